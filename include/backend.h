// Execution-backend tags of the phys-autodiff operator API.
//
// API-compatible with the reference's include/backend.h:3-4: two empty tag types used as the
// template argument of mlp_forward<> / mlp_backward<> (include/mlp.h).  In this build ExecCuda
// means "the sm_100a kernels in phys_autodiff_b200/csrc" and there is no CPU implementation behind
// ExecCpu: the CPU path lives only in the reference (and in oracle/, as a test checker).
#ifndef PHYS_AUTODIFF_BACKEND_H
#define PHYS_AUTODIFF_BACKEND_H

struct ExecCpu {};   // reference CPU path (not provided by this library)
struct ExecCuda {};  // B200 path

#endif  // PHYS_AUTODIFF_BACKEND_H
