// Deeper coordinate MLPs, FAST mode: the hidden -> hidden contractions on the 5th-generation tensor cores
// (tcgen05.mma, accumulators and activations in tensor memory), everything else as in deep_kernels.cu.
//
// ADDITIVE and NOT bit-exact: BASELINE config 5 asks where tensor cores start to pay ("tensor-core crossover"), and
// the north-star allows them "if ncu shows that layer width makes [the hidden-layer contractions] compute-bound" -- ncu
// does (profiles/r02_ncu_deep_h128_summary.json: FMA pipe 92 % active).  Plain bf16 or tf32 operands are useless for
// this path (the time difference amplifies output noise by 1/(2 dt) = 250, residual error 0.3 .. 2),
// so every fp32 operand is split into THREE bf16 terms (v = t1 + t2 + t3 to 2^-24) and a layer is the six products
// t_i(a) t_j(W), i + j <= 4, accumulated in fp32 by the tensor core, smallest terms first: the error class of an
// FFMA-contracted fp32 evaluation -- the reference's own CUDA kernels (src/mlp_cuda.cu) -- at 6/16 of the bf16 rate.
// Layer 1 keeps the strict arithmetic (identical bits to the strict path), the output layer is fp32 FMA on the CUDA cores.
// The reference has no hidden -> hidden layer at all (include/mlp.h:5-6), so there is nothing to pin this against except
// this repository's own strict deep kernel: tests bound max |y_fast - y_strict| and the loss difference.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

#include "deep_kernels.cuh"

namespace physad {

// Bytes of the operand image of ONE hidden -> hidden layer: 3 bf16 terms x H x H, each term in the K-major no-swizzle
// core-matrix layout the MMA descriptor names (deep_tc_pack_layer writes it).
inline size_t deep_tc_layer_bytes(int H) { return size_t(3) * H * H * 2; }
// W: the layer's [H out][H in] row-major fp32 weights (the reference's W2-style layout).  Host function.
void deep_tc_pack_layer(int H, const float* W, uint8_t* image);
// Whether (H, hidden_layers) is built: H in {32, 64, 128}, 2 <= hidden_layers <= 16.  All layer images stay resident in
// shared memory when they fit (H = 128: <= 3 hidden layers, H = 64: <= 10); otherwise two buffers are refilled from L2 one
// layer ahead (96 KB per layer and 128 rows at H = 128 -- that mode is L2-bandwidth-bound, not tensor-pipe-bound).
bool deep_tc_supported(int H, int hidden_layers);
// `a.wh` is ignored; `wparts` = device pointer to (hidden_layers - 1) consecutive layer images; a.bh as in deep_launch.
int deep_tc_launch(int H, bool fields, const void* mlp_const, const DeepArgs& a, const uint8_t* wparts, int grid_blocks, cudaStream_t st);

}  // namespace physad
