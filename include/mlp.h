// Two-layer ReLU MLP operator, tag-dispatched on the backend (API-compatible with the
// reference's include/mlp.h:5-9; only explicit specialisations exist).
//
//   y[i, o] = b2[o] + sum_h W2[o, h] * relu(b1[h] + sum_k W1[h, k] * x[i, k])
//
// x: B x In, y: B x Out (row-major, HOST pointers, caller-owned); W1: H x In, W2: Out x H row-major.
// This library provides mlp_forward<ExecCuda> (bit-exact with the reference's
// mlp_forward<ExecCpu>, src/mlp_cpu.cpp:14-36: fp32, separate multiply and add, k/h ascending).
// mlp_backward<ExecCuda> (MSE weight gradients of mean((y - y_target)^2), src/mlp_cuda.cu:123-184) is
// provided bit-exact with mlp_backward<ExecCpu> as well; it is outside the grid->loss hot path and not tuned.
#ifndef PHYS_AUTODIFF_MLP_H
#define PHYS_AUTODIFF_MLP_H
#include <cstddef>

template <typename Exec>
void mlp_forward(const float* x, const float* W1, const float* b1, const float* W2, const float* b2, float* y,
                 std::size_t B, std::size_t In, std::size_t H, std::size_t Out);

template <typename Exec>
void mlp_backward(const float* x, const float* y_target, const float* W1, const float* b1, const float* W2,
                  const float* b2, float* dW1, float* db1, float* dW2, float* db2,
                  std::size_t B, std::size_t In, std::size_t H, std::size_t Out);

#endif  // PHYS_AUTODIFF_MLP_H
