// Strict-fp32 dense contraction for the generic-shape MLP operator and its MSE backward
// (mlp_forward<ExecCuda> for In/Out != 4 and mlp_backward<ExecCuda>: include/mlp.h:5-9, CPU reference
// src/mlp_cpu.cpp:14-85).  Interface of dense_kernels.cu.
//
//   C[m, n] = epilogue( init[n] + sum over k ASCENDING of A(m, k) * B(k, n) )      every multiply and every add rounded
//                                                                                  separately, in that order
// which is exactly what each of the reference's loops computes for one output entry -- bit-exactness with the CPU
// path fixes the ORDER of an entry's accumulation, not where the operands come from.  So instead of one thread per
// entry walking global memory (the shape of the reference's own kernels, src/mlp_cuda.cu:14-89), this is a
// register-tiled contraction: a 256-thread block owns a 64 x 64 tile of C, stages 16-deep slabs of A and B in shared
// memory (coalesced along whichever index is contiguous in memory), and a thread accumulates a 4 x 4 sub-tile in
// registers with k ascending -- 32 strict operations per two 128-bit shared-memory reads.
#pragma once
#include <cuda_runtime.h>

namespace physad {

enum GemmEpilogue {
    GEMM_STORE = 0,        // C = acc
    GEMM_RELU = 1,         // C = acc > 0 ? acc : 0                                   (hidden layer, src/mlp_cpu.cpp:7-9)
    GEMM_SCALED_DIFF = 2,  // C = scale * (acc - aux[m, n])                           (gz2, src/mlp_cpu.cpp:58)
    GEMM_MASK = 3          // C = acc * (aux[m, n] > 0 ? 1 : 0)                       (gz1 = s * relu'(z1), src/mlp_cpu.cpp:74)
};

struct GemmArgs {
    const float* A; long long a_ms, a_ks;   // A(m, k) = A[m * a_ms + k * a_ks]
    const float* B; long long b_ks, b_ns;   // B(k, n) = B[k * b_ks + n * b_ns]
    int ones_col;                           // column n whose B(k, n) is 1 for every k (column sums: db), or -1
    const float* init;                      // [N] start value of every sum (the bias), or null = 0
    int M, N, K;
    int epilogue;
    const float* aux;                       // [M x N] row-major, epilogues 2 and 3
    float scale;
    float* C; long long c_ms;               // C[m * c_ms + n] for n < n_split
    int n_split;                            // columns >= n_split go to C2[m] (the ones column)
    float* C2;
};

int strict_gemm_launch(const GemmArgs& g, cudaStream_t st);   // returns a cudaError_t value

}  // namespace physad
