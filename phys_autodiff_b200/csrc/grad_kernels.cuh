// Closed-loop gradient  d(L_sigma + L_u) / d(MLP weights)  -- the "MLP backward fed by the physics VJP" the
// reference plans (REQUIREMENT.md:155-169) but never built: its backward stops at dL/dR
// (src/phys_cpu.cpp:151-170) and its MLP backward is the MSE one (src/mlp_cpu.cpp:38-85).  ADDITIVE: no
// reference counterpart, checked against oracle_grad.c (itself pinned by finite differences).
//
// Inputs are what the forward pass left in HBM: the time-t fields (sigma, u channel-major) and the four
// residual arrays.  Two kernels (grad_kernels.cu):
//   k_phys_adjoint  forms, for every grid point q, the adjoint of the time-t network outputs:
//        g      = (2 w / float(N)) * R                         (fp32, as src/phys_cpu.cpp:162-163)
//        A_t[c] = local terms  g_s div(u), g_s d_j sigma + sum_i g_ui d_j u_i
//               + TRANSPOSED central differences of the fluxes  g_s u_j,  g_ui u_j + delta_ij g_s sigma
//                 taken from the six neighbours (clamped edges: the edge point is its own neighbour and
//                 the 1/(2h) divisor stays, src/phys_cpu.cpp:8-10, so its flux enters with the sign flipped);
//   k_phys_grad     back-propagates A_t and A_+ = +g / (2 dt), A_- = -A_+ (time slices t+dt, t-dt: local) through
//      the two-layer ReLU MLP (lane per PAIR of hidden units, warp per point): the hidden pre-activation is
//      recomputed with the forward's exact fp32 operation order (same ReLU mask and activations as the fields
//      that produced R), everything downstream uses packed FMAs.  A_+ = -A_- lets the two outer slices share
//      W2^T A and the dW2 update  A_+ (a_+ - a_-).
// Weight-gradient accumulators: ten per hidden unit (sum dz*x, dz*y, dz*z | sum dz_-, dz_0, dz_+ | sum A_o a),
// fp32 within a 32-point batch, double across batches (per thread, in shared memory); blocks write double
// partials and the last block to finish sums them in block order (deterministic) and forms
// dW1[h,3] = sum_s t_s sum dz_s,  db1[h] = sum_s sum dz_s.
#pragma once
#include <cuda_runtime.h>


namespace physad {

struct GradArgs {
    int nx, ny, nz;            // global grid
    int z_begin, z_end;        // planes whose points are processed
    int z_origin;              // global z of plane 0 of the arrays (may be negative: wrapped halo planes)
    int wrap_z;                // periodic grid AND the arrays hold the whole grid: z neighbours wrap inside the arrays
    int periodic;
    size_t cstride;            // channel stride of u0 (points in the arrays)
    const float* s0;           // time-t sigma
    const float* u0;           // time-t velocity, channel-major
    const float* R[4];         // residuals (same plane layout as the fields)
    const float* cxs; const float* cys; const float* czs;   // axis coordinate tables (global indices)
    float scale_s, scale_u;    // 2 w / float(N)
    float inv2dt, inv2hx, inv2hy, inv2hz;
    float tc[3];               // network time input of the three slices
    const float *W1, *b1, *W2; // device copies, reference layout
    int H;                     // runtime width (<= template width)
    double* partials;          // [gridDim.x][GRAD_NACC*HT + 6]
    unsigned int* ticket;
    double* grad;              // [9*H + 4] (runtime H): dW1 | db1 | dW2 | db2
    double* acc_out;           // optional [2]: sum R_sigma^2, sum |R_u|^2 over the processed points
};

constexpr int GRAD_THREADS = 256;
constexpr int GRAD_NACC = 10;   // per hidden unit: sum dz*x, dz*y, dz*z | sum dz_-, dz_0, dz_+ | sum A_o a (4)

// grad_kernels.cu (its own translation unit, so the kernel can be rebuilt without the rest of the library)
int grad_blocks_per_sm(int HT, int* out);                                   // resident blocks per SM on the current device
int adjoint_launch(const GradArgs& a, float4* adj, cudaStream_t st);        // A_t of every processed point -> adj
int grad_launch(int HT, const GradArgs& a, const float4* adj, unsigned blocks, cudaStream_t st);   // cudaError_t values

}  // namespace physad
