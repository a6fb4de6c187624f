// Analytic ("tangent") physics loss -- BASELINE.json's north_star in its literal wording: evaluate the coordinate MLP
// over the grid, PROPAGATE THE INPUT-DERIVATIVES THROUGH IT (forward mode, tangents in registers) and reduce the PDE
// residual loss.  ADDITIVE and explicitly NOT the parity path: the reference never does this -- every derivative in
// src/phys_cpu.cpp:66-109 is a central finite difference of MLP outputs sampled on the grid,
// with GridSpec.hx independent of the coordinate normalisation, ReLU kinks and periodic wrap of a non-periodic network --
// so its numbers differ from the finite-difference loss by the discretisation error (they agree where the network is
// affine over the stencil; tests/test_oracle_cpu.py).  Checker: oracle.c: oracle_tangent_loss ("parity unpinned").
//
//   z_h = b1[h] + sum_k W1[h,k] c_k           (strict fp32, the forward path's roundings -> the ReLU mask m_h = z_h > 0
//                                               is bit-identical to the reference forward's)
//   y_o = b2[o] + sum_h W2[o,h] relu(z_h)     dy_o/dc_k = sum_h m_h W2[o,h] W1[h,k]          (products precomputed on the host)
//   d/dx_j = s_j d/dc_j,  s_j = dc_j/dx_j from the grid spacing and the normalisation;  d/dt = d/dc_t
//   R_sigma = d_t sigma + u . grad(sigma) + sigma div(u),   R_u = d_t u + (u . grad) u       (the PDE of src/phys_cpu.cpp:103-106)
// No stencil, hence no halo, no ring, no z-march and no second or third network evaluation: one pass of 25 multiply-adds
// per point and hidden unit (FFMA allowed: there is no 1/(2 dt) amplification of rounding noise), points split over the
// ranks with nothing but the 16-byte reduction in common.
#pragma once
#include <cuda_runtime.h>

namespace physad {

template <int H>
struct TangentConst {
    float4 w1[H];      // {W1[h,0], W1[h,1], W1[h,2], fl(W1[h,3] * c_t)}   the last one pre-rounded like MlpConst::pt0
    float b1[H];
    float4 w2[H];      // {W2[0,h], W2[1,h], W2[2,h], W2[3,h]}
    float4 p[H][4];    // p[h][o] = {W2[o,h] W1[h,0], W2[o,h] W1[h,1], W2[o,h] W1[h,2], W2[o,h] W1[h,3]}
    float4 b2;
};

struct TangentArgs {
    int nx, ny, nz, z_begin, z_end;
    const float* cxs; const float* cys; const float* czs;
    float sx, sy, sz;            // dc/dx of the three space coordinates
    double2* partials; unsigned int* ticket; double* acc_out;
    float* R[4];                 // slab-local residual outputs or null
};

// mlp_const: host TangentConst<H> for the template width H in {32, 64, 128}; returns a cudaError_t value
int tangent_launch(int H, const void* tangent_const, const TangentArgs& a, int blocks, cudaStream_t st);
constexpr int TANGENT_THREADS = 256;

}  // namespace physad
