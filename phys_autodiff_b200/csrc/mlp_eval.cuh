// Strict-fp32 evaluation of the reference's two-layer coordinate MLP, written for sm_100a.
//
// "Strict" = every multiply and every add is rounded separately (__fmul_rn / __fadd_rn, which nvcc
// never contracts into FFMA) and accumulated in the reference's order, so the outputs are
// bit-identical to mlp_forward<ExecCpu> (reference src/mlp_cpu.cpp:14-36):
//     a[h] = relu(b1[h] + W1[h,0]*x + W1[h,1]*y + W1[h,2]*z + W1[h,3]*t)     (left to right)
//     y[o] = b2[o] + sum_{h ascending} W2[o,h]*a[h]
// SURVEY.md section 0 fact 3 explains why this matters: the time difference multiplies output
// noise by 1/(2 dt), so FFMA-contracted outputs break the 1e-5 residual tolerance.
//
// What is shared instead of recomputed (none of it changes a single rounding):
//   * the three time slices t-dt, t, t+dt differ only in the LAST layer-1 term, so the prefix
//     ((b1 + W1[h,0]*x) + W1[h,1]*y) + W1[h,2]*z is computed once and the pre-rounded products
//     W1[h,3]*t_slice (the same for every grid point) are formed once on the host;
//   * a thread's P points share x (and the plane's z), so b1 + W1[h,0]*x and W1[h,2]*z are
//     computed once per thread.
// Weights arrive through the kernel-parameter constant bank (MlpConst is a __grid_constant__
// argument): uniform loads, no shared-memory or LSU traffic in the inner loop.
#pragma once
#include <cuda_runtime.h>

namespace physad {

// Weights as the kernels consume them, filled on the host (capi.cu: fill_const) and passed as a
// __grid_constant__ kernel parameter.  Layer 1 is stored per PAIR of hidden units q = (2q, 2q+1) because
// it is evaluated two hidden units at a time with packed f32x2 instructions; arrays whose values are
// multiplied (w0s, w1s, w2s and the layer-2 pairs) hold their pair half-swapped, see mlp_eval.
template <int H>
struct MlpConst {
    static_assert(H % 2 == 0, "hidden units are processed in pairs");
    float2 b1p[H / 2];  // {b1[2q],   b1[2q+1]}
    float2 w0s[H / 2];  // {W1[2q+1,0], W1[2q,0]}   swapped
    float2 w1s[H / 2];  // {W1[2q+1,1], W1[2q,1]}   swapped
    float2 w2s[H / 2];  // {W1[2q+1,2], W1[2q,2]}   swapped
    float2 ptm[H / 2];  // {W1[2q,3]*t_minus, W1[2q+1,3]*t_minus}  separately rounded fp32 products
    float2 pt0[H / 2];  // same for t
    float2 ptp[H / 2];  // same for t_plus
    float4 w2[H];       // {W2[1,h], W2[0,h], W2[3,h], W2[2,h]}   output pairs, swapped
    float4 b2;          // {b2[0..3]}
};

__device__ __forceinline__ float relu_ref(float s) {
    // reference: s > 0 ? s : 0 (src/mlp_cpu.cpp:7-9).  fmaxf gives the same value for every input
    // (NaN -> 0 as well); only the sign of a zero result may differ.
    return fmaxf(s, 0.f);
}

// Axis coordinate of grid index i (reference src/mlp_grid.cpp:25-29): IEEE division, then 2u-1.
__device__ __forceinline__ float axis_coord(int i, int n, bool m1p1) {
    if (n <= 1) return 0.f;
    const float u = __fdiv_rn(float(i), float(n - 1));
    return m1p1 ? __fsub_rn(__fmul_rn(2.f, u), 1.f) : u;
}

// Neighbour index rule of the stencil (reference src/phys_cpu.cpp:8-15): wrap or clamp.
__device__ __forceinline__ int bc_index(int v, int n, bool periodic) {
    if (periodic) {
        int r = v % n;
        return r < 0 ? r + n : r;
    }
    return v < 0 ? 0 : (v > n - 1 ? n - 1 : v);
}

// ---- packed fp32x2 arithmetic (sm_100a FMUL2 / FADD2) ------------------------------------------
// The FP32 pipe retires 128 lane-ops/clk/SM whether they are issued as scalar or as packed f32x2
// instructions (profiles/r01_microbench_fp32_*.json: strict 36.9 TFLOP/s == strict2 37.1), but a packed
// instruction takes ONE issue slot for two lane-ops.  The scalar kernel is issue-bound (ncu: 95 % issue
// utilisation at 78 % FMA-pipe utilisation), so layer 2 -- 24 of the ~31 pipe ops per point and hidden
// unit -- is issued packed: the two halves are two OUTPUTS (y0,y1 | y2,y3) sharing one activation, which
// the hardware broadcasts from a single register (SASS operand `R.F32`), and the weight pair comes
// straight from a 64-bit uniform load.  Each half is still a separately rounded multiply followed by a
// separately rounded add in the reference's h order, so results stay bit-identical.
// ptxas would contract mul.rn.f32x2 feeding add.rn.f32x2 into FFMA2 (single rounding -- measured, even
// with .rn and -fmad=false) when the product has one use; feeding the product half-swapped (a free
// `.LO_HI` operand swizzle in SASS) prevents that, so weight pairs are stored swapped.
// tests/test_boundary.py asserts the kernels contain no FFMA2.
typedef unsigned long long f32x2;

__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 mul2_rn(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// acc + swap_halves(p), each half rounded to nearest
__device__ __forceinline__ f32x2 add2_rn_swapped(f32x2 acc, f32x2 p) {
    float lo, hi;
    unpack2(p, lo, hi);
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(acc), "l"(pack2(hi, lo)));
    return r;
}

__device__ __forceinline__ f32x2 add2_rn(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 pack2(float2 v) { return pack2(v.x, v.y); }
__device__ __forceinline__ f32x2 bcast2(float v) { return pack2(v, v); }  // becomes a `.F32` broadcast operand

// NS = 3: all three time slices (y[j][0|1|2] = t-dt | t | t+dt);  NS = 1: time t only (y[j][0]).
// P points share cx and cz and differ in cy.
// PACKED = true : layer 1 on two hidden units per instruction, layer 2 on two outputs per instruction.
// PACKED = false: the same arithmetic as scalar FMUL/FADD (kept as the in-tree cross-check; bitwise equal).
template <int H, int NS, int P, int UNROLL, bool PACKED>
__device__ __forceinline__ void mlp_eval(const MlpConst<H>& w, float cx, const float (&cy)[P], float cz,
                                         float (&y)[P][NS][4]) {
    const float4 b2 = w.b2;
    f32x2 q01[P][NS], q23[P][NS];  // packed accumulators {y0,y1}, {y2,y3}
#pragma unroll
    for (int j = 0; j < P; ++j)
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            if (PACKED) {
                q01[j][s] = pack2(b2.x, b2.y);
                q23[j][s] = pack2(b2.z, b2.w);
            } else {
                y[j][s][0] = b2.x; y[j][s][1] = b2.y; y[j][s][2] = b2.z; y[j][s][3] = b2.w;
            }
        }
#pragma unroll UNROLL
    for (int q = 0; q < H / 2; ++q) {
        const float2 b1p = w.b1p[q], w0s = w.w0s[q], w1s = w.w1s[q], w2s = w.w2s[q];
        const float2 tm = w.ptm[q], t0 = w.pt0[q], tp = w.ptp[q];
        const float4 ca = w.w2[2 * q], cb = w.w2[2 * q + 1];
        if (PACKED) {
            // {h, h+1} pre-activations; every product is formed half-swapped and added back swapped
            const f32x2 sx = add2_rn_swapped(pack2(b1p), mul2_rn(pack2(w0s), bcast2(cx)));  // b1 + W1[.,0]*x
            const f32x2 mz = mul2_rn(pack2(w2s), bcast2(cz));                               // W1[.,2]*z (swapped)
            const f32x2 ca10 = pack2(ca.x, ca.y), ca32 = pack2(ca.z, ca.w);
            const f32x2 cb10 = pack2(cb.x, cb.y), cb32 = pack2(cb.z, cb.w);
#pragma unroll
            for (int j = 0; j < P; ++j) {
                const f32x2 sxy = add2_rn_swapped(sx, mul2_rn(pack2(w1s), bcast2(cy[j])));
                const f32x2 sxyz = add2_rn_swapped(sxy, mz);
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    const float2 pt = (NS == 1) ? t0 : (s == 0 ? tm : (s == 1 ? t0 : tp));
                    float v0, v1;
                    unpack2(add2_rn(sxyz, pack2(pt)), v0, v1);
                    const f32x2 a0 = bcast2(relu_ref(v0)), a1 = bcast2(relu_ref(v1));
                    q01[j][s] = add2_rn_swapped(q01[j][s], mul2_rn(a0, ca10));   // h = 2q
                    q23[j][s] = add2_rn_swapped(q23[j][s], mul2_rn(a0, ca32));
                    q01[j][s] = add2_rn_swapped(q01[j][s], mul2_rn(a1, cb10));   // h = 2q + 1
                    q23[j][s] = add2_rn_swapped(q23[j][s], mul2_rn(a1, cb32));
                }
            }
        } else {
#pragma unroll
            for (int e = 0; e < 2; ++e) {  // h = 2q + e
                const float b1 = e ? b1p.y : b1p.x, w0 = e ? w0s.x : w0s.y, w1 = e ? w1s.x : w1s.y, w2 = e ? w2s.x : w2s.y;
                const float4 c = e ? cb : ca;
                const float sx = __fadd_rn(b1, __fmul_rn(w0, cx));
                const float mz = __fmul_rn(w2, cz);
#pragma unroll
                for (int j = 0; j < P; ++j) {
                    const float sxyz = __fadd_rn(__fadd_rn(sx, __fmul_rn(w1, cy[j])), mz);
#pragma unroll
                    for (int s = 0; s < NS; ++s) {
                        const float2 pt2 = (NS == 1) ? t0 : (s == 0 ? tm : (s == 1 ? t0 : tp));
                        const float act = relu_ref(__fadd_rn(sxyz, e ? pt2.y : pt2.x));
                        y[j][s][0] = __fadd_rn(y[j][s][0], __fmul_rn(c.y, act));
                        y[j][s][1] = __fadd_rn(y[j][s][1], __fmul_rn(c.x, act));
                        y[j][s][2] = __fadd_rn(y[j][s][2], __fmul_rn(c.w, act));
                        y[j][s][3] = __fadd_rn(y[j][s][3], __fmul_rn(c.z, act));
                    }
                }
            }
        }
    }
    if (PACKED) {
#pragma unroll
        for (int j = 0; j < P; ++j)
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                unpack2(q01[j][s], y[j][s][0], y[j][s][1]);
                unpack2(q23[j][s], y[j][s][2], y[j][s][3]);
            }
    }
}

// ---------------------------------------------------------------------------------------------
// Point residual (reference src/phys_cpu.cpp:66-109), every rounding spelled out so the fused kernel and
// the stage-wise kernel produce identical bits from identical fields.
//   f = [sigma, ux, uy, uz] at the point;  g?[c] = d f_c / d?;  dT[c] = d f_c / dt.
// Two arithmetic modes, selected by the type T of the derivatives:
//   T = double : EXACTLY the CPU reference -- float loads widened to double, differences times
//                1.0/(2.0*double(h)) (:38-41), sums in the order written on :96-106, separate multiply and
//                add (the reference build does not contract), one rounding to float at the end.  Residuals
//                are then bit-identical to cpu_phys_residuals.  Runs on the FP64 / conversion pipes, which
//                the fp32-bound kernels leave idle.
//   T = float  : the same expressions in fp32 with fused multiply-adds (what the reference's own CUDA
//                kernels do, src/phys_cuda_fused.cu:67-99); within ~1e-7 * max|R| of the CPU.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float central_diff(float plus, float minus, float inv2h) {
    return __fmul_rn(__fsub_rn(plus, minus), inv2h);
}
__device__ __forceinline__ double central_diff(float plus, float minus, double inv2h) {
    return __dmul_rn(__dsub_rn(double(plus), double(minus)), inv2h);
}

__device__ __forceinline__ void point_residual(const float (&f)[4], const float (&gx)[4], const float (&gy)[4],
                                               const float (&gz)[4], const float (&dT)[4], float (&R)[4]) {
    const float div = __fadd_rn(__fadd_rn(gx[1], gy[2]), gz[3]);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float adv = __fmaf_rn(f[3], gz[c], __fmaf_rn(f[2], gy[c], __fmul_rn(f[1], gx[c])));
        R[c] = __fadd_rn(dT[c], adv);
    }
    R[0] = __fmaf_rn(f[0], div, R[0]);
}

__device__ __forceinline__ void point_residual(const float (&f)[4], const double (&gx)[4], const double (&gy)[4],
                                               const double (&gz)[4], const double (&dT)[4], float (&R)[4]) {
    const double ux = double(f[1]), uy = double(f[2]), uz = double(f[3]);
    const double div = __dadd_rn(__dadd_rn(gx[1], gy[2]), gz[3]);                                   // :96
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        // u.grad(f_c) = (ux*d/dx + uy*d/dy) + uz*d/dz                                                  :97-101
        const double adv = __dadd_rn(__dadd_rn(__dmul_rn(ux, gx[c]), __dmul_rn(uy, gy[c])), __dmul_rn(uz, gz[c]));
        const double r = __dadd_rn(dT[c], adv);
        R[c] = float(c == 0 ? __dadd_rn(r, __dmul_rn(double(f[0]), div)) : r);                      // :103-106
    }
}

// First-order UPWIND advection (additive switch; the reference plans it, REQUIREMENT.md:123-134, and never ships it:
// parity is against oracle.c's oracle_phys_residuals_upwind only).  As point_residual, except that the advective
// derivative of f_c along axis j is one-sided against the velocity: (f - f_minus)/h if u_j > 0, (f_plus - f)/h otherwise;
// the divergence in sigma * div(u) and the time derivative stay central.  T = float (default) or double (exact mode).
template <typename T>
__device__ __forceinline__ void point_residual_upwind(const float (&f)[4], const float (&xm)[4], const float (&xp)[4],
                                                      const float (&ym)[4], const float (&yp)[4], const float (&zm)[4],
                                                      const float (&zp)[4], const T (&dT)[4], T i1x, T i1y, T i1z, T i2x, T i2y,
                                                      T i2z, float (&R)[4]) {
    const T div = ((T(xp[1]) - T(xm[1])) * i2x + (T(yp[2]) - T(ym[2])) * i2y) + (T(zp[3]) - T(zm[3])) * i2z;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const T ax = f[1] > 0.f ? (T(f[c]) - T(xm[c])) * i1x : (T(xp[c]) - T(f[c])) * i1x;
        const T ay = f[2] > 0.f ? (T(f[c]) - T(ym[c])) * i1y : (T(yp[c]) - T(f[c])) * i1y;
        const T az = f[3] > 0.f ? (T(f[c]) - T(zm[c])) * i1z : (T(zp[c]) - T(f[c])) * i1z;
        T r = dT[c] + ((T(f[1]) * ax + T(f[2]) * ay) + T(f[3]) * az);
        if (c == 0) r += T(f[0]) * div;
        R[c] = float(r);
    }
}

}  // namespace physad
