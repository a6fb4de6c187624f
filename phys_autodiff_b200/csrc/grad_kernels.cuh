// Closed-loop gradient  d(L_sigma + L_u) / d(MLP weights)  -- the "MLP backward fed by the physics VJP" the
// reference plans (REQUIREMENT.md:155-169) but never built: its backward stops at dL/dR
// (src/phys_cpu.cpp:151-170) and its MLP backward is the MSE one (src/mlp_cpu.cpp:38-85).  ADDITIVE: no
// reference counterpart, checked against oracle_grad.c (itself pinned by finite differences).
//
// Inputs are what the forward pass left in HBM: the time-t fields (sigma, u channel-major) and the four
// residual arrays.  For every grid point q the kernel
//   A. forms the adjoint of the three network evaluations at q (thread per point):
//        g      = (2 w / float(N)) * R                         (fp32, as src/phys_cpu.cpp:162-163)
//        A_+    = +g / (2 dt),  A_- = -g / (2 dt)               (time slices t+dt, t-dt: local)
//        A_t[c] = local terms  g_s div(u), g_s d_j sigma + sum_i g_ui d_j u_i
//               + TRANSPOSED central differences of the fluxes  g_s u_j,  g_ui u_j + delta_ij g_s sigma
//                 taken from the six neighbours (clamped edges: the edge point is its own neighbour and
//                 the 1/(2h) divisor stays, src/phys_cpu.cpp:8-10, so its flux enters with the sign flipped);
//   B. back-propagates A through the two-layer ReLU MLP (lane per hidden unit, warp per point): the hidden
//      pre-activation is recomputed with the forward's exact fp32 operation order (same ReLU mask and
//      activations as the fields that produced R), everything downstream uses FMAs.  A_+ = -A_- lets the
//      two outer slices share W2^T A and the dW2 update  A_d (a_+ - a_-).
// Weight-gradient accumulators live in registers (9 per hidden unit: dW1[h,0..3], db1[h], dW2[0..3,h]),
// fp32 within a 32-point batch, double across batches; blocks write double partials and the last block
// to finish sums them in block order (deterministic).
#pragma once
#include <cuda_runtime.h>

#include "fused_loss.cuh"

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
    double* partials;          // [gridDim.x][9*HT + 4]
    unsigned int* ticket;
    double* grad;              // [9*H + 4] (runtime H): dW1 | db1 | dW2 | db2
};

constexpr int GRAD_THREADS = 256;

template <int HT>
__global__ void __launch_bounds__(GRAD_THREADS) k_phys_grad(const GradArgs a) {
    constexpr int HPL = HT / 32;             // hidden units per lane
    constexpr int NW = GRAD_THREADS / 32;
    constexpr int NGT = 9 * HT + 4;
    __shared__ float4 s_x[2][GRAD_THREADS];  // cx, cy, cz, -
    __shared__ float4 s_gt[2][GRAD_THREADS]; // A_t
    __shared__ float4 s_gd[2][GRAD_THREADS]; // A_+ (= -A_-)
    __shared__ unsigned int s_flag;

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const size_t plane = size_t(a.nx) * a.ny;
    const size_t p_begin = size_t(a.z_begin) * plane, p_end = size_t(a.z_end) * plane;
    const size_t nchunks = (p_end - p_begin + GRAD_THREADS - 1) / GRAD_THREADS;

    // this lane's hidden units h = lane + 32 j
    float w1[HPL][4], bb[HPL], pt[HPL][3], w2[HPL][4];
#pragma unroll
    for (int j = 0; j < HPL; ++j) {
        const int h = lane + 32 * j;
        const bool on = h < a.H;
#pragma unroll
        for (int k = 0; k < 4; ++k) w1[j][k] = on ? __ldg(a.W1 + h * 4 + k) : 0.f;
        bb[j] = on ? __ldg(a.b1 + h) : 0.f;
#pragma unroll
        for (int s = 0; s < 3; ++s) pt[j][s] = __fmul_rn(w1[j][3], a.tc[s]);
#pragma unroll
        for (int o = 0; o < 4; ++o) w2[j][o] = on ? __ldg(a.W2 + o * a.H + h) : 0.f;
    }
    double acc[HPL][9];
#pragma unroll
    for (int j = 0; j < HPL; ++j)
#pragma unroll
        for (int k = 0; k < 9; ++k) acc[j][k] = 0.0;
    double db2[4] = {0.0, 0.0, 0.0, 0.0};

    int buf = 0;
    for (size_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x, buf ^= 1) {
        // ---- A: adjoint of the three network outputs at this thread's point -----------------------
        {
            const size_t p = p_begin + ch * GRAD_THREADS + threadIdx.x;
            float4 gt = make_float4(0.f, 0.f, 0.f, 0.f), gd = gt, xc = gt;
            if (p < p_end) {
                const int z = int(p / plane);
                const int rem = int(p - size_t(z) * plane);
                const int y = rem / a.nx, x = rem - y * a.nx;
                const bool per = a.periodic != 0;
                // neighbour indices and the sign with which a neighbour's flux enters (clamped edges flip it)
                const int xm = bc_index(x - 1, a.nx, per), xp = bc_index(x + 1, a.nx, per);
                const int ym = bc_index(y - 1, a.ny, per), yp = bc_index(y + 1, a.ny, per);
                int zm = z - 1, zp = z + 1;
                if (a.wrap_z) { zm = bc_index(zm, a.nz, true); zp = bc_index(zp, a.nz, true); }
                else if (!per) { zm = max(zm, 0); zp = min(zp, a.nz - 1); }
                const float sg[3][2] = {{(per || x >= 1) ? 1.f : -1.f, (per || x <= a.nx - 2) ? 1.f : -1.f},
                                        {(per || y >= 1) ? 1.f : -1.f, (per || y <= a.ny - 2) ? 1.f : -1.f},
                                        {(per || z >= 1) ? 1.f : -1.f, (per || z <= a.nz - 2) ? 1.f : -1.f}};
                const size_t zl = size_t(z - a.z_origin) * plane;
                const size_t row = zl + size_t(y) * a.nx;
                const size_t q = row + x;
                const size_t nbr[3][2] = {{row + xm, row + xp},
                                          {zl + size_t(ym) * a.nx + x, zl + size_t(yp) * a.nx + x},
                                          {size_t(zm - a.z_origin) * plane + size_t(y) * a.nx + x,
                                           size_t(zp - a.z_origin) * plane + size_t(y) * a.nx + x}};
                const float i2h[3] = {a.inv2hx, a.inv2hy, a.inv2hz};
                const float gq[4] = {a.scale_s * __ldg(a.R[0] + q), a.scale_u * __ldg(a.R[1] + q),
                                     a.scale_u * __ldg(a.R[2] + q), a.scale_u * __ldg(a.R[3] + q)};
                float A[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    float f[2][4], g[2][4];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const size_t n = nbr[j][e];
                        f[e][0] = __ldg(a.s0 + n);
                        g[e][0] = a.scale_s * __ldg(a.R[0] + n);
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            f[e][c + 1] = __ldg(a.u0 + size_t(c) * a.cstride + n);
                            g[e][c + 1] = a.scale_u * __ldg(a.R[c + 1] + n);
                        }
                    }
                    // local terms: derivatives of the four fields along j at q
                    float d[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) d[c] = (f[1][c] - f[0][c]) * i2h[j];
                    A[0] += gq[0] * d[j + 1];
                    A[j + 1] += gq[0] * d[0] + gq[1] * d[1] + gq[2] * d[2] + gq[3] * d[3];
                    // transposed difference of the neighbours' fluxes along j
                    float T[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        float fl[2];
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            fl[e] = g[e][c] * f[e][j + 1];
                            if (c == j + 1) fl[e] += g[e][0] * f[e][0];
                        }
                        T[c] = sg[j][0] * fl[0] - sg[j][1] * fl[1];
                        A[c] += i2h[j] * T[c];
                    }
                }
                gt = make_float4(A[0], A[1], A[2], A[3]);
                gd = make_float4(gq[0] * a.inv2dt, gq[1] * a.inv2dt, gq[2] * a.inv2dt, gq[3] * a.inv2dt);
                const int zg = a.wrap_z ? z : bc_index(z, a.nz, true);   // halo planes of a periodic slab carry wrapped coordinates
                xc = make_float4(__ldg(a.cxs + x), __ldg(a.cys + y), __ldg(a.czs + zg), 0.f);
                db2[0] += double(A[0]); db2[1] += double(A[1]); db2[2] += double(A[2]); db2[3] += double(A[3]);
            }
            s_x[buf][threadIdx.x] = xc;
            s_gt[buf][threadIdx.x] = gt;
            s_gd[buf][threadIdx.x] = gd;
        }
        __syncthreads();
        // ---- B: MLP backward, warp `wid` takes 32 of the chunk's points, lane = hidden unit(s) -----
        float f[HPL][9];
#pragma unroll
        for (int j = 0; j < HPL; ++j)
#pragma unroll
            for (int k = 0; k < 9; ++k) f[j][k] = 0.f;
#pragma unroll 2
        for (int i = 0; i < 32; ++i) {
            const float4 xc = s_x[buf][wid * 32 + i];
            const float4 gt = s_gt[buf][wid * 32 + i];
            const float4 gd = s_gd[buf][wid * 32 + i];
#pragma unroll
            for (int j = 0; j < HPL; ++j) {
                // forward's operation order: ((b1 + W1[h,0] x) + W1[h,1] y) + W1[h,2] z, then + W1[h,3] t_s
                float pre = __fadd_rn(bb[j], __fmul_rn(w1[j][0], xc.x));
                pre = __fadd_rn(pre, __fmul_rn(w1[j][1], xc.y));
                pre = __fadd_rn(pre, __fmul_rn(w1[j][2], xc.z));
                const float zm = __fadd_rn(pre, pt[j][0]), z0 = __fadd_rn(pre, pt[j][1]), zp = __fadd_rn(pre, pt[j][2]);
                const float am = fmaxf(zm, 0.f), a0 = fmaxf(z0, 0.f), ap = fmaxf(zp, 0.f);
                const float da_t = w2[j][0] * gt.x + w2[j][1] * gt.y + w2[j][2] * gt.z + w2[j][3] * gt.w;
                const float da_d = w2[j][0] * gd.x + w2[j][1] * gd.y + w2[j][2] * gd.z + w2[j][3] * gd.w;
                const float ad = ap - am;
                f[j][5] += gt.x * a0 + gd.x * ad;
                f[j][6] += gt.y * a0 + gd.y * ad;
                f[j][7] += gt.z * a0 + gd.z * ad;
                f[j][8] += gt.w * a0 + gd.w * ad;
                const float dz0 = z0 > 0.f ? da_t : 0.f;
                const float dzp = zp > 0.f ? da_d : 0.f;
                const float dzm = zm > 0.f ? -da_d : 0.f;
                const float dzs = dz0 + dzp + dzm;
                f[j][0] += dzs * xc.x;
                f[j][1] += dzs * xc.y;
                f[j][2] += dzs * xc.z;
                f[j][3] += dzm * a.tc[0] + dz0 * a.tc[1] + dzp * a.tc[2];
                f[j][4] += dzs;
            }
        }
#pragma unroll
        for (int j = 0; j < HPL; ++j)
#pragma unroll
            for (int k = 0; k < 9; ++k) acc[j][k] += double(f[j][k]);
    }

    // ---- block partial: sum the warps' accumulators through shared memory (reusing the staging arrays) ----
    __syncthreads();
    double* s_acc = reinterpret_cast<double*>(&s_x[0][0]);   // needs NW * 32 doubles per pass: 2 KB of the 8 KB
    double* part = a.partials + size_t(blockIdx.x) * NGT;
#pragma unroll
    for (int j = 0; j < HPL; ++j)
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            s_acc[wid * 32 + lane] = acc[j][k];
            __syncthreads();
            if (wid == 0) {
                double s = 0.0;
#pragma unroll
                for (int w = 0; w < NW; ++w) s += s_acc[w * 32 + lane];
                const int h = lane + 32 * j;
                // template-width layout: dW1[h*4+k] | db1[h] at 4 HT | dW2[o*HT+h] at 5 HT | db2 at 9 HT
                const int idx = k < 4 ? h * 4 + k : (k == 4 ? 4 * HT + h : 5 * HT + (k - 5) * HT + h);
                part[idx] = s;
            }
            __syncthreads();
        }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        const double v = warp_sum(db2[o]);
        if (lane == 0) s_acc[wid * 4 + o] = v;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) s += s_acc[w * 4 + threadIdx.x];
        part[9 * HT + threadIdx.x] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_flag = atomicAdd(a.ticket, 1u);
    __syncthreads();
    if (s_flag != gridDim.x - 1) return;
    __threadfence();
    // last block: block-ordered sums, written in the runtime-width layout
    for (int e = threadIdx.x; e < NGT; e += GRAD_THREADS) {
        double s = 0.0;
        for (unsigned int b = 0; b < gridDim.x; ++b) s += __ldcg(a.partials + size_t(b) * NGT + e);
        int out = -1;
        if (e < 4 * HT) { if (e / 4 < a.H) out = e; }
        else if (e < 5 * HT) { if (e - 4 * HT < a.H) out = 4 * a.H + (e - 4 * HT); }
        else if (e < 9 * HT) { const int o = (e - 5 * HT) / HT, h = (e - 5 * HT) % HT; if (h < a.H) out = 5 * a.H + o * a.H + h; }
        else out = 9 * a.H + (e - 9 * HT);
        if (out >= 0) a.grad[out] = s;
    }
    if (threadIdx.x == 0) *a.ticket = 0u;
}

}  // namespace physad
