// The two backward kernels of the closed loop; see grad_kernels.cuh for what they compute.
#include "grad_kernels.cuh"
#include "mlp_eval.cuh"

#include <cstdint>
#include <type_traits>

namespace physad {

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ f32x2 fma2_rn(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2 sub2_rn(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// ---- A: adjoint of the time-t network outputs, A_t, for QV consecutive x per thread ------------------------
// QV = 4: 128-bit loads of the centre row and of the four y/z neighbour rows of all eight input arrays, two
// scalar loads per array for the x neighbours outside the quad (the access pattern of k_phys_residual_v4);
// QV = 1 is the any-shape fallback.  Output: adj[(z - z_begin) * nx*ny + y*nx + x] = A_t as float4.
template <int QV>
__device__ __forceinline__ void load_q(const float* p, float (&v)[QV]) {
    if constexpr (QV == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
        v[0] = __ldg(p);
    }
}

constexpr int ADJ_THREADS = 128;

template <int QV>
__global__ void __launch_bounds__(ADJ_THREADS, QV == 4 ? 5 : 8) k_phys_adjoint(const GradArgs a, float4* __restrict__ adj) {
    const int nxq = a.nx / QV;
    const size_t nquads = size_t(nxq) * a.ny * (a.z_end - a.z_begin);
    const size_t t = size_t(blockIdx.x) * ADJ_THREADS + threadIdx.x;
    if (t >= nquads) return;
    const int x = int(t % nxq) * QV;
    const size_t r = t / nxq;
    const int y = int(r % a.ny), z = a.z_begin + int(r / a.ny);
    const bool per = a.periodic != 0;
    const size_t plane = size_t(a.nx) * a.ny;
    const int ym = bc_index(y - 1, a.ny, per), yp = bc_index(y + 1, a.ny, per);
    int zm = z - 1, zp = z + 1;
    if (a.wrap_z) { zm = bc_index(zm, a.nz, true); zp = bc_index(zp, a.nz, true); }
    else if (!per) { zm = max(zm, 0); zp = min(zp, a.nz - 1); }
    const size_t zl = size_t(z - a.z_origin) * plane;
    const size_t oc = zl + size_t(y) * a.nx + x;
    // offsets of the neighbour rows (same x): y-1, y+1, z-1, z+1
    const size_t orow[4] = {zl + size_t(ym) * a.nx + x, zl + size_t(yp) * a.nx + x,
                            size_t(zm - a.z_origin) * plane + size_t(y) * a.nx + x,
                            size_t(zp - a.z_origin) * plane + size_t(y) * a.nx + x};
    const size_t oxl = oc - x + bc_index(x - 1, a.nx, per), oxr = oc - x + bc_index(x + QV, a.nx, per);
    const float* arr[8] = {a.s0, a.u0, a.u0 + a.cstride, a.u0 + 2 * a.cstride, a.R[0], a.R[1], a.R[2], a.R[3]};
    const float sc[4] = {a.scale_s, a.scale_u, a.scale_u, a.scale_u};
    // centre row (fields f, loss adjoint g = scale * R) and the x neighbours outside the quad
    float fc[4][QV], gc[4][QV], fl[4], gl[4], fr[4], gr[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        load_q<QV>(arr[c] + oc, fc[c]);
        load_q<QV>(arr[4 + c] + oc, gc[c]);
#pragma unroll
        for (int e = 0; e < QV; ++e) gc[c][e] *= sc[c];
        fl[c] = __ldg(arr[c] + oxl); gl[c] = sc[c] * __ldg(arr[4 + c] + oxl);
        fr[c] = __ldg(arr[c] + oxr); gr[c] = sc[c] * __ldg(arr[4 + c] + oxr);
    }
    float A[QV][4];
#pragma unroll
    for (int e = 0; e < QV; ++e)
#pragma unroll
        for (int c = 0; c < 4; ++c) A[e][c] = 0.f;
    // one axis j at point e: neighbour values (fm, gm) on the minus side, (fp, gp) on the plus side; sm / sp = the sign
    // with which that neighbour's flux enters (a clamped edge point is its own neighbour: sign flipped)
    auto axis = [&](int j, int e, const float (&fm)[4], const float (&gm)[4], const float (&fp)[4], const float (&gp)[4],
                    float sm, float sp, float i2h) {
        float d[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) d[c] = (fp[c] - fm[c]) * i2h;
        A[e][0] += gc[0][e] * d[j + 1];                                                        // g_s div(u)
        A[e][j + 1] += gc[0][e] * d[0] + gc[1][e] * d[1] + gc[2][e] * d[2] + gc[3][e] * d[3];  // g_s d_j sigma + g_ui d_j u_i
#pragma unroll
        for (int c = 0; c < 4; ++c) {   // transposed difference of the fluxes g_c u_j (+ g_s sigma on the diagonal)
            float flm = gm[c] * fm[j + 1], flp = gp[c] * fp[j + 1];
            if (c == j + 1) { flm += gm[0] * fm[0]; flp += gp[0] * fp[0]; }
            A[e][c] += i2h * (sm * flm - sp * flp);
        }
    };
#pragma unroll
    for (int e = 0; e < QV; ++e) {   // x axis: neighbours inside the quad come from the centre row
        float fm[4], gm[4], fp[4], gp[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            fm[c] = e == 0 ? fl[c] : fc[c][e > 0 ? e - 1 : 0];
            gm[c] = e == 0 ? gl[c] : gc[c][e > 0 ? e - 1 : 0];
            fp[c] = e == QV - 1 ? fr[c] : fc[c][e < QV - 1 ? e + 1 : 0];
            gp[c] = e == QV - 1 ? gr[c] : gc[c][e < QV - 1 ? e + 1 : 0];
        }
        const int xe = x + e;
        axis(0, e, fm, gm, fp, gp, (per || xe >= 1) ? 1.f : -1.f, (per || xe <= a.nx - 2) ? 1.f : -1.f, a.inv2hx);
    }
#pragma unroll
    for (int j = 1; j < 3; ++j) {    // y and z axes
        float fm[4][QV], gm[4][QV], fp[4][QV], gp[4][QV];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            load_q<QV>(arr[c] + orow[2 * (j - 1)], fm[c]);
            load_q<QV>(arr[4 + c] + orow[2 * (j - 1)], gm[c]);
            load_q<QV>(arr[c] + orow[2 * (j - 1) + 1], fp[c]);
            load_q<QV>(arr[4 + c] + orow[2 * (j - 1) + 1], gp[c]);
        }
        const int q = j == 1 ? y : z, n = j == 1 ? a.ny : a.nz;
        const float sm = (per || q >= 1) ? 1.f : -1.f, sp = (per || q <= n - 2) ? 1.f : -1.f;
#pragma unroll
        for (int e = 0; e < QV; ++e) {
            float a_fm[4], a_gm[4], a_fp[4], a_gp[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                a_fm[c] = fm[c][e]; a_gm[c] = sc[c] * gm[c][e]; a_fp[c] = fp[c][e]; a_gp[c] = sc[c] * gp[c][e];
            }
            axis(j, e, a_fm, a_gm, a_fp, a_gp, sm, sp, j == 1 ? a.inv2hy : a.inv2hz);
        }
    }
    float4* out = adj + (size_t(z - a.z_begin) * plane + size_t(y) * a.nx + x);
#pragma unroll
    for (int e = 0; e < QV; ++e) out[e] = make_float4(A[e][0], A[e][1], A[e][2], A[e][3]);
}

// Phase B works on PAIRS of hidden units (2q, 2q+1) with packed f32x2 instructions: the kernel is bound by
// the FP32 pipe's issue slots, and a packed instruction retires two lane-operations per slot.  A lane owns
// PPL pairs; HT/2 pairs span LPP lanes, so a warp takes PPW = 32/LPP points per iteration (2 for HT = 32).
template <int HT>
__global__ void __launch_bounds__(GRAD_THREADS, 2) k_phys_grad(const GradArgs a, const float4* __restrict__ adj) {
    constexpr int LPP = HT >= 64 ? 32 : HT / 2;
    constexpr int PPL = HT / 2 / LPP;
    constexpr int PPW = 32 / LPP;
    constexpr int NW = GRAD_THREADS / 32;
    constexpr int NGT = GRAD_NACC * HT + 6;   // accumulators | db2[4] | sum R_sigma^2, sum |R_u|^2
    // ONE staging array (the final sums reuse all of it as doubles, so its three parts must be one object):
    // [0] = cx, cy, cz, -   [1] = A_t   [2] = A_+ (= -A_-), each double-buffered
    __shared__ float4 s_stage[3][2][GRAD_THREADS];
    float4 (*s_x)[GRAD_THREADS] = s_stage[0];
    float4 (*s_gt)[GRAD_THREADS] = s_stage[1];
    float4 (*s_gd)[GRAD_THREADS] = s_stage[2];
    __shared__ unsigned int s_flag;
    static_assert(sizeof(s_stage) >= sizeof(double) * NGT, "final sums reuse the staging array");
    // double accumulators of every thread, [accumulator][thread] (conflict-free): 40-80 registers per thread saved,
    // which is what lets two blocks per SM stay resident at every width
    extern __shared__ double s_dacc[];

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int sub = lane / LPP, lp = lane % LPP;
    const size_t plane = size_t(a.nx) * a.ny;
    const size_t p_begin = size_t(a.z_begin) * plane, p_end = size_t(a.z_end) * plane;
    const size_t nchunks = (p_end - p_begin + GRAD_THREADS - 1) / GRAD_THREADS;

    // this lane's hidden-unit pairs q = lp + LPP j  ->  units (2q, 2q+1).  Multiplied layer-1 weights are kept
    // half-swapped so that ptxas cannot contract the strict products into FFMA2 (mlp_eval.cuh).
    f32x2 b1p[PPL], w0s[PPL], w1s[PPL], w2s[PPL], ptm[PPL], pt0[PPL], ptp[PPL], w2c[PPL][4];
#pragma unroll
    for (int j = 0; j < PPL; ++j) {
        const int h0 = 2 * (lp + LPP * j), h1 = h0 + 1;
        const bool on0 = h0 < a.H, on1 = h1 < a.H;
        float w1a[4], w1b[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            w1a[k] = on0 ? __ldg(a.W1 + h0 * 4 + k) : 0.f;
            w1b[k] = on1 ? __ldg(a.W1 + h1 * 4 + k) : 0.f;
        }
        b1p[j] = pack2(on0 ? __ldg(a.b1 + h0) : 0.f, on1 ? __ldg(a.b1 + h1) : 0.f);
        w0s[j] = pack2(w1b[0], w1a[0]);
        w1s[j] = pack2(w1b[1], w1a[1]);
        w2s[j] = pack2(w1b[2], w1a[2]);
        ptm[j] = pack2(__fmul_rn(w1a[3], a.tc[0]), __fmul_rn(w1b[3], a.tc[0]));
        pt0[j] = pack2(__fmul_rn(w1a[3], a.tc[1]), __fmul_rn(w1b[3], a.tc[1]));
        ptp[j] = pack2(__fmul_rn(w1a[3], a.tc[2]), __fmul_rn(w1b[3], a.tc[2]));
#pragma unroll
        for (int o = 0; o < 4; ++o)
            w2c[j][o] = pack2(on0 ? __ldg(a.W2 + o * a.H + h0) : 0.f, on1 ? __ldg(a.W2 + o * a.H + h1) : 0.f);
    }
#pragma unroll
    for (int k = 0; k < PPL * GRAD_NACC * 2; ++k) s_dacc[k * GRAD_THREADS + threadIdx.x] = 0.0;
    double db2[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};   // db2[0..3], then the two residual square sums

    int buf = 0;
    for (size_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x, buf ^= 1) {
        // ---- stage this thread's point: coordinates, A_t (from k_phys_adjoint) and A_+ = g / (2 dt) --------
        {
            const size_t p = p_begin + ch * GRAD_THREADS + threadIdx.x;
            float4 gt = make_float4(0.f, 0.f, 0.f, 0.f), gd = gt, xc = gt;
            if (p < p_end) {
                const int z = int(p / plane);
                const int rem = int(p - size_t(z) * plane);
                const int y = rem / a.nx, x = rem - y * a.nx;
                const size_t q = p - size_t(a.z_origin) * plane;
                const float rq[4] = {__ldg(a.R[0] + q), __ldg(a.R[1] + q), __ldg(a.R[2] + q), __ldg(a.R[3] + q)};
                gt = __ldg(adj + (p - p_begin));
                gd = make_float4(a.scale_s * rq[0] * a.inv2dt, a.scale_u * rq[1] * a.inv2dt, a.scale_u * rq[2] * a.inv2dt,
                                 a.scale_u * rq[3] * a.inv2dt);
                xc = make_float4(__ldg(a.cxs + x), __ldg(a.cys + y), __ldg(a.czs + z), 0.f);
                db2[0] += double(gt.x); db2[1] += double(gt.y); db2[2] += double(gt.z); db2[3] += double(gt.w);
                db2[4] += double(rq[0]) * double(rq[0]);
                db2[5] += double(rq[1]) * double(rq[1]) + double(rq[2]) * double(rq[2]) + double(rq[3]) * double(rq[3]);
            }
            s_x[buf][threadIdx.x] = xc;
            s_gt[buf][threadIdx.x] = gt;
            s_gd[buf][threadIdx.x] = gd;
        }
        __syncthreads();
        // ---- B: MLP backward; warp `wid` takes 32 of the chunk's points, PPW per iteration ---------------
        f32x2 f[PPL][GRAD_NACC], fD[PPL];
#pragma unroll
        for (int j = 0; j < PPL; ++j) {
            fD[j] = 0ull;
#pragma unroll
            for (int k = 0; k < GRAD_NACC; ++k) f[j][k] = 0ull;
        }
        // ROWB (nx % 32 == 0): the warp's 32 consecutive points lie in one grid row, so W1[h,1]*y and W1[h,2]*z are
        // formed once per batch and the y/z columns of dW1 follow from the batch sum of dz
        const float4 xrow = s_x[buf][wid * 32];
        auto run = [&](auto tag) {
            constexpr bool ROWB = decltype(tag)::value;
            f32x2 P1[PPL], P2[PPL];
            if (ROWB) {
#pragma unroll
                for (int j = 0; j < PPL; ++j) {
                    P1[j] = mul2_rn(w1s[j], bcast2(xrow.y));
                    P2[j] = mul2_rn(w2s[j], bcast2(xrow.z));
                }
            }
            // U points per group: everything up to the mask test for all of them, ONE vote, then the masked part --
            // so that the branch does not separate the points' independent dependency chains
            constexpr int U = PPL == 1 ? 2 : 1;
#pragma unroll 1
            for (int i = 0; i < 32 / PPW; i += U) {
                float tl[U][PPL], th[U][PPL], dl[U][PPL], dh[U][PPL], z0l[U][PPL], z0h[U][PPL], cxs[U];
                bool pml[U][PPL], pmh[U][PPL], ppl[U][PPL], pph[U][PPL];
                f32x2 cy2[U], cz2[U];
                bool mixed = false;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int pi = wid * 32 + (i + u) * PPW + sub;
                    const float4 xc = s_x[buf][pi];
                    const float4 gt = s_gt[buf][pi];
                    const float4 gd = s_gd[buf][pi];
                    const f32x2 cx2 = bcast2(xc.x);
                    cxs[u] = xc.x; cy2[u] = bcast2(xc.y); cz2[u] = bcast2(xc.z);
#pragma unroll
                    for (int j = 0; j < PPL; ++j) {
                        // the forward's operation order, two hidden units at a time:
                        // ((b1 + W1[h,0] x) + W1[h,1] y) + W1[h,2] z, then + (W1[h,3] t_s rounded)
                        f32x2 pre = add2_rn_swapped(b1p[j], mul2_rn(w0s[j], cx2));
                        pre = add2_rn_swapped(pre, ROWB ? P1[j] : mul2_rn(w1s[j], cy2[u]));
                        pre = add2_rn_swapped(pre, ROWB ? P2[j] : mul2_rn(w2s[j], cz2[u]));
                        const f32x2 zm = add2_rn(pre, ptm[j]), z0 = add2_rn(pre, pt0[j]), zp = add2_rn(pre, ptp[j]);
                        float zml, zmh, zpl, zph;
                        unpack2(zm, zml, zmh); unpack2(z0, z0l[u][j], z0h[u][j]); unpack2(zp, zpl, zph);
                        const f32x2 am = pack2(fmaxf(zml, 0.f), fmaxf(zmh, 0.f));
                        const f32x2 a0 = pack2(fmaxf(z0l[u][j], 0.f), fmaxf(z0h[u][j], 0.f));
                        const f32x2 ap = pack2(fmaxf(zpl, 0.f), fmaxf(zph, 0.f));
                        // W2^T A for the time-t adjoint and for A_+ (A_- = -A_+)
                        f32x2 dat = mul2_rn(w2c[j][0], bcast2(gt.x));
                        dat = fma2_rn(w2c[j][1], bcast2(gt.y), dat);
                        dat = fma2_rn(w2c[j][2], bcast2(gt.z), dat);
                        dat = fma2_rn(w2c[j][3], bcast2(gt.w), dat);
                        f32x2 dad = mul2_rn(w2c[j][0], bcast2(gd.x));
                        dad = fma2_rn(w2c[j][1], bcast2(gd.y), dad);
                        dad = fma2_rn(w2c[j][2], bcast2(gd.z), dad);
                        dad = fma2_rn(w2c[j][3], bcast2(gd.w), dad);
                        const f32x2 ad = sub2_rn(ap, am);
                        f[j][6] = fma2_rn(bcast2(gd.x), ad, fma2_rn(bcast2(gt.x), a0, f[j][6]));
                        f[j][7] = fma2_rn(bcast2(gd.y), ad, fma2_rn(bcast2(gt.y), a0, f[j][7]));
                        f[j][8] = fma2_rn(bcast2(gd.z), ad, fma2_rn(bcast2(gt.z), a0, f[j][8]));
                        f[j][9] = fma2_rn(bcast2(gd.w), ad, fma2_rn(bcast2(gt.w), a0, f[j][9]));
                        unpack2(dat, tl[u][j], th[u][j]); unpack2(dad, dl[u][j], dh[u][j]);
                        pml[u][j] = zml > 0.f; pmh[u][j] = zmh > 0.f; ppl[u][j] = zpl > 0.f; pph[u][j] = zph > 0.f;
                        mixed = mixed || (pml[u][j] != ppl[u][j]) || (pmh[u][j] != pph[u][j]);
                    }
                }
                // z_-, z_0, z_+ are monotone in the slice (a rounded constant is added to the same prefix), so the
                // three ReLU masks are equal iff the outer two are.  Warp-uniform fast path for that case
                // (all but ~1e-3 of the units at dt = 2e-3): dz_+ + dz_- = 0 and one mask serves all slices.
                if (!__any_sync(0xffffffffu, mixed)) {
#pragma unroll
                    for (int u = 0; u < U; ++u)
#pragma unroll
                        for (int j = 0; j < PPL; ++j) {
                            const f32x2 dz0 = pack2(ppl[u][j] ? tl[u][j] : 0.f, pph[u][j] ? th[u][j] : 0.f);
                            const f32x2 dzd = pack2(ppl[u][j] ? dl[u][j] : 0.f, pph[u][j] ? dh[u][j] : 0.f);   // dz_+ = -dz_-
                            f[j][0] = fma2_rn(dz0, bcast2(cxs[u]), f[j][0]);
                            if (!ROWB) {
                                f[j][1] = fma2_rn(dz0, cy2[u], f[j][1]);
                                f[j][2] = fma2_rn(dz0, cz2[u], f[j][2]);
                            }
                            f[j][4] = add2_rn(f[j][4], dz0);
                            fD[j] = add2_rn(fD[j], dzd);
                        }
                } else {
#pragma unroll
                    for (int u = 0; u < U; ++u)
#pragma unroll
                        for (int j = 0; j < PPL; ++j) {
                            const f32x2 dz0 = pack2(z0l[u][j] > 0.f ? tl[u][j] : 0.f, z0h[u][j] > 0.f ? th[u][j] : 0.f);
                            const f32x2 dzp = pack2(ppl[u][j] ? dl[u][j] : 0.f, pph[u][j] ? dh[u][j] : 0.f);
                            const f32x2 dzm = pack2(pml[u][j] ? -dl[u][j] : 0.f, pmh[u][j] ? -dh[u][j] : 0.f);
                            const f32x2 dzs = add2_rn(add2_rn(dz0, dzp), dzm);
                            f[j][0] = fma2_rn(dzs, bcast2(cxs[u]), f[j][0]);
                            if (!ROWB) {
                                f[j][1] = fma2_rn(dzs, cy2[u], f[j][1]);
                                f[j][2] = fma2_rn(dzs, cz2[u], f[j][2]);
                            }
                            f[j][3] = add2_rn(f[j][3], dzm);
                            f[j][4] = add2_rn(f[j][4], dz0);
                            f[j][5] = add2_rn(f[j][5], dzp);
                        }
                }
            }
#pragma unroll
            for (int j = 0; j < PPL; ++j) {
                f[j][3] = sub2_rn(f[j][3], fD[j]);   // sum dz_-
                f[j][5] = add2_rn(f[j][5], fD[j]);   // sum dz_+
                if (ROWB) {
                    const f32x2 sdz = add2_rn(add2_rn(f[j][3], f[j][4]), f[j][5]);
                    f[j][1] = mul2_rn(sdz, bcast2(xrow.y));
                    f[j][2] = mul2_rn(sdz, bcast2(xrow.z));
                }
            }
        };
        if ((a.nx & 31) == 0) run(std::true_type{}); else run(std::false_type{});
#pragma unroll
        for (int j = 0; j < PPL; ++j)
#pragma unroll
            for (int k = 0; k < GRAD_NACC; ++k) {
                float lo, hi;
                unpack2(f[j][k], lo, hi);
                s_dacc[((j * GRAD_NACC + k) * 2) * GRAD_THREADS + threadIdx.x] += double(lo);
                s_dacc[((j * GRAD_NACC + k) * 2 + 1) * GRAD_THREADS + threadIdx.x] += double(hi);
            }
    }

    // ---- block partial: sum the warps' accumulators through shared memory (reusing the staging arrays) ----
    // partial layout: [k][h] for the GRAD_NACC accumulators (template width), then db2[4]
    __syncthreads();
    double* s_acc = reinterpret_cast<double*>(&s_stage[0][0][0]);   // NW * 64 doubles per pass: 4 KB of the 24 KB
    double* part = a.partials + size_t(blockIdx.x) * NGT;
#pragma unroll
    for (int j = 0; j < PPL; ++j)
#pragma unroll
        for (int k = 0; k < GRAD_NACC; ++k) {
            double lo = s_dacc[((j * GRAD_NACC + k) * 2) * GRAD_THREADS + threadIdx.x];
            double hi = s_dacc[((j * GRAD_NACC + k) * 2 + 1) * GRAD_THREADS + threadIdx.x];
            if (PPW == 2) {   // the two half-warps hold the same hidden units for different points
                lo += __shfl_xor_sync(0xffffffffu, lo, 16);
                hi += __shfl_xor_sync(0xffffffffu, hi, 16);
            }
            s_acc[(wid * 32 + lane) * 2] = lo;
            s_acc[(wid * 32 + lane) * 2 + 1] = hi;
            __syncthreads();
            if (wid == 0 && sub == 0) {
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int w = 0; w < NW; ++w) { s0 += s_acc[(w * 32 + lane) * 2]; s1 += s_acc[(w * 32 + lane) * 2 + 1]; }
                const int h0 = 2 * (lp + LPP * j);
                part[k * HT + h0] = s0;
                part[k * HT + h0 + 1] = s1;
            }
            __syncthreads();
        }
#pragma unroll
    for (int o = 0; o < 6; ++o) {
        const double v = warp_sum_d(db2[o]);
        if (lane == 0) s_acc[wid * 6 + o] = v;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) s += s_acc[w * 6 + threadIdx.x];
        part[GRAD_NACC * HT + threadIdx.x] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_flag = atomicAdd(a.ticket, 1u);
    __syncthreads();
    if (s_flag != gridDim.x - 1) return;
    __threadfence();
    // last block: block-ordered sums, then the gradient in the reference's layouts (runtime width)
    double* tot = reinterpret_cast<double*>(&s_stage[0][0][0]);
    for (int e = threadIdx.x; e < NGT; e += GRAD_THREADS) {
        double s = 0.0;
        for (unsigned int b = 0; b < gridDim.x; ++b) s += __ldcg(a.partials + size_t(b) * NGT + e);
        tot[e] = s;
    }
    __syncthreads();
    const int H = a.H;
    for (int e = threadIdx.x; e < 9 * H + 4; e += GRAD_THREADS) {
        double v;
        if (e < 4 * H) {
            const int h = e >> 2, k = e & 3;
            v = k < 3 ? tot[k * HT + h]
                      : double(a.tc[0]) * tot[3 * HT + h] + double(a.tc[1]) * tot[4 * HT + h] + double(a.tc[2]) * tot[5 * HT + h];
        } else if (e < 5 * H) {
            const int h = e - 4 * H;
            v = tot[3 * HT + h] + tot[4 * HT + h] + tot[5 * HT + h];
        } else if (e < 9 * H) {
            const int o = (e - 5 * H) / H, h = (e - 5 * H) % H;
            v = tot[(6 + o) * HT + h];
        } else {
            v = tot[GRAD_NACC * HT + (e - 9 * H)];
        }
        a.grad[e] = v;
    }
    if (a.acc_out && threadIdx.x < 2) a.acc_out[threadIdx.x] = tot[GRAD_NACC * HT + 4 + threadIdx.x];
    if (threadIdx.x == 0) *a.ticket = 0u;
}


namespace {
template <int HT>
constexpr size_t grad_smem() { return size_t(HT >= 64 ? HT / 64 : 1) * GRAD_NACC * 2 * GRAD_THREADS * sizeof(double); }

template <int HT>
int blocks_per_sm_t(int* out) {
    cudaError_t e = cudaFuncSetAttribute(k_phys_grad<HT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(grad_smem<HT>()));
    if (e != cudaSuccess) return int(e);
    return int(cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, k_phys_grad<HT>, GRAD_THREADS, grad_smem<HT>()));
}
}  // namespace

int grad_blocks_per_sm(int HT, int* out) {
    switch (HT) {
        case 32: return blocks_per_sm_t<32>(out);
        case 64: return blocks_per_sm_t<64>(out);
        case 128: return blocks_per_sm_t<128>(out);
    }
    return int(cudaErrorInvalidValue);
}

int grad_launch(int HT, const GradArgs& a, const float4* adj, unsigned blocks, cudaStream_t st) {
    switch (HT) {
        case 32: k_phys_grad<32><<<blocks, GRAD_THREADS, grad_smem<32>(), st>>>(a, adj); break;
        case 64: k_phys_grad<64><<<blocks, GRAD_THREADS, grad_smem<64>(), st>>>(a, adj); break;
        case 128: k_phys_grad<128><<<blocks, GRAD_THREADS, grad_smem<128>(), st>>>(a, adj); break;
        default: return int(cudaErrorInvalidValue);
    }
    return int(cudaGetLastError());
}

int adjoint_launch(const GradArgs& a, float4* adj, cudaStream_t st) {
    const size_t pts = size_t(a.nx) * a.ny * (a.z_end - a.z_begin);
    if (pts == 0) return 0;
    bool v4 = a.nx % 4 == 0 && a.cstride % 4 == 0;
    const void* ptrs[6] = {a.s0, a.u0, a.R[0], a.R[1], a.R[2], a.R[3]};
    for (const void* p : ptrs) v4 = v4 && reinterpret_cast<uintptr_t>(p) % 16 == 0;
    if (v4) k_phys_adjoint<4><<<unsigned((pts / 4 + ADJ_THREADS - 1) / ADJ_THREADS), ADJ_THREADS, 0, st>>>(a, adj);
    else k_phys_adjoint<1><<<unsigned((pts + ADJ_THREADS - 1) / ADJ_THREADS), ADJ_THREADS, 0, st>>>(a, adj);
    return int(cudaGetLastError());
}

}  // namespace physad
