// The backward kernel of the closed loop; see grad_kernels.cuh for what it computes.
#include "grad_kernels.cuh"
#include "mlp_eval.cuh"

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

// Phase B works on PAIRS of hidden units (2q, 2q+1) with packed f32x2 instructions: the kernel is bound by
// the FP32 pipe's issue slots, and a packed instruction retires two lane-operations per slot.  A lane owns
// PPL pairs; HT/2 pairs span LPP lanes, so a warp takes PPW = 32/LPP points per iteration (2 for HT = 32).
template <int HT>
__global__ void __launch_bounds__(GRAD_THREADS, 2) k_phys_grad(const GradArgs a) {
    constexpr int LPP = HT >= 64 ? 32 : HT / 2;
    constexpr int PPL = HT / 2 / LPP;
    constexpr int PPW = 32 / LPP;
    constexpr int NW = GRAD_THREADS / 32;
    constexpr int NGT = GRAD_NACC * HT + 6;   // accumulators | db2[4] | sum R_sigma^2, sum |R_u|^2
    __shared__ float4 s_x[2][GRAD_THREADS];  // cx, cy, cz, -
    __shared__ float4 s_gt[2][GRAD_THREADS]; // A_t
    __shared__ float4 s_gd[2][GRAD_THREADS]; // A_+ (= -A_-)
    __shared__ unsigned int s_flag;
    static_assert(sizeof(float4) * 2 * GRAD_THREADS * 3 >= sizeof(double) * NGT, "final sums reuse the staging arrays");
    // double accumulators of every thread, [accumulator][thread] (conflict-free): registers are better spent on
    // a third resident block -- the loop is latency-bound at two
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
                const float rq[4] = {__ldg(a.R[0] + q), __ldg(a.R[1] + q), __ldg(a.R[2] + q), __ldg(a.R[3] + q)};
                const float gq[4] = {a.scale_s * rq[0], a.scale_u * rq[1], a.scale_u * rq[2], a.scale_u * rq[3]};
                db2[4] += double(rq[0]) * double(rq[0]);
                db2[5] += double(rq[1]) * double(rq[1]) + double(rq[2]) * double(rq[2]) + double(rq[3]) * double(rq[3]);
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
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        float fl[2];
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            fl[e] = g[e][c] * f[e][j + 1];
                            if (c == j + 1) fl[e] += g[e][0] * f[e][0];
                        }
                        A[c] += i2h[j] * (sg[j][0] * fl[0] - sg[j][1] * fl[1]);
                    }
                }
                gt = make_float4(A[0], A[1], A[2], A[3]);
                gd = make_float4(gq[0] * a.inv2dt, gq[1] * a.inv2dt, gq[2] * a.inv2dt, gq[3] * a.inv2dt);
                xc = make_float4(__ldg(a.cxs + x), __ldg(a.cys + y), __ldg(a.czs + z), 0.f);
                db2[0] += double(A[0]); db2[1] += double(A[1]); db2[2] += double(A[2]); db2[3] += double(A[3]);
            }
            s_x[buf][threadIdx.x] = xc;
            s_gt[buf][threadIdx.x] = gt;
            s_gd[buf][threadIdx.x] = gd;
        }
        __syncthreads();
        // ---- B: MLP backward; warp `wid` takes 32 of the chunk's points, PPW per iteration ---------------
        f32x2 f[PPL][GRAD_NACC];
#pragma unroll
        for (int j = 0; j < PPL; ++j)
#pragma unroll
            for (int k = 0; k < GRAD_NACC; ++k) f[j][k] = 0ull;
#pragma unroll 2
        for (int i = 0; i < 32 / PPW; ++i) {
            const int pi = wid * 32 + i * PPW + sub;
            const float4 xc = s_x[buf][pi];
            const float4 gt = s_gt[buf][pi];
            const float4 gd = s_gd[buf][pi];
            const f32x2 cx2 = bcast2(xc.x), cy2 = bcast2(xc.y), cz2 = bcast2(xc.z);
#pragma unroll
            for (int j = 0; j < PPL; ++j) {
                // the forward's operation order, two hidden units at a time:
                // ((b1 + W1[h,0] x) + W1[h,1] y) + W1[h,2] z, then + (W1[h,3] t_s rounded)
                f32x2 pre = add2_rn_swapped(b1p[j], mul2_rn(w0s[j], cx2));
                pre = add2_rn_swapped(pre, mul2_rn(w1s[j], cy2));
                pre = add2_rn_swapped(pre, mul2_rn(w2s[j], cz2));
                const f32x2 zm = add2_rn(pre, ptm[j]), z0 = add2_rn(pre, pt0[j]), zp = add2_rn(pre, ptp[j]);
                float zml, zmh, z0l, z0h, zpl, zph;
                unpack2(zm, zml, zmh); unpack2(z0, z0l, z0h); unpack2(zp, zpl, zph);
                const f32x2 am = pack2(fmaxf(zml, 0.f), fmaxf(zmh, 0.f));
                const f32x2 a0 = pack2(fmaxf(z0l, 0.f), fmaxf(z0h, 0.f));
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
                float tl, th, dl, dh;
                unpack2(dat, tl, th); unpack2(dad, dl, dh);
                const f32x2 dz0 = pack2(z0l > 0.f ? tl : 0.f, z0h > 0.f ? th : 0.f);
                const f32x2 dzp = pack2(zpl > 0.f ? dl : 0.f, zph > 0.f ? dh : 0.f);
                const f32x2 dzm = pack2(zml > 0.f ? -dl : 0.f, zmh > 0.f ? -dh : 0.f);
                const f32x2 dzs = add2_rn(add2_rn(dz0, dzp), dzm);
                f[j][0] = fma2_rn(dzs, cx2, f[j][0]);
                f[j][1] = fma2_rn(dzs, cy2, f[j][1]);
                f[j][2] = fma2_rn(dzs, cz2, f[j][2]);
                f[j][3] = add2_rn(f[j][3], dzm);
                f[j][4] = add2_rn(f[j][4], dz0);
                f[j][5] = add2_rn(f[j][5], dzp);
            }
        }
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
    double* s_acc = reinterpret_cast<double*>(&s_x[0][0]);   // NW * 64 doubles per pass: 4 KB of the 8 KB
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
    double* tot = reinterpret_cast<double*>(&s_x[0][0]);
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

int grad_launch(int HT, const GradArgs& a, unsigned blocks, cudaStream_t st) {
    switch (HT) {
        case 32: k_phys_grad<32><<<blocks, GRAD_THREADS, grad_smem<32>(), st>>>(a); break;
        case 64: k_phys_grad<64><<<blocks, GRAD_THREADS, grad_smem<64>(), st>>>(a); break;
        case 128: k_phys_grad<128><<<blocks, GRAD_THREADS, grad_smem<128>(), st>>>(a); break;
        default: return int(cudaErrorInvalidValue);
    }
    return int(cudaGetLastError());
}

}  // namespace physad
