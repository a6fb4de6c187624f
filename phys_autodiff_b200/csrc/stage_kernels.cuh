// Stage-wise kernels behind the reference's existing operator names: MLP over a grid / over a
// coordinate array (strict fp32, bit-exact with src/mlp_cpu.cpp), the finite-difference physics
// residual on externally supplied fields with an on-device double reduction, and the residual VJP.
// These are the device-resident, HBM-facing halves of the path; the metric kernel is in
// fused_loss.cuh.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "fused_loss.cuh"

namespace physad {

// Reduced-precision field I/O (additive, REQUIREMENT.md:123-128 "FP16/BF16: inputs / outputs may be 16-bit, differences and
// reductions accumulate in FP32"): the fields live in HBM as __half or __nv_bfloat16, every load widens to fp32 and the
// arithmetic after the load is the fp32 kernels' unchanged.  E = float is the reference path.
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ldg4(const __half* p) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float4 ldg4(const __nv_bfloat16* p) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float ldg1(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ldg1(const __half* p) { return __half2float(__ldg(p)); }
__device__ __forceinline__ float ldg1(const __nv_bfloat16* p) { return __bfloat162float(__ldg(p)); }
__device__ __forceinline__ void st1(float* p, float v) { *p = v; }
__device__ __forceinline__ void st1(__half* p, float v) { *p = __float2half_rn(v); }
__device__ __forceinline__ void st1(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }


// ---------------------------------------------------------------------------------------------
// MLP over the grid, coordinates from the point index (reference: make_grid_coords +
// mlp_forward, src/mlp_grid.cpp:21-43,53-67).  One thread per point of the slab.
//   FIELDS = false: one time slice, AoS float4 [sigma,ux,uy,uz] per point (mlp_grid_infer_cuda).
//   FIELDS = true : three slices t-dt,t,t+dt split into sigma[n] / u[3n] channel-major arrays
//                   (mlp_generate_fields_cuda + split_outputs_to_fields, src/mlp_grid.cpp:69-106),
//                   sharing the layer-1 prefix between the slices.
// For FIELDS=false the single slice's time products are expected in lt[h].y.
// ---------------------------------------------------------------------------------------------
struct GridInferArgs {
    int nx, ny, nz;
    int z_begin, z_end;
    int m1p1;
    const float* cxs; const float* cys; const float* czs;  // axis-coordinate tables (capi.cu: ensure_coord_tables)
    float4* out_aos;    // FIELDS=false
    float* sigma[3];    // FIELDS=true: t-dt, t, t+dt
    float* u[3];
    size_t cstride;     // FIELDS=true: channel stride of the u arrays; 0 = the slab's point count
};

// Block = 32 x 8 threads on a 32 x 32 (x,y) patch of one z plane; a thread evaluates 4 points that share x
// (rows ty, ty+8, ...), which shares b1 + W1[h,0]x and W1[h,2]z between them exactly as in the fused kernel;
// coordinates come from the index tables (no per-point integer or IEEE division); stores are coalesced rows.
template <int H, bool FIELDS, int UNROLL, typename E = float>
__global__ void __launch_bounds__(256) k_mlp_grid(const __grid_constant__ MlpConst<H> w, const GridInferArgs a) {
    constexpr int P = 4;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y0 = blockIdx.y * 32 + (threadIdx.x >> 5);
    const int z = a.z_begin + blockIdx.z;
    const size_t n = a.cstride ? a.cstride : size_t(a.z_end - a.z_begin) * a.ny * a.nx;
    const float cx = __ldg(a.cxs + min(x, a.nx - 1));
    const float cz = __ldg(a.czs + z);
    float cy[P];
#pragma unroll
    for (int j = 0; j < P; ++j) cy[j] = __ldg(a.cys + min(y0 + 8 * j, a.ny - 1));
    if (FIELDS) {
        float o[P][3][4];
        mlp_eval<H, 3, P, UNROLL, true>(w, cx, cy, cz, o);
#pragma unroll
        for (int j = 0; j < P; ++j) {
            const int y = y0 + 8 * j;
            if (x < a.nx && y < a.ny) {
                const size_t i = (size_t(blockIdx.z) * a.ny + y) * a.nx + x;
#pragma unroll
                for (int s = 0; s < 3; ++s) {
                    E* ps = reinterpret_cast<E*>(a.sigma[s]);   // E != float: the six arrays hold 16-bit elements
                    E* pu = reinterpret_cast<E*>(a.u[s]);
                    st1(ps + i, o[j][s][0]);
                    st1(pu + i, o[j][s][1]);
                    st1(pu + n + i, o[j][s][2]);
                    st1(pu + 2 * n + i, o[j][s][3]);
                }
            }
        }
    } else {
        float o[P][1][4];
        mlp_eval<H, 1, P, UNROLL, true>(w, cx, cy, cz, o);
#pragma unroll
        for (int j = 0; j < P; ++j) {
            const int y = y0 + 8 * j;
            if (x < a.nx && y < a.ny)
                a.out_aos[(size_t(blockIdx.z) * a.ny + y) * a.nx + x] = make_float4(o[j][0][0], o[j][0][1], o[j][0][2], o[j][0][3]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// MLP on an explicit coordinate/feature array (mlp_forward<ExecCuda>, include/mlp.h:5-6), In = Out = 4: weights
// staged once per block in shared memory in their reference layouts, one point per thread.  Any other In / Out, and
// the MSE backward (mlp_backward<ExecCuda>), go through the strict register-tiled contraction of dense_kernels.cuh.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_mlp_forward_4x4(const float4* __restrict__ x, const float* __restrict__ W1,
                                                         const float* __restrict__ b1, const float* __restrict__ W2,
                                                         const float* __restrict__ b2, float4* __restrict__ y, size_t B,
                                                         int H) {
    extern __shared__ float4 sw[];  // [H] {b1,w0,w1,w2}, [H] {w3, W2[0..2][h]}, [H] {W2[3][h],..}
    float4* s_a = sw;
    float4* s_b = sw + H;
    float* s_c = reinterpret_cast<float*>(sw + 2 * H);
    for (int h = threadIdx.x; h < H; h += blockDim.x) {
        s_a[h] = make_float4(b1[h], W1[h * 4 + 0], W1[h * 4 + 1], W1[h * 4 + 2]);
        s_b[h] = make_float4(W1[h * 4 + 3], W2[h], W2[H + h], W2[2 * H + h]);
        s_c[h] = W2[3 * H + h];
    }
    __syncthreads();
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const float4 p = x[i];
    float o0 = b2[0], o1 = b2[1], o2 = b2[2], o3 = b2[3];
#pragma unroll 4
    for (int h = 0; h < H; ++h) {
        const float4 a = s_a[h], c = s_b[h];
        const float d = s_c[h];
        float s = __fadd_rn(a.x, __fmul_rn(a.y, p.x));
        s = __fadd_rn(s, __fmul_rn(a.z, p.y));
        s = __fadd_rn(s, __fmul_rn(a.w, p.z));
        s = __fadd_rn(s, __fmul_rn(c.x, p.w));
        const float act = relu_ref(s);
        o0 = __fadd_rn(o0, __fmul_rn(c.y, act));
        o1 = __fadd_rn(o1, __fmul_rn(c.z, act));
        o2 = __fadd_rn(o2, __fmul_rn(c.w, act));
        o3 = __fadd_rn(o3, __fmul_rn(d, act));
    }
    y[i] = make_float4(o0, o1, o2, o3);
}

// ---------------------------------------------------------------------------------------------
// Physics residual on supplied fields (reference src/phys_cpu.cpp:25-110; fp32 like the reference's
// CUDA kernels).  HBM-bound: 48 B in + 16 B out per point, everything else must stay out of the way.
//   * a block owns a 32 x 8 (x,y) tile and MARCHES over a chunk of z planes: the time-t values of
//     planes z-1, z, z+1 roll through registers, so z neighbours cost no loads; x/y neighbours are L1
//     hits on lines the block's own centre loads brought in (a warp is one 128-byte row segment);
//   * indices come from the 3-D launch geometry (no integer division per point), wrap/clamp of a +-1
//     offset is two compares;
//   * REDUCE accumulates the squares per thread in double over the whole march, so the double
//     shuffles / ticket of grid_reduce2 run once per block, not once per 256 points.
//   WRITE_R : store the four residual arrays
//   REDUCE  : {sum Rs^2, sum |Ru|^2} in double -> acc_out (on-device loss reduction)
//   SCALE   : store scale_s*Rs, scale_u*Ru instead of R (backward recomputed from fields,
//             cuda_phys_loss_backward_fused, include/phys.h:132-143)
// ---------------------------------------------------------------------------------------------
struct PhysArgs {
    int nx, ny, nz;   // nz = number of planes IN THE ARRAYS (the slab's planes in slab mode)
    int periodic;
    int clamp_z;  // 1: clamp the z neighbours at the ends of the arrays even on a periodic grid (the arrays are a
                  //    window of planes whose outermost residuals the caller discards, see physad_fused_loss_grad_slab_dev)
    int zc;  // planes per z chunk
    // slab mode (multi-GPU on supplied fields): the arrays hold only this rank's planes, and the time-t planes
    // just below / above the slab ([4 channels][ny][nx] each, already wrapped/clamped by the host) come from
    // these two buffers.  Null = whole grid in the arrays, wrap/clamp inside them.
    const float* halo_lo;
    const float* halo_hi;
    float inv2dt, inv2hx, inv2hy, inv2hz;
    double inv2dt_d, inv2hx_d, inv2hy_d, inv2hz_d;  // exact-residual mode (DPRES)
    float inv1hx, inv1hy, inv1hz;                   // 1/h: one-sided differences of the upwind scheme
    double inv1hx_d, inv1hy_d, inv1hz_d;
    float scale_s, scale_u;
    const float* s_m; const float* s_0; const float* s_p;
    const float* u_m; const float* u_0; const float* u_p;
    float* R[4];
    double2* partials; unsigned int* ticket; double* acc_out;
};

// time-t plane just below / above plane z of channel c (see PhysArgs::halo_lo/hi)
template <typename E>
__device__ __forceinline__ const E* plane_below(const PhysArgs& a, const E* f0c, int c, int z, size_t pln, bool per) {
    if (z > 0) return f0c + size_t(z - 1) * pln;
    if (a.halo_lo) return reinterpret_cast<const E*>(a.halo_lo) + size_t(c) * pln;
    return f0c + size_t((per && !a.clamp_z) ? a.nz - 1 : 0) * pln;
}
template <typename E>
__device__ __forceinline__ const E* plane_above(const PhysArgs& a, const E* f0c, int c, int z, size_t pln, bool per) {
    if (z + 1 < a.nz) return f0c + size_t(z + 1) * pln;
    if (a.halo_hi) return reinterpret_cast<const E*>(a.halo_hi) + size_t(c) * pln;
    return f0c + size_t((per && !a.clamp_z) ? 0 : a.nz - 1) * pln;
}

// neighbour index for an offset of +-1 (n >= 1): wrap or clamp without a division
__device__ __forceinline__ int nb1(int v, int n, bool periodic) {
    if (periodic) return v < 0 ? v + n : (v >= n ? v - n : v);
    return v < 0 ? 0 : (v >= n ? n - 1 : v);
}

template <bool WRITE_R, bool REDUCE, bool SCALE, bool DPRES, bool UPWIND = false>
__global__ void __launch_bounds__(256) k_phys_residual(const PhysArgs a) {
    using real = typename std::conditional<DPRES, double, float>::type;
    const real i2t = DPRES ? real(a.inv2dt_d) : real(a.inv2dt), i2x = DPRES ? real(a.inv2hx_d) : real(a.inv2hx);
    const real i2y = DPRES ? real(a.inv2hy_d) : real(a.inv2hy), i2z = DPRES ? real(a.inv2hz_d) : real(a.inv2hz);
    __shared__ double2 s_red[8];
    __shared__ unsigned int s_flag;
    const bool per = a.periodic != 0;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int z0 = blockIdx.z * a.zc, z1 = min(a.nz, z0 + a.zc);
    const size_t N = size_t(a.nx) * a.ny * a.nz, pln = size_t(a.nx) * a.ny;
    double acc_s = 0.0, acc_u = 0.0;
    if (x < a.nx && y < a.ny) {
        const size_t row = size_t(y) * a.nx;
        const size_t oc = row + x, oxm = row + nb1(x - 1, a.nx, per), oxp = row + nb1(x + 1, a.nx, per);
        const size_t oym = size_t(nb1(y - 1, a.ny, per)) * a.nx + x, oyp = size_t(nb1(y + 1, a.ny, per)) * a.nx + x;
        const float* f0[4] = {a.s_0, a.u_0, a.u_0 + N, a.u_0 + 2 * N};
        const float* fm[4] = {a.s_m, a.u_m, a.u_m + N, a.u_m + 2 * N};
        const float* fp[4] = {a.s_p, a.u_p, a.u_p + N, a.u_p + 2 * N};
        // Software pipeline: the 12 DRAM-bound centre loads of plane z+1 (time-t value that becomes `hi`, and
        // the t+-dt values) are issued one iteration ahead, so a full plane of arithmetic and L1-hit
        // neighbour loads covers their latency (ncu on the unpipelined form: long-scoreboard stalls 13 per
        // issue, DRAM 55 % busy).
        float lo[4], mid[4], hi[4];      // time-t values at z-1, z, z+1 of this column
        float tp[4], tm[4];              // t+dt / t-dt values at z
        {
            const size_t pc = size_t(z0) * pln + oc;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                lo[c] = __ldg(plane_below(a, f0[c], c, z0, pln, per) + oc);
                mid[c] = __ldg(f0[c] + pc);
                hi[c] = __ldg(plane_above(a, f0[c], c, z0, pln, per) + oc);
                tp[c] = __ldg(fp[c] + pc); tm[c] = __ldg(fm[c] + pc);
            }
        }
        for (int z = z0; z < z1; ++z) {
            const size_t pz = size_t(z) * pln;
            // prefetch for the next iteration (plane z+1's t+-dt values, plane z+2's time-t value)
            const int zn = min(z + 1, a.nz - 1);  // clamped only to stay in bounds on the last iteration
            const size_t pzn = size_t(zn) * pln + oc;
            float ntp[4], ntm[4], nhi[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                ntp[c] = __ldg(fp[c] + pzn); ntm[c] = __ldg(fm[c] + pzn);
                nhi[c] = __ldg(plane_above(a, f0[c], c, zn, pln, per) + oc);
            }
            real dT[4];
            float R[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) dT[c] = central_diff(tp[c], tm[c], i2t);
            if (UPWIND) {
                float xm[4], xp[4], ym[4], yp[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    xp[c] = __ldg(f0[c] + pz + oxp); xm[c] = __ldg(f0[c] + pz + oxm);
                    yp[c] = __ldg(f0[c] + pz + oyp); ym[c] = __ldg(f0[c] + pz + oym);
                }
                point_residual_upwind<real>(mid, xm, xp, ym, yp, lo, hi, dT,
                                            DPRES ? real(a.inv1hx_d) : real(a.inv1hx), DPRES ? real(a.inv1hy_d) : real(a.inv1hy),
                                            DPRES ? real(a.inv1hz_d) : real(a.inv1hz), i2x, i2y, i2z, R);
            } else {
                real gx[4], gy[4], gz[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    gx[c] = central_diff(__ldg(f0[c] + pz + oxp), __ldg(f0[c] + pz + oxm), i2x);
                    gy[c] = central_diff(__ldg(f0[c] + pz + oyp), __ldg(f0[c] + pz + oym), i2y);
                    gz[c] = central_diff(hi[c], lo[c], i2z);
                }
                point_residual(mid, gx, gy, gz, dT, R);
            }
            if (WRITE_R) {
                const size_t i = pz + oc;
                if (a.R[0]) a.R[0][i] = SCALE ? a.scale_s * R[0] : R[0];
                if (a.R[1]) a.R[1][i] = SCALE ? a.scale_u * R[1] : R[1];
                if (a.R[2]) a.R[2][i] = SCALE ? a.scale_u * R[2] : R[2];
                if (a.R[3]) a.R[3][i] = SCALE ? a.scale_u * R[3] : R[3];
            }
            if (REDUCE) {
                acc_s += double(R[0]) * double(R[0]);
                acc_u += double(R[1]) * double(R[1]) + double(R[2]) * double(R[2]) + double(R[3]) * double(R[3]);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) { lo[c] = mid[c]; mid[c] = hi[c]; hi[c] = nhi[c]; tp[c] = ntp[c]; tm[c] = ntm[c]; }
        }
    }
    if (REDUCE) {
        // grid_reduce2 indexes partials by a linear block id
        const unsigned int lin = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        const unsigned int nblk = gridDim.x * gridDim.y * gridDim.z;
        grid_reduce2_lin<8>(acc_s, acc_u, a.partials, a.ticket, a.acc_out, s_red, &s_flag, lin, nblk);
    }
}

// ---------------------------------------------------------------------------------------------
// The same operator with 128-bit accesses: a thread owns FOUR consecutive x of one row and marches z.
// Per plane it issues 12 LDG.128 that miss to DRAM (time-t of plane z+1, t+dt and t-dt of plane z), 8 LDG.128
// for the y neighbours (L1/L2 hits) and 8 scalar loads for the two x neighbours outside its quad, instead of
// 28 scalar loads per point: 3.7x fewer load instructions and 4x the bytes in flight per thread, which is
// what the latency-bound scalar form lacks.  Needs nx % 4 == 0 and 16-byte aligned arrays (capi.cu checks
// and otherwise uses k_phys_residual).  Block = 64 quads (256 columns) x 2 rows (V4_THREADS = 128).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float comp(const float4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }

constexpr int V4_THREADS = 128;   // 64 quads x 2 rows: ~170 registers per thread, so three small blocks fit an SM where one of 256 did
template <bool WRITE_R, bool REDUCE, bool SCALE, bool DPRES, typename E = float>
__global__ void __launch_bounds__(V4_THREADS, DPRES ? 2 : 3) k_phys_residual_v4(const PhysArgs a) {
    using real = typename std::conditional<DPRES, double, float>::type;
    const real i2t = DPRES ? real(a.inv2dt_d) : real(a.inv2dt), i2x = DPRES ? real(a.inv2hx_d) : real(a.inv2hx);
    const real i2y = DPRES ? real(a.inv2hy_d) : real(a.inv2hy), i2z = DPRES ? real(a.inv2hz_d) : real(a.inv2hz);
    __shared__ double2 s_red[V4_THREADS / 32];
    __shared__ unsigned int s_flag;
    const bool per = a.periodic != 0;
    const int x = (blockIdx.x * 64 + (threadIdx.x & 63)) * 4, y = blockIdx.y * (V4_THREADS / 64) + (threadIdx.x >> 6);
    const int z0 = blockIdx.z * a.zc, z1 = min(a.nz, z0 + a.zc);
    const size_t N = size_t(a.nx) * a.ny * a.nz, pln = size_t(a.nx) * a.ny;
    double acc_s = 0.0, acc_u = 0.0;
    if (x < a.nx && y < a.ny) {
        const size_t row = size_t(y) * a.nx;
        const size_t oc = row + x, oxm = row + nb1(x - 1, a.nx, per), oxp = row + nb1(x + 4, a.nx, per);
        const size_t oym = size_t(nb1(y - 1, a.ny, per)) * a.nx + x, oyp = size_t(nb1(y + 1, a.ny, per)) * a.nx + x;
        auto el = [](const float* p) { return reinterpret_cast<const E*>(p); };   // the fields hold elements of type E
        const E* f0[4] = {el(a.s_0), el(a.u_0), el(a.u_0) + N, el(a.u_0) + 2 * N};
        const E* fm[4] = {el(a.s_m), el(a.u_m), el(a.u_m) + N, el(a.u_m) + 2 * N};
        const E* fp[4] = {el(a.s_p), el(a.u_p), el(a.u_p) + N, el(a.u_p) + 2 * N};
        float4 lo[4], mid[4], hi[4];
        {
            const size_t pc = size_t(z0) * pln + oc;
#pragma unroll
            for (int c = 0; c < 4; ++c) { lo[c] = ldg4(plane_below(a, f0[c], c, z0, pln, per) + oc); mid[c] = ldg4(f0[c] + pc); }
        }
        for (int z = z0; z < z1; ++z) {
            const size_t pz = size_t(z) * pln;
            float4 tp[4], tm[4], yp[4], ym[4];
            float xl[4], xr[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                hi[c] = ldg4(plane_above(a, f0[c], c, z, pln, per) + oc);
                tp[c] = ldg4(fp[c] + pz + oc);
                tm[c] = ldg4(fm[c] + pz + oc);
                yp[c] = ldg4(f0[c] + pz + oyp);
                ym[c] = ldg4(f0[c] + pz + oym);
                xl[c] = ldg1(f0[c] + pz + oxm);
                xr[c] = ldg1(f0[c] + pz + oxp);
            }
            float4 Rv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {  // the four points of the quad
                float f[4], R[4];
                real dT[4], gx[4], gy[4], gz[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    f[c] = comp(mid[c], e);
                    const float left = e == 0 ? xl[c] : comp(mid[c], e - 1);
                    const float right = e == 3 ? xr[c] : comp(mid[c], e + 1);
                    dT[c] = central_diff(comp(tp[c], e), comp(tm[c], e), i2t);
                    gx[c] = central_diff(right, left, i2x);
                    gy[c] = central_diff(comp(yp[c], e), comp(ym[c], e), i2y);
                    gz[c] = central_diff(comp(hi[c], e), comp(lo[c], e), i2z);
                }
                point_residual(f, gx, gy, gz, dT, R);
                if (REDUCE) {
                    acc_s += double(R[0]) * double(R[0]);
                    acc_u += double(R[1]) * double(R[1]) + double(R[2]) * double(R[2]) + double(R[3]) * double(R[3]);
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float v = SCALE ? (c == 0 ? a.scale_s : a.scale_u) * R[c] : R[c];
                    if (e == 0) Rv[c].x = v; else if (e == 1) Rv[c].y = v; else if (e == 2) Rv[c].z = v; else Rv[c].w = v;
                }
            }
            if (WRITE_R) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (a.R[c]) *reinterpret_cast<float4*>(a.R[c] + pz + oc) = Rv[c];
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) { lo[c] = mid[c]; mid[c] = hi[c]; }
        }
    }
    if (REDUCE) {
        const unsigned int lin = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        const unsigned int nblk = gridDim.x * gridDim.y * gridDim.z;
        grid_reduce2_lin<V4_THREADS / 32>(acc_s, acc_u, a.partials, a.ticket, a.acc_out, s_red, &s_flag, lin, nblk);
    }
}

// g = scale * R (reference src/phys_cpu.cpp:151-170); four arrays in one launch, float4-vectorised
// when N % 4 == 0 (all pointers come from cudaMalloc or are at least 16-byte aligned by contract).
struct ScaleArgs {
    const float* R[4];
    float* G[4];
    float scale[4];
    size_t n;
};

__global__ void __launch_bounds__(256) k_scale4(const ScaleArgs a, int vec4) {
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    const size_t t0 = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float s = a.scale[c];
        if (vec4) {
            const float4* r = reinterpret_cast<const float4*>(a.R[c]);
            float4* g = reinterpret_cast<float4*>(a.G[c]);
            for (size_t i = t0; i < a.n / 4; i += stride) {
                const float4 v = __ldg(r + i);
                g[i] = make_float4(s * v.x, s * v.y, s * v.z, s * v.w);
            }
        } else {
            for (size_t i = t0; i < a.n; i += stride) a.G[c][i] = s * __ldg(a.R[c] + i);
        }
    }
}

}  // namespace physad
