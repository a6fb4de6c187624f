// Deeper coordinate MLPs (BASELINE config 5: depth sweep): In = 4 -> H -> H -> ... -> H -> Out = 4 with L >= 1
// hidden layers.  ADDITIVE: the reference API (include/mlp.h:5-6) has exactly one hidden layer, so for L > 1
// there is no reference implementation to pin against -- the semantics are the reference's layer rule applied
// again (src/mlp_cpu.cpp:19-24: start from the bias, add W[g,h]*a[h] for h ascending, separate fp32 multiply and
// add, ReLU), restated for the CPU by the test oracle (oracle_mlp_forward_deep, "parity unpinned" for L > 1;
// L = 1 is pinned: it must equal the one-hidden-layer path bit for bit, and the tests check that).
//
// Kernel: one thread per grid point, stage-wise (fields out, physics by the stage-wise kernels).  For each of
// the three time slices the H activations of the current layer live in REGISTERS (the h loop is fully
// unrolled so they are statically indexed); a hidden->hidden layer produces its outputs 16 at a time in
// packed accumulators (FMUL2/FADD2 over output pairs, as in mlp_eval.cuh), streaming the weights with
// uniform 128-bit loads from a host-prepared layout (transposed [h][g], pairs half-swapped for the
// no-contraction trick); outputs go through a per-thread shared-memory column (dynamic group index) and
// come back as the next layer's register activations.  Strict fp32 throughout: the FP32 pipe is the
// roofline, 2*H*H lane-operations per point, slice and extra layer.
#pragma once
#include "mlp_eval.cuh"

namespace physad {

struct DeepArgs {
    int nx, ny, nz;
    int z_begin, z_end;
    int hidden_layers;          // L >= 1
    const float* cxs; const float* cys; const float* czs;
    const float* wh;            // [(L-1)][H][H]: for layer l, input h: H outputs, pairs stored (g+1, g)
    const float* bh;            // [(L-1)][H]
    float4* out_aos;            // FIELDS = false
    float* sigma[3];            // FIELDS = true
    float* u[3];
};

template <int H, bool FIELDS>
__global__ void __launch_bounds__(128) k_mlp_deep(const __grid_constant__ MlpConst<H> w, const DeepArgs a) {
    constexpr int NT = 128, G = 16;
    extern __shared__ float s_col[];  // [H][NT]: this thread's column of layer outputs
    const int x = blockIdx.x * NT + threadIdx.x, y = blockIdx.y, z = a.z_begin + blockIdx.z;
    if (x >= a.nx) return;
    const size_t n = size_t(a.z_end - a.z_begin) * a.ny * a.nx;
    const size_t i = (size_t(blockIdx.z) * a.ny + y) * a.nx + x;
    const float cx = __ldg(a.cxs + x), cy = __ldg(a.cys + y), cz = __ldg(a.czs + z);
    float* col = s_col + threadIdx.x;
    constexpr int NS = FIELDS ? 3 : 1;
    float out[NS][4];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        float act[H];
        // layer 1 (the three slices differ only in the pre-rounded W1[h,3]*t term; FIELDS=false uses pt0)
#pragma unroll
        for (int q = 0; q < H / 2; ++q) {
            const float2 b1p = w.b1p[q], w0s = w.w0s[q], w1s = w.w1s[q], w2s = w.w2s[q];
            const float2 pt = NS == 1 ? w.pt0[q] : (s == 0 ? w.ptm[q] : (s == 1 ? w.pt0[q] : w.ptp[q]));
            const f32x2 sx = add2_rn_swapped(pack2(b1p), mul2_rn(pack2(w0s), bcast2(cx)));
            const f32x2 sxy = add2_rn_swapped(sx, mul2_rn(pack2(w1s), bcast2(cy)));
            const f32x2 sxyz = add2_rn_swapped(sxy, mul2_rn(pack2(w2s), bcast2(cz)));
            float v0, v1;
            unpack2(add2_rn(sxyz, pack2(pt)), v0, v1);
            act[2 * q] = relu_ref(v0);
            act[2 * q + 1] = relu_ref(v1);
        }
        // hidden -> hidden layers
        for (int l = 0; l + 1 < a.hidden_layers; ++l) {
            const float* wl = a.wh + size_t(l) * H * H;
            const float* bl = a.bh + size_t(l) * H;
            for (int g0 = 0; g0 < H; g0 += G) {
                f32x2 q[G / 2];
#pragma unroll
                for (int p = 0; p < G / 2; ++p) q[p] = pack2(__ldg(bl + g0 + 2 * p), __ldg(bl + g0 + 2 * p + 1));
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    const float4* wp = reinterpret_cast<const float4*>(wl + size_t(h) * H + g0);
                    const f32x2 aa = bcast2(act[h]);
#pragma unroll
                    for (int v = 0; v < G / 4; ++v) {
                        const float4 c = __ldg(wp + v);  // {W[g+1,h], W[g,h], W[g+3,h], W[g+2,h]}
                        q[2 * v] = add2_rn_swapped(q[2 * v], mul2_rn(aa, pack2(c.x, c.y)));
                        q[2 * v + 1] = add2_rn_swapped(q[2 * v + 1], mul2_rn(aa, pack2(c.z, c.w)));
                    }
                }
#pragma unroll
                for (int p = 0; p < G / 2; ++p) {
                    float v0, v1;
                    unpack2(q[p], v0, v1);
                    col[(g0 + 2 * p) * NT] = relu_ref(v0);
                    col[(g0 + 2 * p + 1) * NT] = relu_ref(v1);
                }
            }
#pragma unroll
            for (int h = 0; h < H; ++h) act[h] = col[h * NT];  // own column only: no synchronisation needed
        }
        // output layer (h ascending, packed over output pairs)
        const float4 b2 = w.b2;
        f32x2 q01 = pack2(b2.x, b2.y), q23 = pack2(b2.z, b2.w);
#pragma unroll
        for (int h = 0; h < H; ++h) {
            const float4 c = w.w2[h];
            const f32x2 aa = bcast2(act[h]);
            q01 = add2_rn_swapped(q01, mul2_rn(aa, pack2(c.x, c.y)));
            q23 = add2_rn_swapped(q23, mul2_rn(aa, pack2(c.z, c.w)));
        }
        unpack2(q01, out[s][0], out[s][1]);
        unpack2(q23, out[s][2], out[s][3]);
    }
    if (FIELDS) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            a.sigma[s][i] = out[s][0];
            a.u[s][i] = out[s][1];
            a.u[s][n + i] = out[s][2];
            a.u[s][2 * n + i] = out[s][3];
        }
    } else {
        a.out_aos[i] = make_float4(out[0][0], out[0][1], out[0][2], out[0][3]);
    }
}

}  // namespace physad
