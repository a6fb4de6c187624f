// Deeper coordinate MLPs, fast mode: hidden -> hidden layers on tcgen05 with three-term bf16 operands.
// What is computed, and why it is additive / not bit-exact: deep_tc_kernels.cuh.
//
// One 256-thread block per SM, persistent over tiles of 128 points; the three time slices of a tile go through the
// layers one after the other ("row" = (point, slice), 128 rows = the M of one MMA).
//   * thread (m, half): row m = its tensor-memory lane (warp w may touch lanes 32 (w % 4) ..), half = w / 4 picks which
//     H/2 hidden units it produces in layer 1 and which H/2 accumulator columns it drains after a layer;
//   * tensor memory columns: [0, H) the fp32 accumulators D; then the three bf16 terms of the A operand (activations),
//     H/2 columns each, two K elements per 32-bit column -- written by tcgen05.st from the thread that owns the row, so
//     activations never visit shared memory and no proxy fence is needed between an epilogue and the next layer;
//   * shared memory: the operand images of ALL hidden -> hidden layers (3 terms x H x H bf16 each, K-major no-swizzle core
//     matrices, written by deep_tc_pack_layer on the host and bulk-copied once per block), the layer-1 pairs, the output
//     layer, the biases;
//   * a layer = 6 (term pairs) x H/16 (K slices) tcgen05.mma issued by thread 0, smallest products first, then ONE
//     tcgen05.commit -> mbarrier; every thread waits on it, drains its columns (tcgen05.ld), adds the bias, applies
//     ReLU and either splits into the next layer's terms or, after the last hidden layer, accumulates the four outputs.
// Governing roofline: the bf16 tensor pipe at 6 MMA passes per fp32-equivalent contraction: 2 H^2 x 6 flop per row and
// layer against MEASURED_PEAKS' dense bf16 figure.  This first version runs MMA and epilogue of a tile back to back
// (one accumulator, one A operand), so the tensor pipe idles while the CUDA cores split activations and vice versa.
#include "deep_tc_kernels.cuh"
#include "mlp_eval.cuh"
#include "tc_common.cuh"

#include <cstring>

namespace physad {

namespace {

using namespace tc;

constexpr int TC_THREADS = 256;
constexpr int TILE = 128;                       // rows of one MMA
constexpr size_t MIN_SMEM = 120 * 1024;         // more than half an SM's shared memory: exactly one block per SM, so the
                                                // block's tensor-memory allocation can never wait for a neighbour's

template <int H>
struct TmemCols { static constexpr uint32_t value = H == 128 ? 512u : (H == 64 ? 256u : 128u); };   // >= H + 3 H/2, a power of two

template <int H>
size_t smem_for(int nl) {
    return size_t(nl) * 3 * H * H * 2 + 20 * H + 16 * H + size_t(nl) * H * 4 + TILE * 16 + 32;
}

template <int H, bool FIELDS>
__global__ void __launch_bounds__(TC_THREADS, 1)
    k_mlp_deep_tc(const __grid_constant__ MlpConst<H> w, const __grid_constant__ DeepArgs a, const uint8_t* __restrict__ wparts) {
    constexpr int NS = FIELDS ? 3 : 1;
    constexpr int HH = H / 2;                    // hidden units per thread in layer 1 = accumulator columns per thread
    constexpr uint32_t D_COL = 0, A_COL = H;     // A term p: columns A_COL + p * H/2 ...
    constexpr uint32_t LBO = 16 * H, SBO = 128;  // K-neighbour / N-neighbour core matrices of a weight image
    constexpr uint32_t TERM_BYTES = H * H * 2;
    extern __shared__ __align__(128) uint8_t smem[];
    const int nl = a.hidden_layers - 1;
    const size_t w_bytes = size_t(nl) * 3 * TERM_BYTES;
    float2* s_l1 = reinterpret_cast<float2*>(smem + w_bytes);                      // [5][H/2], as deep_kernels.cu
    float4* s_w2 = reinterpret_cast<float4*>(smem + w_bytes + 20 * H);             // {W2[0..3, h]}
    float* s_bh = reinterpret_cast<float*>(smem + w_bytes + 36 * H);               // [nl][H]
    float4* s_part = reinterpret_cast<float4*>(smem + w_bytes + 36 * H + size_t(nl) * H * 4);   // [TILE] upper half's outputs
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_part + TILE);                   // [0] weights landed, [1] layer done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

    const int tid = threadIdx.x, warp = tid >> 5, m = tid & (TILE - 1), half = tid >> 7;
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<TmemCols<H>::value>(tmem_slot);
    for (int q = tid; q < H / 2; q += TC_THREADS) {
        const float4 ra = __ldg(reinterpret_cast<const float4*>(a.W1) + 2 * q), rb = __ldg(reinterpret_cast<const float4*>(a.W1) + 2 * q + 1);
        s_l1[q] = make_float2(__ldg(a.b1 + 2 * q), __ldg(a.b1 + 2 * q + 1));
        s_l1[H / 2 + q] = make_float2(rb.x, ra.x);        // half-swapped for mul2_rn / add2_rn_swapped (mlp_eval.cuh)
        s_l1[2 * (H / 2) + q] = make_float2(rb.y, ra.y);
        s_l1[3 * (H / 2) + q] = make_float2(rb.z, ra.z);
        s_l1[4 * (H / 2) + q] = make_float2(ra.w, rb.w);
    }
    for (int h = tid; h < H; h += TC_THREADS) {
        const float4 v = w.w2[h];                          // {W2[1,h], W2[0,h], W2[3,h], W2[2,h]}
        s_w2[h] = make_float4(v.y, v.x, v.w, v.z);
    }
    for (int i = tid; i < nl * H; i += TC_THREADS) s_bh[i] = __ldg(a.bh + i);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (tid == 0) {
        mbar_expect_tx(&bars[0], uint32_t(w_bytes));
        for (int i = 0; i < nl * 3; ++i) bulk_g2s(smem + size_t(i) * TERM_BYTES, wparts + size_t(i) * TERM_BYTES, TERM_BYTES, &bars[0]);
    }
    mbar_wait(&bars[0], 0);

    const uint32_t tbase = *tmem_slot;
    const uint32_t lane_t = tbase + (uint32_t((warp & 3) * 32) << 16);
    const uint32_t idesc = idesc_bf16_f32(TILE, H);
    const long long n_slab = (long long)(a.z_end - a.z_begin) * a.ny * a.nx;
    const long long tiles = (n_slab + TILE - 1) / TILE;
    const int plane = a.nx * a.ny;
    const size_t n = size_t(n_slab);
    const float4 b2 = w.b2;
    uint32_t phase = 0;

    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long i_pt = tile * TILE + m;
        float cx, cy, cz;
        {
            const long long p = i_pt < n_slab ? i_pt : n_slab - 1;   // tail tile: evaluate a valid point, never stored
            const int zl = int(p / plane), rem = int(p - (long long)zl * plane);
            const int y = rem / a.nx, x = rem - y * a.nx;
            cx = __ldg(a.cxs + x); cy = __ldg(a.cys + y); cz = __ldg(a.czs + a.z_begin + zl);
        }
#pragma unroll 1
        for (int s = 0; s < NS; ++s) {
            const float tcs = FIELDS ? a.tc[s] : a.tc[1];
            // ---- layer 1 (strict fp32, the arithmetic of mlp_eval.cuh): this thread's H/2 hidden units of row m ------------
#pragma unroll 1
            for (int c = 0; c < HH / 16; ++c) {
                uint32_t t1[8], t2[8], t3[8];
                const int q0 = (half * HH + c * 16) / 2;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int q = q0 + j;
                    const float2 b1p = s_l1[q], w0s = s_l1[H / 2 + q], w1s = s_l1[2 * (H / 2) + q], w2s = s_l1[3 * (H / 2) + q];
                    const float2 w3 = s_l1[4 * (H / 2) + q];
                    const float2 pt = make_float2(__fmul_rn(w3.x, tcs), __fmul_rn(w3.y, tcs));
                    const f32x2 sx = add2_rn_swapped(pack2(b1p), mul2_rn(pack2(w0s), bcast2(cx)));
                    const f32x2 sxy = add2_rn_swapped(sx, mul2_rn(pack2(w1s), bcast2(cy)));
                    const f32x2 sxyz = add2_rn_swapped(sxy, mul2_rn(pack2(w2s), bcast2(cz)));
                    float v0, v1;
                    unpack2(add2_rn(sxyz, pack2(pt)), v0, v1);
                    split3(relu_ref(v0), relu_ref(v1), t1[j], t2[j], t3[j]);
                }
                const uint32_t col = lane_t + A_COL + uint32_t(q0);
                tmem_st8(col, t1);
                tmem_st8(col + HH, t2);
                tmem_st8(col + 2 * HH, t3);
            }
            tmem_st_wait();
            fence_before_sync();
            __syncthreads();

            float y0 = 0.f, y1 = 0.f, y2 = 0.f, y3 = 0.f;
#pragma unroll 1
            for (int l = 0; l < nl; ++l) {
                if (tid == 0) {
                    fence_after_sync();
                    const uint32_t wl = smem_u32(smem) + uint32_t(l) * 3 * TERM_BYTES;
                    // t3(a) t1(W), t2 t2, t1 t3, t2 t1, t1 t2, t1 t1: ascending magnitude
                    const int PA[6] = {2, 1, 0, 1, 0, 0}, PB[6] = {0, 1, 2, 0, 1, 0};
                    bool acc = false;
#pragma unroll
                    for (int ps = 0; ps < 6; ++ps) {
#pragma unroll 1
                        for (int ks = 0; ks < H / 16; ++ks) {
                            const uint64_t bd = smem_desc(wl + PB[ps] * TERM_BYTES + ks * 2 * LBO, LBO, SBO);
                            mma_bf16_ts(tbase + D_COL, tbase + A_COL + PA[ps] * HH + ks * 8, bd, idesc, acc);
                            acc = true;
                        }
                    }
                    mma_commit(&bars[1]);
                }
                __syncwarp();
                mbar_wait(&bars[1], phase);
                phase ^= 1;
                __syncwarp();             // tcgen05.ld / .st are warp-collective: leave the polling loop together
                fence_after_sync();
                const bool last = l == nl - 1;
#pragma unroll 1
                for (int c = 0; c < HH / 16; ++c) {
                    const int g0 = half * HH + c * 16;
                    uint32_t r[16];
                    tmem_ld16(lane_t + D_COL + uint32_t(g0), r);
                    tmem_ld_wait();
                    float v[16];
                    const float4* bp = reinterpret_cast<const float4*>(s_bh + l * H + g0);
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        const float4 b = bp[j4];
                        v[4 * j4] = relu_ref(__uint_as_float(r[4 * j4]) + b.x);
                        v[4 * j4 + 1] = relu_ref(__uint_as_float(r[4 * j4 + 1]) + b.y);
                        v[4 * j4 + 2] = relu_ref(__uint_as_float(r[4 * j4 + 2]) + b.z);
                        v[4 * j4 + 3] = relu_ref(__uint_as_float(r[4 * j4 + 3]) + b.w);
                    }
                    if (!last) {
                        uint32_t t1[8], t2[8], t3[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) split3(v[2 * j], v[2 * j + 1], t1[j], t2[j], t3[j]);
                        const uint32_t col = lane_t + A_COL + uint32_t(g0 / 2);
                        tmem_st8(col, t1);
                        tmem_st8(col + HH, t2);
                        tmem_st8(col + 2 * HH, t3);
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float4 o = s_w2[g0 + j];
                            y0 = fmaf(o.x, v[j], y0);
                            y1 = fmaf(o.y, v[j], y1);
                            y2 = fmaf(o.z, v[j], y2);
                            y3 = fmaf(o.w, v[j], y3);
                        }
                    }
                }
                if (!last) tmem_st_wait();
                fence_before_sync();      // this thread's tensor-memory reads / writes are done before the next layer is issued
                __syncthreads();
            }
            // ---- output: lower half adds the upper half's partial sums and stores -------------------------------------------
            if (half == 1) s_part[m] = make_float4(y0, y1, y2, y3);
            __syncthreads();
            if (half == 0 && i_pt < n_slab) {
                const float4 o = s_part[m];
                y0 = (b2.x + y0) + o.x; y1 = (b2.y + y1) + o.y; y2 = (b2.z + y2) + o.z; y3 = (b2.w + y3) + o.w;
                if (FIELDS) {
                    a.sigma[s][i_pt] = y0;
                    a.u[s][i_pt] = y1;
                    a.u[s][n + i_pt] = y2;
                    a.u[s][2 * n + i_pt] = y3;
                } else {
                    a.out_aos[i_pt] = make_float4(y0, y1, y2, y3);
                }
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_free<TmemCols<H>::value>(tbase);
}

template <int H, bool FIELDS>
int launch_t(const void* mlp_const, const DeepArgs& a, const uint8_t* wparts, int grid_blocks, cudaStream_t st) {
    size_t smem = smem_for<H>(a.hidden_layers - 1);
    if (smem < MIN_SMEM) smem = MIN_SMEM;
    cudaError_t e = cudaFuncSetAttribute(k_mlp_deep_tc<H, FIELDS>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return int(e);
    k_mlp_deep_tc<H, FIELDS><<<grid_blocks, TC_THREADS, smem, st>>>(*static_cast<const MlpConst<H>*>(mlp_const), a, wparts);
    return int(cudaGetLastError());
}

uint16_t bf16_rn(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return uint16_t((u >> 16) | 0x40);   // NaN stays NaN
    u += 0x7fffu + ((u >> 16) & 1u);
    return uint16_t(u >> 16);
}
float bf16_to_float(uint16_t b) {
    const uint32_t u = uint32_t(b) << 16;
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

}  // namespace

bool deep_tc_supported(int H, int hidden_layers) {
    if (hidden_layers < 2) return false;
    const int nl = hidden_layers - 1;
    const size_t cap = 227 * 1024;
    switch (H) {
        case 32: return smem_for<32>(nl) <= cap;
        case 64: return smem_for<64>(nl) <= cap;
        case 128: return smem_for<128>(nl) <= cap;
    }
    return false;
}

void deep_tc_pack_layer(int H, const float* W, uint8_t* image) {
    uint16_t* out = reinterpret_cast<uint16_t*>(image);
    const size_t term = size_t(H) * H;
    for (int g = 0; g < H; ++g)          // output = the MMA's N index
        for (int h = 0; h < H; ++h) {    // input = K
            const float v = W[size_t(g) * H + h];
            const uint16_t t1 = bf16_rn(v);
            const float r = v - bf16_to_float(t1);
            const uint16_t t2 = bf16_rn(r);
            const float s = r - bf16_to_float(t2);
            const uint16_t t3 = bf16_rn(s);
            // core matrix (g / 8, h / 8): 8 rows of 8 elements; K-neighbours H/8 core matrices apart
            const size_t off = size_t(h / 8) * (H / 8) * 64 + size_t(g / 8) * 64 + size_t(g % 8) * 8 + size_t(h % 8);
            out[off] = t1;
            out[term + off] = t2;
            out[2 * term + off] = t3;
        }
}

int deep_tc_launch(int H, bool fields, const void* mlp_const, const DeepArgs& a, const uint8_t* wparts, int grid_blocks, cudaStream_t st) {
    if (!deep_tc_supported(H, a.hidden_layers)) return int(cudaErrorInvalidConfiguration);
    switch (H) {
        case 32: return fields ? launch_t<32, true>(mlp_const, a, wparts, grid_blocks, st) : launch_t<32, false>(mlp_const, a, wparts, grid_blocks, st);
        case 64: return fields ? launch_t<64, true>(mlp_const, a, wparts, grid_blocks, st) : launch_t<64, false>(mlp_const, a, wparts, grid_blocks, st);
        case 128: return fields ? launch_t<128, true>(mlp_const, a, wparts, grid_blocks, st) : launch_t<128, false>(mlp_const, a, wparts, grid_blocks, st);
    }
    return int(cudaErrorInvalidValue);
}

}  // namespace physad
