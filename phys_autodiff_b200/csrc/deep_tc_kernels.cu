// Deeper coordinate MLPs, fast mode: hidden -> hidden layers on tcgen05 with three-term bf16 operands.
// What is computed, and why it is additive / not bit-exact: deep_tc_kernels.cuh.
//
// Two kernels, one persistent block per SM over "row tiles" = 128 points x one time slice (row = (point, slice), 128 rows =
// the M of an MMA):
//
// k_mlp_deep_tc (H = 128, and any shape whose layer images must be streamed): ONE row tile in flight, pipelined across the
// two halves of the hidden width.  WPG = 2 epilogue warps per (group, lane quarter) for H >= 64 (544 threads), 1 for H = 32.
//   * epilogue group g (warps [4 WPG g, 4 WPG (g+1))): thread m owns row m = tensor-memory lane m (a warp may only touch
//     lanes 32 (w % 4) ..); the group produces hidden units [g H/2, (g+1) H/2) in layer 1 and drains the accumulator
//     columns of the same range after every layer, each thread H / (2 WPG) of them;  the last warp: one elected lane issues
//     every MMA (and the weight streaming), the warp owns the tensor memory;
//   * tensor memory columns: [0, H) the fp32 accumulators D (halves D0 | D1), then TWO A operands (activations) of
//     3 bf16 terms x H/2 columns each (two K elements per 32-bit column), written by tcgen05.st from the thread that
//     owns the row: activations never visit shared memory.  Layer t reads A[t & 1], its epilogue writes A[(t+1) & 1];
//   * shared memory: the operand images of the hidden -> hidden layers (3 terms x H x H bf16 each, K-major no-swizzle
//     core matrices, written by deep_tc_pack_layer on the host and bulk-copied; all resident when they fit, else two
//     buffers refilled one layer ahead), layer-1 pairs, output layer, biases;
//   * a layer = four blocks (N half n, K half k) of 6 term pairs x H/32 K slices, issued n0k0 n0k1 | commit full[0] |
//     n1k0 n1k1 | commit full[1].  Block (n, k) of the NEXT layer needs only "group k has written its K half of the
//     other A operand and group n has drained D_n" = mbarrier ready[k] / ready[n], so group 0's epilogue runs under this
//     layer's n1 blocks and group 1's under the next layer's n0k0 block: the tensor pipe only waits when an epilogue
//     takes longer than a quarter of a layer;
//   * after the last hidden layer a thread accumulates its share of the four outputs (fp32 FMA) instead of splitting;
//     the next row tile's layer 1 is computed BEFORE waiting for that last layer (its A operand is free), so it is off
//     the critical path as well.  The last column share of a row collects the others' partial outputs through shared
//     memory and stores.
//
// k_mlp_deep_tc2 (H <= 64, every layer image resident): TWO row tiles in flight, see the comment above it.
//
// Governing roofline: the bf16 tensor pipe at 6 MMA passes per fp32-equivalent contraction: 2 H^2 x 6 flop per row and
// layer against MEASURED_PEAKS' dense bf16 figure.
#include "deep_tc_kernels.cuh"
#include "mlp_eval.cuh"
#include "tc_common.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace physad {

namespace {

using namespace tc;

__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

constexpr int TILE = 128;                       // rows of one MMA
// epilogue warps per (group, lane quarter): each takes 1/WPG of the group's columns.  Two for H >= 64: the epilogues are
// latency-bound with one (measured: tensor pipe waiting 25 % of the time), a thread's share must stay a multiple of 16.
template <int H>
struct Wpg { static constexpr int value = H >= 64 ? 2 : 1; };
template <int H>
struct TcThreads { static constexpr int value = (8 * Wpg<H>::value + 1) * 32; };   // 2 groups x 4 lane quarters x WPG warps + the MMA warp
constexpr size_t MIN_SMEM = 120 * 1024;         // more than half an SM's shared memory: exactly one block per SM, so the
                                                // block's tensor-memory allocation can never wait for a neighbour's

template <int H>
struct TmemCols { static constexpr uint32_t value = H == 128 ? 512u : (H == 64 ? 256u : 128u); };   // = H + 2 * 3 H/2

constexpr size_t SMEM_CAP = 227 * 1024;

// nbuf layer images + the small tables of an nl-layer network
template <int H>
size_t smem_for(int nl, int nbuf) {
    return size_t(nbuf) * 3 * H * H * 2 + 28 * H + 16 * H + size_t(nl) * H * 4 + 2 * 3 * TILE * 16 + 96;
}
// all layers resident if they fit, else two buffers that the MMA warp refills one layer ahead
template <int H>
int weight_buffers(int nl) { return smem_for<H>(nl, nl) <= SMEM_CAP ? nl : 2; }

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int H, bool FIELDS>
// 17 warps: one scheduler hosts five of them, so 16384 / (5 x 32) = 102 -> 96 registers per thread is the most that can
// launch (a __maxnreg__(112) build failed with "too many resources"); the <128, fields> variant spills 24 bytes at that cap.
__global__ void __launch_bounds__(TcThreads<H>::value, 1)
    k_mlp_deep_tc(const __grid_constant__ MlpConst<H> w, const __grid_constant__ DeepArgs a, const uint8_t* __restrict__ wparts,
                  const int nbuf, unsigned long long* __restrict__ prof) {
    constexpr int NS = FIELDS ? 3 : 1;
    constexpr int HH = H / 2;                    // hidden units / accumulator columns per group; columns of one A term
    constexpr uint32_t D_COL = 0, A_COL = H, A_BUF = 3 * HH;   // A operand b, term p: columns A_COL + b A_BUF + p HH ...
    constexpr uint32_t LBO = 16 * H, SBO = 128;  // K-neighbour / N-neighbour core matrices of a weight image
    constexpr uint32_t TERM_BYTES = H * H * 2;
    constexpr int KS_HALF = H / 32;              // K = 16 slices per K half
    constexpr int WPG = Wpg<H>::value, TC_THREADS = TcThreads<H>::value, MMA_WARP = 8 * WPG;
    constexpr int CPT = HH / WPG;                // hidden units / accumulator columns per epilogue thread
    constexpr int NCH = CPT / 16;                // ... in chunks of 16
    static_assert(CPT % 16 == 0, "an epilogue thread drains whole 16-column chunks");
    extern __shared__ __align__(128) uint8_t smem[];
    const int nl = a.hidden_layers - 1;
    const size_t w_bytes = size_t(nbuf) * 3 * TERM_BYTES;
    const bool resident = nbuf >= nl;            // else: layer of step s lives in buffer s & 1
    // layer-1 pairs [7][H/2]: 0..2 W1[., 0..2] half-swapped | 3 b1 | 4..6 fl(W1[., 3] t_s) for the three slices
    float2* s_l1 = reinterpret_cast<float2*>(smem + w_bytes);
    float4* s_w2 = reinterpret_cast<float4*>(smem + w_bytes + 28 * H);             // {W2[0..3, h]}
    float* s_bh = reinterpret_cast<float*>(smem + w_bytes + 44 * H);               // [nl][H]
    float4* s_part = reinterpret_cast<float4*>(smem + w_bytes + 44 * H + size_t(nl) * H * 4);   // [2][3][TILE] partial outputs
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_part + 2 * 3 * TILE);
    uint64_t* bar_w = bars;                      // weights landed
    uint64_t* bar_full = bars + 1;               // [2] D_n complete (tcgen05.commit)
    uint64_t* bar_ready = bars + 3;              // [2] group g: K half g of the next A operand written, D_g drained
    uint64_t* bar_part = bars + 5;               // the other threads' partial outputs of a row are in s_part
    uint64_t* bar_wbuf = bars + 6;               // [2] streaming: the layer image in buffer b has landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    const int tid = threadIdx.x, warp = tid >> 5, m = tid & (TILE - 1);
    const int sub = tid >> 7;                    // which CPT columns of the hidden width (2 WPG = the MMA warp)
    const int grp = sub / WPG;                   // K half / accumulator half this thread works for
    if (tid == 0) {
        mbar_init(bar_w, 1);
        mbar_init(&bar_full[0], 1);
        mbar_init(&bar_full[1], 1);
        mbar_init(&bar_ready[0], TILE * WPG);
        mbar_init(&bar_ready[1], TILE * WPG);
        mbar_init(bar_part, TILE * (2 * WPG - 1));
        mbar_init(&bar_wbuf[0], 1);
        mbar_init(&bar_wbuf[1], 1);
        mbar_fence_init();
    }
    if (warp == MMA_WARP) tmem_alloc<TmemCols<H>::value>(tmem_slot);
    for (int q = tid; q < H / 2; q += TC_THREADS) {
        const float4 ra = __ldg(reinterpret_cast<const float4*>(a.W1) + 2 * q), rb = __ldg(reinterpret_cast<const float4*>(a.W1) + 2 * q + 1);
        s_l1[q] = make_float2(rb.x, ra.x);                // half-swapped for mul2_rn / add2_rn_swapped (mlp_eval.cuh)
        s_l1[H / 2 + q] = make_float2(rb.y, ra.y);
        s_l1[2 * (H / 2) + q] = make_float2(rb.z, ra.z);
        s_l1[3 * (H / 2) + q] = make_float2(__ldg(a.b1 + 2 * q), __ldg(a.b1 + 2 * q + 1));
#pragma unroll
        for (int sl = 0; sl < 3; ++sl)                    // fl(W1[.,3] t_s): the product the strict path rounds once per slice
            s_l1[(4 + sl) * (H / 2) + q] = make_float2(__fmul_rn(ra.w, a.tc[sl]), __fmul_rn(rb.w, a.tc[sl]));
    }
    for (int h = tid; h < H; h += TC_THREADS) {
        const float4 v = w.w2[h];                          // {W2[1,h], W2[0,h], W2[3,h], W2[2,h]}
        s_w2[h] = make_float4(v.y, v.x, v.w, v.z);
    }
    for (int i = tid; i < nl * H; i += TC_THREADS) s_bh[i] = __ldg(a.bh + i);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (tid == 0 && resident) {
        mbar_expect_tx(bar_w, uint32_t(w_bytes));
        for (int i = 0; i < nl * 3; ++i) bulk_g2s(smem + size_t(i) * TERM_BYTES, wparts + size_t(i) * TERM_BYTES, TERM_BYTES, bar_w);
    }

    const uint32_t tbase = *tmem_slot;
    const long long n_slab = (long long)(a.z_end - a.z_begin) * a.ny * a.nx;
    const long long tiles = (n_slab + TILE - 1) / TILE;
    const long long my_tiles = blockIdx.x < tiles ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long my_rts = my_tiles * NS;      // row tiles of this block, in order (tile, slice)

    if (warp == MMA_WARP) {
        // ================= MMA warp: lane 0 issues, in layer order; `step` counts layers over all row tiles =================
        {
            const long long total_steps = my_rts * nl;
            // streaming: the image of step s's layer goes to buffer s & 1 (one elected lane issues, everybody waits)
            auto stream_layer = [&](long long s_) {
                if (elect_one()) {
                    uint64_t* bar = &bar_wbuf[s_ & 1];
                    const uint8_t* src = wparts + size_t(s_ % nl) * 3 * TERM_BYTES;
                    uint8_t* dst = smem + size_t(s_ & 1) * 3 * TERM_BYTES;
                    mbar_expect_tx(bar, 3 * TERM_BYTES);
                    for (int i = 0; i < 3; ++i) bulk_g2s(dst + size_t(i) * TERM_BYTES, src + size_t(i) * TERM_BYTES, TERM_BYTES, bar);
                }
                __syncwarp();
            };
            uint32_t ph_w = 0;        // phase bits of bar_wbuf[0], bar_wbuf[1]
            if (resident) {
                mbar_wait(bar_w, 0);
            } else {
                if (total_steps > 0) stream_layer(0);
                if (total_steps > 1) stream_layer(1);
            }
            const uint32_t idesc = idesc_bf16_f32(TILE, HH);
            const uint32_t w0 = smem_u32(smem);
            // t3(a) t1(W), t2 t2, t1 t3, t2 t1, t1 t2, t1 t1: ascending magnitude
            const int PA[6] = {2, 1, 0, 1, 0, 0}, PB[6] = {0, 1, 2, 0, 1, 0};
            uint32_t ph_ready[2] = {0, 0};
            long long step = 0;
            const long long t_begin = clock64();
            long long t_wait = 0;
            for (long long rt = 0; rt < my_rts; ++rt) {
#pragma unroll 1
                for (int l = 0; l < nl; ++l, ++step) {
                    const uint32_t a_in = tbase + A_COL + uint32_t(step & 1) * A_BUF;
                    const uint32_t wl = w0 + uint32_t(resident ? l : int(step & 1)) * 3 * TERM_BYTES;
                    if (!resident) {
                        mbar_wait(&bar_wbuf[step & 1], (ph_w >> (step & 1)) & 1u);
                        ph_w ^= 1u << (step & 1);
                    }
                    const uint32_t b_lo = (wl >> 4) | ((LBO >> 4) << 16), b_hi = (SBO >> 4) | (1u << 14);   // smem_desc(), in two words
#pragma unroll
                    for (int nh = 0; nh < 2; ++nh) {
#pragma unroll
                        for (int kh = 0; kh < 2; ++kh) {
                            // (n0,k0): ready[0];  (n0,k1): ready[1];  the n1 blocks need nothing new
                            if (nh == 0) {
                                const long long t0 = clock64();
                                mbar_wait(&bar_ready[kh], ph_ready[kh]);
                                t_wait += clock64() - t0;
                                ph_ready[kh] ^= 1;
                                fence_after_sync();
                                // group 1 has drained the previous layer, so every MMA of that layer has completed and
                                // its buffer may be overwritten: fetch the NEXT layer's image into it, 3/4 of a layer ahead
                                if (!resident && kh == 1 && step >= 1 && step + 1 < total_steps) stream_layer(step + 1);
                            }
                            if (elect_one()) {
#pragma unroll
                                for (int ps = 0; ps < 6; ++ps) {
#pragma unroll
                                    for (int ks = 0; ks < KS_HALF; ++ks) {
                                        const uint32_t kslice = uint32_t(kh * KS_HALF + ks);
                                        const uint32_t off = PB[ps] * TERM_BYTES + uint32_t(nh) * (HH / 8) * SBO + kslice * 2 * LBO;
                                        mma_bf16_ts(tbase + D_COL + uint32_t(nh) * HH, a_in + PA[ps] * HH + kslice * 8, b_lo + (off >> 4), b_hi,
                                                    idesc, !(kh == 0 && ps == 0 && ks == 0));
                                    }
                                }
                            }
                            __syncwarp();
                        }
                        if (elect_one()) mma_commit(&bar_full[nh]);
                        __syncwarp();
                    }
                }
            }
            if (prof && blockIdx.x == 0 && tid == MMA_WARP * 32) {
                prof[0] = (unsigned long long)(clock64() - t_begin);
                prof[1] = (unsigned long long)t_wait;
                prof[2] = (unsigned long long)my_rts;
            }
        }
        __syncwarp();   // lanes 1-31 wait here for the issuing lane: the block barrier below is reached as a whole warp
    } else {
        // ================= epilogue groups ==============================================================================
        const uint32_t lane_t = tbase + (uint32_t((warp & 3) * 32) << 16);
        const int plane = a.nx * a.ny;
        const size_t n = size_t(n_slab);
        const float4 b2 = w.b2;
        // layer 1 (strict fp32, the arithmetic of mlp_eval.cuh) of row tile rt: this group's H/2 hidden units of row m,
        // split and stored as K half `grp` of A operand `buf`
        auto layer1 = [&](long long rt, uint32_t buf) {
            const long long tile = blockIdx.x + (rt / NS) * gridDim.x;
            const int sl = int(rt % NS);
            const long long i_pt = tile * TILE + m;
            const long long p = i_pt < n_slab ? i_pt : n_slab - 1;   // tail tile: evaluate a valid point, never stored
            const int zl = int(p / plane), rem = int(p - (long long)zl * plane);
            const int y = rem / a.nx, x = rem - y * a.nx;
            const float cx = __ldg(a.cxs + x), cy = __ldg(a.cys + y), cz = __ldg(a.czs + a.z_begin + zl);
            const float2* bt = s_l1 + (4 + (FIELDS ? sl : 1)) * (H / 2);
            const f32x2 cx2 = bcast2(cx), cy2 = bcast2(cy), cz2 = bcast2(cz);
#pragma unroll 1
            for (int c = 0; c < NCH; ++c) {
                uint32_t t1[8], t2[8], t3[8];
                const int q0 = (sub * CPT + c * 16) / 2;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int q = q0 + j;
                    // the strict path's roundings: (((b1 + W1[.,0] x) + W1[.,1] y) + W1[.,2] z) + fl(W1[.,3] t), products rounded
                    // separately (fed half-swapped so that ptxas cannot contract them, mlp_eval.cuh).  Measured against an fp64
                    // evaluation (tools/deep_tc_truth.py): with fused multiply-adds here the loss error grows from <= 7e-6 to ~1e-5.
                    f32x2 z2 = add2_rn_swapped(pack2(s_l1[3 * (H / 2) + q]), mul2_rn(pack2(s_l1[q]), cx2));
                    z2 = add2_rn_swapped(z2, mul2_rn(pack2(s_l1[H / 2 + q]), cy2));
                    z2 = add2_rn_swapped(z2, mul2_rn(pack2(s_l1[2 * (H / 2) + q]), cz2));
                    z2 = add2_rn(z2, pack2(bt[q]));
                    split3_relu(z2, t1[j], t2[j], t3[j]);
                }
                const uint32_t col = lane_t + A_COL + buf * A_BUF + uint32_t(q0);
                tmem_st8(col, t1);
                tmem_st8(col + HH, t2);
                tmem_st8(col + 2 * HH, t3);
            }
        };
        uint32_t ph_full = 0, ph_part = 0, ph_prev = 0;
        long long step = 0;
        long long t_l1 = 0, t_wfull = 0, t_drain = 0, t_out = 0, t_wprev = 0, t_ld = 0, t_math = 0, t_stw = 0;
        const long long e_begin = clock64();
        if (my_rts > 0) {
            layer1(0, 0);
            tmem_st_wait();
            fence_before_sync();
            mbar_arrive(&bar_ready[grp]);
        }
        for (long long rt = 0; rt < my_rts; ++rt) {
            f32x2 y01 = pack2(0.f, 0.f), y23 = pack2(0.f, 0.f);   // this thread's share of the four outputs
#pragma unroll 1
            for (int l = 0; l < nl; ++l, ++step) {
                const bool last = l == nl - 1;
                const uint32_t out_buf = uint32_t((step + 1) & 1);
                // Group 0 also observes the END of the previous layer (its n1 blocks read the operand written next; group 1
                // has seen it as its own full[1]).  One wait per layer, in order: the barrier can never be two phases ahead,
                // because the layer after this one needs this group's next arrival on ready[0].
                long long c0 = clock64();
                if (grp == 0 && step > 0) {
                    mbar_wait(&bar_full[1], ph_prev);
                    ph_prev ^= 1;
                    __syncwarp();
                }
                long long c1 = clock64();
                t_wprev += c1 - c0;
                // the next row tile's layer 1 goes into the A operand this (last) layer does not read, before the wait
                if (last && rt + 1 < my_rts) layer1(rt + 1, out_buf);
                c0 = clock64();
                t_l1 += c0 - c1;
                mbar_wait(&bar_full[grp], ph_full);
                ph_full ^= 1;
                __syncwarp();             // tcgen05.ld / .st are warp-collective: leave the polling loop together
                fence_after_sync();
                c1 = clock64();
                t_wfull += c1 - c0;
                uint32_t r[NCH][16];      // all of this thread's columns in flight, one wait
#pragma unroll
                for (int c = 0; c < NCH; ++c) tmem_ld16(lane_t + D_COL + uint32_t(sub * CPT + c * 16), r[c]);
                tmem_ld_wait();
                const long long c_ld = clock64();
                t_ld += c_ld - c1;
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const int g0 = sub * CPT + c * 16;
                    const float4* bp = reinterpret_cast<const float4*>(s_bh + l * H + g0);
                    f32x2 v2[8];              // accumulator + bias, two columns per register pair
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        const float4 b = bp[j4];
                        v2[2 * j4] = add2_rn(pack2(__uint_as_float(r[c][4 * j4]), __uint_as_float(r[c][4 * j4 + 1])), pack2(b.x, b.y));
                        v2[2 * j4 + 1] = add2_rn(pack2(__uint_as_float(r[c][4 * j4 + 2]), __uint_as_float(r[c][4 * j4 + 3])), pack2(b.z, b.w));
                    }
                    if (!last) {
                        uint32_t t1[8], t2[8], t3[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) split3_relu(v2[j], t1[j], t2[j], t3[j]);
                        const uint32_t col = lane_t + A_COL + out_buf * A_BUF + uint32_t(g0 / 2);
                        tmem_st8(col, t1);
                        tmem_st8(col + HH, t2);
                        tmem_st8(col + 2 * HH, t3);
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float va, vb;
                            unpack2(v2[j], va, vb);
                            const float4 oa = s_w2[g0 + 2 * j], ob = s_w2[g0 + 2 * j + 1];
                            const f32x2 aa = bcast2(relu_ref(va)), ab = bcast2(relu_ref(vb));
                            y01 = fma2(pack2(oa.x, oa.y), aa, y01);
                            y23 = fma2(pack2(oa.z, oa.w), aa, y23);
                            y01 = fma2(pack2(ob.x, ob.y), ab, y01);
                            y23 = fma2(pack2(ob.z, ob.w), ab, y23);
                        }
                    }
                }
                const long long c_st = clock64();
                t_math += c_st - c_ld;
                tmem_st_wait();
                t_stw += clock64() - c_st;
                fence_before_sync();      // this thread's tensor-memory reads and writes are done: D_grp and the K half may be reused / read
                if (last && sub == 2 * WPG - 1) {   // before the arrival that lets the next row tile run: bar_part can never be two phases ahead
                    mbar_wait(bar_part, ph_part);
                    ph_part ^= 1;
                }
                mbar_arrive(&bar_ready[grp]);
                t_drain += clock64() - c1;
            }
            const long long c2 = clock64();
            // ---- outputs: the last column share of a row collects the other shares' partial sums and stores -----------------
            float4* part = s_part + (rt & 1) * 3 * TILE;
            float y0, y1, y2, y3;
            unpack2(y01, y0, y1);
            unpack2(y23, y2, y3);
            if (sub != 2 * WPG - 1) {
                part[sub * TILE + m] = make_float4(y0, y1, y2, y3);
                mbar_arrive(bar_part);
            } else {
                const long long tile = blockIdx.x + (rt / NS) * gridDim.x;
                const int sl = int(rt % NS);
                const long long i_pt = tile * TILE + m;
                if (i_pt < n_slab) {
                    float4 o = b2;
#pragma unroll
                    for (int k = 0; k < 2 * WPG - 1; ++k) {
                        const float4 pk = part[k * TILE + m];
                        o.x += pk.x; o.y += pk.y; o.z += pk.z; o.w += pk.w;
                    }
                    y0 += o.x; y1 += o.y; y2 += o.z; y3 += o.w;
                    if (FIELDS) {
                        a.sigma[sl][i_pt] = y0;
                        a.u[sl][i_pt] = y1;
                        a.u[sl][n + i_pt] = y2;
                        a.u[sl][2 * n + i_pt] = y3;
                    } else {
                        a.out_aos[i_pt] = make_float4(y0, y1, y2, y3);
                    }
                }
            }
            t_out += clock64() - c2;
        }
        if (prof && blockIdx.x == 0 && m == 0 && sub % WPG == 0) {
            unsigned long long* q = prof + 4 + grp * 8;
            q[0] = (unsigned long long)(clock64() - e_begin);
            q[1] = (unsigned long long)t_wprev; q[2] = (unsigned long long)t_l1; q[3] = (unsigned long long)t_wfull;
            q[4] = (unsigned long long)t_drain; q[5] = (unsigned long long)t_out;
            if (grp == 0) { prof[24] = (unsigned long long)t_ld; prof[25] = (unsigned long long)t_math; prof[26] = (unsigned long long)t_stw; }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == MMA_WARP) tmem_free<TmemCols<H>::value>(tbase);
}

// ---- H <= 64: TWO row tiles in flight ------------------------------------------------------------------------------------
// At H <= 64 a tile needs only H + 3 H/2 <= 160 tensor-memory columns and the kernel above is bound by its CUDA-core
// epilogue (the split and drain work is O(H) per row and layer, the MMA work O(H^2)), whose threads wait a quarter of the time
// for "their" MMAs.  Here two row tiles ("slots") alternate instead: while the tensor pipe runs slot 0's layer, slot 1's
// eight epilogue warps drain, split and refill slot 1's A operand, and vice versa -- so a layer is plain full-width MMAs
// (N = H, no N/K halves, ONE A operand per slot: its epilogue and its MMAs never overlap), 24 instead of 48 per layer.
// Used when every layer image is resident in shared memory; the streamed case keeps the kernel above.
constexpr int TC2_THREADS = (16 + 1) * 32;       // 2 slots x 4 lane quarters x 2 column halves + the MMA warp

template <int H>
size_t smem_for2(int nl) {
    return size_t(nl) * 3 * H * H * 2 + 28 * H + 16 * H + size_t(nl) * H * 4 + 2 * 2 * TILE * 16 + 96;
}

template <int H, bool FIELDS>
__global__ void __launch_bounds__(TC2_THREADS, 1)
    k_mlp_deep_tc2(const __grid_constant__ MlpConst<H> w, const __grid_constant__ DeepArgs a, const uint8_t* __restrict__ wparts) {
    static_assert(H == 32 || H == 64, "two tiles of H + 3 H/2 columns each");
    constexpr int NS = FIELDS ? 3 : 1;
    constexpr int HH = H / 2;                    // hidden units / accumulator columns per epilogue thread; columns of one A term
    constexpr uint32_t SLOT_COLS = H + 3 * HH;   // D | A term 1 | term 2 | term 3
    constexpr uint32_t TMEM_COLS = H == 64 ? 512u : 256u;
    constexpr uint32_t LBO = 16 * H, SBO = 128;
    constexpr uint32_t TERM_BYTES = H * H * 2;
    constexpr int NCH = HH / 16;
    constexpr int MMA_WARP = 16;
    extern __shared__ __align__(128) uint8_t smem[];
    const int nl = a.hidden_layers - 1;
    const size_t w_bytes = size_t(nl) * 3 * TERM_BYTES;
    // layer-1 pairs [7][H/2]: 0..2 W1[., 0..2] half-swapped | 3 b1 | 4..6 fl(W1[., 3] t_s) for the three slices
    float2* s_l1 = reinterpret_cast<float2*>(smem + w_bytes);
    float4* s_w2 = reinterpret_cast<float4*>(smem + w_bytes + 28 * H);             // {W2[0..3, h]}
    float* s_bh = reinterpret_cast<float*>(smem + w_bytes + 44 * H);               // [nl][H]
    float4* s_part = reinterpret_cast<float4*>(smem + w_bytes + 44 * H + size_t(nl) * H * 4);   // [slot][2][TILE] lower half's outputs
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_part + 2 * 2 * TILE);
    uint64_t* bar_w = bars;                      // layer images landed
    uint64_t* bar_full = bars + 1;               // [slot] the slot's layer is complete (tcgen05.commit)
    uint64_t* bar_ready = bars + 3;              // [slot] the slot's A operand is written and its D drained (256 arrivals)
    uint64_t* bar_part = bars + 5;               // [slot] the lower column half's partial outputs are in s_part (128 arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

    const int tid = threadIdx.x, warp = tid >> 5, m = tid & (TILE - 1);
    const int slot = tid >> 8, half = (tid >> 7) & 1;
    if (tid == 0) {
        mbar_init(bar_w, 1);
        for (int s_ = 0; s_ < 2; ++s_) {
            mbar_init(&bar_full[s_], 1);
            mbar_init(&bar_ready[s_], 2 * TILE);
            mbar_init(&bar_part[s_], TILE);
        }
        mbar_fence_init();
    }
    if (warp == MMA_WARP) tmem_alloc<TMEM_COLS>(tmem_slot);
    for (int q = tid; q < H / 2; q += TC2_THREADS) {
        const float4 ra = __ldg(reinterpret_cast<const float4*>(a.W1) + 2 * q), rb = __ldg(reinterpret_cast<const float4*>(a.W1) + 2 * q + 1);
        s_l1[q] = make_float2(rb.x, ra.x);                // half-swapped for mul2_rn / add2_rn_swapped (mlp_eval.cuh)
        s_l1[H / 2 + q] = make_float2(rb.y, ra.y);
        s_l1[2 * (H / 2) + q] = make_float2(rb.z, ra.z);
        s_l1[3 * (H / 2) + q] = make_float2(__ldg(a.b1 + 2 * q), __ldg(a.b1 + 2 * q + 1));
#pragma unroll
        for (int sl = 0; sl < 3; ++sl)                    // fl(W1[.,3] t_s): the product the strict path rounds once per slice
            s_l1[(4 + sl) * (H / 2) + q] = make_float2(__fmul_rn(ra.w, a.tc[sl]), __fmul_rn(rb.w, a.tc[sl]));
    }
    for (int h = tid; h < H; h += TC2_THREADS) {
        const float4 v = w.w2[h];                          // {W2[1,h], W2[0,h], W2[3,h], W2[2,h]}
        s_w2[h] = make_float4(v.y, v.x, v.w, v.z);
    }
    for (int i = tid; i < nl * H; i += TC2_THREADS) s_bh[i] = __ldg(a.bh + i);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (tid == 0) {
        mbar_expect_tx(bar_w, uint32_t(w_bytes));
        for (int i = 0; i < nl * 3; ++i) bulk_g2s(smem + size_t(i) * TERM_BYTES, wparts + size_t(i) * TERM_BYTES, TERM_BYTES, bar_w);
    }

    const uint32_t tbase = *tmem_slot;
    const long long n_slab = (long long)(a.z_end - a.z_begin) * a.ny * a.nx;
    const long long tiles = (n_slab + TILE - 1) / TILE;
    const long long my_tiles = blockIdx.x < tiles ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long my_rts = my_tiles * NS;      // row tiles of this block in order (tile, slice); slot s takes s, s + 2, ...

    if (warp == MMA_WARP) {
        mbar_wait(bar_w, 0);
        const uint32_t idesc = idesc_bf16_f32(TILE, H);
        const uint32_t w0 = smem_u32(smem);
        const uint32_t b_hi = (SBO >> 4) | (1u << 14);
        const int PA[6] = {2, 1, 0, 1, 0, 0}, PB[6] = {0, 1, 2, 0, 1, 0};   // t3 t1, t2 t2, t1 t3, t2 t1, t1 t2, t1 t1
        uint32_t ph_ready = 0;                   // phase bits of bar_ready[0], [1]
        for (long long rt0 = 0; rt0 < my_rts; rt0 += 2) {
#pragma unroll 1
            for (int l = 0; l < nl; ++l) {
                const uint32_t b_lo = ((w0 + uint32_t(l) * 3 * TERM_BYTES) >> 4) | ((LBO >> 4) << 16);
#pragma unroll
                for (int s_ = 0; s_ < 2; ++s_) {
                    if (rt0 + s_ < my_rts) {
                        mbar_wait(&bar_ready[s_], (ph_ready >> s_) & 1u);
                        ph_ready ^= 1u << s_;
                        fence_after_sync();
                        const uint32_t d_t = tbase + uint32_t(s_) * SLOT_COLS, a_t = d_t + H;
                        if (elect_one()) {
#pragma unroll
                            for (int ps = 0; ps < 6; ++ps) {
#pragma unroll
                                for (int ks = 0; ks < H / 16; ++ks)
                                    mma_bf16_ts(d_t, a_t + PA[ps] * HH + ks * 8, b_lo + ((PB[ps] * TERM_BYTES + uint32_t(ks) * 2 * LBO) >> 4), b_hi,
                                                idesc, !(ps == 0 && ks == 0));
                            }
                            mma_commit(&bar_full[s_]);
                        }
                        __syncwarp();
                    }
                }
            }
        }
        __syncwarp();
    } else {
        const uint32_t lane_t = tbase + (uint32_t((warp & 3) * 32) << 16) + uint32_t(slot) * SLOT_COLS;   // this slot's D, row m
        const uint32_t a_t = lane_t + H;
        const int plane = a.nx * a.ny;
        const size_t n = size_t(n_slab);
        const float4 b2 = w.b2;
        float4* part = s_part + slot * 2 * TILE;
        uint32_t ph_full = 0, ph_part = 0;
        long long k_rt = 0;                      // how many row tiles this slot has finished (parity picks the s_part buffer)
        for (long long rt = slot; rt < my_rts; rt += 2, ++k_rt) {
            const long long tile = blockIdx.x + (rt / NS) * gridDim.x;
            const int sl = int(rt % NS);
            const long long i_pt = tile * TILE + m;
            // ---- layer 1 (strict fp32, the arithmetic of mlp_eval.cuh): this thread's H/2 hidden units of row m -> A -------------
            {
                const long long p = i_pt < n_slab ? i_pt : n_slab - 1;   // tail tile: evaluate a valid point, never stored
                const int zl = int(p / plane), rem = int(p - (long long)zl * plane);
                const int y = rem / a.nx, x = rem - y * a.nx;
                const f32x2 cx2 = bcast2(__ldg(a.cxs + x)), cy2 = bcast2(__ldg(a.cys + y)), cz2 = bcast2(__ldg(a.czs + a.z_begin + zl));
                const float2* bt = s_l1 + (4 + (FIELDS ? sl : 1)) * (H / 2);
#pragma unroll 1
                for (int c = 0; c < NCH; ++c) {
                    uint32_t t1[8], t2[8], t3[8];
                    const int q0 = (half * HH + c * 16) / 2;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int q = q0 + j;
                        f32x2 z2 = add2_rn_swapped(pack2(s_l1[3 * (H / 2) + q]), mul2_rn(pack2(s_l1[q]), cx2));
                        z2 = add2_rn_swapped(z2, mul2_rn(pack2(s_l1[H / 2 + q]), cy2));
                        z2 = add2_rn_swapped(z2, mul2_rn(pack2(s_l1[2 * (H / 2) + q]), cz2));
                        z2 = add2_rn(z2, pack2(bt[q]));
                        split3_relu(z2, t1[j], t2[j], t3[j]);
                    }
                    tmem_st8(a_t + uint32_t(q0), t1);
                    tmem_st8(a_t + HH + uint32_t(q0), t2);
                    tmem_st8(a_t + 2 * HH + uint32_t(q0), t3);
                }
                tmem_st_wait();
                fence_before_sync();
                mbar_arrive(&bar_ready[slot]);
            }
            f32x2 y01 = pack2(0.f, 0.f), y23 = pack2(0.f, 0.f);
#pragma unroll 1
            for (int l = 0; l < nl; ++l) {
                const bool last = l == nl - 1;
                mbar_wait(&bar_full[slot], ph_full);
                ph_full ^= 1;
                __syncwarp();             // tcgen05.ld / .st are warp-collective
                fence_after_sync();
                uint32_t r[NCH][16];
#pragma unroll
                for (int c = 0; c < NCH; ++c) tmem_ld16(lane_t + uint32_t(half * HH + c * 16), r[c]);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const int g0 = half * HH + c * 16;
                    const float4* bp = reinterpret_cast<const float4*>(s_bh + l * H + g0);
                    f32x2 v2[8];
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        const float4 b = bp[j4];
                        v2[2 * j4] = add2_rn(pack2(__uint_as_float(r[c][4 * j4]), __uint_as_float(r[c][4 * j4 + 1])), pack2(b.x, b.y));
                        v2[2 * j4 + 1] = add2_rn(pack2(__uint_as_float(r[c][4 * j4 + 2]), __uint_as_float(r[c][4 * j4 + 3])), pack2(b.z, b.w));
                    }
                    if (!last) {
                        // the slot's MMAs have completed (full), so its ONE A operand can be overwritten in place
                        uint32_t t1[8], t2[8], t3[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) split3_relu(v2[j], t1[j], t2[j], t3[j]);
                        tmem_st8(a_t + uint32_t(g0 / 2), t1);
                        tmem_st8(a_t + HH + uint32_t(g0 / 2), t2);
                        tmem_st8(a_t + 2 * HH + uint32_t(g0 / 2), t3);
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float va, vb;
                            unpack2(v2[j], va, vb);
                            const float4 oa = s_w2[g0 + 2 * j], ob = s_w2[g0 + 2 * j + 1];
                            const f32x2 aa = bcast2(relu_ref(va)), ab = bcast2(relu_ref(vb));
                            y01 = fma2(pack2(oa.x, oa.y), aa, y01);
                            y23 = fma2(pack2(oa.z, oa.w), aa, y23);
                            y01 = fma2(pack2(ob.x, ob.y), ab, y01);
                            y23 = fma2(pack2(ob.z, ob.w), ab, y23);
                        }
                    }
                }
                if (!last) {
                    tmem_st_wait();
                    fence_before_sync();
                    mbar_arrive(&bar_ready[slot]);
                }
            }
            // ---- outputs: the lower column half hands its partial sums to the upper one, which stores --------------------------------
            float y0, y1, y2, y3;
            unpack2(y01, y0, y1);
            unpack2(y23, y2, y3);
            float4* pbuf = part + (k_rt & 1) * TILE;
            if (half == 0) {
                pbuf[m] = make_float4(y0, y1, y2, y3);
                mbar_arrive(&bar_part[slot]);
            } else {
                // waited for BEFORE this thread's next arrival on ready (the next row tile's layer 1): the lower half cannot be
                // two row tiles ahead, so the barrier is never two phases ahead and the two s_part buffers suffice
                mbar_wait(&bar_part[slot], ph_part);
                ph_part ^= 1;
                if (i_pt < n_slab) {
                    const float4 o = pbuf[m];
                    y0 = (b2.x + o.x) + y0; y1 = (b2.y + o.y) + y1; y2 = (b2.z + o.z) + y2; y3 = (b2.w + o.w) + y3;
                    if (FIELDS) {
                        a.sigma[sl][i_pt] = y0;
                        a.u[sl][i_pt] = y1;
                        a.u[sl][n + i_pt] = y2;
                        a.u[sl][2 * n + i_pt] = y3;
                    } else {
                        a.out_aos[i_pt] = make_float4(y0, y1, y2, y3);
                    }
                }
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == MMA_WARP) tmem_free<TMEM_COLS>(tbase);
}

template <int H, bool FIELDS>
int launch_t(const void* mlp_const, const DeepArgs& a, const uint8_t* wparts, int grid_blocks, cudaStream_t st) {
    const int nl = a.hidden_layers - 1, nbuf = weight_buffers<H>(nl);
    if constexpr (H <= 64) {
        // every layer image resident: two row tiles in flight (k_mlp_deep_tc2); PHYSAD_DEEP_TC_ONE_TILE keeps the kernel above
        static const bool one_tile = getenv("PHYSAD_DEEP_TC_ONE_TILE") != nullptr;
        size_t smem2 = smem_for2<H>(nl);
        if (!one_tile && smem2 <= SMEM_CAP) {
            if (smem2 < MIN_SMEM) smem2 = MIN_SMEM;
            cudaError_t e2 = cudaFuncSetAttribute(k_mlp_deep_tc2<H, FIELDS>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem2));
            if (e2 != cudaSuccess) return int(e2);
            k_mlp_deep_tc2<H, FIELDS><<<grid_blocks, TC2_THREADS, smem2, st>>>(*static_cast<const MlpConst<H>*>(mlp_const), a, wparts);
            return int(cudaGetLastError());
        }
    }
    size_t smem = smem_for<H>(nl, nbuf);
    if (smem < MIN_SMEM) smem = MIN_SMEM;
    cudaError_t e = cudaFuncSetAttribute(k_mlp_deep_tc<H, FIELDS>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return int(e);
    // development aid: PHYSAD_DEEP_TC_PROF=1 prints block 0's cycle counters (MMA thread: total / waiting for the epilogues;
    // thread 0 of each epilogue group: total / waiting / layer 1 / waiting for the MMAs / draining / output) after a sync
    static const bool want_prof = getenv("PHYSAD_DEEP_TC_PROF") != nullptr;
    static unsigned long long* d_prof = nullptr;
    if (want_prof && !d_prof) {
        cudaMalloc(&d_prof, 32 * sizeof(unsigned long long));
    }
    if (want_prof) cudaMemsetAsync(d_prof, 0, 32 * sizeof(unsigned long long), st);
    k_mlp_deep_tc<H, FIELDS><<<grid_blocks, TcThreads<H>::value, smem, st>>>(*static_cast<const MlpConst<H>*>(mlp_const), a, wparts, nbuf,
                                                                             want_prof ? d_prof : nullptr);
    cudaError_t le = cudaGetLastError();
    if (want_prof && le == cudaSuccess) {
        unsigned long long h[32];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, d_prof, sizeof(h), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[deep_tc H=%d L=%d] mma: total %llu wait_ready %llu row_tiles %llu | g0: total %llu wprev %llu l1 %llu wfull %llu drain %llu out %llu"
                        " | g1: total %llu wprev %llu l1 %llu wfull %llu drain %llu out %llu | g0 drain = ld %llu + math %llu + st_wait %llu + rest\n",
                H, a.hidden_layers, h[0], h[1], h[2], h[4], h[5], h[6], h[7], h[8], h[9], h[12], h[13], h[14], h[15], h[16], h[17], h[24], h[25], h[26]);
    }
    return int(le);
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
    if (hidden_layers < 2 || hidden_layers > 16) return false;
    const int nl = hidden_layers - 1;
    switch (H) {
        case 32: return smem_for<32>(nl, weight_buffers<32>(nl)) <= SMEM_CAP;
        case 64: return smem_for<64>(nl, weight_buffers<64>(nl)) <= SMEM_CAP;
        case 128: return smem_for<128>(nl, weight_buffers<128>(nl)) <= SMEM_CAP;
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
