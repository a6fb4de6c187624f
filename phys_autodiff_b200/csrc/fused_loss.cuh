// The metric kernel: fused coordinate-MLP evaluation + central-difference PDE residual + loss
// reduction for a z-slab of the grid, one launch, nothing but 16 bytes leaves the chip.
//
// Work decomposition
//   work   = tiles x planes "tile-planes", linearised tile-major.  The grid is PERSISTENT: one block per
//            resident slot (SMs x blocks/SM), block b owns the contiguous range [ranges[b], ranges[b+1]) of
//            tile-planes, i.e. at most two z-segments of (usually) two different tiles -- every block
//            gets the same COST (planes + the two time-t-only halo planes of every segment it starts;
//            ranges are computed on the host, capi.cu: FusedPlan), so there is no partial last wave and
//            no block that is slow because its range straddles two tiles.
//   block  = marches along z over each of its segments of one (TX x TY) tile of (x,y) columns.
//   thread = P columns that share x (rows ty, ty+TYB, ...); 32 lanes of a warp = 32 consecutive x.
//   step k = plane zk = segment_begin - 1 + k.  Interior steps evaluate all three time slices for
//            the thread's columns (sharing the layer-1 prefix, see mlp_eval.cuh); the first and the
//            last step are the segment's z-halo and evaluate time t only.
//   halo   = the stencil is the 7-point cross, so besides the tile only the ring of 2(TX+TY)
//            columns around it is needed at time t.  Ring columns are recomputed from coordinates
//            (the MLP is a pure function of (x,y,z,t): SURVEY.md section 8e) in 32-column tasks that
//            rotate over the block's warps from plane to plane so no warp/SMSP is the slow one; the
//            coordinate tables cxs/cys/czs (built on the host with the reference's float expressions)
//            replace every per-point division.
//   fields = time-t outputs of the last four planes live in shared memory ([4 planes][4 ch]
//            [(TY+2) x (TX+2)]); the time differences (y(t+dt) - y(t-dt)) / (2 dt) of the plane whose
//            residual is pending stay in registers.  Planes are separated by a split-phase mbarrier
//            (SPLITBAR: arrive after writing plane k, wait for plane k-1's barrier one plane later) or,
//            in the cross-check variants, one __syncthreads per plane.
//   residual (plane zk-1, formed at step k when plane zk is known): reference src/phys_cpu.cpp:66-109
//            in fp32 (the reference's own CUDA kernels are fp32 too, src/phys_cuda_fused.cu:67-99) or, with
//            DPRES, in double exactly as the CPU reference (bit-identical residuals);
//            squares are accumulated per thread in double exactly like src/phys_cpu.cpp:140-145,
//            then warp shuffle -> block -> per-block partial -> last block sums partials in index order
//            (deterministic for a given launch geometry).
// Out-of-range columns of partial tiles evaluate at the wrapped/clamped coordinate, which is what
// their in-range neighbours need as halo; they just never contribute to the sum.
#pragma once
#include <type_traits>

#include "mlp_eval.cuh"

namespace physad {

// ---- in-kernel all-reduce of the two partial sums over NVLink peer memory ------------------------
// Each rank owns an exchange buffer of XCHG_MAX_RANKS x 2 slots that every peer has mapped (CUDA IPC,
// capi.cu: physad_xchg_*).  The last block of rank r stores its {sum_s, sum_u} into slot [epoch&1][r] of
// EVERY rank's buffer (plain P2P stores, then a system-scope release of the slot's flag = epoch), waits
// until all `world` slots of its own buffer carry this epoch, and adds them in rank order -- so every
// rank ends the kernel with the same, deterministic global sums and no separate collective launch
// (NCCL's latency for a 16-byte all-reduce is a sizeable fraction of a ~0.15 ms slab kernel).
// Two slot sets (epoch parity) suffice: a rank can only be one epoch ahead of the slowest peer,
// because finishing epoch e needs every peer's epoch-e flag.  Ranks are one process per GPU, so the
// kernels that wait on each other run on different devices.
constexpr int XCHG_MAX_RANKS = 8;

struct XSlot {
    double a, b;
    unsigned long long flag;
    unsigned long long pad;
};

struct XchgArgs {
    int rank, world;               // world <= 1: exchange disabled
    unsigned long long epoch;      // same on every rank, > 0, +1 per launch
    XSlot* peer[XCHG_MAX_RANKS];   // peer[p] = rank p's buffer as mapped here (peer[rank] = own)
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_relaxed_sys(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// Called by the first warp of one block with this rank's totals; returns the global totals in (a, b)
// on lane 0 and `true` on every lane if all peers arrived.  Lane p owns peer p: it publishes into p's
// buffer, then ACQUIRES the flag of slot p in its own buffer and reads that slot's payload itself (the
// release/acquire pair orders exactly those accesses); the `world` pairs are then added in rank order through
// shuffles, so every rank forms the same sum.
// Co-residency requirement: the kernels that wait on each other run on DIFFERENT devices (one process per
// GPU), and every rank must launch the same sequence of exchanging kernels -- a rank that is late only delays
// the others, a rank that never launches makes them give up after ~4 s: the sums become NaN (so that a consumer
// who ignores the status cannot use them) and the return value / FusedArgs::status report the timeout.
__device__ __forceinline__ bool xchg_allreduce2(const XchgArgs& x, double& a, double& b) {
    const int lane = threadIdx.x & 31;
    const int set = int(x.epoch & 1ull) * XCHG_MAX_RANKS;
    bool ok = true;
    double pa = 0.0, pb = 0.0;
    if (lane < x.world) {
        XSlot* s = x.peer[lane] + set + x.rank;
        s->a = a;
        s->b = b;
        __threadfence_system();
        st_release_sys(&s->flag, x.epoch);
        const XSlot* r = x.peer[x.rank] + set + lane;
        const long long t0 = clock64();
        while (ld_acquire_sys(&r->flag) != x.epoch) {
            if (clock64() - t0 > 8000000000ll) { ok = false; break; }
            __nanosleep(64);
        }
        pa = ld_relaxed_sys(&r->a);
        pb = ld_relaxed_sys(&r->b);
    }
    ok = __all_sync(0xffffffffu, ok);
    double ta = 0.0, tb = 0.0;
    for (int p = 0; p < x.world; ++p) {   // rank order, identical on every rank
        ta += __shfl_sync(0xffffffffu, pa, p);
        tb += __shfl_sync(0xffffffffu, pb, p);
    }
    a = ok ? ta : __longlong_as_double(0x7ff8000000000000ll);
    b = ok ? tb : __longlong_as_double(0x7ff8000000000000ll);
    return ok;
}

// Result record in MAPPED PINNED HOST memory (capi.cu: physad_ctx::h_res): the last block stores the two sums
// and a status there and then releases `seq` at system scope; the host-buffer entry points poll `seq` instead
// of enqueueing a 16-byte cudaMemcpyAsync and synchronising the stream (~10 us per step, which is 6 % of a
// 0.16 ms slab step on 8 GPUs).
struct HostResult {
    double a, b;
    unsigned long long status;   // 0 = ok, 1 = a peer never arrived at the exchange
    unsigned long long seq;      // written last (st.release.sys)
};

// Where a reduction's totals go besides acc_out (all optional).
struct ReduceSink {
    const XchgArgs* x = nullptr;           // cross-rank exchange before publishing
    HostResult* host = nullptr;            // mapped host record
    unsigned long long host_seq = 0;
    unsigned int* status = nullptr;        // device word, set to 1 when the exchange timed out (sticky)
};

// Optional per-block timeline (diagnostics, tools/trace_fused.py): TRACE_SLOTS x {globaltimer ns, clock64}.
constexpr int TRACE_SLOTS = 18;   // 0 start | 1 prologue done | 2+3s,3+3s,4+3s: segment s start / first halo plane done / end | 14 march done | 15 {smid, tile-planes} | 16 block exit
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned int smid() {
    unsigned int r;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
    return r;
}

struct FusedArgs {
    int nx, ny, nz;          // global grid
    int z_begin, z_end;      // slab [z_begin, z_end)
    int tiles_x, tiles_y;    // tile grid; work = tiles_x*tiles_y*(z_end-z_begin) tile-planes over gridDim.x blocks
    // launch plan (device memory, built once per geometry by capi.cu: FusedPlan)
    const float* cxs;        // [nx] axis coordinates of x indices (same fp32 ops as the reference, done on the host)
    const float* cys;        // [ny]
    const float* czs;        // [nz]
    const int* ranges;       // [gridDim.x + 1] block b owns tile-planes [ranges[b], ranges[b+1]) -- equal COST shares
    int m1p1, periodic;
    float inv2dt, inv2hx, inv2hy, inv2hz;
    double inv2dt_d, inv2hx_d, inv2hy_d, inv2hz_d;  // 1.0/(2.0*double(h)): the exact-residual mode (DPRES)
    double2* partials;       // [gridDim.x]
    unsigned int* ticket;    // zero on entry, zero again on exit
    double* acc_out;         // [2]
    float* R[4];             // slab-local residual outputs or null
    XchgArgs x;              // multi-GPU exchange (world <= 1: off)
    HostResult* host_res;    // mapped pinned host record or null (the *_host entry points)
    unsigned long long host_seq;
    unsigned int* status;    // device word: exchange timeout (sticky) or null
    unsigned long long* trace;  // [gridDim.x][TRACE_SLOTS][2] or null (diagnostics)
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide {a,b} sum -> per-block partial -> grid total written by the last block to finish.
template <int NWARPS>
__device__ __forceinline__ void grid_reduce2_lin(double a, double b, double2* partials, unsigned int* ticket, double* out,
                                                 double2* s_red, unsigned int* s_flag, unsigned int block_lin,
                                                 unsigned int nblocks, const ReduceSink* sink = nullptr) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane == 0) s_red[wid] = make_double2(a, b);
    __syncthreads();
    if (threadIdx.x == 0) {
        double sa = 0.0, sb = 0.0;
#pragma unroll
        for (int i = 0; i < NWARPS; ++i) { sa += s_red[i].x; sb += s_red[i].y; }
        partials[block_lin] = make_double2(sa, sb);
        __threadfence();
        *s_flag = atomicAdd(ticket, 1u);
    }
    __syncthreads();
    if (*s_flag != nblocks - 1) return;
    // last block: fixed-order strided sums, then the same tree
    __threadfence();
    double sa = 0.0, sb = 0.0;
    for (unsigned int i = threadIdx.x; i < nblocks; i += NWARPS * 32) {
        const double2 p = __ldcg(&partials[i]);
        sa += p.x; sb += p.y;
    }
    sa = warp_sum(sa);
    sb = warp_sum(sb);
    __syncthreads();
    if (lane == 0) s_red[wid] = make_double2(sa, sb);
    __syncthreads();
    if (threadIdx.x < 32) {  // first warp: block total, optional cross-rank exchange, publish
        double ta = 0.0, tb = 0.0;
#pragma unroll
        for (int i = 0; i < NWARPS; ++i) { ta += s_red[i].x; tb += s_red[i].y; }
        bool ok = true;
        if (sink != nullptr && sink->x != nullptr && sink->x->world > 1) ok = xchg_allreduce2(*sink->x, ta, tb);
        if (threadIdx.x == 0) {
            out[0] = ta;
            out[1] = tb;
            *ticket = 0u;
            if (sink != nullptr) {
                if (!ok && sink->status) *sink->status = 1u;
                if (sink->host) {
                    sink->host->a = ta;
                    sink->host->b = tb;
                    sink->host->status = ok ? 0ull : 1ull;
                    __threadfence_system();
                    st_release_sys(&sink->host->seq, sink->host_seq);
                }
            }
        }
    }
}

template <int NWARPS>
__device__ __forceinline__ void grid_reduce2(double a, double b, double2* partials, unsigned int* ticket, double* out,
                                             double2* s_red, unsigned int* s_flag, const ReduceSink* sink = nullptr) {
    grid_reduce2_lin<NWARPS>(a, b, partials, ticket, out, s_red, s_flag, blockIdx.x, gridDim.x, sink);
}

// Exchange-only launch for a rank whose slab is empty (more ranks than planes): it still has to
// contribute its zeros and must end up with the global sums.
static __global__ void k_xchg_only(const __grid_constant__ XchgArgs x, double* out, HostResult* host, unsigned long long host_seq,
                            unsigned int* status) {
    double a = 0.0, b = 0.0;
    const bool ok = xchg_allreduce2(x, a, b);
    if (threadIdx.x == 0) {
        out[0] = a; out[1] = b;
        if (!ok && status) *status = 1u;
        if (host) {
            host->a = a; host->b = b; host->status = ok ? 0ull : 1ull;
            __threadfence_system();
            st_release_sys(&host->seq, host_seq);
        }
    }
}

// ---- split-phase block barrier (mbarrier): arrive now, wait one plane later ----------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return unsigned(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}

template <int H, int P, int TYB, int UNROLL, int MINB, bool PACKED, bool SPLITBAR = false, bool DPRES = false>
__global__ void __launch_bounds__(32 * TYB, MINB)
k_fused_mlp_phys_loss(const __grid_constant__ MlpConst<H> w, const FusedArgs a) {
    constexpr int TX = 32, TY = TYB * P, NWARPS = TYB;
    static_assert((NWARPS & (NWARPS - 1)) == 0, "warps per block must be a power of two");
    constexpr int SX = TX + 2, SY = TY + 2, CH = SX * SY, PLANE = 4 * CH, NB = 4;
    constexpr int RING = 2 * TX + 2 * TY, NTASK = (RING + 31) / 32;
    extern __shared__ float smem[];
    float* buf = smem;  // [NB][4][SY][SX]
    __shared__ double2 s_red[NWARPS];
    __shared__ unsigned int s_flag;
    // SPLITBAR: two mbarriers used alternately by plane parity.  A warp ARRIVES after writing plane k and
    // only WAITS for plane k-1's barrier (armed a whole plane earlier, so practically never blocking)
    // before it reads plane k-1's neighbours: the ring-duty warps no longer hold the block up.  With four
    // plane buffers a warp can never overwrite data a slower warp still reads (it cannot be two waits ahead).
    __shared__ unsigned long long s_bar[2];
    // diagnostics: thread 0 stamps slot i of this block's timeline (a null check per segment, nothing per plane)
    auto stamp = [&](int slot) {
        if (a.trace != nullptr && threadIdx.x == 0) {
            unsigned long long* t = a.trace + (size_t(blockIdx.x) * TRACE_SLOTS + slot) * 2;
            t[0] = globaltimer_ns();
            t[1] = (unsigned long long)clock64();
        }
    };
    stamp(0);
    if (SPLITBAR) {
        if (threadIdx.x == 0) { mbar_init(&s_bar[0], NWARPS); mbar_init(&s_bar[1], NWARPS); }
        __syncthreads();
    }

    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // warp id == ty
    const bool per = a.periodic != 0;

    const int nzl = a.z_end - a.z_begin;
    const long long W = (long long)a.tiles_x * a.tiles_y * nzl;
    (void)W;
    const long long w_end = a.ranges[blockIdx.x + 1];
    long long wpos = a.ranges[blockIdx.x];
    double acc_s = 0.0, acc_u = 0.0;
    int kbuf = 0;  // running plane-buffer index (keeps rotating across segments)
    int seg = 0;
    if (a.trace != nullptr && threadIdx.x == 0) {
        unsigned long long* t = a.trace + (size_t(blockIdx.x) * TRACE_SLOTS + 15) * 2;
        t[0] = smid();
        t[1] = (unsigned long long)(w_end - wpos);
    }
    stamp(1);

  while (wpos < w_end) {
    const int tile = int(wpos / nzl);
    const int zs = int(wpos - (long long)tile * nzl);
    const int nplanes = int(w_end - wpos < (long long)(nzl - zs) ? w_end - wpos : (long long)(nzl - zs));
    wpos += nplanes;
    const int zc0 = a.z_begin + zs;
    const int x0 = (tile % a.tiles_x) * TX, y0 = (tile / a.tiles_x) * TY;

    // this thread's columns
    const int gx = x0 + tx;
    const float cx = __ldg(a.cxs + bc_index(gx, a.nx, per));
    float cy[P];
    bool live[P];
#pragma unroll
    for (int j = 0; j < P; ++j) {
        const int gy = y0 + ty + j * TYB;
        cy[j] = __ldg(a.cys + bc_index(gy, a.ny, per));
        live[j] = gx < a.nx && gy < a.ny;
    }
    // DPRES: derivatives and residual sums in double exactly as the CPU reference (bit-identical residuals)
    using real = typename std::conditional<DPRES, double, float>::type;
    const real i2t = DPRES ? real(a.inv2dt_d) : real(a.inv2dt), i2x = DPRES ? real(a.inv2hx_d) : real(a.inv2hx);
    const real i2y = DPRES ? real(a.inv2hy_d) : real(a.inv2hy), i2z = DPRES ? real(a.inv2hz_d) : real(a.inv2hz);
    real dT[P][4];  // time derivative of the plane whose residual is pending
#pragma unroll
    for (int j = 0; j < P; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) dT[j][c] = real(0);

    if (seg < 4) stamp(2 + 3 * seg);
    float cz_next = __ldg(a.czs + bc_index(zc0 - 1, a.nz, per));   // one plane ahead: its L2 latency is off the plane's critical path
    for (int k = 0; k <= nplanes + 1; ++k, ++kbuf) {
        const int zk = zc0 - 1 + k;
        const float cz = cz_next;
        cz_next = __ldg(a.czs + bc_index(zk + 1, a.nz, per));
        if (k == 1 && seg < 4) stamp(3 + 3 * seg);
        float* pl = buf + (kbuf & (NB - 1)) * PLANE;
        const bool halo_plane = (k == 0) || (k == nplanes + 1);
        real dTn[P][4];
        if (halo_plane) {
            float y[P][1][4];
            mlp_eval<H, 1, P, UNROLL, PACKED>(w, cx, cy, cz, y);
#pragma unroll
            for (int j = 0; j < P; ++j) {
                const int o = (1 + ty + j * TYB) * SX + 1 + tx;
#pragma unroll
                for (int c = 0; c < 4; ++c) { pl[c * CH + o] = y[j][0][c]; dTn[j][c] = real(0); }
            }
        } else {
            float y[P][3][4];
            mlp_eval<H, 3, P, UNROLL, PACKED>(w, cx, cy, cz, y);
#pragma unroll
            for (int j = 0; j < P; ++j) {
                const int o = (1 + ty + j * TYB) * SX + 1 + tx;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    pl[c * CH + o] = y[j][1][c];
                    dTn[j][c] = central_diff(y[j][2][c], y[j][0][c], i2t);
                }
            }
            // ring duty for this plane (x/y neighbours of the tile edge): NTASK tasks of 32 ring columns,
            // task i goes to warp (i + kbuf) % NWARPS, i.e. this warp has at most one task per plane and the
            // duty rotates over the warps from plane to plane.  Ring slot -> coordinates are recomputed here
            // (two cached table loads) instead of living in registers for the whole march.
            {
                static_assert(NTASK <= NWARPS, "one ring task per warp and plane");
                const int i = (ty - kbuf) & (NWARPS - 1);
                if (i < NTASK) {
                    const int r = i * 32 + tx;
                    int xx, yy;
                    if (r < TX) { xx = 1 + r; yy = 0; }
                    else if (r < 2 * TX) { xx = 1 + r - TX; yy = SY - 1; }
                    else if (r < 2 * TX + TY) { xx = 0; yy = 1 + r - 2 * TX; }
                    else { xx = SX - 1; yy = 1 + r - 2 * TX - TY; }
                    if (r >= RING) { xx = 0; yy = 0; }  // idle lanes of the last task compute a throw-away column
                    const float rcx = __ldg(a.cxs + bc_index(x0 - 1 + xx, a.nx, per));
                    const float rc[1] = {__ldg(a.cys + bc_index(y0 - 1 + yy, a.ny, per))};
                    float yr[1][1][4];
                    mlp_eval<H, 1, 1, UNROLL, PACKED>(w, rcx, rc, cz, yr);
                    if (r < RING) {
#pragma unroll
                        for (int c = 0; c < 4; ++c) pl[c * CH + yy * SX + xx] = yr[0][0][c];
                    }
                }
            }
        }
        if (SPLITBAR) {
            __syncwarp();
            if (tx == 0) mbar_arrive(&s_bar[kbuf & 1]);
            if (kbuf >= 1) mbar_wait(&s_bar[(kbuf - 1) & 1], unsigned((kbuf - 1) >> 1) & 1u);
        } else {
            __syncthreads();
        }
        if (k >= 2) {
            // residual of plane zk-1: centre/x/y neighbours from plane buffer k-1, z neighbours from k-2 and k
            const float* pc = buf + ((kbuf - 1) & (NB - 1)) * PLANE;
            const float* pm = buf + ((kbuf - 2) & (NB - 1)) * PLANE;
            const float* pp = pl;
            const int zl = zk - 1 - a.z_begin;  // slab-local plane of the residual
#pragma unroll
            for (int j = 0; j < P; ++j) {
                const int o = (1 + ty + j * TYB) * SX + 1 + tx;
                float f[4], R[4];
                real gxv[4], gyv[4], gzv[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float* q = pc + c * CH + o;
                    f[c] = q[0];
                    gxv[c] = central_diff(q[1], q[-1], i2x);
                    gyv[c] = central_diff(q[SX], q[-SX], i2y);
                    gzv[c] = central_diff(pp[c * CH + o], pm[c * CH + o], i2z);
                }
                point_residual(f, gxv, gyv, gzv, dT[j], R);
                const float Rs = R[0], Rx = R[1], Ry = R[2], Rz = R[3];
                if (live[j]) {
                    acc_s += double(Rs) * double(Rs);
                    acc_u += double(Rx) * double(Rx) + double(Ry) * double(Ry) + double(Rz) * double(Rz);
                    if (a.R[0]) {
                        const size_t i = (size_t(zl) * a.ny + (y0 + ty + j * TYB)) * a.nx + gx;
                        a.R[0][i] = Rs; a.R[1][i] = Rx; a.R[2][i] = Ry; a.R[3][i] = Rz;
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < P; ++j)
#pragma unroll
            for (int c = 0; c < 4; ++c) dT[j][c] = dTn[j][c];
    }
    // The next segment starts writing plane buffers that the slowest warp may still be reading
    // (SPLITBAR keeps the one-plane lag across segments, so it needs nothing here).
    if (!SPLITBAR) __syncthreads();
    if (seg < 4) stamp(4 + 3 * seg);
    ++seg;
  }
    stamp(14);
    ReduceSink sink;
    sink.x = &a.x; sink.host = a.host_res; sink.host_seq = a.host_seq; sink.status = a.status;
    grid_reduce2<NWARPS>(acc_s, acc_u, a.partials, a.ticket, a.acc_out, s_red, &s_flag, &sink);
    stamp(16);
}

}  // namespace physad
