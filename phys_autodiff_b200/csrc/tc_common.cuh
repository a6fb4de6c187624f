// sm_100a tensor-core plumbing used by deep_tc_kernels.cu (and tools/tc_probe.cu): thin wrappers around the
// tcgen05 / TMEM / mbarrier / bulk-copy PTX, plus the two descriptor encoders.  Nothing here is specific to the
// MLP; the layouts are the "K-major, no swizzle" canonical form: an operand tile is a grid of 8-row x 16-byte
// core matrices (128 contiguous bytes each); SBO is the byte distance between core matrices that are neighbours
// along M/N, LBO between neighbours along K.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace physad {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

// exactly one lane of a CONVERGED warp gets true (the compiler then knows a single lane is active: no broadcast loops)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ---- mbarrier ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol error becomes a trap (a CUDA error on the host), never a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin) {
        if (spin > (1u << 24)) __trap();
#ifdef PHYSAD_TC_BACKOFF_NS
        __nanosleep(PHYSAD_TC_BACKOFF_NS);
#endif
    }
}

// ---- bulk copy global -> shared, completion counted in bytes on an mbarrier ----------------------------------------
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- tensor memory ---------------------------------------------------------------------------------------------
// One warp allocates `cols` (power of two >= 32) columns; the base address lands in shared memory.
template <uint32_t COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_smem) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_smem)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t COLS>
__device__ __forceinline__ void tmem_free(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 32 lanes x 32 bit, 16 consecutive columns: register j of lane i <- TMEM[lane_base + i][col + j]
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                 "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- descriptors -----------------------------------------------------------------------------------------------
// Shared-memory operand, K-major, no swizzle (layout type 0), descriptor version 1 (Blackwell).
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return uint64_t((saddr >> 4) & 0x3fffu) | (uint64_t((lbo_bytes >> 4) & 0x3fffu) << 16) | (uint64_t((sbo_bytes >> 4) & 0x3fffu) << 32) |
           (uint64_t(1) << 46);
}
// Instruction descriptor: D = f32, A = B = bf16, both K-major, dense, M x N.
__host__ __device__ constexpr uint32_t idesc_bf16_f32(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}

// D[tmem] (+)= A[tmem] * B[smem]^T, one K = 16 slice of bf16; issued by ONE thread
__device__ __forceinline__ void mma_bf16_ts(uint32_t d_taddr, uint32_t a_taddr, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_taddr),
        "r"(a_taddr), "l"(b_desc), "r"(idesc), "r"(uint32_t(accumulate))
        : "memory");
}
// the same with the shared-memory descriptor as two words (the low word carries the address: advancing an operand is
// one 32-bit add)
__device__ __forceinline__ void mma_bf16_ts(uint32_t d_taddr, uint32_t a_taddr, uint32_t b_desc_lo, uint32_t b_desc_hi, uint32_t idesc,
                                            bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 bd;\n\t"
        "mov.b64 bd, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], bd, %4, p;\n\t}" ::"r"(d_taddr),
        "r"(a_taddr), "r"(b_desc_lo), "r"(b_desc_hi), "r"(idesc), "r"(uint32_t(accumulate))
        : "memory");
}
// the same with A from shared memory
__device__ __forceinline__ void mma_bf16_ss(uint32_t d_taddr, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_taddr),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(uint32_t(accumulate))
        : "memory");
}
// all MMAs issued so far by this thread -> one arrival on the mbarrier when they have completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- fp32 -> three bf16 terms ------------------------------------------------------------------------------------
// relu(v) = t1 + t2 + t3 EXACTLY: each term is the bf16 TRUNCATION of what the previous ones left (8 significant bits
// each, the remainders are exact in fp32 and stay non-negative), and `.relu` turns a negative input into three zero terms
// (t1 = 0 leaves the negative remainder, which the next conversion clamps again) -- the ReLU costs no instruction.
// Two values at a time as an f32x2 register pair: the packed result holds the low half's term in bits 0..15.
typedef unsigned long long f32pair;
__device__ __forceinline__ uint32_t cvt_bf16x2(float lo, float hi) {   // round to nearest (tools/tc_probe.cu)
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
__device__ __forceinline__ uint32_t cvt_rz_relu_bf16x2(float lo, float hi) {
    uint32_t d;
    asm("cvt.rz.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
__device__ __forceinline__ f32pair sub2(f32pair a, f32pair b) {
    f32pair r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32pair widen_bf16x2(uint32_t t) {   // the two bf16 halves as two floats
    f32pair r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(t << 16), "r"(t & 0xffff0000u));
    return r;
}
__device__ __forceinline__ void split3_relu(f32pair v, uint32_t& t1, uint32_t& t2, uint32_t& t3) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
    t1 = cvt_rz_relu_bf16x2(lo, hi);
    const f32pair r = sub2(v, widen_bf16x2(t1));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r));
    t2 = cvt_rz_relu_bf16x2(lo, hi);
    const f32pair q = sub2(r, widen_bf16x2(t2));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(q));
    t3 = cvt_rz_relu_bf16x2(lo, hi);
}

}  // namespace tc
}  // namespace physad
