// Probe for the tcgen05 operand layouts deep_tc_kernels.cu relies on (one CTA, exact small-integer data, host check).
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I phys_autodiff_b200/csrc -o tools/_bin/tc_probe tools/tc_probe.cu
// Each variant computes D[128 x N] = A[128 x K] * B[N x K]^T in bf16 -> fp32 with A either in tensor memory (written by
// tcgen05.st, row = lane, two K elements per 32-bit column) or in shared memory, B in shared memory in the K-major
// no-swizzle canonical layout, and prints the number of mismatching entries.  Variant knobs flip the assumptions one
// at a time, so a wrong guess shows up as "variant 0 fails, variant k passes".
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include "tc_common.cuh"

using namespace physad::tc;

struct Knobs {
    int N, K;
    int a_in_tmem;      // 1: TS form, 0: SS form
    int swap_lbo_sbo;   // descriptor fields exchanged
    int swap_pack;      // odd K element in the low half
};

__host__ __device__ inline int a_val(int m, int k) { return ((m * 3 + k * 5) % 7) - 3; }
__host__ __device__ inline int b_val(int n, int k) { return ((n * 2 + k * 3) % 5) - 2; }

__global__ void __launch_bounds__(128, 1) k_probe(Knobs kn, float* D) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int N = kn.N, K = kn.K, tid = threadIdx.x, warp = tid >> 5;
    __nv_bfloat16* sB = reinterpret_cast<__nv_bfloat16*>(smem);
    __nv_bfloat16* sA = sB + N * K;
    const uint32_t sbo = 128, lboB = (N / 8) * 128, lboA = (128 / 8) * 128;
    for (int i = tid; i < N * K; i += 128) {
        const int n = i / K, k = i % K;
        sB[(k / 8) * (lboB / 2) + (n / 8) * 64 + (n % 8) * 8 + (k % 8)] = __float2bfloat16(float(b_val(n, k)));
    }
    for (int i = tid; i < 128 * K; i += 128) {
        const int m = i / K, k = i % K;
        sA[(k / 8) * (lboA / 2) + (m / 8) * 64 + (m % 8) * 8 + (k % 8)] = __float2bfloat16(float(a_val(m, k)));
    }
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc<512>(&tmem_slot);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tbase = tmem_slot;
    const uint32_t d_col = 0, a_col = 256;
    if (kn.a_in_tmem) {
        // row = tid; K/2 packed columns
        const uint32_t lane_addr = tbase + (uint32_t(warp * 32) << 16) + a_col;
        for (int c0 = 0; c0 < K / 2; c0 += 8) {
            uint32_t r[8];
            for (int j = 0; j < 8; ++j) {
                const int k0 = 2 * (c0 + j);
                const float e = float(a_val(tid, k0)), o = float(a_val(tid, k0 + 1));
                r[j] = kn.swap_pack ? cvt_bf16x2(o, e) : cvt_bf16x2(e, o);
            }
            tmem_st8(lane_addr + c0, r);
        }
        tmem_st_wait();
    }
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
        fence_after_sync();
        const uint32_t idesc = idesc_bf16_f32(128, N);
        for (int ks = 0; ks < K / 16; ++ks) {
            const uint32_t lb = kn.swap_lbo_sbo ? sbo : lboB, sb = kn.swap_lbo_sbo ? lboB : sbo;
            const uint64_t bd = smem_desc(smem_u32(sB) + ks * 2 * lboB, lb, sb);
            if (kn.a_in_tmem) {
                mma_bf16_ts(tbase + d_col, tbase + a_col + ks * 8, bd, idesc, ks > 0);
            } else {
                const uint32_t la = kn.swap_lbo_sbo ? sbo : lboA, sa = kn.swap_lbo_sbo ? lboA : sbo;
                mma_bf16_ss(tbase + d_col, smem_desc(smem_u32(sA) + ks * 2 * lboA, la, sa), bd, idesc, ks > 0);
            }
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tbase + (uint32_t(warp * 32) << 16) + d_col + c0, r);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) D[tid * N + c0 + j] = __uint_as_float(r[j]);
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_free<512>(tbase);
}

int main() {
    const Knobs variants[] = {
        {128, 128, 1, 0, 0}, {128, 128, 1, 1, 0}, {128, 128, 1, 0, 1}, {128, 128, 1, 1, 1}, {128, 128, 0, 0, 0}, {128, 128, 0, 1, 0},
        {64, 64, 1, 0, 0},   {32, 32, 1, 0, 0},   {64, 64, 0, 0, 0},   {128, 16, 1, 0, 0},  {128, 16, 0, 0, 0},
    };
    float* dD;
    cudaMalloc(&dD, 128 * 128 * sizeof(float));
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    for (const Knobs& kn : variants) {
        cudaMemset(dD, 0xff, 128 * 128 * sizeof(float));
        k_probe<<<1, 128, 128 * 1024>>>(kn, dD);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("N=%d K=%d tmemA=%d swapLS=%d swapPack=%d: CUDA error %s\n", kn.N, kn.K, kn.a_in_tmem, kn.swap_lbo_sbo, kn.swap_pack,
                   cudaGetErrorString(e));
            return 1;
        }
        std::vector<float> D(128 * kn.N);
        cudaMemcpy(D.data(), dD, D.size() * sizeof(float), cudaMemcpyDeviceToHost);
        int bad = 0, first = -1;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < kn.N; ++n) {
                float ref = 0;
                for (int k = 0; k < kn.K; ++k) ref += float(a_val(m, k) * b_val(n, k));
                if (D[m * kn.N + n] != ref) {
                    if (first < 0) first = m * kn.N + n;
                    ++bad;
                }
            }
        printf("N=%d K=%d tmemA=%d swapLS=%d swapPack=%d: %d / %d mismatches", kn.N, kn.K, kn.a_in_tmem, kn.swap_lbo_sbo, kn.swap_pack, bad,
               128 * kn.N);
        if (bad) printf("  (first at m=%d n=%d: got %g)", first / kn.N, first % kn.N, D[first]);
        printf("\n");
    }
    cudaFree(dD);
    return 0;
}
