// See dense_kernels.cuh.
#include "dense_kernels.cuh"

namespace physad {

namespace {

constexpr int GT = 64, GK = 16, GPAD = 68;   // tile edge, slab depth, shared row stride (multiple of 4: 128-bit reads)

__global__ void __launch_bounds__(256) k_strict_gemm(const GemmArgs g) {
    __shared__ __align__(16) float As[GK][GPAD];   // As[k][m]
    __shared__ __align__(16) float Bs[GK][GPAD];   // Bs[k][n]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * GT, n0 = blockIdx.x * GT;
    float acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int n = n0 + 4 * tx + j;
        const float v = (g.init != nullptr && n < g.N) ? __ldg(g.init + n) : 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i][j] = v;
    }
    const bool a_k_fast = g.a_ks == 1, b_k_fast = g.b_ks == 1;
    for (int k0 = 0; k0 < g.K; k0 += GK) {
        // stage the slabs; consecutive threads follow the index that is contiguous in memory
        for (int idx = tid; idx < GT * GK; idx += 256) {
            const int m = a_k_fast ? idx / GK : idx % GT, k = a_k_fast ? idx % GK : idx / GT;
            float v = 0.f;
            if (m0 + m < g.M && k0 + k < g.K) v = __ldg(g.A + (long long)(m0 + m) * g.a_ms + (long long)(k0 + k) * g.a_ks);
            As[k][m] = v;
        }
        for (int idx = tid; idx < GT * GK; idx += 256) {
            const int n = b_k_fast ? idx / GK : idx % GT, k = b_k_fast ? idx % GK : idx / GT;
            float v = 0.f;
            if (n0 + n < g.N && k0 + k < g.K)
                v = (n0 + n == g.ones_col) ? 1.f : __ldg(g.B + (long long)(k0 + k) * g.b_ks + (long long)(n0 + n) * g.b_ns);
            Bs[k][n] = v;
        }
        __syncthreads();
        const int kmax = min(GK, g.K - k0);   // never runs past K: a padded k would add +0 and could flip a -0 sum
#pragma unroll 4
        for (int k = 0; k < kmax; ++k) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[k][4 * ty]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][4 * tx]);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = __fadd_rn(acc[i][j], __fmul_rn(av[i], bv[j]));
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + 4 * ty + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + 4 * tx + j;
            if (n >= g.N) continue;
            float v = acc[i][j];
            if (g.epilogue == GEMM_RELU) v = v > 0.f ? v : 0.f;
            else if (g.epilogue == GEMM_SCALED_DIFF) v = __fmul_rn(g.scale, __fsub_rn(v, __ldg(g.aux + (long long)m * g.N + n)));
            else if (g.epilogue == GEMM_MASK) v = __fmul_rn(v, __ldg(g.aux + (long long)m * g.N + n) > 0.f ? 1.f : 0.f);
            if (n < g.n_split) g.C[(long long)m * g.c_ms + n] = v;
            else g.C2[m] = v;
        }
    }
}

}  // namespace

int strict_gemm_launch(const GemmArgs& g, cudaStream_t st) {
    if (g.M <= 0 || g.N <= 0) return 0;
    const dim3 grid(unsigned((g.N + GT - 1) / GT), unsigned((g.M + GT - 1) / GT));
    if (grid.y > 65535u) return int(cudaErrorInvalidConfiguration);
    k_strict_gemm<<<grid, 256, 0, st>>>(g);
    return int(cudaGetLastError());
}

}  // namespace physad
