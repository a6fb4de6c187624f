// Issue-mix microbenchmark for sm_100a: how packed f32x2 FMAs share an SM sub-partition with the ALU-pipe
// instructions a masked back-propagation needs (FMNMX, FSETP, FSEL).  Each mode runs a fixed group of
// independent instructions per thread in a long unrolled loop; the output is ns per group per SM sub-partition
// normalised to SM cycles with the clock sampled by the caller (1 965 MHz on the pool's B200s under this load),
// so that "cycles per group" can be read against the instruction counts.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o microbench_mix tools/microbench_mix.cu && ./microbench_mix [sm_mhz] [blocks_per_sm]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float sum2(u64 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a + b; }

constexpr int CH = 8, INNER = 32;

template <int MODE>
__global__ void __launch_bounds__(256) k_mix(float* out, int outer, float w0, float w1) {
  u64 acc[CH], xa[CH], xb[CH];
  float s[CH], t[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    acc[c] = pk(threadIdx.x * 1e-3f + c, 1.f + c); xa[c] = pk(w0 + c * 1e-4f, w0 - c * 1e-4f); xb[c] = pk(w1 + c * 1e-5f, w1);
    s[c] = threadIdx.x * 1e-2f - c; t[c] = w1 * c;
  }
  const u64 W0 = pk(w0, w0), W1 = pk(w1, w1);
  for (int o = 0; o < outer; ++o) {
#pragma unroll
    for (int i = 0; i < INNER; ++i) {
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        if (MODE == 0 || MODE == 2 || MODE == 3)   // FFMA2, shared multiplier/addend registers
          asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[c]) : "l"(W0), "l"(W1));
        if (MODE == 1 || MODE == 6 || MODE == 7 || MODE == 9)   // FFMA2, three distinct register pairs
          asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[c]) : "l"(xa[c]), "l"(xb[(c + 3) % CH]));
        // the ALU-pipe instructions read the running accumulator (or, in the ALU-only modes, a value that
        // changes every slot) so that ptxas can neither hoist nor merge them
        float lo, hi;
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[c]));
        if (MODE == 4 || MODE == 5 || MODE == 8) { lo = t[c]; hi = s[(c + 1) % CH]; t[c] = s[c]; }
        if (MODE == 8 || MODE == 9)                // FSET.BF: 1.0f / 0.0f from a comparison
          asm volatile("set.gt.f32.f32 %0, %1, %2;" : "=f"(s[c]) : "f"(lo), "f"(s[c]));
        if (MODE == 2 || MODE == 4 || MODE == 6)   // FMNMX
          if (i & 1) asm volatile("max.f32 %0, %0, %1;" : "+f"(s[c]) : "f"(lo));      // alternate so that two
          else asm volatile("min.f32 %0, %0, %1;" : "+f"(s[c]) : "f"(hi));            // cannot merge into FMNMX3
        if (MODE == 3 || MODE == 5 || MODE == 7) { // FSETP + FSEL
          asm volatile("{ .reg .pred p; setp.gt.f32 p, %1, 0f00000000; selp.f32 %0, %2, %3, p; }"
                       : "=f"(s[c]) : "f"(lo), "f"(hi), "f"(s[c]));
        }
      }
    }
  }
  float r = 0.f;
#pragma unroll
  for (int c = 0; c < CH; ++c) r += sum2(acc[c]) + s[c] + t[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
double run(int sms, int bps, int outer) {
  const int grid = sms * bps;
  float* out; CK(cudaMalloc(&out, sizeof(float) * grid * 256));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int w = 0; w < 2; ++w) k_mix<MODE><<<grid, 256>>>(out, outer, 0.999f, 1e-3f);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 4; ++r) {
    CK(cudaEventRecord(e0));
    k_mix<MODE><<<grid, 256>>>(out, outer, 0.999f, 1e-3f);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  CK(cudaFree(out));
  // groups of CH slots per warp; warps per sub-partition = bps * 8 / 4
  const double groups_per_smsp = double(outer) * INNER * (bps * 8 / 4.0);
  return best * 1e6 / groups_per_smsp;   // ns per group (of CH slots) per sub-partition
}

int main(int argc, char** argv) {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const double mhz = argc > 1 ? atof(argv[1]) : 1965.0;
  const int bps = argc > 2 ? atoi(argv[2]) : 2;
  const int outer = 400, sms = p.multiProcessorCount;
  const char* names[10] = {"ffma2_reuse", "ffma2_3reg", "ffma2_reuse+fmnmx", "ffma2_reuse+fsetp_fsel", "fmnmx", "fsetp_fsel",
                          "ffma2_3reg+fmnmx", "ffma2_3reg+fsetp_fsel", "fset", "ffma2_3reg+fset"};
  double ns[10] = {run<0>(sms, bps, outer), run<1>(sms, bps, outer), run<2>(sms, bps, outer), run<3>(sms, bps, outer),
                  run<4>(sms, bps, outer), run<5>(sms, bps, outer), run<6>(sms, bps, outer), run<7>(sms, bps, outer), run<8>(sms, bps, outer), run<9>(sms, bps, outer)};
  printf("{\"gpu\": \"%s\", \"assumed_sm_mhz\": %.0f, \"warps_per_subpartition\": %d, \"slots_per_group\": %d", p.name, mhz, bps * 2, CH);
  for (int i = 0; i < 10; ++i) printf(", \"%s\": {\"ns_per_group\": %.3f, \"cycles_per_slot\": %.3f}", names[i], ns[i], ns[i] * mhz * 1e-3 / CH);
  printf("}\n");
  return 0;
}
