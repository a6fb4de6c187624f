// FP32 CUDA-core pipe microbenchmark for sm_100a (B200).
// Measures the issue/pipe ceilings the fused MLP+physics kernel is judged against:
//   ffma        : 3-register FFMA chains                     (nominal peak, "fast" mode)
//   strict      : FMUL + FADD pairs (no contraction)         (parity mode: what mlp_cpu.cpp's
//                                                             `s += w*x` means without FMA)
//   strict_mnmx : FMUL+FADD pairs with 1 FMNMX per 8 pairs   (does the ALU pipe co-issue?)
//   ffma2       : packed fma.rn.f32x2
//   strict2     : packed mul.rn.f32x2 + add.rn.f32x2 kept un-contracted (half-swap trick)
//   mul2_add    : FMUL2 products + scalar FADD accumulation
//   strict_lds  : FMUL+FADD with operands fetched by broadcast LDS.128 (1 per 16 math instr)
// Output: one JSON object on stdout; instr/clk/SM from clock64, TFLOP/s from CUDA events.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <string>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int CH = 8;       // independent chains per thread
constexpr int INNER = 64;   // unrolled groups per outer iteration

__device__ __forceinline__ unsigned long long pk(float a, float b) {
  unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r;
}
__device__ __forceinline__ float lo(unsigned long long v) {
  float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a + b;
}

struct Res { double ms, tflops, ipc_sm, mhz; };
struct WBlock { float4 w[64]; };

// mode 7/8: the fused kernel's layer-2 pattern. 7: packed (FMUL2 with a broadcast .F32 activation and a
// uniform-register weight pair, FADD2 on the half-swapped product); 8: the same MACs as scalar FMUL+FADD.
template <int MODE>
__global__ void __launch_bounds__(256) k_l2(const __grid_constant__ WBlock wb, float* out, long long* cyc, int outer) {
  float act[6];
  unsigned long long q[12];
  float y[24];
  #pragma unroll
  for (int c = 0; c < 6; ++c) act[c] = threadIdx.x * 1e-3f + c;
  #pragma unroll
  for (int c = 0; c < 12; ++c) q[c] = pk(float(c), float(c) + 0.5f);
  #pragma unroll
  for (int c = 0; c < 24; ++c) y[c] = float(c);
  long long t0 = clock64();
  for (int o = 0; o < outer; ++o) {
    #pragma unroll 4
    for (int h = 0; h < 64; ++h) {
      const float4 c = wb.w[h];
      if (MODE == 7) {
        const unsigned long long c10 = pk(c.x, c.y), c32 = pk(c.z, c.w);
        #pragma unroll
        for (int j = 0; j < 6; ++j) {
          const unsigned long long aa = pk(act[j], act[j]);
          unsigned long long p; float pl, ph;
          asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(p) : "l"(aa), "l"(c10));
          asm("mov.b64 {%0,%1}, %2;" : "=f"(pl), "=f"(ph) : "l"(p));
          asm("add.rn.f32x2 %0, %1, %2;" : "=l"(q[2 * j]) : "l"(q[2 * j]), "l"(pk(ph, pl)));
          asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(p) : "l"(aa), "l"(c32));
          asm("mov.b64 {%0,%1}, %2;" : "=f"(pl), "=f"(ph) : "l"(p));
          asm("add.rn.f32x2 %0, %1, %2;" : "=l"(q[2 * j + 1]) : "l"(q[2 * j + 1]), "l"(pk(ph, pl)));
        }
      } else {
        #pragma unroll
        for (int j = 0; j < 6; ++j) {
          y[4 * j + 0] = __fadd_rn(y[4 * j + 0], __fmul_rn(c.x, act[j]));
          y[4 * j + 1] = __fadd_rn(y[4 * j + 1], __fmul_rn(c.y, act[j]));
          y[4 * j + 2] = __fadd_rn(y[4 * j + 2], __fmul_rn(c.z, act[j]));
          y[4 * j + 3] = __fadd_rn(y[4 * j + 3], __fmul_rn(c.w, act[j]));
        }
      }
    }
  }
  long long t1 = clock64();
  float s = 0.f;
  #pragma unroll
  for (int c = 0; c < 12; ++c) s += lo(q[c]);
  #pragma unroll
  for (int c = 0; c < 24; ++c) s += y[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
Res run_l2(int sms, int bps, int outer);

template <int MODE>
__global__ void __launch_bounds__(256) k_pipe(float* out, long long* cyc, int outer, float w0, float w1) {
  extern __shared__ float4 sw[];
  float acc[CH];
  unsigned long long acc2[CH];
  #pragma unroll
  for (int c = 0; c < CH; ++c) { acc[c] = threadIdx.x * 1e-3f + c; acc2[c] = pk(acc[c], acc[c] + 1.f); }
  if (MODE == 5) { for (int i = threadIdx.x; i < 64; i += blockDim.x) sw[i] = make_float4(w0, w1, w0 + 1e-3f, w1 - 1e-3f); __syncthreads(); }
  const unsigned long long W0 = pk(w0, w0), W1 = pk(w1, w1);
  long long t0 = clock64();
  for (int o = 0; o < outer; ++o) {
    #pragma unroll
    for (int i = 0; i < INNER; ++i) {
      if (MODE == 0) {
        #pragma unroll
        for (int c = 0; c < CH; ++c) acc[c] = __fmaf_rn(acc[c], w0, w1);
      } else if (MODE == 1) {
        #pragma unroll
        for (int c = 0; c < CH; ++c) acc[c] = __fadd_rn(__fmul_rn(acc[c], w0), w1);
      } else if (MODE == 2) {
        #pragma unroll
        for (int c = 0; c < CH; ++c) acc[c] = __fadd_rn(__fmul_rn(acc[c], w0), w1);
        acc[i % CH] = fmaxf(acc[i % CH], 0.f);
      } else if (MODE == 3) {
        #pragma unroll
        for (int c = 0; c < CH; ++c)
          asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc2[c]) : "l"(W0), "l"(W1));
      } else if (MODE == 4) {
        // ptxas contracts mul.rn.f32x2 -> add.rn.f32x2 into FFMA2 when the product has one use,
        // even with .rn and -fmad=false; feeding the product half-swapped (free .LO_HI operand
        // swizzle in SASS) keeps FMUL2 and FADD2 separate.
        #pragma unroll
        for (int c = 0; c < CH; ++c) {
          unsigned long long p, q; float pl, ph;
          asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(p) : "l"(acc2[c]), "l"(W0));
          asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(pl), "=f"(ph) : "l"(p));
          asm volatile("mov.b64 %0, {%1,%2};" : "=l"(q) : "f"(ph), "f"(pl));
          asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(acc2[c]) : "l"(q), "l"(W1));
        }
      } else if (MODE == 6) {
        // FMUL2 for the products, scalar FADD for the sequential accumulation
        #pragma unroll
        for (int c = 0; c < CH; c += 2) {
          unsigned long long p; float pl, ph;
          asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(p) : "l"(pk(acc[c], acc[c + 1])), "l"(W0));
          asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(pl), "=f"(ph) : "l"(p));
          acc[c] = __fadd_rn(pl, w1); acc[c + 1] = __fadd_rn(ph, w1);
        }
      } else if (MODE == 5) {
        float4 w = sw[(i + o) & 63];
        #pragma unroll
        for (int c = 0; c < CH; c += 2) {
          acc[c]     = __fadd_rn(__fmul_rn(acc[c], w.x), w.y);
          acc[c + 1] = __fadd_rn(__fmul_rn(acc[c + 1], w.z), w.w);
        }
      }
    }
  }
  long long t1 = clock64();
  float s = 0.f;
  #pragma unroll
  for (int c = 0; c < CH; ++c) s += acc[c] + lo(acc2[c]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}


template <int MODE>
Res run(int sms, int blocks_per_sm, int outer, double flops_per_group_per_thread, double instr_per_group) {
  int grid = sms * blocks_per_sm, block = 256;
  float* out; long long* cyc;
  CK(cudaMalloc(&out, sizeof(float) * grid * block));
  CK(cudaMalloc(&cyc, sizeof(long long) * grid));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int w = 0; w < 3; ++w) k_pipe<MODE><<<grid, block, 1024>>>(out, cyc, outer, 0.999f, 1e-3f);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    CK(cudaEventRecord(e0));
    k_pipe<MODE><<<grid, block, 1024>>>(out, cyc, outer, 0.999f, 1e-3f);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  std::vector<long long> h(grid);
  CK(cudaMemcpy(h.data(), cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
  double mc = 0; for (auto v : h) mc += double(v); mc /= grid;
  double groups = double(outer) * INNER;
  double threads = double(grid) * block;
  Res r;
  r.ms = best;
  r.tflops = flops_per_group_per_thread * groups * threads / (best * 1e-3) / 1e12;
  // warp-instructions issued per SM per clock (mean block cycles; blocks_per_sm resident together)
  r.ipc_sm = instr_per_group * groups * (block / 32) * blocks_per_sm / mc;
  r.mhz = mc / (best * 1e-3) / 1e6;
  CK(cudaFree(out)); CK(cudaFree(cyc));
  return r;
}

template <int MODE>
Res run_l2(int sms, int bps, int outer) {
  int grid = sms * bps, block = 256;
  float* out; long long* cyc;
  CK(cudaMalloc(&out, sizeof(float) * grid * block));
  CK(cudaMalloc(&cyc, sizeof(long long) * grid));
  WBlock wb;
  for (int h = 0; h < 64; ++h) wb.w[h] = make_float4(1e-3f * h, -2e-3f * h, 3e-3f, 1e-4f * h);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int w = 0; w < 3; ++w) k_l2<MODE><<<grid, block>>>(wb, out, cyc, outer);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    CK(cudaEventRecord(e0));
    k_l2<MODE><<<grid, block>>>(wb, out, cyc, outer);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  Res r;
  r.ms = best;
  double macs = double(outer) * 64 * 24 * double(grid) * block;   // 24 MACs per h per thread
  r.tflops = 2.0 * macs / (best * 1e-3) / 1e12;
  r.ipc_sm = 0; r.mhz = 0;
  CK(cudaFree(out)); CK(cudaFree(cyc));
  return r;
}

int main(int argc, char** argv) {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  int outer = argc > 1 ? atoi(argv[1]) : 2000;
  int bps = argc > 2 ? atoi(argv[2]) : 4;   // 4 x 256 threads = 32 warps/SM
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"blocks_per_sm\": %d, \"threads_per_block\": 256", p.name, sms, bps);
  struct { const char* name; Res r; const char* note; } rows[9];
  rows[0] = {"ffma",        run<0>(sms, bps, outer, 2.0 * CH, CH),             "flops=2/FFMA"};
  rows[1] = {"strict",      run<1>(sms, bps, outer, 2.0 * CH, 2.0 * CH),       "FMUL+FADD; flops=1/instr"};
  rows[2] = {"strict_mnmx", run<2>(sms, bps, outer, 2.0 * CH, 2.0 * CH + 1),   "FMUL+FADD +1 FMNMX per 16; flops exclude FMNMX"};
  rows[3] = {"ffma2",       run<3>(sms, bps, outer, 4.0 * CH, CH),             "flops=4/FFMA2"};
  rows[4] = {"strict2",     run<4>(sms, bps, outer, 4.0 * CH, 2.0 * CH),       "FMUL2+FADD2; flops=2/instr"};
  rows[5] = {"strict_lds",  run<5>(sms, bps, outer, 2.0 * CH, 2.0 * CH + 1),   "FMUL+FADD + 1 LDS.128 per 16"};
  rows[6] = {"mul2_add",    run<6>(sms, bps, outer, 2.0 * CH, 1.5 * CH),       "FMUL2 + 2 scalar FADD per 2 MACs"};
  rows[7] = {"l2_packed",   run_l2<7>(sms, bps, outer / 8 + 1),                "layer-2 pattern: FMUL2(R.F32 bcast, UR pair)+FADD2(LO_HI); flops=2/MAC"};
  rows[8] = {"l2_scalar",   run_l2<8>(sms, bps, outer / 8 + 1),                "layer-2 pattern: scalar FMUL+FADD with UR operands"};
  for (auto& x : rows)
    printf(", \"%s\": {\"ms\": %.4f, \"tflops\": %.3f, \"warp_instr_per_clk_per_sm\": %.3f, \"sm_mhz_effective\": %.1f, \"note\": \"%s\"}",
           x.name, x.r.ms, x.r.tflops, x.r.ipc_sm, x.r.mhz, x.note);
  printf("}\n");
  return 0;
}
