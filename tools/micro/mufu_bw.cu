// Microbenchmark: what bounds the attention softmax loop on sm_100a?  MUFU.EX2 throughput per SM as a function of the
// number of resident warps, with and without the other instructions of the loop (FFMA scale/shift, FADD row sum,
// F2FP bf16x2 pack), each with 4 independent accumulator chains so no dependent-FADD chain limits the rate.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bw mufu_bw.cu && ./mufu_bw
// modes: 0 = ex2 only, 1 = ffma + ex2 + fadd (softmax without the pack), 2 = ffma + ex2 + fadd + bf16x2 pack (full loop),
//        3 = pack only (cvt.rn.bf16x2.f32), 4 = FMA-pipe polynomial exp2 only
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float a, float b) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r; }
__device__ __forceinline__ float exp2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;
  const float r = x - (t - 12582912.0f);
  float p = fmaf(0.0551716685f, r, 0.2426111251f);
  p = fmaf(p, r, 0.6932609677f);
  p = fmaf(p, r, 0.9999280572f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) bench(int iters, float scale, float mb, float* sink, long long* cycles) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  uint32_t pk = 0;
  float x = (float)threadIdx.x * 1e-3f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 64; j += 4) {
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float in = x + (float)(j + e);
        if (MODE == 1 || MODE == 2) in = fmaf(in, scale, -mb);
        if (MODE == 4) v[e] = exp2_poly(in);
        else if (MODE == 3) v[e] = in;
        else v[e] = ex2(in);
      }
      if (MODE == 0 || MODE == 4) { a0 += v[0]; a1 += v[1]; a2 += v[2]; a3 += v[3]; }
      if (MODE == 1 || MODE == 2) { a0 += v[0] + v[1]; a1 += v[2] + v[3]; }
      if (MODE == 2 || MODE == 3) { pk ^= pack(v[0], v[1]); pk ^= pack(v[2], v[3]); }
    }
    x -= 1.f;
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (a0 + a1 + a2 + a3 == 123.456f || pk == 0x12345u) sink[0] = a0;
}

template <int MODE> void run(const char* name, int threads, int ctas_per_sm) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = sms * ctas_per_sm, iters = 4000;
  float* sink; long long* cyc;
  cudaMalloc(&sink, 4); cudaMalloc(&cyc, grid * sizeof(long long));
  bench<MODE><<<grid, threads>>>(10, 1.01f, 0.5f, sink, cyc);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  bench<MODE><<<grid, threads>>>(iters, 1.01f, 0.5f, sink, cyc);
  cudaEventRecord(b);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, a, b);
  long long h[2048]; cudaMemcpy(h, cyc, sizeof(long long) * (grid < 2048 ? grid : 2048), cudaMemcpyDeviceToHost);
  double clk = 0; for (int i = 0; i < grid && i < 2048; ++i) clk = h[i] > clk ? (double)h[i] : clk;
  const double warps = threads / 32.0 * ctas_per_sm;
  const double elems = warps * 32 * (double)iters * 64.0;               // per SM
  printf("%-34s warps/SM=%2.0f (%.0f/SMSP)  %s  %.3f ms  %.0f clk  %.2f elements/clk/SM  (%.2f /clk/SMSP)\n", name, warps, warps / 4,
         e == cudaSuccess ? "ok" : cudaGetErrorString(e), ms, clk, elems / clk, elems / clk / 4);
  cudaFree(sink); cudaFree(cyc);
}

int main() {
  const int cfg[6][2] = {{128, 1}, {256, 1}, {512, 1}, {512, 2}, {512, 3}, {512, 4}};   // 4, 8, 16, 32, 48, 64 warps per SM
  for (auto& c : cfg) run<0>("ex2 only", c[0], c[1]);
  for (auto& c : cfg) run<1>("ffma + ex2 + fadd", c[0], c[1]);
  for (auto& c : cfg) run<2>("ffma + ex2 + fadd + bf16x2 pack", c[0], c[1]);
  for (auto& c : cfg) run<3>("bf16x2 pack only (per 2 elements)", c[0], c[1]);
  for (auto& c : cfg) run<4>("FMA-pipe polynomial exp2", c[0], c[1]);
  return 0;
}
