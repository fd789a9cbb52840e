// Microbenchmark: tcgen05.ld (TMEM -> registers) and MUFU.EX2 throughput per SM on sm_100a, to decide what bounds
// the attention softmax / dS warps (they read 128 fp32 columns per row per tile and evaluate one exp2 per score).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu && ./tmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}

// mode 0: TMEM loads only; mode 1: exp2 only (64 per iteration per thread); mode 2: both
template <int MODE>
__global__ void __launch_bounds__(256, 2) bench(int iters, float* sink, long long* cycles) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  float x = (float)threadIdx.x * 1e-3f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE != 1) {
      uint32_t r0[32], r1[32], r2[32], r3[32];          // like the softmax warps: four loads in flight, one wait
      ld32(tmem, r0); ld32(tmem + 32, r1); ld32(tmem + 64, r2); ld32(tmem + 96, r3);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += __uint_as_float((r0[0] ^ r1[1] ^ r2[2] ^ r3[3]) & 1u);
    }
    if (MODE != 0) {
#pragma unroll
      for (int j = 0; j < 64; ++j) {
        float y;
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x + (float)j));
        acc += y;
      }
      x -= 1.f;
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(256u) : "memory");
}

template <int MODE> void run(const char* name, int threads, int ctas_per_sm) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = sms * ctas_per_sm, iters = 2000;
  float* sink; long long* cyc;
  cudaMalloc(&sink, 4); cudaMalloc(&cyc, grid * sizeof(long long));
  bench<MODE><<<grid, threads>>>(10, sink, cyc);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  bench<MODE><<<grid, threads>>>(iters, sink, cyc);
  cudaEventRecord(b);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, a, b);
  long long h[1024]; cudaMemcpy(h, cyc, sizeof(long long) * (grid < 1024 ? grid : 1024), cudaMemcpyDeviceToHost);
  const double clk = (double)h[0];
  const double warps = threads / 32.0 * ctas_per_sm;
  const double tmem_bytes = MODE != 1 ? warps * iters * 128.0 * 32 * 4 : 0;   // per SM
  const double exps = MODE != 0 ? warps * 32 * iters * 64.0 : 0;               // per SM
  printf("%-28s threads=%d ctas/sm=%d  %s  %.3f ms  %.0f clk  TMEM %.1f B/clk/SM  EX2 %.2f /clk/SM\n", name, threads, ctas_per_sm,
         e == cudaSuccess ? "ok" : cudaGetErrorString(e), ms, clk, tmem_bytes / clk, exps / clk);
  cudaFree(sink); cudaFree(cyc);
}

int main() {
  run<0>("tmem ld only", 128, 1);
  run<0>("tmem ld only", 128, 2);
  run<0>("tmem ld only", 256, 1);
  run<0>("tmem ld only", 256, 2);
  run<1>("ex2 only", 128, 1);
  run<1>("ex2 only", 128, 2);
  run<1>("ex2 only", 256, 2);
  run<2>("tmem ld + ex2", 128, 2);
  run<2>("tmem ld + ex2", 256, 2);
  return 0;
}
