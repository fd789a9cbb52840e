// Shared helpers for the sm_100a kernels: error plumbing, dtype load/store, row maps.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/avjepa_b200.h"

typedef __nv_bfloat16 bf16;

// ---- error plumbing (thread-local last error; no exceptions cross the C ABI) ----------
void avj_set_error(const char* fmt, ...);

#define AVJ_CHECK(cond, ...)                                                     \
  do {                                                                           \
    if (!(cond)) {                                                               \
      avj_set_error(__VA_ARGS__);                                                \
      return 1;                                                                  \
    }                                                                            \
  } while (0)

#define AVJ_CUDA(expr)                                                           \
  do {                                                                           \
    cudaError_t e__ = (expr);                                                    \
    if (e__ != cudaSuccess) {                                                    \
      avj_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),     \
                    __FILE__, __LINE__);                                         \
      return 2;                                                                  \
    }                                                                            \
  } while (0)

// every kernel launch of the library is followed by AVJ_LAUNCH_CHECK(): it also feeds the launch counter that
// avj_launch_count() reports (bench.py's `gpu_launches` is this counter's difference over the timed region).
extern unsigned long long g_avj_launches;
#define AVJ_COUNT_LAUNCH() (void)__atomic_add_fetch(&g_avj_launches, 1ull, __ATOMIC_RELAXED)

#define AVJ_LAUNCH_CHECK()                                                       \
  do {                                                                           \
    AVJ_COUNT_LAUNCH();                                                          \
    cudaError_t e__ = cudaGetLastError();                                        \
    if (e__ != cudaSuccess) {                                                    \
      avj_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), \
                    __FILE__, __LINE__);                                         \
      return 3;                                                                  \
    }                                                                            \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

int avj_num_sms();

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------
// A step is ~2400 dependent launches of 10-400 us kernels, so the per-boundary cost (grid drain, launch
// processing, the next kernel's barrier-init / TMEM-alloc / descriptor-fetch prologue) is a visible share of
// the step.  Kernels launched through avj_launch_pdl carry cudaLaunchAttributeProgrammaticStreamSerialization:
// they may be scheduled while their predecessor in the stream is still running.  Every such kernel
//   * calls pdl_trigger() first (lets ITS successor be scheduled as early as possible), then
//   * does only predecessor-independent set-up (shared memory, barriers, TMEM, kernel-parameter reads), and
//   * calls pdl_wait() -- which returns once the predecessor grid has completed and flushed -- before its
//     first global-memory access.
// Only kernels that contain pdl_wait() may be launched this way.  AVJ_PDL=0 drops the attribute (the device
// instructions are no-ops for a normally launched grid).
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool avj_pdl_enabled();

template <typename... KArgs, typename... Args>
static inline cudaError_t avj_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                         Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = avj_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- in-library kernel timing (avj_prof_*): when enabled every instrumented entry point brackets its
// launches with CUDA events on the launch stream; bench.py reads per-family totals for the roofline.
enum { AVJ_FAM_GEMM = 0, AVJ_FAM_ATTN_FWD = 1, AVJ_FAM_ATTN_BWD = 2, AVJ_FAM_LN_FWD = 3, AVJ_FAM_LN_BWD = 4,
       AVJ_FAM_COLSUM = 5, AVJ_FAM_OPTIM = 6, AVJ_FAM_OTHER = 7, AVJ_FAM_COUNT = 8 };
extern bool g_avj_prof_on;
void avj_prof_begin(int family, double work, cudaStream_t s, int d0 = 0, int d1 = 0, int d2 = 0, int d3 = 0);
void avj_prof_end(cudaStream_t s);
struct AvjProfScope {
  cudaStream_t s; bool on;
  // d0..d3: free-form shape of the launch (GEMM: layout|epilogue bits, M, N, K; attention: B, N, H, hd;
  // row kernels: rows, D) -- only written out by avj_prof_dump
  AvjProfScope(int family, double work, void* stream, int d0 = 0, int d1 = 0, int d2 = 0, int d3 = 0)
      : s(reinterpret_cast<cudaStream_t>(stream)), on(g_avj_prof_on) {
    if (on) avj_prof_begin(family, work, s, d0, d1, d2, d3);
  }
  ~AvjProfScope() { if (on) avj_prof_end(s); }
};

// ---- dtype helpers ------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<bf16>(bf16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void unpack_bf16x2(uint32_t u, float& lo, float& hi) {
  __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&u);
  lo = __bfloat162float(t.x);
  hi = __bfloat162float(t.y);
}

// Load / store 8 consecutive elements as floats (16 B for bf16, 32 B for fp32).
template <typename T> __device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <> __device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void load8<bf16>(const bf16* p, float (&v)[8]) {
  uint4 a = *reinterpret_cast<const uint4*>(p);
  unpack_bf16x2(a.x, v[0], v[1]); unpack_bf16x2(a.y, v[2], v[3]);
  unpack_bf16x2(a.z, v[4], v[5]); unpack_bf16x2(a.w, v[6], v[7]);
}
template <typename T> __device__ __forceinline__ void store8(T* p, const float (&v)[8]);
template <> __device__ __forceinline__ void store8<float>(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void store8<bf16>(bf16* p, const float (&v)[8]) {
  uint4 a;
  a.x = pack_bf16x2(v[0], v[1]); a.y = pack_bf16x2(v[2], v[3]);
  a.z = pack_bf16x2(v[4], v[5]); a.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = a;
}

// ---- row map -----------------------------------------------------------------------------
__host__ __device__ __forceinline__ int64_t map_row(const avj_rowmap& m, int64_t r) {
  if (m.rows_per_group == 0) return r + m.row_offset;
  return (r / m.rows_per_group) * (int64_t)m.group_stride + (r % m.rows_per_group) + m.row_offset;
}

// ---- math --------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// Production-mode GELU (results are rounded to bf16 anyway; fp32 check mode keeps erff()):
//   Phi(x) ~= sigmoid(u),   u = x (c0 + c1 x^2 + c2 x^4),  x^2 clamped to 50
// fitted to the exact normal CDF: |x Phi - gelu_erf(x)| <= 2.6e-5 and |d/dx - gelu_erf'(x)| <= 1.1e-4 for
// all x with an exact sigmoid (bf16 has a relative step of 3.9e-3).
//   forward : sigmoid(u) = 0.5 + 0.5 tanh(u/2) with tanh.approx.f32 (ONE MUFU op, <= 2^-11 relative):
//             8 instructions per element -- the fc1 epilogue that applies it is MUFU/issue bound.
//   backward: s = 1/(1+e), e = exp(-u); s(1-s) = e s^2 has no cancellation near saturation (tanh.approx
//             would turn its 2^-11 error into a 1e-2 error of GELU' for x > 3.5), 13 instructions.
#define AVJ_GELU_C0 1.5950157685561237f
#define AVJ_GELU_C1 0.07401129205320253f
#define AVJ_GELU_C2 -0.0007030335786691012f
__device__ __forceinline__ float gelu_fast_fwd(float x) {
  const float x2 = fminf(x * x, 50.0f);
  const float h = x * fmaf(fmaf(0.5f * AVJ_GELU_C2, x2, 0.5f * AVJ_GELU_C1), x2, 0.5f * AVJ_GELU_C0);   // u / 2
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return x * fmaf(0.5f, t, 0.5f);
}
__device__ __forceinline__ float gelu_fast_bwd(float x) {
  const float x2 = fminf(x * x, 50.0f);
  const float u = x * fmaf(fmaf(AVJ_GELU_C2, x2, AVJ_GELU_C1), x2, AVJ_GELU_C0);
  float e, s;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(u * -1.4426950408889634f));
  e = fminf(e, 1e30f);                                   // keep e * s * s finite for very negative x
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(1.0f + e));
  const float du = fmaf(fmaf(5.0f * AVJ_GELU_C2, x2, 3.0f * AVJ_GELU_C1), x2, AVJ_GELU_C0);
  return fmaf(x * (e * s * s), du, s);                   // Phi + x Phi'
}
template <bool FAST> __device__ __forceinline__ float gelu_fwd(float x) {
  if (FAST) return gelu_fast_fwd(x);
  return gelu_erf(x);
}
template <bool FAST> __device__ __forceinline__ float gelu_bwd(float x) {
  if (FAST) return gelu_fast_bwd(x);
  return gelu_erf_grad(x);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum for blockDim.x <= 1024 (multiple of 32); `red` is >= 32 floats of smem.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (w == 0) t = warp_sum(t);
  if (threadIdx.x == 0) red[0] = t;
  __syncthreads();
  return red[0];
}

// ---- shared GEMM epilogue (used by the SIMT check-mode GEMM and the tcgen05 GEMM) ----------
// epilogue_prefetch gathers everything that is ADDED to the activated accumulator (fp32 residual row,
// gathered positional row, previous C for `C +=`) into registers; it does not depend on the
// accumulator, so the tcgen05 kernel issues it while the TMEM load is still in flight.
template <int NV>
__device__ __forceinline__ void epilogue_prefetch(const avj_epilogue& ep, const void* C, int ldc, int N,
                                                  int64_t r, int n0, float (&add)[NV]) {
#pragma unroll
  for (int i = 0; i < NV; ++i) add[i] = 0.f;
  const int64_t pr = map_row(ep.out_map, r);
  if (ep.residual) {
    const float* res = ep.residual + pr * (int64_t)ldc + n0;
#pragma unroll
    for (int i = 0; i < NV; i += 4) {
      float4 b = *reinterpret_cast<const float4*>(res + i);
      add[i] += b.x; add[i + 1] += b.y; add[i + 2] += b.z; add[i + 3] += b.w;
    }
  }
  if (ep.pos) {
    const int64_t prow = ep.pos_idx ? ep.pos_idx[r] : (r % ep.pos_rows);
    const float* pp = ep.pos + prow * (int64_t)N + n0;
#pragma unroll
    for (int i = 0; i < NV; i += 4) {
      float4 b = *reinterpret_cast<const float4*>(pp + i);
      add[i] += b.x; add[i + 1] += b.y; add[i + 2] += b.z; add[i + 3] += b.w;
    }
  }
  if (ep.accumulate && ep.out_dtype == AVJ_F32) {
    const float* out = reinterpret_cast<const float*>(C) + pr * (int64_t)ldc + n0;
#pragma unroll
    for (int i = 0; i < NV; i += 4) {
      float4 b = *reinterpret_cast<const float4*>(out + i);
      add[i] += b.x; add[i + 1] += b.y; add[i + 2] += b.z; add[i + 3] += b.w;
    }
  }
}

// Applies bias / activation / GELU' to NV consecutive accumulator columns of logical row r starting
// at column n0, adds the prefetched addend and stores.  NV must be a multiple of 8.
template <typename TAct, int NV, bool FAST>
__device__ __forceinline__ void epilogue_apply_store(const avj_epilogue& ep, void* C, int ldc, int N,
                                                     int64_t r, int n0, float (&acc)[NV], const float (&add)[NV]) {
  if (ep.bias) {
#pragma unroll
    for (int i = 0; i < NV; i += 4) {
      float4 b = *reinterpret_cast<const float4*>(ep.bias + n0 + i);
      acc[i] += b.x; acc[i + 1] += b.y; acc[i + 2] += b.z; acc[i + 3] += b.w;
    }
  }
  if (ep.act == 1) {
    if (ep.pre_out) {
      TAct* pre = reinterpret_cast<TAct*>(ep.pre_out) + r * (int64_t)N + n0;
#pragma unroll
      for (int i = 0; i < NV; i += 8) {
        float t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = acc[i + j];
        store8<TAct>(pre + i, t);
      }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = gelu_fwd<FAST>(acc[i]);
  }
  if (ep.dact_aux) {
    const TAct* aux = reinterpret_cast<const TAct*>(ep.dact_aux) + r * (int64_t)N + n0;
#pragma unroll
    for (int i = 0; i < NV; i += 8) {
      float t[8];
      load8<TAct>(aux + i, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i + j] *= gelu_bwd<FAST>(t[j]);
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] += add[i];
  const int64_t pr = map_row(ep.out_map, r);
  if (ep.out_dtype == AVJ_F32) {
    float* out = reinterpret_cast<float*>(C) + pr * (int64_t)ldc + n0;
#pragma unroll
    for (int i = 0; i < NV; i += 4)
      *reinterpret_cast<float4*>(out + i) = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
  } else {
    bf16* out = reinterpret_cast<bf16*>(C) + pr * (int64_t)ldc + n0;
#pragma unroll
    for (int i = 0; i < NV; i += 8) {
      float t[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) t[j] = acc[i + j];
      store8<bf16>(out + i, t);
    }
  }
}
