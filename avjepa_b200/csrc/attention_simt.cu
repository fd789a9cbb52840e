// Check-mode attention (fp32 FMA, exact softmax bookkeeping in fp32): forward with online
// softmax, backward as two deterministic passes (dQ by query, dK/dV by key), P recomputed from
// the saved log-sum-exp.  Reads q/k/v in place from the qkv-Linear output [B, N, 3, H, hd].
// This is the 1e-4 parity path and the fallback for head dims the tensor-core kernel does
// not take; the production kernel is attention_mma.cu / attention_umma.cu.
#include "common.cuh"

#define AT_WARPS 8
#define AT_QPW 4                       // queries (or keys, in the dK/dV pass) per warp
#define AT_ROWS (AT_WARPS * AT_QPW)    // 32 rows per CTA
#define AT_TILE 32                     // rows of the "other" operand staged per step

template <typename T>
__device__ __forceinline__ void stage_rows(const T* __restrict__ base, int64_t row_stride, int row0, int nrows_valid,
                                           int hd, float* __restrict__ dst, int dst_ld) {
  // copies up to AT_TILE rows of hd elements (zero-fills beyond nrows_valid)
  for (int e = threadIdx.x; e < AT_TILE * hd; e += blockDim.x) {
    const int r = e / hd, d = e % hd;
    dst[r * dst_ld + d] = (r < nrows_valid) ? to_f32<T>(base[(int64_t)(row0 + r) * row_stride + d]) : 0.f;
  }
}

// ---------------------------------------------------------------------------------------- fwd
template <typename T, int ND>
__global__ void __launch_bounds__(AT_WARPS * 32)
attn_fwd_simt(const T* __restrict__ qkv, T* __restrict__ out, float* __restrict__ lse,
              int B, int N, int H, int hd, float scale) {
  extern __shared__ float sm[];
  const int ldk = hd + 1;
  float* Ks = sm;                         // [AT_TILE][ldk]
  float* Vs = Ks + AT_TILE * ldk;         // [AT_TILE][ldk]
  float* Qs = Vs + AT_TILE * ldk;         // [AT_ROWS][ldk]
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * AT_ROWS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t rs = 3 * (int64_t)H * hd;
  const T* qb = qkv + (int64_t)b * N * rs + (int64_t)h * hd;
  const T* kb = qb + (int64_t)H * hd;
  const T* vb = kb + (int64_t)H * hd;

  stage_rows<T>(qb, rs, q0, min(AT_ROWS, N - q0), hd, Qs, ldk);
  float m[AT_QPW], l[AT_QPW], o[AT_QPW][ND];
#pragma unroll
  for (int i = 0; i < AT_QPW; ++i) {
    m[i] = -INFINITY; l[i] = 0.f;
#pragma unroll
    for (int d = 0; d < ND; ++d) o[i][d] = 0.f;
  }
  for (int k0 = 0; k0 < N; k0 += AT_TILE) {
    __syncthreads();
    const int nv = min(AT_TILE, N - k0);
    stage_rows<T>(kb, rs, k0, nv, hd, Ks, ldk);
    stage_rows<T>(vb, rs, k0, nv, hd, Vs, ldk);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < AT_QPW; ++i) {
      const float* qv = Qs + (warp * AT_QPW + i) * ldk;
      float s = 0.f;
      for (int d = 0; d < hd; ++d) s = fmaf(qv[d], Ks[lane * ldk + d], s);
      s = (lane < nv) ? s * scale : -INFINITY;
      const float mn = fmaxf(m[i], warp_max(s));
      const float p = __expf(s - mn);
      const float corr = __expf(m[i] - mn);
      l[i] = l[i] * corr + warp_sum(p);
      m[i] = mn;
#pragma unroll
      for (int d = 0; d < ND; ++d) o[i][d] *= corr;
      for (int j = 0; j < AT_TILE; ++j) {
        const float pj = __shfl_sync(0xffffffffu, p, j);
#pragma unroll
        for (int d = 0; d < ND; ++d) {
          const int dd = lane + 32 * d;
          if (dd < hd) o[i][d] = fmaf(pj, Vs[j * ldk + dd], o[i][d]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < AT_QPW; ++i) {
    const int qi = q0 + warp * AT_QPW + i;
    if (qi >= N) continue;
    const float inv = 1.0f / l[i];
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      const int dd = lane + 32 * d;
      if (dd < hd) out[((int64_t)b * N + qi) * H * hd + (int64_t)h * hd + dd] = from_f32<T>(o[i][d] * inv);
    }
    if (lane == 0) lse[((int64_t)b * H + h) * N + qi] = m[i] + __logf(l[i]);
  }
}

// ------------------------------------------------------------------------------- bwd: delta
// delta[b,h,i] = sum_d dout[b,i,h,d] * out[b,i,h,d]
template <typename T>
__global__ void attn_delta_kernel(const T* __restrict__ out, const T* __restrict__ dout, float* __restrict__ delta,
                                  int B, int N, int H, int hd) {
  const int64_t total = (int64_t)B * N * H;
  const int lane = threadIdx.x & 31;
  for (int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; w < total; w += ((int64_t)gridDim.x * blockDim.x) >> 5) {
    const int h = (int)(w % H);
    const int64_t bn = w / H;
    const int i = (int)(bn % N), b = (int)(bn / N);
    const T* o = out + (bn * H + h) * hd;
    const T* g = dout + (bn * H + h) * hd;
    float s = 0.f;
    for (int d = lane; d < hd; d += 32) s = fmaf(to_f32<T>(o[d]), to_f32<T>(g[d]), s);
    s = warp_sum(s);
    if (lane == 0) delta[((int64_t)b * H + h) * N + i] = s;
  }
}

// ---------------------------------------------------------------------------------- bwd: dQ
template <typename T, int ND>
__global__ void __launch_bounds__(AT_WARPS * 32)
attn_bwd_dq_simt(const T* __restrict__ qkv, const T* __restrict__ dout, const float* __restrict__ lse,
                 const float* __restrict__ delta, T* __restrict__ dqkv, int B, int N, int H, int hd, float scale) {
  extern __shared__ float sm[];
  const int ldk = hd + 1;
  float* Ks = sm;
  float* Vs = Ks + AT_TILE * ldk;
  float* Qs = Vs + AT_TILE * ldk;          // [AT_ROWS][ldk]
  float* Gs = Qs + AT_ROWS * ldk;          // dout rows [AT_ROWS][ldk]
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * AT_ROWS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t rs = 3 * (int64_t)H * hd;
  const T* qb = qkv + (int64_t)b * N * rs + (int64_t)h * hd;
  const T* kb = qb + (int64_t)H * hd;
  const T* vb = kb + (int64_t)H * hd;
  const T* gb = dout + (int64_t)b * N * H * hd + (int64_t)h * hd;
  const int nq = min(AT_ROWS, N - q0);
  stage_rows<T>(qb, rs, q0, nq, hd, Qs, ldk);
  stage_rows<T>(gb, (int64_t)H * hd, q0, nq, hd, Gs, ldk);
  float dq[AT_QPW][ND], L[AT_QPW], Dl[AT_QPW];
#pragma unroll
  for (int i = 0; i < AT_QPW; ++i) {
    const int qi = q0 + warp * AT_QPW + i;
    L[i] = (qi < N) ? lse[((int64_t)b * H + h) * N + qi] : 0.f;
    Dl[i] = (qi < N) ? delta[((int64_t)b * H + h) * N + qi] : 0.f;
#pragma unroll
    for (int d = 0; d < ND; ++d) dq[i][d] = 0.f;
  }
  for (int k0 = 0; k0 < N; k0 += AT_TILE) {
    __syncthreads();
    const int nv = min(AT_TILE, N - k0);
    stage_rows<T>(kb, rs, k0, nv, hd, Ks, ldk);
    stage_rows<T>(vb, rs, k0, nv, hd, Vs, ldk);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < AT_QPW; ++i) {
      const float* qv = Qs + (warp * AT_QPW + i) * ldk;
      const float* gv = Gs + (warp * AT_QPW + i) * ldk;
      float s = 0.f, dp = 0.f;
      for (int d = 0; d < hd; ++d) {
        s = fmaf(qv[d], Ks[lane * ldk + d], s);
        dp = fmaf(gv[d], Vs[lane * ldk + d], dp);
      }
      const float p = (lane < nv) ? __expf(s * scale - L[i]) : 0.f;
      const float ds = p * (dp - Dl[i]) * scale;
      for (int j = 0; j < AT_TILE; ++j) {
        const float dsj = __shfl_sync(0xffffffffu, ds, j);
#pragma unroll
        for (int d = 0; d < ND; ++d) {
          const int dd = lane + 32 * d;
          if (dd < hd) dq[i][d] = fmaf(dsj, Ks[j * ldk + dd], dq[i][d]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < AT_QPW; ++i) {
    const int qi = q0 + warp * AT_QPW + i;
    if (qi >= N) continue;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      const int dd = lane + 32 * d;
      if (dd < hd) dqkv[((int64_t)b * N + qi) * rs + (int64_t)h * hd + dd] = from_f32<T>(dq[i][d]);
    }
  }
}

// ------------------------------------------------------------------------------- bwd: dK, dV
template <typename T, int ND>
__global__ void __launch_bounds__(AT_WARPS * 32)
attn_bwd_dkv_simt(const T* __restrict__ qkv, const T* __restrict__ dout, const float* __restrict__ lse,
                  const float* __restrict__ delta, T* __restrict__ dqkv, int B, int N, int H, int hd, float scale) {
  extern __shared__ float sm[];
  const int ldk = hd + 1;
  float* Qs = sm;                          // staged query tile [AT_TILE][ldk]
  float* Gs = Qs + AT_TILE * ldk;          // staged dout tile  [AT_TILE][ldk]
  float* Ks = Gs + AT_TILE * ldk;          // this CTA's keys   [AT_ROWS][ldk]
  float* Vs = Ks + AT_ROWS * ldk;          // this CTA's values [AT_ROWS][ldk]
  float* Ls = Vs + AT_ROWS * ldk;          // lse of staged queries [AT_TILE]
  float* Ds = Ls + AT_TILE;                // delta of staged queries [AT_TILE]
  const int b = blockIdx.z, h = blockIdx.y, k0 = blockIdx.x * AT_ROWS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t rs = 3 * (int64_t)H * hd;
  const T* qb = qkv + (int64_t)b * N * rs + (int64_t)h * hd;
  const T* kb = qb + (int64_t)H * hd;
  const T* vb = kb + (int64_t)H * hd;
  const T* gb = dout + (int64_t)b * N * H * hd + (int64_t)h * hd;
  const int nk = min(AT_ROWS, N - k0);
  stage_rows<T>(kb, rs, k0, nk, hd, Ks, ldk);
  stage_rows<T>(vb, rs, k0, nk, hd, Vs, ldk);
  float dk[AT_QPW][ND], dv[AT_QPW][ND];
#pragma unroll
  for (int i = 0; i < AT_QPW; ++i)
#pragma unroll
    for (int d = 0; d < ND; ++d) { dk[i][d] = 0.f; dv[i][d] = 0.f; }

  for (int q0 = 0; q0 < N; q0 += AT_TILE) {
    __syncthreads();
    const int nv = min(AT_TILE, N - q0);
    stage_rows<T>(qb, rs, q0, nv, hd, Qs, ldk);
    stage_rows<T>(gb, (int64_t)H * hd, q0, nv, hd, Gs, ldk);
    if (threadIdx.x < AT_TILE) {
      const int qi = q0 + threadIdx.x;
      Ls[threadIdx.x] = (qi < N) ? lse[((int64_t)b * H + h) * N + qi] : 0.f;
      Ds[threadIdx.x] = (qi < N) ? delta[((int64_t)b * H + h) * N + qi] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < AT_QPW; ++i) {
      const float* kv = Ks + (warp * AT_QPW + i) * ldk;
      const float* vv = Vs + (warp * AT_QPW + i) * ldk;
      // lane <-> staged query `lane`
      float s = 0.f, dp = 0.f;
      for (int d = 0; d < hd; ++d) {
        s = fmaf(Qs[lane * ldk + d], kv[d], s);
        dp = fmaf(Gs[lane * ldk + d], vv[d], dp);
      }
      const float p = (lane < nv) ? __expf(s * scale - Ls[lane]) : 0.f;
      const float ds = p * (dp - Ds[lane]) * scale;
      for (int j = 0; j < AT_TILE; ++j) {
        const float pj = __shfl_sync(0xffffffffu, p, j);
        const float dsj = __shfl_sync(0xffffffffu, ds, j);
#pragma unroll
        for (int d = 0; d < ND; ++d) {
          const int dd = lane + 32 * d;
          if (dd < hd) {
            dv[i][d] = fmaf(pj, Gs[j * ldk + dd], dv[i][d]);
            dk[i][d] = fmaf(dsj, Qs[j * ldk + dd], dk[i][d]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < AT_QPW; ++i) {
    const int ki = k0 + warp * AT_QPW + i;
    if (ki >= N) continue;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      const int dd = lane + 32 * d;
      if (dd < hd) {
        const int64_t o = ((int64_t)b * N + ki) * rs + (int64_t)h * hd + dd;
        dqkv[o + (int64_t)H * hd] = from_f32<T>(dk[i][d]);
        dqkv[o + 2 * (int64_t)H * hd] = from_f32<T>(dv[i][d]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------- launchers
template <typename T>
static int fwd_launch(const T* qkv, T* out, float* lse, int B, int N, int H, int hd, float scale, cudaStream_t s) {
  const int nd = (hd + 31) / 32;
  dim3 grid((N + AT_ROWS - 1) / AT_ROWS, H, B);
  const size_t smem = (size_t)(2 * AT_TILE + AT_ROWS) * (hd + 1) * sizeof(float);
#define FWD_CASE(ND) case ND:                                                                               \
    if (smem > 48 * 1024) cudaFuncSetAttribute(attn_fwd_simt<T, ND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    attn_fwd_simt<T, ND><<<grid, AT_WARPS * 32, smem, s>>>(qkv, out, lse, B, N, H, hd, scale); break;
  switch (nd) { FWD_CASE(1) FWD_CASE(2) FWD_CASE(3) FWD_CASE(4)
    default: avj_set_error("attention: head_dim %d > 128 unsupported", hd); return 1; }
#undef FWD_CASE
  AVJ_LAUNCH_CHECK();
  return 0;
}

template <typename T>
static int bwd_launch(const T* qkv, const T* out, const T* dout, const float* lse, T* dqkv, float* ws,
                      int B, int N, int H, int hd, float scale, cudaStream_t s) {
  const int nd = (hd + 31) / 32;
  float* delta = ws;
  {
    const int64_t warps = (int64_t)B * N * H;
    int grid = (int)((warps + 7) / 8);
    const int cap = avj_num_sms() * 16;
    if (grid > cap) grid = cap;
    attn_delta_kernel<T><<<grid, 256, 0, s>>>(out, dout, delta, B, N, H, hd);
    AVJ_LAUNCH_CHECK();
  }
  dim3 grid((N + AT_ROWS - 1) / AT_ROWS, H, B);
  const size_t smem_dq = (size_t)(2 * AT_TILE + 2 * AT_ROWS) * (hd + 1) * sizeof(float);
  const size_t smem_dkv = smem_dq + 2 * AT_TILE * sizeof(float);
#define BWD_CASE(ND)                                                                                              \
  case ND:                                                                                                        \
    if (smem_dkv > 48 * 1024) {                                                                                   \
      cudaFuncSetAttribute(attn_bwd_dq_simt<T, ND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_dq);   \
      cudaFuncSetAttribute(attn_bwd_dkv_simt<T, ND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_dkv); \
    }                                                                                                             \
    attn_bwd_dq_simt<T, ND><<<grid, AT_WARPS * 32, smem_dq, s>>>(qkv, dout, lse, delta, dqkv, B, N, H, hd, scale);   \
    AVJ_COUNT_LAUNCH();                                                                                             \
    attn_bwd_dkv_simt<T, ND><<<grid, AT_WARPS * 32, smem_dkv, s>>>(qkv, dout, lse, delta, dqkv, B, N, H, hd, scale); \
    break;
  switch (nd) { BWD_CASE(1) BWD_CASE(2) BWD_CASE(3) BWD_CASE(4)
    default: avj_set_error("attention: head_dim %d > 128 unsupported", hd); return 1; }
#undef BWD_CASE
  AVJ_LAUNCH_CHECK();
  return 0;
}

int avj_attention_fwd_simt(int dtype, const void* qkv, void* out, float* lse, int B, int N, int H, int hd, float scale, cudaStream_t s) {
  if (dtype == AVJ_BF16) return fwd_launch<bf16>((const bf16*)qkv, (bf16*)out, lse, B, N, H, hd, scale, s);
  return fwd_launch<float>((const float*)qkv, (float*)out, lse, B, N, H, hd, scale, s);
}

int avj_attention_bwd_simt(int dtype, const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                           float* ws, int B, int N, int H, int hd, float scale, cudaStream_t s) {
  if (dtype == AVJ_BF16)
    return bwd_launch<bf16>((const bf16*)qkv, (const bf16*)out, (const bf16*)dout, lse, (bf16*)dqkv, ws, B, N, H, hd, scale, s);
  return bwd_launch<float>((const float*)qkv, (const float*)out, (const float*)dout, lse, (float*)dqkv, ws, B, N, H, hd, scale, s);
}
