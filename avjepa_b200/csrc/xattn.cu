// Few-query cross-attention (AttentivePooler, src/models/attentive_pooler.py:21-102 -> CrossAttention,
// src/models/utils/modules.py:123-159): n learned queries (n = 1 in every classifier of the reference) attend over the N
// tokens of a frozen encoder.  With a handful of queries there is no reuse of K / V across query rows, so this is not a
// tensor-core problem: every K and V row is read exactly once per (batch, head) -- the kernel is HBM-bound and built
// as such (16-byte loads, one pass for the scores, one for the weighted sum).
//
//   q  : [B, n, H, hd]      (the q-Linear output)
//   kv : [B, N, 2, H, hd]   (the kv-Linear output, read in place -- no permute copies)
//   out: [B, n, H*hd], lse: fp32 [B, H, n]
//
// One CTA per (head, batch) walks its n queries.  Scores live in shared memory (N floats).  Algorithmic bytes:
// forward 2*N*hd*sizeof per (b, h) (+ n-fold re-reads that hit L2/L1), backward the same plus the dK / dV rows written.
#include "common.cuh"

#define XA_THREADS 256

template <typename T>
__device__ __forceinline__ float xa_dot(const T* __restrict__ row, const float* __restrict__ qs, int hd) {
  float acc = 0.f;
  for (int c = 0; c < hd; c += 8) {
    float v[8];
    load8<T>(row + c, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc = fmaf(v[e], qs[c + e], acc);
  }
  return acc;
}

__device__ __forceinline__ float xa_block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = (threadIdx.x < (XA_THREADS >> 5)) ? red[threadIdx.x] : -INFINITY;
  if (w == 0) t = warp_max(t);
  if (threadIdx.x == 0) red[0] = t;
  __syncthreads();
  return red[0];
}

// out[d] = sum_j w[j] * rows[j][d]:  thread (chunk, group) accumulates 8 dims over keys j = group, group + G, ...
template <typename T>
__device__ __forceinline__ void xa_weighted_sum(const T* __restrict__ base, int64_t row_stride, const float* __restrict__ w, int N, int hd,
                                                float* __restrict__ part /* [G][hd] */, float* __restrict__ outv /* [hd] */) {
  const int CH = hd / 8, G = XA_THREADS / CH;
  const int ch = threadIdx.x % CH, g = threadIdx.x / CH;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (g < G) {
    for (int j = g; j < N; j += G) {
      float v[8];
      load8<T>(base + (int64_t)j * row_stride + ch * 8, v);
      const float wj = w[j];
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(wj, v[e], acc[e]);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) part[g * hd + ch * 8 + e] = acc[e];
  }
  __syncthreads();
  if (threadIdx.x < hd) {
    float t = 0.f;
    for (int gg = 0; gg < G; ++gg) t += part[gg * hd + threadIdx.x];
    outv[threadIdx.x] = t;
  }
  __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(XA_THREADS)
xattn_fwd_kernel(const T* __restrict__ q, const T* __restrict__ kv, T* __restrict__ out, float* __restrict__ lse,
                 int N, int n, int H, int hd, float scale_log2) {
  extern __shared__ float xs[];            // [N] scores | [hd] q | [hd] result | [G*hd] partials | [32] reduction
  float* s = xs;
  float* qs = s + N;
  float* ov = qs + hd;
  float* part = ov + hd;
  const int G = XA_THREADS / (hd / 8);
  float* red = part + G * hd;
  const int h = blockIdx.x, b = blockIdx.y;
  const int64_t rs = 2 * (int64_t)H * hd;
  const T* kb = kv + (int64_t)b * N * rs + (int64_t)h * hd;
  const T* vb = kb + (int64_t)H * hd;
  for (int qi = 0; qi < n; ++qi) {
    const T* qrow = q + (((int64_t)b * n + qi) * H + h) * hd;
    for (int d = threadIdx.x; d < hd; d += XA_THREADS) qs[d] = to_f32<T>(qrow[d]) * scale_log2;
    __syncthreads();
    float mx = -INFINITY;
    for (int j = threadIdx.x; j < N; j += XA_THREADS) {
      const float v = xa_dot<T>(kb + (int64_t)j * rs, qs, hd);
      s[j] = v;
      mx = fmaxf(mx, v);
    }
    const float m = xa_block_max(mx, red);
    float sum = 0.f;
    for (int j = threadIdx.x; j < N; j += XA_THREADS) {
      const float p = exp2f(s[j] - m);
      s[j] = p;
      sum += p;
    }
    const float l = block_sum(sum, red);
    xa_weighted_sum<T>(vb, rs, s, N, hd, part, ov);
    if (threadIdx.x < hd) out[(((int64_t)b * n + qi) * H + h) * hd + threadIdx.x] = from_f32<T>(ov[threadIdx.x] / l);
    if (threadIdx.x == 0) lse[((int64_t)b * H + h) * n + qi] = (m + log2f(l)) * 0.69314718055994530942f;
    __syncthreads();
  }
}

// dq [B, n, H, hd];  dkv [B, N, 2, H, hd] (WRITTEN by the first query of a (b, h), accumulated by the others)
template <typename T>
__global__ void __launch_bounds__(XA_THREADS)
xattn_bwd_kernel(const T* __restrict__ q, const T* __restrict__ kv, const T* __restrict__ out, const T* __restrict__ dout,
                 const float* __restrict__ lse, T* __restrict__ dq, T* __restrict__ dkv, int N, int n, int H, int hd, float scale) {
  extern __shared__ float xs[];            // [N] ds | [hd] q*scale*log2e | [hd] dO | [hd] q (raw) | [hd] result | [G*hd] partials | [32]
  float* s = xs;
  float* qs = s + N;
  float* dov = qs + hd;
  float* qraw = dov + hd;
  float* ov = qraw + hd;
  float* part = ov + hd;
  const int G = XA_THREADS / (hd / 8);
  float* red = part + G * hd;
  const int h = blockIdx.x, b = blockIdx.y;
  const int64_t rs = 2 * (int64_t)H * hd;
  const T* kb = kv + (int64_t)b * N * rs + (int64_t)h * hd;
  const T* vb = kb + (int64_t)H * hd;
  T* dkb = dkv + (int64_t)b * N * rs + (int64_t)h * hd;
  T* dvb = dkb + (int64_t)H * hd;
  const float scale_log2 = scale * 1.4426950408889634f;
  for (int qi = 0; qi < n; ++qi) {
    const int64_t qoff = (((int64_t)b * n + qi) * H + h) * hd;
    float dl = 0.f;
    for (int d = threadIdx.x; d < hd; d += XA_THREADS) {
      const float qv = to_f32<T>(q[qoff + d]), g = to_f32<T>(dout[qoff + d]);
      qraw[d] = qv; qs[d] = qv * scale_log2; dov[d] = g;
      dl += g * to_f32<T>(out[qoff + d]);
    }
    const float delta = block_sum(dl, red);            // sum_d dO * O
    const float l2 = lse[((int64_t)b * H + h) * n + qi] * 1.4426950408889634f;
    // per key: p = exp2(q.k*scale*log2e - lse2), dp = dO.v, ds = p (dp - delta) scale;  dV_j (+)= p dO, dK_j (+)= ds q
    for (int j = threadIdx.x; j < N; j += XA_THREADS) {
      const float p = exp2f(xa_dot<T>(kb + (int64_t)j * rs, qs, hd) - l2);
      const float dp = xa_dot<T>(vb + (int64_t)j * rs, dov, hd);
      const float ds = p * (dp - delta) * scale;
      s[j] = ds;
      T* dkr = dkb + (int64_t)j * rs;
      T* dvr = dvb + (int64_t)j * rs;
      for (int c = 0; c < hd; c += 8) {
        float a[8], v[8];
        if (qi > 0) { load8<T>(dkr + c, a); load8<T>(dvr + c, v); }
        else {
#pragma unroll
          for (int e = 0; e < 8; ++e) { a[e] = 0.f; v[e] = 0.f; }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) { a[e] = fmaf(ds, qraw[c + e], a[e]); v[e] = fmaf(p, dov[c + e], v[e]); }
        store8<T>(dkr + c, a);
        store8<T>(dvr + c, v);
      }
    }
    __syncthreads();
    xa_weighted_sum<T>(kb, rs, s, N, hd, part, ov);     // dq = sum_j ds_j k_j
    if (threadIdx.x < hd) dq[qoff + threadIdx.x] = from_f32<T>(ov[threadIdx.x]);
    __syncthreads();
  }
}

static size_t xa_smem(int N, int hd, int extra_vecs) {
  const int G = XA_THREADS / (hd / 8);
  return sizeof(float) * ((size_t)N + (size_t)extra_vecs * hd + (size_t)G * hd + 32);
}

extern "C" int avj_xattn_fwd(int dtype, const void* q, const void* kv, void* out, float* lse,
                             int B, int N, int n, int H, int hd, float scale, void* stream) {
  AVJ_CHECK(hd % 8 == 0 && hd >= 8 && hd <= 256, "avj_xattn_fwd: head_dim %d must be a multiple of 8 in [8, 256]", hd);
  AVJ_CHECK(dtype == AVJ_F32 || dtype == AVJ_BF16, "avj_xattn_fwd: bad dtype %d", dtype);
  if (B == 0 || N == 0 || n == 0) return 0;
  const size_t smem = xa_smem(N, hd, 2);
  AVJ_CHECK(smem <= 200 * 1024, "avj_xattn_fwd: N=%d does not fit the shared-memory score buffer", N);
  AvjProfScope prof(AVJ_FAM_OTHER, 2.0 * B * N * H * hd * (dtype == AVJ_BF16 ? 2 : 4), stream, 10, B * N, H * hd);
  dim3 grid(H, B);
  const float sl2 = scale * 1.4426950408889634f;
  if (dtype == AVJ_BF16) {
    AVJ_CUDA(cudaFuncSetAttribute(xattn_fwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    xattn_fwd_kernel<bf16><<<grid, XA_THREADS, smem, as_stream(stream)>>>((const bf16*)q, (const bf16*)kv, (bf16*)out, lse, N, n, H, hd, sl2);
  } else {
    AVJ_CUDA(cudaFuncSetAttribute(xattn_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    xattn_fwd_kernel<float><<<grid, XA_THREADS, smem, as_stream(stream)>>>((const float*)q, (const float*)kv, (float*)out, lse, N, n, H, hd, sl2);
  }
  AVJ_LAUNCH_CHECK();
  return 0;
}

extern "C" int avj_xattn_bwd(int dtype, const void* q, const void* kv, const void* out, const void* dout, const float* lse,
                             void* dq, void* dkv, int B, int N, int n, int H, int hd, float scale, void* stream) {
  AVJ_CHECK(hd % 8 == 0 && hd >= 8 && hd <= 256, "avj_xattn_bwd: head_dim %d must be a multiple of 8 in [8, 256]", hd);
  AVJ_CHECK(dtype == AVJ_F32 || dtype == AVJ_BF16, "avj_xattn_bwd: bad dtype %d", dtype);
  if (B == 0 || N == 0 || n == 0) return 0;
  const size_t smem = xa_smem(N, hd, 4);
  AVJ_CHECK(smem <= 200 * 1024, "avj_xattn_bwd: N=%d does not fit the shared-memory score buffer", N);
  AvjProfScope prof(AVJ_FAM_OTHER, 4.0 * B * N * H * hd * (dtype == AVJ_BF16 ? 2 : 4), stream, 11, B * N, H * hd);
  dim3 grid(H, B);
  if (dtype == AVJ_BF16) {
    AVJ_CUDA(cudaFuncSetAttribute(xattn_bwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    xattn_bwd_kernel<bf16><<<grid, XA_THREADS, smem, as_stream(stream)>>>((const bf16*)q, (const bf16*)kv, (const bf16*)out, (const bf16*)dout,
                                                                         lse, (bf16*)dq, (bf16*)dkv, N, n, H, hd, scale);
  } else {
    AVJ_CUDA(cudaFuncSetAttribute(xattn_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    xattn_bwd_kernel<float><<<grid, XA_THREADS, smem, as_stream(stream)>>>((const float*)q, (const float*)kv, (const float*)out,
                                                                           (const float*)dout, lse, (float*)dq, (float*)dkv, N, n, H, hd, scale);
  }
  AVJ_LAUNCH_CHECK();
  return 0;
}
