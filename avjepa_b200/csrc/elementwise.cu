// HBM-bound kernels of the AV-JEPA step: patch extraction, mask gather/scatter, row-mapped
// copies, mask-token fill, column sums, loss, AdamW+EMA, norms.  All use 16-byte vector
// accesses on the contiguous feature dimension and grids sized in multiples of the SM count.
#include "common.cuh"

#include <stdarg.h>

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

void avj_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int avj_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

extern "C" int avj_version(void) { return AVJ_ABI_VERSION; }
extern "C" const char* avj_last_error_string(void) { return g_err; }
extern "C" int avj_device_ok(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10;
}

static inline int grid_for(int64_t work_items, int per_block, int waves = 8) {
  int64_t need = (work_items + per_block - 1) / per_block;
  int64_t cap = (int64_t)avj_num_sms() * waves;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// ------------------------------------------------------------------------------------------
// K1/K2 patch extraction: fp32 [B,C,T,H,W] -> rows of (c, dt, dh, dw) for the kept tokens.
// One thread moves 8 consecutive dw (32 B read, 16/32 B write).
// ------------------------------------------------------------------------------------------
template <typename TOut>
__global__ void patchify_kernel(const float* __restrict__ x, const int64_t* __restrict__ idx, TOut* __restrict__ out,
                                int B, int C, int T, int H, int W, int tub, int p, int K) {
  const int gh = H / p, gw = W / p;
  const int kdim = C * tub * p * p;
  const int vec_per_row = kdim / 8;
  const int64_t total = (int64_t)B * K * vec_per_row;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / vec_per_row;
    const int c8 = (int)(i % vec_per_row) * 8;
    const int b = (int)(r / K);
    const int tok = idx ? (int)idx[r] : (int)(r % K);
    const int tt = tok / (gh * gw), hh = (tok / gw) % gh, ww = tok % gw;
    const int dw = c8 % p;
    const int dh = (c8 / p) % p;
    const int dt = (c8 / (p * p)) % tub;
    const int ch = c8 / (p * p * tub);
    const float* src = x + ((((int64_t)b * C + ch) * T + (tt * tub + dt)) * H + (hh * p + dh)) * W + ww * p + dw;
    float v[8];
    load8<float>(src, v);
    store8<TOut>(out + r * kdim + c8, v);
  }
}

extern "C" int avj_patchify(const float* x, const int64_t* idx, void* out, int out_dtype,
                            int B, int C, int T, int H, int W, int tub, int patch, int K, void* stream) {
  AVJ_CHECK(patch % 8 == 0 && W % 8 == 0, "avj_patchify: patch and W must be multiples of 8");
  AVJ_CHECK(T % tub == 0 && H % patch == 0 && W % patch == 0, "avj_patchify: dims not divisible by patch");
  if (B * K == 0) return 0;
  AvjProfScope prof(AVJ_FAM_OTHER, (double)B * K * C * tub * patch * patch * (4 + (out_dtype == AVJ_BF16 ? 2 : 4)), stream, 1, B * K, C * tub * patch * patch);
  const int64_t total = (int64_t)B * K * (C * tub * patch * patch / 8);
  const int grid = grid_for(total, 256);
  if (out_dtype == AVJ_BF16)
    patchify_kernel<bf16><<<grid, 256, 0, as_stream(stream)>>>(x, idx, (bf16*)out, B, C, T, H, W, tub, patch, K);
  else
    patchify_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(x, idx, (float*)out, B, C, T, H, W, tub, patch, K);
  AVJ_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// K3 apply_masks gather / scatter-add.  The index is read once per row-chunk (not expanded
// to [B,K,D] like the reference's .repeat).
// ------------------------------------------------------------------------------------------
template <typename T, bool kBackward>
__global__ void gather_rows_kernel(const T* __restrict__ src, const int64_t* __restrict__ idx, T* __restrict__ dst,
                                   int B, int N, int K, int D) {
  const int vpr = D / 8;
  const int64_t total = (int64_t)B * K * vpr;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / vpr;
    const int c = (int)(i % vpr) * 8;
    const int b = (int)(r / K);
    const int64_t full = (int64_t)b * N + idx[r];
    float v[8];
    if (!kBackward) {
      load8<T>(src + full * D + c, v);
      store8<T>(dst + r * D + c, v);
    } else {
      float a[8];
      load8<T>(src + r * D + c, v);
      load8<T>(dst + full * D + c, a);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += v[j];
      store8<T>(dst + full * D + c, a);
    }
  }
}

extern "C" int avj_gather_rows_fwd(int dtype, const void* x, const int64_t* idx, void* out,
                                   int B, int N, int K, int D, void* stream) {
  AVJ_CHECK(D % 8 == 0, "avj_gather_rows_fwd: D must be a multiple of 8");
  if ((int64_t)B * K * D == 0) return 0;
  AvjProfScope prof(AVJ_FAM_OTHER, (double)B * K * D * 2 * (dtype == AVJ_BF16 ? 2 : 4), stream, 2, B * K, D);
  const int grid = grid_for((int64_t)B * K * (D / 8), 256);
  if (dtype == AVJ_BF16)
    gather_rows_kernel<bf16, false><<<grid, 256, 0, as_stream(stream)>>>((const bf16*)x, idx, (bf16*)out, B, N, K, D);
  else
    gather_rows_kernel<float, false><<<grid, 256, 0, as_stream(stream)>>>((const float*)x, idx, (float*)out, B, N, K, D);
  AVJ_LAUNCH_CHECK();
  return 0;
}

extern "C" int avj_gather_rows_bwd(int dtype, const void* dout, const int64_t* idx, void* dx,
                                   int B, int N, int K, int D, void* stream) {
  AVJ_CHECK(D % 8 == 0, "avj_gather_rows_bwd: D must be a multiple of 8");
  if ((int64_t)B * K * D == 0) return 0;
  AvjProfScope prof(AVJ_FAM_OTHER, ((double)B * K + (double)B * N) * D * (dtype == AVJ_BF16 ? 2 : 4), stream, 3, B * K, D);
  const int grid = grid_for((int64_t)B * K * (D / 8), 256);
  if (dtype == AVJ_BF16)
    gather_rows_kernel<bf16, true><<<grid, 256, 0, as_stream(stream)>>>((const bf16*)dout, idx, (bf16*)dx, B, N, K, D);
  else
    gather_rows_kernel<float, true><<<grid, 256, 0, as_stream(stream)>>>((const float*)dout, idx, (float*)dx, B, N, K, D);
  AVJ_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// row-mapped copy / cast / accumulate
// ------------------------------------------------------------------------------------------
template <typename TIn, typename TOut>
__global__ void copy_rows_kernel(const TIn* __restrict__ in, int ld_in, avj_rowmap imap,
                                 TOut* __restrict__ out, int ld_out, avj_rowmap omap,
                                 int rows, int D, int accumulate) {
  const int vpr = D / 8;
  const int64_t total = (int64_t)rows * vpr;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / vpr;
    const int c = (int)(i % vpr) * 8;
    float v[8];
    load8<TIn>(in + map_row(imap, r) * ld_in + c, v);
    TOut* o = out + map_row(omap, r) * ld_out + c;
    if (accumulate) {
      float a[8];
      load8<TOut>(o, a);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += a[j];
    }
    store8<TOut>(o, v);
  }
}

extern "C" int avj_copy_rows(const void* in, int in_dtype, int ld_in, avj_rowmap imap,
                             void* out, int out_dtype, int ld_out, avj_rowmap omap,
                             int rows, int D, int accumulate, void* stream) {
  AVJ_CHECK(D % 8 == 0 && ld_in % 8 == 0 && ld_out % 8 == 0, "avj_copy_rows: D/ld must be multiples of 8");
  if ((int64_t)rows * D == 0) return 0;
  AvjProfScope prof(AVJ_FAM_OTHER, (double)rows * D * ((in_dtype == AVJ_BF16 ? 2 : 4) + (out_dtype == AVJ_BF16 ? 2 : 4) * (accumulate ? 2 : 1)), stream, 4, rows, D);
  const int grid = grid_for((int64_t)rows * (D / 8), 256);
  cudaStream_t s = as_stream(stream);
#define LAUNCH(TI, TO) copy_rows_kernel<TI, TO><<<grid, 256, 0, s>>>((const TI*)in, ld_in, imap, (TO*)out, ld_out, omap, rows, D, accumulate)
  if (in_dtype == AVJ_F32 && out_dtype == AVJ_F32) LAUNCH(float, float);
  else if (in_dtype == AVJ_F32 && out_dtype == AVJ_BF16) LAUNCH(float, bf16);
  else if (in_dtype == AVJ_BF16 && out_dtype == AVJ_F32) LAUNCH(bf16, float);
  else LAUNCH(bf16, bf16);
#undef LAUNCH
  AVJ_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// K4 predictor target rows: x[map(r)] = mask_token + pos[idx[r]]
// ------------------------------------------------------------------------------------------
__global__ void fill_mask_tokens_kernel(const float* __restrict__ tok, const float* __restrict__ pos,
                                        const int64_t* __restrict__ idx, float* __restrict__ x, int ld,
                                        avj_rowmap map, int rows, int D) {
  const int vpr = D / 4;
  const int64_t total = (int64_t)rows * vpr;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / vpr;
    const int c = (int)(i % vpr) * 4;
    const float4 t = *reinterpret_cast<const float4*>(tok + c);
    const float4 p = *reinterpret_cast<const float4*>(pos + idx[r] * (int64_t)D + c);
    *reinterpret_cast<float4*>(x + map_row(map, r) * ld + c) = make_float4(t.x + p.x, t.y + p.y, t.z + p.z, t.w + p.w);
  }
}

extern "C" int avj_fill_mask_tokens(const float* mask_token, const float* pos, const int64_t* idx,
                                    float* x, int ld, avj_rowmap map, int rows, int D, void* stream) {
  AVJ_CHECK(D % 4 == 0, "avj_fill_mask_tokens: D must be a multiple of 4");
  if ((int64_t)rows * D == 0) return 0;
  AvjProfScope prof(AVJ_FAM_OTHER, (double)rows * D * 8, stream, 5, rows, D);
  const int grid = grid_for((int64_t)rows * (D / 4), 256);
  fill_mask_tokens_kernel<<<grid, 256, 0, as_stream(stream)>>>(mask_token, pos, idx, x, ld, map, rows, D);
  AVJ_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// column sum over mapped rows (deterministic two-pass: per-slab partials, then a fixed-order
// reduction).  Slab = 128 rows; one CTA handles a slab x 256 columns (8 cols per thread x 32
// threads x 8 row-lanes).
// ------------------------------------------------------------------------------------------
#define COLSUM_MAX_Y 64

extern "C" int64_t avj_colsum_ws_floats(int rows, int D) {
  (void)rows;
  return (int64_t)COLSUM_MAX_Y * D;
}

// grid (ceil(D/256), Y): block (32, 8) owns 256 columns x one row range; 4 independent 16-byte loads per
// thread are kept in flight.  Deterministic two-stage reduction (no atomics): partials -> ws[Y][D].
template <typename T>
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const T* __restrict__ in, int ld, avj_rowmap map, float* __restrict__ ws,
                      int rows, int D, int chunk) {
  __shared__ float sm[8][32][8 + 1];
  pdl_trigger();
  pdl_wait();
  const int c0 = (blockIdx.x * 32 + threadIdx.x) * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c0 < D) {
    const int r_beg = blockIdx.y * chunk, r_end = min(rows, r_beg + chunk);
    int r = r_beg + threadIdx.y;
    for (; r + 56 < r_end; r += 64) {                  // eight independent 16-byte loads in flight per thread
      float v[8][8];
#pragma unroll
      for (int u = 0; u < 8; ++u) load8<T>(in + map_row(map, r + 8 * u) * ld + c0, v[u]);
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[u][j];
    }
    for (; r < r_end; r += 8) {
      float v[8];
      load8<T>(in + map_row(map, r) * ld + c0, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sm[threadIdx.y][threadIdx.x][j] = acc[j];
  __syncthreads();
  if (threadIdx.y == 0 && c0 < D) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float s = 0.f;
#pragma unroll
      for (int y = 0; y < 8; ++y) s += sm[y][threadIdx.x][j];
      ws[(int64_t)blockIdx.y * D + c0 + j] = s;
    }
  }
}

__global__ void colsum_final_kernel(const float* __restrict__ ws, float* __restrict__ out, int ny, int D) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= D) return;
  float s = 0.f;
#pragma unroll 8
  for (int i = 0; i < ny; ++i) s += ws[(int64_t)i * D + c];
  out[c] += s;
}

extern "C" int avj_colsum(const void* in, int in_dtype, int ld, avj_rowmap map, float* out,
                          int rows, int D, float* ws, void* stream) {
  AVJ_CHECK(D % 8 == 0 && ld % 8 == 0, "avj_colsum: D/ld must be multiples of 8");
  if (rows == 0 || D == 0) return 0;
  AvjProfScope prof(AVJ_FAM_COLSUM, (double)rows * D * (in_dtype == AVJ_BF16 ? 2 : 4), stream, rows, D);
  const int gx = (D + 255) / 256;
  int ny = (6 * avj_num_sms() + gx - 1) / gx;                 // ~6 CTAs per SM
  if (ny > COLSUM_MAX_Y) ny = COLSUM_MAX_Y;
  if (ny > (rows + 63) / 64) ny = (rows + 63) / 64;           // at least 64 rows per CTA
  if (ny < 1) ny = 1;
  int chunk = (rows + ny - 1) / ny;
  chunk = (chunk + 7) / 8 * 8;
  ny = (rows + chunk - 1) / chunk;
  dim3 grid(gx, ny), block(32, 8);
  if (in_dtype == AVJ_BF16)
    avj_launch_pdl(colsum_partial_kernel<bf16>, grid, block, 0, as_stream(stream), (const bf16*)in, ld, map, ws, rows, D, chunk);
  else
    avj_launch_pdl(colsum_partial_kernel<float>, grid, block, 0, as_stream(stream), (const float*)in, ld, map, ws, rows, D, chunk);
  AVJ_LAUNCH_CHECK();
  avj_launch_pdl(colsum_final_kernel, dim3((D + 63) / 64), dim3(64), 0, as_stream(stream), ws, out, ny, D);
  AVJ_LAUNCH_CHECK();
  return 0;
}

// Two column sums over the same rows in one launch: blockIdx.x < gx1 serves (in1, D1), the rest (in2, D2).
template <typename T>
__global__ void __launch_bounds__(256)
colsum2_partial_kernel(const T* __restrict__ in1, int ld1, int D1, const T* __restrict__ in2, int ld2, int D2,
                       float* __restrict__ ws, int rows, int chunk, int gx1) {
  __shared__ float sm[8][32][8 + 1];
  pdl_trigger();
  pdl_wait();
  const bool second = (int)blockIdx.x >= gx1;
  const T* in = second ? in2 : in1;
  const int ld = second ? ld2 : ld1, D = second ? D2 : D1;
  const int c0 = (((int)blockIdx.x - (second ? gx1 : 0)) * 32 + threadIdx.x) * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c0 < D) {
    const int r_beg = blockIdx.y * chunk, r_end = min(rows, r_beg + chunk);
    int r = r_beg + threadIdx.y;
    for (; r + 56 < r_end; r += 64) {                  // eight independent 16-byte loads in flight per thread
      float v[8][8];
#pragma unroll
      for (int u = 0; u < 8; ++u) load8<T>(in + (int64_t)(r + 8 * u) * ld + c0, v[u]);
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[u][j];
    }
    for (; r < r_end; r += 8) {
      float v[8];
      load8<T>(in + (int64_t)r * ld + c0, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sm[threadIdx.y][threadIdx.x][j] = acc[j];
  __syncthreads();
  if (threadIdx.y == 0 && c0 < D) {
    float* w = ws + (int64_t)blockIdx.y * (D1 + D2) + (second ? D1 : 0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float s = 0.f;
#pragma unroll
      for (int y = 0; y < 8; ++y) s += sm[y][threadIdx.x][j];
      w[c0 + j] = s;
    }
  }
}

__global__ void colsum2_final_kernel(const float* __restrict__ ws, float* __restrict__ out1, float* __restrict__ out2,
                                     int ny, int D1, int D2) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= D1 + D2) return;
  float s = 0.f;
#pragma unroll 8
  for (int i = 0; i < ny; ++i) s += ws[(int64_t)i * (D1 + D2) + c];
  if (c < D1) out1[c] += s; else out2[c - D1] += s;
}

extern "C" int avj_colsum2(const void* in1, int ld1, int D1, float* out1, const void* in2, int ld2, int D2, float* out2,
                           int in_dtype, int rows, float* ws, void* stream) {
  AVJ_CHECK(D1 % 8 == 0 && ld1 % 8 == 0 && D2 % 8 == 0 && ld2 % 8 == 0, "avj_colsum2: D/ld must be multiples of 8");
  AVJ_CHECK(out1 && out2 && ws, "avj_colsum2: NULL output/workspace");
  if (rows == 0) return 0;
  AvjProfScope prof(AVJ_FAM_COLSUM, (double)rows * (D1 + D2) * (in_dtype == AVJ_BF16 ? 2 : 4), stream, rows, D1, D2);
  const int gx1 = (D1 + 255) / 256, gx = gx1 + (D2 + 255) / 256;
  int ny = (6 * avj_num_sms() + gx - 1) / gx;
  if (ny > COLSUM_MAX_Y) ny = COLSUM_MAX_Y;
  if (ny > (rows + 63) / 64) ny = (rows + 63) / 64;
  if (ny < 1) ny = 1;
  int chunk = (rows + ny - 1) / ny;
  chunk = (chunk + 7) / 8 * 8;
  ny = (rows + chunk - 1) / chunk;
  dim3 grid(gx, ny), block(32, 8);
  if (in_dtype == AVJ_BF16)
    avj_launch_pdl(colsum2_partial_kernel<bf16>, grid, block, 0, as_stream(stream), (const bf16*)in1, ld1, D1, (const bf16*)in2, ld2, D2, ws, rows, chunk, gx1);
  else
    avj_launch_pdl(colsum2_partial_kernel<float>, grid, block, 0, as_stream(stream), (const float*)in1, ld1, D1, (const float*)in2, ld2, D2, ws, rows, chunk, gx1);
  AVJ_LAUNCH_CHECK();
  avj_launch_pdl(colsum2_final_kernel, dim3((D1 + D2 + 63) / 64), dim3(64), 0, as_stream(stream), ws, out1, out2, ny, D1, D2);
  AVJ_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// K10 loss forward+backward in one pass, deterministic two-stage reduction.
// ------------------------------------------------------------------------------------------
#define LOSS_BLOCK 256
#define LOSS_PER_THREAD 16

extern "C" int64_t avj_loss_ws_floats(int64_t n) {
  int64_t blocks = (n + (int64_t)LOSS_BLOCK * LOSS_PER_THREAD - 1) / ((int64_t)LOSS_BLOCK * LOSS_PER_THREAD);
  int64_t cap = (int64_t)148 * 16;
  return (blocks < cap ? blocks : cap) + 1;
}

__global__ void loss_kernel(const float* __restrict__ z, const float* __restrict__ h, float* __restrict__ dz,
                            float* __restrict__ partial, int64_t n, float p, int mode, float beta, float gscale) {
  __shared__ float red[32];
  float acc = 0.f;
  const int64_t nv = n / 4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(z)[i];
    const float4 b = reinterpret_cast<const float4*>(h)[i];
    float d[4] = {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w};
    float g[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float ad = fabsf(d[j]);
      const float sg = (d[j] > 0.f) ? 1.f : ((d[j] < 0.f) ? -1.f : 0.f);
      if (mode == 1) {            // smooth-L1 (extra mode, not in the reference)
        if (ad < beta) { acc += 0.5f * d[j] * d[j] / beta; g[j] = d[j] / beta; }
        else { acc += ad - 0.5f * beta; g[j] = sg; }
      } else if (p == 1.0f) {
        acc += ad; g[j] = sg;
      } else if (p == 2.0f) {
        acc += ad * ad; g[j] = 2.f * d[j];
      } else {
        acc += powf(ad, p); g[j] = (ad > 0.f) ? p * powf(ad, p - 1.f) * sg : 0.f;
      }
      g[j] *= gscale;
    }
    if (dz) reinterpret_cast<float4*>(dz)[i] = make_float4(g[0], g[1], g[2], g[3]);
  }
  const float tot = block_sum(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}

__global__ void loss_final_kernel(const float* __restrict__ partial, int nparts, float scale, float* __restrict__ out) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) acc += partial[i];
  const float tot = block_sum(acc, red);
  if (threadIdx.x == 0) out[0] += tot * scale;
}

extern "C" int avj_loss_fwd_bwd(const float* z, const float* h, float* dz, float* loss_out,
                                int64_t n, int n_masks, float loss_exp, int mode, float beta,
                                float grad_scale, float* ws, void* stream) {
  AVJ_CHECK(n % 4 == 0 && n > 0, "avj_loss_fwd_bwd: n must be a positive multiple of 4");
  AvjProfScope prof(AVJ_FAM_OTHER, (double)n * 12, stream, 6, (int)(n >> 10), 1024);
  const int nparts = (int)(avj_loss_ws_floats(n) - 1);
  const float denom = (mode == 1) ? 1.0f : loss_exp;
  // d/dz [ (1/(n*n_masks*p)) sum |d|^p ] : the kernel's g is d|d|^p/dd, so fold 1/(n*n_masks*p) here
  const float gscale = grad_scale / ((float)n * (float)n_masks * denom);
  loss_kernel<<<nparts, LOSS_BLOCK, 0, as_stream(stream)>>>(z, h, dz, ws, n, loss_exp, mode, beta, gscale);
  AVJ_LAUNCH_CHECK();
  loss_final_kernel<<<1, 256, 0, as_stream(stream)>>>(ws, nparts, 1.0f / ((float)n * (float)n_masks * denom), loss_out);
  AVJ_LAUNCH_CHECK();
  return 0;
}

// token-variance regulariser value: one warp-column-group per (b, 32 features)
__global__ void reg_accumulate_kernel(const float* __restrict__ z, float* __restrict__ pstd, int B, int K, int D, float inv_masks) {
  // blockDim = (32, 8): x -> feature, y -> token lane
  __shared__ float s1[8][33], s2[8][33];
  const int b = blockIdx.y;
  const int d = blockIdx.x * 32 + threadIdx.x;
  float a = 0.f, q = 0.f;
  if (d < D) {
    // shifted sums (shift = first token) for a stable single pass
    const float sh = z[((int64_t)b * K) * D + d];
    for (int k = threadIdx.y; k < K; k += 8) {
      const float v = z[((int64_t)b * K + k) * D + d] - sh;
      a += v; q += v * v;
    }
  }
  s1[threadIdx.y][threadIdx.x] = a; s2[threadIdx.y][threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.y == 0 && d < D) {
    float sa = 0.f, sq = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) { sa += s1[y][threadIdx.x]; sq += s2[y][threadIdx.x]; }
    const float var = (sq - sa * sa / (float)K) / (float)(K - 1);     // unbiased, torch.var default
    pstd[(int64_t)b * D + d] += sqrtf(fmaxf(var, 0.f) + 0.0001f) * inv_masks;
  }
}

__global__ void reg_finish_kernel(const float* __restrict__ pstd, float* __restrict__ out, int n) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += fmaxf(1.0f - pstd[i], 0.f);
  const float tot = block_sum(acc, red);
  if (threadIdx.x == 0) out[0] = tot / (float)n;
}

extern "C" int avj_reg_accumulate(const float* z, float* pstd, int B, int K, int D, int n_masks, void* stream) {
  AVJ_CHECK(K >= 2, "avj_reg_accumulate: need at least 2 tokens");
  dim3 grid((D + 31) / 32, B), block(32, 8);
  reg_accumulate_kernel<<<grid, block, 0, as_stream(stream)>>>(z, pstd, B, K, D, 1.0f / (float)n_masks);
  AVJ_LAUNCH_CHECK();
  return 0;
}

extern "C" int avj_reg_finish(const float* pstd, float* loss_reg, int n, void* stream) {
  reg_finish_kernel<<<1, 1024, 0, as_stream(stream)>>>(pstd, loss_reg, n);
  AVJ_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// K11-K13 AdamW + unscale/clip + EMA + zero-grad + bf16 shadows, one pass over a flat range.
// Algorithmic bytes/param: read p,g,m,v (16) + write p,m,v (12) [+ g=0 (4)] [+ target r/w (8)]
// [+ shadows (2+2)].
// ------------------------------------------------------------------------------------------
__global__ void adamw_ema_kernel(avj_adamw_args a, float bc1, float bc2_sqrt) {
  const float gs = a.scale_ptr ? a.scale_ptr[0] : 1.0f;
  const float step_size = a.lr / bc1;
  const float decay = 1.0f - a.lr * a.wd;
  const int64_t nv = a.n / 4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
    float4 p4 = reinterpret_cast<float4*>(a.p)[i];
    float p[4] = {p4.x, p4.y, p4.z, p4.w};
    if (!a.skip_update) {
      float4 g4 = reinterpret_cast<float4*>(a.g)[i];
      float4 m4 = reinterpret_cast<float4*>(a.m)[i];
      float4 v4 = reinterpret_cast<float4*>(a.v)[i];
      float g[4] = {g4.x * gs, g4.y * gs, g4.z * gs, g4.w * gs};
      float m[4] = {m4.x, m4.y, m4.z, m4.w};
      float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        p[j] *= decay;
        m[j] = m[j] + (1.0f - a.beta1) * (g[j] - m[j]);
        v[j] = v[j] * a.beta2 + (1.0f - a.beta2) * g[j] * g[j];
        const float denom = sqrtf(v[j]) / bc2_sqrt + a.eps;
        p[j] -= step_size * (m[j] / denom);
      }
      reinterpret_cast<float4*>(a.p)[i] = make_float4(p[0], p[1], p[2], p[3]);
      reinterpret_cast<float4*>(a.m)[i] = make_float4(m[0], m[1], m[2], m[3]);
      reinterpret_cast<float4*>(a.v)[i] = make_float4(v[0], v[1], v[2], v[3]);
      if (a.zero_grad) reinterpret_cast<float4*>(a.g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (a.p_lp) {
      uint2 o; o.x = pack_bf16x2(p[0], p[1]); o.y = pack_bf16x2(p[2], p[3]);
      reinterpret_cast<uint2*>(a.p_lp)[i] = o;
    }
    if (a.target) {
      float4 k4 = reinterpret_cast<float4*>(a.target)[i];
      float k[4] = {k4.x, k4.y, k4.z, k4.w};
      const float om = 1.0f - a.ema_m;
#pragma unroll
      for (int j = 0; j < 4; ++j) k[j] = k[j] * a.ema_m + om * p[j];
      reinterpret_cast<float4*>(a.target)[i] = make_float4(k[0], k[1], k[2], k[3]);
      if (a.target_lp) {
        uint2 o; o.x = pack_bf16x2(k[0], k[1]); o.y = pack_bf16x2(k[2], k[3]);
        reinterpret_cast<uint2*>(a.target_lp)[i] = o;
      }
    }
  }
}

extern "C" int avj_adamw_ema_step(const avj_adamw_args* a, void* stream) {
  AVJ_CHECK(a != nullptr, "avj_adamw_ema_step: null args");
  AVJ_CHECK(a->n % 4 == 0, "avj_adamw_ema_step: range length must be a multiple of 4 (pad the flat buffer)");
  if (a->n == 0) return 0;
  AVJ_CHECK(a->skip_update || (a->g && a->m && a->v), "avj_adamw_ema_step: g/m/v required");
  const double bc1 = 1.0 - pow((double)a->beta1, (double)a->step);
  const double bc2 = 1.0 - pow((double)a->beta2, (double)a->step);
  AvjProfScope prof(AVJ_FAM_OPTIM, (double)a->n * (a->skip_update ? 0 : 28) + (double)a->n * (a->target ? 8 : 0) +
                                       (double)a->n * (a->zero_grad ? 4 : 0) + (double)a->n * ((a->p_lp ? 2 : 0) + (a->target_lp ? 2 : 0)), stream);
  const int grid = grid_for(a->n / 4, 256, 8);
  adamw_ema_kernel<<<grid, 256, 0, as_stream(stream)>>>(*a, (float)bc1, (float)sqrt(bc2));
  AVJ_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// sum of squares / clip coefficient / cast
// ------------------------------------------------------------------------------------------
extern "C" int64_t avj_sumsq_ws_floats(int64_t n) { (void)n; return 148 * 8 + 1; }

__global__ void sumsq_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ partial) {
  __shared__ float red[32];
  float acc = 0.f;
  const int64_t nv = n / 4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(x)[i];
    acc += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
  }
  const float tot = block_sum(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}

__global__ void sumsq_final_kernel(const float* __restrict__ partial, int nparts, float* __restrict__ out) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) acc += partial[i];
  const float tot = block_sum(acc, red);
  if (threadIdx.x == 0) out[0] = tot;
}

extern "C" int avj_sumsq(const float* x, int64_t n, float* out, float* ws, void* stream) {
  AVJ_CHECK(n % 4 == 0, "avj_sumsq: n must be a multiple of 4");
  const int nparts = 148 * 8;
  sumsq_kernel<<<nparts, 256, 0, as_stream(stream)>>>(x, n, ws);
  AVJ_LAUNCH_CHECK();
  sumsq_final_kernel<<<1, 256, 0, as_stream(stream)>>>(ws, nparts, out);
  AVJ_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// Per-parameter statistics over a flat buffer in ONE pass (grad_logger / adamw_logger,
// src/utils/logging.py:91-118): out[s] += sum over segment s of x^2 (mode 0) or |x| (mode 1).
// seg_off[0..nseg] are ascending element offsets (multiples of 4; the zero padding between
// parameters belongs to the preceding segment and contributes nothing).  Each block owns a
// 8192-element chunk: when the chunk lies inside one segment (the common case -- weights are
// 10^5..10^6 elements) it does a block reduction and one fp64 atomic, otherwise warps that lie
// inside one segment reduce by shuffle and the few straddling lanes add individually.
// ------------------------------------------------------------------------------------------
#define SEG_CHUNK 8192
__device__ __forceinline__ int seg_of(const int64_t* __restrict__ off, int nseg, int64_t e) {
  int lo = 0, hi = nseg;                    // largest s with off[s] <= e
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (off[mid] <= e) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void segment_stats_kernel(const float* __restrict__ x, const int64_t* __restrict__ off, int nseg, int mode,
                                     double* __restrict__ out) {
  __shared__ float red[32];
  const int64_t total = off[nseg];
  const int64_t c0 = (int64_t)blockIdx.x * SEG_CHUNK;
  const int64_t c1 = min(c0 + (int64_t)SEG_CHUNK, total);
  if (c0 >= total) return;
  const int s_first = seg_of(off, nseg, c0), s_last = seg_of(off, nseg, c1 - 1);
  if (s_first == s_last) {
    float acc = 0.f;
    for (int64_t e = c0 + 4 * (int64_t)threadIdx.x; e < c1; e += 4 * (int64_t)blockDim.x) {
      const float4 a = *reinterpret_cast<const float4*>(x + e);
      acc += mode == 0 ? (a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w) : (fabsf(a.x) + fabsf(a.y) + fabsf(a.z) + fabsf(a.w));
    }
    const float tot = block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(out + s_first, (double)tot);
    return;
  }
  for (int64_t e0 = c0; e0 < c1; e0 += 4 * (int64_t)blockDim.x) {
    const int64_t e = e0 + 4 * (int64_t)threadIdx.x;
    float acc = 0.f;
    int sg = -1;
    if (e < c1) {
      const float4 a = *reinterpret_cast<const float4*>(x + e);
      acc = mode == 0 ? (a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w) : (fabsf(a.x) + fabsf(a.y) + fabsf(a.z) + fabsf(a.w));
      sg = seg_of(off, nseg, e);
    }
    const int sg0 = __shfl_sync(0xffffffffu, sg, 0);
    if (__all_sync(0xffffffffu, sg == sg0)) {
      const float w = warp_sum(acc);
      if ((threadIdx.x & 31) == 0 && sg0 >= 0) atomicAdd(out + sg0, (double)w);
    } else if (sg >= 0) {
      atomicAdd(out + sg, (double)acc);
    }
  }
}

extern "C" int avj_segment_stats(const float* x, const int64_t* seg_off, int nseg, int64_t total, int mode, double* out,
                                 void* stream) {
  AVJ_CHECK(x && seg_off && out, "avj_segment_stats: NULL argument");
  AVJ_CHECK(mode == 0 || mode == 1, "avj_segment_stats: mode must be 0 (sum of squares) or 1 (sum of |x|)");
  AVJ_CHECK(total % 4 == 0, "avj_segment_stats: total must be a multiple of 4");
  if (nseg <= 0 || total == 0) return 0;
  AvjProfScope prof(AVJ_FAM_OTHER, 4.0 * (double)total, stream, 9, (int)(total >> 20), nseg);
  const int64_t blocks = (total + SEG_CHUNK - 1) / SEG_CHUNK;
  segment_stats_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(x, seg_off, nseg, mode, out);
  AVJ_LAUNCH_CHECK();
  return 0;
}

__global__ void clip_coef_kernel(const float* sumsq, float max_norm, float inv_scale, float* coef) {
  float c = inv_scale;
  if (max_norm > 0.f) {
    const float norm = sqrtf(sumsq[0]) * inv_scale;
    c *= fminf(1.0f, max_norm / (norm + 1e-6f));
  }
  coef[0] = c;
}

extern "C" int avj_clip_coef(const float* sumsq, float max_norm, float inv_loss_scale, float* coef, void* stream) {
  clip_coef_kernel<<<1, 1, 0, as_stream(stream)>>>(sumsq, max_norm, inv_loss_scale, coef);
  AVJ_LAUNCH_CHECK();
  return 0;
}

template <typename TOut>
__global__ void cast_kernel(const float* __restrict__ in, TOut* __restrict__ out, int64_t n) {
  const int64_t nv = n / 8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
    float v[8];
    load8<float>(in + i * 8, v);
    store8<TOut>(out + i * 8, v);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int64_t i = nv * 8; i < n; ++i) out[i] = from_f32<TOut>(in[i]);
}

extern "C" int avj_cast(const float* in, void* out, int out_dtype, int64_t n, void* stream) {
  if (n == 0) return 0;
  AvjProfScope prof(AVJ_FAM_OTHER, (double)n * (4 + (out_dtype == AVJ_BF16 ? 2 : 4)), stream, 7, (int)(n >> 10), 1024);
  const int grid = grid_for(n / 8 + 1, 256);
  if (out_dtype == AVJ_BF16) cast_kernel<bf16><<<grid, 256, 0, as_stream(stream)>>>(in, (bf16*)out, n);
  else cast_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(in, (float*)out, n);
  AVJ_LAUNCH_CHECK();
  return 0;
}

extern "C" int avj_memset_zero(void* ptr, int64_t nbytes, void* stream) {
  if (nbytes <= 0) return 0;
  AvjProfScope prof(AVJ_FAM_OTHER, (double)nbytes, stream, 8, (int)(nbytes >> 10), 1024);
  AVJ_CUDA(cudaMemsetAsync(ptr, 0, (size_t)nbytes, as_stream(stream)));
  return 0;
}
