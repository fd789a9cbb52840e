// K5 LayerNorm forward/backward over the fp32 residual stream (HBM-bound).
// One warp owns one row; the row lives in registers (float4 per lane, D <= 2048), statistics
// are warp-shuffle reductions.  Algorithmic bytes/row: fwd 4D read + sizeof(y)*D write;
// bwd sizeof(dy)*D + 4D (x) [+4D dres] read, 4D [+2D] write.
#include "common.cuh"

#define LN_WARPS 8
#define LN_MAX_BLOCKS (148 * 2)

template <typename TY, int NV4>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     TY* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                     int rows, int D, float eps) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = D / 4;
  for (int64_t r = (int64_t)blockIdx.x * LN_WARPS + warp; r < rows; r += (int64_t)gridDim.x * LN_WARPS) {
    const float4* xr = reinterpret_cast<const float4*>(x + r * D);
    float4 v[NV4];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int c = lane + i * 32;
      v[i] = (c < nvec) ? xr[c] : make_float4(0.f, 0.f, 0.f, 0.f);
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
    const float mean = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
        const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
        q += a * a + b * b + cc * cc + d * d;
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
    if (lane == 0) {
      if (mean_out) mean_out[r] = mean;
      if (rstd_out) rstd_out[r] = rstd;
    }
    TY* yr = y + r * D;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
        float4 g = gamma ? reinterpret_cast<const float4*>(gamma)[c] : make_float4(1.f, 1.f, 1.f, 1.f);
        float4 b = beta ? reinterpret_cast<const float4*>(beta)[c] : make_float4(0.f, 0.f, 0.f, 0.f);
        const float o0 = (v[i].x - mean) * rstd * g.x + b.x;
        const float o1 = (v[i].y - mean) * rstd * g.y + b.y;
        const float o2 = (v[i].z - mean) * rstd * g.z + b.z;
        const float o3 = (v[i].w - mean) * rstd * g.w + b.w;
        if (sizeof(TY) == 4) {
          reinterpret_cast<float4*>(yr)[c] = make_float4(o0, o1, o2, o3);
        } else {
          uint2 o; o.x = pack_bf16x2(o0, o1); o.y = pack_bf16x2(o2, o3);
          reinterpret_cast<uint2*>(yr)[c] = o;
        }
      }
    }
  }
}

template <typename TY>
static int launch_ln_fwd(const float* x, const float* gamma, const float* beta, TY* y, float* mean, float* rstd,
                         int rows, int D, float eps, cudaStream_t s) {
  const int nv4 = (D / 4 + 31) / 32;
  int grid = (rows + LN_WARPS - 1) / LN_WARPS;
  const int cap = avj_num_sms() * 8;
  if (grid > cap) grid = cap;
#define LN_CASE(NV) case NV: avj_launch_pdl(layernorm_fwd_kernel<TY, NV>, dim3(grid), dim3(LN_WARPS * 32), 0, s, x, gamma, beta, y, mean, rstd, rows, D, eps); break;
  switch (nv4) {
    LN_CASE(1) LN_CASE(2) LN_CASE(3) LN_CASE(4) LN_CASE(5) LN_CASE(6) LN_CASE(7) LN_CASE(8)
    LN_CASE(9) LN_CASE(10) LN_CASE(11) LN_CASE(12) LN_CASE(13) LN_CASE(14) LN_CASE(15) LN_CASE(16)
    default: avj_set_error("avj_layernorm_fwd: D=%d too large (max 2048)", D); return 1;
  }
#undef LN_CASE
  AVJ_LAUNCH_CHECK();
  return 0;
}

extern "C" int avj_layernorm_fwd(const float* x, const float* gamma, const float* beta,
                                 void* y, int y_dtype, float* mean, float* rstd,
                                 int rows, int D, float eps, void* stream) {
  AVJ_CHECK(D % 4 == 0 && D > 0, "avj_layernorm_fwd: D must be a positive multiple of 4");
  if (rows == 0) return 0;
  AvjProfScope prof(AVJ_FAM_LN_FWD, (double)rows * D * (4 + (y_dtype == AVJ_BF16 ? 2 : 4)), stream, rows, D);
  if (y_dtype == AVJ_BF16) return launch_ln_fwd<bf16>(x, gamma, beta, (bf16*)y, mean, rstd, rows, D, eps, as_stream(stream));
  return launch_ln_fwd<float>(x, gamma, beta, (float*)y, mean, rstd, rows, D, eps, as_stream(stream));
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
extern "C" int64_t avj_layernorm_bwd_ws_floats(int rows, int D) {
  (void)rows;
  return (int64_t)LN_MAX_BLOCKS * 3 * D;
}

template <typename TDY>
__device__ __forceinline__ float4 load4_as_f32(const TDY* p, int c);
template <>
__device__ __forceinline__ float4 load4_as_f32<float>(const float* p, int c) {
  return reinterpret_cast<const float4*>(p)[c];
}
template <>
__device__ __forceinline__ float4 load4_as_f32<bf16>(const bf16* p, int c) {
  const uint2 u = reinterpret_cast<const uint2*>(p)[c];
  float4 f;
  unpack_bf16x2(u.x, f.x, f.y);
  unpack_bf16x2(u.y, f.z, f.w);
  return f;
}

// One warp per row, rows strided over the grid.  Registers hold only the row being processed (x, dy, and
// the residual gradient, all requested up front so a row costs ONE memory round trip); the per-column
// d(gamma)/d(beta) partial sums live in a warp-private shared-memory slab (each lane owns its columns,
// so no synchronisation is needed until the final cross-warp reduction).  That keeps the kernel at
// <= 96 registers, i.e. 2-3 blocks (16-24 warps) per SM instead of one.
template <typename TDY, typename TLP, int NV4>
__global__ void __launch_bounds__(LN_WARPS * 32, 2)
layernorm_bwd_kernel(const TDY* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ dres,
                     float* __restrict__ dx, TLP* __restrict__ dx_lp, float* __restrict__ ws, int want_dgamma,
                     int want_colsum, int rows, int D) {
  extern __shared__ float sm[];   // [LN_WARPS][3][D] partials: d(gamma), d(beta), column sum of dx
  pdl_trigger();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = D / 4;
  float4* sg = reinterpret_cast<float4*>(sm + (size_t)warp * 3 * D);
  float4* sb = sg + nvec;
  float4* sc = sb + nvec;
  if (want_dgamma || want_colsum) {
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
        sg[c] = make_float4(0.f, 0.f, 0.f, 0.f); sb[c] = make_float4(0.f, 0.f, 0.f, 0.f); sc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }

  pdl_wait();                     // everything above touched shared memory only
  for (int64_t r = (int64_t)blockIdx.x * LN_WARPS + warp; r < rows; r += (int64_t)gridDim.x * LN_WARPS) {
    const float4* xr = reinterpret_cast<const float4*>(x + r * D);
    const TDY* dyr = dy + r * D;
    float4 xv[NV4], d[NV4], dr[NV4];
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
        xv[i] = xr[c];
        d[i] = load4_as_f32<TDY>(dyr, c);
        dr[i] = dres ? reinterpret_cast<const float4*>(dres + r * D)[c] : make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        xv[i] = make_float4(0.f, 0.f, 0.f, 0.f); d[i] = xv[i]; dr[i] = xv[i];
      }
    }
    const float mu = mean[r], rs = rstd[r];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
        const float4 gm = gamma ? reinterpret_cast<const float4*>(gamma)[c] : make_float4(1.f, 1.f, 1.f, 1.f);
        // xv <- normalised x, d stays dy, g = dy * gamma folded into the sums
        xv[i] = make_float4((xv[i].x - mu) * rs, (xv[i].y - mu) * rs, (xv[i].z - mu) * rs, (xv[i].w - mu) * rs);
        if (want_dgamma) {
          float4 a = sg[c], bb = sb[c];
          a.x += d[i].x * xv[i].x; a.y += d[i].y * xv[i].y; a.z += d[i].z * xv[i].z; a.w += d[i].w * xv[i].w;
          bb.x += d[i].x; bb.y += d[i].y; bb.z += d[i].z; bb.w += d[i].w;
          sg[c] = a; sb[c] = bb;
        }
        d[i] = make_float4(d[i].x * gm.x, d[i].y * gm.y, d[i].z * gm.z, d[i].w * gm.w);
        s1 += d[i].x + d[i].y + d[i].z + d[i].w;
        s2 += d[i].x * xv[i].x + d[i].y * xv[i].y + d[i].z * xv[i].z + d[i].w * xv[i].w;
      }
    }
    const float m1 = warp_sum(s1) / (float)D;
    const float m2 = warp_sum(s2) / (float)D;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
        const float4 o = make_float4(dr[i].x + rs * (d[i].x - m1 - xv[i].x * m2), dr[i].y + rs * (d[i].y - m1 - xv[i].y * m2),
                                     dr[i].z + rs * (d[i].z - m1 - xv[i].z * m2), dr[i].w + rs * (d[i].w - m1 - xv[i].w * m2));
        reinterpret_cast<float4*>(dx + r * D)[c] = o;
        if (want_colsum) {
          float4 a = sc[c];
          a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
          sc[c] = a;
        }
        if (dx_lp) {
          if (sizeof(TLP) == 4) {
            reinterpret_cast<float4*>(dx_lp + r * D)[c] = o;
          } else {
            uint2 u; u.x = pack_bf16x2(o.x, o.y); u.y = pack_bf16x2(o.z, o.w);
            reinterpret_cast<uint2*>(dx_lp + r * D)[c] = u;
          }
        }
      }
    }
  }

  if (want_dgamma || want_colsum) {
    __syncthreads();
    for (int c = threadIdx.x; c < 3 * D; c += blockDim.x) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < LN_WARPS; ++w) t += sm[(size_t)w * 3 * D + c];
      ws[(size_t)blockIdx.x * 3 * D + c] = t;
    }
  }
}

// cross-block reduction of the [nblocks][3D] partials: 32 columns x 32 row groups per block, four independent loads in
// flight per thread (the partials sit in L2: the reduction is latency-, not bandwidth-bound)
__global__ void __launch_bounds__(1024)
layernorm_bwd_final_kernel(const float* __restrict__ ws, int nblocks, int D,
                           float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dcolsum) {
  __shared__ float red[32][33];
  pdl_trigger();
  pdl_wait();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
  if (c < 3 * D) {
    const size_t st = (size_t)3 * D;
    int b = ty;
    for (; b + 96 < nblocks; b += 128) {
      t0 += ws[(size_t)b * st + c]; t1 += ws[(size_t)(b + 32) * st + c];
      t2 += ws[(size_t)(b + 64) * st + c]; t3 += ws[(size_t)(b + 96) * st + c];
    }
    for (; b < nblocks; b += 32) t0 += ws[(size_t)b * st + c];
  }
  red[ty][tx] = (t0 + t1) + (t2 + t3);
  __syncthreads();
  if (ty == 0 && c < 3 * D) {
    float v = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) v += red[j][tx];
    if (c < D) { if (dgamma) dgamma[c] += v; }
    else if (c < 2 * D) { if (dbeta) dbeta[c - D] += v; }
    else { if (dcolsum) dcolsum[c - 2 * D] += v; }
  }
}

template <typename TDY, typename TLP>
static int launch_ln_bwd(const TDY* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                         const float* dres, float* dx, TLP* dx_lp, float* dgamma, float* dbeta, float* dcolsum, float* ws,
                         int rows, int D, cudaStream_t s) {
  const int nv4 = (D / 4 + 31) / 32;
  int grid = (rows + LN_WARPS - 1) / LN_WARPS;
  if (grid > LN_MAX_BLOCKS) grid = LN_MAX_BLOCKS;
  const int want = (dgamma != nullptr || dbeta != nullptr) ? 1 : 0;
  const int want_cs = dcolsum != nullptr ? 1 : 0;
  const size_t smem = (want || want_cs) ? (size_t)LN_WARPS * 3 * D * sizeof(float) : 0;
#define LNB_CASE(NV)                                                                                          \
  case NV: {                                                                                                  \
    auto k = layernorm_bwd_kernel<TDY, TLP, NV>;                                                              \
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
    cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);    \
    avj_launch_pdl(k, dim3(grid), dim3(LN_WARPS * 32), smem, s, dy, x, gamma, mean, rstd, dres, dx, dx_lp, ws, want, want_cs, rows, D);  \
  } break;
  switch (nv4) {
    LNB_CASE(1) LNB_CASE(2) LNB_CASE(3) LNB_CASE(4) LNB_CASE(5) LNB_CASE(6) LNB_CASE(7) LNB_CASE(8)
    LNB_CASE(9) LNB_CASE(10) LNB_CASE(11) LNB_CASE(12) LNB_CASE(13) LNB_CASE(14) LNB_CASE(15) LNB_CASE(16)
    default: avj_set_error("avj_layernorm_bwd: D=%d too large (max 2048)", D); return 1;
  }
#undef LNB_CASE
  AVJ_LAUNCH_CHECK();
  if (want || want_cs) {
    avj_launch_pdl(layernorm_bwd_final_kernel, dim3((3 * D + 31) / 32), dim3(1024), 0, s, ws, grid, D, dgamma, dbeta, dcolsum);
    AVJ_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int avj_layernorm_bwd(const void* dy, int dy_dtype, const float* x, const float* gamma,
                                 const float* mean, const float* rstd, const float* dres_in,
                                 float* dx_out, void* dx_lp, int lp_dtype,
                                 float* dgamma, float* dbeta, float* dcolsum, float* ws,
                                 int rows, int D, void* stream) {
  AVJ_CHECK(D % 4 == 0 && D > 0, "avj_layernorm_bwd: D must be a positive multiple of 4");
  AVJ_CHECK(!(dgamma || dbeta || dcolsum) || ws, "avj_layernorm_bwd: workspace required for dgamma/dbeta/dcolsum");
  if (rows == 0) return 0;
  cudaStream_t s = as_stream(stream);
  AvjProfScope prof(AVJ_FAM_LN_BWD, (double)rows * D * ((dy_dtype == AVJ_BF16 ? 2 : 4) + 4 + (dres_in ? 4 : 0) + 4 +
                                                         (dx_lp ? (lp_dtype == AVJ_BF16 ? 2 : 4) : 0)), stream, rows, D, dcolsum ? 1 : 0);
  if (dy_dtype == AVJ_BF16) {
    if (dx_lp && lp_dtype == AVJ_F32)
      return launch_ln_bwd<bf16, float>((const bf16*)dy, x, gamma, mean, rstd, dres_in, dx_out, (float*)dx_lp, dgamma, dbeta, dcolsum, ws, rows, D, s);
    return launch_ln_bwd<bf16, bf16>((const bf16*)dy, x, gamma, mean, rstd, dres_in, dx_out, (bf16*)dx_lp, dgamma, dbeta, dcolsum, ws, rows, D, s);
  }
  if (dx_lp && lp_dtype == AVJ_BF16)
    return launch_ln_bwd<float, bf16>((const float*)dy, x, gamma, mean, rstd, dres_in, dx_out, (bf16*)dx_lp, dgamma, dbeta, dcolsum, ws, rows, D, s);
  return launch_ln_bwd<float, float>((const float*)dy, x, gamma, mean, rstd, dres_in, dx_out, (float*)dx_lp, dgamma, dbeta, dcolsum, ws, rows, D, s);
}
