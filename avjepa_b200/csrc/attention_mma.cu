// K7 tensor-core attention for bf16: flash-style forward (online softmax, never materialises
// the N x N score matrix) and a two-pass deterministic backward (dQ by query block, dK/dV by key
// block; P recomputed from the saved log-sum-exp; no atomics).  Non-causal, no mask, any N.
//
// Reads q/k/v in place from the qkv-Linear output [B, N, 3, H, hd] with cp.async 16-byte copies
// (double-buffered K/V tiles); head dims that are not MMA-friendly (24, 80) are zero-padded to
// HP in SHARED MEMORY only -- HBM traffic stays at hd.  Math: mma.sync.m16n8k16 bf16 -> fp32,
// softmax statistics in fp32 with exp2f.  CTA = 4 warps x 16 rows, 64-row x 64-col tiles.
//
// Roofline: tensor-bound for N >= ~512.  Algorithmic FLOPs: fwd 4*N^2*hd per (b, h);
// bwd 14*N^2*hd here (two extra QK^T / dO V^T recomputations vs. the fused 10*N^2*hd schedule
// buy determinism and zero atomics).
#include "common.cuh"

#define FA_BM 64
#define FA_BN 64
#define FA_WARPS 4
#define FA_THREADS (FA_WARPS * 32)

__device__ __forceinline__ uint32_t fa_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Stage `FA_BM` rows x hd cols (bf16) of a strided matrix into smem [FA_BM][HP+8]; rows >= nvalid and
// columns in [hd, HP) are zero-filled with plain stores.
template <int HP>
__device__ __forceinline__ void stage_tile(bf16* __restrict__ dst, const bf16* __restrict__ src, int64_t row_stride,
                                           int row0, int nrows_total, int hd) {
  constexpr int LD = HP + 8;
  constexpr int CH = HP / 8;                       // 16-byte chunks per row (incl. padding chunks)
  const int hd_ch = hd / 8;
  for (int e = threadIdx.x; e < FA_BM * CH; e += FA_THREADS) {
    const int r = e / CH, c = e % CH;
    bf16* d = dst + r * LD + c * 8;
    if (row0 + r < nrows_total && c < hd_ch) {
      cp_async16(fa_smem(d), src + (int64_t)(row0 + r) * row_stride + c * 8);
    } else {
      *reinterpret_cast<uint4*>(d) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
}

// A-operand fragments (16 rows x HP) of this warp's row slab, kept in registers.
template <int HP>
__device__ __forceinline__ void load_a_frags(uint32_t (&a)[HP / 16][4], const bf16* tile, int warp_row0, int lane) {
  constexpr int LD = HP + 8;
  const int r = warp_row0 + (lane & 7) + 8 * ((lane >> 3) & 1);
  const int c = 8 * (lane >> 4);
#pragma unroll
  for (int k = 0; k < HP / 16; ++k) ldsm_x4(a[k], fa_smem(tile + r * LD + k * 16 + c));
}

// acc[nt][4] (16 x 64) += A(16 x HP, regs) . Bt^T where Bt is smem [64 rows(n)][HP(k)] (k contiguous)
template <int HP>
__device__ __forceinline__ void gemm_a_regs_bt(float (&acc)[8][4], const uint32_t (&a)[HP / 16][4], const bf16* bt, int lane) {
  constexpr int LD = HP + 8;
  const int n = (lane & 7) + 8 * (lane >> 4);
  const int kc = 8 * ((lane >> 3) & 1);
#pragma unroll
  for (int k = 0; k < HP / 16; ++k) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {      // pairs of n-tiles
      uint32_t b[4];
      ldsm_x4(b, fa_smem(bt + (np * 16 + n) * LD + k * 16 + kc));
      mma16816(acc[2 * np], a[k], b[0], b[1]);
      mma16816(acc[2 * np + 1], a[k], b[2], b[3]);
    }
  }
}

// acc[nt][4] (16 x HP) += P(16 x 64, packed bf16 A frags) . Bn where Bn is smem [64 rows(k)][HP(n)] (n contiguous)
template <int HP>
__device__ __forceinline__ void gemm_p_bn(float (&acc)[HP / 8][4], const uint32_t (&p)[4][4], const bf16* bn, int lane) {
  constexpr int LD = HP + 8;
  const int kr = (lane & 7) + 8 * ((lane >> 3) & 1);
  const int nc = 8 * (lane >> 4);
#pragma unroll
  for (int k = 0; k < 4; ++k) {           // 64 keys = 4 k16 steps
#pragma unroll
    for (int np = 0; np < HP / 16; ++np) {
      uint32_t b[4];
      ldsm_x4_t(b, fa_smem(bn + (k * 16 + kr) * LD + np * 16 + nc));
      mma16816(acc[2 * np], p[k], b[0], b[1]);
      mma16816(acc[2 * np + 1], p[k], b[2], b[3]);
    }
  }
}

__device__ __forceinline__ void pack_p(uint32_t (&p)[4][4], const float (&s)[8][4]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    p[k][0] = pack_bf16x2(s[2 * k][0], s[2 * k][1]);
    p[k][1] = pack_bf16x2(s[2 * k][2], s[2 * k][3]);
    p[k][2] = pack_bf16x2(s[2 * k + 1][0], s[2 * k + 1][1]);
    p[k][3] = pack_bf16x2(s[2 * k + 1][2], s[2 * k + 1][3]);
  }
}

// ============================================================================ forward
template <int HP>
__global__ void __launch_bounds__(FA_THREADS)
fa_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, float* __restrict__ lse,
              int N, int H, int hd, float scale_log2) {
  constexpr int LD = HP + 8;
  extern __shared__ __align__(16) uint8_t fa_smem_raw[];
  bf16* sQ = reinterpret_cast<bf16*>(fa_smem_raw);
  bf16* sK = sQ + FA_BM * LD;            // [2][FA_BN][LD]
  bf16* sV = sK + 2 * FA_BN * LD;        // [2][FA_BN][LD]
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * FA_BM;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t rs = 3 * (int64_t)H * hd;
  const bf16* qb = qkv + (int64_t)b * N * rs + (int64_t)h * hd;
  const bf16* kb = qb + (int64_t)H * hd;
  const bf16* vb = kb + (int64_t)H * hd;

  stage_tile<HP>(sQ, qb, rs, q0, N, hd);
  stage_tile<HP>(sK, kb, rs, 0, N, hd);
  stage_tile<HP>(sV, vb, rs, 0, N, hd);
  cp_async_commit();

  float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
  float o[HP / 8][4];
#pragma unroll
  for (int i = 0; i < HP / 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  uint32_t qf[HP / 16][4];

  const int n_tiles = (N + FA_BN - 1) / FA_BN;
  for (int t = 0; t < n_tiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < n_tiles) {
      stage_tile<HP>(sK + (buf ^ 1) * FA_BN * LD, kb, rs, (t + 1) * FA_BN, N, hd);
      stage_tile<HP>(sV + (buf ^ 1) * FA_BN * LD, vb, rs, (t + 1) * FA_BN, N, hd);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (t == 0) load_a_frags<HP>(qf, sQ, warp * 16, lane);

    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
    gemm_a_regs_bt<HP>(s, qf, sK + buf * FA_BN * LD, lane);

    // mask keys beyond N, running max / sum per row (rows r = lane/4 and r+8)
    const int kbase = t * FA_BN + 2 * (lane & 3);
    float mx[2] = {m[0], m[1]};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (kbase + nt * 8 + c >= N) { s[nt][c] = -INFINITY; s[nt][2 + c] = -INFINITY; }
      }
      mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float corr[2], rsum[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) corr[r] = exp2f((m[r] - mx[r]) * scale_log2);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = exp2f((s[nt][0] - mx[0]) * scale_log2);
      s[nt][1] = exp2f((s[nt][1] - mx[0]) * scale_log2);
      s[nt][2] = exp2f((s[nt][2] - mx[1]) * scale_log2);
      s[nt][3] = exp2f((s[nt][3] - mx[1]) * scale_log2);
      rsum[0] += s[nt][0] + s[nt][1];
      rsum[1] += s[nt][2] + s[nt][3];
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) { l[r] = l[r] * corr[r] + rsum[r]; m[r] = mx[r]; }
#pragma unroll
    for (int i = 0; i < HP / 8; ++i) { o[i][0] *= corr[0]; o[i][1] *= corr[0]; o[i][2] *= corr[1]; o[i][3] *= corr[1]; }
    uint32_t pf[4][4];
    pack_p(pf, s);
    gemm_p_bn<HP>(o, pf, sV + buf * FA_BN * LD, lane);
    __syncthreads();
  }

  // finish: full row sums across the quad, normalise, write O and LSE
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
    l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
  }
  const int row_a = q0 + warp * 16 + (lane >> 2), row_b = row_a + 8;
  const float inv_a = 1.f / l[0], inv_b = 1.f / l[1];
  const int64_t os = (int64_t)H * hd;
#pragma unroll
  for (int nt = 0; nt < HP / 8; ++nt) {
    const int d = nt * 8 + 2 * (lane & 3);
    if (d < hd) {
      if (row_a < N) *reinterpret_cast<uint32_t*>(out + ((int64_t)b * N + row_a) * os + (int64_t)h * hd + d) = pack_bf16x2(o[nt][0] * inv_a, o[nt][1] * inv_a);
      if (row_b < N) *reinterpret_cast<uint32_t*>(out + ((int64_t)b * N + row_b) * os + (int64_t)h * hd + d) = pack_bf16x2(o[nt][2] * inv_b, o[nt][3] * inv_b);
    }
  }
  if ((lane & 3) == 0) {
    const float ln2 = 0.69314718055994530942f;
    if (row_a < N) lse[((int64_t)b * H + h) * N + row_a] = (m[0] * scale_log2 + log2f(l[0])) * ln2;
    if (row_b < N) lse[((int64_t)b * H + h) * N + row_b] = (m[1] * scale_log2 + log2f(l[1])) * ln2;
  }
}

// ============================================================================ backward
// delta[b,h,i] = sum_d dO . O  -- one THREAD per (b, i, h): a head slice is hd contiguous bf16 (48..256 B),
// read as 16-byte vectors; consecutive threads take consecutive heads, i.e. consecutive memory.
// Legacy form (lse2_out == NULL): delta[(b*H+h)*N + i], unscaled (mma.sync / SIMT backward).
// tcgen05 form: rows padded to n_pad (multiple of 64, zero filled) so the backward kernels can read whole
// float4 groups without guards; delta is pre-multiplied by `scale` and lse by log2(e).
__global__ void __launch_bounds__(256)
fa_delta_kernel(const bf16* __restrict__ out, const bf16* __restrict__ dout, float* __restrict__ delta,
                const float* __restrict__ lse, float* __restrict__ lse2_out, float scale, int n_pad,
                int B, int N, int H, int hd) {
  pdl_trigger();
  pdl_wait();
  const int64_t total = (int64_t)B * n_pad * H;
  for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < total; w += (int64_t)gridDim.x * blockDim.x) {
    const int h = (int)(w % H);
    const int64_t bn = w / H;
    const int i = (int)(bn % n_pad), b = (int)(bn / n_pad);
    const int64_t oi = ((int64_t)b * H + h) * n_pad + i;
    if (i >= N) {
      delta[oi] = 0.f;
      if (lse2_out) lse2_out[oi] = 0.f;
      continue;
    }
    const int64_t row = ((int64_t)b * N + i) * H + h;
    const bf16* o = out + row * hd;
    const bf16* g = dout + row * hd;
    float s = 0.f;
    if ((hd & 7) == 0) {
      for (int d = 0; d < hd; d += 8) {
        float a[8], c[8];
        load8<bf16>(o + d, a);
        load8<bf16>(g + d, c);
#pragma unroll
        for (int e = 0; e < 8; ++e) s = fmaf(a[e], c[e], s);
      }
    } else {
      for (int d = 0; d < hd; ++d) s = fmaf(__bfloat162float(o[d]), __bfloat162float(g[d]), s);
    }
    delta[oi] = s * scale;
    if (lse2_out) lse2_out[oi] = lse[((int64_t)b * H + h) * N + i] * 1.4426950408889634f;
  }
}

// dQ pass: CTA owns 64 queries, loops over key tiles.
//   S = Q K^T ; P = exp(S*scale - lse) ; dP = dO V^T ; dS = P (dP - delta) scale ; dQ += dS K
template <int HP>
__global__ void __launch_bounds__(FA_THREADS)
fa_bwd_dq_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout, const float* __restrict__ lse,
                 const float* __restrict__ delta, bf16* __restrict__ dqkv, int N, int H, int hd, float scale, float scale_log2) {
  constexpr int LD = HP + 8;
  extern __shared__ __align__(16) uint8_t fa_smem_raw[];
  bf16* sQ = reinterpret_cast<bf16*>(fa_smem_raw);
  bf16* sG = sQ + FA_BM * LD;            // dO rows
  bf16* sK = sG + FA_BM * LD;            // [2][FA_BN][LD]
  bf16* sV = sK + 2 * FA_BN * LD;        // [2][FA_BN][LD]
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * FA_BM;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t rs = 3 * (int64_t)H * hd, os = (int64_t)H * hd;
  const bf16* qb = qkv + (int64_t)b * N * rs + (int64_t)h * hd;
  const bf16* kb = qb + os;
  const bf16* vb = kb + os;
  const bf16* gb = dout + (int64_t)b * N * os + (int64_t)h * hd;

  stage_tile<HP>(sQ, qb, rs, q0, N, hd);
  stage_tile<HP>(sG, gb, os, q0, N, hd);
  stage_tile<HP>(sK, kb, rs, 0, N, hd);
  stage_tile<HP>(sV, vb, rs, 0, N, hd);
  cp_async_commit();

  const int row_a = q0 + warp * 16 + (lane >> 2), row_b = row_a + 8;
  const float ln2inv = 1.4426950408889634f;
  float L[2], Dl[2];
  L[0] = (row_a < N) ? lse[((int64_t)b * H + h) * N + row_a] * ln2inv : 0.f;
  L[1] = (row_b < N) ? lse[((int64_t)b * H + h) * N + row_b] * ln2inv : 0.f;
  Dl[0] = (row_a < N) ? delta[((int64_t)b * H + h) * N + row_a] : 0.f;
  Dl[1] = (row_b < N) ? delta[((int64_t)b * H + h) * N + row_b] : 0.f;

  float dq[HP / 8][4];
#pragma unroll
  for (int i = 0; i < HP / 8; ++i) { dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f; }
  uint32_t qf[HP / 16][4], gf[HP / 16][4];

  const int n_tiles = (N + FA_BN - 1) / FA_BN;
  for (int t = 0; t < n_tiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < n_tiles) {
      stage_tile<HP>(sK + (buf ^ 1) * FA_BN * LD, kb, rs, (t + 1) * FA_BN, N, hd);
      stage_tile<HP>(sV + (buf ^ 1) * FA_BN * LD, vb, rs, (t + 1) * FA_BN, N, hd);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (t == 0) { load_a_frags<HP>(qf, sQ, warp * 16, lane); load_a_frags<HP>(gf, sG, warp * 16, lane); }

    float s[8][4], dp[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f; }
    gemm_a_regs_bt<HP>(s, qf, sK + buf * FA_BN * LD, lane);
    gemm_a_regs_bt<HP>(dp, gf, sV + buf * FA_BN * LD, lane);
    const int kbase = t * FA_BN + 2 * (lane & 3);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const bool valid = (kbase + nt * 8 + c) < N;
        const float pa = valid ? exp2f(s[nt][c] * scale_log2 - L[0]) : 0.f;
        const float pb = valid ? exp2f(s[nt][2 + c] * scale_log2 - L[1]) : 0.f;
        s[nt][c] = pa * (dp[nt][c] - Dl[0]) * scale;
        s[nt][2 + c] = pb * (dp[nt][2 + c] - Dl[1]) * scale;
      }
    }
    uint32_t pf[4][4];
    pack_p(pf, s);
    gemm_p_bn<HP>(dq, pf, sK + buf * FA_BN * LD, lane);
    __syncthreads();
  }
#pragma unroll
  for (int nt = 0; nt < HP / 8; ++nt) {
    const int d = nt * 8 + 2 * (lane & 3);
    if (d < hd) {
      if (row_a < N) *reinterpret_cast<uint32_t*>(dqkv + ((int64_t)b * N + row_a) * rs + (int64_t)h * hd + d) = pack_bf16x2(dq[nt][0], dq[nt][1]);
      if (row_b < N) *reinterpret_cast<uint32_t*>(dqkv + ((int64_t)b * N + row_b) * rs + (int64_t)h * hd + d) = pack_bf16x2(dq[nt][2], dq[nt][3]);
    }
  }
}

// dK/dV pass: CTA owns 64 keys (each warp 16), loops over query tiles.
//   S^T = K Q^T ; P^T = exp(S^T*scale - lse[q]) ; dV += P^T dO ; dP^T = V dO^T ;
//   dS^T = P^T (dP^T - delta[q]) scale ; dK += dS^T Q
template <int HP>
__global__ void __launch_bounds__(FA_THREADS)
fa_bwd_dkv_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout, const float* __restrict__ lse,
                  const float* __restrict__ delta, bf16* __restrict__ dqkv, int N, int H, int hd, float scale, float scale_log2) {
  constexpr int LD = HP + 8;
  extern __shared__ __align__(16) uint8_t fa_smem_raw[];
  bf16* sK = reinterpret_cast<bf16*>(fa_smem_raw);
  bf16* sV = sK + FA_BM * LD;
  bf16* sQ = sV + FA_BM * LD;            // [2][FA_BN][LD]
  bf16* sG = sQ + 2 * FA_BN * LD;        // [2][FA_BN][LD]
  float* sL = reinterpret_cast<float*>(sG + 2 * FA_BN * LD);   // [2][FA_BN] lse (log2 domain)
  float* sD = sL + 2 * FA_BN;                                  // [2][FA_BN] delta
  const int b = blockIdx.z, h = blockIdx.y, k0 = blockIdx.x * FA_BM;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t rs = 3 * (int64_t)H * hd, os = (int64_t)H * hd;
  const bf16* qb = qkv + (int64_t)b * N * rs + (int64_t)h * hd;
  const bf16* kb = qb + os;
  const bf16* vb = kb + os;
  const bf16* gb = dout + (int64_t)b * N * os + (int64_t)h * hd;
  const float* lb = lse + ((int64_t)b * H + h) * N;
  const float* db = delta + ((int64_t)b * H + h) * N;
  const float ln2inv = 1.4426950408889634f;

  auto stage_stats = [&](int buf, int q0) {
    if (threadIdx.x < FA_BN) {
      const int qi = q0 + threadIdx.x;
      sL[buf * FA_BN + threadIdx.x] = (qi < N) ? lb[qi] * ln2inv : 0.f;
      sD[buf * FA_BN + threadIdx.x] = (qi < N) ? db[qi] : 0.f;
    }
  };

  stage_tile<HP>(sK, kb, rs, k0, N, hd);
  stage_tile<HP>(sV, vb, rs, k0, N, hd);
  stage_tile<HP>(sQ, qb, rs, 0, N, hd);
  stage_tile<HP>(sG, gb, os, 0, N, hd);
  stage_stats(0, 0);
  cp_async_commit();

  float dk[HP / 8][4], dv[HP / 8][4];
#pragma unroll
  for (int i = 0; i < HP / 8; ++i) { dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f; dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f; }
  uint32_t kf[HP / 16][4], vf[HP / 16][4];

  const int n_tiles = (N + FA_BN - 1) / FA_BN;
  for (int t = 0; t < n_tiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < n_tiles) {
      stage_tile<HP>(sQ + (buf ^ 1) * FA_BN * LD, qb, rs, (t + 1) * FA_BN, N, hd);
      stage_tile<HP>(sG + (buf ^ 1) * FA_BN * LD, gb, os, (t + 1) * FA_BN, N, hd);
      stage_stats(buf ^ 1, (t + 1) * FA_BN);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (t == 0) { load_a_frags<HP>(kf, sK, warp * 16, lane); load_a_frags<HP>(vf, sV, warp * 16, lane); }

    float s[8][4], dp[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f; }
    gemm_a_regs_bt<HP>(s, kf, sQ + buf * FA_BN * LD, lane);      // S^T tile: rows = my keys, cols = queries
    gemm_a_regs_bt<HP>(dp, vf, sG + buf * FA_BN * LD, lane);     // dP^T
    const int qbase = 2 * (lane & 3);
    float pt[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int ql = nt * 8 + qbase + c;
        const bool valid = (t * FA_BN + ql) < N;
        const float Lq = sL[buf * FA_BN + ql], Dq = sD[buf * FA_BN + ql];
        const float pa = valid ? exp2f(s[nt][c] * scale_log2 - Lq) : 0.f;
        const float pb = valid ? exp2f(s[nt][2 + c] * scale_log2 - Lq) : 0.f;
        pt[nt][c] = pa; pt[nt][2 + c] = pb;
        s[nt][c] = pa * (dp[nt][c] - Dq) * scale;
        s[nt][2 + c] = pb * (dp[nt][2 + c] - Dq) * scale;
      }
    }
    uint32_t pf[4][4];
    pack_p(pf, pt);
    gemm_p_bn<HP>(dv, pf, sG + buf * FA_BN * LD, lane);          // dV += P^T dO
    pack_p(pf, s);
    gemm_p_bn<HP>(dk, pf, sQ + buf * FA_BN * LD, lane);          // dK += dS^T Q
    __syncthreads();
  }
  const int row_a = k0 + warp * 16 + (lane >> 2), row_b = row_a + 8;
#pragma unroll
  for (int nt = 0; nt < HP / 8; ++nt) {
    const int d = nt * 8 + 2 * (lane & 3);
    if (d < hd) {
      if (row_a < N) {
        bf16* base = dqkv + ((int64_t)b * N + row_a) * rs + (int64_t)h * hd + d;
        *reinterpret_cast<uint32_t*>(base + os) = pack_bf16x2(dk[nt][0], dk[nt][1]);
        *reinterpret_cast<uint32_t*>(base + 2 * os) = pack_bf16x2(dv[nt][0], dv[nt][1]);
      }
      if (row_b < N) {
        bf16* base = dqkv + ((int64_t)b * N + row_b) * rs + (int64_t)h * hd + d;
        *reinterpret_cast<uint32_t*>(base + os) = pack_bf16x2(dk[nt][2], dk[nt][3]);
        *reinterpret_cast<uint32_t*>(base + 2 * os) = pack_bf16x2(dv[nt][2], dv[nt][3]);
      }
    }
  }
}

// ============================================================================ host
static int padded_hd(int hd) {
  if (hd % 8) return 0;
  if (hd <= 32) return 32;
  if (hd <= 64) return 64;
  if (hd <= 96) return 96;
  if (hd <= 128) return 128;
  return 0;
}

bool avj_attention_mma_supported(int dtype, int hd) { return dtype == AVJ_BF16 && padded_hd(hd) != 0; }

template <int HP>
static int fwd_launch(const bf16* qkv, bf16* out, float* lse, int B, int N, int H, int hd, float scale, cudaStream_t s) {
  const size_t smem = (size_t)(FA_BM + 4 * FA_BN) * (HP + 8) * sizeof(bf16);
  static bool set = false;
  if (!set) { cudaFuncSetAttribute(fa_fwd_kernel<HP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); set = true; }
  dim3 grid((N + FA_BM - 1) / FA_BM, H, B);
  fa_fwd_kernel<HP><<<grid, FA_THREADS, smem, s>>>(qkv, out, lse, N, H, hd, scale * 1.4426950408889634f);
  AVJ_LAUNCH_CHECK();
  return 0;
}

// Row statistics of the backward: delta = rowsum(dO . O) and (tcgen05 path) lse * log2(e).
//   lse2_out == NULL : delta [B, H, N], unscaled
//   lse2_out != NULL : delta * scale and lse2 as [B, H, n_pad] with zero padding (n_pad % 64 == 0)
int avj_attention_delta(const void* out, const void* dout, float* delta, const float* lse, float* lse2_out, float scale,
                        int n_pad, int B, int N, int H, int hd, cudaStream_t s) {
  const int64_t threads = (int64_t)B * n_pad * H;
  int grid = (int)((threads + 255) / 256);
  const int cap = avj_num_sms() * 16;
  if (grid > cap) grid = cap;
  avj_launch_pdl(fa_delta_kernel, dim3(grid), dim3(256), 0, s, (const bf16*)out, (const bf16*)dout, delta, lse, lse2_out, scale, n_pad, B, N, H, hd);
  AVJ_LAUNCH_CHECK();
  return 0;
}

template <int HP>
static int bwd_launch(const bf16* qkv, const bf16* out, const bf16* dout, const float* lse, bf16* dqkv, float* ws,
                      int B, int N, int H, int hd, float scale, cudaStream_t s) {
  {
    int rc = avj_attention_delta(out, dout, ws, nullptr, nullptr, 1.0f, N, B, N, H, hd, s);
    if (rc) return rc;
  }
  const size_t smem_dq = (size_t)(2 * FA_BM + 4 * FA_BN) * (HP + 8) * sizeof(bf16);
  const size_t smem_dkv = smem_dq + 4 * FA_BN * sizeof(float);
  static bool set = false;
  if (!set) {
    cudaFuncSetAttribute(fa_bwd_dq_kernel<HP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_dq);
    cudaFuncSetAttribute(fa_bwd_dkv_kernel<HP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_dkv);
    set = true;
  }
  dim3 grid((N + FA_BM - 1) / FA_BM, H, B);
  const float sl2 = scale * 1.4426950408889634f;
  fa_bwd_dq_kernel<HP><<<grid, FA_THREADS, smem_dq, s>>>(qkv, dout, lse, ws, dqkv, N, H, hd, scale, sl2);
  AVJ_LAUNCH_CHECK();
  fa_bwd_dkv_kernel<HP><<<grid, FA_THREADS, smem_dkv, s>>>(qkv, dout, lse, ws, dqkv, N, H, hd, scale, sl2);
  AVJ_LAUNCH_CHECK();
  return 0;
}

int avj_attention_fwd_mma(const void* qkv, void* out, float* lse, int B, int N, int H, int hd, float scale, cudaStream_t s) {
  switch (padded_hd(hd)) {
    case 32: return fwd_launch<32>((const bf16*)qkv, (bf16*)out, lse, B, N, H, hd, scale, s);
    case 64: return fwd_launch<64>((const bf16*)qkv, (bf16*)out, lse, B, N, H, hd, scale, s);
    case 96: return fwd_launch<96>((const bf16*)qkv, (bf16*)out, lse, B, N, H, hd, scale, s);
    case 128: return fwd_launch<128>((const bf16*)qkv, (bf16*)out, lse, B, N, H, hd, scale, s);
  }
  avj_set_error("attention: unsupported head_dim %d", hd);
  return 1;
}

int avj_attention_bwd_mma(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                          float* ws, int B, int N, int H, int hd, float scale, cudaStream_t s) {
  switch (padded_hd(hd)) {
    case 32: return bwd_launch<32>((const bf16*)qkv, (const bf16*)out, (const bf16*)dout, lse, (bf16*)dqkv, ws, B, N, H, hd, scale, s);
    case 64: return bwd_launch<64>((const bf16*)qkv, (const bf16*)out, (const bf16*)dout, lse, (bf16*)dqkv, ws, B, N, H, hd, scale, s);
    case 96: return bwd_launch<96>((const bf16*)qkv, (const bf16*)out, (const bf16*)dout, lse, (bf16*)dqkv, ws, B, N, H, hd, scale, s);
    case 128: return bwd_launch<128>((const bf16*)qkv, (const bf16*)out, (const bf16*)dout, lse, (bf16*)dqkv, ws, B, N, H, hd, scale, s);
  }
  avj_set_error("attention: unsupported head_dim %d", hd);
  return 1;
}
