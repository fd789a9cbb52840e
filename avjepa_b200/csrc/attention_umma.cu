// K7 flash-attention FORWARD on the 5th-gen tensor cores (sm_100a): S = Q K^T and O += P V are
// tcgen05.mma instructions with their accumulators in TMEM; softmax runs on 128 threads that each own
// one query row (= one TMEM lane).
//
//   warp 0      : loader -- TMA box loads (hd == 64) or cp.async 16-byte copies of Q / K / V head slices
//                 straight out of the qkv-Linear output [B, N, 3, H, hd] into SWIZZLE_128B shared-memory
//                 tiles (zero padding of hd -> HDP and of rows >= N happens here, in shared memory only;
//                 warps 2-3 help on the cp.async path)
//   warp 1      : TMEM allocator + single-thread MMA issuer
//   warps 4..7  : softmax (208 registers after setmaxnreg; warps 0-3 drop to 48 so that two CTAs fit the
//                 register file of every SM sub-partition): tcgen05.ld S (128 fp32 columns per row), running max / sum in fp32 with
//                 exp2f, P -> bf16 -> swizzled smem (A operand of the PV MMA), lazy rescale of the
//                 TMEM-resident O only when a row maximum moves by more than 2^8, final O / l
//
// CTA = 128 queries of one (batch, head); K/V tiles of 128 keys, double-buffered; S (128 cols) and
// O (HDP cols) share a 256-column TMEM allocation so two CTAs fit per SM and overlap each other's
// MMA and softmax phases.  S[t+1] is issued as soon as the softmax has pulled S[t] into registers.
//
// Roofline: MUFU-bound for hd <= 64 (one exp2 per score: 128x128 per tile per CTA) -- the tensor
// pipe needs ~512 cycles per tile at hd = 64, the 16 exp2/clk/SM special-function units ~1024.
// Algorithmic FLOPs: 4 * N^2 * hd per (batch, head).
#include "common.cuh"

#include "umma_attn.cuh"
#include <mutex>
#include <type_traits>
#include <stdlib.h>
#include <utility>
#include <vector>

#define UA_BM 128
#define UA_BN 128
#define UA_THREADS 256
#define UA_LOADERS 96             // cp.async path: warps 0, 2, 3
#define UA_P_TILE (128 * 128)                     // one 64-key atom of P: 128 rows x 128 B

// K/V ring depth: the stage of tile t is released by PV(t), so with 2 stages the load of tile t+2 is issued when
// PV(t) retires and S(t+2) -- wanted one softmax later -- waits out its full L2 latency every tile.  hd <= 32 tiles
// have 64-byte rows (UaTile) and afford 4 stages next to a second CTA; at hd = 64 there is room for 3 once P lives
// in tensor memory (TS) and its 32 KB of shared memory are gone.
template <int HDP, bool TS> struct UaSmem {
  static constexpr int NST = HDP == 32 ? 4 : (HDP == 128 ? 2 : (TS ? 3 : 2));
  static constexpr uint32_t HALF = 128 * UaTile<HDP>::PITCH;            // one 64-column (or 32-column) sub-tile
  static constexpr uint32_t TILE = HALF * UaTile<HDP>::NH;
  static constexpr uint32_t Q = 0, K = TILE, V = K + NST * TILE, P = V + NST * TILE, BARS = P + (TS ? 0 : 2 * UA_P_TILE),
                            TOTAL = BARS + 256;
};
#define UA_O_COL 128
// TMEM: S 128 columns | O HDP columns | (TS) 64 columns of packed bf16 P.  256 columns (two CTAs per SM) up to HDP = 64;
// HDP = 128 needs 320 -> a 512-column allocation, one CTA per SM (its 160 KB of shared memory allow no second one anyway).
template <int HDP> struct UaTmem {
  static constexpr uint32_t COLS = HDP == 128 ? 512 : 256;
  static constexpr uint32_t P_COL = HDP == 128 ? 256 : 192;
};

// TMA = true (hd == 64): one elected thread issues 16 KB box loads (rows past N are zero-filled by
// the hardware); TMA = false: the loader warp gathers 16-byte chunks with cp.async and pads in smem.
// POLY: every fourth pair of scores takes ua_exp2_poly (FMA pipe) instead of MUFU ex2.
// TS: P goes to tensor memory (packed bf16, columns [192, 256)) and is the A operand of the PV MMA from there.
// ILV: exp2 / row-sum / pack of the softmax interleaved group-wise through a value-neutral dependency (see below).
template <int HDP, bool TMA, bool POLY, bool TS, bool ILV>
__global__ void __launch_bounds__(UA_THREADS, 2)
fa_fwd_umma_kernel(const __grid_constant__ CUtensorMap tmap, const bf16* __restrict__ qkv, bf16* __restrict__ out,
                   float* __restrict__ lse, int N, int H, int hd, float scale_log2, uint32_t zero_u) {
  extern __shared__ __align__(1024) uint8_t ua_raw[];
  const uint32_t base = ua_smem(ua_raw);
  using L = UaSmem<HDP, TS>;
  using TL = UaTile<HDP>;
  constexpr int NST = L::NST;
  constexpr uint32_t UA_TILE_BYTES = L::TILE;
  constexpr int NH = TL::NH;
  constexpr uint32_t UA_TMEM_COLS = UaTmem<HDP>::COLS, UA_P_COL = UaTmem<HDP>::P_COL;
  static_assert(HDP != 128 || (TMA && TS), "head_dim > 64 runs the TMA loaders with P in tensor memory only");
  const uint32_t sQ = base + L::Q;
  const uint32_t sK = base + L::K;                   // [NST]
  const uint32_t sV = base + L::V;                   // [NST]
  const uint32_t sP = base + L::P;                   // 2 atoms of 64 keys
  const uint32_t bars = base + L::BARS;
  const uint32_t kv_full = bars;            // [NST <= 4]
  const uint32_t kv_empty = bars + 32;      // [NST <= 4] count 1
  const uint32_t s_full = bars + 64;        // count 1
  const uint32_t s_free = bars + 72;        // count 128
  const uint32_t p_full = bars + 80;        // count 128
  const uint32_t o_done = bars + 88;        // count 1
  const uint32_t tmem_slot = bars + 96;
  const uint32_t q_ready = bars + 104;      // count 32: Q pad columns zeroed (TMA path, hd < HDP)
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(ua_raw + (tmem_slot - base));

  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * UA_BM;
  // A TMA box always spans HDP columns: for hd < HDP columns [hd, HDP) of every tile hold the NEXT head's
  // values.  They only matter in the contraction of S = Q K^T, so zeroing them in the Q tile is enough; the
  // matching columns of O = P V are computed from V's neighbour and never stored.
  // (HDP = 128: head_dim is a multiple of 16 and the score MMAs simply stop after hd / 16 K-steps.)
  const bool zero_q_pad = TMA && hd < HDP && HDP != 128;
  const int64_t rs = 3 * (int64_t)H * hd;
  const bf16* qb = qkv + (int64_t)b * N * rs + (int64_t)h * hd;
  const bf16* kb = qb + (int64_t)H * hd;
  const bf16* vb = kb + (int64_t)H * hd;
  const int T = (N + UA_BN - 1) / UA_BN;

  if (threadIdx.x == 0) {
    if (base & 1023u) __trap();                      // SWIZZLE_128B tiles need 1024-byte alignment
    for (int i = 0; i < NST; ++i) { ua_mbar_init(kv_full + 8 * i, TMA ? 1 : UA_LOADERS); ua_mbar_init(kv_empty + 8 * i, 1); }
    ua_mbar_init(s_full, 1); ua_mbar_init(s_free, 128); ua_mbar_init(p_full, 128); ua_mbar_init(o_done, 1);
    ua_mbar_init(q_ready, 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)UA_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  ua_fence_before();
  __syncthreads();
  ua_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  pdl_wait();

  if (warp < 4) {
  ua_reg_dec<48>();
  if (warp == 0 || (!TMA && warp >= 2)) {
    // ============================ loader ============================
    const int ld_tid = warp == 0 ? lane : (warp - 1) * 32 + lane;     // 0..95 on the cp.async path
    if (TMA) {
      if (lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
        const int cq = h * hd, ck = (H + h) * hd, cv = (2 * H + h) * hd;
        for (int t = 0; t < T; ++t) {
          const int st = t % NST;
          if (t >= NST) ua_mbar_wait(kv_empty + 8 * st, ((t / NST) & 1) ^ 1);
          const uint32_t fb = kv_full + 8 * st;
          ua_expect_tx(fb, (t == 0 ? 3 : 2) * UA_TILE_BYTES);
#pragma unroll
          for (int hf = 0; hf < NH; ++hf) {                          // one box per 64-column half
            if (t == 0) ua_tma3d(sQ + hf * L::HALF, &tmap, fb, cq + 64 * hf, q0, b);
            ua_tma3d(sK + st * UA_TILE_BYTES + hf * L::HALF, &tmap, fb, ck + 64 * hf, t * UA_BN, b);
            ua_tma3d(sV + st * UA_TILE_BYTES + hf * L::HALF, &tmap, fb, cv + 64 * hf, t * UA_BN, b);
          }
        }
      }
    } else {
      // up to NST-1 K/V tiles are kept in flight.  kv_full[j] is signalled BEFORE waiting for the stage of
      // tile t to drain (the MMA warp issues S(j) ahead of PV(t-NST)); the opposite order would deadlock.
      ua_stage<HDP>(sQ, qb, rs, q0, N, hd, ld_tid, UA_LOADERS);
      for (int t = 0; t < T; ++t) {
        const int st = t % NST;
        if (t >= NST - 1) {
          asm volatile("cp.async.wait_group %0;" ::"n"(NST - 2) : "memory");
          ua_fence_async_smem();
          ua_mbar_arrive(kv_full + 8 * ((t - (NST - 1)) % NST));
        }
        if (t >= NST) ua_mbar_wait(kv_empty + 8 * st, ((t / NST) & 1) ^ 1);
        ua_stage<HDP>(sK + st * UA_TILE_BYTES, kb, rs, t * UA_BN, N, hd, ld_tid, UA_LOADERS);
        ua_stage<HDP>(sV + st * UA_TILE_BYTES, vb, rs, t * UA_BN, N, hd, ld_tid, UA_LOADERS);
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      ua_fence_async_smem();
      for (int j = max(0, T - (NST - 1)); j < T; ++j) ua_mbar_arrive(kv_full + 8 * (j % NST));
    }
  } else if (TMA && warp == 2) {
    if (zero_q_pad) {
      ua_mbar_wait(kv_full, 0);                                       // Q (and tile 0) have landed
      ua_zero_pad<HDP>(sQ, 128, hd, lane);
      ua_fence_async_smem();
      ua_mbar_arrive(q_ready);
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    if (lane == 0) {
      const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(UA_BN >> 3) << 17) | ((uint32_t)(UA_BM >> 4) << 24);
      const uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)((HDP == 128 ? 64 : HDP) >> 3) << 17) |
                               ((uint32_t)(UA_BM >> 4) << 24);
      // HDP = 128: the second half of O has hd - 64 columns (16 .. 64)
      const uint32_t idesc_o2 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)((hd - 64) >> 3) << 17) | ((uint32_t)(UA_BM >> 4) << 24);
      const uint64_t qd = TL::kmajor(sQ);
      const int ksteps = HDP == 128 ? hd / 16 : HDP / 16;
      auto issue_s = [&](int t) {
        const uint64_t kd = TL::kmajor(sK + (t % NST) * UA_TILE_BYTES);
        for (int k = 0; k < ksteps; ++k) {
          const uint32_t ho = (uint32_t)(k >> 2) * (L::HALF >> 4);     // descriptor offset of the half (16-byte units)
          ua_mma(tmem, qd + ho + 2 * (k & 3), kd + ho + 2 * (k & 3), idesc_s, k > 0 ? 1u : 0u);
        }
        ua_commit(s_full);
      };
      ua_mbar_wait(kv_full, 0);   // tile 0 (and Q)
      if (zero_q_pad) ua_mbar_wait(q_ready, 0);
      ua_fence_after();
      issue_s(0);
      for (int t = 0; t < T; ++t) {
        if (t + 1 < T) {
          ua_mbar_wait(kv_full + 8 * ((t + 1) % NST), ((t + 1) / NST) & 1);
          ua_mbar_wait(s_free, t & 1);                 // softmax has S[t] in registers
          ua_fence_after();
          issue_s(t + 1);
        }
        ua_mbar_wait(p_full, t & 1);
        ua_fence_after();
        const uint32_t vt = sV + (t % NST) * UA_TILE_BYTES;
#pragma unroll
        for (int k = 0; k < UA_BN / 16; ++k) {
          const uint64_t pd = ua_desc(sP + (k >> 2) * UA_P_TILE, 1, 64) + 2 * (k & 3);
          const uint64_t vd = TL::mnmajor(vt) + TL::MN_KADV * k;     // MN-major: 16 keys per step
          if (TS) ua_mma_ts(tmem + UA_O_COL, tmem + UA_P_COL + 8 * k, vd, idesc_o, (t > 0 || k > 0) ? 1u : 0u);
          else    ua_mma(tmem + UA_O_COL, pd, vd, idesc_o, (t > 0 || k > 0) ? 1u : 0u);
          if constexpr (HDP == 128)          // columns [64, hd) of O from the second V half
            ua_mma_ts(tmem + UA_O_COL + 64, tmem + UA_P_COL + 8 * k, vd + (L::HALF >> 4), idesc_o2, (t > 0 || k > 0) ? 1u : 0u);
        }
        ua_commit(o_done);
        ua_commit(kv_empty + 8 * (t % NST));
      }
    }
  }
  } else {
    ua_reg_inc<208>();
    // ============================ softmax ============================
    const int q = warp & 3;
    const int row = q * 32 + lane;                                   // TMEM lane == query row in tile
    const uint32_t t_s = tmem + ((uint32_t)(q * 32) << 16);
    const uint32_t t_o = t_s + UA_O_COL;
    float m = -INFINITY, l = 0.f;
    // zero_u: kernel argument, always 0 -- a zero the compiler cannot see through (ILV dependency chain)
    // MASKED is a compile-time tag: only the LAST key tile of a sequence with N % 128 != 0 carries the
    // 128 compare+select pairs that push keys past N to -inf (as a run-time predicate the compiler
    // if-converts them into every tile: +2 instructions per score in an issue/MUFU-bound loop)
    auto tile = [&](const int t, auto masked_tag) {
      constexpr bool MASKED = decltype(masked_tag)::value;
      ua_mbar_wait(s_full, t & 1);
      ua_fence_after();
      float s[128];
      ua_ld32(t_s, s); ua_ld32(t_s + 32, s + 32); ua_ld32(t_s + 64, s + 64); ua_ld32(t_s + 96, s + 96);
      ua_ld_wait();
      ua_fence_before();
      ua_mbar_arrive(s_free);
      if constexpr (MASKED) {
        const int nvalid = N - t * UA_BN;
#pragma unroll
        for (int j = 0; j < 128; ++j) if (j >= nvalid) s[j] = -INFINITY;
      }
      // four independent chains (a single running maximum is 64 dependent FMNMX3 ~ 300 idle cycles per tile)
      float mx0 = fmaxf(s[0], s[1]), mx1 = fmaxf(s[2], s[3]), mx2 = fmaxf(s[4], s[5]), mx3 = fmaxf(s[6], s[7]);
#pragma unroll
      for (int j = 8; j < 128; j += 8) {
        mx0 = fmaxf(mx0, fmaxf(s[j], s[j + 1]));
        mx1 = fmaxf(mx1, fmaxf(s[j + 2], s[j + 3]));
        mx2 = fmaxf(mx2, fmaxf(s[j + 4], s[j + 5]));
        mx3 = fmaxf(mx3, fmaxf(s[j + 6], s[j + 7]));
      }
      const float tmax = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      const float m_new = fmaxf(m, tmax);
      // lazy rescale: keep the stale maximum unless it is off by more than 2^8
      const bool jump = (m_new - m) * scale_log2 > 8.0f;             // true for t == 0 (m = -inf)
      const float m_use = jump ? m_new : m;
      const float corr = exp2f((m - m_use) * scale_log2);            // 1 when unchanged, 0 at t == 0
      const float mb = m_use * scale_log2;
      float rs0 = 0.f, rs1 = 0.f, rs2 = 0.f, rs3 = 0.f;
      uint32_t pk[64];
      if constexpr (ILV) {
        // ptxas emits the straightforward loop as three PHASES (128 FFMA, 128 MUFU.EX2, 128 FADD + 64 F2FP): a warp's
        // MUFU instructions issue 8 cycles apart, so for ~1000 cycles it issues nothing else, and afterwards the MUFU pipe
        // idles while the warp adds and packs.  The two co-resident CTAs share the pipe processor-style, which keeps their
        // phases aligned -- the pipe sits idle during BOTH warps' add / pack / load phases (ncu: MUFU pipe 63-65 %).
        // Here groups of 8 scores are chained through a value-neutral dependency: the offset of group g + 2 is
        // fma(partial row sum of group g, 0, -mb), so the adds and packs of group g MUST be scheduled before the exp2 of
        // group g + 2 -- i.e. inside the issue shadow of the MUFU instructions of group g + 1.
        float nmb_a = -mb, nmb_b = -mb;
#pragma unroll
        for (int g = 0; g < 16; ++g) {
          const float nmb = (g & 1) ? nmb_b : nmb_a;
          float p[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int j = 8 * g + e;
            const float x = fmaf(s[j], scale_log2, nmb);
            p[e] = (POLY && e >= 6) ? ua_exp2_poly(x) : ua_exp2(x);
          }
          const float t = ((p[0] + p[1]) + (p[2] + p[3])) + ((p[4] + p[5]) + (p[6] + p[7]));
          if ((g & 3) == 0) rs0 += t; else if ((g & 3) == 1) rs1 += t; else if ((g & 3) == 2) rs2 += t; else rs3 += t;
#pragma unroll
          for (int e = 0; e < 8; e += 2) pk[4 * g + (e >> 1)] = pack_bf16x2(p[e], p[e + 1]);
          // == -mb (t is finite, zero_u is 0 at run time), but it depends on group g's row sum AND on its packed words
          const uint32_t zbits = ((pk[4 * g] | pk[4 * g + 1]) | (pk[4 * g + 2] | pk[4 * g + 3])) & zero_u;
          const float dep = fmaf(t, 0.0f, -mb) + __uint_as_float(zbits);
          if (g & 1) nmb_b = dep; else nmb_a = dep;
        }
      } else {
#pragma unroll
      for (int j = 0; j < 128; j += 2) {
        const float x0 = fmaf(s[j], scale_log2, -mb), x1 = fmaf(s[j + 1], scale_log2, -mb);
        const bool poly = POLY && ((j >> 1) & 3) == 3;
        const float p0 = poly ? ua_exp2_poly(x0) : ua_exp2(x0);
        const float p1 = poly ? ua_exp2_poly(x1) : ua_exp2(x1);
        if (((j >> 1) & 3) == 0) rs0 += p0 + p1;
        else if (((j >> 1) & 3) == 1) rs1 += p0 + p1;
        else if (((j >> 1) & 3) == 2) rs2 += p0 + p1;
        else rs3 += p0 + p1;
        pk[j >> 1] = pack_bf16x2(p0, p1);
      }
      }
      l = l * corr + ((rs0 + rs1) + (rs2 + rs3));
      m = m_use;
      if (t > 0) ua_mbar_wait(o_done, (t - 1) & 1);                  // PV[t-1] finished: P smem and O are ours
      ua_fence_after();
      if constexpr (TS) {
        ua_st_regs<32>(t_s + UA_P_COL, pk);
        ua_st_regs<32>(t_s + UA_P_COL + 32, pk + 32);
      } else {
#pragma unroll
        for (int c = 0; c < 16; ++c) {                               // 16 chunks of 8 keys
          const uint32_t dst = sP + (c >> 3) * UA_P_TILE + row * 128 + (((c & 7) ^ (row & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk[4 * c]), "r"(pk[4 * c + 1]),
                       "r"(pk[4 * c + 2]), "r"(pk[4 * c + 3]) : "memory");
        }
      }
      if (t > 0 && __any_sync(0xffffffffu, jump)) {
        float o[32];
#pragma unroll
        for (int c = 0; c < HDP; c += 32) {
          if (HDP == 128 && c >= hd) break;                          // columns past head_dim are never written
          ua_ld32(t_o + c, o);
          ua_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) o[j] *= corr;
          ua_st32(t_o + c, o);
        }
      }
      if constexpr (TS) ua_st_wait(); else ua_fence_async_smem();
      ua_fence_before();
      ua_mbar_arrive(p_full);
    };
    const bool tail = (N % UA_BN) != 0;
    for (int t = 0; t + 1 < T; ++t) tile(t, std::false_type{});
    if (tail) tile(T - 1, std::true_type{}); else tile(T - 1, std::false_type{});
    ua_mbar_wait(o_done, (T - 1) & 1);
    ua_fence_after();
    const int qi = q0 + row;
    const float inv = 1.f / l;
    bf16* orow = out + ((int64_t)b * N + qi) * (int64_t)H * hd + (int64_t)h * hd;
#pragma unroll
    for (int c = 0; c < HDP; c += 32) {
      if (HDP == 128 && c >= hd) break;
      float o[32];
      ua_ld32(t_o + c, o);
      ua_ld_wait();
      if (qi < N) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          if (c + j < hd) {
            float v8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v8[e] = o[j + e] * inv;
            store8<bf16>(orow + c + j, v8);
          }
        }
      }
    }
    if (qi < N) lse[((int64_t)b * H + h) * N + qi] = (m * scale_log2 + log2f(l)) * 0.69314718055994530942f;
  }

  ua_fence_before();
  __syncthreads();
  if (warp == 1) {
    ua_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)UA_TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------
bool avj_attention_umma_fwd_supported(int dtype, int hd) {
  // head_dim <= 64: multiples of 8 (padded to 32 / 64 in shared memory only); 80 .. 128: multiples of 16, two 64-column halves
  return dtype == AVJ_BF16 && ((hd % 8 == 0 && hd >= 8 && hd <= 64) || (hd % 16 == 0 && hd > 64 && hd <= 128));
}

typedef CUresult (*PFN_ua_encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// [B, N, cols] bf16 viewed as a 3-D tensor {cols, N, B}; box = {box_cols cols, box_rows rows, 1}:
// 64 columns -> 128-byte rows, SWIZZLE_128B; 32 columns -> 64-byte rows, SWIZZLE_64B (UaTile<32>).
int ua_make_map3d(const void* qkv, int B, int N, int cols, int box_rows, CUtensorMap* out, int box_cols) {
  struct Key { const void* p; int B, N, cols, box_rows, box_cols; };
  AVJ_CHECK(box_cols == 32 || box_cols == 64, "ua_make_map3d: box_cols must be 32 or 64");
  static std::mutex mu;
  static std::vector<std::pair<Key, CUtensorMap>> cache;
  {
    std::lock_guard<std::mutex> g(mu);
    for (auto& e : cache)
      if (e.first.p == qkv && e.first.B == B && e.first.N == N && e.first.cols == cols && e.first.box_rows == box_rows &&
          e.first.box_cols == box_cols) { *out = e.second; return 0; }
  }
  static PFN_ua_encode enc = nullptr;
  if (!enc) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      enc = reinterpret_cast<PFN_ua_encode>(p);
  }
  AVJ_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)N, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)cols * 2, (cuuint64_t)cols * 2 * (cuuint64_t)N};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(qkv), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  AVJ_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(qkv) failed (%d)", (int)r);
  std::lock_guard<std::mutex> g(mu);
  if (cache.size() > 512) cache.clear();
  cache.push_back({Key{qkv, B, N, cols, box_rows, box_cols}, *out});
  return 0;
}

template <int HDP, bool TMA, bool POLY, bool TS, bool ILV>
static int ua_launch(const bf16* qkv, bf16* out, float* lse, int B, int N, int H, int hd, float scale, cudaStream_t s) {
  static bool set = false;
  if (!set) {
    cudaError_t e = cudaFuncSetAttribute(fa_fwd_umma_kernel<HDP, TMA, POLY, TS, ILV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UaSmem<HDP, TS>::TOTAL);
    AVJ_CHECK(e == cudaSuccess, "cudaFuncSetAttribute(fa_fwd_umma_kernel) failed: %s", cudaGetErrorString(e));
    // two CTAs per SM need the full 228 KB shared-memory carve-out
    cudaFuncSetAttribute(fa_fwd_umma_kernel<HDP, TMA, POLY, TS, ILV>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    set = true;
  }
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  if (TMA) {
    int rc = ua_make_map3d(qkv, B, N, 3 * H * hd, 128, &map, HDP == 128 ? 64 : HDP);
    if (rc) return rc;
  }
  dim3 grid((N + UA_BM - 1) / UA_BM, H, B);
  avj_launch_pdl(fa_fwd_umma_kernel<HDP, TMA, POLY, TS, ILV>, grid, dim3(UA_THREADS), UaSmem<HDP, TS>::TOTAL, s, map, qkv, out, lse, N, H, hd, scale * 1.4426950408889634f, 0u);
  AVJ_LAUNCH_CHECK();
  return 0;
}

int avj_attention_fwd_umma(const void* qkv, void* out, float* lse, int B, int N, int H, int hd, float scale, cudaStream_t s) {
  static int use_tma = -1;
  if (use_tma < 0) { const char* e = getenv("AVJ_ATTN_TMA"); use_tma = (e && e[0] == '0') ? 0 : 1; }
  static int use_tma32 = -1;   // AVJ_ATTN_TMA32=0: head_dim <= 32 goes back to the cp.async gather loaders
  if (use_tma32 < 0) { const char* e = getenv("AVJ_ATTN_TMA32"); use_tma32 = (e && e[0] == '0') ? 0 : use_tma; }
  // A quarter of the exp2 evaluations on the FMA pipe.  With the interleaved softmax loop this wins 3 % on long sequences
  // (0.3725 -> 0.3614 ms at N = 1664, hd 64) and loses 4 % on short ones (N = 384), so the default is by length;
  // AVJ_ATTN_POLY=0 / 1 forces it off / on.
  static int poly_env = -2;
  if (poly_env == -2) { const char* e = getenv("AVJ_ATTN_POLY"); poly_env = !e ? -1 : (e[0] == '1' ? 1 : 0); }
  const int use_poly = poly_env >= 0 ? poly_env : (N >= 1024 ? 1 : 0);
  const bool al = (reinterpret_cast<uintptr_t>(qkv) & 15) == 0;
  static int ts = -1;          // AVJ_ATTN_TMEM_P=0: P through shared memory instead of tensor memory
  if (ts < 0) { const char* e = getenv("AVJ_ATTN_TMEM_P"); ts = (e && e[0] == '0') ? 0 : 1; }
#define UA_ARGS (const bf16*)qkv, (bf16*)out, lse, B, N, H, hd, scale, s
  static int ilv = -1;         // AVJ_ATTN_ILV=0: the plain (phase-ordered) softmax loop of round 1
  if (ilv < 0) { const char* e = getenv("AVJ_ATTN_ILV"); ilv = (e && e[0] == '0') ? 0 : 1; }
#define UA_GO(HDP_, TMA_)                                                                             \
  {                                                                                                   \
    if (ts && ilv) return use_poly ? ua_launch<HDP_, TMA_, true, true, true>(UA_ARGS) : ua_launch<HDP_, TMA_, false, true, true>(UA_ARGS);   \
    if (ts) return use_poly ? ua_launch<HDP_, TMA_, true, true, false>(UA_ARGS) : ua_launch<HDP_, TMA_, false, true, false>(UA_ARGS);   \
    return use_poly ? ua_launch<HDP_, TMA_, true, false, false>(UA_ARGS) : ua_launch<HDP_, TMA_, false, false, false>(UA_ARGS);         \
  }
  if (hd <= 32) {
    if (use_tma32 && al) { UA_GO(32, true); }
    UA_GO(32, false);
  }
  if (hd > 64) {
    AVJ_CHECK(al && use_tma, "attention forward: head_dim %d needs 16-byte aligned qkv (TMA loaders)", hd);
    return ilv ? ua_launch<128, true, false, true, true>(UA_ARGS) : ua_launch<128, true, false, true, false>(UA_ARGS);
  }
  if (hd == 64 && use_tma && al) { UA_GO(64, true); }
  UA_GO(64, false);
#undef UA_GO
#undef UA_ARGS
}
