// K7 flash-attention BACKWARD on the 5th-gen tensor cores (sm_100a).  ONE kernel template gives the two deterministic passes
// below (head_dim 64 .. 128: no atomics, no dQ accumulation buffer) and, for head_dim <= 32, the SINGLE-PASS form in which
// the KV pass also produces dQ (fp32 accumulator + TMA reduce-add, see SP further down):
//
//   KV pass (KV = true ): CTA owns 128 keys of one (batch, head) and streams 64-query tiles
//        S^T  = K Q^T          P^T  = exp2(S^T * scale*log2e - lse*log2e)      (TMEM lane = key)
//        dP^T = V dO^T         dS^T = P^T o (dP^T - delta) * scale
//        dV  += P^T dO         dK  += dS^T Q
//   Q pass  (KV = false): CTA owns 128 queries and streams 64-key tiles
//        S    = Q K^T          dP   = dO V^T                                   (TMEM lane = query)
//        dS   = P o (dP - delta) * scale          dQ += dS K
//
// Computing the TRANSPOSED scores in the KV pass makes P^T / dS^T row-per-thread, so they are written to
// shared memory as K-major A operands of the dV / dK MMAs with no transpose anywhere; the streamed
// Q / dO (or K / V) tiles serve both as K-major B operands of the score MMAs and as MN-major B
// operands of the output MMAs from the same shared-memory bytes.
//
//   warp 0      : loader (TMA box loads when hd == 64, else 16-byte cp.async gathers with zero padding
//                 hd -> HDP in shared memory only; warps 2-3 help gathering) + per-column lse/delta
//                 staging for the KV pass
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer
//   warps 4..7  : one thread per TMEM lane: tcgen05.ld of both 128x64 score tiles, exp2 / dS math in
//                 fp32, bf16 pack, swizzled st.shared; final dK/dV (or dQ) TMEM -> bf16 -> dqkv
//
// 256 threads compiled for 128 registers so that two CTAs are co-resident in EVERY SM sub-partition's
// register file; warps 0-3 then shrink to 56 registers (setmaxnreg.dec) and the math warpgroup grows
// to 200 (setmaxnreg.inc) -- it holds two 64-column fp32 score rows per thread.
//
// TMEM: scores 2 x 64 columns + outputs 2 x HDP columns <= 256, so two CTAs share an SM and overlap
// each other's MMA and exp2 phases.  Rows past N are zero (TMA OOB fill / explicit zero) which makes
// every masked contribution vanish without predicates; lse/delta of such rows are staged as 0.
//
// Roofline: MUFU/FMA-bound like the forward (one exp2 + ~6 FMA-pipe ops per score element, twice --
// once per pass).  Algorithmic FLOPs: 10 * N^2 * hd per (batch, head); executed: 14 * N^2 * HDP.
#include "umma_attn.cuh"

#include <mutex>
#include <stdlib.h>
#include <utility>
#include <vector>

#define UB_BM 128
#define UB_BN 64
#define UB_THREADS 256
#define UB_LOADERS 96            // cp.async path: warps 0, 2, 3
#define UB_T2_COL 64
#define UB_O1_COL 128
#define UB_O2_COL 192
#define UB_PD_TILE (128 * 128)       // P^T / dS^T: 128 rows x 64 streamed columns bf16 (always 128-byte rows)

// Streamed-tile ring depth.  The stage of tile t is only released by the OUTPUT MMAs of step t, so with 2
// stages the load of tile t+2 starts when step t ends and its full L2 latency is exposed every step; 3-4
// stages issue it one or two steps earlier.  hd <= 32 tiles are half the size (64-byte rows) so 4 stages
// fit next to a second CTA; at hd = 64 there is room for 3 (112 KB per CTA in the dK/dV pass).
template <int HDP, bool KV> struct UbSmem {
  static constexpr int NST = HDP == 32 ? 4 : (HDP == 128 ? 2 : 3);
  // column-statistics buffers of the KV pass: two (one named barrier per step) where shared memory allows;
  // at hd = 64 two co-resident CTAs leave room for one (two barriers per step)
  static constexpr int SB = HDP == 32 ? 2 : 1;
  // HDP = 128: every tile is two 64-column SWIZZLE_128B halves (RHALF / CHALF apart), 161 KB -> one CTA per SM
  static constexpr uint32_t RHALF = 128 * UaTile<HDP>::PITCH, CHALF = 64 * UaTile<HDP>::PITCH;
  static constexpr uint32_t ROW_TILE = RHALF * UaTile<HDP>::NH, COL_TILE = CHALF * UaTile<HDP>::NH;
  // single pass (HDP == 32 KV pass): TWO dS^T tiles back to back -- the dQ MMAs run once per PAIR of steps with M = 128
  // (queries of both steps), reading both tiles as the two 64-query chunks of one MN-major operand
  static constexpr uint32_t NDS = (KV && HDP == 32) ? 2 : 1;
  static constexpr uint32_t R1 = 0, R2 = ROW_TILE, C1 = 2 * ROW_TILE, C2 = C1 + NST * COL_TILE,
                            DS = C2 + NST * COL_TILE, P = DS + NDS * UB_PD_TILE,
                            STAT = KV ? P + UB_PD_TILE : P,            // KV pass: [SB][lse2 64 | delta 64] floats
                            BARS = STAT + (KV ? SB * 512 : 0), TOTAL = BARS + 192;
};

// MW = number of math warpgroups (1 or 2).  The score math has no cross-column dependency (row statistics come
// from the forward / the delta kernel), so with MW = 2 each thread owns HALF of its row's 64 streamed columns:
// twice the warps in flight per scheduler for a latency-bound exp2/FMA loop, half the per-step critical path.
// TS (HDP == 32 only: 2 x 64 score + 2 x 32 output + 2 x 32 operand columns = 256): P^T / dS^T are written to
// tensor memory as packed bf16 and consumed as the A operand of the output MMAs from there.
// PP (MW == 2 only): the two math warpgroups are DECOUPLED -- each owns its half of the streamed columns with its own
// score / operand / output barriers, the MMA thread serves the halves alternately (score MMAs of N = 32, output MMAs of
// K = 32).  In lockstep (PP = false) both warpgroups wait for the same score MMAs at the same time, so the MUFU pipe
// idles whenever the CTA waits; decoupled, one half's exp2 math runs under the other half's TMEM loads, operand stores
// and MMA round trips (ncu, lockstep: 22 % of the math warps' samples sit at the t_full wait, MUFU pipe 47-59 %).
// SP (single pass; KV pass only, HDP == 32, TMA, TS, MW == 2): the KV-owner CTA ALSO produces dQ, so the Q pass -- and its
// second evaluation of every exp2 -- disappears.  Per step dQ_tile[64 q, hd] = dS[64 q, 128 keys] K[128 keys, hd] is one more
// group of tcgen05.mma (M = 64; A = the dS^T tile in shared memory read MN-major, B = the stationary K tile read MN-major),
// its fp32 result leaves TMEM through a 64 x hd staging tile and ONE TMA reduce-add (cp.reduce.async.bulk.tensor .add)
// into an fp32 dQ accumulator [B, N, H, hd]; dS^T therefore lives in shared memory (it feeds dK K-major and dQ MN-major),
// P^T stays in tensor memory.  TMEM: S^T 64 | dP^T 64 | dV 32 | dK 32 | P^T 32 | dQ 32 = 256 columns.  The dQ MMAs and the
// reduce-add run once per PAIR of steps (M = 128: two dS^T tiles kept side by side), so a pair costs 32 instead of 40 MMAs.
template <int HDP, bool TMA, bool KV, int MW, bool TS, bool PP, bool SP>
__global__ void __launch_bounds__(128 + 128 * MW, 2)
fa_bwd_umma_kernel(const __grid_constant__ CUtensorMap map_qkv128, const __grid_constant__ CUtensorMap map_qkv64,
                   const __grid_constant__ CUtensorMap map_do, const __grid_constant__ CUtensorMap map_dq,
                   const __grid_constant__ CUtensorMap map_dq2, const bf16* __restrict__ qkv, const bf16* __restrict__ dout,
                   const float* __restrict__ lse2, const float* __restrict__ delta, bf16* __restrict__ dqkv,
                   int N, int n_pad, int H, int hd, float scale, float scale_log2, int dbg) {
  using L = UbSmem<HDP, KV>;
  using TL = UaTile<HDP>;
  constexpr int NST = L::NST;
  static_assert(!TS || HDP == 32, "operands in TMEM need the 256-column budget of HDP = 32");
  static_assert(HDP != 128 || (TMA && !TS && !PP && !SP && MW == 2), "head_dim > 64: TMA loaders, operands through shared memory, two lockstep warpgroups");
  constexpr int NH = TL::NH;
  // TMEM: scores 2 x 64 | outputs 2 x HDP.  HDP = 128 needs 384 columns -> a 512-column allocation (one CTA per SM)
  constexpr uint32_t UB_TMEM_COLS = HDP == 128 ? 512 : 256;
  static_assert(!PP || MW == 2, "the decoupled schedule needs two math warpgroups");
  static_assert(!SP || (KV && HDP == 32 && TMA && TS && MW == 2 && !PP), "single pass: KV pass, HDP 32, TMA, P^T in TMEM, two lockstep warpgroups");
  constexpr uint32_t DQ_COL = 224;                            // SP: dQ tile (M = 64: rows 16j + i sit in lane 32j + i)
  constexpr uint32_t O1_COL = UB_O1_COL, O2_COL = TS ? UB_O1_COL + 32 : (HDP == 128 ? UB_O1_COL + 128 : UB_O2_COL);
  constexpr uint32_t PT_COL = 192, DST_COL = 224;             // TS: packed bf16 P^T / dS^T, 32 columns each
  extern __shared__ __align__(1024) uint8_t ub_raw[];
  const uint32_t base = ua_smem(ub_raw);
  const uint32_t sR1 = base + L::R1, sR2 = base + L::R2, sC1 = base + L::C1, sC2 = base + L::C2;
  const uint32_t sDS = base + L::DS, sP = base + L::P;
  const uint32_t bars = base + L::BARS;
  const uint32_t c_full = bars;             // [NST <= 4]
  const uint32_t c_empty = bars + 32;       // [NST <= 4]
  const uint32_t t_full = bars + 64;        // [2] (PP: one per column half; else only [0])
  const uint32_t t_free = bars + 80;        // [2]
  const uint32_t p_full = bars + 96;        // [2]
  const uint32_t o_done = bars + 112;       // [2]
  const uint32_t tmem_slot = bars + 128;
  const uint32_t r_ready = bars + 136;      // count 32: pad columns of the stationary tiles zeroed (TMA, hd < HDP)
  const uint32_t all_done = bars + 144;     // count 1: every output MMA of the CTA has retired
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(ub_raw + L::BARS + 128);

  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, r0 = blockIdx.x * UB_BM;
  const int64_t rs = 3 * (int64_t)H * hd;                    // qkv row stride (elements)
  const int64_t os = (int64_t)H * hd;                        // out / dout row stride
  const bf16* qb = qkv + (int64_t)b * N * rs + (int64_t)h * hd;
  const bf16* kb = qb + (int64_t)H * hd;
  const bf16* vb = kb + (int64_t)H * hd;
  const bf16* dob = dout + (int64_t)b * N * os + (int64_t)h * hd;
  // row statistics prepared by fa_delta_kernel: lse * log2(e) and delta * scale, rows padded to n_pad
  const float* lse_bh = lse2 + ((int64_t)b * H + h) * n_pad;
  const float* delta_bh = delta + ((int64_t)b * H + h) * n_pad;
  const int T = (N + UB_BN - 1) / UB_BN;
  // A TMA box always spans HDP columns: for hd < HDP columns [hd, HDP) hold the neighbouring head.  Both score
  // products contract over those columns with one STATIONARY operand (S: R1, dP: R2), so zeroing the pads of
  // R1 / R2 once per CTA is enough; the pad columns of dQ / dK / dV are never stored.
  // (HDP = 128: head_dim is a multiple of 16 and the score MMAs stop after hd / 16 K-steps instead.)
  const bool zero_pad = TMA && hd < HDP && HDP != 128;

  if (threadIdx.x == 0) {
    if (base & 1023u) __trap();
    ua_mbar_init(r_ready, 32);
    for (int i = 0; i < NST; ++i) { ua_mbar_init(c_full + 8 * i, TMA ? 1 : UB_LOADERS); ua_mbar_init(c_empty + 8 * i, 1); }
    constexpr uint32_t NB = PP ? 2 : 1, CNT = PP ? 128 : 128 * MW;
    for (uint32_t i = 0; i < NB; ++i) {
      ua_mbar_init(t_full + 8 * i, 1); ua_mbar_init(t_free + 8 * i, CNT); ua_mbar_init(p_full + 8 * i, CNT); ua_mbar_init(o_done + 8 * i, 1);
    }
    ua_mbar_init(all_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)UB_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  ua_fence_before();
  __syncthreads();
  ua_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  pdl_wait();

  if (warp < 4) {
  ua_reg_dec<MW == 2 ? 40 : 56>();
  if (warp == 0 || (!TMA && (warp == 2 || warp == 3))) {
    // ============================ loader ============================
    const int ld_tid = warp == 0 ? lane : (warp - 1) * 32 + lane;     // 0..95 on the cp.async path
    const int ld_n = TMA ? 32 : UB_LOADERS;
    // stationary tiles: KV pass K, V ; Q pass Q, dO.   streamed tiles: KV pass Q, dO ; Q pass K, V.
    const int c_r1 = KV ? (H + h) * hd : h * hd;              // column of R1 inside a qkv row
    const int c_c1 = KV ? h * hd : (H + h) * hd;
    if (TMA && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_qkv64) : "memory");
    if (!TMA) {
      ua_stage<HDP, 128>(sR1, KV ? kb : qb, rs, r0, N, hd, ld_tid, ld_n);
      if (KV) ua_stage<HDP, 128>(sR2, vb, rs, r0, N, hd, ld_tid, ld_n);
      else    ua_stage<HDP, 128>(sR2, dob, os, r0, N, hd, ld_tid, ld_n);
    }
    for (int t = 0; t < T; ++t) {
      const int st = t % NST;
      if (!TMA && t >= NST - 1) {
        // tile j = t-(NST-1) has landed once at most NST-2 copy groups are pending.  It is signalled BEFORE
        // waiting for stage `st` to drain: the MMA warp issues scores(j) ahead of outputs(t-NST), so the
        // opposite order would deadlock.
        asm volatile("cp.async.wait_group %0;" ::"n"(NST - 2) : "memory");
        ua_fence_async_smem();
        ua_mbar_arrive(c_full + 8 * ((t - (NST - 1)) % NST));
      }
      if (t >= NST) ua_mbar_wait(c_empty + 8 * st, ((t / NST) & 1) ^ 1);
      const uint32_t c1 = sC1 + st * L::COL_TILE, c2 = sC2 + st * L::COL_TILE;
      if (TMA) {
        if (lane == 0) {
          const uint32_t fb = c_full + 8 * st;
          ua_expect_tx(fb, 2 * L::COL_TILE + (t == 0 ? 2 * L::ROW_TILE : 0));
#pragma unroll
          for (int hf = 0; hf < NH; ++hf) {                       // one box per 64-column half
            const int dc = 64 * hf;
            const uint32_t ro = hf * L::RHALF, co = hf * L::CHALF;
            if (t == 0) {
              ua_tma3d(sR1 + ro, &map_qkv128, fb, c_r1 + dc, r0, b);
              if (KV) ua_tma3d(sR2 + ro, &map_qkv128, fb, (2 * H + h) * hd + dc, r0, b);
              else    ua_tma3d(sR2 + ro, &map_do, fb, h * hd + dc, r0, b);
            }
            ua_tma3d(c1 + co, &map_qkv64, fb, c_c1 + dc, t * UB_BN, b);
            if (KV) ua_tma3d(c2 + co, &map_do, fb, h * hd + dc, t * UB_BN, b);
            else    ua_tma3d(c2 + co, &map_qkv64, fb, (2 * H + h) * hd + dc, t * UB_BN, b);
          }
        }
      } else {
        ua_stage<HDP, 64>(c1, KV ? qb : kb, rs, t * UB_BN, N, hd, ld_tid, ld_n);
        if (KV) ua_stage<HDP, 64>(c2, dob, os, t * UB_BN, N, hd, ld_tid, ld_n);
        else    ua_stage<HDP, 64>(c2, vb, rs, t * UB_BN, N, hd, ld_tid, ld_n);
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
    }
    if (!TMA) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      ua_fence_async_smem();
      for (int j = max(0, T - (NST - 1)); j < T; ++j) ua_mbar_arrive(c_full + 8 * (j % NST));
    }
  } else if (TMA && warp == 2) {
    if (zero_pad) {
      ua_mbar_wait(c_full, 0);                                  // R1, R2 (and tile 0) have landed
      ua_zero_pad<HDP>(sR1, 128, hd, lane);
      ua_zero_pad<HDP>(sR2, 128, hd, lane);
      ua_fence_async_smem();
      ua_mbar_arrive(r_ready);
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    if (lane == 0) {
      const uint32_t idesc_t = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(UB_BN >> 3) << 17) | ((uint32_t)(UB_BM >> 4) << 24);
      const uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)((HDP == 128 ? 64 : HDP) >> 3) << 17) |
                               ((uint32_t)(UB_BM >> 4) << 24);
      // HDP = 128: columns [64, hd) of every output come from the second half of the streamed tile
      const uint32_t idesc_o2 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)((hd - 64) >> 3) << 17) | ((uint32_t)(UB_BM >> 4) << 24);
      const int ksteps = HDP == 128 ? hd / 16 : HDP / 16;
      const uint64_t r1d = TL::kmajor(sR1), r2d = TL::kmajor(sR2);
      auto issue_scores = [&](int t) {
        const uint64_t c1d = TL::kmajor(sC1 + (t % NST) * L::COL_TILE);
        const uint64_t c2d = TL::kmajor(sC2 + (t % NST) * L::COL_TILE);
        if constexpr (HDP == 128) {
          for (int k = 0; k < ksteps; ++k) {
            const uint32_t kr = (uint32_t)(k >> 2) * (L::RHALF >> 4) + 2 * (k & 3), kc = (uint32_t)(k >> 2) * (L::CHALF >> 4) + 2 * (k & 3);
            ua_mma(tmem, r1d + kr, c1d + kc, idesc_t, k > 0 ? 1u : 0u);
          }
          for (int k = 0; k < ksteps; ++k) {
            const uint32_t kr = (uint32_t)(k >> 2) * (L::RHALF >> 4) + 2 * (k & 3), kc = (uint32_t)(k >> 2) * (L::CHALF >> 4) + 2 * (k & 3);
            ua_mma(tmem + UB_T2_COL, r2d + kr, c2d + kc, idesc_t, k > 0 ? 1u : 0u);
          }
        } else {
#pragma unroll
          for (int k = 0; k < HDP / 16; ++k) ua_mma(tmem, r1d + 2 * k, c1d + 2 * k, idesc_t, k > 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < HDP / 16; ++k) ua_mma(tmem + UB_T2_COL, r2d + 2 * k, c2d + 2 * k, idesc_t, k > 0 ? 1u : 0u);
        }
        ua_commit(t_full);
      };
      // one column half (32 streamed rows) of both score products; PP only
      const uint32_t idesc_h = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(UB_BM >> 4) << 24);
      auto issue_scores_half = [&](int t, int hf) {
        const uint32_t off = (uint32_t)hf * 32u * TL::PITCH;                 // 32 streamed rows further into the tile
        const uint64_t c1d = TL::kmajor(sC1 + (t % NST) * L::COL_TILE + off);
        const uint64_t c2d = TL::kmajor(sC2 + (t % NST) * L::COL_TILE + off);
#pragma unroll
        for (int k = 0; k < HDP / 16; ++k) ua_mma(tmem + 32 * hf, r1d + 2 * k, c1d + 2 * k, idesc_h, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < HDP / 16; ++k) ua_mma(tmem + UB_T2_COL + 32 * hf, r2d + 2 * k, c2d + 2 * k, idesc_h, k > 0 ? 1u : 0u);
        ua_commit(t_full + 8 * hf);
      };
      // output MMAs of step t over the K steps [k0, k1) of the 64 streamed columns
      auto issue_outputs = [&](int t, int k0, int k1) {
        const uint32_t c1 = sC1 + (t % NST) * L::COL_TILE, c2 = sC2 + (t % NST) * L::COL_TILE;
        const uint64_t dsd = ua_desc(sDS, 1, 64);
        const uint64_t c1m = TL::mnmajor(c1);                // MN-major view: 16 streamed rows per K step
        if (KV) {
          const uint64_t pd = ua_desc(sP, 1, 64);
          const uint64_t c2m = TL::mnmajor(c2);
          if constexpr (SP) {
            // three INDEPENDENT accumulation chains (dV: 4 MMAs, dK: 4, dQ tile: 8).  Small dependent MMAs are latency-, not
            // throughput-bound (ncu: tensor pipe 21 % busy while the math warps wait for these 16 MMAs), so the chains are
            // issued round-robin instead of one after the other.
            // dQ once per PAIR of steps: M = 128 queries (even step = chunk 0 in sDS, odd step = chunk 1 in sDS + 16 KB, the
            // chunk stride is the descriptor's leading-byte offset), 8 MMAs per pair instead of 8 per step -- the step is
            // bound by the NUMBER of small MMAs (each occupies the tensor pipe ~67 cycles whatever its size).  A last
            // unpaired step (T odd) keeps the M = 64 form on buffer 0.
            const bool pair = (t & 1) != 0, tail = !pair && (t == T - 1);
            const uint32_t idesc_q = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(HDP >> 3) << 17) |
                                     ((uint32_t)((pair ? 128 : 64) >> 4) << 24);
            const uint64_t dsm = ua_desc(sDS, UB_PD_TILE >> 4, 64);   // MN-major view of dS^T: 64 queries contiguous per chunk, 16 keys per K step
            const uint64_t dsk = ua_desc(sDS + (t & 1) * UB_PD_TILE, 1, 64);   // K-major view of THIS step's tile (dK)
            const uint64_t r1m = TL::mnmajor(sR1);           // K tile: 16 keys per K step, hd contiguous
#pragma unroll
            for (int k = 0; k < UB_BN / 16; ++k) {
              const uint32_t acc = (t > 0 || k > 0) ? 1u : 0u;
              if (pair || tail) ua_mma(tmem + DQ_COL, dsm + 128 * (2 * k), r1m + TL::MN_KADV * (2 * k), idesc_q, k > 0 ? 1u : 0u);
              ua_mma_ts(tmem + O1_COL, tmem + PT_COL + 8 * k, c2m + TL::MN_KADV * k, idesc_o, acc);           // dV += P^T dO
              if (pair || tail) ua_mma(tmem + DQ_COL, dsm + 128 * (2 * k + 1), r1m + TL::MN_KADV * (2 * k + 1), idesc_q, 1u);
              ua_mma(tmem + O2_COL, dsk + 2 * k, c1m + TL::MN_KADV * k, idesc_o, acc);                        // dK += dS^T Q
            }
          } else {
#pragma unroll
          for (int k = k0; k < k1; ++k) {                                                      // dV += P^T dO
            const uint32_t acc = (t > 0 || k > 0) ? 1u : 0u;
            if (TS) ua_mma_ts(tmem + O1_COL, tmem + PT_COL + 8 * k, c2m + TL::MN_KADV * k, idesc_o, acc);
            else    ua_mma(tmem + O1_COL, pd + 2 * k, c2m + TL::MN_KADV * k, idesc_o, acc);
            if constexpr (HDP == 128) ua_mma(tmem + O1_COL + 64, pd + 2 * k, c2m + (L::CHALF >> 4) + TL::MN_KADV * k, idesc_o2, acc);
          }
#pragma unroll
          for (int k = k0; k < k1; ++k) {                                                      // dK += dS^T Q
            const uint32_t acc = (t > 0 || k > 0) ? 1u : 0u;
            if (TS) ua_mma_ts(tmem + O2_COL, tmem + DST_COL + 8 * k, c1m + TL::MN_KADV * k, idesc_o, acc);
            else    ua_mma(tmem + O2_COL, dsd + 2 * k, c1m + TL::MN_KADV * k, idesc_o, acc);
            if constexpr (HDP == 128) ua_mma(tmem + O2_COL + 64, dsd + 2 * k, c1m + (L::CHALF >> 4) + TL::MN_KADV * k, idesc_o2, acc);
          }
          }
        } else {
#pragma unroll
          for (int k = k0; k < k1; ++k) {                                                      // dQ += dS K
            const uint32_t acc = (t > 0 || k > 0) ? 1u : 0u;
            if (TS) ua_mma_ts(tmem + O1_COL, tmem + DST_COL + 8 * k, c1m + TL::MN_KADV * k, idesc_o, acc);
            else    ua_mma(tmem + O1_COL, dsd + 2 * k, c1m + TL::MN_KADV * k, idesc_o, acc);
            if constexpr (HDP == 128) ua_mma(tmem + O1_COL + 64, dsd + 2 * k, c1m + (L::CHALF >> 4) + TL::MN_KADV * k, idesc_o2, acc);
          }
        }
      };
      ua_mbar_wait(c_full, 0);
      if (zero_pad) ua_mbar_wait(r_ready, 0);
      ua_fence_after();
      if constexpr (PP) {
        issue_scores_half(0, 0);
        issue_scores_half(0, 1);
        for (int t = 0; t < T; ++t) {
          if (t + 1 < T) ua_mbar_wait(c_full + 8 * ((t + 1) % NST), ((t + 1) / NST) & 1);
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            if (t + 1 < T) {
              ua_mbar_wait(t_free + 8 * hf, t & 1);            // this half's scores of step t are in registers
              ua_fence_after();
              issue_scores_half(t + 1, hf);
            }
            ua_mbar_wait(p_full + 8 * hf, t & 1);
            ua_fence_after();
            issue_outputs(t, 2 * hf, 2 * hf + 2);
            ua_commit(o_done + 8 * hf);
          }
          ua_commit(c_empty + 8 * (t % NST));
        }
      } else {
        issue_scores(0);
        for (int t = 0; t < T; ++t) {
          if (t + 1 < T) {
            ua_mbar_wait(c_full + 8 * ((t + 1) % NST), ((t + 1) / NST) & 1);
            ua_mbar_wait(t_free, t & 1);                       // both score tiles of step t are in registers
            ua_fence_after();
            issue_scores(t + 1);
          }
          ua_mbar_wait(p_full, t & 1);
          ua_fence_after();
          issue_outputs(t, 0, UB_BN / 16);
          ua_commit(o_done);
          ua_commit(c_empty + 8 * (t % NST));
        }
      }
      ua_commit(all_done);
    }
  }
  } else {
    // register pool per CTA: 256 x 128 (MW = 1) -> 128 x 56 + 128 x 200; 384 x 80 (MW = 2) -> 128 x 40 + 256 x 96
    ua_reg_inc<MW == 2 ? 96 : 200>();
    // ============================ score math ============================
    const int q = warp & 3;
    const int wg = (warp - 4) >> 2;                           // math warpgroup: columns [wg * CW, +CW) of every step
    constexpr int CW = UB_BN / MW;
    const int col0 = wg * CW;
    const int row = q * 32 + lane;                            // TMEM lane = stationary row in tile
    const uint32_t t_1 = tmem + ((uint32_t)(q * 32) << 16);
    const int ri = r0 + row;
    float my_l2 = 0.f, my_dl = 0.f;
    if (!KV && ri < N) { my_l2 = lse_bh[ri]; my_dl = delta_bh[ri]; }
    float* stat = reinterpret_cast<float*>(ub_raw + L::STAT);
    float my_stat = 0.f;
    // PP: barriers of my column half; statistics staged per warpgroup (threads 0-31: lse2 of my 32 columns, 32-63: delta)
    const uint32_t bo = PP ? 8u * (uint32_t)wg : 0u;
    const float* pp_src = (row < 32 ? lse_bh + col0 + row : delta_bh + col0 + (row - 32));
    if (KV && !PP && wg == 0) my_stat = __ldg((row < 64 ? lse_bh : delta_bh - 64) + row);   // tile 0 (rows are padded to n_pad with zeros)
    if (KV && PP && row < 64) my_stat = __ldg(pp_src);
    // SP: dQ tile of step tq (fp32 [64 q, hd], M = 64 accumulator: row 16j + i in TMEM lane 32j + i) -> dense staging tile in
    // the (otherwise unused) sP region -> one TMA reduce-add into dq_acc[b, tq*64 .., h, :].  Warp (q, wg) moves columns
    // [16 wg, 16 wg + 16) of rows [16 q, 16 q + 16).  The staging tile is known to be free: the issuing thread waited for the
    // previous reduce to have READ it before this step's statistics barrier (bar.sync 1), which every math thread has passed.
    const bool dq_issuer = (threadIdx.x == 128);
    auto flush_dq = [&](int tq, bool pair) {       // pair: 128 queries (tiles tq, tq + 1), TMEM row == lane; else 64 (M = 64 layout)
      uint32_t v[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
            "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
          : "r"(t_1 + DQ_COL + 16 * wg) : "memory");
      ua_ld_wait();
      if (pair || lane < 16) {
        const uint32_t dst = sP + (uint32_t)((pair ? 32 * q + lane : 16 * q + lane) * hd + 16 * wg) * 4u;
#pragma unroll
        for (int c = 0; c < 16; c += 4)
          if (16 * wg + c < hd)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + 4 * c), "r"(v[c]), "r"(v[c + 1]), "r"(v[c + 2]), "r"(v[c + 3]) : "memory");
      }
      ua_fence_async_smem();
      asm volatile("bar.sync 3, %0;" ::"n"(128 * MW) : "memory");
      if (dq_issuer && !(dbg & 1)) {
        asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                     ::"l"(pair ? &map_dq2 : &map_dq), "r"(sP), "r"(h * hd), "r"(tq * UB_BN), "r"(b) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    };
    for (int t = 0; t < T; ++t) {
      ua_mbar_wait(t_full + bo, t & 1);
      ua_fence_after();
      float s[CW], dp[CW];
#pragma unroll
      for (int c = 0; c < CW; c += 32) { ua_ld32(t_1 + col0 + c, s + c); ua_ld32(t_1 + UB_T2_COL + col0 + c, dp + c); }
      ua_ld_wait();
      ua_fence_before();
      ua_mbar_arrive(t_free + bo);
      uint32_t pk_p[KV ? CW / 2 : 1], pk_d[CW / 2];
      // per-column statistics of this query tile: thread i publishes ONE value (i < 64: lse2 of column i,
      // else delta of column i - 64) that it loaded a whole step earlier, so no thread ever waits on global
      // memory; everybody then reads the 2 x 256 bytes as shared-memory broadcasts
      if constexpr (KV && PP) {
        float* sb = stat + (t % L::SB) * 128 + wg * 64;       // [lse2 of my 32 columns | delta of my 32 columns]
        auto wg_sync = [&] {                                   // named barrier of my warpgroup (immediate ids: a register id reserves all 16)
          if (wg == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
        };
        if (L::SB == 1) wg_sync();                             // my warpgroup's readers of the previous tile are done
        if (row < 64) sb[row] = my_stat;
        wg_sync();
        if (row < 64 && t + 1 < T) my_stat = __ldg(pp_src + (t + 1) * UB_BN);
      } else if constexpr (KV) {
        float* sb = stat + (t % L::SB) * 128;
        if (SP && dq_issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging tile free again
        if (L::SB == 1) asm volatile("bar.sync 1, %0;" ::"n"(128 * MW) : "memory");   // previous tile's readers are done
        if (wg == 0) sb[row] = my_stat;
        asm volatile("bar.sync 1, %0;" ::"n"(128 * MW) : "memory");
        if (wg == 0 && t + 1 < T) my_stat = __ldg((row < 64 ? lse_bh : delta_bh - 64) + (t + 1) * UB_BN + row);
      }
      const float* st_l2 = stat + (KV ? (t % L::SB) * 128 : 0) + (PP ? wg * 64 : col0);
      const float* st_dl = st_l2 + (PP ? 32 : 64);
#pragma unroll
      for (int j = 0; j < CW; j += 4) {
        float l2[4], dl[4];
        if constexpr (KV) {
          const float4 a = *reinterpret_cast<const float4*>(st_l2 + j);
          const float4 d = *reinterpret_cast<const float4*>(st_dl + j);
          l2[0] = a.x; l2[1] = a.y; l2[2] = a.z; l2[3] = a.w;
          dl[0] = d.x; dl[1] = d.y; dl[2] = d.z; dl[3] = d.w;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) { l2[e] = my_l2; dl[e] = my_dl; }
        }
        float p[4], ds[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          p[e] = ua_exp2(fmaf(s[j + e], scale_log2, -l2[e]));
          ds[e] = p[e] * fmaf(dp[j + e], scale, -dl[e]);          // dl is pre-multiplied by scale
        }
        if constexpr (KV) { pk_p[j >> 1] = pack_bf16x2(p[0], p[1]); pk_p[(j >> 1) + 1] = pack_bf16x2(p[2], p[3]); }
        pk_d[j >> 1] = pack_bf16x2(ds[0], ds[1]); pk_d[(j >> 1) + 1] = pack_bf16x2(ds[2], ds[3]);
      }
      if (t > 0) ua_mbar_wait(o_done + bo, (t - 1) & 1);      // output MMAs of step t-1 done: P / dS smem is ours
      ua_fence_after();
      // dQ of the pair of steps (t-2, t-1), complete since the MMAs of the odd step t-1: TMEM -> staging -> TMA reduce-add
      if constexpr (SP) { if (t > 0 && ((t - 1) & 1) && !(dbg & 2)) flush_dq(t - 2, true); }
      if constexpr (SP) {
        // P^T -> tensor memory (A operand of dV), dS^T -> shared memory (A operand of dK, K-major, and of dQ, MN-major)
        ua_st_regs<CW / 2>(t_1 + PT_COL + col0 / 2, pk_p);
#pragma unroll
        for (int c = 0; c < CW / 8; ++c) {
          const uint32_t off = (t & 1) * UB_PD_TILE + row * 128 + (((col0 / 8 + c) ^ (row & 7)) << 4);   // even / odd step buffer
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sDS + off), "r"(pk_d[4 * c]), "r"(pk_d[4 * c + 1]),
                       "r"(pk_d[4 * c + 2]), "r"(pk_d[4 * c + 3]) : "memory");
        }
        ua_st_wait();
        ua_fence_async_smem();
      } else if constexpr (TS) {
        ua_st_regs<CW / 2>(t_1 + DST_COL + col0 / 2, pk_d);
        if constexpr (KV) ua_st_regs<CW / 2>(t_1 + PT_COL + col0 / 2, pk_p);
        ua_st_wait();
      } else {
#pragma unroll
        for (int c = 0; c < CW / 8; ++c) {                    // my chunks of 8 streamed columns
          const uint32_t off = row * 128 + (((col0 / 8 + c) ^ (row & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sDS + off), "r"(pk_d[4 * c]), "r"(pk_d[4 * c + 1]),
                       "r"(pk_d[4 * c + 2]), "r"(pk_d[4 * c + 3]) : "memory");
          if constexpr (KV)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sP + off), "r"(pk_p[4 * c]), "r"(pk_p[4 * c + 1]),
                         "r"(pk_p[4 * c + 2]), "r"(pk_p[4 * c + 3]) : "memory");
        }
        ua_fence_async_smem();
      }
      ua_fence_before();
      ua_mbar_arrive(p_full + bo);
    }
    if constexpr (SP) {
      if (dq_issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      asm volatile("bar.sync 3, %0;" ::"n"(128 * MW) : "memory");
    }
    ua_mbar_wait(all_done, 0);
    ua_fence_after();
    if constexpr (SP) {
      if (T & 1) flush_dq(T - 1, false);             // unpaired last step
      else flush_dq(T - 2, true);
      if (dq_issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    // dqkv row layout [3][H][hd]: slot 0 = dQ, 1 = dK, 2 = dV
    bf16* drow = dqkv + ((int64_t)b * N + ri) * rs + (int64_t)h * hd;
#pragma unroll
    for (int o = 0; o < (KV ? 2 : 1); ++o) {
      const int slot = KV ? (o == 0 ? 2 : 1) : 0;
      const uint32_t t_o = t_1 + (o == 0 ? O1_COL : O2_COL);
#pragma unroll
      for (int c = 0; c < HDP; c += 32) {
        if ((o * (HDP / 32) + c / 32) % MW != wg) continue;   // 32-column output blocks round-robin over the warpgroups
        if (HDP == 128 && c >= hd) continue;                  // never written
        float v[32];
        ua_ld32(t_o + c, v);
        ua_ld_wait();
        if (ri < N) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (c + j < hd) {
              float v8[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) v8[e] = v[j + e];
              store8<bf16>(drow + (int64_t)slot * H * hd + c + j, v8);
            }
          }
        }
      }
    }
  }

  ua_fence_before();
  __syncthreads();
  if (warp == 1) {
    ua_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)UB_TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------
int avj_attention_delta(const void* out, const void* dout, float* delta, const float* lse, float* lse2_out, float scale,
                        int n_pad, int B, int N, int H, int hd, cudaStream_t s);
int ua_make_map3d(const void* ptr, int B, int N, int cols, int box_rows, CUtensorMap* out, int box_cols);

bool avj_attention_umma_bwd_supported(int dtype, int hd) {
  // head_dim <= 64: multiples of 8; 80 .. 128: multiples of 16 (two 64-column halves per tile)
  return dtype == AVJ_BF16 && ((hd % 8 == 0 && hd >= 8 && hd <= 64) || (hd % 16 == 0 && hd > 64 && hd <= 128));
}

static int ub_dbg() {     // AVJ_ATTN_BWD_DBG (timing experiments only; results are WRONG when set): 1 = no TMA reduce, 2 = no dQ flush
  static int v = -1;
  if (v < 0) { const char* e = getenv("AVJ_ATTN_BWD_DBG"); v = e ? atoi(e) : 0; }
  return v;
}

template <int HDP, bool TMA, bool KV, int MW, bool TS, bool PP, bool SP = false>
static int ub_launch(const CUtensorMap& m128, const CUtensorMap& m64, const CUtensorMap& mdo, const CUtensorMap& mdq, const CUtensorMap& mdq2,
                     const bf16* qkv, const bf16* dout,
                     const float* lse2, const float* delta, bf16* dqkv, int B, int N, int n_pad, int H, int hd, float scale,
                     cudaStream_t s) {
  static bool set = false;
  const int smem = (int)UbSmem<HDP, KV>::TOTAL;
  if (!set) {
    cudaError_t e = cudaFuncSetAttribute(fa_bwd_umma_kernel<HDP, TMA, KV, MW, TS, PP, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    AVJ_CHECK(e == cudaSuccess, "cudaFuncSetAttribute(fa_bwd_umma_kernel) failed: %s", cudaGetErrorString(e));
    cudaFuncSetAttribute(fa_bwd_umma_kernel<HDP, TMA, KV, MW, TS, PP, SP>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    set = true;
  }
  dim3 grid((N + UB_BM - 1) / UB_BM, H, B);
  avj_launch_pdl(fa_bwd_umma_kernel<HDP, TMA, KV, MW, TS, PP, SP>, grid, dim3(128 + 128 * MW), (size_t)smem, s, m128, m64, mdo, mdq, mdq2, qkv, dout, lse2, delta, dqkv, N, n_pad,
                 H, hd, scale, scale * 1.4426950408889634f, ub_dbg());
  AVJ_LAUNCH_CHECK();
  return 0;
}

template <int HDP, bool TMA, int MW, bool TS, bool PP>
static int ub_both(const bf16* qkv, const bf16* dout, const float* lse2, const float* delta, bf16* dqkv,
                   int B, int N, int n_pad, int H, int hd, float scale, cudaStream_t s) {
  CUtensorMap m128, m64, mdo64, mdo128;
  memset(&m128, 0, sizeof(m128)); memset(&m64, 0, sizeof(m64)); memset(&mdo64, 0, sizeof(mdo64)); memset(&mdo128, 0, sizeof(mdo128));
  if (TMA) {
    int rc;
    constexpr int BC = HDP == 128 ? 64 : HDP;                   // box columns
    if ((rc = ua_make_map3d(qkv, B, N, 3 * H * hd, 128, &m128, BC))) return rc;
    if ((rc = ua_make_map3d(qkv, B, N, 3 * H * hd, 64, &m64, BC))) return rc;
    if ((rc = ua_make_map3d(dout, B, N, H * hd, 64, &mdo64, BC))) return rc;
    if ((rc = ua_make_map3d(dout, B, N, H * hd, 128, &mdo128, BC))) return rc;
  }
  int rc = ub_launch<HDP, TMA, true, MW, TS, PP>(m128, m64, mdo64, m128, m128, qkv, dout, lse2, delta, dqkv, B, N, n_pad, H, hd, scale, s);
  if (rc) return rc;
  return ub_launch<HDP, TMA, false, MW, TS, PP>(m128, m64, mdo128, m128, m128, qkv, dout, lse2, delta, dqkv, B, N, n_pad, H, hd, scale, s);
}

// fp32 [B, N, cols] viewed as {cols, N, B}; box = {box_cols, box_rows, 1}, no swizzle (TMA reduce-add target of the dQ tiles)
static int ub_make_map3d_f32(const void* ptr, int B, int N, int cols, int box_rows, int box_cols, CUtensorMap* out) {
  typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static PFN enc = nullptr;
  if (!enc) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      enc = reinterpret_cast<PFN>(p);
  }
  AVJ_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)N, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)cols * 4, (cuuint64_t)cols * 4 * (cuuint64_t)N};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  AVJ_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(dq accumulator) failed (%d)", (int)r);
  return 0;
}

int avj_copy_rows(const void* in, int in_dtype, int ld_in, avj_rowmap imap, void* out, int out_dtype, int ld_out, avj_rowmap omap,
                  int rows, int D, int accumulate, void* stream);

// single pass (head_dim <= 32, TMA): zero the fp32 dQ accumulator, ONE kernel for dK / dV / dQ, then dQ -> bf16 into dqkv[..., 0, :, :]
static int ub_single(const bf16* qkv, const bf16* dout, const float* lse2, const float* delta, bf16* dqkv, float* dq_acc,
                     int B, int N, int n_pad, int H, int hd, float scale, cudaStream_t s) {
  CUtensorMap m128, m64, mdo64, mdq, mdq2;
  memset(&m128, 0, sizeof(m128)); memset(&m64, 0, sizeof(m64)); memset(&mdo64, 0, sizeof(mdo64)); memset(&mdq, 0, sizeof(mdq));
  memset(&mdq2, 0, sizeof(mdq2));
  int rc;
  if ((rc = ua_make_map3d(qkv, B, N, 3 * H * hd, 128, &m128, 32))) return rc;
  if ((rc = ua_make_map3d(qkv, B, N, 3 * H * hd, 64, &m64, 32))) return rc;
  if ((rc = ua_make_map3d(dout, B, N, H * hd, 64, &mdo64, 32))) return rc;
  if ((rc = ub_make_map3d_f32(dq_acc, B, N, H * hd, 64, hd, &mdq))) return rc;
  if ((rc = ub_make_map3d_f32(dq_acc, B, N, H * hd, 128, hd, &mdq2))) return rc;     // pairs of query tiles
  rc = ub_launch<32, true, true, 2, true, false, true>(m128, m64, mdo64, mdq, mdq2, qkv, dout, lse2, delta, dqkv, B, N, n_pad, H, hd, scale, s);
  if (rc) return rc;
  const avj_rowmap ident = {0, 0, 0};
  return avj_copy_rows(dq_acc, AVJ_F32, H * hd, ident, dqkv, AVJ_BF16, 3 * H * hd, ident, B * N, H * hd, 0, s);
}

int avj_attention_bwd_umma(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* ws,
                           int B, int N, int H, int hd, float scale, cudaStream_t s) {
  // workspace: delta * scale | lse * log2(e), each [B, H, n_pad]
  const int n_pad = (N + 63) / 64 * 64;
  float* delta = ws;
  float* lse2 = ws + (int64_t)B * H * n_pad;
  float* dq_acc = lse2 + (int64_t)B * H * n_pad;                 // [B, N, H, hd] fp32 (single-pass path)
  static int sp = -1;          // AVJ_ATTN_BWD_SP=0: two passes (KV pass + Q pass) for head_dim <= 32 as well
  if (sp < 0) { const char* e = getenv("AVJ_ATTN_BWD_SP"); sp = (e && e[0] == '0') ? 0 : 1; }
  const bool al16 = ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(dout)) & 15) == 0;
  const bool single = sp && hd <= 32 && al16 && (reinterpret_cast<uintptr_t>(dq_acc) & 15) == 0;
  if (single) AVJ_CUDA(cudaMemsetAsync(dq_acc, 0, (size_t)B * N * H * hd * sizeof(float), s));
  int rc = avj_attention_delta(out, dout, delta, lse, lse2, scale, n_pad, B, N, H, hd, s);
  if (rc) return rc;
  if (single) return ub_single((const bf16*)qkv, (const bf16*)dout, lse2, delta, (bf16*)dqkv, dq_acc, B, N, n_pad, H, hd, scale, s);
  static int use_tma = -1;
  if (use_tma < 0) { const char* e = getenv("AVJ_ATTN_TMA"); use_tma = (e && e[0] == '0') ? 0 : 1; }
  static int use_tma32 = -1;   // AVJ_ATTN_TMA32=0: head_dim <= 32 goes back to the cp.async gather loaders
  if (use_tma32 < 0) { const char* e = getenv("AVJ_ATTN_TMA32"); use_tma32 = (e && e[0] == '0') ? 0 : use_tma; }
  const bool aligned = ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(dout)) & 15) == 0;
  static int mw = -1;          // AVJ_ATTN_BWD_MW=1: one math warpgroup (256-thread CTAs) instead of two
  if (mw < 0) { const char* e = getenv("AVJ_ATTN_BWD_MW"); mw = (e && e[0] == '1') ? 1 : 2; }
  static int ts = -1;          // AVJ_ATTN_TMEM_P=0: P^T / dS^T through shared memory for head_dim <= 32 as well
  if (ts < 0) { const char* e = getenv("AVJ_ATTN_TMEM_P"); ts = (e && e[0] == '0') ? 0 : 1; }
#define UB_ARGS (const bf16*)qkv, (const bf16*)dout, lse2, delta, (bf16*)dqkv, B, N, n_pad, H, hd, scale, s
  static int pp = -1;          // AVJ_ATTN_BWD_PP=1: decoupled math warpgroups (measured SLOWER than lockstep: 0.677 vs 0.550 ms predictor, 1.59 vs 1.27 ms target)
  if (pp < 0) { const char* e = getenv("AVJ_ATTN_BWD_PP"); pp = (e && e[0] == '1') ? 1 : 0; }
#define UB_GO(HDP_, TMA_, TS_)                                                                             \
  {                                                                                                        \
    if (mw == 2 && pp) return ub_both<HDP_, TMA_, 2, TS_, true>(UB_ARGS);                                  \
    return mw == 2 ? ub_both<HDP_, TMA_, 2, TS_, false>(UB_ARGS) : ub_both<HDP_, TMA_, 1, TS_, false>(UB_ARGS); \
  }
  if (hd <= 32) {
    if (use_tma32 && aligned) { if (ts) { UB_GO(32, true, true); } UB_GO(32, true, false); }
    if (ts) { UB_GO(32, false, true); }
    UB_GO(32, false, false);
  }
  if (hd > 64) {
    AVJ_CHECK(aligned && use_tma, "attention backward: head_dim %d needs 16-byte aligned qkv / dout (TMA loaders)", hd);
    return ub_both<128, true, 2, false, false>(UB_ARGS);
  }
  if (hd == 64 && use_tma && aligned) { UB_GO(64, true, false); }
  UB_GO(64, false, false);
#undef UB_ARGS
#undef UB_GO
}
