// K6 production GEMM for sm_100a: persistent, warp-specialised, TMA -> shared memory ->
// tcgen05.mma (bf16 x bf16 -> fp32 in TMEM) -> tcgen05.ld -> fused epilogue -> global.
//
//   warp 0      : TMA producer (one elected lane), mbarrier ring (4 stages; 6 in the CTA-pair kernel)
//   warp 1      : TMEM allocator + MMA issuer (one elected lane), UMMA 128 x BN x 16 (256 x BN x 16 per CTA pair)
//   warps 2, 3  : idle (they join the producer in the two gather variants of the patch embedding)
//   warps 4..   : 8 or 16 epilogue warps; warp w owns TMEM lanes [32*(w%4), +32) and a share of the columns
//
// CTA tile 128 x BN x 64 (CTA pair: 256 x BN x 64) with BN in {64,128,192,256} chosen to divide N.  Two TMEM accumulator
// buffers (2 x 256 columns) let the epilogue of tile i overlap the main loop of tile i+1.
// Operand layouts: K-major (activations x weights^T, "NT"), MN-major B (dgrad, "NN") and
// MN-major A and B (wgrad, "TN") -- the MN-major tiles are loaded as 64x64 TMA boxes and
// described to the tensor core with the MN-major SWIZZLE_128B canonical layout, so no
// transposed copy of any operand is ever materialised.  Weight gradients (tiny M x N, long K)
// are split along K over CTAs and reduced with fp32 vector atomics into the gradient buffer.
//
// Roofline: tensor-bound.  Algorithmic FLOPs per launch = 2*M*N*K.
#include "common.cuh"

#include <cuda.h>
#include <mutex>
#include <unordered_map>
#include <stdlib.h>

#define UG_BM 128
#define UG_BK 64
#define UG_MAX_BN 256
#define UG_STAGES 4
#define UG_A_STAGE_BYTES (UG_BM * UG_BK * 2)        // 16 KB
#define UG_B_STAGE_BYTES (UG_MAX_BN * UG_BK * 2)    // 32 KB
#define UG_EPI_WARPS 8
#define UG_THREADS (128 + 32 * UG_EPI_WARPS)   // warpgroup 0: producer, MMA, 2 idle; warpgroups 1-2: epilogue
#define UG_EPI_STAGE_BYTES 4096                     // per epilogue warp: 32 rows x 32 fp32, swizzled
#define UG_SMEM_BYTES (UG_STAGES * (UG_A_STAGE_BYTES + UG_B_STAGE_BYTES) + 1024 + 256 + UG_EPI_WARPS * UG_EPI_STAGE_BYTES)

struct UmmaParams {
  int M, N, K, ldc;
  int block_n;
  int tiles_m, tiles_n;
  int k_blocks;          // ceil(K / 64)
  int split_k;           // >= 1
  int kb_per_split;
  int a_mn_major, b_mn_major;
  uint32_t mn_lbo, mn_sbo, mn_kadv;   // MN-major descriptor fields / k-advance (16 B units)
  void* C;
  int tma_store;         // thread==row epilogues: bf16 C (and pre_out) leave through smem + TMA bulk stores
  int gelu_exact;        // AVJ_GELU_EXACT=1: erff() GELU / GELU' (generic epilogue) instead of the fitted sigmoid form
  // ---- im2col-free patch embedding (gemm_umma_kernel<..., IM2COL = true> only): A rows are tokens whose 16 x 16 patch
  //      rows are gathered straight out of the clip / spectrogram into the operand tile; both operands are fp32 in
  //      shared memory and the MMA is kind::tf32 (k-block = 32 elements = two patch rows of one (channel, frame))
  const float* ic_x;     // [B, C, T, H, W]
  const int64_t* ic_idx; // [B, ic_kt] kept-token ids (sorted), or nullptr: every token, ic_kt == tokens per sample
  int ic_kt;             // GEMM rows per sample
  int ic_nw, ic_nh;      // tokens per frame row / frame column
  int ic_tub, ic_c, ic_t;// tubelet depth, channels, frames
  avj_epilogue ep;
};

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// TMA store of one 32-row x 32-column bf16 block (SWIZZLE_64B staging tile) at {column c0, row c1}; rows and
// columns past the tensor bounds are clipped by the hardware.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N> __device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void l2_prefetch(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ float4 ld_f4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void f4_add(float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }

// shared-memory matrix descriptor, SWIZZLE_128B, sm_100 version field = 1
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo16, uint32_t sbo16) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(lbo16 & 0x3FFF) << 16) | ((uint64_t)(sbo16 & 0x3FFF) << 32) |
         (1ull << 46) | (2ull << 61);
}

// ------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------
// Epilogue specialisations (compile-time, so the hot loop carries no dead registers or branches):
//   EPI_PLAIN      thread == row, acc (+bias) -> bf16/fp32                       (qkv fprop, dgrads)
//   EPI_GELU       + GELU, optional bf16 pre-activation stash                    (fc1 fprop)
//   EPI_DACT       acc * GELU'(saved pre-activation)                             (fc2 dgrad)
//   EPI_TRANSPOSED fp32 C with bias / fp32 residual / gathered positional rows / C += / split-K atomics;
//                  32x32 blocks go through a swizzled smem transpose so that every global access is a
//                  coalesced 4 rows x 128 B                                      (proj, fc2, wgrads, embeds)
//   EPI_GENERIC    every epilogue field at run time (rare combinations)
enum { EPI_PLAIN = 0, EPI_GELU = 1, EPI_DACT = 2, EPI_TRANSPOSED = 3, EPI_GENERIC = 4 };

template <int R> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_f4(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// One accumulator tile (this warp's 32 rows x [c_lo, c_hi) columns) through the fused epilogue.
// `wait_full` blocks until the tile's MMAs have retired, `release` signals that the TMEM buffer is drained.
template <int EPI, int NBUF, bool TS, class WaitFull, class Release>
__device__ __forceinline__ void epilogue_tile(const UmmaParams& p, const CUtensorMap* tma_c, const CUtensorMap* tma_p,
                                              int row_base, int n_blk, uint32_t taddr, uint32_t stg,
                                              int lane, int c_lo, int c_hi, WaitFull wait_full, Release release) {
  const avj_epilogue& ep = p.ep;
  const int cols_half = c_hi - c_lo;
  const int sub = lane >> 3, c4 = lane & 7;        // transposed phase: row i*4+sub, float4 column c4
  uint32_t raw[32];

  if (EPI == EPI_PLAIN || EPI == EPI_GELU || EPI == EPI_DACT) {
    // ---------------- direct: thread == row, 32 consecutive columns per step ----------------
    const bool has_bias = (EPI != EPI_DACT) && ep.bias != nullptr;
    constexpr bool ts = TS;
    // direct stores: the warp's bias segment goes to shared memory once per tile (broadcast reads afterwards);
    // TMA stores own the staging buffer, so there the bias comes straight from L1 (one address per warp)
    if (has_bias && !ts) {
      __syncwarp();                               // every lane is done with the previous tile's bias
      if (lane * 4 < cols_half) sts_f4(stg + lane * 16, ld_f4(ep.bias + n_blk * p.block_n + c_lo + lane * 4));
      __syncwarp();
    }
    const int64_t row = row_base + lane;
    const bool row_ok = row < p.M;
    const int64_t prow = row_ok ? map_row(ep.out_map, row) : 0;
    // TMA-store staging: NBUF tiles of 32 rows x 64 B (SWIZZLE_64B: 16-byte chunk j of row r sits at j ^ ((r >> 1) & 3))
    const uint32_t st_row = stg + lane * 64;
    const int st_swz = (lane >> 1) & 3;
    int st_buf = 0;
    auto stage_store = [&](const CUtensorMap* map, const float (&v)[32], int n0) {
      if (lane == 0) tma_store_wait_read<NBUF - 1>();          // the tile written NBUF stores ago has been read
      __syncwarp();
      const uint32_t dst = st_row + st_buf * 2048;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + ((j ^ st_swz) << 4)),
                     "r"(pack_bf16x2(v[8 * j], v[8 * j + 1])), "r"(pack_bf16x2(v[8 * j + 2], v[8 * j + 3])),
                     "r"(pack_bf16x2(v[8 * j + 4], v[8 * j + 5])), "r"(pack_bf16x2(v[8 * j + 6], v[8 * j + 7])) : "memory");
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) tma_store_2d(map, stg + st_buf * 2048, n0, row_base);
      st_buf = (st_buf + 1) % NBUF;
    };
    // GELU' operand (the saved pre-activation, thread == row): the loads of step c + 32 are issued before the
    // math of step c and those of the first step before the tile's MMAs have even retired, so their L2 / HBM
    // latency never sits on the epilogue's critical path
    uint4 aux_nxt[4];
    auto load_aux = [&](int c) {
      const uint4* ap = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(ep.dact_aux) + row * (int64_t)p.N +
                                                       n_blk * p.block_n + c);
#pragma unroll
      for (int j = 0; j < 4; ++j) aux_nxt[j] = row_ok ? __ldg(ap + j) : make_uint4(0u, 0u, 0u, 0u);
    };
    if (EPI == EPI_DACT) load_aux(c_lo);
    wait_full();
    tmem_ld32_issue(taddr + c_lo, raw);
    for (int c = c_lo; c < c_hi; c += 32) {
      const int n0 = n_blk * p.block_n + c;
      uint4 aux[4];
      if (EPI == EPI_DACT) {
#pragma unroll
        for (int j = 0; j < 4; ++j) aux[j] = aux_nxt[j];
        if (c + 32 < c_hi) load_aux(c + 32);
      }
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
      if (c + 32 < c_hi) {
        tmem_ld32_issue(taddr + c + 32, raw);
      } else {
        release();
      }
      if (has_bias) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b = ts ? __ldg(reinterpret_cast<const float4*>(ep.bias + n0) + j) : lds_f4(stg + (c - c_lo + j * 4) * 4);
          v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
        }
      }
      if (!ts && !row_ok) continue;
      if (EPI == EPI_GELU) {
        if (ep.pre_out) {
          if (ts) {
            stage_store(tma_p, v, n0);
          } else {
            uint4* pp = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(ep.pre_out) + row * (int64_t)p.N + n0);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              pp[j] = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                 pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
          }
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = gelu_fwd<true>(v[i]);
      }
      if (EPI == EPI_DACT) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t w[4] = {aux[j].x, aux[j].y, aux[j].z, aux[j].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float x0, x1;
            unpack_bf16x2(w[e], x0, x1);
            v[8 * j + 2 * e] *= gelu_bwd<true>(x0);
            v[8 * j + 2 * e + 1] *= gelu_bwd<true>(x1);
          }
        }
      }
      if (ts) {
        stage_store(tma_c, v, n0);
      } else if (ep.out_dtype == AVJ_F32) {
        float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.C) + prow * (int64_t)p.ldc + n0);
#pragma unroll
        for (int j = 0; j < 8; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      } else {
        uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.C) + prow * (int64_t)p.ldc + n0);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          op[j] = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                             pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
      }
    }
  } else if (EPI == EPI_GENERIC) {
    // ---------------- generic direct path: every epilogue field evaluated at run time ----------------
    wait_full();
    tmem_ld32_issue(taddr + c_lo, raw);
    const int64_t row = row_base + lane;
    const bool row_ok = row < p.M;
    for (int c = c_lo; c < c_hi; c += 32) {
      const int n0 = n_blk * p.block_n + c;
      float add[32];
      if (p.split_k == 1 && row_ok) epilogue_prefetch<32>(ep, p.C, p.ldc, p.N, row, n0, add);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
      if (c + 32 < c_hi) {
        tmem_ld32_issue(taddr + c + 32, raw);
      } else {
        release();
      }
      if (row_ok) {
        if (p.split_k > 1) {
          float* out = reinterpret_cast<float*>(p.C) + map_row(ep.out_map, row) * (int64_t)p.ldc + n0;
#pragma unroll
          for (int i = 0; i < 32; i += 4) atomicAdd(reinterpret_cast<float4*>(out + i), make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
        } else {
          if (p.gelu_exact) epilogue_apply_store<bf16, 32, false>(ep, p.C, p.ldc, p.N, row, n0, v, add);
          else epilogue_apply_store<bf16, 32, true>(ep, p.C, p.ldc, p.N, row, n0, v, add);
        }
      }
    }
  } else {
    // ---------------- transposed: coalesced 4 rows x 128 B per warp instruction, fp32 C ----------------
    int prow[8];                                   // physical C row of my 8 rows, -1 = past M
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = row_base + i * 4 + sub;
      prow[i] = r < p.M ? (int)map_row(ep.out_map, r) : -1;
    }
    wait_full();
    tmem_ld32_issue(taddr + c_lo, raw);
    for (int c = c_lo; c < c_hi; c += 32) {
      const int n = n_blk * p.block_n + c + c4 * 4;
      // ---- addends that do not depend on the accumulator: ALL loads are issued back to back (rows past
      // M are clamped to a valid row and discarded) so one chunk costs one memory round trip, and they
      // are in flight while the TMEM load completes
      float4 add[8], add2[8];
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      const bool two = ep.accumulate && (ep.residual || ep.pos);
      if (p.split_k == 1) {
        if (ep.bias) b4 = ld_f4(ep.bias + n);
#pragma unroll
        for (int i = 0; i < 8; ++i) { add[i] = make_float4(0.f, 0.f, 0.f, 0.f); add2[i] = add[i]; }
        if (ep.residual) {
#pragma unroll
          for (int i = 0; i < 8; ++i) add[i] = ld_f4(ep.residual + (int64_t)max(prow[i], 0) * p.ldc + n);
        } else if (ep.pos) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = min(row_base + i * 4 + sub, p.M - 1);
            const int64_t pr = ep.pos_idx ? ep.pos_idx[r] : (int64_t)(r % ep.pos_rows);
            add[i] = ld_f4(ep.pos + pr * (int64_t)p.N + n);
          }
        }
        if (ep.accumulate) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 t = ld_f4(reinterpret_cast<const float*>(p.C) + (int64_t)max(prow[i], 0) * p.ldc + n);
            if (two) add2[i] = t; else add[i] = t;
          }
        }
      }
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t dst = stg + lane * 128 + ((j ^ (lane & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(raw[4 * j]), "r"(raw[4 * j + 1]),
                     "r"(raw[4 * j + 2]), "r"(raw[4 * j + 3]) : "memory");
      }
      if (c + 32 < c_hi) {
        tmem_ld32_issue(taddr + c + 32, raw);      // next block is in flight while this one is stored
      } else {
        release();                                  // accumulator drained: the MMA warp may reuse it
      }
      __syncwarp();
      float4 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rl = i * 4 + sub;
        v[i] = lds_f4(stg + rl * 128 + ((c4 ^ (rl & 7)) << 4));
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (prow[i] < 0) continue;
        float* out = reinterpret_cast<float*>(p.C) + (int64_t)prow[i] * p.ldc + n;
        if (p.split_k > 1) {
          atomicAdd(reinterpret_cast<float4*>(out), v[i]);
        } else {
          f4_add(v[i], b4);
          f4_add(v[i], add[i]);
          if (two) f4_add(v[i], add2[i]);
          *reinterpret_cast<float4*>(out) = v[i];
        }
      }
    }
  }
}

// EW = number of epilogue warps: 8 (two per TMEM lane quarter, half of the columns each) or 16 (four per
// quarter, a quarter of the columns each) for the thread==row epilogues, which are latency- rather than
// issue-bound and double their throughput with twice the warps in flight.
// GATHER: 0 = both operands by TMA; 1 = im2col-free patch embedding (A rows gathered out of the clip, tf32); 2 = its weight
// gradient (B rows = the same patch values, gathered and converted to bf16 by the producer warps, MN-major).
template <int EPI, int EW, bool TS, int GATHER = 0>
__global__ void __launch_bounds__(128 + 32 * EW, 1)
gemm_umma_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                 const __grid_constant__ CUtensorMap tma_c, const __grid_constant__ CUtensorMap tma_p, const UmmaParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t smem_a = base;
  const uint32_t smem_b = base + UG_STAGES * UG_A_STAGE_BYTES;
  const uint32_t epi_stage = smem_b + UG_STAGES * UG_B_STAGE_BYTES;   // [UG_EPI_WARPS] x 4 KB, 1024-byte aligned
  const uint32_t bars = epi_stage + UG_EPI_WARPS * UG_EPI_STAGE_BYTES;
  const uint32_t full_bar = bars;                       // [UG_STAGES]
  const uint32_t empty_bar = bars + 8 * UG_STAGES;      // [UG_STAGES]
  const uint32_t tfull_bar = bars + 16 * UG_STAGES;     // [2]
  const uint32_t tempty_bar = tfull_bar + 16;           // [2]
  const uint32_t tmem_slot = tempty_bar + 16;           // u32
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  constexpr bool IM2COL = GATHER == 1, WGRAD = GATHER == 2;
  const uint32_t tok_tab = bars + 256;                  // WGRAD: [2][64] s32 element offsets of the k-block's tokens (extra 512 bytes)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_b) : "memory");
    for (int i = 0; i < UG_STAGES; ++i) { mbar_init(full_bar + 8 * i, GATHER ? 97 : 1); mbar_init(empty_bar + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(tfull_bar + 8 * i, 1); mbar_init(tempty_bar + 8 * i, 32 * EW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();                 // barrier init, TMEM allocation and descriptor prefetch overlapped the previous kernel

  const int n_units = p.tiles_m * p.tiles_n * p.split_k;

  if (warp < 4) {
    // warpgroup 0 hands its registers to the epilogue warpgroups (the kernel is compiled for 168); the gather loop of
    // the im2col producer needs more than 40, so that instantiation keeps the launch allocation everywhere
    if constexpr (!GATHER) reg_dec<40>();
    if (WGRAD && warp != 1) {
      // ================= producer, patch-embedding weight gradient (warps 0, 2, 3: 96 threads) =================
      // dW[D, kd] += dY^T[D, tokens] . X[tokens, kd]: the contraction runs over the kept tokens, 64 per k-block.  A = dY (bf16,
      // row-major [tokens, D]) arrives by TMA as an MN-major tile like every weight gradient; B = the patch values X is
      // never materialised: each producer thread reads 8 consecutive floats of a token's patch row out of the clip,
      // converts them to bf16 and stores the 16-byte chunk into the MN-major SWIZZLE_128B tile ([token][64 kd values] per
      // 8 KB chunk of 64 columns).  Token element offsets of the k-block come from a small shared-memory table.
      const int pt = warp == 0 ? lane : (warp - 1) * 32 + lane;      // 0 .. 95
      const int HW = p.ic_nh * 16 * p.ic_nw * 16, Wd = p.ic_nw * 16;
      const int cpr = p.block_n / 8;                                  // 16-byte chunks per token row of the stage
      uint32_t stage = 0, phase = 0;
      int it = 0;
      for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
        const int tile = u / p.split_k, ks = u % p.split_k;
        const int m_blk = tile / p.tiles_n, n_blk = tile % p.tiles_n;
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(p.k_blocks, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const uint32_t tab = tok_tab + (it & 1) * 256;
          if (pt < 64) {
            const int t = kb * 64 + pt;
            int off = -1;                                               // past the last token: a zero row
            if (t < p.K) {
              const int b = t / p.ic_kt, jj = t % p.ic_kt;
              const int tok = p.ic_idx ? (int)__ldg(p.ic_idx + (int64_t)b * p.ic_kt + jj) : jj;
              const int wt = tok % p.ic_nw, hrow = tok / p.ic_nw;
              off = ((b * p.ic_c) * p.ic_t + (hrow / p.ic_nh) * p.ic_tub) * HW + (hrow % p.ic_nh) * 16 * Wd + wt * 16;
            }
            asm volatile("st.shared.s32 [%0], %1;" ::"r"(tab + 4 * pt), "r"(off) : "memory");
          }
          asm volatile("bar.sync 1, 96;" ::: "memory");
          const uint32_t fb = full_bar + 8 * stage;
          const uint32_t sa = smem_a + stage * UG_A_STAGE_BYTES;
          const uint32_t sb = smem_b + stage * UG_B_STAGE_BYTES;
          mbar_wait(empty_bar + 8 * stage, phase ^ 1);
          if (pt == 0) {
            mbar_expect_tx(fb, UG_A_STAGE_BYTES);
            tma_load_2d(sa, &tma_a, fb, m_blk * UG_BM, kb * UG_BK);
            tma_load_2d(sa + 8192, &tma_a, fb, m_blk * UG_BM + 64, kb * UG_BK);
          }
          // eight chunks per round: all sixteen 16-byte loads are in flight before the first conversion (the shared-memory
          // stores below are ordered asm statements, loads scheduled between them would each expose a full L2 round trip)
          const int* tabp = reinterpret_cast<const int*>(smem_raw + (tab - raw));
          const int total = 64 * cpr;
          const int cshift = 31 - __clz(cpr);                        // block_n is 64, 128 or 256: cpr is a power of two
          for (int q0 = pt; q0 < total; q0 += 8 * 96) {
            float4 f[8][2];
            uint32_t dst[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int q = q0 + 96 * e;
              f[e][0] = make_float4(0.f, 0.f, 0.f, 0.f);
              f[e][1] = f[e][0];
              dst[e] = 0xffffffffu;
              if (q < total) {
                const int r = q >> cshift, j = q & (cpr - 1);          // token row of the k-block, 16-byte chunk of its row
                const int kk = n_blk * p.block_n + j * 8;              // first of 8 consecutive kd indices: (plane, dh, dw0)
                const int off = tabp[r];
                dst[e] = sb + (uint32_t)(j >> 3) * 8192u + (uint32_t)r * 128u + (uint32_t)(((j & 7) ^ (r & 7)) << 4);
                if (off >= 0) {
                  const int plane = kk >> 8;
                  const float4* src = reinterpret_cast<const float4*>(
                      p.ic_x + off + ((plane / p.ic_tub) * p.ic_t + plane % p.ic_tub) * HW + ((kk >> 4) & 15) * Wd + (kk & 15));
                  f[e][0] = __ldg(src);
                  f[e][1] = __ldg(src + 1);
                }
              }
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              if (dst[e] == 0xffffffffu) continue;
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst[e]), "r"(pack_bf16x2(f[e][0].x, f[e][0].y)),
                           "r"(pack_bf16x2(f[e][0].z, f[e][0].w)), "r"(pack_bf16x2(f[e][1].x, f[e][1].y)),
                           "r"(pack_bf16x2(f[e][1].z, f[e][1].w)) : "memory");
            }
          }
          fence_proxy_async_smem();
          mbar_arrive(fb);
          if (++stage == UG_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    } else if (IM2COL && warp != 1) {
      // ================= producer, im2col-free patch rows (warps 0, 2, 3: 96 threads) =================
      // Row r of the A tile is one token; its k-block kb is 128 bytes: rows dh0, dh0 + 1 (16 floats each) of the token's
      // 16 x 16 patch in (channel c, frame dt of the tubelet), kb = ((c * tub + dt) * 16 + dh0) / 2 -- exactly the
      // (c, dt, dh, dw) order of the Conv3d / Conv2d weight.  Those are 64-byte runs 4*W bytes apart: as TMA boxes they
      // are far too small (measured: ~65 clk per box, 6 x slower than the tensor pipe needs), so the gather is 16-byte
      // cp.async copies straight into the SWIZZLE_128B K-major tile (thread = one 16-byte chunk column of 11 rows); the
      // weight tile comes by TMA as usual.  A stage is published two commits later: cp.async.wait_group, proxy fence,
      // arrive -- full barrier = 96 thread arrivals + the expect_tx arrival of the weight box.
      const int pt = warp == 0 ? lane : (warp - 1) * 32 + lane;      // 0 .. 95
      const int ch = pt & 7, rg = pt >> 3;                             // chunk of the 128-byte row; rows rg, rg + 12, ...
      const int in_row = (ch >> 2), dw0 = (ch & 3) * 4;                // which of the two patch rows, first float
      const uint32_t b_box_bytes = (uint32_t)p.block_n * 128u;
      const int HW = p.ic_nh * 16 * p.ic_nw * 16, Wd = p.ic_nw * 16;
      constexpr int LAG = 2;
      uint32_t stage = 0, phase = 0;
      int issued = 0;
      auto publish = [&](int which) { mbar_arrive(full_bar + 8 * (which % UG_STAGES)); };
      for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
        const int tile = u / p.split_k, ks = u % p.split_k;
        const int m_blk = tile / p.tiles_n, n_blk = tile % p.tiles_n;
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(p.k_blocks, kb0 + p.kb_per_split);
        int base[11];                                                  // element offset of (token, plane 0, patch row 0, dw0)
#pragma unroll
        for (int i = 0; i < 11; ++i) {
          const int r = rg + 12 * i;
          int row = m_blk * UG_BM + r;
          row = min(row, p.M - 1);                                     // rows past M re-load the last token (never stored)
          const int b = row / p.ic_kt, jj = row % p.ic_kt;
          const int tok = p.ic_idx ? (int)__ldg(p.ic_idx + (int64_t)b * p.ic_kt + jj) : jj;
          const int wt = tok % p.ic_nw, hrow = tok / p.ic_nw;
          base[i] = ((b * p.ic_c) * p.ic_t + (hrow / p.ic_nh) * p.ic_tub) * HW + (hrow % p.ic_nh) * 16 * Wd + wt * 16 + dw0;
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          const int plane = kb >> 3;                                   // (c, dt): 8 k-blocks of two patch rows each
          const int off = ((plane / p.ic_tub) * p.ic_t + plane % p.ic_tub) * HW + ((kb & 7) * 2 + in_row) * Wd;
          const uint32_t fb = full_bar + 8 * stage;
          const uint32_t sa = smem_a + stage * UG_A_STAGE_BYTES;
          mbar_wait(empty_bar + 8 * stage, phase ^ 1);
          if (pt == 0) {
            mbar_expect_tx(fb, b_box_bytes);
            tma_load_2d(smem_b + stage * UG_B_STAGE_BYTES, &tma_b, fb, kb * 32, n_blk * p.block_n);
          }
#pragma unroll
          for (int i = 0; i < 11; ++i) {
            const int r = rg + 12 * i;
            if (r < UG_BM) {
              const uint32_t dst = sa + (uint32_t)r * 128u + (uint32_t)((ch ^ (r & 7)) << 4);
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(p.ic_x + base[i] + off) : "memory");
            }
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
          if (++issued > LAG) {
            asm volatile("cp.async.wait_group %0;" ::"n"(LAG) : "memory");
            fence_proxy_async_smem();
            publish(issued - 1 - LAG);
          }
          if (++stage == UG_STAGES) { stage = 0; phase ^= 1; }
        }
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      fence_proxy_async_smem();
      for (int j = max(0, issued - LAG); j < issued; ++j) publish(j);
    } else if (warp == 0) {
      // ================= TMA producer =================
      if (lane == 0) {
        const uint32_t b_box_bytes = (uint32_t)p.block_n * UG_BK * 2;
        uint32_t stage = 0, phase = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
          const int tile = u / p.split_k, ks = u % p.split_k;
          const int m_blk = tile / p.tiles_n, n_blk = tile % p.tiles_n;
          const int kb0 = ks * p.kb_per_split;
          const int kb1 = min(p.k_blocks, kb0 + p.kb_per_split);
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(empty_bar + 8 * stage, phase ^ 1);
            const uint32_t fb = full_bar + 8 * stage;
            mbar_expect_tx(fb, UG_A_STAGE_BYTES + b_box_bytes);
            const uint32_t sa = smem_a + stage * UG_A_STAGE_BYTES;
            const uint32_t sb = smem_b + stage * UG_B_STAGE_BYTES;
            if (!p.a_mn_major) {
              tma_load_2d(sa, &tma_a, fb, kb * UG_BK, m_blk * UG_BM);
            } else {
              tma_load_2d(sa, &tma_a, fb, m_blk * UG_BM, kb * UG_BK);
              tma_load_2d(sa + 8192, &tma_a, fb, m_blk * UG_BM + 64, kb * UG_BK);
            }
            if (!p.b_mn_major) {
              tma_load_2d(sb, &tma_b, fb, kb * UG_BK, n_blk * p.block_n);
            } else {
              for (int j = 0; j < p.block_n / 64; ++j)
                tma_load_2d(sb + j * 8192, &tma_b, fb, n_blk * p.block_n + j * 64, kb * UG_BK);
            }
            if (++stage == UG_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    } else if (warp == 1) {
      // ================= MMA issuer =================
      if (lane == 0) {
        // operand format field: 1 = bf16 (kind::f16), 2 = tf32 (kind::tf32); fp32 accumulator
        const uint32_t fmt = IM2COL ? 2u : 1u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)p.a_mn_major << 15) |
                               ((uint32_t)p.b_mn_major << 16) | ((uint32_t)(p.block_n >> 3) << 17) |
                               ((uint32_t)(UG_BM >> 4) << 24);
        uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
          const int ks = u % p.split_k;
          const int kb0 = ks * p.kb_per_split;
          const int kb1 = min(p.k_blocks, kb0 + p.kb_per_split);
          mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + acc * UG_MAX_BN;
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(full_bar + 8 * stage, phase);
            tc_fence_after();
            const uint32_t sa = smem_a + stage * UG_A_STAGE_BYTES;
            const uint32_t sb = smem_b + stage * UG_B_STAGE_BYTES;
            const uint64_t adesc0 = p.a_mn_major ? make_desc(sa, p.mn_lbo, p.mn_sbo) : make_desc(sa, 1, 64);
            const uint64_t bdesc0 = p.b_mn_major ? make_desc(sb, p.mn_lbo, p.mn_sbo) : make_desc(sb, 1, 64);
            const uint32_t a_adv = p.a_mn_major ? p.mn_kadv : 2u;   // 16-byte units per UMMA_K=16
            const uint32_t b_adv = p.b_mn_major ? p.mn_kadv : 2u;
#pragma unroll
            for (int k = 0; k < UG_BK / 16; ++k) {             // tf32: 4 steps of K = 8 elements, also 32 bytes each
              if (IM2COL) tc_mma_tf32(tmem_d, adesc0 + (uint64_t)(k * a_adv), bdesc0 + (uint64_t)(k * b_adv), idesc,
                                      (kb > kb0 || k > 0) ? 1u : 0u);
              else tc_mma_bf16(tmem_d, adesc0 + (uint64_t)(k * a_adv), bdesc0 + (uint64_t)(k * b_adv), idesc,
                               (kb > kb0 || k > 0) ? 1u : 0u);
            }
            tc_commit(empty_bar + 8 * stage);          // frees the smem stage when these MMAs retire
            if (++stage == UG_STAGES) { stage = 0; phase ^= 1; }
          }
          tc_commit(tfull_bar + 8 * acc);              // accumulator complete -> epilogue
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  } else {
    // ================= epilogue warpgroups (EW warps) =================
    // EW/4 warps per TMEM lane quarter, each taking BN/(EW/4) of the tile's columns, 32 columns at a time.
    if constexpr (!GATHER) reg_inc<EW == 8 ? 232 : 104>();   // pool: 640 x 96 at launch; 128 x (96 - 40) freed >= 512 x (104 - 96) claimed (112 would deadlock)
    const int q = warp & 3;                          // TMEM lane quarter this warp may touch
    const int ew = warp - 4;
    const int part = ew >> 2;
    const int cols_part = p.block_n / (EW / 4);
    const int c_lo = part * cols_part, c_hi = c_lo + cols_part;
    const uint32_t stg = epi_stage + ew * (UG_EPI_WARPS * UG_EPI_STAGE_BYTES / EW);
    uint32_t acc = 0, acc_phase = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
      const int tile = u / p.split_k;
      const int m_blk = tile / p.tiles_n, n_blk = tile % p.tiles_n;
      const int row_base = m_blk * UG_BM + q * 32;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * UG_MAX_BN;
      epilogue_tile<EPI, 16 / EW, TS>(p, &tma_c, &tma_p, row_base, n_blk, taddr, stg, lane, c_lo, c_hi,
                         [&] { mbar_wait(tfull_bar + 8 * acc, acc_phase); tc_fence_after(); },
                         [&] { tc_fence_before(); mbar_arrive(tempty_bar + 8 * acc); });
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (TS && lane == 0) tma_store_wait_read<0>();   // staging smem must outlive the bulk stores' reads
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// 2-CTA kernel: one 256 x BN output tile per CTA PAIR (tcgen05.mma.cta_group::2)
// ------------------------------------------------------------------------------------------
// The two CTAs of a cluster sit on the two SMs of a TPC.  Each loads ITS 128 rows of A and ITS half of the
// B tile (BN/2 rows or columns); the leader's single MMA thread issues M = 256 instructions that read both
// CTAs' shared memory and write both CTAs' TMEM.  Per SM that is 8 KB instead of 12 KB of operand traffic
// per 128x256x16 MACs -- shared-memory bandwidth (TMA writes + MMA reads) is what caps the 1-CTA kernel at
// ~77 % of the tensor pipe -- and the smaller stages make room for a 6-deep ring.
//   TMA      : both CTAs, every load completes on the LEADER's full barrier (expect_tx covers both halves)
//   MMA      : leader only; tcgen05.commit multicasts to both CTAs' empty / tmem-full barriers
//   epilogue : both CTAs on their own 128 accumulator rows; one lane per warp arrives (remotely for the
//              peer) on the leader's tmem-empty barrier
#define UG2_STAGES 6
#define UG2_STAGE_BYTES (UG_BM * UG_BK * 2)                       // 16 KB for A and 16 KB for the B half
#define UG2_SMEM_BYTES (UG2_STAGES * 2 * UG2_STAGE_BYTES + 1024 + 256 + UG_EPI_WARPS * UG_EPI_STAGE_BYTES)

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}

template <int EPI, int EW, bool TS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128 + 32 * EW, 1)
gemm_umma2_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                  const __grid_constant__ CUtensorMap tma_c, const __grid_constant__ CUtensorMap tma_p, const UmmaParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t smem_a = base;
  const uint32_t smem_b = base + UG2_STAGES * UG2_STAGE_BYTES;
  const uint32_t epi_stage = smem_b + UG2_STAGES * UG2_STAGE_BYTES;   // 1024-byte aligned
  const uint32_t bars = epi_stage + UG_EPI_WARPS * UG_EPI_STAGE_BYTES;
  const uint32_t full_bar = bars;                       // [UG2_STAGES]  (used in the leader only)
  const uint32_t empty_bar = bars + 8 * UG2_STAGES;     // [UG2_STAGES]
  const uint32_t tfull_bar = bars + 16 * UG2_STAGES;    // [2]
  const uint32_t tempty_bar = tfull_bar + 16;           // [2]           (used in the leader only)
  const uint32_t tmem_slot = tempty_bar + 16;           // u32
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  pdl_trigger();

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_b) : "memory");
    for (int i = 0; i < UG2_STAGES; ++i) { mbar_init(full_bar + 8 * i, 1); mbar_init(empty_bar + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(tfull_bar + 8 * i, 1); mbar_init(tempty_bar + 8 * i, 2 * EW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();                                   // barrier inits + TMEM allocation visible pair-wide
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();

  const int n_units = p.tiles_m * p.tiles_n * p.split_k;   // tiles_m counts 256-row tiles here
  const int half_n = p.block_n / 2;

  if (warp < 4) {
    reg_dec<40>();
    if (warp == 0) {
      // ================= TMA producer (both CTAs) =================
      if (lane == 0) {
        const uint32_t stage_tx = 2u * (UG2_STAGE_BYTES + (uint32_t)half_n * UG_BK * 2);   // both CTAs' bytes
        uint32_t stage = 0, phase = 0;
        for (int u = cluster_id; u < n_units; u += n_clusters) {
          const int tile = u / p.split_k, ks = u % p.split_k;
          const int m_blk = tile / p.tiles_n, n_blk = tile % p.tiles_n;
          const int m0 = m_blk * 2 * UG_BM + (int)rank * UG_BM;           // my 128 rows of the 256-row tile
          const int n0 = n_blk * p.block_n + (int)rank * half_n;          // my half of the B tile
          const int kb0 = ks * p.kb_per_split;
          const int kb1 = min(p.k_blocks, kb0 + p.kb_per_split);
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(empty_bar + 8 * stage, phase ^ 1);
            const uint32_t fb = mapa_u32(full_bar + 8 * stage, 0);        // the leader's barrier
            if (leader) mbar_expect_tx(full_bar + 8 * stage, stage_tx);
            const uint32_t sa = smem_a + stage * UG2_STAGE_BYTES;
            const uint32_t sb = smem_b + stage * UG2_STAGE_BYTES;
            if (!p.a_mn_major) {
              tma_load_2d_2sm(sa, &tma_a, fb, kb * UG_BK, m0);
            } else {
              tma_load_2d_2sm(sa, &tma_a, fb, m0, kb * UG_BK);
              tma_load_2d_2sm(sa + 8192, &tma_a, fb, m0 + 64, kb * UG_BK);
            }
            if (!p.b_mn_major) {
              tma_load_2d_2sm(sb, &tma_b, fb, kb * UG_BK, n0);
            } else {
              for (int j = 0; j < half_n / 64; ++j)
                tma_load_2d_2sm(sb + j * 8192, &tma_b, fb, n0 + j * 64, kb * UG_BK);
            }
            if (++stage == UG2_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    } else if (warp == 1 && leader) {
      // ================= MMA issuer (leader CTA only) =================
      if (lane == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)p.a_mn_major << 15) |
                               ((uint32_t)p.b_mn_major << 16) | ((uint32_t)(p.block_n >> 3) << 17) |
                               ((uint32_t)((2 * UG_BM) >> 4) << 24);
        uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
        for (int u = cluster_id; u < n_units; u += n_clusters) {
          const int ks = u % p.split_k;
          const int kb0 = ks * p.kb_per_split;
          const int kb1 = min(p.k_blocks, kb0 + p.kb_per_split);
          mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + acc * UG_MAX_BN;
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(full_bar + 8 * stage, phase);
            tc_fence_after();
            const uint32_t sa = smem_a + stage * UG2_STAGE_BYTES;
            const uint32_t sb = smem_b + stage * UG2_STAGE_BYTES;
            const uint64_t adesc0 = p.a_mn_major ? make_desc(sa, p.mn_lbo, p.mn_sbo) : make_desc(sa, 1, 64);
            const uint64_t bdesc0 = p.b_mn_major ? make_desc(sb, p.mn_lbo, p.mn_sbo) : make_desc(sb, 1, 64);
            const uint32_t a_adv = p.a_mn_major ? p.mn_kadv : 2u;
            const uint32_t b_adv = p.b_mn_major ? p.mn_kadv : 2u;
#pragma unroll
            for (int k = 0; k < UG_BK / 16; ++k) {
              tc_mma_bf16_2sm(tmem_d, adesc0 + (uint64_t)(k * a_adv), bdesc0 + (uint64_t)(k * b_adv), idesc,
                              (kb > kb0 || k > 0) ? 1u : 0u);
            }
            tc_commit_2sm(empty_bar + 8 * stage);        // frees this stage in BOTH CTAs
            if (++stage == UG2_STAGES) { stage = 0; phase ^= 1; }
          }
          tc_commit_2sm(tfull_bar + 8 * acc);            // accumulator complete -> both epilogues
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  } else {
    // ================= epilogue warpgroups (both CTAs, own 128 accumulator rows) =================
    reg_inc<EW == 8 ? 232 : 104>();   // pool: 640 x 96 at launch; 128 x (96 - 40) freed >= 512 x (104 - 96) claimed (112 would deadlock)
    const int q = warp & 3;
    const int ew = warp - 4;
    const int part = ew >> 2;
    const int cols_part = p.block_n / (EW / 4);
    const int c_lo = part * cols_part, c_hi = c_lo + cols_part;
    const uint32_t stg = epi_stage + ew * (UG_EPI_WARPS * UG_EPI_STAGE_BYTES / EW);
    uint32_t acc = 0, acc_phase = 0;
    for (int u = cluster_id; u < n_units; u += n_clusters) {
      const int tile = u / p.split_k;
      const int m_blk = tile / p.tiles_n, n_blk = tile % p.tiles_n;
      const int row_base = m_blk * 2 * UG_BM + (int)rank * UG_BM + q * 32;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * UG_MAX_BN;
      const uint32_t leader_tempty = mapa_u32(tempty_bar + 8 * acc, 0);
      epilogue_tile<EPI, 16 / EW, TS>(p, &tma_c, &tma_p, row_base, n_blk, taddr, stg, lane, c_lo, c_hi,
                         [&] { mbar_wait(tfull_bar + 8 * acc, acc_phase); tc_fence_after(); },
                         [&] { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive_cluster(leader_tempty); });
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (TS && lane == 0) tma_store_wait_read<0>();
  }

  tc_fence_before();
  cluster_sync_all();                                   // the peer's smem / barriers / TMEM stay alive until both are done
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// host: tensor-map cache + launch
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

struct MapKey {
  const void* ptr; uint64_t inner, outer, ld; uint32_t box_inner, box_outer, swz64;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && inner == o.inner && outer == o.outer && ld == o.ld && box_inner == o.box_inner && box_outer == o.box_outer &&
           swz64 == o.swz64;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    h = h * 1000003u ^ k.inner; h = h * 1000003u ^ k.outer; h = h * 1000003u ^ k.ld;
    h = h * 1000003u ^ k.box_inner; h = h * 1000003u ^ k.box_outer; h = h * 1000003u ^ k.swz64;
    return h;
  }
};

// 2-D bf16 tensor map: `inner` contiguous elements per row, `outer` rows, row pitch `ld` elements.
// swz64: SWIZZLE_64B (32-element rows of the epilogue's store tiles) instead of SWIZZLE_128B.
static int get_tensor_map(const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                          uint32_t box_outer, CUtensorMap* out, uint32_t swz64 = 0) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key{ptr, inner, outer, ld, box_inner, box_outer, swz64};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return 0; }
  }
  PFN_encodeTiled enc = get_encode_fn();
  AVJ_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available (driver too old?)");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  AVJ_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d): ptr=%p inner=%llu outer=%llu ld=%llu box=%ux%u", (int)r,
            ptr, (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld, box_inner, box_outer);
  {
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 8192) cache.clear();
    cache[key] = *out;
  }
  return 0;
}

static int pick_block_n(int N) {
  const int cands[4] = {256, 192, 128, 64};
  for (int i = 0; i < 4; ++i) if (N % cands[i] == 0) return cands[i];
  return 0;
}

static uint32_t env_u32(const char* name, uint32_t dflt) {
  const char* s = getenv(name);
  return s ? (uint32_t)strtoul(s, nullptr, 0) : dflt;
}

bool avj_gemm_umma_supported(int dtype, int layout, const void* A, const void* B, int M, int N, int K,
                             int lda, int ldb, const avj_epilogue& ep) {
  if (dtype != AVJ_BF16) return false;
  if (M <= 0 || K <= 0 || pick_block_n(N) == 0) return false;
  if ((lda % 8) || (ldb % 8)) return false;
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) return false;
  if (layout == AVJ_GEMM_TN && (M % 8)) return false;
  if (layout != AVJ_GEMM_NT && layout != AVJ_GEMM_NN && layout != AVJ_GEMM_TN) return false;
  (void)ep;
  return true;
}

int avj_gemm_umma(int layout, const void* A, const void* B, void* C, int M, int N, int K,
                  int lda, int ldb, int ldc, const avj_epilogue& ep, cudaStream_t s) {
  UmmaParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K; p.ldc = ldc; p.C = C; p.ep = ep;
  p.k_blocks = (K + UG_BK - 1) / UG_BK;
  p.a_mn_major = (layout == AVJ_GEMM_TN);
  p.b_mn_major = (layout != AVJ_GEMM_NT);
  // MN-major SWIZZLE_128B canonical layout: 64-element MN atoms 8192 B apart (one TMA box each),
  // 8-row K groups 1024 B apart; one UMMA_K=16 step = 2 K groups = 2048 B.
  static const uint32_t mn_lbo = env_u32("AVJ_UMMA_MN_LBO", 8192 / 16);
  static const uint32_t mn_sbo = env_u32("AVJ_UMMA_MN_SBO", 1024 / 16);
  static const uint32_t mn_kadv = env_u32("AVJ_UMMA_MN_KADV", 2048 / 16);
  p.mn_lbo = mn_lbo; p.mn_sbo = mn_sbo; p.mn_kadv = mn_kadv;

  // ---- 1-CTA (128 x BN tiles) or CTA-pair (256 x BN tiles, cta_group::2) kernel
  //   AVJ_GEMM_2CTA = 0 never | 1 whenever legal | 2 (default) when the problem has at least 256 rows
  static const uint32_t mode2 = env_u32("AVJ_GEMM_2CTA", 2);
  p.block_n = pick_block_n(N);
  bool two = mode2 != 0 && (mode2 == 1 || M >= 2 * UG_BM);
  if (two && p.b_mn_major) {
    // each CTA loads BN/2 columns of an MN-major B tile as whole 64-column TMA boxes
    if (N % 256 == 0) p.block_n = 256;
    else if (N % 128 == 0) p.block_n = 128;
    else two = false;
  }
  const int tile_m = two ? 2 * UG_BM : UG_BM;
  p.tiles_m = (M + tile_m - 1) / tile_m;
  p.tiles_n = N / p.block_n;

  const int sms = avj_num_sms();
  const int workers = two ? sms / 2 : sms;               // CTAs or CTA pairs
  const int tiles = p.tiles_m * p.tiles_n;
  p.split_k = 1;
  const bool pure_accumulate = ep.accumulate && ep.out_dtype == AVJ_F32 && !ep.bias && !ep.residual && !ep.pos &&
                               !ep.act && !ep.dact_aux;
  if (pure_accumulate && tiles < 2 * workers && p.k_blocks >= 8) {
    // split-K factor from a wave model: a unit costs its k-blocks plus a fixed epilogue (~6 k-block times for a
    // 256 x 256 fp32 read-modify-write), the launch takes ceil(units / workers) waves of the slowest unit.  The old
    // rule (2 waves' worth of units) left partial last waves: 18 tiles x 9 splits = 162 units on 74 CTA pairs = 3 waves
    // at 73 % occupancy; the model picks 4 splits = 72 units = 1 full wave.
    static const uint32_t old_rule = env_u32("AVJ_GEMM_SPLITK_OLD", 0);
    int max_split = p.k_blocks / 4;
    if (max_split > 32) max_split = 32;
    if (old_rule) {
      int want = (2 * workers + tiles - 1) / tiles;
      if (want > p.k_blocks / 4) want = p.k_blocks / 4;
      p.split_k = want < 1 ? 1 : want;
    } else {
      double best = -1.0;
      for (int sp = 1; sp <= max_split; ++sp) {
        const int kbs = (p.k_blocks + sp - 1) / sp;
        const int eff = (p.k_blocks + kbs - 1) / kbs;
        const int waves = (tiles * eff + workers - 1) / workers;
        const double t = waves * (kbs + 6.0);
        if (best < 0.0 || t < best * 0.98) { best = t; p.split_k = eff; }
      }
    }
  }
  p.kb_per_split = (p.k_blocks + p.split_k - 1) / p.split_k;
  p.split_k = (p.k_blocks + p.kb_per_split - 1) / p.kb_per_split;   // drop empty splits

  CUtensorMap ma, mb;
  int rc;
  if (!p.a_mn_major) rc = get_tensor_map(A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, UG_BK, UG_BM, &ma);
  else               rc = get_tensor_map(A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, UG_BK, &ma);
  if (rc) return rc;
  const uint32_t b_rows = (uint32_t)(two ? p.block_n / 2 : p.block_n);   // K-major B box: rows of W per CTA
  if (!p.b_mn_major) rc = get_tensor_map(B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, UG_BK, b_rows, &mb);
  else               rc = get_tensor_map(B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, UG_BK, &mb);
  if (rc) return rc;

  // ---- epilogue specialisation
  const bool adds = ep.residual || ep.pos || ep.accumulate;
  int epi;
  static const uint32_t force_generic = env_u32("AVJ_EPI_GENERIC", 0);
  // AVJ_GELU_EXACT=1 (A/B aid): bf16 GEMMs evaluate GELU / GELU' with erff() like the fp32 check mode; those launches
  // take the run-time-dispatched epilogue
  static const uint32_t gelu_exact = env_u32("AVJ_GELU_EXACT", 0);
  p.gelu_exact = (gelu_exact && (ep.act || ep.dact_aux)) ? 1 : 0;
  if (force_generic || p.gelu_exact) epi = EPI_GENERIC;
  else if (ep.out_dtype == AVJ_F32) epi = (ep.act || ep.dact_aux) ? EPI_GENERIC : EPI_TRANSPOSED;
  else if (adds || p.split_k > 1 || (ep.act && ep.dact_aux)) epi = EPI_GENERIC;
  else if (ep.act) epi = EPI_GELU;
  else if (ep.dact_aux) epi = EPI_DACT;
  else epi = EPI_PLAIN;

  // thread==row epilogues with a plain [M, N] bf16 C: 32 x 32 blocks leave through shared memory and TMA bulk
  // stores (full 64-byte row segments instead of 32 scattered 16-byte pieces per warp instruction)
  static const uint32_t use_ts = env_u32("AVJ_GEMM_TMA_STORE", 1);
  CUtensorMap mc, mp;
  memset(&mc, 0, sizeof(mc)); memset(&mp, 0, sizeof(mp));
  p.tma_store = 0;
  if (use_ts && epi <= EPI_DACT && ep.out_dtype == AVJ_BF16 && ep.out_map.rows_per_group == 0 && ep.out_map.row_offset == 0 &&
      ldc % 8 == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0 &&
      (!ep.pre_out || (reinterpret_cast<uintptr_t>(ep.pre_out) & 15) == 0)) {
    rc = get_tensor_map(C, (uint64_t)N, (uint64_t)M, (uint64_t)ldc, 32, 32, &mc, 1);
    if (rc) return rc;
    if (epi == EPI_GELU && ep.pre_out) {
      rc = get_tensor_map(ep.pre_out, (uint64_t)N, (uint64_t)M, (uint64_t)N, 32, 32, &mp, 1);
      if (rc) return rc;
    }
    p.tma_store = 1;
  }

  // thread==row epilogues run 16 epilogue warps when the tile splits into four 32-column-aligned parts
  static const uint32_t ew16 = env_u32("AVJ_GEMM_EW16", 1);
  const bool wide = ew16 && epi <= EPI_DACT && p.block_n % 128 == 0;

  typedef void (*kern_t)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const UmmaParams);
  // [two][wide][epi], then the TMA-store variants of the three thread==row epilogues: [two][wide][epi]
  static const kern_t kerns[2][2][5] = {
      {{gemm_umma_kernel<EPI_PLAIN, 8, false>, gemm_umma_kernel<EPI_GELU, 8, false>, gemm_umma_kernel<EPI_DACT, 8, false>,
        gemm_umma_kernel<EPI_TRANSPOSED, 8, false>, gemm_umma_kernel<EPI_GENERIC, 8, false>},
       {gemm_umma_kernel<EPI_PLAIN, 16, false>, gemm_umma_kernel<EPI_GELU, 16, false>, gemm_umma_kernel<EPI_DACT, 16, false>, nullptr, nullptr}},
      {{gemm_umma2_kernel<EPI_PLAIN, 8, false>, gemm_umma2_kernel<EPI_GELU, 8, false>, gemm_umma2_kernel<EPI_DACT, 8, false>,
        gemm_umma2_kernel<EPI_TRANSPOSED, 8, false>, gemm_umma2_kernel<EPI_GENERIC, 8, false>},
       {gemm_umma2_kernel<EPI_PLAIN, 16, false>, gemm_umma2_kernel<EPI_GELU, 16, false>, gemm_umma2_kernel<EPI_DACT, 16, false>, nullptr, nullptr}}};
  static const kern_t kerns_ts[2][2][3] = {
      {{gemm_umma_kernel<EPI_PLAIN, 8, true>, gemm_umma_kernel<EPI_GELU, 8, true>, gemm_umma_kernel<EPI_DACT, 8, true>},
       {gemm_umma_kernel<EPI_PLAIN, 16, true>, gemm_umma_kernel<EPI_GELU, 16, true>, gemm_umma_kernel<EPI_DACT, 16, true>}},
      {{gemm_umma2_kernel<EPI_PLAIN, 8, true>, gemm_umma2_kernel<EPI_GELU, 8, true>, gemm_umma2_kernel<EPI_DACT, 8, true>},
       {gemm_umma2_kernel<EPI_PLAIN, 16, true>, gemm_umma2_kernel<EPI_GELU, 16, true>, gemm_umma2_kernel<EPI_DACT, 16, true>}}};
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    for (int t = 0; t < 2; ++t)
      for (int w = 0; w < 2; ++w)
        for (int i = 0; i < 5 && attr_err == cudaSuccess; ++i)
          if (kerns[t][w][i]) {
            attr_err = cudaFuncSetAttribute(kerns[t][w][i], cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            t == 0 ? UG_SMEM_BYTES : UG2_SMEM_BYTES);
            if (i < 3 && attr_err == cudaSuccess)
              attr_err = cudaFuncSetAttribute(kerns_ts[t][w][i], cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              t == 0 ? UG_SMEM_BYTES : UG2_SMEM_BYTES);
          }
  });
  AVJ_CHECK(attr_err == cudaSuccess, "cudaFuncSetAttribute(gemm_umma_kernel) failed: %s", cudaGetErrorString(attr_err));

  const int units = tiles * p.split_k;
  const int threads = 128 + 32 * (wide ? 16 : 8);
  if (two) {
    const int clusters = units < workers ? units : workers;
    avj_launch_pdl(p.tma_store ? kerns_ts[1][wide][epi] : kerns[1][wide][epi], dim3(2 * clusters), dim3(threads), UG2_SMEM_BYTES, s, ma, mb, mc, mp, p);
  } else {
    const int grid = units < sms ? units : sms;
    avj_launch_pdl(p.tma_store ? kerns_ts[0][wide][epi] : kerns[0][wide][epi], dim3(grid), dim3(threads), UG_SMEM_BYTES, s, ma, mb, mc, mp, p);
  }
  AVJ_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// im2col-free patch embedding: out[row, :] = W[D, kd] . patch(token(row)) + bias + pos[token]   (fp32, tf32 tensor cores)
// ------------------------------------------------------------------------------------------
// Replaces the Conv3d / Conv2d projections of the reference (src/models/utils/patch_embed.py:85-102) for the tokens a
// mask keeps (or all of them): no [tokens, kd] patch matrix is ever materialised, the producer warps of
// gemm_umma_kernel<EPI_TRANSPOSED, 8, false, 1> gather patch rows straight out of the clip into the operand tile.
static int get_tensor_map_f32_2d(const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner, uint32_t box_outer,
                                 CUtensorMap* out) {
  PFN_encodeTiled enc = get_encode_fn();
  AVJ_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available (driver too old?)");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 4};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  AVJ_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(fp32 weight) failed (%d)", (int)r);
  return 0;
}

bool avj_patch_embed_umma_supported(int patch, int H, int W, int T, int tub, int D) {
  return patch == 16 && H % 16 == 0 && W % 16 == 0 && T % tub == 0 && tub >= 1 && pick_block_n(D) != 0;
}

int avj_patch_embed_umma(const float* x, const int64_t* idx, const float* w, float* out, int B, int C, int T, int H, int W, int tub,
                         int patch, int Kt, int D, int ldc, const avj_epilogue& ep, cudaStream_t s) {
  AVJ_CHECK(avj_patch_embed_umma_supported(patch, H, W, T, tub, D), "avj_patch_embed: unsupported geometry (patch %d, %dx%d, D %d)", patch, H, W, D);
  AVJ_CHECK(ep.out_dtype == AVJ_F32 && !ep.act && !ep.dact_aux && !ep.accumulate && !ep.residual, "avj_patch_embed: fp32 output with bias / positional rows only");
  AVJ_CHECK(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w)) & 15) == 0, "avj_patch_embed: x and w must be 16-byte aligned");
  const int n_w = W / 16, n_h = H / 16, n_full = (T / tub) * n_h * n_w;
  const int kd = C * tub * 256;
  AVJ_CHECK(idx != nullptr || Kt == n_full, "avj_patch_embed: without a token list every token is embedded (Kt must be %d)", n_full);
  const int M = B * Kt;
  if (M == 0) return 0;
  UmmaParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = D; p.K = kd; p.ldc = ldc; p.C = out; p.ep = ep;
  p.k_blocks = kd / 32;
  p.block_n = pick_block_n(D);
  p.tiles_m = (M + UG_BM - 1) / UG_BM;
  p.tiles_n = D / p.block_n;
  p.split_k = 1;
  p.kb_per_split = p.k_blocks;
  p.ic_x = x; p.ic_idx = idx; p.ic_kt = Kt; p.ic_nw = n_w; p.ic_nh = n_h; p.ic_tub = tub; p.ic_c = C; p.ic_t = T;
  AVJ_CHECK((int64_t)B * C * T * H * W < (int64_t)1 << 31, "avj_patch_embed: input larger than 2^31 elements");

  CUtensorMap ma, mb, mc, mp;
  memset(&ma, 0, sizeof(ma)); memset(&mc, 0, sizeof(mc)); memset(&mp, 0, sizeof(mp));
  int rc = get_tensor_map_f32_2d(w, (uint64_t)kd, (uint64_t)D, (uint64_t)kd, 32, (uint32_t)p.block_n, &mb);
  if (rc) return rc;

  auto kern = gemm_umma_kernel<EPI_TRANSPOSED, 8, false, 1>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [&] { attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, UG_SMEM_BYTES); });
  AVJ_CHECK(attr_err == cudaSuccess, "cudaFuncSetAttribute(patch-embed kernel) failed: %s", cudaGetErrorString(attr_err));
  const int units = p.tiles_m * p.tiles_n;
  const int sms = avj_num_sms();
  avj_launch_pdl(kern, dim3(units < sms ? units : sms), dim3(128 + 32 * 8), UG_SMEM_BYTES, s, ma, mb, mc, mp, p);
  AVJ_LAUNCH_CHECK();
  return 0;
}

// Weight gradient of the patch embedding without a patch matrix: gw[D, kd] += dy^T[D, tokens] . patches[tokens, kd], the
// second operand gathered (and converted to bf16) out of the clip by the producer warps of
// gemm_umma_kernel<EPI_TRANSPOSED, 8, false, 2>.  dy: bf16 [B * Kt, D] row-major.  Split along the tokens over CTAs like every
// weight gradient (fp32 atomics into gw).
int avj_patch_embed_wgrad_umma(const float* x, const int64_t* idx, const void* dy, float* gw, int B, int C, int T, int H, int W, int tub,
                               int patch, int Kt, int D, cudaStream_t s) {
  AVJ_CHECK(avj_patch_embed_umma_supported(patch, H, W, T, tub, D) && D % 8 == 0, "avj_patch_embed_wgrad: unsupported geometry");
  AVJ_CHECK(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(gw)) & 15) == 0,
            "avj_patch_embed_wgrad: x, dy and gw must be 16-byte aligned");
  const int n_w = W / 16, n_h = H / 16, n_full = (T / tub) * n_h * n_w;
  const int kd = C * tub * 256;
  AVJ_CHECK(idx != nullptr || Kt == n_full, "avj_patch_embed_wgrad: without a token list every token contributes (Kt must be %d)", n_full);
  AVJ_CHECK((int64_t)B * C * T * H * W < (int64_t)1 << 31, "avj_patch_embed_wgrad: input larger than 2^31 elements");
  const int tokens = B * Kt;
  if (tokens == 0) return 0;
  UmmaParams p;
  memset(&p, 0, sizeof(p));
  p.M = D; p.N = kd; p.K = tokens; p.ldc = kd; p.C = gw;
  p.ep.accumulate = 1; p.ep.out_dtype = AVJ_F32;
  p.k_blocks = (tokens + UG_BK - 1) / UG_BK;
  p.a_mn_major = 1; p.b_mn_major = 1;
  p.mn_lbo = env_u32("AVJ_UMMA_MN_LBO", 8192 / 16); p.mn_sbo = env_u32("AVJ_UMMA_MN_SBO", 1024 / 16); p.mn_kadv = env_u32("AVJ_UMMA_MN_KADV", 2048 / 16);
  p.block_n = kd % 256 == 0 ? 256 : (kd % 128 == 0 ? 128 : 64);
  p.tiles_m = (D + UG_BM - 1) / UG_BM;
  p.tiles_n = kd / p.block_n;
  const int sms = avj_num_sms();
  const int tiles = p.tiles_m * p.tiles_n;
  int split = sms / tiles;                                  // one wave of (tile, token range) units
  if (split > p.k_blocks) split = p.k_blocks;
  if (split < 1) split = 1;
  p.kb_per_split = (p.k_blocks + split - 1) / split;
  p.split_k = (p.k_blocks + p.kb_per_split - 1) / p.kb_per_split;
  p.ic_x = x; p.ic_idx = idx; p.ic_kt = Kt; p.ic_nw = n_w; p.ic_nh = n_h; p.ic_tub = tub; p.ic_c = C; p.ic_t = T;

  CUtensorMap ma, mb, mc, mp;
  memset(&mb, 0, sizeof(mb)); memset(&mc, 0, sizeof(mc)); memset(&mp, 0, sizeof(mp));
  int rc = get_tensor_map(dy, (uint64_t)D, (uint64_t)tokens, (uint64_t)D, 64, UG_BK, &ma);
  if (rc) return rc;
  auto kern = gemm_umma_kernel<EPI_TRANSPOSED, 8, false, 2>;
  const int smem = UG_SMEM_BYTES + 512;                     // + the token-offset tables of the gather
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [&] { attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); });
  AVJ_CHECK(attr_err == cudaSuccess, "cudaFuncSetAttribute(patch-embed wgrad kernel) failed: %s", cudaGetErrorString(attr_err));
  const int units = tiles * p.split_k;
  avj_launch_pdl(kern, dim3(units < sms ? units : sms), dim3(128 + 32 * 8), (size_t)smem, s, ma, mb, mc, mp, p);
  AVJ_LAUNCH_CHECK();
  return 0;
}
