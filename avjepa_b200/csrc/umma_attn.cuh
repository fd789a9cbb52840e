// tcgen05 / TMEM / TMA / mbarrier PTX wrappers and the SWIZZLE_128B tile staging shared by the
// attention forward and backward kernels (attention_umma.cu, attention_umma_bwd.cu).
#pragma once
#include "common.cuh"

#include <cuda.h>

__device__ __forceinline__ uint32_t ua_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ua_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void ua_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void ua_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void ua_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ua_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ua_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void ua_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void ua_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// shared-memory matrix descriptor (sm_100 version field = 1); layout type 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
__device__ __forceinline__ uint64_t ua_desc(uint32_t saddr, uint32_t lbo16, uint32_t sbo16, uint32_t layout = 2) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(lbo16 & 0x3FFF) << 16) | ((uint64_t)(sbo16 & 0x3FFF) << 32) |
         (1ull << 46) | ((uint64_t)layout << 61);
}

// Geometry of an operand tile whose rows hold HDP bf16 of one head: HDP = 64 -> 128-byte rows, SWIZZLE_128B;
// HDP = 32 -> 64-byte rows, SWIZZLE_64B (half the shared memory and L2 traffic of padding to 128 bytes).
// HDP = 128 (head_dim 80 .. 128): TWO 64-column SWIZZLE_128B sub-tiles ("halves") of 128 rows each, HALF bytes apart; the
// geometry constants below describe one half.
template <int HDP> struct UaTile {
  static constexpr int NH = HDP == 128 ? 2 : 1;                      // 64-column halves per tile
  static constexpr int PITCH = HDP == 32 ? 64 : 128;                 // bytes per row
  static constexpr uint32_t LAYOUT = HDP == 32 ? 4u : 2u;            // descriptor layout type
  static constexpr uint32_t SBO = 8 * PITCH / 16;                    // 8-row group stride, 16-byte units
  static constexpr uint32_t MN_KADV = 16 * PITCH / 16;               // MN-major: 16 K-rows per UMMA_K step
  __device__ static __forceinline__ uint32_t chunk(int r, int c) {   // physical 16-byte chunk of logical chunk c
    return HDP == 32 ? (uint32_t)(c ^ ((r >> 1) & 3)) : (uint32_t)(c ^ (r & 7));
  }
  __device__ static __forceinline__ uint64_t kmajor(uint32_t saddr) { return ua_desc(saddr, 1, SBO, LAYOUT); }
  __device__ static __forceinline__ uint64_t mnmajor(uint32_t saddr) { return ua_desc(saddr, 512, SBO, LAYOUT); }
};
__device__ __forceinline__ void ua_cp16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src));
}
__device__ __forceinline__ void ua_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void ua_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void ua_st32(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
        "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
        "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
        "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
        "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// tcgen05.mma with the A operand in TENSOR memory (lane = row, one 32-bit column = two consecutive bf16 along K):
// P / dS go from the softmax registers straight to the tensor core -- no shared-memory round trip, no
// generic->async proxy fence.  One instruction covers UMMA_K = 16, i.e. 8 TMEM columns of A.
__device__ __forceinline__ void ua_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// registers -> TMEM, 32 lanes x NR columns (NR = 16 or 32); the caller issues tcgen05.wait::st
template <int NR> __device__ __forceinline__ void ua_st_regs(uint32_t taddr, const uint32_t* v);
template <> __device__ __forceinline__ void ua_st_regs<32>(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
        "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
        "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
template <> __device__ __forceinline__ void ua_st_regs<16>(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void ua_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// `nthreads` threads stage ROWS rows x HDP bf16 (hd real columns, rest zero) into a swizzled UaTile<HDP> tile.
template <int HDP, int ROWS = 128>
__device__ __forceinline__ void ua_stage(uint32_t tile, const bf16* __restrict__ src, int64_t row_stride, int row0,
                                         int nrows_total, int hd, int tid, int nthreads = 32) {
  constexpr int CH = HDP / 8;
  const int hd_ch = hd / 8;
  for (int e = tid; e < ROWS * CH; e += nthreads) {
    const int r = e / CH, c = e % CH;
    const uint32_t dst = tile + r * UaTile<HDP>::PITCH + (UaTile<HDP>::chunk(r, c) << 4);
    if (row0 + r < nrows_total && c < hd_ch) {
      ua_cp16(dst, src + (int64_t)(row0 + r) * row_stride + c * 8);
    } else {
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0u) : "memory");
    }
  }
}

// One warp zeroes columns [hd, HDP) of a TMA-landed `rows` x HDP tile (generic-proxy stores: the caller
// fences with fence.proxy.async before the tile is handed to the tensor core).
template <int HDP>
__device__ __forceinline__ void ua_zero_pad(uint32_t tile, int rows, int hd, int lane) {
  constexpr int CH = HDP / 8;
  const int c0 = hd / 8, npad = CH - c0;
  for (int e = lane; e < rows * npad; e += 32) {
    const int r = e / npad, c = c0 + e % npad;
    const uint32_t dst = tile + r * UaTile<HDP>::PITCH + (UaTile<HDP>::chunk(r, c) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0u) : "memory");
  }
}

// Register re-partitioning between warpgroups (warps 4k..4k+3 must all execute the same one).
template <int R> __device__ __forceinline__ void ua_reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R> __device__ __forceinline__ void ua_reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }

__device__ __forceinline__ float ua_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exp2 on the FMA pipe (Cody-Waite split + degree-3 minimax of 2^r on [-0.5, 0.5], max relative error 7.5e-5,
// far below the bf16 rounding of P): 8 FMA-pipe instructions.  The softmax loops route a fixed fraction of their
// elements here because the 16 exp2/clk/SM MUFU units, not the tensor pipe, bound attention at head_dim <= 64.
__device__ __forceinline__ float ua_exp2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;                       // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float r = x - (t - 12582912.0f);
  float p = fmaf(0.0551716685f, r, 0.2426111251f);
  p = fmaf(p, r, 0.6932609677f);
  p = fmaf(p, r, 0.9999280572f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// 3-D TMA load of one 128-row x 64-col bf16 box of the qkv tensor viewed as {3*H*hd, N, B}
__device__ __forceinline__ void ua_tma3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void ua_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

