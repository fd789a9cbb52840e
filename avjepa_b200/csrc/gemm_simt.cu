// Check-mode GEMM: plain fp32-FMA shared-memory tiled kernel for all three layouts and both
// operand dtypes, sharing the fused epilogue with the tcgen05 kernel.  This is the 1e-4
// "fp32 check mode" path (and the fallback for shapes the tensor-core kernel does not take);
// it is NOT the production path -- see gemm_umma.cu.
#include "common.cuh"

#define SG_BM 64
#define SG_BN 128
#define SG_BK 16

template <typename T, int LAYOUT>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const T* __restrict__ A, const T* __restrict__ B, void* __restrict__ C,
                 int M, int N, int K, int lda, int ldb, int ldc, avj_epilogue ep) {
  __shared__ float As[SG_BK][SG_BM + 4];
  __shared__ float Bs[SG_BK][SG_BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
  const int ty = tid / 16, tx = tid % 16;     // 4 rows x 8 cols per thread
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += SG_BK) {
    // ---- A tile -> As[k][m]
#pragma unroll
    for (int it = 0; it < (SG_BM * SG_BK) / 256; ++it) {
      const int e = tid + it * 256;
      int m, k;
      if (LAYOUT == AVJ_GEMM_TN) { k = e / SG_BM; m = e % SG_BM; } else { m = e / SG_BK; k = e % SG_BK; }
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < K)
        v = to_f32<T>(LAYOUT == AVJ_GEMM_TN ? A[(int64_t)gk * lda + gm] : A[(int64_t)gm * lda + gk]);
      As[k][m] = v;
    }
    // ---- B tile -> Bs[k][n]
#pragma unroll
    for (int it = 0; it < (SG_BN * SG_BK) / 256; ++it) {
      const int e = tid + it * 256;
      int n, k;
      if (LAYOUT == AVJ_GEMM_NT) { n = e / SG_BK; k = e % SG_BK; } else { k = e / SG_BN; n = e % SG_BN; }
      const int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < N && gk < K)
        v = to_f32<T>(LAYOUT == AVJ_GEMM_NT ? B[(int64_t)gn * ldb + gk] : B[(int64_t)gk * ldb + gn]);
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      float a[4], b[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = Bs[k][tx * 8 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  const int nc = n0 + tx * 8;
  if (nc >= N) return;   // N % 8 == 0 is enforced by the caller
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = m0 + ty * 4 + i;
    if (r < M) {
      float add[8];
      epilogue_prefetch<8>(ep, C, ldc, N, r, nc, add);
      epilogue_apply_store<T, 8, false>(ep, C, ldc, N, r, nc, acc[i], add);
    }
  }
}

template <typename T>
static int launch_simt(int layout, const T* A, const T* B, void* C, int M, int N, int K, int lda, int ldb, int ldc,
                       const avj_epilogue& ep, cudaStream_t s) {
  dim3 grid((N + SG_BN - 1) / SG_BN, (M + SG_BM - 1) / SG_BM);
  switch (layout) {
    case AVJ_GEMM_NT: gemm_simt_kernel<T, AVJ_GEMM_NT><<<grid, 256, 0, s>>>(A, B, C, M, N, K, lda, ldb, ldc, ep); break;
    case AVJ_GEMM_NN: gemm_simt_kernel<T, AVJ_GEMM_NN><<<grid, 256, 0, s>>>(A, B, C, M, N, K, lda, ldb, ldc, ep); break;
    case AVJ_GEMM_TN: gemm_simt_kernel<T, AVJ_GEMM_TN><<<grid, 256, 0, s>>>(A, B, C, M, N, K, lda, ldb, ldc, ep); break;
    default: avj_set_error("avj_gemm: bad layout %d", layout); return 1;
  }
  AVJ_LAUNCH_CHECK();
  return 0;
}

int avj_gemm_simt(int dtype, int layout, const void* A, const void* B, void* C, int M, int N, int K,
                  int lda, int ldb, int ldc, const avj_epilogue& ep, cudaStream_t s) {
  if (dtype == AVJ_BF16) return launch_simt<bf16>(layout, (const bf16*)A, (const bf16*)B, C, M, N, K, lda, ldb, ldc, ep, s);
  return launch_simt<float>(layout, (const float*)A, (const float*)B, C, M, N, K, lda, ldb, ldc, ep, s);
}
