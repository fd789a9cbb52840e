// C-ABI dispatch for the compute-bound entry points (GEMM, attention): picks the tensor-core
// kernel when the problem qualifies, the fp32-FMA check-mode kernel otherwise.  The choice can
// be forced with AVJ_FORCE_SIMT=1 (debugging / A-B parity runs).
#include "common.cuh"

#include <stdlib.h>
#include <mutex>
#include <vector>

bool avj_pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("AVJ_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

// ---- launch counter (AVJ_LAUNCH_CHECK feeds it) ------------------------------------------------
unsigned long long g_avj_launches = 0;
extern "C" int64_t avj_launch_count(void) { return (int64_t)__atomic_load_n(&g_avj_launches, __ATOMIC_RELAXED); }

// ---- kernel timing ---------------------------------------------------------------------------
bool g_avj_prof_on = false;
namespace {
struct ProfRec { cudaEvent_t a, b; int family; double work; int d[4]; };
std::vector<ProfRec> g_prof;
std::mutex g_prof_mu;
thread_local int t_prof_open = -1;
}  // namespace

void avj_prof_begin(int family, double work, cudaStream_t s, int d0, int d1, int d2, int d3) {
  std::lock_guard<std::mutex> g(g_prof_mu);
  if (t_prof_open >= 0) return;                    // nested entry point: the outer scope owns the record
  ProfRec r;
  r.family = family; r.work = work;
  r.d[0] = d0; r.d[1] = d1; r.d[2] = d2; r.d[3] = d3;
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
  cudaEventRecord(r.a, s);
  g_prof.push_back(r);
  t_prof_open = (int)g_prof.size() - 1;
}
void avj_prof_end(cudaStream_t s) {
  std::lock_guard<std::mutex> g(g_prof_mu);
  if (t_prof_open < 0) return;
  cudaEventRecord(g_prof[t_prof_open].b, s);
  t_prof_open = -1;
}
extern "C" int avj_prof_enable(int on) {
  std::lock_guard<std::mutex> g(g_prof_mu);
  for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  g_prof.clear();
  t_prof_open = -1;
  g_avj_prof_on = on != 0;
  return 0;
}
// Sums the records of one family (synchronises on their events).  Any output pointer may be NULL.
extern "C" int avj_prof_collect(int family, double* ms, double* work, int* launches) {
  std::lock_guard<std::mutex> g(g_prof_mu);
  double t = 0.0, w = 0.0;
  int n = 0;
  for (auto& r : g_prof) {
    if (r.family != family) continue;
    AVJ_CUDA(cudaEventSynchronize(r.b));
    float e = 0.f;
    AVJ_CUDA(cudaEventElapsedTime(&e, r.a, r.b));
    t += e; w += r.work; ++n;
  }
  if (ms) *ms = t;
  if (work) *work = w;
  if (launches) *launches = n;
  return 0;
}
// One CSV line per record (family, work, ms, d0..d3) in launch order -- tools/step_breakdown.py groups them by shape.
extern "C" int avj_prof_dump(const char* path) {
  std::lock_guard<std::mutex> g(g_prof_mu);
  AVJ_CHECK(path != nullptr, "avj_prof_dump: NULL path");
  FILE* f = fopen(path, "w");
  AVJ_CHECK(f != nullptr, "avj_prof_dump: cannot open %s", path);
  fprintf(f, "family,work,ms,d0,d1,d2,d3\n");
  for (auto& r : g_prof) {
    float e = 0.f;
    if (cudaEventSynchronize(r.b) != cudaSuccess || cudaEventElapsedTime(&e, r.a, r.b) != cudaSuccess) { cudaGetLastError(); continue; }
    fprintf(f, "%d,%.12g,%.6f,%d,%d,%d,%d\n", r.family, r.work, e, r.d[0], r.d[1], r.d[2], r.d[3]);
  }
  fclose(f);
  return 0;
}

int avj_gemm_simt(int dtype, int layout, const void* A, const void* B, void* C, int M, int N, int K,
                  int lda, int ldb, int ldc, const avj_epilogue& ep, cudaStream_t s);
bool avj_gemm_umma_supported(int dtype, int layout, const void* A, const void* B, int M, int N, int K,
                             int lda, int ldb, const avj_epilogue& ep);
int avj_gemm_umma(int layout, const void* A, const void* B, void* C, int M, int N, int K,
                  int lda, int ldb, int ldc, const avj_epilogue& ep, cudaStream_t s);
int avj_attention_fwd_simt(int dtype, const void* qkv, void* out, float* lse, int B, int N, int H, int hd, float scale, cudaStream_t s);
int avj_attention_bwd_simt(int dtype, const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                           float* ws, int B, int N, int H, int hd, float scale, cudaStream_t s);
bool avj_attention_mma_supported(int dtype, int hd);
int avj_attention_fwd_mma(const void* qkv, void* out, float* lse, int B, int N, int H, int hd, float scale, cudaStream_t s);
int avj_attention_bwd_mma(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                          float* ws, int B, int N, int H, int hd, float scale, cudaStream_t s);

bool avj_attention_umma_fwd_supported(int dtype, int hd);
int avj_attention_fwd_umma(const void* qkv, void* out, float* lse, int B, int N, int H, int hd, float scale, cudaStream_t s);

bool avj_attention_umma_bwd_supported(int dtype, int hd);
int avj_attention_bwd_umma(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* ws,
                           int B, int N, int H, int hd, float scale, cudaStream_t s);
static bool attn_bwd_use_umma() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("AVJ_ATTN_BWD"); v = (e && e[0] == 'm') ? 0 : 1; }
  return v == 1;
}

// AVJ_ATTN_FWD = umma (default) | mma : which tensor-core forward kernel serves bf16 attention
static bool attn_fwd_use_umma() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("AVJ_ATTN_FWD"); v = (e && e[0] == 'm') ? 0 : 1; }
  return v == 1;
}

static bool force_simt() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("AVJ_FORCE_SIMT"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}
// AVJ_GEMM_SIMT_FALLBACK=1 lets a bf16 GEMM the tcgen05 kernel refuses run on the fp32-FMA check kernel
// (~100x slower); by default such a call is an ERROR so that a mis-shaped production call cannot hide.
static bool allow_simt_fallback() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("AVJ_GEMM_SIMT_FALLBACK"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}
static bool force_simt_attn() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("AVJ_FORCE_SIMT_ATTN"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1 || force_simt();
}

extern "C" int avj_gemm(int dtype, int layout, const void* A, const void* B, void* C,
                        int M, int N, int K, int lda, int ldb, int ldc,
                        const avj_epilogue* ep_in, void* stream) {
  AVJ_CHECK(ep_in != nullptr, "avj_gemm: epilogue must not be NULL");
  AVJ_CHECK(dtype == AVJ_F32 || dtype == AVJ_BF16, "avj_gemm: bad dtype %d", dtype);
  AVJ_CHECK(N % 8 == 0 && ldc % 8 == 0, "avj_gemm: N and ldc must be multiples of 8 (N=%d ldc=%d)", N, ldc);
  AVJ_CHECK(!(ep_in->accumulate && ep_in->out_dtype != AVJ_F32), "avj_gemm: accumulate needs an fp32 C");
  if (M == 0 || N == 0) return 0;
  avj_epilogue ep = *ep_in;
  if (K == 0) {
    AVJ_CHECK(ep.accumulate, "avj_gemm: K == 0 is only meaningful for accumulate epilogues");
    return 0;
  }
  const int epi_bits = layout | (ep.bias ? 4 : 0) | (ep.act ? 8 : 0) | (ep.residual ? 16 : 0) | (ep.accumulate ? 32 : 0) |
                       (ep.dact_aux ? 64 : 0) | (ep.out_dtype == AVJ_F32 ? 128 : 0);
  AvjProfScope prof(AVJ_FAM_GEMM, 2.0 * M * (double)N * K, stream, epi_bits, M, N, K);
  if (!force_simt() && avj_gemm_umma_supported(dtype, layout, A, B, M, N, K, lda, ldb, ep))
    return avj_gemm_umma(layout, A, B, C, M, N, K, lda, ldb, ldc, ep, as_stream(stream));
  AVJ_CHECK(dtype == AVJ_F32 || force_simt() || allow_simt_fallback(),
            "avj_gemm: bf16 problem not supported by the tcgen05 kernel (layout=%d M=%d N=%d K=%d lda=%d ldb=%d A%%16=%d B%%16=%d; "
            "needs N %% 64 == 0, lda/ldb %% 8 == 0, 16-byte aligned A/B, and M %% 8 == 0 for TN). Set AVJ_GEMM_SIMT_FALLBACK=1 "
            "(or AVJ_FORCE_SIMT=1) to run it on the fp32-FMA check kernel instead.",
            layout, M, N, K, lda, ldb, (int)(reinterpret_cast<uintptr_t>(A) & 15), (int)(reinterpret_cast<uintptr_t>(B) & 15));
  return avj_gemm_simt(dtype, layout, A, B, C, M, N, K, lda, ldb, ldc, ep, as_stream(stream));
}

bool avj_patch_embed_umma_supported(int patch, int H, int W, int T, int tub, int D);
int avj_patch_embed_umma(const float* x, const int64_t* idx, const float* w, float* out, int B, int C, int T, int H, int W, int tub,
                         int patch, int Kt, int D, int ldc, const avj_epilogue& ep, cudaStream_t s);

int avj_patch_embed_wgrad_umma(const float* x, const int64_t* idx, const void* dy, float* gw, int B, int C, int T, int H, int W, int tub,
                               int patch, int Kt, int D, cudaStream_t s);

extern "C" int avj_patch_embed_wgrad(const float* x, const int64_t* idx, const void* dy, float* gw,
                                     int B, int C, int T, int H, int W, int tub, int patch, int K, int D, void* stream) {
  if (B == 0 || K == 0) return 0;
  const int kd = C * tub * patch * patch;
  // profiled as a GEMM (TN, accumulate, fp32 out) of M = D, N = kd, K = B*K tokens
  AvjProfScope prof(AVJ_FAM_GEMM, 2.0 * B * K * (double)D * kd, stream, 2 | 32 | 128, D, kd, B * K);
  return avj_patch_embed_wgrad_umma(x, idx, dy, gw, B, C, T, H, W, tub, patch, K, D, as_stream(stream));
}

extern "C" int avj_patch_embed_supported(int patch, int H, int W, int T, int tub, int D) {
  return avj_patch_embed_umma_supported(patch, H, W, T, tub, D) ? 1 : 0;
}

extern "C" int avj_patch_embed(const float* x, const int64_t* idx, const float* w, float* out,
                               int B, int C, int T, int H, int W, int tub, int patch, int K, int D, int ldc,
                               const avj_epilogue* ep, void* stream) {
  AVJ_CHECK(ep != nullptr, "avj_patch_embed: epilogue must not be NULL");
  if (B == 0 || K == 0) return 0;
  const int kd = C * tub * patch * patch;
  // profiled as a GEMM (NT, bias, fp32 out) of M = B*K, N = D, K = kd
  AvjProfScope prof(AVJ_FAM_GEMM, 2.0 * B * K * (double)D * kd, stream, 0 | 4 | 128 | 256, B * K, D, kd);
  return avj_patch_embed_umma(x, idx, w, out, B, C, T, H, W, tub, patch, K, D, ldc, *ep, as_stream(stream));
}

extern "C" int64_t avj_attention_bwd_ws_floats(int B, int N, int H, int hd) {
  // delta and lse*log2(e) with rows padded to 64, plus the fp32 dQ accumulator [B, N, H, hd] of the single-pass kernel
  return 2 * (int64_t)B * H * ((N + 63) / 64 * 64) + (int64_t)B * N * H * hd;
}

extern "C" int avj_attention_fwd(int dtype, const void* qkv, void* out, float* lse,
                                 int B, int N, int H, int hd, float scale, void* stream) {
  AVJ_CHECK(hd > 0 && hd <= 128, "avj_attention_fwd: head_dim %d out of range (1..128)", hd);
  if (B == 0 || N == 0) return 0;
  AvjProfScope prof(AVJ_FAM_ATTN_FWD, 4.0 * B * H * (double)N * N * hd, stream, B, N, H, hd);
  if (!force_simt_attn() && attn_fwd_use_umma() && avj_attention_umma_fwd_supported(dtype, hd))
    return avj_attention_fwd_umma(qkv, out, lse, B, N, H, hd, scale, as_stream(stream));
  if (!force_simt_attn() && avj_attention_mma_supported(dtype, hd))
    return avj_attention_fwd_mma(qkv, out, lse, B, N, H, hd, scale, as_stream(stream));
  return avj_attention_fwd_simt(dtype, qkv, out, lse, B, N, H, hd, scale, as_stream(stream));
}

extern "C" int avj_attention_bwd(int dtype, const void* qkv, const void* out, const void* dout,
                                 const float* lse, void* dqkv, float* ws,
                                 int B, int N, int H, int hd, float scale, void* stream) {
  AVJ_CHECK(hd > 0 && hd <= 128, "avj_attention_bwd: head_dim %d out of range (1..128)", hd);
  AVJ_CHECK(ws != nullptr, "avj_attention_bwd: workspace required");
  if (B == 0 || N == 0) return 0;
  AvjProfScope prof(AVJ_FAM_ATTN_BWD, 10.0 * B * H * (double)N * N * hd, stream, B, N, H, hd);
  if (!force_simt_attn() && attn_bwd_use_umma() && avj_attention_umma_bwd_supported(dtype, hd))
    return avj_attention_bwd_umma(qkv, out, dout, lse, dqkv, ws, B, N, H, hd, scale, as_stream(stream));
  if (!force_simt_attn() && avj_attention_mma_supported(dtype, hd))
    return avj_attention_bwd_mma(qkv, out, dout, lse, dqkv, ws, B, N, H, hd, scale, as_stream(stream));
  return avj_attention_bwd_simt(dtype, qkv, out, dout, lse, dqkv, ws, B, N, H, hd, scale, as_stream(stream));
}
