/*
 * avjepa_b200.h -- C ABI of libavjepa_sm100.so, the sm_100a kernels behind the AV-JEPA
 * masked encoder/predictor training step.
 *
 * The reference (johnshizhu/AVJEPA) is pure Python/PyTorch and has no FFI of its own; the
 * drop-in boundary is its Python module API.  This header is the
 * boundary UNDER that API: each entry point names the reference call site (file:line,
 * relative to the reference root) whose device work it replaces.  Callers own every buffer
 * and pass raw device pointers, sizes and a cudaStream_t; no torch types cross this line.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; avj_last_error_string()
 *     describes the last failure on the calling thread.  Nothing throws.
 *   - all matrices are row-major; `ld*` are leading dimensions in ELEMENTS.
 *   - dtype codes: AVJ_F32 (fp32 "check mode": SIMT kernels, 1e-4 parity) and
 *     AVJ_BF16 (production: tcgen05/TMEM/TMA tensor-core kernels, fp32 accumulate).
 *   - `stream` is a cudaStream_t passed as void*.
 *   - a "row map" sends logical row r to physical row
 *         (r / rows_per_group) * group_stride + (r % rows_per_group) + row_offset;
 *     rows_per_group == 0 means identity.  It lets kernels read/write the concatenated
 *     [ctx_v | tgt_v | ctx_a | tgt_a] token layouts in place (no torch.cat copies).
 */
#ifndef AVJEPA_B200_H_
#define AVJEPA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AVJ_ABI_VERSION 1

enum { AVJ_F32 = 0, AVJ_BF16 = 1 };

/* GEMM operand layouts: which index is contiguous in memory. */
enum {
  AVJ_GEMM_NT = 0, /* C[M,N] = A[M,K] . B[N,K]^T   (fprop:  y = x W^T)          */
  AVJ_GEMM_NN = 1, /* C[M,N] = A[M,K] . B[K,N]     (dgrad:  dx = dy W)          */
  AVJ_GEMM_TN = 2  /* C[M,N] = A[K,M]^T . B[K,N]   (wgrad:  dW = dy^T x)        */
};

typedef struct avj_rowmap {
  int32_t rows_per_group; /* 0 = identity */
  int32_t group_stride;
  int32_t row_offset;
} avj_rowmap;

/* Fused GEMM epilogue: C[map(r), n] (+)= act(acc + bias[n]) * gelu'(aux[r,n])
 *                                        + residual[map(r), n] + pos[pos_idx[r], n]      */
typedef struct avj_epilogue {
  const float* bias;       /* [N] or NULL                                                  */
  const float* residual;   /* fp32, same physical layout as C (ldc), or NULL               */
  const float* pos;        /* fp32 [pos_rows, N] positional table, or NULL                 */
  const int64_t* pos_idx;  /* [M] token ids into pos; NULL -> r % pos_rows                 */
  int32_t pos_rows;
  int32_t act;             /* 0 none, 1 exact-erf GELU                                     */
  void* pre_out;           /* act!=0: also store the pre-activation, [M,N] ld=N, dtype     */
  const void* dact_aux;    /* != NULL: multiply acc by gelu'(dact_aux[r,n]) ([M,N], dtype) */
  int32_t accumulate;      /* 1: C += result (fp32 C only; weight gradients)               */
  int32_t out_dtype;       /* AVJ_F32 or AVJ_BF16                                          */
  avj_rowmap out_map;      /* physical row of C / residual for logical row r               */
} avj_epilogue;

int avj_version(void);
const char* avj_last_error_string(void);
/* 1 when the device behind the current context is sm_100 and the tcgen05 path is usable. */
int avj_device_ok(void);

/* Number of kernels this library has launched in the calling process so far (every launch site counts itself);
 * bench.py reports the difference over its timed region as `gpu_launches`. */
int64_t avj_launch_count(void);

/* ---- K6: nn.Linear fprop/dgrad/wgrad (src/models/utils/modules.py:25-36,54-77) and the
 *      patch-embed / predictor-embed GEMMs (patch_embed.py:85-101, audiovisionpredictor.py:
 *      232-243).  dtype selects the operand type of A and B.  bf16 problems run on the tcgen05 kernel and must
 *      satisfy N % 64 == 0, lda/ldb % 8 == 0, 16-byte aligned A/B (TN: M % 8 == 0); anything else is an ERROR
 *      (never a silent fp32-FMA fallback) unless AVJ_GEMM_SIMT_FALLBACK=1 / AVJ_FORCE_SIMT=1 is set. */
int avj_gemm(int dtype, int layout, const void* A, const void* B, void* C,
             int M, int N, int K, int lda, int ldb, int ldc,
             const avj_epilogue* ep, void* stream);

/* ---- K1/K2: tubelet / mel patch extraction feeding the patch-embed GEMM
 *      (src/models/utils/patch_embed.py:98-102).  x is fp32 [B, C, T, H, W]; out row
 *      r = b*K + j holds the flattened (c, dt, dh, dw) patch of token idx[r] (or token j
 *      when idx == NULL, K == tokens per clip), i.e. the Conv3d weight's own K order, so
 *      only the kept tokens are ever embedded.  Audio: C=1, T=1, tub=1. */
int avj_patchify(const float* x, const int64_t* idx, void* out, int out_dtype,
                 int B, int C, int T, int H, int W, int tub, int patch, int K, void* stream);

/* ---- K1: the whole patch-embedding projection, im2col-free (src/models/utils/patch_embed.py:85-102:
 *      `self.proj(x).flatten(2).transpose(1, 2)` of PatchEmbed3D / AudioVisionPatchEmbed3D.forward, then the
 *      positional-embedding add of audiovision_transformer.py forward):
 *        out[omap(b*K + j), :] = w[D, C*tub*256] . patch(x[b], token idx[b, j]) + ep->bias + ep->pos[idx[b, j], :]
 *      x fp32 [B, C, T, H, W] (audio: C = 1, T = tub = 1), w the fp32 Conv weight viewed [D, C*tub*16*16], out fp32 with row
 *      pitch ldc.  The patch rows (64-byte runs) are gathered by the producer warps with 16-byte cp.async copies straight out
 *      of x into the swizzled operand tile, the weight tile arrives by TMA, and both are multiplied on the tensor cores
 *      as tf32 (tcgen05.mma.kind::tf32); no patch matrix exists in HBM.  idx == NULL
 *      embeds every token (K == tokens per clip).  Only ep->bias, ep->pos / pos_idx / pos_rows and ep->out_map are
 *      honoured (out_dtype must be AVJ_F32).  patch must be 16 and D a multiple of 64; other geometries are an
 *      error (callers use avj_patchify + avj_gemm there, as the fp32 check mode does). */
int avj_patch_embed(const float* x, const int64_t* idx, const float* w, float* out,
                    int B, int C, int T, int H, int W, int tub, int patch, int K, int D, int ldc,
                    const avj_epilogue* ep, void* stream);
/* Weight gradient of the same projection, also without a patch matrix (autograd of patch_embed.py:85-102 w.r.t. the Conv weight):
 *        gw[D, C*tub*256] += sum over b, j of dy[b*K + j, :]^T (x) patch(x[b], token idx[b, j])
 *      dy bf16 [B*K, D] row-major (the gradient of the embedded rows, compute dtype), gw the fp32 gradient of the Conv weight
 *      viewed [D, C*tub*16*16].  The patch values are gathered out of x and converted to bf16 by the producer warps of the
 *      tcgen05 GEMM (MN-major operand tile); the token dimension is split over CTAs (fp32 atomics into gw). */
int avj_patch_embed_wgrad(const float* x, const int64_t* idx, const void* dy, float* gw,
                          int B, int C, int T, int H, int W, int tub, int patch, int K, int D, void* stream);
/* 1 when avj_patch_embed accepts this geometry. */
int avj_patch_embed_supported(int patch, int H, int W, int T, int tub, int D);

/* ---- device-side mask sampling (src/masks/avmultiblock3d.py:131-234, the per-sample loop of _AVMaskGenerator.__call__ for all
 *      mask generators of one step): block origins are drawn on the GPU from rng_state, a replica of torch's CPU generator
 *      (int32[626] in device memory = 624 MT19937 words, values left in the block, index of the next word; updated in place),
 *      in the reference's draw order (video top, left, start, audio top, left per block; a sample whose video context comes out
 *      empty is drawn again).  gens is a HOST array [n_gen][5] = {t, h, w, blocks per sample, max_context_duration} (the block
 *      size of this step, drawn by the caller from the seeded per-step generator).  For generator g and sample b the ascending
 *      kept video ids go to enc_v[g][b][0 .. counts[g][b][0]), kept audio ids to enc_a (counts[..][1]), dropped video / audio
 *      ids to pred_v / pred_a (counts[..][2], [3]); rows are NV = duration*height*width (NA = a_height*a_width) wide.
 *      status (zeroed by the caller): bit 0 = some index set has exactly one element (the reference raises TypeError there),
 *      bit 1 = gave up resampling.  One single-CTA launch. */
int avj_mask_collate(void* rng_state, const int32_t* gens, int n_gen, int B, int duration, int height, int width,
                     int a_height, int a_width, int a_block_h, int a_block_w, int64_t* enc_v, int64_t* pred_v,
                     int64_t* enc_a, int64_t* pred_a, int32_t* counts, int32_t* status, void* stream);

/* ---- K3: apply_masks (src/masks/utils.py:14-34) for one mask, and its backward.
 *      out[b, j, :] = x[b, idx[b, j], :];  bwd: dx[b, idx[b, j], :] += dout[b, j, :]. */
int avj_gather_rows_fwd(int dtype, const void* x, const int64_t* idx, void* out,
                        int B, int N, int K, int D, void* stream);
int avj_gather_rows_bwd(int dtype, const void* dout, const int64_t* idx, void* dx,
                        int B, int N, int K, int D, void* stream);

/* ---- generic row-mapped copy / cast / accumulate:
 *      out[omap(r), :] (=|+=) in[imap(r), :], r in [0, rows).  Replaces the torch.split /
 *      torch.cat / .to(dtype) glue of app/avjepa/train.py:449-455,476-486 and
 *      audiovisionpredictor.py:272-299. */
int avj_copy_rows(const void* in, int in_dtype, int ld_in, avj_rowmap imap,
                  void* out, int out_dtype, int ld_out, avj_rowmap omap,
                  int rows, int D, int accumulate, void* stream);

/* ---- K4: predictor target rows (audiovisionpredictor.py:245-269):
 *      x[map(r), :] = mask_token[:] + pos[idx[r], :], r in [0, rows). */
int avj_fill_mask_tokens(const float* mask_token, const float* pos, const int64_t* idx,
                         float* x, int ld, avj_rowmap map, int rows, int D, void* stream);

/* ---- column sum over mapped rows: out[n] += sum_r in[map(r), n]  (bias and mask-token
 *      gradients).  ws: fp32 workspace of avj_colsum_ws_floats(rows, D) floats. */
int64_t avj_colsum_ws_floats(int rows, int D);
int avj_colsum(const void* in, int in_dtype, int ld, avj_rowmap map, float* out,
               int rows, int D, float* ws, void* stream);
/* two column sums over the same rows in ONE launch (identity row map): out1[n] += sum_r in1[r, n] (D1
 * columns, ld1) and out2 likewise.  ws: avj_colsum_ws_floats(rows, D1 + D2) floats. */
int avj_colsum2(const void* in1, int ld1, int D1, float* out1, const void* in2, int ld2, int D2, float* out2,
                int in_dtype, int rows, float* ws, void* stream);

/* ---- K5: nn.LayerNorm (modules.py:115,119; eps 1e-6) and F.layer_norm without affine
 *      (app/avjepa/train.py:448; eps 1e-5).  x fp32 [rows, D]; y in y_dtype; gamma/beta may
 *      be NULL (no affine); mean/rstd may be NULL (inference). */
int avj_layernorm_fwd(const float* x, const float* gamma, const float* beta,
                      void* y, int y_dtype, float* mean, float* rstd,
                      int rows, int D, float eps, void* stream);
/* dx_out = dres_in + LN'(dy) (fp32);  dx_lp: optional low-precision copy of dx_out (dtype
 * lp_dtype) feeding the next dgrad/wgrad GEMM;  dgamma/dbeta are ACCUMULATED (may be NULL);
 * dcolsum (may be NULL): dcolsum[n] += sum_r dx_out[r, n] -- the bias gradient of the Linear whose
 * output is added to this residual stream (proj / fc2), fused here instead of a separate pass.
 * ws: avj_layernorm_bwd_ws_floats(rows, D) floats. */
int64_t avj_layernorm_bwd_ws_floats(int rows, int D);
int avj_layernorm_bwd(const void* dy, int dy_dtype, const float* x, const float* gamma,
                      const float* mean, const float* rstd, const float* dres_in,
                      float* dx_out, void* dx_lp, int lp_dtype,
                      float* dgamma, float* dbeta, float* dcolsum, float* ws,
                      int rows, int D, void* stream);

/* ---- K7: F.scaled_dot_product_attention(q, k, v) (modules.py:66-69): non-causal, no
 *      mask, no dropout, scale = hd^-0.5.  qkv is the qkv-Linear output [B, N, 3, H, hd]
 *      read in place (no permute copies); out is [B, N, H*hd]; lse fp32 [B, H, N]. */
int avj_attention_fwd(int dtype, const void* qkv, void* out, float* lse,
                      int B, int N, int H, int hd, float scale, void* stream);
/* dqkv [B, N, 3, H, hd];  ws: avj_attention_bwd_ws_floats(B,N,H,hd) floats (row statistics and, for head_dim <= 32,
 * the fp32 dQ accumulator [B, N, H, hd] of the single-pass kernel: dK / dV / dQ from ONE evaluation of every exp2, the
 * dQ tiles reduced across key blocks with TMA reduce-add). */
int64_t avj_attention_bwd_ws_floats(int B, int N, int H, int hd);
int avj_attention_bwd(int dtype, const void* qkv, const void* out, const void* dout,
                      const float* lse, void* dqkv, float* ws,
                      int B, int N, int H, int hd, float scale, void* stream);

/* ---- few-query cross-attention of the attentive probes (src/models/attentive_pooler.py:21-102 ->
 *      CrossAttention.forward, src/models/utils/modules.py:139-159): n learned queries over N encoder tokens, scale hd^-0.5.
 *      q [B, n, H, hd] (q-Linear output), kv [B, N, 2, H, hd] (kv-Linear output, read in place), out [B, n, H*hd],
 *      lse fp32 [B, H, n].  HBM-bound by construction (every K / V row is read once per (batch, head)).
 *      bwd: dq [B, n, H, hd], dkv [B, N, 2, H, hd] (fully overwritten). */
int avj_xattn_fwd(int dtype, const void* q, const void* kv, void* out, float* lse,
                  int B, int N, int n, int H, int hd, float scale, void* stream);
int avj_xattn_bwd(int dtype, const void* q, const void* kv, const void* out, const void* dout, const float* lse,
                  void* dq, void* dkv, int B, int N, int n, int H, int hd, float scale, void* stream);

/* ---- K10: latent loss sum_i mean(|z_i - h_i|^p)/p / n_masks (app/avjepa/train.py:490-495)
 *      for ONE mask, forward + backward in one pass.  z, h fp32 [n].
 *      loss_out[0] += (1/(n*n_masks*p)) * sum |z-h|^p ;  dz = grad_scale * d/dz of that term.
 *      mode: 0 = |.|^p with p = loss_exp (reference), 1 = smooth-L1 with beta (extra).
 *      ws: avj_loss_ws_floats(n) floats. */
int64_t avj_loss_ws_floats(int64_t n);
int avj_loss_fwd_bwd(const float* z, const float* h, float* dz, float* loss_out,
                     int64_t n, int n_masks, float loss_exp, int mode, float beta,
                     float grad_scale, float* ws, void* stream);
/* token-variance regulariser value (app/avjepa/train.py:497-498,507-508): z fp32 [B,K,D] per
 * mask; pstd[b,d] += sqrt(var_k(z)+1e-4)/n_masks.  Then avj_reg_finish reduces
 * mean(relu(1-pstd)) into loss_reg[0]. */
int avj_reg_accumulate(const float* z, float* pstd, int B, int K, int D, int n_masks, void* stream);
int avj_reg_finish(const float* pstd, float* loss_reg, int n, void* stream);

/* ---- K11-K13: AdamW (torch.optim.AdamW as built by app/avjepa/utils.py:228-282) fused with
 *      grad unscale/clip, the EMA target update (app/avjepa/train.py:534-537), grad zeroing
 *      and the bf16 shadow-weight refresh, over one contiguous parameter range.
 *      scale_ptr: device float multiplied into every gradient (1/loss_scale * clip coef),
 *      may be NULL.  target/p_lp/target_lp may be NULL. */
typedef struct avj_adamw_args {
  float* p; float* g; float* m; float* v;     /* fp32 [n]                                   */
  float* target;                              /* EMA target fp32 [n] or NULL                */
  void* p_lp; void* target_lp;                /* bf16 shadows or NULL                       */
  int64_t n;
  float lr, wd, beta1, beta2, eps;
  int32_t step;                               /* 1-based                                    */
  float ema_m;
  int32_t skip_update;                        /* 1: frozen range -> EMA + shadows only      */
  int32_t zero_grad;
  const float* scale_ptr;
} avj_adamw_args;
int avj_adamw_ema_step(const avj_adamw_args* a, void* stream);

/* sum of squares of a contiguous fp32 range -> out[0] (K13, clip_grad_norm_ train.py:518-520);
 * ws: avj_sumsq_ws_floats(n).  avj_clip_coef: coef[0] = inv_loss_scale * min(1, max_norm /
 * (sqrt(sumsq)*inv_loss_scale + 1e-6)), or just inv_loss_scale when max_norm <= 0. */
int64_t avj_sumsq_ws_floats(int64_t n);
int avj_sumsq(const float* x, int64_t n, float* out, float* ws, void* stream);
int avj_clip_coef(const float* sumsq, float max_norm, float inv_loss_scale, float* coef, void* stream);

/* ---- per-parameter statistics in one pass over a flat fp32 buffer (grad_logger / adamw_logger,
 *      src/utils/logging.py:91-118, which issue one blocking float() per parameter): out[s] += sum over
 *      [seg_off[s], seg_off[s+1]) of x^2 (mode 0) or |x| (mode 1).  seg_off: nseg+1 ascending int64 device
 *      offsets, multiples of 4, seg_off[nseg] == total.  out: fp64 [nseg], ACCUMULATED (caller zeroes). */
int avj_segment_stats(const float* x, const int64_t* seg_off, int nseg, int64_t total, int mode, double* out,
                      void* stream);

/* fp32 -> bf16 (or fp32 copy) over a contiguous range: shadow-weight refresh. */
int avj_cast(const float* in, void* out, int out_dtype, int64_t n, void* stream);
/* stream-ordered zero fill of a raw device range (gradient / scatter targets). */
int avj_memset_zero(void* ptr, int64_t nbytes, void* stream);

/* ---- in-library kernel timing (measurement only): while enabled, every GEMM / attention / LayerNorm /
 *      column-sum / optimizer entry point brackets its launches with CUDA events on its stream.
 *      avj_prof_collect sums one family: 0 gemm (work = FLOPs), 1 attention fwd, 2 attention bwd (FLOPs),
 *      3 layernorm fwd, 4 layernorm bwd, 5 colsum, 6 optimizer, 7 other row kernels (work = algorithmic bytes). */
int avj_prof_enable(int on);
int avj_prof_collect(int family, double* ms, double* work, int* launches);
/* one CSV line per timed launch, in launch order: family,work,ms,d0,d1,d2,d3 (d* = the launch's shape:
 * GEMM layout|epilogue bits,M,N,K; attention B,N,H,hd; row kernels rows,D; family 7 = gather/copy/loss/...
 * with d0 = 1 patchify, 2/3 gather fwd/bwd, 4 copy_rows, 5 mask-token fill, 6 loss, 7 cast, 8 memset). */
int avj_prof_dump(const char* path);

/* ---- whole-stack schedules: L pre-LN transformer blocks (Block.forward / Attention.forward /
 *      MLP.forward, src/models/utils/modules.py:114-120, :61-78, :30-36) issued as ONE call, so the
 *      host pays one FFI crossing per stack instead of one per kernel.  Every pointer is caller
 *      owned; activation slots are the per-layer buffers the backward re-reads (pre may be NULL in
 *      a no-grad forward).  Gradient pointers (gw/gb) may be NULL (frozen parameter). */
typedef struct avj_linear_w { const void* w; const float* b; float* gw; float* gb; } avj_linear_w;
typedef struct avj_norm_w { const float* w; const float* b; float* gw; float* gb; float eps; int32_t pad_; } avj_norm_w;
typedef struct avj_layer {
  avj_norm_w n1; avj_linear_w qkv; avj_linear_w proj; avj_norm_w n2; avj_linear_w fc1; avj_linear_w fc2;
  float* x; float* mean1; float* rstd1; void* h1; void* qkv_act; void* o; float* lse;
  float* x1; float* mean2; float* rstd2; void* h2; void* pre; void* act;
  float* x_out;            /* residual stream after this block (= next layer's x)          */
} avj_layer;
#define AVJ_MAX_SEGMENTS 8
typedef struct avj_stack {
  int32_t dtype;           /* AVJ_F32 / AVJ_BF16: operand dtype of GEMMs and attention      */
  int32_t B, N, D, H, hidden, L;
  /* Variable-length batch: n_seg > 0 -> the token matrix is the concatenation of n_seg groups, group g holding
   * seg_B[g] sequences of seg_N[g] tokens each (rows of group g start at sum_{j<g} seg_B[j]*seg_N[j]); B and N are
   * ignored.  The Linear / LayerNorm kernels see one [sum B*N, D] matrix -- ONE launch per Linear for all groups
   * (the reference's MultiMask wrappers run the whole backbone once per mask, src/models/utils/multimask.py:37-46,
   * 55-71) -- and attention runs per group.  lse of group g starts at sum_{j<g} seg_B[j]*H*seg_N[j] floats.
   * n_seg == 0 -> one group (B, N). */
  int32_t n_seg;
  int32_t seg_B[AVJ_MAX_SEGMENTS];
  int32_t seg_N[AVJ_MAX_SEGMENTS];
} avj_stack;
typedef struct avj_stack_scratch {
  float* dxa; float* dxb;  /* fp32 [B*N, D] residual-gradient ping/pong; dxa holds d x_L on entry
                              and d x_0 on return                                            */
  void* dx_lp;             /* compute-dtype copy of the current residual gradient (in/out)  */
  void* d_hid; void* d_qkv; void* d_h; void* d_o;
  float* ws;               /* max of the layernorm_bwd / colsum / attention_bwd workspaces  */
  void* const* layer_done; /* NULL, or L cudaEvent_t handles (NULL entries allowed): event i is
                              recorded on the stream once every kernel of layer i's backward has
                              been enqueued, i.e. once layer i's parameter gradients are final for
                              this call -- the data-parallel host side starts that layer's
                              gradient all-reduce behind it while lower layers still compute      */
} avj_stack_scratch;
int avj_stack_forward(const avj_stack* s, const avj_layer* layers, void* stream);
int avj_stack_backward(const avj_stack* s, const avj_layer* layers, const avj_stack_scratch* sc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AVJEPA_B200_H_ */
