// Whole-stack kernel schedules: the forward and backward of L pre-LN transformer blocks issued from
// ONE C call (host-side native runtime; the per-kernel entry points stay available for tests).
// Restates Block.forward / Attention.forward / MLP.forward of the reference
// (src/models/utils/modules.py:114-120, :61-78, :30-36) as an explicit launch sequence:
//
//   fwd : LN1 -> qkv GEMM(+bias) -> attention -> proj GEMM(+bias,+x) -> LN2 -> fc1 GEMM(+bias,GELU,
//         pre-activation stash) -> fc2 GEMM(+bias,+x1)
//   bwd : the same chain reversed; weight gradients accumulate in place (C += epilogue), bias
//         gradients are column sums, LayerNorm backward fuses the residual-gradient add and emits the
//         compute-dtype copy that feeds the next dgrad/wgrad GEMMs.
#include "common.cuh"

#define RC(expr)            \
  do {                      \
    int rc__ = (expr);      \
    if (rc__) return rc__;  \
  } while (0)

static inline avj_epilogue epi(int out_dtype) {
  avj_epilogue e;
  memset(&e, 0, sizeof(e));
  e.out_dtype = out_dtype;
  return e;
}

// groups of equal-length sequences inside the token matrix (avj_stack.n_seg / seg_B / seg_N)
struct Segs {
  int n; int B[AVJ_MAX_SEGMENTS]; int N[AVJ_MAX_SEGMENTS]; int64_t row0[AVJ_MAX_SEGMENTS]; int64_t lse0[AVJ_MAX_SEGMENTS]; int64_t R;
};
static int make_segs(const avj_stack* s, Segs* g) {
  g->n = s->n_seg > 0 ? s->n_seg : 1;
  AVJ_CHECK(g->n <= AVJ_MAX_SEGMENTS, "avj_stack: %d segments (max %d)", g->n, AVJ_MAX_SEGMENTS);
  int64_t r = 0, l = 0;
  for (int i = 0; i < g->n; ++i) {
    g->B[i] = s->n_seg > 0 ? s->seg_B[i] : s->B;
    g->N[i] = s->n_seg > 0 ? s->seg_N[i] : s->N;
    AVJ_CHECK(g->B[i] >= 0 && g->N[i] >= 0, "avj_stack: negative segment shape");
    g->row0[i] = r; g->lse0[i] = l;
    r += (int64_t)g->B[i] * g->N[i];
    l += (int64_t)g->B[i] * s->H * g->N[i];
  }
  AVJ_CHECK(r < (1ll << 31), "avj_stack: %lld rows overflow int32", (long long)r);
  g->R = r;
  return 0;
}
static inline size_t esize(int dtype) { return dtype == AVJ_BF16 ? 2 : 4; }

extern "C" int avj_stack_forward(const avj_stack* s, const avj_layer* L, void* stream) {
  AVJ_CHECK(s && L, "avj_stack_forward: NULL argument");
  Segs g;
  RC(make_segs(s, &g));
  const int R = (int)g.R, D = s->D, Hd = s->hidden, cd = s->dtype;
  AVJ_CHECK(s->H > 0 && D % s->H == 0, "avj_stack_forward: D=%d not divisible by heads=%d", D, s->H);
  const int hd = D / s->H;
  const float scale = 1.0f / sqrtf((float)hd);
  const size_t es = esize(cd);
  for (int i = 0; i < s->L; ++i) {
    const avj_layer& w = L[i];
    RC(avj_layernorm_fwd(w.x, w.n1.w, w.n1.b, w.h1, cd, w.mean1, w.rstd1, R, D, w.n1.eps, stream));
    avj_epilogue e = epi(cd);
    e.bias = w.qkv.b;
    RC(avj_gemm(cd, AVJ_GEMM_NT, w.h1, w.qkv.w, w.qkv_act, R, 3 * D, D, D, D, 3 * D, &e, stream));
    for (int k = 0; k < g.n; ++k)
      RC(avj_attention_fwd(cd, (const char*)w.qkv_act + g.row0[k] * 3 * D * es, (char*)w.o + g.row0[k] * D * es, w.lse + g.lse0[k],
                           g.B[k], g.N[k], s->H, hd, scale, stream));
    e = epi(AVJ_F32);
    e.bias = w.proj.b; e.residual = w.x;
    RC(avj_gemm(cd, AVJ_GEMM_NT, w.o, w.proj.w, w.x1, R, D, D, D, D, D, &e, stream));
    RC(avj_layernorm_fwd(w.x1, w.n2.w, w.n2.b, w.h2, cd, w.mean2, w.rstd2, R, D, w.n2.eps, stream));
    e = epi(cd);
    e.bias = w.fc1.b; e.act = 1; e.pre_out = w.pre;
    RC(avj_gemm(cd, AVJ_GEMM_NT, w.h2, w.fc1.w, w.act, R, Hd, D, D, D, Hd, &e, stream));
    e = epi(AVJ_F32);
    e.bias = w.fc2.b; e.residual = w.x1;
    RC(avj_gemm(cd, AVJ_GEMM_NT, w.act, w.fc2.w, w.x_out, R, D, Hd, Hd, Hd, D, &e, stream));
  }
  return 0;
}

extern "C" int avj_stack_backward(const avj_stack* s, const avj_layer* L, const avj_stack_scratch* sc, void* stream) {
  AVJ_CHECK(s && L && sc, "avj_stack_backward: NULL argument");
  Segs g;
  RC(make_segs(s, &g));
  const int R = (int)g.R, D = s->D, Hd = s->hidden, cd = s->dtype;
  const int hd = D / s->H;
  const float scale = 1.0f / sqrtf((float)hd);
  const size_t es = esize(cd);
  const avj_rowmap ident = {0, 0, 0};
  float* cur = sc->dxa;
  float* nxt = sc->dxb;
  for (int i = s->L - 1; i >= 0; --i) {
    const avj_layer& w = L[i];
    AVJ_CHECK(w.pre != nullptr, "avj_stack_backward: layer %d has no saved pre-activation (forward ran without save)", i);
    // Bias gradients: fc2.gb / proj.gb are column sums of the residual-stream gradient and come out of the
    // LayerNorm backward that PRODUCES that gradient (dcolsum); only the top layer's fc2.gb, whose input was
    // produced outside this call, needs its own pass.  fc1.gb / qkv.gb share one launch per layer.
    // ---- MLP: x2 = x1 + fc2(gelu(fc1(LN2(x1))))
    if (w.fc2.gb && i == s->L - 1) RC(avj_colsum(sc->dx_lp, cd, D, ident, w.fc2.gb, R, D, sc->ws, stream));
    avj_epilogue e;
    if (w.fc2.gw) {
      e = epi(AVJ_F32); e.accumulate = 1;
      RC(avj_gemm(cd, AVJ_GEMM_TN, sc->dx_lp, w.act, w.fc2.gw, D, Hd, R, D, Hd, Hd, &e, stream));
    }
    e = epi(cd); e.dact_aux = w.pre;
    RC(avj_gemm(cd, AVJ_GEMM_NN, sc->dx_lp, w.fc2.w, sc->d_hid, R, Hd, D, D, Hd, Hd, &e, stream));
    if (w.fc1.gw) {
      e = epi(AVJ_F32); e.accumulate = 1;
      RC(avj_gemm(cd, AVJ_GEMM_TN, sc->d_hid, w.h2, w.fc1.gw, Hd, D, R, Hd, D, D, &e, stream));
    }
    e = epi(cd);
    RC(avj_gemm(cd, AVJ_GEMM_NN, sc->d_hid, w.fc1.w, sc->d_h, R, D, Hd, Hd, D, D, &e, stream));
    RC(avj_layernorm_bwd(sc->d_h, cd, w.x1, w.n2.w, w.mean2, w.rstd2, cur, nxt, sc->dx_lp, cd, w.n2.gw, w.n2.gb, w.proj.gb,
                         sc->ws, R, D, stream));
    { float* t = cur; cur = nxt; nxt = t; }
    // ---- attention: x1 = x + proj(attn(qkv(LN1(x))))
    if (w.proj.gw) {
      e = epi(AVJ_F32); e.accumulate = 1;
      RC(avj_gemm(cd, AVJ_GEMM_TN, sc->dx_lp, w.o, w.proj.gw, D, D, R, D, D, D, &e, stream));
    }
    e = epi(cd);
    RC(avj_gemm(cd, AVJ_GEMM_NN, sc->dx_lp, w.proj.w, sc->d_o, R, D, D, D, D, D, &e, stream));
    for (int k = 0; k < g.n; ++k)
      RC(avj_attention_bwd(cd, (const char*)w.qkv_act + g.row0[k] * 3 * D * es, (const char*)w.o + g.row0[k] * D * es,
                           (const char*)sc->d_o + g.row0[k] * D * es, w.lse + g.lse0[k], (char*)sc->d_qkv + g.row0[k] * 3 * D * es,
                           sc->ws, g.B[k], g.N[k], s->H, hd, scale, stream));
    if (w.qkv.gb && w.fc1.gb) {
      RC(avj_colsum2(sc->d_qkv, 3 * D, 3 * D, w.qkv.gb, sc->d_hid, Hd, Hd, w.fc1.gb, cd, R, sc->ws, stream));
    } else {
      if (w.qkv.gb) RC(avj_colsum(sc->d_qkv, cd, 3 * D, ident, w.qkv.gb, R, 3 * D, sc->ws, stream));
      if (w.fc1.gb) RC(avj_colsum(sc->d_hid, cd, Hd, ident, w.fc1.gb, R, Hd, sc->ws, stream));
    }
    if (w.qkv.gw) {
      e = epi(AVJ_F32); e.accumulate = 1;
      RC(avj_gemm(cd, AVJ_GEMM_TN, sc->d_qkv, w.h1, w.qkv.gw, 3 * D, D, R, 3 * D, D, D, &e, stream));
    }
    e = epi(cd);
    RC(avj_gemm(cd, AVJ_GEMM_NN, sc->d_qkv, w.qkv.w, sc->d_h, R, D, 3 * D, 3 * D, D, D, &e, stream));
    RC(avj_layernorm_bwd(sc->d_h, cd, w.x, w.n1.w, w.mean1, w.rstd1, cur, nxt, sc->dx_lp, cd, w.n1.gw, w.n1.gb,
                         i > 0 ? L[i - 1].fc2.gb : nullptr, sc->ws, R, D, stream));
    { float* t = cur; cur = nxt; nxt = t; }
    // layer i's weight / bias / LayerNorm gradients are complete (fc2.gb of layer i was written by layer i+1)
    if (sc->layer_done && sc->layer_done[i])
      AVJ_CUDA(cudaEventRecord(reinterpret_cast<cudaEvent_t>(sc->layer_done[i]), as_stream(stream)));
  }
  // two swaps per layer: the result is back in dxa
  return 0;
}
