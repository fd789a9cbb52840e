"""Device engine: raw-pointer launches of the C-ABI kernels and the explicit forward/backward
schedule of a pre-LN transformer stack.

PyTorch is used here only for device memory (``torch.empty``), the current CUDA stream and
the tensors handed back to the caller; every arithmetic step is a kernel of
``libavjepa_sm100.so``.  Buffers inside one forward/backward are carved out of two byte
arenas by integer pointer arithmetic (no per-op tensor objects):

* the *saved* arena of a :class:`StackRun` holds every activation the backward needs
  (per layer: x, LN stats, LN outputs, qkv, attention output + LSE, pre/post-GELU), laid out
  layer-major so one ``torch.empty`` covers the whole stack;
* the *scratch* arena is reused stream-ordered for temporaries (gradients flowing between
  kernels, reduction workspaces).

Residual stream and its gradient are fp32; GEMM/attention operands are the compute dtype
(bf16 in production, fp32 in check mode).
"""
import ctypes as C
import os
import threading

import torch

from avjepa_b200 import _cabi
from avjepa_b200._cabi import BF16, F32, GEMM_NN, GEMM_NT, GEMM_TN, Epilogue, RowMap


def _align(n, a=256):
    return (n + a - 1) // a * a


class Mode(object):
    """Compute dtype of GEMM/attention operands."""

    def __init__(self, code):
        self.code = code
        self.torch_dtype = torch.bfloat16 if code == BF16 else torch.float32
        self.size = 2 if code == BF16 else 4

    @staticmethod
    def current():
        """bf16 under ``torch.autocast('cuda', dtype=bfloat16)`` (what the reference train loop
        enters, app/avjepa/train.py:502), fp32 check mode otherwise."""
        if torch.is_autocast_enabled('cuda') and torch.get_autocast_dtype('cuda') == torch.bfloat16:
            return MODE_BF16
        return MODE_F32


MODE_F32 = Mode(F32)
MODE_BF16 = Mode(BF16)


def stream():
    return torch.cuda.current_stream().cuda_stream


def require_cuda(t, what):
    if not t.is_cuda:
        raise _cabi.AvjError(f'{what}: expected a CUDA tensor (device={t.device}); avjepa_b200 has no CPU path')


class _BufferPool(object):
    """Free list of large device buffers, per (device, stream).  The activation arenas of a step are 1-11 GB and
    their size changes with every mask draw; asked for afresh each time, the torch caching allocator ends up calling
    cudaMalloc (a device synchronisation) inside steps.  Arenas therefore return their buffer here when they die
    and a new arena takes the smallest free buffer that fits (stream-ordered reuse, like the caching allocator's own
    rule for one stream); a new buffer is only allocated, 12 % larger than asked, when none fits."""

    def __init__(self):
        self.free = {}
        self.lock = threading.RLock()      # arenas die on the autograd thread while the main thread builds the next ones

    def take(self, nbytes, device):
        key = (device.index, torch.cuda.current_stream(device).cuda_stream)
        with self.lock:
            lst = self.free.setdefault(key, [])
            best = None
            for i, b in enumerate(lst):
                if nbytes <= b.numel() <= 2 * nbytes and (best is None or b.numel() < lst[best].numel()):
                    best = i
            if best is not None:
                return key, lst.pop(best)
        return key, torch.empty(int(nbytes * 1.12) + (1 << 20), dtype=torch.uint8, device=device)

    def give(self, key, buf):
        with self.lock:
            lst = self.free.setdefault(key, [])
            lst.append(buf)
            if len(lst) > 8:                               # keep the largest ones
                lst.sort(key=lambda b: b.numel())
                del lst[0]


_POOL = _BufferPool()
_POOL_MIN = 64 << 20
_POOL_ON = os.environ.get('AVJ_ARENA_POOL', '1') != '0'     # 0: every arena is a fresh torch.empty


class Arena(object):
    """Byte arena over one torch allocation; hands out raw device pointers."""

    def __init__(self, nbytes, device):
        self.nbytes = int(nbytes)
        need = max(self.nbytes, 256) + 256
        self._pool_key = None
        device = torch.device(device)
        if _POOL_ON and need >= _POOL_MIN and device.type == 'cuda':
            self._pool_key, self.buf = _POOL.take(need, device)
        else:
            self.buf = torch.empty(need, dtype=torch.uint8, device=device)
        self.base = _align(self.buf.data_ptr())
        self.off = 0

    def __del__(self):
        try:
            if self._pool_key is not None and self.buf is not None:
                _POOL.give(self._pool_key, self.buf)
        except Exception:                                   # interpreter shutdown
            pass

    def alloc(self, nbytes):
        p = self.base + self.off
        self.off += _align(int(nbytes))
        if self.off > self.nbytes + 256:
            raise _cabi.AvjError(f'arena overflow: need {self.off} of {self.nbytes} bytes')
        return p


class _Scratch(object):
    """Per-device reusable scratch arena (stream-ordered reuse on the current stream)."""

    def __init__(self):
        self.arenas = {}

    def get(self, nbytes, device):
        key = (device.index, torch.cuda.current_stream().cuda_stream)
        a = self.arenas.get(key)
        if a is None or a.nbytes < nbytes:
            a = Arena(int(nbytes * 1.25) + (1 << 20), device)
            self.arenas[key] = a
        a.off = 0
        return a


SCRATCH = _Scratch()


# ----------------------------------------------------------------------------------------------
# thin launch wrappers (all pointers are ints or None)
# ----------------------------------------------------------------------------------------------
def rowmap(rows_per_group=0, group_stride=0, row_offset=0):
    return RowMap(int(rows_per_group), int(group_stride), int(row_offset))


def gemm(mode, layout, A, B_, Cp, M, N, K, lda, ldb, ldc, out_dtype, bias=None, residual=None, pos=None,
         pos_idx=None, pos_rows=0, act=0, pre_out=None, dact_aux=None, accumulate=0, out_map=None):
    ep = Epilogue(bias, residual, pos, pos_idx, int(pos_rows), int(act), pre_out, dact_aux, int(accumulate),
                  int(out_dtype), out_map if out_map is not None else _cabi.IDENTITY)
    _cabi.call('avj_gemm', mode.code, layout, A, B_, Cp, int(M), int(N), int(K), int(lda), int(ldb), int(ldc),
               C.byref(ep), stream())


def patch_embed_tma_enabled():
    """AVJ_PATCH_EMBED_TMA=0: patch matrix (avj_patchify) + bf16 GEMM instead of the im2col-free tf32 kernel."""
    return os.environ.get('AVJ_PATCH_EMBED_TMA', '1') != '0'


def patch_embed(x, idx, w_f32, out, B, Cin, T, H, W, tub, patch, K, D, ldc, bias=None, pos=None, pos_idx=None, pos_rows=0,
                out_map=None):
    """Conv3d / Conv2d patch projection of the kept tokens straight out of the clip (avj_patch_embed)."""
    ep = Epilogue(bias, None, pos, pos_idx, int(pos_rows), 0, None, None, 0, int(_cabi.F32),
                  out_map if out_map is not None else _cabi.IDENTITY)
    _cabi.call('avj_patch_embed', x, idx, w_f32, out, int(B), int(Cin), int(T), int(H), int(W), int(tub), int(patch), int(K),
               int(D), int(ldc), C.byref(ep), stream())


def patch_embed_wgrad(x, idx, dy_bf16, gw, B, Cin, T, H, W, tub, patch, K, D):
    """Conv weight gradient of the patch projection, patch values gathered straight out of the clip (avj_patch_embed_wgrad)."""
    _cabi.call('avj_patch_embed_wgrad', x, idx, dy_bf16, gw, int(B), int(Cin), int(T), int(H), int(W), int(tub), int(patch), int(K), int(D),
               stream())


def layernorm_fwd(x, gamma, beta, y, y_dtype, mean, rstd, rows, D, eps):
    _cabi.call('avj_layernorm_fwd', x, gamma, beta, y, y_dtype, mean, rstd, int(rows), int(D), float(eps), stream())


def layernorm_bwd(dy, dy_dtype, x, gamma, mean, rstd, dres, dx, dx_lp, lp_dtype, dgamma, dbeta, ws, rows, D, dcolsum=None):
    _cabi.call('avj_layernorm_bwd', dy, dy_dtype, x, gamma, mean, rstd, dres, dx, dx_lp, lp_dtype, dgamma, dbeta, dcolsum, ws,
               int(rows), int(D), stream())


def colsum(inp, in_dtype, ld, rmap, out, rows, D, ws):
    _cabi.call('avj_colsum', inp, in_dtype, int(ld), rmap, out, int(rows), int(D), ws, stream())


def copy_rows(inp, in_dtype, ld_in, imap, out, out_dtype, ld_out, omap, rows, D, accumulate=0):
    _cabi.call('avj_copy_rows', inp, in_dtype, int(ld_in), imap, out, out_dtype, int(ld_out), omap, int(rows), int(D),
               int(accumulate), stream())


def memset0(ptr, nbytes):
    _cabi.call('avj_memset_zero', ptr, int(nbytes), stream())


# ----------------------------------------------------------------------------------------------
# parameter views
# ----------------------------------------------------------------------------------------------
class LinearW(object):
    """Pointers for one nn.Linear: weight in compute dtype, fp32 bias, fp32 grad buffers."""
    __slots__ = ('w', 'b', 'gw', 'gb', 'out_f', 'in_f')

    def __init__(self, lin, shadows, mode, want_grad):
        self.out_f, self.in_f = lin.weight.shape
        self.w = shadows.weight_ptr(lin.weight, mode)
        self.b = lin.bias.data_ptr() if lin.bias is not None else None
        self.gw = grad_ptr(lin.weight) if want_grad else None
        self.gb = grad_ptr(lin.bias) if (want_grad and lin.bias is not None) else None


class NormW(object):
    __slots__ = ('w', 'b', 'gw', 'gb', 'eps')

    def __init__(self, ln, want_grad):
        self.w = ln.weight.data_ptr()
        self.b = ln.bias.data_ptr()
        self.eps = ln.eps
        self.gw = grad_ptr(ln.weight) if want_grad else None
        self.gb = grad_ptr(ln.bias) if want_grad else None


class BlockW(object):
    __slots__ = ('n1', 'qkv', 'proj', 'n2', 'fc1', 'fc2')

    def __init__(self, blk, shadows, mode, want_grad):
        self.n1 = NormW(blk.norm1, want_grad)
        self.qkv = LinearW(blk.attn.qkv, shadows, mode, want_grad)
        self.proj = LinearW(blk.attn.proj, shadows, mode, want_grad)
        self.n2 = NormW(blk.norm2, want_grad)
        self.fc1 = LinearW(blk.mlp.fc1, shadows, mode, want_grad)
        self.fc2 = LinearW(blk.mlp.fc2, shadows, mode, want_grad)


_GRAD_TOUCHES = 0


def grad_touch_count():
    """How many gradient pointers have been handed to backward kernels so far (the fused optimizer compares this
    with the value it saw when it last cleared its buffers to decide whether zero_grad() has work to do)."""
    return _GRAD_TOUCHES


def grad_ptr(p):
    """fp32 gradient buffer of a parameter, created zero-filled on first use.  Backward kernels
    ACCUMULATE straight into it (weight-gradient GEMMs with a `C +=` epilogue), so no
    per-parameter gradient temporaries exist."""
    global _GRAD_TOUCHES
    if not p.requires_grad:
        return None
    _GRAD_TOUCHES += 1
    if p.grad is None:
        p.grad = torch.zeros_like(p, memory_format=torch.contiguous_format)
    return p.grad.data_ptr()


class Shadows(object):
    """Compute-dtype copies of the 2-D+ weights of one module tree.

    In fp32 mode the parameter itself is used.  In bf16 mode each weight has a bf16 shadow that
    is refreshed (one cast kernel) whenever the parameter's version counter or storage changed;
    the fused optimizer (:mod:`avjepa_b200.optim`) refreshes shadows itself inside the AdamW
    kernel and marks them clean, so the steady-state step issues no cast kernels at all.
    """

    def __init__(self):
        self.entries = {}     # id(param) -> [version, data_ptr, bf16 tensor]

    def weight_ptr(self, p, mode):
        if mode.code == F32:
            return p.data_ptr()
        e = self.entries.get(id(p))
        if e is None:
            e = [None, None, torch.empty(p.shape, dtype=torch.bfloat16, device=p.device)]
            self.entries[id(p)] = e
        if e[0] != p._version or e[1] != p.data_ptr():
            _cabi.call('avj_cast', p.data_ptr(), e[2].data_ptr(), BF16, p.numel(), stream())
            e[0], e[1] = p._version, p.data_ptr()
        return e[2].data_ptr()

    def adopt(self, p, shadow_tensor):
        """Register an externally maintained shadow (a view into the optimizer's flat bf16 buffer)."""
        self.entries[id(p)] = [p._version, p.data_ptr(), shadow_tensor]

    def mark_clean(self, p):
        e = self.entries.get(id(p))
        if e is not None:
            e[0], e[1] = p._version, p.data_ptr()


# ----------------------------------------------------------------------------------------------
# transformer stack
# ----------------------------------------------------------------------------------------------
class StackRun(object):
    """One forward (and optionally backward) of L pre-LN blocks + final LayerNorm over a token matrix [R, D].
    Restates Block.forward / Attention.forward / MLP.forward of the reference
    (src/models/utils/modules.py:114-120, :61-78, :30-36) as an explicit kernel schedule.

    The token matrix is a VARIABLE-LENGTH batch: `segs` = [(B_0, N_0), (B_1, N_1), ...] groups of B_g sequences of N_g
    tokens each, rows of group g starting at ``row0[g]``.  The reference's MultiMask wrappers run the whole backbone
    once per mask (src/models/utils/multimask.py:37-46,55-71); here both masks travel through ONE schedule -- every
    Linear / LayerNorm / column-sum is a single launch over all rows, attention runs per group."""

    def __init__(self, segs, D, heads, depth, mode, save, device, hidden=None):
        if isinstance(segs, tuple) and len(segs) == 2 and not isinstance(segs[0], (tuple, list)):
            segs = [segs]
        self.segs = [(int(b), int(n)) for b, n in segs]
        if len(self.segs) > _cabi.MAX_SEGMENTS:
            raise _cabi.AvjError(f'{len(self.segs)} sequence groups in one stack (max {_cabi.MAX_SEGMENTS})')
        self.B, self.N = self.segs[0]                       # single-group callers
        self.D, self.H, self.L = D, heads, depth
        self.row0, self.lse0 = [], []
        r = l = 0
        for b, n in self.segs:
            self.row0.append(r)
            self.lse0.append(l)
            r += b * n
            l += b * heads * n
        self.R = r
        self.hd = D // heads
        self.Hd = hidden if hidden is not None else 4 * D
        self.mode, self.save, self.device = mode, save, device
        R, s = self.R, mode.size
        # per-layer saved layout (byte offsets)
        o, lay = 0, {}
        for name, nbytes in (('x', R * D * 4), ('mean1', R * 4), ('rstd1', R * 4), ('h1', R * D * s),
                             ('qkv', R * 3 * D * s), ('o', R * D * s), ('lse', l * 4),
                             ('x1', R * D * 4), ('mean2', R * 4), ('rstd2', R * 4), ('h2', R * D * s),
                             ('pre', R * self.Hd * s), ('act', R * self.Hd * s)):
            lay[name] = o
            o += _align(nbytes)
        self.lay, self.layer_bytes = lay, o
        n_layers_stored = depth if save else 1
        tail = _align(R * D * 4) + 2 * _align(R * 4)            # x_L, final mean, rstd
        self.arena = Arena(n_layers_stored * o + tail + 4096, device)
        self.layer_base = [self.arena.alloc(o) for _ in range(n_layers_stored)]
        self.x_final = self.arena.alloc(R * D * 4)
        self.mean_f = self.arena.alloc(R * 4)
        self.rstd_f = self.arena.alloc(R * 4)

    # -- addressing ------------------------------------------------------------------------
    def slot(self, layer, name):
        base = self.layer_base[layer if self.save else 0]
        return base + self.lay[name]

    def x_in(self, layer):
        """Input residual stream of `layer` (layer == L -> output of the last block).  Without
        saving, the stream ping-pongs between two buffers."""
        if self.save:
            return self.slot(layer, 'x') if layer < self.L else self.x_final
        return (self.slot(0, 'x'), self.x_final)[layer % 2]

    def x_out(self, layer):
        return self.x_in(layer + 1)

    def x0_rows(self, g):
        """Pointer of the first fp32 input row of group g (where the embedding kernels write)."""
        return self.x_in(0) + self.row0[g] * self.D * 4

    def attn_ws_floats(self):
        lib = _cabi.load()
        return max(lib.avj_attention_bwd_ws_floats(b, n, self.H, self.hd) for b, n in self.segs)

    # -- forward ---------------------------------------------------------------------------
    def forward_layer(self, i, w):
        """x_{i+1} = Block_i(x_i): LN1 -> qkv -> attention -> proj(+x) -> LN2 -> fc1/GELU -> fc2(+x1)."""
        m, R, D, Hd, cd, s = self.mode, self.R, self.D, self.Hd, self.mode.code, self.mode.size
        scale = float(self.hd ** -0.5)
        x, x1, xo = self.x_in(i), self.slot(i, 'x1'), self.x_out(i)
        layernorm_fwd(x, w.n1.w, w.n1.b, self.slot(i, 'h1'), cd, self.slot(i, 'mean1'), self.slot(i, 'rstd1'), R, D, w.n1.eps)
        gemm(m, GEMM_NT, self.slot(i, 'h1'), w.qkv.w, self.slot(i, 'qkv'), R, 3 * D, D, D, D, 3 * D, cd, bias=w.qkv.b)
        for (b, n), r0, l0 in zip(self.segs, self.row0, self.lse0):
            _cabi.call('avj_attention_fwd', cd, self.slot(i, 'qkv') + r0 * 3 * D * s, self.slot(i, 'o') + r0 * D * s,
                       self.slot(i, 'lse') + l0 * 4, b, n, self.H, self.hd, scale, stream())
        gemm(m, GEMM_NT, self.slot(i, 'o'), w.proj.w, x1, R, D, D, D, D, D, F32, bias=w.proj.b, residual=x)
        layernorm_fwd(x1, w.n2.w, w.n2.b, self.slot(i, 'h2'), cd, self.slot(i, 'mean2'), self.slot(i, 'rstd2'), R, D, w.n2.eps)
        gemm(m, GEMM_NT, self.slot(i, 'h2'), w.fc1.w, self.slot(i, 'act'), R, Hd, D, D, D, Hd, cd, bias=w.fc1.b,
             act=1, pre_out=self.slot(i, 'pre') if self.save else None)
        gemm(m, GEMM_NT, self.slot(i, 'act'), w.fc2.w, xo, R, D, Hd, Hd, Hd, D, F32, bias=w.fc2.b, residual=x1)

    # -- whole-stack C schedule --------------------------------------------------------------
    def _stack_desc(self):
        d = _cabi.Stack(self.mode.code, self.B, self.N, self.D, self.H, self.Hd, self.L)
        d.n_seg = len(self.segs)
        for g, (b, n) in enumerate(self.segs):
            d.seg_B[g], d.seg_N[g] = b, n
        return d

    def _layer_array(self, blocks):
        """ctypes array of avj_layer: weight pointers + this run's activation slots."""
        arr = (_cabi.Layer * self.L)()
        for i, w in enumerate(blocks):
            l = arr[i]
            for name in ('n1', 'n2'):
                src, dst = getattr(w, name), getattr(l, name)
                dst.w, dst.b, dst.gw, dst.gb, dst.eps = src.w, src.b, src.gw, src.gb, src.eps
            for name in ('qkv', 'proj', 'fc1', 'fc2'):
                src, dst = getattr(w, name), getattr(l, name)
                dst.w, dst.b, dst.gw, dst.gb = src.w, src.b, src.gw, src.gb
            sl = self.slot
            l.x, l.x_out = self.x_in(i), self.x_out(i)
            l.mean1, l.rstd1, l.h1, l.qkv_act, l.o, l.lse = (sl(i, 'mean1'), sl(i, 'rstd1'), sl(i, 'h1'), sl(i, 'qkv'),
                                                            sl(i, 'o'), sl(i, 'lse'))
            l.x1, l.mean2, l.rstd2, l.h2, l.act = sl(i, 'x1'), sl(i, 'mean2'), sl(i, 'rstd2'), sl(i, 'h2'), sl(i, 'act')
            l.pre = sl(i, 'pre') if self.save else None
        return arr

    def _per_group(self, ptrs):
        """Normalise an output / gradient pointer argument: one pointer for all R rows (contiguous) -> [(ptr, row0,
        rows)], or a list with one pointer per group."""
        if isinstance(ptrs, (list, tuple)):
            assert len(ptrs) == len(self.segs)
            return [(p, r0, b * n) for p, r0, (b, n) in zip(ptrs, self.row0, self.segs)]
        return [(ptrs, 0, self.R)]

    def forward(self, blocks, norm, out_ptr, out_dtype):
        """blocks: list[BlockW]; norm: NormW or None.  Writes LN(x_L) to `out_ptr` -- one pointer for all rows, or a
        list with one (contiguous) destination per sequence group."""
        D = self.D
        if self.L > 0:
            desc, arr = self._stack_desc(), self._layer_array(blocks)
            _cabi.call('avj_stack_forward', C.byref(desc), arr, stream())
        self.x_last = self.x_in(self.L)
        for p, r0, rows in self._per_group(out_ptr):
            if rows == 0:
                continue
            if norm is not None:
                layernorm_fwd(self.x_last + r0 * D * 4, norm.w, norm.b, p, out_dtype, self.mean_f + r0 * 4, self.rstd_f + r0 * 4,
                              rows, D, norm.eps)
            else:
                copy_rows(self.x_last + r0 * D * 4, F32, D, _cabi.IDENTITY, p, out_dtype, D, _cabi.IDENTITY, rows, D)

    # -- backward --------------------------------------------------------------------------
    def scratch_bytes(self):
        R, D, Hd, s = self.R, self.D, self.Hd, self.mode.size
        per = 2 * _align(R * D * 4) + _align(R * D * s) + _align(R * Hd * s) + _align(R * 3 * D * s) + 2 * _align(R * D * s)
        ws = max(_cabi.load().avj_layernorm_bwd_ws_floats(R, D), _cabi.load().avj_colsum_ws_floats(R, 3 * D + Hd),
                 self.attn_ws_floats()) * 4
        return per + _align(ws) + (1 << 16)

    def backward(self, blocks, norm, dy_ptr, dy_dtype, sc, layer_events=None):
        """dy: gradient wrt the LN(x_L) output in dy_dtype -- one pointer ([R, D] contiguous) or one per sequence
        group.  `sc` is a scratch Arena with at least scratch_bytes().  Returns the pointer of d x_0 (fp32, in scratch).
        `layer_events`: optional list of L torch.cuda.Event; event i is recorded (by the C schedule, on the
        launch stream) when layer i's parameter gradients are final for this call."""
        assert self.save, 'backward needs a forward run with save=True'
        m, R, D, Hd, cd, s = self.mode, self.R, self.D, self.Hd, self.mode.code, self.mode.size
        dxa = sc.alloc(R * D * 4)          # fp32 residual-gradient ping
        dxb = sc.alloc(R * D * 4)          # pong
        dx_lp = sc.alloc(R * D * s)        # compute-dtype copy feeding the GEMMs
        d_hid = sc.alloc(R * Hd * s)       # d pre-GELU
        d_qkv = sc.alloc(R * 3 * D * s)
        d_h = sc.alloc(R * D * s)          # d LN-output (h2 / h1)
        d_o = sc.alloc(R * D * s)
        lib = _cabi.load()
        ws = sc.alloc(4 * max(lib.avj_layernorm_bwd_ws_floats(R, D), lib.avj_colsum_ws_floats(R, 3 * D + Hd),
                              self.attn_ws_floats()))
        cur, nxt = dxa, dxb
        esz = 4 if dy_dtype == F32 else 2
        for p, r0, rows in self._per_group(dy_ptr):
            if rows == 0:
                continue
            if norm is not None:
                layernorm_bwd(p, dy_dtype, self.x_in(self.L) + r0 * D * 4, norm.w, self.mean_f + r0 * 4, self.rstd_f + r0 * 4, None,
                              cur + r0 * D * 4, dx_lp + r0 * D * s, cd, norm.gw, norm.gb, ws, rows, D)
            else:
                copy_rows(p, dy_dtype, D, _cabi.IDENTITY, cur + r0 * D * 4, F32, D, _cabi.IDENTITY, rows, D)
                copy_rows(p, dy_dtype, D, _cabi.IDENTITY, dx_lp + r0 * D * s, cd, D, _cabi.IDENTITY, rows, D)
        del esz
        if self.L > 0:
            desc, arr = self._stack_desc(), self._layer_array(blocks)
            ev_arr = None
            if layer_events is not None:
                assert len(layer_events) == self.L
                ev_arr = (C.c_void_p * self.L)(*[int(e.cuda_event) for e in layer_events])
            scs = _cabi.StackScratch(cur, nxt, dx_lp, d_hid, d_qkv, d_h, d_o, ws, ev_arr)
            _cabi.call('avj_stack_backward', C.byref(desc), arr, C.byref(scs), stream())
        return cur
