"""Fused AdamW + EMA optimizer over flat parameter buffers.

``FusedAdamWEMA`` is a ``torch.optim.Optimizer`` with the same ``param_groups`` semantics as
the ``torch.optim.AdamW`` the reference builds in ``app/avjepa/utils.py:243-270`` (so the
reference's LR/WD schedulers, ``state_dict`` consumers and ``adamw_logger`` keep working), but

* every parameter of a group is a view into ONE contiguous fp32 buffer, with matching flat
  buffers for grad / exp_avg / exp_avg_sq -- one kernel launch per group instead of one
  foreach-list per op;
* the kernel also unscales / clips the gradient, zeroes it, writes the bf16 shadow weights the
  next forward will read, and -- for groups that have an EMA twin (the target encoder) --
  performs ``k = m*k + (1-m)*q`` (``app/avjepa/train.py:534-537``) and writes the target's
  bf16 shadow in the same pass;
* the global grad norm for clipping is one multi-block reduction, consumed on the device (no
  host sync).

Algorithmic HBM bytes per parameter: 16 read + 12 write (+4 grad zero, +8 EMA, +4 shadows).

Differences from ``torch.optim.AdamW`` that a caller can observe (documented, covered by tests):

* the kernel updates EVERY trainable element of a range each step, so a parameter that received no gradient still
  has weight decay and moment decay applied (torch skips parameters whose ``.grad`` is ``None``).  In the AV-JEPA
  step every trainable parameter receives a gradient every iteration (both mask-token sets are used, both
  modalities are embedded), so the two rules coincide on the path this package serves;
* ``step()`` leaves the gradients in place by default -- the reference loop reads them AFTER the step
  (``grad_logger``, ``app/avjepa/train.py:526-529``) and then calls ``zero_grad()``.  ``step(zero_grads=True)``
  (what :class:`~avjepa_b200.app.avjepa.train.TrainStep` uses) zeroes them inside the same kernel pass instead;
* ``zero_grad()`` keeps the flat views (never sets ``.grad`` to ``None``) and skips the memset only when the last
  ``step(zero_grads=True)`` already cleared the buffers and no backward has touched them since
  (:func:`avjepa_b200.engine.grad_ptr` counts every hand-out of a gradient pointer).
"""
import math

import torch

from avjepa_b200 import _cabi, engine
from avjepa_b200._cabi import AdamWArgs

import ctypes as C


def _flatten_into(params, pad_to=8):
    """Re-home `params` into one contiguous fp32 buffer; returns (flat, offsets)."""
    dev = params[0].device
    total, offs = 0, []
    for p in params:
        offs.append(total)
        total += (p.numel() + pad_to - 1) // pad_to * pad_to
    flat = torch.zeros(total, dtype=torch.float32, device=dev)
    for p, o in zip(params, offs):
        flat[o:o + p.numel()].copy_(p.data.reshape(-1))
        p.data = flat[o:o + p.numel()].view(p.shape)
    return flat, offs


class FusedAdamWEMA(torch.optim.Optimizer):

    def __init__(self, param_groups, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(param_groups, defaults)
        self._flat = None
        self._ema_pairs = {}        # id(param) -> target param
        self._shadow_owners = []    # (Shadows, params) to mark clean after a step
        self._step = 0
        self._step_t = torch.tensor(0.0)

    # ------------------------------------------------------------------ wiring
    def attach_ema(self, online_module, target_module):
        """Pair every parameter of `online_module` with the same-named parameter of
        `target_module`; the step then performs the momentum update too.  Frozen parameters
        (the sincos tables) are EMA-ed as well, exactly like the reference loop."""
        tgt = dict(target_module.named_parameters())
        self._ema_named = [(p, tgt[n]) for n, p in online_module.named_parameters()]
        self._flat = None

    def attach_shadows(self, module, shadows):
        self._shadow_owners.append((module, shadows))
        self._flat = None

    def _build(self):
        """Flatten every group (and the EMA twins, in the same order) once."""
        dev = self.param_groups[0]['params'][0].device
        if dev.type != 'cuda':
            raise _cabi.AvjError('FusedAdamWEMA needs CUDA parameters; avjepa_b200 has no CPU path')
        ema = {id(p): k for p, k in getattr(self, '_ema_named', [])}
        shadow_of = {}
        for module, sh in self._shadow_owners:
            for p in module.parameters():
                shadow_of[id(p)] = sh
        self._ranges = []
        grouped = set()
        for gi, g in enumerate(self.param_groups):
            ps = [p for p in g['params']]
            if not ps:
                self._ranges.append(None)
                continue
            # trainable params of this group that have / do not have an EMA twin are kept in
            # separate contiguous ranges so each launch has uniform arguments
            for has_ema in (True, False):
                sub = [p for p in ps if (id(p) in ema) == has_ema and p.requires_grad]
                if not sub:
                    continue
                flat, offs = _flatten_into(sub)
                gflat = torch.zeros_like(flat)
                m = torch.zeros_like(flat)
                v = torch.zeros_like(flat)
                lp = torch.empty(flat.numel(), dtype=torch.bfloat16, device=dev)
                _cabi.call('avj_cast', flat.data_ptr(), lp.data_ptr(), _cabi.BF16, flat.numel(), engine.stream())
                kflat = klp = None
                if has_ema:
                    twins = [ema[id(p)] for p in sub]
                    kflat, koffs = _flatten_into(twins)
                    assert koffs == offs
                    klp = torch.empty(flat.numel(), dtype=torch.bfloat16, device=dev)
                    _cabi.call('avj_cast', kflat.data_ptr(), klp.data_ptr(), _cabi.BF16, kflat.numel(), engine.stream())
                for p, o in zip(sub, offs):
                    n = p.numel()
                    old = p.grad
                    p.grad = gflat[o:o + n].view(p.shape)
                    if old is not None:
                        p.grad.copy_(old)
                    st = self.state[p]
                    st['step'] = self._step_t          # one shared counter tensor for all parameters
                    prev_m, prev_v = st.get('exp_avg'), st.get('exp_avg_sq')
                    st['exp_avg'] = m[o:o + n].view(p.shape)
                    st['exp_avg_sq'] = v[o:o + n].view(p.shape)
                    if prev_m is not None:
                        st['exp_avg'].copy_(prev_m)
                        st['exp_avg_sq'].copy_(prev_v)
                    if p.dim() >= 2 and id(p) in shadow_of:
                        shadow_of[id(p)].adopt(p, lp[o:o + n].view(p.shape))
                    grouped.add(id(p))
                self._ranges.append(dict(group=gi, flat=flat, g=gflat, m=m, v=v, lp=lp, k=kflat, klp=klp,
                                         params=sub, offs=offs, frozen=False))
                if has_ema:
                    self._adopt_target_shadows(twins, klp, offs)
        # frozen (requires_grad=False) parameters with an EMA twin: EMA only
        frozen = [p for p, _ in getattr(self, '_ema_named', []) if id(p) not in grouped]
        if frozen:
            flat, offs = _flatten_into(frozen)
            kflat, _ = _flatten_into([ema[id(p)] for p in frozen])
            self._ranges.append(dict(group=None, flat=flat, g=None, m=None, v=None, lp=None, k=kflat, klp=None,
                                     params=frozen, offs=offs, frozen=True))
        self._ranges = [r for r in self._ranges if r is not None]
        self._flat = True
        self._sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
        self._coef = torch.ones(1, dtype=torch.float32, device=dev)
        self._coef_val = 1.0
        self._ws = torch.empty(int(_cabi.load().avj_sumsq_ws_floats(0)), dtype=torch.float32, device=dev)

    def _adopt_target_shadows(self, twins, klp, offs):
        for module, sh in self._shadow_owners:
            ids = {id(p) for p in module.parameters()}
            for p, o in zip(twins, offs):
                if id(p) in ids and p.dim() >= 2:
                    sh.adopt(p, klp[o:o + p.numel()].view(p.shape))

    def ensure_built(self):
        if not self._flat:
            self._build()

    # ------------------------------------------------------------------ grads
    def zero_grad(self, set_to_none=False):
        """Zero the flat gradient buffers (the per-parameter ``.grad`` views stay).  A no-op only when the last
        ``step(zero_grads=True)`` cleared them in-kernel and no backward handed out a gradient pointer since."""
        self.ensure_built()
        if self._clean_at == engine.grad_touch_count():
            return
        for r in self._ranges:
            if r['g'] is not None:
                r['g'].zero_()
        self._clean_at = engine.grad_touch_count()

    _clean_at = -1          # engine.grad_touch_count() at the moment the buffers were last known to be all-zero

    def flat_grads(self):
        self.ensure_built()
        return [r['g'] for r in self._ranges if r['g'] is not None]

    def grad_norm_sq(self, group_filter=None):
        """Sum of squared gradients over the selected ranges, as a device scalar (no sync)."""
        self.ensure_built()
        total = torch.zeros(1, dtype=torch.float32, device=self._sumsq.device)
        for r in self._ranges:
            if r['g'] is None or (group_filter is not None and not group_filter(r)):
                continue
            _cabi.call('avj_sumsq', r['g'].data_ptr(), r['g'].numel(), self._sumsq.data_ptr(), self._ws.data_ptr(),
                       engine.stream())
            total += self._sumsq
        return total

    # ------------------------------------------------------------------ step
    @torch.no_grad()
    def step(self, closure=None, ema_momentum=None, inv_loss_scale=1.0, coef_by_group=None, zero_grads=False):
        """One AdamW step on every group; EMA for paired ranges when `ema_momentum` is given.
        `coef_by_group`: optional {group index -> device float tensor} gradient multipliers
        (unscale x clip), else `inv_loss_scale` is applied.  `zero_grads`: clear the gradients in the same pass."""
        self.begin_step()
        for r in self._ranges:
            self.step_interval(r, 0, r['flat'].numel(), ema_momentum=ema_momentum, inv_loss_scale=inv_loss_scale,
                               coef_by_group=coef_by_group, zero_grads=zero_grads)
        self.end_step(zero_grads)
        return None

    # The step in pieces: begin_step(); step_interval(range, lo, hi) for a partition of every range; end_step().
    # The data-parallel step uses this to update each gradient interval as soon as ITS all-reduce has finished while the
    # all-reduces of later intervals are still in flight (dist.GradSync.finish_pipelined).
    def begin_step(self):
        self.ensure_built()
        self._step += 1

    def range_of_grad(self, g):
        for r in self._ranges:
            if r['g'] is g:
                return r
        raise KeyError('not a flat gradient buffer of this optimizer')

    def step_interval(self, r, lo, hi, ema_momentum=None, inv_loss_scale=1.0, coef_by_group=None, zero_grads=False):
        """AdamW (+ EMA, shadows, gradient zeroing) on elements [lo, hi) of range `r`; lo and hi are multiples of 8."""
        n = int(hi) - int(lo)
        if n <= 0:
            return
        o4, o2 = 4 * int(lo), 2 * int(lo)
        a = AdamWArgs()
        a.p = r['flat'].data_ptr() + o4
        a.n = n
        a.target = r['k'].data_ptr() + o4 if (r['k'] is not None and ema_momentum is not None) else None
        a.target_lp = r['klp'].data_ptr() + o2 if (r['klp'] is not None and ema_momentum is not None) else None
        a.ema_m = float(ema_momentum) if ema_momentum is not None else 1.0
        if r['frozen']:
            a.skip_update = 1
            a.g = a.m = a.v = a.p_lp = None
            a.lr = a.wd = 0.0
            a.beta1, a.beta2, a.eps, a.step = 0.9, 0.999, 1e-8, self._step
            if a.target is None:
                return
        else:
            g = self.param_groups[r['group']]
            a.g, a.m, a.v = r['g'].data_ptr() + o4, r['m'].data_ptr() + o4, r['v'].data_ptr() + o4
            a.p_lp = r['lp'].data_ptr() + o2
            a.lr, a.wd = float(g['lr']), float(g['weight_decay'])
            a.beta1, a.beta2 = float(g['betas'][0]), float(g['betas'][1])
            a.eps = float(g['eps'])
            a.step = self._step
            a.skip_update = 0
            a.zero_grad = 1 if zero_grads else 0
            if coef_by_group is not None and r['group'] in coef_by_group:
                a.scale_ptr = coef_by_group[r['group']].data_ptr()
            elif inv_loss_scale != 1.0:
                if self._coef_val != inv_loss_scale:
                    self._coef.fill_(inv_loss_scale)
                    self._coef_val = inv_loss_scale
                a.scale_ptr = self._coef.data_ptr()
            else:
                a.scale_ptr = None
        _cabi.call('avj_adamw_ema_step', C.byref(a), engine.stream())

    _coef_val = None

    def end_step(self, zero_grads=False):
        self._step_t += 1
        self._clean_at = engine.grad_touch_count() if zero_grads else -1

    def scale_grads(self, factor):
        """Multiply every flat gradient buffer by `factor` (GradScaler.unscale_ semantics for a scale != 1)."""
        self.ensure_built()
        for r in self._ranges:
            if r['g'] is not None:
                r['g'].mul_(float(factor))

    # ------------------------------------------------------------------ per-parameter statistics (logging)
    def _segments(self, r):
        """Device int64 offsets [n_params + 1] of range r (cached)."""
        seg = r.get('seg')
        if seg is None:
            offs = list(r['offs']) + [r['flat'].numel()]
            seg = torch.tensor(offs, dtype=torch.int64, device=r['flat'].device)
            r['seg'] = seg
        return seg

    def segment_stats(self, which, group_filter=None):
        """One pass per flat buffer: fp64 device vector with one entry per parameter of the selected ranges --
        sum of squares of the gradient (`which='g'`) or sum of |.| of a moment buffer (`'m'`, `'v'`).
        Returns (names-aligned parameter list, device tensor).  No host synchronisation."""
        self.ensure_built()
        mode = 0 if which == 'g' else 1
        ps, outs = [], []
        for r in self._ranges:
            buf = r.get(which)
            if buf is None or (group_filter is not None and not group_filter(r)):
                continue
            out = torch.zeros(len(r['params']), dtype=torch.float64, device=buf.device)
            _cabi.call('avj_segment_stats', buf.data_ptr(), self._segments(r).data_ptr(), len(r['params']), buf.numel(), mode,
                       out.data_ptr(), engine.stream())
            ps += r['params']
            outs.append(out)
        if not outs:
            return [], None
        return ps, torch.cat(outs)

    def state_dict(self):
        """torch.optim-compatible state: every parameter gets its OWN `step` tensor and contiguous clones of its
        moments (not views into the flat buffers, not one shared counter), so the dict loads into a stock
        ``torch.optim.AdamW`` -- the reference's optimizer -- and steps correctly there."""
        sd = super().state_dict()
        state = {}
        for k, st in sd['state'].items():             # the packed dicts alias self.state[p]: build fresh ones
            st = dict(st)
            if 'step' in st:
                st['step'] = torch.tensor(float(self._step))
            for name in ('exp_avg', 'exp_avg_sq'):
                if torch.is_tensor(st.get(name)):
                    st[name] = st[name].detach().clone().contiguous()
            state[k] = st
        sd['state'] = state
        return sd

    def load_state_dict(self, state_dict):
        """Accepts a torch.optim.AdamW state dict (reference checkpoints, app/avjepa/utils.py:63);
        moments are copied into the flat buffers on the next (lazy) rebuild."""
        super().load_state_dict(state_dict)
        steps = [float(st['step']) for st in self.state.values() if 'step' in st]
        self._step = int(max(steps)) if steps else 0
        self._step_t = torch.tensor(float(self._step))
        self._flat = None

    def mark_grads_dirty(self):
        """Kept for callers of the round-1 API: forces the next zero_grad() to clear the buffers."""
        self._clean_at = -1
