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
        """Gradients live in flat buffers that the step kernel already zeroed; keep the views."""
        self.ensure_built()
        if self._dirty_grads:
            for r in self._ranges:
                if r['g'] is not None:
                    r['g'].zero_()
            self._dirty_grads = False

    _dirty_grads = True

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
    def step(self, closure=None, ema_momentum=None, inv_loss_scale=1.0, coef_by_group=None):
        """One AdamW step on every group; EMA for paired ranges when `ema_momentum` is given.
        `coef_by_group`: optional {group index -> device float tensor} gradient multipliers
        (unscale x clip), else `inv_loss_scale` is applied."""
        self.ensure_built()
        self._step += 1
        for r in self._ranges:
            a = AdamWArgs()
            a.p = r['flat'].data_ptr()
            a.n = r['flat'].numel()
            a.target = r['k'].data_ptr() if (r['k'] is not None and ema_momentum is not None) else None
            a.target_lp = r['klp'].data_ptr() if (r['klp'] is not None and ema_momentum is not None) else None
            a.ema_m = float(ema_momentum) if ema_momentum is not None else 1.0
            if r['frozen']:
                a.skip_update = 1
                a.g = a.m = a.v = a.p_lp = None
                a.lr = a.wd = 0.0
                a.beta1, a.beta2, a.eps, a.step = 0.9, 0.999, 1e-8, self._step
                if a.target is None:
                    continue
            else:
                g = self.param_groups[r['group']]
                a.g, a.m, a.v = r['g'].data_ptr(), r['m'].data_ptr(), r['v'].data_ptr()
                a.p_lp = r['lp'].data_ptr()
                a.lr, a.wd = float(g['lr']), float(g['weight_decay'])
                a.beta1, a.beta2 = float(g['betas'][0]), float(g['betas'][1])
                a.eps = float(g['eps'])
                a.step = self._step
                a.skip_update = 0
                a.zero_grad = 1
                if coef_by_group is not None and r['group'] in coef_by_group:
                    a.scale_ptr = coef_by_group[r['group']].data_ptr()
                elif inv_loss_scale != 1.0:
                    self._coef.fill_(inv_loss_scale)
                    a.scale_ptr = self._coef.data_ptr()
                else:
                    a.scale_ptr = None
            _cabi.call('avj_adamw_ema_step', C.byref(a), engine.stream())
        self._step_t += 1
        self._dirty_grads = False
        return None

    def load_state_dict(self, state_dict):
        """Accepts a torch.optim.AdamW state dict (reference checkpoints, app/avjepa/utils.py:63);
        moments are copied into the flat buffers on the next (lazy) rebuild."""
        super().load_state_dict(state_dict)
        steps = [float(st['step']) for st in self.state.values() if 'step' in st]
        self._step = int(max(steps)) if steps else 0
        self._step_t = torch.tensor(float(self._step))
        self._flat = None

    def mark_grads_dirty(self):
        self._dirty_grads = True
