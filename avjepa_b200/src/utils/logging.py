"""Timing / CSV / running-statistics helpers used around the step (host side).

Same public surface as the reference's ``src/utils/logging.py`` -- ``gpu_timer :14-31``, ``get_logger :38-41``,
``CSVLogger :44-62``, ``AverageMeter :65-88``, ``grad_logger :91-105``, ``adamw_logger :108-118`` -- so the
train-loop call sites keep working, written independently.  ``grad_logger`` / ``adamw_logger`` return the same
objects as the reference but never issue one blocking ``float()`` per tensor (~1000 device syncs per step at
ViT-L in the reference): with the fused optimizer they are ONE segment-reduction kernel per flat buffer and one
device-to-host copy (:func:`device_param_stats`); for any other optimizer the per-tensor norms are stacked on
the device and cross in a single transfer.
"""
import logging
import math
import sys

import torch

_FMT = '[%(levelname)-8s][%(asctime)s][%(funcName)-25s] %(message)s'
_DATEFMT = '%Y-%m-%d %H:%M:%S'


def gpu_timer(closure, log_timings=True):
    """Run ``closure()`` and measure it with a pair of CUDA events on the current stream.
    Returns ``(closure result, elapsed milliseconds)``; -1 ms when timing is off or CUDA is absent."""
    if not (log_timings and torch.cuda.is_available()):
        return closure(), -1.
    tic, toc = (torch.cuda.Event(enable_timing=True) for _ in range(2))
    tic.record()
    result = closure()
    toc.record()
    toc.synchronize()
    return result, tic.elapsed_time(toc)


def get_logger(name=None, force=False):
    logging.basicConfig(stream=sys.stdout, level=logging.INFO, format=_FMT, datefmt=_DATEFMT, force=force)
    return logging.getLogger(name=name)


class CSVLogger(object):
    """Appends one formatted row per ``log`` call; columns are given as ``(printf format, name)`` pairs."""

    def __init__(self, fname, *columns):
        self.fname = fname
        self.formats = [fmt for fmt, _ in columns]
        self._write(','.join(name for _, name in columns))

    def _write(self, line):
        with open(self.fname, 'a+') as f:
            f.write(line + '\n')

    def log(self, *values):
        self._write(','.join(fmt % v for fmt, v in zip(self.formats, values)))


class AverageMeter(object):
    """Running value / mean / extrema (attributes ``val avg max min sum count``, like the reference's)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0
        self.max, self.min = -math.inf, math.inf

    def update(self, val, n=1):
        self.val = val
        if isinstance(val, (int, float)):       # the reference tolerates non-comparable values silently
            self.max = val if val > self.max else self.max
            self.min = val if val < self.min else self.min
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


def _is_weight(name, p):
    """grad_logger's filter: matrices only -- no biases, no 1-D parameters."""
    return not (name.endswith('.bias') or p.dim() == 1)


def _grad_stats_from(names, norms):
    stats = AverageMeter()
    first = last = None
    for n, g in zip(names, norms):
        stats.update(g)
        if 'qkv' in n:
            last = g
            first = g if first is None else first
    if first is None or last is None:
        first = last = 0.
    stats.first_layer, stats.last_layer = first, last
    return stats


def grad_logger(named_params):
    """Per-weight gradient norms: meter over all weight matrices plus the first / last ``qkv`` norm."""
    names, norms = [], []
    for n, p in named_params:
        if p.grad is not None and _is_weight(n, p):
            names.append(n)
            norms.append(torch.linalg.vector_norm(p.grad.detach()))
    host = torch.stack(norms).tolist() if norms else []          # one D2H transfer for all tensors
    return _grad_stats_from(names, host)


def adamw_logger(optimizer):
    """Mean magnitude of Adam's first / second moment per parameter, as two meters."""
    first, second = AverageMeter(), AverageMeter()
    ranged = getattr(optimizer, 'segment_stats', None)
    if ranged is not None and getattr(optimizer, '_flat', None):
        pm, sm = optimizer.segment_stats('m')
        pv, sv = optimizer.segment_stats('v')
        if sm is not None:
            numel = torch.tensor([p.numel() for p in pm], dtype=torch.float64)
            host = torch.stack([sm, sv]).cpu() / numel
            for a, b in zip(host[0].tolist(), host[1].tolist()):
                first.update(a)
                second.update(b)
        return {'exp_avg': first, 'exp_avg_sq': second}
    state = optimizer.state_dict().get('state')
    ms = [s['exp_avg'].abs().mean() for s in state.values() if s.get('exp_avg') is not None]
    vs = [s['exp_avg_sq'].abs().mean() for s in state.values() if s.get('exp_avg_sq') is not None]
    if ms:
        for a, b in zip(torch.stack(ms).tolist(), torch.stack(vs).tolist()):
            first.update(a)
            second.update(b)
    return {'exp_avg': first, 'exp_avg_sq': second}


class DeviceParamStats(object):
    """Everything ``grad_logger(encoder) / grad_logger(predictor) / adamw_logger(optimizer)`` report, computed on the
    device from the fused optimizer's flat buffers and staged to pinned host memory with one asynchronous copy.

    Usage inside a step (see :class:`avjepa_b200.app.avjepa.train.TrainStep`)::

        stats.capture_grads(opt, coef)      # BEFORE the fused step zeroes the gradients
        opt.step(..., zero_grads=True)
        stats.capture_moments(opt)          # AFTER the step, like the reference (train.py:526-531)
        stats.stage(side_stream)            # one D2H copy, no host wait
        ...
        enc_stats, pred_stats, optim_stats = stats.collect(named_enc, named_pred)   # waits for the copy only
    """

    def __init__(self):
        self._g = self._m = self._v = None
        self._host = self._event = None
        self._layout = None

    def capture_grads(self, opt, coef_by_group=None, scale=1.0):
        """Sum of squares of every parameter's gradient; `coef_by_group` ({group -> device scalar}, the unscale x clip
        multiplier the optimizer kernel applies) or the plain `scale` turn them into the post-clip norms the
        reference's grad_logger sees."""
        ps, sq = opt.segment_stats('g')
        self._g_params = ps
        if sq is None:
            self._g = None
            return
        mult = []
        for r in opt._ranges:
            if r.get('g') is None:
                continue
            c = coef_by_group.get(r['group']) if coef_by_group else None
            c = c.double().reshape(1) if c is not None else torch.full((1,), float(scale), dtype=torch.float64, device=sq.device)
            mult.append(c.expand(len(r['params'])))
        self._g = sq.sqrt() * torch.cat(mult)

    def capture_moments(self, opt):
        pm, sm = opt.segment_stats('m')
        _, sv = opt.segment_stats('v')
        self._m_params = pm
        if sm is None:
            self._m = self._v = None
            return
        numel = torch.tensor([p.numel() for p in pm], dtype=torch.float64, device=sm.device)
        self._m, self._v = sm / numel, sv / numel

    def stage(self, stream=None, extra=None):
        """Start the D2H copy of everything captured (+ optional `extra` device scalars) on `stream`."""
        parts = [t for t in (self._g, self._m, self._v) if t is not None]
        extra = [e.double().reshape(1) for e in (extra or [])]
        if not parts and not extra:
            self._host = None
            return
        payload = torch.cat(parts + extra)
        self._layout = (0 if self._g is None else self._g.numel(), 0 if self._m is None else self._m.numel(), len(extra))
        if self._host is None or self._host.numel() != payload.numel():
            self._host = torch.empty(payload.numel(), dtype=torch.float64).pin_memory()
            self._event = torch.cuda.Event()
        cur = torch.cuda.current_stream(payload.device)
        stream = stream or cur
        stream.wait_stream(cur)
        with torch.cuda.stream(stream):
            self._host.copy_(payload, non_blocking=True)
            self._event.record(stream)
        payload.record_stream(stream)

    def collect(self, named_enc, named_pred):
        """(enc grad stats, pred grad stats, {'exp_avg','exp_avg_sq'} meters, extra scalars) -- waits for the staged
        copy only."""
        if self._host is None:
            return _grad_stats_from([], []), _grad_stats_from([], []), {'exp_avg': AverageMeter(), 'exp_avg_sq': AverageMeter()}, []
        self._event.synchronize()
        ng, nm, ne = self._layout
        host = self._host.tolist()
        g, m, v, extra = host[:ng], host[ng:ng + nm], host[ng + nm:ng + 2 * nm], host[ng + 2 * nm:]
        norm_of = {id(p): val for p, val in zip(self._g_params, g)} if ng else {}
        out = []
        for named in (named_enc, named_pred):
            names, vals = [], []
            for n, p in named:
                if id(p) in norm_of and _is_weight(n, p):
                    names.append(n)
                    vals.append(norm_of[id(p)])
            out.append(_grad_stats_from(names, vals))
        first, second = AverageMeter(), AverageMeter()
        for a, b in zip(m, v):
            first.update(a)
            second.update(b)
        return out[0], out[1], {'exp_avg': first, 'exp_avg_sq': second}, extra
