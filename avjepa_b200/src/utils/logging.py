"""Timing / CSV / meters used around the step (host).

Mirror of the reference's ``src/utils/logging.py``: ``gpu_timer :14-31``, ``get_logger :38-41``,
``CSVLogger :44-62``, ``AverageMeter :65-88``, ``grad_logger :91-105``, ``adamw_logger :108-118``.
``grad_logger`` / ``adamw_logger`` keep their return types but compute every per-tensor norm
on the device and transfer them with ONE copy instead of one blocking ``float()`` per tensor.
"""
import logging
import sys

import torch


def gpu_timer(closure, log_timings=True):
    """Times closure() with CUDA events on the current stream; returns (result, ms)."""
    log_timings = log_timings and torch.cuda.is_available()
    elapsed_time = -1.
    if log_timings:
        start = torch.cuda.Event(enable_timing=True)
        end = torch.cuda.Event(enable_timing=True)
        start.record()
    result = closure()
    if log_timings:
        end.record()
        torch.cuda.synchronize()
        elapsed_time = start.elapsed_time(end)
    return result, elapsed_time


LOG_FORMAT = "[%(levelname)-8s][%(asctime)s][%(funcName)-25s] %(message)s"
DATE_FORMAT = "%Y-%m-%d %H:%M:%S"


def get_logger(name=None, force=False):
    logging.basicConfig(stream=sys.stdout, level=logging.INFO, format=LOG_FORMAT, datefmt=DATE_FORMAT, force=force)
    return logging.getLogger(name=name)


class CSVLogger(object):

    def __init__(self, fname, *argv):
        self.fname = fname
        self.types = [v[0] for v in argv]
        with open(self.fname, '+a') as f:
            print(','.join(v[1] for v in argv), file=f)

    def log(self, *argv):
        with open(self.fname, '+a') as f:
            print(','.join(t % v for t, v in zip(self.types, argv)), file=f)


class AverageMeter(object):
    """computes and stores the average and current value"""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = 0
        self.avg = 0
        self.max = float('-inf')
        self.min = float('inf')
        self.sum = 0
        self.count = 0

    def update(self, val, n=1):
        self.val = val
        try:
            self.max = max(val, self.max)
            self.min = min(val, self.min)
        except Exception:
            pass
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


def grad_logger(named_params):
    stats = AverageMeter()
    stats.first_layer = None
    stats.last_layer = None
    names, norms = [], []
    for n, p in named_params:
        if (p.grad is not None) and not (n.endswith('.bias') or len(p.shape) == 1):
            names.append(n)
            norms.append(torch.linalg.vector_norm(p.grad.data))
    if norms:
        host = torch.stack(norms).tolist()        # one D2H transfer for all tensors
        for n, g in zip(names, host):
            stats.update(g)
            if 'qkv' in n:
                stats.last_layer = g
                if stats.first_layer is None:
                    stats.first_layer = g
    if stats.first_layer is None or stats.last_layer is None:
        stats.first_layer = stats.last_layer = 0.
    return stats


def adamw_logger(optimizer):
    """magnitude of first and second moment buffers in adamw (one D2H transfer)."""
    state = optimizer.state_dict().get('state')
    exp_avg_stats = AverageMeter()
    exp_avg_sq_stats = AverageMeter()
    a = [s.get('exp_avg').abs().mean() for s in state.values() if s.get('exp_avg') is not None]
    b = [s.get('exp_avg_sq').abs().mean() for s in state.values() if s.get('exp_avg_sq') is not None]
    if a:
        for x, y in zip(torch.stack(a).tolist(), torch.stack(b).tolist()):
            exp_avg_stats.update(x)
            exp_avg_sq_stats.update(y)
    return {'exp_avg': exp_avg_stats, 'exp_avg_sq': exp_avg_sq_stats}
