"""Shared host logic of the multiblock3d mask generators.

Both collators (``multiblock3d.MaskCollator`` and ``avmultiblock3d.AVMaskCollator``) draw
(i) one block size per call from a ``torch.Generator`` seeded with a process-shared step
counter and (ii) block positions from the GLOBAL torch CPU RNG.  The bit-exact contract with
the reference (``src/masks/multiblock3d.py:97-203``, ``src/masks/avmultiblock3d.py:105-234``)
is therefore a contract on the ORDER of RNG draws, which lives here in one place.

This is CPU work done in DataLoader workers; it is not on the GPU critical path.
"""
import math
from multiprocessing import Value

import torch


class StepCounter(object):
    """Process-shared iteration counter (picklable into spawned DataLoader workers)."""

    def __init__(self):
        self._v = Value('i', -1)

    def next(self):
        with self._v.get_lock():
            self._v.value += 1
            return self._v.value


def draw_block_size(seed, duration, height, width, temporal_scale, spatial_scale, aspect_ratio_scale):
    """Three seeded uniform draws -> (t, h, w) in patches."""
    gen = torch.Generator()
    gen.manual_seed(seed)
    u_t = torch.rand(1, generator=gen).item()
    u_s = torch.rand(1, generator=gen).item()
    u_ar = torch.rand(1, generator=gen).item()

    lo, hi = temporal_scale
    t = max(1, int(duration * (lo + u_t * (hi - lo))))
    lo, hi = spatial_scale
    n_spatial = int(height * width * (lo + u_s * (hi - lo)))
    lo, hi = aspect_ratio_scale
    ar = lo + u_ar * (hi - lo)
    h = min(int(round(math.sqrt(n_spatial * ar))), height)
    w = min(int(round(math.sqrt(n_spatial / ar))), width)
    return (t, h, w)


def draw_offset(extent, size):
    """One unseeded draw from the global RNG: a valid block origin in [0, extent-size]."""
    return int(torch.randint(0, extent - size + 1, (1,)))


def split_keep_drop(keep):
    """keep: flat bool tensor.  Returns (kept_idx, dropped_idx) as ascending int64 vectors."""
    return torch.nonzero(keep).flatten(), torch.nonzero(~keep).flatten()


def strict_len(v, strict=True):
    """``len`` of an index vector with the reference's quirk: there a one-element index set
    is a 0-d tensor (``argwhere(..).squeeze()``), so ``len()`` raises TypeError.  ``strict``
    reproduces that; ``strict=False`` is the guarded variant."""
    if strict and v.numel() == 1:
        raise TypeError('len() of a 0-d tensor')
    return v.numel()


def stack_truncated(rows, cap=None):
    """Truncate every row to the batch-min length (optionally capped) and stack -> [B, K] int64."""
    k = min(r.numel() for r in rows)
    if cap is not None:
        k = min(k, cap)
    return torch.stack([r[:k] for r in rows], dim=0)
