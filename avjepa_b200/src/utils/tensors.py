"""Small tensor utilities used by the model factories and the train loop (host).

Mirror of the reference's ``src/utils/tensors.py``: ``trunc_normal_ :48-50`` (inverse-CDF
sampling; consumes exactly one ``uniform_`` draw per tensor, so same-seed initialisation is
bit-identical to the reference), ``repeat_interleave_batch :65-71`` and the batch-dim
``apply_masks :53-62`` variant (the models use ``src.masks.utils.apply_masks`` instead).
"""
import math

import torch


def _normal_cdf(x):
    return (1. + math.erf(x / math.sqrt(2.))) / 2.


def trunc_normal_(tensor, mean=0., std=1., a=-2., b=2.):
    """Fill with N(mean, std^2) truncated to [a, b] via the inverse CDF of a uniform draw."""
    with torch.no_grad():
        lo = _normal_cdf((a - mean) / std)
        hi = _normal_cdf((b - mean) / std)
        tensor.uniform_(2 * lo - 1, 2 * hi - 1)
        tensor.erfinv_()
        tensor.mul_(std * math.sqrt(2.))
        tensor.add_(mean)
        tensor.clamp_(min=a, max=b)
        return tensor


def apply_masks(x, masks):
    """Batch-dim variant: gathers each mask and stacks the results along dim 0."""
    from avjepa_b200.src.masks.utils import apply_masks as _gather
    return torch.cat(_gather(x, masks, concat=False), dim=0)


def repeat_interleave_batch(x, B, repeat):
    """Repeat each consecutive chunk of B rows ``repeat`` times (identity for repeat=1)."""
    n_chunks = len(x) // B
    if repeat == 1:
        return x[:n_chunks * B]
    chunks = [x[i * B:(i + 1) * B] for i in range(n_chunks) for _ in range(repeat)]
    return torch.cat(chunks, dim=0)
