"""Process-group bring-up and the three autograd-aware collectives of the reference's
``src/utils/distributed.py``: ``init_distributed :20-47``, ``AllGather :50-78``, ``AllReduceSum :81-97``,
``AllReduce :100-113``.  The AV-JEPA train loop calls ``AllReduce.apply`` on two logging scalars every iteration
(``app/avjepa/train.py:560-561``); the gradient all-reduce of this package's data-parallel step is separate
(:mod:`avjepa_b200.dist`).

``init_distributed`` here is :func:`avjepa_b200.dist.init_distributed`: it understands ``torchrun``'s environment
(RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*) in addition to an explicit ``rank_and_world_size`` and the SLURM
variables the reference reads, and picks NCCL on a GPU box, gloo otherwise.
"""
import torch
import torch.distributed as tdist

from avjepa_b200.dist import init_distributed  # noqa: F401


def _world():
    """World size, or 1 when there is no initialised process group."""
    if tdist.is_available() and tdist.is_initialized():
        return tdist.get_world_size()
    return 1


class AllGather(torch.autograd.Function):
    """Concatenate every rank's tensor along dim 0; the backward all-reduces the incoming gradient and hands each
    rank the slice that corresponds to its own contribution."""

    @staticmethod
    def forward(ctx, x):
        n = _world()
        if n == 1:
            return x
        x = x.contiguous()
        parts = [torch.empty_like(x) for _ in range(n)]
        tdist.all_gather(parts, x)
        return torch.cat(parts, dim=0)

    @staticmethod
    def backward(ctx, grads):
        n = _world()
        if n == 1:
            return grads
        per_rank = grads.shape[0] // n
        lo = per_rank * tdist.get_rank()
        grads = grads.contiguous()
        tdist.all_reduce(grads)
        return grads[lo:lo + per_rank]


class AllReduceSum(torch.autograd.Function):
    """In-place SUM across ranks; gradient passes through unchanged."""

    @staticmethod
    def forward(ctx, x):
        if _world() > 1:
            x = x.contiguous()
            tdist.all_reduce(x)
        return x

    @staticmethod
    def backward(ctx, grads):
        return grads


class AllReduce(torch.autograd.Function):
    """MEAN across ranks (divide locally, then sum); gradient passes through unchanged."""

    @staticmethod
    def forward(ctx, x):
        n = _world()
        if n > 1:
            x = x.contiguous() / n
            tdist.all_reduce(x)
        return x

    @staticmethod
    def backward(ctx, grads):
        return grads
