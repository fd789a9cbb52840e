"""Data-parallel plumbing: one process per GPU, NCCL over NVLink/NVSwitch.

The reference wraps its models in single-process ``nn.DataParallel`` and never synchronises
gradients between ranks; with one visible GPU per process that is a
pass-through.  This build shards the batch across ranks and averages gradients so that N
ranks x B/N clips reproduce the 1-GPU full-batch step:
one all-reduce per flat gradient buffer of the fused optimizer (4 buffers for the AV model,
the largest ~1.2 GB fp32 at ViT-L), issued in bucket-sized chunks so the first chunks are on
the wire while later ones are still being enqueued.
"""
import os

import torch
import torch.distributed as tdist


def init_distributed(port=37123, rank_and_world_size=(None, None)):
    """Returns (world_size, rank).  Reads torchrun's env (RANK / WORLD_SIZE / LOCAL_RANK /
    MASTER_ADDR / MASTER_PORT); falls back to world_size 1 like the reference."""
    if tdist.is_available() and tdist.is_initialized():
        return tdist.get_world_size(), tdist.get_rank()
    rank, world = rank_and_world_size
    if rank is None or world is None:
        if 'RANK' in os.environ and 'WORLD_SIZE' in os.environ:
            rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
        else:
            return 1, 0
    if world <= 1:
        return 1, 0
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    os.environ.setdefault('MASTER_PORT', str(port))
    backend = 'nccl' if torch.cuda.is_available() else 'gloo'
    if backend == 'nccl':
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', rank % max(1, torch.cuda.device_count()))))
    tdist.init_process_group(backend=backend, world_size=world, rank=rank)
    return world, rank


class GradSync(object):
    """Averages the optimizer's flat gradient buffers across ranks."""

    def __init__(self, world_size, bucket_bytes=64 << 20, group=None):
        self.world_size = world_size
        self.bucket_elems = bucket_bytes // 4
        self.group = group

    def all_reduce_flat(self, flats, average=True):
        """SUM every flat buffer across ranks in bucket-sized chunks.  With average=False the
        caller applies the returned 1/world factor itself (the fused optimizer folds it into its
        gradient multiplier).  Returns the factor still owed (1.0 when already applied)."""
        if self.world_size <= 1:
            return 1.0
        works = []
        for g in flats:
            n = g.numel()
            for lo in range(0, n, self.bucket_elems):
                chunk = g[lo:min(n, lo + self.bucket_elems)]
                works.append(tdist.all_reduce(chunk, op=tdist.ReduceOp.SUM, group=self.group, async_op=True))
        for w in works:
            w.wait()
        inv = 1.0 / self.world_size
        if not average:
            return inv
        for g in flats:
            g.mul_(inv)
        return 1.0

    def all_reduce(self, optimizer, average=True):
        if hasattr(optimizer, 'flat_grads'):
            flats = optimizer.flat_grads()
        else:
            flats = [p.grad for grp in optimizer.param_groups for p in grp['params'] if p.grad is not None]
        return self.all_reduce_flat(flats, average=average)
