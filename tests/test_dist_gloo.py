"""N>1 host logic on CPU: world_size-2 gloo run of the gradient averaging (GradSync) --
sharded-batch gradients averaged over ranks equal the full-batch gradient."""
import os
import subprocess
import sys
import textwrap

from conftest import ROOT

WORKER = textwrap.dedent('''
    import os, sys, torch
    import torch.distributed as tdist
    sys.path.insert(0, %r)
    from avjepa_b200.dist import GradSync, init_distributed
    world, rank = init_distributed()
    assert world == 2 and tdist.get_backend() == 'gloo'
    torch.manual_seed(0)
    w = torch.randn(1000, 8)
    x = torch.randn(6, 1000)                       # global batch of 6 "clips"
    def grad(rows):                                # d/dw of mean over rows of sum((x w)^2)
        wp = w.clone().requires_grad_(True)
        ((x[rows] @ wp) ** 2).sum(dim=1).mean().backward()
        return wp.grad
    full = grad(slice(0, 6))
    mine = grad(slice(3 * rank, 3 * rank + 3)).contiguous()
    flats = [mine.view(-1)[:5000].clone(), mine.view(-1)[5000:].clone()]
    GradSync(world, bucket_bytes=4096).all_reduce_flat(flats)
    got = torch.cat(flats).view_as(full)
    assert torch.allclose(got, full, rtol=1e-5, atol=1e-6), (got - full).abs().max()
    print('rank', rank, 'ok')
''') % ROOT


def test_grad_sync_world2_gloo(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT='29531', WORLD_SIZE='2', CUDA_VISIBLE_DEVICES='')
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=180)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert 'ok' in o


def test_gradsync_pending_intervals_cover_exactly_once():
    """Bookkeeping of the overlapped all-reduce (dist.GradSync): whatever order the backward reports finished
    ranges in, every element is reduced exactly once."""
    import random
    from avjepa_b200.dist import GradSync
    rng = random.Random(0)
    for _ in range(200):
        n = rng.randint(1, 300)
        count = [0] * n
        done = []
        for _ in range(rng.randint(0, 8)):
            lo = rng.randint(0, n)
            hi = rng.randint(lo, n)
            for a, b in GradSync.pending_intervals(done, lo, hi):
                assert lo <= a < b <= hi
                for i in range(a, b):
                    count[i] += 1
                done.append((a, b))
        for a, b in GradSync.pending_intervals(done, 0, n):
            for i in range(a, b):
                count[i] += 1
            done.append((a, b))
        assert count == [1] * n
        iv = sorted(done)
        assert iv[0][0] == 0 and iv[-1][1] == n and all(x[1] == y[0] for x, y in zip(iv, iv[1:]))


PIPE_WORKER = textwrap.dedent('''
    import os, sys, torch
    import torch.distributed as tdist
    sys.path.insert(0, %r)
    import avjepa_b200.dist as D
    from avjepa_b200.dist import GradSync, init_distributed
    world, rank = init_distributed()
    assert world == 2 and tdist.get_backend() == 'gloo'

    class Opt(object):                                   # the slice of FusedAdamWEMA the pipelined finish talks to
        def __init__(self):
            torch.manual_seed(7)
            base = [torch.randn(4096), torch.randn(1000)]
            self.full = [b * 3.0 for b in base]                     # what the SUM over both ranks must be: rank r holds (r + 1) * base
            self._ranges = [dict(g=(b * (rank + 1)).clone(), group=i) for i, b in enumerate(base)]
            self.updated = [torch.zeros_like(b) for b in base]
            self.log = []
        def ensure_built(self):
            pass
        def flat_grads(self):
            return [r['g'] for r in self._ranges]
        def range_of_grad(self, g):
            return next(r for r in self._ranges if r['g'] is g)
        def step_interval(self, r, lo, hi):
            i = next(k for k, x in enumerate(self._ranges) if x is r)
            self.updated[i][lo:hi] += r['g'][lo:hi]              # consumes the REDUCED values of exactly this interval
            self.log.append((i, lo, hi))

    opt = Opt()
    sync = GradSync(world, bucket_bytes=4096)            # 1024-element chunks: several collectives per interval
    # arm the bookkeeping the way begin_step does on a GPU (the per-layer CUDA events are not needed for this test)
    sync._opt, sync._works, sync._records = opt, [], []
    sync._done = {id(r['g']): [] for r in opt._ranges}
    D._ACTIVE = sync
    g0, g1 = opt.flat_grads()
    sync._reduce(g0, 2048, 4096)                         # "top layers" first, like the backward reports them
    sync._reduce(g1, 0, 1000)
    sync._reduce(g0, 1024, 3000)                         # overlaps what is already reduced: only [1024, 2048) is new
    assert sync.can_pipeline(opt)
    order = []
    for g, lo, hi in sync.finish_pipelined(opt):
        order.append((0 if g is g0 else 1, lo, hi))
        opt.step_interval(opt.range_of_grad(g), lo, hi)
    assert order == [(0, 2048, 4096), (1, 0, 1000), (0, 1024, 2048), (0, 0, 1024)], order     # issue order, complement last
    for u, f in zip(opt.updated, opt.full):              # every element reduced and consumed exactly once
        assert torch.allclose(u, f, rtol=1e-6, atol=1e-6), (u - f).abs().max()
    assert D._ACTIVE is None and not sync.can_pipeline(opt)
    print('rank', rank, 'ok')
''') % ROOT


def test_pipelined_finish_world2_gloo(tmp_path):
    """GradSync.finish_pipelined under two real ranks: intervals come back in the order their collectives were issued, the
    complement is reduced last, and an optimizer that consumes each interval as it is handed out sees every element's SUM
    exactly once."""
    script = tmp_path / 'pipe_worker.py'
    script.write_text(PIPE_WORKER)
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT='29532', WORLD_SIZE='2', CUDA_VISIBLE_DEVICES='')
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=180)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert 'ok' in o
