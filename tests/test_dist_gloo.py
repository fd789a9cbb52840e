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
