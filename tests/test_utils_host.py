"""Host-side step utilities (section 8f rows 1-2): repeat_interleave_batch against
reference-generated goldens, the logging helpers against the live reference module, the three autograd collectives on a
world_size-2 gloo group, and the bf16 LossScaler's call-order semantics."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import torch

from conftest import REFERENCE, ROOT
from helpers import golden


def test_repeat_interleave_batch_matches_reference_goldens():
    from avjepa_b200.src.utils.tensors import repeat_interleave_batch
    gz = golden('tensors_misc.npz')
    x = torch.from_numpy(gz['rib_in'])
    for B, rep in ((2, 1), (2, 2), (3, 2), (1, 3), (6, 2)):
        out = repeat_interleave_batch(x, B, rep)
        assert np.array_equal(out.numpy(), gz[f'rib_B{B}_r{rep}']), (B, rep)
    # mask tensors are int64: dtype and values survive
    m = torch.arange(8, dtype=torch.int64).reshape(4, 2)
    assert repeat_interleave_batch(m, 2, 2).dtype == torch.int64
    assert repeat_interleave_batch(m, 2, 2).tolist() == [[0, 1], [2, 3], [0, 1], [2, 3], [4, 5], [6, 7], [4, 5], [6, 7]]


def _ref_logging():
    if not os.path.isdir(REFERENCE):
        pytest.skip('reference not mounted')
    sys.path.insert(0, REFERENCE)
    try:
        import importlib
        return importlib.import_module('src.utils.logging')
    finally:
        sys.path.remove(REFERENCE)


def test_meters_and_csv_logger_behave_like_the_reference(tmp_path):
    from avjepa_b200.src.utils import logging as ours
    ref = _ref_logging()
    a, b = ours.AverageMeter(), ref.AverageMeter()
    for v, n in ((0.5, 1), (2.0, 3), (-1.25, 2), (7, 1)):
        a.update(v, n)
        b.update(v, n)
        for f in ('val', 'avg', 'max', 'min', 'sum', 'count'):
            assert getattr(a, f) == getattr(b, f), f
    cols = (('%d', 'epoch'), ('%.5f', 'loss'), ('%d', 'gpu-time(ms)'))
    la, lb = ours.CSVLogger(str(tmp_path / 'a.csv'), *cols), ref.CSVLogger(str(tmp_path / 'b.csv'), *cols)
    for row in ((1, 0.123456, 12.7), (2, 3.0, 99)):
        la.log(*row)
        lb.log(*row)
    assert (tmp_path / 'a.csv').read_text() == (tmp_path / 'b.csv').read_text()
    r, ms = ours.gpu_timer(lambda: 41 + 1)
    assert r == 42 and (ms == -1. or ms >= 0.)


def test_grad_and_adamw_loggers_match_the_reference_on_cpu():
    from avjepa_b200.src.utils import logging as ours
    ref = _ref_logging()
    torch.manual_seed(0)
    net = torch.nn.ModuleDict({'blocks': torch.nn.ModuleList([torch.nn.ModuleDict({'attn': torch.nn.ModuleDict({
        'qkv': torch.nn.Linear(8, 24), 'proj': torch.nn.Linear(8, 8)}), 'norm1': torch.nn.LayerNorm(8)}) for _ in range(3)])})
    opt = torch.optim.AdamW(net.parameters(), lr=1e-3)
    for p in net.parameters():
        p.grad = torch.randn_like(p)
    opt.step()
    ga, gb = ours.grad_logger(net.named_parameters()), ref.grad_logger(net.named_parameters())
    for f in ('avg', 'min', 'max', 'count', 'first_layer', 'last_layer'):
        assert getattr(ga, f) == pytest.approx(getattr(gb, f), rel=1e-6), f
    oa, ob = ours.adamw_logger(opt), ref.adamw_logger(opt)
    for k in ('exp_avg', 'exp_avg_sq'):
        for f in ('avg', 'min', 'max', 'count'):
            assert getattr(oa[k], f) == pytest.approx(getattr(ob[k], f), rel=1e-6), (k, f)
    # no gradients at all -> the reference's zero defaults
    for p in net.parameters():
        p.grad = None
    z = ours.grad_logger(net.named_parameters())
    assert z.first_layer == 0. and z.last_layer == 0. and z.count == 0


COLLECTIVES_WORKER = textwrap.dedent('''
    import os, sys, torch
    import torch.distributed as tdist
    sys.path.insert(0, %r)
    from avjepa_b200.src.utils.distributed import AllGather, AllReduce, AllReduceSum, init_distributed
    world, rank = init_distributed()
    assert world == 2 and tdist.get_backend() == 'gloo'
    x = torch.tensor([1.0 + rank, 10.0 * (rank + 1)], requires_grad=True)
    m = AllReduce.apply(x * 1.0)                      # mean over ranks: [1.5, 15]
    assert torch.allclose(m, torch.tensor([1.5, 15.0])), m
    m.sum().backward()
    assert torch.equal(x.grad, torch.ones(2))         # gradient passes through
    s = AllReduceSum.apply(torch.tensor([1.0 + rank]))
    assert float(s) == 3.0
    y = torch.tensor([[float(rank), 2.0]], requires_grad=True)
    g = AllGather.apply(y * 1.0)
    assert g.tolist() == [[0.0, 2.0], [1.0, 2.0]], g
    (g * torch.tensor([[1.0, 2.0], [3.0, 4.0]])).sum().backward()
    # backward: all-reduce of the incoming gradient, then this rank's slice: 2 x the local rows of the weights
    assert y.grad.tolist() == [[2.0 * (1.0 + 2 * rank), 2.0 * (2.0 + 2 * rank)]], y.grad
    print('rank', rank, 'ok')
''') % ROOT


def test_autograd_collectives_world2_gloo(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(COLLECTIVES_WORKER)
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT='29537', WORLD_SIZE='2', CUDA_VISIBLE_DEVICES='')
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=180)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert 'ok' in o


def test_collectives_are_identity_without_a_process_group():
    from avjepa_b200.src.utils.distributed import AllGather, AllReduce, AllReduceSum
    x = torch.tensor([1.0, 2.0])
    for fn in (AllGather, AllReduce, AllReduceSum):
        assert torch.equal(fn.apply(x), x)


def test_bf16_loss_scaler_is_the_identity_and_keeps_the_reference_call_order():
    """scale -> backward -> unscale_ -> clip -> step -> update (app/avjepa/train.py:514-523): gradients seen between
    unscale_ and step are the TRUE gradients (the old 2^16 scale with a no-op unscale_ made every
    clipped step renormalise to ~1.5e-4)."""
    from avjepa_b200.app.avjepa.utils import LossScaler
    w = torch.nn.Parameter(torch.tensor([1.0, -2.0]))
    opt = torch.optim.SGD([w], lr=0.5)
    sc = LossScaler()
    assert sc.is_enabled() and sc.get_scale() == 1.0
    loss = (w * torch.tensor([3.0, 4.0])).sum()
    assert sc.scale(loss) is loss
    sc.scale(loss).backward()
    sc.unscale_(opt)
    assert w.grad.tolist() == [3.0, 4.0]
    total = torch.nn.utils.clip_grad_norm_([w], 1.0)
    assert float(total) == pytest.approx(5.0)
    sc.step(opt)
    sc.update()
    assert torch.allclose(w.detach(), torch.tensor([1.0 - 0.5 * 0.6, -2.0 - 0.5 * 0.8]), atol=1e-6)
    # a reference checkpoint's GradScaler state (scale 65536, growth tracker, ...) is not adopted
    sc.load_state_dict({'scale': 65536.0, 'growth_factor': 2.0, 'backoff_factor': 0.5, 'growth_interval': 2000, '_growth_tracker': 7})
    assert sc.get_scale() == 1.0 and set(sc.state_dict()) >= {'scale', 'growth_factor', 'backoff_factor', 'growth_interval'}
    # an explicit scale is honoured faithfully: unscale_ divides once, double unscale_ is an error like GradScaler's
    sc.update(new_scale=8.0)
    w.grad = None
    sc.scale((w * 2.0).sum()).backward()
    assert w.grad.tolist() == [16.0, 16.0]
    sc.unscale_(opt)
    assert w.grad.tolist() == [2.0, 2.0]
    with pytest.raises(RuntimeError):
        sc.unscale_(opt)
    sc.step(opt)
    sc.update()
