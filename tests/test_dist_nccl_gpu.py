"""Data-parallel parity on real GPUs (row e): TrainStep + GradSync over NCCL, plain and overlapped all-reduce, one
global mask set sharded by rank, flat gradients after the all-reduce against the 1-GPU full-batch gradients
(fp32 check mode <= 1e-4, bf16 <= 1e-2).  Needs >= 2 GPUs (``gpurun --gpus 2``); skipped otherwise."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_trainstep_gradsync_nccl_matches_single_gpu_full_batch():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT='29541', WORLD_SIZE='2')
    worker = os.path.join(ROOT, 'tests', 'dist_worker.py')
    procs = [subprocess.Popen([sys.executable, worker], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=900)[0].decode() for p in procs]
    print(outs[0])
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o[-4000:]
        assert 'ok' in o
