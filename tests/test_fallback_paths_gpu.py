"""The environment switches of the C library select alternative kernels (shared-memory P instead of TMEM, one
math warpgroup, cp.async loaders, direct-store / 8-warp / 1-CTA GEMM epilogues, no programmatic dependent launch).
They are read once per process, so every combination runs in its own interpreter against the same fp32 PyTorch
references as tests/test_kernels_gpu.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import kernel_checks as kc
F32, BF16, NT, NN, TN = 0, 1, 0, 1, 2
bad = []
for args in ((BF16, 1, 300, 2, 24), (BF16, 2, 257, 3, 64), (BF16, 1, 200, 2, 16), (BF16, 1, 130, 2, 32)):
    ok, err = kc.check_attention(*args)
    if not ok: bad.append(('attention', args, err, kc.LAST_ATTENTION_ERRORS))
for args in ((BF16, NT, 700, 768, 192, 'bias_gelu'), (BF16, NN, 650, 768, 192, 'dact'), (BF16, NT, 333, 576, 192, 'bias'),
             (BF16, NT, 515, 192, 768, 'bias_res'), (BF16, TN, 192, 1536, 777, 'accum')):
    ok, err = kc.check_gemm(*args)
    if not ok: bad.append(('gemm', args, err))
print('BAD', bad)
sys.exit(1 if bad else 0)
''' % (ROOT, os.path.join(ROOT, 'tests'))

ENVS = [
    dict(AVJ_ATTN_TMEM_P='0', AVJ_ATTN_BWD_MW='1', AVJ_GEMM_TMA_STORE='0', AVJ_GEMM_EW16='0'),
    dict(AVJ_ATTN_TMA='0', AVJ_GEMM_2CTA='0', AVJ_PDL='0'),
    dict(AVJ_ATTN_POLY='1', AVJ_ATTN_BWD_MW='1', AVJ_GEMM_EW16='0'),
    dict(AVJ_GELU_EXACT='1', AVJ_ATTN_ILV='0', AVJ_ATTN_BWD_SP='0', AVJ_ATTN_POLY='0'),
    dict(AVJ_ATTN_BWD_PP='1', AVJ_GEMM_SPLITK_OLD='1'),
]


@pytest.mark.parametrize('env', ENVS, ids=lambda e: ','.join(f'{k[4:]}={v}' for k, v in e.items()))
def test_alternative_kernel_paths(env):
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, '-c', SCRIPT], env=e, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
