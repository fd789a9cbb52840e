"""Tiny driver for `ncu --set full`: runs each named kernel case twice (first launch = warm-up) so a
capture with `-k regex:<kernel> -c <2*cases>` sees exactly these launches.
usage: python tools/ncu_cases.py gemm_proj gemm_fc1 gemm_qkv attn_fwd attn_bwd ln_bwd ..."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import kernel_bench as kb  # noqa: E402
from avjepa_b200._cabi import GEMM_NN, GEMM_NT, GEMM_TN  # noqa: E402

R_T, R_C, R_P = 24 * 1664, 24 * 384, 24 * 1216
CASES = {
    'gemm_qkv': lambda: kb.bench_gemm(GEMM_NT, R_T, 3072, 1024, 'bias', 'target qkv'),
    'gemm_proj': lambda: kb.bench_gemm(GEMM_NT, R_T, 1024, 1024, 'res', 'target proj'),
    'gemm_fc1': lambda: kb.bench_gemm(GEMM_NT, R_T, 4096, 1024, 'gelu', 'target fc1'),
    'gemm_fc2': lambda: kb.bench_gemm(GEMM_NT, R_T, 1024, 4096, 'res', 'target fc2'),
    'gemm_dact': lambda: kb.bench_gemm(GEMM_NN, R_C, 4096, 1024, 'dact', 'ctx fc2 dgrad'),
    'gemm_wgrad': lambda: kb.bench_gemm(GEMM_TN, 4096, 1024, R_C, 'accum', 'ctx fc1 wgrad'),
    'gemm_pred_fc1': lambda: kb.bench_gemm(GEMM_NT, R_P, 1536, 384, 'gelu', 'pred fc1'),
    'gemm_pred_qkv': lambda: kb.bench_gemm(GEMM_NT, R_P, 1152, 384, 'bias', 'pred qkv'),
    'gemm_pred_dact': lambda: kb.bench_gemm(GEMM_NN, R_P, 1536, 384, 'dact', 'pred fc2 dgrad'),
    'gemm_square': lambda: kb.bench_gemm(GEMM_NT, 8192, 8192, 8192, 'none', 'square 8192'),
    'attn_target': lambda: kb.bench_attn(24, 1664, 16, 64, 'target enc'),
    'attn_pred': lambda: kb.bench_attn(24, 1216, 16, 24, 'predictor'),
    'attn_pred_step': lambda: kb.bench_attn(24, 1304, 16, 24, 'predictor, the sequence length of the profiled bench step'),
    'patch_embed': lambda: kb.bench_patch_embed(24, 1024),
    'ln': lambda: kb.bench_ln(R_T, 1024),
    'ln_ctx': lambda: kb.bench_ln(24 * 537, 1024),
    'ln_pred': lambda: kb.bench_ln(24 * 2450, 384),
    'colsum': lambda: kb.bench_colsum2(24 * 537, 3072, 4096),
    'adamw': lambda: kb.bench_adamw(300_000_000),
    'gather': lambda: kb.bench_gather(24, 1568, 800, 1024),
    'loss': lambda: kb.bench_loss(24 * 1144, 1024),
}

if __name__ == '__main__':
    kb.ITERS, kb.WARM, kb.WITH_LIBRARY = 1, 1, False
    for name in sys.argv[1:]:
        CASES[name]()
    torch.cuda.synchronize()
