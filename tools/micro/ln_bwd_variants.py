"""Timing experiment: which part of layernorm_bwd costs bandwidth? (dgamma/dbeta partials, column sums, low-precision copy, residual input)"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tools'))
import kernel_bench as kb
from avjepa_b200 import _cabi, engine
from avjepa_b200._cabi import BF16
rows, D = 39936, 1024
x = torch.randn((rows, D), device='cuda'); dy = torch.randn((rows, D), device='cuda').bfloat16(); dres = torch.randn((rows, D), device='cuda')
dx = torch.empty((rows, D), device='cuda'); dxl = torch.empty((rows, D), device='cuda', dtype=torch.bfloat16)
g = torch.randn(D, device='cuda'); mean = torch.randn(rows, device='cuda'); rstd = torch.rand(rows, device='cuda') + 0.5
dg = torch.zeros(D, device='cuda'); db = torch.zeros(D, device='cuda'); cs = torch.zeros(D, device='cuda')
ws = torch.empty(int(_cabi.load().avj_layernorm_bwd_ws_floats(rows, D)), device='cuda')
for name, kw in (('full (dgamma+dbeta, dres, lp)', dict(dg=dg, db=db, cs=None, dres=dres, lp=dxl)), ('+colsum', dict(dg=dg, db=db, cs=cs, dres=dres, lp=dxl)),
                 ('no partial sums', dict(dg=None, db=None, cs=None, dres=dres, lp=dxl)), ('no partials, no lp copy', dict(dg=None, db=None, cs=None, dres=dres, lp=None)),
                 ('no partials, no dres', dict(dg=None, db=None, cs=None, dres=None, lp=dxl))):
    def fn():
        engine.layernorm_bwd(dy.data_ptr(), BF16, x.data_ptr(), g.data_ptr(), mean.data_ptr(), rstd.data_ptr(), kw['dres'].data_ptr() if kw['dres'] is not None else None,
                             dx.data_ptr(), kw['lp'].data_ptr() if kw['lp'] is not None else None, BF16, kw['dg'].data_ptr() if kw['dg'] is not None else None,
                             kw['db'].data_ptr() if kw['db'] is not None else None, ws.data_ptr(), rows, D, dcolsum=kw['cs'].data_ptr() if kw['cs'] is not None else None)
    ms = kb.timeit(fn)
    b = rows * D * (2 + 4 + (4 if kw['dres'] is not None else 0) + 4 + (2 if kw['lp'] is not None else 0))
    print(json.dumps(dict(variant=name, ms=round(ms, 4), gbs=round(b / ms / 1e6, 1))), flush=True)
