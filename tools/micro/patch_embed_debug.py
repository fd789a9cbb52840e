"""Bring-up aid for avj_patch_embed: selector weights expose what the TMA boxes put into the operand tile."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
import torch
from avjepa_b200 import engine
from avjepa_b200._cabi import RowMap

DEV = 'cuda'
torch.manual_seed(0)
B, D = 2, 64
for name, shape, tub, ntok in (('audio', (B, 1, 1, 128, 192), 1, 96), ('video', (B, 3, 16, 224, 224), 2, 1568)):
    _, Cc, T, H, W = shape
    kd = Cc * tub * 256
    for case in ('ones', 'select'):
        if case == 'ones':
            x = torch.ones(shape, device=DEV)
            w = torch.ones((D, kd), device=DEV)
        else:
            x = torch.arange(B * Cc * T * H * W, device=DEV, dtype=torch.float32).reshape(shape) % 2048
            w = torch.zeros((D, kd), device=DEV)
            ks = [(n * 37) % kd for n in range(D)]
            for n, k in enumerate(ks):
                w[n, k] = 1.0
        out = torch.full((B * ntok, D), -7.0, device=DEV)
        engine.patch_embed(x.data_ptr(), None, w.data_ptr(), out.data_ptr(), B, Cc, T, H, W, tub, 16, ntok, D, D)
        torch.cuda.synchronize()
        ref = (torch.nn.functional.conv3d(x.double(), w.double().reshape(D, Cc, tub, 16, 16), stride=(tub, 16, 16))
               .flatten(2).transpose(1, 2).reshape(B * ntok, D).float())
        bad = ~torch.isclose(out, ref, rtol=2e-3, atol=1e-2)
        print(name, case, 'nan', int(torch.isnan(out).sum()), 'untouched', int((out == -7.0).sum()), 'mismatch', int(bad.sum()), 'of', out.numel())
        if bad.any():
            r = torch.nonzero(bad)[:6].tolist()
            for i, j in r:
                print('   row', i, 'col', j, 'got', float(out[i, j]), 'want', float(ref[i, j]))
            rows_bad = bad.any(1)
            print('   bad rows:', torch.nonzero(rows_bad).flatten()[:40].tolist())
            cols_bad = bad.any(0)
            print('   bad cols:', torch.nonzero(cols_bad).flatten()[:64].tolist())
