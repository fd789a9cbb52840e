"""GPU bring-up: runs every kernel check, isolating the tcgen05 GEMM cases in subprocesses with
timeouts (a mis-programmed mbarrier pipeline hangs rather than fails) and, if an MN-major layout
fails, sweeps the descriptor-field conventions through the AVJ_UMMA_MN_* environment knobs.
Writes a table to stdout; run as:  python tools/gpu_bringup.py [group]"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

F32, BF16 = 0, 1
NT, NN, TN = 0, 1, 2

UMMA_GROUPS = {
    'nt': [(BF16, NT, 128, 256, 64, 'none'), (BF16, NT, 128, 64, 64, 'none'), (BF16, NT, 128, 256, 256, 'none'),
           (BF16, NT, 1000, 1024, 1024, 'bias'), (BF16, NT, 333, 576, 192, 'bias'), (BF16, NT, 700, 768, 192, 'bias_gelu'),
           (BF16, NT, 515, 192, 768, 'bias_res'), (BF16, NT, 480, 384, 1536, 'pos_map'), (BF16, NT, 9000, 3072, 1024, 'bias')],
    'nn': [(BF16, NN, 128, 256, 64, 'none'), (BF16, NN, 128, 64, 64, 'none'), (BF16, NN, 128, 128, 128, 'none'),
           (BF16, NN, 900, 1024, 3072, 'none'), (BF16, NN, 650, 768, 192, 'dact')],
    'tn': [(BF16, TN, 128, 256, 64, 'accum'), (BF16, TN, 128, 64, 64, 'accum'), (BF16, TN, 256, 128, 128, 'accum'),
           (BF16, TN, 3072, 1024, 5000, 'accum'), (BF16, TN, 192, 1536, 777, 'accum'), (BF16, TN, 384, 192, 130, 'accum')],
}


def run_group_inproc(name):
    import torch
    import kernel_checks as kc
    for case in UMMA_GROUPS[name]:
        t0 = time.time()
        try:
            ok, err = kc.check_gemm(*case)
            torch.cuda.synchronize()
            print(json.dumps(dict(case=case, ok=bool(ok), err=err, s=round(time.time() - t0, 2))), flush=True)
        except Exception as e:  # noqa: BLE001
            print(json.dumps(dict(case=case, ok=False, exc=repr(e)[:300])), flush=True)


def run_group(name, env_extra=None, timeout=180):
    env = dict(os.environ)
    env.update(env_extra or {})
    try:
        p = subprocess.run([sys.executable, os.path.abspath(__file__), name], env=env, stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, timeout=timeout)
        out = p.stdout.decode()
        rc = p.returncode
    except subprocess.TimeoutExpired as e:
        out = (e.stdout or b'').decode() + '\n<<TIMEOUT>>'
        rc = -9
    res = []
    for line in out.splitlines():
        if line.startswith('{'):
            res.append(json.loads(line))
    n_ok = sum(1 for r in res if r.get('ok'))
    print(f'== umma group {name} env={env_extra} rc={rc}: {n_ok}/{len(UMMA_GROUPS[name])} ok')
    for r in res:
        print('   ', r)
    if rc != 0 and not res:
        print(out[-2000:])
    return n_ok == len(UMMA_GROUPS[name]), res


def main():
    import torch
    print('device', torch.cuda.get_device_name(0), 'torch', torch.__version__)
    import kernel_checks as kc
    simple = [
        ('gemm f32 NT bias', lambda: kc.check_gemm(F32, NT, 100, 192, 192, 'bias')),
        ('gemm f32 NN', lambda: kc.check_gemm(F32, NN, 77, 64, 136, 'none')),
        ('gemm f32 TN accum', lambda: kc.check_gemm(F32, TN, 192, 64, 333, 'accum')),
        ('gemm f32 gelu', lambda: kc.check_gemm(F32, NT, 130, 768, 192, 'bias_gelu')),
        ('gemm f32 res', lambda: kc.check_gemm(F32, NT, 64, 192, 768, 'bias_res')),
        ('gemm f32 dact', lambda: kc.check_gemm(F32, NN, 96, 768, 192, 'dact')),
        ('gemm f32 pos_map', lambda: kc.check_gemm(F32, NT, 120, 192, 256, 'pos_map')),
        ('gemm bf16 simt-fallback N=72', lambda: kc.check_gemm(BF16, NT, 50, 72, 64, 'bias')),
        ('ln f32', lambda: kc.check_layernorm(F32, 333, 192)),
        ('ln bf16 1024', lambda: kc.check_layernorm(BF16, 1000, 1024)),
        ('ln f32 1280', lambda: kc.check_layernorm(F32, 50, 1280)),
        ('ln noaffine', lambda: kc.check_layernorm(F32, 64, 768, False)),
        ('attn f32 hd64', lambda: kc.check_attention(F32, 2, 100, 3, 64)),
        ('attn f32 hd24', lambda: kc.check_attention(F32, 1, 333, 2, 24)),
        ('attn bf16 hd64', lambda: kc.check_attention(BF16, 2, 257, 4, 64)),
        ('attn bf16 hd24', lambda: kc.check_attention(BF16, 1, 130, 2, 24)),
        ('attn bf16 hd80', lambda: kc.check_attention(BF16, 1, 96, 2, 80)),
        ('attn bf16 hd128', lambda: kc.check_attention(BF16, 2, 40, 2, 128)),
        ('gather f32', lambda: kc.check_gather(F32)),
        ('gather bf16', lambda: kc.check_gather(BF16)),
        ('patchify f32 v', lambda: kc.check_patchify(F32)),
        ('patchify bf16 v', lambda: kc.check_patchify(BF16)),
        ('patchify f32 a', lambda: kc.check_patchify(F32, audio=True)),
        ('rows', kc.check_rows),
        ('loss', kc.check_loss),
        ('adamw', kc.check_adamw),
    ]
    for name, fn in simple:
        try:
            ok, err = fn()
            torch.cuda.synchronize()
            print(f'{"PASS" if ok else "FAIL"}  {name:32s} err={err:.3e}', flush=True)
        except Exception as e:  # noqa: BLE001
            print(f'EXC   {name:32s} {repr(e)[:400]}', flush=True)
    ok_nt, _ = run_group('nt')
    for g in ('nn', 'tn'):
        ok, _ = run_group(g)
        if not ok and ok_nt:
            # sweep MN-major descriptor conventions (16-byte units): (LBO, SBO, k-advance)
            for lbo, sbo, kadv in ((64, 512, 128), (512, 64, 64), (64, 512, 64), (1, 64, 128), (512, 128, 128), (128, 512, 128)):
                ok2, _ = run_group(g, dict(AVJ_UMMA_MN_LBO=str(lbo), AVJ_UMMA_MN_SBO=str(sbo), AVJ_UMMA_MN_KADV=str(kadv)))
                if ok2:
                    print(f'!! group {g} passes with LBO={lbo} SBO={sbo} KADV={kadv}')
                    break


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] in UMMA_GROUPS:
        run_group_inproc(sys.argv[1])
    else:
        main()
