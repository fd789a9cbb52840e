"""Benchmark of the AV-JEPA pre-training step (BASELINE.json metric: clips/sec/GPU, ViT-L/16).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the UNMODIFIED reference (baseline/_ref) on host cores
    python bench.py --model vit_base --batch 32              # BASELINE config 2;  --model vit_huge: config 4
    python bench.py --frozen-forward --batch 64              # BASELINE config 5: no-grad encoder forward, N = 1664

A "step" is one full iteration of the hot path (schedules, target fwd, 2x context fwd/bwd, 2x predictor fwd/bwd, L1
latent loss, [grad all-reduce], AdamW, EMA) on one batch of synthetic clips shaped like configs/pretrain/vitl16.yaml:
B=24 per GPU, 16x224x224 video + 128x192 log-mel, multiblock masks from the collator under torch.manual_seed(234),
random-init weights.

Printed JSON (one line, rank 0):
  value          whole-job clips/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e            same through the public API with HOST inputs: per step a pinned-host -> device copy of clips,
                 spectrogram and masks and a device -> host read of the loss
  roofline       the launch class (kernel x shape) with the largest total time in one instrumented step -- across ALL
                 kernel families -- against the measured peaks in MEASURED_PEAKS.json; plus all_gemm / attention / step
  parity         the first thing the run does (N = 1): forward + backward of the SAME model at B = 2 against the fp32
                 CPU oracle (loss, global gradient error, and the oracle's own bf16-autocast error beside it)
  reference_gpu  the unmodified reference (stock PyTorch eager, bf16 autocast, cuBLAS + SDPA) on the same GPU, same
                 batch / masks / inputs, with and without its grad_logger / adamw_logger syncs -- the real competitor
  cpu_baseline   the reference's own CPU path (baseline/_ref; the oracle port if that is absent) on a bounded sample
  comm_exposed_ms  (N > 1) time per step the compute stream waited for gradient collectives
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL_DIMS = {'vit_tiny': (192, 12, 3), 'vit_small': (384, 12, 6), 'vit_base': (768, 12, 12),
              'vit_large': (1024, 24, 16), 'vit_huge': (1280, 32, 16)}
PRED_DIM, PRED_DEPTH = 384, 12

MASK_CFG = [
    dict(aspect_ratio=(0.75, 1.5), num_blocks=8, spatial_scale=(0.15, 0.15), temporal_scale=(1.0, 1.0),
         max_temporal_keep=1.0, max_keep=None),
    dict(aspect_ratio=(0.75, 1.5), num_blocks=2, spatial_scale=(0.7, 0.7), temporal_scale=(1.0, 1.0),
         max_temporal_keep=1.0, max_keep=None),
]

YAML_LIKE = dict(   # configs/pretrain/vitl16.yaml
    meta=dict(seed=234, use_sdpa=True, dtype='bfloat16'),
    mask=MASK_CFG,
    model=dict(model_name='vit_large', pred_depth=PRED_DEPTH, pred_embed_dim=PRED_DIM, uniform_power=True,
               use_mask_tokens=True, zero_init_mask_tokens=True),
    data=dict(batch_size=24, num_frames=16, tubelet_size=2, crop_size=224, patch_size=16, dataset_type='synthetic'),
    loss=dict(loss_exp=1.0, reg_coeff=0.0),
    optimization=dict(ipe=300, ipe_scale=1.25, clip_grad=10.0, weight_decay=0.04, final_weight_decay=0.4, epochs=300,
                      warmup=40, start_lr=0.0002, lr=0.000625, final_lr=1.0e-6, ema=(0.998, 1.0)),
)


def step_flops(model, mask_lens):
    """matmul FLOPs of one step per clip for the actual mask lengths
    [(Kc_v, Kc_a, Kt_v, Kt_a), ...]."""
    D, depth, _ = MODEL_DIMS[model]

    def blk(n, d):
        return 24 * n * d * d + 4 * n * n * d
    pe = 2 * 1568 * 1536 * D + 2 * 96 * 256 * D
    f_target = depth * blk(1664, D) + pe
    f_ctx = sum(depth * blk(kcv + kca, D) + pe for kcv, kca, _, _ in mask_lens)
    f_pred = sum(PRED_DEPTH * blk(kcv + kca + ktv + kta, PRED_DIM) + 2 * (kcv + kca) * D * PRED_DIM
                 + 2 * (ktv + kta) * PRED_DIM * D for kcv, kca, ktv, kta in mask_lens)
    return f_target + 3 * f_ctx - 2 * pe + 3 * f_pred


def sample_masks(n_sets, batch, seed=234):
    """Collator output for n_sets steps (re-drawing on the reference's 0-d crash)."""
    import torch
    from avjepa_b200.src.masks.avmultiblock3d import AVMaskCollator
    torch.manual_seed(seed)
    coll = AVMaskCollator(cfgs_mask=MASK_CFG, crop_size=224, num_frames=16, patch_size=16, tubelet_size=2)
    fake = [([torch.zeros(1)], 0, [0], torch.zeros(1)) for _ in range(batch)]
    sets = []
    while len(sets) < n_sets:
        try:
            _, ev, ea, pv, pa = coll(fake)
        except TypeError:
            continue
        sets.append((ev, ea, pv, pa))
    return sets


def read_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16_burst=d['bf16_tflops'], bf16_sustained=d['bf16_tflops_sustained'], hbm=d['hbm_gbs'], src='measured')
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, src='fallback')


class ClockSampler(object):
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.proc, self.idx = None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.idx), f'--query-gpu={self.QUERY}',
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0].decode()
        except Exception:
            self.proc.kill()
            out = ''
        sm, mx, reasons = [], [], set()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the reference's own implementation of the step on host cores
# ------------------------------------------------------------------------------------------------
def model_cfg(model, batch, fp32=False):
    cfg = json.loads(json.dumps(YAML_LIKE))
    cfg['model']['model_name'] = model
    cfg['data']['batch_size'] = batch
    if fp32:
        cfg['meta']['dtype'] = 'float32'
    return cfg


def cpu_reference_rate(model, sample_batch, steps, warmup, budget_s=None, full_batch=24):
    """The reference step on this host's cores: the live, unmodified reference modules (baseline/_ref, kind
    "reference") when installed, else the oracle port (kind "port").  fp32 -- CUDA autocast / GradScaler disable
    themselves without a GPU, exactly as in the reference.  Returns (cpu_baseline dict, median s/step, timed steps)."""
    import torch
    import logging
    logging.disable(logging.CRITICAL)
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    ev, ea, pv, pa = sample_masks(1, sample_batch)[0]
    g = torch.Generator().manual_seed(1000)
    clips = torch.randn(sample_batch, 3, 16, 224, 224, generator=g)
    asgram = -80.0 * torch.rand(sample_batch, 1, 128, 192, generator=g)
    from baseline import reference_step as R
    if R.available():
        kind = 'reference'
        tr = R.ReferenceTrainer(model_cfg(model, sample_batch), 'cpu', with_loggers=True)
        run = lambda: tr.train_step(clips, asgram, ev, ea, pv, pa)  # noqa: E731
        what = 'the unmodified reference modules from baseline/_ref (restated train_step, grad/adamw loggers on)'
    else:
        kind = 'port'
        from oracle import avjepa_oracle as O
        from avjepa_b200.app.avjepa.utils import init_audio_video_model
        torch.manual_seed(234)
        enc, pred = init_audio_video_model(device=torch.device('cpu'), model_name=model, pred_depth=PRED_DEPTH,
                                           pred_embed_dim=PRED_DIM, uniform_power=True, use_mask_tokens=True, num_mask_tokens=2)
        strip = lambda m: {k[len('backbone.'):]: v.detach().float() for k, v in m.state_dict().items()}  # noqa: E731
        st = O.StepState(strip(enc), strip(pred), heads=MODEL_DIMS[model][2])
        del enc, pred
        run = lambda: O.train_step(st, clips, asgram, ev, ea, pv, pa)  # noqa: E731
        what = 'oracle/avjepa_oracle.py (port of the reference step) on torch CPU'
    times, t_begin = [], time.time()
    for i in range(warmup + steps):
        t0 = time.time()
        run()
        dt = time.time() - t0
        if i >= warmup:
            times.append(dt)
        if budget_s is not None and len(times) >= 3 and time.time() - t_begin + dt > budget_s:
            break                      # bounded leg of the default bench run: at least 3 timed steps, then stop
    med = statistics.median(times)
    return dict(value=sample_batch / med, unit='clips/s', cores=cores, kind=kind, batch=sample_batch,
                sample=f'{model} full step (fwd+bwd+AdamW+EMA) fp32, batch {sample_batch} (bounded sample of the batch-{full_batch} '
                       f'workload), {len(times)} timed steps after {min(warmup, i)} warm-up, median {med:.2f} s/step, {what}'), med, len(times)


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    model = args.model
    base, med, done = cpu_reference_rate(model, args.cpu_sample_batch, args.steps, args.warmup, full_batch=args.batch)
    line = dict(metric='clips/sec/GPU, ViT-L/16 AV-JEPA step', value=base['value'], unit='clips/s', n_gpus=args.gpus,
                steps=done, warmup=args.warmup, ms_per_step=med * 1000.0, higher_is_better=True, scaling='weak',
                vs_baseline=None, dtype='f32', data='synthetic', impl='reference',
                config=dict(workload=f'{model} AV-JEPA pretrain step (configs/pretrain/vitl16.yaml shape), bounded sample: batch '
                                     f'{args.cpu_sample_batch} of the batch-{args.batch} workload, on host cores',
                            batch=args.cpu_sample_batch),
                cpu_baseline=base,
                e2e=dict(value=base['value'], unit='clips/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line), flush=True)


def reference_gpu_rate(model, batch, dev, host_clips, host_asgram, mask_sets, steps=5, warmup=2):
    """The real competitor: the unmodified reference, stock PyTorch eager under bf16 autocast
    (cuBLAS / cuDNN / SDPA), same GPU, same batch, masks and inputs -- with and without its per-parameter logging
    syncs (src/utils/logging.py:91-118).  Returns a dict or {'unavailable': why}."""
    import torch
    from baseline import reference_step as R
    if not R.available():
        return dict(unavailable='baseline/_ref not installed (tools/install_reference.sh)')
    out = dict(batch=batch, steps=steps, warmup=warmup, dtype='bf16 autocast (torch.cuda.amp, GradScaler)',
               how='baseline/reference_step.py: restated train_step around the live reference modules, CUDA events')
    clips, asgram = host_clips.to(dev), host_asgram.to(dev)
    for tag, loggers in (('with_loggers', True), ('no_loggers', False)):
        try:
            tr = R.ReferenceTrainer(model_cfg(model, batch), dev, with_loggers=loggers)
            dm = [[[m.to(dev) for m in grp] for grp in s] for s in mask_sets[:steps + warmup]]
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            last = first = None
            for i in range(steps + warmup):
                if i == warmup:
                    torch.cuda.synchronize()
                    e0.record()
                last = tr.train_step(clips, asgram, *dm[i])
                if first is None:
                    first = last['loss']              # loss of the freshly initialised model on mask set 0
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[tag] = dict(ms_per_step=ms, clips_per_s=batch / (ms * 1e-3), loss=last['loss'], first_step_loss=first)
        except Exception as e:      # the competitor arm must never take the bench line down with it
            out[tag] = dict(error=f'{type(e).__name__}: {e}'[:300])
        finally:
            tr = None
            import gc
            gc.collect()
            torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
FAM = ['gemm', 'attn_fwd', 'attn_bwd', 'ln_fwd', 'ln_bwd', 'colsum', 'optimizer', 'other']
KERNEL_OF = {0: 'gemm_umma2_kernel / gemm_umma_kernel (tcgen05, TMA)', 1: 'fa_fwd_umma_kernel (tcgen05, TMEM softmax)',
             2: 'fa_bwd_umma_kernel (tcgen05, two passes)', 3: 'layernorm_fwd_kernel', 4: 'layernorm_bwd_kernel',
             5: 'colsum kernels', 6: 'adamw_ema_kernel', 7: 'row kernels'}


def launch_label(f, d):
    """Same labels as tools/step_breakdown.py (and the keys of profiles/*ncu_traffic.json)."""
    if f == 0:
        bits = d[0]
        lay = ['NT', 'NN', 'TN'][bits & 3]
        epi = '+'.join(nm for b, nm in ((4, 'bias'), (8, 'gelu'), (16, 'res'), (32, 'accum'), (64, 'dact'), (128, 'f32out')) if bits & b)
        return f'gemm {lay} {d[1]}x{d[2]}x{d[3]} {epi}'.strip()
    if f in (1, 2):
        return f'{FAM[f]} B{d[0]} N{d[1]} H{d[2]} hd{d[3]}'
    return f'{FAM[f]} {d[0]}x{d[1]}'


def launch_classes(csv_path):
    """{(family, label): [launches, total ms, total work]} of the instrumented step (avj_prof_dump CSV)."""
    import csv
    groups = {}
    with open(csv_path) as f:
        for r in csv.DictReader(f):
            fam = int(r['family'])
            d = [int(r[f'd{i}']) for i in range(4)]
            g = groups.setdefault((fam, launch_label(fam, d)), [0, 0.0, 0.0])
            g[0] += 1
            g[1] += float(r['ms'])
            g[2] += float(r['work'])
    return groups


def ncu_traffic(label):
    """dram bytes per launch of `label` from the committed ncu --set full captures (profiles/*ncu_traffic.json)."""
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, 'profiles', '*ncu_traffic.json')), reverse=True):
        try:
            with open(path) as f:
                tab = json.load(f)
        except (OSError, ValueError):
            continue
        e = tab.get(label)
        if e is not None:
            return e['dram_read_bytes'] + e['dram_write_bytes'], (
                f"{os.path.basename(path)}: ncu dram read {e['dram_read_bytes']} + write {e['dram_write_bytes']} B per launch; "
                f"algorithmic {e['algorithmic_bytes']} B; {e['capture']}")
    return None, f'no ncu --set full capture committed for "{label}"'


def build_roofline(dump_path, fam, ms_step, step_tflops, peaks):
    """roofline block: the DOMINANT launch class of the step (largest total device time, any family), with the whole
    GEMM family, both attention families and the whole step beside it."""
    classes = launch_classes(dump_path)
    (f, label), (n, ms, work) = max(classes.items(), key=lambda kv: kv[1][1])
    tensor = f in (0, 1, 2)
    peak = peaks['bf16_sustained'] if tensor else peaks['hbm']
    rate = work / (ms * 1e-3) / (1e12 if tensor else 1e9)
    traffic, note = ncu_traffic(label)
    breakdown = {}
    for name, (ms_f, w, cnt) in fam.items():
        if cnt == 0:
            continue
        t = name in ('gemm', 'attention_fwd', 'attention_bwd')
        r = w / (ms_f * 1e-3) / (1e12 if t else 1e9) if ms_f > 0 else 0.0
        breakdown[name] = dict(ms=round(ms_f, 3), launches=cnt, achieved=round(r, 1), unit='TFLOP/s' if t else 'GB/s',
                               frac=round(r / (peaks['bf16_sustained'] if t else peaks['hbm']), 3),
                               share_of_step=round(ms_f / ms_step, 3))
    top5 = sorted(classes.items(), key=lambda kv: -kv[1][1])[:8]
    g_ms, g_fl, g_n = fam['gemm']
    all_gemm = g_fl / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
    return dict(bound='tensor' if tensor else 'hbm', kernel=f'{KERNEL_OF[f]}: {label}', achieved=rate, peak=peak,
                unit='TFLOP/s' if tensor else 'GB/s', frac=rate / peak, traffic=traffic, traffic_note=note,
                algorithmic_work_per_launch=work / n, avg_launch_us=ms / n * 1e3, launches=n, share_of_step=ms / ms_step,
                peak_source=f'{peaks["src"]} {"sustained bf16" if tensor else "HBM copy"} (kernel timed inside a long step)',
                all_gemm=dict(achieved=all_gemm, frac=all_gemm / peaks['bf16_sustained'], launches=g_n, ms_per_step=g_ms,
                              share_of_step=g_ms / ms_step),
                step=dict(achieved=step_tflops, unit='TFLOP/s', frac_of_sustained=step_tflops / peaks['bf16_sustained'],
                          frac_of_burst=step_tflops / peaks['bf16_burst']),
                top_classes=[dict(label=k[1], ms=round(v[1], 3), launches=v[0],
                                  achieved=round(v[2] / (v[1] * 1e-3) / (1e12 if k[0] < 3 else 1e9), 1)) for k, v in top5],
                families=breakdown)


def parity_block(model, dev):
    """Forward + backward of `model` at B = 2 (golden masks, seeded inputs) on the product path against the fp32 CPU
    oracle -- the checker, never the thing measured -- with the oracle's own bf16-autocast error beside it."""
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import step_support as S
    t0 = time.time()
    res = S.config_parity(model, MODEL_DIMS[model][2], dev)
    fam = res.pop('families')
    res['worst_family'] = max(fam.items(), key=lambda kv: kv[1]['rel'])[0]
    res['worst_family_rel'] = fam[res['worst_family']]['rel']
    res['bars'] = 'loss_rel <= 1e-2; grad_rel <= max(1e-2, 1.25 x autocast_grad_rel)'
    res['ok'] = bool(res['loss_rel'] <= 1e-2 and res['grad_rel'] <= max(1e-2, 1.25 * res.get('autocast_grad_rel', 0.0)))
    res['seconds'] = round(time.time() - t0, 1)
    return res


def run_frozen_forward(args):
    """BASELINE config 5: frozen-encoder forward (no masks, all 1664 tokens, no grad) at large batch -- the
    av_prediction / attentive-probe evaluation's encoder call (app/avprediction/train.py:455-474)."""
    import torch
    import torch.distributed as tdist
    from avjepa_b200 import _cabi
    from avjepa_b200.app.avjepa.utils import init_audio_video_model
    from avjepa_b200.dist import init_distributed
    import logging
    logging.disable(logging.CRITICAL)
    world, rank = init_distributed()
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    _cabi.load()
    torch.manual_seed(234)
    enc, _ = init_audio_video_model(device=dev, model_name=args.model, pred_depth=1, pred_embed_dim=PRED_DIM, uniform_power=True,
                                    use_mask_tokens=True, num_mask_tokens=2)
    for p in enc.parameters():
        p.requires_grad = False
    B, K, W = args.batch, args.steps, args.warmup
    g = torch.Generator().manual_seed(1000 + rank)
    host_clips = torch.randn(B, 3, 16, 224, 224, generator=g).pin_memory()
    host_asgram = (-80.0 * torch.rand(B, 1, 128, 192, generator=g)).pin_memory()
    D, depth, _ = MODEL_DIMS[args.model]
    flops_clip = depth * (24 * 1664 * D * D + 4 * 1664 * 1664 * D) + 2 * 1568 * 1536 * D + 2 * 96 * 256 * D

    def loop(e2e):
        c, a = host_clips.to(dev), host_asgram.to(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = 0
        for i in range(K + W):
            if i == W:
                if world > 1:
                    tdist.barrier()
                torch.cuda.synchronize()
                n0 = _cabi.launch_count()
                e0.record()
            with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16):
                if e2e:
                    c, a = host_clips.to(dev, non_blocking=True), host_asgram.to(dev, non_blocking=True)
                out = enc(c, a)
                if e2e:
                    chk = float(out[:, 0, 0].sum())             # D2H of a per-clip feature every step
        e1.record()
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / K, _cabi.launch_count() - n0

    sampler = ClockSampler(local)
    sampler.start()
    ms, launches = loop(False)
    clocks = sampler.stop()
    ms_e2e, _ = loop(True)
    t = torch.tensor([ms, ms_e2e], device=dev)
    if world > 1:
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        tdist.barrier()
        tdist.destroy_process_group()
    if rank != 0:
        return
    ms, ms_e2e = float(t[0]), float(t[1])
    peaks = read_peaks()
    tf = flops_clip * B / (ms * 1e-3) / 1e12
    line = dict(metric='clips/sec/GPU, ViT-L/16 frozen-encoder forward', value=world * B / (ms * 1e-3), unit='clips/s', n_gpus=world,
                steps=K, warmup=W, ms_per_step=ms, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='bf16',
                data='synthetic',
                config=dict(workload=f'{args.model} frozen-encoder forward (BASELINE config 5), no masks, 1664 tokens, batch {B}/GPU',
                            parallelism=f'dp{world} (replicas, no collective)', l2='activations (>2 GB/step) exceed the 126 MB L2',
                            step_tflops_per_gpu=tf, frac_of_bf16_peak=tf / peaks['bf16_sustained'], flops_per_clip=flops_clip),
                e2e=dict(value=world * B / (ms_e2e * 1e-3), unit='clips/s', h2d_bytes_per_step=int(host_clips.numel() * 4 + host_asgram.numel() * 4),
                         d2h_bytes_per_step=4, ms_per_step=ms_e2e),
                gpu_launches=int(launches), clocks=clocks,
                roofline=dict(bound='tensor', kernel='whole forward (GEMM + attention)', achieved=tf, peak=peaks['bf16_sustained'],
                              unit='TFLOP/s', frac=tf / peaks['bf16_sustained'], traffic=None), cpu_baseline=None)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--model', default='vit_large')
    ap.add_argument('--batch', type=int, default=24, help='clips per GPU (vitl16.yaml: 24)')
    ap.add_argument('--cpu-sample-batch', type=int, default=1)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-parity', action='store_true', help='skip the B=2 oracle parity block')
    ap.add_argument('--no-reference-gpu', action='store_true', help='skip the reference-on-GPU competitor leg')
    ap.add_argument('--device-masks', action='store_true',
                    help='e2e loop: masks sampled on the GPU (DeviceAVMaskCollator, one call ahead) instead of copied from pinned host memory')
    ap.add_argument('--per-rank-masks', action='store_true',
                    help='N>1: every rank seeds its mask collator differently (lengths differ across ranks -> straggler skew). '
                         'Default: identical collator streams on all ranks, which is what the reference does -- it seeds every '
                         'rank with the same meta.seed (app/avjepa/train.py:164-165), so the DataLoader workers of all ranks '
                         'draw the same block positions')
    ap.add_argument('--frozen-forward', action='store_true', help='BASELINE config 5: no-grad encoder forward at N=1664')
    ap.add_argument('--fp32', action='store_true', help='fp32 check mode instead of bf16')
    ap.add_argument('--profile-only', action='store_true',
                    help='run warm-up + timed steps of the resident loop and exit (driver for ncu launch lists; prints no bench line)')
    ap.add_argument('--prof-dump', default=None, help='write the per-launch timing CSV of the instrumented step here')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup

    if args.impl == 'reference':
        run_reference_arm(args)
        return
    if args.frozen_forward:
        run_frozen_forward(args)
        return

    import torch
    import torch.distributed as tdist
    from avjepa_b200 import _cabi
    from avjepa_b200.app.avjepa.train import build_training
    from avjepa_b200.dist import init_distributed
    import logging
    logging.disable(logging.CRITICAL)

    world, rank = init_distributed()
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    _cabi.load()

    # ---- parity first: the same model family at B = 2 against the CPU oracle (rank 0 of an N = 1 run only)
    parity = None
    if world == 1 and not args.no_parity and not args.fp32 and not args.profile_only:
        try:
            parity = parity_block(args.model, dev)
        except Exception as e:
            parity = dict(ok=False, error=f'{type(e).__name__}: {e}'[:300])
        torch.cuda.empty_cache()

    step, _, _ = build_training(model_cfg(args.model, args.batch, args.fp32), dev, world, rank)
    if step.grad_sync is not None:
        step.grad_sync.measure = True
    B, K, W = args.batch, args.steps, args.warmup

    # ---- synthetic inputs: every rank draws its own clips; masks per step from the collator
    g = torch.Generator().manual_seed(1000 + rank)
    host_clips = torch.randn(B, 3, 16, 224, 224, generator=g).pin_memory()
    host_asgram = (-80.0 * torch.rand(B, 1, 128, 192, generator=g)).pin_memory()
    n_sets = 2 * (K + W) + 1
    mask_sets = sample_masks(n_sets, B, seed=234 + (rank if args.per_rank_masks else 0))
    host_masks = [tuple([m.pin_memory() for m in grp] for grp in s) for s in mask_sets]
    lens = [[(s[0][i].shape[1], s[1][i].shape[1], s[2][i].shape[1], s[3][i].shape[1]) for i in range(2)] for s in mask_sets]

    def to_device(s):
        return [[m.to(dev, non_blocking=True) for m in grp] for grp in s]

    # full-size parity probe: the loss of the freshly initialised model on mask set 0 at the FULL batch; the reference-on-GPU
    # leg below starts from the same seed (bit-identical initial weights) and reports the same quantity from its first step
    loss_first = None
    if world == 1 and not args.no_reference_gpu and not args.fp32 and not args.profile_only:
        with torch.no_grad():
            loss_first = float(step.forward_loss(host_clips.to(dev), host_asgram.to(dev), *to_device(host_masks[0]))[0])

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    def timed_loop(first_set, e2e):
        """Returns (ms per step, flops per clip averaged over the timed steps, launches, last loss, exposed comm ms)."""
        clips_d, asgram_d = host_clips.to(dev), host_asgram.to(dev)
        dev_masks = [to_device(host_masks[first_set + i]) for i in range(K + W)] if not e2e else None
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches = 0
        if e2e:
            # public-API path: every step's clips, spectrograms and masks start in PINNED HOST memory and are
            # staged by the package's DevicePrefetcher (copy stream, one batch ahead); the loss is read back
            # to the host every step (sync=True -> float(loss)).
            from avjepa_b200.app.avjepa.prefetch import DevicePrefetcher
            feed = DevicePrefetcher(((host_clips, host_asgram, () if args.device_masks else host_masks[first_set + i])
                                     for i in range(K + W)), dev)
            dev_coll = None
            if args.device_masks:
                from avjepa_b200.src.masks.device_collator import DeviceAVMaskCollator
                torch.manual_seed(234 + (rank if args.per_rank_masks else 0))
                dev_coll = DeviceAVMaskCollator(MASK_CFG, crop_size=224, num_frames=16, patch_size=16, tubelet_size=2, device=dev,
                                                prefetch=True)
        if not e2e and not args.profile_only:
            # allocator steady state: the activation arenas grow with the largest sequence lengths seen so far (mask draws
            # differ per step), and growing means cudaMalloc -- a device synchronisation that a long training run only
            # sees in its first few hundred steps.  Two extra untimed steps with the largest context / predictor
            # sequences of this loop's mask sets come before the W warm-up steps.
            ctx_rows = [sum(lens[first_set + i][m][0] + lens[first_set + i][m][1] for m in range(2)) for i in range(K + W)]
            all_rows = [sum(sum(lens[first_set + i][m]) for m in range(2)) for i in range(K + W)]
            # always exactly two steps: under --per-rank-masks the ranks pick different sets, and a step is a collective
            for j in (max(range(K + W), key=lambda i: ctx_rows[i]), max(range(K + W), key=lambda i: all_rows[i])):
                step(clips_d, asgram_d, *dev_masks[j], epoch=0, sync=False)
        for i in range(K + W):
            if i == W:
                barrier()
                launches = _cabi.launch_count()
                mallocs = torch.cuda.memory_stats(dev).get('num_device_alloc', 0)
                if step.grad_sync is not None:
                    step.grad_sync.exposed_ms()
                ev0.record()
            if e2e:
                c, a, m = next(feed)
                while dev_coll is not None:
                    try:
                        m = dev_coll.sample(B)
                        break
                    except TypeError:                             # the reference collator's one-element quirk: draw again
                        continue
                out = step(c, a, *m, epoch=0, sync=True)          # loss scalars read back to the host every step
            else:
                out = step(clips_d, asgram_d, *dev_masks[i], epoch=0, sync=False)
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1) / K
        launches = (_cabi.launch_count() - launches) / K
        exposed = step.grad_sync.exposed_ms() if step.grad_sync is not None else 0.0
        fl = sum(step_flops(args.model, lens[first_set + i]) for i in range(W, K + W)) / K
        last_loss = float(out[0])
        mallocs = torch.cuda.memory_stats(dev).get('num_device_alloc', 0) - mallocs
        if rank == 0:
            print(f'[bench] {"e2e" if e2e else "resident"} loop: {ms:.2f} ms/step, {mallocs} cudaMalloc calls inside the timed region',
                  file=sys.stderr, flush=True)
        return ms, fl, launches, last_loss, exposed

    sampler = ClockSampler(local)
    sampler.start()
    ms_res, flops_clip, launches, loss_res, exposed_res = timed_loop(0, e2e=False)
    clocks = sampler.stop()
    if args.profile_only:
        if rank == 0:
            print(json.dumps(dict(profile_only=True, ms_per_step=ms_res, gpu_launches_per_step=launches)), flush=True)
        return
    ms_e2e, _, _, loss_e2e, exposed_e2e = timed_loop(K + W, e2e=True)

    # ---- max over ranks
    t = torch.tensor([ms_res, ms_e2e, exposed_res, exposed_e2e], device=dev)
    if world > 1:
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
    ms_res, ms_e2e, exposed_res, exposed_e2e = (float(x) for x in t)

    # ---- roofline: every GEMM / attention / LayerNorm / column-sum / optimizer launch of ONE extra step is bracketed
    # with CUDA events on the launch stream inside the library (avj_prof_*); nothing is serialised, so the sum per
    # class is that class's device time inside a normal step.  The step is COLLECTIVE (it contains the gradient
    # all-reduce), so every rank runs it; only rank 0 instruments and reports.
    peaks = read_peaks()
    roof = None
    step_tflops = flops_clip * B / (ms_res * 1e-3) / 1e12
    if rank == 0:
        _cabi.prof_enable(True)
    step(host_clips.to(dev), host_asgram.to(dev), *to_device(host_masks[-1]), epoch=0, sync=True)
    torch.cuda.synchronize()
    if rank == 0:
        fam = _cabi.prof_collect()
        dump_path = args.prof_dump or os.path.join(tempfile.gettempdir(), f'avj_prof_{os.getpid()}.csv')
        _cabi.prof_dump(dump_path)
        _cabi.prof_enable(False)
        roof = build_roofline(dump_path, fam, ms_res, step_tflops, peaks)

    if world > 1:
        tdist.barrier()
        tdist.destroy_process_group()
    if rank != 0:
        return

    # ---- the competitor and the CPU baseline (N = 1 only: rank 0 would stall a multi-rank job)
    ref_gpu = cpu = None
    if world == 1:
        del step
        torch.cuda.empty_cache()
        if not args.no_reference_gpu and not args.fp32:
            ref_gpu = reference_gpu_rate(args.model, B, dev, host_clips, host_asgram, mask_sets)
            ref_first = ref_gpu.get('no_loggers', {}).get('first_step_loss')
            if loss_first is not None and ref_first:
                ref_gpu['full_size_parity'] = dict(
                    batch=B, first_step_loss=loss_first, reference_first_step_loss=ref_first,
                    rel=abs(loss_first - ref_first) / abs(ref_first), bar=1e-2,
                    what='loss of the freshly initialised model (same seed: bit-identical weights) on the same clips and masks, '
                         'this build (bf16) vs the unmodified reference under bf16 autocast on this GPU')
            ok = ref_gpu.get('no_loggers', {}).get('clips_per_s')
            if ok:
                ref_gpu['speedup_vs_no_loggers'] = (B / (ms_e2e * 1e-3)) / ok
                wl = ref_gpu.get('with_loggers', {}).get('clips_per_s')
                ref_gpu['speedup_vs_with_loggers'] = (B / (ms_e2e * 1e-3)) / wl if wl else None
        if not args.no_cpu_baseline:
            cpu, _, _ = cpu_reference_rate(args.model, args.cpu_sample_batch, 3, 1, budget_s=150.0, full_batch=B)

    clips_per_s = world * B / (ms_res * 1e-3)
    e2e_clips = world * B / (ms_e2e * 1e-3)
    h2d = host_clips.numel() * 4 + host_asgram.numel() * 4
    if not args.device_masks:
        h2d += sum(m.numel() * 8 for grp in host_masks[0] for m in grp)
    line = dict(
        metric='clips/sec/GPU, ViT-L/16 AV-JEPA step', value=clips_per_s, unit='clips/s', n_gpus=world, steps=K, warmup=W,
        ms_per_step=ms_res, higher_is_better=True, scaling='weak', vs_baseline=None,
        dtype='f32' if args.fp32 else 'bf16', data='synthetic',
        config=dict(workload=f'{args.model} AV-JEPA pretrain step (configs/pretrain/vitl16.yaml shape), batch {B}/GPU, '
                             f'16x224x224 video + 128x192 log-mel, 2 multiblock masks, predictor depth {PRED_DEPTH}',
                    parallelism=f'dp{world}', l2='inputs+weights (>1.5 GB/step) exceed the 126 MB L2; no explicit flush',
                    prewarm='2 untimed steps with the largest mask draws before the W warm-up steps (activation arenas at steady-state size)',
                    masks='per-rank collator seeds (lengths differ across ranks)' if args.per_rank_masks else 'collator seeded with meta.seed on every rank, like the reference (same mask lengths on all ranks)',
                    e2e_masks='sampled on the GPU (avj_mask_collate, bit-exact replica of the CPU generator)' if args.device_masks else 'pinned host tensors from the host collator',
                    clips_per_s_per_gpu=clips_per_s / world,
                    step_tflops_per_gpu=step_tflops, frac_of_bf16_peak=step_tflops / peaks['bf16_sustained'],
                    flops_per_clip=flops_clip, loss=loss_res),
        e2e=dict(value=e2e_clips, unit='clips/s', h2d_bytes_per_step=int(h2d),
                 d2h_bytes_per_step=12 + (4 * (len(MASK_CFG) * B * 4 + 1) if args.device_masks else 0),
                 ms_per_step=ms_e2e, loss=loss_e2e),
        gpu_launches=int(round(launches * K)), gpu_launches_per_step=launches, clocks=clocks, roofline=roof, parity=parity,
        reference_gpu=ref_gpu, cpu_baseline=cpu)
    if world > 1:
        line['comm_exposed_ms'] = dict(resident=exposed_res, e2e=exposed_e2e,
                                       how='CUDA events around the compute stream\'s wait on the gradient all-reduces, mean per step, max over ranks')
    print(json.dumps(line), flush=True)


if __name__ == '__main__':
    main()
