"""Benchmark of the AV-JEPA pre-training step (BASELINE.json metric: clips/sec/GPU, ViT-L/16).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores

A "step" is one full iteration of the hot path (schedules, target fwd, 2x context fwd/bwd, 2x
predictor fwd/bwd, L1 latent loss, [grad all-reduce], AdamW, EMA) on one batch of synthetic
clips shaped like configs/pretrain/vitl16.yaml: B=24 per GPU, 16x224x224 video + 128x192 log-mel,
multiblock masks from the collator under torch.manual_seed(234), random-init weights.

Printed JSON (one line, rank 0):
  value        whole-job clips/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e          same through the public API with HOST inputs: per step a pinned-host -> device copy
               of clips, spectrogram and masks and a device -> host read of the loss
  roofline     dominant kernel (tcgen05 GEMM): FLOPs / CUDA-event time of every GEMM launch of one
               instrumented step, against the measured bf16 peak in MEASURED_PEAKS.json
  cpu_baseline the CPU oracle (port of the reference step) timed on this host on a bounded sample
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
# CPU baseline / reference arm: the oracle port of the reference step on host cores
# ------------------------------------------------------------------------------------------------
def cpu_oracle_rate(model, sample_batch, steps, warmup, budget_s=200.0):
    import torch
    from oracle import avjepa_oracle as O
    from avjepa_b200.app.avjepa.utils import init_audio_video_model
    import logging
    logging.disable(logging.CRITICAL)
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    torch.manual_seed(234)
    enc, pred = init_audio_video_model(device=torch.device('cpu'), model_name=model, pred_depth=PRED_DEPTH,
                                       pred_embed_dim=PRED_DIM, uniform_power=True, use_mask_tokens=True, num_mask_tokens=2)
    strip = lambda m: {k[len('backbone.'):]: v.detach().float() for k, v in m.state_dict().items()}  # noqa: E731
    st = O.StepState(strip(enc), strip(pred), heads=MODEL_DIMS[model][2])
    del enc, pred
    ev, ea, pv, pa = sample_masks(1, sample_batch)[0]
    clips, asgram = O.synthetic_batch(sample_batch, 0)
    times = []
    t_begin = time.time()
    done = 0
    for i in range(warmup + steps):
        t0 = time.time()
        O.train_step(st, clips, asgram, ev, ea, pv, pa)
        dt = time.time() - t0
        if i >= warmup:
            times.append(dt)
            done += 1
        # keep the whole run bounded: stop early rather than run for many minutes
        if time.time() - t_begin + dt > budget_s and done >= 1:
            break
    med = statistics.median(times)
    return dict(value=sample_batch / med, unit='clips/s', cores=cores, kind='port',
                sample=f'{model} full step (fwd+bwd+AdamW+EMA) fp32, batch {sample_batch}, {done} timed step(s) after '
                       f'{min(warmup, i)} warm-up, median {med:.2f} s/step, oracle/avjepa_oracle.py on torch CPU'), med, done


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    model = args.model
    base, med, done = cpu_oracle_rate(model, args.cpu_sample_batch, args.steps, min(args.warmup, 1))
    line = dict(metric='clips/sec/GPU, ViT-L/16 AV-JEPA step', value=base['value'], unit='clips/s', n_gpus=args.gpus,
                steps=done, warmup=min(args.warmup, 1), ms_per_step=med * 1000.0, higher_is_better=True, scaling='weak',
                vs_baseline=None, dtype='f32', data='synthetic', impl='reference',
                config=dict(workload=f'{model} AV-JEPA pretrain step, vitl16.yaml shape, bounded sample batch '
                                     f'{args.cpu_sample_batch} on host cores'),
                cpu_baseline=base,
                e2e=dict(value=base['value'], unit='clips/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def dominant_gemm(csv_path):
    """The GEMM launch class with the largest total device time in the instrumented step (avj_prof_dump CSV)."""
    import csv
    groups = {}
    with open(csv_path) as f:
        for r in csv.DictReader(f):
            if int(r['family']) != 0:
                continue
            key = tuple(int(r[f'd{i}']) for i in range(4))
            g = groups.setdefault(key, [0, 0.0, 0.0])
            g[0] += 1
            g[1] += float(r['ms'])
            g[2] += float(r['work'])
    (bits, M, N, K), (n, ms, work) = max(groups.items(), key=lambda kv: kv[1][1])
    lay = ['NT', 'NN', 'TN'][bits & 3]
    epi = '+'.join(nm for b, nm in ((4, 'bias'), (8, 'gelu'), (16, 'res'), (32, 'accum'), (64, 'dact'), (128, 'f32out')) if bits & b)
    return dict(label=f'gemm {lay} {M}x{N}x{K} {epi}'.strip(), n=n, ms=ms, us=ms / n * 1e3, flops=work / n,
                tflops=work / (ms * 1e-3) / 1e12)


def ncu_traffic(label):
    """dram bytes per launch of `label` from the committed ncu --set full capture (profiles/r1_ncu_traffic.json)."""
    path = os.path.join(ROOT, 'profiles', 'r1_ncu_traffic.json')
    try:
        with open(path) as f:
            tab = json.load(f)
    except (OSError, ValueError):
        return None, 'profiles/r1_ncu_traffic.json not found'
    e = tab.get(label)
    if e is None:
        return None, f'no ncu capture for "{label}"'
    return e['dram_read_bytes'] + e['dram_write_bytes'], (f"ncu dram read {e['dram_read_bytes']} + write {e['dram_write_bytes']} B per launch; "
                                                          f"algorithmic {e['algorithmic_bytes']} B; {e['capture']}")


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
    ap.add_argument('--fp32', action='store_true', help='fp32 check mode instead of bf16')
    ap.add_argument('--profile-only', action='store_true',
                    help='run warm-up + timed steps of the resident loop and exit (driver for ncu launch lists; prints no bench line)')
    ap.add_argument('--prof-dump', default=None, help='write the per-launch timing CSV of the instrumented step here')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup

    if args.impl == 'reference':
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as tdist
    from avjepa_b200 import _cabi, engine
    from avjepa_b200.app.avjepa.train import build_training
    from avjepa_b200.dist import init_distributed
    import logging
    logging.disable(logging.CRITICAL)

    world, rank = init_distributed()
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    _cabi.load()

    cfg = json.loads(json.dumps(YAML_LIKE))
    cfg['model']['model_name'] = args.model
    cfg['data']['batch_size'] = args.batch
    if args.fp32:
        cfg['meta']['dtype'] = 'float32'
    step, _, _ = build_training(cfg, dev, world, rank)
    B, K, W = args.batch, args.steps, args.warmup

    # ---- synthetic inputs: every rank draws its own clips; masks per step from the collator
    g = torch.Generator().manual_seed(1000 + rank)
    host_clips = torch.randn(B, 3, 16, 224, 224, generator=g).pin_memory()
    host_asgram = (-80.0 * torch.rand(B, 1, 128, 192, generator=g)).pin_memory()
    n_sets = 2 * (K + W) + 1
    mask_sets = sample_masks(n_sets, B, seed=234 + rank)
    host_masks = [tuple([m.pin_memory() for m in grp] for grp in s) for s in mask_sets]
    lens = [[(s[0][i].shape[1], s[1][i].shape[1], s[2][i].shape[1], s[3][i].shape[1]) for i in range(2)] for s in mask_sets]

    def to_device(s):
        return [[m.to(dev, non_blocking=True) for m in grp] for grp in s]

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    def timed_loop(first_set, e2e):
        """Returns (ms per step, flops per clip averaged over the timed steps, launches)."""
        clips_d, asgram_d = host_clips.to(dev), host_asgram.to(dev)
        dev_masks = [to_device(host_masks[first_set + i]) for i in range(K + W)] if not e2e else None
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches = 0
        if e2e:
            # public-API path: every step's clips, spectrograms and masks start in PINNED HOST memory and are
            # staged by the package's DevicePrefetcher (copy stream, one batch ahead); the loss is read back
            # to the host every step (sync=True -> float(loss)).
            from avjepa_b200.app.avjepa.prefetch import DevicePrefetcher
            feed = DevicePrefetcher(((host_clips, host_asgram, host_masks[first_set + i]) for i in range(K + W)), dev)
        for i in range(K + W):
            if i == W:
                barrier()
                launches = _cabi.launch_count
                ev0.record()
            if e2e:
                c, a, m = next(feed)
                out = step(c, a, *m, epoch=0, sync=True)          # loss scalars read back to the host every step
            else:
                out = step(clips_d, asgram_d, *dev_masks[i], epoch=0, sync=False)
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1) / K
        launches = _cabi.launch_count - launches
        fl = sum(step_flops(args.model, lens[first_set + i]) for i in range(W, K + W)) / K
        last_loss = float(out[0])
        return ms, fl, launches, last_loss

    sampler = ClockSampler(local)
    sampler.start()
    ms_res, flops_clip, launches, loss_res = timed_loop(0, e2e=False)
    clocks = sampler.stop()
    if args.profile_only:
        if rank == 0:
            print(json.dumps(dict(profile_only=True, ms_per_step=ms_res, gpu_launches=int(launches))), flush=True)
        return
    ms_e2e, _, _, loss_e2e = timed_loop(K + W, e2e=True)

    # ---- max over ranks
    t = torch.tensor([ms_res, ms_e2e], device=dev)
    if world > 1:
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
    ms_res, ms_e2e = float(t[0]), float(t[1])

    # ---- roofline of the dominant kernel: every GEMM launch of one instrumented step
    peaks = read_peaks()
    roof = None
    # the library brackets every GEMM / attention / LayerNorm / column-sum / optimizer launch of ONE extra
    # step with CUDA events on the launch stream (avj_prof_*); nothing is serialised, so the sum per
    # family is that family's device time inside a normal step.  The step is COLLECTIVE (it contains the
    # gradient all-reduce), so every rank runs it; only rank 0 instruments and reports.
    if rank == 0:
        _cabi.prof_enable(True)
    step(host_clips.to(dev), host_asgram.to(dev), *to_device(host_masks[-1]), epoch=0, sync=True)
    torch.cuda.synchronize()
    if rank == 0:
        fam = _cabi.prof_collect()
        dump_path = args.prof_dump or os.path.join(tempfile.gettempdir(), f'avj_prof_{os.getpid()}.csv')
        _cabi.prof_dump(dump_path)
        _cabi.prof_enable(False)
        g_ms, g_fl, g_n = fam['gemm']
        peak = peaks['bf16_sustained']
        breakdown = {}
        for name, (ms_f, work, n) in fam.items():
            if n == 0:
                continue
            unit = 'TFLOP/s' if name in ('gemm', 'attention_fwd', 'attention_bwd') else 'GB/s'
            rate = work / (ms_f * 1e-3) / (1e12 if unit == 'TFLOP/s' else 1e9) if ms_f > 0 else 0.0
            breakdown[name] = dict(ms=round(ms_f, 3), launches=n, achieved=round(rate, 1), unit=unit)
        # dominant kernel = the GEMM launch class (layout, epilogue, M, N, K) with the largest total time in the step
        top = dominant_gemm(dump_path)
        all_gemm = g_fl / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
        traffic, traffic_note = ncu_traffic(top['label'])
        roof = dict(bound='tensor', kernel=f"gemm_umma2_kernel (tcgen05 cta_group::2), {top['label']}",
                    achieved=top['tflops'], peak=peak, unit='TFLOP/s', frac=top['tflops'] / peak,
                    traffic=traffic, traffic_note=traffic_note,
                    algorithmic_flops_per_launch=top['flops'], avg_launch_us=top['us'], launches=top['n'],
                    share_of_step=top['ms'] / ms_res,
                    peak_source=f'{peaks["src"]} sustained bf16 (kernel timed inside a long step)',
                    all_gemm=dict(achieved=all_gemm, frac=all_gemm / peak, launches=g_n, ms_per_step=g_ms,
                                  share_of_step=g_ms / ms_res),
                    families=breakdown)

    if world > 1:
        tdist.barrier()
        tdist.destroy_process_group()
    if rank != 0:
        return

    cpu = None
    if not args.no_cpu_baseline and world == 1:        # the CPU baseline is an N = 1 artefact (rank 0 would stall the job)
        cpu, _, _ = cpu_oracle_rate(args.model, args.cpu_sample_batch, 1, 1, budget_s=120.0)

    clips_per_s = world * B / (ms_res * 1e-3)
    e2e_clips = world * B / (ms_e2e * 1e-3)
    h2d = host_clips.numel() * 4 + host_asgram.numel() * 4 + sum(m.numel() * 8 for grp in host_masks[0] for m in grp)
    step_tflops = flops_clip * B / (ms_res * 1e-3) / 1e12
    line = dict(
        metric='clips/sec/GPU, ViT-L/16 AV-JEPA step', value=clips_per_s, unit='clips/s', n_gpus=world, steps=K, warmup=W,
        ms_per_step=ms_res, higher_is_better=True, scaling='weak', vs_baseline=None,
        dtype='f32' if args.fp32 else 'bf16', data='synthetic',
        config=dict(workload=f'{args.model} AV-JEPA pretrain step (configs/pretrain/vitl16.yaml shape), batch {B}/GPU, '
                             f'16x224x224 video + 128x192 log-mel, 2 multiblock masks, predictor depth {PRED_DEPTH}',
                    parallelism=f'dp{world}', l2='inputs+weights (>1.5 GB/step) exceed the 126 MB L2; no explicit flush',
                    clips_per_s_per_gpu=clips_per_s / world,
                    step_tflops_per_gpu=step_tflops, frac_of_bf16_peak=step_tflops / peaks['bf16_sustained'],
                    flops_per_clip=flops_clip, loss=loss_res),
        e2e=dict(value=e2e_clips, unit='clips/s', h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=12,
                 ms_per_step=ms_e2e, loss=loss_e2e),
        gpu_launches=int(launches), clocks=clocks, roofline=roof, cpu_baseline=cpu)
    print(json.dumps(line), flush=True)


if __name__ == '__main__':
    main()
