"""bench.py's roofline bookkeeping on a committed per-launch dump of one ViT-L step (tests/golden/prof_dump_vitl_b24_step.csv,
written by avj_prof_dump on a B200 in round 2): the dominant launch class is found over ALL kernel families, its figures
are consistent with the raw rows, ncu traffic is attached when a capture of that class is committed, and the contract keys
of the `roofline` object are all there."""
import csv
import os

from conftest import GOLDEN, ROOT

DUMP = os.path.join(GOLDEN, 'prof_dump_vitl_b24_step.csv')


def _families(bench, cabi):
    fam = {name: [0.0, 0.0, 0] for name in cabi.PROF_FAMILIES}
    with open(DUMP) as f:
        for r in csv.DictReader(f):
            e = fam[cabi.PROF_FAMILIES[int(r['family'])]]
            e[0] += float(r['ms'])
            e[1] += float(r['work'])
            e[2] += 1
    return {k: tuple(v) for k, v in fam.items()}


def test_roofline_block_from_a_recorded_step():
    import sys
    sys.path.insert(0, ROOT)
    import bench
    from avjepa_b200 import _cabi
    fam = _families(bench, _cabi)
    ms_step = sum(v[0] for v in fam.values())
    peaks = bench.read_peaks()
    step_tflops = bench.step_flops('vit_large', [(184, 40, 959, 56), (70, 15, 1145, 74)]) * 24 / (ms_step * 1e-3) / 1e12
    roof = bench.build_roofline(DUMP, fam, ms_step, step_tflops, peaks)
    for key in ('bound', 'achieved', 'peak', 'unit', 'frac', 'traffic', 'kernel', 'all_gemm', 'step', 'top_classes', 'families'):
        assert key in roof, key
    # the dominant class really is the one with the largest total time, over every family
    classes = bench.launch_classes(DUMP)
    top = max(classes.items(), key=lambda kv: kv[1][1])
    assert top[0][1] in roof['kernel']
    assert abs(roof['share_of_step'] - top[1][1] / ms_step) < 1e-9
    assert roof['bound'] in ('tensor', 'hbm') and roof['unit'] in ('TFLOP/s', 'GB/s')
    assert 0.0 < roof['frac'] < 1.0 and abs(roof['frac'] - roof['achieved'] / roof['peak']) < 1e-12
    assert 0.0 < roof['all_gemm']['frac'] < 1.0 and 0.0 < roof['step']['frac_of_sustained'] < 1.0
    # an ncu --set full capture of the dominant class is committed: traffic is DRAM bytes per launch, not far above the
    # algorithmic bytes (no wasted re-reads)
    assert roof['traffic'] is not None and roof['traffic'] > 0
    # every family that launched shows up with a share; the shares add up to the step
    shares = sum(v['share_of_step'] for v in roof['families'].values())
    assert abs(shares - 1.0) < 0.01
    # the step launches no patch-matrix kernel (family 7 'other', d0 == 1) any more: im2col-free in both directions
    with open(DUMP) as f:
        assert not any(int(r['family']) == 7 and int(r['d0']) == 1 for r in csv.DictReader(f))


def test_step_flops_reproduce_the_surveys_figures():
    """F_step = F_target + 3 F_ctx - 2 PE + 3 F_pred at the mean mask lengths of the vitl16 mask config gives
    0.552 / 0.692 / 1.180 / 2.806 / 5.165 TFLOP per clip for ViT-T/S/B/L/H.  bench.py recomputes it from the actual lengths of
    every timed step with the same function; this pins the function to the survey's numbers."""
    import sys
    sys.path.insert(0, ROOT)
    import bench
    lens = [(366, 9, 748, 59), (111, 48, 1096, 26)]        # (Kc_v, Kc_a, Kt_v, Kt_a) of mask 0 and mask 1
    for model, want in (('vit_tiny', 0.552), ('vit_small', 0.692), ('vit_base', 1.180), ('vit_large', 2.806), ('vit_huge', 5.165)):
        got = bench.step_flops(model, lens) / 1e12
        assert abs(got / want - 1.0) < 2e-3, (model, got, want)
