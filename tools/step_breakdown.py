"""Groups the per-launch timing CSV written by `bench.py --prof-dump` (avj_prof_dump) by kernel family and
shape: launches, total ms, share of the instrumented step, achieved TFLOP/s or GB/s.
usage: python tools/step_breakdown.py gpurun_out/prof_dump.csv [top_n]"""
import csv
import sys
from collections import defaultdict

FAM = ['gemm', 'attn_fwd', 'attn_bwd', 'ln_fwd', 'ln_bwd', 'colsum', 'optimizer', 'other']
OTHER = {1: 'patchify', 2: 'gather_fwd', 3: 'gather_bwd', 4: 'copy_rows', 5: 'fill_mask_tokens', 6: 'loss', 7: 'cast', 8: 'memset'}


def label(f, d):
    if f == 0:
        bits = d[0]
        lay = ['NT', 'NN', 'TN'][bits & 3]
        epi = '+'.join(n for b, n in ((4, 'bias'), (8, 'gelu'), (16, 'res'), (32, 'accum'), (64, 'dact'), (128, 'f32out')) if bits & b)
        return f'gemm {lay} {d[1]}x{d[2]}x{d[3]} {epi}'
    if f in (1, 2):
        return f'{FAM[f]} B{d[0]} N{d[1]} H{d[2]} hd{d[3]}'
    if f == 7:
        return f'{OTHER.get(d[0], "other")} {d[1]}x{d[2]}'
    return f'{FAM[f]} {d[0]}x{d[1]}' + (f' +{d[2]}' if d[2] else '')


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    groups = defaultdict(lambda: [0, 0.0, 0.0])
    fam_tot = defaultdict(lambda: [0, 0.0, 0.0])
    total = 0.0
    for r in csv.DictReader(open(path)):
        f = int(r['family'])
        d = [int(r[f'd{i}']) for i in range(4)]
        ms, work = float(r['ms']), float(r['work'])
        for tab, key in ((groups, label(f, d)), (fam_tot, FAM[f])):
            tab[key][0] += 1
            tab[key][1] += ms
            tab[key][2] += work
        total += ms
    print(f'total timed {total:.2f} ms')
    for tab, name in ((fam_tot, 'family'), (groups, 'shape')):
        print(f'--- by {name}')
        for key, (n, ms, work) in sorted(tab.items(), key=lambda kv: -kv[1][1])[:top]:
            flops = key.startswith(('gemm', 'attn'))
            rate = work / (ms * 1e-3) / (1e12 if flops else 1e9) if ms > 0 else 0.0
            print(f'{ms:9.3f} ms {100 * ms / total:5.1f}%  n={n:4d}  {ms / n * 1e3:8.1f} us/launch  {rate:8.1f} {"TF/s" if flops else "GB/s"}  {key}')


if __name__ == '__main__':
    main()
