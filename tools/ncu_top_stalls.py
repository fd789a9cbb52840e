"""Top stalled SASS instructions of each kernel in an `ncu --page source --csv` dump.
usage: ncu -i X.ncu-rep --page source --csv [--kernel-name ...] > src.csv; python tools/ncu_top_stalls.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
tables, cur, hdr, name = [], None, None, None
for r in rows:
    if r and r[0] == 'Kernel Name':
        name = r[1]
        continue
    if r and r[0] == 'Address':
        hdr = r
        cur = []
        tables.append((name, hdr, cur))
        continue
    if cur is not None and len(r) >= 8:
        cur.append(r)
for name, hdr, data in tables:
    ia, isamp, ie = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    tot = sum(int(r[isamp]) for r in data) or 1
    print('====', name[:110], 'samples', tot, 'sass', len(data))
    agg = {}
    for r in data:
        for i in stall_cols:
            if i < len(r) and r[i].isdigit():
                agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i])
    print('   ', ', '.join(f'{k[6:]}={100 * v / tot:.0f}%' for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    for i, r in sorted(enumerate(data), key=lambda x: -int(x[1][isamp]))[:topn]:
        why = max(stall_cols, key=lambda c: int(r[c]) if c < len(r) and r[c].isdigit() else 0)
        print(f'{i:5d} {int(r[isamp]):6d} {100 * int(r[isamp]) / tot:5.1f}%  exec={r[ie]:>9}  {hdr[why][6:]:12s} {r[ia].strip()[:96]}')
