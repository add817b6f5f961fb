"""Aggregate `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv` by CUDA source line.

    python tools/ncu_lines.py src.csv [topN]
"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]; continue
    if len(r) >= 60 and r[0] == 'Line No':
        hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    if r[2] != '-':   # a SASS row (belongs to the source line above it)
        continue
    key = (cur_file, r[0])
    samp = float(r[hdr.index('# Samples')] or 0)
    inst = float(r[hdr.index('Instructions Executed')] or 0)
    st = {c: float(r[i] or 0) for i, c in enumerate(hdr) if c.startswith('stall_') and 'Not Issued' not in c}
    a = agg.setdefault(key, [0.0, 0.0, collections.Counter(), r[1]])
    a[0] += samp; a[1] += inst; a[2].update(st)
tot = sum(a[0] for a in agg.values()); toti = sum(a[1] for a in agg.values())
print('total samples %d, warp instructions %.3g' % (tot, toti))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    top = ', '.join('%s %.0f%%' % (k[6:], 100 * v / max(a[0], 1)) for k, v in a[2].most_common(3))
    print('%5.1f%% inst %4.1f%% %s:%s  %s   [%s]' % (100 * a[0] / tot, 100 * a[1] / toti, key[0][4:], key[1], a[3].strip()[:90], top))
