"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by source file + line ranges.
usage: ncu_by_region.py src.csv file:lo-hi=name ...   (file = basename; unmatched lines go to 'other')"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
regions = []
for a in sys.argv[2:]:
    spec, name = a.split('=')
    f, rng = spec.split(':')
    lo, hi = rng.split('-')
    regions.append((f, int(lo), int(hi), name))
cur = None; hdr = None
acc = {}
ti = ts = 0
for r in rows:
    if not r: continue
    if r[0] in ('File Name', 'File Path'): cur = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None: continue
    try: ln = int(r[0])
    except ValueError: continue
    def g(n):
        try: return float(r[hdr.index(n)])
        except Exception: return 0.0
    i, s = g('Instructions Executed'), g('# Samples')
    name = 'other:' + str(cur)
    for f, lo, hi, nm in regions:
        if (cur or '').endswith(f) and lo <= ln <= hi: name = nm; break
    a = acc.setdefault(name, [0, 0]); a[0] += i; a[1] += s
    ti += i; ts += s
for k, (i, s) in sorted(acc.items(), key=lambda x: -x[1][0]):
    print(f"{k:28s} inst {i:12.0f} {100*i/ti:5.1f}%   samples {100*s/ts:5.1f}%")
print("total inst", ti)
