"""Opcode mix of a kernel from `ncu -i rep --page source --csv --print-source sass`.
usage: ncu_sass_mix.py sass.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ia = hdr.index('Instructions Executed'); isrc = hdr.index('Source'); it = hdr.index('Thread Instructions Executed')
tot = thr = loc = 0
ops = {}
for r in rows[2:]:
    try: ie = float(r[ia]); te = float(r[it])
    except Exception: continue
    parts = r[isrc].split()
    if not parts: continue
    op = parts[1] if parts[0].startswith('@') and len(parts) > 1 else parts[0]
    op = op.split('.')[0]
    a = ops.setdefault(op, [0, 0]); a[0] += ie; a[1] += te
    tot += ie; thr += te
    if op in ('STL', 'LDL'): loc += ie
print(f"warp-inst {tot:.4g}  thread-inst {thr:.4g}  avg active {thr/tot:.1f}  local {100*loc/tot:.1f}%")
for k, (v, t) in sorted(ops.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{k:12s} {v:14.0f} {100*v/tot:5.1f}%   active {t/max(v,1):4.1f}")
