"""Totals of the warp-state samples of an `ncu --page source --csv --print-source sass` dump, per stall reason,
and the hottest SASS instructions with their dominant stall."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = None; tot = {}; data = []
for r in rows:
    if not r: continue
    if r[0] in ('Line No', 'Address', '#'): hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    rec = dict(zip(hdr, r))
    st = {}
    for k, v in rec.items():
        if k.startswith('stall_') and 'Not Issued' not in k:
            try: st[k] = float(v)
            except ValueError: pass
    if not st: continue
    for k, v in st.items(): tot[k] = tot.get(k, 0) + v
    data.append((sum(st.values()), rec.get('Source', '')[:90], max(st, key=st.get) if st else ''))
s = sum(tot.values())
print('total samples', s)
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    if v: print(f"  {k:28s} {v:8.0f} {100 * v / s:5.1f}%")
for d in sorted(data, reverse=True)[:top]:
    print(f"{100 * d[0] / s:5.1f}% {d[2]:22s} {d[1]}")
