"""Summarise `ncu --page source --csv --print-source cuda,sass` output: hottest source lines."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; hdr = None; data = []
for r in rows:
    if not r: continue
    if r[0] in ('File Name', 'File Path'): cur = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None or r[0] == '': continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    def g(name):
        try: return float(r[hdr.index(name)])
        except Exception: return 0.0
    data.append((g('Instructions Executed'), g('# Samples'), cur, ln, r[1].strip()[:100], g('stall_barrier'), g('stall_long_sb'), g('stall_short_sb'), g('stall_mio'), g('stall_lg'), g('stall_wait'), g('stall_math')))
ti = sum(d[0] for d in data); ts = sum(d[1] for d in data)
print(f"total warp-inst {ti:.3e}  samples {ts:.0f}")
print("by instructions:")
for d in sorted(data, reverse=True)[:top]:
    print(f"{d[0]/ti*100:5.1f}%i {d[1]/ts*100:5.1f}%s {d[2]}:{d[3]} | {d[4]}")
print("by samples (stalls: barrier long_sb short_sb mio lg wait math):")
for d in sorted(data, key=lambda x: -x[1])[:top]:
    print(f"{d[1]/ts*100:5.1f}%s {d[0]/ti*100:5.1f}%i {d[2]}:{d[3]} [{d[5]:.0f} {d[6]:.0f} {d[7]:.0f} {d[8]:.0f} {d[9]:.0f} {d[10]:.0f} {d[11]:.0f}] | {d[4][:70]}")
