"""Thread instructions per sample and stall-sample share per source FUNCTION of one kernel, from
`ncu -i rep --page source --csv --print-source cuda,sass --kernel-name regex:NAME` (needs -lineinfo and --import-source).
Every source line is attributed to the function whose definition precedes it in its file.
usage: ncu_inst_by_function.py source.csv samples_per_launch [launches_in_capture]"""
import collections, csv, os, re, sys
rows = list(csv.reader(open(sys.argv[1])))
samples = float(sys.argv[2]); launches = int(sys.argv[3]) if len(sys.argv) > 3 else 1
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
defs = {}
def func_table(fname):
    if fname in defs: return defs[fname]
    tab = []
    for base in ("flacarray_b200/csrc", ""):
        p = os.path.join(root, base, fname)
        if os.path.exists(p):
            for n, line in enumerate(open(p, errors="ignore"), 1):
                m = re.match(r"^(?:template\s*<[^>]*>\s*)?(?:FA_D|FA_DNOINL|FA_HD|__global__|static|inline)\b.*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", line)
                if m and not line.strip().endswith(";"): tab.append((n, m.group(1)))
            break
    defs[fname] = tab
    return tab
def func_of(fname, ln):
    name = "?"
    for n, f in func_table(fname):
        if n <= ln: name = f
        else: break
    return name
cur = None; hdr = None; agg = collections.OrderedDict(); ti = ts = 0.0
for r in rows:
    if not r: continue
    if r[0] in ("File Name", "File Path"): cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or r[0] == "": continue
    try: ln = int(r[0])
    except ValueError: continue
    def g(n):
        try: return float(r[hdr.index(n)])
        except Exception: return 0.0
    key = f"{cur}:{func_of(cur, ln)}"
    a = agg.setdefault(key, [0.0, 0.0])
    a[0] += g("Thread Instructions Executed"); a[1] += g("# Samples")
    ti += g("Thread Instructions Executed"); ts += g("# Samples")
print(f"thread instructions per sample: {ti / launches / samples:.1f}   (capture of {launches} launch(es), {samples:.4g} samples each)")
print(f"{'function':58s} {'inst/sample':>11s} {'inst %':>7s} {'stall samples %':>16s}")
for k, (i, s) in sorted(agg.items(), key=lambda x: -x[1][0]):
    if i / ti < 0.004: continue
    print(f"{k:58s} {i / launches / samples:11.2f} {100 * i / ti:7.1f} {100 * s / max(ts, 1):16.1f}")
