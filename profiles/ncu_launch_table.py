"""Table of an ncu launch list (--csv --log-file): one row per launch, one column per metric."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hi]
d = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(h): continue
    rec = dict(zip(h, r))
    d.setdefault((rec['ID'], rec['Kernel Name'], rec.get('Grid Size')), {})[rec['Metric Name']] = rec['Metric Value']
short = {'gpu__time_duration.sum': 'ns', 'smsp__inst_executed.sum': 'winst', 'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue%',
         'smsp__inst_executed_op_local_ld.sum': 'lld', 'smsp__inst_executed_op_local_st.sum': 'lst',
         'dram__bytes_read.sum': 'rd', 'dram__bytes_write.sum': 'wr'}
lo, hi_ = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, 10**9)
for k, v in d.items():
    if not (lo <= int(k[0]) < hi_): continue
    print(k[0], k[1][:34].ljust(34), k[2].ljust(14), ' '.join(f"{short[m]}={x}" for m, x in v.items() if m in short))
