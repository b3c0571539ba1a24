"""DRAM bytes, time and instructions of ONE full-size encode call and ONE decode call, from an ncu launch list of
`bench.py` (--csv --log-file with gpu__time_duration.sum, smsp__inst_executed.sum, dram__bytes_read/write.sum).
An encode call = the launches from a k_minmax over all `n_stream` streams up to the next k_enc_finalize.
usage: ncu_traffic_per_call.py launches.csv [n_stream] [samples]  -> JSON on stdout"""
import collections, csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
n_stream = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
samples = float(sys.argv[3]) if len(sys.argv) > 3 else 1e9
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hi]
L = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(h): continue
    rec = dict(zip(h, r))
    L.setdefault(int(rec['ID']), {'name': rec['Kernel Name'], 'grid': rec.get('Grid Size', '')})[rec['Metric Name']] = float(rec['Metric Value'].replace(',', ''))
ids = sorted(L)
def short(n): return n.split('(')[0].replace('void ', '').split('<')[0]
def summarise(sel):
    per = collections.OrderedDict()
    for i in sel:
        k = short(L[i]['name'])
        a = per.setdefault(k, {'launches': 0, 'ms': 0.0, 'dram_bytes_read': 0.0, 'dram_bytes_write': 0.0, 'warp_inst': 0.0})
        a['launches'] += 1
        a['ms'] += L[i].get('gpu__time_duration.sum', 0.0) / 1e6
        a['dram_bytes_read'] += L[i].get('dram__bytes_read.sum', 0.0)
        a['dram_bytes_write'] += L[i].get('dram__bytes_write.sum', 0.0)
        a['warp_inst'] += L[i].get('smsp__inst_executed.sum', 0.0)
    tot = {k: sum(v[k] for v in per.values()) for k in ('ms', 'dram_bytes_read', 'dram_bytes_write', 'warp_inst')}
    for v in list(per.values()) + [tot]:
        v['thread_inst_per_sample'] = round(v['warp_inst'] * 32 / samples, 2)
        v['ms'] = round(v['ms'], 4)
    return {'kernels': per, 'total': tot}
out = {}
start = [i for i in ids if short(L[i]['name']) == 'k_minmax' and f", {n_stream}," in L[i]['grid']]
if start:
    s = start[-1] if len(start) > 1 else start[0]
    sel = []
    for i in ids:
        if i < s: continue
        sel.append(i)
        if short(L[i]['name']) == 'k_enc_finalize': break
    out['encode'] = summarise(sel)
# decode call: the last k_dec_meta .. k_dec_crc group whose tile kernel covers all the frames
tiles = [i for i in ids if short(L[i]['name']) == 'k_dec_tile']
if tiles:
    t = max(tiles, key=lambda i: L[i].get('dram__bytes_write.sum', 0.0))
    sel = [i for i in ids if t - 3 <= i <= t + 3 and short(L[i]['name']).startswith('k_dec')]
    out['decode'] = summarise(sel)
print(json.dumps(out, indent=1))
