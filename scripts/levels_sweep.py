"""Encode/decode time and ratio per compression level (float32 TOD, 200 x 1M)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from flacarray_b200 import libflacarray as lf
dev = torch.device("cuda", 0)
n, L = 200, 1000000
data = bench.make_tod_torch(n, L, 1, dev)
q = torch.full((n,), 1e-4, dtype=torch.float32, device=dev)
for level in (0, 1, 2, 3, 4, 5, 6, 7, 8):
    best_e = best_d = 1e9
    for it in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
        e0.record()
        comp, starts, nbytes, off, gain = lf.encode_device(data.reshape(-1), n, L, level, q)
        e1.record()
        bs = 1152 if level < 3 else 4096
        out = lf.decode_device(comp, starts, nbytes, n, L, -1, -1, False, int(nbytes.max().item()), bs, off, gain)
        e2.record(); torch.cuda.synchronize()
        best_e = min(best_e, e0.elapsed_time(e1)); best_d = min(best_d, e1.elapsed_time(e2))
    err = float((out.view(n, L) - data).abs().max())
    print(f"level {level}: ratio {comp.numel() / (4 * n * L):.4f} enc {best_e:7.2f} ms ({4e-6 * n * L / best_e:6.1f} GB/s) dec {best_d:6.2f} ms ({4e-6 * n * L / best_d:6.1f} GB/s) max err {err:.2e}", flush=True)
