import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import bench
from flacarray_b200 import _lib, libflacarray as lf
dev = torch.device("cuda", 0)
n_stream, n_samp = int(sys.argv[1]) if len(sys.argv) > 1 else 1000, 1000000
data = bench.make_tod_torch(n_stream, n_samp, 1, dev)
quanta = torch.full((n_stream,), 1e-4, dtype=torch.float32, device=dev)
flat = data.reshape(-1)
ctx = _lib.context(dev)
for it in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
    e0.record()
    comp, starts, nbytes, off, gain = lf.encode_device(flat, n_stream, n_samp, 5, quanta)
    e1.record()
    t1 = time.perf_counter()
    mx = int(nbytes.max().item())
    out = lf.decode_device(comp, starts, nbytes, n_stream, n_samp, -1, -1, False, mx, 4096, off, gain)
    e2.record(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"it{it}: enc host {1e3*(t1-t0):.1f} ms gpu {e0.elapsed_time(e1):.1f} | dec host {1e3*(t2-t1):.1f} gpu {e1.elapsed_time(e2):.1f} | mem {torch.cuda.memory_allocated()/1e9:.1f} GB reserved {torch.cuda.memory_reserved()/1e9:.1f}")
ctx.profile(True)
comp, starts, nbytes, off, gain = lf.encode_device(flat, n_stream, n_samp, 5, quanta)
out = lf.decode_device(comp, starts, nbytes, n_stream, n_samp, -1, -1, False, int(nbytes.max().item()), 4096, off, gain)
print("k_encode ms", ctx.profile_ms(0), "k_dec_tile ms", ctx.profile_ms(1))
