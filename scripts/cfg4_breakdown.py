"""cfg4 (float64 (4096, 2M), precision 5): where does a cold FlacArray.from_array spend its time?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import flacarray_b200 as fa
from flacarray_b200 import libflacarray as lf, _lib
from flacarray_b200.utils import quanta_from_precision

dev = torch.device("cuda", 0)
n, L = 4096, 2000000
g = torch.Generator(device=dev); g.manual_seed(1)
x = torch.empty((n, L), dtype=torch.float64, device=dev)
for i in range(0, n, 64):
    x[i:i + 64] = torch.randn((64, L), generator=g, device=dev, dtype=torch.float64)
torch.cuda.synchronize()


def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e


def stats(tag):
    s = torch.cuda.memory_stats()
    print(f"  [{tag}] reserved {s['reserved_bytes.all.current'] / 1e9:.1f} GB, allocated {s['allocated_bytes.all.current'] / 1e9:.1f} GB, "
          f"retries {s['num_alloc_retries']}, cudaMalloc calls {s['segment.all.allocated']}", flush=True)


for rep in range(3):
    t0 = time.perf_counter(); a = ev()
    q = quanta_from_precision(x, 5, (n,))
    b = ev(); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"rep {rep}: quanta_from_precision {a.elapsed_time(b):.1f} ms (wall {1e3 * (t1 - t0):.1f})"); stats("std")
    qd = lf.to_device(np.asarray(q, dtype=np.float64).reshape(-1), dev, torch.float64)
    ctx = _lib.context(dev); ctx.profile(True)
    t0 = time.perf_counter(); a = ev()
    comp, starts, nbytes, off, gain = lf.encode_device(x.reshape(-1), n, L, 5, qd)
    b = ev(); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"rep {rep}: encode_device {a.elapsed_time(b):.1f} ms (wall {1e3 * (t1 - t0):.1f}), kernels {ctx.profile_ms(0)}, "
          f"{x.numel() * 8 / a.elapsed_time(b) / 1e6:.1f} GB/s"); stats("enc")
    ctx.profile(False)
    del comp
    t0 = time.perf_counter(); a = ev()
    far = fa.FlacArray.from_array(x, precision=5)
    b = ev(); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"rep {rep}: from_array {a.elapsed_time(b):.1f} ms (wall {1e3 * (t1 - t0):.1f})"); stats("from_array")
    del far
