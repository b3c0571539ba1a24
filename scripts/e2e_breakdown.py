"""Where does the host-buffer (e2e) path spend its time?  Wraps the staging helpers with synchronised
timers and prints per-phase totals for FlacArray.from_array / to_array on the bench workload."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import flacarray_b200 as fa
from flacarray_b200 import libflacarray as lf

dev = torch.device("cuda", 0)
n_stream = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
n_samp = 1000000
data = bench.make_tod_torch(n_stream, n_samp, 1, dev)
host = torch.empty((n_stream, n_samp), dtype=torch.float32, pin_memory=True)
host.copy_(data)
torch.cuda.synchronize()
host_np = host.numpy()
del data

# raw link bandwidth, pinned
d = torch.empty_like(host, device=dev)
for name, fn in (("H2D pinned", lambda: d.copy_(host, non_blocking=True)), ("D2H pinned", lambda: host.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); t = time.perf_counter() - t0
    print(f"{name}: {host_np.nbytes / t / 1e9:.1f} GB/s ({1e3 * t:.1f} ms for {host_np.nbytes / 1e9:.1f} GB)")
# both directions at once
h2 = torch.empty((n_stream, n_samp), dtype=torch.float32, pin_memory=True)
d2 = torch.empty_like(d)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.perf_counter()
with torch.cuda.stream(s1):
    d.copy_(host, non_blocking=True)
with torch.cuda.stream(s2):
    h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); t = time.perf_counter() - t0
print(f"H2D + D2H concurrently: {2 * host_np.nbytes / t / 1e9:.1f} GB/s aggregate ({1e3 * t:.1f} ms)")
del d, d2, h2

acc = {}


def timed(mod, name):
    f = getattr(mod, name)

    def w(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = f(*a, **k)
        torch.cuda.synchronize(); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
        return r
    setattr(mod, name, w)


if os.environ.get("FAB_BREAKDOWN", "1") == "1":
    for nm in ("to_device", "to_host", "encode_device", "decode_device"):
        timed(lf, nm)

for it in range(3):
    acc.clear()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    far = fa.FlacArray.from_array(host_np, quanta=1e-4)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    a1 = dict(acc); acc.clear()
    back = far.to_array()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"it{it}: from_array {1e3 * (t1 - t0):.1f} ms {({k: round(1e3 * v, 1) for k, v in a1.items()})} | "
          f"to_array {1e3 * (t2 - t1):.1f} ms {({k: round(1e3 * v, 1) for k, v in acc.items()})} | "
          f"e2e {2 * host_np.nbytes / (t2 - t0) / 1e9:.2f} GB/s")
