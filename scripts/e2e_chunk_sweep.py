"""e2e (pinned host buffers, public API) vs pipeline chunk size."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import flacarray_b200 as fa
from flacarray_b200 import libflacarray as lf
dev = torch.device("cuda", 0)
n, L = 1000, 1000000
data = bench.make_tod_torch(n, L, 1, dev)
host = torch.empty((n, L), dtype=torch.float32, pin_memory=True); host.copy_(data); torch.cuda.synchronize()
x = host.numpy(); del data
for mb in (192, 96, 64, 32):
    lf._PIPE_CHUNK_BYTES = mb << 20
    far = back = None
    for _ in range(3):
        far = fa.FlacArray.from_array(x, quanta=1e-4); back = far.to_array()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3):
        far = fa.FlacArray.from_array(x, quanta=1e-4); t1 = time.perf_counter(); back = far.to_array()
    torch.cuda.synchronize(); t = (time.perf_counter() - t0) / 3
    print(f"chunk {mb:4d} MB: e2e {2 * x.nbytes / t / 1e9:.2f} GB/s ({1e3 * t:.1f} ms per round trip)", flush=True)
