"""Host-buffer path with PAGEABLE numpy input (what a user of the reference passes) vs pinned input.
Prints GB/s of raw samples for FlacArray.from_array / to_array, plus the CPU staging-copy rate."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import flacarray_b200 as fa

dev = torch.device("cuda", 0)
n_stream = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
n_samp = 1000000
data = bench.make_tod_torch(n_stream, n_samp, 1, dev)
pinned = torch.empty((n_stream, n_samp), dtype=torch.float32, pin_memory=True)
pinned.copy_(data)
torch.cuda.synchronize()
del data
pageable = np.array(pinned.numpy(), copy=True)
print("cpu threads", torch.get_num_threads(), "cores", os.cpu_count())
t0 = time.perf_counter(); pinned.copy_(torch.from_numpy(pageable)); t = time.perf_counter() - t0
print(f"pageable -> pinned CPU copy: {pageable.nbytes / t / 1e9:.1f} GB/s")
d = torch.empty((n_stream, n_samp), dtype=torch.float32, device=dev)
torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(torch.from_numpy(pageable)); torch.cuda.synchronize()
print(f"pageable -> device cudaMemcpy: {pageable.nbytes / (time.perf_counter() - t0) / 1e9:.1f} GB/s")
del d

for name, arr in (("pinned", pinned.numpy()), ("pageable", pageable)):
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        far = fa.FlacArray.from_array(arr, quanta=1e-4)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        comp = far.compressed if name == "pinned" else np.array(far.compressed, copy=True)
        far2 = fa.FlacArray(far, compressed=comp) if False else far
        t1b = time.perf_counter()
        back = far.to_array()
        torch.cuda.synchronize(); t2 = time.perf_counter()
        print(f"{name} it{it}: from_array {arr.nbytes / (t1 - t0) / 1e9:.1f} GB/s ({1e3 * (t1 - t0):.0f} ms) | "
              f"to_array {arr.nbytes / (t2 - t1b) / 1e9:.1f} GB/s ({1e3 * (t2 - t1b):.0f} ms)")
    assert np.allclose(back, arr, atol=0.6e-4, rtol=0)
# decode with a pageable compressed buffer
from flacarray_b200.decompress import array_decompress
comp = np.array(far.compressed, copy=True)
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = array_decompress(comp, n_samp, far.stream_starts, far.stream_nbytes, stream_offsets=far.stream_offsets,
                           stream_gains=far.stream_gains)
    torch.cuda.synchronize(); t = time.perf_counter() - t0
    print(f"decode from pageable compressed it{it}: {pageable.nbytes / t / 1e9:.1f} GB/s ({1e3 * t:.0f} ms)")
