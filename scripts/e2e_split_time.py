"""from_array / to_array of the bench workload through host (pinned) buffers, timed separately (wall clock around a
synchronize): which half of the e2e figure is further from its copy limit?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import flacarray_b200 as fa

dev = torch.device("cuda", 0)
n_stream, n_samp = (int(sys.argv[1]) if len(sys.argv) > 1 else 1000), 1000000
data = bench.make_tod_torch(n_stream, n_samp, 1, dev)
host = torch.empty((n_stream, n_samp), dtype=torch.float32, pin_memory=True)
host.copy_(data); torch.cuda.synchronize()
del data
x = host.numpy()
for it in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    far = fa.FlacArray.from_array(x, quanta=1e-4, level=5)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    y = far.to_array()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"it{it}: from_array {1e3 * (t1 - t0):7.1f} ms ({x.nbytes / (t1 - t0) / 1e9:5.1f} GB/s)   to_array {1e3 * (t2 - t1):7.1f} ms "
          f"({x.nbytes / (t2 - t1) / 1e9:5.1f} GB/s)   compressed {far.nbytes / 1e9:.2f} GB", flush=True)
