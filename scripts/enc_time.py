"""Encode-only timing (kernel experiments): python scripts/enc_time.py [streams]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from flacarray_b200 import _lib, libflacarray as lf
dev = torch.device("cuda", 0)
n_stream, n_samp = int(sys.argv[1]) if len(sys.argv) > 1 else 1000, 1000000
level = int(sys.argv[2]) if len(sys.argv) > 2 else 5
data = bench.make_tod_torch(n_stream, n_samp, 1, dev)
quanta = torch.full((n_stream,), 1e-4, dtype=torch.float32, device=dev)
flat = data.reshape(-1)
ctx = _lib.context(dev)
for it in range(4):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    if it == 3: ctx.profile(True)
    e0.record()
    try:
        comp, starts, nbytes, off, gain = lf.encode_device(flat, n_stream, n_samp, level, quanta)
    except Exception as ex:
        print("err", ex)
    e1.record(); torch.cuda.synchronize()
    print(f"it{it}: enc gpu {e0.elapsed_time(e1):.2f} ms")
print("encoder kernels ms", ctx.profile_ms(0))
