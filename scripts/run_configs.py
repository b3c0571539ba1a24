"""BASELINE.json configs[0..3] at full size on one B200, device-resident, through the public API.

Every config is checked with size-independent properties (bit-exact integer round trip, float error
bound 0.5 * quanta, keep/slice window equal to the same window of a full decode) and timed with CUDA
events.  One JSON line per config.

    python scripts/run_configs.py [1 2 3 4] [--scale F]     (F < 1 shrinks the stream count)
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import flacarray_b200 as fa
from flacarray_b200 import _lib
from flacarray_b200 import libflacarray as lf

SEED = 123456789
dev = torch.device("cuda", 0)


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def timed(fn, reps=3):
    fn()
    best = 1e30
    for _ in range(reps):
        a = ev(); r = fn(); b = ev(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return r, best


def cfg1():
    g = torch.Generator(device=dev); g.manual_seed(SEED)
    x = (torch.cumsum(torch.randint(-1000, 1001, (4, 100000), generator=g, device=dev), 1)
         + torch.randint(-50, 51, (4, 100000), generator=g, device=dev)).to(torch.int32)
    far, t_enc = timed(lambda: fa.FlacArray.from_array(x, level=5))
    y, t_dec = timed(lambda: far.to_array())
    ok = bool(torch.equal(y, x))
    return dict(cfg=1, workload="int32 random walk (4, 100000) level 5", ok=ok, ratio=far.nbytes / (x.numel() * 4),
                enc_ms=t_enc, dec_ms=t_dec, raw_gb=x.numel() * 4 / 1e9)


def cfg2(scale):
    n = max(1, int(1000 * scale)); L = 1000000
    x = bench.make_tod_torch(n, L, SEED, dev)
    far, t_enc = timed(lambda: fa.FlacArray.from_array(x, quanta=1e-4))
    y, t_dec = timed(lambda: far.to_array())
    err = float((y - x).abs().max())
    return dict(cfg=2, workload=f"float32 TOD ({n}, {L}) quanta 1e-4", ok=err <= 0.5e-4 + 2e-6, max_err=err,
                ratio=far.nbytes / (x.numel() * 4), enc_ms=t_enc, dec_ms=t_dec, raw_gb=x.numel() * 4 / 1e9,
                enc_gbs=x.numel() * 4 / t_enc / 1e6, dec_gbs=x.numel() * 4 / t_dec / 1e6)


def cfg3(scale):
    n = max(1, int(2000 * scale)); L = 500000
    g = torch.Generator(device=dev); g.manual_seed(SEED)
    x = torch.empty((n, L), dtype=torch.int64, device=dev)
    step = 250
    for i in range(0, n, step):
        j = min(n, i + step)
        x[i:j] = torch.cumsum(torch.randint(-(1 << 20), (1 << 20) + 1, (j - i, L), generator=g, device=dev), 1)
        x[i:j] += (1 << 40) * torch.randint(-4, 5, (j - i, L), generator=g, device=dev)
    x[0, 0] = torch.iinfo(torch.int64).min; x[0, 1] = torch.iinfo(torch.int64).max
    x[0, 2] = 1 << 32; x[0, 3] = -(1 << 32)
    far, t_enc = timed(lambda: fa.FlacArray.from_array(x, level=5), reps=2)
    y, t_dec = timed(lambda: far.to_array(), reps=2)
    ok = bool(torch.equal(y, x))
    return dict(cfg=3, workload=f"int64 ({n}, {L}) two-channel", ok=ok, ratio=far.nbytes / (x.numel() * 8),
                enc_ms=t_enc, dec_ms=t_dec, raw_gb=x.numel() * 8 / 1e9, enc_gbs=x.numel() * 8 / t_enc / 1e6,
                dec_gbs=x.numel() * 8 / t_dec / 1e6)


def cfg4(scale):
    n = max(2, int(4096 * scale)); L = 2000000
    g = torch.Generator(device=dev); g.manual_seed(SEED)
    x = torch.empty((n, L), dtype=torch.float64, device=dev)
    t = torch.arange(L, device=dev, dtype=torch.float64)
    minf = 5.0 / L
    wave = 2.0 * torch.sin(2 * np.pi * 3 * minf * t) + 6.0 * torch.sin(2 * np.pi * minf * t)
    step = 64
    for i in range(0, n, step):
        j = min(n, i + step)
        dc = 5.0 * (torch.rand((j - i, 1), generator=g, device=dev, dtype=torch.float64) - 0.5)
        sc = torch.rand((j - i, 1), generator=g, device=dev, dtype=torch.float64)
        x[i:j] = torch.randn((j - i, L), generator=g, device=dev, dtype=torch.float64)
        x[i:j] += dc + sc * wave
    del wave, t
    torch.cuda.synchronize()
    def mem(tag):
        if os.environ.get("FAB_CFG_DEBUG"):
            s = torch.cuda.memory_stats()
            print(f"# cfg4 [{tag}] reserved {s['reserved_bytes.all.current'] / 1e9:.1f} GB, allocated "
                  f"{s['allocated_bytes.all.current'] / 1e9:.1f} GB, retries {s['num_alloc_retries']}, "
                  f"segments {s['segment.all.allocated']}", flush=True)

    mem("before")
    a = ev()
    far = fa.FlacArray.from_array(x, precision=5)
    b = ev(); torch.cuda.synchronize()
    t_enc = a.elapsed_time(b)
    mem("after from_array")
    if os.environ.get("FAB_CFG_DEBUG"):
        for rep in range(2):
            a2 = ev(); far2 = fa.FlacArray.from_array(x, precision=5); b2 = ev(); torch.cuda.synchronize()
            print(f"# cfg4 from_array again: {a2.elapsed_time(b2):.1f} ms", flush=True)
            del far2
        mem("after repeats")
    keep = (np.arange(n) % 2) == 0
    sl = slice(L // 2 - 50000, L // 2 + 50000)
    (part, idx), t_part = timed(lambda: far.to_array(keep=keep, stream_slice=sl, keep_indices=True), reps=2)
    # property: the window equals the same window of a full decode of a few kept streams, and is within
    # 0.5 quanta of the input
    quanta = torch.std(x[:8], dim=-1, unbiased=False) / 1e5
    k8 = np.zeros(n, bool); k8[[0, 2, 4, 6]] = True
    full = far.to_array(keep=k8)
    ok = bool(torch.equal(full[:, sl], part[:4]))
    err = float(((part[:4] - x[[0, 2, 4, 6]][:, sl]).abs() / quanta[[0, 2, 4, 6], None]).max())
    ok = ok and err <= 0.5 * 1.0001 and len(idx) == int(keep.sum())
    ctx = _lib.context(dev)
    t_dec512 = 1e30
    for rep in range(2):
        ctx.profile(True)
        a = ev()
        y = far.to_array(keep=(np.arange(n) < min(n, 512)))
        b = ev(); torch.cuda.synchronize()
        print(f"# cfg4 decode of 512 streams, rep {rep}: {a.elapsed_time(b):.2f} ms, k_dec_tile {ctx.profile_ms(1)}", flush=True)
        t_dec512 = min(t_dec512, a.elapsed_time(b))
        del y
    ctx.profile(False)
    return dict(cfg=4, workload=f"float64 ({n}, {L}) precision 5; keep rows%2==0 + slice 100k", ok=ok, max_err_quanta=err,
                ratio=far.nbytes / (x.numel() * 8), enc_ms=t_enc, raw_gb=x.numel() * 8 / 1e9,
                enc_gbs=x.numel() * 8 / t_enc / 1e6, partial_ms=t_part, partial_out_gb=part.numel() * 8 / 1e9,
                full_decode_512_streams_ms=t_dec512, dec_gbs=min(n, 512) * L * 8 / t_dec512 / 1e6)


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    scale = 1.0
    if "--scale" in sys.argv:
        scale = float(sys.argv[sys.argv.index("--scale") + 1])
        args = [a for a in args if a != str(scale) and a != sys.argv[sys.argv.index("--scale") + 1]]
    which = [int(a) for a in args] or [1, 2, 3, 4]
    for c in which:
        t0 = time.perf_counter()
        r = {1: cfg1, 2: lambda: cfg2(scale), 3: lambda: cfg3(scale), 4: lambda: cfg4(scale)}[c]()
        r["wall_s"] = time.perf_counter() - t0
        print(json.dumps(r), flush=True)
        torch.cuda.empty_cache()
