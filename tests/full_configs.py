"""BASELINE.json configs at full size, through the public API (test infrastructure: it uses the CPU oracle).

configs 1-4 run on one B200, device-resident; config 5 runs one process per GPU under torchrun:

    python tests/full_configs.py [1 2 3 4] [--scale F]          (F < 1 shrinks the stream count)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tests/full_configs.py 5 [--scale F]

Every config is checked with size-independent properties (bit-exact integer round trip, float error bound
0.5 * quanta, keep/slice window equal to the same window of a full decode), and a sample of the GPU-encoded
streams is decoded by the CPU oracle (oracle/flac_oracle.c) and compared with the input / with the GPU decoder.
Timed with CUDA events.  One JSON line per config.  `scripts/run_configs.py` forwards here.
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
from oracle import oracle as O

SEED = 123456789
LOCAL_RANK = int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device("cuda", LOCAL_RANK)
torch.cuda.set_device(dev)


def oracle_streams(far, rows, is_int64=False):
    """Decode the streams `rows` (flat indices) of a FlacArray with the CPU oracle -> integer array [len(rows), stream_size]."""
    comp = far.compressed
    starts = np.asarray(far.stream_starts).reshape(-1)
    nbytes = np.asarray(far.stream_nbytes).reshape(-1)
    parts, st, nb, pos = [], [], [], 0
    for r in rows:
        a, n = int(starts[r]), int(nbytes[r])
        piece = comp[a:a + n]
        parts.append(piece.cpu().numpy() if torch.is_tensor(piece) else np.asarray(piece))
        st.append(pos); nb.append(n); pos += n
    return O.decode(np.concatenate(parts), np.array(st), np.array(nb), far.stream_size, is_int64=is_int64)


def sample_rows(n, k=8):
    return sorted(set(int(v) for v in np.linspace(0, n - 1, k).round()))


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
    ok_or = bool(np.array_equal(oracle_streams(far, [0, 1, 2, 3]), x.cpu().numpy()))
    # ... and the other direction: oracle-encoded bytes through the GPU decoder
    oc, os_, on = O.encode(x.cpu().numpy(), 5)
    ok_or = ok_or and bool(np.array_equal(fa.array_decompress(oc, x.shape[1], os_, on), x.cpu().numpy()))
    ok = ok and ok_or
    return dict(cfg=1, workload="int32 random walk (4, 100000) level 5", ok=ok, oracle_streams=4, oracle_ok=ok_or,
                oracle_ratio=oc.size / (x.numel() * 4), ratio=far.nbytes / (x.numel() * 4),
                enc_ms=t_enc, dec_ms=t_dec, raw_gb=x.numel() * 4 / 1e9)


def cfg2(scale):
    n = max(1, int(1000 * scale)); L = 1000000
    x = bench.make_tod_torch(n, L, SEED, dev)
    far, t_enc = timed(lambda: fa.FlacArray.from_array(x, quanta=1e-4))
    y, t_dec = timed(lambda: far.to_array())
    err = float((y - x).abs().max())
    # oracle decode of sampled GPU-encoded streams: bit-exact with the oracle's quantiser applied to the input, and its
    # restore bit-exact with the GPU decoder's float output
    rows = sample_rows(n)
    oi = oracle_streams(far, rows)
    xi, xo, xg = O.float_to_int(x[rows].cpu().numpy(), np.full(len(rows), 1e-4, np.float32))
    ok_or = bool(np.array_equal(oi, xi)) and bool(np.array_equal(O.int_to_float(oi, xo, xg), y[rows].cpu().numpy()))
    return dict(cfg=2, workload=f"float32 TOD ({n}, {L}) quanta 1e-4", ok=err <= 0.5e-4 + 2e-6 and ok_or, oracle_streams=len(rows),
                oracle_ok=ok_or, max_err=err,
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
    rows = sample_rows(n)
    ok_or = bool(np.array_equal(oracle_streams(far, rows, is_int64=True), x[rows].cpu().numpy()))
    ok = ok and ok_or
    return dict(cfg=3, workload=f"int64 ({n}, {L}) two-channel", ok=ok, oracle_streams=len(rows), oracle_ok=ok_or,
                ratio=far.nbytes / (x.numel() * 8),
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
    rows = [0, 2, 4, 6]
    oi = oracle_streams(far, rows, is_int64=True)
    og = O.int_to_float(oi, np.asarray(far.stream_offsets).reshape(-1)[rows], np.asarray(far.stream_gains).reshape(-1)[rows])
    ok_or = bool(np.array_equal(og, full.cpu().numpy()))
    ok = ok and ok_or
    del full
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
    return dict(cfg=4, workload=f"float64 ({n}, {L}) precision 5; keep rows%2==0 + slice 100k", ok=ok, oracle_streams=len(rows),
                oracle_ok=ok_or, max_err_quanta=err,
                ratio=far.nbytes / (x.numel() * 8), enc_ms=t_enc, raw_gb=x.numel() * 8 / 1e9,
                enc_gbs=x.numel() * 8 / t_enc / 1e6, partial_ms=t_part, partial_out_gb=part.numel() * 8 / 1e9,
                full_decode_512_streams_ms=t_dec512, dec_gbs=min(n, 512) * L * 8 / t_dec512 / 1e6)


def cfg5(scale):
    """10k streams x 4M float32 samples sharded over the ranks (1250 per GPU at 8): encode, NCCL all-gather of the
    byte counts -> global stream offsets, gather of every rank's block to rank 0 over NCCL (serial writer,
    io_common.write_compressed / reference hdf5.py:247-308, io_common.py:401-595), write into an HDF5-layout group,
    read back a keep + stream_slice window, compare with an oracle decode of sampled streams."""
    import torch.distributed as dist
    from flacarray_b200.memgroup import MemGroup
    from flacarray_b200.mpi import TorchComm, _even_split
    from flacarray_b200 import io_common

    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    comm = TorchComm()
    rank, world = comm.rank, comm.size
    n_glob = max(world, int(10000 * scale)); L = 4000000
    # the writer holds the whole compressed array (~0.5 x raw) in host memory next to one rank's block in flight
    try:
        with open("/proc/meminfo") as fh:
            avail = [int(ln.split()[1]) * 1024 for ln in fh if ln.startswith("MemAvailable:")][0]
        fit = int(avail / 3 / (0.5 * L * 4))
        if fit < n_glob:
            print(f"# cfg5: host memory ({avail / 1e9:.0f} GB available) limits the run to {fit} streams", file=sys.stderr, flush=True)
            n_glob = max(world, fit)
    except (OSError, IndexError):
        pass
    a, b = _even_split(n_glob, world)[rank]
    x = bench.make_tod_torch(b - a, L, SEED + 1000 * rank, dev)
    torch.cuda.synchronize(); comm.barrier()
    fa.FlacArray.from_array(x[:4], quanta=1e-4, mpi_comm=comm)          # warm-up: contexts, scratch, NCCL channels
    torch.cuda.synchronize(); comm.barrier()
    t0 = time.perf_counter(); e0 = ev()
    far = fa.FlacArray.from_array(x, quanta=1e-4, mpi_comm=comm)
    e1 = ev(); torch.cuda.synchronize()
    t_enc_dev = e0.elapsed_time(e1)
    t_enc = torch.tensor([t_enc_dev], device=dev)
    dist.all_reduce(t_enc, op=dist.ReduceOp.MAX)
    t_enc = float(t_enc)
    # bookkeeping identities of the distributed array (array.py:586-637, mpi.py:156-187)
    gstarts = np.asarray(far.global_stream_starts).reshape(-1)
    lstarts = np.asarray(far.stream_starts).reshape(-1)
    proc = far.global_process_nbytes
    checks = {}
    checks["global_starts"] = int(gstarts[0]) == int(sum(proc[:rank])) and bool(np.array_equal(gstarts - gstarts[0], lstarts))
    checks["global_nbytes"] = int(far.global_nbytes) == int(sum(proc))
    checks["global_shape"] = tuple(far.global_shape) == (n_glob, L)
    ok = all(checks.values())
    # own streams: round trip error of a sample of streams
    rows = sample_rows(b - a, 4)
    back = far.to_array(keep=np.isin(np.arange(b - a), rows))
    err = float((back - x[rows]).abs().max())
    checks["roundtrip_err"] = err <= 0.5e-4 + 2e-6
    ok = ok and checks["roundtrip_err"]
    del back
    comm.barrier()
    # gather to the writer
    grp = MemGroup() if rank == 0 else None
    tw0 = time.perf_counter()
    far.write_hdf5(grp)
    torch.cuda.synchronize(); comm.barrier()
    t_write = time.perf_counter() - tw0
    res = None
    if rank == 0:
        # read back a window of every 97th (5th when scaled down) stream, compare with (a) the oracle's decode of those streams from the bytes
        # in the group and (b) the input of the streams rank 0 owns
        keep = (np.arange(n_glob) % (97 if n_glob >= 970 else 5)) == 0
        sl = slice(L // 2 - 25000, L // 2 + 25000)
        tr0 = time.perf_counter()
        part, idx = io_common.read_array(grp, keep=keep, stream_slice=sl, keep_indices=True, no_flatten=True)
        torch.cuda.synchronize()
        t_read = time.perf_counter() - tr0
        part = part.cpu().numpy() if torch.is_tensor(part) else np.asarray(part)
        g_starts = np.asarray(grp["stream_starts"][...]).reshape(-1)
        g_nbytes = np.asarray(grp["stream_bytes"][...]).reshape(-1)
        g_off = np.asarray(grp["stream_offsets"][...]).reshape(-1)
        g_gain = np.asarray(grp["stream_gains"][...]).reshape(-1)
        comp = grp["compressed"]
        kept = np.flatnonzero(keep)
        pick = [int(kept[i]) for i in sample_rows(len(kept), 8)]
        ok_or = True
        for s_ in pick:
            buf = np.asarray(comp[int(g_starts[s_]):int(g_starts[s_] + g_nbytes[s_])])
            oi = O.decode(buf, np.array([0]), np.array([buf.size]), L)
            of = O.int_to_float(oi, g_off[s_:s_ + 1], g_gain[s_:s_ + 1])
            row = int(np.searchsorted(kept, s_))
            ok_or = ok_or and bool(np.array_equal(of[0, sl], part[row]))
            if s_ < b - a:
                ok_or = ok_or and float(np.abs(part[row] - x[s_, sl].cpu().numpy()).max()) <= 0.5e-4 + 2e-6
        ok_idx = len(idx) == int(keep.sum())
        res = dict(oracle_ok=ok_or, oracle_streams=len(pick), window=[int(keep.sum()), sl.stop - sl.start], read_s=t_read,
                   file_nbytes=int(comp.shape[0]) if hasattr(comp, "shape") else int(far.global_nbytes), ok_idx=ok_idx)
    oks = comm.allgather(bool(ok))
    allchecks = comm.gather(checks, root=0)
    if rank != 0:
        return None
    raw = n_glob * L * 4
    return dict(cfg=5, workload=f"float32 TOD ({n_glob}, {L}) over {world} GPUs, quanta 1e-4: encode + all-gather + gather to rank 0 + "
                f"HDF5-layout write (in-memory group) + keep/slice read-back", ok=all(oks) and res["oracle_ok"] and res["ok_idx"],
                rank_checks=allchecks, n_gpus=world, streams_per_gpu=b - a, ratio=far.global_nbytes / raw, raw_gb=raw / 1e9, enc_ms=t_enc,
                enc_gbs_aggregate=raw / t_enc / 1e6, gather_write_s=t_write, gather_write_gbs=far.global_nbytes / t_write / 1e9,
                max_err=err, **res)


def main(argv):
    args = [a for a in argv if not a.startswith("--")]
    scale = 1.0
    if "--scale" in argv:
        scale = float(argv[argv.index("--scale") + 1])
        args = [a for a in args if a != str(scale) and a != argv[argv.index("--scale") + 1]]
    which = [int(a) for a in args] or [1, 2, 3, 4]
    out = []
    for c in which:
        t0 = time.perf_counter()
        r = {1: cfg1, 2: lambda: cfg2(scale), 3: lambda: cfg3(scale), 4: lambda: cfg4(scale), 5: lambda: cfg5(scale)}[c]()
        if r is not None:
            r["wall_s"] = time.perf_counter() - t0
            print(json.dumps(r), flush=True)
            out.append(r)
        torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    main(sys.argv[1:])
