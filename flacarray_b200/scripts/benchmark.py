"""I/O benchmark CLI with the options and phase names of /root/reference/src/flacarray/scripts/benchmark.py
(`--out_dir`, `--global_shape`, `--use_threads`; full pass, then even streams x 100 middle samples).

Compression and decompression run on the GPU.  Containers are real HDF5 / Zarr files when h5py / zarr
import, otherwise in-memory groups (`memgroup.MemGroup`) -- the line printed for each phase says which.
Launch under torchrun for one process per GPU; the leading axis is split with the reference's rule.

    python -m flacarray_b200.scripts.benchmark --global_shape "(64,3,1000000)"
"""
import argparse
import ast
import contextlib
import os
import time

import numpy as np

from .. import hdf5 as fh5
from .. import zarr as fzr
from ..array import FlacArray
from ..demo import create_fake_data
from ..hdf5_utils import H5File, have_hdf5
from ..memgroup import MemGroup
from ..mpi import TorchComm, distribute_and_verify
from ..utils import print_timers


class _MemStore:
    """Keeps in-memory groups alive between the write and the read phase of one benchmark."""

    def __init__(self):
        self.groups = dict()

    @contextlib.contextmanager
    def open(self, path, mode, comm, zarr_style):
        rank = 0 if comm is None else comm.rank
        if mode == "w" and rank == 0:
            self.groups[path] = MemGroup(zarr_style=zarr_style)
        yield self.groups.get(path) if rank == 0 else None
        if comm is not None:
            comm.barrier()


@contextlib.contextmanager
def _h5(path, mode, comm, store):
    if have_hdf5:
        with H5File(path, mode, comm=comm) as hf:
            yield hf.handle
    else:
        with store.open(path, mode, comm, False) as g:
            yield g


@contextlib.contextmanager
def _zr(path, mode, comm, store):
    if fzr.have_zarr:
        with fzr.ZarrGroup(path, mode=mode, comm=comm) as zf:
            yield zf
    else:
        with store.open(path, mode, comm, True) as g:
            yield g


def _timed(label, comm, fn):
    import torch

    start = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    if comm is not None:
        comm.barrier()
    if comm is None or comm.rank == 0:
        print(f"  {label} in {time.perf_counter() - start:0.3f} seconds", flush=True)
    return out


def benchmark(global_shape, dir=".", keep=None, stream_slice=None, mpi_comm=None, use_threads=False):
    """Same sequence as the reference's benchmark(): FlacArray compress / write / read for HDF5 and Zarr,
    then direct write_array / read_array (benchmark.py:98-290)."""
    rank = 0 if mpi_comm is None else mpi_comm.rank
    if rank == 0:
        os.makedirs(dir, exist_ok=True)
    if mpi_comm is not None:
        mpi_comm.barrier()
    dist = distribute_and_verify(mpi_comm, global_shape[0])
    local_shape = (dist[rank][1] - dist[rank][0],) + tuple(global_shape[1:])
    arr, mpi_dist = create_fake_data(local_shape, comm=mpi_comm)
    shpstr = "x".join(f"{x}" for x in global_shape)
    store = _MemStore()
    kinds = (("HDF5", "h5", _h5, fh5, have_hdf5), ("Zarr", "zarr", _zr, fzr, fzr.have_zarr))

    for name, ext, opener, _, real in kinds:
        where = name if real else f"{name} layout (in-memory group)"
        flcarr = _timed("FlacArray compress", mpi_comm,
                        lambda: FlacArray.from_array(arr, quanta=1.0e-15, mpi_comm=mpi_comm, use_threads=use_threads))
        out_file = os.path.join(dir, f"io_bench_{shpstr}.{ext}")

        def wr():
            with opener(out_file, "w", mpi_comm, store) as g:
                (flcarr.write_hdf5 if ext == "h5" else flcarr.write_zarr)(g)

        def rd():
            with opener(out_file, "r", mpi_comm, store) as g:
                reader = FlacArray.read_hdf5 if ext == "h5" else FlacArray.read_zarr
                return reader(g, keep=keep, mpi_comm=mpi_comm, mpi_dist=mpi_dist)

        _timed(f"FlacArray write {where}", mpi_comm, wr)
        check = _timed(f"FlacArray read {where}", mpi_comm, rd)
        del flcarr, check

    for name, ext, opener, mod, real in kinds:
        where = name if real else f"{name} layout (in-memory group)"
        out_file = os.path.join(dir, f"io_bench_direct_{shpstr}.{ext}")

        def wr():
            with opener(out_file, "w", mpi_comm, store) as g:
                mod.write_array(arr, g, level=5, quanta=1.0e-15, mpi_comm=mpi_comm, use_threads=use_threads)

        def rd():
            with opener(out_file, "r", mpi_comm, store) as g:
                return mod.read_array(g, keep=keep, stream_slice=stream_slice, mpi_comm=mpi_comm,
                                      use_threads=use_threads, mpi_dist=mpi_dist)

        _timed(f"Direct compress and write {where}", mpi_comm, wr)
        check = _timed(f"Direct read {where} and decompress", mpi_comm, rd)
        del check
    del arr
    if rank == 0:
        print_timers()


def cli(argv=None):
    parser = argparse.ArgumentParser(description="Run Benchmarks")
    parser.add_argument("--out_dir", required=False, default="flacarray_benchmark_out", help="Output directory")
    parser.add_argument("--global_shape", required=False, default="(4,3,100000)", help="Global data shape (as a string)")
    parser.add_argument("--use_threads", required=False, default=False, action="store_true",
                        help="Accepted for compatibility; the GPU path has no thread switch")
    args = parser.parse_args(argv)
    shape = tuple(int(x) for x in ast.literal_eval(args.global_shape))

    comm = None
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        if not dist.is_initialized():
            dist.init_process_group("nccl")
        comm = TorchComm()
    rank = 0 if comm is None else comm.rank

    if rank == 0:
        print("Full Data Tests:", flush=True)
    benchmark(shape, dir=os.path.join(args.out_dir, "full"), use_threads=args.use_threads, mpi_comm=comm)

    keep = np.zeros(shape[:-1], dtype=bool)
    keep[0::2] = True
    mid = shape[-1] // 2
    if rank == 0:
        print("Sliced Data Tests (100 samples from even stream indices):", flush=True)
    benchmark(shape, dir=os.path.join(args.out_dir, "sliced"), keep=keep, stream_slice=slice(mid - 50, mid + 50, 1),
              use_threads=args.use_threads, mpi_comm=comm)


if __name__ == "__main__":
    cli()
