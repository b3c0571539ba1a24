"""Zarr front end of the FlacArray group layout: same functions and arguments as
/root/reference/src/flacarray/zarr.py:145-524 on top of `io_common` (identical dataset / attribute
names, zarr_load_v1.py:21-29; arrays are made with `create_array` and accessed by slicing because zarr
arrays have no read_direct / write_direct).  Format versions 0 and 1 are read, version 1 is written.
"""
from . import io_common as _io
from .utils import function_timer

try:
    import zarr as _zarr  # noqa: F401

    have_zarr = True
except Exception:  # pragma: no cover - optional dependency
    _zarr = None
    have_zarr = False

zarr_names = dict(_io.NAMES)


class ZarrGroup(object):
    """Context manager opening a zarr group on rank 0 only (reference zarr.py:30-62)."""

    def __init__(self, name, mode, comm=None):
        self.handle = None
        if not have_zarr:
            raise RuntimeError("zarr is not importable")
        if comm is None or comm.rank == 0:
            self.handle = _zarr.open_group(name, mode=mode)
        if comm is not None:
            comm.barrier()

    def close(self):
        self.handle = None

    def __enter__(self):
        return self.handle

    def __exit__(self, *args):
        self.close()


@function_timer
def write_compressed(zgrp, leading_shape, global_leading_shape, stream_size, stream_starts, global_stream_starts,
                     stream_nbytes, stream_offsets, stream_gains, compressed, n_channels, local_nbytes, global_nbytes,
                     global_process_nbytes, mpi_comm, mpi_dist):
    from . import __version__

    return _io.write_compressed(zgrp, leading_shape, global_leading_shape, stream_size, stream_starts,
                                global_stream_starts, stream_nbytes, stream_offsets, stream_gains, compressed, n_channels,
                                local_nbytes, global_nbytes, global_process_nbytes, mpi_comm, mpi_dist,
                                software_version=__version__)


@function_timer
def write_array(arr, zgrp, level=5, quanta=None, precision=None, mpi_comm=None, use_threads=False):
    return _io.write_array(arr, zgrp, level=level, quanta=quanta, precision=precision, mpi_comm=mpi_comm,
                           use_threads=use_threads)


@function_timer
def read_compressed(zgrp, keep=None, mpi_comm=None, mpi_dist=None):
    return _io.read_compressed(zgrp, keep=keep, mpi_comm=mpi_comm, mpi_dist=mpi_dist)


@function_timer
def read_array(zgrp, keep=None, stream_slice=None, keep_indices=False, mpi_comm=None, mpi_dist=None, use_threads=False):
    return _io.read_array(zgrp, keep=keep, stream_slice=stream_slice, keep_indices=keep_indices, mpi_comm=mpi_comm,
                          mpi_dist=mpi_dist, use_threads=use_threads, no_flatten=False)
