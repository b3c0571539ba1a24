"""HDF5 front end of the FlacArray group layout: same functions and arguments as
/root/reference/src/flacarray/hdf5.py:96-525 (`write_compressed`, `write_array`, `read_compressed`,
`read_array`) on top of `io_common`.  Works with h5py groups when h5py is installed and with any
object implementing the group protocol (e.g. `memgroup.MemGroup`); format versions 0 and 1
are read (hdf5_load_v0.py / hdf5_load_v1.py), version 1 is written.

Distributed arrays: see io_common -- rank 0 owns the file unless every rank passes a handle
(MPI-enabled h5py), in which case every rank writes its own hyperslabs.
"""
from . import io_common as _io
from .utils import function_timer

try:
    import h5py  # noqa: F401

    have_hdf5 = True
except Exception:  # pragma: no cover - optional dependency
    have_hdf5 = False

hdf5_names = dict(_io.NAMES)


@function_timer
def write_compressed(hgrp, leading_shape, global_leading_shape, stream_size, stream_starts, global_stream_starts,
                     stream_nbytes, stream_offsets, stream_gains, compressed, n_channels, local_nbytes, global_nbytes,
                     global_process_nbytes, mpi_comm, mpi_dist):
    from . import __version__

    return _io.write_compressed(hgrp, leading_shape, global_leading_shape, stream_size, stream_starts,
                                global_stream_starts, stream_nbytes, stream_offsets, stream_gains, compressed, n_channels,
                                local_nbytes, global_nbytes, global_process_nbytes, mpi_comm, mpi_dist,
                                software_version=__version__)


@function_timer
def write_array(arr, hgrp, level=5, quanta=None, precision=None, mpi_comm=None, use_threads=False):
    return _io.write_array(arr, hgrp, level=level, quanta=quanta, precision=precision, mpi_comm=mpi_comm,
                           use_threads=use_threads)


@function_timer
def read_compressed(hgrp, keep=None, mpi_comm=None, mpi_dist=None):
    return _io.read_compressed(hgrp, keep=keep, mpi_comm=mpi_comm, mpi_dist=mpi_dist)


@function_timer
def read_array(hgrp, keep=None, stream_slice=None, keep_indices=False, mpi_comm=None, mpi_dist=None, use_threads=False):
    return _io.read_array(hgrp, keep=keep, stream_slice=stream_slice, keep_indices=keep_indices, mpi_comm=mpi_comm,
                          mpi_dist=mpi_dist, use_threads=use_threads, no_flatten=False)
