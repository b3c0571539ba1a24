"""HDF5 file helpers with the names of /root/reference/src/flacarray/hdf5_utils.py:25-168.

Processes here are torch.distributed ranks (one per GPU), not MPI ranks, so h5py's `mpio` driver never
applies: files are opened by rank 0 and the other ranks hold `None` -- the serial-writer mode of
`io_common` (reference hdf5.py:247-308).
"""
from .utils import log

try:
    import h5py

    have_hdf5 = True
except Exception:  # pragma: no cover - optional dependency
    h5py = None
    have_hdf5 = False


def have_hdf5_parallel():
    """hdf5_utils.py:29-57: parallel HDF5 needs mpi4py communicators; never available over torch.distributed."""
    return False


def hdf5_use_serial(hgrp, mpi_comm):
    """True when the group is not open on every rank (hdf5_utils.py:60-80)."""
    if mpi_comm is None or mpi_comm.size == 1:
        return True
    have = mpi_comm.allgather(1 if hgrp is not None else 0)
    return sum(have) != mpi_comm.size


def hdf5_open(path, mode, comm=None, force_serial=False):
    """Open `path` on rank 0; other ranks get None (hdf5_utils.py:83-119, serial branch)."""
    if not have_hdf5:
        raise RuntimeError("h5py is not importable")
    rank = 0 if comm is None else comm.rank
    if rank != 0:
        return None
    log.debug(f"Opened file {path} serially")
    return h5py.File(path, mode)


class H5File(object):
    """Context manager around `hdf5_open`; `.handle` is None away from rank 0 (hdf5_utils.py:122-146)."""

    def __init__(self, name, mode, comm=None, force_serial=False):
        self.handle = hdf5_open(name, mode, comm=comm, force_serial=force_serial)

    def close(self):
        if getattr(self, "handle", None) is not None:
            self.handle.flush()
            self.handle.close()
            self.handle = None

    def __del__(self):
        self.close()

    def __enter__(self):
        return self

    def __exit__(self, *args):
        self.close()


def check_dataset_buffer_size(msg, slices, dtype, parallel):
    """Warn about > 2 GiB buffers under parallel HDF5 (hdf5_utils.py:149-168); a no-op in serial mode."""
    if not parallel:
        return
    nelem = 1
    for slc in slices:
        nelem *= slc.stop - slc.start
    nbytes = nelem * dtype.itemsize
    if nbytes >= 2147483647:
        log.warning(f"{msg}:  buffer size of {nbytes} bytes > 2^31 - 1.   HDF5 parallel I/O will likely fail.")
