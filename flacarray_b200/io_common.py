"""FlacArray group layout (format version 1) shared by the HDF5 and Zarr front ends.

Layout written by the reference (hdf5.py:96-308, zarr.py:145-304; names hdf5_load_v1.py:22-30):

    group.attrs   flacarray_format_version = "1", flacarray_software_version, flac_channels = "1" | "2"
    stream_starts int64  [global leading shape]   (attrs: stream_size)   global byte offset of each stream
    stream_bytes  int64  [global leading shape]                          compressed bytes of each stream
    stream_offsets / stream_gains  float32|float64 [global leading shape]  (float data only)
    compressed    uint8  [global_nbytes]                                 all streams back to back

A single stream is stored with leading shape (1,).  The functions here work on any object with the
group protocol (h5py.Group, zarr.Group, flacarray_b200.memgroup.MemGroup).

Distributed use: one process per GPU, `mpi_comm` is a TorchComm (or any mpi4py-style communicator with
rank/size/bcast/send/recv).  Like the reference's serial-HDF5 path (io_common.py:262-595) rank 0 owns
the group and the other ranks ship their blocks to it; when every rank passes a group handle each rank
writes / reads its own slices directly (the reference's parallel-HDF5 and Zarr behaviour).
"""
import numpy as np

from .utils import keep_select, select_keep_indices

FORMAT_VERSION = "1"
READ_VERSIONS = ("0", "1")
NAMES = {
    "compressed": "compressed",
    "stream_starts": "stream_starts",
    "stream_bytes": "stream_bytes",
    "stream_size": "stream_size",
    "stream_offsets": "stream_offsets",
    "stream_gains": "stream_gains",
    "flac_channels": "flac_channels",
}


def _host(x):
    """numpy view of numpy / torch (any device) data; None stays None."""
    if x is None:
        return None
    if hasattr(x, "detach"):
        return x.detach().cpu().numpy()
    return np.asarray(x)


def _create(grp, name, shape, dtype):
    if hasattr(grp, "create_dataset"):
        return grp.create_dataset(name, shape=tuple(int(s) for s in shape), dtype=np.dtype(dtype))
    return grp.create_array(name, shape=tuple(int(s) for s in shape), dtype=np.dtype(dtype))


def _write(ds, buf, dest):
    buf = np.ascontiguousarray(buf)
    if buf.size == 0:
        return
    if hasattr(ds, "write_direct"):
        ds.write_direct(buf, None, dest)
    else:
        ds[dest] = buf


def _read(ds, src, dtype=None):
    shape = tuple(s.stop - s.start for s in src)
    out = np.empty(shape, dtype=np.dtype(ds.dtype if dtype is None else dtype))
    if out.size == 0:
        return out
    if hasattr(ds, "read_direct"):
        ds.read_direct(out, src, None)
    else:
        out[...] = ds[src]
    return out


def _rank_size(comm):
    if comm is None:
        return 0, 1
    return comm.rank, comm.size


def _all_have_group(grp, comm):
    """True when every rank holds a group handle (each rank then does its own I/O)."""
    if comm is None:
        return True
    flags = comm.allgather(1 if grp is not None else 0)
    return all(int(f) == 1 for f in flags)


def _h5_driver(grp):
    """Driver name of the file behind an h5py group ("mpio", "sec2", ...); None for zarr / in-memory groups."""
    f = getattr(grp, "file", None)
    return getattr(f, "driver", None) if f is not None else None


def _write_mode(grp, comm):
    """How a distributed write proceeds: (direct, collective_create).

    Every rank writing its own slices is only sound when the store allows uncoordinated writers (zarr, in-memory
    groups) or coordinates them itself (h5py over the mpio driver, where dataset creation and attributes are
    COLLECTIVE: every rank has to issue them, hdf5.py:247-268 of the reference does the same).  Serial h5py handles
    on several ranks would have independent processes write one file: those go through the serial writer, with rank 0
    the only one touching its handle."""
    direct = _all_have_group(grp, comm)
    if comm is None or not direct:
        return direct, False
    drivers = comm.allgather(_h5_driver(grp))
    if all(d is None for d in drivers):
        return True, False
    if all(d == "mpio" for d in drivers):
        return True, True
    return False, False


def _block_slices(dist_range, aux_shape):
    return (slice(dist_range[0], dist_range[1]),) + tuple(slice(0, int(x)) for x in aux_shape[1:])


def write_compressed(grp, leading_shape, global_leading_shape, stream_size, stream_starts, global_stream_starts,
                     stream_nbytes, stream_offsets, stream_gains, compressed, n_channels, local_nbytes, global_nbytes,
                     global_process_nbytes, mpi_comm, mpi_dist, software_version="0"):
    """Write one (possibly distributed) compressed array into `grp` (same arguments as the reference's
    hdf5.write_compressed / zarr.write_compressed)."""
    rank, nproc = _rank_size(mpi_comm)
    aux_global = tuple(int(x) for x in global_leading_shape)
    aux_local = tuple(int(x) for x in leading_shape)
    if len(aux_global) == 0:
        aux_global, aux_local = (1,), (1,)
    gstarts = _host(global_stream_starts).astype(np.int64).reshape(aux_local)
    nbytes = _host(stream_nbytes).astype(np.int64).reshape(aux_local)
    offs = None if stream_offsets is None else _host(stream_offsets).reshape(aux_local)
    gains = None if stream_gains is None else _host(stream_gains).reshape(aux_local)
    if mpi_dist is None:
        mpi_dist = [(0, aux_global[0])]
    direct, collective_create = _write_mode(grp, mpi_comm)
    # a rank that only ships its block to the writer keeps device-resident bytes on the device: a communicator that can
    # move tensors (TorchComm over NCCL) sends them from there, instead of device -> host -> device -> wire
    ship_dev = (not direct and rank != 0 and hasattr(compressed, "is_cuda") and compressed.is_cuda
                and getattr(mpi_comm, "sends_tensors", False))
    comp = compressed.reshape(-1) if ship_dev else _host(compressed).reshape(-1)

    dsets = None
    if rank == 0 or collective_create:
        if grp is None:
            raise RuntimeError("rank 0 needs the group handle")
        grp.attrs["flacarray_format_version"] = FORMAT_VERSION
        grp.attrs["flacarray_software_version"] = str(software_version)
        grp.attrs[NAMES["flac_channels"]] = f"{int(n_channels)}"
        ds_starts = _create(grp, NAMES["stream_starts"], aux_global, np.int64)
        ds_starts.attrs[NAMES["stream_size"]] = int(stream_size)
        _create(grp, NAMES["stream_bytes"], aux_global, np.int64)
        if offs is not None:
            _create(grp, NAMES["stream_offsets"], aux_global, offs.dtype)
        if gains is not None:
            _create(grp, NAMES["stream_gains"], aux_global, gains.dtype)
        _create(grp, NAMES["compressed"], (int(global_nbytes),), np.uint8)
    if mpi_comm is not None and direct:
        mpi_comm.barrier()   # the datasets exist before any other rank touches them

    byte_off = [0]
    for p in range(nproc):
        byte_off.append(byte_off[-1] + int(global_process_nbytes[p]))

    def put(g, p, block):
        b_starts, b_nbytes, b_offs, b_gains, b_comp = block
        sl = _block_slices(mpi_dist[p], aux_global)
        _write(g[NAMES["stream_starts"]], b_starts, sl)
        _write(g[NAMES["stream_bytes"]], b_nbytes, sl)
        if b_offs is not None:
            _write(g[NAMES["stream_offsets"]], b_offs, sl)
        if b_gains is not None:
            _write(g[NAMES["stream_gains"]], b_gains, sl)
        _write(g[NAMES["compressed"]], b_comp, (slice(byte_off[p], byte_off[p] + b_comp.size),))

    mine = (gstarts, nbytes, offs, gains, comp)
    if direct:
        put(grp, rank, mine)
        if mpi_comm is not None:
            mpi_comm.barrier()
        return
    # serial writer: rank 0 receives one block at a time (bounded memory) and writes it
    if rank == 0:
        put(grp, 0, mine)
        for p in range(1, nproc):
            put(grp, p, mpi_comm.recv(source=p))
    else:
        mpi_comm.send(mine, dest=0)
    mpi_comm.barrier()


def _read_block(grp, meta, dist_range, keep_block):
    """One rank's block of a version-1 group: (local_shape, starts, nbytes, offsets, gains, compressed, indices)."""
    aux_global = meta["aux_shape"]
    sl = _block_slices(dist_range, aux_global)
    lead = tuple(s.stop - s.start for s in sl)
    raw_starts = _read(grp[NAMES["stream_starts"]], sl)
    raw_nbytes = _read(grp[NAMES["stream_bytes"]], sl)
    raw_offs = _read(grp[NAMES["stream_offsets"]], sl) if meta["has_offsets"] else None
    raw_gains = _read(grp[NAMES["stream_gains"]], sl) if meta["has_gains"] else None
    dcomp = grp[NAMES["compressed"]]
    if keep_block is None:
        total = int(raw_nbytes.sum())
        if total == 0 or raw_starts.size == 0:
            return (None, None, None, None, None, None, None)
        first = int(raw_starts.reshape(-1)[0])
        comp = _read(dcomp, (slice(first, first + total),), np.uint8)
        return (lead + (meta["stream_size"],), raw_starts - first, raw_nbytes, raw_offs, raw_gains, comp, None)
    starts, nbytes, indices = keep_select(np.asarray(keep_block, dtype=bool), raw_starts, raw_nbytes)
    if len(starts) == 0:
        return (None, None, None, None, None, None, None)
    # only the kept streams are read from the file (io_common.py:52-74), packed back to back
    rel = np.zeros_like(starts)
    rel[1:] = np.cumsum(nbytes)[:-1]
    comp = np.empty(int(nbytes.sum()), dtype=np.uint8)
    for i in range(len(starts)):
        comp[rel[i]:rel[i] + nbytes[i]] = _read(dcomp, (slice(int(starts[i]), int(starts[i] + nbytes[i])),), np.uint8)
    return ((len(starts), meta["stream_size"]), rel, nbytes, select_keep_indices(raw_offs, indices),
            select_keep_indices(raw_gains, indices), comp, indices)


def _read_compressed_versioned(grp, keep=None, mpi_comm=None, mpi_dist=None):
    """`read_compressed` plus the format version string as a last element."""
    from .mpi import distribute_and_verify

    rank, nproc = _rank_size(mpi_comm)
    direct = _all_have_group(grp, mpi_comm)
    meta = None
    if grp is not None and (rank == 0 or direct):
        if "flacarray_format_version" not in grp.attrs:
            raise RuntimeError("Group does not contain a FlacArray")
        ver = grp.attrs["flacarray_format_version"]
        ver = ver.decode() if isinstance(ver, bytes) else str(ver)
        if ver not in READ_VERSIONS:
            raise RuntimeError(f"Unsupported FlacArray format version {ver} (this reader handles versions 0 and 1)")
        dstarts = grp[NAMES["stream_starts"]]
        meta = {
            # version 0 predates 2-channel streams and has no channel attribute (hdf5_load_v0.py:262-270)
            "n_channel": 1 if ver == "0" else int(grp.attrs[NAMES["flac_channels"]]),
            "version": ver,
            "stream_size": int(dstarts.attrs[NAMES["stream_size"]]),
            "aux_shape": tuple(int(x) for x in dstarts.shape),
            "has_offsets": NAMES["stream_offsets"] in grp,
            "has_gains": NAMES["stream_gains"] in grp,
        }
    if mpi_comm is not None and not direct:
        meta = mpi_comm.bcast(meta, root=0)
    global_shape = meta["aux_shape"] + (meta["stream_size"],)
    mpi_dist = distribute_and_verify(mpi_comm, global_shape[0], mpi_dist=mpi_dist)
    if keep is not None:
        keep = np.asarray(_host(keep), dtype=bool)
        if keep.shape != meta["aux_shape"]:
            raise RuntimeError("The keep array should have the same shape as the leading dimensions of the array")

    def block_of(g, p):
        kb = None if keep is None else keep[_block_slices(mpi_dist[p], meta["aux_shape"])]
        return _read_block(g, meta, mpi_dist[p], kb)

    if direct:
        blk = block_of(grp, rank)
    elif rank == 0:
        blk = block_of(grp, 0)
        for p in range(1, nproc):
            mpi_comm.send(block_of(grp, p), dest=p)
    else:
        blk = mpi_comm.recv(source=0)
    local_shape, starts, nbytes, offs, gains, comp, indices = blk
    return (local_shape, global_shape, comp, meta["n_channel"], starts, nbytes, offs, gains, mpi_dist, indices,
            meta["version"])


def read_compressed(grp, keep=None, mpi_comm=None, mpi_dist=None):
    """Load (this rank's block of) a compressed array.  Returns the reference's tuple
    (local_shape, global_shape, compressed, n_channel, stream_starts, stream_nbytes, stream_offsets,
    stream_gains, mpi_dist, keep_indices) -- hdf5_load_v1.py:93-246, hdf5_load_v0.py:92-277."""
    return _read_compressed_versioned(grp, keep=keep, mpi_comm=mpi_comm, mpi_dist=mpi_dist)[:-1]


def _decompress_v0(compressed, stream_size, starts, nbytes, offs, gains, first, last, use_threads, no_flatten):
    """Version-0 payloads are always 1-channel FLAC (hdf5_load_v0.py:357-411).  Offsets without gains mark
    the legacy int64 encoding -- 32-bit samples plus one int64 offset per stream; float64 data was stored
    as 32-bit integers with float64 offsets/gains and is restored in float32, then promoted."""
    from .decompress import array_decompress

    kw = dict(first_stream_sample=first, last_stream_sample=last, is_int64=False, use_threads=use_threads,
              no_flatten=no_flatten)
    if offs is not None and gains is None:
        arr = _host(array_decompress(compressed, stream_size, starts, nbytes, **kw))
        return arr.astype(np.int64) + np.asarray(offs).reshape(np.asarray(offs).shape + (1,))
    if offs is None:
        return array_decompress(compressed, stream_size, starts, nbytes, **kw)
    want64 = np.asarray(gains).dtype == np.dtype(np.float64)
    arr = array_decompress(compressed, stream_size, starts, nbytes,
                           stream_offsets=np.asarray(offs, dtype=np.float32),
                           stream_gains=np.asarray(gains, dtype=np.float32), **kw)
    return arr.astype(np.float64) if want64 else arr


def read_array(grp, keep=None, stream_slice=None, keep_indices=False, mpi_comm=None, mpi_dist=None, use_threads=False,
               no_flatten=False):
    """Read and decompress (hdf5_load_v1.py:249-375); the decode runs on the GPU."""
    from .decompress import array_decompress

    (local_shape, global_shape, compressed, n_channel, starts, nbytes, offs, gains, mpi_dist, indices,
     version) = _read_compressed_versioned(grp, keep=keep, mpi_comm=mpi_comm, mpi_dist=mpi_dist)
    first = last = None
    if stream_slice is not None:
        if stream_slice.step is not None and stream_slice.step != 1:
            raise RuntimeError("Only stream slices with a step size of 1 are supported")
        first, last = stream_slice.start, stream_slice.stop
    if compressed is None:
        arr = None      # this rank holds no streams (empty keep selection)
    elif version == "0":
        arr = _decompress_v0(compressed, local_shape[-1], starts, nbytes, offs, gains, first, last, use_threads,
                             no_flatten)
    else:
        arr = array_decompress(compressed, local_shape[-1], starts, nbytes, stream_offsets=offs, stream_gains=gains,
                               first_stream_sample=first, last_stream_sample=last, is_int64=(n_channel == 2),
                               use_threads=use_threads, no_flatten=no_flatten)
    if keep_indices:
        return arr, indices
    return arr


def write_array(arr, grp, level=5, quanta=None, precision=None, mpi_comm=None, use_threads=False):
    """Compress on the GPU and write (hdf5.py:311-398)."""
    from .compress import array_compress
    from .libflacarray import np_dtype
    from .mpi import global_array_properties, global_bytes

    props = global_array_properties(tuple(arr.shape), mpi_comm=mpi_comm)
    n_channels = 2 if np_dtype(arr).itemsize == 8 else 1
    compressed, starts, nbytes, offsets, gains = array_compress(arr, level=level, quanta=quanta, precision=precision,
                                                                use_threads=use_threads)
    compressed, starts, nbytes = _host(compressed), _host(starts), _host(nbytes)
    local_nbytes = int(compressed.size)
    global_nbytes, proc_nbytes, global_starts = global_bytes(local_nbytes, starts, mpi_comm)
    leading = (1,) if len(arr.shape) == 1 else tuple(arr.shape[:-1])
    write_compressed(grp, leading, props["shape"][:-1], arr.shape[-1], starts, global_starts, nbytes, _host(offsets),
                     _host(gains), compressed, n_channels, local_nbytes, global_nbytes, proc_nbytes, mpi_comm,
                     props["dist"])
