"""HDF5 / Zarr group layout (format version 1) without a GPU: the compressed representation comes from the
CPU oracle (test infrastructure), the layout code is the product's (flacarray_b200/io_common.py).
Mirrors the reference's tests/hdf5.py and tests/zarr.py (write -> read -> __eq__), plus keep masks and the
two-rank serial-writer path over torch.distributed (gloo)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_flacarray(oracle, data, comm=None, dist_l=None, global_shape=None):
    from flacarray_b200.array import FlacArray

    d2 = data.reshape(1, -1) if data.ndim == 1 else data.reshape(-1, data.shape[-1])
    comp, starts, nbytes = oracle.encode(d2, 5)
    lead = (1,) if data.ndim == 1 else data.shape[:-1]
    gshape = global_shape if global_shape is not None else ((1,) + data.shape if data.ndim == 1 else data.shape)
    return FlacArray(None, shape=data.shape, global_shape=gshape, compressed=comp, dtype=data.dtype,
                     stream_starts=starts.reshape(lead), stream_nbytes=nbytes.reshape(lead), mpi_comm=comm, mpi_dist=dist_l)


@pytest.mark.parametrize("zarr_style", [False, True])
def test_group_layout_roundtrip(oracle, zarr_style):
    from flacarray_b200 import hdf5 as fh5
    from flacarray_b200 import zarr as fzr
    from flacarray_b200.array import FlacArray
    from flacarray_b200.memgroup import MemGroup

    mod = fzr if zarr_style else fh5
    rng = np.random.default_rng(3)
    for data in (np.cumsum(rng.integers(-99, 100, (3, 4, 3000)), axis=-1).astype(np.int32),
                 np.cumsum(rng.integers(-2 ** 33, 2 ** 33, (5, 2500)), axis=-1).astype(np.int64),
                 rng.integers(-1000, 1000, 7000).astype(np.int32)):
        far = _oracle_flacarray(oracle, data)
        grp = MemGroup(zarr_style=zarr_style)
        if zarr_style:
            far.write_zarr(grp)
        else:
            far.write_hdf5(grp)
        # the layout the reference writes (hdf5.py:194-247, hdf5_load_v1.py:22-30)
        assert grp.attrs["flacarray_format_version"] == "1"
        assert grp.attrs["flac_channels"] == ("2" if data.dtype == np.int64 else "1")
        lead = (1,) if data.ndim == 1 else data.shape[:-1]
        assert grp["stream_starts"].shape == lead and grp["stream_starts"].dtype == np.int64
        assert grp["stream_starts"].attrs["stream_size"] == data.shape[-1]
        assert grp["stream_bytes"].shape == lead
        assert grp["compressed"].shape == (far.nbytes,) and grp["compressed"].dtype == np.uint8
        assert "stream_offsets" not in grp and "stream_gains" not in grp
        back = FlacArray.read_zarr(grp) if zarr_style else FlacArray.read_hdf5(grp)
        assert back == far and back.shape == data.shape and back.dtype == data.dtype
        # module-level reader returns the reference's 10-tuple
        tup = mod.read_compressed(grp)
        assert tup[0] == ((1,) + data.shape if data.ndim == 1 else data.shape) and tup[3] == (2 if data.dtype == np.int64 else 1)
        assert np.array_equal(oracle.decode(tup[2], tup[4].reshape(-1), tup[5].reshape(-1), data.shape[-1],
                                            is_int64=data.dtype == np.int64).reshape(data.shape), data)
        if data.ndim == 3:
            keep = np.zeros(data.shape[:-1], bool)
            keep[0, 1] = keep[2, 3] = keep[1, 0] = True
            tup = mod.read_compressed(grp, keep=keep)
            assert tup[0] == (3, data.shape[-1]) and tup[9] == [(0, 1), (1, 0), (2, 3)]
            got = oracle.decode(tup[2], tup[4], tup[5], data.shape[-1])
            assert np.array_equal(got, data[keep])
            empty = mod.read_compressed(grp, keep=np.zeros(data.shape[:-1], bool))
            assert empty[2] is None and empty[0] is None
            with pytest.raises(RuntimeError):
                mod.read_compressed(grp, keep=np.zeros((2, 2), bool))
    with pytest.raises(RuntimeError):
        FlacArray.read_hdf5(MemGroup())


def test_float_aux_datasets(oracle):
    """offsets / gains datasets keep the float dtype of the original data (hdf5.py:218-235)."""
    from flacarray_b200.array import FlacArray
    from flacarray_b200.memgroup import MemGroup

    rng = np.random.default_rng(4)
    f = rng.normal(0, 1, (4, 3000)).astype(np.float32)
    ints, off, gain = oracle.float_to_int(f, np.full(4, 1e-4, np.float32))
    comp, starts, nbytes = oracle.encode(ints, 5)
    far = FlacArray(None, shape=f.shape, global_shape=f.shape, compressed=comp, dtype=np.float32, stream_starts=starts,
                    stream_nbytes=nbytes, stream_offsets=off, stream_gains=gain)
    grp = MemGroup()
    far.write_hdf5(grp)
    assert grp["stream_offsets"].dtype == np.float32 and grp["stream_gains"].shape == (4,)
    back = FlacArray.read_hdf5(grp)
    assert back == far and back.dtype == np.float32


def _v0_group(oracle, ints32, offsets=None, gains=None, zarr_style=False):
    """A format-version-0 group as the first flacarray releases wrote it (hdf5_load_v0.py:22-31): same
    dataset names as version 1, 1-channel FLAC only, and no `flac_channels` attribute."""
    from flacarray_b200.memgroup import MemGroup

    comp, starts, nbytes = oracle.encode(ints32.reshape(-1, ints32.shape[-1]), 5)
    lead = ints32.shape[:-1]
    grp = MemGroup(zarr_style=zarr_style)
    grp.attrs["flacarray_format_version"] = "0"
    grp.attrs["flacarray_software_version"] = "0.1.0"
    ds = grp.create_dataset("stream_starts", data=starts.reshape(lead))
    ds.attrs["stream_size"] = ints32.shape[-1]
    grp.create_dataset("stream_bytes", data=nbytes.reshape(lead))
    grp.create_dataset("compressed", data=comp)
    if offsets is not None:
        grp.create_dataset("stream_offsets", data=offsets)
    if gains is not None:
        grp.create_dataset("stream_gains", data=gains)
    return grp


def test_version0_groups_are_readable(oracle):
    """hdf5.py:430-446 dispatches on the version attribute; version 0 has one channel and no channel attr."""
    from flacarray_b200 import hdf5 as fh5
    from flacarray_b200.memgroup import MemGroup

    rng = np.random.default_rng(8)
    ints = np.cumsum(rng.integers(-50, 51, (2, 3, 2000)), axis=-1).astype(np.int32)
    offs = rng.integers(-2 ** 40, 2 ** 40, (2, 3)).astype(np.int64)
    grp = _v0_group(oracle, ints, offsets=offs)
    tup = fh5.read_compressed(grp)
    assert len(tup) == 10 and tup[0] == ints.shape and tup[3] == 1
    assert np.array_equal(tup[6], offs) and tup[7] is None
    assert np.array_equal(oracle.decode(tup[2], tup[4].reshape(-1), tup[5].reshape(-1), 2000).reshape(ints.shape), ints)
    keep = np.zeros((2, 3), bool)
    keep[1, 2] = True
    tup = fh5.read_compressed(grp, keep=keep)
    assert tup[0] == (1, 2000) and np.array_equal(tup[6], offs[keep])
    bad = MemGroup()
    bad.attrs["flacarray_format_version"] = "7"
    with pytest.raises(RuntimeError):
        fh5.read_compressed(bad)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    try:
        if ROOT not in sys.path:
            sys.path.insert(0, ROOT)
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        import torch.distributed as dist

        dist.init_process_group("gloo", rank=rank, world_size=world)
        from flacarray_b200.array import FlacArray
        from flacarray_b200.memgroup import MemGroup
        from flacarray_b200.mpi import TorchComm, global_array_properties
        from oracle import oracle as O

        comm = TorchComm()
        comm._BIG = 1 << 10          # the compressed blocks of this test travel as raw tensors in chunks
        comm._CHUNK = 3000
        rng = np.random.default_rng(5)
        n_total, L = 5, 4000
        full = np.cumsum(rng.integers(-500, 501, (n_total, L)), axis=1).astype(np.int32)
        lo, hi = [(0, 3), (3, 5)][rank]
        local = full[lo:hi]
        props = global_array_properties(local.shape, comm)
        far = _oracle_flacarray(O, local, comm=comm, dist_l=props["dist"], global_shape=props["shape"])
        # serial writer: only rank 0 holds the group; rank 1 ships its block
        grp = MemGroup() if rank == 0 else None
        far.write_hdf5(grp)
        if rank == 0:
            assert grp["stream_starts"].shape == (n_total,) and grp["compressed"].shape == (far.global_nbytes,)
            st, nb = grp["stream_starts"][...], grp["stream_bytes"][...]
            assert np.array_equal(O.decode(grp["compressed"][...], st, nb, L), full)
        # distributed read with a different distribution and a keep mask
        back = FlacArray.read_hdf5(grp, mpi_comm=comm, mpi_dist=[(0, 2), (2, 5)])
        blo, bhi = [(0, 2), (2, 5)][rank]
        assert back.shape == (bhi - blo, L) and back.global_shape == (n_total, L)
        got = O.decode(back.compressed, back.stream_starts, back.stream_nbytes, L)
        assert np.array_equal(got, full[blo:bhi])
        keep = np.array([True, False, False, True, True])
        from flacarray_b200 import hdf5 as fh5

        tup = fh5.read_compressed(grp, keep=keep, mpi_comm=comm)
        klo, khi = tup[8][rank]
        want = full[klo:khi][keep[klo:khi]]
        assert np.array_equal(O.decode(tup[2], tup[4], tup[5], L), want)
        comm.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except BaseException as e:  # noqa: BLE001
        import traceback

        q.put((rank, "FAIL: " + "".join(traceback.format_exception(type(e), e, e.__traceback__))))


def test_two_rank_serial_writer(oracle):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}: {msg}"


def test_hdf5_utils_serial_rules():
    """hdf5_utils.py:60-80: a group that is not open on every rank means rank 0 does the I/O."""
    from flacarray_b200 import hdf5_utils as hu

    class _Comm:
        def __init__(self, size, have):
            self.size, self._have, self.rank = size, have, 0

        def allgather(self, v):
            return self._have

    assert hu.hdf5_use_serial(None, None) and hu.hdf5_use_serial(object(), _Comm(1, [1]))
    assert hu.hdf5_use_serial(object(), _Comm(2, [1, 0])) and not hu.hdf5_use_serial(object(), _Comm(2, [1, 1]))
    assert hu.have_hdf5_parallel() is False
    hu.check_dataset_buffer_size("x", (slice(0, 4),), np.dtype(np.int64), False)
    if not hu.have_hdf5:
        with pytest.raises(RuntimeError):
            hu.H5File("/nonexistent/x.h5", "r")


def test_host_copy_pool_matches_plain_copy():
    """libflacarray._host_copy: with one intra-op thread (torchrun's default) large staging copies are
    split over a thread pool; the bytes must be the same as a plain copy."""
    import torch

    from flacarray_b200 import libflacarray as lf

    src = torch.from_numpy(np.random.default_rng(5).integers(0, 255, (37, 1 << 20), dtype=np.uint8))   # 37 MB, odd split
    dst = torch.zeros_like(src)
    old = torch.get_num_threads()
    try:
        torch.set_num_threads(1)
        lf._host_copy(dst, src)
    finally:
        torch.set_num_threads(old)
    assert torch.equal(dst, src)
    small = torch.zeros(1000, dtype=torch.float32)
    lf._host_copy(small, torch.arange(1000, dtype=torch.float32))
    assert float(small[-1]) == 999.0
