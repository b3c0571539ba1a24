"""N > 1 path on CPU: two processes over torch.distributed (gloo) exercise the host logic that one-GPU-per-
process runs use (flacarray_b200/mpi.py, reference mpi.py:33-187): leading-axis distribution, global
shape negotiation, and the ONLY collective of the hot path -- the all-gather of per-rank compressed byte
counts that turns local stream starts into global ones.  The compressed bytes come from the CPU oracle
(test infrastructure); no GPU kernel runs here.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


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
        from flacarray_b200.mpi import TorchComm, distribute_and_verify, global_array_properties, global_bytes
        from oracle import oracle as O

        comm = TorchComm()
        assert comm.rank == rank and comm.size == world
        # ---- distribution rule (mpi.py:84-90): first n % size ranks get one extra
        n_total, L = 7, 6000
        dist_l = distribute_and_verify(comm, n_total)
        assert dist_l == [(0, 4), (4, 7)]
        assert distribute_and_verify(comm, n_total, mpi_dist=dist_l) == dist_l
        with pytest.raises(RuntimeError):
            distribute_and_verify(comm, n_total, mpi_dist=[(0, 3), (4, 7)])      # gap
        with pytest.raises(RuntimeError):
            distribute_and_verify(comm, 1)                                        # fewer streams than ranks
        lo, hi = dist_l[rank]
        # ---- this rank's shard, compressed by the oracle
        rng = np.random.default_rng(99)
        full = (np.cumsum(rng.integers(-500, 501, (n_total, L)), axis=1)).astype(np.int32)
        local = full[lo:hi]
        comp, starts, nbytes = O.encode(local, 5)
        props = global_array_properties(local.shape, comm)
        assert props["shape"] == (n_total, L) and props["dist"] == dist_l
        # ---- the collective: byte counts -> global offsets (mpi.py:156-187)
        gtot, per_rank, gstarts = global_bytes(comp.size, starts, comm)
        all_sizes = comm.allgather(int(comp.size))
        assert per_rank == all_sizes and gtot == sum(all_sizes)
        assert np.array_equal(gstarts, starts + sum(all_sizes[:rank]))
        far = FlacArray(None, shape=local.shape, global_shape=props["shape"], compressed=comp, dtype=np.int32,
                        stream_starts=starts, stream_nbytes=nbytes, mpi_comm=comm, mpi_dist=props["dist"])
        assert far.global_nbytes == gtot and far.global_process_nbytes == all_sizes
        assert far.global_shape == (n_total, L) and far.nstreams == hi - lo and far.global_nstreams == n_total
        assert np.array_equal(far.global_stream_starts, gstarts)
        assert f"Rank {rank:04d}" in repr(far)
        # concatenating the shards at the global offsets gives one valid global container
        pieces = comm.gather((gstarts, nbytes, comp), root=0)
        if rank == 0:
            blob = np.zeros(gtot, np.uint8)
            gs, gn = [], []
            for s_, n_, c_ in pieces:
                blob[s_[0]:s_[0] + c_.size] = c_
                gs.append(s_); gn.append(n_)
            gs, gn = np.concatenate(gs), np.concatenate(gn)
            assert np.array_equal(O.decode(blob, gs, gn, L), full)
        # inconsistent trailing shapes are detected on every rank (mpi.py:147-152)
        with pytest.raises(RuntimeError):
            global_array_properties((2, L + rank), comm)
        comm.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except BaseException as e:  # noqa: BLE001
        import traceback

        q.put((rank, "FAIL: " + "".join(traceback.format_exception(type(e), e, e.__traceback__))))


def test_two_rank_byte_count_allgather():
    import torch.multiprocessing as mp

    from oracle import oracle as O

    O.lib()          # build the checker once, before the workers race for it
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
