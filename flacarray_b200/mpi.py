"""Distribution helpers with the interface of /root/reference/src/flacarray/mpi.py:33-187.

The reference distributes the leading axis over an mpi4py communicator.  Here one process drives one
GPU and the communicator is `TorchComm`, a thin adapter over torch.distributed (NCCL on B200, gloo on
CPU) exposing the mpi4py-style members the reference's helpers use (rank, size, allgather, gather,
bcast).  Streams are independent, so the ONLY exchange on the hot path is the all-gather of each
rank's compressed byte count (mpi.py:177) -- one int64 per rank, sent as a device tensor over NCCL.
A real mpi4py communicator also works, since only those members are used.
"""
import numpy as np

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None

use_mpi = False
MPI = None


class TorchComm:
    """mpi4py-like view of a torch.distributed process group."""

    def __init__(self, group=None, device=None):
        if dist is None or not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)
        backend = dist.get_backend(group)
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        self.device = device

    def Get_rank(self):
        return self.rank

    def Get_size(self):
        return self.size

    def allgather_int64(self, value):
        """All-gather one int64 per rank as a tensor collective (NCCL over NVLink on B200)."""
        t = torch.tensor([int(value)], dtype=torch.int64, device=self.device)
        out = torch.empty(self.size, dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(out, t, group=self.group)
        return [int(x) for x in out.cpu().tolist()]

    def allgather(self, obj):
        if isinstance(obj, (int, np.integer)):
            return self.allgather_int64(obj)
        out = [None] * self.size
        dist.all_gather_object(out, obj, group=self.group)
        return out

    def gather(self, obj, root=0):
        out = [None] * self.size if self.rank == root else None
        dist.gather_object(obj, out, dst=root, group=self.group)
        return out

    def bcast(self, obj, root=0):
        box = [obj]
        dist.broadcast_object_list(box, src=root, group=self.group)
        return box[0]

    # Point-to-point transfer of (possibly nested tuples / lists of) numpy arrays, used by the serial-writer
    # file I/O (io_common.py).  Small objects are pickled; arrays above `_BIG` bytes travel as raw uint8
    # tensors in bounded chunks (through device memory over NCCL, directly from host memory over gloo), so
    # a rank's multi-GB compressed block is never pickled.
    _BIG = 1 << 20
    _CHUNK = 256 << 20

    def _split(self, obj, big):
        if isinstance(obj, np.ndarray) and obj.nbytes >= self._BIG:
            big.append(np.ascontiguousarray(obj))
            return ("__ndarray__", len(big) - 1, obj.dtype.str, obj.shape)
        if isinstance(obj, (tuple, list)):
            return type(obj)(self._split(o, big) for o in obj)
        return obj

    def _join(self, obj, big):
        if isinstance(obj, tuple) and len(obj) == 4 and obj[0] == "__ndarray__":
            return big[obj[1]].view(np.dtype(obj[2])).reshape(obj[3])
        if isinstance(obj, (tuple, list)):
            return type(obj)(self._join(o, big) for o in obj)
        return obj

    def send(self, obj, dest):
        big = []
        skel = self._split(obj, big)
        dist.send_object_list([(skel, [b.nbytes for b in big])], dst=dest, group=self.group)
        for b in big:
            flat = torch.from_numpy(b.reshape(-1).view(np.uint8))
            for o in range(0, flat.numel(), self._CHUNK):
                piece = flat[o:o + self._CHUNK]
                dist.send(piece.to(self.device) if self.device.type == "cuda" else piece, dst=dest, group=self.group)

    def recv(self, source):
        box = [None]
        dist.recv_object_list(box, src=source, group=self.group)
        skel, sizes = box[0]
        big = []
        for nb in sizes:
            host = np.empty(nb, dtype=np.uint8)
            flat = torch.from_numpy(host)
            for o in range(0, nb, self._CHUNK):
                n = min(self._CHUNK, nb - o)
                if self.device.type == "cuda":
                    buf = torch.empty(n, dtype=torch.uint8, device=self.device)
                    dist.recv(buf, src=source, group=self.group)
                    flat[o:o + n].copy_(buf)
                else:
                    dist.recv(flat[o:o + n], src=source, group=self.group)
            big.append(host)
        return self._join(skel, big)

    def barrier(self):
        dist.barrier(group=self.group)


def distribute_and_verify(mpi_comm, n_elem, mpi_dist=None):
    """Compute or verify the distribution of `n_elem` leading-axis elements (mpi.py:33-90)."""
    if mpi_dist is not None:
        if mpi_comm is None:
            if len(mpi_dist) != 1 or mpi_dist[0][0] != 0 or mpi_dist[0][1] != n_elem:
                msg = "mpi_comm is None and mpi_dist does not contain single range "
                msg += "of all elements"
                raise RuntimeError(msg)
            return mpi_dist
        if mpi_comm.size != len(mpi_dist):
            msg = f"If specified, mpi_dist (len={len(mpi_dist)}) should have same "
            msg += f"length as comm size ({mpi_comm.size})"
            raise RuntimeError(msg)
        if mpi_dist[0][0] != 0 or mpi_dist[-1][1] != n_elem:
            msg = f"If specified, mpi_dist ({mpi_dist[0][0]} ... {mpi_dist[-1][1]})"
            msg += f" should span the full range of elements ({n_elem})"
            raise RuntimeError(msg)
        for proc in range(1, mpi_comm.size):
            if mpi_dist[proc][0] != mpi_dist[proc - 1][1]:
                raise RuntimeError("mpi_dist must have contiguous ranges of first, last (exclusive)")
            if mpi_dist[proc][1] <= mpi_dist[proc][0]:
                raise RuntimeError(f"mpi_dist has no data for process {proc}")
        return mpi_dist
    if mpi_comm is None:
        return [(0, n_elem)]
    # uniform split, first n % size ranks get one extra (np.array_split rule, mpi.py:84-90)
    size = mpi_comm.size
    base, extra = divmod(int(n_elem), size)
    if base == 0:
        msg = f"Cannot distribute {n_elem} streams among {size}"
        msg += " processes."
        raise RuntimeError(msg)
    dist_out = []
    off = 0
    for proc in range(size):
        n = base + (1 if proc < extra else 0)
        dist_out.append((off, off + n))
        off += n
    return dist_out


def global_array_properties(local_shape, mpi_comm):
    """Global shape and per-rank leading-axis ranges (mpi.py:93-153)."""
    props = dict()
    local_shape = tuple(int(x) for x in local_shape)
    if mpi_comm is None:
        if len(local_shape) == 1:
            props["shape"] = (1, local_shape[0])
            props["dist"] = [(0, 1)]
        else:
            props["shape"] = local_shape
            props["dist"] = [(0, local_shape[0])]
        return props
    all_shapes = mpi_comm.gather(local_shape, root=0)
    err = False
    if mpi_comm.rank == 0:
        dist_l = list()
        shp = all_shapes[0]
        if len(shp) == 1:
            lda = 1
            trl = shp
        else:
            lda = shp[0]
            trl = shp[1:]
        dist_l.append((0, lda))
        ldoff = lda
        for s in all_shapes[1:]:
            if len(s) == 1:
                lda += 1
                dist_l.append((ldoff, ldoff + 1))
                ldoff += 1
                if s != trl:
                    err = True
                    break
            else:
                lda += s[0]
                dist_l.append((ldoff, ldoff + s[0]))
                ldoff += s[0]
                if s[1:] != trl:
                    err = True
                    break
        props["shape"] = (lda,) + tuple(trl)
        props["dist"] = dist_l
    err = mpi_comm.bcast(err, root=0)
    props = mpi_comm.bcast(props, root=0)
    if err:
        raise RuntimeError("Inconsistent array dimensions across processes")
    return props


def global_bytes(local_nbytes, stream_starts, mpi_comm):
    """(total global bytes, bytes per process, global byte offsets of the local streams) -- mpi.py:156-187."""
    if mpi_comm is None or mpi_comm.size == 1:
        return (local_nbytes, [local_nbytes], stream_starts)
    rank = mpi_comm.rank
    all_nbytes = mpi_comm.allgather(int(local_nbytes))
    global_nbytes = int(np.sum(all_nbytes))
    byte_offset = int(np.sum(all_nbytes[:rank]))
    global_starts = stream_starts + byte_offset
    return (global_nbytes, all_nbytes, global_starts)
