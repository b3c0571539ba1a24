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
    sends_tensors = True    # torch tensors (host or device) inside a sent object arrive as numpy arrays

    def _split(self, obj, big):
        if isinstance(obj, np.ndarray) and obj.nbytes >= self._BIG:
            big.append(np.ascontiguousarray(obj))
            return ("__ndarray__", len(big) - 1, obj.dtype.str, obj.shape)
        if torch is not None and isinstance(obj, torch.Tensor):
            if obj.numel() * obj.element_size() < self._BIG:
                return obj.detach().cpu().numpy()
            # stays where it is (a CUDA tensor goes on the wire straight from device memory)
            big.append(obj.detach().contiguous().reshape(-1).view(torch.uint8))
            return ("__ndarray__", len(big) - 1, np.dtype(str(obj.dtype).replace("torch.", "")).str, tuple(obj.shape))
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
        dist.send_object_list([(skel, [int(b.nbytes) if isinstance(b, np.ndarray) else int(b.numel()) for b in big])],
                              dst=dest, group=self.group)
        for b in big:
            flat = torch.from_numpy(b.reshape(-1).view(np.uint8)) if isinstance(b, np.ndarray) else b
            for o in range(0, flat.numel(), self._CHUNK):
                piece = flat[o:o + self._CHUNK]
                if piece.device != self.device:
                    piece = piece.to(self.device)
                dist.send(piece, dst=dest, group=self.group)

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


def _even_split(n_elem, parts):
    """Contiguous [first, last) ranges of `n_elem` items over `parts` owners, the first n % parts one longer (the
    np.array_split rule the reference uses, mpi.py:84-90)."""
    base, extra = divmod(int(n_elem), parts)
    edges = [p * base + min(p, extra) for p in range(parts + 1)]
    return [(edges[p], edges[p + 1]) for p in range(parts)]


def distribute_and_verify(mpi_comm, n_elem, mpi_dist=None):
    """The leading-axis distribution [(first, last), ...] over the communicator's ranks: computed when `mpi_dist` is
    None, otherwise checked (one non-empty range per rank, contiguous, covering [0, n_elem)) -- mpi.py:33-90."""
    nproc = 1 if mpi_comm is None else mpi_comm.size
    if mpi_dist is None:
        if n_elem < nproc:
            raise RuntimeError(f"Cannot distribute {n_elem} streams among {nproc} processes.")
        return _even_split(n_elem, nproc)
    if len(mpi_dist) != nproc:
        if mpi_comm is None:
            raise RuntimeError("mpi_comm is None and mpi_dist does not contain single range of all elements")
        raise RuntimeError(f"If specified, mpi_dist (len={len(mpi_dist)}) should have same length as comm size ({nproc})")
    if mpi_dist[0][0] != 0 or mpi_dist[-1][1] != n_elem:
        if mpi_comm is None:
            raise RuntimeError("mpi_comm is None and mpi_dist does not contain single range of all elements")
        raise RuntimeError(f"If specified, mpi_dist ({mpi_dist[0][0]} ... {mpi_dist[-1][1]}) should span the full "
                           f"range of elements ({n_elem})")
    for proc, (lo, hi) in enumerate(mpi_dist):
        if proc > 0 and lo != mpi_dist[proc - 1][1]:
            raise RuntimeError("mpi_dist must have contiguous ranges of first, last (exclusive)")
        if hi <= lo and nproc > 1:
            raise RuntimeError(f"mpi_dist has no data for process {proc}")
    return mpi_dist


def global_array_properties(local_shape, mpi_comm):
    """{"shape": global shape, "dist": [(first, last) per rank]} from every rank's local shape (mpi.py:93-153).  The
    ranks hold consecutive blocks of the leading axis; a 1-D local array counts as one stream.  Trailing dimensions
    must agree everywhere."""
    mine = tuple(int(x) for x in local_shape)
    if len(mine) == 1:
        mine = (1,) + mine
    shapes = [mine] if mpi_comm is None else [tuple(s) for s in mpi_comm.allgather(mine)]
    if any(s[1:] != shapes[0][1:] for s in shapes):
        raise RuntimeError("Inconsistent array dimensions across processes")
    edges = np.concatenate([[0], np.cumsum([s[0] for s in shapes])])
    return {"shape": (int(edges[-1]),) + shapes[0][1:],
            "dist": [(int(edges[p]), int(edges[p + 1])) for p in range(len(shapes))]}


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
