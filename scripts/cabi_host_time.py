"""Throughput of the reference-signature C entry points (host pointers, pageable numpy memory):
encode_i32 / decode_i32 on a (256, 1M) int32 random walk (1 GB).  FLACARRAY_B200_NO_PIPE=1 = plain path."""
import ctypes as C
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from flacarray_b200 import _lib

L = C.CDLL(_lib.SO_PATH)
libc = C.CDLL(None)
n, size = 256, 1000000
rng = np.random.default_rng(1)
x = np.cumsum(rng.integers(-300, 301, (n, size), dtype=np.int32), axis=1, dtype=np.int32)
starts = np.zeros(n, np.int64)
out = np.zeros((n, size), np.int32)
L.encode_i32.restype = C.c_int
L.decode_i32.restype = C.c_int
for it in range(4):
    nb = C.c_int64(0); buf = C.POINTER(C.c_ubyte)()
    t0 = time.perf_counter()
    rc = L.encode_i32(C.c_void_p(x.ctypes.data), C.c_int64(n), C.c_int64(size), C.c_uint32(5), C.byref(nb),
                      C.c_void_p(starts.ctypes.data), C.byref(buf))
    t1 = time.perf_counter()
    assert rc == 0
    comp = np.ctypeslib.as_array(buf, shape=(nb.value,)).copy()
    libc.free(buf)
    nbytes = np.empty(n, np.int64); nbytes[:-1] = np.diff(starts); nbytes[-1] = nb.value - starts[-1]
    t2 = time.perf_counter()
    rc = L.decode_i32(C.c_void_p(comp.ctypes.data), C.c_void_p(starts.ctypes.data), C.c_void_p(nbytes.ctypes.data),
                      C.c_int64(n), C.c_int64(size), C.c_int64(-1), C.c_int64(-1), C.c_void_p(out.ctypes.data), C.c_bool(True))
    t3 = time.perf_counter()
    assert rc == 0 and np.array_equal(out, x)
    print(f"it{it}: encode_i32 {x.nbytes / (t1 - t0) / 1e9:.1f} GB/s ({1e3 * (t1 - t0):.0f} ms), "
          f"decode_i32 {x.nbytes / (t3 - t2) / 1e9:.1f} GB/s ({1e3 * (t3 - t2):.0f} ms), ratio {nb.value / x.nbytes:.3f}", flush=True)
