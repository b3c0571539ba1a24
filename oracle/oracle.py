"""ctypes front-end of the CPU parity oracle (oracle/flac_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package `flacarray_b200` never does.

`build()` compiles oracle/liboracle.so (and oracle/_ref/libfa_utils.so from the reference's own
utils.c when /root/reference is present) with the Makefile in this directory.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "flac_oracle.c")
    stale = (not os.path.exists(so)) or os.path.getmtime(so) < os.path.getmtime(src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True, capture_output=True)
    ref_so = os.path.join(_HERE, "_ref", "libfa_utils.so")
    if os.path.isdir("/root/reference") and (force or not os.path.exists(ref_so)):
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    so = build()
    L = C.CDLL(so)
    L.orc_encode_bound.restype = C.c_int64
    L.orc_encode_bound.argtypes = [C.c_int64, C.c_int, C.c_int]
    L.orc_encode_stream.restype = C.c_int64
    L.orc_encode_stream.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, _u8p, C.c_int64]
    L.orc_decode_stream.restype = C.c_int
    L.orc_decode_stream.argtypes = [_u8p, C.c_int64, C.c_int64, C.c_int, C.c_int64, C.c_int64, C.c_void_p]
    L.orc_index_frames.restype = C.c_int64
    L.orc_index_frames.argtypes = [_u8p, C.c_int64, _i64p, _i32p, _i32p, C.c_int64]
    for name in ("orc_encode_i32", "orc_encode_i64"):
        f = getattr(L, name)
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_uint32, C.POINTER(C.c_int64), _i64p,
                      C.POINTER(C.c_void_p), C.c_int]
    for name in ("orc_decode_i32", "orc_decode_i64"):
        f = getattr(L, name)
        f.restype = C.c_int
        f.argtypes = [_u8p, _i64p, _i64p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int]
    L.orc_free.argtypes = [C.c_void_p]
    L.orc_float32_to_int32.restype = C.c_int
    L.orc_float32_to_int32.argtypes = [_f32p, C.c_int64, C.c_int64, C.c_void_p, _i32p, _f32p, _f32p]
    L.orc_float64_to_int64.restype = C.c_int
    L.orc_float64_to_int64.argtypes = [_f64p, C.c_int64, C.c_int64, C.c_void_p, _i64p, _f64p, _f64p]
    L.orc_int32_to_float32.argtypes = [_i32p, C.c_int64, C.c_int64, _f32p, _f32p, _f32p]
    L.orc_int64_to_float64.argtypes = [_i64p, C.c_int64, C.c_int64, _f64p, _f64p, _f64p]
    L.orc_crc8.restype = C.c_uint8
    L.orc_crc8.argtypes = [_u8p, C.c_int64]
    L.orc_crc16.restype = C.c_uint16
    L.orc_crc16.argtypes = [_u8p, C.c_int64]
    _LIB = L
    return L


def ref_utils():
    """The reference's own utils.c (oracle/_ref/libfa_utils.so), or None when not built."""
    global _REF
    if _REF is not None:
        return _REF
    so = os.path.join(_HERE, "_ref", "libfa_utils.so")
    if not os.path.exists(so):
        try:
            build()
        except Exception:
            pass
    if not os.path.exists(so):
        return None
    R = C.CDLL(so)
    R.float32_to_int32.restype = C.c_int
    R.float32_to_int32.argtypes = [_f32p, C.c_int64, C.c_int64, C.c_void_p, _i32p, _f32p, _f32p]
    R.float64_to_int64.restype = C.c_int
    R.float64_to_int64.argtypes = [_f64p, C.c_int64, C.c_int64, C.c_void_p, _i64p, _f64p, _f64p]
    R.int32_to_float32.argtypes = [_i32p, C.c_int64, C.c_int64, _f32p, _f32p, _f32p]
    R.int64_to_float64.argtypes = [_i64p, C.c_int64, C.c_int64, _f64p, _f64p, _f64p]
    _REF = R
    return R


# ---------------------------------------------------------------------------------------------
# numpy-level helpers
# ---------------------------------------------------------------------------------------------

def _flat2d(arr):
    arr = np.ascontiguousarray(arr)
    if arr.ndim == 1:
        return arr.reshape(1, -1)
    return arr.reshape(-1, arr.shape[-1])


def encode(arr, level=5, use_threads=False):
    """int32/int64 [..., stream_size] -> (compressed u8, starts i64[n], nbytes i64[n])."""
    L = lib()
    a = _flat2d(arr)
    n_stream, stream_size = a.shape
    starts = np.zeros(n_stream, np.int64)
    nb = C.c_int64(0)
    ptr = C.c_void_p()
    fn = L.orc_encode_i64 if a.dtype == np.int64 else L.orc_encode_i32
    if a.dtype not in (np.dtype(np.int32), np.dtype(np.int64)):
        raise ValueError("int32/int64 only")
    err = fn(a.ctypes.data, n_stream, stream_size, level, C.byref(nb), starts, C.byref(ptr), int(use_threads))
    if err != 0:
        raise RuntimeError(f"Encoding failed, return code = {err}")
    out = np.ctypeslib.as_array((C.c_uint8 * nb.value).from_address(ptr.value)).copy()
    L.orc_free(ptr)
    nbytes = np.empty(n_stream, np.int64)
    nbytes[:-1] = np.diff(starts)
    nbytes[-1] = nb.value - starts[-1]
    return out, starts, nbytes


def decode(compressed, starts, nbytes, stream_size, first=-1, last=-1, is_int64=False, use_threads=False):
    L = lib()
    starts = np.ascontiguousarray(starts, np.int64).reshape(-1)
    nbytes = np.ascontiguousarray(nbytes, np.int64).reshape(-1)
    n_stream = starts.size
    n_decode = stream_size
    if first >= 0 and last >= 0:
        n_decode = max(last - first, 0)
    out = np.zeros((n_stream, max(n_decode, 0)), np.int64 if is_int64 else np.int32)
    fn = L.orc_decode_i64 if is_int64 else L.orc_decode_i32
    err = fn(np.ascontiguousarray(compressed, np.uint8), starts, nbytes, n_stream, stream_size, first, last,
             out.ctypes.data, int(use_threads))
    if err != 0:
        raise RuntimeError(f"Decoding failed, return code = {err}")
    return out


def encode_stream(samples, level=5, stereo_mode=-1):
    """samples int32 [n] or [n, 2] -> bytes of one complete FLAC stream."""
    L = lib()
    s = np.ascontiguousarray(samples, np.int32)
    nch = 1 if s.ndim == 1 else s.shape[1]
    n = s.shape[0]
    cap = L.orc_encode_bound(n, nch, level)
    buf = np.zeros(cap, np.uint8)
    sz = L.orc_encode_stream(s.ctypes.data, n, nch, level, stereo_mode, buf, cap)
    if sz < 0:
        raise RuntimeError("oracle encode failed")
    return buf[:sz].copy()


def decode_stream(buf, stream_size, n_channels, first=0, n_decode=None):
    L = lib()
    if n_decode is None:
        n_decode = stream_size - first
    out = np.zeros((n_decode, n_channels), np.int32)
    b = np.ascontiguousarray(buf, np.uint8)
    err = L.orc_decode_stream(b, b.size, stream_size, n_channels, first, n_decode, out.ctypes.data)
    if err != 0:
        raise RuntimeError(f"Decoding failed, return code = {err}")
    return out


def index_frames(buf, cap=1 << 20):
    L = lib()
    b = np.ascontiguousarray(buf, np.uint8)
    offs = np.zeros(cap, np.int64)
    bss = np.zeros(cap, np.int32)
    cas = np.zeros(cap, np.int32)
    n = L.orc_index_frames(b, b.size, offs, bss, cas, cap)
    if n < 0:
        raise RuntimeError("oracle frame index failed")
    return offs[:n].copy(), bss[:n].copy(), cas[:n].copy()


def float_to_int(data, quanta=None, use_ref=False):
    """2-D float32/float64 [n_stream, stream_size], quanta None or array[n_stream]."""
    d = np.ascontiguousarray(data)
    n_stream, stream_size = d.shape
    L = ref_utils() if use_ref else lib()
    if L is None:
        raise RuntimeError("reference utils.c oracle not built")
    pre = "" if use_ref else "orc_"
    if d.dtype == np.float32:
        out = np.empty(d.shape, np.int32); off = np.empty(n_stream, np.float32); gain = np.empty(n_stream, np.float32)
        q = None if quanta is None else np.ascontiguousarray(quanta, np.float32)
        fn = getattr(L, pre + "float32_to_int32")
    else:
        out = np.empty(d.shape, np.int64); off = np.empty(n_stream, np.float64); gain = np.empty(n_stream, np.float64)
        q = None if quanta is None else np.ascontiguousarray(quanta, np.float64)
        fn = getattr(L, pre + "float64_to_int64")
    fn(d.reshape(-1), n_stream, stream_size, None if q is None else q.ctypes.data, out.reshape(-1), off, gain)
    return out, off, gain


def int_to_float(idata, offsets, gains, use_ref=False):
    d = np.ascontiguousarray(idata)
    n_stream, stream_size = d.shape
    L = ref_utils() if use_ref else lib()
    pre = "" if use_ref else "orc_"
    if d.dtype == np.int32:
        out = np.empty(d.shape, np.float32)
        getattr(L, pre + "int32_to_float32")(d.reshape(-1), n_stream, stream_size,
                                              np.ascontiguousarray(offsets, np.float32),
                                              np.ascontiguousarray(gains, np.float32), out.reshape(-1))
    else:
        out = np.empty(d.shape, np.float64)
        getattr(L, pre + "int64_to_float64")(d.reshape(-1), n_stream, stream_size,
                                              np.ascontiguousarray(offsets, np.float64),
                                              np.ascontiguousarray(gains, np.float64), out.reshape(-1))
    return out
