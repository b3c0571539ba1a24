"""Drop-in replacement for the reference's compiled extension `flacarray.libflacarray`.

Same function names, argument meaning, return shapes/dtypes and error behaviour as
/root/reference/src/flacarray/libflacarray/libflacarray.pyx:113-823, but every array operation runs in
hand-written sm_100a CUDA kernels behind the C ABI of include/flacarray_b200.h (ctypes).  PyTorch is
used only to own device memory and CUDA streams.

Inputs may be numpy arrays (host; results come back as numpy, staged through pinned memory) or CUDA
torch tensors (device-resident; results stay on the device as torch tensors).  `use_threads` and the
`*_threaded` variants are accepted and ignored: the work is always parallel on the GPU.
There is no CPU path: without the built library or without a CUDA device these functions raise.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

flac_i32_dtype = np.dtype(np.int32)
flac_i64_dtype = np.dtype(np.int64)
compressed_dtype = np.dtype(np.uint8)
offset_dtype = np.dtype(np.int64)

_NP2TORCH = {
    np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64, np.dtype(np.float32): torch.float32,
    np.dtype(np.float64): torch.float64, np.dtype(np.uint8): torch.uint8, np.dtype(np.bool_): torch.bool,
}
_TORCH2NP = {v: k for k, v in _NP2TORCH.items()}
_FAB = {np.dtype(np.int32): 0, np.dtype(np.int64): 1, np.dtype(np.float32): 2, np.dtype(np.float64): 3}


def is_torch(x):
    return isinstance(x, torch.Tensor)


def np_dtype(x):
    """numpy dtype of a numpy array or torch tensor."""
    if is_torch(x):
        return _TORCH2NP[x.dtype]
    return np.dtype(x.dtype)


def _is_contiguous(x):
    if is_torch(x):
        return x.is_contiguous()
    return x.flags.c_contiguous


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("flacarray_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def to_device(x, device=None, dtype=None):
    """numpy / torch (any device) -> contiguous CUDA tensor (async H2D when the source is pinned)."""
    if device is None:
        device = x.device if (is_torch(x) and x.is_cuda) else _device()
    if is_torch(x):
        t = x
    else:
        a = np.ascontiguousarray(x)
        if not a.flags.writeable:
            a = a.copy()
        t = torch.from_numpy(a)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    if not t.is_cuda:
        t = t.to(device, non_blocking=True)
    return t.contiguous()


def to_host(t):
    """CUDA tensor -> numpy (through a pinned staging tensor that keeps the memory alive)."""
    if not t.is_cuda:
        return t.numpy()
    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    h.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return h.numpy()


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _check(rc, ctx, stream, what):
    """Collect the device-side error mask; raise like the reference binding (pyx:325-328, :649-652)."""
    if rc == 0:
        rc = _lib.lib().fab_finish(ctx.handle, stream)
    if rc != 0:
        if rc & (1 << 21):
            raise RuntimeError("Cannot convert data with NaNs to integers")
        msg = f"{what} failed, return code = {rc}"
        if rc & (1 << 20):
            msg += f" ({ctx.last_error()})"
        raise RuntimeError(msg)


# -------------------------------------------------------------------------------------------------
# float <-> int (pyx:113-282)
# -------------------------------------------------------------------------------------------------

def _float_to_int(flatdata, n_stream, stream_size, quanta, fdt, idt, check_nan=False):
    on_dev = is_torch(flatdata) and flatdata.is_cuda
    d = to_device(flatdata, dtype=_NP2TORCH[fdt])
    dev = d.device
    with torch.cuda.device(dev):
        ctx = _lib.context(dev)
        q = None
        if quanta is not None and len(quanta) == n_stream:  # pyx:142-144: wrong length => computed from the data
            q = to_device(quanta, dev, _NP2TORCH[fdt])
        out = torch.empty(n_stream * stream_size, dtype=_NP2TORCH[idt], device=dev)
        off = torch.empty(n_stream, dtype=_NP2TORCH[fdt], device=dev)
        gain = torch.empty(n_stream, dtype=_NP2TORCH[fdt], device=dev)
        st = _stream(dev)
        rc = _lib.lib().fab_float_to_int(ctx.handle, _ptr(d), _FAB[fdt], n_stream, stream_size, _ptr(q), _ptr(out),
                                         _ptr(off), _ptr(gain), st)
        if rc == 0:
            rc = _lib.lib().fab_finish(ctx.handle, st)
        if not check_nan:
            rc &= ~(1 << 21)  # the compiled converter does not look for NaNs (utils.py:268 does, before calling it)
        if rc != 0:
            _check(rc, ctx, st, "Encoding")
    if on_dev:
        return out, off, gain
    return to_host(out), to_host(off), to_host(gain)


def wrap_float32_to_int32(flatdata, n_stream, stream_size, quanta, check_nan=False):
    """pyx:113-161.  Returns (int32 flat, offsets float32[n], gains float32[n])."""
    return _float_to_int(flatdata, n_stream, stream_size, quanta, np.dtype(np.float32), np.dtype(np.int32), check_nan)


def wrap_float64_to_int64(flatdata, n_stream, stream_size, quanta, check_nan=False):
    """pyx:164-212."""
    return _float_to_int(flatdata, n_stream, stream_size, quanta, np.dtype(np.float64), np.dtype(np.int64), check_nan)


def _int_to_float(idata, n_stream, stream_size, offsets, gains, idt, fdt):
    on_dev = is_torch(idata) and idata.is_cuda
    d = to_device(idata, dtype=_NP2TORCH[idt])
    dev = d.device
    with torch.cuda.device(dev):
        ctx = _lib.context(dev)
        off = to_device(offsets, dev, _NP2TORCH[fdt])
        gain = to_device(gains, dev, _NP2TORCH[fdt])
        out = torch.empty(n_stream * stream_size, dtype=_NP2TORCH[fdt], device=dev)
        st = _stream(dev)
        rc = _lib.lib().fab_int_to_float(ctx.handle, _ptr(d), _FAB[idt], n_stream, stream_size, _ptr(off), _ptr(gain),
                                         _ptr(out), st)
        _check(rc, ctx, st, "Decoding")
    return out if on_dev else to_host(out)


def wrap_int32_to_float32(idata, n_stream, stream_size, offsets, gains):
    """pyx:215-247."""
    return _int_to_float(idata, n_stream, stream_size, offsets, gains, np.dtype(np.int32), np.dtype(np.float32))


def wrap_int64_to_float64(idata, n_stream, stream_size, offsets, gains):
    """pyx:250-282."""
    return _int_to_float(idata, n_stream, stream_size, offsets, gains, np.dtype(np.int64), np.dtype(np.float64))


# -------------------------------------------------------------------------------------------------
# encode (pyx:285-594)
# -------------------------------------------------------------------------------------------------

def encode_device(d, n_stream, stream_size, level, quanta=None):
    """Device-level encode.  d: CUDA tensor (int32/int64/float32/float64), flat or 2-D.

    Returns CUDA tensors (compressed u8[total], starts i64[n], nbytes i64[n], offsets, gains); offsets /
    gains are None for integer input.  For float input the quantisation (utils.c:160-328) is fused in
    front of the encoder; `quanta` is None (auto) or a CUDA tensor [n_stream].
    """
    dev = d.device
    dt = _TORCH2NP[d.dtype]
    with torch.cuda.device(dev):
        ctx = _lib.context(dev)
        L = _lib.lib()
        if level < 0 or level > 8:
            raise RuntimeError("Encoding failed, return code = 2")
        bound = L.fab_encode_bound(n_stream, stream_size, _FAB[dt], level)
        out = torch.empty(max(bound, 1), dtype=torch.uint8, device=dev)
        aux = torch.empty(2 * n_stream + 1, dtype=torch.int64, device=dev)
        starts, nbytes, total = aux[:n_stream], aux[n_stream:2 * n_stream], aux[2 * n_stream:]
        off = gain = None
        if dt.kind == "f":
            off = torch.empty(n_stream, dtype=d.dtype, device=dev)
            gain = torch.empty(n_stream, dtype=d.dtype, device=dev)
        st = _stream(dev)
        rc = L.fab_encode(ctx.handle, _ptr(d), _FAB[dt], n_stream, stream_size, level, _ptr(quanta), _ptr(off),
                          _ptr(gain), _ptr(out), bound, _ptr(starts), _ptr(nbytes), _ptr(total), st)
        _check(rc, ctx, st, "Encoding")
        n_total = int(total.item())
    return out[:n_total], starts, nbytes, off, gain


def _wrap_encode(flatdata, n_stream, stream_size, level, dt):
    on_dev = is_torch(flatdata) and flatdata.is_cuda
    d = to_device(flatdata, dtype=_NP2TORCH[dt])
    comp, starts, nbytes, _, _ = encode_device(d, int(n_stream), int(stream_size), int(level))
    if on_dev:
        return comp, starts, nbytes
    return to_host(comp), to_host(starts), to_host(nbytes)


def wrap_encode_i32(flatdata, n_stream, stream_size, level):
    """pyx:285-343.  Returns (compressed bytes, flat starts, flat nbytes)."""
    return _wrap_encode(flatdata, n_stream, stream_size, level, flac_i32_dtype)


def wrap_encode_i32_threaded(flatdata, n_stream, stream_size, level):
    """pyx:346-404 (the GPU path is always parallel)."""
    return _wrap_encode(flatdata, n_stream, stream_size, level, flac_i32_dtype)


def wrap_encode_i64(flatdata, n_stream, stream_size, level):
    """pyx:407-465."""
    return _wrap_encode(flatdata, n_stream, stream_size, level, flac_i64_dtype)


def wrap_encode_i64_threaded(flatdata, n_stream, stream_size, level):
    """pyx:468-526."""
    return _wrap_encode(flatdata, n_stream, stream_size, level, flac_i64_dtype)


def encode_flac(data, level, use_threads=False):
    """Compress an integer array to a FLAC representation (pyx:529-594).

    Returns (compressed bytestream, stream starting bytes, stream nbytes); starts / nbytes have the
    leading shape of `data` (or (1,) for a single stream).
    """
    dt = np_dtype(data)
    if dt != flac_i32_dtype and dt != flac_i64_dtype:
        raise RuntimeError("Only 32bit or 64bit integer data is supported")
    if not _is_contiguous(data):
        raise RuntimeError("Only C-contiguous arrays are supported")
    if level < 0 or level > 8:
        raise RuntimeError("FLAC only supports compression levels 0-8")
    shape = tuple(data.shape)
    stream_size = shape[-1]
    if len(shape[:-1]) == 0:
        n_stream = 1
        starts_shape = (1,)
    else:
        n_stream = int(np.prod(shape[:-1]))
        starts_shape = shape[:-1]
    flatdata = data.reshape((-1,))
    if dt == flac_i32_dtype:
        compressed, flatstarts, flatnbytes = wrap_encode_i32(flatdata, n_stream, stream_size, level)
    else:
        compressed, flatstarts, flatnbytes = wrap_encode_i64(flatdata, n_stream, stream_size, level)
    return (compressed, flatstarts.reshape(starts_shape), flatnbytes.reshape(starts_shape))


# -------------------------------------------------------------------------------------------------
# decode (pyx:597-823)
# -------------------------------------------------------------------------------------------------

def blocksize_hint(compressed, starts):
    """Nominal FLAC blocksize from the first stream's STREAMINFO (sizes the device frame table)."""
    try:
        s0 = int(np.asarray(starts.cpu() if is_torch(starts) else starts).reshape(-1)[0])
        head = compressed[s0:s0 + 12]
        head = head.cpu().numpy() if is_torch(head) else np.asarray(head)
        if head.size >= 12 and bytes(head[:4]) == b"fLaC":
            return (int(head[8]) << 8) | int(head[9])
    except Exception:
        pass
    return 0


def decode_device(comp, starts, nbytes, n_stream, stream_size, first_sample, last_sample, is_int64, max_nbytes,
                  bs_hint, offsets=None, gains=None):
    """Device-level decode.  All arrays are CUDA tensors; returns a flat CUDA tensor.

    With offsets/gains the int->float restore (utils.c:330-368) is applied on the device.
    """
    dev = comp.device
    n_decode = stream_size
    if first_sample >= 0 and last_sample >= 0:
        n_decode = last_sample - first_sample
    with torch.cuda.device(dev):
        ctx = _lib.context(dev)
        out = torch.empty(n_stream * max(n_decode, 0), dtype=torch.int64 if is_int64 else torch.int32, device=dev)
        st = _stream(dev)
        rc = _lib.lib().fab_decode(ctx.handle, _ptr(comp), _ptr(starts), _ptr(nbytes), n_stream, stream_size,
                                   1 if is_int64 else 0, first_sample, last_sample, _ptr(out), _ptr(offsets),
                                   _ptr(gains), int(max_nbytes), int(bs_hint), st)
        _check(rc, ctx, st, "Decoding")
    if offsets is not None and gains is not None:
        out = out.view(torch.float64 if is_int64 else torch.float32)
    return out


def _wrap_decode(compressed, starts, nbytes, n_stream, stream_size, first_sample, last_sample, is_int64):
    on_dev = is_torch(compressed) and compressed.is_cuda
    h_starts = starts.cpu().numpy() if is_torch(starts) else np.asarray(starts)
    h_nbytes = nbytes.cpu().numpy() if is_torch(nbytes) else np.asarray(nbytes)
    max_nb = int(h_nbytes.max()) if h_nbytes.size else 0
    hint = blocksize_hint(compressed, h_starts)
    if on_dev:
        comp = compressed
        d_starts = to_device(h_starts, comp.device, torch.int64)
    else:
        # upload only the byte range the selected streams cover (keep masks select sparse subsets)
        lo = int(h_starts.min())
        hi = int((h_starts + h_nbytes).max())
        comp = to_device(compressed[lo:hi])
        d_starts = to_device(h_starts - lo, comp.device, torch.int64)
    d_nbytes = to_device(h_nbytes, comp.device, torch.int64)
    out = decode_device(comp, d_starts, d_nbytes, int(n_stream), int(stream_size), int(first_sample),
                        int(last_sample), is_int64, max_nb, hint)
    return out if on_dev else to_host(out)


def wrap_decode_i32(compressed, starts, nbytes, n_stream, stream_size, first_sample, last_sample, use_threads):
    """pyx:597-653.  Returns the flat int32 array."""
    return _wrap_decode(compressed, starts, nbytes, n_stream, stream_size, first_sample, last_sample, False)


def wrap_decode_i64(compressed, starts, nbytes, n_stream, stream_size, first_sample, last_sample, use_threads):
    """pyx:655-711."""
    return _wrap_decode(compressed, starts, nbytes, n_stream, stream_size, first_sample, last_sample, True)


def decode_flac(compressed, starts, nbytes, stream_size, first_sample=-1, last_sample=-1, use_threads=False,
                is_int64=False):
    """Decompress a FLAC compressed bytestream (pyx:713-823)."""
    if np_dtype(compressed) != compressed_dtype:
        raise RuntimeError("Compressed data should be of type uint8")
    if not _is_contiguous(compressed):
        raise RuntimeError("Only C-contiguous arrays are supported")
    if np_dtype(starts) != offset_dtype:
        raise RuntimeError("starts data should be of type int64")
    if not _is_contiguous(starts):
        raise RuntimeError("Only C-contiguous arrays are supported")
    if np_dtype(nbytes) != offset_dtype:
        raise RuntimeError("nbytes data should be of type int64")
    if not _is_contiguous(nbytes):
        raise RuntimeError("Only C-contiguous arrays are supported")
    if stream_size <= 0:
        raise RuntimeError("You must specify the non-zero output stream size")
    if len(compressed.shape) != 1:
        raise RuntimeError("Compressed byte array should be one dimensional")

    n_decode = stream_size
    if first_sample >= 0 and last_sample >= 0:
        if last_sample > stream_size:
            raise RuntimeError("last_sample is beyond end of stream")
        if first_sample > stream_size - 1:
            raise RuntimeError("first_sample is beyond last element of stream")
        if first_sample >= last_sample:
            raise RuntimeError("first_sample is larger than last_sample")
        n_decode = last_sample - first_sample

    output_shape = tuple(starts.shape) + (n_decode,)
    n_stream = int(np.prod(starts.shape))
    flat_starts = starts.reshape((-1,))
    flat_nbytes = nbytes.reshape((-1,))
    if is_int64:
        flat_output = wrap_decode_i64(compressed, flat_starts, flat_nbytes, n_stream, stream_size, first_sample,
                                      last_sample, use_threads)
    else:
        flat_output = wrap_decode_i32(compressed, flat_starts, flat_nbytes, n_stream, stream_size, first_sample,
                                      last_sample, use_threads)
    return flat_output.reshape(output_shape)


# -------------------------------------------------------------------------------------------------
# Fused entry points used by compress.py / decompress.py (no integer intermediate in host memory)
# -------------------------------------------------------------------------------------------------

def encode_flac_float(data, level, quanta):
    """float32/float64 [..., stream_size] -> (compressed, starts, nbytes, offsets, gains).

    Equivalent to utils.float_to_int (utils.c:160-328) followed by encode_flac, with the quantisation
    fused into the encoder's frame load.  `quanta`: None (derive from the data range) or an array with
    one value per stream.
    """
    dt = np_dtype(data)
    shape = tuple(data.shape)
    stream_size = shape[-1]
    if len(shape[:-1]) == 0:
        n_stream, lead = 1, (1,)
    else:
        n_stream, lead = int(np.prod(shape[:-1])), shape[:-1]
    on_dev = is_torch(data) and data.is_cuda
    d = to_device(data.reshape((-1,)), dtype=_NP2TORCH[dt])
    q = None
    if quanta is not None:
        q = to_device(quanta.reshape((-1,)), d.device, _NP2TORCH[dt])
        if q.numel() != n_stream:
            q = None
    comp, starts, nbytes, off, gain = encode_device(d, n_stream, stream_size, int(level), q)
    if not on_dev:
        comp, starts, nbytes, off, gain = (to_host(comp), to_host(starts), to_host(nbytes), to_host(off), to_host(gain))
    return comp, starts.reshape(lead), nbytes.reshape(lead), off, gain


def decode_flac_float(compressed, starts, nbytes, stream_size, offsets, gains, first_sample=-1, last_sample=-1,
                      is_int64=False):
    """decode_flac followed by utils.int_to_float (utils.c:330-368) without leaving the device."""
    n_decode = stream_size
    if first_sample >= 0 and last_sample >= 0:
        n_decode = last_sample - first_sample
    output_shape = tuple(starts.shape) + (n_decode,)
    n_stream = int(np.prod(starts.shape))
    on_dev = is_torch(compressed) and compressed.is_cuda
    h_starts = (starts.cpu().numpy() if is_torch(starts) else np.asarray(starts)).reshape(-1)
    h_nbytes = (nbytes.cpu().numpy() if is_torch(nbytes) else np.asarray(nbytes)).reshape(-1)
    max_nb = int(h_nbytes.max()) if h_nbytes.size else 0
    hint = blocksize_hint(compressed, h_starts)
    if on_dev:
        comp = compressed
        d_starts = to_device(h_starts, comp.device, torch.int64)
    else:
        lo = int(h_starts.min())
        hi = int((h_starts + h_nbytes).max())
        comp = to_device(compressed[lo:hi])
        d_starts = to_device(h_starts - lo, comp.device, torch.int64)
    d_nbytes = to_device(h_nbytes, comp.device, torch.int64)
    fdt = torch.float64 if is_int64 else torch.float32
    off = to_device(offsets.reshape((-1,)), comp.device, fdt)
    gain = to_device(gains.reshape((-1,)), comp.device, fdt)
    out = decode_device(comp, d_starts, d_nbytes, n_stream, int(stream_size), int(first_sample), int(last_sample),
                        is_int64, max_nb, hint, off, gain)
    out = out if on_dev else to_host(out)
    return out.reshape(output_shape)
