"""Host-side helpers mirroring /root/reference/src/flacarray/utils.py (same names, arguments,
errors): logging, function timers, float<->int conversion wrappers, keep-mask selection.

The array work is done by the CUDA library through `libflacarray`; numpy inputs give numpy results,
CUDA torch tensors give CUDA torch tensors.
"""
import logging
import os
import time
from functools import wraps

import numpy as np

from .libflacarray import (
    is_torch,
    np_dtype,
    wrap_float32_to_int32,
    wrap_float64_to_int64,
    wrap_int32_to_float32,
    wrap_int64_to_float64,
)

# ---- logging / timers (reference utils.py:21-175) -------------------------------------------------
log = logging.getLogger("flacarray")
_lvl = os.environ.get("FLACARRAY_LOGLEVEL", os.environ.get("FLACARRAY_LOG_LEVEL", None))
if _lvl is not None and hasattr(logging, _lvl.upper()):
    log.setLevel(getattr(logging, _lvl.upper()))

_use_timing = os.environ.get("FLACARRAY_TIMING", "").lower() in ("1", "true", "yes")
_timers = dict()
_stack = list()


def use_function_timers():
    return _use_timing


def function_timer(f):
    """Accumulate wall time per call-stack-qualified function name when FLACARRAY_TIMING is set."""
    if not _use_timing:
        return f

    @wraps(f)
    def wrapper(*args, **kwargs):
        _stack.append(f.__qualname__)
        name = ":".join(_stack)
        t0 = time.perf_counter()
        try:
            return f(*args, **kwargs)
        finally:
            dt = time.perf_counter() - t0
            ent = _timers.setdefault(name, [0, 0.0])
            ent[0] += 1
            ent[1] += dt
            _stack.pop()

    return wrapper


def clear_timers():
    _timers.clear()


def get_timers():
    return {k: {"calls": v[0], "seconds": v[1]} for k, v in _timers.items()}


def print_timers():
    for name in sorted(_timers):
        calls, secs = _timers[name]
        print(f"{name}: {secs:0.3e} s in {calls} calls", flush=True)


# ---- small helpers (reference utils.py:178-243) ------------------------------------------------------
def ensure_one_element(input, dtype=None):
    """Promote a scalar to a 1-element array, or check that an array is a single element of dtype."""
    if isinstance(input, np.ndarray):
        if input.shape != (1,):
            raise ValueError("Input array does not have a single element.")
        if dtype is not None and input.dtype != np.dtype(dtype):
            raise ValueError(f"Input has dtype {input.dtype}, not {dtype}")
        return input
    if is_torch(input):
        if tuple(input.shape) != (1,):
            raise ValueError("Input array does not have a single element.")
        return input
    if dtype is None:
        raise ValueError("Input is a scalar, dtype must be specified")
    return np.array([input], dtype=dtype)


def compressed_dtype(n_channel, offsets, gains):
    """dtype of the decompressed data from channel count and presence of offsets/gains."""
    if n_channel == 2:
        return np.dtype(np.int64) if (offsets is None or gains is None) else np.dtype(np.float64)
    return np.dtype(np.int32) if (offsets is None or gains is None) else np.dtype(np.float32)


def _std_last_axis(data):
    """np.std(data, axis=-1, keepdims=True) (reference utils.py:284), reduced on the device
    (libflacarray.stream_std -> fab_stream_std): one read of the data, double-precision moments.  numpy sums
    float32 input pairwise in single precision, so the two differ by rounding only."""
    from .libflacarray import stream_std

    return np.asarray(stream_std(data))[..., None]


def quanta_from_precision(data, precision, leading_shape):
    """reference utils.py:282-296."""
    rms = _std_last_axis(data)
    try:
        len(precision)
        if precision.shape != leading_shape:
            msg = f"precision array ({precision}) has shape that does not "
            msg += f"match leading shape of data ({precision.shape} != "
            msg += f"{leading_shape})"
            raise RuntimeError(msg)
        return rms.reshape(leading_shape) / 10 ** precision.reshape(leading_shape)
    except TypeError:
        return rms.reshape(leading_shape) / 10**precision


@function_timer
def float_to_int(data, quanta=None, precision=None):
    """Convert floating point data to integers (reference utils.py:246-342, utils.c:160-328).

    Returns (integer data, offset array, gain array).
    """
    dt = np_dtype(data)
    if quanta is not None and precision is not None:
        raise RuntimeError("Cannot specify both quanta and precision")
    if dt != np.dtype(np.float32) and dt != np.dtype(np.float64):
        raise ValueError("Only float32 and float64 data are supported")

    shape = tuple(data.shape)
    leading_shape = shape[:-1]
    n_stream = 1 if len(leading_shape) == 0 else int(np.prod(leading_shape))
    stream_size = shape[-1]

    if precision is not None:
        quanta = quanta_from_precision(data, precision, leading_shape)

    if quanta is None:
        quanta = np.zeros(0, dtype=dt)  # "fake value": wrong length => derived from the data range
    else:
        try:
            len(quanta)
            if tuple(quanta.shape) != leading_shape:
                msg = f"quanta array ({quanta}) has shape that does not "
                msg += f"match leading shape of data ({quanta.shape} != "
                msg += f"{leading_shape})"
                raise RuntimeError(msg)
        except TypeError:
            quanta = quanta * np.ones(leading_shape, dtype=dt)
    if not is_torch(quanta):
        quanta = np.asarray(quanta).reshape((-1,)).astype(dt)
    else:
        quanta = quanta.reshape((-1,))

    wrap = wrap_float32_to_int32 if dt == np.dtype(np.float32) else wrap_float64_to_int64
    output, offsets, gains = wrap(data.reshape((-1,)), n_stream, stream_size, quanta, check_nan=True)

    if len(leading_shape) == 0:
        return (output.reshape(shape), offsets.reshape((-1,)), gains.reshape((-1,)))
    return (output.reshape(shape), offsets.reshape(leading_shape), gains.reshape(leading_shape))


@function_timer
def int_to_float(idata, offset, gain):
    """Restore floating point data from integers (reference utils.py:346-408, utils.c:330-368)."""
    dt = np_dtype(idata)
    if dt != np.dtype(np.int32) and dt != np.dtype(np.int64):
        raise ValueError("Input data should be int32 or int64")
    is_int64 = dt == np.dtype(np.int64)
    fdt = np.float64 if is_int64 else np.float32

    shape = tuple(idata.shape)
    leading_shape = shape[:-1]
    if len(leading_shape) == 0 or (len(leading_shape) == 1 and leading_shape[0] == 1):
        n_stream = 1
        offset = ensure_one_element(offset, fdt)
        gain = ensure_one_element(gain, fdt)
    else:
        n_stream = int(np.prod(leading_shape))
        if tuple(offset.shape) != leading_shape:
            msg = f"Offset array has shape {offset.shape}, expected "
            msg += f"shape {leading_shape}"
            raise ValueError(msg)
        if tuple(gain.shape) != leading_shape:
            msg = f"Gain array has shape {gain.shape}, expected "
            msg += f"shape {leading_shape}"
            raise ValueError(msg)
    stream_size = shape[-1]
    wrap = wrap_int64_to_float64 if is_int64 else wrap_int32_to_float32
    result = wrap(idata.reshape((-1,)), n_stream, stream_size, offset.reshape((-1,)), gain.reshape((-1,)))
    return result.reshape(shape)


def keep_select(keep, stream_starts, stream_nbytes):
    """Filter a subset of streams (reference utils.py:411-449).

    Returns (starts, nbytes, indices): compacted int64 arrays in C order of the True entries of `keep`
    and the list of index tuples (None when keep is None).  Vectorised with np.nonzero instead of the
    reference's nditer loop; same results.
    """
    if keep is None:
        return (stream_starts, stream_nbytes, None)
    if keep.shape != stream_starts.shape:
        raise RuntimeError("The keep array should have the same shape as stream_starts")
    if keep.shape != stream_nbytes.shape:
        raise RuntimeError("The keep array should have the same shape as stream_starts")
    nz = np.nonzero(keep)
    starts = np.ascontiguousarray(np.asarray(stream_starts)[nz], dtype=np.int64)
    nbytes = np.ascontiguousarray(np.asarray(stream_nbytes)[nz], dtype=np.int64)
    indices = [tuple(int(x) for x in idx) for idx in zip(*nz)]
    return (starts, nbytes, indices)


def select_keep_indices(arr, indices):
    """Extract array elements with a list of index tuples (reference utils.py:452-459)."""
    if arr is None:
        return None
    if indices is None:
        return arr
    if len(indices) == 0:
        return np.zeros(0, dtype=arr.dtype)
    idx = tuple(np.array([i[d] for i in indices]) for d in range(len(indices[0])))
    return np.ascontiguousarray(np.asarray(arr)[idx], dtype=arr.dtype)
