"""array_compress: same contract as /root/reference/src/flacarray/compress.py:12-84.

Float input is quantised (utils.c:160-328) and encoded in one fused device pass; the returned
offsets / gains have the leading shape of the input exactly as the reference's float_to_int returns
them.  Reference quirk Q1 (array-valued `quanta` raising AttributeError, compress.py:60-70) is fixed as
a superset: an array of the leading shape is accepted.
"""
import numpy as np

from .libflacarray import encode_flac, encode_flac_float, is_torch, np_dtype
from .utils import function_timer


@function_timer
def array_compress(arr, level=5, quanta=None, precision=None, use_threads=False):
    """Compress an array with optional floating point conversion.

    Returns (compressed bytes, stream starts, stream_nbytes, stream offsets, stream gains); offsets and
    gains are None for integer input.
    """
    size = arr.numel() if is_torch(arr) else arr.size
    if size == 0:
        raise ValueError("Cannot compress a zero-sized array!")
    shape = tuple(arr.shape)
    leading_shape = shape[:-1]
    dt = np_dtype(arr)

    if dt == np.dtype(np.float32) or dt == np.dtype(np.float64):
        if quanta is None and precision is None:
            msg = f"Compressing floating point data ('{dt}') "
            msg += "requires specifying either quanta or precision."
            raise RuntimeError(msg)
        if quanta is not None:
            if precision is not None:
                raise RuntimeError("Cannot set both quanta and precision")
            try:
                len(quanta)
                if tuple(quanta.shape) != leading_shape:
                    msg = "If not a scalar, quanta must have the same shape as the "
                    msg += "leading dimensions of the array"
                    raise ValueError(msg)
                dquanta = quanta if is_torch(quanta) else np.asarray(quanta).astype(dt)
            except TypeError:
                dquanta = quanta * np.ones(leading_shape, dtype=dt)
        else:
            # utils.py:282-296: the standard deviations are reduced on the device, inside the encode call (a host
            # array crosses the bus once: every chunk of streams gets its quanta right before it is encoded)
            dquanta = None
            try:
                len(precision)
                if tuple(precision.shape) != leading_shape:
                    msg = f"precision array ({precision}) has shape that does not "
                    msg += f"match leading shape of data ({precision.shape} != "
                    msg += f"{leading_shape})"
                    raise RuntimeError(msg)
            except TypeError:
                pass
        if level < 0 or level > 8:
            raise RuntimeError("FLAC only supports compression levels 0-8")
        compressed, starts, nbytes, foff, gains = encode_flac_float(arr, level, dquanta,
                                                                    precision=None if quanta is not None else precision)
        if len(leading_shape) == 0:
            foff, gains = foff.reshape((-1,)), gains.reshape((-1,))
        else:
            foff, gains = foff.reshape(leading_shape), gains.reshape(leading_shape)
        return (compressed, starts, nbytes, foff, gains)
    elif dt == np.dtype(np.int32) or dt == np.dtype(np.int64):
        (compressed, starts, nbytes) = encode_flac(arr, level, use_threads=use_threads)
        return (compressed, starts, nbytes, None, None)
    else:
        raise ValueError(f"Unsupported data type '{dt}'")
