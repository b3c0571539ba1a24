"""array_decompress / array_decompress_slice: same contract as
/root/reference/src/flacarray/decompress.py:18-205 (keep mask, sample window, single-stream
flattening rules, error messages).  The FLAC decode and the int->float restore run on the device."""
import numpy as np

from .libflacarray import decode_flac, decode_flac_float, is_torch
from .utils import ensure_one_element, function_timer, keep_select, select_keep_indices


@function_timer
def array_decompress_slice(
    compressed,
    stream_size,
    stream_starts,
    stream_nbytes,
    stream_offsets=None,
    stream_gains=None,
    keep=None,
    first_stream_sample=None,
    last_stream_sample=None,
    is_int64=False,
    use_threads=False,
    no_flatten=False,
):
    """Decompress a slice of a FLAC encoded array and restore the original data type.

    Returns (output array, list of stream indices); the list is None when `keep` is None.
    """
    if first_stream_sample is None:
        first_stream_sample = -1
    if last_stream_sample is None:
        last_stream_sample = -1

    is_scalar = False
    if not (isinstance(stream_starts, np.ndarray) or is_torch(stream_starts)) or (
        len(stream_starts.shape) == 1 and stream_starts.shape[0] == 1
    ):
        is_scalar = True
        stream_starts = ensure_one_element(stream_starts, np.int64)
        stream_nbytes = ensure_one_element(stream_nbytes, np.int64)
        if stream_offsets is not None:
            fdt = np.float64 if is_int64 else np.float32
            stream_offsets = ensure_one_element(stream_offsets, fdt)
            stream_gains = ensure_one_element(stream_gains, fdt)
    if is_torch(stream_starts):
        stream_starts = stream_starts.cpu().numpy()
    if is_torch(stream_nbytes):
        stream_nbytes = stream_nbytes.cpu().numpy()

    starts, nbytes, indices = keep_select(keep, stream_starts, stream_nbytes)
    if stream_offsets is not None and is_torch(stream_offsets):
        stream_offsets = stream_offsets.cpu().numpy()
    if stream_gains is not None and is_torch(stream_gains):
        stream_gains = stream_gains.cpu().numpy()
    offsets = select_keep_indices(stream_offsets, indices)
    gains = select_keep_indices(stream_gains, indices)

    if stream_offsets is not None:
        if stream_gains is not None:
            arr = decode_flac_float(
                compressed,
                starts,
                nbytes,
                stream_size,
                offsets,
                gains,
                first_sample=first_stream_sample,
                last_sample=last_stream_sample,
                is_int64=is_int64,
            ) if starts.size > 0 else _empty(starts, stream_size, first_stream_sample, last_stream_sample, is_int64, True)
        else:
            raise RuntimeError("When specifying offsets, you must also provide the gains")
    else:
        if stream_gains is not None:
            raise RuntimeError("When specifying gains, you must also provide the offsets")
        arr = decode_flac(
            compressed,
            starts,
            nbytes,
            stream_size,
            first_sample=first_stream_sample,
            last_sample=last_stream_sample,
            use_threads=use_threads,
            is_int64=is_int64,
        ) if starts.size > 0 else _empty(starts, stream_size, first_stream_sample, last_stream_sample, is_int64, False)
    if is_scalar and not no_flatten:
        return (arr.reshape((-1,)), indices)
    return (arr, indices)


def _empty(starts, stream_size, first, last, is_int64, is_float):
    """Zero selected streams (all-False keep mask): the reference returns an empty (0, n) array."""
    n = stream_size if not (first >= 0 and last >= 0) else last - first
    if is_float:
        dt = np.float64 if is_int64 else np.float32
    else:
        dt = np.int64 if is_int64 else np.int32
    return np.zeros(tuple(starts.shape) + (n,), dtype=dt)


@function_timer
def array_decompress(
    compressed,
    stream_size,
    stream_starts,
    stream_nbytes,
    stream_offsets=None,
    stream_gains=None,
    first_stream_sample=None,
    last_stream_sample=None,
    is_int64=False,
    use_threads=False,
    no_flatten=False,
):
    """Decompress a FLAC encoded array and restore the original data type."""
    arr, _ = array_decompress_slice(
        compressed,
        stream_size,
        stream_starts,
        stream_nbytes,
        stream_offsets=stream_offsets,
        stream_gains=stream_gains,
        keep=None,
        first_stream_sample=first_stream_sample,
        last_stream_sample=last_stream_sample,
        is_int64=is_int64,
        use_threads=use_threads,
        no_flatten=no_flatten,
    )
    return arr
