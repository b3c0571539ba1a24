"""Synthetic data with the model of /root/reference/src/flacarray/demo.py:11-114 (seed 123456789)."""
import numpy as np

from .mpi import global_array_properties


def create_fake_data(local_shape, sigma=1.0, dtype=np.float64, seed=123456789, comm=None, dc_sigma=5):
    """Fake random data for testing: uniform full-range values when sigma is None (with the dtype's
    extremes planted), else DC level + two sinusoids + Gaussian noise per stream.

    Returns (local data, distribution of the leading axis).
    """
    rank = 0 if comm is None else comm.rank
    gprops = global_array_properties(local_shape, comm)
    shape = gprops["shape"]
    mpi_dist = gprops["dist"]
    flatshape = int(np.prod(shape))
    stream_size = shape[-1]
    leading_shape = shape[:-1]
    leading_shape_ext = leading_shape + (1,)
    dtype = np.dtype(dtype)

    rng = np.random.default_rng(seed=seed)
    global_data = None
    if rank == 0:
        if sigma is None:
            if dtype.kind == "i":
                low, high = np.iinfo(dtype).min, np.iinfo(dtype).max
                flat_data = rng.integers(low=low, high=high, size=flatshape, dtype=np.int64).astype(dtype)
            else:
                low, high = np.finfo(dtype).min, np.finfo(dtype).max
                flat_data = rng.uniform(low=low, high=high, size=flatshape).astype(dtype)
            flat_data[0] = low
            flat_data[1] = high
            global_data = flat_data.reshape(shape)
        else:
            dc = 0 if dc_sigma is None else dc_sigma * sigma * (rng.random(size=leading_shape_ext) - 0.5)
            wave = np.zeros(stream_size, dtype=dtype)
            t = np.arange(stream_size)
            minf = 5 / stream_size
            for freq, amp in zip([3 * minf, minf], [2 * sigma, 6 * sigma]):
                wave[:] += amp * np.sin(2 * np.pi * freq * t)
            scale = rng.random(size=leading_shape_ext)
            global_data = np.empty(shape, dtype=dtype)
            global_data[...] = dc
            global_data[...] += scale * wave
            global_data[:] += rng.normal(0.0, sigma, flatshape).reshape(shape)
    if comm is not None:
        global_data = comm.bcast(global_data, root=0)

    if len(leading_shape) == 0 or (len(leading_shape) == 1 and leading_shape[0] == 1):
        data = global_data
    else:
        local_slice = (slice(mpi_dist[rank][0], mpi_dist[rank][1], 1),) + tuple(slice(None) for _ in shape[1:])
        data = global_data[local_slice]
    if len(data.shape) == 2 and data.shape[0] == 1:
        data = data.reshape((-1))
    return data, mpi_dist
