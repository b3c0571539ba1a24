"""Synthetic test data in the spirit of the reference's demo module (src/flacarray/demo.py): either values that
span the whole range of the dtype, or detector-like timestreams.  Used by the benchmark CLI and the tests; the
numbers are this package's own (a different draw order than the reference's generator, same statistics)."""
import numpy as np

from .mpi import global_array_properties

_SEED = 123456789      # the reference's default seed (demo.py:12)


def _full_range(rng, shape, dtype):
    """Uniform over everything `dtype` can hold, with both extremes guaranteed to occur (the lossless-int and
    clipping edge cases of the reference's tests, tests/bindings.py:165-230)."""
    n = int(np.prod(shape))
    if dtype.kind == "i":
        info = np.iinfo(dtype)
        flat = rng.integers(info.min, info.max, size=n, dtype=np.int64, endpoint=True).astype(dtype)
    else:
        info = np.finfo(dtype)
        # (uniform(min, max) overflows: draw a sign and a magnitude)
        flat = (rng.choice(np.array([-1.0, 1.0]), size=n) * rng.uniform(0.0, float(info.max), size=n)).astype(dtype)
    flat[:2] = (info.min, info.max)
    return flat.reshape(shape)


def _timestreams(rng, shape, dtype, sigma, dc_sigma):
    """Per stream: a DC level, a slow pair of sinusoids (5 and 15 periods per stream) scaled by a random factor,
    and white noise of standard deviation `sigma`."""
    n = shape[-1]
    lead = shape[:-1] + (1,)
    phase = 2.0 * np.pi * np.arange(n) / n
    drift = 6.0 * sigma * np.sin(5.0 * phase) + 2.0 * sigma * np.sin(15.0 * phase)
    out = rng.normal(0.0, sigma, size=shape)
    out += rng.random(size=lead) * drift
    if dc_sigma is not None:
        out += dc_sigma * sigma * (rng.random(size=lead) - 0.5)
    return out.astype(dtype)


def create_fake_data(local_shape, sigma=1.0, dtype=np.float64, seed=_SEED, comm=None, dc_sigma=5):
    """Random data for tests and benchmarks.

    `sigma=None` gives full-range values of `dtype`; otherwise detector-like timestreams with noise `sigma`.  With a
    communicator the global array (leading axis split over the ranks) is generated on rank 0 and every rank gets its
    block.  Returns (local data, leading-axis distribution [(first, last), ...]); a single stream comes back 1-D.
    """
    dtype = np.dtype(dtype)
    props = global_array_properties(tuple(local_shape), comm)
    shape, dist = props["shape"], props["dist"]
    rank = 0 if comm is None else comm.rank
    full = None
    if rank == 0:
        rng = np.random.default_rng(seed)
        full = _full_range(rng, shape, dtype) if sigma is None else _timestreams(rng, shape, dtype, sigma, dc_sigma)
    if comm is not None:
        full = comm.bcast(full, root=0)
    if len(shape) > 1 and shape[0] > 1:
        lo, hi = dist[rank]
        full = full[lo:hi]
    if full.ndim == 2 and full.shape[0] == 1:
        full = full.reshape(-1)
    return full, dist
