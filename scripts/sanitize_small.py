"""Small encode/decode workload for compute-sanitizer (memcheck / racecheck): all dtypes, full + short
frames, levels 0/5/8, constant / wasted / wide / noise streams, keep + slice decode."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import flacarray_b200 as fa

rng = np.random.default_rng(1)
L = 3 * 4096 + 700
cases = {
    "walk": np.cumsum(rng.integers(-900, 901, (3, L)), axis=1).astype(np.int32),
    "wide": np.cumsum(rng.integers(-2 ** 24, 2 ** 24, (2, L)), axis=1).astype(np.uint32).astype(np.int32),
    "const": np.full((1, L), 5, np.int32),
    "wasted": (np.cumsum(rng.integers(-900, 901, (1, L)), axis=1) << 7).astype(np.int32),
    "noise": rng.integers(-2 ** 31, 2 ** 31 - 1, (1, L), dtype=np.int64).astype(np.int32),
    "i64": (np.cumsum(rng.integers(-2 ** 20, 2 ** 20, (2, L)), axis=1) + 2 ** 40 * rng.integers(-4, 5, (2, L))).astype(np.int64),
    "f32": rng.normal(0, 1, (2, L)).astype(np.float32),
    "f64": rng.normal(0, 1, (2, L)),
}
for level in (0, 5, 8):
    for name, x in cases.items():
        kw = {}
        if x.dtype.kind == "f":
            kw = {"quanta": 1e-4} if x.dtype == np.float32 else {"precision": 6}
        far = fa.FlacArray.from_array(x, level=level, **kw)
        y = far.to_array()
        if x.dtype.kind == "i":
            assert np.array_equal(y, x), (name, level)
        keep = np.zeros(x.shape[0], bool); keep[-1] = True
        z = far.to_array(keep=keep, stream_slice=slice(5000, 5100))
        assert np.array_equal(z, y[keep][:, 5000:5100]), (name, level)
print("sanitize workload ok")
