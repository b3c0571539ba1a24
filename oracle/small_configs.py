"""Small, seeded versions of the five BASELINE.json configs (test infrastructure: tests/ and scripts/ only)."""
import numpy as np


def small_configs():
    """Small versions of the five BASELINE.json configs, seeded like the reference's demo (demo.py:12)."""
    rng = np.random.default_rng(123456789)
    n = 20000
    t = np.arange(n) / n
    tod = (5 * (rng.random((3, 1)) - 0.5) + 2 * np.sin(2 * np.pi * 15 * t) + 6 * np.sin(2 * np.pi * 5 * t) + rng.normal(0, 1, (3, n)))
    walk64 = np.cumsum(rng.integers(-2 ** 20, 2 ** 20 + 1, (2, n)), axis=1) + 2 ** 40 * rng.integers(-4, 5, (2, n))
    walk64[0, :4] = [-2 ** 63, 2 ** 63 - 1, 2 ** 32, -2 ** 32]
    return {
        "cfg1_i32_walk": (np.cumsum(rng.integers(-1000, 1001, (4, n)), axis=1) + rng.integers(-50, 51, (4, n))).astype(np.int32),
        "cfg2_tod_q1e-4": np.round(tod / 1e-4).astype(np.int32),
        "cfg3_i64_walk": walk64.astype(np.int64),
        "cfg4_tod_f64": np.round(tod / 1e-5).astype(np.int64),
        "full_range": rng.integers(-2 ** 31, 2 ** 31, (2, n), dtype=np.int64).astype(np.int32),
    }
