"""Logic tests of the CUDA kernel bodies through the OS-thread SIMT emulator (tests/hostsim).

These run the SAME source the GPU runs (flacarray_b200/csrc/*.h) on the CPU, purely as a unit-test
harness for the GPU-less container; the oracle is the checker.  The emulator is not part of the product.
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))


@pytest.fixture(scope="module")
def H():
    import hostsim

    hostsim.lib()
    return hostsim


def _cases(rng):
    walk = (np.cumsum(rng.integers(-1000, 1001, (2, 9000)), axis=1) + rng.integers(-50, 51, (2, 9000))).astype(np.int32)
    full = rng.integers(-2 ** 31, 2 ** 31, (1, 5000), dtype=np.int64).astype(np.int32)
    full[0, 0], full[0, 1] = -2 ** 31, 2 ** 31 - 1
    return {
        "walk": walk, "full": full, "const": np.full((1, 5000), -77, np.int32),
        "wasted": (walk[:1] << 5).astype(np.int32), "tiny": rng.integers(-5, 6, (2, 7)).astype(np.int32),
        "one": np.array([[42]], np.int32), "odd": rng.integers(-100, 100, (1, 1000)).astype(np.int32),
        # wide samples (64-bit statistics path): a 2^29-range walk, and one that wraps like the low word of an int64
        "widewalk": np.cumsum(rng.integers(-2 ** 22, 2 ** 22, (2, 9000)), axis=1).astype(np.int32),
        "widewasted": (np.cumsum(rng.integers(-2 ** 15, 2 ** 15, (1, 9000)), axis=1) << 9).astype(np.int32),
        "wrap": np.cumsum(rng.integers(-2 ** 26, 2 ** 26, (1, 9000)), axis=1).astype(np.uint32).astype(np.int32),
    }


@pytest.mark.parametrize("level", [0, 5, 8])
def test_encoder_body_roundtrips_through_oracle(H, oracle, level):
    rng = np.random.default_rng(21)
    for name, x in _cases(rng).items():
        c, s, n, _, _ = H.encode(x, level)
        assert s[0] == 0 and np.array_equal(np.cumsum(n) - n, s) and n.sum() == c.size
        assert np.array_equal(oracle.decode(c, s, n, x.shape[1]), x), name
        oc, _, _ = oracle.encode(x, level)
        if x.size > 4000:
            assert c.size <= 1.02 * oc.size, (name, c.size, oc.size)


def test_decoder_bodies_on_own_and_foreign_streams(H, oracle):
    rng = np.random.default_rng(22)
    for name, x in _cases(rng).items():
        c, s, n, _, _ = H.encode(x, 5)          # carries the frame-size table
        oc, os_, on = oracle.encode(x, 5)       # foreign: needs the sync scan
        for comp, st, nb in ((c, s, n), (oc, os_, on)):
            y, walked = H.decode(comp, st, nb, x.shape[1])
            assert np.array_equal(y, x) and walked == 0, name
            y, walked = H.decode(comp, st, nb, x.shape[1], mode=2)  # warp-tile throughput path
            assert np.array_equal(y, x) and walked == 0, name
            y, walked = H.decode(comp, st, nb, x.shape[1], mode=1)  # sequential walker
            assert np.array_equal(y, x) and walked == x.shape[0], name
            if x.shape[1] > 20:
                f, l = x.shape[1] // 2 - 5, x.shape[1] // 2 + 5
                y, _ = H.decode(comp, st, nb, x.shape[1], f, l)
                assert np.array_equal(y, x[:, f:l]), name


@pytest.mark.parametrize("log2_amp", [14, 20, 30])
def test_tile_decoder_32bit_prediction_guess_and_miss(H, oracle, log2_amp):
    """Signals that are ~0 at every frame start and grow to 2^log2_amp inside the frame: the warp-tile decoder guesses
    "32-bit prediction sums are exact" from the warm-up; small amplitudes stay there, large ones fail the end-of-subframe
    check and must come back bit-exact from the general decoder (no walker)."""
    rng = np.random.default_rng(500 + log2_amp)
    n = 2 * 4096 + 700
    i = np.arange(n)
    env = np.sin(np.pi * (i % 4096) / 4096.0) ** 2
    amp = float(2 ** log2_amp - 2 ** (log2_amp - 4))
    x = np.array([np.clip(np.rint(amp * env * np.sin(2 * np.pi * i / per + 0.3) + rng.normal(0, 3.0, n)), -2 ** 31, 2 ** 31 - 1)
                  for per in (41.0, 113.0, 390.0)]).astype(np.int32)
    for own, (comp, st, nb) in enumerate((oracle.encode(x, 5), H.encode(x, 5)[:3])):
        y, info = H.decode(comp, st, nb, n, mode=2)       # info = streams walked + 1000 * frames left to the general decoder
        assert np.array_equal(y, x) and info % 1000 == 0
        if log2_amp == 14:
            assert info == 0                              # the guess holds: nothing leaves the tile path
        if log2_amp == 30 and not own:
            assert info >= 1000                           # 15-bit coefficients x 2^30 samples: the check must fire


def test_int64_bodies(H, oracle):
    rng = np.random.default_rng(23)
    a = rng.integers(-2 ** 63, 2 ** 63 - 1, (1, 5000), dtype=np.int64)
    a[0, :4] = [-2 ** 63, 2 ** 63 - 1, 2 ** 32, -2 ** 32]
    b = (np.cumsum(rng.integers(-2 ** 20, 2 ** 20, (1, 9000)), axis=1) + 2 ** 40 * 3).astype(np.int64)
    # BASELINE configs[2] model: the high word is a multiple of 256 (wasted bits) while the walk is positive
    c3 = (np.cumsum(rng.integers(-2 ** 20, 2 ** 20 + 1, (2, 9000)), axis=1) + 2 ** 40 * rng.integers(-4, 5, (2, 9000))).astype(np.int64)
    for x in (a, b, c3):
        c, s, n, _, _ = H.encode(x, 5)
        assert np.array_equal(oracle.decode(c, s, n, x.shape[1], is_int64=True), x)
        oc, os_, on = oracle.encode(x, 5)  # stereo search: exercises side/mid decoding
        for comp, st, nb in ((c, s, n), (oc, os_, on)):
            y, walked = H.decode(comp, st, nb, x.shape[1], is_int64=True)
            assert np.array_equal(y, x) and walked == 0


def test_decoder_bodies_on_golden_third_party_streams(H, golden_dir):
    import glob

    paths = sorted(glob.glob(os.path.join(golden_dir, "ffmpeg_*.npz")) + glob.glob(os.path.join(golden_dir, "handmade_*.npz")))
    for p in paths:
        with np.load(p) as z:
            stream, samples = z["stream"], z["samples"]
        n, nch = samples.shape
        st = np.zeros(1, np.int64)
        nb = np.array([stream.size], np.int64)
        want = samples.reshape(1, -1).view(np.int64) if nch == 2 else samples.reshape(1, -1)
        for mode in (0, 2):
            y, walked = H.decode(stream, st, nb, n, is_int64=(nch == 2), mode=mode)
            assert np.array_equal(y, want), (p, mode)
            assert walked % 1000 == 0, (p, mode)   # no stream needed the sequential walker


def test_float_bodies_match_reference_vectors(H, golden_dir):
    import ctypes as C

    L = H.lib()
    for name, is64, fdt, idt in (("quant_f32", 0, np.float32, np.int32), ("quant_f64", 1, np.float64, np.int64)):
        with np.load(os.path.join(golden_dir, name + ".npz")) as z:
            g = {k: z[k] for k in z.files}
        data = np.ascontiguousarray(g["data"])
        ns, ss = data.shape
        for key, q in (("0", g["quanta0"]), ("1", g["quanta1"]), ("_auto", None)):
            out = np.zeros(data.shape, idt); off = np.zeros(ns, fdt); gain = np.zeros(ns, fdt)
            qq = None if q is None else np.ascontiguousarray(q, fdt)
            L.hs_float_to_int(data.ctypes.data, is64, ns, ss, None if qq is None else qq.ctypes.data, out.ctypes.data,
                              off.ctypes.data, gain.ctypes.data)
            assert np.array_equal(out, g["ints" + key]) and np.array_equal(off, g["off" + key]) and np.array_equal(gain, g["gain" + key])
            rest = np.zeros(data.shape, fdt)
            L.hs_int_to_float(out.ctypes.data, is64, ns, ss, off.ctypes.data, gain.ctypes.data, rest.ctypes.data)
            assert np.array_equal(rest, g["restored" + key])
        # fused quantise + encode of floats decodes to the reference's integers
        c, s, n, off, gain = H.encode(data, 5, quanta=g["quanta0"])
        from oracle import oracle as O
        assert np.array_equal(O.decode(c, s, n, ss, is_int64=bool(is64)), g["ints0"])
        assert np.array_equal(off, g["off0"]) and np.array_equal(gain, g["gain0"])


def test_corrupt_frame_is_caught_by_the_crc_pass(H):
    """A flipped bit inside a frame body still decodes structurally; the frame-parallel CRC-16 pass
    (crc_frame_warp) flags the stream and the walker reports ERROR_DECODE_PROCESS."""
    rng = np.random.default_rng(25)
    x = np.cumsum(rng.integers(-900, 901, (2, 9000)), axis=1).astype(np.int32)
    c, s, n, _, _ = H.encode(x, 5)
    y, walked = H.decode(c, s, n, 9000, mode=2)
    assert np.array_equal(y, x) and walked == 0
    bad = c.copy()
    bad[s[1] + 3000] ^= 0x10
    for mode in (0, 2):
        with pytest.raises(RuntimeError, match="16384"):
            H.decode(bad, s, n, 9000, mode=mode)


def test_fast_quantiser_matches_reference_sequence(H):
    """quant_f32_fast (k_enc_analyze) == quant_f32 (utils.c:232-240 restated) on ties, wide values, NaN, odd gains."""
    import ctypes as C

    L = H.lib()
    L.hs_quant_fast_mismatches.restype = C.c_int64
    L.hs_quant_fast_mismatches.argtypes = [C.c_void_p, C.c_int64, C.c_float, C.c_float]
    rng = np.random.default_rng(31)
    for off, gain in ((0.0, 1.0), (0.0, 1e4), (1.2345, 1e4), (-3.3, 1e7), (0.5, 2.0), (0.0, -1e4), (0.0, 0.0), (7.0, 3.4e38)):
        parts = [
            rng.normal(0, 1, 200000), rng.normal(0, 1e3, 50000), rng.uniform(-500, 500, 50000),
            np.arange(-4000, 4000) * 0.5, np.arange(-4000, 4000) * 0.5 / 1e4, np.arange(-4000, 4000) * 0.25 + off,
            (np.arange(-2000, 2000) + 0.5) / max(abs(gain), 1e-30) + off,
            np.array([0.0, -0.0, np.nan, np.inf, -np.inf, 1e-45, -1e-45, 3.4e38, -3.4e38, 4194303.5, 4194304.0, -4194304.5,
                      8388607.5, 8388608.0, 2147483520.0, 2147483648.0, -2147483648.0, -2147483904.0, 0.49999997, -0.49999997]),
            np.float32(2.0) ** rng.integers(-30, 40, 2000) * rng.choice([-1, 1], 2000),
        ]
        x = np.ascontiguousarray(np.concatenate(parts), np.float32)
        bad = L.hs_quant_fast_mismatches(x.ctypes.data, x.size, C.c_float(off), C.c_float(gain))
        assert bad == 0, (off, gain, bad)
