"""GPU ports of the reference's own tests (src/flacarray/tests/{bindings,utils,array}.py): round-trip
identities, slice semantics, float tolerances, FlacArray slicing shapes and keep masks."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fa():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import __graft_entry__ as g

    g.build()
    import flacarray_b200

    return flacarray_b200


def test_wrappers_i32_i64(fa):
    """reference tests/bindings.py:27-163"""
    from flacarray_b200 import libflacarray as lf

    rng = np.random.default_rng()
    n_stream, stream_size = 3, 10000
    for dt, enc, dec, ext in ((np.int32, lf.wrap_encode_i32_threaded, lf.wrap_decode_i32, [2147483647, -2147483647]),
                              (np.int64, lf.wrap_encode_i64_threaded, lf.wrap_decode_i64,
                               [2 ** 63 - 1, -(2 ** 63 - 1), 2 ** 32, -2 ** 32])):
        ii = np.iinfo(dt)
        data = rng.integers(ii.min, ii.max, (n_stream * stream_size), dtype=np.int64).astype(dt)
        data[:len(ext)] = ext
        comp, starts, nbytes = enc(data, n_stream, stream_size, 5)
        out = dec(comp, starts, nbytes, n_stream, stream_size, -1, -1, True)
        assert np.array_equal(out, data)
        first, last = stream_size // 2 - 5, stream_size // 2 + 5
        out = dec(comp, starts, nbytes, n_stream, stream_size, first, last, True)
        assert np.array_equal(out.reshape(n_stream, -1), data.reshape(n_stream, -1)[:, first:last])


def test_encode_decode_roundtrip(fa):
    """reference tests/bindings.py:165-230"""
    from flacarray_b200.demo import create_fake_data
    from flacarray_b200.libflacarray import decode_flac, encode_flac

    for shape in ((4, 3, 1000), (10000,)):
        for dt in (np.int32, np.int64):
            data, _ = create_fake_data(shape, sigma=None, dtype=dt)
            comp, starts, nbytes = encode_flac(data, 5)
            assert starts.shape == (shape[:-1] if len(shape) > 1 else (1,))
            out = decode_flac(comp, starts, nbytes, shape[-1], is_int64=(dt == np.int64))
            assert np.array_equal(out.reshape(data.shape), data)
            first, last = shape[-1] // 2 - 5, shape[-1] // 2 + 5
            out = decode_flac(comp, starts, nbytes, shape[-1], first_sample=first, last_sample=last, is_int64=(dt == np.int64))
            assert np.array_equal(out.reshape(data.shape[:-1] + (10,)), data[..., first:last])


def test_float_to_int_roundtrips(fa):
    """reference tests/utils.py:22-108"""
    from flacarray_b200.demo import create_fake_data

    data, _ = create_fake_data((4, 3, 1000), sigma=1.0, dtype=np.float64)
    for kw, tol in (({"quanta": 1e-16}, 1e-14), ({"quanta": 1e-5}, 1e-5), ({"precision": 5}, 1e-4),
                    ({"precision": 5 * np.ones((4, 3), dtype=np.int32)}, 1e-4)):
        idata, off, gain = fa.float_to_int(data, **kw)
        assert idata.dtype == np.int64 and off.shape == (4, 3)
        assert np.allclose(fa.int_to_float(idata, off, gain), data, rtol=0, atol=tol * 10)
    d32 = data.astype(np.float32)
    for kw, tol in (({"quanta": 1e-6}, 1e-5), ({"quanta": 1e-5}, 1e-5), ({"precision": 5}, 1e-4)):
        idata, off, gain = fa.float_to_int(d32, **kw)
        assert idata.dtype == np.int32
        assert np.allclose(fa.int_to_float(idata, off, gain), d32, rtol=0, atol=max(tol, 1e-5) * 10)


def test_precision_quanta_on_device(fa):
    """reference utils.py:282-296: quanta = np.std(data, axis=-1) / 10**precision.  The standard deviation is reduced on
    the device (fab_stream_std, double-precision moments); numpy sums float32 pairwise in single precision, so the two
    agree to rounding (tolerances below), and encoding with `precision` is byte-identical to encoding with the quanta
    derived from the device's standard deviation -- in one shot, through the chunked host pipeline, and for a CUDA tensor."""
    import torch
    from flacarray_b200 import libflacarray as lf

    rng = np.random.default_rng(20261018)
    for dt, shape, rtol in ((np.float32, (6, 50000), 3e-6), (np.float64, (5, 70001), 1e-12), (np.float32, (2, 3, 1000), 3e-6)):
        scale = 10.0 ** rng.integers(-3, 4, shape[:-1] + (1,))
        data = (rng.normal(0, 1, shape) * scale + 100.0 * scale).astype(dt)
        sd = lf.stream_std(data)
        assert sd.dtype == dt and sd.shape == shape[:-1]
        assert np.allclose(sd, np.std(data.astype(np.float64), axis=-1), rtol=rtol, atol=0)
        assert np.allclose(sd, np.std(data, axis=-1), rtol=30 * rtol, atol=0)
        sd_t = lf.stream_std(torch.from_numpy(data).cuda())
        assert np.array_equal(sd_t, sd)
        for prec in (4, rng.integers(2, 6, shape[:-1])):
            q = np.asarray(lf.quanta_from_std(sd, prec)).astype(dt)
            a = fa.array_compress(data, precision=prec)
            b = fa.array_compress(data, quanta=q)
            c = fa.array_compress(torch.from_numpy(data).cuda(), precision=prec)
            for x, y, z in zip(a, b, c):
                assert np.array_equal(x, y)
                assert np.array_equal(x, z.cpu().numpy() if torch.is_tensor(z) else z)
            idata, off, gain = fa.float_to_int(data, precision=prec)
            idata2, off2, gain2 = fa.float_to_int(data, quanta=q)
            assert np.array_equal(idata, idata2) and np.array_equal(off, off2) and np.array_equal(gain, gain2)
    # a host array large enough for the chunked pipeline (chunks of whole streams, quanta per chunk on the device)
    big = (rng.normal(0, 1, (64, 300000)) * np.linspace(0.5, 50, 64)[:, None]).astype(np.float32)
    sd = lf.stream_std(big)
    assert np.allclose(sd, np.std(big.astype(np.float64), axis=-1), rtol=3e-6, atol=0)
    q = np.asarray(lf.quanta_from_std(sd, 5)).astype(np.float32)
    a = fa.array_compress(big, precision=5)
    b = fa.array_compress(big, quanta=q)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    out = fa.array_decompress(a[0], big.shape[-1], a[1], a[2], stream_offsets=a[3], stream_gains=a[4])
    assert np.max(np.abs(out - big) / sd[:, None]) < 1e-5


def test_helpers_all_dtypes(fa):
    """reference tests/array.py:26-146"""
    from flacarray_b200.demo import create_fake_data

    for shape in ((4, 3, 1000), (10000,)):
        first, last = shape[-1] // 2 - 5, shape[-1] // 2 + 5
        for dt, kw, tol in ((np.int32, {}, 0), (np.int64, {}, 0), (np.float32, {"quanta": 1e-6}, 1e-5),
                            (np.float64, {"quanta": 1e-7}, 1e-6)):
            sig = None if np.dtype(dt).kind == "i" else 1.0
            data, _ = create_fake_data(shape, sigma=sig, dtype=dt)
            comp, starts, nbytes, off, gain = fa.array_compress(data, level=5, **kw)
            is64 = np.dtype(dt).itemsize == 8
            out = fa.array_decompress(comp, shape[-1], starts, nbytes, stream_offsets=off, stream_gains=gain, is_int64=is64)
            assert out.shape == data.shape and out.dtype == np.dtype(dt)
            assert np.allclose(out, data, rtol=0, atol=tol)
            out = fa.array_decompress(comp, shape[-1], starts, nbytes, stream_offsets=off, stream_gains=gain,
                                      first_stream_sample=first, last_stream_sample=last, is_int64=is64)
            assert np.allclose(out, data[..., first:last], rtol=0, atol=tol)


def test_flacarray_slicing_and_keep(fa):
    """reference tests/array.py:148-232 (values are compared too, not only shapes)"""
    from flacarray_b200.demo import create_fake_data

    data, _ = create_fake_data((4, 3, 1000), sigma=1.0, dtype=np.float64)
    far = fa.FlacArray.from_array(data, quanta=1e-16)
    assert far.shape == data.shape and far.dtype == np.float64 and far.stream_size == 1000
    assert far.nstreams == 12 and far.global_nbytes == far.nbytes == far.compressed.nbytes
    assert np.allclose(far.to_array(), data, rtol=0, atol=1e-13)
    assert np.allclose(far.to_array(stream_slice=slice(100, 200)), data[..., 100:200], rtol=0, atol=1e-13)
    keys = [(slice(None),), (1,), (slice(1, 3), slice(None), slice(10, 20)), (1, 2, slice(None)),
            (slice(None), 1, slice(-20, None)), (3, slice(0, 2), 500), (2, 1, 17), (slice(None), slice(None), 999),
            (slice(0, 0),), (7,), (1, slice(None), slice(30, 30))]
    for key in keys:
        got = far[key]
        try:
            want = data[key]
        except IndexError:
            want = np.zeros(got.shape)
        assert got.shape == want.shape, key
        if want.size:
            assert np.allclose(got, want, rtol=0, atol=1e-13), key
    keep = np.zeros((4, 3), bool)
    keep[1, 2] = keep[3, 0] = keep[0, 0] = True
    arr, idx = far.to_array(keep=keep, stream_slice=slice(200, 300), keep_indices=True)
    assert arr.shape == (3, 100) and idx == [(0, 0), (1, 2), (3, 0)]
    for row, i in zip(arr, idx):
        assert np.allclose(row, data[i][200:300], rtol=0, atol=1e-13)
    one, _ = create_fake_data((5000,), sigma=1.0, dtype=np.float32)
    f1 = fa.FlacArray.from_array(one, quanta=1e-6)
    assert f1.to_array().shape == (5000,) and f1[10:20].shape == (10,) and f1.stream_starts.shape == (1,)
    assert np.allclose(f1[10:20], one[10:20], rtol=0, atol=1e-5)
    cp = fa.FlacArray(far)
    assert cp == far and not (cp == f1)
    with pytest.raises(RuntimeError):
        far[0] = 1


def test_quantization_error_bounds(fa):
    """reference tests/array.py:234-283"""
    rng = np.random.default_rng(7)
    n = 10000
    quanta = 1e-3
    for dc in (0.0, 0.5, -0.5, 10.0, -10.0, -10.51, -10.4):
        data = (rng.normal(0, 1, n) + dc).astype(np.float64)
        far = fa.FlacArray.from_array(data, quanta=quanta)
        assert np.max(np.abs(far.to_array() - data)) <= 0.5 * quanta * (1 + 1e-9)
        pre = np.round(data / quanta) * quanta
        far = fa.FlacArray.from_array(pre, quanta=quanta)
        assert np.max(np.abs(far.to_array() - pre)) <= 2 * np.max(np.abs(pre)) * np.finfo(np.float64).eps + 1e-15


def test_host_pipeline_matches_single_shot(fa, monkeypatch):
    """The chunked host-buffer pipeline (H2D / kernels / D2H on three streams) returns exactly what one
    device call returns: same bytes, starts, nbytes, offsets, gains; decode with keep + slice; the
    capacity-overflow repair path (first chunk compresses far better than the rest)."""
    import torch
    from flacarray_b200 import libflacarray as lf

    rng = np.random.default_rng(11)
    n, L = 37, 9000
    cases = []
    walk = (np.cumsum(rng.integers(-300, 301, (n, L)), axis=1)).astype(np.int32)
    cases.append((walk, {}))
    w64 = np.cumsum(rng.integers(-(1 << 33), 1 << 33, (n, L)), axis=1).astype(np.int64)
    cases.append((w64, {}))
    f32 = (rng.normal(0, 1, (n, L)) + np.linspace(-3, 3, n)[:, None]).astype(np.float32)
    cases.append((f32, {"quanta": 1e-4}))
    f64 = rng.normal(0, 1, (n, L))
    cases.append((f64, {"precision": 6}))
    skew = walk.copy()
    skew[:6] = 7                      # constant streams first: the capacity guess is far too small
    skew[6:] = rng.integers(-2 ** 31, 2 ** 31 - 1, (n - 6, L), dtype=np.int64).astype(np.int32)
    cases.append((skew, {}))
    for data, kw in cases:
        monkeypatch.setattr(lf, "_PIPE_MIN_BYTES", 1 << 40)
        ref = fa.array_compress(data, level=5, **kw)
        monkeypatch.setattr(lf, "_PIPE_MIN_BYTES", 1 << 10)
        monkeypatch.setattr(lf, "_PIPE_CHUNK_BYTES", 6 * L * data.itemsize)
        got = fa.array_compress(data, level=5, **kw)
        for a, b in zip(ref, got):
            assert (a is None) == (b is None)
            if a is not None:
                assert a.dtype == b.dtype and np.array_equal(a, b)
        comp, starts, nbytes, off, gain = got
        is64 = data.itemsize == 8
        full = fa.array_decompress(comp, L, starts, nbytes, stream_offsets=off, stream_gains=gain, is_int64=is64)
        monkeypatch.setattr(lf, "_PIPE_MIN_BYTES", 1 << 40)
        full1 = fa.array_decompress(comp, L, starts, nbytes, stream_offsets=off, stream_gains=gain, is_int64=is64)
        assert np.array_equal(full, full1) and full.dtype == data.dtype
        if data.dtype.kind == "i":
            assert np.array_equal(full, data)
        monkeypatch.setattr(lf, "_PIPE_MIN_BYTES", 1 << 10)
        keep = (np.arange(n) % 3) != 1
        part, idx = fa.array_decompress_slice(comp, L, starts, nbytes, stream_offsets=off, stream_gains=gain, keep=keep,
                                              first_stream_sample=4000, last_stream_sample=4700, is_int64=is64)
        assert np.array_equal(part, full[keep][:, 4000:4700]) and len(idx) == int(keep.sum())
    # torch host tensors (pinned) are accepted like numpy arrays
    pin = torch.from_numpy(walk).pin_memory()
    c2 = lf.encode_flac(pin, 5)
    assert np.array_equal(c2[0], fa.array_compress(walk, level=5)[0])


@pytest.mark.parametrize("zarr_style", [False, True])
def test_write_read_array_through_group_layout(fa, zarr_style):
    """reference tests/hdf5.py:27-173 and tests/zarr.py: write_array -> read_array (full, keep mask,
    stream slice) and FlacArray.write_* / read_* with the compression and decompression on the GPU."""
    from flacarray_b200 import hdf5 as fh5
    from flacarray_b200 import zarr as fzr
    from flacarray_b200.demo import create_fake_data
    from flacarray_b200.memgroup import MemGroup

    mod = fzr if zarr_style else fh5
    for dt, kw, tol in ((np.int32, {}, 0), (np.int64, {}, 0), (np.float32, {"quanta": 1e-5}, 1e-5),
                        (np.float64, {"precision": 7}, 1e-6)):
        sig = None if np.dtype(dt).kind == "i" else 1.0
        data, _ = create_fake_data((4, 3, 5000), sigma=sig, dtype=dt)
        grp = MemGroup(zarr_style=zarr_style)
        mod.write_array(data, grp, level=5, **kw)
        assert grp.attrs["flacarray_format_version"] == "1" and grp["stream_starts"].shape == (4, 3)
        assert ("stream_offsets" in grp) == (np.dtype(dt).kind == "f")
        full = mod.read_array(grp)
        assert full.shape == data.shape and full.dtype == np.dtype(dt)
        assert np.allclose(full, data, rtol=0, atol=tol)
        keep = np.zeros((4, 3), bool)
        keep[0, 2] = keep[3, 1] = True
        part, idx = mod.read_array(grp, keep=keep, stream_slice=slice(1000, 1200), keep_indices=True)
        assert part.shape == (2, 200) and idx == [(0, 2), (3, 1)]
        assert np.array_equal(part, full[keep][:, 1000:1200])
        far = fa.FlacArray.from_array(data, **kw)
        g2 = MemGroup(zarr_style=zarr_style)
        (far.write_zarr if zarr_style else far.write_hdf5)(g2)
        back = (fa.FlacArray.read_zarr if zarr_style else fa.FlacArray.read_hdf5)(g2)
        assert back == far and np.array_equal(back.to_array(), far.to_array())
        assert np.array_equal(g2["compressed"][...], grp["compressed"][...])     # deterministic bytes
    one, _ = create_fake_data((6000,), sigma=1.0, dtype=np.float32)
    g3 = MemGroup(zarr_style=zarr_style)
    mod.write_array(one, g3, quanta=1e-6)
    assert g3["stream_starts"].shape == (1,) and mod.read_array(g3).shape == (6000,)


def test_version0_group_read(fa, oracle):
    """hdf5_load_v0.py:357-411: legacy int64 (32-bit samples + int64 stream offsets), float32, and float64
    stored as 32-bit integers with float64 offsets / gains."""
    from flacarray_b200 import hdf5 as fh5
    from test_io_layout import _v0_group

    rng = np.random.default_rng(12)
    ints = np.cumsum(rng.integers(-50, 51, (2, 3, 4000)), axis=-1).astype(np.int32)
    offs = rng.integers(-2 ** 40, 2 ** 40, (2, 3)).astype(np.int64)
    got = fh5.read_array(_v0_group(oracle, ints, offsets=offs))
    assert got.dtype == np.int64 and np.array_equal(got, ints.astype(np.int64) + offs[..., None])
    got = fh5.read_array(_v0_group(oracle, ints))
    assert got.dtype == np.int32 and np.array_equal(got, ints)
    part = fh5.read_array(_v0_group(oracle, ints, offsets=offs), stream_slice=slice(100, 300))
    assert np.array_equal(part, (ints.astype(np.int64) + offs[..., None])[..., 100:300])
    for fdt in (np.float32, np.float64):
        foff = rng.normal(0, 1, (2, 3)).astype(fdt)
        fgain = np.full((2, 3), 1.0e4, dtype=fdt)
        got = fh5.read_array(_v0_group(oracle, ints, offsets=foff, gains=fgain))
        want = oracle.int_to_float(ints.reshape(6, -1), foff.astype(np.float32).reshape(-1),
                                   fgain.astype(np.float32).reshape(-1)).reshape(ints.shape).astype(fdt)
        assert got.dtype == np.dtype(fdt) and np.array_equal(np.asarray(got), want)


def test_benchmark_cli_runs(fa, tmp_path, capsys):
    """scripts/benchmark.py:293-354: full pass, then even streams x 100 middle samples, timer table at the end."""
    from flacarray_b200.scripts.benchmark import cli
    cli(["--out_dir", str(tmp_path / "bench"), "--global_shape", "(4,3,20000)"])
    out = capsys.readouterr().out
    assert "Full Data Tests:" in out and "Sliced Data Tests" in out
    assert out.count("FlacArray compress in") == 4 and out.count("Direct read") == 4


def test_two_host_threads_encode_concurrently(fa):
    """The reference's entry points are re-entrant (no global state); here every thread has its
    own C context and scratch buffers, and ctypes drops the GIL during the calls."""
    import threading

    import torch

    rng = np.random.default_rng(21)
    arrays = [np.cumsum(rng.integers(-300, 301, (40, 50000)), axis=1).astype(np.int32) for _ in range(2)]
    results = [None, None]

    def work(i):
        try:
            ok = True
            for _ in range(6):
                d = torch.from_numpy(arrays[i]).cuda()
                far = fa.FlacArray.from_array(d)
                ok = ok and bool(torch.equal(far.to_array(), d))
                ok = ok and np.array_equal(fa.FlacArray.from_array(arrays[i]).to_array(), arrays[i])
            results[i] = ok
        except BaseException as e:  # noqa: BLE001
            results[i] = e

    threads = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    assert results == [True, True], results


def test_device_encode_capacity_guess_and_overflow_fallback(fa):
    """Device-resident encodes size their output from recent compression ratios; data that compresses
    worse than the guess must fall back to the worst-case buffer and still round-trip exactly."""
    import torch

    from flacarray_b200 import libflacarray as lf

    g = torch.Generator(device="cuda").manual_seed(5)
    n, L = 24, 100000                                     # 9.6 MB of int32: above the guessing threshold
    smooth = torch.cumsum(torch.randint(-3, 4, (n, L), device="cuda", generator=g), dim=1).to(torch.int32)
    noise = torch.randint(-2 ** 31, 2 ** 31 - 1, (n, L), device="cuda", generator=g, dtype=torch.int64).to(torch.int32)
    lf._ratio_history().clear()
    sizes = []
    for data in (smooth, smooth, noise, noise, smooth):
        comp, starts, nbytes, _, _ = lf.encode_device(data.reshape(-1), n, L, 5)
        assert int(nbytes.sum()) == comp.numel() and int(starts[-1] + nbytes[-1]) == comp.numel()
        out = lf.decode_device(comp, starts, nbytes, n, L, -1, -1, False, int(nbytes.max()), 4096)
        assert torch.equal(out.view(n, L), data)
        sizes.append(comp.numel())
    assert sizes[0] == sizes[1] == sizes[4] and sizes[2] == sizes[3] > 3 * sizes[0]
    hist = lf._ratio_history()[(torch.cuda.current_device(), torch.int32, 5)]
    assert len(hist) == 4 and max(hist) > 0.9


def test_caller_supplied_workspace(fa):
    """fab_set_workspace: the calls use the caller's block and never allocate; a block that is too small is an
    ERROR_ALLOC with the needed size in fab_last_error, not a hidden cudaMalloc (include/flacarray_b200.h)."""
    import torch

    from flacarray_b200 import _lib, libflacarray as lf

    L = _lib.lib()
    dev = torch.device("cuda", 0)
    ctx = _lib.context(dev)
    n_stream, n_samp = 24, 30000
    rng = np.random.default_rng(5)
    x = np.cumsum(rng.integers(-200, 201, (n_stream, n_samp)), axis=1).astype(np.int32)
    need_e = L.fab_encode_workspace_bytes(n_stream, n_samp, 0, 5)
    need_d = L.fab_decode_workspace_bytes(n_stream, n_samp, 4096)
    assert need_e > 0 and need_d > 0
    assert L.fab_encode_workspace_bytes(0, n_samp, 0, 5) == 0 and L.fab_encode_workspace_bytes(1, 1, 0, 9) == 0
    ref = fa.array_compress(x, level=5)
    try:
        ws = torch.empty(max(need_e, need_d), dtype=torch.uint8, device=dev)
        ctx.set_workspace(ws)
        before = torch.cuda.memory_allocated(dev)
        got = fa.array_compress(x, level=5)
        assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])
        y = fa.array_decompress(got[0], n_samp, got[1], got[2])
        assert np.array_equal(y, x)
        # too small: refused, with the size in the message
        small = torch.empty(4096, dtype=torch.uint8, device=dev)
        ctx.set_workspace(small)
        with pytest.raises(Exception) as ei:
            fa.array_compress(x, level=5)
        assert "workspace too small" in str(ei.value) or "workspace too small" in ctx.last_error()
        del before
    finally:
        ctx.set_workspace(None)
    got = fa.array_compress(x, level=5)
    assert np.array_equal(got[0], ref[0])


def test_reference_signature_entry_points_are_reentrant(fa):
    """Four host threads call encode_i32 / decode_i32 (host pointers, reference signatures) at once: each leases its
    own host slot (context + staging buffers), results are identical to a serial call."""
    import ctypes as C
    import threading

    from flacarray_b200 import _lib

    L = _lib.lib()
    rng = np.random.default_rng(77)
    arrays = [np.cumsum(rng.integers(-500, 501, (16, 40000)), axis=1).astype(np.int32) for _ in range(4)]

    def enc(a):
        n_bytes = C.c_int64(0)
        starts = np.zeros(a.shape[0], np.int64)
        buf = C.POINTER(C.c_ubyte)()
        rc = L.encode_i32(C.c_void_p(a.ctypes.data), C.c_int64(a.shape[0]), C.c_int64(a.shape[1]), C.c_uint32(5),
                          C.byref(n_bytes), C.c_void_p(starts.ctypes.data), C.byref(buf))
        assert rc == 0, rc
        out = np.ctypeslib.as_array(buf, shape=(n_bytes.value,)).copy()
        C.CDLL(None).free(buf)
        return out, starts

    def dec(b, starts, shape):
        nb = np.diff(np.append(starts, b.size)).astype(np.int64)
        out = np.zeros(shape, np.int32)
        rc = L.decode_i32(C.c_void_p(b.ctypes.data), C.c_void_p(starts.ctypes.data), C.c_void_p(nb.ctypes.data),
                          C.c_int64(shape[0]), C.c_int64(shape[1]), C.c_int64(-1), C.c_int64(-1), C.c_void_p(out.ctypes.data),
                          C.c_bool(False))
        assert rc == 0, rc
        return out

    L.encode_i32.restype = C.c_int
    L.decode_i32.restype = C.c_int
    serial = [enc(a) for a in arrays]
    results = [None] * 4

    def work(i):
        try:
            ok = True
            for _ in range(4):
                b, st = enc(arrays[i])
                ok = ok and np.array_equal(b, serial[i][0]) and np.array_equal(st, serial[i][1])
                ok = ok and np.array_equal(dec(b, st, arrays[i].shape), arrays[i])
            results[i] = ok
        except BaseException as e:  # noqa: BLE001
            results[i] = e

    threads = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=180)
    assert results == [True] * 4, results
