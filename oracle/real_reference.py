"""Locator + thin adapters for a REAL libFLAC-backed reference, when the machine has one.

TEST INFRASTRUCTURE ONLY (imported by tests/test_real_reference.py and scripts/run_configs.py; never by the
flacarray_b200 package).  The codec arithmetic of hpc4cmb/flacarray lives in libFLAC (meson.build:13, >= 1.4.0;
wheels pin 1.5.0), which is absent from the build image and from /root/reference, so byte / size parity with it is
"parity unpinned" (DESIGN.md section 2).  This module is the hook that pins it as soon as a reference appears:

  1. `import flacarray`                      -- the reference package itself (its own compiled extension);
  2. `baseline/_ref` on sys.path             -- a driver-provided install of the reference (see .gitignore);
  3. `ctypes.util.find_library("FLAC")`      -- a bare libFLAC: driven through ctypes with the same calls, in the
                                                same order and with the same settings as the reference's
                                                compress.c:184-237 and decompress.c:256-305.

`find()` returns an object with `encode(int32/int64 [n, L], level) -> (bytes, starts, nbytes)` and
`decode(bytes, starts, nbytes, L, is_int64) -> ints`, plus `.kind`, or None.
"""
import ctypes as C
import ctypes.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _PackageReference:
    """The reference Python package (libflacarray.pyx:529-594 encode_flac, :713-823 decode_flac)."""

    def __init__(self, mod, kind):
        self.kind = kind
        self._lib = mod

    def encode(self, x, level=5):
        x = np.ascontiguousarray(x)
        comp, starts, nbytes = self._lib.encode_flac(x, int(level), False)
        return np.asarray(comp), np.asarray(starts).reshape(-1), np.asarray(nbytes).reshape(-1)

    def decode(self, comp, starts, nbytes, stream_size, is_int64=False, first=-1, last=-1):
        return np.asarray(self._lib.decode_flac(np.ascontiguousarray(comp, np.uint8), np.ascontiguousarray(starts, np.int64),
                                                np.ascontiguousarray(nbytes, np.int64), int(stream_size), first, last, False, is_int64))


# ---- bare libFLAC through ctypes --------------------------------------------------------------------------------
_WRITE_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_ubyte), C.c_size_t, C.c_uint32, C.c_uint32, C.c_void_p)
_READ_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_ubyte), C.POINTER(C.c_size_t), C.c_void_p)
_SEEK_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint64, C.c_void_p)
_TELL_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p)
_LEN_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p)
_EOF_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p)
_DWRITE_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.POINTER(C.c_int32)), C.c_void_p)
_META_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_void_p)
_ERR_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_void_p)


class _LibFlacReference:
    """libFLAC's stream encoder / decoder with the reference's settings: bits_per_sample 32, channels 1 | 2,
    blocksize 0 (the level's default), compression level, no seek callback on the encoder (compress.c:207-214);
    full decode with process_until_end_of_stream (decompress.c:274-280)."""

    kind = "libFLAC (ctypes)"

    def __init__(self, path):
        L = C.CDLL(path)
        self.L = L
        for name, res, args in (
            ("FLAC__stream_encoder_new", C.c_void_p, []),
            ("FLAC__stream_encoder_delete", None, [C.c_void_p]),
            ("FLAC__stream_encoder_set_compression_level", C.c_int, [C.c_void_p, C.c_uint32]),
            ("FLAC__stream_encoder_set_blocksize", C.c_int, [C.c_void_p, C.c_uint32]),
            ("FLAC__stream_encoder_set_channels", C.c_int, [C.c_void_p, C.c_uint32]),
            ("FLAC__stream_encoder_set_bits_per_sample", C.c_int, [C.c_void_p, C.c_uint32]),
            ("FLAC__stream_encoder_init_stream", C.c_int, [C.c_void_p, _WRITE_CB, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
            ("FLAC__stream_encoder_process_interleaved", C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32]),
            ("FLAC__stream_encoder_finish", C.c_int, [C.c_void_p]),
            ("FLAC__stream_decoder_new", C.c_void_p, []),
            ("FLAC__stream_decoder_delete", None, [C.c_void_p]),
            ("FLAC__stream_decoder_init_stream", C.c_int, [C.c_void_p, _READ_CB, _SEEK_CB, _TELL_CB, _LEN_CB, _EOF_CB, _DWRITE_CB,
                                                           _META_CB, _ERR_CB, C.c_void_p]),
            ("FLAC__stream_decoder_process_until_end_of_stream", C.c_int, [C.c_void_p]),
            ("FLAC__stream_decoder_finish", C.c_int, [C.c_void_p]),
        ):
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args

    def _encode_stream(self, inter, nch, level):
        L = self.L
        chunks = []

        def wr(enc, buf, nbytes, samples, frame, client):
            chunks.append(C.string_at(buf, nbytes))
            return 0

        cb = _WRITE_CB(wr)
        enc = L.FLAC__stream_encoder_new()
        try:
            ok = (L.FLAC__stream_encoder_set_compression_level(enc, level) and L.FLAC__stream_encoder_set_blocksize(enc, 0)
                  and L.FLAC__stream_encoder_set_channels(enc, nch) and L.FLAC__stream_encoder_set_bits_per_sample(enc, 32))
            if not ok or L.FLAC__stream_encoder_init_stream(enc, cb, None, None, None, None) != 0:
                raise RuntimeError("libFLAC encoder init failed")
            if not L.FLAC__stream_encoder_process_interleaved(enc, inter.ctypes.data, inter.size // nch):
                raise RuntimeError("libFLAC process_interleaved failed")
            if not L.FLAC__stream_encoder_finish(enc):
                raise RuntimeError("libFLAC finish failed")
        finally:
            L.FLAC__stream_encoder_delete(enc)
        return b"".join(chunks)

    def encode(self, x, level=5):
        x = np.ascontiguousarray(x)
        nch = 2 if x.dtype == np.int64 else 1
        x2 = x.reshape(-1, x.shape[-1])
        parts = [self._encode_stream(np.ascontiguousarray(row).view(np.int32), nch, int(level)) for row in x2]   # utils.c:112-116
        nbytes = np.array([len(p) for p in parts], np.int64)
        starts = np.cumsum(nbytes) - nbytes
        return np.frombuffer(b"".join(parts), np.uint8).copy(), starts, nbytes

    def _decode_stream(self, data, stream_size, nch):
        L = self.L
        state = {"pos": 0, "out": np.zeros((stream_size, nch), np.int32), "n": 0, "err": 0}

        def rd(dec, buf, nbytes, client):
            want = nbytes[0]
            left = len(data) - state["pos"]
            if left <= 0:
                nbytes[0] = 0
                return 1            # END_OF_STREAM
            n = min(want, left)
            C.memmove(buf, data[state["pos"]:state["pos"] + n], n)
            state["pos"] += n
            nbytes[0] = n
            return 0

        def seek(dec, off, client):
            state["pos"] = int(off)
            return 0

        def tell(dec, off, client):
            off[0] = state["pos"]
            return 0

        def length(dec, ln, client):
            ln[0] = len(data)
            return 0

        def eof(dec, client):
            return 1 if state["pos"] >= len(data) else 0

        def wr(dec, frame, bufs, client):
            bs = C.cast(frame, C.POINTER(C.c_uint32))[0]     # FLAC__Frame.header.blocksize is the first field
            n0 = state["n"]
            take = min(bs, stream_size - n0)
            for c in range(nch):
                state["out"][n0:n0 + take, c] = np.ctypeslib.as_array(bufs[c], (bs,))[:take]
            state["n"] = n0 + take
            return 0

        def meta(dec, m, client):
            return None

        def err(dec, status, client):
            state["err"] += 1

        cbs = (_READ_CB(rd), _SEEK_CB(seek), _TELL_CB(tell), _LEN_CB(length), _EOF_CB(eof), _DWRITE_CB(wr), _META_CB(meta), _ERR_CB(err))
        dec = L.FLAC__stream_decoder_new()
        try:
            if L.FLAC__stream_decoder_init_stream(dec, *cbs, None) != 0:
                raise RuntimeError("libFLAC decoder init failed")
            if not L.FLAC__stream_decoder_process_until_end_of_stream(dec) or state["err"]:
                raise RuntimeError("libFLAC decode failed")
            L.FLAC__stream_decoder_finish(dec)
        finally:
            L.FLAC__stream_decoder_delete(dec)
        if state["n"] != stream_size:
            raise RuntimeError(f"libFLAC decoded {state['n']} of {stream_size} samples")
        return state["out"]

    def decode(self, comp, starts, nbytes, stream_size, is_int64=False, first=-1, last=-1):
        comp = bytes(np.ascontiguousarray(comp, np.uint8))
        nch = 2 if is_int64 else 1
        rows = []
        for s, n in zip(np.asarray(starts).reshape(-1), np.asarray(nbytes).reshape(-1)):
            o = self._decode_stream(comp[int(s):int(s) + int(n)], stream_size, nch)
            rows.append(np.ascontiguousarray(o).reshape(-1).view(np.int64) if is_int64 else o[:, 0].copy())
        out = np.stack(rows)
        if first >= 0 and last >= 0:
            out = out[:, first:last]
        return out


def find():
    """The best available real reference, or None (the usual case in the build image)."""
    try:
        import flacarray.libflacarray as m      # noqa: F401 - the reference package, compiled against libFLAC

        return _PackageReference(m, "flacarray package (import flacarray)")
    except Exception:  # noqa: BLE001
        pass
    ref = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(ref):
        sys.path.insert(0, ref)
        try:
            import flacarray.libflacarray as m

            return _PackageReference(m, "flacarray package (baseline/_ref)")
        except Exception:  # noqa: BLE001
            sys.path.remove(ref)
    path = ctypes.util.find_library("FLAC")
    if path:
        try:
            return _LibFlacReference(path)
        except Exception:  # noqa: BLE001
            return None
    return None
