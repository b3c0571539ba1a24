"""Third-party FLAC codec (FFmpeg's native `flac` encoder/decoder) through ctypes.

TEST INFRASTRUCTURE ONLY -- an independent implementation of the FLAC bit-stream used to pin
the oracle (and, through it, the CUDA path) in both directions.  It is NOT libFLAC and says
nothing about size parity with the reference; it validates format conformance.

The library is the `libavcodec` that ships inside the opencv_python_headless wheel of this
image (same image on the GPU box).  No headers are installed, so the few struct offsets used
are discovered / sanity-checked at run time.  If anything does not line up `available()`
returns False and the tests that need it skip.
"""
import ctypes as C
import glob
import os
import sys

import numpy as np

_state = {}


def _libdir():
    for p in sys.path:
        d = os.path.join(p, "opencv_python_headless.libs")
        if os.path.isdir(d):
            return d
    return None


def _load():
    if "av" in _state:
        return _state["av"]
    _state["av"] = None
    d = _libdir()
    if d is None:
        return None
    try:
        mode = C.RTLD_GLOBAL
        # dependency order: everything libavcodec needs must already be global
        order = ["libcrypto", "libssl", "libdrm", "libpng16", "libavutil", "libswresample", "libvpx",
                 "libaom", "libavif", "libavcodec"]
        libs = {}
        for name in order:
            hits = sorted(glob.glob(os.path.join(d, name + "-*.so*")))
            if not hits:
                continue
            try:
                libs[name] = C.CDLL(hits[0], mode=mode)
            except OSError:
                if name in ("libavutil", "libavcodec"):
                    raise
        avc, avu = libs["libavcodec"], libs["libavutil"]
    except (OSError, KeyError):
        return None
    vp = C.c_void_p
    avc.avcodec_find_decoder_by_name.restype = vp
    avc.avcodec_find_decoder_by_name.argtypes = [C.c_char_p]
    avc.avcodec_find_encoder_by_name.restype = vp
    avc.avcodec_find_encoder_by_name.argtypes = [C.c_char_p]
    avc.avcodec_alloc_context3.restype = vp
    avc.avcodec_alloc_context3.argtypes = [vp]
    avc.avcodec_open2.restype = C.c_int
    avc.avcodec_open2.argtypes = [vp, vp, vp]
    avc.avcodec_free_context.argtypes = [C.POINTER(vp)]
    avc.av_packet_alloc.restype = vp
    avc.av_packet_free.argtypes = [C.POINTER(vp)]
    avc.av_new_packet.restype = C.c_int
    avc.av_new_packet.argtypes = [vp, C.c_int]
    avc.av_packet_unref.argtypes = [vp]
    avc.avcodec_send_packet.restype = C.c_int
    avc.avcodec_send_packet.argtypes = [vp, vp]
    avc.avcodec_receive_frame.restype = C.c_int
    avc.avcodec_receive_frame.argtypes = [vp, vp]
    avc.avcodec_send_frame.restype = C.c_int
    avc.avcodec_send_frame.argtypes = [vp, vp]
    avc.avcodec_receive_packet.restype = C.c_int
    avc.avcodec_receive_packet.argtypes = [vp, vp]
    avu.av_frame_alloc.restype = vp
    avu.av_frame_free.argtypes = [C.POINTER(vp)]
    avu.av_frame_unref.argtypes = [vp]
    avu.av_frame_get_buffer.restype = C.c_int
    avu.av_frame_get_buffer.argtypes = [vp, C.c_int]
    avu.av_opt_set.restype = C.c_int
    avu.av_opt_set.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_int]
    avu.av_opt_set_int.restype = C.c_int
    avu.av_opt_set_int.argtypes = [vp, C.c_char_p, C.c_int64, C.c_int]
    avu.av_opt_get_int.restype = C.c_int
    avu.av_opt_get_int.argtypes = [vp, C.c_char_p, C.c_int, C.POINTER(C.c_int64)]
    _state["av"] = (avc, avu)
    return _state["av"]


# AVPacket: buf(8) pts(8) dts(8) data(8)@24 size(4)@32
_PKT_DATA, _PKT_SIZE = 24, 32
# AVFrame: data[8]@0, linesize[8]@64, extended_data@96, width@104, height@108, nb_samples@112, format@116
_FR_DATA0, _FR_NB, _FR_FMT = 0, 112, 116
_FR_SR, _FR_CHL = 180, 384  # verified by the self-test in available()
_S32 = 2  # AV_SAMPLE_FMT_S32 (interleaved)


def _rd(ptr, off, ctype):
    return ctype.from_address(ptr + off).value


def _wr(ptr, off, ctype, val):
    ctype.from_address(ptr + off).value = val


def _ctx_offsets(avc, avu, codec):
    """Find AVCodecContext.sample_rate by setting a sentinel through the AVOption API."""
    if "ctxoff" in _state:
        return _state["ctxoff"]
    ctx = avc.avcodec_alloc_context3(codec)
    avu.av_opt_set_int(ctx, b"ar", 44123, 0)
    off = None
    for o in range(0, 1024, 4):
        if _rd(ctx, o, C.c_int32) == 44123:
            off = o
            break
    p = C.c_void_p(ctx)
    avc.avcodec_free_context(C.byref(p))
    if off is None:
        raise RuntimeError("cannot locate AVCodecContext.sample_rate")
    chl = (off + 8 + 7) // 8 * 8
    _state["ctxoff"] = {"sample_rate": off, "sample_fmt": off + 4, "ch_layout": chl, "frame_size": chl + 24}
    return _state["ctxoff"]


def decode_frames(frames, n_channels):
    """Decode a list of raw FLAC frames (bytes, each starting FF F8) -> int32 [n, n_channels]."""
    avc, avu = _load()
    codec = avc.avcodec_find_decoder_by_name(b"flac")
    ctx = avc.avcodec_alloc_context3(codec)
    if avc.avcodec_open2(ctx, codec, None) < 0:
        raise RuntimeError("avcodec_open2(flac decoder) failed")
    pkt = avc.av_packet_alloc()
    fr = avu.av_frame_alloc()
    out = []
    try:
        for fb in frames:
            if avc.av_new_packet(pkt, len(fb)) < 0:
                raise RuntimeError("av_new_packet failed")
            C.memmove(_rd(pkt, _PKT_DATA, C.c_void_p), fb, len(fb))
            r = avc.avcodec_send_packet(ctx, pkt)
            avc.av_packet_unref(pkt)
            if r < 0:
                raise RuntimeError(f"avcodec_send_packet failed ({r})")
            while True:
                r = avc.avcodec_receive_frame(ctx, fr)
                if r < 0:
                    break
                nb = _rd(fr, _FR_NB, C.c_int32)
                fmt = _rd(fr, _FR_FMT, C.c_int32)
                if fmt != _S32:
                    raise RuntimeError(f"unexpected sample format {fmt}")
                d0 = _rd(fr, _FR_DATA0, C.c_void_p)
                arr = np.ctypeslib.as_array((C.c_int32 * (nb * n_channels)).from_address(d0)).copy()
                out.append(arr.reshape(nb, n_channels))
                avu.av_frame_unref(fr)
    finally:
        p = C.c_void_p(fr); avu.av_frame_free(C.byref(p))
        p = C.c_void_p(pkt); avc.av_packet_free(C.byref(p))
        p = C.c_void_p(ctx); avc.avcodec_free_context(C.byref(p))
    return np.concatenate(out, axis=0) if out else np.zeros((0, n_channels), np.int32)


def encode_frames(samples, level=5, ch_mode=None, options=None):
    """Encode int32 [n, n_channels] at 32 bps -> (list of frame bytes, frame_size)."""
    avc, avu = _load()
    samples = np.ascontiguousarray(samples, dtype=np.int32)
    n, nch = samples.shape
    codec = avc.avcodec_find_encoder_by_name(b"flac")
    off = _ctx_offsets(avc, avu, codec)
    ctx = avc.avcodec_alloc_context3(codec)
    avu.av_opt_set_int(ctx, b"ar", 44100, 0)
    avu.av_opt_set(ctx, b"ch_layout", b"mono" if nch == 1 else b"stereo", 0)
    avu.av_opt_set(ctx, b"strict", b"-2", 0)
    avu.av_opt_set_int(ctx, b"bits_per_raw_sample", 32, 0)
    avu.av_opt_set_int(ctx, b"compression_level", level, 0)
    _wr(ctx, off["sample_fmt"], C.c_int32, _S32)
    if ch_mode is not None:
        if avu.av_opt_set(ctx, b"ch_mode", ch_mode.encode(), 1) < 0:
            raise RuntimeError("ch_mode option rejected")
    for k, v in (options or {}).items():
        if avu.av_opt_set(ctx, k.encode(), str(v).encode(), 1) < 0:
            raise RuntimeError(f"option {k} rejected")
    if avc.avcodec_open2(ctx, codec, None) < 0:
        raise RuntimeError("avcodec_open2(flac encoder) failed")
    fsize = _rd(ctx, off["frame_size"], C.c_int32)
    pkt = avc.av_packet_alloc()
    fr = avu.av_frame_alloc()
    frames = []

    def drain():
        while True:
            r = avc.avcodec_receive_packet(ctx, pkt)
            if r < 0:
                break
            sz = _rd(pkt, _PKT_SIZE, C.c_int32)
            if sz > 0:
                frames.append(C.string_at(_rd(pkt, _PKT_DATA, C.c_void_p), sz))
            avc.av_packet_unref(pkt)

    try:
        pos = 0
        while pos < n:
            nb = min(fsize, n - pos)
            _wr(fr, _FR_NB, C.c_int32, nb)
            _wr(fr, _FR_FMT, C.c_int32, _S32)
            _wr(fr, _FR_SR, C.c_int32, 44100)
            # ch_layout: {order=1 (native), nb_channels, mask, opaque}
            _wr(fr, _FR_CHL, C.c_int32, 1)
            _wr(fr, _FR_CHL + 4, C.c_int32, nch)
            _wr(fr, _FR_CHL + 8, C.c_uint64, 0x4 if nch == 1 else 0x3)
            if avu.av_frame_get_buffer(fr, 0) < 0:
                raise RuntimeError("av_frame_get_buffer failed")
            d0 = _rd(fr, _FR_DATA0, C.c_void_p)
            blk = samples[pos:pos + nb]
            C.memmove(d0, blk.ctypes.data, blk.nbytes)
            r = avc.avcodec_send_frame(ctx, fr)
            avu.av_frame_unref(fr)
            if r < 0:
                raise RuntimeError(f"avcodec_send_frame failed ({r})")
            drain()
            pos += nb
        avc.avcodec_send_frame(ctx, None)
        drain()
    finally:
        p = C.c_void_p(fr); avu.av_frame_free(C.byref(p))
        p = C.c_void_p(pkt); avc.av_packet_free(C.byref(p))
        p = C.c_void_p(ctx); avc.avcodec_free_context(C.byref(p))
    return frames, fsize


def make_stream(frames, n_channels, blocksize, total_samples=0):
    """Wrap raw frames into a complete FLAC file: fLaC + STREAMINFO(last) + frames."""
    si = bytearray(34)
    si[0] = si[2] = (blocksize >> 8) & 0xFF
    si[1] = si[3] = blocksize & 0xFF
    sr = 44100
    si[10] = (sr >> 12) & 0xFF
    si[11] = (sr >> 4) & 0xFF
    si[12] = ((sr & 0xF) << 4) | ((n_channels - 1) << 1) | 1
    si[13] = 0xF0 | ((total_samples >> 32) & 0xF)
    si[14:18] = int(total_samples & 0xFFFFFFFF).to_bytes(4, "big")
    return b"fLaC" + bytes([0x80, 0, 0, 34]) + bytes(si) + b"".join(frames)


def available():
    """True when the FFmpeg codec loads and a hand-made VERBATIM frame round-trips through it."""
    if "ok" in _state:
        return _state["ok"]
    _state["ok"] = False
    try:
        if _load() is None:
            return False
        x = (np.arange(192, dtype=np.int64) * 7919 % 2001 - 1000).astype(np.int32).reshape(-1, 1)
        frames, _ = encode_frames(x, level=5)
        y = decode_frames(frames, 1)
        _state["ok"] = bool(np.array_equal(x, y))
    except Exception:
        _state["ok"] = False
    return _state["ok"]
