// fa_decode_tile.h -- throughput decoder: one warp decodes 32 frames (one per lane) in lock step.
//
// Same role as frame_body() in fa_decode.h (libFLAC frame decode + the write callback
// decompress.c:66-101), restructured for the machine:
//   * predictor history and coefficients live in registers (order <= 12, 32-bit samples);
//   * every lane appends its samples to a [32 lanes][32 samples] shared-memory tile; after 32 samples
//     the warp transposes the tile so that each store instruction writes 128 contiguous bytes of one
//     frame (fully coalesced), optionally restoring int32 -> float32 on the way (utils.c:350-368);
//   * CRC-16 is folded into the bit reader (slice-by-4 per fetched word, tables in shared memory);
//   * frames this path does not handle (33-bit side channel, order > 12, ...) are flagged and left to
//     the general per-thread decoder.
#pragma once
#include "fa_decode.h"
#include "fa_quant.h"

namespace fa {

constexpr int kTileOrd = 12;
constexpr int kTileStride = 33;
constexpr int kTileWarps = 4;  // warps per CTA

struct TileRow {          // per-lane output description, read by all lanes during the flush
    int32_t* out;         // address of (frame sample 0, channel 0)
    int lo, hi;           // valid sample range inside the frame
    float off, coeff;     // int32 -> float32 restore (when enabled)
};

struct TileShared {       // per warp
    int32_t tile[32 * kTileStride];
    TileRow row[32];
};

// Bit reader with a CRC-16 that lags two words behind the fetch pointer (so that the frame end, which
// is only known after the last sample, can be handled exactly).
struct BitRdC {
    const uint32_t* wp;
    const uint32_t* wend;
    uint64_t buf;
    int n;
    int nwords;        // words fetched so far
    uint32_t crc;      // CRC state before word (nwords - 2)
    uint32_t w1, w2;   // last fetched word (w1) and the one before (w2)
    int err;
};

FA_D uint32_t crc16_word(const uint16_t* T, uint32_t c, uint32_t w) {
    // T: [4][256] slice tables in shared memory
    return (uint32_t)(T[3 * 256 + (((c >> 8) ^ (w >> 24)) & 0xFF)] ^ T[2 * 256 + (((c & 0xFF) ^ (w >> 16)) & 0xFF)] ^
                      T[256 + ((w >> 8) & 0xFF)] ^ T[w & 0xFF]);
}
FA_D uint32_t crc16_b(const uint16_t* T, uint32_t c, uint32_t byte) {
    return ((c << 8) & 0xFFFF) ^ T[((c >> 8) ^ byte) & 0xFF];
}

template <bool CRC>
FA_D void brc_fetch(BitRdC& br, const uint16_t* T) {
    uint32_t w = 0;
    if (br.wp < br.wend) w = bswap32(ldg32(br.wp));
    br.wp++;
    if (CRC) {
        if (br.nwords >= 2) br.crc = crc16_word(T, br.crc, br.w2);
        br.w2 = br.w1;
        br.w1 = w;
    }
    br.nwords++;
    br.buf |= (uint64_t)w << (32 - br.n);
    br.n += 32;
}

// Start reading at byte `start` (crc0 = CRC state over the frame bytes before `start`).
template <bool CRC>
FA_D void brc_init(BitRdC& br, const uint8_t* start, const uint8_t* end, uint32_t crc0, const uint16_t* T) {
    uintptr_t s = (uintptr_t)start;
    br.wp = (const uint32_t*)(s & ~(uintptr_t)3);
    br.wend = (const uint32_t*)(((uintptr_t)end + 3) & ~(uintptr_t)3);
    int a = (int)(s & 3);
    br.buf = 0; br.n = 0; br.nwords = 0; br.err = 0; br.w1 = br.w2 = 0;
    br.crc = crc0;
    uint32_t w = 0;
    if (br.wp < br.wend) w = bswap32(ldg32(br.wp));
    br.wp++;
    // the first word is consumed byte-wise by the CRC (its leading `a` bytes precede `start`), so it
    // does not enter the lagging word queue
    if (CRC) for (int b = a; b < 4; ++b) br.crc = crc16_b(T, br.crc, (w >> (24 - 8 * b)) & 0xFF);
    br.buf = ((uint64_t)w << 32) << (8 * a);
    br.n = 32 - 8 * a;
}

template <bool CRC>
FA_D void brc_refill(BitRdC& br, const uint16_t* T) {
    if (br.n <= 32) brc_fetch<CRC>(br, T);
}

template <bool CRC>
FA_D uint32_t brc_read(BitRdC& br, int nb, const uint16_t* T) {  // nb in [0, 32]
    brc_refill<CRC>(br, T);
    uint32_t v = nb ? (uint32_t)(br.buf >> (64 - nb)) : 0u;
    br.buf = nb ? (br.buf << nb) : br.buf;
    br.n -= nb;
    return v;
}
template <bool CRC>
FA_D int32_t brc_read_signed(BitRdC& br, int nb, const uint16_t* T) {  // nb in [0, 32]
    if (nb == 0) return 0;
    uint32_t v = brc_read<CRC>(br, nb, T);
    uint32_t sign = 1u << (nb - 1);
    return (int32_t)((v ^ sign) - sign);
}
template <bool CRC>
FA_D uint32_t brc_unary(BitRdC& br, const uint16_t* T) {
    uint32_t q = 0;
    for (;;) {
        brc_refill<CRC>(br, T);
        if (br.buf != 0) {
            int z = clz64(br.buf);
            q += (uint32_t)z;
            br.buf = (br.buf << z) << 1;
            br.n -= z + 1;
            return q;
        }
        q += (uint32_t)br.n;
        br.n = 0;
        if (br.wp > br.wend + 1) { br.err = 1; return q; }
    }
}
// Rice code with parameter k (< 32): fast path when the whole code sits in the buffer.
template <bool CRC>
FA_D int32_t brc_rice(BitRdC& br, int k, const uint16_t* T) {
    brc_refill<CRC>(br, T);
    uint32_t q, low;
    int z = clz64(br.buf);
    if (br.buf != 0 && z + 1 + k <= br.n) {
        q = (uint32_t)z;
        uint64_t t = (br.buf << z) << 1;
        low = k ? (uint32_t)(t >> (64 - k)) : 0u;
        br.buf = k ? (t << k) : t;
        br.n -= z + 1 + k;
    } else {
        q = brc_unary<CRC>(br, T);
        low = brc_read<CRC>(br, k, T);
    }
    uint32_t u = (q << k) | low;
    return (int32_t)(u >> 1) ^ -(int32_t)(u & 1);
}

// Bits consumed since brc_init, and the frame-relative end / CRC check.
FA_D int64_t brc_pos_bits(const BitRdC& br, int a) { return (int64_t)(br.nwords + 1) * 32 - br.n - 8 * a; }

struct TileLane {
    int mode;       // 0 const, 1 verbatim, 2 predictive, -1 idle
    int lpc;        // predictive: 1 = LPC subframe (parameters follow the warm-up), 0 = fixed predictor
    int order, shift, wasted, bps;
    int raw_left;
    int32_t cval;
    int plen, esc, psize, part, nparts, left, k, rawbits;
    int32_t h[kTileOrd];
    int32_t c[kTileOrd];
};

// Parse one subframe header (up to and including the residual header).  Returns false if this path
// cannot decode the subframe (caller flags the frame for the general decoder) or the stream is bad.
template <bool CRC>
FA_D bool tile_subframe_begin(BitRdC& br, const uint16_t* T, int bs, int bps, TileLane& L) {
    if (brc_read<CRC>(br, 1, T) != 0) return false;
    int type = (int)brc_read<CRC>(br, 6, T);
    int wasted = 0;
    if (brc_read<CRC>(br, 1, T)) wasted = (int)brc_unary<CRC>(br, T) + 1;
    bps -= wasted;
    if (bps <= 0 || bps > 32 || br.err) return false;
    L.wasted = wasted;
    L.bps = bps;
    L.order = 0;
    L.lpc = 0;
    L.shift = 0;
    L.left = 0;
    L.part = -1;
    L.k = 0;
#pragma unroll
    for (int j = 0; j < kTileOrd; ++j) { L.h[j] = 0; L.c[j] = 0; }
    if (type == 0) {
        L.mode = 0;
        L.raw_left = 0;
        L.cval = brc_read_signed<CRC>(br, bps, T);
        return true;
    }
    if (type == 1) {
        L.mode = 1;
        L.raw_left = bs;
        return true;
    }
    bool lpc = type >= 32;
    int order;
    if (lpc) order = (type & 31) + 1;
    else if (type >= 8 && type <= 12) order = type - 8;
    else return false;
    if (order > kTileOrd || order > bs) return false;
    L.mode = 2;
    L.lpc = lpc ? 1 : 0;
    L.order = order;
    L.raw_left = order;
    return true;
}

// After the warm-up samples: LPC parameters + residual header.
template <bool CRC>
FA_D bool tile_subframe_params(BitRdC& br, const uint16_t* T, int bs, TileLane& L) {
    const int order = L.order;
    if (L.lpc) {
        int prec = (int)brc_read<CRC>(br, 4, T) + 1;
        if (prec == 16) return false;
        int sh = (int)brc_read<CRC>(br, 5, T);
        if (sh & 16) return false;  // negative shift
        L.shift = sh;
#pragma unroll
        for (int j = 0; j < kTileOrd; ++j)
            if (j < order) L.c[j] = brc_read_signed<CRC>(br, prec, T);
    } else {
        const int32_t fx[5][4] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, -1, 0, 0}, {3, -3, 1, 0}, {4, -6, 4, -1}};
#pragma unroll
        for (int j = 0; j < 4; ++j) L.c[j] = fx[order][j];
    }
    uint32_t method = brc_read<CRC>(br, 2, T);
    if (method > 1) return false;
    L.plen = method == 0 ? 4 : 5;
    L.esc = method == 0 ? 15 : 31;
    int porder = (int)brc_read<CRC>(br, 4, T);
    L.psize = bs >> porder;
    L.nparts = 1 << porder;
    if (porder > 0 && (L.psize << porder) != bs) return false;
    if (porder > 0 && L.psize < order) return false;
    L.part = -1;
    L.left = 0;
    return !br.err;
}

template <bool CRC>
FA_D int32_t tile_next_sample(BitRdC& br, const uint16_t* T, TileLane& L) {
    int32_t v;
    if (L.mode == 0) {
        v = L.cval;
    } else if (L.raw_left > 0) {
        v = brc_read_signed<CRC>(br, L.bps, T);
        L.raw_left--;
    } else {
        while (L.left == 0) {
            L.part++;
            if (L.part >= L.nparts) { br.err = 1; return 0; }
            L.left = L.psize - (L.part == 0 ? L.order : 0);
            int k = (int)brc_read<CRC>(br, L.plen, T);
            if (k == L.esc) { L.k = -1; L.rawbits = (int)brc_read<CRC>(br, 5, T); }
            else L.k = k;
        }
        L.left--;
        int32_t r = (L.k >= 0) ? brc_rice<CRC>(br, L.k, T) : brc_read_signed<CRC>(br, L.rawbits, T);
        int64_t sum = 0;
#pragma unroll
        for (int j = 0; j < kTileOrd; ++j) sum += (int64_t)L.c[j] * (int64_t)L.h[j];
        v = (int32_t)((int64_t)r + (sum >> L.shift));
    }
#pragma unroll
    for (int j = kTileOrd - 1; j > 0; --j) L.h[j] = L.h[j - 1];
    L.h[0] = v;
    return (int32_t)((uint32_t)v << L.wasted);
}

struct TileParams {
    DecParams D;
    int64_t j0, nwin;        // frame window (for the hinted blocksize)
    unsigned char* frame_flag;  // [n_sel][nframes_cap]: 1 = leave to the general decoder
    int restore;             // 1: nch == 1 and D.data receives float32 (offsets/gains given)
    const float* offsets;
    const float* gains;
};

// One warp: 32 consecutive (stream, frame) work items.  `ws` = this warp's shared storage, `T` = CRC
// slice tables in shared memory.
template <bool CRC>
FA_D void tile_warp_body(const TileParams& P, int64_t item0, TileShared* ws, const uint16_t* T) {
    const DecParams& D = P.D;
    const int ln = lane();
    const int nch = D.nch;
    int64_t idx = item0 + ln;
    int64_t k = idx / P.nwin;
    bool active = k < D.n_sel;
    int64_t j = 0;
    int bs_nom = 0, bs = 0;
    const uint8_t* fp = nullptr;
    const uint8_t* end = nullptr;
    long long off = 0, next = -1;
    if (active) {
        if (D.stream_flag[k] != 0) active = false;
    }
    if (active) {
        const StreamMeta m = D.meta[k];
        bs_nom = m.blocksize;
        int64_t jj0 = D.first / bs_nom, jj1 = (D.first + D.n_decode - 1) / bs_nom;
        if (jj1 - jj0 + 1 > P.nwin) {
            if (idx % P.nwin == 0) atom_or_global(&D.stream_flag[k], 1);
            active = false;
        } else {
            j = jj0 + idx % P.nwin;
            if (j > jj1) active = false;
        }
    }
    FrameHdr fh;
    uint32_t crc0 = 0;
    if (active) {
        const long long* fo = D.frame_off + k * (int64_t)(D.nframes_cap + 1);
        off = fo[j];
        next = fo[j + 1];
        if (off < 0) { atom_or_global(&D.stream_flag[k], 2); active = false; }
    }
    if (active) {
        const uint8_t* buf = D.bytes + D.starts[k];
        end = buf + D.nbytes[k];
        fp = buf + off;
        bool ok = parse_frame_header(fp, end - fp, D.crc, fh);
        int bps = ok ? (fh.bps ? fh.bps : D.meta[k].bps) : 0;
        if (!ok || fh.nch != nch || bps != 32 || fh.variable || (int64_t)fh.number != j) {
            atom_or_global(&D.stream_flag[k], 2);
            active = false;
        } else if (fh.ca >= 8) {
            P.frame_flag[k * (int64_t)D.nframes_cap + j] = 1;  // side channel: general decoder
            active = false;
        } else {
            bs = fh.blocksize;
            if (CRC) for (int b = 0; b < fh.hdr_bytes; ++b) crc0 = crc16_b(T, crc0, fp[b]);
        }
    }
    // publish the row description
    {
        TileRow r;
        r.out = nullptr; r.lo = 0; r.hi = 0; r.off = 0.f; r.coeff = 0.f;
        if (active) {
            int64_t s0 = j * (int64_t)bs_nom;
            r.out = D.data + (k * D.n_decode + (s0 - D.first)) * nch;
            int64_t lo = D.first - s0, hi = D.first + D.n_decode - s0;
            r.lo = (int)(lo < 0 ? 0 : lo);
            r.hi = (int)(hi > bs ? bs : hi);
            if (P.restore) { r.off = P.offsets[k]; r.coeff = restore_coeff_f32(P.gains[k]); }
        }
        ws->row[ln] = r;
    }
    syncwarp();
    uint32_t bsmax = (uint32_t)(active ? bs : 0);
    for (int m = 16; m >= 1; m >>= 1) { uint32_t o = shfl_xor(bsmax, m); bsmax = o > bsmax ? o : bsmax; }

    BitRdC br;
    br.err = 0;
    int a = 0;
    if (active) {
        const uint8_t* body = fp + fh.hdr_bytes;
        a = (int)((uintptr_t)body & 3);
        brc_init<CRC>(br, body, end, crc0, T);
    } else {
        br.wp = br.wend = nullptr; br.buf = 0; br.n = 64; br.nwords = 0; br.crc = 0; br.w1 = br.w2 = 0;
    }
    bool fail = false;      // stream problem -> walker
    bool punt = false;      // unsupported subframe -> general decoder
    TileLane L;
    L.mode = -1;
    for (int c = 0; c < nch; ++c) {
        if (active && !fail && !punt) {
            if (!tile_subframe_begin<CRC>(br, T, bs, 32, L)) {
                if (br.err) fail = true; else punt = true;
            }
        }
        bool run = active && !fail && !punt;
        // The warm-up samples are read inside the main loop (raw_left); the LPC parameters and the
        // residual header follow them in the stream and are parsed when the warm-up is exhausted.
        bool need_params = run && L.mode == 2;
        for (uint32_t base = 0; base < bsmax; base += 32) {
            for (int s = 0; s < 32; ++s) {
                int i = (int)base + s;
                int32_t v = 0;
                if (run && i < bs) {
                    if (need_params && L.raw_left == 0) {
                        if (!tile_subframe_params<CRC>(br, T, bs, L)) { run = false; if (br.err) fail = true; else punt = true; }
                        need_params = false;
                    }
                    if (run) v = tile_next_sample<CRC>(br, T, L);
                }
                ws->tile[ln * kTileStride + s] = v;
            }
            syncwarp();
            // transpose: each iteration stores 32 consecutive samples of one frame (128 B)
            for (int r = 0; r < 32; ++r) {
                const TileRow row = ws->row[r];
                int i = (int)base + ln;
                if (row.out != nullptr && i >= row.lo && i < row.hi) {
                    int32_t v = ws->tile[r * kTileStride + ln];
                    if (P.restore) ((float*)row.out)[i] = restore_f32(v, row.off, row.coeff);
                    else row.out[(int64_t)i * nch + c] = v;
                }
            }
            syncwarp();
        }
        if (run && br.err) fail = true;
    }
    if (active) {
        if (punt && !fail) {
            P.frame_flag[k * (int64_t)D.nframes_cap + j] = 1;
        } else if (!fail) {
            // frame end: pad to a byte, CRC-16, chain check against the next frame's start
            int64_t bits = brc_pos_bits(br, a);            // relative to the body start
            int64_t body_bytes = (bits + 7) >> 3;
            int64_t len = fh.hdr_bytes + body_bytes + 2;
            if (fp + len > end) fail = true;
            if (!fail && next >= 0 && off + len != next) fail = true;
            if (!fail && CRC) {
                // bytes of the body counted from the first fetched word: E = a + body_bytes
                int64_t E = a + body_bytes;
                int64_t wE = E >> 2;                       // word holding the first CRC byte
                uint32_t c16 = br.crc;                     // state before word (nwords + 1 - 2) in absolute numbering
                // absolute word numbering: word 0 = first word (consumed byte-wise in brc_init);
                // fetched word t (t = 1..nwords) ; br.crc covers words < nwords - 1 (i.e. up to t = nwords - 2)
                int64_t covered = (br.nwords >= 2) ? (int64_t)br.nwords - 1 : 1;   // first uncovered absolute word
                // when fewer than two words were fetched nothing sits in the lag queue beyond w1
                uint32_t q[2] = {br.w2, br.w1};
                int qn = br.nwords >= 2 ? 2 : br.nwords;   // words in the queue: absolute indices covered .. covered+qn-1
                if (br.nwords == 1) q[0] = br.w1;
                for (int t = 0; t < qn; ++t) {
                    int64_t wabs = covered + t;
                    if (wabs > wE) break;
                    int nb = wabs < wE ? 4 : (int)(E & 3);
                    for (int b = 0; b < nb; ++b) c16 = crc16_b(T, c16, (q[t] >> (24 - 8 * b)) & 0xFF);
                }
                uint32_t want = ((uint32_t)fp[len - 2] << 8) | fp[len - 1];
                if (c16 != want) fail = true;
            }
        }
        if (fail) atom_or_global(&D.stream_flag[k], 2);
    }
}

}  // namespace fa
