/*
 * flac_oracle.c -- CPU restatement of the flacarray hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for the B200 CUDA path.  It is NOT part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  The product library (flacarray_b200/csrc) never links or calls it.
 *
 * What it restates (file:line are relative to /root/reference/src/flacarray/libflacarray/):
 *   - orc_encode_i32/i64      : compress.c:133-270 (serial) / :274-435 (OpenMP over streams),
 *                               compress.c:440-540 (i32/i64 helpers, 1 or 2 channels)
 *   - orc_decode_i32/i64      : decompress.c:194-313 (range checks, per-stream loop, slice
 *                               window), decompress.c:66-101 (interleave + clip to n_decode),
 *                               decompress.c:318-375 (helpers)
 *   - int64 <-> 2 x int32     : utils.c:96-125 (little-endian reinterpret: ch0 = low word
 *                               as signed int32, ch1 = high word)
 *   - orc_float32_to_int32 &c : utils.c:160-368 (same operation order and intermediate types)
 *
 * The codec arithmetic itself lives in libFLAC (xiph/flac; the reference requires >= 1.4.0,
 * meson.build:13, and its wheels pin 1.5.0, packaging/wheels/install_deps_linux.sh:53).
 * libFLAC is NOT in /root/reference and is not installed in this image, so the codec below
 * restates the published bit-stream (RFC 9639) for the decoder and libFLAC's published
 * encoder procedure (stream_encoder.c process_subframe_/find_best_partition_order_, lpc.c,
 * fixed.c, window.c of 1.4/1.5) for the encoder.
 *
 * PARITY PIN: the reference's own tests hold no compressed bytes (round-trip properties only),
 * and no libFLAC exists here, so byte/size parity against libFLAC itself is "parity unpinned".
 * What IS pinned (tests/test_oracle_*.py):
 *   - the decoder against FFmpeg 8.0's independent native FLAC encoder (32 bps, mono, and all
 *     four stereo assignments incl. the 33-bit side channel) and a hand-assembled VERBATIM frame;
 *   - the encoder against FFmpeg's independent FLAC decoder (bit-exact samples);
 *   - the float<->int converters against the reference's own utils.c compiled into oracle/_ref.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <stdbool.h>

/* Error codes: flacarray.h:20-40 */
#define ERROR_NONE 0
#define ERROR_ALLOC (1 << 0)
#define ERROR_INVALID_LEVEL (1 << 1)
#define ERROR_ZERO_NSTREAM (1 << 2)
#define ERROR_ZERO_STREAMSIZE (1 << 3)
#define ERROR_ENCODE_PROCESS (1 << 9)
#define ERROR_ENCODE_COLLECT (1 << 11)
#define ERROR_DECODE_INIT (1 << 13)
#define ERROR_DECODE_PROCESS (1 << 14)
#define ERROR_DECODE_SAMPLE_RANGE (1 << 17)
#define ERROR_DECODE_SEEK (1 << 18)

#define MAX_LPC_ORDER 32
#define MAX_FIXED_ORDER 4
#define MAX_BLOCKSIZE 65535

/* ------------------------------------------------------------------------------------------ */
/* CRC-8 (poly 0x07) and CRC-16 (poly 0x8005), both MSB-first, init 0 (RFC 9639 9.1.8, 9.3)    */
/* ------------------------------------------------------------------------------------------ */
static uint8_t crc8_tab[256];
static uint16_t crc16_tab[256];
static int crc_ready = 0;

static void crc_init(void) {
    if (crc_ready) return;
    for (int i = 0; i < 256; ++i) {
        uint8_t c = (uint8_t)i;
        for (int b = 0; b < 8; ++b) c = (c & 0x80) ? (uint8_t)((c << 1) ^ 0x07) : (uint8_t)(c << 1);
        crc8_tab[i] = c;
        uint16_t d = (uint16_t)(i << 8);
        for (int b = 0; b < 8; ++b) d = (d & 0x8000) ? (uint16_t)((d << 1) ^ 0x8005) : (uint16_t)(d << 1);
        crc16_tab[i] = d;
    }
    crc_ready = 1;
}

uint8_t orc_crc8(const uint8_t *p, int64_t n) {
    crc_init();
    uint8_t c = 0;
    for (int64_t i = 0; i < n; ++i) c = crc8_tab[c ^ p[i]];
    return c;
}

uint16_t orc_crc16(const uint8_t *p, int64_t n) {
    crc_init();
    uint16_t c = 0;
    for (int64_t i = 0; i < n; ++i) c = (uint16_t)((c << 8) ^ crc16_tab[(c >> 8) ^ p[i]]);
    return c;
}

/* ------------------------------------------------------------------------------------------ */
/* Bit reader, MSB first                                                                       */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    const uint8_t *p;
    int64_t nbits; /* total bits available */
    int64_t pos;   /* bit position */
    int err;
} BitReader;

static inline uint32_t br_bit(BitReader *br) {
    if (br->pos >= br->nbits) { br->err = 1; return 0; }
    uint32_t b = (br->p[br->pos >> 3] >> (7 - (br->pos & 7))) & 1u;
    br->pos++;
    return b;
}

/* read n <= 57 bits unsigned */
static inline uint64_t br_read(BitReader *br, int n) {
    if (n == 0) return 0;
    if (br->pos + n > br->nbits) { br->err = 1; br->pos = br->nbits; return 0; }
    uint64_t v = 0;
    int64_t pos = br->pos;
    int left = n;
    while (left > 0) {
        int off = (int)(pos & 7);
        int take = 8 - off;
        if (take > left) take = left;
        uint32_t byte = br->p[pos >> 3];
        uint32_t bits = (byte >> (8 - off - take)) & ((1u << take) - 1u);
        v = (v << take) | bits;
        pos += take;
        left -= take;
    }
    br->pos = pos;
    return v;
}

static inline int64_t br_read_signed(BitReader *br, int n) {
    if (n == 0) return 0;
    uint64_t v = br_read(br, n);
    uint64_t sign = 1ull << (n - 1);
    return (int64_t)((v ^ sign) - sign);
}

static inline uint32_t br_unary(BitReader *br) {
    /* count zeros before the terminating one */
    uint32_t q = 0;
    for (;;) {
        if (br->pos >= br->nbits) { br->err = 1; return q; }
        int off = (int)(br->pos & 7);
        uint32_t byte = (uint32_t)(br->p[br->pos >> 3] << off) & 0xFFu; /* remaining bits at top */
        if (byte == 0) { q += 8 - off; br->pos += 8 - off; continue; }
        int lz = __builtin_clz(byte) - 24;
        q += lz;
        br->pos += lz + 1;
        if (br->pos > br->nbits) br->err = 1;
        return q;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Decoder (RFC 9639)                                                                          */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int blocksize;
    int channel_assignment; /* raw 4-bit field */
    int n_channels;
    int bps;
    uint64_t number; /* frame number (fixed) or sample number (variable) */
    int variable;
    int header_bytes;
} FrameHeader;

typedef struct {
    int min_blocksize, max_blocksize;
    int sample_rate, channels, bps;
    uint64_t total_samples;
    int64_t first_frame_byte;
} StreamInfo;

/* Parse the metadata chain.  Returns 0 on success. */
static int parse_metadata(const uint8_t *buf, int64_t nbytes, StreamInfo *si) {
    if (nbytes < 8 || memcmp(buf, "fLaC", 4) != 0) return -1;
    int64_t pos = 4;
    int have_si = 0;
    for (;;) {
        if (pos + 4 > nbytes) return -1;
        int last = buf[pos] >> 7;
        int type = buf[pos] & 0x7F;
        int64_t len = ((int64_t)buf[pos + 1] << 16) | ((int64_t)buf[pos + 2] << 8) | buf[pos + 3];
        pos += 4;
        if (pos + len > nbytes) return -1;
        if (type == 0) {
            if (len < 34) return -1;
            const uint8_t *s = buf + pos;
            si->min_blocksize = (s[0] << 8) | s[1];
            si->max_blocksize = (s[2] << 8) | s[3];
            si->sample_rate = (s[10] << 12) | (s[11] << 4) | (s[12] >> 4);
            si->channels = ((s[12] >> 1) & 7) + 1;
            si->bps = (((s[12] & 1) << 4) | (s[13] >> 4)) + 1;
            si->total_samples = ((uint64_t)(s[13] & 0xF) << 32) | ((uint64_t)s[14] << 24) |
                                ((uint64_t)s[15] << 16) | ((uint64_t)s[16] << 8) | s[17];
            have_si = 1;
        }
        pos += len;
        if (last) break;
    }
    if (!have_si) return -1;
    si->first_frame_byte = pos;
    return 0;
}

static int parse_frame_header(const uint8_t *p, int64_t avail, const StreamInfo *si, FrameHeader *fh) {
    crc_init();
    if (avail < 6) return -1;
    if (p[0] != 0xFF || (p[1] & 0xFE) != 0xF8) return -1; /* sync + reserved 0 */
    fh->variable = p[1] & 1;
    int bs_code = p[2] >> 4;
    int sr_code = p[2] & 0xF;
    int ch_code = p[3] >> 4;
    int ss_code = (p[3] >> 1) & 7;
    if (p[3] & 1) return -1; /* reserved */
    if (bs_code == 0 || sr_code == 15) return -1;
    if (ch_code > 10) return -1;
    if (ss_code == 3) return -1;
    int pos = 4;
    /* UTF-8 style coded number */
    uint32_t b0 = p[pos++];
    uint64_t v;
    int extra;
    if (!(b0 & 0x80)) { v = b0; extra = 0; }
    else if ((b0 & 0xE0) == 0xC0) { v = b0 & 0x1F; extra = 1; }
    else if ((b0 & 0xF0) == 0xE0) { v = b0 & 0x0F; extra = 2; }
    else if ((b0 & 0xF8) == 0xF0) { v = b0 & 0x07; extra = 3; }
    else if ((b0 & 0xFC) == 0xF8) { v = b0 & 0x03; extra = 4; }
    else if ((b0 & 0xFE) == 0xFC) { v = b0 & 0x01; extra = 5; }
    else if (b0 == 0xFE && fh->variable) { v = 0; extra = 6; }
    else return -1;
    if (pos + extra + 1 > avail) return -1;
    for (int i = 0; i < extra; ++i) {
        uint32_t b = p[pos++];
        if ((b & 0xC0) != 0x80) return -1;
        v = (v << 6) | (b & 0x3F);
    }
    fh->number = v;
    int bs;
    if (bs_code == 1) bs = 192;
    else if (bs_code >= 2 && bs_code <= 5) bs = 576 << (bs_code - 2);
    else if (bs_code == 6) { if (pos + 1 > avail) return -1; bs = p[pos++] + 1; }
    else if (bs_code == 7) { if (pos + 2 > avail) return -1; bs = ((p[pos] << 8) | p[pos + 1]) + 1; pos += 2; }
    else bs = 256 << (bs_code - 8);
    if (sr_code == 12) pos += 1;
    else if (sr_code == 13 || sr_code == 14) pos += 2;
    if (pos + 1 > avail) return -1;
    uint8_t c = 0;
    for (int i = 0; i < pos; ++i) c = crc8_tab[c ^ p[i]];
    if (c != p[pos]) return -1;
    pos++;
    fh->blocksize = bs;
    fh->channel_assignment = ch_code;
    fh->n_channels = (ch_code < 8) ? ch_code + 1 : 2;
    static const int ss_tab[8] = {0, 8, 12, 0, 16, 20, 24, 32};
    fh->bps = ss_code == 0 ? si->bps : ss_tab[ss_code];
    fh->header_bytes = pos;
    return 0;
}

/* Decode the residual section into res[order..bs).  Returns 0 on success. */
static int decode_residual(BitReader *br, int bs, int order, int64_t *res) {
    int method = (int)br_read(br, 2);
    if (method > 1) return -1;
    int plen = method == 0 ? 4 : 5;
    uint32_t esc = method == 0 ? 15 : 31;
    int porder = (int)br_read(br, 4);
    int nparts = 1 << porder;
    if ((bs >> porder) << porder != bs && porder > 0) return -1;
    int i = order;
    for (int part = 0; part < nparts; ++part) {
        int n = (porder == 0) ? bs - order : (part == 0 ? (bs >> porder) - order : (bs >> porder));
        if (n < 0) return -1;
        uint32_t k = (uint32_t)br_read(br, plen);
        if (k == esc) {
            int raw = (int)br_read(br, 5);
            for (int j = 0; j < n; ++j) res[i++] = br_read_signed(br, raw);
        } else {
            for (int j = 0; j < n; ++j) {
                uint32_t q = br_unary(br);
                uint64_t low = br_read(br, (int)k);
                uint64_t u = ((uint64_t)q << k) | low;
                res[i++] = (int64_t)(u >> 1) ^ -(int64_t)(u & 1);
            }
        }
        if (br->err) return -1;
    }
    return br->err ? -1 : 0;
}

/* Decode one subframe of `bps` bits into out[0..bs) (int64 because the side channel is 33 bit). */
static int decode_subframe(BitReader *br, int bs, int bps, int64_t *out) {
    if (br_bit(br) != 0) return -1; /* padding */
    int type = (int)br_read(br, 6);
    int wasted = 0;
    if (br_bit(br)) wasted = (int)br_unary(br) + 1;
    if (br->err) return -1;
    bps -= wasted;
    if (bps <= 0) return -1;
    if (type == 0) {
        int64_t v = br_read_signed(br, bps);
        for (int i = 0; i < bs; ++i) out[i] = v;
    } else if (type == 1) {
        for (int i = 0; i < bs; ++i) out[i] = br_read_signed(br, bps);
    } else if (type >= 8 && type <= 12) {
        int order = type - 8;
        if (order > bs) return -1;
        for (int i = 0; i < order; ++i) out[i] = br_read_signed(br, bps);
        if (decode_residual(br, bs, order, out)) return -1;
        switch (order) {
        case 0: break;
        case 1: for (int i = 1; i < bs; ++i) out[i] += out[i - 1]; break;
        case 2: for (int i = 2; i < bs; ++i) out[i] += 2 * out[i - 1] - out[i - 2]; break;
        case 3: for (int i = 3; i < bs; ++i) out[i] += 3 * out[i - 1] - 3 * out[i - 2] + out[i - 3]; break;
        case 4: for (int i = 4; i < bs; ++i) out[i] += 4 * out[i - 1] - 6 * out[i - 2] + 4 * out[i - 3] - out[i - 4]; break;
        }
    } else if (type >= 32) {
        int order = (type & 31) + 1;
        if (order > bs) return -1;
        for (int i = 0; i < order; ++i) out[i] = br_read_signed(br, bps);
        int prec = (int)br_read(br, 4) + 1;
        if (prec == 16) return -1;
        int shift = (int)br_read_signed(br, 5);
        if (shift < 0) return -1;
        int32_t coef[MAX_LPC_ORDER];
        for (int j = 0; j < order; ++j) coef[j] = (int32_t)br_read_signed(br, prec);
        if (decode_residual(br, bs, order, out)) return -1;
        for (int i = order; i < bs; ++i) {
            int64_t sum = 0;
            for (int j = 0; j < order; ++j) sum += (int64_t)coef[j] * out[i - 1 - j];
            out[i] += sum >> shift;
        }
    } else {
        return -1; /* reserved */
    }
    if (br->err) return -1;
    if (wasted)
        for (int i = 0; i < bs; ++i) out[i] = (int64_t)((uint64_t)out[i] << wasted);
    return 0;
}

/*
 * Decode one frame starting at p.  ch[c][0..bs) receives the channel samples (after stereo undo).
 * Returns frame length in bytes, or -1.
 */
static int64_t decode_frame(const uint8_t *p, int64_t avail, const StreamInfo *si, FrameHeader *fh,
                            int64_t *ch0, int64_t *ch1) {
    if (parse_frame_header(p, avail, si, fh)) return -1;
    if (fh->n_channels > 2) return -1; /* flacarray only writes 1 or 2 channels */
    BitReader br = {p, avail * 8, (int64_t)fh->header_bytes * 8, 0};
    int bs = fh->blocksize;
    int ca = fh->channel_assignment;
    int bps0 = fh->bps + (ca == 9 ? 1 : 0);
    int bps1 = fh->bps + ((ca == 8 || ca == 10) ? 1 : 0);
    if (decode_subframe(&br, bs, bps0, ch0)) return -1;
    if (fh->n_channels == 2 && decode_subframe(&br, bs, bps1, ch1)) return -1;
    if (ca == 8) { /* left/side */
        for (int i = 0; i < bs; ++i) ch1[i] = ch0[i] - ch1[i];
    } else if (ca == 9) { /* side/right */
        for (int i = 0; i < bs; ++i) ch0[i] = ch0[i] + ch1[i];
    } else if (ca == 10) { /* mid/side */
        for (int i = 0; i < bs; ++i) {
            int64_t m = ch0[i], s = ch1[i];
            m = (int64_t)((uint64_t)m << 1) | (s & 1);
            ch0[i] = (m + s) >> 1;
            ch1[i] = (m - s) >> 1;
        }
    }
    /* byte align, CRC-16 */
    int64_t end = (br.pos + 7) >> 3;
    if (end + 2 > avail) return -1;
    uint16_t want = (uint16_t)((p[end] << 8) | p[end + 1]);
    if (orc_crc16(p, end) != want) return -1;
    return end + 2;
}

/*
 * Decode samples [first, first + n_decode) of one stream into interleaved int32 `out`
 * (decompress.c:66-101 interleave/clip semantics).  Returns an ERROR_* bitmask.
 */
int orc_decode_stream(const uint8_t *buf, int64_t nbytes, int64_t stream_size, int n_channels,
                      int64_t first, int64_t n_decode, int32_t *out) {
    StreamInfo si;
    if (parse_metadata(buf, nbytes, &si)) return ERROR_DECODE_INIT;
    if (si.channels != n_channels) return ERROR_DECODE_PROCESS;
    int64_t *ch0 = (int64_t *)malloc(sizeof(int64_t) * 2 * (MAX_BLOCKSIZE + 1));
    if (!ch0) return ERROR_ALLOC;
    int64_t *ch1 = ch0 + MAX_BLOCKSIZE + 1;
    int64_t pos = si.first_frame_byte;
    int64_t sample = 0; /* index of first sample of the current frame */
    int64_t last = first + n_decode;
    int err = ERROR_NONE;
    uint64_t expect_frame = 0;
    while (sample < last) {
        if (pos >= nbytes) { err = ERROR_DECODE_PROCESS; break; }
        FrameHeader fh;
        int64_t flen = decode_frame(buf + pos, nbytes - pos, &si, &fh, ch0, ch1);
        if (flen < 0 || fh.n_channels != n_channels || fh.bps != 32) { err = ERROR_DECODE_PROCESS; break; }
        if (!fh.variable) {
            if (fh.number != expect_frame) { err = ERROR_DECODE_PROCESS; break; }
        } else if ((int64_t)fh.number != sample) { err = ERROR_DECODE_PROCESS; break; }
        expect_frame++;
        int64_t lo = sample < first ? first : sample;
        int64_t hi = sample + fh.blocksize < last ? sample + fh.blocksize : last;
        for (int64_t s = lo; s < hi; ++s) {
            out[(s - first) * n_channels] = (int32_t)ch0[s - sample];
            if (n_channels == 2) out[(s - first) * n_channels + 1] = (int32_t)ch1[s - sample];
        }
        sample += fh.blocksize;
        pos += flen;
    }
    free(ch0);
    return err;
}

/* List frame byte offsets of a stream (test helper).  Returns count or -1. */
int64_t orc_index_frames(const uint8_t *buf, int64_t nbytes, int64_t *offsets, int32_t *blocksizes,
                         int32_t *assignments, int64_t cap) {
    StreamInfo si;
    if (parse_metadata(buf, nbytes, &si)) return -1;
    int64_t *ch0 = (int64_t *)malloc(sizeof(int64_t) * 2 * (MAX_BLOCKSIZE + 1));
    int64_t *ch1 = ch0 + MAX_BLOCKSIZE + 1;
    int64_t pos = si.first_frame_byte, n = 0;
    while (pos < nbytes) {
        FrameHeader fh;
        int64_t flen = decode_frame(buf + pos, nbytes - pos, &si, &fh, ch0, ch1);
        if (flen < 0) { free(ch0); return -1; }
        if (n < cap) { offsets[n] = pos; blocksizes[n] = fh.blocksize; assignments[n] = fh.channel_assignment; }
        n++;
        pos += flen;
    }
    free(ch0);
    return n;
}

/* decompress.c:194-313 */
static int orc_decode(const uint8_t *bytes, const int64_t *starts, const int64_t *nbytes, int64_t n_stream,
                      int64_t stream_size, int n_channels, int64_t first_sample, int64_t last_sample,
                      int32_t *data, int use_threads) {
    int64_t first_decode = 0;
    int64_t n_decode = stream_size;
    if (first_sample >= 0 && last_sample >= 0) {
        if (last_sample > stream_size) return ERROR_DECODE_SAMPLE_RANGE;
        if (first_sample > stream_size - 1) return ERROR_DECODE_SAMPLE_RANGE;
        if (first_sample >= last_sample) return ERROR_DECODE_SAMPLE_RANGE;
        first_decode = first_sample;
        n_decode = last_sample - first_sample;
    }
    int errors = ERROR_NONE;
#pragma omp parallel for schedule(static) reduction(| : errors) if (use_threads)
    for (int64_t i = 0; i < n_stream; ++i) {
        errors |= orc_decode_stream(bytes + starts[i], nbytes[i], stream_size, n_channels, first_decode,
                                    n_decode, data + i * n_decode * n_channels);
    }
    return errors;
}

int orc_decode_i32(const uint8_t *bytes, const int64_t *starts, const int64_t *nbytes, int64_t n_stream,
                   int64_t stream_size, int64_t first_sample, int64_t last_sample, int32_t *data,
                   int use_threads) {
    return orc_decode(bytes, starts, nbytes, n_stream, stream_size, 1, first_sample, last_sample, data,
                      use_threads);
}

/* decompress.c:343-375, little-endian branch of utils.c:112-116 */
int orc_decode_i64(const uint8_t *bytes, const int64_t *starts, const int64_t *nbytes, int64_t n_stream,
                   int64_t stream_size, int64_t first_sample, int64_t last_sample, int64_t *data,
                   int use_threads) {
    return orc_decode(bytes, starts, nbytes, n_stream, stream_size, 2, first_sample, last_sample,
                      (int32_t *)data, use_threads);
}

/* ------------------------------------------------------------------------------------------ */
/* Bit writer, MSB first                                                                       */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    uint8_t *p;
    int64_t cap;  /* bytes */
    int64_t pos;  /* bits */
    int err;
} BitWriter;

static inline void bw_write(BitWriter *bw, uint64_t v, int n) {
    /* n <= 64 */
    if (n == 0) return;
    if (((bw->pos + n + 7) >> 3) > bw->cap) { bw->err = 1; return; }
    for (int i = n - 1; i >= 0; --i) {
        if ((v >> i) & 1u) bw->p[bw->pos >> 3] |= (uint8_t)(0x80u >> (bw->pos & 7));
        bw->pos++;
    }
}

static inline void bw_zeros(BitWriter *bw, int64_t n) {
    if (((bw->pos + n + 7) >> 3) > bw->cap) { bw->err = 1; return; }
    bw->pos += n; /* buffer is pre-zeroed */
}

static inline void bw_rice(BitWriter *bw, int64_t r, int k) {
    uint64_t u = ((uint64_t)r << 1) ^ (uint64_t)(r >> 63);
    uint64_t q = u >> k;
    bw_zeros(bw, (int64_t)q);
    bw_write(bw, 1, 1);
    if (k) bw_write(bw, u & ((1ull << k) - 1), k);
}

/* ------------------------------------------------------------------------------------------ */
/* Encoder: libFLAC presets (stream_encoder.c compression_levels_[]) -- SURVEY App. B          */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int do_mid_side;
    int max_lpc_order;
    int blocksize;
    int max_partition_order;
    int n_apod; /* number of apodization windows tried */
} Preset;

static const Preset presets[9] = {
    {0, 0, 1152, 3, 1}, {1, 0, 1152, 3, 1}, {1, 0, 1152, 3, 1}, {0, 6, 4096, 4, 1}, {1, 8, 4096, 4, 1},
    {1, 8, 4096, 5, 1}, {1, 8, 4096, 6, 2}, {1, 12, 4096, 6, 2}, {1, 12, 4096, 6, 3},
};

/* A chosen subframe encoding */
typedef struct {
    int type;  /* 0 constant, 1 verbatim, 2 fixed, 3 lpc */
    int order;
    int wasted;
    int bps; /* after wasted bits removed */
    int qlp_precision, qlp_shift;
    int32_t qlp[MAX_LPC_ORDER];
    int porder;
    int rice2;
    int params[1 << 8];
    uint64_t bits; /* estimate incl. subframe header */
} SubframePlan;

static int max_porder_from_blocksize(int bs) {
    int p = 0;
    while (!(bs & 1)) { p++; bs >>= 1; }
    return p > 15 ? 15 : p;
}

/*
 * libFLAC find_best_partition_order_/set_partitioned_rice_ (estimate-based, no escape codes).
 * res[0..bs) with res[0..order) unused.  Returns estimated residual bits (incl. 6 header bits).
 */
static uint64_t rice_search(const int64_t *res, int bs, int order, int max_porder_level, int *best_porder,
                            int *params, int *rice2) {
    int maxp = max_porder_from_blocksize(bs);
    if (maxp > max_porder_level) maxp = max_porder_level;
    while (maxp > 0 && (bs >> maxp) <= order) maxp--;
    int nfin = 1 << maxp;
    uint64_t sums[9][256];
    /* finest level sums */
    {
        int psize = bs >> maxp;
        int i = order;
        for (int part = 0; part < nfin; ++part) {
            int end = (part + 1) * psize;
            uint64_t s = 0;
            for (; i < end; ++i) s += (uint64_t)(res[i] < 0 ? -res[i] : res[i]);
            sums[maxp][part] = s;
        }
    }
    for (int p = maxp - 1; p >= 0; --p)
        for (int part = 0; part < (1 << p); ++part) sums[p][part] = sums[p + 1][2 * part] + sums[p + 1][2 * part + 1];
    uint64_t best_bits = UINT64_MAX;
    int tmp[256];
    for (int p = maxp; p >= 0; --p) {
        uint64_t bits = 6;
        for (int part = 0; part < (1 << p); ++part) {
            uint64_t n = (uint64_t)(bs >> p) - (part == 0 ? (uint64_t)order : 0);
            uint64_t mean = sums[p][part];
            int k = 0;
            uint64_t kk = n;
            while (kk < mean) { k++; kk <<= 1; }
            if (k >= 31) k = 30; /* rice_parameter_limit - 1 for > 16 bps */
            tmp[part] = k;
            uint64_t pb = 4 + (uint64_t)(1 + k) * n + (k ? (mean >> (k - 1)) : (mean << 1)) - (n >> 1);
            bits += pb;
        }
        if (bits < best_bits) {
            best_bits = bits;
            *best_porder = p;
            memcpy(params, tmp, sizeof(int) * (size_t)(1 << p));
        }
    }
    *rice2 = 0;
    for (int part = 0; part < (1 << *best_porder); ++part)
        if (params[part] >= 15) *rice2 = 1;
    return best_bits;
}

static int fits_i32(int64_t v) { return v >= -2147483647LL && v <= 2147483647LL; } /* INT32_MIN not codable */

/* window.c FLAC__window_tukey(p = 0.5) for the nominal blocksize; short last frames use a prefix */
static void make_tukey(float *w, int L) {
    for (int n = 0; n < L; ++n) w[n] = 1.0f;
    int Np = (int)(0.5f / 2.0f * L) - 1;
    if (Np > 0) {
        for (int n = 0; n <= Np; ++n) {
            w[n] = (float)(0.5f - 0.5f * cosf((float)(M_PI * n / Np)));
            w[L - Np - 1 + n] = (float)(0.5f - 0.5f * cosf((float)(M_PI * (n + Np) / Np)));
        }
    }
}

/* partial tukey / punchout tukey of libFLAC are approximated by one extra tukey for levels 6-8:
 * the oracle only needs a faithful size model at the default level (5) and a valid stream at all. */

static double expected_bits_per_sample(double lpc_error, double error_scale) {
    if (lpc_error > 0.0) {
        double bps = 0.5 * log(error_scale * lpc_error) / M_LN2;
        return bps >= 0.0 ? bps : 0.0;
    } else if (lpc_error < 0.0) {
        return 1e32;
    }
    return 0.0;
}

/*
 * Plan one subframe for signal x[0..bs) of `bps` bits (stream_encoder.c process_subframe_).
 * res_out receives the residual of the chosen predictor.
 */
static void plan_subframe(const int64_t *xin, int bs, int bps_in, const Preset *ps, const float *window,
                          SubframePlan *pl, int64_t *res_out, int64_t *work) {
    int64_t *x = work;            /* bs */
    int64_t *res = work + bs;     /* bs */
    memset(pl, 0, sizeof(*pl));
    /* wasted bits */
    uint64_t orv = 0;
    for (int i = 0; i < bs; ++i) orv |= (uint64_t)xin[i];
    int wasted = 0;
    if (orv != 0) wasted = __builtin_ctzll(orv);
    if (wasted >= bps_in) wasted = 0;
    int bps = bps_in - wasted;
    for (int i = 0; i < bs; ++i) x[i] = xin[i] >> wasted;
    pl->wasted = wasted;
    pl->bps = bps;
    uint64_t hdr = 8 + (wasted ? (uint64_t)wasted : 0);
    /* verbatim */
    pl->type = 1;
    pl->bits = hdr + (uint64_t)bps * bs;
    /* constant */
    int constant = 1;
    for (int i = 1; i < bs; ++i) if (x[i] != x[0]) { constant = 0; break; }
    if (constant) {
        pl->type = 0;
        pl->bits = hdr + bps;
        return;
    }
    if (bs <= MAX_FIXED_ORDER) return;
    /* fixed: FLAC__fixed_compute_best_predictor_wide / _limit_residual */
    {
        uint64_t te[5] = {0, 0, 0, 0, 0};
        int ok[5] = {1, 1, 1, 1, 1};
        for (int i = MAX_FIXED_ORDER; i < bs; ++i) {
            int64_t e0 = x[i];
            int64_t e1 = x[i] - x[i - 1];
            int64_t e2 = x[i] - 2 * x[i - 1] + x[i - 2];
            int64_t e3 = x[i] - 3 * x[i - 1] + 3 * x[i - 2] - x[i - 3];
            int64_t e4 = x[i] - 4 * x[i - 1] + 6 * x[i - 2] - 4 * x[i - 3] + x[i - 4];
            int64_t e[5] = {e0, e1, e2, e3, e4};
            for (int o = 0; o < 5; ++o) {
                if (!fits_i32(e[o])) ok[o] = 0;
                te[o] += (uint64_t)(e[o] < 0 ? -e[o] : e[o]);
            }
        }
        for (int o = 0; o < 5; ++o) if (!ok[o]) te[o] = UINT64_MAX;
        int order;
        uint64_t m1234 = te[1] < te[2] ? te[1] : te[2];
        uint64_t m34 = te[3] < te[4] ? te[3] : te[4];
        if (m34 < m1234) m1234 = m34;
        uint64_t m234 = te[2] < m34 ? te[2] : m34;
        if (te[0] < m1234) order = 0;
        else if (te[1] < m234) order = 1;
        else if (te[2] < m34) order = 2;
        else if (te[3] < te[4]) order = 3;
        else order = 4;
        if (te[order] != UINT64_MAX) {
            double n = (double)(bs - MAX_FIXED_ORDER);
            double rbps = te[order] > 0 ? log(M_LN2 * (double)te[order] / n) / M_LN2 : 0.0;
            if (rbps < (double)bps) {
                /* residual over the whole block for the chosen order */
                int good = 1;
                for (int i = order; i < bs; ++i) {
                    int64_t e;
                    switch (order) {
                    case 0: e = x[i]; break;
                    case 1: e = x[i] - x[i - 1]; break;
                    case 2: e = x[i] - 2 * x[i - 1] + x[i - 2]; break;
                    case 3: e = x[i] - 3 * x[i - 1] + 3 * x[i - 2] - x[i - 3]; break;
                    default: e = x[i] - 4 * x[i - 1] + 6 * x[i - 2] - 4 * x[i - 3] + x[i - 4]; break;
                    }
                    if (!fits_i32(e)) { good = 0; break; }
                    res[i] = e;
                }
                if (good) {
                    int porder, rice2, params[256];
                    uint64_t rb = rice_search(res, bs, order, ps->max_partition_order, &porder, params, &rice2);
                    uint64_t bits = hdr + (uint64_t)order * bps + rb;
                    if (bits < pl->bits) {
                        pl->type = 2; pl->order = order; pl->porder = porder; pl->rice2 = rice2;
                        memcpy(pl->params, params, sizeof(int) * (size_t)(1 << porder));
                        pl->bits = bits;
                        memcpy(res_out, res, sizeof(int64_t) * bs);
                    }
                }
            }
        }
    }
    /* LPC */
    if (ps->max_lpc_order > 0) {
        int max_order = ps->max_lpc_order;
        if (max_order >= bs) max_order = bs - 1;
        int precision = ps->blocksize <= 384 ? 13 : (ps->blocksize <= 1152 ? 14 : 15);
        /* window the data: lpc.c FLAC__lpc_window_data (float data, float window) */
        float *wd = (float *)(work + 2 * bs);
        for (int i = 0; i < bs; ++i) wd[i] = (float)x[i] * window[i];
        double autoc[MAX_LPC_ORDER + 1];
        for (int l = 0; l <= max_order; ++l) {
            double s = 0.0;
            for (int i = l; i < bs; ++i) s += (double)wd[i] * (double)wd[i - l];
            autoc[l] = s;
        }
        if (autoc[0] != 0.0) {
            /* Levinson-Durbin: lpc.c FLAC__lpc_compute_lp_coefficients */
            double lpc[MAX_LPC_ORDER], lp_coeff[MAX_LPC_ORDER][MAX_LPC_ORDER], error[MAX_LPC_ORDER];
            double err = autoc[0];
            int mo = max_order;
            for (int i = 0; i < mo; ++i) {
                double r = -autoc[i + 1];
                for (int j = 0; j < i; ++j) r -= lpc[j] * autoc[i - j];
                r /= err;
                lpc[i] = r;
                int j;
                for (j = 0; j < (i >> 1); ++j) {
                    double tmp = lpc[j];
                    lpc[j] += r * lpc[i - 1 - j];
                    lpc[i - 1 - j] += r * tmp;
                }
                if (i & 1) lpc[j] += lpc[j] * r;
                err *= (1.0 - r * r);
                for (j = 0; j <= i; ++j) lp_coeff[i][j] = (double)(float)(-lpc[j]);
                error[i] = err;
                if (err == 0.0) { mo = i + 1; break; }
            }
            /* FLAC__lpc_compute_best_order */
            int best_order;
            {
                double error_scale = 0.5 / (double)bs;
                double best_bits = 1e300;
                int best_index = 0;
                int overhead = bps + precision;
                for (int idx = 0, order = 1; idx < mo; ++idx, ++order) {
                    double bits = expected_bits_per_sample(error[idx], error_scale) * (double)(bs - order) +
                                  (double)(order * overhead);
                    if (bits < best_bits) { best_index = idx; best_bits = bits; }
                }
                best_order = best_index + 1;
            }
            int order = best_order;
            double rbps = expected_bits_per_sample(error[order - 1], 0.5 / (double)(bs - order));
            if (rbps < (double)bps) {
                /* FLAC__lpc_quantize_coefficients */
                int prec = precision - 1;
                int32_t qmax = (1 << prec) - 1, qmin = -(1 << prec);
                double cmax = 0.0;
                for (int i = 0; i < order; ++i) { double d = fabs(lp_coeff[order - 1][i]); if (d > cmax) cmax = d; }
                int okq = cmax > 0.0;
                int shift = 0;
                int32_t q[MAX_LPC_ORDER];
                if (okq) {
                    int log2cmax;
                    (void)frexp(cmax, &log2cmax);
                    log2cmax--;
                    shift = prec - log2cmax - 1;
                    if (shift > 15) shift = 15;
                    else if (shift < 0) okq = 0; /* negative shifts are not emitted (RFC 9639 forbids them) */
                }
                if (okq) {
                    double e = 0.0;
                    for (int i = 0; i < order; ++i) {
                        e += lp_coeff[order - 1][i] * (double)(1 << shift);
                        long qq = lround(e);
                        if (qq > qmax) qq = qmax; else if (qq < qmin) qq = qmin;
                        e -= (double)qq;
                        q[i] = (int32_t)qq;
                    }
                    int good = 1;
                    for (int i = order; i < bs; ++i) {
                        int64_t sum = 0;
                        for (int j = 0; j < order; ++j) sum += (int64_t)q[j] * x[i - 1 - j];
                        int64_t e2 = x[i] - (sum >> shift);
                        if (!fits_i32(e2)) { good = 0; break; }
                        res[i] = e2;
                    }
                    if (good) {
                        int porder, rice2, params[256];
                        uint64_t rb = rice_search(res, bs, order, ps->max_partition_order, &porder, params, &rice2);
                        uint64_t bits = hdr + 4 + 5 + (uint64_t)order * (uint64_t)(precision + bps) + rb;
                        if (bits < pl->bits) {
                            pl->type = 3; pl->order = order; pl->porder = porder; pl->rice2 = rice2;
                            pl->qlp_precision = precision; pl->qlp_shift = shift;
                            memcpy(pl->qlp, q, sizeof(int32_t) * (size_t)order);
                            memcpy(pl->params, params, sizeof(int) * (size_t)(1 << porder));
                            pl->bits = bits;
                            memcpy(res_out, res, sizeof(int64_t) * bs);
                        }
                    }
                }
            }
        }
    }
}

static void write_subframe(BitWriter *bw, const SubframePlan *pl, const int64_t *xin, const int64_t *res, int bs) {
    int w = pl->wasted, bps = pl->bps;
    int typebits = pl->type == 0 ? 0 : pl->type == 1 ? 1 : pl->type == 2 ? (8 + pl->order) : (32 + pl->order - 1);
    bw_write(bw, 0, 1);
    bw_write(bw, (uint64_t)typebits, 6);
    if (w) { bw_write(bw, 1, 1); bw_zeros(bw, w - 1); bw_write(bw, 1, 1); }
    else bw_write(bw, 0, 1);
    uint64_t mask = bps == 64 ? ~0ull : ((1ull << bps) - 1);
    if (pl->type == 0) { bw_write(bw, (uint64_t)(xin[0] >> w) & mask, bps); return; }
    if (pl->type == 1) { for (int i = 0; i < bs; ++i) bw_write(bw, (uint64_t)(xin[i] >> w) & mask, bps); return; }
    for (int i = 0; i < pl->order; ++i) bw_write(bw, (uint64_t)(xin[i] >> w) & mask, bps);
    if (pl->type == 3) {
        bw_write(bw, (uint64_t)(pl->qlp_precision - 1), 4);
        bw_write(bw, (uint64_t)pl->qlp_shift & 31, 5);
        for (int j = 0; j < pl->order; ++j)
            bw_write(bw, (uint64_t)(int64_t)pl->qlp[j] & ((1ull << pl->qlp_precision) - 1), pl->qlp_precision);
    }
    bw_write(bw, (uint64_t)pl->rice2, 2);
    bw_write(bw, (uint64_t)pl->porder, 4);
    int i = pl->order;
    for (int part = 0; part < (1 << pl->porder); ++part) {
        int n = (bs >> pl->porder) - (part == 0 ? pl->order : 0);
        bw_write(bw, (uint64_t)pl->params[part], pl->rice2 ? 5 : 4);
        for (int j = 0; j < n; ++j, ++i) bw_rice(bw, res[i], pl->params[part]);
    }
}

static int utf8_put(uint8_t *p, uint64_t v) {
    if (v < 0x80) { p[0] = (uint8_t)v; return 1; }
    int n = v < 0x800 ? 2 : v < 0x10000 ? 3 : v < 0x200000 ? 4 : v < 0x4000000 ? 5 : v < 0x80000000ull ? 6 : 7;
    static const uint8_t lead[8] = {0, 0, 0xC0, 0xE0, 0xF0, 0xF8, 0xFC, 0xFE};
    for (int i = n - 1; i > 0; --i) { p[i] = (uint8_t)(0x80 | (v & 0x3F)); v >>= 6; }
    p[0] = (uint8_t)(lead[n] | v);
    return n;
}

static const char vendor[] = "reference libFLAC 1.5.0 20250211";

/* Upper bound on the encoded size of one stream */
int64_t orc_encode_bound(int64_t stream_size, int n_channels, int level) {
    int bs = presets[level].blocksize;
    int64_t nframes = (stream_size + bs - 1) / bs;
    return 4 + 4 + 34 + 4 + 8 + (int64_t)sizeof(vendor) + nframes * (32 + ((int64_t)bs * 33 * n_channels + 7) / 8 + 64) + 1024;
}

/*
 * Encode one stream (interleaved int32, n_channels 1|2) as a complete FLAC file the way the
 * reference drives libFLAC (compress.c:184-237): fLaC + STREAMINFO (unknown totals, no MD5: no
 * seek callback is given) + VORBIS_COMMENT(vendor) + fixed-blocksize frames.
 * stereo_mode: -1 = per preset (search), 0 = independent, 1 = left/side, 2 = side/right, 3 = mid/side.
 * Returns bytes written or -1.
 */
int64_t orc_encode_stream(const int32_t *data, int64_t stream_size, int n_channels, int level, int stereo_mode,
                          uint8_t *out, int64_t cap) {
    crc_init();
    if (level < 0 || level > 8) return -1;
    const Preset *ps = &presets[level];
    int bs_nom = ps->blocksize;
    memset(out, 0, (size_t)cap);
    int64_t pos = 0;
    memcpy(out, "fLaC", 4); pos = 4;
    /* STREAMINFO */
    out[pos++] = 0x00; out[pos++] = 0; out[pos++] = 0; out[pos++] = 34;
    {
        uint8_t *s = out + pos;
        s[0] = (uint8_t)(bs_nom >> 8); s[1] = (uint8_t)bs_nom; s[2] = s[0]; s[3] = s[1];
        /* min/max framesize = 0 (unknown) */
        uint32_t sr = 44100;
        s[10] = (uint8_t)(sr >> 12); s[11] = (uint8_t)(sr >> 4);
        s[12] = (uint8_t)(((sr & 0xF) << 4) | ((n_channels - 1) << 1) | ((31 >> 4) & 1));
        s[13] = (uint8_t)((31 & 0xF) << 4); /* total samples = 0 */
        pos += 34;
    }
    /* VORBIS_COMMENT, last */
    {
        uint32_t vlen = (uint32_t)strlen(vendor);
        uint32_t len = 4 + vlen + 4;
        out[pos++] = 0x84; out[pos++] = (uint8_t)(len >> 16); out[pos++] = (uint8_t)(len >> 8); out[pos++] = (uint8_t)len;
        out[pos++] = (uint8_t)vlen; out[pos++] = (uint8_t)(vlen >> 8); out[pos++] = (uint8_t)(vlen >> 16); out[pos++] = (uint8_t)(vlen >> 24);
        memcpy(out + pos, vendor, vlen); pos += vlen;
        pos += 4; /* zero comments */
    }
    float *window = (float *)malloc(sizeof(float) * (size_t)bs_nom);
    int64_t *buf = (int64_t *)malloc(sizeof(int64_t) * (size_t)bs_nom * 16);
    if (!window || !buf) { free(window); free(buf); return -1; }
    make_tukey(window, bs_nom);
    int64_t *L = buf, *R = buf + bs_nom, *M = buf + 2 * bs_nom, *S = buf + 3 * bs_nom;
    int64_t *resbuf[4] = {buf + 4 * bs_nom, buf + 5 * bs_nom, buf + 6 * bs_nom, buf + 7 * bs_nom};
    int64_t *work = buf + 8 * bs_nom; /* 3*bs needed */
    int64_t nframes = (stream_size + bs_nom - 1) / bs_nom;
    int64_t ret = 0;
    for (int64_t f = 0; f < nframes; ++f) {
        int bs = (int)((f + 1) * bs_nom <= stream_size ? bs_nom : stream_size - f * bs_nom);
        const int32_t *src = data + f * bs_nom * n_channels;
        for (int i = 0; i < bs; ++i) {
            L[i] = src[i * n_channels];
            if (n_channels == 2) R[i] = src[i * n_channels + 1];
        }
        SubframePlan plans[4];
        int ca = 0; /* channel assignment field */
        const int64_t *sig0 = L, *sig1 = R;
        int p0 = 0, p1 = 1;
        plan_subframe(L, bs, 32, ps, window, &plans[0], resbuf[0], work);
        if (n_channels == 2) {
            plan_subframe(R, bs, 32, ps, window, &plans[1], resbuf[1], work);
            ca = 1;
            int mode = stereo_mode;
            int search = (mode < 0 && ps->do_mid_side);
            if (search || mode > 0) {
                for (int i = 0; i < bs; ++i) { M[i] = (L[i] + R[i]) >> 1; S[i] = L[i] - R[i]; }
                plan_subframe(M, bs, 32, ps, window, &plans[2], resbuf[2], work);
                plan_subframe(S, bs, 33, ps, window, &plans[3], resbuf[3], work);
                if (search) {
                    uint64_t bits[4] = {plans[0].bits + plans[1].bits, plans[0].bits + plans[3].bits,
                                        plans[1].bits + plans[3].bits, plans[2].bits + plans[3].bits};
                    mode = 0;
                    for (int m = 1; m < 4; ++m) if (bits[m] < bits[mode]) mode = m;
                }
            }
            if (mode < 0) mode = 0;
            if (mode == 1) { ca = 8; sig0 = L; sig1 = S; p0 = 0; p1 = 3; }
            else if (mode == 2) { ca = 9; sig0 = S; sig1 = R; p0 = 3; p1 = 1; }
            else if (mode == 3) { ca = 10; sig0 = M; sig1 = S; p0 = 2; p1 = 3; }
        }
        /* frame header */
        uint8_t *fp = out + pos;
        int64_t fcap = cap - pos;
        if (fcap < 32) { ret = -1; break; }
        int h = 0;
        fp[h++] = 0xFF; fp[h++] = 0xF8;
        int bs_code;
        if (bs == 192) bs_code = 1;
        else if (bs == 576 || bs == 1152 || bs == 2304 || bs == 4608) bs_code = bs == 576 ? 2 : bs == 1152 ? 3 : bs == 2304 ? 4 : 5;
        else if (bs == 256 || bs == 512 || bs == 1024 || bs == 2048 || bs == 4096 || bs == 8192 || bs == 16384 || bs == 32768) {
            bs_code = 8; int t = bs >> 8; while (t > 1) { bs_code++; t >>= 1; }
        } else bs_code = bs <= 256 ? 6 : 7;
        fp[h++] = (uint8_t)((bs_code << 4) | 9); /* 44.1 kHz */
        fp[h++] = (uint8_t)((ca << 4) | (7 << 1)); /* 32 bps */
        h += utf8_put(fp + h, (uint64_t)f);
        if (bs_code == 6) fp[h++] = (uint8_t)(bs - 1);
        else if (bs_code == 7) { fp[h++] = (uint8_t)((bs - 1) >> 8); fp[h++] = (uint8_t)(bs - 1); }
        fp[h] = orc_crc8(fp, h); h++;
        BitWriter bw = {fp, fcap - 2, (int64_t)h * 8, 0};
        write_subframe(&bw, &plans[p0], sig0, resbuf[p0], bs);
        if (n_channels == 2) write_subframe(&bw, &plans[p1], sig1, resbuf[p1], bs);
        if (bw.err) { ret = -1; break; }
        int64_t end = (bw.pos + 7) >> 3;
        uint16_t c = orc_crc16(fp, end);
        fp[end] = (uint8_t)(c >> 8); fp[end + 1] = (uint8_t)c;
        pos += end + 2;
    }
    free(window); free(buf);
    return ret < 0 ? -1 : pos;
}

/* compress.c:133-435: per-stream encode, then exclusive prefix sum of sizes and concatenation. */
static int orc_encode(const int32_t *data, int64_t n_stream, int64_t stream_size, int n_channels, uint32_t level,
                      int64_t *n_bytes, int64_t *starts, unsigned char **bytes, int use_threads) {
    if (level > 8) return ERROR_INVALID_LEVEL;
    if (n_stream == 0) return ERROR_ZERO_NSTREAM;
    if (stream_size == 0) return ERROR_ZERO_STREAMSIZE;
    *n_bytes = 0;
    *bytes = NULL;
    uint8_t **bufs = (uint8_t **)calloc((size_t)n_stream, sizeof(uint8_t *));
    int64_t *sizes = (int64_t *)calloc((size_t)n_stream, sizeof(int64_t));
    if (!bufs || !sizes) { free(bufs); free(sizes); return ERROR_ALLOC; }
    int64_t bound = orc_encode_bound(stream_size, n_channels, (int)level);
    int errors = ERROR_NONE;
#pragma omp parallel for schedule(static) reduction(| : errors) if (use_threads)
    for (int64_t i = 0; i < n_stream; ++i) {
        bufs[i] = (uint8_t *)malloc((size_t)bound);
        if (!bufs[i]) { errors |= ERROR_ALLOC; continue; }
        sizes[i] = orc_encode_stream(data + i * stream_size * n_channels, stream_size, n_channels, (int)level, -1,
                                     bufs[i], bound);
        if (sizes[i] < 0) errors |= ERROR_ENCODE_PROCESS;
    }
    if (errors == ERROR_NONE) {
        for (int64_t i = 0; i < n_stream; ++i) { starts[i] = *n_bytes; *n_bytes += sizes[i]; }
        *bytes = (unsigned char *)malloc((size_t)(*n_bytes));
        if (!*bytes) errors |= ERROR_ALLOC;
        else for (int64_t i = 0; i < n_stream; ++i) memcpy(*bytes + starts[i], bufs[i], (size_t)sizes[i]);
    }
    for (int64_t i = 0; i < n_stream; ++i) free(bufs[i]);
    free(bufs); free(sizes);
    return errors;
}

int orc_encode_i32(const int32_t *data, int64_t n_stream, int64_t stream_size, uint32_t level, int64_t *n_bytes,
                   int64_t *starts, unsigned char **bytes, int use_threads) {
    return orc_encode(data, n_stream, stream_size, 1, level, n_bytes, starts, bytes, use_threads);
}

/* compress.c:482-540 with utils.c:112-116 (LE reinterpret of int64 as [lo, hi] int32 pairs) */
int orc_encode_i64(const int64_t *data, int64_t n_stream, int64_t stream_size, uint32_t level, int64_t *n_bytes,
                   int64_t *starts, unsigned char **bytes, int use_threads) {
    return orc_encode((const int32_t *)data, n_stream, stream_size, 2, level, n_bytes, starts, bytes, use_threads);
}

void orc_free(void *p) { free(p); }

/* ------------------------------------------------------------------------------------------ */
/* Float <-> int converters: restatement of utils.c:160-368 (same types at every step).        */
/* Compiled with -std=c11 -ffp-contract=off like the reference build (meson.build:8).          */
/* ------------------------------------------------------------------------------------------ */
int orc_float32_to_int32(const float *input, int64_t n_stream, int64_t stream_size, const float *quanta,
                         int32_t *output, float *offsets, float *gains) {
    int32_t flac_max = 2147483647;
    for (int64_t is = 0; is < n_stream; ++is) {
        const float *in = input + is * stream_size;
        float smin = in[0], smax = in[0];
        for (int64_t i = 1; i < stream_size; ++i) {
            float v = in[i];
            if (v < smin) smin = v;
            if (v > smax) smax = v;
        }
        offsets[is] = 0.5 * (smin + smax);               /* utils.c:194 float add, double mul, ->float */
        float amp;
        if ((smin - offsets[is]) > (smax - offsets[is])) amp = 1.01 * (smin - offsets[is]);
        else amp = 1.01 * (smax - offsets[is]);          /* utils.c:198-202 */
        float min_quanta = amp / flac_max;               /* utils.c:203 int->float conversion */
        float squanta = quanta == NULL ? min_quanta : quanta[is];
        int64_t nquant = (int64_t)((double)offsets[is] / (double)squanta); /* utils.c:221 */
        offsets[is] = (float)((double)squanta * (double)nquant);           /* utils.c:222 */
        if (squanta == 0) gains[is] = 1.0;
        else gains[is] = 1.0 / squanta;                  /* utils.c:229 double div -> float */
        int32_t *o = output + is * stream_size;
        for (int64_t i = 0; i < stream_size; ++i) {
            float st = in[i] - offsets[is];
            if (st >= 0) o[i] = (int32_t)(gains[is] * st + 0.5);   /* float mul, double add, trunc */
            else o[i] = (int32_t)(gains[is] * st - 0.5);
        }
    }
    return ERROR_NONE;
}

int orc_float64_to_int64(const double *input, int64_t n_stream, int64_t stream_size, const double *quanta,
                         int64_t *output, double *offsets, double *gains) {
    int64_t flac_max = 9223372036854775807LL;
    for (int64_t is = 0; is < n_stream; ++is) {
        const double *in = input + is * stream_size;
        double smin = in[0], smax = in[0];
        for (int64_t i = 1; i < stream_size; ++i) {
            double v = in[i];
            if (v < smin) smin = v;
            if (v > smax) smax = v;
        }
        offsets[is] = 0.5 * (smin + smax);
        double amp;
        if ((smin - offsets[is]) > (smax - offsets[is])) amp = 1.01 * (smin - offsets[is]);
        else amp = 1.01 * (smax - offsets[is]);
        double min_quanta = amp / flac_max;
        double squanta = quanta == NULL ? min_quanta : quanta[is];
        int64_t nquant = (int64_t)(offsets[is] / squanta);
        offsets[is] = squanta * (double)nquant;
        if (squanta == 0) gains[is] = 1.0;
        else gains[is] = 1.0 / squanta;
        int64_t *o = output + is * stream_size;
        for (int64_t i = 0; i < stream_size; ++i) {
            double st = in[i] - offsets[is];
            if (st >= 0) o[i] = (int64_t)(gains[is] * st + 0.5);
            else o[i] = (int64_t)(gains[is] * st - 0.5);
        }
    }
    return ERROR_NONE;
}

void orc_int64_to_float64(const int64_t *input, int64_t n_stream, int64_t stream_size, const double *offsets,
                          const double *gains, double *output) {
    for (int64_t is = 0; is < n_stream; ++is) {
        double coeff = 1.0 / gains[is];
        for (int64_t i = 0; i < stream_size; ++i)
            output[is * stream_size + i] = offsets[is] + coeff * (double)input[is * stream_size + i];
    }
}

void orc_int32_to_float32(const int32_t *input, int64_t n_stream, int64_t stream_size, const float *offsets,
                          const float *gains, float *output) {
    for (int64_t is = 0; is < n_stream; ++is) {
        float coeff = 1.0 / gains[is];
        for (int64_t i = 0; i < stream_size; ++i)
            output[is * stream_size + i] = offsets[is] + coeff * (float)input[is * stream_size + i];
    }
}
