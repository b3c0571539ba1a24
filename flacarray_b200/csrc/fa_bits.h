// fa_bits.h -- MSB-first bit reader over global memory, CRC tables, frame-header parsing.
// Shared by the decoder kernels.  Bit-stream layout: RFC 9639 (the reference delegates this to
// libFLAC behind decompress.c:256-305 / compress.c:184-237).
#pragma once
#include "fa_simt.h"

namespace fa {

// Error bits: identical to the reference's flacarray.h:20-40 so callers see the same codes.
enum : int {
    kErrNone = 0,
    kErrAlloc = 1 << 0,
    kErrInvalidLevel = 1 << 1,
    kErrZeroNstream = 1 << 2,
    kErrZeroStreamsize = 1 << 3,
    kErrEncodeProcess = 1 << 9,
    kErrEncodeCollect = 1 << 11,
    kErrDecodeInit = 1 << 13,
    kErrDecodeProcess = 1 << 14,
    kErrDecodeStreamsize = 1 << 16,
    kErrDecodeSampleRange = 1 << 17,
    kErrDecodeSeek = 1 << 18,
    kErrConvertType = 1 << 19,
    kErrCuda = 1 << 20,  // extension: CUDA runtime failure (no device, launch error, OOM)
};

// ---- CRC tables (filled on the host once, uploaded to global memory) ---------------------------
struct CrcTables {
    uint8_t crc8[256];        // poly 0x07
    uint16_t crc16[4][256];   // poly 0x8005, slice-by-4: [0] = plain byte table
    uint16_t shift_hi[17][256];  // multiply state by x^(8 * 2^j) mod P: contribution of the high byte
    uint16_t shift_lo[17][256];  //                                      ... of the low byte
};

inline void crc_tables_init(CrcTables* t) {
    for (int i = 0; i < 256; ++i) {
        uint8_t c = (uint8_t)i;
        for (int b = 0; b < 8; ++b) c = (c & 0x80) ? (uint8_t)((c << 1) ^ 0x07) : (uint8_t)(c << 1);
        t->crc8[i] = c;
        uint16_t d = (uint16_t)(i << 8);
        for (int b = 0; b < 8; ++b) d = (d & 0x8000) ? (uint16_t)((d << 1) ^ 0x8005) : (uint16_t)(d << 1);
        t->crc16[0][i] = d;
    }
    // slice tables: crc16[k][b] = CRC state after byte b followed by k zero bytes
    for (int k = 1; k < 4; ++k)
        for (int i = 0; i < 256; ++i) {
            uint16_t c = t->crc16[k - 1][i];
            t->crc16[k][i] = (uint16_t)((c << 8) ^ t->crc16[0][c >> 8]);
        }
    // shift tables: advancing a state s over n zero bytes is linear in s.
    // level 0 = one zero byte.
    for (int i = 0; i < 256; ++i) {
        uint16_t hi = (uint16_t)(i << 8), lo = (uint16_t)i;
        t->shift_hi[0][i] = (uint16_t)((hi << 8) ^ t->crc16[0][hi >> 8]);
        t->shift_lo[0][i] = (uint16_t)((lo << 8) ^ t->crc16[0][lo >> 8]);
    }
    for (int j = 1; j < 17; ++j)
        for (int i = 0; i < 256; ++i) {
            // apply level j-1 twice
            uint16_t a = t->shift_hi[j - 1][i];
            t->shift_hi[j][i] = (uint16_t)(t->shift_hi[j - 1][a >> 8] ^ t->shift_lo[j - 1][a & 0xFF]);
            uint16_t b = t->shift_lo[j - 1][i];
            t->shift_lo[j][i] = (uint16_t)(t->shift_hi[j - 1][b >> 8] ^ t->shift_lo[j - 1][b & 0xFF]);
        }
}

FA_D uint32_t crc16_byte(const CrcTables* t, uint32_t crc, uint32_t byte) {
    return ((crc << 8) & 0xFFFF) ^ t->crc16[0][((crc >> 8) ^ byte) & 0xFF];
}
// state after `crc` is followed by 2^j zero bytes
FA_D uint32_t crc16_shift_pow2(const CrcTables* t, uint32_t crc, int j) {
    return (uint32_t)(t->shift_hi[j][(crc >> 8) & 0xFF] ^ t->shift_lo[j][crc & 0xFF]);
}
// state after `crc` is followed by n zero bytes (n < 2^17)
FA_D uint32_t crc16_shift(const CrcTables* t, uint32_t crc, uint32_t n) {
    for (int j = 0; n; ++j, n >>= 1)
        if (n & 1) crc = crc16_shift_pow2(t, crc, j);
    return crc;
}

// CRC-16 over one big-endian word / one byte with the slice-by-4 tables copied to shared memory
// (T = [4][256] uint16, T[k][b] = state after byte b followed by k zero bytes).
FA_D uint32_t crc16_word(const uint16_t* T, uint32_t c, uint32_t w) {
    // T: [4][256] slice tables in shared memory
    return (uint32_t)(T[3 * 256 + (((c >> 8) ^ (w >> 24)) & 0xFF)] ^ T[2 * 256 + (((c & 0xFF) ^ (w >> 16)) & 0xFF)] ^
                      T[256 + ((w >> 8) & 0xFF)] ^ T[w & 0xFF]);
}
// The same word step without tables.  P = x^16 + x^15 + x^2 + 1 = (x + 1)(x^15 + x + 1): modulo the trinomial a
// high part folds back with two shifts (x^15 = x + 1), modulo x + 1 a polynomial is its parity, and the two
// residues are recombined by adding P's cofactor 0x8003 when the parities disagree.  ~20 integer instructions per
// word and no shared-memory traffic, against four table reads whose random indices cost ~2.7 bank-conflict
// wavefronts each.
FA_D uint32_t crc16_word_alu(uint32_t c, uint32_t w) {
    const uint32_t x = w ^ (c << 16);                 // x(t) * t^16 mod P is the new state
    const uint32_t h1 = x >> 15;
    const uint32_t v1 = (x & 0x7FFFu) ^ h1 ^ (h1 << 1);      // x mod (t^15 + t + 1), 18 bits
    const uint32_t h2 = v1 >> 15;
    const uint32_t m = (v1 & 0x7FFFu) ^ h2 ^ (h2 << 1);      // ... 15 bits
    const uint32_t b = (m << 1) ^ (m << 2);                  // * t^16 = * (t^2 + t) mod the trinomial
    const uint32_t h3 = b >> 15;
    const uint32_t a = (b & 0x7FFFu) ^ h3 ^ (h3 << 1);
    const uint32_t q = (uint32_t)popc32(a ^ x) & 1u;         // parity(a) != parity(x)
    return a ^ ((0u - q) & 0x8003u);
}
FA_D uint32_t crc16_b(const uint16_t* T, uint32_t c, uint32_t byte) {
    return ((c << 8) & 0xFFFF) ^ T[((c >> 8) ^ byte) & 0xFF];
}
// ---- CRC-16 in the trinomial domain (the bulk passes: k_dec_crc, k_enc_compact) -----------------------------------
// P = (t + 1) T with T = t^15 + t + 1.  Instead of the CRC state the bulk passes carry the residue R = M mod T of the
// MESSAGE polynomial (lazily reduced: any 32-bit value congruent to it) and the XOR of all message words (its parity
// is M mod (t + 1)); the CRC is formed once per frame by crct_finish.  Everything is shifts and XORs:
//   * t^15 = t + 1, so a value v folds as (v & 0x7FFF) ^ h ^ (h << 1), h = v >> 15 (32 bits -> 18, 29 bits -> 15);
//   * t^(2^k) mod T only has terms out of {t, t^2, t^4, t^8} for every k >= 3 (t^8, t^16 = t^2 + t, t^32 = t^4 + t^2, ...,
//     period 15 in k), so advancing a residue over 2^j bytes is at most four shifts -- no tables, no shared memory.
FA_D uint32_t crct_fold(uint32_t v) {
    const uint32_t h = v >> 15;
    return (v & 0x7FFFu) ^ h ^ (h << 1);
}
// multiply by the constant with terms t^1, t^2, t^4, t^8 selected by `m` (bits 1, 2, 4, 8 of m); a < 2^24
template <uint32_t M>
FA_D uint32_t crct_mulc(uint32_t a) {
    uint32_t r = 0;
    if (M & 2u) r ^= a << 1;
    if (M & 4u) r ^= a << 2;
    if (M & 16u) r ^= a << 4;
    if (M & 256u) r ^= a << 8;
    return r;
}
FA_D uint32_t crct_mulv(uint32_t a, uint32_t m) {     // the same with a run-time constant
    return ((0u - ((m >> 1) & 1u)) & (a << 1)) ^ ((0u - ((m >> 2) & 1u)) & (a << 2)) ^ ((0u - ((m >> 4) & 1u)) & (a << 4)) ^
           ((0u - ((m >> 8) & 1u)) & (a << 8));
}
// t^(8 * 2^j) mod T = t^(2^(j + 3)), j = 0..14 (then it repeats)
FA_D uint32_t crct_pow2_const(int j) {
    const uint32_t K[15] = {0x100u, 0x006u, 0x014u, 0x110u, 0x106u, 0x012u, 0x104u, 0x016u, 0x114u, 0x116u, 0x112u, 0x102u,
                            0x002u, 0x004u, 0x010u};
    return K[j % 15];
}
// residue after `r` is followed by n zero bytes; returns a value below 2^15
FA_D uint32_t crct_shift_bytes(uint32_t r, uint64_t n) {
    r = crct_fold(crct_fold(r));
    for (int j = 0; n; ++j, n >>= 1)
        if (n & 1) {
            const uint32_t m = crct_pow2_const(j);
            r = crct_fold(crct_mulv(r, m));
        }
    return r;
}
// residue of 16 message bytes (q = the four little-endian words as loaded); 32 bits, lazily reduced
FA_D uint32_t crct_chunk(const U4& q) {
    uint32_t r = bswap32(q.x);
    r = crct_mulc<0x14u>(crct_fold(r)) ^ bswap32(q.y);       // * t^32, next word
    r = crct_mulc<0x14u>(crct_fold(r)) ^ bswap32(q.z);
    r = crct_mulc<0x14u>(crct_fold(r)) ^ bswap32(q.w);
    return r;
}
FA_D uint32_t crct_byte(uint32_t r, uint32_t byte) { return crct_fold((r << 8) ^ byte); }   // r < 2^24
// CRC-16 of the message from its residue mod T (any 32-bit representative) and the XOR of its words (any word size)
FA_D uint32_t crct_finish(uint32_t r, uint32_t xor_words) {
    const uint32_t m = crct_fold(crct_fold(r));                       // 15 bits
    const uint32_t a = crct_fold(crct_mulc<0x6u>(m));                 // * t^16: the CRC is M t^16 mod P
    const uint32_t q = (uint32_t)popc32(a ^ xor_words) & 1u;          // parity(a) != parity(M): add P's cofactor T
    return a ^ ((0u - q) & 0x8003u);
}

// word steps of the CRC pass: table-free arithmetic (FAB_CRC_ALU words out of every 2) or shared-memory tables
#ifndef FAB_CRC_ALU
#define FAB_CRC_ALU 1
#endif
#define FAB_CRC_TAB_(T, c, w) crc16_word(T, c, w)
#define FAB_CRC_ALU_(T, c, w) crc16_word_alu(c, w)
#if FAB_CRC_ALU == 2
#define FAB_CRC_STEP0 FAB_CRC_ALU_
#define FAB_CRC_STEP1 FAB_CRC_ALU_
#elif FAB_CRC_ALU == 1
#define FAB_CRC_STEP0 FAB_CRC_ALU_
#define FAB_CRC_STEP1 FAB_CRC_TAB_
#else
#define FAB_CRC_STEP0 FAB_CRC_TAB_
#define FAB_CRC_STEP1 FAB_CRC_TAB_
#endif

// ---- Bit reader ---------------------------------------------------------------------------------
// Reads aligned 32-bit words (coalescing-friendly, L1/L2 sector reuse) and never touches a word
// that holds no valid byte of [start, end).
struct BitRd {
    const uint32_t* wp;    // next aligned word to fetch
    const uint32_t* w0;    // aligned word holding the first byte
    const uint32_t* wend;  // first aligned word with no valid byte
    uint64_t buf;          // unread bits, MSB aligned; bits below `n` are zero
    int n;                 // number of valid bits in buf
    int a8;                // 8 * (start byte offset inside *w0)
    int err;
};

FA_D void br_init(BitRd& br, const uint8_t* start, const uint8_t* end) {
    uintptr_t s = (uintptr_t)start;
    br.w0 = (const uint32_t*)(s & ~(uintptr_t)3);
    br.wend = (const uint32_t*)(((uintptr_t)end + 3) & ~(uintptr_t)3);
    br.a8 = (int)(s & 3) * 8;
    br.wp = br.w0;
    br.buf = 0;
    br.n = 0;
    br.err = 0;
    if (br.wp < br.wend) {
        uint32_t w = bswap32(ldg32(br.wp));
        br.wp++;
        br.buf = ((uint64_t)w << 32) << br.a8;
        br.n = 32 - br.a8;
    }
}

FA_D void br_refill(BitRd& br) {
    if (br.n <= 32) {
        uint32_t w = 0;
        if (br.wp < br.wend) w = bswap32(ldg32(br.wp));
        br.wp++;
        br.buf |= (uint64_t)w << (32 - br.n);
        br.n += 32;
    }
}

// bits consumed since br_init
FA_D int64_t br_pos(const BitRd& br) { return (int64_t)(br.wp - br.w0) * 32 - br.n - br.a8; }

// nb in [0, 32]
FA_D uint32_t br_read(BitRd& br, int nb) {
    br_refill(br);
    uint32_t v = nb ? (uint32_t)(br.buf >> (64 - nb)) : 0u;
    br.buf = nb ? (br.buf << nb) : br.buf;
    br.n -= nb;
    return v;
}

// signed, nb in [0, 33]
FA_D int64_t br_read_signed(BitRd& br, int nb) {
    if (nb == 0) return 0;
    uint64_t v;
    if (nb > 32) {
        uint64_t hi = br_read(br, nb - 32);
        v = (hi << 32) | br_read(br, 32);
    } else {
        v = br_read(br, nb);
    }
    uint64_t sign = 1ull << (nb - 1);
    return (int64_t)((v ^ sign) - sign);
}

// number of zero bits before the next one bit; consumes the one bit
FA_D uint32_t br_unary(BitRd& br) {
    uint32_t q = 0;
    for (;;) {
        br_refill(br);
        if (br.buf != 0) {
            int z = clz64(br.buf);
            q += (uint32_t)z;
            br.buf = (br.buf << z) << 1;  // two steps: z + 1 may be 64
            br.n -= z + 1;
            return q;
        }
        q += (uint32_t)br.n;
        br.n = 0;
        if (br.wp > br.wend + 1) {  // ran off the end of the stream
            br.err = 1;
            return q;
        }
    }
}

// ---- Frame header -------------------------------------------------------------------------------
struct FrameHdr {
    int blocksize;
    int ca;        // raw channel-assignment field
    int nch;
    int bps;       // 0 = take from STREAMINFO
    int variable;  // blocking strategy bit
    int hdr_bytes;
    uint64_t number;
};

// Parse + validate (reserved fields, UTF-8 number, CRC-8) a frame header at p; `avail` bytes readable.
FA_D bool parse_frame_header(const uint8_t* p, int64_t avail, const CrcTables* t, FrameHdr& fh) {
    if (avail < 6) return false;
    if (p[0] != 0xFF || (p[1] & 0xFE) != 0xF8) return false;
    fh.variable = p[1] & 1;
    int bs_code = p[2] >> 4, sr_code = p[2] & 0xF, ch_code = p[3] >> 4, ss_code = (p[3] >> 1) & 7;
    if ((p[3] & 1) || bs_code == 0 || sr_code == 15 || ch_code > 10 || ss_code == 3) return false;
    int pos = 4;
    uint32_t b0 = p[pos++];
    uint64_t v;
    int extra;
    if (!(b0 & 0x80)) { v = b0; extra = 0; }
    else if ((b0 & 0xE0) == 0xC0) { v = b0 & 0x1F; extra = 1; }
    else if ((b0 & 0xF0) == 0xE0) { v = b0 & 0x0F; extra = 2; }
    else if ((b0 & 0xF8) == 0xF0) { v = b0 & 0x07; extra = 3; }
    else if ((b0 & 0xFC) == 0xF8) { v = b0 & 0x03; extra = 4; }
    else if ((b0 & 0xFE) == 0xFC) { v = b0 & 0x01; extra = 5; }
    else if (b0 == 0xFE && fh.variable) { v = 0; extra = 6; }
    else return false;
    int need = pos + extra + (bs_code == 6 ? 1 : bs_code == 7 ? 2 : 0) +
               (sr_code == 12 ? 1 : (sr_code == 13 || sr_code == 14) ? 2 : 0) + 1;
    if (need > avail) return false;
    for (int i = 0; i < extra; ++i) {
        uint32_t b = p[pos++];
        if ((b & 0xC0) != 0x80) return false;
        v = (v << 6) | (b & 0x3F);
    }
    fh.number = v;
    int bs;
    if (bs_code == 1) bs = 192;
    else if (bs_code <= 5) bs = 576 << (bs_code - 2);
    else if (bs_code == 6) bs = p[pos++] + 1;
    else if (bs_code == 7) { bs = ((p[pos] << 8) | p[pos + 1]) + 1; pos += 2; }
    else bs = 256 << (bs_code - 8);
    if (sr_code == 12) pos += 1;
    else if (sr_code == 13 || sr_code == 14) pos += 2;
    uint32_t c = 0;
    for (int i = 0; i < pos; ++i) c = t->crc8[c ^ p[i]];
    if (c != p[pos]) return false;
    pos++;
    fh.blocksize = bs;
    fh.ca = ch_code;
    fh.nch = ch_code < 8 ? ch_code + 1 : 2;
    fh.bps = ss_code == 0 ? 0 : ss_code == 1 ? 8 : ss_code == 2 ? 12 : ss_code == 4 ? 16 : ss_code == 5 ? 20 : ss_code == 6 ? 24 : 32;
    fh.hdr_bytes = pos;
    return true;
}

}  // namespace fa
