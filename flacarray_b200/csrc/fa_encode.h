// fa_encode.h -- FLAC frame encoder body: one 256-thread CTA encodes one (stream, frame).
//
// Replaces the libFLAC encoder the reference drives per stream in compress.c:184-237 (serial) and
// compress.c:337-390 (OpenMP), plus the byte bookkeeping of the write callbacks
// (compress.c:13-104) and the final prefix-sum/concatenation (compress.c:402-429).
//
// Per (frame, channel): wasted-bits/constant detection -> fixed-predictor error sums (orders 0-4)
// -> tukey window + autocorrelation (FP64 accumulate, fixed reduction order => deterministic)
// -> Levinson-Durbin + order choice + coefficient quantisation (one thread, FP64)
// -> residual partition sums for the fixed and the LPC candidate -> Rice partition/parameter search
// (one warp per candidate) -> exact code lengths -> block prefix sum -> bit packing into a
// shared-memory staged frame -> CRC-8 / CRC-16 -> decoupled look-back scan over frame sizes gives the
// frame's final byte offset -> coalesced copy to HBM.  Input is read from HBM exactly once and the
// compressed bytes are written exactly once.
#pragma once
#include "fa_bits.h"
#include "fa_quant.h"
#include <math.h>

namespace fa {

constexpr int kEncThreads = 256;
constexpr int kSpt = 16;        // samples per thread
constexpr int kMaxBs = kEncThreads * kSpt;  // 4096
constexpr int kMaxOrd = 12;     // libFLAC presets never exceed order 12
constexpr int kMaxParts = 64;   // partition order <= 6
constexpr int kSmpWords = kMaxBs + kMaxBs / 16 + 16;

enum { kI32 = 0, kI64 = 1, kF32 = 2, kF64 = 3 };

// libFLAC compression_levels_[]: blocksize for set_blocksize(0), max LPC order, max
// residual partition order.  Stereo decorrelation is not searched: on (low word, high word) pairs it
// buys < 0.01 %, so 2-channel frames are always "independent".
struct LevelPreset { int blocksize, max_lpc_order, max_porder; };
inline LevelPreset level_preset(int level) {
    const LevelPreset t[9] = {{1152, 0, 3}, {1152, 0, 3}, {1152, 0, 3}, {4096, 6, 4}, {4096, 8, 4},
                              {4096, 8, 5}, {4096, 8, 6}, {4096, 12, 6}, {4096, 12, 6}};
    return t[level];
}

// tukey(0.5) window of the nominal blocksize (libFLAC window.c FLAC__window_tukey); short last frames
// use its prefix.  Host side, computed once per (context, blocksize).
inline void make_tukey_window(float* w, int L) {
    for (int n = 0; n < L; ++n) w[n] = 1.0f;
    int Np = (int)(0.5f / 2.0f * (float)L) - 1;
    if (Np > 0) {
        for (int n = 0; n <= Np; ++n) {
            w[n] = (float)(0.5 - 0.5 * cos(3.14159265358979323846 * n / Np));
            w[L - Np - 1 + n] = (float)(0.5 - 0.5 * cos(3.14159265358979323846 * (n + Np) / Np));
        }
    }
}

// Bytes before the first frame of every stream: "fLaC" + STREAMINFO + APPLICATION(faB2 table, last).
inline int stream_header_bytes(int nframes) { return 4 + 4 + 34 + 4 + 8 + 3 * nframes; }

struct EncParams {
    const void* data;
    int dtype;                 // kI32 / kI64 / kF32 / kF64
    const void* offsets;       // per-stream float/double (kF32/kF64), written by the quantise pre-pass
    const void* gains;
    int64_t n_stream, stream_size;
    int nch;
    int blocksize, nframes;    // per stream
    int max_lpc_order, max_porder, qlp_precision;
    const float* window;       // tukey(0.5) of length blocksize
    const CrcTables* crc;
    uint8_t* out;
    int64_t out_capacity;
    long long* starts;         // [n_stream] byte offset of every stream; must be preset to -1
    long long* ends;           // [n_stream]
    unsigned long long* desc;  // [n_stream * nframes] look-back descriptors, zeroed
    uint32_t* ticket;          // zeroed
    int* err;
    int hdr_bytes;
};

struct Plan {
    int type;      // 0 constant, 1 verbatim, 2 fixed, 3 lpc
    int order, wasted, shift, prec, porder, rice2;
    uint32_t res_bits;  // estimated bits of the residual section
    int32_t qlp[kMaxOrd];
};

struct EncShared {
    Plan cand[2];  // [0] fixed, [1] lpc
    Plan plan;     // winner
    int cand_ok[2];
    int maxp[2];
    unsigned long long psum[2][kMaxParts];
    uint8_t kpar[2][2 * kMaxParts];  // params for porder p at offset (1 << p) - 1
    uint8_t params[kMaxParts];
    double autoc[kMaxOrd + 1];
    unsigned long long fix_err[5];
    uint32_t fix_bad;
    uint32_t red[8];
    uint32_t scan[8];
    uint32_t crc_part[kEncThreads];
    int bitpos;
    int sub_total_bits;
    long long frame_off;
    long long stream_start;
    uint32_t g;
};

FA_D int pidx(int i) { return i + (i >> 4); }

// ---- block-wide helpers (256 threads = 8 warps) ---------------------------------------------------
FA_D uint32_t block_or(uint32_t v, uint32_t* red) {
    for (int m = 16; m >= 1; m >>= 1) v |= shfl_xor(v, m);
    if (lane() == 0) red[warp()] = v;
    sync();
    uint32_t r = 0;
    for (int w = 0; w < kEncThreads / 32; ++w) r |= red[w];
    sync();
    return r;
}

FA_D uint32_t block_excl_scan(uint32_t v, uint32_t* wt, uint32_t& total) {
    uint32_t inc = v;
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t n = shfl_up(inc, d);
        if (lane() >= d) inc += n;
    }
    if (lane() == 31) wt[warp()] = inc;
    sync();
    uint32_t base = 0, tot = 0;
    for (int w = 0; w < kEncThreads / 32; ++w) {
        uint32_t x = wt[w];
        if (w < warp()) base += x;
        tot += x;
    }
    sync();
    total = tot;
    return base + inc - v;
}

FA_D double shfl_xor_d(double v, int m) {
    unsigned long long u;
    memcpy(&u, &v, 8);
    uint32_t lo = shfl_xor((uint32_t)u, m), hi = shfl_xor((uint32_t)(u >> 32), m);
    u = ((unsigned long long)hi << 32) | lo;
    memcpy(&v, &u, 8);
    return v;
}
FA_D unsigned long long shfl_xor_u64(unsigned long long u, int m) {
    uint32_t lo = shfl_xor((uint32_t)u, m), hi = shfl_xor((uint32_t)(u >> 32), m);
    return ((unsigned long long)hi << 32) | lo;
}

// ---- per-thread bit packer into the shared-memory frame buffer ------------------------------------
// The buffer is pre-zeroed, so runs of zero bits (unary quotients) only advance the position.
// A thread owns every word that lies entirely inside its bit range (plain store); its first and last
// words may be shared with neighbours (atomic OR).
struct BitPk {
    uint32_t* out;
    uint32_t cur;
    int pos;
    bool first;
};
FA_D void pk_begin(BitPk& pk, uint32_t* out, int pos) { pk.out = out; pk.cur = 0; pk.pos = pos; pk.first = true; }
FA_D void pk_flush(BitPk& pk, int word) {
    if (pk.cur) {
        if (pk.first) atom_or_shared(&pk.out[word], pk.cur);
        else pk.out[word] = pk.cur;
    }
    pk.first = false;
    pk.cur = 0;
}
// nb in [1, 32], v < 2^nb
FA_D void pk_emit(BitPk& pk, uint32_t v, int nb) {
    int off = pk.pos & 31, space = 32 - off;
    if (nb < space) {
        pk.cur |= v << (space - nb);
    } else if (nb == space) {
        pk.cur |= v;
        pk_flush(pk, pk.pos >> 5);
    } else {
        int rem = nb - space;  // 1..31
        pk.cur |= v >> rem;
        pk_flush(pk, pk.pos >> 5);
        pk.cur = v << (32 - rem);
    }
    pk.pos += nb;
}
FA_D void pk_emit64(BitPk& pk, uint64_t v, int nb) {  // nb in [1, 64]
    if (nb > 32) { pk_emit(pk, (uint32_t)(v >> 32) & (nb == 64 ? 0xFFFFFFFFu : ((1u << (nb - 32)) - 1u)), nb - 32); nb = 32; }
    pk_emit(pk, nb == 32 ? (uint32_t)v : ((uint32_t)v & ((1u << nb) - 1u)), nb);
}
FA_D void pk_skip(BitPk& pk, uint32_t q) {
    int np = pk.pos + (int)q;
    if ((np >> 5) != (pk.pos >> 5)) pk_flush(pk, pk.pos >> 5);
    pk.pos = np;
}
FA_D void pk_end(BitPk& pk) {
    if (pk.cur) atom_or_shared(&pk.out[pk.pos >> 5], pk.cur);
    pk.cur = 0;
}

// ---- residual of the thread's 16 samples for a compile-time predictor order ------------------------
// xw[j] = sample (i0 - kMaxOrd + j).  res[j] valid for i0 + j in [order, bs).
template <int ORD>
FA_D void residual_block(const int32_t* xw, const int32_t* coef, int shift, int64_t* res) {
#pragma unroll
    for (int j = 0; j < kSpt; ++j) {
        int64_t sum = 0;
#pragma unroll
        for (int m = 0; m < ORD; ++m) sum += (int64_t)coef[m] * (int64_t)xw[kMaxOrd + j - 1 - m];
        res[j] = (int64_t)xw[kMaxOrd + j] - (sum >> shift);
    }
}
FA_D void residual_dispatch(int order, const int32_t* xw, const int32_t* coef, int shift, int64_t* res) {
    switch (order) {
    case 0: residual_block<0>(xw, coef, shift, res); break;
    case 1: residual_block<1>(xw, coef, shift, res); break;
    case 2: residual_block<2>(xw, coef, shift, res); break;
    case 3: residual_block<3>(xw, coef, shift, res); break;
    case 4: residual_block<4>(xw, coef, shift, res); break;
    case 5: residual_block<5>(xw, coef, shift, res); break;
    case 6: residual_block<6>(xw, coef, shift, res); break;
    case 7: residual_block<7>(xw, coef, shift, res); break;
    case 8: residual_block<8>(xw, coef, shift, res); break;
    case 9: residual_block<9>(xw, coef, shift, res); break;
    case 10: residual_block<10>(xw, coef, shift, res); break;
    case 11: residual_block<11>(xw, coef, shift, res); break;
    default: residual_block<12>(xw, coef, shift, res); break;
    }
}

FA_D void fixed_coefs(int order, int32_t* c) {
    const int32_t fx[5][4] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, -1, 0, 0}, {3, -3, 1, 0}, {4, -6, 4, -1}};
    for (int j = 0; j < kMaxOrd; ++j) c[j] = (j < 4) ? fx[order][j] : 0;
}

FA_D bool fits_res(int64_t v) { return v >= -2147483647LL && v <= 2147483647LL; }

FA_D int max_porder_for(int bs, int order, int level_max) {
    int p = 0;
    while (p < level_max && !((bs >> p) & 1)) p++;   // largest p with 2^p | bs, capped
    while (p > 0 && (bs >> p) <= order) p--;
    return p;
}

// Accumulate |res| of the thread's samples into the finest-level partition sums (shared atomics at
// partition boundaries only).  Returns false if some residual does not fit the Rice coder.
FA_D bool partition_sums(const int64_t* res, int i0, int bs, int order, int psize, unsigned long long* psum) {
    bool ok = true;
    int i = i0 < order ? order : i0;
    int iend = i0 + kSpt < bs ? i0 + kSpt : bs;
    if (i >= iend) return true;
    int part = i / psize;
    int next = (part + 1) * psize;
    unsigned long long acc = 0;
    for (; i < iend; ++i) {
        if (i == next) {
            atom_add_shared64(&psum[part], acc);
            acc = 0;
            part++;
            next += psize;
        }
        int64_t r = res[i - i0];
        ok = ok && fits_res(r);
        acc += (unsigned long long)(r < 0 ? -r : r);
    }
    atom_add_shared64(&psum[part], acc);
    return ok;
}

// One warp: libFLAC's estimate-based partition-order / Rice-parameter search
// (find_best_partition_order_ / set_partitioned_rice_ without escape codes).
FA_D void rice_search_warp(EncShared* sh, int cand, int bs, int order, int maxp) {
    unsigned long long* ps = sh->psum[cand];
    uint32_t best_bits = 0xFFFFFFFFu;
    int best_p = maxp;
    for (int p = maxp; p >= 0; --p) {
        int nparts = 1 << p;
        uint32_t bits = 0;
        for (int part = lane(); part < nparts; part += 32) {
            unsigned long long sum = ps[part];
            uint32_t n = (uint32_t)(bs >> p) - (part == 0 ? (uint32_t)order : 0u);
            int k = 0;
            while (k < 30 && ((unsigned long long)n << k) < sum) k++;
            sh->kpar[cand][(1 << p) - 1 + part] = (uint8_t)k;
            unsigned long long pb = 4ull + (unsigned long long)(1 + k) * n + (k ? (sum >> (k - 1)) : (sum << 1));
            pb -= (n >> 1);
            bits += pb > (1ull << 25) ? (1u << 25) : (uint32_t)pb;
        }
        for (int m = 16; m >= 1; m >>= 1) bits += shfl_xor(bits, m);
        bits += 6;
        if (bits < best_bits) { best_bits = bits; best_p = p; }
        // merge to the next coarser level
        if (p > 0) {
            unsigned long long a[2] = {0, 0};
            int half = nparts >> 1;
            int cnt = 0;
            for (int j = lane(); j < half; j += 32) a[cnt++] = ps[2 * j] + ps[2 * j + 1];
            syncwarp();
            cnt = 0;
            for (int j = lane(); j < half; j += 32) ps[j] = a[cnt++];
            syncwarp();
        }
    }
    if (lane() == 0) {
        sh->cand[cand].porder = best_p;
        sh->cand[cand].res_bits = best_bits;
    }
    syncwarp();
}

// Levinson-Durbin, order choice and coefficient quantisation (libFLAC lpc.c procedure), one thread.
FA_D void lpc_design(EncShared* sh, int bs, int bps, int max_order, int precision) {
    Plan& pl = sh->cand[1];
    sh->cand_ok[1] = 0;
    const double* autoc = sh->autoc;
    if (!(autoc[0] != 0.0)) return;
    double lpc[kMaxOrd], coefs[kMaxOrd][kMaxOrd], error[kMaxOrd];
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
        for (j = 0; j <= i; ++j) coefs[i][j] = (double)(float)(-lpc[j]);
        error[i] = err;
        if (err == 0.0) { mo = i + 1; break; }
    }
    // FLAC__lpc_compute_best_order
    const double ln2 = 0.69314718055994530942;
    double escale = 0.5 / (double)bs;
    double best_bits = 1e300;
    int order = 1;
    for (int idx = 0; idx < mo; ++idx) {
        double e = error[idx], b;
        if (e > 0.0) { b = 0.5 * log(escale * e) / ln2; if (b < 0.0) b = 0.0; }
        else if (e < 0.0) b = 1e32;
        else b = 0.0;
        double bits = b * (double)(bs - (idx + 1)) + (double)((idx + 1) * (bps + precision));
        if (bits < best_bits) { best_bits = bits; order = idx + 1; }
    }
    {
        double e = error[order - 1], b;
        double es = 0.5 / (double)(bs - order);
        if (e > 0.0) { b = 0.5 * log(es * e) / ln2; if (b < 0.0) b = 0.0; }
        else if (e < 0.0) b = 1e32;
        else b = 0.0;
        if (!(b < (double)bps)) return;
    }
    // FLAC__lpc_quantize_coefficients
    int prec = precision - 1;
    int32_t qmax = (1 << prec) - 1, qmin = -(1 << prec);
    double cmax = 0.0;
    for (int i = 0; i < order; ++i) { double d = fabs(coefs[order - 1][i]); if (d > cmax) cmax = d; }
    if (!(cmax > 0.0)) return;
    int log2cmax;
    (void)frexp(cmax, &log2cmax);
    log2cmax--;
    int shift = prec - log2cmax - 1;
    if (shift > 15) shift = 15;
    if (shift < 0) return;
    double e = 0.0;
    for (int i = 0; i < order; ++i) {
        e += coefs[order - 1][i] * (double)(1 << shift);
        double rq = e < 0.0 ? -floor(-e + 0.5) : floor(e + 0.5);  // lround: half away from zero
        long long q = (long long)rq;
        if (q > qmax) q = qmax; else if (q < qmin) q = qmin;
        e -= (double)q;
        pl.qlp[i] = (int32_t)q;
    }
    for (int i = order; i < kMaxOrd; ++i) pl.qlp[i] = 0;
    pl.type = 3;
    pl.order = order;
    pl.shift = shift;
    pl.prec = precision;
    sh->cand_ok[1] = 1;
}

FA_D int utf8_put(uint8_t* p, uint64_t v) {
    if (v < 0x80) { p[0] = (uint8_t)v; return 1; }
    int n = v < 0x800 ? 2 : v < 0x10000 ? 3 : v < 0x200000 ? 4 : v < 0x4000000 ? 5 : v < 0x80000000ull ? 6 : 7;
    const uint8_t lead[8] = {0, 0, 0xC0, 0xE0, 0xF0, 0xF8, 0xFC, 0xFE};
    for (int i = n - 1; i > 0; --i) { p[i] = (uint8_t)(0x80 | (v & 0x3F)); v >>= 6; }
    p[0] = (uint8_t)(lead[n] | v);
    return n;
}

FA_D int blocksize_code(int bs) {
    if (bs == 192) return 1;
    if (bs == 576) return 2;
    if (bs == 1152) return 3;
    if (bs == 2304) return 4;
    if (bs == 4608) return 5;
    for (int c = 8; c <= 15; ++c) if (bs == (256 << (c - 8))) return c;
    return bs <= 256 ? 6 : 7;
}

// byte k of the staged frame (words hold the bit-stream MSB first)
FA_D uint32_t out_byte(const uint32_t* out, int k) { return (out[k >> 2] >> (24 - 8 * (k & 3))) & 0xFFu; }

// ------------------------------------------------------------------------------------------------
// The CTA body.  `smem_raw` is the dynamic shared memory base (>= enc_smem_bytes(nch)).
// ------------------------------------------------------------------------------------------------
inline size_t enc_out_words(int nch) { return (size_t)(nch * (kMaxBs * 32 / 8 + 8) + 64) / 4; }
inline size_t enc_smem_bytes(int nch) {
    return ((sizeof(EncShared) + 15) & ~(size_t)15) + (size_t)nch * kSmpWords * 4 + (size_t)kSmpWords * 4 +
           enc_out_words(nch) * 4 + 16;
}

FA_D void encode_frame_cta(const EncParams& P, unsigned char* smem_raw) {
    EncShared* sh = (EncShared*)smem_raw;
    int32_t* smp = (int32_t*)(smem_raw + ((sizeof(EncShared) + 15) & ~(size_t)15));
    float* wd = (float*)(smp + (size_t)P.nch * kSmpWords);      // also reduction scratch
    uint32_t* out = (uint32_t*)(wd + kSmpWords);
    const int out_words = (int)(((size_t)P.nch * (kMaxBs * 32 / 8 + 8) + 64) / 4);
    const int t = tid();

    // ---- work assignment: tickets are handed out in launch order so that look-back never waits on
    // a CTA that has not started (decoupled look-back, Merrill & Garland).
    if (t == 0) sh->g = atom_add_global(P.ticket, 1u);
    sync();
    const uint32_t g = sh->g;
    const int64_t s = (int64_t)(g / (uint32_t)P.nframes);
    const int f = (int)(g % (uint32_t)P.nframes);
    const int64_t samp0 = (int64_t)f * P.blocksize;
    const int bs = (int)((P.stream_size - samp0) < P.blocksize ? (P.stream_size - samp0) : P.blocksize);
    const int nch = P.nch;

    // ---- load (and quantise) the frame: coalesced reads, padded-blocked layout in shared memory
    {
        float off32 = 0.f, gain32 = 0.f;
        double off64 = 0., gain64 = 0.;
        if (P.dtype == kF32) { off32 = ((const float*)P.offsets)[s]; gain32 = ((const float*)P.gains)[s]; }
        if (P.dtype == kF64) { off64 = ((const double*)P.offsets)[s]; gain64 = ((const double*)P.gains)[s]; }
        const int64_t base = s * P.stream_size + samp0;
        for (int i = t; i < bs; i += kEncThreads) {
            if (P.dtype == kI32) {
                smp[pidx(i)] = ((const int32_t*)P.data)[base + i];
            } else if (P.dtype == kF32) {
                smp[pidx(i)] = quant_f32(((const float*)P.data)[base + i], off32, gain32);
            } else {
                long long v;
                if (P.dtype == kI64) v = ((const long long*)P.data)[base + i];
                else v = quant_f64(((const double*)P.data)[base + i], off64, gain64);
                smp[pidx(i)] = (int32_t)(uint32_t)((unsigned long long)v & 0xFFFFFFFFull);  // ch0 = low word
                smp[kSmpWords + pidx(i)] = (int32_t)(v >> 32);                              // ch1 = high word
            }
        }
        for (int w = t; w < out_words; w += kEncThreads) out[w] = 0;
    }
    sync();

    // ---- frame header (thread 0): RFC 9639 9.1
    if (t == 0) {
        uint8_t h[16];
        int n = 0;
        int bc = blocksize_code(bs);
        h[n++] = 0xFF; h[n++] = 0xF8;
        h[n++] = (uint8_t)((bc << 4) | 9);                       // 44.1 kHz like the reference's default
        h[n++] = (uint8_t)(((nch == 2 ? 1 : 0) << 4) | (7 << 1));  // independent channels, 32 bps
        n += utf8_put(h + n, (uint64_t)f);
        if (bc == 6) h[n++] = (uint8_t)(bs - 1);
        else if (bc == 7) { h[n++] = (uint8_t)((bs - 1) >> 8); h[n++] = (uint8_t)(bs - 1); }
        uint32_t c = 0;
        for (int i = 0; i < n; ++i) c = P.crc->crc8[c ^ h[i]];
        h[n++] = (uint8_t)c;
        for (int i = 0; i < n; ++i) out[i >> 2] |= (uint32_t)h[i] << (24 - 8 * (i & 3));
        sh->bitpos = n * 8;
    }
    sync();

    const int i0 = t * kSpt;
    for (int c = 0; c < nch; ++c) {
        int32_t* x = smp + (size_t)c * kSmpWords;
        // ---- wasted bits / constant ----
        uint32_t orv = 0, diff = 0;
        {
            int32_t x0 = x[0];
            for (int j = 0; j < kSpt; ++j) {
                int i = i0 + j;
                if (i < bs) { int32_t v = x[pidx(i)]; orv |= (uint32_t)v; diff |= (uint32_t)(v ^ x0); }
            }
        }
        orv = block_or(orv, sh->red);
        diff = block_or(diff, sh->red);
        int wasted = (orv == 0) ? 0 : ctz32(orv);
        const int bps = 32 - wasted;
        if (wasted) {
            for (int j = 0; j < kSpt; ++j) { int i = i0 + j; if (i < bs) x[pidx(i)] >>= wasted; }
            sync();
        }
        const bool constant = (diff == 0);
        const uint32_t verbatim_bits = (uint32_t)bps * (uint32_t)bs;
        const bool try_pred = !constant && bs > 4;

        // thread-local window of samples: xw[j] = x[i0 - kMaxOrd + j]
        int32_t xw[kMaxOrd + kSpt];
#pragma unroll
        for (int j = 0; j < kMaxOrd + kSpt; ++j) {
            int i = i0 - kMaxOrd + j;
            xw[j] = (i >= 0 && i < bs) ? x[pidx(i)] : 0;
        }

        if (t == 0) { sh->cand_ok[0] = 0; sh->cand_ok[1] = 0; sh->fix_bad = 0; }
        if (t < 2 * kMaxParts) ((unsigned long long*)sh->psum)[t] = 0;

        if (try_pred) {
            // ---- fixed predictors: sum |e_k| over i >= 4, k = 0..4 (libFLAC fixed.c) ----
            unsigned long long fe[5] = {0, 0, 0, 0, 0};
            uint32_t bad = 0;
#pragma unroll
            for (int j = 0; j < kSpt; ++j) {
                int i = i0 + j;
                if (i >= 4 && i < bs) {
                    int64_t a0 = xw[kMaxOrd + j], a1 = xw[kMaxOrd + j - 1], a2 = xw[kMaxOrd + j - 2],
                            a3 = xw[kMaxOrd + j - 3], a4 = xw[kMaxOrd + j - 4];
                    int64_t e0 = a0, e1 = a0 - a1, e2 = e1 - (a1 - a2), e3 = e2 - (a1 - 2 * a2 + a3),
                            e4 = e3 - (a1 - 3 * a2 + 3 * a3 - a4);
                    int64_t e[5] = {e0, e1, e2, e3, e4};
#pragma unroll
                    for (int k = 0; k < 5; ++k) {
                        if (!fits_res(e[k])) bad |= 1u << k;
                        fe[k] += (unsigned long long)(e[k] < 0 ? -e[k] : e[k]);
                    }
                }
            }
            // reduce: pairs by shuffle, then [k][128] in scratch, warps 0..4 finish
            unsigned long long* scr = (unsigned long long*)wd;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                unsigned long long o = shfl_xor_u64(fe[k], 1);
                if (!(t & 1)) scr[k * 128 + (t >> 1)] = fe[k] + o;
            }
            bad = block_or(bad, sh->red);  // includes barriers: scratch visible
            if (warp() < 5) {
                int k = warp();
                unsigned long long v = scr[k * 128 + lane()] + scr[k * 128 + 32 + lane()] + scr[k * 128 + 64 + lane()] +
                                       scr[k * 128 + 96 + lane()];
                for (int m = 16; m >= 1; m >>= 1) v += shfl_xor_u64(v, m);
                if (lane() == 0) sh->fix_err[k] = v;
            }
            sync();
            if (t == 0) {
                unsigned long long te[5];
                for (int k = 0; k < 5; ++k) te[k] = ((bad >> k) & 1) ? ~0ull : sh->fix_err[k];
                unsigned long long m34 = te[3] < te[4] ? te[3] : te[4];
                unsigned long long m234 = te[2] < m34 ? te[2] : m34;
                unsigned long long m1234 = te[1] < m234 ? te[1] : m234;
                int order;
                if (te[0] < m1234) order = 0;
                else if (te[1] < m234) order = 1;
                else if (te[2] < m34) order = 2;
                else if (te[3] < te[4]) order = 3;
                else order = 4;
                Plan& pl = sh->cand[0];
                pl.type = 2; pl.order = order; pl.shift = 0; pl.prec = 0;
                bool ok = te[order] != ~0ull;
                if (ok) {
                    double n = (double)(bs - 4);
                    double rb = te[order] > 0 ? log(0.69314718055994530942 * (double)te[order] / n) / 0.69314718055994530942 : 0.0;
                    ok = rb < (double)bps;
                }
                sh->cand_ok[0] = ok ? 1 : 0;
                sh->maxp[0] = max_porder_for(bs, order, P.max_porder);
            }
            sync();  // scratch (wd) is free again

            // ---- LPC analysis ----
            int max_order = P.max_lpc_order;
            if (max_order >= bs) max_order = bs - 1;
            if (max_order > 0) {
                // window (libFLAC lpc.c FLAC__lpc_window_data: float data * float window)
#pragma unroll
                for (int j = 0; j < kSpt; ++j) {
                    int i = i0 + j;
                    if (i < bs) wd[pidx(i)] = fmul((float)xw[kMaxOrd + j], P.window[i]);
                }
                sync();
                double ac[kMaxOrd + 1];
#pragma unroll
                for (int l = 0; l <= kMaxOrd; ++l) ac[l] = 0.0;
                {
                    double w[kMaxOrd + kSpt];
#pragma unroll
                    for (int j = 0; j < kMaxOrd + kSpt; ++j) {
                        int i = i0 - kMaxOrd + j;
                        w[j] = (i >= 0 && i < bs) ? (double)wd[pidx(i)] : 0.0;
                    }
#pragma unroll
                    for (int l = 0; l <= kMaxOrd; ++l) {
                        if (l <= max_order) {
#pragma unroll
                            for (int j = 0; j < kSpt; ++j) ac[l] = dfma(w[kMaxOrd + j], w[kMaxOrd + j - l], ac[l]);
                        }
                    }
                }
                sync();  // everyone has read wd: reuse it as reduction scratch
                double* dscr = (double*)wd;
#pragma unroll
                for (int l = 0; l <= kMaxOrd; ++l) {
                    double o = shfl_xor_d(ac[l], 1);
                    if (!(t & 1)) dscr[l * 128 + (t >> 1)] = ac[l] + o;
                }
                sync();
                for (int l = warp(); l <= max_order; l += kEncThreads / 32) {
                    double v = (dscr[l * 128 + lane()] + dscr[l * 128 + 32 + lane()]) +
                               (dscr[l * 128 + 64 + lane()] + dscr[l * 128 + 96 + lane()]);
                    for (int m = 16; m >= 1; m >>= 1) v += shfl_xor_d(v, m);
                    if (lane() == 0) sh->autoc[l] = v;
                }
                sync();
                if (t == 0) {
                    lpc_design(sh, bs, bps, max_order, P.qlp_precision);
                    if (sh->cand_ok[1]) sh->maxp[1] = max_porder_for(bs, sh->cand[1].order, P.max_porder);
                }
                sync();
            }

            // ---- residual partition sums for both candidates ----
            uint32_t resbad = 0;
            for (int cd = 0; cd < 2; ++cd) {
                if (!sh->cand_ok[cd]) continue;   // block-uniform
                int32_t coef[kMaxOrd];
                const Plan& pl = sh->cand[cd];
                if (cd == 0) fixed_coefs(pl.order, coef);
                else { for (int m = 0; m < kMaxOrd; ++m) coef[m] = pl.qlp[m]; }
                int64_t res[kSpt];
                residual_dispatch(pl.order, xw, coef, pl.shift, res);
                int psize = bs >> sh->maxp[cd];
                if (!partition_sums(res, i0, bs, pl.order, psize, sh->psum[cd])) resbad |= 1u << cd;
            }
            resbad = block_or(resbad, sh->red);
            if (warp() < 2 && sh->cand_ok[warp()] && !((resbad >> warp()) & 1))
                rice_search_warp(sh, warp(), bs, sh->cand[warp()].order, sh->maxp[warp()]);
            sync();
            if (t == 0) {
                // ---- choose: verbatim vs fixed vs lpc by estimated size (stream_encoder.c process_subframe_)
                uint32_t best = verbatim_bits;
                int win = -1;
                for (int cd = 0; cd < 2; ++cd) {
                    if (!sh->cand_ok[cd] || ((resbad >> cd) & 1)) continue;
                    const Plan& pl = sh->cand[cd];
                    uint32_t bits = (uint32_t)pl.order * (uint32_t)bps + pl.res_bits;
                    if (cd == 1) bits += 4 + 5 + (uint32_t)pl.order * (uint32_t)pl.prec;
                    if (bits < best) { best = bits; win = cd; }
                }
                if (win < 0) { sh->plan.type = 1; sh->plan.order = 0; }
                else sh->plan = sh->cand[win];
                if (win >= 0) {
                    int p = sh->plan.porder;
                    int rice2 = 0;
                    for (int q = 0; q < (1 << p); ++q) {
                        uint8_t k = sh->kpar[win][(1 << p) - 1 + q];
                        sh->params[q] = k;
                        if (k >= 15) rice2 = 1;
                    }
                    sh->plan.rice2 = rice2;
                }
                sh->plan.wasted = wasted;
            }
            sync();
        } else {
            if (t == 0) { sh->plan.type = constant ? 0 : 1; sh->plan.order = 0; sh->plan.wasted = wasted; }
            sync();
        }

        // ---- exact size of the predictive subframe; fall back to verbatim if it does not pay ----
        int64_t res[kSpt];
        uint32_t lens = 0;
        int ptype = sh->plan.type;
        int order = sh->plan.order;
        int porder = sh->plan.porder;
        int psize = bs >> porder;
        int plen = sh->plan.rice2 ? 5 : 4;
        if (ptype >= 2) {
            int32_t coef[kMaxOrd];
            if (ptype == 2) fixed_coefs(order, coef);
            else { for (int m = 0; m < kMaxOrd; ++m) coef[m] = sh->plan.qlp[m]; }
            residual_dispatch(order, xw, coef, sh->plan.shift, res);
            int i = i0 < order ? order : i0;
            int iend = i0 + kSpt < bs ? i0 + kSpt : bs;
            if (i < iend) {
                int part = i / psize;
                int next = (part + 1) * psize;
                int k = sh->params[part];
                if (i == (part == 0 ? order : part * psize)) lens += plen;
                for (; i < iend; ++i) {
                    if (i == next) { part++; next += psize; k = sh->params[part]; lens += plen; }
                    int64_t r = res[i - i0];
                    uint32_t u = ((uint32_t)r << 1) ^ (uint32_t)(r >> 63);
                    lens += (u >> k) + 1 + k;
                }
            }
        }
        uint32_t total;
        uint32_t excl = block_excl_scan(lens, sh->scan, total);
        if (ptype >= 2) {
            uint32_t hdr_bits = (uint32_t)order * (uint32_t)bps + 6 + (ptype == 3 ? 9 + (uint32_t)order * (uint32_t)sh->plan.prec : 0);
            if (hdr_bits + total >= verbatim_bits) ptype = 1;  // block-uniform
        }

        // ---- pack ----
        int pos0 = sh->bitpos;
        int sub_hdr_bits = 8 + (wasted ? wasted : 0);
        if (t == 0) {
            BitPk pk;
            pk_begin(pk, out, pos0);
            int typebits = ptype == 0 ? 0 : ptype == 1 ? 1 : ptype == 2 ? (8 + order) : (32 + order - 1);
            pk_emit(pk, (uint32_t)typebits << 1 | (wasted ? 1u : 0u), 8);
            if (wasted) { pk_skip(pk, (uint32_t)(wasted - 1)); pk_emit(pk, 1, 1); }
            if (ptype == 0) {
                pk_emit64(pk, (uint64_t)(uint32_t)x[0], bps);
            } else if (ptype >= 2) {
                for (int j = 0; j < order; ++j) pk_emit64(pk, (uint64_t)(uint32_t)x[pidx(j)], bps);
                if (ptype == 3) {
                    pk_emit(pk, (uint32_t)(sh->plan.prec - 1), 4);
                    pk_emit(pk, (uint32_t)sh->plan.shift & 31u, 5);
                    for (int j = 0; j < order; ++j)
                        pk_emit(pk, (uint32_t)sh->plan.qlp[j] & ((1u << sh->plan.prec) - 1u), sh->plan.prec);
                }
                pk_emit(pk, (uint32_t)sh->plan.rice2, 2);
                pk_emit(pk, (uint32_t)porder, 4);
            }
            pk_end(pk);
        }
        int body0;  // first bit of the per-sample payload
        if (ptype == 0) body0 = pos0 + sub_hdr_bits + bps;
        else if (ptype == 1) body0 = pos0 + sub_hdr_bits;
        else body0 = pos0 + sub_hdr_bits + order * bps + 6 + (ptype == 3 ? 9 + order * sh->plan.prec : 0);
        int sub_end;
        if (ptype == 0) {
            sub_end = body0;
        } else if (ptype == 1) {
            BitPk pk;
            int start = i0 < bs ? i0 : bs;
            pk_begin(pk, out, body0 + start * bps);
            for (int j = 0; j < kSpt; ++j) {
                int i = i0 + j;
                if (i < bs) pk_emit64(pk, (uint64_t)(uint32_t)xw[kMaxOrd + j] & (bps == 32 ? 0xFFFFFFFFull : ((1ull << bps) - 1)), bps);
            }
            pk_end(pk);
            sub_end = body0 + bs * bps;
        } else {
            BitPk pk;
            pk_begin(pk, out, body0 + (int)excl);
            int i = i0 < order ? order : i0;
            int iend = i0 + kSpt < bs ? i0 + kSpt : bs;
            if (i < iend) {
                int part = i / psize;
                int next = (part + 1) * psize;
                int k = sh->params[part];
                if (i == (part == 0 ? order : part * psize)) pk_emit(pk, (uint32_t)k, plen);
                for (; i < iend; ++i) {
                    if (i == next) { part++; next += psize; k = sh->params[part]; pk_emit(pk, (uint32_t)k, plen); }
                    int64_t r = res[i - i0];
                    uint32_t u = ((uint32_t)r << 1) ^ (uint32_t)(r >> 63);
                    pk_skip(pk, u >> k);
                    pk_emit(pk, (1u << k) | (u & ((1u << k) - 1u)), k + 1);
                }
            }
            pk_end(pk);
            sub_end = body0 + (int)total;
        }
        sync();
        if (t == 0) sh->bitpos = sub_end;
        sync();
    }

    // ---- frame footer: pad to a byte, CRC-16 over the whole frame ----
    const int nbytes_body = (sh->bitpos + 7) >> 3;
    {
        // each thread: CRC of a contiguous chunk (front-padded so that all chunks have equal length),
        // then a log-step combine with the "advance by 2^j zero bytes" tables.
        const CrcTables* T = P.crc;
        int chunk = (nbytes_body + kEncThreads - 1) / kEncThreads;
        int lg = 0;
        while ((1 << lg) < chunk) lg++;
        chunk = 1 << lg;                      // power of two => combine uses one table level per step
        int pad = chunk * kEncThreads - nbytes_body;
        int b0 = t * chunk - pad, b1 = b0 + chunk;
        uint32_t c = 0;
        int k = b0 < 0 ? 0 : b0;
        // byte-wise to a word boundary, then slice-by-4
        for (; k < b1 && (k & 3); ++k) c = crc16_byte(T, c, out_byte(out, k));
        for (; k + 4 <= b1; k += 4) {
            uint32_t w = out[k >> 2];
            c = (uint32_t)(T->crc16[3][((c >> 8) ^ (w >> 24)) & 0xFF] ^ T->crc16[2][((c & 0xFF) ^ (w >> 16)) & 0xFF] ^
                           T->crc16[1][(w >> 8) & 0xFF] ^ T->crc16[0][w & 0xFF]);
        }
        for (; k < b1; ++k) c = crc16_byte(T, c, out_byte(out, k));
        // combine: at step j, thread t (with bit j clear) merges partner t + 2^j whose block is 2^j chunks long
        sh->crc_part[t] = c;
        sync();
        for (int j = 0; j < 8; ++j) {
            if ((t & ((2 << j) - 1)) == 0) {
                uint32_t a = sh->crc_part[t], b = sh->crc_part[t + (1 << j)];
                sh->crc_part[t] = crc16_shift_pow2(T, a, lg + j) ^ b;
            }
            sync();
        }
    }
    const int frame_bytes = nbytes_body + 2;
    if (t == 0) {
        uint32_t c = sh->crc_part[0];
        int k = nbytes_body;
        out[k >> 2] |= ((c >> 8) & 0xFF) << (24 - 8 * (k & 3));
        k++;
        out[k >> 2] |= (c & 0xFF) << (24 - 8 * (k & 3));
    }

    // ---- decoupled look-back over frame sizes: exclusive prefix = final byte offset ----
    if (t == 0) {
        const unsigned long long kAgg = 1ull << 62, kPre = 2ull << 62, kMask = (1ull << 62) - 1;
        unsigned long long mine = (unsigned long long)frame_bytes + (f == 0 ? (unsigned long long)P.hdr_bytes : 0ull);
        unsigned long long excl = 0;
        if (g == 0) {
            st_release_u64(&P.desc[0], kPre | mine);
        } else {
            st_release_u64(&P.desc[g], kAgg | mine);
            long long j = (long long)g - 1;
            for (;;) {
                unsigned long long d = ld_acquire_u64(&P.desc[j]);
                unsigned long long st = d >> 62;
                if (st == 0) { spin_pause(); continue; }
                excl += d & kMask;
                if (st == 2) break;
                j--;
            }
            st_release_u64(&P.desc[g], kPre | (excl + mine));
        }
        long long off = (long long)excl + (f == 0 ? P.hdr_bytes : 0);
        sh->frame_off = off;
        if (f == 0) {
            st_release_u64((unsigned long long*)&P.starts[s], excl);
            sh->stream_start = (long long)excl;
        } else {
            unsigned long long v;
            while ((long long)(v = ld_acquire_u64((const unsigned long long*)&P.starts[s])) < 0) spin_pause();
            sh->stream_start = (long long)v;
        }
        if (f == P.nframes - 1) P.ends[s] = off + frame_bytes;
        if (off + frame_bytes > P.out_capacity) { atom_or_global(P.err, kErrEncodeCollect); sh->frame_off = -1; }
    }
    sync();
    const long long off = sh->frame_off;
    if (off >= 0) {
        uint8_t* dst = P.out + off;
        // aligned 32-bit stores in the middle, byte stores at the ragged ends
        int a = (int)((uintptr_t)dst & 3);
        int head = a ? 4 - a : 0;
        if (head > frame_bytes) head = frame_bytes;
        if (t < head) dst[t] = (uint8_t)out_byte(out, t);
        int nwords = (frame_bytes - head) >> 2;
        uint32_t* dw = (uint32_t*)(dst + head);
        for (int w = t; w < nwords; w += kEncThreads) {
            int b = head + 4 * w;   // stream byte index of this word's first byte
            uint32_t hi = out[b >> 2], lo = out[(b >> 2) + 1];
            uint32_t v = funnel_l(lo, hi, 8u * (uint32_t)(b & 3));
            dw[w] = bswap32(v);
        }
        int tail0 = head + 4 * nwords;
        if (t < frame_bytes - tail0) dst[tail0 + t] = (uint8_t)out_byte(out, tail0 + t);

        // stream header (frame 0) and this frame's entry in the frame-size table
        uint8_t* sp = P.out + sh->stream_start;
        if (f == 0 && t == 0) {
            sp[0] = 'f'; sp[1] = 'L'; sp[2] = 'a'; sp[3] = 'C';
            sp[4] = 0x00; sp[5] = 0; sp[6] = 0; sp[7] = 34;
            uint8_t* si = sp + 8;
            for (int i = 0; i < 34; ++i) si[i] = 0;
            si[0] = si[2] = (uint8_t)(P.blocksize >> 8);
            si[1] = si[3] = (uint8_t)P.blocksize;
            const uint32_t sr = 44100;
            si[10] = (uint8_t)(sr >> 12); si[11] = (uint8_t)(sr >> 4);
            si[12] = (uint8_t)(((sr & 0xF) << 4) | ((nch - 1) << 1) | 1);
            unsigned long long ts = (unsigned long long)P.stream_size;
            if (ts >> 36) ts = 0;  // does not fit the 36-bit field: "unknown"
            si[13] = (uint8_t)(0xF0 | ((ts >> 32) & 0xF));
            si[14] = (uint8_t)(ts >> 24); si[15] = (uint8_t)(ts >> 16); si[16] = (uint8_t)(ts >> 8); si[17] = (uint8_t)ts;
            uint8_t* ap = sp + 42;
            uint32_t alen = 8 + 3 * (uint32_t)P.nframes;
            ap[0] = 0x82; ap[1] = (uint8_t)(alen >> 16); ap[2] = (uint8_t)(alen >> 8); ap[3] = (uint8_t)alen;
            ap[4] = 'f'; ap[5] = 'a'; ap[6] = 'B'; ap[7] = '2';
            uint32_t nf = (uint32_t)P.nframes;
            ap[8] = (uint8_t)(nf >> 24); ap[9] = (uint8_t)(nf >> 16); ap[10] = (uint8_t)(nf >> 8); ap[11] = (uint8_t)nf;
        }
        if (t == 0) {
            uint8_t* e = sp + 54 + 3 * f;
            e[0] = (uint8_t)(frame_bytes >> 16); e[1] = (uint8_t)(frame_bytes >> 8); e[2] = (uint8_t)frame_bytes;
        }
    }
}

}  // namespace fa
