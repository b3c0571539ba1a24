// fa_quant.h -- float <-> integer conversion with the reference's exact rounding sequence.
//
// Restates /root/reference/src/flacarray/libflacarray/utils.c:160-368 operation by operation: the
// same intermediate types (float vs double) at every step, no FMA contraction (explicit _rn
// intrinsics), and x86 `cvttsd2si` semantics for the out-of-range casts that C leaves undefined
// (the reference deliberately allows clipping, utils.c:211-217).
#pragma once
#include "fa_simt.h"

namespace fa {

// (int64_t)double as compiled for x86-64: out-of-range and NaN give INT64_MIN ("integer indefinite")
FA_D long long cvt_d2ll_x86(double v) {
    if (v >= -9223372036854775808.0 && v < 9223372036854775808.0) return (long long)v;
    return (long long)0x8000000000000000ull;
}
// (int32_t)double as compiled for x86-64 (cvttsd2si r32)
FA_D int32_t cvt_d2i_x86(double v) {
    if (v > -2147483649.0 && v < 2147483648.0) return (int32_t)v;
    return (int32_t)0x80000000u;
}

// utils.c:194-230: per-stream offset and gain from (min, max) and an optional quanta.
FA_D void quant_params_f32(float smin, float smax, bool have_q, float q, float* off_out, float* gain_out) {
    float off = (float)dmul(0.5, (double)fadd(smin, smax));
    float d1 = fsub(smin, off), d2 = fsub(smax, off);
    float amp = (d1 > d2) ? (float)dmul(1.01, (double)d1) : (float)dmul(1.01, (double)d2);
    float min_quanta = fdiv(amp, 2147483648.0f);  // (float)INT32_MAX == 2^31
    float sq = have_q ? q : min_quanta;
    long long nquant = cvt_d2ll_x86(ddiv((double)off, (double)sq));
    off = (float)dmul((double)sq, (double)nquant);
    float gain = (sq == 0.0f) ? 1.0f : (float)ddiv(1.0, (double)sq);
    *off_out = off;
    *gain_out = gain;
}

FA_D void quant_params_f64(double smin, double smax, bool have_q, double q, double* off_out, double* gain_out) {
    double off = dmul(0.5, dadd(smin, smax));
    double d1 = dsub(smin, off), d2 = dsub(smax, off);
    double amp = (d1 > d2) ? dmul(1.01, d1) : dmul(1.01, d2);
    double min_quanta = ddiv(amp, 9223372036854775808.0);  // (double)INT64_MAX == 2^63
    double sq = have_q ? q : min_quanta;
    long long nquant = cvt_d2ll_x86(ddiv(off, sq));
    off = dmul(sq, (double)nquant);
    double gain = (sq == 0.0) ? 1.0 : ddiv(1.0, sq);
    *off_out = off;
    *gain_out = gain;
}

// utils.c:232-240
FA_D int32_t quant_f32(float x, float off, float gain) {
    float st = fsub(x, off);
    double v = (double)fmul(gain, st);
    v = (st >= 0.0f) ? dadd(v, 0.5) : dsub(v, 0.5);
    return cvt_d2i_x86(v);
}
// utils.c:317-325
FA_D long long quant_f64(double x, double off, double gain) {
    double st = dsub(x, off);
    double v = dmul(gain, st);
    v = (st >= 0.0) ? dadd(v, 0.5) : dsub(v, 0.5);
    return cvt_d2ll_x86(v);
}

// utils.c:350-368 / :330-348: coeff = 1/gain (double division, then the storage type), mul then add.
FA_D float restore_coeff_f32(float gain) { return (float)ddiv(1.0, (double)gain); }
FA_D float restore_f32(int32_t x, float off, float coeff) { return fadd(off, fmul(coeff, (float)x)); }
FA_D double restore_coeff_f64(double gain) { return ddiv(1.0, gain); }
FA_D double restore_f64(long long x, double off, double coeff) { return dadd(off, dmul(coeff, (double)x)); }

}  // namespace fa
