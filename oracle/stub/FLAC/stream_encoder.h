/*
 * Minimal stand-in for <FLAC/stream_encoder.h>, written for this repo (not copied from libFLAC).
 * The reference's flacarray.h includes the libFLAC headers only to spell callback prototypes;
 * utils.c (the float<->int converters and the int64 split) uses no libFLAC symbol.  These opaque
 * typedefs let gcc compile /root/reference/src/flacarray/libflacarray/utils.c unchanged into
 * oracle/_ref/libfa_utils.so without libFLAC being installed.  TEST INFRASTRUCTURE ONLY.
 */
#ifndef ORACLE_STUB_FLAC_STREAM_ENCODER_H
#define ORACLE_STUB_FLAC_STREAM_ENCODER_H
#include <stdint.h>
#include <stddef.h>
typedef int FLAC__bool;
typedef uint8_t FLAC__byte;
typedef int32_t FLAC__int32;
typedef uint64_t FLAC__uint64;
typedef struct FLAC__StreamEncoder FLAC__StreamEncoder;
typedef struct FLAC__StreamDecoder FLAC__StreamDecoder;
typedef struct FLAC__Frame FLAC__Frame;
typedef int FLAC__StreamEncoderWriteStatus;
typedef int FLAC__StreamDecoderReadStatus;
typedef int FLAC__StreamDecoderWriteStatus;
typedef int FLAC__StreamDecoderErrorStatus;
typedef int FLAC__StreamDecoderSeekStatus;
typedef int FLAC__StreamDecoderTellStatus;
typedef int FLAC__StreamDecoderLengthStatus;
#endif
