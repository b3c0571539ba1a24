/* See stream_encoder.h in this directory: all stand-in typedefs live there. */
#include "stream_encoder.h"
