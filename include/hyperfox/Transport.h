/* Transport.h -- same include name as the reference header; the class lives in hyperfox.h (C++ mirror over the C ABI of libhfx.so) */
#include "hyperfox.h"
