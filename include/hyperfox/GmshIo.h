/* GmshIo.h -- same include style as the reference headers; the class lives in hyperfox.h (C++ mirror over the C ABI of libhfx.so) */
#include "hyperfox.h"
