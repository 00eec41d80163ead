/* RKType.h -- same include name as the reference header; the enum lives in hyperfox.h */
#include "hyperfox.h"
