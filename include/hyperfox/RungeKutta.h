/* RungeKutta.h -- same include name as the reference header; the class lives in hyperfox.h */
#include "hyperfox.h"
