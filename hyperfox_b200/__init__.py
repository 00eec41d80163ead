"""hyperfox_b200 -- B200-native (sm_100a) implementation of HyperFox's element-by-element HDG assembly path.

The compute path is libhfx.so (hand-written CUDA, C ABI in include/hfx.h); this package is the thin host-side
mirror of the reference's class surface (hfox.py) plus synthetic mesh generation (meshgen.py).
"""
from .capi import ErrorHandle, LIB_PATH, device_count  # noqa: F401
