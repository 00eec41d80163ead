"""The C-ABI library loads on a CPU-only box and exports every symbol include/hfx.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

from hyperfox_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "hfx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hfx_[a-z0-9_]+)\s*\(", src)))


def test_header_matches_binding_list():
    assert header_symbols() == sorted(capi.SYMBOLS)


def test_library_exports_every_symbol():
    lib = ctypes.CDLL(capi.LIB_PATH)
    for s in header_symbols():
        assert hasattr(lib, s), s


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to create a context (it must never fall back to the oracle or the CPU)."""
    import pytest
    if capi.device_count() > 0:
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    assert capi.lib().hfx_ctx_create(0, ctypes.byref(h)) != 0
    assert b"no CPU fallback" in capi.lib().hfx_last_error(None)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "hyperfox_b200")
    for dp, dn, fn in os.walk(pkg):
        for f in fn:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".inc")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                for bad in ("from oracle", "import oracle", "liboracle", "oracle.h", "oracle/src"):
                    assert bad not in txt, (f, bad)
