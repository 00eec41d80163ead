import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


_UNPACKED = {}


def unpacked_fixtures(kind):
    """Directory holding the decompressed tests/golden/meshes/<kind>/*.gz fixtures (kind = "msh": Gmsh sources of the reference's
    regression meshes, "h5": a few of its HDF5 mesh files; stored gzip-compressed, unpacked once per session into a temporary directory)."""
    import atexit, gzip, shutil, tempfile
    if kind not in _UNPACKED:
        d = tempfile.mkdtemp(prefix="hfx_%s_" % kind)
        atexit.register(shutil.rmtree, d, True)
        src = os.path.join(ROOT, "tests", "golden", "meshes", kind)
        for f in sorted(os.listdir(src)):
            if f.endswith(".gz"):
                with gzip.open(os.path.join(src, f), "rb") as i, open(os.path.join(d, f[:-3]), "wb") as o:
                    shutil.copyfileobj(i, o)
        _UNPACKED[kind] = d
    return _UNPACKED[kind]


def load_mesh(name):
    z = np.load(os.path.join(ROOT, "tests", "golden", "meshes", name + ".npz"))
    return z["nodes"], z["cells"]


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import lib as O
    O.lib()
    return O


@pytest.fixture(scope="session", autouse=True)
def _built():
    # the C-ABI library must exist for both CPU (symbol / host-table tests) and GPU tests
    so = os.path.join(ROOT, "hyperfox_b200", "libhfx.so")
    if not os.path.exists(so):
        import __graft_entry__ as g
        g.build()
