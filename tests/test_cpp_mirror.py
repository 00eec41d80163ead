"""The C++ host mirror (include/hyperfox/*.h) of the reference's class surface, driven by the reference's own solver tests
rewritten against it (tests/cpp/test_hdg_path.cpp).  CPU: it compiles, links against libhfx.so and honours the call-order
contract without a device.  GPU: the full known-answer cases."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_hdg_path.cpp")
BIN = os.path.join(ROOT, "tests", "cpp", "test_hdg_path")
MESHES = ["lightTri", "lightTri2", "regression_dim-2_h-1e-1_ord-2", "regression_dim-3_h-2e-1_ord-3", "regression_dim-2_h-1e-1_ord-3", "regression_dim-2_h-2e-1_ord-2"]


def build_cpp_test():
    deps = [SRC, os.path.join(ROOT, "include", "hyperfox", "hyperfox.h"), os.path.join(ROOT, "include", "hfx.h"), os.path.join(ROOT, "hyperfox_b200", "libhfx.so")]
    if os.path.exists(BIN) and all(os.path.getmtime(BIN) >= os.path.getmtime(d) for d in deps):
        return BIN
    libdir = os.path.join(ROOT, "hyperfox_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Wno-comment", "-I", os.path.join(ROOT, "include", "hyperfox"), SRC, "-o", BIN,
                           "-L", libdir, "-lhfx", "-Wl,-rpath," + libdir, "-Wl,-rpath,/usr/local/cuda/lib64"])
    return BIN


@pytest.fixture(scope="module")
def mesh_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("meshes")
    for name in MESHES:
        z = np.load(os.path.join(ROOT, "tests", "golden", "meshes", name + ".npz"))
        nodes, cells = z["nodes"], z["cells"]
        with open(os.path.join(d, name + ".txt"), "w") as f:
            f.write("%d %d %d %d\n" % (nodes.shape[0], nodes.shape[1], cells.shape[0], cells.shape[1]))
            np.savetxt(f, nodes, fmt="%.17g")
            np.savetxt(f, cells, fmt="%d")
    import shutil
    from tests.conftest import unpacked_fixtures
    for name in ("regression_dim-2_h-2e-1", "regression_dim-3_h-2e-1"):
        shutil.copyfile(os.path.join(unpacked_fixtures("msh"), name + ".msh"), os.path.join(d, name + ".msh"))
    for name in ("regression_dim-2_h-2e-1_ord-2", "regression_dim-3_h-2e-1_ord-3", "lightTri2", "fieldTest"):
        shutil.copyfile(os.path.join(unpacked_fixtures("h5"), name + ".h5"), os.path.join(d, name + ".h5"))
    return str(d)


def run(mesh_dir, section):
    p = subprocess.run([build_cpp_test(), mesh_dir, section], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    return p.stdout


def test_cpp_mirror_builds_and_call_order_contract(mesh_dir):
    """TestHDGSolver.cpp:37-80: every step throws before its prerequisite (no device needed up to initialize())."""
    out = run(mesh_dir, "contract")
    assert " 0 failed" in out


def test_cpp_mirror_gmsh_io_regenerates_reference_fixtures(mesh_dir):
    """GmshIo (Io mirror) -> hfx_host_read_msh / hfx_host_high_order_mesh: the reference's .h5 fixtures from their .msh sources (host only)."""
    out = run(mesh_dir, "meshio")
    assert " 0 failed" in out, out


def test_cpp_mirror_non_linear_wrapper(mesh_dir):
    """TestNonLinearWrapper.cpp restated against the C++ mirror (host only)."""
    out = run(mesh_dir, "nlw")
    assert " 0 failed" in out, out


def test_cpp_mirror_partitioner(mesh_dir):
    """tests/parallel/TestZoltanPartitioner.cpp restated against the mirror's Partitioner (host C++ plan behind the C ABI; every rank built in one process)."""
    out = run(mesh_dir, "partitioner")
    assert " 0 failed" in out, out


@pytest.mark.gpu
def test_cpp_mirror_reference_tests(mesh_dir):
    out = run(mesh_dir, "all")
    assert " 0 failed" in out, out
