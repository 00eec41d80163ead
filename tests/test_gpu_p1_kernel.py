"""Linear-tet kernels (hfx_p1.cuh; Laplace-type models on straight-sided cells) against the oracle and against the element-group kernel:
HFX_P1=1 (the default) = sixteen lanes per element, one trace column per lane; HFX_P1=2 = one thread per element (kept for comparison)."""
import numpy as np
import pytest

from tests import helpers as H
from tests.test_gpu_parity import compare, TOL_ENTRIES, TOL_RECOVERY, TOL_SOLUTION

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("model,diff,bc,tau_double", [("laplace", "none", "dirichlet", False), ("laplace", "none", "integrated", False),
                                                      ("diffsrc", "none", "dirichlet", True), ("diffsrc", "const", "integrated", True),
                                                      ("diffsrc", "const", "dirichlet", False)])
@pytest.mark.parametrize("variant", ["1", "2"])
def test_p1_kernel_matches_oracle(model, diff, bc, tau_double, variant, monkeypatch):
    monkeypatch.setenv("HFX_P1", variant)
    o, s, fm = compare(H.make_case(3, 1, N=3, perturb=0.12, model=model, diff=diff, bc=bc, tau_double=tau_double, seed=43))
    assert s.lastAssembleKernel() == "p1"


def test_p1_kernel_is_the_default_on_linear_tets():
    o, s, fm = compare(H.make_case(3, 1, N=4, perturb=0.1, model="laplace", seed=61))
    assert s.lastAssembleKernel() == "p1"


def test_p1_kernel_ragged_tail():
    """the reference's own Gmsh tets (729 cells: neither a multiple of the sixteen elements of a CTA nor of the two of a warp)"""
    case = H.make_case(3, 1, mesh="regression_dim-3_h-2e-1_ord-1", model="diffsrc", bc="integrated", tau_double=True, seed=67)
    assert case["cells"].shape[0] % 2 == 1
    o, s, fm = compare(case)
    assert s.lastAssembleKernel() == "p1"


@pytest.mark.parametrize("model", ["laplace", "diffsrc"])
def test_p1_kernel_matches_element_group_kernel(model, monkeypatch):
    case = H.make_case(3, 1, N=3, perturb=0.1, model=model, tau_double=model == "diffsrc", seed=47)
    monkeypatch.setenv("HFX_P1", "1")
    s1, fm1, _ = H.run_device(case)
    assert s1.lastAssembleKernel() == "p1"
    l1 = s1.getLocal(); v1 = s1.getCSR()[2].copy(); r1 = s1.getCSR()[3].copy(); sol1 = fm1["Solution"].values.copy()
    monkeypatch.setenv("HFX_NO_P1", "1")
    s2, fm2, _ = H.run_device(case)
    assert s2.lastAssembleKernel() == "fused"
    l2 = s2.getLocal()
    for name in ("S", "S0"):
        assert H.rel_err(l1[name], l2[name]) < TOL_ENTRIES, name
    for name in ("U", "Q", "U0", "Q0"):
        assert H.rel_err(l1[name], l2[name]) < TOL_RECOVERY, name
    assert H.rel_err(v1, s2.getCSR()[2]) < TOL_ENTRIES and H.rel_err(r1, s2.getCSR()[3]) < TOL_ENTRIES
    assert H.rel_err(sol1, fm2["Solution"].values) < TOL_SOLUTION


def test_other_models_on_linear_tets_keep_the_element_group_kernel(monkeypatch):
    monkeypatch.setenv("HFX_P1", "1")
    o, s, fm = compare(H.make_case(3, 1, N=2, perturb=0.1, model="cdrs", diff="scalar", seed=53))
    assert s.lastAssembleKernel() == "fused"


def test_p1_kernel_reassembly_is_bit_reproducible(monkeypatch):
    monkeypatch.setenv("HFX_P1", "1")
    case = H.make_case(3, 1, N=3, perturb=0.1, model="diffsrc", seed=59)
    s, fm, m = H.run_device(case, solve=False)
    v1 = s.getCSR()[2].copy(); r1 = s.getCSR()[3].copy()
    s.assemble()
    assert np.array_equal(v1, s.getCSR()[2]) and np.array_equal(r1, s.getCSR()[3])
