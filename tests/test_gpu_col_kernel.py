"""Order-2 tetrahedra through the column-per-lane kernel (hfx_col.cuh: one warp per element, one trace column per lane; Laplace-type models on straight-sided
cells) against the oracle and against the element-group kernel (HFX_COL=0)."""
import numpy as np
import pytest

from tests import helpers as H
from tests.test_gpu_parity import compare, TOL_ENTRIES, TOL_RECOVERY, TOL_SOLUTION

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("model,diff,bc,tau_double", [("laplace", "none", "dirichlet", False), ("laplace", "none", "integrated", False),
                                                      ("diffsrc", "none", "dirichlet", True), ("diffsrc", "const", "integrated", True),
                                                      ("diffsrc", "const", "dirichlet", False)])
def test_col_kernel_matches_oracle(model, diff, bc, tau_double):
    o, s, fm = compare(H.make_case(3, 2, N=3, perturb=0.12, model=model, diff=diff, bc=bc, tau_double=tau_double, seed=43))
    assert s.lastAssembleKernel() == "col"


def test_col_kernel_on_the_reference_gmsh_mesh():
    """the reference's own order-2 Gmsh tets (729 cells: not a multiple of the warps of a CTA)"""
    case = H.make_case(3, 2, mesh="regression_dim-3_h-2e-1_ord-2", model="diffsrc", bc="integrated", tau_double=True, seed=67)
    o, s, fm = compare(case)
    assert s.lastAssembleKernel() == "col"


@pytest.mark.parametrize("model", ["laplace", "diffsrc"])
def test_col_kernel_matches_element_group_kernel(model, monkeypatch):
    case = H.make_case(3, 2, N=3, perturb=0.1, model=model, tau_double=model == "diffsrc", seed=47)
    s1, fm1, _ = H.run_device(case)
    assert s1.lastAssembleKernel() == "col"
    l1 = s1.getLocal(); v1 = s1.getCSR()[2].copy(); r1 = s1.getCSR()[3].copy(); sol1 = fm1["Solution"].values.copy()
    monkeypatch.setenv("HFX_COL", "0")
    s2, fm2, _ = H.run_device(case)
    assert s2.lastAssembleKernel() == "fused"
    l2 = s2.getLocal()
    for name in ("S", "S0"):
        assert H.rel_err(l1[name], l2[name]) < TOL_ENTRIES, name
    for name in ("U", "Q", "U0", "Q0"):
        assert H.rel_err(l1[name], l2[name]) < TOL_RECOVERY, name
    assert H.rel_err(v1, s2.getCSR()[2]) < TOL_ENTRIES and H.rel_err(r1, s2.getCSR()[3]) < TOL_ENTRIES
    assert H.rel_err(sol1, fm2["Solution"].values) < TOL_SOLUTION


def test_other_models_and_curved_cells_keep_the_element_group_kernel():
    o, s, fm = compare(H.make_case(3, 2, N=2, perturb=0.1, model="cdrs", diff="scalar", seed=53))
    assert s.lastAssembleKernel() == "fused"
    o, s, fm = compare(H.make_case(3, 2, N=2, perturb=0.1, model="laplace", curved=0.05, seed=54))
    assert s.lastAssembleKernel() == "fused"


def test_col_kernel_reassembly_is_bit_reproducible():
    case = H.make_case(3, 2, N=3, perturb=0.1, model="diffsrc", seed=59)
    s, fm, m = H.run_device(case, solve=False)
    assert s.lastAssembleKernel() == "col"
    v1 = s.getCSR()[2].copy(); r1 = s.getCSR()[3].copy()
    s.assemble()
    assert np.array_equal(v1, s.getCSR()[2]) and np.array_equal(r1, s.getCSR()[3])


def test_col_kernel_fine_mesh():
    """elements of the benchmark's size (h = 1/69 for the order-2 point of the sweep): the refinement step of U keeps the entries at the parity bars"""
    o, s, fm = compare(H.make_case(3, 2, N=3, perturb=0.1, model="laplace", scale=3.0 / 69.0, seed=71))
    assert s.lastAssembleKernel() == "col"
