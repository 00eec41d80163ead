"""Large-element kernel (hfx_big.cuh: 3-D order 4, straight-sided cells, D = c I) against the oracle and against the general kernel.

Same bars as tests/test_gpu_parity.py.  Every in-scope scalar model goes through it: Laplace (tau constant on each face: the face masses are scalar
multiples of the reference face mass), tau varying along the faces (cubature contraction of the tau mass), double-valued tau, sources, reaction,
convection (the configs[3] fields too), implicit Euler, both boundary models; meshes with a curved cell must fall back to the general kernel."""
import numpy as np
import pytest

from tests import helpers as H
from tests.test_gpu_parity import compare, TOL_ENTRIES, TOL_RECOVERY, TOL_SOLUTION

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("model,diff,bc,tau_double", [("laplace", "none", "dirichlet", False), ("laplace", "none", "integrated", False),
                                                      ("diffsrc", "none", "dirichlet", True), ("diffsrc", "const", "integrated", False),
                                                      ("cdrs", "none", "dirichlet", True), ("cdrs", "const", "dirichlet", False),
                                                      ("euler", "none", "dirichlet", False), ("euler", "const", "dirichlet", True)])
def test_big_kernel_matches_oracle(model, diff, bc, tau_double):
    o, s, fm = compare(H.make_case(3, 4, N=2, perturb=0.12, model=model, diff=diff, bc=bc, tau_double=tau_double, seed=23))
    assert s.lastAssembleKernel() == "big"


def test_big_kernel_configs3_fields():
    """BASELINE configs[3]: D = 1e-2, v = 4(-(y-1/2), x-1/2, 0), tau = |v.n| + D / sqrt(D dt), h = 1/8."""
    case = H.make_case(3, 4, N=2, perturb=0.1, model="cd", scale=0.25, seed=5)
    H.config4_fields(case)
    o, s, fm = compare(case)
    assert s.lastAssembleKernel() == "big"


@pytest.mark.parametrize("model", ["laplace", "cdrs"])
def test_big_kernel_matches_general_kernel(model, monkeypatch):
    case = H.make_case(3, 4, N=2, perturb=0.1, model=model, tau_double=model == "cdrs", seed=29)
    s1, fm1, _ = H.run_device(case)
    assert s1.lastAssembleKernel() == "big"
    l1 = s1.getLocal(); v1 = s1.getCSR()[2].copy(); sol1 = fm1["Solution"].values.copy()
    monkeypatch.setenv("HFX_NO_BIG", "1")
    s2, fm2, _ = H.run_device(case)
    assert s2.lastAssembleKernel() == "general"
    l2 = s2.getLocal()
    for name in ("S", "S0"):
        assert H.rel_err(l1[name], l2[name]) < TOL_ENTRIES, name
    for name in ("U", "Q", "U0", "Q0"):
        assert H.rel_err(l1[name], l2[name]) < TOL_RECOVERY, name
    assert H.rel_err(v1, s2.getCSR()[2]) < TOL_ENTRIES
    assert H.rel_err(sol1, fm2["Solution"].values) < TOL_SOLUTION


def test_curved_cell_falls_back_to_general_kernel():
    case = H.make_case(3, 4, N=2, perturb=0.1, model="laplace", curved=0.03, seed=31)
    o, s, fm = compare(case)
    assert s.lastAssembleKernel() == "general"


def test_big_kernel_reassembly_is_bit_reproducible():
    case = H.make_case(3, 4, N=2, perturb=0.1, model="cdrs", seed=37)
    s, fm, m = H.run_device(case, solve=False)
    v1 = s.getCSR()[2].copy(); r1 = s.getCSR()[3].copy()
    s.assemble()
    assert np.array_equal(v1, s.getCSR()[2]) and np.array_equal(r1, s.getCSR()[3])


@pytest.mark.parametrize("model", ["laplace", "cdrs", "euler"])
def test_recovery_by_recomputation(model):
    """HFX_RECOMPUTE_RECOVERY: U, Q are never stored; hfx_recover re-condenses each element and applies them out of shared memory.
    Same solution fields as the stored-operator recovery (to rounding: the sums run in another order) and as the oracle."""
    case = H.make_case(3, 4, N=2, perturb=0.1, model=model, tau_double=model == "cdrs", seed=41)
    o = H.run_oracle(case)
    s1, fm1, _ = H.run_device(case)
    s2, fm2, _ = H.run_device(case, recompute=True)
    assert s2.lastAssembleKernel() == "big"
    for name in ("Trace", "Solution", "Flux"):
        assert H.rel_err(fm2[name].values, fm1[name].values) < 1e-12, name
    assert H.rel_err(fm2["Solution"].values, o.sol.ravel()) < TOL_SOLUTION
    assert H.rel_err(fm2["Flux"].values, o.flux.ravel()) < TOL_SOLUTION
    with pytest.raises(Exception):
        s2.getLocal()


@pytest.mark.parametrize("model,diff,tau_double,expect", [("laplace", "none", False, "fused"), ("diffsrc", "none", True, "big"), ("cdrs", "none", True, "big"),
                                                          ("euler", "const", False, "big"), ("cdrs", "scalar", False, "fused")])
def test_order_3_dispatch_and_parity(model, diff, tau_double, expect):
    """3-D order 3 on straight-sided cells: Laplace-type models with a face-constant tau keep the all-reference path of the element-group kernel; a tau that varies
    along faces (tau_double draws random nodal values), convection, reaction or implicit Euler take hdg_big_kernel<3,3,256>; a diffusion FIELD keeps the element-group kernel."""
    o, s, fm = compare(H.make_case(3, 3, N=2, perturb=0.1, model=model, diff=diff, tau_double=tau_double, seed=61))
    assert s.lastAssembleKernel() == expect
