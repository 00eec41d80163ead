"""GPU parity on the reference's orthotope elements (quads of order 1-5, hexes of order 1-2: ReferenceElement.cpp:624-627,885-1004),
the "hex variant" of SURVEY.md section 8d.  Same bars and the same comparison as tests/test_gpu_parity.py: topology, scatter indices
and CSR structure bit-exact, assembled entries within 1e-12, solution fields within 1e-10 of the oracle.  Orthotope cells take the
general kernel (hfx_generic.cuh) with Jacobians and normals evaluated at every cubature point (multilinear geometry)."""
import numpy as np
import pytest

from tests import helpers as H
from tests.test_gpu_parity import compare

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dim,order", [(2, 1), (2, 2), (2, 3), (2, 4), (2, 5), (3, 1), (3, 2)])
@pytest.mark.parametrize("perturb", [0.0, 0.2])
def test_laplace_box_mesh(dim, order, perturb):
    compare(H.make_case(dim, order, N=3 if dim == 3 else 4, perturb=perturb, geom="orthotope"))


@pytest.mark.parametrize("dim,order,model,diff,bc", [(2, 3, "cdrs", "scalar", "dirichlet"), (3, 2, "cdrs", "tensor", "dirichlet"),
                                                     (3, 2, "diffsrc", "scalar", "integrated"), (2, 4, "euler", "scalar", "dirichlet"),
                                                     (3, 1, "diffsrc", "const", "dirichlet"), (2, 2, "burgers", "scalar", "integrated")])
def test_models_on_orthotopes(dim, order, model, diff, bc):
    """Every in-scope model on perturbed (non-affine) quads / hexes, double-valued tau, curved interior nodes for convection."""
    compare(H.make_case(dim, order, N=3, perturb=0.15, model=model, diff=diff, bc=bc, tau_double=model not in ("laplace",), seed=23,
                        curved=0.03 if model == "cdrs" else 0.0, geom="orthotope"), solve=model != "burgers")


@pytest.mark.parametrize("dim,order", [(2, 2), (3, 2)])
def test_constant_solution_on_orthotopes(dim, order):
    """TestHDGSolver.cpp:16-100 restated on quads / hexes: tau = 1, Dirichlet = 3 => Solution = 3, Flux = 0, Trace = 3 (1e-12)."""
    case = H.make_case(dim, order, N=2, perturb=0.1, geom="orthotope")
    case["fields"]["Dirichlet"][case["topo"]["boundary"]] = 3.0
    s, fm, m = H.run_device(case, rtol=1e-15)
    assert np.abs(fm["Solution"].values - 3.0).max() < 1e-12
    assert np.abs(fm["Flux"].values).max() < 1e-11
    assert np.abs(fm["Trace"].values - 3.0).max() < 1e-12


@pytest.mark.parametrize("model,diff,bc,tau_double", [("laplace", "none", "dirichlet", False), ("laplace", "none", "integrated", False), ("diffsrc", "const", "dirichlet", True),
                                                      ("cdrs", "none", "dirichlet", True), ("euler", "none", "dirichlet", False)])
def test_structured_hexahedra_take_the_large_element_kernel(model, diff, bc, tau_double, monkeypatch):
    """Parallelepiped hexahedra of order 2 (every cell of a structured mesh) with D = c I: hdg_big_kernel<BigHexP2, 512> (hfx_big.cuh with the orthotope frame) against the
    oracle and against the general kernel; a perturbed (trilinear) mesh keeps the general kernel."""
    case = H.make_case(3, 2, N=3, perturb=0.0, model=model, diff=diff, bc=bc, tau_double=tau_double, seed=67, geom="orthotope")
    o, s, fm = compare(case)
    assert s.lastAssembleKernel() == "big"
    l1 = s.getLocal(); v1 = s.getCSR()[2].copy()
    monkeypatch.setenv("HFX_NO_BIG", "1")
    s2, fm2, _ = H.run_device(case)
    assert s2.lastAssembleKernel() == "general"
    assert H.rel_err(l1["S"], s2.getLocal()["S"]) < 1e-12 and H.rel_err(v1, s2.getCSR()[2]) < 1e-12
    monkeypatch.delenv("HFX_NO_BIG")
    s3, _, _ = H.run_device(H.make_case(3, 2, N=2, perturb=0.15, geom="orthotope"), solve=False)
    assert s3.lastAssembleKernel() == "general"
