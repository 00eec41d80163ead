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
