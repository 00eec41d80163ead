"""The CG oracle (oracle/cg.py) against the reference's own known answers."""
import numpy as np
import pytest

from oracle import cg
from oracle.mesh import compute_faces
from oracle.refel import ReferenceElement
from tests.conftest import load_mesh


@pytest.mark.parametrize("dim", [1, 2])
@pytest.mark.parametrize("order", [1, 2, 3, 4, 5])
def test_diffusion_operator_monomial_energies(dim, order):
    """tests/unittests/operator/TestDiffusion.cpp:29-95: v = sum_d x_d^p at the nodes, v^T A v = dim * 2 p^2 / (2 p - 1) on the reference element and on a rotated,
    shifted copy (margin 1e-12)"""
    ore = ReferenceElement(dim, order, "simplex")
    v = (ore.nodes ** order).sum(axis=1)
    want = dim * 2.0 * order ** 2 / (2.0 * order - 1.0)
    assert abs(v @ cg.diffusion_matrix(ore, ore.nodes) @ v - want) < 1e-12
    if dim == 2:
        th = 1.0
        R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        X = ore.nodes @ R.T + 2.0
        assert abs(v @ cg.diffusion_matrix(ore, X) @ v - want) < 1e-12


def test_cg_solver_constant_state():
    """tests/unittests/solver/TestCGSolver.cpp: lightTri (order 1), LaplaceModel, Dirichlet = 3 on every boundary face => Solution = 3 (1e-12)"""
    nodes, cells = load_mesh("lightTri")
    ore = ReferenceElement(2, 1, "simplex")
    topo = compute_faces(cells, ore)
    o = cg.CGOracle(ore, nodes, cells, topo["faces"], topo["boundary"])
    o.assemble(np.full((topo["faces"].shape[0], 2), 3.0))
    assert np.abs(o.solve() - 3.0).max() < 1e-12


@pytest.mark.parametrize("dim,order,name", [(2, 2, "regression_dim-2_h-2e-1_ord-2"), (3, 1, "regression_dim-3_h-2e-1_ord-1"), (3, 3, "regression_dim-3_h-3e-1_ord-3")])
def test_cg_laplace_regression(dim, order, name):
    """tests/regression/CG/TestCGLaplace.cpp: u = sin x e^y (harmonic) as Dirichlet data, nodal l2 error under the reference's ceiling (1e-2); an affine u is
    reproduced to rounding at every order; the matrix is symmetric away from the Dirichlet rows and annihilates constants there"""
    nodes, cells = load_mesh(name)
    ore = ReferenceElement(dim, order, "simplex")
    topo = compute_faces(cells, ore)
    o = cg.CGOracle(ore, nodes, cells, topo["faces"], topo["boundary"])
    ana = np.sin(nodes[:, 0]) * np.exp(nodes[:, 1])
    o.assemble(ana[topo["faces"]])
    sol = o.solve()
    assert np.sqrt(((sol - ana) ** 2).sum() / (ana ** 2).sum()) < 1e-2
    lin = 1.0 + nodes @ np.arange(1, dim + 1)
    o.assemble(lin[topo["faces"]])
    assert np.abs(o.solve() - lin).max() < 1e-11
    bn = np.unique(topo["faces"][topo["boundary"]])
    inner = np.setdiff1d(np.arange(nodes.shape[0]), bn)
    A = o.A.toarray()
    assert np.abs(A[inner].sum(axis=1)).max() < 1e-11
    assert np.abs(A[np.ix_(inner, inner)] - A[np.ix_(inner, inner)].T).max() < 1e-12


def _intx2n(dim, n):
    """tests/TestUtils.h.in:88-96"""
    return 2.0 / (2.0 * n + 1) if dim < 3 else 1.0 / (2.0 * n + 3.0) + 1.0 / (2.0 * n + 1.0)


@pytest.mark.parametrize("dim", [1, 2, 3])
@pytest.mark.parametrize("order", [1, 2, 3])
def test_convection_and_mass_known_answers(dim, order):
    """tests/unittests/operator/TestConvection.cpp:26-48 (v = e_d, u = x_d^(p) / p ... : w^T C u = int x^(2(p-1))) and TestMass.cpp:24-41 (1^T M 1 = volume of the
    reference simplex, (x^p)^T M x^p = int x^2p) on the reference element"""
    ore = ReferenceElement(dim, order, "simplex")
    X = ore.nodes
    j = order - 1
    for d in range(dim):
        vel = np.zeros((ore.nNodes, dim)); vel[:, d] = 1.0
        u = X[:, d] ** (j + 1) / (j + 1.0); w = X[:, d] ** j
        assert abs(w @ cg.convection_matrix(ore, X, vel) @ u - _intx2n(dim, j)) < 1e-12
    M = cg.mass_matrix(ore, X)
    one = np.ones(ore.nNodes)
    assert abs(one @ M @ one - [2.0, 2.0, 4.0 / 3.0][dim - 1]) < 1e-12
    xn = X[:, 0] ** order
    assert abs(xn @ M @ xn - _intx2n(dim, order)) < 1e-12


def test_euler_apply_identity():
    """tests/unittests/operator/TestEuler.cpp: stiffness -> M + dt K, rhs -> M u_old + dt f"""
    ore = ReferenceElement(2, 2, "simplex")
    X = ore.nodes * 0.3 + 1.0
    K = cg.diffusion_matrix(ore, X); M = cg.mass_matrix(ore, X)
    rng = np.random.default_rng(0)
    f, u = rng.standard_normal(ore.nNodes), rng.standard_normal(ore.nNodes)
    A, F = cg.euler_apply(K, f, M, u, 0.01)
    assert np.abs(A - (M + 0.01 * K)).max() < 1e-15 and np.abs(F - (M @ u + 0.01 * f)).max() < 1e-15
