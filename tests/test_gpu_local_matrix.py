"""The per-element Model surface on the DEVICE (FEModel::compute / getLocalMatrix / getLocalRHS, src/model/FEModel.h:43-78): the reference's
model tests restated against what the GPU assembles before the static condensation (hfx_get_local_matrix).

* tests/unittests/model/TestHDGLaplaceModel.cpp: call-order contract, model matrix = HDGBase + HDGDiffusion on the reference element
  (the oracle's operators, which tests/test_oracle_operators.py pins on the reference's TestHDGBase / TestHDGDiffusion identities), zero right-hand side;
* the other in-scope models on a physical (perturbed) element against the oracle's local system: HDGDiffusionSource with a tensor field,
  HDGConvectionDiffusionReactionSource, implicit Euler, HDGBurgersModel (nDOF = dim, Newton-linearised HDGUNabU);
* HDGSolver.getLocalMatrix(iEl) on a mesh: every element against the oracle, and the block identities of TestHDGBase (S_qq = M (x) I, S_ul = -S_lu^T for Laplace).
"""
import numpy as np
import pytest

from hyperfox_b200 import hfox
from hyperfox_b200.capi import ErrorHandle
from oracle import lib as O
from oracle.refel import ReferenceElement as OracleRefEl
from tests import helpers as H

pytestmark = pytest.mark.gpu

TOL = 1e-12


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("dim,order", [(2, 1), (2, 2), (2, 3), (2, 4), (2, 5), (3, 1), (3, 2), (3, 3), (3, 4), (3, 5)])
def test_hdg_laplace_model_on_the_reference_element(dim, order):
    re = hfox.ReferenceElement(dim, order, "simplex")
    mod = hfox.HDGLaplaceModel(re)
    with pytest.raises(ErrorHandle):
        mod.compute()
    with pytest.raises(ErrorHandle):
        mod.setFieldMap({})
    taus = np.ones(re.getNumFaces() * re.getFaceElement().getNumNodes())
    mod.setFieldMap({"Tau": taus})
    with pytest.raises(ErrorHandle):
        mod.compute()
    mod.setElementNodes(re.getNodes())
    with pytest.raises(ErrorHandle):
        mod.compute()                      # not allocated yet
    mod.allocate(1)
    mod.compute()
    ore = OracleRefEl(dim, order)
    rc = O.RefElC(ore)
    nodes = np.asarray(re.getNodes())
    test = O.op_base(rc, 1, nodes, taus) + O.op_diffusion(rc, 1, nodes)
    d = test - mod.getLocalMatrix()
    assert (d.T @ d).sum() < TOL            # the reference's own check (TestHDGLaplaceModel.cpp)
    assert rel(mod.getLocalMatrix(), test) < TOL
    assert abs(mod.getLocalRHS().sum()) < TOL


def _physical_element(ore, seed):
    rng = np.random.default_rng(seed)
    dim = ore.dim
    A = np.eye(dim) + 0.25 * rng.standard_normal((dim, dim))
    if np.linalg.det(A) < 0:
        A[:, 0] *= -1
    return 0.3 * (ore.nodes @ A.T) + rng.standard_normal(dim)[None, :]


@pytest.mark.parametrize("dim,order", [(2, 3), (3, 2), (3, 4)])
def test_diffusion_source_and_cdrs_models_on_a_physical_element(dim, order):
    re = hfox.ReferenceElement(dim, order, "simplex")
    ore = OracleRefEl(dim, order); rc = O.RefElC(ore)
    nodes = _physical_element(ore, 3)
    rng = np.random.default_rng(7)
    nN, l = re.getNumNodes(), re.getNumFaces() * re.getFaceElement().getNumNodes()
    tau = 0.5 + rng.random(l)
    Araw = rng.standard_normal((nN, dim, dim)) * 0.2
    D = (np.eye(dim)[None] + Araw @ Araw.transpose(0, 2, 1)).transpose(0, 2, 1).reshape(nN, dim * dim)
    vel = rng.standard_normal((nN, dim))
    src = lambda x: np.exp(-sum((xi - 0.1) ** 2 for xi in x))
    reac = lambda x: 1.0 + 0.5 * x[0]
    xip = ore.ipShape @ nodes
    # HDGDiffusionSource with a tensor field
    mod = hfox.HDGDiffusionSource(re)
    mod.allocate(1); mod.setSourceFunction(src)
    mod.setElementNodes(nodes); mod.setFieldMap({"Tau": tau, "DiffusionTensor": D})
    mod.compute()
    A, F = O.local_system(rc, O.make_model(1, O.OP_DIFFUSION | O.OP_SOURCE, dim * dim), nodes=nodes, tau=tau, diff=D, srcIP=np.array([src(p) for p in xip]))
    assert rel(mod.getLocalMatrix(), A) < TOL and rel(mod.getLocalRHS(), F) < TOL
    # HDGConvectionDiffusionReactionSource: velocity + scalar diffusion + reaction + source
    mod = hfox.HDGConvectionDiffusionReactionSource(re)
    mod.allocate(1); mod.setSourceFunction(src); mod.setReactionFunction(reac)
    Ds = 0.5 + rng.random((nN, 1))
    mod.setElementNodes(nodes); mod.setFieldMap({"Tau": tau, "DiffusionTensor": Ds, "Velocity": vel})
    mod.compute()
    A, F = O.local_system(rc, O.make_model(1, O.OP_DIFFUSION | O.OP_CONVECTION | O.OP_REACTION | O.OP_SOURCE, 1), nodes=nodes, tau=tau, diff=Ds, vel=vel,
                          srcIP=np.array([src(p) for p in xip]), reacIP=np.array([reac(p) for p in xip]))
    assert rel(mod.getLocalMatrix(), A) < TOL and rel(mod.getLocalRHS(), F) < TOL
    # implicit Euler on HDGDiffusionSource (Euler.cpp:18-37): dt * operators + mass, dt * source + mass * old solution
    mod = hfox.HDGDiffusionSource(re)
    ts = hfox.Euler(re); ts.setTimeStep(0.05)
    mod.setTimeScheme(ts)
    mod.allocate(1); mod.setSourceFunction(src)
    old = rng.random(nN)
    mod.setElementNodes(nodes); mod.setFieldMap({"Tau": tau, "Solution": old})
    mod.compute()
    A, F = O.local_system(rc, O.make_model(1, O.OP_DIFFUSION | O.OP_SOURCE, 0, O.TS_EULER_IMPLICIT, 0.05), nodes=nodes, tau=tau, solOld=old,
                          srcIP=np.array([src(p) for p in xip]))
    assert rel(mod.getLocalMatrix(), A) < TOL and rel(mod.getLocalRHS(), F) < TOL


@pytest.mark.parametrize("dim,order", [(2, 2), (2, 4), (3, 2)])
def test_burgers_model_on_a_physical_element(dim, order):
    """TestHDGBurgersModel.cpp: model = Base + HDGUNabU (+ Diffusion), right-hand side = the UNabU residual (+ one source per component)."""
    re = hfox.ReferenceElement(dim, order, "simplex")
    ore = OracleRefEl(dim, order); rc = O.RefElC(ore)
    nodes = _physical_element(ore, 5)
    rng = np.random.default_rng(11)
    nN, nFc, nNf = re.getNumNodes(), re.getNumFaces(), re.getFaceElement().getNumNodes()
    blk = 2.0 * np.eye(dim)[None] + 0.3 * rng.random((nFc * nNf, dim, dim))
    sol = 0.3 * rng.standard_normal((nN, dim)); tr = 0.3 * rng.standard_normal((nFc * nNf, dim))
    Ds = 0.5 + rng.random((nN, 1))
    mod = hfox.HDGBurgersModel(re)
    with pytest.raises(ErrorHandle):
        mod.allocate(1)                     # nDOF must equal the dimension
    mod.allocate(dim)
    mod.setElementNodes(nodes); mod.setFieldMap({"Tau": blk.reshape(-1), "BufferSolution": sol, "Trace": tr, "DiffusionTensor": Ds})
    mod.compute()
    A, F = O.local_system(rc, O.make_model(dim, O.OP_UNABU | O.OP_DIFFUSION, 1), nodes=nodes, tau=blk.reshape(-1), diff=Ds, bufSol=sol.ravel(), trace=tr.ravel())
    assert rel(mod.getLocalMatrix(), A) < TOL and rel(mod.getLocalRHS(), F) < TOL


def test_solver_local_matrices_on_a_mesh_and_block_identities():
    """HDGSolver.getLocalMatrix(iEl) for every element of a perturbed 3-D mesh against the oracle, and TestHDGBase's block identities on the device blocks."""
    case = H.make_case(3, 3, N=2, perturb=0.1, model="laplace", seed=2)
    s, fm, m = H.run_device(case, solve=False)
    ore = case["ore"]; rc = O.RefElC(ore)
    nN, dim = ore.nNodes, 3
    u, q, l, n = O.sizes(rc, 1)
    from tests.test_gpu_fine_mesh import element_tau
    M = np.einsum("p,pi,pj->ij", ore.ipWeights, ore.ipShape, ore.ipShape)
    for e in range(0, case["cells"].shape[0], 5):
        A, F = s.getLocalMatrix(e)
        Ao, Fo = O.local_system(rc, O.make_model(1, O.OP_DIFFUSION), nodes=case["nodes"][case["cells"][e]], tau=element_tau(case, e))
        assert rel(A, Ao) < TOL and np.abs(F - Fo).max() < TOL
        # S_qq = M (x) I_dim with the physical mass matrix (HDGBase.cpp:152): straight-sided element, M = detJ * M_ref
        X = case["nodes"][case["cells"][e]]
        detJ = abs(np.linalg.det(0.5 * (X[1:4] - X[0])))
        Sqq = A[u:u + q, u:u + q]
        assert rel(Sqq, np.kron(detJ * M, np.eye(dim))) < 1e-11
        # S_lu = tau-mass on the faces, S_ul = -S_lu^T (HDGBase.cpp:125-131), S_ll = -tau-mass (symmetric)
        assert rel(A[:u, u + q:], -A[u + q:, :u].T) < TOL
        assert rel(A[u + q:, u + q:], A[u + q:, u + q:].T) < TOL
