"""Pins the oracle's operators (oracle/src/oracle.cpp) with the reference's own operator/model known-answer tests.

Restates tests/unittests/operator/TestOperator.cpp:43-163 (analytic J, J^-1, det J on affine, offset and non-linear elements),
TestHDGBase.cpp:59-133 (block identities vs Mass / Convection / reference normals), TestHDGDiffusion.cpp:38-102 (vs Convection +
face Mass, incl. D = 3 I), TestHDGConvection.cpp, TestReaction/TestSource/TestMass, TestEuler, and the model = sum-of-operators
tests of tests/unittests/model/*.cpp, all on the reference element itself for dim 2-3, order 1-5."""
import numpy as np
import pytest

from oracle import lib as O
from oracle.refel import ReferenceElement

DIMS_ORDERS = [(d, p) for d in (2, 3) for p in (1, 2, 3, 4, 5)]


def setup(dim, order):
    re = ReferenceElement(dim, order)
    return re, O.RefElC(re)


@pytest.mark.parametrize("dim,order", [(2, 3), (3, 4)])
def test_jacobians_analytic(dim, order):
    """TestOperator.cpp:43-163. Note the reference stores J(r, m) = d x_m / d xi_r, i.e. the transpose of the map's matrix."""
    re, rc = setup(dim, order)
    Jac = np.array([[3.0, 2.0], [1.0, 4.0]]) if dim == 2 else np.diag([1.0, 1.0, 5.0]) + np.array([[0, 0, 0], [0, 0, 0], [0, 0, 0.0]])
    if dim == 3:
        Jac = np.array([[1.0, 0, 0], [0, 1.0, 0], [0, 0, 5.0]])
    for off in (np.zeros(dim), np.arange(1, dim + 1, dtype=float)):
        if off.any():
            Jac = Jac + np.pad(np.array([[2.0, -1.0], [1.0, 3.0]]), ((0, dim - 2), (0, dim - 2)))
        nodes = re.nodes @ Jac.T + off
        jac, inv, dV, nrm = O.element_geometry(rc, nodes)
        for ip in range(re.nIP):
            assert np.abs(jac[ip] - Jac.T).max() < 1e-12
            assert np.abs(inv[ip] - np.linalg.inv(Jac.T)).max() < 1e-12
            assert abs(dV[ip] / re.ipWeights[ip] - np.linalg.det(Jac)) < 1e-12
    # non-linear element x_j = xi_j^ord (ord = dim as in the reference test): J = diag(ord * xi^(ord-1)) at every cubature point
    o = dim
    jac, inv, dV, nrm = O.element_geometry(rc, re.nodes ** o)
    for ip in range(re.nIP):
        ana = np.diag(o * re.ipCoords[ip] ** (o - 1))
        assert np.abs(jac[ip] - ana.T).max() < 1e-12
        assert abs(dV[ip] / re.ipWeights[ip] - np.linalg.det(ana)) < 1e-12


def face_masses(re, rc, dV):
    return [O.op_mass(re.faceElement.ipShape, dV[re.nIP + f * rc.nIPf: re.nIP + (f + 1) * rc.nIPf]) for f in range(rc.nFc)]


REFN = {2: [[0, -1], [2 ** -0.5] * 2, [-1, 0]], 3: [[0, -1, 0], [3 ** -0.5] * 3, [-1, 0, 0], [0, 0, -1]]}


@pytest.mark.parametrize("dim,order", DIMS_ORDERS)
def test_hdg_base_blocks(dim, order):
    """TestHDGBase.cpp:59-133 with tau = 1 on the reference element."""
    re, rc = setup(dim, order)
    nN, nNf, nFc = rc.nN, rc.nNf, rc.nFc
    A = O.op_base(rc, 1, re.nodes, np.ones(nFc * nNf))
    jac, inv, dV, nrm = O.element_geometry(rc, re.nodes)
    M = O.op_mass(re.ipShape, re.ipWeights)
    sQ, sL = nN, nN * (dim + 1)
    assert np.abs(A[sQ:sL, sQ:sL] - np.kron(M, np.eye(dim))).max() < 1e-12          # Sqq = M (x) I
    # Squ from the convection operator with unit velocities (:75-88): Squ[(l,k), :] = C_k[:, l]^T
    for k in range(dim):
        vel = np.zeros((nN, dim)); vel[:, k] = 1.0
        C = -O.op_convection(rc, 1, re.nodes, vel)[:nN, :nN].T    # HDGConvection puts -C^T into Suu; here no face part since v.n mass goes to Sul/Sll
        # C[k_,l] = sum dV (e_k . grad phi_l) phi_k_ ; Squ[(l,k), j] = sum dV (grad phi_l)_k phi_j = C[j, l]
        assert np.abs(A[sQ + k:sL:dim, :nN] - C.T).max() < 1e-12
    Mf = face_masses(re, rc, dV)
    fn = np.array(re.faceNodes)
    Slu = np.zeros((nFc * nNf, nN)); Suu = np.zeros((nN, nN)); Sql = np.zeros((nN * dim, nFc * nNf))
    for f in range(nFc):
        blk = slice(sL + f * nNf, sL + (f + 1) * nNf)
        assert np.abs(A[blk, blk] + Mf[f]).max() < 1e-12                              # Sll = -face mass
        Slu[f * nNf:(f + 1) * nNf, fn[f]] += Mf[f]
        Suu[np.ix_(fn[f], fn[f])] += Mf[f]
        for d in range(dim):
            Sql[fn[f] * dim + d, f * nNf:(f + 1) * nNf] -= Mf[f] * REFN[dim][f][d]
    assert np.abs(A[sL:, :nN] - Slu).max() < 1e-12
    assert np.abs(A[:nN, sL:] + Slu.T).max() < 1e-12
    assert np.abs(A[:nN, :nN] - Suu).max() < 1e-12
    assert np.abs(A[sQ:sL, sL:] - Sql).max() < 1e-12
    assert np.abs(A[sL:, sQ:sL]).max() == 0 and np.abs(A[:nN, sQ:sL]).max() == 0     # Slq = Suq = 0 in the base operator


@pytest.mark.parametrize("dim,order", DIMS_ORDERS)
def test_hdg_diffusion_vs_convection_and_face_mass(dim, order):
    """TestHDGDiffusion.cpp:38-102 incl. the D = 3 I variant."""
    re, rc = setup(dim, order)
    nN, nNf, nFc = rc.nN, rc.nNf, rc.nFc
    jac, inv, dV, nrm = O.element_geometry(rc, re.nodes)
    Mf = face_masses(re, rc, dV)
    fn = np.array(re.faceNodes)
    sQ, sL = nN, nN * (dim + 1)
    Suq = np.zeros((nN, nN * dim)); Slq = np.zeros((nFc * nNf, nN * dim))
    for k in range(dim):
        vel = np.zeros((nN, dim)); vel[:, k] = 1.0
        C = -O.op_convection(rc, 1, re.nodes, vel)[:nN, :nN].T
        Suq[:, k::dim] += C.T
    for f in range(nFc):
        for d in range(dim):
            Suq[np.ix_(fn[f], fn[f] * dim + d)] -= Mf[f] * REFN[dim][f][d]
            Slq[f * nNf:(f + 1) * nNf, fn[f] * dim + d] -= Mf[f] * REFN[dim][f][d]
    for scale, diff, comps in ((1.0, None, 0), (3.0, 3.0 * np.ones(nN), 1), (3.0, np.tile((3.0 * np.eye(dim)).ravel(), (nN, 1)), dim * dim)):
        A = O.op_diffusion(rc, 1, re.nodes, diff, comps)
        T = A.copy()
        T[:nN, sQ:sL] -= scale * Suq
        T[sL:, sQ:sL] -= scale * Slq
        assert (T * T).sum() < 1e-12


@pytest.mark.parametrize("dim,order", DIMS_ORDERS)
def test_models_are_sums_of_operators_and_structure(dim, order):
    """tests/unittests/model/TestHDGLaplaceModel.cpp:46-86, TestHDGConvectionDiffusionReactionSource.cpp:102-143 (incl. Euler): the local
    matrix is the sum of the operator matrices; also the sparsity structure the CUDA kernel relies on (S_qq = M (x) I, face coupling)."""
    re, rc = setup(dim, order)
    rng = np.random.default_rng(dim * 10 + order)
    nN, nNf, nFc = rc.nN, rc.nNf, rc.nFc
    nodes = re.nodes * 0.4 + 0.01 * rng.standard_normal(re.nodes.shape)
    tau = 0.5 + rng.random(nFc * nNf); vel = rng.standard_normal((nN, dim)); diff = 0.5 + rng.random(nN)
    src = rng.random(rc.nIP); reac = rng.random(rc.nIP); sold = rng.random(nN)
    base = O.op_base(rc, 1, nodes, tau); dif = O.op_diffusion(rc, 1, nodes, diff, 1); conv = O.op_convection(rc, 1, nodes, vel)
    A, F = O.local_system(rc, O.make_model(1, O.OP_DIFFUSION), nodes=nodes, tau=tau)
    assert ((A - base - O.op_diffusion(rc, 1, nodes)) ** 2).sum() < 1e-12 and np.abs(F).max() == 0
    md = O.make_model(1, O.OP_DIFFUSION | O.OP_CONVECTION | O.OP_REACTION | O.OP_SOURCE, diffComps=1)
    A, F = O.local_system(rc, md, nodes=nodes, tau=tau, diff=diff, vel=vel, srcIP=src, reacIP=reac)
    jac, inv, dV, nrm = O.element_geometry(rc, nodes)
    R = O.op_mass(re.ipShape, reac * dV[:rc.nIP])
    ref = base + dif + conv
    ref[:nN, :nN] += R
    assert ((A - ref) ** 2).sum() < 1e-12
    assert np.abs(F[:nN] - re.ipShape.T @ (src * dV[:rc.nIP])).max() < 1e-13 and np.abs(F[nN:]).max() == 0
    # implicit Euler as coded (Euler.cpp:18-37, hook HDGModel.cpp:38-47): the u rows Su, Fu are scaled by dt, then Suu += M, Fu += M u_old
    # (TestEuler.cpp:70-76: backEuler.apply(conv, 1) == mass + dt*conv, mass*sol + dt*1)
    dt = 0.1
    mdE = O.make_model(1, O.OP_DIFFUSION | O.OP_SOURCE, diffComps=1, timeScheme=O.TS_EULER_IMPLICIT, dt=dt)
    AE, FE = O.local_system(rc, mdE, nodes=nodes, tau=tau, diff=diff, srcIP=src, solOld=sold)
    M = O.op_mass(re.ipShape, dV[:rc.nIP])
    refE = base + dif
    refE[:nN, :] *= dt
    refE[:nN, :nN] += M
    assert ((AE - refE) ** 2).sum() < 1e-12
    assert np.abs(FE[:nN] - (dt * (re.ipShape.T @ (src * dV[:rc.nIP])) + M @ sold)).max() < 1e-12 and np.abs(FE[nN:]).max() == 0
    # structure
    sQ, sL = nN, nN * (dim + 1)
    assert np.abs(A[sQ:sL, sQ:sL] - np.kron(O.op_mass(re.ipShape, dV[:rc.nIP]), np.eye(dim))).max() < 1e-13


@pytest.mark.parametrize("dim,order", [(2, 1), (2, 3), (3, 2), (3, 3)])
def test_condensation_qr_vs_lu_and_direct(dim, order):
    """HDGSolver.cpp:331-348: the Householder-QR path (what the reference does) agrees with LU and with a dense numpy solve."""
    re, rc = setup(dim, order)
    rng = np.random.default_rng(5)
    nodes = re.nodes * 0.3 + 0.01 * rng.standard_normal(re.nodes.shape)
    A, F = O.local_system(rc, O.make_model(1, O.OP_DIFFUSION | O.OP_SOURCE), nodes=nodes, tau=1 + rng.random(rc.nFc * rc.nNf), srcIP=rng.random(rc.nIP))
    u, q, l, n = O.sizes(rc, 1)
    qr = O.condense(u, q, l, A, F, 0); lu = O.condense(u, q, l, A, F, 1)
    for a, b in zip(qr, lu):
        assert np.abs(a - b).max() <= 1e-12 * max(1.0, np.abs(a).max())
    # direct: eliminate (u,q) from the dense system
    L = A[:u + q, :u + q]; Bm = A[:u + q, u + q:]; C = A[u + q:, :u + q]; D = A[u + q:, u + q:]
    X = np.linalg.solve(L, Bm)
    S = D - C @ X
    assert np.abs(S - qr[2]).max() < 1e-11 * np.abs(S).max()
    assert np.abs(-X[:u] - qr[0]).max() < 1e-10 * np.abs(X).max() and np.abs(-X[u:] - qr[1]).max() < 1e-10 * np.abs(X).max()


def _unabu_forms(dim, order, power):
    """u = (x^power, 0, ..) on the reference element, trace = u on the faces: (u^T A_uu u + u^T A_ul t, rhs_u . u) of HDGUNabU."""
    re = ReferenceElement(dim, order)
    rc = O.RefElC(re)
    nN, nD = re.nNodes, dim
    sol = np.zeros((nN, nD))
    sol[:, 0] = re.nodes[:, 0] ** power
    trace = sol[np.asarray(re.faceNodes)]                       # [nFc, nNf, nD]
    A, r = O.op_unabu(rc, nD, re.nodes, sol, trace)
    u_len = nN * nD
    u, t = sol.reshape(-1), trace.reshape(-1)
    sL = u_len * (dim + 1)                                      # the trace block starts after the u and q blocks (TestHDGUNabU.cpp:98)
    return float(u @ A[:u_len, :u_len] @ u + u @ A[:u_len, sL:sL + t.size] @ t), float(r[:u_len] @ u)


def _intx2n(dim, k):
    """tests/TestUtils.h.in:88-96: closed form of the integral the reference compares with."""
    return 2.0 / (2.0 * k + 1) if dim < 3 else 1.0 / (2.0 * k + 3.0) + 1.0 / (2.0 * k + 1.0)


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("order", [1, 2, 3, 4, 5])
def test_hdg_unabu_closed_forms(dim, order):
    """tests/unittests/operator/TestHDGUNabU.cpp:31-105 (u = x^(n+1) e_0, odd n+1, closed-form integrals, margin 1e-12) and :107-170
    (a constant state gives zero for both forms)."""
    for n in range(0, 2 * (order - 1) // 3):
        if (n + 1) % 2 == 0:
            continue
        quad, lin = _unabu_forms(dim, order, n + 1)
        ref = _intx2n(dim, (3 * n + 2) // 2)
        assert abs(quad / (2.0 * (n + 1)) - ref) < 1e-12, (n, quad, ref)
        assert abs(lin / (1.0 * (n + 1)) - ref) < 1e-12, (n, lin, ref)
    quad, lin = _unabu_forms(dim, order, 0)                      # constant solution (1, 0, ..)
    assert quad < 1e-12 and lin < 1e-12


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("order", [1, 2, 3])
def test_burgers_model_is_the_sum_of_its_operators(dim, order):
    """tests/unittests/model/TestHDGBurgersModel.cpp:105-138: HDGBurgersModel (nDOF = dim) = HDGBase + HDGUNabU + HDGDiffusion, local
    right-hand side = the HDGUNabU one; with Euler the u-rows are scaled by dt and the mass matrix enters (the same fields as the
    reference's test: tau = I, D = 3 I, Solution = BufferSolution = Trace = 2)."""
    re = ReferenceElement(dim, order)
    rc = O.RefElC(re)
    nN, nD, nFc, nNf = re.nNodes, dim, re.nFaces, re.faceElement.nNodes
    u, q, l, n = O.sizes(rc, nD)
    tau = np.tile(np.eye(nD).reshape(-1), (nFc, nNf, 1))                     # [nFc, nNf, nD*nD], col-major per node
    diff = np.tile((3.0 * np.eye(dim)).reshape(-1), (nN, 1))                 # [nN, dim*dim]
    sol = np.full((nN, nD), 2.0)
    trace = np.full((nFc, nNf, nD), 2.0)
    base = O.op_base(rc, nD, re.nodes, tau)
    conv, rhs = O.op_unabu(rc, nD, re.nodes, sol, trace)
    dif = O.op_diffusion(rc, nD, re.nodes, diff, dim * dim)
    f = dict(nodes=re.nodes, tau=tau, diff=diff, bufSol=sol, trace=trace)
    mask = O.OP_UNABU | O.OP_DIFFUSION
    A, F = O.local_system(rc, O.make_model(nD, mask, dim * dim), **f)
    ana = base + conv + dif
    assert ((ana - A) ** 2).sum() < 1e-12 and ((rhs - F) ** 2).sum() < 1e-12
    # Euler (TestHDGBurgersModel.cpp:117-138, Euler.cpp:28-32)
    dt = 1e-2
    A2, F2 = O.local_system(rc, O.make_model(nD, mask, dim * dim, O.TS_EULER_IMPLICIT, dt), solOld=sol.reshape(-1), **f)
    jac, inv, dV, nrm = O.element_geometry(rc, re.nodes)
    M1 = O.op_mass(re.ipShape, dV[:re.nIP])
    M = np.kron(M1, np.eye(nD))
    ana2 = ana.copy()
    ana2[:u] *= dt
    ana2[:u, :u] += M
    assert ((ana2 - A2) ** 2).sum() < 1e-12
    exp = rhs.copy()
    exp[:u] = dt * rhs[:u] + M @ sol.reshape(-1)
    assert ((exp - F2) ** 2).sum() < 1e-12
    # the reference's test compares the u-segment with M sol alone: true because the HDGUNabU right-hand side of a constant state vanishes
    assert np.abs(rhs[:u]).max() < 1e-12


@pytest.mark.parametrize("dim,order", DIMS_ORDERS)
def test_transport_model_is_base_plus_convection(dim, order):
    """tests/unittests/model/TestHDGTransport.cpp (src/model/HDGTransport.cpp:47-69): localMatrix = Base + Convection, zero right-hand side -- the operator
    descriptor HFX_OP_CONVECTION of the product's HDGTransport; with an implicit Euler step the u rows are scaled by dt and the mass terms added."""
    re, rc = setup(dim, order)
    rng = np.random.default_rng(dim * 100 + order)
    nN, nNf, nFc = rc.nN, rc.nNf, rc.nFc
    nodes = re.nodes * 0.3 + 0.01 * rng.standard_normal(re.nodes.shape)
    tau = 0.5 + rng.random(nFc * nNf); vel = rng.standard_normal((nN, dim)); sold = rng.random(nN)
    base = O.op_base(rc, 1, nodes, tau); conv = O.op_convection(rc, 1, nodes, vel)
    A, F = O.local_system(rc, O.make_model(1, O.OP_CONVECTION), nodes=nodes, tau=tau, vel=vel)
    assert ((A - base - conv) ** 2).sum() < 1e-12 and np.abs(F).max() == 0
    dt = 0.05
    AE, FE = O.local_system(rc, O.make_model(1, O.OP_CONVECTION, timeScheme=O.TS_EULER_IMPLICIT, dt=dt), nodes=nodes, tau=tau, vel=vel, solOld=sold)
    jac, inv, dV, nrm = O.element_geometry(rc, nodes)
    M = O.op_mass(re.ipShape, dV[:rc.nIP])
    ref = base + conv
    ref[:nN, :] *= dt
    ref[:nN, :nN] += M
    assert ((AE - ref) ** 2).sum() < 1e-12
    assert np.abs(FE[:nN] - M @ sold).max() < 1e-12 and np.abs(FE[nN:]).max() == 0
