"""Pins the oracle's solver-level restatement (condensation, Dirichlet rows, CSR scatter, GMRES(30)+Jacobi, recovery) with the
reference's end-to-end known answers, and checks host-side product pieces that need no GPU (topology builder, mesh generator)."""
import numpy as np
import pytest

from hyperfox_b200 import capi, meshgen
from oracle import lib as O
from oracle.mesh import compute_faces
from oracle.refel import ReferenceElement
from tests import helpers as H
from tests.conftest import load_mesh


def test_hdgsolver_constant_solution_oracle():
    """tests/unittests/solver/TestHDGSolver.cpp:16-100 (lightTri2: 2 P2 triangles, tau = 1, Dirichlet = 3)."""
    case = H.make_case(2, 2, mesh="lightTri2")
    case["fields"]["Dirichlet"][:] = 3.0
    for lu in (0, 1):
        o = H.run_oracle(case, useLU=lu, rtol=1e-12)
        assert np.abs(o.sol - 3.0).max() < 1e-12 and np.abs(o.flux).max() < 1e-12 and np.abs(o.trace - 3.0).max() < 1e-12


@pytest.mark.parametrize("name,dim,order", [("regression_dim-2_h-3e-1_ord-1", 2, 1), ("regression_dim-2_h-2e-1_ord-2", 2, 2),
                                             ("regression_dim-2_h-1e-1_ord-3", 2, 3), ("regression_dim-3_h-2e-1_ord-1", 3, 1),
                                             ("regression_dim-3_h-2e-1_ord-2", 3, 2), ("regression_dim-3_h-2e-1_ord-3", 3, 3)])
def test_laplace_regression_ceiling(name, dim, order):
    """tests/regression/HDG/TestHDGLaplace.cpp:110-139: u = sin(x) e^y, error ceiling 1e-2 (rtol 1e-16 there; 1e-13 here)."""
    case = H.make_case(dim, order, mesh=name)
    o = H.run_oracle(case)
    ana = case["ana"][case["cells"]]
    assert np.sqrt(((o.sol - ana) ** 2).sum() / (ana ** 2).sum()) < 1e-2
    assert o.resnorm < 1e-10


def test_csr_pattern_counts():
    """Interior rows couple to (2 nFc - 1) faces, boundary rows to nFc faces (70 / 40 nnz per row at p = 3)."""
    case = H.make_case(3, 3, N=2, perturb=0.0)
    o = H.run_oracle(case, solve=False)
    t = 10
    nnz = np.diff(o.rowptr)
    isb = np.zeros(case["topo"]["faces"].shape[0], dtype=bool)
    isb[case["topo"]["boundary"]] = True
    assert np.all(nnz.reshape(-1, t)[isb] == 4 * t) and np.all(nnz.reshape(-1, t)[~isb] == 7 * t)
    for r in (0, nnz.size - 1):
        cols = o.colidx[o.rowptr[r]:o.rowptr[r + 1]]
        assert np.all(np.diff(cols) > 0)          # sorted, unique (PETSc AIJ)


@pytest.mark.parametrize("name,dim,order", [("lightTri2", 2, 2), ("regression_dim-2_h-1e-1_ord-4", 2, 4), ("regression_dim-3_h-3e-1_ord-5", 3, 5),
                                             ("regression_dim-3_h-2e-1_ord-3", 3, 3)])
def test_product_topology_matches_oracle(name, dim, order):
    """csrc/host/hfx_topology.cpp vs oracle/mesh.py (restating Mesh.cpp:183-274,377-537): bit exact, on reference meshes."""
    nodes, cells = load_mesh(name)
    a = compute_faces(cells, ReferenceElement(dim, order))
    b = capi.host_compute_faces(dim, order, cells)
    for k in ("faces", "cell2face", "face2cell", "boundary"):
        assert np.array_equal(a[k], b[k]), k
    # consistency checks of tests/unittests/mesh/TestMesh.cpp:194-312 (membership, not numbering)
    for F in range(b["faces"].shape[0]):
        c0 = b["face2cell"][F, 0]
        assert F in b["cell2face"][c0] and set(b["faces"][F]) <= set(cells[c0])


def test_lightTri2_topology_values():
    nodes, cells = load_mesh("lightTri2")
    tp = capi.host_compute_faces(2, 2, cells)
    assert tp["faces"].shape == (5, 3)
    assert tp["cell2face"].tolist() == [[0, 1, 2], [1, 3, 4]]
    assert tp["face2cell"].tolist() == [[0, -1], [0, 1], [0, -1], [1, -1], [1, -1]]
    assert tp["boundary"].tolist() == [0, 2, 3, 4]


@pytest.mark.parametrize("dim,order,N", [(2, 3, 4), (3, 1, 3), (3, 3, 3), (3, 4, 2)])
def test_kuhn_mesh_generator(dim, order, N):
    nodes, cells = meshgen.kuhn_mesh(N, order, dim, perturb=0.1)
    re = ReferenceElement(dim, order)
    lam = np.concatenate([(1 - 0.5 * (re.nodes + 1).sum(1))[:, None], 0.5 * (re.nodes + 1)], 1)
    X = nodes[cells]
    assert np.abs(np.einsum("nk,ckd->cnd", lam, X[:, :dim + 1]) - X).max() < 1e-14     # straight-sided images of the reference nodes
    assert cells.shape[0] == N ** dim * (6 if dim == 3 else 2)
    vol = np.abs(np.linalg.det(X[:, 1:dim + 1] - X[:, :1])) / (6 if dim == 3 else 2)
    assert abs(vol.sum() - 1.0) < 1e-12
    tp = capi.host_compute_faces(dim, order, cells)
    nb = tp["boundary"].size
    assert nb == (12 * N * N if dim == 3 else 4 * N)
    # Euler: every interior face has two cells
    assert (tp["face2cell"][:, 1] >= 0).sum() * 2 + nb == cells.shape[0] * (dim + 1)


def test_gmres_stock_systems_oracle():
    """tests/unittests/resolution/TestLinAlgebraInterfaces.cpp:70-170 on the oracle's GMRES (non-symmetric systems => GMRES mandatory)."""
    import ctypes as C
    for n in (1, 5, 100):
        for kind in ("identity", "triangular", "hinge"):
            M = {"identity": np.eye(n), "triangular": np.tril(np.ones((n, n))), "hinge": 2 * np.eye(n) - np.eye(n, k=-1)}[kind]
            xs = np.random.default_rng(n).random(n)
            b = M @ xs
            rows, cols = np.nonzero(M)
            rowptr = np.r_[0, np.cumsum(np.bincount(rows, minlength=n))].astype(np.int64)
            colidx = cols.astype(np.int32); vals = M[rows, cols].astype(np.float64)
            x = np.zeros(n); res = C.c_double(0)
            O.lib().orc_gmres(n, rowptr.ctypes.data_as(O._lp), O._i(colidx), O._d(vals), O._d(b), O._d(x), 30, 1, 1, 1e-16, 1000, C.byref(res))
            assert np.abs(x - xs).max() < 1e-10, (n, kind)


@pytest.mark.parametrize("dim,order", [(2, 2), (2, 3), (3, 2)])
@pytest.mark.parametrize("bc", ["dirichlet", "integrated"])
def test_boundary_model_rows_of_the_global_system(dim, order, bc):
    """tests/unittests/model/TestDirichletModel.cpp (local matrix = identity, rhs = the Dirichlet values, assembly type Set) and
    TestIntegratedDirichletModel.cpp:46-54 (local matrix = face mass, rhs = mass x values), seen where they end up: the rows of the
    boundary faces in the assembled trace system (HDGSolver.cpp:361-529: Set semantics, every other entry of those rows is zero)."""
    import scipy.sparse as sp
    from oracle.refel import ReferenceElement
    case = H.make_case(dim, order, N=2, perturb=0.0, bc=bc, model="diffsrc", seed=2)
    o = H.run_oracle(case, solve=False)
    topo, nodes = case["topo"], case["nodes"]
    faces, bnd = topo["faces"], topo["boundary"]
    t = faces.shape[1]
    A = sp.csr_matrix((o.vals, o.colidx, o.rowptr))
    g = case["fields"]["Dirichlet"][:, :, 0]
    fe = ReferenceElement(dim - 1, order)
    Mref = np.einsum("p,pi,pj->ij", fe.ipWeights, fe.ipShape, fe.ipShape)
    for F in bnd:
        rows = A[F * t:(F + 1) * t].toarray()
        diag = rows[:, F * t:(F + 1) * t]
        off = rows.copy()
        off[:, F * t:(F + 1) * t] = 0.0
        assert np.abs(off).max() == 0.0
        if bc == "dirichlet":
            assert np.array_equal(diag, np.eye(t))
            assert np.array_equal(o.rhs[F * t:(F + 1) * t], g[F])
        else:
            x = nodes[faces[F][:dim]]                                       # vertices of the straight-sided face
            meas = np.linalg.norm(x[1] - x[0]) if dim == 2 else 0.5 * np.linalg.norm(np.cross(x[1] - x[0], x[2] - x[0]))
            M = Mref * meas / 2.0                                           # reference measure of the face element: 2 (segment and triangle)
            assert np.abs(diag - M).max() < 1e-13 * np.abs(M).max()
            assert np.abs(o.rhs[F * t:(F + 1) * t] - M @ g[F]).max() < 1e-13 * max(1.0, np.abs(g[F]).max()) * np.abs(M).max() * t


def test_explicit_solver_types_of_the_oracle():
    """HDGSolverOpts.type = WEXPLICIT / SEXPLICIT (HDGSolver.cpp:346-354,626-667,709-729).  The reference holds no known answers for them, so the restatement is
    pinned through what the algebra implies: (i) with the converged implicit Solution / Flux as data, the explicit trace problem returns the implicit trace (the
    l rows of the local systems summed over the elements ARE the global equations, and S_ll couples one face only); (ii) the global solve of WEXPLICIT and the
    per-face solves of SEXPLICIT agree."""
    from tests import helpers as H
    case = H.make_case(2, 3, N=3, perturb=0.1, model="diffsrc", tau_double=True, seed=9)
    o = H.run_oracle(case)
    case["solCur"], case["fluxCur"] = o.sol.copy(), o.flux.copy()
    w = H.run_oracle(case, solverType=1)
    s = H.run_oracle(case, solverType=2)
    assert H.rel_err(w.trace, o.trace) < 1e-10 and H.rel_err(s.trace, o.trace) < 1e-10
    assert H.rel_err(s.trace, w.trace) < 1e-10
    assert H.rel_err(w.sol, o.sol) < 1e-9
    # S = S_ll: no coupling between different faces of an element
    l, t = w.l, w.t
    S = w.S.reshape(w.nCells, l, l)
    for f in range(l // t):
        for g in range(l // t):
            if f != g:
                assert np.abs(S[:, g * t:(g + 1) * t, f * t:(f + 1) * t]).max() == 0.0
