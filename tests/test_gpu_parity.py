"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every test drives the CUDA path through the reference-shaped API
(hyperfox_b200.hfox -> C ABI of libhfx.so) and compares with the oracle on the same inputs.

Bars (BASELINE.json north_star): CSR structure and scatter indices bit-exact; assembled entries within 1e-12 relative (to the
largest magnitude of the compared block -- individual entries may be exact zeros); solution fields within 1e-10 relative."""
import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu

TOL_ENTRIES = 1e-12      # assembled entries: per-element S, S0 and the global CSR values / RHS
TOL_RECOVERY = 2e-11     # local recovery operators U, Q, U0, Q0 = K^-1(...): conditioning-limited (the oracle's own QR and LU variants
                         # differ by up to ~5e-12 at order 5); they only feed Solution/Flux, whose bar is 1e-10
TOL_SOLUTION = 1e-10


def compare(case, solve=True):
    o = H.run_oracle(case, solve=solve)
    s, fm, m = H.run_device(case, solve=solve)
    # topology + scatter indices: bit exact
    assert np.array_equal(m.faces, case["topo"]["faces"])
    assert np.array_equal(m.cell2FaceMap, case["topo"]["cell2face"])
    assert np.array_equal(m.face2CellMap, case["topo"]["face2cell"])
    assert np.array_equal(s.getElemDofs(), o.elem_dofs())
    rowptr, col, vals, rhs = s.getCSR()
    assert np.array_equal(rowptr, o.rowptr)
    assert np.array_equal(col, o.colidx)
    # per-element condensed blocks
    loc = s.getLocal()
    for name, ref in (("U", o.U), ("Q", o.Q), ("S", o.S), ("U0", o.U0), ("Q0", o.Q0), ("S0", o.S0)):
        e = H.rel_err(loc[name], ref)
        assert e < (TOL_ENTRIES if name in ("S", "S0") else TOL_RECOVERY), (name, e)
    assert H.rel_err(vals, o.vals) < TOL_ENTRIES
    assert H.rel_err(rhs, o.rhs) < TOL_ENTRIES
    if solve:
        assert s.stats.converged == 1
        assert H.rel_err(fm["Trace"].values, o.trace) < TOL_SOLUTION
        assert H.rel_err(fm["Solution"].values, o.sol.ravel()) < TOL_SOLUTION
        assert H.rel_err(fm["Flux"].values, o.flux.ravel()) < TOL_SOLUTION
    return o, s, fm


@pytest.mark.parametrize("dim,order", [(2, 1), (2, 2), (2, 3), (2, 4), (2, 5), (3, 1), (3, 2), (3, 3)])
def test_laplace_kuhn_perturbed(dim, order):
    compare(H.make_case(dim, order, N=3 if dim == 3 else 4, perturb=0.15))


@pytest.mark.parametrize("dim,order,model", [(2, 2, "laplace"), (2, 4, "cdrs"), (3, 2, "diffsrc"), (3, 3, "laplace"), (3, 3, "cdrs")])
def test_curved_elements(dim, order, model):
    """Non-affine geometry: every non-vertex node displaced, so det J, J^-1 and the normals vary over the element."""
    compare(H.make_case(dim, order, N=3, model=model, diff="scalar" if model != "laplace" else "none", tau_double=model != "laplace",
                        curved=0.04, seed=11))


@pytest.mark.parametrize("name,dim,order", [("regression_dim-2_h-1e-1_ord-2", 2, 2), ("regression_dim-3_h-2e-1_ord-3", 3, 3),
                                             ("regression_dim-3_h-3e-1_ord-2", 3, 2)])
def test_laplace_reference_meshes(name, dim, order):
    """tests/regression/HDG/TestHDGLaplace.cpp:110-139 -- u = sin(x) e^y, l2 error ceiling 1e-2 (here: nodal relative l2)."""
    case = H.make_case(dim, order, mesh=name)
    o, s, fm = compare(case)
    sol = fm["Solution"].values.reshape(case["cells"].shape)
    ana = case["ana"][case["cells"]]
    assert np.sqrt(((sol - ana) ** 2).sum() / (ana ** 2).sum()) < 1e-2


def test_hdgsolver_constant_solution():
    """tests/unittests/solver/TestHDGSolver.cpp:16-100: lightTri2, tau = 1, Dirichlet = 3 => u = 3, q = 0, lambda = 3 (1e-12)."""
    case = H.make_case(2, 2, mesh="lightTri2")
    case["fields"]["Dirichlet"][:] = 3.0
    s, fm, m = H.run_device(case, rtol=1e-12)
    assert np.abs(fm["Solution"].values - 3.0).max() < 1e-12
    assert np.abs(fm["Flux"].values).max() < 1e-12
    assert np.abs(fm["Trace"].values - 3.0).max() < 1e-12


@pytest.mark.parametrize("dim,order,diff", [(2, 3, "scalar"), (3, 2, "scalar"), (3, 3, "tensor"), (2, 2, "tensor"), (3, 3, "none")])
def test_diffusion_source(dim, order, diff):
    compare(H.make_case(dim, order, N=3, model="diffsrc", diff=diff, tau_double=True, seed=3))


@pytest.mark.parametrize("dim,order,diff", [(2, 2, "scalar"), (3, 3, "scalar"), (3, 2, "tensor"), (3, 3, "none")])
def test_convection_diffusion_reaction_source(dim, order, diff):
    compare(H.make_case(dim, order, N=3, model="cdrs", diff=diff, tau_double=True, seed=5))


@pytest.mark.parametrize("dim,order", [(2, 2), (3, 3)])
def test_euler_time_scheme(dim, order):
    compare(H.make_case(dim, order, N=3, model="euler", diff="scalar", seed=7))


@pytest.mark.parametrize("dim,order", [(2, 3), (3, 2)])
def test_integrated_dirichlet(dim, order):
    compare(H.make_case(dim, order, N=3, bc="integrated", seed=9))


@pytest.mark.parametrize("dim,order,model,bc,mixed", [(3, 3, "laplace", "dirichlet", False), (3, 3, "diffsrc", "dirichlet", False), (3, 2, "diffsrc", "integrated", False),
                                                       (2, 3, "diffsrc", "dirichlet", False), (2, 5, "laplace", "integrated", False), (3, 3, "diffsrc", "dirichlet", True),
                                                       (3, 1, "laplace", "dirichlet", False)])
def test_all_reference_path(dim, order, model, bc, mixed, monkeypatch):
    """Straight-sided elements of a Laplace-type model whose tau is constant on each face (a different value per face and per side)
    take every block from reference matrices; `mixed` makes tau vary along some faces, so both paths meet inside one mesh.
    The same inputs with the path disabled (HFX_NO_REFPATH) must give the same condensed blocks."""
    case = H.make_case(dim, order, N=3, perturb=0.12, model=model, bc=bc, tau_double=True, seed=21)
    tau = case["fields"]["Tau"]
    rng = np.random.default_rng(5)
    keep = rng.random(tau.shape[0]) < 0.5 if mixed else np.zeros(tau.shape[0], dtype=bool)
    tau[~keep] = tau[~keep][:, :1, :]          # constant along the face, one value per side
    o, s, fm = compare(case)
    monkeypatch.setenv("HFX_NO_REFPATH", "1")
    s2, fm2, _ = H.run_device(case)
    a, b = s.getLocal(), s2.getLocal()
    for name in ("S", "S0", "U", "Q", "U0", "Q0"):
        assert H.rel_err(a[name], b[name]) < TOL_ENTRIES, name


@pytest.mark.parametrize("dim,order,model", [(3, 3, "diffsrc"), (2, 2, "diffsrc"), (3, 3, "cdrs"), (2, 4, "cdrs"), (3, 4, "cdrs"), (3, 2, "euler")])
def test_constant_scalar_diffusion_field(dim, order, model, monkeypatch):
    """A scalar DiffusionTensor field that is constant over the mesh is D = c I: the kernels scale the D = I blocks instead of
    interpolating the field (and a Laplace-type model stays on the all-reference path).  Same result as the field path."""
    case = H.make_case(dim, order, N=2 if order == 4 and dim == 3 else 3, perturb=0.1, model=model, diff="const", tau_double=model != "euler", seed=31)
    if model == "diffsrc":
        case["fields"]["Tau"][:] = case["fields"]["Tau"][:, :1, :]      # face-constant tau: all-reference path with c != 1
    o, s, fm = compare(case)
    monkeypatch.setenv("HFX_NO_CONST_DIFF", "1")
    s2, fm2, _ = H.run_device(case)
    a, b = s.getLocal(), s2.getLocal()
    for name in ("S", "S0", "U", "Q", "U0", "Q0"):
        assert H.rel_err(a[name], b[name]) < TOL_RECOVERY, name


@pytest.mark.parametrize("dim,order,model", [(3, 3, "laplace"), (2, 4, "cdrs"), (3, 2, "diffsrc")])
def test_face_block_jacobi_preconditioner(dim, order, model):
    """pc = 2: Jacobi on the t x t diagonal face blocks of the trace system.  Same solution (1e-10), fewer GMRES iterations than
    point Jacobi."""
    from hyperfox_b200 import hfox
    case = H.make_case(dim, order, N=3, perturb=0.1, model=model, diff="scalar" if model != "laplace" else "none", seed=41)
    o = H.run_oracle(case)
    s, fm, m = H.run_device(case)
    its_point = s.stats.iterations
    s.linSystem.opts.preconditionnerType = hfox.PCBJACOBI
    s.solve()
    assert s.stats.converged == 1
    assert H.rel_err(fm["Trace"].values, o.trace) < TOL_SOLUTION
    assert H.rel_err(fm["Solution"].values, o.sol.ravel()) < TOL_SOLUTION
    assert H.rel_err(fm["Flux"].values, o.flux.ravel()) < TOL_SOLUTION
    assert s.stats.iterations < its_point, (s.stats.iterations, its_point)


def test_reassembly_is_bit_reproducible():
    """Deterministic scatter: two assemblies of the same inputs give bit-identical CSR values (<= 2 contributors per entry)."""
    case = H.make_case(3, 3, N=3, model="cdrs", diff="scalar", tau_double=True)
    s, fm, m = H.run_device(case, solve=False)
    v1 = s.getCSR(with_cols=False)[2].copy()
    s.assemble()
    v2 = s.getCSR(with_cols=False)[2]
    assert np.array_equal(v1, v2)


def test_call_order_contract():
    """tests/unittests/solver/TestHDGSolver.cpp:37-80: every step throws before its prerequisite."""
    from hyperfox_b200 import hfox
    s = hfox.HDGSolver()
    with pytest.raises(hfox.ErrorHandle):
        s.solve()
    with pytest.raises(hfox.ErrorHandle):
        s.assemble()
    with pytest.raises(hfox.ErrorHandle):
        s.allocate()
    s.initialize()
    with pytest.raises(hfox.ErrorHandle):
        s.allocate()


# ---- LinAlgebraInterface mirror: tests/unittests/resolution/TestLinAlgebraInterfaces.cpp:29-170 ---------------------
def _lai(rtol=1e-16):
    from hyperfox_b200 import hfox
    return hfox, hfox.CudaLinAlgebraInterface(hfox.PetscOpts(rtol=rtol, maxits=1000))


def test_lai_state_machine():
    hfox, a = _lai()
    for fn in (a.configure, lambda: a.allocate(3), lambda: a.addValMatrix(0, 0, 1.0), a.assemble, a.solve):
        with pytest.raises(hfox.ErrorHandle):
            fn()
    a.initialize()
    with pytest.raises(hfox.ErrorHandle):
        a.allocate(3)
    a.configure()
    with pytest.raises(hfox.ErrorHandle):
        a.addValMatrix(0, 0, 1.0)
    a.allocate(3)
    with pytest.raises(hfox.ErrorHandle):
        a.solve()
    with pytest.raises(hfox.ErrorHandle):
        a.clearSystem()


@pytest.mark.parametrize("n", [1, 2, 5, 10, 100])
def test_lai_stock_systems(n):
    hfox, a = _lai()
    rng = np.random.default_rng(n)
    # identity
    a.initialize(); a.configure(); a.allocate(n)
    b = rng.random(n)
    for i in range(n):
        a.addValMatrix(i, i, 1.0)
        a.addValRHS(i, b[i])
    a.assemble()
    assert np.abs(a.solve() - b).max() < 1e-12
    # lower triangular ones: x = (b0, b1-b0, ...)
    a.destroySystem(); a.initialize(); a.configure(); a.allocate(n)
    rows = np.arange(n)
    a.addValsMatrix(rows, rows, np.tril(np.ones((n, n))))      # row-major |is| x |js| (TestPetscInterface.cpp:57-63)
    a.addValsRHS(rows, np.arange(1, n + 1, dtype=float))
    a.assemble()
    assert np.abs(a.solve() - 1.0).max() < 1e-10
    # bidiagonal "hinge": 2 on the diagonal, -1 on the sub-diagonal
    a.clearSystem()
    M = 2 * np.eye(n) - np.eye(n, k=-1)
    a.setValsMatrix(rows, rows, M)
    xs = rng.random(n)
    a.setValsRHS(rows, M @ xs)
    a.assemble()
    assert np.abs(a.solve() - xs).max() < 1e-10


# ---- general kernel (hfx_generic.cuh): 3-D orders 4-5, nDOFsPerNode > 1, HDGUNabU --------------------------------------------
@pytest.mark.parametrize("dim,order,model,diff,bc", [(2, 2, "cdrs", "scalar", "dirichlet"), (3, 2, "diffsrc", "tensor", "dirichlet"),
                                                     (3, 3, "laplace", "none", "dirichlet"), (2, 3, "euler", "scalar", "dirichlet"),
                                                     (3, 2, "laplace", "none", "integrated"), (3, 3, "cdrs", "scalar", "dirichlet")])
def test_general_kernel_matches_oracle_and_fused(dim, order, model, diff, bc, monkeypatch):
    """The general kernel forced on configurations the fused kernel also covers: same parity bars against the oracle."""
    monkeypatch.setenv("HFX_FORCE_GENERIC", "1")
    compare(H.make_case(dim, order, N=3, model=model, diff=diff, bc=bc, tau_double=model != "laplace", seed=13, curved=0.03 if model == "cdrs" else 0.0))


@pytest.mark.parametrize("order,model", [(4, "laplace"), (5, "laplace"), (4, "cdrs")])
def test_3d_orders_4_and_5(order, model):
    """BASELINE.json configs[3]/[4]: order 4 (convection-diffusion) and the order sweep up to the reference's maximum (5) on tets."""
    compare(H.make_case(3, order, N=2, perturb=0.1, model=model, diff="scalar" if model == "cdrs" else "none", tau_double=model == "cdrs", seed=17))


@pytest.mark.parametrize("dim,order,diff,tau_double", [(2, 1, "scalar", False), (2, 2, "scalar", True), (2, 3, "tensor", False), (2, 4, "scalar", True),
                                                       (3, 2, "scalar", False)])
def test_burgers_model_one_newton_linearisation(dim, order, diff, tau_double):
    """HDGBurgersModel (nDOFsPerNode = dim): Base + HDGUNabU + Diffusion with per-component sources, linearised about a random
    previous iterate (BufferSolution, Trace) -- BASELINE.json configs[1] at every order of its sweep."""
    compare(H.make_case(dim, order, N=3, model="burgers", diff=diff, tau_double=tau_double, seed=19))


def _burgers_stat_analytic(x, D=1.0, A=0.5, x0=(1.0, 0.0)):
    """tests/regression/HDG/TestHDGBurgersStat.cpp:46-90: Cole-Hopf solution u = -2 D grad(phi) / phi."""
    ex, em = np.exp(A * (x[:, 0] - x0[0])), np.exp(-A * (x[:, 0] - x0[0]))
    cy, sy = np.cos(A * (x[:, 1] - x0[1])), np.sin(A * (x[:, 1] - x0[1]))
    pot = 0.001 * A * np.exp((1 + x0[0]) * A) * (1 + x[:, 0]) + (ex + em) * cy
    g0 = 0.001 * A * np.exp((1 + x0[0]) * A) + A * (ex - em) * cy
    g1 = -A * (ex + em) * sy
    return np.stack([g0, g1], axis=1) * (-2.0 * D / pot)[:, None]


@pytest.mark.parametrize("order,hname,h", [(1, "2e-1", 0.2), (2, "2e-1", 0.2), (3, "2e-1", 0.2), (4, "2e-1", 0.2),
                                           (1, "1e-1", 0.1), (2, "1e-1", 0.1), (3, "1e-1", 0.1), (4, "1e-1", 0.1)])
def test_burgers_stationary_newton_regression(order, hname, h):
    """BASELINE.json configs[1] = tests/regression/HDG/TestHDGBurgersStat.cpp: HDGBurgersModel + IntegratedDirichletModel on
    regression_dim-2_h-2e-1_ord-p, D = 1, tau = (2D/h) I on both sides, NonLinearWrapper(<= 10 iterations, tol 1e-6) from a zero state.
    The device Newton loop must (i) converge to the Cole-Hopf solution and (ii) follow the oracle's Newton iterates."""
    from hyperfox_b200 import hfox
    from oracle import lib as O
    from oracle.mesh import compute_faces
    from oracle.refel import ReferenceElement as OracleRefEl
    from tests.conftest import load_mesh
    dim, D = 2, 1.0     # both meshes of the reference's sweep (TestHDGBurgersStat.cpp: meshSizes {2e-1, 1e-1} x orders {1..4})
    nodes, cells = load_mesh("regression_dim-2_h-%s_ord-%d" % (hname, order))
    m = hfox.Mesh(dim, order, "simplex")
    m.setMesh(nodes, cells)
    re = m.getReferenceElement()
    nN, nNf, nF, nC = re.getNumNodes(), re.getFaceElement().getNumNodes(), m.getNumberFaces(), m.getNumberCells()
    ana = _burgers_stat_analytic(nodes, D)
    fm = {"Solution": hfox.Field(m, hfox.Cell, nN, dim), "BufferSolution": hfox.Field(m, hfox.Cell, nN, dim), "Flux": hfox.Field(m, hfox.Cell, nN, dim * dim),
          "Trace": hfox.Field(m, hfox.Face, nNf, dim), "Tau": hfox.Field(m, hfox.Face, nNf, 2 * dim * dim), "Dirichlet": hfox.Field(m, hfox.Face, nNf, dim),
          "DiffusionTensor": hfox.Field(m, hfox.Node, 1, 1)}
    fm["Tau"].setDoubleValued(True)
    tau = np.zeros((nF, nNf, 2, dim, dim)); tau[..., 0, 0] = tau[..., 1, 1] = 2.0 * D / h
    fm["Tau"].values[:] = tau.ravel()
    fm["DiffusionTensor"].values[:] = D
    dirv = np.zeros((nF, nNf, dim)); b = m.boundaryFaces
    dirv[b] = ana[m.faces[b]]
    fm["Dirichlet"].values[:] = dirv.ravel()
    mod = hfox.HDGBurgersModel(re)
    s = hfox.HDGSolver()
    s.setMesh(m); s.setFieldMap(fm); s.setLinSystem(hfox.CudaLinAlgebraInterface(hfox.PetscOpts(rtol=1e-14, maxits=10000)))
    s.setModel(mod); s.setBoundaryModel(hfox.IntegratedDirichletModel(re.getFaceElement()))
    s.initialize(); s.allocate()
    w = hfox.NonLinearWrapper()
    w.setMaxIterations(10); w.setResidualTolerance(1e-6); w.setSolutionFields(fm["Solution"], fm["BufferSolution"]); w.setSolver(s)
    w.solve()
    assert w.getResidual() < 1e-6
    sol = fm["Solution"].values.reshape(nC, nN, dim)
    err = np.sqrt(((sol - ana[cells]) ** 2).sum() / (ana[cells] ** 2).sum())
    assert err < [3e-2, 2e-3, 3e-4, 5e-5][order - 1], err      # reference ceiling: l2Err < 1 (TestHDGBurgersStat.cpp:363)
    # the same Newton loop on the oracle
    ore = OracleRefEl(dim, order)
    topo = compute_faces(cells, ore)
    f = {"Tau": tau.reshape(nF, nNf, 2 * dim * dim), "Dirichlet": dirv, "DiffusionTensor": np.full((nodes.shape[0], 1), D),
         "BufferSolution": np.zeros((nC, nN, dim)), "Trace": np.zeros((nF, nNf, dim))}
    o = O.HDGOracle(O.RefElC(ore), dict(nodes=nodes, cells=cells, **topo), O.make_model(dim, O.OP_UNABU | O.OP_DIFFUSION, 1), f, bcKind=O.BC_INTEGRATED_DIRICHLET)
    cur, prev = np.zeros((nC, nN * dim)), np.zeros((nC, nN * dim))
    for _ in range(10):
        f["BufferSolution"] = prev.reshape(nC, nN, dim)
        o.set_fields(f); o.assemble(); o.solve(rtol=1e-14, maxits=10000)
        cur = o.sol.copy(); f["Trace"] = o.trace.reshape(nF, nNf, dim)
        diff, ref = ((cur - prev) ** 2).sum(), (prev ** 2).sum()
        res = np.sqrt(diff / ref) if ref != 0 else np.sqrt(diff)
        if res < 1e-6:
            break
        prev = cur.copy()
    assert H.rel_err(fm["Solution"].values, cur.ravel()) < 1e-8


def _run_dist(nproc, cubes=3, order=3, geom="simplex"):
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1", "--master-port", "29531",
           os.path.join(root, "tests", "dist_solve_check.py"), str(cubes), str(order), geom]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and "DIST_OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]


def test_distributed_solve_single_rank_communicator():
    """The NCCL path with a one-rank communicator (what a 1-GPU box can run): ownership mask, masked right-hand side, all-reduce."""
    _run_dist(1)


@pytest.mark.parametrize("transport", ["p2p", "nccl"])
def test_distributed_solve_two_gpus(transport, monkeypatch):
    """Trace-halo exchange + summed dots on two GPUs -- over NVLink peer memory (direct stores into the neighbour's ghost buffers, one-shot all-reduce;
    the default) and over NCCL (HFX_P2P=0): solution fields within 1e-10 of the single-process oracle, same GMRES iteration count."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    if transport == "nccl":
        monkeypatch.setenv("HFX_P2P", "0")
    _run_dist(2, cubes=4)


def test_distributed_solve_of_hexahedra_two_gpus():
    """The multi-GPU harness on orthotope cells (order-2 hexahedra, perturbed so that they are genuinely trilinear): plan of the hex mesh in host C++, quadrilateral
    face blocks exchanged in their canonical node order, solution within 1e-10 of the single-process oracle."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    _run_dist(2, cubes=4, order=2, geom="orthotope")


def test_distributed_harness_of_hexahedra_single_rank():
    _run_dist(1, cubes=3, order=2, geom="orthotope")


def _diffsrc_analytic(t, x):
    """tests/regression/HDG/TestHDGDiffusionSource.cpp:32-47 (analyticalDiffSrc) and :23-30 (gaussianSrc)."""
    from scipy.special import erf
    a = x - 0.5
    res = (a * erf(a) + np.exp(-a ** 2) / np.sqrt(np.pi)).sum(axis=1)
    return res + np.exp(-x.shape[1] * (np.pi / 2) ** 2 * t) * np.cos(np.pi / 2 * x.sum(axis=1))


@pytest.mark.parametrize("rk,meshname,nSteps", [("BEuler", "regression_dim-2_h-2e-1_ord-2", 6), ("CrankNicolson", "regression_dim-2_h-2e-1_ord-2", 6),
                                                ("QZ2", "regression_dim-2_h-2e-1_ord-2", 6), ("BEuler", "regression_dim-2_h-1e-1_ord-2", 100)])
def test_diffusion_source_runge_kutta_time_loop(rk, meshname, nSteps):
    """BASELINE.json configs[0] = tests/regression/HDG/TestHDGDiffusionSource.cpp: HDGDiffusionSource + RungeKutta(type, {Flux, Trace}) +
    DirichletModel on regression_dim-2_h-2e-1_ord-2, tau = 1/sqrt(dt), gaussian source, dt = 1e-2 -- the reference's time loop
    (OldX <- X; per stage: assemble, solve, computeStage; computeSolution).  Device vs the oracle running the same loop, and vs the
    analytic solution."""
    from hyperfox_b200 import hfox
    from oracle import lib as O
    from oracle.mesh import compute_faces
    from oracle.refel import ReferenceElement as OracleRefEl
    from tests.conftest import load_mesh
    dim, order, dt = 2, 2, 1e-2     # the last case is BASELINE configs[0] at its full length: h = 1e-1, order 2, 100 time steps to t = 1
    nodes, cells = load_mesh(meshname)
    m = hfox.Mesh(dim, order, "simplex"); m.setMesh(nodes, cells)
    re = m.getReferenceElement()
    nN, nNf, nF, nC = re.getNumNodes(), re.getFaceElement().getNumNodes(), m.getNumberFaces(), m.getNumberCells()
    ts = hfox.RungeKutta(re, getattr(hfox, rk), ["Flux", "Trace"]); ts.setTimeStep(dt)
    nSt = ts.getNumStages()
    fm = {"Solution": hfox.Field(m, hfox.Cell, nN, 1), "Flux": hfox.Field(m, hfox.Cell, nN, dim), "Trace": hfox.Field(m, hfox.Face, nNf, 1),
          "Tau": hfox.Field(m, hfox.Face, nNf, 1), "Dirichlet": hfox.Field(m, hfox.Face, nNf, 1), "DiffusionTensor": hfox.Field(m, hfox.Node, 1, 1),
          "OldSolution": hfox.Field(m, hfox.Cell, nN, 1), "OldFlux": hfox.Field(m, hfox.Cell, nN, dim), "OldTrace": hfox.Field(m, hfox.Face, nNf, 1)}
    for k in range(nSt):
        fm["RKStage_%d" % k] = hfox.Field(m, hfox.Cell, nN, 1); fm["RKStage_Flux_%d" % k] = hfox.Field(m, hfox.Cell, nN, dim)
        fm["RKStage_Trace_%d" % k] = hfox.Field(m, hfox.Face, nNf, 1)
    fm["Tau"].values[:] = 1.0 / np.sqrt(dt); fm["DiffusionTensor"].values[:] = 1.0
    fm["Solution"].values[:] = _diffsrc_analytic(0.0, nodes)[cells].ravel()
    fm["Trace"].values[:] = _diffsrc_analytic(0.0, nodes)[m.faces].ravel()
    src = lambda x: -2.0 / np.sqrt(np.pi) * sum(np.exp(-(xi - 0.5) ** 2) for xi in x)
    mod = hfox.HDGDiffusionSource(re); mod.setTimeScheme(ts)
    s = hfox.HDGSolver()
    s.setMesh(m); s.setFieldMap(fm); s.setLinSystem(hfox.CudaLinAlgebraInterface(hfox.PetscOpts(rtol=1e-13, maxits=20000)))
    s.setModel(mod); s.setBoundaryModel(hfox.DirichletModel(re.getFaceElement()))
    s.initialize(); s.allocate()
    mod.setSourceFunction(src)
    # oracle twin
    ore = OracleRefEl(dim, order); topo = compute_faces(cells, ore)
    xip = np.einsum("pi,cid->cpd", ore.ipShape, nodes[cells])
    of = {"Tau": np.full((nF, nNf, 1), 1.0 / np.sqrt(dt)), "DiffusionTensor": np.ones((nodes.shape[0], 1)), "Dirichlet": np.zeros((nF, nNf, 1)),
          "srcIP": np.array([[src(p) for p in el] for el in xip])}
    osol = fm["Solution"].values.reshape(nC, nN).copy(); oflux = np.zeros((nC, nN * dim)); otr = fm["Trace"].values.reshape(nF, nNf).copy()
    b = m.boundaryFaces
    t = 0.0
    for step in range(nSteps):
        t += dt
        dirv = np.zeros((nF, nNf)); dirv[b] = _diffsrc_analytic(t, nodes)[m.faces[b]]
        fm["Dirichlet"].values[:] = dirv.ravel()
        for a in ("Solution", "Flux", "Trace"):
            fm["Old" + a].values[:] = fm[a].values
        for k in range(nSt):
            s.assemble(); s.solve(); ts.computeStage(fm)
        ts.computeSolution(fm)
        # oracle: same loop (RungeKutta.cpp:90-213)
        of["Dirichlet"] = dirv.reshape(nF, nNf, 1)
        old = dict(Solution=osol.copy(), Flux=oflux.copy(), Trace=otr.copy())
        cur = dict(Solution=osol, Flux=oflux, Trace=otr)
        st = {a: [] for a in cur}
        tab = ts.bTable
        for k in range(nSt):
            row = tab[k, 1:]
            of.update(solOld=old["Solution"], fluxOld=old["Flux"], traceOld=old["Trace"].reshape(nF, nNf, 1))
            if k > 0:
                of.update(rkSol=np.array(st["Solution"]), rkFlux=np.array(st["Flux"]), rkTrace=np.array(st["Trace"]).reshape(k, nF, nNf, 1))
            o = O.HDGOracle(O.RefElC(ore), dict(nodes=nodes, cells=cells, **topo), O.make_model(1, O.OP_DIFFUSION | O.OP_SOURCE, 1, O.TS_RK, dt, k, row), of)
            o.assemble(); o.solve(rtol=1e-13, maxits=20000)
            new = dict(Solution=o.sol, Flux=o.flux, Trace=o.trace.reshape(nF, nNf))
            for a in cur:
                st[a].append((new[a] - old[a]) / dt)
                cur[a] = old[a] + dt * sum(row[j] * st[a][j] for j in range(k + 1))
        bs = tab[nSt, 1:]
        osol, oflux, otr = (old[a] + dt * sum(bs[k] * st[a][k] for k in range(nSt)) for a in ("Solution", "Flux", "Trace"))
    assert H.rel_err(fm["Solution"].values, osol.ravel()) < (1e-9 if nSteps <= 6 else 1e-8)
    assert H.rel_err(fm["Flux"].values, oflux.ravel()) < (1e-8 if nSteps <= 6 else 1e-7)
    ana = _diffsrc_analytic(t, nodes)[cells]
    sol = fm["Solution"].values.reshape(nC, nN)
    assert np.sqrt(((sol - ana) ** 2).sum() / (ana ** 2).sum()) < 1e-2     # reference ceiling on the time-integrated l2 error (TestHDGDiffusionSource.cpp)
