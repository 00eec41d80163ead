"""HDGTransport (src/model/HDGTransport.cpp) and the WEXPLICIT / SEXPLICIT solver types (src/solver/HDGSolverOpts.h, HDGSolver.cpp:346-354,605-667,709-729)
on the device against the oracle: the trace problem explicit in the current Solution / Flux fields, S = S_ll (block diagonal per face), global solve (WEXPLICIT)
or one dense solve per face (SEXPLICIT)."""
import numpy as np
import pytest

from hyperfox_b200 import hfox
from tests import helpers as H
from tests.test_gpu_parity import compare, TOL_ENTRIES, TOL_RECOVERY, TOL_SOLUTION

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dim,order", [(2, 2), (2, 4), (3, 1), (3, 3)])
def test_transport_model_with_implicit_euler(dim, order):
    """Base + Convection + Euler mass (the reference's Transport regression steps this model in time), double-valued tau"""
    o, s, fm = compare(H.make_case(dim, order, N=3, perturb=0.1, model="transport_euler", tau_double=True, seed=5))


def _explicit_case(dim, order, model, seed, **kw):
    case = H.make_case(dim, order, N=3, perturb=0.1, model=model, seed=seed, **kw)
    rng = np.random.default_rng(seed + 1)
    nC, nN = case["cells"].shape[0], case["ore"].nNodes
    if "solOld" not in case:
        case["solCur"] = rng.standard_normal((nC, nN))
    case["fluxCur"] = rng.standard_normal((nC, nN * dim))
    return case


def _compare_explicit(case, solverType):
    o = H.run_oracle(case, solverType=solverType)
    s, fm, m = H.run_device(case, solverType=solverType)
    assert s.lastAssembleKernel() == "general"
    loc = s.getLocal()
    for name, ref in (("U", o.U), ("Q", o.Q), ("S", o.S), ("U0", o.U0), ("Q0", o.Q0), ("S0", o.S0)):
        e = H.rel_err(loc[name], ref)
        assert e < (TOL_ENTRIES if name in ("S", "S0") else TOL_RECOVERY), (name, e)
    rowptr, col, vals, rhs = s.getCSR()
    if solverType == 1:
        assert np.array_equal(rowptr, o.rowptr) and np.array_equal(col, o.colidx)
        assert H.rel_err(vals, o.vals) < TOL_ENTRIES
    assert H.rel_err(rhs, o.rhs) < TOL_ENTRIES
    assert H.rel_err(fm["Trace"].values, o.trace) < TOL_SOLUTION
    assert H.rel_err(fm["Solution"].values, o.sol.ravel()) < TOL_SOLUTION
    assert H.rel_err(fm["Flux"].values, o.flux.ravel()) < TOL_SOLUTION
    return o, s, fm


@pytest.mark.parametrize("solverType", [hfox.WEXPLICIT, hfox.SEXPLICIT])
@pytest.mark.parametrize("dim,order,model,kw", [(2, 3, "diffsrc", dict(tau_double=True)), (3, 2, "transport_euler", dict(tau_double=True)),
                                                 (3, 3, "laplace", dict(bc="integrated")), (2, 2, "euler", dict(diff="scalar"))])
def test_explicit_solver_types_match_oracle(solverType, dim, order, model, kw):
    _compare_explicit(_explicit_case(dim, order, model, 21, **kw), solverType)


def test_weak_and_strong_explicit_agree():
    """the WEXPLICIT system is block diagonal per face: GMRES on it and the per-face solves of SEXPLICIT give the same trace"""
    case = _explicit_case(3, 2, "diffsrc", 33, tau_double=True)
    s1, fm1, _ = H.run_device(case, solverType=hfox.WEXPLICIT)
    t1 = fm1["Trace"].values.copy()
    s2, fm2, _ = H.run_device(case, solverType=hfox.SEXPLICIT)
    assert s2.stats.iterations == 0
    assert H.rel_err(fm2["Trace"].values, t1) < TOL_SOLUTION


def test_explicit_step_is_a_fixed_point_of_the_implicit_solution():
    """with the converged implicit Solution / Flux as the explicit data, the l rows of the local systems, summed over the elements, give back the implicit trace"""
    case = H.make_case(3, 2, N=3, perturb=0.1, model="diffsrc", tau_double=True, seed=9)
    s0, fm0, _ = H.run_device(case)
    case["solCur"] = fm0["Solution"].values.reshape(case["cells"].shape[0], -1).copy()
    case["fluxCur"] = fm0["Flux"].values.reshape(case["cells"].shape[0], -1).copy()
    trace0 = fm0["Trace"].values.copy()
    interior = np.ones(case["topo"]["faces"].shape[0], dtype=bool); interior[case["topo"]["boundary"]] = False
    s1, fm1, _ = H.run_device(case, solverType=hfox.SEXPLICIT)
    t = trace0.size // interior.size
    assert H.rel_err(fm1["Trace"].values.reshape(-1, t), trace0.reshape(-1, t)) < 1e-9


def test_sexplicit_needs_no_linear_system():
    """HDGSolver.cpp:12-14: the linear system is optional for the SEXPLICIT type"""
    case = _explicit_case(2, 2, "laplace", 3)
    m = hfox.Mesh(2, 2); m.setMesh(case["nodes"], case["cells"])
    re = m.getReferenceElement(); nN, nNf = re.getNumNodes(), re.getFaceElement().getNumNodes()
    fm = {"Solution": hfox.Field(m, hfox.Cell, nN, 1), "Flux": hfox.Field(m, hfox.Cell, nN, 2), "Trace": hfox.Field(m, hfox.Face, nNf, 1),
          "Tau": hfox.Field(m, hfox.Face, nNf, 1), "Dirichlet": hfox.Field(m, hfox.Face, nNf, 1)}
    fm["Tau"].values[:] = 1.0
    fm["Solution"].values[:] = case["solCur"].ravel(); fm["Flux"].values[:] = case["fluxCur"].ravel()
    s = hfox.HDGSolver(); s.setVerbosity(False); s.setOptions(hfox.HDGSolverOpts(type=hfox.SEXPLICIT, verbosity=False))
    s.setMesh(m); s.setFieldMap(fm); s.setModel(hfox.HDGLaplaceModel(re)); s.setBoundaryModel(hfox.DirichletModel(re.getFaceElement()))
    s.initialize(); s.allocate(); s.assemble(); s.solve()
    assert np.all(np.isfinite(fm["Trace"].values))


def _morlet(t, x, v):
    """tests/regression/HDG/TestHDGTransport.cpp:22-30 with TestUtils::morelet (tests/TestUtils.h.in:272-275): centre 0.3, deviation 1/16, frequency 8 pi"""
    r = np.linalg.norm(x - t * v - 0.3, axis=-1)
    return 0.5 * np.exp(-0.5 * (r * 16.0) ** 2) * np.cos(8.0 * np.pi * r)


def _morlet_grad(t, x, v):
    p = x - t * v - 0.3
    r = np.maximum(np.linalg.norm(p, axis=-1), 1e-300)
    return 0.5 * (p / r[..., None]) * (np.exp(-0.5 * (r * 16.0) ** 2) * ((r * 16.0) * np.cos(8.0 * np.pi * r) + 8.0 * np.pi * np.sin(8.0 * np.pi * r)))[..., None]


@pytest.mark.parametrize("solverType", [hfox.IMPLICIT, hfox.WEXPLICIT, hfox.SEXPLICIT])
def test_transport_regression_time_loop(solverType):
    """tests/regression/HDG/TestHDGTransport.cpp restated: HDGTransport + RungeKutta(BEuler, {Flux, Trace}) + DirichletModel on regression_dim-2_h-2e-1_ord-3,
    v = (1, 1)/sqrt 2, upwind tau (|v.n| on the outflow side of a face, 0 on the other), Morlet wavelet carried by v, dt = 1e-2, for each of the three solver types.
    Device loop against the oracle running the same loop (8 steps), and the reference's own ceiling on the error (l2Err < 1)."""
    from oracle import lib as O
    from oracle.mesh import compute_faces
    from oracle.refel import ReferenceElement as OracleRefEl
    from tests.conftest import load_mesh
    dim, order, dt, nSteps = 2, 3, 1e-2, 8
    nodes, cells = load_mesh("regression_dim-2_h-2e-1_ord-3")
    m = hfox.Mesh(dim, order, "simplex"); m.setMesh(nodes, cells)
    re = m.getReferenceElement()
    nN, nNf, nF, nC = re.getNumNodes(), re.getFaceElement().getNumNodes(), m.getNumberFaces(), m.getNumberCells()
    ts = hfox.RungeKutta(re, hfox.BEuler, ["Flux", "Trace"]); ts.setTimeStep(dt)
    assert ts.getNumStages() == 1
    v = np.full(dim, 1.0 / np.sqrt(2.0))
    # upwind tau: outward normal of the face seen from its first cell (straight faces)
    f2c, faces = m.face2CellMap, m.faces
    fx = nodes[faces[:, :2]]
    tv = fx[:, 1] - fx[:, 0]
    nrm = np.stack([tv[:, 1], -tv[:, 0]], axis=1); nrm /= np.linalg.norm(nrm, axis=1)[:, None]
    cc = nodes[cells[f2c[:, 0], :3]].mean(axis=1)
    flip = np.einsum("fd,fd->f", fx.mean(axis=1) - cc, nrm) < 0
    nrm[flip] *= -1
    proj = nrm @ v
    tau = np.zeros((nF, nNf, 2))
    tau[proj > 0, :, 0] = np.abs(proj[proj > 0])[:, None]
    tau[proj <= 0, :, 1] = np.abs(proj[proj <= 0])[:, None]
    fm = {"Solution": hfox.Field(m, hfox.Cell, nN, 1), "Flux": hfox.Field(m, hfox.Cell, nN, dim), "Trace": hfox.Field(m, hfox.Face, nNf, 1),
          "Tau": hfox.Field(m, hfox.Face, nNf, 2), "Dirichlet": hfox.Field(m, hfox.Face, nNf, 1), "Velocity": hfox.Field(m, hfox.Node, 1, dim),
          "OldSolution": hfox.Field(m, hfox.Cell, nN, 1), "OldFlux": hfox.Field(m, hfox.Cell, nN, dim), "OldTrace": hfox.Field(m, hfox.Face, nNf, 1),
          "RKStage_0": hfox.Field(m, hfox.Cell, nN, 1), "RKStage_Flux_0": hfox.Field(m, hfox.Cell, nN, dim), "RKStage_Trace_0": hfox.Field(m, hfox.Face, nNf, 1)}
    fm["Tau"].setDoubleValued(True); fm["Tau"].values[:] = tau.ravel()
    fm["Velocity"].values[:] = np.tile(v, nodes.shape[0])
    fm["Solution"].values[:] = _morlet(0.0, nodes, v)[cells].ravel()
    fm["Flux"].values[:] = _morlet_grad(0.0, nodes, v)[cells].ravel()
    fm["Trace"].values[:] = _morlet(0.0, nodes, v)[faces].ravel()
    mod = hfox.HDGTransport(re); mod.setTimeScheme(ts)
    s = hfox.HDGSolver()
    s.setVerbosity(False); s.setOptions(hfox.HDGSolverOpts(type=solverType, verbosity=False))
    s.setMesh(m); s.setFieldMap(fm); s.setLinSystem(hfox.CudaLinAlgebraInterface(hfox.PetscOpts(rtol=1e-13, maxits=20000)))
    s.setModel(mod); s.setBoundaryModel(hfox.DirichletModel(re.getFaceElement()))
    s.initialize(); s.allocate()
    ore = OracleRefEl(dim, order); topo = compute_faces(cells, ore)
    of = {"Tau": tau, "Velocity": np.tile(v, (nodes.shape[0], 1)), "Dirichlet": np.zeros((nF, nNf, 1))}
    osol = fm["Solution"].values.reshape(nC, nN).copy(); oflux = fm["Flux"].values.reshape(nC, nN * dim).copy(); otr = fm["Trace"].values.reshape(nF, nNf).copy()
    b = m.boundaryFaces
    row = ts.bTable[0, 1:]
    t = 0.0
    for step in range(nSteps):
        t += dt
        dirv = np.zeros((nF, nNf)); dirv[b] = _morlet(t, nodes, v)[faces[b]]
        fm["Dirichlet"].values[:] = dirv.ravel()
        for a in ("Solution", "Flux", "Trace"):
            fm["Old" + a].values[:] = fm[a].values
        s.assemble(); s.solve(); ts.computeStage(fm)
        ts.computeSolution(fm)
        # oracle: the same loop (RungeKutta.cpp:90-213 with one stage)
        of["Dirichlet"] = dirv.reshape(nF, nNf, 1)
        of.update(solOld=osol.copy(), fluxOld=oflux.copy(), traceOld=otr.reshape(nF, nNf, 1).copy(), Solution=osol.copy(), Flux=oflux.copy())
        o = O.HDGOracle(O.RefElC(ore), dict(nodes=nodes, cells=cells, **topo), O.make_model(1, O.OP_CONVECTION, 0, O.TS_RK, dt, 0, row), of)
        o.solverType = solverType
        o.assemble()
        if solverType == hfox.SEXPLICIT:
            o.solve_faces()
        else:
            o.solve(rtol=1e-13, maxits=20000)
        old = dict(Solution=osol, Flux=oflux, Trace=otr)
        new = dict(Solution=o.sol, Flux=o.flux, Trace=o.trace.reshape(nF, nNf))
        st = {a: (new[a] - old[a]) / dt for a in old}
        bs = ts.bTable[1, 1:]
        osol, oflux, otr = (old[a] + dt * bs[0] * st[a] for a in ("Solution", "Flux", "Trace"))
    assert H.rel_err(fm["Solution"].values, osol.ravel()) < 1e-9
    assert H.rel_err(fm["Flux"].values, oflux.ravel()) < 1e-8
    assert H.rel_err(fm["Trace"].values, otr.ravel()) < 1e-9
    ana = _morlet(t, nodes, v)[cells]
    sol = fm["Solution"].values.reshape(nC, nN)
    assert np.sqrt(((sol - ana) ** 2).sum() / (ana ** 2).sum()) < 1.0     # TestHDGTransport.cpp: CHECK(it->l2Err < 1)
