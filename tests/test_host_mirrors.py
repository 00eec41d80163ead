"""Host logic of the Python mirrors that needs no GPU, against the reference's own unit tests."""
import numpy as np
import pytest

from hyperfox_b200 import hfox
from tests.conftest import load_mesh


@pytest.mark.parametrize("start", [1.0, 2.0, -1.0, 0.25, 42.0, 1e-8])
def test_non_linear_wrapper_quadratic_fixed_point(start):
    """tests/unittests/solver/TestNonLinearWrapper.cpp:11-46: Newton on x^2 = 0 through setLinearizedSolver on Node fields of
    lightTri; the loop must end with residual < 1e-6 and |x| < 1e-4 from every starting value."""
    nodes, cells = load_mesh("lightTri")
    m = hfox.Mesh(2, 1, "simplex")
    m.setMesh(nodes, cells)
    sol, inter = hfox.Field(m, hfox.Node, 1, 1), hfox.Field(m, hfox.Node, 1, 1)

    def linearized(solver):
        inter_v = inter.values
        sol.values[:] = inter_v - inter_v ** 2 / (2.0 * inter_v)

    wrap = hfox.NonLinearWrapper()
    wrap.setVerbosity(0)
    wrap.setSolutionFields(sol, inter)
    wrap.setSolver(object())                      # never touched: the linearized solver replaces assemble + solve
    wrap.setLinearizedSolver(linearized)
    inter.values[:] = start
    sol.values[:] = 0.0
    wrap.solve()
    assert wrap.getResidual() < 1e-6
    assert np.abs(inter.values).max() < 1e-4


def test_non_linear_wrapper_contract_and_dampening():
    """NonLinearWrapper.cpp:41-47 (throws before the solver / fields are set) and :60-66 (dampened update of both fields)."""
    wrap = hfox.NonLinearWrapper()
    with pytest.raises(hfox.ErrorHandle, match="the Solver must be set"):
        wrap.solve()
    wrap.setSolver(object())
    with pytest.raises(hfox.ErrorHandle, match="current and previous Solutions"):
        wrap.solve()
    nodes, cells = load_mesh("lightTri")
    m = hfox.Mesh(2, 1, "simplex")
    m.setMesh(nodes, cells)
    cur, prev = hfox.Field(m, hfox.Node, 1, 1), hfox.Field(m, hfox.Node, 1, 1)
    prev.values[:] = 1.0
    calls = []

    def lin(solver):
        calls.append(prev.values.copy())
        cur.values[:] = 0.5 * prev.values       # contraction towards 0

    wrap.setSolutionFields(cur, prev)
    wrap.setLinearizedSolver(lin)
    wrap.setDampening(0.5)
    wrap.setMaxIterations(3)
    wrap.setResidualTolerance(1e-30)
    wrap.solve()
    # each iteration: cur = 0.5 prev, then both <- 0.5 cur + 0.5 prev = 0.75 prev
    assert len(calls) == 3 and np.allclose(calls[1], 0.75) and np.allclose(calls[2], 0.75 ** 2)
    assert np.allclose(prev.values, 0.75 ** 3) and np.allclose(cur.values, prev.values)
    assert abs(wrap.getResidual() - 0.5) < 1e-14
    wrap.setResidualComputer(lambda a, b: 0.0)
    calls.clear()
    wrap.solve()
    assert len(calls) == 1


class _F:
    def __init__(self, v):
        self.values = np.array(v, dtype=float)


@pytest.mark.parametrize("rk,nst", [("BEuler", 1), ("FEuler", 1), ("CrankNicolson", 2), ("RK4", 4)])
def test_runge_kutta_stage_bookkeeping(rk, nst):
    """tests/unittests/operator/TestRungeKutta.cpp:32-43 (Butcher table characteristics) and :112-123, :190-201, :271-300 (computeStage
    stores (Solution - OldSolution)/dt, advances the stage counter, computeSolution = OldSolution + dt sum b_k RKStage_k and resets)."""
    re = hfox.ReferenceElement(2, 2, "simplex")
    ts = hfox.RungeKutta(re, getattr(hfox, rk))
    assert ts.getNumStages() == nst and ts.getStage() == 0
    dt, n = 1e-1, re.getNumNodes()
    ts.setTimeStep(dt)
    fm = {"Solution": _F(np.full(n, 1.0)), "OldSolution": _F(np.full(n, 2.0))}
    for k in range(nst):
        fm["RKStage_%d" % k] = _F(np.full(n, 3.0 + k))
    with pytest.raises(hfox.ErrorHandle, match="all stages must be computed"):
        ts.computeSolution(fm)
    stages = []
    for k in range(nst):
        before = fm["Solution"].values.copy()
        ts.computeStage(fm)
        assert ts.getStage() == k + 1
        stages.append((before - 2.0) / dt)
        assert np.array_equal(fm["RKStage_%d" % k].values, stages[-1])          # TestRungeKutta.cpp:115-118: exactly (1 - 2)/dt at stage 0
        row = ts.bTable[k, 1:]
        assert np.allclose(fm["Solution"].values, 2.0 + dt * sum(row[j] * stages[j] for j in range(k + 1)), rtol=0, atol=1e-14)
        fm["Solution"].values[:] = 1.0 + 0.1 * (k + 1)                          # the next stage solve would overwrite Solution
    with pytest.raises(hfox.ErrorHandle, match="cannot compute more stages"):
        ts.computeStage(fm)
    ts.computeSolution(fm)
    assert ts.getStage() == 0
    bs = ts.bTable[nst, 1:]
    assert np.abs(fm["Solution"].values - (2.0 + dt * sum(bs[k] * stages[k] for k in range(nst)))).max() < 1e-12


def test_runge_kutta_butcher_table_checks():
    """TestRungeKutta.cpp:32-43, same calls: a 2x2 table of ones and a 3x2 table are refused, the 3x3 identity and RK4 are accepted."""
    re = hfox.ReferenceElement(2, 1, "simplex")
    ts = hfox.RungeKutta(re, hfox.CrankNicolson)
    assert ts.getNumStages() == 2 and ts.getStage() == 0
    with pytest.raises(hfox.ErrorHandle, match="upper triangular"):
        ts.setButcherTable(np.ones((2, 2)))
    with pytest.raises(hfox.ErrorHandle, match="square"):
        ts.setButcherTable(np.eye(3)[:, :2])
    ts.setButcherTable(np.eye(3))
    ts.setButcherTable(hfox.RK4)
    assert ts.getNumStages() == 4
    for t in range(14):                                     # every table of the database loads (RungeKutta.cpp:215-291)
        ts.setButcherTable(t)
        assert ts.bTable.shape[0] == ts.bTable.shape[1] and np.all(ts.bTable[-1, 0] == 0) and np.all(np.triu(ts.bTable[:-1, 1:], 1) == 0)
        assert abs(ts.bTable[-1, 1:].sum() - 1.0) < 1e-14  # consistency: sum b = 1


def test_host_builders_edge_cases():
    """Empty meshes, a single cell (every face on the boundary) and a non-manifold input through the host C ABI (no GPU)."""
    from hyperfox_b200 import capi, meshio
    tp = capi.host_compute_faces(3, 2, np.zeros((0, 10), dtype=np.int32))
    assert tp["faces"].shape == (0, 6) and tp["cell2face"].shape == (0, 4) and tp["boundary"].size == 0
    n, c = meshio.high_order_from_linear(3, 2, np.zeros((0, 3)), np.zeros((0, 4), dtype=np.int32))
    assert n.shape == (0, 3) and c.shape == (0, 10)
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1.0]])
    n, h = meshio.high_order_from_linear(3, 4, v, np.array([[0, 1, 2, 3]], dtype=np.int32))
    assert n.shape == (35, 3) and np.array_equal(np.sort(h.ravel()), np.arange(35))
    tp = capi.host_compute_faces(3, 4, h)
    assert tp["faces"].shape == (4, 15) and tp["boundary"].tolist() == [0, 1, 2, 3] and np.array_equal(tp["face2cell"], [[0, -1]] * 4)
    with pytest.raises(capi.ErrorHandle, match="shared by more than two cells"):
        capi.host_compute_faces(3, 1, np.array([[0, 1, 2, 3], [0, 1, 2, 4], [0, 1, 2, 5]], dtype=np.int32))
    m = hfox.Mesh(3, 2, "simplex")
    with pytest.raises(hfox.ErrorHandle, match="connectivity does not match"):
        m.setMesh(v, np.array([[0, 1, 2, 3]], dtype=np.int32))          # 4 nodes per cell for an order-2 reference element (10)
    with pytest.raises(hfox.ErrorHandle, match="not yet supported"):
        hfox.ReferenceElement(3, 2, "prism")


@pytest.mark.parametrize("dim,order,geom,msg", [(3, 3, 1, "interpolation order 3 is not yet supported for dimension 3"),
                                                (3, 6, 0, "interpolation order 6 is not yet supported"),
                                                (2, 6, 1, "interpolation order 6 is not yet supported"),
                                                (4, 1, 0, "spatial dimension 4 is not yet supported")])
def test_reference_element_limits(dim, order, geom, msg):
    """ReferenceElement.cpp:614-634 (maximum orders: 5, hexes 2) and the constructor's checks, TestReferenceElement.cpp:24-44: same
    refusals, same "Class : function : message" text, from the product's host builder."""
    from hyperfox_b200 import capi
    with pytest.raises(capi.ErrorHandle, match=msg):
        capi.host_refel_tables(dim, order, geom)
