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
