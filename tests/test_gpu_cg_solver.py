"""CGSolver on the device (hfx_cg_*, hyperfox_b200/csrc/hfx_cg.cuh) against the CG oracle (oracle/cg.py) and the reference's own tests:
tests/unittests/solver/TestCGSolver.cpp (call-order contract, constant state) and tests/regression/CG/TestCGLaplace.cpp (harmonic solution)."""
import numpy as np
import pytest

from hyperfox_b200 import hfox, meshgen
from oracle import cg
from oracle.mesh import compute_faces
from oracle.refel import ReferenceElement as OracleRefEl
from tests import helpers as H
from tests.conftest import load_mesh

pytestmark = pytest.mark.gpu


def _setup(dim, order, nodes, cells, model="laplace", diff=None, source=None, geom="simplex"):
    m = hfox.Mesh(dim, order, geom); m.setMesh(nodes, cells)
    re = m.getReferenceElement()
    nNf = re.getFaceElement().getNumNodes()
    fm = {"Solution": hfox.Field(m, hfox.Node, 1, 1), "Dirichlet": hfox.Field(m, hfox.Face, nNf, 1)}
    if diff is not None:
        fm["DiffusionTensor"] = hfox.Field(m, hfox.Node, 1, diff.shape[1]); fm["DiffusionTensor"].values[:] = diff.ravel()
    mod = hfox.LaplaceModel(re) if model == "laplace" else hfox.DiffusionSource(re)
    s = hfox.CGSolver()
    s.setVerbosity(False); s.setMesh(m); s.setFieldMap(fm)
    s.setLinSystem(hfox.CudaLinAlgebraInterface(hfox.PetscOpts(rtol=1e-14, maxits=20000)))
    s.setModel(mod); s.setBoundaryModel(hfox.DirichletModel(re.getFaceElement()))
    s.initialize(); s.allocate()
    if source is not None:
        mod.setSourceFunction(source)
    return m, fm, s


def _compare(dim, order, nodes, cells, model="laplace", diff=None, source=None, geom="simplex"):
    m, fm, s = _setup(dim, order, nodes, cells, model, diff, source, geom)
    ore = OracleRefEl(dim, order, geom)
    topo = compute_faces(cells, ore)
    assert np.array_equal(m.faces, topo["faces"])
    ana = np.sin(nodes[:, 0]) * np.exp(nodes[:, 1])
    dirv = np.zeros(topo["faces"].shape); dirv[topo["boundary"]] = ana[topo["faces"][topo["boundary"]]]
    fm["Dirichlet"].values[:] = dirv.ravel()
    s.assemble(); s.solve()
    o = cg.CGOracle(ore, nodes, cells, topo["faces"], topo["boundary"], diff=diff, source=source)
    o.assemble(dirv); o.solve()
    rowptr, col, vals, rhs = s.getCSR()
    assert np.array_equal(rowptr, o.rowptr) and np.array_equal(col, o.colidx)          # CSR structure: bit exact
    assert H.rel_err(vals, o.vals) < 1e-12 and H.rel_err(rhs, o.b) < 1e-12               # assembled entries
    assert s.stats.converged == 1
    assert H.rel_err(fm["Solution"].values, o.sol) < 1e-10                               # solution field
    return m, fm, s, o, ana


@pytest.mark.parametrize("dim,order", [(2, 1), (2, 3), (2, 5), (3, 1), (3, 2), (3, 4)])
def test_cg_laplace_matches_oracle(dim, order):
    nodes, cells = meshgen.kuhn_mesh(3, order, dim, perturb=0.12)
    _compare(dim, order, nodes, cells)


@pytest.mark.parametrize("dim,order,comps", [(2, 2, 1), (3, 2, 1), (2, 3, 4), (3, 2, 9)])
def test_cg_diffusion_source_matches_oracle(dim, order, comps):
    nodes, cells = meshgen.kuhn_mesh(3, order, dim, perturb=0.1)
    rng = np.random.default_rng(5)
    if comps == 1:
        diff = 0.5 + rng.random((nodes.shape[0], 1))
    else:
        A = rng.standard_normal((nodes.shape[0], dim, dim)) * 0.2
        D = np.eye(dim)[None] + A @ A.transpose(0, 2, 1) + 0.1 * A                       # not symmetric on purpose: exercises the column-major layout
        diff = D.transpose(0, 2, 1).reshape(nodes.shape[0], dim * dim)
    src = lambda x: np.exp(-10 * sum((xi - 0.5) ** 2 for xi in x))
    _compare(dim, order, nodes, cells, model="diffsrc", diff=diff, source=src)


def test_cg_curved_and_orthotope_elements():
    nodes, cells = meshgen.kuhn_mesh(3, 3, 2, perturb=0.1)
    isv = np.zeros(nodes.shape[0], dtype=bool); isv[np.unique(cells[:, :3])] = True
    nodes = nodes.copy(); nodes[~isv] += 0.01 * np.random.default_rng(2).uniform(-1, 1, size=(int((~isv).sum()), 2))
    _compare(2, 3, nodes, cells)
    qn, qc = meshgen.box_mesh(3, 2, 3, perturb=0.1)
    _compare(3, 2, qn, qc, geom="orthotope")


@pytest.mark.parametrize("dim,order,name", [(2, 2, "regression_dim-2_h-1e-1_ord-2"), (3, 3, "regression_dim-3_h-2e-1_ord-3")])
def test_cg_laplace_regression(dim, order, name):
    """tests/regression/CG/TestCGLaplace.cpp: harmonic u = sin x e^y, nodal l2 error under the reference's ceiling"""
    nodes, cells = load_mesh(name)
    m, fm, s, o, ana = _compare(dim, order, nodes, cells)
    sol = fm["Solution"].values
    assert np.sqrt(((sol - ana) ** 2).sum() / (ana ** 2).sum()) < 1e-2


def test_cg_solver_reference_unit_test():
    """tests/unittests/solver/TestCGSolver.cpp:37-66 restated: every step throws before its prerequisite; lightTri, Dirichlet = 3 => Solution = 3 (1e-12)"""
    nodes, cells = load_mesh("lightTri")
    m = hfox.Mesh(2, 1, "simplex"); m.setMesh(nodes, cells)
    re = m.getReferenceElement()
    fm = {"Solution": hfox.Field(m, hfox.Node, 1, 1), "Dirichlet": hfox.Field(m, hfox.Face, re.getFaceElement().getNumNodes(), 1)}
    fm["Dirichlet"].values[:] = 3.0
    s = hfox.CGSolver()
    s.setVerbosity(False)

    def all_throw(alloc=True):
        for fn in (s.solve, s.assemble) + ((s.allocate,) if alloc else ()):
            with pytest.raises(hfox.ErrorHandle):
                fn()
    all_throw(); s.setMesh(m)
    all_throw(); s.setFieldMap(fm)
    all_throw(); s.setLinSystem(hfox.CudaLinAlgebraInterface(hfox.PetscOpts(rtol=1e-14)))
    all_throw(); s.setModel(hfox.LaplaceModel(re))
    all_throw(); s.setBoundaryModel(hfox.DirichletModel(re.getFaceElement()))
    all_throw(); s.initialize()
    s.allocate()
    with pytest.raises(hfox.ErrorHandle):
        s.solve()
    s.assemble(); s.solve()
    assert np.abs(fm["Solution"].values - 3.0).max() < 1e-12


def test_cg_reassembly_and_unsupported_models():
    nodes, cells = meshgen.kuhn_mesh(2, 2, 3)
    m, fm, s = _setup(3, 2, nodes, cells)
    fm["Dirichlet"].values[:] = 1.0
    s.assemble()
    v1 = s.getCSR()[2].copy()
    s.assemble()
    assert H.rel_err(s.getCSR()[2], v1) < 1e-14          # clearSystem between assemblies (atomics: equal to rounding, not bitwise; HFX_CG_GATHER=1 for bitwise)
    with pytest.raises(hfox.ErrorHandle, match="nodal field"):
        fm2 = dict(fm); fm2["Solution"] = hfox.Field(m, hfox.Cell, m.getReferenceElement().getNumNodes(), 1)
        s2 = hfox.CGSolver(); s2.setMesh(m); s2.setFieldMap(fm2); s2.setLinSystem(hfox.CudaLinAlgebraInterface(hfox.PetscOpts()))
        s2.setModel(hfox.LaplaceModel(m.getReferenceElement())); s2.setBoundaryModel(hfox.DirichletModel(m.getReferenceElement().getFaceElement()))
        s2.initialize(); s2.allocate()


def _compare_time(dim, order, model, nSteps=3, dt=0.05):
    """implicit Euler steps through the mirror (Solution is the old state and receives the new one) against the oracle running the same steps"""
    nodes, cells = meshgen.kuhn_mesh(3, order, dim, perturb=0.1)
    m = hfox.Mesh(dim, order, "simplex"); m.setMesh(nodes, cells)
    re = m.getReferenceElement()
    nNf = re.getFaceElement().getNumNodes()
    fm = {"Solution": hfox.Field(m, hfox.Node, 1, 1), "Dirichlet": hfox.Field(m, hfox.Face, nNf, 1)}
    rng = np.random.default_rng(4)
    vel = diff = src = None
    if model == "transport":
        vel = np.zeros_like(nodes); vel[:, 0], vel[:, 1] = -(nodes[:, 1] - 0.5), nodes[:, 0] - 0.5
        fm["Velocity"] = hfox.Field(m, hfox.Node, 1, dim); fm["Velocity"].values[:] = vel.ravel()
        mod = hfox.Transport(re)
    else:
        diff = 0.5 + rng.random((nodes.shape[0], 1))
        fm["DiffusionTensor"] = hfox.Field(m, hfox.Node, 1, 1); fm["DiffusionTensor"].values[:] = diff.ravel()
        src = lambda x: np.exp(-10 * sum((xi - 0.5) ** 2 for xi in x))
        mod = hfox.DiffusionSource(re)
    ts = hfox.Euler(re); ts.setTimeStep(dt)
    mod.setTimeScheme(ts)
    s = hfox.CGSolver()
    s.setVerbosity(False); s.setMesh(m); s.setFieldMap(fm); s.setLinSystem(hfox.CudaLinAlgebraInterface(hfox.PetscOpts(rtol=1e-14, maxits=20000)))
    s.setModel(mod); s.setBoundaryModel(hfox.DirichletModel(re.getFaceElement())); s.initialize(); s.allocate()
    if src is not None:
        mod.setSourceFunction(src)
    ore = OracleRefEl(dim, order)
    topo = compute_faces(cells, ore)
    u0 = np.sin(3 * nodes[:, 0]) * np.cos(2 * nodes[:, 1])
    fm["Solution"].values[:] = u0
    dirv = np.zeros(topo["faces"].shape); dirv[topo["boundary"]] = u0[topo["faces"][topo["boundary"]]]
    fm["Dirichlet"].values[:] = dirv.ravel()
    uo = u0.copy()
    for step in range(nSteps):
        s.assemble(); s.solve()
        o = cg.CGOracle(ore, nodes, cells, topo["faces"], topo["boundary"], diff=diff, source=src, vel=vel, diffusion=model != "transport", dt=dt, solOld=uo)
        o.assemble(dirv); uo = o.solve()
        if step == 0:
            rowptr, col, vals, rhs = s.getCSR()
            assert np.array_equal(rowptr, o.rowptr) and np.array_equal(col, o.colidx)
            assert H.rel_err(vals, o.vals) < 1e-12 and H.rel_err(rhs, o.b) < 1e-12
    assert s.stats.converged == 1
    assert H.rel_err(fm["Solution"].values, uo) < 1e-10


@pytest.mark.parametrize("dim,order,model", [(2, 2, "diffsrc"), (3, 2, "diffsrc"), (2, 3, "transport"), (3, 1, "transport")])
def test_cg_implicit_euler_steps_match_oracle(dim, order, model):
    """DiffusionSource / Transport (src/model/DiffusionSource.cpp, Transport.cpp) + Euler (FEModel::compute, Euler.cpp:18-37) through CGSolver: dt A + M, dt F + M u_old"""
    _compare_time(dim, order, model)


def test_cg_steady_convection_diffusion_terms_match_oracle():
    """the Convection entries on their own (steady Transport is singular: assembled entries only)"""
    dim, order = 2, 3
    nodes, cells = meshgen.kuhn_mesh(3, order, dim, perturb=0.1)
    m = hfox.Mesh(dim, order, "simplex"); m.setMesh(nodes, cells)
    re = m.getReferenceElement()
    fm = {"Solution": hfox.Field(m, hfox.Node, 1, 1), "Dirichlet": hfox.Field(m, hfox.Face, re.getFaceElement().getNumNodes(), 1), "Velocity": hfox.Field(m, hfox.Node, 1, dim)}
    vel = np.stack([1.0 + nodes[:, 1], 0.5 - nodes[:, 0]], axis=1)
    fm["Velocity"].values[:] = vel.ravel()
    s = hfox.CGSolver()
    s.setVerbosity(False); s.setMesh(m); s.setFieldMap(fm); s.setLinSystem(hfox.CudaLinAlgebraInterface(hfox.PetscOpts()))
    s.setModel(hfox.Transport(re)); s.setBoundaryModel(hfox.DirichletModel(re.getFaceElement())); s.initialize(); s.allocate()
    s.assemble()
    ore = OracleRefEl(dim, order); topo = compute_faces(cells, ore)
    o = cg.CGOracle(ore, nodes, cells, topo["faces"], topo["boundary"], vel=vel, diffusion=False)
    o.assemble(np.zeros(topo["faces"].shape))
    assert H.rel_err(s.getCSR()[2], o.vals) < 1e-12
    with pytest.raises(hfox.ErrorHandle, match="Velocity"):
        fm2 = {k: v for k, v in fm.items() if k != "Velocity"}
        s2 = hfox.CGSolver(); s2.setVerbosity(False); s2.setMesh(m); s2.setFieldMap(fm2); s2.setLinSystem(hfox.CudaLinAlgebraInterface(hfox.PetscOpts()))
        s2.setModel(hfox.Transport(re)); s2.setBoundaryModel(hfox.DirichletModel(re.getFaceElement())); s2.initialize(); s2.allocate(); s2.assemble()


def test_cg_affine_fast_path_and_mixed_meshes(monkeypatch):
    """cells that are the affine image of the reference element take cg_affine_kernel (reference stiffness matrices), the others the cubature loop: a mesh with a few
    curved cells against the oracle, and the fast path against the cubature loop on the same straight-sided mesh"""
    nodes, cells = meshgen.kuhn_mesh(3, 3, 3, perturb=0.1)
    isv = np.zeros(nodes.shape[0], dtype=bool); isv[np.unique(cells[:, :4])] = True
    bend = np.setdiff1d(np.unique(cells[:7]), np.nonzero(isv)[0])           # the high-order nodes of the first seven cells
    mixed = nodes.copy(); mixed[bend] += 0.004 * np.random.default_rng(8).uniform(-1, 1, size=(bend.size, 3))
    _compare(3, 3, mixed, cells)
    m, fm, s = _setup(3, 3, nodes, cells)
    fm["Dirichlet"].values[:] = 1.5
    s.assemble(); v1 = s.getCSR()[2].copy()
    monkeypatch.setenv("HFX_CG_NO_AFFINE", "1")
    s.assemble(); v2 = s.getCSR()[2].copy()
    assert H.rel_err(v1, v2) < 1e-13 and np.abs(v1).max() > 0


@pytest.mark.parametrize("dim,order,geom", [(2, 4, "simplex"), (3, 2, "simplex"), (3, 2, "orthotope")])
def test_cg_affine_fast_path_with_source_and_euler(dim, order, geom):
    """DiffusionSource without a DiffusionTensor field (D = I) + source + implicit Euler on straight-sided cells: every term of the fast path (C_rs K^_rs, detJ M^,
    detJ w phi f) against the oracle"""
    nodes, cells = (meshgen.kuhn_mesh(3, order, dim, perturb=0.1) if geom == "simplex" else meshgen.box_mesh(3, order, dim))
    m = hfox.Mesh(dim, order, geom); m.setMesh(nodes, cells)
    re = m.getReferenceElement()
    fm = {"Solution": hfox.Field(m, hfox.Node, 1, 1), "Dirichlet": hfox.Field(m, hfox.Face, re.getFaceElement().getNumNodes(), 1)}
    src = lambda x: 1.0 + x[0] * x[1]
    mod = hfox.DiffusionSource(re)
    ts = hfox.Euler(re); ts.setTimeStep(0.03); mod.setTimeScheme(ts)
    s = hfox.CGSolver()
    s.setVerbosity(False); s.setMesh(m); s.setFieldMap(fm); s.setLinSystem(hfox.CudaLinAlgebraInterface(hfox.PetscOpts(rtol=1e-14, maxits=20000)))
    s.setModel(mod); s.setBoundaryModel(hfox.DirichletModel(re.getFaceElement())); s.initialize(); s.allocate()
    mod.setSourceFunction(src)
    ore = OracleRefEl(dim, order, geom); topo = compute_faces(cells, ore)
    u0 = np.cos(2 * nodes[:, 0]) + nodes[:, 1] ** 2
    fm["Solution"].values[:] = u0
    dirv = np.zeros(topo["faces"].shape); dirv[topo["boundary"]] = u0[topo["faces"][topo["boundary"]]]
    fm["Dirichlet"].values[:] = dirv.ravel()
    s.assemble(); s.solve()
    o = cg.CGOracle(ore, nodes, cells, topo["faces"], topo["boundary"], source=src, dt=0.03, solOld=u0)
    o.assemble(dirv); o.solve()
    rowptr, col, vals, rhs = s.getCSR()
    assert H.rel_err(vals, o.vals) < 1e-12 and H.rel_err(rhs, o.b) < 1e-12
    assert H.rel_err(fm["Solution"].values, o.sol) < 1e-10


def test_cg_gather_form_matches_the_atomic_scatter_and_is_bit_reproducible(monkeypatch):
    """the gather form (HFX_CG_GATHER=1: one warp per row, fixed summation order) against the default element-wise scatter with atomics, matrix and right-hand side,
    with a source and an implicit Euler step; two assemblies of the gather form agree bit for bit"""
    monkeypatch.setenv("HFX_CG_GATHER", "1")
    dim, order = 3, 2
    nodes, cells = meshgen.kuhn_mesh(4, order, dim, perturb=0.1)
    m = hfox.Mesh(dim, order, "simplex"); m.setMesh(nodes, cells)
    re = m.getReferenceElement()
    fm = {"Solution": hfox.Field(m, hfox.Node, 1, 1), "Dirichlet": hfox.Field(m, hfox.Face, re.getFaceElement().getNumNodes(), 1)}
    mod = hfox.DiffusionSource(re)
    ts = hfox.Euler(re); ts.setTimeStep(0.02); mod.setTimeScheme(ts)
    s = hfox.CGSolver()
    s.setVerbosity(False); s.setMesh(m); s.setFieldMap(fm); s.setLinSystem(hfox.CudaLinAlgebraInterface(hfox.PetscOpts(rtol=1e-12)))
    s.setModel(mod); s.setBoundaryModel(hfox.DirichletModel(re.getFaceElement())); s.initialize(); s.allocate()
    mod.setSourceFunction(lambda x: x[0] - x[2] ** 2)
    u0 = np.sin(nodes[:, 0] + 2 * nodes[:, 1])
    fm["Solution"].values[:] = u0; fm["Dirichlet"].values[:] = 0.7
    s.assemble(); _, _, v1, r1 = s.getCSR()
    s._upload("Solution"); s.assemble(); _, _, v2, r2 = s.getCSR()
    assert np.array_equal(v1, v2) and np.array_equal(r1, r2)
    monkeypatch.setenv("HFX_CG_GATHER", "0")
    s.assemble(); _, _, v3, r3 = s.getCSR()
    assert H.rel_err(v1, v3) < 1e-13 and H.rel_err(r1, r3) < 1e-13 and np.abs(r1).max() > 0
