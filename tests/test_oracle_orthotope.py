"""The oracle on the reference's orthotope elements (quads order 1-5, hexes order 1-2: ReferenceElement.cpp:624-627,885-1004), CPU only.
Pins: the reference's smallest end-to-end known answer (TestHDGSolver.cpp:16-100: tau = 1, Dirichlet = 3 => Solution = 3, Flux = 0,
Trace = 3 to 1e-12) restated on quads / hexes, the outward unit normals of the reference orthotope, the harmonic regression solution
of TestHDGLaplace.cpp:20-28 converging at the spectral rate, and the structured mesh generator (hyperfox_b200.meshgen.box_mesh)."""
import numpy as np
import pytest

from hyperfox_b200 import meshgen
from oracle import lib as O
from oracle.refel import ReferenceElement
from tests import helpers as H


@pytest.mark.parametrize("dim,order", [(2, 1), (2, 3), (2, 5), (3, 1), (3, 2)])
def test_box_mesh_is_conforming(dim, order):
    N = 3
    nodes, cells = meshgen.box_mesh(N, order, dim, perturb=0.2)
    re = ReferenceElement(dim, order, "orthotope")
    assert nodes.shape == ((N * order + 1) ** dim, dim) and cells.shape == (N ** dim, re.nNodes)
    assert np.array_equal(np.unique(cells), np.arange(nodes.shape[0]))
    # every node of every cell sits at the multilinear image of its reference position (shared nodes agree between cells)
    lin = ReferenceElement(dim, 1, "orthotope")
    phi = np.array([lin.interpolate(p) for p in re.nodes]).reshape(re.nNodes, 2 ** dim)
    img = np.einsum("nv,cvd->cnd", phi, nodes[cells[:, :2 ** dim]])
    assert np.abs(img - nodes[cells]).max() < 1e-14
    topo = H.compute_faces(cells, re)
    nB = {2: 4 * N, 3: 6 * N * N}[dim]
    assert topo["boundary"].size == nB
    assert topo["faces"].shape[0] == (dim * N ** dim * 2 + nB) // 2
    rc = O.RefElC(re)
    for c in range(cells.shape[0]):
        jac, inv, dV, nrm = O.element_geometry(rc, nodes[cells[c]])
        assert (dV > 0).all()
    vol = sum(O.element_geometry(rc, nodes[cells[c]])[2][:re.nIP].sum() for c in range(cells.shape[0]))
    assert abs(vol - 1.0) < 1e-13


@pytest.mark.parametrize("dim", [2, 3])
def test_reference_orthotope_normals(dim):
    """Outward unit normals of the reference orthotope at every face cubature point (HDGBase.cpp:33-65): +-e_d."""
    re = ReferenceElement(dim, 2, "orthotope")
    rc = O.RefElC(re)
    jac, inv, dV, nrm = O.element_geometry(rc, re.nodes)
    nIPf = re.faceElement.nIP
    nrm = nrm.reshape(re.nFaces, nIPf, dim)
    for f in range(re.nFaces):
        ctr = re.nodes[re.faceNodes[f]].mean(0)                  # centre of the face = its outward direction on [-1,1]^dim
        assert np.abs(nrm[f] - ctr[None, :]).max() < 1e-14
    assert abs(dV[:re.nIP].sum() - 2.0 ** dim) < 1e-13
    assert np.abs(dV[re.nIP:].reshape(re.nFaces, nIPf).sum(1) - 2.0 ** (dim - 1)).max() < 1e-13


@pytest.mark.parametrize("dim,order", [(2, 1), (2, 2), (2, 4), (3, 1), (3, 2)])
def test_constant_solution(dim, order):
    case = H.make_case(dim, order, N=2, perturb=0.15, geom="orthotope")
    case["fields"]["Dirichlet"][case["topo"]["boundary"]] = 3.0
    o = H.run_oracle(case, rtol=1e-15)
    assert np.abs(o.sol - 3.0).max() < 1e-12
    assert np.abs(o.flux).max() < 1e-11
    assert np.abs(o.trace - 3.0).max() < 1e-12


def test_harmonic_solution_converges_spectrally():
    errs = []
    for order in (1, 2, 3, 4, 5):
        case = H.make_case(2, order, N=3, perturb=0.15, geom="orthotope")
        o = H.run_oracle(case)
        ana = case["ana"][case["cells"]]
        errs.append(np.abs(o.sol.reshape(ana.shape) - ana).max())
    assert all(b < 0.25 * a for a, b in zip(errs, errs[1:])), errs
    assert errs[-1] < 1e-6


def test_qr_and_lu_condensation_agree_on_hexes():
    case = H.make_case(3, 2, N=2, perturb=0.15, model="cdrs", diff="tensor", tau_double=True, geom="orthotope", seed=5)
    a, b = H.run_oracle(case, useLU=0, solve=False), H.run_oracle(case, useLU=1, solve=False)
    assert H.rel_err(a.S, b.S) < 1e-12 and H.rel_err(a.vals, b.vals) < 1e-12


def _both_topologies(dim, order, cells):
    """Oracle (numpy) and product (host C++ behind the C ABI: hfx_host_compute_faces) topology of the same cells."""
    from hyperfox_b200 import capi
    cells = np.asarray(cells, dtype=np.int32)
    o = H.compute_faces(cells, ReferenceElement(dim, order, "orthotope"))
    p = capi.host_compute_faces(dim, order, cells, 1)
    for k in ("faces", "cell2face", "face2cell", "boundary"):
        assert np.array_equal(np.asarray(o[k]).reshape(-1), np.asarray(p[k]).reshape(-1)), k
    return o


def test_reference_quad_mesh_known_answers():
    """tests/unittests/mesh/TestMesh.cpp:20-69,194-312 (four linear quads around node 4) and :314-420 (one order-2 quad): face counts,
    face membership, face-to-cell adjacency and the boundary set, for the oracle and for the product's host topology builder."""
    quads = [[0, 5, 4, 8], [1, 6, 4, 5], [2, 7, 4, 6], [3, 8, 4, 7]]
    t = _both_topologies(2, 1, quads)
    assert t["faces"].shape == (12, 2) and t["boundary"].size == 8
    expected = [{0, 5}, {5, 1}, {1, 6}, {6, 2}, {2, 7}, {7, 3}, {3, 8}, {8, 0}, {5, 4}, {6, 4}, {7, 4}, {8, 4}]
    got = [set(f) for f in t["faces"].tolist()]
    assert sorted(map(sorted, got)) == sorted(map(sorted, expected))
    boundary_nodes = {0, 1, 2, 3, 5, 6, 7, 8}
    for F in t["boundary"]:
        assert set(t["faces"][F].tolist()) <= boundary_nodes
    for F, (c0, c1) in enumerate(t["face2cell"].tolist()):     # interior faces carry node 4 and two ascending cells
        assert (c1 >= 0) == (4 in got[F])
        if c1 >= 0:
            assert c0 < c1 and all(set(got[F]) <= set(quads[c]) for c in (c0, c1))
    for c in range(4):
        for f in range(4):
            assert set(got[t["cell2face"][c, f]]) <= set(quads[c])
    # order 2: one 9-node quad, faces {0,1,5}, {1,2,6}, {2,3,7}, {3,0,8}, all on the boundary, all adjacent to cell 0 only
    t2 = _both_topologies(2, 2, [[0, 1, 2, 3, 5, 6, 7, 8, 4]])
    assert sorted(map(sorted, t2["faces"].tolist())) == sorted(map(sorted, [[0, 1, 5], [1, 2, 6], [2, 3, 7], [3, 0, 8]]))
    assert t2["boundary"].size == 4 and (t2["face2cell"][:, 0] == 0).all() and (t2["face2cell"][:, 1] == -1).all()


@pytest.mark.parametrize("dim,order", [(2, 3), (3, 2)])
def test_box_mesh_topology_matches_between_oracle_and_product(dim, order):
    nodes, cells = meshgen.box_mesh(3, order, dim)
    _both_topologies(dim, order, cells)
