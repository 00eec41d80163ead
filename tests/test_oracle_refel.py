"""Pins the oracle's reference-element machinery (oracle/refel.py) with the reference's own known-answer tests, and checks the
product's independent host table builder (libhfx.so: hfx_refel_host_tables, no GPU needed) against the oracle.

Restates tests/unittests/element/TestCubature.cpp:44-49,99-289 and TestReferenceElement.cpp:46-57,117-236."""
import math

import numpy as np
import pytest

from hyperfox_b200 import capi
from oracle.refel import Cubature, ReferenceElement

NIP = {  # TestCubature.cpp:44-49 / SURVEY.md section 8: degree 2p rules
    ("simplex", 3): {2: 4, 4: 14, 6: 24, 8: 46, 10: 81},
    ("simplex", 2): {2: 3, 4: 6, 6: 12, 8: 16, 10: 25},
}
VOL = {("simplex", 1): 2.0, ("simplex", 2): 2.0, ("simplex", 3): 4.0 / 3.0, ("orthotope", 1): 2.0, ("orthotope", 2): 4.0, ("orthotope", 3): 8.0}


@pytest.mark.parametrize("geom,dim", [("simplex", 2), ("simplex", 3)])
def test_cubature_nip_table(geom, dim):
    for deg, n in NIP[(geom, dim)].items():
        assert Cubature(dim, deg, geom).nIP == n


@pytest.mark.parametrize("geom", ["simplex", "orthotope"])
@pytest.mark.parametrize("dim", [1, 2, 3])
def test_cubature_volumes_and_monomials(geom, dim):
    """Weights sum to the reference volume (TestCubature.cpp:99-140); monomials are integrated exactly (:141-289)."""
    maxdeg = 10 if dim == 3 else 12
    for deg in range(0, maxdeg + 1):
        try:
            c = Cubature(dim, deg, geom)
        except KeyError:
            continue
        assert abs(c.weights.sum() - VOL[(geom, dim)]) < 1e-12
        if geom == "orthotope":
            for e in range(0, deg + 1):   # int_{-1}^{1} x^e = 2/(e+1) (e even) per direction
                exact = (2.0 / (e + 1) if e % 2 == 0 else 0.0) * 2.0 ** (dim - 1)
                assert abs((c.weights * c.coords[:, 0] ** e).sum() - exact) < 1e-11, (dim, deg, e)
        else:
            # simplex {x_i >= -1, sum x_i <= 2 - dim}: with y = (x+1)/2 on the unit simplex, int y_0^a = a!/(a+dim)! * 2^dim... (Dirichlet)
            for a in range(0, deg + 1):
                exact = VOL[(geom, dim)] * math.factorial(dim) * math.factorial(a) / math.factorial(a + dim)
                val = (c.weights * ((c.coords[:, 0] + 1.0) / 2.0) ** a).sum()
                assert abs(val - exact) < 1e-11, (dim, deg, a)


def test_reference_element_construction_errors():
    """TestReferenceElement.cpp:14-31."""
    with pytest.raises(ValueError):
        ReferenceElement(2, 2, "non existant polytope")
    with pytest.raises(Exception):
        ReferenceElement(100, 1, "simplex")
    with pytest.raises(Exception):
        ReferenceElement(2, 100, "simplex")
    with pytest.raises(capi.ErrorHandle):
        capi.host_refel_tables(3, 6)       # 3-D simplex order <= 5 (ReferenceElement.cpp:623)


@pytest.mark.parametrize("dim", [1, 2, 3])
@pytest.mark.parametrize("order", [0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("geom", ["simplex", "orthotope"])
def test_reference_element_counts_lagrange_monomials(dim, order, geom):
    if geom == "orthotope" and dim == 3 and order > 2:
        pytest.skip("3-D orthotope order <= 2 (ReferenceElement.cpp:627)")
    re = ReferenceElement(dim, order, geom)
    nN = math.comb(order + dim, dim) if geom == "simplex" else (order + 1) ** dim
    assert re.nNodes == nN and re.nodes.shape == (nN, dim)
    assert re.nFaces == (dim + 1 if geom == "simplex" else 2 * dim)
    assert re.faceElement.nNodes == (math.comb(order + dim - 1, dim - 1) if geom == "simplex" else (order + 1) ** (dim - 1))
    if order > 0:
        assert len(re.faceNodes) == re.nFaces and len(re.faceNodes[0]) == re.faceElement.nNodes
    L = np.array([re.interpolate(p) for p in re.nodes])
    assert np.abs(L - np.eye(nN)).max() < 1e-10                      # Lagrange property, :117-145
    pt = [-0.5] * dim
    vals = (re.nodes ** order).sum(1)
    assert re.interpolate(pt) @ vals == pytest.approx(sum(x ** order for x in pt), rel=1e-9, abs=1e-10)   # :146-190
    d = re.interpolate_deriv(pt)
    for k in range(dim):
        exact = order * pt[k] ** (order - 1) if order > 0 else 0.0
        assert d[:, k] @ vals == pytest.approx(exact, rel=1e-8, abs=1e-9)                                   # :191-236


def test_tet_face_order_and_reference_normals():
    """Face vertex sets of the tet {3,1,0},{2,1,3},{2,3,0},{0,1,2} (ReferenceElement.cpp:1051-1054) and the reference normals of
    tests/TestUtils.h.in:251-264."""
    from oracle import lib as O
    re = ReferenceElement(3, 1)
    assert re.faceNodes == [[3, 1, 0], [2, 1, 3], [2, 3, 0], [0, 1, 2]]
    for dim, refn in ((2, [[0, -1], [2 ** -0.5] * 2, [-1, 0]]), (3, [[0, -1, 0], [3 ** -0.5] * 3, [-1, 0, 0], [0, 0, -1]])):
        for p in (1, 2, 3):
            r = ReferenceElement(dim, p)
            rc = O.RefElC(r)
            jac, inv, dV, nrm = O.element_geometry(rc, r.nodes)
            assert np.abs(jac[:r.nIP] - np.eye(dim)).max() < 1e-12
            for f in range(rc.nFc):
                assert np.abs(nrm[f * rc.nIPf:(f + 1) * rc.nIPf] - np.array(refn[f])).max() < 1e-12


@pytest.mark.parametrize("dim,order", [(2, 1), (2, 2), (2, 3), (2, 4), (2, 5), (3, 1), (3, 2), (3, 3), (3, 4), (3, 5)])
def test_product_host_tables_match_oracle(dim, order):
    """The product builds its tables with its own C++ code (csrc/host/hfx_refel.cpp); they must agree with the oracle's numpy build."""
    t = capi.host_refel_tables(dim, order)
    o = ReferenceElement(dim, order).tables()
    for k in ("nN", "nNf", "nFc", "nIP", "nIPf"):
        assert t[k] == o[k]
    assert np.array_equal(t["faceNodes"], o["faceNodes"])
    assert np.abs(t["nodes"] - ReferenceElement(dim, order).nodes).max() < 1e-14
    for k in ("w", "fw"):
        assert np.array_equal(t[k], o[k])                                     # same table data, bit for bit
    for k, tol in (("shape", 5e-13), ("dshape", 5e-12), ("fshape", 5e-13), ("fdshape", 5e-12)):
        assert np.abs(t[k].reshape(o[k].shape) - o[k]).max() < tol, k
