"""ORACLE (test infrastructure, never imported by the product path).

CPU restatement, in numpy, of the reference's reference-element machinery:

  * Cubature                 -> /root/reference/src/element/Cubature.cpp:52-59 (rule lookup), tables :66-2730
  * ReferenceElement ctor    -> /root/reference/src/element/ReferenceElement.cpp:5-26
  * node database            -> ReferenceElement.cpp:614-1150 (1-D Lobatto :636-878, orthotope :885-1004, simplex :1006-1149)
  * face-node maps           -> ReferenceElement.cpp:136-202
  * node-to-mode map         -> ReferenceElement.cpp:204-218
  * orthonormal modes        -> ReferenceElement.cpp:231-317 (collapsed coordinates + Jacobi polynomials)
  * derivative modes         -> ReferenceElement.cpp:320-445 (incl. the epsilon = 1e-6 guard at collapsed vertices)
  * inverse Vandermonde      -> ReferenceElement.cpp:447-455
  * interpolate / Deriv      -> ReferenceElement.cpp:542-577

Third-party arithmetic restated here: Boost.Math 1.76 `jacobi`, `jacobi_derivative` (three-term recurrence;
d/dx P_n^{a,b} = (n+a+b+1)/2 P_{n-1}^{a+1,b+1}) and Eigen `inverse()` (numpy.linalg.inv).

Parity status: pinned by the reference's own known-answer tests (tests/test_oracle_refel.py restates
tests/unittests/element/TestCubature.cpp and TestReferenceElement.cpp).
"""
import itertools
import json
import math
import os

import numpy as np

_TABLES = None


def _tables():
    global _TABLES
    if _TABLES is None:
        raw = json.load(open(os.path.join(os.path.dirname(__file__), "tables", "tables.json")))
        fh = float.fromhex
        _TABLES = {
            "nip_map": {tuple(int(x) if i < 2 else x for i, x in enumerate(k.split(","))): v
                        for k, v in raw["nip_map"].items()},
            "rules": {tuple(int(x) if i < 2 else x for i, x in enumerate(k.split(","))):
                      (np.array([[fh(x) for x in p] for p in v["coords"]], dtype=np.float64).reshape(len(v["weights"]), -1),
                       np.array([fh(x) for x in v["weights"]], dtype=np.float64))
                      for k, v in raw["rules"].items()},
            "lobatto": {int(k): [fh(x) for x in v] for k, v in raw["lobatto"].items()},
        }
    return _TABLES


MAX_ORDER = {(0, "simplex"): 10, (1, "simplex"): 10, (2, "simplex"): 10, (3, "simplex"): 5,
             (0, "orthotope"): 5, (1, "orthotope"): 5, (2, "orthotope"): 5, (3, "orthotope"): 2}
MAX_CUB_ORDER = {0: 20, 1: 20, 2: 20, 3: 10}


class Cubature:
    """Cubature.cpp:5-26,52-59."""

    def __init__(self, dim, degree, geom):
        if dim > 3:
            raise ValueError("Cubature : Constructor : dimension too large")
        if degree > MAX_CUB_ORDER[dim]:
            raise ValueError("Cubature : Constructor : the requested polynomial order (%d) is too large for dimension %d"
                             % (degree, dim))
        t = _tables()
        self.dim, self.degree, self.geom = dim, degree, geom
        if dim == 0:
            self.nIP = 0
            self.coords = np.zeros((0, 0))
            self.weights = np.zeros((0,))
            return
        self.nIP = t["nip_map"][(dim, degree, geom)]
        c, w = t["rules"][(dim, self.nIP, geom)]
        self.coords = c.copy()
        self.weights = w.copy()


# ---------------------------------------------------------------------------------------------
# Jacobi polynomials (Boost.Math jacobi.hpp restated)
def jacobi(n, a, b, x):
    if n == 0:
        return 1.0
    y0 = 1.0
    y1 = (a + 1.0) + (a + b + 2.0) * (x - 1.0) / 2.0
    k = 2
    while k <= n:
        # Boost: gamma1 = 2k(k+a+b)(2k+a+b-2); ...
        k_ = float(k)
        denom = 2.0 * k_ * (k_ + a + b) * (2.0 * k_ + a + b - 2.0)
        g1 = (2.0 * k_ + a + b - 1.0) * ((2.0 * k_ + a + b) * (2.0 * k_ + a + b - 2.0) * x + a * a - b * b)
        g0 = -2.0 * (k_ + a - 1.0) * (k_ + b - 1.0) * (2.0 * k_ + a + b)
        y2 = (g1 * y1 + g0 * y0) / denom
        y0, y1 = y1, y2
        k += 1
    return y1


def jacobi_derivative(n, a, b, x, k=1):
    assert k == 1
    if n == 0:
        return 0.0
    return (n + a + b + 1.0) / 2.0 * jacobi(n - 1, a + 1.0, b + 1.0, x)


# ---------------------------------------------------------------------------------------------
_FACE_ORIENT = {
    (1, "simplex"): ([[-1.0], [1.0]], [[0], [1]]),
    (2, "simplex"): ([[-1.0, -1.0], [1.0, -1.0], [-1.0, 1.0]], [[0, 1], [1, 2], [2, 0]]),
    (3, "simplex"): ([[-1.0, -1.0, -1.0], [1.0, -1.0, -1.0], [-1.0, 1.0, -1.0], [-1.0, -1.0, 1.0]],
                     [[0, 1], [1, 2], [2, 0], [3, 0], [3, 1], [3, 2],
                      [3, 1, 0], [2, 1, 3], [2, 3, 0], [0, 1, 2]]),
    (1, "orthotope"): ([[-1.0], [1.0]], [[0], [1]]),
    (2, "orthotope"): ([[-1.0, -1.0], [1.0, -1.0], [1.0, 1.0], [-1.0, 1.0]], [[0, 1], [1, 2], [2, 3], [3, 0]]),
    (3, "orthotope"): ([[-1.0, -1.0, -1.0], [1.0, -1.0, -1.0], [1.0, 1.0, -1.0], [-1.0, 1.0, -1.0],
                        [-1.0, -1.0, 1.0], [1.0, -1.0, 1.0], [1.0, 1.0, 1.0], [-1.0, 1.0, 1.0]],
                       [[0, 1], [1, 2], [2, 3], [3, 0], [0, 4], [1, 5], [2, 6], [3, 7], [4, 5], [5, 6], [6, 7], [7, 4],
                        [1, 0, 4, 5], [2, 1, 5, 6], [3, 2, 6, 7], [0, 3, 7, 4], [0, 1, 2, 3], [5, 4, 7, 6]]),
}

_NODE_DB = {}


def _remove_duplicates(v):
    """ReferenceElement.cpp:1164-1188 (keep first occurrence, tol 1e-12 per coordinate)."""
    s = len(v)
    dup = set()
    for i in range(s):
        for j in range(i + 1, s):
            if all(abs(v[i][d] - v[j][d]) < 1e-12 for d in range(len(v[i]))):
                dup.add(j)
    return [v[i] for i in range(s) if i not in dup]


def node_database(dim, order, geom):
    key = (dim, order, geom)
    if key in _NODE_DB:
        return _NODE_DB[key]
    if dim == 0:
        res = [[]]
    elif dim == 1:
        res = [[x] for x in _tables()["lobatto"][order]]
    elif order == 0:
        res = [[0.0] * dim]
    else:
        principal, face_list = _FACE_ORIENT[(dim, geom)]
        val = [list(p) for p in principal]
        for q in range(1, dim):
            nm1 = _FACE_ORIENT[(q, geom)][0]
            inter = np.array([[nm1[l + 1][k] - nm1[0][k] for l in range(q)] for k in range(q)], dtype=np.float64)
            inv_inter = np.linalg.inv(inter)
            fe_nodes = node_database(q, order, geom)
            for f in face_list:
                if len(f) != len(nm1):
                    continue
                T = np.array([[principal[f[l + 1]][d] - principal[f[0]][d] for l in range(q)] for d in range(dim)])
                T = T @ inv_inter
                v0 = np.array(principal[f[0]])
                origin = np.array(nm1[0]) if geom == "simplex" else np.array(fe_nodes[0])
                for nd in fe_nodes:
                    val.append(list(T @ (np.array(nd) - origin) + v0))
        n_face_nodes = len(val)
        if order > 1:
            lob = list(_tables()["lobatto"][order])
            end = lob.pop(1)
            lob.append(end)
            combos = [list(c) for c in itertools.product(range(1, order), repeat=dim)]
            if geom == "orthotope":
                for c in combos:
                    val.append([lob[c[l]] for l in range(dim)])
            else:
                combos = [c for c in combos if sum(c) < order]
                coef = {}
                for c in combos:
                    x = 2.0 + dim * lob[c[0]]
                    for l in range(1, dim):
                        x -= lob[c[l]]
                    x -= lob[order - sum(c)]
                    x *= 1.0 / (dim + 1.0)
                    x -= 1.0
                    coef[tuple(c)] = x
                for c in combos:
                    if dim == 2:
                        val.append([coef[(c[0], c[1])], coef[(c[1], c[0])]])
                    else:
                        val.append([coef[(c[0], c[1], c[2])], coef[(c[1], c[0], c[2])], coef[(c[2], c[1], c[0])]])
        res = _remove_duplicates(val)
    _NODE_DB[key] = res
    return res


class ReferenceElement:
    def __init__(self, dim, order, geom="simplex"):
        if geom in ("quad", "hex"):
            geom = "orthotope"
        if geom not in ("simplex", "orthotope"):
            raise ValueError("ReferenceElement : setGeometry : Element type %s is not yet supported." % geom)
        if dim > 3 or dim < 0:
            raise ValueError("ReferenceElement : setDim : dimension %d not supported" % dim)
        if order > MAX_ORDER[(dim, geom)] or order < 0:
            raise ValueError("ReferenceElement : setOrder : The interpolation order %d is not yet supported for dimension %d."
                             % (order, dim))
        self.dim, self.order, self.geom = dim, order, geom
        db = node_database(dim, order, geom)
        self.nodes = np.array(db, dtype=np.float64).reshape(len(db), dim)
        self.nNodes = len(db)
        if dim > 0:
            if geom == "simplex":
                assert self.nNodes == math.comb(order + dim, dim)
            else:
                assert self.nNodes == (order + 1) ** dim
        self.cubature = Cubature(dim, 2 * order if geom == "simplex" else 4 * order, geom)
        self.nFaces = dim + 1 if geom == "simplex" else 2 * dim
        self.faceElement = ReferenceElement(dim - 1, order, geom) if dim != 0 else None
        self._face_nodes()
        if dim > 0:
            combos = [list(c) for c in itertools.product(range(order + 1), repeat=dim)]
            if geom == "simplex":
                combos = [c for c in combos if sum(c) <= order]
            self.modeMap = combos
            V = np.array([self.compute_modes(self.nodes[i]) for i in range(self.nNodes)])
            self.invV = np.linalg.inv(V)
            self.ipCoords = self.cubature.coords
            self.ipWeights = self.cubature.weights
            self.nIP = self.cubature.nIP
            self.ipShape = np.array([self.interpolate(p) for p in self.ipCoords]).reshape(self.nIP, self.nNodes)
            self.ipDShape = np.array([self.interpolate_deriv(p) for p in self.ipCoords]).reshape(self.nIP, self.nNodes, dim)
        else:
            self.nIP = 0

    # ReferenceElement.cpp:136-202
    def _face_nodes(self):
        dim, order, geom = self.dim, self.order, self.geom
        self.faceNodes = [[] for _ in range(self.nFaces)]
        self.innerNodes = []
        if dim == 0 or order == 0:
            return
        if dim == 1:
            self.faceNodes = [[0], [1]]
            self.innerNodes = list(range(2, order + 1))
            return
        nm1 = _FACE_ORIENT[(dim - 1, geom)][0]
        q = dim - 1
        inter = np.array([[nm1[l + 1][k] - nm1[0][k] for l in range(q)] for k in range(q)], dtype=np.float64)
        inv_inter = np.linalg.inv(inter)
        principal, face_list = _FACE_ORIENT[(dim, geom)]
        fe_nodes = self.faceElement.nodes
        pre0 = np.array(nm1[0])
        idx = 0
        for f in face_list:
            if len(f) != len(nm1):
                continue
            T = np.array([[principal[f[l + 1]][d] - principal[f[0]][d] for l in range(q)] for d in range(dim)]) @ inv_inter
            v0 = np.array(principal[f[0]])
            lst = []
            for l in range(fe_nodes.shape[0]):
                nv = T @ (fe_nodes[l] - pre0) + v0
                found = -1
                for k in range(self.nNodes):
                    if np.all(np.abs(self.nodes[k] - nv) <= 1e-12):
                        found = k
                        break
                if found < 0:
                    raise RuntimeError("ReferenceElement : determineFaceNodes : could not find one of the face nodes.")
                lst.append(found)
            self.faceNodes[idx] = lst
            idx += 1
        n_face_nodes = max(max(f) for f in self.faceNodes) + 1
        self.innerNodes = list(range(n_face_nodes, self.nNodes))

    # ReferenceElement.cpp:231-267
    def map_coords_simplex(self, c):
        d = self.dim
        if d == 1:
            return [c[0]]
        if d == 2:
            m0 = 2.0 * (1.0 + c[0]) / (1.0 - c[1]) - 1.0 if c[1] != 1.0 else -1.0
            return [m0, c[1]]
        m0 = -2.0 * (1.0 + c[0]) / (c[1] + c[2]) - 1.0 if (c[1] + c[2]) != 0.0 else -1.0
        m1 = 2.0 * (1.0 + c[1]) / (1.0 - c[2]) - 1.0 if c[2] != 1.0 else -1.0
        return [m0, m1, c[2]]

    # ReferenceElement.cpp:269-317
    def compute_modes(self, point):
        res = np.ones(self.nNodes)
        if self.geom == "orthotope":
            for i, mode in enumerate(self.modeMap):
                for k in range(self.dim):
                    res[i] *= jacobi(mode[k], 0.0, 0.0, point[k])
            return res
        mc = self.map_coords_simplex(point)
        for j, mode in enumerate(self.modeMap):
            for k in range(self.dim):
                if k == 0:
                    res[j] *= jacobi(mode[0], 0.0, 0.0, mc[0])
                elif k == 1:
                    res[j] *= math.sqrt(2.0) * jacobi(mode[1], 2.0 * mode[0] + 1.0, 0.0, mc[1]) * (1 - mc[1]) ** mode[0]
                else:
                    res[j] *= 2.0 * jacobi(mode[2], 2.0 * (mode[1] + mode[0] + 1.0), 0.0, mc[2]) * (1 - mc[2]) ** (mode[1] + mode[0])
        return res

    # ReferenceElement.cpp:320-445
    def compute_deriv_modes(self, point):
        dim = self.dim
        res = np.ones((self.nNodes, dim))
        if self.geom == "orthotope":
            for i, mode in enumerate(self.modeMap):
                for k in range(dim):
                    for l in range(dim):
                        if l != k:
                            res[i, l] *= jacobi(mode[k], 0.0, 0.0, point[k])
                        else:
                            res[i, l] *= jacobi_derivative(mode[k], 0.0, 0.0, point[k], 1)
            return res
        mc = self.map_coords_simplex(point)
        J = np.zeros((dim, dim))
        eps = 1e-6
        if dim == 1:
            J[0, 0] = 1.0
        elif dim == 2:
            if point[1] != 1.0:
                J[0, 0] = 2.0 / (1.0 - point[1])
                J[1, 0] = 2.0 * (1.0 + point[0]) * (1.0 / (1.0 - point[1]) ** 2.0)
            else:
                J[0, 0] = 2.0 / eps
                J[1, 0] = 0.0
            J[1, 1] = 1.0
        else:
            if (point[1] + point[2]) != 0.0:
                J[0, 0] = -2.0 / (point[1] + point[2])
                J[1, 0] = 2.0 * (1.0 + point[0]) * (1.0 / (point[1] + point[2]) ** 2.0)
                J[2, 0] = 2.0 * (1.0 + point[0]) * (1.0 / (point[1] + point[2]) ** 2.0)
            else:
                J[0, 0] = -2.0 / eps
            if point[2] != 1.0:
                J[1, 1] = 2.0 / (1.0 - point[2])
                J[2, 1] = 2.0 * (1.0 + point[1]) * (1.0 / (1.0 - point[2]) ** 2.0)
            else:
                J[1, 1] = 2.0 / eps
            J[2, 2] = 1.0
        s2 = math.sqrt(2.0)
        for j, mode in enumerate(self.modeMap):
            g = np.ones(dim)
            for k in range(dim):
                for l in range(dim):
                    if l == 0:
                        if l != k:
                            g[k] *= jacobi(mode[0], 0.0, 0.0, mc[0])
                        else:
                            g[k] *= jacobi_derivative(mode[0], 0.0, 0.0, mc[0], 1)
                    elif l == 1:
                        a = 2.0 * mode[0] + 1.0
                        if l != k:
                            g[k] *= s2 * jacobi(mode[1], a, 0.0, mc[1]) * (1 - mc[1]) ** mode[0]
                        else:
                            power = max(mode[0] - 1, 0)
                            g[k] *= (s2 * jacobi_derivative(mode[1], a, 0.0, mc[1], 1) * (1 - mc[1]) ** mode[0]
                                     + s2 * jacobi(mode[1], a, 0.0, mc[1]) * (-mode[0] * (1 - mc[1]) ** power))
                    else:
                        a = 2.0 * (mode[1] + mode[0] + 1.0)
                        if l != k:
                            g[k] *= 2.0 * jacobi(mode[2], a, 0.0, mc[2]) * (1 - mc[2]) ** (mode[1] + mode[0])
                        else:
                            power = max(mode[1] + mode[0] - 1, 0)
                            g[k] *= (2.0 * jacobi_derivative(mode[2], a, 0.0, mc[2], 1) * (1 - mc[2]) ** (mode[1] + mode[0])
                                     + 2.0 * jacobi(mode[2], a, 0.0, mc[2]) * (-(mode[1] + mode[0]) * (1 - mc[2]) ** power))
            res[j] = J @ g
        return res

    # ReferenceElement.cpp:542-577
    def interpolate(self, point):
        if self.dim == 0:
            return np.ones(self.nNodes)
        return self.compute_modes(point) @ self.invV

    def interpolate_deriv(self, point):
        if self.dim == 0:
            return np.zeros((self.nNodes, 1))
        dm = self.compute_deriv_modes(point)
        return np.stack([dm[:, k] @ self.invV for k in range(self.dim)], axis=1)

    def tables(self):
        """Flat constant tensors consumed by the per-element oracle (and compared with the product's builder)."""
        fe = self.faceElement
        return dict(dim=self.dim, order=self.order, nN=self.nNodes, nNf=fe.nNodes, nFc=self.nFaces,
                    nIP=self.nIP, nIPf=fe.nIP,
                    shape=np.ascontiguousarray(self.ipShape), dshape=np.ascontiguousarray(self.ipDShape),
                    w=np.ascontiguousarray(self.ipWeights),
                    fshape=np.ascontiguousarray(fe.ipShape), fdshape=np.ascontiguousarray(fe.ipDShape),
                    fw=np.ascontiguousarray(fe.ipWeights),
                    faceNodes=np.array(self.faceNodes, dtype=np.int32))
