"""ORACLE (test infrastructure): Gmsh 2.2 ASCII reader and the straight-sided order-p mesh generator,
numpy / pure-Python restatement; the product's implementation is host C++ behind the C ABI (hyperfox_b200/csrc/host/hfx_meshio.cpp,
hyperfox_b200/meshio.py).

Restates /root/reference/tools/convertGmsh2H5HO.cpp:117-257 (generateHigherOrderMesh), :259-363 (readMesh through MOAB) and
:366-397 (generateCellNodes) for simplex meshes, i.e. what produced every ressources/meshes/regression/*_ord-p.h5 fixture from its
.msh file.  MOAB (absent here) decides the numbering of the intermediate entities; the convention restated is the one SURVEY.md
section 8c documents for Mesh::computeFaces:

  * entities present in the .msh file keep their file order, per type (in 3-D meshes the boundary triangles of the file are entities
    1..nB of dimension 2, with the file's vertex order);
  * `get_adjacencies(cells, k, create=true, UNION)` walks the cells in ascending id and, inside a cell, the sub-entities in the
    canonical (MBCN) order -- edges (0,1),(1,2),(2,0),(0,3),(1,3),(2,3); tet faces (0,1,3),(1,2,3),(0,3,2),(0,2,1) -- and a sub-entity
    that does not exist yet gets the next id, with the vertex order of that first appearance;
  * the adjacent sub-entities of one cell are returned in ascending id (a MOAB Range is sorted).

The node numbering of the generated mesh depends on all three, and the reference's own .h5 fixtures pin it:
tests/test_meshio.py regenerates them from the .msh files (with this restatement and with the product) and compares cells bit-exactly,
node coordinates to rounding.  Pinned: 22 reference fixtures.
"""
import numpy as np

from .refel import ReferenceElement

_EDGES = {2: [(0, 1), (1, 2), (2, 0)], 3: [(0, 1), (1, 2), (2, 0), (0, 3), (1, 3), (2, 3)]}
_TET_FACES = [(0, 1, 3), (1, 2, 3), (0, 3, 2), (0, 2, 1)]
_GMSH_NODES = {15: 1, 1: 2, 2: 3, 4: 4}          # point, line, triangle, tetrahedron
_GMSH_DIM = {15: 0, 1: 1, 2: 2, 4: 3}


def read_msh(path):
    """Gmsh 2.2 ASCII.  Returns (nodes [n,3] ordered by ascending node tag, {topological dim: int array [m, dim+1]} in file order)."""
    with open(path) as f:
        tok = f.read().split("\n")
    i = 0
    nodes, ids, elems = None, None, {1: [], 2: [], 3: []}
    while i < len(tok):
        line = tok[i].strip()
        if line == "$MeshFormat":
            ver = tok[i + 1].split()
            if not ver[0].startswith("2") or ver[1] != "0":
                raise ValueError("meshio : read_msh : only the Gmsh 2.x ASCII format is supported")
            i += 3
        elif line == "$Nodes":
            n = int(tok[i + 1])
            raw = np.array([tok[i + 2 + k].split() for k in range(n)], dtype=np.float64)
            ids, nodes = raw[:, 0].astype(np.int64), raw[:, 1:4]
            i += n + 3
        elif line == "$Elements":
            n = int(tok[i + 1])
            for k in range(n):
                p = tok[i + 2 + k].split()
                ty, ntags = int(p[1]), int(p[2])
                if ty not in _GMSH_NODES:
                    raise ValueError("meshio : read_msh : element type %d is not supported (linear simplices only)" % ty)
                if _GMSH_DIM[ty] > 0:
                    elems[_GMSH_DIM[ty]].append([int(x) for x in p[3 + ntags:3 + ntags + _GMSH_NODES[ty]]])
            i += n + 3
        else:
            i += 1
    order = np.argsort(ids, kind="stable")
    remap = -np.ones(ids.max() + 1, dtype=np.int64)
    remap[ids[order]] = np.arange(ids.size)
    out = {d: remap[np.array(v, dtype=np.int64).reshape(-1, d + 1)] for d, v in elems.items() if v}
    return nodes[order], out


def _sub_entities(dim, cells, existing):
    """The entities of every topological dimension below `dim`, numbered as restated above.
    Returns ({k: connectivity [m, k+1]}, {k: per-cell ascending ids [nCells, nSub]})."""
    conn, adj = {}, {}
    for k in range(1, dim):
        table = _EDGES[dim] if k == 1 else _TET_FACES
        ents, key2id = [], {}
        for e in (existing.get(k, np.zeros((0, k + 1), dtype=np.int64))).tolist():
            key2id.setdefault(tuple(sorted(e)), len(ents))
            ents.append(e)
        percell = np.empty((cells.shape[0], len(table)), dtype=np.int64)
        for c, cell in enumerate(cells.tolist()):
            for s, loc in enumerate(table):
                vs = [cell[v] for v in loc]
                key = tuple(sorted(vs))
                j = key2id.get(key)
                if j is None:
                    j = key2id[key] = len(ents)
                    ents.append(vs)
                percell[c, s] = j
        conn[k] = np.array(ents, dtype=np.int64)
        adj[k] = np.sort(percell, axis=1)
    return conn, adj


def _cell_nodes(ref, lin):
    """generateCellNodes (convertGmsh2H5HO.cpp:366-397): affine image of the reference nodes, vertex 0 + T (xi - xi_0)."""
    td = ref.shape[1]
    locT = (ref[1:td + 1] - ref[0]).T
    T = (lin[1:td + 1] - lin[0]).T @ np.linalg.inv(locT)
    return (ref - ref[0]) @ T.T + lin[0]


def high_order_from_linear(dim, order, lin_nodes, cells, existing=None):
    """generateHigherOrderMesh (convertGmsh2H5HO.cpp:117-257).  lin_nodes [n, >=dim], cells [nCells, dim+1] (0-based vertex ids),
    existing: {k: lower-dimensional entities already present in the input file}.  Returns (nodes [N, dim], cells [nCells, nN])."""
    lin_nodes = np.asarray(lin_nodes, dtype=np.float64)[:, :dim]
    cells = np.asarray(cells, dtype=np.int64)
    conn, adj = _sub_entities(dim, cells, existing or {})
    conn[dim] = cells
    adj[dim] = np.arange(cells.shape[0], dtype=np.int64)[:, None]
    refs, inner = {}, {}
    for k in range(1, dim + 1):
        re = ReferenceElement(k, order, "simplex")
        refs[k] = np.asarray(re.nodes, dtype=np.float64).reshape(-1, k)
        fn = set(np.asarray(re.faceNodes).ravel().tolist())
        inner[k] = [i for i in range(refs[k].shape[0]) if i not in fn]      # ReferenceElement::getInnerNodes
    nN = refs[dim].shape[0]
    ho_nodes, vert_id = [], -np.ones(lin_nodes.shape[0], dtype=np.int64)
    ent_nodes = {k: {} for k in range(1, dim + 1)}
    ho_cells = -np.ones((cells.shape[0], nN), dtype=np.int64)
    for e, cell in enumerate(cells.tolist()):
        for i, v in enumerate(cell):
            if vert_id[v] < 0:
                vert_id[v] = len(ho_nodes)
                ho_nodes.append(lin_nodes[v])
            ho_cells[e, i] = vert_id[v]
        el_nodes = _cell_nodes(refs[dim], lin_nodes[cell])
        for k in range(1, dim + 1):
            if not inner[k]:
                continue
            for cid in adj[k][e].tolist():
                got = ent_nodes[k].get(cid)
                if got is None:
                    pts = _cell_nodes(refs[k], lin_nodes[conn[k][cid]])[inner[k]]
                    got = ent_nodes[k][cid] = list(range(len(ho_nodes), len(ho_nodes) + len(inner[k])))
                    ho_nodes.extend(pts)
                for nid in got:
                    hit = np.flatnonzero((np.abs(el_nodes - ho_nodes[nid]) < 1e-8).all(axis=1))
                    if hit.size == 0:
                        raise RuntimeError("convertGmsh2H5HO : generateHigherOrderMesh : one of the cell nodes could not be found in element")
                    ho_cells[e, hit[0]] = nid
    return np.array(ho_nodes).reshape(-1, dim), ho_cells.astype(np.int32)


def high_order_from_msh(path, dim, order):
    nodes, elems = read_msh(path)
    return high_order_from_linear(dim, order, nodes, elems[dim], {k: v for k, v in elems.items() if k < dim})
