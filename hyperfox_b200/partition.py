"""Element partition for the multi-GPU runs (one process per GPU).

The reference partitions cells with Zoltan GRAPH/PHG (src/parallel/ZoltanPartitioner.cpp:14-32,42-56); Zoltan is not available here
and its output is not pinned by any reference test, so the harness uses a deterministic stand-in: contiguous element ranges of the
lexicographically numbered Kuhn mesh (= coordinate slabs; `box_partition_vector` gives 2/4/8 coordinate boxes), recursive coordinate
bisection of the cell centroids for any mesh (`rcb_partition_vector`), or any externally supplied partition vector (`load_partition_vector`).  Global ids are assigned
before partitioning, as in the reference (Partitioner.cpp:13-36), so assembled entries do not depend on the partition.
Assembly needs no exchange; the faces shared between ranks are what a trace halo exchange would carry.
"""
import numpy as np


def slab_range(n_total, rank, world):
    return n_total * rank // world, n_total * (rank + 1) // world


def partition_vector(n_total, world):
    p = np.zeros(n_total, dtype=np.int32)
    for r in range(world):
        e0, e1 = slab_range(n_total, r, world)
        p[e0:e1] = r
    return p


def rcb_partition_vector(verts, lin_cells, world):
    """Recursive coordinate bisection of the cell centroids (the deterministic stand-in for the Zoltan partition on
    ANY mesh, structured or not).  A set of cells that has to feed k ranks is cut perpendicular to the longest axis of its bounding box
    into floor(k/2) : ceil(k/2) parts by cell count (ties broken by cell id), so every world size is balanced to within one cell."""
    verts = np.asarray(verts, dtype=np.float64)
    cen = verts[np.asarray(lin_cells)].mean(axis=1)
    part = np.zeros(cen.shape[0], dtype=np.int32)
    stack = [(np.arange(cen.shape[0]), 0, int(world))]
    while stack:
        ids, r0, k = stack.pop()
        if k == 1:
            part[ids] = r0
            continue
        c = cen[ids]
        ax = int(np.argmax(c.max(axis=0) - c.min(axis=0)))
        kl = k // 2
        nl = (ids.size * kl + k // 2) // k          # nearest integer to ids.size * kl / k
        order = np.lexsort((ids, c[:, ax]))
        stack.append((np.sort(ids[order[:nl]]), r0, kl))
        stack.append((np.sort(ids[order[nl:]]), r0 + kl, k - kl))
    return part


def load_partition_vector(path, n_total, world):
    """An externally supplied cell partition (one rank id per cell, global cell order): what the reference gets from Zoltan
    (ZoltanPartitioner.cpp:35-167) and what SURVEY.md section 8e asks the harness to accept.  `.npy`, or text with one integer per cell."""
    p = np.load(path) if str(path).endswith(".npy") else np.loadtxt(path, dtype=np.int64, ndmin=1)
    p = np.asarray(p).reshape(-1)
    if p.size != n_total:
        raise ValueError("Partitioner : computePartition : the partition file holds %d entries for %d cells" % (p.size, n_total))
    if p.min() < 0 or p.max() >= world:
        raise ValueError("Partitioner : computePartition : rank ids must lie in [0, %d)" % world)
    if np.unique(p).size != world:
        raise ValueError("Partitioner : computePartition : every rank must own at least one cell")
    return p.astype(np.int32)


def box_partition_vector(N, world, dim=3):
    """Coordinate boxes of the N^dim Kuhn mesh (cell id = cube id * dim! + k, cube id lexicographic with x fastest): the stand-in for a
    graph partition with a small cut (recursive coordinate bisection into 2/4/8 boxes).  world = 2^a is split
    along z, then y, then x, one bisection per axis in turn; any other world size falls back to slabs."""
    import math
    a = int(round(math.log2(world))) if world > 0 else 0
    fact = {2: 2, 3: 6}[dim]
    if world < 1 or (1 << a) != world:
        return partition_vector(N ** dim * fact, world)
    p = [1] * dim                                  # boxes per axis (x, y[, z])
    for k in range(a):
        p[dim - 1 - (k % dim)] *= 2
    ci = np.arange(N)
    if dim == 3:
        kk, jj, ii = np.meshgrid(ci, ci, ci, indexing="ij")
        b = ((kk * p[2] // N) * p[1] + (jj * p[1] // N)) * p[0] + (ii * p[0] // N)
    else:
        jj, ii = np.meshgrid(ci, ci, indexing="ij")
        b = (jj * p[1] // N) * p[0] + (ii * p[0] // N)
    return np.repeat(b.ravel().astype(np.int32), fact)


def extract_submesh(verts, lin_cells, owned):
    """Local vertex numbering of the owned cells. Returns (local verts, local cells, global vertex ids)."""
    cells = lin_cells[owned]
    used = np.unique(cells)
    remap = -np.ones(verts.shape[0], dtype=np.int64)
    remap[used] = np.arange(used.size)
    return verts[used], remap[cells].astype(np.int32), used


def face_keys(lin_cells, dim):
    """Sorted global-vertex tuples of the faces of each (linear) simplex, local face order of the reference element
    (tet: {3,1,0},{2,1,3},{2,3,0},{0,1,2}; tri: {0,1},{1,2},{2,0})."""
    fv = {3: [[3, 1, 0], [2, 1, 3], [2, 3, 0], [0, 1, 2]], 2: [[0, 1], [1, 2], [2, 0]]}[dim]
    k = np.sort(lin_cells[:, fv], axis=2)            # [nC, nFc, dim]
    return k


def shared_faces(lin_cells, part, rank, dim):
    """Faces of rank's cells whose other cell belongs to another rank: array of sorted global vertex tuples + the other rank
    (the content of the reference's sharedFaceList, Partitioner.h:223)."""
    keys = face_keys(lin_cells, dim)
    nC, nFc, nv = keys.shape
    flat = keys.reshape(-1, nv)
    owner = np.repeat(part, nFc)
    uniq, inv, counts = np.unique(flat, axis=0, return_inverse=True, return_counts=True)
    inv = inv.reshape(-1)
    # for interior faces (count 2) find the pair of owners
    order = np.argsort(inv, kind="stable")
    o_sorted = owner[order]
    starts = np.r_[0, np.cumsum(counts)[:-1]]
    two = counts == 2
    a, b = o_sorted[starts[two]], o_sorted[starts[two] + 1]
    cut = a != b
    mine = cut & ((a == rank) | (b == rank))
    other = np.where(a[mine] == rank, b[mine], a[mine])
    return uniq[two][mine], other


# ---- distributed trace solve: ownership + halo plan (the reference's ZoltanPartitioner.cpp:83-133 + Partitioner.cpp:42-107) --------
def global_linear_topology(lin_cells, dim):
    """Global face numbering of the linear skeleton (MOAB convention restated in csrc/host/hfx_topology.cpp): c2f [nC,nFc], f2c [nF,2]."""
    from . import capi
    tp = capi.host_compute_faces(dim, 1, lin_cells)
    return tp["cell2face"], tp["face2cell"]


def face_owner(f2c, part):
    """A face travels with its first adjacent cell (the lowest global cell id): every face is owned by exactly one rank."""
    return part[f2c[:, 0]]


def rank_cells(c2f, f2c, part, rank):
    """Owned cells of `rank` + the ghost cells across the faces it owns (overlap-1: the owner recomputes the neighbour's element)."""
    owned = np.flatnonzero(part == rank)
    own_face = face_owner(f2c, part) == rank
    second = f2c[own_face, 1]
    second = second[second >= 0]
    ghosts = np.unique(second[part[second] != rank])
    return owned, ghosts


def rank_problem(verts, lin_cells, part, rank, dim, c2f=None, f2c=None):
    """Everything rank `rank` needs, computed from the global linear mesh without communication (it is deterministic, so every rank
    derives the same plan): local cells (owned first, then ghosts), the global id and owner of every local face, and the halo lists in
    LOCAL face ids ordered by global face id.  Returns a dict."""
    if c2f is None:
        c2f, f2c = global_linear_topology(lin_cells, dim)
    owner = face_owner(f2c, part)
    world = int(part.max()) + 1
    owned, ghosts = rank_cells(c2f, f2c, part, rank)
    cells_g = np.concatenate([owned, ghosts])
    lverts, lcells, vids = extract_submesh(verts, lin_cells, cells_g)
    from . import capi
    ltp = capi.host_compute_faces(dim, 1, lcells)                      # local face numbering (same local face order as the global one)
    nFl = ltp["faces"].shape[0]
    gface = np.full(nFl, -1, dtype=np.int64)
    gface[ltp["cell2face"].ravel()] = c2f[cells_g].ravel()
    fowner = owner[gface]
    mine = fowner == rank
    # faces each rank holds as ghosts (its non-owned local faces), by global id
    need = {}
    for r in range(world):
        if r == rank:
            gf = gface
        else:
            o_r, g_r = rank_cells(c2f, f2c, part, r)
            gf = np.unique(c2f[np.concatenate([o_r, g_r])].ravel())
        need[r] = np.sort(gf[owner[gf] != r])
    order = np.argsort(gface)
    sorted_g = gface[order]

    def to_local(gids):
        pos = np.searchsorted(sorted_g, gids)
        assert np.array_equal(sorted_g[pos], gids)
        return order[pos].astype(np.int32)

    nbrs, send, recv = [], [], []
    for r in range(world):
        if r == rank:
            continue
        s = need[r][owner[need[r]] == rank]          # my owned faces that r sees as ghosts
        rr = need[rank][owner[need[rank]] == r]      # my ghosts owned by r
        if s.size or rr.size:
            nbrs.append(r); send.append(to_local(s)); recv.append(to_local(rr))
    return dict(owned_cells=owned, ghost_cells=ghosts, cells_global=cells_g, verts=lverts, lin_cells=lcells, vertex_ids=vids,
                face_global=gface, face_owner=fowner.astype(np.int32), owned_face=mine.astype(np.uint8), nbrs=np.array(nbrs, dtype=np.int32),
                send=send, recv=recv, local_topology=ltp)


def face_canonical_positions(dim, order, faces, node_vertex_gid):
    """canon[F][a] = position of local face node a of face F in the rank-independent order of that face: the face-element node order
    obtained when the face's vertices are taken in ascending global vertex id.  faces: [nF, nNf] local high-order face connectivity
    (vertices first, in face-element vertex order); node_vertex_gid[node] = global vertex id of a vertex node."""
    import itertools
    from . import capi
    nF, nNf = faces.shape
    nv = dim                                                     # vertices of a face
    if order == 1:
        ranks = np.argsort(np.argsort(node_vertex_gid[faces[:, :nv]], axis=1), axis=1)
        return ranks.astype(np.uint8)
    ref = capi.host_refel_tables(dim - 1, order)["nodes"]        # [nNf, dim-1] on [-1,1]^(dim-1)
    lam = np.concatenate([(1.0 - 0.5 * (ref + 1.0).sum(1))[:, None], 0.5 * (ref + 1.0)], axis=1)   # barycentric wrt the face-element vertices
    table = {}
    for rho in itertools.permutations(range(nv)):                # rho[k] = rank of local vertex k in the sorted order
        lam_c = np.zeros_like(lam)
        for k in range(nv):
            lam_c[:, rho[k]] = lam[:, k]
        d = np.abs(lam_c[:, None, :] - lam[None, :, :]).max(axis=2)
        pos = d.argmin(axis=1)
        assert d[np.arange(nNf), pos].max() < 1e-10 and np.unique(pos).size == nNf
        table[rho] = pos
    ranks = np.argsort(np.argsort(node_vertex_gid[faces[:, :nv]], axis=1), axis=1)
    out = np.zeros((nF, nNf), dtype=np.uint8)
    for rho, pos in table.items():
        sel = np.all(ranks == np.array(rho)[None, :], axis=1)
        out[sel] = pos[None, :]
    return out


def set_halo(ctx_handle, prob, canon):
    """Hand the halo plan of rank_problem() and the canonical face-node positions to the library (hfx_comm_set_halo)."""
    from . import capi
    from .capi import check, lib, pi
    import ctypes as C
    sc = np.array([a.size for a in prob["send"]], dtype=np.int32); rc = np.array([a.size for a in prob["recv"]], dtype=np.int32)
    sf = np.concatenate(prob["send"]).astype(np.int32) if prob["send"] else np.zeros(0, dtype=np.int32)
    rf = np.concatenate(prob["recv"]).astype(np.int32) if prob["recv"] else np.zeros(0, dtype=np.int32)
    own = np.ascontiguousarray(prob["owned_face"], dtype=np.uint8)
    check(lib().hfx_comm_set_halo(ctx_handle, int(prob["nbrs"].size), pi(prob["nbrs"]), pi(sc), pi(np.ascontiguousarray(sf)), pi(rc), pi(np.ascontiguousarray(rf)),
                                  own.ctypes.data_as(C.POINTER(C.c_ubyte)), np.ascontiguousarray(canon, dtype=np.uint8).ctypes.data_as(C.POINTER(C.c_ubyte))), ctx_handle)


# ---- the same plan through the C ABI (host C++: csrc/host/hfx_partition.cpp) -- what bench.py, dist.py and the C++ mirror use -----------------
def rcb_partition_vector_c(verts, lin_cells, world, geom=0):
    """hfx_host_rcb_partition: recursive coordinate bisection in host C++ (same cuts as rcb_partition_vector)."""
    import ctypes as C
    from .capi import lib, pd, pi, ErrorHandle
    verts = np.ascontiguousarray(verts, dtype=np.float64); cells = np.ascontiguousarray(lin_cells, dtype=np.int32)
    part = np.zeros(cells.shape[0], dtype=np.int32)
    L = lib()
    L.hfx_plan_last_error.restype = C.c_char_p
    if L.hfx_host_rcb_partition(verts.shape[1], geom, C.c_longlong(verts.shape[0]), pd(verts), C.c_longlong(cells.shape[0]), pi(cells), world, pi(part)):
        raise ErrorHandle(L.hfx_plan_last_error().decode())
    return part


def graph_partition_vector_c(lin_cells, world, dim=3, geom=0):
    """hfx_host_graph_partition: recursive bisection of the dual graph by greedy graph growing (host C++; the stand-in closest to Zoltan GRAPH)."""
    import ctypes as C
    from .capi import lib, pi, ErrorHandle
    cells = np.ascontiguousarray(lin_cells, dtype=np.int32)
    part = np.zeros(cells.shape[0], dtype=np.int32)
    L = lib()
    L.hfx_plan_last_error.restype = C.c_char_p
    if L.hfx_host_graph_partition(dim, geom, C.c_longlong(cells.shape[0]), pi(cells), world, pi(part)):
        raise ErrorHandle(L.hfx_plan_last_error().decode())
    return part


class Plan:
    """hfx_plan handle: the partition / halo plan of one rank, built in host C++ (hfx_plan_create)."""

    def __init__(self, dim, lin_cells, part, rank, world, geom=0):
        import ctypes as C
        from .capi import lib, pi, ErrorHandle
        self.L = lib()
        self.L.hfx_plan_last_error.restype = C.c_char_p
        self.L.hfx_plan_destroy.restype = None
        cells = np.ascontiguousarray(lin_cells, dtype=np.int32); part = np.ascontiguousarray(part, dtype=np.int32)
        self.h = C.c_void_p()
        if self.L.hfx_plan_create(dim, geom, C.c_longlong(cells.shape[0]), pi(cells), pi(part), rank, world, C.byref(self.h)):
            raise ErrorHandle(self.L.hfx_plan_last_error().decode())
        sz = (C.c_longlong * 8)()
        self.L.hfx_plan_sizes(self.h, sz)
        nO, nG, nV, nF, nN, nS, nR, nSh = [int(x) for x in sz]
        nv = cells.shape[1]
        lp = C.POINTER(C.c_longlong)
        self.cells_global = np.zeros(nO + nG, dtype=np.int64); self.vertex_ids = np.zeros(nV, dtype=np.int64); self.lin_cells = np.zeros((nO + nG, nv), dtype=np.int32)
        self.face_global = np.zeros(nF, dtype=np.int64); self.face_owner = np.zeros(nF, dtype=np.int32); self.owned_face = np.zeros(nF, dtype=np.uint8)
        self.nbrs = np.zeros(nN, dtype=np.int32); sc = np.zeros(nN, dtype=np.int32); rc = np.zeros(nN, dtype=np.int32)
        sf = np.zeros(nS, dtype=np.int32); rf = np.zeros(nR, dtype=np.int32); self.shared_face_list = np.zeros((nSh, 3), dtype=np.int64)
        self.L.hfx_plan_get(self.h, self.cells_global.ctypes.data_as(lp), self.vertex_ids.ctypes.data_as(lp), pi(self.lin_cells), self.face_global.ctypes.data_as(lp),
                            pi(self.face_owner), self.owned_face.ctypes.data_as(C.POINTER(C.c_ubyte)), pi(self.nbrs), pi(sc), pi(rc), pi(sf), pi(rf),
                            self.shared_face_list.ctypes.data_as(lp))
        self.n_owned, self.n_ghost = nO, nG
        self.owned_cells, self.ghost_cells = self.cells_global[:nO], self.cells_global[nO:]
        so, ro = np.r_[0, np.cumsum(sc)], np.r_[0, np.cumsum(rc)]
        self.send = [sf[so[k]:so[k + 1]] for k in range(nN)]; self.recv = [rf[ro[k]:ro[k + 1]] for k in range(nN)]

    def as_problem(self, verts):
        """The dict rank_problem() returns (same keys), from the C++ plan."""
        return dict(owned_cells=self.owned_cells, ghost_cells=self.ghost_cells, cells_global=self.cells_global, verts=np.asarray(verts)[self.vertex_ids],
                    lin_cells=self.lin_cells, vertex_ids=self.vertex_ids, face_global=self.face_global, face_owner=self.face_owner, owned_face=self.owned_face,
                    nbrs=self.nbrs, send=self.send, recv=self.recv, plan=self)

    def set_halo(self, ctx_handle, canon):
        import ctypes as C
        from .capi import check
        check(self.L.hfx_comm_set_halo_plan(ctx_handle, self.h, np.ascontiguousarray(canon, dtype=np.uint8).ctypes.data_as(C.POINTER(C.c_ubyte))), ctx_handle)

    def __del__(self):
        try:
            if self.h:
                self.L.hfx_plan_destroy(self.h); self.h = None
        except Exception:
            pass


def face_canonical_positions_c(dim, order, faces, node_vertex_gid, geom=0):
    """hfx_host_face_canonical_positions[_geom] (host C++): same result as face_canonical_positions; geom = 1: faces of orthotope cells."""
    import ctypes as C
    from .capi import lib, pi, ErrorHandle
    faces = np.ascontiguousarray(faces, dtype=np.int32); gv = np.ascontiguousarray(node_vertex_gid, dtype=np.int64)
    out = np.zeros(faces.shape, dtype=np.uint8)
    L = lib()
    L.hfx_plan_last_error.restype = C.c_char_p
    if L.hfx_host_face_canonical_positions_geom(dim, order, geom, C.c_longlong(faces.shape[0]), faces.shape[1], pi(faces), gv.ctypes.data_as(C.POINTER(C.c_longlong)),
                                                out.ctypes.data_as(C.POINTER(C.c_ubyte))):
        raise ErrorHandle(L.hfx_plan_last_error().decode())
    return out
