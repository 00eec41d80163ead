"""Synthetic Kuhn meshes for the oracle's own legs (bench.py --impl reference, cpu_baseline): the same construction as the product's
hyperfox_b200/meshgen.py (unit cube split into Kuhn simplices, order-p straight-sided nodes as the affine image of the reference nodes,
what tools/convertGmsh2H5HO.cpp:117-257 does for Gmsh meshes), built on the ORACLE's reference element so that the reference arm never
imports the product package.  Test infrastructure."""
import itertools

import numpy as np

from .refel import ReferenceElement


def kuhn_linear(N, dim=3):
    """(N+1)^dim lattice vertices of [0,1]^dim and the dim! Kuhn simplices of every cube; cell id = cube id * dim! + k."""
    ax = np.arange(N + 1)
    if dim == 3:
        zz, yy, xx = np.meshgrid(ax, ax, ax, indexing="ij")
        verts = np.stack([xx.ravel(), yy.ravel(), zz.ravel()], axis=1).astype(np.float64) / N
        vid = lambda i, j, k: (k * (N + 1) + j) * (N + 1) + i
        ci = np.arange(N)
        kk, jj, ii = np.meshgrid(ci, ci, ci, indexing="ij")
        ii, jj, kk = ii.ravel(), jj.ravel(), kk.ravel()
        cells = []
        for perm in itertools.permutations(range(3)):
            cur = [ii.copy(), jj.copy(), kk.copy()]
            vs = [vid(*cur)]
            for ax_ in perm:
                cur[ax_] = cur[ax_] + 1
                vs.append(vid(*cur))
            cells.append(np.stack(vs, axis=1))
        cells = np.stack(cells, axis=1).reshape(-1, 4)
    else:
        yy, xx = np.meshgrid(ax, ax, indexing="ij")
        verts = np.stack([xx.ravel(), yy.ravel()], axis=1).astype(np.float64) / N
        vid = lambda i, j: j * (N + 1) + i
        ci = np.arange(N)
        jj, ii = np.meshgrid(ci, ci, indexing="ij")
        ii, jj = ii.ravel(), jj.ravel()
        cells = np.stack([np.stack([vid(ii, jj), vid(ii + 1, jj), vid(ii + 1, jj + 1)], axis=1),
                          np.stack([vid(ii, jj), vid(ii + 1, jj + 1), vid(ii, jj + 1)], axis=1)], axis=1).reshape(-1, 3)
    # positive orientation (det > 0) so that dV > 0 like the Gmsh-generated reference meshes
    p = verts[cells]
    e = p[:, 1:] - p[:, :1]
    neg = np.linalg.det(e) < 0
    cells[neg, -2], cells[neg, -1] = cells[neg, -1].copy(), cells[neg, -2].copy()
    return verts, cells.astype(np.int32)


def high_order(verts, cells, order, perturb=0.0, seed=20240229):
    """Order-p nodes/cells from a linear simplex mesh. Returns (nodes [nNodes,dim], cells [nCells,nN])."""
    dim = verts.shape[1]
    if perturb > 0.0:
        rng = np.random.default_rng(seed)
        lo, hi = verts.min(0), verts.max(0)
        interior = np.all((verts > lo + 1e-12) & (verts < hi - 1e-12), axis=1)
        h = (hi - lo).max() / round((verts.shape[0]) ** (1.0 / dim) - 1)
        verts = verts.copy()
        verts[interior] += perturb * h * rng.uniform(-1, 1, size=(int(interior.sum()), dim))
    if order == 1:
        return verts, cells.astype(np.int32)
    ref = ReferenceElement(dim, order).nodes           # [nN, dim] on the reference simplex [-1,1]^dim
    nN = ref.shape[0]
    lam = np.concatenate([(1.0 - 0.5 * (ref + 1.0).sum(1))[:, None], 0.5 * (ref + 1.0)], axis=1)   # barycentric, [nN, dim+1]
    # classes of barycentric values (0 = exactly zero)
    vals = np.sort(lam.ravel())
    reps = []
    for v in vals:
        if abs(v) < 1e-12:
            continue
        if not reps or abs(v - reps[-1]) > 1e-10:
            reps.append(v)
    reps = np.array(reps)
    cls = np.where(np.abs(lam) < 1e-12, 0, 1 + np.argmin(np.abs(lam[..., None] - reps[None, None, :]), axis=2))   # [nN, dim+1]
    nC = cells.shape[0]
    # key of node (c, i): sorted list of (vertex id, class) over the vertices with non-zero weight
    vids = np.broadcast_to(cells[:, None, :], (nC, nN, dim + 1)).astype(np.int64)
    cl = np.broadcast_to(cls[None, :, :], (nC, nN, dim + 1)).astype(np.int64)
    pair = np.where(cl > 0, (vids + 1) * 64 + cl, 0)   # 0 = absent ; vertex ids < 2^25
    pair = -np.sort(-pair, axis=2)                     # descending, zeros last
    nbits = 32
    k0 = (pair[..., 0] << nbits) | pair[..., 1]
    k1 = (pair[..., 2] << nbits) | (pair[..., 3] if dim == 3 else 0)
    k0, k1 = k0.ravel(), k1.ravel()
    order_ = np.lexsort((k1, k0))
    s0, s1 = k0[order_], k1[order_]
    new = np.r_[True, (s0[1:] != s0[:-1]) | (s1[1:] != s1[:-1])]
    gid_sorted = np.cumsum(new) - 1
    # number nodes by first appearance so that vertex-like ordering is deterministic: first = smallest flat index
    first_flat = np.full(gid_sorted[-1] + 1, np.iinfo(np.int64).max)
    np.minimum.at(first_flat, gid_sorted, order_)
    rank = np.empty_like(first_flat)
    rank[np.argsort(first_flat, kind="stable")] = np.arange(first_flat.size)
    gid = np.empty(nC * nN, dtype=np.int64)
    gid[order_] = rank[gid_sorted]
    hcells = gid.reshape(nC, nN).astype(np.int32)
    # coordinates from the first appearance
    nodes = np.zeros((first_flat.size, dim))
    flat = first_flat[np.argsort(rank)] if False else None
    src = np.empty(first_flat.size, dtype=np.int64)
    src[rank] = first_flat
    c_idx, n_idx = src // nN, src % nN
    nodes = np.einsum("nk,nkd->nd", lam[n_idx], verts[cells[c_idx]])
    return nodes, hcells


def kuhn_mesh(N, order, dim=3, perturb=0.0):
    v, c = kuhn_linear(N, dim)
    return high_order(v, c, order, perturb)


