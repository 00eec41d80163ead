"""Element partition for the multi-GPU runs (one process per GPU).

The reference partitions cells with Zoltan GRAPH/PHG (src/parallel/ZoltanPartitioner.cpp:14-32,42-56); Zoltan is not available here
and its output is not pinned by any reference test, so the harness uses a deterministic stand-in: contiguous element ranges of the
lexicographically numbered Kuhn mesh (= coordinate slabs), or any externally supplied partition vector.  Global ids are assigned
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
