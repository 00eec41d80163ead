"""Host logic of the multi-GPU path, on CPU: face ownership, overlap-1 ghost cells, halo send/recv lists and
the rank-independent face-node order in which trace blocks travel."""
import numpy as np
import pytest

from hyperfox_b200 import capi, meshgen, partition as P


@pytest.mark.parametrize("world,boxes", [(2, False), (4, False), (8, False), (2, True), (4, True), (8, True)])
def test_ownership_and_halo_lists_are_consistent(world, boxes):
    v, c = meshgen.kuhn_linear(4, 3)
    part = P.box_partition_vector(4, world, 3) if boxes else P.partition_vector(c.shape[0], world)
    assert part.shape[0] == c.shape[0] and np.array_equal(np.bincount(part, minlength=world), np.full(world, c.shape[0] // world))
    c2f, f2c = P.global_linear_topology(c, 3)
    probs = [P.rank_problem(v, c, part, r, 3, c2f, f2c) for r in range(world)]
    # every face is owned by exactly one rank (ZoltanPartitioner.cpp:83-133: the face travels with one of its cells)
    owned = np.concatenate([p["face_global"][p["owned_face"] == 1] for p in probs])
    assert np.array_equal(np.sort(owned), np.arange(f2c.shape[0]))
    for r, p in enumerate(probs):
        # rows of owned faces are complete: both adjacent cells are local
        loc_cells = set(p["cells_global"].tolist())
        for F in p["face_global"][p["owned_face"] == 1]:
            assert all(int(cc) in loc_cells for cc in f2c[F] if cc >= 0)
        # what r sends to s is exactly what s receives from r, in the same (global id) order: the sharedFaceList contract
        for k, s in enumerate(p["nbrs"]):
            q = probs[s]
            ks = list(q["nbrs"]).index(r)
            assert np.array_equal(p["face_global"][p["send"][k]], q["face_global"][q["recv"][ks]])
            assert np.all(p["owned_face"][p["send"][k]] == 1) and np.all(p["owned_face"][p["recv"][k]] == 0)
        # every ghost face is received from exactly one neighbour
        ghosts = np.flatnonzero(p["owned_face"] == 0)
        recv = np.concatenate(p["recv"]) if p["recv"] else np.zeros(0, dtype=int)
        assert np.array_equal(np.sort(recv), ghosts)


@pytest.mark.parametrize("dim,order", [(2, 2), (2, 4), (3, 1), (3, 2), (3, 3), (3, 4)])
def test_canonical_face_node_order_is_rank_independent(dim, order):
    """The two cells of a face order its nodes differently; the canonical positions must map the same physical node to the same slot."""
    v, c = meshgen.kuhn_linear(2, dim)
    nodes, cells = meshgen.high_order(v, c, order)
    tp = capi.host_compute_faces(dim, order, cells)
    gv = np.full(nodes.shape[0], -1, dtype=np.int64)
    gv[cells[:, :dim + 1]] = c
    fn = capi.host_refel_tables(dim, order)["faceNodes"]
    canon1 = P.face_canonical_positions(dim, order, tp["faces"], gv)
    checked = 0
    for F in range(tp["faces"].shape[0]):
        c2 = tp["face2cell"][F, 1]
        if c2 < 0:
            continue
        k = list(tp["cell2face"][c2]).index(F)
        f2 = cells[c2][fn[k]][None, :]
        canon2 = P.face_canonical_positions(dim, order, f2, gv)[0]
        a = np.empty(f2.shape[1], dtype=int); a[canon1[F]] = tp["faces"][F]
        b = np.empty(f2.shape[1], dtype=int); b[canon2] = f2[0]
        assert np.array_equal(a, b)
        checked += 1
    assert checked > 0


def test_supplied_partition_vector(tmp_path):
    """An arbitrary (non-contiguous, graph-partitioner-like) partition vector from a file gives a consistent plan too, and the
    assembled system does not depend on it: ownership covers every face exactly once."""
    v, c = meshgen.kuhn_linear(3, 3)
    rng = np.random.default_rng(5)
    part = rng.integers(0, 3, size=c.shape[0]).astype(np.int32)
    np.save(tmp_path / "part.npy", part)
    np.savetxt(tmp_path / "part.txt", part, fmt="%d")
    for f in ("part.npy", "part.txt"):
        assert np.array_equal(P.load_partition_vector(str(tmp_path / f), c.shape[0], 3), part)
    with pytest.raises(ValueError, match="holds"):
        P.load_partition_vector(str(tmp_path / "part.npy"), c.shape[0] + 1, 3)
    with pytest.raises(ValueError, match="rank ids"):
        P.load_partition_vector(str(tmp_path / "part.npy"), c.shape[0], 2)
    with pytest.raises(ValueError, match="at least one cell"):
        P.load_partition_vector(str(tmp_path / "part.npy"), c.shape[0], 4)      # rank 3 owns nothing
    c2f, f2c = P.global_linear_topology(c, 3)
    probs = [P.rank_problem(v, c, part, r, 3, c2f, f2c) for r in range(3)]
    owned = np.concatenate([p["face_global"][p["owned_face"] == 1] for p in probs])
    assert np.array_equal(np.sort(owned), np.arange(f2c.shape[0]))
    assert sum(int(p["owned_cells"].size) for p in probs) == c.shape[0]


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_rcb_on_an_unstructured_reference_mesh(world):
    """Recursive coordinate bisection on the reference's Gmsh tet mesh regression_dim-3_h-2e-1 (729 cells): balanced
    to within one cell per bisection level, every face owned once, halo lists symmetric, and a cut much smaller than a random partition's."""
    from tests.conftest import load_mesh
    nodes, cells = load_mesh("regression_dim-3_h-2e-1_ord-1")
    part = P.rcb_partition_vector(nodes, cells, world)
    counts = np.bincount(part, minlength=world)
    assert counts.sum() == cells.shape[0] and counts.max() - counts.min() <= 3
    assert np.array_equal(part, P.rcb_partition_vector(nodes, cells, world))           # deterministic
    c2f, f2c = P.global_linear_topology(cells, 3)
    interior = f2c[:, 1] >= 0
    cut = int((part[f2c[interior, 0]] != part[f2c[interior, 1]]).sum())
    rnd = np.random.default_rng(1).integers(0, world, size=cells.shape[0])
    cut_rnd = int((rnd[f2c[interior, 0]] != rnd[f2c[interior, 1]]).sum())
    assert cut < 0.45 * cut_rnd
    probs = [P.rank_problem(nodes[:, :3], cells, part, r, 3, c2f, f2c) for r in range(world)]
    owned = np.concatenate([p["face_global"][p["owned_face"] == 1] for p in probs])
    assert np.array_equal(np.sort(owned), np.arange(f2c.shape[0]))
    for r, p in enumerate(probs):
        for k, s in enumerate(p["nbrs"]):
            q = probs[s]
            ks = list(q["nbrs"]).index(r)
            assert np.array_equal(p["face_global"][p["send"][k]], q["face_global"][q["recv"][ks]])


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_host_cpp_plan_equals_the_numpy_plan(world):
    """The plan behind the C ABI (hfx_plan_create, host C++) against the numpy restatement above: same partition vector, same local numbering,
    same ownership, same send / receive lists -- and the reference's sharedFaceList triples [global face, other rank, global adjacent cell]."""
    v, c = meshgen.kuhn_linear(5, 3)
    v = v + 0.01 * np.sin(7.0 * v[:, ::-1])            # break the lattice symmetry: the bisection cuts must not depend on ties
    part = P.rcb_partition_vector(v, c, world)
    assert np.array_equal(P.rcb_partition_vector_c(v, c, world), part)
    c2f, f2c = P.global_linear_topology(c, 3)
    for r in range(world):
        ref = P.rank_problem(v, c, part, r, 3, c2f, f2c)
        pl = P.Plan(3, c, part, r, world)
        q = pl.as_problem(v)
        for k in ("owned_cells", "ghost_cells", "cells_global", "lin_cells", "vertex_ids", "face_global", "face_owner", "owned_face", "nbrs", "verts"):
            assert np.array_equal(np.asarray(ref[k]), np.asarray(q[k])), k
        for k in range(len(ref["nbrs"])):
            assert np.array_equal(ref["send"][k], q["send"][k]) and np.array_equal(ref["recv"][k], q["recv"][k])
        # sharedFaceList: every cut face of this rank once, with the rank and the cell on the other side
        keys, other = P.shared_faces(c, part, r, 3)
        sfl = pl.shared_face_list
        assert sfl.shape[0] == keys.shape[0]
        for F, orank, ocell in sfl:
            cc = f2c[F]
            assert part[ocell] == orank != r and ocell in cc and part[cc[0] if cc[1] == ocell else cc[1]] == r


@pytest.mark.parametrize("dim,order", [(2, 3), (3, 1), (3, 3), (3, 4)])
def test_host_cpp_canonical_positions_equal_numpy(dim, order):
    v, c = meshgen.kuhn_linear(2, dim)
    nodes, cells = meshgen.high_order(v, c, order)
    tp = capi.host_compute_faces(dim, order, cells)
    gv = np.full(nodes.shape[0], -1, dtype=np.int64)
    gv[cells[:, :dim + 1]] = np.random.default_rng(3).permutation(v.shape[0])[c]     # arbitrary global vertex ids
    assert np.array_equal(P.face_canonical_positions_c(dim, order, tp["faces"], gv), P.face_canonical_positions(dim, order, tp["faces"], gv))


def test_plan_errors_are_reported():
    v, c = meshgen.kuhn_linear(2, 3)
    with pytest.raises(Exception, match="Partitioner"):
        P.Plan(3, c, np.full(c.shape[0], 5, dtype=np.int32), 0, 2)
    with pytest.raises(Exception, match="Partitioner"):
        P.Plan(3, c, np.zeros(c.shape[0], dtype=np.int32), 1, 2)          # rank 1 owns nothing


@pytest.mark.parametrize("mesh,world", [("kuhn", 2), ("kuhn", 5), ("kuhn", 8), ("gmsh", 3), ("gmsh", 8)])
def test_graph_partition_is_balanced_connected_and_plans(mesh, world):
    """hfx_host_graph_partition (recursive bisection of the dual graph by greedy graph growing, the stand-in closest to Zoltan GRAPH): balanced to one cell,
    every part connected in the dual graph, no coordinates needed, and a valid input of the halo plan (ownership / send = receive lists)."""
    if mesh == "kuhn":
        v, c = meshgen.kuhn_linear(5, 3)
    else:
        from tests.conftest import load_mesh
        nodes, cells = load_mesh("regression_dim-3_h-2e-1_ord-1")
        v, c = nodes, cells[:, :4]
    part = P.graph_partition_vector_c(c, world)
    counts = np.bincount(part, minlength=world)
    assert counts.min() >= c.shape[0] // world - 1 and counts.max() <= -(-c.shape[0] // world) + 1
    c2f, f2c = P.global_linear_topology(c, 3)
    for r in range(world):           # compactness of each part: graph growing leaves one dominant connected piece (the tail of a BFS order may split off a few cells)
        ids = np.flatnonzero(part == r); inpart = np.zeros(c.shape[0], dtype=bool); inpart[ids] = True
        seen = np.zeros(c.shape[0], dtype=bool); comps = []
        for s0 in ids:
            if seen[s0]:
                continue
            stack = [s0]; seen[s0] = True; n = 0
            while stack:
                x = stack.pop(); n += 1
                for F in c2f[x]:
                    for y in f2c[F]:
                        if y >= 0 and inpart[y] and not seen[y]:
                            seen[y] = True; stack.append(y)
            comps.append(n)
        assert max(comps) >= 0.8 * ids.size, (r, comps)
    plans = [P.Plan(3, c, part, r, world) for r in range(world)]
    owned = np.concatenate([pl.face_global[pl.owned_face == 1] for pl in plans])
    assert np.array_equal(np.sort(owned), np.arange(f2c.shape[0]))
    for r, pl in enumerate(plans):
        for k, s2 in enumerate(pl.nbrs):
            q = plans[s2]; ks = list(q.nbrs).index(r)
            assert np.array_equal(pl.face_global[pl.send[k]], q.face_global[q.recv[ks]])
    # the cut: a graph partition of the structured mesh should not be worse than twice the coordinate bisection's
    if mesh == "kuhn":
        cut = lambda pv: int(((f2c[:, 1] >= 0) & (pv[f2c[:, 0]] != pv[np.maximum(f2c[:, 1], 0)])).sum())
        assert cut(part) <= 2 * cut(P.rcb_partition_vector(v, c, world))


@pytest.mark.parametrize("dim,order", [(2, 2), (2, 4), (3, 1), (3, 2)])
def test_canonical_face_node_order_of_orthotope_faces(dim, order):
    """Quadrilateral faces of hexahedra (and the edges of quads): the two cells of an interior face order its nodes differently (any of the 8 symmetries of the square);
    the canonical positions must put the same physical node in the same slot -- with arbitrary global vertex ids."""
    nodes, cells = meshgen.box_mesh(2, order, dim, perturb=0.1)
    tp = capi.host_compute_faces(dim, order, cells, 1)
    nv = 2 ** dim
    gv = np.full(nodes.shape[0], -1, dtype=np.int64)
    vids = np.unique(cells[:, :nv])
    gv[vids] = np.random.default_rng(5).permutation(10 * vids.size)[:vids.size]
    fn = capi.host_refel_tables(dim, order, 1)["faceNodes"]
    canon1 = P.face_canonical_positions_c(dim, order, tp["faces"], gv, geom=1)
    assert all(np.array_equal(np.sort(r), np.arange(r.size)) for r in canon1)
    checked = 0
    for F in range(tp["faces"].shape[0]):
        c2 = tp["face2cell"][F, 1]
        if c2 < 0:
            continue
        k = list(tp["cell2face"][c2]).index(F)
        f2 = np.ascontiguousarray(cells[c2][fn[k]][None, :])
        canon2 = P.face_canonical_positions_c(dim, order, f2, gv, geom=1)[0]
        a = np.empty(f2.shape[1], dtype=int); a[canon1[F]] = tp["faces"][F]
        b = np.empty(f2.shape[1], dtype=int); b[canon2] = f2[0]
        assert np.array_equal(a, b)
        checked += 1
    assert checked > 0


@pytest.mark.parametrize("world", [2, 4])
def test_plan_of_a_hexahedral_mesh(world):
    nodes, cells = meshgen.box_mesh(4, 1, 3)
    part = P.rcb_partition_vector_c(nodes, cells, world, geom=1)
    assert np.bincount(part, minlength=world).tolist() == [cells.shape[0] // world] * world
    tp = capi.host_compute_faces(3, 1, cells, 1)
    plans = [P.Plan(3, cells, part, r, world, geom=1) for r in range(world)]
    owned = np.concatenate([pl.face_global[pl.owned_face == 1] for pl in plans])
    assert np.array_equal(np.sort(owned), np.arange(tp["faces"].shape[0]))
    for r, pl in enumerate(plans):
        for k, s2 in enumerate(pl.nbrs):
            q = plans[s2]; ks = list(q.nbrs).index(r)
            assert np.array_equal(pl.face_global[pl.send[k]], q.face_global[q.recv[ks]])
        ghosts = np.flatnonzero(pl.owned_face == 0)
        recv = np.concatenate(pl.recv) if pl.recv else np.zeros(0, dtype=int)
        assert np.array_equal(np.sort(recv), ghosts)
