"""Gmsh reader + straight-sided order-p mesh generation (product: hyperfox_b200.meshio = host C++ behind the C ABI; oracle:
oracle.meshio = numpy restatement) against the reference's own
fixtures: every ressources/meshes/regression/regression_dim-D_h-H_ord-P.h5 was produced by the reference's tools/convertGmsh2H5HO from
regression_dim-D_h-H.msh (ressources/meshes/regression/generateH5FromMsh.py).  Regenerating them pins

  * the generator (cells bit-exact, node coordinates to rounding), and with it
  * the MOAB numbering convention the topology builders restate (first appearance over ascending cell ids, MBCN canonical sub-entity
    order, entities of the input file first): the node numbering of an order >= 2 (edges) / >= 3 (faces) mesh depends on the relative
    ids of the edges / faces of every cell, so a different convention gives different cells (checked below).

Fixtures: tests/golden/meshes/*.npz (converted .h5) and tests/golden/meshes/msh/*.msh.gz (the reference's files, gzip-compressed), both by
tools/make_golden_meshes.py."""
import os

import numpy as np
import pytest

from hyperfox_b200 import meshio as product
from oracle import meshio
from tests.conftest import load_mesh, unpacked_fixtures

MSH = unpacked_fixtures("msh")
CASES = [(2, "3e-1", o) for o in range(1, 6)] + [(2, "2e-1", o) for o in range(1, 6)] + [(2, "1e-1", o) for o in range(1, 5)] + \
        [(3, "3e-1", o) for o in range(1, 6)] + [(3, "2e-1", o) for o in range(1, 4)]


@pytest.mark.parametrize("impl", ["product", "oracle"])
@pytest.mark.parametrize("dim,h,order", CASES)
def test_regenerates_reference_fixture(dim, h, order, impl):
    nodes, cells = load_mesh("regression_dim-%d_h-%s_ord-%d" % (dim, h, order))
    n, c = (product if impl == "product" else meshio).high_order_from_msh(os.path.join(MSH, "regression_dim-%d_h-%s.msh" % (dim, h)), dim, order)
    assert c.shape == cells.shape and np.array_equal(c, cells)
    assert n.shape == nodes.shape and np.abs(n - nodes).max() < 1e-15


@pytest.mark.parametrize("impl", ["product", "oracle"])
def test_read_msh_counts(impl):
    meshio = product if impl == "product" else globals()["meshio"]
    nodes, el = meshio.read_msh(os.path.join(MSH, "regression_dim-3_h-3e-1.msh"))
    assert el[3].shape == (340, 4) and el[2].shape == (240, 3) and 1 not in el      # tets + the boundary triangles of the file
    assert nodes.shape[1] == 3 and el[3].min() == 0 and el[3].max() == nodes.shape[0] - 1
    nodes, el = meshio.read_msh(os.path.join(MSH, "regression_dim-2_h-3e-1.msh"))
    assert el[2].shape == (52, 3) and nodes.shape == (35, 3) and np.all(nodes[:, 2] == 0.0)


def test_fixture_pins_the_numbering_convention(monkeypatch):
    """Any other order of the tet faces / edges inside a cell, or another starting vertex of a face, fails to reproduce the fixture."""
    path = os.path.join(MSH, "regression_dim-3_h-3e-1.msh")
    g4 = load_mesh("regression_dim-3_h-3e-1_ord-4")[1]
    g2 = load_mesh("regression_dim-3_h-3e-1_ord-2")[1]
    faces, edges = list(meshio._TET_FACES), list(meshio._EDGES[3])
    for perm in ([1, 0, 2, 3], [0, 1, 3, 2], [3, 2, 1, 0], [0, 2, 1, 3]):
        monkeypatch.setattr(meshio, "_TET_FACES", [faces[i] for i in perm])
        assert not np.array_equal(meshio.high_order_from_msh(path, 3, 4)[1], g4)
    monkeypatch.setattr(meshio, "_TET_FACES", [f[1:] + f[:1] for f in faces])
    assert not np.array_equal(meshio.high_order_from_msh(path, 3, 4)[1], g4)
    monkeypatch.setattr(meshio, "_TET_FACES", faces)
    for perm in ([1, 0, 2, 3, 4, 5], [0, 1, 2, 4, 3, 5], [5, 4, 3, 2, 1, 0]):
        monkeypatch.setitem(meshio._EDGES, 3, [edges[i] for i in perm])
        assert not np.array_equal(meshio.high_order_from_msh(path, 3, 2)[1], g2)
    monkeypatch.setitem(meshio._EDGES, 3, edges)
    assert np.array_equal(meshio.high_order_from_msh(path, 3, 4)[1], g4)


def test_topology_builder_uses_the_pinned_convention():
    """The face numbering of the product's topology builder (hfx_host_compute_faces) is the same walk: faces by first appearance over
    ascending cells in the canonical face order.  Cross-check: number the faces of the linear tets with meshio's sub-entity walk (no
    pre-existing entities, as in Mesh::computeFaces where the mesh comes from the .h5 file) and compare cell2face."""
    from hyperfox_b200 import capi
    nodes, cells = load_mesh("regression_dim-3_h-3e-1_ord-1")
    conn, adj = meshio._sub_entities(3, cells.astype(np.int64), {})
    tp = capi.host_compute_faces(3, 1, cells)
    assert tp["faces"].shape[0] == conn[2].shape[0]
    assert np.array_equal(np.sort(np.asarray(tp["cell2face"]).reshape(-1, 4), axis=1), adj[2])
    assert np.array_equal(np.sort(np.asarray(tp["faces"]).reshape(-1, 3), axis=1), np.sort(conn[2], axis=1))


def test_product_reader_errors():
    from hyperfox_b200.capi import ErrorHandle
    with pytest.raises(ErrorHandle, match="MeshIo : readMsh : could not load mesh file"):
        product.read_msh(os.path.join(MSH, "does_not_exist.msh"))
    nodes, el = product.read_msh(os.path.join(MSH, "regression_dim-2_h-3e-1.msh"))
    bad = el[2].copy()
    bad[0, 0] = nodes.shape[0] + 5
    with pytest.raises(ErrorHandle, match="generateHigherOrderMesh"):
        product.high_order_from_linear(2, 2, nodes, bad)


def test_product_and_oracle_agree_on_a_kuhn_mesh():
    """A mesh without pre-existing lower-dimensional entities (the synthetic meshes of the benchmarks)."""
    from hyperfox_b200 import meshgen
    for dim, order in ((2, 4), (3, 3)):
        v, c = meshgen.kuhn_linear(2, dim)
        a, b = product.high_order_from_linear(dim, order, v, c), meshio.high_order_from_linear(dim, order, v, c)
        assert np.array_equal(a[1], b[1]) and np.abs(a[0] - b[0]).max() < 1e-15


@pytest.mark.parametrize("impl", ["product", "oracle"])
def test_reader_rejects_what_the_generator_cannot_raise(tmp_path, impl):
    """Binary files, other format versions and non-simplex elements are refused with a message instead of being misread."""
    m = product if impl == "product" else meshio
    err = Exception if impl == "product" else ValueError
    quad = "$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n4\n1 0 0 0\n2 1 0 0\n3 1 1 0\n4 0 1 0\n$EndNodes\n$Elements\n1\n1 3 2 1 1 1 2 3 4\n$EndElements\n"
    (tmp_path / "quad.msh").write_text(quad)
    with pytest.raises(err, match="not supported"):
        m.read_msh(str(tmp_path / "quad.msh"))
    (tmp_path / "bin.msh").write_text(quad.replace("2.2 0 8", "2.2 1 8"))
    with pytest.raises(err, match="ASCII"):
        m.read_msh(str(tmp_path / "bin.msh"))
    (tmp_path / "v4.msh").write_text(quad.replace("2.2 0 8", "4.1 0 8"))
    with pytest.raises(err, match="2.x"):
        m.read_msh(str(tmp_path / "v4.msh"))
    # node tags need not be contiguous or sorted: they are compacted in ascending tag order
    tri = "$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n3\n30 0 1 0\n10 0 0 0\n20 1 0 0\n$EndNodes\n$Elements\n2\n1 15 2 0 1 10\n2 2 2 0 1 10 20 30\n$EndElements\n"
    (tmp_path / "tri.msh").write_text(tri)
    nodes, el = m.read_msh(str(tmp_path / "tri.msh"))
    assert np.array_equal(nodes[:, :2], [[0, 0], [1, 0], [0, 1]]) and np.array_equal(el[2], [[0, 1, 2]]) and list(el) == [2]


H5 = unpacked_fixtures("h5")


@pytest.mark.parametrize("name", ["lightTri2", "regression_dim-2_h-2e-1_ord-2", "regression_dim-3_h-2e-1_ord-3", "regression_dim-3_h-3e-1_ord-5"])
def test_h5_mesh_reader_on_reference_files(name):
    """hfx_host_read_h5_mesh (HDF5Io::loadMesh without libhdf5) on verbatim copies of the reference's mesh files.  The content is known
    independently of any HDF5 reader: the regression files are regenerated bit-exactly from their .msh sources above, and lightTri2 is
    written out in tests/unittests/solver/TestHDGSolver.cpp / SURVEY.md section 8c (cells [[0,1,3,5,4,8],[3,1,2,4,6,7]], 9 nodes)."""
    nodes, cells = product.read_h5_mesh(os.path.join(H5, name + ".h5"))
    gn, gc = load_mesh(name)
    assert nodes.shape == gn.shape and cells.shape == gc.shape
    assert np.array_equal(nodes, gn) and np.array_equal(cells, gc)
    if name == "lightTri2":
        assert cells.tolist() == [[0, 1, 3, 5, 4, 8], [3, 1, 2, 4, 6, 7]] and nodes.shape == (9, 2)
    elif "h-3e-1" in name or "h-2e-1" in name:
        dim, order = int(name.split("dim-")[1][0]), int(name[-1])
        h = name.split("_h-")[1].split("_")[0]
        n, c = product.high_order_from_msh(os.path.join(MSH, "regression_dim-%d_h-%s.msh" % (dim, h)), dim, order)
        assert np.array_equal(c, cells) and np.abs(n - nodes).max() < 1e-15


def test_h5_mesh_reader_errors(tmp_path):
    from hyperfox_b200.capi import ErrorHandle
    with pytest.raises(ErrorHandle, match="HDF5Io : loadMesh : could not open"):
        product.read_h5_mesh(str(tmp_path / "missing.h5"))
    (tmp_path / "not.h5").write_bytes(b"\x89HDX" + bytes(200))
    with pytest.raises(ErrorHandle, match="not an HDF5 file"):
        product.read_h5_mesh(str(tmp_path / "not.h5"))
    raw = open(os.path.join(H5, "lightTri2.h5"), "rb").read()
    (tmp_path / "cut.h5").write_bytes(raw[:1500])
    with pytest.raises(ErrorHandle, match="HDF5Io : loadMesh"):
        product.read_h5_mesh(str(tmp_path / "cut.h5"))


def test_h5_mesh_reader_survives_corrupt_offsets(tmp_path):
    """Mesh files are untrusted input: every 8-byte word of a reference file replaced in turn by values that wrap around a 64-bit sum
    (2^64-8, 2^64-1 = HADDR_UNDEF, 2^63) or point far outside the file.  The reader must either raise its ErrorHandle or return the
    untouched mesh -- never read outside the buffer (`off + n > size` overflowed)."""
    from hyperfox_b200.capi import ErrorHandle
    raw = bytearray(open(os.path.join(H5, "lightTri2.h5"), "rb").read())
    gn, gc = load_mesh("lightTri2")
    path = str(tmp_path / "fuzz.h5")
    rejected = 0
    for off in range(8, min(len(raw) - 8, 4096), 8):
        for val in (2 ** 64 - 8, 2 ** 64 - 1, 2 ** 63, len(raw) - 4, 2 ** 40):
            b = bytearray(raw)
            b[off:off + 8] = int(val).to_bytes(8, "little")
            open(path, "wb").write(b)
            try:
                nodes, cells = product.read_h5_mesh(path)
            except ErrorHandle:
                rejected += 1
                continue
            assert nodes.shape[0] < 10 ** 6 and cells.shape[0] < 10 ** 6
            if nodes.shape == gn.shape and cells.shape == gc.shape and off < 2048:
                pass    # a word the reader does not interpret (or padding): the mesh may or may not be the original, but it came from inside the file
    assert rejected > 20


def test_python_io_mirrors():
    """hfox.HDF5Io / hfox.GmshIo (the reference's Io interface): both routes give the same Mesh, faces included."""
    from hyperfox_b200 import hfox
    a, b = hfox.Mesh(3, 3, "simplex"), hfox.Mesh(3, 3, "simplex")
    hfox.HDF5Io(a).load(os.path.join(H5, "regression_dim-3_h-2e-1_ord-3.h5"))
    io = hfox.GmshIo()
    io.setMesh(b)
    io.load(os.path.join(MSH, "regression_dim-3_h-2e-1.msh"))
    assert a.getNumberCells() == b.getNumberCells() == 729 and a.getNumberPoints() == b.getNumberPoints() == 4249
    assert np.array_equal(a.cells, b.cells) and np.array_equal(a.faces, b.faces) and np.array_equal(a.face2CellMap, b.face2CellMap)
    assert np.abs(a.nodes - b.nodes).max() < 1e-15
    with pytest.raises(hfox.ErrorHandle, match="is not an hdf5 file"):
        hfox.HDF5Io(a).load(os.path.join(MSH, "regression_dim-3_h-2e-1.msh"))
    with pytest.raises(hfox.ErrorHandle, match="mesh must be set"):
        hfox.GmshIo().load(os.path.join(MSH, "regression_dim-3_h-2e-1.msh"))
    with pytest.raises(hfox.ErrorHandle, match="connectivity does not match"):
        hfox.HDF5Io(hfox.Mesh(3, 2, "simplex")).load(os.path.join(H5, "regression_dim-3_h-2e-1_ord-3.h5"))
    with pytest.raises(hfox.ErrorHandle, match="write"):
        io.write("out.msh")


@pytest.mark.parametrize("dim,order,N", [(2, 5, 6), (3, 3, 5), (3, 4, 3)])
def test_generator_agrees_with_the_synthetic_mesh_generator(dim, order, N):
    """The order-p Kuhn meshes of the benchmarks (hyperfox_b200.meshgen, barycentric construction) and the reference-numbered generator
    place the same nodes in every cell (same element geometry to rounding, same node set); only the global numbering differs."""
    from hyperfox_b200 import meshgen
    v, c = meshgen.kuhn_linear(N, dim)
    n1, c1 = product.high_order_from_linear(dim, order, v, c)
    n2, c2 = meshgen.high_order(v, c, order)
    assert n1.shape == n2.shape and c1.shape == c2.shape
    assert np.abs(n1[c1] - n2[c2]).max() < 1e-15
    # the two numberings are related by one global permutation
    perm = np.full(n1.shape[0], -1, dtype=np.int64)
    perm[c1.ravel()] = c2.ravel()
    assert perm.min() >= 0 and np.unique(perm).size == perm.size and np.array_equal(perm[c1], c2)


@pytest.mark.parametrize("name", ["lightTri2", "regression_dim-2_h-2e-1_ord-2", "regression_dim-3_h-2e-1_ord-3", "regression_dim-3_h-3e-1_ord-5"])
def test_h5_writer_reproduces_the_reference_files_byte_for_byte(name, tmp_path):
    """hfx_host_write_h5 (HDF5Io::write without libhdf5) against the files the reference's own writer produced (tools/convertGmsh2H5HO.cpp -> HDF5Io::write ->
    libhdf5): written from the same mesh, the file is IDENTICAL to the reference's up to the two 4-byte modification times -- superblock, group B-trees, local
    heaps, symbol-table nodes, object headers and the placement of every block included."""
    ref = open(os.path.join(H5, name + ".h5"), "rb").read()
    nodes, cells = product.read_h5_mesh(os.path.join(H5, name + ".h5"))
    out = str(tmp_path / "out.h5")
    product.write_h5(out, nodes, cells, mtime=12345)
    got = open(out, "rb").read()
    assert len(got) == len(ref)
    diff = [i for i in range(len(ref)) if ref[i] != got[i]]
    # the modification-time messages of the two datasets: 4 bytes each, 148 bytes into a 272-byte object header that starts at 1832 (Nodes) / after the Nodes data (Cells)
    assert len(diff) <= 8 and all(1988 <= i < 1992 or i > 4000 for i in diff)
    mt = int.from_bytes(ref[1988:1992], "little")
    product.write_h5(out, nodes, cells, mtime=mt)
    got = open(out, "rb").read()
    rest = [i for i in range(len(ref)) if ref[i] != got[i]]
    assert all(i > 4000 for i in rest) and len(rest) <= 4        # only the second dataset's time stamp can still differ
    n2, c2 = product.read_h5_mesh(out)
    assert np.array_equal(n2, nodes) and np.array_equal(c2, cells)


def test_h5_field_reader_on_the_reference_file():
    """tests/unittests/io/TestHDF5Io.cpp "Load test field": ressources/meshes/fieldTest.h5 holds NodeField = 0..8 (Node, 1 x 1) and CellField = (i, -i) (Cell, 2 x 1);
    that file was written by h5py (scalar int64 ftype attribute), the reference's own writer stores an int32 [1] attribute: both are read."""
    path = os.path.join(H5, "fieldTest.h5")
    hasMesh, names = product.h5_info(path)
    assert not hasMesh and names == ["CellField", "NodeField"]
    ft, v = product.read_h5_field(path, "NodeField")
    assert ft == product.H5_NODE and v.shape == (9, 1, 1) and np.array_equal(v.ravel(), np.arange(9.0))
    ft, v = product.read_h5_field(path, "CellField")
    assert ft == product.H5_CELL and v.shape == (8, 2, 1) and np.array_equal(v.ravel(), np.array([[i, -i] for i in range(8)], dtype=float).ravel())
    err = Exception
    with pytest.raises(err, match="was not found in FieldData"):
        product.read_h5_field(path, "Nope")
    with pytest.raises(err, match="no FieldData group"):
        product.read_h5_field(os.path.join(H5, "lightTri2.h5"), "NodeField")


def test_h5_write_then_load_fields_and_mesh(tmp_path):
    """TestHDF5Io.cpp "Write test field" through the Python mirror: HDF5Io.write (Mesh + FieldData) then HDF5Io.load into fresh objects; and a file with more
    fields than one symbol-table node holds (8), names longer than the initial heap."""
    from hyperfox_b200 import hfox
    m = hfox.Mesh(2, 2, "simplex")
    hfox.HDF5Io(m).load(os.path.join(H5, "lightTri2.h5"))
    nodeField, cellField, faceField = hfox.Field(m, hfox.Node, 1, 1), hfox.Field(m, hfox.Cell, 2, 1), hfox.Field(m, hfox.Face, 3, 2)
    nodeField.values[:] = np.arange(9.0)
    cellField.values[:] = np.array([[i, -i] for i in range(2)], dtype=float).ravel()
    faceField.values[:] = np.random.default_rng(1).standard_normal(faceField.values.size)
    io = hfox.HDF5Io(m)
    io.setField("NodeField", nodeField); io.setField("CellField", cellField); io.setField("FaceField", faceField)
    out = str(tmp_path / "tmp.h5")
    io.write(out)
    m2 = hfox.Mesh(2, 2, "simplex")
    io2 = hfox.HDF5Io(m2)
    hfox.HDF5Io(m2).load(os.path.join(H5, "lightTri2.h5"))
    n2, c2, f2 = hfox.Field(m2, hfox.Node, 1, 1), hfox.Field(m2, hfox.Cell, 1, 1), hfox.Field(m2, hfox.Face, 1, 1)
    io2.setField("NodeField", n2); io2.setField("CellField", c2); io2.setField("FaceField", f2)
    io2.load(out)
    assert np.array_equal(m2.nodes, m.nodes) and np.array_equal(m2.cells, m.cells)
    assert np.array_equal(n2.values, nodeField.values) and np.array_equal(c2.values, cellField.values) and np.array_equal(f2.values, faceField.values)
    assert (c2.type, c2.nObj, c2.nVals) == (hfox.Cell, 2, 1) and (f2.type, f2.nObj, f2.nVals) == (hfox.Face, 3, 2)
    # the independent dev-time reader (tools/mini_h5.py, written against libhdf5's own files) parses the written file as well
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from mini_h5 import MiniH5
    f = MiniH5(out)
    assert np.array_equal(f.read("/Mesh/Nodes"), m.nodes) and np.array_equal(f.read("/Mesh/Cells"), m.cells)
    assert np.array_equal(f.read("/FieldData/FaceField").ravel(), faceField.values)
    # many fields, long names
    rng = np.random.default_rng(3)
    fields = {"RKStage_Flux_%d_with_a_rather_long_name" % k: (product.H5_CELL, rng.standard_normal((5, 3, 2))) for k in range(21)}
    big = str(tmp_path / "many.h5")
    product.write_h5(big, fields=fields)
    hasMesh, names = product.h5_info(big)
    assert not hasMesh and names == sorted(fields)
    for k in names:
        ft, v = product.read_h5_field(big, k)
        assert ft == product.H5_CELL and np.array_equal(v, fields[k][1])
    g = MiniH5(big)
    for k in names:
        assert np.array_equal(g.read("/FieldData/" + k), fields[k][1])
    with pytest.raises(Exception, match="could not find anything to write"):
        product.write_h5(str(tmp_path / "empty.h5"))


def test_h5_writer_rejects_bad_input(tmp_path):
    out = str(tmp_path / "bad.h5")
    with pytest.raises(Exception, match="non-empty name"):
        product.write_h5(out, fields={"a/b": (product.H5_NODE, np.zeros((2, 1, 1)))})
    with pytest.raises(Exception, match="problem writing values"):
        product.write_h5(out, fields={"f": (product.H5_NODE, np.zeros((2, 1)))})
    # an empty field is legal (zero entities): written and read back
    product.write_h5(out, fields={"empty": (product.H5_CELL, np.zeros((0, 3, 1))), "one": (product.H5_FACE, np.ones((1, 1, 1)))})
    ft, v = product.read_h5_field(out, "empty")
    assert ft == product.H5_CELL and v.shape == (0, 3, 1)
    ft, v = product.read_h5_field(out, "one")
    assert ft == product.H5_FACE and v.ravel().tolist() == [1.0]
