"""ORACLE (test infrastructure): mesh topology as the reference derives it through MOAB.

Restates /root/reference/src/mesh/Mesh.cpp:183-274 (computeFaces), :377-402 (face2Cell), :404-425 (cell2Face),
:496-515 (boundary), :517-537 (face connectivity).

MOAB (bitbucket HEAD, unpinned, absent here) decides the *global face numbering*.  Convention restated (SURVEY.md
section 8c): `get_adjacencies(cells, dim-1, create=true, UNION)` creates faces while walking the cells in
ascending id and, inside a cell, the local faces in the reference element's face order (same vertex sets as MBCN's
canonical sides); a face seen for the first time gets the next id.  face2Cell lists adjacent cells in ascending id,
second entry -1 on the boundary; a face's node list is taken from its lowest-id cell through the order-p faceNodes.
No reference test pins the numbering itself (TestMesh.cpp checks membership only), but the reference's mesh fixtures do: the
.h5 regression meshes were generated from the .msh files by tools/convertGmsh2H5HO, whose node numbering depends on the relative
MOAB ids of the faces of every cell; tests/test_meshio.py regenerates them bit-exactly with this convention (and with no other
order of the local faces) and checks that this builder and the product's number the faces by the same walk.
"""
import numpy as np


def compute_faces(cells, refel, skel_face_nodes=None):
    """cells: [nCells, nN] int. refel: oracle.refel.ReferenceElement (order p). Returns dict of int32 arrays."""
    from .refel import ReferenceElement
    cells = np.asarray(cells)
    nC = cells.shape[0]
    dim = refel.dim
    skel = ReferenceElement(dim, 1 if refel.order != 0 else 0, refel.geom)
    sfn = np.array(skel.faceNodes)                       # local face -> skeleton vertex indices
    nFc, nVf = sfn.shape
    nSk = skel.nNodes
    corner = cells[:, :nSk]                              # first nSkelNodes of each cell (Mesh.cpp:222-229)
    fv = corner[:, sfn]                                  # [nC, nFc, nVf]
    keys = np.sort(fv.reshape(nC * nFc, nVf), axis=1)
    uniq, first, inv = np.unique(keys, axis=0, return_index=True, return_inverse=True)
    inv = inv.reshape(-1)
    order = np.argsort(first, kind="stable")             # faces numbered by first appearance in (cell, local face) order
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    fid = rank[inv]                                      # global face id of every (cell, local face)
    nF = uniq.shape[0]
    cell2face = fid.reshape(nC, nFc).astype(np.int32)
    face2cell = -np.ones((nF, 2), dtype=np.int32)
    cell_of = np.repeat(np.arange(nC), nFc)
    # ascending cell id: first occurrence is the lowest id
    so = np.lexsort((cell_of, fid))
    fs, cs = fid[so], cell_of[so]
    starts = np.r_[0, np.flatnonzero(np.diff(fs)) + 1]
    counts = np.diff(np.r_[starts, fs.size])
    assert counts.max() <= 2, "non-manifold mesh"
    face2cell[fs[starts], 0] = cs[starts]
    two = counts == 2
    face2cell[fs[starts[two]], 1] = cs[starts[two] + 1]
    # face connectivity from the lowest-id cell (Mesh.cpp:529-536)
    fn = np.array(refel.faceNodes)                       # [nFc, nNf]
    c0 = face2cell[:, 0]
    lf0 = np.argmax(cell2face[c0] == np.arange(nF)[:, None], axis=1)
    faces = cells[c0[:, None], fn[lf0]].astype(np.int32)
    boundary = np.flatnonzero(face2cell[:, 1] < 0).astype(np.int32)
    return dict(faces=faces, cell2face=cell2face, face2cell=face2cell, boundary=boundary)
