"""Mesh input of the path's callers: Gmsh 2.2 ASCII reader and the straight-sided order-p mesh
generator of the reference's tools/convertGmsh2H5HO.cpp:117-397, without MOAB.  The implementation is host C++
(csrc/host/hfx_meshio.cpp) behind the C ABI (hfx_host_read_msh, hfx_host_high_order_mesh, hfx_host_read_h5_mesh: the reference's .h5 mesh
files without libhdf5); this module is the ctypes wrapper.
The generated meshes carry the reference's node numbering: tests/test_meshio.py regenerates the reference's .h5 regression fixtures
from their .msh sources, cells bit-exact."""
import ctypes as C

import numpy as np

from . import capi
from .capi import ErrorHandle, check, lib, pd, pi


def read_msh(path):
    """Returns (nodes [n,3] by ascending node tag, {topological dim k: int32 [m, k+1]} in file order, 0-based vertex ids)."""
    L = lib()
    n, counts = C.c_int(0), (C.c_int * 4)()
    check(L.hfx_host_read_msh(str(path).encode(), C.byref(n), counts, None, None, None, None))
    nodes = np.zeros((n.value, 3))
    el = {k: np.zeros((counts[k], k + 1), dtype=np.int32) for k in (1, 2, 3)}
    check(L.hfx_host_read_msh(str(path).encode(), C.byref(n), counts, pd(nodes), pi(el[1]), pi(el[2]), pi(el[3])))
    return nodes, {k: v for k, v in el.items() if v.shape[0]}


def high_order_from_linear(dim, order, lin_nodes, cells, existing=None):
    """generateHigherOrderMesh: (nodes [N, dim], cells [nCells, nN]) of the order-p mesh; `existing` = {k: entities of topological
    dimension k < dim already present in the input file} (they are numbered before the generated ones, as MOAB does)."""
    L = lib()
    lin = np.ascontiguousarray(np.asarray(lin_nodes, dtype=np.float64)[:, :dim])
    cells = np.ascontiguousarray(cells, dtype=np.int32)
    ex = {k: np.ascontiguousarray((existing or {}).get(k, np.zeros((0, k + 1))), dtype=np.int32) for k in (1, 2)}
    nN = capi.host_refel_tables(dim, order)["nodes"].reshape(-1, dim).shape[0]
    n = C.c_int(0)
    args = (dim, order, lin.shape[0], pd(lin), cells.shape[0], pi(cells), ex[1].shape[0], pi(ex[1]), ex[2].shape[0] if dim == 3 else 0, pi(ex[2]))
    check(L.hfx_host_high_order_mesh(*args, C.byref(n), None, None))
    nodes, ho = np.zeros((n.value, dim)), np.zeros((cells.shape[0], nN), dtype=np.int32)
    check(L.hfx_host_high_order_mesh(*args, C.byref(n), pd(nodes), pi(ho)))
    return nodes, ho


def high_order_from_msh(path, dim, order):
    nodes, elems = read_msh(path)
    return high_order_from_linear(dim, order, nodes, elems[dim], {k: v for k, v in elems.items() if k < dim})


def read_h5_mesh(path):
    """HDF5Io::loadMesh (src/io/HDF5Io.cpp:111-152) without libhdf5: (nodes [nNodes, dimNodeSpace], cells [nCells, nN]) of a mesh file of
    the reference (ressources/meshes/**/*.h5)."""
    L = lib()
    nn, d, nc, npc = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
    check(L.hfx_host_read_h5_mesh(str(path).encode(), C.byref(nn), C.byref(d), C.byref(nc), C.byref(npc), None, None))
    nodes, cells = np.zeros((nn.value, d.value)), np.zeros((nc.value, npc.value), dtype=np.int32)
    check(L.hfx_host_read_h5_mesh(str(path).encode(), C.byref(nn), C.byref(d), C.byref(nc), C.byref(npc), pd(nodes), pi(cells)))
    return nodes, cells


# FieldType values stored in the files: the reference's enum (src/field/FieldTypes.h)
H5_NODE, H5_EDGE, H5_FACE, H5_CELL = 0, 1, 2, 3


def write_h5(path, nodes=None, cells=None, fields=None, mtime=0):
    """HDF5Io::write (src/io/HDF5Io.cpp:66-109) without libhdf5.  fields: {name: (ftype, values[nEntities, nObjPerEnt, nValsPerObj])} with the reference's
    FieldType values (H5_NODE / H5_FACE / H5_CELL)."""
    L = lib()
    fields = fields or {}
    names = sorted(fields)
    arrs = [np.ascontiguousarray(fields[n][1], dtype=np.float64) for n in names]
    for n, a in zip(names, arrs):
        if a.ndim != 3:
            raise ErrorHandle("HDF5Io : writeFields : problem writing values of field: " + n)
    cn = (C.c_char_p * max(1, len(names)))(*[n.encode() for n in names])
    ft = np.array([int(fields[n][0]) for n in names] or [0], dtype=np.int32)
    sh = np.array([a.shape for a in arrs] or [[0, 0, 0]], dtype=np.int64)
    vp = (C.c_void_p * max(1, len(names)))(*[a.ctypes.data for a in arrs])
    if nodes is not None:
        nodes = np.ascontiguousarray(nodes, dtype=np.float64); cells = np.ascontiguousarray(cells, dtype=np.int32)
        check(L.hfx_host_write_h5(str(path).encode(), C.c_uint(int(mtime)), nodes.shape[1], C.c_longlong(nodes.shape[0]), pd(nodes), C.c_longlong(cells.shape[0]), cells.shape[1],
                                  pi(cells), len(names), cn, pi(ft), sh.ctypes.data_as(C.POINTER(C.c_longlong)), vp))
    else:
        check(L.hfx_host_write_h5(str(path).encode(), C.c_uint(int(mtime)), 0, C.c_longlong(0), None, C.c_longlong(0), 0, None, len(names), cn, pi(ft),
                                  sh.ctypes.data_as(C.POINTER(C.c_longlong)), vp))


def h5_info(path):
    """(has a Mesh group, names of the datasets of the FieldData group)"""
    L = lib()
    hm, nf = C.c_int(0), C.c_int(0)
    buf = C.create_string_buffer(1 << 16)
    check(L.hfx_host_h5_info(str(path).encode(), C.byref(hm), C.byref(nf), buf, len(buf)))
    return bool(hm.value), [s for s in buf.value.decode().split("\n") if s]


def read_h5_field(path, name):
    """HDF5Io::loadFields (src/io/HDF5Io.cpp:154-187): (ftype, values [nEntities, nObjPerEnt, nValsPerObj]) of /FieldData/<name>"""
    L = lib()
    sh = (C.c_longlong * 3)(); ft = C.c_int(-1)
    check(L.hfx_host_read_h5_field(str(path).encode(), name.encode(), sh, C.byref(ft), None))
    vals = np.zeros(tuple(sh))
    check(L.hfx_host_read_h5_field(str(path).encode(), name.encode(), sh, C.byref(ft), pd(vals)))
    return ft.value, vals
