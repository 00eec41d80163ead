"""ctypes binding of libhfx.so (include/hfx.h).  Fails loudly when the CUDA library is missing: there is no CPU fallback."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhfx.so")
_lib = None

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)
lp = C.POINTER(C.c_longlong)


class ErrorHandle(RuntimeError):
    """hfox::ErrorHandle ("Class : function : message", reference src/globals/ErrorHandle.cpp:5-7,41-44)."""


class ModelDesc(C.Structure):
    _fields_ = [("nDOF", C.c_int), ("opmask", C.c_int), ("timeScheme", C.c_int), ("dt", C.c_double)]


class SolveOpts(C.Structure):
    _fields_ = [("ksp", C.c_int), ("pc", C.c_int), ("restart", C.c_int), ("maxits", C.c_int), ("rtol", C.c_double)]


class SolveStats(C.Structure):
    _fields_ = [("iterations", C.c_int), ("resnorm", C.c_double), ("bnorm", C.c_double), ("converged", C.c_int)]


class SolveInfo(C.Structure):
    _fields_ = [("msPerIteration", C.c_float), ("allReduces", C.c_longlong), ("haloExchanges", C.c_longlong), ("haloBytesPerExchange", C.c_longlong),
                ("ownedFaces", C.c_longlong), ("interiorFaces", C.c_longlong), ("boundaryFaces", C.c_longlong), ("nNeighbours", C.c_int), ("msPhase", C.c_float * 4), ("transport", C.c_int)]


# every symbol include/hfx.h declares (tests/test_capi_symbols.py checks the header against this list and the .so)
SYMBOLS = [
    "hfx_ctx_create", "hfx_ctx_destroy", "hfx_last_error", "hfx_device_count", "hfx_fp64_peak", "hfx_dmma_peak", "hfx_refel_set", "hfx_refel_info", "hfx_refel_tables",
    "hfx_refel_host_tables", "hfx_mesh_set", "hfx_mesh_set_topology", "hfx_mesh_sizes", "hfx_mesh_get_topology", "hfx_host_compute_faces", "hfx_host_read_msh", "hfx_host_read_h5_mesh", "hfx_host_write_h5", "hfx_host_h5_info", "hfx_host_read_h5_field", "hfx_host_high_order_mesh",
    "hfx_field_set", "hfx_field_set_async", "hfx_field_get", "hfx_field_size", "hfx_field_lincomb", "hfx_field_diff_norm2", "hfx_model_describe", "hfx_time_scheme_rk", "hfx_ip_coords", "hfx_source_values", "hfx_source_values_n", "hfx_reaction_values",
    "hfx_boundary_describe", "hfx_allocate", "hfx_assemble", "hfx_cg_allocate", "hfx_cg_assemble", "hfx_cg_solve", "hfx_cg_get_csr", "hfx_solver_type", "hfx_solve", "hfx_solve_info", "hfx_recover", "hfx_sync", "hfx_last_assemble_ms", "hfx_last_assemble_kernel", "hfx_assemble_profile", "hfx_get_csr",
    "hfx_get_local", "hfx_get_local_matrix", "hfx_residual", "hfx_get_elem_dofs", "hfx_comm_unique_id", "hfx_comm_init", "hfx_comm_set_halo", "hfx_comm_halo_field", "hfx_host_rcb_partition", "hfx_host_graph_partition", "hfx_plan_create", "hfx_plan_destroy", "hfx_plan_last_error", "hfx_plan_sizes", "hfx_plan_get", "hfx_host_face_canonical_positions", "hfx_host_face_canonical_positions_geom", "hfx_comm_set_halo_plan", "hfx_lai_create", "hfx_lai_destroy", "hfx_lai_last_error", "hfx_lai_set_opts", "hfx_lai_initialize",
    "hfx_lai_configure", "hfx_lai_allocate", "hfx_lai_add_val_matrix", "hfx_lai_add_vals_matrix", "hfx_lai_add_val_rhs", "hfx_lai_add_vals_rhs",
    "hfx_lai_set_val_matrix", "hfx_lai_set_vals_matrix", "hfx_lai_set_val_rhs", "hfx_lai_set_vals_rhs", "hfx_lai_zero_out_rows",
    "hfx_lai_assemble", "hfx_lai_assemble_flush", "hfx_lai_solve", "hfx_lai_get_solution_ownership", "hfx_lai_clear_system",
    "hfx_lai_destroy_system", "hfx_lai_get_num_dofs",
]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ErrorHandle("hfx : load : %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(make -C hyperfox_b200/csrc); the product has no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.hfx_last_error.restype = C.c_char_p
        L.hfx_last_error.argtypes = [C.c_void_p]
        L.hfx_lai_last_error.restype = C.c_char_p
        L.hfx_lai_last_error.argtypes = [C.c_void_p]
        L.hfx_lai_add_val_matrix.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
        L.hfx_lai_set_val_matrix.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
        L.hfx_lai_add_val_rhs.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.hfx_lai_set_val_rhs.argtypes = [C.c_void_p, C.c_int, C.c_double]
        _lib = L
    return _lib


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def pd(a):
    return None if a is None else a.ctypes.data_as(dp)


def pi(a):
    return None if a is None else a.ctypes.data_as(ip)


def check(rc, ctx=None):
    if rc != 0:
        msg = lib().hfx_last_error(ctx)
        raise ErrorHandle(msg.decode() if msg else "hfx : unknown error")


def lcheck(rc, lai):
    if rc != 0:
        msg = lib().hfx_lai_last_error(lai)
        raise ErrorHandle(msg.decode() if msg else "hfx : unknown error")


def device_count():
    return lib().hfx_device_count()


def host_refel_tables(dim, order, geom=0):
    """Product's host table builder (no GPU needed)."""
    L = lib()
    sizes = (C.c_int * 5)()
    check(L.hfx_refel_host_tables(dim, order, geom, sizes, None, None, None, None, None, None, None, None, None))
    nN, nNf, nFc, nIP, nIPf = list(sizes)
    out = dict(dim=dim, order=order, nN=nN, nNf=nNf, nFc=nFc, nIP=nIP, nIPf=nIPf,
               nodes=np.zeros((nN, dim)), ipCoords=np.zeros((nIP, dim)), w=np.zeros(nIP), shape=np.zeros((nIP, nN)),
               dshape=np.zeros((nIP, nN, dim)), fshape=np.zeros((nIPf, nNf)), fdshape=np.zeros((nIPf, nNf, max(dim - 1, 1))), fw=np.zeros(nIPf),
               faceNodes=np.zeros((nFc, nNf), dtype=np.int32))
    check(L.hfx_refel_host_tables(dim, order, geom, sizes, pd(out["nodes"]), pd(out["ipCoords"]), pd(out["w"]), pd(out["shape"]),
                                  pd(out["dshape"]), pd(out["fshape"]), pd(out["fdshape"]), pd(out["fw"]), pi(out["faceNodes"])))
    return out


def host_compute_faces(dim, order, cells, geom=0):
    L = lib()
    cells = i32(cells)
    nC, nN = cells.shape
    nF, nB = C.c_int(0), C.c_int(0)
    t = host_refel_tables(dim, order, geom)
    check(L.hfx_host_compute_faces(dim, order, geom, nC, pi(cells), C.byref(nF), None, None, None, C.byref(nB), None))
    faces = np.zeros((nF.value, t["nNf"]), dtype=np.int32)
    c2f = np.zeros((nC, t["nFc"]), dtype=np.int32)
    f2c = np.zeros((nF.value, 2), dtype=np.int32)
    bnd = np.zeros(nB.value, dtype=np.int32)
    check(L.hfx_host_compute_faces(dim, order, geom, nC, pi(cells), C.byref(nF), pi(faces), pi(c2f), pi(f2c), C.byref(nB), pi(bnd)))
    return dict(faces=faces, cell2face=c2f, face2cell=f2c, boundary=bnd)
