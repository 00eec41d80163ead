"""Dev aid: time per GMRES iteration on the assembled p=3 trace system, free of first-call costs (difference of a 90- and a 30-iteration solve)."""
import ctypes as C, sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyperfox_b200 import capi, meshgen
from hyperfox_b200.capi import check, lib, pd, pi
N = int(sys.argv[1]) if len(sys.argv) > 1 else 40
order = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dim = 3
nodes, cells = meshgen.kuhn_mesh(N, order, dim)
tp = capi.host_compute_faces(dim, order, cells)
nF, nNf = tp["faces"].shape
L = lib(); h = C.c_void_p()
check(L.hfx_ctx_create(0, C.byref(h)))
check(L.hfx_refel_set(h, dim, order, 0), h)
check(L.hfx_mesh_set(h, nodes.shape[0], pd(nodes), cells.shape[0], pi(cells)), h)
ana = np.sin(nodes[:, 0]) * np.exp(nodes[:, 1])
dirv = np.zeros((nF, nNf)); dirv[tp["boundary"]] = ana[tp["faces"][tp["boundary"]]]
check(L.hfx_field_set(h, b"Tau", 2, nNf, 1, pd(np.ones((nF, nNf))), 0), h)
check(L.hfx_field_set(h, b"Dirichlet", 2, nNf, 1, pd(dirv), 0), h)
md = capi.ModelDesc(1, 1, 0, 0.0)
check(L.hfx_model_describe(h, C.byref(md)), h)
check(L.hfx_boundary_describe(h, 0, 0, None), h)
check(L.hfx_allocate(h, 0), h)
check(L.hfx_assemble(h), h)
nnz = C.c_longlong(0); nrows = C.c_longlong(0)
check(L.hfx_get_csr(h, C.byref(nrows), C.byref(nnz), None, None, None, None), h)
def run(its):
    so = capi.SolveOpts(0, 1, 30, its, 1e-30); st = capi.SolveStats()
    check(L.hfx_sync(h), h); t0 = time.time()
    check(L.hfx_solve(h, C.byref(so), C.byref(st)), h); check(L.hfx_sync(h), h)
    return time.time() - t0, st.iterations
run(30)
best = 1e9
for _ in range(3):
    t30, i30 = run(30); t90, i90 = run(90)
    best = min(best, (t90 - t30) / max(i90 - i30, 1))
print("tets %d, trace dofs %d, nnz %d: %.3f ms per GMRES(30) iteration; matrix stream alone = %.0f GB/s" % (cells.shape[0], nrows.value, nnz.value, best * 1e3, 8.0 * nnz.value / best / 1e9))
