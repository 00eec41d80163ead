"""Dev aid: throughput of the assemble path for any (dim, order, nDOF) -- general kernel when the fused one has no instantiation."""
import ctypes as C, sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyperfox_b200 import capi, meshgen
from hyperfox_b200.capi import check, lib, pd, pi
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10
order = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dim = int(sys.argv[3]) if len(sys.argv) > 3 else 3
geom = 1 if len(sys.argv) > 4 and sys.argv[4] in ("orthotope", "hex", "quad") else 0     # N^dim quads / hexes instead of Kuhn simplices
nodes, cells = meshgen.box_mesh(N, order, dim) if geom else meshgen.kuhn_mesh(N, order, dim)
tp = capi.host_compute_faces(dim, order, cells, geom)
nF, nNf = tp["faces"].shape
L = lib(); h = C.c_void_p()
check(L.hfx_ctx_create(0, C.byref(h)))
check(L.hfx_refel_set(h, dim, order, geom), h)
check(L.hfx_mesh_set(h, nodes.shape[0], pd(nodes), cells.shape[0], pi(cells)), h)
tau = np.ones((nF, nNf)); dirv = np.zeros((nF, nNf))
check(L.hfx_field_set(h, b"Tau", 2, nNf, 1, pd(tau), 0), h)
check(L.hfx_field_set(h, b"Dirichlet", 2, nNf, 1, pd(dirv), 0), h)
md = capi.ModelDesc(1, 1, 0, 0.0)
check(L.hfx_model_describe(h, C.byref(md)), h)
check(L.hfx_boundary_describe(h, 0, 0, None), h)
check(L.hfx_allocate(h, 0), h)
a, b = C.c_float(0), C.c_float(0)
for _ in range(3):
    check(L.hfx_assemble(h), h)
    L.hfx_last_assemble_ms(h, C.byref(a), C.byref(b))
print(("orthotope " if geom else "simplex ") + "dim %d order %d: %d elements, kernel %.3f ms -> %.3f M el/s" % (dim, order, cells.shape[0], b.value, cells.shape[0] / b.value / 1e3))
