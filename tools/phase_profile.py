"""Dev aid: per-phase cycle counts of the fused assemble kernel (CTA 0), run on the GPU box."""
import ctypes as C, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyperfox_b200 import capi, meshgen
from hyperfox_b200.capi import check, lib, pd, pi
N = int(sys.argv[1]) if len(sys.argv) > 1 else 12
order = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dim = 3
nodes, cells = meshgen.kuhn_mesh(N, order, dim)
tp = capi.host_compute_faces(dim, order, cells)
nF, nNf = tp["faces"].shape
L = lib(); h = C.c_void_p()
check(L.hfx_ctx_create(0, C.byref(h)))
check(L.hfx_refel_set(h, dim, order, 0), h)
check(L.hfx_mesh_set(h, nodes.shape[0], pd(nodes), cells.shape[0], pi(cells)), h)
tau = np.ones((nF, nNf)); dirv = np.zeros((nF, nNf))
check(L.hfx_field_set(h, b"Tau", 2, nNf, 1, pd(tau), 0), h)
check(L.hfx_field_set(h, b"Dirichlet", 2, nNf, 1, pd(dirv), 0), h)
import os
md = capi.ModelDesc(1, 1 | ((1 << 30) if os.environ.get('HFX_EXCL') else 0), 0, 0.0)
check(L.hfx_model_describe(h, C.byref(md)), h)
check(L.hfx_boundary_describe(h, 0, 0, None), h)
check(L.hfx_allocate(h, 0), h)
for _ in range(3):
    check(L.hfx_assemble(h), h)
cyc = np.zeros(16, dtype=np.int64)
check(L.hfx_assemble_profile(h, cyc.ctypes.data_as(capi.lp)), h)
a, b = C.c_float(0), C.c_float(0)
check(L.hfx_assemble(h), h)
L.hfx_last_assemble_ms(h, C.byref(a), C.byref(b))
nC = cells.shape[0]
grid = min(nC, 148 * int(os.environ.get('HFX_CTAS_PER_SM', '2')))
per = nC / grid
names = ["P0 commit gather", "P1b per-IP geometry (+prefetch issue)", "P2 g", "P3a M", "P3b contractions", "P1a raw jacobians", "P3c Suq(D)", "P3d faceparts", "P4 A,B", "P5 K", "P6b invK (CTA)", "P7 U", "P8 Q", "P9 S+scatter", "P3b0 invM (CTA)", "P10 write-out"]
if order >= 4 or os.environ.get("HFX_FORCE_GENERIC"):
    names = ["gather+zero Lm", "geometry @ IPs", "gradients, face matrices, M", "uu,uq,qu blocks", "ul,lu,ql,lq,ll + src", "UNabU rhs / time scheme", "invert M", "A,B = W [Squ|Sql]", "K, R", "invert K", "U", "Q", "write U,Q", "S + scatter", "-", "-"]
kk = C.c_int(0)
L.hfx_last_assemble_kernel(h, C.byref(kk), None)
if kk.value == 2:
    names = ["P0 gather", "PG geometry", "PA SJ + point weights", "PB face masses / CG", "PC Suu, Fu", "PD K, R", "PE invert K", "PF U + refinement", "PQ Q", "PZ Zq (+ U,Q bulk stores)", "PS S", "PW write-out", "-", "-", "-", "-"]
    grid = min(nC, 148)
    per = nC / grid
tot = cyc.sum()
print("elements %d, kernel %.3f ms -> %.2f M el/s ; CTA0 handled ~%.1f elements, %.0f cycles/element" % (nC, b.value, nC / b.value / 1e3, per, tot / per))
for n, c in zip(names, cyc):
    if c:
        print("  %-22s %9.0f cyc/elem  %5.1f%%" % (n, c / per, 100.0 * c / tot))
