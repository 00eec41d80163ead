"""BASELINE.json configs[4]: order sweep p = 1..5 on 3-D tets at a fixed DOF count (nCells * nN(p) ~ 2e6 here, a 10th of SURVEY 8d's
2e7 so that the whole sweep takes seconds): assemble+condense throughput per order, Laplace, straight-sided Kuhn meshes.
p <= 3 run the fused kernel, p = 4, 5 the general kernel."""
import ctypes as C, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyperfox_b200 import capi, meshgen
from hyperfox_b200.capi import check, lib, pd, pi
FLOPS = {1: 16610, 2: 201490, 3: 1325333, 4: 6220333, 5: 23364077}
NN = {1: 4, 2: 10, 3: 20, 4: 35, 5: 56}
target = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0e6
L = lib()
L.hfx_fp64_peak.restype = C.c_double
peak = L.hfx_fp64_peak(0)
out = []
orders = tuple(int(x) for x in sys.argv[2].split(",")) if len(sys.argv) > 2 else (1, 2, 3, 4, 5)
for order in orders:
    N = max(2, int(round((target / NN[order] / 6.0) ** (1.0 / 3.0))))
    nodes, cells = meshgen.kuhn_mesh(N, order, 3)
    tp = capi.host_compute_faces(3, order, cells)
    nF, nNf = tp["faces"].shape
    h = C.c_void_p()
    check(L.hfx_ctx_create(0, C.byref(h)))
    check(L.hfx_refel_set(h, 3, order, 0), h)
    check(L.hfx_mesh_set(h, nodes.shape[0], pd(nodes), cells.shape[0], pi(cells)), h)
    check(L.hfx_field_set(h, b"Tau", 2, nNf, 1, pd(np.ones((nF, nNf))), 0), h)
    check(L.hfx_field_set(h, b"Dirichlet", 2, nNf, 1, pd(np.zeros((nF, nNf))), 0), h)
    md = capi.ModelDesc(1, 1, 0, 0.0)
    check(L.hfx_model_describe(h, C.byref(md)), h)
    check(L.hfx_boundary_describe(h, 0, 0, None), h)
    check(L.hfx_allocate(h, 0), h)
    a, b = C.c_float(0), C.c_float(0)
    ms = []
    for i in range(5):
        check(L.hfx_assemble(h), h)
        L.hfx_last_assemble_ms(h, C.byref(a), C.byref(b))
        if i >= 2:
            ms.append(b.value)
    t = float(np.mean(ms)) * 1e-3
    nC = cells.shape[0]
    kk = C.c_int(0)
    L.hfx_last_assemble_kernel(h, C.byref(kk), None)
    row = dict(order=order, cubes=N, elements=nC, dofs=nC * NN[order], kernel=("fused", "general", "big", "p1", "col")[kk.value], ms=t * 1e3, elements_per_s=nC / t,
               tflops_algorithmic=FLOPS[order] * nC / t / 1e12, frac_fp64_peak=FLOPS[order] * nC / t / 1e12 / peak)
    out.append(row)
    print(json.dumps(row))
    L.hfx_ctx_destroy(h)
print(json.dumps({"fp64_peak_tflops_measured": peak}))
