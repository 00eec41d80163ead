"""BASELINE.json configs[3] on a reduced mesh: 3-D convection-diffusion HDG, order 4 (HDGConvectionDiffusionReactionSource with Velocity +
DiffusionTensor, no reaction/source; DirichletModel), synthetic Kuhn mesh N^3 x 6 tets on one B200: assemble+condense throughput of the
general kernel.  Fields as SURVEY.md section 8d: D = 1e-2 (scalar Node field), v = 4 (-(y-1/2), x-1/2, 0), tau = |v.n| + D / sqrt(D dt), dt = 1e-2.
usage: python tools/bench_config4.py [N] [order]"""
import ctypes as C, sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyperfox_b200 import capi, meshgen
from hyperfox_b200.capi import check, lib, pd, pi
N = int(sys.argv[1]) if len(sys.argv) > 1 else 30
order = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dim = 3
FLOPS = {1: 16610, 2: 201490, 3: 1325333, 4: 6220333, 5: 23364077}
nodes, cells = meshgen.kuhn_mesh(N, order, dim)
tp = capi.host_compute_faces(dim, order, cells)
faces = tp["faces"]; nF, nNf = faces.shape
D, dt = 1e-2, 1e-2
vel = np.zeros_like(nodes); vel[:, 0] = -4 * (nodes[:, 1] - 0.5); vel[:, 1] = 4 * (nodes[:, 0] - 0.5)
# face normals of the straight-sided faces from their first three nodes (vertices)
fx = nodes[faces[:, :3]]
nrm = np.cross(fx[:, 1] - fx[:, 0], fx[:, 2] - fx[:, 0]); nrm /= np.linalg.norm(nrm, axis=1)[:, None]
tau = np.abs(np.einsum("fnd,fd->fn", vel[faces], nrm)) + D / np.sqrt(D * dt)
L = lib(); h = C.c_void_p()
L.hfx_fp64_peak.restype = C.c_double
check(L.hfx_ctx_create(0, C.byref(h)))
check(L.hfx_refel_set(h, dim, order, 0), h)
check(L.hfx_mesh_set(h, nodes.shape[0], pd(nodes), cells.shape[0], pi(cells)), h)
check(L.hfx_field_set(h, b"Tau", 2, nNf, 1, pd(np.ascontiguousarray(tau)), 0), h)
check(L.hfx_field_set(h, b"Dirichlet", 2, nNf, 1, pd(np.zeros((nF, nNf))), 0), h)
check(L.hfx_field_set(h, b"DiffusionTensor", 0, 1, 1, pd(np.full((nodes.shape[0], 1), D)), 0), h)
check(L.hfx_field_set(h, b"Velocity", 0, 1, dim, pd(np.ascontiguousarray(vel)), 0), h)
md = capi.ModelDesc(1, 1 | 2, 0, 0.0)        # Diffusion + Convection
check(L.hfx_model_describe(h, C.byref(md)), h)
check(L.hfx_boundary_describe(h, 0, 0, None), h)
check(L.hfx_allocate(h, 0), h)
a, b = C.c_float(0), C.c_float(0)
ms = []
for i in range(4):
    check(L.hfx_assemble(h), h)
    L.hfx_last_assemble_ms(h, C.byref(a), C.byref(b))
    if i: ms.append(b.value)
t = float(np.mean(ms)) * 1e-3
nC = cells.shape[0]
peak = L.hfx_fp64_peak(0)
print(json.dumps(dict(config="3D convection-diffusion HDG, order %d, Kuhn %d^3 x 6" % (order, N), elements=nC, kernel="fused" if order <= 3 else "general",
                      ms=t * 1e3, elements_per_s=nC / t, tflops_algorithmic_laplace_count=FLOPS[order] * nC / t / 1e12,
                      frac_fp64_peak=FLOPS[order] * nC / t / 1e12 / peak)))
