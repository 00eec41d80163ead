"""Throughput of the device CG assembly (hfx_cg_assemble: LaplaceModel + DirichletModel) on a synthetic Kuhn mesh: elements/s, HBM-side traffic estimate, and the
Krylov iteration time on the node-based CSR.  usage: python tools/bench_cg.py [cubes=30] [order=3]"""
import ctypes as C, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hyperfox_b200 import capi, meshgen
from hyperfox_b200.capi import check, lib, pd, pi
N = int(sys.argv[1]) if len(sys.argv) > 1 else 30
order = int(sys.argv[2]) if len(sys.argv) > 2 else 3
L = lib()
nodes, cells = meshgen.kuhn_mesh(N, order, 3)
tp = capi.host_compute_faces(3, order, cells)
nF, nNf = tp["faces"].shape
h = C.c_void_p()
check(L.hfx_ctx_create(0, C.byref(h)))
check(L.hfx_refel_set(h, 3, order, 0), h)
check(L.hfx_mesh_set(h, nodes.shape[0], pd(nodes), cells.shape[0], pi(cells)), h)
sol = np.zeros(nodes.shape[0]); dirv = np.zeros((nF, nNf))
ana = np.sin(nodes[:, 0]) * np.exp(nodes[:, 1])
dirv[tp["boundary"]] = ana[tp["faces"][tp["boundary"]]]
check(L.hfx_field_set(h, b"Solution", 0, 1, 1, pd(sol), 0), h)
check(L.hfx_field_set(h, b"Dirichlet", 2, nNf, 1, pd(dirv), 0), h)
md = capi.ModelDesc(1, 1, 0, 0.0)
check(L.hfx_model_describe(h, C.byref(md)), h)
check(L.hfx_boundary_describe(h, 0, 0, None), h)
t0 = time.time(); check(L.hfx_cg_allocate(h), h); tAlloc = time.time() - t0
n, nnz = C.c_longlong(0), C.c_longlong(0)
check(L.hfx_cg_get_csr(h, C.byref(n), C.byref(nnz), None, None, None, None), h)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ms = []
for i in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    check(L.hfx_cg_assemble(h), h)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    if i >= 3:
        ms.append(dt * 1e3)
t = float(np.mean(ms)) * 1e-3
o = capi.SolveOpts(0, 1, 30, 20000, 1e-12)
st = capi.SolveStats()
torch.cuda.synchronize(); t0 = time.perf_counter()
check(L.hfx_cg_solve(h, C.byref(o), C.byref(st)), h)
torch.cuda.synchronize(); tSolve = time.perf_counter() - t0
check(L.hfx_field_get(h, b"Solution", pd(sol)), h)
err = float(np.sqrt(((sol - ana) ** 2).sum() / (ana ** 2).sum()))
nC, nN = cells.shape
print(json.dumps({"metric": "CG elements assembled/s (LaplaceModel, order %d tets)" % order, "value": nC / t, "unit": "elements/s", "elements": nC, "nodes": int(n.value), "csr_nnz": int(nnz.value),
                  "ms_per_assemble_wall": t * 1e3, "timing": "wall clock around hfx_cg_assemble with device synchronisation (includes the clearing memsets and the Dirichlet rows)",
                  "pattern_build_s_host": tAlloc, "matrix_GBs_written_lower_bound": nnz.value * 8 / t / 1e9, "atomic_adds_per_s": nC * nN * nN / t,
                  "gmres_iterations": st.iterations, "gmres_ms_per_iteration_wall": tSolve * 1e3 / max(1, st.iterations), "converged": st.converged,
                  "nodal_rel_l2_error_vs_harmonic": err}))
