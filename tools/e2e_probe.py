"""Dev aid: wall time of hfox.HDGSolver.assemble with pinned host fields next to the event times inside hfx_assemble (1M p=3 tets)."""
import ctypes as C, sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyperfox_b200 import capi, meshgen, hfox
from hyperfox_b200.capi import check, lib
import bench
N=55; dim=3; order=3
verts, lin = meshgen.kuhn_linear(N, dim)
nodes, cells = meshgen.high_order(verts, lin, order)
tp, tau, dirv = bench.poisson_inputs(nodes, cells, order, dim)
nNf = tp["faces"].shape[1]
m = hfox.Mesh(dim, order, "simplex")
m.nodes, m.cells = capi.f64(nodes), capi.i32(cells)
m.faces, m.cell2FaceMap, m.face2CellMap, m.boundaryFaces = tp["faces"], tp["cell2face"], tp["face2cell"], tp["boundary"]
re = m.getReferenceElement()
fm = {"Solution": hfox.Field(m, hfox.Cell, re.getNumNodes(), 1), "Flux": hfox.Field(m, hfox.Cell, re.getNumNodes(), dim),
      "Trace": hfox.Field(m, hfox.Face, nNf, 1), "Tau": hfox.Field(m, hfox.Face, nNf, 1), "Dirichlet": hfox.Field(m, hfox.Face, nNf, 1)}
pin = {k: torch.empty(fm[k].values.size, dtype=torch.float64).pin_memory() for k in ("Tau", "Dirichlet")}
for k, src in (("Tau", tau), ("Dirichlet", dirv)):
    fm[k].values = pin[k].numpy(); fm[k].values[:] = src.ravel()
s = hfox.HDGSolver(device=0)
s.setMesh(m); s.setFieldMap(fm); s.setLinSystem(hfox.CudaLinAlgebraInterface(hfox.PetscOpts(), device=0))
s.setModel(hfox.HDGLaplaceModel(re)); s.setBoundaryCondition(hfox.DirichletModel(re.getFaceElement()), tp["boundary"].astype(np.int32).tolist())
s.initialize(); s.allocate()
L = lib(); a, b = C.c_float(0), C.c_float(0)
for i in range(6):
    torch.cuda.synchronize(); t0 = time.time()
    s.assemble()
    torch.cuda.synchronize(); t1 = time.time()
    L.hfx_last_assemble_ms(s._h(), C.byref(a), C.byref(b))
    print("wall %.2f ms ; inside hfx_assemble: total %.2f ms (clear+kernels %.2f)" % ((t1 - t0) * 1e3, a.value, b.value))
