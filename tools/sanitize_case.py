"""Dev aid: one small assemble + solve per kernel, for compute-sanitizer (memcheck / racecheck / synccheck) runs.
usage: python tools/sanitize_case.py [case index ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import helpers as H
# (dim, order, model, extra): 0 fused p=3, 1 linear tets (sixteen lanes per element), 2 order-2 tets (column per lane), 3 2-D conv-diff, 4 3-D conv-diff (large-element kernel <3,3,256>),
# 5 order-4 tets (large-element kernel <3,4> + chunked block SpMV), 6 explicit solver type (general kernel), 7 SEXPLICIT face solve, 8 continuous Galerkin
cases = [(3, 3, "laplace", None), (3, 1, "laplace", None), (3, 2, "diffsrc", None), (2, 3, "cdrs", None), (3, 3, "cdrs", None), (3, 4, "laplace", None),
         (3, 2, "diffsrc", "wexplicit"), (2, 3, "laplace", "sexplicit"), (3, 2, "laplace", "cg")]
if len(sys.argv) > 1:
    cases = [cases[int(a)] for a in sys.argv[1:]]
for dim, order, model, extra in cases:
    if extra == "cg":
        from hyperfox_b200 import hfox, meshgen
        nodes, cells = meshgen.kuhn_mesh(2, order, dim, perturb=0.1)
        m = hfox.Mesh(dim, order, "simplex"); m.setMesh(nodes, cells)
        re = m.getReferenceElement()
        fm = {"Solution": hfox.Field(m, hfox.Node, 1, 1), "Dirichlet": hfox.Field(m, hfox.Face, re.getFaceElement().getNumNodes(), 1)}
        fm["Dirichlet"].values[:] = 2.0
        s = hfox.CGSolver(); s.setVerbosity(False); s.setMesh(m); s.setFieldMap(fm); s.setLinSystem(hfox.CudaLinAlgebraInterface(hfox.PetscOpts(rtol=1e-8, maxits=50)))
        s.setModel(hfox.LaplaceModel(re)); s.setBoundaryModel(hfox.DirichletModel(re.getFaceElement())); s.initialize(); s.allocate(); s.assemble(); s.solve()
        print("ok", dim, order, model, extra, s.stats.iterations)
        continue
    case = H.make_case(dim, order, N=2, perturb=0.1 if order < 4 else 0.0, model=model, diff="scalar" if model == "cdrs" else "none", tau_double=model not in ("laplace",))
    st = {None: 0, "wexplicit": 1, "sexplicit": 2}[extra]
    if st:
        rng = np.random.default_rng(1)
        nC, nN = case["cells"].shape[0], case["ore"].nNodes
        case["solCur"] = rng.standard_normal((nC, nN)); case["fluxCur"] = rng.standard_normal((nC, nN * dim))
    s, fm, m = H.run_device(case, rtol=1e-8, maxits=50, solverType=st)
    print("ok", dim, order, model, extra, s.lastAssembleKernel(), s.stats.iterations)
