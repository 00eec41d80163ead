"""Diagnostic (GPU box): device blocks vs the oracle on meshes scaled to the benchmark's element size, Laplace and the
convection-dominated configs[3] fields.  Prints normwise (max|a-b| / max|b|) and entrywise (max |a-b| / (|b| + floor max|b|)) errors."""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
from tests import helpers as H


def errs(a, b, floor=1e-3):
    a, b = np.asarray(a, float).ravel(), np.asarray(b, float).ravel()
    sc = np.abs(b).max() or 1.0
    return float(np.abs(a - b).max() / sc), float((np.abs(a - b) / (np.abs(b) + floor * sc)).max())


out = []
for dim, order, N in ((3, 3, 3), (3, 4, 2), (3, 2, 3), (2, 2, 4), (3, 1, 3), (3, 5, 2)):
    for model in ("laplace", "cd"):
        for scale in (1.0, 3.0 / 55.0, 3.0 / 110.0):
            for perturb in (0.0, 0.12):
                case = H.make_case(dim, order, N=N, perturb=perturb, model=model, scale=scale, seed=3)
                if model == "cd":
                    H.config4_fields(case)
                o = H.run_oracle(case, solve=True, rtol=1e-13)
                try:
                    s, fm, m = H.run_device(case, solve=True, rtol=1e-13)
                except Exception as ex:
                    print("FAIL", dim, order, model, scale, perturb, ex, flush=True)
                    continue
                loc = s.getLocal()
                rowptr, col, vals, rhs = s.getCSR()
                rec = dict(dim=dim, order=order, model=model, scale=scale, perturb=perturb, its=int(s.stats.iterations), oits=int(o.its))
                for name, ref in (("U", o.U), ("Q", o.Q), ("S", o.S), ("U0", o.U0), ("Q0", o.Q0), ("S0", o.S0)):
                    rec[name] = errs(loc[name], ref)
                rec["vals"] = errs(vals, o.vals)
                rec["rhs"] = errs(rhs, o.rhs)
                rec["Trace"] = errs(fm["Trace"].values, o.trace)
                rec["Solution"] = errs(fm["Solution"].values, o.sol)
                rec["Flux"] = errs(fm["Flux"].values, o.flux)
                out.append(rec)
                print(json.dumps(rec), flush=True)
json.dump(out, open("gpurun_out/fine_mesh_diag.json", "w"), indent=1)
