"""Dev aid: one small assemble + solve per element size, for compute-sanitizer (memcheck / racecheck / synccheck) runs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import helpers as H
cases = [(3, 3, "laplace"), (3, 1, "laplace"), (3, 2, "diffsrc"), (2, 3, "cdrs"), (3, 3, "cdrs")]
if len(sys.argv) > 1:
    cases = [cases[int(a)] for a in sys.argv[1:]]
for dim, order, model in cases:
    case = H.make_case(dim, order, N=2, perturb=0.1, model=model, diff="scalar" if model == "cdrs" else "none", tau_double=model != "laplace")
    s, fm, m = H.run_device(case, rtol=1e-8, maxits=50)
    print("ok", dim, order, model, s.stats.iterations)
