import os, subprocess, sys, tempfile
sys.path.insert(0, '/root/repo')
import numpy as np
from tests.test_cpp_mirror import build_cpp_test, MESHES, ROOT
d = tempfile.mkdtemp()
for name in MESHES:
    z = np.load(os.path.join(ROOT, "tests", "golden", "meshes", name + ".npz"))
    nodes, cells = z["nodes"], z["cells"]
    with open(os.path.join(d, name + ".txt"), "w") as f:
        f.write("%d %d %d %d\n" % (nodes.shape[0], nodes.shape[1], cells.shape[0], cells.shape[1]))
        np.savetxt(f, nodes, fmt="%.17g"); np.savetxt(f, cells, fmt="%d")
b = build_cpp_test()
for sec in ("model", "nlw", "partitioner", "solver", "lai", "laplace", "diffsrc", "rk"):
    p = subprocess.run([b, d, sec], capture_output=True, text=True, timeout=600)
    print(sec, p.returncode, p.stdout[-300:], p.stderr[-300:], flush=True)
