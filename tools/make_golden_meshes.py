#!/usr/bin/env python3
"""Dev-time: convert the reference's mesh fixtures into tests/golden/meshes/*.npz (nodes f8, cells i4).

Sources: /root/reference/ressources/meshes/lightTri2.h5 (tests/unittests/solver/TestHDGSolver.cpp:16-24) and
ressources/meshes/regression/regression_dim-{2,3}_h-*_ord-*.h5 (tests/regression/HDG/*.cpp).
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from mini_h5 import MiniH5

REF = "/root/reference/ressources/meshes"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "meshes")
os.makedirs(OUT, exist_ok=True)
todo = ["lightTri2.h5", "lightTri.h5"]
for h in ("3e-1", "2e-1", "1e-1"):
    for o in range(1, 6):
        if h == "1e-1" and o == 5:
            continue
        todo.append("regression/regression_dim-2_h-%s_ord-%d.h5" % (h, o))
for o in range(1, 6):
    todo.append("regression/regression_dim-3_h-3e-1_ord-%d.h5" % o)
for o in range(1, 4):
    todo.append("regression/regression_dim-3_h-2e-1_ord-%d.h5" % o)
for t in todo:
    f = MiniH5(os.path.join(REF, t))
    nodes, cells = f.read("/Mesh/Nodes"), f.read("/Mesh/Cells")
    name = os.path.basename(t)[:-3] + ".npz"
    np.savez_compressed(os.path.join(OUT, name), nodes=nodes, cells=cells.astype(np.int32))
    print(name, nodes.shape, cells.shape)

# the Gmsh sources of the regression fixtures (inputs of tools/convertGmsh2H5HO in the reference): small ones only, stored gzip-compressed as
# test data so that tests/test_meshio.py can regenerate the .h5 fixtures above from them
import gzip
import shutil
os.makedirs(os.path.join(OUT, "msh"), exist_ok=True)
for t in ("regression_dim-2_h-3e-1", "regression_dim-2_h-2e-1", "regression_dim-2_h-1e-1", "regression_dim-3_h-3e-1", "regression_dim-3_h-2e-1"):
    with open(os.path.join(REF, "regression", t + ".msh"), "rb") as i, gzip.GzipFile(os.path.join(OUT, "msh", t + ".msh.gz"), "wb", mtime=0) as o:
        shutil.copyfileobj(i, o)
    print("msh/" + t + ".msh.gz")

# a few of the reference's .h5 files, gzip-compressed (test data for the dependency-free HDF5 reader of the product, hfx_host_read_h5_mesh)
os.makedirs(os.path.join(OUT, "h5"), exist_ok=True)
for t in ("lightTri2.h5", "regression/regression_dim-2_h-2e-1_ord-2.h5", "regression/regression_dim-3_h-2e-1_ord-3.h5", "regression/regression_dim-3_h-3e-1_ord-5.h5"):
    with open(os.path.join(REF, t), "rb") as i, gzip.GzipFile(os.path.join(OUT, "h5", os.path.basename(t) + ".gz"), "wb", mtime=0) as o:
        shutil.copyfileobj(i, o)
    print("h5/" + os.path.basename(t) + ".gz")
