"""Per-kernel view of the Krylov iteration (run under `ncu --metrics gpu__time_duration.sum`, B200_PROFILING.md): the benchmark's
1M-tet order-3 Poisson system (or --cubes N), one assemble, then `--its` GMRES iterations."""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cubes", type=int, default=55)
    ap.add_argument("--order", type=int, default=3)
    ap.add_argument("--its", type=int, default=31)
    ap.add_argument("--pc", type=int, default=1)
    a = ap.parse_args()
    from bench import poisson_inputs
    from hyperfox_b200 import capi, meshgen
    from hyperfox_b200.capi import check, lib, pd, pi
    nodes, cells = meshgen.kuhn_mesh(a.cubes, a.order, 3)
    tp, tau, dirv = poisson_inputs(nodes, cells, a.order, 3)
    L = lib()
    h = C.c_void_p()
    check(L.hfx_ctx_create(0, C.byref(h)))
    check(L.hfx_refel_set(h, 3, a.order, 0), h)
    check(L.hfx_mesh_set(h, nodes.shape[0], pd(nodes), cells.shape[0], pi(cells)), h)
    nNf = tp["faces"].shape[1]
    check(L.hfx_field_set(h, b"Tau", 2, nNf, 1, pd(tau), 0), h)
    check(L.hfx_field_set(h, b"Dirichlet", 2, nNf, 1, pd(dirv), 0), h)
    md = capi.ModelDesc(1, 1, 0, 0.0)
    check(L.hfx_model_describe(h, C.byref(md)), h)
    check(L.hfx_boundary_describe(h, 0, 0, None), h)
    check(L.hfx_allocate(h, 0), h)
    check(L.hfx_assemble(h), h)
    info = capi.SolveInfo()
    for its in (a.its, a.its):
        so = capi.SolveOpts(0, a.pc, 30, its, 1e-30)
        st = capi.SolveStats()
        check(L.hfx_solve(h, C.byref(so), C.byref(st)), h)
        check(L.hfx_solve_info(h, C.byref(info)), h)
        print("solve: %d iterations, %.3f ms / iteration (device); phases op %.3f dots %.3f reduce+step %.3f update %.3f" % ((st.iterations, info.msPerIteration) + tuple(info.msPhase)), flush=True)


if __name__ == "__main__":
    main()
