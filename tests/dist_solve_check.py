"""Launched by torchrun (one rank per GPU): distributed Poisson solve vs the single-process oracle on the same global mesh.
usage: torchrun --nproc-per-node N tests/dist_solve_check.py [cubes] [order]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("HFX_DIST_WATCHDOG", "240")), exit=True)   # a hung collective must not hold the GPUs
    import torch
    import torch.distributed as dist
    from hyperfox_b200 import meshgen, partition
    from hyperfox_b200.dist import DistributedPoisson
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    order = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    geom = sys.argv[3] if len(sys.argv) > 3 else "simplex"      # "orthotope": hexahedra (the reference element supports orders 1-2 there)
    rank, world, lrank = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lrank)
    if world > 1:
        dist.init_process_group("gloo")
    dbg = bool(os.environ.get("HFX_DIST_DEBUG"))
    if geom == "orthotope":
        gnodes, gcells = meshgen.box_mesh(N, order, 3, perturb=0.1)
        verts, lin = gnodes, np.ascontiguousarray(gcells[:, :8])
        part = partition.rcb_partition_vector_c(verts, lin, world, geom=1) if world > 1 else np.zeros(lin.shape[0], dtype=np.int32)
        dp = DistributedPoisson(verts, lin, part, rank, world, order, device=lrank, rtol=1e-13, geom="orthotope", global_mesh=(gnodes, gcells))
    else:
        verts, lin = meshgen.kuhn_linear(N, 3)
        rng = np.random.default_rng(3)
        interior = np.all((verts > 1e-12) & (verts < 1 - 1e-12), axis=1)
        verts[interior] += 0.1 / N * rng.uniform(-1, 1, size=(int(interior.sum()), 3))
        part = partition.partition_vector(lin.shape[0], world)
        if dbg: print("rank", rank, "building", flush=True)
        dp = DistributedPoisson(verts, lin, part, rank, world, order, device=lrank, rtol=1e-13)
    if dbg: print("rank", rank, "comm ready; nbrs", dp.prob["nbrs"], [a.size for a in dp.prob["send"]], [a.size for a in dp.prob["recv"]], flush=True)
    dp.assemble()
    if dbg: print("rank", rank, "assembled", flush=True)
    dp.solve()
    if dbg: print("rank", rank, "solved", dp.solver.stats.iterations, flush=True)
    ids, sol = dp.owned_solution()
    out = [None] * world
    if world > 1:
        dist.all_gather_object(out, (ids, sol, dp.solver.stats.iterations))
    else:
        out = [(ids, sol, dp.solver.stats.iterations)]
    dp.close()   # collective teardown at the same point on every rank
    if rank == 0:
        from oracle import lib as O
        from oracle.mesh import compute_faces
        from oracle.refel import ReferenceElement as OracleRefEl
        if geom == "orthotope":
            nodes, cells = gnodes, gcells
            ore = OracleRefEl(3, order, "orthotope")
        else:
            nodes, cells = meshgen.high_order(verts, lin, order)
            ore = OracleRefEl(3, order)
        topo = compute_faces(cells, ore)
        nF, nNf = topo["faces"].shape
        ana = np.sin(nodes[:, 0]) * np.exp(nodes[:, 1])
        dirv = np.zeros((nF, nNf, 1)); b = topo["boundary"]; dirv[b, :, 0] = ana[topo["faces"][b]]
        o = O.HDGOracle(O.RefElC(ore), dict(nodes=nodes, cells=cells, **topo), O.make_model(1, O.OP_DIFFUSION), dict(Tau=np.ones((nF, nNf, 1)), Dirichlet=dirv))
        o.assemble(); o.solve(rtol=1e-13, maxits=20000)
        full = np.zeros_like(o.sol)
        seen = np.zeros(cells.shape[0], dtype=int)
        for ids_r, sol_r, _ in out:
            # node order inside a cell: the local high-order mesh is generated from the same linear cell, so the element-local order agrees
            full[ids_r] = sol_r; seen[ids_r] += 1
        assert np.all(seen == 1), "every cell must be owned by exactly one rank"
        err = np.abs(full - o.sol).max() / np.abs(o.sol).max()
        print("dist_solve_check: world %d, %d %s, order %d, gmres its %s (oracle %d), max rel err vs oracle %.3e" % (world, cells.shape[0], "hexes" if geom == "orthotope" else "tets", order, [x[2] for x in out], o.its, err))
        print("DIST_OK" if err < 1e-10 else "DIST_FAILED")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
