#!/usr/bin/env python3
"""Hex variant of the headline (SURVEY.md section 8d "Hex variant: the N^3 hexes directly"; north_star: "synthetic structured tet/hex meshes ... at 1, 2, 4 and 8 GPUs"):
3-D Poisson HDG on N^3 order-2 hexahedra (the reference element's maximum order for hexes), assemble + condense throughput and the distributed GMRES iteration.
One process per GPU (torchrun for N > 1), strong scaling on the fixed mesh, rank 0 prints one JSON line.  Structured (parallelepiped) order-2 hexahedra take hdg_big_kernel<BigHexP2, 512>; everything else on orthotopes the general kernel.
  python tools/bench_hex.py [--cubes 48] [--order 2] [--steps 5] [--warmup 3]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 tools/bench_hex.py --gpus 2"""
import argparse, ctypes as C, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1); ap.add_argument("--cubes", type=int, default=48); ap.add_argument("--order", type=int, default=2)
    ap.add_argument("--steps", type=int, default=5); ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    import torch, torch.distributed as dist
    from hyperfox_b200 import capi, meshgen, partition
    from hyperfox_b200.capi import check, lib
    from hyperfox_b200.dist import DistributedPoisson
    rank, world, lrank = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lrank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
    gnodes, gcells = meshgen.box_mesh(a.cubes, a.order, 3)
    verts, lin = gnodes, np.ascontiguousarray(gcells[:, :8])
    part = partition.rcb_partition_vector_c(verts, lin, world, geom=1) if world > 1 else np.zeros(lin.shape[0], dtype=np.int32)
    dp = DistributedPoisson(verts, lin, part, rank, world, a.order, device=lrank, rtol=1e-30, maxits=60, geom="orthotope", global_mesh=(gnodes, gcells))
    L, h = lib(), dp.solver._h()
    ms = []
    x, y = C.c_float(0), C.c_float(0)
    for i in range(a.warmup + a.steps):
        if i == a.warmup:
            if world > 1: dist.barrier()
            torch.cuda.synchronize()
        dp.assemble()
        L.hfx_last_assemble_ms(h, C.byref(x), C.byref(y))
        if i >= a.warmup: ms.append(x.value)
    dp.solve()                                    # 60 GMRES iterations (not to convergence): per-iteration device time
    info = capi.SolveInfo(); check(L.hfx_solve_info(h, C.byref(info)), h)
    red = torch.tensor([float(np.mean(ms)), info.msPerIteration], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(dp.prob["owned_cells"].size)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX); dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    kk = C.c_int(0)
    L.hfx_last_assemble_kernel(h, C.byref(kk), None)
    nCellsLocal, nOwnedLocal = int(dp.mesh.getNumberCells()), int(dp.prob["owned_cells"].size)
    dp.close()   # collective teardown at the same point on every rank
    kname = ("hdg_assemble_kernel", "hdg_generic_kernel (general kernel)", "hdg_big_kernel<BigHexP2, 512> (large-element formulation with the orthotope frame)", "hdg_p1_kernel")[kk.value]
    if rank == 0:
        print(json.dumps({"metric": "HDG elements assembled+condensed/s (p=%d 3D hexes)" % a.order, "value": float(tot.item()) / (float(red[0]) * 1e-3), "unit": "elements/s", "n_gpus": world,
                          "steps": a.steps, "warmup": a.warmup, "ms_per_step": float(red[0]), "higher_is_better": True, "scaling": "strong", "dtype": "f64", "data": "synthetic",
                          "config": {"workload": "3D Poisson HDG order %d on %d^3 = %d structured hexahedra, HDGLaplaceModel + DirichletModel, tau = 1; %s" % (a.order, a.cubes, a.cubes ** 3, kname),
                                     "partition": "recursive coordinate bisection, plan in host C++ (hfx_plan_create, orthotope cells)" if world > 1 else "single rank",
                                     "elements_rank0": nCellsLocal, "owned_elements_rank0": nOwnedLocal},
                          "gmres_ms_per_iteration": float(red[1]), "transport": int(info.transport), "halo_bytes_per_exchange_rank0": int(info.haloBytesPerExchange)}), flush=True)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
