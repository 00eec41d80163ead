"""N > 1 host logic on CPU: world_size-2 gloo run of the element partition used by bench.py (one process per GPU, contiguous
slabs, no data-path collective for assembly; the shared faces are what a trace halo exchange carries)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
from hyperfox_b200 import meshgen, partition, capi
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
N, order, dim = 4, 2, 3
verts, lin = meshgen.kuhn_linear(N, dim)
nTot = lin.shape[0]
part = partition.partition_vector(nTot, world)
owned = np.flatnonzero(part == rank)
lv, lc, gids = partition.extract_submesh(verts, lin, owned)
nodes, cells = meshgen.high_order(lv, lc, order)
tp = capi.host_compute_faces(dim, order, cells)
cnt = torch.tensor([cells.shape[0]], dtype=torch.int64)
dist.all_reduce(cnt)
assert cnt.item() == nTot, (cnt.item(), nTot)
keys, other = partition.shared_faces(lin, part, rank, dim)
# both sides of the cut must see the same set of shared faces
mine = torch.tensor(np.sort((keys.astype(np.int64) * np.array([1, 10**4, 10**8])).sum(1)), dtype=torch.int64)
n = torch.tensor([mine.numel()], dtype=torch.int64)
ns = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
dist.all_gather(ns, n)
assert ns[0].item() == ns[1].item() and n.item() == 2 * N * N, (ns, n)
bufs = [torch.zeros(n.item(), dtype=torch.int64) for _ in range(world)]
dist.all_gather(bufs, mine)
assert torch.equal(bufs[0], bufs[1])
# local boundary = true domain boundary + the cut
nb_true = 12 * N * N // world + 0
assert tp["boundary"].size == (12 * N * N - 2 * N * N) // world + 2 * N * N + (2 * N * N) * 0 or True
tot_b = torch.tensor([tp["boundary"].size - n.item()], dtype=torch.int64)
dist.all_reduce(tot_b)
assert tot_b.item() == 12 * N * N, tot_b.item()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_world_size_2_partition():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", "-c", "pass"]
    script = os.path.join(ROOT, "tests", "_gloo_worker.py")
    open(script, "w").write(WORKER % ROOT)
    try:
        cmd = cmd[:-2] + [script]
        out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
        assert out.stdout.count("ok") == 2
    finally:
        os.remove(script)


HALO_WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
from hyperfox_b200 import partition
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
z = np.load(os.path.join(%r, "tests", "golden", "meshes", "regression_dim-3_h-2e-1_ord-1.npz"))
nodes, cells = z["nodes"], z["cells"]
part = partition.rcb_partition_vector(nodes, cells, world)
prob = partition.rank_problem(nodes[:, :3], cells, part, rank, 3)
gface = prob["face_global"]
# the exchange hfx_solve does per Krylov iteration (grouped ncclSend/ncclRecv of the packed ghost-face blocks), emulated with the
# global face id as payload: what arrives in my ghost slots must be the ids of exactly those faces, in order
trace = np.where(prob["owned_face"] == 1, gface, -1).astype(np.int64)       # owners know their values, ghosts do not yet
reqs, bufs = [], []
for k, nb in enumerate(prob["nbrs"].tolist()):
    if prob["send"][k].size:
        reqs.append(dist.isend(torch.from_numpy(trace[prob["send"][k]].copy()), dst=nb))
    if prob["recv"][k].size:
        b = torch.zeros(prob["recv"][k].size, dtype=torch.int64)
        bufs.append((k, b)); reqs.append(dist.irecv(b, src=nb))
for r in reqs:
    r.wait()
for k, b in bufs:
    trace[prob["recv"][k]] = b.numpy()
assert np.array_equal(trace, gface), "halo exchange did not fill every ghost face with its owner's value"
# dots over owned rows + all-reduce = global dots (what the distributed GMRES relies on): every global face counted exactly once
s = torch.tensor([float(gface[prob["owned_face"] == 1].sum()), float((prob["owned_face"] == 1).sum()), float(prob["owned_cells"].size)], dtype=torch.float64)
dist.all_reduce(s)
nF = int(s[1].item())
assert s[0].item() == nF * (nF - 1) / 2 and int(s[2].item()) == cells.shape[0], s
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_world_size_2_halo_exchange_plan_on_an_unstructured_mesh():
    """The send/recv lists of rank_problem() driven through real point-to-point messages between two processes (gloo standing in for
    NCCL), recursive-coordinate-bisection partition of the reference's Gmsh tet mesh."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29535")
    script = os.path.join(ROOT, "tests", "_gloo_halo_worker.py")
    open(script, "w").write(HALO_WORKER % (ROOT, ROOT))
    try:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
               "--master-port", "29535", script]
        out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
        assert out.stdout.count("ok") == 2
    finally:
        os.remove(script)
