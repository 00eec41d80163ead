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
