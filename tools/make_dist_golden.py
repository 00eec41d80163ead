"""Golden vector of the multi-GPU parity gate (bench.py --gpus N, tests/test_gpu_dist_gate.py): the ORACLE's solution of the Poisson problem
of tests/dist_solve_check.py (Kuhn 6^3 x 6 = 1296 perturbed tets, order 3, tau = 1, g = sin x e^y), written once, here, on CPU.
bench.py compares the distributed product solve with this file (it must not execute the oracle on its own arm).
usage: python tools/make_dist_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def gate_mesh(N=6):
    """Linear mesh of the gate: same perturbation as tests/dist_solve_check.py."""
    from oracle import meshgen
    verts, lin = meshgen.kuhn_linear(N, 3)
    rng = np.random.default_rng(3)
    interior = np.all((verts > 1e-12) & (verts < 1 - 1e-12), axis=1)
    verts[interior] += 0.1 / N * rng.uniform(-1, 1, size=(int(interior.sum()), 3))
    return verts, lin


def main():
    from oracle import lib as O
    from oracle import meshgen
    from oracle.mesh import compute_faces
    from oracle.refel import ReferenceElement
    N, order = 6, 3
    verts, lin = gate_mesh(N)
    nodes, cells = meshgen.high_order(verts, lin, order)
    ore = ReferenceElement(3, order)
    topo = compute_faces(cells, ore)
    nF, nNf = topo["faces"].shape
    ana = np.sin(nodes[:, 0]) * np.exp(nodes[:, 1])
    dirv = np.zeros((nF, nNf, 1)); b = topo["boundary"]; dirv[b, :, 0] = ana[topo["faces"][b]]
    o = O.HDGOracle(O.RefElC(ore), dict(nodes=nodes, cells=cells, **topo), O.make_model(1, O.OP_DIFFUSION), dict(Tau=np.ones((nF, nNf, 1)), Dirichlet=dirv))
    o.assemble(); o.solve(rtol=1e-13, maxits=20000)
    out = os.path.join(ROOT, "tests", "golden", "dist_gate_kuhn6_p3.npz")
    np.savez_compressed(out, verts=verts, lin=lin, solution=o.sol, iterations=np.int64(o.its), N=np.int64(N), order=np.int64(order))
    print("wrote", out, o.sol.shape, "oracle gmres iterations", o.its)


if __name__ == "__main__":
    main()
