"""One process per GPU: distributed HDG solve of the reference-shaped problem.

The mesh is partitioned by a cell partition vector (the reference: Zoltan, src/parallel/ZoltanPartitioner.cpp); every rank builds its
local problem deterministically from the global linear mesh (hfx_plan_create, host C++), assembles its owned + ghost elements with no
exchange, and the Krylov solve exchanges only ghost-face trace blocks (NCCL send/recv) and dot products (NCCL all-reduce) inside
libhfx.so.  torch.distributed is used for plumbing only: broadcasting the ncclUniqueId and gathering results.
"""
import ctypes as C

import numpy as np

from . import capi, hfox, meshgen, partition
from .capi import check, lib


def broadcast_unique_id(rank, world):
    """ncclUniqueId created by rank 0 (hfx_comm_unique_id), broadcast through torch.distributed (any backend)."""
    import torch.distributed as dist
    buf = C.create_string_buffer(128)
    if rank == 0:
        check(lib().hfx_comm_unique_id(buf))
    obj = [bytes(buf.raw) if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(obj, src=0)
    return obj[0]


class DistributedPoisson:
    """HDGLaplaceModel + DirichletModel (g = analytic function of x) on a partitioned simplex mesh; rank-local hfox objects."""

    def __init__(self, verts, lin_cells, part, rank, world, order, dim=3, device=0, g=lambda x: np.sin(x[:, 0]) * np.exp(x[:, 1]), rtol=1e-12, maxits=20000,
                 topology=None, geom="simplex", global_mesh=None):
        """geom = "orthotope" (quads / hexes): pass the global order-p mesh as global_mesh = (nodes, cells); verts / lin_cells are then its nodes and the
        vertex columns of its cells (global node ids), and the local mesh is cut out of it (simplices are raised to order p from the local linear cells)."""
        self.rank, self.world, self.dim, self.order = rank, world, dim, order
        gcode = 0 if geom == "simplex" else 1
        if topology is not None:
            c2f, f2c = topology
        else:
            tpg = capi.host_compute_faces(dim, 1, lin_cells, gcode)
            c2f, f2c = tpg["cell2face"], tpg["face2cell"]
        self.plan = partition.Plan(dim, lin_cells, part, rank, world, geom=gcode)          # host C++ behind the C ABI (hfx_plan_create)
        self.prob = p = self.plan.as_problem(verts)
        nv = lin_cells.shape[1]
        if global_mesh is None:
            nodes, cells = meshgen.high_order(p["verts"], p["lin_cells"], order)
            vgid = p["vertex_ids"][p["lin_cells"]]
        else:                     # cut the local cells (owned first, then ghosts) out of the global order-p mesh: local nodes in ascending global id
            gnodes, gcells = global_mesh
            loc = gcells[p["cells_global"]]
            used = np.unique(loc)
            remap = -np.ones(gnodes.shape[0], dtype=np.int64); remap[used] = np.arange(used.size)
            nodes, cells = np.ascontiguousarray(gnodes[used]), remap[loc].astype(np.int32)
            vgid = loc[:, :nv]     # any globally consistent id of the vertices defines the canonical face-node order: the global node id
        self.mesh = m = hfox.Mesh(dim, order, geom)
        m.setMesh(nodes, cells)
        re = m.getReferenceElement()
        nN, nNf = re.getNumNodes(), re.getFaceElement().getNumNodes()
        self.fm = fm = {"Solution": hfox.Field(m, hfox.Cell, nN, 1), "Flux": hfox.Field(m, hfox.Cell, nN, dim), "Trace": hfox.Field(m, hfox.Face, nNf, 1),
                        "Tau": hfox.Field(m, hfox.Face, nNf, 1), "Dirichlet": hfox.Field(m, hfox.Face, nNf, 1)}
        fm["Tau"].values[:] = 1.0
        # Dirichlet data on the faces of the true domain boundary (cuts of the partition are interior faces of the global mesh)
        on_bnd = f2c[p["face_global"], 1] < 0
        self.bfaces = np.flatnonzero(on_bnd).astype(np.int32)
        dirv = np.zeros((m.getNumberFaces(), nNf))
        dirv[self.bfaces] = g(nodes)[m.faces[self.bfaces]]
        fm["Dirichlet"].values[:] = dirv.ravel()
        self.solver = s = hfox.HDGSolver(device=device)
        s.setMesh(m); s.setFieldMap(fm)
        s.setLinSystem(hfox.CudaLinAlgebraInterface(hfox.PetscOpts(rtol=rtol, maxits=maxits), device=device))
        s.setModel(hfox.HDGLaplaceModel(re)); s.setBoundaryCondition(hfox.DirichletModel(re.getFaceElement()), self.bfaces.tolist())
        s.initialize(); s.allocate()
        uid = broadcast_unique_id(rank, world)
        check(lib().hfx_comm_init(s._h(), world, rank, uid), s._h())
        gv = np.full(nodes.shape[0], -1, dtype=np.int64)          # global vertex id of the vertex nodes of the local high-order mesh
        gv[cells[:, :nv]] = vgid
        self.canon = partition.face_canonical_positions_c(dim, order, m.faces, gv, geom=gcode)
        self.plan.set_halo(s._h(), self.canon)      # hfx_comm_set_halo_plan checks that the plan's local mesh is the context's mesh
        self.nodes, self.cells = nodes, cells

    def assemble(self): self.solver.assemble()
    def solve(self): self.solver.solve()

    def close(self):
        """Collective teardown: every rank calls it at the same point (a barrier first: the peers map this rank's halo buffers over CUDA IPC and must be done with them
        before they are freed); the device context is destroyed here and not by the garbage collector."""
        import torch
        torch.cuda.synchronize()
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        ctx = getattr(self.solver, "ctx", None)
        if ctx is not None:
            ctx.close()
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()

    def owned_solution(self):
        """(global cell ids, Solution values [nOwned, nN]) of the cells this rank owns."""
        nO = self.prob["owned_cells"].size
        nN = self.mesh.getReferenceElement().getNumNodes()
        return self.prob["owned_cells"], self.fm["Solution"].values.reshape(-1, nN)[:nO].copy()
