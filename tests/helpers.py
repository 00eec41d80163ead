"""Shared problem set-ups for the parity tests: the same inputs go to the oracle (CPU restatement of the reference)
and to the CUDA path through the reference-shaped API (hyperfox_b200.hfox)."""
import numpy as np

from hyperfox_b200 import hfox, meshgen
from oracle import lib as O
from oracle.mesh import compute_faces
from oracle.refel import ReferenceElement as OracleRefEl
from tests.conftest import load_mesh


def sin_exp(x):
    return np.sin(x[0]) * np.exp(x[1])


def make_case(dim, order, mesh="kuhn", N=3, perturb=0.1, model="laplace", bc="dirichlet", tau_double=False, diff="none", seed=0, curved=0.0, geom="simplex",
              scale=1.0):
    """Returns dict with numpy inputs in the reference's Field layouts.  curved > 0 displaces every non-vertex node by
    curved*h*U(-1,1)^dim: genuinely curved (non-affine) elements, Jacobians and normals vary from cubature point to point.
    scale shrinks the whole mesh (scale = 3/55 turns the N = 3 Kuhn cube into elements of the benchmark's h = 1/55).
    model "cd": HDGConvectionDiffusionReactionSource with Velocity + DiffusionTensor only (BASELINE configs[3])."""
    rng = np.random.default_rng(seed)
    if geom == "orthotope":     # structured quads / hexes (the reference's orthotope elements: ReferenceElement.cpp:885-1004)
        nodes, cells = meshgen.box_mesh(N, order, dim, perturb=perturb)
    elif mesh == "kuhn":
        nodes, cells = meshgen.kuhn_mesh(N, order, dim, perturb=perturb)
    else:
        nodes, cells = load_mesh(mesh)
    if scale != 1.0:
        nodes = nodes * scale
    if curved > 0.0 and order > 1:
        isv = np.zeros(nodes.shape[0], dtype=bool)
        isv[np.unique(cells[:, :(2 ** dim if geom == "orthotope" else dim + 1)])] = True
        h = np.linalg.norm(nodes[cells[:, 1]] - nodes[cells[:, 0]], axis=1).min()
        nodes = nodes.copy()
        nodes[~isv] += curved * h * np.random.default_rng(seed + 77).uniform(-1, 1, size=(int((~isv).sum()), dim))
    ore = OracleRefEl(dim, order, geom)
    topo = compute_faces(cells, ore)
    nF, nNf, nN = topo["faces"].shape[0], ore.faceElement.nNodes, ore.nNodes
    ana = np.sin(nodes[:, 0]) * np.exp(nodes[:, 1])
    nD = dim if model == "burgers" else 1
    dirv = np.zeros((nF, nNf, nD))
    b = topo["boundary"]
    dirv[b, :, 0] = ana[topo["faces"][b]]
    if nD > 1:
        dirv[b, :, 1] = (np.cos(nodes[:, 0]) * np.exp(-nodes[:, 1]))[topo["faces"][b]]
        if nD > 2:
            dirv[b, :, 2] = (nodes[:, 2] * nodes[:, 0])[topo["faces"][b]]
    case = dict(dim=dim, order=order, geom=geom, nodes=nodes, cells=cells, topo=topo, ore=ore, ana=ana, model=model, bc=bc, nD=nD)
    fields = {"Dirichlet": dirv}
    if nD > 1:     # tau is a full nDOF x nDOF matrix per face node (col-major), HDGBase.cpp:18-32
        blk = 2.0 * np.eye(nD)[None, None] + 0.3 * rng.random((nF, nNf, nD, nD))
        fields["Tau"] = np.concatenate([blk.reshape(nF, nNf, nD * nD), (blk + 0.2).reshape(nF, nNf, nD * nD)], axis=2) if tau_double else blk.reshape(nF, nNf, nD * nD)
        fields["BufferSolution"] = 0.3 * rng.standard_normal((cells.shape[0], nN, nD))
        fields["Trace"] = 0.3 * rng.standard_normal((nF, nNf, nD))
    elif tau_double:
        fields["Tau"] = 0.5 + rng.random((nF, nNf, 2))
    else:
        fields["Tau"] = np.ones((nF, nNf, 1)) if model == "laplace" else 0.5 + rng.random((nF, nNf, 1))
    if diff == "const":      # a scalar diffusion field that happens to be constant: D = c I (the kernels then skip all field work)
        fields["DiffusionTensor"] = np.full((nodes.shape[0], 1), 0.37)
    elif diff == "scalar":
        fields["DiffusionTensor"] = 0.5 + rng.random((nodes.shape[0], 1))
    elif diff == "tensor":
        A = rng.standard_normal((nodes.shape[0], dim, dim)) * 0.2
        D = np.eye(dim)[None] + A @ A.transpose(0, 2, 1) + 0.1 * A    # not symmetric on purpose: exercises the col-major layout
        fields["DiffusionTensor"] = D.transpose(0, 2, 1).reshape(nodes.shape[0], dim * dim)   # col-major per node
    if model in ("cdrs", "cd", "transport", "transport_euler"):
        c = nodes - 0.5
        vel = np.zeros_like(nodes)
        vel[:, 0], vel[:, 1] = -4 * c[:, 1], 4 * c[:, 0]
        fields["Velocity"] = vel
    case["fields"] = fields
    case["source"] = (lambda x: np.exp(-10 * sum((xi - 0.5) ** 2 for xi in x))) if model in ("diffsrc", "cdrs", "euler") else None
    if model == "burgers":
        case["source"] = lambda x, c: (c + 1.0) * np.exp(-3 * sum((xi - 0.4) ** 2 for xi in x))
    case["reaction"] = (lambda x: 1.0 + x[0]) if model == "cdrs" else None
    if model in ("euler", "transport_euler"):
        case["solOld"] = rng.random((cells.shape[0], nN))
    return case


def config4_fields(case, D=1e-2, dt=1e-2):
    """BASELINE configs[3] (tests/parallel/TestParHDGConvectionDiffusionReactionSource.cpp): D = 1e-2 scalar Node field,
    v = 4(-(y-1/2), x-1/2, 0), Tau = |v.n| + D / sqrt(D dt) on both sides of every face (double valued)."""
    nodes, topo, dim = case["nodes"], case["topo"], case["dim"]
    faces = topo["faces"]
    vel = case["fields"]["Velocity"]
    fx = nodes[faces[:, :dim]]                       # the face's vertices come first in its node list
    if dim == 2:
        tvec = fx[:, 1] - fx[:, 0]
        nrm = np.stack([tvec[:, 1], -tvec[:, 0]], axis=1)
    else:
        nrm = np.cross(fx[:, 1] - fx[:, 0], fx[:, 2] - fx[:, 0])
    nrm /= np.linalg.norm(nrm, axis=1)[:, None]
    vdn = np.abs(np.einsum("fad,fd->fa", vel[faces], nrm))
    tau = vdn + D / np.sqrt(D * dt)
    case["fields"]["Tau"] = np.stack([tau, tau], axis=2)
    case["fields"]["DiffusionTensor"] = np.full((nodes.shape[0], 1), D)
    return case


def run_oracle(case, useLU=0, rtol=1e-13, maxits=20000, solve=True, solverType=0):
    ore, topo = case["ore"], case["topo"]
    rc = O.RefElC(ore)
    model = case["model"]
    f = dict(case["fields"])
    mask, diffComps, ts = O.OP_DIFFUSION, 0, O.TS_NONE
    if model == "laplace":
        f.pop("DiffusionTensor", None)
    if "DiffusionTensor" in f:
        diffComps = f["DiffusionTensor"].shape[1]
    if model in ("cdrs", "cd"):
        mask = O.OP_CONVECTION | (O.OP_DIFFUSION if "DiffusionTensor" in f else 0)
    if model in ("transport", "transport_euler"):      # HDGTransport.cpp:47-61: Base + Convection
        mask = O.OP_CONVECTION
    xip = np.einsum("pi,cid->cpd", ore.ipShape, case["nodes"][case["cells"]])
    nD = case.get("nD", 1)
    if model == "burgers":
        mask = O.OP_UNABU | (O.OP_DIFFUSION if "DiffusionTensor" in f else 0) | O.OP_SOURCE
        f["srcIP"] = np.array([[[case["source"](p, c) for p in el] for c in range(case["dim"])] for el in xip])
    elif case["source"] is not None:
        mask |= O.OP_SOURCE
        f["srcIP"] = np.array([[case["source"](p) for p in el] for el in xip])
    if case["reaction"] is not None:
        mask |= O.OP_REACTION
        f["reacIP"] = np.array([[case["reaction"](p) for p in el] for el in xip])
    if model in ("euler", "transport_euler"):
        ts = O.TS_EULER_IMPLICIT
        f["solOld"] = case["solOld"]
    if solverType != 0:      # WEXPLICIT / SEXPLICIT: explicit in the current Solution / Flux (HDGSolver.cpp:346-354)
        f["Solution"] = case["solOld"] if "solOld" in case else case["solCur"]
        f["Flux"] = case["fluxCur"]
    md = O.make_model(nD, mask, diffComps, ts, 0.1)
    mesh = dict(nodes=case["nodes"], cells=case["cells"], **topo)
    h = O.HDGOracle(rc, mesh, md, f, bcKind=O.BC_DIRICHLET if case["bc"] == "dirichlet" else O.BC_INTEGRATED_DIRICHLET, useLU=useLU)
    h.solverType = solverType
    h.assemble()
    if solve:
        if solverType == 2:
            h.solve_faces()
        else:
            h.solve(rtol=rtol, maxits=maxits)
    return h


def run_device(case, rtol=1e-13, maxits=20000, solve=True, keepS=True, recompute=False, solverType=0):
    dim, order = case["dim"], case["order"]
    m = hfox.Mesh(dim, order, case.get("geom", "simplex"))
    m.setMesh(case["nodes"], case["cells"])
    re = m.getReferenceElement()
    nN, nNf = re.getNumNodes(), re.getFaceElement().getNumNodes()
    fm = {}
    nD = case.get("nD", 1)
    fm["Solution"] = hfox.Field(m, hfox.Cell, nN, nD)
    fm["Flux"] = hfox.Field(m, hfox.Cell, nN, dim * nD)
    fm["Trace"] = hfox.Field(m, hfox.Face, nNf, nD)
    tau = case["fields"]["Tau"]
    fm["Tau"] = hfox.Field(m, hfox.Face, nNf, tau.shape[2])
    fm["Tau"].values[:] = tau.ravel()
    if tau.shape[2] == 2 * nD * nD:
        fm["Tau"].setDoubleValued(True)
    if "BufferSolution" in case["fields"]:
        fm["BufferSolution"] = hfox.Field(m, hfox.Cell, nN, nD)
        fm["BufferSolution"].values[:] = case["fields"]["BufferSolution"].ravel()
        fm["Trace"].values[:] = case["fields"]["Trace"].ravel()
    fm["Dirichlet"] = hfox.Field(m, hfox.Face, nNf, nD)
    fm["Dirichlet"].values[:] = case["fields"]["Dirichlet"].ravel()
    if "DiffusionTensor" in case["fields"]:
        d = case["fields"]["DiffusionTensor"]
        fm["DiffusionTensor"] = hfox.Field(m, hfox.Node, 1, d.shape[1])
        fm["DiffusionTensor"].values[:] = d.ravel()
    if "Velocity" in case["fields"]:
        fm["Velocity"] = hfox.Field(m, hfox.Node, 1, dim)
        fm["Velocity"].values[:] = case["fields"]["Velocity"].ravel()
    model = case["model"]
    if model == "laplace":
        mod = hfox.HDGLaplaceModel(re)
    elif model in ("diffsrc", "euler"):
        mod = hfox.HDGDiffusionSource(re)
    elif model == "burgers":
        mod = hfox.HDGBurgersModel(re)
    elif model in ("transport", "transport_euler"):
        mod = hfox.HDGTransport(re)
    else:
        mod = hfox.HDGConvectionDiffusionReactionSource(re)
    if solverType != 0:
        fm["Solution"].values[:] = case.get("solCur", np.zeros(0)).ravel() if "solOld" not in case else case["solOld"].ravel()
        fm["Flux"].values[:] = case["fluxCur"].ravel()
    if model in ("euler", "transport_euler"):
        ts = hfox.Euler(re)
        ts.setTimeStep(0.1)
        mod.setTimeScheme(ts)
        fm["Solution"].values[:] = case["solOld"].ravel()
    bm = hfox.DirichletModel(re.getFaceElement()) if case["bc"] == "dirichlet" else hfox.IntegratedDirichletModel(re.getFaceElement())
    opts = hfox.PetscOpts(rtol=rtol, maxits=maxits, verbose=False)
    lai = hfox.CudaLinAlgebraInterface(opts)
    s = hfox.HDGSolver(keepLocalS=keepS, recomputeRecovery=recompute)
    s.setVerbosity(False)
    if solverType != 0:
        s.setOptions(hfox.HDGSolverOpts(type=solverType, verbosity=False))
    s.setMesh(m)
    s.setFieldMap(fm)
    s.setLinSystem(lai)
    s.setModel(mod)
    s.setBoundaryModel(bm)
    s.initialize()
    s.allocate()
    if case["source"] is not None:
        mod.setSourceFunction(case["source"])
    if case["reaction"] is not None:
        mod.setReactionFunction(case["reaction"])
    s.assemble()
    if solve:
        s.solve()
    return s, fm, m


def rel_err(a, b):
    sc = np.abs(b).max()
    return np.abs(a - b).max() / (sc if sc > 0 else 1.0)
