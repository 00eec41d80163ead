"""ORACLE (test infrastructure): ctypes binding of oracle/liboracle.so + the serial HDGSolver driver.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
Driver semantics follow /root/reference/src/solver/HDGSolver.cpp (allocate :5-106, assemble :166-174, solve :677-779).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

OP_DIFFUSION, OP_CONVECTION, OP_REACTION, OP_SOURCE, OP_UNABU = 1, 2, 4, 8, 16
TS_NONE, TS_EULER_IMPLICIT, TS_EULER_EXPLICIT, TS_RK = 0, 1, 2, 3
BC_DIRICHLET, BC_INTEGRATED_DIRICHLET = 0, 1

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_longlong)


class RefEl(C.Structure):
    _fields_ = [("dim", C.c_int), ("nN", C.c_int), ("nNf", C.c_int), ("nFc", C.c_int), ("nIP", C.c_int), ("nIPf", C.c_int),
                ("shape", _dp), ("dshape", _dp), ("w", _dp), ("fshape", _dp), ("fdshape", _dp), ("fw", _dp), ("faceNodes", _ip)]


class Model(C.Structure):
    _fields_ = [("nDOF", C.c_int), ("opmask", C.c_int), ("diffComps", C.c_int), ("timeScheme", C.c_int), ("dt", C.c_double),
                ("rkStage", C.c_int), ("rkNumStages", C.c_int), ("rkRow", _dp)]


class ElFields(C.Structure):
    _fields_ = [(n, _dp) for n in ("nodes", "tau", "diff", "vel", "srcIP", "reacIP", "bufSol", "trace", "solOld", "fluxOld",
                                   "traceOld", "rkSol", "rkFlux", "rkTrace")]


class Mesh(C.Structure):
    _fields_ = [("dim", C.c_int), ("nCells", C.c_int), ("nFaces", C.c_int), ("nNodes", C.c_int),
                ("nodes", _dp), ("cells", _ip), ("faces", _ip), ("cell2face", _ip), ("face2cell", _ip)]


class Fields(C.Structure):
    _fields_ = [("tau", _dp), ("tauVals", C.c_int), ("diff", _dp), ("diffType", C.c_int), ("vel", _dp), ("srcIP", _dp),
                ("reacIP", _dp), ("bufSol", _dp), ("trace", _dp), ("solOld", _dp), ("fluxOld", _dp), ("traceOld", _dp),
                ("rkSol", _dp), ("rkFlux", _dp), ("rkTrace", _dp), ("dirichlet", _dp), ("bFaces", _ip), ("nBFaces", C.c_int),
                ("bcKind", C.c_int)]


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "src", "oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_csr_pattern.restype = C.c_longlong
        _LIB.orc_bench_assemble.restype = C.c_double
        _LIB.orc_gmres.restype = C.c_int
        _LIB.orc_gmres.argtypes = [C.c_longlong, _lp, _ip, _dp, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, _dp]
    return _LIB


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def _c64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _c32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


class RefElC:
    """Keeps the numpy tables alive next to the C struct."""

    def __init__(self, re):
        t = re.tables() if hasattr(re, "tables") else re
        self.t = {k: (_c64(v) if k not in ("faceNodes",) and isinstance(v, np.ndarray) else v) for k, v in t.items()}
        self.t["faceNodes"] = _c32(t["faceNodes"])
        tt = self.t
        self.c = RefEl(tt["dim"], tt["nN"], tt["nNf"], tt["nFc"], tt["nIP"], tt["nIPf"], _d(tt["shape"]), _d(tt["dshape"]), _d(tt["w"]),
                       _d(tt["fshape"]), _d(tt["fdshape"]), _d(tt["fw"]), _i(tt["faceNodes"]))
        for k in ("dim", "nN", "nNf", "nFc", "nIP", "nIPf"):
            setattr(self, k, tt[k])


def make_model(nDOF=1, opmask=OP_DIFFUSION, diffComps=0, timeScheme=TS_NONE, dt=0.0, rkStage=0, rkRow=None):
    row = _c64(rkRow) if rkRow is not None else None
    m = Model(nDOF, opmask, diffComps, timeScheme, dt, rkStage, 0 if row is None else row.size, _d(row))
    m._keep = row
    return m


def sizes(rc, nDOF):
    u = rc.nN * nDOF
    q = u * rc.dim
    l = rc.nFc * rc.nNf * nDOF
    return u, q, l, u + q + l


def element_geometry(rc, nodes):
    nJ = rc.nIP + rc.nFc * rc.nIPf
    d = rc.dim
    jac = np.zeros((nJ, d, d)); inv = np.zeros((nJ, d, d)); dV = np.zeros(nJ); nrm = np.zeros((rc.nFc * rc.nIPf, d))
    nodes = _c64(nodes)
    lib().orc_element_geometry(C.byref(rc.c), _d(nodes), _d(jac), _d(inv), _d(dV), _d(nrm))
    return jac, inv, dV, nrm


def local_system(rc, model, **f):
    u, q, l, n = sizes(rc, model.nDOF)
    keep = {k: _c64(v) for k, v in f.items()}
    ef = ElFields(**{k: _d(v) for k, v in keep.items()})
    A = np.zeros((n, n), order="F"); F = np.zeros(n)
    lib().orc_local_system(C.byref(rc.c), C.byref(model), C.byref(ef), _d(A), _d(F))
    return A, F


def op_base(rc, nDOF, nodes, tau):
    n = sizes(rc, nDOF)[3]
    A = np.zeros((n, n), order="F"); nodes, tau = _c64(nodes), _c64(tau)
    lib().orc_op_base(C.byref(rc.c), nDOF, _d(nodes), _d(tau), _d(A))
    return A


def op_diffusion(rc, nDOF, nodes, diff=None, diffComps=0):
    n = sizes(rc, nDOF)[3]
    A = np.zeros((n, n), order="F"); nodes, diff = _c64(nodes), _c64(diff)
    lib().orc_op_diffusion(C.byref(rc.c), nDOF, _d(nodes), _d(diff), diffComps, _d(A))
    return A


def op_convection(rc, nDOF, nodes, vel):
    n = sizes(rc, nDOF)[3]
    A = np.zeros((n, n), order="F"); nodes, vel = _c64(nodes), _c64(vel)
    lib().orc_op_convection(C.byref(rc.c), nDOF, _d(nodes), _d(vel), _d(A))
    return A


def op_unabu(rc, nDOF, nodes, sol, trace):
    n = sizes(rc, nDOF)[3]
    A = np.zeros((n, n), order="F"); r = np.zeros(n)
    nodes, sol, trace = _c64(nodes), _c64(sol), _c64(trace)
    lib().orc_op_unabu(C.byref(rc.c), nDOF, _d(nodes), _d(sol), _d(trace), _d(A), _d(r))
    return A, r


def op_mass(shape, dV):
    shape, dV = _c64(shape), _c64(dV)
    nIP, nN = shape.shape
    M = np.zeros((nN, nN), order="F")
    lib().orc_op_mass(nN, nIP, _d(shape), _d(dV), _d(M))
    return M


def condense(u, q, l, A, F, useLU=0):
    A = np.asfortranarray(A, dtype=np.float64); F = _c64(F)
    U = np.zeros((u, l), order="F"); Q = np.zeros((q, l), order="F"); S = np.zeros((l, l), order="F")
    U0 = np.zeros(u); Q0 = np.zeros(q); S0 = np.zeros(l)
    lib().orc_condense(u, q, l, _d(A), _d(F), useLU, _d(U), _d(Q), _d(S), _d(U0), _d(Q0), _d(S0))
    return U, Q, S, U0, Q0, S0


class HDGOracle:
    """Serial restatement of HDGSolver driven over a whole mesh.

    mesh: dict(nodes, cells, faces, cell2face, face2cell, boundary); fields: dict of numpy arrays with the reference's
    Field layouts: Tau [nFaces,nNf,tauVals], DiffusionTensor, Velocity, Dirichlet [nFaces,nNf,nDOF],
    srcIP [nCells,nSrc,nIP], reacIP, BufferSolution, Trace, ...
    """

    def __init__(self, refel, mesh, model, fields, bcKind=BC_DIRICHLET, bFaces=None, useLU=0):
        self.rc = refel if isinstance(refel, RefElC) else RefElC(refel)
        self.model = model
        self.useLU = useLU
        self.nDOF = model.nDOF
        self.m = dict(nodes=_c64(mesh["nodes"]), cells=_c32(mesh["cells"]), faces=_c32(mesh["faces"]),
                      cell2face=_c32(mesh["cell2face"]), face2cell=_c32(mesh["face2cell"]))
        self.nCells, self.nFaces = self.m["cells"].shape[0], self.m["faces"].shape[0]
        self.cm = Mesh(self.rc.dim, self.nCells, self.nFaces, self.m["nodes"].shape[0], _d(self.m["nodes"]), _i(self.m["cells"]),
                       _i(self.m["faces"]), _i(self.m["cell2face"]), _i(self.m["face2cell"]))
        self.bFaces = _c32(np.sort(mesh["boundary"] if bFaces is None else bFaces))
        self.bcKind = bcKind
        self.set_fields(fields)
        self.u, self.q, self.l, self.n = sizes(self.rc, self.nDOF)
        self.t = self.rc.nNf * self.nDOF
        self.ndofs = self.nFaces * self.t

    def set_fields(self, fields):
        f = {k: _c64(v) for k, v in fields.items()}
        self.f = f
        tau = f["Tau"]
        tauVals = tau.size // (self.nFaces * self.rc.nNf)
        diff = f.get("DiffusionTensor")
        diffType = 0
        if diff is not None and self.model.diffComps > 0:
            diffType = 0 if diff.size == self.m["nodes"].shape[0] * self.model.diffComps else 1
        self.cf = Fields(_d(tau), tauVals, _d(diff), diffType, _d(f.get("Velocity")), _d(f.get("srcIP")), _d(f.get("reacIP")),
                         _d(f.get("BufferSolution")), _d(f.get("Trace")), _d(f.get("solOld")), _d(f.get("fluxOld")), _d(f.get("traceOld")),
                         _d(f.get("rkSol")), _d(f.get("rkFlux")), _d(f.get("rkTrace")), _d(f.get("Dirichlet")),
                         _i(self.bFaces), self.bFaces.size, self.bcKind)

    def assemble_local(self):
        nC, u, q, l = self.nCells, self.u, self.q, self.l
        self.U = np.zeros((nC, u * l)); self.Q = np.zeros((nC, q * l)); self.S = np.zeros((nC, l * l))
        self.U0 = np.zeros((nC, u)); self.Q0 = np.zeros((nC, q)); self.S0 = np.zeros((nC, l))
        if getattr(self, "solverType", 0) != 0:   # HDGSolverOpts.type = WEXPLICIT (1) / SEXPLICIT (2): explicit in the current Solution / Flux (HDGSolver.cpp:346-354)
            self._sol = _c64(self.f["Solution"]).reshape(nC, u); self._flux = _c64(self.f["Flux"]).reshape(nC, q)
            lib().orc_assemble_local_explicit(C.byref(self.rc.c), C.byref(self.model), C.byref(self.cm), C.byref(self.cf), 0, nC, self.useLU,
                                              _d(self._sol), _d(self._flux), _d(self.U), _d(self.Q), _d(self.S), _d(self.U0), _d(self.Q0), _d(self.S0))
        else:
            lib().orc_assemble_local(C.byref(self.rc.c), C.byref(self.model), C.byref(self.cm), C.byref(self.cf), 0, nC, self.useLU,
                                     _d(self.U), _d(self.Q), _d(self.S), _d(self.U0), _d(self.Q0), _d(self.S0))
        lib().orc_apply_bc(C.byref(self.rc.c), C.byref(self.model), C.byref(self.cm), C.byref(self.cf), _d(self.S), _d(self.S0))

    def pattern(self):
        self.rowptr = np.zeros(self.ndofs + 1, dtype=np.int64)
        nnz = lib().orc_csr_pattern(C.byref(self.rc.c), self.nDOF, C.byref(self.cm), self.rowptr.ctypes.data_as(_lp), None)
        self.colidx = np.zeros(nnz, dtype=np.int32)
        lib().orc_csr_pattern(C.byref(self.rc.c), self.nDOF, C.byref(self.cm), self.rowptr.ctypes.data_as(_lp), _i(self.colidx))
        return self.rowptr, self.colidx

    def elem_dofs(self):
        out = np.zeros((self.nCells, self.l), dtype=np.int32)
        for e in range(self.nCells):
            lib().orc_elem_dofs(C.byref(self.rc.c), self.nDOF, C.byref(self.cm), e, _i(out[e]))
        return out

    def assemble(self):
        self.assemble_local()
        if not hasattr(self, "rowptr"):
            self.pattern()
        self.vals = np.zeros(self.colidx.size)
        self.rhs = np.zeros(self.ndofs)
        lib().orc_scatter(C.byref(self.rc.c), self.nDOF, C.byref(self.cm), _d(self.S), _d(self.S0),
                          self.rowptr.ctypes.data_as(_lp), _i(self.colidx), _d(self.vals), _d(self.rhs))

    def solve_faces(self):
        """HDGSolverOpts.type = SEXPLICIT (HDGSolver.cpp:626-667,709-729): the element blocks S_ll / S0 are accumulated per FACE (face-node order) and every
        face is solved on its own (HouseholderQR there, LAPACK here); then the usual local recovery."""
        t = self.t
        Sf = np.zeros((self.nFaces, t, t)); S0f = np.zeros((self.nFaces, t))
        dofs = self.elem_dofs()
        for e in range(self.nCells):
            Se = self.S[e].reshape(self.l, self.l).T      # column-major element block -> [row, col]
            for i in range(self.rc.nFc):
                d = dofs[e, i * t:(i + 1) * t]; F = d[0] // t; loc = d - F * t
                Sf[F][np.ix_(loc, loc)] += Se[i * t:(i + 1) * t, i * t:(i + 1) * t]
                S0f[F][loc] += self.S0[e, i * t:(i + 1) * t]
        x = np.concatenate([np.linalg.solve(Sf[F], S0f[F]) for F in range(self.nFaces)])
        self.its, self.resnorm = 0, 0.0
        return self._recover(x)

    def _recover(self, x):
        self.trace = x
        self.sol = np.zeros((self.nCells, self.u)); self.flux = np.zeros((self.nCells, self.q))
        lib().orc_recover(C.byref(self.rc.c), self.nDOF, C.byref(self.cm), _d(x), _d(self.U), _d(self.Q), _d(self.U0), _d(self.Q0),
                          _d(self.sol), _d(self.flux))
        return self.trace, self.sol, self.flux

    def solve(self, rtol=1e-6, maxits=1000, restart=30, pc=1, bs=0):
        x = np.zeros(self.ndofs)
        res = C.c_double(0)
        its = lib().orc_gmres(self.ndofs, self.rowptr.ctypes.data_as(_lp), _i(self.colidx), _d(self.vals), _d(self.rhs), _d(x),
                              restart, pc, bs if bs else self.t, rtol, maxits, C.byref(res))
        self.its, self.resnorm = its, res.value
        self.trace = x
        self.sol = np.zeros((self.nCells, self.u)); self.flux = np.zeros((self.nCells, self.q))
        lib().orc_recover(C.byref(self.rc.c), self.nDOF, C.byref(self.cm), _d(x), _d(self.U), _d(self.Q), _d(self.U0), _d(self.Q0),
                          _d(self.sol), _d(self.flux))
        return self.trace, self.sol, self.flux

    def bench_assemble(self, nThreads, useLU=0):
        if not hasattr(self, "rowptr"):
            self.pattern()
        vals = np.zeros(self.colidx.size); rhs = np.zeros(self.ndofs)
        return lib().orc_bench_assemble(C.byref(self.rc.c), C.byref(self.model), C.byref(self.cm), C.byref(self.cf), nThreads, useLU,
                                        self.rowptr.ctypes.data_as(_lp), _i(self.colidx), _d(vals), _d(rhs)), vals, rhs
