"""Host-side mirror (Python) of the reference's class surface for the HDG path, over the C ABI of libhfx.so.

Same names, argument meaning and error behaviour as the reference C++ classes, so the parity tests read like the
reference's own tests (tests/unittests/solver/TestHDGSolver.cpp, tests/regression/HDG/TestHDGLaplace.cpp):

  ReferenceElement  <- src/element/ReferenceElement.h          Mesh   <- src/mesh/Mesh.h
  Field             <- src/field/Field.h                        HDG*Model, DirichletModel <- src/model/*.h
  PetscOpts         <- src/resolution/PetscOpts.h               HDGSolver <- src/solver/HDGSolver.h + Solver.h
  CudaLinAlgebraInterface (replaces PetscInterface) <- src/resolution/LinAlgebraInterface.h:23-162
  NonLinearWrapper  <- src/solver/NonLinearWrapper.h

Host Fields hold the values (as in the reference); HDGSolver.assemble() copies the input fields to the GPU, HDGSolver.solve()
copies Trace / Solution / Flux back.  The C++ mirror with the same names lives in include/hyperfox/.
"""
import ctypes as C
import os

import numpy as np

from . import capi
from .capi import ErrorHandle, check, f64, i32, lcheck, lib, pd, pi

Node, Cell, Face = 0, 1, 2      # FieldType (src/field/FieldTypes.h); C ABI: HFX_FIELD_NODE/CELL/FACE
Add, Set = 0, 1                 # AssemblyType.h
KSPGMRES, KSPCG = 0, 1
PCNONE, PCJACOBI, PCBJACOBI = 0, 1, 2   # PCBJACOBI: Jacobi on the t x t diagonal blocks of the faces (the block structure of the trace system)
IMPLICIT, WEXPLICIT, SEXPLICIT = 0, 1, 2    # HDGSolverType (src/solver/HDGSolverOpts.h:6-10)

OP_DIFFUSION, OP_CONVECTION, OP_REACTION, OP_SOURCE, OP_UNABU = 1, 2, 4, 8, 16


class ReferenceElement:
    """ReferenceElement(dim, order, geom): tables come from the product's host builder (csrc/host/hfx_refel.cpp)."""

    def __init__(self, dim, order, geom="simplex"):
        if geom not in ("simplex", "orthotope", "quad", "hex"):
            raise ErrorHandle("ReferenceElement : setGeometry : Element type %s is not yet supported." % geom)
        self._geom = 0 if geom == "simplex" else 1
        self.t = capi.host_refel_tables(dim, order, self._geom)
        self.dim, self.order, self.geom = dim, order, geom
        self._face = None

    def getDimension(self): return self.dim
    def getOrder(self): return self.order
    def getNumNodes(self): return self.t["nN"]
    def getNumIPs(self): return self.t["nIP"]
    def getNumFaces(self): return self.t["nFc"]
    def getNodes(self): return self.t["nodes"]
    def getFaceNodes(self): return self.t["faceNodes"]
    def getIPCoords(self): return self.t["ipCoords"]
    def getIPWeights(self): return self.t["w"]
    def getIPShapeFunctions(self): return self.t["shape"]
    def getIPDerivShapeFunctions(self): return self.t["dshape"]

    def getFaceElement(self):
        if self._face is None and self.dim > 0:
            self._face = ReferenceElement(self.dim - 1, self.order, self.geom)
        return self._face


class Mesh:
    """Mesh(dim, order, geom) + setMesh(nodes, cells); faces/adjacency from the product's topology builder."""

    def __init__(self, dim, order, geom="simplex"):
        self.refEl = ReferenceElement(dim, order, geom)
        self.dim, self.order = dim, order
        self.nodes = self.cells = None
        self.partitioner = None

    def setMesh(self, nodes, cells):
        self.nodes, self.cells = f64(nodes), i32(cells)
        if self.cells.shape[1] != self.refEl.getNumNodes():
            raise ErrorHandle("Mesh : setMesh : the connectivity does not match the reference element")
        if self.nodes.ndim != 2 or self.nodes.shape[1] != self.dim:
            # the device path reads nodes with stride dim (hfx_mesh_set): a node space of another dimension would be mis-strided silently
            raise ErrorHandle("Mesh : setMesh : the device path needs dimNodeSpace == dimension of the reference element (%d coordinates per node given, %d expected)"
                              % (self.nodes.shape[1] if self.nodes.ndim == 2 else -1, self.dim))
        tp = capi.host_compute_faces(self.dim, self.order, self.cells, self.refEl._geom)
        self.faces, self.cell2FaceMap, self.face2CellMap, self.boundaryFaces = tp["faces"], tp["cell2face"], tp["face2cell"], tp["boundary"]

    def getReferenceElement(self): return self.refEl
    def getNodeSpaceDimension(self): return self.nodes.shape[1]
    def getNumberPoints(self): return self.nodes.shape[0]
    def getNumberCells(self): return self.cells.shape[0]
    def getNumberFaces(self): return self.faces.shape[0]
    def getBoundaryFaces(self): return self.boundaryFaces
    def getCell(self, i): return self.cells[i]
    def getFace(self, i): return self.faces[i]
    def getCell2Face(self, i): return self.cell2FaceMap[i]
    def getFace2Cell(self, i): return self.face2CellMap[i][self.face2CellMap[i] >= 0]
    def getPoint(self, i): return self.nodes[i]
    def getSlicePoints(self, ids): return self.nodes[np.asarray(ids)]


class Io:
    """src/io/Io.h: load / write / setMesh / setField."""

    def __init__(self, mesh=None):
        self.myMesh, self.fieldMap = mesh, {}

    def setMesh(self, mesh): self.myMesh = mesh
    def setField(self, name, field): self.fieldMap[name] = field
    def load(self, filename): raise NotImplementedError
    def write(self, filename): raise ErrorHandle("%s : write : writing is not supported by this format" % type(self).__name__)

    def _need_mesh(self, filename, ext):
        if self.myMesh is None:
            raise ErrorHandle("%s : load : the mesh must be set before loading" % type(self).__name__)
        if not str(filename).endswith(ext):
            raise ErrorHandle("%s : load : the file extension must be %s" % (type(self).__name__, ext))


class HDF5Io(Io):
    """HDF5Io (src/io/HDF5Io.cpp) without libhdf5: load = Mesh group (:111-152) + the fields registered with setField (:154-187), write = Mesh group (:189-302) +
    FieldData group (:304-391).  The files carry the reference's FieldType values (FieldTypes.h: Node 0, Face 2, Cell 3)."""
    _TO_FILE = {Node: 0, Face: 2, Cell: 3}
    _FROM_FILE = {0: Node, 2: Face, 3: Cell}

    def load(self, filename):
        from . import meshio
        if self.myMesh is None:
            raise ErrorHandle("HDF5Io : load : must enter a mesh into the io before loading a file.")
        hasMesh, names = meshio.h5_info(filename)
        if not hasMesh and not names:
            raise ErrorHandle("HDFIo : load : could not find Mesh or FieldData groups in file")
        if hasMesh:
            nodes, cells = meshio.read_h5_mesh(filename)
            self.myMesh.setMesh(nodes, cells)
        if names:
            for name, f in self.fieldMap.items():
                if name not in names:
                    raise ErrorHandle("HDF5Io : loadFields : field with name " + name + " was not found in FieldData")
                ft, vals = meshio.read_h5_field(filename, name)
                if ft not in self._FROM_FILE:
                    raise ErrorHandle("HDF5Io : loadFields : edge fields are not supported yet.")
                f.type, f.nObj, f.nVals = self._FROM_FILE[ft], vals.shape[1], vals.shape[2]
                f._deviceNewer = False
                f._values = np.ascontiguousarray(vals.ravel())

    def write(self, filename):
        from . import meshio
        if self.myMesh is None and not self.fieldMap:
            raise ErrorHandle("HDFIo : write : could not find anything to write")
        fields = {}
        for name, f in self.fieldMap.items():
            v = f.values
            nEnt = v.size // max(1, f.nObj * f.nVals)
            fields[name] = (self._TO_FILE[f.type], v.reshape(nEnt, f.nObj, f.nVals))
        m = self.myMesh
        hasMesh = m is not None and getattr(m, "nodes", None) is not None
        meshio.write_h5(filename, m.nodes if hasMesh else None, m.cells if hasMesh else None, fields, mtime=int(__import__("time").time()))


class GmshIo(Io):
    """A Gmsh 2.2 file of linear simplices raised to the order of the mesh's reference element with the node numbering of the
    reference's tools/convertGmsh2H5HO.cpp:117-257 (hfx_host_read_msh + hfx_host_high_order_mesh)."""

    def load(self, filename):
        from . import meshio
        self._need_mesh(filename, ".msh")
        if self.myMesh.refEl.geom != "simplex":
            raise ErrorHandle("GmshIo : load : only simplex meshes can be generated from a Gmsh file")
        nodes, cells = meshio.high_order_from_msh(filename, self.myMesh.dim, self.myMesh.order)
        self.myMesh.setMesh(nodes, cells)


class Field:
    """Field(mesh, type, nObjPerEnt, nValsPerObj): values[(ent*nObj + o)*nVals + v] (src/field/Field.cpp:41-61)."""

    def __init__(self, mesh, ftype, nObjPerEnt, nValsPerObj):
        self.mesh, self.type, self.nObj, self.nVals = mesh, ftype, nObjPerEnt, nValsPerObj
        nEnt = {Node: mesh.getNumberPoints(), Cell: mesh.getNumberCells(), Face: mesh.getNumberFaces()}[ftype]
        self._dev = self._devName = None      # the HDGSolver whose device context holds a copy of this field, and its name there
        self._deviceNewer = False             # the device copy is ahead of the host array (solve / device field arithmetic wrote it)
        self.values = np.zeros(nEnt * nObjPerEnt * nValsPerObj)
        self.doubleValued = False

    # The host array is the user's view.  Results the device produces (HDGSolver.solve, RungeKutta.computeStage / computeSolution, NonLinearWrapper) stay
    # on the device until somebody looks: reading `values` brings them back (into the SAME array object), and from then on the host copy is
    # authoritative again (it may be modified in place, which cannot be observed), so the next assemble uploads it.
    @property
    def values(self):
        if self._deviceNewer:
            self._deviceNewer = False
            check(lib().hfx_field_get(self._dev._h(), self._devName.encode(), pd(self._values)), self._dev._h())
        return self._values

    @values.setter
    def values(self, v):
        self._values = v
        self._deviceNewer = False

    def _on_device(self, solver):
        return self._dev is solver and self._devName is not None

    def getValues(self): return self.values
    def getFieldType(self): return self.type
    def getNumObjPerEnt(self): return self.nObj
    def getNumValsPerObj(self): return self.nVals
    def setDoubleValued(self, b): self.doubleValued = bool(b)
    def isDoubleValued(self): return self.doubleValued
    def getLength(self): return self.values.size


class Euler:
    """TimeScheme: implicit Euler as coded in the reference (src/operator/Euler.cpp:18-37; the explicit flavour has no device kernel)."""

    def __init__(self, refEl, isExplicit=False):
        if isExplicit:
            raise ErrorHandle("Euler : Euler : the explicit Euler scheme has no device kernel")
        self.dt = 0.0

    def setTimeStep(self, dt): self.dt = dt


# RKType (src/operator/RKType.h) and the Butcher tables of RungeKutta::setUpDB (src/operator/RungeKutta.cpp:215-291): row s < nStages =
# [c_s | a_s0 .. a_s,nStages-1], last row = [0 | b_0 .. b_nStages-1]
FEuler, EMidpoint, Heun, Kutta3, Heun3, SSPRK3, RK4, BEuler, IMidpoint, CrankNicolson, KS2, QZ2, ALX2, RK43 = range(14)
_g = 1.0 - np.sqrt(2.0) / 2.0
BUTCHER = {
    FEuler: [[0, 0], [0, 1]],
    EMidpoint: [[0, 0, 0], [0.5, 0.5, 0], [0, 0, 1]],
    Heun: [[0, 0, 0], [1, 1, 0], [0, 0.5, 0.5]],
    Kutta3: [[0, 0, 0, 0], [0.5, 0.5, 0, 0], [1, -1, 2, 0], [0, 1 / 6, 2 / 3, 1 / 6]],
    Heun3: [[0, 0, 0, 0], [1 / 3, 1 / 3, 0, 0], [2 / 3, 0, 2 / 3, 0], [0, 1 / 4, 0, 3 / 4]],
    SSPRK3: [[0, 0, 0, 0], [1, 1, 0, 0], [0.5, 0.25, 0.25, 0], [0, 1 / 6, 1 / 6, 2 / 3]],
    RK4: [[0, 0, 0, 0, 0], [0.5, 0.5, 0, 0, 0], [0.5, 0, 0.5, 0, 0], [1, 0, 0, 1, 0], [0, 1 / 6, 1 / 3, 1 / 3, 1 / 6]],
    BEuler: [[1, 1], [0, 1]],
    IMidpoint: [[0.5, 0.5], [0, 1]],
    CrankNicolson: [[0, 0, 0], [1, 0.5, 0.5], [0, 0.5, 0.5]],
    KS2: [[0.5, 0.5, 0], [1.5, -0.5, 2], [0, -0.5, 1.5]],
    QZ2: [[0.25, 0.25, 0], [0.75, 0.5, 0.25], [0, 0.5, 0.5]],
    ALX2: [[_g, _g, 0], [1, 1 - _g, _g], [0, 1 - _g, _g]],
    RK43: [[0.5, 0.5, 0, 0, 0], [2 / 3, 1 / 6, 0.5, 0, 0], [0.5, -0.5, 0.5, 0.5, 0], [1, 1.5, -1.5, 0.5, 0.5], [0, 1.5, -1.5, 0.5, 0.5]],
}


class RungeKutta:
    """TimeScheme: RungeKutta(refEl, type, auxiliaryFields) (src/operator/RungeKutta.cpp).  apply() runs inside the device assembly
    (hfx_time_scheme_rk); computeStage / computeSolution are the reference's field updates (:145-213) on the host Fields."""
    isRK = True

    def __init__(self, refEl, type=CrankNicolson, fields=()):
        self.setButcherTable(type)
        self.auxiliaryFields = list(fields)
        self.stageCounter = 0
        self.dt = 0.0

    def setButcherTable(self, t):
        if not isinstance(t, (list, np.ndarray)):          # RKType: straight from the database (RungeKutta.cpp:15-17)
            self.bTable = np.array(BUTCHER[t], dtype=float)
            return
        tab = np.array(t, dtype=float)                     # explicit table: the reference's checks, RungeKutta.cpp:19-32 -- the strict upper
        if tab.ndim != 2 or tab.shape[0] != tab.shape[1]:  # triangle of the WHOLE [c | a ; 0 | b] table must vanish (followed as coded)
            raise ErrorHandle("RungeKutta : setButcherTable : the Butcher table should be square")
        if np.any(np.triu(tab, 1) != 0):
            raise ErrorHandle("RungeKutta : setButcherTable : the upper triangular part of the Butcher table should be null (no fully implicit implementation as of yet)")
        self.bTable = tab

    def setTimeStep(self, dt): self.dt = dt
    def getStage(self): return self.stageCounter
    def getNumStages(self): return self.bTable.shape[1] - 1
    def stageRow(self): return np.ascontiguousarray(self.bTable[self.stageCounter, 1:])

    def _device_solver(self, fm):
        """The HDGSolver on whose device the solution fields of `fm` live (after its solve()), if its field map is `fm`: the stage arithmetic then
        runs there.  None: host arithmetic (numpy), e.g. for fields that never met a solver.  At most 8 terms per combination (7 stages)."""
        sol = fm.get("Solution")
        s = getattr(sol, "_dev", None)
        if s is None or getattr(s, "fieldMap", None) is not fm or self.getNumStages() > 7 or os.environ.get("HFX_HOST_FIELD_ARITHMETIC"):
            return None
        return s

    def fieldNames(self):
        """Everything RungeKutta::setFieldMap requires (:44-88) for the current stage."""
        names = ["OldSolution"] + ["Old" + a for a in self.auxiliaryFields]
        for k in range(self.stageCounter):
            names += ["RKStage_%d" % k] + ["RKStage_%s_%d" % (a, k) for a in self.auxiliaryFields]
        return names

    def computeStage(self, fm):
        if self.stageCounter >= self.getNumStages():
            raise ErrorHandle("RungeKutta : computeStage : cannot compute more stages than the method allows, think about computing the solution")
        row, s = self.bTable[self.stageCounter, 1:], self.stageCounter
        solver = self._device_solver(fm)
        if solver is not None:     # RungeKutta.cpp:145-179 as device AXPYs: the fields do not leave the GPU between the stages
            for base in ["Solution"] + self.auxiliaryFields:
                sname = lambda k: "RKStage_%d" % k if base == "Solution" else "RKStage_%s_%d" % (base, k)
                solver.fieldLinComb(sname(s), [1.0 / self.dt, -1.0 / self.dt], [base, "Old" + base])
                solver.fieldLinComb(base, [1.0] + [self.dt * row[j] for j in range(s + 1)], ["Old" + base] + [sname(j) for j in range(s + 1)])
            self.stageCounter += 1
            return
        for base in ["Solution"] + self.auxiliaryFields:
            stage = lambda k: fm["RKStage_%d" % k if base == "Solution" else "RKStage_%s_%d" % (base, k)]
            sol, old = fm[base].values, fm["Old" + base].values
            stage(s).values[:] = (sol - old) / self.dt
            sol[:] = old + self.dt * sum(row[j] * stage(j).values for j in range(s + 1))
        self.stageCounter += 1

    def computeSolution(self, fm):
        if self.stageCounter != self.getNumStages():
            raise ErrorHandle("RungeKutta : computeSolution : all stages must be computed before computing the solution")
        bs = self.bTable[self.stageCounter, 1:]
        solver = self._device_solver(fm)
        if solver is not None:     # RungeKutta.cpp:181-213
            nSt = self.getNumStages()
            for base in ["Solution"] + self.auxiliaryFields:
                sname = lambda k: "RKStage_%d" % k if base == "Solution" else "RKStage_%s_%d" % (base, k)
                solver.fieldLinComb(base, [1.0] + [self.dt * bs[k] for k in range(nSt)], ["Old" + base] + [sname(k) for k in range(nSt)])
            self.stageCounter = 0
            return
        for base in ["Solution"] + self.auxiliaryFields:
            stage = lambda k: fm["RKStage_%d" % k if base == "Solution" else "RKStage_%s_%d" % (base, k)]
            fm[base].values[:] = fm["Old" + base].values + self.dt * sum(bs[k] * stage(k).values for k in range(self.getNumStages()))
        self.stageCounter = 0


class _HDGModel:
    opmask = 0
    needsSource = False

    def __init__(self, refEl):
        self.refEl = refEl
        self.allocated = False
        self.timeScheme = None
        self.sourceFunc = self.reactionFunc = None
        self.nDOF = 1

    def setTimeScheme(self, ts):
        if self.allocated:
            raise ErrorHandle("FEModel : setTimeScheme : the time scheme must be set before allocation or field setting")
        self.timeScheme = ts

    def allocate(self, nDOFsPerNode):
        if nDOFsPerNode < 1:
            raise ErrorHandle("HDGOperator : allocate : the number of DOFs per node must be at least one")
        self.nDOF = nDOFsPerNode
        self.allocated = True

    def getAssemblyType(self): return (Add, Add)

    def _mask(self, fieldNames, strict=True):
        return self.opmask

    # ---- the per-element surface of the reference (src/model/FEModel.h:43-78): setElementNodes / setFieldMap(local values) / compute /
    #      getLocalMatrix / getLocalRHS.  compute() runs the DEVICE operators on a one-element mesh (hfx_get_local_matrix: the general kernel's
    #      dense local system before the condensation), so the reference's model tests (TestHDGLaplaceModel, TestHDGBase identities, ...) can
    #      be run against what the GPU actually assembles.
    def setElementNodes(self, nodes):
        self.elementNodes = f64(np.asarray(nodes, dtype=np.float64))
        if self.elementNodes.ndim != 2 or self.elementNodes.shape[0] != self.refEl.getNumNodes():
            raise ErrorHandle("FEModel : setElementNodes : the number of nodes does not match the reference element")

    def setFieldMap(self, fm):
        """Element-local field values (std::map<std::string, std::vector<double>>): Tau [nFaces x nNodesPerFace x nDOF^2 (x 2 never: one side)],
        DiffusionTensor [nNodes x (1 | dim^2)], Velocity [nNodes x dim], BufferSolution / Solution [nNodes x nDOF], Trace [nFaces x nNodesPerFace x nDOF]."""
        if "Tau" not in fm:
            raise ErrorHandle("HDGModel : setFieldMap : must provide a Tau field")
        self.localFieldMap = {k: np.asarray(v, dtype=np.float64).ravel() for k, v in fm.items()}

    def compute(self, device=0):
        if not self.allocated:
            raise ErrorHandle("FEModel : compute : the model must be allocated before computing")
        if getattr(self, "elementNodes", None) is None:
            raise ErrorHandle("FEModel : compute : the nodes have not been set")
        if getattr(self, "localFieldMap", None) is None:
            raise ErrorHandle("FEModel : compute : the field map has not been set")
        re, nD, lf = self.refEl, self.nDOF, self.localFieldMap
        dim, nN, nFc, nNf = re.getDimension(), re.getNumNodes(), re.getNumFaces(), re.getFaceElement().getNumNodes()
        m = Mesh(dim, re.getOrder(), "simplex" if re._geom == 0 else "orthotope")
        m.setMesh(self.elementNodes, np.arange(nN, dtype=np.int32)[None, :])
        # the faces of a one-element mesh are its local faces, their node lists the element's face nodes: local field values map one to one
        if not np.array_equal(m.cell2FaceMap[0], np.arange(nFc)) or not np.array_equal(m.faces, np.asarray(re.getFaceNodes())):
            raise ErrorHandle("FEModel : compute : unexpected face numbering of the one-element mesh")
        fm = {"Solution": Field(m, Cell, nN, nD), "Flux": Field(m, Cell, nN, nD * dim), "Trace": Field(m, Face, nNf, nD), "Dirichlet": Field(m, Face, nNf, nD)}
        tau = lf["Tau"]
        if tau.size != nFc * nNf * nD * nD:
            raise ErrorHandle("HDGModel : setFieldMap : the Tau field does not have the right size")
        fm["Tau"] = Field(m, Face, nNf, nD * nD); fm["Tau"].values[:] = tau
        for name, ftype, nObj in (("DiffusionTensor", Node, 1), ("Velocity", Node, 1), ("BufferSolution", Cell, nN), ("Solution", Cell, nN), ("Trace", Face, nNf)):
            if name in lf:
                ents = {Node: nN, Cell: 1, Face: nFc}[ftype]
                if lf[name].size % (ents * nObj) != 0:
                    raise ErrorHandle("HDGModel : setFieldMap : the %s field does not have the right size" % name)
                fm[name] = Field(m, ftype, nObj, lf[name].size // (ents * nObj)); fm[name].values[:] = lf[name]
        s = HDGSolver(device=device)
        s.setVerbosity(False); s.setMesh(m); s.setFieldMap(fm); s.setLinSystem(CudaLinAlgebraInterface(PetscOpts(), device=device))
        s.setModel(self); s.setBoundaryModel(DirichletModel(re.getFaceElement())); s.initialize(); s.allocate()
        self.localMatrix, self.localRHS = s.getLocalMatrix(0)

    def getLocalMatrix(self): return self.localMatrix
    def getLocalRHS(self): return self.localRHS


class HDGLaplaceModel(_HDGModel):
    """Base + Diffusion(D = I) (src/model/HDGLaplaceModel.cpp:18-30)."""
    opmask = OP_DIFFUSION

    def _mask(self, fieldNames, strict=True):
        return OP_DIFFUSION

    usesDiffusionField = False


class HDGDiffusionSource(_HDGModel):
    """Base + Diffusion(DiffusionTensor field if given) ; rhs = Source (src/model/HDGDiffusionSource.cpp:43-85)."""
    usesDiffusionField = True

    def setSourceFunction(self, s):
        if not self.allocated:
            raise ErrorHandle("HDGDiffusionSource : setSourceFunction : the model must be allocated before setting the source function")
        self.sourceFunc = s

    def _mask(self, fieldNames, strict=True):
        if self.sourceFunc is None and strict:   # computeLocalRHS always evaluates the source (HDGDiffusionSource.cpp:81-85)
            raise ErrorHandle("Source : calcSource : must set a source function before calculating the source.")
        return OP_DIFFUSION | (OP_SOURCE if self.sourceFunc is not None else 0)


class HDGConvectionDiffusionReactionSource(_HDGModel):
    """Base [+ Convection if Velocity] [+ Diffusion if DiffusionTensor] [+ Reaction] ; rhs = [Source]
    (src/model/HDGConvectionDiffusionReactionSource.cpp:69-108)."""
    usesDiffusionField = True

    def setSourceFunction(self, s):
        if not self.allocated:
            raise ErrorHandle("HDGConvectionDiffusionReactionSource : setSourceFunction : the model must be allocated before setting the source function")
        self.sourceFunc = s

    def setReactionFunction(self, r):
        if not self.allocated:
            raise ErrorHandle("HDGConvectionDiffusionReactionSource : setReactionFunction : the model must be allocated before setting the reaction function")
        self.reactionFunc = r

    def _mask(self, fieldNames, strict=True):
        if "Velocity" not in fieldNames and "DiffusionTensor" not in fieldNames:
            raise ErrorHandle("HDGConvectionDiffusionReactionSource : setFieldMap : must provide at least either a Velocity field or a DiffusionTensor field")
        m = 0
        if "Velocity" in fieldNames: m |= OP_CONVECTION
        if "DiffusionTensor" in fieldNames: m |= OP_DIFFUSION
        if self.reactionFunc is not None: m |= OP_REACTION
        if self.sourceFunc is not None: m |= OP_SOURCE
        return m


class HDGTransport(_HDGModel):
    """HDGTransport (src/model/HDGTransport.cpp:5-69): localMatrix = Base + Convection, zero right-hand side; a Velocity field is required."""
    usesDiffusionField = False

    def setFieldMap(self, fm):
        if "Velocity" not in fm:
            raise ErrorHandle("HDGTransport : setFieldMap : one must provide a Velocity field to use the Transport model.")
        super().setFieldMap(fm)

    def _mask(self, fieldNames, strict=True):
        if strict and "Velocity" not in fieldNames:
            raise ErrorHandle("HDGTransport : setFieldMap : one must provide a Velocity field to use the Transport model.")
        return OP_CONVECTION


class HDGBurgersModel(_HDGModel):
    """Base + HDGUNabU (Newton-linearised convection, needs BufferSolution and Trace) [+ Diffusion if DiffusionTensor] ; rhs = UNabU
    rhs [+ one scalar Source per component]  (src/model/HDGBurgersModel.cpp:5-124)."""
    usesDiffusionField = True
    isBurgers = True

    def allocate(self, nDOFsPerNode):
        if nDOFsPerNode != self.refEl.getDimension():
            raise ErrorHandle("HDGBurgersModel : allocate : the number of DOFs per node must be equal to the number of spatial dimensions for the Burgers equation")
        _HDGModel.allocate(self, nDOFsPerNode)

    def setSourceFunction(self, s):
        """s(x, i): source of component i (std::function<double(const std::vector<double>&, int)>)."""
        if not self.allocated:
            raise ErrorHandle("HDGBurgersModel : setSourceFunction : the model must be allocated before setting the source function")
        self.sourceFunc = s

    def _mask(self, fieldNames, strict=True):
        if "BufferSolution" not in fieldNames:
            raise ErrorHandle("HDGBurgersModel : setFieldMap : must provide a BufferSolution field for the Newton-Raphson iterations")
        m = OP_UNABU
        if "DiffusionTensor" in fieldNames: m |= OP_DIFFUSION
        if self.sourceFunc is not None: m |= OP_SOURCE
        return m


class DirichletModel:
    """assembly = {Set, Set}, localMatrix = I, localRHS = Dirichlet (src/model/DirichletModel.cpp:19-44)."""
    kind = 0

    def __init__(self, faceRefEl):
        self.refEl = faceRefEl

    def allocate(self, nDOFsPerNode): pass
    def getAssemblyType(self): return (Set, Set)


class IntegratedDirichletModel(DirichletModel):
    """localMatrix = face mass, localRHS = M g (src/model/IntegratedDirichletModel.cpp)."""
    kind = 1


class PetscOpts:
    """Same defaults as src/resolution/PetscOpts.h:12-28."""

    def __init__(self, solverType=KSPGMRES, preconditionnerType=PCJACOBI, rtol=1e-6, maxits=1000, verbose=False, restart=30):
        self.solverType, self.preconditionnerType, self.rtol, self.maxits, self.verbose, self.restart = solverType, preconditionnerType, rtol, maxits, verbose, restart

    def c(self):
        return capi.SolveOpts(self.solverType, self.preconditionnerType, self.restart, self.maxits, self.rtol)


class Context:
    def __init__(self, device=0):
        self.h = C.c_void_p()
        check(lib().hfx_ctx_create(device, C.byref(self.h)))
        self.device = device

    def close(self):
        if self.h:
            lib().hfx_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CudaLinAlgebraInterface:
    """Drop-in for PetscInterface behind LinAlgebraInterface (src/resolution/LinAlgebraInterface.h:23-162)."""

    def __init__(self, options=None, context=None, device=0):
        self.opts = options or PetscOpts()
        self.ctx = context or Context(device)
        self.h = C.c_void_p()
        check(lib().hfx_lai_create(self.ctx.h, C.byref(self.h)))
        o = self.opts.c()
        lib().hfx_lai_set_opts(self.h, C.byref(o))
        self.stats = capi.SolveStats()

    def _c(self, rc): lcheck(rc, self.h)
    def initialize(self): self._c(lib().hfx_lai_initialize(self.h))
    def configure(self): self._c(lib().hfx_lai_configure(self.h))

    def allocate(self, ndofs, diagSparsePattern=None, offSparsePattern=None):
        self._c(lib().hfx_lai_allocate(self.h, int(ndofs), None, None))

    def addValMatrix(self, i, j, val): self._c(lib().hfx_lai_add_val_matrix(self.h, int(i), int(j), float(val)))
    def setValMatrix(self, i, j, val): self._c(lib().hfx_lai_set_val_matrix(self.h, int(i), int(j), float(val)))
    def addValRHS(self, i, val): self._c(lib().hfx_lai_add_val_rhs(self.h, int(i), float(val)))
    def setValRHS(self, i, val): self._c(lib().hfx_lai_set_val_rhs(self.h, int(i), float(val)))

    def addValsMatrix(self, is_, js, vals):
        is_, js, vals = i32(is_), i32(js), f64(vals)
        self._c(lib().hfx_lai_add_vals_matrix(self.h, is_.size, pi(is_), js.size, pi(js), pd(vals)))

    def setValsMatrix(self, is_, js, vals):
        is_, js, vals = i32(is_), i32(js), f64(vals)
        self._c(lib().hfx_lai_set_vals_matrix(self.h, is_.size, pi(is_), js.size, pi(js), pd(vals)))

    def addValsRHS(self, is_, vals):
        is_, vals = i32(is_), f64(vals)
        self._c(lib().hfx_lai_add_vals_rhs(self.h, is_.size, pi(is_), pd(vals)))

    def setValsRHS(self, is_, vals):
        is_, vals = i32(is_), f64(vals)
        self._c(lib().hfx_lai_set_vals_rhs(self.h, is_.size, pi(is_), pd(vals)))

    def zeroOutRows(self, is_):
        is_ = i32(is_)
        self._c(lib().hfx_lai_zero_out_rows(self.h, is_.size, pi(is_)))

    def assemble(self): self._c(lib().hfx_lai_assemble(self.h))
    def assembleFlush(self): self._c(lib().hfx_lai_assemble_flush(self.h))
    def clearSystem(self): self._c(lib().hfx_lai_clear_system(self.h))
    def destroySystem(self): self._c(lib().hfx_lai_destroy_system(self.h))

    def getNumDofs(self):
        n = C.c_int(0)
        lib().hfx_lai_get_num_dofs(self.h, C.byref(n))
        return n.value

    def getSolutionOwnership(self):
        lo, hi = C.c_int(0), C.c_int(0)
        self._c(lib().hfx_lai_get_solution_ownership(self.h, C.byref(lo), C.byref(hi)))
        return list(range(lo.value, hi.value))

    def solve(self, solution=None):
        n = self.getNumDofs()
        out = np.zeros(n)
        self._c(lib().hfx_lai_solve(self.h, pd(out), C.byref(self.stats)))
        if solution is not None:
            solution.resize(n, refcheck=False)
            solution[:] = out
        return out

    def __del__(self):
        try:
            if self.h:
                lib().hfx_lai_destroy(self.h)
        except Exception:
            pass


class HDGSolverOpts:
    def __init__(self, type=IMPLICIT, verbosity=False):
        self.type, self.verbosity = type, verbosity


class HDGSolver:
    """HDGSolver: set*/initialize/allocate/assemble/solve with the reference's call-order contract
    (src/solver/Solver.h:35-88, src/solver/HDGSolver.cpp:5-106,166-174,677-779)."""

    INPUT_FIELDS = ("Tau", "Dirichlet", "DiffusionTensor", "Velocity")

    def __init__(self, device=0, keepLocalS=False, recomputeRecovery=False):
        self.myMesh = self.fieldMap = self.linSystem = self.model = None
        self.boundaries = []
        self.initialized = self.allocated = self.assembled = False
        self.device = device
        self.keepLocalS = keepLocalS
        self.recomputeRecovery = recomputeRecovery   # HFX_RECOMPUTE_RECOVERY: U, Q are not stored, the recovery re-condenses each element
        self.ctx = None
        self.verbose = False
        self.stats = capi.SolveStats()
        self._meshUploaded = False
        self.myOpts = HDGSolverOpts()

    def setVerbosity(self, v): self.verbose = bool(v)
    def setOptions(self, opts):
        """HDGSolver::setOptions (HDGSolver.h:41).  WEXPLICIT / SEXPLICIT: the trace problem is explicit in the current Solution / Flux (HDGSolver.cpp:346-354)."""
        if opts.type not in (IMPLICIT, WEXPLICIT, SEXPLICIT):
            raise ErrorHandle("HDGSolver : setOptions : unknown solver type")
        self.myOpts = opts
        self.verbose = bool(opts.verbosity)
    def setMesh(self, m): self.myMesh = m; self._meshUploaded = False
    def setFieldMap(self, fm): self.fieldMap = fm
    def setLinSystem(self, lai): self.linSystem = lai
    def setModel(self, m): self.model = m

    def setBoundaryModel(self, bm):
        if self.myMesh is None:
            raise ErrorHandle("Solver : setBoundaryModel : must set the Mesh before the boundary model.")
        self.boundaries.append((bm, None))

    def setBoundaryCondition(self, bm, faces):
        self.boundaries.append((bm, np.array(sorted(faces), dtype=np.int32)))

    def initialize(self):
        if self.linSystem is not None:
            self.linSystem.destroySystem()
            self.linSystem.initialize()
            self.linSystem.configure()
        self.initialized = True

    def _h(self): return self.ctx.h

    def _upload_field(self, name, asynchronous=False):
        f = self.fieldMap[name]
        if f._deviceNewer and f._dev is self and f._devName == name:
            return                             # produced on the device and not looked at since: nothing to upload
        v = f64(f.values)
        f._dev, f._devName = self, name
        if asynchronous:   # the copy overlaps the assembly; the field's storage is not touched before hfx_assemble returns
            self._keep = getattr(self, "_keep", {}); self._keep[name] = v
            check(lib().hfx_field_set_async(self._h(), name.encode(), f.type, f.nObj, f.nVals, pd(v), int(f.doubleValued)), self._h())
        else:
            check(lib().hfx_field_set(self._h(), name.encode(), f.type, f.nObj, f.nVals, pd(v), int(f.doubleValued)), self._h())

    def allocate(self):
        if not self.initialized:
            raise ErrorHandle("HDGSolver : allocate : must initialize the solver before allocating.")
        if self.myMesh is None:
            raise ErrorHandle("HDGSolver : allocate : must set the Mesh before allocating.")
        if self.linSystem is None and self.myOpts.type != SEXPLICIT:      # HDGSolver.cpp:12-14
            raise ErrorHandle("HDGSolver : allocate : must set the linear system before allocating.")
        if self.model is None:
            raise ErrorHandle("HDGSolver : allocate : must set the model before allocating.")
        if not self.boundaries:
            raise ErrorHandle("HDGSolver : allocate : must set the boundary model before allocating.")
        if not self.fieldMap:
            raise ErrorHandle("HDGSolver : allocate : must set the fields before allocating.")
        fm, mesh, re = self.fieldMap, self.myMesh, self.myMesh.getReferenceElement()
        for nm, msg in (("Solution", "the field map must have a Solution field."), ("Flux", "the field map must have a Flux field."),
                        ("Tau", "the field map must have a Tau field."), ("Trace", "the field map must have a Trace field.")):
            if nm not in fm:
                raise ErrorHandle("HDGSolver : allocate : " + msg)
        sol = fm["Solution"]
        if sol.type != Cell:
            raise ErrorHandle("HDGSolver : allocate : the Solution field must be a cell field.")
        if sol.nObj != re.getNumNodes():
            raise ErrorHandle("HDGSolver : allocate : the Solution field must have an object per element node.")
        self.nDOFsPerNode = sol.nVals
        fl = fm["Flux"]
        if fl.type != Cell or fl.nObj != re.getNumNodes() or fl.nVals != self.nDOFsPerNode * mesh.getNodeSpaceDimension():
            raise ErrorHandle("HDGSolver : allocate : the Flux field must represent a spatial derivative of the Solution field.")
        if fm["Trace"].type != Face or fm["Trace"].nVals != self.nDOFsPerNode:
            raise ErrorHandle("HDGSolver : allocate : the Trace field must have the same number of values per object as the Solution field.")
        self.ctx = getattr(self.linSystem, "ctx", None) or Context(self.device)
        L, h = lib(), self._h()
        if not self._meshUploaded:
            check(L.hfx_refel_set(h, mesh.dim, mesh.order, re._geom), h)
            check(L.hfx_mesh_set(h, mesh.getNumberPoints(), pd(mesh.nodes), mesh.getNumberCells(), pi(mesh.cells)), h)
            self._meshUploaded = True
        self.model.allocate(self.nDOFsPerNode)
        for name in self._input_fields():
            self._upload_field(name)
        self._describe_model(strict=False)
        for bm, faces in self.boundaries:
            bm.allocate(self.nDOFsPerNode)
            check(L.hfx_boundary_describe(h, bm.kind, 0 if faces is None else faces.size, None if faces is None else pi(faces)), h)
        check(L.hfx_solver_type(h, self.myOpts.type), h)
        check(L.hfx_allocate(h, (1 if self.keepLocalS else 0) | (2 if self.recomputeRecovery else 0)), h)
        self.allocated = True

    def _input_fields(self):
        names = [n for n in self.INPUT_FIELDS if n in self.fieldMap]
        if not getattr(self.model, "usesDiffusionField", True) and "DiffusionTensor" in names:
            names.remove("DiffusionTensor")   # HDGLaplaceModel never reads it (HDGLaplaceModel.cpp:18-30)
        ts = self.model.timeScheme
        if ts is not None and getattr(ts, "isRK", False):
            if self.allocated:
                names += [n for n in ts.fieldNames() if n in self.fieldMap]
        elif ts is not None and self.allocated:
            names.append("Solution")
        if getattr(self.model, "isBurgers", False):
            names += [n for n in ("BufferSolution", "Trace") if n in self.fieldMap]
        if self.myOpts.type != IMPLICIT and self.allocated:
            names += [n for n in ("Solution", "Flux") if n not in names]
        return names

    def _describe_model(self, strict=True):
        names = set(self._input_fields())
        mask = self.model._mask(names, strict)
        tso = self.model.timeScheme
        ts = 0 if tso is None else (2 if getattr(tso, "isRK", False) else 1)
        md = capi.ModelDesc(self.nDOFsPerNode, mask, ts, tso.dt if ts else 0.0)
        check(lib().hfx_model_describe(self._h(), C.byref(md)), self._h())
        if ts == 2:
            if sorted(tso.auxiliaryFields) != ["Flux", "Trace"]:
                raise ErrorHandle("RungeKutta : apply : the stiffness matrix does not have the correct dimensions (the HDG path needs the auxiliary fields Flux and Trace)")
            if strict:
                for n in tso.fieldNames():
                    if n not in self.fieldMap:
                        raise ErrorHandle("RungeKutta : setFieldMap : the field map must provide the field " + n)
            row = tso.stageRow() if tso.getStage() < tso.getNumStages() else np.zeros(tso.getNumStages())
            check(lib().hfx_time_scheme_rk(self._h(), min(tso.getStage(), tso.getNumStages() - 1), tso.getNumStages(), pd(row)), self._h())
        self._mask = mask

    def _eval_callbacks(self):
        L, h, mesh = lib(), self._h(), self.myMesh
        if self._mask & (OP_SOURCE | OP_REACTION):
            nC, nIP, d = mesh.getNumberCells(), mesh.getReferenceElement().getNumIPs(), mesh.dim
            if getattr(self, "_xip", None) is None:
                self._xip = np.zeros((nC, nIP, d))
                check(L.hfx_ip_coords(h, pd(self._xip)), h)
            pts = self._xip.reshape(-1, d)
            if self._mask & OP_SOURCE:
                if getattr(self.model, "isBurgers", False):   # one scalar Source per component (HDGBurgersModel.cpp:51-56,112-122)
                    v = f64([[[self.model.sourceFunc(list(p), c) for p in el] for c in range(d)] for el in self._xip])
                    check(L.hfx_source_values_n(h, d, pd(v)), h)
                else:
                    v = f64([self.model.sourceFunc(list(p)) for p in pts])
                    check(L.hfx_source_values(h, pd(v)), h)
            if self._mask & OP_REACTION:
                v = f64([self.model.reactionFunc(list(p)) for p in pts])
                check(L.hfx_reaction_values(h, pd(v)), h)

    def assemble(self):
        if not (self.initialized and self.allocated):
            raise ErrorHandle("HDGSolver : assemble : the solver must be initialized and allocated before assembling.")
        for name in self._input_fields():
            self._upload_field(name, asynchronous=True)
        self._describe_model()
        self._eval_callbacks()
        check(lib().hfx_assemble(self._h()), self._h())
        self.assembled = True

    def solve(self):
        if not self.assembled:
            raise ErrorHandle("HDGSolver : solve : system must be assembled before solving")
        o = self.linSystem.opts.c() if hasattr(self.linSystem, "opts") else PetscOpts().c()
        check(lib().hfx_solver_type(self._h(), self.myOpts.type), self._h())
        check(lib().hfx_solve(self._h(), C.byref(o), C.byref(self.stats)), self._h())
        for name in ("Trace", "Solution", "Flux"):      # left on the device; Field.values fetches them when somebody looks
            f = self.fieldMap[name]
            f._values = f64(f._values)
            f._dev, f._devName, f._deviceNewer = self, name, True

    # ---- device field arithmetic (hfx_field_lincomb / hfx_field_diff_norm2) ---------------------------------------------------------------
    def _ensure_on_device(self, name):
        f = self.fieldMap[name]
        if not (f._on_device(self) and f._devName == name) or not f._deviceNewer:
            v = f64(f._values)
            check(lib().hfx_field_set(self._h(), name.encode(), f.type, f.nObj, f.nVals, pd(v), int(f.doubleValued)), self._h())
            f._dev, f._devName = self, name

    def fieldLinComb(self, dst, coefs, names):
        """fieldMap[dst] <- sum_k coefs[k] * fieldMap[names[k]] on the device; the result stays there until its `values` are read."""
        for n in names:
            self._ensure_on_device(n)
        arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
        cf = f64(np.asarray(coefs, dtype=np.float64))
        check(lib().hfx_field_lincomb(self._h(), dst.encode(), len(names), pd(cf), arr), self._h())
        f = self.fieldMap[dst]
        f._values = f64(f._values)
        f._dev, f._devName, f._deviceNewer = self, dst, True

    def fieldDiffNorm2(self, a, b):
        """(||a - b||^2, ||b||^2) of two fields of the field map, on the device (owned cells + all-reduce on a partitioned mesh)."""
        for n in (a, b):
            self._ensure_on_device(n)
        d2, r2 = C.c_double(0), C.c_double(0)
        check(lib().hfx_field_diff_norm2(self._h(), a.encode(), b.encode(), C.byref(d2), C.byref(r2)), self._h())
        return d2.value, r2.value

    # ---- parity hooks -------------------------------------------------------------------------------------------
    def getCSR(self, with_cols=True):
        L, h = lib(), self._h()
        n, nnz = C.c_longlong(0), C.c_longlong(0)
        check(L.hfx_get_csr(h, C.byref(n), C.byref(nnz), None, None, None, None), h)
        rowptr = np.zeros(n.value + 1, dtype=np.int64)
        col = np.zeros(nnz.value, dtype=np.int32) if with_cols else None
        vals = np.zeros(nnz.value); rhs = np.zeros(n.value)
        check(L.hfx_get_csr(h, None, None, rowptr.ctypes.data_as(capi.lp), pi(col), pd(vals), pd(rhs)), h)
        return rowptr, col, vals, rhs

    def getLocal(self, iEl=0, nEl=None):
        mesh, re = self.myMesh, self.myMesh.getReferenceElement()
        nEl = mesh.getNumberCells() - iEl if nEl is None else nEl
        nD = self.nDOFsPerNode
        u = re.getNumNodes() * nD; q = u * mesh.dim; l = re.getNumFaces() * re.getFaceElement().getNumNodes() * nD
        out = dict(U=np.zeros((nEl, u * l)), U0=np.zeros((nEl, u)), Q=np.zeros((nEl, q * l)), Q0=np.zeros((nEl, q)))
        if self.keepLocalS:
            out.update(S=np.zeros((nEl, l * l)), S0=np.zeros((nEl, l)))
        check(lib().hfx_get_local(self._h(), iEl, nEl, pd(out.get("S")), pd(out.get("S0")), pd(out["U"]), pd(out["U0"]), pd(out["Q"]), pd(out["Q0"])), self._h())
        return out

    def getLocalMatrix(self, iEl):
        """(A [n, n], F [n]): the dense local system of element iEl as Model::compute leaves it (operators + time scheme, before the condensation),
        from the device (hfx_get_local_matrix).  Unknown order [u | q | lambda] as in the reference."""
        if not (self.initialized and self.allocated):
            raise ErrorHandle("HDGSolver : assemble : the solver must be initialized and allocated before assembling.")
        for name in self._input_fields():
            self._upload_field(name)
        self._describe_model()
        self._eval_callbacks()
        re, nD = self.myMesh.getReferenceElement(), self.nDOFsPerNode
        n = re.getNumNodes() * nD * (1 + self.myMesh.dim) + re.getNumFaces() * re.getFaceElement().getNumNodes() * nD
        A = np.zeros((n, n)); F = np.zeros(n)
        check(lib().hfx_get_local_matrix(self._h(), int(iEl), pd(A), pd(F)), self._h())
        return np.ascontiguousarray(A.T), F          # the library writes column-major

    def getElemDofs(self):
        mesh, re = self.myMesh, self.myMesh.getReferenceElement()
        l = re.getNumFaces() * re.getFaceElement().getNumNodes() * self.nDOFsPerNode
        out = np.zeros((mesh.getNumberCells(), l), dtype=np.int32)
        check(lib().hfx_get_elem_dofs(self._h(), 0, mesh.getNumberCells(), pi(out)), self._h())
        return out

    def lastAssembleMs(self):
        a, b = C.c_float(0), C.c_float(0)
        lib().hfx_last_assemble_ms(self._h(), C.byref(a), C.byref(b))
        return a.value, b.value

    def lastAssembleKernel(self):
        """Which device kernel served the last assemble: "fused" (hfx_assemble.cuh), "general" (hfx_generic.cuh), "big" (hfx_big.cuh), "p1" (hfx_p1.cuh) or "col" (hfx_col.cuh)."""
        k, pf = C.c_int(0), C.c_int(0)
        lib().hfx_last_assemble_kernel(self._h(), C.byref(k), C.byref(pf))
        return ("fused", "general", "big", "p1", "col")[k.value]


class LaplaceModel:
    """LaplaceModel (src/model/LaplaceModel.cpp:15-52): localMatrix = Diffusion (DiffusionTensor field if given), no right-hand side; assembly = {Add, None}."""
    isCG = True
    usesDiffusionField = True

    def __init__(self, refEl):
        self.refEl, self.allocated, self.sourceFunc, self.timeScheme = refEl, False, None, None

    def allocate(self, nDOFsPerNode):
        self.nDOF, self.allocated = nDOFsPerNode, True

    def setTimeScheme(self, ts):
        raise ErrorHandle("LaplaceModel : setTimeScheme : the Laplace model is stationary")

    def _mask(self, fieldNames=()):
        return OP_DIFFUSION


class DiffusionSource(LaplaceModel):
    """DiffusionSource (src/model/DiffusionSource.cpp): Diffusion + Source right-hand side, optionally under an implicit Euler step (FEModel::compute, FEModel.cpp:22-33)."""

    def setSourceFunction(self, s):
        if not self.allocated:
            raise ErrorHandle("DiffusionSource : setSourceFunction : the model must be allocated before setting the source function")
        self.sourceFunc = s

    def setTimeScheme(self, ts):
        if self.allocated:
            raise ErrorHandle("FEModel : setTimeScheme : the time scheme must be set before allocation or field setting")
        if getattr(ts, "isRK", False) or getattr(ts, "isExplicit", False):
            raise ErrorHandle("DiffusionSource : setTimeScheme : the device CG path serves the implicit Euler scheme")
        self.timeScheme = ts

    def _mask(self, fieldNames=()):
        if self.sourceFunc is None:
            raise ErrorHandle("Source : calcSource : must set a source function before calculating the source.")
        return OP_DIFFUSION | OP_SOURCE


class Transport(DiffusionSource):
    """Transport (src/model/Transport.cpp): localMatrix = Convection (Velocity node field), zero right-hand side, optionally under an implicit Euler step."""
    usesDiffusionField = False

    def setSourceFunction(self, s):
        raise ErrorHandle("Transport : setSourceFunction : the transport model has no source")

    def _mask(self, fieldNames=()):
        if "Velocity" not in fieldNames:
            raise ErrorHandle("Transport : setFieldMap : one must provide a Velocity field to use the Transport model.")
        return OP_CONVECTION


class CGSolver:
    """CGSolver (src/solver/CGSolver.cpp) on the device (hfx_cg_*): node-based CSR, element loop, DirichletModel rows, Krylov solve into the nodal Solution field.
    Same call-order contract as the reference (tests/unittests/solver/TestCGSolver.cpp:37-66)."""

    def __init__(self, device=0):
        self.myMesh = self.fieldMap = self.linSystem = self.model = None
        self.boundaries = []
        self.initialized = self.allocated = self.assembled = False
        self.device, self.ctx, self.verbose = device, None, False
        self.stats = capi.SolveStats()

    def setVerbosity(self, v): self.verbose = bool(v)
    def setMesh(self, m): self.myMesh = m
    def setFieldMap(self, fm): self.fieldMap = fm
    def setLinSystem(self, lai): self.linSystem = lai
    def setModel(self, m): self.model = m

    def setBoundaryModel(self, bm):
        if self.myMesh is None:
            raise ErrorHandle("Solver : setBoundaryModel : must set the Mesh before the boundary model.")
        self.boundaries.append((bm, None))

    def setBoundaryCondition(self, bm, faces):
        self.boundaries.append((bm, np.array(sorted(faces), dtype=np.int32)))

    def initialize(self):
        if self.linSystem is not None:
            self.linSystem.destroySystem(); self.linSystem.initialize(); self.linSystem.configure()
        self.initialized = True

    def _h(self): return self.ctx.h

    def allocate(self):
        if not self.initialized:
            raise ErrorHandle("CGSolver : allocate : must initialize the solver before allocating.")
        if self.myMesh is None:
            raise ErrorHandle("CGSolver : allocate : must set the Mesh before allocating.")
        if self.linSystem is None:
            raise ErrorHandle("CGSolver : allocate : must set the linear system before allocating.")
        if self.model is None:
            raise ErrorHandle("CGSolver : allocate : must set the model before allocating.")
        if not self.boundaries:
            raise ErrorHandle("CGSolver : allocate : must set the boundary model before allocating.")
        if not self.fieldMap:
            raise ErrorHandle("CGSolver : allocate : must set the fields before allocating.")
        if "Solution" not in self.fieldMap:
            raise ErrorHandle("CGSolver : allocate : the field map must have a Solution field.")
        sol, mesh, re = self.fieldMap["Solution"], self.myMesh, self.myMesh.getReferenceElement()
        if sol.type != Node:
            raise ErrorHandle("CGSolver : allocate : the Solution field must be a nodal field.")
        self.nDOFsPerNode = sol.nObj * sol.nVals
        self.ctx = getattr(self.linSystem, "ctx", None) or Context(self.device)
        L, h = lib(), self._h()
        check(L.hfx_refel_set(h, mesh.dim, mesh.order, re._geom), h)
        check(L.hfx_mesh_set(h, mesh.getNumberPoints(), pd(mesh.nodes), mesh.getNumberCells(), pi(mesh.cells)), h)
        self.model.allocate(self.nDOFsPerNode)
        self._upload("Solution")
        md = capi.ModelDesc(self.nDOFsPerNode, OP_DIFFUSION, 0, 0.0)
        check(L.hfx_model_describe(h, C.byref(md)), h)
        for bm, faces in self.boundaries:
            bm.allocate(self.nDOFsPerNode)
            if getattr(bm, "kind", None) != 0:
                raise ErrorHandle("CGSolver : allocate : the device CG path serves DirichletModel boundaries")
            check(L.hfx_boundary_describe(h, 0, 0 if faces is None else faces.size, None if faces is None else pi(faces)), h)
        check(L.hfx_cg_allocate(h), h)
        self.allocated = True

    def _upload(self, name):
        f = self.fieldMap[name]
        v = f64(f.values)
        check(lib().hfx_field_set(self._h(), name.encode(), f.type, f.nObj, f.nVals, pd(v), int(f.doubleValued)), self._h())

    def assemble(self):
        if not (self.initialized and self.allocated):
            raise ErrorHandle("CGSolver : assemble : must initialize and allocate the solver before allocating.")
        L, h, mesh = lib(), self._h(), self.myMesh
        names = ["Dirichlet"] + (["DiffusionTensor"] if self.model.usesDiffusionField else []) + ["Velocity"]
        ts = self.model.timeScheme
        if ts is not None:
            names.append("Solution")            # the old state of the Euler step (Euler.cpp:31)
        for name in names:
            if name in self.fieldMap:
                self._upload(name)
        if "Dirichlet" not in self.fieldMap:
            raise ErrorHandle("DirichletModel : setFieldMap : must give a field named Dirichlet to the DirichletModel")
        mask = self.model._mask(set(self.fieldMap))
        md = capi.ModelDesc(self.nDOFsPerNode, mask, 1 if ts is not None else 0, ts.dt if ts is not None else 0.0)
        check(L.hfx_model_describe(h, C.byref(md)), h)
        if mask & OP_SOURCE:
            nC, nIP, d = mesh.getNumberCells(), mesh.getReferenceElement().getNumIPs(), mesh.dim
            xip = np.zeros((nC, nIP, d))
            check(L.hfx_ip_coords(h, pd(xip)), h)
            v = f64([self.model.sourceFunc(list(p)) for p in xip.reshape(-1, d)])
            check(L.hfx_source_values(h, pd(v)), h)
        check(L.hfx_cg_assemble(h), h)
        self.assembled = True

    def solve(self):
        if not self.assembled:
            raise ErrorHandle("CGSolver : solve : system must be assembled before solving")
        o = self.linSystem.opts.c() if hasattr(self.linSystem, "opts") else PetscOpts().c()
        check(lib().hfx_cg_solve(self._h(), C.byref(o), C.byref(self.stats)), self._h())
        f = self.fieldMap["Solution"]
        f._deviceNewer = False
        check(lib().hfx_field_get(self._h(), b"Solution", pd(f._values)), self._h())

    def getCSR(self):
        n, nnz = C.c_longlong(0), C.c_longlong(0)
        check(lib().hfx_cg_get_csr(self._h(), C.byref(n), C.byref(nnz), None, None, None, None), self._h())
        rowptr, col, vals, rhs = np.zeros(n.value + 1, dtype=np.int64), np.zeros(nnz.value, dtype=np.int32), np.zeros(nnz.value), np.zeros(n.value)
        check(lib().hfx_cg_get_csr(self._h(), C.byref(n), C.byref(nnz), rowptr.ctypes.data_as(C.POINTER(C.c_longlong)), pi(col), pd(vals), pd(rhs)), self._h())
        return rowptr, col, vals, rhs


class NonLinearWrapper:
    """Fixed-point / Newton driver with damping (src/solver/NonLinearWrapper.cpp:41-79)."""

    def __init__(self):
        self.mySolver = self.currentSolution = self.previousSolution = None
        self.resTol, self.maxIters, self.dampening, self.residual, self.verbose = 1e-6, 1000, 0.0, 0.0, False
        self.linearizedSolver, self.residualComputer = self.vanillaLinearizedSolver, self.vanillaResidualComputer

    @staticmethod
    def vanillaLinearizedSolver(solver):          # NonLinearWrapper.cpp:36-39
        solver.assemble()
        solver.solve()

    @staticmethod
    def vanillaResidualComputer(cur, prev):       # NonLinearWrapper.cpp:12-34: ||cur - prev|| / ||prev||, the two sums all-reduced over the ranks
        s = getattr(cur, "_dev", None)
        if s is not None and not os.environ.get("HFX_HOST_FIELD_ARITHMETIC"):
            names = {id(f): n for n, f in s.fieldMap.items()}
            if id(cur) in names and id(prev) in names:      # on the device: one fused pass, owned cells + ncclAllReduce on a partitioned mesh
                diff, ref = s.fieldDiffNorm2(names[id(cur)], names[id(prev)])
                return np.sqrt(diff / ref) if ref != 0 else np.sqrt(diff)
        diff, ref = float(np.sum((cur.values - prev.values) ** 2)), float(np.sum(prev.values ** 2))
        return np.sqrt(diff / ref) if ref != 0 else np.sqrt(diff)

    def setLinearizedSolver(self, f): self.linearizedSolver = f
    def setResidualComputer(self, f): self.residualComputer = f

    def setSolver(self, s): self.mySolver = s
    def setSolutionFields(self, cur, prev): self.currentSolution, self.previousSolution = cur, prev
    def setMaxIterations(self, n): self.maxIters = n
    def setResidualTolerance(self, t): self.resTol = t
    def setDampening(self, d): self.dampening = d
    def setVerbosity(self, v): self.verbose = v
    def getResidual(self): return self.residual

    def solve(self):
        if self.mySolver is None:
            raise ErrorHandle("NonLinearWrapper : solve : the Solver must be set before attempting to solve")
        if self.currentSolution is None:
            raise ErrorHandle("NonLinearWrapper : solve : the current and previous Solutions should be set before attempting to solve")
        s = self.mySolver
        names = {id(f): n for n, f in getattr(s, "fieldMap", {}).items()} if hasattr(s, "fieldLinComb") else {}
        onDevice = id(self.currentSolution) in names and id(self.previousSolution) in names and not os.environ.get("HFX_HOST_FIELD_ARITHMETIC")
        for _ in range(self.maxIters):
            self.linearizedSolver(self.mySolver)
            self.residual = self.residualComputer(self.currentSolution, self.previousSolution)
            if self.residual < self.resTol:
                break
            if onDevice:      # NonLinearWrapper.cpp:55-70 as device AXPYs
                c, p_ = names[id(self.currentSolution)], names[id(self.previousSolution)]
                s.fieldLinComb(c, [1.0 - self.dampening, self.dampening], [c, p_])
                s.fieldLinComb(p_, [1.0], [c])
                continue
            cur, prev = self.currentSolution.values, self.previousSolution.values
            val = (1.0 - self.dampening) * cur + self.dampening * prev
            cur[:] = val
            prev[:] = val
