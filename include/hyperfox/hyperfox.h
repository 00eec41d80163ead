/* hyperfox.h -- C++ host mirror of the reference's class surface for the element-by-element HDG path, over the C ABI
 * of libhfx.so (include/hfx.h).  Header-only; link with -lhfx.
 *
 * Same class names, method names, argument meaning and error behaviour (hfox::ErrorHandle, "Class : function : message") as
 * the reference, so code written against the reference's headers compiles against these after swapping the include path
 * and PetscInterface -> CudaLinAlgebraInterface:
 *
 *   ErrorHandle                <- src/globals/ErrorHandle.h             ReferenceElement <- src/element/ReferenceElement.h
 *   Mesh                       <- src/mesh/Mesh.h                       Field, FieldType <- src/field/Field.h, FieldTypes.h
 *   PetscOpts                  <- src/resolution/PetscOpts.h            LinAlgebraInterface <- src/resolution/LinAlgebraInterface.h:23-162
 *   CudaLinAlgebraInterface    (replaces PetscInterface, src/resolution/PetscInterface.cpp)
 *   TimeScheme, Euler, RungeKutta  <- src/operator/{TimeScheme,Euler,RungeKutta}.h
 *   FEModel, HDGModel, HDGLaplaceModel, HDGDiffusionSource, HDGConvectionDiffusionReactionSource, HDGBurgersModel,
 *   BoundaryModel, DirichletModel, IntegratedDirichletModel  <- src/model/*.h
 *   Solver, HDGSolver, HDGSolverOpts  <- src/solver/{Solver,HDGSolver,HDGSolverOpts}.h
 *   NonLinearWrapper           <- src/solver/NonLinearWrapper.h
 *
 * What differs by design: models are operator *descriptors* (the per-element virtual compute() of the reference cannot run on
 * the device; std::function source/reaction callbacks are evaluated by the host at x(IP) and uploaded); Fields stay host-side
 * std::vector<double> (as in the reference) and are copied to / from HBM by HDGSolver::assemble / solve.  There is no CPU
 * fallback: without a CUDA device every compute call throws ErrorHandle.
 */
#ifndef HYPERFOX_B200_HYPERFOX_H
#define HYPERFOX_B200_HYPERFOX_H

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <exception>
#include <functional>
#include <iostream>
#include <map>
#include <set>
#include <string>
#include <tuple>
#include <vector>

#include "../hfx.h"

namespace hfox {

// ---- src/globals/ErrorHandle.h ------------------------------------------------------------------------------------------
class ErrorHandle : public std::exception {
 public:
  ErrorHandle() {}
  explicit ErrorHandle(const std::string& full) : message(full) {}
  ErrorHandle(const std::string& cls, const std::string& fn, const std::string& msg) : message(cls + " : " + fn + " : " + msg) {}
  const char* what() const noexcept override { return message.c_str(); }

 protected:
  std::string message;
};

namespace detail {
inline void check(int rc, const hfx_ctx* c) { if (rc != 0) { const char* m = hfx_last_error(c); throw ErrorHandle(m && *m ? m : "hfx : call : unknown error"); } }
inline void lcheck(int rc, const hfx_lai* l) { if (rc != 0) { const char* m = hfx_lai_last_error(l); throw ErrorHandle(m && *m ? m : "hfx : call : unknown error"); } }
// one device context shared by the solver and its linear system
class Context {
 public:
  explicit Context(int device = 0) { detail::check(hfx_ctx_create(device, &h), nullptr); }
  ~Context() { if (h) hfx_ctx_destroy(h); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  hfx_ctx* h = nullptr;
};
}  // namespace detail

// ---- src/field/FieldTypes.h, src/model/AssemblyType.h ------------------------------------------------------------------
enum FieldType : int { None = -1, Node = 0, Edge = 1, Face = 2, Cell = 3 };
enum UnitAssemblyType { NoAssembly, Add, Set };
struct AssemblyType { UnitAssemblyType matrix; UnitAssemblyType rhs; };

// ---- src/element/ReferenceElement.h ------------------------------------------------------------------------------------
enum elementGeometry { simplex, orthotope };

class ReferenceElement {
 public:
  ReferenceElement(int dim, int ord, std::string geom) : dimension(dim), order(ord) {
    if (geom == "simplex") geometry = simplex;
    else if (geom == "orthotope") geometry = orthotope;
    else throw ErrorHandle("ReferenceElement", "setGeometry", "Element type " + geom + " is not yet supported.");
    geomName = geom;
    if (dim < 0) throw ErrorHandle("ReferenceElement", "setDim", "Dimension must be positive.");
    if (dim == 0) { nNodes = 1; nFaces = 0; nodes.assign(1, std::vector<double>()); ipShapeFunctions.assign(1, std::vector<double>(1, 1.0)); ipWeights.assign(1, 1.0); ipCoords.assign(1, std::vector<double>()); return; }
    int sz[5] = {0, 0, 0, 0, 0};
    detail::check(hfx_refel_host_tables(dim, ord, geometry == simplex ? HFX_SIMPLEX : HFX_ORTHOTOPE, sz, 0, 0, 0, 0, 0, 0, 0, 0, 0), nullptr);
    const int nN = sz[0], nNf = sz[1], nFc = sz[2], nIP = sz[3], nIPf = sz[4];
    std::vector<double> nd((size_t)nN * dim), ipc((size_t)nIP * dim), w(nIP), sh((size_t)nIP * nN), dsh((size_t)nIP * nN * dim), fsh((size_t)nIPf * nNf),
        fdsh((size_t)nIPf * nNf * std::max(dim - 1, 1)), fw(nIPf);
    std::vector<int> fn((size_t)nFc * nNf);
    detail::check(hfx_refel_host_tables(dim, ord, geometry == simplex ? HFX_SIMPLEX : HFX_ORTHOTOPE, sz, nd.data(), ipc.data(), w.data(), sh.data(), dsh.data(),
                                        fsh.data(), fdsh.data(), fw.data(), fn.data()), nullptr);
    nNodes = nN; nFaces = nFc;
    nodes.assign(nN, std::vector<double>(dim));
    for (int i = 0; i < nN; i++) for (int d = 0; d < dim; d++) nodes[i][d] = nd[(size_t)i * dim + d];
    ipCoords.assign(nIP, std::vector<double>(dim));
    for (int i = 0; i < nIP; i++) for (int d = 0; d < dim; d++) ipCoords[i][d] = ipc[(size_t)i * dim + d];
    ipWeights = w;
    ipShapeFunctions.assign(nIP, std::vector<double>(nN));
    ipDerivShapeFunctions.assign(nIP, std::vector<std::vector<double> >(nN, std::vector<double>(dim)));
    for (int ip = 0; ip < nIP; ip++) for (int i = 0; i < nN; i++) {
      ipShapeFunctions[ip][i] = sh[(size_t)ip * nN + i];
      for (int d = 0; d < dim; d++) ipDerivShapeFunctions[ip][i][d] = dsh[((size_t)ip * nN + i) * dim + d];
    }
    faceNodes.assign(nFc, std::vector<int>(nNf));
    for (int f = 0; f < nFc; f++) for (int j = 0; j < nNf; j++) faceNodes[f][j] = fn[(size_t)f * nNf + j];
    faceElement = new ReferenceElement(dim - 1, ord, geom);
  }
  ~ReferenceElement() { delete faceElement; }
  ReferenceElement(const ReferenceElement&) = delete;
  ReferenceElement& operator=(const ReferenceElement&) = delete;
  int getDimension() const { return dimension; }
  int getOrder() const { return order; }
  elementGeometry getGeometry() const { return geometry; }
  const std::string& getGeometryName() const { return geomName; }
  int getNumNodes() const { return nNodes; }
  int getNumIPs() const { return (int)ipWeights.size(); }
  int getNumFaces() const { return nFaces; }
  const std::vector<std::vector<double> >* getNodes() const { return &nodes; }
  const std::vector<std::vector<int> >* getFaceNodes() const { return &faceNodes; }
  const std::vector<std::vector<double> >* getIPCoords() const { return &ipCoords; }
  const std::vector<double>* getIPWeights() const { return &ipWeights; }
  const std::vector<std::vector<double> >* getIPShapeFunctions() const { return &ipShapeFunctions; }
  const std::vector<std::vector<std::vector<double> > >* getIPDerivShapeFunctions() const { return &ipDerivShapeFunctions; }
  const ReferenceElement* getFaceElement() const { return faceElement; }

 protected:
  int dimension, order, nNodes = 0, nFaces = 0;
  elementGeometry geometry = simplex;
  std::string geomName;
  std::vector<std::vector<double> > nodes, ipCoords, ipShapeFunctions;
  std::vector<double> ipWeights;
  std::vector<std::vector<std::vector<double> > > ipDerivShapeFunctions;
  std::vector<std::vector<int> > faceNodes;
  ReferenceElement* faceElement = nullptr;
};

// ---- src/mesh/Mesh.h -----------------------------------------------------------------------------------------------------
class Partitioner;
class Mesh {
 public:
  Mesh() {}
  Mesh(int dim, int order, std::string geom) { setReferenceElement(dim, order, geom); }
  Mesh(int dim, int order, std::string geom, int dimPointSpace, std::vector<double>& points, std::vector<int>& connectivity) {
    setReferenceElement(dim, order, geom);
    setMesh(dimPointSpace, points, connectivity);
  }
  ~Mesh() { delete refElement; }
  Mesh(const Mesh&) = delete;
  Mesh& operator=(const Mesh&) = delete;
  void setReferenceElement(int dim, int order, std::string geom) {
    delete refElement;
    refElement = new ReferenceElement(dim, order, geom);
    nNodesPerCell = refElement->getNumNodes();
    nNodesPerFace = refElement->getFaceElement()->getNumNodes();
    nFacesPerCell = refElement->getNumFaces();
  }
  // src/mesh/Mesh.cpp:31-46: copies the candidates, then computeFaces()
  void setMesh(int dimPointSpace, std::vector<double>& points_candidate, std::vector<int>& connectivity_candidate) {
    if (!refElement) throw ErrorHandle("Mesh", "setMesh", "the reference element must be set before the mesh");
    if (dimPointSpace <= 0 || points_candidate.size() % dimPointSpace != 0) throw ErrorHandle("Mesh", "setMesh", "the points are not consistent with the dimension of the point space");
    if (connectivity_candidate.size() % nNodesPerCell != 0) throw ErrorHandle("Mesh", "setMesh", "the connectivity is not consistent with the reference element");
    dimNodeSpace = dimPointSpace;
    nodes = points_candidate; cells = connectivity_candidate;
    nNodes = (int)nodes.size() / dimNodeSpace; nCells = (int)cells.size() / nNodesPerCell;
    computeFaces();
  }
  void setPartitioner(Partitioner* p) { part = p; }
  Partitioner* getPartitioner() { return part; }
  const std::vector<double>* getPoints() const { return &nodes; }
  const std::vector<int>* getCells() const { return &cells; }
  const std::vector<int>* getFaces() const { return &faces; }
  void getPoint(int i, std::vector<double>* p) const { p->assign(nodes.begin() + (size_t)i * dimNodeSpace, nodes.begin() + (size_t)(i + 1) * dimNodeSpace); }
  void getCell(int i, std::vector<int>* c) const { c->assign(cells.begin() + (size_t)i * nNodesPerCell, cells.begin() + (size_t)(i + 1) * nNodesPerCell); }
  void getFace(int i, std::vector<int>* f) const { f->assign(faces.begin() + (size_t)i * nNodesPerFace, faces.begin() + (size_t)(i + 1) * nNodesPerFace); }
  void getSlicePoints(const std::vector<int>& slice, std::vector<std::vector<double> >* pts) const {
    pts->resize(slice.size());
    for (size_t k = 0; k < slice.size(); k++) getPoint(slice[k], &(*pts)[k]);
  }
  void getSliceCells(const std::vector<int>& slice, std::vector<std::vector<int> >* out) const {
    out->resize(slice.size());
    for (size_t k = 0; k < slice.size(); k++) getCell(slice[k], &(*out)[k]);
  }
  void getSliceFaces(const std::vector<int>& slice, std::vector<std::vector<int> >* out) const {
    out->resize(slice.size());
    for (size_t k = 0; k < slice.size(); k++) getFace(slice[k], &(*out)[k]);
  }
  const ReferenceElement* getReferenceElement() const { return refElement; }
  int getNumberPoints() const { return nNodes; }
  int getNumberCells() const { return nCells; }
  int getNumberFaces() const { return nFaces; }
  int getDimension() const { return refElement->getDimension(); }
  int getNodeSpaceDimension() const { return dimNodeSpace; }
  int getNumFacesPerCell() const { return nFacesPerCell; }
  const std::vector<int>* getCell2FaceMap() const { return &cell2FaceMap; }
  void getCell2Face(int i, std::vector<int>* c2f) const { c2f->assign(cell2FaceMap.begin() + (size_t)i * nFacesPerCell, cell2FaceMap.begin() + (size_t)(i + 1) * nFacesPerCell); }
  const std::vector<int>* getFace2CellMap() const { return &face2CellMap; }
  void getFace2Cell(int i, std::vector<int>* f2c) const {   // src/mesh/Mesh.cpp: boundary faces list one cell only
    f2c->clear();
    for (int k = 0; k < 2; k++) if (face2CellMap[(size_t)2 * i + k] >= 0) f2c->push_back(face2CellMap[(size_t)2 * i + k]);
  }
  std::set<int>* getBoundaryFaces() { return &boundaryFaces; }

 protected:
  // src/mesh/Mesh.cpp:183-274,377-537 (MOAB) -> the product's host topology builder
  void computeFaces() {
    const int dim = refElement->getDimension(), order = refElement->getOrder(), geom = refElement->getGeometry() == simplex ? HFX_SIMPLEX : HFX_ORTHOTOPE;
    int nF = 0, nB = 0;
    detail::check(hfx_host_compute_faces(dim, order, geom, nCells, cells.data(), &nF, 0, 0, 0, &nB, 0), nullptr);
    faces.assign((size_t)nF * nNodesPerFace, 0); cell2FaceMap.assign((size_t)nCells * nFacesPerCell, 0); face2CellMap.assign((size_t)nF * 2, -1);
    std::vector<int> b(nB);
    detail::check(hfx_host_compute_faces(dim, order, geom, nCells, cells.data(), &nF, faces.data(), cell2FaceMap.data(), face2CellMap.data(), &nB, b.data()), nullptr);
    nFaces = nF;
    boundaryFaces = std::set<int>(b.begin(), b.end());
  }
  std::vector<double> nodes;
  std::vector<int> cells, faces, face2CellMap, cell2FaceMap;
  int dimNodeSpace = 0, nNodes = 0, nCells = 0, nFaces = 0, nNodesPerCell = 0, nNodesPerFace = 0, nFacesPerCell = 0;
  ReferenceElement* refElement = nullptr;
  Partitioner* part = nullptr;
  std::set<int> boundaryFaces;
};

// ---- src/io/Io.h + tools/convertGmsh2H5HO.cpp ------------------------------------------------------------------------------
// The reference's Io interface (load / write / setMesh / setField) for the one input format the path's callers start from: a Gmsh
// 2.2 file of linear simplices, raised to the order of the mesh's reference element exactly as tools/convertGmsh2H5HO does
// (generateHigherOrderMesh, :117-257) -- same node numbering, without MOAB (host C++ inside libhfx).
class Field;
class Io {
 public:
  virtual ~Io() {}
  virtual void load(std::string filename) = 0;
  virtual void write(std::string filename) = 0;
  virtual void setMesh(Mesh* mesh) { myMesh = mesh; }
  virtual void setField(std::string name, Field* field) { fieldMap[name] = field; }

 protected:
  static std::string getExtension(std::string filename) {
    const size_t p = filename.find_last_of('.');
    return p == std::string::npos ? std::string() : filename.substr(p);
  }
  Mesh* myMesh = nullptr;
  std::map<std::string, Field*> fieldMap;
};

// src/io/HDF5Io.h without libhdf5 (hfx_host_read_h5_mesh / hfx_host_read_h5_field / hfx_host_write_h5): load = Mesh group (HDF5Io.cpp:111-152) + the fields registered
// with setField (:154-187); write = Mesh group (:189-302) + FieldData group (:304-391).  Bodies after the Field class below.
class HDF5Io : public Io {
 public:
  HDF5Io() {}
  explicit HDF5Io(Mesh* mesh) { setMesh(mesh); }
  void load(std::string filename) override;
  void write(std::string filename) override;
};

class GmshIo : public Io {
 public:
  GmshIo() {}
  explicit GmshIo(Mesh* mesh) { setMesh(mesh); }
  void load(std::string filename) override {
    if (!myMesh) throw ErrorHandle("GmshIo", "load", "the mesh must be set before loading");
    if (!myMesh->getReferenceElement()) throw ErrorHandle("GmshIo", "load", "the mesh needs a reference element (dimension and order) before loading");
    if (getExtension(filename) != ".msh") throw ErrorHandle("GmshIo", "load", "the file extension must be .msh");
    if (myMesh->getReferenceElement()->getGeometry() != simplex) throw ErrorHandle("GmshIo", "load", "only simplex meshes can be generated from a Gmsh file");
    const int dim = myMesh->getDimension(), order = myMesh->getReferenceElement()->getOrder();
    int nLin = 0, counts[4] = {0, 0, 0, 0};
    detail::check(hfx_host_read_msh(filename.c_str(), &nLin, counts, 0, 0, 0, 0), nullptr);
    std::vector<double> raw((size_t)nLin * 3), lin((size_t)nLin * dim);
    std::vector<int> el[4];
    for (int k = 1; k <= 3; k++) el[k].resize((size_t)counts[k] * (k + 1));
    detail::check(hfx_host_read_msh(filename.c_str(), &nLin, counts, raw.data(), el[1].data(), el[2].data(), el[3].data()), nullptr);
    if (counts[dim] == 0) throw ErrorHandle("GmshIo", "load", "the file holds no cells of the dimension of the mesh");
    for (int i = 0; i < nLin; i++) for (int d = 0; d < dim; d++) lin[(size_t)i * dim + d] = raw[(size_t)i * 3 + d];
    const int nEx2 = dim == 3 ? counts[2] : 0;
    int nNodes = 0;
    detail::check(hfx_host_high_order_mesh(dim, order, nLin, lin.data(), counts[dim], el[dim].data(), counts[1], el[1].data(), nEx2, el[2].data(), &nNodes, 0, 0), nullptr);
    std::vector<double> pts((size_t)nNodes * dim);
    std::vector<int> cells((size_t)counts[dim] * myMesh->getReferenceElement()->getNumNodes());
    detail::check(hfx_host_high_order_mesh(dim, order, nLin, lin.data(), counts[dim], el[dim].data(), counts[1], el[1].data(), nEx2, el[2].data(), &nNodes, pts.data(), cells.data()), nullptr);
    myMesh->setMesh(dim, pts, cells);
  }
  void write(std::string) override { throw ErrorHandle("GmshIo", "write", "writing Gmsh files is not supported"); }
};

// ---- src/field/Field.h ---------------------------------------------------------------------------------------------------
class Field {
 public:
  Field() {}
  explicit Field(Mesh* mesh) : pmesh(mesh) {}
  Field(Mesh* mesh, FieldType t, int nObjPerEnt, int nValPerObj) : pmesh(mesh), type(t), numObjPerEnt(nObjPerEnt), numValsPerObj(nValPerObj) { allocate(); }
  virtual ~Field() {}
  Mesh* getMesh() const { return pmesh; }
  void setMesh(Mesh* m) { pmesh = m; }
  int getLength() const { return (int)values.size(); }
  void computeNumEntities() {   // src/field/Field.cpp:25-39
    switch (type) {
      case Node: numEntities = pmesh->getNumberPoints(); break;
      case Face: numEntities = pmesh->getNumberFaces(); break;
      case Cell: numEntities = pmesh->getNumberCells(); break;
      default: throw ErrorHandle("Field", "computeNumEntities", "the field type is not supported");
    }
  }
  void allocate() { computeNumEntities(); values.assign((size_t)numEntities * numObjPerEnt * numValsPerObj, 0.0); deviceNewer = false; }
  // Results the device produces (HDGSolver::solve, RungeKutta::computeStage / computeSolution) stay there until somebody looks: getValues() brings them back, and from then
  // on the host vector is authoritative again (it may be modified in place), so the next assemble uploads it.
  std::vector<double>* getValues() { syncFromDevice(); return &values; }
  void syncFromDevice() {
    if (!deviceNewer) return;
    deviceNewer = false;
    detail::check(hfx_field_get(devCtx, devName.c_str(), values.data()), devCtx);
  }
  void markOnDevice(hfx_ctx* c, const std::string& name, bool newer) { devCtx = c; devName = name; deviceNewer = newer; }
  bool isDeviceNewer(const hfx_ctx* c, const std::string& name) const { return deviceNewer && devCtx == c && devName == name; }
  hfx_ctx* deviceContext() const { return devCtx; }
  const std::string& deviceName() const { return devName; }
  std::vector<double>* hostValues() { return &values; }   // no synchronisation: for code that is about to overwrite or upload
  FieldType* getFieldType() { return &type; }
  int* getNumEntities() { return &numEntities; }
  int* getNumObjPerEnt() { return &numObjPerEnt; }
  int* getNumValsPerObj() { return &numValsPerObj; }
  void getValues(int i, std::vector<double>* vals) {
    syncFromDevice();
    const size_t n = (size_t)numObjPerEnt * numValsPerObj;
    vals->assign(values.begin() + i * n, values.begin() + (i + 1) * n);
  }
  void getSliceValues(std::vector<int>& is, std::vector<double>* vals) {
    syncFromDevice();
    const size_t n = (size_t)numObjPerEnt * numValsPerObj;
    vals->resize(is.size() * n);
    for (size_t k = 0; k < is.size(); k++) std::copy(values.begin() + is[k] * n, values.begin() + (is[k] + 1) * n, vals->begin() + k * n);
  }
  void setDoubleValued(bool d) { doubleValued = d; }
  bool isDoubleValued() { return doubleValued; }

 protected:
  std::vector<double> values;
  Mesh* pmesh = nullptr;
  FieldType type = None;
  int numEntities = 0, numObjPerEnt = 0, numValsPerObj = 0;
  bool doubleValued = false;
  hfx_ctx* devCtx = nullptr;   // device context that holds a copy of this field (under devName) ...
  std::string devName;
  bool deviceNewer = false;    // ... which is ahead of `values`
};

inline void HDF5Io::load(std::string filename) {
  if (!myMesh) throw ErrorHandle("HDF5Io", "load", "must enter a mesh into the io before loading a file.");
  int hasMesh = 0, nFields = 0;
  std::vector<char> names(1 << 16);
  detail::check(hfx_host_h5_info(filename.c_str(), &hasMesh, &nFields, names.data(), (int)names.size()), nullptr);
  if (!hasMesh && nFields == 0) throw ErrorHandle("HDFIo", "load", "could not find Mesh or FieldData groups in file");
  if (hasMesh) {
    int nNodes = 0, dimSpace = 0, nCells = 0, nPerCell = 0;
    detail::check(hfx_host_read_h5_mesh(filename.c_str(), &nNodes, &dimSpace, &nCells, &nPerCell, 0, 0), nullptr);
    if (!myMesh->getReferenceElement() || myMesh->getReferenceElement()->getNumNodes() != nPerCell)
      throw ErrorHandle("HDF5Io", "loadMesh", "the cells of the file do not have the number of nodes of the reference element of the mesh");
    std::vector<double> pts((size_t)nNodes * dimSpace);
    std::vector<int> cells((size_t)nCells * nPerCell);
    detail::check(hfx_host_read_h5_mesh(filename.c_str(), &nNodes, &dimSpace, &nCells, &nPerCell, pts.data(), cells.data()), nullptr);
    myMesh->setMesh(dimSpace, pts, cells);
  }
  if (nFields > 0) {
    for (std::map<std::string, Field*>::iterator it = fieldMap.begin(); it != fieldMap.end(); ++it) {
      long long shape[3] = {0, 0, 0};
      int ft = -1;
      detail::check(hfx_host_read_h5_field(filename.c_str(), it->first.c_str(), shape, &ft, 0), nullptr);
      Field* f = it->second;
      *f->getFieldType() = (FieldType)ft; *f->getNumEntities() = (int)shape[0]; *f->getNumObjPerEnt() = (int)shape[1]; *f->getNumValsPerObj() = (int)shape[2];
      f->getValues()->assign((size_t)(shape[0] * shape[1] * shape[2]), 0.0);
      detail::check(hfx_host_read_h5_field(filename.c_str(), it->first.c_str(), shape, &ft, f->getValues()->data()), nullptr);
    }
  }
}
inline void HDF5Io::write(std::string filename) {
  const bool meshExists = myMesh != NULL && myMesh->getNumberPoints() > 0;
  if (!meshExists && fieldMap.empty()) throw ErrorHandle("HDFIo", "write", "could not find anything to write");
  std::vector<const char*> names; std::vector<int> ftypes; std::vector<long long> shapes; std::vector<const double*> vals;
  for (std::map<std::string, Field*>::iterator it = fieldMap.begin(); it != fieldMap.end(); ++it) {
    Field* f = it->second;
    const std::vector<double>* v = f->getValues();
    const long long per = (long long)*f->getNumObjPerEnt() * *f->getNumValsPerObj();
    names.push_back(it->first.c_str()); ftypes.push_back((int)*f->getFieldType());
    shapes.push_back(per ? (long long)v->size() / per : 0); shapes.push_back(*f->getNumObjPerEnt()); shapes.push_back(*f->getNumValsPerObj());
    vals.push_back(v->data());
  }
  const int nPerCell = meshExists ? myMesh->getReferenceElement()->getNumNodes() : 0;
  detail::check(hfx_host_write_h5(filename.c_str(), (unsigned)std::time(nullptr), meshExists ? myMesh->getNodeSpaceDimension() : 0, meshExists ? myMesh->getNumberPoints() : 0,
                                  meshExists ? myMesh->getPoints()->data() : nullptr, meshExists ? myMesh->getNumberCells() : 0, nPerCell,
                                  meshExists ? myMesh->getCells()->data() : nullptr, (int)names.size(), names.data(), ftypes.data(), shapes.data(), vals.data()), nullptr);
}

// ---- src/parallel/Partitioner.h, ZoltanPartitioner.h ----------------------------------------------------------------------------------
// One process per GPU.  The reference reads rank / size from MPI_COMM_WORLD, lets Zoltan cut the cell graph and MIGRATES cells, nodes, faces and
// field values between the ranks (Partitioner.cpp:203-563), then negotiates the shared faces over MPI (:42-107).  Here every process starts
// from the same global mesh, so the same end state is a pure function of (global mesh, cell partition vector, rank): update() replaces the
// mesh and the fields by this rank's part -- the owned cells followed by the ghost cells across the faces it owns (overlap 1: the face owner
// recomputes the element on the other side, so assembly needs no exchange) -- through the host C++ plan behind the C ABI (hfx_plan_create).
// Difference kept on purpose: the partitioned Mesh holds LOCAL ids in its connectivity (the device path indexes with them); the reference keeps
// global ids there and translates with global2Local*.  local2Global* / global2Local* / getSharedFaceList have the reference's meaning.
class Partitioner {
 public:
  Partitioner() {}
  explicit Partitioner(Mesh* pmesh) { setMesh(pmesh); }
  virtual ~Partitioner() { if (plan) hfx_plan_destroy(plan); }
  virtual void initialize() {   // Partitioner.cpp:5-11 with the launcher's RANK / WORLD_SIZE (torchrun, mpirun wrappers) instead of MPI_Comm_rank / size
    const char* r = std::getenv("RANK"); const char* w = std::getenv("WORLD_SIZE");
    initialize(r ? std::atoi(r) : 0, w ? std::atoi(w) : 1);
  }
  virtual void initialize(int rank_, int nPartitions_) {
    if (nPartitions_ < 1 || rank_ < 0 || rank_ >= nPartitions_) throw ErrorHandle("Partitioner", "initialize", "the rank must lie in [0, nPartitions)");
    rank = rank_; nPartitions = nPartitions_; initialized = 1;
  }
  virtual void computePartition() = 0;   // fills partitionVector: one rank id per cell of the global mesh
  virtual void setMesh(Mesh* pmesh) { myMesh = pmesh; if (pmesh) pmesh->setPartitioner(this); }
  virtual void setFields(std::vector<Field*> fieldList) { fields = fieldList; }
  virtual void update() {   // Partitioner.cpp:203-563: afterwards the mesh and the fields are this rank's part
    if (!initialized) throw ErrorHandle("Partitioner", "update", "must initialize the partitioner before updating.");
    if (!myMesh || !myMesh->getReferenceElement()) throw ErrorHandle("Partitioner", "update", "must set the mesh before updating.");
    const ReferenceElement* re = myMesh->getReferenceElement();
    const int dim = re->getDimension(), geom = re->getGeometry() == simplex ? HFX_SIMPLEX : HFX_ORTHOTOPE;
    const int nN = re->getNumNodes(), nv = geom == HFX_SIMPLEX ? dim + 1 : (1 << dim), nC = myMesh->getNumberCells(), dsp = myMesh->getNodeSpaceDimension();
    if ((int)partitionVector.size() != nC) throw ErrorHandle("Partitioner", "update", "must compute the partition before updating (one rank id per cell).");
    totNodes = myMesh->getNumberPoints(); totCells = nC; totFaces = myMesh->getNumberFaces();
    const std::vector<int>& gc = *myMesh->getCells();
    std::vector<int> lin((size_t)nC * nv);   // linear skeleton: the vertices come first in the node list of a cell
    for (int c = 0; c < nC; c++) for (int k = 0; k < nv; k++) lin[(size_t)c * nv + k] = gc[(size_t)c * nN + k];
    if (plan) { hfx_plan_destroy(plan); plan = nullptr; }
    if (hfx_plan_create(dim, geom, nC, lin.data(), partitionVector.data(), rank, nPartitions, &plan)) throw ErrorHandle(hfx_plan_last_error());
    long long sz[8];
    hfx_plan_sizes(plan, sz);
    nOwned = (int)sz[0];
    const int nL = (int)(sz[0] + sz[1]), nFl = (int)sz[3];
    std::vector<long long> cg((size_t)nL), fg((size_t)nFl), sfl((size_t)sz[7] * 3);
    hfx_plan_get(plan, cg.data(), 0, 0, fg.data(), 0, 0, 0, 0, 0, 0, 0, sfl.data());
    elementIDs.assign(cg.begin(), cg.end()); faceIDs.assign(fg.begin(), fg.end()); sharedFaceList.assign(sfl.begin(), sfl.end());
    // local nodes: ascending global id over the nodes of the local cells
    std::vector<int> used; used.reserve((size_t)nL * nN);
    for (int i = 0; i < nL; i++) for (int k = 0; k < nN; k++) used.push_back(gc[(size_t)elementIDs[i] * nN + k]);
    std::sort(used.begin(), used.end()); used.erase(std::unique(used.begin(), used.end()), used.end());
    nodeIDs = used;
    computeGlobal2LocalMaps();
    std::vector<double> pts((size_t)nodeIDs.size() * dsp);
    const std::vector<double>& gp = *myMesh->getPoints();
    for (size_t i = 0; i < nodeIDs.size(); i++) for (int d = 0; d < dsp; d++) pts[i * dsp + d] = gp[(size_t)nodeIDs[i] * dsp + d];
    std::vector<int> lc((size_t)nL * nN);
    for (int i = 0; i < nL; i++) for (int k = 0; k < nN; k++) lc[(size_t)i * nN + k] = glob2LocNodeIDs[gc[(size_t)elementIDs[i] * nN + k]];
    // global vertex id of the vertex nodes (the rank-independent order in which face blocks travel is defined with them)
    std::vector<long long> gv(nodeIDs.size(), -1);
    for (int i = 0; i < nL; i++) for (int k = 0; k < nv; k++) gv[(size_t)lc[(size_t)i * nN + k]] = gc[(size_t)elementIDs[i] * nN + k];
    // fields first (they still refer to the global mesh), then the mesh
    std::vector<std::vector<double> > newVals(fields.size());
    for (size_t f = 0; f < fields.size(); f++) {
      Field* F = fields[f];
      const std::vector<int>* ids = *F->getFieldType() == Node ? &nodeIDs : *F->getFieldType() == Face ? &faceIDs : *F->getFieldType() == Cell ? &elementIDs : nullptr;
      if (!ids) throw ErrorHandle("Partitioner", "update", "the field type is not supported");
      const size_t n = (size_t)(*F->getNumObjPerEnt()) * (*F->getNumValsPerObj());
      newVals[f].resize(ids->size() * n);
      for (size_t i = 0; i < ids->size(); i++) std::copy(F->getValues()->begin() + (size_t)(*ids)[i] * n, F->getValues()->begin() + ((size_t)(*ids)[i] + 1) * n, newVals[f].begin() + i * n);
    }
    myMesh->setMesh(dsp, pts, lc);
    myMesh->setPartitioner(this);
    if (myMesh->getNumberFaces() != nFl) throw ErrorHandle("Partitioner", "update", "the local face numbering of the mesh differs from the plan's");
    for (size_t f = 0; f < fields.size(); f++) { fields[f]->computeNumEntities(); *fields[f]->getValues() = newVals[f]; }
    const int nNf = re->getFaceElement()->getNumNodes();
    canon.assign((size_t)nFl * nNf, 0);
    if (hfx_host_face_canonical_positions(dim, re->getOrder(), nFl, nNf, myMesh->getFaces()->data(), gv.data(), canon.data())) throw ErrorHandle(hfx_plan_last_error());
    // faces of the true domain boundary: a local face with one local cell is either on the boundary or on a cut of the partition
    updated = true;
  }
  // Partitioner.cpp:565-826 exchanges the values of cell / face fields on the shared entities over MPI.  On the device path the only shared values
  // are the traces of ghost faces: HDGSolver::solve exchanges them over NCCL (hfx_comm_halo_field) before the local recovery.
  virtual void updateSharedInformation() {
    if (!updated) throw ErrorHandle("Partitioner", "updateSharedInformation", "must update the partition before sharing information.");
  }
  virtual int getNumPartitions() const { return nPartitions; }
  virtual int getRank() const { return rank; }
  virtual int getTotalNumberNodes() const { return totNodes; }
  virtual int getTotalNumberEls() const { return totCells; }
  virtual int getTotalNumberFaces() const { return totFaces; }
  virtual int getNumberOwnedCells() const { return nOwned; }   // the local cells [0, nOwned) are owned, the rest are ghosts
  virtual const std::vector<int>* getSharedFaceList() const { return &sharedFaceList; }
  virtual int local2GlobalNode(int loc) const { return nodeIDs[loc]; }
  virtual int local2GlobalFace(int loc) const { return faceIDs[loc]; }
  virtual int local2GlobalEl(int loc) const { return elementIDs[loc]; }
  virtual void local2GlobalNodeSlice(const std::vector<int>& loc, std::vector<int>* glob) const { slice(loc, glob, nodeIDs); }
  virtual void local2GlobalFaceSlice(const std::vector<int>& loc, std::vector<int>* glob) const { slice(loc, glob, faceIDs); }
  virtual void local2GlobalElementSlice(const std::vector<int>& loc, std::vector<int>* glob) const { slice(loc, glob, elementIDs); }
  virtual int global2LocalNode(int glob) const { return find(glob2LocNodeIDs, glob); }
  virtual int global2LocalFace(int glob) const { return find(glob2LocFaceIDs, glob); }
  virtual int global2LocalElement(int glob) const { return find(glob2LocElementIDs, glob); }
  virtual void global2LocalNodeSlice(const std::vector<int>& glob, std::vector<int>* loc) const { loc->resize(glob.size()); for (size_t i = 0; i < glob.size(); i++) (*loc)[i] = global2LocalNode(glob[i]); }
  virtual void global2LocalFaceSlice(const std::vector<int>& glob, std::vector<int>* loc) const { loc->resize(glob.size()); for (size_t i = 0; i < glob.size(); i++) (*loc)[i] = global2LocalFace(glob[i]); }
  virtual void global2LocalElementSlice(const std::vector<int>& glob, std::vector<int>* loc) const { loc->resize(glob.size()); for (size_t i = 0; i < glob.size(); i++) (*loc)[i] = global2LocalElement(glob[i]); }
  virtual const std::vector<int>* getNodeIds() const { return &nodeIDs; }
  virtual const std::vector<int>* getFaceIds() const { return &faceIDs; }
  virtual const std::vector<int>* getCellIds() const { return &elementIDs; }
  const std::vector<int>* getPartitionVector() const { return &partitionVector; }
  // device side: the halo plan HDGSolver::allocate hands to the library, and the NCCL id every rank must share (created by rank 0 with
  // hfx_comm_unique_id and distributed by the host program, e.g. MPI_Bcast or exchangeCommunicatorId below)
  const hfx_plan* getPlan() const { return plan; }
  const std::vector<unsigned char>* getCanonicalFacePositions() const { return &canon; }
  void setCommunicatorId(const char id128[128]) { commId.assign(id128, id128 + 128); }
  bool hasCommunicatorId() const { return commId.size() == 128; }
  const char* getCommunicatorId() const { return commId.data(); }
  void exchangeCommunicatorId(const std::string& path) {   // rank 0 writes the id to `path` (atomically), the other ranks wait for it: for launchers without MPI
    char id[128];
    if (rank == 0) {
      detail::check(hfx_comm_unique_id(id), nullptr);
      const std::string tmp = path + ".tmp";
      FILE* f = std::fopen(tmp.c_str(), "wb");
      if (!f || std::fwrite(id, 1, 128, f) != 128) throw ErrorHandle("Partitioner", "exchangeCommunicatorId", "cannot write " + tmp);
      std::fclose(f);
      if (std::rename(tmp.c_str(), path.c_str())) throw ErrorHandle("Partitioner", "exchangeCommunicatorId", "cannot publish " + path);
    } else {
      for (int tries = 0;; tries++) {
        FILE* f = std::fopen(path.c_str(), "rb");
        if (f) { const size_t n = std::fread(id, 1, 128, f); std::fclose(f); if (n == 128) break; }
        if (tries > 6000) throw ErrorHandle("Partitioner", "exchangeCommunicatorId", "timed out waiting for " + path);
        struct timespec ts = {0, 10000000}; nanosleep(&ts, nullptr);
      }
    }
    setCommunicatorId(id);
  }

 protected:
  static void slice(const std::vector<int>& loc, std::vector<int>* glob, const std::vector<int>& ids) { glob->resize(loc.size()); for (size_t i = 0; i < loc.size(); i++) (*glob)[i] = ids[loc[i]]; }
  static int find(const std::map<int, int>& m, int g) { std::map<int, int>::const_iterator it = m.find(g); return it == m.end() ? -1 : it->second; }
  void computeGlobal2LocalMaps() {   // Partitioner.cpp:109-125
    glob2LocNodeIDs.clear(); glob2LocFaceIDs.clear(); glob2LocElementIDs.clear();
    for (size_t i = 0; i < nodeIDs.size(); i++) glob2LocNodeIDs[nodeIDs[i]] = (int)i;
    for (size_t i = 0; i < faceIDs.size(); i++) glob2LocFaceIDs[faceIDs[i]] = (int)i;
    for (size_t i = 0; i < elementIDs.size(); i++) glob2LocElementIDs[elementIDs[i]] = (int)i;
  }
  Mesh* myMesh = NULL;
  std::vector<Field*> fields;
  std::vector<int> nodeIDs, faceIDs, elementIDs, sharedFaceList, partitionVector;
  std::map<int, int> glob2LocNodeIDs, glob2LocFaceIDs, glob2LocElementIDs;
  int nPartitions = 1, rank = 0, totNodes = 0, totCells = 0, totFaces = 0, nOwned = 0;
  bool initialized = 0, updated = false;
  hfx_plan* plan = nullptr;
  std::vector<unsigned char> canon;
  std::vector<char> commId;
};

// Recursive coordinate bisection of the cell centroids: the deterministic stand-in for ZoltanPartitioner (Zoltan PHG is not available; its cuts are
// not pinned by any reference test, and solution fields do not depend on them).
class RcbPartitioner : public Partitioner {
 public:
  using Partitioner::Partitioner;
  void computePartition() override {
    if (!initialized) throw ErrorHandle("RcbPartitioner", "computePartition", "must initialize the partitioner before computing the partition.");
    if (!myMesh || !myMesh->getReferenceElement()) throw ErrorHandle("RcbPartitioner", "computePartition", "must set the mesh before computing the partition.");
    const ReferenceElement* re = myMesh->getReferenceElement();
    const int dim = re->getDimension(), geom = re->getGeometry() == simplex ? HFX_SIMPLEX : HFX_ORTHOTOPE, nN = re->getNumNodes(), nv = geom == HFX_SIMPLEX ? dim + 1 : (1 << dim);
    if (myMesh->getNodeSpaceDimension() != dim) throw ErrorHandle("RcbPartitioner", "computePartition", "the node space dimension must equal the dimension of the reference element");
    const int nC = myMesh->getNumberCells();
    std::vector<int> lin((size_t)nC * nv);
    for (int c = 0; c < nC; c++) for (int k = 0; k < nv; k++) lin[(size_t)c * nv + k] = (*myMesh->getCells())[(size_t)c * nN + k];
    partitionVector.assign((size_t)nC, 0);
    if (hfx_host_rcb_partition(dim, geom, myMesh->getNumberPoints(), myMesh->getPoints()->data(), nC, lin.data(), nPartitions, partitionVector.data())) throw ErrorHandle(hfx_plan_last_error());
  }
};
// Recursive bisection of the dual graph (cells adjacent through a face: what ZoltanPartitioner.cpp:169-260 gives Zoltan) by greedy graph growing.
class GraphPartitioner : public Partitioner {
 public:
  using Partitioner::Partitioner;
  void computePartition() override {
    if (!initialized) throw ErrorHandle("GraphPartitioner", "computePartition", "must initialize the partitioner before computing the partition.");
    if (!myMesh || !myMesh->getReferenceElement()) throw ErrorHandle("GraphPartitioner", "computePartition", "must set the mesh before computing the partition.");
    const ReferenceElement* re = myMesh->getReferenceElement();
    const int dim = re->getDimension(), geom = re->getGeometry() == simplex ? HFX_SIMPLEX : HFX_ORTHOTOPE, nN = re->getNumNodes(), nv = geom == HFX_SIMPLEX ? dim + 1 : (1 << dim);
    const int nC = myMesh->getNumberCells();
    std::vector<int> lin((size_t)nC * nv);
    for (int c = 0; c < nC; c++) for (int k = 0; k < nv; k++) lin[(size_t)c * nv + k] = (*myMesh->getCells())[(size_t)c * nN + k];
    partitionVector.assign((size_t)nC, 0);
    if (hfx_host_graph_partition(dim, geom, nC, lin.data(), nPartitions, partitionVector.data())) throw ErrorHandle(hfx_plan_last_error());
  }
};
// A partition computed elsewhere (e.g. the reference's Zoltan run: ZoltanPartitioner.cpp:35-167), one rank id per global cell.
class VectorPartitioner : public Partitioner {
 public:
  VectorPartitioner(Mesh* pmesh, const std::vector<int>& cellRanks) : Partitioner(pmesh), given(cellRanks) {}
  void computePartition() override { partitionVector = given; }
 protected:
  std::vector<int> given;
};

// ---- src/resolution/PetscOpts.h ------------------------------------------------------------------------------------------
typedef const char* KSPType;
typedef const char* PCType;
#ifndef KSPGMRES
#define KSPGMRES "gmres"
#define KSPCG "cg"
#define PCNONE "none"
#define PCJACOBI "jacobi"
#define PCBJACOBI "bjacobi"
#endif
struct PetscOpts {
  KSPType solverType{KSPGMRES};
  PCType preconditionnerType{PCJACOBI};
  double rtol{1e-6};
  int maxits{1000};
  bool verbose{1};
};

// ---- src/resolution/LinAlgebraInterface.h:23-162 -------------------------------------------------------------------------
class LinAlgebraInterface {
 public:
  virtual ~LinAlgebraInterface() {}
  virtual void initialize() = 0;
  virtual void configure() = 0;
  virtual void allocate(int ndofs, const std::vector<int>* diagSparsePattern = NULL, const std::vector<int>* offSparsePattern = NULL) = 0;
  virtual void addValMatrix(int i, int j, const double& val) = 0;
  virtual void addValsMatrix(std::vector<int>& is, std::vector<int>& js, const double* vals) = 0;   // vals row-major |is| x |js|
  virtual void addValRHS(int i, const double& val) = 0;
  virtual void addValsRHS(std::vector<int>& is, const double* vals) = 0;
  virtual void setValMatrix(int i, int j, const double& val) = 0;
  virtual void setValsMatrix(std::vector<int>& is, std::vector<int>& js, const double* vals) = 0;
  virtual void setValRHS(int i, const double& val) = 0;
  virtual void setValsRHS(std::vector<int>& is, const double* vals) = 0;
  virtual void zeroOutRows(std::vector<int>& is) = 0;
  virtual void assemble() = 0;
  virtual void assembleFlush() = 0;
  virtual void solve(std::vector<double>* solution) = 0;
  virtual void getSolutionOwnership(std::vector<int>* ownership) = 0;
  virtual void clearSystem() = 0;
  virtual void destroySystem() = 0;
  virtual int getNumDofs() const = 0;
};

// Drop-in for PetscInterface: device CSR + GMRES(30)/CG with Jacobi (src/resolution/PetscInterface.cpp:60-265)
class CudaLinAlgebraInterface : public LinAlgebraInterface {
 public:
  explicit CudaLinAlgebraInterface(PetscOpts options = PetscOpts(), int device = 0) : myOptions(options), dev(device) {}
  ~CudaLinAlgebraInterface() { if (lai) hfx_lai_destroy(lai); delete ctx; }
  void setOptions(PetscOpts options) { myOptions = options; if (lai) pushOpts(); }
  void initialize() override { ensure(); detail::lcheck(hfx_lai_initialize(lai), lai); }
  void configure() override { ensure(); pushOpts(); detail::lcheck(hfx_lai_configure(lai), lai); }
  void allocate(int ndofs, const std::vector<int>* diag = NULL, const std::vector<int>* off = NULL) override {
    ensure();
    detail::lcheck(hfx_lai_allocate(lai, ndofs, diag ? diag->data() : NULL, off ? off->data() : NULL), lai);
  }
  void addValMatrix(int i, int j, const double& val) override { ensure(); detail::lcheck(hfx_lai_add_val_matrix(lai, i, j, val), lai); }
  void addValsMatrix(std::vector<int>& is, std::vector<int>& js, const double* vals) override { ensure(); detail::lcheck(hfx_lai_add_vals_matrix(lai, (int)is.size(), is.data(), (int)js.size(), js.data(), vals), lai); }
  void addValRHS(int i, const double& val) override { ensure(); detail::lcheck(hfx_lai_add_val_rhs(lai, i, val), lai); }
  void addValsRHS(std::vector<int>& is, const double* vals) override { ensure(); detail::lcheck(hfx_lai_add_vals_rhs(lai, (int)is.size(), is.data(), vals), lai); }
  void setValMatrix(int i, int j, const double& val) override { ensure(); detail::lcheck(hfx_lai_set_val_matrix(lai, i, j, val), lai); }
  void setValsMatrix(std::vector<int>& is, std::vector<int>& js, const double* vals) override { ensure(); detail::lcheck(hfx_lai_set_vals_matrix(lai, (int)is.size(), is.data(), (int)js.size(), js.data(), vals), lai); }
  void setValRHS(int i, const double& val) override { ensure(); detail::lcheck(hfx_lai_set_val_rhs(lai, i, val), lai); }
  void setValsRHS(std::vector<int>& is, const double* vals) override { ensure(); detail::lcheck(hfx_lai_set_vals_rhs(lai, (int)is.size(), is.data(), vals), lai); }
  void zeroOutRows(std::vector<int>& is) override { ensure(); detail::lcheck(hfx_lai_zero_out_rows(lai, (int)is.size(), is.data()), lai); }
  void assemble() override { ensure(); detail::lcheck(hfx_lai_assemble(lai), lai); }
  void assembleFlush() override { ensure(); detail::lcheck(hfx_lai_assemble_flush(lai), lai); }
  void solve(std::vector<double>* solution) override {
    ensure();
    int n = 0;
    hfx_lai_get_num_dofs(lai, &n);
    solution->resize(n);
    detail::lcheck(hfx_lai_solve(lai, solution->data(), &stats), lai);
  }
  void getSolutionOwnership(std::vector<int>* own) override {
    ensure();
    int lo = 0, hi = 0;
    detail::lcheck(hfx_lai_get_solution_ownership(lai, &lo, &hi), lai);
    own->resize(hi - lo);
    for (int i = lo; i < hi; i++) (*own)[i - lo] = i;
  }
  void clearSystem() override { ensure(); detail::lcheck(hfx_lai_clear_system(lai), lai); }
  void destroySystem() override { ensure(); detail::lcheck(hfx_lai_destroy_system(lai), lai); }
  int getNumDofs() const override { int n = 0; if (lai) hfx_lai_get_num_dofs(lai, &n); return n; }
  // device plumbing shared with HDGSolver
  hfx_ctx* context() { ensure(); return ctx->h; }
  hfx_solve_opts cOpts() const {
    hfx_solve_opts o;
    o.ksp = std::string(myOptions.solverType) == KSPCG ? 1 : 0;
    const std::string pc(myOptions.preconditionnerType);
    o.pc = pc == PCNONE ? 0 : (pc == PCBJACOBI ? 2 : 1);
    o.restart = 30; o.maxits = myOptions.maxits; o.rtol = myOptions.rtol;
    return o;
  }
  const hfx_solve_stats& getStats() const { return stats; }
  hfx_solve_stats stats{0, 0.0, 0.0, 0};

 protected:
  void ensure() {
    if (!ctx) ctx = new detail::Context(dev);
    if (!lai) detail::check(hfx_lai_create(ctx->h, &lai), ctx->h);
  }
  void pushOpts() { hfx_solve_opts o = cOpts(); hfx_lai_set_opts(lai, &o); }
  PetscOpts myOptions;
  int dev = 0;
  detail::Context* ctx = nullptr;
  hfx_lai* lai = nullptr;
};

// ---- src/operator/TimeScheme.h, Euler.h --------------------------------------------------------------------------------------
class TimeScheme {
 public:
  explicit TimeScheme(const ReferenceElement* re) : refEl(re) {}
  virtual ~TimeScheme() {}
  void setTimeStep(double timeStep) { deltat = timeStep; }
  double getTimeStep() const { return deltat; }
  virtual int cKind() const = 0;

 protected:
  const ReferenceElement* refEl;
  double deltat = 0.0;
};
class Euler : public TimeScheme {
 public:
  Euler(const ReferenceElement* re, bool isExplicitUser = false) : TimeScheme(re), isExplicit(isExplicitUser) {
    if (isExplicit) throw ErrorHandle("Euler", "Euler", "the explicit Euler scheme has no device kernel");
  }
  int cKind() const override { return HFX_TS_EULER_IMPLICIT; }

 protected:
  bool isExplicit = false;
};

// ---- src/operator/RKType.h, RungeKutta.h -------------------------------------------------------------------------------------------
enum RKType { FEuler, EMidpoint, Heun, Kutta3, Heun3, SSPRK3, RK4, BEuler, IMidpoint, CrankNicolson, KS2, QZ2, ALX2, RK43 };
class RungeKutta : public TimeScheme {
 public:
  RungeKutta(const ReferenceElement* re, RKType type = CrankNicolson, std::vector<std::string> fields = std::vector<std::string>()) : TimeScheme(re) {
    setButcherTable(type);
    auxiliaryFields = fields;
    stageCounter = 0;
  }
  int cKind() const override { return HFX_TS_RUNGE_KUTTA; }
  // RungeKutta::setUpDB (RungeKutta.cpp:215-291): row s < nStages = [c_s | a_s0 .. a_s,nStages-1], last row = [0 | b]
  void setButcherTable(RKType type) {
    const double g = 1.0 - std::sqrt(2.0) / 2.0;
    switch (type) {
      case FEuler: tab(2, {0, 0, 0, 1}); break;
      case EMidpoint: tab(3, {0, 0, 0, 0.5, 0.5, 0, 0, 0, 1}); break;
      case Heun: tab(3, {0, 0, 0, 1, 1, 0, 0, 0.5, 0.5}); break;
      case Kutta3: tab(4, {0, 0, 0, 0, 0.5, 0.5, 0, 0, 1, -1, 2, 0, 0, 1.0 / 6, 2.0 / 3, 1.0 / 6}); break;
      case Heun3: tab(4, {0, 0, 0, 0, 1.0 / 3, 1.0 / 3, 0, 0, 2.0 / 3, 0, 2.0 / 3, 0, 0, 0.25, 0, 0.75}); break;
      case SSPRK3: tab(4, {0, 0, 0, 0, 1, 1, 0, 0, 0.5, 0.25, 0.25, 0, 0, 1.0 / 6, 1.0 / 6, 2.0 / 3}); break;
      case RK4: tab(5, {0, 0, 0, 0, 0, 0.5, 0.5, 0, 0, 0, 0.5, 0, 0.5, 0, 0, 1, 0, 0, 1, 0, 0, 1.0 / 6, 1.0 / 3, 1.0 / 3, 1.0 / 6}); break;
      case BEuler: tab(2, {1, 1, 0, 1}); break;
      case IMidpoint: tab(2, {0.5, 0.5, 0, 1}); break;
      case CrankNicolson: tab(3, {0, 0, 0, 1, 0.5, 0.5, 0, 0.5, 0.5}); break;
      case KS2: tab(3, {0.5, 0.5, 0, 1.5, -0.5, 2, 0, -0.5, 1.5}); break;
      case QZ2: tab(3, {0.25, 0.25, 0, 0.75, 0.5, 0.25, 0, 0.5, 0.5}); break;
      case ALX2: tab(3, {g, g, 0, 1, 1 - g, g, 0, 1 - g, g}); break;
      case RK43: tab(5, {0.5, 0.5, 0, 0, 0, 2.0 / 3, 1.0 / 6, 0.5, 0, 0, 0.5, -0.5, 0.5, 0.5, 0, 1, 1.5, -1.5, 0.5, 0.5, 0, 1.5, -1.5, 0.5, 0.5}); break;
    }
  }
  void setAuxiliaryFields(std::vector<std::string> fields) { auxiliaryFields = fields; }
  const std::vector<std::string>& getAuxiliaryFields() const { return auxiliaryFields; }
  int getStage() const { return stageCounter; }
  int getNumStages() const { return nT - 1; }
  std::vector<double> stageRow() const { return std::vector<double>(bTable.begin() + stageCounter * nT + 1, bTable.begin() + (stageCounter + 1) * nT); }
  // every field RungeKutta::setFieldMap requires for the current stage (:44-88)
  std::vector<std::string> fieldNames() const {
    std::vector<std::string> n(1, "OldSolution");
    for (size_t a = 0; a < auxiliaryFields.size(); a++) n.push_back("Old" + auxiliaryFields[a]);
    for (int k = 0; k < stageCounter; k++) {
      n.push_back("RKStage_" + std::to_string(k));
      for (size_t a = 0; a < auxiliaryFields.size(); a++) n.push_back("RKStage_" + auxiliaryFields[a] + "_" + std::to_string(k));
    }
    return n;
  }
  void computeStage(std::map<std::string, Field*>* fm) {   // RungeKutta.cpp:145-180
    if (stageCounter >= getNumStages()) throw ErrorHandle("RungeKutta", "computeStage", "cannot compute more stages than the method allows, think about computing the solution");
    const double invdt = 1.0 / deltat;
    if (hfx_ctx* h = deviceOf(fm)) {   // device AXPYs (hfx_field_lincomb): the fields do not leave the GPU between the stages
      for (size_t b = 0; b <= auxiliaryFields.size(); b++) {
        const std::string base = b == 0 ? "Solution" : auxiliaryFields[b - 1];
        lincomb(h, fm, stageName(base, stageCounter), {invdt, -invdt}, {base, "Old" + base});
        std::vector<double> cf(1, 1.0); std::vector<std::string> nm(1, "Old" + base);
        for (int j = 0; j < stageCounter + 1; j++) { cf.push_back(deltat * bTable[stageCounter * nT + 1 + j]); nm.push_back(stageName(base, j)); }
        lincomb(h, fm, base, cf, nm);
      }
      stageCounter += 1;
      return;
    }
    for (size_t b = 0; b <= auxiliaryFields.size(); b++) {
      const std::string base = b == 0 ? "Solution" : auxiliaryFields[b - 1];
      std::vector<double>& sol = *fm->at(base)->getValues();
      const std::vector<double>& old = *fm->at("Old" + base)->getValues();
      std::vector<double>& rk = *fm->at(stageName(base, stageCounter))->getValues();
      for (size_t i = 0; i < sol.size(); i++) {
        rk[i] = invdt * (sol[i] - old[i]);
        double buffer = 0.0;
        for (int j = 0; j < stageCounter + 1; j++) buffer += bTable[stageCounter * nT + 1 + j] * (*fm->at(stageName(base, j))->getValues())[i];
        sol[i] = old[i] + deltat * buffer;
      }
    }
    stageCounter += 1;
  }
  void computeSolution(std::map<std::string, Field*>* fm) {   // RungeKutta.cpp:182-213
    if (stageCounter != getNumStages()) throw ErrorHandle("RungeKutta", "computeSolution", "all stages must be computed before computing the solution");
    if (hfx_ctx* h = deviceOf(fm)) {
      for (size_t b = 0; b <= auxiliaryFields.size(); b++) {
        const std::string base = b == 0 ? "Solution" : auxiliaryFields[b - 1];
        std::vector<double> cf(1, 1.0); std::vector<std::string> nm(1, "Old" + base);
        for (int k = 0; k < getNumStages(); k++) { cf.push_back(deltat * bTable[stageCounter * nT + 1 + k]); nm.push_back(stageName(base, k)); }
        lincomb(h, fm, base, cf, nm);
      }
      stageCounter = 0;
      return;
    }
    for (size_t b = 0; b <= auxiliaryFields.size(); b++) {
      const std::string base = b == 0 ? "Solution" : auxiliaryFields[b - 1];
      std::vector<double>& sol = *fm->at(base)->getValues();
      sol = *fm->at("Old" + base)->getValues();
      for (int k = 0; k < getNumStages(); k++) {
        const double bkdt = bTable[stageCounter * nT + 1 + k] * deltat;
        const std::vector<double>& rk = *fm->at(stageName(base, k))->getValues();
        for (size_t i = 0; i < rk.size(); i++) sol[i] += bkdt * rk[i];
      }
    }
    stageCounter = 0;
  }

 protected:
  static std::string stageName(const std::string& base, int k) { return base == "Solution" ? "RKStage_" + std::to_string(k) : "RKStage_" + base + "_" + std::to_string(k); }
  // the device context on which the solution fields of `fm` live (after HDGSolver::solve), or NULL: host arithmetic.  At most 8 terms per combination (7 stages).
  hfx_ctx* deviceOf(std::map<std::string, Field*>* fm) const {
    if (std::getenv("HFX_HOST_FIELD_ARITHMETIC") || getNumStages() > 7) return nullptr;
    std::map<std::string, Field*>::iterator it = fm->find("Solution");
    return it == fm->end() ? nullptr : it->second->deviceContext();
  }
  static void lincomb(hfx_ctx* h, std::map<std::string, Field*>* fm, const std::string& dst, const std::vector<double>& cf, const std::vector<std::string>& nm) {
    std::vector<const char*> names(nm.size());
    for (size_t k = 0; k < nm.size(); k++) {
      Field* f = fm->at(nm[k]);
      if (!f->isDeviceNewer(h, nm[k])) {   // the host copy is the newest (or the field never met the device): upload it
        detail::check(hfx_field_set(h, nm[k].c_str(), *f->getFieldType() == Node ? HFX_FIELD_NODE : (*f->getFieldType() == Face ? HFX_FIELD_FACE : HFX_FIELD_CELL), *f->getNumObjPerEnt(),
                                    *f->getNumValsPerObj(), f->getValues()->data(), f->isDoubleValued() ? 1 : 0), h);
        f->markOnDevice(h, nm[k], false);
      }
      names[k] = nm[k].c_str();
    }
    detail::check(hfx_field_lincomb(h, dst.c_str(), (int)nm.size(), cf.data(), names.data()), h);
    fm->at(dst)->markOnDevice(h, dst, true);
  }
  void tab(int n, std::initializer_list<double> v) { nT = n; bTable.assign(v.begin(), v.end()); }
  std::vector<double> bTable;
  int nT = 0;
  std::vector<std::string> auxiliaryFields;
  int stageCounter = 0;
};

// ---- src/model/FEModel.h, HDGModel.h, the HDG models ----------------------------------------------------------------------------
class FEModel {
 public:
  explicit FEModel(const ReferenceElement* re) : refEl(re) {}
  virtual ~FEModel() {}
  virtual void allocate(int nDOFsPerNode) = 0;
  void setTimeScheme(TimeScheme* ts) {   // src/model/FEModel.cpp:15-20
    if (allocated) throw ErrorHandle("FEModel", "setTimeScheme", "the time scheme must be set before allocation or field setting");
    timeScheme = ts;
  }
  virtual const AssemblyType* getAssemblyType() const { return &assembly; }
  const ReferenceElement* getReferenceElement() const { return refEl; }
  TimeScheme* getTimeScheme() const { return timeScheme; }
  typedef std::function<double(const std::vector<double>&)> ScalarFunction;

 protected:
  const ReferenceElement* refEl;
  TimeScheme* timeScheme = NULL;
  bool allocated = 0;
  AssemblyType assembly{Add, Add};
};

class HDGModel : public FEModel {
 public:
  using FEModel::FEModel;
  void allocate(int nDOFsPerNode) override {
    if (nDOFsPerNode < 1) throw ErrorHandle("HDGOperator", "allocate", "the number of DOFs per node must be at least one");
    nDOFsPNode = nDOFsPerNode; allocated = 1;
  }
  // operator descriptor for the device (Base is always present, src/model/HDGModel.cpp:28-32); fieldNames = names in the solver's field map
  virtual int opmask(const std::set<std::string>& fieldNames, bool strict) const = 0;
  virtual bool usesDiffusionField() const { return true; }
  virtual bool isBurgers() const { return false; }
  const ScalarFunction& sourceFunction() const { return source; }
  const ScalarFunction& reactionFunction() const { return reaction; }
  int getNumDOFsPerNode() const { return nDOFsPNode; }

  // ---- the per-element surface of the reference (src/model/FEModel.h:43-78) ------------------------------------------------------------------
  // compute() runs the DEVICE operators on a one-element mesh (hfx_get_local_matrix: the general kernel's dense local system before the
  // condensation), so the reference's model tests run against what the GPU assembles.  Field values are element-local, as the reference hands them
  // to a Model: Tau [nFaces x nNodesPerFace x nDOF^2], DiffusionTensor [nNodes x (1 | dim^2)], Velocity [nNodes x dim], BufferSolution / Solution
  // [nNodes x nDOF], Trace [nFaces x nNodesPerFace x nDOF].  Runge-Kutta models need the solver's stage fields and are served by HDGSolver only.
  struct LocalMatrix {
    int n = 0; std::vector<double> a;   // column-major n x n, unknown order [u | q | lambda]
    double operator()(int i, int j) const { return a[(size_t)i + (size_t)n * j]; }
    int rows() const { return n; } int cols() const { return n; }
  };
  void setElementNodes(const std::vector<std::vector<double> >* ns) {
    if (!ns || (int)ns->size() != refEl->getNumNodes()) throw ErrorHandle("FEModel", "setElementNodes", "the number of nodes does not match the reference element");
    elementNodes = ns;
  }
  virtual void setFieldMap(const std::map<std::string, std::vector<double> >* fm) {
    if (!fm || !fm->count("Tau")) throw ErrorHandle("HDGModel", "setFieldMap", "must provide a Tau field");
    localFieldMap = fm;
  }
  void setDevice(int d) { device = d; }
  virtual void compute() {
    if (!allocated) throw ErrorHandle("FEModel", "compute", "the model must be allocated before computing");
    if (!elementNodes) throw ErrorHandle("FEModel", "compute", "the nodes have not been set");
    if (!localFieldMap) throw ErrorHandle("FEModel", "compute", "the field map has not been set");
    if (timeScheme && timeScheme->cKind() != HFX_TS_EULER_IMPLICIT) throw ErrorHandle("FEModel", "compute", "a Runge-Kutta model needs the stage fields of the solver: use HDGSolver");
    const int dim = refEl->getDimension(), nN = refEl->getNumNodes(), nFc = refEl->getNumFaces(), nNf = refEl->getFaceElement()->getNumNodes(), nD = nDOFsPNode;
    const int dsp = (int)(*elementNodes)[0].size();
    struct Ctx { hfx_ctx* h = nullptr; ~Ctx() { if (h) hfx_ctx_destroy(h); } } C;
    detail::check(hfx_ctx_create(device, &C.h), nullptr);
    hfx_ctx* h = C.h;
    std::vector<double> pts((size_t)nN * dsp);
    for (int i = 0; i < nN; i++) for (int d = 0; d < dsp; d++) pts[(size_t)i * dsp + d] = (*elementNodes)[i][d];
    std::vector<int> cell(nN);
    for (int i = 0; i < nN; i++) cell[i] = i;
    detail::check(hfx_refel_set(h, dim, refEl->getOrder(), refEl->getGeometry() == simplex ? HFX_SIMPLEX : HFX_ORTHOTOPE), h);
    detail::check(hfx_mesh_set(h, nN, pts.data(), 1, cell.data()), h);
    // the faces of a one-element mesh are its local faces and their node lists the element's face nodes: local values map one to one
    const std::map<std::string, std::vector<double> >& lf = *localFieldMap;
    std::set<std::string> names;
    auto put = [&](const char* name, int type, int ents, int nObj, int nValsDefault) {
      std::map<std::string, std::vector<double> >::const_iterator it = lf.find(name);
      std::vector<double> zeros;
      const std::vector<double>* v = nullptr;
      int nVals = nValsDefault;
      if (it != lf.end()) {
        v = &it->second;
        if (v->size() % ((size_t)ents * nObj) != 0) throw ErrorHandle("HDGModel", "setFieldMap", std::string("the ") + name + " field does not have the right size");
        nVals = (int)(v->size() / ((size_t)ents * nObj));
        names.insert(name);
      } else if (nValsDefault > 0) { zeros.assign((size_t)ents * nObj * nValsDefault, 0.0); v = &zeros; }
      if (v) detail::check(hfx_field_set(h, name, type, nObj, nVals, v->data(), 0), h);
    };
    if (lf.at("Tau").size() != (size_t)nFc * nNf * nD * nD) throw ErrorHandle("HDGModel", "setFieldMap", "the Tau field does not have the right size");
    put("Tau", HFX_FIELD_FACE, nFc, nNf, nD * nD);
    put("Dirichlet", HFX_FIELD_FACE, nFc, nNf, nD);
    put("Trace", HFX_FIELD_FACE, nFc, nNf, nD);
    put("Solution", HFX_FIELD_CELL, 1, nN, nD);
    put("Flux", HFX_FIELD_CELL, 1, nN, nD * dim);
    if (usesDiffusionField()) put("DiffusionTensor", HFX_FIELD_NODE, nN, 1, 0);
    put("Velocity", HFX_FIELD_NODE, nN, 1, 0);
    put("BufferSolution", HFX_FIELD_CELL, 1, nN, 0);
    hfx_model_desc md;
    md.nDOF = nD; md.opmask = opmask(names, true);
    md.timeScheme = timeScheme ? timeScheme->cKind() : HFX_TS_NONE; md.dt = timeScheme ? timeScheme->getTimeStep() : 0.0;
    detail::check(hfx_model_describe(h, &md), h);
    static const int none = 0;
    detail::check(hfx_boundary_describe(h, 0, 0, &none), h);
    detail::check(hfx_allocate(h, 0), h);
    if (md.opmask & (HFX_OP_SOURCE | HFX_OP_REACTION)) {
      const int nIP = refEl->getNumIPs();
      std::vector<double> xip((size_t)nIP * dsp), pt(dsp);
      detail::check(hfx_ip_coords(h, xip.data()), h);
      evalLocalCallbacks(h, md.opmask, xip, nIP, dsp);
    }
    const int n = nN * nD * (1 + dim) + nFc * nNf * nD;
    localMatrix.n = n; localMatrix.a.assign((size_t)n * n, 0.0); localRHS.assign((size_t)n, 0.0);
    detail::check(hfx_get_local_matrix(h, 0, localMatrix.a.data(), localRHS.data()), h);
  }
  virtual const LocalMatrix* getLocalMatrix() const { return &localMatrix; }
  virtual const std::vector<double>* getLocalRHS() const { return &localRHS; }

 protected:
  virtual void evalLocalCallbacks(hfx_ctx* h, int mask, const std::vector<double>& xip, int nIP, int d) {
    std::vector<double> v((size_t)nIP), pt(d);
    for (int pass = 0; pass < 2; pass++) {
      const bool src = pass == 0;
      if (!(mask & (src ? HFX_OP_SOURCE : HFX_OP_REACTION))) continue;
      const ScalarFunction& fn = src ? source : reaction;
      for (int k = 0; k < nIP; k++) { pt.assign(xip.begin() + (size_t)k * d, xip.begin() + (size_t)(k + 1) * d); v[k] = fn(pt); }
      detail::check(src ? hfx_source_values(h, v.data()) : hfx_reaction_values(h, v.data()), h);
    }
  }
  int nDOFsPNode = 1, device = 0;
  ScalarFunction source, reaction;
  const std::vector<std::vector<double> >* elementNodes = nullptr;
  const std::map<std::string, std::vector<double> >* localFieldMap = nullptr;
  LocalMatrix localMatrix;
  std::vector<double> localRHS;
};

class HDGLaplaceModel : public HDGModel {   // Base + Diffusion(D = I)  (src/model/HDGLaplaceModel.cpp:18-30)
 public:
  using HDGModel::HDGModel;
  int opmask(const std::set<std::string>&, bool) const override { return HFX_OP_DIFFUSION; }
  bool usesDiffusionField() const override { return false; }
};

class HDGDiffusionSource : public HDGModel {   // Base + Diffusion(DiffusionTensor) ; rhs = Source  (src/model/HDGDiffusionSource.cpp:43-85)
 public:
  using HDGModel::HDGModel;
  void setSourceFunction(ScalarFunction s) {
    if (!allocated) throw ErrorHandle("HDGDiffusionSource", "setSourceFunction", "the model must be allocated before setting the source function");
    source = s;
  }
  int opmask(const std::set<std::string>&, bool strict) const override {
    if (!source && strict) throw ErrorHandle("Source", "calcSource", "must set a source function before calculating the source.");
    return HFX_OP_DIFFUSION | (source ? HFX_OP_SOURCE : 0);
  }
};

class HDGConvectionDiffusionReactionSource : public HDGModel {   // src/model/HDGConvectionDiffusionReactionSource.cpp:69-108
 public:
  using HDGModel::HDGModel;
  void setSourceFunction(ScalarFunction s) {
    if (!allocated) throw ErrorHandle("HDGConvectionDiffusionReactionSource", "setSourceFunction", "the model must be allocated before setting the source function");
    source = s;
  }
  void setReactionFunction(ScalarFunction r) {
    if (!allocated) throw ErrorHandle("HDGConvectionDiffusionReactionSource", "setReactionFunction", "the model must be allocated before setting the reaction function");
    reaction = r;
  }
  int opmask(const std::set<std::string>& names, bool) const override {
    const bool v = names.count("Velocity"), d = names.count("DiffusionTensor");
    if (!v && !d) throw ErrorHandle("HDGConvectionDiffusionReactionSource", "setFieldMap", "must provide at least either a Velocity field or a DiffusionTensor field");
    return (v ? HFX_OP_CONVECTION : 0) | (d ? HFX_OP_DIFFUSION : 0) | (reaction ? HFX_OP_REACTION : 0) | (source ? HFX_OP_SOURCE : 0);
  }
};

class HDGTransport : public HDGModel {   // Base + Convection, zero right-hand side  (src/model/HDGTransport.cpp:5-69)
 public:
  using HDGModel::HDGModel;
  bool usesDiffusionField() const override { return false; }
  int opmask(const std::set<std::string>& names, bool strict) const override {
    if (strict && !names.count("Velocity")) throw ErrorHandle("HDGTransport", "setFieldMap", "one must provide a Velocity field to use the Transport model.");
    return HFX_OP_CONVECTION;
  }
};

// Base + HDGUNabU [+ Diffusion if DiffusionTensor] ; rhs = UNabU rhs [+ one scalar Source per component]  (src/model/HDGBurgersModel.cpp:5-124)
class HDGBurgersModel : public HDGModel {
 public:
  using HDGModel::HDGModel;
  typedef std::function<double(const std::vector<double>&, int)> ComponentFunction;
  void allocate(int nDOFsPerNodeUser) override {
    if (nDOFsPerNodeUser != refEl->getDimension())
      throw ErrorHandle("HDGBurgersModel", "allocate", "the number of DOFs per node must be equal to the number of spatial dimensions for the Burgers equation");
    HDGModel::allocate(nDOFsPerNodeUser);
  }
  void setSourceFunction(ComponentFunction s) {
    if (!allocated) throw ErrorHandle("HDGBurgersModel", "setSourceFunction", "the model must be allocated before setting the source function");
    componentSource = s;
  }
  const ComponentFunction& componentSourceFunction() const { return componentSource; }
  bool isBurgers() const override { return true; }
  void evalLocalCallbacks(hfx_ctx* h, int mask, const std::vector<double>& xip, int nIP, int d) override {   // one scalar Source per component
    if (!(mask & HFX_OP_SOURCE)) return;
    std::vector<double> vc((size_t)d * nIP), pt(d);
    for (int c = 0; c < d; c++) for (int ip = 0; ip < nIP; ip++) { pt.assign(xip.begin() + (size_t)ip * d, xip.begin() + (size_t)(ip + 1) * d); vc[(size_t)c * nIP + ip] = componentSource(pt, c); }
    detail::check(hfx_source_values_n(h, d, vc.data()), h);
  }
  int opmask(const std::set<std::string>& names, bool) const override {
    if (!names.count("BufferSolution")) throw ErrorHandle("HDGBurgersModel", "setFieldMap", "must provide a BufferSolution field for the Newton-Raphson iterations");
    return HFX_OP_UNABU | (names.count("DiffusionTensor") ? HFX_OP_DIFFUSION : 0) | (componentSource ? HFX_OP_SOURCE : 0);
  }

 protected:
  ComponentFunction componentSource;
};

enum BoundaryModelType { CGType, HDGType };
class BoundaryModel : public FEModel {
 public:
  using FEModel::FEModel;
  BoundaryModelType getBoundaryModelType() const { return myType; }
  virtual int cKind() const = 0;

 protected:
  BoundaryModelType myType = CGType;
};
class DirichletModel : public BoundaryModel {   // assembly = {Set, Set}, I, g  (src/model/DirichletModel.cpp:19-44)
 public:
  explicit DirichletModel(const ReferenceElement* re) : BoundaryModel(re) { assembly.matrix = Set; assembly.rhs = Set; }
  void allocate(int) override { allocated = 1; }
  int cKind() const override { return HFX_BC_DIRICHLET; }
};
class IntegratedDirichletModel : public BoundaryModel {   // face mass, M g  (src/model/IntegratedDirichletModel.cpp)
 public:
  explicit IntegratedDirichletModel(const ReferenceElement* re) : BoundaryModel(re) { assembly.matrix = Set; assembly.rhs = Set; }
  void allocate(int) override { allocated = 1; }
  int cKind() const override { return HFX_BC_INTEGRATED_DIRICHLET; }
};

// ---- src/solver/Solver.h, HDGSolverOpts.h, HDGSolver.h ------------------------------------------------------------------------
enum HDGSolverType { IMPLICIT, WEXPLICIT, SEXPLICIT };
struct HDGSolverOpts { HDGSolverType type = IMPLICIT; bool verbosity = 1; };

class Solver {
 public:
  Solver() {}
  virtual ~Solver() {}
  virtual void setModel(FEModel* m) { model = m; }
  virtual void setBoundaryModel(BoundaryModel* m) {
    if (myMesh == NULL) throw ErrorHandle("Solver", "setBoundaryModel", "must set the Mesh before the boundary model.");
    boundaryList.push_back(std::make_tuple(m, myMesh->getBoundaryFaces()));
  }
  virtual void setBoundaryCondition(BoundaryModel* m, std::set<int>* faces) { boundaryList.push_back(std::make_tuple(m, faces)); }
  virtual void setLinSystem(LinAlgebraInterface* lai) { linSystem = lai; }
  virtual void setFieldMap(std::map<std::string, Field*>* fm) { fieldMap = fm; }
  virtual void setMesh(Mesh* m) { myMesh = m; meshUploaded = false; }
  virtual void setVerbosity(bool v) { verbose = v; }
  virtual void initialize() {   // src/solver/Solver.cpp:5-12
    if (linSystem != NULL) { linSystem->destroySystem(); linSystem->initialize(); linSystem->configure(); }
    initialized = 1;
  }
  virtual void allocate() = 0;
  virtual void assemble() = 0;
  virtual void solve() = 0;

 protected:
  LinAlgebraInterface* linSystem = NULL;
  FEModel* model = NULL;
  std::vector<std::tuple<BoundaryModel*, std::set<int>*> > boundaryList;
  std::map<std::string, Field*>* fieldMap = NULL;
  Mesh* myMesh = NULL;
  int nDOFsPerNode = 1;
  bool initialized = 0, allocated = 0, assembled = 0, verbose = 1, meshUploaded = false;
};

class HDGSolver : public Solver {
 public:
  using Solver::Solver;
  ~HDGSolver() {   // fields whose newest copy lives in this solver's context come home before the context goes away
    if (fieldMap) for (std::map<std::string, Field*>::iterator it = fieldMap->begin(); it != fieldMap->end(); ++it)
      if (it->second && it->second->deviceContext() && ownCtx && it->second->deviceContext() == ownCtx->h) { try { it->second->syncFromDevice(); } catch (...) {} it->second->markOnDevice(nullptr, "", false); }
    delete ownCtx;
  }
  // HDGSolver.h:41.  WEXPLICIT / SEXPLICIT: the trace problem is explicit in the current Solution / Flux (HDGSolver.cpp:346-354, hfx_solver_type)
  void setOptions(HDGSolverOpts opts) { myOpts = opts; verbose = opts.verbosity; }
  void setDevice(int d) { device = d; }
  void keepLocalS(bool k) { keepS = k; }
  void recomputeRecovery(bool r) { recompute = r; }   // HFX_RECOMPUTE_RECOVERY: U, Q are not stored (large meshes of order-4 tets)

  void allocate() override {   // src/solver/HDGSolver.cpp:5-106
    if (!initialized) throw ErrorHandle("HDGSolver", "allocate", "must initialize the solver before allocating.");
    if (myMesh == NULL) throw ErrorHandle("HDGSolver", "allocate", "must set the Mesh before allocating.");
    if (linSystem == NULL && myOpts.type != SEXPLICIT) throw ErrorHandle("HDGSolver", "allocate", "must set the linear system before allocating.");   // HDGSolver.cpp:12-14
    if (model == NULL) throw ErrorHandle("HDGSolver", "allocate", "must set the model before allocating.");
    if (boundaryList.empty()) throw ErrorHandle("HDGSolver", "allocate", "must set the boundary model before allocating.");
    if (fieldMap == NULL || fieldMap->size() == 0) throw ErrorHandle("HDGSolver", "allocate", "must set the fields before allocating.");
    const ReferenceElement* re = myMesh->getReferenceElement();
    const int nN = re->getNumNodes(), nNf = re->getFaceElement()->getNumNodes();
    Field* f = need("Solution");
    if (*f->getFieldType() != Cell) throw ErrorHandle("HDGSolver", "allocate", "the Solution field must be a cell field.");
    if (*f->getNumObjPerEnt() != nN) throw ErrorHandle("HDGSolver", "allocate", "the Solution field must have an object per element node.");
    nDOFsPerNode = *f->getNumValsPerObj();
    f = need("Flux");
    if (*f->getFieldType() != Cell) throw ErrorHandle("HDGSolver", "allocate", "the Flux field must be a cell field.");
    if (*f->getNumObjPerEnt() != nN) throw ErrorHandle("HDGSolver", "allocate", "the Flux field must have an object per element node.");
    if (*f->getNumValsPerObj() != nDOFsPerNode * myMesh->getNodeSpaceDimension()) throw ErrorHandle("HDGSolver", "allocate", "the Flux field must represent a spatial derivative of the Solution field.");
    f = need("Tau");
    if (*f->getFieldType() != Face) throw ErrorHandle("HDGSolver", "allocate", "the Tau field must be a face field.");
    if (*f->getNumObjPerEnt() != nNf) throw ErrorHandle("HDGSolver", "allocate", "the Tau field must have an object per element node.");
    if (*f->getNumValsPerObj() != nDOFsPerNode * nDOFsPerNode && *f->getNumValsPerObj() != 2 * nDOFsPerNode * nDOFsPerNode)
      throw ErrorHandle("HDGSolver", "allocate", "the Tau field must have the same or twice the number of values per object as the Solution field.");
    f = need("Trace");
    if (*f->getFieldType() != Face) throw ErrorHandle("HDGSolver", "allocate", "the Trace field must be a face field.");
    if (*f->getNumObjPerEnt() != nNf) throw ErrorHandle("HDGSolver", "allocate", "the Trace field must have an object per element node.");
    if (*f->getNumValsPerObj() != nDOFsPerNode) throw ErrorHandle("HDGSolver", "allocate", "the Trace field must have the same number of values per object as the Solution field.");
    hfx_ctx* h = ctx();
    if (!meshUploaded) {
      detail::check(hfx_refel_set(h, re->getDimension(), re->getOrder(), re->getGeometry() == simplex ? HFX_SIMPLEX : HFX_ORTHOTOPE), h);
      detail::check(hfx_mesh_set(h, myMesh->getNumberPoints(), myMesh->getPoints()->data(), myMesh->getNumberCells(), myMesh->getCells()->data()), h);
      detail::check(hfx_mesh_set_topology(h, myMesh->getNumberFaces(), myMesh->getFaces()->data(), myMesh->getCell2FaceMap()->data(), myMesh->getFace2CellMap()->data()), h);
      meshUploaded = true;
      xip.clear();
    }
    model->allocate(nDOFsPerNode);
    uploadInputs();
    describeModel(false);
    for (size_t i = 0; i < boundaryList.size(); i++) {
      BoundaryModel* bm = std::get<0>(boundaryList[i]);
      bm->allocate(nDOFsPerNode);
      if (bm->getBoundaryModelType() != CGType) throw ErrorHandle("HDGSolver", "applyBoundaryConditions", "only CGType boundary models have a device path");
      std::vector<int> ids(std::get<1>(boundaryList[i])->begin(), std::get<1>(boundaryList[i])->end());
      static const int none = 0;
      detail::check(hfx_boundary_describe(h, bm->cKind(), (int)ids.size(), ids.empty() ? &none : ids.data()), h);
    }
    detail::check(hfx_solver_type(h, (int)myOpts.type), h);
    detail::check(hfx_allocate(h, (keepS ? HFX_KEEP_LOCAL_S : 0) | (recompute ? HFX_RECOMPUTE_RECOVERY : 0)), h);
    // several GPUs (src/solver/HDGSolver.cpp:64-76 asks the Partitioner for the shared faces): NCCL communicator + halo plan of the partitioned mesh
    Partitioner* pp = myMesh->getPartitioner();
    if (pp && pp->getNumPartitions() > 1) {
      if (!pp->getPlan()) throw ErrorHandle("HDGSolver", "allocate", "the partitioner of the mesh must be updated before allocating.");
      if (!pp->hasCommunicatorId()) throw ErrorHandle("HDGSolver", "allocate", "the partitioner needs the communicator id shared by all ranks (setCommunicatorId / exchangeCommunicatorId).");
      detail::check(hfx_comm_init(h, pp->getNumPartitions(), pp->getRank(), pp->getCommunicatorId()), h);
      detail::check(hfx_comm_set_halo_plan(h, pp->getPlan(), pp->getCanonicalFacePositions()->data()), h);
    }
    allocated = 1;
  }

  void assemble() override {   // src/solver/HDGSolver.cpp:166-174
    if (!(initialized && allocated)) throw ErrorHandle("HDGSolver", "assemble", "the solver must be initialized and allocated before assembling.");
    uploadInputs();
    describeModel(true);
    evalCallbacks();
    detail::check(hfx_assemble(ctx()), ctx());
    assembled = 1;
  }

  void solve() override {   // src/solver/HDGSolver.cpp:677-779: linSystem->solve(Trace) + local recovery
    if (!assembled) throw ErrorHandle("HDGSolver", "solve", "system must be assembled before solving");
    hfx_solve_opts o{0, 1, 30, 1000, 1e-6};
    CudaLinAlgebraInterface* cl = dynamic_cast<CudaLinAlgebraInterface*>(linSystem);
    if (cl) o = cl->cOpts();
    detail::check(hfx_solver_type(ctx(), (int)myOpts.type), ctx());
    detail::check(hfx_solve(ctx(), &o, &stats), ctx());
    if (cl) cl->stats = stats;
    const char* out[3] = {"Trace", "Solution", "Flux"};   // left on the device; Field::getValues() fetches them when somebody looks
    for (int k = 0; k < 3; k++) fieldMap->at(out[k])->markOnDevice(ctx(), out[k], true);
  }

  // parity hooks
  void getCSR(std::vector<long long>* rowptr, std::vector<int>* colidx, std::vector<double>* vals, std::vector<double>* rhs) {
    long long n = 0, nnz = 0;
    detail::check(hfx_get_csr(ctx(), &n, &nnz, 0, 0, 0, 0), ctx());
    if (rowptr) rowptr->resize(n + 1);
    if (colidx) colidx->resize(nnz);
    if (vals) vals->resize(nnz);
    if (rhs) rhs->resize(n);
    detail::check(hfx_get_csr(ctx(), 0, 0, rowptr ? rowptr->data() : 0, colidx ? colidx->data() : 0, vals ? vals->data() : 0, rhs ? rhs->data() : 0), ctx());
  }
  const hfx_solve_stats& getStats() const { return stats; }
  hfx_ctx* context() { return ctx(); }

 protected:
  Field* need(const char* name) {
    std::map<std::string, Field*>::iterator it = fieldMap->find(name);
    if (it == fieldMap->end()) throw ErrorHandle("HDGSolver", "allocate", std::string("the field map must have a ") + name + " field.");
    return it->second;
  }
  hfx_ctx* ctx() {
    CudaLinAlgebraInterface* cl = dynamic_cast<CudaLinAlgebraInterface*>(linSystem);
    if (cl) return cl->context();
    if (!ownCtx) ownCtx = new detail::Context(device);
    return ownCtx->h;
  }
  static int cType(FieldType t) { return t == Node ? HFX_FIELD_NODE : (t == Face ? HFX_FIELD_FACE : HFX_FIELD_CELL); }
  std::set<std::string> inputNames() const {
    static const char* in[] = {"Tau", "Dirichlet", "DiffusionTensor", "Velocity"};
    std::set<std::string> s;
    for (int k = 0; k < 4; k++) if (fieldMap->count(in[k])) s.insert(in[k]);
    const HDGModel* hm = dynamic_cast<const HDGModel*>(model);
    if (hm && !hm->usesDiffusionField()) s.erase("DiffusionTensor");   // HDGLaplaceModel never reads it
    if (RungeKutta* rk = dynamic_cast<RungeKutta*>(model->getTimeScheme())) {
      if (allocated) { const std::vector<std::string> nm = rk->fieldNames(); for (size_t k = 0; k < nm.size(); k++) if (fieldMap->count(nm[k])) s.insert(nm[k]); }
    } else if (model->getTimeScheme() && allocated) s.insert("Solution");
    if (hm && hm->isBurgers()) { if (fieldMap->count("BufferSolution")) s.insert("BufferSolution"); s.insert("Trace"); }
    if (myOpts.type != IMPLICIT && allocated) { s.insert("Solution"); s.insert("Flux"); }   // the explicit data of HDGSolver.cpp:349-353
    return s;
  }
  void uploadInputs() {
    const std::set<std::string> names = inputNames();
    for (std::set<std::string>::const_iterator it = names.begin(); it != names.end(); ++it) {
      Field* f = fieldMap->at(*it);
      if (f->isDeviceNewer(ctx(), *it)) continue;   // produced on the device and not looked at since
      detail::check(hfx_field_set(ctx(), it->c_str(), cType(*f->getFieldType()), *f->getNumObjPerEnt(), *f->getNumValsPerObj(), f->getValues()->data(), f->isDoubleValued() ? 1 : 0), ctx());
      f->markOnDevice(ctx(), *it, false);
    }
  }
  void describeModel(bool strict) {
    const HDGModel* hm = dynamic_cast<const HDGModel*>(model);
    if (!hm) throw ErrorHandle("HDGSolver", "allocate", "the model must be an HDGModel");
    hfx_model_desc md;
    md.nDOF = nDOFsPerNode; md.opmask = hm->opmask(inputNames(), strict);
    TimeScheme* ts = model->getTimeScheme();
    md.timeScheme = ts ? ts->cKind() : HFX_TS_NONE; md.dt = ts ? ts->getTimeStep() : 0.0;
    detail::check(hfx_model_describe(ctx(), &md), ctx());
    if (RungeKutta* rk = dynamic_cast<RungeKutta*>(ts)) {   // RungeKutta::apply runs inside the device assembly
      std::vector<std::string> aux = rk->getAuxiliaryFields();
      std::sort(aux.begin(), aux.end());
      if (aux.size() != 2 || aux[0] != "Flux" || aux[1] != "Trace")
        throw ErrorHandle("RungeKutta", "apply", "the stiffness matrix does not have the correct dimensions (the HDG path needs the auxiliary fields Flux and Trace)");
      if (strict) { const std::vector<std::string> nm = rk->fieldNames(); for (size_t k = 0; k < nm.size(); k++) if (!fieldMap->count(nm[k])) throw ErrorHandle("RungeKutta", "setFieldMap", "the field map must provide the field " + nm[k]); }
      const int nSt = rk->getNumStages(), st = std::min(rk->getStage(), nSt - 1);
      std::vector<double> row = rk->getStage() < nSt ? rk->stageRow() : std::vector<double>(nSt, 0.0);
      detail::check(hfx_time_scheme_rk(ctx(), st, nSt, row.data()), ctx());
    }
    mask = md.opmask;
  }
  void evalCallbacks() {   // std::function callbacks run on the host at x(IP) (Source.cpp:5-22, Reaction.cpp:5-22)
    if (!(mask & (HFX_OP_SOURCE | HFX_OP_REACTION))) return;
    const HDGModel* hm = static_cast<const HDGModel*>(model);
    const int nC = myMesh->getNumberCells(), nIP = myMesh->getReferenceElement()->getNumIPs(), d = myMesh->getNodeSpaceDimension();
    if (xip.empty()) { xip.resize((size_t)nC * nIP * d); detail::check(hfx_ip_coords(ctx(), xip.data()), ctx()); }
    std::vector<double> v((size_t)nC * nIP), pt(d);
    if (hm->isBurgers()) {   // one scalar Source per component, [nCells][dim][nIP] (HDGBurgersModel.cpp:51-56,112-122)
      const HDGBurgersModel::ComponentFunction& fn = static_cast<const HDGBurgersModel*>(hm)->componentSourceFunction();
      std::vector<double> vc((size_t)nC * d * nIP);
      for (int e = 0; e < nC; e++) for (int c = 0; c < d; c++) for (int ip = 0; ip < nIP; ip++) {
        const size_t k = (size_t)e * nIP + ip;
        pt.assign(xip.begin() + k * d, xip.begin() + (k + 1) * d);
        vc[((size_t)e * d + c) * nIP + ip] = fn(pt, c);
      }
      detail::check(hfx_source_values_n(ctx(), d, vc.data()), ctx());
      return;
    }
    for (int pass = 0; pass < 2; pass++) {
      const bool src = pass == 0;
      if (!(mask & (src ? HFX_OP_SOURCE : HFX_OP_REACTION))) continue;
      const FEModel::ScalarFunction& fn = src ? hm->sourceFunction() : hm->reactionFunction();
      for (size_t k = 0; k < v.size(); k++) { pt.assign(xip.begin() + k * d, xip.begin() + (k + 1) * d); v[k] = fn(pt); }
      detail::check(src ? hfx_source_values(ctx(), v.data()) : hfx_reaction_values(ctx(), v.data()), ctx());
    }
  }
  HDGSolverOpts myOpts;
  detail::Context* ownCtx = nullptr;
  int device = 0, mask = 0;
  bool keepS = false, recompute = false;
  std::vector<double> xip;
  hfx_solve_stats stats{0, 0.0, 0.0, 0};
};

// ---- src/solver/NonLinearWrapper.h / .cpp:41-79 -----------------------------------------------------------------------------
// ---- continuous Galerkin: src/model/LaplaceModel.h, DiffusionSource.h, src/solver/CGSolver.h (hfx_cg_*) --------------------------------------------------------
class LaplaceModel : public FEModel {   // localMatrix = Diffusion (DiffusionTensor field if given), assembly = {Add, None}  (src/model/LaplaceModel.cpp:15-52)
 public:
  using FEModel::FEModel;
  void allocate(int nDOFsPerNode) override { nDOFsPNode = nDOFsPerNode; allocated = 1; assembly.matrix = Add; }
  virtual int cgMask(const std::map<std::string, Field*>&) const { return HFX_OP_DIFFUSION; }
  virtual const ScalarFunction* sourceFunction() const { return nullptr; }
  virtual bool usesDiffusionField() const { return true; }

 protected:
  int nDOFsPNode = 1;
};
class DiffusionSource : public LaplaceModel {   // Diffusion + Source, no time scheme on the device CG path  (src/model/DiffusionSource.cpp)
 public:
  using LaplaceModel::LaplaceModel;
  void setSourceFunction(ScalarFunction s) {
    if (!allocated) throw ErrorHandle("DiffusionSource", "setSourceFunction", "the model must be allocated before setting the source function");
    source = s;
  }
  int cgMask(const std::map<std::string, Field*>&) const override {
    if (!source) throw ErrorHandle("Source", "calcSource", "must set a source function before calculating the source.");
    return HFX_OP_DIFFUSION | HFX_OP_SOURCE;
  }
  const ScalarFunction* sourceFunction() const override { return &source; }

 protected:
  ScalarFunction source;
};
class Transport : public LaplaceModel {   // Convection (Velocity node field), zero right-hand side  (src/model/Transport.cpp)
 public:
  using LaplaceModel::LaplaceModel;
  int cgMask(const std::map<std::string, Field*>& fm) const override {
    if (!fm.count("Velocity")) throw ErrorHandle("Transport", "setFieldMap", "one must provide a Velocity field to use the Transport model.");
    return HFX_OP_CONVECTION;
  }
  bool usesDiffusionField() const override { return false; }
};

class CGSolver : public Solver {   // src/solver/CGSolver.cpp: allocate :5-40 (+ calcSparsityPattern :261-335), assemble :42-246, solve :248-259
 public:
  using Solver::Solver;
  ~CGSolver() { delete ownCtx; }
  void setDevice(int d) { device = d; }
  void allocate() override {
    if (!initialized) throw ErrorHandle("CGSolver", "allocate", "must initialize the solver before allocating.");
    if (myMesh == NULL) throw ErrorHandle("CGSolver", "allocate", "must set the Mesh before allocating.");
    if (linSystem == NULL) throw ErrorHandle("CGSolver", "allocate", "must set the linear system before allocating.");
    if (model == NULL) throw ErrorHandle("CGSolver", "allocate", "must set the model before allocating.");
    if (boundaryList.empty()) throw ErrorHandle("CGSolver", "allocate", "must set the boundary model before allocating.");
    if (fieldMap == NULL || fieldMap->size() == 0) throw ErrorHandle("CGSolver", "allocate", "must set the fields before allocating.");
    std::map<std::string, Field*>::iterator it = fieldMap->find("Solution");
    if (it == fieldMap->end()) throw ErrorHandle("CGSolver", "allocate", "the field map must have a Solution field.");
    if (*it->second->getFieldType() != Node) throw ErrorHandle("CGSolver", "allocate", "the Solution field must be a nodal field.");
    nDOFsPerNode = *it->second->getNumObjPerEnt() * *it->second->getNumValsPerObj();
    if (!dynamic_cast<LaplaceModel*>(model)) throw ErrorHandle("CGSolver", "allocate", "the device CG path serves LaplaceModel and DiffusionSource");
    const ReferenceElement* re = myMesh->getReferenceElement();
    hfx_ctx* h = ctx();
    detail::check(hfx_refel_set(h, re->getDimension(), re->getOrder(), re->getGeometry() == simplex ? HFX_SIMPLEX : HFX_ORTHOTOPE), h);
    detail::check(hfx_mesh_set(h, myMesh->getNumberPoints(), myMesh->getPoints()->data(), myMesh->getNumberCells(), myMesh->getCells()->data()), h);
    detail::check(hfx_mesh_set_topology(h, myMesh->getNumberFaces(), myMesh->getFaces()->data(), myMesh->getCell2FaceMap()->data(), myMesh->getFace2CellMap()->data()), h);
    model->allocate(nDOFsPerNode);
    upload("Solution");
    hfx_model_desc md; md.nDOF = nDOFsPerNode; md.opmask = HFX_OP_DIFFUSION; md.timeScheme = HFX_TS_NONE; md.dt = 0.0;
    detail::check(hfx_model_describe(h, &md), h);
    for (size_t i = 0; i < boundaryList.size(); i++) {
      BoundaryModel* bm = std::get<0>(boundaryList[i]);
      bm->allocate(nDOFsPerNode);
      if (bm->cKind() != HFX_BC_DIRICHLET) throw ErrorHandle("CGSolver", "allocate", "the device CG path serves DirichletModel boundaries");
      std::vector<int> ids(std::get<1>(boundaryList[i])->begin(), std::get<1>(boundaryList[i])->end());
      static const int none = 0;
      detail::check(hfx_boundary_describe(h, HFX_BC_DIRICHLET, (int)ids.size(), ids.empty() ? &none : ids.data()), h);
    }
    detail::check(hfx_cg_allocate(h), h);
    allocated = 1;
  }
  void assemble() override {
    if (!(initialized && allocated)) throw ErrorHandle("CGSolver", "assemble", "must initialize and allocate the solver before allocating.");
    hfx_ctx* h = ctx();
    if (!fieldMap->count("Dirichlet")) throw ErrorHandle("DirichletModel", "setFieldMap", "must give a field named Dirichlet to the DirichletModel");
    upload("Dirichlet");
    const LaplaceModel* lm = static_cast<const LaplaceModel*>(model);
    if (lm->usesDiffusionField() && fieldMap->count("DiffusionTensor")) upload("DiffusionTensor");
    if (fieldMap->count("Velocity")) upload("Velocity");
    TimeScheme* ts = model->getTimeScheme();   // FEModel::compute (FEModel.cpp:22-33): implicit Euler on the device CG path, the nodal Solution is the old state
    if (ts && ts->cKind() != HFX_TS_EULER_IMPLICIT) throw ErrorHandle("CGSolver", "assemble", "the device CG path serves the implicit Euler scheme");
    if (ts) upload("Solution");
    hfx_model_desc md; md.nDOF = nDOFsPerNode; md.opmask = lm->cgMask(*fieldMap); md.timeScheme = ts ? HFX_TS_EULER_IMPLICIT : HFX_TS_NONE; md.dt = ts ? ts->getTimeStep() : 0.0;
    detail::check(hfx_model_describe(h, &md), h);
    if (md.opmask & HFX_OP_SOURCE) {
      const int nC = myMesh->getNumberCells(), nIP = myMesh->getReferenceElement()->getNumIPs(), d = myMesh->getNodeSpaceDimension();
      std::vector<double> xip((size_t)nC * nIP * d), v((size_t)nC * nIP), pt(d);
      detail::check(hfx_ip_coords(h, xip.data()), h);
      for (size_t k = 0; k < v.size(); k++) { pt.assign(xip.begin() + k * d, xip.begin() + (k + 1) * d); v[k] = (*lm->sourceFunction())(pt); }
      detail::check(hfx_source_values(h, v.data()), h);
    }
    detail::check(hfx_cg_assemble(h), h);
    assembled = 1;
  }
  void solve() override {
    if (!assembled) throw ErrorHandle("CGSolver", "solve", "system must be assembled before solving");
    hfx_solve_opts o{0, 1, 30, 1000, 1e-6};
    CudaLinAlgebraInterface* cl = dynamic_cast<CudaLinAlgebraInterface*>(linSystem);
    if (cl) o = cl->cOpts();
    detail::check(hfx_cg_solve(ctx(), &o, &stats), ctx());
    if (cl) cl->stats = stats;
    Field* f = fieldMap->at("Solution");
    f->markOnDevice(nullptr, "", false);
    detail::check(hfx_field_get(ctx(), "Solution", f->getValues()->data()), ctx());
  }
  const hfx_solve_stats& getStats() const { return stats; }

 private:
  static int cType(FieldType t) { return t == Node ? HFX_FIELD_NODE : (t == Face ? HFX_FIELD_FACE : HFX_FIELD_CELL); }
  void upload(const char* name) {
    Field* f = fieldMap->at(name);
    detail::check(hfx_field_set(ctx(), name, cType(*f->getFieldType()), *f->getNumObjPerEnt(), *f->getNumValsPerObj(), f->getValues()->data(), f->isDoubleValued() ? 1 : 0), ctx());
  }
  hfx_ctx* ctx() {
    CudaLinAlgebraInterface* cl = dynamic_cast<CudaLinAlgebraInterface*>(linSystem);
    if (cl) return cl->context();
    if (!ownCtx) ownCtx = new detail::Context(device);
    return ownCtx->h;
  }
  detail::Context* ownCtx = nullptr;
  int device = 0;
  hfx_solve_stats stats{0, 0.0, 0.0, 0};
};

class NonLinearWrapper {
 public:
  NonLinearWrapper() : mySolver(NULL), previousSolution(NULL), currentSolution(NULL), residual(0.0) {
    residualComputer = vanillaResidualComputer;
    linearizedSolver = vanillaLinearizedSolver;
  }
  void solve() {
    if (mySolver == NULL) throw ErrorHandle("NonLinearWrapper", "solve", "the Solver must be set before attempting to solve");
    if (currentSolution == NULL || previousSolution == NULL) throw ErrorHandle("NonLinearWrapper", "solve", "the current and previous Solutions should be set before attempting to solve");
    std::vector<double>* cur = currentSolution->getValues();
    std::vector<double>* prev = previousSolution->getValues();
    for (int it = 0; it < maxIters; it++) {
      linearizedSolver(mySolver);
      residual = residualComputer(currentSolution, previousSolution);
      if (verbose) std::cout << "Non-linear iteration " << it << " : residual = " << residual << std::endl;
      if (residual < resTol) break;
      for (size_t k = 0; k < cur->size(); k++) { const double v = (1.0 - dampening) * (*cur)[k] + dampening * (*prev)[k]; (*cur)[k] = v; (*prev)[k] = v; }
    }
  }
  double getResidual() { return residual; }
  void setResidualTolerance(double tol) { resTol = tol; }
  void setMaxIterations(int iters) { maxIters = iters; }
  void setVerbosity(bool v) { verbose = v; }
  void setSolutionFields(Field* currentSol, Field* prevSol) { currentSolution = currentSol; previousSolution = prevSol; }
  void setSolver(Solver* solver) { mySolver = solver; }
  void setResidualComputer(std::function<double(Field*, Field*)> rc) { residualComputer = rc; }
  void setLinearizedSolver(std::function<void(Solver*)> sc) { linearizedSolver = sc; }
  void setDampening(double damp) { dampening = damp; }

 protected:
  static double vanillaResidualComputer(Field* cur, Field* prev) {   // relative l2 change (NonLinearWrapper.cpp:12-39)
    double diff = 0.0, ref = 0.0;
    const std::vector<double>&c = *cur->getValues(), &p = *prev->getValues();
    for (size_t k = 0; k < c.size(); k++) { diff += (c[k] - p[k]) * (c[k] - p[k]); ref += p[k] * p[k]; }
    return ref != 0.0 ? std::sqrt(diff / ref) : std::sqrt(diff);
  }
  static void vanillaLinearizedSolver(Solver* s) { s->assemble(); s->solve(); }
  std::function<void(Solver*)> linearizedSolver;
  Solver* mySolver;
  Field* previousSolution;
  Field* currentSolution;
  std::function<double(Field*, Field*)> residualComputer;
  double residual, resTol = 1e-6, dampening = 0.0;
  int maxIters = 1000;   // the constructor overrides the header default of 20 (NonLinearWrapper.cpp:5)
  bool verbose = true;
};

}  // namespace hfox
#endif
