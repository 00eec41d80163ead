// Host-side mesh input of the product (C++17, no dependencies): Gmsh 2.2 ASCII reader and the straight-sided order-p mesh generator.
// What the reference does with MOAB in tools/convertGmsh2H5HO.cpp:117-397; same node numbering (pinned by the reference's own
// .msh -> .h5 fixture pairs, tests/test_meshio.py).
#pragma once
#include <string>
#include <vector>

namespace hfx {

struct MshFile {
  std::vector<double> nodes;            // [nNodes][3], ascending node tag
  std::vector<int> elems[4];            // elems[k]: [m][k+1] linear simplices of topological dimension k (k = 1..3), file order, 0-based
};

// Throws std::runtime_error("MeshIo : readMsh : ...").
void read_msh(const std::string& path, MshFile* out);

// lin: [nLin][dim]; cells: [nCells][dim+1]; existing[k] (k = 1..dim-1): lower-dimensional entities already present in the input file.
// Returns nodes [N][dim] and cells [nCells][nN] of the order-p mesh.
void high_order_mesh(int dim, int order, int nLin, const double* lin, int nCells, const int* cells, const std::vector<int> (&existing)[4],
                     std::vector<double>* nodes, std::vector<int>* hoCells);

// HDF5 mesh file of the reference (HDF5Io::loadMesh, src/io/HDF5Io.cpp:111-152): datasets /Mesh/Nodes (f8, [nNodes][dimNodeSpace]) and
// /Mesh/Cells (i4 or i8, [nCells][nN]).  Dependency-free reader of the subset of the format those files use: superblock version 0,
// symbol-table groups, version-1 object headers, contiguous layout.  Throws std::runtime_error("HDF5Io : loadMesh : ...").
struct H5Mesh {
  int dimNodeSpace = 0, nodesPerCell = 0;
  std::vector<double> nodes;
  std::vector<int> cells;
};
void read_h5_mesh(const std::string& path, H5Mesh* out);

// Field datasets of the reference's files (HDF5Io::loadFields / writeFields, src/io/HDF5Io.cpp:154-187,304-391): /FieldData/<name>, f8 [nEntities][nObjPerEnt][nValsPerObj]
// with an integer attribute "ftype" (FieldTypes.h: Node 0, Edge 1, Face 2, Cell 3).
struct H5Field {
  std::string name;
  int ftype = -1;
  long long shape[3] = {0, 0, 0};
  std::vector<double> vals;
};
bool h5_has_mesh(const std::string& path);
std::vector<std::string> h5_field_names(const std::string& path);
void read_h5_field(const std::string& path, const std::string& name, H5Field* out);   // throws std::runtime_error("HDF5Io : loadFields : ...")

// HDF5Io::write (src/io/HDF5Io.cpp:66-109, writeMesh :189-302, writeFields :304-391) without libhdf5: the same subset of the format, laid out the way libhdf5 1.10 lays
// out the files of the reference's tools (2 KB metadata / small-data aggregator blocks), so that a mesh-only file is byte-identical to the one convertGmsh2H5HO wrote
// for the same mesh, up to the modification times.  mesh may be NULL (fields only); mtime: seconds since the epoch stored in the dataset headers.
void write_h5(const std::string& path, const H5Mesh* mesh, const std::vector<H5Field>& fields, unsigned mtime);

}  // namespace hfx
