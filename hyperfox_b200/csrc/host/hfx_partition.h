// Host-side element partition and trace-halo plan of the multi-GPU path (one process per GPU).
// Replaces, for this path, src/parallel/Partitioner.cpp:42-107 (computeSharedFaces), :565-826 (updateSharedInformation: who sends which
// faces to whom) and the ownership rule of src/parallel/ZoltanPartitioner.cpp:83-133 (a face travels with its first adjacent cell).
// The plan is a pure function of (global linear mesh, cell partition vector, rank): every rank derives it without communication.
#pragma once
#include <cstdint>
#include <vector>

namespace hfx {

// Recursive coordinate bisection of the cell centroids (deterministic stand-in for Zoltan PHG, which is not available; any externally supplied
// partition vector can be used instead).  verts [nVerts][dim], cells [nCells][nv] (vertex ids), part [nCells] out.
void rcb_partition(int dim, long long nVerts, const double* verts, long long nCells, int nv, const int* cells, int world, int* part);

// Graph partition of the dual graph (cells adjacent through a face -- the graph the reference hands to Zoltan GRAPH / PHG, ZoltanPartitioner.cpp:169-260) by
// recursive bisection with greedy graph growing: breadth-first from a pseudo-peripheral cell of the subset, the first floor(k/2)/k of the cells in BFS order
// (ties by cell id) form one side; every part of a connected mesh is a union of at most a few BFS shells, balanced to one cell.  Deterministic; needs no coordinates.
void graph_partition(int dim, int geom, long long nCells, const int* cells, int world, int* part);

struct PartitionPlan {
  int dim = 0, geom = 0, rank = 0, world = 1, nv = 0, nFc = 0;
  long long nOwned = 0, nGhost = 0;
  std::vector<long long> cellsGlobal;       // owned cells (ascending global id) then ghost cells (ascending): the local cell order
  std::vector<long long> vertexIds;         // global vertex id of every local vertex (ascending)
  std::vector<int> localCells;              // [nOwned + nGhost][nv] local vertex ids
  std::vector<long long> faceGlobal;        // [nLocalFaces] global face id of every local face (local numbering = compute_faces on localCells)
  std::vector<int> faceOwner;               // [nLocalFaces] rank that owns (assembles and solves) the face
  std::vector<uint8_t> ownedFace;           // [nLocalFaces] 1 if this rank owns it
  std::vector<int> nbrs, sendCount, recvCount;   // neighbour ranks (ascending); faces per neighbour
  std::vector<int> sendFaces, recvFaces;    // LOCAL face ids, per neighbour in ascending global face id: owned faces the neighbour holds as ghosts / ghosts it owns
  std::vector<long long> sharedFaceList;    // triples [global face id, rank of the other partition, global id of the adjacent cell there] (Partitioner.h:223)
  std::vector<int> localCell2Face, localFace2Cell;   // local topology (linear skeleton)
};

// cells: global linear connectivity [nCells][nv]; part [nCells]; geom: 0 simplex, 1 orthotope
void build_partition_plan(int dim, int geom, long long nCells, const int* cells, const int* part, int rank, int world, PartitionPlan* plan);

// canon[F][a] = position of local face node a of face F in the rank-independent node order of that face: the face-element node order obtained when the
// face's vertices are taken in ascending GLOBAL vertex id (the local order of a face comes from its first LOCAL cell and differs between ranks).
// faces [nFaces][nNf] local high-order face connectivity (vertices first); nodeVertexGid [nNodes]: global vertex id of the vertex nodes (-1 elsewhere).
void face_canonical_positions(int dim, int order, long long nFaces, int nNf, const int* faces, const long long* nodeVertexGid, uint8_t* canon);
// same for the faces of orthotope cells (lines in 2-D: as above; quadrilaterals in 3-D): the canonical frame of a quadrilateral puts the corner with the smallest global
// id at (-1,-1) and, of its two neighbours, the one with the smaller id at (1,-1); a node keeps its bilinear weights on the corners
void face_canonical_positions_geom(int dim, int order, int geom, long long nFaces, int nNf, const int* faces, const long long* nodeVertexGid, uint8_t* canon);

}  // namespace hfx
