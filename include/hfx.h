/* hfx.h -- C ABI of the B200-native HyperFox element-by-element HDG path (libhfx.so).
 *
 * The reference has no FFI: the path sits behind C++ classes.  Each entry point below names the reference
 * interface it replaces (file:line under the reference tree).  The C++ mirror classes in include/hyperfox/
 * (hfox::Mesh, Field, HDGSolver, HDGLaplaceModel, ..., CudaLinAlgebraInterface) call only these functions.
 *
 * Conventions: every function returns 0 on success, non-zero on error; hfx_last_error() then returns a message in
 * the reference's ErrorHandle format "Class : function : message" (src/globals/ErrorHandle.cpp:5-7,41-44).
 * Host buffers are caller-owned; device buffers are library-owned.  All reals are FP64, ids are 32-bit int
 * (row pointers 64-bit).  Layouts are the reference's.
 */
#ifndef HFX_H
#define HFX_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct hfx_ctx hfx_ctx;

/* ---- context ------------------------------------------------------------------------------------------------ */
int hfx_ctx_create(int device, hfx_ctx** ctx);
int hfx_ctx_destroy(hfx_ctx* ctx);
const char* hfx_last_error(const hfx_ctx* ctx); /* ctx may be NULL: last error of a failed hfx_ctx_create */
int hfx_device_count(void);                     /* 0 when no CUDA device is usable: the product has no CPU fallback */
double hfx_fp64_peak(int device);               /* measured DFMA peak of the device in TFLOP/s (roofline denominator; <0 on error) */
double hfx_dmma_peak(int device);               /* measured DMMA (mma.sync.m8n8k4.f64) issue peak in TFLOP/s: the tensor sub-pipe's own ceiling (<0 on error) */

/* ---- reference element: ReferenceElement(dim, order, geom)  src/element/ReferenceElement.cpp:5-26 ------------ */
enum { HFX_SIMPLEX = 0, HFX_ORTHOTOPE = 1 };
/* dim 2 or 3.  Simplices: orders 1-5 (fused kernel for triangles of every order and tets of order <= 3 with one DOF per node, general kernel otherwise);
   orthotopes (ReferenceElement.cpp:624-627): quads of order 1-5, hexes of order 1-2, general kernel, multilinear geometry. */
int hfx_refel_set(hfx_ctx* ctx, int dim, int order, int geom);
int hfx_refel_info(const hfx_ctx* ctx, int* nN, int* nNf, int* nFc, int* nIP, int* nIPf);
/* getNodes/getFaceNodes/getIPCoords/getIPWeights/getIPShapeFunctions/getIPDerivShapeFunctions (ReferenceElement.cpp:499-540),
   any pointer may be NULL.  Shapes: nodes[nN][dim], ipCoords[nIP][dim], w[nIP], shape[nIP][nN], dshape[nIP][nN][dim],
   fshape[nIPf][nNf], fdshape[nIPf][nNf][dim-1], fw[nIPf], faceNodes[nFc][nNf]. */
int hfx_refel_tables(const hfx_ctx* ctx, double* nodes, double* ipCoords, double* w, double* shape, double* dshape,
                     double* fshape, double* fdshape, double* fw, int* faceNodes);
/* stateless variant usable without a GPU (host table builder only) */
int hfx_refel_host_tables(int dim, int order, int geom, int* sizes /*[5] nN,nNf,nFc,nIP,nIPf*/, double* nodes, double* ipCoords,
                          double* w, double* shape, double* dshape, double* fshape, double* fdshape, double* fw, int* faceNodes);

/* ---- mesh: Mesh::setMesh + computeFaces  src/mesh/Mesh.cpp:31-46,183-274 ------------------------------------- */
int hfx_mesh_set(hfx_ctx* ctx, int nNodes, const double* nodes /*[nNodes][dim]*/, int nCells, const int* cells /*[nCells][nN]*/);
/* optional: caller-supplied topology (e.g. MOAB's) instead of the built-in face builder */
int hfx_mesh_set_topology(hfx_ctx* ctx, int nFaces, const int* faces, const int* cell2face, const int* face2cell);
int hfx_mesh_sizes(const hfx_ctx* ctx, int* nNodes, int* nCells, int* nFaces, int* nBoundary);
int hfx_mesh_get_topology(const hfx_ctx* ctx, int* faces, int* cell2face, int* face2cell, int* boundary);
/* host-only variant of the face builder (no GPU needed): returns nFaces, fills caller arrays sized for the worst case
   (nCells*nFc faces); pass NULL arrays to only count */
int hfx_host_compute_faces(int dim, int order, int geom, int nCells, const int* cells, int* nFaces, int* faces, int* cell2face,
                           int* face2cell, int* nBoundary, int* boundary);

/* ---- mesh input, host only (no GPU needed): what tools/convertGmsh2H5HO.cpp:117-397 does through MOAB ------------- */
/* Gmsh 2.2 ASCII file (linear simplices).  First call with NULL arrays to get the sizes: nNodes and counts[k] = number of entities of
   topological dimension k (k = 1..3; counts[0] unused); then with nodes[nNodes][3] (ascending node tag) and elemsK[counts[K]][K+1]
   (0-based vertex ids, file order; any pointer may be NULL). */
int hfx_host_read_msh(const char* path, int* nNodes, int counts[4], double* nodes, int* elems1, int* elems2, int* elems3);
/* HDF5Io::loadMesh (src/io/HDF5Io.cpp:111-152) without libhdf5: /Mesh/Nodes [nNodes][dimNodeSpace] (f8) and /Mesh/Cells
   [nCells][nodesPerCell] (i4 / i8) of the reference's mesh files (superblock 0, symbol-table groups, contiguous datasets).  First call
   with NULL arrays for the sizes. */
int hfx_host_read_h5_mesh(const char* path, int* nNodes, int* dimNodeSpace, int* nCells, int* nodesPerCell, double* nodes, int* cells);
/* HDF5Io::write (src/io/HDF5Io.cpp:66-109; writeMesh :189-302, writeFields :304-391) without libhdf5: group Mesh (Nodes f8, Cells i4; nodes / cells NULL: no Mesh group)
   and group FieldData with one f8 dataset [nEntities][nObjPerEnt][nValsPerObj] per field (shapes[3k..3k+2]) carrying the int attribute "ftype" (the reference's
   FieldTypes.h values: Node 0, Edge 1, Face 2, Cell 3).  Same subset of the format as the reader, laid out as libhdf5 lays out the reference's files: a mesh-only file is
   byte-identical to the one tools/convertGmsh2H5HO.cpp wrote for that mesh, up to the modification times (mtime: seconds since the epoch). */
int hfx_host_write_h5(const char* path, unsigned mtime, int dimNodeSpace, long long nNodes, const double* nodes, long long nCells, int nodesPerCell, const int* cells,
                      int nFields, const char* const* names, const int* ftypes, const long long* shapes, const double* const* vals);
/* HDF5Io::load (src/io/HDF5Io.cpp:13-64): which groups a file holds; names: '\n'-separated field names (may be NULL) */
int hfx_host_h5_info(const char* path, int* hasMesh, int* nFields, char* names, int namesCap);
/* HDF5Io::loadFields (src/io/HDF5Io.cpp:154-187): shape, ftype attribute and values of /FieldData/<name>; vals NULL: sizes only */
int hfx_host_read_h5_field(const char* path, const char* name, long long shape[3], int* ftype, double* vals);
/* generateHigherOrderMesh (convertGmsh2H5HO.cpp:117-257): straight-sided order-p simplex mesh from a linear one, the reference's node
   numbering.  lin[nLin][dim], cells[nCells][dim+1]; existing1 / existing2: edges / triangles already present in the input file
   (they precede the generated ones, as in MOAB).  Pass NULL output arrays to only count; nodesOut[nNodesOut][dim], cellsOut[nCells][nN]. */
int hfx_host_high_order_mesh(int dim, int order, int nLin, const double* lin, int nCells, const int* cells, int nExisting1,
                             const int* existing1, int nExisting2, const int* existing2, int* nNodesOut, double* nodesOut, int* cellsOut);

/* ---- fields: Field(mesh, type, nObjPerEnt, nValsPerObj)  src/field/Field.cpp:41-61 ---------------------------- */
enum { HFX_FIELD_NODE = 0, HFX_FIELD_FACE = 2, HFX_FIELD_CELL = 1 };
/* names with a meaning on the path: "Tau", "Dirichlet", "DiffusionTensor", "Velocity", "Solution", "Flux", "Trace",
   "BufferSolution" (Solver.h:125-141, HDGSolver.cpp:24-73).  Copies host -> device. */
int hfx_field_set(hfx_ctx* ctx, const char* name, int type, int nObjPerEnt, int nValsPerObj, const double* vals, int doubleValued);
/* Same, asynchronous: returns once the copy is enqueued; `vals` (pinned memory for a real overlap) must stay unchanged until the next
   hfx_assemble / hfx_sync returns.  hfx_assemble starts on the first elements while the rest of a Face field is still in flight. */
int hfx_field_set_async(hfx_ctx* ctx, const char* name, int type, int nObjPerEnt, int nValsPerObj, const double* vals, int doubleValued);
int hfx_field_get(hfx_ctx* ctx, const char* name, double* vals); /* device -> host */
/* dst <- sum_k coefs[k] * field names[k] on the device (1 <= nTerms <= 8, equal lengths, dst may be one of the sources and is created with the
   layout of the first source if it does not exist): RungeKutta::computeStage / computeSolution (src/operator/RungeKutta.cpp:145-213) and the
   damped update of NonLinearWrapper (src/solver/NonLinearWrapper.cpp:55-70) without a host round trip of the fields */
int hfx_field_lincomb(hfx_ctx* ctx, const char* dst, int nTerms, const double* coefs, const char* const* names);
/* ||a - b||_2^2 and ||b||_2^2 of two device fields: the residual of NonLinearWrapper (NonLinearWrapper.cpp:12-34).  On a partitioned mesh cell fields count
   the owned cells only and the two sums are all-reduced over the ranks (the reference's two MPI_Allreduce) */
int hfx_field_diff_norm2(hfx_ctx* ctx, const char* a, const char* b, double* diff2, double* ref2);
int hfx_field_size(const hfx_ctx* ctx, const char* name, long long* n);

/* ---- model: the four HDG models as an operator descriptor  src/model/*.cpp (computeLocalMatrix/RHS) ----------- */
enum { HFX_OP_DIFFUSION = 1, HFX_OP_CONVECTION = 2, HFX_OP_REACTION = 4, HFX_OP_SOURCE = 8, HFX_OP_UNABU = 16 };
enum { HFX_TS_NONE = 0, HFX_TS_EULER_IMPLICIT = 1, HFX_TS_RUNGE_KUTTA = 2 };
typedef struct {
  int nDOF;       /* FEModel::allocate(nDOFsPerNode)                                   */
  int opmask;     /* Base is always present (HDGModel.cpp:28-32)                       */
  int timeScheme; /* FEModel::setTimeScheme (HDGModel.cpp:38-47, Euler.cpp:18-37)      */
  double dt;
} hfx_model_desc;
int hfx_model_describe(hfx_ctx* ctx, const hfx_model_desc* md);
/* RungeKutta::apply (src/operator/RungeKutta.cpp:90-143) with auxiliary fields {Flux, Trace}: Butcher row a_s0..a_s,nStages-1 of the current
   stage (lower triangular tables only).  Fields read at assemble: OldSolution, OldFlux (cell), OldTrace (face), and for k < stage
   RKStage_k, RKStage_Flux_k (cell), RKStage_Trace_k (face) -- the reference's names (RungeKutta.cpp:44-88) */
int hfx_time_scheme_rk(hfx_ctx* ctx, int stage, int nStages, const double* row);
/* std::function source/reaction callbacks (Source.h:36, Reaction.h:38) are evaluated by the host at x(IP):
   hfx_ip_coords returns x(ip) = sum_i phi_i(ip) x_i (Source.cpp:5-22), [nCells][nIP][dim] */
int hfx_ip_coords(hfx_ctx* ctx, double* xip);
int hfx_source_values(hfx_ctx* ctx, const double* vals /*[nCells][nIP]*/);
/* per-component sources of the Burgers model (HDGBurgersModel.cpp:51-56,112-122: one scalar Source per component), [nCells][nComp][nIP] */
int hfx_source_values_n(hfx_ctx* ctx, int nComp, const double* vals);
int hfx_reaction_values(hfx_ctx* ctx, const double* vals /*[nCells][nIP]*/);
/* Solver::setBoundaryCondition(BoundaryModel*, faces)  Solver.h:41-48; DirichletModel / IntegratedDirichletModel */
enum { HFX_BC_DIRICHLET = 0, HFX_BC_INTEGRATED_DIRICHLET = 1 };
int hfx_boundary_describe(hfx_ctx* ctx, int kind, int nFaces, const int* faceIds /* NULL: mesh boundary */);

/* ---- solver: HDGSolver::allocate / assemble / solve  src/solver/HDGSolver.cpp:5-106,166-174,677-779 ----------- */
enum { HFX_KEEP_LOCAL_S = 1,        /* also store the per-element S,S0 Fields the reference keeps (HDGSolver.cpp:101-104) */
       HFX_RECOMPUTE_RECOVERY = 2 /* do NOT store the recovery operators U,Q,U0,Q0 (HDGSolver.cpp:93-100; 67 KB per order-4 tet): hfx_recover re-condenses each
                                     element and applies them out of shared memory.  Straight-sided 3-D order-4 meshes (the large-element kernel) only. */ };
int hfx_allocate(hfx_ctx* ctx, int flags);
int hfx_assemble(hfx_ctx* ctx);
typedef struct {
  int ksp;       /* 0 GMRES (PetscOpts.h:14) ; 1 CG                                   */
  int pc;        /* 0 none, 1 point Jacobi (PetscOpts.h:16), 2 face-block Jacobi       */
  int restart;   /* 30 = PETSc default                                                  */
  int maxits;    /* PetscOpts.h:24: 1000                                                */
  double rtol;   /* PetscOpts.h:20: 1e-6                                                */
} hfx_solve_opts;
typedef struct { int iterations; double resnorm; double bnorm; int converged; } hfx_solve_stats;
/* HDGSolver::setOptions (src/solver/HDGSolver.h:41, HDGSolverOpts.h:6-15): IMPLICIT (0, default), WEXPLICIT (1), SEXPLICIT (2).  The explicit types keep U, Q, U0, Q0
   but make the trace problem explicit in the current Solution / Flux fields (HDGSolver.cpp:346-354: S = S_ll, S0 = F_l - S_lu u - S_lq q); WEXPLICIT assembles and
   solves the (face-block-diagonal) global system as usual (:605-624), SEXPLICIT solves every face on its own in hfx_solve (:626-667,709-729) */
enum { HFX_SOLVER_IMPLICIT = 0, HFX_SOLVER_WEXPLICIT = 1, HFX_SOLVER_SEXPLICIT = 2 };
int hfx_solver_type(hfx_ctx* ctx, int type);
int hfx_solve(hfx_ctx* ctx, const hfx_solve_opts* opts, hfx_solve_stats* stats); /* Trace <- solution; then recovery */
int hfx_recover(hfx_ctx* ctx);                                                    /* HDGSolver.cpp:741-775 */
/* what the last hfx_solve cost: device time per Krylov iteration (CUDA events around the whole solve / iterations), and on several GPUs the
   collectives it issued (replaces the MPI_Allreduce / VecScatter counts of KSPSolve's -log_view) */
typedef struct { float msPerIteration; long long allReduces, haloExchanges, haloBytesPerExchange, ownedFaces, interiorFaces, boundaryFaces; int nNeighbours;
                 float msPhase[4]; /* mean ms per iteration: operator (SpMV + halo + preconditioner), dots, reduction (+ all-reduce) + Hessenberg step, Gram-Schmidt update */
                 int transport;    /* 0 single GPU, 1 NCCL (ncclSend/Recv + ncclAllReduce), 2 NVLink peer memory (CUDA IPC: direct stores into the neighbours' ghost
                                      buffers + one-shot all-reduce; HFX_P2P=0 selects 1) */ } hfx_solve_info_t;
int hfx_solve_info(const hfx_ctx* ctx, hfx_solve_info_t* info);

/* ---- continuous-Galerkin path: CGSolver (src/solver/CGSolver.h, CGSolver.cpp) for the Laplace-type models on the same device backend ----------
   Models: LaplaceModel (src/model/LaplaceModel.cpp: Diffusion with an optional DiffusionTensor node field) and DiffusionSource without a time scheme (Diffusion +
   Source; hfx_model_describe with HFX_OP_DIFFUSION [| HFX_OP_SOURCE], nDOF = 1, source values through hfx_source_values at hfx_ip_coords); boundary: DirichletModel
   (hfx_boundary_describe kind HFX_BC_DIRICHLET, values in the Face field "Dirichlet", face-node order).  Needs hfx_refel_set, hfx_mesh_set and a Node field "Solution".
   hfx_cg_allocate  = CGSolver::allocate (:5-40) + calcSparsityPattern (:261-335): node-based CSR, sorted columns, explicit zeros
   hfx_cg_assemble  = CGSolver::assemble (:42-246): element loop (Add), zeroOutRows + Set of the boundary rows
   hfx_cg_solve     = CGSolver::solve (:248-259): Krylov on the device CSR, result in the "Solution" field (hfx_field_get)
   hfx_cg_get_csr   = parity hook (NULL arrays: sizes only) */
int hfx_cg_allocate(hfx_ctx* ctx);
int hfx_cg_assemble(hfx_ctx* ctx);
int hfx_cg_solve(hfx_ctx* ctx, const hfx_solve_opts* opts, hfx_solve_stats* stats);
int hfx_cg_get_csr(hfx_ctx* ctx, long long* nrows, long long* nnz, long long* rowptr, int* colidx, double* vals, double* rhs);
int hfx_sync(hfx_ctx* ctx);
/* timing of the last hfx_assemble (CUDA events on the library's stream), milliseconds */
int hfx_last_assemble_ms(const hfx_ctx* ctx, float* msTotal, float* msKernel);
/* which device kernel served the last hfx_assemble: 0 fused element-group kernel (hfx_assemble.cuh), 1 general kernel (hfx_generic.cuh),
   2 large-element kernel (hfx_big.cuh), 3 linear-tet kernel, one thread per element (hfx_p1.cuh); pivotFallback = 1 if the assembly was redone with partial pivoting after a vanishing pivot */
int hfx_last_assemble_kernel(const hfx_ctx* ctx, int* kernel, int* pivotFallback);
/* development aid: one assemble with per-phase clock64 counters of CTA 0 (cycles16[16]) */
int hfx_assemble_profile(hfx_ctx* ctx, long long* cycles16);

/* ---- multi-GPU: one process per GPU, NCCL over NVLink  (replaces src/parallel/Partitioner.cpp:42-107,565-826 for this path) -------- */
/* Every rank holds its owned cells plus the ghost cells across the faces it owns (overlap-1 recompute: assembly needs no exchange).
   The owner of a face assembles and solves its t trace rows; the other rank sees the face as a ghost column.  Per Krylov iteration:
   one grouped ncclSend/ncclRecv of the packed ghost-face blocks + one ncclAllReduce of the Gram-Schmidt dots; one more halo exchange of
   Trace before the local recovery (HDGSolver.cpp:730-732). */
int hfx_comm_unique_id(char* id128 /*[128] ncclUniqueId, created by rank 0 and broadcast by the host program*/);
int hfx_comm_init(hfx_ctx* ctx, int nRanks, int rank, const char* id128);
/* halo plan in LOCAL face ids: for neighbour k, sendFaces = owned faces that are ghosts on nbrRank[k] and recvFaces = ghost faces owned
   by nbrRank[k], both ordered by global face id (the sharedFaceList contract of Partitioner.h:223); ownedFace[nFaces] = 1 if owned;
   canonPos[nFaces][nNf] = position of each local face node in the rank-independent node order of its face (blocks travel in that
   order: the local order of a face comes from its first LOCAL cell and differs between ranks) */
int hfx_comm_set_halo(hfx_ctx* ctx, int nNbr, const int* nbrRank, const int* sendCount, const int* sendFaces, const int* recvCount,
                      const int* recvFaces, const unsigned char* ownedFace, const unsigned char* canonPos);
int hfx_comm_halo_field(hfx_ctx* ctx, const char* faceFieldName); /* Partitioner::updateSharedInformation for one face field */

/* ---- element partition and halo plan on the host (C++, csrc/host/hfx_partition.cpp): what the reference negotiates over MPI in
   src/parallel/Partitioner.cpp:42-107 (computeSharedFaces), :565-826 (updateSharedInformation) and ZoltanPartitioner.cpp:35-167, as a pure
   function of (global linear mesh, cell partition vector, rank): every rank derives its plan without communication. -------------------- */
/* recursive coordinate bisection of the cell centroids: deterministic stand-in for the Zoltan PHG partition (ZoltanPartitioner.cpp:14-32);
   verts [nVerts][dim], linCells [nCells][verticesPerCell], part [nCells] out.  A Zoltan partition vector can be passed to hfx_plan_create instead. */
int hfx_host_rcb_partition(int dim, int geom, long long nVerts, const double* verts, long long nCells, const int* linCells, int world, int* part);
/* graph partition of the dual graph of the mesh (cells adjacent through a face: the graph ZoltanPartitioner.cpp:169-260 hands to Zoltan GRAPH / PHG) by recursive
   bisection with greedy graph growing from a pseudo-peripheral cell; deterministic, balanced to one cell, needs no coordinates */
int hfx_host_graph_partition(int dim, int geom, long long nCells, const int* linCells, int world, int* part);
typedef struct hfx_plan hfx_plan;
/* plan of `rank`: owned cells + the ghost cells across the faces it owns (a face travels with its first adjacent cell, ZoltanPartitioner.cpp:83-133),
   local vertex / cell / face numbering, face ownership, send / receive lists per neighbour rank, the reference's sharedFaceList */
int hfx_plan_create(int dim, int geom, long long nCells, const int* linCells, const int* part, int rank, int world, hfx_plan** plan);
void hfx_plan_destroy(hfx_plan* plan);
const char* hfx_plan_last_error(void);
/* sizes[8] = nOwnedCells, nGhostCells, nLocalVertices, nLocalFaces, nNeighbours, nSendFaces, nRecvFaces, nSharedFaces */
int hfx_plan_sizes(const hfx_plan* plan, long long sizes[8]);
/* any pointer may be NULL.  cellsGlobal [nOwned+nGhost] (owned first), vertexIds [nLocalVertices], localCells [(nOwned+nGhost)][verticesPerCell],
   faceGlobal / faceOwner / ownedFace [nLocalFaces], nbrRank / sendCount / recvCount [nNeighbours], sendFaces [nSendFaces], recvFaces [nRecvFaces]
   (LOCAL face ids, per neighbour in ascending global face id), sharedFaceList [3 nSharedFaces] = [global face, other rank, global adjacent cell] (Partitioner.h:223) */
int hfx_plan_get(const hfx_plan* plan, long long* cellsGlobal, long long* vertexIds, int* localCells, long long* faceGlobal, int* faceOwner, unsigned char* ownedFace,
                 int* nbrRank, int* sendCount, int* recvCount, int* sendFaces, int* recvFaces, long long* sharedFaceList);
/* canonPos of hfx_comm_set_halo from the local high-order face connectivity and the global vertex id of every vertex node (-1 for the other nodes) */
int hfx_host_face_canonical_positions(int dim, int order, long long nFaces, int nNf, const int* faces, const long long* nodeVertexGid, unsigned char* canonPos);
/* the same for any cell geometry (HFX_SIMPLEX / HFX_ORTHOTOPE: quadrilateral faces of hexahedra keep their bilinear weights on the corners) */
int hfx_host_face_canonical_positions_geom(int dim, int order, int geom, long long nFaces, int nNf, const int* faces, const long long* nodeVertexGid, unsigned char* canonPos);
/* hfx_comm_set_halo with the lists of a plan (the local mesh of ctx must be the plan's local cells at the context's order) */
int hfx_comm_set_halo_plan(hfx_ctx* ctx, const hfx_plan* plan, const unsigned char* canonPos);

/* ---- parity hooks ----------------------------------------------------------------------------------------------- */
/* CSR of the global trace system: sorted columns, explicit zeros (PetscInterface.cpp:99-103).  nnz query with NULLs. */
int hfx_get_csr(hfx_ctx* ctx, long long* nrows, long long* nnz, long long* rowptr, int* colidx, double* vals, double* rhs);
/* per-element condensed blocks, column-major as the reference stores them (HDGSolver.cpp:336-341; the device keeps U, Q row-major and
   transposes here); any may be NULL */
/* parity hook: || rhs - A * Trace ||_2 and || rhs ||_2 of the assembled trace system with the current "Trace" field (one SpMV on the
   device; lets a test check the global system at sizes where hfx_get_csr is too large to bring back) */
int hfx_residual(hfx_ctx* ctx, double* rnorm, double* bnorm);
int hfx_get_local(hfx_ctx* ctx, int iEl, int nEl, double* S, double* S0, double* U, double* U0, double* Q, double* Q0);
/* The per-element Model surface (FEModel::compute / getLocalMatrix / getLocalRHS, src/model/FEModel.h:43-78; Operator::assemble / getMatrix,
   src/operator/Operator.h:27-41): the dense local system of element iEl as Model::compute leaves it -- operators + time scheme, BEFORE the static
   condensation -- computed on the device by the general kernel.  A: n x n column-major, n = nN nDOF (1 + dim) + nFc nNf nDOF, unknown order
   [u | q | lambda] as the reference (q index = (node * dim + d) * nDOF + dof); F: n.  Needs hfx_allocate; uses the current fields. */
int hfx_get_local_matrix(hfx_ctx* ctx, int iEl, double* A, double* F);
/* element -> global trace dof ids (matRowCols, HDGSolver.cpp:596) */
int hfx_get_elem_dofs(hfx_ctx* ctx, int iEl, int nEl, int* dofs);

/* ---- LinAlgebraInterface mirror  src/resolution/LinAlgebraInterface.h:23-162, PetscInterface.cpp ----------------- */
typedef struct hfx_lai hfx_lai;
int hfx_lai_create(hfx_ctx* ctx, hfx_lai** lai);
int hfx_lai_destroy(hfx_lai* lai);
const char* hfx_lai_last_error(const hfx_lai* lai);
int hfx_lai_set_opts(hfx_lai* lai, const hfx_solve_opts* opts);
int hfx_lai_initialize(hfx_lai* lai);
int hfx_lai_configure(hfx_lai* lai);
int hfx_lai_allocate(hfx_lai* lai, int ndofs, const int* diagPattern, const int* offPattern);
int hfx_lai_add_val_matrix(hfx_lai* lai, int i, int j, double val);
int hfx_lai_add_vals_matrix(hfx_lai* lai, int ni, const int* is, int nj, const int* js, const double* vals /*row-major ni x nj*/);
int hfx_lai_add_val_rhs(hfx_lai* lai, int i, double val);
int hfx_lai_add_vals_rhs(hfx_lai* lai, int ni, const int* is, const double* vals);
int hfx_lai_set_val_matrix(hfx_lai* lai, int i, int j, double val);
int hfx_lai_set_vals_matrix(hfx_lai* lai, int ni, const int* is, int nj, const int* js, const double* vals);
int hfx_lai_set_val_rhs(hfx_lai* lai, int i, double val);
int hfx_lai_set_vals_rhs(hfx_lai* lai, int ni, const int* is, const double* vals);
int hfx_lai_zero_out_rows(hfx_lai* lai, int ni, const int* is);
int hfx_lai_assemble(hfx_lai* lai);
int hfx_lai_assemble_flush(hfx_lai* lai);
int hfx_lai_solve(hfx_lai* lai, double* solution /*[ndofs]*/, hfx_solve_stats* stats);
int hfx_lai_get_solution_ownership(hfx_lai* lai, int* lo, int* hi);
int hfx_lai_clear_system(hfx_lai* lai);
int hfx_lai_destroy_system(hfx_lai* lai);
int hfx_lai_get_num_dofs(const hfx_lai* lai, int* n);

#ifdef __cplusplus
}
#endif
#endif /* HFX_H */
