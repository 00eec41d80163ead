/* ORACLE -- test infrastructure only.  Never linked, imported or executed by the product path
 * (hyperfox_b200/), only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 *
 * Plain C++17 CPU restatement (no Eigen / PETSc / MOAB) of the reference's element-by-element HDG path:
 *   src/operator/{Operator,HDGBase,HDGDiffusion,HDGConvection,Convection,Mass,Source,Reaction,Euler,RungeKutta,HDGUNabU}.cpp
 *   src/model/{HDGModel,HDGLaplaceModel,HDGDiffusionSource,HDGConvectionDiffusionReactionSource,HDGBurgersModel,
 *              DirichletModel,IntegratedDirichletModel}.cpp
 *   src/solver/{HDGSolver,NonLinearWrapper}.cpp, src/resolution/PetscInterface.cpp (CSR + GMRES(30)/Jacobi).
 * Each function cites the file:line it follows.  All matrices are column-major unless stated.
 */
#ifndef HFX_ORACLE_H
#define HFX_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int dim, nN, nNf, nFc, nIP, nIPf;
  const double* shape;    /* [nIP][nN]            ReferenceElement::getIPShapeFunctions      */
  const double* dshape;   /* [nIP][nN][dim]       getIPDerivShapeFunctions                   */
  const double* w;        /* [nIP]                                                           */
  const double* fshape;   /* [nIPf][nNf]          face element                               */
  const double* fdshape;  /* [nIPf][nNf][dim-1]                                              */
  const double* fw;       /* [nIPf]                                                          */
  const int* faceNodes;   /* [nFc][nNf]                                                      */
} orc_refel;

/* operator mask */
enum { ORC_OP_DIFFUSION = 1, ORC_OP_CONVECTION = 2, ORC_OP_REACTION = 4, ORC_OP_SOURCE = 8, ORC_OP_UNABU = 16 };
/* time scheme */
enum { ORC_TS_NONE = 0, ORC_TS_EULER_IMPLICIT = 1, ORC_TS_EULER_EXPLICIT = 2, ORC_TS_RK = 3 };
/* boundary kinds */
enum { ORC_BC_DIRICHLET = 0, ORC_BC_INTEGRATED_DIRICHLET = 1 };

typedef struct {
  int nDOF;
  int opmask;
  int diffComps;          /* 0: D = identity default (HDGDiffusion.cpp:102-105); 1: scalar per node; dim*dim: tensor per node */
  int timeScheme;
  double dt;
  /* RK: row of the Butcher table for the current stage (a_s0..a_s,nStages-1), stage index, number of stages */
  int rkStage, rkNumStages;
  const double* rkRow;
} orc_model;

/* element-local fields (already gathered / side-selected / permuted to element-local order) */
typedef struct {
  const double* nodes;     /* [nN][dim]                                       */
  const double* tau;       /* [nFc*nNf][nDOF*nDOF] (col-major nDOF x nDOF)    */
  const double* diff;      /* [nN][diffComps] or NULL                         */
  const double* vel;       /* [nN][dim] or NULL                               */
  const double* srcIP;     /* [nSrc][nIP] host-evaluated callbacks, or NULL (nSrc = 1, or dim for UNabU/Burgers) */
  const double* reacIP;    /* [nIP] or NULL                                   */
  const double* bufSol;    /* [nN][nDOF]  BufferSolution (UNabU)              */
  const double* trace;     /* [nFc*nNf][nDOF] previous trace (UNabU)          */
  const double* solOld;    /* [nN*nDOF]   Euler: "Solution"; RK: "OldSolution"*/
  const double* fluxOld;   /* RK aux OldFlux  [nN*dim*nDOF] or NULL           */
  const double* traceOld;  /* RK aux OldTrace [nFc*nNf*nDOF] or NULL          */
  const double* rkSol;     /* RK stages [rkStage][nN*nDOF]                    */
  const double* rkFlux;    /* [rkStage][q]                                    */
  const double* rkTrace;   /* [rkStage][l]                                    */
} orc_elfields;

/* geometry of one element: arrays of length nJ = nIP + nFc*nIPf (HDGModel.cpp:53-85) */
void orc_element_geometry(const orc_refel* re, const double* nodes, double* jac /*[nJ][dim*dim] row-major J(r,m); faces use (dim-1) rows*/,
                          double* invjac /*[nJ][dim*dim] row-major invJ(m,r) (stride dim; faces: dim x (dim-1) pseudo-inverse)*/, double* dV /*[nJ]*/,
                          double* normals /*[nFc*nIPf][dim]*/);

/* one element: local n x n matrix (col-major) + rhs  (Model::compute) */
void orc_local_system(const orc_refel* re, const orc_model* md, const orc_elfields* f, double* A, double* F);

/* individual operators for the operator-level parity tests (all n x n col-major, zeroed first) */
void orc_op_base(const orc_refel* re, int nDOF, const double* nodes, const double* tau, double* A);
void orc_op_diffusion(const orc_refel* re, int nDOF, const double* nodes, const double* diff, int diffComps, double* A);
void orc_op_convection(const orc_refel* re, int nDOF, const double* nodes, const double* vel, double* A);
void orc_op_mass(int nN, int nIP, const double* shape, const double* dV, double* M /*nN x nN*/);
void orc_op_unabu(const orc_refel* re, int nDOF, const double* nodes, const double* bufSol, const double* trace, double* A, double* rhs);

/* static condensation, HDGSolver.cpp:331-348 (Householder QR like Eigen::HouseholderQR; useLU=1 -> partial-pivot LU) */
void orc_condense(int u, int q, int l, const double* A, const double* F, int useLU,
                  double* U, double* Q, double* S, double* U0, double* Q0, double* S0);

typedef struct {
  int dim, nCells, nFaces, nNodes;
  const double* nodes;     /* [nNodes][dim] */
  const int* cells;        /* [nCells][nN]  */
  const int* faces;        /* [nFaces][nNf] */
  const int* cell2face;    /* [nCells][nFc] */
  const int* face2cell;    /* [nFaces][2]   */
} orc_mesh;

typedef struct {
  const double* tau; int tauVals;              /* Face field [nFaces][nNf][tauVals]; tauVals = nDOF^2 or 2 nDOF^2 (double valued) */
  const double* diff; int diffType;            /* 0 node field [nNodes][diffComps], 1 cell field [nCells][nN][diffComps] */
  const double* vel;                           /* node field [nNodes][dim] */
  const double* srcIP;                         /* [nCells][nSrc][nIP] */
  const double* reacIP;                        /* [nCells][nIP] */
  const double* bufSol;                        /* cell field [nCells][nN][nDOF] */
  const double* trace;                         /* face field [nFaces][nNf][nDOF] */
  const double* solOld;                        /* cell field */
  const double* fluxOld; const double* traceOld;
  const double* rkSol; const double* rkFlux; const double* rkTrace;  /* [stage][field] */
  const double* dirichlet;                     /* face field [nFaces][nNf][nDOF] */
  const int* bFaces; int nBFaces; int bcKind;  /* boundary face list (ascending, std::set order) */
} orc_fields;

/* HDGSolver::calcElementalMatrices + applyBoundaryConditions for elements [e0,e1): writes per-element col-major blocks. */
void orc_assemble_local(const orc_refel* re, const orc_model* md, const orc_mesh* m, const orc_fields* f, int e0, int e1, int useLU,
                        double* U, double* Q, double* S, double* U0, double* Q0, double* S0);
/* the same for HDGSolverOpts.type = WEXPLICIT / SEXPLICIT (HDGSolver.cpp:346-354): S = S_ll, S0 = F_l - S_lu sol - S_lq flux (cell fields Solution [nCells][u], Flux [nCells][q]) */
void orc_assemble_local_explicit(const orc_refel* re, const orc_model* md, const orc_mesh* m, const orc_fields* f, int e0, int e1, int useLU,
                                 const double* solCur, const double* fluxCur, double* U, double* Q, double* S, double* U0, double* Q0, double* S0);
void orc_apply_bc(const orc_refel* re, const orc_model* md, const orc_mesh* m, const orc_fields* f, double* S, double* S0);

/* CSR pattern (sorted columns, explicit zeros) HDGSolver.cpp:117-164 + PETSc AIJ; returns nnz; pass NULL colidx to count */
long long orc_csr_pattern(const orc_refel* re, int nDOF, const orc_mesh* m, long long* rowptr, int* colidx);
/* HDGSolver::assembleSystem: ADD element S (converted to row-major) + S0 into CSR / rhs */
void orc_scatter(const orc_refel* re, int nDOF, const orc_mesh* m, const double* S, const double* S0,
                 const long long* rowptr, const int* colidx, double* vals, double* rhs);
/* element -> global dof map (matRowCols, HDGSolver.cpp:579-599) */
void orc_elem_dofs(const orc_refel* re, int nDOF, const orc_mesh* m, int iEl, int* dofs /*[l]*/);

/* GMRES(restart) left-preconditioned (pc: 0 none, 1 point Jacobi, 2 block Jacobi of size bs), classical Gram-Schmidt,
   zero initial guess, stop on ||M^-1 r|| <= max(rtol ||M^-1 b||, 1e-50).  returns iterations; *resnorm = final prec. residual */
int orc_gmres(long long n, const long long* rowptr, const int* colidx, const double* vals, const double* b, double* x,
              int restart, int pc, int bs, double rtol, int maxits, double* resnorm);

/* HDGSolver::solve recovery :741-775 */
void orc_recover(const orc_refel* re, int nDOF, const orc_mesh* m, const double* traceVals,
                 const double* U, const double* Q, const double* U0, const double* Q0, double* sol, double* flux);

/* CPU baseline: assemble+condense+scatter for elements [0,nEl) with nThreads workers (one per "MPI rank"); returns seconds */
double orc_bench_assemble(const orc_refel* re, const orc_model* md, const orc_mesh* m, const orc_fields* f, int nThreads, int useLU,
                          const long long* rowptr, const int* colidx, double* vals, double* rhs);

#ifdef __cplusplus
}
#endif
#endif
