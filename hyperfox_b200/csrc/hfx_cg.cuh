// Continuous-Galerkin path: CGSolver::assemble (src/solver/CGSolver.cpp:42-246) for the scalar models -- LaplaceModel (src/model/LaplaceModel.cpp:
// Diffusion), DiffusionSource (Diffusion + Source) and Transport (src/model/Transport.cpp: Convection), each optionally under an implicit Euler step (FEModel::compute,
// src/model/FEModel.cpp:22-33 + Euler.cpp:18-37) -- with DirichletModel boundaries, on the device.  Node-based CSR (sorted columns, explicit
// zeros: CGSolver::calcSparsityPattern :261-335 + PETSc AIJ), one CTA per element pass:
//   J(ip) = sum_i dphi_i/dxi x_i, dV = w det J            (Operator.cpp:14-84)
//   G(ip, i) = J^-1 grad^ phi_i                            (Diffusion.cpp:33-36)
//   A_ij = sum_ip dV (D(ip) G(ip, i)) . G(ip, j)           (Diffusion.cpp:37-45; D interpolated from the nodes, scalar or dim x dim column-major: setDiffTensor :65-73)
//   F_i  = sum_ip dV f(x_ip) phi_i(ip)                     (Source.cpp:24-48; f evaluated by the host callback at hfx_ip_coords)
//   C_kl = sum_ip dV (v(ip) . G(ip, l)) phi_k(ip)          (Convection.cpp:5-49; v interpolated from the Velocity node field)
//   Euler: A <- dt A + M, F <- dt F + M u_old,  M_jk = sum_ip dV phi_j phi_k   (Mass.cpp:5-38, Euler.cpp:28-32)
// and the element block is added into the global rows of its nodes (linSystem->addValsMatrix / addValsRHS): floating-point atomics -- a CG entry has as many
// contributors as cells share the node pair, so unlike the HDG trace system the sum order is not fixed.  Any element the reference element supports (runtime sizes,
// curved or multilinear geometry: the Jacobian is evaluated at every cubature point).
#pragma once
#include "hfx_assemble.cuh"

namespace hfx {

struct CgParams {
  int nCells, dim, nN, nIP;
  const double* nodes; const int* cells;
  const double* shape; const double* dshape; const double* w;
  const double* diff; int diffComps;      // DiffusionTensor node field [nNodes][1 | dim^2] or NULL (identity)
  const double* srcIP;                    // [nCells][nIP] or NULL
  const double* vel;                      // Velocity node field [nNodes][dim] or NULL (Convection, src/operator/Convection.cpp)
  int hasDiffusion;                       // 0: Transport (Convection only)
  double eulerDt;                         // > 0: implicit Euler (Euler.cpp:18-37): A <- dt A + M, F <- dt F + M u_old
  const double* solOld;                   // Solution node field [nNodes] (the old state of the Euler step)
  const long long* rowptr; const int* colidx;
  const unsigned char* affine;            // [nCells] 1: the cell is the affine image of the reference element (served by cg_affine_kernel when skipAffine is set) or NULL
  int skipAffine;
  const double* refTab;                   // cg_affine_kernel: K^_rs [dim*dim][nN*nN], M^ [nN*nN], w phi [nIP][nN]
  int fv[4];                              // vertices spanning the affine frame (0,1,2,3 simplices; 0,1,3,4 orthotopes)
  // gather form of the fast path (cg_gather_kernel): per node the cells it belongs to (ascending) with its local index there, per cell C = detJ Jinv^T Jinv and detJ
  const long long* n2c; const int* n2cCell; const unsigned char* n2cLoc; const double* cellGeo; int nNodes; int maxRow;
  const unsigned short* pos;              // [nCells][nN][nN] position of column cells[e][j] inside row cells[e][i] (built once by cg_positions_kernel) or NULL
  double* vals; double* rhs; int* status;
};

__device__ __forceinline__ long long cg_find(const long long* __restrict__ rowptr, const int* __restrict__ colidx, int row, int col) {
  long long lo = rowptr[row], hi = rowptr[row + 1] - 1;
  while (lo < hi) { const long long mid = (lo + hi) >> 1; if (colidx[mid] < col) lo = mid + 1; else hi = mid; }
  return lo;   // the pattern holds every node pair of every cell
}

__global__ void __launch_bounds__(128) cg_element_kernel(const CgParams p) {
  extern __shared__ __align__(16) double smcg[];
  const int dim = p.dim, nN = p.nN, nIP = p.nIP, D2 = dim * dim, tid = threadIdx.x, NT = blockDim.x;
  double* const X = smcg;                               // [nN][dim]
  double* const JI = X + ((nN * dim + 1) & ~1);         // [nIP][dim][dim]
  double* const DP = JI + ((nIP * D2 + 1) & ~1);        // [nIP][dim][dim] row-major D(a, b)
  double* const DV = DP + ((nIP * D2 + 1) & ~1);        // [nIP]
  double* const VP = DV + ((nIP + 1) & ~1);             // [nIP][dim] velocity at the cubature points
  double* const UO = VP + ((nIP * dim + 1) & ~1);       // [nIP] old solution at the cubature points
  double* const G = UO + ((nIP + 1) & ~1);              // [nIP][nN][dim]
  int* const ID = reinterpret_cast<int*>(G + (size_t)nIP * nN * dim);
  for (int e = blockIdx.x; e < p.nCells; e += gridDim.x) {
    if (p.skipAffine && p.affine[e]) continue;      // (block-uniform)
    for (int i = tid; i < nN; i += NT) {
      const int n = p.cells[(size_t)e * nN + i];
      ID[i] = n;
      for (int m = 0; m < dim; m++) X[i * dim + m] = p.nodes[(size_t)n * dim + m];
    }
    __syncthreads();
    for (int ip = tid; ip < nIP; ip += NT) {
      double J[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
#pragma unroll
      for (int r = 0; r < 3; r++)
#pragma unroll
        for (int m = 0; m < 3; m++)
          if (r < dim && m < dim) {   // (compile-time indices: J and I stay in registers)
            double s = 0.0;
            for (int i = 0; i < nN; i++) s = fma(p.dshape[((size_t)ip * nN + i) * dim + r], X[i * dim + m], s);
            J[r][m] = s;
          }
      double det, I[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
      if (dim == 1) { det = J[0][0]; I[0][0] = 1.0 / det; }
      else if (dim == 2) {
        det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
        const double id = 1.0 / det;
        I[0][0] = J[1][1] * id; I[0][1] = -J[0][1] * id; I[1][0] = -J[1][0] * id; I[1][1] = J[0][0] * id;
      } else det_inv(J, det, I);
      if (!(fabs(det) > 1e-300)) atomicOr(p.status, 1);
#pragma unroll
      for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) if (a < dim && b < dim) JI[ip * D2 + a * dim + b] = I[a][b];
      DV[ip] = p.w[ip] * det;
      if (p.vel)
        for (int a = 0; a < dim; a++) {
          double s = 0.0;
          for (int i = 0; i < nN; i++) s = fma(p.shape[(size_t)ip * nN + i], p.vel[(size_t)ID[i] * dim + a], s);
          VP[ip * dim + a] = s;
        }
      if (p.eulerDt > 0.0) {
        double s = 0.0;
        for (int i = 0; i < nN; i++) s = fma(p.shape[(size_t)ip * nN + i], p.solOld[ID[i]], s);
        UO[ip] = s;
      }
      if (p.diff) {
        for (int a = 0; a < dim; a++)
          for (int b = 0; b < dim; b++) {
            double s = 0.0;
            if (p.diffComps == 1) { if (a == b) for (int i = 0; i < nN; i++) s = fma(p.shape[(size_t)ip * nN + i], p.diff[ID[i]], s); }
            else for (int i = 0; i < nN; i++) s = fma(p.shape[(size_t)ip * nN + i], p.diff[(size_t)ID[i] * D2 + a + dim * b], s);   // column-major per node
            DP[ip * D2 + a * dim + b] = s;
          }
      }
    }
    __syncthreads();
    for (int k = tid; k < nIP * nN; k += NT) {
      const int ip = k / nN;
      const double* ds = p.dshape + (size_t)k * dim;
      for (int a = 0; a < dim; a++) {
        double s = 0.0;
        for (int b = 0; b < dim; b++) s = fma(JI[ip * D2 + a * dim + b], ds[b], s);
        G[(size_t)k * dim + a] = s;
      }
    }
    __syncthreads();
    for (int idx = tid; idx < nN * nN; idx += NT) {
      const int i = idx / nN, j = idx - i * nN;
      double acc = 0.0, mss = 0.0;
      for (int ip = 0; ip < nIP; ip++) {
        const double* gi = G + ((size_t)ip * nN + i) * dim; const double* gj = G + ((size_t)ip * nN + j) * dim;
        double s = 0.0;
        if (p.hasDiffusion) {
          if (p.diff) {
            const double* Dm = DP + ip * D2;
            for (int a = 0; a < dim; a++) { double t = 0.0; for (int b = 0; b < dim; b++) t = fma(Dm[a * dim + b], gi[b], t); s = fma(t, gj[a], s); }
          } else for (int a = 0; a < dim; a++) s = fma(gi[a], gj[a], s);
        }
        const double pi_ = p.shape[(size_t)ip * nN + i];
        if (p.vel) {   // Convection.cpp:36-44: op(k, l) += dV (v . grad phi_l) phi_k
          double vg = 0.0;
          for (int a = 0; a < dim; a++) vg = fma(VP[ip * dim + a], gj[a], vg);
          s = fma(vg, pi_, s);
        }
        acc = fma(DV[ip], s, acc);
        if (p.eulerDt > 0.0) mss = fma(DV[ip] * pi_, p.shape[(size_t)ip * nN + j], mss);   // Mass.cpp:5-38
      }
      if (p.eulerDt > 0.0) acc = fma(p.eulerDt, acc, mss);                                    // Euler.cpp:28-32
      const long long at = p.pos ? p.rowptr[ID[i]] + p.pos[(size_t)e * nN * nN + idx] : cg_find(p.rowptr, p.colidx, ID[i], ID[j]);
      atomicAdd(p.vals + at, acc);
    }
    if (p.srcIP || p.eulerDt > 0.0)
      for (int i = tid; i < nN; i += NT) {
        double s = 0.0, mu = 0.0;
        for (int ip = 0; ip < nIP; ip++) {
          const double wphi = DV[ip] * p.shape[(size_t)ip * nN + i];
          if (p.srcIP) s = fma(wphi, p.srcIP[(size_t)e * nIP + ip], s);
          if (p.eulerDt > 0.0) mu = fma(wphi, UO[ip], mu);            // (M u_old)_i = sum_ip dV phi_i u_old(ip)
        }
        atomicAdd(p.rhs + ID[i], p.eulerDt > 0.0 ? fma(p.eulerDt, s, mu) : s);
      }
    __syncthreads();
  }
}

// scatter map of the element blocks (once per hfx_cg_allocate): the binary search of every (row, column) pair is done here, not in every assembly
__global__ void cg_positions_kernel(long long nEntries, int nN, const int* __restrict__ cells, const long long* __restrict__ rowptr, const int* __restrict__ colidx,
                                    unsigned short* __restrict__ pos) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nEntries) return;
  const long long e = k / (nN * nN);
  const int ij = (int)(k - e * nN * nN), i = ij / nN, j = ij - i * nN;
  const int row = cells[e * nN + i], col = cells[e * nN + j];
  pos[k] = (unsigned short)(cg_find(rowptr, colidx, row, col) - rowptr[row]);
}

// Cells that are the affine image of the reference element, D = I, no convection: the Jacobian is constant, so
//   A = sum_rs C_rs K^_rs,  C = detJ Jinv^T Jinv,  K^_rs[i][j] = sum_ip w d_r phi_i d_s phi_j   (reference stiffness matrices, resident in shared memory)
//   M = detJ M^,  F_i = detJ sum_ip w phi_i f(x_ip)
// -- dim^2 multiply-adds per entry instead of a loop over the cubature points.  One warp per element, no barrier.
__global__ void __launch_bounds__(256) cg_affine_kernel(const CgParams p) {
  extern __shared__ __align__(16) double smca[];
  const int dim = p.dim, nN = p.nN, nIP = p.nIP, D2 = dim * dim, NN = nN * nN;
  const int nTab = D2 * NN + NN + nIP * nN;
  for (int i = threadIdx.x; i < nTab; i += blockDim.x) smca[i] = p.refTab[i];
  __syncthreads();
  const double* const KH = smca; const double* const MH = smca + D2 * NN; const double* const PW = MH + NN;
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int e = blockIdx.x * wpb + (threadIdx.x >> 5); e < p.nCells; e += gridDim.x * wpb) {
    if (!p.affine[e]) continue;
    const int* cell = p.cells + (size_t)e * nN;
    double J[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    {   // (loops with compile-time bounds: J, I and C stay in registers)
      const double* x0 = p.nodes + (size_t)cell[p.fv[0]] * dim;
#pragma unroll
      for (int r = 0; r < 3; r++)
        if (r < dim) {
          const double* xr = p.nodes + (size_t)cell[p.fv[r + 1]] * dim;
#pragma unroll
          for (int m = 0; m < 3; m++) if (m < dim) J[r][m] = 0.5 * (xr[m] - x0[m]);
        }
    }
    double det, I[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    if (dim == 1) { det = J[0][0]; I[0][0] = 1.0 / det; }
    else if (dim == 2) {
      det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
      const double id = 1.0 / det;
      I[0][0] = J[1][1] * id; I[0][1] = -J[0][1] * id; I[1][0] = -J[1][0] * id; I[1][1] = J[0][0] * id;
    } else det_inv(J, det, I);
    if (!(fabs(det) > 1e-300) && lane == 0) atomicOr(p.status, 1);
    double Cm[9];   // C(r, s) at Cm[3 r + s]
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int s2 = 0; s2 < 3; s2++) {
        double a = 0.0;
#pragma unroll
        for (int k = 0; k < 3; k++) if (k < dim) a = fma(I[k][r], I[k][s2], a);
        Cm[r * 3 + s2] = det * a;
      }
    const double dt = p.eulerDt;
    for (int idx = lane; idx < NN; idx += 32) {
      double acc = 0.0;
#pragma unroll
      for (int r = 0; r < 3; r++)
#pragma unroll
        for (int s2 = 0; s2 < 3; s2++) if (r < dim && s2 < dim) acc = fma(Cm[r * 3 + s2], KH[(r * dim + s2) * NN + idx], acc);
      if (dt > 0.0) acc = fma(dt, acc, det * MH[idx]);
      const int i = idx / nN;
      const long long at = p.pos ? p.rowptr[cell[i]] + p.pos[(size_t)e * NN + idx] : cg_find(p.rowptr, p.colidx, cell[i], cell[idx - i * nN]);
      atomicAdd(p.vals + at, acc);
    }
    if (p.srcIP || dt > 0.0)
      for (int i = lane; i < nN; i += 32) {
        double s = 0.0, mu = 0.0;
        if (p.srcIP) for (int ip = 0; ip < nIP; ip++) s = fma(PW[ip * nN + i], p.srcIP[(size_t)e * nIP + ip], s);
        if (dt > 0.0) for (int j = 0; j < nN; j++) mu = fma(MH[i * nN + j], p.solOld[cell[j]], mu);
        atomicAdd(p.rhs + cell[i], det * (dt > 0.0 ? fma(dt, s, mu) : s));
      }
  }
}

// C = detJ Jinv^T Jinv (row-major dim x dim) and detJ of every affine cell: the geometry does not change between assemblies (once per hfx_cg_allocate)
__global__ void cg_cell_geometry_kernel(int nCells, int nN, int dim, int fv0, int fv1, int fv2, int fv3, const double* __restrict__ nodes, const int* __restrict__ cells,
                                        const unsigned char* __restrict__ affine, double* __restrict__ geo /*[nCells][10]*/, int* __restrict__ status) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nCells || !affine[e]) return;
  const int fv[4] = {fv0, fv1, fv2, fv3};
  const int* cell = cells + (size_t)e * nN;
  double J[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  const double* x0 = nodes + (size_t)cell[fv[0]] * dim;
  for (int r = 0; r < dim; r++) { const double* xr = nodes + (size_t)cell[fv[r + 1]] * dim; for (int m = 0; m < dim; m++) J[r][m] = 0.5 * (xr[m] - x0[m]); }
  double det, I[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  if (dim == 1) { det = J[0][0]; I[0][0] = 1.0 / det; }
  else if (dim == 2) {
    det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    const double id = 1.0 / det;
    I[0][0] = J[1][1] * id; I[0][1] = -J[0][1] * id; I[1][0] = -J[1][0] * id; I[1][1] = J[0][0] * id;
  } else det_inv(J, det, I);
  if (!(fabs(det) > 1e-300)) atomicOr(status, 1);
  double* g = geo + (size_t)e * 10;
  for (int q = 0; q < 9; q++) g[q] = 0.0;
  for (int r = 0; r < dim; r++) for (int s2 = 0; s2 < dim; s2++) { double a = 0.0; for (int k = 0; k < dim; k++) a = fma(I[k][r], I[k][s2], a); g[r * dim + s2] = det * a; }
  g[9] = det;
}

// Gather form of the fast path: ONE WARP PER ROW (node).  The warp walks the cells the node belongs to in ascending order and adds each cell's row of
// sum_rs C_rs K^_rs (+ the Euler mass) into the row's image in shared memory at the precomputed positions, then stores the row once: no atomics, every entry summed in
// a fixed order (bit-reproducible), the matrix written exactly once.  The right-hand side entry of the node is reduced the same way.
__global__ void __launch_bounds__(256) cg_gather_kernel(const CgParams p) {
  extern __shared__ __align__(16) double smcq[];
  const int dim = p.dim, nN = p.nN, nIP = p.nIP, D2 = dim * dim, NN = nN * nN;
  const int nTab = D2 * NN + NN + nIP * nN;
  for (int i = threadIdx.x; i < nTab; i += blockDim.x) smcq[i] = p.refTab[i];
  __syncthreads();
  const double* const KH = smcq; const double* const MH = smcq + D2 * NN; const double* const PW = MH + NN;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  double* const buf = smcq + ((nTab + 1) & ~1) + (size_t)w * p.maxRow;
  const double dt = p.eulerDt;
  for (int r = blockIdx.x * wpb + w; r < p.nNodes; r += gridDim.x * wpb) {
    const long long r0 = p.rowptr[r];
    const int len = (int)(p.rowptr[r + 1] - r0);
    for (int k = lane; k < len; k += 32) buf[k] = 0.0;
    __syncwarp();
    double frhs = 0.0;
    for (long long k = p.n2c[r]; k < p.n2c[r + 1]; k++) {
      const int e = p.n2cCell[k];
      if (!p.affine[e]) continue;                                  // (warp-uniform) curved cells come afterwards, through cg_element_kernel
      const int i = p.n2cLoc[k];
      const double* g = p.cellGeo + (size_t)e * 10;
      double Cm[9];   // C(r, s) at Cm[3 r + s] (compile-time indices: registers)
#pragma unroll
      for (int r2 = 0; r2 < 3; r2++)
#pragma unroll
        for (int s2 = 0; s2 < 3; s2++) Cm[r2 * 3 + s2] = (r2 < dim && s2 < dim) ? g[r2 * dim + s2] : 0.0;
      const double det = g[9];
      const unsigned short* ps = p.pos + (size_t)e * NN + i * nN;
      for (int j = lane; j < nN; j += 32) {
        double acc = 0.0;
#pragma unroll
        for (int r2 = 0; r2 < 3; r2++)
#pragma unroll
          for (int s2 = 0; s2 < 3; s2++) if (r2 < dim && s2 < dim) acc = fma(Cm[r2 * 3 + s2], KH[(r2 * dim + s2) * NN + i * nN + j], acc);
        if (dt > 0.0) acc = fma(dt, acc, det * MH[i * nN + j]);
        buf[ps[j]] += acc;                                           // the nodes of a cell are distinct: no two lanes share a position
      }
      if (p.srcIP || dt > 0.0) {
        const int* cell = p.cells + (size_t)e * nN;
        double s = 0.0, mu = 0.0;
        if (p.srcIP) for (int ip = lane; ip < nIP; ip += 32) s = fma(PW[ip * nN + i], p.srcIP[(size_t)e * nIP + ip], s);
        if (dt > 0.0) for (int j = lane; j < nN; j += 32) mu = fma(MH[i * nN + j], p.solOld[cell[j]], mu);
        double v = det * (dt > 0.0 ? fma(dt, s, mu) : s);
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        frhs += v;
      }
      __syncwarp();
    }
    for (int k = lane; k < len; k += 32) p.vals[r0 + k] = buf[k];
    if (lane == 0 && (p.srcIP || dt > 0.0)) p.rhs[r] = frhs;
    __syncwarp();
  }
}

// affine image of the reference element? (same test as elem_affine_kernel of the HDG path, on the node-based mesh)
__global__ void cg_affine_flags_kernel(int nCells, int nN, int dim, int fv0, int fv1, int fv2, int fv3, const double* __restrict__ nodes, const int* __restrict__ cells,
                                       const double* __restrict__ bary, unsigned char* __restrict__ affine) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nCells) return;
  const int fv[4] = {fv0, fv1, fv2, fv3};
  const int* cell = cells + (size_t)e * nN;
  double V[4][3];
  for (int v = 0; v <= dim; v++) for (int m = 0; m < dim; m++) V[v][m] = nodes[(size_t)cell[fv[v]] * dim + m];
  double h = 0.0;
  for (int v = 1; v <= dim; v++) for (int m = 0; m < dim; m++) h = fmax(h, fabs(V[v][m] - V[0][m]));
  bool ok = true;
  for (int i = 0; i < nN && ok; i++)
    for (int m = 0; m < dim; m++) {
      double s = 0.0;
      for (int v = 0; v <= dim; v++) s = fma(bary[i * (dim + 1) + v], V[v][m], s);
      if (!(fabs(s - nodes[(size_t)cell[i] * dim + m]) <= 1e-13 * h)) ok = false;
    }
  affine[e] = ok ? 1 : 0;
}

inline size_t cg_smem_bytes(int dim, int nN, int nIP) {
  const size_t D2 = (size_t)dim * dim;
  const size_t d = (((size_t)nN * dim + 1) & ~(size_t)1) + 2 * (((size_t)nIP * D2 + 1) & ~(size_t)1) + 2 * (((size_t)nIP + 1) & ~(size_t)1) + (((size_t)nIP * dim + 1) & ~(size_t)1) + (size_t)nIP * nN * dim;
  return d * 8 + (size_t)((nN + 1) & ~1) * 4 + 16;
}

// DirichletModel through CGSolver (CGSolver.cpp:139-243): the rows of the nodes of every boundary face are zeroed (zeroOutRows), then the face's identity block and its
// Dirichlet values (face-node order) are Set
__global__ void cg_dirichlet_kernel(int nFaces, int nNf, const uint8_t* __restrict__ faceBC, const int* __restrict__ faces, const double* __restrict__ dirichlet,
                                    const long long* __restrict__ rowptr, const int* __restrict__ colidx, double* __restrict__ vals, double* __restrict__ rhs) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= (long long)nFaces * nNf) return;
  const int F = (int)(k / nNf);
  if (faceBC[F] != 1) return;
  const int n = faces[k];
  for (long long q = rowptr[n]; q < rowptr[n + 1]; q++) vals[q] = colidx[q] == n ? 1.0 : 0.0;
  rhs[n] = dirichlet[k];
}

}  // namespace hfx
