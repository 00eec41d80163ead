// General per-element HDG kernel (runtime sizes): any simplex or orthotope dimension / order the reference element supports, any number of
// DOFs per node, every in-scope operator including the Newton-linearised HDGUNabU (HDGBurgersModel).  It covers what the fused
// shared-memory kernel (hfx_assemble.cuh) has no instantiation for: 3-D orders 4-5, nDOFsPerNode > 1, HDGUNabU.
//
// One CTA per element pass; the dense local system lives in a per-CTA global scratch (L2 resident), the two small inverses
// (M: nN x nN, K: u x u) in shared memory.  Same reference semantics as the fused kernel:
//   geometry          src/operator/Operator.cpp:14-84, src/model/HDGModel.cpp:53-85, src/operator/HDGBase.cpp:18-65
//   operators         HDGBase.cpp:67-158, HDGDiffusion.cpp:31-145, HDGConvection.cpp:31-104, Reaction.cpp, Source.cpp,
//                     HDGUNabU.cpp:27-191, Euler.cpp:18-37 (+ hook HDGModel.cpp:38-47)
//   models            HDGLaplaceModel / HDGDiffusionSource / HDGConvectionDiffusionReactionSource / HDGBurgersModel (computeLocal*)
//   condensation      src/solver/HDGSolver.cpp:331-348, with S_qq = M (x) I_{dim*nDOF} (HDGBase.cpp:152) exploited for the q-block
//   boundary + scatter HDGSolver.cpp:361-529 (CGType models), :531-675
// Every matrix entry is produced by exactly one thread (gather form): no atomics inside the element, bit-reproducible.
#pragma once
#include "hfx_assemble.cuh"

namespace hfx {

struct GenParams {
  AsmParams a;
  int dim, nN, nNf, nFc, nIP, nIPf, nD;
  const double* bufSol;      // BufferSolution, cell field [nCells][nN][nD]        (HDGBurgersModel.cpp:87-124)
  const double* tracePrev;   // Trace of the previous iterate, face field [nFaces][nNf][nD]
  int frameV[4];             // local ids of the vertices spanning the affine frame of a straight-sided element (and, first dim of them, of a face):
                             // 0,1,2,3 for simplices, 0,1,3,4 for orthotopes (ReferenceElement.cpp:885-943)
  int forcePivot;            // partial pivoting in K^-1 whatever the model (fallback after a vanishing pivot in the unpivoted path)
  int nSrc;                  // source components: 1, or dim for the Burgers model (HDGBurgersModel.cpp:112-122)
  int smOpt[6];              // offsets (doubles, from the optional area) of Um, Rm, Aq, Bq, shape table, GM when they live in shared memory; -1: global scratch
  double* ws;                // per-CTA scratch
  long long wsStride;        // doubles per CTA
  // RungeKutta::apply (src/operator/RungeKutta.cpp:90-143), auxiliary fields {Flux, Trace}: time scheme code 2
  int rkStage, rkNumStages;
  double rkRow[8];           // Butcher row of the current stage (a_s0 .. a_s,nStages-1)
  const double* oldSol; const double* oldFlux; const double* oldTrace;          // OldSolution / OldFlux (cell), OldTrace (face)
  const double* rkSol[8]; const double* rkFlux[8]; const double* rkTrace[8];   // RKStage_k, RKStage_Flux_k (cell), RKStage_Trace_k (face)  // per-element Model surface (FEModel::compute / getLocalMatrix / getLocalRHS, src/model/FEModel.h:43-78): build the local system of ONE element,
  // write it out dense (column-major n x n, S_qq = M (x) I included) with its right-hand side, and stop before the condensation
  int dumpElem; double* dumpA; double* dumpF;
  // HDGSolverOpts.type = WEXPLICIT / SEXPLICIT (HDGSolver.cpp:346-354): S = S_ll, S0 = F_l - S_lu sol - S_lq flux with the element's current Solution / Flux (cell fields)
  int explicitS; const double* solCur; const double* fluxCur;
};

// scratch layout (offsets in doubles), identical on host and device
struct GenWs {
  int u, q, l, n, t, nJ, nFf, dd, sQ, sL;
  int oLm, oF, oGM, oDV, oIJ, oNRM, oTAUS, oDIP, oVIP, oVDN, oFS, oTDN, oSIP, oDIVS, oX, oTAUn, oDN, oVN, oSOL, oTR, oSOLD, oMM, oW, oFT, oFCN,
      oFNd, oFDN, oFONE, oBUU, oAq, oBq, oRm, oUm, oQm, oLW, total;
  __host__ __device__ GenWs(int dim, int nN, int nNf, int nFc, int nIP, int nIPf, int nD) {
    u = nN * nD; q = u * dim; t = nNf * nD; l = nFc * t; n = u + q + l; nJ = nIP + nFc * nIPf; nFf = nFc * nIPf; dd = dim * dim; sQ = u; sL = u + q;
    int o = 0;
    auto take = [&](int k) { int r = o; o += (k + 1) & ~1; return r; };
    oLm = take(n * n); oF = take(n); oGM = take(nIP * nN * dim); oDV = take(nJ); oIJ = take(nIP * dd);
    oNRM = take(nFf * dim); oTAUS = take(nFf * nD * nD); oDIP = take(nJ * dd); oVIP = take(nIP * dim);
    oVDN = take(nFf); oFS = take(nFf * nD); oTDN = take(nFf); oSIP = take(nIP * nD); oDIVS = take(nIP);
    oX = take(nN * dim); oTAUn = take(nFc * nNf * nD * nD); oDN = take(nN * dd); oVN = take(nN * dim);
    oSOL = take(u); oTR = take(l); oSOLD = take(u); oMM = take(nN * nN); oW = take(nN * nN);
    oFT = take(nFc * t * t); oFCN = take(nFc * t * t); oFNd = take(nFc * dim * nNf * nNf);
    oFDN = take(nFc * dim * nNf * nNf); oFONE = take(nFc * nNf * nNf); oBUU = take(u * u);
    oAq = take(q * u); oBq = take(q * (l + 1)); oRm = take(u * (l + 1)); oUm = take(u * (l + 1));
    oQm = take(q * (l + 1)); oLW = take(2 * nIP + nIP * nD);
    total = o;
  }
};

constexpr int kGenThreads = 256;
#define HFX_GPROF(i) do { if (p.prof && blockIdx.x == 0 && tid == 0) { long long c_ = clock64(); p.prof[i] += c_ - tprev; tprev = c_; } } while (0)

// Warp task on the FP64 tensor cores with run-time sizes: rows [8 mt, 8 mt + 8) x NTW column tiles of 8 starting at tile nt0 of the
// M x N product with reduction length K.  fa(m, k) / fb(k, n) return operand entries and are only called with IN-RANGE indices
// (rows / columns / reduction steps beyond the matrix are clamped here; a clamped reduction step gets an exact zero left operand),
// so they are straight-line loads: the operands of four reduction steps are in flight before the first DMMA of a group, which is
// what hides the L2 latency of operands that live in the per-CTA scratch.  fs(m, n, v0, v1) receives C[m][n], C[m][n+1] (n even)
// and does its own range checks.
template <int NTW, class FA, class FB, class FS>
__device__ __forceinline__ void mma_task_rt(int mt, int nt0, int lane, int M, int N, int K, FA fa, FB fb, FS fs) {
  const int lr = lane >> 2, lc = lane & 3;
  const int m = mt * 8 + lr, mc = m < M ? m : M - 1;
  int nc[NTW];
#pragma unroll
  for (int j = 0; j < NTW; j++) { const int nn = (nt0 + j) * 8 + lr; nc[j] = nn < N ? nn : N - 1; }
  double c[NTW][2];
#pragma unroll
  for (int j = 0; j < NTW; j++) { c[j][0] = 0.0; c[j][1] = 0.0; }
  for (int k0 = 0; k0 < K; k0 += 16) {
    double a[4], b[4][NTW];
#pragma unroll
    for (int s = 0; s < 4; s++) {
      const int k = k0 + 4 * s + lc, kc = k < K ? k : K - 1;
      const double av = fa(mc, kc);
      a[s] = k < K ? av : 0.0;
#pragma unroll
      for (int j = 0; j < NTW; j++) b[s][j] = fb(kc, nc[j]);
    }
#pragma unroll
    for (int s = 0; s < 4; s++)
#pragma unroll
      for (int j = 0; j < NTW; j++) dmma(c[j], a[s], b[s][j]);
  }
#pragma unroll
  for (int j = 0; j < NTW; j++) fs(m, (nt0 + j) * 8 + 2 * lc, c[j][0], c[j][1]);
}

// Gauss-Jordan inverse with partial pivoting of the left half of the row-major n x 2n matrix aug (right half = identity on entry,
// inverse on exit).  Whole CTA.  scr: n + 2n doubles; ipiv: 1 int in shared memory.
__device__ inline void cta_invert(double* aug, int n, double* scr, int* ipiv, int* status) {
  const int tid = threadIdx.x, NT = blockDim.x, n2 = 2 * n;
  double* fac = scr; double* rk = scr + n;
  for (int k = 0; k < n; k++) {
    if (tid < 32) {   // pivot search by warp 0
      double best = -1.0; int bi = k;
      for (int i = k + tid; i < n; i += 32) { const double v = fabs(aug[i * n2 + k]); if (v > best) { best = v; bi = i; } }
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (tid == 0) { *ipiv = bi; if (!(best > 1e-300)) atomicOr(status, 1); }
    }
    __syncthreads();
    const int pv = *ipiv;
    if (pv != k) for (int j = tid; j < n2; j += NT) { const double a = aug[k * n2 + j]; aug[k * n2 + j] = aug[pv * n2 + j]; aug[pv * n2 + j] = a; }
    __syncthreads();
    const double ip = 1.0 / aug[k * n2 + k];
    for (int j = tid; j < n2; j += NT) rk[j] = aug[k * n2 + j] * ip;
    for (int i = tid; i < n; i += NT) fac[i] = aug[i * n2 + k];
    __syncthreads();
    for (int idx = tid; idx < n * n2; idx += NT) {
      const int i = idx / n2, j = idx - i * n2;
      aug[idx] = (i == k) ? rk[j] : fma(-fac[i], rk[j], aug[idx]);
    }
    __syncthreads();
  }
}

// Unpivoted in-place-style Gauss-Jordan inverse of a row-major n x n matrix, ping-pong between two buffers, one barrier per step, no
// integer division.  For the definite blocks (mass matrix; K of the linear models, as the fused kernel does); a vanishing pivot raises
// bit 0 of *status.  Returns the buffer that holds the inverse.
__device__ inline double* cta_invert_np(double* b0, double* b1, int n, int* status) {
  double* src = b0; double* dst = b1;
  const double scale = fabs(b0[0]);
  const float rn = 1.0f / (float)n;
  bool bad = false;
  for (int k = 0; k < n; k++) {
    const double piv = src[k * n + k];
    if (!(fabs(piv) > 1e-14 * scale)) bad = true;
    const double ip = 1.0 / piv;
    const double* __restrict__ sp = src; double* __restrict__ dp = dst;   // distinct buffers: the loads of all entries may run ahead of the stores
#pragma unroll 2
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
      const int i = __float2int_rd(((float)idx + 0.5f) * rn), j = idx - i * n;   // exact for idx < 2^22
      const double pkj = sp[k * n + j], fik = sp[i * n + k] * ip;
      double v;
      if (i == k) v = (j == k) ? ip : pkj * ip;
      else v = (j == k) ? -fik : fma(-fik, pkj, sp[idx]);
      dp[idx] = v;
    }
    __syncthreads();
    double* tsw = dst; dst = src; src = tsw;
  }
  if (bad && threadIdx.x == 0) atomicOr(status, 1);
  return src;
}

#ifndef HFX_GEN_MINBLOCKS
#define HFX_GEN_MINBLOCKS 2
#endif
__global__ void __launch_bounds__(kGenThreads, HFX_GEN_MINBLOCKS) hdg_generic_kernel(const GenParams P) {
  const AsmParams& p = P.a;
  const int dim = P.dim, nN = P.nN, nNf = P.nNf, nFc = P.nFc, nIP = P.nIP, nIPf = P.nIPf, nD = P.nD;
  const GenWs z(dim, nN, nNf, nFc, nIP, nIPf, nD);
  const int u = z.u, q = z.q, l = z.l, n = z.n, t = z.t, nJ = z.nJ, nFf = z.nFf, dd = z.dd, sQ = z.sQ, sL = z.sL, sT = nD * nD;
  const int tid = threadIdx.x, NT = kGenThreads;
  const bool hasDiff = p.opmask & 1, hasConv = p.opmask & 2, hasReac = (p.opmask & 4) && p.reacIP, hasSrc = (p.opmask & 8) && P.a.srcIP, hasUN = p.opmask & 16;
  const bool euler = p.timeScheme == 1;
  const bool diffField = hasDiff && p.diffComps > 0 && p.diff;
  double* ws = P.ws + (size_t)blockIdx.x * P.wsStride;
  double* Lm = ws + z.oLm; double* Fv = ws + z.oF; double* GM = ws + z.oGM; double* DV = ws + z.oDV; double* IJ = ws + z.oIJ; double* NRM;
  double* TAUS; double* DIP = ws + z.oDIP; double* VIP = ws + z.oVIP; double* VDN; double* FS; double* TDN;
  double* SIP = ws + z.oSIP; double* DIVS = ws + z.oDIVS; double* X = ws + z.oX; double* TAUn = ws + z.oTAUn; double* DN = ws + z.oDN; double* VN = ws + z.oVN;
  double* SOL = ws + z.oSOL; double* TR = ws + z.oTR; double* SOLD = ws + z.oSOLD; double* MM = ws + z.oMM; double* FT = ws + z.oFT;
  double* FCN = ws + z.oFCN; double* FNd = ws + z.oFNd; double* FDN = ws + z.oFDN; double* FONE = ws + z.oFONE; double* BUU = ws + z.oBUU; double* Aq = ws + z.oAq;
  double* Bq = ws + z.oBq; double* Rm = ws + z.oRm; double* Um = ws + z.oUm; double* LW = ws + z.oLW;
  // shared: augmented matrix for the two inverses, scratch rows, integer maps
  extern __shared__ __align__(16) double gsm[];
  const int nmax = u > nN ? u : nN;
  double* AUG = gsm;                                 // [nmax][2 nmax]
  double* SCR = AUG + (size_t)nmax * 2 * nmax;       // [3 nmax]
  // per-face-cubature-point weights and the face shape table: read nNf^2 times each by the face-matrix loop
  double* FSHs = SCR + 3 * nmax + 2;                 // [nIPf][nNf]
  double* DVF = FSHs + ((nIPf * nNf + 1) & ~1);      // [nFf] dV
  VDN = DVF + nFf; TDN = VDN + nFf;                  // [nFf] dV v.n ; [nFf] dV (trace . n)
  NRM = TDN + nFf;                                   // [nFf][dim]
  double* DNV = NRM + nFf * dim;                     // [nFf][dim] (D n)
  TAUS = DNV + nFf * dim;                            // [nFf][nD*nD]
  FS = TAUS + nFf * sT;                              // [nFf][nD]
  long long* ROWS = reinterpret_cast<long long*>(FS + ((nFf * nD + 1) & ~1));   // [nFc] first entry of row (F,0) in vals
  int* PERM = reinterpret_cast<int*>(ROWS + nFc);    // [nFc*nNf]
  int* NIF = PERM + nFc * nNf;                       // [nFc*nN]
  int* FNo = NIF + nFc * nN;                         // [nFc*nNf]
  int* FACE = FNo + nFc * nNf;                       // [nFc] global face ids
  int* BCF = FACE + nFc; int* INTF = BCF + nFc; int* POS = INTF + nFc;   // [nFc], [nFc], [nFc*nFc]
  int* RLEN = POS + nFc * nFc; int* OPP = RLEN + nFc; int* IPIV = OPP + nFc;
  int* COLOFF = IPIV + 2;                            // [nFc][l] offset of element-local column c inside the block row of face f
  long long* ROWOFF = reinterpret_cast<long long*>((reinterpret_cast<uintptr_t>(COLOFF + nFc * l) + 15) & ~(uintptr_t)15);   // [l] first entry of element-local trace row r
  // the big operands of the condensation products live in shared memory when they fit (decided on the host, largest benefit first)
  double* OPT = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(ROWOFF + l) + 15) & ~(uintptr_t)15);
  if (P.smOpt[0] >= 0) Um = OPT + P.smOpt[0];
  if (P.smOpt[1] >= 0) Rm = OPT + P.smOpt[1];
  if (P.smOpt[2] >= 0) Aq = OPT + P.smOpt[2];
  if (P.smOpt[3] >= 0) Bq = OPT + P.smOpt[3];
  double* const Qm = Bq;                             // Q = -A U - B overwrites B entry by entry
  const double* SHP = p.shape;
  if (P.smOpt[4] >= 0) { double* d = OPT + P.smOpt[4]; for (int i = tid; i < nIP * nN; i += NT) d[i] = p.shape[i]; SHP = d; }
  if (P.smOpt[5] >= 0) GM = OPT + P.smOpt[5];

  for (int i = tid; i < nIPf * nNf; i += NT) FSHs[i] = p.fshape[i];
  for (int i = tid; i < nFc * nNf; i += NT) FNo[i] = p.faceNodes[i];
  for (int i = tid; i < nFc * nN; i += NT) NIF[i] = p.nodeInFace[i];
  if (tid < nFc) { int vn = 0; for (int kk = 0; kk < nN; kk++) if (p.nodeInFace[tid * nN + kk] < 0) { vn = kk; break; } OPP[tid] = vn; }
  __syncthreads();

  long long tprev = clock64();
  const bool dump = P.dumpA != nullptr;
  for (int e = dump ? P.dumpElem : blockIdx.x; e < (dump ? P.dumpElem + 1 : p.nCells); e += gridDim.x) {
    // ---- gather (HDGSolver.cpp:231-326) ----------------------------------------------------------------------------------------
    const int* cell = p.cells + (size_t)e * nN;
    for (int i = tid; i < nN * dim; i += NT) X[i] = p.elemX[(size_t)e * nN * dim + i];
    for (int i = tid; i < nFc * nNf; i += NT) PERM[i] = p.fperm[(size_t)e * nFc * nNf + i];
    if (tid < nFc) {
      const int F = p.cell2face[(size_t)e * nFc + tid];
      FACE[tid] = F; ROWS[tid] = p.faceRowStart[F]; RLEN[tid] = (int)p.faceNnb[F] * t; BCF[tid] = p.faceBC[F]; INTF[tid] = p.faceInterior[F];
    }
    for (int i = tid; i < nFc * nFc; i += NT) POS[i] = p.elemPos[(size_t)e * nFc * nFc + i];
    __syncthreads();
    for (int i = tid; i < nFc * l; i += NT) {   // block CSR: neighbour block of face f2 in the row of face f, column inside the t x t block
      const int f = i / l, c = i - f * l, f2 = c / t, cc = c - f2 * t, b = cc / nD, k2 = cc - b * nD;
      COLOFF[i] = POS[f * nFc + f2] * t * t + PERM[f2 * nNf + b] * nD + k2;
    }
    for (int r = tid; r < l; r += NT) { const int f = r / t, rr = r - f * t, a = rr / nD, k1 = rr - a * nD; ROWOFF[r] = ROWS[f] + (long long)(PERM[f * nNf + a] * nD + k1) * t; }
    for (int i = tid; i < nFc * nNf * sT; i += NT) {   // Tau: side selection :277-304 then permutation :306-326
      const int fa = i / sT, c = i - fa * sT, f = fa / nNf;
      const int side = (p.tauVals == 2 * sT) ? p.tauSide[(size_t)e * nFc + f] : 0;
      TAUn[i] = p.tau[((size_t)FACE[f] * nNf + PERM[fa]) * p.tauVals + side * sT + c];
    }
    if (diffField) for (int i = tid; i < nN * dd; i += NT) {
      const int nd = i / dd, c = i - nd * dd;
      const size_t ent = p.diffIsCell ? ((size_t)e * nN + nd) : (size_t)cell[nd];
      DN[i] = (p.diffComps == 1) ? (((c / dim) == (c % dim)) ? p.diff[ent] : 0.0) : p.diff[ent * dd + c];
    }
    if (hasConv) for (int i = tid; i < nN * dim; i += NT) VN[i] = p.vel[(size_t)cell[i / dim] * dim + (i % dim)];
    if (hasUN) {
      for (int i = tid; i < u; i += NT) SOL[i] = P.bufSol[(size_t)e * u + i];
      for (int i = tid; i < l; i += NT) { const int fa = i / nD, k = i - fa * nD, f = fa / nNf; TR[i] = P.tracePrev[((size_t)FACE[f] * nNf + PERM[fa]) * nD + k]; }
    }
    if (euler) for (int i = tid; i < u; i += NT) SOLD[i] = p.solOld[(size_t)e * u + i];
    {   // zero the blocks of the local matrix that are read later (column-major, leading dimension n); the Sqq block is neither stored nor read
      const int lane_ = tid & 31, warp_ = tid >> 5, nw_ = NT / 32;
      for (int c = warp_; c < n; c += nw_) {
        double* col = Lm + n * c;
        const bool qcol = c >= sQ && c < sL;
        for (int r = lane_; r < n; r += 32) if (!(qcol && r >= sQ && r < sL)) col[r] = 0.0;
      }
    }
    for (int i = tid; i < n; i += NT) Fv[i] = 0.0;
    __syncthreads();

    HFX_GPROF(0);
    const bool affE = p.affine && p.affine[e];
    // ---- geometry and coefficients at the cubature points -------------------------------------------------------------------------
    for (int k = tid; k < nJ; k += NT) {
      if (k < nIP) {
        const int ip = k;
        double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        if (affE) {   // straight-sided element: constant Jacobian straight from the vertices
          for (int r = 0; r < dim; r++) for (int m = 0; m < dim; m++) J[r][m] = 0.5 * (X[P.frameV[r + 1] * dim + m] - X[P.frameV[0] * dim + m]);
        } else for (int i = 0; i < nN; i++) {
          const double* d = p.dshape + (ip * nN + i) * dim;
          for (int r = 0; r < dim; r++) for (int m = 0; m < dim; m++) J[r][m] = fma(d[r], X[i * dim + m], J[r][m]);
        }
        double det, I[3][3];
        if (dim == 2) {
          det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
          const double id = 1.0 / det;
          I[0][0] = J[1][1] * id; I[0][1] = -J[0][1] * id; I[1][0] = -J[1][0] * id; I[1][1] = J[0][0] * id;
        } else {
          const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1], c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2], c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
          det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
          const double id = 1.0 / det;
          I[0][0] = c00 * id; I[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id; I[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
          I[1][0] = c01 * id; I[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id; I[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
          I[2][0] = c02 * id; I[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id; I[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
        }
        const double dv = p.w[ip] * det;
        DV[ip] = dv;
        for (int m = 0; m < dim; m++) for (int r = 0; r < dim; r++) IJ[ip * dd + m * dim + r] = I[m][r];   // invJ(m,r): x_m <- xi_r
        for (int c = 0; c < dd; c++) {
          double s = ((c / dim) == (c % dim)) ? (diffField ? 1.0 : p.diffConst) : 0.0;
          if (diffField) { s = 0.0; for (int i = 0; i < nN; i++) s = fma(p.shape[ip * nN + i], DN[i * dd + c], s); }
          DIP[ip * dd + c] = s;
        }
        if (hasConv) for (int d = 0; d < dim; d++) { double s = 0.0; for (int i = 0; i < nN; i++) s = fma(p.shape[ip * nN + i], VN[i * dim + d], s); VIP[ip * dim + d] = s; }
        if (hasUN) for (int k2 = 0; k2 < nD; k2++) { double s = 0.0; for (int i = 0; i < nN; i++) s = fma(SOL[i * nD + k2], p.shape[ip * nN + i], s); SIP[ip * nD + k2] = s; }
        LW[ip] = hasReac ? p.reacIP[(size_t)e * nIP + ip] * dv : 0.0;
      } else {
        const int fi = k - nIP, f = fi / nIPf, ip = fi - f * nIPf;
        const int* fn = FNo + f * nNf;
        double J[2][3] = {{0, 0, 0}, {0, 0, 0}};
        if (affE) {
          for (int r = 0; r < dim - 1; r++) for (int m = 0; m < dim; m++) J[r][m] = 0.5 * (X[fn[P.frameV[r + 1]] * dim + m] - X[fn[P.frameV[0]] * dim + m]);
        } else for (int a = 0; a < nNf; a++) {
          const double* d = p.fdshape + (ip * nNf + a) * (dim - 1);
          for (int r = 0; r < dim - 1; r++) for (int m = 0; m < dim; m++) J[r][m] = fma(d[r], X[fn[a] * dim + m], J[r][m]);
        }
        double nv[3] = {0, 0, 0}, area;
        if (dim == 2) { nv[0] = -J[0][1]; nv[1] = J[0][0]; area = sqrt(J[0][0] * J[0][0] + J[0][1] * J[0][1]); }
        else {
          nv[0] = J[0][1] * J[1][2] - J[0][2] * J[1][1]; nv[1] = J[0][2] * J[1][0] - J[0][0] * J[1][2]; nv[2] = J[0][0] * J[1][1] - J[0][1] * J[1][0];
          const double g00 = J[0][0] * J[0][0] + J[0][1] * J[0][1] + J[0][2] * J[0][2], g11 = J[1][0] * J[1][0] + J[1][1] * J[1][1] + J[1][2] * J[1][2];
          const double g01 = J[0][0] * J[1][0] + J[0][1] * J[1][1] + J[0][2] * J[1][2];
          area = sqrt(g00 * g11 - g01 * g01);   // sqrt(det(J J^T)) (Operator.cpp:66-69)
        }
        double nrm = 0.0;
        for (int m = 0; m < dim; m++) nrm = fma(nv[m], nv[m], nrm);
        nrm = sqrt(nrm);
        double prod = 0.0;
        for (int m = 0; m < dim; m++) { nv[m] /= nrm; prod = fma(X[OPP[f] * dim + m] - X[fn[0] * dim + m], nv[m], prod); }   // HDGBase.cpp:43-62
        if (prod > 0.0) for (int m = 0; m < dim; m++) nv[m] = -nv[m];
        const double dvf = p.fw[ip] * area;
        DV[nIP + fi] = dvf; DVF[fi] = dvf;
        for (int m = 0; m < dim; m++) NRM[fi * dim + m] = nv[m];
        for (int c = 0; c < sT; c++) { double s = 0.0; for (int a = 0; a < nNf; a++) s = fma(TAUn[(f * nNf + a) * sT + c], FSHs[ip * nNf + a], s); TAUS[fi * sT + c] = s; }
        for (int c = 0; c < dd; c++) {
          double s = ((c / dim) == (c % dim)) ? (diffField ? 1.0 : p.diffConst) : 0.0;
          if (diffField) { s = 0.0; for (int a = 0; a < nNf; a++) s = fma(FSHs[ip * nNf + a], DN[fn[a] * dd + c], s); }
          DIP[(nIP + fi) * dd + c] = s;
        }
        double vdn = 0.0;
        if (hasConv) for (int d = 0; d < dim; d++) { double s = 0.0; for (int a = 0; a < nNf; a++) s = fma(FSHs[ip * nNf + a], VN[fn[a] * dim + d], s); vdn = fma(s, nv[d], vdn); }
        VDN[fi] = dvf * vdn;
        for (int d = 0; d < dim; d++) {   // (D n)_d, D col-major
          double dn = 0.0;
          for (int b2 = 0; b2 < dim; b2++) dn = fma(DIP[(nIP + fi) * dd + b2 * dim + d], nv[b2], dn);
          DNV[fi * dim + d] = dn;
        }
        if (hasUN) {   // HDGUNabU.cpp:107-124
          double tdn = 0.0;
          for (int d = 0; d < dim; d++) {
            double tr = 0.0, fs = 0.0;
            for (int a = 0; a < nNf; a++) { tr = fma(TR[(f * nNf + a) * nD + d], FSHs[ip * nNf + a], tr); fs = fma(SOL[fn[a] * nD + d], FSHs[ip * nNf + a], fs); }
            tdn = fma(tr, nv[d], tdn);
            FS[fi * nD + d] = fs;
          }
          TDN[fi] = tdn * dvf;
        }
      }
    }
    __syncthreads();
    HFX_GPROF(1);
    // physical gradients gm(d,i) = (J^-1 grad_ref phi_i)_d at the bulk points
    for (int idx = tid; idx < nIP * nN; idx += NT) {
      const int ip = idx / nN, i = idx - ip * nN;
      const double* dp = p.dshape + (ip * nN + i) * dim;
      for (int d = 0; d < dim; d++) { double s = 0.0; for (int r = 0; r < dim; r++) s = fma(IJ[ip * dd + d * dim + r], dp[r], s); GM[(ip * nN + i) * dim + d] = s; }
    }
    __syncthreads();
    if (hasUN) for (int ip = tid; ip < nIP; ip += NT) {   // div of the previous iterate (HDGUNabU.cpp:153-165)
      double dv = 0.0;
      for (int i = 0; i < nN; i++) for (int d = 0; d < dim; d++) dv = fma(GM[(ip * nN + i) * dim + d], SOL[i * nD + d], dv);
      DIVS[ip] = dv;
    }
    // ---- weighted face matrices (gather form of the face loops of HDGBase / HDGDiffusion / HDGConvection / HDGUNabU) -----------
    // scalar problems without the Newton-linearised block: one tensor-core product, rows = node pairs (a,b), reduction over the face
    // cubature points, columns = (face, weight kind): 1, v.n, n_d, (D n)_d, tau
    if (nD == 1 && !hasUN) {
      const int lane = tid & 31, warp = tid >> 5, NWARP = NT / 32;
      const int NK = 3 + 2 * dim, NCOL = nFc * NK, MR = nNf * nNf, MT = (MR + 7) / 8, NG = (NCOL + 23) / 24;
      for (int task = warp; task < MT * NG; task += NWARP) {
        mma_task_rt<3>(task % MT, (task / MT) * 3, lane, MR, NCOL, nIPf,
            [&](int m, int ip) { const int a = m / nNf, b = m - a * nNf; return FSHs[ip * nNf + a] * FSHs[ip * nNf + b]; },
            [&](int ip, int c) {
              const int f = c / NK, kind = c - f * NK, fi = f * nIPf + ip;
              const double dv = DVF[fi];
              if (kind == 0) return dv;
              if (kind == 1) return VDN[fi];
              if (kind < 2 + dim) return dv * NRM[fi * dim + kind - 2];
              if (kind < 2 + 2 * dim) return dv * DNV[fi * dim + kind - 2 - dim];
              return dv * TAUS[fi];
            },
            [&](int m, int c0, double v0, double v1) {
              if (m >= MR) return;
              const int a = m / nNf, b = m - a * nNf;
#pragma unroll
              for (int h = 0; h < 2; h++) {
                const int c = c0 + h;
                if (c >= NCOL) continue;
                const double v = h ? v1 : v0;
                const int f = c / NK, kind = c - f * NK;
                if (kind == 0) FONE[f * MR + m] = v;
                else if (kind == 1) FCN[f * t * t + a + t * b] = hasConv ? v : 0.0;
                else if (kind < 2 + dim) FNd[(f * dim + kind - 2) * MR + m] = v;
                else if (kind < 2 + 2 * dim) FDN[(f * dim + kind - 2 - dim) * MR + m] = v;
                else FT[f * t * t + a + t * b] = v;
              }
            });
      }
    } else
    for (int idx = tid; idx < nFc * nNf * nNf; idx += NT) {
      const int f = idx / (nNf * nNf), ab = idx - f * nNf * nNf, a = ab / nNf, b = ab - a * nNf;
      double one = 0.0, cn = 0.0, tdn = 0.0, nd3[3] = {0, 0, 0}, dn3[3] = {0, 0, 0}, tt[9], fsn[9];
      for (int c = 0; c < sT; c++) { tt[c] = 0.0; fsn[c] = 0.0; }
      for (int ip = 0; ip < nIPf; ip++) {
        const int fi = f * nIPf + ip;
        const double ss = FSHs[ip * nNf + a] * FSHs[ip * nNf + b], dv = DVF[fi], w = dv * ss;
        one += w;
        cn = fma(VDN[fi], ss, cn);
        for (int d = 0; d < dim; d++) {
          nd3[d] = fma(w, NRM[fi * dim + d], nd3[d]);
          dn3[d] = fma(w, DNV[fi * dim + d], dn3[d]);
        }
        for (int c = 0; c < sT; c++) tt[c] = fma(w, TAUS[fi * sT + c], tt[c]);
        if (hasUN) {
          tdn = fma(TDN[fi], ss, tdn);
          for (int k1 = 0; k1 < nD; k1++) for (int k2 = 0; k2 < nD; k2++) fsn[k2 * nD + k1] = fma(FS[fi * nD + k1] * NRM[fi * dim + k2], w, fsn[k2 * nD + k1]);
        }
      }
      FONE[idx] = one;
      for (int d = 0; d < dim; d++) { FNd[(f * dim + d) * nNf * nNf + ab] = nd3[d]; FDN[(f * dim + d) * nNf * nNf + ab] = dn3[d]; }
      for (int k1 = 0; k1 < nD; k1++) for (int k2 = 0; k2 < nD; k2++) {   // row dof k1, column dof k2
        const size_t o = f * t * t + (a * nD + k1) + t * (b * nD + k2);
        FT[o] = tt[k2 * nD + k1];                                                   // tau(nd, md) stored col-major nD x nD: index md*nD + nd
        FCN[o] = (k1 == k2 ? (hasConv ? cn : 0.0) + tdn : 0.0) + fsn[k2 * nD + k1];   // convection (diag in dofs) + UNabU face block
      }
    }
    // reference-to-physical mass matrix (Mass.cpp:5-38): read by the time schemes and, for curved elements, inverted for Sqq^-1 = M^-1 (x) I;
    // a straight-sided element without time scheme never reads it (Sqq itself is not stored: the condensation only needs W)
    const bool affW = affE && p.mhinv;
    if (!affW || euler || p.timeScheme == 2 || dump) {
      const int lane = tid & 31, warp = tid >> 5, NWARP = NT / 32, MT = (nN + 7) / 8, NG = (nN + 23) / 24;
      for (int task = warp; task < MT * NG; task += NWARP) {
        mma_task_rt<3>(task % MT, (task / MT) * 3, lane, nN, nN, nIP,
            [&](int i, int ip) { return SHP[ip * nN + i] * DV[ip]; },
            [&](int ip, int j) { return SHP[ip * nN + j]; },
            [&](int i, int j, double v0, double v1) { if (i < nN) { if (j < nN) MM[i * nN + j] = v0; if (j + 1 < nN) MM[i * nN + j + 1] = v1; } });
      }
    }
    __syncthreads();

    HFX_GPROF(2);
    // ---- local matrix, block by block ------------------------------------------------------------------------------------------------
    // uu
    for (int idx = tid; idx < nN * nN; idx += NT) {
      const int i = idx / nN, j = idx - i * nN;
      double sc = 0.0, un_same = 0.0, un[9];
      for (int c = 0; c < sT; c++) un[c] = 0.0;
      if (hasReac || hasConv || hasUN) for (int ip = 0; ip < nIP; ip++) {
        const double pi_ = p.shape[ip * nN + i], pj = p.shape[ip * nN + j], dv = DV[ip];
        const double* gi = GM + (ip * nN + i) * dim; const double* gj = GM + (ip * nN + j) * dim;
        sc = fma(LW[ip] * pi_, pj, sc);                                                      // Reaction.cpp:24-36
        if (hasConv) { double vg = 0.0; for (int d = 0; d < dim; d++) vg = fma(VIP[ip * dim + d], gi[d], vg); sc = fma(-dv * vg, pj, sc); }   // -C^T
        if (hasUN) {   // HDGUNabU.cpp:153-177
          double sg = 0.0;
          for (int d = 0; d < dim; d++) sg = fma(SIP[ip * nD + d], gi[d], sg);
          un_same = fma(-(DIVS[ip] * pi_ + sg) * dv, pj, un_same);
          for (int k1 = 0; k1 < nD; k1++) for (int k2 = 0; k2 < nD; k2++) un[k2 * nD + k1] = fma(-SIP[ip * nD + k1] * dv, fma(gj[k2], pi_, gi[k2] * pj), un[k2 * nD + k1]);
        }
      }
      double ft[9];
      for (int c = 0; c < sT; c++) ft[c] = 0.0;
      for (int f = 0; f < nFc; f++) {
        const int a = NIF[f * nN + i], b = NIF[f * nN + j];
        if (a >= 0 && b >= 0) for (int k1 = 0; k1 < nD; k1++) for (int k2 = 0; k2 < nD; k2++) ft[k2 * nD + k1] += FT[f * t * t + (a * nD + k1) + t * (b * nD + k2)];
      }
      for (int k1 = 0; k1 < nD; k1++) for (int k2 = 0; k2 < nD; k2++) {
        const double bu = hasUN ? un[k2 * nD + k1] + (k1 == k2 ? un_same : 0.0) : 0.0;
        if (hasUN) BUU[(i * nD + k1) + u * (j * nD + k2)] = bu;
        Lm[(i * nD + k1) + n * (j * nD + k2)] = (k1 == k2 ? sc : 0.0) + bu + ft[k2 * nD + k1];
      }
    }
    // uq (HDGDiffusion bulk + faces) and qu (HDGBase bulk): C[(i,d)][j] = sum_ip (g_i)_d dV phi_j on the tensor cores; with a diffusion
    // field a second pass contracts (D g_i)_d for Suq (HDGDiffusion.cpp:130-144), otherwise Suq's bulk part equals Squ's
    {
      const int lane = tid & 31, warp = tid >> 5, NWARP = NT / 32;
      const int Mr = nN * dim, MT = (Mr + 7) / 8, NG = (nN + 23) / 24;
      const int npass = (hasDiff && diffField) ? 2 : 1;
      for (int task = warp; task < npass * MT * NG; task += NWARP) {
        const int pass = task / (MT * NG), r = task - pass * (MT * NG), mt = r % MT, ng = r / MT;
        mma_task_rt<3>(mt, ng * 3, lane, Mr, nN, nIP,
            [&](int m, int ip) {
              const int i = m / dim, d = m - i * dim;
              const double* gi = GM + (ip * nN + i) * dim;
              double g = gi[d];
              if (pass == 1) { g = 0.0; for (int b2 = 0; b2 < dim; b2++) g = fma(DIP[ip * dd + b2 * dim + d], gi[b2], g); }
              return g * DV[ip];
            },
            [&](int ip, int j) { return SHP[ip * nN + j]; },
            [&](int m, int j0, double v0, double v1) {
              if (m >= Mr) return;
              const int i = m / dim, d = m - i * dim;
#pragma unroll
              for (int h = 0; h < 2; h++) {
                const int j = j0 + h;
                if (j >= nN) continue;
                const double v = h ? v1 : v0;
                const bool writeSuq = (pass == 1) || !diffField;
                double suq = hasDiff ? (diffField ? v : p.diffConst * v) : 0.0;
                if (writeSuq && hasDiff) for (int f = 0; f < nFc; f++) {
                  const int a = NIF[f * nN + i], b = NIF[f * nN + j];
                  if (a >= 0 && b >= 0) suq -= FDN[(f * dim + d) * nNf * nNf + a * nNf + b];
                }
                for (int k = 0; k < nD; k++) {
                  if (writeSuq) Lm[(i * nD + k) + n * (sQ + (j * dim + d) * nD + k)] = suq;     // Suq[(i,k),(j,d,k)]
                  if (pass == 0) {
                    Lm[(sQ + (i * dim + d) * nD + k) + n * (j * nD + k)] = v;                   // Squ[(i,d,k),(j,k)]
                  }
                }
              }
            });
      }
    }
    HFX_GPROF(3);
    // ul, lu, ql, lq, ll
    for (int idx = tid; idx < nN * nFc * nNf; idx += NT) {
      const int i = idx / (nFc * nNf), fb = idx - i * nFc * nNf, f = fb / nNf, b = fb - f * nNf;
      const int a = NIF[f * nN + i];
      if (a < 0) continue;
      for (int k1 = 0; k1 < nD; k1++) {
        for (int k2 = 0; k2 < nD; k2++) {
          const size_t o = f * t * t + (a * nD + k1) + t * (b * nD + k2);
          Lm[(i * nD + k1) + n * (sL + (f * nNf + b) * nD + k2)] = -FT[o] + FCN[o];        // Sul[(fn_a,k1),(f,b,k2)]
          const size_t ot = f * t * t + (b * nD + k1) + t * (a * nD + k2);
          Lm[(sL + (f * nNf + b) * nD + k1) + n * (i * nD + k2)] = FT[ot];                // Slu[(f,b,k1),(fn_a,k2)]
        }
        for (int d = 0; d < dim; d++) {
          Lm[(sQ + (i * dim + d) * nD + k1) + n * (sL + (f * nNf + b) * nD + k1)] = -FNd[(f * dim + d) * nNf * nNf + a * nNf + b];   // Sql
          if (hasDiff) Lm[(sL + (f * nNf + b) * nD + k1) + n * (sQ + (i * dim + d) * nD + k1)] = -FDN[(f * dim + d) * nNf * nNf + b * nNf + a];   // Slq
        }
      }
    }
    for (int idx = tid; idx < nFc * t * t; idx += NT) {
      const int f = idx / (t * t), rc = idx - f * t * t, r = rc % t, c = rc / t;
      Lm[(sL + f * t + r) + n * (sL + f * t + c)] = -FT[idx] + FCN[idx];   // Sll
    }
    // right-hand side: Source.cpp:24-48 (per component for the Burgers model)
    if (hasSrc) for (int idx = tid; idx < nN * P.nSrc; idx += NT) {
      const int i = idx / P.nSrc, c = idx - i * P.nSrc;
      double s = 0.0;
      for (int ip = 0; ip < nIP; ip++) s = fma(p.shape[ip * nN + i], p.srcIP[((size_t)e * P.nSrc + c) * nIP + ip] * DV[ip], s);
      Fv[i * nD + c] = s;
    }
    __syncthreads();
    HFX_GPROF(4);
    if (hasUN) {   // rhs = 1/2 op [u0; lambda0] on the u and lambda segments (HDGUNabU.cpp:178-190); op = the UNabU blocks only
      for (int r = tid; r < u + l; r += NT) {
        double s = 0.0;
        if (r < u) {
          const int i = r / nD, k1 = r - i * nD;
          for (int j = 0; j < u; j++) s = fma(BUU[r + u * j], 0.5 * SOL[j], s);
          for (int f = 0; f < nFc; f++) {
            const int a = NIF[f * nN + i];
            if (a >= 0) for (int c = 0; c < t; c++) s = fma(FCN[f * t * t + (a * nD + k1) + t * c], 0.5 * TR[f * t + c], s);   // (no model combines HDGUNabU with HDGConvection: FCN is the UNabU block)
          }
          Fv[r] += s;
        } else {
          const int rl = r - u, f = rl / t, rr = rl - f * t;
          for (int c = 0; c < t; c++) s = fma(FCN[f * t * t + rr + t * c], 0.5 * TR[f * t + c], s);
          Fv[sL + rl] += s;
        }
      }
      __syncthreads();
    }
    // time scheme: Euler.cpp:18-37 on the u rows (hook HDGModel.cpp:38-47)
    if (euler) {
      for (int idx = tid; idx < u * n; idx += NT) { const int c = idx / u, r = idx - c * u; Lm[r + n * c] *= p.dt; }
      __syncthreads();
      for (int idx = tid; idx < nN * nN; idx += NT) { const int i = idx / nN, j = idx - i * nN; for (int k = 0; k < nD; k++) Lm[(i * nD + k) + n * (j * nD + k)] += MM[idx]; }
      for (int r = tid; r < u; r += NT) {
        const int i = r / nD, k = r - i * nD;
        double s = 0.0;
        for (int j = 0; j < nN; j++) s = fma(MM[i * nN + j], SOLD[j * nD + k], s);
        Fv[r] = fma(Fv[r], p.dt, s);
      }
      __syncthreads();
    }

    if (p.timeScheme == 2) {   // RungeKutta::apply on the u rows (stiffness = Su [u x n], columns [Solution | Flux | Trace])
      // UJ = dt sum_{s < stage} a_s [RKStage_s ; RKStage_Flux_s ; RKStage_Trace_s] + UT,  UT = [OldSolution ; OldFlux ; OldTrace]   (element-local order)
      double* UT = Bq; double* UJ = Bq + n;     // Bq [q x (l+1)] >= 2n doubles is free until the condensation
      for (int j = tid; j < n; j += NT) {
        double ut, uj = 0.0;
        if (j < u) { ut = P.oldSol[(size_t)e * u + j]; for (int s2 = 0; s2 < P.rkStage; s2++) uj = fma(P.rkRow[s2], P.rkSol[s2][(size_t)e * u + j], uj); }
        else if (j < sL) { const int jq = j - u; ut = P.oldFlux[(size_t)e * q + jq]; for (int s2 = 0; s2 < P.rkStage; s2++) uj = fma(P.rkRow[s2], P.rkFlux[s2][(size_t)e * q + jq], uj); }
        else {
          const int jl = j - sL, fa = jl / nD, k = jl - fa * nD, f = fa / nNf;
          const size_t g = ((size_t)FACE[f] * nNf + PERM[fa]) * nD + k;
          ut = P.oldTrace[g];
          for (int s2 = 0; s2 < P.rkStage; s2++) uj = fma(P.rkRow[s2], P.rkTrace[s2][g], uj);
        }
        UT[j] = ut; UJ[j] = fma(uj, p.dt, ut);
      }
      __syncthreads();
      const double ass = P.rkRow[P.rkStage];
      for (int r = tid; r < u; r += NT) {   // s <- dt s - (dt Su) uj ; Su <- a_ss dt Su + [M 0 0] ; s += Su ut
        const int i = r / nD, k = r - i * nD;
        double ex = 0.0, im = 0.0;
        for (int j = 0; j < n; j++) {
          double v = Lm[r + n * j] * p.dt;
          ex = fma(v, UJ[j], ex);
          v *= ass;
          if (j < u && (j % nD) == k) v += MM[i * nN + j / nD];
          Lm[r + n * j] = v;
          im = fma(v, UT[j], im);
        }
        Fv[r] = fma(Fv[r], p.dt, -ex) + im;
      }
      __syncthreads();
    }

    if (dump) {   // the local system as Model::compute leaves it (operators + time scheme), S_qq = M (x) I_{dim nDOF} (HDGBase.cpp:152)
      const int sd = dim * nD;
      for (int idx = tid; idx < n * n; idx += NT) {
        const int c = idx / n, r = idx - c * n;
        double v;
        if (r >= sQ && r < sL && c >= sQ && c < sL) { const int rq = r - sQ, cq = c - sQ; v = (rq % sd) == (cq % sd) ? MM[(rq / sd) * nN + cq / sd] : 0.0; }
        else v = Lm[idx];
        P.dumpA[idx] = v;
      }
      for (int i = tid; i < n; i += NT) P.dumpF[i] = Fv[i];
      return;
    }
    HFX_GPROF(5);
    // ---- static condensation (HDGSolver.cpp:331-348) -------------------------------------------------------------------------------
    // W = M^-1: a straight-sided element has M = det J * M_ref (constant det J), so W = M_ref^-1 / det J comes from the reference table;
    // curved elements invert M (definite: unpivoted Gauss-Jordan)
    const double* WI; int ldW; double wscale = 1.0;
    if (affW) { WI = p.mhinv; ldW = (nN + 1) & ~1; wscale = p.w[0] / DV[0]; }
    else {
      for (int idx = tid; idx < nN * nN; idx += NT) AUG[idx] = MM[idx];
      __syncthreads();
      WI = cta_invert_np(AUG, AUG + nN * nN, nN, p.status); ldW = nN;
    }
    HFX_GPROF(6);
    const int lane = tid & 31, warp = tid >> 5, NWARP = NT / 32;
    {   // A = Sqq^-1 Squ, B = Sqq^-1 Sql (last column of B = 0): one nN x (u+l) x nN product per (direction, dof) slot s2 of the q rows
      const int sd = dim * nD, NC1 = u + l + 1, MT = (nN + 7) / 8, NG = (NC1 + 23) / 24;
      for (int task = warp; task < sd * MT * NG; task += NWARP) {
        const int s2 = task / (MT * NG), r = task - s2 * (MT * NG), mt = r % MT, ng = r / MT;
        mma_task_rt<3>(mt, ng * 3, lane, nN, u + l, nN,
            [&](int i, int j) { return WI[i * ldW + j] * wscale; },
            [&](int j, int c) { return Lm[(sQ + j * sd + s2) + n * (c < u ? c : sL + (c - u))]; },
            [&](int i, int c, double v0, double v1) {
              if (i < nN) {
                const size_t rq = i * sd + s2;
                if (c < u) Aq[rq + q * c] = v0; else if (c < NC1) Bq[rq + q * (c - u)] = c < u + l ? v0 : 0.0;
                if (c + 1 < u) Aq[rq + q * (c + 1)] = v1; else if (c + 1 < NC1) Bq[rq + q * (c + 1 - u)] = c + 1 < u + l ? v1 : 0.0;
              }
            });
      }
    }
    __syncthreads();
    HFX_GPROF(7);
    const bool pivK = hasUN || P.forcePivot;   // the Newton-linearised convection block can make K indefinite: partial pivoting there, definite otherwise
    {   // K = Suu - Suq A (into the augmented shared matrix) ; R = [Sul - Suq B | -Fu]
      const int MT = (u + 7) / 8, NC = u + l, NG = (NC + 23) / 24;
      for (int task = warp; task < MT * NG; task += NWARP) {
        const int mt = task % MT, ng = task / MT;
        mma_task_rt<3>(mt, ng * 3, lane, u, NC, q,
            [&](int r, int rq) { return Lm[r + n * (sQ + rq)]; },
            [&](int rq, int c) { const double* pc = c < u ? Aq + q * c : Bq + q * (c - u); return pc[rq]; },
            [&](int r, int c, double v0, double v1) {
              if (r < u) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                  const int cc = c + h; const double v = h ? v1 : v0;
                  if (cc < u) {
                    const double kv = Lm[r + n * cc] - v;
                    BUU[r * u + cc] = kv;   // copy of K for the refinement of U (BUU is dead once the UNabU right-hand side is formed)
                    if (pivK) { AUG[r * 2 * u + cc] = kv; AUG[r * 2 * u + u + cc] = (r == cc) ? 1.0 : 0.0; }
                    else AUG[r * u + cc] = kv;
                  }
                  else if (cc < NC) Rm[r + u * (cc - u)] = Lm[r + n * (sL + cc - u)] - v;
                }
              }
            });
      }
      for (int r = tid; r < u; r += NT) Rm[r + u * l] = -Fv[r];
    }
    __syncthreads();
    HFX_GPROF(8);
    const double* KI; int ldK;
    if (pivK) { cta_invert(AUG, u, SCR, IPIV, p.status); KI = AUG + u; ldK = 2 * u; }
    else { KI = cta_invert_np(AUG, AUG + u * u, u, p.status); ldK = u; }
    HFX_GPROF(9);
    {   // U = -K^-1 R (column l: U0 = K^-1 Fu)
      const int MT = (u + 7) / 8, NG = (l + 1 + 15) / 16;
      for (int task = warp; task < MT * NG; task += NWARP) {
        const int mt = task % MT, ng = task / MT;
        mma_task_rt<2>(mt, ng * 2, lane, u, l + 1, u,
            [&](int r, int j) { return KI[r * ldK + j]; },
            [&](int j, int c) { return Rm[j + u * c]; },
            [&](int r, int c, double v0, double v1) {
              if (r < u) { if (c <= l) Um[r + u * c] = -v0; if (c + 1 <= l) Um[r + u * (c + 1)] = -v1; }
            });
      }
    }
    __syncthreads();
    {   // one step of iterative refinement, U <- U - K^-1 (K U + R): the explicit inverse alone loses a factor ~ 1 / h on fine meshes (see hfx_assemble.cuh P7r)
      const int MT = (u + 7) / 8, NG = (l + 1 + 15) / 16;
      for (int task = warp; task < MT * NG; task += NWARP) {
        const int mt = task % MT, ng = task / MT;
        mma_task_rt<2>(mt, ng * 2, lane, u, l + 1, u,
            [&](int r, int j) { return BUU[r * u + j]; },
            [&](int j, int c) { return Um[j + u * c]; },
            [&](int r, int c, double v0, double v1) {
              if (r < u) { if (c <= l) Rm[r + u * c] += v0; if (c + 1 <= l) Rm[r + u * (c + 1)] += v1; }
            });
      }
      __syncthreads();
      for (int task = warp; task < MT * NG; task += NWARP) {
        const int mt = task % MT, ng = task / MT;
        mma_task_rt<2>(mt, ng * 2, lane, u, l + 1, u,
            [&](int r, int j) { return KI[r * ldK + j]; },
            [&](int j, int c) { return Rm[j + u * c]; },
            [&](int r, int c, double v0, double v1) {
              if (r < u) { if (c <= l) Um[r + u * c] -= v0; if (c + 1 <= l) Um[r + u * (c + 1)] -= v1; }
            });
      }
      __syncthreads();
    }
    HFX_GPROF(10);
    {   // Q = -A U - B (column l: Q0 = -A U0)
      const int MT = (q + 7) / 8, NG = (l + 1 + 23) / 24;
      for (int task = warp; task < MT * NG; task += NWARP) {
        const int mt = task % MT, ng = task / MT;
        mma_task_rt<3>(mt, ng * 3, lane, q, l + 1, u,
            [&](int rq, int j) { return Aq[rq + q * j]; },
            [&](int j, int c) { return Um[j + u * c]; },
            [&](int rq, int c, double v0, double v1) {
              if (rq < q) {
                if (c <= l) Qm[rq + q * c] = -v0 - Bq[rq + q * c];
                if (c + 1 <= l) Qm[rq + q * (c + 1)] = -v1 - Bq[rq + q * (c + 1)];
              }
            });
      }
    }
    __syncthreads();
    HFX_GPROF(11);
    // write U, Q, U0, Q0 (HDGSolver.cpp:336-341; kept row-major per element on the device, see recover_kernel)
    {   // one warp per row: consecutive lanes -> consecutive addresses of the row-major global block
      const int lane_ = tid & 31, warp_ = tid >> 5, nw_ = NT / 32;
      double* gU = p.U + (size_t)e * u * l; double* gQ = p.Q + (size_t)e * q * l;
      for (int r = warp_; r < u + q; r += nw_) {
        if (r < u) for (int cc = lane_; cc < l; cc += 32) gU[r * l + cc] = Um[r + u * cc];
        else { const int rq = r - u; for (int cc = lane_; cc < l; cc += 32) gQ[rq * l + cc] = Qm[rq + q * cc]; }
      }
    }
    for (int i = tid; i < u; i += NT) p.U0[(size_t)e * u + i] = Um[u * l + i];
    for (int i = tid; i < q; i += NT) p.Q0[(size_t)e * q + i] = Qm[q * l + i];
    // S = Slu U + Slq Q + Sll ; S0 = Fl - Slu U0 - Slq Q0 ; boundary rows (:489-501) ; scatter (:596-618)
    double* gS = p.S ? p.S + (size_t)e * l * l : nullptr;
    double* gS0 = p.S0 ? p.S0 + (size_t)e * l : nullptr;
    HFX_GPROF(12);
    auto emitS = [&](int r, int c, double s) {
      const int f = r / t, bc = BCF[f];
      if (c < l) {
        s += Lm[(sL + r) + n * (sL + c)];
        const int f2 = c / t;
        if (bc) {
          const int rr = r - f * t, a = rr / nD, k1 = rr - a * nD, cc = c - f2 * t, b = cc / nD, k2 = cc - b * nD;
          if (bc == 1) s = (r == c) ? 1.0 : 0.0;                                                       // DirichletModel: identity row
          else s = (f2 == f && k1 == k2) ? FONE[f * nNf * nNf + a * nNf + b] : 0.0;                    // IntegratedDirichletModel: face mass (x) I
        }
        if (gS) gS[r + l * c] = s;
        double* dst = p.vals + ROWOFF[r] + COLOFF[f * l + c];
        if (f2 == f && INTF[f]) atomicAdd(dst, s); else *dst = s;
      } else {
        const int rr = r - f * t, a = rr / nD, k1 = rr - a * nD, F = FACE[f];
        double s0 = Fv[sL + r] - s;
        if (bc == 1) s0 = p.dirichlet[((size_t)F * nNf + a) * nD + k1];
        else if (bc == 2) { s0 = 0.0; for (int b = 0; b < nNf; b++) s0 = fma(FONE[f * nNf * nNf + a * nNf + b], p.dirichlet[((size_t)F * nNf + b) * nD + k1], s0); }
        if (gS0) gS0[r] = s0;
        double* dst = p.rhs + ((size_t)F * nNf + PERM[f * nNf + a]) * nD + k1;
        if (INTF[f]) atomicAdd(dst, s0); else *dst = s0;
      }
    };
    {
      const int MT = (l + 7) / 8, NG = (l + 1 + 23) / 24;
      for (int task = warp; task < MT * NG; task += NWARP) {
        const int mt = task % MT, ng = task / MT;
        mma_task_rt<3>(mt, ng * 3, lane, l, l + 1, u + q,
            [&](int r, int k) { return Lm[(sL + r) + n * k]; },
            [&](int k, int c) {
              if (P.explicitS) return c >= l ? (k < u ? P.solCur[(size_t)e * u + k] : P.fluxCur[(size_t)e * q + (k - u)]) : 0.0;   // explicit trace problem (:349-353)
              const double* pc = k < u ? Um + u * c + k : Qm + q * c + (k - u); return *pc; },
            [&](int r, int c, double v0, double v1) {
              if (r < l) { if (c <= l) emitS(r, c, v0); if (c + 1 <= l) emitS(r, c + 1, v1); }
            });
      }
    }
    __syncthreads();
    HFX_GPROF(13);
  }
}

inline size_t gen_smem_bytes(int nN, int nNf, int nFc, int nD, int dim, int nIPf) {
  const int u = nN * nD, nmax = u > nN ? u : nN, nFf = nFc * nIPf;
  size_t doubles = (size_t)nmax * 2 * nmax + 3 * nmax + 2;
  doubles += (size_t)((nIPf * nNf + 1) & ~1) + 3 * (size_t)nFf + 2 * (size_t)nFf * dim + (size_t)nFf * nD * nD + (size_t)((nFf * nD + 1) & ~1) + 2;
  size_t ints = (size_t)nFc * nNf * 2 + (size_t)nFc * nN + (size_t)nFc * 5 + (size_t)nFc * nFc + 2;
  const int l = nFc * nNf * nD;
  ints += 2 + (size_t)((nFc * l + 1) & ~1);   // IPIV pad, COLOFF
  ints = (ints + 3) & ~(size_t)3;
  return doubles * 8 + ints * 4 + 8 * (size_t)nFc + 8 * (size_t)l + 64;
}
// Optional shared-memory residents (Um, Rm, Aq, Bq, shape, GM) within `budget` bytes, in that order of priority; returns the extra bytes.
inline size_t gen_smem_optional(int dim, int nN, int nNf, int nFc, int nIP, int nD, size_t budget, int (&off)[6]) {
  const size_t u = (size_t)nN * nD, q = u * dim, l = (size_t)nFc * nNf * nD;
  const size_t need[6] = {u * (l + 1), u * (l + 1), q * u, q * (l + 1), (size_t)nIP * nN, (size_t)nIP * nN * dim};
  size_t used = 0;
  for (int k = 0; k < 6; k++) {
    const size_t nd = (need[k] + 1) & ~(size_t)1;
    if ((used + nd) * 8 <= budget) { off[k] = (int)used; used += nd; } else off[k] = -1;
  }
  return used * 8;
}

}  // namespace hfx
