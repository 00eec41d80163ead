// General per-element HDG kernel (runtime sizes): any simplex dimension / order the reference element supports, any number of
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
  int nSrc;                  // source components: 1, or dim for the Burgers model (HDGBurgersModel.cpp:112-122)
  double* ws;                // per-CTA scratch
  long long wsStride;        // doubles per CTA
  // RungeKutta::apply (src/operator/RungeKutta.cpp:90-143), auxiliary fields {Flux, Trace}: time scheme code 2
  int rkStage, rkNumStages;
  double rkRow[8];           // Butcher row of the current stage (a_s0 .. a_s,nStages-1)
  const double* oldSol; const double* oldFlux; const double* oldTrace;          // OldSolution / OldFlux (cell), OldTrace (face)
  const double* rkSol[8]; const double* rkFlux[8]; const double* rkTrace[8];   // RKStage_k, RKStage_Flux_k (cell), RKStage_Trace_k (face)
};

// scratch layout (offsets in doubles), identical on host and device
struct GenWs {
  int u, q, l, n, t, nJ, nFf, dd, sQ, sL;
  long long oLm, oF, oGM, oDV, oIJ, oNRM, oTAUS, oDIP, oVIP, oVDN, oFS, oTDN, oSIP, oDIVS, oX, oTAUn, oDN, oVN, oSOL, oTR, oSOLD, oMM, oW, oFT, oFCN,
      oFNd, oFDN, oFONE, oBUU, oAq, oBq, oRm, oUm, oQm, oLW, total;
  __host__ __device__ GenWs(int dim, int nN, int nNf, int nFc, int nIP, int nIPf, int nD) {
    u = nN * nD; q = u * dim; t = nNf * nD; l = nFc * t; n = u + q + l; nJ = nIP + nFc * nIPf; nFf = nFc * nIPf; dd = dim * dim; sQ = u; sL = u + q;
    long long o = 0;
    auto take = [&](long long k) { long long r = o; o += (k + 1) & ~1LL; return r; };
    oLm = take((long long)n * n); oF = take(n); oGM = take((long long)nIP * nN * dim); oDV = take(nJ); oIJ = take((long long)nIP * dd);
    oNRM = take((long long)nFf * dim); oTAUS = take((long long)nFf * nD * nD); oDIP = take((long long)nJ * dd); oVIP = take((long long)nIP * dim);
    oVDN = take(nFf); oFS = take((long long)nFf * nD); oTDN = take(nFf); oSIP = take((long long)nIP * nD); oDIVS = take(nIP);
    oX = take((long long)nN * dim); oTAUn = take((long long)nFc * nNf * nD * nD); oDN = take((long long)nN * dd); oVN = take((long long)nN * dim);
    oSOL = take(u); oTR = take(l); oSOLD = take(u); oMM = take((long long)nN * nN); oW = take((long long)nN * nN);
    oFT = take((long long)nFc * t * t); oFCN = take((long long)nFc * t * t); oFNd = take((long long)nFc * dim * nNf * nNf);
    oFDN = take((long long)nFc * dim * nNf * nNf); oFONE = take((long long)nFc * nNf * nNf); oBUU = take((long long)u * u);
    oAq = take((long long)q * u); oBq = take((long long)q * (l + 1)); oRm = take((long long)u * (l + 1)); oUm = take((long long)u * (l + 1));
    oQm = take((long long)q * (l + 1)); oLW = take(2LL * nIP + (long long)nIP * nD);
    total = o;
  }
};

constexpr int kGenThreads = 256;

// Gauss-Jordan inverse with partial pivoting of the left half of the row-major n x 2n matrix aug (right half = identity on entry,
// inverse on exit).  Whole CTA.  scr: n + 2n doubles; ipiv: 1 int in shared memory.
__device__ inline void cta_invert(double* aug, int n, double* scr, int* ipiv, int* status) {
  const int tid = threadIdx.x, NT = blockDim.x, n2 = 2 * n;
  double* fac = scr; double* rk = scr + n;
  for (int k = 0; k < n; k++) {
    if (tid < 32) {   // pivot search by warp 0
      double best = -1.0; int bi = k;
      for (int i = k + tid; i < n; i += 32) { const double v = fabs(aug[(size_t)i * n2 + k]); if (v > best) { best = v; bi = i; } }
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (tid == 0) { *ipiv = bi; if (!(best > 1e-300)) atomicOr(status, 1); }
    }
    __syncthreads();
    const int pv = *ipiv;
    if (pv != k) for (int j = tid; j < n2; j += NT) { const double a = aug[(size_t)k * n2 + j]; aug[(size_t)k * n2 + j] = aug[(size_t)pv * n2 + j]; aug[(size_t)pv * n2 + j] = a; }
    __syncthreads();
    const double ip = 1.0 / aug[(size_t)k * n2 + k];
    for (int j = tid; j < n2; j += NT) rk[j] = aug[(size_t)k * n2 + j] * ip;
    for (int i = tid; i < n; i += NT) fac[i] = aug[(size_t)i * n2 + k];
    __syncthreads();
    for (int idx = tid; idx < n * n2; idx += NT) {
      const int i = idx / n2, j = idx - i * n2;
      aug[idx] = (i == k) ? rk[j] : fma(-fac[i], rk[j], aug[idx]);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kGenThreads, 2) hdg_generic_kernel(const GenParams P) {
  const AsmParams& p = P.a;
  const int dim = P.dim, nN = P.nN, nNf = P.nNf, nFc = P.nFc, nIP = P.nIP, nIPf = P.nIPf, nD = P.nD;
  const GenWs z(dim, nN, nNf, nFc, nIP, nIPf, nD);
  const int u = z.u, q = z.q, l = z.l, n = z.n, t = z.t, nJ = z.nJ, nFf = z.nFf, dd = z.dd, sQ = z.sQ, sL = z.sL, sT = nD * nD;
  const int tid = threadIdx.x, NT = kGenThreads;
  const bool hasDiff = p.opmask & 1, hasConv = p.opmask & 2, hasReac = (p.opmask & 4) && p.reacIP, hasSrc = (p.opmask & 8) && P.a.srcIP, hasUN = p.opmask & 16;
  const bool euler = p.timeScheme == 1;
  const bool diffField = hasDiff && p.diffComps > 0 && p.diff;
  double* ws = P.ws + (size_t)blockIdx.x * P.wsStride;
  double* Lm = ws + z.oLm; double* Fv = ws + z.oF; double* GM = ws + z.oGM; double* DV = ws + z.oDV; double* IJ = ws + z.oIJ; double* NRM = ws + z.oNRM;
  double* TAUS = ws + z.oTAUS; double* DIP = ws + z.oDIP; double* VIP = ws + z.oVIP; double* VDN = ws + z.oVDN; double* FS = ws + z.oFS; double* TDN = ws + z.oTDN;
  double* SIP = ws + z.oSIP; double* DIVS = ws + z.oDIVS; double* X = ws + z.oX; double* TAUn = ws + z.oTAUn; double* DN = ws + z.oDN; double* VN = ws + z.oVN;
  double* SOL = ws + z.oSOL; double* TR = ws + z.oTR; double* SOLD = ws + z.oSOLD; double* MM = ws + z.oMM; double* W = ws + z.oW; double* FT = ws + z.oFT;
  double* FCN = ws + z.oFCN; double* FNd = ws + z.oFNd; double* FDN = ws + z.oFDN; double* FONE = ws + z.oFONE; double* BUU = ws + z.oBUU; double* Aq = ws + z.oAq;
  double* Bq = ws + z.oBq; double* Rm = ws + z.oRm; double* Um = ws + z.oUm; double* Qm = ws + z.oQm; double* LW = ws + z.oLW;
  // shared: augmented matrix for the two inverses, scratch rows, integer maps
  extern __shared__ __align__(16) double gsm[];
  const int nmax = u > nN ? u : nN;
  double* AUG = gsm;                                 // [nmax][2 nmax]
  double* SCR = AUG + (size_t)nmax * 2 * nmax;       // [3 nmax]
  long long* ROWS = reinterpret_cast<long long*>(SCR + 3 * nmax + 2);   // [nFc] first entry of row (F,0) in vals
  int* PERM = reinterpret_cast<int*>(ROWS + nFc);    // [nFc*nNf]
  int* NIF = PERM + nFc * nNf;                       // [nFc*nN]
  int* FNo = NIF + nFc * nN;                         // [nFc*nNf]
  int* FACE = FNo + nFc * nNf;                       // [nFc] global face ids
  int* BCF = FACE + nFc; int* INTF = BCF + nFc; int* POS = INTF + nFc;   // [nFc], [nFc], [nFc*nFc]
  int* RLEN = POS + nFc * nFc; int* OPP = RLEN + nFc; int* IPIV = OPP + nFc;

  for (int i = tid; i < nFc * nNf; i += NT) FNo[i] = p.faceNodes[i];
  for (int i = tid; i < nFc * nN; i += NT) NIF[i] = p.nodeInFace[i];
  if (tid < nFc) { int vn = 0; for (int kk = 0; kk < nN; kk++) if (p.nodeInFace[tid * nN + kk] < 0) { vn = kk; break; } OPP[tid] = vn; }
  __syncthreads();

  for (int e = blockIdx.x; e < p.nCells; e += gridDim.x) {
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
    for (long long i = tid; i < (long long)n * n; i += NT) Lm[i] = 0.0;
    for (int i = tid; i < n; i += NT) Fv[i] = 0.0;
    __syncthreads();

    // ---- geometry and coefficients at the cubature points -------------------------------------------------------------------------
    for (int k = tid; k < nJ; k += NT) {
      if (k < nIP) {
        const int ip = k;
        double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        for (int i = 0; i < nN; i++) {
          const double* d = p.dshape + ((size_t)ip * nN + i) * dim;
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
          double s = ((c / dim) == (c % dim)) ? 1.0 : 0.0;
          if (diffField) { s = 0.0; for (int i = 0; i < nN; i++) s = fma(p.shape[(size_t)ip * nN + i], DN[i * dd + c], s); }
          DIP[ip * dd + c] = s;
        }
        if (hasConv) for (int d = 0; d < dim; d++) { double s = 0.0; for (int i = 0; i < nN; i++) s = fma(p.shape[(size_t)ip * nN + i], VN[i * dim + d], s); VIP[ip * dim + d] = s; }
        if (hasUN) for (int k2 = 0; k2 < nD; k2++) { double s = 0.0; for (int i = 0; i < nN; i++) s = fma(SOL[i * nD + k2], p.shape[(size_t)ip * nN + i], s); SIP[ip * nD + k2] = s; }
        LW[ip] = hasReac ? p.reacIP[(size_t)e * nIP + ip] * dv : 0.0;
      } else {
        const int fi = k - nIP, f = fi / nIPf, ip = fi - f * nIPf;
        const int* fn = FNo + f * nNf;
        double J[2][3] = {{0, 0, 0}, {0, 0, 0}};
        for (int a = 0; a < nNf; a++) {
          const double* d = p.fdshape + ((size_t)ip * nNf + a) * (dim - 1);
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
        DV[nIP + fi] = dvf;
        for (int m = 0; m < dim; m++) NRM[fi * dim + m] = nv[m];
        for (int c = 0; c < sT; c++) { double s = 0.0; for (int a = 0; a < nNf; a++) s = fma(TAUn[(f * nNf + a) * sT + c], p.fshape[(size_t)ip * nNf + a], s); TAUS[fi * sT + c] = s; }
        for (int c = 0; c < dd; c++) {
          double s = ((c / dim) == (c % dim)) ? 1.0 : 0.0;
          if (diffField) { s = 0.0; for (int a = 0; a < nNf; a++) s = fma(p.fshape[(size_t)ip * nNf + a], DN[fn[a] * dd + c], s); }
          DIP[(nIP + fi) * dd + c] = s;
        }
        double vdn = 0.0;
        if (hasConv) for (int d = 0; d < dim; d++) { double s = 0.0; for (int a = 0; a < nNf; a++) s = fma(p.fshape[(size_t)ip * nNf + a], VN[fn[a] * dim + d], s); vdn = fma(s, nv[d], vdn); }
        VDN[fi] = dvf * vdn;
        if (hasUN) {   // HDGUNabU.cpp:107-124
          double tdn = 0.0;
          for (int d = 0; d < dim; d++) {
            double tr = 0.0, fs = 0.0;
            for (int a = 0; a < nNf; a++) { tr = fma(TR[(f * nNf + a) * nD + d], p.fshape[(size_t)ip * nNf + a], tr); fs = fma(SOL[fn[a] * nD + d], p.fshape[(size_t)ip * nNf + a], fs); }
            tdn = fma(tr, nv[d], tdn);
            FS[fi * nD + d] = fs;
          }
          TDN[fi] = tdn * dvf;
        }
      }
    }
    __syncthreads();
    // physical gradients gm(d,i) = (J^-1 grad_ref phi_i)_d at the bulk points
    for (int idx = tid; idx < nIP * nN; idx += NT) {
      const int ip = idx / nN, i = idx - ip * nN;
      const double* dp = p.dshape + ((size_t)ip * nN + i) * dim;
      for (int d = 0; d < dim; d++) { double s = 0.0; for (int r = 0; r < dim; r++) s = fma(IJ[ip * dd + d * dim + r], dp[r], s); GM[((size_t)ip * nN + i) * dim + d] = s; }
    }
    __syncthreads();
    if (hasUN) for (int ip = tid; ip < nIP; ip += NT) {   // div of the previous iterate (HDGUNabU.cpp:153-165)
      double dv = 0.0;
      for (int i = 0; i < nN; i++) for (int d = 0; d < dim; d++) dv = fma(GM[((size_t)ip * nN + i) * dim + d], SOL[i * nD + d], dv);
      DIVS[ip] = dv;
    }
    // ---- weighted face matrices (gather form of the face loops of HDGBase / HDGDiffusion / HDGConvection / HDGUNabU) -----------
    for (int idx = tid; idx < nFc * nNf * nNf; idx += NT) {
      const int f = idx / (nNf * nNf), ab = idx - f * nNf * nNf, a = ab / nNf, b = ab - a * nNf;
      double one = 0.0, cn = 0.0, tdn = 0.0, nd3[3] = {0, 0, 0}, dn3[3] = {0, 0, 0}, tt[9], fsn[9];
      for (int c = 0; c < sT; c++) { tt[c] = 0.0; fsn[c] = 0.0; }
      for (int ip = 0; ip < nIPf; ip++) {
        const int fi = f * nIPf + ip;
        const double ss = p.fshape[(size_t)ip * nNf + a] * p.fshape[(size_t)ip * nNf + b], dv = DV[nIP + fi], w = dv * ss;
        one += w;
        cn = fma(VDN[fi], ss, cn);
        for (int d = 0; d < dim; d++) {
          nd3[d] = fma(w, NRM[fi * dim + d], nd3[d]);
          double dn = 0.0;
          for (int b2 = 0; b2 < dim; b2++) dn = fma(DIP[(nIP + fi) * dd + b2 * dim + d], NRM[fi * dim + b2], dn);   // (D n)_d, D col-major
          dn3[d] = fma(w, dn, dn3[d]);
        }
        for (int c = 0; c < sT; c++) tt[c] = fma(w, TAUS[fi * sT + c], tt[c]);
        if (hasUN) {
          tdn = fma(TDN[fi], ss, tdn);
          for (int k1 = 0; k1 < nD; k1++) for (int k2 = 0; k2 < nD; k2++) fsn[k2 * nD + k1] = fma(FS[fi * nD + k1] * NRM[fi * dim + k2], w, fsn[k2 * nD + k1]);
        }
      }
      FONE[idx] = one;
      for (int d = 0; d < dim; d++) { FNd[((size_t)f * dim + d) * nNf * nNf + ab] = nd3[d]; FDN[((size_t)f * dim + d) * nNf * nNf + ab] = dn3[d]; }
      for (int k1 = 0; k1 < nD; k1++) for (int k2 = 0; k2 < nD; k2++) {   // row dof k1, column dof k2
        const size_t o = (size_t)f * t * t + (size_t)(a * nD + k1) + (size_t)t * (b * nD + k2);
        FT[o] = tt[k2 * nD + k1];                                                   // tau(nd, md) stored col-major nD x nD: index md*nD + nd
        FCN[o] = (k1 == k2 ? (hasConv ? cn : 0.0) + tdn : 0.0) + fsn[k2 * nD + k1];   // convection (diag in dofs) + UNabU face block
      }
    }
    // reference-to-physical mass matrix
    for (int idx = tid; idx < nN * nN; idx += NT) {
      const int i = idx / nN, j = idx - i * nN;
      double s = 0.0;
      for (int ip = 0; ip < nIP; ip++) s = fma(p.shape[(size_t)ip * nN + i] * p.shape[(size_t)ip * nN + j], DV[ip], s);
      MM[idx] = s;
    }
    __syncthreads();

    // ---- local matrix, block by block ------------------------------------------------------------------------------------------------
    // uu
    for (int idx = tid; idx < nN * nN; idx += NT) {
      const int i = idx / nN, j = idx - i * nN;
      double sc = 0.0, un_same = 0.0, un[9];
      for (int c = 0; c < sT; c++) un[c] = 0.0;
      for (int ip = 0; ip < nIP; ip++) {
        const double pi_ = p.shape[(size_t)ip * nN + i], pj = p.shape[(size_t)ip * nN + j], dv = DV[ip];
        const double* gi = GM + ((size_t)ip * nN + i) * dim; const double* gj = GM + ((size_t)ip * nN + j) * dim;
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
        if (a >= 0 && b >= 0) for (int k1 = 0; k1 < nD; k1++) for (int k2 = 0; k2 < nD; k2++) ft[k2 * nD + k1] += FT[(size_t)f * t * t + (a * nD + k1) + (size_t)t * (b * nD + k2)];
      }
      for (int k1 = 0; k1 < nD; k1++) for (int k2 = 0; k2 < nD; k2++) {
        const double bu = hasUN ? un[k2 * nD + k1] + (k1 == k2 ? un_same : 0.0) : 0.0;
        if (hasUN) BUU[(size_t)(i * nD + k1) + (size_t)u * (j * nD + k2)] = bu;
        Lm[(size_t)(i * nD + k1) + (size_t)n * (j * nD + k2)] = (k1 == k2 ? sc : 0.0) + bu + ft[k2 * nD + k1];
      }
    }
    // uq (HDGDiffusion bulk + faces) and qu (HDGBase bulk)
    for (int idx = tid; idx < nN * nN * dim; idx += NT) {
      const int i = idx / (nN * dim), jd = idx - i * nN * dim, j = jd / dim, d = jd - j * dim;
      double suq = 0.0, squ = 0.0;
      for (int ip = 0; ip < nIP; ip++) {
        const double dv = DV[ip];
        const double* gi = GM + ((size_t)ip * nN + i) * dim;
        if (hasDiff) { double dg = 0.0; for (int b2 = 0; b2 < dim; b2++) dg = fma(DIP[ip * dd + b2 * dim + d], gi[b2], dg); suq = fma(dg * dv, p.shape[(size_t)ip * nN + j], suq); }
        squ = fma(gi[d] * dv, p.shape[(size_t)ip * nN + j], squ);
      }
      if (hasDiff) for (int f = 0; f < nFc; f++) {
        const int a = NIF[f * nN + i], b = NIF[f * nN + j];
        if (a >= 0 && b >= 0) suq -= FDN[((size_t)f * dim + d) * nNf * nNf + a * nNf + b];
      }
      for (int k = 0; k < nD; k++) {
        Lm[(size_t)(i * nD + k) + (size_t)n * (sQ + (j * dim + d) * nD + k)] = suq;     // Suq[(i,k),(j,d,k)]
        Lm[(size_t)(sQ + (i * dim + d) * nD + k) + (size_t)n * (j * nD + k)] = squ;     // Squ[(i,d,k),(j,k)]
        Lm[(size_t)(sQ + (i * dim + d) * nD + k) + (size_t)n * (sQ + (j * dim + d) * nD + k)] = MM[i * nN + j];   // Sqq = M (x) I
      }
    }
    // ul, lu, ql, lq, ll
    for (int idx = tid; idx < nN * nFc * nNf; idx += NT) {
      const int i = idx / (nFc * nNf), fb = idx - i * nFc * nNf, f = fb / nNf, b = fb - f * nNf;
      const int a = NIF[f * nN + i];
      if (a < 0) continue;
      for (int k1 = 0; k1 < nD; k1++) {
        for (int k2 = 0; k2 < nD; k2++) {
          const size_t o = (size_t)f * t * t + (a * nD + k1) + (size_t)t * (b * nD + k2);
          Lm[(size_t)(i * nD + k1) + (size_t)n * (sL + (f * nNf + b) * nD + k2)] = -FT[o] + FCN[o];        // Sul[(fn_a,k1),(f,b,k2)]
          const size_t ot = (size_t)f * t * t + (b * nD + k1) + (size_t)t * (a * nD + k2);
          Lm[(size_t)(sL + (f * nNf + b) * nD + k1) + (size_t)n * (i * nD + k2)] = FT[ot];                // Slu[(f,b,k1),(fn_a,k2)]
        }
        for (int d = 0; d < dim; d++) {
          Lm[(size_t)(sQ + (i * dim + d) * nD + k1) + (size_t)n * (sL + (f * nNf + b) * nD + k1)] = -FNd[((size_t)f * dim + d) * nNf * nNf + a * nNf + b];   // Sql
          if (hasDiff) Lm[(size_t)(sL + (f * nNf + b) * nD + k1) + (size_t)n * (sQ + (i * dim + d) * nD + k1)] = -FDN[((size_t)f * dim + d) * nNf * nNf + b * nNf + a];   // Slq
        }
      }
    }
    for (int idx = tid; idx < nFc * t * t; idx += NT) {
      const int f = idx / (t * t), rc = idx - f * t * t, r = rc % t, c = rc / t;
      Lm[(size_t)(sL + f * t + r) + (size_t)n * (sL + f * t + c)] = -FT[idx] + FCN[idx];   // Sll
    }
    // right-hand side: Source.cpp:24-48 (per component for the Burgers model)
    if (hasSrc) for (int idx = tid; idx < nN * P.nSrc; idx += NT) {
      const int i = idx / P.nSrc, c = idx - i * P.nSrc;
      double s = 0.0;
      for (int ip = 0; ip < nIP; ip++) s = fma(p.shape[(size_t)ip * nN + i], p.srcIP[((size_t)e * P.nSrc + c) * nIP + ip] * DV[ip], s);
      Fv[i * nD + c] = s;
    }
    __syncthreads();
    if (hasUN) {   // rhs = 1/2 op [u0; lambda0] on the u and lambda segments (HDGUNabU.cpp:178-190); op = the UNabU blocks only
      for (int r = tid; r < u + l; r += NT) {
        double s = 0.0;
        if (r < u) {
          const int i = r / nD, k1 = r - i * nD;
          for (int j = 0; j < u; j++) s = fma(BUU[(size_t)r + (size_t)u * j], 0.5 * SOL[j], s);
          for (int f = 0; f < nFc; f++) {
            const int a = NIF[f * nN + i];
            if (a >= 0) for (int c = 0; c < t; c++) s = fma(FCN[(size_t)f * t * t + (a * nD + k1) + (size_t)t * c], 0.5 * TR[f * t + c], s);   // (no model combines HDGUNabU with HDGConvection: FCN is the UNabU block)
          }
          Fv[r] += s;
        } else {
          const int rl = r - u, f = rl / t, rr = rl - f * t;
          for (int c = 0; c < t; c++) s = fma(FCN[(size_t)f * t * t + rr + (size_t)t * c], 0.5 * TR[f * t + c], s);
          Fv[sL + rl] += s;
        }
      }
      __syncthreads();
    }
    // time scheme: Euler.cpp:18-37 on the u rows (hook HDGModel.cpp:38-47)
    if (euler) {
      for (long long idx = tid; idx < (long long)u * n; idx += NT) { const int r = (int)(idx % u); const long long c = idx / u; Lm[(size_t)r + (size_t)n * c] *= p.dt; }
      __syncthreads();
      for (int idx = tid; idx < nN * nN; idx += NT) { const int i = idx / nN, j = idx - i * nN; for (int k = 0; k < nD; k++) Lm[(size_t)(i * nD + k) + (size_t)n * (j * nD + k)] += MM[idx]; }
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
          double v = Lm[(size_t)r + (size_t)n * j] * p.dt;
          ex = fma(v, UJ[j], ex);
          v *= ass;
          if (j < u && (j % nD) == k) v += MM[i * nN + j / nD];
          Lm[(size_t)r + (size_t)n * j] = v;
          im = fma(v, UT[j], im);
        }
        Fv[r] = fma(Fv[r], p.dt, -ex) + im;
      }
      __syncthreads();
    }

    // ---- static condensation (HDGSolver.cpp:331-348) -------------------------------------------------------------------------------
    for (int idx = tid; idx < nN * 2 * nN; idx += NT) { const int i = idx / (2 * nN), j = idx - i * 2 * nN; AUG[idx] = j < nN ? MM[i * nN + j] : (j - nN == i ? 1.0 : 0.0); }
    __syncthreads();
    cta_invert(AUG, nN, SCR, IPIV, p.status);
    for (int idx = tid; idx < nN * nN; idx += NT) { const int i = idx / nN, j = idx - i * nN; W[idx] = AUG[(size_t)i * 2 * nN + nN + j]; }
    __syncthreads();
    {   // A = Sqq^-1 Squ, B = Sqq^-1 Sql (last column of B = 0)
      const int sd = dim * nD;
      for (long long idx = tid; idx < (long long)q * (u + l + 1); idx += NT) {
        const int rq = (int)(idx % q); const int c = (int)(idx / q);
        const int i = rq / sd, s2 = rq - i * sd;
        double s = 0.0;
        if (c < u + l) {
          const double* col = Lm + (size_t)n * (c < u ? c : sL + (c - u)) + sQ + s2;
          for (int j = 0; j < nN; j++) s = fma(W[i * nN + j], col[(size_t)j * sd], s);
        }
        if (c < u) Aq[(size_t)rq + (size_t)q * c] = s; else Bq[(size_t)rq + (size_t)q * (c - u)] = s;
      }
    }
    __syncthreads();
    // K = Suu - Suq A (into the augmented shared matrix) ; R = [Sul - Suq B | -Fu]
    for (long long idx = tid; idx < (long long)u * (u + l + 1); idx += NT) {
      const int r = (int)(idx % u); const int c = (int)(idx / u);
      double s;
      if (c < u) {
        s = Lm[(size_t)r + (size_t)n * c];
        for (int rq = 0; rq < q; rq++) s = fma(-Lm[(size_t)r + (size_t)n * (sQ + rq)], Aq[(size_t)rq + (size_t)q * c], s);
        AUG[(size_t)r * 2 * u + c] = s;
        AUG[(size_t)r * 2 * u + u + c] = (r == c) ? 1.0 : 0.0;
      } else {
        const int cl = c - u;
        s = cl < l ? Lm[(size_t)r + (size_t)n * (sL + cl)] : -Fv[r];
        if (cl < l) for (int rq = 0; rq < q; rq++) s = fma(-Lm[(size_t)r + (size_t)n * (sQ + rq)], Bq[(size_t)rq + (size_t)q * cl], s);
        Rm[(size_t)r + (size_t)u * cl] = s;
      }
    }
    __syncthreads();
    cta_invert(AUG, u, SCR, IPIV, p.status);
    // U = -K^-1 R (column l: U0 = K^-1 Fu)
    for (long long idx = tid; idx < (long long)u * (l + 1); idx += NT) {
      const int r = (int)(idx % u); const int c = (int)(idx / u);
      double s = 0.0;
      for (int j = 0; j < u; j++) s = fma(AUG[(size_t)r * 2 * u + u + j], Rm[(size_t)j + (size_t)u * c], s);
      Um[idx] = -s;
    }
    __syncthreads();
    // Q = -A U - B (column l: Q0 = -A U0)
    for (long long idx = tid; idx < (long long)q * (l + 1); idx += NT) {
      const int rq = (int)(idx % q); const int c = (int)(idx / q);
      double s = 0.0;
      for (int j = 0; j < u; j++) s = fma(Aq[(size_t)rq + (size_t)q * j], Um[(size_t)j + (size_t)u * c], s);
      Qm[idx] = -s - Bq[idx];
    }
    __syncthreads();
    // write U, Q, U0, Q0 (HDGSolver.cpp:336-341; kept row-major per element on the device, see recover_kernel)
    for (long long idx = tid; idx < (long long)u * l; idx += NT) { const int r = (int)(idx / l), cc = (int)(idx - (long long)r * l); p.U[(size_t)e * u * l + idx] = Um[(size_t)r + (size_t)u * cc]; }
    for (long long idx = tid; idx < (long long)q * l; idx += NT) { const int r = (int)(idx / l), cc = (int)(idx - (long long)r * l); p.Q[(size_t)e * q * l + idx] = Qm[(size_t)r + (size_t)q * cc]; }
    for (int i = tid; i < u; i += NT) p.U0[(size_t)e * u + i] = Um[(size_t)u * l + i];
    for (int i = tid; i < q; i += NT) p.Q0[(size_t)e * q + i] = Qm[(size_t)q * l + i];
    // S = Slu U + Slq Q + Sll ; S0 = Fl - Slu U0 - Slq Q0 ; boundary rows (:489-501) ; scatter (:596-618)
    double* gS = p.S ? p.S + (size_t)e * l * l : nullptr;
    double* gS0 = p.S0 ? p.S0 + (size_t)e * l : nullptr;
    for (long long idx = tid; idx < (long long)l * (l + 1); idx += NT) {
      const int r = (int)(idx % l); const int c = (int)(idx / l);
      const double* lrow = Lm + sL + r;
      double s = c < l ? lrow[(size_t)n * (sL + c)] : 0.0;
      for (int j = 0; j < u; j++) s = fma(lrow[(size_t)n * j], Um[(size_t)j + (size_t)u * c], s);
      for (int rq = 0; rq < q; rq++) s = fma(lrow[(size_t)n * (sQ + rq)], Qm[(size_t)rq + (size_t)q * c], s);
      const int f = r / t, rr = r - f * t, a = rr / nD, k1 = rr - a * nD, F = FACE[f], bc = BCF[f];
      const long long rowOff = ROWS[f] + (long long)(PERM[f * nNf + a] * nD + k1) * t;   // block CSR: row inside each t x t neighbour block
      if (c < l) {
        const int f2 = c / t, cc = c - f2 * t, b = cc / nD, k2 = cc - b * nD;
        if (bc == 1) s = (r == c) ? 1.0 : 0.0;                                                       // DirichletModel: identity row
        else if (bc == 2) s = (f2 == f && k1 == k2) ? FONE[f * nNf * nNf + a * nNf + b] : 0.0;       // IntegratedDirichletModel: face mass (x) I
        if (gS) gS[(size_t)r + (size_t)l * c] = s;
        double* dst = p.vals + rowOff + (long long)POS[f * nFc + f2] * t * t + PERM[f2 * nNf + b] * nD + k2;
        if (f2 == f && INTF[f]) atomicAdd(dst, s); else *dst = s;
      } else {
        double s0 = Fv[sL + r] - s;
        if (bc == 1) s0 = p.dirichlet[((size_t)F * nNf + a) * nD + k1];
        else if (bc == 2) { s0 = 0.0; for (int b = 0; b < nNf; b++) s0 = fma(FONE[f * nNf * nNf + a * nNf + b], p.dirichlet[((size_t)F * nNf + b) * nD + k1], s0); }
        if (gS0) gS0[r] = s0;
        double* dst = p.rhs + ((size_t)F * nNf + PERM[f * nNf + a]) * nD + k1;
        if (INTF[f]) atomicAdd(dst, s0); else *dst = s0;
      }
    }
    __syncthreads();
  }
}

inline size_t gen_smem_bytes(int nN, int nNf, int nFc, int nD) {
  const int u = nN * nD, nmax = u > nN ? u : nN;
  size_t doubles = (size_t)nmax * 2 * nmax + 3 * nmax + 2;
  size_t ints = (size_t)nFc * nNf * 2 + (size_t)nFc * nN + (size_t)nFc * 5 + (size_t)nFc * nFc + 2;
  ints = (ints + 1) & ~(size_t)1;
  return doubles * 8 + ints * 4 + 8 * (size_t)nFc + 16;
}

}  // namespace hfx
