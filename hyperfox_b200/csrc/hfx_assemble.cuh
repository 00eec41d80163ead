// Fused per-element HDG kernel for sm_100a: geometry -> operator contractions -> structured static condensation ->
// Dirichlet masking -> deterministic scatter into the global trace matrix, one CTA per element, everything in shared memory.
//
// Replaces, per element, the reference's HDGSolver::calcElementalMatrices (src/solver/HDGSolver.cpp:176-359) with
// Model::compute (src/model/HDG*.cpp), the operators (src/operator/HDGBase.cpp:67-158, HDGDiffusion.cpp:74-145,
// HDGConvection.cpp:60-104, Reaction.cpp, Source.cpp, Euler.cpp:28-30), applyBoundaryConditions (:361-529, CGType models)
// and assembleSystem (:531-675).
//
// Structure exploited (valid for every in-scope model, checked numerically against the oracle in tests/):
//   S_qq = M (x) I_dim        (HDGBase.cpp:152; no other operator writes the q rows)      -> one nN x nN inverse W = M^-1
//   q-rows / l-rows couple to u,q only through the face nodes                              -> t x t weighted face mass matrices
// All local matrices are column-major ("row index contiguous").
#pragma once
#include <cstdint>

namespace hfx {

struct AsmParams {
  int nCells;
  // mesh
  const double* nodes; const int* cells; const int* cell2face;
  // per-element maps built on device by build_elem_maps (hfx_allocate)
  const uint8_t* fperm;       // [nCells][nFc*nNf] position in faces[F] of element-local face node (HDGSolver.cpp:258-275)
  const uint8_t* tauSide;     // [nCells][nFc]     0 if this cell is face2Cell[F][0] (HDGSolver.cpp:290-293)
  const uint8_t* elemPos;     // [nCells][nFc*nFc] position of face f2 in the sorted neighbour list of face f
  // per-face
  const long long* faceRowStart;  // [nFaces] offset of row (F,0) in vals
  const uint8_t* faceNnb;         // [nFaces] number of neighbour faces (row length = nnb*t)
  const uint8_t* faceBC;          // [nFaces] 0 none, 1 Dirichlet, 2 integrated Dirichlet
  const uint8_t* faceInterior;    // [nFaces] 1 if two adjacent cells
  // fields
  const double* tau; int tauVals;
  const double* diff; int diffComps; int diffIsCell;
  const double* vel;
  const double* srcIP; const double* reacIP;
  const double* solOld;
  const double* dirichlet;
  int opmask; int timeScheme;
  // tables (device global, read-only)
  const double* shape; const double* dshape; const double* w;
  const double* fshape; const double* fdshape; const double* fw;
  const double* ffs;          // [nIPf][nNf*nNf] phi_a*phi_b products of the face element
  const int* faceNodes;       // [nFc][nNf]
  const int8_t* nodeInFace;   // [nFc][nN] inverse of faceNodes (-1 if not on the face)
  // outputs
  double* U; double* Q; double* U0; double* Q0; double* S; double* S0;  // S,S0 may be NULL
  double* vals; double* rhs;
  int* status;                // bit 0: (near) zero pivot met in a local inverse
};

template <int DIM, int P> struct ElemCfg;
#define HFX_CFG(D, PP, NN, NNF, NIP, NIPF) \
  template <> struct ElemCfg<D, PP> { static constexpr int nN = NN, nNf = NNF, nIP = NIP, nIPf = NIPF, nFc = D + 1; };
// nN, nNf, nIP (degree 2p), nIPf   (Cubature.cpp nIP map)
HFX_CFG(2, 1, 3, 2, 3, 2)
HFX_CFG(2, 2, 6, 3, 6, 3)
HFX_CFG(2, 3, 10, 4, 12, 4)
HFX_CFG(2, 4, 15, 5, 16, 5)
HFX_CFG(2, 5, 21, 6, 25, 6)
HFX_CFG(3, 1, 4, 3, 4, 3)
HFX_CFG(3, 2, 10, 6, 14, 6)
HFX_CFG(3, 3, 20, 10, 24, 12)
HFX_CFG(3, 4, 35, 15, 46, 16)
HFX_CFG(3, 5, 56, 21, 81, 25)
#undef HFX_CFG

constexpr int kAsmThreads = 256;

// ---- register-tiled batched GEMM on shared-memory operands -----------------------------------------------------------
// C_b(m,n) = sum_k A_b(m,k) * B_b(k,n) for b<BATCH, m<M, n<N; fa(b,m,k)/fb(b,k,n) are loaders, fs(b,m,n,acc) the epilogue.
// Tiles are dealt to `nt` threads; consecutive threads take consecutive m-tiles so that column-major A operands are read
// conflict-free and B operands are warp-broadcasts.
template <int BATCH, int M, int N, int K, int TM, int TN, class FA, class FB, class FS>
__device__ __forceinline__ void tile_gemm(int tid, int nt, FA fa, FB fb, FS fs) {
  constexpr int MT = (M + TM - 1) / TM, NTT = (N + TN - 1) / TN;
  for (int tile = tid; tile < BATCH * MT * NTT; tile += nt) {
    const int bt = tile / (MT * NTT), tl = tile % (MT * NTT);
    const int m0 = (tl % MT) * TM, n0 = (tl / MT) * TN;
    double acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
      for (int j = 0; j < TN; j++) acc[i][j] = 0.0;
#pragma unroll 4
    for (int k = 0; k < K; k++) {
      double a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i++) a[i] = (M % TM == 0 || m0 + i < M) ? fa(bt, m0 + i, k) : 0.0;
#pragma unroll
      for (int j = 0; j < TN; j++) b[j] = (N % TN == 0 || n0 + j < N) ? fb(bt, k, n0 + j) : 0.0;
#pragma unroll
      for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
      for (int j = 0; j < TN; j++)
        if ((M % TM == 0 || m0 + i < M) && (N % TN == 0 || n0 + j < N)) fs(bt, m0 + i, n0 + j, acc[i][j]);
  }
}

// ---- in-register Gauss-Jordan inverse by one warp ----------------------------------------------------------------------
// Lane r holds rows r, r+32, ... of the n x n matrix (column-major in shared memory, leading dimension n); unpivoted
// (the local matrices M and K are definite for a coercive HDG local problem); *flag |= 1 if a pivot underflows.
template <int n>
__device__ __forceinline__ void warp_invert(const double* __restrict__ src, double* __restrict__ dst, int lane, int* flag) {
  constexpr int R = (n + 31) / 32;
  double row[R][n];
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int i = lane + 32 * r;
#pragma unroll
    for (int j = 0; j < n; j++) row[r][j] = (i < n) ? src[i + n * j] : ((i == j) ? 1.0 : 0.0);
  }
  double scale = 0.0;
#pragma unroll
  for (int r = 0; r < R; r++)
#pragma unroll
    for (int j = 0; j < n; j++) scale = fmax(scale, fabs(row[r][j]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) scale = fmax(scale, __shfl_xor_sync(0xffffffffu, scale, o));
  bool bad = false;
#pragma unroll
  for (int k = 0; k < n; k++) {
    const int kl = k & 31, kr = k >> 5;
    double piv = __shfl_sync(0xffffffffu, row[kr][k], kl);
    if (!(fabs(piv) > 1e-14 * scale)) bad = true;
    const double ip = 1.0 / piv;
    double f[R];
#pragma unroll
    for (int r = 0; r < R; r++) f[r] = row[r][k] * ip;
#pragma unroll
    for (int j = 0; j < n; j++) {
      const double pj = __shfl_sync(0xffffffffu, row[kr][j], kl);  // pivot row entry (before update)
#pragma unroll
      for (int r = 0; r < R; r++) {
        const bool isPivotRow = (r == kr) && (lane == kl);
        if (j == k) row[r][j] = isPivotRow ? ip : -f[r];
        else row[r][j] = isPivotRow ? pj * ip : fma(-f[r], pj, row[r][j]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int i = lane + 32 * r;
    if (i < n) {
#pragma unroll
      for (int j = 0; j < n; j++) dst[i + n * j] = row[r][j];
    }
  }
  if (bad && lane == 0) atomicOr(flag, 1);
}

__device__ __forceinline__ void det_inv(const double (&J)[2][2], double& det, double (&I)[2][2]) {
  det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
  const double id = 1.0 / det;
  I[0][0] = J[1][1] * id; I[0][1] = -J[0][1] * id; I[1][0] = -J[1][0] * id; I[1][1] = J[0][0] * id;
}
__device__ __forceinline__ void det_inv(const double (&J)[3][3], double& det, double (&I)[3][3]) {
  const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1], c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2], c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
  det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
  const double id = 1.0 / det;
  I[0][0] = c00 * id; I[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id; I[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
  I[1][0] = c01 * id; I[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id; I[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
  I[2][0] = c02 * id; I[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id; I[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
}

template <int DIM, int P>
struct AsmSmem {
  using C = ElemCfg<DIM, P>;
  static constexpr int nN = C::nN, t = C::nNf, nFc = C::nFc, nIP = C::nIP, nIPf = C::nIPf;
  static constexpr int l = nFc * t, l1 = l + 1, NW = 3 + 2 * DIM;  // face weight kinds: tau, n_d, (Dn)_d, v.n, 1
  // offsets in doubles
  static constexpr int oX = 0;                               // coords [nN][DIM]
  static constexpr int oIJ = oX + nN * DIM;                  // dV * invJ  [nIP][DIM*DIM]  (m,r)
  static constexpr int oDV = oIJ + nIP * DIM * DIM;          // dV [nIP]
  static constexpr int oDIP = oDV + nIP;                     // D at bulk IPs [nIP][DIM*DIM] col-major
  static constexpr int oVIP = oDIP + nIP * DIM * DIM;        // v at bulk IPs [nIP][DIM]
  static constexpr int oLW = oVIP + nIP * DIM;               // (reac*dV + euler*dV) [nIP], src*dV [nIP]
  static constexpr int oFWT = oLW + 2 * nIP;                 // face IP weights [nFc*nIPf][NW]
  static constexpr int oTAU = oFWT + nFc * nIPf * NW;        // tau at element-local face nodes [l]
  static constexpr int oDN = oTAU + l;                       // D at nodes [nN][DIM*DIM]
  static constexpr int oVN = oDN + nN * DIM * DIM;           // v at nodes [nN][DIM]
  static constexpr int oG = oVN + nN * DIM;                  // g [nIP][DIM][nN]   -> later A_d [DIM][nN x nN]
  static constexpr int szG = (nIP * DIM * nN > DIM * nN * nN) ? nIP * DIM * nN : DIM * nN * nN;
  static constexpr int oCG = oG + szG;                       // suu left operand [nIP][nN]
  static constexpr int oM = oCG + nIP * nN;                  // M
  static constexpr int oW = oM + nN * nN;                    // W = M^-1
  static constexpr int oSQU = oW + nN * nN;                  // Squ_d [DIM][nN x nN] -> later U [nN x l1]
  static constexpr int szSQU = (DIM * nN * nN > nN * l1) ? DIM * nN * nN : nN * l1;
  static constexpr int oSUQ = oSQU + szSQU;                  // Suq_d [DIM][nN x nN]
  static constexpr int oSUU = oSUQ + DIM * nN * nN;          // Suu -> K -> K^-1
  static constexpr int oFW = oSUU + nN * nN;                 // weighted face mass matrices [nFc][NW][t x t]
  static constexpr int oB = oFW + nFc * NW * t * t;          // B_d [DIM][nN x l1] -> Q_d
  static constexpr int oR = oB + DIM * nN * l1;              // R [nN x l1]
  static constexpr int oFU = oR + nN * l1;                   // Fu [nN]
  static constexpr int oEnd = oFU + nN;
  static constexpr int nDoubles = oEnd;
  // after the doubles: row starts (nFc x int64) then a small int area
  static constexpr int nInts = 8 + 2 * nFc * t + nFc * nN + nFc * nFc + 3 * nFc + 8;
  static constexpr size_t bytes = (size_t)nDoubles * 8 + 8 * nFc + 4 * (size_t)nInts;
};

template <int DIM, int P>
__global__ void __launch_bounds__(kAsmThreads, 2) hdg_assemble_kernel(const AsmParams p) {
  using L = AsmSmem<DIM, P>;
  constexpr int nN = L::nN, t = L::t, nFc = L::nFc, nIP = L::nIP, nIPf = L::nIPf, l = L::l, l1 = L::l1, NW = L::NW;
  constexpr int D2 = DIM * DIM, NT = kAsmThreads;
  constexpr int kTau = 0, kN = 1, kDN = 1 + DIM, kC = 1 + 2 * DIM, kOne = 2 + 2 * DIM;
  extern __shared__ double sm[];
  double* X = sm + L::oX; double* IJ = sm + L::oIJ; double* DV = sm + L::oDV; double* DIP = sm + L::oDIP; double* VIP = sm + L::oVIP;
  double* LW = sm + L::oLW; double* FWT = sm + L::oFWT; double* TAU = sm + L::oTAU; double* DN = sm + L::oDN; double* VN = sm + L::oVN;
  double* G = sm + L::oG; double* CG = sm + L::oCG; double* Mm = sm + L::oM; double* W = sm + L::oW; double* SQU = sm + L::oSQU;
  double* SUQ = sm + L::oSUQ; double* SUU = sm + L::oSUU; double* FW = sm + L::oFW; double* B = sm + L::oB; double* R = sm + L::oR;
  double* FU = sm + L::oFU;
  long long* ROWS = reinterpret_cast<long long*>(sm + L::nDoubles);   // [nFc] first entry of row (F,0) in vals
  int* ISM = reinterpret_cast<int*>(ROWS + nFc);                      // [nFc] global face ids
  int* FN = ISM + 8;                                                  // [nFc*t] faceNodes
  int* PERM = FN + nFc * t;                                           // [nFc*t] element-local -> face-node position
  int* NIF = PERM + nFc * t;                                          // [nFc*nN] node -> position in face (or -1)
  int* POS = NIF + nFc * nN;                                          // [nFc*nFc]
  int* RLEN = POS + nFc * nFc;                                        // [nFc] row length
  int* BCF = RLEN + nFc;                                              // [nFc] boundary kind
  int* INTF = BCF + nFc;                                              // [nFc] interior flag
  double* A = G;    // A_d aliases g (dead after the contractions)
  double* Um = SQU; // U aliases Squ (dead after A)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool hasDiff = p.opmask & 1, hasConv = p.opmask & 2, hasReac = (p.opmask & 4) && p.reacIP, hasSrc = (p.opmask & 8) && p.srcIP;
  const bool euler = p.timeScheme == 1;
  const bool diffField = hasDiff && p.diffComps > 0 && p.diff;

  for (int i = tid; i < nFc * t; i += NT) FN[i] = p.faceNodes[i];
  for (int i = tid; i < nFc * nN; i += NT) NIF[i] = p.nodeInFace[i];
  __syncthreads();

  for (int e = blockIdx.x; e < p.nCells; e += gridDim.x) {
    // ---- P0: gather ------------------------------------------------------------------------------------------------
    const int* cell = p.cells + (size_t)e * nN;
    for (int i = tid; i < nN * DIM; i += NT) X[i] = p.nodes[(size_t)cell[i / DIM] * DIM + (i % DIM)];
    if (tid < nFc) {
      const int F = p.cell2face[(size_t)e * nFc + tid];
      ISM[tid] = F;
      ROWS[tid] = p.faceRowStart[F];
      RLEN[tid] = (int)p.faceNnb[F] * t;
      BCF[tid] = p.faceBC[F];
      INTF[tid] = p.faceInterior[F];
    }
    for (int i = tid; i < nFc * nFc; i += NT) POS[i] = p.elemPos[(size_t)e * nFc * nFc + i];
    for (int i = tid; i < l; i += NT) {
      const int f = i / t;
      const int F = p.cell2face[(size_t)e * nFc + f];
      const int pos = p.fperm[(size_t)e * l + i];
      PERM[i] = pos;
      const int side = (p.tauVals == 2) ? p.tauSide[(size_t)e * nFc + f] : 0;
      TAU[i] = p.tau[((size_t)F * t + pos) * p.tauVals + side];
    }
    if (diffField) {
      for (int i = tid; i < nN * D2; i += NT) {
        const int nd = i / D2, c = i % D2;
        const size_t ent = p.diffIsCell ? ((size_t)e * nN + nd) : (size_t)cell[nd];
        double v;
        if (p.diffComps == 1) v = ((c / DIM) == (c % DIM)) ? p.diff[ent] : 0.0;
        else v = p.diff[ent * D2 + c];
        DN[i] = v;
      }
    }
    if (hasConv) for (int i = tid; i < nN * DIM; i += NT) VN[i] = p.vel[(size_t)cell[i / DIM] * DIM + (i % DIM)];
    __syncthreads();

    // ---- P1: geometry at bulk and face integration points (Operator.cpp:14-84, HDGModel.cpp:53-85, HDGBase.cpp:18-65) --
    for (int k = tid; k < nIP + nFc * nIPf; k += NT) {
      if (k < nIP) {
        const int ip = k;
        double J[DIM][DIM];
#pragma unroll
        for (int r = 0; r < DIM; r++)
#pragma unroll
          for (int m = 0; m < DIM; m++) J[r][m] = 0.0;
        for (int i = 0; i < nN; i++) {
          const double* d = p.dshape + ((size_t)ip * nN + i) * DIM;
#pragma unroll
          for (int r = 0; r < DIM; r++) {
            const double dr = __ldg(d + r);
#pragma unroll
            for (int m = 0; m < DIM; m++) J[r][m] = fma(dr, X[i * DIM + m], J[r][m]);
          }
        }
        double det, I[DIM][DIM];
        det_inv(J, det, I);
        const double dv = __ldg(p.w + ip) * det;
        DV[ip] = dv;
#pragma unroll
        for (int m = 0; m < DIM; m++)
#pragma unroll
          for (int r = 0; r < DIM; r++) IJ[ip * D2 + m * DIM + r] = I[m][r] * dv;
        if (diffField) {
          double Dc[D2];
#pragma unroll
          for (int c = 0; c < D2; c++) Dc[c] = 0.0;
          for (int i = 0; i < nN; i++) {
            const double s = __ldg(p.shape + (size_t)ip * nN + i);
#pragma unroll
            for (int c = 0; c < D2; c++) Dc[c] = fma(s, DN[i * D2 + c], Dc[c]);
          }
#pragma unroll
          for (int c = 0; c < D2; c++) DIP[ip * D2 + c] = Dc[c];
        }
        if (hasConv) {
          double v[DIM];
#pragma unroll
          for (int d = 0; d < DIM; d++) v[d] = 0.0;
          for (int i = 0; i < nN; i++) {
            const double s = __ldg(p.shape + (size_t)ip * nN + i);
#pragma unroll
            for (int d = 0; d < DIM; d++) v[d] = fma(s, VN[i * DIM + d], v[d]);
          }
#pragma unroll
          for (int d = 0; d < DIM; d++) VIP[ip * DIM + d] = v[d];
        }
        double lw = 0.0;
        if (hasReac) lw += p.reacIP[(size_t)e * nIP + ip] * dv;
        if (euler) lw += dv;
        LW[ip] = lw;
        LW[nIP + ip] = hasSrc ? p.srcIP[(size_t)e * nIP + ip] * dv : 0.0;
      } else {
        const int fi = k - nIP, f = fi / nIPf, ip = fi % nIPf;
        const int* fn = FN + f * t;
        double J[DIM - 1][DIM];
#pragma unroll
        for (int r = 0; r < DIM - 1; r++)
#pragma unroll
          for (int m = 0; m < DIM; m++) J[r][m] = 0.0;
        double tauip = 0.0, Dc[D2], v[DIM];
#pragma unroll
        for (int c = 0; c < D2; c++) Dc[c] = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; d++) v[d] = 0.0;
        for (int a = 0; a < t; a++) {
          const int nd = fn[a];
          const double* d = p.fdshape + ((size_t)ip * t + a) * (DIM - 1);
          const double s = __ldg(p.fshape + (size_t)ip * t + a);
#pragma unroll
          for (int r = 0; r < DIM - 1; r++) {
            const double dr = __ldg(d + r);
#pragma unroll
            for (int m = 0; m < DIM; m++) J[r][m] = fma(dr, X[nd * DIM + m], J[r][m]);
          }
          tauip = fma(s, TAU[f * t + a], tauip);
          if (diffField) {
#pragma unroll
            for (int c = 0; c < D2; c++) Dc[c] = fma(s, DN[nd * D2 + c], Dc[c]);
          }
          if (hasConv) {
#pragma unroll
            for (int d2 = 0; d2 < DIM; d2++) v[d2] = fma(s, VN[nd * DIM + d2], v[d2]);
          }
        }
        double nv[DIM], area;
        if (DIM == 2) {
          nv[0] = -J[0][1]; nv[1] = J[0][0];
          area = sqrt(J[0][0] * J[0][0] + J[0][1] * J[0][1]);
        } else {
          nv[0] = J[0][1] * J[DIM - 2][2 % DIM] - J[0][2 % DIM] * J[DIM - 2][1];
          nv[1] = J[0][2 % DIM] * J[DIM - 2][0] - J[0][0] * J[DIM - 2][2 % DIM];
          nv[DIM - 1] = J[0][0] * J[DIM - 2][1] - J[0][1] * J[DIM - 2][0];
          // sqrt(det(J J^T)) (Operator.cpp:66-69)
          const double g00 = J[0][0] * J[0][0] + J[0][1] * J[0][1] + J[0][2 % DIM] * J[0][2 % DIM];
          const double g11 = J[DIM - 2][0] * J[DIM - 2][0] + J[DIM - 2][1] * J[DIM - 2][1] + J[DIM - 2][2 % DIM] * J[DIM - 2][2 % DIM];
          const double g01 = J[0][0] * J[DIM - 2][0] + J[0][1] * J[DIM - 2][1] + J[0][2 % DIM] * J[DIM - 2][2 % DIM];
          area = sqrt(g00 * g11 - g01 * g01);
        }
        double nrm = 0.0;
#pragma unroll
        for (int m = 0; m < DIM; m++) nrm = fma(nv[m], nv[m], nrm);
        nrm = sqrt(nrm);
        // outward orientation: (x_opposite - x_v0) . n <= 0   (HDGBase.cpp:43-62)
        const int v0 = fn[0];
        int vn = 0;
        for (int kk = 0; kk < nN; kk++) if (NIF[f * nN + kk] < 0) { vn = kk; break; }
        double prod = 0.0;
#pragma unroll
        for (int m = 0; m < DIM; m++) { nv[m] /= nrm; prod = fma(X[vn * DIM + m] - X[v0 * DIM + m], nv[m], prod); }
        if (prod > 0.0) {
#pragma unroll
          for (int m = 0; m < DIM; m++) nv[m] = -nv[m];
        }
        const double dvf = __ldg(p.fw + ip) * area;
        double* wt = FWT + (size_t)fi * NW;
        wt[kTau] = dvf * tauip;
        double vdn = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; d++) {
          wt[kN + d] = dvf * nv[d];
          double dn = nv[d];
          if (diffField) {
            dn = 0.0;
#pragma unroll
            for (int b = 0; b < DIM; b++) dn = fma(Dc[b * DIM + d], nv[b], dn);  // (D n)_d, D col-major
          }
          wt[kDN + d] = hasDiff ? dvf * dn : 0.0;
          vdn = fma(v[d], nv[d], vdn);
        }
        wt[kC] = hasConv ? dvf * vdn : 0.0;
        wt[kOne] = dvf;
      }
    }
    __syncthreads();

    // ---- P2: g[ip][d][i] = dV (J^-1 grad_ref phi_i)_d ; cg = suu left operand ----------------------------------------
    for (int idx = tid; idx < nIP * nN; idx += NT) {
      const int ip = idx / nN, i = idx % nN;
      const double* d = p.dshape + ((size_t)ip * nN + i) * DIM;
      double dr[DIM], gg[DIM];
#pragma unroll
      for (int r = 0; r < DIM; r++) dr[r] = __ldg(d + r);
#pragma unroll
      for (int m = 0; m < DIM; m++) {
        double s = 0.0;
#pragma unroll
        for (int r = 0; r < DIM; r++) s = fma(IJ[ip * D2 + m * DIM + r], dr[r], s);
        gg[m] = s;
        G[(ip * DIM + m) * nN + i] = s;
      }
      double c = LW[ip] * __ldg(p.shape + (size_t)ip * nN + i);
      if (hasConv) {
#pragma unroll
        for (int m = 0; m < DIM; m++) c = fma(-VIP[ip * DIM + m], gg[m], c);
      }
      CG[ip * nN + i] = c;
    }
    __syncthreads();

    // ---- P3a: M = sum_ip dV phi phi^T ---------------------------------------------------------------------------------
    tile_gemm<1, nN, nN, nIP, 2, 1>(tid, NT,
        [&](int, int m, int k) { return DV[k] * __ldg(p.shape + (size_t)k * nN + m); },
        [&](int, int k, int n) { return __ldg(p.shape + (size_t)k * nN + n); },
        [&](int, int m, int n, double v) { Mm[m + nN * n] = v; });
    __syncthreads();

    // ---- P3b: warp 0 inverts M while the other warps do the remaining contractions ---------------------------------
    if (warp == 0) {
      warp_invert<nN>(Mm, W, lane, p.status);
    } else {
      const int t2 = tid - 32, nt2 = NT - 32;
      // Squ_d[k][j] = sum_ip g[ip][d][k] phi[ip][j]     rows m = (d,k)   (HDGBase.cpp:150)
      tile_gemm<1, DIM * nN, nN, nIP, 2, 2>(t2, nt2,
          [&](int, int m, int k) { return G[k * DIM * nN + m]; },
          [&](int, int k, int n) { return __ldg(p.shape + (size_t)k * nN + n); },
          [&](int, int m, int n, double v) { const int d = m / nN, kk = m % nN; SQU[(d * nN + n) * nN + kk] = v; });
      // Suu (bulk part): -C^T (Convection.cpp:5-49) + reaction mass + Euler mass
      tile_gemm<1, nN, nN, nIP, 2, 2>(t2, nt2,
          [&](int, int m, int k) { return CG[k * nN + m]; },
          [&](int, int k, int n) { return __ldg(p.shape + (size_t)k * nN + n); },
          [&](int, int m, int n, double v) { SUU[m + nN * n] = v; });
      // weighted face mass matrices FW[f][kind][a + t b] = sum_ip wt[f][ip][kind] phi_a phi_b
      tile_gemm<1, t * t, nFc * NW, nIPf, 2, 2>(t2, nt2,
          [&](int, int m, int k) { return __ldg(p.ffs + (size_t)k * t * t + m); },
          [&](int, int k, int n) { const int f = n / NW, kind = n % NW; return FWT[(size_t)(f * nIPf + k) * NW + kind]; },
          [&](int, int m, int n, double v) { FW[(size_t)n * t * t + m] = v; });
      // Fu = source (Source.cpp:24-48)
      for (int i = t2; i < nN; i += nt2) {
        double s = 0.0;
        if (hasSrc) for (int ip = 0; ip < nIP; ip++) s = fma(__ldg(p.shape + (size_t)ip * nN + i), LW[nIP + ip], s);
        FU[i] = s;
      }
    }
    __syncthreads();

    // ---- P3c: Suq bulk part (HDGDiffusion.cpp:130-144).  D = I: identical to Squ; no diffusion: zero ---------------------
    if (diffField) {
      for (int idx = tid; idx < nIP * nN; idx += NT) {   // g <- D g in place
        const int ip = idx / nN, i = idx % nN;
        double gg[DIM], o[DIM];
#pragma unroll
        for (int m = 0; m < DIM; m++) gg[m] = G[(ip * DIM + m) * nN + i];
#pragma unroll
        for (int a = 0; a < DIM; a++) {
          double s = 0.0;
#pragma unroll
          for (int b = 0; b < DIM; b++) s = fma(DIP[ip * D2 + b * DIM + a], gg[b], s);
          o[a] = s;
        }
#pragma unroll
        for (int m = 0; m < DIM; m++) G[(ip * DIM + m) * nN + i] = o[m];
      }
      __syncthreads();
      tile_gemm<1, DIM * nN, nN, nIP, 2, 2>(tid, NT,
          [&](int, int m, int k) { return G[k * DIM * nN + m]; },
          [&](int, int k, int n) { return __ldg(p.shape + (size_t)k * nN + n); },
          [&](int, int m, int n, double v) { const int d = m / nN, i = m % nN; SUQ[(d * nN + n) * nN + i] = v; });
    } else {
      // Suq_d[i][j] (bulk) equals Squ_d[i][j] entry by entry when D = I
      for (int idx = tid; idx < DIM * nN * nN; idx += NT) SUQ[idx] = hasDiff ? SQU[idx] : 0.0;
    }
    // Euler: Fu += Mass * Solution_old (Euler.cpp:29-30)
    if (euler && tid < nN) {
      double s = FU[tid];
      const double* so = p.solOld + (size_t)e * nN;
      for (int j = 0; j < nN; j++) s = fma(Mm[tid + nN * j], so[j], s);
      FU[tid] = s;
    }
    __syncthreads();
    // face parts of Suu (+tau mass, HDGBase.cpp:128) and Suq (-(Dn) mass, HDGDiffusion.cpp:121), gather form
    for (int idx = tid; idx < nN * nN; idx += NT) {
      const int i = idx % nN, j = idx / nN;
      double suu = SUU[idx], suq[DIM];
#pragma unroll
      for (int d = 0; d < DIM; d++) suq[d] = SUQ[d * nN * nN + idx];
      for (int f = 0; f < nFc; f++) {
        const int a = NIF[f * nN + i], b = NIF[f * nN + j];
        if (a >= 0 && b >= 0) {
          const double* fw = FW + (size_t)f * NW * t * t + a + t * b;
          suu += fw[kTau * t * t];
#pragma unroll
          for (int d = 0; d < DIM; d++) suq[d] -= fw[(kDN + d) * t * t];
        }
      }
      SUU[idx] = suu;
#pragma unroll
      for (int d = 0; d < DIM; d++) SUQ[d * nN * nN + idx] = suq[d];
    }
    __syncthreads();

    // ---- P4: A_d = W Squ_d ;  B_d = W Sql_d with Sql[(fn_f(a),d),(f,b)] = -N_fd[a][b] (HDGBase.cpp:134) -----------------
    tile_gemm<1, nN, DIM * nN, nN, 2, 3>(tid, NT,
        [&](int, int m, int k) { return W[m + nN * k]; },
        [&](int, int k, int n) { return SQU[n * nN + k]; },          // n = (d,j)
        [&](int, int m, int n, double v) { A[n * nN + m] = v; });    // A[(d*nN + j)*nN + i] = A_d[i][j]
    tile_gemm<nFc, nN, DIM * t, t, 2, (t % 5 == 0 ? 5 : (t % 3 == 0 ? 3 : (t % 2 == 0 ? 2 : 1)))>(tid, NT,
        [&](int f, int m, int k) { return W[m + nN * FN[f * t + k]]; },
        [&](int f, int k, int n) { const int d = n / t, b = n % t; return FW[((size_t)f * NW + kN + d) * t * t + k + t * b]; },
        [&](int f, int m, int n, double v) { const int d = n / t, b = n % t; B[(size_t)d * nN * l1 + m + nN * (f * t + b)] = -v; });
    for (int idx = tid; idx < DIM * nN; idx += NT) B[(size_t)(idx / nN) * nN * l1 + (idx % nN) + nN * l] = 0.0;  // Q0 column
    __syncthreads();

    // ---- P5: K = Suu - sum_d Suq_d A_d  (HDGSolver.cpp:335) ------------------------------------------------------------
    tile_gemm<1, nN, nN, DIM * nN, 2, 1>(tid, NT,
        [&](int, int m, int k) { return SUQ[k * nN + m]; },                                   // k = (d,k')
        [&](int, int k, int n) { const int d = k / nN, kk = k % nN; return A[(d * nN + n) * nN + kk]; },
        [&](int, int m, int n, double v) { SUU[m + nN * n] -= v; });
    __syncthreads();

    // ---- P6: warp 0 inverts K in place; the others form R = Sul - sum_d Suq_d B_d, last column -Fu (:342-343) ----------
    if (warp == 0) {
      warp_invert<nN>(SUU, SUU, lane, p.status);
    } else {
      const int t2 = tid - 32, nt2 = NT - 32;
      tile_gemm<1, nN, l, DIM * nN, 2, 2>(t2, nt2,
          [&](int, int m, int k) { return SUQ[k * nN + m]; },
          [&](int, int k, int n) { const int d = k / nN, j = k % nN; return B[(size_t)d * nN * l1 + j + nN * n]; },
          [&](int, int m, int n, double v) {
            const int f = n / t, b = n % t, a = NIF[f * nN + m];
            double sul = 0.0;
            if (a >= 0) { const double* fw = FW + (size_t)f * NW * t * t + a + t * b; sul = fw[kC * t * t] - fw[kTau * t * t]; }
            R[m + nN * n] = sul - v;
          });
      for (int i = t2; i < nN; i += nt2) R[i + nN * l] = -FU[i];
    }
    __syncthreads();

    // ---- P7: U = -K^-1 R ; U0 = K^-1 Fu ------------------------------------------------------------------------------
    {
      double* gU = p.U + (size_t)e * nN * l;
      double* gU0 = p.U0 + (size_t)e * nN;
      tile_gemm<1, nN, l1, nN, 2, 2>(tid, NT,
          [&](int, int m, int k) { return SUU[m + nN * k]; },
          [&](int, int k, int n) { return R[k + nN * n]; },
          [&](int, int m, int n, double v) {
            Um[m + nN * n] = -v;
            if (n < l) gU[m + nN * n] = -v; else gU0[m] = -v;
          });
    }
    __syncthreads();

    // ---- P8: Q_d = -A_d U - B_d ; Q0_d = -A_d U0  (:344-345) -----------------------------------------------------------
    {
      double* gQ = p.Q + (size_t)e * (DIM * nN) * l;
      double* gQ0 = p.Q0 + (size_t)e * (DIM * nN);
      tile_gemm<DIM, nN, l1, nN, 2, 4>(tid, NT,
          [&](int d, int m, int k) { return A[(d * nN + k) * nN + m]; },
          [&](int, int k, int n) { return Um[k + nN * n]; },
          [&](int d, int m, int n, double v) {
            double* bq = B + (size_t)d * nN * l1 + m + nN * n;
            const double qv = -v - *bq;
            *bq = qv;
            if (n < l) gQ[(m * DIM + d) + (size_t)(DIM * nN) * n] = qv; else gQ0[m * DIM + d] = qv;
          });
    }
    __syncthreads();

    // ---- P9: S = Slu U + Slq Q + Sll ; S0 = Fl - Slu U0 - Slq Q0 (:347-348); Dirichlet rows (:489-501); scatter (:596-618) --
    {
      double* gS = p.S ? p.S + (size_t)e * l * l : nullptr;
      double* gS0 = p.S0 ? p.S0 + (size_t)e * l : nullptr;
      tile_gemm<nFc, t, l1, (1 + DIM) * t, 2, 2>(tid, NT,
          [&](int f, int m, int k) {
            const int kind = k / t, b = k % t;
            const double* fw = FW + (size_t)f * NW * t * t + m + t * b;
            return kind == 0 ? fw[kTau * t * t] : -fw[(kDN + kind - 1) * t * t];
          },
          [&](int f, int k, int n) {
            const int kind = k / t, b = k % t, nd = FN[f * t + b];
            return kind == 0 ? Um[nd + nN * n] : B[(size_t)(kind - 1) * nN * l1 + nd + nN * n];
          },
          [&](int f, int a, int n, double v) {
            const int F = ISM[f], bc = BCF[f];
            const bool inter = INTF[f];
            const int rowDof = F * t + PERM[f * t + a];
            if (n == l) {   // S0
              double s0 = -v;
              if (bc == 1) s0 = p.dirichlet[(size_t)F * t + a];
              else if (bc == 2) {
                s0 = 0.0;
                for (int b = 0; b < t; b++) s0 = fma(FW[((size_t)f * NW + kOne) * t * t + a + t * b], p.dirichlet[(size_t)F * t + b], s0);
              }
              if (gS0) gS0[f * t + a] = s0;
              if (inter) atomicAdd(p.rhs + rowDof, s0); else p.rhs[rowDof] = s0;
              return;
            }
            const int f2 = n / t, b2 = n % t;
            double sv = v;
            if (f2 == f) { const double* fw = FW + (size_t)f * NW * t * t + a + t * b2; sv += fw[kC * t * t] - fw[kTau * t * t]; }
            if (bc == 1) sv = (f2 == f && b2 == a) ? 1.0 : 0.0;
            else if (bc == 2) sv = (f2 == f) ? FW[((size_t)f * NW + kOne) * t * t + a + t * b2] : 0.0;
            if (gS) gS[(f * t + a) + (size_t)l * n] = sv;
            double* dst = p.vals + ROWS[f] + (long long)PERM[f * t + a] * RLEN[f] + POS[f * nFc + f2] * t + PERM[f2 * t + b2];
            if (f2 == f && inter) atomicAdd(dst, sv); else *dst = sv;
          });
    }
    __syncthreads();
  }
}

// host-side launch helper: returns false if (dim, order) has no shared-memory-resident instantiation
template <int DIM, int P>
inline cudaError_t launch_assemble_t(const AsmParams& p, int nSM, cudaStream_t st) {
  using L = AsmSmem<DIM, P>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(hdg_assemble_kernel<DIM, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  int perSM = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, hdg_assemble_kernel<DIM, P>, kAsmThreads, L::bytes);
  if (perSM < 1) perSM = 1;
  long long grid = (long long)nSM * perSM;
  if (grid > p.nCells) grid = p.nCells;
  if (grid < 1) grid = 1;
  hdg_assemble_kernel<DIM, P><<<(int)grid, kAsmThreads, L::bytes, st>>>(p);
  return cudaGetLastError();
}

}  // namespace hfx
