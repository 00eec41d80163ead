// Fused per-element HDG kernel for LARGE straight-sided simplices (3-D order 4: 35 nodes, 15 face nodes, local system 35 + 105 + 60),
// one 512-thread CTA per SM, one element per CTA pass, everything in shared memory.  The element image of hfx_assemble.cuh does not fit
// twice (nor once) at this size, so the algebra is rearranged around what a straight-sided element with D = c I offers
// (Operator.cpp:14-84 constant Jacobian; HDGBase.cpp:67-158, HDGDiffusion.cpp:74-145, HDGConvection.cpp:60-104, Reaction.cpp, Source.cpp,
// Euler.cpp:28-30; condensation HDGSolver.cpp:331-348; boundary rows :361-529; scatter :531-675):
//
//   A_d = Sqq^-1 Squ_d = sum_r Jinv(d,r) A^_r        B_d = Sqq^-1 Sql_d = -(area_f n_fd / detJ) B^_f        (reference matrices A^_r, B^_f)
//   SJ_r := sum_d Jinv(d,r) Suq_d   (formed directly from S^_r^T and the face masses: Suq_d itself is never stored)
//   K = Suu - sum_r SJ_r A^_r                          R_f = Sul_f + (area_f / detJ) sum_r nu_fr (SJ_r B^_f),  nu_fr = sum_d J(r,d) n_fd
//   U = -K^-1 [R | -Fu]  (+ one refinement step)       Q_d = -sum_r Jinv(d,r) (A^_r U) + (area_f n_fd / detJ) B^_f
//   S_f = FT_f (U_f - I_f) + (area_f M^f) Zq_f + FC_f I_f,   Zq_f = -c sum_d n_fd Q_d[faceNodes_f]   (reduction length 2t instead of (1+dim) t)
//
// so that A^_r and B^ stay RESIDENT in shared memory for the whole kernel (they are the right operands of K, R and the left operand of Q) and
// the per-element image is SJ (3 nN^2), K (x2 for the ping-pong Gauss-Jordan), R, U, Q and the face matrices.  Every product runs on the FP64
// tensor cores (mma.sync.m8n8k4.f64 -> DMMA), three accumulator chains (one per reference direction r) per warp task.
// Operand layouts are chosen so that both fragment patterns (lane = 4 lr + lc: address lc*ld + lr or lr*ld + lc) are bank-conflict free:
// leading dimensions = 4 mod 16 doubles.
// Curved elements, diffusion fields, nDOF > 1 and the Newton-linearised operators stay with hfx_generic.cuh.
#pragma once
#include "hfx_assemble.cuh"

namespace hfx {


// Element descriptions: sizes (Cubature.cpp nIP map), the vertices that span the affine frame of a straight-sided cell (origin + one per
// reference axis) and of a face.  Simplices: 0,1,2,3; orthotopes: 0,1,3,4 and 0,1,3 (vertex order of ReferenceElement.cpp:885-943).
template <int DIM_, int P>
struct BigSimplex {
  using E = ElemCfg<DIM_, P>;
  static constexpr int DIM = DIM_, nN = E::nN, nNf = E::nNf, nFc = E::nFc, nIP = E::nIP, nIPf = E::nIPf;
  __host__ __device__ static constexpr int fv(int i) { return i; }
  __host__ __device__ static constexpr int ff(int i) { return i; }
};
struct BigHexP2 {   // 27-node hexahedron (the reference element's maximum order for orthotopes in 3-D)
  static constexpr int DIM = 3, nN = 27, nNf = 9, nFc = 6, nIP = 58, nIPf = 20;
  __host__ __device__ static constexpr int fv(int i) { return i == 0 ? 0 : (i == 1 ? 1 : (i == 2 ? 3 : 4)); }
  __host__ __device__ static constexpr int ff(int i) { return i == 0 ? 0 : (i == 1 ? 1 : 3); }
};
__host__ __device__ constexpr int big_ld(int x) {   // smallest even leading dimension >= x that is 4 or 12 mod 16 (both DMMA fragment patterns conflict free)
  int v = (x + 3) & ~3;
  while (v % 16 != 4 && v % 16 != 12) v += 4;
  return v;
}

template <class C>
struct BigSmem {
  static constexpr int DIM = C::DIM;
  static constexpr int nN = C::nN, t = C::nNf, nFc = C::nFc, nIP = C::nIP, nIPf = C::nIPf, l = nFc * t;
  static constexpr int nNp = ((nN + 3) / 4) * 4;           // reduction pad (zeros) up to a multiple of 4
  static constexpr int npe = ev(nN);                       // size of the Gauss-Jordan (even), and the leading dimension of the global reference tables
  static_assert(npe == nNp, "the Gauss-Jordan pad and the reduction pad coincide");
  static constexpr int KSN = nNp / 4;                      // reduction steps over the element nodes
  static constexpr int MTN = (nN + 7) / 8;                 // 8-row tiles over the element nodes
  static constexpr int ldc = ((l + 2 + 11) / 16) * 16 + 4; // R, U, Q, Zq rows: >= l + 2, = 4 mod 16
  static constexpr int ldb = big_ld(l);                    // resident B^ rows
  static constexpr int tq = ((t + 3) / 4) * 4;             // face reduction pad
  static constexpr int ldf = big_ld(tq);                   // face matrices, column-major [b][a]
  static constexpr int FSZ = ldf * tq;
  static constexpr int nIPp = ((nIP + 3) / 4) * 4, nIPfp = ((nIPf + 3) / 4) * 4;
  static constexpr int NWT = (2 * nFc + 7) / 8;            // column tiles of the face weights [ip][(f, kind)]
  static constexpr int ldw = big_ld(8 * NWT);
  // resident tables
  static constexpr int oAR = 0;                            // A^_r row-major [DIM][nNp][nNp]
  static constexpr int oBH = oAR + DIM * nNp * nNp;        // B^ row-major [nNp][ldb], column (f, b)
  static constexpr int oMF = oBH + nNp * ldb;              // M^f [b][a] ld ldf
  // per element, small
  static constexpr int oX = oMF + FSZ;                     // [nN][DIM]
  static constexpr int oTAU = oX + ev(nN * DIM);           // [l]
  static constexpr int oVN = oTAU + ev(l);                 // [nN][DIM]
  static constexpr int oGEO = oVN + ev(nN * DIM);          // see the GEO_* offsets in the kernel
  static constexpr int szGEO = 2 * DIM * DIM + 2 + 6 + nFc * (4 * DIM + 4);
  static constexpr int oFU = oGEO + ev(szGEO);             // [nNp]
  static constexpr int oLW = oFU + nNp;                    // [2][nIPp]: Suu weight, Fu weight
  static constexpr int oVT = oLW + 2 * nIPp;               // [nIPp][4]: ts dV (Jinv^T v) at the bulk points
  static constexpr int oFWT = oVT + 4 * nIPp;              // [nIPfp][ldw]
  static constexpr int oSOLD = oFWT + nIPfp * ldw;         // [nNp] old solution (Euler)
  static constexpr int oFT = oSOLD + nNp;                  // FT_f [nFc][FSZ] tau-weighted face mass
  static constexpr int oFC = oFT + nFc * FSZ;              // FC_f [nFc][FSZ] (v.n)-weighted face mass
  // per element, large; the span [oSJ, oUU) is reused by the S phase (Zq + staging)
  static constexpr int oSJ = oFC + nFc * FSZ;              // SJ_r [DIM][nNp k'][nNp m]
  static constexpr int oKA = oSJ + DIM * nNp * nNp;        // K, column-major ld nNp (Gauss-Jordan ping)
  static constexpr int oKB = oKA + nNp * nNp;              // (pong)
  static constexpr int oRR = oKB + nNp * nNp;              // R row-major [nNp][ldc]
  static constexpr int oUU = oRR + nNp * ldc;              // U row-major [nNp][ldc]
  static constexpr int oQQ = oUU + nNp * ldc;              // Q_d row-major [DIM][nN][ldc]; before the condensation: PHI [nIPp][nNp], CG [nIPp][nNp]; K copy
  static constexpr int oEnd = oQQ + DIM * nN * ldc;
  static constexpr int oPHI = oQQ, oCG = oPHI + nIPp * nNp, oKC = oCG + nIPp * nNp;
  static_assert(oKC + nNp * nNp <= oEnd, "PHI, CG and the K copy live in the Q region");
  static constexpr int oZQ = oSJ;                          // Zq_f [nFc][tq][ldc]
  static constexpr int szZQ = nFc * tq * ldc, szST = ev(l * l + l);
  // S staging (nFc^2 blocks of t x t in face-node positions, then S0 [l]): behind Zq inside the span when it fits (order 4), else in its own area
  static constexpr bool stInSpan = oZQ + szZQ + szST <= oUU;
  static_assert(oZQ + szZQ <= oUU, "Zq reuses the SJ / K / R span");
  static constexpr int oST = stInSpan ? oZQ + szZQ : oEnd;
  static constexpr int nDoubles = stInSpan ? oEnd : oEnd + szST;
  static constexpr int nInts = 2 * nFc * t + nFc * nN + nFc * nFc + 6 * nFc + (l + 2) + 8;
  static constexpr size_t bytes = (size_t)nDoubles * 8 + 8 * (size_t)nFc + 4 * (size_t)nInts + 16;
};

template <class C, int NT_>
__global__ void __launch_bounds__(NT_, NT_ >= 512 ? 1 : 2) hdg_big_kernel(const AsmParams p) {
  using L = BigSmem<C>;
  constexpr int DIM = C::DIM;
  constexpr int nN = L::nN, t = L::t, nFc = L::nFc, nIP = L::nIP, nIPf = L::nIPf, l = L::l;
  constexpr int nNp = L::nNp, npe = L::npe, KSN = L::KSN, MTN = L::MTN, ldc = L::ldc, ldb = L::ldb, tq = L::tq, ldf = L::ldf, FSZ = L::FSZ;
  constexpr int nIPp = L::nIPp, nIPfp = L::nIPfp, ldw = L::ldw, D2 = DIM * DIM;
  constexpr int NT = NT_, NWARP = NT / 32;
  constexpr int L1T = (l + 1 + 7) / 8, LT = (l + 7) / 8;
  constexpr int kIPBase = ((nFc * nIPf + 31) / 32) * 32;   // threads [0, nFc nIPf): face cubature points; [kIPBase, kIPBase + nIP): bulk cubature points
  static_assert(l + 1 <= 128 && 160 + nFc * nFc <= NT && kIPBase + nIP <= NT && 128 + nN + DIM * nN <= NT && nFc <= 32, "thread roles");
  static_assert(L::bytes <= 232448, "shared memory of one CTA");
  extern __shared__ __align__(16) double sm[];
  double* const AR = sm + L::oAR; double* const BH = sm + L::oBH; double* const MF = sm + L::oMF;
  double* const X = sm + L::oX; double* const TAU = sm + L::oTAU; double* const VN = sm + L::oVN; double* const GEO = sm + L::oGEO;
  double* const FU = sm + L::oFU; double* const LW = sm + L::oLW; double* const VT = sm + L::oVT; double* const FWT = sm + L::oFWT;
  double* const SOLD = sm + L::oSOLD; double* const FT = sm + L::oFT; double* const FC = sm + L::oFC;
  double* const SJ = sm + L::oSJ; double* const KA = sm + L::oKA; double* const KB = sm + L::oKB; double* const RR = sm + L::oRR;
  double* const UU = sm + L::oUU; double* const QQ = sm + L::oQQ; double* const PHI = sm + L::oPHI; double* const CG = sm + L::oCG;
  double* const KC = sm + L::oKC; double* const ZQ = sm + L::oZQ; double* const ST = sm + L::oST; double* const ST0 = ST + l * l;
  long long* const ROWS = reinterpret_cast<long long*>(sm + L::nDoubles);   // [nFc] first entry of block row F in vals
  int* const ISM = reinterpret_cast<int*>(ROWS + nFc);                      // [nFc] global face ids
  int* const FN = ISM + nFc;                                                // [nFc*t]
  int* const PERM = FN + nFc * t;                                           // [l] element-local face node -> position in faces[F]
  int* const NIF = PERM + nFc * t;                                          // [nFc*nN]
  int* const POS = NIF + nFc * nN;                                          // [nFc*nFc]
  int* const BCF = POS + nFc * nFc; int* const INTF = BCF + nFc; int* const OPP = INTF + nFc;
  int* const CMAP = OPP + 3 * nFc;                                          // [l+1] position -> element-local trace index
  // GEO layout
  constexpr int G_I = 0, G_J = D2, G_DET = 2 * D2, G_RDET = G_DET + 1, G_GG = G_DET + 2, G_F = G_GG + 6;
  constexpr int GF = 4 * DIM + 4;   // per face: n[DIM], h[DIM] (SJ face coefficient), cR[DIM] (R epilogue), cQ[DIM] (Q epilogue), area, tau*area, -, -
  constexpr int GF_N = 0, GF_H = DIM, GF_CR = 2 * DIM, GF_CQ = 3 * DIM, GF_AREA = 4 * DIM, GF_TA = 4 * DIM + 1;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, lr = lane >> 2, lc = lane & 3;
  const bool hasDiff = p.opmask & 1, hasConv = p.opmask & 2, hasReac = (p.opmask & 4) && p.reacIP, hasSrc = (p.opmask & 8) && p.srcIP;
  const bool euler = p.timeScheme == 1;
  const double ts = euler ? p.dt : 1.0;                 // Euler::apply scales the u rows by dt before adding the mass terms (Euler.cpp:28-32)
  const double dsc = hasDiff ? p.diffConst : 0.0;       // D = c I
  const bool needSuu = hasConv || hasReac || euler;     // bulk part of Suu: -C^T, reaction mass, Euler mass
  const bool needPhi = needSuu || hasSrc;

  // ---- once per CTA: zero (pads must hold zeros / finite numbers), resident reference tables ---------------------------------------
  for (int i = tid; i < L::nDoubles; i += NT) sm[i] = 0.0;
  __syncthreads();
  for (int idx = tid; idx < DIM * nN * nN; idx += NT) {   // aref: [r][j][m] (column-major A^_r, ld npe) -> AR[r][m][j]
    const int r = idx / (nN * nN), rem = idx - r * nN * nN, j = rem / nN, m = rem - j * nN;
    AR[(r * nNp + m) * nNp + j] = p.aref[((size_t)r * nN + j) * npe + m];
  }
  for (int idx = tid; idx < nFc * nN * t; idx += NT) {    // bref [f][m][b]
    const int f = idx / (nN * t), rem = idx - f * nN * t, m = rem / t, b = rem - m * t;
    BH[m * ldb + f * t + b] = p.bref[idx];
  }
  for (int idx = tid; idx < t * t; idx += NT) { const int b = idx / t, a = idx - b * t; MF[a + ldf * b] = p.mfref[a + ev(t) * b]; }
  for (int i = tid; i < nFc * t; i += NT) FN[i] = p.faceNodes[i];
  for (int i = tid; i < nFc * nN; i += NT) NIF[i] = p.nodeInFace[i];
  if (tid < nFc) { int vn = 0; for (int kk = 0; kk < nN; kk++) if (p.nodeInFace[tid * nN + kk] < 0) { vn = kk; break; } OPP[tid] = vn; }
  if ((nN & 1) && tid == 0) KA[nN + nNp * nN] = 1.0;      // odd size: unit pad diagonal of K for the 2x2-block Gauss-Jordan (pad row / column stay zero)
  __syncthreads();

  // ---- software prefetch of the next element's gather (registers): the two dependent levels (face ids -> tau / row starts) are issued in two stages
  //      behind the S phase of the current element, so their DRAM latency never sits on the critical path -----------------------------------------------
  static_assert(nN * DIM <= NT, "one thread per coordinate");
  double pfX = 0.0, pfTau = 0.0;
  int pfF = 0, pfPerm = 0, pfSide = 0, pfPos = 0, pfBc = 0, pfInt = 0;
  long long pfRow = 0;
  auto prefetchA = [&](int e) {
    if (e >= p.eEnd) return;
    if (tid < nN * DIM) pfX = p.elemX[(size_t)e * nN * DIM + tid];
    if (tid < l) {
      const int f = tid / t;
      pfF = p.cell2face[(size_t)e * nFc + f];
      pfPerm = p.fperm[(size_t)e * l + tid];
      pfSide = (p.tauVals == 2) ? p.tauSide[(size_t)e * nFc + f] : 0;
    } else if (tid >= 128 && tid < 128 + nFc) pfF = p.cell2face[(size_t)e * nFc + (tid - 128)];
    else if (tid >= 160 && tid < 160 + nFc * nFc) pfPos = p.elemPos[(size_t)e * nFc * nFc + (tid - 160)];
  };
  auto prefetchB = [&](int e) {
    if (e >= p.eEnd) return;
    if (tid < l) pfTau = p.tau[((size_t)pfF * t + pfPerm) * p.tauVals + pfSide];
    else if (tid >= 128 && tid < 128 + nFc) { pfRow = p.faceRowStart[pfF]; pfBc = p.faceBC[pfF]; pfInt = p.faceInterior[pfF]; }
  };
  prefetchA(p.eBegin + blockIdx.x);
  prefetchB(p.eBegin + blockIdx.x);

  long long tprev = clock64();
  for (int e = p.eBegin + blockIdx.x; e < p.eEnd; e += gridDim.x) {
    // ---- P0: commit the prefetched gather (HDGSolver.cpp:231-326) ----------------------------------------------------------------------------
    if (tid < nN * DIM) X[tid] = pfX;
    if (tid < l) {
      PERM[tid] = pfPerm; CMAP[(tid / t) * t + pfPerm] = tid;
      TAU[tid] = pfTau;
    } else if (tid == l) CMAP[l] = l;
    else if (tid >= 128 && tid < 128 + nFc) {
      const int f = tid - 128;
      ISM[f] = pfF; ROWS[f] = pfRow; BCF[f] = pfBc; INTF[f] = pfInt;
    } else if (tid >= 160 && tid < 160 + nFc * nFc) POS[tid - 160] = pfPos;
    if (hasConv) { const int* cell = p.cells + (size_t)e * nN; for (int i = tid; i < nN * DIM; i += NT) VN[i] = p.vel[(size_t)cell[i / DIM] * DIM + (i % DIM)]; }
    if (needPhi) for (int i = tid; i < nIP * nN; i += NT) PHI[(i / nN) * nNp + (i % nN)] = p.shape[i];
    if (euler) for (int i = tid; i < nN; i += NT) SOLD[i] = p.solOld[(size_t)e * nN + i];
    __syncthreads();
    HFX_PROF(0);

    // ---- PG: constant geometry (Operator.cpp:14-84 for an affine map), outward normals (HDGBase.cpp:43-62), is tau constant on each face? ----
    int bad = 0;
    if (tid < nFc) {
      const int f = tid;
      double J[DIM][DIM], det, I[DIM][DIM];
#pragma unroll
      for (int r = 0; r < DIM; r++)
#pragma unroll
        for (int m = 0; m < DIM; m++) J[r][m] = 0.5 * (X[C::fv(r + 1) * DIM + m] - X[C::fv(0) * DIM + m]);
      det_inv(J, det, I);
      const double rdet = fast_rcp(det);
      const int* fn = FN + f * t;
      double Jf[DIM - 1][DIM];
#pragma unroll
      for (int r = 0; r < DIM - 1; r++)
#pragma unroll
        for (int m = 0; m < DIM; m++) Jf[r][m] = 0.5 * (X[fn[C::ff(r + 1)] * DIM + m] - X[fn[C::ff(0)] * DIM + m]);
      double nv[DIM];
      if (DIM == 2) { nv[0] = -Jf[0][1]; nv[1] = Jf[0][0]; }
      else {
        nv[0] = Jf[0][1] * Jf[DIM - 2][2 % DIM] - Jf[0][2 % DIM] * Jf[DIM - 2][1];
        nv[1] = Jf[0][2 % DIM] * Jf[DIM - 2][0] - Jf[0][0] * Jf[DIM - 2][2 % DIM];
        nv[DIM - 1] = Jf[0][0] * Jf[DIM - 2][1] - Jf[0][1] * Jf[DIM - 2][0];
      }
      double nn = 0.0;
#pragma unroll
      for (int m = 0; m < DIM; m++) nn = fma(nv[m], nv[m], nn);
      const double inrm = fast_rsqrt(nn), area = nn * inrm;   // |J_0 x J_1| = sqrt(det(J J^T)) (Operator.cpp:66-69)
      double prod = 0.0;
#pragma unroll
      for (int m = 0; m < DIM; m++) prod = fma(X[OPP[f] * DIM + m] - X[fn[0] * DIM + m], nv[m], prod);
      const double sg = (prod > 0.0 ? -1.0 : 1.0) * inrm;
      double* g = GEO + G_F + f * GF;
#pragma unroll
      for (int d = 0; d < DIM; d++) nv[d] *= sg;              // unit outward normal
#pragma unroll
      for (int d = 0; d < DIM; d++) { g[GF_N + d] = nv[d]; g[GF_CQ + d] = area * nv[d] * rdet; }
#pragma unroll
      for (int r = 0; r < DIM; r++) {
        double h = 0.0, nu = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; d++) { h = fma(I[d][r], -area * nv[d], h); nu = fma(J[r][d], nv[d], nu); }
        g[GF_H + r] = h; g[GF_CR + r] = area * rdet * nu;
      }
      g[GF_AREA] = area; g[GF_TA] = TAU[f * t] * area;
    } else if (tid == 32) {
      double J[DIM][DIM], det, I[DIM][DIM];
#pragma unroll
      for (int r = 0; r < DIM; r++)
#pragma unroll
        for (int m = 0; m < DIM; m++) J[r][m] = 0.5 * (X[C::fv(r + 1) * DIM + m] - X[C::fv(0) * DIM + m]);
      det_inv(J, det, I);
#pragma unroll
      for (int m = 0; m < DIM; m++)
#pragma unroll
        for (int r = 0; r < DIM; r++) { GEO[G_I + m * DIM + r] = I[m][r]; GEO[G_J + m * DIM + r] = J[m][r]; }
      GEO[G_DET] = det; GEO[G_RDET] = fast_rcp(det);
      int o = 0;
#pragma unroll
      for (int r = 0; r < DIM; r++)
#pragma unroll
        for (int r2 = r; r2 < DIM; r2++) {
          double s = 0.0;
#pragma unroll
          for (int d = 0; d < DIM; d++) s = fma(I[d][r], I[d][r2], s);
          GEO[G_GG + o++] = s;
        }
    } else if (tid >= 64) {
      for (int i = tid - 64; i < l; i += NT - 64) bad |= (TAU[i] != TAU[(i / t) * t]);
    }
    const bool tauConst = !__syncthreads_or(bad);
    const bool faceContr = !tauConst || hasConv;   // tau or v.n vary along a face: weighted face masses by cubature (HDGBase.cpp:112-131, HDGConvection.cpp:60-104)
    const double det = GEO[G_DET];
    HFX_PROF(1);

    // ---- PA: SJ_r = ts c (detJ sum_r' G(r,r') S^_r'^T + sum_f h_fr E_f) ; face weights / tau masses ; bulk point weights ----------------
    if (faceContr) {
      if (tid < nFc * nIPf) {   // face cubature weights: dV tau, dV v.n
        const int f = tid / nIPf, ip = tid - f * nIPf;
        const int* fn = FN + f * t;
        const double* g = GEO + G_F + f * GF;
        double tauip = 0.0, v[DIM];
#pragma unroll
        for (int d = 0; d < DIM; d++) v[d] = 0.0;
        for (int a = 0; a < t; a++) {
          const double s = __ldg(p.fshape + ip * t + a);
          tauip = fma(s, TAU[f * t + a], tauip);
          if (hasConv) {
#pragma unroll
            for (int d = 0; d < DIM; d++) v[d] = fma(s, VN[fn[a] * DIM + d], v[d]);
          }
        }
        double vdn = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; d++) vdn = fma(v[d], g[GF_N + d], vdn);
        const double dvf = __ldg(p.fw + ip) * g[GF_AREA];
        FWT[ip * ldw + 2 * f] = dvf * tauip;
        FWT[ip * ldw + 2 * f + 1] = hasConv ? dvf * vdn : 0.0;
      }
    } else {
      for (int idx = tid; idx < nFc * t * t; idx += NT) {
        const int f = idx / (t * t), ab = idx - f * t * t, b = ab / t, a = ab - b * t;
        FT[f * FSZ + a + ldf * b] = GEO[G_F + f * GF + GF_TA] * MF[a + ldf * b];
      }
    }
    if (needPhi && tid >= kIPBase && tid < kIPBase + nIP) {   // bulk cubature points: weights of the Suu / Fu contractions, convective velocity
      const int ip = tid - kIPBase;
      const double dv = __ldg(p.w + ip) * det;
      double lw = 0.0;
      if (hasReac) lw += ts * p.reacIP[(size_t)e * nIP + ip] * dv;
      if (euler) lw += dv;
      LW[ip] = lw;
      double rw = hasSrc ? ts * p.srcIP[(size_t)e * nIP + ip] * dv : 0.0;
      if (euler) { double uo = 0.0; for (int i = 0; i < nN; i++) uo = fma(PHI[ip * nNp + i], SOLD[i], uo); rw = fma(dv, uo, rw); }
      LW[nIPp + ip] = rw;
      if (hasConv) {
        double v[DIM];
#pragma unroll
        for (int d = 0; d < DIM; d++) v[d] = 0.0;
        for (int i = 0; i < nN; i++) {
          const double s = PHI[ip * nNp + i];
#pragma unroll
          for (int d = 0; d < DIM; d++) v[d] = fma(s, VN[i * DIM + d], v[d]);
        }
#pragma unroll
        for (int r = 0; r < DIM; r++) {
          double s = 0.0;
#pragma unroll
          for (int d = 0; d < DIM; d++) s = fma(v[d], GEO[G_I + d * DIM + r], s);
          VT[ip * 4 + r] = ts * dv * s;
        }
      }
    }
    {
      double GG[DIM][DIM], hf[nFc][DIM];
      { int o = 0;
#pragma unroll
        for (int r = 0; r < DIM; r++)
#pragma unroll
          for (int r2 = r; r2 < DIM; r2++) { GG[r][r2] = GEO[G_GG + o]; GG[r2][r] = GEO[G_GG + o]; o++; } }
#pragma unroll
      for (int f = 0; f < nFc; f++)
#pragma unroll
        for (int r = 0; r < DIM; r++) hf[f][r] = GEO[G_F + f * GF + GF_H + r];
      const double sc = ts * dsc;
      for (int idx = tid; idx < nN * nN; idx += NT) {
        const int kq = idx / nN, m = idx - kq * nN;
        double s[DIM], ef[nFc];
#pragma unroll
        for (int r = 0; r < DIM; r++) s[r] = __ldg(p.srefT + ((size_t)r * nN + kq) * npe + m);   // S^_r[m][k']
#pragma unroll
        for (int f = 0; f < nFc; f++) {
          const int a = NIF[f * nN + m], b = NIF[f * nN + kq];
          ef[f] = (a >= 0 && b >= 0) ? MF[a + ldf * b] : 0.0;
        }
#pragma unroll
        for (int r = 0; r < DIM; r++) {
          double v = 0.0, w = 0.0;
#pragma unroll
          for (int r2 = 0; r2 < DIM; r2++) v = fma(GG[r][r2], s[r2], v);
#pragma unroll
          for (int f = 0; f < nFc; f++) w = fma(hf[f][r], ef[f], w);
          SJ[(r * nNp + kq) * nNp + m] = sc * fma(det, v, w);
        }
      }
    }
    __syncthreads();
    HFX_PROF(2);

    // ---- PB: weighted face masses FT_f, FC_f[a][b] = sum_ip wt[ip][(f,kind)] phi_a phi_b (tensor cores; left operand from L2) ; CG = Suu left operand ----
    if (faceContr || needSuu) {
      if (faceContr) {
        constexpr int MR = t * t, FW_MT = (MR + 7) / 8, KSF = nIPfp / 4;
        for (int task = warp; task < FW_MT; task += NWARP) {
          const int m = task * 8 + lr, mc = imin(m, MR - 1);
          double a[KSF], c[L::NWT][2];
          zero_c(c);
#pragma unroll
          for (int ks = 0; ks < KSF; ks++) { const int k = ks * 4 + lc; a[ks] = k < nIPf ? __ldg(p.ffs + (size_t)k * MR + mc) : 0.0; }
#pragma unroll
          for (int ks = 0; ks < KSF; ks++)
#pragma unroll
            for (int j = 0; j < L::NWT; j++) dmma(c[j], a[ks], FWT[(ks * 4 + lc) * ldw + 8 * j + lr]);
          if (m < MR) {   // column tile j, lane lc holds (f = 4 j + lc, kind 0 | 1); row m = b * t + a
            const int b = m / t, a2 = m - b * t;
#pragma unroll
            for (int j = 0; j < L::NWT; j++) { const int f = 4 * j + lc; if (f < nFc) { FT[f * FSZ + a2 + ldf * b] = c[j][0]; FC[f * FSZ + a2 + ldf * b] = c[j][1]; } }
          }
        }
      }
      if (needSuu) for (int idx = tid; idx < nIP * nN; idx += NT) {
        const int ip = idx / nN, i = idx - ip * nN;
        double c = LW[ip] * PHI[ip * nNp + i];
        if (hasConv) {
          const double* d = p.dshape + (size_t)idx * DIM;
#pragma unroll
          for (int r = 0; r < DIM; r++) c = fma(-VT[ip * 4 + r], __ldg(d + r), c);   // -dV (v . grad phi_i) (Convection.cpp:5-49, transposed)
        }
        CG[ip * nNp + i] = c;
      }
      if (needSuu) for (int i = tid; i < (nIPp - nIP) * nNp; i += NT) CG[nIP * nNp + i] = 0.0;   // reduction pad rows (the region held Q before)
      __syncthreads();
    }
    HFX_PROF(3);

    // ---- PC: Suu = bulk part (cubature contraction) + tau masses scattered to the element nodes (HDGBase.cpp:128) -> KA ; Fu ----------------
    if (needSuu) {
      constexpr int KSIP = nIPp / 4;
      for (int task = warp; task < MTN * MTN; task += NWARP) {
        const int mt = task % MTN, nt = task / MTN;
        const double* pa = CG + imin(mt * 8 + lr, nN - 1);
        const double* pb = PHI + imin(nt * 8 + lr, nN - 1);
        double c0[2] = {0.0, 0.0}, c1[2] = {0.0, 0.0};
#pragma unroll
        for (int ks = 0; ks < KSIP; ks += 2) {
          const int k = ks * 4 + lc;
          dmma(c0, pa[k * nNp], pb[k * nNp]);
          if (ks + 1 < KSIP) dmma(c1, pa[(k + 4) * nNp], pb[(k + 4) * nNp]);
        }
        const int m = mt * 8 + lr;
        if (m < nN) {
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const int n = nt * 8 + 2 * lc + h;
            if (n < nN) {
              double v = c0[h] + c1[h];
              for (int f = 0; f < nFc; f++) { const int a = NIF[f * nN + m], b = NIF[f * nN + n]; if (a >= 0 && b >= 0) v = fma(ts, FT[f * FSZ + a + ldf * b], v); }
              KA[m + nNp * n] = v;
            }
          }
        }
      }
    } else {
      for (int idx = tid; idx < nN * nN; idx += NT) {
        const int n = idx / nN, m = idx - n * nN;
        double v = 0.0;
#pragma unroll
        for (int f = 0; f < nFc; f++) { const int a = NIF[f * nN + m], b = NIF[f * nN + n]; if (a >= 0 && b >= 0) v += FT[f * FSZ + a + ldf * b]; }
        KA[m + nNp * n] = ts * v;
      }
    }
    if (tid >= NT - 64 && tid < NT - 64 + nN) {   // Fu = source (Source.cpp:24-48) + Euler mass * old solution (Euler.cpp:29-30)
      const int i = tid - (NT - 64);
      double s2 = 0.0;
      if (hasSrc || euler) for (int ip = 0; ip < nIP; ip++) s2 = fma(PHI[ip * nNp + i], LW[nIPp + ip], s2);
      FU[i] = s2;
    }
    __syncthreads();
    HFX_PROF(4);

    // ---- PD: K = Suu - sum_r SJ_r A^_r (also copied for the refinement) ; R_f = Sul_f + sum_r cR_fr (SJ_r B^_f) ; column l of R = -Fu ---------
    {
      constexpr int T_K = MTN * MTN, T_R = MTN * LT;
      for (int task = warp; task < T_K + T_R; task += NWARP) {
        const bool isK = task < T_K;
        const int tk = isK ? task : task - T_K;
        const int mt = tk % MTN, nt = tk / MTN;
        const int m = mt * 8 + lr, mc = imin(m, nN - 1);
        const int ncB = imin(nt * 8 + lr, (isK ? nN : l) - 1);
        const double* pa = SJ + mc;                                   // SJ_r[m][k'] at (r nNp + k') nNp + m
        const double* pb = isK ? AR + ncB : BH + ncB;                 // K: A^_r[k'][n] ; R: B^[k'][n]
        const int sbk = isK ? nNp : ldb, sbr = isK ? nNp * nNp : 0;
        double c[DIM][2];
        zero_c(c);
#pragma unroll
        for (int ks = 0; ks < KSN; ks++) {
          const int k = ks * 4 + lc;
          double a[DIM], b[DIM];
#pragma unroll
          for (int r = 0; r < DIM; r++) { a[r] = pa[(r * nNp + k) * nNp]; b[r] = (isK || r == 0) ? pb[r * sbr + k * sbk] : b[0]; }
#pragma unroll
          for (int r = 0; r < DIM; r++) dmma(c[r], a[r], b[r]);
        }
        if (m < nN) {
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const int cc = nt * 8 + 2 * lc + h;
            if (isK) {
              if (cc < nN) {
                double v = c[0][h];
#pragma unroll
                for (int r = 1; r < DIM; r++) v += c[r][h];
                const double kv = KA[m + nNp * cc] - v;
                KA[m + nNp * cc] = kv; KC[m + nNp * cc] = kv;
              }
            } else if (cc < l) {
              const int f = cc / t, b2 = cc - f * t, a2 = NIF[f * nN + m];
              const double* g = GEO + G_F + f * GF + GF_CR;
              double v = 0.0;
#pragma unroll
              for (int r = 0; r < DIM; r++) v = fma(g[r], c[r][h], v);
              if (a2 >= 0) v += ts * ((hasConv ? FC[f * FSZ + a2 + ldf * b2] : 0.0) - FT[f * FSZ + a2 + ldf * b2]);   // Sul = -tau mass + (v.n) mass
              RR[m * ldc + cc] = v;
            }
          }
        }
      }
      if (tid < nN) { RR[tid * ldc + l] = -FU[tid]; RR[tid * ldc + l + 1] = 0.0; }
    }
    __syncthreads();
    HFX_PROF(5);

    // ---- PE: K^-1 (2x2-block-pivot Gauss-Jordan, all warps; the inverse ends in KA or KB) ------------------------------------------------------
    // (the pivot chain is serial and every thread redoes the pivot-block determinant / reciprocal: fewer threads = less redundant FP64 issue)
    double* KI;
    if (p.gjThreads == 512) { group_invert<npe, nNp, NT, true>(KA, KB, tid, p.status, 1); KI = ((npe / 2) & 1) ? KB : KA; }   // (2x2-block pivots on CUDA cores: HFX_BIG_GJ=512)
    else { block4_invert<npe, nNp>(KA, KB, tid, p.status, 1); __syncthreads(); KI = ((npe / 4) & 1) ? KB : KA; }               // 4x4-block pivots, tensor-core updates
    HFX_PROF(6);

    // ---- PF: U = -K^-1 R, one refinement step U -= K^-1 (K U + R) (see hfx_assemble.cuh P7); a warp owns (row tile, column tile) -------------------
    {
      constexpr int T_U = MTN * L1T;
      // Cm[tile] = beta Cm[tile] + sgn Am Bm[:, tile] ; Am column-major ld nNp, Bm / Cm row-major ld ldc
      auto u_pass = [&](const double* Am, const double* Bm, double* Cm, double sgn, bool acc) {
        for (int task = warp; task < T_U; task += NWARP) {
          const int mt = task % MTN, nt = task / MTN;
          const int m = mt * 8 + lr;
          const double* pa = Am + imin(m, nN - 1);
          const double* pb = Bm + imin(nt * 8 + lr, l);
          double c0[2] = {0.0, 0.0}, c1[2] = {0.0, 0.0};
#pragma unroll
          for (int ks = 0; ks < KSN; ks += 2) {
            const int k = ks * 4 + lc;
            dmma(c0, pa[k * nNp], pb[k * ldc]);
            if (ks + 1 < KSN) dmma(c1, pa[(k + 4) * nNp], pb[(k + 4) * ldc]);
          }
          const int n = nt * 8 + 2 * lc;
          if (m < nN && n <= l) {
            double2* d2 = reinterpret_cast<double2*>(Cm + m * ldc + n);
            const double2 o = acc ? *d2 : make_double2(0.0, 0.0);
            *d2 = make_double2(fma(sgn, c0[0] + c1[0], o.x), fma(sgn, c0[1] + c1[1], o.y));
          }
        }
      };
      u_pass(KI, RR, UU, -1.0, false);
      __syncthreads();
      u_pass(KC, UU, RR, 1.0, true);      // V = R + K U (in place over R)
      __syncthreads();
      u_pass(KI, RR, UU, -1.0, true);     // U -= K^-1 V
      __syncthreads();
    }
    HFX_PROF(7);

    // ---- PQ: Q_d = -sum_r Jinv(d,r) (A^_r U) + cQ_fd B^_f ; Q0_d (column l) -------------------------------------------------------------------
    {
      double Ii[DIM][DIM];
#pragma unroll
      for (int d = 0; d < DIM; d++)
#pragma unroll
        for (int r = 0; r < DIM; r++) Ii[d][r] = GEO[G_I + d * DIM + r];
      constexpr int T_Q = MTN * L1T;
      for (int task = warp; task < T_Q; task += NWARP) {
        const int mt = task % MTN, nt = task / MTN;
        const int m = mt * 8 + lr, mc = imin(m, nN - 1);
        const double* pa = AR + mc * nNp;                  // A^_r[m][k'] at (r nNp + m) nNp + k'
        const double* pb = UU + imin(nt * 8 + lr, l);
        double c[DIM][2];
        zero_c(c);
#pragma unroll
        for (int ks = 0; ks < KSN; ks++) {
          const int k = ks * 4 + lc;
          const double b = pb[k * ldc];
          double a[DIM];
#pragma unroll
          for (int r = 0; r < DIM; r++) a[r] = pa[r * nNp * nNp + k];
#pragma unroll
          for (int r = 0; r < DIM; r++) dmma(c[r], a[r], b);
        }
        const int n = nt * 8 + 2 * lc;
        if (m < nN && n <= l) {
          double bh[2] = {0.0, 0.0}; int fc[2] = {0, 0};
#pragma unroll
          for (int h = 0; h < 2; h++) if (n + h < l) { bh[h] = BH[m * ldb + n + h]; fc[h] = (n + h) / t; }
          double qv[DIM][2];
#pragma unroll
          for (int d = 0; d < DIM; d++) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
              double s = 0.0;
#pragma unroll
              for (int r = 0; r < DIM; r++) s = fma(-Ii[d][r], c[r][h], s);
              qv[d][h] = fma(GEO[G_F + fc[h] * GF + GF_CQ + d], bh[h], s);
            }
            *reinterpret_cast<double2*>(QQ + (d * nN + m) * ldc + n) = make_double2(qv[d][0], qv[d][1]);
          }
          // Zq_f[b][:] = -c sum_d n_fd Q_d[faceNodes_f(b)][:] for the faces this node lies on (straight face, D = c I: Slq_d = -c area n_fd M^f): the S phase reads
          // it instead of the (1 + dim) t long reduction.  The span that held SJ / K / R is free since the refinement of U.
          if (!p.recover) {
#pragma unroll
            for (int f = 0; f < nFc; f++) {
              const int b = NIF[f * nN + m];
              if (b >= 0) {
                const double* g = GEO + G_F + f * GF + GF_N;
                double z0 = 0.0, z1 = 0.0;
#pragma unroll
                for (int d = 0; d < DIM; d++) { const double w = -dsc * g[d]; z0 = fma(w, qv[d][0], z0); z1 = fma(w, qv[d][1], z1); }
                *reinterpret_cast<double2*>(ZQ + (f * tq + b) * ldc + n) = make_double2(z0, z1);
              }
            }
          }
        }
      }
    }
    fence_proxy_async();
    __syncthreads();
    HFX_PROF(8);
    prefetchA(e + gridDim.x);
    if (p.recover) {   // recovery by recomputation: u_e = U lambda_e + U0, q_e = Q lambda_e + Q0 straight out of shared memory; nothing else leaves
      constexpr int q = DIM * nN;
      double* const LAM = SJ;   // (dead since PD)
      if (tid < l) LAM[tid] = p.recTrace[(size_t)ISM[tid / t] * t + PERM[tid]];
      __syncthreads();
      if (tid < nN + q) {
        const int rq = tid - nN;
        const double* row = tid < nN ? UU + tid * ldc : QQ + ((rq % DIM) * nN + rq / DIM) * ldc;
        double s0 = row[l], s1 = 0.0;
#pragma unroll 4
        for (int c2 = 0; c2 < l; c2 += 2) { s0 = fma(row[c2], LAM[c2], s0); s1 = fma(row[c2 + 1], LAM[c2 + 1], s1); }
        if (tid < nN) p.recSol[(size_t)e * nN + tid] = s0 + s1; else p.recFlux[(size_t)e * q + rq] = s0 + s1;
      }
      prefetchB(e + gridDim.x);
      __syncthreads();
      continue;
    }
    // U, Q leave as whole rows (row-major per element in HBM): one bulk copy (TMA) per row, in flight during the S phase
    if (p.U) {
      constexpr int q = DIM * nN;
      if (p.gjThreads != 2) {   // one TMA row copy per thread: the issuing warps spend ~3 k cycles at order 4 (140 serialised issues) while the others start the S phase
        if (tid < nN + q) {
          const int row = tid;
          if (row < nN) bulk_store(p.U + ((size_t)e * nN + row) * l, UU + row * ldc, l * 8);
          else { const int rq = row - nN; bulk_store(p.Q + ((size_t)e * q + rq) * l, QQ + ((rq % DIM) * nN + rq / DIM) * ldc, l * 8); }
          bulk_commit();
        }
      } else {                  // (experiment HFX_GJ=2: coalesced 16-byte stores by the whole CTA: same time, the store path is the limit either way)
        static_assert(l % 2 == 0, "16-byte pieces");
        constexpr int L2 = l / 2;
        double2* const gU = reinterpret_cast<double2*>(p.U + (size_t)e * nN * l);
        double2* const gQ = reinterpret_cast<double2*>(p.Q + (size_t)e * q * l);
        for (int idx = tid; idx < (nN + q) * L2; idx += NT) {
          const int row = idx / L2, c2 = idx - row * L2;
          if (row < nN) gU[idx] = *reinterpret_cast<const double2*>(UU + row * ldc + 2 * c2);
          else { const int rq = row - nN; gQ[rq * L2 + c2] = *reinterpret_cast<const double2*>(QQ + ((rq % DIM) * nN + rq / DIM) * ldc + 2 * c2); }
        }
      }
    }

    HFX_PROF(9);

    // ---- PS: S_f = FT_f (U_f - I_f) + (area_f M^f) Zq_f (+ FC_f on the diagonal block: convection part of Sll) ; S0 = -(column l) (:347-348);
    //      Dirichlet rows (:489-501).  Output rows / columns are face-node POSITIONS: the staging holds the nFc^2 blocks of the global block CSR.
    {
      constexpr int TT = (t + 7) / 8, NTW = 2, NG = (L1T + NTW - 1) / NTW, KSF = tq / 4;
      for (int task = warp; task < nFc * NG; task += NWARP) {
        const int f = task / NG, ng = task - f * NG;
        const double* ftf = FT + f * FSZ;
        const double areaf = GEO[G_F + f * GF + GF_AREA];
        int acl[TT], ncl[NTW];
#pragma unroll
        for (int i = 0; i < TT; i++) acl[i] = CMAP[f * t + imin(i * 8 + lr, t - 1)] - f * t;
#pragma unroll
        for (int j = 0; j < NTW; j++) ncl[j] = CMAP[imin((ng * NTW + j) * 8 + lr, l)];
        double c[TT][NTW][2];
#pragma unroll
        for (int i = 0; i < TT; i++) zero_c(c[i]);
#pragma unroll
        for (int ks = 0; ks < 2 * KSF; ks++) {
          const bool first = ks < KSF;
          const int b = (first ? ks : ks - KSF) * 4 + lc;      // reduction index: face node b (b >= t: zero pad column of the face matrices)
          double av[TT], bv[NTW];
          if (first) {
            const double* urow = UU + FN[f * t + imin(b, t - 1)] * ldc;
#pragma unroll
            for (int i = 0; i < TT; i++) av[i] = ftf[acl[i] + ldf * b];
#pragma unroll
            for (int j = 0; j < NTW; j++) bv[j] = urow[ncl[j]] - ((ncl[j] == f * t + b) ? 1.0 : 0.0);   // Sll = -tau mass rides along: tau mass (U - I)
          } else {
#pragma unroll
            for (int i = 0; i < TT; i++) av[i] = areaf * MF[acl[i] + ldf * b];
#pragma unroll
            for (int j = 0; j < NTW; j++) bv[j] = ZQ[(f * tq + b) * ldc + ncl[j]];
          }
#pragma unroll
          for (int i = 0; i < TT; i++)
#pragma unroll
            for (int j = 0; j < NTW; j++) dmma(c[i][j], av[i], bv[j]);
        }
        const int bcf = BCF[f];
#pragma unroll
        for (int i = 0; i < TT; i++) {
          const int a = i * 8 + lr;   // row position
          if (a < t) {
#pragma unroll
            for (int j = 0; j < NTW; j++) {
#pragma unroll
              for (int h = 0; h < 2; h++) {
                const int cp = (ng * NTW + j) * 8 + 2 * lc + h;   // column position
                double v = c[i][j][h];
                if (cp < l) {
                  const int f2 = cp / t, pb = cp - f2 * t;
                  if (hasConv || bcf) {
                    const int bl = CMAP[cp] - f2 * t;
                    if (f2 == f) {
                      if (hasConv) v += FC[f * FSZ + acl[i] + ldf * bl];
                      if (bcf == 2) v = areaf * MF[acl[i] + ldf * bl];                        // IntegratedDirichletModel row: face mass
                    } else if (bcf == 2) v = 0.0;
                    if (bcf == 1) v = (f2 == f && pb == a) ? 1.0 : 0.0;                       // DirichletModel row (Set)
                  }
                  ST[((f * nFc + f2) * t + a) * t + pb] = v;
                } else if (cp == l) ST0[f * t + a] = v;
              }
            }
          }
        }
      }
    }
    __syncthreads();
    HFX_PROF(10);
    prefetchB(e + gridDim.x);

    // ---- PW: write-out.  Block (f, f2) of the element is one contiguous t x t block of the global block CSR (face-node order): coalesced copy,
    //      atomic add where the second element of an interior face adds to the same diagonal block (two contributors on zeroed storage: order independent).
    {
      constexpr int q = DIM * nN;
      for (int k = warp; k < nFc * nFc; k += NWARP) {
        const int f = k / nFc, f2 = k - f * nFc;
        const double* src = ST + k * t * t;
        double* dst = p.vals + ROWS[f] + (long long)POS[k] * t * t;
        if (f2 == f && INTF[f]) { for (int i = lane; i < t * t; i += 32) atomicAdd(dst + i, src[i]); }
        else for (int i = lane; i < t * t; i += 32) dst[i] = src[i];
      }
      if (p.S) {
        double* gS = p.S + (size_t)e * l * l;
        for (int idx = tid; idx < l * l; idx += NT) {
          const int cc = idx / l, r = idx - cc * l, f = r / t, f2 = cc / t;
          gS[idx] = ST[((f * nFc + f2) * t + PERM[r]) * t + PERM[cc]];
        }
      }
      if (tid < l) {
        const int r = tid, f = r / t, a = r - f * t, F = ISM[f], bc = BCF[f];
        double s0 = -ST0[f * t + PERM[r]];
        if (bc == 1) s0 = p.dirichlet[(size_t)F * t + a];
        else if (bc == 2) {
          const double areaf = GEO[G_F + f * GF + GF_AREA];
          s0 = 0.0;
          for (int b = 0; b < t; b++) s0 = fma(areaf * MF[a + ldf * b], p.dirichlet[(size_t)F * t + b], s0);
        }
        if (p.S0) p.S0[(size_t)e * l + r] = s0;
        const int rowDof = F * t + PERM[r];
        if (INTF[f]) atomicAdd(p.rhs + rowDof, s0); else p.rhs[rowDof] = s0;
      } else if (p.U && tid >= 128 && tid < 128 + nN + q) {
        const int row = tid - 128;
        if (row < nN) p.U0[(size_t)e * nN + row] = UU[row * ldc + l];
        else { const int rq = row - nN; p.Q0[(size_t)e * q + rq] = QQ[((rq % DIM) * nN + rq / DIM) * ldc + l]; }
      }
      bulk_wait_read();   // U, Q rows have left shared memory: the regions are rewritten by the next element pass
    }
    __syncthreads();
    // the S-phase span held K: restore the pad row / column of the 2x2-block Gauss-Jordan (zeros, unit diagonal); every other pad of the span only has
    // to be FINITE (its partner operand is an exact zero of the resident tables / of U), and it is
    if ((nN & 1) && tid < nNp) { KA[nN + nNp * tid] = 0.0; KA[tid + nNp * nN] = tid == nN ? 1.0 : 0.0; }
    HFX_PROF(11);
  }
}

template <class C, int NT_>
inline cudaError_t launch_big(const AsmParams& p, int nSM, cudaStream_t st) {
  using L = BigSmem<C>;
  cudaError_t e = cudaFuncSetAttribute(hdg_big_kernel<C, NT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes);
  if (e != cudaSuccess) return e;
  int perSM = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, hdg_big_kernel<C, NT_>, NT_, L::bytes);
  if (perSM < 1) perSM = 1;
  long long grid = (long long)nSM * perSM;
  if (grid > p.eEnd - p.eBegin) grid = p.eEnd - p.eBegin;
  if (grid < 1) grid = 1;
  hdg_big_kernel<C, NT_><<<(int)grid, NT_, L::bytes, st>>>(p);
  return cudaGetLastError();
}

}  // namespace hfx
