// Small straight-sided tetrahedra (3-D order 2: 10 nodes, 6 face nodes, local system 10 + 30 + 24), ONE WARP PER ELEMENT, ONE TRACE COLUMN PER LANE.
// Once K^-1 is known the columns of the condensation are independent: lane c < l owns trace column c = (face, face node) -- its column of R, U, Q and S --
// lane l the right-hand-side column (U0, Q0, S0).  A column is a chain of small matrix-vector products whose matrices (SJ_r, K, K^-1 of the element; A^_r of the
// reference element) are read from shared memory as 16-byte broadcasts, so there is no tile padding (a 10-node element fills 39 % of the 16 x 16 DMMA tiles of the
// element-group kernel), no block-wide barrier, and the rows of U and Q leave as contiguous runs.  The shared part of the element (geometry, tau masses, SJ_r, K and the
// in-warp Gauss-Jordan) is spread over the 32 lanes.  The algebra is the SJ_r formulation of hfx_big.cuh / hfx_p1.cuh:
//   SJ_r = c (detJ sum_r' G(r,r') S^_r' + sum_f h_fr E_f)           K = Suu - sum_r SJ_r A^_r          K^-1: unpivoted Gauss-Jordan, U refined once
//   column c = (f, b):  R = Sul + cR_f . (SJ_r B^_f)                 U = -K^-1 R, U -= K^-1 (K U + R)   Q_d = -sum_r Jinv(d,r) (A^_r U) + cQ_fd B^_f
//                       S[(f',a)][c] = FT_f' (U_f' - I) + area_f' M^f Zq_f',   Zq_f' = -c sum_d n_f'd Q_d[faceNodes_f']  (accumulated while Q is formed)
// Models: Base + Diffusion with D = c I (+ Source), straight-sided cells, any tau, both boundary models -- the eligibility of hdg_p1_kernel.
// Reference semantics: Operator.cpp:14-84, HDGBase.cpp:18-158, HDGDiffusion.cpp:74-145, Source.cpp:24-48, HDGSolver.cpp:231-348 (gather + condensation),
// :361-529 (boundary rows), :531-675 (scatter).
#pragma once
#include "hfx_assemble.cuh"

namespace hfx {

// face-node map of the order-2 tetrahedron and its inverse (node -> position in face or -1) for the kernel: namespace-scope constants, so that indices that are
// compile-time constants after unrolling fold away and per-node register arrays stay in registers
__device__ constexpr int kC2FN[4][6] = {{3, 1, 0, 8, 4, 7}, {2, 1, 3, 5, 8, 9}, {2, 3, 0, 9, 7, 6}, {0, 1, 2, 4, 5, 6}};
__device__ constexpr int kC2NIF[4][10] = {{2, 1, -1, 0, 4, -1, -1, 5, 3, -1}, {-1, 1, 0, 2, -1, 3, -1, -1, 4, 5}, {2, -1, 0, 1, -1, -1, 5, 4, -1, 3}, {0, 1, 2, -1, 3, 4, 5, -1, -1, -1}};
template <int P> struct ColEl;
template <> struct ColEl<2> {
  __device__ static __forceinline__ int dfn(int f, int b) { return kC2FN[f][b]; }
  __device__ static __forceinline__ int dnif(int f, int m) { return kC2NIF[f][m]; }
  static constexpr int nN = 10, t = 6, nIP = 14, nIPf = 6;
  // face-node map of the order-2 tetrahedron (ReferenceElement.cpp:636-878; checked against the host tables before the kernel is used)
  __host__ __device__ static constexpr int fn(int f, int b) {
    constexpr int T[4][6] = {{3, 1, 0, 8, 4, 7}, {2, 1, 3, 5, 8, 9}, {2, 3, 0, 9, 7, 6}, {0, 1, 2, 4, 5, 6}};
    return T[f][b];
  }
};
template <class C> __host__ __device__ constexpr int col_nif(int f, int m) {   // node -> position in face (or -1)
  for (int b = 0; b < C::t; b++) if (C::fn(f, b) == m) return b;
  return -1;
}

template <int P>
struct ColLayout {
  using C = ColEl<P>;
  static constexpr int nN = C::nN, t = C::t, l = 4 * t, nIP = C::nIP;
  // CTA-wide tables (doubles)
  static constexpr int tA = 0, tS = tA + 3 * nN * nN, tMF = tS + 3 * nN * nN, tBH = tMF + ev(t * t), tT3 = tBH + 4 * nN * t, tPHIW = tT3 + ev(t * t * t),
                       tEF = tPHIW + ev(nIP * nN), tNIF = tEF + 4 * nN * nN, tEnd = tNIF + ev(4 * nN) / 2 + 2;   // NIF: 4 nN ints
  // per element (doubles)
  static constexpr int oX = 0, oN = 12, oAR = 24, oHF = 28, oCR = 40, oCQ = 52, oFU = 64, oTAU = oFU + ev(nN), oFT = oTAU + ev(l), oSJ = oFT + 4 * t * t,
                       oK = oSJ + 3 * nN * nN, oKC = oK + ev(nN * nN), oRS = oKC + ev(nN * nN), oINT = oRS + 4;
  static constexpr int iF = 0, iBC = 4, iIN = 8, iSD = 12, iPOS = 16, iPERM = 32, nInts = 32 + l;
  static constexpr int stride = ev(oINT + (nInts + 1) / 2) + 2;    // (+2: consecutive slices start 4 banks apart)
  static constexpr int NQ = (nN * nN + 31) / 32;
  static_assert(l <= 28 && (nN % 2) == 0, "one trace column per lane (lanes 28-31 gather the faces); 16-byte rows");
};

// host-side image of the tables (filled by hfx_refel_set)
template <int P>
inline void col_fill_tables(std::vector<double>& T, const double* aref, const double* sref, int np, const double* mf, int tp, const double* bref,
                            const double* fw, const double* fshape, const double* w, const double* shape) {
  using L = ColLayout<P>; using C = ColEl<P>;
  constexpr int nN = L::nN, t = L::t;
  T.assign(L::tEnd, 0.0);
  for (int r = 0; r < 3; r++) for (int m = 0; m < nN; m++) for (int k = 0; k < nN; k++) {
    T[L::tA + (r * nN + m) * nN + k] = aref[((size_t)r * nN + k) * np + m];
    T[L::tS + (r * nN + m) * nN + k] = sref[((size_t)r * nN + m) * np + k];
  }
  for (int a = 0; a < t; a++) for (int b = 0; b < t; b++) T[L::tMF + a * t + b] = mf[(size_t)a + (size_t)tp * b];
  for (int f = 0; f < 4; f++) for (int m = 0; m < nN; m++) for (int b = 0; b < t; b++) T[L::tBH + (f * nN + m) * t + b] = bref[((size_t)f * nN + m) * t + b];
  for (int a = 0; a < t; a++) for (int b = 0; b < t; b++) for (int c = 0; c < t; c++) {
    double s = 0.0;
    for (int ip = 0; ip < C::nIPf; ip++) s += fw[ip] * fshape[(size_t)ip * t + a] * fshape[(size_t)ip * t + b] * fshape[(size_t)ip * t + c];
    T[L::tT3 + (a * t + b) * t + c] = s;
  }
  for (int ip = 0; ip < C::nIP; ip++) for (int i = 0; i < nN; i++) T[L::tPHIW + ip * nN + i] = w[ip] * shape[(size_t)ip * nN + i];
  for (int m = 0; m < nN; m++) for (int k = 0; k < nN; k++) for (int f = 0; f < 4; f++) {
    const int a = col_nif<C>(f, m), b = col_nif<C>(f, k);
    T[L::tEF + (m * nN + k) * 4 + f] = (a >= 0 && b >= 0) ? mf[(size_t)a + (size_t)tp * b] : 0.0;
  }
  int* nifp = reinterpret_cast<int*>(T.data() + L::tNIF);
  for (int f = 0; f < 4; f++) for (int m = 0; m < nN; m++) nifp[f * nN + m] = col_nif<C>(f, m);
}
template <int P>
inline bool col_face_nodes_match(const int* faceNodes) {
  using C = ColEl<P>;
  if (P == 2) {   // the device copy of the inverse map
    constexpr int nifDev[4][10] = {{2, 1, -1, 0, 4, -1, -1, 5, 3, -1}, {-1, 1, 0, 2, -1, 3, -1, -1, 4, 5}, {2, -1, 0, 1, -1, -1, 5, 4, -1, 3}, {0, 1, 2, -1, 3, 4, 5, -1, -1, -1}};
    for (int f = 0; f < 4; f++) for (int m = 0; m < 10; m++) if (nifDev[f][m] != col_nif<C>(f, m)) return false;
  }
  for (int f = 0; f < 4; f++) for (int a = 0; a < C::t; a++) if (faceNodes[(size_t)f * C::t + a] != C::fn(f, a)) return false;
  for (int f = 0; f < 4; f++) for (int a = 0; a < 3; a++) if (C::fn(f, a) > 3) return false;   // the first three face nodes are the vertices of the face
  return true;
}

template <int P, int NW>
__global__ void __launch_bounds__(NW * 32) hdg_col_kernel(const AsmParams p) {
  using L = ColLayout<P>; using C = ColEl<P>;
  constexpr int nN = L::nN, t = L::t, l = L::l, nIP = L::nIP, TT = t * t;
  extern __shared__ __align__(16) double smc[];
  const int tid = threadIdx.x, g = tid >> 5, lane = tid & 31;
  double* const TB = smc;
  double* const E = smc + ev(L::tEnd) + g * L::stride;
  int* const EI = reinterpret_cast<int*>(E + L::oINT);
  long long* const RS = reinterpret_cast<long long*>(E + L::oRS);
  for (int i = tid; i < L::tEnd; i += NW * 32) TB[i] = p.colTab[i];
  __syncthreads();
  const double* const tA = TB + L::tA; const double* const tS = TB + L::tS; const double* const tMF = TB + L::tMF; const double* const tBH = TB + L::tBH;
  const double* const tT3 = TB + L::tT3; const double* const tPHIW = TB + L::tPHIW; const double* const tEF = TB + L::tEF;
  const int* const tNIF = reinterpret_cast<const int*>(TB + L::tNIF);
  const bool hasDiff = p.opmask & 1, hasSrc = (p.opmask & 8) && p.srcIP;
  const double dsc = hasDiff ? p.diffConst : 0.0;
  const int tv = p.tauVals;
  double* const K = E + L::oK; double* const KC = E + L::oKC; double* const SJ = E + L::oSJ; double* const FT = E + L::oFT;
  for (long long e = (long long)p.eBegin + (long long)blockIdx.x * NW + g; e < p.eEnd; e += (long long)gridDim.x * NW) {   // (warp-uniform)
    // ---- stage 1: gather ---------------------------------------------------------------------------------------------------------------------------------
    if (lane < 12) E[L::oX + lane] = p.elemX[(size_t)e * (nN * 3) + lane];                 // the four vertices span the element
    if (lane < l) EI[L::iPERM + lane] = p.fperm[(size_t)e * l + lane];
    if (lane < 16) EI[L::iPOS + lane] = p.elemPos[(size_t)e * 16 + lane];
    if (lane >= 28) {
      const int f = lane - 28, Ff = p.cell2face[(size_t)e * 4 + f];
      EI[L::iF + f] = Ff; EI[L::iBC + f] = p.faceBC[Ff]; EI[L::iIN + f] = p.faceInterior[Ff]; RS[f] = p.faceRowStart[Ff];
      EI[L::iSD + f] = tv == 2 ? p.tauSide[(size_t)e * 4 + f] : 0;
    }
    __syncwarp();
    // ---- stage 2: geometry (every lane keeps Jinv and det), per-face data by lanes 0-3, tau by lanes < l, source by lanes < nN --------------------------------------
    double J[3][3], det, I[3][3];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int m = 0; m < 3; m++) J[r][m] = 0.5 * (E[L::oX + (r + 1) * 3 + m] - E[L::oX + m]);
    det_inv(J, det, I);
    if (lane < 4) {
      const int f = lane;
      const int v0 = C::dfn(f, 0), v1 = C::dfn(f, 1), v2 = C::dfn(f, 2), vo = 6 - v0 - v1 - v2;
      double a0[3], a1[3], xo[3];
#pragma unroll
      for (int m = 0; m < 3; m++) {
        const double x0 = E[L::oX + v0 * 3 + m];
        a0[m] = 0.5 * (E[L::oX + v1 * 3 + m] - x0); a1[m] = 0.5 * (E[L::oX + v2 * 3 + m] - x0); xo[m] = E[L::oX + vo * 3 + m] - x0;
      }
      const double nv[3] = {a0[1] * a1[2] - a0[2] * a1[1], a0[2] * a1[0] - a0[0] * a1[2], a0[0] * a1[1] - a0[1] * a1[0]};
      const double nn = nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2];
      const double ar = sqrt(nn), inv = 1.0 / ar;
      const double prod = fma(xo[2], nv[2], fma(xo[1], nv[1], xo[0] * nv[0]));
      const double sg = prod > 0.0 ? -inv : inv;
      const double n0 = sg * nv[0], n1 = sg * nv[1], n2 = sg * nv[2], rdet = 1.0 / det;
      E[L::oN + f * 3] = n0; E[L::oN + f * 3 + 1] = n1; E[L::oN + f * 3 + 2] = n2; E[L::oAR + f] = ar;
#pragma unroll
      for (int r = 0; r < 3; r++) {
        E[L::oHF + f * 3 + r] = -ar * (I[0][r] * n0 + I[1][r] * n1 + I[2][r] * n2);
        E[L::oCR + f * 3 + r] = ar * rdet * (J[r][0] * n0 + J[r][1] * n1 + J[r][2] * n2);
      }
      E[L::oCQ + f * 3] = ar * rdet * n0; E[L::oCQ + f * 3 + 1] = ar * rdet * n1; E[L::oCQ + f * 3 + 2] = ar * rdet * n2;
    }
    if (lane < l) {
      const int f = lane / t;
      E[L::oTAU + lane] = p.tau[((size_t)EI[L::iF + f] * t + EI[L::iPERM + lane]) * tv + EI[L::iSD + f]];
    }
    if (lane < nN) {
      double fu = 0.0;
      if (hasSrc) {
#pragma unroll
        for (int ip = 0; ip < nIP; ip++) fu = fma(tPHIW[ip * nN + lane], p.srcIP[(size_t)e * nIP + ip] * det, fu);
      }
      E[L::oFU + lane] = fu;
    }
    __syncwarp();
    // ---- stage 3: tau masses FT_f = area_f sum_c T3[a][b][c] tau_fc and SJ_r ----------------------------------------------------------------------------------------
    for (int idx = lane; idx < 4 * TT; idx += 32) {
      const int f = idx / TT, ab = idx - f * TT;
      const double* t3 = tT3 + ab * t; const double* tau = E + L::oTAU + f * t;
      double s = 0.0;
#pragma unroll
      for (int c = 0; c < t; c++) s = fma(t3[c], tau[c], s);
      FT[idx] = E[L::oAR + f] * s;
    }
    {
      double G[3][3];
#pragma unroll
      for (int r = 0; r < 3; r++)
#pragma unroll
        for (int r2 = 0; r2 < 3; r2++) G[r][r2] = I[0][r] * I[0][r2] + I[1][r] * I[1][r2] + I[2][r] * I[2][r2];
      for (int idx = lane; idx < nN * nN; idx += 32) {
        const double s0 = tS[idx], s1 = tS[nN * nN + idx], s2 = tS[2 * nN * nN + idx];
        const double e0 = tEF[idx * 4], e1 = tEF[idx * 4 + 1], e2 = tEF[idx * 4 + 2], e3 = tEF[idx * 4 + 3];
#pragma unroll
        for (int r = 0; r < 3; r++) {
          const double v = G[r][0] * s0 + G[r][1] * s1 + G[r][2] * s2;
          const double w = E[L::oHF + r] * e0 + E[L::oHF + 3 + r] * e1 + E[L::oHF + 6 + r] * e2 + E[L::oHF + 9 + r] * e3;
          SJ[r * nN * nN + idx] = dsc * fma(det, v, w);
        }
      }
    }
    __syncwarp();
    // ---- stage 4: K = Suu - sum_r SJ_r A^_r (entry (m, n)) -------------------------------------------------------------------------------------------------------------
    for (int idx = lane; idx < nN * nN; idx += 32) {
      const int m = idx / nN, n = idx - m * nN;
      double v = 0.0;
#pragma unroll
      for (int f = 0; f < 4; f++) { const int a = tNIF[f * nN + m], b = tNIF[f * nN + n]; if (a >= 0 && b >= 0) v += FT[(f * t + a) * t + b]; }
#pragma unroll
      for (int r = 0; r < 3; r++) {
        const double* sj = SJ + (r * nN + m) * nN; const double* ar = tA + r * nN * nN + n;
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < nN; k++) s = fma(sj[k], ar[k * nN], s);
        v -= s;
      }
      K[idx] = v; KC[idx] = v;
    }
    __syncwarp();
    // ---- stage 5: K^-1 in place (unpivoted Gauss-Jordan, the warp holds each step's new entries in registers; K is definite for these models) ---------------------------
    {
      bool bad = false;
#pragma unroll 1
      for (int pv = 0; pv < nN; pv++) {
        const double kpp = K[pv * nN + pv];
        bad = bad || !(fabs(kpp) > 1e-300);
        const double piv = 1.0 / kpp;
        double nv[L::NQ];
#pragma unroll
        for (int q = 0; q < L::NQ; q++) {
          const int idx = lane + 32 * q;
          if (idx < nN * nN) {
            const int i = idx / nN, j2 = idx - i * nN;
            const double kip = K[i * nN + pv], kpj = K[pv * nN + j2], kij = K[idx];
            nv[q] = i == pv ? (j2 == pv ? piv : kpj * piv) : (j2 == pv ? -kip * piv : fma(-kip * piv, kpj, kij));
          }
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < L::NQ; q++) { const int idx = lane + 32 * q; if (idx < nN * nN) K[idx] = nv[q]; }
        __syncwarp();
      }
      if (bad && lane == 0) atomicOr(p.status, 1);
    }
    // ---- stage 6: one column per lane --------------------------------------------------------------------------------------------------------------------------------
    if (lane <= l) {
      const bool rhsCol = lane == l;
      const int c = rhsCol ? 0 : lane, fc = c / t, bcol = c - t * fc;
      const double* bh = tBH + fc * nN * t + bcol;      // B^_f[m][bcol] at bh[m * t]
      double U[nN], V[nN];
      if (!rhsCol) {
        const double cr0 = E[L::oCR + fc * 3], cr1 = E[L::oCR + fc * 3 + 1], cr2 = E[L::oCR + fc * 3 + 2];
        double bm[nN];
#pragma unroll
        for (int k = 0; k < nN; k++) bm[k] = bh[k * t];
#pragma unroll
        for (int m = 0; m < nN; m++) {
          double s[3];
#pragma unroll
          for (int r = 0; r < 3; r++) {
            const double2* sj = reinterpret_cast<const double2*>(SJ + (r * nN + m) * nN);
            double a = 0.0;
#pragma unroll
            for (int k2 = 0; k2 < nN / 2; k2++) { const double2 x = sj[k2]; a = fma(x.x, bm[2 * k2], a); a = fma(x.y, bm[2 * k2 + 1], a); }
            s[r] = a;
          }
          double v = fma(cr2, s[2], fma(cr1, s[1], cr0 * s[0]));
          const int a = tNIF[fc * nN + m];
          if (a >= 0) v -= FT[(fc * t + a) * t + bcol];      // Sul = -tau mass
          V[m] = v;                                           // R
        }
      } else {
#pragma unroll
        for (int m = 0; m < nN; m++) V[m] = -E[L::oFU + m];
      }
      auto matvec = [&](const double* M, const double (&x)[nN], double (&y)[nN], double sgn, bool acc) {   // y = (acc ? y : 0) + sgn M x
#pragma unroll
        for (int m = 0; m < nN; m++) {
          const double2* row = reinterpret_cast<const double2*>(M + m * nN);
          double a = 0.0;
#pragma unroll
          for (int k2 = 0; k2 < nN / 2; k2++) { const double2 v2 = row[k2]; a = fma(v2.x, x[2 * k2], a); a = fma(v2.y, x[2 * k2 + 1], a); }
          y[m] = acc ? fma(sgn, a, y[m]) : sgn * a;
        }
      };
      matvec(K, V, U, -1.0, false);      // U = -K^-1 R
      matvec(KC, U, V, 1.0, true);       // V = R + K U
      matvec(K, V, U, -1.0, true);       // U -= K^-1 V
      // Q_d[m] = cQ_fd B^_f[m] - sum_r Jinv(d,r) (A^_r U)[m], stored at once; Zq_f' accumulated on the way
      double zq[4][t];
#pragma unroll
      for (int f = 0; f < 4; f++)
#pragma unroll
        for (int b = 0; b < t; b++) zq[f][b] = 0.0;
      const double cq0 = rhsCol ? 0.0 : E[L::oCQ + fc * 3], cq1 = rhsCol ? 0.0 : E[L::oCQ + fc * 3 + 1], cq2 = rhsCol ? 0.0 : E[L::oCQ + fc * 3 + 2];
      double* const gQ = rhsCol ? p.Q0 + (size_t)e * (3 * nN) : p.Q + (size_t)e * (3 * nN) * l + c;
      const int qs = rhsCol ? 1 : l;
#pragma unroll
      for (int m = 0; m < nN; m++) {
        double pr[3];
#pragma unroll
        for (int r = 0; r < 3; r++) {
          const double2* row = reinterpret_cast<const double2*>(tA + (r * nN + m) * nN);
          double a = 0.0;
#pragma unroll
          for (int k2 = 0; k2 < nN / 2; k2++) { const double2 v2 = row[k2]; a = fma(v2.x, U[2 * k2], a); a = fma(v2.y, U[2 * k2 + 1], a); }
          pr[r] = a;
        }
        const double bmv = bh[m * t];
        const double q0 = fma(cq0, bmv, -(I[0][0] * pr[0] + I[0][1] * pr[1] + I[0][2] * pr[2]));
        const double q1 = fma(cq1, bmv, -(I[1][0] * pr[0] + I[1][1] * pr[1] + I[1][2] * pr[2]));
        const double q2 = fma(cq2, bmv, -(I[2][0] * pr[0] + I[2][1] * pr[1] + I[2][2] * pr[2]));
        gQ[(m * 3) * qs] = q0; gQ[(m * 3 + 1) * qs] = q1; gQ[(m * 3 + 2) * qs] = q2;
#pragma unroll
        for (int f = 0; f < 4; f++) {
          const int b = C::dnif(f, m);
          if (b >= 0) zq[f][b] = -dsc * (E[L::oN + f * 3] * q0 + E[L::oN + f * 3 + 1] * q1 + E[L::oN + f * 3 + 2] * q2);
        }
      }
      if (!rhsCol) {
        double* const gU = p.U + (size_t)e * nN * l + c;
#pragma unroll
        for (int m = 0; m < nN; m++) gU[m * l] = U[m];
      } else {
#pragma unroll
        for (int m = 0; m < nN; m++) p.U0[(size_t)e * nN + m] = U[m];
      }
      const int permC = rhsCol ? 0 : EI[L::iPERM + c];
      double* const gS = p.S ? p.S + (size_t)e * l * l : nullptr;
#pragma unroll
      for (int f = 0; f < 4; f++) {
        const double ar = E[L::oAR + f];
        double uf[t];
#pragma unroll
        for (int b = 0; b < t; b++) uf[b] = U[C::dfn(f, b)] - ((!rhsCol && f == fc && b == bcol) ? 1.0 : 0.0);
        const int bcf = EI[L::iBC + f], Ff = EI[L::iF + f];
        const bool inter = EI[L::iIN + f] != 0;
        double* const blk = p.vals + RS[f] + (long long)EI[L::iPOS + f * 4 + fc] * TT + permC;
#pragma unroll
        for (int a = 0; a < t; a++) {
          const double* ft = FT + (f * t + a) * t; const double* mfa = tMF + a * t;
          double s1 = 0.0, s2 = 0.0;
#pragma unroll
          for (int b = 0; b < t; b++) { s1 = fma(ft[b], uf[b], s1); s2 = fma(mfa[b], zq[f][b], s2); }
          double sv = fma(ar, s2, s1);
          const int r = f * t + a, pr = EI[L::iPERM + r];
          if (!rhsCol) {
            if (bcf == 1) sv = (r == c) ? 1.0 : 0.0;                                          // DirichletModel row (Set)
            else if (bcf == 2) sv = (f == fc) ? ar * mfa[bcol] : 0.0;                          // IntegratedDirichletModel row: face mass
            if (gS) gS[r + l * c] = sv;
            double* dst = blk + pr * t;
            if (f == fc && inter) atomicAdd(dst, sv); else *dst = sv;
          } else {
            double s0 = -sv;
            if (bcf == 1) s0 = p.dirichlet[(size_t)Ff * t + a];
            else if (bcf == 2) {
              s0 = 0.0;
#pragma unroll
              for (int b = 0; b < t; b++) s0 = fma(ar * mfa[b], p.dirichlet[(size_t)Ff * t + b], s0);
            }
            if (p.S0) p.S0[(size_t)e * l + r] = s0;
            double* dst = p.rhs + (size_t)Ff * t + pr;
            if (inter) atomicAdd(dst, s0); else *dst = s0;
          }
        }
      }
    }
    __syncwarp();
  }
}

template <int P, int NW>
inline cudaError_t launch_col(const AsmParams& p, int nSM, cudaStream_t st) {
  using L = ColLayout<P>;
  const size_t bytes = (size_t)(ev(L::tEnd) + NW * L::stride) * sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(hdg_col_kernel<P, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return e;
  int perSM = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, hdg_col_kernel<P, NW>, NW * 32, bytes);
  if (perSM < 1) perSM = 1;
  long long grid = (long long)nSM * perSM;
  const long long need = ((long long)(p.eEnd - p.eBegin) + NW - 1) / NW;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  hdg_col_kernel<P, NW><<<(int)grid, NW * 32, bytes, st>>>(p);
  return cudaGetLastError();
}

}  // namespace hfx
