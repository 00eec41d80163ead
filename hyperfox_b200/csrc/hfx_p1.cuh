// Linear tetrahedra (3-D order 1: 4 nodes, 3 face nodes, local system 4 + 12 + 12), ONE THREAD PER ELEMENT.
// A linear tet carries 16.6 kflop and 3.1 KB of traffic in the reference's count: it is bound by HBM, not by the FP64 pipe, provided enough elements are in
// flight -- which the element-group kernel of hfx_assemble.cuh (one warp and 13 KB of shared memory per element: 16 elements per SM) cannot offer.  Here the whole
// element lives in the registers and a 0.7 KB shared-memory slice of one thread (256 elements in flight per SM), the reference matrices in constant memory,
// and the algebra is the SJ_r formulation of hfx_big.cuh with closed forms where a 4 x 4 system allows them:
//   SJ_r = ts c (detJ sum_r' G(r,r') S^_r' + sum_f h_fr E_f)        K = Suu - sum_r SJ_r A^_r        K^-1 by the adjugate (2x2 minors), refined through U
//   per trace column c = (f, b):  R = Sul + cR_f . (SJ_r B^_f)      U = -K^-1 R, U -= K^-1 (K U + R)   Q_d = -sum_r Jinv(d,r) (A^_r U) + cQ_fd B^_f
//                                 S[(f',a)][c] = FT_f' (U_f' - I) + area_f' M^f Zq_f'
// Models: Base + Diffusion with D = c I (+ Source), straight-sided cells, any tau (the tau mass is the cubature sum of the reference: area sum_c T3[a][b][c] tau_c),
// both boundary models.  Everything else on linear tets stays with the element-group kernel.
// Reference semantics: Operator.cpp:14-84, HDGBase.cpp:18-158, HDGDiffusion.cpp:74-145, Source.cpp:24-48, HDGSolver.cpp:231-348 (gather + condensation),
// :361-529 (boundary rows), :531-675 (scatter).
#pragma once
#include <type_traits>
#include "hfx_assemble.cuh"

namespace hfx {

struct P1Tables {
  double A[3][4][4];     // A^_r[m][k']  = M_ref^-1 S^_r
  double S[3][4][4];     // S^_r[m][k']  = sum_ip w dphi_m/dxi_r phi_k'
  double MF[3][3];       // reference face mass
  double BH[4][4][3];    // B^_f[m][b]   = M_ref^-1[:, faceNodes_f] M^f
  double T3[3][3][3];    // sum_ip fw phi_a phi_b phi_c on the face element (the reference's face cubature, not the exact integral)
  double PHIW[4][4];     // w[ip] phi_i(ip)
};
__constant__ P1Tables c_p1;
// face-node map of the linear tetrahedron (ReferenceElement.cpp:636-878; checked against the host tables before the kernel is used): compile-time, so
// that every per-node array of the element stays in registers
__device__ constexpr int kP1FN[4][3] = {{3, 1, 0}, {2, 1, 3}, {2, 3, 0}, {0, 1, 2}};
__device__ constexpr int kP1NIF[4][4] = {{2, 1, -1, 0}, {-1, 1, 0, 2}, {2, -1, 0, 1}, {0, 1, 2, -1}};   // node -> position in face (or -1)
__device__ constexpr int kP1OPP[4] = {2, 0, 1, 3};                                                        // first node not on the face

constexpr int kP1Threads = 128;
constexpr int kP1SmemDoubles = 48 + 36 + 24;   // per thread: SJ [3][4][4], FT [4][3][3], cR [4][3], cQ [4][3]

__global__ void __launch_bounds__(kP1Threads, 2) hdg_p1_kernel(const AsmParams p) {
  extern __shared__ double sm1[];
  const int tid = threadIdx.x;
  // shared-memory slice of this thread: index-major, thread-minor (conflict free)
#define SJ_(r, m, k) sm1[(((r) * 4 + (m)) * 4 + (k)) * kP1Threads + tid]
#define FT_(f, a, b) sm1[(48 + ((f) * 3 + (a)) * 3 + (b)) * kP1Threads + tid]
#define CR_(f, r) sm1[(84 + (f) * 3 + (r)) * kP1Threads + tid]
#define CQ_(f, d) sm1[(96 + (f) * 3 + (d)) * kP1Threads + tid]
  const P1Tables& T = c_p1;
  const bool hasDiff = p.opmask & 1, hasSrc = (p.opmask & 8) && p.srcIP;
  const double dsc = hasDiff ? p.diffConst : 0.0;
  for (long long e = (long long)p.eBegin + (long long)blockIdx.x * kP1Threads + tid; e < p.eEnd; e += (long long)gridDim.x * kP1Threads) {
    // ---- gather ------------------------------------------------------------------------------------------------------------------------
    double X[4][3];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int m = 0; m < 3; m++) X[i][m] = p.elemX[(size_t)e * 12 + i * 3 + m];
    int F[4];
#pragma unroll
    for (int f = 0; f < 4; f++) F[f] = p.cell2face[(size_t)e * 4 + f];
    unsigned long long permBits = 0ull, posBits = 0ull;   // 12 face-node positions and 16 block positions, four bits each (indexed at run time)
#pragma unroll
    for (int i = 0; i < 12; i++) permBits |= (unsigned long long)p.fperm[(size_t)e * 12 + i] << (4 * i);
#pragma unroll
    for (int i = 0; i < 16; i++) posBits |= (unsigned long long)p.elemPos[(size_t)e * 16 + i] << (4 * i);
    auto perm = [&](int i) { return (int)((permBits >> (4 * i)) & 15ull); };
    // ---- geometry (constant Jacobian; outward unit normals, HDGBase.cpp:43-62) ------------------------------------------------------------
    double J[3][3], det, I[3][3];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int m = 0; m < 3; m++) J[r][m] = 0.5 * (X[r + 1][m] - X[0][m]);
    det_inv(J, det, I);
    const double rdet = 1.0 / det;
    double nrm[4][3], area[4];
#pragma unroll
    for (int f = 0; f < 4; f++) {
      const int v0 = kP1FN[f][0], v1 = kP1FN[f][1], v2 = kP1FN[f][2], vo = kP1OPP[f];
      double a0[3], a1[3];
#pragma unroll
      for (int m = 0; m < 3; m++) { a0[m] = 0.5 * (X[v1][m] - X[v0][m]); a1[m] = 0.5 * (X[v2][m] - X[v0][m]); }
      double nv[3] = {a0[1] * a1[2] - a0[2] * a1[1], a0[2] * a1[0] - a0[0] * a1[2], a0[0] * a1[1] - a0[1] * a1[0]};
      const double nn = nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2];
      const double ar = sqrt(nn), inv = 1.0 / ar;
      double prod = 0.0;
#pragma unroll
      for (int m = 0; m < 3; m++) prod = fma(X[vo][m] - X[v0][m], nv[m], prod);
      const double sg = prod > 0.0 ? -inv : inv;
#pragma unroll
      for (int m = 0; m < 3; m++) nrm[f][m] = sg * nv[m];
      area[f] = ar;
    }
    // ---- tau masses FT_f = area_f sum_c T3[a][b][c] tau_fc (HDGBase.cpp:112-131 with the face cubature of the reference) --------------------
#pragma unroll
    for (int f = 0; f < 4; f++) {
      const int side = (p.tauVals == 2) ? p.tauSide[(size_t)e * 4 + f] : 0;
      double tau[3];
#pragma unroll
      for (int c = 0; c < 3; c++) tau[c] = p.tau[((size_t)F[f] * 3 + perm(f * 3 + c)) * p.tauVals + side];
#pragma unroll
      for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) FT_(f, a, b) = area[f] * (T.T3[a][b][0] * tau[0] + T.T3[a][b][1] * tau[1] + T.T3[a][b][2] * tau[2]);
    }
    // ---- SJ_r and K ----------------------------------------------------------------------------------------------------------------------
    double G[3][3], hf[4][3];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int r2 = 0; r2 < 3; r2++) G[r][r2] = I[0][r] * I[0][r2] + I[1][r] * I[1][r2] + I[2][r] * I[2][r2];
#pragma unroll
    for (int f = 0; f < 4; f++)
#pragma unroll
      for (int r = 0; r < 3; r++) hf[f][r] = -area[f] * (I[0][r] * nrm[f][0] + I[1][r] * nrm[f][1] + I[2][r] * nrm[f][2]);
    // per-face epilogue coefficients: cR_fr = area_f nu_fr / detJ (nu_fr = sum_d J(r,d) n_fd), cQ_fd = area_f n_fd / detJ
#pragma unroll
    for (int f = 0; f < 4; f++)
#pragma unroll
      for (int r = 0; r < 3; r++) {
        CR_(f, r) = area[f] * rdet * (J[r][0] * nrm[f][0] + J[r][1] * nrm[f][1] + J[r][2] * nrm[f][2]);
        CQ_(f, r) = area[f] * rdet * nrm[f][r];
      }
    double K[4][4];
#pragma unroll
    for (int m = 0; m < 4; m++)
#pragma unroll
      for (int n = 0; n < 4; n++) {
        double v = 0.0;
#pragma unroll
        for (int f = 0; f < 4; f++) { const int a = kP1NIF[f][m], b = kP1NIF[f][n]; if (a >= 0 && b >= 0) v += FT_(f, a, b); }
        K[m][n] = v;
      }
#pragma unroll
    for (int m = 0; m < 4; m++)
#pragma unroll
      for (int k = 0; k < 4; k++) {
        double ef[4];
#pragma unroll
        for (int f = 0; f < 4; f++) { const int a = kP1NIF[f][m], b = kP1NIF[f][k]; ef[f] = (a >= 0 && b >= 0) ? T.MF[a][b] : 0.0; }
#pragma unroll
        for (int r = 0; r < 3; r++) {
          const double v = G[r][0] * T.S[0][m][k] + G[r][1] * T.S[1][m][k] + G[r][2] * T.S[2][m][k];
          const double w = hf[0][r] * ef[0] + hf[1][r] * ef[1] + hf[2][r] * ef[2] + hf[3][r] * ef[3];
          SJ_(r, m, k) = dsc * fma(det, v, w);
        }
      }
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int m = 0; m < 4; m++) {
        const double s0 = SJ_(r, m, 0), s1 = SJ_(r, m, 1), s2 = SJ_(r, m, 2), s3 = SJ_(r, m, 3);
#pragma unroll
        for (int n = 0; n < 4; n++) K[m][n] -= s0 * T.A[r][0][n] + s1 * T.A[r][1][n] + s2 * T.A[r][2][n] + s3 * T.A[r][3][n];
      }
    // ---- K^-1 (adjugate from the 2x2 minors of the two row pairs) ---------------------------------------------------------------------------
    double Ki[4][4];
    {
      const double s0 = K[0][0] * K[1][1] - K[1][0] * K[0][1], s1 = K[0][0] * K[1][2] - K[1][0] * K[0][2], s2 = K[0][0] * K[1][3] - K[1][0] * K[0][3];
      const double s3 = K[0][1] * K[1][2] - K[1][1] * K[0][2], s4 = K[0][1] * K[1][3] - K[1][1] * K[0][3], s5 = K[0][2] * K[1][3] - K[1][2] * K[0][3];
      const double c5 = K[2][2] * K[3][3] - K[3][2] * K[2][3], c4 = K[2][1] * K[3][3] - K[3][1] * K[2][3], c3 = K[2][1] * K[3][2] - K[3][1] * K[2][2];
      const double c2 = K[2][0] * K[3][3] - K[3][0] * K[2][3], c1 = K[2][0] * K[3][2] - K[3][0] * K[2][2], c0 = K[2][0] * K[3][1] - K[3][0] * K[2][1];
      const double dK = s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0;
      if (!(fabs(dK) > 1e-300)) atomicOr(p.status, 1);
      const double id = 1.0 / dK;
      Ki[0][0] = (K[1][1] * c5 - K[1][2] * c4 + K[1][3] * c3) * id;  Ki[0][1] = (-K[0][1] * c5 + K[0][2] * c4 - K[0][3] * c3) * id;
      Ki[0][2] = (K[3][1] * s5 - K[3][2] * s4 + K[3][3] * s3) * id;  Ki[0][3] = (-K[2][1] * s5 + K[2][2] * s4 - K[2][3] * s3) * id;
      Ki[1][0] = (-K[1][0] * c5 + K[1][2] * c2 - K[1][3] * c1) * id; Ki[1][1] = (K[0][0] * c5 - K[0][2] * c2 + K[0][3] * c1) * id;
      Ki[1][2] = (-K[3][0] * s5 + K[3][2] * s2 - K[3][3] * s1) * id; Ki[1][3] = (K[2][0] * s5 - K[2][2] * s2 + K[2][3] * s1) * id;
      Ki[2][0] = (K[1][0] * c4 - K[1][1] * c2 + K[1][3] * c0) * id;  Ki[2][1] = (-K[0][0] * c4 + K[0][1] * c2 - K[0][3] * c0) * id;
      Ki[2][2] = (K[3][0] * s4 - K[3][1] * s2 + K[3][3] * s0) * id;  Ki[2][3] = (-K[2][0] * s4 + K[2][1] * s2 - K[2][3] * s0) * id;
      Ki[3][0] = (-K[1][0] * c3 + K[1][1] * c1 - K[1][2] * c0) * id; Ki[3][1] = (K[0][0] * c3 - K[0][1] * c1 + K[0][2] * c0) * id;
      Ki[3][2] = (-K[3][0] * s3 + K[3][1] * s1 - K[3][2] * s0) * id; Ki[3][3] = (K[2][0] * s3 - K[2][1] * s1 + K[2][2] * s0) * id;
    }
    long long rowStart[4]; int bc[4], inter[4];
#pragma unroll
    for (int f = 0; f < 4; f++) { rowStart[f] = p.faceRowStart[F[f]]; bc[f] = p.faceBC[F[f]]; inter[f] = p.faceInterior[F[f]]; }
    // right-hand side: Fu = detJ sum_ip w phi_i src (Source.cpp:24-48)
    double Fu[4] = {0.0, 0.0, 0.0, 0.0};
    if (hasSrc) {
#pragma unroll
      for (int ip = 0; ip < 4; ip++) {
        const double sv = p.srcIP[(size_t)e * 4 + ip] * det;
#pragma unroll
        for (int i = 0; i < 4; i++) Fu[i] = fma(T.PHIW[ip][i], sv, Fu[i]);
      }
    }
    double* const gU = p.U + (size_t)e * 4 * 12;
    double* const gQ = p.Q + (size_t)e * 12 * 12;
    double* const gS = p.S ? p.S + (size_t)e * 144 : nullptr;
    // ---- one trace column at a time.  fc is a compile-time constant of each expansion (per-face register arrays stay in registers), bcol a run-time index
    //      into constant / shared memory only; rhsCol: the right-hand-side column (U0, Q0, S0) ---------------------------------------------------------
    auto column = [&](auto FC, int bcol, bool rhsCol) {
      constexpr int fc = decltype(FC)::value;
      const int c = fc * 3 + bcol;
      double R[4];
      if (!rhsCol) {
        const double b0 = T.BH[fc][0][bcol], b1 = T.BH[fc][1][bcol], b2 = T.BH[fc][2][bcol], b3 = T.BH[fc][3][bcol];
        const double cr0 = CR_(fc, 0), cr1 = CR_(fc, 1), cr2 = CR_(fc, 2);
#pragma unroll
        for (int m = 0; m < 4; m++) {
          double v = cr0 * (SJ_(0, m, 0) * b0 + SJ_(0, m, 1) * b1 + SJ_(0, m, 2) * b2 + SJ_(0, m, 3) * b3);
          v = fma(cr1, SJ_(1, m, 0) * b0 + SJ_(1, m, 1) * b1 + SJ_(1, m, 2) * b2 + SJ_(1, m, 3) * b3, v);
          v = fma(cr2, SJ_(2, m, 0) * b0 + SJ_(2, m, 1) * b1 + SJ_(2, m, 2) * b2 + SJ_(2, m, 3) * b3, v);
          if (kP1NIF[fc][m] >= 0) v -= FT_(fc, kP1NIF[fc][m] >= 0 ? kP1NIF[fc][m] : 0, bcol);      // Sul = -tau mass
          R[m] = v;
        }
      } else {
#pragma unroll
        for (int m = 0; m < 4; m++) R[m] = -Fu[m];
      }
      double U[4], V[4];
#pragma unroll
      for (int m = 0; m < 4; m++) U[m] = -(Ki[m][0] * R[0] + Ki[m][1] * R[1] + Ki[m][2] * R[2] + Ki[m][3] * R[3]);
#pragma unroll
      for (int m = 0; m < 4; m++) V[m] = R[m] + K[m][0] * U[0] + K[m][1] * U[1] + K[m][2] * U[2] + K[m][3] * U[3];
#pragma unroll
      for (int m = 0; m < 4; m++) U[m] -= Ki[m][0] * V[0] + Ki[m][1] * V[1] + Ki[m][2] * V[2] + Ki[m][3] * V[3];
      double Q[3][4];
      {
        double Pr[3][4];
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
          for (int m = 0; m < 4; m++) Pr[r][m] = T.A[r][m][0] * U[0] + T.A[r][m][1] * U[1] + T.A[r][m][2] * U[2] + T.A[r][m][3] * U[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
          const double cq = rhsCol ? 0.0 : CQ_(fc, d);
#pragma unroll
          for (int m = 0; m < 4; m++) {
            double v = -(I[d][0] * Pr[0][m] + I[d][1] * Pr[1][m] + I[d][2] * Pr[2][m]);
            if (!rhsCol) v = fma(cq, T.BH[fc][m][bcol], v);
            Q[d][m] = v;
          }
        }
      }
      if (!rhsCol) {
#pragma unroll
        for (int m = 0; m < 4; m++) gU[m * 12 + c] = U[m];
#pragma unroll
        for (int m = 0; m < 4; m++)
#pragma unroll
          for (int d = 0; d < 3; d++) gQ[(m * 3 + d) * 12 + c] = Q[d][m];
      } else {
#pragma unroll
        for (int m = 0; m < 4; m++) p.U0[(size_t)e * 4 + m] = U[m];
#pragma unroll
        for (int m = 0; m < 4; m++)
#pragma unroll
          for (int d = 0; d < 3; d++) p.Q0[(size_t)e * 12 + m * 3 + d] = Q[d][m];
      }
      const int permC = rhsCol ? 0 : perm(c);
      // S column: rows (f, a)
#pragma unroll
      for (int f = 0; f < 4; f++) {
        double zq[3], uf[3];
#pragma unroll
        for (int b = 0; b < 3; b++) {
          constexpr int dummy = 0; (void)dummy;
          const int nd = kP1FN[f][b];
          zq[b] = -dsc * (nrm[f][0] * Q[0][nd] + nrm[f][1] * Q[1][nd] + nrm[f][2] * Q[2][nd]);
          uf[b] = U[nd] - ((!rhsCol && f == fc && b == bcol) ? 1.0 : 0.0);
        }
#pragma unroll
        for (int a = 0; a < 3; a++) {
          double s2 = FT_(f, a, 0) * uf[0] + FT_(f, a, 1) * uf[1] + FT_(f, a, 2) * uf[2] + area[f] * (T.MF[a][0] * zq[0] + T.MF[a][1] * zq[1] + T.MF[a][2] * zq[2]);
          const int r = f * 3 + a;
          if (!rhsCol) {
            if (bc[f] == 1) s2 = (r == c) ? 1.0 : 0.0;                                          // DirichletModel row (Set)
            else if (bc[f] == 2) s2 = (f == fc) ? area[f] * T.MF[a][bcol] : 0.0;                // IntegratedDirichletModel row: face mass
            if (gS) gS[r + 12 * c] = s2;
            double* dst = p.vals + rowStart[f] + (long long)((posBits >> (4 * (f * 4 + fc))) & 15ull) * 9 + perm(r) * 3 + permC;
            if (f == fc && inter[f]) atomicAdd(dst, s2); else *dst = s2;
          } else {
            double s0 = -s2;
            if (bc[f] == 1) s0 = p.dirichlet[(size_t)F[f] * 3 + a];
            else if (bc[f] == 2) {
              s0 = 0.0;
#pragma unroll
              for (int b = 0; b < 3; b++) s0 = fma(area[f] * T.MF[a][b], p.dirichlet[(size_t)F[f] * 3 + b], s0);
            }
            if (p.S0) p.S0[(size_t)e * 12 + r] = s0;
            double* dst = p.rhs + (size_t)F[f] * 3 + perm(r);
            if (inter[f]) atomicAdd(dst, s0); else *dst = s0;
          }
        }
      }
    };
#pragma unroll 1
    for (int bcol = 0; bcol < 3; bcol++) {
      column(std::integral_constant<int, 0>{}, bcol, false);
      column(std::integral_constant<int, 1>{}, bcol, false);
      column(std::integral_constant<int, 2>{}, bcol, false);
      column(std::integral_constant<int, 3>{}, bcol, false);
    }
    column(std::integral_constant<int, 0>{}, 0, true);
  }
#undef SJ_
#undef FT_
#undef CR_
#undef CQ_
}

inline cudaError_t launch_p1(const AsmParams& p, int nSM, cudaStream_t st) {
  const size_t bytes = (size_t)kP1SmemDoubles * kP1Threads * sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(hdg_p1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return e;
  int perSM = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, hdg_p1_kernel, kP1Threads, bytes);
  if (perSM < 1) perSM = 1;
  long long grid = (long long)nSM * perSM;
  const long long need = ((long long)(p.eEnd - p.eBegin) + kP1Threads - 1) / kP1Threads;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  hdg_p1_kernel<<<(int)grid, kP1Threads, bytes, st>>>(p);
  return cudaGetLastError();
}


// ---------------------------------------------------------------------------------------------------------------------------------------------------------
// SIXTEEN LANES PER ELEMENT (two elements per warp).  The columns of the condensation are independent once K^-1 is known: lane c < 12 owns trace column
// c = (face, face node) -- its column of R, U, Q and S -- lane 12 the right-hand-side column (U0, Q0, S0); the shared part of the element (geometry, tau masses,
// SJ_r, K, K^-1: one entry per lane) lives in a 1.7 KB shared-memory slice.  A lane holds one column (about 45 doubles) instead of the whole element, the rows of
// U and Q leave as 96-byte runs, and 32 elements are in flight per SM.  Same algebra, tables and eligibility as hdg_p1_kernel above.
constexpr int kPGThreads = 256, kPGElems = kPGThreads / 16;
constexpr int PG_X = 0, PG_N = 12, PG_AR = 24, PG_FT = 28, PG_SJ = 64, PG_CR = 112, PG_CQ = 124, PG_K = 136, PG_KI = 152, PG_FU = 168, PG_TAU = 172, PG_HF = 184, PG_RS = 196,
              PG_INT = 200, PG_STRIDE = 224;   // ints at PG_INT: F[4], BC[4], INTR[4], SIDE[4], PERM[12], POS[16]
constexpr int PGI_F = 0, PGI_BC = 4, PGI_IN = 8, PGI_SD = 12, PGI_PERM = 16, PGI_POS = 28;
constexpr int kPGTab = (int)(sizeof(P1Tables) / sizeof(double)), kPGTabAll = kPGTab + 64;   // + EF[(m,k)][f] = MF[nif(f,m)][nif(f,k)] or 0
constexpr int kHNIF[4][4] = {{2, 1, -1, 0}, {-1, 1, 0, 2}, {2, -1, 0, 1}, {0, 1, 2, -1}};
__host__ __device__ constexpr unsigned long long p1_nif_bits() {
  unsigned long long b = 0ull;
  for (int f = 0; f < 4; f++) for (int m = 0; m < 4; m++) b |= (unsigned long long)(kHNIF[f][m] & 15) << (4 * (f * 4 + m));
  return b;
}
constexpr unsigned long long kNifBits = p1_nif_bits();
constexpr int kHFN[4][3] = {{3, 1, 0}, {2, 1, 3}, {2, 3, 0}, {0, 1, 2}};
constexpr int kHOPP[4] = {2, 0, 1, 3};
__host__ __device__ constexpr unsigned p1_fn_bits() {   // two bits per (f, b), then two bits per opposite node
  unsigned b = 0u;
  for (int f = 0; f < 4; f++) for (int a = 0; a < 3; a++) b |= (unsigned)kHFN[f][a] << (2 * (f * 3 + a));
  for (int f = 0; f < 4; f++) b |= (unsigned)kHOPP[f] << (24 + 2 * f);
  return b;
}
constexpr unsigned kFnBits = p1_fn_bits();

__device__ __forceinline__ int lane_base16(int tid) { return (tid & 31) & 16; }

template <int MINB>
__global__ void __launch_bounds__(kPGThreads, MINB) hdg_p1g_kernel(const AsmParams p) {
  extern __shared__ __align__(16) double smg[];
  const int tid = threadIdx.x, g = tid >> 4, j = tid & 15;
  double* const TB = smg;                                             // tables (CTA-wide)
  double* const E = smg + ((kPGTabAll + 1) & ~1) + g * PG_STRIDE;     // this element's slice
  int* const EI = reinterpret_cast<int*>(E + PG_INT);
  long long* const RS = reinterpret_cast<long long*>(E + PG_RS);
  {
    const double* cp = reinterpret_cast<const double*>(&c_p1);
    for (int i = tid; i < kPGTab; i += kPGThreads) TB[i] = cp[i];
    if (tid < 64) {
      const int mk = tid >> 2, f = tid & 3, m = mk >> 2, k = mk & 3;
      const int a = (int)((kNifBits >> (4 * (f * 4 + m))) & 15ull), b = (int)((kNifBits >> (4 * (f * 4 + k))) & 15ull);
      TB[kPGTab + tid] = (a != 15 && b != 15) ? c_p1.MF[a][b] : 0.0;
    }
  }
  __syncthreads();
  const P1Tables& T = *reinterpret_cast<const P1Tables*>(TB);
  const double* const EF = TB + kPGTab;
  auto nif = [](int f, int m) { return (int)((kNifBits >> (4 * (f * 4 + m))) & 15ull); };   // 15: the node is not on the face
  const bool hasDiff = p.opmask & 1, hasSrc = (p.opmask & 8) && p.srcIP;
  const double dsc = hasDiff ? p.diffConst : 0.0;
  const int tv = p.tauVals;
  // The gather of an element (coordinates, face ids -> row starts / flags, permutations, then tau through them) is two dependent trips to global memory: it is
  // issued one element ahead into registers, behind the column stage of the previous element.
  double nX = 0.0, nTau = 0.0; int nPerm = 0, nPos = 0, nF = 0, nBC = 0, nIN = 0, nSD = 0; long long nRS = 0;
  const unsigned full = 0xffffffffu;
  const int half = lane_base16(tid);
  auto loadElem = [&](long long e) {
    if (j < 12) { nX = p.elemX[(size_t)e * 12 + j]; nPerm = p.fperm[(size_t)e * 12 + j]; }
    nPos = p.elemPos[(size_t)e * 16 + j];
    if (j >= 12) {
      const int f = j - 12;
      nF = p.cell2face[(size_t)e * 4 + f]; nBC = p.faceBC[nF]; nIN = p.faceInterior[nF]; nRS = p.faceRowStart[nF];
      nSD = tv == 2 ? p.tauSide[(size_t)e * 4 + f] : 0;
    }
  };
  auto loadTau = [&]() {   // lanes 0-11: tau of face j / 3 at face node j % 3; the face id and the side sit in lane 12 + j / 3 of the same half warp
    const int src = half + 12 + (j < 12 ? j / 3 : 0);
    const int Ff = __shfl_sync(full, nF, src), sd = __shfl_sync(full, nSD, src);
    if (j < 12) nTau = p.tau[((size_t)Ff * 3 + nPerm) * tv + sd];
  };
  const long long eFirst = (long long)p.eBegin + (long long)blockIdx.x * kPGElems;
  if (eFirst < p.eEnd) { loadElem(eFirst + g < p.eEnd ? eFirst + g : (long long)p.eEnd - 1); loadTau(); }
  for (long long e0 = eFirst; e0 < p.eEnd; e0 += (long long)gridDim.x * kPGElems) {
    const bool live = e0 + g < p.eEnd;
    const long long e = live ? e0 + g : (long long)p.eEnd - 1;
    // ---- stage 1: the prefetched gather goes to the element's slice ------------------------------------------------------------------------------------------
    if (j < 12) { E[PG_X + j] = nX; EI[PGI_PERM + j] = nPerm; E[PG_TAU + j] = nTau; }
    EI[PGI_POS + j] = nPos;
    if (j >= 12) { const int f = j - 12; EI[PGI_F + f] = nF; EI[PGI_BC + f] = nBC; EI[PGI_IN + f] = nIN; RS[f] = nRS; EI[PGI_SD + f] = nSD; }
    {
      const long long en0 = e0 + (long long)gridDim.x * kPGElems;
      if (en0 < p.eEnd) loadElem(en0 + g < p.eEnd ? en0 + g : (long long)p.eEnd - 1);      // (CTA-uniform condition)
    }
    __syncwarp();
    // ---- stage 2: geometry (every lane keeps Jinv and det), per-face data by lanes 0-3, tau by lanes 0-11, source by lane 12 --------------------------------
    double J[3][3], det, I[3][3];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int m = 0; m < 3; m++) J[r][m] = 0.5 * (E[PG_X + (r + 1) * 3 + m] - E[PG_X + m]);
    det_inv(J, det, I);
    if (j < 4) {
      const int f = j;
      const int v0 = (kFnBits >> (2 * (f * 3))) & 3, v1 = (kFnBits >> (2 * (f * 3 + 1))) & 3, v2 = (kFnBits >> (2 * (f * 3 + 2))) & 3, vo = (kFnBits >> (24 + 2 * f)) & 3;
      double a0[3], a1[3], xo[3];
#pragma unroll
      for (int m = 0; m < 3; m++) {
        const double x0 = E[PG_X + v0 * 3 + m];
        a0[m] = 0.5 * (E[PG_X + v1 * 3 + m] - x0); a1[m] = 0.5 * (E[PG_X + v2 * 3 + m] - x0); xo[m] = E[PG_X + vo * 3 + m] - x0;
      }
      const double nv[3] = {a0[1] * a1[2] - a0[2] * a1[1], a0[2] * a1[0] - a0[0] * a1[2], a0[0] * a1[1] - a0[1] * a1[0]};
      const double nn = nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2];
      const double ar = sqrt(nn), inv = 1.0 / ar;
      const double prod = fma(xo[2], nv[2], fma(xo[1], nv[1], xo[0] * nv[0]));
      const double sg = prod > 0.0 ? -inv : inv;
      const double n0 = sg * nv[0], n1 = sg * nv[1], n2 = sg * nv[2], rdet = 1.0 / det;
      E[PG_N + f * 3] = n0; E[PG_N + f * 3 + 1] = n1; E[PG_N + f * 3 + 2] = n2; E[PG_AR + f] = ar;
#pragma unroll
      for (int r = 0; r < 3; r++) {
        E[PG_HF + f * 3 + r] = -ar * (I[0][r] * n0 + I[1][r] * n1 + I[2][r] * n2);
        E[PG_CR + f * 3 + r] = ar * rdet * (J[r][0] * n0 + J[r][1] * n1 + J[r][2] * n2);
      }
      E[PG_CQ + f * 3] = ar * rdet * n0; E[PG_CQ + f * 3 + 1] = ar * rdet * n1; E[PG_CQ + f * 3 + 2] = ar * rdet * n2;
    }
    if (j == 12) {
      double Fu[4] = {0.0, 0.0, 0.0, 0.0};
      if (hasSrc) {
#pragma unroll
        for (int ip = 0; ip < 4; ip++) {
          const double sv = p.srcIP[(size_t)e * 4 + ip] * det;
#pragma unroll
          for (int i = 0; i < 4; i++) Fu[i] = fma(T.PHIW[ip][i], sv, Fu[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; i++) E[PG_FU + i] = Fu[i];
    }
    __syncwarp();
    // ---- stage 3: tau masses (lanes 0-11: row (f, a)) and SJ_r (lane = (m, k)) --------------------------------------------------------------------------------
    if (j < 12) {
      const int f = j / 3, a = j - 3 * f;
      const double ar = E[PG_AR + f], t0 = E[PG_TAU + f * 3], t1 = E[PG_TAU + f * 3 + 1], t2 = E[PG_TAU + f * 3 + 2];
#pragma unroll
      for (int b = 0; b < 3; b++) E[PG_FT + j * 3 + b] = ar * (T.T3[a][b][0] * t0 + T.T3[a][b][1] * t1 + T.T3[a][b][2] * t2);
    }
    {
      const int m = j >> 2, k = j & 3;
      const double e0f = EF[j * 4], e1f = EF[j * 4 + 1], e2f = EF[j * 4 + 2], e3f = EF[j * 4 + 3];
      const double s0 = T.S[0][m][k], s1 = T.S[1][m][k], s2 = T.S[2][m][k];
#pragma unroll
      for (int r = 0; r < 3; r++) {
        const double g0 = I[0][r] * I[0][0] + I[1][r] * I[1][0] + I[2][r] * I[2][0], g1 = I[0][r] * I[0][1] + I[1][r] * I[1][1] + I[2][r] * I[2][1],
                     g2 = I[0][r] * I[0][2] + I[1][r] * I[1][2] + I[2][r] * I[2][2];
        const double v = g0 * s0 + g1 * s1 + g2 * s2;
        const double w = E[PG_HF + r] * e0f + E[PG_HF + 3 + r] * e1f + E[PG_HF + 6 + r] * e2f + E[PG_HF + 9 + r] * e3f;
        E[PG_SJ + (r * 4 + m) * 4 + k] = dsc * fma(det, v, w);
      }
    }
    __syncwarp();
    // ---- stage 4: K (lane = (m, n)) ------------------------------------------------------------------------------------------------------------------------------
    {
      const int m = j >> 2, n = j & 3;
      double v = 0.0;
#pragma unroll
      for (int f = 0; f < 4; f++) { const int a = nif(f, m), b = nif(f, n); if (a != 15 && b != 15) v += E[PG_FT + (f * 3 + a) * 3 + b]; }
#pragma unroll
      for (int r = 0; r < 3; r++) {
        const double* sj = E + PG_SJ + (r * 4 + m) * 4;
        v -= sj[0] * T.A[r][0][n] + sj[1] * T.A[r][1][n] + sj[2] * T.A[r][2][n] + sj[3] * T.A[r][3][n];
      }
      E[PG_K + j] = v;
    }
    __syncwarp();
    // ---- stage 5: K^-1 by cofactors, lane (i, c) forms Ki[i][c] = cof(c, i) / det K; det K = sum_i K[c][i] cof(c, i) over the four lanes of a column c --------------
    {
      const int i = j >> 2, c = j & 3;
      const int r0 = c <= 0 ? 1 : 0, r1 = c <= 1 ? 2 : 1, r2 = c <= 2 ? 3 : 2;      // rows without c
      const int c0 = i <= 0 ? 1 : 0, c1 = i <= 1 ? 2 : 1, c2 = i <= 2 ? 3 : 2;      // columns without i
      const double* Kp = E + PG_K;
      const double a00 = Kp[r0 * 4 + c0], a01 = Kp[r0 * 4 + c1], a02 = Kp[r0 * 4 + c2], a10 = Kp[r1 * 4 + c0], a11 = Kp[r1 * 4 + c1], a12 = Kp[r1 * 4 + c2],
                   a20 = Kp[r2 * 4 + c0], a21 = Kp[r2 * 4 + c1], a22 = Kp[r2 * 4 + c2];
      double cof = a00 * (a11 * a22 - a12 * a21) - a01 * (a10 * a22 - a12 * a20) + a02 * (a10 * a21 - a11 * a20);
      if ((i + c) & 1) cof = -cof;
      double dK = Kp[c * 4 + i] * cof;
      dK += __shfl_xor_sync(0xffffffffu, dK, 4);
      dK += __shfl_xor_sync(0xffffffffu, dK, 8);
      if (!(fabs(dK) > 1e-300) && live) atomicOr(p.status, 1);
      E[PG_KI + j] = cof / dK;
    }
    __syncwarp();
    // ---- stage 6: one column per lane ----------------------------------------------------------------------------------------------------------------------------
    if (j < 13 && live) {
      const bool rhsCol = j == 12;
      const int c = rhsCol ? 0 : j, fc = c / 3, bcol = c - 3 * fc;
      const double b0 = T.BH[fc][0][bcol], b1 = T.BH[fc][1][bcol], b2 = T.BH[fc][2][bcol], b3 = T.BH[fc][3][bcol];
      double R[4];
      if (!rhsCol) {
        const double cr0 = E[PG_CR + fc * 3], cr1 = E[PG_CR + fc * 3 + 1], cr2 = E[PG_CR + fc * 3 + 2];
#pragma unroll
        for (int m = 0; m < 4; m++) {
          const double2* s0 = reinterpret_cast<const double2*>(E + PG_SJ + m * 4);
          const double2* s1 = reinterpret_cast<const double2*>(E + PG_SJ + (4 + m) * 4);
          const double2* s2 = reinterpret_cast<const double2*>(E + PG_SJ + (8 + m) * 4);
          const double2 p0 = s0[0], p1 = s0[1], q0 = s1[0], q1 = s1[1], w0 = s2[0], w1 = s2[1];
          double v = cr0 * (p0.x * b0 + p0.y * b1 + p1.x * b2 + p1.y * b3);
          v = fma(cr1, q0.x * b0 + q0.y * b1 + q1.x * b2 + q1.y * b3, v);
          v = fma(cr2, w0.x * b0 + w0.y * b1 + w1.x * b2 + w1.y * b3, v);
          const int a = nif(fc, m);
          if (a != 15) v -= E[PG_FT + (fc * 3 + a) * 3 + bcol];      // Sul = -tau mass
          R[m] = v;
        }
      } else {
#pragma unroll
        for (int m = 0; m < 4; m++) R[m] = -E[PG_FU + m];
      }
      double U[4], V[4];
      const double2* Ki2 = reinterpret_cast<const double2*>(E + PG_KI);
      const double2* K2 = reinterpret_cast<const double2*>(E + PG_K);
#pragma unroll
      for (int m = 0; m < 4; m++) { const double2 x = Ki2[2 * m], y = Ki2[2 * m + 1]; U[m] = -(x.x * R[0] + x.y * R[1] + y.x * R[2] + y.y * R[3]); }
#pragma unroll
      for (int m = 0; m < 4; m++) { const double2 x = K2[2 * m], y = K2[2 * m + 1]; V[m] = R[m] + x.x * U[0] + x.y * U[1] + y.x * U[2] + y.y * U[3]; }
#pragma unroll
      for (int m = 0; m < 4; m++) { const double2 x = Ki2[2 * m], y = Ki2[2 * m + 1]; U[m] -= x.x * V[0] + x.y * V[1] + y.x * V[2] + y.y * V[3]; }
      double Q[3][4];
      {
        double Pr[3][4];
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
          for (int m = 0; m < 4; m++) Pr[r][m] = T.A[r][m][0] * U[0] + T.A[r][m][1] * U[1] + T.A[r][m][2] * U[2] + T.A[r][m][3] * U[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
          const double cq = rhsCol ? 0.0 : E[PG_CQ + fc * 3 + d];
#pragma unroll
          for (int m = 0; m < 4; m++) {
            const double bm = m == 0 ? b0 : (m == 1 ? b1 : (m == 2 ? b2 : b3));
            Q[d][m] = fma(cq, bm, -(I[d][0] * Pr[0][m] + I[d][1] * Pr[1][m] + I[d][2] * Pr[2][m]));
          }
        }
      }
      if (!rhsCol) {
        double* const gU = p.U + (size_t)e * 48;
        double* const gQ = p.Q + (size_t)e * 144;
#pragma unroll
        for (int m = 0; m < 4; m++) gU[m * 12 + c] = U[m];
#pragma unroll
        for (int m = 0; m < 4; m++)
#pragma unroll
          for (int d = 0; d < 3; d++) gQ[(m * 3 + d) * 12 + c] = Q[d][m];
      } else {
#pragma unroll
        for (int m = 0; m < 4; m++) p.U0[(size_t)e * 4 + m] = U[m];
#pragma unroll
        for (int m = 0; m < 4; m++)
#pragma unroll
          for (int d = 0; d < 3; d++) p.Q0[(size_t)e * 12 + m * 3 + d] = Q[d][m];
      }
      const int permC = rhsCol ? 0 : EI[PGI_PERM + c];
      double* const gS = p.S ? p.S + (size_t)e * 144 : nullptr;
#pragma unroll
      for (int f = 0; f < 4; f++) {
        const double n0 = E[PG_N + f * 3], n1 = E[PG_N + f * 3 + 1], n2 = E[PG_N + f * 3 + 2], ar = E[PG_AR + f];
        double zq[3], uf[3];
#pragma unroll
        for (int b = 0; b < 3; b++) {
          const int nd = kP1FN[f][b];
          zq[b] = -dsc * (n0 * Q[0][nd] + n1 * Q[1][nd] + n2 * Q[2][nd]);
          uf[b] = U[nd] - ((!rhsCol && f == fc && b == bcol) ? 1.0 : 0.0);
        }
        const int bcf = EI[PGI_BC + f], Ff = EI[PGI_F + f];
        const bool inter = EI[PGI_IN + f] != 0;
        double* const blk = p.vals + RS[f] + (long long)EI[PGI_POS + f * 4 + fc] * 9 + permC;
#pragma unroll
        for (int a = 0; a < 3; a++) {
          const double* ft = E + PG_FT + (f * 3 + a) * 3;
          double s2 = ft[0] * uf[0] + ft[1] * uf[1] + ft[2] * uf[2] + ar * (T.MF[a][0] * zq[0] + T.MF[a][1] * zq[1] + T.MF[a][2] * zq[2]);
          const int r = f * 3 + a, pr = EI[PGI_PERM + r];
          if (!rhsCol) {
            if (bcf == 1) s2 = (r == c) ? 1.0 : 0.0;                                          // DirichletModel row (Set)
            else if (bcf == 2) s2 = (f == fc) ? ar * T.MF[a][bcol] : 0.0;                      // IntegratedDirichletModel row: face mass
            if (gS) gS[r + 12 * c] = s2;
            double* dst = blk + pr * 3;
            if (f == fc && inter) atomicAdd(dst, s2); else *dst = s2;
          } else {
            double s0 = -s2;
            if (bcf == 1) s0 = p.dirichlet[(size_t)Ff * 3 + a];
            else if (bcf == 2) {
              s0 = 0.0;
#pragma unroll
              for (int b = 0; b < 3; b++) s0 = fma(ar * T.MF[a][b], p.dirichlet[(size_t)Ff * 3 + b], s0);
            }
            if (p.S0) p.S0[(size_t)e * 12 + r] = s0;
            double* dst = p.rhs + (size_t)Ff * 3 + pr;
            if (inter) atomicAdd(dst, s0); else *dst = s0;
          }
        }
      }
    }
    if (e0 + (long long)gridDim.x * kPGElems < p.eEnd) loadTau();   // the next element's face ids have arrived by now
    __syncwarp();
  }
}

template <int MINB>
inline cudaError_t launch_p1g_t(const AsmParams& p, int nSM, cudaStream_t st) {
  const size_t bytes = (size_t)(((kPGTabAll + 1) & ~1) + kPGElems * PG_STRIDE) * sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(hdg_p1g_kernel<MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return e;
  int perSM = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, hdg_p1g_kernel<MINB>, kPGThreads, bytes);
  if (perSM < 1) perSM = 1;
  long long grid = (long long)nSM * perSM;
  const long long need = ((long long)(p.eEnd - p.eBegin) + kPGElems - 1) / kPGElems;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  hdg_p1g_kernel<MINB><<<(int)grid, kPGThreads, bytes, st>>>(p);
  return cudaGetLastError();
}
inline cudaError_t launch_p1g(const AsmParams& p, int nSM, cudaStream_t st) {
  const int minb = getenv("HFX_P1_MINB") ? atoi(getenv("HFX_P1_MINB")) : 2;
  return minb == 3 ? launch_p1g_t<3>(p, nSM, st) : (minb == 4 ? launch_p1g_t<4>(p, nSM, st) : launch_p1g_t<2>(p, nSM, st));
}

}  // namespace hfx
