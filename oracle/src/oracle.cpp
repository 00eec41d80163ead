// ORACLE -- test infrastructure only (see oracle.h).  CPU restatement of the reference's HDG element path.
// Parity status: pinned by the reference's known-answer tests restated in tests/test_oracle_*.py
// (operator identities, analytic Jacobians, reference normals, TestHDGSolver constant solution, regression L2 ceilings).
#include "oracle.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <thread>
#include <vector>

namespace {

struct Mat {  // column-major dense matrix
  int r = 0, c = 0;
  std::vector<double> a;
  Mat() {}
  Mat(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_, 0.0) {}
  double& operator()(int i, int j) { return a[(size_t)j * r + i]; }
  double operator()(int i, int j) const { return a[(size_t)j * r + i]; }
};

struct Sizes {
  int dim, nN, nNf, nFc, nIP, nIPf, nDOF, u, q, l, n, t, nJ;
  Sizes(const orc_refel* re, int nDOF_) {
    dim = re->dim; nN = re->nN; nNf = re->nNf; nFc = re->nFc; nIP = re->nIP; nIPf = re->nIPf; nDOF = nDOF_;
    u = nN * nDOF; q = u * dim; l = nFc * nNf * nDOF; n = u + q + l; t = nNf * nDOF; nJ = nIP + nFc * nIPf;
  }
};

// ---- small dense helpers -------------------------------------------------------------------
double det_small(const double* M, int d) {  // row-major d x d
  if (d == 1) return M[0];
  if (d == 2) return M[0] * M[3] - M[1] * M[2];
  return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}
void inv_small(const double* M, int d, double* R) {  // row-major d x d (cofactor formula, as Eigen does for d <= 4)
  double dt = det_small(M, d);
  if (d == 1) { R[0] = 1.0 / M[0]; return; }
  if (d == 2) {
    double id = 1.0 / dt;
    R[0] = M[3] * id; R[1] = -M[1] * id; R[2] = -M[2] * id; R[3] = M[0] * id;
    return;
  }
  double id = 1.0 / dt;
  R[0] = (M[4] * M[8] - M[5] * M[7]) * id; R[1] = (M[2] * M[7] - M[1] * M[8]) * id; R[2] = (M[1] * M[5] - M[2] * M[4]) * id;
  R[3] = (M[5] * M[6] - M[3] * M[8]) * id; R[4] = (M[0] * M[8] - M[2] * M[6]) * id; R[5] = (M[2] * M[3] - M[0] * M[5]) * id;
  R[6] = (M[3] * M[7] - M[4] * M[6]) * id; R[7] = (M[1] * M[6] - M[0] * M[7]) * id; R[8] = (M[0] * M[4] - M[1] * M[3]) * id;
}

// Geometry of one element. Operator.cpp:14-84 + HDGModel.cpp:53-85 + HDGBase.cpp:34-65.
// jac[k]: row-major [r][m] (stride dim); invjac[k]: row-major [m][r] (stride dim) i.e. invJ(m, r).
struct Geom {
  std::vector<double> jac, invjac, dV, normals;
};

void jacobians(const double* pts, int nPts, const double* dshape, int nIPs, int dimRef, int dimMesh, int stride, double* jac) {
  // Operator::calcJacobians, Operator.cpp:14-39:  J_ip[r][m] = sum_i dphi_i/dxi_r(ip) x_i[m]
  for (int ip = 0; ip < nIPs; ip++) {
    double* J = jac + (size_t)ip * stride * stride;
    for (int k = 0; k < stride * stride; k++) J[k] = 0.0;
  }
  for (int i = 0; i < nPts; i++)
    for (int ip = 0; ip < nIPs; ip++) {
      double* J = jac + (size_t)ip * stride * stride;
      const double* d = dshape + ((size_t)ip * nPts + i) * dimRef;
      for (int r = 0; r < dimRef; r++)
        for (int m = 0; m < dimMesh; m++) J[r * stride + m] += d[r] * pts[i * dimMesh + m];
    }
}

void geometry(const orc_refel* re, const double* nodes, Geom& g) {
  const int dim = re->dim, nIP = re->nIP, nIPf = re->nIPf, nFc = re->nFc, nNf = re->nNf, nN = re->nN;
  const int nJ = nIP + nFc * nIPf, dd = dim * dim;
  g.jac.assign((size_t)nJ * dd, 0.0); g.invjac.assign((size_t)nJ * dd, 0.0); g.dV.assign(nJ, 0.0);
  g.normals.assign((size_t)nFc * nIPf * dim, 0.0);
  jacobians(nodes, nN, re->dshape, nIP, dim, dim, dim, g.jac.data());
  for (int ip = 0; ip < nIP; ip++) {
    const double* J = &g.jac[(size_t)ip * dd];
    g.dV[ip] = re->w[ip] * det_small(J, dim);   // calcDetJacobians :59-75, calcMeasure :78-84
    inv_small(J, dim, &g.invjac[(size_t)ip * dd]);  // calcInvJacobians :41-57 (square)
  }
  std::vector<double> fpts((size_t)nNf * dim);
  const int dr = dim - 1;
  for (int f = 0; f < nFc; f++) {
    for (int i = 0; i < nNf; i++)
      for (int m = 0; m < dim; m++) fpts[i * dim + m] = nodes[re->faceNodes[f * nNf + i] * dim + m];
    double* Jf = &g.jac[(size_t)(nIP + f * nIPf) * dd];
    jacobians(fpts.data(), nNf, re->fdshape, nIPf, dr, dim, dim, Jf);
    // outward orientation test vector HDGBase.cpp:43-53
    int v0 = re->faceNodes[f * nNf + 0], vn = -1;
    for (int k = 0; k < nN && vn < 0; k++) {
      bool in = false;
      for (int i = 0; i < nNf; i++) in = in || (re->faceNodes[f * nNf + i] == k);
      if (!in) vn = k;
    }
    double tv[3] = {0, 0, 0};
    for (int m = 0; m < dim; m++) tv[m] = nodes[vn * dim + m] - nodes[v0 * dim + m];
    for (int ip = 0; ip < nIPf; ip++) {
      const double* J = Jf + (size_t)ip * dd;  // rows r < dr
      double G[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};               // J J^T (dr x dr)
      for (int a = 0; a < dr; a++)
        for (int b = 0; b < dr; b++) {
          double s = 0;
          for (int m = 0; m < dim; m++) s += J[a * dim + m] * J[b * dim + m];
          G[a * dr + b] = s;
        }
      double dG = det_small(G, dr);
      int k = nIP + f * nIPf + ip;
      g.dV[k] = re->fw[ip] * std::sqrt(dG);    // sqrt(det(J J^T))
      double Gi[9];
      inv_small(G, dr, Gi);
      double* P = &g.invjac[(size_t)k * dd];   // pseudo-inverse J^T (J J^T)^-1 : dim x dr, stored [m][r] stride dim
      for (int m = 0; m < dim; m++)
        for (int r = 0; r < dr; r++) {
          double s = 0;
          for (int a = 0; a < dr; a++) s += J[a * dim + m] * Gi[a * dr + r];
          P[m * dim + r] = s;
        }
      // normal = kernel of the face Jacobian, HDGBase.cpp:54-62
      double nv[3] = {0, 0, 0};
      if (dim == 2) { nv[0] = -J[1]; nv[1] = J[0]; }
      else {
        const double* a = J; const double* b = J + dim;
        nv[0] = a[1] * b[2] - a[2] * b[1]; nv[1] = a[2] * b[0] - a[0] * b[2]; nv[2] = a[0] * b[1] - a[1] * b[0];
      }
      double nrm = 0;
      for (int m = 0; m < dim; m++) nrm += nv[m] * nv[m];
      nrm = std::sqrt(nrm);
      double prod = 0;
      for (int m = 0; m < dim; m++) { nv[m] /= nrm; prod += tv[m] * nv[m]; }
      if (prod > 0) for (int m = 0; m < dim; m++) nv[m] = -nv[m];
      for (int m = 0; m < dim; m++) g.normals[(size_t)(f * nIPf + ip) * dim + m] = nv[m];
    }
  }
}

// Mass::assemble, Mass.cpp:5-38 (nDOF = 1 part)
void mass(int nN, int nIP, const double* shape, const double* dV, Mat& M) {
  M = Mat(nN, nN);
  for (int j = 0; j < nN; j++)
    for (int k = j; k < nN; k++) {
      double s = 0;
      for (int ip = 0; ip < nIP; ip++) s += shape[ip * nN + j] * shape[ip * nN + k] * dV[ip];
      M(j, k) = s; M(k, j) = s;
    }
}

// HDGBase::setTau :18-32 + HDGBase::assemble :67-158
void op_base(const orc_refel* re, const Sizes& z, const Geom& g, const double* tau, Mat& A) {
  const int nD = z.nDOF, dim = z.dim, sT = nD * nD;
  const int sQ = z.u, sL = z.u + z.q;
  std::vector<double> taus((size_t)z.nFc * z.nIPf * sT, 0.0);
  for (int f = 0; f < z.nFc; f++)
    for (int ip = 0; ip < z.nIPf; ip++)
      for (int c = 0; c < sT; c++) {
        double s = 0;
        for (int j = 0; j < z.nNf; j++) s += tau[(size_t)(f * z.nNf + j) * sT + c] * re->fshape[ip * z.nNf + j];
        taus[(size_t)(f * z.nIPf + ip) * sT + c] = s;
      }
  for (int ip = 0; ip < z.nIPf; ip++) {
    const double* sh = re->fshape + (size_t)ip * z.nNf;
    for (int f = 0; f < z.nFc; f++) {
      int off = f * z.nIPf + ip;
      const int* fn = re->faceNodes + f * z.nNf;
      double dv = g.dV[z.nIP + off];
      for (int iN = 0; iN < z.nNf; iN++)
        for (int nd = 0; nd < nD; nd++)
          for (int jN = 0; jN < z.nNf; jN++) {
            double ss = sh[iN] * sh[jN];
            for (int md = 0; md < nD; md++) {
              double val = (dv * taus[(size_t)off * sT + md * nD + nd]) * ss;
              A(sL + (f * z.nNf + iN) * nD + nd, sL + (f * z.nNf + jN) * nD + md) -= val;  // Sll
              A(sL + (f * z.nNf + iN) * nD + nd, fn[jN] * nD + md) += val;                   // Slu
              A(fn[iN] * nD + nd, fn[jN] * nD + md) += val;                                  // Suu
              A(fn[iN] * nD + nd, sL + (f * z.nNf + jN) * nD + md) -= val;                   // Sul
            }
            for (int d = 0; d < dim; d++)                                                    // Sql
              A(sQ + (fn[iN] * dim + d) * nD + nd, sL + (f * z.nNf + jN) * nD + nd) -= (dv * g.normals[(size_t)off * dim + d]) * ss;
          }
    }
  }
  for (int ip = 0; ip < z.nIP; ip++) {
    const double* sh = re->shape + (size_t)ip * z.nN;
    const double* iJ = &g.invjac[(size_t)ip * dim * dim];
    for (int iN = 0; iN < z.nN; iN++) {
      const double* dp = re->dshape + ((size_t)ip * z.nN + iN) * dim;
      double vm[3];
      for (int d = 0; d < dim; d++) {
        double s = 0;
        for (int r = 0; r < dim; r++) s += iJ[d * dim + r] * dp[r];
        vm[d] = s * g.dV[ip];
      }
      for (int d = 0; d < dim; d++)
        for (int nd = 0; nd < nD; nd++)
          for (int jN = 0; jN < z.nN; jN++) {
            A(sQ + (iN * dim + d) * nD + nd, jN * nD + nd) += vm[d] * sh[jN];                               // Squ
            A(sQ + (iN * dim + d) * nD + nd, sQ + (jN * dim + d) * nD + nd) += g.dV[ip] * (sh[iN] * sh[jN]); // Sqq
          }
    }
  }
}

// HDGDiffusion::setDiffusionTensor :31-72 + assemble :74-145.  D col-major per node; diffComps 0 => identity.
void op_diffusion(const orc_refel* re, const Sizes& z, const Geom& g, const double* diff, int diffComps, Mat& A) {
  const int nD = z.nDOF, dim = z.dim, dd = dim * dim;
  std::vector<double> Ds((size_t)z.nJ * dd, 0.0);
  if (diffComps == 0 || diff == nullptr) {
    for (int k = 0; k < z.nJ; k++)
      for (int d = 0; d < dim; d++) Ds[(size_t)k * dd + d * dim + d] = 1.0;
  } else {
    std::vector<double> nodeD((size_t)z.nN * dd, 0.0);
    for (int i = 0; i < z.nN; i++) {
      if (diffComps == 1) for (int d = 0; d < dim; d++) nodeD[(size_t)i * dd + d * dim + d] = diff[i];
      else for (int c = 0; c < dd; c++) nodeD[(size_t)i * dd + c] = diff[(size_t)i * dd + c];
    }
    for (int ip = 0; ip < z.nIP; ip++)
      for (int c = 0; c < dd; c++) {
        double s = 0;
        for (int i = 0; i < z.nN; i++) s += re->shape[ip * z.nN + i] * nodeD[(size_t)i * dd + c];
        Ds[(size_t)ip * dd + c] = s;
      }
    for (int f = 0; f < z.nFc; f++)
      for (int ip = 0; ip < z.nIPf; ip++)
        for (int c = 0; c < dd; c++) {
          double s = 0;
          for (int i = 0; i < z.nNf; i++) s += re->fshape[ip * z.nNf + i] * nodeD[(size_t)re->faceNodes[f * z.nNf + i] * dd + c];
          Ds[(size_t)(z.nIP + f * z.nIPf + ip) * dd + c] = s;
        }
  }
  const int lenU = z.u, lenQ = z.q;
  for (int f = 0; f < z.nFc; f++) {
    const int* fn = re->faceNodes + f * z.nNf;
    for (int ip = 0; ip < z.nIPf; ip++) {
      const double* sh = re->fshape + (size_t)ip * z.nNf;
      int off = f * z.nIPf + ip;
      const double* D = &Ds[(size_t)(z.nIP + off) * dd];
      double bv[3];
      for (int a = 0; a < dim; a++) {
        double s = 0;
        for (int b = 0; b < dim; b++) s += D[b * dim + a] * g.normals[(size_t)off * dim + b];
        bv[a] = s * g.dV[z.nIP + off];
      }
      for (int iN = 0; iN < z.nNf; iN++)
        for (int nd = 0; nd < nD; nd++)
          for (int jN = 0; jN < z.nNf; jN++)
            for (int d = 0; d < dim; d++) {
              double v = (bv[d] * sh[iN]) * sh[jN];
              A(lenU + lenQ + (f * z.nNf + iN) * nD + nd, lenU + (fn[jN] * dim + d) * nD + nd) -= v;
              A(fn[iN] * nD + nd, lenU + (fn[jN] * dim + d) * nD + nd) -= v;
            }
    }
  }
  for (int ip = 0; ip < z.nIP; ip++) {
    const double* sh = re->shape + (size_t)ip * z.nN;
    const double* iJ = &g.invjac[(size_t)ip * dd];
    const double* D = &Ds[(size_t)ip * dd];
    for (int iN = 0; iN < z.nN; iN++) {
      const double* dp = re->dshape + ((size_t)ip * z.nN + iN) * dim;
      double gr[3], bv[3];
      for (int d = 0; d < dim; d++) {
        double s = 0;
        for (int r = 0; r < dim; r++) s += iJ[d * dim + r] * dp[r];
        gr[d] = s;
      }
      for (int a = 0; a < dim; a++) {
        double s = 0;
        for (int b = 0; b < dim; b++) s += D[b * dim + a] * gr[b];
        bv[a] = s * g.dV[ip];
      }
      for (int nd = 0; nd < nD; nd++)
        for (int jN = 0; jN < z.nN; jN++)
          for (int d = 0; d < dim; d++) A(iN * nD + nd, lenU + (jN * dim + d) * nD + nd) += bv[d] * sh[jN];
    }
  }
}

// HDGOperator::multiplyDOFs, HDGOperator.cpp:17-28
void multiply_dofs(Mat& A, int nDOF) {
  if (nDOF <= 1) return;
  int nn = A.c / nDOF;
  Mat buf(nn, nn);
  for (int i = 0; i < nn; i++) for (int j = 0; j < nn; j++) { buf(i, j) = A(i, j); A(i, j) = 0.0; }
  for (int i = 0; i < nn; i++)
    for (int j = 0; j < nn; j++)
      for (int k = 0; k < nDOF; k++) A(i * nDOF + k, j * nDOF + k) = buf(i, j);
}

// HDGConvection::setVelocity :31-58 + assemble :60-104 (+ Convection.cpp:5-49, Mass.cpp)
void op_convection(const orc_refel* re, const Sizes& z, const Geom& g, const double* vel, Mat& A) {
  const int dim = z.dim, nN = z.nN, nNf = z.nNf;
  Mat C(nN, nN);
  for (int ip = 0; ip < z.nIP; ip++) {
    const double* sh = re->shape + (size_t)ip * nN;
    double v[3] = {0, 0, 0};
    for (int i = 0; i < nN; i++) for (int d = 0; d < dim; d++) v[d] += vel[i * dim + d] * sh[i];
    const double* iJ = &g.invjac[(size_t)ip * dim * dim];
    double lm[3];
    for (int r = 0; r < dim; r++) {
      double s = 0;
      for (int d = 0; d < dim; d++) s += v[d] * iJ[d * dim + r];
      lm[r] = g.dV[ip] * s;
    }
    for (int k = 0; k < nN; k++)
      for (int l2 = 0; l2 < nN; l2++) {
        const double* dp = re->dshape + ((size_t)ip * nN + l2) * dim;
        double s = 0;
        for (int r = 0; r < dim; r++) s += lm[r] * dp[r];
        C(k, l2) += s * sh[k];
      }
  }
  for (int i = 0; i < nN; i++) for (int j = 0; j < nN; j++) A(i, j) -= C(j, i);
  std::vector<double> lm(z.nIPf);
  for (int f = 0; f < z.nFc; f++) {
    const int* fn = re->faceNodes + f * nNf;
    for (int ip = 0; ip < z.nIPf; ip++) {
      const double* sh = re->fshape + (size_t)ip * nNf;
      double v[3] = {0, 0, 0};
      for (int i = 0; i < nNf; i++) for (int d = 0; d < dim; d++) v[d] += vel[fn[i] * dim + d] * sh[i];
      double s = 0;
      for (int d = 0; d < dim; d++) s += v[d] * g.normals[(size_t)(f * z.nIPf + ip) * dim + d];
      lm[ip] = g.dV[z.nIP + f * z.nIPf + ip] * s;
    }
    Mat Mf;
    mass(nNf, z.nIPf, re->fshape, lm.data(), Mf);
    int off = nN * (dim + 1) + f * nNf;
    for (int j = 0; j < nNf; j++)
      for (int k = 0; k < nNf; k++) { A(fn[k], off + j) += Mf(k, j); A(off + k, off + j) += Mf(k, j); }
  }
  multiply_dofs(A, z.nDOF);
}

// HDGUNabU::setSolution/setTrace/assemble, HDGUNabU.cpp:27-191 (nDOF == dim)
void op_unabu(const orc_refel* re, const Sizes& z, const Geom& g, const double* sol, const double* trace, Mat& A, std::vector<double>& rhs) {
  const int dim = z.dim, nD = z.nDOF, nN = z.nN, nNf = z.nNf;
  const int lenU = z.u, sL = z.u + z.q, lenL = z.l;
  for (int f = 0; f < z.nFc; f++) {
    const int* fn = re->faceNodes + f * nNf;
    for (int ip = 0; ip < z.nIPf; ip++) {
      const double* sh = re->fshape + (size_t)ip * nNf;
      int fo = f * z.nIPf + ip;
      double tr[3] = {0, 0, 0}, fs[3] = {0, 0, 0};
      for (int i = 0; i < nNf; i++)
        for (int d = 0; d < dim; d++) { tr[d] += trace[(f * nNf + i) * dim + d] * sh[i]; fs[d] += sol[fn[i] * dim + d] * sh[i]; }
      const double* nv = &g.normals[(size_t)fo * dim];
      double dv = g.dV[z.nIP + fo];
      double tdn = 0;
      for (int d = 0; d < dim; d++) tdn += tr[d] * nv[d];
      tdn *= dv;
      for (int iN = 0; iN < nNf; iN++)
        for (int nd = 0; nd < nD; nd++)
          for (int jN = 0; jN < nNf; jN++) {
            A(sL + (f * nNf + iN) * nD + nd, sL + (f * nNf + jN) * nD + nd) += tdn * sh[jN] * sh[iN];
            for (int md = 0; md < nD; md++)
              A(sL + (f * nNf + iN) * nD + nd, sL + (f * nNf + jN) * nD + md) += fs[nd] * sh[jN] * sh[iN] * nv[md] * dv;
          }
    }
  }
  for (int f = 0; f < z.nFc; f++) {
    const int* fn = re->faceNodes + f * nNf;
    for (int iN = 0; iN < nNf; iN++)
      for (int nd = 0; nd < nD; nd++)
        for (int c = 0; c < lenL; c++) A(fn[iN] * nD + nd, sL + c) += A(sL + (f * nNf + iN) * nD + nd, sL + c);
  }
  Mat gm(dim, nN);
  for (int ip = 0; ip < z.nIP; ip++) {
    const double* sh = re->shape + (size_t)ip * nN;
    const double* iJ = &g.invjac[(size_t)ip * dim * dim];
    double sip[3] = {0, 0, 0};
    for (int i = 0; i < nN; i++) for (int d = 0; d < dim; d++) sip[d] += sol[i * dim + d] * sh[i];
    double gs[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // gradSol(dim x nDOF)
    for (int iN = 0; iN < nN; iN++) {
      const double* dp = re->dshape + ((size_t)ip * nN + iN) * dim;
      for (int d = 0; d < dim; d++) {
        double s = 0;
        for (int r = 0; r < dim; r++) s += iJ[d * dim + r] * dp[r];
        gm(d, iN) = s;
      }
      for (int d = 0; d < dim; d++) for (int k = 0; k < nD; k++) gs[d * nD + k] += gm(d, iN) * sol[iN * dim + k];
    }
    double divS = 0;
    for (int d = 0; d < dim; d++) divS += gs[d * nD + d];
    for (int iN = 0; iN < nN; iN++) {
      double sg = 0;
      for (int d = 0; d < dim; d++) sg += sip[d] * gm(d, iN);
      for (int nd = 0; nd < nD; nd++)
        for (int jN = 0; jN < nN; jN++) {
          A(iN * nD + nd, jN * nD + nd) -= (divS * sh[iN] + sg) * g.dV[ip] * sh[jN];
          for (int md = 0; md < nD; md++)
            A(iN * nD + nd, jN * nD + md) -= sip[nd] * (gm(md, jN) * sh[iN] + gm(md, iN) * sh[jN]) * g.dV[ip];
        }
    }
  }
  rhs.assign(z.n, 0.0);
  for (int i = 0; i < lenU; i++) {
    double s = 0;
    for (int j = 0; j < lenU; j++) s += A(i, j) * (sol[j] / 2.0);
    for (int j = 0; j < lenL; j++) s += A(i, sL + j) * (trace[j] / 2.0);
    rhs[i] += s;
  }
  for (int i = 0; i < lenL; i++) {
    double s = 0;
    for (int j = 0; j < lenU; j++) s += A(sL + i, j) * (sol[j] / 2.0);
    for (int j = 0; j < lenL; j++) s += A(sL + i, sL + j) * (trace[j] / 2.0);
    rhs[sL + i] += s;
  }
}

// Model::compute for the four HDG models + HDGModel::compute time-scheme hook (HDGModel.cpp:35-51)
void local_system(const orc_refel* re, const orc_model* md, const orc_elfields* f, Mat& A, std::vector<double>& F) {
  Sizes z(re, md->nDOF);
  Geom g;
  geometry(re, f->nodes, g);
  A = Mat(z.n, z.n);
  F.assign(z.n, 0.0);
  op_base(re, z, g, f->tau, A);
  if (md->opmask & ORC_OP_UNABU) {       // HDGBurgersModel.cpp:87-124
    Mat B(z.n, z.n);
    std::vector<double> r;
    op_unabu(re, z, g, f->bufSol, f->trace, B, r);
    for (size_t k = 0; k < A.a.size(); k++) A.a[k] += B.a[k];
    F = r;
  }
  if (md->opmask & ORC_OP_CONVECTION) {  // HDGConvectionDiffusionReactionSource.cpp:78-83
    Mat B(z.n, z.n);
    op_convection(re, z, g, f->vel, B);
    for (size_t k = 0; k < A.a.size(); k++) A.a[k] += B.a[k];
  }
  if (md->opmask & ORC_OP_DIFFUSION) {
    Mat B(z.n, z.n);
    op_diffusion(re, z, g, f->diff, md->diffComps, B);
    for (size_t k = 0; k < A.a.size(); k++) A.a[k] += B.a[k];
  }
  if ((md->opmask & ORC_OP_REACTION) && f->reacIP) {  // Reaction.cpp:24-36 -> Mass with r*dV, multiplyDOFs, added to Suu
    std::vector<double> lm(z.nIP);
    for (int ip = 0; ip < z.nIP; ip++) lm[ip] = f->reacIP[ip] * g.dV[ip];
    Mat M;
    mass(z.nN, z.nIP, re->shape, lm.data(), M);
    for (int i = 0; i < z.nN; i++) for (int j = 0; j < z.nN; j++) for (int k = 0; k < z.nDOF; k++) A(i * z.nDOF + k, j * z.nDOF + k) += M(i, j);
  }
  if ((md->opmask & ORC_OP_SOURCE) && f->srcIP) {     // Source.cpp:24-48
    int nSrc = (md->opmask & ORC_OP_UNABU) ? z.dim : 1;
    for (int c = 0; c < nSrc; c++)
      for (int i = 0; i < z.nN; i++) {
        double s = 0;
        for (int ip = 0; ip < z.nIP; ip++) s += re->shape[ip * z.nN + i] * (f->srcIP[c * z.nIP + ip] * g.dV[ip]);
        if (md->opmask & ORC_OP_UNABU) F[i * z.dim + c] += s;   // HDGBurgersModel.cpp:112-122
        else F[i] = s;                                            // segment(0, nN) = source
      }
  }
  if (md->timeScheme != ORC_TS_NONE) {
    Mat M1;
    mass(z.nN, z.nIP, re->shape, g.dV.data(), M1);
    Mat M(z.u, z.u);
    for (int i = 0; i < z.nN; i++) for (int j = 0; j < z.nN; j++) for (int k = 0; k < z.nDOF; k++) M(i * z.nDOF + k, j * z.nDOF + k) = M1(i, j);
    const int u = z.u, n = z.n;
    if (md->timeScheme == ORC_TS_EULER_IMPLICIT || md->timeScheme == ORC_TS_EULER_EXPLICIT) {   // Euler.cpp:28-29: Su *= dt; Fu *= dt
      for (int i = 0; i < u; i++) {
        for (int j = 0; j < n; j++) A(i, j) *= md->dt;
        F[i] *= md->dt;
      }
    }
    if (md->timeScheme == ORC_TS_EULER_IMPLICIT) {          // Euler.cpp:30-32
      for (int i = 0; i < u; i++) {
        double s = 0;
        for (int j = 0; j < u; j++) { A(i, j) += M(i, j); s += M(i, j) * f->solOld[j]; }
        F[i] += s;
      }
    } else if (md->timeScheme == ORC_TS_EULER_EXPLICIT) {   // Euler.cpp:33-36
      for (int i = 0; i < u; i++) {
        double s = 0;
        for (int j = 0; j < u; j++) s += (M(i, j) - A(i, j)) * f->solOld[j];
        F[i] += s;
      }
      for (int i = 0; i < u; i++) for (int j = 0; j < u; j++) A(i, j) = M(i, j);
    } else {                                                 // RungeKutta::apply, RungeKutta.cpp:90-143 (aux = {Flux, Trace})
      std::vector<double> uj(n, 0.0), ut(n, 0.0);
      for (int s = 0; s < md->rkStage; s++) {
        double a = md->rkRow[s];
        for (int j = 0; j < u; j++) uj[j] += a * f->rkSol[(size_t)s * u + j];
        for (int j = 0; j < z.q; j++) uj[u + j] += a * f->rkFlux[(size_t)s * z.q + j];
        for (int j = 0; j < z.l; j++) uj[u + z.q + j] += a * f->rkTrace[(size_t)s * z.l + j];
      }
      for (int j = 0; j < n; j++) uj[j] *= md->dt;
      for (int j = 0; j < u; j++) ut[j] = f->solOld[j];
      for (int j = 0; j < z.q; j++) ut[u + j] = f->fluxOld[j];
      for (int j = 0; j < z.l; j++) ut[u + z.q + j] = f->traceOld[j];
      for (int j = 0; j < n; j++) uj[j] += ut[j];
      for (int i = 0; i < u; i++) {
        for (int j = 0; j < n; j++) A(i, j) *= md->dt;
        F[i] *= md->dt;
        double s = 0;
        for (int j = 0; j < n; j++) s += A(i, j) * uj[j];
        F[i] -= s;
      }
      double ass = md->rkRow[md->rkStage];
      for (int i = 0; i < u; i++) {
        for (int j = 0; j < n; j++) A(i, j) *= ass;
        for (int j = 0; j < u; j++) A(i, j) += M(i, j);
        double s = 0;
        for (int j = 0; j < n; j++) s += A(i, j) * ut[j];
        F[i] += s;
      }
    }
  }
}

// ---- dense factorizations -------------------------------------------------------------------
// Unpivoted Householder QR (what Eigen::HouseholderQR does) of a square matrix, with solve for multiple RHS.
struct HQR {
  int n; Mat qr; std::vector<double> tau;
  void compute(const Mat& A) {
    n = A.r; qr = A; tau.assign(n, 0.0);
    for (int k = 0; k < n; k++) {
      double nrm2 = 0;
      for (int i = k + 1; i < n; i++) nrm2 += qr(i, k) * qr(i, k);
      double c0 = qr(k, k);
      if (nrm2 == 0.0) { tau[k] = 0.0; continue; }
      double beta = std::sqrt(c0 * c0 + nrm2);
      if (c0 >= 0) beta = -beta;
      for (int i = k + 1; i < n; i++) qr(i, k) /= (c0 - beta);
      tau[k] = (beta - c0) / beta;
      qr(k, k) = beta;
      for (int j = k + 1; j < n; j++) {  // apply H = I - tau v v^T to trailing columns
        double s = qr(k, j);
        for (int i = k + 1; i < n; i++) s += qr(i, k) * qr(i, j);
        s *= tau[k];
        qr(k, j) -= s;
        for (int i = k + 1; i < n; i++) qr(i, j) -= s * qr(i, k);
      }
    }
  }
  void solve(Mat& B) const {  // in place
    for (int c = 0; c < B.c; c++) {
      for (int k = 0; k < n; k++) {
        if (tau[k] == 0.0) continue;
        double s = B(k, c);
        for (int i = k + 1; i < n; i++) s += qr(i, k) * B(i, c);
        s *= tau[k];
        B(k, c) -= s;
        for (int i = k + 1; i < n; i++) B(i, c) -= s * qr(i, k);
      }
      for (int i = n - 1; i >= 0; i--) {
        double s = B(i, c);
        for (int j = i + 1; j < n; j++) s -= qr(i, j) * B(j, c);
        B(i, c) = s / qr(i, i);
      }
    }
  }
};
struct PLU {
  int n; Mat lu; std::vector<int> piv;
  void compute(const Mat& A) {
    n = A.r; lu = A; piv.resize(n);
    for (int k = 0; k < n; k++) {
      int p = k; double mx = std::fabs(lu(k, k));
      for (int i = k + 1; i < n; i++) if (std::fabs(lu(i, k)) > mx) { mx = std::fabs(lu(i, k)); p = i; }
      piv[k] = p;
      if (p != k) for (int j = 0; j < n; j++) std::swap(lu(k, j), lu(p, j));
      double d = 1.0 / lu(k, k);
      for (int i = k + 1; i < n; i++) lu(i, k) *= d;
      for (int j = k + 1; j < n; j++) {
        double s = lu(k, j);
        for (int i = k + 1; i < n; i++) lu(i, j) -= lu(i, k) * s;
      }
    }
  }
  void solve(Mat& B) const {
    for (int c = 0; c < B.c; c++) {
      for (int k = 0; k < n; k++) if (piv[k] != k) std::swap(B(k, c), B(piv[k], c));
      for (int k = 0; k < n; k++) { double s = B(k, c); for (int i = k + 1; i < n; i++) B(i, c) -= lu(i, k) * s; }
      for (int i = n - 1; i >= 0; i--) {
        double s = B(i, c);
        for (int j = i + 1; j < n; j++) s -= lu(i, j) * B(j, c);
        B(i, c) = s / lu(i, i);
      }
    }
  }
};

Mat block(const Mat& A, int r0, int c0, int nr, int nc) {
  Mat B(nr, nc);
  for (int j = 0; j < nc; j++) for (int i = 0; i < nr; i++) B(i, j) = A(r0 + i, c0 + j);
  return B;
}
void gemm_acc(const Mat& A, const Mat& B, Mat& C, double alpha) {  // C += alpha A B
  for (int j = 0; j < B.c; j++)
    for (int k = 0; k < A.c; k++) {
      double b = alpha * B(k, j);
      if (b == 0.0) continue;
      for (int i = 0; i < A.r; i++) C(i, j) += A(i, k) * b;
    }
}

template <class Fac>
void condense_t(int u, int q, int l, const Mat& A, const double* F, double* Uo, double* Qo, double* So, double* U0o, double* Q0o, double* S0o) {
  // HDGSolver.cpp:331-348
  const int sU = 0, sQ = u, sL = u + q;
  Fac fq; fq.compute(block(A, sQ, sQ, q, q));
  Mat iSqu = block(A, sQ, sU, q, u); fq.solve(iSqu);          // invSqqSqu
  Mat iSql = block(A, sQ, sL, q, l); fq.solve(iSql);          // invSqqSql
  Mat Suq = block(A, sU, sQ, u, q);
  Mat K = block(A, sU, sU, u, u); gemm_acc(Suq, iSqu, K, -1.0);
  Fac fk; fk.compute(K);
  Mat R = block(A, sU, sL, u, l); gemm_acc(Suq, iSql, R, -1.0);
  fk.solve(R);
  Mat U(u, l);
  for (size_t k = 0; k < U.a.size(); k++) U.a[k] = -R.a[k];
  Mat U0(u, 1);
  for (int i = 0; i < u; i++) U0(i, 0) = F[sU + i];
  fk.solve(U0);
  Mat Q(q, l);
  gemm_acc(iSqu, U, Q, -1.0);
  for (size_t k = 0; k < Q.a.size(); k++) Q.a[k] -= iSql.a[k];
  Mat Q0(q, 1);
  gemm_acc(iSqu, U0, Q0, -1.0);
  Mat Slu = block(A, sL, sU, l, u), Slq = block(A, sL, sQ, l, q);
  Mat S = block(A, sL, sL, l, l);
  // locS = Slu*U + Slq*Q + Sll  (sum order as written)
  Mat T(l, l);
  gemm_acc(Slu, U, T, 1.0); gemm_acc(Slq, Q, T, 1.0);
  for (size_t k = 0; k < S.a.size(); k++) S.a[k] = T.a[k] + S.a[k];
  Mat S0(l, 1);
  for (int i = 0; i < l; i++) S0(i, 0) = F[sL + i];
  gemm_acc(Slu, U0, S0, -1.0); gemm_acc(Slq, Q0, S0, -1.0);
  std::memcpy(Uo, U.a.data(), sizeof(double) * U.a.size());
  std::memcpy(Qo, Q.a.data(), sizeof(double) * Q.a.size());
  std::memcpy(So, S.a.data(), sizeof(double) * S.a.size());
  std::memcpy(U0o, U0.a.data(), sizeof(double) * u);
  std::memcpy(Q0o, Q0.a.data(), sizeof(double) * q);
  std::memcpy(S0o, S0.a.data(), sizeof(double) * l);
}

// face2CellMap of one element: position in faces[face_f] of node cells[cell][faceNodes[f][j]]  (HDGSolver.cpp:258-275)
void face_perm(const orc_refel* re, const orc_mesh* m, int iEl, int* perm /*[nFc*nNf]*/) {
  const int nN = re->nN, nNf = re->nNf, nFc = re->nFc;
  for (int f = 0; f < nFc; f++) {
    int gf = m->cell2face[(size_t)iEl * nFc + f];
    const int* face = m->faces + (size_t)gf * nNf;
    for (int j = 0; j < nNf; j++) {
      int node = m->cells[(size_t)iEl * nN + re->faceNodes[f * nNf + j]];
      int pos = -1;
      for (int k = 0; k < nNf; k++) if (face[k] == node) { pos = k; break; }
      if (pos < 0) throw std::runtime_error("HDGSolver : calcElementalMatrices : couldn't find cell node in face.");
      perm[f * nNf + j] = pos;
    }
  }
}

// gather of the element-local fields, HDGSolver.cpp:231-326 (serial semantics)
struct ElGather {
  std::vector<double> nodes, tau, diff, vel, bufSol, trace, solOld, fluxOld, traceOld, rkSol, rkFlux, rkTrace;
  std::vector<int> perm;
  orc_elfields ef;
};

void gather_element(const orc_refel* re, const orc_model* md, const orc_mesh* m, const orc_fields* f, int iEl, ElGather& G) {
  Sizes z(re, md->nDOF);
  const int dim = z.dim, nN = z.nN, nNf = z.nNf, nFc = z.nFc, nD = z.nDOF;
  G.perm.resize(nFc * nNf);
  face_perm(re, m, iEl, G.perm.data());
  G.nodes.resize((size_t)nN * dim);
  for (int i = 0; i < nN; i++)
    for (int d = 0; d < dim; d++) G.nodes[i * dim + d] = m->nodes[(size_t)m->cells[(size_t)iEl * nN + i] * dim + d];
  std::memset(&G.ef, 0, sizeof(G.ef));
  G.ef.nodes = G.nodes.data();
  // Tau: double-valued side selection :277-304 then permutation :306-326
  const int sT = nD * nD;
  G.tau.resize((size_t)nFc * nNf * sT);
  for (int fc = 0; fc < nFc; fc++) {
    int gf = m->cell2face[(size_t)iEl * nFc + fc];
    int side = 0;
    if (f->tauVals == 2 * sT) side = (m->face2cell[2 * (size_t)gf] == iEl) ? 0 : 1;
    for (int j = 0; j < nNf; j++)
      for (int c = 0; c < sT; c++)
        G.tau[(size_t)(fc * nNf + j) * sT + c] = f->tau[((size_t)gf * nNf + G.perm[fc * nNf + j]) * f->tauVals + side * sT + c];
  }
  G.ef.tau = G.tau.data();
  if (f->diff && md->diffComps > 0) {
    int dc = md->diffComps;
    G.diff.resize((size_t)nN * dc);
    for (int i = 0; i < nN; i++)
      for (int c = 0; c < dc; c++)
        G.diff[i * dc + c] = (f->diffType == 0) ? f->diff[(size_t)m->cells[(size_t)iEl * nN + i] * dc + c] : f->diff[((size_t)iEl * nN + i) * dc + c];
    G.ef.diff = G.diff.data();
  }
  if (f->vel) {
    G.vel.resize((size_t)nN * dim);
    for (int i = 0; i < nN; i++) for (int d = 0; d < dim; d++) G.vel[i * dim + d] = f->vel[(size_t)m->cells[(size_t)iEl * nN + i] * dim + d];
    G.ef.vel = G.vel.data();
  }
  int nSrc = (md->opmask & ORC_OP_UNABU) ? dim : 1;
  if (f->srcIP) G.ef.srcIP = f->srcIP + (size_t)iEl * nSrc * z.nIP;
  if (f->reacIP) G.ef.reacIP = f->reacIP + (size_t)iEl * z.nIP;
  auto gatherFace = [&](const double* src, std::vector<double>& dst, int stage, size_t stageStride) {
    dst.resize((size_t)z.l);
    for (int fc = 0; fc < nFc; fc++) {
      int gf = m->cell2face[(size_t)iEl * nFc + fc];
      for (int j = 0; j < nNf; j++)
        for (int k = 0; k < nD; k++)
          dst[(fc * nNf + j) * nD + k] = src[stage * stageStride + ((size_t)gf * nNf + G.perm[fc * nNf + j]) * nD + k];
    }
  };
  if (f->bufSol) G.ef.bufSol = f->bufSol + (size_t)iEl * z.u;
  if (f->trace) { gatherFace(f->trace, G.trace, 0, 0); G.ef.trace = G.trace.data(); }
  if (f->solOld) G.ef.solOld = f->solOld + (size_t)iEl * z.u;
  if (f->fluxOld) G.ef.fluxOld = f->fluxOld + (size_t)iEl * z.q;
  if (f->traceOld) { gatherFace(f->traceOld, G.traceOld, 0, 0); G.ef.traceOld = G.traceOld.data(); }
  if (md->timeScheme == ORC_TS_RK && md->rkStage > 0) {
    G.rkSol.resize((size_t)md->rkStage * z.u); G.rkFlux.resize((size_t)md->rkStage * z.q); G.rkTrace.resize((size_t)md->rkStage * z.l);
    std::vector<double> tmp;
    for (int s = 0; s < md->rkStage; s++) {
      std::memcpy(&G.rkSol[(size_t)s * z.u], f->rkSol + (size_t)s * m->nCells * z.u + (size_t)iEl * z.u, sizeof(double) * z.u);
      std::memcpy(&G.rkFlux[(size_t)s * z.q], f->rkFlux + (size_t)s * m->nCells * z.q + (size_t)iEl * z.q, sizeof(double) * z.q);
      gatherFace(f->rkTrace, tmp, s, (size_t)m->nFaces * z.t);
      std::memcpy(&G.rkTrace[(size_t)s * z.l], tmp.data(), sizeof(double) * z.l);
    }
    G.ef.rkSol = G.rkSol.data(); G.ef.rkFlux = G.rkFlux.data(); G.ef.rkTrace = G.rkTrace.data();
  }
}

}  // namespace

extern "C" {

void orc_element_geometry(const orc_refel* re, const double* nodes, double* jac, double* invjac, double* dV, double* normals) {
  Geom g;
  geometry(re, nodes, g);
  std::memcpy(jac, g.jac.data(), sizeof(double) * g.jac.size());
  std::memcpy(invjac, g.invjac.data(), sizeof(double) * g.invjac.size());
  std::memcpy(dV, g.dV.data(), sizeof(double) * g.dV.size());
  std::memcpy(normals, g.normals.data(), sizeof(double) * g.normals.size());
}

void orc_local_system(const orc_refel* re, const orc_model* md, const orc_elfields* f, double* A, double* F) {
  Mat M; std::vector<double> r;
  local_system(re, md, f, M, r);
  std::memcpy(A, M.a.data(), sizeof(double) * M.a.size());
  std::memcpy(F, r.data(), sizeof(double) * r.size());
}

void orc_op_base(const orc_refel* re, int nDOF, const double* nodes, const double* tau, double* A) {
  Sizes z(re, nDOF); Geom g; geometry(re, nodes, g);
  Mat M(z.n, z.n); op_base(re, z, g, tau, M);
  std::memcpy(A, M.a.data(), sizeof(double) * M.a.size());
}
void orc_op_diffusion(const orc_refel* re, int nDOF, const double* nodes, const double* diff, int diffComps, double* A) {
  Sizes z(re, nDOF); Geom g; geometry(re, nodes, g);
  Mat M(z.n, z.n); op_diffusion(re, z, g, diff, diffComps, M);
  std::memcpy(A, M.a.data(), sizeof(double) * M.a.size());
}
void orc_op_convection(const orc_refel* re, int nDOF, const double* nodes, const double* vel, double* A) {
  Sizes z(re, nDOF); Geom g; geometry(re, nodes, g);
  Mat M(z.n, z.n); op_convection(re, z, g, vel, M);
  std::memcpy(A, M.a.data(), sizeof(double) * M.a.size());
}
void orc_op_mass(int nN, int nIP, const double* shape, const double* dV, double* Mo) {
  Mat M; mass(nN, nIP, shape, dV, M);
  std::memcpy(Mo, M.a.data(), sizeof(double) * M.a.size());
}
void orc_op_unabu(const orc_refel* re, int nDOF, const double* nodes, const double* bufSol, const double* trace, double* A, double* rhs) {
  Sizes z(re, nDOF); Geom g; geometry(re, nodes, g);
  Mat M(z.n, z.n); std::vector<double> r;
  op_unabu(re, z, g, bufSol, trace, M, r);
  std::memcpy(A, M.a.data(), sizeof(double) * M.a.size());
  std::memcpy(rhs, r.data(), sizeof(double) * r.size());
}

void orc_condense(int u, int q, int l, const double* A, const double* F, int useLU,
                  double* U, double* Q, double* S, double* U0, double* Q0, double* S0) {
  Mat M(u + q + l, u + q + l);
  std::memcpy(M.a.data(), A, sizeof(double) * M.a.size());
  if (useLU) condense_t<PLU>(u, q, l, M, F, U, Q, S, U0, Q0, S0);
  else condense_t<HQR>(u, q, l, M, F, U, Q, S, U0, Q0, S0);
}

void orc_assemble_local(const orc_refel* re, const orc_model* md, const orc_mesh* m, const orc_fields* f, int e0, int e1, int useLU,
                        double* U, double* Q, double* S, double* U0, double* Q0, double* S0) {
  Sizes z(re, md->nDOF);
  ElGather G; Mat A; std::vector<double> F;
  for (int e = e0; e < e1; e++) {
    gather_element(re, md, m, f, e, G);
    local_system(re, md, &G.ef, A, F);
    size_t k = (size_t)e;
    if (useLU) condense_t<PLU>(z.u, z.q, z.l, A, F.data(), U + k * z.u * z.l, Q + k * z.q * z.l, S + k * z.l * z.l, U0 + k * z.u, Q0 + k * z.q, S0 + k * z.l);
    else condense_t<HQR>(z.u, z.q, z.l, A, F.data(), U + k * z.u * z.l, Q + k * z.q * z.l, S + k * z.l * z.l, U0 + k * z.u, Q0 + k * z.q, S0 + k * z.l);
  }
}

// HDGSolver::calcElementalMatrices with myOpts.type = WEXPLICIT / SEXPLICIT (HDGSolver.cpp:346-354): U, Q, U0, Q0 as in the implicit case, but the trace problem
// is explicit in the element's CURRENT Solution / Flux fields:  S = S_ll,  S0 = F_l - S_lu sol - S_lq flux
void orc_assemble_local_explicit(const orc_refel* re, const orc_model* md, const orc_mesh* m, const orc_fields* f, int e0, int e1, int useLU,
                                 const double* solCur, const double* fluxCur, double* U, double* Q, double* S, double* U0, double* Q0, double* S0) {
  Sizes z(re, md->nDOF);
  ElGather G; Mat A; std::vector<double> F;
  const int sL = z.u + z.q;
  for (int e = e0; e < e1; e++) {
    gather_element(re, md, m, f, e, G);
    local_system(re, md, &G.ef, A, F);
    size_t k = (size_t)e;
    double* locS = S + k * z.l * z.l; double* locS0 = S0 + k * z.l;
    if (useLU) condense_t<PLU>(z.u, z.q, z.l, A, F.data(), U + k * z.u * z.l, Q + k * z.q * z.l, locS, U0 + k * z.u, Q0 + k * z.q, locS0);
    else condense_t<HQR>(z.u, z.q, z.l, A, F.data(), U + k * z.u * z.l, Q + k * z.q * z.l, locS, U0 + k * z.u, Q0 + k * z.q, locS0);
    const double* sol = solCur + k * z.u; const double* flux = fluxCur + k * z.q;
    for (int i = 0; i < z.l; i++) {
      double s0 = F[sL + i];
      for (int j = 0; j < z.u; j++) s0 -= A(sL + i, j) * sol[j];
      for (int j = 0; j < z.q; j++) s0 -= A(sL + i, z.u + j) * flux[j];
      locS0[i] = s0;
      for (int j = 0; j < z.l; j++) locS[(size_t)j * z.l + i] = A(sL + i, sL + j);
    }
  }
}

// HDGSolver::applyBoundaryConditions :361-529 (serial, CGType boundary models: DirichletModel / IntegratedDirichletModel)
void orc_apply_bc(const orc_refel* re, const orc_model* md, const orc_mesh* m, const orc_fields* f, double* S, double* S0) {
  Sizes z(re, md->nDOF);
  const int nNf = z.nNf, nD = z.nDOF, t = z.t, l = z.l, dim = z.dim;
  for (int b = 0; b < f->nBFaces; b++) {
    int gf = f->bFaces[b];
    Mat LM(t, t); std::vector<double> LR(t, 0.0);
    const double* g = f->dirichlet + (size_t)gf * t;
    if (f->bcKind == ORC_BC_DIRICHLET) {              // DirichletModel.cpp:19-25,38-44
      for (int i = 0; i < t; i++) { LM(i, i) = 1.0; LR[i] = g[i]; }
    } else {                                          // IntegratedDirichletModel.cpp: face mass, rhs = M g
      std::vector<double> fpts((size_t)nNf * dim), jac((size_t)z.nIPf * dim * dim), dv(z.nIPf);
      for (int i = 0; i < nNf; i++) for (int d = 0; d < dim; d++) fpts[i * dim + d] = m->nodes[(size_t)m->faces[(size_t)gf * nNf + i] * dim + d];
      jacobians(fpts.data(), nNf, re->fdshape, z.nIPf, dim - 1, dim, dim, jac.data());
      for (int ip = 0; ip < z.nIPf; ip++) {
        const double* J = &jac[(size_t)ip * dim * dim];
        double G2[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int a = 0; a < dim - 1; a++) for (int c = 0; c < dim - 1; c++) { double s = 0; for (int k = 0; k < dim; k++) s += J[a * dim + k] * J[c * dim + k]; G2[a * (dim - 1) + c] = s; }
        dv[ip] = re->fw[ip] * std::sqrt(det_small(G2, dim - 1));
      }
      Mat M1; mass(nNf, z.nIPf, re->fshape, dv.data(), M1);
      for (int i = 0; i < nNf; i++) for (int j = 0; j < nNf; j++) for (int k = 0; k < nD; k++) LM(i * nD + k, j * nD + k) = M1(i, j);
      for (int i = 0; i < t; i++) { double s = 0; for (int j = 0; j < t; j++) s += LM(i, j) * g[j]; LR[i] = s; }
    }
    for (int side = 0; side < 2; side++) {
      int c = m->face2cell[2 * (size_t)gf + side];
      if (c < 0) continue;
      int lf = -1;
      for (int k = 0; k < z.nFc; k++) if (m->cell2face[(size_t)c * z.nFc + k] == gf) { lf = k; break; }
      double* locS = S + (size_t)c * l * l; double* locS0 = S0 + (size_t)c * l;
      for (int i = 0; i < t; i++) { for (int j = 0; j < l; j++) locS[(size_t)j * l + (t * lf + i)] = 0.0; locS0[t * lf + i] = 0.0; }  // Set/Set
      for (int i = 0; i < t; i++) { for (int j = 0; j < t; j++) locS[(size_t)(t * lf + j) * l + (t * lf + i)] += LM(i, j); locS0[t * lf + i] += LR[i]; }
    }
  }
}

void orc_elem_dofs(const orc_refel* re, int nDOF, const orc_mesh* m, int iEl, int* dofs) {
  const int nNf = re->nNf, nFc = re->nFc;
  std::vector<int> perm(nFc * nNf);
  face_perm(re, m, iEl, perm.data());
  for (int f = 0; f < nFc; f++)
    for (int j = 0; j < nNf; j++)
      for (int k = 0; k < nDOF; k++)
        dofs[(f * nNf + j) * nDOF + k] = (m->cell2face[(size_t)iEl * nFc + f] * nNf + perm[f * nNf + j]) * nDOF + k;  // HDGSolver.cpp:596
}

long long orc_csr_pattern(const orc_refel* re, int nDOF, const orc_mesh* m, long long* rowptr, int* colidx) {
  // rows of face F couple to every dof of every face of every cell adjacent to F (HDGSolver.cpp:130-148); columns sorted (PETSc AIJ)
  const int nNf = re->nNf, nFc = re->nFc, t = nNf * nDOF;
  long long nnz = 0;
  std::vector<int> nb;
  for (int F = 0; F < m->nFaces; F++) {
    nb.clear();
    for (int s = 0; s < 2; s++) {
      int c = m->face2cell[2 * (size_t)F + s];
      if (c < 0) continue;
      for (int k = 0; k < nFc; k++) nb.push_back(m->cell2face[(size_t)c * nFc + k]);
    }
    std::sort(nb.begin(), nb.end());
    nb.erase(std::unique(nb.begin(), nb.end()), nb.end());
    for (int a = 0; a < t; a++) {
      long long row = (long long)F * t + a;
      if (rowptr) rowptr[row] = nnz;
      if (colidx)
        for (size_t g = 0; g < nb.size(); g++) for (int b2 = 0; b2 < t; b2++) colidx[nnz + (long long)g * t + b2] = nb[g] * t + b2;
      nnz += (long long)nb.size() * t;
    }
  }
  if (rowptr) rowptr[(long long)m->nFaces * t] = nnz;
  return nnz;
}

static inline long long csr_find(const long long* rowptr, const int* colidx, long long row, int col) {
  const int* b = colidx + rowptr[row]; const int* e = colidx + rowptr[row + 1];
  const int* p = std::lower_bound(b, e, col);
  if (p == e || *p != col) throw std::runtime_error("oracle: CSR entry not in pattern");
  return p - colidx;
}

void orc_scatter(const orc_refel* re, int nDOF, const orc_mesh* m, const double* S, const double* S0,
                 const long long* rowptr, const int* colidx, double* vals, double* rhs) {
  const int l = re->nFc * re->nNf * nDOF;
  std::vector<int> dofs(l);
  for (int e = 0; e < m->nCells; e++) {
    orc_elem_dofs(re, nDOF, m, e, dofs.data());
    const double* locS = S + (size_t)e * l * l;
    for (int i = 0; i < l; i++) {
      for (int j = 0; j < l; j++) vals[csr_find(rowptr, colidx, dofs[i], dofs[j])] += locS[(size_t)j * l + i];
      rhs[dofs[i]] += S0[(size_t)e * l + i];
    }
  }
}

int orc_gmres(long long n, const long long* rowptr, const int* colidx, const double* vals, const double* b, double* x,
              int restart, int pc, int bs, double rtol, int maxits, double* resnorm) {
  // PETSc KSPGMRES defaults: restart 30, classical Gram-Schmidt (no refinement), left preconditioning,
  // convergence on the preconditioned residual norm, zero initial guess (PetscInterface.cpp:222-247, PetscOpts.h:12-24).
  std::vector<double> dinv;   // point Jacobi or dense inverse blocks
  if (pc == 1) {
    dinv.assign(n, 1.0);
    for (long long i = 0; i < n; i++) {
      double d = 0;
      for (long long k = rowptr[i]; k < rowptr[i + 1]; k++) if (colidx[k] == i) d = vals[k];
      dinv[i] = (d != 0.0) ? 1.0 / d : 1.0;
    }
  } else if (pc == 2) {
    long long nb = n / bs;
    dinv.assign((size_t)nb * bs * bs, 0.0);
    for (long long B = 0; B < nb; B++) {
      Mat D(bs, bs);
      for (int i = 0; i < bs; i++) {
        long long row = B * bs + i;
        for (long long k = rowptr[row]; k < rowptr[row + 1]; k++) {
          long long c = colidx[k];
          if (c >= B * bs && c < (B + 1) * bs) D(i, (int)(c - B * bs)) = vals[k];
        }
      }
      PLU f; f.compute(D);
      Mat I(bs, bs);
      for (int i = 0; i < bs; i++) I(i, i) = 1.0;
      f.solve(I);
      std::memcpy(&dinv[(size_t)B * bs * bs], I.a.data(), sizeof(double) * bs * bs);
    }
  }
  auto applyPC = [&](const std::vector<double>& r, std::vector<double>& z) {
    if (pc == 0) { z = r; return; }
    if (pc == 1) { for (long long i = 0; i < n; i++) z[i] = dinv[i] * r[i]; return; }
    long long nb = n / bs;
    for (long long B = 0; B < nb; B++)
      for (int i = 0; i < bs; i++) {
        double s = 0;
        for (int j = 0; j < bs; j++) s += dinv[(size_t)B * bs * bs + (size_t)j * bs + i] * r[B * bs + j];
        z[B * bs + i] = s;
      }
  };
  auto spmv = [&](const double* v, std::vector<double>& y) {
    for (long long i = 0; i < n; i++) {
      double s = 0;
      for (long long k = rowptr[i]; k < rowptr[i + 1]; k++) s += vals[k] * v[colidx[k]];
      y[i] = s;
    }
  };
  auto nrm2 = [&](const std::vector<double>& v) { double s = 0; for (long long i = 0; i < n; i++) s += v[i] * v[i]; return std::sqrt(s); };
  const int m = restart;
  std::vector<std::vector<double>> V(m + 1, std::vector<double>(n));
  std::vector<double> H((size_t)(m + 1) * m, 0.0), cs(m), sn(m), gg(m + 1), w(n), tmp(n), bz(n), hcol(m + 1);
  for (long long i = 0; i < n; i++) x[i] = 0.0;
  std::vector<double> bvec(b, b + n);
  applyPC(bvec, bz);
  double bnorm = nrm2(bz);
  double tol = std::max(rtol * bnorm, 1e-50);
  int its = 0;
  double res = bnorm;
  if (res <= tol) { if (resnorm) *resnorm = res; return 0; }
  while (its < maxits) {
    spmv(x, tmp);
    for (long long i = 0; i < n; i++) tmp[i] = b[i] - tmp[i];
    applyPC(tmp, V[0]);
    double beta = nrm2(V[0]);
    res = beta;
    if (res <= tol) break;
    for (long long i = 0; i < n; i++) V[0][i] /= beta;
    std::fill(gg.begin(), gg.end(), 0.0);
    gg[0] = beta;
    int k = 0;
    for (; k < m && its < maxits; k++) {
      spmv(V[k].data(), tmp);
      applyPC(tmp, w);
      for (int j = 0; j <= k; j++) { double s = 0; for (long long i = 0; i < n; i++) s += w[i] * V[j][i]; hcol[j] = s; }   // CGS: all dots first
      for (int j = 0; j <= k; j++) for (long long i = 0; i < n; i++) w[i] -= hcol[j] * V[j][i];
      double hn = nrm2(w);
      hcol[k + 1] = hn;
      if (hn != 0.0) for (long long i = 0; i < n; i++) V[k + 1][i] = w[i] / hn;
      for (int j = 0; j < k; j++) { double a = cs[j] * hcol[j] + sn[j] * hcol[j + 1]; hcol[j + 1] = -sn[j] * hcol[j] + cs[j] * hcol[j + 1]; hcol[j] = a; }
      double dn = std::sqrt(hcol[k] * hcol[k] + hcol[k + 1] * hcol[k + 1]);
      if (dn == 0.0) dn = 1e-300;
      cs[k] = hcol[k] / dn; sn[k] = hcol[k + 1] / dn;
      hcol[k] = dn; hcol[k + 1] = 0.0;
      gg[k + 1] = -sn[k] * gg[k]; gg[k] = cs[k] * gg[k];
      for (int j = 0; j <= k; j++) H[(size_t)j * m + k] = hcol[j];
      its++;
      res = std::fabs(gg[k + 1]);
      if (res <= tol || hn == 0.0) { k++; break; }
    }
    std::vector<double> y(k);
    for (int i = k - 1; i >= 0; i--) {
      double s = gg[i];
      for (int j = i + 1; j < k; j++) s -= H[(size_t)i * m + j] * y[j];
      y[i] = s / H[(size_t)i * m + i];
    }
    for (int j = 0; j < k; j++) for (long long i = 0; i < n; i++) x[i] += y[j] * V[j][i];
    if (res <= tol) break;
  }
  if (resnorm) *resnorm = res;
  return its;
}

void orc_recover(const orc_refel* re, int nDOF, const orc_mesh* m, const double* traceVals,
                 const double* U, const double* Q, const double* U0, const double* Q0, double* sol, double* flux) {
  const int u = re->nN * nDOF, q = u * re->dim, l = re->nFc * re->nNf * nDOF;
  std::vector<int> dofs(l);
  std::vector<double> lam(l);
  for (int e = 0; e < m->nCells; e++) {
    orc_elem_dofs(re, nDOF, m, e, dofs.data());
    for (int i = 0; i < l; i++) lam[i] = traceVals[dofs[i]];
    const double* Ue = U + (size_t)e * u * l; const double* Qe = Q + (size_t)e * q * l;
    for (int i = 0; i < u; i++) { double s = 0; for (int j = 0; j < l; j++) s += Ue[(size_t)j * u + i] * lam[j]; sol[(size_t)e * u + i] = s + U0[(size_t)e * u + i]; }
    for (int i = 0; i < q; i++) { double s = 0; for (int j = 0; j < l; j++) s += Qe[(size_t)j * q + i] * lam[j]; flux[(size_t)e * q + i] = s + Q0[(size_t)e * q + i]; }
  }
}

double orc_bench_assemble(const orc_refel* re, const orc_model* md, const orc_mesh* m, const orc_fields* f, int nThreads, int useLU,
                          const long long* rowptr, const int* colidx, double* vals, double* rhs) {
  // One worker per host core, contiguous element ranges -- the stand-in for `mpirun -np <cores>` of the reference
  // (HDGSolver::assemble = calcElementalMatrices + applyBoundaryConditions + assembleSystem). Shared rows are added atomically,
  // which plays the role of PETSc's off-process stash.
  Sizes z(re, md->nDOF);
  const int l = z.l;
  // element -> has boundary faces (serial BC pass would follow; here done per element right after condensation)
  std::vector<char> isB(m->nFaces, 0);
  for (int b = 0; b < f->nBFaces; b++) isB[f->bFaces[b]] = 1;
  auto t0 = std::chrono::steady_clock::now();
  auto work = [&](int e0, int e1) {
    ElGather G; Mat A; std::vector<double> F;
    std::vector<double> U((size_t)z.u * l), Q((size_t)z.q * l), S((size_t)l * l), U0(z.u), Q0(z.q), S0(l);
    std::vector<int> dofs(l);
    for (int e = e0; e < e1; e++) {
      gather_element(re, md, m, f, e, G);
      local_system(re, md, &G.ef, A, F);
      if (useLU) condense_t<PLU>(z.u, z.q, z.l, A, F.data(), U.data(), Q.data(), S.data(), U0.data(), Q0.data(), S0.data());
      else condense_t<HQR>(z.u, z.q, z.l, A, F.data(), U.data(), Q.data(), S.data(), U0.data(), Q0.data(), S0.data());
      for (int k = 0; k < z.nFc; k++) {
        int gf = m->cell2face[(size_t)e * z.nFc + k];
        if (!isB[gf]) continue;
        const double* g = f->dirichlet + (size_t)gf * z.t;
        for (int i = 0; i < z.t; i++) { for (int j = 0; j < l; j++) S[(size_t)j * l + (z.t * k + i)] = 0.0; S[(size_t)(z.t * k + i) * l + (z.t * k + i)] = 1.0; S0[z.t * k + i] = g[i]; }
      }
      orc_elem_dofs(re, md->nDOF, m, e, dofs.data());
      for (int i = 0; i < l; i++) {
        for (int j = 0; j < l; j++) {
          long long p = csr_find(rowptr, colidx, dofs[i], dofs[j]);
          std::atomic_ref<double>(vals[p]).fetch_add(S[(size_t)j * l + i], std::memory_order_relaxed);
        }
        std::atomic_ref<double>(rhs[dofs[i]]).fetch_add(S0[i], std::memory_order_relaxed);
      }
    }
  };
  std::vector<std::thread> th;
  int nEl = m->nCells;
  for (int tI = 0; tI < nThreads; tI++) {
    int e0 = (int)((long long)nEl * tI / nThreads), e1 = (int)((long long)nEl * (tI + 1) / nThreads);
    th.emplace_back(work, e0, e1);
  }
  for (auto& t : th) t.join();
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // extern "C"
