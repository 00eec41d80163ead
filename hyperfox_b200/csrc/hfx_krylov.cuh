// Device-resident GMRES(m) for the condensed trace system (replaces KSPSolve, src/resolution/PetscInterface.cpp:222-247, with the
// reference's PetscOpts: left preconditioning, classical Gram-Schmidt, zero initial guess, test on the preconditioned residual).
//
// What is different from a textbook host-driven loop, and why:
//   * the Hessenberg column, the Givens rotations, the residual estimate and the convergence test live on the device (one tiny
//     kernel per iteration); the host synchronises ONCE PER RESTART CYCLE, not twice per iteration;
//   * the Krylov basis is stored UNNORMALISED, v_j = s_j V_j with s_j = 1 / ||V_j||: the norm of the newest vector is reduced together
//     with the Gram-Schmidt dots of the NEXT iteration, so an iteration has exactly one reduction (one ncclAllReduce of k + 2 doubles on
//     several GPUs) and no separate "scale" pass over the vector.  Column k-1 of the Hessenberg matrix is therefore closed (last entry,
//     rotation, residual) at the start of iteration k; the arithmetic is that of classical Gram-Schmidt GMRES up to the rounding of the
//     deferred scaling;
//   * an iteration is three passes over HBM: SpMV with the Jacobi scaling in its epilogue, one pass that forms all dots (z read once,
//     every basis vector once), one pass that forms V_{k+1} = z - sum_j c_j V_j and the partial sums of its norm;
//   * every kernel starts by reading the `done` flag and returns when the solve has converged inside a cycle.
// Distributed solve: vectors are indexed by LOCAL face; rows of ghost faces are never written by the SpMV / Gram-Schmidt passes (mask),
// their entries of a basis vector are filled by the halo exchange right before that vector is multiplied.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace hfx {

constexpr int kMaxRestart = 30;
constexpr int kDotBlocks = 592;   // 4 x 148

struct GmresDev {
  double Hcol[kMaxRestart + 2];                 // raw column k (h_jk, j <= k), waiting for h_{k+1,k}
  double R[kMaxRestart * kMaxRestart];          // rotated Hessenberg = upper triangular, R[j + kMaxRestart * k]
  double cs[kMaxRestart], sn[kMaxRestart], g[kMaxRestart + 1];
  double s[kMaxRestart + 1];                    // 1 / ||V_j||
  double c[kMaxRestart + 1];                    // coefficients of the current Gram-Schmidt pass, then of the solution update
  double tol, bnorm, res, rtol;
  int its, done, kc, first, maxits, converged;
};

// partial[j * gridDim.x + block] = sum over the block's share of z[i] * V[j][i] for the NV vectors j = NV blockIdx.y + (0..NV-1), j < nv.
// A block column handles NV = 8 vectors: 38 registers, every SM full of warps that each keep eight 16-byte loads in flight (measured
// 6.5 TB/s on the B200; with all 30 accumulators in one thread the kernel needs 154 registers and drops to 4 TB/s).  z is re-read once per
// block column: 1/8 more traffic.
template <int NV>
__global__ void __launch_bounds__(256) kry_dots_kernel(long long n, int nv, const double* __restrict__ V, long long ldv, const double* __restrict__ z,
                                                       double* __restrict__ partial, const int* __restrict__ done) {
  if (done && *done) return;
  const int j0 = blockIdx.y * NV;
  V += (size_t)j0 * ldv; nv -= j0;
  double acc[NV];
#pragma unroll
  for (int j = 0; j < NV; j++) acc[j] = 0.0;
  const long long n2 = n >> 1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
    const double2 zz = reinterpret_cast<const double2*>(z)[i];
#pragma unroll
    for (int j = 0; j < NV; j++)
      if (j < nv) {
        const double2 v = *reinterpret_cast<const double2*>(V + (size_t)j * ldv + 2 * i);
        acc[j] = fma(zz.x, v.x, fma(zz.y, v.y, acc[j]));
      }
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
#pragma unroll
    for (int j = 0; j < NV; j++) if (j < nv) acc[j] = fma(z[n - 1], V[(size_t)j * ldv + n - 1], acc[j]);
  }
  __shared__ double red[8][NV];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < NV; j++) {
    double v = acc[j];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[w][j] = v;
  }
  __syncthreads();
  if (threadIdx.x < NV && (int)threadIdx.x < nv) {
    double s = 0.0;
    for (int k = 0; k < 8; k++) s += red[k][threadIdx.x];
    partial[(size_t)(j0 + threadIdx.x) * gridDim.x + blockIdx.x] = s;
  }
}

// out[i] = base[i] + sum_j coef[j] V[j][i] on the rows of `mask` (NULL: all rows); npart[block] = partial sum of out[i]^2 (NULL: none).
// The coefficients (sign * coef[j], coef on the device) sit in shared memory and the vectors are consumed eight at a time, so the kernel
// keeps ~40 registers whatever nv is.
__global__ void __launch_bounds__(256) kry_lincomb_kernel(long long n, int nv, const double* __restrict__ V, long long ldv, const double* base,
                                                          const double* __restrict__ coef, double sign, double* out,
                                                          const uint8_t* __restrict__ mask, double* __restrict__ npart, const int* __restrict__ done) {
  if (done && *done) return;
  __shared__ double cj[kMaxRestart + 2];
  __shared__ double red[8];
  if (threadIdx.x < kMaxRestart + 2) cj[threadIdx.x] = (int)threadIdx.x < nv ? sign * coef[threadIdx.x] : 0.0;
  __syncthreads();
  double nrm = 0.0;
  const long long n2 = n >> 1;
  const int nv8 = (nv + 7) & ~7;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
    double2 s = reinterpret_cast<const double2*>(base)[i];
    for (int j0 = 0; j0 < nv8; j0 += 8) {
      double2 v[8];
#pragma unroll
      for (int j = 0; j < 8; j++) v[j] = (j0 + j < nv) ? *reinterpret_cast<const double2*>(V + (size_t)(j0 + j) * ldv + 2 * i) : make_double2(0.0, 0.0);
#pragma unroll
      for (int j = 0; j < 8; j++) { const double c = cj[(j0 + j) < kMaxRestart + 2 ? j0 + j : 0]; s.x = fma(c, v[j].x, s.x); s.y = fma(c, v[j].y, s.y); }
    }
    if (mask) {
      const uchar2 m = reinterpret_cast<const uchar2*>(mask)[i];
      if (m.x) { out[2 * i] = s.x; nrm = fma(s.x, s.x, nrm); }
      if (m.y) { out[2 * i + 1] = s.y; nrm = fma(s.y, s.y, nrm); }
    } else {
      reinterpret_cast<double2*>(out)[i] = s;
      nrm = fma(s.x, s.x, fma(s.y, s.y, nrm));
    }
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0 && (!mask || mask[n - 1])) {
    double s = base[n - 1];
    for (int j = 0; j < nv; j++) s = fma(cj[j], V[(size_t)j * ldv + n - 1], s);
    out[n - 1] = s; nrm = fma(s, s, nrm);
  }
  if (npart) {
    for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = nrm;
    __syncthreads();
    if (threadIdx.x == 0) { double s = 0.0; for (int k = 0; k < 8; k++) s += red[k]; npart[blockIdx.x] = s; }
  }
}

// red[j] = sum_b partial[j][b] (j < nv), red[nv] = sum_b npart[b]: deterministic, one block per scalar
__global__ void kry_reduce_kernel(int nv, int nb, const double* __restrict__ partial, const double* __restrict__ npart, double* __restrict__ red,
                                  const int* __restrict__ done) {
  if (done && *done) return;
  const int j = blockIdx.x;
  const double* src = j < nv ? partial + (size_t)j * nb : npart;
  __shared__ double sm[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) s += src[i];
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) { if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) red[j] = sm[0];
}

// Start of iteration k of a cycle (k = 0: cycle start; last != 0: the cycle's tail, no new column): red[0..k-1+1] = <z, V_j> (j <= k, unless
// last) and red[nv] = ||V_k||^2, already summed over the ranks.  Closes column k - 1 (h_{k,k-1} = s_{k-1} ||V_k||, rotations, residual,
// convergence), then opens column k: s_k, c_j = s_j^2 <z, V_j> (Gram-Schmidt coefficients on the unnormalised basis), raw h_jk = s_k s_j <z, V_j>.
__global__ void kry_step_kernel(GmresDev* S, int k, int nv, int last, const double* __restrict__ red) {
  if (threadIdx.x != 0 || S->done) return;
  const int m = kMaxRestart;
  const double nrm2 = red[nv];
  const double nk = sqrt(nrm2);
  if (k == 0) {
    if (S->first) { S->bnorm = nk; const double t = S->rtol * nk; S->tol = t > 1e-50 ? t : 1e-50; S->first = 0; }
    S->g[0] = nk; S->res = nk;
    if (nk <= S->tol) { S->done = 1; S->kc = 0; S->converged = 1; return; }
  } else {
    double* h = S->Hcol;
    const int c = k - 1;
    h[k] = S->s[c] * nk;
    for (int j = 0; j < c; j++) { const double a = S->cs[j] * h[j] + S->sn[j] * h[j + 1]; h[j + 1] = -S->sn[j] * h[j] + S->cs[j] * h[j + 1]; h[j] = a; }
    double dn = sqrt(h[c] * h[c] + h[c + 1] * h[c + 1]);
    if (dn == 0.0) dn = 1e-300;
    S->cs[c] = h[c] / dn; S->sn[c] = h[c + 1] / dn;
    h[c] = dn; h[c + 1] = 0.0;
    S->g[c + 1] = -S->sn[c] * S->g[c]; S->g[c] = S->cs[c] * S->g[c];
    for (int j = 0; j <= c; j++) S->R[j + m * c] = h[j];
    S->its++;
    S->res = fabs(S->g[c + 1]);
    S->kc = k;
    if (S->res <= S->tol) { S->done = 1; S->converged = 1; return; }
    if (nrm2 == 0.0 || S->its >= S->maxits) { S->done = 1; S->converged = S->res <= S->tol ? 1 : 0; return; }
  }
  if (last) return;
  const double sk = 1.0 / nk;
  S->s[k] = sk;
  for (int j = 0; j <= k; j++) {
    const double sj = S->s[j];
    S->c[j] = sj * sj * red[j];
    S->Hcol[j] = sk * sj * red[j];
  }
}

// End of a cycle: y = R^-1 g on the kc closed columns; the update x += sum_j (y_j s_j) V_j uses c[] as its coefficients.  Not gated by `done`.
__global__ void kry_backsolve_kernel(GmresDev* S) {
  if (threadIdx.x != 0) return;
  const int m = kMaxRestart, kc = S->kc;
  double y[kMaxRestart];
  for (int i = kc - 1; i >= 0; i--) {
    double s = S->g[i];
    for (int j = i + 1; j < kc; j++) s -= S->R[i + m * j] * y[j];
    y[i] = s / S->R[i + m * i];
  }
  for (int j = 0; j < m + 1; j++) S->c[j] = j < kc ? y[j] * S->s[j] : 0.0;
}

// z = dinv .* (b - y) (b != NULL) or dinv .* y on the rows of mask; other rows get 0
__global__ void kry_pc_kernel(long long n, const double* __restrict__ dinv, const double* __restrict__ y, const double* __restrict__ b, double* __restrict__ z,
                              const uint8_t* __restrict__ mask) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double r = 0.0;
  if (!mask || mask[i]) { r = b ? b[i] - (y ? y[i] : 0.0) : y[i]; if (dinv) r *= dinv[i]; }
  z[i] = r;
}
__global__ void kry_scale_rows_kernel(long long n, const double* __restrict__ dinv, double* __restrict__ y, const int* __restrict__ done) {
  if (done && *done) return;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] *= dinv[i];
}

}  // namespace hfx
