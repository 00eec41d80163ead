// Isolated timing of the Gauss-Jordan step variants used by the HDG kernel (20x20, ld 20, 128 threads).
#include <cstdio>
#include <cuda_runtime.h>
__host__ __device__ constexpr int ev(int x) { return (x + 1) & ~1; }
__device__ __forceinline__ double fast_rcp(double x) {
  double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(fma(-x, r, 1.0), r, r); r = fma(fma(-x, r, 1.0), r, r); return r;
}
__device__ __forceinline__ void bar_named(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
constexpr int GT = 128;
template <int n, int ld, int VARIANT>
__device__ __noinline__ void gj(double* b0, double* b1, int tid) {
  constexpr int MT = ev(n) / 2, NS = MT * n, NQ = (NS + GT - 1) / GT;
  const double* src = b0; double* dst = b1;
#pragma unroll 1
  for (int k = 0; k < n; k++) {
    const double piv = src[k + ld * k];
    double2 a[NQ], c[NQ]; double pj[NQ];
#pragma unroll
    for (int q = 0; q < NQ; q++) { const int s2 = tid + q * GT; if (s2 < NS) { const int i0 = (s2 % MT) * 2, j = s2 / MT;
      a[q] = *reinterpret_cast<const double2*>(src + i0 + ld * j); c[q] = *reinterpret_cast<const double2*>(src + i0 + ld * k); pj[q] = src[k + ld * j]; } }
    double ip;
    if (VARIANT == 0 || VARIANT >= 4) ip = fast_rcp(piv); else if (VARIANT == 1) ip = 1.0 / piv; else ip = piv;   // 2: no reciprocal at all (timing only)
#pragma unroll
    for (int q = 0; q < NQ; q++) { const int s2 = tid + q * GT; if (s2 < NS) { const int i0 = (s2 % MT) * 2, j = s2 / MT;
      double r0, r1;
      if (VARIANT >= 4) {
        const bool diag = (j == k);
        const double tt = diag ? ip : pj[q] * ip;
        const double ax = diag ? 0.0 : a[q].x, ay = diag ? 0.0 : a[q].y;
        r0 = (i0 == k) ? tt : fma(-c[q].x, tt, ax);
        r1 = (i0 + 1 == k) ? tt : fma(-c[q].y, tt, ay);
      } else if (j == k) { r0 = (i0 == k) ? ip : -c[q].x * ip; r1 = (i0 + 1 == k) ? ip : -c[q].y * ip; }
      else { const double pji = pj[q] * ip; r0 = (i0 == k) ? pji : fma(-c[q].x, pji, a[q].x); r1 = (i0 + 1 == k) ? pji : fma(-c[q].y, pji, a[q].y); }
      *reinterpret_cast<double2*>(dst + i0 + ld * j) = make_double2(r0, r1); } }
    if (VARIANT == 3) __syncthreads(); else if (VARIANT == 5) __syncwarp(); else bar_named(1, GT);
    const double* t = dst; dst = const_cast<double*>(src); src = t;
  }
}
template <int VARIANT>
__global__ void kern(long long* cyc, double* out, int reps) {
  __shared__ __align__(16) double A[400], B[400];
  const int tid = threadIdx.x;
  for (int i = tid; i < 400; i += blockDim.x) { A[i] = (i % 21 == 0) ? 4.0 + i * 0.01 : 0.01 * ((i * 7) % 13); B[i] = 0; }
  __syncthreads();
  long long t0 = clock64();
  for (int r = 0; r < reps; r++) { if (tid < GT) gj<20, 20, VARIANT>(A, B, tid); __syncthreads(); }
  long long t1 = clock64();
  if (tid == 0) { cyc[blockIdx.x] = t1 - t0; out[0] = A[5]; }
}
int main() {
  long long* c; double* d; cudaMalloc(&c, 8 * 1024); cudaMalloc(&d, 8);
  long long h[1024]; const int reps = 50;
  auto run = [&](const char* name, auto k, int threads, int blocks) {
    k<<<blocks, threads>>>(c, d, reps); cudaDeviceSynchronize(); cudaMemcpy(h, c, 8 * blocks, cudaMemcpyDeviceToHost);
    printf("%-70s %8.0f cycles per inversion (%.0f per pivot)\n", name, (double)h[0] / reps, (double)h[0] / reps / 20);
  };
  run("fast_rcp, named barrier(128), CTA of 128, 1 CTA", kern<0>, 128, 1);
  run("fast_rcp, named barrier(128), CTA of 256 (4 idle warps at __syncthreads)", kern<0>, 256, 1);
  run("IEEE division", kern<1>, 128, 1);
  run("no reciprocal (timing of everything else)", kern<2>, 128, 1);
  run("fast_rcp, __syncthreads (CTA of 128)", kern<3>, 128, 1);
  run("fast_rcp, named barrier, 2 CTAs per SM x 148 SMs", kern<0>, 128, 296);
  run("branch-free, fast_rcp, named barrier", kern<4>, 128, 1);
  run("branch-free, fast_rcp, NO cta barrier (timing only)", kern<5>, 128, 1);
  return 0;
}
