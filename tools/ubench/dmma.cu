// FP64 tensor-core (mma.sync.m8n8k4.f64) latency / throughput on B200 vs. the DFMA pipe.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int ILP>
__global__ void k_dmma(long long* cyc, double* out, int iters) {
  double c[ILP][2];
  for (int i = 0; i < ILP; i++) { c[i][0] = threadIdx.x; c[i][1] = i; }
  double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int i = 0; i < ILP; i++) dmma(c[i][0], c[i][1], a, b);
  }
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1];
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (s == 12345.0) out[0] = s;
}
// mixed: DMMA and DFMA interleaved (are the pipes independent?)
__global__ void k_mixed(long long* cyc, double* out, int iters) {
  double c[4][2]; for (int i = 0; i < 4; i++) { c[i][0] = threadIdx.x; c[i][1] = i; }
  double f[8]; for (int i = 0; i < 8; i++) f[i] = i + threadIdx.x * 1e-3;
  double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int i = 0; i < 4; i++) dmma(c[i][0], c[i][1], a, b);
#pragma unroll
      for (int i = 0; i < 8; i++) f[i] = fma(f[i], a, b);
    }
  }
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < 4; i++) s += c[i][0] + c[i][1]; for (int i = 0; i < 8; i++) s += f[i];
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (s == 12345.0) out[0] = s;
}
int main() {
  long long* c; double* d; cudaMalloc(&c, 8 * 4096); cudaMalloc(&d, 8);
  long long h[4096]; const int it = 500;
  auto rep = [&](const char* name, double n_mma_per_warp) {
    cudaDeviceSynchronize(); cudaMemcpy(h, c, 8 * 4096, cudaMemcpyDeviceToHost);
    printf("%-64s %7.2f cycles per DMMA per warp  (%s)\n", name, h[0] / n_mma_per_warp, cudaGetErrorString(cudaGetLastError()));
  };
  k_dmma<1><<<1, 32>>>(c, d, it); rep("DMMA dependent chain (latency), 1 warp", 8.0 * it);
  k_dmma<2><<<1, 32>>>(c, d, it); rep("DMMA ILP=2, 1 warp", 16.0 * it);
  k_dmma<4><<<1, 32>>>(c, d, it); rep("DMMA ILP=4, 1 warp", 32.0 * it);
  k_dmma<8><<<1, 32>>>(c, d, it); rep("DMMA ILP=8, 1 warp", 64.0 * it);
  k_dmma<4><<<1, 128>>>(c, d, it); rep("DMMA ILP=4, 4 warps (1/SMSP)", 32.0 * it);
  k_dmma<4><<<1, 256>>>(c, d, it); rep("DMMA ILP=4, 8 warps (2/SMSP)", 32.0 * it);
  k_dmma<4><<<1, 512>>>(c, d, it); rep("DMMA ILP=4, 16 warps (4/SMSP)", 32.0 * it);
  k_mixed<<<1, 128>>>(c, d, it); rep("mixed 4 DMMA + 8 DFMA per round, 4 warps: cycles per DMMA", 32.0 * it);
  // full-chip throughput
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int blocks = 148 * 4, threads = 256, it2 = 4000;
  for (int r = 0; r < 3; r++) {
    cudaEventRecord(e0); k_dmma<4><<<blocks, threads>>>(c, d, it2); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 256 * 32.0 * it2 * (threads / 32) * blocks;
    printf("full chip DMMA: %.2f TFLOP/s (%.3f ms)\n", flops / (ms * 1e-3) / 1e12, ms);
  }
  return 0;
}
