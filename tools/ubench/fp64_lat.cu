// Micro-benchmarks of the constants the HDG kernel design depends on (B200, sm_100a): DFMA latency / throughput per SMSP,
// FP64 division latency, shared-memory load latency, __syncthreads cost.  Run: nvcc -arch=sm_100a -O3 fp64_lat.cu && ./a.out
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma_chain(double* out, long long* cyc, int iters) {
  double a = threadIdx.x * 1e-9, b = 1.0000001, c = 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) a = fma(a, b, c);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (a == 123.456) out[0] = a;
}
template <int ILP>
__global__ void dfma_ilp(double* out, long long* cyc, int iters) {
  double a[ILP];
  for (int k = 0; k < ILP; k++) a[k] = threadIdx.x * 1e-9 + k;
  const double b = 1.0000001, c = 1e-9;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int k = 0; k < ILP; k++) a[k] = fma(a[k], b, c);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  double s = 0; for (int k = 0; k < ILP; k++) s += a[k];
  if (s == 123.456) out[0] = s;
}
__global__ void ddiv_chain(double* out, long long* cyc, int iters) {
  double a = 1.0 + threadIdx.x * 1e-3;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) a = 1.0 / (a + 0.5);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (a == 123.456) out[0] = a;
}
__global__ void lds_chain(double* out, long long* cyc, int iters) {
  __shared__ int idx[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) idx[i] = (i * 7 + 3) & 1023;
  __syncthreads();
  int j = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) j = idx[j];
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (j == -5) out[0] = j;
}
__global__ void sync_cost(double* out, long long* cyc, int iters) {
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) __syncthreads();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  double* d; long long* c; cudaMalloc(&d, 8); cudaMalloc(&c, 8 * 4096);
  long long h[4096];
  auto rep = [&](const char* name, double per) { cudaDeviceSynchronize(); cudaMemcpy(h, c, 8 * 4096, cudaMemcpyDeviceToHost); printf("%-60s %8.2f cycles/op (CTA 0)\n", name, h[0] / per); };
  const int it = 2000;
  dfma_chain<<<1, 32>>>(d, c, it); rep("DFMA dependent chain, 1 warp on the SM", 16.0 * it);
  dfma_chain<<<1, 128>>>(d, c, it); rep("DFMA dependent chain, 4 warps (1 per SMSP)", 16.0 * it);
  dfma_chain<<<1, 256>>>(d, c, it); rep("DFMA dependent chain, 8 warps (2 per SMSP)", 16.0 * it);
  dfma_chain<<<1, 512>>>(d, c, it); rep("DFMA dependent chain, 16 warps (4 per SMSP)", 16.0 * it);
  dfma_chain<<<1, 1024>>>(d, c, it); rep("DFMA dependent chain, 32 warps (8 per SMSP)", 16.0 * it);
  dfma_ilp<2><<<1, 32>>>(d, c, it); rep("DFMA ILP=2, 1 warp: cycles per DFMA", 8.0 * it);
  dfma_ilp<4><<<1, 32>>>(d, c, it); rep("DFMA ILP=4, 1 warp: cycles per DFMA", 16.0 * it);
  dfma_ilp<8><<<1, 32>>>(d, c, it); rep("DFMA ILP=8, 1 warp: cycles per DFMA", 32.0 * it);
  dfma_ilp<8><<<1, 128>>>(d, c, it); rep("DFMA ILP=8, 4 warps: cycles per DFMA per warp", 32.0 * it);
  dfma_ilp<8><<<1, 256>>>(d, c, it); rep("DFMA ILP=8, 8 warps: cycles per DFMA per warp", 32.0 * it);
  dfma_ilp<8><<<1, 512>>>(d, c, it); rep("DFMA ILP=8, 16 warps: cycles per DFMA per warp", 32.0 * it);
  ddiv_chain<<<1, 32>>>(d, c, it); rep("double division dependent chain (1/(a+0.5)), 1 warp", 8.0 * it);
  lds_chain<<<1, 32>>>(d, c, it); rep("LDS pointer chase, 1 warp", 16.0 * it);
  sync_cost<<<1, 256>>>(d, c, it); rep("__syncthreads, 256 threads", 16.0 * it);
  sync_cost<<<1, 128>>>(d, c, it); rep("__syncthreads, 128 threads", 16.0 * it);
  int dev; cudaGetDevice(&dev); int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev); printf("SM clock attr %d kHz\n", clk);
  return 0;
}
