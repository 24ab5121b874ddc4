// Micro-benchmarks that bound the assembly kernel design on B200 (run with gpurun):
//   dfma     fp64 FMA issue rate (register operands)
//   dmma     mma.sync m8n8k4 f64 rate
//   scatter  fp64 atomicAdd (RED) throughput with the element-scatter address pattern of a 128^3 p=2 mesh
//   stream   plain store bandwidth into the same array
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void k_dfma(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void k_dmma(double* out, int iters) {
  double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
  double c0 = 0, c1 = 0, d0 = 0, d1 = 0, e0 = 0, e1 = 0, f0 = 0, f1 = 0;
  for (int i = 0; i < iters; i++) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(e0), "+d"(e1) : "d"(a), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(f0), "+d"(f1) : "d"(a), "d"(b));
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1 + d0 + d1 + e0 + e1 + f0 + f1;
}

// mode 0: atomicAdd per (a,b) entry of a 27x27 block, warp per element, lanes stride over entries (row-major)
// mode 1: same addresses, plain store (racy; measures the store path)
// mode 2: atomicAdd, but each warp writes 729 CONTIGUOUS doubles (ideal coalescing, no sharing)
// mode 3: symmetric-half + smem-free: only 378 entries (upper triangle) per element, atomics
__global__ void k_scatter(double* vals, int n, int mode, long long nelem) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int nd = n + 2;
  for (long long e = warp; e < nelem; e += nwarps) {
    const int e3 = e % n, e2 = (e / n) % n, e1 = e / ((long long)n * n);
    if (mode == 2) {
      for (int k = lane; k < 729; k += 32) atomicAdd(vals + e * 729 % ((long long)nd * nd * nd * 125 - 729) + k, 1.0);
      continue;
    }
    for (int k = lane; k < 729; k += 32) {
      const int a = k / 27, b = k % 27;
      if (mode == 3 && b < a) continue;
      const int a1 = a / 9, a2 = (a / 3) % 3, a3 = a % 3, b1 = b / 9, b2 = (b / 3) % 3, b3 = b % 3;
      const long long I = ((long long)(e1 + a1) * nd + e2 + a2) * nd + e3 + a3;
      const int pos = ((b1 - a1 + 2) * 5 + (b2 - a2 + 2)) * 5 + (b3 - a3 + 2);
      double* p = vals + I * 125 + pos;
      if (mode == 1) *p = 1.0; else atomicAdd(p, 1.0);
    }
  }
}

__global__ void k_stream(double* vals, long long n) {
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  const long long stride = (long long)gridDim.x * blockDim.x * 2;
  for (; i + 1 < n; i += stride) *reinterpret_cast<double2*>(vals + i) = make_double2(1.0, 2.0);
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s sms %d clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float ms;
  double* out;
  CK(cudaMalloc(&out, sizeof(double) * 148 * 16 * 1024));
  for (int threads : {128, 256, 512, 1024}) {
    const int blocks = prop.multiProcessorCount * (2048 / threads), iters = 20000;
    k_dfma<<<blocks, threads>>>(out, 100, 1.0000001, 1e-9);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    k_dfma<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("dfma threads=%d: %.2f TFLOP/s (%.3f ms)\n", threads, 2.0 * 8 * iters * (double)blocks * threads / ms / 1e9, ms);
  }
  {
    const int threads = 256, blocks = prop.multiProcessorCount * 8, iters = 20000;
    k_dmma<<<blocks, threads>>>(out, 100);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    k_dmma<<<blocks, threads>>>(out, iters);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("dmma m8n8k4: %.2f TFLOP/s (%.3f ms)\n", 2.0 * 256 * 4 * iters * (double)blocks * threads / 32 / ms / 1e9, ms);
  }
  const int n = 128, nd = n + 2;
  const long long nvals = (long long)nd * nd * nd * 125, nelem = (long long)n * n * n;
  double* vals;
  CK(cudaMalloc(&vals, sizeof(double) * nvals));
  CK(cudaMemset(vals, 0, sizeof(double) * nvals));
  for (int rep = 0; rep < 2; rep++) {
    CK(cudaEventRecord(e0));
    k_stream<<<prop.multiProcessorCount * 16, 256>>>(vals, nvals);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("stream store: %.1f GB/s (%.3f ms for %.2f GB)\n", nvals * 8.0 / ms / 1e6, ms, nvals * 8.0 / 1e9);
    CK(cudaEventRecord(e0));
    CK(cudaMemsetAsync(vals, 0, sizeof(double) * nvals));
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("cudaMemset: %.1f GB/s (%.3f ms)\n", nvals * 8.0 / ms / 1e6, ms);
  }
  for (int mode = 0; mode < 4; mode++)
    for (int rep = 0; rep < 2; rep++) {
      CK(cudaEventRecord(e0));
      k_scatter<<<prop.multiProcessorCount * 8, 256>>>(vals, n, mode, nelem);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
      const double cnt = (mode == 3 ? 378.0 : 729.0) * nelem;
      printf("scatter mode %d: %.3f ms, %.1f G atomics/s, %.1f GB/s payload\n", mode, ms, cnt / ms / 1e6, cnt * 8 / ms / 1e6);
    }
  CK(cudaGetLastError());
  return 0;
}
