// Micro-benchmark, part 3: FP64 tensor-core rate (mma.sync.aligned.m8n8k4.f64, SASS DMMA.8x8x4) against the DFMA rate,
// with NACC independent accumulator tiles per warp and WPB warps per block.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench3 ubench3.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void k_dmma(double* out, int iters) {
  double c[NACC][2];
  for (int i = 0; i < NACC; i++) c[i][0] = c[i][1] = 0.;
  double a = threadIdx.x * 1e-3, b = 1. + blockIdx.x * 1e-6;
  for (int it = 0; it < iters; it++)
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma(c[i][0], c[i][1], a, b);
  double s = 0;
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dfma(double* out, int iters) {
  double c[NACC];
  for (int i = 0; i < NACC; i++) c[i] = i;
  double a = 1. + threadIdx.x * 1e-9, b = blockIdx.x * 1e-6;
  for (int it = 0; it < iters; it++)
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
  double s = 0;
  for (int i = 0; i < NACC; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float timeit(F f) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  f();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
  const int iters = 20000;
  for (int wpb : {4, 8, 16, 32}) {
    const int threads = wpb * 32, blocks = sms;
    float ms = timeit([&] { k_dmma<8><<<blocks, threads>>>(out, iters); });
    double flop = 2.0 * 256 * 8 * (double)iters * wpb * blocks;
    printf("DMMA m8n8k4  8 acc tiles/warp, %2d warps/SM: %7.3f ms  %6.2f TFLOP/s\n", wpb, ms, flop / ms * 1e-9);
    ms = timeit([&] { k_dmma<2><<<blocks, threads>>>(out, iters); });
    flop = 2.0 * 256 * 2 * (double)iters * wpb * blocks;
    printf("DMMA m8n8k4  2 acc tiles/warp, %2d warps/SM: %7.3f ms  %6.2f TFLOP/s\n", wpb, ms, flop / ms * 1e-9);
    ms = timeit([&] { k_dfma<16><<<blocks, threads>>>(out, iters); });
    flop = 2.0 * 32 * 16 * (double)iters * wpb * blocks;
    printf("DFMA        16 chains/thread,  %2d warps/SM: %7.3f ms  %6.2f TFLOP/s\n", wpb, ms, flop / ms * 1e-9);
  }
  return 0;
}
