// Micro-benchmarks, part 2: alternatives to per-lane RED for the scatter.
//   bulkred   cp.reduce.async.bulk.global.shared::cta.add.f64 (TMA reduce-add) of contiguous chunks
//   smematom  atomicAdd(double) on shared memory with the element-scatter pattern
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench2 ubench2.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void bulk_reduce_add_f64(double* gdst, const double* ssrc, uint32_t bytes) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(gdst), "r"(s), "r"(bytes) : "memory");
}

// each CTA: fill smem with ones, then reduce `nchunks` chunks of `chunk` doubles to pseudo-random-ish (block-strided) places
__global__ void k_bulkred(double* vals, long long nvals, int chunk, int chunks_per_cta_iter, int iters) {
  extern __shared__ __align__(128) double sbuf[];
  const int total = chunk * chunks_per_cta_iter;
  for (int i = threadIdx.x; i < total; i += blockDim.x) sbuf[i] = 1.0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  for (int it = 0; it < iters; it++) {
    if (threadIdx.x < chunks_per_cta_iter) {
      // destination: emulate neighbouring CTAs overlapping: stride by 2/3 of the CTA footprint
      long long base = ((long long)(blockIdx.x + (long long)it * gridDim.x) * total * 2 / 3 + (long long)threadIdx.x * chunk) % (nvals - chunk);
      base &= ~1LL;  // 16-byte alignment
      bulk_reduce_add_f64(vals + base, sbuf + threadIdx.x * chunk, chunk * 8);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncthreads();
  }
  if (threadIdx.x < chunks_per_cta_iter) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// smem atomics: each warp adds a 27x27 block into a 6x6x6x125-entry staging area (as design A would)
__global__ void k_smematom(double* out, int iters, int mode) {
  extern __shared__ __align__(16) double stage[];  // 216*125 = 27000 doubles = 216 KB
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i = threadIdx.x; i < 216 * 125; i += blockDim.x) stage[i] = 0.;
  __syncthreads();
  for (int it = 0; it < iters; it++) {
    for (int e = warp; e < 64; e += nw) {
      const int e1 = e >> 4, e2 = (e >> 2) & 3, e3 = e & 3;
      for (int k = lane; k < 729; k += 32) {
        const int a = k / 27, b = k % 27;
        const int a1 = a / 9, a2 = (a / 3) % 3, a3 = a % 3, b1 = b / 9, b2 = (b / 3) % 3, b3 = b % 3;
        const int I = ((e1 + a1) * 6 + e2 + a2) * 6 + e3 + a3;
        const int pos = ((b1 - a1 + 2) * 5 + (b2 - a2 + 2)) * 5 + (b3 - a3 + 2);
        if (mode == 0) atomicAdd(&stage[I * 125 + pos], 1.0);
        else stage[I * 125 + pos] += 1.0;  // racy plain RMW: cost floor
      }
    }
    __syncthreads();
  }
  double s = 0;
  for (int i = threadIdx.x; i < 216 * 125; i += blockDim.x) s += stage[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float ms;
  const long long nvals = 130LL * 130 * 130 * 125;
  double* vals;
  CK(cudaMalloc(&vals, sizeof(double) * nvals));
  CK(cudaMemset(vals, 0, sizeof(double) * nvals));
  CK(cudaFuncSetAttribute(k_bulkred, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int chunk : {126, 750, 3000}) {
    for (int ctas_per_sm : {1, 2}) {
      const int per_iter = (ctas_per_sm == 1 ? 24000 : 12000) / chunk;  // ~192 KB or ~96 KB of smem
      const int iters = 50, blocks = prop.multiProcessorCount * ctas_per_sm;
      const size_t smem = (size_t)per_iter * chunk * 8;
      for (int rep = 0; rep < 2; rep++) {
        CK(cudaEventRecord(e0));
        k_bulkred<<<blocks, 128, smem>>>(vals, nvals, chunk, per_iter, iters);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        CK(cudaGetLastError());
      }
      const double bytes = (double)blocks * iters * per_iter * chunk * 8;
      printf("bulkred chunk=%d doubles, %d CTA/SM, %d chunks/iter: %.3f ms, %.1f GB/s reduced (%.1f G f64 adds/s)\n", chunk, ctas_per_sm, per_iter, ms, bytes / ms / 1e6, bytes / 8 / ms / 1e6);
    }
  }
  double* out;
  CK(cudaMalloc(&out, sizeof(double) * 148 * 1024));
  CK(cudaFuncSetAttribute(k_smematom, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 125 * 8));
  for (int mode = 0; mode < 2; mode++)
    for (int threads : {256, 512, 1024}) {
      const int iters = 20;
      for (int rep = 0; rep < 2; rep++) {
        CK(cudaEventRecord(e0));
        k_smematom<<<prop.multiProcessorCount, threads, 216 * 125 * 8>>>(out, iters, mode);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        CK(cudaGetLastError());
      }
      const double adds = (double)prop.multiProcessorCount * iters * 64 * 729;
      printf("smem %s threads=%d: %.3f ms, %.1f G adds/s chip-wide, %.1f cycles/element/SM @1.9GHz\n", mode ? "plain RMW" : "atomicAdd", threads, ms, adds / ms / 1e6,
             ms * 1e-3 * 1.9e9 / (iters * 64.0));
    }
  return 0;
}
