// CSR pattern of an element set (trimmed / subset topology with a pruned dof numbering).
//
// Replaces the reference's post-loop argsort / unique / compress_indices over the COO keys of the selected
// elements (src/nutils/evaluable.py:588-616, 5646-5682; numeric.py:687-711).  No keys are materialised: the
// candidates of row I are the dofs of the box prod_d [lo_d, lo_d + wid_d) of the parent tensor space (pattern.cu),
// enumerated in C order = increasing parent index = increasing NEW index (the renumbering of a pruned basis is
// monotone, function.py:3121-3122); candidate J is kept iff at least one SELECTED element lies in
// supp(N_I) n supp(N_J).  One warp per row; pass 0 counts, an exclusive scan gives rowptr, pass 1 writes the
// columns with a ballot/popc ranking, so rows come out sorted without a sort.

#include <algorithm>

#include "common.cuh"

namespace {

struct PatParams {
  BasisView B;
  const int* efirst[B2_MAXD];  // [ndofs_d] first / last element in the support of the 1-D function
  const int* elast[B2_MAXD];
  const long long* dofmap;     // [nbasis_new] parent index of the new row, null = identity
  const int* renumber;         // parent -> new, null = identity
  const unsigned char* selmask;  // [ntotal] or null (all selected)
  long long nbasis_new;
  long long* counts;           // pass 0: [nbasis_new]
  const long long* rowptr_b;   // pass 1
  int* colidx_b;
};

template <int DIM>
__global__ void __launch_bounds__(256) k_pattern_elemset(const PatParams P, const int pass) {
  const BasisView& B = P.B;
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long In = warp0; In < P.nbasis_new; In += nwarps) {
    long long I = P.dofmap ? P.dofmap[In] : In;
    int i[3] = {0, 0, 0}, lo[3] = {0, 0, 0}, wid[3] = {1, 1, 1}, ef[3] = {0, 0, 0}, el[3] = {0, 0, 0};
    for (int d = DIM - 1; d >= 0; d--) {
      i[d] = (int)(I % B.ndofs[d]);
      I /= B.ndofs[d];
    }
    for (int d = 0; d < DIM; d++) {
      lo[d] = B.lo[d][i[d]];
      wid[d] = B.wid[d][i[d]];
      ef[d] = P.efirst[d][i[d]];
      el[d] = P.elast[d][i[d]];
    }
    const int w = wid[0] * wid[1] * wid[2];
    int count = 0;
    const long long out0 = pass ? P.rowptr_b[In] : 0;
    for (int k0 = 0; k0 < w; k0 += 32) {
      const int k = k0 + lane;
      bool keep = false;
      long long J = 0;
      if (k < w) {
        int j[3] = {0, 0, 0}, r = k;
        for (int d = DIM - 1; d >= 0; d--) {
          j[d] = lo[d] + r % wid[d];
          r /= wid[d];
        }
        int a[3] = {0, 0, 0}, b[3] = {0, 0, 0};
        bool any = true;
        for (int d = 0; d < DIM; d++) {
          a[d] = max(ef[d], P.efirst[d][j[d]]);
          b[d] = min(el[d], P.elast[d][j[d]]);
          if (a[d] > b[d]) any = false;
          J = J * B.ndofs[d] + j[d];
        }
        if (any) {
          if (!P.selmask) keep = true;
          else {
            for (int e0 = a[0]; e0 <= b[0] && !keep; e0++)
              for (int e1 = a[1]; e1 <= (DIM > 1 ? b[1] : a[1]) && !keep; e1++)
                for (int e2 = a[2]; e2 <= (DIM > 2 ? b[2] : a[2]) && !keep; e2++) {
                  long long e = e0;
                  if (DIM > 1) e = e * B.nel[1] + e1;
                  if (DIM > 2) e = e * B.nel[2] + e2;
                  keep = P.selmask[e] != 0;
                }
          }
        }
      }
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (pass && keep) P.colidx_b[out0 + count + __popc(m & ((1u << lane) - 1))] = P.renumber ? P.renumber[J] : (int)J;
      count += __popc(m);
    }
    if (!pass && lane == 0) P.counts[In] = count;
  }
}

// ---- exclusive scan of int64 counts (three small kernels; one-off work, not on the timed path) ----
constexpr int SCAN_T = 256, SCAN_I = 8, SCAN_B = SCAN_T * SCAN_I;

__global__ void __launch_bounds__(SCAN_T) k_scan_block(long long* data, long long n, long long* blocksum) {
  __shared__ long long sh[SCAN_T];
  const long long base = (long long)blockIdx.x * SCAN_B + (long long)threadIdx.x * SCAN_I;
  long long v[SCAN_I], s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_I; k++) {
    v[k] = base + k < n ? data[base + k] : 0;
    s += v[k];
  }
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int off = 1; off < SCAN_T; off <<= 1) {
    const long long t = threadIdx.x >= off ? sh[threadIdx.x - off] : 0;
    __syncthreads();
    sh[threadIdx.x] += t;
    __syncthreads();
  }
  long long run = sh[threadIdx.x] - s;  // exclusive prefix of this thread inside the block
#pragma unroll
  for (int k = 0; k < SCAN_I; k++) {
    if (base + k < n) data[base + k] = run;
    run += v[k];
  }
  if (threadIdx.x == SCAN_T - 1) blocksum[blockIdx.x] = sh[threadIdx.x];
}

__global__ void k_scan_sums(long long* blocksum, long long nblocks, long long* total) {
  // one thread: nblocks = n / 2048 is small (a 10^8-row pattern has 5 10^4 blocks)
  if (blockIdx.x || threadIdx.x) return;
  long long run = 0;
  for (long long b = 0; b < nblocks; b++) {
    const long long t = blocksum[b];
    blocksum[b] = run;
    run += t;
  }
  *total = run;
}

__global__ void __launch_bounds__(SCAN_T) k_scan_add(long long* data, long long n, const long long* blocksum) {
  const long long base = (long long)blockIdx.x * SCAN_B + (long long)threadIdx.x * SCAN_I;
  const long long add = blocksum[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_I; k++)
    if (base + k < n) data[base + k] += add;
}

// rowptr / colidx at dof level from the basis-level pattern: row (I, c) holds (J, e) for J in the basis row, e < ncomp
__global__ void __launch_bounds__(256) k_pattern_export_general(const long long* __restrict__ rowptr_b, const int* __restrict__ colidx_b, long long nbasis, int nc,
                                                                 long long* __restrict__ rowptr, long long* __restrict__ colidx) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long I = warp0; I < nbasis; I += nwarps) {
    const long long r0 = rowptr_b[I];
    const int len = (int)(rowptr_b[I + 1] - r0), rowlen = len * nc;
    const long long base = r0 * nc * nc;
    if (lane < nc) rowptr[I * nc + lane] = base + (long long)lane * rowlen;
    if (I == nbasis - 1 && lane == 0) rowptr[nbasis * nc] = base + (long long)nc * rowlen;
    for (int c = 0; c < nc; c++)
      for (int k = lane; k < rowlen; k += 32) colidx[base + (long long)c * rowlen + k] = (long long)colidx_b[r0 + k / nc] * nc + k % nc;
  }
}

}  // namespace

int launch_exclusive_scan(b2_ctx* ctx, long long* data, long long n) {
  const long long nblocks = (n + SCAN_B - 1) / SCAN_B;
  long long* blocksum = nullptr;
  B2_CUDA(ctx, cudaMalloc((void**)&blocksum, sizeof(long long) * std::max<long long>(nblocks, 1)));
  if (nblocks) k_scan_block<<<(unsigned)nblocks, SCAN_T, 0, ctx->stream>>>(data, n, blocksum);
  k_scan_sums<<<1, 32, 0, ctx->stream>>>(blocksum, nblocks, data + n);
  if (nblocks) k_scan_add<<<(unsigned)nblocks, SCAN_T, 0, ctx->stream>>>(data, n, blocksum);
  ctx->launches += 3;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(blocksum);
  if (e != cudaSuccess) return b2_cuda_fail(ctx, e, "exclusive scan");
  return B2_OK;
}

int launch_pattern_elemset_impl(b2_ctx* ctx, const BasisView& B, const int* const* efirst, const int* const* elast, const long long* dofmap, const int* renumber,
                                const unsigned char* selmask, long long nbasis_new, int pass, long long* counts_or_rowptr, int* colidx_b) {
  PatParams P;
  P.B = B;
  for (int d = 0; d < B2_MAXD; d++) {
    P.efirst[d] = efirst[d];
    P.elast[d] = elast[d];
  }
  P.dofmap = dofmap;
  P.renumber = renumber;
  P.selmask = selmask;
  P.nbasis_new = nbasis_new;
  P.counts = pass ? nullptr : counts_or_rowptr;
  P.rowptr_b = counts_or_rowptr;
  P.colidx_b = colidx_b;
  const int threads = 256;
  const long long want = (nbasis_new + 7) / 8;
  const int blocks = (int)std::min<long long>(std::max<long long>(want, 1), (long long)ctx->sm_count * 16);
  switch (B.ndims) {
    case 1: k_pattern_elemset<1><<<blocks, threads, 0, ctx->stream>>>(P, pass); break;
    case 2: k_pattern_elemset<2><<<blocks, threads, 0, ctx->stream>>>(P, pass); break;
    default: k_pattern_elemset<3><<<blocks, threads, 0, ctx->stream>>>(P, pass); break;
  }
  ctx->launches++;
  B2_CUDA(ctx, cudaGetLastError());
  return B2_OK;
}

int launch_pattern_export_general(b2_ctx* ctx, const long long* rowptr_b, const int* colidx_b, long long nbasis, int ncomp, long long* rowptr, long long* colidx) {
  const int threads = 256;
  const long long want = (nbasis + 7) / 8;
  const int blocks = (int)std::min<long long>(std::max<long long>(want, 1), (long long)ctx->sm_count * 16);
  k_pattern_export_general<<<blocks, threads, 0, ctx->stream>>>(rowptr_b, colidx_b, nbasis, ncomp, rowptr, colidx);
  ctx->launches++;
  B2_CUDA(ctx, cudaGetLastError());
  return B2_OK;
}
