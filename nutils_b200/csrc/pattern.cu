// Analytic CSR pattern of a tensor-product spline space.
//
// Replaces the reference's post-loop argsort/unique/compress_indices over nelems*n_e^2 COO keys
// (src/nutils/evaluable.py:588-616, 5646-5682; numeric.py:687-711).  Dof i_d along dimension d
// couples with the contiguous range [lo_d(i_d), lo_d(i_d)+wid_d(i_d)) (union of the dof ranges of
// the elements in its support); the coupled set of the tensor dof I is the box product, so the
// sorted column list of row (I, c) is { J*ncomp + e : J in box(I) in C order, e < ncomp } and
// rowptr follows from per-dimension prefix sums.  One warp writes the rows of one basis function;
// lanes stride over the row, so colidx stores are coalesced.  HBM-bound: 8 B per stored entry.

#include <algorithm>

#include "common.cuh"

template <int DIM>
__global__ void __launch_bounds__(256) k_pattern_export(const BasisView B, const long long nbasis, long long* __restrict__ rowptr,
                                                         long long* __restrict__ colidx) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int nc = B.ncomp;
  for (long long I = warp0; I < nbasis; I += nwarps) {
    int i[3] = {0, 0, 0};
    long long r = I;
    for (int d = DIM - 1; d >= 0; d--) {
      i[d] = (int)(r % B.ndofs[d]);
      r /= B.ndofs[d];
    }
    int lo[3] = {0, 0, 0}, wid[3] = {1, 1, 1};
    for (int d = 0; d < DIM; d++) {
      lo[d] = B.lo[d][i[d]];
      wid[d] = B.wid[d][i[d]];
    }
    const long long R = row_start_basis<DIM>(B, i);
    const int w = wid[0] * wid[1] * wid[2];
    const long long base = R * nc * nc;
    const int rowlen = w * nc;
    if (lane < nc) rowptr[I * nc + lane] = base + (long long)lane * rowlen;
    if (I == nbasis - 1 && lane == 0) rowptr[nbasis * nc] = base + (long long)nc * rowlen;
    for (int c = 0; c < nc; c++) {
      long long* out = colidx + base + (long long)c * rowlen;
      for (int k = lane; k < rowlen; k += 32) {
        const int e = k % nc;
        int pos = k / nc;
        long long J;
        if (DIM == 1) {
          J = lo[0] + pos;
        } else if (DIM == 2) {
          J = (long long)(lo[0] + pos / wid[1]) * B.ndofs[1] + lo[1] + pos % wid[1];
        } else {
          const int j2 = pos % wid[2];
          pos /= wid[2];
          J = ((long long)(lo[0] + pos / wid[1]) * B.ndofs[1] + lo[1] + pos % wid[1]) * B.ndofs[2] + lo[2] + j2;
        }
        out[k] = J * nc + e;
      }
    }
  }
}

int launch_pattern_export(b2_ctx* ctx, const BasisView& B, long long* rowptr, long long* colidx) {
  long long nbasis = 1;
  for (int d = 0; d < B.ndims; d++) nbasis *= B.ndofs[d];
  const int threads = 256;
  const long long want = (nbasis + 7) / 8;
  const int blocks = (int)std::min<long long>(std::max<long long>(want, 1), (long long)ctx->sm_count * 16);
  switch (B.ndims) {
    case 1: k_pattern_export<1><<<blocks, threads, 0, ctx->stream>>>(B, nbasis, rowptr, colidx); break;
    case 2: k_pattern_export<2><<<blocks, threads, 0, ctx->stream>>>(B, nbasis, rowptr, colidx); break;
    default: k_pattern_export<3><<<blocks, threads, 0, ctx->stream>>>(B, nbasis, rowptr, colidx); break;
  }
  ctx->launches++;
  B2_CUDA(ctx, cudaGetLastError());
  return B2_OK;
}
