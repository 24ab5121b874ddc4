// General dof maps: CSR patterns that do not come from a structured grid, and element integration on
// them.
//
//   b2_pattern_create_csr   a pattern from explicit (rowptr, colidx) arrays -- what nutils.matrix.assemble_csr hands to a
//                           backend (src/nutils/matrix/__init__.py:30-70); gives the device-resident Matrix its SpMV / CG.
//   b2_pattern_general      the pattern of sum_e rowdofs(e) x coldofs(e) from element dof lists: what the reference computes
//                           after its loop by flatten -> stable argsort -> unique -> compress_indices
//                           (evaluable.py:588-616, 5646-5682; numeric.py:687-711).  Here: 64-bit keys row << 32 | col generated
//                           on the device, cub::DeviceRadixSort, cub::DeviceSelect::Unique, row histogram + scan.
//   b2_assemble_general_*   element loop for ANY element type described by tabulated reference data (simplex and mixed meshes,
//                           topology.py:2493, mesh.py:737-753): per element type the basis values and reference gradients at
//                           its quadrature points and the (multi)linear geometry shape functions at those points; per element
//                           the vertex coordinates and the dof list.  One CTA per element: J, J^-1, |det| per point, physical
//                           jets in shared memory, block entries per thread, fp64 RED into the slot found by bisection in the
//                           row's sorted columns.  Coverage-grade (O(nq n_e^2) per element like the reference).

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>

#include <algorithm>
#include <new>

#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) k_general_keys(long long nelems, const long long* __restrict__ rowoff, const long long* __restrict__ rowdofs,
                                                       const long long* __restrict__ coloff, const long long* __restrict__ coldofs,
                                                       const long long* __restrict__ keyoff, unsigned long long* __restrict__ keys) {
  // one warp per element
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long e = warp0; e < nelems; e += nwarps) {
    const long long r0 = rowoff[e], nr = rowoff[e + 1] - r0, c0 = coloff[e], nc = coloff[e + 1] - c0, k0 = keyoff[e];
    for (long long t = lane; t < nr * nc; t += 32) keys[k0 + t] = ((unsigned long long)rowdofs[r0 + t / nc] << 32) | (unsigned long long)coldofs[c0 + t % nc];
  }
}

__global__ void __launch_bounds__(256) k_general_count(long long nelems, const long long* __restrict__ rowoff, const long long* __restrict__ coloff, long long* __restrict__ keyoff) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nelems; e += (long long)gridDim.x * blockDim.x)
    keyoff[e] = (rowoff[e + 1] - rowoff[e]) * (coloff[e + 1] - coloff[e]);
}

__global__ void __launch_bounds__(256) k_general_split(long long nuniq, const unsigned long long* __restrict__ keys, long long nrows, long long* __restrict__ rowptr, int* __restrict__ colidx) {
  // rowptr[r] = first position whose row >= r: every position writes the rows it starts
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k <= nuniq; k += (long long)gridDim.x * blockDim.x) {
    const long long row = k < nuniq ? (long long)(keys[k] >> 32) : nrows;
    const long long prev = k > 0 ? (long long)(keys[k - 1] >> 32) : -1;
    for (long long r = prev + 1; r <= row; r++) rowptr[r] = k;
    if (k < nuniq) colidx[k] = (int)(keys[k] & 0xffffffffu);
  }
}

struct GenParams {
  int nd, nc, ntypes;
  long long nelems;
  const int* etype;
  const int* nq;
  const int* nfun;
  const int* nvert;
  const double* const* weights;
  const double* const* phi;
  const double* const* dphi;
  const double* const* gphi;
  const double* const* gdphi;
  const long long* dofoff;
  const long long* dofs;
  const long long* vertoff;
  const double* vertcoords;
  const long long* rowptr_b;
  const int* colidx_b;
  int nmat, nvec;
  const double* D;  // [nmat][nc][na][nc][na]
  const double* C;  // [nvec][nc][na]
  double* values[B2_MAX_FORMS];
  double* rhs[B2_MAX_FORMS];
};

__global__ void __launch_bounds__(128) k_assemble_tabulated(const GenParams P) {
  extern __shared__ double sm[];
  const int nd = P.nd, na = nd + 1, nc = P.nc, tid = threadIdx.x;
  for (long long e = blockIdx.x; e < P.nelems; e += gridDim.x) {
    const int t = P.etype[e], nq = P.nq[t], nf = P.nfun[t], nv = P.nvert[t];
    double* sJ = sm;                  // [nq][nd*nd] inverse Jacobians
    double* sW = sJ + nq * nd * nd;   // [nq] w |det J|
    double* sG = sW + nq;             // [nq][nf][na] value and physical gradient
    const double* X = P.vertcoords + P.vertoff[e] * nd;
    const long long* dofs = P.dofs + P.dofoff[e];
    for (int q = tid; q < nq; q += blockDim.x) {
      double J[9] = {0., 0., 0., 0., 0., 0., 0., 0., 0.};
      const double* gd = P.gdphi[t] + (long long)q * nv * nd;
      for (int v = 0; v < nv; v++)
        for (int i = 0; i < nd; i++)
          for (int k = 0; k < nd; k++) J[i * nd + k] = fma(X[v * nd + i], gd[v * nd + k], J[i * nd + k]);
      double det, inv[9];
      if (nd == 1) { det = J[0]; inv[0] = 1. / J[0]; }
      else if (nd == 2) {
        det = J[0] * J[3] - J[1] * J[2];
        const double r = 1. / det;
        inv[0] = J[3] * r; inv[1] = -J[1] * r; inv[2] = -J[2] * r; inv[3] = J[0] * r;
      } else {
        const double a0 = J[4] * J[8] - J[5] * J[7], a1 = J[5] * J[6] - J[3] * J[8], a2 = J[3] * J[7] - J[4] * J[6];
        det = J[0] * a0 + J[1] * a1 + J[2] * a2;
        const double r = 1. / det;
        inv[0] = a0 * r; inv[1] = (J[2] * J[7] - J[1] * J[8]) * r; inv[2] = (J[1] * J[5] - J[2] * J[4]) * r;
        inv[3] = a1 * r; inv[4] = (J[0] * J[8] - J[2] * J[6]) * r; inv[5] = (J[2] * J[3] - J[0] * J[5]) * r;
        inv[6] = a2 * r; inv[7] = (J[1] * J[6] - J[0] * J[7]) * r; inv[8] = (J[0] * J[4] - J[1] * J[3]) * r;
      }
      for (int k = 0; k < nd * nd; k++) sJ[q * nd * nd + k] = inv[k];
      sW[q] = P.weights[t][q] * fabs(det);
    }
    __syncthreads();
    for (int it = tid; it < nq * nf; it += blockDim.x) {
      const int q = it / nf, a = it % nf;
      double* g = sG + (long long)it * na;
      g[0] = P.phi[t][it];
      const double* dp = P.dphi[t] + (long long)it * nd;
      const double* Ji = sJ + q * nd * nd;
      for (int j = 0; j < nd; j++) {
        double s = 0.;
        for (int k = 0; k < nd; k++) s = fma(dp[k], Ji[k * nd + j], s);   // d/dx_j = sum_k d/dxi_k (J^-1)[k][j]
        g[1 + j] = s;
      }
    }
    __syncthreads();
    // matrices: one (a, b) pair per thread and pass
    for (int it = tid; it < nf * nf && P.nmat; it += blockDim.x) {
      const int a = it / nf, b = it % nf;
      const long long ra = dofs[a], cb = dofs[b];
      const long long r0 = P.rowptr_b[ra];
      int lo = 0, hi = (int)(P.rowptr_b[ra + 1] - r0);
      const int len = hi;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (P.colidx_b[r0 + mid] < cb) lo = mid + 1; else hi = mid;
      }
      for (int m = 0; m < P.nmat; m++)
        for (int c = 0; c < nc; c++)
          for (int f = 0; f < nc; f++) {
            const double* D = P.D + (((long long)m * nc + c) * na) * nc * na + (long long)f * na;   // D[m][c][x][f][y] at x stride nc*na
            bool any = false;
            for (int x = 0; x < na; x++)
              for (int y = 0; y < na; y++) any |= D[(long long)x * nc * na + y] != 0.;
            if (!any) continue;
            double s = 0.;
            for (int q = 0; q < nq; q++) {
              const double* ga = sG + ((long long)q * nf + a) * na;
              const double* gb = sG + ((long long)q * nf + b) * na;
              double u = 0.;
              for (int x = 0; x < na; x++) {
                double v = 0.;
                for (int y = 0; y < na; y++) v = fma(D[(long long)x * nc * na + y], gb[y], v);
                u = fma(ga[x], v, u);
              }
              s = fma(sW[q], u, s);
            }
            atomicAdd(P.values[m] + (r0 * nc + (long long)c * len) * nc + (long long)lo * nc + f, s);
          }
    }
    for (int a = tid; a < nf && P.nvec; a += blockDim.x)
      for (int v = 0; v < P.nvec; v++)
        for (int c = 0; c < nc; c++) {
          const double* C = P.C + ((long long)v * nc + c) * na;
          double s = 0.;
          for (int q = 0; q < nq; q++) {
            const double* ga = sG + ((long long)q * nf + a) * na;
            double u = 0.;
            for (int x = 0; x < na; x++) u = fma(C[x], ga[x], u);
            s = fma(sW[q], u, s);
          }
          if (s != 0.) atomicAdd(P.rhs[v] + dofs[a] * nc + c, s);
        }
    __syncthreads();
  }
}

template <class T>
int to_device(b2_ctx* ctx, const T* h, size_t n, T** d, std::vector<void*>& owned) {
  *d = nullptr;
  B2_CUDA(ctx, cudaMalloc((void**)d, std::max<size_t>(n, 1) * sizeof(T)));
  owned.push_back(*d);
  if (n) B2_CUDA(ctx, cudaMemcpyAsync(*d, h, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  return B2_OK;
}

struct Owned {
  std::vector<void*> p;
  ~Owned() { for (void* q : p) cudaFree(q); }
};

}  // namespace

extern "C" int b2_pattern_create_csr(b2_ctx* ctx, int64_t nrows, int64_t ncols, const int64_t* rowptr, const int64_t* colidx, b2_pattern** out) {
  if (!ctx || !rowptr || !out || nrows < 0 || ncols < 0 || ncols > 0x7fffffffLL) return b2_fail(ctx, B2_EINVAL, "invalid argument");
  *out = nullptr;
  const int64_t nnz = rowptr[nrows];
  if (rowptr[0] != 0 || nnz < 0 || (nnz && !colidx)) return b2_fail(ctx, B2_EINVAL, "invalid rowptr");
  B2_CUDA(ctx, cudaSetDevice(ctx->device));
  b2_pattern* p = new (std::nothrow) b2_pattern();
  if (!p) return B2_ENOMEM;
  p->ctx = ctx;
  p->basis = nullptr;
  p->nrows = nrows;
  p->ncols = ncols;
  p->nnz = nnz;
  p->ncomp_gen = 1;
  p->rowptr_b.assign(rowptr, rowptr + nrows + 1);
  std::vector<int> ci((size_t)nnz);
  for (int64_t k = 0; k < nnz; k++) {
    if (colidx[k] < 0 || colidx[k] >= ncols) { delete p; return b2_fail(ctx, B2_EINVAL, "column index out of bounds"); }
    ci[(size_t)k] = (int)colidx[k];
  }
  cudaError_t e = cudaMalloc((void**)&p->d_rowptr_b, sizeof(long long) * (size_t)(nrows + 1));
  if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_colidx_b, sizeof(int) * (size_t)std::max<int64_t>(nnz, 1));
  if (e == cudaSuccess) e = cudaMemcpy(p->d_rowptr_b, rowptr, sizeof(long long) * (size_t)(nrows + 1), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && nnz) e = cudaMemcpy(p->d_colidx_b, ci.data(), sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { b2_pattern_destroy(p); return b2_cuda_fail(ctx, e, "pattern upload"); }
  *out = p;
  return B2_OK;
}

extern "C" int b2_pattern_general(b2_ctx* ctx, int64_t nelems, const int64_t* rowoff, const int64_t* rowdofs, const int64_t* coloff, const int64_t* coldofs,
                                  int64_t nrows, int64_t ncols, int ncomp, b2_pattern** out) {
  if (!ctx || !rowoff || !coloff || !out || nelems < 0 || nrows < 0 || ncols < 0 || ncomp < 1) return b2_fail(ctx, B2_EINVAL, "invalid argument");
  if (nrows > 0x7fffffffLL || ncols > 0x7fffffffLL) return b2_fail(ctx, B2_EUNSUPPORTED, "more than 2^31 basis functions");
  *out = nullptr;
  B2_CUDA(ctx, cudaSetDevice(ctx->device));
  Owned own;
  long long *d_ro, *d_rd, *d_co, *d_cd, *d_keyoff;
  int rc;
  if ((rc = to_device(ctx, (const long long*)rowoff, (size_t)nelems + 1, &d_ro, own.p)) != B2_OK) return rc;
  if ((rc = to_device(ctx, (const long long*)rowdofs, (size_t)rowoff[nelems], &d_rd, own.p)) != B2_OK) return rc;
  if ((rc = to_device(ctx, (const long long*)coloff, (size_t)nelems + 1, &d_co, own.p)) != B2_OK) return rc;
  if ((rc = to_device(ctx, (const long long*)coldofs, (size_t)coloff[nelems], &d_cd, own.p)) != B2_OK) return rc;
  for (int64_t k = 0; k < rowoff[nelems]; k++)
    if (rowdofs[k] < 0 || rowdofs[k] >= nrows) return b2_fail(ctx, B2_EINVAL, "row dof out of bounds");
  for (int64_t k = 0; k < coloff[nelems]; k++)
    if (coldofs[k] < 0 || coldofs[k] >= ncols) return b2_fail(ctx, B2_EINVAL, "column dof out of bounds");
  B2_CUDA(ctx, cudaMalloc((void**)&d_keyoff, sizeof(long long) * (size_t)(nelems + 1)));
  own.p.push_back(d_keyoff);
  const int blocks = (int)std::min<long long>(std::max<long long>((nelems + 7) / 8, 1), (long long)ctx->sm_count * 16);
  k_general_count<<<blocks, 256, 0, ctx->stream>>>(nelems, d_ro, d_co, d_keyoff);
  ctx->launches++;
  if ((rc = launch_exclusive_scan(ctx, d_keyoff, nelems)) != B2_OK) return rc;
  long long nkeys = 0;
  B2_CUDA(ctx, cudaMemcpy(&nkeys, d_keyoff + nelems, sizeof(long long), cudaMemcpyDeviceToHost));
  unsigned long long *d_keys, *d_sorted, *d_uniq;
  long long* d_nuniq;
  B2_CUDA(ctx, cudaMalloc((void**)&d_keys, sizeof(unsigned long long) * (size_t)std::max<long long>(nkeys, 1)));
  own.p.push_back(d_keys);
  B2_CUDA(ctx, cudaMalloc((void**)&d_sorted, sizeof(unsigned long long) * (size_t)std::max<long long>(nkeys, 1)));
  own.p.push_back(d_sorted);
  B2_CUDA(ctx, cudaMalloc((void**)&d_uniq, sizeof(unsigned long long) * (size_t)std::max<long long>(nkeys, 1)));
  own.p.push_back(d_uniq);
  B2_CUDA(ctx, cudaMalloc((void**)&d_nuniq, sizeof(long long)));
  own.p.push_back(d_nuniq);
  k_general_keys<<<blocks, 256, 0, ctx->stream>>>(nelems, d_ro, d_rd, d_co, d_cd, d_keyoff, d_keys);
  ctx->launches++;
  B2_CUDA(ctx, cudaGetLastError());
  long long nuniq = 0;
  if (nkeys) {
    if (nkeys > 0x7fffffffLL) return b2_fail(ctx, B2_EUNSUPPORTED, "more than 2^31 element block entries");
    size_t tmp1 = 0, tmp2 = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, tmp1, d_keys, d_sorted, (int)nkeys, 0, 64, ctx->stream);
    cub::DeviceSelect::Unique(nullptr, tmp2, d_sorted, d_uniq, d_nuniq, (int)nkeys, ctx->stream);
    void* d_tmp;
    B2_CUDA(ctx, cudaMalloc(&d_tmp, std::max(tmp1, tmp2)));
    own.p.push_back(d_tmp);
    B2_CUDA(ctx, cub::DeviceRadixSort::SortKeys(d_tmp, tmp1, d_keys, d_sorted, (int)nkeys, 0, 64, ctx->stream));
    B2_CUDA(ctx, cub::DeviceSelect::Unique(d_tmp, tmp2, d_sorted, d_uniq, d_nuniq, (int)nkeys, ctx->stream));
    ctx->launches += 2;
    B2_CUDA(ctx, cudaMemcpyAsync(&nuniq, d_nuniq, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    B2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  b2_pattern* p = new (std::nothrow) b2_pattern();
  if (!p) return B2_ENOMEM;
  p->ctx = ctx;
  p->basis = nullptr;
  p->nrows = nrows * ncomp;
  p->ncols = ncols * ncomp;
  p->nnz = nuniq * ncomp * ncomp;
  p->ncomp_gen = ncomp;
  cudaError_t e = cudaMalloc((void**)&p->d_rowptr_b, sizeof(long long) * (size_t)(nrows + 1));
  if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_colidx_b, sizeof(int) * (size_t)std::max<long long>(nuniq, 1));
  if (e != cudaSuccess) { b2_pattern_destroy(p); return b2_cuda_fail(ctx, e, "pattern allocation"); }
  const int sblocks = (int)std::min<long long>(std::max<long long>((nuniq + 256) / 256, 1), (long long)ctx->sm_count * 16);
  k_general_split<<<sblocks, 256, 0, ctx->stream>>>(nuniq, d_uniq, nrows, p->d_rowptr_b, p->d_colidx_b);
  ctx->launches++;
  p->rowptr_b.resize((size_t)nrows + 1);
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(p->rowptr_b.data(), p->d_rowptr_b, sizeof(long long) * (size_t)(nrows + 1), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) { b2_pattern_destroy(p); return b2_cuda_fail(ctx, e, "pattern construction"); }
  *out = p;
  return B2_OK;
}

extern "C" int b2_assemble_general_host(b2_ctx* ctx, const b2_pattern* pattern, int ndims, int ncomp, int64_t nelems, int ntypes, const int32_t* etype,
                                        const int32_t* nq, const int32_t* nfun, const int32_t* nvert, const double* const* weights, const double* const* phi,
                                        const double* const* dphi, const double* const* gphi, const double* const* gdphi, const int64_t* dofoff, const int64_t* dofs,
                                        const int64_t* vertoff, const double* vertcoords, int nmat, const double* const* D_host, double* const* values_host, int nvec,
                                        const double* const* C_host, double* const* rhs_host) {
  if (!ctx || !pattern || pattern->basis || !pattern->d_rowptr_b) return b2_fail(ctx, B2_EINVAL, "a general pattern (b2_pattern_general) is needed");
  if (ndims < 1 || ndims > 3 || ncomp != pattern->ncomp_gen || nelems < 0 || ntypes < 1) return b2_fail(ctx, B2_EINVAL, "invalid argument");
  if (nmat < 0 || nmat > B2_MAX_FORMS || nvec < 0 || nvec > B2_MAX_FORMS) return b2_fail(ctx, B2_EINVAL, "0..4 matrix and vector forms per call");
  (void)gphi;
  B2_CUDA(ctx, cudaSetDevice(ctx->device));
  const int na = ndims + 1;
  Owned own;
  int rc;
  GenParams P;
  P.nd = ndims; P.nc = ncomp; P.ntypes = ntypes; P.nelems = nelems;
  P.nmat = nmat; P.nvec = nvec;
  P.rowptr_b = pattern->d_rowptr_b;
  P.colidx_b = pattern->d_colidx_b;
  int *d_etype, *d_nq, *d_nf, *d_nv;
  if ((rc = to_device(ctx, (const int*)etype, (size_t)nelems, &d_etype, own.p)) != B2_OK) return rc;
  if ((rc = to_device(ctx, (const int*)nq, (size_t)ntypes, &d_nq, own.p)) != B2_OK) return rc;
  if ((rc = to_device(ctx, (const int*)nfun, (size_t)ntypes, &d_nf, own.p)) != B2_OK) return rc;
  if ((rc = to_device(ctx, (const int*)nvert, (size_t)ntypes, &d_nv, own.p)) != B2_OK) return rc;
  P.etype = d_etype; P.nq = d_nq; P.nfun = d_nf; P.nvert = d_nv;
  std::vector<const double*> hw(ntypes), hp(ntypes), hd(ntypes), hgd(ntypes);
  size_t smem = 0;
  for (int t = 0; t < ntypes; t++) {
    double* d;
    if ((rc = to_device(ctx, weights[t], (size_t)nq[t], &d, own.p)) != B2_OK) return rc;
    hw[t] = d;
    if ((rc = to_device(ctx, phi[t], (size_t)nq[t] * nfun[t], &d, own.p)) != B2_OK) return rc;
    hp[t] = d;
    if ((rc = to_device(ctx, dphi[t], (size_t)nq[t] * nfun[t] * ndims, &d, own.p)) != B2_OK) return rc;
    hd[t] = d;
    if ((rc = to_device(ctx, gdphi[t], (size_t)nq[t] * nvert[t] * ndims, &d, own.p)) != B2_OK) return rc;
    hgd[t] = d;
    smem = std::max(smem, sizeof(double) * ((size_t)nq[t] * ndims * ndims + nq[t] + (size_t)nq[t] * nfun[t] * na));
  }
  if (smem > 200 * 1024) return b2_fail(ctx, B2_EUNSUPPORTED, "element tables exceed shared memory");
  const double** dd;
  if ((rc = to_device(ctx, hw.data(), (size_t)ntypes, &dd, own.p)) != B2_OK) return rc;
  P.weights = dd;
  if ((rc = to_device(ctx, hp.data(), (size_t)ntypes, &dd, own.p)) != B2_OK) return rc;
  P.phi = dd;
  if ((rc = to_device(ctx, hd.data(), (size_t)ntypes, &dd, own.p)) != B2_OK) return rc;
  P.dphi = dd;
  if ((rc = to_device(ctx, hgd.data(), (size_t)ntypes, &dd, own.p)) != B2_OK) return rc;
  P.gdphi = dd;
  P.gphi = nullptr;
  long long *d_dofoff, *d_dofs, *d_vertoff;
  double* d_vc;
  if ((rc = to_device(ctx, (const long long*)dofoff, (size_t)nelems + 1, &d_dofoff, own.p)) != B2_OK) return rc;
  if ((rc = to_device(ctx, (const long long*)dofs, (size_t)dofoff[nelems], &d_dofs, own.p)) != B2_OK) return rc;
  if ((rc = to_device(ctx, (const long long*)vertoff, (size_t)nelems + 1, &d_vertoff, own.p)) != B2_OK) return rc;
  if ((rc = to_device(ctx, vertcoords, (size_t)vertoff[nelems] * ndims, &d_vc, own.p)) != B2_OK) return rc;
  P.dofoff = d_dofoff; P.dofs = d_dofs; P.vertoff = d_vertoff; P.vertcoords = d_vc;
  std::vector<double> hD((size_t)nmat * ncomp * na * ncomp * na), hC((size_t)nvec * ncomp * na);
  for (int m = 0; m < nmat; m++) std::copy(D_host[m], D_host[m] + (size_t)ncomp * na * ncomp * na, hD.begin() + (size_t)m * ncomp * na * ncomp * na);
  for (int v = 0; v < nvec; v++) std::copy(C_host[v], C_host[v] + (size_t)ncomp * na, hC.begin() + (size_t)v * ncomp * na);
  double *dD, *dC;
  if ((rc = to_device(ctx, hD.data(), hD.size(), &dD, own.p)) != B2_OK) return rc;
  if ((rc = to_device(ctx, hC.data(), hC.size(), &dC, own.p)) != B2_OK) return rc;
  P.D = dD; P.C = dC;
  for (int m = 0; m < nmat; m++) {
    B2_CUDA(ctx, cudaMalloc((void**)&P.values[m], sizeof(double) * (size_t)std::max<int64_t>(pattern->nnz, 1)));
    own.p.push_back(P.values[m]);
    B2_CUDA(ctx, cudaMemsetAsync(P.values[m], 0, sizeof(double) * (size_t)pattern->nnz, ctx->stream));
  }
  for (int v = 0; v < nvec; v++) {
    B2_CUDA(ctx, cudaMalloc((void**)&P.rhs[v], sizeof(double) * (size_t)std::max<int64_t>(pattern->nrows, 1)));
    own.p.push_back(P.rhs[v]);
    B2_CUDA(ctx, cudaMemsetAsync(P.rhs[v], 0, sizeof(double) * (size_t)pattern->nrows, ctx->stream));
  }
  if (nelems) {
    B2_CUDA(ctx, cudaFuncSetAttribute(k_assemble_tabulated, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int blocks = (int)std::min<long long>(nelems, (long long)ctx->sm_count * 8);
    {
      KernelTimer timer(ctx);
      k_assemble_tabulated<<<blocks, 128, smem, ctx->stream>>>(P);
    }
    ctx->launches++;
    B2_CUDA(ctx, cudaGetLastError());
  }
  for (int m = 0; m < nmat; m++) B2_CUDA(ctx, cudaMemcpyAsync(values_host[m], P.values[m], sizeof(double) * (size_t)pattern->nnz, cudaMemcpyDeviceToHost, ctx->stream));
  for (int v = 0; v < nvec; v++) B2_CUDA(ctx, cudaMemcpyAsync(rhs_host[v], P.rhs[v], sizeof(double) * (size_t)pattern->nrows, cudaMemcpyDeviceToHost, ctx->stream));
  B2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return B2_OK;
}
