// Host side of the C ABI declared in include/b200fem.h: object lifetime, validation, table
// preparation, kernel dispatch.  No torch, no Python: plain CUDA runtime.

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>

#include "common.cuh"

// ---- errors ----------------------------------------------------------------------------------------

extern "C" const char* b2_strerror(int code) {
  switch (code) {
    case B2_OK: return "ok";
    case B2_EINVAL: return "invalid argument";
    case B2_ENOMEM: return "out of memory";
    case B2_ECUDA: return "CUDA runtime error";
    case B2_ENODEV: return "no usable CUDA device";
    case B2_EUNSUPPORTED: return "unsupported configuration";
    case B2_ENOTCONVERGED: return "tolerance not reached";
    default: return "unknown error";
  }
}

int b2_fail(b2_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return code;
}

int b2_cuda_fail(b2_ctx* ctx, cudaError_t e, const char* what) {
  if (ctx) ctx->err = std::string(what) + ": " + cudaGetErrorString(e);
  cudaGetLastError();  // clear the sticky-free error state
  return e == cudaErrorMemoryAllocation ? B2_ENOMEM : B2_ECUDA;
}

extern "C" const char* b2_last_error(const b2_ctx* ctx) { return ctx ? ctx->err.c_str() : ""; }
extern "C" int b2_version(void) { return 100; }

// ---- context ---------------------------------------------------------------------------------------

extern "C" int b2_device_count(int* count) {
  if (!count) return B2_EINVAL;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    *count = 0;
    return B2_ENODEV;
  }
  *count = n;
  return B2_OK;
}

extern "C" int b2_ctx_create(int device, b2_ctx** out) {
  if (!out) return B2_EINVAL;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return B2_ENODEV;
  }
  if (device < 0 || device >= n) return B2_EINVAL;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return B2_ECUDA;
  if (prop.major != 10) return B2_ENODEV;  // the library only carries sm_100a code
  b2_ctx* ctx = new (std::nothrow) b2_ctx();
  if (!ctx) return B2_ENOMEM;
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess) {
    cudaGetLastError();
    delete ctx;
    return B2_ECUDA;
  }
  ctx->stream = ctx->own_stream;
  *out = ctx;
  return B2_OK;
}

extern "C" int b2_ctx_destroy(b2_ctx* ctx) {
  if (!ctx) return B2_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->flush_buf) cudaFree(ctx->flush_buf);
  if (ctx->scratch) cudaFree(ctx->scratch);
  if (ctx->gbuf) cudaFree(ctx->gbuf);
  if (ctx->queue) cudaFree(ctx->queue);
  if (ctx->formbuf) cudaFree(ctx->formbuf);
  if (ctx->copy_stream) { cudaStreamDestroy(ctx->copy_stream); cudaEventDestroy(ctx->copy_event); }
  for (auto* v : {&ctx->kernel_events, &ctx->event_pool})
    for (auto& ev : *v) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
  cudaEventDestroy(ctx->ev0);
  cudaEventDestroy(ctx->ev1);
  cudaStreamDestroy(ctx->own_stream);
  delete ctx;
  return B2_OK;
}

extern "C" int b2_ctx_set_stream(b2_ctx* ctx, void* s) {
  if (!ctx) return B2_EINVAL;
  ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
  return B2_OK;
}

extern "C" int b2_ctx_synchronize(b2_ctx* ctx) {
  if (!ctx) return B2_EINVAL;
  B2_CUDA(ctx, cudaSetDevice(ctx->device));
  B2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return B2_OK;
}

extern "C" int b2_ctx_timer_start(b2_ctx* ctx) {
  if (!ctx) return B2_EINVAL;
  B2_CUDA(ctx, cudaSetDevice(ctx->device));
  B2_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  return B2_OK;
}

extern "C" int b2_ctx_timer_stop(b2_ctx* ctx, float* ms) {
  if (!ctx || !ms) return B2_EINVAL;
  B2_CUDA(ctx, cudaSetDevice(ctx->device));
  B2_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  B2_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
  B2_CUDA(ctx, cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
  return B2_OK;
}

extern "C" int64_t b2_ctx_launch_count(const b2_ctx* ctx) { return ctx ? ctx->launches : 0; }

KernelTimer::KernelTimer(b2_ctx* c) : ctx(c) {
  auto it = ctx->opts.find("time_kernels");
  if (it == ctx->opts.end() || !it->second) return;
  if (!ctx->event_pool.empty()) {
    e0 = ctx->event_pool.back().first;
    e1 = ctx->event_pool.back().second;
    ctx->event_pool.pop_back();
  } else if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) {
    cudaGetLastError();
    e0 = e1 = nullptr;
    return;
  }
  cudaEventRecord(e0, ctx->stream);
}

KernelTimer::~KernelTimer() {
  if (!e0) return;
  cudaEventRecord(e1, ctx->stream);
  ctx->kernel_events.emplace_back(e0, e1);
}

extern "C" int b2_ctx_kernel_time(b2_ctx* ctx, double* ms, int64_t* launches) {
  if (!ctx || !ms || !launches) return B2_EINVAL;
  B2_CUDA(ctx, cudaSetDevice(ctx->device));
  B2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  // option time_kernels = 2: the longest single launch instead of the sum (the dominant kernel of a multi-kernel step)
  const bool want_max = ctx->opts.count("time_kernels") && ctx->opts["time_kernels"] == 2;
  double total = 0.;
  for (auto& ev : ctx->kernel_events) {
    float t = 0.f;
    B2_CUDA(ctx, cudaEventElapsedTime(&t, ev.first, ev.second));
    total = want_max ? std::max<double>(total, t) : total + t;
  }
  *ms = total;
  *launches = (int64_t)ctx->kernel_events.size();
  ctx->event_pool.insert(ctx->event_pool.end(), ctx->kernel_events.begin(), ctx->kernel_events.end());
  ctx->kernel_events.clear();
  return B2_OK;
}

extern "C" int b2_ctx_set_option(b2_ctx* ctx, const char* name, int64_t value) {
  if (!ctx || !name) return B2_EINVAL;
  ctx->opts[name] = value;
  return B2_OK;
}

extern "C" int b2_host_alloc(b2_ctx* ctx, int64_t nbytes, void** out) {
  if (!ctx || !out || nbytes < 0) return B2_EINVAL;
  B2_CUDA(ctx, cudaSetDevice(ctx->device));
  B2_CUDA(ctx, cudaHostAlloc(out, (size_t)std::max<int64_t>(nbytes, 1), cudaHostAllocDefault));
  return B2_OK;
}

extern "C" int b2_host_free(b2_ctx* ctx, void* p) {
  if (!ctx) return B2_EINVAL;
  if (p) B2_CUDA(ctx, cudaFreeHost(p));
  return B2_OK;
}

extern "C" int b2_device_alloc(b2_ctx* ctx, int64_t nbytes, void** out) {
  if (!ctx || !out || nbytes < 0) return B2_EINVAL;
  B2_CUDA(ctx, cudaSetDevice(ctx->device));
  B2_CUDA(ctx, cudaMalloc(out, (size_t)std::max<int64_t>(nbytes, 1)));
  return B2_OK;
}

extern "C" int b2_device_free(b2_ctx* ctx, void* p) {
  if (!ctx) return B2_EINVAL;
  B2_CUDA(ctx, cudaSetDevice(ctx->device));
  if (p) B2_CUDA(ctx, cudaFree(p));
  return B2_OK;
}

extern "C" int b2_memcpy_d2h(b2_ctx* ctx, void* dst, const void* src, int64_t nbytes) {
  if (!ctx || nbytes < 0 || (nbytes && (!dst || !src))) return B2_EINVAL;
  B2_CUDA(ctx, cudaSetDevice(ctx->device));
  B2_CUDA(ctx, cudaMemcpyAsync(dst, src, (size_t)nbytes, cudaMemcpyDeviceToHost, ctx->stream));
  B2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return B2_OK;
}

extern "C" int b2_memcpy_h2d(b2_ctx* ctx, void* dst, const void* src, int64_t nbytes) {
  if (!ctx || nbytes < 0 || (nbytes && (!dst || !src))) return B2_EINVAL;
  B2_CUDA(ctx, cudaSetDevice(ctx->device));
  B2_CUDA(ctx, cudaMemcpyAsync(dst, src, (size_t)nbytes, cudaMemcpyHostToDevice, ctx->stream));
  B2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return B2_OK;
}

extern "C" int b2_memset_zero(b2_ctx* ctx, void* dev, int64_t nbytes) {
  if (!ctx || nbytes < 0 || (nbytes && !dev)) return B2_EINVAL;
  B2_CUDA(ctx, cudaSetDevice(ctx->device));
  B2_CUDA(ctx, cudaMemsetAsync(dev, 0, (size_t)nbytes, ctx->stream));
  return B2_OK;
}

extern "C" int b2_flush_l2(b2_ctx* ctx) {
  if (!ctx) return B2_EINVAL;
  B2_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!ctx->flush_buf) {
    ctx->flush_bytes = (size_t)256 << 20;  // 256 MiB > 126 MB L2
    B2_CUDA(ctx, cudaMalloc(&ctx->flush_buf, ctx->flush_bytes));
  }
  B2_CUDA(ctx, cudaMemsetAsync(ctx->flush_buf, 0x5a, ctx->flush_bytes, ctx->stream));
  return B2_OK;
}

// ---- small upload helper ---------------------------------------------------------------------------

template <class T>
static int upload(b2_ctx* ctx, const std::vector<T>& h, T** d) {
  *d = nullptr;
  B2_CUDA(ctx, cudaMalloc((void**)d, std::max<size_t>(h.size(), 1) * sizeof(T)));
  if (!h.empty()) B2_CUDA(ctx, cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return B2_OK;
}

// ---- basis -----------------------------------------------------------------------------------------

BasisView b2_basis::view() const {
  BasisView v;
  memset(&v, 0, sizeof(v));
  v.ndims = ndims;
  v.ncomp = ncomp;
  v.nb = 1;
  for (int d = 0; d < ndims; d++) {
    v.p[d] = p[d];
    v.nel[d] = (int)nel[d];
    v.ndofs[d] = (int)ndofs_d[d];
    v.setidx[d] = d_setidx[d];
    v.start[d] = d_start[d];
    v.lo[d] = d_lo[d];
    v.wid[d] = d_wid[d];
    v.cum[d] = d_cum[d];
    v.W[d] = W[d];
    v.nsets[d] = nsets[d];
    v.nb *= p[d] + 1;
  }
  return v;
}

extern "C" int b2_basis_create(b2_ctx* ctx, int ndims, const int64_t* nelems, const int32_t* degree, const int32_t* nsets,
                               const double* const* coeffs, const int32_t* const* setidx, const int64_t* const* start,
                               const int64_t* ndofs, int ncomp, b2_basis** out) {
  if (!ctx || !out || !nelems || !degree || !nsets || !coeffs || !setidx || !start || !ndofs) return b2_fail(ctx, B2_EINVAL, "null argument");
  *out = nullptr;
  if (ndims < 1 || ndims > B2_MAXD) return b2_fail(ctx, B2_EINVAL, "ndims must be 1..3");
  if (ncomp < 1 || ncomp > 3) return b2_fail(ctx, B2_EINVAL, "ncomp must be 1..3");
  B2_CUDA(ctx, cudaSetDevice(ctx->device));
  b2_basis* b = new (std::nothrow) b2_basis();
  if (!b) return B2_ENOMEM;
  b->ctx = ctx;
  b->ndims = ndims;
  b->ncomp = ncomp;
  b->nbasis = 1;
  double nelems_total = 1;
  for (int d = 0; d < B2_MAXD; d++) b->d_setidx[d] = b->d_start[d] = b->d_lo[d] = b->d_wid[d] = b->d_cum[d] = nullptr;
  for (int d = 0; d < ndims; d++) {
    const int p = degree[d];
    const int64_t n = nelems[d];
    if (p < 0 || p > B2_MAX_DEGREE) { delete b; return b2_fail(ctx, B2_EUNSUPPORTED, "degree must be 0..4"); }
    if (n < 1 || nsets[d] < 1 || ndofs[d] < 1 || ndofs[d] > (1 << 30) || n > (1 << 30)) { delete b; return b2_fail(ctx, B2_EINVAL, "invalid sizes"); }
    b->p[d] = p;
    b->nel[d] = n;
    b->ndofs_d[d] = ndofs[d];
    b->nsets[d] = nsets[d];
    b->coeffs[d].assign(coeffs[d], coeffs[d] + (size_t)nsets[d] * (p + 1) * (p + 1));
    b->setidx[d].resize(n);
    b->start[d].resize(n);
    for (int64_t e = 0; e < n; e++) {
      if (setidx[d][e] < 0 || setidx[d][e] >= nsets[d]) { delete b; return b2_fail(ctx, B2_EINVAL, "setidx out of range"); }
      if (start[d][e] < 0 || start[d][e] + p >= ndofs[d]) { delete b; return b2_fail(ctx, B2_EUNSUPPORTED, "element dofs out of range (periodic bases are not supported)"); }
      if (e && start[d][e] < start[d][e - 1]) { delete b; return b2_fail(ctx, B2_EINVAL, "start dofs must be non-decreasing"); }
      b->setidx[d][e] = setidx[d][e];
      b->start[d][e] = (int)start[d][e];
    }
    // column structure: dof i couples with the union of the dof ranges of its supporting elements
    const int nd = (int)ndofs[d];
    b->lo[d].assign(nd, nd);
    std::vector<int> hi(nd, 0);
    for (int64_t e = 0; e < n; e++)
      for (int a = 0; a <= p; a++) {
        int i = b->start[d][e] + a;
        b->lo[d][i] = std::min(b->lo[d][i], b->start[d][e]);
        hi[i] = std::max(hi[i], b->start[d][e] + p + 1);
      }
    b->wid[d].resize(nd);
    b->cum[d].resize(nd + 1);
    long long s = 0;
    for (int i = 0; i < nd; i++) {
      if (hi[i] <= b->lo[d][i]) { b->lo[d][i] = 0; hi[i] = 0; }  // dof without support: empty row
      b->wid[d][i] = hi[i] - b->lo[d][i];
      b->cum[d][i] = (int)s;
      s += b->wid[d][i];
      if (s > 0x7fffffffLL) { delete b; return b2_fail(ctx, B2_EUNSUPPORTED, "dimension too large"); }
    }
    b->cum[d][nd] = (int)s;
    b->W[d] = s;
    b->nbasis *= ndofs[d];
    nelems_total *= (double)n;
  }
  if ((double)b->nbasis * ncomp > 2.0e9 || nelems_total > 2.0e9) { delete b; return b2_fail(ctx, B2_EUNSUPPORTED, "more than 2^31 dofs or elements"); }
  int rc = B2_OK;
  for (int d = 0; d < ndims && rc == B2_OK; d++) {
    rc = upload(ctx, b->setidx[d], &b->d_setidx[d]);
    if (rc == B2_OK) rc = upload(ctx, b->start[d], &b->d_start[d]);
    if (rc == B2_OK) rc = upload(ctx, b->lo[d], &b->d_lo[d]);
    if (rc == B2_OK) rc = upload(ctx, b->wid[d], &b->d_wid[d]);
    if (rc == B2_OK) rc = upload(ctx, b->cum[d], &b->d_cum[d]);
  }
  if (rc != B2_OK) { b2_basis_destroy(b); return rc; }
  *out = b;
  return B2_OK;
}

extern "C" int b2_basis_destroy(b2_basis* b) {
  if (!b) return B2_OK;
  cudaSetDevice(b->ctx->device);
  for (int d = 0; d < B2_MAXD; d++) {
    if (b->d_setidx[d]) cudaFree(b->d_setidx[d]);
    if (b->d_start[d]) cudaFree(b->d_start[d]);
    if (b->d_lo[d]) cudaFree(b->d_lo[d]);
    if (b->d_wid[d]) cudaFree(b->d_wid[d]);
    if (b->d_cum[d]) cudaFree(b->d_cum[d]);
  }
  for (auto& kv : b->tabs)
    for (int d = 0; d < B2_MAXD; d++)
      if (kv.second.tab[d]) cudaFree(kv.second.tab[d]);
  delete b;
  return B2_OK;
}

extern "C" int64_t b2_basis_ndofs(const b2_basis* b) { return b ? b->nbasis * b->ncomp : 0; }

// ---- quadrature ------------------------------------------------------------------------------------

extern "C" int b2_quad_create_tensor(b2_ctx* ctx, int ndims, const int32_t* nq, const double* const* pts, const double* const* wts, b2_quad** out) {
  if (!ctx || !out || !nq || !pts || !wts) return b2_fail(ctx, B2_EINVAL, "null argument");
  *out = nullptr;
  if (ndims < 1 || ndims > B2_MAXD) return b2_fail(ctx, B2_EINVAL, "ndims must be 1..3");
  B2_CUDA(ctx, cudaSetDevice(ctx->device));
  b2_quad* q = new (std::nothrow) b2_quad();
  if (!q) return B2_ENOMEM;
  q->ctx = ctx;
  q->serial = ++ctx->serial;
  q->ndims = ndims;
  for (int d = 0; d < B2_MAXD; d++) q->d_x[d] = q->d_w[d] = nullptr;
  for (int d = 0; d < ndims; d++) {
    if (nq[d] < 1 || nq[d] > 16) { delete q; return b2_fail(ctx, B2_EUNSUPPORTED, "1..16 quadrature points per dimension"); }
    q->nq[d] = nq[d];
    q->pts[d].assign(pts[d], pts[d] + nq[d]);
    q->wts[d].assign(wts[d], wts[d] + nq[d]);
  }
  int rc = B2_OK;
  for (int d = 0; d < ndims && rc == B2_OK; d++) {
    rc = upload(ctx, q->pts[d], &q->d_x[d]);
    if (rc == B2_OK) rc = upload(ctx, q->wts[d], &q->d_w[d]);
  }
  if (rc != B2_OK) { b2_quad_destroy(q); return rc; }
  *out = q;
  return B2_OK;
}

extern "C" int b2_quad_destroy(b2_quad* q) {
  if (!q) return B2_OK;
  cudaSetDevice(q->ctx->device);
  for (int d = 0; d < B2_MAXD; d++) {
    if (q->d_x[d]) cudaFree(q->d_x[d]);
    if (q->d_w[d]) cudaFree(q->d_w[d]);
  }
  delete q;
  return B2_OK;
}

// ---- geometry --------------------------------------------------------------------------------------

extern "C" int b2_geom_create_nodal(b2_ctx* ctx, int ndims, const int64_t* nelems, const double* nodes_host, b2_geom** out) {
  if (!ctx || !out || !nelems || !nodes_host) return b2_fail(ctx, B2_EINVAL, "null argument");
  *out = nullptr;
  if (ndims < 1 || ndims > B2_MAXD) return b2_fail(ctx, B2_EINVAL, "ndims must be 1..3");
  B2_CUDA(ctx, cudaSetDevice(ctx->device));
  b2_geom* g = new (std::nothrow) b2_geom();
  if (!g) return B2_ENOMEM;
  g->ctx = ctx;
  g->ndims = ndims;
  g->nnodes = 1;
  g->d_nodes = nullptr;
  for (int d = 0; d < ndims; d++) {
    if (nelems[d] < 1) { delete g; return b2_fail(ctx, B2_EINVAL, "invalid element count"); }
    g->nel[d] = nelems[d];
    g->nnodes *= nelems[d] + 1;
  }
  cudaError_t e = cudaMalloc((void**)&g->d_nodes, sizeof(double) * g->nnodes * ndims);
  if (e != cudaSuccess) { delete g; return b2_cuda_fail(ctx, e, "cudaMalloc(nodes)"); }
  int rc = b2_geom_update_nodal(g, nodes_host);
  if (rc != B2_OK) { b2_geom_destroy(g); return rc; }
  *out = g;
  return B2_OK;
}

extern "C" int b2_geom_update_nodal(b2_geom* g, const double* nodes_host) {
  if (!g || !nodes_host || g->gbasis) return B2_EINVAL;
  b2_ctx* ctx = g->ctx;
  B2_CUDA(ctx, cudaMemcpyAsync(g->d_nodes, nodes_host, sizeof(double) * g->nnodes * g->ndims, cudaMemcpyHostToDevice, ctx->stream));
  return B2_OK;
}

extern "C" int b2_geom_destroy(b2_geom* g) {
  if (!g) return B2_OK;
  cudaSetDevice(g->ctx->device);
  if (g->d_nodes) cudaFree(g->d_nodes);
  if (g->d_ctrl) cudaFree(g->d_ctrl);
  if (g->d_wts) cudaFree(g->d_wts);
  for (int d = 0; d < B2_MAXD; d++)
    if (g->d_gcoeffs[d]) cudaFree(g->d_gcoeffs[d]);
  delete g;
  return B2_OK;
}

// ---- pattern ---------------------------------------------------------------------------------------

extern "C" int b2_pattern_create(b2_ctx* ctx, const b2_basis* basis, b2_pattern** out) {
  if (!ctx || !basis || !out) return b2_fail(ctx, B2_EINVAL, "null argument");
  b2_pattern* p = new (std::nothrow) b2_pattern();
  if (!p) return B2_ENOMEM;
  p->ctx = ctx;
  p->basis = basis;
  long long n = 1;
  for (int d = 0; d < basis->ndims; d++) n *= basis->W[d];
  p->nnz = n * basis->ncomp * basis->ncomp;
  p->nrows = basis->nbasis * basis->ncomp;
  *out = p;
  return B2_OK;
}

extern "C" int b2_pattern_destroy(b2_pattern* p) {
  if (!p) return B2_OK;
  if (p->aux_pattern) b2_pattern_destroy(p->aux_pattern);
  if (p->aux_elemset) b2_elemset_destroy(p->aux_elemset);
  if (p->d_rowptr_b || p->d_colidx_b) {
    cudaSetDevice(p->ctx->device);
    if (p->d_rowptr_b) cudaFree(p->d_rowptr_b);
    if (p->d_colidx_b) cudaFree(p->d_colidx_b);
  }
  delete p;
  return B2_OK;
}

extern "C" int64_t b2_pattern_nnz(const b2_pattern* p) { return p ? p->nnz : 0; }
extern "C" int64_t b2_pattern_nrows(const b2_pattern* p) { return p ? p->nrows : 0; }

extern "C" int64_t b2_pattern_row_offset(const b2_pattern* p, int64_t row) {
  if (!p || row < 0 || row > p->nrows) return -1;
  if (row == p->nrows) return p->nnz;
  const b2_basis* b = p->basis;
  const int nc = pattern_ncomp(p);
  if (pattern_general(p)) {
    const int64_t In = row / nc, c = row % nc;
    const long long r0 = p->rowptr_b[In], len = p->rowptr_b[In + 1] - r0;
    return (r0 * nc + c * len) * nc;
  }
  int64_t I = row / nc;
  const int c = (int)(row % nc);
  int i[B2_MAXD] = {0, 0, 0};
  for (int d = b->ndims - 1; d >= 0; d--) { i[d] = (int)(I % b->ndofs_d[d]); I /= b->ndofs_d[d]; }
  long long r = b->cum[0][i[0]], w = b->wid[0][i[0]];
  for (int d = 1; d < b->ndims; d++) {
    r = r * b->W[d] + w * b->cum[d][i[d]];
    w *= b->wid[d][i[d]];
  }
  return (r * nc + (long long)c * w) * nc;
}

extern "C" int b2_pattern_export_device(b2_pattern* p, int64_t* rowptr_dev, int64_t* colidx_dev) {
  if (!p || !rowptr_dev || !colidx_dev) return B2_EINVAL;
  B2_CUDA(p->ctx, cudaSetDevice(p->ctx->device));
  if (pattern_general(p)) return launch_pattern_export_general(p->ctx, p->d_rowptr_b, p->d_colidx_b, p->nrows / pattern_ncomp(p), pattern_ncomp(p), (long long*)rowptr_dev, (long long*)colidx_dev);
  return launch_pattern_export(p->ctx, p->basis->view(), (long long*)rowptr_dev, (long long*)colidx_dev);
}

extern "C" int b2_pattern_export_host(b2_pattern* p, int64_t* rowptr_host, int64_t* colidx_host) {
  if (!p || !rowptr_host || !colidx_host) return B2_EINVAL;
  b2_ctx* ctx = p->ctx;
  B2_CUDA(ctx, cudaSetDevice(ctx->device));
  long long *d_rp = nullptr, *d_ci = nullptr;
  B2_CUDA(ctx, cudaMalloc((void**)&d_rp, sizeof(long long) * (p->nrows + 1)));
  cudaError_t e = cudaMalloc((void**)&d_ci, sizeof(long long) * std::max<int64_t>(p->nnz, 1));
  if (e != cudaSuccess) { cudaFree(d_rp); return b2_cuda_fail(ctx, e, "cudaMalloc(colidx)"); }
  int rc = pattern_general(p) ? launch_pattern_export_general(ctx, p->d_rowptr_b, p->d_colidx_b, p->nrows / pattern_ncomp(p), pattern_ncomp(p), d_rp, d_ci)
                      : launch_pattern_export(ctx, p->basis->view(), d_rp, d_ci);
  if (rc == B2_OK) {
    e = cudaMemcpyAsync(rowptr_host, d_rp, sizeof(long long) * (p->nrows + 1), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(colidx_host, d_ci, sizeof(long long) * p->nnz, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = b2_cuda_fail(ctx, e, "pattern download");
  }
  cudaFree(d_rp);
  cudaFree(d_ci);
  return rc;
}

// ---- assembly --------------------------------------------------------------------------------------

// 1-D tables of values and xi-derivatives of every coefficient set at the 1-D quadrature points,
// cached per (basis, quadrature) pair: [nsets][2][p+1][nq]
static int get_tabs(b2_ctx* ctx, const b2_basis* cb, const b2_quad* q, TabDev* out) {
  b2_basis* b = const_cast<b2_basis*>(cb);
  auto it = b->tabs.find(q->serial);
  if (it != b->tabs.end()) { *out = it->second; return B2_OK; }
  TabDev t;
  for (int d = 0; d < b->ndims; d++) {
    const int p = b->p[d], nq = q->nq[d];
    std::vector<double> h((size_t)b->nsets[d] * 2 * (p + 1) * nq);
    for (int s = 0; s < b->nsets[d]; s++)
      for (int a = 0; a <= p; a++) {
        const double* c = &b->coeffs[d][((size_t)s * (p + 1) + a) * (p + 1)];
        for (int k = 0; k < nq; k++) {
          const double x = q->pts[d][k];
          double v = c[0], g = 0.;
          for (int j = 1; j <= p; j++) { g = g * x + v; v = v * x + c[j]; }  // Horner with derivative
          h[(((size_t)s * 2 + 0) * (p + 1) + a) * nq + k] = v;
          h[(((size_t)s * 2 + 1) * (p + 1) + a) * nq + k] = g;
        }
      }
    int rc = upload(ctx, h, &t.tab[d]);
    if (rc != B2_OK) return rc;
  }
  b->tabs[q->serial] = t;
  *out = t;
  return B2_OK;
}

// Sparse term lists of the coefficient tensors D_m / C_v, uploaded into the context's form buffer (synchronous small copy).
int b2_upload_forms(b2_ctx* ctx, int nd, int nc, int nmat, const double* const* D_host, double* const* values_dev,
                    int nvec, const double* const* C_host, double* const* rhs_dev, FormView* out) {
  const int na = nd + 1;
  // sparse term lists of the coefficient tensors
  std::vector<int> termptr(1, 0), termxy;
  std::vector<double> termval, vcoef;
  for (int m = 0; m < nmat; m++) {
    for (int c = 0; c < nc; c++)
      for (int e = 0; e < nc; e++) {
        for (int x = 0; x < na; x++)
          for (int y = 0; y < na; y++) {
            const double v = D_host[m][((c * na + x) * nc + e) * na + y];
            if (v != 0.) { termxy.push_back(x | y << 8); termval.push_back(v); }
          }
        termptr.push_back((int)termxy.size());
      }
  }
  for (int v = 0; v < nvec; v++) {
    vcoef.insert(vcoef.end(), C_host[v], C_host[v] + nc * na);
  }
  // dense 4x4 blocks per (m, c, e) for the tensor-core path
  std::vector<double> jobD;
  std::vector<int> jobinfo;
  for (int m = 0; m < nmat; m++)
    for (int c = 0; c < nc; c++)
      for (int e = 0; e < nc; e++) {
        double blk[16] = {0};
        bool any = false, only00 = true;
        for (int x = 0; x < na; x++)
          for (int y = 0; y < na; y++) {
            const double v = D_host[m][((c * na + x) * nc + e) * na + y];
            blk[x * 4 + y] = v;
            if (v != 0.) { any = true; if (x || y) only00 = false; }
          }
        if (!any) continue;
        jobD.insert(jobD.end(), blk, blk + 16);
        jobinfo.push_back(m | c << 8 | e << 16 | (only00 ? 1 : 0) << 24);
      }
  const size_t b_ptr = termptr.size() * sizeof(int), b_xy = std::max<size_t>(termxy.size(), 1) * sizeof(int);
  const size_t b_val = std::max<size_t>(termval.size(), 1) * sizeof(double), b_vc = std::max<size_t>(vcoef.size(), 1) * sizeof(double);
  const size_t b_jd = std::max<size_t>(jobD.size(), 1) * sizeof(double), b_ji = std::max<size_t>(jobinfo.size(), 1) * sizeof(int);
  const size_t off_val = 0, off_vc = off_val + b_val, off_jd = off_vc + b_vc, off_ptr = off_jd + b_jd, off_xy = off_ptr + ((b_ptr + 7) & ~size_t(7));
  const size_t off_ji = off_xy + ((b_xy + 7) & ~size_t(7));
  const size_t total = off_ji + b_ji;
  if (ctx->formbuf_bytes < total) {
    if (ctx->formbuf) cudaFree(ctx->formbuf);
    ctx->formbuf = nullptr;
    ctx->formbuf_bytes = 0;
    B2_CUDA(ctx, cudaMalloc(&ctx->formbuf, std::max<size_t>(total, 4096)));
    ctx->formbuf_bytes = std::max<size_t>(total, 4096);
  }
  std::vector<unsigned char> stage(total, 0);
  if (!termval.empty()) memcpy(&stage[off_val], termval.data(), termval.size() * sizeof(double));
  if (!vcoef.empty()) memcpy(&stage[off_vc], vcoef.data(), vcoef.size() * sizeof(double));
  memcpy(&stage[off_ptr], termptr.data(), b_ptr);
  if (!termxy.empty()) memcpy(&stage[off_xy], termxy.data(), termxy.size() * sizeof(int));
  if (!jobD.empty()) memcpy(&stage[off_jd], jobD.data(), jobD.size() * sizeof(double));
  if (!jobinfo.empty()) memcpy(&stage[off_ji], jobinfo.data(), jobinfo.size() * sizeof(int));
  // synchronous small copy: the staging vector dies at return
  B2_CUDA(ctx, cudaMemcpyAsync(ctx->formbuf, stage.data(), total, cudaMemcpyHostToDevice, ctx->stream));
  B2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));

  FormView F;
  memset(&F, 0, sizeof(F));
  F.nmat = nmat;
  F.nvec = nvec;
  unsigned char* fb = (unsigned char*)ctx->formbuf;
  F.termval = (const double*)(fb + off_val);
  F.vcoef = (const double*)(fb + off_vc);
  F.termptr = (const int*)(fb + off_ptr);
  F.termxy = (const int*)(fb + off_xy);
  F.jobD = (const double*)(fb + off_jd);
  F.jobinfo = (const int*)(fb + off_ji);
  F.njobs = (int)jobinfo.size();
  for (int m = 0; m < nmat; m++) F.values[m] = values_dev[m];
  for (int v = 0; v < nvec; v++) F.rhs[v] = rhs_dev[v];

  *out = F;
  return B2_OK;
}

// Coverage configurations in 2-D and 3-D (C0 bases of degree >= 3, over-integration, mixed value/gradient forms, two
// components, ...) run the element-set kernel with its tensor-core block product on the element set "everything": its
// materialised pattern has exactly the slots of the analytic one (tests: test_full_selection_equals_structured).  Built
// once per pattern; if it cannot be built (memory) the scalar generic kernel takes over.
static int assemble_via_elemset(b2_ctx* ctx, const b2_pattern* cpattern, const b2_basis* basis, const b2_quad* quad, const b2_geom* geom,
                                int64_t elem_begin, int64_t elem_end, int nmat, const double* const* D_host, double* const* values_dev,
                                int nvec, const double* const* C_host, double* const* rhs_dev) {
  b2_pattern* pattern = const_cast<b2_pattern*>(cpattern);
  if (pattern->aux_failed) return B2_EUNSUPPORTED;
  if (!pattern->aux_pattern) {
    int64_t ntot = 1;
    for (int d = 0; d < basis->ndims; d++) ntot *= basis->nel[d];
    int rc = b2_elemset_create(ctx, basis, ntot, nullptr, nullptr, nullptr, nullptr, nullptr, 0, nullptr, 0, &pattern->aux_elemset);
    if (rc == B2_OK) rc = b2_pattern_create_elemset(ctx, pattern->aux_elemset, &pattern->aux_pattern);
    if (rc != B2_OK || pattern->aux_pattern->nnz != pattern->nnz) {
      if (pattern->aux_pattern) b2_pattern_destroy(pattern->aux_pattern);
      if (pattern->aux_elemset) b2_elemset_destroy(pattern->aux_elemset);
      pattern->aux_pattern = nullptr;
      pattern->aux_elemset = nullptr;
      pattern->aux_failed = true;
      return B2_EUNSUPPORTED;
    }
  }
  return b2_assemble_elemset_device(ctx, pattern->aux_pattern, pattern->aux_elemset, quad, geom, elem_begin, elem_end, nmat, D_host, values_dev, nvec, C_host, rhs_dev);
}

// rows == false: integrate elements [elem_begin, elem_end) and ACCUMULATE into the outputs.
// rows == true:  elem_begin/elem_end are dof planes of dimension 0; every stored value of those planes is WRITTEN.
static int assemble_impl(b2_ctx* ctx, const b2_pattern* pattern, const b2_basis* basis, const b2_quad* quad, const b2_geom* geom,
                         int64_t elem_begin, int64_t elem_end, int nmat, const double* const* D_host, double* const* values_dev,
                         int nvec, const double* const* C_host, double* const* rhs_dev, bool rows = false) {
  if (!ctx || !pattern || !basis || !quad || !geom) return b2_fail(ctx, B2_EINVAL, "null argument");
  if (nmat < 0 || nmat > B2_MAX_FORMS || nvec < 0 || nvec > B2_MAX_FORMS) return b2_fail(ctx, B2_EINVAL, "0..4 matrix and vector forms per call");
  if ((nmat && (!D_host || !values_dev)) || (nvec && (!C_host || !rhs_dev))) return b2_fail(ctx, B2_EINVAL, "null form argument");
  if (pattern->basis != basis) return b2_fail(ctx, B2_EINVAL, "pattern was built for a different basis");
  if (pattern->elemset) return b2_fail(ctx, B2_EINVAL, "pattern of an element set: use b2_assemble_elemset_*");
  if (geom->gbasis) return b2_fail(ctx, B2_EUNSUPPORTED, "spline geometries are integrated through b2_assemble_elemset_*");
  if (quad->ndims != basis->ndims || geom->ndims != basis->ndims) return b2_fail(ctx, B2_EINVAL, "dimension mismatch");
  int64_t ntot = 1;
  for (int d = 0; d < basis->ndims; d++) {
    if (geom->nel[d] != basis->nel[d]) return b2_fail(ctx, B2_EINVAL, "geometry and basis live on different topologies");
    ntot *= basis->nel[d];
  }
  int64_t plane_begin = 0, plane_end = basis->ndofs_d[0];
  if (rows) {
    plane_begin = elem_begin;
    plane_end = elem_end < 0 ? basis->ndofs_d[0] : elem_end;
    if (plane_begin < 0 || plane_begin > plane_end || plane_end > basis->ndofs_d[0]) return b2_fail(ctx, B2_EINVAL, "invalid dof-plane range");
    if (plane_begin == plane_end || (nmat == 0 && nvec == 0)) return B2_OK;
    // elements of dimension 0 whose dofs touch the planes (start is non-decreasing)
    int64_t e_lo = basis->nel[0], e_hi = -1;
    for (int64_t e = 0; e < basis->nel[0]; e++)
      if (basis->start[0][e] < plane_end && basis->start[0][e] + basis->p[0] >= plane_begin) { e_lo = std::min(e_lo, e); e_hi = std::max(e_hi, e); }
    const int64_t per_layer = ntot / basis->nel[0];
    elem_begin = e_lo * per_layer;
    elem_end = (e_hi + 1) * per_layer;
    if (e_hi < e_lo) elem_begin = elem_end = 0;
  } else {
    if (elem_end < 0) elem_end = ntot;
    if (elem_begin < 0 || elem_begin > elem_end || elem_end > ntot) return b2_fail(ctx, B2_EINVAL, "invalid element range");
    if (elem_begin == elem_end || (nmat == 0 && nvec == 0)) return B2_OK;
  }
  B2_CUDA(ctx, cudaSetDevice(ctx->device));

  const int nd = basis->ndims, nc = basis->ncomp;
  TabDev tabs;
  int rc = get_tabs(ctx, basis, quad, &tabs);
  if (rc != B2_OK) return rc;

  BasisView B = basis->view();
  QuadView Q;
  memset(&Q, 0, sizeof(Q));
  Q.nqt = 1;
  for (int d = 0; d < nd; d++) {
    Q.nq[d] = quad->nq[d];
    Q.nqt *= quad->nq[d];
    Q.x[d] = quad->d_x[d];
    Q.w[d] = quad->d_w[d];
    Q.tab[d] = tabs.tab[d];
  }
  GeomView G;
  memset(&G, 0, sizeof(G));
  G.nodes = geom->d_nodes;
  G.nnodes = geom->nnodes;
  {
    long long s = 1;
    for (int d = nd - 1; d >= 0; d--) { G.stride[d] = s; s *= geom->nel[d] + 1; }
  }

  const int64_t kernel_opt = ctx->opts.count("kernel") ? ctx->opts["kernel"] : 0;
  for (int m = 0; m < nmat; m++)
    if (!D_host[m] || !values_dev[m]) return b2_fail(ctx, B2_EINVAL, "null matrix form");
  for (int v = 0; v < nvec; v++)
    if (!C_host[v] || !rhs_dev[v]) return b2_fail(ctx, B2_EINVAL, "null vector form");
  if (rows && kernel_opt != 1) {
    // the specialised owner-computes kernel takes its coefficients as kernel parameters: no upload, no synchronisation
    FormView F0;
    memset(&F0, 0, sizeof(F0));
    F0.nmat = nmat;
    F0.nvec = nvec;
    for (int m = 0; m < nmat; m++) F0.values[m] = values_dev[m];
    for (int v = 0; v < nvec; v++) F0.rhs[v] = rhs_dev[v];
    rc = launch_assemble_rows(ctx, basis, quad, B, Q, G, F0, D_host, C_host, plane_begin, plane_end);
    if (rc != B2_EUNSUPPORTED) return rc;
    if (kernel_opt >= 2) return b2_fail(ctx, B2_EUNSUPPORTED, "requested specialised kernel does not cover this configuration");
  }

  FormView F;
  rc = b2_upload_forms(ctx, nd, nc, nmat, D_host, values_dev, nvec, C_host, rhs_dev, &F);
  if (rc != B2_OK) return rc;

  if (rows) {
    // coverage path: zero the planes' slots, then scatter only the rows inside the planes
    const int64_t nb_plane = basis->nbasis / basis->ndofs_d[0] * nc;
    const int64_t s0 = b2_pattern_row_offset(pattern, plane_begin * nb_plane), s1 = b2_pattern_row_offset(pattern, plane_end * nb_plane);
    for (int m = 0; m < nmat; m++) B2_CUDA(ctx, cudaMemsetAsync(values_dev[m] + s0, 0, sizeof(double) * (size_t)(s1 - s0), ctx->stream));
    for (int v = 0; v < nvec; v++) B2_CUDA(ctx, cudaMemsetAsync(rhs_dev[v] + plane_begin * nb_plane, 0, sizeof(double) * (size_t)((plane_end - plane_begin) * nb_plane), ctx->stream));
    if (elem_begin == elem_end) return B2_OK;
    if (plane_begin == 0 && plane_end == basis->ndofs_d[0] && kernel_opt != 1) {
      // all planes (no row filter needed): the specialised element-scatter kernel where it applies -- e.g. C0 ('std')
      // bases of degree 1 and 2, which the owner-computes kernel (maximal smoothness) does not cover
      rc = launch_assemble_fast(ctx, B, Q, G, F, D_host, C_host, elem_begin, elem_end);
      if (rc != B2_EUNSUPPORTED) return rc;
      if (kernel_opt == 0 && nd >= 2 && nmat > 0 && B.nb <= 128) {
        rc = assemble_via_elemset(ctx, pattern, basis, quad, geom, elem_begin, elem_end, nmat, D_host, values_dev, nvec, C_host, rhs_dev);
        if (rc != B2_EUNSUPPORTED) return rc;
      }
    }
    return launch_assemble_generic(ctx, B, Q, G, F, elem_begin, elem_end, (int)plane_begin, (int)plane_end);
  }
  if (kernel_opt != 1) {
    rc = launch_assemble_fast(ctx, B, Q, G, F, D_host, C_host, elem_begin, elem_end);
    if (rc != B2_EUNSUPPORTED) return rc;
    if (kernel_opt >= 2) return b2_fail(ctx, B2_EUNSUPPORTED, "requested specialised kernel does not cover this configuration");
    if (kernel_opt == 0 && nd >= 2 && nmat > 0 && B.nb <= 128) {
      rc = assemble_via_elemset(ctx, pattern, basis, quad, geom, elem_begin, elem_end, nmat, D_host, values_dev, nvec, C_host, rhs_dev);
      if (rc != B2_EUNSUPPORTED) return rc;
    }
  }
  return launch_assemble_generic(ctx, B, Q, G, F, elem_begin, elem_end, 0, 0x7fffffff);
}

extern "C" int b2_assemble_device(b2_ctx* ctx, const b2_pattern* pattern, const b2_basis* basis, const b2_quad* quad, const b2_geom* geom,
                                  int64_t elem_begin, int64_t elem_end, int nmat, const double* const* D_host, double* const* values_dev,
                                  int nvec, const double* const* C_host, double* const* rhs_dev) {
  return assemble_impl(ctx, pattern, basis, quad, geom, elem_begin, elem_end, nmat, D_host, values_dev, nvec, C_host, rhs_dev);
}

extern "C" int b2_assemble_rows_device(b2_ctx* ctx, const b2_pattern* pattern, const b2_basis* basis, const b2_quad* quad, const b2_geom* geom,
                                       int64_t plane_begin, int64_t plane_end, int nmat, const double* const* D_host, double* const* values_dev,
                                       int nvec, const double* const* C_host, double* const* rhs_dev) {
  return assemble_impl(ctx, pattern, basis, quad, geom, plane_begin, plane_end, nmat, D_host, values_dev, nvec, C_host, rhs_dev, true);
}

extern "C" int b2_assemble_host(b2_ctx* ctx, const b2_pattern* pattern, const b2_basis* basis, const b2_quad* quad, const b2_geom* geom,
                                int64_t elem_begin, int64_t elem_end, int nmat, const double* const* D_host, double* const* values_host,
                                int nvec, const double* const* C_host, double* const* rhs_host) {
  if (!ctx || !pattern) return B2_EINVAL;
  if (nmat < 0 || nmat > B2_MAX_FORMS || nvec < 0 || nvec > B2_MAX_FORMS) return b2_fail(ctx, B2_EINVAL, "0..4 matrix and vector forms per call");
  if ((nmat && !values_host) || (nvec && !rhs_host)) return b2_fail(ctx, B2_EINVAL, "null output argument");
  B2_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t bm = sizeof(double) * (size_t)pattern->nnz, bv = sizeof(double) * (size_t)pattern->nrows;
  const size_t need = std::max<size_t>(bm * nmat + bv * nvec, 8);
  if (ctx->scratch_bytes < need) {
    if (ctx->scratch) cudaFree(ctx->scratch);
    ctx->scratch = nullptr;
    ctx->scratch_bytes = 0;
    B2_CUDA(ctx, cudaMalloc(&ctx->scratch, need));
    ctx->scratch_bytes = need;
  }
  int64_t ntot = 1;
  if (basis) for (int d = 0; d < basis->ndims; d++) ntot *= basis->nel[d];
  // option "path" = 1 forces the element-scatter (accumulate) kernels also for the whole topology
  const bool whole = basis && elem_begin == 0 && (elem_end < 0 || elem_end == ntot) && !(ctx->opts.count("path") && ctx->opts["path"] == 1);
  if (!whole) B2_CUDA(ctx, cudaMemsetAsync(ctx->scratch, 0, need, ctx->stream));
  double* vals[B2_MAX_FORMS];
  double* rhs[B2_MAX_FORMS];
  unsigned char* base = (unsigned char*)ctx->scratch;
  for (int m = 0; m < nmat; m++) vals[m] = (double*)(base + bm * m);
  for (int v = 0; v < nvec; v++) rhs[v] = (double*)(base + bm * nmat + bv * v);
  for (int m = 0; m < nmat; m++)
    if (!values_host[m]) return b2_fail(ctx, B2_EINVAL, "null output buffer");
  for (int v = 0; v < nvec; v++)
    if (!rhs_host[v]) return b2_fail(ctx, B2_EINVAL, "null output buffer");
  // The whole topology: owner-computes rows (every value written once, no zero fill).  Large results are produced in
  // chunks of dof planes so that the device-to-host copy of one chunk (copy stream) overlaps the integration of the next.
  if (whole && need >= ((size_t)64 << 20) && basis->ndofs_d[0] >= 16) {
    if (!ctx->copy_stream) {
      B2_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
      B2_CUDA(ctx, cudaEventCreateWithFlags(&ctx->copy_event, cudaEventDisableTiming));
    }
    const int64_t np = basis->ndofs_d[0], per_plane = pattern->nrows / np;
    const int nchunk = (int)std::min<int64_t>(8, np / 8);
    for (int k = 0; k < nchunk; k++) {
      const int64_t p0 = np * k / nchunk, p1 = np * (k + 1) / nchunk;
      int rc = assemble_impl(ctx, pattern, basis, quad, geom, p0, p1, nmat, D_host, vals, nvec, C_host, rhs, true);
      if (rc != B2_OK) { cudaStreamSynchronize(ctx->copy_stream); return rc; }
      B2_CUDA(ctx, cudaEventRecord(ctx->copy_event, ctx->stream));
      B2_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_event, 0));
      const int64_t s0 = b2_pattern_row_offset(pattern, p0 * per_plane), s1 = b2_pattern_row_offset(pattern, p1 * per_plane);
      for (int m = 0; m < nmat; m++)
        B2_CUDA(ctx, cudaMemcpyAsync(values_host[m] + s0, vals[m] + s0, sizeof(double) * (size_t)(s1 - s0), cudaMemcpyDeviceToHost, ctx->copy_stream));
      for (int v = 0; v < nvec; v++)
        B2_CUDA(ctx, cudaMemcpyAsync(rhs_host[v] + p0 * per_plane, rhs[v] + p0 * per_plane, sizeof(double) * (size_t)((p1 - p0) * per_plane), cudaMemcpyDeviceToHost, ctx->copy_stream));
    }
    B2_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    B2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B2_OK;
  }
  // otherwise one pass: rows for the whole topology, accumulate for a slab of elements
  int rc = whole ? assemble_impl(ctx, pattern, basis, quad, geom, 0, -1, nmat, D_host, vals, nvec, C_host, rhs, true)
                 : assemble_impl(ctx, pattern, basis, quad, geom, elem_begin, elem_end, nmat, D_host, vals, nvec, C_host, rhs);
  if (rc != B2_OK) return rc;
  for (int m = 0; m < nmat; m++) {
    B2_CUDA(ctx, cudaMemcpyAsync(values_host[m], vals[m], bm, cudaMemcpyDeviceToHost, ctx->stream));
  }
  for (int v = 0; v < nvec; v++) {
    B2_CUDA(ctx, cudaMemcpyAsync(rhs_host[v], rhs[v], bv, cudaMemcpyDeviceToHost, ctx->stream));
  }
  B2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return B2_OK;
}
