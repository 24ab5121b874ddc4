// Host side of the element-set entry points of include/b200fem.h: element subsets with ragged quadrature, pruned
// numberings and rational functions (b2_elemset_*), spline geometries (b2_geom_create_spline), the general CSR
// pattern (b2_pattern_create_elemset) and the element-scatter assembly over them (b2_assemble_elemset_*).

#include <algorithm>
#include <cstring>
#include <new>

#include "common.cuh"

int b2_upload_forms(b2_ctx* ctx, int nd, int nc, int nmat, const double* const* D_host, double* const* values_dev,
                    int nvec, const double* const* C_host, double* const* rhs_dev, FormView* out);

namespace {

template <class T>
int upload_n(b2_ctx* ctx, const T* h, size_t n, T** d) {
  *d = nullptr;
  B2_CUDA(ctx, cudaMalloc((void**)d, std::max<size_t>(n, 1) * sizeof(T)));
  if (n) B2_CUDA(ctx, cudaMemcpy(*d, h, n * sizeof(T), cudaMemcpyHostToDevice));
  return B2_OK;
}

int64_t total_elems(const b2_basis* b) {
  int64_t n = 1;
  for (int d = 0; d < b->ndims; d++) n *= b->nel[d];
  return n;
}

}  // namespace

extern "C" int b2_elemset_create(b2_ctx* ctx, const b2_basis* basis, int64_t nsel, const int64_t* elem_ids, const int64_t* qoff, const double* qcoords,
                                 const double* qweights, const int64_t* renumber, int64_t nbasis_new, const double* scale, int rational, b2_elemset** out) {
  if (!ctx || !basis || !out) return b2_fail(ctx, B2_EINVAL, "null argument");
  *out = nullptr;
  const int64_t ntot = total_elems(basis);
  if (nsel < 0 || nsel > ntot || (!elem_ids && nsel != ntot)) return b2_fail(ctx, B2_EINVAL, "invalid number of selected elements");
  if (qoff && (!qcoords || !qweights)) return b2_fail(ctx, B2_EINVAL, "qoff without qcoords/qweights");
  if (rational < 0 || rational > 2 || (rational == 1 && !scale)) return b2_fail(ctx, B2_EINVAL, "invalid rational mode");
  if (!renumber) nbasis_new = basis->nbasis;
  if (nbasis_new < 0 || nbasis_new > basis->nbasis) return b2_fail(ctx, B2_EINVAL, "invalid nbasis_new");
  B2_CUDA(ctx, cudaSetDevice(ctx->device));
  b2_elemset* es = new (std::nothrow) b2_elemset();
  if (!es) return B2_ENOMEM;
  es->ctx = ctx;
  es->basis = basis;
  es->nsel = nsel;
  es->nbasis_new = nbasis_new;
  es->rational = rational;
  es->npoints = 0;
  es->max_nq = 0;
  int rc = B2_OK;
  if (elem_ids) {
    for (int64_t k = 0; k < nsel; k++)
      if (elem_ids[k] < 0 || elem_ids[k] >= ntot || (k && elem_ids[k] <= elem_ids[k - 1])) { delete es; return b2_fail(ctx, B2_EINVAL, "elem_ids must be strictly increasing element indices"); }
    es->elem_ids.assign(elem_ids, elem_ids + nsel);
    std::vector<unsigned char> mask((size_t)ntot, 0);
    for (int64_t k = 0; k < nsel; k++) mask[(size_t)elem_ids[k]] = 1;
    rc = upload_n(ctx, (const long long*)elem_ids, (size_t)nsel, &es->d_elem_ids);
    if (rc == B2_OK) rc = upload_n(ctx, mask.data(), mask.size(), &es->d_selmask);
  }
  if (rc == B2_OK && qoff) {
    if (qoff[0] != 0) rc = b2_fail(ctx, B2_EINVAL, "qoff[0] must be 0");
    for (int64_t k = 0; k < nsel && rc == B2_OK; k++) {
      const int64_t n = qoff[k + 1] - qoff[k];
      if (n < 0 || n > (1 << 24)) rc = b2_fail(ctx, B2_EINVAL, "qoff must be non-decreasing");
      es->max_nq = std::max<int>(es->max_nq, (int)n);
    }
    if (rc == B2_OK) {
      es->npoints = qoff[nsel];
      rc = upload_n(ctx, (const long long*)qoff, (size_t)nsel + 1, &es->d_qoff);
      if (rc == B2_OK) rc = upload_n(ctx, qcoords, (size_t)es->npoints * basis->ndims, &es->d_qcoords);
      if (rc == B2_OK) rc = upload_n(ctx, qweights, (size_t)es->npoints, &es->d_qweights);
      if (rc == B2_OK) {
        // longest-processing-time-first: the element queue of the kernel starts with the cells that carry the most points
        std::vector<long long> order((size_t)nsel);
        for (int64_t k = 0; k < nsel; k++) order[(size_t)k] = k;
        std::stable_sort(order.begin(), order.end(), [&](long long a, long long b) { return qoff[a + 1] - qoff[a] > qoff[b + 1] - qoff[b]; });
        rc = upload_n(ctx, order.data(), order.size(), &es->d_order);
      }
    }
  }
  if (rc == B2_OK && renumber) {
    es->renumber.resize((size_t)basis->nbasis);
    std::vector<long long> dofmap((size_t)nbasis_new, -1);
    int64_t prev = -1;
    for (int64_t I = 0; I < basis->nbasis && rc == B2_OK; I++) {
      int64_t r = renumber[I];
      if (r < 0 || r >= nbasis_new) r = -1;
      es->renumber[(size_t)I] = (int)r;
      if (r >= 0) {
        if (r <= prev) rc = b2_fail(ctx, B2_EUNSUPPORTED, "renumber must be increasing over the kept basis functions");
        prev = r;
        dofmap[(size_t)r] = I;
      }
    }
    for (int64_t r = 0; r < nbasis_new && rc == B2_OK; r++)
      if (dofmap[(size_t)r] < 0) rc = b2_fail(ctx, B2_EINVAL, "renumber does not cover 0..nbasis_new-1");
    if (rc == B2_OK) rc = upload_n(ctx, es->renumber.data(), es->renumber.size(), &es->d_renumber);
    if (rc == B2_OK) rc = upload_n(ctx, dofmap.data(), dofmap.size(), &es->d_dofmap);
  }
  if (rc == B2_OK && scale) rc = upload_n(ctx, scale, (size_t)basis->nbasis, &es->d_scale);
  for (int d = 0; d < basis->ndims && rc == B2_OK; d++) {
    rc = upload_n(ctx, basis->coeffs[d].data(), basis->coeffs[d].size(), &es->d_coeffs[d]);
    // first / last element in the support of every 1-D function (start is non-decreasing)
    const int nd = (int)basis->ndofs_d[d], p = basis->p[d];
    std::vector<int> ef(nd, 0), el(nd, -1);
    for (int i = 0; i < nd; i++) ef[i] = (int)basis->nel[d];
    for (int e = 0; e < (int)basis->nel[d]; e++)
      for (int a = 0; a <= p; a++) {
        const int i = basis->start[d][e] + a;
        ef[i] = std::min(ef[i], e);
        el[i] = std::max(el[i], e);
      }
    if (rc == B2_OK) rc = upload_n(ctx, ef.data(), ef.size(), &es->d_efirst[d]);
    if (rc == B2_OK) rc = upload_n(ctx, el.data(), el.size(), &es->d_elast[d]);
  }
  if (rc != B2_OK) { b2_elemset_destroy(es); return rc; }
  *out = es;
  return B2_OK;
}

extern "C" int b2_elemset_set_faces(b2_elemset* es, const int8_t* face_dim) {
  if (!es) return B2_EINVAL;
  b2_ctx* ctx = es->ctx;
  B2_CUDA(ctx, cudaSetDevice(ctx->device));
  if (es->d_face_dim) { cudaFree(es->d_face_dim); es->d_face_dim = nullptr; }
  if (!face_dim) return B2_OK;
  for (int64_t k = 0; k < es->nsel; k++)
    if (face_dim[k] < -1 || face_dim[k] >= es->basis->ndims) return b2_fail(ctx, B2_EINVAL, "face_dim must be -1 or a reference direction");
  return upload_n(ctx, (const signed char*)face_dim, (size_t)es->nsel, &es->d_face_dim);
}

extern "C" int b2_elemset_set_normals(b2_elemset* es, const double* nref, int64_t npoints) {
  if (!es) return B2_EINVAL;
  b2_ctx* ctx = es->ctx;
  B2_CUDA(ctx, cudaSetDevice(ctx->device));
  if (es->d_normals) { cudaFree(es->d_normals); es->d_normals = nullptr; }
  if (!nref) return B2_OK;
  if (!es->d_qoff || npoints != es->npoints) return b2_fail(ctx, B2_EINVAL, "facet normals need an element set with its own points, one normal per point");
  return upload_n(ctx, nref, (size_t)npoints * es->basis->ndims, &es->d_normals);
}

extern "C" int b2_elemset_set_coefficient(b2_elemset* es, int which, const double* coef, int64_t npoints) {
  if (!es || which < 0 || which >= 2 * B2_MAX_FORMS) return B2_EINVAL;
  b2_ctx* ctx = es->ctx;
  B2_CUDA(ctx, cudaSetDevice(ctx->device));
  if (es->d_coef[which]) { cudaFree(es->d_coef[which]); es->d_coef[which] = nullptr; es->coef_len[which] = 0; }
  if (!coef) return B2_OK;
  if (npoints < 0 || (es->d_qoff && npoints != es->npoints)) return b2_fail(ctx, B2_EINVAL, "one coefficient per quadrature point is needed");
  es->coef_len[which] = npoints;
  return upload_n(ctx, coef, (size_t)npoints, &es->d_coef[which]);
}

extern "C" int b2_elemset_set_coefficient_field(b2_elemset* es, int which, const double* field_dev, int power, double scale) {
  if (!es || which < 0 || which >= 2 * B2_MAX_FORMS) return B2_EINVAL;
  if (field_dev && (power < 0 || power > 8)) return b2_fail(es->ctx, B2_EINVAL, "power of the field coefficient must be 0..8");
  if (field_dev && (es->basis->ncomp != 1 || es->rational)) return b2_fail(es->ctx, B2_EUNSUPPORTED, "field coefficients: scalar, non-rational spaces");
  es->d_field[which] = field_dev;
  es->field_power[which] = power;
  es->field_scale[which] = scale;
  return B2_OK;
}

extern "C" int b2_elemset_destroy(b2_elemset* es) {
  if (!es) return B2_OK;
  cudaSetDevice(es->ctx->device);
  if (es->d_face_dim) cudaFree(es->d_face_dim);
  if (es->d_normals) cudaFree(es->d_normals);
  for (double* p : es->d_coef)
    if (p) cudaFree(p);
  void* ptrs[] = {es->d_order, es->d_elem_ids, es->d_qoff, es->d_qcoords, es->d_qweights, es->d_renumber, es->d_scale, es->d_selmask, es->d_dofmap,
                  es->d_coeffs[0], es->d_coeffs[1], es->d_coeffs[2], es->d_efirst[0], es->d_efirst[1], es->d_efirst[2], es->d_elast[0], es->d_elast[1], es->d_elast[2]};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  delete es;
  return B2_OK;
}

extern "C" int64_t b2_elemset_ndofs(const b2_elemset* es) { return es ? es->nbasis_new * es->basis->ncomp : 0; }
extern "C" int64_t b2_elemset_npoints(const b2_elemset* es) { return es ? es->npoints : 0; }

extern "C" int b2_geom_create_spline(b2_ctx* ctx, const b2_basis* gbasis, const double* ctrl_host, const double* weights_host, b2_geom** out) {
  if (!ctx || !gbasis || !ctrl_host || !out) return b2_fail(ctx, B2_EINVAL, "null argument");
  *out = nullptr;
  if (gbasis->ncomp != 1) return b2_fail(ctx, B2_EINVAL, "the geometry basis must be scalar");
  B2_CUDA(ctx, cudaSetDevice(ctx->device));
  b2_geom* g = new (std::nothrow) b2_geom();
  if (!g) return B2_ENOMEM;
  g->ctx = ctx;
  g->ndims = gbasis->ndims;
  g->nnodes = 0;
  g->d_nodes = nullptr;
  for (int d = 0; d < g->ndims; d++) g->nel[d] = gbasis->nel[d];
  g->gbasis = gbasis;
  int rc = upload_n(ctx, ctrl_host, (size_t)gbasis->nbasis * g->ndims, &g->d_ctrl);
  if (rc == B2_OK && weights_host) rc = upload_n(ctx, weights_host, (size_t)gbasis->nbasis, &g->d_wts);
  for (int d = 0; d < g->ndims && rc == B2_OK; d++) rc = upload_n(ctx, gbasis->coeffs[d].data(), gbasis->coeffs[d].size(), &g->d_gcoeffs[d]);
  if (rc != B2_OK) { b2_geom_destroy(g); return rc; }
  *out = g;
  return B2_OK;
}

extern "C" int b2_pattern_create_elemset(b2_ctx* ctx, const b2_elemset* es, b2_pattern** out) {
  if (!ctx || !es || !out) return b2_fail(ctx, B2_EINVAL, "null argument");
  *out = nullptr;
  const b2_basis* basis = es->basis;
  B2_CUDA(ctx, cudaSetDevice(ctx->device));
  b2_pattern* p = new (std::nothrow) b2_pattern();
  if (!p) return B2_ENOMEM;
  p->ctx = ctx;
  p->basis = basis;
  p->elemset = es;
  const long long nb = es->nbasis_new;
  const int nc = basis->ncomp;
  p->nrows = nb * nc;
  cudaError_t e = cudaMalloc((void**)&p->d_rowptr_b, sizeof(long long) * (size_t)(nb + 1));
  if (e != cudaSuccess) { delete p; return b2_cuda_fail(ctx, e, "cudaMalloc(rowptr)"); }
  const BasisView B = basis->view();
  int rc = launch_pattern_elemset_impl(ctx, B, es->d_efirst, es->d_elast, es->d_dofmap, es->d_renumber, es->d_selmask, nb, 0, p->d_rowptr_b, nullptr);
  if (rc == B2_OK) rc = launch_exclusive_scan(ctx, p->d_rowptr_b, nb);
  if (rc == B2_OK) {
    p->rowptr_b.resize((size_t)nb + 1);
    e = cudaMemcpy(p->rowptr_b.data(), p->d_rowptr_b, sizeof(long long) * (size_t)(nb + 1), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = b2_cuda_fail(ctx, e, "rowptr download");
  }
  if (rc == B2_OK) {
    const long long nnz_b = p->rowptr_b[(size_t)nb];
    p->nnz = nnz_b * nc * nc;
    e = cudaMalloc((void**)&p->d_colidx_b, sizeof(int) * (size_t)std::max<long long>(nnz_b, 1));
    if (e != cudaSuccess) rc = b2_cuda_fail(ctx, e, "cudaMalloc(colidx)");
  }
  if (rc == B2_OK) rc = launch_pattern_elemset_impl(ctx, B, es->d_efirst, es->d_elast, es->d_dofmap, es->d_renumber, es->d_selmask, nb, 1, p->d_rowptr_b, p->d_colidx_b);
  if (rc == B2_OK) {
    e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = b2_cuda_fail(ctx, e, "pattern construction");
  }
  if (rc != B2_OK) { b2_pattern_destroy(p); return rc; }
  *out = p;
  return B2_OK;
}

// device views of an element set with its rule and geometry (shared by the assembly and the evaluation entry points)
struct ESViews {
  BasisView B;
  QuadView Q;
  GeomView G;
  SplineGeomView SG;
  ElemSetView E;
};

static int build_views(b2_ctx* ctx, const b2_pattern* pattern, const b2_elemset* es, const b2_quad* quad, const b2_geom* geom, ESViews* out) {
  const b2_basis* basis = es->basis;
  const int nd = basis->ndims;
  if (!es->d_qoff && !quad) return b2_fail(ctx, B2_EINVAL, "the element set carries no points and no quadrature rule was given");
  if ((quad && quad->ndims != nd) || geom->ndims != nd) return b2_fail(ctx, B2_EINVAL, "dimension mismatch");
  for (int d = 0; d < nd; d++)
    if (geom->nel[d] != basis->nel[d]) return b2_fail(ctx, B2_EINVAL, "geometry and basis live on different topologies");
  if (es->rational == 2 && !(geom->gbasis && geom->d_wts)) return b2_fail(ctx, B2_EINVAL, "rational mode 2 needs a rational spline geometry");
  BasisView B = basis->view();
  QuadView Q;
  memset(&Q, 0, sizeof(Q));
  Q.nqt = 0;
  if (quad && !es->d_qoff) {
    Q.nqt = 1;
    for (int d = 0; d < nd; d++) {
      Q.nq[d] = quad->nq[d];
      Q.nqt *= quad->nq[d];
      Q.x[d] = quad->d_x[d];
      Q.w[d] = quad->d_w[d];
    }
  }
  GeomView G;
  memset(&G, 0, sizeof(G));
  SplineGeomView SG;
  memset(&SG, 0, sizeof(SG));
  if (geom->gbasis) {
    SG.enabled = 1;
    SG.rational = geom->d_wts != nullptr;
    SG.GB = geom->gbasis->view();
    for (int d = 0; d < nd; d++) SG.coeffs[d] = geom->d_gcoeffs[d];
    SG.ctrl = geom->d_ctrl;
    SG.wts = geom->d_wts;
    SG.nbasis = geom->gbasis->nbasis;
  } else {
    G.nodes = geom->d_nodes;
    G.nnodes = geom->nnodes;
    long long s = 1;
    for (int d = nd - 1; d >= 0; d--) { G.stride[d] = s; s *= geom->nel[d] + 1; }
  }
  ElemSetView E;
  memset(&E, 0, sizeof(E));
  E.nsel = es->nsel;
  E.elem_ids = es->d_elem_ids;
  E.qoff = es->d_qoff;
  E.order = es->d_order;
  E.qcoords = es->d_qcoords;
  E.qweights = es->d_qweights;
  E.renumber = es->d_renumber;
  E.scale = es->d_scale;
  E.rational = es->rational;
  E.face_dim = es->d_face_dim;
  E.normals = es->d_normals;
  E.nq_uniform = Q.nqt;
  for (int k = 0; k < 2 * B2_MAX_FORMS; k++) {
    E.coef[k] = es->d_coef[k];
    E.field[k] = es->d_field[k];
    E.field_power[k] = es->field_power[k];
    E.field_scale[k] = es->field_scale[k];
    if (es->d_coef[k] && !es->d_qoff && es->coef_len[k] != es->nsel * (int64_t)Q.nqt) return b2_fail(ctx, B2_EINVAL, "coefficient array does not match nsel x points of the tensor rule");
  }
  E.nbasis_new = es->nbasis_new;
  for (int d = 0; d < nd; d++) E.coeffs[d] = es->d_coeffs[d];
  E.rowptr_b = pattern ? pattern->d_rowptr_b : nullptr;
  E.colidx_b = pattern ? pattern->d_colidx_b : nullptr;

  out->B = B;
  out->Q = Q;
  out->G = G;
  out->SG = SG;
  out->E = E;
  return B2_OK;
}

extern "C" int b2_assemble_elemset_device(b2_ctx* ctx, const b2_pattern* pattern, const b2_elemset* es, const b2_quad* quad, const b2_geom* geom,
                                          int64_t sel_begin, int64_t sel_end, int nmat, const double* const* D_host, double* const* values_dev,
                                          int nvec, const double* const* C_host, double* const* rhs_dev) {
  if (!ctx || !pattern || !es || !geom) return b2_fail(ctx, B2_EINVAL, "null argument");
  if (nmat < 0 || nmat > B2_MAX_FORMS || nvec < 0 || nvec > B2_MAX_FORMS) return b2_fail(ctx, B2_EINVAL, "0..4 matrix and vector forms per call");
  if ((nmat && (!D_host || !values_dev)) || (nvec && (!C_host || !rhs_dev))) return b2_fail(ctx, B2_EINVAL, "null form argument");
  if (pattern->elemset != es) return b2_fail(ctx, B2_EINVAL, "pattern was built for a different element set");
  const b2_basis* basis = es->basis;
  const int nd = basis->ndims, nc = basis->ncomp;
  if (!es->d_qoff && !quad) return b2_fail(ctx, B2_EINVAL, "the element set carries no points and no quadrature rule was given");
  if ((quad && quad->ndims != nd) || geom->ndims != nd) return b2_fail(ctx, B2_EINVAL, "dimension mismatch");
  for (int d = 0; d < nd; d++)
    if (geom->nel[d] != basis->nel[d]) return b2_fail(ctx, B2_EINVAL, "geometry and basis live on different topologies");
  if (es->rational == 2 && !(geom->gbasis && geom->d_wts)) return b2_fail(ctx, B2_EINVAL, "rational mode 2 needs a rational spline geometry");
  if (sel_end < 0) sel_end = es->nsel;
  if (sel_begin < 0 || sel_begin > sel_end || sel_end > es->nsel) return b2_fail(ctx, B2_EINVAL, "invalid element range");
  for (int m = 0; m < nmat; m++)
    if (!D_host[m] || !values_dev[m]) return b2_fail(ctx, B2_EINVAL, "null matrix form");
  for (int v = 0; v < nvec; v++)
    if (!C_host[v] || !rhs_dev[v]) return b2_fail(ctx, B2_EINVAL, "null vector form");
  if (sel_begin == sel_end || (nmat == 0 && nvec == 0)) return B2_OK;
  B2_CUDA(ctx, cudaSetDevice(ctx->device));

  ESViews V;
  {
    int rcv = build_views(ctx, pattern, es, quad, geom, &V);
    if (rcv != B2_OK) return rcv;
  }
  const BasisView& B = V.B;
  const QuadView& Q = V.Q;
  const GeomView& G = V.G;
  const SplineGeomView& SG = V.SG;
  const ElemSetView& E = V.E;

  FormView F;
  int rc = b2_upload_forms(ctx, nd, nc, nmat, D_host, values_dev, nvec, C_host, rhs_dev, &F);
  if (rc != B2_OK) return rc;
  return launch_assemble_elemset(ctx, B, Q, G, SG, E, F, sel_begin, sel_end, es->d_qoff ? es->max_nq : Q.nqt);
}

extern "C" int b2_assemble_elemset_host(b2_ctx* ctx, const b2_pattern* pattern, const b2_elemset* es, const b2_quad* quad, const b2_geom* geom,
                                        int nmat, const double* const* D_host, double* const* values_host,
                                        int nvec, const double* const* C_host, double* const* rhs_host) {
  if (!ctx || !pattern || !es) return B2_EINVAL;
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
  B2_CUDA(ctx, cudaMemsetAsync(ctx->scratch, 0, need, ctx->stream));
  double* vals[B2_MAX_FORMS];
  double* rhs[B2_MAX_FORMS];
  unsigned char* base = (unsigned char*)ctx->scratch;
  for (int m = 0; m < nmat; m++) vals[m] = (double*)(base + bm * m);
  for (int v = 0; v < nvec; v++) rhs[v] = (double*)(base + bm * nmat + bv * v);
  for (int m = 0; m < nmat; m++)
    if (!values_host[m]) return b2_fail(ctx, B2_EINVAL, "null output buffer");
  for (int v = 0; v < nvec; v++)
    if (!rhs_host[v]) return b2_fail(ctx, B2_EINVAL, "null output buffer");
  int rc = b2_assemble_elemset_device(ctx, pattern, es, quad, geom, 0, -1, nmat, D_host, vals, nvec, C_host, rhs);
  if (rc != B2_OK) return rc;
  for (int m = 0; m < nmat; m++) B2_CUDA(ctx, cudaMemcpyAsync(values_host[m], vals[m], bm, cudaMemcpyDeviceToHost, ctx->stream));
  for (int v = 0; v < nvec; v++) B2_CUDA(ctx, cudaMemcpyAsync(rhs_host[v], rhs[v], bv, cudaMemcpyDeviceToHost, ctx->stream));
  B2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return B2_OK;
}

extern "C" int b2_evaluate_elemset_device(b2_ctx* ctx, const b2_elemset* es, const b2_quad* quad, const b2_geom* geom, int nfields, const double* coef_dev,
                                          double* x_dev, double* wdet_dev, double* values_dev, double* grads_dev) {
  if (!ctx || !es || !geom) return b2_fail(ctx, B2_EINVAL, "null argument");
  if (nfields < 0 || (nfields && !coef_dev) || ((values_dev || grads_dev) && !nfields)) return b2_fail(ctx, B2_EINVAL, "fields without coefficients");
  B2_CUDA(ctx, cudaSetDevice(ctx->device));
  ESViews V;
  int rc = build_views(ctx, nullptr, es, quad, geom, &V);
  if (rc != B2_OK) return rc;
  if (es->nsel == 0) return B2_OK;
  return launch_evaluate_elemset(ctx, V.B, V.Q, V.G, V.SG, V.E, es->d_qoff ? es->max_nq : V.Q.nqt, nfields, coef_dev, es->nbasis_new * es->basis->ncomp, x_dev, wdet_dev,
                                 values_dev, grads_dev);
}
