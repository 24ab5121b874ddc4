// Shared host/device definitions of the b200fem CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <utility>
#include <vector>

#include "../../include/b200fem.h"

#define B2_MAXD B2_MAX_DIMS

// ---- device-side views ---------------------------------------------------------------------------

// Per-dimension structure of a tensor-product spline space, resident in HBM.
struct BasisView {
  int ndims, ncomp;
  int p[B2_MAXD];          // degree
  int nel[B2_MAXD];        // elements
  int ndofs[B2_MAXD];      // dofs
  const int* setidx[B2_MAXD];  // [nel]    coefficient set of the element
  const int* start[B2_MAXD];   // [nel]    first dof of the element
  // column structure of the CSR pattern per dof index along d (see pattern.cu)
  const int* lo[B2_MAXD];      // [ndofs]  first coupled dof
  const int* wid[B2_MAXD];     // [ndofs]  number of coupled dofs
  const int* cum[B2_MAXD];     // [ndofs+1] exclusive prefix sum of wid
  long long W[B2_MAXD];        // sum of wid over the dimension
  int nb;                      // prod (p+1): basis functions per element
  int nsets[B2_MAXD];          // number of coefficient sets per dimension
};

struct QuadView {
  int nq[B2_MAXD];
  int nqt;
  const double* x[B2_MAXD];
  const double* w[B2_MAXD];
  // tabulated 1-D basis values/derivatives at the 1-D points: [nsets][2][p+1][nq]
  const double* tab[B2_MAXD];
};

struct GeomView {
  const double* nodes;       // [ndims][nnodes]
  long long nnodes;
  long long stride[B2_MAXD]; // node stride per dimension
};

// Spline (optionally rational) geometry: control points in the scalar basis GB (same element grid).
struct SplineGeomView {
  int enabled;               // 0: nodal multilinear geometry (GeomView)
  int rational;              // weights present: x = sum B w X / sum B w
  BasisView GB;
  const double* coeffs[B2_MAXD];  // [nsets][p+1][p+1] local polynomials of GB per dimension (highest power first)
  const double* ctrl;        // [ndims][nbasis_g]
  const double* wts;         // [nbasis_g] or null
  long long nbasis;
};

// Element set: subset of elements, ragged points, pruned numbering, rational scaling (include/b200fem.h).
struct ElemSetView {
  long long nsel;
  const long long* elem_ids;   // [nsel] or null (all elements)
  const long long* qoff;       // [nsel+1] or null (tensor rule)
  const long long* order;      // [nsel] elements by decreasing cost, or null
  const double* qcoords;       // [npoints][ndims]
  const double* qweights;      // [npoints]
  const int* renumber;         // [nbasis_parent] -> new index, < 0: dropped; null = identity
  const double* scale;         // [nbasis_parent] or null
  int rational;                // 0 none, 1 own weight function, 2 the geometry's weight function
  const double* normals;       // [npoints][ndims] scaled reference normals of the points' facets (immersed boundaries), or null
  const signed char* face_dim; // [nsel] -1 volume, k: points on a face normal to reference direction k (surface measure); null = volume
  const double* coef[2 * B2_MAX_FORMS];  // per-point scalar coefficient of matrix form m / vector form B2_MAX_FORMS + v, or null
  // solution-dependent coefficient of form k: scale * u_h(x_q)^power with u_h = sum_i field[i] N_i evaluated in the kernel
  const double* field[2 * B2_MAX_FORMS];
  int field_power[2 * B2_MAX_FORMS];
  double field_scale[2 * B2_MAX_FORMS];
  long long nq_uniform;        // points per element of the tensor rule (indexing of coef when qoff is null)
  long long nbasis_new;
  const double* coeffs[B2_MAXD];  // local polynomials of the solution basis per dimension
  // general CSR pattern at basis level (rows = new basis indices)
  const long long* rowptr_b;   // [nbasis_new+1]
  const int* colidx_b;         // [nnz_b] new basis indices, sorted per row
};

// Sparse representation of the coefficient tensors D_m: per (m, c, e) a list of (x, y, value).
struct FormView {
  int nmat, nvec;
  const int* termptr;     // [nmat*ncomp*ncomp + 1]
  const int* termxy;      // [nterms] x | y<<8
  const double* termval;  // [nterms]
  const double* vcoef;    // [nvec][ncomp][ndims+1]
  // dense view of the same tensors for the tensor-core path: job j = (matrix m, row component c, column component e) with a
  // non-zero block D_m[c][.][e][.]; jobD[j][x*4+y] (zero padded to 4x4), jobinfo[j] = m | c << 8 | e << 16 | mode << 24
  // (mode 1: only the value-value entry D[0][0] is non-zero -- mass-like)
  const double* jobD;
  const int* jobinfo;
  int njobs;
  double* values[B2_MAX_FORMS];
  double* rhs[B2_MAX_FORMS];
};

// ---- host-side objects ---------------------------------------------------------------------------

struct b2_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // second stream for the device-to-host copies of b2_assemble_host (overlap with the next chunk's kernel)
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t copy_event = nullptr;
  std::string err;
  int64_t launches = 0;
  void* flush_buf = nullptr;
  size_t flush_bytes = 0;
  std::map<std::string, int64_t> opts;
  // scratch for b2_assemble_host
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
  // per-point geometry coefficients of the owner-computes path (k_geom3d -> TMA loads of k_rows3d)
  void* gbuf = nullptr;
  size_t gbuf_bytes = 0;
  // small upload buffer for form descriptions
  void* formbuf = nullptr;
  size_t formbuf_bytes = 0;
  int64_t serial = 0;
  // element queue head of the element-set kernel (dynamic hand-out of elements of very different cost)
  unsigned long long* queue = nullptr;
  // optional per-kernel timing (option "time_kernels")
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> kernel_events;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> event_pool;
};

// brackets one kernel launch with events when the context option "time_kernels" is set
struct KernelTimer {
  b2_ctx* ctx;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  explicit KernelTimer(b2_ctx* c);
  ~KernelTimer();
};

struct TabDev {
  double* tab[B2_MAXD] = {nullptr, nullptr, nullptr};
};

struct b2_basis {
  b2_ctx* ctx;
  int ndims, ncomp;
  int p[B2_MAXD];
  int64_t nel[B2_MAXD];
  int64_t ndofs_d[B2_MAXD];
  int nsets[B2_MAXD];
  std::vector<double> coeffs[B2_MAXD];
  std::vector<int> setidx[B2_MAXD];
  std::vector<int> start[B2_MAXD];
  std::vector<int> lo[B2_MAXD], wid[B2_MAXD], cum[B2_MAXD];
  int* d_setidx[B2_MAXD];
  int* d_start[B2_MAXD];
  int* d_lo[B2_MAXD];
  int* d_wid[B2_MAXD];
  int* d_cum[B2_MAXD];
  long long W[B2_MAXD];
  int64_t nbasis;
  std::map<int64_t, TabDev> tabs;  // keyed by quad serial
  BasisView view() const;
};

struct b2_quad {
  b2_ctx* ctx;
  int64_t serial;
  int ndims;
  int nq[B2_MAXD];
  std::vector<double> pts[B2_MAXD], wts[B2_MAXD];
  double* d_x[B2_MAXD];
  double* d_w[B2_MAXD];
};

struct b2_geom {
  b2_ctx* ctx;
  int ndims;
  int64_t nel[B2_MAXD];
  int64_t nnodes;
  double* d_nodes;
  // spline geometry (b2_geom_create_spline): control points in gbasis
  const b2_basis* gbasis = nullptr;
  double* d_ctrl = nullptr;
  double* d_wts = nullptr;
  double* d_gcoeffs[B2_MAXD] = {nullptr, nullptr, nullptr};
};

struct b2_elemset {
  b2_ctx* ctx;
  const b2_basis* basis;
  int64_t nsel, npoints, nbasis_new;
  int rational;
  int max_nq;                // largest number of points of one element (ragged), 0 for tensor rules
  std::vector<int64_t> elem_ids;   // host copy (empty = all)
  std::vector<int> renumber;       // host copy (empty = identity)
  long long* d_elem_ids = nullptr;
  long long* d_qoff = nullptr;
  long long* d_order = nullptr;   // ragged point sets: the elements by decreasing number of points (hand-out order of the element queue)
  double* d_qcoords = nullptr;
  double* d_qweights = nullptr;
  int* d_renumber = nullptr;
  double* d_scale = nullptr;
  double* d_coeffs[B2_MAXD] = {nullptr, nullptr, nullptr};
  signed char* d_face_dim = nullptr;
  double* d_normals = nullptr;
  double* d_coef[2 * B2_MAX_FORMS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int64_t coef_len[2 * B2_MAX_FORMS] = {0, 0, 0, 0, 0, 0, 0, 0};
  const double* d_field[2 * B2_MAX_FORMS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // caller-owned device vectors
  int field_power[2 * B2_MAX_FORMS] = {0, 0, 0, 0, 0, 0, 0, 0};
  double field_scale[2 * B2_MAX_FORMS] = {1., 1., 1., 1., 1., 1., 1., 1.};
  unsigned char* d_selmask = nullptr;  // [ntotal] 1 = element selected (pattern construction); null = all
  long long* d_dofmap = nullptr;       // [nbasis_new] parent index of each kept basis function; null = identity
  int* d_efirst[B2_MAXD] = {nullptr, nullptr, nullptr};  // [ndofs_d] first / last supporting element of the 1-D functions
  int* d_elast[B2_MAXD] = {nullptr, nullptr, nullptr};
};

struct b2_pattern {
  b2_ctx* ctx;
  const b2_basis* basis;
  int64_t nnz, nrows;
  // explicit patterns (b2_pattern_create_csr, b2_pattern_general): basis == nullptr, rows at basis level in d_rowptr_b /
  // d_colidx_b with ncomp_gen components per basis function, ncols columns
  int64_t ncols = 0;
  int ncomp_gen = 1;
  // general pattern of an element set (b2_pattern_create_elemset); analytic otherwise
  const b2_elemset* elemset = nullptr;
  long long* d_rowptr_b = nullptr;  // [nbasis_new+1]
  int* d_colidx_b = nullptr;        // [nnz_b]
  std::vector<long long> rowptr_b;  // host copy
  // analytic patterns: lazily built element set "everything" + its materialised pattern (same slots as the analytic one), so
  // that the coverage path can run the tensor-core element-set kernel instead of the scalar generic kernel
  b2_elemset* aux_elemset = nullptr;
  b2_pattern* aux_pattern = nullptr;
  bool aux_failed = false;
};

// components per basis function / is the pattern materialised (element sets, explicit patterns)?
inline int pattern_ncomp(const b2_pattern* p) { return p->basis ? p->basis->ncomp : p->ncomp_gen; }
inline bool pattern_general(const b2_pattern* p) { return p->elemset != nullptr || p->basis == nullptr; }

// ---- error helpers -------------------------------------------------------------------------------

int b2_fail(b2_ctx* ctx, int code, const std::string& msg);
int b2_cuda_fail(b2_ctx* ctx, cudaError_t e, const char* what);

#define B2_CUDA(ctx, call)                                   \
  do {                                                       \
    cudaError_t e__ = (call);                                \
    if (e__ != cudaSuccess) return b2_cuda_fail(ctx, e__, #call); \
  } while (0)

// ---- kernel launchers (defined in the .cu files) -------------------------------------------------

int launch_pattern_export(b2_ctx* ctx, const BasisView& B, long long* rowptr, long long* colidx);
// only rows whose dimension-0 dof index lies in [plane_lo, plane_hi) are scattered (0, INT_MAX = all)
int launch_assemble_generic(b2_ctx* ctx, const BasisView& B, const QuadView& Q, const GeomView& G, const FormView& F,
                            long long elem_begin, long long elem_end, int plane_lo, int plane_hi);
// returns B2_EUNSUPPORTED when no specialised kernel covers the request
int launch_assemble_fast(b2_ctx* ctx, const BasisView& B, const QuadView& Q, const GeomView& G, const FormView& F,
                         const double* const* D_host, const double* const* C_host, long long elem_begin, long long elem_end);

// element-set path (assemble_elemset.cu, pattern_elemset.cu)
int launch_assemble_elemset(b2_ctx* ctx, const BasisView& B, const QuadView& Q, const GeomView& G, const SplineGeomView& SG, const ElemSetView& E, const FormView& F,
                            long long sel_begin, long long sel_end, int max_nq);
int launch_evaluate_elemset(b2_ctx* ctx, const BasisView& B, const QuadView& Q, const GeomView& G, const SplineGeomView& SG, const ElemSetView& E, int max_nq,
                            int nfields, const double* coef, long long ndofs, double* x, double* wdet, double* values, double* grads);
// counts[new basis row] = number of coupled columns (pass 0) / fills colidx_b (pass 1)
int launch_pattern_elemset_impl(b2_ctx* ctx, const BasisView& B, const int* const* efirst, const int* const* elast, const long long* dofmap, const int* renumber,
                                const unsigned char* selmask, long long nbasis_new, int pass, long long* counts_or_rowptr, int* colidx_b);
int launch_exclusive_scan(b2_ctx* ctx, long long* data, long long n);  // in place, data[n] receives the total (n+1 entries)
int launch_pattern_export_general(b2_ctx* ctx, const long long* rowptr_b, const int* colidx_b, long long nbasis, int ncomp, long long* rowptr, long long* colidx);

// owner-computes kernel: WRITES every stored value of the dof planes [plane_begin, plane_end) of dimension 0
int launch_assemble_rows(b2_ctx* ctx, const b2_basis* basis, const b2_quad* quad, const BasisView& B, const QuadView& Q, const GeomView& G, const FormView& F,
                         const double* const* D_host, const double* const* C_host, long long plane_begin, long long plane_end);

// ---- device helpers ------------------------------------------------------------------------------

// CSR slot arithmetic for the analytic tensor-product pattern.
//   basis row I = (i_0, i_1, i_2): first entry  R(I) = cum0[i0] W1 W2 + wid0[i0] (cum1[i1] W2 + wid1[i1] cum2[i2])
//   (in units of basis columns), width w(I) = prod wid_d[i_d];
//   dof row (I, c) starts at ncomp^2 R(I) + c ncomp w(I); column (J, e) sits at pos(J) ncomp + e with
//   pos(J) = ((j0-lo0) wid1 + (j1-lo1)) wid2 + (j2-lo2).
template <int DIM>
__device__ __forceinline__ long long row_start_basis(const BasisView& B, const int* i) {
  long long r = B.cum[0][i[0]];
  if (DIM > 1) r = r * B.W[1] + (long long)B.wid[0][i[0]] * B.cum[1][i[1]];
  if (DIM > 2) r = r * B.W[2] + (long long)B.wid[0][i[0]] * B.wid[1][i[1]] * B.cum[2][i[2]];
  return r;
}
