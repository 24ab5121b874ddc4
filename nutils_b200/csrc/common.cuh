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

// Sparse representation of the coefficient tensors D_m: per (m, c, e) a list of (x, y, value).
struct FormView {
  int nmat, nvec;
  const int* termptr;     // [nmat*ncomp*ncomp + 1]
  const int* termxy;      // [nterms] x | y<<8
  const double* termval;  // [nterms]
  const double* vcoef;    // [nvec][ncomp][ndims+1]
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
  // small upload buffer for form descriptions
  void* formbuf = nullptr;
  size_t formbuf_bytes = 0;
  int64_t serial = 0;
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
};

struct b2_pattern {
  b2_ctx* ctx;
  const b2_basis* basis;
  int64_t nnz, nrows;
};

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
