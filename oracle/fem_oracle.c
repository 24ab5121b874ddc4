/*
 * CPU restatement (plain C, pthreads) of the reference's element-integration hot path.
 *
 * TEST INFRASTRUCTURE ONLY: loaded from tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs (as the checker and the reported CPU baseline).
 * Nothing under nutils_b200/ links, loads or calls this file.
 *
 * Reference: evalf/nutils @ 37d1cc5 (10a8), /root/reference/src/nutils.  The reference is pure
 * Python: evaluable.compile (evaluable.py:6532-6838) emits one Python `for` loop over elements
 * (listing: SURVEY.md appendix A) whose body is ~60 numpy calls, run by nproc forked workers that
 * pull element indices from a shared counter (parallel.py:27-154), followed by a serial
 * argsort/unique/bincount (evaluable.py:588-616, 5646-5682; numeric.py:434-460, 687-711).
 * This file restates the same arithmetic, in the same order of steps:
 *
 *   oracle_element_loop   per element:
 *     (ix,iy,iz) = unravel(ielem)                           transformseq.py:563-579
 *     per-dim coefficient rows, N and dN/dxi at the points  function.py:3080-3100, evaluable.py:4328-4374, 4584-4634
 *     dofs = C-order ravel of per-dim ranges (mod ndofs)    function.py:3087-3093
 *     J = dx/dxi of the degree-1 nodal geometry             mesh.py:55-57, function.py:1266-1295
 *     J^-1 (numeric.inv), det J, w |det J|                  numeric.py:221-242, evaluable.py:1403, 1463
 *     grad = dN J^-1                                        function.py:1207-1231
 *     block = sum_q w|J| integrand (full n_e x n_e einsum,  sample.py:951-956 and the generated einsums
 *             no symmetry shortcut, like the reference)
 *     append block / row dofs / col dofs to the COO arrays  evaluable.py:5383-5501
 *     rhs[dofs] += element vector                           evaluable.py:3582-3620 (numpy.add.at under a lock)
 *   oracle_sort_unique    flat = row*ncols+col; STABLE sort; unique mask; inverse
 *                                                           evaluable.py:593-609, 5577-5581, 5604-5640
 *   oracle_accumulate     bincount(inverse, weights=values) numeric.py:448-455
 *   oracle_compress_rows  rowptr from sorted rows           numeric.py:687-711
 *
 * The worker threads (pthreads) replace the reference's forked processes, with the same dynamic,
 * one-element-at-a-time scheduling from a shared counter; the sort stays serial like the reference's.
 *
 * Parity pin: tests/test_oracle_golden.py compares this library with tests/golden/<case>.npz (written by the
 * unmodified reference, oracle/make_golden.py): rowptr/colidx bit-exact, values <= 1e-13 relative.
 */

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

#define MAXD 3
#define MAXP 8

enum { FORM_MASS = 0, FORM_STIFFNESS = 1, FORM_ELASTICITY = 2, FORM_GENERIC = 3, FORM_LOAD = 4 };

typedef struct {
  int32_t ndims;
  int32_t ncomp;
  int64_t nelems[MAXD];
  int32_t degree[MAXD];
  int32_t nq[MAXD];
  int64_t ndofs_d[MAXD];
  const double* coeffs[MAXD];   /* [nsets][p+1][p+1], highest power first */
  const int32_t* setidx[MAXD];  /* [nelems_d] */
  const int64_t* start[MAXD];   /* [nelems_d] */
  const double* qpts[MAXD];     /* [nq_d] */
  const double* qwts[MAXD];     /* [nq_d] */
  const double* nodes;          /* [ndims][n0+1][n1+1][n2+1] */
} oracle_problem;

static double horner(const double* c, int p, double x) {
  double v = c[0];
  for (int k = 1; k <= p; k++) v = v * x + c[k];
  return v;
}

static double horner_deriv(const double* c, int p, double x) {
  if (p == 0) return 0.;
  double v = c[0] * p;
  for (int k = 1; k < p; k++) v = v * x + c[k] * (p - k);
  return v;
}

/* inverse and determinant of a small dense matrix (row-major n x n), n <= 3 */
static double inv_det(int n, const double* J, double* Ji) {
  if (n == 1) {
    Ji[0] = 1. / J[0];
    return J[0];
  }
  if (n == 2) {
    double det = J[0] * J[3] - J[1] * J[2];
    Ji[0] = J[3] / det; Ji[1] = -J[1] / det;
    Ji[2] = -J[2] / det; Ji[3] = J[0] / det;
    return det;
  }
  double c00 = J[4] * J[8] - J[5] * J[7];
  double c01 = J[5] * J[6] - J[3] * J[8];
  double c02 = J[3] * J[7] - J[4] * J[6];
  double det = J[0] * c00 + J[1] * c01 + J[2] * c02;
  Ji[0] = c00 / det; Ji[1] = (J[2] * J[7] - J[1] * J[8]) / det; Ji[2] = (J[1] * J[5] - J[2] * J[4]) / det;
  Ji[3] = c01 / det; Ji[4] = (J[0] * J[8] - J[2] * J[6]) / det; Ji[5] = (J[2] * J[3] - J[0] * J[5]) / det;
  Ji[6] = c02 / det; Ji[7] = (J[1] * J[6] - J[0] * J[7]) / det; Ji[8] = (J[0] * J[4] - J[1] * J[3]) / det;
  return det;
}

/*
 * Element loop.  Outputs (caller-allocated):
 *   values[m]  : double[nelems_total][ne][ne]   ne = prod(p_d+1) * ncomp          (m < nmat)
 *   rows, cols : int64[nelems_total][ne][ne]     (may be NULL to skip)
 *   rhs[v]     : double[ndofs], accumulated into (caller zero-fills)                (v < nvec)
 * mat_kind[m] / mat_param[m]: FORM_MASS, FORM_STIFFNESS (no params), FORM_ELASTICITY (lambda, mu),
 *   FORM_GENERIC (D[ncomp][ndims+1][ncomp][ndims+1]).
 * vec_kind[v] / vec_param[v]: FORM_LOAD (no params), FORM_GENERIC (c[ncomp][ndims+1]).
 * Elements [elem_begin, elem_end) are processed; outputs are indexed relative to elem_begin.
 * Returns 0, or -1 on invalid input / allocation failure.
 */
int oracle_max_threads(void);

typedef struct {
  const oracle_problem* pb;
  int64_t elem_begin, elem_end;
  int nmat; const int32_t* mat_kind; const double* const* mat_param; double* const* values;
  int64_t* rows; int64_t* cols;
  int nvec; const int32_t* vec_kind; const double* const* vec_param; double* const* rhs;
  atomic_llong* counter;  /* the shared element counter of parallel.py:128-146 */
  atomic_int* failed;
} loop_args;

static void atomic_add_double(double* addr, double v) {
  /* stands in for the reference's `with lock: numpy.add.at(...)` (evaluable.py:7116-7133) */
  uint64_t* p = (uint64_t*)addr;
  uint64_t old = __atomic_load_n(p, __ATOMIC_RELAXED), neu;
  do {
    double d;
    memcpy(&d, &old, 8);
    d += v;
    memcpy(&neu, &d, 8);
  } while (!__atomic_compare_exchange_n(p, &old, neu, 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
}

static void* loop_worker(void* argp) {
  const loop_args* A = (const loop_args*)argp;
  const oracle_problem* pb = A->pb;
  const int nd = pb->ndims, nc = pb->ncomp;
  int nb = 1, nqt = 1;
  for (int d = 0; d < nd; d++) { nb *= pb->degree[d] + 1; nqt *= pb->nq[d]; }
  const int ne = nb * nc;
  const int na = nd + 1;
  int64_t nstride[MAXD];
  {
    int64_t s = 1;
    for (int d = nd - 1; d >= 0; d--) { nstride[d] = s; s *= pb->nelems[d] + 1; }
  }
  int64_t nnodes = 1;
  for (int d = 0; d < nd; d++) nnodes *= pb->nelems[d] + 1;
  const int64_t elem_begin = A->elem_begin;
  const int nmat = A->nmat, nvec = A->nvec;
  const int32_t* mat_kind = A->mat_kind; const double* const* mat_param = A->mat_param; double* const* values = A->values;
  int64_t* rows = A->rows; int64_t* cols = A->cols;
  const int32_t* vec_kind = A->vec_kind; const double* const* vec_param = A->vec_param; double* const* rhs = A->rhs;
  double* N = (double*)malloc(sizeof(double) * nqt * nb);          /* [q][a] */
  double* G = (double*)malloc(sizeof(double) * nqt * nb * nd);     /* [q][a][i] physical gradient */
  double* B = (double*)malloc(sizeof(double) * nqt * nb * na);     /* [q][a][alpha] value+gradient */
  double* wdet = (double*)malloc(sizeof(double) * nqt);
  double* blk = (double*)malloc(sizeof(double) * ne * ne);
  int64_t* dofs = (int64_t*)malloc(sizeof(int64_t) * ne);
  double* vec = (double*)malloc(sizeof(double) * ne);
  if (!N || !G || !B || !wdet || !blk || !dofs || !vec) {
    atomic_store(A->failed, 1);
  } else
    for (;;) {
      const int64_t ielem = atomic_fetch_add(A->counter, 1);  /* one element per acquisition, like parallel.range */
      if (ielem >= A->elem_end) break;
      {
        /* (ix,iy,iz) */
        int64_t idx[MAXD], rem = ielem;
        for (int d = nd - 1; d >= 0; d--) { idx[d] = rem % pb->nelems[d]; rem /= pb->nelems[d]; }
        /* 1-D values and derivatives at the 1-D points */
        double val[MAXD][MAXP + 1][2 * MAXP + 2], der[MAXD][MAXP + 1][2 * MAXP + 2];
        for (int d = 0; d < nd; d++) {
          const int p = pb->degree[d];
          const double* c = pb->coeffs[d] + (int64_t)pb->setidx[d][idx[d]] * (p + 1) * (p + 1);
          for (int a = 0; a <= p; a++)
            for (int q = 0; q < pb->nq[d]; q++) {
              val[d][a][q] = horner(c + a * (p + 1), p, pb->qpts[d][q]);
              der[d][a][q] = horner_deriv(c + a * (p + 1), p, pb->qpts[d][q]);
            }
        }
        /* dofs: C-order ravel of per-dim ranges */
        for (int a = 0; a < nb; a++) {
          int ar = a;
          int ad[MAXD];
          for (int d = nd - 1; d >= 0; d--) { ad[d] = ar % (pb->degree[d] + 1); ar /= pb->degree[d] + 1; }
          int64_t dof = 0;
          for (int d = 0; d < nd; d++) dof = dof * pb->ndofs_d[d] + (pb->start[d][idx[d]] + ad[d]) % pb->ndofs_d[d];
          for (int c = 0; c < nc; c++) dofs[a * nc + c] = dof * nc + c;
        }
        /* vertex coordinates of the element, C-order vertices */
        double X[MAXD][8];
        int64_t base = 0;
        for (int d = 0; d < nd; d++) base += idx[d] * nstride[d];
        for (int v = 0; v < (1 << nd); v++) {
          int64_t off = base;
          for (int d = 0; d < nd; d++)
            if (v >> (nd - 1 - d) & 1) off += nstride[d];
          for (int i = 0; i < nd; i++) X[i][v] = pb->nodes[i * nnodes + off];
        }
        for (int q = 0; q < nqt; q++) {
          int qd[MAXD], qr = q;
          for (int d = nd - 1; d >= 0; d--) { qd[d] = qr % pb->nq[d]; qr /= pb->nq[d]; }
          double w = 1.;
          for (int d = 0; d < nd; d++) w *= pb->qwts[d][qd[d]];
          /* J[i][k] = sum_v X[i][v] d phi_v / d xi_k */
          double J[9], Ji[9];
          for (int i = 0; i < nd; i++)
            for (int k = 0; k < nd; k++) {
              double s = 0.;
              for (int v = 0; v < (1 << nd); v++) {
                double f = 1.;
                for (int d = 0; d < nd; d++) {
                  int bit = v >> (nd - 1 - d) & 1;
                  double xi = pb->qpts[d][qd[d]];
                  f *= d == k ? (bit ? 1. : -1.) : (bit ? xi : 1. - xi);
                }
                s += X[i][v] * f;
              }
              J[i * nd + k] = s;
            }
          double det = inv_det(nd, J, Ji);
          wdet[q] = w * fabs(det);
          for (int a = 0; a < nb; a++) {
            int ar = a, ad[MAXD];
            for (int d = nd - 1; d >= 0; d--) { ad[d] = ar % (pb->degree[d] + 1); ar /= pb->degree[d] + 1; }
            double v = 1., dxi[MAXD];
            for (int d = 0; d < nd; d++) v *= val[d][ad[d]][qd[d]];
            for (int k = 0; k < nd; k++) {
              double g = 1.;
              for (int d = 0; d < nd; d++) g *= d == k ? der[d][ad[d]][qd[d]] : val[d][ad[d]][qd[d]];
              dxi[k] = g;
            }
            N[q * nb + a] = v;
            B[(q * nb + a) * na] = v;
            for (int i = 0; i < nd; i++) {
              double s = 0.;
              for (int k = 0; k < nd; k++) s += dxi[k] * Ji[k * nd + i];
              G[(q * nb + a) * nd + i] = s;
              B[(q * nb + a) * na + 1 + i] = s;
            }
          }
        }
        const int64_t eoff = (ielem - elem_begin) * (int64_t)ne * ne;
        for (int m = 0; m < nmat; m++) {
          memset(blk, 0, sizeof(double) * ne * ne);
          const int kind = mat_kind[m];
          const double* par = mat_param ? mat_param[m] : NULL;
          for (int q = 0; q < nqt; q++) {
            const double wq = wdet[q];
            for (int a = 0; a < nb; a++)
              for (int b = 0; b < nb; b++) {
                const double* ga = G + (q * nb + a) * nd;
                const double* gb = G + (q * nb + b) * nd;
                if (kind == FORM_MASS) {
                  blk[a * ne + b] += wq * N[q * nb + a] * N[q * nb + b];
                } else if (kind == FORM_STIFFNESS) {
                  double s = 0.;
                  for (int i = 0; i < nd; i++) s += ga[i] * gb[i];
                  blk[a * ne + b] += wq * s;
                } else if (kind == FORM_ELASTICITY) {
                  /* d2/du2 of eps:sigma = 2 (lambda div v div u + 2 mu eps(v):eps(u)) */
                  const double lm = par[0], mu = par[1];
                  double dot = 0.;
                  for (int i = 0; i < nd; i++) dot += ga[i] * gb[i];
                  for (int c = 0; c < nd; c++)
                    for (int e = 0; e < nd; e++)
                      blk[(a * nc + c) * ne + b * nc + e] += 2. * wq * (lm * ga[c] * gb[e] + mu * ga[e] * gb[c] + (c == e ? mu * dot : 0.));
                } else { /* FORM_GENERIC */
                  const double* Ba = B + (q * nb + a) * na;
                  const double* Bb = B + (q * nb + b) * na;
                  for (int c = 0; c < nc; c++)
                    for (int e = 0; e < nc; e++) {
                      double s = 0.;
                      for (int x = 0; x < na; x++)
                        for (int y = 0; y < na; y++) s += Ba[x] * par[((c * na + x) * nc + e) * na + y] * Bb[y];
                      blk[(a * nc + c) * ne + b * nc + e] += wq * s;
                    }
                }
              }
          }
          memcpy(values[m] + eoff, blk, sizeof(double) * ne * ne);
        }
        if (rows && cols)
          for (int r = 0; r < ne; r++)
            for (int c = 0; c < ne; c++) {
              rows[eoff + r * ne + c] = dofs[r];
              cols[eoff + r * ne + c] = dofs[c];
            }
        for (int v = 0; v < nvec; v++) {
          const double* par = vec_param ? vec_param[v] : NULL;
          for (int r = 0; r < ne; r++) vec[r] = 0.;
          for (int q = 0; q < nqt; q++)
            for (int a = 0; a < nb; a++) {
              if (vec_kind[v] == FORM_LOAD) {
                vec[a] += wdet[q] * N[q * nb + a];
              } else {
                const double* Ba = B + (q * nb + a) * na;
                for (int c = 0; c < nc; c++) {
                  double s = 0.;
                  for (int x = 0; x < na; x++) s += par[c * na + x] * Ba[x];
                  vec[a * nc + c] += wdet[q] * s;
                }
              }
            }
          for (int r = 0; r < ne; r++) {
            atomic_add_double(&rhs[v][dofs[r]], vec[r]);
          }
        }
      }
    }
  free(N); free(G); free(B); free(wdet); free(blk); free(dofs); free(vec);
  return NULL;
}

int oracle_element_loop(const oracle_problem* pb, int64_t elem_begin, int64_t elem_end,
                        int nmat, const int32_t* mat_kind, const double* const* mat_param, double* const* values,
                        int64_t* rows, int64_t* cols,
                        int nvec, const int32_t* vec_kind, const double* const* vec_param, double* const* rhs,
                        int nthreads) {
  if (pb->ndims < 1 || pb->ndims > MAXD || pb->ncomp < 1) return -1;
  for (int d = 0; d < pb->ndims; d++)
    if (pb->degree[d] > MAXP || pb->nq[d] > 2 * MAXP + 2) return -1;
  if (nthreads <= 0) nthreads = oracle_max_threads();
  if (nthreads > 1024) nthreads = 1024;
  atomic_llong counter;
  atomic_int failed;
  atomic_init(&counter, elem_begin);
  atomic_init(&failed, 0);
  loop_args A = {pb, elem_begin, elem_end, nmat, mat_kind, mat_param, values, rows, cols, nvec, vec_kind, vec_param, rhs, &counter, &failed};
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
  if (!th) return -1;
  int started = 0;
  for (int t = 1; t < nthreads; t++)
    if (pthread_create(&th[started], NULL, loop_worker, &A) == 0) started++;
  loop_worker(&A);
  for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
  free(th);
  return atomic_load(&failed) ? -1 : 0;
}

/*
 * flat = rows*ncols+cols; stable sort; unique.  Outputs: inverse[n] (position of every COO entry
 * in the unique list), ukeys[<=n] (sorted unique flat keys).  Returns nnz, or -1.
 * Stable LSD radix sort on 16-bit digits == numpy.argsort(kind='stable') on the same keys.
 */
int64_t oracle_sort_unique(int64_t n, const int64_t* rows, const int64_t* cols, int64_t ncols, int64_t* inverse, int64_t* ukeys) {
  if (n == 0) return 0;
  uint64_t* key = (uint64_t*)malloc(sizeof(uint64_t) * n);
  uint64_t* key2 = (uint64_t*)malloc(sizeof(uint64_t) * n);
  int64_t* ord = (int64_t*)malloc(sizeof(int64_t) * n);
  int64_t* ord2 = (int64_t*)malloc(sizeof(int64_t) * n);
  size_t* hist = (size_t*)malloc(sizeof(size_t) * 65536);
  if (!key || !key2 || !ord || !ord2 || !hist) { free(key); free(key2); free(ord); free(ord2); free(hist); return -1; }
  uint64_t maxkey = 0;
  for (int64_t i = 0; i < n; i++) {
    key[i] = (uint64_t)(rows[i] * ncols + cols[i]);
    ord[i] = i;
    if (key[i] > maxkey) maxkey = key[i];
  }
  for (int shift = 0; shift < 64 && (maxkey >> shift) != 0; shift += 16) {
    memset(hist, 0, sizeof(size_t) * 65536);
    for (int64_t i = 0; i < n; i++) hist[(key[i] >> shift) & 0xffff]++;
    size_t sum = 0;
    for (int b = 0; b < 65536; b++) { size_t c = hist[b]; hist[b] = sum; sum += c; }
    for (int64_t i = 0; i < n; i++) {
      size_t pos = hist[(key[i] >> shift) & 0xffff]++;
      key2[pos] = key[i];
      ord2[pos] = ord[i];
    }
    uint64_t* tk = key; key = key2; key2 = tk;
    int64_t* to = ord; ord = ord2; ord2 = to;
  }
  int64_t nnz = 0;
  for (int64_t i = 0; i < n; i++) {
    if (i == 0 || key[i] != key[i - 1]) ukeys[nnz++] = (int64_t)key[i];
    inverse[ord[i]] = nnz - 1;
  }
  free(key); free(key2); free(ord); free(ord2); free(hist);
  return nnz;
}

/* numpy.bincount(inverse, weights=values, minlength=nnz): sequential, in array order */
void oracle_accumulate(int64_t n, const int64_t* inverse, const double* values, int64_t nnz, double* out) {
  memset(out, 0, sizeof(double) * nnz);
  for (int64_t i = 0; i < n; i++) out[inverse[i]] += values[i];
}

/* ukeys -> rowptr[nrows+1], colidx[nnz] */
void oracle_compress_rows(int64_t nnz, const int64_t* ukeys, int64_t nrows, int64_t ncols, int64_t* rowptr, int64_t* colidx) {
  int64_t r = 0;
  rowptr[0] = 0;
  for (int64_t i = 0; i < nnz; i++) {
    int64_t row = ukeys[i] / ncols;
    colidx[i] = ukeys[i] % ncols;
    while (r < row) rowptr[++r] = i;
  }
  while (r < nrows) rowptr[++r] = nnz;
}

int oracle_max_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}
