// Device-resident matrix operations on the assembled CSR values: y = A x, diag(A), and a Jacobi-preconditioned
// conjugate-gradient solve with constraints -- so that a matrix assembled in HBM never has to cross PCIe.
//
// What it replaces in the reference: nutils.matrix.Matrix.__matmul__ / diagonal / solve with
// constrain + lhs0 (src/nutils/matrix/_base.py:62-75, 100-173: "solve A dx = b - A lhs0 on the free dofs", tolerance
// |A x - b| <= max(atol, rtol |b_reduced|)), as used by solver.System.solve (solver.py:318-425) right after assembly.
//
// The structured pattern is ANALYTIC: the SpMV kernel never reads a column index.  One warp owns the rows of one
// basis function; lanes stride over the row's box of coupled dofs (decoded with multiply-shift divisions by the box
// widths), values are streamed with evict-first loads (read once: 8 B per stored entry is the algorithmic traffic and
// the HBM bound), x is gathered through L2.  Element-set patterns use their materialised basis-level column list.

#include <algorithm>
#include <cmath>
#include <cstring>

#include "common.cuh"

namespace {

// evict-first load of a value that is read exactly once (keeps x resident in L2)
__device__ __forceinline__ double ld_stream(const double* p) { return __ldcs(p); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-level sum of one double per thread into *target (one atomic per block)
__device__ __forceinline__ void block_accumulate(double v, double* target) {
  __shared__ double sh[32];
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = lane < nw ? sh[lane] : 0.;
    v = warp_sum(v);
    if (lane == 0 && v != 0.) atomicAdd(target, v);
  }
}

struct RowBox {
  long long base;   // first slot of dof row (I, 0)
  int lo[3], wid[3], w;
  unsigned m1, m2;  // multiply-shift reciprocals of wid[1], wid[2]: n / d == (n * m) >> 20 for n < 8192, d <= 9
};

template <int DIM>
__device__ __forceinline__ RowBox row_box(const BasisView& B, long long I, int* idx) {
  RowBox r;
  int i[3] = {0, 0, 0};
  for (int d = DIM - 1; d >= 0; d--) {
    i[d] = (int)(I % B.ndofs[d]);
    I /= B.ndofs[d];
  }
  for (int d = 0; d < 3; d++) {
    r.lo[d] = d < DIM ? B.lo[d][i[d]] : 0;
    r.wid[d] = d < DIM ? B.wid[d][i[d]] : 1;
    idx[d] = i[d];
  }
  r.w = r.wid[0] * r.wid[1] * r.wid[2];
  r.base = row_start_basis<DIM>(B, i) * B.ncomp * B.ncomp;
  r.m1 = ((1u << 20) + r.wid[1] - 1) / max(r.wid[1], 1);
  r.m2 = ((1u << 20) + r.wid[2] - 1) / max(r.wid[2], 1);
  return r;
}

// column (basis index) of position pos in the row's box, C order
template <int DIM>
__device__ __forceinline__ long long box_column(const BasisView& B, const RowBox& r, int pos) {
  if (DIM == 1) return r.lo[0] + pos;
  if (DIM == 2) {
    const int q = (int)(((unsigned)pos * r.m1) >> 20);
    return (long long)(r.lo[0] + q) * B.ndofs[1] + r.lo[1] + (pos - q * r.wid[1]);
  }
  const int q2 = (int)(((unsigned)pos * r.m2) >> 20), j2 = pos - q2 * r.wid[2];
  const int q1 = (int)(((unsigned)q2 * r.m1) >> 20), j1 = q2 - q1 * r.wid[1];
  return ((long long)(r.lo[0] + q1) * B.ndofs[1] + r.lo[1] + j1) * B.ndofs[2] + r.lo[2] + j2;
}

// y = A x.  mask (optional): rows with mask != 0 are constrained: y = 0 there (the caller keeps x = 0 on constrained
// columns, so the product is the one of the free-free submatrix).  dot (optional): *dot += x . y.
// all ncomp dof rows of basis function I, one warp (any pattern shape; the coverage path of the kernels below)
template <int DIM>
__device__ __noinline__ void spmv_basis_row(const BasisView& B, const long long I, const double* __restrict__ values, const double* __restrict__ x,
                                            double* __restrict__ y, const unsigned char* __restrict__ mask, const bool want_dot, double& local) {
  const int lane = threadIdx.x & 31;
  const int nc = B.ncomp;
  int idx[3];
  const RowBox r = row_box<DIM>(B, I, idx);
  const int rowlen = r.w * nc;
  for (int c = 0; c < nc; c++) {
    const long long row = I * nc + c;
    if (mask && mask[row]) {
      if (lane == 0) y[row] = 0.;
      continue;
    }
    const double* v = values + r.base + (long long)c * rowlen;
    double s = 0.;
    for (int k0 = 0; k0 < rowlen; k0 += 128) {
      double a[4], b[4];
#pragma unroll
      for (int t = 0; t < 4; t++) {
        const int k = k0 + t * 32 + lane;
        a[t] = k < rowlen ? ld_stream(v + k) : 0.;
      }
#pragma unroll
      for (int t = 0; t < 4; t++) {
        const int k = k0 + t * 32 + lane;
        b[t] = 0.;
        if (k < rowlen) {
          if (nc == 1) b[t] = x[box_column<DIM>(B, r, k)];
          else {
            const int pos = k / nc, e = k - pos * nc;
            b[t] = x[box_column<DIM>(B, r, pos) * nc + e];
          }
        }
      }
#pragma unroll
      for (int t = 0; t < 4; t++) s = fma(a[t], b[t], s);
    }
    s = warp_sum(s);
    if (lane == 0) {
      y[row] = s;
      if (want_dot) local = fma(x[row], s, local);
    }
  }
}

// y = A x.  mask (optional): rows with mask != 0 are constrained: y = 0 there (the caller keeps x = 0 on constrained
// columns, so the product is the one of the free-free submatrix).  dot (optional): *dot += x . y.
template <int DIM>
__global__ void __launch_bounds__(256) k_spmv(const BasisView B, const long long nbasis, const double* __restrict__ values, const double* __restrict__ x,
                                               double* __restrict__ y, const unsigned char* __restrict__ mask, double* dot) {
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  double local = 0.;
  for (long long I = warp0; I < nbasis; I += nwarps) spmv_basis_row<DIM>(B, I, values, x, y, mask, dot != nullptr, local);
  if (dot) block_accumulate(local, dot);
}

// Scalar spaces with rows of at most 128 entries (degree <= 2 in 3-D): RUNS of eight consecutive rows per warp.
// The plain kernel is bound by instruction issue, not by memory (ncu: issue slots 75 % busy at 24 % of the DRAM bandwidth:
// ~440 instructions per row for index decoding, slot arithmetic and a 5-step butterfly per row).  Eight consecutive rows
// along the last dimension whose boxes are translates of each other (all interior rows) share everything but a shift:
// values base += w, first column += 1 -- so per row a lane issues 4 streamed value loads + 4 gathers of x (L1) + 4 DFMA,
// and the eight row sums are reduced TOGETHER by a transposing butterfly (9 shuffles for 8 rows instead of 40).  Runs that
// touch the boundary of the box structure take the plain per-row code.
template <int DIM, int NCH>
__global__ void __launch_bounds__(256, 3) k_spmv_run(const BasisView B, const int nbasis, const double* __restrict__ values, const double* __restrict__ x,
                                                      double* __restrict__ y, const unsigned char* __restrict__ mask, double* dot) {
  constexpr int LD = DIM - 1, R = 8;
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nd1 = DIM > 1 ? B.ndofs[1] : 1, nd2 = DIM > 2 ? B.ndofs[2] : 1, ndl = B.ndofs[LD];
  double local = 0.;
  int sig = -1, coff[NCH];
#pragma unroll
  for (int t = 0; t < NCH; t++) coff[t] = 0;
  for (long long g = blockIdx.x; g * (8 * R) < nbasis; g += gridDim.x) {
    const int I0 = (int)(g * (8 * R)) + warp * R;
    if (I0 >= nbasis) continue;
    int i[3] = {0, 0, 0};
    {
      unsigned r = (unsigned)I0;
      if (DIM > 2) { i[2] = r % (unsigned)nd2; r /= (unsigned)nd2; }
      if (DIM > 1) { i[1] = r % (unsigned)nd1; r /= (unsigned)nd1; }
      i[0] = r;
    }
    // are the boxes of the eight rows translates of the first one along the last dimension?
    bool run = I0 + R <= nbasis && i[LD] + R <= ndl;
    if (run) {
      const int rr = lane & (R - 1);
      const int lo_l = B.lo[LD][i[LD] + rr], wid_l = B.wid[LD][i[LD] + rr];
      const int lo_0 = __shfl_sync(FULL, lo_l, 0), wid_0 = __shfl_sync(FULL, wid_l, 0);
      run = __all_sync(FULL, lo_l == lo_0 + rr && wid_l == wid_0 && wid_l > 0);
    }
    if (!run) {
      const int nrows = min(R, nbasis - I0);
      for (int r = 0; r < nrows; r++) spmv_basis_row<DIM>(B, I0 + r, values, x, y, mask, dot != nullptr, local);
      continue;
    }
    int lo[3] = {0, 0, 0}, wid[3] = {1, 1, 1};
#pragma unroll
    for (int d = 0; d < DIM; d++) {
      lo[d] = B.lo[d][i[d]];
      wid[d] = B.wid[d][i[d]];
    }
    const int w = wid[0] * wid[1] * wid[2];
    const int nsig = wid[0] | wid[1] << 8 | wid[2] << 16;
    if (nsig != sig) {
      sig = nsig;
#pragma unroll
      for (int t = 0; t < NCH; t++) {
        const int k = min(t * 32 + lane, max(w - 1, 0));
        const int j2 = k % wid[2], q = k / wid[2], j1 = q % wid[1], j0 = q / wid[1];
        coff[t] = (j0 * nd1 + j1) * nd2 + j2;
      }
    }
    if (w == 0) {  // functions without support: empty rows
      if (lane < R) y[I0 + lane] = 0.;
      continue;
    }
    const long long base = row_start_basis<DIM>(B, i);
    const long long c0 = ((long long)lo[0] * nd1 + lo[1]) * nd2 + lo[2];
    bool ok[NCH];
    const double* pv[NCH];
    const double* px[NCH];
#pragma unroll
    for (int t = 0; t < NCH; t++) {
      ok[t] = t * 32 + lane < w;
      pv[t] = values + base + (ok[t] ? t * 32 + lane : w - 1);  // clamped into the row: loads need no predicate
      px[t] = x + c0 + coff[t];
    }
    double s[R];
    if (NCH == 4) {
#pragma unroll
      for (int r = 0; r < R; r++) {
        // the eight loads of a row in ONE asm block: issued back to back, all in flight before the first product waits
        double a0, a1, a2, a3, b0, b1, b2, b3;
        asm volatile(
            "ld.global.cs.f64 %0, [%8];\n\t"
            "ld.global.cs.f64 %1, [%9];\n\t"
            "ld.global.cs.f64 %2, [%10];\n\t"
            "ld.global.cs.f64 %3, [%11];\n\t"
            "ld.global.nc.f64 %4, [%12];\n\t"
            "ld.global.nc.f64 %5, [%13];\n\t"
            "ld.global.nc.f64 %6, [%14];\n\t"
            "ld.global.nc.f64 %7, [%15];"
            : "=d"(a0), "=d"(a1), "=d"(a2), "=d"(a3), "=d"(b0), "=d"(b1), "=d"(b2), "=d"(b3)
            : "l"(pv[0] + r * w), "l"(pv[1 % NCH] + r * w), "l"(pv[2 % NCH] + r * w), "l"(pv[3 % NCH] + r * w), "l"(px[0] + r), "l"(px[1 % NCH] + r), "l"(px[2 % NCH] + r),
              "l"(px[3 % NCH] + r));
        double t = (ok[0] ? a0 : 0.) * b0;
        t = fma(ok[1 % NCH] ? a1 : 0., b1, t);
        t = fma(ok[2 % NCH] ? a2 : 0., b2, t);
        s[r] = fma(ok[3 % NCH] ? a3 : 0., b3, t);
      }
    } else {
      // short rows (degree 1: 27 entries): the loads of all eight rows first, then the products
      double a[R][NCH], b[R][NCH];
#pragma unroll
      for (int r = 0; r < R; r++)
#pragma unroll
        for (int t = 0; t < NCH; t++) {
          a[r][t] = __ldcs(pv[t] + r * w);
          b[r][t] = __ldg(px[t] + r);
        }
#pragma unroll
      for (int r = 0; r < R; r++) {
        double acc = 0.;
#pragma unroll
        for (int t = 0; t < NCH; t++) acc = fma(ok[t] ? a[r][t] : 0., b[r][t], acc);
        s[r] = acc;
      }
    }
    // transposing butterfly: after the xor-16/8/4 steps a lane holds ONE row (row = lane >> 2), then two plain steps
    double t4[4], t2[2], v;
    {
      const bool hi = lane & 16;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const double send = hi ? s[j] : s[j + 4], keep = hi ? s[j + 4] : s[j];
        t4[j] = keep + __shfl_xor_sync(FULL, send, 16);
      }
    }
    {
      const bool hi = lane & 8;
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const double send = hi ? t4[j] : t4[j + 2], keep = hi ? t4[j + 2] : t4[j];
        t2[j] = keep + __shfl_xor_sync(FULL, send, 8);
      }
    }
    {
      const bool hi = lane & 4;
      const double send = hi ? t2[0] : t2[1], keep = hi ? t2[1] : t2[0];
      v = keep + __shfl_xor_sync(FULL, send, 4);
    }
    v += __shfl_xor_sync(FULL, v, 2);
    v += __shfl_xor_sync(FULL, v, 1);
    if ((lane & 3) == 0) {
      const int row = I0 + (lane >> 2);
      const bool fixed = mask && mask[row];
      y[row] = fixed ? 0. : v;
      if (dot && !fixed) local = fma(x[row], v, local);
    }
  }
  if (dot) block_accumulate(local, dot);
}

// Fast variant for rows of at most 32 NCH entries (scalar spaces up to degree 2 with NCH = 4, three components / degree 3
// with NCH = 12).  The instruction count per entry is what bounds the plain kernel (ncu: issue slots 75 % busy at 24 % of
// the DRAM bandwidth), so here (a) a warp walks rows I, I+8, I+16, ... and advances the dof multi-index incrementally (no
// 64-bit divisions), and (b) the column OFFSETS of a lane's entries relative to the first coupled dof depend only on the
// box widths (wid0, wid1, wid2), which are the same for all interior rows: they are kept in registers and recomputed only
// when the widths change.  Per entry: one streamed load, one add, one gather, one DFMA.
template <int DIM, int NCH, bool ALL>
__global__ void __launch_bounds__(256, 3) k_spmv_fast(const BasisView B, const int nbasis, const int rows_per_warp, const double* __restrict__ values,
                                                    const double* __restrict__ x, double* __restrict__ y, const unsigned char* __restrict__ mask, double* dot) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nc = B.ncomp, nd1 = DIM > 1 ? B.ndofs[1] : 1, nd2 = DIM > 2 ? B.ndofs[2] : 1;
  const int rb = 8 * rows_per_warp;
  double local = 0.;
  for (long long row0 = (long long)blockIdx.x * rb; row0 < nbasis; row0 += (long long)gridDim.x * rb) {
    int I = (int)row0 + warp;
    if (I >= nbasis) continue;
    int i[3] = {0, 0, 0};
    {
      unsigned r = (unsigned)I;
      if (DIM > 2) { i[2] = r % (unsigned)nd2; r /= (unsigned)nd2; }
      if (DIM > 1) { i[1] = r % (unsigned)nd1; r /= (unsigned)nd1; }
      i[0] = r;
      if (DIM == 1) i[0] = I;
    }
    int sig = -1, coff[NCH];
#pragma unroll
    for (int t = 0; t < NCH; t++) coff[t] = 0;
    for (int j = 0; j < rows_per_warp && I < nbasis; j++, I += 8) {
      int lo[3] = {0, 0, 0}, wid[3] = {1, 1, 1};
#pragma unroll
      for (int d = 0; d < DIM; d++) {
        lo[d] = B.lo[d][i[d]];
        wid[d] = B.wid[d][i[d]];
      }
      const int w = wid[0] * wid[1] * wid[2], rowlen = w * nc;
      const int nsig = wid[0] | wid[1] << 8 | wid[2] << 16;
      if (nsig != sig) {
        sig = nsig;
#pragma unroll
        for (int t = 0; t < NCH; t++) {
          const int k = t * 32 + lane;
          const int pos = k / nc, e = k - pos * nc;
          const int j2 = pos % wid[2], q = pos / wid[2], j1 = q % wid[1], j0 = q / wid[1];
          coff[t] = ((j0 * nd1 + j1) * nd2 + j2) * nc + e;
        }
      }
      const long long base = row_start_basis<DIM>(B, i) * nc * nc;
      const long long c0 = (((long long)lo[0] * nd1 + lo[1]) * nd2 + lo[2]) * nc;
      const double* xr = x + c0;
      // The ncomp dof rows (I, c) of a basis function couple with the SAME columns: x is gathered once per group of four
      // chunks and used for every component.  Addresses are clamped into the row instead of predicating the loads, and the
      // loads of a group sit in asm blocks issued back to back, so they are all in flight before the first product waits
      // (ptxas otherwise serialises load pair -> DFMA -> load pair: one HBM round trip per chunk instead of one per group).
      double s[3] = {0., 0., 0.};
      if (ALL && rowlen > 0) {
        // scalar space, long rows (degree 3): the loads of ALL chunks of the row are issued before the first product, one
        // HBM round trip per row instead of one per group of four chunks
        double a[NCH], b[NCH];
        bool ok[NCH];
#pragma unroll
        for (int t0 = 0; t0 < NCH; t0 += 4) {
          const double* pa[4];
          const double* pb[4];
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const int k = (t0 + u) * 32 + lane;
            ok[t0 + u] = k < rowlen;
            pa[u] = values + base + (ok[t0 + u] ? k : rowlen - 1);
            pb[u] = xr + (ok[t0 + u] ? coff[t0 + u] : 0);
          }
          asm volatile(
              "ld.global.cs.f64 %0, [%8];\n\t"
              "ld.global.cs.f64 %1, [%9];\n\t"
              "ld.global.cs.f64 %2, [%10];\n\t"
              "ld.global.cs.f64 %3, [%11];\n\t"
              "ld.global.nc.f64 %4, [%12];\n\t"
              "ld.global.nc.f64 %5, [%13];\n\t"
              "ld.global.nc.f64 %6, [%14];\n\t"
              "ld.global.nc.f64 %7, [%15];"
              : "=d"(a[t0]), "=d"(a[t0 + 1]), "=d"(a[t0 + 2]), "=d"(a[t0 + 3]), "=d"(b[t0]), "=d"(b[t0 + 1]), "=d"(b[t0 + 2]), "=d"(b[t0 + 3])
              : "l"(pa[0]), "l"(pa[1]), "l"(pa[2]), "l"(pa[3]), "l"(pb[0]), "l"(pb[1]), "l"(pb[2]), "l"(pb[3]));
        }
        double s0 = 0., s1 = 0.;
#pragma unroll
        for (int t = 0; t < NCH; t += 2) {
          s0 = fma(ok[t] ? a[t] : 0., b[t], s0);
          s1 = fma(ok[t + 1] ? a[t + 1] : 0., b[t + 1], s1);
        }
        s[0] = s0 + s1;
      } else if (rowlen > 0) {
#pragma unroll
        for (int t0 = 0; t0 < NCH; t0 += 4) {
          const double* pb[4];
          int ka[4];
          bool ok[4];
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const int k = (t0 + u) * 32 + lane;
            ok[u] = k < rowlen;
            ka[u] = ok[u] ? k : rowlen - 1;
            pb[u] = xr + (ok[u] ? coff[t0 + u] : 0);
          }
          double a[3][4], b0, b1, b2, b3;
          const double* v0 = values + base;
          asm volatile(
              "ld.global.cs.f64 %0, [%8];\n\t"
              "ld.global.cs.f64 %1, [%9];\n\t"
              "ld.global.cs.f64 %2, [%10];\n\t"
              "ld.global.cs.f64 %3, [%11];\n\t"
              "ld.global.nc.f64 %4, [%12];\n\t"
              "ld.global.nc.f64 %5, [%13];\n\t"
              "ld.global.nc.f64 %6, [%14];\n\t"
              "ld.global.nc.f64 %7, [%15];"
              : "=d"(a[0][0]), "=d"(a[0][1]), "=d"(a[0][2]), "=d"(a[0][3]), "=d"(b0), "=d"(b1), "=d"(b2), "=d"(b3)
              : "l"(v0 + ka[0]), "l"(v0 + ka[1]), "l"(v0 + ka[2]), "l"(v0 + ka[3]), "l"(pb[0]), "l"(pb[1]), "l"(pb[2]), "l"(pb[3]));
#pragma unroll
          for (int c = 1; c < 3; c++) {
            if (c < nc) {
              const double* vc = values + base + (long long)c * rowlen;
              asm volatile(
                  "ld.global.cs.f64 %0, [%4];\n\t"
                  "ld.global.cs.f64 %1, [%5];\n\t"
                  "ld.global.cs.f64 %2, [%6];\n\t"
                  "ld.global.cs.f64 %3, [%7];"
                  : "=d"(a[c][0]), "=d"(a[c][1]), "=d"(a[c][2]), "=d"(a[c][3])
                  : "l"(vc + ka[0]), "l"(vc + ka[1]), "l"(vc + ka[2]), "l"(vc + ka[3]));
            }
          }
          const double bb[4] = {ok[0] ? b0 : 0., ok[1] ? b1 : 0., ok[2] ? b2 : 0., ok[3] ? b3 : 0.};
#pragma unroll
          for (int c = 0; c < 3; c++) {
            if (c < nc) {
#pragma unroll
              for (int u = 0; u < 4; u++) s[c] = fma(a[c][u], bb[u], s[c]);
            }
          }
        }
      }
#pragma unroll
      for (int c = 0; c < 3; c++) {
        if (c < nc) {
          const long long row = (long long)I * nc + c;
          const double sc = warp_sum(s[c]);
          if (lane == 0) {
            const bool fixed = mask && mask[row];
            y[row] = fixed ? 0. : sc;
            if (dot && !fixed) local = fma(x[row], sc, local);
          }
        }
      }
      // next row of this warp: I + 8
      if (DIM == 1) i[0] += 8;
      else {
        i[DIM - 1] += 8;
        if (DIM == 3) {
          while (i[2] >= nd2) { i[2] -= nd2; i[1]++; }
          while (i[1] >= nd1) { i[1] -= nd1; i[0]++; }
        } else {
          while (i[1] >= nd1) { i[1] -= nd1; i[0]++; }
        }
      }
    }
  }
  if (dot) block_accumulate(local, dot);
}

// the same on the materialised basis-level pattern of an element set
__global__ void __launch_bounds__(256) k_spmv_general(const long long* __restrict__ rowptr_b, const int* __restrict__ colidx_b, const long long nbasis, const int nc,
                                                       const double* __restrict__ values, const double* __restrict__ x, double* __restrict__ y,
                                                       const unsigned char* __restrict__ mask, double* dot) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  double local = 0.;
  for (long long I = warp0; I < nbasis; I += nwarps) {
    const long long r0 = rowptr_b[I];
    const int len = (int)(rowptr_b[I + 1] - r0), rowlen = len * nc;
    for (int c = 0; c < nc; c++) {
      const long long row = I * nc + c;
      if (mask && mask[row]) {
        if (lane == 0) y[row] = 0.;
        continue;
      }
      const double* v = values + (r0 * nc + (long long)c * len) * nc;
      double s = 0.;
      for (int k = lane; k < rowlen; k += 32) {
        const int pos = k / nc, e = k - pos * nc;
        s = fma(ld_stream(v + k), x[(long long)colidx_b[r0 + pos] * nc + e], s);
      }
      s = warp_sum(s);
      if (lane == 0) {
        y[row] = s;
        if (dot) local = fma(x[row], s, local);
      }
    }
  }
  if (dot) block_accumulate(local, dot);
}

template <int DIM>
__global__ void __launch_bounds__(256) k_diagonal(const BasisView B, const long long nbasis, const double* __restrict__ values, double* __restrict__ diag) {
  const int nc = B.ncomp;
  for (long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x; row < nbasis * nc; row += (long long)gridDim.x * blockDim.x) {
    const long long I = row / nc;
    const int c = (int)(row - I * nc);
    int idx[3];
    const RowBox r = row_box<DIM>(B, I, idx);
    const int pos = ((idx[0] - r.lo[0]) * r.wid[1] + (idx[1] - r.lo[1])) * r.wid[2] + (idx[2] - r.lo[2]);
    diag[row] = r.w ? values[r.base + (long long)c * r.w * nc + (long long)pos * nc + c] : 0.;
  }
}

__global__ void __launch_bounds__(256) k_diagonal_general(const long long* __restrict__ rowptr_b, const int* __restrict__ colidx_b, const long long nbasis, const int nc,
                                                           const double* __restrict__ values, double* __restrict__ diag) {
  for (long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x; row < nbasis * nc; row += (long long)gridDim.x * blockDim.x) {
    const long long I = row / nc;
    const int c = (int)(row - I * nc);
    const long long r0 = rowptr_b[I];
    const int len = (int)(rowptr_b[I + 1] - r0);
    int lo = 0, hi = len;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (colidx_b[r0 + mid] < I) lo = mid + 1;
      else hi = mid;
    }
    diag[row] = (lo < len && colidx_b[r0 + lo] == I) ? values[(r0 * nc + (long long)c * len) * nc + (long long)lo * nc + c] : 0.;
  }
}

// ---- conjugate gradients: device scalars, no host round trip inside an iteration ----
// S[0] rz_old  S[1] rz_new  S[2] p.q  S[3] r.r (current)  S[4] r.r (last completed iteration)
struct CgVec {
  double *x, *r, *p, *q, *dinv;
  const unsigned char* mask;
  long long n;
};

// r = (b - q) on free rows, 0 on constrained; dinv = 1 / diag (1 where the diagonal vanishes); p = dinv r; S[0] = r.p; S[4] = r.r
__global__ void __launch_bounds__(256) k_cg_init(const CgVec V, const double* __restrict__ b, const double* diag, double* S) {
  double rz = 0., rr = 0.;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < V.n; i += (long long)gridDim.x * blockDim.x) {
    const bool fixed = V.mask && V.mask[i];
    const double r = fixed ? 0. : (b ? b[i] : 0.) - V.q[i];
    const double d = diag[i], di = d != 0. ? 1. / d : 1.;
    V.r[i] = r;
    V.dinv[i] = di;
    V.p[i] = di * r;
    rz = fma(r * di, r, rz);
    rr = fma(r, r, rr);
  }
  block_accumulate(rz, S + 0);
  __syncthreads();
  block_accumulate(rr, S + 4);
}

// alpha = rz_old / p.q;  x += alpha p;  r -= alpha q;  rz_new += r.dinv.r;  rr += r.r
__global__ void __launch_bounds__(256) k_cg_update1(const CgVec V, double* S) {
  const double pq = S[2], alpha = pq != 0. ? S[0] / pq : 0.;
  double rz = 0., rr = 0.;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < V.n; i += (long long)gridDim.x * blockDim.x) {
    V.x[i] = fma(alpha, V.p[i], V.x[i]);
    const double r = fma(-alpha, V.q[i], V.r[i]);
    V.r[i] = r;
    rz = fma(r * V.dinv[i], r, rz);
    rr = fma(r, r, rr);
  }
  block_accumulate(rz, S + 1);
  __syncthreads();
  block_accumulate(rr, S + 3);
}

// beta = rz_new / rz_old;  p = dinv r + beta p
__global__ void __launch_bounds__(256) k_cg_update2(const CgVec V, const double* S) {
  const double beta = S[0] != 0. ? S[1] / S[0] : 0.;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < V.n; i += (long long)gridDim.x * blockDim.x)
    V.p[i] = fma(beta, V.p[i], V.dinv[i] * V.r[i]);
}

__global__ void k_cg_rotate(double* S) {
  S[0] = S[1];
  S[1] = 0.;
  S[2] = 0.;
  S[4] = S[3];
  S[3] = 0.;
}

int grid_for(b2_ctx* ctx, long long work_items, int per_block) {
  const long long want = (work_items + per_block - 1) / per_block;
  return (int)std::min<long long>(std::max<long long>(want, 1), (long long)ctx->sm_count * 16);
}

int spmv_launch(b2_ctx* ctx, const b2_pattern* p, const double* values, const double* x, double* y, const unsigned char* mask, double* dot) {
  const b2_basis* basis = p->basis;
  const int nc = pattern_ncomp(p);
  const long long nbasis = p->nrows / nc;
  const int blocks = grid_for(ctx, nbasis, 8);  // 8 warps per block, one basis row per warp and step
  {
    KernelTimer timer(ctx);
    if (pattern_general(p)) {
      k_spmv_general<<<blocks, 256, 0, ctx->stream>>>(p->d_rowptr_b, p->d_colidx_b, nbasis, nc, values, x, y, mask, dot);
    } else {
      const BasisView B = basis->view();
      int maxrow = nc;  // longest row: product of the largest box widths
      for (int d = 0; d < basis->ndims; d++) maxrow *= *std::max_element(basis->wid[d].begin(), basis->wid[d].end());
      const bool plain = ctx->opts.count("spmv_plain") && ctx->opts["spmv_plain"];
      if (nc == 1 && maxrow <= 128 && !plain) {
        const int rblocks = grid_for(ctx, nbasis, 64);
        const bool one = maxrow <= 32;  // one 32-entry chunk per row (degree 1)
        switch (basis->ndims) {
          case 1: k_spmv_run<1, 1><<<rblocks, 256, 0, ctx->stream>>>(B, (int)nbasis, values, x, y, mask, dot); break;
          case 2:
            if (one) k_spmv_run<2, 1><<<rblocks, 256, 0, ctx->stream>>>(B, (int)nbasis, values, x, y, mask, dot);
            else k_spmv_run<2, 4><<<rblocks, 256, 0, ctx->stream>>>(B, (int)nbasis, values, x, y, mask, dot);
            break;
          default:
            if (one) k_spmv_run<3, 1><<<rblocks, 256, 0, ctx->stream>>>(B, (int)nbasis, values, x, y, mask, dot);
            else k_spmv_run<3, 4><<<rblocks, 256, 0, ctx->stream>>>(B, (int)nbasis, values, x, y, mask, dot);
        }
      } else if (maxrow <= 384 && !plain) {
        const int rpw = 16, fblocks = grid_for(ctx, nbasis, 8 * rpw);
        const bool small = maxrow <= 128;
        switch (basis->ndims) {
          case 1: k_spmv_fast<1, 4, false><<<fblocks, 256, 0, ctx->stream>>>(B, (int)nbasis, rpw, values, x, y, mask, dot); break;
          case 2:
            if (small) k_spmv_fast<2, 4, false><<<fblocks, 256, 0, ctx->stream>>>(B, (int)nbasis, rpw, values, x, y, mask, dot);
            else if (nc == 1) k_spmv_fast<2, 12, true><<<fblocks, 256, 0, ctx->stream>>>(B, (int)nbasis, rpw, values, x, y, mask, dot);
            else k_spmv_fast<2, 12, false><<<fblocks, 256, 0, ctx->stream>>>(B, (int)nbasis, rpw, values, x, y, mask, dot);
            break;
          default:
            if (small) k_spmv_fast<3, 4, false><<<fblocks, 256, 0, ctx->stream>>>(B, (int)nbasis, rpw, values, x, y, mask, dot);
            else if (nc == 1) k_spmv_fast<3, 12, true><<<fblocks, 256, 0, ctx->stream>>>(B, (int)nbasis, rpw, values, x, y, mask, dot);
            else k_spmv_fast<3, 12, false><<<fblocks, 256, 0, ctx->stream>>>(B, (int)nbasis, rpw, values, x, y, mask, dot);
        }
      } else {
        switch (basis->ndims) {
          case 1: k_spmv<1><<<blocks, 256, 0, ctx->stream>>>(B, nbasis, values, x, y, mask, dot); break;
          case 2: k_spmv<2><<<blocks, 256, 0, ctx->stream>>>(B, nbasis, values, x, y, mask, dot); break;
          default: k_spmv<3><<<blocks, 256, 0, ctx->stream>>>(B, nbasis, values, x, y, mask, dot); break;
        }
      }
    }
  }
  ctx->launches++;
  B2_CUDA(ctx, cudaGetLastError());
  return B2_OK;
}

int diagonal_launch(b2_ctx* ctx, const b2_pattern* p, const double* values, double* diag) {
  const b2_basis* basis = p->basis;
  const int nc = pattern_ncomp(p);
  const long long nbasis = p->nrows / nc;
  const int blocks = grid_for(ctx, p->nrows, 256);
  if (pattern_general(p)) {
    k_diagonal_general<<<blocks, 256, 0, ctx->stream>>>(p->d_rowptr_b, p->d_colidx_b, nbasis, nc, values, diag);
  } else {
    const BasisView B = basis->view();
    switch (basis->ndims) {
      case 1: k_diagonal<1><<<blocks, 256, 0, ctx->stream>>>(B, nbasis, values, diag); break;
      case 2: k_diagonal<2><<<blocks, 256, 0, ctx->stream>>>(B, nbasis, values, diag); break;
      default: k_diagonal<3><<<blocks, 256, 0, ctx->stream>>>(B, nbasis, values, diag); break;
    }
  }
  ctx->launches++;
  B2_CUDA(ctx, cudaGetLastError());
  return B2_OK;
}

}  // namespace

extern "C" int b2_spmv_device(b2_ctx* ctx, const b2_pattern* pattern, const double* values_dev, const double* x_dev, double* y_dev) {
  if (!ctx || !pattern || !values_dev || !x_dev || !y_dev) return b2_fail(ctx, B2_EINVAL, "null argument");
  if (x_dev == y_dev) return b2_fail(ctx, B2_EINVAL, "x and y must not alias");
  B2_CUDA(ctx, cudaSetDevice(ctx->device));
  return spmv_launch(ctx, pattern, values_dev, x_dev, y_dev, nullptr, nullptr);
}

extern "C" int b2_diagonal_device(b2_ctx* ctx, const b2_pattern* pattern, const double* values_dev, double* diag_dev) {
  if (!ctx || !pattern || !values_dev || !diag_dev) return b2_fail(ctx, B2_EINVAL, "null argument");
  B2_CUDA(ctx, cudaSetDevice(ctx->device));
  return diagonal_launch(ctx, pattern, values_dev, diag_dev);
}

extern "C" int b2_cg_device(b2_ctx* ctx, const b2_pattern* pattern, const double* values_dev, const double* rhs_dev, double* x_dev,
                            const unsigned char* constrained_dev, double atol, double rtol, int maxiter, int* iterations, double* resnorm) {
  if (!ctx || !pattern || !values_dev || !x_dev) return b2_fail(ctx, B2_EINVAL, "null argument");
  if (atol < 0. || rtol < 0. || maxiter < 0) return b2_fail(ctx, B2_EINVAL, "negative tolerance or iteration count");
  B2_CUDA(ctx, cudaSetDevice(ctx->device));
  const long long n = pattern->nrows;
  if (iterations) *iterations = 0;
  if (resnorm) *resnorm = 0.;
  if (n == 0) return B2_OK;
  double* work = nullptr;
  B2_CUDA(ctx, cudaMalloc((void**)&work, sizeof(double) * (size_t)(4 * n + 8)));
  CgVec V;
  V.x = x_dev;
  V.r = work;
  V.p = work + n;
  V.q = work + 2 * n;
  V.dinv = work + 3 * n;
  V.mask = constrained_dev;
  V.n = n;
  double* S = work + 4 * n;
  const int vblocks = grid_for(ctx, n, 256 * 4);
  int rc = B2_OK, it = 0;
  double rr = 0., target = 0.;
  bool explicit_tol = atol > 0. || rtol > 0.;
  auto fail = [&](cudaError_t e, const char* what) { rc = b2_cuda_fail(ctx, e, what); };
  do {
    cudaError_t e = cudaMemsetAsync(S, 0, sizeof(double) * 8, ctx->stream);
    if (e != cudaSuccess) { fail(e, "cudaMemsetAsync"); break; }
    // q = A x0 with ALL columns (constrained values included), then r = b - q on the free rows
    rc = spmv_launch(ctx, pattern, values_dev, x_dev, V.q, nullptr, nullptr);
    if (rc != B2_OK) break;
    rc = diagonal_launch(ctx, pattern, values_dev, V.dinv);  // dinv holds the diagonal until k_cg_init inverts it in place
    if (rc != B2_OK) break;
    k_cg_init<<<vblocks, 256, 0, ctx->stream>>>(V, rhs_dev, V.dinv, S);
    ctx->launches++;
    e = cudaMemcpyAsync(&rr, S + 4, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { fail(e, "cg init"); break; }
    const double r0 = std::sqrt(rr);
    // matrix/_base.py:199-206: tolerance max(atol, rtol |b_reduced|); both zero = "machine precision"
    target = explicit_tol ? std::max(atol, rtol * r0) : 1e-13 * r0;
    if (maxiter == 0) maxiter = (int)std::min<long long>(10 * n + 100, 100000);
    const int check = 10;
    while (std::sqrt(rr) > target && it < maxiter) {
      for (int k = 0; k < check && it < maxiter; k++, it++) {
        rc = spmv_launch(ctx, pattern, values_dev, V.p, V.q, constrained_dev, S + 2);
        if (rc != B2_OK) break;
        k_cg_update1<<<vblocks, 256, 0, ctx->stream>>>(V, S);
        k_cg_update2<<<vblocks, 256, 0, ctx->stream>>>(V, S);
        k_cg_rotate<<<1, 1, 0, ctx->stream>>>(S);
        ctx->launches += 3;
      }
      if (rc != B2_OK) break;
      e = cudaMemcpyAsync(&rr, S + 4, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
      if (e != cudaSuccess) { fail(e, "cg iteration"); break; }
      if (!(rr == rr)) { rc = b2_fail(ctx, B2_ENOTCONVERGED, "conjugate gradients broke down (matrix not positive definite on the free dofs?)"); break; }
    }
  } while (false);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(work);
  if (iterations) *iterations = it;
  if (resnorm) *resnorm = std::sqrt(rr);
  if (rc != B2_OK) return rc;
  if (explicit_tol && std::sqrt(rr) > target) return b2_fail(ctx, B2_ENOTCONVERGED, "tolerance not reached");
  // atol = rtol = 0 asks for machine precision (matrix/_base.py:126-129): running out of iterations far above it is not a solution
  if (!explicit_tol && it >= maxiter && std::sqrt(rr) > 1e3 * target) return b2_fail(ctx, B2_ENOTCONVERGED, "iteration limit reached before machine precision");
  return B2_OK;
}
