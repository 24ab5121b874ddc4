// Element-set integration kernel: the element loop of the reference for topologies that live on a structured grid
// but are NOT the plain tensor case -- trimmed / subset topologies with per-element point sets (finite cell method),
// pruned dof numberings, rational (NURBS) functions and spline geometries.
//
// What it replaces in the reference (src/nutils): the generated loop body (evaluable.py:6773-6787) with
//   - the per-element point set coords/weights looked up through PointsSequence.get (pointsseq.py:324-332; ragged
//     ConcatPoints of cut cells, points.py:257-337),
//   - Polyval / PolyGrad of the element's coefficient rows at ARBITRARY local points (evaluable.py:4328-4374,
//     4584-4634) -- here Horner on the 1-D factors,
//   - PrunedBasis.f_dofs_coeffs (function.py:3130-3133): parent dofs renumbered,
//   - rational functions c_i B_i / W and their gradients (quotient rule the reference derives symbolically,
//     examples/platewithhole.py:66-83),
//   - J, J^-1, |det J| of a multilinear OR (rational) spline geometry (function.py:1207-1231, 1266-1295),
//   - the integrand einsums, the weighted point sum (sample.py:951-956), COO append + accumulate
//     (evaluable.py:5383-5501, numeric.py:434-460) as fp64 atomicAdd into the CSR slot found by bisection in the
//     row's sorted column list, and numpy.add.at for load vectors (evaluable.py:3582-3620).
//
// Mapping: one CTA per selected element (grid-stride), quadrature points in chunks through shared memory:
//   P0 points of the chunk -> P1 1-D values/derivatives (Horner) -> P2 geometry per point -> P3 (N, grad N) per
//   (point, function) -> P4 block entries in registers; scatter at the end of the element.
//
// P4 on the tensor cores (2-D and 3-D; the default).  For one coefficient block D_m[c][.][e][.] the element block is a
// dense GEMM  C[a][b] = sum_q sum_y ( w_q sum_x B[q][a][x] D[x][y] ) B[q][b][y]  with M = N = n_e and K = nq (d+1):
// FP64 mma.sync.m8n8k4 (SASS DMMA.8x8x4), one warp per row of 8x8 tiles, one k-block per quadrature point (k = y; mass-like
// blocks batch four points per k-block instead).  A DMMA carries 256 FMAs per issue slot against 32 of a DFMA and
// tolerates the low occupancy of a big-shared-memory kernel (profiles/microbench/ubench3.cu: 37 TFLOP/s with 4 warps per
// SM and two accumulator tiles per warp; DFMA 34 TFLOP/s only with 16 independent chains per thread) -- this is the
// "genuinely dense per-element GEMM" case of the high-order / many-point elements (p = 4: 125 x 125 x 500; cut cells:
// thousands of points per element).  Context option "elemset_mma" = 0 selects the scalar FMA loop (kept for A/B parity).

#include <algorithm>
#include <cstring>

#include "common.cuh"

namespace {

constexpr int EPT = 8;  // block entries per thread and pass

// evaluation mode (b2_evaluate_elemset_device): instead of integrating, write per point the physical coordinates, the
// weight w |det J| (surface measure on faces), and the values / physical gradients of discrete fields sum_i coef[i] N_i
struct EvalView {
  int nfields;            // 0: integration mode
  const double* coef;     // [nfields][ndofs]
  double* x;              // [npoints][DIM] or null
  double* wdet;           // [npoints] or null
  double* values;         // [npoints][nfields ncomp] or null
  double* grads;          // [npoints][nfields ncomp][DIM] or null
  long long ndofs;
  int enabled;
};

struct ESParams {
  EvalView ev;
  BasisView B;
  QuadView Q;
  GeomView G;
  SplineGeomView SG;
  ElemSetView E;
  FormView F;
  long long sel_begin, sel_end;
  unsigned long long* queue;  // next element to hand out (zeroed before the launch); null: static round robin
  int qchunk;  // points per chunk
  int pm1;     // max(p)+1 of the solution basis
  int pgm1;    // max(p)+1 of the geometry basis (spline geometry)
  int nbg;     // functions per element of the geometry basis
  int ne;      // nb * ncomp
};

template <int DIM>
__device__ __forceinline__ void inv_det(const double* J, double* Ji, double& det) {
  if (DIM == 1) {
    det = J[0];
    Ji[0] = 1. / J[0];
  } else if (DIM == 2) {
    det = J[0] * J[3] - J[1] * J[2];
    const double r = 1. / det;
    Ji[0] = J[3] * r; Ji[1] = -J[1] * r;
    Ji[2] = -J[2] * r; Ji[3] = J[0] * r;
  } else {
    const double c00 = J[4] * J[8] - J[5] * J[7];
    const double c01 = J[5] * J[6] - J[3] * J[8];
    const double c02 = J[3] * J[7] - J[4] * J[6];
    det = J[0] * c00 + J[1] * c01 + J[2] * c02;
    const double r = 1. / det;
    Ji[0] = c00 * r; Ji[1] = (J[2] * J[7] - J[1] * J[8]) * r; Ji[2] = (J[1] * J[5] - J[2] * J[4]) * r;
    Ji[3] = c01 * r; Ji[4] = (J[0] * J[8] - J[2] * J[6]) * r; Ji[5] = (J[2] * J[3] - J[0] * J[5]) * r;
    Ji[6] = c02 * r; Ji[7] = (J[1] * J[6] - J[0] * J[7]) * r; Ji[8] = (J[0] * J[4] - J[1] * J[3]) * r;
  }
}

// value and xi-gradient of local function `a` (C-order multi-index over per-dimension degrees p[]) at chunk point ql,
// from the 1-D table tab[ql][d][pm1][2]
// `mi` = packed multi-index of the function (a_0 | a_1 << 8 | a_2 << 16, tabulated once per CTA: no integer divisions here)
template <int DIM>
__device__ __forceinline__ void tensor_eval(const double* tab, int pm1, int mi, double& N, double* dxi) {
  const int ad[3] = {mi & 255, (mi >> 8) & 255, (mi >> 16) & 255};
  double val[DIM], der[DIM];
#pragma unroll
  for (int d = 0; d < DIM; d++) {
    val[d] = tab[(d * pm1 + ad[d]) * 2];
    der[d] = tab[(d * pm1 + ad[d]) * 2 + 1];
  }
  N = 1.;
#pragma unroll
  for (int d = 0; d < DIM; d++) N *= val[d];
#pragma unroll
  for (int k = 0; k < DIM; k++) {
    double g = 1.;
#pragma unroll
    for (int d = 0; d < DIM; d++) g *= d == k ? der[d] : val[d];
    dxi[k] = g;
  }
}

// D(8x8) += A(8x4) B(4x8) in FP64 on the tensor cores: lane holds A[lane/4][lane%4], B[lane%4][lane/4], C[lane/4][2 (lane%4) + {0,1}]
__device__ __forceinline__ void dmma884(double& c0, double& c1, const double a, const double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// TCMAX = 0: block entries by scalar FMAs (any dimension).  TCMAX > 0: tensor-core path, up to TCMAX tile columns
// (n_e <= 8 TCMAX) and NJ coefficient blocks per pass over the points.
template <int DIM, int TCMAX, int NJ>
__global__ void __launch_bounds__(TCMAX == 4 ? 128 : 512, TCMAX == 4 ? 4 : 1) k_assemble_elemset(const ESParams P) {
  constexpr int NA = DIM + 1;
  constexpr int NV = 1 << DIM;
  constexpr int JS = DIM * DIM + 1;
  const BasisView& B = P.B;
  const ElemSetView& E = P.E;
  const int tid = threadIdx.x, T = blockDim.x;
  const int nb = B.nb, nc = B.ncomp, ne = P.ne, ne2 = ne * ne;
  const int qc = P.qchunk, pm1 = P.pm1, pgm1 = P.pgm1, nbg = P.nbg;
  const bool spline = P.SG.enabled != 0;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sC = reinterpret_cast<double*>(smem_raw);     // [DIM][pm1][pm1]   coefficient rows of the element
  double* sCg = sC + DIM * pm1 * pm1;                   // [DIM][pgm1][pgm1] same for the geometry basis
  double* sX = sCg + DIM * pgm1 * pgm1;                 // nodal: [DIM][NV]; spline: [DIM+1][nbg] (w X, w)
  double* sS = sX + (spline ? (DIM + 1) * nbg : DIM * NV);  // [nb] scale c_a
  double* sPt = sS + nb;                                // [qc][NA]  xi, weight
  double* sA = sPt + qc * NA;                           // [qc][DIM][pm1][2]
  double* sAg = sA + qc * DIM * pm1 * 2;                // [qc][DIM][pgm1][2]
  double* sJ = sAg + qc * DIM * pgm1 * 2;               // [qc][JS]  J^-1 (k,i) then w|det|
  double* sW = sJ + qc * JS;                            // [qc][NA]  W, dW/dxi (rational functions)
  double* sK = sW + qc * NA;                            // [qc][2 B2_MAX_FORMS] pointwise coefficient of every form (1 if none)
  double* sB = sK + qc * 2 * B2_MAX_FORMS;              // [qc][nb][NA]
  sB += (reinterpret_cast<unsigned long long>(sB) >> 3) & 1;  // 16-byte aligned (the launch reserves the slack): 3-D entries move as 128-bit vectors
  double* sV = sB + qc * nb * NA;                       // [nvec][ne]
  long long* sRow = reinterpret_cast<long long*>(sV + B2_MAX_FORMS * ne);  // [nb] first slot of the basis row
  int* sLen = reinterpret_cast<int*>(sRow + nb);        // [nb] columns of the basis row
  int* sDof = sLen + nb;                                // [nb] new basis index, -1: dropped
  double* sD = reinterpret_cast<double*>(sDof + nb);  // [NJ][16] coefficient blocks of the current pass (tensor-core path)
  int* sMi = reinterpret_cast<int*>(sD + NJ * 16);      // [nb] packed multi-index of the local functions
  int* sMg = sMi + nb;                                  // [nbg] same for the geometry basis
  constexpr bool MMA = TCMAX > 0;
  const int ntr = (nb + 7) / 8;                          // rows / columns of 8x8 tiles
  for (int a = tid; a < nb + (spline ? nbg : 0); a += T) {
    const bool geo = a >= nb;
    const int* pp = geo ? P.SG.GB.p : B.p;
    int r = geo ? a - nb : a, mi = 0;
    for (int d = DIM - 1; d >= 0; d--) {
      mi |= (r % (pp[d] + 1)) << (8 * d);
      r /= pp[d] + 1;
    }
    (geo ? sMg : sMi)[geo ? a - nb : a] = mi;
  }

  // Elements differ in cost by two orders of magnitude (27 points in an uncut cell, thousands in a cut one): they are handed out
  // one at a time from a global counter instead of round robin, so that no CTA is left with a tail of expensive cells
  __shared__ long long s_sel;
  const bool use_order = E.order && P.sel_begin == 0 && P.sel_end == E.nsel;  // (a partial range keeps the order of the set)
  for (long long sel = P.sel_begin + blockIdx.x;; sel += gridDim.x) {
    if (P.queue) {
      if (tid == 0) s_sel = P.sel_begin + (long long)atomicAdd(P.queue, 1ULL);
      __syncthreads();   // (the next write of s_sel lies behind the barriers of this element)
      sel = s_sel;
    }
    if (sel >= P.sel_end) break;
    if (P.queue && use_order) sel = E.order[sel];
    const long long elem = E.elem_ids ? E.elem_ids[sel] : sel;
    int ie[3] = {0, 0, 0};
    {
      long long r = elem;
      for (int d = DIM - 1; d >= 0; d--) {
        ie[d] = (int)(r % B.nel[d]);
        r /= B.nel[d];
      }
    }
    const int fdim = E.face_dim ? E.face_dim[sel] : -1;   // >= 0: the points lie on a face, surface measure
    const long long qbeg = E.qoff ? E.qoff[sel] : sel * E.nq_uniform;
    const int nqt = E.qoff ? (int)(E.qoff[sel + 1] - qbeg) : P.Q.nqt;
    __syncthreads();  // previous element fully scattered before shared memory is reused
    for (int t = tid; t < DIM * pm1 * pm1; t += T) {
      const int j = t % pm1, a = (t / pm1) % pm1, d = t / (pm1 * pm1);
      const int p = B.p[d];
      sC[t] = (a <= p && j <= p) ? E.coeffs[d][((long long)B.setidx[d][ie[d]] * (p + 1) + a) * (p + 1) + j] : 0.;
    }
    if (spline) {
      const BasisView& GB = P.SG.GB;
      for (int t = tid; t < DIM * pgm1 * pgm1; t += T) {
        const int j = t % pgm1, a = (t / pgm1) % pgm1, d = t / (pgm1 * pgm1);
        const int p = GB.p[d];
        sCg[t] = (a <= p && j <= p) ? P.SG.coeffs[d][((long long)GB.setidx[d][ie[d]] * (p + 1) + a) * (p + 1) + j] : 0.;
      }
      for (int a = tid; a < nbg; a += T) {
        long long I = 0;
        int r = a, ad[3] = {0, 0, 0};
        for (int d = DIM - 1; d >= 0; d--) {
          ad[d] = r % (GB.p[d] + 1);
          r /= GB.p[d] + 1;
        }
        for (int d = 0; d < DIM; d++) I = I * GB.ndofs[d] + GB.start[d][ie[d]] + ad[d];
        const double w = P.SG.wts ? P.SG.wts[I] : 1.;
        for (int i = 0; i < DIM; i++) sX[i * nbg + a] = w * P.SG.ctrl[i * P.SG.nbasis + I];
        sX[DIM * nbg + a] = w;
      }
    } else {
      for (int t = tid; t < DIM * NV; t += T) {
        const int v = t % NV, i = t / NV;
        long long off = 0;
        for (int d = 0; d < DIM; d++) off += (long long)(ie[d] + ((v >> (DIM - 1 - d)) & 1)) * P.G.stride[d];
        sX[t] = P.G.nodes[i * P.G.nnodes + off];
      }
    }
    for (int a = tid; a < nb; a += T) {
      long long I = 0;
      int r = a, ad[3] = {0, 0, 0};
      for (int d = DIM - 1; d >= 0; d--) {
        ad[d] = r % (B.p[d] + 1);
        r /= B.p[d] + 1;
      }
      for (int d = 0; d < DIM; d++) I = I * B.ndofs[d] + B.start[d][ie[d]] + ad[d];
      int In = E.renumber ? E.renumber[I] : (int)I;
      if (In < 0 || In >= E.nbasis_new) In = -1;
      sDof[a] = In;
      sS[a] = E.scale ? E.scale[I] : 1.;
      sRow[a] = (In >= 0 && E.rowptr_b) ? E.rowptr_b[In] : 0;
      sLen[a] = (In >= 0 && E.rowptr_b) ? (int)(E.rowptr_b[In + 1] - E.rowptr_b[In]) : 0;
    }
    for (int t = tid; t < P.F.nvec * ne; t += T) sV[t] = 0.;
    __syncthreads();

    // a pass = one sweep over the points: EPT block entries per thread (FMA path) or NJ coefficient blocks (tensor cores)
    const int pass_step = MMA ? NJ : T * EPT, pass_end = MMA ? P.F.njobs : ne2;
    for (int pass0 = 0; pass0 < pass_end || pass0 == 0; pass0 += pass_step) {
      double acc[MMA ? 1 : B2_MAX_FORMS][MMA ? 1 : EPT];
      double cacc[MMA ? NJ : 1][MMA ? TCMAX : 1][2];
      if constexpr (MMA) {
#pragma unroll
        for (int j = 0; j < NJ; j++)
#pragma unroll
          for (int tc = 0; tc < TCMAX; tc++) cacc[j][tc][0] = cacc[j][tc][1] = 0.;
        __syncthreads();  // the previous pass has read sD
        for (int t = tid; t < NJ * 16; t += T) sD[t] = pass0 + t / 16 < P.F.njobs ? P.F.jobD[(pass0 + t / 16) * 16 + t % 16] : 0.;
      } else {
#pragma unroll
        for (int m = 0; m < (MMA ? 1 : B2_MAX_FORMS); m++)
#pragma unroll
          for (int e = 0; e < (MMA ? 1 : EPT); e++) acc[m][e] = 0.;
      }

      for (int q0 = 0; q0 < nqt; q0 += qc) {
        const int nqc = min(qc, nqt - q0);
        // P0: local coordinates and weights of the chunk
        for (int ql = tid; ql < nqc; ql += T) {
          double* pt = sPt + ql * NA;
          if (E.qoff) {
            const long long q = qbeg + q0 + ql;
#pragma unroll
            for (int d = 0; d < DIM; d++) pt[d] = E.qcoords[q * DIM + d];
            pt[DIM] = E.qweights[q];
          } else {
            int r = q0 + ql;
            double w = 1.;
            for (int d = DIM - 1; d >= 0; d--) {
              const int k = r % P.Q.nq[d];
              r /= P.Q.nq[d];
              pt[d] = P.Q.x[d][k];
              w *= P.Q.w[d][k];
            }
            pt[DIM] = w;
          }
#pragma unroll
          for (int k = 0; k < 2 * B2_MAX_FORMS; k++) sK[ql * 2 * B2_MAX_FORMS + k] = E.coef[k] ? E.coef[k][qbeg + q0 + ql] : 1.;
        }
        __syncthreads();
        // P1: 1-D values and derivatives (Horner with derivative; rows are highest power first)
        for (int t = tid; t < nqc * DIM * pm1; t += T) {
          const int a = t % pm1, d = (t / pm1) % DIM, ql = t / (pm1 * DIM);
          const int p = B.p[d];
          double v = 0., g = 0.;
          if (a <= p) {
            const double x = sPt[ql * NA + d];
            const double* c = sC + (d * pm1 + a) * pm1;
            v = c[0];
            for (int j = 1; j <= p; j++) { g = g * x + v; v = v * x + c[j]; }
          }
          sA[t * 2] = v;
          sA[t * 2 + 1] = g;
        }
        if (spline) {
          for (int t = tid; t < nqc * DIM * pgm1; t += T) {
            const int a = t % pgm1, d = (t / pgm1) % DIM, ql = t / (pgm1 * DIM);
            const int p = P.SG.GB.p[d];
            double v = 0., g = 0.;
            if (a <= p) {
              const double x = sPt[ql * NA + d];
              const double* c = sCg + (d * pgm1 + a) * pgm1;
              v = c[0];
              for (int j = 1; j <= p; j++) { g = g * x + v; v = v * x + c[j]; }
            }
            sAg[t * 2] = v;
            sAg[t * 2 + 1] = g;
          }
        }
        __syncthreads();
        // P1b: solution-dependent coefficients: u_h at the points from the element's dofs of the field vector,
        // folded into the pointwise coefficient of the form -- what the reference evaluates per element as
        // Polyval(coeffs, points) . take(arg, dofs) inside its loop (function.py:2758-2762, 2598-2626)
        {
          bool anyfield = false;
#pragma unroll
          for (int k = 0; k < 2 * B2_MAX_FORMS; k++) anyfield |= E.field[k] != nullptr;
          if (anyfield) {
            for (int ql = tid; ql < nqc; ql += T) {
              double uh[2 * B2_MAX_FORMS];
#pragma unroll
              for (int k = 0; k < 2 * B2_MAX_FORMS; k++) uh[k] = 0.;
              for (int a = 0; a < nb; a++) {
                if (sDof[a] < 0) continue;
                double N, dxi[DIM];
                tensor_eval<DIM>(sA + ql * DIM * pm1 * 2, pm1, sMi[a], N, dxi);
#pragma unroll
                for (int k = 0; k < 2 * B2_MAX_FORMS; k++)
                  if (E.field[k]) uh[k] = fma(N, E.field[k][sDof[a]], uh[k]);
              }
#pragma unroll
              for (int k = 0; k < 2 * B2_MAX_FORMS; k++)
                if (E.field[k]) {
                  double c = E.field_scale[k];
                  for (int j = 0; j < E.field_power[k]; j++) c *= uh[k];
                  sK[ql * 2 * B2_MAX_FORMS + k] *= c;
                }
            }
            __syncthreads();
          }
        }
        // P2: geometry at the points of the chunk (and the weight function of rational bases)
        for (int ql = tid; ql < nqc; ql += T) {
          const double* pt = sPt + ql * NA;
          double J[DIM * DIM];
          double Wg = 1., dWg[DIM];
#pragma unroll
          for (int k = 0; k < DIM; k++) dWg[k] = 0.;
          if (spline) {
            double Xh[DIM], dXh[DIM * DIM];
#pragma unroll
            for (int i = 0; i < DIM; i++) Xh[i] = 0.;
#pragma unroll
            for (int i = 0; i < DIM * DIM; i++) dXh[i] = 0.;
            double W = 0.;
            for (int a = 0; a < nbg; a++) {
              double N, dxi[DIM];
              tensor_eval<DIM>(sAg + ql * DIM * pgm1 * 2, pgm1, sMg[a], N, dxi);
              const double w = sX[DIM * nbg + a];
              W = fma(N, w, W);
#pragma unroll
              for (int k = 0; k < DIM; k++) dWg[k] = fma(dxi[k], w, dWg[k]);
#pragma unroll
              for (int i = 0; i < DIM; i++) {
                const double xw = sX[i * nbg + a];
                Xh[i] = fma(N, xw, Xh[i]);
#pragma unroll
                for (int k = 0; k < DIM; k++) dXh[i * DIM + k] = fma(dxi[k], xw, dXh[i * DIM + k]);
              }
            }
            // x = Xh / W,  dx/dxi = (dXh - x dW) / W   (W = 1, dW = 0 for polynomial splines)
            const double rW = 1. / W;
#pragma unroll
            for (int i = 0; i < DIM; i++) {
              const double x = Xh[i] * rW;
#pragma unroll
              for (int k = 0; k < DIM; k++) J[i * DIM + k] = (dXh[i * DIM + k] - x * dWg[k]) * rW;
            }
            Wg = W;
          } else {
#pragma unroll
            for (int i = 0; i < DIM * DIM; i++) J[i] = 0.;
#pragma unroll
            for (int v = 0; v < NV; v++) {
#pragma unroll
              for (int k = 0; k < DIM; k++) {
                double f = 1.;
#pragma unroll
                for (int d = 0; d < DIM; d++) {
                  const bool bit = (v >> (DIM - 1 - d)) & 1;
                  f *= d == k ? (bit ? 1. : -1.) : (bit ? pt[d] : 1. - pt[d]);
                }
#pragma unroll
                for (int i = 0; i < DIM; i++) J[i * DIM + k] += sX[i * NV + v] * f;
              }
            }
          }
          double Ji[DIM * DIM], det;
          inv_det<DIM>(J, Ji, det);
#pragma unroll
          for (int i = 0; i < DIM * DIM; i++) sJ[ql * JS + i] = Ji[i];
          double meas = fabs(det);
          if (E.normals) {
            // immersed facet with scaled reference normal n: surface measure |det J| |J^-T n|
            const double* nr = E.normals + (qbeg + q0 + ql) * DIM;
            double nn = 0.;
#pragma unroll
            for (int i = 0; i < DIM; i++) {
              double vi = 0.;
#pragma unroll
              for (int k = 0; k < DIM; k++) vi = fma(Ji[k * DIM + i], nr[k], vi);
              nn = fma(vi, vi, nn);
            }
            meas *= sqrt(nn);
          } else if (fdim >= 0) {
            // surface measure of the face normal to reference direction fdim: |det J| |J^-T e_fdim| (row fdim of J^-1)
            double nn = 0.;
#pragma unroll
            for (int k = 0; k < DIM; k++)
#pragma unroll
              for (int i = 0; i < DIM; i++)
                if (k == fdim) nn = fma(Ji[k * DIM + i], Ji[k * DIM + i], nn);
            meas *= sqrt(nn);
          }
          sJ[ql * JS + DIM * DIM] = pt[DIM] * meas;
          if (E.rational == 1) {
            double W = 0., dW[DIM];
#pragma unroll
            for (int k = 0; k < DIM; k++) dW[k] = 0.;
            for (int a = 0; a < nb; a++) {
              double N, dxi[DIM];
              tensor_eval<DIM>(sA + ql * DIM * pm1 * 2, pm1, sMi[a], N, dxi);
              W = fma(N, sS[a], W);
#pragma unroll
              for (int k = 0; k < DIM; k++) dW[k] = fma(dxi[k], sS[a], dW[k]);
            }
            sW[ql * NA] = W;
#pragma unroll
            for (int k = 0; k < DIM; k++) sW[ql * NA + 1 + k] = dW[k];
          } else if (E.rational == 2) {
            sW[ql * NA] = Wg;
#pragma unroll
            for (int k = 0; k < DIM; k++) sW[ql * NA + 1 + k] = dWg[k];
          }
        }
        __syncthreads();
        // P3: values and physical gradients of the (scaled, rational) functions
        for (int t = tid; t < nqc * nb; t += T) {
          const int a = t % nb, ql = t / nb;
          double N, dxi[DIM];
          tensor_eval<DIM>(sA + ql * DIM * pm1 * 2, pm1, sMi[a], N, dxi);
          const double c = sS[a];
          N *= c;
#pragma unroll
          for (int k = 0; k < DIM; k++) dxi[k] *= c;
          if (E.rational) {
            // R = N / W,  dR = (dN - R dW) / W
            const double rW = 1. / sW[ql * NA];
            N *= rW;
#pragma unroll
            for (int k = 0; k < DIM; k++) dxi[k] = (dxi[k] - N * sW[ql * NA + 1 + k]) * rW;
          }
          double* b = sB + (ql * nb + a) * NA;
          double gph[DIM];
#pragma unroll
          for (int i = 0; i < DIM; i++) {
            double s = 0.;
#pragma unroll
            for (int k = 0; k < DIM; k++) s += dxi[k] * sJ[ql * JS + k * DIM + i];
            gph[i] = s;
          }
          if constexpr (NA == 4) {
            // one entry = 32 aligned bytes: two 128-bit stores (four 64-bit stores at stride 32 bytes across the lanes are an 8-way bank conflict)
            reinterpret_cast<double2*>(b)[0] = make_double2(N, gph[0]);
            reinterpret_cast<double2*>(b)[1] = make_double2(gph[1], gph[2]);
          } else {
            b[0] = N;
#pragma unroll
            for (int i = 0; i < DIM; i++) b[1 + i] = gph[i];
          }
        }
        __syncthreads();
        if (!MMA && P.ev.enabled) {
          // evaluation mode: outputs of the chunk, no integration
          const long long g0 = qbeg + q0;  // global index of the first point of the chunk
          for (int ql = tid; ql < nqc; ql += T) {
            if (P.ev.wdet) P.ev.wdet[g0 + ql] = sJ[ql * JS + DIM * DIM];
            if (P.ev.x) {
              const double* pt = sPt + ql * NA;
              double xq[DIM];
#pragma unroll
              for (int i = 0; i < DIM; i++) xq[i] = 0.;
              if (spline) {
                double W = 0.;
                for (int a = 0; a < nbg; a++) {
                  double N, dxi[DIM];
                  tensor_eval<DIM>(sAg + ql * DIM * pgm1 * 2, pgm1, sMg[a], N, dxi);
                  W = fma(N, sX[DIM * nbg + a], W);
#pragma unroll
                  for (int i = 0; i < DIM; i++) xq[i] = fma(N, sX[i * nbg + a], xq[i]);
                }
#pragma unroll
                for (int i = 0; i < DIM; i++) xq[i] /= W;
              } else {
#pragma unroll
                for (int v = 0; v < NV; v++) {
                  double f = 1.;
#pragma unroll
                  for (int d = 0; d < DIM; d++) f *= ((v >> (DIM - 1 - d)) & 1) ? pt[d] : 1. - pt[d];
#pragma unroll
                  for (int i = 0; i < DIM; i++) xq[i] = fma(sX[i * NV + v], f, xq[i]);
                }
              }
#pragma unroll
              for (int i = 0; i < DIM; i++) P.ev.x[(g0 + ql) * DIM + i] = xq[i];
            }
          }
          const int nfc = P.ev.nfields * nc;
          for (int t = tid; t < nqc * nfc; t += T) {
            const int ql = t / nfc, fc = t - ql * nfc, f = fc / nc, c = fc - f * nc;
            double u[NA];
#pragma unroll
            for (int x = 0; x < NA; x++) u[x] = 0.;
            for (int a = 0; a < nb; a++) {
              if (sDof[a] < 0) continue;
              const double ca = P.ev.coef[f * P.ev.ndofs + (long long)sDof[a] * nc + c];
              const double* Ba = sB + (ql * nb + a) * NA;
#pragma unroll
              for (int x = 0; x < NA; x++) u[x] = fma(ca, Ba[x], u[x]);
            }
            if (P.ev.values) P.ev.values[(g0 + ql) * nfc + fc] = u[0];
            if (P.ev.grads)
#pragma unroll
              for (int k = 0; k < DIM; k++) P.ev.grads[((g0 + ql) * nfc + fc) * DIM + k] = u[1 + k];
          }
          __syncthreads();
          continue;
        }
        // P4: block entries
        if constexpr (MMA) {
          const int warp = tid >> 5, lane = tid & 31;
          if (warp < ntr) {
            const int arow = warp * 8 + (lane >> 2), kk = lane & 3;
#pragma unroll
            for (int j = 0; j < NJ; j++) {
              if (pass0 + j < P.F.njobs) {
                const double* Dj = sD + j * 16;
                if ((P.F.jobinfo[pass0 + j] >> 24) == 0) {
                  // k-block = one point, k = y: A[a][y] = w sum_x B[q][a][x] D[x][y],  B[y][b] = B[q][b][y]
                  // (four points per trip: their shared-memory loads and the A fragments overlap; padding points add zeros)
                  double dcol[NA];
#pragma unroll
                  for (int x = 0; x < NA; x++) dcol[x] = kk < NA ? Dj[x * 4 + kk] : 0.;
                  const bool arow_ok = arow < nb && kk < NA;
                  const int fm = P.F.jobinfo[pass0 + j] & 255;  // matrix form of the block: its pointwise coefficient
                  for (int ql0 = 0; ql0 < nqc; ql0 += 4) {
                    double af[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                      const int ql = ql0 + u;
                      af[u] = 0.;
                      if (arow_ok && ql < nqc) {
                        const double* Ba = sB + (ql * nb + arow) * NA;
                        double t = 0.;
                        if constexpr (NA == 4) {
                          const double2 b01 = reinterpret_cast<const double2*>(Ba)[0], b23 = reinterpret_cast<const double2*>(Ba)[1];
                          t = fma(b23.y, dcol[3], fma(b23.x, dcol[2], fma(b01.y, dcol[1], b01.x * dcol[0])));
                        } else {
#pragma unroll
                          for (int x = 0; x < NA; x++) t = fma(Ba[x], dcol[x], t);
                        }
                        af[u] = t * sJ[ql * JS + DIM * DIM] * sK[ql * 2 * B2_MAX_FORMS + fm];
                      }
                    }
#pragma unroll
                    for (int tc = 0; tc < TCMAX; tc++) {
                      if (tc < ntr) {
                        const int bcol = tc * 8 + (lane >> 2);
                        const bool bok = bcol < nb && kk < NA;
                        double bf[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) bf[u] = (bok && ql0 + u < nqc) ? sB[((ql0 + u) * nb + bcol) * NA + kk] : 0.;
#pragma unroll
                        for (int u = 0; u < 4; u++) dmma884(cacc[j][tc][0], cacc[j][tc][1], af[u], bf[u]);
                      }
                    }
                  }
                } else {
                  // mass-like block (only D[0][0]): k-block = four points
                  const double d00 = Dj[0];
                  const int fm = P.F.jobinfo[pass0 + j] & 255;
                  for (int q4 = 0; q4 < nqc; q4 += 4) {
                    const int ql = q4 + kk;
                    const bool okq = ql < nqc;
                    const double af = (okq && arow < nb) ? sB[(ql * nb + arow) * NA] * sJ[ql * JS + DIM * DIM] * sK[ql * 2 * B2_MAX_FORMS + fm] * d00 : 0.;
#pragma unroll
                    for (int tc = 0; tc < TCMAX; tc++) {
                      if (tc < ntr) {
                        const int bcol = tc * 8 + (lane >> 2);
                        const double bf = (okq && bcol < nb) ? sB[(ql * nb + bcol) * NA] : 0.;
                        dmma884(cacc[j][tc][0], cacc[j][tc][1], af, bf);
                      }
                    }
                  }
                }
              }
            }
          }
        } else {
#pragma unroll
        for (int e = 0; e < EPT; e++) {
          const int entry = pass0 + e * T + tid;
          if (entry < ne2) {
            const int r = entry / ne, c = entry % ne;
            const int a = r / nc, ci = r % nc, b = c / nc, cj = c % nc;
            for (int ql = 0; ql < nqc; ql++) {
              const double* Ba = sB + (ql * nb + a) * NA;
              const double* Bb = sB + (ql * nb + b) * NA;
              const double w = sJ[ql * JS + DIM * DIM];
              double va[NA], vb[NA];
#pragma unroll
              for (int x = 0; x < NA; x++) {
                va[x] = Ba[x] * w;
                vb[x] = Bb[x];
              }
#pragma unroll
              for (int m = 0; m < B2_MAX_FORMS; m++) {
                if (m < P.F.nmat) {
                  const int t0 = P.F.termptr[(m * nc + ci) * nc + cj], t1 = P.F.termptr[(m * nc + ci) * nc + cj + 1];
                  double s = 0.;
                  for (int t = t0; t < t1; t++) {
                    const int xy = P.F.termxy[t];
                    double ax = va[0], by = vb[0];
#pragma unroll
                    for (int x = 1; x < NA; x++) {
                      if ((xy & 255) == x) ax = va[x];
                      if ((xy >> 8) == x) by = vb[x];
                    }
                    s += P.F.termval[t] * ax * by;
                  }
                  acc[m][e] += s * sK[ql * 2 * B2_MAX_FORMS + m];
                }
              }
            }
          }
        }
        }
        // linear forms: thread r owns sV[.][r]
        if (pass0 == 0 && P.F.nvec) {
          // the points are dealt to T / ne groups of threads (an element with few functions would leave most warps waiting at the
          // barrier); partial sums meet in shared memory
          const int nparts = max(1, T / ne);
          for (int it = tid; it < ne * nparts; it += T) {
            const int r = it % ne, part = it / ne;
            const int a = r / nc, ci = r % nc;
            for (int v = 0; v < P.F.nvec; v++) {
              double s = 0.;
              for (int ql = part; ql < nqc; ql += nparts) {
                const double* Ba = sB + (ql * nb + a) * NA;
                double t = 0.;
#pragma unroll
                for (int x = 0; x < NA; x++) t += P.F.vcoef[(v * nc + ci) * NA + x] * Ba[x];
                s += t * sJ[ql * JS + DIM * DIM] * sK[ql * 2 * B2_MAX_FORMS + B2_MAX_FORMS + v];
              }
              if (nparts > 1) atomicAdd(&sV[v * ne + r], s);
              else sV[v * ne + r] += s;
            }
          }
        }
        __syncthreads();
      }
      if (!MMA && P.ev.enabled) break;  // evaluation mode: one sweep over the points, nothing to scatter
      // scatter: bisection of the column in the row's sorted list
      if constexpr (MMA) {
        const int warp = tid >> 5, lane = tid & 31;
        if (warp < ntr) {
          const int a = warp * 8 + (lane >> 2);
          const int In = a < nb ? sDof[a] : -1;
          if (In >= 0) {
            const int len = sLen[a];
            const int* cols = E.colidx_b + sRow[a];
            // a row that kept its whole box of (2p+1)^DIM coupled functions (the rule away from trimmed cells and domain boundaries)
            // holds function b at the analytic position sum_d (b_d - a_d + p_d) stride_d: one load confirms it, the bisection
            // (log2(len) DEPENDENT global loads) is the fallback
            int full = 1;
#pragma unroll
            for (int d = 0; d < DIM; d++) full *= 2 * B.p[d] + 1;
            const int mia = sMi[a];
#pragma unroll
            for (int tc = 0; tc < TCMAX; tc++) {
              if (tc < ntr) {
#pragma unroll
                for (int i = 0; i < 2; i++) {
                  const int b = tc * 8 + (lane & 3) * 2 + i;
                  const int Jn = b < nb ? sDof[b] : -1;
                  if (Jn < 0) continue;
                  int lo = -1;
                  if (len == full) {
                    const int mib = sMi[b];
                    int pos = 0;
#pragma unroll
                    for (int d = 0; d < DIM; d++) pos = pos * (2 * B.p[d] + 1) + ((mib >> (8 * d)) & 255) - ((mia >> (8 * d)) & 255) + B.p[d];
                    if (cols[pos] == Jn) lo = pos;
                  }
                  if (lo < 0) {
                    int hi = len;
                    lo = 0;
                    while (lo < hi) {
                      const int mid = (lo + hi) >> 1;
                      if (cols[mid] < Jn) lo = mid + 1;
                      else hi = mid;
                    }
                  }
                  if (lo < len && cols[lo] == Jn) {
#pragma unroll
                    for (int j = 0; j < NJ; j++) {
                      if (pass0 + j < P.F.njobs) {
                        const int info = P.F.jobinfo[pass0 + j];
                        const int m = info & 255, ci = (info >> 8) & 255, cj = (info >> 16) & 255;
                        atomicAdd(P.F.values[m] + (sRow[a] * nc + (long long)ci * len) * nc + (long long)lo * nc + cj, cacc[j][tc][i]);
                      }
                    }
                  }
                }
              }
            }
          }
        }
      } else {
#pragma unroll
      for (int e = 0; e < EPT; e++) {
        const int entry = pass0 + e * T + tid;
        if (entry < ne2 && P.F.nmat) {
          const int r = entry / ne, c = entry % ne;
          const int a = r / nc, ci = r % nc, b = c / nc, cj = c % nc;
          const int In = sDof[a], Jn = sDof[b];
          if (In >= 0 && Jn >= 0) {
            const int len = sLen[a];
            const int* cols = E.colidx_b + sRow[a];
            int lo = 0, hi = len;
            while (lo < hi) {
              const int mid = (lo + hi) >> 1;
              if (cols[mid] < Jn) lo = mid + 1;
              else hi = mid;
            }
            if (lo < len && cols[lo] == Jn) {
              const long long slot = (sRow[a] * nc + (long long)ci * len) * nc + (long long)lo * nc + cj;
#pragma unroll
              for (int m = 0; m < B2_MAX_FORMS; m++)
                if (m < P.F.nmat) atomicAdd(P.F.values[m] + slot, acc[m][e]);
            }
          }
        }
      }
          }
    }
    for (int t = tid; t < P.F.nvec * ne; t += T) {
      const int v = t / ne, r = t % ne;
      const int a = r / nc, ci = r % nc;
      if (sDof[a] >= 0) atomicAdd(P.F.rhs[v] + (long long)sDof[a] * nc + ci, sV[t]);
    }
  }
}

// per-element point sets of different length: the case the element queue is for
static inline bool E_ragged(const ESParams& P) { return P.E.qoff != nullptr; }

template <int DIM, int TCMAX, int NJ>
int launch_cfg(b2_ctx* ctx, ESParams& P, int max_nq) {
  const int nb = P.B.nb, ne = P.ne, ne2 = ne * ne;
  constexpr int NA = DIM + 1, NV = 1 << DIM, JS = DIM * DIM + 1;
  int threads = (ne2 + EPT - 1) / EPT;
  threads = std::min(512, std::max(64, (threads + 31) / 32 * 32));
  if (TCMAX > 0) threads = std::min(512, std::max(64, 32 * ((nb + 7) / 8)));  // one warp per row of 8x8 tiles
  const bool spline = P.SG.enabled != 0;
  const size_t fixed = sizeof(double) * (DIM * P.pm1 * P.pm1 + DIM * P.pgm1 * P.pgm1 + (spline ? (DIM + 1) * P.nbg : DIM * NV) + nb + B2_MAX_FORMS * ne + NJ * 16) +
                       sizeof(long long) * nb + sizeof(int) * (3 * nb + P.nbg);
  const size_t per_q = sizeof(double) * (NA + DIM * P.pm1 * 2 + DIM * P.pgm1 * 2 + JS + NA + 2 * B2_MAX_FORMS + nb * NA);
  // shared memory per CTA: small CTAs (tensor-core path at low degree: 4 warps) want several CTAs per SM
  const size_t budget = (TCMAX == 4 ? 52 : 96) * 1024;
  int qc = (int)std::max<size_t>(1, (budget > fixed ? (budget - fixed) / per_q : 1));
  qc = std::min(qc, std::max(max_nq, 1));
  if (TCMAX > 0 && qc >= 4) qc &= ~3;  // mass-like blocks batch four points per k-block
  P.qchunk = qc;
  const size_t smem = fixed + per_q * qc + 16;
  if (smem > 200 * 1024) return b2_fail(ctx, B2_EUNSUPPORTED, "element too large for the element-set kernel");
  auto kern = k_assemble_elemset<DIM, TCMAX, NJ>;
  B2_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  B2_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
  per_sm = std::max(per_sm, 1);
  const long long nel = P.sel_end - P.sel_begin;
  P.queue = nullptr;
  if (E_ragged(P) && !(ctx->opts.count("elemset_queue") && ctx->opts["elemset_queue"] == 0)) {
    if (!ctx->queue && cudaMalloc(&ctx->queue, sizeof(unsigned long long)) != cudaSuccess) {
      cudaGetLastError();
      ctx->queue = nullptr;
    }
    if (ctx->queue) {
      B2_CUDA(ctx, cudaMemsetAsync(ctx->queue, 0, sizeof(unsigned long long), ctx->stream));
      P.queue = ctx->queue;
    }
  }
  const int blocks = (int)std::min<long long>(nel, (long long)ctx->sm_count * per_sm * 4);
  {
    KernelTimer timer(ctx);
    kern<<<blocks, threads, smem, ctx->stream>>>(P);
  }
  ctx->launches++;
  B2_CUDA(ctx, cudaGetLastError());
  return B2_OK;
}

template <int DIM>
int launch_dim(b2_ctx* ctx, ESParams& P, int max_nq) {
  const bool mma = DIM > 1 && P.F.njobs > 0 && P.B.nb <= 128 && !(ctx->opts.count("elemset_mma") && ctx->opts["elemset_mma"] == 0);
  if (DIM > 1 && mma) {
    if (P.B.nb <= 32) return launch_cfg<DIM == 1 ? 2 : DIM, 4, 3>(ctx, P, max_nq);
    return launch_cfg<DIM == 1 ? 2 : DIM, 16, 1>(ctx, P, max_nq);
  }
  return launch_cfg<DIM, 0, 1>(ctx, P, max_nq);
}

}  // namespace

int launch_evaluate_elemset(b2_ctx* ctx, const BasisView& B, const QuadView& Q, const GeomView& G, const SplineGeomView& SG, const ElemSetView& E, int max_nq,
                            int nfields, const double* coef, long long ndofs, double* x, double* wdet, double* values, double* grads) {
  ESParams P;
  memset(&P, 0, sizeof(P));
  P.ev.enabled = 1;
  P.ev.nfields = nfields;
  P.ev.coef = coef;
  P.ev.ndofs = ndofs;
  P.ev.x = x;
  P.ev.wdet = wdet;
  P.ev.values = values;
  P.ev.grads = grads;
  P.B = B;
  P.Q = Q;
  P.G = G;
  P.SG = SG;
  P.E = E;
  P.sel_begin = 0;
  P.sel_end = E.nsel;
  P.pm1 = 1;
  P.pgm1 = 1;
  P.nbg = 1;
  for (int d = 0; d < B.ndims; d++) {
    P.pm1 = std::max(P.pm1, B.p[d] + 1);
    if (SG.enabled) {
      P.pgm1 = std::max(P.pgm1, SG.GB.p[d] + 1);
      P.nbg *= SG.GB.p[d] + 1;
    }
  }
  P.ne = B.nb * B.ncomp;
  P.qchunk = 1;
  switch (B.ndims) {
    case 1: return launch_cfg<1, 0, 1>(ctx, P, max_nq);
    case 2: return launch_cfg<2, 0, 1>(ctx, P, max_nq);
    default: return launch_cfg<3, 0, 1>(ctx, P, max_nq);
  }
}

int launch_assemble_elemset(b2_ctx* ctx, const BasisView& B, const QuadView& Q, const GeomView& G, const SplineGeomView& SG, const ElemSetView& E, const FormView& F,
                            long long sel_begin, long long sel_end, int max_nq) {
  ESParams P;
  memset(&P.ev, 0, sizeof(P.ev));
  P.B = B;
  P.Q = Q;
  P.G = G;
  P.SG = SG;
  P.E = E;
  P.F = F;
  P.sel_begin = sel_begin;
  P.sel_end = sel_end;
  P.pm1 = 1;
  P.pgm1 = 1;
  P.nbg = 1;
  for (int d = 0; d < B.ndims; d++) {
    P.pm1 = std::max(P.pm1, B.p[d] + 1);
    if (SG.enabled) {
      P.pgm1 = std::max(P.pgm1, SG.GB.p[d] + 1);
      P.nbg *= SG.GB.p[d] + 1;
    }
  }
  P.ne = B.nb * B.ncomp;
  P.qchunk = 1;
  switch (B.ndims) {
    case 1: return launch_dim<1>(ctx, P, max_nq);
    case 2: return launch_dim<2>(ctx, P, max_nq);
    default: return launch_dim<3>(ctx, P, max_nq);
  }
}
