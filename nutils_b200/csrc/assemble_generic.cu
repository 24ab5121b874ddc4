// Generic element-integration kernel: any dimension 1..3, degree 0..4, 1..3 components, up to four
// bilinear forms (sparse coefficient tensors D_m) and four linear forms per launch.
//
// Replaces the body of the element loop the reference generates (src/nutils/evaluable.py:6773-6787,
// SURVEY.md appendix A): per element basis values/gradients at the quadrature points (Polyval /
// PolyGrad, evaluable.py:4328-4374, 4584-4634 -- here as products of tabulated 1-D factors),
// J, J^-1 and |det J| of the multilinear geometry (function.py:1207-1231, 1266-1295), the integrand
// contractions and the weighted sum over points (sample.py:951-956), the COO append + post-loop
// accumulate (evaluable.py:5383-5501, numeric.py:434-460) -- here an fp64 atomicAdd straight into
// the analytic CSR slot -- and numpy.add.at for load vectors (evaluable.py:3582-3620).
//
// Mapping: one CTA per element (grid-stride over elements).  Phase 1: threads over quadrature points
// compute J^-1 and w|det J| into shared memory.  Phase 2: threads over (point, basis function) fill
// B[q][a][0..DIM] = (N_a, grad N_a).  Phase 3: threads over the entries of the (n_e ncomp)^2 block
// accumulate over the points in registers, then scatter.  Quadrature points are processed in chunks
// so that B fits in shared memory at any degree.  This is the coverage kernel; the specialised
// kernels in assemble_fast.cu take over for the configurations that carry the benchmark.

#include <algorithm>

#include "common.cuh"

namespace {

constexpr int EPT = 8;  // block entries per thread and pass

struct GenericParams {
  BasisView B;
  QuadView Q;
  GeomView G;
  FormView F;
  long long elem_begin, elem_end;
  int plane_lo, plane_hi;  // only rows whose dimension-0 dof index lies in [plane_lo, plane_hi) are scattered
  int qchunk;   // quadrature points per chunk
  int pm1;      // max(p)+1
  int nqm;      // max(nq)
  int ne;       // nb * ncomp
};

template <int DIM>
__device__ __forceinline__ void inverse_det(const double* J, double* Ji, double& det) {
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

template <int DIM>
__global__ void __launch_bounds__(512, 1) k_assemble_generic(const GenericParams P) {
  constexpr int NA = DIM + 1;
  constexpr int NV = 1 << DIM;
  constexpr int JS = DIM * DIM + 1;
  const BasisView& B = P.B;
  const int tid = threadIdx.x, T = blockDim.x;
  const int nb = B.nb, nc = B.ncomp, ne = P.ne, ne2 = ne * ne;
  const int nqt = P.Q.nqt, qc = P.qchunk;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sA = reinterpret_cast<double*>(smem_raw);  // [DIM][2][pm1][nqm]
  double* sX = sA + DIM * 2 * P.pm1 * P.nqm;         // [DIM][NV]
  double* sJ = sX + DIM * NV;                        // [qc][JS]   J^-1 (k,i) then w|det|
  double* sB = sJ + qc * JS;                         // [qc][nb][NA]
  double* sV = sB + qc * nb * NA;                    // [nvec][ne]
  long long* sR = reinterpret_cast<long long*>(sV + B2_MAX_FORMS * ne);  // [nb] row start (basis units)
  int* sI = reinterpret_cast<int*>(sR + nb);         // [nb][DIM] dof index per dim
  int* sLo = sI + nb * DIM;                          // [nb][DIM]
  int* sW = sLo + nb * DIM;                          // [nb][DIM]

  for (long long elem = P.elem_begin + blockIdx.x; elem < P.elem_end; elem += gridDim.x) {
    int ie[3] = {0, 0, 0};
    {
      long long r = elem;
      for (int d = DIM - 1; d >= 0; d--) {
        ie[d] = (int)(r % B.nel[d]);
        r /= B.nel[d];
      }
    }
    __syncthreads();  // previous element fully scattered before shared memory is reused
    // 1-D tables of this element
    for (int t = tid; t < DIM * 2 * P.pm1 * P.nqm; t += T) {
      const int q = t % P.nqm;
      int r = t / P.nqm;
      const int a = r % P.pm1;
      r /= P.pm1;
      const int k = r % 2, d = r / 2;
      double v = 0.;
      if (a <= B.p[d] && q < P.Q.nq[d]) v = P.Q.tab[d][((B.setidx[d][ie[d]] * 2 + k) * (B.p[d] + 1) + a) * P.Q.nq[d] + q];
      sA[t] = v;
    }
    // vertex coordinates, C-order vertices
    for (int t = tid; t < DIM * NV; t += T) {
      const int v = t % NV, i = t / NV;
      long long off = 0;
      for (int d = 0; d < DIM; d++) off += (long long)(ie[d] + ((v >> (DIM - 1 - d)) & 1)) * P.G.stride[d];
      sX[t] = P.G.nodes[i * P.G.nnodes + off];
    }
    // dofs and CSR row data of the element's basis functions
    for (int a = tid; a < nb; a += T) {
      int r = a, i[3] = {0, 0, 0};
      for (int d = DIM - 1; d >= 0; d--) {
        i[d] = B.start[d][ie[d]] + r % (B.p[d] + 1);
        r /= B.p[d] + 1;
      }
      for (int d = 0; d < DIM; d++) {
        sI[a * DIM + d] = i[d];
        sLo[a * DIM + d] = B.lo[d][i[d]];
        sW[a * DIM + d] = B.wid[d][i[d]];
      }
      sR[a] = row_start_basis<DIM>(B, i);
    }
    for (int t = tid; t < P.F.nvec * ne; t += T) sV[t] = 0.;
    __syncthreads();

    for (int pass0 = 0; pass0 < ne2 || pass0 == 0; pass0 += T * EPT) {
      double acc[B2_MAX_FORMS][EPT];
#pragma unroll
      for (int m = 0; m < B2_MAX_FORMS; m++)
#pragma unroll
        for (int e = 0; e < EPT; e++) acc[m][e] = 0.;

      for (int q0 = 0; q0 < nqt; q0 += qc) {
        const int nqc = min(qc, nqt - q0);
        // phase 1: geometry at the points of the chunk
        for (int ql = tid; ql < nqc; ql += T) {
          int qd[3] = {0, 0, 0}, r = q0 + ql;
          for (int d = DIM - 1; d >= 0; d--) {
            qd[d] = r % P.Q.nq[d];
            r /= P.Q.nq[d];
          }
          double xi[3] = {0., 0., 0.}, w = 1.;
          for (int d = 0; d < DIM; d++) {
            xi[d] = P.Q.x[d][qd[d]];
            w *= P.Q.w[d][qd[d]];
          }
          double J[DIM * DIM];
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
                f *= d == k ? (bit ? 1. : -1.) : (bit ? xi[d] : 1. - xi[d]);
              }
#pragma unroll
              for (int i = 0; i < DIM; i++) J[i * DIM + k] += sX[i * NV + v] * f;
            }
          }
          double Ji[DIM * DIM], det;
          inverse_det<DIM>(J, Ji, det);
#pragma unroll
          for (int i = 0; i < DIM * DIM; i++) sJ[ql * JS + i] = Ji[i];
          sJ[ql * JS + DIM * DIM] = w * fabs(det);
        }
        __syncthreads();
        // phase 2: values and physical gradients
        for (int t = tid; t < nqc * nb; t += T) {
          const int a = t % nb, ql = t / nb;
          int qd[3] = {0, 0, 0}, ad[3] = {0, 0, 0};
          {
            int r = q0 + ql;
            for (int d = DIM - 1; d >= 0; d--) {
              qd[d] = r % P.Q.nq[d];
              r /= P.Q.nq[d];
            }
            r = a;
            for (int d = DIM - 1; d >= 0; d--) {
              ad[d] = r % (B.p[d] + 1);
              r /= B.p[d] + 1;
            }
          }
          double val[DIM], der[DIM];
#pragma unroll
          for (int d = 0; d < DIM; d++) {
            val[d] = sA[((d * 2 + 0) * P.pm1 + ad[d]) * P.nqm + qd[d]];
            der[d] = sA[((d * 2 + 1) * P.pm1 + ad[d]) * P.nqm + qd[d]];
          }
          double N = 1., dxi[DIM];
#pragma unroll
          for (int d = 0; d < DIM; d++) N *= val[d];
#pragma unroll
          for (int k = 0; k < DIM; k++) {
            double g = 1.;
#pragma unroll
            for (int d = 0; d < DIM; d++) g *= d == k ? der[d] : val[d];
            dxi[k] = g;
          }
          double* b = sB + (ql * nb + a) * NA;
          b[0] = N;
#pragma unroll
          for (int i = 0; i < DIM; i++) {
            double s = 0.;
#pragma unroll
            for (int k = 0; k < DIM; k++) s += dxi[k] * sJ[ql * JS + k * DIM + i];
            b[1 + i] = s;
          }
        }
        __syncthreads();
        // phase 3: block entries
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
                  acc[m][e] += s;
                }
              }
            }
          }
        }
        // linear forms: thread r owns sV[.][r]
        if (pass0 == 0) {
          for (int r = tid; r < ne; r += T) {
            const int a = r / nc, ci = r % nc;
            for (int v = 0; v < P.F.nvec; v++) {
              double s = 0.;
              for (int ql = 0; ql < nqc; ql++) {
                const double* Ba = sB + (ql * nb + a) * NA;
                double t = 0.;
#pragma unroll
                for (int x = 0; x < NA; x++) t += P.F.vcoef[(v * nc + ci) * NA + x] * Ba[x];
                s += t * sJ[ql * JS + DIM * DIM];
              }
              sV[v * ne + r] += s;
            }
          }
        }
        __syncthreads();
      }
      // scatter
#pragma unroll
      for (int e = 0; e < EPT; e++) {
        const int entry = pass0 + e * T + tid;
        if (entry < ne2) {
          const int r = entry / ne, c = entry % ne;
          const int a = r / nc, ci = r % nc, b = c / nc, cj = c % nc;
          int wa[3] = {1, 1, 1};
          long long pos = 0;
#pragma unroll
          for (int d = 0; d < DIM; d++) {
            wa[d] = sW[a * DIM + d];
            pos = pos * wa[d] + (sI[b * DIM + d] - sLo[a * DIM + d]);
          }
          const long long w = (long long)wa[0] * wa[1] * wa[2];
          const long long slot = (sR[a] * nc + (long long)ci * w) * nc + pos * nc + cj;
          if (sI[a * DIM] >= P.plane_lo && sI[a * DIM] < P.plane_hi) {
#pragma unroll
            for (int m = 0; m < B2_MAX_FORMS; m++)
              if (m < P.F.nmat) atomicAdd(P.F.values[m] + slot, acc[m][e]);
          }
        }
      }
    }
    for (int t = tid; t < P.F.nvec * ne; t += T) {
      const int v = t / ne, r = t % ne;
      const int a = r / nc, ci = r % nc;
      long long I = 0;
      for (int d = 0; d < DIM; d++) I = I * B.ndofs[d] + sI[a * DIM + d];
      if (sI[a * DIM] >= P.plane_lo && sI[a * DIM] < P.plane_hi) atomicAdd(P.F.rhs[v] + I * nc + ci, sV[t]);
    }
  }
}

template <int DIM>
int launch_dim(b2_ctx* ctx, GenericParams& P) {
  const int nb = P.B.nb, ne = P.ne, ne2 = ne * ne;
  constexpr int NA = DIM + 1, NV = 1 << DIM, JS = DIM * DIM + 1;
  int threads = (ne2 + EPT - 1) / EPT;
  threads = std::min(512, std::max(32, (threads + 31) / 32 * 32));
  // shared memory budget: chunk the quadrature points so that everything fits in ~96 KB (>= 2 CTAs/SM)
  const size_t fixed = sizeof(double) * (DIM * 2 * P.pm1 * P.nqm + DIM * NV + B2_MAX_FORMS * ne) + sizeof(long long) * nb + sizeof(int) * 3 * nb * DIM;
  const size_t per_q = sizeof(double) * (JS + nb * NA);
  const size_t budget = 96 * 1024;
  int qc = (int)std::max<size_t>(1, (budget > fixed ? (budget - fixed) / per_q : 1));
  qc = std::min(qc, P.Q.nqt);
  P.qchunk = qc;
  const size_t smem = fixed + per_q * qc + 16;
  if (smem > 200 * 1024) return b2_fail(ctx, B2_EUNSUPPORTED, "element too large for the generic kernel");
  B2_CUDA(ctx, cudaFuncSetAttribute(k_assemble_generic<DIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  B2_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_assemble_generic<DIM>, threads, smem));
  per_sm = std::max(per_sm, 1);
  const long long nel = P.elem_end - P.elem_begin;
  const int blocks = (int)std::min<long long>(nel, (long long)ctx->sm_count * per_sm * 4);
  {
    KernelTimer timer(ctx);
    k_assemble_generic<DIM><<<blocks, threads, smem, ctx->stream>>>(P);
  }
  ctx->launches++;
  B2_CUDA(ctx, cudaGetLastError());
  return B2_OK;
}

}  // namespace

int launch_assemble_generic(b2_ctx* ctx, const BasisView& B, const QuadView& Q, const GeomView& G, const FormView& F,
                            long long elem_begin, long long elem_end, int plane_lo, int plane_hi) {
  GenericParams P;
  P.B = B;
  P.Q = Q;
  P.G = G;
  P.F = F;
  P.elem_begin = elem_begin;
  P.elem_end = elem_end;
  P.plane_lo = plane_lo;
  P.plane_hi = plane_hi;
  P.pm1 = 1;
  P.nqm = 1;
  for (int d = 0; d < B.ndims; d++) {
    P.pm1 = std::max(P.pm1, B.p[d] + 1);
    P.nqm = std::max(P.nqm, Q.nq[d]);
  }
  P.ne = B.nb * B.ncomp;
  P.qchunk = 1;
  switch (B.ndims) {
    case 1: return launch_dim<1>(ctx, P);
    case 2: return launch_dim<2>(ctx, P);
    default: return launch_dim<3>(ctx, P);
  }
}
