// Specialised element-integration kernel for scalar forms on 3-D tensor-product splines (p = 1, 2):
// the configuration that carries the benchmark (BASELINE.json configs[1]).
//
// Same arithmetic as the generic kernel (and as the reference's generated loop),
// reorganised for the FP64 pipe of sm_100a:
//
//   * one warp per element (north star), grid-stride in C order so neighbouring elements -- which
//     share CSR rows -- are in flight together and their atomics meet in L2;
//   * geometry once per quadrature point: lane q evaluates J of the trilinear map, J^-1, |det J| and
//     the reference-space coefficient  Ghat = w |det J| J^-1 Kc J^-T  (6 values) and  m = rho w |det J|;
//   * SUM FACTORISATION instead of the reference's O(nq n_e^2) einsums (evaluable.py:1552-1564,
//     1949-1960): the element matrix  A[a,b] = sum_{kl} sum_q Ghat_kl(q) prod_d W_d[a_d b_d][q_d]
//     is contracted one direction at a time (q3, then q2, then q1) through per-warp shared memory,
//     ~20 kFMA per element for K and M together at p=2 instead of ~79 k;
//   * scatter: fp64 atomicAdd (RED) into the analytic CSR slot (pattern.cu), load vector likewise.
//
// Bound: see DESIGN.md -- with per-lane RED the kernel is limited by the LSU's one fp64 atomic lane
// per clock per SM (profiles/microbench), not by HBM; that is the next thing to remove.

#include <algorithm>
#include <cstring>

#include "common.cuh"

namespace {

enum { KIND_STIFF = 0, KIND_MASS = 1 };

struct FastParams {
  BasisView B;
  QuadView Q;
  GeomView G;
  long long elem_begin, elem_end;
  int nforms;
  int kind[2];
  double coef[2][6];  // stiffness: symmetric conductivity (00,01,02,11,12,22); mass: density in [0]
  double* values[2];
  int nvec;
  double vcoef;       // load: int vcoef N_a
  double* rhs;
  int nsets[3];
};

template <int P>
struct Cfg {
  static constexpr int NB1 = P + 1, NQ1 = P + 1;
  static constexpr int NPAIR = NB1 * NB1;
  static constexpr int NB = NB1 * NB1 * NB1, NQ = NQ1 * NQ1 * NQ1;
  static constexpr int NSLOT = NPAIR * NPAIR;
  static constexpr int SPL = (NSLOT + 31) / 32;  // slots per lane
  // per-warp shared memory (doubles)
  static constexpr int G_OFF = 0, G_SZ = NQ * 7;
  static constexpr int T1_OFF = G_OFF + G_SZ, T1_SZ = 10 * NQ1 * NPAIR;
  static constexpr int T2_OFF = T1_OFF + T1_SZ, T2_SZ = 5 * NSLOT;
  static constexpr int WARP_DOUBLES = T2_OFF + T2_SZ;
  static constexpr int WARP_INTS = 9 * 3 + 3;  // lo, wid, cum per dim and local index; start per dim
};

__device__ __forceinline__ int sym6(int k, int l) {
  // index of (k,l) in (00,01,02,11,12,22)
  const int a = k < l ? k : l, b = k < l ? l : k;
  return a == 0 ? b : (a == 1 ? 2 + b : 5);
}

template <int P, int NF>
__global__ void __launch_bounds__(256, 2) k_assemble_scalar3d(const FastParams prm) {
  using C = Cfg<P>;
  constexpr int NB1 = C::NB1, NQ1 = C::NQ1, NPAIR = C::NPAIR;
  const BasisView& B = prm.B;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nwarp = blockDim.x >> 5;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  // CTA-wide 1-D tables of every coefficient set: [dim][set][2][NB1][NQ1]
  double* sTab = reinterpret_cast<double*>(smem_raw);
  const int tabsz[3] = {prm.nsets[0] * 2 * NB1 * NQ1, prm.nsets[1] * 2 * NB1 * NQ1, prm.nsets[2] * 2 * NB1 * NQ1};
  const int taboff[3] = {0, tabsz[0], tabsz[0] + tabsz[1]};
  const int tabtot = tabsz[0] + tabsz[1] + tabsz[2];
  double* sQ = sTab + tabtot;  // [3][NQ1] points, [3][NQ1] weights
  double* sWarp = sQ + 6 * NQ1 + (wib * C::WARP_DOUBLES);
  int* sInt = reinterpret_cast<int*>(sQ + 6 * NQ1 + nwarp * C::WARP_DOUBLES) + wib * C::WARP_INTS;
  for (int d = 0; d < 3; d++)
    for (int t = threadIdx.x; t < tabsz[d]; t += blockDim.x) sTab[taboff[d] + t] = prm.Q.tab[d][t];
  for (int t = threadIdx.x; t < 3 * NQ1; t += blockDim.x) {
    sQ[t] = prm.Q.x[t / NQ1][t % NQ1];
    sQ[3 * NQ1 + t] = prm.Q.w[t / NQ1][t % NQ1];
  }
  __syncthreads();
  double* sG = sWarp + C::G_OFF;
  double* sT1 = sWarp + C::T1_OFF;
  double* sT2 = sWarp + C::T2_OFF;
  int* sLo = sInt;        // [3][NB1]  start_d - lo_d[i_d]
  int* sWid = sInt + 9;   // [3][NB1]
  int* sCum = sInt + 18;  // [3][NB1]
  int* sSt = sInt + 27;   // [3]

  const long long warp0 = (long long)blockIdx.x * nwarp + wib;
  const long long nwarps = (long long)gridDim.x * nwarp;
  for (long long elem = prm.elem_begin + warp0; elem < prm.elem_end; elem += nwarps) {
    int ie[3];
    {
      long long r = elem;
      ie[2] = (int)(r % B.nel[2]); r /= B.nel[2];
      ie[1] = (int)(r % B.nel[1]); r /= B.nel[1];
      ie[0] = (int)r;
    }
    const double* A[3];  // tables of the element's coefficient sets: [2][NB1][NQ1]
#pragma unroll
    for (int d = 0; d < 3; d++) A[d] = sTab + taboff[d] + B.setidx[d][ie[d]] * 2 * NB1 * NQ1;
    __syncwarp();
    if (lane < 3 * NB1) {
      const int d = lane / NB1, a = lane % NB1;
      const int st = B.start[d][ie[d]], i = st + a;
      sLo[d * 3 + a] = st - B.lo[d][i];
      sWid[d * 3 + a] = B.wid[d][i];
      sCum[d * 3 + a] = B.cum[d][i];
      if (a == 0) sSt[d] = st;
    }
    // ---- geometry: lane q ----
    for (int q = lane; q < C::NQ; q += 32) {
      const int q3 = q % NQ1, q2 = (q / NQ1) % NQ1, q1 = q / (NQ1 * NQ1);
      const double x1 = sQ[q1], x2 = sQ[NQ1 + q2], x3 = sQ[2 * NQ1 + q3];
      const double w = sQ[3 * NQ1 + q1] * sQ[4 * NQ1 + q2] * sQ[5 * NQ1 + q3];
      double J[9];
      const long long base = ie[0] * prm.G.stride[0] + ie[1] * prm.G.stride[1] + ie[2] * prm.G.stride[2];
#pragma unroll
      for (int i = 0; i < 3; i++) {
        const double* X = prm.G.nodes + i * prm.G.nnodes + base;
        const double c000 = __ldg(X), c001 = __ldg(X + prm.G.stride[2]);
        const double c010 = __ldg(X + prm.G.stride[1]), c011 = __ldg(X + prm.G.stride[1] + prm.G.stride[2]);
        const double c100 = __ldg(X + prm.G.stride[0]), c101 = __ldg(X + prm.G.stride[0] + prm.G.stride[2]);
        const double c110 = __ldg(X + prm.G.stride[0] + prm.G.stride[1]), c111 = __ldg(X + prm.G.stride[0] + prm.G.stride[1] + prm.G.stride[2]);
        // along xi3
        const double d00 = c001 - c000, d01 = c011 - c010, d10 = c101 - c100, d11 = c111 - c110;
        const double m00 = fma(x3, d00, c000), m01 = fma(x3, d01, c010), m10 = fma(x3, d10, c100), m11 = fma(x3, d11, c110);
        const double e0 = fma(x2, d01 - d00, d00), e1 = fma(x2, d11 - d10, d10);
        J[i * 3 + 2] = fma(x1, e1 - e0, e0);
        // along xi2
        const double f0 = m01 - m00, f1 = m11 - m10;
        const double n0 = fma(x2, f0, m00), n1 = fma(x2, f1, m10);
        J[i * 3 + 1] = fma(x1, f1 - f0, f0);
        J[i * 3 + 0] = n1 - n0;
      }
      const double c00 = J[4] * J[8] - J[5] * J[7];
      const double c01 = J[5] * J[6] - J[3] * J[8];
      const double c02 = J[3] * J[7] - J[4] * J[6];
      const double det = J[0] * c00 + J[1] * c01 + J[2] * c02;
      const double r = 1. / det;
      double Ji[9];  // Ji[k*3+i] = d xi_k / d x_i
      Ji[0] = c00 * r; Ji[1] = (J[2] * J[7] - J[1] * J[8]) * r; Ji[2] = (J[1] * J[5] - J[2] * J[4]) * r;
      Ji[3] = c01 * r; Ji[4] = (J[0] * J[8] - J[2] * J[6]) * r; Ji[5] = (J[2] * J[3] - J[0] * J[5]) * r;
      Ji[6] = c02 * r; Ji[7] = (J[1] * J[6] - J[0] * J[7]) * r; Ji[8] = (J[0] * J[4] - J[1] * J[3]) * r;
      const double wd = w * fabs(det);
      double* g = sG + q * 7;
      // stiffness coefficient of form 0 (if any form is a stiffness form they share Ghat up to Kc; NF<=2 keeps one of each kind)
      double kc[6] = {1., 0., 0., 1., 0., 1.};
#pragma unroll
      for (int f = 0; f < NF; f++) {
        if (prm.kind[f] == KIND_STIFF) {
#pragma unroll
          for (int t = 0; t < 6; t++) kc[t] = prm.coef[f][t];
        }
      }
      // T = Ji Kc  (3x3), Ghat = T Ji^T
      double T[9];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        T[k * 3 + 0] = Ji[k * 3] * kc[0] + Ji[k * 3 + 1] * kc[1] + Ji[k * 3 + 2] * kc[2];
        T[k * 3 + 1] = Ji[k * 3] * kc[1] + Ji[k * 3 + 1] * kc[3] + Ji[k * 3 + 2] * kc[4];
        T[k * 3 + 2] = Ji[k * 3] * kc[2] + Ji[k * 3 + 1] * kc[4] + Ji[k * 3 + 2] * kc[5];
      }
      int t = 0;
#pragma unroll
      for (int k = 0; k < 3; k++)
#pragma unroll
        for (int l = k; l < 3; l++) g[t++] = wd * (T[k * 3] * Ji[l * 3] + T[k * 3 + 1] * Ji[l * 3 + 1] + T[k * 3 + 2] * Ji[l * 3 + 2]);
      g[6] = wd;
    }
    __syncwarp();

    // ---- load vector: f_a = vcoef sum_q w|det| N_a(q) ----
    if (prm.nvec) {
      for (int a = lane; a < C::NB; a += 32) {
        const int a3 = a % NB1, a2 = (a / NB1) % NB1, a1 = a / (NB1 * NB1);
        double s = 0.;
        for (int q1 = 0; q1 < NQ1; q1++) {
          double s2 = 0.;
          for (int q2 = 0; q2 < NQ1; q2++) {
            double s3 = 0.;
#pragma unroll
            for (int q3 = 0; q3 < NQ1; q3++) s3 = fma(A[2][a3 * NQ1 + q3], sG[((q1 * NQ1 + q2) * NQ1 + q3) * 7 + 6], s3);
            s2 = fma(A[1][a2 * NQ1 + q2], s3, s2);
          }
          s = fma(A[0][a1 * NQ1 + q1], s2, s);
        }
        const long long I = ((long long)(sSt[0] + a1) * B.ndofs[1] + sSt[1] + a2) * B.ndofs[2] + sSt[2] + a3;
        atomicAdd(prm.rhs + I, s * prm.vcoef);
      }
    }

    double acc[NF][C::SPL][NPAIR];
#pragma unroll
    for (int f = 0; f < NF; f++)
#pragma unroll
      for (int s = 0; s < C::SPL; s++)
#pragma unroll
        for (int t = 0; t < NPAIR; t++) acc[f][s][t] = 0.;

    for (int q1 = 0; q1 < NQ1; q1++) {
      // ---- stage 1: contract q3.  work item (q2, pr3) -> T1[term][q2][pr3], term = k*3+l (9) and 9 = mass ----
      for (int wi = lane; wi < NQ1 * NPAIR; wi += 32) {
        const int pr3 = wi % NPAIR, q2 = wi / NPAIR;
        const int a3 = pr3 / NB1, b3 = pr3 % NB1;
        double t1[10];
#pragma unroll
        for (int t = 0; t < 10; t++) t1[t] = 0.;
#pragma unroll
        for (int q3 = 0; q3 < NQ1; q3++) {
          const double* g = sG + ((q1 * NQ1 + q2) * NQ1 + q3) * 7;
          const double va = A[2][a3 * NQ1 + q3], da = A[2][(NB1 + a3) * NQ1 + q3];
          const double vb = A[2][b3 * NQ1 + q3], db = A[2][(NB1 + b3) * NQ1 + q3];
          const double w00 = va * vb, w01 = va * db, w10 = da * vb, w11 = da * db;
#pragma unroll
          for (int k = 0; k < 3; k++)
#pragma unroll
            for (int l = 0; l < 3; l++) {
              const double w = k == 2 ? (l == 2 ? w11 : w10) : (l == 2 ? w01 : w00);
              t1[k * 3 + l] = fma(w, g[sym6(k, l)], t1[k * 3 + l]);
            }
          t1[9] = fma(w00, g[6], t1[9]);
        }
#pragma unroll
        for (int t = 0; t < 10; t++) sT1[(t * NQ1 + q2) * NPAIR + pr3] = t1[t];
      }
      __syncwarp();
      // ---- stage 2: contract q2.  work item (pr3, b2) -> T2[g][pr2][pr3] for a2 = 0..P; g = (k==0)*2 + (l==0), 4 = mass ----
      for (int wi = lane; wi < NPAIR * NB1; wi += 32) {
        const int pr3 = wi % NPAIR, b2 = wi / NPAIR;
        double t2[5][NB1];
#pragma unroll
        for (int g = 0; g < 5; g++)
#pragma unroll
          for (int a2 = 0; a2 < NB1; a2++) t2[g][a2] = 0.;
#pragma unroll
        for (int q2 = 0; q2 < NQ1; q2++) {
          double t[10];
#pragma unroll
          for (int k = 0; k < 10; k++) t[k] = sT1[(k * NQ1 + q2) * NPAIR + pr3];
          const double vb = A[1][b2 * NQ1 + q2], db = A[1][(NB1 + b2) * NQ1 + q2];
          // per group: the part multiplied by the VALUE of N_a2 (pv) and by its DERIVATIVE (pd); dim index 1 <-> k,l == 1
          // terms (k,l): k==1 -> derivative on a2; l==1 -> derivative on b2
          double pv[5], pd[5];
          // group 3: k==0,l==0 -> term 00
          pv[3] = vb * t[0]; pd[3] = 0.;
          // group 2: k==0,l!=0 -> terms 01 (l==1: db), 02 (vb)
          pv[2] = db * t[1] + vb * t[2]; pd[2] = 0.;
          // group 1: k!=0,l==0 -> terms 10 (k==1: deriv on a2), 20 (value)
          pv[1] = vb * t[6]; pd[1] = vb * t[3];
          // group 0: k!=0,l!=0 -> terms 11 (da,db), 12 (da,vb), 21 (va,db), 22 (va,vb)
          pv[0] = db * t[7] + vb * t[8]; pd[0] = db * t[4] + vb * t[5];
          pv[4] = vb * t[9]; pd[4] = 0.;
#pragma unroll
          for (int a2 = 0; a2 < NB1; a2++) {
            const double va = A[1][a2 * NQ1 + q2], da = A[1][(NB1 + a2) * NQ1 + q2];
            t2[0][a2] = fma(va, pv[0], fma(da, pd[0], t2[0][a2]));
            t2[1][a2] = fma(va, pv[1], fma(da, pd[1], t2[1][a2]));
            t2[2][a2] = fma(va, pv[2], t2[2][a2]);
            t2[3][a2] = fma(va, pv[3], t2[3][a2]);
            t2[4][a2] = fma(va, pv[4], t2[4][a2]);
          }
        }
#pragma unroll
        for (int g = 0; g < 5; g++)
#pragma unroll
          for (int a2 = 0; a2 < NB1; a2++) sT2[(g * NPAIR + a2 * NB1 + b2) * NPAIR + pr3] = t2[g][a2];
      }
      __syncwarp();
      // ---- stage 3: contract q1 into the accumulators.  slot = pr2*NPAIR + pr3 ----
      double va1[NB1], da1[NB1];
#pragma unroll
      for (int a = 0; a < NB1; a++) {
        va1[a] = A[0][a * NQ1 + q1];
        da1[a] = A[0][(NB1 + a) * NQ1 + q1];
      }
#pragma unroll
      for (int s = 0; s < C::SPL; s++) {
        const int slot = s * 32 + lane;
        if (slot < C::NSLOT) {
          const double g0 = sT2[0 * C::NSLOT + slot], g1 = sT2[1 * C::NSLOT + slot], g2 = sT2[2 * C::NSLOT + slot], g3 = sT2[3 * C::NSLOT + slot];
          const double gm = sT2[4 * C::NSLOT + slot];
          // group bits: g = (k==0)*2 + (l==0): g3 -> d/d on both a1,b1; g2 -> derivative on a1 only; g1 -> on b1 only; g0 -> values
#pragma unroll
          for (int b1 = 0; b1 < NB1; b1++) {
            const double ud = da1[b1] * g3 + va1[b1] * g2;  // multiplied by derivative of a1
            const double uv = da1[b1] * g1 + va1[b1] * g0;  // multiplied by value of a1
            const double um = va1[b1] * gm;
#pragma unroll
            for (int a1 = 0; a1 < NB1; a1++) {
#pragma unroll
              for (int f = 0; f < NF; f++) {
                if (prm.kind[f] == KIND_STIFF)
                  acc[f][s][a1 * NB1 + b1] = fma(da1[a1], ud, fma(va1[a1], uv, acc[f][s][a1 * NB1 + b1]));
                else
                  acc[f][s][a1 * NB1 + b1] = fma(va1[a1], um, acc[f][s][a1 * NB1 + b1]);
              }
            }
          }
        }
      }
      __syncwarp();
    }

    // ---- scatter ----
#pragma unroll
    for (int s = 0; s < C::SPL; s++) {
      const int slot = s * 32 + lane;
      if (slot < C::NSLOT) {
        const int pr3 = slot % NPAIR, pr2 = slot / NPAIR;
        const int a3 = pr3 / NB1, b3 = pr3 % NB1, a2 = pr2 / NB1, b2 = pr2 % NB1;
#pragma unroll
        for (int a1 = 0; a1 < NB1; a1++) {
          const int w0 = sWid[a1], w1 = sWid[3 + a2], w2 = sWid[6 + a3];
          const long long R = ((long long)sCum[a1] * B.W[1] + (long long)w0 * sCum[3 + a2]) * B.W[2] + (long long)w0 * w1 * sCum[6 + a3];
#pragma unroll
          for (int b1 = 0; b1 < NB1; b1++) {
            const long long pos = ((long long)(sLo[a1] + b1) * w1 + (sLo[3 + a2] + b2)) * w2 + (sLo[6 + a3] + b3);
#pragma unroll
            for (int f = 0; f < NF; f++)
              atomicAdd(prm.values[f] + R + pos, prm.kind[f] == KIND_MASS ? acc[f][s][a1 * NB1 + b1] * prm.coef[f][0] : acc[f][s][a1 * NB1 + b1]);
          }
        }
      }
    }
  }
}

template <int P, int NF>
int launch_cfg(b2_ctx* ctx, const FastParams& prm) {
  using C = Cfg<P>;
  const int warps = 8, threads = warps * 32;
  const int tabtot = (prm.nsets[0] + prm.nsets[1] + prm.nsets[2]) * 2 * C::NB1 * C::NQ1;
  const size_t smem = sizeof(double) * (tabtot + 6 * C::NQ1 + warps * C::WARP_DOUBLES) + sizeof(int) * warps * C::WARP_INTS + 16;
  if (smem > 200 * 1024) return B2_EUNSUPPORTED;
  auto kern = k_assemble_scalar3d<P, NF>;
  B2_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  B2_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
  per_sm = std::max(per_sm, 1);
  const long long nel = prm.elem_end - prm.elem_begin;
  const int blocks = (int)std::min<long long>((nel + warps - 1) / warps, (long long)ctx->sm_count * per_sm);
  {
    KernelTimer timer(ctx);
    kern<<<blocks, threads, smem, ctx->stream>>>(prm);
  }
  ctx->launches++;
  B2_CUDA(ctx, cudaGetLastError());
  return B2_OK;
}

}  // namespace

int launch_assemble_fast(b2_ctx* ctx, const BasisView& B, const QuadView& Q, const GeomView& G, const FormView& F,
                         const double* const* D_host, const double* const* C_host, long long elem_begin, long long elem_end) {
  if (B.ndims != 3 || B.ncomp != 1) return B2_EUNSUPPORTED;
  const int P = B.p[0];
  if (B.p[1] != P || B.p[2] != P || (P != 1 && P != 2)) return B2_EUNSUPPORTED;
  for (int d = 0; d < 3; d++)
    if (Q.nq[d] != P + 1) return B2_EUNSUPPORTED;
  if (F.nmat < 1 || F.nmat > 2 || F.nvec > 1) return B2_EUNSUPPORTED;
  FastParams prm;
  memset(&prm, 0, sizeof(prm));
  prm.B = B;
  prm.Q = Q;
  prm.G = G;
  prm.elem_begin = elem_begin;
  prm.elem_end = elem_end;
  prm.nforms = F.nmat;
  int nstiff = 0, nmass = 0;
  for (int m = 0; m < F.nmat; m++) {
    const double* D = D_host[m];  // [4][4]
    bool gradgrad = false, mass = D[0] != 0., mixed = false, symmetric = true;
    for (int x = 1; x < 4; x++) {
      if (D[x] != 0. || D[x * 4] != 0.) mixed = true;
      for (int y = 1; y < 4; y++) {
        if (D[x * 4 + y] != 0.) gradgrad = true;
        if (D[x * 4 + y] != D[y * 4 + x]) symmetric = false;
      }
    }
    if (mixed || !symmetric || (gradgrad && mass) || (!gradgrad && !mass)) return B2_EUNSUPPORTED;
    if (gradgrad) {
      prm.kind[m] = KIND_STIFF;
      const double kc[6] = {D[5], D[6], D[7], D[10], D[11], D[15]};
      for (int t = 0; t < 6; t++) prm.coef[m][t] = kc[t];
      nstiff++;
    } else {
      prm.kind[m] = KIND_MASS;
      prm.coef[m][0] = D[0];
      nmass++;
    }
    prm.values[m] = F.values[m];
  }
  if (nstiff > 1 || nmass > 1) return B2_EUNSUPPORTED;
  prm.nvec = F.nvec;
  if (F.nvec) {
    const double* Cv = C_host[0];  // [4]
    if (Cv[1] != 0. || Cv[2] != 0. || Cv[3] != 0.) return B2_EUNSUPPORTED;
    prm.vcoef = Cv[0];
    prm.rhs = F.rhs[0];
  }
  for (int d = 0; d < 3; d++) prm.nsets[d] = B.nsets[d];
  if (P == 1) return F.nmat == 1 ? launch_cfg<1, 1>(ctx, prm) : launch_cfg<1, 2>(ctx, prm);
  return F.nmat == 1 ? launch_cfg<2, 1>(ctx, prm) : launch_cfg<2, 2>(ctx, prm);
}
