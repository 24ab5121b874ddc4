// Owner-computes ("row-sum") integration kernel for scalar forms on 3-D maximal-smoothness
// tensor-product splines: every stored CSR value is produced by exactly ONE thread and written
// once with a plain store -- no atomics, no zero-fill, no cross-GPU exchange.
//
// What it replaces in the reference: the whole generated element loop (SURVEY.md appendix A;
// src/nutils/evaluable.py:6773-6787) PLUS the post-loop argsort/unique/accumulate that sums the
// element blocks into CSR values (evaluable.py:588-616, 5646-5682; numeric.py:434-460).  Instead of
// forming 27x27 element blocks and scattering them (1.5 G scatter-adds per matrix at 128^3, which
// is what bounds assemble_fast.cu: the LSU retires ~280 G fp64 RED/s chip-wide), the sum over the
// elements in supp(N_i) n supp(N_j) is folded into a GLOBAL sum factorisation:
//
//   A[i,j] = sum_{Q0} W0[i0,j0](Q0) sum_{Q1} W1[i1,j1](Q1) sum_{Q2} W2[i2,j2](Q2) Ghat(Q0,Q1,Q2)
//
// with Q_d running over the quadrature points of ALL elements in the common support along d.
//
// Mapping.  A CTA owns a tile of T1 x T2 dofs in dimensions 1 and 2 and MARCHES along dimension 0
// over the element layers e0 that support its dof planes [r0, r1).  Per layer and chunk of QC
// points q0 it runs four barrier-separated stages through shared memory:
//   G   geometry at the halo points (T1+P)(P+1) x (T2+P)(P+1): trilinear J, adj(J), det ->
//       Ghat = w/|det| adj Kc adj^T (6 values) and w|det|                        [~115 flop/pt]
//   S1  contract Q2: lanes = (q0, Q1), warp-uniform i2 -> T1[Q1][9 terms][q0][pair2]
//   S2  contract Q1: lanes = (q0, pair2), warp-uniform i1 -> T2[q0][5 groups][pair1 x pair2]
//   S3  contract q0 into register accumulators acc[(P+1)^2] per dof pair and matrix.
// After a layer the entries with min(a0,b0)=0 are complete (dof e0 has no further support) and are
// stored; the other P^2 are carried into the next layer (register shift).  Flop count ~10 kFMA per
// element-equivalent for K+M at p=2 versus ~20 k for the per-element sum factorisation; the bound
// is the FP64 pipe (64 DFMA/clk/SM), then shared-memory bandwidth -- see DESIGN.md.
//
// Dimension names inside this file: 0 = marching (slowest dof index), 1, 2 = tile dimensions.

#include <cuda.h>

#include <algorithm>
#include <cmath>
#include <cstring>

#include "common.cuh"

namespace {

struct RowParams {
  BasisView B;
  QuadView Q;
  GeomView G;
  int plane_begin, plane_end;  // dof planes of dimension 0 written by this launch
  int nseg, tiles1, tiles2;
  int has_f, iso;
  double kc[6];  // symmetric conductivity (00,01,02,11,12,22)
  double rho, vcoef;
  // vector-valued spaces (ncomp > 1): one launch writes the rows of component `crow`; form e = column component e:
  //   stiffness-like:  dm[e][x][y] = D[crow][1+x][e][1+y]  (any 3x3 block, not necessarily symmetric)
  //   mass-like:       rhoe[e]     = D[crow][0][e][0]
  int ncomp, crow;
  int ecol0;  // column component of form 0 of a stiffness-like launch (forms = consecutive column components ecol0, ecol0 + 1, ..)
  double dm[3][9];
  double rhoe[3];
  // (value, derivative) of the local functions of the DOMINANT coefficient set of each dimension at the 1-D points,
  // [dim][q][a][2]: kernel parameters live in the constant bank, so these are free FP64 operands
  int cset[3];
  double ctab[3][B2_MAX_DEGREE + 1][B2_MAX_DEGREE + 1][2];
  // products of the dimension-2 factors of the dominant set, [q][a][b][va vb, va db, da vb, da db]: S1 multiplies the
  // coefficient straight into the accumulators (one DFMA with a constant-bank operand per term, no pre-scaling DMULs)
  double cp2[3][3][3][4];  // degrees 1 and 2 only (larger tables fall out of the constant cache)
  double* valK;
  double* valM;
  double* rhs;
  int g_ebeg;  // first element layer held by the precomputed geometry array (GPRE launches)
  int dbg;  // B2_EXPERIMENT builds: skip stages to measure their share of the critical path (1 stores, 2 S3, 4 S1, 8 S2)
  void* host_gcache;  // host only: GeomCache shared by the launches of one b2_assemble_rows_device call
};

template <int P_, int T1_, int T2_, int QC_, int NT_, int SPLIT_, int NG_ = 7, int OCC_ = 1>
struct RCfg {
  static constexpr int OCC = OCC_;  // CTAs per SM the configuration is sized for (registers, shared memory)
  static constexpr int NG = NG_;  // components of the coefficient per point: 7 = symmetric Ghat (6) + w|det|, 10 = general Ghat (9) + w|det|
  static constexpr int P = P_, NB = P + 1, NQ = P + 1, WD = 2 * P + 1;
  static constexpr int T1 = T1_, T2 = T2_, QC = QC_, NT = NT_, NW = NT / 32;
  // one CTA per SM; the register file is split over the 4 sub-partitions (16384 each), warps are dealt round-robin
  static constexpr int MAXREG = (16384 / (((NW + 3) / 4) * OCC) / 32 / 8 * 8) > 255 ? 255 : (16384 / (((NW + 3) / 4) * OCC) / 32 / 8 * 8);
  static constexpr int SPLIT = SPLIT_;                    // bit 0: S1 warp items split in three term groups, bit 1: S2 items in two (true = both)
  static constexpr int H1 = T1 + P, H2 = T2 + P;          // halo elements per tile dimension
  static constexpr int NQ1 = H1 * NQ, NQ2 = H2 * NQ;      // halo points
  static constexpr int NP1 = T1 * WD, NP2 = T2 * WD;      // dof pairs (i, i-P..i+P)
  static constexpr int LS = QC * NQ1;                     // S1 lane items (q0, Q1)
  // T1[Q1][term][q0][pair2] is stored by S1 lanes (q0, Q1) and loaded by S2 lanes (q0, pair2): the paddings NP2P (pairs per
  // q0) and T1QS (stride per Q1) are chosen so that both are free of bank conflicts (banks = 8-byte index mod 16 per half warp)
  static constexpr bool t1_ok(int np2p, int t1qs) {
    for (int h = 0; h * 16 < QC * NQ1; h++) {
      unsigned used = 0;
      for (int l = h * 16; l < h * 16 + 16 && l < QC * NQ1; l++) {
        const int b = ((l % NQ1) * t1qs + (l / NQ1) * np2p) % 16;
        if (used >> b & 1) return false;
        used |= 1u << b;
      }
    }
    return true;
  }
  static constexpr int pick_np2p() {
    for (int a = NP2; a < NP2 + 4; a++)
      for (int b = 9 * QC * a; b < 9 * QC * a + 16; b++)
        if (t1_ok(a, b)) return a;
    return NP2 | 1;
  }
  static constexpr int NP2P = pick_np2p();
  static constexpr int L2S = QC * NP2P;                   // S2 lane items (q0, pair2)
  static constexpr int pick_t1qs() {
    for (int b = 9 * L2S; b < 9 * L2S + 16; b++)
      if (t1_ok(NP2P, b)) return b;
    return (9 * L2S) | 1;
  }
  static constexpr int T1QS = pick_t1qs();                // T1 stride per Q1
  static constexpr int N12 = NP1 * NP2;                   // S3 items
  static constexpr int N12P = N12 + 4;
  static constexpr int pick_t2qs() {                      // T2 stride per q0: lane (q0, pair2) -> consecutive banks
    for (int b = 5 * N12P; b < 5 * N12P + 16; b++)
      if (b % 16 == NP2P % 16) return b;
    return 5 * N12P;
  }
  static constexpr int T2QS = pick_t2qs();
  static constexpr int IPT = (N12 + NT - 1) / NT;
  static constexpr int WPI1 = (LS + 31) / 32, WPI2 = (L2S + 31) / 32;
  // shared memory, in doubles
  static constexpr int SZ_G = NG * NQ2 * LS;
  static constexpr int SZ_T2 = QC * T2QS;
  static constexpr int OFF_G = 0, OFF_T2 = SZ_G;             // separate buffers: G(l+1) is produced while T2(l) is consumed
  static constexpr int OFF_T1 = SZ_G + SZ_T2, SZ_T1 = NQ1 * T1QS;
  static constexpr int OFF_L1 = OFF_T1 + SZ_T1, SZ_L1 = NQ1 * QC * T2;
  static constexpr int OFF_L2 = OFF_L1 + SZ_L1, SZ_L2 = QC * T1 * T2;
  static constexpr int OFF_NOD = OFF_L2 + SZ_L2, SZ_NOD = 3 * 2 * (H1 + 1) * (H2 + 1);   // x2 (double buffer)
  static constexpr int OFF_TB1 = OFF_NOD + 2 * SZ_NOD, SZ_TB1 = H1 * NQ * NB * 2;
  static constexpr int OFF_TB2 = OFF_TB1 + SZ_TB1, SZ_TB2 = H2 * NQ * NB * 2;
  static constexpr int OFF_TB0 = OFF_TB2 + SZ_TB2, SZ_TB0 = NQ * NB * 2;                 // x2 (double buffer)
  static constexpr int OFF_PW = OFF_TB0 + 2 * SZ_TB0, SZ_PW = 6 * NQ;
  static constexpr int OFF_ROW = OFF_PW + SZ_PW, SZ_ROW = 2 * NB;                        // ints [NB][lo, wid, cum, pad], x2
  static constexpr int MAXL = 512;                                                       // layers per marching segment (sSet0)
  static constexpr int OFF_SET = OFF_ROW + 2 * SZ_ROW;
  static constexpr int OFF_IC = OFF_SET + MAXL / 2 + 2;                                  // [IPT][NT] 64-bit row-start factors of the S3 items
  static constexpr int IPTS = (T1 * T2 * ((WD * WD + 1) / 2) + NT - 1) / NT;              // items per thread of the symmetric variant
  static constexpr int TOTAL = OFF_IC + (NG == 10 ? 2 * IPT : IPT > 2 * IPTS ? IPT : 2 * IPTS) * NT;  // NG == 10: off-diagonal blocks store every pair transposed as well
  static constexpr int NPF = (SZ_NOD + NT - 1) / NT;                                     // node values prefetched per thread
  static_assert(SZ_TB0 <= NT && 4 * NB <= NT, "layer tables are prefetched by one pass of the CTA");
};

// ---- S1: contract Q2.  One warp item = (dof i2 of the tile, 32 lanes of (q0, Q1), term group PART) ----
// terms: 0 A0=G00 v.v  1 A1=G01 v.v  2 A2=G11 v.v  3 B0=G02 v.d  4 B1=G12 v.d  5 C0=G02 d.v  6 C1=G12 d.v  7 D=G22 d.d  (a-side.b-side), NTK: M
template <class C, bool FK, bool FM, int PART, int NPARTS, bool CONST>
__device__ __forceinline__ void s1_item(const RowParams& prm, const double* __restrict__ sG, const double* __restrict__ sTb2, double* __restrict__ sT1, double* __restrict__ sL1,
                                        int i2, int i2l, int L, int n2) {
  constexpr int P = C::P, NB = C::NB, NQ = C::NQ, WD = C::WD, NQ1 = C::NQ1, NQ2 = C::NQ2, LS = C::LS;
  constexpr int NTK = FK ? 8 : 0, GS = NQ2 * LS;
  // term groups: TA = {A0,A1,A2} (v.v), TB = {B0,B1} (v.d) with the mass term and the load chain, TC = {C0,C1,D} (d.*)
  constexpr bool TA = FK && (NPARTS == 1 || PART == 0), TB = FK && (NPARTS == 1 || PART == 1), TC = FK && (NPARTS == 1 || PART == 2);
  constexpr bool TM = NPARTS == 1 || PART == 1;
  const int q0l = L / NQ1, Q1 = L % NQ1;
  double t[WD][9];
#pragma unroll
  for (int d = 0; d < WD; d++)
#pragma unroll
    for (int m = 0; m < 9; m++) t[d][m] = 0.;
  double l1 = 0.;
#pragma unroll
  for (int k = 0; k <= P; k++) {
    const int e2 = i2 - P + k;
    if (e2 < 0 || e2 >= n2) continue;
    const int a = P - k, e2l = i2l + k;
    // the point loop indexes no registers: at degree >= 3 it stays rolled -- fully unrolled the item is several thousand
    // instructions per term group and the kernel thrashes the instruction cache (ncu at p=4: 60 % of the S1 samples "no instruction")
#pragma unroll(P <= 2 ? NQ : 1)
    for (int q2 = 0; q2 < NQ; q2++) {
      const double* g = sG + (e2l * NQ + q2) * LS + L;
      const double* tb = sTb2 + (e2l * NQ + q2) * NB * 2;
      // component of Ghat[k][l]: symmetric storage (00,01,02,11,12,22) or general k*3+l
      constexpr bool SYM = C::NG == 7;
      constexpr int I00 = 0, I01 = 1, I02 = 2, I10 = SYM ? 1 : 3, I11 = SYM ? 3 : 4, I12 = SYM ? 4 : 5, I20 = SYM ? 2 : 6, I21 = SYM ? 4 : 7, I22 = SYM ? 5 : 8;
      if (CONST && P <= 2) {
        // 1-D factor products from the constant bank: every term is one DFMA
        double G00 = 0., G01 = 0., G11 = 0., G10 = 0., G02 = 0., G12 = 0., G20 = 0., G21 = 0., G22 = 0., Gw = 0.;
        if (TA) { G00 = g[I00 * GS]; G01 = g[I01 * GS]; G11 = g[I11 * GS]; if (!SYM) G10 = g[I10 * GS]; }
        if (TB) { G02 = g[I02 * GS]; G12 = g[I12 * GS]; }
        if (TC) { G20 = SYM && TB ? G02 : g[I20 * GS]; G21 = SYM && TB ? G12 : g[I21 * GS]; G22 = g[I22 * GS]; }
        if (TM) { Gw = g[(C::NG - 1) * GS]; l1 = fma(prm.ctab[2][q2][a][0], Gw, l1); }
#pragma unroll
        for (int b = 0; b <= P; b++) {
          const double* cp = prm.cp2[q2][a][b];
          const int d = b + k;
          if (TA) {
            t[d][0] = fma(G00, cp[0], t[d][0]);
            t[d][1] = fma(G01, cp[0], t[d][1]);
            t[d][2] = fma(G11, cp[0], t[d][2]);
            if (!SYM) t[d][8] = fma(G10, cp[0], t[d][8]);   // A10 takes the slot of the mass term (general forms carry no mass)
          }
          if (TB) {
            t[d][3] = fma(G02, cp[1], t[d][3]);
            t[d][4] = fma(G12, cp[1], t[d][4]);
          }
          if (TC) {
            t[d][5] = fma(G20, cp[2], t[d][5]);
            t[d][6] = fma(G21, cp[2], t[d][6]);
            t[d][7] = fma(G22, cp[3], t[d][7]);
          }
          if (TM && FM) t[d][8] = fma(Gw, cp[0], t[d][8]);
        }
      } else {
      const double va = CONST ? prm.ctab[2][q2][a][0] : tb[a * 2], da = CONST ? prm.ctab[2][q2][a][1] : tb[a * 2 + 1];
      double pa[9];
      double pw = 0., p10 = 0.;
      if (TA) { pa[0] = va * g[I00 * GS]; pa[1] = va * g[I01 * GS]; pa[2] = va * g[I11 * GS]; if (!SYM) p10 = va * g[I10 * GS]; }
      if (TB) { pa[3] = va * g[I02 * GS]; pa[4] = va * g[I12 * GS]; }
      if (TC) { pa[5] = da * g[I20 * GS]; pa[6] = da * g[I21 * GS]; pa[7] = da * g[I22 * GS]; }
      if (TM) { pw = va * g[(C::NG - 1) * GS]; l1 += pw; }
#pragma unroll
      for (int b = 0; b <= P; b++) {
        const double vb = CONST ? prm.ctab[2][q2][b][0] : tb[b * 2], db = CONST ? prm.ctab[2][q2][b][1] : tb[b * 2 + 1];
        const int d = b + k;
        if (TA) {
          t[d][0] = fma(pa[0], vb, t[d][0]);
          t[d][1] = fma(pa[1], vb, t[d][1]);
          t[d][2] = fma(pa[2], vb, t[d][2]);
          if (!SYM) t[d][8] = fma(p10, vb, t[d][8]);
        }
        if (TB) {
          t[d][3] = fma(pa[3], db, t[d][3]);
          t[d][4] = fma(pa[4], db, t[d][4]);
        }
        if (TC) {
          t[d][5] = fma(pa[5], vb, t[d][5]);
          t[d][6] = fma(pa[6], vb, t[d][6]);
          t[d][7] = fma(pa[7], db, t[d][7]);
        }
        if (TM && FM) t[d][8] = fma(pw, vb, t[d][8]);
      }
      }
    }
  }
  double* o = sT1 + Q1 * C::T1QS + q0l * C::NP2P + i2l * WD;
#pragma unroll
  for (int d = 0; d < WD; d++) {
    if (TA) { o[0 * C::L2S + d] = t[d][0]; o[1 * C::L2S + d] = t[d][1]; o[2 * C::L2S + d] = t[d][2]; if (C::NG == 10) o[8 * C::L2S + d] = t[d][8]; }
    if (TB) { o[3 * C::L2S + d] = t[d][3]; o[4 * C::L2S + d] = t[d][4]; }
    if (TC) { o[5 * C::L2S + d] = t[d][5]; o[6 * C::L2S + d] = t[d][6]; o[7 * C::L2S + d] = t[d][7]; }
    if (TM && FM) o[NTK * C::L2S + d] = t[d][8];
  }
  if (TM) sL1[i2l * LS + L] = l1;
}

// ---- S2: contract Q1.  One warp item = (dof i1 of the tile, 32 lanes of (q0, pair2), group set PART) ----
// groups by the dimension-0 factor still to be applied (a-side, b-side): 0 DD, 1 DV, 2 VD, 3 VV, NGK: M
template <class C, bool FK, bool FM, int PART, int NPARTS, bool CONST, bool SYM = false>
__device__ __forceinline__ void s2_item(const RowParams& prm, const double* __restrict__ sT1, const double* __restrict__ sTb1, const double* __restrict__ sL1, double* __restrict__ sT2,
                                        double* __restrict__ sL2, int i1, int i1l, int L, int n1, bool want_f) {
  constexpr int P = C::P, NB = C::NB, NQ = C::NQ, WD = C::WD, T1 = C::T1, T2 = C::T2, NP2 = C::NP2, L2S = C::L2S, N12P = C::N12P;
  constexpr int NTK = FK ? 8 : 0, NGK = FK ? 4 : 0;
  constexpr bool GA = FK && (NPARTS == 1 || PART == 0), GB = FK && (NPARTS == 1 || PART == 1), GM = FM && (NPARTS == 1 || PART == 1), GF = NPARTS == 1 || PART == 0;
  const int q0l = L / C::NP2P, pair2 = L % C::NP2P;
  const int i2l = pair2 / WD;
  const bool diag = GF && want_f && (pair2 % WD) == P;
  double u[WD][5];
#pragma unroll
  for (int d = 0; d < WD; d++)
#pragma unroll
    for (int m = 0; m < 5; m++) u[d][m] = 0.;
  double l2 = 0.;
#pragma unroll
  for (int k = 0; k <= P; k++) {
    const int e1 = i1 - P + k;
    if (e1 < 0 || e1 >= n1) continue;
    const int a = P - k, e1l = i1l + k;
#pragma unroll(P <= 2 ? NQ : 1)
    for (int q1 = 0; q1 < NQ; q1++) {
      const int Q1 = e1l * NQ + q1;
      const double* x = sT1 + Q1 * C::T1QS + L;
      const double* tb = sTb1 + Q1 * NB * 2;
      const double va = CONST ? prm.ctab[1][q1][a][0] : tb[a * 2], da = CONST ? prm.ctab[1][q1][a][1] : tb[a * 2 + 1];
      double X[7];
      if (GA) {
        const double A0 = x[0], A1 = x[L2S], B0 = x[3 * L2S], C0 = x[5 * L2S];
        const double A10 = C::NG == 10 ? x[8 * L2S] : A1;
        X[0] = va * A0; X[1] = va * A1; X[2] = va * B0; X[3] = fma(da, A10, va * C0);
      }
      if (GB) {
        const double A2 = x[2 * L2S], B1 = x[4 * L2S], C1 = x[6 * L2S], D = x[7 * L2S];
        X[4] = fma(da, A2, va * C1); X[5] = fma(da, B1, va * D);
      }
      if (GM) X[6] = va * x[NTK * L2S];
      if (diag) l2 = fma(va, sL1[i2l * C::LS + q0l * C::NQ1 + Q1], l2);
#pragma unroll
      for (int b = 0; b <= P; b++) {
        const double vb = CONST ? prm.ctab[1][q1][b][0] : tb[b * 2], db = CONST ? prm.ctab[1][q1][b][1] : tb[b * 2 + 1];
        const int d = b + k;
        if (SYM && d < P) continue;  // symmetric forms: the pairs (i1, j1 < i1) are produced by the tile of j1 and stored transposed
        if (GA) {
          u[d][0] = fma(X[0], vb, u[d][0]);
          u[d][1] = fma(X[1], db, fma(X[2], vb, u[d][1]));
          u[d][2] = fma(X[3], vb, u[d][2]);
        }
        if (GB) u[d][3] = fma(X[4], db, fma(X[5], vb, u[d][3]));
        if (GM) u[d][4] = fma(X[6], vb, u[d][4]);
      }
    }
  }
  double* o = sT2 + q0l * C::T2QS + (i1l * WD) * NP2 + pair2;
#pragma unroll
  for (int d = SYM ? P : 0; d < WD; d++) {
    if (GA) { o[0 * N12P + d * NP2] = u[d][0]; o[1 * N12P + d * NP2] = u[d][1]; o[2 * N12P + d * NP2] = u[d][2]; }
    if (GB) o[3 * N12P + d * NP2] = u[d][3];
    if (GM) o[NGK * N12P + d * NP2] = u[d][4];
  }
  if (diag) sL2[(q0l * T1 + i1l) * T2 + i2l] = l2;
}

// Out-of-line variants for halos that contain elements with a non-dominant coefficient set (domain boundary): the 1-D factors
// come from shared memory.  Kept out of the main instruction stream (the kernel is I-cache sensitive: ~6 k instructions).
template <class C, bool FK, bool FM, int PART, int NPARTS>
__device__ __noinline__ void s1_item_tab(const RowParams& prm, const double* sG, const double* sTb2, double* sT1, double* sL1, int i2, int i2l, int L, int n2) {
  s1_item<C, FK, FM, PART, NPARTS, false>(prm, sG, sTb2, sT1, sL1, i2, i2l, L, n2);
}
template <class C, bool FK, bool FM, int PART, int NPARTS, bool SYM = false>
__device__ __noinline__ void s2_item_tab(const RowParams& prm, const double* sT1, const double* sTb1, const double* sL1, double* sT2, double* sL2, int i1, int i1l, int L,
                                         int n1, bool want_f) {
  s2_item<C, FK, FM, PART, NPARTS, false, SYM>(prm, sT1, sTb1, sL1, sT2, sL2, i1, i1l, L, n1, want_f);
}

// 1/a for the normal, positive |det J| of a valid mesh: hardware seed (>= 20 bits) + two Newton steps (error < 2 ulp);
// replaces the IEEE division sequence (slow-path checks included) at every quadrature point
__device__ __forceinline__ double rcp_pos(double a) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
  double e = fma(-a, r, 1.);
  r = fma(r, e, r);
  e = fma(-a, r, 1.);
  return fma(r, e, r);
}

// ---- G: geometry of one (Q1, Q2) column of the halo, the QC points q0 = qc.. of the current chunk ----
// Trilinear map x = sum_v phi_v X_v of the element: the interpolations along dimensions 2 and 1 are shared by the points of the
// column (J[:,0] does not depend on xi0; J[:,1], J[:,2] are linear in xi0); adj(J), det and Ghat = w/|det| adj K adj^T per point.
template <class C, bool FK>
__device__ __forceinline__ void g_column(const RowParams& prm, int form, const double* __restrict__ sNod, const double* __restrict__ sPt, const double* __restrict__ sWt,
                                         double* __restrict__ sG, int col, int qc, int e1base, int e2base, int n1, int n2) {
  constexpr int NQ = C::NQ, QC = C::QC, H1 = C::H1, H2 = C::H2, NQ1 = C::NQ1, NQ2 = C::NQ2, LS = C::LS, GS = NQ2 * LS;
  const int Q2 = col / NQ1, Q1 = col % NQ1;
  const int e1l = Q1 / NQ, q1 = Q1 % NQ, e2l = Q2 / NQ, q2 = Q2 % NQ;
  const int e1 = e1base + e1l, e2 = e2base + e2l;
  if (e1 < 0 || e1 >= n1 || e2 < 0 || e2 >= n2) return;
  const double x1 = sPt[NQ + q1], x2 = sPt[2 * NQ + q2];
  const double w12 = sWt[NQ + q1] * sWt[2 * NQ + q2];
  double J0[3], D1a[3], D1d[3], D2a[3], D2d[3];  // J[:,0]; J[:,k](xi0) = Dka + xi0 Dkd
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const double* X = sNod + (i * 2 * (H1 + 1) + e1l) * (H2 + 1) + e2l;
    constexpr int SP = (H1 + 1) * (H2 + 1), S1 = H2 + 1;
    const double c000 = X[0], c001 = X[1], c010 = X[S1], c011 = X[S1 + 1];
    const double c100 = X[SP], c101 = X[SP + 1], c110 = X[SP + S1], c111 = X[SP + S1 + 1];
    const double d00 = c001 - c000, d01 = c011 - c010, d10 = c101 - c100, d11 = c111 - c110;                              // d/dxi2 on the 4 edges
    const double m00 = fma(x2, d00, c000), m01 = fma(x2, d01, c010), m10 = fma(x2, d10, c100), m11 = fma(x2, d11, c110);  // values at xi2
    const double g0 = fma(x1, d01 - d00, d00), g1 = fma(x1, d11 - d10, d10);                                              // d/dxi2 at (xi1, xi2), planes 0/1
    const double f0 = m01 - m00, f1 = m11 - m10;                                                                          // d/dxi1 at xi2, planes 0/1
    const double h0 = fma(x1, f0, m00), h1 = fma(x1, f1, m10);                                                            // values at (xi1, xi2)
    J0[i] = h1 - h0;
    D1a[i] = f0; D1d[i] = f1 - f0;
    D2a[i] = g0; D2d[i] = g1 - g0;
  }
  double* g = sG + Q2 * LS + Q1;
  // pass 1: A[0] = J[:,1] x J[:,2], det = J[:,0] . A[0] and the scale w/|det| of every point -- the reciprocal chains interleave
  double A0[QC][3], sc[QC], wd[QC];
#pragma unroll
  for (int q0 = 0; q0 < QC; q0++) {
    const double x0 = sPt[qc + q0];
    double J1[3], J2[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      J1[i] = fma(x0, D1d[i], D1a[i]);
      J2[i] = fma(x0, D2d[i], D2a[i]);
    }
    A0[q0][0] = J1[1] * J2[2] - J1[2] * J2[1];
    A0[q0][1] = J1[2] * J2[0] - J1[0] * J2[2];
    A0[q0][2] = J1[0] * J2[1] - J1[1] * J2[0];
  }
#pragma unroll
  for (int q0 = 0; q0 < QC; q0++) {
    const double det = J0[0] * A0[q0][0] + J0[1] * A0[q0][1] + J0[2] * A0[q0][2];
    const double adet = fabs(det), w = sWt[qc + q0] * w12;
    sc[q0] = w * rcp_pos(adet);
    wd[q0] = w * adet;
  }
  // pass 2: the other two adjugate rows A[1] = J[:,2] x J[:,0], A[2] = J[:,0] x J[:,1] and Ghat = s A K A^T
#pragma unroll
  for (int q0 = 0; q0 < QC; q0++) {
    const double x0 = sPt[qc + q0];
    double J1[3], J2[3], A[9];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      J1[i] = fma(x0, D1d[i], D1a[i]);
      J2[i] = fma(x0, D2d[i], D2a[i]);
      A[i] = A0[q0][i];
    }
    A[3] = J2[1] * J0[2] - J2[2] * J0[1]; A[4] = J2[2] * J0[0] - J2[0] * J0[2]; A[5] = J2[0] * J0[1] - J2[1] * J0[0];
    A[6] = J0[1] * J1[2] - J0[2] * J1[1]; A[7] = J0[2] * J1[0] - J0[0] * J1[2]; A[8] = J0[0] * J1[1] - J0[1] * J1[0];
    double* o = g + q0 * NQ1;
    if (FK && C::NG == 10) {
      // general coefficient of form `form`: Ghat[k][l] = s sum_xy A[k][x] M[x][y] A[l][y]
      const double s = sc[q0];
      const double* M = prm.dm[form];
      double T[9];
#pragma unroll
      for (int k = 0; k < 3; k++)
#pragma unroll
        for (int y = 0; y < 3; y++) T[k * 3 + y] = A[k * 3] * M[y] + A[k * 3 + 1] * M[3 + y] + A[k * 3 + 2] * M[6 + y];
#pragma unroll
      for (int k = 0; k < 3; k++)
#pragma unroll
        for (int l = 0; l < 3; l++) o[(k * 3 + l) * GS] = s * (T[k * 3] * A[l * 3] + T[k * 3 + 1] * A[l * 3 + 1] + T[k * 3 + 2] * A[l * 3 + 2]);
    } else if (FK) {
      const double s = sc[q0];
      if (prm.iso) {
        const double sk = s * prm.kc[0];
        o[0 * GS] = sk * (A[0] * A[0] + A[1] * A[1] + A[2] * A[2]);
        o[1 * GS] = sk * (A[0] * A[3] + A[1] * A[4] + A[2] * A[5]);
        o[2 * GS] = sk * (A[0] * A[6] + A[1] * A[7] + A[2] * A[8]);
        o[3 * GS] = sk * (A[3] * A[3] + A[4] * A[4] + A[5] * A[5]);
        o[4 * GS] = sk * (A[3] * A[6] + A[4] * A[7] + A[5] * A[8]);
        o[5 * GS] = sk * (A[6] * A[6] + A[7] * A[7] + A[8] * A[8]);
      } else {
        double T[9];
#pragma unroll
        for (int k = 0; k < 3; k++) {
          T[k * 3 + 0] = A[k * 3] * prm.kc[0] + A[k * 3 + 1] * prm.kc[1] + A[k * 3 + 2] * prm.kc[2];
          T[k * 3 + 1] = A[k * 3] * prm.kc[1] + A[k * 3 + 1] * prm.kc[3] + A[k * 3 + 2] * prm.kc[4];
          T[k * 3 + 2] = A[k * 3] * prm.kc[2] + A[k * 3 + 1] * prm.kc[4] + A[k * 3 + 2] * prm.kc[5];
        }
        int t = 0;
#pragma unroll
        for (int k = 0; k < 3; k++)
#pragma unroll
          for (int l = k; l < 3; l++) o[(t++) * GS] = s * (T[k * 3] * A[l * 3] + T[k * 3 + 1] * A[l * 3 + 1] + T[k * 3 + 2] * A[l * 3 + 2]);
      }
    }
    o[(C::NG - 1) * GS] = wd[q0];
  }
}

// ---- precomputed geometry (GPRE launches): Ghat and w|det J| of EVERY quadrature point are computed ONCE by k_geom3d into HBM,
// laid out [component][Q2][Q0][Q1] so that the halo box of a CTA and layer -- (T1+P)(P+1) x QC x (T2+P)(P+1) points x NG
// components -- is ONE TMA tensor copy (cp.async.bulk.tensor.4d) straight into sG, out-of-domain points zero-filled by the
// hardware.  The tile kernel then spends no FP64 instruction on geometry (the halo made it 2.25x redundant at tile 4x4: a third of
// all instructions); the price is HBM traffic (8 NG bytes per point written once, read ~(1+P/T)^2 times), which this FP64-bound
// kernel has to spare.
struct GeomParams {
  GeomView G;
  QuadView Q;
  int n0, n1, n2, ebeg, iso;
  double kc[6];
  int nform;        // GEN: number of general 3x3 coefficient blocks (column components of a vector-valued launch)
  double dm[3][9];  // GEN: Ghat^(f)[k][l] = w/|det| sum_xy adj[k][x] dm[f][x][y] adj[l][y]
  double* out;
  long long pitch1, sQ2, scomp;  // strides in doubles: Q0 -> pitch1, Q2 -> sQ2, component -> scomp
};

template <int P, bool GEN>
__global__ void __launch_bounds__(128) k_geom3d(const GeomParams prm) {
  constexpr int NQ = P + 1;
  const int Q1 = blockIdx.x * 128 + threadIdx.x, Q2 = blockIdx.y, e0 = prm.ebeg + blockIdx.z;
  if (Q1 >= NQ * prm.n1) return;
  const int e1 = Q1 / NQ, q1 = Q1 % NQ, e2 = Q2 / NQ, q2 = Q2 % NQ;
  const double x1 = prm.Q.x[1][q1], x2 = prm.Q.x[2][q2];
  const double w12 = prm.Q.w[1][q1] * prm.Q.w[2][q2];
  double J0[3], D1a[3], D1d[3], D2a[3], D2d[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const double* X = prm.G.nodes + i * prm.G.nnodes + (long long)e0 * prm.G.stride[0] + (long long)e1 * prm.G.stride[1] + e2;
    const long long SP = prm.G.stride[0], S1 = prm.G.stride[1];
    const double c000 = __ldg(X), c001 = __ldg(X + 1), c010 = __ldg(X + S1), c011 = __ldg(X + S1 + 1);
    const double c100 = __ldg(X + SP), c101 = __ldg(X + SP + 1), c110 = __ldg(X + SP + S1), c111 = __ldg(X + SP + S1 + 1);
    const double d00 = c001 - c000, d01 = c011 - c010, d10 = c101 - c100, d11 = c111 - c110;
    const double m00 = fma(x2, d00, c000), m01 = fma(x2, d01, c010), m10 = fma(x2, d10, c100), m11 = fma(x2, d11, c110);
    const double g0 = fma(x1, d01 - d00, d00), g1 = fma(x1, d11 - d10, d10);
    const double f0 = m01 - m00, f1 = m11 - m10;
    const double h0 = fma(x1, f0, m00), h1 = fma(x1, f1, m10);
    J0[i] = h1 - h0;
    D1a[i] = f0; D1d[i] = f1 - f0;
    D2a[i] = g0; D2d[i] = g1 - g0;
  }
  double* o = prm.out + (long long)Q2 * prm.sQ2 + (long long)blockIdx.z * NQ * prm.pitch1 + Q1;
#pragma unroll
  for (int q0 = 0; q0 < NQ; q0++) {
    const double x0 = prm.Q.x[0][q0];
    double J1[3], J2[3], A[9];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      J1[i] = fma(x0, D1d[i], D1a[i]);
      J2[i] = fma(x0, D2d[i], D2a[i]);
    }
    A[0] = J1[1] * J2[2] - J1[2] * J2[1]; A[1] = J1[2] * J2[0] - J1[0] * J2[2]; A[2] = J1[0] * J2[1] - J1[1] * J2[0];
    A[3] = J2[1] * J0[2] - J2[2] * J0[1]; A[4] = J2[2] * J0[0] - J2[0] * J0[2]; A[5] = J2[0] * J0[1] - J2[1] * J0[0];
    A[6] = J0[1] * J1[2] - J0[2] * J1[1]; A[7] = J0[2] * J1[0] - J0[0] * J1[2]; A[8] = J0[0] * J1[1] - J0[1] * J1[0];
    const double det = J0[0] * A[0] + J0[1] * A[1] + J0[2] * A[2];
    const double adet = fabs(det), w = prm.Q.w[0][q0] * w12;
    const double sc = w * rcp_pos(adet);
    double* oq = o + q0 * prm.pitch1;
    if (GEN) {
      // per form f: 9 components of the general coefficient + w |det J| (the layout [NG = 10] of the tile kernel's sG)
      for (int f = 0; f < prm.nform; f++) {
        const double* M = prm.dm[f];
        double T[9];
#pragma unroll
        for (int k = 0; k < 3; k++)
#pragma unroll
          for (int y = 0; y < 3; y++) T[k * 3 + y] = A[k * 3] * M[y] + A[k * 3 + 1] * M[3 + y] + A[k * 3 + 2] * M[6 + y];
#pragma unroll
        for (int k = 0; k < 3; k++)
#pragma unroll
          for (int l = 0; l < 3; l++) oq[(f * 10 + k * 3 + l) * prm.scomp] = sc * (T[k * 3] * A[l * 3] + T[k * 3 + 1] * A[l * 3 + 1] + T[k * 3 + 2] * A[l * 3 + 2]);
        oq[(f * 10 + 9) * prm.scomp] = w * adet;
      }
      continue;
    }
    if (prm.iso) {
      const double sk = sc * prm.kc[0];
      oq[0 * prm.scomp] = sk * (A[0] * A[0] + A[1] * A[1] + A[2] * A[2]);
      oq[1 * prm.scomp] = sk * (A[0] * A[3] + A[1] * A[4] + A[2] * A[5]);
      oq[2 * prm.scomp] = sk * (A[0] * A[6] + A[1] * A[7] + A[2] * A[8]);
      oq[3 * prm.scomp] = sk * (A[3] * A[3] + A[4] * A[4] + A[5] * A[5]);
      oq[4 * prm.scomp] = sk * (A[3] * A[6] + A[4] * A[7] + A[5] * A[8]);
      oq[5 * prm.scomp] = sk * (A[6] * A[6] + A[7] * A[7] + A[8] * A[8]);
    } else {
      double T[9];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        T[k * 3 + 0] = A[k * 3] * prm.kc[0] + A[k * 3 + 1] * prm.kc[1] + A[k * 3 + 2] * prm.kc[2];
        T[k * 3 + 1] = A[k * 3] * prm.kc[1] + A[k * 3 + 1] * prm.kc[3] + A[k * 3 + 2] * prm.kc[4];
        T[k * 3 + 2] = A[k * 3] * prm.kc[2] + A[k * 3 + 1] * prm.kc[4] + A[k * 3 + 2] * prm.kc[5];
      }
      int t = 0;
#pragma unroll
      for (int k = 0; k < 3; k++)
#pragma unroll
        for (int l = k; l < 3; l++) oq[(t++) * prm.scomp] = sc * (T[k * 3] * A[l * 3] + T[k * 3 + 1] * A[l * 3 + 1] + T[k * 3 + 2] * A[l * 3 + 2]);
    }
    oq[6 * prm.scomp] = w * adet;
  }
}

__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(unsigned dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst), "l"(map), "r"(c0),
               "r"(c1), "r"(c2), "r"(c3), "r"(bar)
               : "memory");
}

// The kernel.  Per element layer l two barrier-separated phases, software-pipelined so that every phase mixes FP64-heavy and
// shared-memory-heavy work of different layers and idle warps of one task pick up the other (dynamic warp-level work queue):
//   phase X:  S3(l-1) + store(l-1)  [static: the thread owns its dof pair's accumulators]   ||   S1(l)   [queue]
//   phase Y:  S2(l)  [queue]   ||   G(l+1)  [queue, 32 columns per item]
// NFORM > 1: vector-valued stiffness-like launch, forms = column components, one pipeline step per (layer, form, chunk)
// TRN (vector-valued spaces, off-diagonal block (crow, e > crow) of a SYMMETRIC form): every entry ((I, crow), (J, e)) is also stored
// transposed as ((J, e), (I, crow)), so the blocks below the diagonal are never integrated
template <class C, bool FK, bool FM, int NFORM, bool VEC, bool GPRE, bool SYM, bool TRN = false>
__global__ void __maxnreg__(C::MAXREG) k_rows3d(const RowParams prm, const __grid_constant__ CUtensorMap gmap) {
  static_assert(!SYM || (NFORM == 1 && C::NG == 7), "symmetric variant: one form with a symmetric coefficient (scalar space, or a diagonal block of a vector-valued one)");
  static_assert(!TRN || (VEC && !SYM && FK && !FM), "transposed stores: off-diagonal blocks of vector-valued stiffness-like forms");
  constexpr int NCV = VEC ? 3 : 1;  // components of a vector-valued space (the launcher admits 3 only)
  constexpr bool TWO = SYM || TRN;  // the thread stores its entries twice: it keeps the slot data of the transposed entry as well
  static_assert(!GPRE || C::NG == 7 || C::NG == 10, "precomputed geometry: symmetric scalar forms or general forms");
  static_assert(NFORM == 1 || (FK && !FM && C::NG == 10), "several forms per launch: general stiffness-like forms only");
  constexpr int P = C::P, NB = C::NB, NQ = C::NQ, WD = C::WD, QC = C::QC, NT = C::NT, NW = C::NW;
  constexpr int T1 = C::T1, T2 = C::T2, H1 = C::H1, H2 = C::H2, NQ1 = C::NQ1, NQ2 = C::NQ2;
  constexpr int NP2 = C::NP2, LS = C::LS, L2S = C::L2S;
  constexpr int N12 = C::N12, N12P = C::N12P;
  // SYM: a thread owns a dof pair {(i1, i2), (j1, j2)} with (j1, j2) >= (i1, i2) -- (WD^2 + 1) / 2 partners per tile dof -- and stores
  // its entries twice, as (i, j) and transposed as (j, i): half the S3 work and two thirds of the S2 work of the full variant
  constexpr int NSYM = (WD * WD + 1) / 2, N12S = T1 * T2 * NSYM;
  constexpr int IPT = SYM ? (N12S + NT - 1) / NT : C::IPT;
  static_assert(!SYM || IPT == C::IPTS, "the transposed row-start factors share the sIc array");
  static_assert(!TRN || C::NG == 10, "sIc is sized for the transposed entries of general-coefficient configurations");
  constexpr int NGK = FK ? 4 : 0;  // S2 output groups: DD DV VD VV [M]
  constexpr int NPARTS1 = ((C::SPLIT & 1) && FK) ? 3 : 1, NPARTS = ((C::SPLIT & 2) && FK) ? 2 : 1;  // term groups of S1 / S2 items
  static_assert(NQ % QC == 0, "the points q0 of a layer are processed in NQ/QC chunks");
  constexpr int NCH = NQ / QC, NSUB = NCH * NFORM;
  constexpr int NS1 = T2 * C::WPI1 * NPARTS1, NS2 = T1 * C::WPI2 * NPARTS, NGC = (NQ1 * NQ2 + 31) / 32;
  const BasisView& B = prm.B;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  extern __shared__ __align__(128) double smem[];
  double* sG = smem + C::OFF_G;       // [7][NQ2][q0 NQ1 + Q1]
  double* sT2 = smem + C::OFF_T2;     // [q0][group][N12P] with stride T2QS per q0
  double* sT1 = smem + C::OFF_T1;     // [NQ1][T1QS] = [NQ1][term][q0 NP2P + pair2]
  double* sL1 = smem + C::OFF_L1;     // [T2][QC][NQ1]
  double* sL2 = smem + C::OFF_L2;     // [QC][T1][T2]
  double* sNodB = smem + C::OFF_NOD;  // 2 x [3][2][H1+1][H2+1]
  double* sTb1 = smem + C::OFF_TB1;   // [H1][NQ][NB][2]  (value, derivative) of local function a at point q of halo element
  double* sTb2 = smem + C::OFF_TB2;   // [H2][NQ][NB][2]
  double* sTb0B = smem + C::OFF_TB0;  // 2 x [NQ][NB][2]  layer tables of dimension 0
  double* sPt = smem + C::OFF_PW;     // [3][NQ]
  double* sWt = sPt + 3 * NQ;         // [3][NQ]
  int* sRowB = reinterpret_cast<int*>(smem + C::OFF_ROW);  // 2 x [NB][4] = lo, wid, cum of dof e0+a along dimension 0
  int* sSet0 = reinterpret_cast<int*>(smem + C::OFF_SET);  // [MAXL] coefficient set of the layers of this segment
  int* sCnt = sSet0 + C::MAXL;                             // [2] work-queue heads of phases X and Y
  const unsigned gbar = smem_addr(sCnt + 2);               // mbarrier of the TMA loads of sG (GPRE)
  unsigned gphase = 0;

  // ---- work unit ----
  const int t2 = blockIdx.x % prm.tiles2, t1 = (blockIdx.x / prm.tiles2) % prm.tiles1, seg = blockIdx.x / (prm.tiles2 * prm.tiles1);
  const int i1lo = t1 * T1, i2lo = t2 * T2, e1base = i1lo - P, e2base = i2lo - P;
  const int n0 = B.nel[0], n1 = B.nel[1], n2 = B.nel[2], nd1 = B.ndofs[1], nd2 = B.ndofs[2];
  const long long npl = prm.plane_end - prm.plane_begin;
  const int r0 = prm.plane_begin + (int)(npl * seg / prm.nseg), r1 = prm.plane_begin + (int)(npl * (seg + 1) / prm.nseg);
  if (r0 >= r1) return;
  const int ebeg = max(0, r0 - P), eend = min(n0 - 1, r1 - 1);

  // ---- per-unit tables ----
  for (int t = tid; t < C::SZ_TB1; t += NT) {
    const int k = t & 1, a = (t >> 1) % NB, q = (t / (2 * NB)) % NQ, el = t / (2 * NB * NQ);
    const int e = e1base + el;
    sTb1[t] = (e >= 0 && e < n1) ? prm.Q.tab[1][((B.setidx[1][e] * 2 + k) * NB + a) * NQ + q] : 0.;
  }
  for (int t = tid; t < C::SZ_TB2; t += NT) {
    const int k = t & 1, a = (t >> 1) % NB, q = (t / (2 * NB)) % NQ, el = t / (2 * NB * NQ);
    const int e = e2base + el;
    sTb2[t] = (e >= 0 && e < n2) ? prm.Q.tab[2][((B.setidx[2][e] * 2 + k) * NB + a) * NQ + q] : 0.;
  }
  for (int t = tid; t < 3 * NQ; t += NT) {
    sPt[t] = prm.Q.x[t / NQ][t % NQ];
    sWt[t] = prm.Q.w[t / NQ][t % NQ];
  }
  for (int t = tid; t <= eend - ebeg; t += NT) sSet0[t] = B.setidx[0][ebeg + t];
  if (tid < 2) sCnt[tid] = NW;
  if (GPRE && tid == 0) mbar_init(gbar, 1);
  // do all elements of the halo carry the dominant coefficient set?  (then S1/S2 take their 1-D factors from the constant bank)
  bool mine = true;
  if (tid < H1) { const int e = e1base + tid; if (e >= 0 && e < n1 && B.setidx[1][e] != prm.cset[1]) mine = false; }
  const bool uni1 = __syncthreads_and(mine);
  mine = true;
  if (tid < H2) { const int e = e2base + tid; if (e >= 0 && e < n2 && B.setidx[2][e] != prm.cset[2]) mine = false; }
  const bool uni2 = __syncthreads_and(mine);

  // ---- layer tables, fetched ahead into registers and parked in a double buffer (index = layer parity):
  //      nodes of planes e, e+1 over the halo; 1-D table and CSR row data (lo, wid, cum) of dimension 0 ----
  constexpr int NPF = C::NPF;
  const double* nod_src[NPF];
  double pf_nod[NPF];
#pragma unroll
  for (int r = 0; r < NPF; r++) {
    const int t = tid + r * NT;
    nod_src[r] = nullptr;
    pf_nod[r] = 0.;
    if (t < C::SZ_NOD) {
      const int c2 = t % (H2 + 1), c1 = (t / (H2 + 1)) % (H1 + 1), pl = (t / ((H2 + 1) * (H1 + 1))) % 2, i = t / (2 * (H2 + 1) * (H1 + 1));
      const int v1 = min(max(e1base + c1, 0), n1), v2 = min(max(e2base + c2, 0), n2);
      nod_src[r] = prm.G.nodes + i * prm.G.nnodes + (long long)pl * prm.G.stride[0] + (long long)v1 * prm.G.stride[1] + v2;
    }
  }
  double pf_tb0 = 0.;
  int pf_row = 0;
  auto fetch_nodes = [&](int e) {
    if (e > eend) return;
#pragma unroll
    for (int r = 0; r < NPF; r++)
      if (nod_src[r]) pf_nod[r] = __ldg(nod_src[r] + (long long)e * prm.G.stride[0]);
  };
  auto park_nodes = [&](int e) {
    if (e > eend) return;
#pragma unroll
    for (int r = 0; r < NPF; r++)
      if (nod_src[r]) sNodB[(e & 1) * C::SZ_NOD + tid + r * NT] = pf_nod[r];
  };
  auto fetch_row = [&](int e) {
    if (e > eend) return;
    if (tid < C::SZ_TB0) {
      const int k = tid & 1, a = (tid >> 1) % NB, q = tid / (2 * NB);
      pf_tb0 = prm.Q.tab[0][((sSet0[e - ebeg] * 2 + k) * NB + a) * NQ + q];
    }
    if (tid < 4 * NB) {
      const int a = tid >> 2, w = tid & 3, i0 = e + a;
      pf_row = w == 0 ? B.lo[0][i0] : w == 1 ? B.wid[0][i0] : w == 2 ? B.cum[0][i0] : 0;
    }
  };
  auto park_row = [&](int e) {
    if (e > eend) return;
    if (tid < C::SZ_TB0) sTb0B[(e & 1) * C::SZ_TB0 + tid] = pf_tb0;
    if (tid < 4 * NB) sRowB[(e & 1) * 4 * NB + tid] = pf_row;
  };
  if (!GPRE) {
    fetch_nodes(ebeg);
    park_nodes(ebeg);
  }

  // ---- S3 items owned by this thread: dof pairs (i1, j1) x (i2, j2) ----
  // packed per item: wid1*wid2 | position of (j1, j2) in that box << 8 | tile-local (i1, i2) << 16 | valid << 24 | diagonal << 25;
  // the 64-bit row-start factor lives in shared memory (read once per layer), the rest is re-derived when needed
  int imeta[IPT];
  int imetaT[TWO ? IPT : 1];   // SYM: the same for the transposed entry (row (j1, j2), column (i1, i2)); bit 24 = it exists (off-diagonal pair)
  int t2idx[IPT];              // index of the pair in the T2 arrays
  constexpr bool REGBASE = SYM || NT <= 256;  // (512-thread ordered-pair configurations: the two extra registers per pair cost more in spills than the rebuilt base: p=3 13.5 -> 13.9 ms)
  long long ibase[REGBASE ? IPT : 1], ibaseT[TWO && REGBASE ? IPT : 1];  // thread part of the slot of an entry on interior layers (see the store)
  long long* sIc = reinterpret_cast<long long*>(smem + C::OFF_IC);  // [IPT][NT] (SYM: [2 IPT][NT], direct then transposed)
#pragma unroll
  for (int it = 0; it < IPT; it++) {
    const int item = tid + it * NT;
    int i1l, d1, i2l, d2;
    bool inrange;
    if (SYM) {
      const int t = item / NSYM, k = item % NSYM;
      i1l = t / T2; i2l = t % T2;
      if (k <= P) { d1 = P; d2 = P + k; }                                  // (0, 0..P)
      else { d1 = P + 1 + (k - P - 1) / WD; d2 = (k - P - 1) % WD; }        // (1..P, -P..P)
      inrange = item < N12S;
    } else {
      const int pair1 = item / NP2, pair2 = item % NP2;
      i1l = pair1 / WD; d1 = pair1 % WD; i2l = pair2 / WD; d2 = pair2 % WD;
      inrange = item < N12;
    }
    t2idx[it] = inrange ? (i1l * WD + d1) * NP2 + i2l * WD + d2 : 0;
    const int i1 = i1lo + i1l, j1 = i1 + d1 - P, i2 = i2lo + i2l, j2 = i2 + d2 - P;
    const bool v = inrange && i1 < nd1 && i2 < nd2 && j1 >= 0 && j1 < nd1 && j2 >= 0 && j2 < nd2;
    imeta[it] = (i1l * T2 + i2l) << 16;
    sIc[it * NT + tid] = 0;
    if (REGBASE) ibase[it] = 0;
    if (TWO) {
      imetaT[it] = 0;
      if (REGBASE) ibaseT[it] = 0;
      sIc[(IPT + it) * NT + tid] = 0;
    }
    if (v) {
      const int w1 = B.wid[1][i1], w2 = B.wid[2][i2];
      const long long ic = (long long)B.cum[1][i1] * B.W[2] + (long long)w1 * B.cum[2][i2];
      const int io = (j1 - B.lo[1][i1]) * w2 + (j2 - B.lo[2][i2]);
      sIc[it * NT + tid] = ic;
      if (REGBASE) ibase[it] = VEC ? ((long long)WD * ic * NCV + (long long)prm.crow * WD * (w1 * w2)) * NCV + io * NCV + prm.ecol0 : (long long)WD * ic + io;
      imeta[it] |= (w1 * w2) | io << 8 | 1 << 24 | (d1 == P && d2 == P ? 1 << 25 : 0);
      // SYM: the pair {i, i} holds both orders of its dimension-0 entries itself; TRN: its transposed entries lie in another block
      if (TRN || (SYM && !(d1 == P && d2 == P))) {
        const int v1 = B.wid[1][j1], v2 = B.wid[2][j2];
        const long long icT = (long long)B.cum[1][j1] * B.W[2] + (long long)v1 * B.cum[2][j2];
        const int ioT = (i1 - B.lo[1][j1]) * v2 + (i2 - B.lo[2][j2]);
        sIc[(IPT + it) * NT + tid] = icT;
        if (REGBASE) ibaseT[it] = VEC ? ((long long)WD * icT * NCV + (long long)prm.ecol0 * WD * (v1 * v2)) * NCV + ioT * NCV + prm.crow : (long long)WD * icT + ioT;
        imetaT[it] = (v1 * v2) | ioT << 8 | 1 << 24;
      }
    }
  }
  const long long W12 = B.W[1] * B.W[2], W12N = W12 * (NCV * NCV), WDW12 = WD * W12N;

  double accK[IPT][NFORM][NB][NB], accM[IPT][NB][NB], accF[IPT][NB];
#pragma unroll
  for (int it = 0; it < IPT; it++)
#pragma unroll
    for (int a = 0; a < NB; a++) {
      accF[it][a] = 0.;
#pragma unroll
      for (int b = 0; b < NB; b++) {
        accM[it][a][b] = 0.;
#pragma unroll
        for (int f = 0; f < NFORM; f++) accK[it][f][a][b] = 0.;
      }
    }

  __syncthreads();

  // one pipeline step = one chunk of QC point-planes q0 of one form of one element layer (order: layer, form, chunk)
  // (the loop starts at s = -1 with phase Y only: that is the prologue G(0), through the one call site of g_column)
  const int nstep = (eend - ebeg + 1) * NSUB;
  for (int s = -1; s <= nstep; s++) {
    if (s >= 0) {
    const int l = ebeg + s / NSUB, ch = s % NSUB;
    // ======================= phase X: S3(s-1) [+ store of its layer]  ||  S1(s) =======================
    if (!GPRE && ch == NSUB - 1) fetch_nodes(l + 1);
    if (ch == 0) fetch_row(l);
    if (tid == 0) sCnt[1] = NW;
    if (s > 0) {
      const int e0 = ebeg + (s - 1) / NSUB, qc = ((s - 1) % NCH) * QC, form = ((s - 1) % NSUB) / NCH;
      const double* sTb0 = sTb0B + (e0 & 1) * C::SZ_TB0;
      const int* sRow = sRowB + (e0 & 1) * 4 * NB;
      // the S2 results of the whole chunk are fetched before the first use (one exposed shared-memory latency per step, not QC)
      constexpr bool PRE3 = IPT == 1 && P <= 2 && !VEC && NFORM == 1 && NT <= 256;  // register-rich configurations only
      double xpre[PRE3 ? QC : 1][5];
      if (PRE3 && (SYM ? (imeta[0] >> 24 & 1) : (tid < N12))) {
#pragma unroll
        for (int q0l = 0; q0l < QC; q0l++) {
          const double* x = sT2 + q0l * C::T2QS + t2idx[0];
          if (FK) { xpre[q0l][0] = x[0]; xpre[q0l][1] = x[N12P]; xpre[q0l][2] = x[2 * N12P]; xpre[q0l][3] = x[3 * N12P]; }
          if (FM) xpre[q0l][4] = x[NGK * N12P];
        }
      }
#pragma unroll
      for (int q0l = 0; q0l < QC; q0l++) {
        double va[NB], da[NB];
#pragma unroll
        for (int a = 0; a < NB; a++) {
          va[a] = sTb0[((qc + q0l) * NB + a) * 2];
          da[a] = sTb0[((qc + q0l) * NB + a) * 2 + 1];
        }
#pragma unroll
        for (int it = 0; it < IPT; it++) {
#ifdef B2_EXPERIMENT
          if (prm.dbg & 2) continue;
#endif
          if (SYM ? (imeta[it] >> 24 & 1) : (tid + it * NT < N12)) {
            const double* x = sT2 + q0l * C::T2QS + t2idx[it];
            double gDD = 0., gDV = 0., gVD = 0., gVV = 0., gM = 0.;
            if (PRE3) {
              if (FK) { gDD = xpre[q0l][0]; gDV = xpre[q0l][1]; gVD = xpre[q0l][2]; gVV = xpre[q0l][3]; }
              if (FM) gM = xpre[q0l][4];
            } else {
              if (FK) { gDD = x[0]; gDV = x[N12P]; gVD = x[2 * N12P]; gVV = x[3 * N12P]; }
              if (FM) gM = x[NGK * N12P];
            }
#pragma unroll
            for (int b = 0; b < NB; b++) {
              const double ud = fma(da[b], gDD, va[b] * gDV), uv = fma(da[b], gVD, va[b] * gVV), um = va[b] * gM;
#pragma unroll
              for (int a = 0; a < NB; a++) {
                if (FK) {
#pragma unroll
                  for (int f = 0; f < NFORM; f++)
                    if (f == form) accK[it][f][a][b] = fma(da[a], ud, fma(va[a], uv, accK[it][f][a][b]));
                }
                if (FM) accM[it][a][b] = fma(va[a], um, accM[it][a][b]);
              }
            }
            if (prm.has_f && form == 0 && (imeta[it] >> 25 & 1)) {
              const double l2 = sL2[q0l * T1 * T2 + (imeta[it] >> 16 & 255)];
#pragma unroll
              for (int a = 0; a < NB; a++) accF[it][a] = fma(va[a], l2, accF[it][a]);
            }
          }
        }
      }
      if ((s - 1) % NSUB == NSUB - 1) {
      // store the completed entries of layer e0, carry the rest
      const bool last = e0 == n0 - 1;
      // interior layer of a scalar space (the common case): all P+1 rows have the full width 2P+1, their first coupled dof is
      // i0-P and they all lie in the planes of this launch -- the slot arithmetic collapses to one 64-bit base per row
      const bool interior = (!VEC || (REGBASE && !FM)) && e0 >= P && e0 <= n0 - 1 - P && e0 >= r0 && e0 + P < r1;
#pragma unroll
      for (int it = 0; it < IPT; it++) {
#ifdef B2_EXPERIMENT
        if (prm.dbg & 1) {
        } else
#endif
        if (interior) {
          // rows e0 .. e0+P have the full width WD, so cum0[e0 + a] = cum0[e0] + WD a: the slot of entry (a, b) is
          // cum0[e0] W12 [layer, CTA-uniform] + ibase [thread, fixed] + a WD W12 [uniform] + (P - a + b) w12 [thread, small]
          if (!REGBASE) {
            // register-starved configurations (512 threads): the thread part of the slot comes from shared memory
            if (imeta[it] >> 24 & 1) {
              const int iw12 = imeta[it] & 255;
              const long long b12 = (long long)WD * sIc[it * NT + tid] + (imeta[it] >> 8 & 255);
#pragma unroll
              for (int a = 0; a < NB; a++) {
                const long long ra = (long long)sRow[a * 4 + 2] * W12 + b12;
#pragma unroll
                for (int b = 0; b < NB; b++) {
                  if (a == 0 || b == 0) {
                    const long long slot = ra + (P - a + b) * iw12;
                    if (FK) prm.valK[slot] = accK[it][0][a][b];
                    if (FM && prm.valM) prm.valM[slot] = accM[it][a][b] * prm.rho;
                  }
                }
              }
              if (prm.has_f && (imeta[it] >> 25 & 1)) {
                const int il = imeta[it] >> 16 & 255;
                prm.rhs[((long long)e0 * nd1 + i1lo + il / T2) * nd2 + i2lo + il % T2] = accF[it][0] * prm.vcoef;
              }
            }
          } else if (imeta[it] >> 24 & 1) {
            // (vector-valued spaces: W12 and w12 carry the factors ncomp^2 and ncomp of the component-interleaved slots, the
            // component offsets are part of ibase; form f = column component ecol0 + f)
            const int iw12 = (imeta[it] & 255) * NCV;
            const long long lbase = (long long)sRow[2] * W12N;
            double* __restrict__ pK = prm.valK + (lbase + ibase[REGBASE ? it : 0]);
            double* __restrict__ pM = prm.valM + (lbase + ibase[REGBASE ? it : 0]);
#pragma unroll
            for (int a = 0; a < NB; a++) {
#pragma unroll
              for (int b = 0; b < NB; b++) {
                if (a == 0 || b == 0) {
                  const long long off = (long long)a * WDW12 + (P - a + b) * iw12;
                  if (FK) {
#pragma unroll
                    for (int f = 0; f < NFORM; f++) pK[off + f] = accK[it][f][a][b];
                  }
                  if (FM && !VEC && prm.valM) pM[off] = accM[it][a][b] * prm.rho;
                }
              }
            }
            if (prm.has_f && (imeta[it] >> 25 & 1)) {
              const int il = imeta[it] >> 16 & 255;
              prm.rhs[(((long long)e0 * nd1 + i1lo + il / T2) * nd2 + i2lo + il % T2) * NCV + (VEC ? prm.crow : 0)] = accF[it][0] * prm.vcoef;
            }
            if (TWO && (imetaT[it] >> 24 & 1)) {
              // the transposed entries: row (e0 + b, j1, j2) [component ecol0 + f], column (e0 + a, i1, i2) [component crow]
              const int iwT = (imetaT[it] & 255) * NCV;
              double* __restrict__ qK = prm.valK + (lbase + ibaseT[REGBASE ? it : 0]);
              double* __restrict__ qM = prm.valM + (lbase + ibaseT[REGBASE ? it : 0]);
#pragma unroll
              for (int b = 0; b < NB; b++) {
#pragma unroll
                for (int a = 0; a < NB; a++) {
                  if (a == 0 || b == 0) {
                    const long long off = (long long)b * WDW12 + (P - b + a) * iwT;
                    if (FK) {
#pragma unroll
                      for (int f = 0; f < NFORM; f++) qK[off + (long long)f * WD * iwT] = accK[it][f][a][b];
                    }
                    if (FM && !VEC && prm.valM) qM[off] = accM[it][a][b] * prm.rho;
                  }
                }
              }
            }
          }
        } else if (imeta[it] >> 24 & 1) {
          const long long ic12 = sIc[it * NT + tid];
          const int iw12 = imeta[it] & 255, io12 = imeta[it] >> 8 & 255;
#pragma unroll
          for (int a = 0; a < NB; a++) {
            const int i0 = e0 + a;
            if (i0 < r0 || i0 >= r1) continue;
            const int rlo = sRow[a * 4], rwid = sRow[a * 4 + 1], rcum = sRow[a * 4 + 2];
            // scalar space: slot = R + pos;  ncomp components: row (I, crow) starts at (R ncomp + crow w) ncomp and holds
            // (J, e) at pos ncomp + e, with R = first slot of basis row I, w = its width, pos = position of J in the row
            const int nc = VEC ? prm.ncomp : 1;
            const long long rowslot = (((long long)rcum * W12 + (long long)rwid * ic12) * nc + (long long)prm.crow * rwid * iw12) * nc + (long long)io12 * nc;
#pragma unroll
            for (int b = 0; b < NB; b++) {
              if (a == 0 || b == 0 || last) {
                const long long slot = rowslot + (long long)(e0 + b - rlo) * iw12 * nc;
                if (FK) {
#pragma unroll
                  for (int f = 0; f < NFORM; f++) prm.valK[slot + (VEC ? prm.ecol0 : 0) + f] = accK[it][f][a][b];
                }
                if (FM && prm.valM) {
                  if (nc == 1) prm.valM[slot] = accM[it][a][b] * prm.rho;
                  else
                    for (int e = 0; e < nc; e++) prm.valM[slot + e] = accM[it][a][b] * prm.rhoe[e];
                }
              }
            }
            if (prm.has_f && (imeta[it] >> 25 & 1) && (a == 0 || last)) {
              const int il = imeta[it] >> 16 & 255;
              prm.rhs[(((long long)i0 * nd1 + i1lo + il / T2) * nd2 + i2lo + il % T2) * nc + prm.crow] = accF[it][a] * prm.vcoef;
            }
          }
          if (TWO && (imetaT[it] >> 24 & 1)) {
            // transposed entries: row (e0 + b, j1, j2) [component ecol0 + f], column (e0 + a, i1, i2) [component crow]
            const long long icT = sIc[(IPT + it) * NT + tid];
            const int iwT = imetaT[it] & 255, ioT = imetaT[it] >> 8 & 255;
            const int nc = VEC ? prm.ncomp : 1;
#pragma unroll
            for (int b = 0; b < NB; b++) {
              const int j0 = e0 + b;
              if (j0 < r0 || j0 >= r1) continue;
              const int rlo = sRow[b * 4], rwid = sRow[b * 4 + 1], rcum = sRow[b * 4 + 2];
#pragma unroll
              for (int f = 0; f < NFORM; f++) {
                const long long rowslot = VEC ? (((long long)rcum * W12 + (long long)rwid * icT) * nc + (long long)(prm.ecol0 + f) * rwid * iwT) * nc + (long long)ioT * nc + prm.crow
                                              : (long long)rcum * W12 + (long long)rwid * icT + ioT;
#pragma unroll
                for (int a = 0; a < NB; a++) {
                  if (a == 0 || b == 0 || last) {
                    const long long slot = rowslot + (long long)(e0 + a - rlo) * iwT * nc;
                    if (FK) prm.valK[slot] = accK[it][f][a][b];
                    if (FM && !VEC && prm.valM) prm.valM[slot] = accM[it][a][b] * prm.rho;
                  }
                }
              }
            }
          }
        }
#pragma unroll
        for (int a = 0; a < NB; a++) {
          accF[it][a] = a < P ? accF[it][a + 1] : 0.;
#pragma unroll
          for (int b = 0; b < NB; b++) {
            if (FK) {
#pragma unroll
              for (int f = 0; f < NFORM; f++) accK[it][f][a][b] = (a < P && b < P) ? accK[it][f][a + 1][b + 1] : 0.;
            }
            if (FM) accM[it][a][b] = (a < P && b < P) ? accM[it][a + 1][b + 1] : 0.;
          }
        }
      }
      }
    }
    if (s < nstep) {
      if (GPRE
#ifdef B2_EXPERIMENT
          && !(prm.dbg & 16)
#endif
      ) {  // the box of this step has landed in sG
        mbar_wait(gbar, gphase);
        gphase ^= 1;
      }
      // work queue: the first item of a warp is static, the rest is handed out by a shared counter (starts at NW)
      for (int wi = warp; wi < NS1;) {
        const int part = wi % NPARTS1, wj = wi / NPARTS1;
        const int i2l = wj / C::WPI1, L = (wj % C::WPI1) * 32 + lane, i2 = i2lo + i2l;
        const int e1 = e1base + (L % NQ1) / NQ;
#ifdef B2_EXPERIMENT
        if (prm.dbg & 4) {
        } else
#endif
        if (i2 < nd2 && L < LS && e1 >= 0 && e1 < n1) {
          if (uni2) {
            if (NPARTS1 == 1) s1_item<C, FK, FM, 0, 1, true>(prm, sG, sTb2, sT1, sL1, i2, i2l, L, n2);
            else if (part == 0) s1_item<C, FK, FM, 0, 3, true>(prm, sG, sTb2, sT1, sL1, i2, i2l, L, n2);
            else if (part == 1) s1_item<C, FK, FM, 1, 3, true>(prm, sG, sTb2, sT1, sL1, i2, i2l, L, n2);
            else s1_item<C, FK, FM, 2, 3, true>(prm, sG, sTb2, sT1, sL1, i2, i2l, L, n2);
          } else {
            if (NPARTS1 == 1) s1_item_tab<C, FK, FM, 0, 1>(prm, sG, sTb2, sT1, sL1, i2, i2l, L, n2);
            else if (part == 0) s1_item_tab<C, FK, FM, 0, 3>(prm, sG, sTb2, sT1, sL1, i2, i2l, L, n2);
            else if (part == 1) s1_item_tab<C, FK, FM, 1, 3>(prm, sG, sTb2, sT1, sL1, i2, i2l, L, n2);
            else s1_item_tab<C, FK, FM, 2, 3>(prm, sG, sTb2, sT1, sL1, i2, i2l, L, n2);
          }
        }
        if (P <= 2 && NS1 <= NW) break;  // every warp has its one static item: no queue traffic
        __syncwarp();
        if (lane == 0) wi = atomicAdd(&sCnt[0], 1);
        wi = __shfl_sync(0xffffffffu, wi, 0);
      }
    }
    if (!GPRE && ch == NSUB - 1) park_nodes(l + 1);
    if (ch == 0) park_row(l);
    __syncthreads();
    if (s >= nstep) break;
    }

    // ======================= phase Y: S2(s)  ||  G(s+1) =======================
    if (tid == 0) sCnt[0] = NW;
    const int ngc = (!GPRE && s + 1 < nstep) ? NGC : 0;
    const int ln = ebeg + (s + 1) / NSUB, qcn = ((s + 1) % NCH) * QC, formn = ((s + 1) % NSUB) / NCH;
    if (GPRE && s + 1 < nstep && tid == 0
#ifdef B2_EXPERIMENT
        && !(prm.dbg & 16)
#endif
    ) {
      // sG was last read by S1 of step s (phase X, generic proxy): order those reads before the async-proxy writes
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(gbar, (unsigned)(sizeof(double) * C::SZ_G));
      tma_load_4d(smem_addr(sG), &gmap, NQ * e1base, (ln - prm.g_ebeg) * NQ + qcn, NQ * e2base, formn * C::NG, gbar);
    }
    const int ns2 = s >= 0 ? NS2 : 0;
    for (int wi = warp; wi < ns2 + ngc;) {
      if (wi >= ns2) {
        const int col = (wi - ns2) * 32 + lane;
        if (col < NQ1 * NQ2) g_column<C, FK>(prm, formn, sNodB + (ln & 1) * C::SZ_NOD, sPt, sWt, sG, col, qcn, e1base, e2base, n1, n2);
      } else {
        const int part = wi % NPARTS, wj = wi / NPARTS;
        const int i1l = wj / C::WPI2, L = (wj % C::WPI2) * 32 + lane, i1 = i1lo + i1l;
#ifdef B2_EXPERIMENT
        if (prm.dbg & 8) {
        } else
#endif
        if (i1 < nd1 && L < L2S && (L % C::NP2P) < NP2) {
          if (uni1) {
            if (NPARTS == 1) s2_item<C, FK, FM, 0, 1, true, SYM>(prm, sT1, sTb1, sL1, sT2, sL2, i1, i1l, L, n1, prm.has_f);
            else if (part == 0) s2_item<C, FK, FM, 0, 2, true, SYM>(prm, sT1, sTb1, sL1, sT2, sL2, i1, i1l, L, n1, prm.has_f);
            else s2_item<C, FK, FM, 1, 2, true, SYM>(prm, sT1, sTb1, sL1, sT2, sL2, i1, i1l, L, n1, prm.has_f);
          } else {
            if (NPARTS == 1) s2_item_tab<C, FK, FM, 0, 1, SYM>(prm, sT1, sTb1, sL1, sT2, sL2, i1, i1l, L, n1, prm.has_f);
            else if (part == 0) s2_item_tab<C, FK, FM, 0, 2, SYM>(prm, sT1, sTb1, sL1, sT2, sL2, i1, i1l, L, n1, prm.has_f);
            else s2_item_tab<C, FK, FM, 1, 2, SYM>(prm, sT1, sTb1, sL1, sT2, sL2, i1, i1l, L, n1, prm.has_f);
          }
        }
      }
      if (P <= 2 && GPRE && NS2 <= NW) break;  // one static S2 item per warp and no geometry items
      __syncwarp();
      if (lane == 0) wi = atomicAdd(&sCnt[1], 1);
      wi = __shfl_sync(0xffffffffu, wi, 0);
    }
    __syncthreads();
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// driver entry point without linking libcuda (the library links the static runtime only)
EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
  }
  return fn;
}

// the geometry array of one assembly call and its TMA descriptor (the K and M launches of degrees 3 and 4 share it)
struct GeomCache {
  bool valid = false;
  int box[4] = {0, 0, 0, 0};
  int g_ebeg = 0;
  double kc[6] = {0, 0, 0, 0, 0, 0};  // the conductivity the array was computed with
  CUtensorMap map;
};

// GPRE launches: run k_geom3d over the element layers the launch touches and describe the result to the TMA unit
template <class C, int NFORM>
int prepare_geometry(b2_ctx* ctx, RowParams& prm, CUtensorMap* map) {
  constexpr int P = C::P, NQ = C::NQ;
  constexpr bool GEN = C::NG == 10;
  constexpr int NCOMP = C::NG * NFORM;
  GeomCache* gc = (GeomCache*)prm.host_gcache;
  if (!GEN && gc && gc->valid && gc->box[0] == C::NQ1 && gc->box[1] == C::QC && gc->box[2] == C::NQ2 && gc->box[3] == C::NG && !memcmp(gc->kc, prm.kc, sizeof(prm.kc))) {
    *map = gc->map;
    prm.g_ebeg = gc->g_ebeg;
    return B2_OK;
  }
  const int n0 = prm.B.nel[0], n1 = prm.B.nel[1], n2 = prm.B.nel[2];
  const int ebeg = std::max(0, prm.plane_begin - P), eend = std::min(n0 - 1, prm.plane_end - 1);
  const long long nlay = eend - ebeg + 1;
  if (nlay <= 0) return B2_OK;
  EncodeTiledFn encode = get_encode_tiled();
  if (!encode) return b2_fail(ctx, B2_ECUDA, "cuTensorMapEncodeTiled is not available");
  const long long pitch1 = ((long long)NQ * n1 + 1) & ~1LL;  // rows of Q1 start on 16-byte boundaries
  const long long sQ2 = nlay * NQ * pitch1, scomp = (long long)NQ * n2 * sQ2;
  const size_t need = sizeof(double) * (size_t)(NCOMP * scomp);
  if (ctx->gbuf_bytes < need) {
    if (ctx->gbuf) cudaFree(ctx->gbuf);
    ctx->gbuf = nullptr;
    ctx->gbuf_bytes = 0;
    if (cudaMalloc(&ctx->gbuf, need) != cudaSuccess) {
      cudaGetLastError();
      ctx->gbuf = nullptr;
      return B2_ENOMEM;
    }
    ctx->gbuf_bytes = need;
  }
  GeomParams gp;
  gp.G = prm.G;
  gp.Q = prm.Q;
  gp.n0 = n0; gp.n1 = n1; gp.n2 = n2;
  gp.ebeg = ebeg;
  gp.iso = prm.iso;
  for (int t = 0; t < 6; t++) gp.kc[t] = prm.kc[t];
  gp.nform = NFORM;
  for (int f = 0; f < 3; f++)
    for (int t = 0; t < 9; t++) gp.dm[f][t] = prm.dm[f][t];
  gp.out = (double*)ctx->gbuf;
  gp.pitch1 = pitch1; gp.sQ2 = sQ2; gp.scomp = scomp;
  if ((long long)NQ * n2 > 65535 || nlay > 65535) return B2_EUNSUPPORTED;
  {
    KernelTimer timer(ctx);
    k_geom3d<P, GEN><<<dim3((unsigned)((NQ * n1 + 127) / 128), (unsigned)(NQ * n2), (unsigned)nlay), 128, 0, ctx->stream>>>(gp);
  }
  ctx->launches++;
  {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
      return b2_fail(ctx, B2_ECUDA, std::string("k_geom3d launch failed: ") + cudaGetErrorString(e) + " grid " + std::to_string((NQ * n1 + 127) / 128) + "x" + std::to_string(NQ * n2) + "x" +
                                        std::to_string(nlay) + " ebeg " + std::to_string(ebeg) + " planes " + std::to_string(prm.plane_begin) + ":" + std::to_string(prm.plane_end));
  }
  prm.g_ebeg = ebeg;
  const cuuint64_t gdim[4] = {(cuuint64_t)NQ * n1, (cuuint64_t)(nlay * NQ), (cuuint64_t)NQ * n2, (cuuint64_t)NCOMP};
  const cuuint64_t gstr[3] = {(cuuint64_t)pitch1 * 8, (cuuint64_t)sQ2 * 8, (cuuint64_t)scomp * 8};
  const cuuint32_t box[4] = {(cuuint32_t)C::NQ1, (cuuint32_t)C::QC, (cuuint32_t)C::NQ2, (cuuint32_t)C::NG};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, ctx->gbuf, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return b2_fail(ctx, B2_ECUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
  if (gc && !GEN) {
    gc->valid = true;
    gc->box[0] = C::NQ1; gc->box[1] = C::QC; gc->box[2] = C::NQ2; gc->box[3] = C::NG;
    gc->g_ebeg = ebeg;
    memcpy(gc->kc, prm.kc, sizeof(prm.kc));
    gc->map = *map;
  }
  return B2_OK;
}

template <class C, bool FK, bool FM, int NFORM = 1, bool VEC = false, bool GPRE = false, bool SYM = false, bool TRN = false>
int launch_rows_cfg(b2_ctx* ctx, RowParams& prm) {
  auto kern = k_rows3d<C, FK, FM, NFORM, VEC, GPRE, SYM, TRN>;
  const size_t smem = sizeof(double) * C::TOTAL;
  static_assert(sizeof(double) * C::TOTAL <= 227 * 1024, "tile does not fit in shared memory");
  B2_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUtensorMap gmap;
  memset(&gmap, 0, sizeof(gmap));
  if (GPRE) {
    const int rc = prepare_geometry<C, NFORM>(ctx, prm, &gmap);
    if (rc != B2_OK) return rc;
  }
  prm.tiles1 = (prm.B.ndofs[1] + C::T1 - 1) / C::T1;
  prm.tiles2 = (prm.B.ndofs[2] + C::T2 - 1) / C::T2;
  // segments along the marching direction: trade redundant layers (P per segment) against wave quantisation
  const long long tiles = (long long)prm.tiles1 * prm.tiles2, npl = prm.plane_end - prm.plane_begin;
  const int slots = ctx->sm_count * C::OCC;  // CTAs resident at a time
  int best = 1;
  double bestcost = 1e300;
  const int64_t forced = ctx->opts.count("rows_nseg") ? ctx->opts["rows_nseg"] : 0;
  for (int s = 1; s <= 64 && s <= npl; s++) {
    if ((npl + s - 1) / s + C::P > C::MAXL) continue;  // sSet0 holds the layers of one segment
    const long long waves = (tiles * s + slots - 1) / slots;
    const double cost = (double)waves * ((double)(npl + s - 1) / s + C::P + 1.5);  // +1.5: per-CTA set-up in layer units
    if (cost < bestcost) { bestcost = cost; best = s; }
  }
  prm.nseg = forced > 0 ? (int)std::min<int64_t>(forced, npl) : best;
  if ((npl + prm.nseg - 1) / prm.nseg + C::P > C::MAXL) prm.nseg = (int)((npl + C::MAXL - C::P - 1) / (C::MAXL - C::P));
  const long long blocks = tiles * prm.nseg;
  if (blocks > 0x7fffffffLL) return B2_EUNSUPPORTED;
  {
    KernelTimer timer(ctx);
    kern<<<(unsigned)blocks, C::NT, smem, ctx->stream>>>(prm, gmap);
  }
  ctx->launches++;
  {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
      return b2_fail(ctx, B2_ECUDA, std::string("k_rows3d launch failed: ") + cudaGetErrorString(e) + " blocks " + std::to_string(blocks) + " threads " + std::to_string(C::NT) + " smem " +
                                        std::to_string(smem) + " gpre " + std::to_string((int)GPRE) + " planes " + std::to_string(prm.plane_begin) + ":" + std::to_string(prm.plane_end));
  }
  return B2_OK;
}

// scalar forms: with the geometry precomputed (default; option "rows_gpre" = 0 disables it, as does a failed allocation of
// the geometry array) or evaluated inside the tile kernel
template <class C, bool FK, bool FM>
int launch_rows_scalar(b2_ctx* ctx, RowParams& prm) {
  const bool gpre = !(ctx->opts.count("rows_gpre") && ctx->opts["rows_gpre"] == 0);
  // the coefficient of a scalar form is symmetric (checked by the caller).  Measured (profiles/r02/README.md): the symmetric variant
  // (half the S3 and two thirds of the S2 work against scattered transposed stores) wins at degree 4 (K and M in one launch: -30 %);
  // at degree 2 the full variant is 3 % faster since its interior store takes the slot bases from registers (4.95 against 5.09 ms
  // at 128^3), at degree 3 2 %; at degree 1 the symmetric variant wins (128^3: 1.33 against 1.55 ms)
  const bool sym = ctx->opts.count("rows_sym") ? ctx->opts["rows_sym"] != 0 : (C::P == 4 || C::P == 1);
  if (gpre) {
    const int rc = sym ? launch_rows_cfg<C, FK, FM, 1, false, true, true>(ctx, prm) : launch_rows_cfg<C, FK, FM, 1, false, true, false>(ctx, prm);
    if (rc != B2_ENOMEM) return rc;
  }
  return launch_rows_cfg<C, FK, FM, 1, false, false, false>(ctx, prm);
}

// vector-valued launches: the precomputed geometry (10 components per column component) is available but NOT the default -- measured
// on 96^3 p=2 elasticity it changes nothing (31.2 against 30.5 ms: with 16 warps the in-kernel geometry already runs in the shadow of
// S2) and costs 5.7 GB; option "rows_gpre_vec" = 1 selects it
template <class C, bool FK, bool FM, int NFORM, bool TRN = false>
int launch_rows_vec(b2_ctx* ctx, RowParams& prm) {
  const bool gpre = ctx->opts.count("rows_gpre_vec") && ctx->opts["rows_gpre_vec"] != 0;
  if (gpre) {
    const int rc = launch_rows_cfg<C, FK, FM, NFORM, true, true, false, TRN>(ctx, prm);
    if (rc != B2_ENOMEM) return rc;
  }
  return launch_rows_cfg<C, FK, FM, NFORM, true, false, false, TRN>(ctx, prm);
}

// diagonal block (crow, crow) of a symmetric vector-valued form: the symmetric scalar pipeline (precomputed geometry, unordered dof
// pairs) writing into the slots of the vector-valued pattern
template <class C>
int launch_rows_vec_diag(b2_ctx* ctx, RowParams& prm) {
  const bool gpre = !(ctx->opts.count("rows_gpre") && ctx->opts["rows_gpre"] == 0);
  if (gpre) {
    const int rc = launch_rows_cfg<C, true, false, 1, true, true, true>(ctx, prm);
    if (rc != B2_ENOMEM) return rc;
  }
  return launch_rows_cfg<C, true, false, 1, true, false, true>(ctx, prm);
}

template <class C>
int launch_rows_forms(b2_ctx* ctx, RowParams& prm, bool fk, bool fm) {
  if (fk && fm) return launch_rows_scalar<C, true, true>(ctx, prm);
  if (fk) return launch_rows_scalar<C, true, false>(ctx, prm);
  return launch_rows_scalar<C, false, true>(ctx, prm);
}

// Vector-valued space (3 components): one launch per (matrix form, row component).  Stiffness-like forms (any 3x3 block
// D[c][1+x][e][1+y], e.g. elasticity) run the general-coefficient pipeline with the 3 column components as forms; mass-like
// forms (D[c][0][e][0]) run the scalar mass chain once and store rho[c][e] M; the load vector rides along.
int launch_rows_vector(b2_ctx* ctx, RowParams& base, const FormView& F, const double* const* D_host, const double* const* C_host, int P) {
  constexpr int nc = 3, na = 4;
  if (P < 1 || P > 4) return B2_EUNSUPPORTED;
  if (F.nvec > 1 || (F.nmat == 0 && F.nvec == 0)) return B2_EUNSUPPORTED;
  int kind[B2_MAX_FORMS];  // 0 stiffness-like, 1 mass-like
  bool symm[B2_MAX_FORMS];
  for (int m = 0; m < F.nmat; m++) {
    bool gg = false, mass = false, mixed = false;
    for (int c = 0; c < nc; c++)
      for (int e = 0; e < nc; e++)
        for (int x = 0; x < na; x++)
          for (int y = 0; y < na; y++) {
            const double v = D_host[m][((c * na + x) * nc + e) * na + y];
            if (v == 0.) continue;
            if (x && y) gg = true;
            else if (!x && !y) mass = true;
            else mixed = true;
          }
    if (mixed || (gg && mass) || (!gg && !mass)) return B2_EUNSUPPORTED;
    kind[m] = gg ? 0 : 1;
    // D[c][x][e][y] == D[e][y][c][x]: the matrix is symmetric (elasticity and every form that derives from an energy)
    symm[m] = gg && !(ctx->opts.count("rows_vecsym") && ctx->opts["rows_vecsym"] == 0);
    double dmax = 0.;
    for (int t = 0; t < nc * na * nc * na; t++) dmax = std::max(dmax, std::fabs(D_host[m][t]));
    for (int c = 0; c < nc && symm[m]; c++)
      for (int e = 0; e < nc; e++)
        for (int x = 1; x < na; x++)
          for (int y = 1; y < na; y++)
            if (std::fabs(D_host[m][((c * na + x) * nc + e) * na + y] - D_host[m][((e * na + y) * nc + c) * na + x]) > 4e-16 * dmax) symm[m] = false;
  }
  if (F.nvec)
    for (int c = 0; c < nc; c++)
      for (int x = 1; x < na; x++)
        if (C_host[0][c * na + x] != 0.) return B2_EUNSUPPORTED;
  for (int c = 0; c < nc; c++) {
    bool want_f = F.nvec > 0;
    for (int m = 0; m <= F.nmat; m++) {
      if (m == F.nmat && !want_f) break;
      RowParams prm = base;
      prm.crow = c;
      prm.has_f = want_f;
      if (want_f) {
        prm.vcoef = C_host[0][c * na];
        prm.rhs = F.rhs[0];
        want_f = false;
      }
      int rc;
      if (m < F.nmat && kind[m] == 0 && symm[m]) {
        // symmetric form: the diagonal block (c, c) by the symmetric scalar pipeline, the blocks (c, e > c) by the general one with every
        // entry stored transposed into block (e, c) as well -- 6 of the 9 blocks are integrated, 3 of them at the symmetric cost
        auto blk = [&](int e, int x, int y) { return D_host[m][((c * na + 1 + x) * nc + e) * na + 1 + y]; };
        RowParams pd = prm;
        pd.ecol0 = c;
        pd.valK = F.values[m];
        const double kc[6] = {blk(c, 0, 0), blk(c, 0, 1), blk(c, 0, 2), blk(c, 1, 1), blk(c, 1, 2), blk(c, 2, 2)};
        for (int t = 0; t < 6; t++) pd.kc[t] = kc[t];
        pd.iso = kc[1] == 0. && kc[2] == 0. && kc[4] == 0. && kc[0] == kc[3] && kc[0] == kc[5];
        rc = P == 1   ? launch_rows_vec_diag<RCfg<1, 9, 9, 2, 512, 0>>(ctx, pd)
             : P == 2 ? launch_rows_vec_diag<RCfg<2, 4, 4, 3, 256, 0>>(ctx, pd)
             : P == 3 ? launch_rows_vec_diag<RCfg<3, 3, 3, 2, 256, 3>>(ctx, pd)
                      : launch_rows_vec_diag<RCfg<4, 2, 3, 1, 256, 3>>(ctx, pd);
        const int nf = nc - 1 - c;  // column components e > c
        if (rc == B2_OK && nf > 0) {
          RowParams po = prm;
          po.has_f = 0;
          po.ecol0 = c + 1;
          po.valK = F.values[m];
          for (int f = 0; f < nf; f++)
            for (int x = 0; x < 3; x++)
              for (int y = 0; y < 3; y++) po.dm[f][x * 3 + y] = blk(c + 1 + f, x, y);
          const int64_t ovar = ctx->opts.count("rows_vec_offdiag") ? ctx->opts["rows_vec_offdiag"] : 0;
          if (P == 2 && ovar != 2) {
            // one launch per block on the register-rich 256-thread configuration
            for (int f = 0; f < nf && rc == B2_OK; f++) {
              RowParams pe = po;
              for (int t = 0; t < 9; t++) pe.dm[0][t] = po.dm[f][t];
              pe.ecol0 = c + 1 + f;
              rc = launch_rows_vec<RCfg<2, 4, 4, 3, 256, 0, 10>, true, false, 1, true>(ctx, pe);
            }
          } else if (P <= 2) {
            if (nf == 2) rc = P == 1 ? launch_rows_vec<RCfg<1, 8, 8, 2, 512, 0, 10>, true, false, 2, true>(ctx, po) : launch_rows_vec<RCfg<2, 4, 4, 3, 512, 1, 10>, true, false, 2, true>(ctx, po);
            else rc = P == 1 ? launch_rows_vec<RCfg<1, 8, 8, 2, 512, 0, 10>, true, false, 1, true>(ctx, po) : launch_rows_vec<RCfg<2, 4, 4, 3, 512, 1, 10>, true, false, 1, true>(ctx, po);
          } else {
            for (int f = 0; f < nf && rc == B2_OK; f++) {
              RowParams pe = po;
              for (int t = 0; t < 9; t++) pe.dm[0][t] = po.dm[f][t];
              pe.ecol0 = c + 1 + f;
              rc = P == 3 ? launch_rows_vec<RCfg<3, 3, 3, 2, 256, 3, 10>, true, false, 1, true>(ctx, pe) : launch_rows_vec<RCfg<4, 2, 2, 1, 256, 3, 10>, true, false, 1, true>(ctx, pe);
            }
          }
        }
      } else if (m < F.nmat && kind[m] == 0) {
        for (int e = 0; e < nc; e++)
          for (int x = 0; x < 3; x++)
            for (int y = 0; y < 3; y++) prm.dm[e][x * 3 + y] = D_host[m][((c * na + 1 + x) * nc + e) * na + 1 + y];
        prm.valK = F.values[m];
        if (P <= 2) {
          // degree 2: 512 threads (one dof pair per thread: 3 x 9 accumulators) with the S1 items split in term groups -- 48^3 elasticity 6.25 -> 4.82 ms
          rc = P == 1 ? launch_rows_vec<RCfg<1, 8, 8, 2, 512, 0, 10>, true, false, 3>(ctx, prm) : launch_rows_vec<RCfg<2, 4, 4, 3, 512, 1, 10>, true, false, 3>(ctx, prm);
        } else {
          // degrees 3 and 4: the accumulators of three column components do not fit the register file together --
          // one launch per column component e, writing the slots (J, e) of the rows (I, crow)
          rc = B2_OK;
          for (int e = 0; e < nc && rc == B2_OK; e++) {
            RowParams pe = prm;
            for (int t = 0; t < 9; t++) pe.dm[0][t] = prm.dm[e][t];
            pe.ecol0 = e;
            if (e) pe.has_f = 0;
            rc = P == 3 ? launch_rows_vec<RCfg<3, 3, 3, 2, 256, 3, 10>, true, false, 1>(ctx, pe) : launch_rows_vec<RCfg<4, 2, 2, 1, 256, 3, 10>, true, false, 1>(ctx, pe);
          }
        }
      } else {
        // mass-like form, or the load vector alone (mass chain without a matrix)
        prm.valM = nullptr;
        if (m < F.nmat) {
          for (int e = 0; e < nc; e++) prm.rhoe[e] = D_host[m][((c * na) * nc + e) * na];
          prm.valM = F.values[m];
        }
        rc = P == 1   ? launch_rows_cfg<RCfg<1, 8, 8, 2, 256, 0>, false, true, 1, true>(ctx, prm)
             : P == 2 ? launch_rows_cfg<RCfg<2, 4, 4, 3, 256, 0>, false, true, 1, true>(ctx, prm)
             : P == 3 ? launch_rows_cfg<RCfg<3, 3, 3, 2, 256, 3>, false, true, 1, true>(ctx, prm)
                      : launch_rows_cfg<RCfg<4, 2, 2, 1, 256, 3>, false, true, 1, true>(ctx, prm);
      }
      if (rc != B2_OK) return rc;
    }
  }
  return B2_OK;
}

}  // namespace

// Owner-computes assembly of the dof planes [plane_begin, plane_end) of dimension 0.  Returns
// B2_EUNSUPPORTED when the configuration is outside the specialised kernel (caller falls back).
int launch_assemble_rows(b2_ctx* ctx, const b2_basis* basis, const b2_quad* quad, const BasisView& B, const QuadView& Q, const GeomView& G, const FormView& F,
                         const double* const* D_host, const double* const* C_host, long long plane_begin, long long plane_end) {
  if (B.ndims != 3 || (B.ncomp != 1 && B.ncomp != 3)) return B2_EUNSUPPORTED;
  const int P = B.p[0];
  if (B.p[1] != P || B.p[2] != P || P < 1 || P > 4) return B2_EUNSUPPORTED;
  for (int d = 0; d < 3; d++) {
    if (Q.nq[d] != P + 1) return B2_EUNSUPPORTED;
    // maximal smoothness: element e carries dofs e..e+P
    if (B.ndofs[d] != B.nel[d] + P) return B2_EUNSUPPORTED;
    for (int e = 0; e < B.nel[d]; e++)
      if (basis->start[d][e] != e) return B2_EUNSUPPORTED;
  }
  if (F.nmat > 2 || F.nvec > 1 || (F.nmat == 0 && F.nvec == 0)) return B2_EUNSUPPORTED;
  RowParams prm;
  memset(&prm, 0, sizeof(prm));
  prm.B = B;
  prm.Q = Q;
  prm.G = G;
  prm.plane_begin = (int)plane_begin;
  prm.plane_end = (int)plane_end;
  prm.rho = 1.;
  prm.kc[0] = prm.kc[3] = prm.kc[5] = 1.;
  prm.ncomp = B.ncomp;
  prm.dbg = ctx->opts.count("rows_dbg") ? (int)ctx->opts["rows_dbg"] : 0;
  GeomCache gcache;
  prm.host_gcache = &gcache;
  // dominant coefficient set per dimension and its table at the 1-D points (host copy of what get_tabs uploaded)
  for (int d = 0; d < 3; d++) {
    std::vector<int> count(basis->nsets[d], 0);
    for (int e = 0; e < B.nel[d]; e++) count[basis->setidx[d][e]]++;
    const int cs = (int)(std::max_element(count.begin(), count.end()) - count.begin());
    prm.cset[d] = cs;
    for (int q = 0; q <= P; q++)
      for (int a = 0; a <= P; a++) {
        const double* c = &basis->coeffs[d][((size_t)cs * (P + 1) + a) * (P + 1)];
        const double x = quad->pts[d][q];
        double v = c[0], g = 0.;
        for (int j = 1; j <= P; j++) { g = g * x + v; v = v * x + c[j]; }
        prm.ctab[d][q][a][0] = v;
        prm.ctab[d][q][a][1] = g;
      }
  }
  for (int q = 0; q <= P && P <= 2; q++)
    for (int a = 0; a <= P; a++)
      for (int b = 0; b <= P; b++) {
        const double va = prm.ctab[2][q][a][0], da = prm.ctab[2][q][a][1], vb = prm.ctab[2][q][b][0], db = prm.ctab[2][q][b][1];
        prm.cp2[q][a][b][0] = va * vb;
        prm.cp2[q][a][b][1] = va * db;
        prm.cp2[q][a][b][2] = da * vb;
        prm.cp2[q][a][b][3] = da * db;
      }
  if (B.ncomp == 3) return launch_rows_vector(ctx, prm, F, D_host, C_host, P);
  bool fk = false, fm = false;
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
      if (fk) return B2_EUNSUPPORTED;
      fk = true;
      const double kc[6] = {D[5], D[6], D[7], D[10], D[11], D[15]};
      for (int t = 0; t < 6; t++) prm.kc[t] = kc[t];
      prm.valK = F.values[m];
    } else {
      if (fm) return B2_EUNSUPPORTED;
      fm = true;
      prm.rho = D[0];
      prm.valM = F.values[m];
    }
  }
  prm.iso = prm.kc[1] == 0. && prm.kc[2] == 0. && prm.kc[4] == 0. && prm.kc[0] == prm.kc[3] && prm.kc[0] == prm.kc[5];
  if (F.nvec) {
    const double* Cv = C_host[0];  // [4]
    if (Cv[1] != 0. || Cv[2] != 0. || Cv[3] != 0.) return B2_EUNSUPPORTED;
    prm.has_f = 1;
    prm.vcoef = Cv[0];
    prm.rhs = F.rhs[0];
    if (!fk && !fm) fm = true, prm.valM = nullptr;  // load vector only: run the mass chain without storing it
  }
  // Tile / thread configuration.  Measured at 128^3 (profiles/r01/variants.txt): few register-rich warps with unsplit
  // S1/S2 items (256 threads, 230 registers, no spills) beat more, thinner warps (512 threads cap at 128 registers and spill).
  const int64_t variant = ctx->opts.count("rows_variant") ? ctx->opts["rows_variant"] : 0;
  (void)variant;
  // degree 1: accumulators are small (2 x 2 per pair and form), so 16 thin warps hide more latency than 8 register-rich ones, and a
  // 9 x 9 tile still fits (128^3: 1.59 ms with <8,8,256 threads> -> 1.28 ms)
  if (P == 1) return launch_rows_forms<RCfg<1, 9, 9, 2, 512, 0>>(ctx, prm, fk, fm);
  if (P == 4) {
    // degree 4 (the high-order IGA case): 2 x 3 dof columns per CTA, one point-plane per pipeline step, S1/S2 items split
    // in term groups; K and M in separate launches (25 accumulators per dof pair and form, two dof pairs per thread)
    using C4 = RCfg<4, 2, 3, 1, 256, 3>;  // 2 x 3 dof columns: 486 of 512 accumulator slots used, halo 6 x 7 elements (48^3: 14.4 -> 13.7 ms against 2 x 2)
    const bool sym4 = !(ctx->opts.count("rows_sym") && ctx->opts["rows_sym"] == 0) && !(ctx->opts.count("rows_gpre") && ctx->opts["rows_gpre"] == 0);
    // symmetric variant: one dof pair per thread (246 of 256), 2 x 25 accumulators: K and M in ONE launch
    if (!(fk && fm) || (sym4 && !(ctx->opts.count("rows_split_forms") && ctx->opts["rows_split_forms"]))) return launch_rows_forms<C4>(ctx, prm, fk, fm);
    RowParams pk = prm;
    pk.valM = nullptr;
    int rc = launch_rows_scalar<C4, true, false>(ctx, pk);
    if (rc != B2_OK) return rc;
    RowParams pm = prm;
    pm.valK = nullptr;
    pm.has_f = 0;
    return launch_rows_scalar<C4, false, true>(ctx, pm);
  }
  if (P == 3) {
    // two chunks of two point-planes per layer
    using C3 = RCfg<3, 3, 3, 2, 256, 3>;
    // K and M together: 512 threads (one dof pair per thread: 2 x 16 accumulators, 128 registers) -- 96^3: 14.5 ms
    // (the 512-thread one-pair-per-thread configuration of round 1 is what the symmetric variant gives with 256 threads)
    if (fk && fm && !(ctx->opts.count("rows_sym") && ctx->opts["rows_sym"] != 0) && !(ctx->opts.count("rows_gpre") && ctx->opts["rows_gpre"] == 0) && !(ctx->opts.count("rows_split_forms") && ctx->opts["rows_split_forms"])) return launch_rows_cfg<RCfg<3, 3, 3, 2, 512, 3>, true, true, 1, false, true, false>(ctx, prm);
    // K and M in one launch (the geometry stage runs once; 2 x 2 x 16 accumulators per thread spill ~0.7 KB to L1, still
    // 10 % faster than two launches: 96^3 17.7 -> 15.9 ms); option "rows_split_forms" = 1 selects the two launches
    if (!(fk && fm) || !(ctx->opts.count("rows_split_forms") && ctx->opts["rows_split_forms"])) return launch_rows_forms<C3>(ctx, prm, fk, fm);
    RowParams pk = prm;
    pk.valM = nullptr;
    int rc = launch_rows_scalar<C3, true, false>(ctx, pk);
    if (rc != B2_OK) return rc;
    RowParams pm = prm;
    pm.valK = nullptr;
    pm.has_f = 0;
    return launch_rows_scalar<C3, false, true>(ctx, pm);
  }
#ifdef B2_EXPERIMENT
  if (fk && fm) {
    if (variant == 1) return launch_rows_cfg<RCfg<2, 4, 4, 3, 512, 3>, true, true>(ctx, prm);
    if (variant == 2) return launch_rows_cfg<RCfg<2, 4, 4, 3, 256, 3>, true, true>(ctx, prm);
    if (variant == 3) return launch_rows_cfg<RCfg<2, 5, 3, 3, 384, 3>, true, true>(ctx, prm);
    if (variant == 4) return launch_rows_cfg<RCfg<2, 3, 5, 3, 384, 3>, true, true>(ctx, prm);
    if (variant == 5) return launch_rows_cfg<RCfg<2, 4, 4, 3, 384, 3>, true, true>(ctx, prm);
    if (variant == 6) return launch_rows_cfg<RCfg<2, 4, 4, 3, 320, 0>, true, true>(ctx, prm);
    if (variant == 7) return launch_rows_cfg<RCfg<2, 4, 4, 3, 384, 0>, true, true>(ctx, prm);
    if (variant == 8) return launch_rows_cfg<RCfg<2, 4, 4, 3, 256, 2>, true, true>(ctx, prm);
    if (variant == 9) return launch_rows_cfg<RCfg<2, 4, 4, 3, 256, 1>, true, true>(ctx, prm);
    if (variant == 10) return launch_rows_cfg<RCfg<2, 4, 4, 3, 320, 2>, true, true>(ctx, prm);
    if (variant == 11) return launch_rows_cfg<RCfg<2, 4, 4, 3, 512, 1>, true, true>(ctx, prm);
    if (variant == 12) return launch_rows_cfg<RCfg<2, 4, 4, 3, 512, 0>, true, true>(ctx, prm);
    if (variant == 13) return launch_rows_cfg<RCfg<2, 4, 4, 3, 384, 1>, true, true>(ctx, prm);
    if (variant == 30) return launch_rows_cfg<RCfg<2, 4, 4, 3, 256, 0>, true, true, 1, false, true>(ctx, prm);
    if (variant == 40) return launch_rows_cfg<RCfg<2, 4, 4, 3, 256, 0>, true, true, 1, false, true, true>(ctx, prm);
    if (variant == 41) return launch_rows_cfg<RCfg<2, 4, 4, 3, 256, 0>, true, true, 1, false, false, true>(ctx, prm);
    if (variant == 47) return launch_rows_cfg<RCfg<2, 4, 4, 3, 256, 1>, true, true, 1, false, true, false>(ctx, prm);
    if (variant == 48) return launch_rows_cfg<RCfg<2, 4, 4, 3, 256, 3>, true, true, 1, false, true, false>(ctx, prm);
    if (variant == 49) return launch_rows_cfg<RCfg<2, 4, 4, 3, 256, 2>, true, true, 1, false, true, false>(ctx, prm);
    if (variant == 42) return launch_rows_cfg<RCfg<2, 4, 4, 3, 512, 1>, true, true, 1, false, true, true>(ctx, prm);
    if (variant == 43) return launch_rows_cfg<RCfg<2, 4, 4, 3, 384, 1>, true, true, 1, false, true, true>(ctx, prm);
    if (variant == 44) return launch_rows_cfg<RCfg<2, 4, 4, 1, 256, 1, 7, 2>, true, true, 1, false, true, true>(ctx, prm);
    if (variant == 45) return launch_rows_cfg<RCfg<2, 4, 4, 1, 256, 3, 7, 2>, true, true, 1, false, true, true>(ctx, prm);
    if (variant == 46) return launch_rows_cfg<RCfg<2, 4, 4, 3, 512, 3>, true, true, 1, false, true, true>(ctx, prm);
    if (variant == 31) return launch_rows_cfg<RCfg<2, 4, 4, 3, 512, 1>, true, true, 1, false, true>(ctx, prm);
    if (variant == 20) return launch_rows_cfg<RCfg<2, 4, 4, 1, 256, 1, 7, 2>, true, true>(ctx, prm);
    if (variant == 21) return launch_rows_cfg<RCfg<2, 4, 4, 1, 256, 3, 7, 2>, true, true>(ctx, prm);
    if (variant == 22) return launch_rows_cfg<RCfg<2, 4, 4, 1, 256, 0, 7, 2>, true, true>(ctx, prm);
    if (variant == 24) return launch_rows_cfg<RCfg<2, 4, 4, 1, 256, 0, 7, 1>, true, true>(ctx, prm);
    if (variant == 25) return launch_rows_cfg<RCfg<2, 4, 4, 1, 256, 1, 7, 1>, true, true>(ctx, prm);
    if (variant == 26) return launch_rows_cfg<RCfg<2, 4, 4, 1, 128, 1, 7, 3>, true, true>(ctx, prm);
  }
#endif
  return launch_rows_forms<RCfg<2, 4, 4, 3, 256, 0>>(ctx, prm, fk, fm);
}
