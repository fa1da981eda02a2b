// pms_device_fast.cuh — flop-reduced hex8 corner metrics for the default (non-strict) build.
//
// The metric + gradient pass is FP64-pipe-bound on B200 (measured 64 FP64 lanes/clk/SM, tools/fp64_peak.cu), so
// the lever is the FP64 instruction count.  The formulas below are algebraically identical to
// HexMeshSmootherMetric::{metric,grad_metric} (SM/SmootherMetric.hpp:613-967) but cost ~0.6x the instructions:
//
//  * canonical edge vectors: corner c of a hex (node = i+2j+4k) uses the three edges along +i,+j,+k that touch
//    it; [Ei Ej Ek] is the reference's corner Jacobian up to an even signed column permutation, which leaves
//    det A, det W and T = A W^-1 unchanged (checked against the locs table, sgrid_metric_helpers.hpp:67-69).
//  * T' = A cof(W)^T (= det W * T): ||T||_F^2 = ||T'||_F^2 / detW^2, no inverse of W is formed.
//  * one rsqrt replaces sqrt + three divisions:  with y = ||T'||_F^2, P = 3 sqrt(3) |detW| detA,
//        r = rsqrt(y P^2) = 1/(P sqrt y):   mu_c = f^3/(3 sqrt3 d) detW - detW = y^2 r - detW
//  * gradient (lever (i) of SURVEY section 7: adjT(A W^-1) W^-T = adjT(A)/detW):
//        dmu/dA = k1 (T' cof W) - k2 cof(A),   k1 = 3 y r,   k2 = y^2 r / detA
//  * rsqrt / reciprocal: MUFU seed + 2 Newton steps (no slow-path branches); relative error a few ulp.
// Results agree with the reference-order code to rounding level (tests/test_gpu_parity.py: <=1e-12 on totals).
// The strict_fp build does not use this file.
#pragma once
#include "pms_device.cuh"

#ifndef PMS_GRAD_VIA_GRAM
#define PMS_GRAD_VIA_GRAM 1
#endif
#ifndef PMS_GRAD_EDGE_ACC
#define PMS_GRAD_EDGE_ACC 1  // gradient accumulated per hex EDGE first (see hex_metric_fast_v)
#endif
#ifndef PMS_RSQRT_CUBIC
#define PMS_RSQRT_CUBIC 1
#endif
#ifndef PMS_CORNER_ORDER
#define PMS_CORNER_ORDER 1  // Gray code: 1.3 % faster element pass than the natural order (profiles/r02_ab_corner_order.log)
#endif
#ifndef PMS_NO_OK_SELECT
#define PMS_NO_OK_SELECT 1  // ShapeB1 gradient: no per-corner selects on detA > 0 (the element-level rule covers them)
#endif
#ifndef PMS_USE_BRANCH
#define PMS_USE_BRANCH 0    // invalid element: branch around the 24 stores instead of 24 selects
#endif


namespace pmsdev {

PMS_DI double fast_rsqrt(double x) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#if PMS_RSQRT_CUBIC
  // one third-order step: with e = (1 - x r^2)/2 the exact value is r (1 - 2e)^(-1/2) = r (1 + e + 1.5 e^2 + O(e^3)); the
  // seed carries >= 20 bits, so the remainder 2.5 e^3 is below 2^-58.  5 FP64 on a 4-deep chain instead of 6 on a 6-deep one.
  const double h = (0.5 * x) * r;
  const double e = fma(-h, r, 0.5);
  return fma(r * e, fma(1.5, e, 1.0), r);
#else
  const double hx = 0.5 * x;
#pragma unroll
  for (int it = 0; it < 2; it++) {
    const double h = hx * r;
    const double e = fma(-h, r, 0.5);
    r = fma(r, e, r);
  }
  return r;
#endif
}
PMS_DI double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#pragma unroll
  for (int it = 0; it < 2; it++) {
    const double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
  }
  return r;
}

PMS_DI void cross3(const double (&a)[3], const double (&b)[3], double (&c)[3]) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
PMS_DI double dot3(const double (&a)[3], const double (&b)[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// vertex accessors: V(p, r) = component r of the vertex at structured position p (node = i+2j+4k)
struct RegVerts {  // vertices already in registers
  const double (&v)[8][3];
  PMS_DI explicit RegVerts(const double (&vv)[8][3]) : v(vv) {}
  PMS_DI double operator()(int p, int r) const { return v[p][r]; }
};
template <bool EXODUS, int STRIDE>
struct SmemVerts {  // vertices read in place from a thread-private shared-memory column sm[(a*3+r)*STRIDE]
  const double* base;  // slot of (local node 0, comp 0) of the wanted array for this thread
  PMS_DI explicit SmemVerts(const double* b) : base(b) {}
  PMS_DI double operator()(int p, int r) const {
    const int a = EXODUS ? (p == 2 ? 3 : (p == 3 ? 2 : (p == 6 ? 7 : (p == 7 ? 6 : p)))) : p;  // structured -> Exodus local node
    return base[(a * 3 + r) * STRIDE];
  }
};

template <bool EXODUS, int STRIDE>
struct SmemVertsReload {  // as SmemVerts, but every access is a fresh LDS: no vertex value stays live across corners
  const volatile double* base;
  PMS_DI explicit SmemVertsReload(const double* b) : base(b) {}
  PMS_DI double operator()(int p, int r) const {
    const int a = EXODUS ? (p == 2 ? 3 : (p == 3 ? 2 : (p == 6 ? 7 : (p == 7 ? 6 : p)))) : p;
    return base[(a * 3 + r) * STRIDE];
  }
};
template <int NN>
struct IdxVerts {  // vertices of a brick staged once per distinct node: base[r*NN + local node index]
  const double* base;
  const int (&off)[8];  // local node index of structured position p
  PMS_DI IdxVerts(const double* b, const int (&o)[8]) : base(b), off(o) {}
  PMS_DI double operator()(int p, int r) const { return base[r * NN + off[p]]; }
};

// canonical edge along axis AX touching corner c (vertices in structured numbering node = i+2j+4k)
template <int AX, class V>
PMS_DI void corner_edge(const V& v, int c, double (&e)[3]) {
  const int lo = c & ~(1 << AX), hi = lo | (1 << AX);
#pragma unroll
  for (int r = 0; r < 3; r++) e[r] = v(hi, r) - v(lo, r);
}

// Hex element metric (+ gradient when GRAD) in structured vertex numbering.  STAGE 0 untangle, 1 ShapeB1 with
// reference mesh.  grad[a][r] accumulates d(metric)/d(x_a,r); the caller applies the element-level validity rule.
// Written branch-free (selects instead of `if (detA > 0)`): the 8 corners then sit in ONE basic block and ptxas can
// interleave their dependent FP64 chains, which is what hides the ~9-cycle DFMA latency at 2 warps per scheduler.
//
// Gradient assembly (PMS_GRAD_EDGE_ACC): a corner's derivative dM[r][s] w.r.t. its canonical edge along axis s goes
// +dM to the edge's high vertex and -dM to its low vertex, and each of the 12 hex edges is the canonical edge of exactly
// its two end corners.  So the two corners' terms  k1 (A G)[r][s] - k2 cof(A)[r][s]  are chained by FMA into one edge
// accumulator E (1 DMUL + 3 DFMA per component), and a vertex's gradient is the signed sum of its three edges' E
// (2 DADD per component): 12*12 + 8*6 = 192 FP64 per hex instead of 8*(18 + 18) = 288 for "form dM, then add it to
// both ends".  Same terms, different association: differs from the other form by rounding only.
struct NoMid {
  PMS_DI void operator()() const {}
};
// `mid` is invoked once between corners 3 and 4: the one-kernel gradient (pms_elem_ws.cuh) places the second half of its
// node gather there, half an iteration after the first.
template <int STAGE, bool GRAD, class VC, class VO, class MID = NoMid>
PMS_DI double hex_metric_fast_v(const VC& vc, const VO& vo, bool& valid, double (&grad)[8][3], MID mid = MID()) {
  valid = true;
  double val0 = 0.0, val1 = 0.0;
  double E[3][4][3];  // [axis s][edge index among the 4 along s][component r]
  if (GRAD && !PMS_GRAD_EDGE_ACC) {
#pragma unroll
    for (int a = 0; a < 8; a++)
#pragma unroll
      for (int r = 0; r < 3; r++) grad[a][r] = 0.0;
  }
#pragma unroll
  for (int cc = 0; cc < 8; cc++) {
    if (cc == 4) mid();
    // corner visiting order (PMS_CORNER_ORDER): 0 natural; 1 Gray code (consecutive corners share an edge, so every edge
    // accumulator is completed as early as possible); 2 k-pairs first.  The edge accumulators are keyed on the corner's
    // position (lo/hi), so the order only changes which of an edge's two corners is visited first.
    constexpr int ORD[3][8] = {{0, 1, 2, 3, 4, 5, 6, 7}, {0, 1, 3, 2, 6, 7, 5, 4}, {0, 4, 1, 5, 3, 7, 2, 6}};
    const int c = ORD[PMS_CORNER_ORDER][cc];
    double a0[3], a1[3], a2[3], w0[3], w1[3], w2[3];
    corner_edge<0>(vc, c, a0);
    corner_edge<1>(vc, c, a1);
    corner_edge<2>(vc, c, a2);
    corner_edge<0>(vo, c, w0);
    corner_edge<1>(vo, c, w1);
    corner_edge<2>(vo, c, w2);
    double K[3][3];  // K[s] = cofactor column s of A = [a0 a1 a2]
    cross3(a1, a2, K[0]);
    const double detA = dot3(a0, K[0]);
    const bool ok = detA > 0.0;
    if (!ok) valid = false;
    // d(term)/d(edge s, component r) = k1 * Tp[r][s] + nk2 * K[s][r]
    double Tp[3][3], k1 = 0.0, nk2 = 0.0;
    if (STAGE == 0) {
      double C0[3];
      cross3(w1, w2, C0);
      const double detW = dot3(w0, C0);
      const double vv = BETA_MULT * detW - detA;
      const double vp = vv > 0.0 ? vv : 0.0;  // contributes 0 (value and gradient) when vv <= 0
      ((c & 1) ? val1 : val0) += vp * vp;
      if (GRAD) {
        cross3(a2, a0, K[1]);
        cross3(a0, a1, K[2]);
        nk2 = -2.0 * vp;
      }
    } else {
      double C0[3], C1[3], C2[3];
      cross3(w1, w2, C0);
      cross3(w2, w0, C1);
      cross3(w0, w1, C2);
      const double detW = dot3(w0, C0);
      double ys[3];
      if (GRAD && PMS_GRAD_VIA_GRAM) {
        // (T'C)[r][j] = sum_i a_i[r] (C_i.C_j): with the Gram matrix G = C C^T of the reference cofactors (6 dot
        // products) B = A G serves both y = |T'|^2 = A:B and the derivative, and T' itself is never formed:
        // 18 + 27 + 9 FP64 instead of 27 + 9 + 27 per corner.  Tp holds B here.
        const double g00 = dot3(C0, C0), g01 = dot3(C0, C1), g02 = dot3(C0, C2), g11 = dot3(C1, C1), g12 = dot3(C1, C2),
                     g22 = dot3(C2, C2);
#pragma unroll
        for (int r = 0; r < 3; r++) {
          Tp[r][0] = a0[r] * g00 + a1[r] * g01 + a2[r] * g02;
          Tp[r][1] = a0[r] * g01 + a1[r] * g11 + a2[r] * g12;
          Tp[r][2] = a0[r] * g02 + a1[r] * g12 + a2[r] * g22;
          ys[r] = a0[r] * Tp[r][0] + a1[r] * Tp[r][1] + a2[r] * Tp[r][2];
        }
      } else {
#pragma unroll
        for (int r = 0; r < 3; r++) {
#pragma unroll
          for (int s = 0; s < 3; s++) Tp[r][s] = a0[r] * C0[s] + a1[r] * C1[s] + a2[r] * C2[s];
          ys[r] = Tp[r][0] * Tp[r][0] + Tp[r][1] * Tp[r][1] + Tp[r][2] * Tp[r][2];
        }
      }
      const double y = (ys[0] + ys[1]) + ys[2];
      const double P = FAC_3SQRT3 * fabs(detW) * detA;
      double yr, yrd = 0.0;  // sqrt(y)/P and, for the derivative, sqrt(y)/(P detA)
      if (GRAD && PMS_GRAD_VIA_GRAM) {
        // one rsqrt serves both: u = rsqrt(y (P detA)^2) = 1/(sqrt(y) P detA), y u = sqrt(y)/(P detA), times detA = sqrt(y)/P
        // (saves the separate reciprocal of detA: 7 FP64 per corner)
        // No select on the argument: a corner with detA <= 0 makes the element invalid, the ShapeB1 gradient of an
        // invalid element is discarded as a whole (gradient_functors.hpp:391) and the value below is selected by `ok`,
        // so whatever inf/NaN the rsqrt of a zero argument produces never reaches a result.
        const double Pd = P * detA;
        yrd = y * fast_rsqrt((PMS_NO_OK_SELECT || ok) ? y * Pd * Pd : 1.0);
        yr = yrd * detA;
      } else {
        yr = y * fast_rsqrt(ok ? y * P * P : 1.0);
      }
      const double q = y * yr;   // y^1.5/P
      ((c & 1) ? val1 : val0) += ok ? q - detW : 0.0;
      if (GRAD) {
        cross3(a2, a0, K[1]);
        cross3(a0, a1, K[2]);
        // detA <= 0: the corner contributes nothing (SmootherMetric.hpp:835) - with the Gram form that is implied by
        // the element-level rule above (the caller drops the whole gradient when !valid), so no per-corner selects
        if (PMS_GRAD_VIA_GRAM && PMS_NO_OK_SELECT) {
          k1 = 3.0 * yr;
          nk2 = -y * yrd;
        } else if (PMS_GRAD_VIA_GRAM) {
          k1 = ok ? 3.0 * yr : 0.0;
          nk2 = ok ? -y * yrd : 0.0;
        } else {
          k1 = ok ? 3.0 * yr : 0.0;
          nk2 = ok ? -q * fast_rcp(detA) : 0.0;
        }
        if (!PMS_GRAD_VIA_GRAM) {  // Tp held T' = A C^T: the derivative needs T' C
#pragma unroll
          for (int r = 0; r < 3; r++) {
            const double t0 = Tp[r][0], t1 = Tp[r][1], t2 = Tp[r][2];
            Tp[r][0] = t0 * C0[0] + t1 * C0[1] + t2 * C0[2];
            Tp[r][1] = t0 * C1[0] + t1 * C1[1] + t2 * C1[2];
            Tp[r][2] = t0 * C2[0] + t1 * C2[1] + t2 * C2[2];
          }
        }
      }
    }
    if (GRAD) {
#pragma unroll
      for (int s = 0; s < 3; s++) {
        const int lo = c & ~(1 << s), hi = lo | (1 << s);
        if (PMS_GRAD_EDGE_ACC) {
          const int ei = ((lo >> (s + 1)) << s) | (lo & ((1 << s) - 1));  // lo with bit s squeezed out
          // first of the edge's two corners in the visiting order?
          bool first = true;
#pragma unroll
          for (int q = 0; q < 8; q++)
            if (q < cc && ORD[PMS_CORNER_ORDER][q] == (c ^ (1 << s))) first = false;
#pragma unroll
          for (int r = 0; r < 3; r++) {
            if (STAGE == 0)
              E[s][ei][r] = first ? nk2 * K[s][r] : fma(nk2, K[s][r], E[s][ei][r]);
            else
              E[s][ei][r] = fma(nk2, K[s][r], first ? k1 * Tp[r][s] : fma(k1, Tp[r][s], E[s][ei][r]));
          }
        } else {
#pragma unroll
          for (int r = 0; r < 3; r++) {
            const double d = STAGE == 0 ? nk2 * K[s][r] : k1 * Tp[r][s] + nk2 * K[s][r];
            grad[hi][r] += d;
            grad[lo][r] -= d;
          }
        }
      }
    }
  }
  if (GRAD && PMS_GRAD_EDGE_ACC) {
#pragma unroll
    for (int v = 0; v < 8; v++)
#pragma unroll
      for (int r = 0; r < 3; r++) {
        const double e0 = E[0][v >> 1][r], e1 = E[1][((v >> 2) << 1) | (v & 1)][r], e2 = E[2][v & 3][r];
        grad[v][r] = (((v & 1) ? e0 : -e0) + ((v & 2) ? e1 : -e1)) + ((v & 4) ? e2 : -e2);
      }
  }
  return val0 + val1;
}

// Hex metric at NB trial step lengths in one pass (line search).  With x(alpha) = x + alpha*s every corner edge is
// a + alpha*b, so per corner   det A(alpha) = d0 + d1 alpha + d2 alpha^2 + d3 alpha^3   (triple-product expansion) and
// T'(alpha) = A(alpha) cof(W)^T = Tx + alpha Ts, hence   y(alpha) = |T'|^2 = y0 + 2 y1 alpha + y2 alpha^2.
// The 8 coefficients cost about two single evaluations of the corner; each further alpha costs ~20 FP64 instructions
// (two Horner forms, one rsqrt) instead of ~94.  vs must already be zero at fixed nodes (update_coordinates.cpp:255).
// sum[k] += metric of this element at al[k]; bit k of invmask set when any corner has det A(al[k]) <= 0.
// NF < NB: only the LAST NF step lengths get their metric; for the first NB - NF only det A(alpha) > 0 is tested (3 FMA +
// a compare per corner instead of ~17 FP64) and their sums stay untouched.  The line search on an untangled mesh rejects a
// trial with any invalid element whatever its metric, and its trial sequence runs from far too long steps down to the
// accepted one, so the driver asks for values only near the previously accepted step (Driver::speculate).
template <int STAGE, int NB, int CORNER_UNROLL = 8, int NF = NB, class VC, class VS, class VO>
PMS_DI void hex_metric_multi(const VC& vc, const VS& vs, const VO& vo, const double (&al)[NB], double (&sum)[NB], unsigned& invmask) {
  invmask = 0;
  // CORNER_UNROLL < 8 keeps fewer corners' edge vectors and coefficients live (no spills), at the price of run-time
  // vertex addressing; measured per kernel
#pragma unroll CORNER_UNROLL
  for (int c = 0; c < 8; c++) {
    double a0[3], a1[3], a2[3], b0[3], b1[3], b2[3], w0[3], w1[3], w2[3];
    corner_edge<0>(vc, c, a0);
    corner_edge<1>(vc, c, a1);
    corner_edge<2>(vc, c, a2);
    corner_edge<0>(vs, c, b0);
    corner_edge<1>(vs, c, b1);
    corner_edge<2>(vs, c, b2);
    corner_edge<0>(vo, c, w0);
    corner_edge<1>(vo, c, w1);
    corner_edge<2>(vo, c, w2);
    double K0[3], L[3], L2[3], M[3], C0[3];
    cross3(a1, a2, K0);
    cross3(a1, b2, L);
    cross3(b1, a2, L2);
    cross3(b1, b2, M);
#pragma unroll
    for (int r = 0; r < 3; r++) L[r] += L2[r];
    const double d0 = dot3(a0, K0), d1 = dot3(a0, L) + dot3(b0, K0), d2 = dot3(a0, M) + dot3(b0, L), d3 = dot3(b0, M);
    cross3(w1, w2, C0);
    const double detW = dot3(w0, C0);
    if (STAGE == 0) {
      const double bw = BETA_MULT * detW;
#pragma unroll
      for (int k = 0; k < NB; k++) {
        const double detA = fma(al[k], fma(al[k], fma(al[k], d3, d2), d1), d0);
        if (!(detA > 0.0)) invmask |= 1u << k;
        if (k < NB - NF) continue;
        const double vv = bw - detA;
        const double vp = vv > 0.0 ? vv : 0.0;
        sum[k] += vp * vp;
      }
    } else {
      double C1[3], C2[3];
      cross3(w2, w0, C1);
      cross3(w0, w1, C2);
      double y0 = 0.0, y1 = 0.0, y2 = 0.0;
#pragma unroll
      for (int r = 0; r < 3; r++)
#pragma unroll
        for (int q = 0; q < 3; q++) {
          const double tx = a0[r] * C0[q] + a1[r] * C1[q] + a2[r] * C2[q];
          const double ts = b0[r] * C0[q] + b1[r] * C1[q] + b2[r] * C2[q];
          y0 += tx * tx;
          y1 += tx * ts;
          y2 += ts * ts;
        }
      const double y1x2 = 2.0 * y1, cw = FAC_3SQRT3 * fabs(detW);
#pragma unroll
      for (int k = 0; k < NB; k++) {
        const double detA = fma(al[k], fma(al[k], fma(al[k], d3, d2), d1), d0);
        const bool ok = detA > 0.0;
        if (k < NB - NF) {  // validity only
          if (!ok) invmask |= 1u << k;
          continue;
        }
        const double y = fma(al[k], fma(al[k], y2, y1x2), y0);
        const double P = cw * detA;
        // (no select on the argument: y P^2 >= 0, and whatever inf/NaN a zero argument gives is dropped by the select below)
        const double rr = fast_rsqrt(y * P * P);
        const double yr = y * rr;
        sum[k] += ok ? y * yr - detW : 0.0;
        if (!ok) invmask |= 1u << k;
      }
    }
  }
}

template <int STAGE, bool GRAD>
PMS_DI double hex_metric_fast(const double (&vc)[8][3], const double (&vo)[8][3], bool& valid, double (&grad)[8][3]) {
  return hex_metric_fast_v<STAGE, GRAD>(RegVerts(vc), RegVerts(vo), valid, grad);
}

// ShapeB1 metric only, restructured for instruction-level parallelism: the per-corner dependent chains (9-term
// Frobenius sum, rsqrt seed + 2 Newton steps) are the latency bottleneck at 2 warps per scheduler (ncu: "wait" is the
// top stall), so phase 1 computes y_c, P_c, detW_c for all 8 corners (3-way split sums), phase 2 runs the 8
// independent rsqrt/Newton chains interleaved.  Same arithmetic per corner as hex_metric_fast<1,false>.
PMS_DI double hex_shapeb1_metric_ilp(const double (&vc)[8][3], const double (&vo)[8][3], bool& valid) {
  double z[8], y[8], dW[8];
  bool ok[8];
  valid = true;
#pragma unroll
  for (int c = 0; c < 8; c++) {
    double a0[3], a1[3], a2[3], w0[3], w1[3], w2[3];
    corner_edge<0>(RegVerts(vc), c, a0);
    corner_edge<1>(RegVerts(vc), c, a1);
    corner_edge<2>(RegVerts(vc), c, a2);
    corner_edge<0>(RegVerts(vo), c, w0);
    corner_edge<1>(RegVerts(vo), c, w1);
    corner_edge<2>(RegVerts(vo), c, w2);
    double K0[3], C0[3], C1[3], C2[3];
    cross3(a1, a2, K0);
    const double detA = dot3(a0, K0);
    ok[c] = detA > 0.0;
    if (!ok[c]) valid = false;
    cross3(w1, w2, C0);
    cross3(w2, w0, C1);
    cross3(w0, w1, C2);
    dW[c] = dot3(w0, C0);
    double ys[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
      const double t0 = a0[r] * C0[0] + a1[r] * C1[0] + a2[r] * C2[0];
      const double t1 = a0[r] * C0[1] + a1[r] * C1[1] + a2[r] * C2[1];
      const double t2 = a0[r] * C0[2] + a1[r] * C1[2] + a2[r] * C2[2];
      ys[r] = t0 * t0 + t1 * t1 + t2 * t2;
    }
    y[c] = (ys[0] + ys[1]) + ys[2];
    const double P = FAC_3SQRT3 * fabs(dW[c]) * detA;
    z[c] = y[c] * P * P;  // >= 0; an invalid corner's inf/NaN is dropped by the select on ok[c] below
  }
  double r[8], hz[8];
#pragma unroll
  for (int c = 0; c < 8; c++) {
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r[c]) : "d"(z[c]));
    hz[c] = 0.5 * z[c];
  }
#pragma unroll
  for (int it = 0; it < 2; it++) {
    double e[8];
#pragma unroll
    for (int c = 0; c < 8; c++) e[c] = fma(-(hz[c] * r[c]), r[c], 0.5);
#pragma unroll
    for (int c = 0; c < 8; c++) r[c] = fma(r[c], e[c], r[c]);
  }
  double v0 = 0.0, v1 = 0.0;
#pragma unroll
  for (int c = 0; c < 8; c += 2) {
    v0 += ok[c] ? y[c] * (y[c] * r[c]) - dW[c] : 0.0;
    v1 += ok[c + 1] ? y[c + 1] * (y[c + 1] * r[c + 1]) - dW[c + 1] : 0.0;
  }
  return v0 + v1;
}

// One corner of a hex from its three edge vectors e_s = v[neighbour along s] - v[corner] (a* current, w* reference) and
// the orientation sign p = (-1)^popc(c) that relates them to the canonical +axis frame.  Returns the corner's metric
// term; detA_out = p*det[a0 a1 a2] is the reference's detA (validity test); dM[r][s] = d(term)/d(a_s[r]) when `have`.
template <int STAGE, bool GRAD>
PMS_DI double hex_corner_fast(const double (&a0)[3], const double (&a1)[3], const double (&a2)[3], const double (&w0)[3],
                              const double (&w1)[3], const double (&w2)[3], const double p, double& detA_out,
                              double (&dM)[3][3], bool& have) {
  double K0[3], K1[3], K2[3];
  cross3(a1, a2, K0);
  const double detAr = dot3(a0, K0);  // raw determinant; the reference's is p*detAr
  const double detA = p * detAr;
  detA_out = detA;
  have = false;
  double val = 0.0;
  if (STAGE == 0) {
    double C0[3];
    cross3(w1, w2, C0);
    const double detW = p * dot3(w0, C0);
    const double vv = BETA_MULT * detW - detA;
    if (vv > 0.0) {
      val = vv * vv;
      if (GRAD) {
        have = true;
        cross3(a2, a0, K1);
        cross3(a0, a1, K2);
        const double s = -2.0 * vv * p;  // d(detA)/d(a_s) = p * cof column s
#pragma unroll
        for (int r = 0; r < 3; r++) {
          dM[r][0] = s * K0[r];
          dM[r][1] = s * K1[r];
          dM[r][2] = s * K2[r];
        }
      }
    }
  } else {
    if (detA > 0.0) {
      double C0[3], C1[3], C2[3];
      cross3(w1, w2, C0);
      cross3(w2, w0, C1);
      cross3(w0, w1, C2);
      const double detWr = dot3(w0, C0);
      double Tp[3][3];
      double y = 0.0;
#pragma unroll
      for (int r = 0; r < 3; r++)
#pragma unroll
        for (int s = 0; s < 3; s++) {
          Tp[r][s] = a0[r] * C0[s] + a1[r] * C1[s] + a2[r] * C2[s];
          y += Tp[r][s] * Tp[r][s];
        }
      const double P = FAC_3SQRT3 * fabs(detWr) * detA;
      const double rr = fast_rsqrt(y * P * P);
      const double yr = y * rr;
      const double q = y * yr;
      val = q - p * detWr;
      if (GRAD) {
        have = true;
        cross3(a2, a0, K1);
        cross3(a0, a1, K2);
        const double k1 = 3.0 * yr;
        const double nk2 = -q * fast_rcp(detAr);  // q/detA_raw: the raw-frame cofactors carry the same sign p
#pragma unroll
        for (int r = 0; r < 3; r++) {
          const double (&t)[3] = Tp[r];
          dM[r][0] = k1 * (t[0] * C0[0] + t[1] * C0[1] + t[2] * C0[2]) + nk2 * K0[r];
          dM[r][1] = k1 * (t[0] * C1[0] + t[1] * C1[1] + t[2] * C1[2]) + nk2 * K1[r];
          dM[r][2] = k1 * (t[0] * C2[0] + t[1] * C2[1] + t[2] * C2[2]) + nk2 * K2[r];
        }
      }
    }
  }
  return val;
}

// Exodus/Shards hex8 local node a  <->  structured numbering (BlockStructuredGrid.cpp:20 permute_to_unstructured)
PMS_DI constexpr int hex_perm(int a) { return a == 2 ? 3 : (a == 3 ? 2 : (a == 6 ? 7 : (a == 7 ? 6 : a))); }

}  // namespace pmsdev
