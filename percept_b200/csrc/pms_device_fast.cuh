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

namespace pmsdev {

PMS_DI double fast_rsqrt(double x) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double hx = 0.5 * x;
#pragma unroll
  for (int it = 0; it < 2; it++) {
    const double h = hx * r;
    const double e = fma(-h, r, 0.5);
    r = fma(r, e, r);
  }
  return r;
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

// canonical edge along axis AX touching corner c (vertices in structured numbering node = i+2j+4k)
template <int AX>
PMS_DI void corner_edge(const double (&v)[8][3], int c, double (&e)[3]) {
  const int lo = c & ~(1 << AX), hi = lo | (1 << AX);
#pragma unroll
  for (int r = 0; r < 3; r++) e[r] = v[hi][r] - v[lo][r];
}

// Hex element metric (+ gradient when GRAD) in structured vertex numbering.  STAGE 0 untangle, 1 ShapeB1 with
// reference mesh.  grad[a][r] accumulates d(metric)/d(x_a,r); the caller applies the element-level validity rule.
template <int STAGE, bool GRAD>
PMS_DI double hex_metric_fast(const double (&vc)[8][3], const double (&vo)[8][3], bool& valid, double (&grad)[8][3]) {
  valid = true;
  double val = 0.0;
  if (GRAD) {
#pragma unroll
    for (int a = 0; a < 8; a++)
#pragma unroll
      for (int r = 0; r < 3; r++) grad[a][r] = 0.0;
  }
#pragma unroll
  for (int c = 0; c < 8; c++) {
    double a0[3], a1[3], a2[3], w0[3], w1[3], w2[3];
    corner_edge<0>(vc, c, a0);
    corner_edge<1>(vc, c, a1);
    corner_edge<2>(vc, c, a2);
    corner_edge<0>(vo, c, w0);
    corner_edge<1>(vo, c, w1);
    corner_edge<2>(vo, c, w2);
    double K0[3], K1[3], K2[3];
    cross3(a1, a2, K0);
    const double detA = dot3(a0, K0);
    if (detA <= 0.0) valid = false;
    double dM[3][3];  // dM[r][s]: derivative w.r.t. component r of the edge along axis s
    bool have = false;
    if (STAGE == 0) {
      double C0[3];
      cross3(w1, w2, C0);
      const double detW = dot3(w0, C0);
      const double vv = BETA_MULT * detW - detA;
      if (vv > 0.0) {
        val += vv * vv;
        if (GRAD) {
          have = true;
          cross3(a2, a0, K1);
          cross3(a0, a1, K2);
          const double s = -2.0 * vv;
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
        const double detW = dot3(w0, C0);
        double Tp[3][3];
        double y = 0.0;
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
          for (int s = 0; s < 3; s++) {
            Tp[r][s] = a0[r] * C0[s] + a1[r] * C1[s] + a2[r] * C2[s];
            y += Tp[r][s] * Tp[r][s];
          }
        const double P = FAC_3SQRT3 * fabs(detW) * detA;
        const double rr = fast_rsqrt(y * P * P);
        const double yr = y * rr;  // sqrt(y)/P
        const double q = y * yr;   // y^1.5/P
        val += q - detW;
        if (GRAD) {
          have = true;
          cross3(a2, a0, K1);
          cross3(a0, a1, K2);
          const double k1 = 3.0 * yr;
          const double nk2 = -q * fast_rcp(detA);
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
    if (GRAD && have) {
#pragma unroll
      for (int s = 0; s < 3; s++) {
        const int lo = c & ~(1 << s), hi = lo | (1 << s);
#pragma unroll
        for (int r = 0; r < 3; r++) {
          grad[hi][r] += dM[r][s];
          grad[lo][r] -= dM[r][s];
        }
      }
    }
  }
  return val;
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
