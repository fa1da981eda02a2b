// pms_device.cuh — per-element Jacobian metrics and analytic gradients (device code).
//
// Restates, for sm_100a, the arithmetic of
//   SM/sgrid_metric_helpers.hpp:37-208      (3x3 helpers, sGridJacobianUtil, grad_util)
//   SM/SmootherMetric.hpp:602-968           (HexMeshSmootherMetric, structured blocks)
//   SM/SmootherMetric.hpp:970-1052,1207-1448 + SM/SmootherMetric.cpp:33-58
//                                           (SmootherMetricUntangle/ShapeB1, STK hex8 / tet4)
//   SM/JacobianUtil.hpp:133-205, JacobianUtil.cpp:32-35,162-243,388-396,448-457
// (SM/ = /root/reference/src/percept/mesh/mod/smoother/).
//
// Expression order follows the reference exactly, so that the translation unit compiled with
// -fmad=false ("strict_fp") is bit-comparable with a -ffp-contract=off CPU build, and the
// default build differs only by FMA contraction.  Everything is fully unrolled over corners with
// compile-time corner tables so vertices, Jacobians and gradients stay in registers.
#pragma once
#include <cstdint>

namespace pmsdev {

// element "flavors": which reference code path the element follows
enum : int {
  FL_SGRID = 0,  // structured block cell: node = i+2j+4k, sgrid locs table, mutator product order 0,1,2
  FL_HEX8 = 1,   // STK hex8 in Exodus/Shards order, DenseMatrix operator* order 0,2,1
  FL_TET4 = 2    // STK tet4, one Jacobian referenced to the equilateral tet, replicated 4x
};

#define PMS_DI __device__ __forceinline__

constexpr double BETA_MULT = 0.05;                          // SmootherMetric.hpp:609
constexpr double FAC_3SQRT3 = 0x1.4c8dc2e423980p+2;         // 3.0*std::sqrt(3.0) evaluated in double = 5.196152422706632
constexpr double ISQRT3 = 5.77350269189625797959429519858e-01;  // JacobianUtil.hpp:167
constexpr double ISQRT6 = 4.08248290463863052509822647505e-01;  // JacobianUtil.hpp:171

// corner -> (x0,x1,x2,x3): sgrid_metric_helpers.hpp:67-69 and JacobianUtil.cpp:32-35
template <int FL>
PMS_DI constexpr int loc(int c, int j) {
  constexpr int S[8][4] = {{0, 1, 2, 4}, {1, 3, 0, 5}, {2, 0, 3, 6}, {3, 2, 1, 7},
                           {4, 6, 5, 0}, {5, 4, 7, 1}, {6, 7, 4, 2}, {7, 5, 6, 3}};
  constexpr int H[8][4] = {{0, 1, 3, 4}, {1, 2, 0, 5}, {2, 3, 1, 6}, {3, 0, 2, 7},
                           {4, 7, 5, 0}, {5, 4, 6, 1}, {6, 5, 7, 2}, {7, 6, 4, 3}};
  return FL == FL_SGRID ? S[c][j] : (FL == FL_HEX8 ? H[c][j] : j);
}

PMS_DI double det3(const double (&m)[3][3]) {  // sgrid_metric_helpers.hpp:38-42
  return m[0][0] * (m[1][1] * m[2][2] - m[2][1] * m[1][2]) + m[0][1] * (m[2][0] * m[1][2] - m[1][0] * m[2][2]) +
         m[0][2] * (m[1][0] * m[2][1] - m[2][0] * m[1][1]);
}

// Jacobian of corner c: hex = [x1-x0 | x2-x0 | x3-x0]; tet = jacobian_matrix_tet_3D
template <int FL>
PMS_DI void corner_jacobian(const double (&v)[8][3], int c, double (&A)[3][3], double& detJ) {
  const int i0 = loc<FL>(c, 0), i1 = loc<FL>(c, 1), i2 = loc<FL>(c, 2), i3 = loc<FL>(c, 3);
  if (FL == FL_TET4) {  // JacobianUtil.hpp:175-205
#pragma unroll
    for (int r = 0; r < 3; r++) {
      const double f = v[i1][r] + v[i0][r];
      A[r][0] = v[i1][r] - v[i0][r];
      A[r][1] = (2.0 * v[i2][r] - f) * ISQRT3;
      A[r][2] = (3.0 * v[i3][r] - v[i2][r] - f) * ISQRT6;
    }
  } else {  // sgrid_metric_helpers.hpp:45-62
#pragma unroll
    for (int r = 0; r < 3; r++) {
      A[r][0] = v[i1][r] - v[i0][r];
      A[r][1] = v[i2][r] - v[i0][r];
      A[r][2] = v[i3][r] - v[i0][r];
    }
  }
  detJ = det3(A);
}

PMS_DI void matrix_inverse(const double (&m)[3][3], const double detm, double (&r)[3][3]) {  // :85-99
  const double detInv = 1.0 / detm;
  r[0][0] = (m[1][1] * m[2][2] - m[1][2] * m[2][1]) * detInv;
  r[0][1] = (m[0][2] * m[2][1] - m[0][1] * m[2][2]) * detInv;
  r[0][2] = (m[0][1] * m[1][2] - m[0][2] * m[1][1]) * detInv;
  r[1][0] = (m[1][2] * m[2][0] - m[1][0] * m[2][2]) * detInv;
  r[1][1] = (m[0][0] * m[2][2] - m[0][2] * m[2][0]) * detInv;
  r[1][2] = (m[0][2] * m[1][0] - m[0][0] * m[1][2]) * detInv;
  r[2][0] = (m[1][0] * m[2][1] - m[1][1] * m[2][0]) * detInv;
  r[2][1] = (m[0][1] * m[2][0] - m[0][0] * m[2][1]) * detInv;
  r[2][2] = (m[0][0] * m[1][1] - m[0][1] * m[1][0]) * detInv;
}

PMS_DI void matrix_product(const double (&x)[3][3], const double (&y)[3][3], double (&z)[3][3]) {  // :102-117
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) z[i][j] = x[i][0] * y[0][j] + x[i][1] * y[1][j] + x[i][2] * y[2][j];
}

PMS_DI double sqr_frobenius(const double (&m)[3][3]) {  // :142-158
  double sum = 0.0;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) sum += m[i][j] * m[i][j];
  return sum;
}

PMS_DI void transpose_adj(const double (&m)[3][3], double (&r)[3][3]) {  // :161-174 (scalar applied by caller)
  r[0][0] = (m[1][1] * m[2][2] - m[1][2] * m[2][1]);
  r[0][1] = (m[1][2] * m[2][0] - m[1][0] * m[2][2]);
  r[0][2] = (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
  r[1][0] = (m[0][2] * m[2][1] - m[0][1] * m[2][2]);
  r[1][1] = (m[0][0] * m[2][2] - m[0][2] * m[2][0]);
  r[1][2] = (m[0][1] * m[2][0] - m[0][0] * m[2][1]);
  r[2][0] = (m[0][1] * m[1][2] - m[0][2] * m[1][1]);
  r[2][1] = (m[0][2] * m[1][0] - m[0][0] * m[1][2]);
  r[2][2] = (m[0][0] * m[1][1] - m[0][1] * m[1][0]);
}

template <int FL>
__host__ __device__ constexpr int npe() { return FL == FL_TET4 ? 4 : 8; }
// number of distinct corner Jacobians evaluated (tet4 replicates one, JacobianUtil.cpp:388-396)
template <int FL>
__host__ __device__ constexpr int ncorner_eval() { return FL == FL_TET4 ? 1 : 8; }

// ShapeB1 value of one corner given A, W and their determinants (detA > 0).
//   SmootherMetric.hpp:665-712 / :1269-1310.  Returns shape_metric*detWiC; T, WI, d, f are outputs for the gradient.
template <bool USE_REF>
PMS_DI double shapeb1_corner(const double (&A)[3][3], const double (&W)[3][3], const double detA, const double detW,
                             double (&T)[3][3], double (&WI)[3][3], double& d, double& f, double& den) {
  double detWi = detW;
  if (USE_REF) {
    matrix_inverse(W, detWi, WI);
    matrix_product(A, WI, T);
  } else {
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
      for (int b = 0; b < 3; b++) T[a][b] = A[a][b];
    detWi = 1.0;
  }
  d = detA / detWi;
  f = sqrt(sqr_frobenius(T));
  den = FAC_3SQRT3 * d;
  const double shape_metric = (f * f * f) / den - 1.0;
  return shape_metric * detW;
}

// ---------------------------------------------------------------------------------------------
// element metric:  STAGE 0 = untangle, 1 = ShapeB1.  vc = current, vo = reference vertices.
// ---------------------------------------------------------------------------------------------
template <int FL, int STAGE, bool USE_REF>
PMS_DI double element_metric(const double (&vc)[8][3], const double (&vo)[8][3], bool& valid) {
  valid = true;
  double val = 0.0;
  constexpr int NC = ncorner_eval<FL>();
  constexpr int NN = npe<FL>();
#pragma unroll
  for (int c = 0; c < NC; c++) {
    double A[3][3], W[3][3], detA, detW;
    corner_jacobian<FL>(vc, c, A, detA);
    corner_jacobian<FL>(vo, c, W, detW);
    if (detA <= 0.0) valid = false;
    double m;
    if (STAGE == 0) {
      double vv = BETA_MULT * detW - detA;
      vv = (vv < 0.0 ? 0.0 : vv);
      m = vv * vv;
    } else {
      m = 0.0;
      if (detA > 0.0) {
        double T[3][3], WI[3][3], d, f, den;
        m = shapeb1_corner<USE_REF>(A, W, detA, detW, T, WI, d, f, den);
      }
    }
    if (FL == FL_TET4) {
      // m_detJ[i] = m, m_J[i] = m_J[0] for i<4: the same corner value is accumulated 4 times
#pragma unroll
      for (int k = 0; k < NN; k++) val += m;
    } else {
      val += m;
    }
  }
  return val;
}

// ---------------------------------------------------------------------------------------------
// tet4 with the reference-mesh part cached per element: a tet has ONE corner Jacobian (JacobianUtil.cpp:388-396), so
// W^-1 (9 values) and det W depend on coordinates_NM1 only and are computed once per run by k_tet_ref — the very
// operations of corner_jacobian + matrix_inverse below — instead of gathering the 4 reference vertices (12 scattered
// 8-byte loads per tet and pass; the tet kernels are L1/LSU-bound on exactly those gathers).  ref[v/3][v%3], v = 0..8:
// WI row-major, v = 9: det W.  Same expression order as element_metric / element_grad_metric<FL_TET4, STAGE, true>.
// ---------------------------------------------------------------------------------------------
PMS_DI void tet_reference(const double (&vo)[8][3], double (&ref)[10]) {
  double W[3][3], detW, WI[3][3];
  corner_jacobian<FL_TET4>(vo, 0, W, detW);
  matrix_inverse(W, detW, WI);
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int b = 0; b < 3; b++) ref[a * 3 + b] = WI[a][b];
  ref[9] = detW;
}

template <int STAGE, bool GRAD>
PMS_DI double tet_metric_cached(const double (&vc)[8][3], const double (&ref)[8][3], bool& valid, double (&grad)[8][3]) {
  valid = true;
  double A[3][3], detA;
  corner_jacobian<FL_TET4>(vc, 0, A, detA);
  const double detW = ref[3][0];
  if (detA <= 0.0) valid = false;
  if (GRAD) {
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int r = 0; r < 3; r++) grad[a][r] = 0.0;
  }
  double dM[3][3];
  double m = 0.0;
  bool have = false;
  if (STAGE == 0) {
    const double vv = BETA_MULT * detW - detA;
    const double vv1 = (vv > 0.0 ? vv : 0.0);
    m = vv1 * vv1;
    if (GRAD && vv > 0.0) {
      have = true;
      double TA[3][3];
      transpose_adj(A, TA);
      const double s = -2.0 * vv;
#pragma unroll
      for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) dM[a][b] = s * TA[a][b];
    }
  } else if (detA > 0.0) {
    have = true;
    double WI[3][3], T[3][3];
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
      for (int b = 0; b < 3; b++) WI[a][b] = ref[a][b];
    matrix_product(A, WI, T);
    const double d = detA / detW;
    const double f = sqrt(sqr_frobenius(T));
    const double den = FAC_3SQRT3 * d;
    const double shape_metric = (f * f * f) / den - 1.0;
    m = shape_metric * detW;
    if (GRAD) {
      const double norm = f;
      const double iden = 1.0 / den;
      const double s1 = (3 * norm * iden);
#pragma unroll
      for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) dM[a][b] = s1 * T[a][b];
      const double norm_cube = norm * norm * norm;
      double TAT[3][3];
      transpose_adj(T, TAT);
      const double s2 = (norm_cube * iden / d);
#pragma unroll
      for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) dM[a][b] -= s2 * TAT[a][b];
      double tmp[3][3];
#pragma unroll
      for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) {  // DenseMatrix operator*, DenseMatrix.hpp:502-533 (k = 0,2,1)
          double t = dM[a][0] * WI[b][0];
          t += dM[a][2] * WI[b][2];
          t += dM[a][1] * WI[b][1];
          tmp[a][b] = t;
        }
#pragma unroll
      for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) dM[a][b] = tmp[a][b] * detW;
    }
  }
  double val = 0.0;
#pragma unroll
  for (int k = 0; k < 4; k++) val += m;
  if (GRAD && have) {  // grad_util_tet_3d, JacobianUtil.cpp:215-243, then grad = ((G+G)+G)+G over the 4 identical corners
#pragma unroll
    for (int r = 0; r < 3; r++) {
      double G0 = 0.0, G1 = 0.0, G2 = 0.0, G3 = 0.0;
      G1 += dM[r][0] * (+1);
      G0 += dM[r][0] * (-1);
      G2 += dM[r][1] * (+2.0 * ISQRT3);
      G1 += dM[r][1] * (-1.0 * ISQRT3);
      G0 += dM[r][1] * (-1.0 * ISQRT3);
      G3 += dM[r][2] * (+3.0 * ISQRT6);
      G2 += dM[r][2] * (-1.0 * ISQRT6);
      G1 += dM[r][2] * (-1.0 * ISQRT6);
      G0 += dM[r][2] * (-1.0 * ISQRT6);
      grad[0][r] = ((G0 + G0) + G0) + G0;
      grad[1][r] = ((G1 + G1) + G1) + G1;
      grad[2][r] = ((G2 + G2) + G2) + G2;
      grad[3][r] = ((G3 + G3) + G3) + G3;
    }
  }
  return val;
}

// ---------------------------------------------------------------------------------------------
// element metric + gradient.  grad[a][r] = d(metric)/d(x_a,r) following grad_metric + grad_util;
// returns the metric value as total_metric would compute it (scaled by detW for ShapeB1).
// In smoothing mode corners with detA<=0 contribute nothing and the caller drops the whole
// element's gradient when !valid (gradient_functors.hpp:391, Def.hpp:1219).
// ---------------------------------------------------------------------------------------------
template <int FL, int STAGE, bool USE_REF>
PMS_DI double element_grad_metric(const double (&vc)[8][3], const double (&vo)[8][3], bool& valid,
                                  double (&grad)[8][3]) {
  valid = true;
  double val = 0.0;
  constexpr int NC = ncorner_eval<FL>();
  constexpr int NN = npe<FL>();
#pragma unroll
  for (int a = 0; a < NN; a++)
#pragma unroll
    for (int r = 0; r < 3; r++) grad[a][r] = 0.0;

#pragma unroll
  for (int c = 0; c < NC; c++) {
    double A[3][3], W[3][3], detA, detW;
    corner_jacobian<FL>(vc, c, A, detA);
    corner_jacobian<FL>(vo, c, W, detW);
    if (detA <= 0.0) valid = false;
    double dM[3][3];
    double m = 0.0;
    bool have = false;
    if (STAGE == 0) {
      const double vv = BETA_MULT * detW - detA;
      const double vv1 = (vv > 0.0 ? vv : 0.0);
      m = vv1 * vv1;
      if (vv > 0.0) {
        have = true;
        double TA[3][3];
        transpose_adj(A, TA);
        const double s = -2.0 * vv;
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
          for (int b = 0; b < 3; b++) dM[a][b] = s * TA[a][b];
      }
    } else {
      if (detA > 0.0) {
        have = true;
        double T[3][3], WI[3][3], d, f, den;
        m = shapeb1_corner<USE_REF>(A, W, detA, detW, T, WI, d, f, den);
        const double norm = f;
        const double iden = 1.0 / den;
        const double s1 = (3 * norm * iden);
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
          for (int b = 0; b < 3; b++) dM[a][b] = s1 * T[a][b];
        const double norm_cube = norm * norm * norm;
        double TAT[3][3];
        transpose_adj(T, TAT);
        const double s2 = (norm_cube * iden / d);
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
          for (int b = 0; b < 3; b++) dM[a][b] -= s2 * TAT[a][b];
        if (USE_REF) {
          double tmp[3][3];
#pragma unroll
          for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++) {
              // dM * WI^T : WIT[k][b] = WI[b][k]
              if (FL == FL_SGRID)  // matrix_product_mutator, sgrid_metric_helpers.hpp:120-139 (k = 0,1,2)
                tmp[a][b] = dM[a][0] * WI[b][0] + dM[a][1] * WI[b][1] + dM[a][2] * WI[b][2];
              else {  // DenseMatrix operator*, DenseMatrix.hpp:502-533 (k = 0,2,1)
                double t = dM[a][0] * WI[b][0];
                t += dM[a][2] * WI[b][2];
                t += dM[a][1] * WI[b][1];
                tmp[a][b] = t;
              }
            }
#pragma unroll
          for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++) dM[a][b] = tmp[a][b];
        }
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
          for (int b = 0; b < 3; b++) dM[a][b] = dM[a][b] * detW;
      }
    }
    if (FL == FL_TET4) {
#pragma unroll
      for (int k = 0; k < NN; k++) val += m;
      if (have) {
        // grad_util_tet_3d, JacobianUtil.cpp:215-243, then grad = ((G+G)+G)+G over the 4 identical corners
#pragma unroll
        for (int r = 0; r < 3; r++) {
          double G0 = 0.0, G1 = 0.0, G2 = 0.0, G3 = 0.0;
          G1 += dM[r][0] * (+1);
          G0 += dM[r][0] * (-1);
          G2 += dM[r][1] * (+2.0 * ISQRT3);
          G1 += dM[r][1] * (-1.0 * ISQRT3);
          G0 += dM[r][1] * (-1.0 * ISQRT3);
          G3 += dM[r][2] * (+3.0 * ISQRT6);
          G2 += dM[r][2] * (-1.0 * ISQRT6);
          G1 += dM[r][2] * (-1.0 * ISQRT6);
          G0 += dM[r][2] * (-1.0 * ISQRT6);
          grad[0][r] = ((G0 + G0) + G0) + G0;
          grad[1][r] = ((G1 + G1) + G1) + G1;
          grad[2][r] = ((G2 + G2) + G2) + G2;
          grad[3][r] = ((G3 + G3) + G3) + G3;
        }
      }
    } else {
      val += m;
      if (have) {
        // grad_util (sgrid_metric_helpers.hpp:178-196): +col_m to node idx[m+1], -(col0+col1+col2) to idx[0];
        // corners are combined in ascending corner order (SmootherMetric.hpp:958-961).
        const int i0 = loc<FL>(c, 0), i1 = loc<FL>(c, 1), i2 = loc<FL>(c, 2), i3 = loc<FL>(c, 3);
#pragma unroll
        for (int r = 0; r < 3; r++) {
          double g0 = -dM[r][0];
          g0 += -dM[r][1];
          g0 += -dM[r][2];
          grad[i0][r] += g0;
          grad[i1][r] += dM[r][0];
          grad[i2][r] += dM[r][1];
          grad[i3][r] += dM[r][2];
        }
      }
    }
  }
  return val;
}

}  // namespace pmsdev
