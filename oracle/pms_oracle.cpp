// pms_oracle.cpp — CPU restatement of Percept's reference-mesh CG smoother hot path.
//
// *** TEST INFRASTRUCTURE ONLY. ***  Nothing under percept_b200/ may include, link
// or call this file.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs use it, as the checker / CPU baseline.
//
// Parity status: PINNED by the reference's own golden values
//   test/regression_tests/HeavyTestMeshSmoother.cpp:1352-1358 (nele=100/200/300:
//   n_invalid, mtot, |g|_untangle, |g|_smooth) and :809-810 (bump mesh nele=5),
//   see tests/test_oracle_golden.py.  tet4 has no golden in the reference; it is
//   cross-checked against central finite differences (the reference's own dormant
//   check, ReferenceMeshSmootherConjugateGradientDef.hpp:1252-1335).
//   PARITY UNPINNED (no golden vector, known-answer test or fixture in the reference's tests; pinned through properties
//   instead, see the test files): the StructuredGridRefiner restatement (orc_refine_structured_*, tests/test_refine.py)
//   and the smooth_surfaces hooks with analytic evaluators (orc_add_analytic_surface, orc_set_node_classification,
//   orc_snap_nodes, orc_get_surface_normals; tests/test_surface_hooks.py - the reference evaluates third-party CAD).
//
// All file:line citations are relative to /root/reference/ ;
//   SM/ = src/percept/mesh/mod/smoother/ , Def.hpp = SM/ReferenceMeshSmootherConjugateGradientDef.hpp
//
// Two accumulation modes, selected with orc_set_option("real_is_double", v):
//   0 (default): Double = long double  — the reference CPU build (MeshType.hpp:41-46)
//   1          : Double = double       — the reference KOKKOS_ENABLE_CUDA build
// Build with -O2 -ffp-contract=off (no FMA, like the reference's plain CXX flags).
// Exports the same C ABI as include/percept_smoother_c.h with prefix orc_.

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <limits>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

typedef long double LD;

// ---------------------------------------------------------------------------------
// 3x3 algebra — SM/sgrid_metric_helpers.hpp:37-174, math/DenseMatrix.hpp:567-970
// ---------------------------------------------------------------------------------
typedef double M33[3][3];

inline double det3(const M33 m) {  // sgrid_metric_helpers.hpp:38-42 == DenseMatrix.hpp:567-571
  return m[0][0] * (m[1][1] * m[2][2] - m[2][1] * m[1][2]) + m[0][1] * (m[2][0] * m[1][2] - m[1][0] * m[2][2]) +
         m[0][2] * (m[1][0] * m[2][1] - m[2][0] * m[1][1]);
}

inline void jacobian_matrix_3D(double& detJ, M33 A, const double* x0, const double* x1, const double* x2,
                               const double* x3) {  // sgrid_metric_helpers.hpp:45-62, JacobianUtil.hpp:133-150
  A[0][0] = (x1[0] - x0[0]);
  A[0][1] = (x2[0] - x0[0]);
  A[0][2] = (x3[0] - x0[0]);
  A[1][0] = (x1[1] - x0[1]);
  A[1][1] = (x2[1] - x0[1]);
  A[1][2] = (x3[1] - x0[1]);
  A[2][0] = (x1[2] - x0[2]);
  A[2][1] = (x2[2] - x0[2]);
  A[2][2] = (x3[2] - x0[2]);
  detJ = det3(A);
}

#define ISQRT3 5.77350269189625797959429519858e-01   /* JacobianUtil.hpp:167 */
#define ISQRT6 4.08248290463863052509822647505e-01   /* JacobianUtil.hpp:171 */

inline void jacobian_matrix_tet_3D(double& detJ, M33 A, const double* x0, const double* x1, const double* x2,
                                   const double* x3) {  // JacobianUtil.hpp:175-205
  double* matr = &A[0][0];
  double f;
  f = x1[0] + x0[0];
  matr[0] = x1[0] - x0[0];
  matr[1] = (2.0 * x2[0] - f) * ISQRT3;
  matr[2] = (3.0 * x3[0] - x2[0] - f) * ISQRT6;
  f = x1[1] + x0[1];
  matr[3] = x1[1] - x0[1];
  matr[4] = (2.0 * x2[1] - f) * ISQRT3;
  matr[5] = (3.0 * x3[1] - x2[1] - f) * ISQRT6;
  f = x1[2] + x0[2];
  matr[6] = x1[2] - x0[2];
  matr[7] = (2.0 * x2[2] - f) * ISQRT3;
  matr[8] = (3.0 * x3[2] - x2[2] - f) * ISQRT6;
  detJ = det3(A);
}

inline void matrix_inverse(const M33 m, const double detm, M33 r) {  // sgrid_metric_helpers.hpp:85-99, DenseMatrix.hpp:938-952
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

inline void matrix_product(const M33 x, const M33 y, M33 z) {  // sgrid_metric_helpers.hpp:102-117, DenseMatrix.hpp:955-970
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) z[i][j] = x[i][0] * y[0][j] + x[i][1] * y[1][j] + x[i][2] * y[2][j];
}

// DenseMatrix operator* (DenseMatrix.hpp:502-533): the fall-through switch adds k=0, then k=2, then k=1.
inline void matrix_product_densematrix_op(const M33 x, const M33 y, M33 z) {
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) {
      double tmp = x[r][0] * y[0][c];
      tmp += x[r][2] * y[2][c];
      tmp += x[r][1] * y[1][c];
      z[r][c] = tmp;
    }
}

inline double sqr_frobenius(const M33 m) {  // sgrid_metric_helpers.hpp:142-158, DenseMatrix.hpp:889-906
  double sum = 0.0;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) sum += m[i][j] * m[i][j];
  return sum;
}

inline void transpose_adj(const M33 m, M33 r, double scalar = 1.0) {  // sgrid_metric_helpers.hpp:161-174, DenseMatrix.hpp:615-649
  r[0][0] = scalar * (m[1][1] * m[2][2] - m[1][2] * m[2][1]);
  r[0][1] = scalar * (m[1][2] * m[2][0] - m[1][0] * m[2][2]);
  r[0][2] = scalar * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
  r[1][0] = scalar * (m[0][2] * m[2][1] - m[0][1] * m[2][2]);
  r[1][1] = scalar * (m[0][0] * m[2][2] - m[0][2] * m[2][0]);
  r[1][2] = scalar * (m[0][1] * m[2][0] - m[0][0] * m[2][1]);
  r[2][0] = scalar * (m[0][1] * m[1][2] - m[0][2] * m[1][1]);
  r[2][1] = scalar * (m[0][2] * m[1][0] - m[0][0] * m[1][2]);
  r[2][2] = scalar * (m[0][0] * m[1][1] - m[0][1] * m[1][0]);
}

// corner -> (x0,x1,x2,x3) tables
static const int LOCS_SGRID[8][4] = {{0, 1, 2, 4}, {1, 3, 0, 5}, {2, 0, 3, 6}, {3, 2, 1, 7},
                                     {4, 6, 5, 0}, {5, 4, 7, 1}, {6, 7, 4, 2}, {7, 5, 6, 3}};  // sgrid_metric_helpers.hpp:67-69
static const int LOCS_HEX[8][4] = {{0, 1, 3, 4}, {1, 2, 0, 5}, {2, 3, 1, 6}, {3, 0, 2, 7},
                                   {4, 7, 5, 0}, {5, 4, 6, 1}, {6, 5, 7, 2}, {7, 6, 4, 3}};  // JacobianUtil.cpp:32-35

// grad_util for hex corners (sgrid_metric_helpers.hpp:178-196, JacobianUtil.cpp:162-180)
inline void grad_util_hex(const M33 dMdA, double grad[8][3], const int* ind) {
  for (int i = 0; i < 8; i++)
    for (int j = 0; j < 3; j++) grad[i][j] = 0.0;
  for (int r = 0; r < 3; r++) {
    grad[ind[1]][r] += dMdA[r][0] * (+1);
    grad[ind[0]][r] += dMdA[r][0] * (-1);
    grad[ind[2]][r] += dMdA[r][1] * (+1);
    grad[ind[0]][r] += dMdA[r][1] * (-1);
    grad[ind[3]][r] += dMdA[r][2] * (+1);
    grad[ind[0]][r] += dMdA[r][2] * (-1);
  }
}

// grad_util_tet_3d (JacobianUtil.cpp:215-243), indices_tet = {0,1,2,3}
inline void grad_util_tet(const M33 dMdA, double grad[4][3]) {
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 3; j++) grad[i][j] = 0.0;
  for (int r = 0; r < 3; r++) {
    grad[1][r] += dMdA[r][0] * (+1);
    grad[0][r] += dMdA[r][0] * (-1);
    grad[2][r] += dMdA[r][1] * (+2.0 * ISQRT3);
    grad[1][r] += dMdA[r][1] * (-1.0 * ISQRT3);
    grad[0][r] += dMdA[r][1] * (-1.0 * ISQRT3);
    grad[3][r] += dMdA[r][2] * (+3.0 * ISQRT6);
    grad[2][r] += dMdA[r][2] * (-1.0 * ISQRT6);
    grad[1][r] += dMdA[r][2] * (-1.0 * ISQRT6);
    grad[0][r] += dMdA[r][2] * (-1.0 * ISQRT6);
  }
}

const double BETA_MULT = 0.05;  // SmootherMetric.hpp:609, SmootherMetric.cpp:27

// ---------------------------------------------------------------------------------
// HexMeshSmootherMetric (structured) — SM/SmootherMetric.hpp:602-968
// ---------------------------------------------------------------------------------
struct HexMeshSmootherMetric {
  bool m_use_ref_mesh = true;
  bool m_untangling = true;

  double metric(double v_cur[8][3], double v_org[8][3], bool& valid) const {  // :613-720
    valid = true;
    double nodal_A[8], nodal_W[8];
    M33 J_A[8], J_W[8];
    for (int i = 0; i < 8; i++) {
      const int* l = LOCS_SGRID[i];
      jacobian_matrix_3D(nodal_A[i], J_A[i], v_cur[l[0]], v_cur[l[1]], v_cur[l[2]], v_cur[l[3]]);
      jacobian_matrix_3D(nodal_W[i], J_W[i], v_org[l[0]], v_org[l[1]], v_org[l[2]], v_org[l[3]]);
    }
    if (m_untangling) {
      double val_untangle = 0.0;
      for (int i = 0; i < 8; i++) {
        double detAi = nodal_A[i], detWi = nodal_W[i];
        if (detAi <= 0.) valid = false;
        double vv = BETA_MULT * detWi - detAi;
        (vv < 0.0 ? vv = 0.0 : vv);
        val_untangle += vv * vv;
      }
      return val_untangle;
    }
    double val_shape = 0.0;
    M33 T, WI;
    for (int i = 0; i < 8; i++) {
      const double detAi = nodal_A[i];
      double detWi = nodal_W[i];
      const double detWiC = detWi;
      if (detAi <= 0) valid = false;
      double shape_metric = 0.0;
      if (detAi > 0.0) {
        if (m_use_ref_mesh) {
          matrix_inverse(J_W[i], detWi, WI);
          matrix_product(J_A[i], WI, T);
        } else {
          for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) T[a][b] = J_A[i][a][b];
          detWi = 1.0;
        }
        double d = detAi / detWi;
        double f = sqr_frobenius(T);
        f = std::sqrt(f);
        const double fac = 3.0 * std::sqrt(3.0);
        double den = fac * d;
        shape_metric = (f * f * f) / den - 1.0;
        shape_metric = shape_metric * detWiC;
      }
      val_shape += shape_metric;
    }
    return val_shape;
  }

  // :722-967.  grad[8][3]; element-level `valid` decides (in the caller) whether grad is used.
  double grad_metric(double v_cur[8][3], double v_org[8][3], bool& valid, double grad[8][3]) const {
    valid = true;
    double nodal_A[8], nodal_W[8];
    M33 J_A[8], J_W[8];
    static thread_local double dMetric_dA[8][3][3];  // corners with detA<=0 keep stale values in smoothing mode (:835)
    for (int i = 0; i < 8; i++) {
      const int* l = LOCS_SGRID[i];
      jacobian_matrix_3D(nodal_A[i], J_A[i], v_cur[l[0]], v_cur[l[1]], v_cur[l[2]], v_cur[l[3]]);
      jacobian_matrix_3D(nodal_W[i], J_W[i], v_org[l[0]], v_org[l[1]], v_org[l[2]], v_org[l[3]]);
    }
    double val = 0.0;
    if (m_untangling) {
      for (int i = 0; i < 8; i++) {
        double detAi = nodal_A[i], detWi = nodal_W[i];
        if (detAi <= 0) valid = false;
        double vv = BETA_MULT * detWi - detAi;
        double vv1 = (vv > 0.0 ? vv : 0);
        if (vv > 0.0)
          transpose_adj(J_A[i], dMetric_dA[i], -2.0 * vv);
        else
          for (int j = 0; j < 3; j++)
            for (int k = 0; k < 3; k++) dMetric_dA[i][j][k] = 0.0;
        val += vv1 * vv1;
      }
    } else {
      M33 WIT, TAT, T, WI;
      for (int i = 0; i < 8; i++) {
        double detAi = nodal_A[i], detWi = nodal_W[i];
        double detWiC = detWi;
        if (detAi <= 0) valid = false;
        double shape_metric = 0.0;
        if (detAi > 0) {
          if (m_use_ref_mesh) {
            matrix_inverse(J_W[i], detWi, WI);
            matrix_product(J_A[i], WI, T);
          } else {
            for (int a = 0; a < 3; a++)
              for (int b = 0; b < 3; b++) T[a][b] = J_A[i][a][b];
            T[1][2] = J_A[i][1][0];  // sic, SmootherMetric.hpp:850
            detWi = 1.0;
            for (int a = 0; a < 3; a++)
              for (int b = 0; b < 3; b++) WI[a][b] = (a == b ? 1.0 : 0.0);
          }
          double d = detAi / detWi;
          double f = sqr_frobenius(T);
          f = std::sqrt(f);
          const double fac = 3.0 * std::sqrt(3.0);
          double den = fac * d;
          shape_metric = (f * f * f) / den - 1.0;
          double norm = f;
          double iden = 1.0 / den;
          for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) dMetric_dA[i][a][b] = (3 * norm * iden) * T[a][b];
          double norm_cube = norm * norm * norm;
          transpose_adj(T, TAT);
          for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) dMetric_dA[i][a][b] -= (norm_cube * iden / d) * TAT[a][b];
          if (m_use_ref_mesh) {
            for (int a = 0; a < 3; a++)
              for (int b = 0; b < 3; b++) WIT[a][b] = WI[b][a];
            M33 temp;  // matrix_product_mutator, sgrid_metric_helpers.hpp:120-139
            for (int a = 0; a < 3; a++)
              for (int b = 0; b < 3; b++)
                temp[a][b] = dMetric_dA[i][a][0] * WIT[0][b] + dMetric_dA[i][a][1] * WIT[1][b] + dMetric_dA[i][a][2] * WIT[2][b];
            for (int a = 0; a < 3; a++)
              for (int b = 0; b < 3; b++) dMetric_dA[i][a][b] = temp[a][b];
          }
          for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) dMetric_dA[i][a][b] = dMetric_dA[i][a][b] * detWiC;
        }
        val += shape_metric;
      }
    }
    double grads_fld[8][8][3];
    for (int i = 0; i < 8; i++) grad_util_hex(dMetric_dA[i], grads_fld[i], LOCS_SGRID[i]);
    for (int i = 0; i < 8; i++)
      for (int j = 0; j < 3; j++) grad[i][j] = 0.0;
    for (int k = 0; k < 8; k++)
      for (int i = 0; i < 8; i++)
        for (int j = 0; j < 3; j++) grad[i][j] += grads_fld[k][i][j];
    return val;
  }
};

// ---------------------------------------------------------------------------------
// STK metrics: JacobianUtilImpl<STKMesh>::operator() (JacobianUtil.cpp:316-499),
// SmootherMetricUntangleImpl (SmootherMetric.cpp:33-58, SmootherMetric.hpp:994-1050),
// SmootherMetricShapeB1Impl (SmootherMetric.hpp:1207-1448)
// ---------------------------------------------------------------------------------
struct JacUtil {
  int nn;
  double detJ[8];
  M33 J[8];
  M33 dMetric_dA[8];
  double grad[8][8][3];
  void jac(int topo, double v[8][3]) {
    if (topo == 8) {
      nn = 8;
      for (int i = 0; i < 8; i++) {
        const int* l = LOCS_HEX[i];
        jacobian_matrix_3D(detJ[i], J[i], v[l[0]], v[l[1]], v[l[2]], v[l[3]]);
      }
    } else {
      nn = 4;
      jacobian_matrix_tet_3D(detJ[0], J[0], v[0], v[1], v[2], v[3]);
      for (int i = 1; i < 4; i++) {
        detJ[i] = detJ[0];
        memcpy(J[i], J[0], sizeof(M33));
      }
    }
  }
  void grad_metric_util(int topo) {  // JacobianUtil.cpp:503-635
    if (topo == 8) {
      for (int i = 0; i < 8; i++) grad_util_hex(dMetric_dA[i], grad[i], LOCS_HEX[i]);
    } else {
      for (int i = 0; i < 4; i++) {
        double g4[4][3];
        grad_util_tet(dMetric_dA[i], g4);
        for (int a = 0; a < 4; a++)
          for (int b = 0; b < 3; b++) grad[i][a][b] = g4[a][b];
      }
    }
  }
};

double stk_untangle_metric(int topo, double vc[8][3], double vo[8][3], bool& valid) {  // SmootherMetric.cpp:33-58
  valid = true;
  JacUtil jacA, jacW;
  jacA.jac(topo, vc);
  jacW.jac(topo, vo);
  double val_untangle = 0.0;
  for (int i = 0; i < jacA.nn; i++) {
    double detAi = jacA.detJ[i], detWi = jacW.detJ[i];
    if (detAi <= 0.) valid = false;
    double vv = BETA_MULT * detWi - detAi;
    vv = std::max(vv, 0.0);
    val_untangle += vv * vv;
  }
  return val_untangle;
}

double stk_untangle_grad_metric(int topo, double vc[8][3], double vo[8][3], bool& valid, double grad[8][3]) {  // SmootherMetric.hpp:994-1050
  JacUtil jacA, jacW;
  jacA.jac(topo, vc);
  jacW.jac(topo, vo);
  double val = 0.0;
  const int nn = jacA.nn;
  for (int i = 0; i < nn; i++) {
    double detAi = jacA.detJ[i], detWi = jacW.detJ[i];
    if (detAi <= 0) valid = false;
    double vv = BETA_MULT * detWi - detAi;
    double vv1 = std::max(vv, 0.0);
    if (vv > 0.0) {
      // wrt_A = -2.0*vv*transpose_adj(A): scalar (-2.0*vv) times each entry
      transpose_adj(jacA.J[i], jacA.dMetric_dA[i], 1.0);
      const double s = -2.0 * vv;
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) jacA.dMetric_dA[i][a][b] = jacA.dMetric_dA[i][a][b] * s;
    } else
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) jacA.dMetric_dA[i][a][b] = 0.0;
    val += vv1 * vv1;
  }
  jacA.grad_metric_util(topo);
  for (int i = 0; i < nn; i++)
    for (int j = 0; j < 3; j++) grad[i][j] = 0.0;
  for (int k = 0; k < nn; k++)
    for (int i = 0; i < nn; i++)
      for (int j = 0; j < 3; j++) grad[i][j] += jacA.grad[k][i][j];
  return val;
}

double stk_shapeb1_metric(int topo, bool use_ref, double vc[8][3], double vo[8][3], bool& valid) {  // SmootherMetric.hpp:1224-1319
  valid = true;
  JacUtil jacA, jacW;
  jacA.jac(topo, vc);
  jacW.jac(topo, vo);
  double val_shape = 0.0;
  M33 WI, T;
  for (int i = 0; i < jacA.nn; i++) {
    const double detAi = jacA.detJ[i];
    double detWi = jacW.detJ[i];
    const double detWiC = detWi;
    if (detAi <= 0) valid = false;
    double shape_metric = 0.0;
    if (detAi > 0.0) {
      if (use_ref) {
        matrix_inverse(jacW.J[i], detWi, WI);
        matrix_product(jacA.J[i], WI, T);
      } else {
        memcpy(T, jacA.J[i], sizeof(M33));
        detWi = 1.0;
      }
      double d = detAi / detWi;
      double f = sqr_frobenius(T);
      f = std::sqrt(f);
      const double fac = 3.0 * std::sqrt(3.0);
      double den = fac * d;
      shape_metric = (f * f * f) / den - 1.0;
      shape_metric = shape_metric * detWiC;
    }
    val_shape += shape_metric;
  }
  return val_shape;
}

double stk_shapeb1_grad_metric(int topo, bool use_ref, double vc[8][3], double vo[8][3], bool& valid,
                               double grad[8][3]) {  // SmootherMetric.hpp:1322-1446
  JacUtil jacA, jacW;
  jacA.jac(topo, vc);
  jacW.jac(topo, vo);
  for (int i = 0; i < 8; i++) memset(jacA.dMetric_dA[i], 0, sizeof(M33));
  double val_shape = 0.0;
  M33 WI, T, WIT, TAT;
  const int nn = jacA.nn;
  for (int i = 0; i < nn; i++) {
    double detAi = jacA.detJ[i], detWi = jacW.detJ[i];
    double detWiC = detWi;
    if (detAi <= 0) valid = false;
    double shape_metric = 0.0;
    if (detAi > 0) {
      if (use_ref) {
        matrix_inverse(jacW.J[i], detWi, WI);
        matrix_product(jacA.J[i], WI, T);
      } else {
        memcpy(T, jacA.J[i], sizeof(M33));
        detWi = 1.0;
        for (int a = 0; a < 3; a++)
          for (int b = 0; b < 3; b++) WI[a][b] = (a == b ? 1.0 : 0.0);
      }
      double d = detAi / detWi;
      double f = sqr_frobenius(T);
      f = std::sqrt(f);
      const double fac = 3.0 * std::sqrt(3.0);
      double den = fac * d;
      shape_metric = (f * f * f) / den - 1.0;
      M33& wrt_A = jacA.dMetric_dA[i];
      double norm = f;
      double iden = 1.0 / den;
      const double s1 = (3 * norm * iden);
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) wrt_A[a][b] = T[a][b] * s1;
      double norm_cube = norm * norm * norm;
      transpose_adj(T, TAT);
      const double s2 = (norm_cube * iden / d);
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) wrt_A[a][b] -= TAT[a][b] * s2;
      if (use_ref) {
        for (int a = 0; a < 3; a++)
          for (int b = 0; b < 3; b++) WIT[a][b] = WI[b][a];
        M33 tmp;
        matrix_product_densematrix_op(wrt_A, WIT, tmp);  // wrt_A = wrt_A * WIT  (DenseMatrix operator*)
        memcpy(wrt_A, tmp, sizeof(M33));
      }
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) wrt_A[a][b] = wrt_A[a][b] * detWiC;
    }
    val_shape += shape_metric;
  }
  jacA.grad_metric_util(topo);
  for (int i = 0; i < nn; i++)
    for (int j = 0; j < 3; j++) grad[i][j] = 0.0;
  for (int k = 0; k < nn; k++)
    for (int i = 0; i < nn; i++)
      for (int j = 0; j < 3; j++) grad[i][j] += jacA.grad[k][i][j];
  return val_shape;
}

// ---------------------------------------------------------------------------------
// mesh container
// ---------------------------------------------------------------------------------
struct BC {
  int beg[3], end[3];
};
struct Zone {
  int donor;
  int transform[3], owner_beg[3], owner_end[3], donor_beg[3], donor_end[3];
};
struct Block {
  unsigned n[3];
  size_t off;
  std::vector<BC> bcs;
  std::vector<Zone> zones;
  size_t nnodes() const { return (size_t)n[0] * n[1] * n[2]; }
  size_t ncells() const { return (size_t)(n[0] - 1) * (n[1] - 1) * (n[2] - 1); }
  size_t node(unsigned i, unsigned j, unsigned k) const { return off + i + (size_t)n[0] * (j + (size_t)n[1] * k); }
};
struct Field {
  int ncomp = 0;
  std::vector<double> d;  // SoA: comp c of node n at d[c*N + n]
};

static const char* FIELD_NAMES3[] = {"coordinates", "coordinates_N", "coordinates_NM1", "coordinates_lagged",
                                     "cg_g",        "cg_r",          "cg_d",            "cg_s"};
static const char* FIELD_NAMES1[] = {"cg_edge_length", "num_adj_elems"};

}  // namespace

struct orc_ctx {
  bool structured = false, committed = false;
  std::vector<Block> blocks;
  int topo = 0;
  size_t N = 0, nelems = 0;
  std::vector<int32_t> conn;
  std::vector<uint8_t> elem_owned, node_owned, fixed;
  std::vector<int64_t> csr_ptr, csr_ent;
  std::map<std::string, Field> fields;
  bool use_ref_mesh = true, real_is_double = false, adj_fixed = false;
  int omp_threads = 1;
  // surface smoothing hooks: node classification as MeshGeometry::classify_node returns it
  // (0 vertex, 1 curve, 2 surface, 3 volume) + the evaluator (analytic surface) of each surface node
  bool smooth_surfaces = false;
  struct Surface {
    int kind;       // 0 plane (p, n), 1 sphere (c, R), 2 cylinder (a, u, R)
    double p[3], u[3], R;
  };
  std::vector<Surface> surfaces;
  std::vector<int8_t> node_dof;
  std::vector<int32_t> node_surf;
  std::vector<double> elem_metric;
  std::vector<int32_t> elem_flag;
  std::string err;
  double* f(const char* name) { return fields.at(name).d.data(); }
};

namespace {

// Analytic stand-ins for MeshGeometry::normal_at / snap_points_to_geometry (mesh/geometry/kernel/MeshGeometry.hpp:155-265;
// the reference evaluates OpenNURBS / PGEOM CAD, third party and absent here).  Plain double arithmetic, left to right.
inline void surf_normal(const orc_ctx::Surface& s, const double x[3], double n[3], double* t_out = nullptr) {
  if (s.kind == 0) {
    for (int i = 0; i < 3; i++) n[i] = s.u[i];
    return;
  }
  double d[3] = {x[0] - s.p[0], x[1] - s.p[1], x[2] - s.p[2]};
  if (s.kind == 2) {
    const double t = d[0] * s.u[0] + d[1] * s.u[1] + d[2] * s.u[2];
    for (int i = 0; i < 3; i++) d[i] = d[i] - t * s.u[i];
    if (t_out) *t_out = t;
  }
  const double r = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  for (int i = 0; i < 3; i++) n[i] = r > 0.0 ? d[i] / r : 0.0;
}
inline void surf_snap(const orc_ctx::Surface& s, double x[3]) {
  double n[3], t = 0.0;
  surf_normal(s, x, n, &t);
  if (s.kind == 0) {
    const double h = (x[0] - s.p[0]) * n[0] + (x[1] - s.p[1]) * n[1] + (x[2] - s.p[2]) * n[2];
    for (int i = 0; i < 3; i++) x[i] = x[i] - h * n[i];
  } else if (s.kind == 1) {
    for (int i = 0; i < 3; i++) x[i] = s.p[i] + s.R * n[i];
  } else {
    for (int i = 0; i < 3; i++) x[i] = (s.p[i] + t * s.u[i]) + s.R * n[i];
  }
}
// get_fixed_flag with a geometry and no boundary selector (MeshSmoother.cpp:293-340)
inline void apply_classification(orc_ctx& c) {
  if (c.node_dof.empty()) return;
  c.fixed.assign(c.node_dof.size(), 0);
  for (size_t n = 0; n < c.node_dof.size(); n++) {
    const int dof = c.node_dof[n];
    c.fixed[n] = (dof == 0 || dof == 1 || (dof == 2 && !c.smooth_surfaces)) ? 1 : 0;
  }
}

struct OrcError : std::runtime_error {
  int code;
  OrcError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

// device_safe_transform_block_indices, MeshType.hpp:86-131
inline void transform_block_indices(const int idx[3], const int tr[3], const int lbeg[3], const int dbeg[3], int donor[3]) {
  int t[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) t[3 * i + j] = (tr[j] < 0 ? -1 : 1) * (int)(std::abs(tr[j]) == i + 1);
  int diff[3] = {idx[0] - lbeg[0], idx[1] - lbeg[1], idx[2] - lbeg[2]};
  for (int r = 0; r < 3; r++) donor[r] = t[3 * r] * diff[0] + t[3 * r + 1] * diff[1] + t[3 * r + 2] * diff[2] + dbeg[r];
}

// iterate the local nodes of a zone connectivity with their donor nodes
template <class F>
void for_zone_nodes(const orc_ctx& c, int ib, const Zone& z, F fn) {
  int lb[3], le[3];
  unsigned ls[3];
  for (int d = 0; d < 3; d++) {
    lb[d] = std::min(z.owner_beg[d], z.owner_end[d]);
    le[d] = std::max(z.owner_beg[d], z.owner_end[d]);
    ls[d] = 1 + le[d] - lb[d];
  }
  const Block& B = c.blocks[ib];
  const Block& D = c.blocks[z.donor];
  size_t n = (size_t)ls[0] * ls[1] * ls[2];
  for (size_t iNode = 0; iNode < n; iNode++) {
    // gradient_functors.hpp:194-208 (note: uses owner_beg, not min(beg,end), as origin)
    unsigned ix[3];
    ix[0] = (z.owner_beg[0] - 1) + (iNode % ls[0]);
    ix[1] = (z.owner_beg[1] - 1) + (iNode / ls[0]) % ls[1];
    ix[2] = (z.owner_beg[2] - 1) + ((iNode / ls[0]) / ls[1]) % ls[2];
    int li[3] = {(int)ix[0] + 1, (int)ix[1] + 1, (int)ix[2] + 1};
    int di[3];
    transform_block_indices(li, z.transform, z.owner_beg, z.donor_beg, di);
    fn(B.node(ix[0], ix[1], ix[2]), D.node(di[0] - 1, di[1] - 1, di[2] - 1));
  }
}

// sum pass (iblock<donor: loc += don) then propagate pass (iblock>donor: loc = don)
// gradient_functors.hpp:141-220, get_edge_len_avg.hpp:146-249
void sum_and_propagate(orc_ctx& c, double* fld, int ncomp) {
  for (int pass = 0; pass < 2; pass++)
    for (int ib = 0; ib < (int)c.blocks.size(); ib++)
      for (const Zone& z : c.blocks[ib].zones) {
        bool doSum = pass == 0;
        if (doSum && ib < z.donor)
          for_zone_nodes(c, ib, z, [&](size_t l, size_t d) {
            for (int k = 0; k < ncomp; k++) fld[k * c.N + l] += fld[k * c.N + d];
          });
        else if (!doSum && ib > z.donor)
          for_zone_nodes(c, ib, z, [&](size_t l, size_t d) {
            for (int k = 0; k < ncomp; k++) fld[k * c.N + l] = fld[k * c.N + d];
          });
      }
}

// SB_nodal_field_dot ownership filter: BlockStructuredGrid.cpp:448-458, BlockStructuredGrid.hpp:156-300
bool sgrid_node_counted(const Block& B, int blockID, unsigned i, unsigned j, unsigned k) {
  bool on_end = (i == 0 || i == B.n[0] - 1 || j == 0 || j == B.n[1] - 1 || k == 0 || k == B.n[2] - 1);
  if (!on_end) return true;
  bool inAny = false;
  unsigned min_block = blockID;
  for (const Zone& z : B.zones) {
    unsigned lb[3], le[3];
    for (int d = 0; d < 3; d++) {
      lb[d] = std::min(z.owner_beg[d], z.owner_end[d]) - 1;
      le[d] = std::max(z.owner_beg[d], z.owner_end[d]) - 1;
    }
    bool in = lb[0] <= i && i <= le[0] && lb[1] <= j && j <= le[1] && lb[2] <= k && k <= le[2];
    if (in) {
      inAny = true;
      if (min_block > (unsigned)z.donor) min_block = z.donor;
    }
  }
  if (!inAny) return true;                 // boundary and not on an interface
  return min_block == (unsigned)blockID;   // interface: lowest block id counts it
}

inline void gather_cell(const orc_ctx& c, const Block& B, const double* X, unsigned ci, unsigned cj, unsigned ck,
                        double v[8][3], size_t nid[8]) {  // total_element_metric.cpp:286-302
  int cnt = 0;
  for (unsigned dk = 0; dk < 2; dk++)
    for (unsigned dj = 0; dj < 2; dj++)
      for (unsigned di = 0; di < 2; di++) {
        size_t n = B.node(ci + di, cj + dj, ck + dk);
        nid[cnt] = n;
        for (int ic = 0; ic < 3; ic++) v[cnt][ic] = X[ic * c.N + n];
        ++cnt;
      }
}

inline void gather_elem(const orc_ctx& c, const double* X, size_t e, double v[8][3]) {
  for (int a = 0; a < c.topo; a++) {
    size_t n = c.conn[e * c.topo + a];
    for (int ic = 0; ic < 3; ic++) v[a][ic] = X[ic * c.N + n];
  }
}

void copy_field(orc_ctx& c, const char* dst, const char* src) {
  Field& d = c.fields.at(dst);
  Field& s = c.fields.at(src);
  if (d.ncomp != s.ncomp) throw OrcError(1, "copy_field: component mismatch");
  d.d = s.d;
}

// ---------------------------------------------------------------------------------
// smoother, templated on Double (MeshType.hpp:41-46)
// ---------------------------------------------------------------------------------
template <class Real>
struct Smoother {
  orc_ctx& c;
  explicit Smoother(orc_ctx& cc) : c(cc) {}

  // ---- BLAS-1 ----
  Real dot(const char* xn, const char* yn) {
    Field& fx = c.fields.at(xn);
    Field& fy = c.fields.at(yn);
    const double* x = fx.d.data();
    const double* y = fy.d.data();
    const int nc = fx.ncomp;
    Real sum = 0.0;
    if (c.structured) {  // BlockStructuredGrid.cpp:519-539
      for (int ib = 0; ib < (int)c.blocks.size(); ib++) {
        const Block& B = c.blocks[ib];
        Real tot = 0;
        for (unsigned k = 0; k < B.n[2]; k++)
          for (unsigned j = 0; j < B.n[1]; j++)
            for (unsigned i = 0; i < B.n[0]; i++)
              if (sgrid_node_counted(B, ib, i, j, k)) {
                size_t n = B.node(i, j, k);
                for (int d = 0; d < nc; d++) tot += x[d * c.N + n] * y[d * c.N + n];
              }
        Real m_sum = 0.0;
        m_sum += tot;
        sum += m_sum;
      }
    } else {  // PerceptMesh.cpp:4690-4726 (owned nodes only)
      for (size_t n = 0; n < c.N; n++)
        if (c.node_owned[n])
          for (int d = 0; d < nc; d++) sum += x[d * c.N + n] * y[d * c.N + n];
    }
    return sum;
  }

  size_t get_number_nodes() {  // PerceptMesh.cpp:730-756
    size_t cnt = 0;
    if (c.structured) {
      for (int ib = 0; ib < (int)c.blocks.size(); ib++) {
        const Block& B = c.blocks[ib];
        for (unsigned k = 0; k < B.n[2]; k++)
          for (unsigned j = 0; j < B.n[1]; j++)
            for (unsigned i = 0; i < B.n[0]; i++)
              if (sgrid_node_counted(B, ib, i, j, k)) cnt++;
      }
    } else
      for (size_t n = 0; n < c.N; n++) cnt += c.node_owned[n] ? 1 : 0;
    return cnt;
  }

  // ---- update_coordinates: x += alpha*s on unfixed nodes (update_coordinates.cpp:94-118, 241-268) ----
  void coord_update(Real alpha, bool calc_deltas, Real& dmax, Real& dmax_relative) {
    copy_field(c, "coordinates_lagged", "coordinates");
    double* x = c.f("coordinates");
    const double* s = c.f("cg_s");
    const double* el = c.f("cg_edge_length");
    if (calc_deltas) {
      dmax = 0;
      dmax_relative = 0;
    }
    for (size_t n = 0; n < c.N; n++) {
      if (c.fixed[n]) continue;
      if (!c.structured && !c.node_owned[n] && false) continue;
      Real maxDelta = 0.0;
      for (int d = 0; d < 3; d++) {
        Real dt = alpha * s[d * c.N + n];
        x[d * c.N + n] += dt;
        Real a = dt < 0 ? -dt : dt;
        if (calc_deltas && maxDelta < a) maxDelta = a;
      }
      if (calc_deltas) {
        Real e = el[n];
        dmax = std::max(maxDelta, dmax);
        dmax_relative = std::max(maxDelta / e, dmax_relative);
      }
    }
  }

  // ---- element metric over the mesh at the CURRENT "coordinates" ----
  Real sum_metric(int stage, size_t& n_invalid, bool keep) {
    Real mtot = 0.0;
    n_invalid = 0;
    const double* X = c.f("coordinates");
    const double* W = c.f("coordinates_NM1");
    if (keep) {
      c.elem_metric.assign(c.nelems, 0.0);
      c.elem_flag.assign(c.nelems, 0);
    }
    size_t eoff = 0;
    if (c.structured) {
      HexMeshSmootherMetric hm;
      hm.m_untangling = stage == 0;
      hm.m_use_ref_mesh = c.use_ref_mesh;
      for (const Block& B : c.blocks) {  // Def.hpp:353-360: per-block reduce, then added
        Real bt = 0;
        size_t bi = 0;
        for (unsigned k = 0; k + 1 < B.n[2]; k++)
          for (unsigned j = 0; j + 1 < B.n[1]; j++)
            for (unsigned i = 0; i + 1 < B.n[0]; i++) {
              double vc[8][3], vo[8][3];
              size_t nid[8];
              gather_cell(c, B, X, i, j, k, vc, nid);
              gather_cell(c, B, W, i, j, k, vo, nid);
              bool valid = false;
              Real mm = hm.metric(vc, vo, valid);
              bt += mm;
              bi += !valid;
              if (keep) {
                c.elem_metric[eoff] = (double)mm;
                c.elem_flag[eoff] = !valid;
              }
              eoff++;
            }
        mtot += bt;
        n_invalid += bi;
      }
    } else {
      for (size_t e = 0; e < c.nelems; e++) {
        if (!c.elem_owned[e]) continue;
        double vc[8][3], vo[8][3];
        gather_elem(c, X, e, vc);
        gather_elem(c, W, e, vo);
        bool valid = false;
        Real mm = stage == 0 ? stk_untangle_metric(c.topo, vc, vo, valid)
                             : stk_shapeb1_metric(c.topo, c.use_ref_mesh, vc, vo, valid);
        mtot += mm;
        n_invalid += !valid;
        if (keep) {
          c.elem_metric[e] = (double)mm;
          c.elem_flag[e] = !valid;
        }
      }
    }
    return mtot;
  }

  // total_metric, Def.hpp:330-388
  Real total_metric(int stage, Real alpha, bool& valid, size_t* num_invalid, bool keep = false) {
    Real dm, dr;
    coord_update(alpha, false, dm, dr);
    size_t n_invalid = 0;
    Real mtot = sum_metric(stage, n_invalid, keep);
    valid = (n_invalid == 0);
    copy_field(c, "coordinates", "coordinates_lagged");
    if (num_invalid) *num_invalid = n_invalid;
    return mtot;
  }

  size_t count_invalid() {  // MeshSmoother.cpp:192-234 (untangle metric's valid flag)
    size_t n_invalid = 0;
    sum_metric(0, n_invalid, false);
    return n_invalid;
  }

  // get_gradient, Def.hpp:1481-1520
  void get_gradient(int stage) {
    double* g = c.f("cg_g");
    std::fill(c.fields.at("cg_g").d.begin(), c.fields.at("cg_g").d.end(), 0.0);
    const double* X = c.f("coordinates");
    const double* W = c.f("coordinates_NM1");
    if (c.structured) {  // gradient_functors.hpp:349-486
      HexMeshSmootherMetric hm;
      hm.m_untangling = stage == 0;
      hm.m_use_ref_mesh = true;  // get_gradient() never forwards m_use_ref_mesh to m_sgrid_gradient (Def.hpp:1485-1495)
      for (const Block& B : c.blocks) {
        for (unsigned k = 0; k + 1 < B.n[2]; k++)
          for (unsigned j = 0; j + 1 < B.n[1]; j++)
            for (unsigned i = 0; i + 1 < B.n[0]; i++) {
              double vc[8][3], vo[8][3], grad[8][3];
              size_t nid[8];
              gather_cell(c, B, X, i, j, k, vc, nid);
              gather_cell(c, B, W, i, j, k, vo, nid);
              bool gmvalid = true;
              hm.grad_metric(vc, vo, gmvalid, grad);
              if (gmvalid || hm.m_untangling)
                for (int in = 0; in < 8; in++)
                  for (int d = 0; d < 3; d++) g[d * c.N + nid[in]] += grad[in][d];
            }
        for (size_t n = B.off; n < B.off + B.nnodes(); n++)  // fixup, gradient_functors.hpp:223-243
          if (c.fixed[n])
            for (int d = 0; d < 3; d++) g[d * c.N + n] = 0.0;
      }
      sum_and_propagate(c, g, 3);
    } else {  // Def.hpp:1195-1250
      for (size_t e = 0; e < c.nelems; e++) {
        double vc[8][3], vo[8][3], grad[8][3];
        gather_elem(c, X, e, vc);
        gather_elem(c, W, e, vo);
        bool gmvalid = true;
        if (stage == 0)
          stk_untangle_grad_metric(c.topo, vc, vo, gmvalid, grad);
        else
          stk_shapeb1_grad_metric(c.topo, c.use_ref_mesh, vc, vo, gmvalid, grad);
        if (gmvalid || stage == 0)
          for (int in = 0; in < c.topo; in++) {
            size_t n = c.conn[e * c.topo + in];
            if (c.fixed[n]) continue;
            for (int d = 0; d < 3; d++) {
              if (c.node_owned[n])
                g[d * c.N + n] += grad[in][d];
              else
                g[d * c.N + n] = 0.0;
            }
          }
      }
      copy_field(c, "cg_r", "cg_g");
      if (c.smooth_surfaces && !c.node_dof.empty())  // GenericAlgorithm_get_gradient_2, Def.hpp:1401-1447
        for (size_t n = 0; n < c.N; n++) {
          if (c.fixed[n] || c.node_dof[n] != 2) continue;
          double x[3] = {X[n], X[c.N + n], X[2 * c.N + n]}, nrm[3];
          surf_normal(c.surfaces.at(c.node_surf[n]), x, nrm);
          double dot = 0.0;  // project_delta_to_tangent_plane, MeshSmoother.cpp:437-460
          for (int i = 0; i < 3; i++) dot += g[i * c.N + n] * nrm[i];
          for (int i = 0; i < 3; i++) g[i * c.N + n] -= dot * nrm[i];
        }
    }
  }

  // snap_nodes (Def.hpp:236-316): non-fixed MS_SURFACE nodes go back onto their surface; returns max displacement
  double snap_nodes() {
    double dmax = 0.0;
    if (c.node_dof.empty()) return dmax;
    double* X = c.f("coordinates");
    for (size_t n = 0; n < c.N; n++) {
      if (c.fixed[n] || c.node_dof[n] != 2) continue;
      double x[3] = {X[n], X[c.N + n], X[2 * c.N + n]};
      const double o[3] = {x[0], x[1], x[2]};
      surf_snap(c.surfaces.at(c.node_surf[n]), x);
      double dm = 0.0;
      for (int i = 0; i < 3; i++) {
        dm += (o[i] - x[i]) * (o[i] - x[i]);
        X[i * c.N + n] = x[i];
      }
      dmax = std::max(dmax, std::sqrt(dm));
    }
    return dmax;
  }

  // ---- edge lengths ----
  // sgrid: get_edge_len_avg.hpp:64-144, 302-383 ; STK: Def.hpp:39-65 + PerceptMesh.cpp:3173-3220
  void get_edge_lengths() {
    const double* W = c.f("coordinates_NM1");
    double* el = c.f("cg_edge_length");
    double* na = c.f("num_adj_elems");
    if (c.structured) {
      static const int edges[12][2] = {{0, 1}, {1, 3}, {3, 2}, {2, 0}, {4, 5}, {5, 7},
                                       {7, 6}, {6, 4}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};
      for (const Block& B : c.blocks)
        for (unsigned k = 0; k < B.n[2]; k++)
          for (unsigned j = 0; j < B.n[1]; j++)
            for (unsigned i = 0; i < B.n[0]; i++) {
              Real nm = 0.0;
              unsigned num_adj = 0;
              for (int i0 = (int)i - 1; i0 <= (int)i; ++i0) {
                if (i0 < 0 || i0 > (int)B.n[0] - 2) continue;
                for (int i1 = (int)j - 1; i1 <= (int)j; ++i1) {
                  if (i1 < 0 || i1 > (int)B.n[1] - 2) continue;
                  for (int i2 = (int)k - 1; i2 <= (int)k; ++i2) {
                    if (i2 < 0 || i2 > (int)B.n[2] - 2) continue;
                    num_adj++;
                    size_t en[8];
                    int cnt = 0;
                    for (int a2 = i2; a2 <= i2 + 1; a2++)
                      for (int a1 = i1; a1 <= i1 + 1; a1++)
                        for (int a0 = i0; a0 <= i0 + 1; a0++) en[cnt++] = B.node(a0, a1, a2);
                    double edge_length_ave = 0.0;
                    for (int ie = 0; ie < 12; ie++) {
                      Real n0[3], n1[3];  // `Double node_coord_data_*` get_edge_len_avg.hpp:80-88
                      for (int d = 0; d < 3; d++) {
                        n0[d] = W[d * c.N + en[edges[ie][0]]];
                        n1[d] = W[d * c.N + en[edges[ie][1]]];
                      }
                      double edge_length = 0.0;
                      for (int d = 0; d < 3; d++) edge_length += (n0[d] - n1[d]) * (n0[d] - n1[d]);
                      edge_length = std::sqrt(edge_length);
                      edge_length_ave += edge_length / ((double)12);
                    }
                    nm += edge_length_ave;
                  }
                }
              }
              size_t n = B.node(i, j, k);
              if (!c.adj_fixed) na[n] = (double)num_adj;
              el[n] = (double)nm;
            }
      // interface sum + propagate (num_adj only the first time, get_edge_len_avg.hpp:333-341)
      sum_and_propagate(c, el, 1);
      if (!c.adj_fixed) sum_and_propagate(c, na, 1);
      c.adj_fixed = true;
      for (size_t n = 0; n < c.N; n++) el[n] /= na[n];
    } else {
      static const int hex_edges[12][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}, {4, 5}, {5, 6},
                                           {6, 7}, {7, 4}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};  // Shards Hexahedron<8>
      static const int tet_edges[6][2] = {{0, 1}, {1, 2}, {2, 0}, {0, 3}, {1, 3}, {2, 3}};  // Shards Tetrahedron<4>
      const int ne = c.topo == 8 ? 12 : 6;
      for (size_t n = 0; n < c.N; n++) {
        Real nm = 0.0, nele = 0.0;
        for (int64_t p = c.csr_ptr[n]; p < c.csr_ptr[n + 1]; p++) {
          size_t e = c.csr_ent[p] >> 3;
          double edge_length_ave = 0.0;
          for (int ie = 0; ie < ne; ie++) {
            int a = c.topo == 8 ? hex_edges[ie][0] : tet_edges[ie][0];
            int b = c.topo == 8 ? hex_edges[ie][1] : tet_edges[ie][1];
            size_t na_ = c.conn[e * c.topo + a], nb_ = c.conn[e * c.topo + b];
            double edge_length = 0.0;
            for (int d = 0; d < 3; d++)
              edge_length += (W[d * c.N + na_] - W[d * c.N + nb_]) * (W[d * c.N + na_] - W[d * c.N + nb_]);
            edge_length = std::sqrt(edge_length);
            edge_length_ave += edge_length / ((double)ne);
          }
          nm += edge_length_ave;
          nele += 1.0;
        }
        nm /= nele;
        el[n] = (double)nm;
        na[n] = (double)nele;
      }
    }
  }

  // get_alpha_0: sgrid get_alpha_0_refmesh.hpp:111-193 ; STK Def.hpp:843-901
  Real get_alpha_0() {
    const double* s = c.f("cg_s");
    const double* el = c.f("cg_edge_length");
    Real alpha = std::numeric_limits<Real>::max();
    if (c.structured) {
      for (size_t n = 0; n < c.N; n++) {
        Real edge_length_ave = el[n];
        Real sn = 0.0;
        for (int d = 0; d < 3; d++) sn += s[d * c.N + n] * s[d * c.N + n];
        sn = std::sqrt(sn);
        Real cand = sn > 0.0 ? edge_length_ave / sn : std::numeric_limits<Real>::max();
        if (alpha > cand) alpha = cand;
      }
      if (!(alpha > 0.0)) throw OrcError(4, "bad alpha");
      if (alpha == std::numeric_limits<Real>::max()) alpha = 1.0;
    } else {
      bool alpha_set = false;
      for (size_t n = 0; n < c.N; n++) {
        if (c.fixed[n]) continue;
        Real edge_length_ave = el[n];
        Real sn = 0.0;
        for (int d = 0; d < 3; d++) sn += s[d * c.N + n] * s[d * c.N + n];
        sn = std::sqrt(sn);
        if (sn > 0.0) {
          Real alpha_new = edge_length_ave / sn;
          if (!alpha_set) {
            alpha_set = true;
            alpha = alpha_new;
          } else if (alpha_new < alpha)
            alpha = alpha_new;
        }
      }
      if (!alpha_set) alpha = 1.0;
      if (!(alpha > 0.0)) throw OrcError(4, "bad alpha");
    }
    return alpha;
  }

  // ---- driver state (ReferenceMeshSmootherBase.hpp:71-101) ----
  Real m_dmax = 0, m_dmax_relative = 0, m_dnew = 0, m_dold = 0, m_d0 = 1, m_dmid = 0, m_alpha = 0, m_alpha_0 = 0,
       m_grad_norm_scaled = 0, m_total_metric = 0;
  int m_stage = 0, m_iter = 0;
  size_t m_num_invalid = 0, m_num_nodes = 0;
  Real m_global_metric = std::numeric_limits<double>::max();
  bool m_untangled = false;
  double gradNorm = 1e-8;
  uint64_t n_metric_evals = 0, n_gradient_evals = 0, n_halvings = 0;

  Real tm(Real alpha, bool& valid, size_t* ninv = 0) {
    n_metric_evals++;
    return total_metric(m_stage, alpha, valid, ninv);
  }

  void axpby(double alpha, const char* xn, double beta, const char* yn) {  // PerceptMesh.cpp:4611-4650, BlockStructuredGrid.cpp:383-405
    Field& fx = c.fields.at(xn);
    Field& fy = c.fields.at(yn);
    for (size_t i = 0; i < fy.d.size(); i++) fy.d[i] = alpha * fx.d[i] + beta * fy.d[i];
  }

  // line_search, Def.hpp:466-591
  Real line_search(bool& restarted, const Real mfac_mult = 1.0) {
    const Real alpha_fac = 10.0, tau = 0.5, c0 = 1.e-1, min_alpha_factor = 1.e-12;
    m_alpha_0 = get_alpha_0();
    Real alpha = alpha_fac * m_alpha_0;
    bool total_valid = false;
    size_t n_invalid = 0;
    const Real metric_0 = tm(0.0, total_valid, &n_invalid);
    Real metric = 0.0;
    Real sDotGrad = dot("cg_s", "cg_g");
    if (sDotGrad >= 0.0) {
      copy_field(c, "cg_s", "cg_r");
      m_alpha_0 = get_alpha_0();
      alpha = alpha_fac * m_alpha_0;
      sDotGrad = dot("cg_s", "cg_g");
      restarted = true;
    }
    if (!(sDotGrad < 0.0)) throw OrcError(4, "bad sDotGrad");
    const Real armijo_offset_factor = c0 * sDotGrad;
    bool converged = false;
    total_valid = false;
    int liter = 0, niter = 1000;
    while (!converged && liter < niter) {
      metric = tm(alpha, total_valid, &n_invalid);
      const Real mfac = alpha * armijo_offset_factor * mfac_mult;
      converged = (metric < metric_0 + mfac);
      if (m_untangled) converged = converged && total_valid;
      if (!converged) {
        alpha *= tau;
        n_halvings++;
      }
      if (alpha < min_alpha_factor * m_alpha_0) break;
      ++liter;
    }
    if (!converged) {
      restarted = true;
      copy_field(c, "cg_s", "cg_r");
      m_alpha_0 = get_alpha_0();
      alpha = alpha_fac * m_alpha_0;
      sDotGrad = dot("cg_s", "cg_g");
      if (!(sDotGrad < 0.0)) throw OrcError(4, "bad sDotGrad 2nd time");
    }
    liter = 0;
    while (!converged && liter < niter) {
      metric = tm(alpha, total_valid);
      const Real mfac = alpha * armijo_offset_factor * mfac_mult;
      converged = (metric < metric_0 + mfac);
      if (m_untangled) {
        converged = converged && total_valid;
        if (total_valid && m_dmax_relative < gradNorm) converged = true;
      }
      if (!converged) {
        alpha *= tau;
        n_halvings++;
      }
      if (alpha < min_alpha_factor * m_alpha_0) break;
      ++liter;
    }
    if (!converged) throw OrcError(4, "can't reduce metric");
    Real a1 = alpha / 2., a2 = alpha;
    Real f0 = metric_0, f1 = tm(a1, total_valid), f2 = tm(a2, total_valid);
    Real den = 2. * (a2 * (-f0 + f1) + a1 * (f0 - f2));
    Real num = a2 * a2 * (f1 - f0) + a1 * a1 * (f0 - f2);
    if (std::fabs(den) > 1.e-10) {
      Real alpha_quadratic = num / den;
      if (alpha_quadratic > 1.e-10 && alpha_quadratic < 2 * alpha) {
        Real fm = tm(alpha_quadratic, total_valid);
        if (fm < f2 && (m_stage == 0 || total_valid)) alpha = alpha_quadratic;
      }
    }
    return alpha;
  }

  bool check_convergence() {  // Def.hpp:989-1008
    Real grad_check = gradNorm;
    bool retval = false;
    if (m_stage == 0 && (m_dnew == 0.0 || m_total_metric == 0.0))
      retval = true;
    else if (m_stage == 0 && m_num_invalid == 0 && (m_dmax_relative < grad_check))
      retval = true;
    else if (m_num_invalid == 0 &&
             (m_iter > 10 && (m_dmax_relative < grad_check &&
                              (m_dnew < grad_check * grad_check * m_d0 || m_grad_norm_scaled < grad_check))))
      retval = true;
    return retval;
  }

  double run_one_iteration() {  // Def.hpp:1010-1093
    bool total_valid = false;
    if (m_iter == 0) {
      get_edge_lengths();
      for (const char* nm : {"cg_r", "cg_d", "cg_s"}) std::fill(c.fields.at(nm).d.begin(), c.fields.at(nm).d.end(), 0.0);
    }
    get_gradient(m_stage);
    n_gradient_evals++;
    axpby(-1.0, "cg_g", 0.0, "cg_r");
    m_dold = dot("cg_d", "cg_d");
    m_dmid = dot("cg_r", "cg_d");
    m_dnew = dot("cg_r", "cg_r");
    if (m_iter == 0) m_d0 = m_dnew;
    Real metric_check = tm(0.0, total_valid);
    m_total_metric = metric_check;
    if (check_convergence() || metric_check == 0.0) return tm(0.0, total_valid);
    Real cg_beta = 0.0;
    if (m_dold == 0.0)
      cg_beta = 0.0;
    else if (m_iter > 0)
      cg_beta = (m_dnew - m_dmid) / m_dold;
    size_t N = m_num_nodes;
    if (m_iter % N == 0 || cg_beta <= 0.0)
      copy_field(c, "cg_s", "cg_r");
    else
      axpby(1.0, "cg_r", (double)cg_beta, "cg_s");
    copy_field(c, "cg_d", "cg_r");
    bool restarted = false;
    Real alpha = line_search(restarted);
    Real snorm = dot("cg_s", "cg_s");
    m_grad_norm_scaled = m_alpha_0 * std::sqrt(snorm) / Real(m_num_nodes);
    m_alpha = alpha;
    coord_update(alpha, true, m_dmax, m_dmax_relative);  // update_node_positions, Def.hpp:395-410
    if (c.smooth_surfaces) {  // check if re-snapped geometry is acceptable, Def.hpp:1078-1089
      snap_nodes();
      if (m_stage != 0) {
        bool total_valid_0 = true;
        tm(0.0, total_valid_0);
        if (!total_valid_0) throw OrcError(3, "bad mesh after snap_node_positions...");
      }
    }
    Real t = tm(0.0, total_valid);
    return t;
  }

  // run_algorithm, ReferenceMeshSmootherBase.cpp:240-334, log format :57-81, :216-238
  void run_algorithm(int innerIter, double grad_norm, const char* log_path) {
    gradNorm = grad_norm;
    std::ofstream myFile;
    const unsigned width = 14;
    if (log_path) {
      myFile.open(log_path, std::ios::out);
      myFile.precision(6);
      myFile.setf(std::ios_base::left, std::ios_base::adjustfield);
      myFile.setf(std::ios_base::scientific);
      myFile << std::setw(8) << "%iter" << std::setw(width) << "global_metric" << std::setw(width) << "invalid_elems"
             << std::setw(width) << "dmax" << std::setw(width) << "dmax_rel" << std::setw(width) << "grad/g0"
             << std::setw(width) << "gradScaled" << std::setw(width) << "alpha" << std::setw(width) << "alpha_0"
             << std::setw(6) << "stage" << std::setw(width) << "untangled_elems" << std::setw(width) << "noNodes"
             << std::endl;
    }
    m_num_nodes = get_number_nodes();
    copy_field(c, "coordinates_lagged", "coordinates_NM1");
    {
      // coordinates = 1.0*coordinates_N + 0.0*coordinates_NM1 + 0.0*coordinates (axpbypgz)
      Field& z = c.fields.at("coordinates");
      Field& x = c.fields.at("coordinates_N");
      Field& y = c.fields.at("coordinates_NM1");
      for (size_t i = 0; i < z.d.size(); i++) z.d[i] = 1.0 * x.d[i] + 0.0 * y.d[i] + 0.0 * z.d[i];
    }
    m_num_invalid = count_invalid();
    m_untangled = (m_num_invalid == 0);
    for (int stage = 0; stage < 2; stage++) {
      m_stage = stage;
      if (stage == 1) {
        size_t num_invalid_1 = count_invalid();
        if (num_invalid_1 != 0) throw OrcError(4, "Invalid elements exist for start of stage 2, quiting");
      }
      for (int iter = 0; iter < innerIter; ++iter) {
        m_iter = iter;
        m_num_invalid = count_invalid();
        int num_invalid_0 = (int)m_num_invalid;
        if (!m_untangled && m_num_invalid == 0) m_untangled = true;
        m_global_metric = run_one_iteration();
        bool conv = check_convergence();
        if (log_path) {
          myFile << std::setw(8) << m_iter << std::setw(width) << m_global_metric << std::setw(width) << num_invalid_0
                 << std::setw(width) << m_dmax << std::setw(width) << m_dmax_relative << std::setw(width)
                 << std::sqrt(m_dnew / m_d0) << std::setw(width) << m_grad_norm_scaled << std::setw(width) << m_alpha
                 << std::setw(width) << m_alpha_0 << std::setw(6) << m_stage << std::setw(width) << m_untangled
                 << std::setw(width) << m_num_nodes << std::endl;
        }
        if (!m_untangled && m_num_invalid == 0) m_untangled = true;
        if (conv && m_untangled) break;
      }
    }
    copy_field(c, "coordinates_lagged", "coordinates");
  }
};

}  // namespace

// ---------------------------------------------------------------------------------
// C ABI (orc_ twin of include/percept_smoother_c.h)
// ---------------------------------------------------------------------------------
struct orc_state {
  double dmax, dmax_relative, dnew, dold, d0, dmid, alpha, alpha_0, grad_norm_scaled, total_metric, global_metric;
  int stage, iter, untangled, converged;
  uint64_t num_invalid, num_nodes, n_metric_evals, n_gradient_evals, n_line_search_halvings, restarted;
};

#define ORC_TRY try {
#define ORC_CATCH                       \
  }                                     \
  catch (const OrcError& e) {           \
    c->err = e.what();                  \
    return e.code;                      \
  }                                     \
  catch (const std::exception& e) {     \
    c->err = e.what();                  \
    return 1;                           \
  }                                     \
  return 0;

template <class Real>
static void state_to(const Smoother<Real>& s, orc_state* st) {
  st->dmax = (double)s.m_dmax;
  st->dmax_relative = (double)s.m_dmax_relative;
  st->dnew = (double)s.m_dnew;
  st->dold = (double)s.m_dold;
  st->d0 = (double)s.m_d0;
  st->dmid = (double)s.m_dmid;
  st->alpha = (double)s.m_alpha;
  st->alpha_0 = (double)s.m_alpha_0;
  st->grad_norm_scaled = (double)s.m_grad_norm_scaled;
  st->total_metric = (double)s.m_total_metric;
  st->global_metric = (double)s.m_global_metric;
  st->stage = s.m_stage;
  st->iter = s.m_iter;
  st->untangled = s.m_untangled;
  st->num_invalid = s.m_num_invalid;
  st->num_nodes = s.m_num_nodes;
  st->n_metric_evals = s.n_metric_evals;
  st->n_gradient_evals = s.n_gradient_evals;
  st->n_line_search_halvings = s.n_halvings;
}
template <class Real>
static void state_from(Smoother<Real>& s, const orc_state* st) {
  s.m_dmax = st->dmax;
  s.m_dmax_relative = st->dmax_relative;
  s.m_dnew = st->dnew;
  s.m_dold = st->dold;
  s.m_d0 = st->d0;
  s.m_dmid = st->dmid;
  s.m_alpha = st->alpha;
  s.m_alpha_0 = st->alpha_0;
  s.m_grad_norm_scaled = st->grad_norm_scaled;
  s.m_total_metric = st->total_metric;
  s.m_global_metric = st->global_metric;
  s.m_stage = st->stage;
  s.m_iter = st->iter;
  s.m_untangled = st->untangled;
  s.m_num_invalid = st->num_invalid;
  s.m_num_nodes = st->num_nodes;
}

extern "C" {

int orc_create(orc_ctx** out, int) {
  *out = new orc_ctx();
  return 0;
}
void orc_destroy(orc_ctx* c) { delete c; }
const char* orc_last_error(const orc_ctx* c) { return c->err.c_str(); }

int orc_set_option(orc_ctx* c, const char* key, double v) {
  std::string k(key);
  if (k == "use_ref_mesh")
    c->use_ref_mesh = v != 0;
  else if (k == "real_is_double")
    c->real_is_double = v != 0;
  else if (k == "omp_threads")
    c->omp_threads = (int)v;
  else if (k == "smooth_surfaces") {
    c->smooth_surfaces = v != 0;
    apply_classification(*c);
  }
  else if (k == "strict_fp" || k == "cache_metric" || k == "keep_element_metrics")
    ;
  else {
    c->err = "unknown option " + k;
    return 1;
  }
  return 0;
}

int orc_add_structured_block(orc_ctx* c, const unsigned node_size[3], int* block_id) {
  Block b;
  for (int d = 0; d < 3; d++) b.n[d] = node_size[d];
  b.off = 0;
  c->structured = true;
  c->blocks.push_back(b);
  if (block_id) *block_id = (int)c->blocks.size() - 1;
  return 0;
}
int orc_block_add_boundary_condition(orc_ctx* c, int block, const int rb[3], const int re[3]) {
  BC bc;
  for (int d = 0; d < 3; d++) {
    bc.beg[d] = rb[d];
    bc.end[d] = re[d];
  }
  c->blocks.at(block).bcs.push_back(bc);
  return 0;
}
int orc_block_add_zone_connectivity(orc_ctx* c, int block, int donor, const int tr[3], const int ob[3], const int oe[3],
                                    const int db[3], const int de[3]) {
  Zone z;
  z.donor = donor;
  for (int d = 0; d < 3; d++) {
    z.transform[d] = tr[d];
    z.owner_beg[d] = ob[d];
    z.owner_end[d] = oe[d];
    z.donor_beg[d] = db[d];
    z.donor_end[d] = de[d];
  }
  c->blocks.at(block).zones.push_back(z);
  return 0;
}
int orc_set_unstructured(orc_ctx* c, int topology, size_t nnodes, size_t nelems, const int32_t* conn,
                         const uint8_t* elem_owned, const uint8_t* node_owned) {
  if (topology != 8 && topology != 4) {
    c->err = "topology must be HEX8 or TET4";
    return 1;
  }
  c->structured = false;
  c->topo = topology;
  c->N = nnodes;
  c->nelems = nelems;
  c->conn.assign(conn, conn + nelems * topology);
  c->elem_owned.assign(nelems, 1);
  c->node_owned.assign(nnodes, 1);
  if (elem_owned) c->elem_owned.assign(elem_owned, elem_owned + nelems);
  if (node_owned) c->node_owned.assign(node_owned, node_owned + nnodes);
  return 0;
}
int orc_commit(orc_ctx* c) {
  if (c->structured) {
    size_t off = 0, ne = 0;
    for (Block& b : c->blocks) {
      b.off = off;
      off += b.nnodes();
      ne += b.ncells();
    }
    c->N = off;
    c->nelems = ne;
    c->topo = 8;
  } else {
    // node -> element CSR sorted by (element, local node)
    c->csr_ptr.assign(c->N + 1, 0);
    for (size_t e = 0; e < c->nelems; e++)
      for (int a = 0; a < c->topo; a++) c->csr_ptr[c->conn[e * c->topo + a] + 1]++;
    for (size_t n = 0; n < c->N; n++) c->csr_ptr[n + 1] += c->csr_ptr[n];
    c->csr_ent.assign(c->csr_ptr[c->N], 0);
    std::vector<int64_t> fill(c->csr_ptr.begin(), c->csr_ptr.end() - 1);
    for (size_t e = 0; e < c->nelems; e++)
      for (int a = 0; a < c->topo; a++) c->csr_ent[fill[c->conn[e * c->topo + a]]++] = (int64_t)e * 8 + a;
  }
  if (c->fixed.size() != c->N) c->fixed.assign(c->N, 0);
  for (const char* n : FIELD_NAMES3) {
    Field& f = c->fields[n];
    f.ncomp = 3;
    f.d.assign(3 * c->N, 0.0);
  }
  for (const char* n : FIELD_NAMES1) {
    Field& f = c->fields[n];
    f.ncomp = 1;
    f.d.assign(c->N, 0.0);
  }
  c->committed = true;
  return 0;
}
int orc_set_fixed_mask(orc_ctx* c, const uint8_t* fixed) {
  size_t N = c->N;
  if (c->structured && !c->committed) {
    N = 0;
    for (Block& b : c->blocks) N += b.nnodes();
  }
  if (fixed)
    c->fixed.assign(fixed, fixed + N);
  else
    c->fixed.assign(N, 0);
  return 0;
}
int orc_select_boundary_sgrid(orc_ctx* c) {  // SGridBoundarySelector, ReferenceMeshSmootherConjugateGradient.hpp:31-69
  size_t N = 0;
  for (Block& b : c->blocks) {
    b.off = N;
    N += b.nnodes();
  }
  c->fixed.assign(N, 0);
  for (Block& b : c->blocks)
    for (unsigned k = 0; k < b.n[2]; k++)
      for (unsigned j = 0; j < b.n[1]; j++)
        for (unsigned i = 0; i < b.n[0]; i++) {
          unsigned idx[3] = {i, j, k};
          bool isBC = false;
          for (const BC& bc : b.bcs) {
            unsigned beg[3], end[3];
            for (int d = 0; d < 3; d++) {
              beg[d] = (unsigned)std::min(bc.beg[d], bc.end[d]) - 1;
              end[d] = (unsigned)std::max(bc.beg[d], bc.end[d]) - 1;
            }
            isBC = beg[0] <= idx[0] && idx[0] <= end[0] && beg[1] <= idx[1] && idx[1] <= end[1] && beg[2] <= idx[2] &&
                   idx[2] <= end[2];
            if (isBC) break;
          }
          c->fixed[b.node(i, j, k)] = isBC;
        }
  return 0;
}
int orc_num_nodes_local(const orc_ctx* c, size_t* n) {
  *n = c->N;
  return 0;
}
int orc_num_elements_local(const orc_ctx* c, size_t* n) {
  *n = c->nelems;
  return 0;
}
int orc_get_node_elem_csr(const orc_ctx* c, int64_t* row_ptr, int64_t* entries) {
  if (c->structured) return 1;
  memcpy(row_ptr, c->csr_ptr.data(), c->csr_ptr.size() * sizeof(int64_t));
  memcpy(entries, c->csr_ent.data(), c->csr_ent.size() * sizeof(int64_t));
  return 0;
}

static bool block_range(orc_ctx* c, int block, size_t& off, size_t& n) {
  if (block < 0) {
    off = 0;
    n = c->N;
    return true;
  }
  if (block >= (int)c->blocks.size()) return false;
  off = c->blocks[block].off;
  n = c->blocks[block].nnodes();
  return true;
}
int orc_field_upload(orc_ctx* c, const char* name, int block, const double* host, int layout) {
  auto it = c->fields.find(name);
  size_t off, n;
  if (it == c->fields.end() || !block_range(c, block, off, n)) {
    c->err = std::string("unknown field/block ") + name;
    return 1;
  }
  Field& f = it->second;
  for (int k = 0; k < f.ncomp; k++)
    for (size_t i = 0; i < n; i++) f.d[k * c->N + off + i] = layout == 1 ? host[k * n + i] : host[i * f.ncomp + k];
  return 0;
}
int orc_field_download(orc_ctx* c, const char* name, int block, double* host, int layout) {
  auto it = c->fields.find(name);
  size_t off, n;
  if (it == c->fields.end() || !block_range(c, block, off, n)) {
    c->err = std::string("unknown field/block ") + name;
    return 1;
  }
  Field& f = it->second;
  for (int k = 0; k < f.ncomp; k++)
    for (size_t i = 0; i < n; i++) (layout == 1 ? host[k * n + i] : host[i * f.ncomp + k]) = f.d[k * c->N + off + i];
  return 0;
}
int orc_copy_field(orc_ctx* c, const char* dst, const char* src) {
  ORC_TRY copy_field(*c, dst, src);
  ORC_CATCH
}
int orc_nodal_field_axpby(orc_ctx* c, double alpha, const char* x, double beta, const char* y) {
  ORC_TRY Smoother<double> s(*c);
  s.axpby(alpha, x, beta, y);
  ORC_CATCH
}
int orc_nodal_field_axpbypgz(orc_ctx* c, double alpha, const char* x, double beta, const char* y, double gamma,
                             const char* z) {
  ORC_TRY Field& fx = c->fields.at(x);
  Field& fy = c->fields.at(y);
  Field& fz = c->fields.at(z);
  for (size_t i = 0; i < fz.d.size(); i++) fz.d[i] = alpha * fx.d[i] + beta * fy.d[i] + gamma * fz.d[i];
  ORC_CATCH
}
int orc_nodal_field_dot(orc_ctx* c, const char* x, const char* y, double* result) {
  ORC_TRY if (c->real_is_double) *result = Smoother<double>(*c).dot(x, y);
  else *result = (double)Smoother<LD>(*c).dot(x, y);
  ORC_CATCH
}
// long double result (x87 80-bit) for tolerance checks below double epsilon
int orc_nodal_field_dot_ld(orc_ctx* c, const char* x, const char* y, long double* result) {
  ORC_TRY* result = Smoother<LD>(*c).dot(x, y);
  ORC_CATCH
}
int orc_nodal_field_set_value(orc_ctx* c, const char* name, double value) {
  ORC_TRY Field& f = c->fields.at(name);
  std::fill(f.d.begin(), f.d.end(), value);
  ORC_CATCH
}
int orc_get_number_nodes(orc_ctx* c, size_t* n) {
  ORC_TRY* n = Smoother<double>(*c).get_number_nodes();
  ORC_CATCH
}
int orc_get_edge_lengths(orc_ctx* c) {
  ORC_TRY if (c->real_is_double) Smoother<double>(*c).get_edge_lengths();
  else Smoother<LD>(*c).get_edge_lengths();
  ORC_CATCH
}
int orc_total_metric(orc_ctx* c, int stage, double alpha, double* mtot, size_t* n_invalid) {
  ORC_TRY bool valid;
  if (c->real_is_double)
    *mtot = Smoother<double>(*c).total_metric(stage, alpha, valid, n_invalid, true);
  else
    *mtot = (double)Smoother<LD>(*c).total_metric(stage, alpha, valid, n_invalid, true);
  ORC_CATCH
}
int orc_total_metric_ld(orc_ctx* c, int stage, double alpha, long double* mtot, size_t* n_invalid) {
  ORC_TRY bool valid;
  *mtot = Smoother<LD>(*c).total_metric(stage, alpha, valid, n_invalid, true);
  ORC_CATCH
}
int orc_get_gradient(orc_ctx* c, int stage) {
  ORC_TRY Smoother<double>(*c).get_gradient(stage);
  ORC_CATCH
}
int orc_metric_and_gradient(orc_ctx* c, int stage, double* mtot, size_t* n_invalid) {
  ORC_TRY Smoother<double>(*c).get_gradient(stage);
  bool valid;
  if (c->real_is_double)
    *mtot = Smoother<double>(*c).total_metric(stage, 0.0, valid, n_invalid, true);
  else
    *mtot = (double)Smoother<LD>(*c).total_metric(stage, 0.0, valid, n_invalid, true);
  ORC_CATCH
}
int orc_count_invalid_elements(orc_ctx* c, size_t* n) {
  ORC_TRY* n = Smoother<double>(*c).count_invalid();
  ORC_CATCH
}
int orc_get_alpha_0(orc_ctx* c, double* a) {
  ORC_TRY if (c->real_is_double) *a = Smoother<double>(*c).get_alpha_0();
  else *a = (double)Smoother<LD>(*c).get_alpha_0();
  ORC_CATCH
}
int orc_update_node_positions(orc_ctx* c, double alpha, double* dmax, double* dmax_rel) {
  ORC_TRY if (c->real_is_double) {
    double a = 0, b = 0;
    Smoother<double>(*c).coord_update(alpha, true, a, b);
    *dmax = a;
    *dmax_rel = b;
  }
  else {
    LD a = 0, b = 0;
    Smoother<LD>(*c).coord_update(alpha, true, a, b);
    *dmax = (double)a;
    *dmax_rel = (double)b;
  }
  ORC_CATCH
}
int orc_get_element_metrics(orc_ctx* c, double* metric, int32_t* flag) {
  if (c->elem_metric.size() != c->nelems) {
    c->err = "no element metrics kept";
    return 1;
  }
  if (metric) memcpy(metric, c->elem_metric.data(), c->nelems * sizeof(double));
  if (flag) memcpy(flag, c->elem_flag.data(), c->nelems * sizeof(int32_t));
  return 0;
}
// ---- StructuredGridRefiner ----------------------------------------------------------------------
// Host part of do_refine_structured (structured/StructuredGridRefiner.cpp:40-170): every block gets 2n-1 nodes per
// direction (get_new_sizes, :19-38); the 1-based boundary-condition and zone-connectivity ranges map r -> 2(r-1)+1
// (class member m_index_base = 1, StructuredGridRefiner.hpp:26; :107-154).
int orc_refine_structured_topology(const orc_ctx* coarse, orc_ctx* fine, unsigned long long* num_refined_cells) {
  orc_ctx* c = fine;
  ORC_TRY
  if (!coarse->structured) throw OrcError(2, "refiner input is not a structured grid");
  if (fine->committed || fine->structured || fine->topo != 0) throw OrcError(2, "refiner output ctx must be empty");
  const int base = 1;
  unsigned long long ncell = 0;
  for (const Block& b : coarse->blocks) {
    Block nb;
    for (int d = 0; d < 3; d++) nb.n[d] = 2 * b.n[d] - 1;
    nb.off = 0;
    for (const BC& bc : b.bcs) {
      BC r = bc;
      for (int d = 0; d < 3; d++) {
        r.beg[d] = 2 * (bc.beg[d] - base) + base;
        r.end[d] = 2 * (bc.end[d] - base) + base;
      }
      nb.bcs.push_back(r);
    }
    for (const Zone& z : b.zones) {
      Zone r = z;
      for (int d = 0; d < 3; d++) {
        r.owner_beg[d] = 2 * (z.owner_beg[d] - base) + base;
        r.owner_end[d] = 2 * (z.owner_end[d] - base) + base;
        r.donor_beg[d] = 2 * (z.donor_beg[d] - base) + base;
        r.donor_end[d] = 2 * (z.donor_end[d] - base) + base;
      }
      nb.zones.push_back(r);
    }
    ncell += (unsigned long long)nb.ncells();
    fine->blocks.push_back(nb);
  }
  fine->structured = true;
  if (num_refined_cells) *num_refined_cells = ncell;
  ORC_CATCH
}

// StructuredGridRefinerImpl::operator() (structured/StructuredGridRefinerDef.hpp:134-222), one call per INPUT node:
// vertex copy, the 3 edge midpoints, the 3 face centres and the cell centroid it owns, sums in the reference's order.
int orc_refine_structured_field(orc_ctx* coarse, const char* src_field, orc_ctx* fine, const char* dst_field) {
  orc_ctx* c = fine;
  ORC_TRY
  if (!coarse->committed || !fine->committed) throw OrcError(2, "refiner needs committed grids");
  if (coarse->blocks.size() != fine->blocks.size()) throw OrcError(2, "grids do not match");
  const Field& fi = coarse->fields.at(src_field);
  Field& fo = fine->fields.at(dst_field);
  if (fi.ncomp != 3 || fo.ncomp != 3) throw OrcError(2, "refiner prolongs 3-component nodal fields");
  const size_t Ni = coarse->N, No = fine->N;
  for (size_t ib = 0; ib < coarse->blocks.size(); ib++) {
    const Block& bi = coarse->blocks[ib];
    const Block& bo = fine->blocks[ib];
    for (int d = 0; d < 3; d++)
      if (bo.n[d] != 2 * bi.n[d] - 1) throw OrcError(2, "output block is not the 2x refinement of the input block");
    auto in = [&](unsigned i, unsigned j, unsigned k, unsigned ic) { return fi.d[ic * Ni + bi.node(i, j, k)]; };
    auto out = [&](unsigned i, unsigned j, unsigned k, unsigned ic) -> double& { return fo.d[ic * No + bo.node(i, j, k)]; };
    const unsigned node_max[3] = {bi.n[0] - 1, bi.n[1] - 1, bi.n[2] - 1};
    for (unsigned k = 0; k < bi.n[2]; k++)
      for (unsigned j = 0; j < bi.n[1]; j++)
        for (unsigned i = 0; i < bi.n[0]; i++) {
          const unsigned indx[3] = {i, j, k};
          const unsigned oindx[3] = {2 * i, 2 * j, 2 * k};
          for (unsigned ic = 0; ic < 3; ic++) out(oindx[0], oindx[1], oindx[2], ic) = in(i, j, k, ic);  // vertices
          const unsigned delta[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
          for (unsigned ll = 0; ll < 3; ll++)  // edges
            if (indx[ll] < node_max[ll])
              for (unsigned ic = 0; ic < 3; ic++)
                out(oindx[0] + delta[ll][0], oindx[1] + delta[ll][1], oindx[2] + delta[ll][2], ic) =
                    0.5 * (in(i, j, k, ic) + in(i + delta[ll][0], j + delta[ll][1], k + delta[ll][2], ic));
          for (unsigned ll = 0; ll < 3; ll++) {  // faces
            const unsigned addo[3][3] = {{0, 1, 1}, {1, 0, 1}, {1, 1, 0}};
            const unsigned lc[3] = {ll, (ll + 1) % 3, (ll + 2) % 3};
            if (indx[lc[1]] < node_max[lc[1]] && indx[lc[2]] < node_max[lc[2]])
              for (unsigned ic = 0; ic < 3; ic++) {
                double& o = out(oindx[0] + addo[ll][0], oindx[1] + addo[ll][1], oindx[2] + addo[ll][2], ic);
                o = 0;
                for (unsigned ii = 0; ii < 2; ii++)
                  for (unsigned jj = 0; jj < 2; jj++) {
                    const unsigned add[3][3] = {{0, ii, jj}, {ii, 0, jj}, {ii, jj, 0}};
                    o += 0.25 * in(i + add[ll][0], j + add[ll][1], k + add[ll][2], ic);
                  }
              }
          }
          if (i < node_max[0] && j < node_max[1] && k < node_max[2])  // centroid
            for (unsigned ic = 0; ic < 3; ic++) {
              double& o = out(oindx[0] + 1, oindx[1] + 1, oindx[2] + 1, ic);
              o = 0;
              for (unsigned a = 0; a < 2; a++)
                for (unsigned b = 0; b < 2; b++)
                  for (unsigned cc = 0; cc < 2; cc++) o += 0.125 * in(i + a, j + b, k + cc, ic);
            }
        }
  }
  ORC_CATCH
}

// ---- surface smoothing hooks --------------------------------------------------------------------------
int orc_add_analytic_surface(orc_ctx* c, int kind, const double* params, int* surface_id) {
  ORC_TRY
  if (kind < 0 || kind > 2) throw OrcError(2, "unknown analytic surface kind");
  orc_ctx::Surface s{};
  s.kind = kind;
  for (int i = 0; i < 3; i++) s.p[i] = params[i];
  if (kind == 1) {
    s.R = params[3];
  } else {
    const double l = std::sqrt(params[3] * params[3] + params[4] * params[4] + params[5] * params[5]);
    if (!(l > 0.0)) throw OrcError(2, "analytic surface: zero direction");
    for (int i = 0; i < 3; i++) s.u[i] = params[3 + i] / l;
    s.R = kind == 2 ? params[6] : 0.0;
  }
  c->surfaces.push_back(s);
  if (surface_id) *surface_id = (int)c->surfaces.size() - 1;
  ORC_CATCH
}
int orc_set_node_classification(orc_ctx* c, const int8_t* dof, const int32_t* surface_id) {
  ORC_TRY
  if (!c->committed || c->structured) throw OrcError(2, "node classification needs a committed unstructured mesh (sgrid: not impl)");
  for (size_t n = 0; n < c->N; n++)
    if (dof[n] == 2 && (surface_id[n] < 0 || surface_id[n] >= (int)c->surfaces.size())) throw OrcError(2, "surface node without a surface");
  c->node_dof.assign(dof, dof + c->N);
  c->node_surf.assign(surface_id, surface_id + c->N);
  apply_classification(*c);
  ORC_CATCH
}
int orc_snap_nodes(orc_ctx* c, double* dmax) {
  ORC_TRY
  Smoother<double> s(*c);
  const double d = s.snap_nodes();
  if (dmax) *dmax = d;
  ORC_CATCH
}
// get_surface_normals (Def.hpp:596-706): normal at every non-fixed surface node, zero elsewhere; SoA (3, N)
int orc_get_surface_normals(orc_ctx* c, double* out) {
  ORC_TRY
  const double* X = c->f("coordinates");
  std::fill(out, out + 3 * c->N, 0.0);
  for (size_t n = 0; n < c->N && !c->node_dof.empty(); n++) {
    if (c->fixed[n] || c->node_dof[n] != 2) continue;
    double x[3] = {X[n], X[c->N + n], X[2 * c->N + n]}, nrm[3];
    surf_normal(c->surfaces.at(c->node_surf[n]), x, nrm);
    for (int i = 0; i < 3; i++) out[i * c->N + n] = nrm[i];
  }
  ORC_CATCH
}

int orc_state_init(orc_ctx* c, orc_state* st) {
  memset(st, 0, sizeof(*st));
  st->d0 = 1;
  st->global_metric = std::numeric_limits<double>::max();
  size_t n;
  orc_get_number_nodes(c, &n);
  st->num_nodes = n;
  return 0;
}
int orc_run_one_iteration(orc_ctx* c, orc_state* st, double grad_norm, double* metric_out) {
  ORC_TRY if (c->real_is_double) {
    Smoother<double> s(*c);
    state_from(s, st);
    s.gradNorm = grad_norm;
    *metric_out = s.run_one_iteration();
    s.m_global_metric = *metric_out;
    state_to(s, st);
  }
  else {
    Smoother<LD> s(*c);
    state_from(s, st);
    s.gradNorm = grad_norm;
    *metric_out = (double)s.run_one_iteration();
    s.m_global_metric = *metric_out;
    state_to(s, st);
  }
  ORC_CATCH
}
// line_search(bool& restarted, double mfac_mult) on its own (ReferenceMeshSmootherConjugateGradient.hpp:155, Def.hpp:466-591)
int orc_line_search(orc_ctx* c, orc_state* st, double mfac_mult, int* restarted, double* alpha) {
  ORC_TRY bool r = false;
  if (c->real_is_double) {
    Smoother<double> s(*c);
    state_from(s, st);
    *alpha = s.line_search(r, mfac_mult);
    state_to(s, st);
  } else {
    Smoother<LD> s(*c);
    state_from(s, st);
    *alpha = (double)s.line_search(r, (LD)mfac_mult);
    state_to(s, st);
  }
  if (restarted) *restarted = r ? 1 : 0;
  ORC_CATCH
}
int orc_run_algorithm(orc_ctx* c, int inner_iterations, double grad_norm, const char* log_path, orc_state* st) {
  ORC_TRY if (c->real_is_double) {
    Smoother<double> s(*c);
    s.run_algorithm(inner_iterations, grad_norm, log_path);
    if (st) state_to(s, st);
  }
  else {
    Smoother<LD> s(*c);
    s.run_algorithm(inner_iterations, grad_norm, log_path);
    if (st) state_to(s, st);
  }
  ORC_CATCH
}

// ---------------------------------------------------------------------------------
// fixtures: BlockStructuredGrid::fixture_1 (structured/StructuredBlock.cpp:101-168) and
// build_perturbed_cube (test/regression_tests/HeavyTestMeshSmoother.cpp:93-177).
// out: SoA planes x|y|z of (n+1)^3 nodes, i fastest.  Uses glibc rand() like the test.
// ---------------------------------------------------------------------------------
void orc_fixture_cube(unsigned nx, unsigned ny, unsigned nz, const double width[3], const double offset[3], double* out) {
  size_t N = (size_t)nx * ny * nz;
  unsigned sz[3] = {nx, ny, nz};
  for (unsigned k = 0; k < nz; k++)
    for (unsigned j = 0; j < ny; j++)
      for (unsigned i = 0; i < nx; i++) {
        unsigned idx[3] = {i, j, k};
        size_t n = i + (size_t)nx * (j + (size_t)ny * k);
        for (int ic = 0; ic < 3; ic++) {
          double xyz = double(idx[ic]) / double(sz[ic] - 1);
          out[ic * N + n] = width[ic] * xyz + offset[ic];
        }
      }
}
void orc_perturb_cube_rand(unsigned nele, double epsilon, unsigned seed, double* xyz) {
  const unsigned nn = nele + 1;
  size_t N = (size_t)nn * nn * nn;
  const double tol = 1e-6;
  std::srand(seed);
  for (unsigned k = 0; k < nn; k++)
    for (unsigned j = 0; j < nn; j++)
      for (unsigned i = 0; i < nn; i++) {
        size_t n = i + (size_t)nn * (j + (size_t)nn * k);
        double data[3] = {xyz[n], xyz[N + n], xyz[2 * N + n]};
        const double x = data[0], y = data[1], z = data[2];
        if (x > tol && x < 1 - tol && y > tol && y < 1 - tol && z > tol && z < 1 - tol)
          for (int d = 0; d < 3; d++) data[d] = data[d] + epsilon * ((double)std::rand() / (double)RAND_MAX - 0.5);
        for (int d = 0; d < 3; d++) xyz[d * N + n] = data[d];
      }
}

// ---------------------------------------------------------------------------------
// CPU baseline ("Kokkos-OpenMP-equivalent" restatement): same loop structure as the
// reference functors — cell-parallel metric parallel_reduce
// (total_element_metric.cpp:271-307), cell-parallel gradient with atomic scatter
// (gradient_functors.hpp:349-407 -> #pragma omp atomic), long double reduction.
// Returns seconds for `reps` repetitions of {get_gradient ; total_metric(0)}.
// ---------------------------------------------------------------------------------
double orc_bench_metric_and_gradient(orc_ctx* c, int stage, int reps, double* mtot_out, size_t* ninv_out,
                                     double* t_metric, double* t_gradient) {
  const double* X = c->f("coordinates");
  const double* W = c->f("coordinates_NM1");
  double* g = c->f("cg_g");
  const size_t N = c->N;
  double tm = 0, tg = 0;
  LD mtot = 0;
  size_t ninv = 0;
#ifdef _OPENMP
  if (c->omp_threads > 0) omp_set_num_threads(c->omp_threads);
#endif
  auto now = []() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
  };
  for (int rep = 0; rep < reps; rep++) {
    double t0 = now();
    // ---- gradient ----
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < 3 * N; i++) g[i] = 0.0;
    if (c->structured) {
      HexMeshSmootherMetric hm;
      hm.m_untangling = stage == 0;
      for (const Block& B : c->blocks) {
        const long long nci = B.n[0] - 1, ncj = B.n[1] - 1, nck = B.n[2] - 1;
#pragma omp parallel for schedule(static)
        for (long long idx = 0; idx < nci * ncj * nck; idx++) {
          unsigned i = idx % nci, j = (idx / nci) % ncj, k = idx / (nci * ncj);
          double vc[8][3], vo[8][3], grad[8][3];
          size_t nid[8];
          gather_cell(*c, B, X, i, j, k, vc, nid);
          gather_cell(*c, B, W, i, j, k, vo, nid);
          bool gmvalid = true;
          hm.grad_metric(vc, vo, gmvalid, grad);
          if (gmvalid || hm.m_untangling)
            for (int in = 0; in < 8; in++)
              for (int d = 0; d < 3; d++) {
#pragma omp atomic
                g[d * N + nid[in]] += grad[in][d];
              }
        }
#pragma omp parallel for schedule(static)
        for (size_t n = B.off; n < B.off + B.nnodes(); n++)
          if (c->fixed[n])
            for (int d = 0; d < 3; d++) g[d * N + n] = 0.0;
      }
    } else {
      const int topo = c->topo;
#pragma omp parallel for schedule(static)
      for (long long e = 0; e < (long long)c->nelems; e++) {
        double vc[8][3], vo[8][3], grad[8][3];
        gather_elem(*c, X, e, vc);
        gather_elem(*c, W, e, vo);
        bool gmvalid = true;
        if (stage == 0)
          stk_untangle_grad_metric(topo, vc, vo, gmvalid, grad);
        else
          stk_shapeb1_grad_metric(topo, c->use_ref_mesh, vc, vo, gmvalid, grad);
        if (gmvalid || stage == 0)
          for (int in = 0; in < topo; in++) {
            size_t n = c->conn[e * topo + in];
            if (c->fixed[n]) continue;
            for (int d = 0; d < 3; d++) {
#pragma omp atomic
              g[d * N + n] += grad[in][d];
            }
          }
      }
    }
    double t1 = now();
    // ---- metric ----
    mtot = 0;
    ninv = 0;
    if (c->structured) {
      HexMeshSmootherMetric hm;
      hm.m_untangling = stage == 0;
      hm.m_use_ref_mesh = c->use_ref_mesh;
      for (const Block& B : c->blocks) {
        const long long nci = B.n[0] - 1, ncj = B.n[1] - 1, nck = B.n[2] - 1;
        LD bt = 0;
        size_t bi = 0;
#pragma omp parallel for schedule(static) reduction(+ : bt, bi)
        for (long long idx = 0; idx < nci * ncj * nck; idx++) {
          unsigned i = idx % nci, j = (idx / nci) % ncj, k = idx / (nci * ncj);
          double vc[8][3], vo[8][3];
          size_t nid[8];
          gather_cell(*c, B, X, i, j, k, vc, nid);
          gather_cell(*c, B, W, i, j, k, vo, nid);
          bool valid = false;
          bt += hm.metric(vc, vo, valid);
          bi += !valid;
        }
        mtot += bt;
        ninv += bi;
      }
    } else {
      const int topo = c->topo;
      LD bt = 0;
      size_t bi = 0;
#pragma omp parallel for schedule(static) reduction(+ : bt, bi)
      for (long long e = 0; e < (long long)c->nelems; e++) {
        if (!c->elem_owned[e]) continue;
        double vc[8][3], vo[8][3];
        gather_elem(*c, X, e, vc);
        gather_elem(*c, W, e, vo);
        bool valid = false;
        bt += stage == 0 ? stk_untangle_metric(topo, vc, vo, valid) : stk_shapeb1_metric(topo, c->use_ref_mesh, vc, vo, valid);
        bi += !valid;
      }
      mtot = bt;
      ninv = bi;
    }
    double t2 = now();
    tg += t1 - t0;
    tm += t2 - t1;
  }
  if (mtot_out) *mtot_out = (double)mtot;
  if (ninv_out) *ninv_out = ninv;
  if (t_metric) *t_metric = tm;
  if (t_gradient) *t_gradient = tg;
  return tm + tg;
}

int orc_omp_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
