// C++ host-API test: reads like the reference's own tests
//   heavy_perceptMeshSmoother.total_metric_and_gradient  (test/regression_tests/HeavyTestMeshSmoother.cpp:1344-1429)
//   heavy_perceptMeshSmoother.sgrid_perturbed_smooth     (:1431-1468)
// but against percept_b200::ReferenceMeshSmootherConjugateGradientImpl<StructuredGrid> (CUDA path behind the C ABI).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "percept_b200/ReferenceMeshSmootherConjugateGradient.hpp"

using namespace percept_b200;

static int failures = 0;
#define EXPECT_NEAR(a, b, tol)                                                                    \
  do {                                                                                            \
    if (!(std::fabs((double)(a) - (double)(b)) <= (tol))) {                                       \
      std::printf("EXPECT_NEAR failed %s:%d: %.15g vs %.15g (tol %g)\n", __FILE__, __LINE__, (double)(a), (double)(b), (double)(tol)); \
      failures++;                                                                                 \
    }                                                                                             \
  } while (0)
#define EXPECT_EQ(a, b)                                                                           \
  do {                                                                                            \
    if (!((a) == (b))) {                                                                          \
      std::printf("EXPECT_EQ failed %s:%d: %lld vs %lld\n", __FILE__, __LINE__, (long long)(a), (long long)(b)); \
      failures++;                                                                                 \
    }                                                                                             \
  } while (0)

// build_perturbed_cube (HeavyTestMeshSmoother.cpp:93-177)
static void build_perturbed_cube(PerceptMesh& eMesh, const unsigned nele) {
  const unsigned nxyz = nele + 1;
  std::array<unsigned, 3> sizes{{nxyz, nxyz, nxyz}};
  auto bsg = BlockStructuredGrid::fixture_1(sizes);
  eMesh.set_block_structured_grid(bsg);
  eMesh.add_coordinate_state_fields();
  const size_t N = eMesh.num_nodes_local();
  std::vector<double> x(3 * N);
  eMesh.get_field("coordinates", x.data());
  eMesh.set_field("coordinates_NM1", x.data());  // save state of original mesh
  pms_fixture_perturb_rand(nxyz, nxyz, nxyz, 1.0 / (double)nele, 1, x.data());
  eMesh.set_field("coordinates", x.data());
  eMesh.set_field("coordinates_N", x.data());  // save state of projected mesh
}

int main() {
  {  // TEST(heavy_perceptMeshSmoother, total_metric_and_gradient), nele = 100
    const unsigned nele = 100;
    PerceptMesh eMesh(3u);
    build_perturbed_cube(eMesh, nele);
    ReferenceMeshSmootherConjugateGradientImpl<StructuredGrid> sm(&eMesh, NULL, NULL, 0, 1, 1e-3, 1);  // no selector: nothing fixed
    sm.m_stage = 1;
    sm.get_gradient();
    double g = std::sqrt(eMesh.nodal_field_dot("cg_g", "cg_g"));
    EXPECT_NEAR(g, 389575833.965, 1.0);
    sm.m_stage = 0;
    sm.get_gradient();
    g = std::sqrt(eMesh.nodal_field_dot("cg_g", "cg_g"));
    EXPECT_NEAR(g, 1.82363504678e-07, 1e-6);
    bool valid = true;
    size_t n_invalid = 0;
    Double mtot = sm.total_metric(0.0, 1.0, valid, &n_invalid);
    EXPECT_NEAR(mtot, 5.52566e-08, 1.e-7);
    EXPECT_EQ(n_invalid, (size_t)391365);
    EXPECT_EQ(valid, false);
    EXPECT_EQ(ReferenceMeshSmootherConjugateGradientImpl<StructuredGrid>::parallel_count_invalid_elements(&eMesh), (size_t)391365);
    std::printf("total_metric_and_gradient: |g|_smooth-golden ok, mtot=%.6e n_invalid=%zu\n", (double)mtot, n_invalid);
  }
  {  // TEST(heavy_perceptMeshSmoother, sgrid_perturbed_smooth), nele = 25, tol 1e-3, up to 1001 iterations/stage
    const unsigned nele = 25;
    PerceptMesh eMesh(3u);
    build_perturbed_cube(eMesh, nele);
    eMesh.setProperty("in_filename", "perturbed_25x25x25");
    SGridBoundarySelector boundarySelector(&eMesh);
    ReferenceMeshSmootherConjugateGradientImpl<StructuredGrid> pmmpsi(&eMesh, NULL, &boundarySelector, 0, 1001, 1.e-3, 1);
    pmmpsi.m_do_animation = 0;
    const size_t inv0 = pmmpsi.parallel_count_invalid_elements(&eMesh);
    pmmpsi.run();
    EXPECT_EQ(pmmpsi.m_num_invalid, (size_t)0);
    EXPECT_EQ(pmmpsi.m_untangled, true);
    EXPECT_EQ(pmmpsi.m_stage, 1);
    std::printf("sgrid_perturbed_smooth: %zu invalid -> 0, final metric %.6e after stage %d iter %d, dmax_rel %.3e\n", inv0,
                (double)pmmpsi.m_global_metric, pmmpsi.m_stage, pmmpsi.m_iter, (double)pmmpsi.m_dmax_relative);
    // the smoothed mesh is close to the reference cube again (the perturbation was the only distortion)
    const size_t N = eMesh.num_nodes_local();
    std::vector<double> x(3 * N), w(3 * N);
    eMesh.get_field("coordinates", x.data());
    eMesh.get_field("coordinates_NM1", w.data());
    double maxd = 0;
    for (size_t i = 0; i < 3 * N; i++) maxd = std::fmax(maxd, std::fabs(x[i] - w[i]));
    EXPECT_NEAR(maxd, 0.0, 0.02 / nele);
    std::remove("smootherlogperturbed_25x25x25.txt");
  }
  {  // error convention: failures surface as std::runtime_error
    bool threw = false;
    try {
      PerceptMesh eMesh(3u);
      eMesh.nodal_field_dot("cg_g", "cg_g");  // no mesh / no fields yet
    } catch (const std::runtime_error& e) {
      threw = true;
    }
    EXPECT_EQ(threw, true);
  }
  std::printf(failures ? "FAILED %d\n" : "ALL PASSED\n", failures);
  return failures ? 1 : 0;
}
