// C++ host-API test: reads like the reference's own tests
//   heavy_perceptMeshSmoother.total_metric_and_gradient  (test/regression_tests/HeavyTestMeshSmoother.cpp:1344-1429)
//   heavy_perceptMeshSmoother.sgrid_perturbed_smooth     (:1431-1468)
// but against percept_b200::ReferenceMeshSmootherConjugateGradientImpl<StructuredGrid> (CUDA path behind the C ABI).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
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
  {  // the same test on the STK side (HeavyTestMeshSmoother.cpp:1376-1404 runs the STK hex8 mesh of the same perturbed
     // cube and expects the same numbers): unstructured hex8 in Exodus order through the C++ facade, no Python in between
    const unsigned nele = 100, nn = nele + 1;
    const size_t N = (size_t)nn * nn * nn;
    std::vector<double> w(3 * N), x;
    const double width[3] = {1, 1, 1}, offset[3] = {0, 0, 0};
    pms_fixture_cube(nn, nn, nn, width, offset, w.data());
    x = w;
    pms_fixture_perturb_rand(nn, nn, nn, 1.0 / (double)nele, 1, x.data());
    std::vector<int32_t> conn;
    conn.reserve((size_t)8 * nele * nele * nele);
    for (unsigned k = 0; k < nele; k++)
      for (unsigned j = 0; j < nele; j++)
        for (unsigned i = 0; i < nele; i++) {
          auto id = [&](unsigned a, unsigned b, unsigned c) { return (int32_t)((i + a) + nn * ((j + b) + nn * (k + c))); };
          for (int32_t v : {id(0, 0, 0), id(1, 0, 0), id(1, 1, 0), id(0, 1, 0), id(0, 0, 1), id(1, 0, 1), id(1, 1, 1), id(0, 1, 1)}) conn.push_back(v);
        }
    PerceptMesh eMesh(3u);
    eMesh.set_unstructured_mesh(PMS_HEX8, N, conn.size() / 8, conn.data());
    eMesh.add_coordinate_state_fields();
    eMesh.set_field("coordinates_NM1", w.data());
    eMesh.set_field("coordinates", x.data());
    eMesh.set_field("coordinates_N", x.data());
    ReferenceMeshSmootherConjugateGradientImpl<STKMesh> sm(&eMesh, NULL, NULL, 0, 1, 1e-3, 1);
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
    EXPECT_EQ(ReferenceMeshSmootherConjugateGradientImpl<STKMesh>::parallel_count_invalid_elements(&eMesh), (size_t)391365);
    std::printf("total_metric_and_gradient (STK hex8): mtot=%.6e n_invalid=%zu\n", (double)mtot, n_invalid);
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
  {  // TEST(regr_PerceptCGNS, test6_sgrid_refine) (RegressionTestPerceptCGNS.cpp:409-429): fixture -> refine -> smooth,
     // and the refine-then-smooth set-up of SGridSmootherMultiBlockvsSingleBlock.cpp:150-175
    PerceptMesh eMesh(3u);
    std::array<unsigned, 3> sizes{{4, 4, 4}};
    std::shared_ptr<BlockStructuredGrid> bsg = BlockStructuredGrid::fixture_1(sizes);
    bsg->m_sblocks[0]->dump_vtk("cube", 0, 1);
    StructuredGridRefiner sgr1(bsg, 0);
    sgr1.m_input->dump_vtk("cube_init");
    const unsigned ncell = sgr1.do_refine();
    sgr1.m_output->dump_vtk("cube_ref");
    EXPECT_EQ(ncell, 6u * 6u * 6u);
    EXPECT_EQ(sgr1.m_output->m_sblocks[0]->node_size[0], 7u);
    EXPECT_EQ(sgr1.m_output->m_sblocks[0]->m_boundaryConditions[1].beg[0], 7);  // imax face: 4 -> 7
    EXPECT_EQ(sgr1.m_output->m_sblocks[0]->m_boundaryConditions[5].end[2], 7);
    StructuredGridRefiner sgr2(sgr1.m_output, 0);  // refine again: 13^3 nodes
    EXPECT_EQ(sgr2.do_refine(), 12u * 12u * 12u);
    bsg = sgr2.m_output;
    {  // twice-refined fixture == finer fixture (edge/face/centroid averages of a uniform lattice)
      auto fine = BlockStructuredGrid::fixture_1({{13, 13, 13}});
      double maxd = 0;
      for (size_t i = 0; i < fine->m_sblocks[0]->m_sgrid_coords.size(); i++)
        maxd = std::fmax(maxd, std::fabs(fine->m_sblocks[0]->m_sgrid_coords[i] - bsg->m_sblocks[0]->m_sgrid_coords[i]));
      EXPECT_NEAR(maxd, 0.0, 2.3e-16);
    }
    // do_smooth_block_structured (RegressionTestPerceptCGNS.cpp:381-407): perturb the interior, smooth with fixed boundary
    eMesh.set_block_structured_grid(bsg);
    eMesh.add_coordinate_state_fields();
    const size_t N = eMesh.num_nodes_local();
    std::vector<double> x(3 * N);
    eMesh.get_field("coordinates", x.data());
    eMesh.set_field("coordinates_NM1", x.data());
    pms_fixture_perturb_rand(13, 13, 13, 0.25 / 12, 1, x.data());
    eMesh.set_field("coordinates", x.data());
    eMesh.set_field("coordinates_N", x.data());
    eMesh.setProperty("in_filename", "cube_ref");
    SGridBoundarySelector boundarySelector(&eMesh);
    ReferenceMeshSmootherConjugateGradientImpl<StructuredGrid> pmmpsi(&eMesh, NULL, &boundarySelector, 0, 200, 1.e-4, 1);
    pmmpsi.run();
    EXPECT_EQ(pmmpsi.m_num_invalid, (size_t)0);
    std::vector<double> w(3 * N);
    eMesh.get_field("coordinates", x.data());
    eMesh.get_field("coordinates_NM1", w.data());
    double maxd = 0;
    for (size_t i = 0; i < 3 * N; i++) maxd = std::fmax(maxd, std::fabs(x[i] - w[i]));
    EXPECT_NEAR(maxd, 0.0, 0.02 / 12);
    std::printf("sgrid_refine: 3^3 -> 6^3 -> 12^3 cells, smoothed back to the lattice within %.2e\n", maxd);
    // the .vts piece is the reference's ASCII layout: header + one "%16.4f" triple per node
    std::ifstream vts("cube_ref.0.vts");
    std::string line;
    std::getline(vts, line);
    EXPECT_EQ(line.find("<VTKFile type=\"StructuredGrid\"") == 0, true);
    std::getline(vts, line);
    EXPECT_EQ(line.find("WholeExtent=\"0 6 0 6 0 6\"") != std::string::npos, true);
    for (const char* f : {"cube.0.vts", "cube_init.pvd", "cube_init.0.vts", "cube_ref.pvd", "cube_ref.0.vts", "smootherlogcube_ref.txt"}) std::remove(f);
  }
  {  // STKMesh + MeshGeometry with smooth_surfaces: hex8 box whose top face is a cylinder patch; face-interior nodes
     // are MS_SURFACE (slide on their surface), box edges MS_CURVE and corners MS_VERTEX stay put (MeshSmoother.cpp:293-340)
    const unsigned ne = 6, nn = ne + 1;
    const size_t N = (size_t)nn * nn * nn;
    std::vector<double> w(3 * N), x(3 * N);
    std::vector<int8_t> dof(N);
    std::vector<int32_t> ev(N, -1);
    auto ztop = [](double xx) { return -2.0 + std::sqrt(9.0 - (xx - 0.5) * (xx - 0.5)); };
    std::srand(7);
    for (unsigned k = 0; k < nn; k++)
      for (unsigned j = 0; j < nn; j++)
        for (unsigned i = 0; i < nn; i++) {
          const size_t n = i + (size_t)nn * (j + (size_t)nn * k);
          const double X = (double)i / ne, Y = (double)j / ne, Z = (double)k / ne * ztop(X);
          w[n] = X, w[N + n] = Y, w[2 * N + n] = Z;
          const int nb = (i == 0 || i == ne) + (j == 0 || j == ne) + (k == 0 || k == ne);
          dof[n] = nb == 3 ? 0 : nb == 2 ? 1 : nb == 1 ? 2 : 3;
          if (nb == 1) ev[n] = i == 0 ? 0 : i == ne ? 1 : j == 0 ? 2 : j == ne ? 3 : k == 0 ? 4 : 5;
          for (int d = 0; d < 3; d++)
            x[d * N + n] = w[d * N + n] + (nb <= 1 ? 0.2 / ne * ((double)std::rand() / RAND_MAX - 0.5) : 0.0);
        }
    std::vector<int32_t> conn;
    for (unsigned k = 0; k < ne; k++)
      for (unsigned j = 0; j < ne; j++)
        for (unsigned i = 0; i < ne; i++) {
          auto id = [&](unsigned a, unsigned b, unsigned c) { return (int32_t)((i + a) + nn * ((j + b) + nn * (k + c))); };
          for (int32_t v : {id(0, 0, 0), id(1, 0, 0), id(1, 1, 0), id(0, 1, 0), id(0, 0, 1), id(1, 0, 1), id(1, 1, 1), id(0, 1, 1)}) conn.push_back(v);
        }
    MeshGeometry geom;
    geom.add_plane({{0, 0, 0}}, {{-1, 0, 0}});
    geom.add_plane({{1, 0, 0}}, {{1, 0, 0}});
    geom.add_plane({{0, 0, 0}}, {{0, -1, 0}});
    geom.add_plane({{0, 1, 0}}, {{0, 1, 0}});
    geom.add_plane({{0, 0, 0}}, {{0, 0, -1}});
    geom.add_cylinder({{0.5, 0, -2}}, {{0, 1, 0}}, 3.0);
    geom.set_classification(dof, ev);
    PerceptMesh eMesh(3u);
    eMesh.set_unstructured_mesh(PMS_HEX8, N, conn.size() / 8, conn.data());
    eMesh.add_coordinate_state_fields();
    eMesh.set_smooth_surfaces(true);
    eMesh.set_field("coordinates_NM1", w.data());
    eMesh.set_field("coordinates", x.data());
    eMesh.setProperty("in_filename", "curved_box");
    ReferenceMeshSmootherConjugateGradientImpl<STKMesh> sm(&eMesh, NULL, NULL, &geom, 15, 1.e-6, 1);
    EXPECT_EQ(sm.get_fixed_flag((size_t)0).second, (int)MS_VERTEX);
    EXPECT_EQ(sm.get_fixed_flag((size_t)(nn / 2)).second, (int)MS_CURVE);
    EXPECT_EQ(sm.get_fixed_flag((size_t)(nn / 2 + nn * (nn / 2))).first, false);  // bottom-face interior: MS_SURFACE, free
    EXPECT_EQ(sm.get_fixed_flag((size_t)(nn / 2 + nn * (nn / 2))).second, (int)MS_SURFACE);
    sm.snap_nodes();  // the perturbed start goes onto the geometry (Refiner::snapAndSmooth snaps before smoothing)
    eMesh.copy_field("coordinates_N", "coordinates");
    sm.m_stage = 1;
    bool valid = true;
    const Double m0 = sm.total_metric(0.0, 1.0, valid);
    sm.run();
    EXPECT_EQ(sm.m_num_invalid, (size_t)0);
    EXPECT_EQ(sm.m_global_metric < m0, true);
    eMesh.get_field("coordinates", x.data());
    double off_surface = 0, moved_fixed = 0;
    for (size_t n = 0; n < N; n++) {
      if (dof[n] < 2)
        for (int d = 0; d < 3; d++) moved_fixed = std::fmax(moved_fixed, std::fabs(x[d * N + n] - w[d * N + n]));
      if (ev[n] == 4) off_surface = std::fmax(off_surface, std::fabs(x[2 * N + n]));
      if (ev[n] == 5) off_surface = std::fmax(off_surface, std::fabs(std::hypot(x[n] - 0.5, x[2 * N + n] + 2.0) - 3.0));
    }
    EXPECT_NEAR(off_surface, 0.0, 4e-15);
    EXPECT_EQ(moved_fixed, 0.0);
    sm.get_surface_normals(&eMesh);
    const size_t top = nn / 2 + nn * (nn / 2 + (size_t)nn * ne);  // a node of the cylinder patch: unit radial normal
    EXPECT_NEAR(std::hypot(sm.m_surface_normals[top], sm.m_surface_normals[2 * N + top]), 1.0, 1e-15);
    std::printf("smooth_surfaces: metric %.6e -> %.6e, surface nodes within %.1e of the geometry\n", (double)m0,
                (double)sm.m_global_metric, off_surface);
    std::remove("smootherlogcurved_box.txt");
  }
  {  // the virtual members are live: a class derived from the smoother that overrides total_metric / get_gradient /
     // update_node_positions sees them called by run_one_iteration() and line_search() (Def.hpp:1010-1093, :466-591),
     // and gets the same iterates as the exact class (which runs an iteration as one C call)
    struct Counting : ReferenceMeshSmootherConjugateGradientImpl<StructuredGrid> {
      using ReferenceMeshSmootherConjugateGradientImpl<StructuredGrid>::ReferenceMeshSmootherConjugateGradientImpl;
      int n_metric = 0, n_grad = 0, n_update = 0;
      Double total_metric(Double alpha, double scale, bool& valid, size_t* num_invalid = 0) override {
        n_metric++;
        return ReferenceMeshSmootherConjugateGradientImpl<StructuredGrid>::total_metric(alpha, scale, valid, num_invalid);
      }
      void get_gradient() override {
        n_grad++;
        ReferenceMeshSmootherConjugateGradientImpl<StructuredGrid>::get_gradient();
      }
      void update_node_positions(Double alpha) override {
        n_update++;
        ReferenceMeshSmootherConjugateGradientImpl<StructuredGrid>::update_node_positions(alpha);
      }
    };
    const unsigned nele = 12;
    std::vector<double> xa, xb;
    Double ma = 0, mb = 0;
    for (int derived = 0; derived < 2; derived++) {
      PerceptMesh eMesh(3u);
      build_perturbed_cube(eMesh, nele);
      const size_t N = eMesh.num_nodes_local();
      std::vector<double> x(3 * N);
      eMesh.get_field("coordinates_NM1", x.data());
      pms_fixture_perturb_rand(nele + 1, nele + 1, nele + 1, 0.3 / (double)nele, 1, x.data());  // a valid start
      eMesh.set_field("coordinates", x.data());
      eMesh.set_field("coordinates_N", x.data());
      eMesh.setProperty("in_filename", "virtuals");
      SGridBoundarySelector boundarySelector(&eMesh);
      if (derived) {
        Counting sm(&eMesh, NULL, &boundarySelector, 0, 4, 0.0, 1);
        sm.run();
        EXPECT_EQ(sm.n_grad >= 5, true);       // one gradient per CG iteration (stage 0 stops early on a valid mesh, stage 1 runs its 4)
        EXPECT_EQ(sm.n_metric > sm.n_grad, true);  // the line search's trials went through the override
        EXPECT_EQ(sm.n_update >= 4, true);
        mb = sm.m_global_metric;
        // line_search on its own: a descent step along -g that lowers the metric and leaves the nodes where they are
        sm.m_stage = 1;
        sm.get_gradient();
        eMesh.nodal_field_axpby(-1.0, "cg_g", 0.0, "cg_r");
        eMesh.copy_field("cg_s", "cg_r");
        std::vector<double> before(3 * N), after(3 * N);
        eMesh.get_field("coordinates", before.data());
        bool restarted = false, valid = true;
        const int nm0 = sm.n_metric;
        const Double alpha = sm.line_search(restarted, 1.0);
        eMesh.get_field("coordinates", after.data());
        EXPECT_EQ(alpha > 0.0, true);
        EXPECT_EQ(sm.n_metric > nm0, true);
        EXPECT_EQ(before == after, true);
        EXPECT_EQ(sm.total_metric(alpha, 1.0, valid) <= sm.total_metric(0.0, 1.0, valid), true);
        xb.resize(3 * N);
        eMesh.get_field("coordinates", xb.data());
        std::printf("virtual dispatch: %d total_metric, %d get_gradient, %d update_node_positions calls; line_search alpha %.3e\n",
                    sm.n_metric, sm.n_grad, sm.n_update, (double)alpha);
      } else {
        ReferenceMeshSmootherConjugateGradientImpl<StructuredGrid> sm(&eMesh, NULL, &boundarySelector, 0, 4, 0.0, 1);
        sm.run();
        ma = sm.m_global_metric;
        // the exact class: line_search through pms_line_search, same contract
        sm.m_stage = 1;
        sm.get_gradient();
        eMesh.nodal_field_axpby(-1.0, "cg_g", 0.0, "cg_r");
        eMesh.copy_field("cg_s", "cg_r");
        bool restarted = false;
        const Double alpha = sm.line_search(restarted, 1.0);
        EXPECT_EQ(alpha > 0.0, true);
        xa.resize(3 * N);
        eMesh.get_field("coordinates", xa.data());
        // helpers of the reference's public surface
        bool valid = false;
        const Double m_el = sm.metric((size_t)0, valid);
        EXPECT_EQ(valid, true);
        EXPECT_EQ(m_el >= 0.0, true);
        EXPECT_NEAR(sm.nodal_edge_length_ave((size_t)(N / 2)), 1.0 / nele, 1e-12);
      }
      std::remove("smootherlogvirtuals.txt");
    }
    double maxd = 0;
    for (size_t i = 0; i < xa.size(); i++) maxd = std::fmax(maxd, std::fabs(xa[i] - xb[i]));
    EXPECT_NEAR(maxd, 0.0, 1e-9);
    EXPECT_NEAR(mb, ma, 1e-9 * std::fabs((double)ma));
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
