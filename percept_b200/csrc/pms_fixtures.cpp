// pms_fixtures.cpp — synthetic test/benchmark meshes of the reference's own tests (host-only helpers).
//   fixture_1 cube:        structured/StructuredBlock.cpp:101-168 (node (i,j,k) at width*idx/(n-1)+offset)
//   build_perturbed_cube:  test/regression_tests/HeavyTestMeshSmoother.cpp:93-177 (glibc srand/rand stream)
#include <cstdlib>

#include "../../include/percept_fixtures_c.h"

extern "C" {

void pms_fixture_cube(unsigned nx, unsigned ny, unsigned nz, const double width[3], const double offset[3], double* out) {
  const size_t N = (size_t)nx * ny * nz;
  const unsigned sz[3] = {nx, ny, nz};
  for (unsigned k = 0; k < nz; k++)
    for (unsigned j = 0; j < ny; j++)
      for (unsigned i = 0; i < nx; i++) {
        const unsigned idx[3] = {i, j, k};
        const size_t n = i + (size_t)nx * (j + (size_t)ny * k);
        for (int ic = 0; ic < 3; ic++) out[ic * N + n] = width[ic] * (double(idx[ic]) / double(sz[ic] - 1)) + offset[ic];
      }
}

void pms_fixture_perturb_rand(unsigned nx, unsigned ny, unsigned nz, double epsilon, unsigned seed, double* xyz) {
  const size_t N = (size_t)nx * ny * nz;
  const double tol = 1e-6;
  std::srand(seed);
  for (unsigned k = 0; k < nz; k++)
    for (unsigned j = 0; j < ny; j++)
      for (unsigned i = 0; i < nx; i++) {
        const size_t n = i + (size_t)nx * (j + (size_t)ny * k);
        double data[3] = {xyz[n], xyz[N + n], xyz[2 * N + n]};
        const double x = data[0], y = data[1], z = data[2];
        if (x > tol && x < 1 - tol && y > tol && y < 1 - tol && z > tol && z < 1 - tol)
          for (int d = 0; d < 3; d++) data[d] = data[d] + epsilon * ((double)std::rand() / (double)RAND_MAX - 0.5);
        for (int d = 0; d < 3; d++) xyz[d * N + n] = data[d];
      }
}

}  // extern "C"
