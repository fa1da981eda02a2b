/* percept_fixtures_c.h — synthetic meshes used by the reference's own smoother tests (host helpers,
 * no GPU work).  Outputs are SoA planes x|y|z, i fastest (the layout pms_field_upload takes as PMS_SOA). */
#ifndef PERCEPT_FIXTURES_C_H
#define PERCEPT_FIXTURES_C_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
/* BlockStructuredGrid::fixture_1 / StructuredBlock::fixture_1 (structured/StructuredBlock.cpp:101-168). */
void pms_fixture_cube(unsigned nx, unsigned ny, unsigned nz, const double width[3], const double offset[3], double* out);
/* build_perturbed_cube (test/regression_tests/HeavyTestMeshSmoother.cpp:124-159): strictly interior nodes of the
 * unit cube get x_d += epsilon*(rand()/RAND_MAX - 0.5), loop order k,j,i,d after srand(seed). */
void pms_fixture_perturb_rand(unsigned nx, unsigned ny, unsigned nz, double epsilon, unsigned seed, double* xyz);
#ifdef __cplusplus
}
#endif
#endif
