"""Pins the CPU oracle to the reference's own golden values.

Source of the numbers: /root/reference/test/regression_tests/HeavyTestMeshSmoother.cpp
  :1352-1358  heavy_perceptMeshSmoother.total_metric_and_gradient (nele = 100, 200, 300, 400)
  :809-810    bump mesh nele=5 untangle metric 0.0179978, n_invalid 25
The tolerances are the reference test's own EXPECT_NEAR tolerances (:1393, :1402, :1423) —
the printed goldens carry 12 significant digits, which we additionally check to 1e-11 relative.
"""
import numpy as np
import pytest

from tests.orc import oracle, fixture_cube, perturb_cube_rand

GOLD = {  # nele: (n_invalid, mtot, |g|_untangle, |g|_smooth)
    100: (391365, 5.52566e-08, 1.82363504678e-07, 389575833.965),
    200: (3189265, 7.08e-09, 1.63634415837e-08, 1124552264.63),
    300: (10820052, 2.10961e-09, 3.97221589862e-09, 89552984978.3),
}


def build_perturbed_cube(api, nele):
    W = fixture_cube(nele + 1)
    X = perturb_cube_rand(W, nele, 1.0 / nele, 1)
    c = api.create()
    b = c.add_structured_block((nele + 1,) * 3)
    c.add_fixture_1_bcs(b)
    c.commit()  # no boundary selector => nothing fixed (gg_1_sgrid(&eMesh, NULL), :1380)
    c.upload("coordinates", X)
    c.upload("coordinates_N", X)
    c.upload("coordinates_NM1", W)
    return c


@pytest.mark.parametrize("nele", [100, 200, 300])
def test_total_metric_and_gradient_golden(nele):
    n_inv, mtot, g_unt, g_smooth = GOLD[nele]
    c = build_perturbed_cube(oracle, nele)
    c.get_gradient(1)
    g = np.sqrt(c.dot("cg_g", "cg_g"))
    assert abs(g - g_smooth) < 1.0 * max(1.0, g_smooth / 389575833.965)  # (the reference's 1.0 at nele = 100, scaled with |g|)
    assert abs(g - g_smooth) / g_smooth < 1e-11
    c.get_gradient(0)
    g = np.sqrt(c.dot("cg_g", "cg_g"))
    assert abs(g - g_unt) < 1e-6
    assert abs(g - g_unt) / g_unt < 1e-11
    m, ninv = c.total_metric(0, 0.0)
    assert abs(m - mtot) < 1e-7
    assert abs(m - mtot) / mtot < 1e-5  # golden printed with 6 digits
    assert ninv == n_inv


def bump_mesh(api, nele):
    """do_hex_4_structured (HeavyTestMeshSmoother.cpp:705-790): z <- z^2, then bump on z=0."""
    nn = nele + 1
    X = fixture_cube(nn)
    X[2] = X[2] * X[2]
    W = X.copy()
    scale = 2.0 * 4.0 * (5.0 / nele)
    ix, iy = X[0], X[1]
    bot = np.abs(X[2]) < 1e-5
    X[2] = np.where(bot, ix * (1.0 - ix) * iy * (1.0 - iy) * scale, X[2])
    c = api.create()
    b = c.add_structured_block((nn,) * 3)
    c.add_fixture_1_bcs(b)
    c.select_boundary_sgrid()
    c.commit()
    c.upload("coordinates", X)
    c.upload("coordinates_N", X)
    c.upload("coordinates_NM1", W)
    return c


def test_bump_mesh_untangle_golden():
    c = bump_mesh(oracle, 5)
    m, ninv = c.total_metric(0, 0.0)
    assert abs(m - 0.0179978) < 1e-6
    assert ninv == 25
