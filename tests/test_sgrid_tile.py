"""The one-pass structured tile kernel (pms_sgrid_tile.cuh: staged node planes, in-CTA node gather, tile-seam scratch)
against the two-pass element/gather path and the CPU oracle: same nodal gradient BITWISE in the strict build (the
additions per node happen in the same (dk,dj,di) order), <= 1e-13 in the default build; metric, invalid counts and
per-element values alike.  Sizes are chosen so that tiles are clipped, seams exist in i and j, and k is cut in chunks."""
import numpy as np
import pytest

import percept_b200 as pb
from tests.orc import fixture_cube, oracle

pytestmark = pytest.mark.gpu


def block_coords(nn, amp, seed):
    W = fixture_cube(nn)
    rng = np.random.default_rng(seed)
    h = np.array([1.0 / (n - 1) for n in nn])[:, None]
    X = W + amp * h * (rng.random(W.shape) - 0.5)
    return X, W


def build(api, nn, X, W, fixed, options):
    c = api.create()
    for k, v in options.items():
        c.set_option(k, v)
    b = c.add_structured_block(nn)
    c.add_fixture_1_bcs(b)
    if fixed:
        c.select_boundary_sgrid()
    c.commit()
    for name, F in (("coordinates", X), ("coordinates_N", X), ("coordinates_NM1", W)):
        c.upload(name, F)
    return c


SIZES = [(70, 37, 21), (34, 18, 3), (2, 2, 2), (33, 17, 9), (65, 49, 40), (5, 100, 7)]


@pytest.mark.parametrize("nn", SIZES)
@pytest.mark.parametrize("stage", [0, 1])
@pytest.mark.parametrize("strict", [0, 1])
def test_fused_equals_two_pass_and_oracle(nn, stage, strict):
    # amp 1.2 inverts a good share of the cells (the untangle stage's input, and ShapeB1's dropped elements)
    X, W = block_coords(nn, 1.2 if stage == 0 else 0.5, seed=3)
    for fixed in (True, False):
        opts = {"strict_fp": strict, "cache_metric": 0, "keep_element_metrics": 1}
        f = build(pb.api, nn, X, W, fixed, dict(opts, sgrid_fused=1))
        t = build(pb.api, nn, X, W, fixed, dict(opts, sgrid_fused=0))
        o = build(oracle, nn, X, W, fixed, {"real_is_double": 1})
        n0 = f.launch_count()
        mf, nf = f.metric_and_gradient(stage)
        assert f.launch_count() - n0 <= 2  # tile kernel (+ seam sums), no element/gather pair
        mt, nt = t.metric_and_gradient(stage)
        mo, no = o.metric_and_gradient(stage)
        assert nf == nt == no
        assert abs(mf - mo) <= 1e-12 * max(abs(mo), 1e-300) and abs(mf - mt) <= 1e-13 * max(abs(mt), 1e-300)
        gf, gt, go = f.download("cg_g"), t.download("cg_g"), o.download("cg_g")
        scale = max(np.max(np.abs(go)), 1e-300)
        if strict:
            assert np.array_equal(gf, gt)
            assert np.array_equal(gf, go)
            assert np.array_equal(f.element_metrics()[0], o.element_metrics()[0])
        else:
            assert np.max(np.abs(gf - gt)) <= 1e-13 * scale
            assert np.max(np.abs(gf - go)) <= 1e-12 * scale
        assert np.array_equal(f.element_metrics()[1], o.element_metrics()[1])
        # deterministic: a second evaluation is bitwise the same
        f.metric_and_gradient(stage)
        assert np.array_equal(f.download("cg_g"), gf)
        for c in (f, t, o):
            c.close()


def test_fused_at_128_cubed_matches_oracle():
    """BASELINE config 2 size (128^3 cells): 4 x 8 tiles, several k chunks, 32-bit/64-bit index paths."""
    nn = (129, 129, 129)
    X, W = block_coords(nn, 0.3, seed=11)
    f = build(pb.api, nn, X, W, True, {"cache_metric": 0})
    o = build(oracle, nn, X, W, True, {"real_is_double": 1})
    mf, nf = f.metric_and_gradient(1)
    mo, no = o.metric_and_gradient(1)
    assert nf == no
    assert abs(mf - mo) <= 1e-12 * abs(mo)
    gf, go = f.download("cg_g"), o.download("cg_g")
    err = np.max(np.abs(gf - go)) / np.max(np.abs(go))
    assert err <= 1e-12, err
    s = build(pb.api, nn, X, W, True, {"cache_metric": 0, "strict_fp": 1})
    s.metric_and_gradient(1)
    assert np.array_equal(s.download("cg_g"), go)
