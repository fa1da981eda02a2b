"""The one-kernel unstructured gradient (pms_elem_ws.cuh: persistent element loop + lagged node gather through a ring
of per-element contributions) against the two-pass element/gather path: BITWISE the same nodal gradient (same CSR walk
per node), same metric and invalid count; and against the CPU oracle.  Meshes must be large enough for the plan to
exist (4 slabs of 32 x 8 x #SM elements), so these run at 64^3 hexes / 40^3 x 6 tets."""
import numpy as np
import pytest

import percept_b200 as pb
from tests import meshes
from tests.orc import oracle

pytestmark = pytest.mark.gpu


def build(api, kind, nele, X, W, opts, perm=None):
    c = api.create()
    for k, v in opts.items():
        c.set_option(k, v)
    nn = nele + 1
    conn = meshes.hex8_conn(nele) if kind == "hex8" else meshes.tet4_conn(nele)
    if perm is not None:
        conn = np.ascontiguousarray(conn[perm])
    c.set_unstructured(8 if kind == "hex8" else 4, nn ** 3, conn)
    c.set_fixed_mask(meshes.boundary_mask(nn))
    c.commit()
    for f, v in (("coordinates", X), ("coordinates_N", X), ("coordinates_NM1", W)):
        c.upload(f, v)
    return c


@pytest.mark.parametrize("kind,nele", [("hex8", 64), ("tet4", 40)])
@pytest.mark.parametrize("stage", [0, 1])
def test_one_kernel_gradient_equals_two_pass_and_oracle(kind, nele, stage):
    X, W = meshes.perturbed_coords(nele, 1.0 if stage == 0 else 0.25)
    a = build(pb.api, kind, nele, X, W, {"cache_metric": 0, "elem_ws": 1})
    b = build(pb.api, kind, nele, X, W, {"cache_metric": 0, "elem_ws": 0})
    n0 = a.launch_count()
    ma, na = a.metric_and_gradient(stage)
    assert a.launch_count() - n0 == 1  # one kernel: no element pass + gather pair
    mb, nb = b.metric_and_gradient(stage)
    assert na == nb and abs(ma - mb) <= 1e-13 * abs(mb)
    ga, gb = a.download("cg_g"), b.download("cg_g")
    assert np.array_equal(ga, gb)
    assert np.array_equal(a.download("cg_r"), b.download("cg_r"))
    for _ in range(3):  # deterministic across launches (the counters are reset, the ring is re-used)
        a.metric_and_gradient(stage)
        assert np.array_equal(a.download("cg_g"), ga)
    o = build(oracle, kind, nele, X, W, {"real_is_double": 1})
    mo, no = o.metric_and_gradient(stage)
    go = o.download("cg_g")
    assert na == no and abs(ma - mo) <= 1e-12 * abs(mo)
    assert np.max(np.abs(ga - go)) <= 1e-12 * np.max(np.abs(go))
    if stage == 1:  # a CG run through the driver uses it too: same iterates as with the two-pass gradient
        sa, sb = a.run_algorithm(3, 0.0), b.run_algorithm(3, 0.0)
        assert np.array_equal(a.download("coordinates"), b.download("coordinates"))
        assert sa.global_metric == sb.global_metric


def test_wildly_numbered_mesh_keeps_the_two_pass_path():
    """Elements in random order: the elements around a node span the whole index range, the ring would be as large as the
    full scratch, so the plan is refused and the gradient still comes out right."""
    nele = 56
    X, W = meshes.perturbed_coords(nele, 0.25)
    perm = np.random.default_rng(5).permutation(nele ** 3)
    a = build(pb.api, "hex8", nele, X, W, {"cache_metric": 0, "elem_ws": 1}, perm)
    b = build(pb.api, "hex8", nele, X, W, {"cache_metric": 0, "elem_ws": 0}, perm)
    n0 = a.launch_count()
    ma, na = a.metric_and_gradient(1)
    assert a.launch_count() - n0 == 2
    mb, nb = b.metric_and_gradient(1)
    assert (ma, na) == (mb, nb)
    assert np.array_equal(a.download("cg_g"), b.download("cg_g"))
