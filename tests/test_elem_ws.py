"""The one-kernel unstructured hex8 gradient (pms_elem_ws.cuh: persistent element loop with the node gather riding on it
through cp.async, nodes listed by completion iteration, cross-CTA progress words) against the two-pass element/gather
path: BITWISE the same nodal gradient (same CSR walk per node), same metric and invalid count; and against the CPU
oracle.  tet4 keeps the two-pass path."""
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


@pytest.mark.parametrize("kind,nele", [("hex8", 64), ("hex8", 37), ("tet4", 24)])
@pytest.mark.parametrize("stage", [0, 1])
def test_one_kernel_gradient_equals_two_pass_and_oracle(kind, nele, stage):
    X, W = meshes.perturbed_coords(nele, 1.0 if stage == 0 else 0.25)
    a = build(pb.api, kind, nele, X, W, {"cache_metric": 0, "elem_ws": 1})
    b = build(pb.api, kind, nele, X, W, {"cache_metric": 0, "elem_ws": 0})
    n0 = a.launch_count()
    ma, na = a.metric_and_gradient(stage)
    # one kernel: no element pass + gather pair (tet4: the two passes + the once-per-run reference cache of its elements)
    assert a.launch_count() - n0 == (1 if kind == "hex8" else 3)
    mb, nb = b.metric_and_gradient(stage)
    assert na == nb and abs(ma - mb) <= 1e-13 * abs(mb)
    ga, gb = a.download("cg_g"), b.download("cg_g")
    assert np.array_equal(ga, gb)
    assert np.array_equal(a.download("cg_r"), b.download("cg_r"))
    for _ in range(3):  # deterministic across launches (the progress words are reset, the scratch is re-used)
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


def test_wildly_numbered_mesh():
    """Elements in random order: every node completes only in the last iterations, so the gather trails the element
    loop instead of riding on it (or the plan is refused); the gradient still comes out bit for bit."""
    nele = 56
    X, W = meshes.perturbed_coords(nele, 0.25)
    perm = np.random.default_rng(5).permutation(nele ** 3)
    a = build(pb.api, "hex8", nele, X, W, {"cache_metric": 0, "elem_ws": 1}, perm)
    b = build(pb.api, "hex8", nele, X, W, {"cache_metric": 0, "elem_ws": 0}, perm)
    ma, na = a.metric_and_gradient(1)
    mb, nb = b.metric_and_gradient(1)
    assert na == nb and abs(ma - mb) <= 1e-13 * abs(mb)
    assert np.array_equal(a.download("cg_g"), b.download("cg_g"))


def test_nodes_with_more_than_eight_elements_and_isolated_nodes():
    """Irregular valence: two lattice nodes merged into one (16 adjacent hexes: beyond the 8-entry table, handled by the
    list-driven CSR gather) leaves the other one without elements.  Untangle stage (the merged mesh is tangled)."""
    nele = 40
    nn = nele + 1
    X, W = meshes.perturbed_coords(nele, 0.25)
    conn = meshes.hex8_conn(nele).copy()
    keep = 10 + nn * (11 + nn * 12)
    for gone in (20 + nn * (21 + nn * 22), 30 + nn * (5 + nn * 31)):
        conn[conn == gone] = keep
    res = []
    for ws in (1, 0):
        c = pb.api.create()
        c.set_option("cache_metric", 0)
        c.set_option("elem_ws", ws)
        c.set_unstructured(8, nn ** 3, conn)
        c.set_fixed_mask(meshes.boundary_mask(nn))
        c.commit()
        for f, v in (("coordinates", X), ("coordinates_N", X), ("coordinates_NM1", W)):
            c.upload(f, v)
        n0 = c.launch_count()
        m, ninv = c.metric_and_gradient(0)
        res.append((m, ninv, c.download("cg_g"), c.launch_count() - n0))
    assert res[0][3] == 2 and res[1][3] == 2  # one-kernel path: main kernel + overflow list; two-pass: element + gather
    assert res[0][1] == res[1][1] and abs(res[0][0] - res[1][0]) <= 1e-13 * abs(res[1][0])
    assert np.array_equal(res[0][2], res[1][2])
