"""Surface smoothing hooks: the `smooth_surfaces` branch of the CG smoother.

Reference: get_fixed_flag (mesh/mod/smoother/MeshSmoother.cpp:271-344), GenericAlgorithm_get_gradient_2 +
project_delta_to_tangent_plane (ReferenceMeshSmootherConjugateGradientDef.hpp:1401-1447, MeshSmoother.cpp:437-460),
snap_nodes and the post-snap validity check (Def.hpp:236-316, 1078-1089), get_surface_normals (Def.hpp:596-706).
The reference evaluates CAD (OpenNURBS / PGEOM, third party, absent; its geometry tests need .3dm files that are not in
the repository), so parity is "unpinned by goldens": the oracle restates the control flow with the same analytic
evaluators the product uses, and both are pinned by the properties the reference itself checks
(check_project_all_delta_to_tangent_plane: |g.n| <= 1e-6 |g|) plus on-surface / idempotence / fixed-node invariants.

Test body: a box whose top face is a cylinder patch; corners are vertices (dof 0), box edges curves (dof 1), face
interiors surfaces (dof 2: five planes + the cylinder), the rest volume (dof 3).
"""
import numpy as np
import pytest

from tests import meshes
from tests.orc import fixture_cube, oracle

AXIS_P, AXIS_U, RAD = (0.5, 0.0, -2.0), (0.0, 1.0, 0.0), 3.0


def ztop(x):
    return AXIS_P[2] + np.sqrt(RAD ** 2 - (x - AXIS_P[0]) ** 2)


def curved_box(nele, seed=5, amp=0.2):
    """reference coords W, perturbed coords X (surface nodes perturbed tangentially-ish, NOT yet snapped), dof, surf"""
    nn = nele + 1
    L = fixture_cube(nn)  # lattice in [0,1]^3
    W = np.stack([L[0], L[1], L[2] * ztop(L[0])])
    ii = np.rint(L * nele).astype(int)
    on = [(ii[d] == 0) | (ii[d] == nele) for d in range(3)]
    nb = on[0].astype(int) + on[1].astype(int) + on[2].astype(int)
    dof = np.where(nb == 3, 0, np.where(nb == 2, 1, np.where(nb == 1, 2, 3))).astype(np.int8)
    # surfaces: 0 x=0, 1 x=1, 2 y=0, 3 y=1, 4 z=0 (planes), 5 top cylinder
    surf = np.full(nn ** 3, -1, dtype=np.int32)
    face = nb == 1
    surf[face & (ii[0] == 0)] = 0
    surf[face & (ii[0] == nele)] = 1
    surf[face & (ii[1] == 0)] = 2
    surf[face & (ii[1] == nele)] = 3
    surf[face & (ii[2] == 0)] = 4
    surf[face & (ii[2] == nele)] = 5
    rng = np.random.default_rng(seed)
    X = W + amp / nele * (rng.random(W.shape) - 0.5) * (dof >= 2)[None, :]
    return np.ascontiguousarray(W), np.ascontiguousarray(X), dof, surf


def add_surfaces(c):
    ids = [c.add_analytic_surface(0, (0, 0, 0, -1, 0, 0)), c.add_analytic_surface(0, (1, 0, 0, 1, 0, 0)),
           c.add_analytic_surface(0, (0, 0, 0, 0, -1, 0)), c.add_analytic_surface(0, (0, 1, 0, 0, 2, 0)),  # normal gets normalised
           c.add_analytic_surface(0, (0, 0, 0, 0, 0, -1)), c.add_analytic_surface(2, AXIS_P + AXIS_U + (RAD,))]
    assert ids == list(range(6))


def build(api, nele, smooth, options=None, snap=True):
    W, X, dof, surf = curved_box(nele)
    c = api.create()
    for k, v in (options or {}).items():
        c.set_option(k, v)
    c.set_unstructured(8, W.shape[1], meshes.hex8_conn(nele))
    c.commit()
    add_surfaces(c)
    c.set_option("smooth_surfaces", smooth)
    c.set_node_classification(dof, surf)
    for f, v in (("coordinates", X), ("coordinates_N", X), ("coordinates_NM1", W)):
        c.upload(f, v)
    if snap and smooth:
        c.snap_nodes()
        c.copy_field("coordinates_N", "coordinates")
    return c, W, X, dof, surf


def surface_residual(x, surf):
    """distance of every surface node from its surface"""
    r = np.zeros(x.shape[1])
    for s, (d, v) in enumerate(((0, 0.0), (0, 1.0), (1, 0.0), (1, 1.0), (2, 0.0))):
        m = surf == s
        r[m] = np.abs(x[d, m] - v)
    m = surf == 5
    r[m] = np.abs(np.hypot(x[0, m] - AXIS_P[0], x[2, m] - AXIS_P[2]) - RAD)
    return r


def check_invariants(c, W, dof, surf, smooth):
    x = c.download("coordinates")
    assert np.array_equal(x[:, dof < 2], W[:, dof < 2])  # vertices and curves never move
    if smooth:
        assert surface_residual(x, surf).max() <= 4e-15  # surface nodes stay on their surface
    else:
        x0 = c.download("coordinates_N")
        assert np.array_equal(x[:, dof == 2], x0[:, dof == 2])


def test_oracle_classification_projection_and_snap():
    c, W, X, dof, surf = build(oracle, 6, 1, {"real_is_double": 1}, snap=False)
    assert surface_residual(X, surf).max() > 1e-3  # the perturbation left the surfaces
    d = c.snap_nodes()
    x = c.download("coordinates")
    assert 1e-3 < d < 0.05 and surface_residual(x, surf).max() <= 4e-15
    assert np.array_equal(x[:, dof != 2], X[:, dof != 2])  # only surface nodes are snapped
    assert c.snap_nodes() <= 1e-15  # idempotent
    nrm = c.get_surface_normals()
    assert np.allclose(np.linalg.norm(nrm[:, dof == 2], axis=0), 1.0, atol=1e-15) and not nrm[:, dof != 2].any()
    assert np.allclose(nrm[:, surf == 3].T, (0.0, 1.0, 0.0))
    for stage in (0, 1):
        c.get_gradient(stage)
        g, r = c.download("cg_g"), c.download("cg_r")
        m = dof == 2
        gn = (g[:, m] * nrm[:, m]).sum(axis=0)
        # the reference's own check (check_project_all_delta_to_tangent_plane): |g.n| <= 1e-6 |g|
        assert np.all(np.abs(gn) <= 1e-12 * np.maximum(np.linalg.norm(g[:, m], axis=0), 1e-300))
        rn = (r[:, m] * nrm[:, m]).sum(axis=0)
        assert np.allclose(g[:, m], r[:, m] - rn * nrm[:, m], rtol=0, atol=1e-12 * np.abs(r).max())  # cg_r keeps the raw gradient
        assert np.array_equal(g[:, dof == 3], r[:, dof == 3]) and not g[:, dof < 2].any()
        if stage == 1:  # (the untangle gradient of this valid mesh is identically zero)
            assert np.abs(rn).max() > 1e-6 * np.abs(r).max() > 0  # the projection did remove something


@pytest.mark.parametrize("smooth", [0, 1])
def test_oracle_smoothing_with_and_without_surface_motion(smooth):
    c, W, X, dof, surf = build(oracle, 6, smooth, {"real_is_double": 1})
    m0, _ = c.total_metric(1, 0.0)
    st = c.run_algorithm(12, 1e-6)
    assert st.num_invalid == 0 and st.global_metric < m0
    check_invariants(c, W, dof, surf, smooth)
    if smooth:
        x, x0 = c.download("coordinates"), c.download("coordinates_N")
        assert np.abs(x[:, dof == 2] - x0[:, dof == 2]).max() > 1e-4  # surface nodes did slide


def test_oracle_sliding_surfaces_reach_a_lower_metric_than_fixed_surfaces():
    out = {}
    for smooth in (0, 1):
        c, W, X, dof, surf = build(oracle, 5, smooth, {"real_is_double": 1})
        if not smooth:  # same starting mesh as the smooth run: snap by hand
            s, *_ = build(oracle, 5, 1, {"real_is_double": 1})
            x0 = s.download("coordinates")
            for f in ("coordinates", "coordinates_N"):
                c.upload(f, x0)
        out[smooth] = c.run_algorithm(25, 1e-7).global_metric
    assert out[1] < out[0]


def test_structured_grids_are_not_implemented_like_the_reference():
    from percept_b200.capi import SmootherError
    c = meshes.build(oracle, "sgrid", 3, *meshes.perturbed_coords(3, 0.2))
    with pytest.raises(SmootherError):
        c.set_node_classification(np.zeros(64, np.int8), np.zeros(64, np.int32))


# ------------------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("strict", [0, 1])
def test_gpu_projection_snap_normals_match_oracle(strict):
    import percept_b200 as pb
    g, W, X, dof, surf = build(pb.api, 7, 1, {"strict_fp": strict}, snap=False)
    o, *_ = build(oracle, 7, 1, {"real_is_double": 1}, snap=False)
    dg, do = g.snap_nodes(), o.snap_nodes()
    xg, xo = g.download("coordinates"), o.download("coordinates")
    if strict:
        assert dg == do and np.array_equal(xg, xo)
    else:
        assert abs(dg - do) <= 2e-15 and np.abs(xg - xo).max() <= 2e-15  # FMA contraction: a few ulp
    assert surface_residual(xg, surf).max() <= 4e-15
    g.upload("coordinates", xo)
    ng, no = g.get_surface_normals(), o.get_surface_normals()
    assert np.abs(ng - no).max() <= (0 if strict else 1e-15)
    for stage in (0, 1):
        mg, ig = g.metric_and_gradient(stage)
        mo, io = o.metric_and_gradient(stage)
        assert ig == io and abs(mg - mo) <= 1e-12 * max(abs(mo), 1e-300)
        gg, go = g.download("cg_g"), o.download("cg_g")
        assert np.abs(gg - go).max() <= 1e-12 * np.abs(go).max()
        assert np.abs(g.download("cg_r") - o.download("cg_r")).max() <= 1e-12 * np.abs(go).max()
        m = dof == 2
        gn = (gg[:, m] * no[:, m]).sum(axis=0)
        assert np.all(np.abs(gn) <= 1e-12 * np.maximum(np.linalg.norm(gg[:, m], axis=0), 1e-300))


@pytest.mark.gpu
@pytest.mark.parametrize("smooth", [0, 1])
@pytest.mark.parametrize("variant", [2, 5])
def test_gpu_surface_smoothing_run_matches_oracle(smooth, variant):
    import percept_b200 as pb
    opts = {"variant_grad_corners": variant}
    if variant >= 4:
        opts["keep_element_metrics"] = 0
    g, W, X, dof, surf = build(pb.api, 8, smooth, opts)
    o, *_ = build(oracle, 8, smooth, {"real_is_double": 1})
    sg, so = g.run_algorithm(10, 0.0), o.run_algorithm(10, 0.0)
    assert (sg.stage, sg.iter, sg.num_invalid) == (so.stage, so.iter, so.num_invalid)
    assert abs(sg.global_metric - so.global_metric) <= 1e-9 * abs(so.global_metric)
    assert np.abs(g.download("coordinates") - o.download("coordinates")).max() <= 1e-9
    check_invariants(g, W, dof, surf, smooth)
    g.set_option("variant_grad_corners", 2)


@pytest.mark.gpu
def test_gpu_surface_hook_errors():
    import percept_b200 as pb
    from percept_b200.capi import SmootherError
    c = meshes.build(pb.api, "sgrid", 3, *meshes.perturbed_coords(3, 0.2))
    with pytest.raises(SmootherError, match="not impl"):
        c.set_node_classification(np.zeros(64, np.int8), np.zeros(64, np.int32))
    h = pb.api.create()
    h.set_unstructured(8, 27, meshes.hex8_conn(2))
    h.commit()
    with pytest.raises(SmootherError, match="evaluator"):
        h.set_node_classification(np.full(27, 2, np.int8), np.zeros(27, np.int32))  # no surface registered
    with pytest.raises(SmootherError, match="kind"):
        h.add_analytic_surface(7, (0, 0, 0))
    with pytest.raises(SmootherError, match="direction"):
        h.add_analytic_surface(0, (0, 0, 0, 0, 0, 0))


def test_host_classifier_reproduces_the_box_classification():
    """percept_b200.geometry.classify_nodes (set-up code): skin nodes on 1 / 2 / >=3 analytic surfaces are
    surface / curve / vertex nodes - the classification the curved-box tests construct by hand."""
    from percept_b200 import exodus, geometry
    nele = 6
    W, X, dof, surf = curved_box(nele)
    surfaces = [{"kind": "plane", "point": [0, 0, 0], "normal": [-1, 0, 0]}, {"kind": "plane", "point": [1, 0, 0], "normal": [1, 0, 0]},
                {"kind": "plane", "point": [0, 0, 0], "normal": [0, -1, 0]}, {"kind": "plane", "point": [0, 1, 0], "normal": [0, 2, 0]},
                {"kind": "plane", "point": [0, 0, 0], "normal": [0, 0, -1]},
                {"kind": "cylinder", "point": list(AXIS_P), "axis": list(AXIS_U), "radius": RAD}]
    skin = exodus.skin_nodes(meshes.hex8_conn(nele), 8)
    d2, s2 = geometry.classify_nodes(W, skin, surfaces, 1e-12)
    assert np.array_equal(d2, dof) and np.array_equal(s2, surf)
    kind, p = geometry.surface_params(surfaces[5])
    assert kind == 2 and np.allclose(p, AXIS_P + AXIS_U + (RAD,))
    # a skin node on no listed surface is held fixed
    d3, _ = geometry.classify_nodes(W, skin, surfaces[:5], 1e-12)
    top_interior = (surf == 5)
    assert np.all(d3[top_interior] == 0)


@pytest.mark.gpu
def test_smooth_cli_with_analytic_surfaces(tmp_path):
    """`python -m percept_b200.smooth --smooth_surfaces=1 --surfaces geo.json`: the curved box written as Exodus; the
    classifier finds faces / edges / corners from the geometry, face nodes slide and stay on their surface."""
    import json
    from scipy.io import netcdf_file
    from percept_b200 import smooth
    nele = 6
    W, X, dof, surf = curved_box(nele)
    conn = meshes.hex8_conn(nele)

    def write(path, C):
        nc = netcdf_file(path, "w", version=1)
        nc.createDimension("num_nodes", C.shape[1])
        nc.createDimension("num_el_in_blk1", conn.shape[0])
        nc.createDimension("num_nod_per_el1", 8)
        for i, n in enumerate(("coordx", "coordy", "coordz")):
            v = nc.createVariable(n, "d", ("num_nodes",))
            v[:] = C[i]
        c = nc.createVariable("connect1", "i", ("num_el_in_blk1", "num_nod_per_el1"))
        c[:] = conn + 1
        nc.close()
    # start from a mesh whose surface nodes already lie on the geometry (perturbed tangentially, then snapped by the oracle)
    o, *_ = build(oracle, nele, 1, {"real_is_double": 1})
    X0 = o.download("coordinates")
    write(str(tmp_path / "x.e"), X0)
    write(str(tmp_path / "w.e"), W)
    geo = [{"kind": "plane", "point": [0, 0, 0], "normal": [-1, 0, 0]}, {"kind": "plane", "point": [1, 0, 0], "normal": [1, 0, 0]},
           {"kind": "plane", "point": [0, 0, 0], "normal": [0, -1, 0]}, {"kind": "plane", "point": [0, 1, 0], "normal": [0, 1, 0]},
           {"kind": "plane", "point": [0, 0, 0], "normal": [0, 0, -1]},
           {"kind": "cylinder", "point": list(AXIS_P), "axis": list(AXIS_U), "radius": RAD}]
    json.dump(geo, open(tmp_path / "geo.json", "w"))
    smooth.main([str(tmp_path / "x.e"), str(tmp_path / "out.npz"), "--reference_mesh", str(tmp_path / "w.e"), "--smoother_niter", "12",
                 "--smoother_tol", "1e-6", "--smooth_surfaces", "1", "--surfaces", str(tmp_path / "geo.json"), "--surface_tol", "1e-12",
                 "--log", str(tmp_path / "log.txt")])
    out = np.load(tmp_path / "out.npz")["coords"]
    assert np.array_equal(out[:, dof < 2], W[:, dof < 2])
    assert surface_residual(out, surf).max() <= 4e-15
    assert np.abs(out[:, dof == 2] - X0[:, dof == 2]).max() > 1e-4
    so = o.run_algorithm(12, 1e-6)
    assert np.abs(out - o.download("coordinates")).max() <= 1e-9
