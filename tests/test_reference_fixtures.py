"""Tests on the reference's own mesh fixtures (tests/golden/*.npz, extracted by tests/golden/make_fixtures.py from
build-cmake/hex8-ref.e[.3.*] and test/unit_tests/exodus_files/tet_from_hex_fixture_0.e): real refiner node numbering
(scrambled ids), the reference's own 3-rank Nemesis decomposition, and a real tet4 mesh.

tet4 has no numeric golden anywhere in the reference (its tet_4 test asserts nothing), so the tet restatement is
cross-checked the way the reference's dormant check does it (ReferenceMeshSmootherConjugateGradientDef.hpp:1252-1335):
analytic gradient vs central finite differences of the metric, thresholds diff > 1e-8 && diff > 1e-3*comp."""
import os

import numpy as np
import pytest

from percept_b200 import partition as part
from tests.orc import oracle

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    d = np.load(os.path.join(G, name))
    X = np.ascontiguousarray(d["coords"])
    conn = np.ascontiguousarray(d["conn"].astype(np.int32))
    tol = 1e-9
    fixed = ((X < tol) | (X > 1 - tol)).any(axis=0).astype(np.uint8)
    return X, conn, fixed


def perturbed(W, fixed, amp, seed=5):
    rng = np.random.default_rng(seed)
    X = W + amp * (rng.random(W.shape) - 0.5) * (fixed == 0)[None, :]
    return np.ascontiguousarray(X)


def build(api, topo, X, W, conn, fixed, options=None):
    c = api.create()
    for k, v in (options or {}).items():
        c.set_option(k, v)
    c.set_unstructured(topo, X.shape[1], conn)
    c.set_fixed_mask(fixed)
    c.commit()
    c.upload("coordinates", X)
    c.upload("coordinates_N", X)
    c.upload("coordinates_NM1", W)
    return c


CASES = [("hex8_ref.npz", 8, 1.0 / 12), ("tet_from_hex_fixture_0.npz", 4, 1.0 / 8)]


@pytest.mark.parametrize("name,topo,h", CASES)
@pytest.mark.parametrize("stage,amp", [(1, 0.3), (0, 1.2)])
def test_oracle_analytic_gradient_matches_finite_differences(name, topo, h, stage, amp):
    W, conn, fixed = load(name)
    X = perturbed(W, fixed, amp * h)
    c = build(oracle, topo, X, W, conn, fixed)
    m0, ninv = c.total_metric(stage, 0.0)
    if stage == 1:
        assert ninv == 0
    c.get_gradient(stage)
    g = c.download("cg_g")
    free = np.nonzero(fixed == 0)[0]
    rng = np.random.default_rng(11)
    eps1 = np.sqrt(np.finfo(float).eps) * h  # sqrt_eps*edge_length_ave, Def.hpp:1276
    checked = 0
    for n in rng.choice(free, size=40, replace=False):
        for d in range(3):
            Xp = X.copy()
            Xp[d, n] += eps1
            c.upload("coordinates", Xp)
            mp, _ = c.total_metric(stage, 0.0)
            Xp[d, n] -= 2 * eps1
            c.upload("coordinates", Xp)
            mm, _ = c.total_metric(stage, 0.0)
            fd = (mp - mm) / (2 * eps1)
            ag = g[d, n]
            diff, comp = abs(ag - fd), (abs(ag) + abs(fd)) / 2
            assert not (comp > 1e-6 and diff > 1e-8 and diff > 1e-3 * comp), (name, stage, n, d, ag, fd)
            checked += 1
    assert checked == 120
    assert np.abs(g[:, fixed == 1]).max() == 0.0


def test_partition_reproduces_the_references_nemesis_decomposition():
    """hex8-ref.e.3.*: Percept's own 3-rank decomposition.  Feeding its element->rank assignment to our builder must
    reproduce each rank's node set and its node communication maps (n_comm_nids / n_comm_proc) exactly."""
    h = np.load(os.path.join(G, "hex8_ref.npz"))
    p = np.load(os.path.join(G, "hex8_ref_3ranks.npz"))
    conn = h["conn"]
    # node and element ids differ between the serial and the parallel refinement (new entities are numbered per
    # rank), so nodes are matched by position on the 13^3 lattice and elements by their node sets
    key = lambda X: np.rint(X * 12).astype(np.int64) @ np.array([1, 13, 169])
    pos2node = {int(k): i for i, k in enumerate(key(h["coords"].T))}
    assert len(pos2node) == 2197
    elem_of = {tuple(sorted(row)): e for e, row in enumerate(conn.tolist())}
    elem_rank = np.full(conn.shape[0], -1, np.int32)
    gls = []
    for r in range(3):
        gl = np.array([pos2node[int(k)] for k in key(p[f"r{r}_coords"].T)])
        gls.append(gl)
        for row in gl[p[f"r{r}_conn"]].tolist():
            e = elem_of[tuple(sorted(row))]
            assert elem_rank[e] == -1
            elem_rank[e] = r
    assert (elem_rank >= 0).all() and np.bincount(elem_rank).tolist() == [576, 576, 576]
    nodes_of = [np.unique(conn[elem_rank == r]) for r in range(3)]
    for r in range(3):
        gl = gls[r]
        assert np.array_equal(np.sort(gl), nodes_of[r])  # node set of the rank's owned elements
        for q in range(3):
            if q == r:
                continue
            ours = np.intersect1d(nodes_of[r], nodes_of[q])
            sel = p[f"r{r}_n_comm_proc"] == q
            theirs = np.unique(gl[p[f"r{r}_n_comm_nids"][sel]])
            assert np.array_equal(ours, theirs), (r, q)
    # and our ownership / aura construction on that decomposition is consistent
    owned = np.zeros(conn.max() + 1, int)
    for r in range(3):
        lm = part.local_mesh(conn, elem_rank, r, 3)
        owned[lm.node_gid[lm.node_owned == 1]] += 1
        assert lm.n_owned_elems == 576
    assert (owned == 1).all()


@pytest.mark.gpu
@pytest.mark.parametrize("name,topo,h", CASES)
@pytest.mark.parametrize("stage,amp", [(1, 0.3), (0, 1.2)])
def test_gpu_parity_on_reference_fixture_meshes(name, topo, h, stage, amp):
    import percept_b200 as pb
    W, conn, fixed = load(name)
    X = perturbed(W, fixed, amp * h)
    o = build(oracle, topo, X, W, conn, fixed, {"real_is_double": 1})
    o.get_gradient(stage)
    mo, no = o.total_metric(stage, 0.0)
    go = o.download("cg_g")
    for strict in (1, 0):
        g = build(pb.api, topo, X, W, conn, fixed, {"strict_fp": strict, "keep_element_metrics": 1, "cache_metric": 0})
        g.get_gradient(stage)
        mg, ng = g.total_metric(stage, 0.0)
        gg = g.download("cg_g")
        assert ng == no and abs(mg - mo) <= 1e-12 * abs(mo)
        if strict:
            assert np.array_equal(gg, go)
            assert np.array_equal(g.element_metrics()[0], o.element_metrics()[0])
        else:
            assert np.max(np.abs(gg - go)) <= 1e-12 * np.max(np.abs(go))
        rg, eg = g.node_elem_csr(topo)
        ro, eo = o.node_elem_csr(topo)
        assert np.array_equal(rg, ro) and np.array_equal(eg, eo)


@pytest.mark.gpu
@pytest.mark.parametrize("name,topo,h", CASES)
def test_gpu_smoothing_of_reference_fixture_meshes(name, topo, h):
    """Tangled start (untangle stage does real work), then ShapeB1: 8 iterations per stage, tol = 0."""
    import percept_b200 as pb
    W, conn, fixed = load(name)
    X = perturbed(W, fixed, 1.0 * h)
    o = build(oracle, topo, X, W, conn, fixed, {"real_is_double": 1})
    g = build(pb.api, topo, X, W, conn, fixed)
    assert g.count_invalid_elements() == o.count_invalid_elements() > 0
    try:
        so = o.run_algorithm(8, 0.0)
        oracle_failed = None
    except Exception as ex:  # the reference throws if untangling did not finish in the given iterations
        oracle_failed = str(ex)
    if oracle_failed:
        with pytest.raises(pb.SmootherError) as ei:
            g.run_algorithm(8, 0.0)
        assert "Invalid elements" in str(ei.value) and "Invalid elements" in oracle_failed
    else:
        sg = g.run_algorithm(8, 0.0)
        assert sg.num_invalid == so.num_invalid
        d = np.max(np.abs(g.download("coordinates") - o.download("coordinates")))
        assert d <= 1e-9


def test_skin_selection_finds_the_cube_faces():
    from percept_b200 import exodus
    for name, topo, h in CASES:
        W, conn, fixed = load(name)
        assert np.array_equal(exodus.skin_nodes(conn, topo), fixed)  # skin of the unit cube == its six faces


def test_exodus_reader_roundtrip(tmp_path):
    """Write a tiny classic-netCDF Exodus file the way the reference fixtures are laid out and read it back."""
    from scipy.io import netcdf_file
    from percept_b200 import exodus
    W, conn, _ = load("hex8_ref.npz")
    p = str(tmp_path / "m.e")
    nc = netcdf_file(p, "w", version=1)
    nc.createDimension("num_nodes", W.shape[1])
    nc.createDimension("num_el_in_blk1", conn.shape[0])
    nc.createDimension("num_nod_per_el1", 8)
    for i, n in enumerate(("coordx", "coordy", "coordz")):
        v = nc.createVariable(n, "d", ("num_nodes",))
        v[:] = W[i]
    c = nc.createVariable("connect1", "i", ("num_el_in_blk1", "num_nod_per_el1"))
    c[:] = conn + 1
    c.elem_type = b"HEX8"
    nc.close()
    m = exodus.read_exodus(p)
    assert np.array_equal(m["coords"], W) and m["blocks"][0][0] == 8 and np.array_equal(m["blocks"][0][1], conn)


@pytest.mark.gpu
def test_smooth_cli_on_reference_fixture(tmp_path):
    """mesh_adapt-style entry point on hex8-ref (written as Exodus, perturbed), skin fixed."""
    from scipy.io import netcdf_file
    from percept_b200 import smooth
    W, conn, fixed = load("hex8_ref.npz")
    X = perturbed(W, fixed, 0.4 / 12)

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
    write(str(tmp_path / "x.e"), X)
    write(str(tmp_path / "w.e"), W)
    smooth.main([str(tmp_path / "x.e"), str(tmp_path / "out.npz"), "--reference_mesh", str(tmp_path / "w.e"), "--smoother_niter", "200",
                 "--smoother_tol", "1e-4", "--log", str(tmp_path / "log.txt")])
    out = np.load(tmp_path / "out.npz")["coords"]
    assert np.array_equal(out[:, fixed == 1], W[:, fixed == 1])           # skin did not move
    assert np.max(np.abs(out - W)) < 0.05 * np.max(np.abs(X - W))          # interior went back to the reference mesh
    assert len(open(tmp_path / "log.txt").read().splitlines()) > 2
