"""StructuredGridRefiner: 2x refinement of a block-structured grid feeding the smoother.

Reference: structured/StructuredGridRefiner.cpp:19-170 (sizes, BC / zone-connectivity ranges r -> 2r-1) and
structured/StructuredGridRefinerDef.hpp:134-222 (vertex / edge / face / centroid insertion).  The reference's tests hold
no numeric golden for the refiner (regr_PerceptCGNS.test6_sgrid_refine only dumps files; the mbvsb tests refine and then
compare multi-block with single-block smoothing), so the restatement is pinned through properties: the refined
fixture_1 cube is the finer fixture_1 cube, affine fields are reproduced exactly, multi-block == single block, and the
refined boundary conditions select exactly the boundary of the refined block.  The CUDA kernel is output-centric (one
thread per fine node), the oracle input-centric like the reference; they must agree bit for bit.
"""
import numpy as np
import pytest

from tests import meshes
from tests.orc import fixture_cube, oracle
from tests.test_multiblock import build_multiblock, gather_single


def single_block(api, nn, X, options=None):
    c = api.create()
    for k, v in (options or {}).items():
        c.set_option(k, v)
    b = c.add_structured_block(nn)
    c.add_fixture_1_bcs(b)
    c.select_boundary_sgrid()
    c.commit()
    for f in ("coordinates", "coordinates_N", "coordinates_NM1"):
        c.upload(f, X)
    return c


def refine(api, coarse, fields=("coordinates",), options=None):
    fine = api.create()
    for k, v in (options or {}).items():
        fine.set_option(k, v)
    ncell = coarse.refine_topology(fine)
    fine.select_boundary_sgrid()
    fine.commit()
    for f in fields:
        coarse.refine_field(f, fine, f)
    return fine, ncell


def moved_nodes(c):
    """mask of nodes update_node_positions moves (= not fixed)"""
    n = c.num_nodes_local
    before = c.download("coordinates").copy()
    c.upload("cg_s", np.ones((3, n)))
    c.upload("cg_edge_length", np.ones((1, n)))
    c.update_node_positions(0.125)
    after = c.download("coordinates")
    c.upload("coordinates", before)
    return (after != before).any(axis=0)


def test_oracle_refined_fixture_is_the_finer_fixture():
    nn = (5, 4, 7)
    coarse = single_block(oracle, nn, fixture_cube(nn))
    fine, ncell = refine(oracle, coarse)
    fn = tuple(2 * v - 1 for v in nn)
    assert fine.block_sizes == [fn] and ncell == (fn[0] - 1) * (fn[1] - 1) * (fn[2] - 1)
    assert fine.num_nodes_local == fn[0] * fn[1] * fn[2]
    assert np.max(np.abs(fine.download("coordinates") - fixture_cube(fn))) <= 2.3e-16
    # refined boundary conditions (r -> 2r-1) select exactly the 6 faces of the refined block
    i = np.arange(fn[0])[None, None, :]
    j = np.arange(fn[1])[None, :, None]
    k = np.arange(fn[2])[:, None, None]
    interior = ((i > 0) & (i < fn[0] - 1) & (j > 0) & (j < fn[1] - 1) & (k > 0) & (k < fn[2] - 1)).reshape(-1)
    assert np.array_equal(moved_nodes(fine), interior)


def test_oracle_refinement_reproduces_affine_fields_exactly():
    nn = (4, 6, 3)
    W = fixture_cube(nn, width=(3.0, 5.0, 2.0))  # lattice coordinates 0..n-1: dyadic, so all averages are exact
    X = np.stack([2 * W[0] - 3 * W[1] + 0.5 * W[2] + 1, W[1] + 4 * W[2], -W[0] + 0.25 * W[2] - 7])
    coarse = single_block(oracle, nn, np.ascontiguousarray(X))
    fine, _ = refine(oracle, coarse)
    fn = tuple(2 * v - 1 for v in nn)
    Wf = fixture_cube(fn, width=(3.0, 5.0, 2.0))
    Xf = np.stack([2 * Wf[0] - 3 * Wf[1] + 0.5 * Wf[2] + 1, Wf[1] + 4 * Wf[2], -Wf[0] + 0.25 * Wf[2] - 7])
    assert np.array_equal(fine.download("coordinates"), Xf)


def test_oracle_stencils_against_direct_numpy_averages():
    """Every fine node is the mean of the 1/2/4/8 coarse nodes of its cell entity (to rounding; order pinned elsewhere)."""
    nn = (4, 5, 3)
    rng = np.random.default_rng(3)
    X = rng.random((3, nn[0] * nn[1] * nn[2]))
    coarse = single_block(oracle, nn, X)
    fine, _ = refine(oracle, coarse)
    fn = tuple(2 * v - 1 for v in nn)
    F = fine.download("coordinates").reshape(3, fn[2], fn[1], fn[0])
    C = X.reshape(3, nn[2], nn[1], nn[0])
    for K in range(fn[2]):
        for J in range(fn[1]):
            for I in range(fn[0]):
                sl = tuple(slice(v >> 1, (v >> 1) + (v & 1) + 1) for v in (K, J, I))
                ref = C[(slice(None),) + sl].reshape(3, -1).mean(axis=1)
                assert np.max(np.abs(F[:, K, J, I] - ref)) <= 4e-16
    assert np.array_equal(F[:, ::2, ::2, ::2], C)  # vertices are copies


def test_oracle_refined_multiblock_equals_refined_single_block():
    """Zone connectivity ranges map r -> 2r-1 on both sides, so the refined two-block grid is still consistently
    connected (the property the reference's mbvsb tests rely on after do_refine).  Two blocks: the reference's
    sum + propagate interface passes, which the oracle restates literally, are complete only then."""
    nele = 4
    X, W = meshes.perturbed_coords(nele, 0.3)
    multi, maps = build_multiblock(oracle, nele, (2, 1, 1), X, W, {"real_is_double": 1})
    single = meshes.build(oracle, "sgrid", nele, X, W, True, {"real_is_double": 1})
    fm, nm = refine(oracle, multi, ("coordinates", "coordinates_NM1"))
    fs, ns = refine(oracle, single, ("coordinates", "coordinates_NM1"))
    assert nm == ns == (2 * nele) ** 3
    nf = 2 * nele + 1
    assert fm.get_number_nodes() == fs.get_number_nodes() == nf ** 3  # interface duplicates counted once
    # fine maps: block (x0..x1) -> (2x0..2x1)
    fmaps = []
    for m in maps:
        k, r = np.divmod(m, (nele + 1) ** 2)
        j, i = np.divmod(r, nele + 1)
        I = np.arange(2 * i.min(), 2 * i.max() + 1)[None, None, :]
        J = np.arange(2 * j.min(), 2 * j.max() + 1)[None, :, None]
        K = np.arange(2 * k.min(), 2 * k.max() + 1)[:, None, None]
        fmaps.append((I + nf * (J + nf * K)).reshape(-1))
    xs = fs.download("coordinates")
    assert np.array_equal(gather_single(fm, fmaps, "coordinates", 3, nf ** 3), xs)
    ms, _ = fs.metric_and_gradient(1)
    mm, _ = fm.metric_and_gradient(1)
    assert abs(mm - ms) <= 1e-10 * abs(ms)
    gs = fs.download("cg_g")
    assert np.max(np.abs(gather_single(fm, fmaps, "cg_g", 3, nf ** 3) - gs)) <= 1e-12 * np.max(np.abs(gs))


def test_refiner_argument_errors_oracle():
    from percept_b200.capi import SmootherError
    coarse = single_block(oracle, (3, 3, 3), fixture_cube(3))
    other = single_block(oracle, (3, 3, 3), fixture_cube(3))
    with pytest.raises(SmootherError):
        coarse.refine_topology(other)  # output ctx must be empty
    with pytest.raises(SmootherError):
        coarse.refine_field("coordinates", other, "coordinates")  # not the 2x refinement


# ------------------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("strict", [0, 1])
@pytest.mark.parametrize("nn", [(5, 4, 7), (2, 2, 2), (33, 17, 9)])
def test_gpu_refinement_is_bit_identical_to_the_oracle(nn, strict):
    import percept_b200 as pb
    rng = np.random.default_rng(7)
    X = np.ascontiguousarray(fixture_cube(nn) + 0.3 / max(nn) * (rng.random((3, nn[0] * nn[1] * nn[2])) - 0.5))
    g = single_block(pb.api, nn, X, {"strict_fp": strict})
    o = single_block(oracle, nn, X)
    fg, ng = refine(pb.api, g, ("coordinates", "coordinates_NM1"), {"strict_fp": strict})
    fo, no = refine(oracle, o, ("coordinates", "coordinates_NM1"))
    assert ng == no and fg.block_sizes == fo.block_sizes
    assert np.array_equal(fg.download("coordinates"), fo.download("coordinates"))
    assert np.array_equal(fg.download("coordinates_NM1"), fo.download("coordinates_NM1"))
    assert np.array_equal(moved_nodes(fg), moved_nodes(fo))  # refined boundary conditions -> same fixed nodes


@pytest.mark.gpu
def test_gpu_refine_twice_then_smooth_matches_oracle():
    """The mbvsb workflow (SGridSmootherMultiBlockvsSingleBlock.cpp:150-175): fixture -> do_refine x2 -> perturb ->
    smooth; multi-block and single-block agree and both agree with the oracle."""
    import percept_b200 as pb
    nele = 3
    W0 = fixture_cube(nele + 1)

    def workflow(api, procs, opts):
        if procs is None:
            c = meshes.build(api, "sgrid", nele, W0, W0, True, opts)
            maps = None
        else:
            c, maps = build_multiblock(api, nele, procs, W0, W0, opts)
        for _ in range(2):
            c, _n = refine(api, c)
        return c, maps

    ne = 4 * nele
    Xp, Wp = meshes.perturbed_coords(ne, 0.25)
    runs = {}
    for name, api, procs, opts in (("g1", pb.api, None, {}), ("o1", oracle, None, {"real_is_double": 1}),
                                   ("g2", pb.api, (2, 1, 1), {})):
        c, maps = workflow(api, procs, opts)
        ref = c.download("coordinates") if procs is None else None
        if procs is None:
            assert np.max(np.abs(ref - Wp)) <= 2.3e-16  # twice-refined fixture == finer fixture
            for f, v in (("coordinates", Xp), ("coordinates_N", Xp), ("coordinates_NM1", Wp)):
                c.upload(f, v)
        else:
            nn, nf = nele + 1, ne + 1
            for b, m in enumerate(maps):
                k, r = np.divmod(m, nn ** 2)
                j, i = np.divmod(r, nn)
                I = np.arange(4 * i.min(), 4 * i.max() + 1)[None, None, :]
                J = np.arange(4 * j.min(), 4 * j.max() + 1)[None, :, None]
                K = np.arange(4 * k.min(), 4 * k.max() + 1)[:, None, None]
                fm = (I + nf * (J + nf * K)).reshape(-1)
                for f, v in (("coordinates", Xp), ("coordinates_N", Xp), ("coordinates_NM1", Wp)):
                    c.upload(f, np.ascontiguousarray(v[:, fm]), block=b)
                maps[b] = fm
        st = c.run_algorithm(6, 0.0)
        x = c.download("coordinates") if procs is None else gather_single(c, maps, "coordinates", 3, (ne + 1) ** 3)
        runs[name] = (st.global_metric, x)
    assert abs(runs["g1"][0] - runs["o1"][0]) <= 1e-9 * abs(runs["o1"][0])
    assert np.max(np.abs(runs["g1"][1] - runs["o1"][1])) <= 1e-9
    assert np.max(np.abs(runs["g2"][1] - runs["g1"][1])) <= 1e-9


@pytest.mark.gpu
def test_gpu_multiblock_refinement_bit_identical_and_connected():
    import percept_b200 as pb
    nele = 6
    X, W = meshes.perturbed_coords(nele, 0.3)
    g, maps = build_multiblock(pb.api, nele, (2, 1, 2), X, W)
    o, _ = build_multiblock(oracle, nele, (2, 1, 2), X, W, {"real_is_double": 1})
    fg, ng = refine(pb.api, g, ("coordinates", "coordinates_NM1", "coordinates_N"))
    fo, no = refine(oracle, o, ("coordinates", "coordinates_NM1", "coordinates_N"))
    assert ng == no == (2 * nele) ** 3
    for b in range(len(maps)):
        assert np.array_equal(fg.download("coordinates", block=b), fo.download("coordinates", block=b))
    assert fg.get_number_nodes() == fo.get_number_nodes() == (2 * nele + 1) ** 3
    mg, ig = fg.metric_and_gradient(1)
    mo, io = fo.metric_and_gradient(1)
    assert ig == io and abs(mg - mo) <= 1e-12 * abs(mo)


@pytest.mark.gpu
def test_gpu_refiner_argument_errors():
    import percept_b200 as pb
    from percept_b200.capi import SmootherError
    coarse = single_block(pb.api, (3, 3, 3), fixture_cube(3))
    other = single_block(pb.api, (3, 3, 3), fixture_cube(3))
    with pytest.raises(SmootherError, match="empty"):
        coarse.refine_topology(other)
    with pytest.raises(SmootherError, match="2x refinement"):
        coarse.refine_field("coordinates", other, "coordinates")
    hexc = pb.api.create()
    hexc.set_unstructured(8, 27, meshes.hex8_conn(2))
    hexc.commit()
    with pytest.raises(SmootherError, match="structured"):
        hexc.refine_topology(pb.api.create())


def test_oracle_refined_rotated_blocks_stay_connected():
    """Non-identity zone-connectivity transform ((-2, 1, 3) / (2, -1, 3), ranges with reversed ends): the r -> 2r-1 map
    of owner and donor ranges (StructuredGridRefiner.cpp:107-131) keeps the refined blocks consistently connected -
    duplicates counted once, interface gradient equal to the refined single block."""
    from tests.test_multiblock import build_two_blocks_rotated
    nele = 4
    X, W = meshes.perturbed_coords(nele, 0.3)
    multi, maps = build_two_blocks_rotated(oracle, nele, X, W, {"real_is_double": 1})
    single = meshes.build(oracle, "sgrid", nele, X, W, True, {"real_is_double": 1})
    fm, nm = refine(oracle, multi, ("coordinates", "coordinates_NM1"))
    fs, ns = refine(oracle, single, ("coordinates", "coordinates_NM1"))
    nf = 2 * nele + 1
    assert nm == ns == (2 * nele) ** 3 and fm.get_number_nodes() == fs.get_number_nodes() == nf ** 3
    # fine index maps of the two blocks (block 1 stored with rotated local axes (a, b, c) = (j, m1 - (i - m), k))
    m = nele // 2
    m1 = nele - m
    i0 = np.arange(0, 2 * m + 1)[None, None, :]
    j0 = np.arange(nf)[None, :, None]
    k0 = np.arange(nf)[:, None, None]
    map0 = (i0 + nf * (j0 + nf * k0)).reshape(-1)
    a = np.arange(nf)[None, None, :]
    b = np.arange(2 * m1 + 1)[None, :, None]
    c = np.arange(nf)[:, None, None]
    map1 = ((2 * m + 2 * m1 - b) + nf * (a + nf * c)).reshape(-1)
    fmaps = [map0, map1]
    # (block 1 sums its face / cell averages in its own rotated axis order: equal to rounding, not bitwise)
    assert np.max(np.abs(gather_single(fm, fmaps, "coordinates", 3, nf ** 3) - fs.download("coordinates"))) <= 4e-16
    for stage in (0, 1):
        ms, is_ = fs.metric_and_gradient(stage)
        mm, im = fm.metric_and_gradient(stage)
        assert im == is_ and abs(mm - ms) <= 1e-10 * max(abs(ms), 1e-300)
        gs = fs.download("cg_g")
        assert np.max(np.abs(gather_single(fm, fmaps, "cg_g", 3, nf ** 3) - gs)) <= 1e-12 * max(np.max(np.abs(gs)), 1e-300)
