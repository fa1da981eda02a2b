"""`python -m percept_b200.smooth` — the smoothing step of `mesh_adapt --smooth_geometry=1` on the B200 path.

Option names and defaults follow mesh_adapt (src/adapt/main/MeshAdapt.cpp:795-821, MeshAdaptMemberVarInit.hpp:64-75):
  --smoother_niter (1000)  --smoother_tol (1e-4)  --smooth_use_reference_mesh (1)  --fix_all_block_boundaries (0: fix the
  mesh skin, like the surface-part boundary selector of Refiner::snapAndSmooth, Refiner.cpp:1536-1565).
--smooth_surfaces=1 needs --surfaces FILE (JSON list of analytic surfaces, percept_b200/geometry.py): skin nodes on one
surface slide on it (tangent-plane projection + snapping on the GPU), nodes on two surfaces (curves) or three (vertices)
stay fixed; a CAD kernel (OpenNURBS / PGEOM) is not on this path.

Input: an Exodus II file (classic netCDF-3) with hex8 or tet4 blocks holding the mesh to smooth; --reference_mesh gives
the undistorted mesh (coordinates_NM1) with identical connectivity; without it the input itself is the reference mesh.
"""
import argparse
import sys

import numpy as np


def main(argv=None):
    ap = argparse.ArgumentParser(prog="percept_b200.smooth", description=__doc__)
    ap.add_argument("input_mesh")
    ap.add_argument("output_npz")
    ap.add_argument("--reference_mesh", default=None)
    ap.add_argument("--smoother_niter", type=int, default=1000)
    ap.add_argument("--smoother_tol", type=float, default=1e-4)
    ap.add_argument("--smooth_use_reference_mesh", type=int, default=1)
    ap.add_argument("--fix_all_block_boundaries", type=int, default=0)
    ap.add_argument("--smooth_surfaces", type=int, default=0)
    ap.add_argument("--surfaces", default=None, help="JSON list of analytic surfaces (plane / sphere / cylinder)")
    ap.add_argument("--surface_tol", type=float, default=1e-8, help="classification distance, in mesh units")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--log", default=None, help="smootherlog file (default smootherlog<input>.txt)")
    a = ap.parse_args(argv)
    if a.smooth_surfaces and not a.surfaces:
        sys.exit("--smooth_surfaces=1 needs --surfaces FILE (analytic surfaces); CAD geometry queries are not on the B200 path")
    import percept_b200 as pb
    from percept_b200 import exodus
    m = exodus.read_exodus(a.input_mesh)
    topos = {t for t, _ in m["blocks"]}
    if len(topos) != 1:
        sys.exit("mixed hex8/tet4 meshes: smooth one topology per run")
    topo = topos.pop()
    conn = np.concatenate([c for _, c in m["blocks"]])
    X = m["coords"]
    W = exodus.read_exodus(a.reference_mesh)["coords"] if a.reference_mesh else X.copy()
    fixed = exodus.skin_nodes(conn, topo)
    if a.fix_all_block_boundaries and len(m["blocks"]) > 1:  # also the faces between element blocks
        for _, c in m["blocks"]:
            fixed |= np.pad(exodus.skin_nodes(c, topo), (0, fixed.size - int(c.max()) - 1))
    ctx = pb.create(a.device)
    ctx.set_option("use_ref_mesh", a.smooth_use_reference_mesh)
    ctx.set_unstructured(topo, X.shape[1], conn)
    ctx.set_fixed_mask(fixed)
    ctx.commit()
    if a.smooth_surfaces:
        from percept_b200 import geometry
        surfaces = geometry.load_surfaces(a.surfaces)
        for s_ in surfaces:
            ctx.add_analytic_surface(*geometry.surface_params(s_))
        dof, surf = geometry.classify_nodes(W, fixed, surfaces, a.surface_tol)
        ctx.set_option("smooth_surfaces", 1)
        ctx.set_node_classification(dof, surf)   # replaces the skin mask: surface nodes slide, curves / vertices stay
        print(f"MeshSmoother: {int((dof == 2).sum())} surface, {int((dof == 1).sum())} curve, {int((dof == 0).sum())} vertex nodes")
    ctx.upload("coordinates_NM1", W)
    ctx.upload("coordinates_N", X)
    ctx.upload("coordinates", X)
    n0 = ctx.count_invalid_elements()
    print(f"MeshSmoother: number of invalid elements before untangling = {n0}")
    st = ctx.run_algorithm(a.smoother_niter, a.smoother_tol, a.log or f"smootherlog{a.input_mesh.replace('/', '_')}.txt")
    print(f"MeshSmoother: stage {st.stage} iter {st.iter} metric {st.global_metric:.6e} invalid {st.num_invalid} dmax_rel {st.dmax_relative:.3e}")
    exodus.write_coordinates_npz(a.output_npz, ctx.download("coordinates"), node_num_map=m["node_num_map"], fixed=fixed)


if __name__ == "__main__":
    main()
