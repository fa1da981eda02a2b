"""First-look device timings of the hot kernels (CUDA events per kernel family on the ctx stream)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import percept_b200 as pb
from percept_b200 import fixtures as fx

def run(kind, nele, eps, reps=int(os.environ.get("QB_REPS", "5"))):
    nn = nele + 1
    W = fx.cube(nn); X = fx.perturb_rand(W, nn, eps / nele, 1)
    c = pb.create()
    t0 = time.time()
    if kind == "sgrid":
        b = c.add_structured_block((nn,) * 3); c.add_fixture_1_bcs(b); c.select_boundary_sgrid()
    else:
        conn = fx.hex8_connectivity(nele)
        if kind == "tet4": conn = fx.tet4_from_hex8(conn)
        c.set_unstructured(8 if kind == "hex8" else 4, nn ** 3, conn); c.set_fixed_mask(fx.cube_boundary_mask(nn))
    c.commit()
    c.upload("coordinates", X); c.upload("coordinates_N", X); c.upload("coordinates_NM1", W)
    c.set_option("cache_metric", 0)
    if os.environ.get("QB_FUSED"):
        c.set_option("sgrid_fused", int(os.environ["QB_FUSED"]))
    if os.environ.get("QB_ELEMWS"):
        c.set_option("elem_ws", int(os.environ["QB_ELEMWS"]))
    if os.environ.get("QB_VARIANT"):
        vm, vg = os.environ["QB_VARIANT"].split(",")
        c.set_option("variant_metric_corners", int(vm)); c.set_option("variant_grad_corners", int(vg))
    ne, nnod = c.num_elements_local, c.num_nodes_local
    print(f"--- {kind} {nele}^3: {ne} elems, {nnod} nodes, setup {time.time()-t0:.1f}s")
    s = 1e-3 / nele * np.random.default_rng(0).standard_normal((3, nnod)); c.upload("cg_s", s)
    c.set_timing(1)
    ops = os.environ.get("QB_OPS", "metric0,metricA,mgrad").split(",")
    for stage in [int(v) for v in os.environ.get("QB_STAGES", "0,1").split(",") if v != ""]:
        for what in ops:
            for it in range(reps + 2):
                if it == 2: c.kernel_time_reset()
                if what == "multi": c.total_metric_multi(stage, [0.5 * 0.5 ** k for k in range(8)]); continue
                if what == "multi12": c.total_metric_multi(stage, [0.5 * 0.5 ** k for k in range(12)]); continue
                if what == "multi16": c.total_metric_multi(stage, [0.5 * 0.5 ** k for k in range(16)]); continue
                if what == "metric0": c.total_metric(stage, 0.0)
                elif what == "metricA": c.total_metric(stage, 0.5)
                else: c.metric_and_gradient(stage)
            fam = "gradient" if what == "mgrad" else "metric"
            ms, n = c.kernel_time_ms(fam)
            ms /= reps
            bpn = {"metric0": 48, "metricA": 72, "mgrad": 72, "multi": 72, "multi12": 72, "multi16": 72}[what]
            extra = 0 if kind == "sgrid" else (32 if kind == "hex8" else 16) * ne
            gb = (bpn * nnod + extra) / 1e9
            print(f"stage {stage} {what:8s}: {ms:8.3f} ms  {ne/ms/1e6:8.2f} Gelem/s  alg {gb/ms*1e3:7.1f} GB/s ({gb/ms*1e3/6561.6*100:5.1f}% of 6561.6)")
    if kind == "sgrid" and "refine" in os.environ.get("QB_OPS", "refine"):
        # StructuredGridRefiner prolongation: nele^3 -> (2 nele)^3 cells
        f = pb.create(); c.refine_topology(f); f.commit(); f.set_timing(1)
        for it in range(reps + 2):
            if it == 2: f.kernel_time_reset()
            c.refine_field("coordinates", f, "coordinates")
        ms, n = f.kernel_time_ms("refine"); ms /= reps
        gb = 24 * (nnod + f.num_nodes_local) / 1e9
        print(f"refine {nele}^3 -> {2*nele}^3: {ms:8.3f} ms  {f.num_nodes_local/ms/1e6:8.2f} Gnode/s  alg {gb/ms*1e3:7.1f} GB/s ({gb/ms*1e3/6561.6*100:5.1f}% of 6561.6)")
        f.close()
    if os.environ.get("QB_OPS"): c.close(); return
    # BLAS-1 style kernels
    for name, fn, bpn in (("dot", lambda: c.dot("cg_s", "cg_g"), 48), ("axpby", lambda: c.axpby(1.0, "cg_g", 0.5, "cg_s"), 72),
                          ("copy", lambda: c.copy_field("cg_d", "cg_s"), 48)):
        for it in range(reps + 2):
            if it == 2: c.kernel_time_reset()
            fn()
        ms, n = c.kernel_time_ms("blas1"); ms /= reps
        print(f"{name:8s}: {ms:8.3f} ms  {bpn*nnod/ms/1e6:8.1f} GB/s")
    st = c.state_init()
    t0 = time.time(); c.set_timing(0); c.set_option("cache_metric", 1)
    st.stage = 0
    for it in range(5):
        st.iter = it; st.num_invalid = c.count_invalid_elements()
        m = c.run_one_iteration(st, 0.0)
    c.synchronize()
    dt = time.time() - t0
    print(f"5 untangle CG iterations: {dt/5*1e3:.2f} ms/iter wall, metric evals {st.n_metric_evals}, grad evals {st.n_gradient_evals}, metric {m:.6e} ninv {st.num_invalid}")
    c.close()

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    for kind in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["sgrid", "hex8", "tet4"]):
        run(kind, n, 1.0)
