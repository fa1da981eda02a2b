"""Per-kernel-family device time of CG iterations (ShapeB1 stage) on the bench mesh: python tools/cg_breakdown.py [nele] [kind]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import percept_b200 as pb
from percept_b200 import fixtures as fx

nele = int(sys.argv[1]) if len(sys.argv) > 1 else 256
kind = sys.argv[2] if len(sys.argv) > 2 else "hex8"
nn = nele + 1
W = fx.cube(nn); X = fx.perturb_rand(W, nn, 0.25 / nele, 1)
c = pb.create()
if kind == "sgrid":
    b = c.add_structured_block((nn,) * 3); c.add_fixture_1_bcs(b); c.select_boundary_sgrid()
else:
    c.set_unstructured(8, nn ** 3, fx.hex8_connectivity(nele)); c.set_fixed_mask(fx.cube_boundary_mask(nn))
c.commit()
for f, v in (("coordinates", X), ("coordinates_N", X), ("coordinates_NM1", W)):
    c.upload(f, v)
st = c.state_init(); st.stage = 1
niter = 6
halv = []
for it in range(niter):
    h0 = st.n_line_search_halvings
    if it == 1:
        c.synchronize(); c.set_timing(1); c.kernel_time_reset(); t0 = time.time(); ev0 = (st.n_metric_evals, st.n_gradient_evals)
    st.iter = it; st.num_invalid = c.count_invalid_elements()
    m = c.run_one_iteration(st, 0.0)
    halv.append(st.n_line_search_halvings - h0)
c.synchronize(); wall = (time.time() - t0) / (niter - 1) * 1e3
tot = 0
for fam in ("metric", "gradient", "blas1", "update", "alpha0", "edge", "halo"):
    ms, n = c.kernel_time_ms(fam); tot += ms
    print(f"{fam:9s} {ms/(niter-1):8.3f} ms/iter  {n/(niter-1):6.1f} launches/iter")
print(f"kernels {tot/(niter-1):.3f} ms/iter, wall (timing on) {wall:.3f} ms/iter, metric evals/iter {(st.n_metric_evals-ev0[0])/(niter-1):.1f}, grad evals/iter {(st.n_gradient_evals-ev0[1])/(niter-1):.1f}, metric {m:.6e}")
c.set_timing(0)
t0 = time.time()
for it in range(niter, niter + 5):
    h0 = st.n_line_search_halvings
    st.iter = it; st.num_invalid = c.count_invalid_elements(); c.run_one_iteration(st, 0.0)
    halv.append(st.n_line_search_halvings - h0)
c.synchronize(); print(f"wall (timing off) {(time.time()-t0)/5*1e3:.3f} ms/iter")
print("line-search halvings per iteration:", halv)
