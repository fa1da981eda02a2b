"""Small odd-sized cases that touch every kernel (all element kinds, stages, variants, multi-block, strict build);
run under compute-sanitizer --tool memcheck / racecheck."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import percept_b200 as pb
from percept_b200 import fixtures as fx

def run(kind, nn, strict, variant):
    W = fx.cube(nn); rng = np.random.default_rng(1)
    X = W + 0.2 / max(nn) * (rng.random(W.shape) - 0.5) * (fx.cube_boundary_mask(nn) == 0)[None, :]
    c = pb.create()
    c.set_option("strict_fp", strict); c.set_option("variant_metric_corners", variant); c.set_option("variant_grad_corners", variant)
    c.set_option("keep_element_metrics", 1)
    if kind == "sgrid":
        b = c.add_structured_block(nn); c.add_fixture_1_bcs(b); c.select_boundary_sgrid()
    else:
        conn = fx.hex8_connectivity((nn[0] - 1, nn[1] - 1, nn[2] - 1))
        if kind == "tet4": conn = fx.tet4_from_hex8(conn)
        c.set_unstructured(8 if kind == "hex8" else 4, W.shape[1], conn); c.set_fixed_mask(fx.cube_boundary_mask(nn))
    c.commit()
    for f, v in (("coordinates", X), ("coordinates_N", X), ("coordinates_NM1", W)): c.upload(f, v)
    st = c.run_algorithm(3, 0.0)
    c.set_option("cache_metric", 0)
    c.total_metric(1, 0.3); c.metric_and_gradient(0); c.get_alpha_0(); c.dot("cg_g", "cg_s")
    c.close()
    return st.global_metric

for kind in ("sgrid", "hex8", "tet4"):
    for strict, variant in ((0, 2), (0, 3), (0, 1), (0, 0), (1, 2)):
        print(kind, strict, variant, run(kind, (14, 8, 6), strict, variant), flush=True)
print("SANITIZE_CASE_DONE")
