"""A/B of the end-to-end pipelines on one box: sequential, read-back || next upload, fully staged three-stream."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import percept_b200 as pb
from percept_b200 import fixtures as fx
nele = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nn = nele + 1
W = fx.cube(nn); X = fx.perturb_rand(W, nn, 0.125 / nele, 1)
c = pb.create(); c.set_unstructured(8, nn ** 3, fx.hex8_connectivity(nele)); c.set_fixed_mask(fx.cube_boundary_mask(nn)); c.commit()
for f, v in (("coordinates", X), ("coordinates_N", X), ("coordinates_NM1", W)): c.upload(f, v)
hx = torch.from_numpy(np.ascontiguousarray(X)).pin_memory().numpy()
hg = [torch.empty((3, nn ** 3), dtype=torch.float64).pin_memory().numpy() for _ in range(2)]
ne = c.num_elements_local
def seq(n):
    for k in range(n):
        c.upload("coordinates", hx); c.metric_and_gradient(1); c.download("cg_g", out=hg[0])
def two(n):
    c.upload("coordinates", hx); c.metric_and_gradient(1)
    for k in range(n):
        c.download_async("cg_g", hg[k % 2]); c.upload("coordinates", hx); c.metric_and_gradient(1)
    c.transfer_wait()
def three(n):
    c.upload("coordinates", hx); c.metric_and_gradient(1); c.upload_async("coordinates", hx)
    for k in range(n):
        c.download_async("cg_g", hg[k % 2]); c.upload_finish(); c.upload_async("coordinates", hx); c.metric_and_gradient(1)
    c.upload_finish(); c.transfer_wait()
def h2d_only(n):
    for k in range(n): c.upload("coordinates", hx)
def d2h_only(n):
    for k in range(n): c.download("cg_g", out=hg[0])
for name, fn in (("h2d only", h2d_only), ("d2h only", d2h_only), ("sequential", seq), ("two-stream", two), ("three-stream", three), ("two-stream", two), ("three-stream", three)):
    fn(2); c.synchronize()
    t0 = time.perf_counter(); fn(10); c.synchronize(); dt = (time.perf_counter() - t0) / 10
    print(f"{name:13s} {dt*1e3:7.2f} ms/step  {ne/dt/1e6:8.0f} Melem/s   ({24*nn**3/dt/1e9:5.1f} GB/s per direction)")
