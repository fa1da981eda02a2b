"""Times individual multi-rank operations (2+ ranks under torchrun) to locate collective overheads."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import percept_b200 as pb
from percept_b200 import partition as part
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local); dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
procs = {2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[world]; cells = tuple(n * p for p in procs); h = 1.0 / n
lm = part.box_local_mesh(cells, procs, rank)
W = lm.global_lattice_coords(h); X = part.perturb_by_global_id(W, lm.node_gid, lm.node_on_global_boundary, 0.25 * h)
c = pb.create(local); c.set_unstructured(pb.HEX8, lm.nnodes, lm.conn, lm.elem_owned, lm.node_owned); c.set_fixed_mask(lm.node_on_global_boundary); c.commit()
uid = torch.zeros(128, dtype=torch.uint8)
if rank == 0:
    buf = C.create_string_buffer(128); pb.lib.pms_comm_unique_id(buf); uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
uid = uid.cuda(); dist.broadcast(uid, 0); c.comm_init(uid.cpu().numpy().tobytes(), rank, world)
c.set_halo(lm.nb_ranks, lm.send_off, lm.send_nodes, lm.recv_off, lm.recv_nodes)
for f, v in (("coordinates", X), ("coordinates_N", X), ("coordinates_NM1", W)): c.upload(f, v)
c.set_option("cache_metric", 0)
c.get_edge_lengths(); c.get_gradient(1); c.axpby(-1.0, "cg_g", 0.0, "cg_s")
def T(name, fn, reps=20):
    fn(); c.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    c.synchronize(); dt = (time.perf_counter() - t0) / reps
    if rank == 0: print(f"{name:28s} {dt*1e3:9.3f} ms", flush=True)
T("dot (allreduce sum)", lambda: c.dot("cg_s", "cg_g"))
T("total_metric a=0", lambda: c.total_metric(1, 0.0))
T("total_metric a=1e-3", lambda: c.total_metric(1, 1e-3))
T("get_alpha_0 (min)", lambda: c.get_alpha_0())
T("halo_exchange cg_s", lambda: c.halo_exchange("cg_s"))
T("copy_field s<-r (+halo)", lambda: c.copy_field("cg_s", "cg_s" if False else "cg_r"))
T("metric_and_gradient", lambda: c.metric_and_gradient(1))
T("count_invalid", lambda: c.count_invalid_elements())
st = c.state_init(); st.stage = 1
c.set_option("cache_metric", 1)
t0 = time.perf_counter(); st.iter = 0; st.num_invalid = c.count_invalid_elements(); c.run_one_iteration(st, 0.0); c.synchronize()
if rank == 0: print("first CG iteration", time.perf_counter() - t0, "s, metric evals", st.n_metric_evals, flush=True)
t0 = time.perf_counter(); st.iter = 1; st.num_invalid = c.count_invalid_elements(); c.run_one_iteration(st, 0.0); c.synchronize()
if rank == 0: print("second CG iteration", time.perf_counter() - t0, "s, metric evals", st.n_metric_evals, flush=True)
dist.barrier(); dist.destroy_process_group()
