#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native Percept mesh-smoothing hot path.

Metric (BASELINE.json): Melem metric+grad evals/s (and CG iterations/s) on a 256^3 hex mesh, fraction of the HBM
roofline, 1/2/4/8 GPUs.  A *step* is one fused ShapeB1 metric+gradient evaluation (get_gradient + total_metric(0),
SM/ReferenceMeshSmootherConjugateGradientDef.hpp:330-388,1481-1520) over the whole mesh.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--nele 256]
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload at N = 1 (config.workload): BASELINE config 3 — synthetic 256^3 unstructured hex8 mesh in Exodus order
(16.78 M elements, 16.97 M nodes), unit cube, interior nodes randomly perturbed (glibc srand(1)/rand(), the
reference test's build_perturbed_cube) by +-0.125 h so the mesh stays valid and every corner takes the full ShapeB1
path; cube faces fixed.  N > 1 is a weak-scaling run: the cube grows to (256 px) x (256 py) x (256 pz) cells and is
cut geometrically into px*py*pz boxes, one per GPU, each with one layer of ghost elements (owner-computes gradient,
no gradient exchange; metric scalars go through ncclAllReduce).

The JSON line carries: value (device-resident throughput), e2e (same metric through the C ABI with pinned HOST
buffers: coordinates H2D, gradient + scalars D2H inside the timed region), roofline (algorithmic bytes / CUDA-event
time of the metric+gradient kernels vs MEASURED_PEAKS.json), cpu_baseline (the CPU restatement of the reference's
functors, OpenMP, timed on this box's host cores on a bounded sample), clocks, gpu_launches.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_BYTES_PER_NODE_MGRAD = 72   # read x (24) + read x_ref (24) + write g (24)   [SURVEY 8(d)]
ALG_BYTES_PER_HEX8_CONN = 32    # int32 connectivity read once per element


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi SM clock / throttle-reason samples during the timed region."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        clocks = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = max([int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()] or [0])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": clocks[len(clocks) // 2] if clocks else None, "sm_max_mhz": mx or None, "reasons": reasons,
                "samples": len(clocks)}


def proc_grid(n):
    return {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}.get(n) or (n, 1, 1)


# ----------------------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU implementation of the path (oracle OpenMP restatement) on this box's host cores
# ----------------------------------------------------------------------------------------------------------------
def cpu_reference(nele_sample, target_seconds, threads=None):
    """Times {get_gradient ; total_metric(0)} of the CPU restatement on an nele_sample^3 hex8 mesh (same generator,
    same perturbation amplitude in units of h).  Returns (elements/s, info dict)."""
    import ctypes as C
    import numpy as np
    from percept_b200.capi import Api
    lib = C.CDLL(os.path.join(ROOT, "oracle", "liborc_omp.so"))
    api = Api(lib, "orc_", product=False)
    lib.orc_fixture_cube.argtypes = [C.c_uint] * 3 + [C.POINTER(C.c_double)] * 3
    lib.orc_perturb_cube_rand.argtypes = [C.c_uint, C.c_double, C.c_uint, C.POINTER(C.c_double)]
    lib.orc_bench_metric_and_gradient.restype = C.c_double
    lib.orc_bench_metric_and_gradient.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_size_t),
                                                  C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.orc_omp_max_threads.restype = C.c_int
    ncores = threads or os.cpu_count() or 1
    n, nn = nele_sample, nele_sample + 1
    W = np.zeros((3, nn ** 3))
    dp = C.POINTER(C.c_double)
    lib.orc_fixture_cube(nn, nn, nn, (C.c_double * 3)(1, 1, 1), (C.c_double * 3)(0, 0, 0), W.ctypes.data_as(dp))
    X = W.copy()
    lib.orc_perturb_cube_rand(n, 0.25 / n, 1, X.ctypes.data_as(dp))
    # same mesh builders as the GPU arm (numpy only, no oracle code on the product path and vice versa)
    ci, cj, ck = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    ci, cj, ck = (a.transpose(2, 1, 0).reshape(-1) for a in (ci, cj, ck))
    nid = lambda i, j, k: i + nn * (j + nn * k)
    offs = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
    conn = np.ascontiguousarray(np.stack([nid(ci + a, cj + b, ck + c) for a, b, c in offs], axis=1).astype(np.int32))
    i = np.arange(nn)
    m = (i[None, None, :] % n == 0) | (i[None, :, None] % n == 0) | (i[:, None, None] % n == 0)
    c = api.create()
    c.set_option("omp_threads", ncores)
    c.set_unstructured(8, nn ** 3, conn)
    c.set_fixed_mask(np.ascontiguousarray(m.reshape(-1).astype(np.uint8)))
    c.commit()
    c.upload("coordinates", X)
    c.upload("coordinates_N", X)
    c.upload("coordinates_NM1", W)
    mt, ni, tm, tg = C.c_double(), C.c_size_t(), C.c_double(), C.c_double()

    def run(reps):
        return lib.orc_bench_metric_and_gradient(c.p, 1, reps, C.byref(mt), C.byref(ni), C.byref(tm), C.byref(tg))
    run(1)  # warm-up (page faults, OpenMP team start)
    t1 = run(1)
    reps = max(1, min(200, int(target_seconds / max(t1, 1e-6))))
    t = run(reps)
    nelem = n ** 3
    return nelem * reps / t, {"cores": ncores, "reps": reps, "seconds": t, "nele_sample": n, "metric": mt.value,
                              "t_metric_s": tm.value, "t_gradient_s": tg.value,
                              "omp_max_threads": lib.orc_omp_max_threads()}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nsample = min(args.nele, 128)
    # K timed steps, each a bounded sample (nsample^3 hexes) of the workload; warm-up inside cpu_reference
    per_step_target = 1.0
    t0 = time.time()
    vals = []
    info = None
    for _ in range(max(1, min(args.steps, 10))):
        v, info = cpu_reference(nsample, per_step_target)
        vals.append(v)
    v = sum(vals) / len(vals)
    out = {
        "impl": "reference", "metric": "Melem metric+grad evals/s", "value": v / 1e6, "unit": "Melem/s", "n_gpus": args.gpus,
        "steps": len(vals), "warmup": 1, "ms_per_step": 1e3 * (nsample ** 3) / v, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.nele}^3 unstructured hex8, random perturbation 0.125h (valid), ShapeB1 fused metric+gradient; "
                               f"each step evaluates a {nsample}^3-hex sample of it on the host cores"},
        "cpu_baseline": {"value": v / 1e6, "unit": "Melem/s", "cores": info["cores"], "kind": "port",
                         "sample": f"{nsample}^3 hex8 x {info['reps']} evaluations per step, OpenMP restatement of the reference functors "
                                   f"(oracle/liborc_omp.so; the reference itself needs Kokkos/Trilinos/STK and cannot be built here)"},
        "e2e": {"value": v / 1e6, "unit": "Melem/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.time() - t0,
    }
    print(json.dumps(out), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import percept_b200 as pb
    from percept_b200 import fixtures as fx
    from percept_b200 import partition as part

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    T0 = time.time()

    def log(msg):
        if rank == 0:
            print(f"[bench {time.time() - T0:7.1f}s] {msg}", file=sys.stderr, flush=True)
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n = args.nele
    px, py, pz = proc_grid(world)
    NX, NY, NZ = n * px, n * py, n * pz
    h = 1.0 / n  # every rank's box is a unit cube in its own frame: identical per-GPU work (weak scaling)

    # ---- local mesh: owned n^3 box + one layer of ghost elements --------------------------------------------
    t_setup = time.time()
    if args.workload == "sgrid":
        return run_ours_sgrid(args, rank, world, local_rank, dist, (NX, NY, NZ), (px, py, pz), h, log, T0)
    lm = part.box_local_mesh((NX, NY, NZ), (px, py, pz), rank)
    # coordinates: global fixture cube scaled so that h is the same for all N; perturbation from the rank-local
    # glibc stream on the owned+ghost box is not globally consistent, so use the deterministic hash-free form:
    # perturb the GLOBAL lattice with the sinusoid-free random generator evaluated per global node id.
    W = lm.global_lattice_coords(h)
    X = part.perturb_by_global_id(W, lm.node_gid, lm.node_on_global_boundary, 0.25 * h, seed=1) if world > 1 else None
    if world == 1:
        Wc = fx.cube(n + 1)
        Xc = fx.perturb_rand(Wc, n + 1, 0.25 / n, 1)  # BASELINE config 3 generator (build_perturbed_cube stream)
        W, X = Wc, Xc
    ctx = pb.create(local_rank)
    ctx.set_unstructured(pb.HEX8, lm.nnodes, lm.conn, lm.elem_owned if world > 1 else None, lm.node_owned if world > 1 else None)
    ctx.set_fixed_mask(lm.node_on_global_boundary.astype(np.uint8))
    ctx.commit()
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            import ctypes as C
            buf = C.create_string_buffer(128)
            assert pb.lib.pms_comm_unique_id(buf) == 0
            uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        ctx.comm_init(bytes(uid.cpu().numpy().tobytes()), rank, world)
        ctx.set_halo(lm.nb_ranks, lm.send_off, lm.send_nodes, lm.recv_off, lm.recv_nodes)
    ctx.upload("coordinates", X)
    ctx.upload("coordinates_N", X)
    ctx.upload("coordinates_NM1", W)
    nel_owned = int(lm.n_owned_elems)
    nnod_local = lm.nnodes
    t_setup = time.time() - t_setup
    log(f"setup done ({t_setup:.1f}s): {nel_owned} owned elements, {nnod_local} local nodes")

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    stream = torch.cuda.ExternalStream(ctx.stream())
    STAGE = 1

    # ---- (1) device-resident: K fused metric+gradient evaluations ---------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()  # samples every 100 ms from warm-up to the end of the timed work (all under load)
    t_w = time.time()
    for _ in range(args.warmup):
        ctx.metric_and_gradient(STAGE)
    # about 0.5 s more of identical load before timing so the clocks are settled and sampled.  The count is decided on
    # rank 0 and broadcast: every call contains an all-reduce, so all ranks must issue the same number of calls.
    ctx.synchronize()
    per_call = max((time.time() - t_w) / args.warmup, 1e-4)
    n_spin = torch.tensor([int(min(2000, 0.5 / per_call))], dtype=torch.int64, device="cuda")
    if dist:
        dist.broadcast(n_spin, 0)
    for _ in range(int(n_spin.item())):
        ctx.metric_and_gradient(STAGE)
    if world > 1:
        ctx.halo_exchange("cg_s")  # first ncclSend/Recv between a pair sets up the P2P channel (~0.7 s, once)
    launches0 = ctx.launch_count()
    ctx.set_timing(1)
    ctx.kernel_time_reset()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(args.steps):
            mtot, ninv = ctx.metric_and_gradient(STAGE)
        e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.launch_count() - launches0
    kern_ms, kern_launches = ctx.kernel_time_ms("gradient")
    ctx.set_timing(0)
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    log(f"device-resident loop done: {ms_step:.3f} ms/step")
    total_elems = nel_owned * world
    value = total_elems / (ms_step * 1e-3) / 1e6  # Melem/s, whole job

    # ---- (2) end to end through the C ABI with pinned host buffers ---------------------------------------------
    hx = torch.from_numpy(np.ascontiguousarray(X)).pin_memory()
    hg = torch.empty((3, nnod_local), dtype=torch.float64).pin_memory()
    hx_np, hg_np = hx.numpy(), hg.numpy()
    e2e_steps = max(3, min(args.steps, 10))

    def e2e_step():
        ctx.upload("coordinates", hx_np)                # H2D 24 B/node
        m, ni = ctx.metric_and_gradient(STAGE)          # kernels + D2H of {mtot, n_invalid}
        ctx.download("cg_g", out=hg_np)                 # D2H 24 B/node
        return m
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_sync = total_elems * e2e_steps / float(dt.item()) / 1e6
    log(f"e2e loop (one evaluation at a time) done: {e2e_sync:.0f} Melem/s")

    # pipelined over consecutive evaluations: every step still uploads its coordinates and downloads its gradient, but
    # the read-back of evaluation k (second stream) and the upload of evaluation k+2 (third stream) run beside the
    # kernels of evaluation k+1 (device staging buffers on both sides, see pms_field_*_async)
    hg2 = torch.empty((3, nnod_local), dtype=torch.float64).pin_memory()
    hbuf = [hg_np, hg2.numpy()]

    def pipelined(nsteps):
        ctx.upload("coordinates", hx_np)                 # input of evaluation 0
        ctx.metric_and_gradient(STAGE)                   # evaluation 0
        ctx.upload_async("coordinates", hx_np)           # input of evaluation 1 on its way
        for k in range(nsteps):
            ctx.download_async("cg_g", hbuf[k % 2])      # D2H 24 B/node: gradient of evaluation k
            ctx.upload_finish()                          # coordinates <- staged input of evaluation k+1
            ctx.upload_async("coordinates", hx_np)       # H2D 24 B/node: input of evaluation k+2
            ctx.metric_and_gradient(STAGE)               # evaluation k+1 (+ D2H of {mtot, n_invalid})
        ctx.upload_finish()
        ctx.transfer_wait()
    pipelined(2)
    barrier()
    t0 = time.perf_counter()
    pipelined(e2e_steps)                                 # e2e_steps+1 uploads+evaluations, e2e_steps read-backs timed
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_val = total_elems * e2e_steps / float(dt.item()) / 1e6
    assert np.array_equal(hbuf[0], hbuf[1])              # same coordinates every step: identical gradients came back
    log(f"e2e loop (transfers overlapped with the kernels of the neighbouring evaluations) done: {e2e_val:.0f} Melem/s")

    # ---- (3) CG iterations/s (device resident; the reference's run_one_iteration incl. line search) ------------
    cg = {}
    try:
        st = ctx.state_init()
        st.stage = 1
        ctx.upload("coordinates", X)
        n_it = 10
        st.num_invalid = ctx.count_invalid_elements()
        gm = ctx.run_one_iteration(st, 0.0)  # iteration 0 (edge lengths, first direction) is set-up: untimed
        ev0 = int(st.n_metric_evals)
        barrier()
        t0 = time.perf_counter()
        for it in range(1, n_it + 1):
            st.iter = it
            st.num_invalid = ctx.count_invalid_elements()
            gm = ctx.run_one_iteration(st, 0.0)
        barrier()
        dtc = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if dist:
            dist.all_reduce(dtc, op=dist.ReduceOp.MAX)
        cg = {"cg_iters_per_s": n_it / float(dtc.item()), "iters": n_it, "metric_evals": int(st.n_metric_evals) - ev0,
              "gradient_evals": int(st.n_gradient_evals), "final_metric": gm, "stage": "ShapeB1"}
    except pb.SmootherError as ex:  # report, do not hide
        cg = {"error": str(ex)}

    log(f"cg loop done: {cg}")
    clocks = sampler.stop()
    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    # ---- roofline of the fused metric+gradient evaluation (element pass + deterministic gather) -----------------
    peak, peak_src = hbm_peak()
    alg_bytes = ALG_BYTES_PER_NODE_MGRAD * nnod_local + ALG_BYTES_PER_HEX8_CONN * lm.nelems
    kernel_ms_per_eval = kern_ms / max(args.steps, 1)
    achieved = alg_bytes / (kernel_ms_per_eval * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("mgrad_dram_bytes_per_eval")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "kernel": "k_elem_pipe<hex8,ShapeB1,grad> + k_grad_gather_csr (one evaluation)",
                "kernel_ms_per_eval": kernel_ms_per_eval, "algorithmic_bytes_per_eval": alg_bytes,
                "note": "the evaluation is FP64-pipe-bound (about 1400 FP64 instr/hex vs 64 lanes/clk/SM), see DESIGN.md; "
                        "fp64 pipe utilisation is in profiles/"}
    # the binding resource, stated beside the contract's HBM figure: FP64 thread-instructions of the evaluation
    # (1282 per hex: 1410 by SASS count before the Gram-matrix form of the derivative and the merged rsqrt took 16 per corner off) against the measured FP64 pipe rate of
    # this part (148 SMs x 64 lanes/clk, profiles/r01_fp64_peak.txt) at the SM clock sampled during the run
    fp64_instr = 1282.0 * lm.nelems
    fp64_peak = 148 * 64 * (clocks.get("sm_mhz") or 1965) * 1e6
    roofline["fp64_pipe"] = {"achieved_ginstr_s": fp64_instr / (kernel_ms_per_eval * 1e-3) / 1e9, "peak_ginstr_s": fp64_peak / 1e9,
                             "frac": fp64_instr / (kernel_ms_per_eval * 1e-3) / fp64_peak,
                             "note": "whole evaluation incl. the HBM-bound gather pass (25 % of the time); the element pass alone "
                                     "ran the FP64 pipe at 71.0 % in the final ncu capture of profiles/r01_ncu_summary.txt"}

    # ---- CPU baseline on this box's host cores, bounded sample --------------------------------------------------
    try:
        cpu_v, info = cpu_reference(min(n, 128), 12.0)
        cpu_baseline = {"value": cpu_v / 1e6, "unit": "Melem/s", "cores": info["cores"], "kind": "port",
                        "sample": f"{info['nele_sample']}^3 hex8 (1/{(n // info['nele_sample']) ** 3} of the workload) x {info['reps']} evaluations, "
                                  f"{info['seconds']:.1f} s; OpenMP restatement of the reference functors (oracle/liborc_omp.so)"}
    except Exception as ex:
        cpu_baseline = {"value": None, "unit": "Melem/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}

    out = {
        "metric": "Melem metric+grad evals/s", "value": value, "unit": "Melem/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{n}^3 unstructured hex8 per GPU ({nel_owned} owned elements, {nnod_local} local nodes), unit-cube lattice, "
                               f"random interior perturbation 0.125h (valid mesh), cube faces fixed, ShapeB1 fused metric+gradient",
                   "global_cells": [NX, NY, NZ], "partition": f"{px}x{py}x{pz} boxes, 1 ghost layer, owner-computes",
                   "l2": "inputs (815 MB of coordinates per GPU) exceed the 126 MB L2; no explicit flush"},
        "e2e": {"value": e2e_val, "unit": "Melem/s", "h2d_bytes_per_step": 24 * nnod_local, "d2h_bytes_per_step": 24 * nnod_local + 16,
                "steps": e2e_steps, "api": "per step: pms_field_download_async(cg_g of evaluation k) || pms_field_upload_async(coordinates of k+2) || pms_metric_and_gradient(k+1); pinned host buffers, three streams",
                "one_evaluation_at_a_time": e2e_sync},
        "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clocks,
        "cg": cg, "check": {"total_metric": mtot, "n_invalid": int(ninv)}, "setup_s": t_setup,
    }
    print(json.dumps(out), flush=True)
    if dist:
        dist.destroy_process_group()


def run_ours_sgrid(args, rank, world, local_rank, dist, cells, procs, h, log, T0):
    """Structured variant (BASELINE configs 2 and 5): one n^3-cell StructuredGrid block per GPU, sinusoidal interior
    perturbation, interface-node sums over NCCL.  Same JSON keys; config.workload names it."""
    import numpy as np
    import torch
    import percept_b200 as pb
    from percept_b200 import partition as part
    n = args.nele
    lb = part.sgrid_block(cells, procs, rank)
    W = lb.global_lattice_coords(h)
    L = [c * h for c in cells]
    b = 0.2 * h * np.sin(2 * np.pi * 3 * W[0] / L[0]) * np.sin(2 * np.pi * 5 * W[1] / L[1]) * np.sin(2 * np.pi * 7 * W[2] / L[2])
    X = np.ascontiguousarray(W + b[None, :] * (lb.fixed == 0)[None, :])
    ctx = pb.create(local_rank)
    ctx.add_structured_block(lb.node_size)
    ctx.set_fixed_mask(lb.fixed)
    ctx.commit()
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            import ctypes as C
            buf = C.create_string_buffer(128)
            assert pb.lib.pms_comm_unique_id(buf) == 0
            uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        ctx.comm_init(bytes(uid.cpu().numpy().tobytes()), rank, world)
        ctx.set_interface(lb.nb_ranks, lb.if_off, lb.if_nodes, lb.owned)
    for f in ("coordinates", "coordinates_N"):
        ctx.upload(f, X)
    ctx.upload("coordinates_NM1", W)
    nel, nnod = ctx.num_elements_local, ctx.num_nodes_local
    log(f"sgrid setup done: {nel} cells, {nnod} nodes per GPU")

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
    stream = torch.cuda.ExternalStream(ctx.stream())
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 3) + 20):
        ctx.metric_and_gradient(1)
    launches0 = ctx.launch_count()
    ctx.set_timing(1)
    ctx.kernel_time_reset()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(args.steps):
            mtot, ninv = ctx.metric_and_gradient(1)
        e1.record(stream)
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    launches = ctx.launch_count() - launches0
    kern_ms, _ = ctx.kernel_time_ms("gradient")
    ctx.set_timing(0)
    # metric-only (line-search) evaluations and CG iterations
    ctx.set_option("cache_metric", 0)
    s = 1e-3 * h * np.random.default_rng(0).standard_normal((3, nnod)) * (lb.fixed == 0)[None, :]
    ctx.upload("cg_s", s)
    if world > 1:
        pass
    barrier()
    t0 = time.perf_counter()
    for _ in range(10):
        ctx.total_metric(1, 0.5)
    barrier()
    ms_metric = (time.perf_counter() - t0) / 10 * 1e3
    ctx.set_option("cache_metric", 1)
    st = ctx.state_init()
    st.stage = 1
    st.num_invalid = ctx.count_invalid_elements()
    ctx.run_one_iteration(st, 0.0)
    ev0 = int(st.n_metric_evals)
    barrier()
    t0 = time.perf_counter()
    for it in range(1, 11):
        st.iter = it
        st.num_invalid = ctx.count_invalid_elements()
        gm = ctx.run_one_iteration(st, 0.0)
    barrier()
    dtc = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(dtc, op=dist.ReduceOp.MAX)
    clocks = sampler.stop()
    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return
    peak, peak_src = hbm_peak()
    alg_bytes = ALG_BYTES_PER_NODE_MGRAD * nnod
    kernel_ms_per_eval = kern_ms / args.steps
    out = {"metric": "Melem metric+grad evals/s", "value": nel * world / (ms_step * 1e-3) / 1e6, "unit": "Melem/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"one {n}^3-cell StructuredGrid block per GPU ({nnod} nodes), sinusoidal interior perturbation 0.2h, "
                                  f"block faces on the global boundary fixed, ShapeB1 fused metric+gradient", "global_cells": list(cells),
                      "partition": f"{procs[0]}x{procs[1]}x{procs[2]} blocks, interface-node sums over NCCL"},
           "gpu_launches": int(launches),
           "roofline": {"bound": "hbm", "achieved": alg_bytes / (kernel_ms_per_eval * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": alg_bytes / (kernel_ms_per_eval * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                        "kernel_ms_per_eval": kernel_ms_per_eval, "algorithmic_bytes_per_eval": alg_bytes},
           "metric_only": {"ms_per_eval_with_direction": ms_metric, "Melem_per_s": nel * world / (ms_metric * 1e-3) / 1e6},
           "cg": {"cg_iters_per_s": 10 / float(dtc.item()), "iters": 10, "metric_evals": int(st.n_metric_evals) - ev0, "final_metric": gm},
           "clocks": clocks, "check": {"total_metric": mtot, "n_invalid": int(ninv)}}
    print(json.dumps(out), flush=True)
    if dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nele", type=int, default=256, help="cells per direction per GPU")
    ap.add_argument("--workload", default="hex8", choices=["hex8", "sgrid"],
                    help="hex8: BASELINE config 3 (headline, default); sgrid: one structured block per GPU (configs 2/5)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
