#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native Percept mesh-smoothing hot path.

Metric (BASELINE.json): Melem metric+grad evals/s (and CG iterations/s) on a 256^3 hex mesh, fraction of the HBM
roofline, 1/2/4/8 GPUs.  A *step* is one fused ShapeB1 metric+gradient evaluation (get_gradient + total_metric(0),
SM/ReferenceMeshSmootherConjugateGradientDef.hpp:330-388,1481-1520) over the whole mesh.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--nele 256] [--configs C2,C4,C5]
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload (config.workload; `value`, `e2e`, `roofline`, `cg`, `untangle`): BASELINE config 3 — synthetic 256^3
unstructured hex8 mesh in Exodus order (16.78 M elements, 16.97 M nodes) per GPU, unit cube, cube faces fixed.  The
timed step runs ShapeB1 on the mesh perturbed by +-0.125 h (valid, every corner takes the full path); the `untangle`
block runs stage 0 on the mesh perturbed by +-0.5 h (P(1/256): ~40 % of the elements inverted).  N > 1 is a weak-scaling
run: (256 px) x (256 py) x (256 pz) cells cut into boxes, one per GPU, one layer of ghost elements (owner-computes
gradient; cg_s ghost exchange per CG iteration; scalars through ncclAllReduce).

`configs` holds the other BASELINE configurations measured in the same run, each with its own value / roofline / cg:
  C2  one 128^3-cell StructuredGrid block, sinusoidal perturbation (N = 1 only)
  C4  119^3 x 6 = 10.1 M tet4, deterministic RCB over N ranks, cg_s halo exchange
  C5  one 512^3-cell StructuredGrid block PER GPU (2x2x2 at N = 8), interface-node gradient sums over ncclSend/Recv
      INSIDE the timed step (exchange_ms = their share)
`check` compares the GPU path with the CPU oracle on identical inputs (128^3 at N = 1; a partitioned 2x-smaller global
mesh against the single-block oracle at N > 1, including ghost copies == owners after 3 CG iterations).
"""
import argparse
import ctypes as C
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
ALG_BYTES_PER_TET4_CONN = 16
FP64_INSTR_PER_HEX_MGRAD = 1223.0  # ShapeB1 metric+gradient, FP64 thread-instructions per hex (static SASS count of the element
                                   # kernel, tools/sass_mix.py; profiles/r02_traffic.json)


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
# the CPU oracle (oracle/liborc*.so): the checker, and the timed CPU baseline.  Bound through percept_b200.capi.Api
# (a ctypes signature table); importing that module does not load the CUDA library.
# ----------------------------------------------------------------------------------------------------------------
class Oracle:
    def __init__(self, omp=True):
        from percept_b200.capi import Api
        self.lib = C.CDLL(os.path.join(ROOT, "oracle", "liborc_omp.so" if omp else "liborc.so"))
        self.api = Api(self.lib, "orc_", product=False)
        dp = C.POINTER(C.c_double)
        self.lib.orc_fixture_cube.argtypes = [C.c_uint] * 3 + [dp] * 3
        self.lib.orc_perturb_cube_rand.argtypes = [C.c_uint, C.c_double, C.c_uint, dp]
        self.lib.orc_bench_metric_and_gradient.restype = C.c_double
        self.lib.orc_bench_metric_and_gradient.argtypes = [C.c_void_p, C.c_int, C.c_int, dp, C.POINTER(C.c_size_t), dp, dp]
        self.lib.orc_omp_max_threads.restype = C.c_int

    def cube(self, nn):
        import numpy as np
        W = np.zeros((3, nn ** 3))
        self.lib.orc_fixture_cube(nn, nn, nn, (C.c_double * 3)(1, 1, 1), (C.c_double * 3)(0, 0, 0), W.ctypes.data_as(C.POINTER(C.c_double)))
        return W

    def perturb(self, W, n, eps, seed=1):
        X = W.copy()
        self.lib.orc_perturb_cube_rand(n, eps, seed, X.ctypes.data_as(C.POINTER(C.c_double)))
        return X


def hex8_conn_np(n):
    import numpy as np
    nn = n + 1
    ci, cj, ck = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    ci, cj, ck = (a.transpose(2, 1, 0).reshape(-1) for a in (ci, cj, ck))
    nid = lambda i, j, k: i + nn * (j + nn * k)
    offs = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
    return np.ascontiguousarray(np.stack([nid(ci + a, cj + b, ck + c) for a, b, c in offs], axis=1).astype(np.int32))


def cube_face_mask_np(nn):
    import numpy as np
    i = np.arange(nn)
    m = (i[None, None, :] % (nn - 1) == 0) | (i[None, :, None] % (nn - 1) == 0) | (i[:, None, None] % (nn - 1) == 0)
    return np.ascontiguousarray(m.reshape(-1).astype(np.uint8))


def oracle_hex8_ctx(orc, n, eps_factor, threads):
    """The oracle on an n^3 hex8 cube (build_perturbed_cube stream, amplitude eps_factor/n), same generator as the GPU arm."""
    nn = n + 1
    W = orc.cube(nn)
    X = orc.perturb(W, n, eps_factor / n, 1)
    c = orc.api.create()
    c.set_option("omp_threads", threads)
    c.set_unstructured(8, nn ** 3, hex8_conn_np(n))
    c.set_fixed_mask(cube_face_mask_np(nn))
    c.commit()
    for f, v in (("coordinates", X), ("coordinates_N", X), ("coordinates_NM1", W)):
        c.upload(f, v)
    return c, X, W


def cpu_reference(orc, nele_sample, target_seconds, threads=None, ctx=None):
    """Times {get_gradient ; total_metric(0)} of the CPU restatement on an nele_sample^3 hex8 mesh (same generator,
    same perturbation amplitude in units of h).  Returns (elements/s, info dict)."""
    ncores = threads or os.cpu_count() or 1
    c = ctx or oracle_hex8_ctx(orc, nele_sample, 0.25, ncores)[0]
    mt, ni, tm, tg = C.c_double(), C.c_size_t(), C.c_double(), C.c_double()

    def run(reps):
        return orc.lib.orc_bench_metric_and_gradient(c.p, 1, reps, C.byref(mt), C.byref(ni), C.byref(tm), C.byref(tg))
    run(1)  # warm-up (page faults, OpenMP team start)
    t1 = run(1)
    reps = max(1, min(200, int(target_seconds / max(t1, 1e-6))))
    t = run(reps)
    nelem = nele_sample ** 3
    return nelem * reps / t, {"cores": ncores, "reps": reps, "seconds": t, "nele_sample": nele_sample, "metric": mt.value,
                              "t_metric_s": tm.value, "t_gradient_s": tg.value, "omp_max_threads": orc.lib.orc_omp_max_threads()}


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path (OpenMP restatement of its functors; the reference
    itself needs Kokkos/Trilinos/STK/MPI and cannot be built here) on this box's host cores.  Loads oracle/ only —
    percept_b200's CUDA library is not mapped into this process."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.time()
    orc = Oracle(omp=True)
    ncores = os.cpu_count() or 1
    n = args.nele
    nsample = min(n, 128)
    full = None
    if n > nsample:  # the full workload once (untimed steps below use the sample): shows the sample is representative
        cf = oracle_hex8_ctx(orc, n, 0.25, ncores)[0]
        v_full, inf = cpu_reference(orc, n, 0.0, ncores, ctx=cf)
        full = {"Melem_per_s": v_full / 1e6, "evaluations": inf["reps"], "seconds": inf["seconds"], "metric": inf["metric"]}
        cf.close()
    cs = oracle_hex8_ctx(orc, nsample, 0.25, ncores)[0]
    mt, ni, tm, tg = C.c_double(), C.c_size_t(), C.c_double(), C.c_double()
    run = lambda reps: orc.lib.orc_bench_metric_and_gradient(cs.p, 1, reps, C.byref(mt), C.byref(ni), C.byref(tm), C.byref(tg))
    evals_per_step = (n // nsample) ** 3  # one step = the whole workload's element count, evaluated on the sample mesh
    # bound the run: at most ~120 s of timed work
    t1 = run(1)
    budget = 120.0
    steps = max(1, min(args.steps, int(budget / max(t1 * evals_per_step, 1e-6))))
    for _ in range(min(args.warmup, 2)):
        run(1)
    tt = 0.0
    for _ in range(steps):
        tt += run(evals_per_step)
    v = nsample ** 3 * evals_per_step * steps / tt
    out = {
        "impl": "reference", "metric": "Melem metric+grad evals/s", "value": v / 1e6, "unit": "Melem/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": min(args.warmup, 2), "ms_per_step": 1e3 * tt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{n}^3 unstructured hex8, random perturbation 0.125h (valid), ShapeB1 fused metric+gradient; one step = "
                               f"{evals_per_step} evaluations of a {nsample}^3-hex sample (= the workload's {n ** 3} elements) on the host cores",
                   "steps_requested": args.steps, "steps_run": steps, "bounded_to_s": budget},
        "cpu_baseline": {"value": v / 1e6, "unit": "Melem/s", "cores": ncores, "kind": "port",
                         "sample": f"{nsample}^3 hex8 x {evals_per_step} evaluations per step x {steps} steps, OpenMP restatement of the "
                                   f"reference functors (oracle/liborc_omp.so; the reference itself needs Kokkos/Trilinos/STK and cannot "
                                   f"be built here)", "full_workload_once": full},
        "e2e": {"value": v / 1e6, "unit": "Melem/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.time() - t0,
    }
    print(json.dumps(out), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------
class Env:
    pass


def nccl_uid(E):
    import torch
    import percept_b200 as pb
    uid = torch.zeros(128, dtype=torch.uint8)
    if E.rank == 0:
        buf = C.create_string_buffer(128)
        assert pb.lib.pms_comm_unique_id(buf) == 0
        uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    uid = uid.cuda()
    E.dist.broadcast(uid, 0)
    return bytes(uid.cpu().numpy().tobytes())


def barrier(E, ctx=None):
    import torch
    if ctx is not None:
        ctx.synchronize()
    torch.cuda.synchronize()
    if E.dist:
        E.dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(E, v):
    import torch
    t = torch.tensor([v], dtype=torch.float64, device="cuda")
    if E.dist:
        E.dist.all_reduce(t, op=E.dist.ReduceOp.MAX)
    return float(t.item())


def timed_evals(E, ctx, call, steps, families=("gradient",)):
    """K calls bracketed by barrier + synchronize, CUDA events on the ctx stream, max over ranks.  Returns
    (ms per step, {family: kernel ms per step on this rank}, launches per step)."""
    import torch
    stream = torch.cuda.ExternalStream(ctx.stream())
    l0 = ctx.launch_count()
    ctx.set_timing(1)
    ctx.kernel_time_reset()
    barrier(E, ctx)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(steps):
            call()
        e1.record(stream)
    barrier(E, ctx)
    ms = max_over_ranks(E, e0.elapsed_time(e1)) / steps
    fam = {f: ctx.kernel_time_ms(f)[0] / steps for f in families}
    ctx.set_timing(0)
    return ms, fam, (ctx.launch_count() - l0) / steps


def cg_iterations(E, ctx, stage, n_it=10):
    """run_one_iteration (Def.hpp:1010-1093, incl. line search) n_it times after the set-up iteration; wall clock
    around barriers, max over ranks."""
    import percept_b200 as pb
    try:
        st = ctx.state_init()
        st.stage = stage
        st.num_invalid = ctx.count_invalid_elements()
        ninv0 = int(st.num_invalid)
        st.untangled = 1 if ninv0 == 0 else 0  # run_algorithm: m_untangled = (m_num_invalid == 0), ReferenceMeshSmootherBase.cpp:262
        gm = ctx.run_one_iteration(st, 0.0)  # iteration 0 (edge lengths, first direction) is set-up: untimed
        ev0 = int(st.n_metric_evals)
        ctx.set_timing(1)
        ctx.kernel_time_reset()
        barrier(E, ctx)
        t0 = time.perf_counter()
        for it in range(1, n_it + 1):
            st.iter = it
            st.num_invalid = ctx.count_invalid_elements()
            if not st.untangled and st.num_invalid == 0:
                st.untangled = 1  # ReferenceMeshSmootherBase.cpp:300
            gm = ctx.run_one_iteration(st, 0.0)
        barrier(E, ctx)
        dt = max_over_ranks(E, time.perf_counter() - t0)
        halo_ms = ctx.kernel_time_ms("halo")[0] / n_it
        ctx.set_timing(0)
        return {"cg_iters_per_s": n_it / dt, "iters": n_it, "metric_evals": int(st.n_metric_evals) - ev0,
                "gradient_evals": int(st.n_gradient_evals), "final_metric": gm, "stage": "ShapeB1" if stage else "Untangle",
                "n_invalid_start": ninv0, "n_invalid_end": int(ctx.count_invalid_elements()), "exchange_ms_per_iter": halo_ms}
    except pb.SmootherError as ex:  # report, do not hide
        return {"error": str(ex)}


def roofline_dict(alg_bytes, kernel_ms, kernel, traffic=None, traffic_source=None, extra=None):
    peak, peak_src = hbm_peak()
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    d = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
         "traffic_source": traffic_source, "peak_source": peak_src, "kernel": kernel, "kernel_ms_per_eval": kernel_ms,
         "algorithmic_bytes_per_eval": alg_bytes}
    if extra:
        d.update(extra)
    return d


def profile_traffic(key):
    """DRAM bytes per evaluation of the dominant kernel from this round's `ncu --set full` capture (profiles/)."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            try:
                v = json.load(open(p)).get(key)
                if v:
                    return v, f"profiles/{name} (ncu capture of the N=1 workload; not re-measured in this run)"
            except Exception:
                pass
    return None, None


def tet_traffic(E):
    """DRAM bytes of the tet4 evaluation from the ncu capture: only meaningful for the single-GPU mesh it was taken on."""
    if E.world != 1:
        return None, None
    return profile_traffic("tet4_mgrad_dram_bytes_per_eval")


# ---- headline: BASELINE config 3 ------------------------------------------------------------------------------------
def bench_hex8(E, args):
    import numpy as np
    import torch
    import percept_b200 as pb
    from percept_b200 import fixtures as fx
    from percept_b200 import partition as part
    n, world, rank = args.nele, E.world, E.rank
    px, py, pz = proc_grid(world)
    NX, NY, NZ = n * px, n * py, n * pz
    h = 1.0 / n  # every rank's box is a unit cube in its own frame: identical per-GPU work (weak scaling)
    t_setup = time.time()
    lm = part.box_local_mesh((NX, NY, NZ), (px, py, pz), rank)
    if world == 1:
        W = fx.cube(n + 1)
        X = fx.perturb_rand(W, n + 1, 0.25 / n, 1)   # BASELINE config 3 generator (build_perturbed_cube stream), valid mesh
        XT = fx.perturb_rand(W, n + 1, 1.0 / n, 1)   # P(1/n): tangled (the untangle stage's input)
    else:
        W = lm.global_lattice_coords(h)
        X = part.perturb_by_global_id(W, lm.node_gid, lm.node_on_global_boundary, 0.25 * h, seed=1)
        XT = part.perturb_by_global_id(W, lm.node_gid, lm.node_on_global_boundary, 1.0 * h, seed=1)
    ctx = pb.create(E.local_rank)
    ctx.set_unstructured(pb.HEX8, lm.nnodes, lm.conn, lm.elem_owned if world > 1 else None, lm.node_owned if world > 1 else None)
    ctx.set_fixed_mask(lm.node_on_global_boundary.astype(np.uint8))
    ctx.commit()
    if world > 1:
        ctx.comm_init(nccl_uid(E), rank, world)
        ctx.set_halo(lm.nb_ranks, lm.send_off, lm.send_nodes, lm.recv_off, lm.recv_nodes)
    for f, v in (("coordinates", X), ("coordinates_N", X), ("coordinates_NM1", W)):
        ctx.upload(f, v)
    nel_owned, nnod = int(lm.n_owned_elems), lm.nnodes
    t_setup = time.time() - t_setup
    E.log(f"hex8 setup done ({t_setup:.1f}s): {nel_owned} owned elements, {nnod} local nodes")
    STAGE = 1

    # ---- (1) device-resident: K fused metric+gradient evaluations ---------------------------------------------
    t_w = time.time()
    for _ in range(args.warmup):
        ctx.metric_and_gradient(STAGE)
    ctx.synchronize()
    per_call = max((time.time() - t_w) / args.warmup, 1e-4)
    # ~0.5 s more of identical load before timing so the clocks are settled and sampled; the count is decided on rank 0
    # and broadcast (every call contains an all-reduce, so all ranks must issue the same number of calls)
    n_spin = torch.tensor([int(min(2000, 0.5 / per_call))], dtype=torch.int64, device="cuda")
    if E.dist:
        E.dist.broadcast(n_spin, 0)
    for _ in range(int(n_spin.item())):
        ctx.metric_and_gradient(STAGE)
    if world > 1:
        ctx.halo_exchange("cg_s")  # first ncclSend/Recv between a pair sets up the P2P channel (~0.7 s, once)
    res = {}
    ms_step, fam, launches = timed_evals(E, ctx, lambda: res.__setitem__("m", ctx.metric_and_gradient(STAGE)), args.steps)
    mtot, ninv = res["m"]
    E.log(f"hex8 device-resident loop done: {ms_step:.3f} ms/step")
    total_elems = nel_owned * world
    value = total_elems / (ms_step * 1e-3) / 1e6  # Melem/s, whole job

    # ---- (2) end to end through the C ABI with pinned host buffers ---------------------------------------------
    hx = torch.from_numpy(np.ascontiguousarray(X)).pin_memory()
    hw = torch.from_numpy(np.ascontiguousarray(W)).pin_memory()
    hg = torch.empty((3, nnod), dtype=torch.float64).pin_memory()
    hx_np, hw_np, hg_np = hx.numpy(), hw.numpy(), hg.numpy()
    e2e_steps = max(3, min(args.steps, 10))

    def e2e_step():
        ctx.upload("coordinates", hx_np)                # H2D 24 B/node
        m, ni = ctx.metric_and_gradient(STAGE)          # kernels + D2H of {mtot, n_invalid}
        ctx.download("cg_g", out=hg_np)                 # D2H 24 B/node
        return m
    for _ in range(2):
        e2e_step()
    barrier(E, ctx)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier(E, ctx)
    e2e_seq = total_elems * e2e_steps / max_over_ranks(E, time.perf_counter() - t0) / 1e6
    E.log(f"e2e (one evaluation at a time) done: {e2e_seq:.0f} Melem/s")

    # pipelined over consecutive evaluations: every step still uploads its coordinates and downloads its gradient, but
    # the read-back of evaluation k (second stream) and the upload of evaluation k+2 (third stream) run beside the
    # kernels of evaluation k+1 (device staging buffers on both sides, see pms_field_*_async)
    hg2 = torch.empty((3, nnod), dtype=torch.float64).pin_memory()
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
    barrier(E, ctx)
    t0 = time.perf_counter()
    pipelined(e2e_steps)                                 # e2e_steps+1 uploads+evaluations, e2e_steps read-backs timed
    barrier(E, ctx)
    e2e_pipe = total_elems * e2e_steps / max_over_ranks(E, time.perf_counter() - t0) / 1e6
    assert np.array_equal(hbuf[0], hbuf[1])              # same coordinates every step: identical gradients came back
    E.log(f"e2e (transfers beside the neighbouring evaluations' kernels) done: {e2e_pipe:.0f} Melem/s")

    # the drop-in's real unit of work: run() from host buffers — upload the three coordinate fields, K CG iterations of
    # ShapeB1 on the device, download the smoothed coordinates
    run_it = 10
    barrier(E, ctx)
    t0 = time.perf_counter()
    ctx.upload("coordinates_NM1", hw_np)
    ctx.upload("coordinates_N", hx_np)
    ctx.upload("coordinates", hx_np)
    st_run = ctx.run_algorithm(run_it, 0.0)
    ctx.download("coordinates", out=hg_np)
    barrier(E, ctx)
    dt_run = max_over_ranks(E, time.perf_counter() - t0)
    n_iter_run = int(st_run.n_gradient_evals)  # one gradient evaluation per CG iteration (stage 0 stops at once on a valid mesh)
    e2e_run = {"seconds": dt_run, "cg_iterations": n_iter_run, "cg_iters_per_s": n_iter_run / dt_run, "h2d_bytes": 72 * nnod,
               "d2h_bytes": 24 * nnod, "final_metric": st_run.global_metric,
               "api": "pms_field_upload x3 (host) ; pms_run_algorithm(inner_iterations = 10 per stage) ; pms_field_download(coordinates)"}
    E.log(f"e2e run(): {dt_run:.3f} s for {run_it} CG iterations from host buffers")

    # ---- (3) CG iterations/s (device resident; the reference's run_one_iteration incl. line search) ------------
    ctx.upload("coordinates", X)
    cg = cg_iterations(E, ctx, 1)
    E.log(f"hex8 ShapeB1 cg done: {cg}")

    # ---- (4) untangle stage on the tangled mesh (BASELINE config 3 as stated: Untangle, then ShapeB1) ----------
    ctx.upload("coordinates", XT)
    ctx.upload("coordinates_N", XT)
    for _ in range(3):
        ctx.metric_and_gradient(0)
    r0 = {}
    ms_u, fam_u, _ = timed_evals(E, ctx, lambda: r0.__setitem__("m", ctx.metric_and_gradient(0)), max(5, min(args.steps, 20)))
    cg_u = cg_iterations(E, ctx, 0)
    untangle = {"mesh": "perturbation +-0.5h (P(1/n), HeavyTestMeshSmoother.cpp:124-159)", "n_invalid": int(r0["m"][1]),
                "invalid_fraction": int(r0["m"][1]) / float(total_elems), "metric_and_gradient_ms": ms_u,
                "Melem_per_s": total_elems / (ms_u * 1e-3) / 1e6, "kernel_ms_per_eval": fam_u["gradient"], "cg": cg_u}
    E.log(f"hex8 untangle done: {ms_u:.3f} ms/eval, cg {cg_u}")

    alg_bytes = ALG_BYTES_PER_NODE_MGRAD * nnod + ALG_BYTES_PER_HEX8_CONN * lm.nelems
    traffic, tsrc = profile_traffic("mgrad_dram_bytes_per_eval")
    out = {"value": value, "ms_per_step": ms_step, "launches_per_step": launches, "nel_owned": nel_owned, "nnod": nnod,
           "nelems_local": lm.nelems, "alg_bytes": alg_bytes, "kernel_ms": fam["gradient"], "traffic": traffic, "traffic_source": tsrc,
           "e2e_seq": e2e_seq, "e2e_pipe": e2e_pipe, "e2e_steps": e2e_steps, "e2e_run": e2e_run, "cg": cg, "untangle": untangle,
           "check": {"total_metric": mtot, "n_invalid": int(ninv)}, "setup_s": t_setup, "cells": [NX, NY, NZ], "procs": [px, py, pz]}
    ctx.close()
    return out


# ---- BASELINE configs 2 and 5: StructuredGrid blocks ----------------------------------------------------------------
def device_field_view(ctx, name, nnod):
    """torch view (3, nnod) of a ctx field's device memory (pms_field_device_ptr: component c at ptr + c*stride)."""
    import torch

    class _Arr:
        pass
    ptr, stride = ctx.field_device_ptr(name)
    a = _Arr()
    a.__cuda_array_interface__ = {"shape": (3, stride), "typestr": "<f8", "data": (ptr, False), "version": 3}
    return torch.as_tensor(a, device="cuda")[:, :nnod]


def bench_sgrid(E, args, n, label, block_per_gpu):
    """One n^3-cell StructuredGrid block per GPU (world > 1: 2x2x2-style layout, interface-node sums over NCCL inside
    get_gradient), sinusoidal interior perturbation (BASELINE config 2's generator on the global cube)."""
    import numpy as np
    import torch
    import percept_b200 as pb
    from percept_b200 import partition as part
    world, rank = (E.world, E.rank) if block_per_gpu else (1, 0)
    procs = proc_grid(world)
    cells = tuple(n * p for p in procs)
    h = 1.0 / n
    t_setup = time.time()
    lb = part.sgrid_block(cells, procs, rank)
    ctx = pb.create(E.local_rank)
    ctx.add_structured_block(lb.node_size)
    ctx.set_fixed_mask(lb.fixed)
    ctx.commit()
    if world > 1:
        ctx.comm_init(nccl_uid(E), rank, world)
        ctx.set_interface(lb.nb_ranks, lb.if_off, lb.if_nodes, lb.owned)
    nel, nnod = ctx.num_elements_local, ctx.num_nodes_local
    # coordinates are generated on the device (torch is plumbing) straight into the context's fields
    dev = torch.device("cuda", E.local_rank)
    ni, nj, nk = lb.node_size
    b = part._box_of_rank(rank, procs)
    g1 = [torch.arange(lb.node_size[d], device=dev, dtype=torch.float64) + b[d] * n for d in range(3)]
    Wx = (g1[0] * h).view(1, 1, ni).expand(nk, nj, ni)
    Wy = (g1[1] * h).view(1, nj, 1).expand(nk, nj, ni)
    Wz = (g1[2] * h).view(nk, 1, 1).expand(nk, nj, ni)
    L = [c * h for c in cells]
    bump = 0.2 * h * (torch.sin(2 * np.pi * 3 * g1[0] * h / L[0]).view(1, 1, ni) * torch.sin(2 * np.pi * 5 * g1[1] * h / L[1]).view(1, nj, 1) *
                      torch.sin(2 * np.pi * 7 * g1[2] * h / L[2]).view(nk, 1, 1))
    free = torch.from_numpy(lb.fixed == 0).to(dev).view(nk, nj, ni)
    bump = bump * free
    for name in ("coordinates_NM1", "coordinates_N", "coordinates"):
        v = device_field_view(ctx, name, nnod)
        for c, Wc in enumerate((Wx, Wy, Wz)):
            v[c].view(nk, nj, ni).copy_(Wc if name == "coordinates_NM1" else Wc + bump)
        torch.cuda.synchronize()
        ctx.field_release_ptr(name)
    del bump, free, Wx, Wy, Wz
    torch.cuda.empty_cache()
    t_setup = time.time() - t_setup
    E.log(f"{label} setup done ({t_setup:.1f}s): {nel} cells, {nnod} nodes per GPU")
    for _ in range(5):
        ctx.metric_and_gradient(1)
    res = {}
    steps = max(5, min(args.steps, 20))
    ms_step, fam, launches = timed_evals(E, ctx, lambda: res.__setitem__("m", ctx.metric_and_gradient(1)), steps, ("gradient", "halo"))
    mtot, ninv = res["m"]
    # node count (lowest-block-id ownership, BlockStructuredGrid.cpp:471-516) against the closed form
    nn_glob = ctx.get_number_nodes()
    nn_expect = (cells[0] + 1) * (cells[1] + 1) * (cells[2] + 1)
    # metric-only evaluation at a trial step length (the line search's pass)
    s = device_field_view(ctx, "cg_s", nnod)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    s.copy_(1e-3 * h * torch.randn((3, nnod), device=dev, dtype=torch.float64, generator=gen) * torch.from_numpy(lb.fixed == 0).to(dev)[None, :])
    torch.cuda.synchronize()
    ctx.field_release_ptr("cg_s")
    ctx.set_option("cache_metric", 0)
    ms_metric, fam_m, _ = timed_evals(E, ctx, lambda: ctx.total_metric(1, 0.5), 5, ("metric",))
    ctx.set_option("cache_metric", 1)
    cg = cg_iterations(E, ctx, 1)
    E.log(f"{label}: {ms_step:.3f} ms/eval, exchange {fam['halo']:.3f} ms, cg {cg}")
    alg_bytes = ALG_BYTES_PER_NODE_MGRAD * nnod
    traffic, tsrc = profile_traffic("sgrid_mgrad_dram_bytes_per_node")
    if traffic:
        traffic = traffic * nnod
    out = {"workload": f"one {n}^3-cell StructuredGrid block per GPU ({nnod} nodes), sinusoidal interior perturbation 0.2h, faces on the "
                       f"global boundary fixed, ShapeB1 fused metric+gradient (one-pass tile kernel)",
           "global_cells": list(cells), "partition": f"{procs[0]}x{procs[1]}x{procs[2]} blocks" +
           (", interface-node gradient sums over ncclSend/Recv inside the step" if world > 1 else ""),
           "n_gpus": world, "value": nel * world / (ms_step * 1e-3) / 1e6, "unit": "Melem/s", "ms_per_step": ms_step, "steps": steps,
           "kernel_ms_per_eval": fam["gradient"], "exchange_ms": fam["halo"], "exchange_share": fam["halo"] / ms_step,
           "launches_per_step": launches,
           "roofline": roofline_dict(alg_bytes, fam["gradient"], "k_sgrid_tile_ws<ShapeB1> + k_sgrid_seam_sum", traffic, tsrc),
           "metric_only": {"ms_per_eval_with_direction": ms_metric, "Melem_per_s": nel * world / (ms_metric * 1e-3) / 1e6},
           "cg": cg, "check": {"total_metric": mtot, "n_invalid": int(ninv), "num_nodes": int(nn_glob), "num_nodes_expected": nn_expect,
                               "num_nodes_equal": int(nn_glob) == nn_expect}, "setup_s": t_setup}
    ctx.close()
    return out


# ---- BASELINE config 4: tet4, RCB ---------------------------------------------------------------------------------
def bench_tet4(E, args, n=119):
    import numpy as np
    import percept_b200 as pb
    from percept_b200 import fixtures as fx
    from percept_b200 import partition as part
    world, rank = E.world, E.rank
    t_setup = time.time()
    nn = n + 1
    conn = fx.tet4_from_hex8(fx.hex8_connectivity(n))
    W = fx.cube(nn)
    fixed = fx.cube_boundary_mask(nn)
    X = part.perturb_by_global_id(W, np.arange(nn ** 3), fixed, 0.15 / n, seed=4)
    nel_glob = conn.shape[0]
    ctx = pb.create(E.local_rank)
    if world > 1:
        conn64 = conn.astype(np.int64)
        er = part.rcb(W.T[conn64].mean(axis=1), world)  # deterministic RCB on element centroids
        lm = part.local_mesh(conn64, er, rank, world, fixed)
        ctx.set_unstructured(pb.TET4, lm.nnodes, lm.conn, lm.elem_owned, lm.node_owned)
        ctx.set_fixed_mask(lm.node_on_global_boundary)
        ctx.commit()
        ctx.comm_init(nccl_uid(E), rank, world)
        ctx.set_halo(lm.nb_ranks, lm.send_off, lm.send_nodes, lm.recv_off, lm.recv_nodes)
        Xl, Wl = np.ascontiguousarray(X[:, lm.node_gid]), np.ascontiguousarray(W[:, lm.node_gid])
        nnod, nel_local, nel_owned = lm.nnodes, lm.nelems, int(lm.n_owned_elems)
        part_sizes = np.bincount(er, minlength=world).tolist()
    else:
        ctx.set_unstructured(pb.TET4, nn ** 3, conn)
        ctx.set_fixed_mask(fixed)
        ctx.commit()
        Xl, Wl = X, W
        nnod, nel_local, nel_owned = nn ** 3, nel_glob, nel_glob
        part_sizes = [nel_glob]
    for f, v in (("coordinates", Xl), ("coordinates_N", Xl), ("coordinates_NM1", Wl)):
        ctx.upload(f, v)
    t_setup = time.time() - t_setup
    E.log(f"C4 tet4 setup done ({t_setup:.1f}s): {nel_owned} owned of {nel_local} local tets, {nnod} local nodes")
    for _ in range(5):
        ctx.metric_and_gradient(1)
    if world > 1:
        ctx.halo_exchange("cg_s")
    res = {}
    steps = max(5, min(args.steps, 50))
    ms_step, fam, launches = timed_evals(E, ctx, lambda: res.__setitem__("m", ctx.metric_and_gradient(1)), steps)
    mtot, ninv = res["m"]
    cg = cg_iterations(E, ctx, 1)
    E.log(f"C4 tet4: {ms_step:.3f} ms/eval, cg {cg}")
    alg_bytes = ALG_BYTES_PER_NODE_MGRAD * nnod + ALG_BYTES_PER_TET4_CONN * nel_local
    out = {"workload": f"{n}^3 hexes x 6 Kuhn tets = {nel_glob} tet4 ({nn ** 3} nodes), perturbation 0.15/{n}, ShapeB1 fused "
                       f"metric+gradient; strong scaling (the one mesh is cut over the ranks)",
           "partition": f"deterministic RCB on element centroids into {world} parts {part_sizes}, one aura layer, owner-computes gradient, "
                        f"cg_s ghost exchange per CG iteration" if world > 1 else "single GPU",
           "n_gpus": world, "scaling": "strong", "value": nel_glob / (ms_step * 1e-3) / 1e6, "unit": "Melem/s", "ms_per_step": ms_step,
           "steps": steps, "kernel_ms_per_eval": fam["gradient"], "launches_per_step": launches,
           "roofline": roofline_dict(alg_bytes, fam["gradient"], "k_elem_pipe<tet4,ShapeB1,grad> + k_grad_gather_csr (this rank's part)",
                                     *tet_traffic(E)),
           "cg": cg, "check": {"total_metric": mtot, "n_invalid": int(ninv)}, "setup_s": t_setup}
    ctx.close()
    return out


# ---- parity against the CPU oracle, printed in the JSON line -------------------------------------------------------
def rel_err(a, b):
    import numpy as np
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def check_single_gpu(E, orc, octx, n):
    """GPU vs oracle on the n^3 hex8 mesh the cpu_baseline leg evaluates: same generator, same connectivity."""
    import percept_b200 as pb
    import numpy as np
    nn = n + 1
    W = orc.cube(nn)
    X = orc.perturb(W, n, 0.25 / n, 1)
    g = pb.create(E.local_rank)
    g.set_unstructured(pb.HEX8, nn ** 3, hex8_conn_np(n))
    g.set_fixed_mask(cube_face_mask_np(nn))
    g.commit()
    for f, v in (("coordinates", X), ("coordinates_N", X), ("coordinates_NM1", W)):
        g.upload(f, v)
    out = {"mesh": f"{n}^3 hex8, perturbation 0.125h"}
    for stage, name in ((1, "shapeb1"), (0, "untangle")):
        mg, ng = g.metric_and_gradient(stage)
        mo, no = octx.metric_and_gradient(stage)
        out[name] = {"metric_rel_err": abs(mg - mo) / max(abs(mo), 1e-300), "n_invalid_equal": int(ng) == int(no),
                     "grad_rel_err": rel_err(g.download("cg_g"), octx.download("cg_g"))}
    # the structured path on the same coordinates (the sgrid tile kernel) against the same oracle gradient
    s = pb.create(E.local_rank)
    b = s.add_structured_block((nn, nn, nn))
    s.add_fixture_1_bcs(b)
    s.select_boundary_sgrid()
    s.commit()
    for f, v in (("coordinates", X), ("coordinates_N", X), ("coordinates_NM1", W)):
        s.upload(f, v)
    ms, ns = s.metric_and_gradient(1)
    mo, no = octx.metric_and_gradient(1)
    out["sgrid_tile_kernel_shapeb1"] = {"metric_rel_err": abs(ms - mo) / abs(mo), "n_invalid_equal": int(ns) == int(no),
                                        "grad_rel_err": rel_err(s.download("cg_g"), octx.download("cg_g"))}
    out["ok"] = all(v["metric_rel_err"] <= 1e-12 and v["grad_rel_err"] <= 1e-12 and v["n_invalid_equal"]
                    for k, v in out.items() if isinstance(v, dict))
    g.close()
    s.close()
    return out


def check_multi_gpu(E, n=40, niter=3):
    """N > 1: a (n px) x (n py) x (n pz) hex8 cube partitioned over the ranks (ghost layer, halo exchange) against the
    single-block oracle on rank 0: all-reduced metric, invalid count, node count, assembled gradient, coordinates
    after `niter` CG iterations per stage, and ghost copies == owner values."""
    import numpy as np
    import torch
    import percept_b200 as pb
    from percept_b200 import partition as part
    world, rank, dist = E.world, E.rank, E.dist
    procs = proc_grid(world)
    cells = tuple(n * p for p in procs)
    h = 1.0 / n
    lm = part.box_local_mesh(cells, procs, rank)
    W = lm.global_lattice_coords(h)
    X = part.perturb_by_global_id(W, lm.node_gid, lm.node_on_global_boundary, 0.3 * h, seed=3)
    c = pb.create(E.local_rank)
    c.set_unstructured(pb.HEX8, lm.nnodes, lm.conn, lm.elem_owned, lm.node_owned)
    c.set_fixed_mask(lm.node_on_global_boundary)
    c.commit()
    c.comm_init(nccl_uid(E), rank, world)
    c.set_halo(lm.nb_ranks, lm.send_off, lm.send_nodes, lm.recv_off, lm.recv_nodes)
    for f, v in (("coordinates", X), ("coordinates_N", X), ("coordinates_NM1", W)):
        c.upload(f, v)
    m1, n1 = c.metric_and_gradient(1)
    g_local = c.download("cg_g")
    st = c.run_algorithm(niter, 0.0)
    x_local = c.download("coordinates")
    ngl = (cells[0] + 1) * (cells[1] + 1) * (cells[2] + 1)

    def assemble(f):
        full = torch.zeros((3, ngl), dtype=torch.float64, device="cuda")
        o = lm.node_owned == 1
        full[:, torch.from_numpy(lm.node_gid[o]).cuda()] = torch.from_numpy(np.ascontiguousarray(f[:, o])).cuda()
        dist.all_reduce(full)
        return full.cpu().numpy()
    xg, gg = assemble(x_local), assemble(g_local)
    ghost_ok = torch.tensor([1.0 if np.array_equal(xg[:, lm.node_gid], x_local) else 0.0], dtype=torch.float64, device="cuda")
    dist.all_reduce(ghost_ok, op=dist.ReduceOp.MIN)
    out = None
    if rank == 0:
        orc = Oracle(omp=True)
        one = part.box_local_mesh(cells, (1, 1, 1), 0)
        W1 = one.global_lattice_coords(h)
        X1 = part.perturb_by_global_id(W1, one.node_gid, one.node_on_global_boundary, 0.3 * h, seed=3)
        o = orc.api.create()
        o.set_option("omp_threads", os.cpu_count() or 1)
        o.set_option("real_is_double", 1)
        o.set_unstructured(8, one.nnodes, one.conn)
        o.set_fixed_mask(one.node_on_global_boundary)
        o.commit()
        for f, v in (("coordinates", X1), ("coordinates_N", X1), ("coordinates_NM1", W1)):
            o.upload(f, v)
        mo, no = o.metric_and_gradient(1)
        go = o.download("cg_g")
        so = o.run_algorithm(niter, 0.0)
        xo = o.download("coordinates")
        out = {"mesh": f"{cells[0]}x{cells[1]}x{cells[2]} hex8 over {world} ranks vs the single-block CPU oracle, {niter} CG iterations per stage",
               "metric_rel_err": abs(m1 - mo) / abs(mo), "n_invalid_equal": int(n1) == int(no), "grad_rel_err": rel_err(gg, go),
               "coord_rel_err_after_cg": rel_err(xg, xo), "final_metric_rel_err": abs(st.global_metric - so.global_metric) / abs(so.global_metric),
               "num_nodes": int(st.num_nodes), "num_nodes_expected": ngl, "ghost_copies_equal_owners": bool(ghost_ok.item() == 1.0)}
        out["ok"] = bool(out["metric_rel_err"] <= 1e-12 and out["grad_rel_err"] <= 1e-12 and out["n_invalid_equal"] and
                         out["coord_rel_err_after_cg"] <= 1e-9 and out["num_nodes"] == ngl and out["ghost_copies_equal_owners"])
        o.close()
    c.close()
    dist.barrier()
    return out


def check_multi_gpu_sgrid(E, n=32, niter=3):
    """N > 1: one structured block per rank (interface sums over NCCL) against the single-block oracle on rank 0."""
    import numpy as np
    import torch
    import percept_b200 as pb
    from percept_b200 import partition as part
    world, rank, dist = E.world, E.rank, E.dist
    procs = proc_grid(world)
    cells = tuple(n * p for p in procs)
    h = 1.0 / n
    lb = part.sgrid_block(cells, procs, rank)
    W = lb.global_lattice_coords(h)
    X = part.perturb_by_global_id(W, lb.node_gid, lb.fixed, 0.3 * h, seed=5)
    c = pb.create(E.local_rank)
    c.add_structured_block(lb.node_size)
    c.set_fixed_mask(lb.fixed)
    c.commit()
    c.comm_init(nccl_uid(E), rank, world)
    c.set_interface(lb.nb_ranks, lb.if_off, lb.if_nodes, lb.owned)
    for f, v in (("coordinates", X), ("coordinates_N", X), ("coordinates_NM1", W)):
        c.upload(f, v)
    m1, n1 = c.metric_and_gradient(1)
    g_local = c.download("cg_g")
    st = c.run_algorithm(niter, 0.0)
    x_local = c.download("coordinates")
    ngl = (cells[0] + 1) * (cells[1] + 1) * (cells[2] + 1)

    def assemble(f):  # owner copies; duplicates must agree (checked below)
        full = torch.zeros((3, ngl), dtype=torch.float64, device="cuda")
        o = lb.owned == 1
        full[:, torch.from_numpy(lb.node_gid[o]).cuda()] = torch.from_numpy(np.ascontiguousarray(f[:, o])).cuda()
        dist.all_reduce(full)
        return full.cpu().numpy()
    xg, gg = assemble(x_local), assemble(g_local)
    dup_ok = torch.tensor([1.0 if (np.array_equal(xg[:, lb.node_gid], x_local) and np.array_equal(gg[:, lb.node_gid], g_local)) else 0.0],
                          dtype=torch.float64, device="cuda")
    dist.all_reduce(dup_ok, op=dist.ReduceOp.MIN)
    out = None
    if rank == 0:
        orc = Oracle(omp=True)
        one = part.sgrid_block(cells, (1, 1, 1), 0)
        W1 = one.global_lattice_coords(h)
        X1 = part.perturb_by_global_id(W1, one.node_gid, one.fixed, 0.3 * h, seed=5)
        o = orc.api.create()
        o.set_option("omp_threads", os.cpu_count() or 1)
        o.set_option("real_is_double", 1)
        o.add_structured_block(one.node_size)
        o.set_fixed_mask(one.fixed)
        o.commit()
        for f, v in (("coordinates", X1), ("coordinates_N", X1), ("coordinates_NM1", W1)):
            o.upload(f, v)
        mo, no = o.metric_and_gradient(1)
        go = o.download("cg_g")
        so = o.run_algorithm(niter, 0.0)
        xo = o.download("coordinates")
        out = {"mesh": f"{world} structured blocks of {n}^3 cells ({procs[0]}x{procs[1]}x{procs[2]}) vs the single-block CPU oracle, {niter} CG iterations per stage",
               "metric_rel_err": abs(m1 - mo) / abs(mo), "n_invalid_equal": int(n1) == int(no), "grad_rel_err": rel_err(gg, go),
               "coord_rel_err_after_cg": rel_err(xg, xo), "num_nodes": int(st.num_nodes), "num_nodes_expected": ngl,
               "interface_copies_equal": bool(dup_ok.item() == 1.0)}
        out["ok"] = bool(out["metric_rel_err"] <= 1e-12 and out["grad_rel_err"] <= 1e-12 and out["n_invalid_equal"] and
                         out["coord_rel_err_after_cg"] <= 1e-9 and out["num_nodes"] == ngl and out["interface_copies_equal"])
        o.close()
    c.close()
    dist.barrier()
    return out


def run_ours(args):
    import torch
    E = Env()
    E.rank = int(os.environ.get("RANK", "0"))
    E.world = int(os.environ.get("WORLD_SIZE", "1"))
    E.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    T0 = time.time()

    def log(msg):
        if E.rank == 0:
            print(f"[bench {time.time() - T0:7.1f}s] {msg}", file=sys.stderr, flush=True)
    E.log = log
    if E.world != args.gpus and E.world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={E.world}")
    torch.cuda.set_device(E.local_rank)
    E.dist = None
    if E.world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", E.local_rank))
        E.dist = dist
    world, n = E.world, args.nele
    want = [c for c in args.configs.split(",") if c]

    sampler = ClockSampler(E.local_rank)
    sampler.start()  # samples every 100 ms from warm-up to the end of the timed work (all under load)
    H = bench_hex8(E, args)
    clocks = sampler.stop()

    configs = {}
    errors = {}

    def guarded(key, fn):
        try:
            configs[key] = fn()
        except Exception as ex:  # a failed sub-config is reported in the line, it does not take the headline down
            errors[key] = f"{type(ex).__name__}: {ex}"
            log(f"{key} FAILED: {errors[key]}")
            if E.dist:
                raise
    if "C2" in want and world == 1:
        guarded("C2_sgrid_128", lambda: bench_sgrid(E, args, 128, "C2 sgrid 128^3", False))
    if "C4" in want:
        guarded("C4_tet4_10M", lambda: bench_tet4(E, args))
    if "C5" in want:
        guarded("C5_sgrid_512_per_gpu", lambda: bench_sgrid(E, args, args.nele_c5, f"C5 sgrid {args.nele_c5}^3/GPU", True))

    check = dict(H["check"])
    if world > 1 and "check" in want:
        check["multi_gpu_hex8_vs_oracle"] = check_multi_gpu(E)
        check["multi_gpu_sgrid_vs_oracle"] = check_multi_gpu_sgrid(E)
    if E.rank != 0:
        if E.dist:
            E.dist.destroy_process_group()
        return

    # ---- roofline of the fused metric+gradient evaluation (element pass + deterministic gather) -----------------
    kernel_ms = H["kernel_ms"]
    fp64_instr = FP64_INSTR_PER_HEX_MGRAD * H["nelems_local"]
    fp64_peak = 148 * 64 * (clocks.get("sm_mhz") or 1965) * 1e6
    roofline = roofline_dict(H["alg_bytes"], kernel_ms, "k_elem_pipe<hex8,ShapeB1,grad> + k_grad_gather_csr (one evaluation)",
                             H["traffic"], H["traffic_source"],
                             {"note": f"the evaluation is FP64-pipe-bound ({FP64_INSTR_PER_HEX_MGRAD:.0f} FP64 instr/hex vs 64 lanes/clk/SM), see DESIGN.md",
                              "fp64_pipe": {"achieved_ginstr_s": fp64_instr / (kernel_ms * 1e-3) / 1e9, "peak_ginstr_s": fp64_peak / 1e9,
                                            "frac": fp64_instr / (kernel_ms * 1e-3) / fp64_peak,
                                            "note": f"{FP64_INSTR_PER_HEX_MGRAD:.0f} FP64 thread-instructions per hex is the static SASS count of the element kernel "
                                                    f"(profiles/r02_traffic.json), not re-measured in this run; whole evaluation incl. the HBM-bound gather pass"}})

    # ---- CPU baseline on this box's host cores, bounded sample; the same oracle context then checks the GPU ----------
    cpu_baseline = None
    if world == 1:
        try:
            orc = Oracle(omp=True)
            ns = min(n, 128)
            octx = oracle_hex8_ctx(orc, ns, 0.25, os.cpu_count() or 1)[0]
            cpu_v, info = cpu_reference(orc, ns, 12.0, ctx=octx)
            cpu_baseline = {"value": cpu_v / 1e6, "unit": "Melem/s", "cores": info["cores"], "kind": "port",
                            "sample": f"{info['nele_sample']}^3 hex8 (1/{(n // info['nele_sample']) ** 3} of the workload) x {info['reps']} evaluations, "
                                      f"{info['seconds']:.1f} s; OpenMP restatement of the reference functors (oracle/liborc_omp.so)"}
            if "check" in want:
                octx.set_option("real_is_double", 1)
                check["gpu_vs_oracle"] = check_single_gpu(E, orc, octx, ns)
            octx.close()
        except Exception as ex:
            cpu_baseline = {"value": None, "unit": "Melem/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {type(ex).__name__}: {ex}"}
    nnod = H["nnod"]
    out = {
        "metric": "Melem metric+grad evals/s", "value": H["value"], "unit": "Melem/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": H["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{n}^3 unstructured hex8 per GPU ({H['nel_owned']} owned elements, {nnod} local nodes), unit-cube lattice, "
                               f"random interior perturbation 0.125h (valid mesh), cube faces fixed, ShapeB1 fused metric+gradient",
                   "global_cells": H["cells"], "partition": f"{H['procs'][0]}x{H['procs'][1]}x{H['procs'][2]} boxes, 1 ghost layer, owner-computes",
                   "l2": "inputs (815 MB of coordinates per GPU) exceed the 126 MB L2; no explicit flush"},
        "e2e": {"value": H["e2e_seq"], "unit": "Melem/s", "h2d_bytes_per_step": 24 * nnod, "d2h_bytes_per_step": 24 * nnod + 16,
                "steps": H["e2e_steps"], "api": "per step: pms_field_upload(coordinates, host) ; pms_metric_and_gradient ; pms_field_download(cg_g, host); "
                                                "pinned host buffers, one evaluation at a time",
                "pipelined": {"value": H["e2e_pipe"], "api": "pms_field_download_async(cg_g of evaluation k) || pms_field_upload_async(coordinates of "
                                                             "k+2) || pms_metric_and_gradient(k+1); three streams"},
                "run": H["e2e_run"]},
        "gpu_launches": int(round(H["launches_per_step"] * args.steps)), "roofline": roofline, "clocks": clocks,
        "cg": H["cg"], "untangle": H["untangle"], "configs": configs, "check": check, "setup_s": H["setup_s"],
    }
    if cpu_baseline is not None:
        out["cpu_baseline"] = cpu_baseline
    if errors:
        out["config_errors"] = errors
    print(json.dumps(out), flush=True)
    if E.dist:
        E.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nele", type=int, default=256, help="cells per direction per GPU of the headline hex8 workload")
    ap.add_argument("--nele-c5", dest="nele_c5", type=int, default=512, help="cells per direction of the structured block per GPU (C5)")
    ap.add_argument("--configs", default="C2,C4,C5,check", help="additional BASELINE configurations / oracle checks to run (comma list; '' = none)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
