"""Multi-GPU parity worker (run under torchrun, one rank per GPU, NCCL): partitioned smoothing vs the same mesh on one GPU."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

import percept_b200 as pb
from percept_b200 import partition as part


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    procs = {2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[world]
    n = 10
    cells = (n * procs[0], n * procs[1], n * procs[2])
    h = 1.0 / n
    niter = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    eps = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
    mode = sys.argv[3] if len(sys.argv) > 3 else "hex8"
    if mode == "sgrid":
        return main_sgrid(rank, world, local, procs, n, cells, h, niter, eps)
    if mode == "tet4":
        return main_tet4(rank, world, local, niter, eps)

    lm = part.box_local_mesh(cells, procs, rank)
    W = lm.global_lattice_coords(h)
    X = part.perturb_by_global_id(W, lm.node_gid, lm.node_on_global_boundary, eps * h, seed=3)
    c = pb.create(local)
    c.set_unstructured(pb.HEX8, lm.nnodes, lm.conn, lm.elem_owned, lm.node_owned)
    c.set_fixed_mask(lm.node_on_global_boundary)
    c.commit()
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = C.create_string_buffer(128)
        assert pb.lib.pms_comm_unique_id(buf) == 0
        uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    uid = uid.cuda()
    dist.broadcast(uid, 0)
    c.comm_init(uid.cpu().numpy().tobytes(), rank, world)
    c.set_halo(lm.nb_ranks, lm.send_off, lm.send_nodes, lm.recv_off, lm.recv_nodes)
    for f, v in (("coordinates", X), ("coordinates_N", X), ("coordinates_NM1", W)):
        c.upload(f, v)
    m1, n1 = c.metric_and_gradient(1)
    g_local = c.download("cg_g")
    st = c.run_algorithm(niter, 0.0)
    x_local = c.download("coordinates")
    ngl = (cells[0] + 1) * (cells[1] + 1) * (cells[2] + 1)

    # assemble the owned-node results on rank 0
    def assemble(f):
        full = torch.zeros((3, ngl), dtype=torch.float64, device="cuda")
        o = lm.node_owned == 1
        full[:, torch.from_numpy(lm.node_gid[o]).cuda()] = torch.from_numpy(np.ascontiguousarray(f[:, o])).cuda()
        dist.all_reduce(full)
        return full.cpu().numpy()
    xg, gg = assemble(x_local), assemble(g_local)

    # every ghost copy must equal its owner's value after the run (consistent duplicates)
    own_vals = torch.from_numpy(xg[:, lm.node_gid]).numpy()
    ghost_ok = np.array_equal(own_vals, x_local)

    if rank == 0:
        one = part.box_local_mesh(cells, (1, 1, 1), 0)
        W1 = one.global_lattice_coords(h)
        X1 = part.perturb_by_global_id(W1, one.node_gid, one.node_on_global_boundary, eps * h, seed=3)
        s = pb.create(local)
        s.set_unstructured(pb.HEX8, one.nnodes, one.conn)
        s.set_fixed_mask(one.node_on_global_boundary)
        s.commit()
        for f, v in (("coordinates", X1), ("coordinates_N", X1), ("coordinates_NM1", W1)):
            s.upload(f, v)
        ms, ns = s.metric_and_gradient(1)
        gs = s.download("cg_g")
        ss = s.run_algorithm(niter, 0.0)
        xs = s.download("coordinates")
        rel = lambda a, b: np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)
        res = {"metric_rel": abs(m1 - ms) / abs(ms), "ninv": (n1, ns), "grad_rel": rel(gg, gs), "coord_rel": rel(xg, xs),
               "final_metric_rel": abs(st.global_metric - ss.global_metric) / abs(ss.global_metric), "ghost_ok": bool(ghost_ok),
               "num_nodes": (int(st.num_nodes), int(ss.num_nodes))}
        print("MGRESULT", res, flush=True)
        assert n1 == ns and res["metric_rel"] < 1e-12 and res["grad_rel"] < 1e-12, res
        assert res["coord_rel"] < 1e-9 and res["final_metric_rel"] < 1e-9, res
        assert st.num_nodes == ss.num_nodes == ngl
    assert ghost_ok, f"rank {rank}: ghost coordinates differ from owner coordinates"
    dist.barrier()
    dist.destroy_process_group()
    print("MGOK", rank, flush=True)


def nccl_uid(rank):
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = C.create_string_buffer(128)
        assert pb.lib.pms_comm_unique_id(buf) == 0
        uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    uid = uid.cuda()
    dist.broadcast(uid, 0)
    return uid.cpu().numpy().tobytes()


def main_tet4(rank, world, local, niter, eps):
    """BASELINE config 4 at test size: tet4 mesh from a subdivided cube, deterministic RCB on element centroids,
    one aura layer, owner->ghost exchange; compared with the same mesh on one GPU."""
    from percept_b200 import fixtures as fx
    n = 12
    nn = n + 1
    conn = fx.tet4_from_hex8(fx.hex8_connectivity(n)).astype(np.int64)
    W = fx.cube(nn)
    fixed = fx.cube_boundary_mask(nn)
    gid_all = np.arange(nn ** 3)
    X = part.perturb_by_global_id(W, gid_all, fixed, eps / n, seed=4)
    er = part.rcb(W.T[conn].mean(axis=1), world)
    lm = part.local_mesh(conn, er, rank, world, fixed)
    c = pb.create(local)
    c.set_unstructured(pb.TET4, lm.nnodes, lm.conn, lm.elem_owned, lm.node_owned)
    c.set_fixed_mask(lm.node_on_global_boundary)
    c.commit()
    c.comm_init(nccl_uid(rank), rank, world)
    c.set_halo(lm.nb_ranks, lm.send_off, lm.send_nodes, lm.recv_off, lm.recv_nodes)
    for f, v in (("coordinates", X[:, lm.node_gid]), ("coordinates_N", X[:, lm.node_gid]), ("coordinates_NM1", W[:, lm.node_gid])):
        c.upload(f, np.ascontiguousarray(v))
    m1, n1 = c.metric_and_gradient(1)
    g_local = c.download("cg_g")
    st = c.run_algorithm(niter, 0.0)
    x_local = c.download("coordinates")
    ngl = nn ** 3

    def assemble(f):
        full = torch.zeros((3, ngl), dtype=torch.float64, device="cuda")
        o = lm.node_owned == 1
        full[:, torch.from_numpy(lm.node_gid[o]).cuda()] = torch.from_numpy(np.ascontiguousarray(f[:, o])).cuda()
        dist.all_reduce(full)
        return full.cpu().numpy()
    xg, gg = assemble(x_local), assemble(g_local)
    ghost_ok = np.array_equal(xg[:, lm.node_gid], x_local)
    if rank == 0:
        s = pb.create(local)
        s.set_unstructured(pb.TET4, ngl, conn.astype(np.int32))
        s.set_fixed_mask(fixed)
        s.commit()
        for f, v in (("coordinates", X), ("coordinates_N", X), ("coordinates_NM1", W)):
            s.upload(f, v)
        ms, ns = s.metric_and_gradient(1)
        gs = s.download("cg_g")
        ss = s.run_algorithm(niter, 0.0)
        xs = s.download("coordinates")
        rel = lambda a, b: np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)
        res = {"metric_rel": abs(m1 - ms) / abs(ms), "ninv": (n1, ns), "grad_rel": rel(gg, gs), "coord_rel": rel(xg, xs),
               "final_metric_rel": abs(st.global_metric - ss.global_metric) / abs(ss.global_metric), "ghost_ok": bool(ghost_ok),
               "parts": np.bincount(er).tolist()}
        print("MGRESULT", res, flush=True)
        assert n1 == ns and res["metric_rel"] < 1e-12 and res["grad_rel"] < 1e-12, res
        assert res["coord_rel"] < 1e-9 and res["final_metric_rel"] < 1e-9, res
        assert st.num_nodes == ss.num_nodes == ngl
    assert ghost_ok
    dist.barrier()
    dist.destroy_process_group()
    print("MGOK", rank, flush=True)


def main_sgrid(rank, world, local, procs, n, cells, h, niter, eps):
    """One structured block per GPU (BASELINE config 5) vs the single block holding the whole lattice."""
    lb = part.sgrid_block(cells, procs, rank)
    W = lb.global_lattice_coords(h)
    X = part.perturb_by_global_id(W, lb.node_gid, lb.fixed, eps * h, seed=3)
    c = pb.create(local)
    c.add_structured_block(lb.node_size)
    c.set_fixed_mask(lb.fixed)
    c.commit()
    c.comm_init(nccl_uid(rank), rank, world)
    c.set_interface(lb.nb_ranks, lb.if_off, lb.if_nodes, lb.owned)
    for f, v in (("coordinates", X), ("coordinates_N", X), ("coordinates_NM1", W)):
        c.upload(f, v)
    m1, n1 = c.metric_and_gradient(1)
    g_local = c.download("cg_g")
    c.get_edge_lengths()
    el_local = c.download("cg_edge_length")
    st = c.run_algorithm(niter, 0.0)
    x_local = c.download("coordinates")
    ngl = (cells[0] + 1) * (cells[1] + 1) * (cells[2] + 1)

    def assemble(f):
        full = torch.zeros((f.shape[0], ngl), dtype=torch.float64, device="cuda")
        o = lb.owned == 1
        full[:, torch.from_numpy(lb.node_gid[o]).cuda()] = torch.from_numpy(np.ascontiguousarray(f[:, o])).cuda()
        dist.all_reduce(full)
        return full.cpu().numpy()
    xg, gg, eg = assemble(x_local), assemble(g_local), assemble(el_local)
    copies_ok = np.array_equal(xg[:, lb.node_gid], x_local) and np.array_equal(gg[:, lb.node_gid], g_local)
    if rank == 0:
        NXn, NYn, NZn = cells[0] + 1, cells[1] + 1, cells[2] + 1
        gid = np.arange(ngl)
        one = part.sgrid_block(cells, (1, 1, 1), 0)
        W1 = one.global_lattice_coords(h)
        X1 = part.perturb_by_global_id(W1, one.node_gid, one.fixed, eps * h, seed=3)
        s = pb.create(local)
        s.add_structured_block((NXn, NYn, NZn))
        s.set_fixed_mask(one.fixed)
        s.commit()
        for f, v in (("coordinates", X1), ("coordinates_N", X1), ("coordinates_NM1", W1)):
            s.upload(f, v)
        ms, ns = s.metric_and_gradient(1)
        gs = s.download("cg_g")
        s.get_edge_lengths()
        es = s.download("cg_edge_length")
        ss = s.run_algorithm(niter, 0.0)
        xs = s.download("coordinates")
        rel = lambda a, b: np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)
        res = {"metric_rel": abs(m1 - ms) / abs(ms), "ninv": (n1, ns), "grad_rel": rel(gg, gs), "edge_rel": rel(eg, es),
               "coord_rel": rel(xg, xs), "final_metric_rel": abs(st.global_metric - ss.global_metric) / abs(ss.global_metric),
               "copies_ok": bool(copies_ok), "num_nodes": (int(st.num_nodes), int(ss.num_nodes))}
        print("MGRESULT", res, flush=True)
        # SGridSmootherMultiBlockvsSingleBlock.cpp tolerances: mtot 1e-10, nodal cg_g 1e-12, edge_len 1e-12
        assert n1 == ns and res["metric_rel"] < 1e-12 and res["grad_rel"] < 1e-12 and res["edge_rel"] < 1e-12, res
        assert res["coord_rel"] < 1e-9 and res["final_metric_rel"] < 1e-9, res
        assert st.num_nodes == ss.num_nodes == ngl
    assert copies_ok, f"rank {rank}: interface copies differ"
    dist.barrier()
    dist.destroy_process_group()
    print("MGOK", rank, flush=True)


if __name__ == "__main__":
    main()
