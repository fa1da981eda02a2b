"""Host-side multi-GPU logic on CPU: deterministic RCB, ownership, aura layer, halo plans.
The 2-rank exchange runs over torch.distributed/gloo (world_size 2) exactly as the NCCL path is driven."""
import os
import subprocess
import sys

import numpy as np
import pytest

from percept_b200 import partition as part
from tests import meshes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def centroids(conn, X):
    return X.T[conn].mean(axis=1)


@pytest.mark.parametrize("nparts", [2, 4, 8])
def test_rcb_is_balanced_deterministic_and_gives_boxes(nparts):
    n = 8
    conn = meshes.hex8_conn(n)
    X, _ = meshes.sinus_coords(n, 0.0)
    c = centroids(conn, X)
    p1, p2 = part.rcb(c, nparts), part.rcb(c.copy(), nparts)
    assert np.array_equal(p1, p2)
    counts = np.bincount(p1, minlength=nparts)
    assert counts.min() == counts.max() == n ** 3 // nparts
    # on a cube RCB cuts boxes: every part's centroid bounding box has volume fraction 1/nparts
    for q in range(nparts):
        ext = c[p1 == q].max(axis=0) - c[p1 == q].min(axis=0) + 1.0 / n
        assert abs(np.prod(ext) - 1.0 / nparts) < 1e-12


@pytest.mark.parametrize("procs", [(2, 1, 1), (2, 2, 1), (2, 2, 2)])
def test_box_local_mesh_equals_general_builder(procs):
    cells = (6, 4, 4)
    px, py, pz = procs
    nr = px * py * pz
    NX, NY, NZ = cells
    # global hex8 lattice
    ci, cj, ck = np.meshgrid(np.arange(NX), np.arange(NY), np.arange(NZ), indexing="ij")
    ci, cj, ck = (a.transpose(2, 1, 0).reshape(-1) for a in (ci, cj, ck))
    nid = lambda i, j, k: i + (NX + 1) * (j + (NY + 1) * k)
    offs = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
    conn = np.stack([nid(ci + a, cj + b, ck + c) for a, b, c in offs], axis=1).astype(np.int64)
    er = part.box_elem_rank(cells, procs)
    for r in range(nr):
        a = part.box_local_mesh(cells, procs, r)
        b = part.local_mesh(conn, er, r, nr)
        assert a.nnodes == b.nnodes and a.nelems == b.nelems and a.n_owned_elems == b.n_owned_elems
        assert np.array_equal(a.node_gid, b.node_gid) and np.array_equal(a.elem_gid, b.elem_gid)
        assert np.array_equal(a.conn, b.conn)
        assert np.array_equal(a.elem_owned, b.elem_owned) and np.array_equal(a.node_owned, b.node_owned)
        assert np.array_equal(a.nb_ranks, b.nb_ranks)
        assert np.array_equal(a.send_off, b.send_off) and np.array_equal(a.send_nodes, b.send_nodes)
        assert np.array_equal(a.recv_off, b.recv_off) and np.array_equal(a.recv_nodes, b.recv_nodes)


def test_every_node_has_exactly_one_owner_and_owned_gradients_are_local():
    n, nr = 6, 4
    conn = meshes.hex8_conn(n).astype(np.int64)
    X, _ = meshes.sinus_coords(n, 0.0)
    er = part.rcb(centroids(conn, X), nr)
    owned_count = np.zeros((n + 1) ** 3, int)
    for r in range(nr):
        lm = part.local_mesh(conn, er, r, nr)
        owned_count[lm.node_gid[lm.node_owned == 1]] += 1
        # every element adjacent to an owned node is present locally (complete owner-computes gradient)
        adj_global = np.bincount(conn.reshape(-1), minlength=(n + 1) ** 3)
        adj_local = np.bincount(lm.conn.reshape(-1), minlength=lm.nnodes)
        o = lm.node_owned == 1
        assert np.array_equal(adj_local[o], adj_global[lm.node_gid[o]])
    assert (owned_count == 1).all()


def test_perturbation_by_global_id_is_rank_independent():
    cells, procs = (4, 4, 4), (2, 1, 1)
    a, b = part.box_local_mesh(cells, procs, 0), part.box_local_mesh(cells, procs, 1)
    Xa = part.perturb_by_global_id(a.global_lattice_coords(0.25), a.node_gid, a.node_on_global_boundary, 0.05)
    Xb = part.perturb_by_global_id(b.global_lattice_coords(0.25), b.node_gid, b.node_on_global_boundary, 0.05)
    common, ia, ib = np.intersect1d(a.node_gid, b.node_gid, return_indices=True)
    assert common.size > 0 and np.array_equal(Xa[:, ia], Xb[:, ib])
    assert np.array_equal(Xa[:, a.node_on_global_boundary == 1], a.global_lattice_coords(0.25)[:, a.node_on_global_boundary == 1])


WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from percept_b200 import partition as part
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
cells, procs = (6, 4, 4), (2, 1, 1)
lm = part.box_local_mesh(cells, procs, rank)
# a global nodal field f(gid); owners hold the truth, ghosts start wrong
truth = lambda g: np.stack([np.sin(g * 0.37), np.cos(g * 0.11), g * 1e-3])
f = truth(lm.node_gid)
f[:, lm.node_owned == 0] = -777.0
# owner -> ghost exchange exactly as pms_halo_exchange does it (pack, send/recv per neighbour, unpack)
reqs, bufs = [], []
for q, nb in enumerate(lm.nb_ranks):
    s = lm.send_nodes[lm.send_off[q]:lm.send_off[q + 1]]
    r = lm.recv_nodes[lm.recv_off[q]:lm.recv_off[q + 1]]
    if s.size:
        reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(f[:, s])), int(nb)))
    if r.size:
        t = torch.empty((3, r.size), dtype=torch.float64)
        bufs.append((r, t))
        reqs.append(dist.irecv(t, int(nb)))
for q in reqs:
    q.wait()
for r, t in bufs:
    f[:, r] = t.numpy()
assert np.array_equal(f, truth(lm.node_gid)), "ghost values differ from owner values"
# scalar reductions: owned-node dot and owned-element count add up to the global values
d = torch.tensor([float((truth(lm.node_gid)[:, lm.node_owned == 1] ** 2).sum()), float(lm.n_owned_elems)], dtype=torch.float64)
dist.all_reduce(d)
NX, NY, NZ = cells
g = np.arange((NX + 1) * (NY + 1) * (NZ + 1))
assert abs(d[0].item() - (truth(g) ** 2).sum()) < 1e-9 and d[1].item() == NX * NY * NZ
dist.destroy_process_group()
print("OK", rank)
'''


def test_two_rank_halo_exchange_over_gloo(tmp_path):
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", str(w), ROOT]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert p.returncode == 0, p.stdout + p.stderr
    assert p.stdout.count("OK") == 2
