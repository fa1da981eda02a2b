"""Deterministic mesh partitioning for the multi-GPU path (host side, numpy).

Replaces, for this path only, what Percept gets from Zoltan RCB + STK ownership/aura ghosting
(src/percept/stk_rebalance/ZoltanPartition.cpp:133-150,622; stk owned/shared/aura parts):

  * `rcb(centroids, nparts)`      geometric recursive coordinate bisection, split the longest axis at the median of
                                  (coordinate, element id) — fully deterministic (Zoltan's own result is unpinned by
                                  the reference's tests).
  * `local_mesh(...)`             rank-local view: owned elements + the elements adjacent to owned nodes (one aura
                                  layer), node owner = min rank among the owners of adjacent elements, owned masks,
                                  and the owner->ghost exchange plan (matching send/recv lists ordered by global id).
  * `box_local_mesh(...)`         the same result computed analytically for a lattice cut into px*py*pz boxes
                                  (what RCB produces on a cube), without materialising the global mesh.

With owner-computes + one ghost layer the nodal gradient of owned nodes is complete locally (mirrors
ReferenceMeshSmootherConjugateGradientDef.hpp:1222-1246, 1355-1386), so only cg_s travels, once per CG iteration.
"""
from dataclasses import dataclass, field

import numpy as np


def rcb(centroids, nparts):
    """centroids (nelem, 3) -> part id per element; nparts must be a power of two."""
    n = centroids.shape[0]
    part = np.zeros(n, dtype=np.int32)
    assert nparts & (nparts - 1) == 0, "nparts must be a power of two"

    def split(ids, base, k):
        if k == 1:
            part[ids] = base
            return
        c = centroids[ids]
        ext = c.max(axis=0) - c.min(axis=0)
        ax = int(np.argmax(ext))  # first longest axis
        order = np.lexsort((ids, c[:, ax]))  # by coordinate, ties by element id
        half = len(ids) // 2
        split(ids[order[:half]], base, k // 2)
        split(ids[order[half:]], base + k // 2, k // 2)

    split(np.arange(n), 0, nparts)
    return part


@dataclass
class LocalMesh:
    nnodes: int
    nelems: int
    conn: np.ndarray               # (nelems, npe) local node ids
    elem_owned: np.ndarray         # uint8
    node_owned: np.ndarray         # uint8
    node_gid: np.ndarray           # global node id of each local node
    elem_gid: np.ndarray
    node_on_global_boundary: np.ndarray
    n_owned_elems: int
    nb_ranks: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    send_off: np.ndarray = field(default_factory=lambda: np.zeros(1, np.int64))
    send_nodes: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    recv_off: np.ndarray = field(default_factory=lambda: np.zeros(1, np.int64))
    recv_nodes: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    lattice: tuple = None          # (NX, NY, NZ) cells of the global lattice (box meshes only)

    def global_lattice_coords(self, h):
        """Unperturbed lattice coordinates of the local nodes, spacing h (SoA (3, nnodes))."""
        NX, NY, NZ = self.lattice
        g = self.node_gid
        i = g % (NX + 1)
        j = (g // (NX + 1)) % (NY + 1)
        k = g // ((NX + 1) * (NY + 1))
        return np.ascontiguousarray(np.stack([i * h, j * h, k * h]).astype(np.float64))


def node_owner(conn, elem_rank, nnodes):
    """owner(node) = min rank among the owners of its adjacent elements."""
    owner = np.full(nnodes, np.iinfo(np.int32).max, dtype=np.int32)
    np.minimum.at(owner, conn.reshape(-1), np.repeat(elem_rank.astype(np.int32), conn.shape[1]))
    return owner


def _halo_plan(rank, nranks, local_gids_of, owner_g, g2l):
    """send/recv lists ordered by global node id. local_gids_of(r) -> sorted global ids of rank r's local nodes."""
    mine = local_gids_of(rank)
    own = owner_g[mine]
    recv, send = {}, {}
    for q in np.unique(own):
        if q != rank:
            recv[int(q)] = g2l[mine[own == q]]
    for r in range(nranks):
        if r == rank:
            continue
        theirs = local_gids_of(r)
        sel = theirs[owner_g[theirs] == rank]
        if sel.size:
            send[r] = g2l[sel]
    nb = sorted(set(recv) | set(send))
    so, ro, sn, rn = [0], [0], [], []
    for q in nb:
        s = send.get(q, np.zeros(0, np.int64))
        r_ = recv.get(q, np.zeros(0, np.int64))
        sn.append(s)
        rn.append(r_)
        so.append(so[-1] + len(s))
        ro.append(ro[-1] + len(r_))
    cat = lambda xs: np.concatenate(xs).astype(np.int32) if xs else np.zeros(0, np.int32)
    return (np.array(nb, np.int32), np.array(so, np.int64), cat(sn), np.array(ro, np.int64), cat(rn))


def local_elements(conn, elem_rank, owner, r):
    """Elements a rank holds: its owned elements plus every element touching one of its owned nodes (aura layer)."""
    owned = elem_rank == r
    touches = (owner[conn] == r).any(axis=1)
    return np.nonzero(owned | touches)[0]


def local_mesh(conn, elem_rank, rank, nranks, boundary_mask=None):
    """General (global-view) builder. conn (nelem, npe) global node ids; elem_rank part id per element."""
    nnodes = int(conn.max()) + 1
    owner = node_owner(conn, elem_rank, nnodes)
    cache = {}

    def gids_of(r):
        if r not in cache:
            cache[r] = np.unique(conn[local_elements(conn, elem_rank, owner, r)])
        return cache[r]

    le = local_elements(conn, elem_rank, owner, rank)
    lg = gids_of(rank)  # sorted global ids = local numbering
    g2l = np.full(nnodes, -1, dtype=np.int64)
    g2l[lg] = np.arange(lg.size)
    lconn = g2l[conn[le]].astype(np.int32)
    nb, so, sn, ro, rn = _halo_plan(rank, nranks, gids_of, owner, g2l)
    bm = np.zeros(nnodes, np.uint8) if boundary_mask is None else boundary_mask
    return LocalMesh(nnodes=int(lg.size), nelems=int(le.size), conn=np.ascontiguousarray(lconn),
                     elem_owned=(elem_rank[le] == rank).astype(np.uint8), node_owned=(owner[lg] == rank).astype(np.uint8),
                     node_gid=lg.astype(np.int64), elem_gid=le.astype(np.int64), node_on_global_boundary=bm[lg].astype(np.uint8),
                     n_owned_elems=int((elem_rank[le] == rank).sum()), nb_ranks=nb, send_off=so, send_nodes=sn, recv_off=ro,
                     recv_nodes=rn)


# ------------------------------------------------------------------------------------------------------------------
# analytic box decomposition of an (NX, NY, NZ)-cell hex8 lattice into (px, py, pz) boxes
# ------------------------------------------------------------------------------------------------------------------
def _box_of_rank(rank, procs):
    px, py, pz = procs
    return rank % px, (rank // px) % py, rank // (px * py)


def _ranges(N, p, b):
    """owned cell range [lo, hi), local cell range [lo, lhi), local node range [lo, lnhi] of box b in one direction."""
    n = N // p
    lo, hi = b * n, (b + 1) * n if b < p - 1 else N
    lhi = min(hi + 1, N)  # one ghost cell layer on the high side (the cells that touch owned nodes at index hi)
    return lo, hi, lhi


def _owner_box_1d(idx, N, p):
    n = N // p
    ob = np.where(idx == 0, 0, (idx - 1) // n)
    return np.minimum(ob, p - 1)


def _local_node_box(rank, cells, procs):
    b = _box_of_rank(rank, procs)
    out = []
    for d in range(3):
        lo, hi, lhi = _ranges(cells[d], procs[d], b[d])
        out.append((lo, lhi))  # node range [lo, lhi] inclusive
    return out


def box_local_mesh(cells, procs, rank):
    """Rank-local hex8 mesh (Exodus order) of the lattice cut into boxes; equals local_mesh() with RCB boxes."""
    NX, NY, NZ = cells
    px, py, pz = procs
    nranks = px * py * pz
    b = _box_of_rank(rank, procs)
    r3 = [_ranges(cells[d], procs[d], b[d]) for d in range(3)]
    (lx, hx, lhx), (ly, hy, lhy), (lz, hz, lhz) = r3
    ncx, ncy, ncz = lhx - lx, lhy - ly, lhz - lz         # local cells
    nnx, nny, nnz = ncx + 1, ncy + 1, ncz + 1           # local nodes
    # local nodes, lexicographic (i fastest) == ascending global id within the box
    ii, jj, kk = np.meshgrid(np.arange(nnx), np.arange(nny), np.arange(nnz), indexing="ij")
    ii, jj, kk = (a.transpose(2, 1, 0).reshape(-1) for a in (ii, jj, kk))
    gi, gj, gk = ii + lx, jj + ly, kk + lz
    gid = gi + (NX + 1) * (gj + (NY + 1) * gk.astype(np.int64))
    owner = (_owner_box_1d(gi, NX, px) + px * (_owner_box_1d(gj, NY, py) + py * _owner_box_1d(gk, NZ, pz))).astype(np.int32)
    boundary = ((gi == 0) | (gi == NX) | (gj == 0) | (gj == NY) | (gk == 0) | (gk == NZ)).astype(np.uint8)
    # local cells
    ci, cj, ck = np.meshgrid(np.arange(ncx), np.arange(ncy), np.arange(ncz), indexing="ij")
    ci, cj, ck = (a.transpose(2, 1, 0).reshape(-1) for a in (ci, cj, ck))
    nid = lambda i, j, k: i + nnx * (j + nny * k)
    offs = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
    conn = np.ascontiguousarray(np.stack([nid(ci + a, cj + b_, ck + c) for a, b_, c in offs], axis=1).astype(np.int32))
    e_owned = ((ci + lx < hx) & (cj + ly < hy) & (ck + lz < hz)).astype(np.uint8)
    egid = (ci + lx) + NX * ((cj + ly) + NY * (ck + lz).astype(np.int64))
    lm = LocalMesh(nnodes=int(gid.size), nelems=int(conn.shape[0]), conn=conn, elem_owned=e_owned,
                   node_owned=(owner == rank).astype(np.uint8), node_gid=gid, elem_gid=egid, node_on_global_boundary=boundary,
                   n_owned_elems=int(e_owned.sum()), lattice=(NX, NY, NZ))
    if nranks == 1:
        return lm
    # halo plan: receive every non-owned local node from its owner; send my owned nodes that lie in a neighbour's box
    recv, send = {}, {}
    for q in np.unique(owner):
        if q != rank:
            recv[int(q)] = np.nonzero(owner == q)[0]  # ascending local id == ascending global id
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                bb = (b[0] + dx, b[1] + dy, b[2] + dz)
                if (dx, dy, dz) == (0, 0, 0) or not all(0 <= bb[d] < procs[d] for d in range(3)):
                    continue
                r = bb[0] + px * (bb[1] + py * bb[2])
                (ax0, ax1), (ay0, ay1), (az0, az1) = _local_node_box(r, cells, procs)
                sel = (owner == rank) & (gi >= ax0) & (gi <= ax1) & (gj >= ay0) & (gj <= ay1) & (gk >= az0) & (gk <= az1)
                ids = np.nonzero(sel)[0]
                if ids.size:
                    send[r] = ids
    nb = sorted(set(recv) | set(send))
    so, ro, sn, rn = [0], [0], [], []
    for q in nb:
        s = send.get(q, np.zeros(0, np.int64))
        r_ = recv.get(q, np.zeros(0, np.int64))
        sn.append(s)
        rn.append(r_)
        so.append(so[-1] + len(s))
        ro.append(ro[-1] + len(r_))
    lm.nb_ranks = np.array(nb, np.int32)
    lm.send_off, lm.recv_off = np.array(so, np.int64), np.array(ro, np.int64)
    lm.send_nodes = np.concatenate(sn).astype(np.int32) if sn else np.zeros(0, np.int32)
    lm.recv_nodes = np.concatenate(rn).astype(np.int32) if rn else np.zeros(0, np.int32)
    return lm


@dataclass
class LocalBlock:
    """One structured block per rank (BASELINE config 5): no ghost cells, interface nodes duplicated per block."""
    node_size: tuple
    origin: tuple                # global node index of local node (0,0,0)
    node_gid: np.ndarray
    fixed: np.ndarray            # global boundary nodes
    owned: np.ndarray            # 1 where this rank is the lowest rank holding the node (dots / node count)
    nb_ranks: np.ndarray
    if_off: np.ndarray
    if_nodes: np.ndarray
    lattice: tuple

    def global_lattice_coords(self, h):
        NX, NY, NZ = self.lattice
        g = self.node_gid
        i = g % (NX + 1)
        j = (g // (NX + 1)) % (NY + 1)
        k = g // ((NX + 1) * (NY + 1))
        return np.ascontiguousarray(np.stack([i * h, j * h, k * h]).astype(np.float64))


def sgrid_block(cells, procs, rank):
    """Block `rank` of the lattice cut into px*py*pz structured blocks, with the cross-rank interface lists
    (shared nodes per neighbouring rank, ascending rank, each list in ascending global node id)."""
    NX, NY, NZ = cells
    px, py, pz = procs
    b = _box_of_rank(rank, procs)
    lo, hi = [], []
    for d in range(3):
        l, h_, _ = _ranges(cells[d], procs[d], b[d])
        lo.append(l)
        hi.append(h_)
    nn = tuple(hi[d] - lo[d] + 1 for d in range(3))
    ii, jj, kk = np.meshgrid(np.arange(nn[0]), np.arange(nn[1]), np.arange(nn[2]), indexing="ij")
    ii, jj, kk = (a.transpose(2, 1, 0).reshape(-1) for a in (ii, jj, kk))
    gi, gj, gk = ii + lo[0], jj + lo[1], kk + lo[2]
    gid = gi + (NX + 1) * (gj + (NY + 1) * gk.astype(np.int64))
    owner = (_owner_box_1d(gi, NX, px) + px * (_owner_box_1d(gj, NY, py) + py * _owner_box_1d(gk, NZ, pz))).astype(np.int32)
    fixed = ((gi == 0) | (gi == NX) | (gj == 0) | (gj == NY) | (gk == 0) | (gk == NZ)).astype(np.uint8)
    nb, off, nodes = [], [0], []
    for r in range(px * py * pz):  # ascending rank
        if r == rank:
            continue
        bb = _box_of_rank(r, procs)
        if any(abs(bb[d] - b[d]) > 1 for d in range(3)):
            continue
        sel = np.ones(gid.size, dtype=bool)
        for d, g in enumerate((gi, gj, gk)):
            l, h_, _ = _ranges(cells[d], procs[d], bb[d])
            sel &= (g >= l) & (g <= h_)
        ids = np.nonzero(sel)[0]
        if ids.size:
            nb.append(r)
            nodes.append(ids)  # ascending local id == ascending global id inside the shared face/edge/corner
            off.append(off[-1] + ids.size)
    return LocalBlock(node_size=nn, origin=tuple(lo), node_gid=gid, fixed=fixed, owned=(owner == rank).astype(np.uint8),
                      nb_ranks=np.array(nb, np.int32), if_off=np.array(off, np.int64),
                      if_nodes=np.concatenate(nodes).astype(np.int32) if nodes else np.zeros(0, np.int32), lattice=(NX, NY, NZ))


def box_elem_rank(cells, procs):
    """Part id of every cell of the global lattice for the box decomposition (element id = ci + NX*(cj + NY*ck))."""
    NX, NY, NZ = cells
    px, py, pz = procs
    ci, cj, ck = np.meshgrid(np.arange(NX), np.arange(NY), np.arange(NZ), indexing="ij")
    ci, cj, ck = (a.transpose(2, 1, 0).reshape(-1) for a in (ci, cj, ck))
    bx = np.minimum(ci // (NX // px), px - 1)
    by = np.minimum(cj // (NY // py), py - 1)
    bz = np.minimum(ck // (NZ // pz), pz - 1)
    return (bx + px * (by + py * bz)).astype(np.int32)


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def perturb_by_global_id(W, gid, fixed, eps, seed=1):
    """x_d += eps*(u - 0.5), u a counter-based uniform of (seed, global node id, d): identical on every rank that
    holds a copy of the node (a rank-local rand() stream would not be)."""
    X = W.copy()
    with np.errstate(over="ignore"):
        for d in range(3):
            key = gid.astype(np.uint64) * np.uint64(3) + np.uint64(d) + np.uint64(seed) * np.uint64(0x100000001B3)
            u = (_splitmix64(key) >> np.uint64(11)).astype(np.float64) / float(1 << 53)
            X[d] += np.where(fixed != 0, 0.0, eps * (u - 0.5))
    return X
