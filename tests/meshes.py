"""Builds identical meshes on any backend (product `pms_` ctx or oracle `orc_` ctx)."""
import numpy as np

from tests.orc import fixture_cube, perturb_cube_rand


def boundary_mask(nn):
    i = np.arange(nn)[None, None, :]
    j = np.arange(nn)[None, :, None]
    k = np.arange(nn)[:, None, None]
    m = (i == 0) | (i == nn - 1) | (j == 0) | (j == nn - 1) | (k == 0) | (k == nn - 1)
    return np.ascontiguousarray(np.broadcast_to(m, (nn, nn, nn)).reshape(-1).astype(np.uint8))


def hex8_conn(n):
    ci, cj, ck = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    ci, cj, ck = (a.transpose(2, 1, 0).reshape(-1) for a in (ci, cj, ck))
    nid = lambda i, j, k: i + (n + 1) * (j + (n + 1) * k)
    offs = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
    return np.ascontiguousarray(np.stack([nid(ci + a, cj + b, ck + c) for a, b, c in offs], axis=1).astype(np.int32))


def tet4_conn(n):
    h = hex8_conn(n)
    paths = [(1, 2), (2, 3), (3, 7), (7, 4), (4, 5), (5, 1)]
    return np.ascontiguousarray(np.stack([h[:, [0, a, b, 6]] for a, b in paths], axis=1).reshape(-1, 4).astype(np.int32))


def perturbed_coords(nele, eps_factor, seed=1):
    W = fixture_cube(nele + 1)
    X = perturb_cube_rand(W, nele, eps_factor / nele, seed)
    return X, W


def sinus_coords(nele, amp=0.2):
    W = fixture_cube(nele + 1)
    h = 1.0 / nele
    b = amp * h * np.sin(2 * np.pi * 3 * W[0]) * np.sin(2 * np.pi * 5 * W[1]) * np.sin(2 * np.pi * 7 * W[2])
    return W + b[None, :], W


def build(api, kind, nele, X, W, fixed=True, options=None):
    """kind in {'sgrid','hex8','tet4'}; X current/projected coords, W reference coords (3,N) SoA."""
    c = api.create()
    for k, v in (options or {}).items():
        c.set_option(k, v)
    nn = nele + 1
    if kind == "sgrid":
        b = c.add_structured_block((nn,) * 3)
        c.add_fixture_1_bcs(b)
        if fixed:
            c.select_boundary_sgrid()
    else:
        conn = hex8_conn(nele) if kind == "hex8" else tet4_conn(nele)
        c.set_unstructured(8 if kind == "hex8" else 4, nn ** 3, conn)
        if fixed:
            c.set_fixed_mask(boundary_mask(nn))
    c.commit()
    c.upload("coordinates", X)
    c.upload("coordinates_N", X)
    c.upload("coordinates_NM1", W)
    return c


def seeded_field(n, seed, scale=1.0):
    rng = np.random.default_rng(seed)
    return scale * rng.standard_normal((3, n))
