"""The algebra behind pms_total_metric_multi (hex_metric_multi in percept_b200/csrc/pms_device_fast.cuh), checked on the
CPU against the oracle: with x(alpha) = x + alpha*s every corner edge is a + alpha*b, so per corner

    det A(alpha)            = d0 + d1 alpha + d2 alpha^2 + d3 alpha^3        (triple-product expansion)
    |A(alpha) cof(W)^T|_F^2 = y0 + 2 y1 alpha + y2 alpha^2
    ShapeB1:  mu = y^1.5 / (3 sqrt3 |det W| det A) - det W      (= (|A W^-1|_F^3 / (3 sqrt3 det A / det W) - 1) det W,
                                                                    SmootherMetric.hpp:613-720)
    Untangle: mu = max(0, 0.05 det W - det A)^2

This numpy restatement of the kernel's formulation must reproduce the oracle's total metric and invalid-element count at
every trial step length of a backtracking sequence.
"""
import numpy as np
import pytest

from tests import meshes
from tests.orc import oracle

# corner c (structured numbering node = i + 2j + 4k): canonical +axis edges are v[c | bit] - v[c & ~bit]
EXO_OF_POS = [0, 1, 3, 2, 4, 5, 7, 6]  # structured position -> Exodus local node (permute_to_unstructured)


def corner_edges(V):
    """V: (E, 8, 3) vertices in structured position order -> (E, 8 corners, 3 axes, 3 comps) canonical edge vectors"""
    out = np.empty((V.shape[0], 8, 3, 3))
    for c in range(8):
        for ax in range(3):
            lo, hi = c & ~(1 << ax), c | (1 << ax)
            out[:, c, ax] = V[:, hi] - V[:, lo]
    return out


def multi_alpha_metric(X, S, W, conn, fixed, alphas, stage):
    pos = conn[:, EXO_OF_POS]
    Sm = S * (fixed == 0)[None, :]  # fixed nodes do not move (update_coordinates.cpp:255-256)
    a = corner_edges(X.T[pos])
    b = corner_edges(Sm.T[pos])
    w = corner_edges(W.T[pos])
    cr = np.cross
    K0, L, M = cr(a[:, :, 1], a[:, :, 2]), cr(a[:, :, 1], b[:, :, 2]) + cr(b[:, :, 1], a[:, :, 2]), cr(b[:, :, 1], b[:, :, 2])
    dot = lambda u, v: (u * v).sum(-1)
    d0, d1 = dot(a[:, :, 0], K0), dot(a[:, :, 0], L) + dot(b[:, :, 0], K0)
    d2, d3 = dot(a[:, :, 0], M) + dot(b[:, :, 0], L), dot(b[:, :, 0], M)
    C = np.stack([cr(w[:, :, 1], w[:, :, 2]), cr(w[:, :, 2], w[:, :, 0]), cr(w[:, :, 0], w[:, :, 1])], axis=2)  # cof(W) rows
    detW = dot(w[:, :, 0], C[:, :, 0])
    Tx = np.einsum("ecir,eciq->ecrq", a, C)  # sum_i a_i[r] C_i[q]
    Ts = np.einsum("ecir,eciq->ecrq", b, C)
    y0, y1, y2 = (Tx * Tx).sum((-1, -2)), (Tx * Ts).sum((-1, -2)), (Ts * Ts).sum((-1, -2))
    res = []
    for al in alphas:
        detA = d0 + al * (d1 + al * (d2 + al * d3))
        ok = detA > 0
        if stage == 0:
            mu = np.maximum(0.05 * detW - detA, 0.0) ** 2
        else:
            y = y0 + al * (2 * y1 + al * y2)
            with np.errstate(all="ignore"):
                mu = np.where(ok, y ** 1.5 / (3 * np.sqrt(3.0) * np.abs(detW) * detA) - detW, 0.0)
        res.append((mu.sum(), int((~ok).any(axis=1).sum())))
    return res


@pytest.mark.parametrize("eps", [0.25, 1.0])
@pytest.mark.parametrize("stage", [0, 1])
def test_polynomial_in_alpha_formulation_matches_oracle(eps, stage):
    nele = 7
    X, W = meshes.perturbed_coords(nele, eps)
    o = meshes.build(oracle, "hex8", nele, X, W, True, {"real_is_double": 1})
    S = meshes.seeded_field(X.shape[1], 4, 0.03 / nele)
    o.upload("cg_s", S)
    conn = meshes.hex8_conn(nele)
    fixed = meshes.boundary_mask(nele + 1)
    alphas = [6.0 * 0.5 ** k for k in range(12)] + [0.0, -0.4]
    mine = multi_alpha_metric(X, S, W, conn, fixed, alphas, stage)
    saw_invalid = saw_valid = False
    for al, (m, ninv) in zip(alphas, mine):
        mo, no = o.total_metric(stage, al)
        assert ninv == no
        tol = 1e-12 if eps < 1.0 else 2e-11  # tangled mesh + ShapeB1: ill-conditioned near det A -> 0+ (see the GPU test)
        assert abs(m - mo) <= tol * abs(mo) + 1e-300
        saw_invalid |= no > 0
        saw_valid |= no == 0
    if eps < 1.0:
        assert saw_invalid and saw_valid  # the sequence crosses from inverted to valid meshes
