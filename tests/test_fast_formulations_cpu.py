"""The algebra of the flop-reduced device code, checked on the CPU against the oracle (no GPU needed):

* hex_metric_fast_v (percept_b200/csrc/pms_device_fast.cuh): canonical +axis edge frame, Gram matrix G = C C^T of the
  reference cofactor rows, B = A G serving both y = |A cof(W)^T|^2 = A:B and the derivative, ONE reciprocal square root
  u = 1/sqrt(y (P detA)^2) refined by a single third-order step from a 20-bit seed, and the gradient assembled per hex
  EDGE (the two end corners' terms k1 B[r][s] - k2 cof(A)[s][r] chained into one accumulator, a vertex = the signed sum of
  its three edges);
* tet_reference / tet_metric_cached (pms_device.cuh): W^-1 and det W of a tet computed once from the reference vertices
  and re-used by every later evaluation.

These numpy restatements follow the kernels' formulation, not the reference's; they must reproduce the oracle's (= the
reference's) total metric, invalid counts and nodal gradient to rounding.
"""
import numpy as np
import pytest

from tests import meshes
from tests.orc import oracle

EXO_OF_POS = [0, 1, 3, 2, 4, 5, 7, 6]  # structured position (i + 2j + 4k) -> Exodus local node
FAC = 3.0 * np.sqrt(3.0)


def truncated_rsqrt_seed(z):
    """A seed with ~20 good bits, as rsqrt.approx.ftz.f64 delivers: exact value with the mantissa cut to 20 bits."""
    r = 1.0 / np.sqrt(z)
    m, e = np.frexp(r)
    return np.ldexp(np.floor(m * 2.0 ** 20) / 2.0 ** 20, e)


def fast_rsqrt(z):
    r = truncated_rsqrt_seed(z)
    h = (0.5 * z) * r
    e = 0.5 - h * r
    return r + (r * e) * (1.0 + 1.5 * e)  # r (1 + e + 1.5 e^2): remainder 2.5 e^3


def hex_metric_and_gradient(X, W, conn, stage):
    """-> (total metric, invalid elements, nodal gradient (3, N)); element contributions added in ascending element order"""
    pos = conn[:, EXO_OF_POS]
    vc, vo = X.T[pos], W.T[pos]  # (E, 8, 3)
    E = conn.shape[0]
    val = np.zeros(E)
    invalid = np.zeros(E, dtype=bool)
    edge = np.zeros((E, 3, 4, 3))  # [axis s][edge index][component r]
    cr = np.cross
    dot = lambda u, v: (u * v).sum(-1)
    for c in range(8):
        a = [vc[:, c | (1 << s)] - vc[:, c & ~(1 << s)] for s in range(3)]
        w = [vo[:, c | (1 << s)] - vo[:, c & ~(1 << s)] for s in range(3)]
        K = [cr(a[1], a[2]), cr(a[2], a[0]), cr(a[0], a[1])]  # cofactor columns of A
        detA = dot(a[0], K[0])
        ok = detA > 0
        invalid |= ~ok
        if stage == 0:
            detW = dot(w[0], cr(w[1], w[2]))
            vp = np.maximum(0.05 * detW - detA, 0.0)
            val += vp * vp
            k1, nk2 = np.zeros(E), -2.0 * vp
            B = np.zeros((E, 3, 3))
        else:
            C = [cr(w[1], w[2]), cr(w[2], w[0]), cr(w[0], w[1])]
            detW = dot(w[0], C[0])
            G = np.array([[dot(C[i], C[j]) for j in range(3)] for i in range(3)])  # (3, 3, E)
            A = np.stack(a, axis=2)  # A[e][r][i] = a_i[r]
            B = np.einsum("eri,ije->erj", A, G)
            y = (A * B).sum((1, 2))
            P = FAC * np.abs(detW) * detA
            Pd = P * detA
            with np.errstate(all="ignore"):
                yrd = y * fast_rsqrt(y * Pd * Pd)  # sqrt(y) / (P detA)
            yr = yrd * detA
            q = y * yr
            val += np.where(ok, q - detW, 0.0)
            k1, nk2 = 3.0 * yr, -y * yrd
        for s in range(3):
            lo = c & ~(1 << s)
            ei = ((lo >> (s + 1)) << s) | (lo & ((1 << s) - 1))
            edge[:, s, ei] += k1[:, None] * B[:, :, s] + nk2[:, None] * K[s]
    grad = np.zeros((E, 8, 3))
    for v in range(8):
        e0, e1, e2 = edge[:, 0, v >> 1], edge[:, 1, ((v >> 2) << 1) | (v & 1)], edge[:, 2, v & 3]
        grad[:, v] = ((e0 if v & 1 else -e0) + (e1 if v & 2 else -e1)) + (e2 if v & 4 else -e2)
    use = np.ones(E, dtype=bool) if stage == 0 else ~invalid  # gradient_functors.hpp:391
    g = np.zeros((X.shape[1], 3))
    np.add.at(g, pos[use].reshape(-1), grad[use].reshape(-1, 3))
    return val.sum(), int(invalid.sum()), g.T


@pytest.mark.parametrize("eps", [0.25, 1.0])
@pytest.mark.parametrize("stage", [0, 1])
def test_edge_assembled_gram_gradient_matches_oracle(eps, stage):
    nele = 6
    X, W = meshes.perturbed_coords(nele, eps)
    o = meshes.build(oracle, "hex8", nele, X, W, False, {"real_is_double": 1})  # nothing fixed: every node's sum is checked
    mo, no = o.metric_and_gradient(stage)
    go = o.download("cg_g")
    m, ninv, g = hex_metric_and_gradient(X, W, meshes.hex8_conn(nele), stage)
    assert ninv == no
    tol = 1e-12 if eps < 1.0 else 2e-11  # tangled mesh + ShapeB1: corners with det A -> 0+ dominate (see the GPU tests)
    assert abs(m - mo) <= tol * abs(mo)
    assert np.max(np.abs(g - go)) <= tol * np.max(np.abs(go))
    if eps == 1.0:
        assert no > 0  # the tangled mesh exercises the invalid-corner / dropped-element paths


def test_third_order_rsqrt_step_reaches_double_precision_from_a_20_bit_seed():
    z = np.exp(np.random.default_rng(0).uniform(-60, 60, 100000))
    r = fast_rsqrt(z)
    assert np.max(np.abs(r * np.sqrt(z) - 1.0)) < 4e-16
    seed_err = np.max(np.abs(truncated_rsqrt_seed(z) * np.sqrt(z) - 1.0))
    assert 1e-7 < seed_err < 2.1e-6  # the seed really is that coarse


ISQRT3, ISQRT6 = 1.0 / np.sqrt(3.0), 1.0 / np.sqrt(6.0)


def tet_jacobian(v):  # jacobian_matrix_tet_3D, JacobianUtil.hpp:175-205; v: (E, 4, 3) -> A[e][r][c]
    f = v[:, 1] + v[:, 0]
    return np.stack([v[:, 1] - v[:, 0], (2.0 * v[:, 2] - f) * ISQRT3, (3.0 * v[:, 3] - v[:, 2] - f) * ISQRT6], axis=2)


def tet_metric_and_gradient_cached(X, ref, conn, stage):
    """ref = (WI (E,3,3), detW (E,)): the per-element cache k_tet_ref builds"""
    WI, detW = ref
    A = tet_jacobian(X.T[conn])
    detA = np.linalg.det(A)
    E = conn.shape[0]
    invalid = ~(detA > 0)
    dM = np.zeros((E, 3, 3))
    if stage == 0:
        vv = 0.05 * detW - detA
        m = np.maximum(vv, 0.0) ** 2
        TA = np.linalg.det(A)[:, None, None] * np.transpose(np.linalg.inv(A), (0, 2, 1))  # transpose adjugate = cof(A)
        dM = np.where((vv > 0)[:, None, None], (-2.0 * vv)[:, None, None] * TA, 0.0)
    else:
        T = A @ WI
        d = detA / detW
        with np.errstate(all="ignore"):
            f = np.sqrt((T * T).sum((1, 2)))
            den = FAC * d
            m = np.where(~invalid, (f ** 3 / den - 1.0) * detW, 0.0)
            cofT = np.linalg.det(T)[:, None, None] * np.transpose(np.linalg.inv(T), (0, 2, 1))
            d_dT = (3.0 * f / den)[:, None, None] * T - (f ** 3 / den / d)[:, None, None] * cofT
            dM = np.where((~invalid)[:, None, None], (d_dT @ np.transpose(WI, (0, 2, 1))) * detW[:, None, None], 0.0)
    # grad_util_tet_3d (JacobianUtil.cpp:215-243), the same corner counted 4 times
    G = np.zeros((E, 4, 3))
    G[:, 1] += dM[:, :, 0]
    G[:, 0] -= dM[:, :, 0]
    G[:, 2] += dM[:, :, 1] * (2.0 * ISQRT3)
    G[:, 1] -= dM[:, :, 1] * ISQRT3
    G[:, 0] -= dM[:, :, 1] * ISQRT3
    G[:, 3] += dM[:, :, 2] * (3.0 * ISQRT6)
    for a in range(3):
        G[:, a] -= dM[:, :, 2] * ISQRT6
    G *= 4.0
    use = np.ones(E, dtype=bool) if stage == 0 else ~invalid
    g = np.zeros((X.shape[1], 3))
    np.add.at(g, conn[use].reshape(-1), G[use].reshape(-1, 3))
    return 4.0 * m.sum(), int(invalid.sum()), g.T


@pytest.mark.parametrize("stage", [0, 1])
def test_cached_tet_reference_matches_oracle_and_is_independent_of_the_current_mesh(stage):
    nele = 5
    X, W = meshes.perturbed_coords(nele, 1.0 if stage == 0 else 0.25)
    conn = meshes.tet4_conn(nele)
    Wj = tet_jacobian(W.T[conn])
    ref = (np.linalg.inv(Wj), np.linalg.det(Wj))  # computed ONCE from the reference mesh ...
    o = meshes.build(oracle, "tet4", nele, X, W, False, {"real_is_double": 1})
    for it in range(2):  # ... and re-used for a second, different current mesh
        mo, no = o.metric_and_gradient(stage)
        go = o.download("cg_g")
        m, ninv, g = tet_metric_and_gradient_cached(X, ref, conn, stage)
        assert ninv == no
        assert abs(m - mo) <= 1e-12 * abs(mo)
        assert np.max(np.abs(g - go)) <= 1e-11 * np.max(np.abs(go))
        X = X + 0.02 / nele * np.sin(37.0 * X[::-1])
        for f in ("coordinates", "coordinates_N"):
            o.upload(f, X)


# ------------------------------------------------------------------------------------------------------------------
# Validity-first line-search batches (Driver::speculate / op_total_metric in pms_api.cpp): on an untangled mesh a trial
# with invalid elements is answered with (NaN, ninv) instead of its metric.  The reference's Armijo loops
# (ReferenceMeshSmootherConjugateGradientDef.hpp:497-548) must take the same decisions either way.
# ------------------------------------------------------------------------------------------------------------------
def armijo_loops(trial, metric_0, alpha_init, offset_factor, untangled, dmax_relative, grad_norm, tau=0.5, niter=1000):
    """The two back-tracking loops of line_search; trial(alpha) -> (metric, n_invalid).  -> (converged, alpha, visited)"""
    alpha, visited, converged = alpha_init, [], False
    liter = 0
    while not converged and liter < niter:
        metric, ninv = trial(alpha)
        visited.append(alpha)
        converged = metric < metric_0 + alpha * offset_factor
        if untangled:
            converged = converged and ninv == 0
        if not converged:
            alpha *= tau
        if alpha < 1e-12 * alpha_init:
            break
        liter += 1
    liter = 0
    while not converged and liter < niter:  # (second loop, after the restart of the search direction)
        metric, ninv = trial(alpha)
        visited.append(alpha)
        converged = metric < metric_0 + alpha * offset_factor
        if untangled:
            converged = converged and ninv == 0
            if ninv == 0 and dmax_relative < grad_norm:
                converged = True
        if not converged:
            alpha *= tau
        if alpha < 1e-12 * alpha_init:
            break
        liter += 1
    return converged, alpha, visited


@pytest.mark.parametrize("seed", range(40))
def test_validity_only_answers_leave_the_armijo_decisions_unchanged(seed):
    rng = np.random.default_rng(seed)
    metric_0 = 1.0
    n_bad = int(rng.integers(0, 9))        # leading trials at which the mesh has invalid elements
    n_reject = int(rng.integers(0, 8))     # valid trials that still fail the Armijo test
    sometimes_lower = rng.random() < 0.5   # an invalid trial whose metric would pass the Armijo test on its own
    table = {}

    def full(alpha):
        k = int(round(-np.log2(alpha)))
        if k not in table:
            if k < n_bad:
                m = 0.5 if (sometimes_lower and k % 2 == 0) else 3.0 + rng.random()
                table[k] = (m, int(rng.integers(1, 1000)))
            elif k < n_bad + n_reject:
                table[k] = (1.5 + rng.random(), 0)
            else:
                table[k] = (0.9, 0)
        return table[k]

    def validity_only(alpha):
        m, ninv = full(alpha)
        return (float("nan"), ninv) if ninv > 0 else (m, ninv)

    args = (metric_0, 1.0, -1e-3, True, 1.0, 1e-6)
    a = armijo_loops(full, *args)
    b = armijo_loops(validity_only, *args)
    assert a == b
    assert a[0] and len(a[2]) == n_bad + n_reject + 1
