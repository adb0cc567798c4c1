"""Parity of the CUDA hot path (through the C-ABI) against the oracle and the reference golden vector.
Tolerance: 1e-8 absolute on beliefs, pair beliefs and per-node Bethe free energy (BASELINE.json north_star)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import mpbp_b200 as M
from mpbp_b200 import _lib
from oracle import mpbp as O, tt as OT
from tests.common import build_pair, compare, otrunc
from tests.test_oracle_golden import SIS_INFINITE_GOLDEN

TOL = 1e-8


def _qr(A):
    b, m, n = A.shape
    k = min(m, n)
    R = np.zeros((b, k, n))
    Ac = np.ascontiguousarray(A)
    _lib.check(_lib.lib().mpbp_test_qr(Ac.ctypes.data_as(_lib.c_dp), b, m, n, R.ctypes.data_as(_lib.c_dp)))
    return R


@pytest.mark.parametrize("m,n", [(1, 1), (5, 3), (3, 5), (40, 8), (160, 40), (257, 17), (700, 100), (1600, 400), (2048, 33), (9, 9)])
def test_qr_r_factor(m, n):
    rng = np.random.default_rng(m * 1000 + n)
    A = rng.standard_normal((3, m, n))
    A[1, :, n // 2] = A[1, :, 0] * 2.0  # exactly dependent column
    A[2] *= np.logspace(0, -12, n)[None, :]  # badly scaled
    R = _qr(A)
    for b in range(3):
        Rn = np.linalg.qr(A[b], mode="r")
        assert np.allclose(np.tril(R[b], -1), 0)
        scale = np.abs(Rn).max()
        assert np.max(np.abs(np.abs(R[b]) - np.abs(Rn))) < 1e-12 * max(scale, 1) * max(m, n) ** 0.5 or \
            np.max(np.abs(R[b].T @ R[b] - A[b].T @ A[b])) < 1e-12 * np.abs(A[b].T @ A[b]).max()
        assert np.max(np.abs(R[b].T @ R[b] - A[b].T @ A[b])) < 1e-11 * np.abs(A[b].T @ A[b]).max()


@pytest.mark.parametrize("m,n,H", [(1, 1, 32), (5, 3, 32), (3, 5, 32), (40, 8, 32), (160, 40, 16), (257, 17, 32), (700, 100, 32),
                                   (1600, 400, 32), (4000, 400, 32), (333, 77, 16), (64, 600, 16), (900, 450, 16), (1600, 400, 64), (333, 77, 64),
                                   (70, 9, 64)])
def test_qr_flat_tree_dmma(m, n, H):
    rng = np.random.default_rng(m * 1000 + n)
    A = rng.standard_normal((3, m, n))
    A[1, :, n // 2] = A[1, :, 0] * 2.0
    A[2] *= np.logspace(0, -12, n)[None, :]
    R = np.zeros((3, n, n))
    ms = np.zeros(1)
    Ac = np.ascontiguousarray(A)
    _lib.check(_lib.lib().mpbp_test_qr_ft(Ac.ctypes.data_as(_lib.c_dp), 3, m, n, H, R.ctypes.data_as(_lib.c_dp), ms.ctypes.data_as(_lib.c_dp)))
    for b in range(3):
        # R is n x n upper triangular (for m < n up to m rows are non-negligible, not necessarily the first m)
        assert np.allclose(np.tril(R[b], -1), 0)
        G = A[b].T @ A[b]
        assert np.max(np.abs(R[b].T @ R[b] - G)) < 1e-11 * np.abs(G).max()


@pytest.mark.parametrize("p,n,d,decay", [(200, 400, 20, 0.85), (440, 400, 20, 0.9), (80, 400, 20, 0.7), (100, 100, 10, 0.8),
                                          (60, 30, 10, 0.5), (640, 400, 20, 0.93), (200, 400, 20, 0.97), (660, 900, 30, 0.9),
                                          (72, 100, 10, 0.6)])
def test_truncation_svd_core_vs_numpy(p, n, d, decay):
    """the op-truncation SVD (direct Jacobi / blocked subspace iteration) against LAPACK: what matters is the
    rank-d projection P M, weighted by the singular values"""
    rng = np.random.default_rng(p + n)
    batch, c = 2, min(p, n)
    Ms = []
    for b in range(batch):
        Uq, _ = np.linalg.qr(rng.standard_normal((p, c)))
        Vq, _ = np.linalg.qr(rng.standard_normal((n, c)))
        s = decay ** np.arange(c) * (1 + 0.3 * rng.random(c))
        Ms.append((Uq * s) @ Vq.T)
    Mall = np.ascontiguousarray(np.stack([m.T for m in Ms]))  # column-major p x n == row-major n x p
    U = np.zeros((batch, d, p)); S = np.zeros((batch, d)); st = np.zeros(5); ms = np.zeros(1)
    _lib.check(_lib.lib().mpbp_test_svd(Mall.ctypes.data_as(_lib.c_dp), batch, p, n, d, U.ctypes.data_as(_lib.c_dp),
                                        S.ctypes.data_as(_lib.c_dp), st.ctypes.data_as(_lib.c_dp), ms.ctypes.data_as(_lib.c_dp)))
    for b in range(batch):
        Un, sn, _ = np.linalg.svd(Ms[b], full_matrices=False)
        Ud = U[b].T  # p x d
        assert np.allclose(S[b], sn[:d], rtol=1e-9, atol=1e-13 * sn[0])
        assert np.max(np.abs(Ud.T @ Ud - np.eye(d))) < 1e-10
        err = np.linalg.norm(Un[:, :d] @ (Un[:, :d].T @ Ms[b]) - Ud @ (Ud.T @ Ms[b])) / sn[0]
        # (decay 0.97: sigma_21/sigma_20 = 0.97, the truncation itself discards half of the weight; the subspace is then
        # resolved to 1e-10 of sigma_1 instead of 1e-11)
        assert err < (1e-10 if decay >= 0.97 else 1e-11), (err, st)


@pytest.mark.parametrize("p,c", [(4, 2), (80, 40), (80, 80), (33, 7), (40, 60), (135, 45)])
def test_jacobi_singular_values(p, c):
    rng = np.random.default_rng(p * 100 + c)
    A = rng.standard_normal((2, c, p)) * np.logspace(0, -9, c)[None, :, None]  # batch of col-major p x c
    A0 = A.copy()
    sig = np.zeros((2, c))
    order = np.zeros((2, c), dtype=np.int32)
    _lib.check(_lib.lib().mpbp_test_jacobi(A.ctypes.data_as(_lib.c_dp), 2, p, c, sig.ctypes.data_as(_lib.c_dp), order.ctypes.data_as(_lib.c_i32p)))
    for b in range(2):
        Mx = A0[b].T  # p x c
        s = np.linalg.svd(Mx, compute_uv=False)
        k = min(p, c)
        assert np.allclose(sig[b][:k], s[:k], rtol=1e-10, atol=1e-13 * s[0])
        U = A[b].T[:, order[b][:k]] / sig[b][:k]
        G = U.T @ U
        good = sig[b][:k] > 1e-13 * s[0]
        assert np.max(np.abs(G[np.ix_(good, good)] - np.eye(good.sum()))) < 1e-10
        # projector onto the leading left singular vectors matches numpy's
        Un = np.linalg.svd(Mx, full_matrices=False)[0]
        kk = max(1, min(k, 5))
        assert np.max(np.abs(U[:, :kk] @ U[:, :kk].T - Un[:, :kk] @ Un[:, :kk].T)) < 1e-7


def test_sis_infinite_graph_golden_device():
    # /root/reference/test/sis_infinite_graph.jl:3-29 through the CUDA path
    T, k, gamma, lam, rho = 6, 3, 0.1, 0.1, 0.2
    w = [M.SISFactor(lam, rho)] * (T + 1)
    phi = [np.array([1 - gamma, gamma]) if t == 0 else np.ones(2) for t in range(T + 1)]
    bp = M.mpbp_infinite_graph(k, w, 2, phi, dmax=10)
    iters, cb = M.iterate_(bp, maxiter=200, svd_trunc=M.TruncBond(10), tol=1e-14)
    assert iters < 200
    b = M.beliefs(bp)[0]
    assert np.max(np.abs(b - SIS_INFINITE_GOLDEN)) < TOL


def _tree_case(seed, T=2):
    rng = np.random.default_rng(seed)
    und = [(0, 1), (1, 2), (1, 3), (3, 4)]
    N = 5
    h = rng.standard_normal(N)
    kinds = [("glauber", (1.0, float(h[i]), 1.0)) for i in range(N)]
    phi = [[np.ones(2) for _ in range(T + 1)] for _ in range(N)]
    for i in range(N):
        phi[i][0] = np.array([0.75, 0.25])
        t = int(rng.integers(1, T + 1))
        o = np.full(2, 0.05)
        o[rng.integers(2)] = 1.0
        phi[i][t] = phi[i][t] * o
    return N, und, kinds, phi


@pytest.mark.parametrize("schedule", ["sequential", "parallel"])
def test_glauber_small_tree_vs_oracle(schedule):
    T = 2
    N, und, kinds, phi = _tree_case(111, T)
    bo, bd = build_pair(N, und, T, kinds, [2] * N, phi, dmax=10)
    tr = M.TruncBondThresh(10)
    O.iterate(bo, maxiter=6, trunc=otrunc(tr), tol=0.0, schedule=schedule)
    M.iterate_(bd, maxiter=6, svd_trunc=tr, tol=0.0, shuffle_nodes=False, schedule=schedule)
    eb, ef, ep = compare(bo, bd)
    assert eb < TOL and ef < TOL and ep < TOL, (eb, ef, ep)
    # exactness on a tree: Z_bp == Z_exact (reference test/glauber_small_tree.jl:63)
    from oracle import exact
    p, Z, logZ = exact.exact_prob(bo)
    assert abs(-M.bethe_free_energy(bd) - logZ) < 1e-8


@pytest.mark.parametrize("schedule", ["sequential", "parallel"])
def test_sis_loopy_truncated_vs_oracle(schedule):
    # graph of test/sis_heterogeneous_compare_homogeneous.jl:5-10, truncation active (TruncBond(3))
    T = 4
    und = [(0, 1), (0, 2), (1, 2), (2, 3), (3, 4)]
    N = 5
    kinds = [("sis", (0.15 + 0.02 * i, 0.12, 0.01)) for i in range(N)]
    phi = [[np.array([0.87, 0.13]) if t == 0 else np.ones(2) for t in range(T + 1)] for _ in range(N)]
    phi[2][3] = np.array([0.2, 1.0])
    rng = np.random.default_rng(3)
    go = O.BiDiGraph(N, und)
    psi = []
    for e in range(go.ne):
        psi.append(None)
    for e in range(go.ne):
        if psi[e] is None:
            ps = [0.5 + rng.random((2, 2)) for _ in range(T + 1)]
            psi[e] = ps
            psi[go.rev[e]] = [p.T.copy() for p in ps]
    bo, bd = build_pair(N, und, T, kinds, [2] * N, phi, psi=psi, dmax=4)
    tr = M.TruncBond(3)
    O.iterate(bo, maxiter=4, trunc=otrunc(tr), tol=0.0, schedule=schedule)
    M.iterate_(bd, maxiter=4, svd_trunc=tr, tol=0.0, shuffle_nodes=False, schedule=schedule)
    eb, ef, ep = compare(bo, bd)
    assert eb < TOL and ef < TOL and ep < TOL, (eb, ef, ep)
    for e in range(bd.E2):  # test/normalizations.jl:48-52
        msg = bd.get_message(e)
        l = np.ones((1, 1))
        for a in msg:
            l = l @ a.sum(axis=(2, 3))
        assert abs(l[0, 0] - 1) < 1e-10


def test_sirs_chain_vs_oracle():
    T = 3
    und = [(0, 1), (1, 2), (2, 3)]
    N = 4
    kinds = [("sirs", (0.4, 0.15, 0.2, 0.05))] * N
    phi = [[np.array([0.7, 0.3, 0.0]) if t == 0 else np.ones(3) for t in range(T + 1)] for _ in range(N)]
    phi[0][2] = np.array([0.1, 1.0, 0.3])
    bo, bd = build_pair(N, und, T, kinds, [3] * N, phi, dmax=9)
    tr = M.TruncThresh(1e-9)
    O.iterate(bo, maxiter=5, trunc=otrunc(tr), tol=0.0)
    M.iterate_(bd, maxiter=5, svd_trunc=tr, tol=0.0, shuffle_nodes=False)
    eb, ef, ep = compare(bo, bd)
    assert eb < TOL and ef < TOL and ep < TOL, (eb, ef, ep)


def test_glauber_degree5_star_truncated_vs_oracle():
    # exercises nstates growth (l+1), all cavity levels of a z=5 node and a TSQR-free heavy op
    T = 3
    und = [(0, k) for k in range(1, 6)] + [(1, 2)]
    N = 6
    kinds = [("glauber", (0.4, 0.1 * (i - 2), 1.0)) for i in range(N)]
    phi = [[np.array([0.2, 0.8]) if t == 0 else np.ones(2) for t in range(T + 1)] for _ in range(N)]
    bo, bd = build_pair(N, und, T, kinds, [2] * N, phi, dmax=4)
    tr = M.TruncBond(4)
    O.iterate(bo, maxiter=3, trunc=otrunc(tr), tol=0.0)
    M.iterate_(bd, maxiter=3, svd_trunc=tr, tol=0.0, shuffle_nodes=False)
    eb, ef, ep = compare(bo, bd)
    assert eb < TOL and ef < TOL and ep < TOL, (eb, ef, ep)


@pytest.mark.parametrize("schedule", ["sequential", "parallel"])
def test_mixed_degrees_level_staggering_is_invisible(schedule):
    # hub of degree 6, nodes of degree 1..4 and a cycle: the engine staggers the cavity levels of independent nodes
    # (option level_balance) to fill the GPU; the results must not depend on it, bit for bit, and match the oracle
    T = 3
    und = [(0, k) for k in range(1, 7)] + [(1, 2), (2, 3), (3, 4), (4, 7), (7, 8), (8, 4), (5, 8)]
    N = 9
    kinds = [("glauber", (0.3 + 0.05 * i, 0.1 * (i - 4), 1.0)) for i in range(N)]
    phi = [[np.array([0.3, 0.7]) if t == 0 else np.ones(2) for t in range(T + 1)] for _ in range(N)]
    tr = M.TruncBond(4)
    res = []
    for lb in (1.0, 0.0):
        bo, bd = build_pair(N, und, T, kinds, [2] * N, phi, dmax=4)
        bd.set_option("level_balance", lb)
        M.iterate_(bd, maxiter=3, svd_trunc=tr, tol=0.0, shuffle_nodes=False, schedule=schedule)
        res.append((np.concatenate([np.array(b).ravel() for b in M.beliefs(bd)]), M.api.free_energy_contributions(bd).copy()))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    O.iterate(bo, maxiter=3, trunc=otrunc(tr), tol=0.0, schedule=schedule)
    eb, ef, ep = compare(bo, bd)
    assert eb < TOL and ef < TOL and ep < TOL, (eb, ef, ep)


def test_glauber_4regular_bond10_subspace_svd_vs_oracle():
    # D = 100, d~X up to 100 > 64: the truncating SVDs take the blocked subspace-iteration path
    import networkx as nx
    T, N, d = 5, 6, 10
    G = nx.random_regular_graph(4, N, seed=3)
    und = [(int(a), int(b)) for a, b in G.edges()]
    kinds = [("glauber", (0.5, 0.1 + 0.03 * i, 1.0)) for i in range(N)]
    phi = [[np.array([0.2, 0.8]) if t == 0 else np.ones(2) for t in range(T + 1)] for _ in range(N)]
    phi[1][3] = np.array([0.7, 0.4])
    bo, bd = build_pair(N, und, T, kinds, [2] * N, phi, dmax=d)
    tr = M.TruncBond(d)
    O.iterate(bo, maxiter=4, trunc=otrunc(tr), tol=0.0, schedule="parallel")
    M.iterate_(bd, maxiter=4, svd_trunc=tr, tol=0.0, shuffle_nodes=False, schedule="parallel")
    eb, ef, ep = compare(bo, bd)
    assert eb < TOL and ef < TOL and ep < TOL, (eb, ef, ep)


def test_time_dependent_factors_vs_oracle():
    # w[i][t] differs over t -> the nt = T+1 table path (time-dependent Pxy / Pyy / W slices)
    from oracle import factors as OF
    T, N = 4, 4
    und = [(0, 1), (1, 2), (2, 3), (3, 0), (0, 2)]
    go = O.BiDiGraph(N, und)
    gd = M.IndexedBiDiGraph(N, und)
    lam = [0.1 + 0.05 * t for t in range(T + 1)]
    wo = [[OF.SISFactor(lam[t] + 0.01 * i, 0.2, 0.01) for t in range(T + 1)] for i in range(N)]
    wd = [[M.SISFactor(lam[t] + 0.01 * i, 0.2, 0.01) for t in range(T + 1)] for i in range(N)]
    phi = [[np.array([0.85, 0.15]) if t == 0 else np.ones(2) for t in range(T + 1)] for _ in range(N)]
    phi[1][2] = np.array([0.4, 0.9])
    bo = O.MPBP(go, wo, [2] * N, T, phi=[[p.copy() for p in ph] for ph in phi])
    bd = M.mpbp(gd, wd, [2] * N, T, phi=phi, dmax=5)
    tr = M.TruncBond(5)
    O.iterate(bo, maxiter=3, trunc=otrunc(tr), tol=0.0)
    M.iterate_(bd, maxiter=3, svd_trunc=tr, tol=0.0, shuffle_nodes=False)
    eb, ef, ep = compare(bo, bd)
    assert eb < TOL and ef < TOL and ep < TOL, (eb, ef, ep)


@pytest.mark.parametrize("kind", ["pmj", "intglauber"])
def test_pmj_and_integer_glauber_vs_oracle(kind):
    # factors whose prob_xy depends on the neighbour index k (glauber_bp.jl:58-91,144-179)
    T, N = 3, 4
    und = [(0, 1), (0, 2), (0, 3), (1, 2)]
    deg = {0: 3, 1: 2, 2: 2, 3: 1}
    kinds = []
    for i in range(N):
        z = deg[i]
        if kind == "pmj":
            kinds.append(("pmj", ([1, -1, 1][:z], 0.6, 0.1 * i, 1.0)))
        else:
            kinds.append(("intglauber", ([1, -2, 1][:z], 0.1 * i, 0.7)))
    phi = [[np.array([0.3, 0.7]) if t == 0 else np.ones(2) for t in range(T + 1)] for _ in range(N)]
    bo, bd = build_pair(N, und, T, kinds, [2] * N, phi, dmax=8)
    tr = M.TruncBond(6)
    O.iterate(bo, maxiter=3, trunc=otrunc(tr), tol=0.0)
    M.iterate_(bd, maxiter=3, svd_trunc=tr, tol=0.0, shuffle_nodes=False)
    eb, ef, ep = compare(bo, bd)
    assert eb < TOL and ef < TOL and ep < TOL, (eb, ef, ep)


def test_nodes_subset_and_visiting_order_vs_oracle():
    # iterate!(bp; nodes=...) with an explicit (shuffled) visiting order: the level-scheduled device sweep must
    # equal the serial in-place sweep of src/mpbp.jl:189-192
    T = 3
    und = [(0, 1), (1, 2), (2, 3), (3, 4), (4, 0), (1, 3)]
    N = 5
    kinds = [("sis", (0.25, 0.1, 0.02))] * N
    phi = [[np.array([0.7, 0.3]) if t == 0 else np.ones(2) for t in range(T + 1)] for _ in range(N)]
    bo, bd = build_pair(N, und, T, kinds, [2] * N, phi, dmax=4)
    tr = M.TruncBond(4)
    order = [3, 0, 4, 1]  # node 2 is never updated
    for it in range(2):
        O.iterate(bo, maxiter=1, trunc=otrunc(tr), tol=0.0, nodes=order)
        M.iterate_(bd, maxiter=1, svd_trunc=tr, tol=0.0, nodes=order, shuffle_nodes=False)
    eb, ef, ep = compare(bo, bd)
    assert eb < TOL and ef < TOL and ep < TOL, (eb, ef, ep)


def test_sis_heterogeneous_star_and_model_vs_oracle():
    # per-neighbour infection probabilities (src/Models/epidemics/sis_heterogeneous_bp.jl:69-72): the Pxy tables are per
    # (node class, neighbour); star of test/sis_heterogeneous.jl plus a loopy edge, truncation active
    T, N = 3, 5
    rng = np.random.default_rng(3)
    und = [(0, 1), (0, 2), (0, 3), (1, 2), (3, 4)]
    gd = M.IndexedBiDiGraph(N, und)
    lam = np.zeros((N, N))
    for a, b in und:
        lam[a, b], lam[b, a] = rng.random(), rng.random()
    rho, alpha = rng.random(N), 0.2 * rng.random(N)
    kinds = [("sishet", ([lam[int(j), i] for j in gd.neighbors(i)], float(rho[i]), float(alpha[i]))) for i in range(N)]
    phi = [[np.array([0.6, 0.4]) if t == 0 else np.ones(2) for t in range(T + 1)] for _ in range(N)]
    phi[2][T] = np.array([0.1, 1.0])
    bo, bd = build_pair(N, und, T, kinds, [2] * N, phi, dmax=4)
    tr = M.TruncBond(3)
    O.iterate(bo, maxiter=4, trunc=otrunc(tr), tol=0.0)
    M.iterate_(bd, maxiter=4, svd_trunc=tr, tol=0.0, shuffle_nodes=False)
    eb, ef, ep = compare(bo, bd)
    assert eb < TOL and ef < TOL and ep < TOL, (eb, ef, ep)
    # the model constructor builds the same factors (sis_heterogeneous_factors, sis_heterogeneous.jl:51-53)
    bm = M.mpbp(M.SIS_heterogeneous(gd, lam, rho, T, alpha=alpha, phi=phi), dmax=4)
    M.iterate_(bm, maxiter=4, svd_trunc=tr, tol=0.0, shuffle_nodes=False)
    assert np.array_equal(np.array(M.beliefs(bm)), np.array(M.beliefs(bd)))
    # uniform rates reproduce the homogeneous model (test/sis_heterogeneous_compare_homogeneous.jl:12-35)
    bu = M.mpbp(M.SIS(gd, 0.15, 0.12, T, gamma=0.13), dmax=4)
    bh = M.mpbp(M.SIS_heterogeneous(gd, 0.15 * (lam > 0), 0.12, T, gamma=0.13), dmax=4)
    for b in (bu, bh):
        M.iterate_(b, maxiter=30, svd_trunc=tr, tol=1e-12, shuffle_nodes=False)
    assert np.allclose(np.array(M.beliefs(bu)), np.array(M.beliefs(bh)), atol=1e-12)


def test_glauber_infinite_graph_free_energy_known_answer():
    # /root/reference/test/glauber_infinite_graph.jl:7-18 inputs, run without damping: f = 0.98812749675847 per node
    # (derived known answer, SURVEY.md header fact 3(ii)).  The reference test uses TruncThresh(0.0) (no device bond
    # capacity there); here TruncBondThresh(30, 1e-13): the exact ranks of the X = 8 intermediates can exceed dmax.
    T, k, m0 = 3, 3, 0.5
    w = [M.HomogeneousGlauberFactor(1.0, 0.0, 1.0)] * (T + 1)
    phi = [np.array([(1 + m0) / 2, (1 - m0) / 2]) if t == 0 else np.ones(2) for t in range(T + 1)]
    phi[1] = np.array([0.4, 0.6])
    phi[-1] = np.array([0.95, 0.05])
    bp = M.mpbp_infinite_graph(k, w, 2, phi, dmax=30)
    iters, cb = M.iterate_(bp, maxiter=150, svd_trunc=M.TruncBondThresh(30, 1e-13), tol=1e-14)
    assert abs(M.bethe_free_energy(bp) - 0.98812749675847) < 1e-8


def test_generic_factor_exhaustive_trace_vs_oracle():
    # GenericFactor forces the f_bp / f_bp_dummy_neighbor path (reference test/glauber_small_tree.jl:133-169,
    # src/bp_core.jl:18-93); on a tree with a non-binding truncation it must agree with the oracle's generic path
    # and with brute force
    from oracle import factors as OF, exact
    T, N = 2, 5
    und = [(0, 1), (1, 2), (1, 3)]
    rng = np.random.default_rng(7)
    h = rng.standard_normal(N)
    go = O.BiDiGraph(N, und)
    gd = M.IndexedBiDiGraph(N, und)
    wo = [[OF.GenericFactor(OF.HomogeneousGlauberFactor(1.0, float(h[i]), 1.0))] * (T + 1) for i in range(N)]
    wd = [[M.GenericFactor(M.HomogeneousGlauberFactor(1.0, float(h[i]), 1.0))] * (T + 1) for i in range(N)]
    # node 4 is isolated: give it a recursive factor (the reference handles degree 0 through the recursive path)
    wo[4] = [OF.HomogeneousGlauberFactor(1.0, float(h[4]), 1.0)] * (T + 1)
    wd[4] = [M.HomogeneousGlauberFactor(1.0, float(h[4]), 1.0)] * (T + 1)
    phi = [[np.array([0.75, 0.25]) if t == 0 else np.ones(2) for t in range(T + 1)] for _ in range(N)]
    phi[2][1] = np.array([1.0, 0.1])
    phi[0][2] = np.array([0.2, 1.0])
    bo = O.MPBP(go, wo, [2] * N, T, phi=[[p.copy() for p in ph] for ph in phi])
    bd = M.mpbp(gd, wd, [2] * N, T, phi=phi, dmax=16)
    O.iterate(bo, maxiter=4, trunc=OT.TruncThresh(0.0), tol=0.0)
    M.iterate_(bd, maxiter=4, svd_trunc=M.TruncBondThresh(16, 0.0), tol=0.0, shuffle_nodes=False)
    eb, ef, ep = compare(bo, bd)
    assert eb < TOL and ef < TOL and ep < TOL, (eb, ef, ep)
    p, Z, logZ = exact.exact_prob(bo)
    assert abs(-M.bethe_free_energy(bd) - logZ) < 1e-8


def test_pair_observations_generic_glauber_vs_oracle_and_exact():
    # test/pair_observations.jl:62-112: random couplings (GenericGlauberFactor -> exhaustive trace) with psi on four
    # (edge, time) pairs on a tree; device vs oracle vs brute force
    from oracle import factors as OF, exact
    T, N = 2, 5
    rng = np.random.default_rng(11)
    und = [(0, 1), (1, 2), (2, 3), (2, 4)]
    go = O.BiDiGraph(N, und)
    gd = M.IndexedBiDiGraph(N, und)
    obs = [(0, 1, 1, np.array([[0.1, 0.9], [0.3, 0.4]])), (2, 3, 2, np.array([[0.4, 0.6], [0.5, 0.9]])),
           (2, 4, 2, rng.random((2, 2))), (1, 2, T, rng.random((2, 2)))]
    psi = [[np.ones((2, 2)) for _ in range(T + 1)] for _ in range(go.ne)]
    for (i, j, t, m) in obs:
        for e in range(go.ne):
            if go.src[e] == i and go.dst[e] == j:
                psi[e][t] = psi[e][t] * m
            if go.src[e] == j and go.dst[e] == i:
                psi[e][t] = psi[e][t] * m.T
    h = rng.standard_normal(N)
    J = {frozenset(e): float(rng.standard_normal()) for e in und}
    wo, wd = [], []
    for i in range(N):
        Ji = [J[frozenset((i, int(j)))] for j in gd.neighbors(i)]
        wo.append([OF.GenericGlauberFactor(Ji, float(h[i]), 1.0)] * (T + 1))
        wd.append([M.GenericGlauberFactor(Ji, float(h[i]), 1.0)] * (T + 1))
    phi = [[np.array([0.15, 0.85]) if t == 0 else np.ones(2) for t in range(T + 1)] for _ in range(N)]
    phi[3][1] = np.array([1.0, 0.05])
    bo = O.MPBP(go, wo, [2] * N, T, phi=[[p.copy() for p in ph] for ph in phi], psi=[[m.copy() for m in ps] for ps in psi])
    bd = M.mpbp(gd, wd, [2] * N, T, phi=phi, psi=psi, dmax=16)
    O.iterate(bo, maxiter=5, trunc=OT.TruncThresh(0.0), tol=0.0)
    M.iterate_(bd, maxiter=5, svd_trunc=M.TruncBondThresh(16, 0.0), tol=0.0, shuffle_nodes=False)
    eb, ef, ep = compare(bo, bd)
    assert eb < TOL and ef < TOL and ep < TOL, (eb, ef, ep)
    p, Z, logZ = exact.exact_prob(bo)
    assert abs(-M.bethe_free_energy(bd) - logZ) < 1e-8
    assert np.allclose(np.array(M.beliefs(bd)), np.array(exact.exact_marginals(bo, p)), atol=1e-8)


@pytest.mark.parametrize("schedule", ["sequential", "parallel"])
def test_sirs_loopy_truncated_vs_oracle(schedule):
    # q = 3 on a loopy graph with an active bond cap: 3x3 message blocks through every kernel of the recursion
    T = 3
    und = [(0, 1), (1, 2), (0, 2), (2, 3)]
    N = 4
    kinds = [("sirs", (0.3 + 0.05 * i, 0.15, 0.2, 0.02)) for i in range(N)]
    phi = [[np.array([0.7, 0.25, 0.05]) if t == 0 else np.ones(3) for t in range(T + 1)] for _ in range(N)]
    phi[1][2] = np.array([0.1, 1.0, 0.3])
    bo, bd = build_pair(N, und, T, kinds, [3] * N, phi, dmax=5)
    tr = M.TruncBond(5)
    O.iterate(bo, maxiter=3, trunc=otrunc(tr), tol=0.0, schedule=schedule)
    M.iterate_(bd, maxiter=3, svd_trunc=tr, tol=0.0, shuffle_nodes=False, schedule=schedule)
    eb, ef, ep = compare(bo, bd)
    assert eb < TOL and ef < TOL and ep < TOL, (eb, ef, ep)


def test_trunc_bond_thresh_loopy_vs_oracle():
    # TruncBondThresh(d, eps): both the cap and the relative threshold bind at different sites
    T = 4
    und = [(0, 1), (0, 2), (1, 2), (2, 3), (3, 4), (4, 0)]
    N = 5
    kinds = [("glauber", (0.35, 0.05 * (i - 2), 1.0)) for i in range(N)]
    phi = [[np.array([0.3, 0.7]) if t == 0 else np.ones(2) for t in range(T + 1)] for _ in range(N)]
    bo, bd = build_pair(N, und, T, kinds, [2] * N, phi, dmax=6)
    tr = M.TruncBondThresh(5, 1e-3)
    O.iterate(bo, maxiter=3, trunc=otrunc(tr), tol=0.0)
    M.iterate_(bd, maxiter=3, svd_trunc=tr, tol=0.0, shuffle_nodes=False)
    bonds = [bd.get_message(e)[k].shape[0] for e in range(bd.E2) for k in range(T + 1)]
    assert max(bonds) == 5 and any(1 < b < 4 for b in bonds[2:])
    eb, ef, ep = compare(bo, bd)
    assert eb < TOL and ef < TOL and ep < TOL, (eb, ef, ep)


def test_observe_everything_free_energy_is_logprob_on_device():
    # test/glauber_small_tree.jl:74-86: every (i, t) observed with hard one-hot reweightings (exact zeros in phi): the
    # messages have zero blocks everywhere, -f_bethe must equal logprob(X) and the beliefs are the observed trajectory
    from oracle import factors as OF, exact
    T, N = 3, 5
    und = [(0, 1), (1, 2), (1, 3)]  # node 4 isolated
    rng = np.random.default_rng(5)
    h = rng.standard_normal(N)
    go = O.BiDiGraph(N, und)
    gd = M.IndexedBiDiGraph(N, und)
    wo = [[OF.HomogeneousGlauberFactor(1.0, float(h[i]), 1.0)] * (T + 1) for i in range(N)]
    wd = [[M.HomogeneousGlauberFactor(1.0, float(h[i]), 1.0)] * (T + 1) for i in range(N)]
    phi = [[np.array([0.75, 0.25]) if t == 0 else np.ones(2) for t in range(T + 1)] for _ in range(N)]
    bo = O.MPBP(go, wo, [2] * N, T, phi=[[p.copy() for p in ph] for ph in phi])
    X = exact.onesample(bo, rng)
    for i in range(N):
        for t in range(T + 1):
            bo.phi[i][t] = bo.phi[i][t] * (np.arange(1, 3) == X[i, t])
    bd = M.mpbp(gd, wd, [2] * N, T, phi=[[p.copy() for p in ph] for ph in bo.phi], dmax=10)
    tr = M.TruncBondThresh(10, 0.0)
    O.iterate(bo, maxiter=6, trunc=OT.TruncBondThresh(10), tol=0.0)
    M.iterate_(bd, maxiter=6, svd_trunc=tr, tol=0.0, shuffle_nodes=False)
    lp = exact.logprob(bo, X)
    assert abs(-M.bethe_free_energy(bd) - lp) < 1e-8 and abs(-O.bethe_free_energy(bo) - lp) < 1e-8
    b = np.array(M.beliefs(bd))
    assert np.allclose(b, (np.arange(1, 3)[None, None, :] == X[:, :, None]), atol=1e-10)
    eb, ef, ep = compare(bo, bd)
    assert eb < TOL and ef < TOL and ep < TOL, (eb, ef, ep)


def test_damped_factor_vs_oracle_and_exact():
    # DampedFactor(w, p) (src/recursive_bp_factor.jl:183-206, test/glauber_small_tree.jl:88-131)
    from oracle import factors as OF, exact
    T, N = 2, 5
    und = [(0, 1), (1, 2), (1, 3), (3, 4)]
    rng = np.random.default_rng(9)
    h = rng.standard_normal(N)
    go = O.BiDiGraph(N, und)
    gd = M.IndexedBiDiGraph(N, und)
    wo = [[OF.DampedFactor(OF.HomogeneousGlauberFactor(1.0, float(h[i]), 1.0), 0.2)] * (T + 1) for i in range(N)]
    wd = [[M.DampedFactor(M.HomogeneousGlauberFactor(1.0, float(h[i]), 1.0), 0.2)] * (T + 1) for i in range(N)]
    phi = [[np.array([0.75, 0.25]) if t == 0 else np.ones(2) for t in range(T + 1)] for _ in range(N)]
    phi[2][1] = np.array([1.0, 1e-3])
    phi[4][2] = np.array([1e-3, 1.0])
    bo = O.MPBP(go, wo, [2] * N, T, phi=[[p.copy() for p in ph] for ph in phi])
    bd = M.mpbp(gd, wd, [2] * N, T, phi=phi, dmax=10)
    O.iterate(bo, maxiter=8, trunc=OT.TruncBondThresh(10), tol=0.0)
    M.iterate_(bd, maxiter=8, svd_trunc=M.TruncBondThresh(10, 0.0), tol=0.0, shuffle_nodes=False)
    eb, ef, ep = compare(bo, bd)
    assert eb < TOL and ef < TOL and ep < TOL, (eb, ef, ep)
    p, Z, logZ = exact.exact_prob(bo)
    assert abs(-M.bethe_free_energy(bd) - logZ) < 1e-8
    assert np.allclose(np.array(M.beliefs(bd)), np.array(exact.exact_marginals(bo, p)), atol=1e-8)


@pytest.mark.parametrize("schedule", ["sequential", "parallel"])
def test_two_time_marginals_and_autocorrelations_vs_oracle(schedule):
    # beliefs_tu / autocorrelations / autocovariances (src/mpbp.jl:239-255,289-296) on a loopy graph with an isolated
    # node and an active truncation: the device computes them with the beliefs (option "twovar")
    T = 4
    und = [(0, 1), (0, 2), (1, 2), (2, 3)]
    N = 5
    kinds = [("glauber", (0.4, 0.1 * (i - 2), 1.0)) for i in range(N)]
    phi = [[np.array([0.3, 0.7]) if t == 0 else np.ones(2) for t in range(T + 1)] for _ in range(N)]
    phi[1][2] = np.array([1.0, 0.2])
    bo, bd = build_pair(N, und, T, kinds, [2] * N, phi, dmax=4)
    bd.set_option("twovar", T)
    tr = M.TruncBond(4)
    O.iterate(bo, maxiter=3, trunc=otrunc(tr), tol=0.0, schedule=schedule)
    M.iterate_(bd, maxiter=3, svd_trunc=tr, tol=0.0, shuffle_nodes=False, schedule=schedule)
    tvo, tvd = O.beliefs_tu(bo), M.beliefs_tu(bd)
    for i in range(N):
        for t in range(T + 1):
            for u in range(T + 1):
                if u <= t:
                    assert tvd[i][t][u] is None
                else:
                    assert np.max(np.abs(np.asarray(tvo[i][t][u]).reshape(2, 2) - tvd[i][t][u])) < TOL, (i, t, u)
    f = lambda x, i: 2 * x - 3
    assert np.max(np.abs(np.array(O.autocorrelations(bo, f)) - np.array(M.autocorrelations(f, bd)))) < TOL
    assert np.max(np.abs(np.array(O.autocovariances(bo, f)) - np.array(M.autocovariances(f, bd)))) < TOL
    eb, ef, ep = compare(bo, bd)  # the usual read-outs are untouched by the option
    assert eb < TOL and ef < TOL and ep < TOL, (eb, ef, ep)


def test_two_time_marginals_sirs_tree_exact_maxdist_and_loud_when_off():
    from oracle import exact
    T = 2
    und = [(0, 1), (1, 2)]
    N = 3
    kinds = [("sirs", (0.4, 0.15, 0.2, 0.05))] * N
    phi = [[np.array([0.7, 0.3, 0.0]) if t == 0 else np.ones(3) for t in range(T + 1)] for _ in range(N)]
    phi[0][2] = np.array([0.1, 1.0, 0.3])
    bo, bd = build_pair(N, und, T, kinds, [3] * N, phi, dmax=9)
    with pytest.raises(M.MPBPError):
        M.beliefs_tu(bd)  # not switched on
    bd.set_option("twovar", T)
    tr = M.TruncBondThresh(9, 0.0)
    O.iterate(bo, maxiter=5, trunc=OT.TruncThresh(0.0), tol=0.0)
    M.iterate_(bd, maxiter=5, svd_trunc=tr, tol=0.0, shuffle_nodes=False)
    p, Z, _ = exact.exact_prob(bo)
    f = lambda x, i: x - 1
    rex = exact.exact_autocorrelations(bo, p, f)
    assert np.max(np.abs(np.array(M.autocorrelations(f, bd)) - np.array(rex))) < TOL
    # maxdist = 1: only adjacent times are computed
    bd.set_option("twovar", 1)
    M.iterate_(bd, maxiter=1, svd_trunc=tr, tol=0.0, shuffle_nodes=False)
    tv = M.beliefs_tu(bd)
    assert tv[1][0][1] is not None and tv[1][1][2] is not None and tv[1][0][2] is None


def test_alternate_marginals_vs_oracle_and_exact():
    # alternate_marginals (src/mpbp.jl:270-280): p(x_i^t, x_j^{t+1}) per directed edge; tree with psi (exact available)
    # and a loopy truncated SIRS case (q = 3) against the oracle
    from oracle import exact
    T, N = 2, 5
    rng = np.random.default_rng(2)
    und = [(0, 1), (1, 2), (1, 3)]
    go = O.BiDiGraph(N, und)
    psi = [None] * go.ne
    for e in range(go.ne):
        if psi[e] is None:
            ps = [0.5 + rng.random((2, 2)) for _ in range(T + 1)]
            psi[e] = ps
            psi[go.rev[e]] = [p.T.copy() for p in ps]
    kinds = [("glauber", (1.0, float(rng.standard_normal()), 1.0)) for _ in range(N)]
    phi = [[np.array([0.75, 0.25]) if t == 0 else np.ones(2) for t in range(T + 1)] for _ in range(N)]
    phi[2][1] = np.array([1.0, 0.1])
    bo, bd = build_pair(N, und, T, kinds, [2] * N, phi, psi=psi, dmax=10)
    O.iterate(bo, maxiter=5, trunc=OT.TruncBondThresh(10), tol=0.0)
    M.iterate_(bd, maxiter=5, svd_trunc=M.TruncBondThresh(10, 0.0), tol=0.0, shuffle_nodes=False)
    p, Z, _ = exact.exact_prob(bo)
    amd, amo, amx = M.alternate_marginals(bd), O.alternate_marginals(bo), exact.exact_alternate_marginals(bo, p)
    assert np.max(np.abs(np.array(amd) - np.array(amo))) < TOL and np.max(np.abs(np.array(amd) - np.array(amx))) < TOL
    # loopy, q = 3, truncation active
    T = 3
    und = [(0, 1), (1, 2), (0, 2), (2, 3)]
    N = 4
    kinds = [("sirs", (0.3 + 0.05 * i, 0.15, 0.2, 0.02)) for i in range(N)]
    phi = [[np.array([0.7, 0.25, 0.05]) if t == 0 else np.ones(3) for t in range(T + 1)] for _ in range(N)]
    bo, bd = build_pair(N, und, T, kinds, [3] * N, phi, dmax=5)
    tr = M.TruncBond(5)
    O.iterate(bo, maxiter=3, trunc=otrunc(tr), tol=0.0)
    M.iterate_(bd, maxiter=3, svd_trunc=tr, tol=0.0, shuffle_nodes=False)
    amd, amo = M.alternate_marginals(bd), O.alternate_marginals(bo)
    assert np.max(np.abs(np.array(amd) - np.array(amo))) < TOL
    f = lambda x: x - 1
    co = [[float(np.array([f(x + 1) for x in range(3)]) @ p @ np.array([f(x + 1) for x in range(3)])) for p in am] for am in amo]
    assert np.max(np.abs(np.array(M.alternate_correlations(f, bd)) - np.array(co))) < TOL


def test_k4_bond16_full_size_paths_vs_oracle():
    # complete graph K4, T=5, TruncBond(16): at the middle cuts D = 256 -> H=64 flat-tree QR, TSQR split (few ops per
    # launch) and the subspace-iteration SVD (d~X = 96..128 > 48) all run inside a real BP iteration
    T, N, d = 5, 4, 16
    und = [(a, b) for a in range(N) for b in range(a + 1, N)]
    kinds = [("glauber", (0.4, 0.05 * (i + 1), 1.0)) for i in range(N)]
    phi = [[np.array([0.2, 0.8]) if t == 0 else np.ones(2) for t in range(T + 1)] for _ in range(N)]
    phi[2][3] = np.array([0.6, 0.3])
    bo, bd = build_pair(N, und, T, kinds, [2] * N, phi, dmax=d)
    tr = M.TruncBond(d)
    O.iterate(bo, maxiter=3, trunc=otrunc(tr), tol=0.0, schedule="parallel")
    M.iterate_(bd, maxiter=3, svd_trunc=tr, tol=0.0, shuffle_nodes=False, schedule="parallel")
    assert max(bd.get_message(0)[k].shape[0] for k in range(T + 1)) == d  # the cap is reached
    assert bd.counters()["svd_calls"] > 0  # the subspace path ran
    eb, ef, ep = compare(bo, bd)
    assert eb < TOL and ef < TOL and ep < TOL, (eb, ef, ep)


@pytest.mark.parametrize("schedule", ["sequential", "parallel"])
def test_damping_vs_oracle(schedule):
    # set_msg! with damp > 0 (src/recursive_bp_factor.jl:168-179): TT sum of new and old message + compress! + normalize!
    T = 4
    und = [(0, 1), (0, 2), (1, 2), (2, 3), (3, 4)]
    N = 5
    kinds = [("sis", (0.2 + 0.02 * i, 0.12, 0.01)) for i in range(N)]
    phi = [[np.array([0.87, 0.13]) if t == 0 else np.ones(2) for t in range(T + 1)] for _ in range(N)]
    phi[2][3] = np.array([0.2, 1.0])
    bo, bd = build_pair(N, und, T, kinds, [2] * N, phi, dmax=4)
    tr = M.TruncBond(3)
    O.iterate(bo, maxiter=4, trunc=otrunc(tr), tol=0.0, damp=0.3, schedule=schedule)
    M.iterate_(bd, maxiter=4, svd_trunc=tr, tol=0.0, damp=0.3, shuffle_nodes=False, schedule=schedule)
    eb, ef, ep = compare(bo, bd)
    assert eb < TOL and ef < TOL and ep < TOL, (eb, ef, ep)


def test_damping_infinite_graph_vs_oracle():
    # reference test/glauber_infinite_graph.jl:22 iterates the infinite graph with damp = 0.1: the k recomputed messages
    # are damped one after the other against the evolving bp.mu[1]
    from oracle import factors as OF
    T, k, m0 = 3, 3, 0.5
    phi = [np.array([(1 + m0) / 2, (1 - m0) / 2]) if t == 0 else np.ones(2) for t in range(T + 1)]
    phi[1] = np.array([0.4, 0.6])
    phi[-1] = np.array([0.95, 0.05])
    bo = O.mpbp_infinite_graph(k, [OF.HomogeneousGlauberFactor(1.0, 0.1, 1.0)] * (T + 1), 2, [p.copy() for p in phi])
    bd = M.mpbp_infinite_graph(k, [M.HomogeneousGlauberFactor(1.0, 0.1, 1.0)] * (T + 1), 2, phi, dmax=6)
    O.iterate(bo, maxiter=5, trunc=OT.TruncBond(6), tol=0.0, damp=0.1)
    M.iterate_(bd, maxiter=5, svd_trunc=M.TruncBond(6), tol=0.0, damp=0.1)
    assert np.max(np.abs(np.array(O.beliefs(bo)[0]) - M.beliefs(bd)[0])) < TOL
    assert abs(O.bethe_free_energy(bo) - M.bethe_free_energy(bd)) < TOL


def test_isolated_node_and_leaf_vs_oracle():
    # degree-0 node (cavity of an empty neighbourhood) next to a 2-chain
    T = 3
    und = [(0, 1)]
    N = 3
    kinds = [("glauber", (0.3, 0.2, 1.0)), ("glauber", (0.3, -0.1, 1.0)), ("glauber", (0.3, 0.4, 1.0))]
    phi = [[np.array([0.3, 0.7]) if t == 0 else np.ones(2) for t in range(T + 1)] for _ in range(N)]
    phi[2][2] = np.array([0.9, 0.2])
    bo, bd = build_pair(N, und, T, kinds, [2] * N, phi, dmax=4)
    tr = M.TruncBond(4)
    O.iterate(bo, maxiter=2, trunc=otrunc(tr), tol=0.0)
    M.iterate_(bd, maxiter=2, svd_trunc=tr, tol=0.0, shuffle_nodes=False)
    eb, ef, ep = compare(bo, bd)
    assert eb < TOL and ef < TOL and ep < TOL, (eb, ef, ep)


def test_pack_unpack_roundtrip():
    # multi-GPU plumbing: device-side pack/unpack of message slots must round-trip bit-exactly
    import torch
    T = 3
    und = [(0, 1), (1, 2), (0, 2)]
    kinds = [("sis", (0.2, 0.1, 0.0))] * 3
    phi = [[np.array([0.8, 0.2]) if t == 0 else np.ones(2) for t in range(T + 1)] for _ in range(3)]
    bo, bd = build_pair(3, und, T, kinds, [2] * 3, phi, dmax=4)
    M.iterate_(bd, maxiter=2, svd_trunc=M.TruncBond(3), tol=0.0, shuffle_nodes=False)
    before = [bd.get_message(e) for e in range(bd.E2)]
    sb = int(_lib.lib().mpbp_message_slot_bytes(bd._h))
    edges = np.arange(bd.E2, dtype=np.int64)
    buf = torch.empty(bd.E2 * sb, dtype=torch.uint8, device="cuda:0")
    _lib.check(_lib.lib().mpbp_pack_messages_dev(bd._h, bd.E2, edges.ctypes.data_as(_lib.c_i64p), buf.data_ptr()))
    M.reset_messages_(bd)
    perm = edges[::-1].copy()  # unpack into the reversed edge order, then back
    _lib.check(_lib.lib().mpbp_unpack_messages_dev(bd._h, bd.E2, edges.ctypes.data_as(_lib.c_i64p), buf.data_ptr()))
    after = [bd.get_message(e) for e in range(bd.E2)]
    for a, b in zip(before, after):
        assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_errors_are_loud():
    T = 1
    g = M.IndexedBiDiGraph(2, [(0, 1)])
    w = [[M.SISFactor(0.1, 0.1)] * (T + 1)] * 2
    bp = M.mpbp(g, w, [2, 2], T, dmax=2)
    with pytest.raises(M.MPBPError):
        M.iterate_(bp, maxiter=1, svd_trunc=M.TruncBond(5))  # exceeds dmax
    with pytest.raises(M.MPBPError):
        M.iterate_(bp, maxiter=1, svd_trunc=M.TruncBond(2), damp=1.5)  # invalid damping -> loud
