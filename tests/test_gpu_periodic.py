"""Periodic-in-time MPBP on the device (/root/reference/src/mpems.jl:96-155,
/root/reference/src/recursive_bp_factor.jl:89-101, /root/reference/src/mpbp.jl:399-409) against oracle/periodic.py and
brute-force enumeration of the wrapped dynamics, on the structures of /root/reference/test/periodic.jl.  Tolerance 1e-8."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import mpbp_b200 as M
from oracle import factors as OF, mpbp as O, periodic as P, tt

TOL = 1e-8


def _tree_case():
    rng = np.random.default_rng(111)
    T, N = 2, 5
    und = [(0, 1), (1, 2), (1, 3)]  # node 4 is isolated, like J[5,:] = 0 of test/periodic.jl:6-10
    go = O.BiDiGraph(N, und)
    h = rng.standard_normal(N)
    obs = [(0, 1, 1, np.array([[0.1, 0.9], [0.3, 0.4]])), (1, 3, 2, np.array([[0.4, 0.6], [0.5, 0.9]])), (1, 2, T, rng.random((2, 2)))]
    psi = [[np.ones((2, 2)) for _ in range(T + 1)] for _ in range(go.ne)]
    for (i, j, t, m) in obs:
        for e in range(go.ne):
            if go.src[e] == i and go.dst[e] == j:
                psi[e][t] = psi[e][t] * m
            if go.src[e] == j and go.dst[e] == i:
                psi[e][t] = psi[e][t] * m.T
    phi = [[np.ones(2) for _ in range(T + 1)] for _ in range(N)]
    for i in range(N):
        phi[i][0] = np.array([0.75, 0.25])
    phi[2][1] = np.array([1.0, 0.1])
    phi[0][2] = np.array([0.3, 1.0])
    return go, und, h, T, N, phi, psi


def _pair(go, und, N, T, q, wo, wd, phi, psi, dmax):
    gd = M.IndexedBiDiGraph(N, und)
    assert list(gd.src) == go.src and list(gd.dst) == go.dst and list(gd.rev) == go.rev
    bo = P.PeriodicMPBP(go, wo, [q] * N, T, phi=[[p.copy() for p in ph] for ph in phi], psi=[[p.copy() for p in ps] for ps in psi])
    bd = M.periodic_mpbp(gd, wd, [q] * N, T, phi=phi, psi=psi, dmax=dmax)
    return bo, bd


def _exact_marginals(p, N, L):
    return [[p.sum(axis=tuple(a for a in range(N * L) if a != i * L + t)) for t in range(L)] for i in range(N)]


@pytest.mark.parametrize("schedule", ["sequential", "parallel"])
def test_periodic_glauber_tree_vs_oracle_and_exact(schedule):
    go, und, h, T, N, phi, psi = _tree_case()
    wo = [[OF.HomogeneousGlauberFactor(1.0, float(h[i]), 1.0)] * (T + 1) for i in range(N)]
    wd = [[M.HomogeneousGlauberFactor(1.0, float(h[i]), 1.0)] * (T + 1) for i in range(N)]
    bo, bd = _pair(go, und, N, T, 2, wo, wd, phi, psi, dmax=10)
    P.iterate(bo, maxiter=8, trunc=tt.TruncBondThresh(10))
    bd.set_option("twovar", T)  # two-time marginals of every belief, all time pairs
    iters, cb = M.iterate_(bd, maxiter=8, svd_trunc=M.TruncBondThresh(10), tol=0.0, shuffle_nodes=False, schedule=schedule)
    assert iters == 8
    L = T + 1
    p, logZ = P.exact_prob(bo)
    b_d = M.beliefs(bd)
    be = _exact_marginals(p, N, L)
    for i in range(N):
        assert np.allclose(np.array(b_d[i]), np.array(be[i]), atol=TOL)
    assert abs(-M.bethe_free_energy(bd) - logZ) < TOL
    if schedule == "sequential":  # same sweep as the oracle: every intermediate agrees, not only the fixed point
        assert np.allclose(np.array(P.beliefs(bo)), np.array([np.array(b) for b in b_d]), atol=TOL)
        assert np.allclose(bo.f, M.api.free_energy_contributions(bd), atol=TOL)
    # two-time marginals / autocorrelations (test/periodic.jl:43-47, 62-63) against brute force
    btu = M.beliefs_tu(bd)
    spin = lambda x, i: 3 - 2 * x  # potts2spin on states numbered from 1
    r_bp = M.autocorrelations(spin, bd)
    for i in range(N):
        for t in range(L):
            for u in range(t + 1, L):
                ex = p.sum(axis=tuple(a for a in range(N * L) if a not in (i * L + t, i * L + u)))
                assert np.allclose(btu[i][t][u], ex, atol=TOL)
                r_ex = sum(spin(a + 1, i) * spin(b + 1, i) * ex[a, b] for a in range(2) for b in range(2))
                assert abs(r_bp[i][t, u] - r_ex) < TOL
    # pair beliefs and their free-energy weights (test/periodic.jl:49-60)
    pb_d, lz_d = M.pair_beliefs(bd)
    pb_o, lz_o = P.pair_beliefs(bo)
    for e in range(go.ne):
        i, j = go.src[e], go.dst[e]
        for t in range(L):
            ex = p.sum(axis=tuple(a for a in range(N * L) if a not in (i * L + t, j * L + t)))
            assert np.allclose(np.array(pb_d[e][t]), ex if i < j else ex.T, atol=TOL)
    assert np.allclose(lz_d, lz_o, atol=TOL)
    # messages: ring-normalised, every bond (the closing one included) under the cap, same function as the oracle's
    rng = np.random.default_rng(0)
    for e in range(go.ne):
        A = tt.TT(bd.get_message(e))
        assert abs(P.lognormalization(A)) < 1e-9
        assert A[0].shape[0] == A[-1].shape[1] and max(max(a.shape[:2]) for a in A) <= 10
        if schedule == "sequential":
            for _ in range(4):
                x = [tuple(rng.integers(0, 2, size=2)) for _ in range(L)]
                assert abs(P.evaluate(A, x) - P.evaluate(bo.mu[e], x)) < TOL


def test_periodic_reference_flow_glauber_model_observations_autocovariances():
    # the whole flow of /root/reference/test/periodic.jl:1-75 through the model constructors: Ising -> Glauber(psi) ->
    # periodic_mpbp(model) -> draw_node_observations!(bp, N) (hard one-hot phi) -> iterate!(maxiter=20, TruncBondThresh(10)) ->
    # Z, beliefs, autocorrelations, autocovariances, pair beliefs against brute force (oracle.periodic.exact_prob)
    rng = np.random.default_rng(111)
    T, N = 2, 5
    L = T + 1
    und = [(0, 1), (1, 2), (1, 3)]
    g = M.IndexedBiDiGraph(N, und)
    h = rng.standard_normal(N)
    ising = M.Ising(g, J=np.ones(len(g.undirected)), h=h, beta=1.0)
    O_ = {(0, 1): (1, np.array([[0.1, 0.9], [0.3, 0.4]])), (1, 3): (2, np.array([[0.4, 0.6], [0.5, 0.9]])), (1, 2): (T, rng.random((2, 2)))}
    psi_und = []
    for (a, b) in g.undirected.tolist():
        ps = [np.ones((2, 2)) for _ in range(L)]
        if (a, b) in O_:
            t, m = O_[(a, b)]
            ps[t] = ps[t] * m
        psi_und.append(ps)
    gl = M.Glauber(ising, T, psi=psi_und)
    for i in range(N):
        gl.phi[i][0] = gl.phi[i][0] * np.array([0.75, 0.25])
    bp = M.periodic_mpbp(gl, dmax=10)
    X, observed = M.draw_node_observations_(bp, N, rng=5)
    assert len(observed) == N
    bp.set_option("twovar", T)
    iters, cb = M.iterate_(bp, maxiter=20, svd_trunc=M.TruncBondThresh(10), tol=0.0, shuffle_nodes=False)
    # brute force on the same inputs
    go = O.BiDiGraph(N, und)
    assert list(g.src) == go.src and list(g.dst) == go.dst
    wo = [[OF.HomogeneousGlauberFactor(1.0, float(h[i]), 1.0)] * L for i in range(N)]
    psi_dir = [[np.asarray(p) if g.src[e] < g.dst[e] else np.asarray(p).T for p in psi_und[g.und_of[e]]] for e in range(g.ne)]
    bo = P.PeriodicMPBP(go, wo, [2] * N, T, phi=[[np.asarray(p, dtype=float).copy() for p in ph] for ph in bp.phi], psi=psi_dir)
    p, logZ = P.exact_prob(bo)
    assert abs(-M.bethe_free_energy(bp) - logZ) < TOL                                  # Z_exact ~ Z_bp
    be = _exact_marginals(p, N, L)
    b = M.beliefs(bp)
    for i in range(N):
        assert np.allclose(np.array(b[i]), np.array(be[i]), atol=TOL)                  # p_ex ~ p_bp
    f = lambda x, i: 2 * x - 3                                                         # test/periodic.jl:41
    r_bp, c_bp, mu = M.autocorrelations(f, bp), M.autocovariances(f, bp), M.means(f, bp)
    fx = np.array([f(1, 0), f(2, 0)], dtype=float)
    for i in range(N):
        for t in range(L):
            for u in range(t + 1, L):
                ex = p.sum(axis=tuple(a for a in range(N * L) if a not in (i * L + t, i * L + u)))
                r_ex = fx @ ex @ fx
                assert abs(r_bp[i][t, u] - r_ex) < TOL                                 # r_bp ~ r_exact
                assert abs(c_bp[i][t, u] - (r_ex - (fx @ be[i][t]) * (fx @ be[i][u]))) < TOL  # c_bp ~ c_exact
    pb, _ = M.pair_beliefs(bp)
    for e in range(g.ne):
        i, j = int(g.src[e]), int(g.dst[e])
        for t in range(L):
            ex = p.sum(axis=tuple(a for a in range(N * L) if a not in (i * L + t, j * L + t)))
            assert np.allclose(np.array(pb[e][t]), ex if i < j else ex.T, atol=TOL)


def test_periodic_sis_tree_truncthresh0_vs_exact():
    # T = 3, TruncThresh(0.0): the bonds are whatever the exact ranks are (dmax is only a capacity)
    T, N = 3, 3
    und = [(0, 1), (1, 2)]
    go = O.BiDiGraph(N, und)
    wo = [[OF.SISFactor(0.3, 0.25, 0.05)] * (T + 1) for _ in range(N)]
    wd = [[M.SISFactor(0.3, 0.25, 0.05)] * (T + 1) for _ in range(N)]
    phi = [[np.ones(2) for _ in range(T + 1)] for _ in range(N)]
    phi[0][1] = np.array([0.2, 1.0])
    phi[2][3] = np.array([1.0, 0.4])
    psi = [[np.ones((2, 2)) for _ in range(T + 1)] for _ in range(go.ne)]
    bo, bd = _pair(go, und, N, T, 2, wo, wd, phi, psi, dmax=12)
    P.iterate(bo, maxiter=6, trunc=tt.TruncThresh(0.0))
    M.iterate_(bd, maxiter=6, svd_trunc=M.TruncThresh(0.0), tol=0.0, shuffle_nodes=False)
    p, logZ = P.exact_prob(bo)
    L = T + 1
    be = _exact_marginals(p, N, L)
    b_d = M.beliefs(bd)
    for i in range(N):
        assert np.allclose(np.array(b_d[i]), np.array(be[i]), atol=TOL)
    assert abs(-M.bethe_free_energy(bd) - logZ) < TOL
    assert np.allclose(bo.f, M.api.free_energy_contributions(bd), atol=TOL)


def test_periodic_set_get_message_roundtrip_and_one_node_update():
    # random ring messages with a non-trivial closing bond go in through mpbp_set_message; ONE node update must equal the oracle's
    rng = np.random.default_rng(7)
    T, N, d = 3, 4, 3
    und = [(0, 1), (0, 2), (0, 3)]
    go = O.BiDiGraph(N, und)
    wo = [[OF.SISFactor(0.3, 0.2, 0.05)] * (T + 1) for _ in range(N)]
    wd = [[M.SISFactor(0.3, 0.2, 0.05)] * (T + 1) for _ in range(N)]
    phi = [[0.3 + rng.random(2) for _ in range(T + 1)] for _ in range(N)]
    psi = [[0.5 + rng.random((2, 2)) for _ in range(T + 1)] for _ in range(go.ne)]
    for e in range(go.ne):  # psi of the two directions of an edge are transposes of each other
        r = go.rev[e]
        if e < r:
            psi[r] = [p.T.copy() for p in psi[e]]
    bo, bd = _pair(go, und, N, T, 2, wo, wd, phi, psi, dmax=4)
    for e in go.in_edges[0]:
        A = tt.TT([rng.random((d, d, 2, 2)) for _ in range(T + 1)])
        P.normalize(A)
        bo.mu[e] = A.copy()
        bd.set_message(e, [a.copy() for a in A])
        B = bd.get_message(e)
        assert all(np.allclose(a, b, atol=1e-14) for a, b in zip(A, B))
    P.onebpiter(bo, 0, tt.TruncBond(4))
    M.iterate_(bd, maxiter=1, svd_trunc=M.TruncBond(4), tol=0.0, nodes=[0], shuffle_nodes=False)
    assert np.allclose(np.array(P.marginals(bo.b[0])), np.array(M.beliefs(bd)[0]), atol=TOL)
    assert abs(bo.f[0] - M.api.free_energy_contributions(bd)[0]) < TOL
    for e in go.out_edges[0]:
        A = tt.TT(bd.get_message(e))
        for _ in range(6):
            x = [tuple(rng.integers(0, 2, size=2)) for _ in range(T + 1)]
            assert abs(P.evaluate(A, x) - P.evaluate(bo.mu[e], x)) < TOL


def test_periodic_loopy_sirs_binding_truncation_time_dependent_factors():
    rng = np.random.default_rng(5)
    T, N = 2, 4
    und = [(0, 1), (1, 2), (2, 0), (2, 3)]
    go = O.BiDiGraph(N, und)
    par = [[(0.3 + 0.1 * t, 0.2, 0.15, 0.05) for t in range(T + 1)] for _ in range(N)]
    wo = [[OF.SIRSFactor(*par[i][t]) for t in range(T + 1)] for i in range(N)]
    wd = [[M.SIRSFactor(*par[i][t]) for t in range(T + 1)] for i in range(N)]
    phi = [[0.2 + rng.random(3) for _ in range(T + 1)] for _ in range(N)]
    psi = [[np.ones((3, 3)) for _ in range(T + 1)] for _ in range(go.ne)]
    bo, bd = _pair(go, und, N, T, 3, wo, wd, phi, psi, dmax=3)
    P.iterate(bo, maxiter=3, trunc=tt.TruncBond(3))
    M.iterate_(bd, maxiter=3, svd_trunc=M.TruncBond(3), tol=0.0, shuffle_nodes=False)
    assert np.allclose(np.array(P.beliefs(bo)), np.array([np.array(b) for b in M.beliefs(bd)]), atol=TOL)
    assert np.allclose(bo.f, M.api.free_energy_contributions(bd), atol=TOL)


def test_periodic_infinite_graph_vs_oracle_and_complete_graph():
    # test/periodic.jl:78-118 (periodic_mpbp_infinite_graph against the complete graph of k+1 nodes).  On a loopy graph the
    # exact ring bonds double with every iteration (2, 4, 8, 32, ...), so every longer run is bound by the truncation, where
    # the ring sweeps of TensorTrains.jl are unpinned (oracle/periodic.py header): compared here while the bonds fit.
    T, k, d = 2, 3, 8
    L = T + 1
    fac = (0.4, 0.1, 1.0)
    phi = [np.array([0.75, 0.25]), np.array([0.4, 0.6]), np.array([0.95, 0.05])]
    wd = [M.HomogeneousGlauberFactor(*fac)] * L
    bo = P.PeriodicMPBP(O.InfiniteRegularGraph(k), [[OF.HomogeneousGlauberFactor(*fac)] * L], [2], T, phi=[[p.copy() for p in phi]])
    bp = M.periodic_mpbp_infinite_graph(k, wd, 2, phi=[p.copy() for p in phi], dmax=d)
    for it in range(3):
        P.iterate(bo, maxiter=1, trunc=tt.TruncBondThresh(d, 1e-12))
        M.iterate_(bp, maxiter=1, svd_trunc=M.TruncBondThresh(d, 1e-12), tol=0.0, shuffle_nodes=False)
        assert np.allclose(np.array(P.beliefs(bo)[0]), np.array(M.beliefs(bp)[0]), atol=TOL)
        assert abs(bo.f[0] - M.api.free_energy_contributions(bp)[0]) < TOL
        if it == 1:
            b_inf2 = np.array(M.beliefs(bp)[0])
            pb_inf2 = np.array(M.pair_beliefs(bp)[0][0])
    # Jacobi BP on the complete graph K_{k+1} with identical nodes = the infinite-graph iteration, message by message
    N = k + 1
    g = M.IndexedBiDiGraph(N, [(i, j) for i in range(N) for j in range(i + 1, N)])
    be = M.periodic_mpbp(g, [list(wd) for _ in range(N)], [2] * N, T, phi=[[p.copy() for p in phi] for _ in range(N)], dmax=d)
    M.iterate_(be, maxiter=2, svd_trunc=M.TruncBondThresh(d, 1e-12), tol=0.0, schedule="parallel")
    for i in range(N):
        assert np.allclose(np.array(M.beliefs(be)[i]), b_inf2, atol=TOL)
    assert np.allclose(np.array(M.pair_beliefs(be)[0][0]), pb_inf2, atol=TOL)


def test_periodic_damped_sweep_vs_oracle():
    # set_msg! with damp > 0 on ring messages (block-diagonal sum, compress!, normalize!).  Damped ring messages grow by the
    # bond of the old message at every sweep, so only the first sweeps stay below the cap (non-binding, exactly comparable)
    go, und, h, T, N, phi, psi = _tree_case()
    wo = [[OF.HomogeneousGlauberFactor(1.0, float(h[i]), 1.0)] * (T + 1) for i in range(N)]
    wd = [[M.HomogeneousGlauberFactor(1.0, float(h[i]), 1.0)] * (T + 1) for i in range(N)]
    bo, bd = _pair(go, und, N, T, 2, wo, wd, phi, psi, dmax=16)
    P.iterate(bo, maxiter=1, trunc=tt.TruncBondThresh(16, 1e-13), damp=0.3)
    M.iterate_(bd, maxiter=1, svd_trunc=M.TruncBondThresh(16, 1e-13), tol=0.0, shuffle_nodes=False, damp=0.3)
    assert np.allclose(np.array(P.beliefs(bo)), np.array([np.array(b) for b in M.beliefs(bd)]), atol=TOL)
    assert np.allclose(bo.f, M.api.free_energy_contributions(bd), atol=TOL)
    rng = np.random.default_rng(1)
    for e in range(go.ne):
        A = tt.TT(bd.get_message(e))
        assert abs(P.lognormalization(A)) < 1e-9
        assert [a.shape[0] for a in A] == [a.shape[0] for a in bo.mu[e]]
        for _ in range(4):
            x = [tuple(rng.integers(0, 2, size=2)) for _ in range(T + 1)]
            assert abs(P.evaluate(A, x) - P.evaluate(bo.mu[e], x)) < TOL


def test_periodic_path_fails_loudly_where_it_is_not_defined():
    T, N = 2, 3
    g = M.IndexedBiDiGraph(N, [(0, 1), (1, 2)])
    w = [[M.SISFactor(0.3, 0.2)] * (T + 1) for _ in range(N)]
    bp = M.periodic_mpbp(g, w, [2] * N, T, dmax=4)
    with pytest.raises(M.MPBPError):
        M.alternate_marginals(bp)
    # the forward sampler is an input generator: like the reference's onesample! it ignores the wrap-around factor, so a
    # periodic state draws the same trajectory as the open one (test/periodic.jl:31 draws observations from a periodic bp)
    bo = M.mpbp(g, w, [2] * N, T, dmax=4)
    assert np.array_equal(M.sample_prior(bp, 7), M.sample_prior(bo, 7))
    with pytest.raises(M.MPBPError):  # a truncation that lets a bond outgrow dmax is an error, never a silent cut
        M.iterate_(M.periodic_mpbp(g, [[M.SISFactor(0.3, 0.2)] * (T + 1) for _ in range(N)], [2] * N, T, phi=[[np.array([0.3, 0.7])] * (T + 1)] * N, dmax=1),
                   maxiter=2, svd_trunc=M.TruncThresh(1e-12), tol=0.0, shuffle_nodes=False)
    with pytest.raises(M.MPBPError):
        M.periodic_mpbp(g, w, [2] * N, T, dmax=20)
