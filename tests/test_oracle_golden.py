"""Pins the oracle (CPU restatement) against the reference's own known-answer test and against
brute-force enumeration, mirroring the reference's test strategy."""
import numpy as np
import pytest

from oracle import exact, factors as F, mpbp as O, tt


# /root/reference/test/sis_infinite_graph.jl:21-29 -- the only numeric literals in the reference test-suite
SIS_INFINITE_GOLDEN = np.array([
    [0.9000000001671186, 0.0999999998328814],
    [0.8932690998131098, 0.10673090018689023],
    [0.8899420329322244, 0.11005796706777556],
    [0.8884643888492034, 0.11153561115079656],
    [0.8880305235706524, 0.1119694764293476],
    [0.8882121515614524, 0.11178784843854758],
    [0.8887717202217936, 0.1112282797782064],
])


def test_sis_infinite_graph_golden():
    # inputs: /root/reference/test/sis_infinite_graph.jl:3-18
    T, k, gamma, lam, rho = 6, 3, 0.1, 0.1, 0.2
    w = [F.SISFactor(lam, rho) for _ in range(T + 1)]
    phi = [np.array([1 - gamma, gamma]) if t == 0 else np.ones(2) for t in range(T + 1)]
    bp = O.mpbp_infinite_graph(k, w, 2, phi)
    iters, _ = O.iterate(bp, maxiter=200, trunc=tt.TruncBond(10), tol=1e-14)
    b = np.array(O.beliefs(bp)[0])
    assert iters < 200
    # reference asserts isapprox with rtol = sqrt(eps); we are ~4e-11 away
    assert np.max(np.abs(b - SIS_INFINITE_GOLDEN)) < 1e-9


def test_glauber_infinite_vs_complete_graph():
    # /root/reference/test/glauber_infinite_graph.jl:7-45 (deterministic inputs, run without damping)
    T, k, m0 = 3, 3, 0.5
    w = [F.HomogeneousGlauberFactor(1.0, 0.0, 1.0) for _ in range(T + 1)]
    phi = [np.array([(1 + m0) / 2, (1 - m0) / 2]) if t == 0 else np.ones(2) for t in range(T + 1)]
    phi[1] = np.array([0.4, 0.6])
    phi[-1] = np.array([0.95, 0.05])
    bp = O.mpbp_infinite_graph(k, w, 2, phi)
    O.iterate(bp, maxiter=150, trunc=tt.TruncThresh(0.0), tol=1e-15)
    f_inf = O.bethe_free_energy(bp)
    N = k + 1
    g = O.BiDiGraph(N, [(a, b) for a in range(N) for b in range(a + 1, N)])
    bpx = O.MPBP(g, [w] * N, [2] * N, T, phi=[phi] * N)
    O.iterate(bpx, maxiter=150, trunc=tt.TruncThresh(0.0), tol=1e-15)
    f_k4 = O.bethe_free_energy(bpx) / N
    assert abs(f_inf - 0.98812749675847) < 1e-10  # derived known answer (SURVEY.md header fact 3(ii))
    assert abs(f_inf - f_k4) < 1e-10
    assert np.allclose(np.array(O.beliefs(bp)[0]), np.array(O.beliefs(bpx)[0]), atol=1e-10)


def _small_tree(rng, T=2):
    # /root/reference/test/glauber_small_tree.jl:6-24 (own RNG stream: Julia's cannot be replayed)
    und = [(0, 1), (1, 2), (1, 3)]
    N = 5
    g = O.BiDiGraph(N, und)
    h = rng.standard_normal(N)
    w = [[F.HomogeneousGlauberFactor(1.0, h[i], 1.0) for _ in range(T + 1)] for i in range(N)]
    phi = [[np.ones(2) for _ in range(T + 1)] for _ in range(N)]
    for i in range(N):
        phi[i][0] = np.array([0.75, 0.25])
    # N random (soft-ish) one-hot observations
    for _ in range(N):
        i, t = rng.integers(N), rng.integers(1, T + 1)
        o = np.full(2, 1e-3)
        o[rng.integers(2)] = 1.0
        phi[i][t] = phi[i][t] * o
    return g, w, phi


@pytest.mark.parametrize("generic", [False, True])
def test_glauber_small_tree_vs_exact(generic):
    rng = np.random.default_rng(111)
    T = 2
    g, w, phi = _small_tree(rng, T)
    if generic:
        w = [[F.GenericFactor(x) for x in wi] for wi in w]
    bp = O.MPBP(g, w, [2] * g.N, T, phi=phi)
    trunc = tt.TruncThresh(0.0) if generic else tt.TruncBondThresh(10)
    O.iterate(bp, maxiter=20, trunc=trunc, tol=0.0)
    p, Z, logZ = exact.exact_prob(bp)
    assert abs(np.exp(-O.bethe_free_energy(bp)) - Z) < 1e-9 * Z
    be = exact.exact_marginals(bp, p)
    assert np.allclose(np.array(O.beliefs(bp)), np.array(be), atol=1e-10)
    pb, _ = O.pair_beliefs(bp)
    pe = exact.exact_pair_marginals(bp, p)
    assert np.allclose(np.array(pb), np.array(pe), atol=1e-10)
    am, amex = O.alternate_marginals(bp), exact.exact_alternate_marginals(bp, p)  # test/glauber_small_tree.jl:58-61,66
    assert np.allclose(np.array(am), np.array(amex), atol=1e-10)
    f = lambda x, i: 2 * x - 3  # test/glauber_small_tree.jl:43-50: autocorrelations / autocovariances vs exact
    r, rex = O.autocorrelations(bp, f), exact.exact_autocorrelations(bp, p, f)
    assert np.allclose(np.array(r), np.array(rex), atol=1e-10)
    mu = O.means(bp, f)
    assert np.allclose(np.array(O.autocovariances(bp, f)), np.array([a - np.outer(m, m) for a, m in zip(rex, mu)]), atol=1e-10)
    for A in bp.mu:  # test/normalizations.jl:48-52
        assert abs(tt.lognormalization(A)) < 1e-10


def test_sis_small_tree_vs_exact():
    # /root/reference/test/sis_small_tree.jl:2-51 structure: star graph, T=3, bond cap 4 is exact
    T = 3
    g = O.BiDiGraph(4, [(0, 1), (0, 2), (0, 3)])
    lam, rho, gamma, alpha = 0.5, 0.4, 0.5, 0.1
    w = [[F.SISFactor(lam, rho, alpha) for _ in range(T + 1)] for _ in range(4)]
    phi = [[np.array([1 - gamma, gamma]) if t == 0 else np.ones(2) for t in range(T + 1)] for _ in range(4)]
    rng = np.random.default_rng(5)
    for i in range(4):
        o = np.full(2, 0.2)
        o[rng.integers(2)] = 1.0
        phi[i][T] = phi[i][T] * o
    bp = O.MPBP(g, w, [2] * 4, T, phi=phi)
    O.iterate(bp, maxiter=10, trunc=tt.TruncBondMax(4), tol=0.0)
    p, Z, _ = exact.exact_prob(bp)
    assert abs(np.exp(-O.bethe_free_energy(bp)) - Z) < 1e-9 * Z
    assert np.allclose(np.array(O.beliefs(bp)), np.array(exact.exact_marginals(bp, p)), atol=1e-10)
    pb, _ = O.pair_beliefs(bp)
    assert np.allclose(np.array(pb), np.array(exact.exact_pair_marginals(bp, p)), atol=1e-10)


def test_sis_heterogeneous_small_tree_vs_exact():
    # /root/reference/test/sis_heterogeneous.jl:1-47 structure: star, T=3, per-edge lambda, per-node rho/alpha, cap 8 exact
    T, N = 3, 4
    rng = np.random.default_rng(0)
    und = [(0, 1), (0, 2), (0, 3)]
    g = O.BiDiGraph(N, und)
    lam = np.zeros((N, N))
    for a, b in und:
        lam[a, b], lam[b, a] = rng.random(), rng.random()
    rho, alpha, gamma = rng.random(N), rng.random(N), 0.5
    w = []
    for i in range(N):
        nb = [g.dst[e] for e in range(len(g.src)) if g.src[e] == i]
        assert nb == sorted(nb)
        w.append([F.SIS_heterogeneousFactor([lam[j, i] for j in nb], rho[i], alpha[i])] * (T + 1))
    phi = [[np.array([1 - gamma, gamma]) if t == 0 else np.ones(2) for t in range(T + 1)] for _ in range(N)]
    for i in range(N):
        o = np.full(2, 0.1)
        o[rng.integers(2)] = 1.0
        phi[i][T] = phi[i][T] * o
    bp = O.MPBP(g, w, [2] * N, T, phi=phi)
    O.iterate(bp, maxiter=10, trunc=tt.TruncBondMax(8), tol=0.0)
    p, Z, _ = exact.exact_prob(bp)
    assert abs(np.exp(-O.bethe_free_energy(bp)) - Z) < 1e-9 * Z
    assert np.allclose(np.array(O.beliefs(bp)), np.array(exact.exact_marginals(bp, p)), atol=1e-10)
    pb, _ = O.pair_beliefs(bp)
    assert np.allclose(np.array(pb), np.array(exact.exact_pair_marginals(bp, p)), atol=1e-10)
    # the recursive interface reproduces the factor itself (what the reference checks with RestrictedRecursiveBPFactor, :49-60)
    f = w[0][0]
    for xn in (1, 2):
        for x in (1, 2):
            for xs in np.ndindex(2, 2, 2):
                xs = [v + 1 for v in xs]
                assert abs(F.RecursiveBPFactor.__call__(f, xn, xs, x) - f(xn, xs, x)) < 1e-14


def test_sis_heterogeneous_with_uniform_rates_equals_homogeneous():
    # /root/reference/test/sis_heterogeneous_compare_homogeneous.jl:1-35: loopy 5-node graph, TruncBond(3), tol=1e-12
    T, N = 3, 5
    A = np.array([[0, 1, 1, 0, 0], [1, 0, 1, 0, 0], [1, 1, 0, 1, 0], [0, 0, 1, 0, 1], [0, 0, 0, 1, 0]])
    und = [(i, j) for i in range(N) for j in range(i + 1, N) if A[i, j]]
    lam, rho, gamma = 0.15, 0.12, 0.13
    g = O.BiDiGraph(N, und)
    phi = [[np.array([1 - gamma, gamma]) if t == 0 else np.ones(2) for t in range(T + 1)] for _ in range(N)]
    wu = [[F.SISFactor(lam, rho)] * (T + 1) for _ in range(N)]
    wh = [[F.SIS_heterogeneousFactor([lam] * int(A[i].sum()), rho)] * (T + 1) for i in range(N)]
    res = []
    for w in (wu, wh):
        bp = O.MPBP(g, w, [2] * N, T, phi=[[p.copy() for p in ph] for ph in phi])
        O.iterate(bp, maxiter=200, trunc=tt.TruncBond(3), tol=1e-12)
        res.append(np.array(O.beliefs(bp)))
    assert np.allclose(res[0], res[1], atol=1e-12)


def test_pair_observations_tree_vs_exact():
    # /root/reference/test/pair_observations.jl:1-60: Glauber on a 5-node tree, T=2, psi on four (edge, time) pairs,
    # TruncThresh(0.0); recursive and generic (random couplings) factors
    T, N = 2, 5
    rng = np.random.default_rng(111)
    und = [(0, 1), (1, 2), (2, 3), (2, 4)]
    g = O.BiDiGraph(N, und)
    obs = [(0, 1, 1, np.array([[0.1, 0.9], [0.3, 0.4]])), (2, 3, 2, np.array([[0.4, 0.6], [0.5, 0.9]])),
           (2, 4, 2, rng.random((2, 2))), (1, 2, T, rng.random((2, 2)))]
    E2 = len(g.src)
    psi = [[np.ones((2, 2)) for _ in range(T + 1)] for _ in range(E2)]
    for (i, j, t, m) in obs:  # pair_observations_nondirected: psi_{i->j}[t][x_i, x_j] = m, psi_{j->i} = m^T
        for e in range(E2):
            if g.src[e] == i and g.dst[e] == j:
                psi[e][t] = psi[e][t] * m
            if g.src[e] == j and g.dst[e] == i:
                psi[e][t] = psi[e][t] * m.T
    h = rng.standard_normal(N)
    for generic in (False, True):
        if generic:
            J = {frozenset(e): rng.standard_normal() for e in und}
            w = []
            for i in range(N):
                nb = [g.dst[e] for e in range(E2) if g.src[e] == i]
                w.append([F.GenericGlauberFactor([J[frozenset((i, j))] for j in nb], h[i], 1.0)] * (T + 1))
        else:
            w = [[F.HomogeneousGlauberFactor(1.0, h[i], 1.0)] * (T + 1) for i in range(N)]
        phi = [[np.ones(2) for _ in range(T + 1)] for _ in range(N)]
        for _ in range(N):
            i, t = rng.integers(N), rng.integers(1, T + 1)
            o = np.full(2, 1e-2)
            o[rng.integers(2)] = 1.0
            phi[i][t] = phi[i][t] * o
        bp = O.MPBP(g, w, [2] * N, T, phi=phi, psi=[[m.copy() for m in ps] for ps in psi])
        O.iterate(bp, maxiter=10, trunc=tt.TruncThresh(0.0), tol=0.0)
        p, Z, _ = exact.exact_prob(bp)
        assert abs(np.exp(-O.bethe_free_energy(bp)) - Z) < 1e-9 * Z
        assert np.allclose(np.array(O.beliefs(bp)), np.array(exact.exact_marginals(bp, p)), atol=1e-10)
        pb, _ = O.pair_beliefs(bp)
        assert np.allclose(np.array(pb), np.array(exact.exact_pair_marginals(bp, p)), atol=1e-10)


def test_glauber_observe_everything_free_energy_is_logprob():
    # /root/reference/test/glauber_small_tree.jl:74-86: with every (i, t) observed (hard, zeros in phi) the Bethe free
    # energy is minus the log-probability of the observed trajectory
    rng = np.random.default_rng(111)
    T = 2
    g, w, phi = _small_tree(rng, T)
    phi = [[np.ones(2) for _ in range(T + 1)] for _ in range(g.N)]
    for i in range(g.N):
        phi[i][0] = np.array([0.75, 0.25])
    bp = O.MPBP(g, w, [2] * g.N, T, phi=phi)
    X = exact.onesample(bp, rng)
    for i in range(g.N):
        for t in range(T + 1):
            bp.phi[i][t] = bp.phi[i][t] * (np.arange(1, 3) == X[i, t])
    O.iterate(bp, maxiter=10, trunc=tt.TruncBondThresh(10), tol=0.0)
    assert abs(-O.bethe_free_energy(bp) - exact.logprob(bp, X)) < 1e-10
    for i in range(g.N):
        assert np.allclose(np.array(O.beliefs(bp)[i]), (np.arange(1, 3)[None, :] == X[i][:, None]), atol=1e-12)


def test_damped_factor_small_tree_vs_exact():
    # /root/reference/test/glauber_small_tree.jl:88-131: DampedFactor(w, 0.2) on the same tree, TruncBondThresh(10)
    rng = np.random.default_rng(111)
    T = 2
    g, w, phi = _small_tree(rng, T)
    w = [[F.DampedFactor(x, 0.2) for x in wi] for wi in w]
    bp = O.MPBP(g, w, [2] * g.N, T, phi=phi)
    O.iterate(bp, maxiter=20, trunc=tt.TruncBondThresh(10), tol=0.0)
    p, Z, _ = exact.exact_prob(bp)
    assert abs(np.exp(-O.bethe_free_energy(bp)) - Z) < 1e-9 * Z
    assert np.allclose(np.array(O.beliefs(bp)), np.array(exact.exact_marginals(bp, p)), atol=1e-10)
    pb, _ = O.pair_beliefs(bp)
    assert np.allclose(np.array(pb), np.array(exact.exact_pair_marginals(bp, p)), atol=1e-10)


def test_sirs_small_tree_vs_exact():
    # /root/reference/test/sirs_small_tree.jl structure (q=3, TruncThresh(0.0))
    T = 2
    g = O.BiDiGraph(3, [(0, 1), (1, 2)])
    w = [[F.SIRSFactor(0.4, 0.15, 0.2, 0.05) for _ in range(T + 1)] for _ in range(3)]
    gamma = 0.3
    phi = [[np.array([1 - gamma, gamma, 0.0]) if t == 0 else np.ones(3) for t in range(T + 1)] for _ in range(3)]
    phi[0][2] = np.array([0.1, 1.0, 0.3])
    bp = O.MPBP(g, w, [3] * 3, T, phi=phi)
    O.iterate(bp, maxiter=6, trunc=tt.TruncThresh(0.0), tol=0.0)
    p, Z, _ = exact.exact_prob(bp)
    assert abs(np.exp(-O.bethe_free_energy(bp)) - Z) < 1e-9 * Z
    assert np.allclose(np.array(O.beliefs(bp)), np.array(exact.exact_marginals(bp, p)), atol=1e-10)


def test_schedules_share_fixed_point():
    # test/sis_heterogeneous_compare_homogeneous.jl:5-35 graph: loopy, TruncBond(3)
    T = 3
    und = [(0, 1), (0, 2), (1, 2), (2, 3), (3, 4)]
    g = O.BiDiGraph(5, und)
    w = [[F.SISFactor(0.15, 0.12) for _ in range(T + 1)] for _ in range(5)]
    phi = [[np.array([0.87, 0.13]) if t == 0 else np.ones(2) for t in range(T + 1)] for _ in range(5)]
    res = []
    for sched in ("sequential", "parallel"):
        bp = O.MPBP(g, w, [2] * 5, T, phi=phi)
        O.iterate(bp, maxiter=200, trunc=tt.TruncBond(3), tol=1e-13, schedule=sched)
        res.append(np.array(O.beliefs(bp)))
    assert np.allclose(res[0], res[1], atol=1e-7)


def test_mpem_evaluate_is_invariant_under_orthogonalisation_and_mpem2():
    # /root/reference/test/mpems.jl:3-40: evaluate(A, x) unchanged by orthogonalize_left! (MPEM1, MPEM2), by compress!
    # with a non-binding truncation, and by the MPEM3 -> MPEM2 conversion
    rng = np.random.default_rng(4)
    for phys, T, d in (((2,), 10, 4), ((2, 2), 5, 4), ((3, 2), 4, 3)):
        A = tt.rand_tt([1] + [d] * T + [1], *phys, rng=rng)
        xs = [[tuple(rng.integers(0, p) for p in phys) for _ in range(T + 1)] for _ in range(5)]
        e1 = [A.evaluate(x) for x in xs]
        B = A.copy()
        tt.orthogonalize_left(B, tt.TruncThresh(0.0))
        assert np.allclose([B.evaluate(x) for x in xs], e1, rtol=1e-10)
        B = A.copy()
        tt.orthogonalize_right(B, tt.TruncThresh(0.0))
        assert np.allclose([B.evaluate(x) for x in xs], e1, rtol=1e-10)
        B = A.copy()
        tt.compress(B, tt.TruncBond(d * d))
        assert np.allclose([B.evaluate(x) for x in xs], e1, rtol=1e-10)
        assert abs(tt.lognormalization(A)) < 1e-10  # rand_tt normalises
    Bs = [rng.random((1, 3, 2, 2, 2)), rng.random((3, 4, 2, 2, 2)), rng.random((4, 1, 2, 2, 2))]
    Bs[-1][:, :, :, :, 1] = Bs[-1][:, :, :, :, 0]
    C = O.mpem2(Bs, 0.0)
    for _ in range(8):
        x = [tuple(rng.integers(0, 2, size=2)) for _ in range(3)]
        M = np.ones((1, 1))
        for t in range(3):
            xn = x[t + 1][0] if t < 2 else 0
            M = M @ Bs[t][:, :, x[t][0], x[t][1], xn]
        assert abs(C.evaluate(x) - M[0, 0]) < 1e-12 * max(1.0, abs(M[0, 0]))


@pytest.mark.parametrize("kind", ["pmj", "integer"])
def test_pmj_and_integer_glauber_small_tree_vs_exact(kind):
    # /root/reference/test/glauber_pmJ_small_tree.jl:1-63 (J = +-1 on a 4-node star-like tree, beta = 2, T = 3,
    # TruncThresh(0.0)) and the IntegerGlauber testset of test/glauber_small_tree.jl:174-230 (integer couplings)
    rng = np.random.default_rng(111)
    T, N = 3, 4
    und = [(0, 1), (1, 2), (1, 3)]
    Jm = {frozenset((0, 1)): -1, frozenset((1, 2)): 1, frozenset((1, 3)): 1} if kind == "pmj" else \
         {frozenset((0, 1)): -2, frozenset((1, 2)): 1, frozenset((1, 3)): 3}
    g = O.BiDiGraph(N, und)
    h = rng.standard_normal(N)
    beta = 2.0 if kind == "pmj" else 0.7
    w = []
    for i in range(N):
        nb = [g.dst[e] for e in g.out_edges[i]]
        Ji = [Jm[frozenset((i, j))] for j in nb]
        if kind == "pmj":
            w.append([F.PMJGlauberFactor(Ji, 1.0, float(h[i]), beta)] * (T + 1))
        else:
            w.append([F.IntegerGlauberFactor(Ji, float(h[i]), beta)] * (T + 1))
    phi = [[np.array([0.75, 0.25]) if t == 0 else np.ones(2) for t in range(T + 1)] for _ in range(N)]
    for _ in range(N):
        i, t = rng.integers(N), rng.integers(1, T + 1)
        o = np.full(2, 1e-2)
        o[rng.integers(2)] = 1.0
        phi[i][t] = phi[i][t] * o
    bp = O.MPBP(g, w, [2] * N, T, phi=phi)
    O.iterate(bp, maxiter=20, trunc=tt.TruncThresh(0.0), tol=0.0)
    p, Z, _ = exact.exact_prob(bp)
    assert abs(np.exp(-O.bethe_free_energy(bp)) - Z) < 1e-9 * Z
    assert np.allclose(np.array(O.beliefs(bp)), np.array(exact.exact_marginals(bp, p)), atol=1e-10)
    f = lambda x, i: 2 * x - 3
    assert np.allclose(np.array(O.autocorrelations(bp, f)), np.array(exact.exact_autocorrelations(bp, p, f)), atol=1e-10)
    # the recursive interface reproduces the factor's own definition
    for i in range(N):
        z = len(g.out_edges[i])
        for xs in np.ndindex(*([2] * z)):
            for xn in (1, 2):
                for x in (1, 2):
                    xs1 = [v + 1 for v in xs]
                    assert abs(F.RecursiveBPFactor.__call__(w[i][0], xn, xs1, x) - w[i][0](xn, xs1, x)) < 1e-14


def test_periodic_glauber_tree_with_pair_observations_vs_exact():
    # /root/reference/test/periodic.jl:1-75: periodic-in-time Glauber on a small tree with pair observations,
    # TruncBondThresh(10) (non-binding); Z and marginals against brute force of the wrapped dynamics
    from oracle import periodic as P
    rng = np.random.default_rng(111)
    T, N = 2, 5
    und = [(0, 1), (1, 2), (1, 3)]
    g = O.BiDiGraph(N, und)
    h = rng.standard_normal(N)
    w = [[F.HomogeneousGlauberFactor(1.0, float(h[i]), 1.0)] * (T + 1) for i in range(N)]
    obs = [(0, 1, 1, np.array([[0.1, 0.9], [0.3, 0.4]])), (1, 3, 2, np.array([[0.4, 0.6], [0.5, 0.9]])), (1, 2, T, rng.random((2, 2)))]
    psi = [[np.ones((2, 2)) for _ in range(T + 1)] for _ in range(g.ne)]
    for (i, j, t, m) in obs:
        for e in range(g.ne):
            if g.src[e] == i and g.dst[e] == j:
                psi[e][t] = psi[e][t] * m
            if g.src[e] == j and g.dst[e] == i:
                psi[e][t] = psi[e][t] * m.T
    phi = [[np.ones(2) for _ in range(T + 1)] for _ in range(N)]
    for i in range(N):
        phi[i][0] = np.array([0.75, 0.25])
    phi[2][1] = np.array([1.0, 0.1])
    phi[0][2] = np.array([0.3, 1.0])
    bp = P.PeriodicMPBP(g, w, [2] * N, T, phi=phi, psi=psi)
    P.iterate(bp, maxiter=8, trunc=tt.TruncBondThresh(10))
    p, logZ = P.exact_prob(bp)
    assert abs(-P.bethe_free_energy(bp) - logZ) < 1e-10
    L = T + 1
    be = [[p.sum(axis=tuple(a for a in range(N * L) if a != i * L + t)) for t in range(L)] for i in range(N)]
    assert np.allclose(np.array(P.beliefs(bp)), np.array(be), atol=1e-10)
    for A in bp.mu:
        assert abs(P.lognormalization(A)) < 1e-10
        assert max(max(a.shape[:2]) for a in A) <= 10  # every bond, the closing one included, obeys the cap


def test_periodic_sis_tree_vs_exact_and_ring_invariances():
    from oracle import periodic as P
    rng = np.random.default_rng(3)
    T, N = 3, 3
    g = O.BiDiGraph(N, [(0, 1), (1, 2)])
    w = [[F.SISFactor(0.3, 0.25, 0.05)] * (T + 1) for _ in range(N)]
    phi = [[np.ones(2) for _ in range(T + 1)] for _ in range(N)]
    phi[0][1] = np.array([0.2, 1.0])
    phi[2][3] = np.array([1.0, 0.4])
    bp = P.PeriodicMPBP(g, w, [2] * N, T, phi=phi)
    P.iterate(bp, maxiter=6, trunc=tt.TruncThresh(0.0))
    p, logZ = P.exact_prob(bp)
    assert abs(-P.bethe_free_energy(bp) - logZ) < 1e-10
    L = T + 1
    be = [[p.sum(axis=tuple(a for a in range(N * L) if a != i * L + t)) for t in range(L)] for i in range(N)]
    assert np.allclose(np.array(P.beliefs(bp)), np.array(be), atol=1e-10)
    # /root/reference/test/mpems.jl:42-65: evaluate is invariant under ring orthogonalisation and periodic mpem2
    A = tt.TT([rng.random((4, 4, 2, 2)) for _ in range(6)])
    xs = [[tuple(rng.integers(0, 2, size=2)) for _ in range(6)] for _ in range(6)]
    e1 = [P.evaluate(A, x) for x in xs]
    for fn in (lambda B: P.orthogonalize_left(B, tt.TruncThresh(0.0)), lambda B: P.orthogonalize_right(B, tt.TruncThresh(0.0)),
               lambda B: P.compress(B, tt.TruncBond(64))):
        B = A.copy()
        fn(B)
        assert np.allclose([P.evaluate(B, x) for x in xs], e1, rtol=1e-9)
    Bs = [rng.random((2, 3, 2, 2, 2)), rng.random((3, 4, 2, 2, 2)), rng.random((4, 2, 2, 2, 2))]
    C = P.mpem2(Bs, 0.0)
    for _ in range(8):
        x = [tuple(rng.integers(0, 2, size=2)) for _ in range(3)]
        M = np.eye(2)
        for t in range(3):
            M = M @ Bs[t][:, :, x[t][0], x[t][1], x[(t + 1) % 3][0]]
        assert abs(P.evaluate(C, x) - np.trace(M)) < 1e-12 * max(1.0, abs(np.trace(M)))


@pytest.mark.parametrize("seed", range(6))
def test_random_small_trees_are_exact(seed):
    """property the reference relies on throughout its test-suite: on a tree, with a non-binding truncation, MPBP is
    exact (Z, one- and two-node marginals) for every model, any reweightings and any pair observations"""
    rng = np.random.default_rng(1000 + seed)
    N = int(rng.integers(3, 6))
    T = int(rng.integers(1, 3))
    und = [(int(rng.integers(0, k)), k) for k in range(1, N)]  # random recursive tree
    g = O.BiDiGraph(N, und)
    model = ["glauber", "sis", "sirs"][seed % 3]
    q = 3 if model == "sirs" else 2
    w = []
    for i in range(N):
        if model == "glauber":
            f = F.HomogeneousGlauberFactor(float(rng.normal()), float(rng.normal()), 1.0)
        elif model == "sis":
            f = F.SISFactor(float(rng.random()), float(rng.random()), 0.2 * float(rng.random()))
        else:
            f = F.SIRSFactor(float(rng.random()), float(rng.random()), float(rng.random()), 0.2 * float(rng.random()))
        w.append([f] * (T + 1))
    phi = [[0.1 + rng.random(q) for _ in range(T + 1)] for _ in range(N)]
    psi = [None] * g.ne
    for e in range(g.ne):
        if psi[e] is None:
            ps = [0.2 + rng.random((q, q)) for _ in range(T + 1)]
            psi[e] = ps
            psi[g.rev[e]] = [p.T.copy() for p in ps]
    bp = O.MPBP(g, w, [q] * N, T, phi=phi, psi=psi)
    O.iterate(bp, maxiter=N + 2, trunc=tt.TruncThresh(0.0), tol=0.0)
    p, Z, logZ = exact.exact_prob(bp)
    assert abs(-O.bethe_free_energy(bp) - logZ) < 1e-9
    assert np.allclose(np.array(O.beliefs(bp)), np.array(exact.exact_marginals(bp, p)), atol=1e-10)
    pb, _ = O.pair_beliefs(bp)
    assert np.allclose(np.array(pb), np.array(exact.exact_pair_marginals(bp, p)), atol=1e-10)
    assert np.allclose(np.array(O.alternate_marginals(bp)), np.array(exact.exact_alternate_marginals(bp, p)), atol=1e-10)
    fobs = lambda x, i: x * x - 1.5
    assert np.allclose(np.array(O.autocorrelations(bp, fobs)), np.array(exact.exact_autocorrelations(bp, p, fobs)), atol=1e-10)


def test_glauber_infinite_bipartite_graph_known_answer():
    """/root/reference/test/glauber_infinite_graph.jl:48-100: BP on the infinite bipartite (3,2)-regular graph equals BP on the
    complete bipartite graph K_{2,3} (free energy per node and beliefs of the two classes); TruncThresh(0.0), damp = 0.1 on
    the infinite side, as in the reference test."""
    T, k, m0 = 3, (3, 2), 0.5
    JA, JB, h, beta = 1.0, -0.2, -0.1, 1.0
    w = [[F.HomogeneousGlauberFactor(JA, h, beta)] * (T + 1), [F.HomogeneousGlauberFactor(JB, h, beta)] * (T + 1)]
    phi = [[np.array([(1 + m0) / 2, (1 - m0) / 2]) if t == 0 else np.ones(2) for t in range(T + 1)] for _ in range(2)]
    phi[0][1] = np.array([0.4, 0.6])
    phi[1][T] = np.array([0.95, 0.05])
    bp = O.mpbp_infinite_bipartite_graph(k, w, (2, 2), phi=phi)
    it, _ = O.iterate(bp, maxiter=150, trunc=tt.TruncThresh(0.0), tol=1e-15, damp=0.1)
    assert it < 150
    N = sum(k)
    und = [(a, b) for a in range(k[1]) for b in range(k[1], N)]
    g = O.BiDiGraph(N, und)
    wex = [[F.HomogeneousGlauberFactor(JA if i < k[1] else JB, h, beta)] * (T + 1) for i in range(N)]
    phiex = [[p.copy() for p in (phi[0] if i < k[1] else phi[1])] for i in range(N)]
    be = O.MPBP(g, wex, [2] * N, T, phi=phiex)
    it2, _ = O.iterate(be, maxiter=150, trunc=tt.TruncThresh(0.0), tol=1e-15)
    assert it2 < 150
    assert abs(np.exp(-O.bethe_free_energy(bp)) - np.exp(-O.bethe_free_energy(be) / N)) < 1e-10
    b, bex = O.beliefs(bp), O.beliefs(be)
    assert np.max(np.abs(np.array(b[0]) - np.array(bex[0]))) < 1e-10
    assert np.max(np.abs(np.array(b[1]) - np.array(bex[k[1]]))) < 1e-10
