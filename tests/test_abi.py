"""CPU-side checks of the boundary: the shared library loads and exports every symbol include/mpbp.h declares."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mpbp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mpbp_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    import __graft_entry__ as G
    G.build()
    from mpbp_b200 import _lib
    lib = C.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mpbp.h but not exported"
    for n in names:
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in _lib.py"


def test_version_and_error_string_without_gpu():
    from mpbp_b200 import _lib
    L = _lib.lib()
    assert L.mpbp_version() == 100
    assert isinstance(L.mpbp_last_error(), bytes)
    # argument validation happens before any CUDA call
    assert L.mpbp_destroy(None) == 0


def test_missing_library_is_loud(monkeypatch, tmp_path):
    from mpbp_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.MPBPError):
        _lib.lib()


def test_host_tabulation_matches_oracle_factors():
    """the product's host-side factor tables agree with the oracle's independent factor definitions"""
    import numpy as np
    import mpbp_b200 as M
    from oracle import factors as OF
    pairs = [
        (OF.HomogeneousGlauberFactor(0.7, -0.2, 1.3), M.HomogeneousGlauberFactor(0.7, -0.2, 1.3), 2),
        (OF.SISFactor(0.3, 0.2, 0.05), M.SISFactor(0.3, 0.2, 0.05), 2),
        (OF.SIS_heterogeneousFactor([0.3, 0.6, 0.1], 0.2, 0.05), M.SIS_heterogeneousFactor([0.3, 0.6, 0.1], 0.2, 0.05), 2),
        (OF.SIRSFactor(0.3, 0.2, 0.1, 0.05), M.SIRSFactor(0.3, 0.2, 0.1, 0.05), 3),
        (OF.PMJGlauberFactor([1, -1, 1], 0.5, 0.1, 1.0), M.PMJGlauberFactor([1, -1, 1], 0.5, 0.1, 1.0), 2),
        (OF.IntegerGlauberFactor([1, -2, 1], 0.1, 0.8), M.IntegerGlauberFactor([1, -2, 1], 0.1, 0.8), 2),
    ]
    for a, b, q in pairs:
        z = 3
        for l in range(z + 1):
            assert a.nstates(l) == b.nstates(l)
        for xn in range(1, q + 1):
            for x in range(1, q + 1):
                for y in range(1, a.nstates(z) + 1):
                    assert abs(a.prob_y(xn, x, y, z) - b.prob_y(xn, x, y, z)) < 1e-15
                for xk in range(1, q + 1):
                    for y in range(1, a.nstates(z - 1) + 1):
                        assert abs(a.prob_y_partial(xn, x, xk, y, z - 1, 2) - b.prob_y_partial(xn, x, xk, y, z - 1, 2)) < 1e-15
        import itertools
        for xs in itertools.product(range(1, q + 1), repeat=z):
            for xn in range(1, q + 1):
                for x in range(1, q + 1):
                    assert abs(a(xn, list(xs), x) - b(xn, list(xs), x)) < 1e-14


def test_cavity_pairs_cover_recursion():
    from mpbp_b200.factors import cavity_pairs
    for z in range(1, 8):
        ps = set(cavity_pairs(z))
        if z == 1:
            assert ps == {(1, 0)}
            continue
        need = {(k, 1) for k in range(1, z)} | {(z, 0)} | {(1, l) for l in range(0, z - 1)} | {(k, z - 1 - k) for k in range(1, z)}
        assert need <= ps


def test_graph_edge_order_matches_reference_convention():
    import mpbp_b200 as M
    g = M.IndexedBiDiGraph(4, [(2, 0), (0, 1), (3, 0)])
    assert list(zip(g.src, g.dst)) == [(0, 1), (0, 2), (0, 3), (1, 0), (2, 0), (3, 0)]
    assert list(g.rev) == [3, 4, 5, 0, 1, 2]
    assert list(g.colptr) == [0, 3, 4, 5, 6]


def test_host_sampler_and_observation_drawing():
    """src/sampling.jl:30-66,191-210 restated on the host: prior forward sampling and (soft) one-hot reweightings"""
    import numpy as np
    import mpbp_b200 as M
    from mpbp_b200.sampling import draw_node_observations_, onesample
    N, T = 12, 6
    g = M.IndexedBiDiGraph(N, [(i, (i + 1) % N) for i in range(N)] + [(0, 6)])
    w = [[M.SIRSFactor(0.6, 0.3, 0.2, 0.05)] * (T + 1) for _ in range(N)]
    q = [3] * N
    phi = [[np.array([0.5, 0.5, 0.0]) if t == 0 else np.ones(3) for t in range(T + 1)] for _ in range(N)]
    X, wgt = onesample(g, w, q, T, phi, rng=7)
    X2, _ = onesample(g, w, q, T, phi, rng=7)
    assert X.shape == (N, T + 1) and np.array_equal(X, X2) and wgt == 1.0
    assert set(np.unique(X)) <= {1, 2, 3} and not np.any(X[:, 0] == 3)
    # SIRS: S -> R and R -> I never happen in one step
    for t in range(T):
        assert not np.any((X[:, t] == 1) & (X[:, t + 1] == 3)) and not np.any((X[:, t] == 3) & (X[:, t + 1] == 2))
    _, obs = draw_node_observations_(phi, X, 10, rng=1)
    assert len(obs) == 10 and obs == sorted(obs) and len(set(obs)) == 10
    for (i, t) in obs:
        assert phi[i][t][X[i, t] - 1] > 0 and np.count_nonzero(phi[i][t]) == 1
    phi2 = [[np.ones(3) for _ in range(T + 1)] for _ in range(N)]
    _, obs2 = draw_node_observations_(phi2, X, N, softinf=100.0, last_time=True, rng=2)
    assert all(t == T for _, t in obs2) and len(obs2) == N
    i, t = obs2[0]
    assert abs(phi2[i][t][X[i, t] - 1] - 100.0 / 101.0) < 1e-12 and abs(phi2[i][t].min() - 1.0 / 101.0) < 1e-12
    # the likelihood weight accounts for reweightings at t > 0
    Xw, wgt2 = onesample(g, w, q, T, phi2, rng=7)
    assert 0.0 < wgt2 < 1.0


def test_glauber_factor_selection_matches_the_transition_probability():
    """glauber_factors (src/Models/glauber/glauber_bp.jl:121-142): homogeneous / +-J / integer / generic couplings pick
    different factor types, all of which must equal 1 / (1 + exp(-2 beta s' (sum_j J_ij s_j + h_i)))"""
    import itertools
    import math
    import numpy as np
    import mpbp_b200 as M
    g = M.IndexedBiDiGraph(5, [(0, 1), (1, 2), (1, 3), (3, 4), (0, 3)])
    rng = np.random.default_rng(0)
    nund = len(g.undirected)
    cases = {
        M.HomogeneousGlauberFactor: np.full(nund, 0.7),
        M.PMJGlauberFactor: 0.7 * rng.choice([-1.0, 1.0], size=nund) * np.array([1, -1] + [1] * (nund - 2)),
        M.IntegerGlauberFactor: np.array([1.0, -2.0, 3.0, 1.0, 2.0][:nund]),
        M.GenericGlauberFactor: rng.standard_normal(nund),
    }
    h, beta, T = rng.standard_normal(5), 1.3, 2
    spin = lambda x: 3 - 2 * x  # potts2spin: 1 -> +1, 2 -> -1
    for cls, J in cases.items():
        w = M.glauber_factors(M.Ising(g, J=J, h=h, beta=beta), T)
        assert len(w) == 5 and all(len(wi) == T + 1 for wi in w)
        assert all(type(wi[0]) is cls for wi in w), (cls, [type(wi[0]).__name__ for wi in w])
        for i in range(5):
            nb = g.neighbors(i)
            Ji = [J[g.und_of[e]] for e in g.outedges(i)]
            for xs in itertools.product((1, 2), repeat=len(nb)):
                field = sum(j * spin(x) for j, x in zip(Ji, xs)) + h[i]
                for xn in (1, 2):
                    p = 1.0 / (1.0 + math.exp(-2.0 * beta * spin(xn) * field))
                    assert abs(w[i][0](xn, list(xs), 1) - p) < 1e-12, (cls.__name__, i, xs, xn)


def test_truncation_policies_and_graph_builders():
    """(kind, d, eps) passed through the C ABI must select the same number of singular values as the oracle's
    TensorTrains restatement; the graph builders agree on the reference's CSC edge order"""
    import numpy as np
    import networkx as nx
    import mpbp_b200 as M
    from oracle import tt as OT
    from tests.common import otrunc
    lam = np.array([3.0, 1.0, 0.5, 1e-3, 1e-9, 0.0])

    def keep_abi(tr):  # the rule the device applies (csrc/common.cuh trunc_keep), restated
        n = len(lam)
        if tr.kind == 0:
            return min(n, tr.d)
        k = max(1, max([i + 1 for i in range(n) if lam[i] > tr.eps * np.linalg.norm(lam)], default=1))
        return min(k, tr.d) if tr.kind == 2 else k

    for tr in (M.TruncBond(2), M.TruncBond(10), M.TruncBondMax(3), M.TruncThresh(0.0), M.TruncThresh(1e-2), M.TruncThresh(1e-6),
               M.TruncBondThresh(2, 1e-6), M.TruncBondThresh(5, 1e-2), M.TruncBondThresh(4)):
        assert (tr.kind, tr.d >= 0, tr.eps >= 0.0) == (tr.kind, True, True)
        assert keep_abi(tr) == otrunc(tr).keep(lam), tr
    G = nx.petersen_graph()
    g1 = M.IndexedBiDiGraph.from_networkx(G)
    g2 = M.IndexedBiDiGraph.from_adjacency(nx.to_numpy_array(G))
    g3 = M.IndexedBiDiGraph(10, list(G.edges()))
    for g in (g2, g3):
        assert np.array_equal(g.src, g1.src) and np.array_equal(g.dst, g1.dst) and np.array_equal(g.rev, g1.rev)
    assert np.all(np.diff(g1.src) >= 0) and all(np.all(np.diff(g1.neighbors(i)) > 0) for i in range(10))  # sorted by (src, dst)
    assert np.array_equal(g1.src[g1.rev], g1.dst) and np.array_equal(g1.dst[g1.rev], g1.src)
    assert g1.ne == 30 and all(g1.degree(i) == 3 for i in range(10))
