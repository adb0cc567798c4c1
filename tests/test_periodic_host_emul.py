"""CPU tier: the periodic node-update KERNEL SOURCE (matrixproductbp.jl_b200/csrc/periodic.cuh), compiled unchanged by g++
as a one-thread emulation (tests/host_emul/periodic_host.cpp), against oracle/periodic.py and brute force.  This checks the
arithmetic and the indexing of the device code without a GPU; thread-level behaviour is covered by tests/test_gpu_periodic.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from mpbp_b200 import factors as PF
from oracle import factors as OF, mpbp as O, periodic as P, tt

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_emul", "periodic_host.cpp")
OUT = os.path.join(HERE, "host_emul", "_build", "libper_host.so")
DEPS = [SRC] + [os.path.join(HERE, "..", "matrixproductbp.jl_b200", "csrc", f) for f in ("periodic.cuh", "periodic_plan.h")]

dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)


def _p(a, t):
    return a.ctypes.data_as(t)


@pytest.fixture(scope="module")
def emu():
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in DEPS):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", OUT, SRC], check=True)
    lib = C.CDLL(OUT)
    lib.per_host_svd.restype = C.c_int
    lib.per_host_node_update.restype = C.c_int
    lib.per_host_pair_belief.restype = C.c_int
    return lib


@pytest.mark.parametrize("R,Cc", [(12, 5), (5, 12), (7, 7), (1, 4), (4, 1), (40, 9)])
@pytest.mark.parametrize("wl", [0, 1])
def test_svd_factor(emu, R, Cc, wl):
    rng = np.random.default_rng(R * 100 + Cc)
    M = np.asfortranarray(rng.standard_normal((R, Cc)) * 3.7)
    if R == 40:
        M[:, 5:] = M[:, :4] @ rng.standard_normal((4, 4))  # rank 5: the zero singular values must be dropped
    Lf, Rf = np.zeros(R * Cc), np.zeros(R * Cc)
    ls, err = C.c_double(), C.c_int()
    r = emu.per_host_svd(_p(M, dp), R, Cc, wl, 1, 0, C.c_double(0.0), _p(Lf, dp), _p(Rf, dp), C.byref(ls), C.byref(err))
    assert err.value == 0
    assert r == (5 if R == 40 else min(R, Cc))
    Lm = Lf[:R * r].reshape((R, r), order="F")
    Rm = Rf[:r * Cc].reshape((r, Cc), order="F")
    assert np.allclose(np.exp(ls.value) * Lm @ Rm, M, atol=1e-13 * np.abs(M).max())
    s = np.linalg.svd(M / np.abs(M).max(), compute_uv=False)[:r]
    if wl == 0:
        assert np.allclose(Rm @ Rm.T, np.eye(r), atol=1e-12)
        assert np.allclose(np.linalg.norm(Lm, axis=0), s, atol=1e-12)
    else:
        assert np.allclose(Lm.T @ Lm, np.eye(r), atol=1e-12)
        assert np.allclose(np.linalg.norm(Rm, axis=1), s, atol=1e-12)
    # truncation: TruncBond(2)
    r2 = emu.per_host_svd(_p(M, dp), R, Cc, wl, 0, 2, C.c_double(0.0), _p(Lf, dp), _p(Rf, dp), C.byref(ls), C.byref(err))
    assert r2 == min(2, min(R, Cc))


class HostPeriodicBP:
    """drives per_host_node_update over a graph: messages live in numpy in the device slot format"""

    def __init__(self, emu, g, w, q, T, phi, psi, dmax):
        self.emu, self.g, self.w, self.q, self.T, self.dmax = emu, g, w, q, T, dmax
        self.L = T + 1
        self.phi, self.psi = phi, psi
        qmax = max(q)
        self.ss = dmax * dmax * qmax * qmax
        ne = g.ne
        self.bonds = np.ones((ne, self.L + 1), dtype=np.int32)
        self.data = np.zeros((ne, self.L * self.ss))
        self.ls = np.zeros(ne)
        for e in range(ne):
            P_ = q[g.src[e]] * q[g.dst[e]]
            for t in range(self.L):
                self.data[e, t * self.ss:t * self.ss + P_] = 1.0
            self.ls[e] = -self.L * np.log(P_)
        self.marg = [np.full((self.L, q[i]), 1.0 / q[i]) for i in range(g.N)]
        self.tv = [np.zeros((self.L, self.L, qmax * qmax)) for i in range(g.N)]
        self.f = np.zeros(g.N)

    def update(self, i, trunc, damp=0.0):
        g, q, L = self.g, self.q, self.L
        eout = g.out_edges[i]
        ein = g.in_edges[i]
        z = len(eout)
        qi = q[i]
        qn = np.array([q[g.dst[e]] for e in eout], dtype=np.int32)
        wi = self.w[i]
        same = all(x is wi[0] for x in wi)
        ws = [wi[0]] if same else list(wi)
        tab = PF.tabulate_class(ws, z, qi, qn)
        d1 = np.array([p[0] for p in tab["pairs"]], dtype=np.int32)
        d2 = np.array([p[1] for p in tab["pairs"]], dtype=np.int32)
        phi = np.ascontiguousarray(np.concatenate([np.asarray(p, dtype=np.float64) for p in self.phi[i]]))
        psi = np.ascontiguousarray(np.concatenate([np.asarray(p, dtype=np.float64).ravel(order="F") for e in eout for p in self.psi[e]])) if z else np.zeros(1)
        ib = np.ascontiguousarray(self.bonds[ein]) if z else np.zeros((1, L + 1), dtype=np.int32)
        idat = np.ascontiguousarray(self.data[ein]) if z else np.zeros((1, 1))
        ils = np.ascontiguousarray(self.ls[ein]) if z else np.zeros(1)
        # the out slots start as the old messages (set_msg! damps against them)
        ob = np.ascontiguousarray(self.bonds[eout]) if z else np.zeros((1, L + 1), dtype=np.int32)
        odat = np.ascontiguousarray(self.data[eout]) if z else np.zeros((1, 1))
        ols = np.ascontiguousarray(self.ls[eout]) if z else np.zeros(1)
        marg = np.zeros(L * qi)
        lzi, f = C.c_double(), C.c_double()
        lzij = np.zeros(max(z, 1))
        err = C.c_int()
        ny = np.ascontiguousarray(tab["ny"], dtype=np.int32)
        tv = np.zeros(L * L * self.tv[i].shape[2])
        rc = self.emu.per_host_node_update(z, qi, _p(qn, ip), self.T, self.dmax, _p(ny, ip), len(ws), _p(tab["pxy"], dp), len(d1),
                                           _p(d1, ip), _p(d2, ip), _p(tab["pyy"], dp), _p(tab["w"], dp), _p(tab["wd"], dp),
                                           _p(tab["minit"], dp), _p(phi, dp), _p(psi, dp), trunc.kind, trunc.d, C.c_double(trunc.eps),
                                           C.c_double(damp), self.ss, _p(ib, ip), _p(idat, dp), _p(ils, dp), _p(ob, ip), _p(odat, dp), _p(ols, dp),
                                           _p(marg, dp), C.byref(lzi), _p(lzij, dp), C.byref(f), _p(tv, dp), L, self.tv[i].shape[2],
                                           C.byref(err))
        assert rc == 0 and err.value == 0, (rc, err.value)
        for k, e in enumerate(eout):
            self.bonds[e], self.data[e], self.ls[e] = ob[k], odat[k], ols[k]
        self.marg[i] = marg.reshape(L, qi)
        self.tv[i] = tv.reshape(L, L, -1)
        self.f[i] = f.value

    def pair_beliefs(self):
        g, q, L = self.g, self.q, self.L
        b, logz = [None] * g.ne, np.zeros(g.N)
        for e in range(g.ne):
            qi, qj = q[g.src[e]], q[g.dst[e]]
            r = g.rev[e]
            psi = np.ascontiguousarray(np.concatenate([np.asarray(p, dtype=np.float64).ravel(order="F") for p in self.psi[e]]))
            out = np.zeros(L * qi * qj)
            lz, err = C.c_double(), C.c_int()
            ba, bb = np.ascontiguousarray(self.bonds[e]), np.ascontiguousarray(self.bonds[r])
            da, db = np.ascontiguousarray(self.data[e]), np.ascontiguousarray(self.data[r])
            self.emu.per_host_pair_belief(qi, qj, self.T, self.dmax, self.ss, _p(ba, ip), _p(da, dp), C.c_double(self.ls[e]), _p(bb, ip),
                                          _p(db, dp), C.c_double(self.ls[r]), _p(psi, dp), _p(out, dp), C.byref(lz), C.byref(err))
            assert err.value == 0
            b[e] = [out[t * qi * qj:(t + 1) * qi * qj].reshape((qi, qj), order="F") for t in range(L)]
            j = g.src[e]
            logz[j] += (1.0 / g.degree(j) - 0.5) * lz.value
        return b, logz

    def message(self, e):
        """site tensors with the normalisation folded in, like mpbp_get_message"""
        P_ = (self.q[self.g.src[e]], self.q[self.g.dst[e]])
        b = self.bonds[e]
        sc = np.exp(self.ls[e] / self.L)
        return [sc * self.data[e, t * self.ss:t * self.ss + b[t] * b[t + 1] * P_[0] * P_[1]].reshape((b[t], b[t + 1]) + P_, order="F") for t in range(self.L)]


class Tr:
    def __init__(self, kind, d, eps):
        self.kind, self.d, self.eps = kind, d, eps


def _glauber_tree():
    rng = np.random.default_rng(111)
    T, N = 2, 5
    und = [(0, 1), (1, 2), (1, 3)]
    g = O.BiDiGraph(N, und)
    h = rng.standard_normal(N)
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
    return g, h, T, N, phi, psi


def test_periodic_glauber_tree_kernel_source_vs_oracle_and_exact(emu):
    # the structure of /root/reference/test/periodic.jl:1-75 (tree, pair observations, TruncBondThresh(10))
    g, h, T, N, phi, psi = _glauber_tree()
    wo = [[OF.HomogeneousGlauberFactor(1.0, float(h[i]), 1.0)] * (T + 1) for i in range(N)]
    wp = [[PF.HomogeneousGlauberFactor(1.0, float(h[i]), 1.0)] * (T + 1) for i in range(N)]
    bo = P.PeriodicMPBP(g, wo, [2] * N, T, phi=phi, psi=psi)
    P.iterate(bo, maxiter=6, trunc=tt.TruncBondThresh(10))
    hb = HostPeriodicBP(emu, g, wp, [2] * N, T, phi, psi, dmax=10)
    for _ in range(6):
        for i in range(N):
            hb.update(i, Tr(2, 10, 0.0))
    assert np.allclose(np.array(P.beliefs(bo)), np.array(hb.marg), atol=1e-9)
    assert np.allclose(bo.f, hb.f, atol=1e-9)
    p, logZ = P.exact_prob(bo)
    assert abs(-hb.f.sum() - logZ) < 1e-9
    # pair beliefs (test/periodic.jl:49-60) against the oracle and against brute force; logz sums to the same free energy
    pb_o, lz_o = P.pair_beliefs(bo)
    pb_h, lz_h = hb.pair_beliefs()
    L = T + 1
    for e in range(g.ne):
        assert np.allclose(np.array(pb_o[e]), np.array(pb_h[e]), atol=1e-9)
        i, j = g.src[e], g.dst[e]
        for t in range(L):
            ex = p.sum(axis=tuple(a for a in range(N * L) if a not in (i * L + t, j * L + t)))
            assert np.allclose(pb_h[e][t], ex if i < j else ex.T, atol=1e-9)
    assert np.allclose(lz_o, lz_h, atol=1e-9)
    # two-time marginals b_i(x^t, x^u) of every belief (autocorrelations of test/periodic.jl:43-47) against brute force
    for i in range(N):
        for t in range(L):
            for u in range(t + 1, L):
                ex = p.sum(axis=tuple(a for a in range(N * L) if a not in (i * L + t, i * L + u)))
                assert np.allclose(hb.tv[i][t, u, :4].reshape(2, 2, order="F"), ex, atol=1e-9)
    for e in range(g.ne):
        A = tt.TT(hb.message(e))
        assert abs(P.lognormalization(A)) < 1e-10
        assert max(max(a.shape[:2]) for a in A) <= 10
        # same function on the ring as the oracle's message
        rng = np.random.default_rng(e)
        for _ in range(5):
            x = [tuple(rng.integers(0, 2, size=2)) for _ in range(T + 1)]
            assert abs(P.evaluate(A, x) - P.evaluate(bo.mu[e], x)) < 1e-9


def test_periodic_sirs_binding_truncation_matches_the_oracle_sweeps(emu):
    # a BINDING truncation on a loopy graph: the kernel performs the same ring sweeps as oracle/periodic.py (whose sweep order
    # is itself unpinned against TensorTrains.jl, see its header), so the two agree to rounding as long as no singular
    # values cross; q = 3, time-dependent factors, degree 3
    rng = np.random.default_rng(5)
    T, N = 2, 4
    und = [(0, 1), (1, 2), (2, 0), (2, 3)]
    g = O.BiDiGraph(N, und)
    par = [[(0.3 + 0.1 * t, 0.2, 0.15, 0.05) for t in range(T + 1)] for _ in range(N)]
    wo = [[OF.SIRSFactor(*par[i][t]) for t in range(T + 1)] for i in range(N)]
    wp = [[PF.SIRSFactor(*par[i][t]) for t in range(T + 1)] for i in range(N)]
    phi = [[0.2 + rng.random(3) for _ in range(T + 1)] for _ in range(N)]
    psi = [[np.ones((3, 3)) for _ in range(T + 1)] for _ in range(g.ne)]
    bo = P.PeriodicMPBP(g, wo, [3] * N, T, phi=phi, psi=psi)
    P.iterate(bo, maxiter=3, trunc=tt.TruncBond(3))
    hb = HostPeriodicBP(emu, g, wp, [3] * N, T, phi, psi, dmax=3)
    for _ in range(3):
        for i in range(N):
            hb.update(i, Tr(0, 3, 0.0))
    assert np.allclose(np.array(P.beliefs(bo)), np.array(hb.marg), atol=1e-8)
    assert np.allclose(bo.f, hb.f, atol=1e-8)


def test_periodic_sis_tree_truncthresh0_vs_exact(emu):
    # oracle twin: tests/test_oracle_golden.py::test_periodic_sis_tree_vs_exact_and_ring_invariances (T = 3, TruncThresh(0.0))
    T, N = 3, 3
    g = O.BiDiGraph(N, [(0, 1), (1, 2)])
    wo = [[OF.SISFactor(0.3, 0.25, 0.05)] * (T + 1) for _ in range(N)]
    wp = [[PF.SISFactor(0.3, 0.25, 0.05)] * (T + 1) for _ in range(N)]
    phi = [[np.ones(2) for _ in range(T + 1)] for _ in range(N)]
    phi[0][1] = np.array([0.2, 1.0])
    phi[2][3] = np.array([1.0, 0.4])
    psi = [[np.ones((2, 2)) for _ in range(T + 1)] for _ in range(g.ne)]
    bo = P.PeriodicMPBP(g, wo, [2] * N, T, phi=phi, psi=psi)
    P.iterate(bo, maxiter=6, trunc=tt.TruncThresh(0.0))
    hb = HostPeriodicBP(emu, g, wp, [2] * N, T, phi, psi, dmax=12)
    for _ in range(6):
        for i in range(N):
            hb.update(i, Tr(1, 0, 0.0))
    p, logZ = P.exact_prob(bo)
    L = T + 1
    be = [[p.sum(axis=tuple(a for a in range(N * L) if a != i * L + t)) for t in range(L)] for i in range(N)]
    assert np.allclose(np.array(hb.marg), np.array(be), atol=1e-9)
    assert abs(-hb.f.sum() - logZ) < 1e-9
    assert np.allclose(np.array(P.beliefs(bo)), np.array(hb.marg), atol=1e-9)


def test_periodic_sirs_tree_q3_vs_exact(emu):
    # q = 3 (SIRS), degree-2 centre, random reweightings and pair observations: kernel source vs brute force and the oracle
    rng = np.random.default_rng(9)
    T, N = 2, 3
    g = O.BiDiGraph(N, [(0, 1), (1, 2)])
    par = (0.35, 0.2, 0.25, 0.05)
    wo = [[OF.SIRSFactor(*par)] * (T + 1) for _ in range(N)]
    wp = [[PF.SIRSFactor(*par)] * (T + 1) for _ in range(N)]
    phi = [[0.2 + rng.random(3) for _ in range(T + 1)] for _ in range(N)]
    psi = [[np.ones((3, 3)) for _ in range(T + 1)] for _ in range(g.ne)]
    m = 0.3 + rng.random((3, 3))
    for e in range(g.ne):
        if (g.src[e], g.dst[e]) == (0, 1):
            psi[e][1] = m
        if (g.src[e], g.dst[e]) == (1, 0):
            psi[e][1] = m.T
    bo = P.PeriodicMPBP(g, wo, [3] * N, T, phi=phi, psi=psi)
    P.iterate(bo, maxiter=5, trunc=tt.TruncBondThresh(12, 1e-13))
    hb = HostPeriodicBP(emu, g, wp, [3] * N, T, phi, psi, dmax=12)
    for _ in range(5):
        for i in range(N):
            hb.update(i, Tr(2, 12, 1e-13))
    p, logZ = P.exact_prob(bo)
    L = T + 1
    be = [[p.sum(axis=tuple(a for a in range(N * L) if a != i * L + t)) for t in range(L)] for i in range(N)]
    assert np.allclose(np.array(hb.marg), np.array(be), atol=1e-9)
    assert abs(-hb.f.sum() - logZ) < 1e-9
    assert np.allclose(np.array(P.beliefs(bo)), np.array(hb.marg), atol=1e-9)
    pb, _ = hb.pair_beliefs()
    for e in range(g.ne):
        i, j = g.src[e], g.dst[e]
        for t in range(L):
            ex = p.sum(axis=tuple(a for a in range(N * L) if a not in (i * L + t, j * L + t)))
            assert np.allclose(pb[e][t], ex if i < j else ex.T, atol=1e-9)
