"""ORACLE (test infrastructure only -- never imported by the product path).

numpy restatement of the MPBP message-update hot path of MatrixProductBP.jl:

* MPBP struct / constructor / reset     /root/reference/src/mpbp.jl:1-102
* recursive node update                 /root/reference/src/recursive_bp_factor.jl:64-179
* generic (exhaustive trace) update     /root/reference/src/bp_core.jl:18-109, src/mpbp.jl:117-154
* MPEM3, mpem2, marginalize             /root/reference/src/mpems.jl:27-94
* iterate!, CB_BP, beliefs, pair_beliefs, bethe_free_energy   /root/reference/src/mpbp.jl:157-237,298
* InfiniteRegularGraph                  /root/reference/src/infinite_graph.jl:8-43
* CavityTools.cavity (un-vendored, compat "0.3, 1"): restated as in SURVEY.md section 8c.

Pinned by tests/test_oracle_golden.py (reference golden vector + brute force).
"""
from __future__ import annotations

import numpy as np

from . import tt as T_
from .tt import TT


# --------------------------------------------------------------------------------------
# graph (IndexedGraphs.IndexedBiDiGraph restated: directed edge index = position in the CSC of
# the adjacency matrix, source = column, destination = row  ->  sorted by (src, dst))
# --------------------------------------------------------------------------------------
class BiDiGraph:
    def __init__(self, n, undirected_edges):
        und = sorted({(min(a, b), max(a, b)) for a, b in undirected_edges if a != b})
        self.N = int(n)
        self.undirected = und
        dire = sorted([(a, b) for a, b in und] + [(b, a) for a, b in und])
        self.src = [e[0] for e in dire]
        self.dst = [e[1] for e in dire]
        idx = {e: k for k, e in enumerate(dire)}
        self.rev = [idx[(b, a)] for a, b in dire]
        self.out_edges = [[] for _ in range(n)]
        for k, (a, b) in enumerate(dire):
            self.out_edges[a].append(k)
        # in-edges in the same neighbour order as out-edges (recursive_bp_factor.jl:149-152)
        self.in_edges = [[self.rev[k] for k in self.out_edges[i]] for i in range(n)]
        self.infinite_k = None

    @property
    def ne(self):
        return len(self.src)

    def degree(self, i):
        return len(self.out_edges[i])


class InfiniteRegularGraph:
    """one node, one stored message repeated k times (infinite_graph.jl:8-20)."""

    def __init__(self, k):
        self.N = 1
        self.infinite_k = int(k)
        self.src = [0]
        self.dst = [0]
        self.rev = [0]
        self.out_edges = [[0] * k]
        self.in_edges = [[0] * k]

    @property
    def ne(self):
        return 1

    def degree(self, i):
        return self.infinite_k


class InfiniteBipartiteRegularGraph:
    """two node classes of degrees k = (kA, kB); edge e (0-based) is the message INTO node e from the other class
    (infinite_graph.jl:62-85: inedges(g, i) carry idx i, outedges idx 3-i)."""

    def __init__(self, k):
        self.N = 2
        self.k = (int(k[0]), int(k[1]))
        self.infinite_k = None
        self.bipartite_k = self.k
        self.src = [1, 0]
        self.dst = [0, 1]
        self.rev = [1, 0]
        self.in_edges = [[0] * self.k[0], [1] * self.k[1]]
        self.out_edges = [[1] * self.k[0], [0] * self.k[1]]

    @property
    def ne(self):
        return 2

    def degree(self, i):
        return self.k[i]


def mpbp_infinite_bipartite_graph(k, w, q, phi=None, psi=None, d=1):
    """infinite_graph.jl:87-105"""
    T = len(w[0]) - 1
    return MPBP(InfiniteBipartiteRegularGraph(k), [list(w[0]), list(w[1])], list(q), T, phi=phi, psi=psi, d=d)


class MPBP:
    def __init__(self, g, w, q, T, phi=None, psi=None, d=1):
        self.g, self.w, self.q, self.T = g, w, list(q), int(T)
        L = T + 1
        N = g.N
        self.phi = phi if phi is not None else [[np.ones(q[i]) for _ in range(L)] for i in range(N)]
        self.psi = (
            psi
            if psi is not None
            else [[np.ones((q[g.src[e]], q[g.dst[e]])) for _ in range(L)] for e in range(g.ne)]
        )
        bonds = [1] + [d] * T + [1]
        self.mu = [T_.flat_tt(bonds, q[g.src[e]], q[g.dst[e]]) for e in range(g.ne)]
        self.b = [T_.flat_tt(bonds, q[i]) for i in range(N)]
        self.f = np.zeros(N)

    def reset_messages(self):
        for A in self.mu:
            for t in range(len(A)):
                A[t] = np.ones_like(A[t])
            A.ls = 0.0
            T_.normalize(A)


def mpbp_infinite_graph(k, w, q, phi=None, psi=None, d=1):
    T = len(w) - 1
    g = InfiniteRegularGraph(k)
    bp = MPBP(g, [w], [q], T, phi=[phi] if phi is not None else None, psi=[psi] if psi is not None else None, d=d)
    return bp


# --------------------------------------------------------------------------------------
# cavity (CavityTools.cavity restated)
# --------------------------------------------------------------------------------------
def cavity(src, op, init):
    n = len(src)
    if n == 0:
        return [], init
    if n == 1:
        return [init], op(src[0], init)
    dest = [src[0]]
    for i in range(1, n):
        dest.append(op(dest[i - 1], src[i]))
    full = op(dest[n - 1], init)
    right = init
    for i in range(n - 1, 0, -1):
        dest[i] = op(dest[i - 1], right)
        right = op(src[i], right)
    dest[0] = right
    return dest, full


# --------------------------------------------------------------------------------------
# recursive node update
# --------------------------------------------------------------------------------------
def _tab(fn, *dims):
    out = np.zeros(dims)
    for idx in np.ndindex(*dims):
        out[idx] = fn(*[i + 1 for i in idx])
    return out


def compute_prob_ys(wi, qi, mu_in, psi_out, T, trunc):
    """recursive_bp_factor.jl:104-143"""
    L = T + 1
    B = []
    for k in range(len(psi_out)):
        tens = []
        for t in range(L):
            w, m, psi = wi[t], mu_in[k][t], psi_out[k][t]
            Pxy = _tab(lambda y, xk, xi: w.prob_xy(y, xk, xi, k + 1) * psi[xi - 1, xk - 1], w.nstates(1), m.shape[2], qi)
            tens.append(np.einsum("ykx,mnkx->mnyx", Pxy, m))
        B.append((TT(tens, mu_in[k].ls), 1))

    def op(a, b):
        (B1, d1), (B2, d2) = a, b
        tens = []
        for t in range(L):
            w, b1, b2 = wi[t], B1[t], B2[t]
            Pyy = _tab(lambda y, y1, y2, xi: w.prob_yy(y, y1, y2, xi, d1, d2), w.nstates(d1 + d2), b1.shape[2], b2.shape[2], b1.shape[3])
            B3 = np.einsum("yabx,mnax,opbx->monpyx", Pyy, b1, b2, optimize=True)
            s = B3.shape
            tens.append(B3.reshape(s[0] * s[1], s[2] * s[3], s[4], s[5], order="F"))
        Bout = TT(tens, B1.ls + B2.ls)
        T_.compress(Bout, trunc)
        T_.normalize_eachmatrix(Bout)
        return Bout, d1 + d2

    Minit = [
        _tab(lambda a, b, y, xi: wi[t].prob_y0(y, xi), 1, 1, wi[t].nstates(0), qi) for t in range(L)
    ]
    init = (TT(Minit), 0)
    dest, full = cavity(B, op, init)
    C = [d[0] for d in dest]
    return C, full[0], B


def f_bp_partial(A, wi, phii, d, prob, qj, j):
    """_f_bp_partial, recursive_bp_factor.jl:73-87.  prob(w, x', x, xj, y, d, j)"""
    q = len(phii[0])
    L = len(A)
    Bs = []
    for t in range(L - 1):
        At = A[t]
        W = _tab(lambda xn, x, xj, y: prob(wi[t], xn, x, xj, y, d, j) * phii[t][x - 1], q, q, qj, At.shape[2])
        Bs.append(np.einsum("zxjy,mnyx->mnxjz", W, At))
    AT = A[L - 1]
    last = np.einsum("mnyx,x->mnx", AT, phii[L - 1])
    Bs.append(np.broadcast_to(last[:, :, :, None, None], last.shape + (qj, q)).copy())
    return Bs, A.ls


def mpem2(Bs, ls):
    """mpems.jl:67-94: L->R sweep of un-truncated SVDs turning an MPEM3 into an MPEM2."""
    L = len(Bs)
    q, qj = Bs[0].shape[2], Bs[0].shape[3]
    C = [None] * L
    logc = 0.0
    Bnew = Bs[0]
    for t in range(L - 1):
        m, n = Bnew.shape[0], Bnew.shape[1]
        # M[(xi,xj,m),(n,xi')]
        M = np.transpose(Bnew, (2, 3, 0, 1, 4)).reshape(q * qj * m, n * q, order="F")
        mt = np.max(np.abs(M))
        if np.isfinite(mt) and mt != 0:
            M = M / mt
            logc += np.log(mt)
        U, lam, Vt = np.linalg.svd(M, full_matrices=False)
        r = len(lam)
        C[t] = np.transpose(U.reshape(q, qj, m, r, order="F"), (2, 3, 0, 1))
        V3 = Vt.reshape(r, n, q, order="F")
        Bnew = np.einsum("m,mlx,lnxyz->mnxyz", lam, V3, Bs[t + 1])
    C[L - 1] = Bnew[:, :, :, :, 0]
    return TT(C, ls + logc)


def marginalize(A: TT):
    return TT([a.sum(axis=3) for a in A], A.ls)


def set_msg(mu_old, muj, damp, trunc):
    """recursive_bp_factor.jl:168-179"""
    logz = T_.normalize(muj)
    if damp > 0:
        muj = T_.tt_sum(muj, mu_old, damp / (1 - damp))
        T_.compress(muj, trunc)
        T_.normalize(muj)
    return muj, logz


def onebpiter_recursive(bp: MPBP, i, trunc, damp=0.0, mu_read=None, mu_write=None):
    """recursive_bp_factor.jl:146-165.  mu_read / mu_write allow a Jacobi (double-buffered) schedule."""
    g = bp.g
    mu_read = bp.mu if mu_read is None else mu_read
    mu_write = bp.mu if mu_write is None else mu_write
    ein, eout = g.in_edges[i], g.out_edges[i]
    wi, phii, di = bp.w[i], bp.phi[i], len(ein)
    qi = bp.q[i]
    mu_in = [mu_read[e] for e in ein]  # snapshot (references)
    C, full, _ = compute_prob_ys(wi, qi, mu_in, [bp.psi[e] for e in eout], bp.T, trunc)
    sumlogz = 0.0
    for j, e in enumerate(eout):
        qj = bp.q[g.dst[e]]
        Bs, ls = f_bp_partial(C[j], wi, phii, di - 1, lambda w, *a: w.prob_y_partial(*a), qj, j + 1)
        muj = T_.compress(mpem2(Bs, ls), trunc, "left")
        T_.normalize_eachmatrix(muj)
        muj, lz = set_msg(mu_write[e], muj, damp, trunc)
        mu_write[e] = muj
        sumlogz += lz
    Bs, ls = f_bp_partial(full, wi, phii, di, lambda w, *a: w.prob_y_dummy(*a), 1, 1)
    bp.b[i] = marginalize(mpem2(Bs, ls))
    logzi = T_.normalize(bp.b[i])
    bp.f[i] = (di / 2 - 1) * logzi - 0.5 * sumlogz


# --------------------------------------------------------------------------------------
# generic (exhaustive) node update
# --------------------------------------------------------------------------------------
def _kron_slices(mats):
    out = np.ones((1, 1))
    for m in mats:
        # Julia kron(A,B): index (a,b) with b fastest; bond fusion order is irrelevant for the function
        out = np.kron(out, m)
    return out


def f_bp(A, wi, phii, psi_out, j_index, qj):
    """bp_core.jl:18-57 (j_index 0-based; j_index=None -> dummy neighbour, bp_core.jl:60-93)"""
    import itertools

    L = len(phii)
    q = len(phii[0])
    dummy = j_index is None
    notj = [k for k in range(len(A)) if dummy or k != j_index]
    qn = [psi_out[k][0].shape[1] for k in notj]
    Bs = []
    for t in range(L):
        ml = int(np.prod([A[k][t].shape[0] for k in notj])) if notj else 1
        nr = int(np.prod([A[k][t].shape[1] for k in notj])) if notj else 1
        Bt = np.zeros((ml, nr, q, 1 if dummy else qj, q))
        for xi in range(q):
            for xn in itertools.product(*[range(v) for v in qn]):
                At = _kron_slices([A[k][t][:, :, xk, xi] * psi_out[k][t][xi, xk] for k, xk in zip(notj, xn)])
                for xj in range(1 if dummy else qj):
                    for xnext in range(q):
                        wgt = phii[t][xi]
                        if t < L - 1:
                            if dummy:
                                xfull = [v + 1 for v in xn]
                            else:
                                xfull = [v + 1 for v in xn[:j_index]] + [xj + 1] + [v + 1 for v in xn[j_index:]]
                            wgt = wgt * wi[t](xnext + 1, xfull, xi + 1)
                        if wgt != 0:
                            Bt[:, :, xi, xj, xnext] += At * wgt
        Bs.append(Bt)
    ls = sum(A[k].ls for k in notj)
    return Bs, ls


def onebpiter_generic(bp: MPBP, i, trunc, mu_read=None, mu_write=None):
    """mpbp.jl:117-154"""
    g = bp.g
    mu_read = bp.mu if mu_read is None else mu_read
    mu_write = bp.mu if mu_write is None else mu_write
    ein, eout = g.in_edges[i], g.out_edges[i]
    A = [mu_read[e] for e in ein]
    psi_out = [bp.psi[e] for e in eout]
    sumlogz = 0.0
    for j, e in enumerate(eout):
        Bs, ls = f_bp(A, bp.w[i], bp.phi[i], psi_out, j, bp.q[g.dst[e]])
        muj = T_.compress(mpem2(Bs, ls), trunc, "left")
        sumlogz += T_.normalize(muj)
        mu_write[e] = muj
    di = len(ein)
    Bs, ls = f_bp(A, bp.w[i], bp.phi[i], psi_out, None, 1)
    bp.b[i] = marginalize(T_.compress(mpem2(Bs, ls), trunc, "left"))
    logzi = T_.lognormalization(bp.b[i])
    bp.f[i] = (di / 2 - 1) * logzi - 0.5 * sumlogz


def onebpiter(bp, i, trunc, damp=0.0, mu_read=None, mu_write=None):
    if bp.w[i][0].recursive:
        onebpiter_recursive(bp, i, trunc, damp, mu_read, mu_write)
    else:
        onebpiter_generic(bp, i, trunc, mu_read, mu_write)


# --------------------------------------------------------------------------------------
# driver and read-outs
# --------------------------------------------------------------------------------------
def means(bp, f=lambda x, i: x):
    """mpbp.jl:257-261 with states numbered from 1 as in Julia."""
    out = []
    for i in range(bp.g.N):
        out.append([sum(f(x + 1, i) * p[x] for x in range(len(p))) for p in T_.marginals(bp.b[i])])
    return out


def iterate(bp: MPBP, maxiter=5, trunc=None, tol=1e-10, damp=0.0, nodes=None, schedule="sequential", f=lambda x, i: x):
    """mpbp.jl:185-198 with shuffle_nodes=false and one thread (schedule='sequential': in-place
    Gauss-Seidel sweep in `nodes` order) or the Jacobi variant (schedule='parallel': every node
    reads the messages of the previous iteration)."""
    trunc = trunc if trunc is not None else T_.TruncThresh(1e-6)
    nodes = list(range(bp.g.N)) if nodes is None else list(nodes)
    m_old = means(bp, f)
    deltas = []
    for it in range(1, maxiter + 1):
        if schedule == "sequential":
            for i in nodes:
                onebpiter(bp, i, trunc, damp)
        else:
            new = list(bp.mu)
            for i in nodes:
                onebpiter(bp, i, trunc, damp, mu_read=bp.mu, mu_write=new)
            bp.mu = new
        m_new = means(bp, f)
        delta = max(max(abs(a - b) for a, b in zip(mn, mo)) for mn, mo in zip(m_new, m_old))
        deltas.append(delta)
        m_old = m_new
        if delta < tol:
            return it, deltas
    return maxiter, deltas


def beliefs(bp):
    return [T_.marginals(b) for b in bp.b]


def beliefs_tu(bp, maxdist=None):
    """mpbp.jl:239: two-time marginals b_i(x^t, x^u), t < u <= t + maxdist, of every node's belief; [i][t][u] is a
    q x q array (None where not computed)."""
    out = []
    for b in bp.b:
        tv = T_.twovar_marginals(b)
        L = len(tv)
        if maxdist is not None:
            for t in range(L):
                for u in range(L):
                    if u - t > maxdist:
                        tv[t][u] = None
        out.append(tv)
    return out


def autocorrelations(bp, f=lambda x, i: x, maxdist=None):
    """mpbp.jl:245-255: r_i[t, u] = <f(x_i^t) f(x_i^u)> for t < u (zero elsewhere), states numbered from 1."""
    out = []
    for i, tv in enumerate(beliefs_tu(bp, maxdist)):
        L = len(tv)
        q = bp.q[i]
        fx = np.array([f(x + 1, i) for x in range(q)], dtype=float)
        r = np.zeros((L, L))
        for t in range(L):
            for u in range(t + 1, L):
                if tv[t][u] is not None:
                    r[t, u] = fx @ np.asarray(tv[t][u]).reshape(q, q) @ fx
        out.append(r)
    return out


def autocovariances(bp, f=lambda x, i: x, maxdist=None):
    """mpbp.jl:289-296: covariance.(r, mu) = r - mu mu' (on the whole matrix, like the reference)."""
    mu = means(bp, f)
    return [r - np.outer(m, m) for r, m in zip(autocorrelations(bp, f, maxdist), mu)]


def bethe_free_energy(bp):
    k = getattr(bp.g, "bipartite_k", None)
    if k is not None:  # infinite_graph.jl:120-122
        return float((bp.f[0] * k[1] + bp.f[1] * k[0]) / (k[0] + k[1]))
    return float(np.sum(bp.f))


def pair_belief_tt(Aij, Aji, psi):
    """bp_core.jl:95-101"""
    tens = []
    for a, b, p in zip(Aij, Aji, psi):
        c = np.einsum("acij,bdji,ij->abcdij", a, b, p)
        s = c.shape
        tens.append(c.reshape(s[0] * s[1], s[2] * s[3], s[4], s[5], order="F"))
    return TT(tens, Aij.ls + Aji.ls)


def pair_beliefs(bp):
    """mpbp.jl:202-235 (+ infinite_graph.jl:37-43).  returns (b[e][t][xs,xt], logz[i])"""
    g = bp.g
    logz = np.zeros(g.N)
    b = [None] * g.ne
    if getattr(g, "bipartite_k", None) is not None:
        # infinite_graph.jl:110-118
        for i in range(2):
            P = pair_belief_tt(bp.mu[i], bp.mu[1 - i], bp.psi[i])
            b[i] = T_.marginals(P)
            logz[i] = (1 / (g.bipartite_k[i] - 1) - 0.5) * T_.lognormalization(P)
        return b, logz
    if g.infinite_k is not None:
        Aij = bp.mu[0]
        P = pair_belief_tt(Aij, Aij, bp.psi[0])
        b[0] = T_.marginals(P)
        logz[0] = (1 / (g.infinite_k - 1) - 0.5) * T_.lognormalization(P)
        return b, logz
    for e in range(g.ne):
        j = g.dst[e]
        P = pair_belief_tt(bp.mu[e], bp.mu[g.rev[e]], bp.psi[e])
        b[e] = T_.marginals(P)
        logz[j] += (1 / g.degree(j) - 0.5) * T_.lognormalization(P)
    return b, logz


def alternate_marginals(bp):
    """mpbp.jl:270-280: p(x_i^t, x_j^{t+1}) for every directed edge i->j and t = 0..T-1, from the two-time marginals
    of the pair-belief MPEM (twovar_marginals(pb)[t, t+1] summed over x_j^t and x_i^{t+1})."""
    g = bp.g
    L = bp.T + 1
    out = []
    for e in range(g.ne):
        qi, qj = bp.q[g.src[e]], bp.q[g.dst[e]]
        P = pair_belief_tt(bp.mu[e], bp.mu[g.rev[e]], bp.psi[e])
        tv = T_.twovar_marginals(P)
        out.append([np.asarray(tv[t][t + 1]).reshape(qi, qj, qi, qj).sum(axis=(1, 2)) for t in range(L - 1)])
    return out
