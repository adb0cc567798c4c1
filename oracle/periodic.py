"""TEST INFRASTRUCTURE ONLY -- oracle restatement of the periodic-in-time variant.

Periodic MPEMs close the matrix product with a trace instead of 1x1 boundaries; the dynamics wraps around: the factor
at the last time maps (x_neigh^T, x_i^T) to x_i^0.  Reference code followed here:
  /root/reference/src/mpems.jl:96-155            PeriodicMPEM3, evaluate (trace), mpem2(::PeriodicMPEM3)
  /root/reference/src/recursive_bp_factor.jl:89-101  _f_bp_partial for PeriodicMPEM2 (every site carries the factor)
  /root/reference/src/mpbp.jl:399-409            periodic_mpbp (flat periodic messages / beliefs)
  /root/reference/test/periodic.jl:1-75          Glauber on a small tree with pair observations vs brute force

PARITY STATUS.  The ring versions of compress! / orthogonalize! / normalize! / marginals live in the un-vendored
TensorTrains.jl (PeriodicTensorTrain); their sweep order around the ring could not be recalled with confidence, so the
restatement below (full ring sweeps: every bond, including the one closing the ring, is visited once per sweep) is
"parity unpinned" whenever a truncation actually binds.  With a non-binding truncation -- what test/periodic.jl uses
(TruncBondThresh(10) on a tree) -- every valid compression is exact, and this module is pinned against brute-force
enumeration of the periodic dynamics (tests/test_oracle_golden.py::test_periodic_*).  Device counterpart:
matrixproductbp.jl_b200/csrc/periodic.cuh (the same full-turn sweeps), checked against this module by
tests/test_periodic_host_emul.py (CPU emulation of the kernel source) and tests/test_gpu_periodic.py.
"""
from __future__ import annotations

import numpy as np

from . import tt as T_
from .mpbp import _tab, cavity, marginalize
from .tt import TT, TruncThresh, _site_sum


# ---------------------------------------------------------------------------------------------
# ring tensor trains
# ---------------------------------------------------------------------------------------------
def flat_periodic_tt(L, d, *q):
    A = TT([np.ones((d, d) + tuple(q)) for _ in range(L)])
    normalize(A)
    return A


def evaluate(A: TT, x):
    M = np.eye(A[0].shape[0])
    for a, xt in zip(A, x):
        M = M @ a[(slice(None), slice(None)) + tuple(np.atleast_1d(xt))]
    return float(np.trace(M) * np.exp(A.ls))


def lognormalization(A: TT) -> float:
    M, acc = None, A.ls
    for a in A:
        S = _site_sum(a)
        M = S if M is None else M @ S
        m = np.max(np.abs(M))
        if np.isfinite(m) and m != 0.0:
            M = M / m
            acc += np.log(m)
    return acc + np.log(np.abs(np.trace(M)))


def normalize(A: TT) -> float:
    lz = lognormalization(A)
    f = np.exp((A.ls - lz) / len(A))
    for t in range(len(A)):
        A[t] = A[t] * f
    A.ls = 0.0
    return lz


def marginals(A: TT):
    L = len(A)
    S = [_site_sum(a) for a in A]

    def scaled(M):
        m = np.max(np.abs(M))
        return M / m if np.isfinite(m) and m != 0.0 else M

    pre = [np.eye(A[0].shape[0])]
    for t in range(L - 1):
        pre.append(scaled(pre[-1] @ S[t]))
    suf = [None] * L
    suf[L - 1] = np.eye(A[L - 1].shape[1])
    for t in range(L - 2, -1, -1):
        suf[t] = scaled(S[t + 1] @ suf[t + 1])
    out = []
    for t in range(L):
        a = A[t].reshape(A[t].shape[0], A[t].shape[1], -1)
        p = np.einsum("im,mnx,ni->x", pre[t], a, suf[t])
        out.append((p / p.sum()).reshape(A[t].shape[2:]))
    return out


def orthogonalize_right(A: TT, trunc):
    """one full turn around the ring, right to left: A[t] <- V', carry U diag(lam) into the site on its left
    (site 0 carries into site L-1 across the closing bond)."""
    L = len(A)
    for t in range(L - 1, -1, -1):
        C = A[t]
        M = C.reshape(C.shape[0], -1, order="F")
        m = np.max(np.abs(M))
        if np.isfinite(m) and m != 0.0:
            M = M / m
            A.ls += np.log(m)
        U, lam, Vt = trunc(M)
        A[t] = Vt.reshape((Vt.shape[0],) + C.shape[1:], order="F")
        p = (t - 1) % L
        A[p] = np.einsum("mk...,kr->mr...", A[p], U * lam[None, :])
    return A


def orthogonalize_left(A: TT, trunc):
    L = len(A)
    for t in range(L):
        C = A[t]
        m_, n_ = C.shape[:2]
        phys = C.shape[2:]
        P = int(np.prod(phys)) if len(phys) else 1
        M = np.moveaxis(C.reshape(m_, n_, P, order="F"), 1, 2).reshape(m_ * P, n_, order="F")
        mx = np.max(np.abs(M))
        if np.isfinite(mx) and mx != 0.0:
            M = M / mx
            A.ls += np.log(mx)
        U, lam, Vt = trunc(M)
        r = len(lam)
        A[t] = np.moveaxis(U.reshape(m_, P, r, order="F"), 1, 2).reshape((m_, r) + phys, order="F")
        nx = (t + 1) % L
        A[nx] = np.einsum("rl,ln...->rn...", lam[:, None] * Vt, A[nx])
    return A


def compress(A: TT, trunc, is_orthogonal="none"):
    if is_orthogonal == "none":
        orthogonalize_right(A, TruncThresh(0.0))
        orthogonalize_left(A, trunc)
    elif is_orthogonal == "left":
        orthogonalize_right(A, trunc)
    else:
        orthogonalize_left(A, trunc)
    return A


# ---------------------------------------------------------------------------------------------
# BP on periodic messages
# ---------------------------------------------------------------------------------------------
class PeriodicMPBP:
    def __init__(self, g, w, q, T, phi=None, psi=None, d=1):
        self.g, self.w, self.q, self.T = g, w, list(q), int(T)
        L, N = T + 1, g.N
        self.phi = phi if phi is not None else [[np.ones(q[i]) for _ in range(L)] for i in range(N)]
        self.psi = psi if psi is not None else [[np.ones((q[g.src[e]], q[g.dst[e]])) for _ in range(L)] for e in range(g.ne)]
        self.mu = [flat_periodic_tt(L, d, q[g.src[e]], q[g.dst[e]]) for e in range(g.ne)]
        self.b = [flat_periodic_tt(L, d, q[i]) for i in range(N)]
        self.f = np.zeros(N)


def compute_prob_ys(wi, qi, mu_in, psi_out, T, trunc):
    """recursive_bp_factor.jl:104-143 on ring tensors (the Kronecker product also squares the closing bond)"""
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
        compress(Bout, trunc)
        T_.normalize_eachmatrix(Bout)
        return Bout, d1 + d2

    init = (TT([_tab(lambda a, b, y, xi: wi[t].prob_y0(y, xi), 1, 1, wi[t].nstates(0), qi) for t in range(L)]), 0)
    dest, full = cavity(B, op, init)
    return [d[0] for d in dest], full[0]


def f_bp_partial(A, wi, phii, d, prob, qj, j):
    """recursive_bp_factor.jl:89-101: every site carries the factor; x^{t+1} of the last site is x^0"""
    q = len(phii[0])
    Bs = []
    for t in range(len(A)):
        W = _tab(lambda xn, x, xj, y: prob(wi[t], xn, x, xj, y, d, j) * phii[t][x - 1], q, q, qj, A[t].shape[2])
        Bs.append(np.einsum("zxjy,mnyx->mnxjz", W, A[t]))
    return Bs, A.ls


def mpem2(Bs, ls):
    """mpems.jl:123-155 (mpem2 of a PeriodicMPEM3): un-truncated L->R SVD sweep; the last carry closes onto site 0"""
    L = len(Bs)
    q, qj = Bs[0].shape[2], Bs[0].shape[3]
    C = [None] * L
    logc = 0.0
    Bnew = Bs[0]
    for t in range(L):
        m, n = Bnew.shape[0], Bnew.shape[1]
        M = np.transpose(Bnew, (2, 3, 0, 1, 4)).reshape(q * qj * m, n * q, order="F")
        mt = np.max(np.abs(M))
        if np.isfinite(mt) and mt != 0:
            M = M / mt
            logc += np.log(mt)
        U, lam, Vt = np.linalg.svd(M, full_matrices=False)
        r = len(lam)
        C[t] = np.transpose(U.reshape(q, qj, m, r, order="F"), (2, 3, 0, 1))
        V3 = Vt.reshape(r, n, q, order="F")
        if t < L - 1:
            Bnew = np.einsum("m,mlx,lnxyz->mnxyz", lam, V3, Bs[t + 1])
        else:
            C[0] = np.einsum("m,mkx,knxy->mnxy", lam, V3, C[0])
    return TT(C, ls + logc)


def ring_sum(A: TT, B: TT, coefB: float) -> TT:
    """A + coefB * B as a ring train: every site block-diagonal, the closing bond included (TensorTrains._compose on
    PeriodicTensorTrains, as used by set_msg!, recursive_bp_factor.jl:172-176)"""
    L = len(A)
    fa, fb = np.exp(A.ls / L), np.exp(B.ls / L)
    out = []
    for t in range(L):
        a, b = A[t] * fa, B[t] * fb * (coefB if t == 0 else 1.0)
        c = np.zeros((a.shape[0] + b.shape[0], a.shape[1] + b.shape[1]) + a.shape[2:])
        c[: a.shape[0], : a.shape[1]] = a
        c[a.shape[0]:, a.shape[1]:] = b
        out.append(c)
    return TT(out, 0.0)


def onebpiter(bp: PeriodicMPBP, i, trunc, damp=0.0):
    """recursive_bp_factor.jl:146-165 on periodic messages; set_msg! damping :168-179"""
    g = bp.g
    ein, eout = g.in_edges[i], g.out_edges[i]
    wi, phii, di, qi = bp.w[i], bp.phi[i], len(ein), bp.q[i]
    C, full = compute_prob_ys(wi, qi, [bp.mu[e] for e in ein], [bp.psi[e] for e in eout], bp.T, trunc)
    sumlogz = 0.0
    for j, e in enumerate(eout):
        qj = bp.q[g.dst[e]]
        Bs, ls = f_bp_partial(C[j], wi, phii, di - 1, lambda w, *a: w.prob_y_partial(*a), qj, j + 1)
        muj = compress(mpem2(Bs, ls), trunc, "left")
        T_.normalize_eachmatrix(muj)
        sumlogz += normalize(muj)
        if damp > 0:
            muj = ring_sum(muj, bp.mu[e], damp / (1 - damp))
            compress(muj, trunc)
            normalize(muj)
        bp.mu[e] = muj
    Bs, ls = f_bp_partial(full, wi, phii, di, lambda w, *a: w.prob_y_dummy(*a), 1, 1)
    bp.b[i] = marginalize(mpem2(Bs, ls))
    logzi = normalize(bp.b[i])
    bp.f[i] = (di / 2 - 1) * logzi - 0.5 * sumlogz


def iterate(bp: PeriodicMPBP, maxiter=5, trunc=None, nodes=None, damp=0.0):
    trunc = trunc if trunc is not None else TruncThresh(1e-6)
    nodes = list(range(bp.g.N)) if nodes is None else list(nodes)
    for _ in range(maxiter):
        for i in nodes:
            onebpiter(bp, i, trunc, damp)
    return maxiter


def beliefs(bp):
    return [marginals(b) for b in bp.b]


def bethe_free_energy(bp):
    return float(np.sum(bp.f))


def pair_belief_tt(Aij, Aji, psi):
    """bp_core.jl:95-101 on ring trains: the bond-d^2 product of the two messages of an edge, reweighted by psi"""
    tens = []
    for a, b, p in zip(Aij, Aji, psi):
        t = np.einsum("acij,bdji,ij->abcdij", a, b, np.asarray(p))
        s = t.shape
        tens.append(t.reshape(s[0] * s[1], s[2] * s[3], s[4], s[5], order="F"))
    return TT(tens, Aij.ls + Aji.ls)


def pair_beliefs(bp):
    """mpbp.jl:202-235 with the ring marginals / normalisation: (b[e][t][x_src, x_dst], logz[i])"""
    g = bp.g
    logz = np.zeros(g.N)
    b = [None] * g.ne
    for e in range(g.ne):
        j = g.dst[e]
        Pt = pair_belief_tt(bp.mu[e], bp.mu[g.rev[e]], bp.psi[e])
        b[e] = marginals(Pt)
        logz[j] += (1 / g.degree(j) - 0.5) * lognormalization(Pt)
    return b, logz


# ---------------------------------------------------------------------------------------------
# brute force for the periodic dynamics
# ---------------------------------------------------------------------------------------------
def exact_prob(bp):
    """like oracle.exact.exact_prob with the wrap-around factor w_i^T(x_i^0 | x_neigh^T, x_i^T) included"""
    g, w, phi, psi, q = bp.g, bp.w, bp.phi, bp.psi, bp.q
    N, L = g.N, bp.T + 1
    dims = [q[i] for i in range(N) for _ in range(L)]
    grids = np.indices(dims).reshape(len(dims), -1)
    X = lambda i, t: grids[i * L + (t % L)]
    logp = np.zeros(grids.shape[1])
    neigh = [[g.dst[e] for e in g.out_edges[i]] for i in range(N)]
    with np.errstate(divide="ignore"):
        for i in range(N):
            for t in range(L):
                qn = [q[k] for k in neigh[i]]
                tab = np.zeros([q[i]] + qn + [q[i]])
                for idx in np.ndindex(*tab.shape):
                    tab[idx] = w[i][t](idx[0] + 1, [v + 1 for v in idx[1:-1]], idx[-1] + 1)
                logp += np.log(tab[(X(i, t + 1),) + tuple(X(k, t) for k in neigh[i]) + (X(i, t),)])
                logp += np.log(np.asarray(phi[i][t])[X(i, t)])
        for e in range(g.ne):
            for t in range(L):
                logp += 0.5 * np.log(np.asarray(psi[e][t])[X(g.src[e], t), X(g.dst[e], t)])
    mx = logp.max()
    logZ = mx + np.log(np.exp(logp - mx).sum())
    return np.exp(logp - logZ).reshape(dims), float(logZ)
