"""ORACLE (test infrastructure only).  Brute-force enumeration of the joint distribution,
restating /root/reference/src/exact.jl:5-41 (exact_prob), :43-75 (site marginals) and the
pair marginals used by every small-tree test of the reference (test/glauber_small_tree.jl:40-72).
Vectorised over the q^{N(T+1)} configurations with numpy; usable up to ~2e6 configurations."""
from __future__ import annotations

import itertools

import numpy as np


def exact_prob(bp):
    g, w, phi, psi, q = bp.g, bp.w, bp.phi, bp.psi, bp.q
    N, L = g.N, bp.T + 1
    dims = [q[i] for i in range(N) for _ in range(L)]  # variable (i,t) -> axis i*L+t
    Q = int(np.prod(dims))
    assert Q <= 4_000_000, f"{Q} configurations"
    grids = np.indices(dims).reshape(len(dims), -1)  # [var, config] 0-based
    X = lambda i, t: grids[i * L + t]
    logp = np.zeros(Q)
    neigh = [[g.dst[e] for e in g.out_edges[i]] for i in range(N)]
    with np.errstate(divide="ignore"):
        for i in range(N):
            logp += np.log(np.asarray(phi[i][0])[X(i, 0)])
            for t in range(L - 1):
                # tabulate w over (x', x_neigh..., x)
                qn = [q[k] for k in neigh[i]]
                tab = np.zeros([q[i]] + qn + [q[i]])
                for idx in np.ndindex(*tab.shape):
                    tab[idx] = w[i][t](idx[0] + 1, [v + 1 for v in idx[1:-1]], idx[-1] + 1)
                sel = (X(i, t + 1),) + tuple(X(k, t) for k in neigh[i]) + (X(i, t),)
                logp += np.log(tab[sel])
                logp += np.log(np.asarray(phi[i][t + 1])[X(i, t + 1)])
        for e in range(g.ne):
            i, j = g.src[e], g.dst[e]
            for t in range(L):
                logp += 0.5 * np.log(np.asarray(psi[e][t])[X(i, t), X(j, t)])
    mx = logp.max()
    logZ = mx + np.log(np.exp(logp - mx).sum())
    p = np.exp(logp - logZ)
    return p.reshape(dims), float(np.exp(logZ)), float(logZ)


def exact_marginals(bp, p):
    N, L = bp.g.N, bp.T + 1
    out = []
    for i in range(N):
        out.append([p.sum(axis=tuple(a for a in range(N * L) if a != i * L + t)) for t in range(L)])
    return out


def exact_pair_marginals(bp, p):
    g = bp.g
    N, L = g.N, bp.T + 1
    out = []
    for e in range(g.ne):
        i, j = g.src[e], g.dst[e]
        row = []
        for t in range(L):
            a, b = i * L + t, j * L + t
            m = p.sum(axis=tuple(c for c in range(N * L) if c not in (a, b)))
            row.append(m if a < b else m.T)
        out.append(row)
    return out


def exact_autocorrelations(bp, p, f=lambda x, i: x):
    """<f(x_i^t) f(x_i^u)> for t < u from the joint distribution (exact_autocorrelations of the reference's src/exact.jl,
    as used by test/glauber_small_tree.jl:46-50); zero elsewhere."""
    N, L = bp.g.N, bp.T + 1
    out = []
    for i in range(N):
        fx = np.array([f(x + 1, i) for x in range(bp.q[i])], dtype=float)
        r = np.zeros((L, L))
        for t in range(L):
            for u in range(t + 1, L):
                m = p.sum(axis=tuple(c for c in range(N * L) if c not in (i * L + t, i * L + u)))
                r[t, u] = fx @ m @ fx
        out.append(r)
    return out


def exact_alternate_marginals(bp, p):
    """p(x_i^t, x_j^{t+1}) for every directed edge i->j, t = 0..T-1 (exact_alternate_marginals of src/exact.jl)."""
    g = bp.g
    N, L = g.N, bp.T + 1
    out = []
    for e in range(g.ne):
        i, j = g.src[e], g.dst[e]
        row = []
        for t in range(L - 1):
            a, b = i * L + t, j * L + t + 1
            m = p.sum(axis=tuple(c for c in range(N * L) if c not in (a, b)))
            row.append(m if a < b else m.T)
        out.append(row)
    return out


def onesample(bp, rng):
    """forward sample of the prior dynamics: x_i^0 ~ phi_i^0 (normalised), x_i^{t+1} ~ w_i^t(. | x_neigh^t, x_i^t)
    (the prior part of /root/reference/src/sampling.jl onesample; reweightings at t > 0 are NOT applied).  1-based states."""
    g, w, q = bp.g, bp.w, bp.q
    N, L = g.N, bp.T + 1
    neigh = [[g.dst[e] for e in g.out_edges[i]] for i in range(N)]
    X = np.zeros((N, L), dtype=int)
    for i in range(N):
        p0 = np.asarray(bp.phi[i][0], dtype=float)
        X[i, 0] = 1 + rng.choice(q[i], p=p0 / p0.sum())
    for t in range(L - 1):
        for i in range(N):
            pr = np.array([w[i][t](xn, [X[k, t] for k in neigh[i]], X[i, t]) for xn in range(1, q[i] + 1)])
            X[i, t + 1] = 1 + rng.choice(q[i], p=pr / pr.sum())
    return X


def logprob(bp, X):
    """log of the un-normalised weight of trajectory X (N x (T+1), 1-based): /root/reference/src/mpbp.jl:301-323."""
    g, w, phi, psi = bp.g, bp.w, bp.phi, bp.psi
    N, L = g.N, bp.T + 1
    neigh = [[g.dst[e] for e in g.out_edges[i]] for i in range(N)]
    lp = 0.0
    for i in range(N):
        lp += np.log(phi[i][0][X[i, 0] - 1])
    for t in range(L - 1):
        for i in range(N):
            lp += np.log(w[i][t](X[i, t + 1], [X[k, t] for k in neigh[i]], X[i, t]))
            lp += np.log(phi[i][t + 1][X[i, t + 1] - 1])
    for t in range(L):
        for e in range(g.ne):
            lp += 0.5 * np.log(np.asarray(psi[e][t])[X[g.src[e], t] - 1, X[g.dst[e], t] - 1])
    return float(lp)
