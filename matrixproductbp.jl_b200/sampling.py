"""Observation drawing and a host-side forward sampler (reference: src/sampling.jl:30-66,191-210).

The API's `onesample` / `draw_node_observations_` draw the trajectory ON THE DEVICE (mpbp_sample_prior); the numpy `onesample`
below evaluates the factors' functors on the host and only serves the CPU-only checks of the factor definitions.
`onesample` draws a trajectory from the prior dynamics and returns it with its likelihood weight, exactly like the
reference's `onesample!`; `draw_node_observations_` turns entries of a trajectory into (soft) one-hot reweightings.
States are numbered from 1, as in the reference."""
from __future__ import annotations

import math

import numpy as np


def onesample(g, w, q, T, phi, psi=None, rng=None):
    """X[i, t] (N x (T+1), 1-based) ~ prior: x_i^0 ~ phi_i^0 / sum, x_i^{t+1} ~ w_i^t(. | x_neigh^t, x_i^t); second return
    value: exp(sum_t>0 log phi_i^t[x_i^t] + 1/2 sum_e,t log psi_e^t[x_i^t, x_j^t])  (src/sampling.jl:30-59)."""
    rng = np.random.default_rng(rng)
    N, L = g.N, T + 1
    X = np.zeros((N, L), dtype=np.int64)
    neigh = [np.asarray(g.neighbors(i), dtype=np.int64) for i in range(N)]
    for i in range(N):
        p0 = np.asarray(phi[i][0], dtype=float)
        X[i, 0] = 1 + rng.choice(int(q[i]), p=p0 / p0.sum())
    logl = 0.0
    for t in range(T):
        for i in range(N):
            xn = [int(v) for v in X[neigh[i], t]]
            p = np.array([w[i][t](xx, xn, int(X[i, t])) for xx in range(1, int(q[i]) + 1)], dtype=float)
            X[i, t + 1] = 1 + rng.choice(int(q[i]), p=p / p.sum())
            v = float(phi[i][t + 1][X[i, t + 1] - 1])
            logl += math.log(v) if v > 0 else -math.inf
    if psi is not None:
        for e in range(g.ne):
            i, j = int(g.src[e]), int(g.dst[e])
            for t in range(L):
                v = float(np.asarray(psi[e][t])[X[i, t] - 1, X[j, t] - 1])
                logl += 0.5 * math.log(v) if v > 0 else -math.inf
    return X, math.exp(logl)


def draw_node_observations_(phi, X, nobs, softinf=math.inf, last_time=False, times=None, rng=None):
    """multiply phi[i][t] by a (soft) indicator of X[i, t] for `nobs` distinct (i, t) pairs drawn without replacement
    from all nodes x `times` (default: every time, or only the last one with last_time) -- src/sampling.jl:191-203.
    softinf = inf gives hard observations (weights 1 / 0).  Returns (phi, observed) with observed sorted."""
    rng = np.random.default_rng(rng)
    N, L = X.shape
    times = list(times) if times is not None else ([L - 1] if last_time else list(range(L)))
    pairs = [(i, t) for i in range(N) for t in times]
    assert 0 <= nobs <= len(pairs)
    observed = sorted(pairs[k] for k in rng.choice(len(pairs), size=nobs, replace=False))
    if math.isinf(softinf):
        softone, softzero = 1.0, 0.0
    else:
        softone = 1.0 / (1.0 + math.exp(-math.log(softinf)))
        softzero = 1.0 / (1.0 + math.exp(math.log(softinf)))
    for (i, t) in observed:
        ph = np.asarray(phi[i][t], dtype=float)
        phi[i][t] = ph * np.where(np.arange(1, len(ph) + 1) == X[i, t], softone, softzero)
    return phi, observed
