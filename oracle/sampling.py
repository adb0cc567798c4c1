"""ORACLE (test infrastructure only).  Forward sampler of the prior dynamics, restating

* onesample!                        /root/reference/src/sampling.jl:30-59
* (w::RecursiveBPFactor)(x', xn, x)  /root/reference/src/recursive_bp_factor.jl:34-46  (neighbours folded one by one)

with the counter-based uniform numbers of the device sampler (csrc/kernels.cuh: samp_uniform, a splitmix64 finaliser of
(seed, node, time)) instead of Julia's global RNG, so that device and oracle trajectories can be compared bit for bit:
the reference's own streams are not reproducible outside Julia.  Plain Python floats, same operation order as the kernel.
"""
import math

import numpy as np

MASK = (1 << 64) - 1


def samp_uniform(seed, i, t):
    zz = (seed + 0x9E3779B97F4A7C15 * (i + 1) + 0xD1B54A32D192ED03 * (t + 2)) & MASK
    zz = ((zz ^ (zz >> 30)) * 0xBF58476D1CE4E5B9) & MASK
    zz = ((zz ^ (zz >> 27)) * 0x94D049BB133111EB) & MASK
    zz = zz ^ (zz >> 31)
    return float(zz >> 11) * (1.0 / 9007199254740992.0)


def _draw(p, u):
    S = 0.0
    for v in p:
        S = S + float(v)
    c = 0.0
    for x, v in enumerate(p):
        c = c + float(v) / S
        if u < c:
            return x
    return len(p) - 1


def transition(w, q, xn, x):
    """[P(x' | xn, x) for x' in 0..q-1] with 0-based states; folding order of the cavity recursion (prefix products, then
    the initial term), the same function as the factor's functor for a valid RecursiveBPFactor"""
    if not w.recursive:
        return [float(w(xp + 1, [v + 1 for v in xn], x + 1)) for xp in range(q)]
    z = len(xn)
    ny0 = w.nstates(0)
    init = [float(w.prob_y0(y + 1, x + 1)) for y in range(ny0)]
    if z == 0:
        P = init
    else:
        ny1 = w.nstates(1)
        P = [float(w.prob_xy(y + 1, xn[0] + 1, x + 1, 1)) for y in range(ny1)]
        for k in range(1, z):
            px = [float(w.prob_xy(y + 1, xn[k] + 1, x + 1, k + 1)) for y in range(ny1)]
            nyn = w.nstates(k + 1)
            Pn = []
            for y in range(nyn):
                acc = 0.0
                for y1 in range(len(P)):
                    for y2 in range(ny1):
                        acc = acc + (float(w.prob_yy(y + 1, y1 + 1, y2 + 1, x + 1, k, 1)) * P[y1]) * px[y2]
                Pn.append(acc)
            P = Pn
        nyz = w.nstates(z)
        Pn = []
        for y in range(nyz):
            acc = 0.0
            for y1 in range(len(P)):
                for y0 in range(ny0):
                    acc = acc + (float(w.prob_yy(y + 1, y1 + 1, y0 + 1, x + 1, z, 0)) * P[y1]) * init[y0]
            Pn.append(acc)
        P = Pn
    out = []
    for xp in range(q):
        acc = 0.0
        for y in range(len(P)):
            acc = acc + float(w.prob_y(xp + 1, x + 1, y + 1, z)) * P[y]
        out.append(acc)
    return out


def sample_prior(bp, seed):
    """X[i, t] (0-based states) and the likelihood weight exp(sum_{t>0} log phi + 1/2 sum_e log psi) of sampling.jl:52-58"""
    g = bp.g
    N, L = g.N, bp.T + 1
    X = np.zeros((N, L), dtype=np.int64)
    neigh = [[g.dst[e] for e in g.out_edges[i]] for i in range(N)]
    for i in range(N):
        X[i, 0] = _draw([float(v) for v in bp.phi[i][0]], samp_uniform(seed, i, -1))
    logl = 0.0
    for t in range(bp.T):
        for i in range(N):
            p = transition(bp.w[i][t], bp.q[i], [int(X[j, t]) for j in neigh[i]], int(X[i, t]))
            X[i, t + 1] = _draw(p, samp_uniform(seed, i, t))
            v = float(bp.phi[i][t + 1][X[i, t + 1]])
            logl += math.log(v) if v > 0 else -math.inf
    for e in range(g.ne):
        i, j = g.src[e], g.dst[e]
        for t in range(L):
            v = float(np.asarray(bp.psi[e][t])[X[i, t], X[j, t]])
            logl += 0.5 * math.log(v) if v > 0 else -math.inf
    return X, math.exp(logl)
