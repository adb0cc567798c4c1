"""Factor definitions of the host layer: the user-facing ``BPFactor`` / ``RecursiveBPFactor`` interface
and the model factors, mirroring the reference (names, argument meaning, 1-based states):

* interface           /root/reference/src/bp_core.jl:1-13, src/recursive_bp_factor.jl:6-61
* Glauber factors     /root/reference/src/Models/glauber/glauber_bp.jl:1-91,144-179
* SIS / SIRS factors  /root/reference/src/Models/epidemics/sis_bp.jl:4-78, sirs_bp.jl:3-44
* DampedFactor        /root/reference/src/recursive_bp_factor.jl:183-206

The factor definitions stay on the host; only their tabulated values cross the C-ABI
(``tabulate_class`` below builds the arrays documented at ``mpbp_add_node_class`` in include/mpbp.h).
"""
from __future__ import annotations

import math

import numpy as np

SUSCEPTIBLE, INFECTIOUS, RECOVERED = 1, 2, 3


def potts2spin(x):
    return 3 - 2 * x


class BPFactor:
    """A factor must implement ``w(x_next, x_neighbours, x)`` (bp_core.jl:1-13)."""

    def __call__(self, xnext, xneigh, x):
        raise NotImplementedError

    def key(self):
        """hashable identity used to share tabulated classes between nodes; None = never shared."""
        return None


class RecursiveBPFactor(BPFactor):
    """Minimal interface: nstates, prob_y, prob_xy, prob_yy (+ optional prob_y0, prob_y_partial)."""

    def nstates(self, l):
        raise NotImplementedError("nstates")

    def prob_y(self, xnext, x, y, d):
        raise NotImplementedError("prob_y")

    def prob_xy(self, yk, xk, xi, k=None):
        raise NotImplementedError("prob_xy")

    def prob_yy(self, y, y1, y2, xi, d1=None, d2=None):
        raise NotImplementedError("prob_yy")

    def prob_y0(self, y, x):
        return 1.0 if y == 1 else 0.0

    def prob_y_partial(self, xnext, x, xk, y1, d, k):
        n1 = self.nstates(1)
        return sum(
            self.prob_y(xnext, x, y, d + 1) * self.prob_xy(y2, xk, x, k) * self.prob_yy(y, y1, y2, x, d, 1)
            for y in range(1, self.nstates(d + 1) + 1)
            for y2 in range(1, n1 + 1)
        )

    def prob_y_dummy(self, xnext, x, xk, y1, d, j):
        return self.prob_y(xnext, x, y1, d)

    def __call__(self, xnext, xneigh, x):
        d = len(xneigh)
        P = [self.prob_y0(y, x) for y in range(1, self.nstates(0) + 1)]
        for k in range(1, d + 1):
            P = [
                sum(
                    self.prob_yy(y, y1, y2, x, 1, k - 1) * self.prob_xy(y1, xneigh[k - 1], x, k) * P[y2 - 1]
                    for y1 in range(1, self.nstates(1) + 1)
                    for y2 in range(1, len(P) + 1)
                )
                for y in range(1, self.nstates(k) + 1)
            ]
        return sum(P[y - 1] * self.prob_y(xnext, x, y, d) for y in range(1, len(P) + 1))


def _sigm(E):
    return 1.0 / (1.0 + math.exp(2 * E))


class HomogeneousGlauberFactor(RecursiveBPFactor):
    def __init__(self, J, h, beta=1.0):
        self.bJ, self.bh = float(J) * beta, float(h) * beta

    def key(self):
        return ("HG", self.bJ, self.bh)

    def nstates(self, l):
        return l + 1

    def prob_y(self, xnext, x, z, d):
        return _sigm(-potts2spin(xnext) * (self.bJ * (2 * z - 2 - d) + self.bh))

    def prob_xy(self, yk, xk, xi, k=None):
        return 1.0 if yk != xk else 0.0

    def prob_yy(self, y, y1, y2, xi, d1=None, d2=None):
        return 1.0 if y == y1 + y2 - 1 else 0.0

    def __call__(self, xnext, xneigh, x):
        return _sigm(-potts2spin(xnext) * (self.bJ * sum(potts2spin(v) for v in xneigh) + self.bh))


class PMJGlauberFactor(RecursiveBPFactor):
    def __init__(self, signs, J, h, beta=1.0):
        self.signs = tuple(int(s) for s in signs)
        self.bJ, self.bh = float(J) * beta, float(h) * beta

    def key(self):
        return ("PMJ", self.signs, self.bJ, self.bh)

    def nstates(self, d):
        return 2 * d + 1

    def prob_y(self, xnext, x, y, d):
        return _sigm(-potts2spin(xnext) * (self.bJ * (y - d - 1) + self.bh))

    def prob_xy(self, yk, xk, xi, k=None):
        return 1.0 if yk == potts2spin(xk) * self.signs[k - 1] + 2 else 0.0

    def prob_yy(self, y, y1, y2, xi, d1=None, d2=None):
        return 1.0 if y == y1 + y2 - 1 else 0.0

    def __call__(self, xnext, xneigh, x):
        hji = self.bJ * sum(s * potts2spin(v) for v, s in zip(xneigh, self.signs))
        return _sigm(-potts2spin(xnext) * (hji + self.bh))


class IntegerGlauberFactor(RecursiveBPFactor):
    def __init__(self, J, h, beta):
        self.J = tuple(int(j) for j in J)
        self.h, self.beta = float(h), float(beta)
        self.K = sum(abs(j) for j in self.J) + 1

    def key(self):
        return ("IG", self.J, self.h, self.beta)

    def nstates(self, l):
        return 2 * self.K - 1

    def prob_y(self, xnext, x, y, d):
        return _sigm(-potts2spin(xnext) * self.beta * ((y - self.K) + self.h))

    def prob_xy(self, yk, xk, xi, k=None):
        return 1.0 if yk == potts2spin(xk) * self.J[k - 1] + self.K else 0.0

    def prob_yy(self, y, y1, y2, xi, d1=None, d2=None):
        return 1.0 if y + self.K == y1 + y2 else 0.0

    def prob_y0(self, y, x):
        return 1.0 if y == self.K else 0.0

    def __call__(self, xnext, xneigh, x):
        ht = sum(j * potts2spin(v) for j, v in zip(self.J, xneigh))
        return _sigm(-potts2spin(xnext) * self.beta * (ht + self.h))


class GenericGlauberFactor(BPFactor):
    def __init__(self, J, h, beta=1.0):
        self.bJ = tuple(float(j) * beta for j in J)
        self.bh = float(h) * beta

    def key(self):
        return ("GG", self.bJ, self.bh)

    def __call__(self, xnext, xneigh, x):
        hji = sum(j * potts2spin(v) for v, j in zip(xneigh, self.bJ))
        return _sigm(-potts2spin(xnext) * (hji + self.bh))


class SISFactor(RecursiveBPFactor):
    def __init__(self, lam, rho, alpha=0.0):
        for v in (lam, rho, alpha):
            assert 0 <= v <= 1
        self.lam, self.rho, self.alpha = float(lam), float(rho), float(alpha)

    def key(self):
        return ("SIS", self.lam, self.rho, self.alpha)

    def nstates(self, l):
        return 1 if l == 0 else 2

    def prob_y(self, xnext, x, y, d):
        w = (y == SUSCEPTIBLE) * (1 - self.alpha)
        if xnext == INFECTIOUS:
            return (x == INFECTIOUS) * (1 - self.rho) + (x == SUSCEPTIBLE) * (1 - w)
        return (x == INFECTIOUS) * self.rho + (x == SUSCEPTIBLE) * w

    def prob_xy(self, yk, xk, xi, k=None):
        inf = xk == INFECTIOUS
        return (yk == INFECTIOUS) * self.lam * inf + (yk == SUSCEPTIBLE) * (1 - self.lam * inf)

    def prob_yy(self, y, y1, y2, xi, d1=None, d2=None):
        return 1.0 if (y == INFECTIOUS) == ((y1 == INFECTIOUS) or (y2 == INFECTIOUS)) else 0.0

    def __call__(self, xnext, xneigh, x):
        if x == INFECTIOUS:
            return self.rho if xnext == SUSCEPTIBLE else 1 - self.rho
        p = (1 - self.alpha) * (1 - self.lam) ** sum(v == INFECTIOUS for v in xneigh)
        return p if xnext == SUSCEPTIBLE else 1 - p


class SIS_heterogeneousFactor(SISFactor):
    """SIS with per-neighbour incoming infection probabilities lam[k], k = position of the neighbour in the node's
    (ascending) neighbour list (src/Models/epidemics/sis_heterogeneous_bp.jl:4-15, prob_xy :69-72)."""

    def __init__(self, lam, rho, alpha=0.0):
        lam = [float(v) for v in lam]
        for v in lam + [rho, alpha]:
            assert 0 <= v <= 1
        self.lam, self.rho, self.alpha = lam, float(rho), float(alpha)

    def key(self):
        return ("SIShet", tuple(self.lam), self.rho, self.alpha)

    def prob_xy(self, yk, xk, xi, k=None):
        lam, inf = self.lam[k - 1], xk == INFECTIOUS
        return (yk == INFECTIOUS) * lam * inf + (yk == SUSCEPTIBLE) * (1 - lam * inf)

    def __call__(self, xnext, xneigh, x):
        if x == INFECTIOUS:
            return self.rho if xnext == SUSCEPTIBLE else 1 - self.rho
        p = 1 - self.alpha
        for v, lam in zip(xneigh, self.lam):
            p *= 1 - lam * (v == INFECTIOUS)
        return p if xnext == SUSCEPTIBLE else 1 - p


class SIRSFactor(RecursiveBPFactor):
    def __init__(self, lam, rho, sigma, alpha=0.0):
        for v in (lam, rho, sigma, alpha):
            assert 0 <= v <= 1
        self.lam, self.rho, self.sigma, self.alpha = float(lam), float(rho), float(sigma), float(alpha)

    def key(self):
        return ("SIRS", self.lam, self.rho, self.sigma, self.alpha)

    def nstates(self, l):
        return 1 if l == 0 else 2

    def prob_y(self, xnext, x, y, d):
        w = (y == SUSCEPTIBLE) * (1 - self.alpha)
        if xnext == INFECTIOUS:
            return (x == INFECTIOUS) * (1 - self.rho) + (x == SUSCEPTIBLE) * (1 - w)
        if xnext == SUSCEPTIBLE:
            return (x == RECOVERED) * self.sigma + (x == SUSCEPTIBLE) * w
        return (x == INFECTIOUS) * self.rho + (x == RECOVERED) * (1 - self.sigma)

    def prob_xy(self, yk, xk, xi, k=None):
        inf = xk == INFECTIOUS
        return (yk == INFECTIOUS) * self.lam * inf + (yk == SUSCEPTIBLE) * (1 - self.lam * inf)

    def prob_yy(self, y, y1, y2, xi, d1=None, d2=None):
        return 1.0 if (y == INFECTIOUS) == ((y1 == INFECTIOUS) or (y2 == INFECTIOUS)) else 0.0


class DampedFactor(RecursiveBPFactor):
    """adds a transition x->x with probability p and rescales the others by 1-p."""

    def __init__(self, w, p):
        assert 0 <= p <= 1
        self.w, self.p = w, float(p)

    def key(self):
        k = self.w.key()
        return None if k is None else ("Damped", k, self.p)

    def nstates(self, l):
        return self.w.nstates(l)

    def prob_xy(self, *a):
        return self.w.prob_xy(*a)

    def prob_yy(self, *a):
        return self.w.prob_yy(*a)

    def prob_y0(self, y, x):
        return self.w.prob_y0(y, x)

    def prob_y(self, xnext, x, y, d):
        return (1 - self.p) * self.w.prob_y(xnext, x, y, d) + self.p * (xnext == x)

    def __call__(self, xnext, xneigh, x):
        return (1 - self.p) * self.w(xnext, xneigh, x) + self.p * (xnext == x)


class GenericFactor(BPFactor):
    """wraps any factor so that only its functor is used: forces the exhaustive-trace path (src/test_factors.jl:41-45)"""

    def __init__(self, w):
        self.w = w

    def key(self):
        k = self.w.key()
        return None if k is None else ("Generic", k)

    def __call__(self, xnext, xneigh, x):
        return self.w(xnext, xneigh, x)


# --------------------------------------------------------------------------------------
# tabulation for the C-ABI (mpbp_add_node_class)
# --------------------------------------------------------------------------------------
def cavity_pairs(z):
    """(d1,d2) operand sizes met by the cavity recursion of a degree-z node."""
    pairs = []

    def add(p):
        if p not in pairs:
            pairs.append(p)

    if z == 1:
        add((1, 0))
        return pairs
    for k in range(1, z):
        add((k, 1))
    add((z, 0))
    for k in range(z - 1, 0, -1):
        add((1, z - 1 - k))
    for k in range(1, z):
        add((k, z - 1 - k))
    return pairs


def _f(a):
    return np.asarray(a, dtype=np.float64).ravel(order="F")


def tabulate_class(ws, z, q, qn):
    """ws: list of RecursiveBPFactor over time (length nt = 1 or T+1).  Returns the argument arrays of
    mpbp_add_node_class: ny, pxy, pairs, pyy, w, wd, minit (all flattened column-major)."""
    w0 = ws[0]
    ny = np.array([w0.nstates(l) for l in range(z + 1)], dtype=np.int32)
    pairs = cavity_pairs(z)
    pxy, wj, wd, minit = [], [], [], []
    for w in ws:
        for k in range(z):
            a = np.zeros((ny[1], qn[k], q))
            for y in range(ny[1]):
                for xk in range(qn[k]):
                    for xi in range(q):
                        a[y, xk, xi] = w.prob_xy(y + 1, xk + 1, xi + 1, k + 1)
            pxy.append(_f(a))
        for j in range(z):
            a = np.zeros((q, q, qn[j], ny[z - 1]))
            for xn in range(q):
                for x in range(q):
                    for xj in range(qn[j]):
                        for y in range(ny[z - 1]):
                            a[xn, x, xj, y] = w.prob_y_partial(xn + 1, x + 1, xj + 1, y + 1, z - 1, j + 1)
            wj.append(_f(a))
        a = np.zeros((q, q, ny[z]))
        for xn in range(q):
            for x in range(q):
                for y in range(ny[z]):
                    a[xn, x, y] = w.prob_y(xn + 1, x + 1, y + 1, z)
        wd.append(_f(a))
        a = np.zeros((ny[0], q))
        for y in range(ny[0]):
            for x in range(q):
                a[y, x] = w.prob_y0(y + 1, x + 1)
        minit.append(_f(a))
    pyy = []
    for (d1, d2) in pairs:
        for w in ws:
            a = np.zeros((ny[d1 + d2], ny[d1], ny[d2], q))
            for y in range(ny[d1 + d2]):
                for y1 in range(ny[d1]):
                    for y2 in range(ny[d2]):
                        for x in range(q):
                            a[y, y1, y2, x] = w.prob_yy(y + 1, y1 + 1, y2 + 1, x + 1, d1, d2)
            pyy.append(_f(a))
    cat = lambda l: np.ascontiguousarray(np.concatenate(l)) if l else np.zeros(1)
    return dict(ny=ny, pxy=cat(pxy), pairs=pairs, pyy=cat(pyy), w=cat(wj), wd=cat(wd), minit=cat(minit))
