"""ORACLE (test infrastructure only).  Factor definitions restated from the reference:

* RecursiveBPFactor interface      /root/reference/src/recursive_bp_factor.jl:6-61
* HomogeneousGlauberFactor         /root/reference/src/Models/glauber/glauber_bp.jl:22-56
* PMJGlauberFactor                 /root/reference/src/Models/glauber/glauber_bp.jl:58-91
* IntegerGlauberFactor             /root/reference/src/Models/glauber/glauber_bp.jl:144-179
* GenericGlauberFactor             /root/reference/src/Models/glauber/glauber_bp.jl:1-20
* SISFactor                        /root/reference/src/Models/epidemics/sis_bp.jl:4-18,61-78
* SIRSFactor                       /root/reference/src/Models/epidemics/sirs_bp.jl:3-44
* DampedFactor                     /root/reference/src/recursive_bp_factor.jl:183-206
* GenericFactor (test wrapper)     /root/reference/src/test_factors.jl:41-45

All state / auxiliary-variable arguments are 1-based, exactly as in the Julia source, so that the
formulas can be compared line by line.
"""
from __future__ import annotations

import math

SUSCEPTIBLE, INFECTIOUS, RECOVERED = 1, 2, 3


def potts2spin(x):
    return 3 - 2 * x


class BPFactor:
    """generic factor: only the functor w(x_next, x_neighbours, x) is available (bp_core.jl:1-13)."""

    recursive = False

    def __call__(self, xnext, xneigh, x):
        raise NotImplementedError


class RecursiveBPFactor(BPFactor):
    recursive = True

    def nstates(self, l):
        raise NotImplementedError

    def prob_y(self, xnext, x, y, d):
        raise NotImplementedError

    def prob_xy(self, yk, xk, xi, k=None):
        raise NotImplementedError

    def prob_yy(self, y, y1, y2, xi, d1=None, d2=None):
        raise NotImplementedError

    def prob_y0(self, y, x):
        return float(y == 1)

    # recursive_bp_factor.jl:34-46
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

    # recursive_bp_factor.jl:49-54
    def prob_y_partial(self, xnext, x, xk, y1, d, k):
        return sum(
            self.prob_y(xnext, x, y, d + 1) * self.prob_xy(y2, xk, x, k) * self.prob_yy(y, y1, y2, x, d, 1)
            for y in range(1, self.nstates(d + 1) + 1)
            for y2 in range(1, self.nstates(1) + 1)
        )

    # recursive_bp_factor.jl:59-61
    def prob_y_dummy(self, xnext, x, xk, y1, d, j):
        return self.prob_y(xnext, x, y1, d)


class HomogeneousGlauberFactor(RecursiveBPFactor):
    def __init__(self, J, h, beta=1.0):
        self.bJ, self.bh = J * beta, h * beta

    def nstates(self, l):
        return l + 1

    def prob_y(self, xnext, x, z, d):
        y = 2 * z - 2 - d
        hji = self.bJ * y + self.bh
        E = -potts2spin(xnext) * hji
        return 1.0 / (1.0 + math.exp(2 * E))

    def prob_xy(self, yk, xk, xi, k=None):
        return float(yk != xk)

    def prob_yy(self, y, y1, y2, xi, d1=None, d2=None):
        return float(y == y1 + y2 - 1)

    def __call__(self, xnext, xneigh, x):
        hji = self.bJ * sum(potts2spin(v) for v in xneigh)
        E = -potts2spin(xnext) * (hji + self.bh)
        return 1.0 / (1.0 + math.exp(2 * E))


class PMJGlauberFactor(RecursiveBPFactor):
    def __init__(self, signs, J, h, beta=1.0):
        self.signs = [int(s) for s in signs]
        self.bJ, self.bh = J * beta, h * beta

    def nstates(self, d):
        return 2 * d + 1

    def prob_y(self, xnext, x, y, d):
        ht = y - d - 1
        E = -potts2spin(xnext) * (self.bJ * ht + self.bh)
        return 1.0 / (1.0 + math.exp(2 * E))

    def prob_xy(self, yk, xk, xi, k=None):
        return float(yk == potts2spin(xk) * self.signs[k - 1] + 2)

    def prob_yy(self, y, y1, y2, xi, d1=None, d2=None):
        return float(y == y1 + y2 - 1)

    def __call__(self, xnext, xneigh, x):
        hji = self.bJ * sum(s * potts2spin(v) for v, s in zip(xneigh, self.signs))
        E = -potts2spin(xnext) * (hji + self.bh)
        return 1.0 / (1.0 + math.exp(2 * E))


class IntegerGlauberFactor(RecursiveBPFactor):
    def __init__(self, J, h, beta):
        self.J = [int(j) for j in J]
        self.h, self.beta = h, beta
        self.K = sum(abs(j) for j in self.J) + 1

    def nstates(self, l):
        return 2 * self.K - 1

    def prob_y(self, xnext, x, y, d):
        ht = y - self.K
        E = -potts2spin(xnext) * self.beta * (ht + self.h)
        return 1.0 / (1.0 + math.exp(2 * E))

    def prob_xy(self, yk, xk, xi, k=None):
        return float(yk == potts2spin(xk) * self.J[k - 1] + self.K)

    def prob_yy(self, y, y1, y2, xi, d1=None, d2=None):
        return float(y + self.K == y1 + y2)

    def prob_y0(self, y, x):
        return float(y == self.K)

    def __call__(self, xnext, xneigh, x):
        ht = sum(j * potts2spin(v) for j, v in zip(self.J, xneigh))
        E = -potts2spin(xnext) * self.beta * (ht + self.h)
        return 1.0 / (1.0 + math.exp(2 * E))


class GenericGlauberFactor(BPFactor):
    def __init__(self, J, h, beta=1.0):
        self.bJ = [j * beta for j in J]
        self.bh = h * beta

    def __call__(self, xnext, xneigh, x):
        hji = sum(j * potts2spin(v) for v, j in zip(xneigh, self.bJ))
        E = -potts2spin(xnext) * (hji + self.bh)
        return 1.0 / (1.0 + math.exp(2 * E))


class SISFactor(RecursiveBPFactor):
    def __init__(self, lam, rho, alpha=0.0):
        self.lam, self.rho, self.alpha = lam, rho, alpha

    def nstates(self, l):
        return 1 if l == 0 else 2

    def prob_y(self, xnext, x, y, d):
        z = 1.0  # neighbour j susceptible (sis_bp.jl:63-64)
        w = (y == SUSCEPTIBLE) * (1 - self.alpha)
        if xnext == INFECTIOUS:
            return (x == INFECTIOUS) * (1 - self.rho) + (x == SUSCEPTIBLE) * (1 - z * w)
        return (x == INFECTIOUS) * self.rho + (x == SUSCEPTIBLE) * z * w

    def prob_xy(self, yk, xk, xi, k=None):
        lam = self.lam
        return (yk == INFECTIOUS) * lam * (xk == INFECTIOUS) + (yk == SUSCEPTIBLE) * (1 - lam * (xk == INFECTIOUS))

    def prob_yy(self, y, y1, y2, xi, d1=None, d2=None):
        return 1.0 * ((y == INFECTIOUS) == ((y1 == INFECTIOUS) or (y2 == INFECTIOUS)))

    def __call__(self, xnext, xneigh, x):
        if x == INFECTIOUS:
            return self.rho if xnext == SUSCEPTIBLE else 1 - self.rho
        p = (1 - self.alpha) * (1 - self.lam) ** sum(v == INFECTIOUS for v in xneigh)
        return p if xnext == SUSCEPTIBLE else 1 - p


class SIS_heterogeneousFactor(SISFactor):
    """per-neighbour infection probabilities lam[k] (src/Models/epidemics/sis_heterogeneous_bp.jl:4-15,52-72)."""

    def __init__(self, lam, rho, alpha=0.0):
        self.lam, self.rho, self.alpha = [float(v) for v in lam], rho, alpha

    def prob_xy(self, yk, xk, xi, k=None):
        lam = self.lam[k - 1]
        return (yk == INFECTIOUS) * lam * (xk == INFECTIOUS) + (yk == SUSCEPTIBLE) * (1 - lam * (xk == INFECTIOUS))

    def __call__(self, xnext, xneigh, x):
        if x == INFECTIOUS:
            return self.rho if xnext == SUSCEPTIBLE else 1 - self.rho
        p = 1 - self.alpha
        for v, lam in zip(xneigh, self.lam):
            p *= 1 - lam * (v == INFECTIOUS)
        return p if xnext == SUSCEPTIBLE else 1 - p


class SIRSFactor(RecursiveBPFactor):
    def __init__(self, lam, rho, sigma, alpha=0.0):
        self.lam, self.rho, self.sigma, self.alpha = lam, rho, sigma, alpha

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
        lam = self.lam
        return (yk == INFECTIOUS) * lam * (xk == INFECTIOUS) + (yk == SUSCEPTIBLE) * (1 - lam * (xk == INFECTIOUS))

    def prob_yy(self, y, y1, y2, xi, d1=None, d2=None):
        return 1.0 * ((y == INFECTIOUS) == ((y1 == INFECTIOUS) or (y2 == INFECTIOUS)))


class DampedFactor(RecursiveBPFactor):
    def __init__(self, w, p):
        assert 0 <= p <= 1
        self.w, self.p = w, p

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
    """forces the exhaustive-trace path (test_factors.jl:41-45)."""

    def __init__(self, w):
        self.w = w

    def __call__(self, xnext, xneigh, x):
        return self.w(xnext, xneigh, x)
