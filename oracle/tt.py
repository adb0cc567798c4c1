"""ORACLE (test infrastructure only -- never imported by the product path).

CPU/numpy restatement of the tensor-train arithmetic that MatrixProductBP.jl delegates to the
un-vendored dependency TensorTrains.jl (compat "0.12", /root/reference/Project.toml:49) and to
LinearAlgebra.svd (LAPACK gesdd).  The semantics restated here are the ones listed in
SURVEY.md section 8c; they are pinned end-to-end by tests/test_oracle_golden.py against the only
numeric literals in the reference's test-suite (/root/reference/test/sis_infinite_graph.jl:21-29)
and against brute-force enumeration (oracle/exact.py).

Conventions
-----------
* A tensor train is a list of numpy arrays ``A[t][m, n, x...]`` (same index order as the Julia
  ``Array{F,3/4}`` of /root/reference/src/mpems.jl:1-17) plus a log-scale ``ls``:
  ``value(x) = exp(ls) * prod_t A[t][:, :, x_t]``  (the reference stores ``z`` with
  ``value = prod / z``, i.e. ``ls = -log z``).
* Fused indices ``(a, b)`` follow Julia/TensorCast: ``a`` fastest == Fortran-order reshape.
"""
from __future__ import annotations

import numpy as np


# --------------------------------------------------------------------------------------
# SVD truncators  (TensorTrains.jl: TruncThresh / TruncBond / TruncBondMax / TruncBondThresh;
# re-exported at /root/reference/src/MatrixProductBP.jl:42,69)
# --------------------------------------------------------------------------------------
class SVDTrunc:
    def keep(self, lam: np.ndarray) -> int:
        raise NotImplementedError

    def __call__(self, M: np.ndarray):
        U, lam, Vt = np.linalg.svd(M, full_matrices=False)
        k = max(1, int(self.keep(lam)))
        return U[:, :k], lam[:k], Vt[:k, :]


class TruncThresh(SVDTrunc):
    """keep lam_k > eps * ||lam||_2 (last index satisfying it)."""

    def __init__(self, eps: float):
        self.eps = float(eps)

    def keep(self, lam):
        nrm = np.linalg.norm(lam)
        idx = np.nonzero(lam > self.eps * nrm)[0]
        return (idx[-1] + 1) if len(idx) else 1

    def __repr__(self):
        return f"TruncThresh({self.eps})"


class TruncBond(SVDTrunc):
    """keep the first min(length(lam), d)."""

    def __init__(self, d: int):
        self.d = int(d)

    def keep(self, lam):
        return min(len(lam), self.d)

    def __repr__(self):
        return f"TruncBond({self.d})"


class TruncBondMax(TruncBond):
    """TruncBond that also records the max truncation error (error not used by the hot path)."""

    def __init__(self, d: int):
        super().__init__(d)
        self.maxerr = 0.0

    def keep(self, lam):
        k = min(len(lam), self.d)
        if k < len(lam) and lam[0] > 0:
            self.maxerr = max(self.maxerr, float(np.sqrt(np.sum(lam[k:] ** 2) / np.sum(lam ** 2))))
        return k


class TruncBondThresh(SVDTrunc):
    def __init__(self, d: int, eps: float = 0.0):
        self.d = int(d)
        self.eps = float(eps)

    def keep(self, lam):
        return min(TruncThresh(self.eps).keep(lam), self.d)

    def __repr__(self):
        return f"TruncBondThresh({self.d},{self.eps})"


# --------------------------------------------------------------------------------------
# Tensor train container
# --------------------------------------------------------------------------------------
class TT:
    __slots__ = ("tensors", "ls")

    def __init__(self, tensors, ls: float = 0.0):
        self.tensors = [np.asarray(a, dtype=np.float64) for a in tensors]
        self.ls = float(ls)

    def __len__(self):
        return len(self.tensors)

    def __getitem__(self, t):
        return self.tensors[t]

    def __setitem__(self, t, v):
        self.tensors[t] = v

    def __iter__(self):
        return iter(self.tensors)

    def copy(self):
        return TT([a.copy() for a in self.tensors], self.ls)

    def bond_dims(self):
        return [a.shape[0] for a in self.tensors] + [self.tensors[-1].shape[1]]

    def evaluate(self, x):
        """x: sequence over sites of index tuples (0-based)."""
        M = np.ones((1, 1))
        for a, xt in zip(self.tensors, x):
            M = M @ a[(slice(None), slice(None)) + tuple(np.atleast_1d(xt))]
        return float(np.trace(M) * np.exp(self.ls))


def flat_tt(bondsizes, *q):
    """TensorTrains.flat_tt: constant tensors, overall normalisation 1
    (used by flat_mpem1/flat_mpem2, /root/reference/src/mpems.jl:6,20)."""
    L = len(bondsizes) - 1
    A = TT([np.ones((bondsizes[t], bondsizes[t + 1]) + tuple(q)) for t in range(L)])
    normalize(A)
    return A


def rand_tt(bondsizes, *q, rng=None):
    rng = np.random.default_rng(rng)
    L = len(bondsizes) - 1
    A = TT([rng.random((bondsizes[t], bondsizes[t + 1]) + tuple(q)) for t in range(L)])
    normalize(A)
    return A


def _rescale(M, A: TT):
    """max-abs rescale with the guards of /root/reference/src/mpems.jl:76-80; factor folded into ls."""
    m = np.max(np.abs(M)) if M.size else 0.0
    if np.isfinite(m) and m != 0.0:
        M = M / m
        A.ls += np.log(m)
    return M


def orthogonalize_right(A: TT, trunc: SVDTrunc):
    """TensorTrains.orthogonalize_right!: R->L sweep of (truncated) SVDs; A[t] <- V' ;
    carry A[t-1] * U * diag(lam)."""
    L = len(A)
    C = A[L - 1]
    for t in range(L - 1, 0, -1):
        phys = C.shape[2:]
        M = C.reshape(C.shape[0], -1, order="F")  # M[m,(n,x)]
        U, lam, Vt = trunc(M)
        A[t] = Vt.reshape((Vt.shape[0], C.shape[1]) + phys, order="F")
        UL = U * lam[None, :]
        C = np.einsum("mk...,kr->mr...", A[t - 1], UL)
        C = _rescale(C, A)
    A[0] = C
    return A


def orthogonalize_left(A: TT, trunc: SVDTrunc):
    """TensorTrains.orthogonalize_left!: L->R sweep; A[t] <- U ; carry diag(lam) V' * A[t+1]."""
    L = len(A)
    C = A[0]
    for t in range(0, L - 1):
        m, n = C.shape[:2]
        phys = C.shape[2:]
        P = int(np.prod(phys)) if len(phys) else 1
        # M[(m,x),n]
        M = np.moveaxis(C.reshape(m, n, P, order="F"), 1, 2).reshape(m * P, n, order="F")
        U, lam, Vt = trunc(M)
        r = len(lam)
        A[t] = np.moveaxis(U.reshape(m, P, r, order="F"), 1, 2).reshape((m, r) + phys, order="F")
        LV = lam[:, None] * Vt
        C = np.einsum("rl,ln...->rn...", LV, A[t + 1])
        C = _rescale(C, A)
    A[L - 1] = C
    return A


def compress(A: TT, trunc: SVDTrunc, is_orthogonal: str = "none"):
    """TensorTrains.compress!(A; svd_trunc, is_orthogonal)."""
    if is_orthogonal == "none":
        orthogonalize_right(A, TruncThresh(0.0))
        orthogonalize_left(A, trunc)
    elif is_orthogonal == "left":
        orthogonalize_right(A, trunc)
    elif is_orthogonal == "right":
        orthogonalize_left(A, trunc)
    else:
        raise ValueError(is_orthogonal)
    return A


def normalize_eachmatrix(A: TT):
    for t in range(len(A)):
        m = np.max(np.abs(A[t]))
        if np.isfinite(m) and m != 0.0:
            A[t] = A[t] / m
            A.ls += np.log(m)
    return A


def _site_sum(a):
    return a.reshape(a.shape[0], a.shape[1], -1).sum(axis=2)


def lognormalization(A: TT) -> float:
    """log of TensorTrains.normalization(A) = sum_x prod_t A[t](x_t) / z (assumed positive)."""
    l = np.ones((1,))
    acc = A.ls
    first = True
    for a in A:
        S = _site_sum(a)
        l = S if first else l @ S
        if first:
            l = S
            first = False
        m = np.max(np.abs(l))
        if np.isfinite(m) and m != 0.0:
            l = l / m
            acc += np.log(m)
    val = np.trace(np.atleast_2d(l))
    return acc + np.log(np.abs(val))


def normalize(A: TT) -> float:
    """TensorTrains.normalize!: rescale so that normalization == 1 and z == 1;
    returns log of the previous normalization."""
    lz = lognormalization(A)
    # fold everything into the tensors: value(x) = prod_t A'[t]
    L = len(A)
    shift = (A.ls - lz) / L
    f = np.exp(shift)
    for t in range(L):
        A[t] = A[t] * f
    A.ls = 0.0
    return lz


def accumulate_L(A: TT):
    """left environments l[t] = prod_{s<=t} sum_x A[s][:,:,x]  (each rescaled; scale irrelevant for marginals)."""
    Ls = []
    l = np.ones((1, A[0].shape[0]))
    for a in A:
        l = l @ _site_sum(a)
        m = np.max(np.abs(l))
        if np.isfinite(m) and m != 0.0:
            l = l / m
        Ls.append(l)
    return Ls


def accumulate_R(A: TT):
    Rs = [None] * len(A)
    r = np.ones((A[-1].shape[1], 1))
    for t in range(len(A) - 1, -1, -1):
        r = _site_sum(A[t]) @ r
        m = np.max(np.abs(r))
        if np.isfinite(m) and m != 0.0:
            r = r / m
        Rs[t] = r
    return Rs


def marginals(A: TT):
    """TensorTrains.marginals: p_t[x] ~ tr(L_{t-1} A[t][:,:,x] R_{t+1}), normalised to sum 1."""
    Ls = accumulate_L(A)
    Rs = accumulate_R(A)
    L = len(A)
    out = []
    for t in range(L):
        l = Ls[t - 1] if t > 0 else np.ones((1, A[0].shape[0]))
        r = Rs[t + 1] if t < L - 1 else np.ones((A[-1].shape[1], 1))
        a = A[t]
        p = np.einsum("im,mn...,nj->...", l, a, r)
        out.append(p / p.sum())
    return out


def twovar_marginals(A: TT):
    """TensorTrains.twovar_marginals (used by autocorrelations, /root/reference/src/mpbp.jl:239-255):
    p[t][u][x_t, x_u] for t<u (flattened physical index per site)."""
    L = len(A)
    Ls = accumulate_L(A)
    Rs = accumulate_R(A)
    res = [[None] * L for _ in range(L)]
    for t in range(L - 1):
        l = Ls[t - 1] if t > 0 else np.ones((1, A[0].shape[0]))
        at = A[t].reshape(A[t].shape[0], A[t].shape[1], -1)
        M = np.einsum("im,mnx->xn", l, at)  # [x_t, n]
        for u in range(t + 1, L):
            au = A[u].reshape(A[u].shape[0], A[u].shape[1], -1)
            r = Rs[u + 1] if u < L - 1 else np.ones((A[-1].shape[1], 1))
            p = np.einsum("xn,nky,kj->xy", M, au, r)
            res[t][u] = p / p.sum()
            M = M @ _site_sum(A[u])
            m = np.max(np.abs(M))
            if m > 0:
                M = M / m
    return res


def tt_sum(A: TT, B: TT, coefB: float = 1.0) -> TT:
    """block-diagonal TT for A + coefB * B (TensorTrains._compose as used by set_msg!,
    /root/reference/src/recursive_bp_factor.jl:172-176; both operands must have ls folded in)."""
    assert len(A) == len(B)
    L = len(A)
    fa, fb = np.exp(A.ls / L), np.exp(B.ls / L)
    out = []
    for t in range(L):
        a, b = A[t] * fa, B[t] * fb
        if t == 0:
            b = b * coefB
        phys = a.shape[2:]
        ml = a.shape[0] + b.shape[0] if t > 0 else 1
        nr = a.shape[1] + b.shape[1] if t < L - 1 else 1
        c = np.zeros((ml, nr) + phys)
        if t == 0 and L == 1:
            c = a + b
        elif t == 0:
            c[:, : a.shape[1]] = a
            c[:, a.shape[1]:] = b
        elif t == L - 1:
            c[: a.shape[0], :] = a
            c[a.shape[0]:, :] = b
        else:
            c[: a.shape[0], : a.shape[1]] = a
            c[a.shape[0]:, a.shape[1]:] = b
        out.append(c)
    return TT(out, 0.0)
