"""Host-side mirror of the MatrixProductBP.jl API for the message-update hot path, with a CUDA-backed MPBP
state driven through the C-ABI of include/mpbp.h (libmpbp_b200.so).  There is no CPU path.

reference entry points mirrored here (Julia `!` suffix -> trailing underscore):
  mpbp(g, w, q, T; phi, psi)            /root/reference/src/mpbp.jl:60-70
  mpbp(::Glauber) / mpbp(::SIS) / mpbp(::SIRS)   src/Models/glauber/glauber_bp.jl:94-100, epidemics/sis_bp.jl:42-46, sirs_bp.jl:21-25
  mpbp_infinite_graph(k, w, q, phi)     src/infinite_graph.jl:22-35
  iterate_(bp; maxiter, svd_trunc, tol, damp, nodes, shuffle_nodes)  src/mpbp.jl:185-198  -> (iters, cb)
  beliefs(bp), pair_beliefs(bp), bethe_free_energy(bp), means(f, bp), reset_messages_(bp)   src/mpbp.jl:72-80,202-261,298
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .factors import (BPFactor, DampedFactor, GenericFactor, GenericGlauberFactor, HomogeneousGlauberFactor, IntegerGlauberFactor,
                      PMJGlauberFactor, RecursiveBPFactor, SIRSFactor, SIS_heterogeneousFactor, SISFactor, tabulate_class)
from .truncations import SVDTrunc, TruncBond, TruncBondMax, TruncBondThresh, TruncThresh

__all__ = [
    "BPFactor", "RecursiveBPFactor", "HomogeneousGlauberFactor", "PMJGlauberFactor", "IntegerGlauberFactor",
    "GenericGlauberFactor", "SISFactor", "SIS_heterogeneousFactor", "SIS_heterogeneous", "SIRSFactor", "DampedFactor", "TruncBond", "TruncBondMax", "TruncThresh",
    "TruncBondThresh", "GenericFactor", "IndexedBiDiGraph", "InfiniteRegularGraph", "InfiniteBipartiteRegularGraph", "mpbp_infinite_bipartite_graph", "Ising", "Glauber", "SIS", "SIRS", "MPBP", "CB_BP",
    "mpbp", "mpbp_infinite_graph", "periodic_mpbp", "periodic_mpbp_infinite_graph", "iterate_", "beliefs", "beliefs_tu", "autocorrelations", "autocovariances", "alternate_marginals", "alternate_correlations", "pair_correlations", "pair_beliefs", "bethe_free_energy", "means",
    "reset_messages_", "glauber_factors", "MPBPError", "onesample", "sample_prior", "draw_node_observations_",
]

MPBPError = _lib.MPBPError


def _p(a, t):
    return a.ctypes.data_as(t)


# --------------------------------------------------------------------------------------
# graphs
# --------------------------------------------------------------------------------------
class IndexedBiDiGraph:
    """Directed edge index = position in the CSC of the symmetric adjacency matrix (source = column), i.e.
    edges sorted by (src, dst) -- the order of IndexedGraphs.IndexedBiDiGraph used at src/mpbp.jl:40-58."""

    def __init__(self, n, undirected_edges):
        und = np.asarray(sorted({(min(int(a), int(b)), max(int(a), int(b))) for a, b in undirected_edges if a != b}), dtype=np.int64).reshape(-1, 2)
        self.N = int(n)
        self.undirected = und
        s = np.concatenate([und[:, 0], und[:, 1]])
        d = np.concatenate([und[:, 1], und[:, 0]])
        order = np.lexsort((d, s))
        self.src, self.dst = np.ascontiguousarray(s[order]), np.ascontiguousarray(d[order])
        E2 = len(self.src)
        key = self.src * self.N + self.dst
        rkey = self.dst * self.N + self.src
        self.rev = np.ascontiguousarray(np.searchsorted(key, rkey).astype(np.int64))
        self.colptr = np.zeros(self.N + 1, dtype=np.int64)
        np.add.at(self.colptr, self.src + 1, 1)
        self.colptr = np.cumsum(self.colptr)
        assert E2 == 0 or np.all(key[self.rev] == rkey)
        # index of the undirected edge (i<j, sorted) of every directed edge: J / psi of the models are given per undirected edge
        ukey = und[:, 0] * self.N + und[:, 1]
        self.und_of = np.searchsorted(ukey, np.minimum(self.src, self.dst) * self.N + np.maximum(self.src, self.dst))

    @classmethod
    def from_adjacency(cls, A):
        A = np.asarray(A)
        i, j = np.nonzero(np.triu(A != 0, 1) | np.triu((A != 0).T, 1))
        return cls(A.shape[0], list(zip(i.tolist(), j.tolist())))

    @classmethod
    def from_networkx(cls, G):
        return cls(G.number_of_nodes(), list(G.edges()))

    @property
    def ne(self):
        return len(self.src)

    def degree(self, i):
        return int(self.colptr[i + 1] - self.colptr[i])

    def outedges(self, i):
        return range(int(self.colptr[i]), int(self.colptr[i + 1]))

    def neighbors(self, i):
        return self.dst[self.colptr[i]:self.colptr[i + 1]]


class InfiniteRegularGraph:
    """src/infinite_graph.jl:8-20"""

    def __init__(self, k):
        self.k = int(k)
        self.N = 1

    @property
    def ne(self):
        return 1

    def degree(self, i):
        return self.k


class InfiniteBipartiteRegularGraph:
    """src/infinite_graph.jl:62-85: two node classes of degrees k = (kA, kB); message slot e (0-based) holds the message INTO
    node e (from the other class), exactly like the reference's edge indices."""

    def __init__(self, k):
        self.k = (int(k[0]), int(k[1]))
        self.N = 2

    @property
    def ne(self):
        return 2

    def degree(self, i):
        return self.k[i]

    def neighbors(self, i):
        return [1 - i] * self.k[i]


# --------------------------------------------------------------------------------------
# models (src/Models/glauber/glauber.jl, epidemics/sis.jl, sirs.jl)
# --------------------------------------------------------------------------------------
class Ising:
    def __init__(self, g: IndexedBiDiGraph, J=None, h=None, beta=1.0):
        self.g = g
        nund = len(g.undirected)
        self.J = np.ones(nund) if J is None else np.asarray(J, dtype=float)
        self.h = np.zeros(g.N) if h is None else np.asarray(h, dtype=float)
        self.beta = float(beta)
        assert len(self.J) == nund and len(self.h) == g.N


class Glauber:
    def __init__(self, ising: Ising, T: int, phi=None, psi=None):
        self.ising, self.T = ising, int(T)
        N = ising.g.N
        self.phi = [[np.ones(2) for _ in range(T + 1)] for _ in range(N)] if phi is None else phi
        self.psi = psi  # per undirected edge [t][2x2] or None


class SIS:
    def __init__(self, g: IndexedBiDiGraph, lam, rho, T, gamma=0.5, alpha=0.0, phi=None, psi=None):
        self.g, self.lam, self.rho, self.alpha, self.T = g, lam, rho, alpha, int(T)
        gam = np.broadcast_to(np.asarray(gamma, dtype=float), (g.N,))
        self.phi = [[np.array([1 - gam[i], gam[i]]) if t == 0 else np.ones(2) for t in range(T + 1)] for i in range(g.N)] if phi is None else phi
        self.psi = psi  # per directed edge


class SIS_heterogeneous:
    """src/Models/epidemics/sis_heterogeneous.jl:1-49.  `lam` is an N x N matrix (dense or scipy sparse) whose entry
    [j, i] is the probability that j infects i; node i's factor takes column i restricted to its neighbours, in
    ascending neighbour order (`sis_heterogeneous_factors`, :51-53).  `rho`, `alpha`: per-node vectors."""

    def __init__(self, g: IndexedBiDiGraph, lam, rho, T, alpha=None, gamma=0.5, phi=None, psi=None):
        lam = np.asarray(lam.todense() if hasattr(lam, "todense") else lam, dtype=float)
        assert lam.shape == (g.N, g.N)
        self.g, self.lam, self.T = g, lam, int(T)
        self.rho = np.broadcast_to(np.asarray(rho, dtype=float), (g.N,))
        self.alpha = np.zeros(g.N) if alpha is None else np.broadcast_to(np.asarray(alpha, dtype=float), (g.N,))
        gam = np.broadcast_to(np.asarray(gamma, dtype=float), (g.N,))
        self.phi = [[np.array([1 - gam[i], gam[i]]) if t == 0 else np.ones(2) for t in range(T + 1)] for i in range(g.N)] if phi is None else phi
        self.psi = psi  # per directed edge

    def factors(self):
        g = self.g
        return [[SIS_heterogeneousFactor([self.lam[int(j), i] for j in g.neighbors(i)], self.rho[i], self.alpha[i])] * (self.T + 1)
                for i in range(g.N)]


class SIRS:
    def __init__(self, g: IndexedBiDiGraph, lam, rho, sigma, T, gamma=0.5, alpha=0.0, phi=None, psi=None):
        self.g, self.lam, self.rho, self.sigma, self.alpha, self.T = g, lam, rho, sigma, alpha, int(T)
        self.phi = [[np.array([1 - gamma, gamma, 0.0]) if t == 0 else np.ones(3) for t in range(T + 1)] for _ in range(g.N)] if phi is None else phi
        self.psi = psi


def glauber_factors(ising: Ising, T: int):
    """src/Models/glauber/glauber_bp.jl:121-142"""
    g, beta = ising.g, ising.beta
    absconst = len(ising.J) == 0 or np.all(np.abs(ising.J) == abs(ising.J[0]))
    homog = len(ising.J) == 0 or np.all(ising.J == ising.J[0])
    out = []
    for i in range(g.N):
        J = ising.J[g.und_of[g.colptr[i]:g.colptr[i + 1]]]
        h = ising.h[i]
        if absconst:
            Ji = 0.0 if len(J) == 0 else J[0]
            if homog:
                w = HomogeneousGlauberFactor(Ji, h, beta)
            else:
                w = PMJGlauberFactor(np.sign(J).astype(int), beta * abs(Ji), beta * h)
        elif np.all(J == np.round(J)):
            w = IntegerGlauberFactor(J.astype(int), h, beta)
        else:
            w = GenericGlauberFactor(J, h, beta)
        out.append([w] * (T + 1))
    return out


# --------------------------------------------------------------------------------------
# the CUDA-backed MPBP state
# --------------------------------------------------------------------------------------
class MPBP:
    """MPBP{G,F,V,M2,M1} with a device message store (src/mpbp.jl:1-33).  Fields g, w, phi, psi live on the host
    and are mirrored to the device; mu, b, f live on the device and are read through beliefs / pair_beliefs /
    bethe_free_energy / get_message."""

    def __init__(self, g, w, q, T, phi=None, psi=None, dmax=None, device=0, periodic=False):
        L = _lib.lib()
        self.g, self.w, self.T = g, w, int(T)
        self.periodic = bool(periodic)  # periodic_mpbp, src/mpbp.jl:399-409: PeriodicMPEM2 messages (ring tensor trains)
        self.q = np.ascontiguousarray(np.asarray(q, dtype=np.int32))
        self.N = g.N
        self.bipartite = isinstance(g, InfiniteBipartiteRegularGraph)
        self.infinite = isinstance(g, InfiniteRegularGraph) or self.bipartite
        assert len(w) == self.N and all(len(wi) == T + 1 for wi in w), "w must hold T+1 factors per node"
        self.dmax = int(dmax) if dmax is not None else (8 if self.periodic else 16)
        self._h = C.c_void_p()
        # _emap: reference edge index -> engine edge index (identity except for the bipartite infinite graph, whose
        # reference slot e = "message into node e" is the engine's edge 1 - e = (1-e -> e))
        if self.bipartite:
            _lib.check(L.mpbp_create_infinite_bipartite(g.k[0], g.k[1], self.T, int(self.q[0]), int(self.q[1]), self.dmax, device, C.byref(self._h)))
            self._src = np.array([0, 1], dtype=np.int64)
            self._dst = np.array([1, 0], dtype=np.int64)
            self._emap = [1, 0]
        elif self.infinite:
            _lib.check(L.mpbp_create_infinite(g.k, self.T, int(self.q[0]), self.dmax, device, C.byref(self._h)))
            self._src = np.zeros(1, dtype=np.int64)
            self._dst = np.zeros(1, dtype=np.int64)
        else:
            create = L.mpbp_create_periodic if self.periodic else L.mpbp_create
            _lib.check(create(self.N, g.ne, self.T, _p(self.q, _lib.c_i32p), _p(g.colptr, _lib.c_i64p), _p(g.dst, _lib.c_i64p),
                              _p(g.rev, _lib.c_i64p), self.dmax, device, C.byref(self._h)))
            self._src, self._dst = g.src, g.dst
        if self.periodic and self.infinite:
            _lib.check(L.mpbp_set_option(self._h, b"periodic", 1.0))
        self.E2 = len(self._src)
        if not self.bipartite:
            self._emap = list(range(self.E2))
        self.phi = [[np.ones(self.q[i]) for _ in range(T + 1)] for i in range(self.N)] if phi is None else phi
        # psi[e] (reference index) is indexed [x_src, x_dst] of the ENGINE edge _emap[e]
        self.psi = [[np.ones((self.q[self._src[self._emap[e]]], self.q[self._dst[self._emap[e]]])) for _ in range(T + 1)] for e in range(self.E2)] if psi is None else psi
        self._classes_dirty = True
        self.sync_reweightings()

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                _lib.lib().mpbp_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    # ---- host -> device mirrors ----
    def sync_reweightings(self):
        """upload bp.phi / bp.psi (call again after editing them in place, e.g. after drawing observations)."""
        L = _lib.lib()
        phi = np.ascontiguousarray(np.concatenate([np.asarray(p, dtype=np.float64).ravel() for ph in self.phi for p in ph]))
        assert len(phi) == int(np.sum(self.q)) * (self.T + 1), "phi must be [N][T+1][q_i]"
        _lib.check(L.mpbp_set_phi(self._h, _p(phi, _lib.c_dp)))
        inv = np.argsort(self._emap)  # engine edge -> reference index
        psi = np.ascontiguousarray(np.concatenate([np.asarray(p, dtype=np.float64).ravel(order="F") for ee in range(self.E2) for p in self.psi[int(inv[ee])]]))
        _lib.check(L.mpbp_set_psi(self._h, _p(psi, _lib.c_dp)))

    def upload_phi(self, flat):
        """upload phi from one contiguous float64 host buffer laid out [i][t][x] (e.g. pinned memory), no staging copy"""
        flat = np.asarray(flat)
        assert flat.dtype == np.float64 and flat.flags.c_contiguous and flat.size == int(np.sum(self.q)) * (self.T + 1)
        _lib.check(_lib.lib().mpbp_set_phi(self._h, _p(flat, _lib.c_dp)))

    def sync_factors(self, active=None):
        """tabulate the factors (host) and upload one class per distinct (factor, degree, neighbour states)."""
        L = _lib.lib()
        cache = {}
        cls = np.zeros(self.N, dtype=np.int32)
        _lib.check(L.mpbp_clear_node_classes(self._h))  # every table is re-uploaded below
        for i in range(self.N):
            z = self.g.degree(i)
            if active is not None and not active[i]:
                cls[i] = -1
                continue
            qn = np.array([self.q[0]] * z if (self.infinite and not self.bipartite) else [self.q[j] for j in self.g.neighbors(i)], dtype=np.int32)
            wi = self.w[i]
            keys = [w.key() for w in wi]
            if not isinstance(wi[0], RecursiveBPFactor):
                # generic BPFactor: only the functor w(x', x_neighbours, x) exists -> dense table, exhaustive-trace path
                same = all(w is wi[0] for w in wi) or (keys[0] is not None and all(k == keys[0] for k in keys))
                ck = ("generic", keys[0], z, int(self.q[i]), tuple(qn.tolist())) if same and keys[0] is not None else None
                if ck is not None and ck in cache:
                    cls[i] = cache[ck]
                    continue
                if z == 0:
                    raise MPBPError("generic BPFactor on an isolated node is not supported")
                ws = [wi[0]] if same else list(wi)
                qi = int(self.q[i])
                tabs = []
                for w in ws:
                    tab = np.zeros([qi] + [int(v) for v in qn] + [qi])
                    for idx in np.ndindex(*tab.shape):
                        tab[idx] = w(idx[0] + 1, [v + 1 for v in idx[1:-1]], idx[-1] + 1)
                    tabs.append(tab.ravel(order="F"))
                wt = np.ascontiguousarray(np.concatenate(tabs))
                cid = C.c_int32()
                _lib.check(L.mpbp_add_generic_class(self._h, z, qi, _p(qn, _lib.c_i32p), len(ws), _p(wt, _lib.c_dp), C.byref(cid)))
                cls[i] = cid.value
                if ck is not None:
                    cache[ck] = cid.value
                continue
            same = all(w is wi[0] for w in wi) or (keys[0] is not None and all(k == keys[0] for k in keys))
            ck = None
            if same and keys[0] is not None:
                ck = (keys[0], z, int(self.q[i]), tuple(qn.tolist()))
                if ck in cache:
                    cls[i] = cache[ck]
                    continue
            ws = [wi[0]] if same else list(wi)
            tab = tabulate_class(ws, z, int(self.q[i]), qn)
            d1 = np.array([p[0] for p in tab["pairs"]], dtype=np.int32)
            d2 = np.array([p[1] for p in tab["pairs"]], dtype=np.int32)
            cid = C.c_int32()
            _lib.check(L.mpbp_add_node_class(self._h, z, int(self.q[i]), _p(qn, _lib.c_i32p), len(ws), _p(tab["ny"], _lib.c_i32p),
                                             _p(tab["pxy"], _lib.c_dp), len(d1), _p(d1, _lib.c_i32p), _p(d2, _lib.c_i32p),
                                             _p(tab["pyy"], _lib.c_dp), _p(tab["w"], _lib.c_dp), _p(tab["wd"], _lib.c_dp),
                                             _p(tab["minit"], _lib.c_dp), C.byref(cid)))
            cls[i] = cid.value
            if ck is not None:
                cache[ck] = cid.value
        _lib.check(L.mpbp_set_node_classes(self._h, _p(cls, _lib.c_i32p)))
        self._classes_dirty = False

    # ---- device <-> host messages (checkpoint / resume, parity) ----
    def get_message(self, e):
        L = _lib.lib()
        e = self._emap[e]
        bonds = np.zeros(self.T + 2, dtype=np.int32)
        need = C.c_int64()
        _lib.check(L.mpbp_get_message(self._h, e, _p(bonds, _lib.c_i32p), None, 0, C.byref(need)))
        data = np.zeros(need.value)
        _lib.check(L.mpbp_get_message(self._h, e, _p(bonds, _lib.c_i32p), _p(data, _lib.c_dp), need.value, C.byref(need)))
        qs, qd = int(self.q[self._src[e]]), int(self.q[self._dst[e]])
        out, off = [], 0
        for t in range(self.T + 1):
            n = bonds[t] * bonds[t + 1] * qs * qd
            out.append(data[off:off + n].reshape((bonds[t], bonds[t + 1], qs, qd), order="F").copy())
            off += n
        return out

    def set_message(self, e, tensors):
        L = _lib.lib()
        e = self._emap[e]
        bonds = np.array([t.shape[0] for t in tensors] + [tensors[-1].shape[1]], dtype=np.int32)
        data = np.ascontiguousarray(np.concatenate([np.asarray(t, dtype=np.float64).ravel(order="F") for t in tensors]))
        _lib.check(L.mpbp_set_message(self._h, e, _p(bonds, _lib.c_i32p), _p(data, _lib.c_dp)))

    def counters(self, reset=False):
        fam = np.zeros(4)
        _lib.check(_lib.lib().mpbp_family_flops(self._h, _p(fam, _lib.c_dp)))  # (read before the reset below clears them)
        out = np.zeros(8)
        _lib.check(_lib.lib().mpbp_counters(self._h, _p(out, _lib.c_dp), int(reset)))
        return dict(launches=out[0], qr_flops=out[1], qr_ms=out[3], ops=out[4], edge_updates=out[5], svd_calls=out[2], svd_iters=out[6], svd_unconverged=out[7],
                    qr_split_extra_flops=fam[1], kron_carry_flops=fam[2], svd_subspace_flops=fam[3])

    def kernel_times(self, reset=False):
        out = np.zeros(9)
        _lib.check(_lib.lib().mpbp_kernel_times(self._h, _p(out, _lib.c_dp), 9, int(reset)))
        names = ["qr_sweep1", "kron_carry", "kron_proj", "gemm_m2t", "qr_small", "jacobi_project", "finalize", "belief", "btilde"]
        return dict(zip(names, out.tolist()))

    def set_stream(self, cuda_stream_ptr):
        _lib.check(_lib.lib().mpbp_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def set_option(self, name, value):
        _lib.check(_lib.lib().mpbp_set_option(self._h, name.encode(), float(value)))


class CB_BP:
    """convergence callback record (src/mpbp.jl:157-183): Delta per iteration; `f` is the observable f(x, i)."""

    def __init__(self, bp=None, f=None, showprogress=False):
        self.f = f
        self.deltas = []
        self.showprogress = showprogress

    @property
    def Δs(self):
        return self.deltas


def mpbp(*args, **kw):
    """mpbp(g, w, q, T; phi, psi, dmax) or mpbp(model; dmax)"""
    if len(args) == 1:
        m = args[0]
        if isinstance(m, Glauber):
            g = m.ising.g
            w = glauber_factors(m.ising, m.T)
            psi = None
            if m.psi is not None:  # pair_obs_undirected_to_directed, src/mpbp.jl:375-396
                psi = []
                for e in range(g.ne):
                    u = g.und_of[e]
                    psi.append([np.asarray(p) if g.src[e] < g.dst[e] else np.asarray(p).T for p in m.psi[u]])
            return MPBP(g, w, [2] * g.N, m.T, phi=m.phi, psi=psi, **kw)
        if isinstance(m, SIS):
            w = [[SISFactor(m.lam, m.rho, m.alpha)] * (m.T + 1)] * m.g.N
            return MPBP(m.g, w, [2] * m.g.N, m.T, phi=m.phi, psi=m.psi, **kw)
        if isinstance(m, SIS_heterogeneous):
            return MPBP(m.g, m.factors(), [2] * m.g.N, m.T, phi=m.phi, psi=m.psi, **kw)
        if isinstance(m, SIRS):
            w = [[SIRSFactor(m.lam, m.rho, m.sigma, m.alpha)] * (m.T + 1)] * m.g.N
            return MPBP(m.g, w, [3] * m.g.N, m.T, phi=m.phi, psi=m.psi, **kw)
        raise TypeError(f"no mpbp method for {type(m)}")
    g, w, q, T = args
    return MPBP(g, w, q, T, **kw)


def periodic_mpbp(*args, **kw):
    """periodic_mpbp(g, w, q, T; phi, psi, dmax) or periodic_mpbp(model; dmax) -- src/mpbp.jl:399-409: the messages are
    periodic-in-time MPEMs (the factor at the last time maps (x_neigh^T, x_i^T) to x_i^0)."""
    return mpbp(*args, periodic=True, **kw)


def periodic_mpbp_infinite_graph(k, w, q, phi=None, psi=None, **kw):
    """periodic_mpbp_infinite_graph(k, w, q, phi) -- test/periodic.jl:78-93"""
    return mpbp_infinite_graph(k, w, q, phi=phi, psi=psi, periodic=True, **kw)


def mpbp_infinite_bipartite_graph(k, w, q, phi=None, psi=None, **kw):
    """mpbp_infinite_bipartite_graph((kA, kB), [wA, wB], (qA, qB), [phiA, phiB]; psi) -- src/infinite_graph.jl:87-105"""
    T = len(w[0]) - 1
    assert len(w[1]) == T + 1
    return MPBP(InfiniteBipartiteRegularGraph(k), [list(w[0]), list(w[1])], [int(q[0]), int(q[1])], T, phi=phi, psi=psi, **kw)


def mpbp_infinite_graph(k, w, q, phi=None, psi=None, **kw):
    T = len(w) - 1
    return MPBP(InfiniteRegularGraph(k), [list(w)], [q], T, phi=None if phi is None else [phi], psi=None if psi is None else [psi], **kw)


def iterate_(bp: MPBP, maxiter=5, svd_trunc: SVDTrunc = None, showprogress=False, cb: CB_BP = None, tol=1e-10, nodes=None,
             shuffle_nodes=True, damp=0.0, schedule="sequential", rng=None):
    """iterate!(bp; ...) -> (iters, cb).  `schedule`: "sequential" = the reference's in-place sweep in node
    order (one thread), "parallel" = Jacobi (all nodes read the previous iteration's messages).  With
    shuffle_nodes the visiting order of every iteration is a fresh permutation drawn from `rng`."""
    L = _lib.lib()
    svd_trunc = TruncThresh(1e-6) if svd_trunc is None else svd_trunc  # default_truncator, src/mpems.jl:161
    if bp._classes_dirty:
        bp.sync_factors()
    cb = CB_BP(bp) if cb is None else cb
    nodes_arr = np.arange(bp.N, dtype=np.int64) if nodes is None else np.ascontiguousarray(np.asarray(nodes, dtype=np.int64))
    shuffle = shuffle_nodes and schedule == "sequential" and len(nodes_arr) > 1
    obs = None
    if cb.f is not None:
        qmax = int(bp.q.max())
        obs = np.zeros((bp.N, qmax))
        for i in range(bp.N):
            for x in range(int(bp.q[i])):
                obs[i, x] = cb.f(x + 1, i)
        obs = np.ascontiguousarray(obs)
    iters = C.c_int()
    sched = {"sequential": 0, "parallel": 1}[schedule]

    def call(n_it, nodes_now, deltas):
        _lib.check(L.mpbp_iterate(bp._h, int(n_it), svd_trunc.kind, svd_trunc.d, svd_trunc.eps, float(tol), float(damp), sched,
                                  _p(nodes_now, _lib.c_i64p), len(nodes_now), None,
                                  None if obs is None else _p(obs, _lib.c_dp), C.byref(iters), _p(deltas, _lib.c_dp)))

    if shuffle:
        # src/mpbp.jl:188-196: the first sweep visits `nodes` in the given order; after every sweep
        # `sample!(nodes, vertices(bp.g), replace=false)` redraws the list from ALL vertices (a permutation when `nodes`
        # covers the graph).  One C call per iteration with that iteration's list: nothing of size maxiter x N is
        # materialised.  Delta is taken over the nodes just updated (the others did not move).
        rng = np.random.default_rng(rng)
        d1 = np.zeros(1)
        for it in range(int(maxiter)):
            order = nodes_arr if it == 0 else np.ascontiguousarray(rng.choice(bp.N, size=len(nodes_arr), replace=False).astype(np.int64))
            call(1, order, d1)
            cb.deltas.append(float(d1[0]))
            if d1[0] < tol:
                return it + 1, cb
        return int(maxiter), cb
    deltas = np.zeros(max(maxiter, 1))
    call(maxiter, nodes_arr, deltas)
    cb.deltas.extend(deltas[:iters.value].tolist())
    return iters.value, cb


def beliefs(bp: MPBP):
    """beliefs(bp)[i][t][x]"""
    out = np.zeros(int(np.sum(bp.q)) * (bp.T + 1))
    _lib.check(_lib.lib().mpbp_beliefs(bp._h, _p(out, _lib.c_dp)))
    res, off = [], 0
    for i in range(bp.N):
        qi = int(bp.q[i])
        res.append(out[off:off + qi * (bp.T + 1)].reshape(bp.T + 1, qi).copy())
        off += qi * (bp.T + 1)
    return res


def means(f, bp: MPBP):
    return [[sum(f(x + 1, i) * p[x] for x in range(len(p))) for p in b] for i, b in enumerate(beliefs(bp))]


def beliefs_tu(bp: MPBP):
    """beliefs_tu(bp)[i][t][u] = b_i(x^t, x^u) as a q x q array for t < u <= t + maxdist (src/mpbp.jl:239), None
    elsewhere.  The two-time marginals are computed together with the beliefs during `iterate_`, so they must be
    switched on first: ``bp.set_option("twovar", maxdist)`` (maxdist = bp.T for all pairs)."""
    L, qm = bp.T + 1, int(np.max(bp.q))
    out = np.zeros(bp.N * L * L * qm * qm)
    _lib.check(_lib.lib().mpbp_twovar_marginals(bp._h, _p(out, _lib.c_dp)))
    out = out.reshape(bp.N, L, L, qm * qm)
    res = []
    for i in range(bp.N):
        qi = int(bp.q[i])
        row = [[None] * L for _ in range(L)]
        for t in range(L):
            for u in range(t + 1, L):
                blk = out[i, t, u, :qi * qi]
                if blk.any():
                    row[t][u] = blk.reshape(qi, qi, order="F").copy()
        res.append(row)
    return res


def autocorrelations(f, bp: MPBP):
    """autocorrelations(f, bp)[i][t, u] = <f(x_i^t, i) f(x_i^u, i)> for t < u, zero elsewhere (src/mpbp.jl:245-255);
    states are numbered from 1 as in the reference."""
    res = []
    for i, tv in enumerate(beliefs_tu(bp)):
        L, qi = bp.T + 1, int(bp.q[i])
        fx = np.array([f(x + 1, i) for x in range(qi)], dtype=float)
        r = np.zeros((L, L))
        for t in range(L):
            for u in range(t + 1, L):
                if tv[t][u] is not None:
                    r[t, u] = fx @ tv[t][u] @ fx
        res.append(r)
    return res


def autocovariances(f, bp: MPBP):
    """covariance.(r, mu) = r - mu mu' on the whole matrix (src/mpbp.jl:287-296)"""
    mu = means(f, bp)
    return [r - np.outer(m, m) for r, m in zip(autocorrelations(f, bp), mu)]


def pair_beliefs(bp: MPBP):
    """pair_beliefs(bp) -> (b[e][t][x_src, x_dst], logz[i])"""
    sizes = [int(bp.q[bp._src[e]]) * int(bp.q[bp._dst[e]]) for e in range(bp.E2)]
    out = np.zeros(sum(sizes) * (bp.T + 1))
    logz = np.zeros(bp.N)
    _lib.check(_lib.lib().mpbp_pair_beliefs(bp._h, _p(out, _lib.c_dp), _p(logz, _lib.c_dp)))
    res, off = [], 0
    for e in range(bp.E2):
        qs, qd = int(bp.q[bp._src[e]]), int(bp.q[bp._dst[e]])
        n = qs * qd * (bp.T + 1)
        res.append(np.stack([out[off + t * qs * qd: off + (t + 1) * qs * qd].reshape(qs, qd, order="F") for t in range(bp.T + 1)]))
        off += n
    return [res[bp._emap[e]] for e in range(bp.E2)], logz


def pair_correlations(f, bp: MPBP):
    """<f(x_i^t) f(x_j^t)> per directed edge i->j (src/mpbp.jl:263-267); states numbered from 1"""
    res = []
    for pb in pair_beliefs(bp)[0]:
        fi = np.array([f(x + 1) for x in range(pb.shape[1])], dtype=float)
        fj = np.array([f(x + 1) for x in range(pb.shape[2])], dtype=float)
        res.append([float(fi @ p @ fj) for p in pb])
    return res


def alternate_marginals(bp: MPBP):
    """alternate_marginals(bp)[e][t][x_i^t, x_j^{t+1}] for every directed edge e = i->j, t = 0..T-1 (src/mpbp.jl:270-280)"""
    sizes = [int(bp.q[bp._src[e]]) * int(bp.q[bp._dst[e]]) for e in range(bp.E2)]
    out = np.zeros(sum(sizes) * (bp.T + 1))
    _lib.check(_lib.lib().mpbp_alternate_marginals(bp._h, _p(out, _lib.c_dp)))
    res, off = [], 0
    for e in range(bp.E2):
        qs, qd = int(bp.q[bp._src[e]]), int(bp.q[bp._dst[e]])
        res.append([out[off + t * qs * qd: off + (t + 1) * qs * qd].reshape(qs, qd, order="F").copy() for t in range(bp.T)])
        off += qs * qd * (bp.T + 1)
    return [res[bp._emap[e]] for e in range(bp.E2)]


def alternate_correlations(f, bp: MPBP):
    """<f(x_i^t) f(x_j^{t+1})> per directed edge (src/mpbp.jl:282-286); states numbered from 1"""
    res = []
    for am in alternate_marginals(bp):
        row = []
        for p in am:
            fi = np.array([f(x + 1) for x in range(p.shape[0])], dtype=float)
            fj = np.array([f(x + 1) for x in range(p.shape[1])], dtype=float)
            row.append(float(fi @ p @ fj))
        res.append(row)
    return res


def free_energy_contributions(bp: MPBP):
    f = np.zeros(bp.N)
    _lib.check(_lib.lib().mpbp_free_energy(bp._h, _p(f, _lib.c_dp)))
    return f


def bethe_free_energy(bp: MPBP):
    f = free_energy_contributions(bp)
    if bp.bipartite:  # reweighted by the fraction of nodes in each block (src/infinite_graph.jl:120-122)
        k = bp.g.k
        return float((f[0] * k[1] + f[1] * k[0]) / (k[0] + k[1]))
    return float(np.sum(f))


def sample_prior(bp: MPBP, seed):
    """X[i, t] (0-based states, N x (T+1) int32) of one forward simulation of the prior dynamics, drawn ON THE DEVICE from the
    uploaded factor tables (mpbp_sample_prior); deterministic in `seed`"""
    if bp._classes_dirty:
        bp.sync_factors()
    X = np.zeros((bp.N, bp.T + 1), dtype=np.int32)
    _lib.check(_lib.lib().mpbp_sample_prior(bp._h, C.c_uint64(int(seed) & ((1 << 64) - 1)), _p(X, _lib.c_i32p)))
    return X


def onesample(bp: MPBP, rng=None):
    """onesample(bp) -> (X, weight): forward sample of the prior dynamics of bp (src/sampling.jl:30-66), states numbered from 1
    as in the reference.  The trajectory is drawn on the device; the likelihood weight exp(sum_{t>0} log phi + 1/2 sum_e log
    psi) is a vectorised host reduction over it."""
    rng = np.random.default_rng(rng)
    X0 = sample_prior(bp, int(rng.integers(0, 2 ** 63 - 1)))
    L = bp.T + 1
    logl = 0.0
    with np.errstate(divide="ignore"):
        for i in range(bp.N):
            ph = np.asarray([np.asarray(p, dtype=float) for p in bp.phi[i]])  # [t][x]
            logl += float(np.sum(np.log(ph[np.arange(1, L), X0[i, 1:]])))
        for e in range(bp.E2):
            ee = bp._emap[e]
            i, j = int(bp._src[ee]), int(bp._dst[ee])
            ps = np.asarray([np.asarray(p, dtype=float) for p in bp.psi[e]])  # [t][xi, xj]
            logl += 0.5 * float(np.sum(np.log(ps[np.arange(L), X0[i], X0[j]])))
    return X0.astype(np.int64) + 1, float(np.exp(logl))


def draw_node_observations_(bp: MPBP, nobs, rng=None, **kw):
    """draw_node_observations!(bp, nobs): sample a trajectory from the prior, observe `nobs` of its entries (reweightings
    bp.phi updated in place and uploaded), return (X, observed) -- src/sampling.jl:205-210"""
    from .sampling import draw_node_observations_ as _draw
    rng = np.random.default_rng(rng)
    X, _ = onesample(bp, rng)
    _, observed = _draw(bp.phi, X, nobs, rng=rng, **kw)
    bp.sync_reweightings()
    return X, observed


def reset_messages_(bp: MPBP):
    _lib.check(_lib.lib().mpbp_reset_messages(bp._h))
