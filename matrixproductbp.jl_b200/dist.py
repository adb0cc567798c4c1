"""Multi-GPU driver: node-partitioned MPBP with one halo exchange of cut-edge messages per BP iteration.

No reference counterpart (the reference is single-process).  One process per GPU;
``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU tests) is the plumbing.  Each rank
builds the LOCAL graph = owned nodes + their remote neighbours ("halo" nodes, never updated locally) with
all directed edges between them, runs the Jacobi node update for its owned nodes, then exchanges the
freshly written messages whose destination is owned by a peer: message (i -> j), i owned here, j owned by
peer p, is p's in-message.  Fixed-capacity slots make the payload sizes static.

The driver is backend-agnostic: ``backend`` needs ``iterate_owned()``, ``pack(local_edges) -> uint8 tensor``,
``unpack(local_edges, tensor)`` and ``slot_bytes``.  The product backend is :class:`CudaBackend` (C-ABI
pack/unpack on device buffers); the tests inject an oracle-based backend to exercise the exchange logic
on CPU with gloo.
"""
from __future__ import annotations

import numpy as np


def partition_contiguous(N, world):
    """owner[i] for a contiguous block partition (NVSwitch bandwidth is uniform: locality does not matter,
    only balance; callers may pass any other owner array)."""
    return np.minimum(np.arange(N) * world // N, world - 1).astype(np.int64)


def node_update_cost(z, q=2):
    """Relative cost of one node update of degree z: the sum over the heavy cavity ops (prefix, suffix and
    destination products, both operands of full bond) of the row multiplier X = nstates * q of the sweep-1
    matrix, plus one unit for the light ops.  The QR and the truncating SVD are both linear in X at capped bonds."""
    z = int(z)
    w = 1.0
    for k in range(1, z):
        w += q * (k + 2)
    for k in range(1, z - 1):
        w += q * (z - k + 1) + q * z
    return w


def partition_balanced(N, und_edges, world, cost=node_update_cost):
    """owner[i] balancing the summed node-update cost over ranks (longest-processing-time greedy on the degree
    cost model; ties keep lower node ids on lower ranks).  The result of the run does not depend on the
    partition (:class:`LocalProblem` preserves neighbour order), only the time per iteration does."""
    und = np.asarray(und_edges, dtype=np.int64).reshape(-1, 2)
    deg = np.bincount(und.reshape(-1), minlength=N)
    c = np.array([cost(z) for z in deg], dtype=np.float64)
    owner = np.zeros(N, dtype=np.int64)
    load = np.zeros(world)
    count = np.zeros(world, dtype=np.int64)
    for i in np.argsort(-c, kind="stable"):
        r = int(np.lexsort((count, load))[0])  # least loaded, then fewest nodes
        owner[i] = r
        load[r] += c[i]
        count[r] += 1
    return owner


class LocalProblem:
    """Local view of rank `rank`: node and edge maps between the global graph and the local subgraph."""

    def __init__(self, N, und_edges, owner, rank):
        und = np.asarray(und_edges, dtype=np.int64).reshape(-1, 2)
        self.rank = rank
        self.owner = np.asarray(owner)
        mine = self.owner == rank
        keep = mine[und[:, 0]] | mine[und[:, 1]]
        lund = und[keep]
        owned = np.nonzero(mine)[0]
        # local ids preserve the global order, so every node sees its neighbours in the same order as in the
        # single-process run (the order of the truncated cavity products, hence the results, depend on it)
        self.nodes = np.union1d(owned, np.unique(lund))  # local -> global, ascending
        self.n_owned = len(owned)
        self.g2l = -np.ones(N, dtype=np.int64)
        self.g2l[self.nodes] = np.arange(len(self.nodes))
        self.local_und = [(int(self.g2l[a]), int(self.g2l[b])) for a, b in lund]
        self.owned_local = self.g2l[owned].astype(np.int64)

    def build_exchange(self, lsrc, ldst, world):
        """lsrc/ldst: local directed edge arrays (local node ids, reference edge order).
        send[p] = local edges (i->j) with i owned, j owned by p ; recv[p] = local edges (j->i) with j owned by p,
        i owned -- both sorted by the GLOBAL (src,dst) key so that sender and receiver agree on the order."""
        gs, gd = self.nodes[lsrc], self.nodes[ldst]
        N = len(self.owner)
        key = gs * N + gd
        self.send, self.recv = [], []
        for p in range(world):
            if p == self.rank:
                self.send.append(np.zeros(0, dtype=np.int64))
                self.recv.append(np.zeros(0, dtype=np.int64))
                continue
            s = np.nonzero((self.owner[gs] == self.rank) & (self.owner[gd] == p))[0]
            r = np.nonzero((self.owner[gs] == p) & (self.owner[gd] == self.rank))[0]
            self.send.append(s[np.argsort(key[s])].astype(np.int64))
            self.recv.append(r[np.argsort(key[r])].astype(np.int64))


class DistMPBP:
    """Jacobi MPBP over `world` ranks.  `dist` is torch.distributed (already initialised) or None for world=1."""

    def __init__(self, local: LocalProblem, backend, dist=None, device="cpu"):
        self.local, self.backend, self.dist, self.device = local, backend, dist, device
        self.world = dist.get_world_size() if dist is not None else 1

    def halo_exchange(self):
        import torch
        if self.world == 1:
            return 0
        sb = self.backend.slot_bytes
        send_edges = np.concatenate(self.local.send)
        recv_edges = np.concatenate(self.local.recv)
        sbuf = self.backend.pack(send_edges)  # uint8 tensor on self.device, len(send_edges)*sb
        rbuf = torch.empty(len(recv_edges) * sb, dtype=torch.uint8, device=self.device)
        in_split = [len(e) * sb for e in self.local.send]
        out_split = [len(e) * sb for e in self.local.recv]
        self.dist.all_to_all_single(rbuf, sbuf, output_split_sizes=out_split, input_split_sizes=in_split)
        self.backend.unpack(recv_edges, rbuf)
        return int(sbuf.numel())

    def iterate(self, maxiter, tol=0.0):
        """returns (iters, deltas): Delta is the max over ranks (all-reduce MAX), as CB_BP would see it."""
        import torch
        import time
        deltas = []
        for it in range(maxiter):
            t0 = time.perf_counter()
            d = self.backend.iterate_owned()  # synchronous (mpbp_iterate returns after the device is done)
            self.compute_s = getattr(self, "compute_s", 0.0) + (time.perf_counter() - t0)
            self.halo_exchange()
            if self.world > 1:
                t = torch.tensor([d], dtype=torch.float64, device=self.device)
                self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
                d = float(t.item())
            deltas.append(d)
            if d < tol:
                return it + 1, deltas
        return maxiter, deltas


class CudaBackend:
    """product backend: the CUDA engine through the C-ABI (no CPU fallback)."""

    def __init__(self, bp, owned_local, svd_trunc):
        import torch
        from . import _lib
        self._lib = _lib
        self.bp, self.owned, self.trunc = bp, np.ascontiguousarray(owned_local, dtype=np.int64), svd_trunc
        self.slot_bytes = int(_lib.lib().mpbp_message_slot_bytes(bp._h))
        self.torch = torch
        active = np.zeros(bp.N, dtype=bool)
        active[self.owned] = True
        bp.sync_factors(active=active)

    def iterate_owned(self):
        from .api import iterate_
        iters, cb = iterate_(self.bp, maxiter=1, svd_trunc=self.trunc, tol=0.0, nodes=self.owned, shuffle_nodes=False,
                             schedule="parallel")
        return cb.deltas[-1]

    def pack(self, edges):
        edges = np.ascontiguousarray(edges, dtype=np.int64)
        buf = self.torch.empty(len(edges) * self.slot_bytes, dtype=self.torch.uint8, device=f"cuda:{self.torch.cuda.current_device()}")
        if len(edges):
            self._lib.check(self._lib.lib().mpbp_pack_messages_dev(self.bp._h, len(edges), edges.ctypes.data_as(self._lib.c_i64p), buf.data_ptr()))
        return buf

    def unpack(self, edges, buf):
        edges = np.ascontiguousarray(edges, dtype=np.int64)
        # The collective that filled `buf` is ordered on torch's CURRENT stream (all_to_all_single makes it wait for
        # NCCL); the engine scatters on its OWN stream, which knows nothing about either: drain the current stream first.
        self.torch.cuda.current_stream().synchronize()
        if len(edges):
            self._lib.check(self._lib.lib().mpbp_unpack_messages_dev(self.bp._h, len(edges), edges.ctypes.data_as(self._lib.c_i64p), buf.data_ptr()))
