"""N>1 path on CPU: world_size-2 gloo run of the node-partitioned driver (mpbp_b200.dist) with an oracle-backed
compute backend injected by the test, compared with the single-process Jacobi run of the oracle.  This covers the
partition, the halo send/recv lists, the all_to_all exchange and the Delta all-reduce; the CUDA backend's pack /
unpack are covered by tests/test_gpu_parity.py::test_pack_unpack_roundtrip."""
import os
import socket

import numpy as np
import pytest

from oracle import factors as OF, mpbp as O, tt as OT

T, D = 3, 4
UND = [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 0), (1, 4), (2, 6), (6, 7), (7, 3)]
N = 8


def _problem():
    w = [[OF.SISFactor(0.2 + 0.01 * i, 0.15, 0.02)] * (T + 1) for i in range(N)]
    phi = [[np.array([0.8 - 0.02 * i, 0.2 + 0.02 * i]) if t == 0 else np.ones(2) for t in range(T + 1)] for i in range(N)]
    phi[3][2] = np.array([0.3, 1.0])
    return w, phi


class OracleBackend:
    """test-only backend: the CPU oracle does the node updates; messages travel in fixed-size float64 slots"""

    def __init__(self, bp, owned, trunc):
        import torch
        self.torch = torch
        self.bp, self.owned, self.trunc = bp, list(owned), trunc
        self.L = bp.T + 1
        self.slot_doubles = self.L * D * D * 4 + (self.L + 1) + 1
        self.slot_bytes = 8 * self.slot_doubles
        self.prev = [O.means(bp)[i] for i in self.owned]

    def iterate_owned(self):
        new = list(self.bp.mu)
        for i in self.owned:
            O.onebpiter(self.bp, i, self.trunc, mu_read=self.bp.mu, mu_write=new)
        self.bp.mu = new
        m = O.means(self.bp)
        cur = [m[i] for i in self.owned]
        d = max(max(abs(a - b) for a, b in zip(x, y)) for x, y in zip(cur, self.prev))
        self.prev = cur
        return d

    def pack(self, edges):
        buf = np.zeros((len(edges), self.slot_doubles))
        for k, e in enumerate(edges):
            A = self.bp.mu[int(e)]
            bonds = A.bond_dims()
            off = 0
            for t in range(self.L):
                a = A[t].ravel(order="F")
                buf[k, off:off + a.size] = a
                off += D * D * 4
            buf[k, self.L * D * D * 4: self.L * D * D * 4 + self.L + 1] = bonds
            buf[k, -1] = A.ls
        return self.torch.from_numpy(buf.reshape(-1).view(np.uint8).copy())

    def unpack(self, edges, tbuf):
        buf = tbuf.numpy().view(np.float64).reshape(len(edges), self.slot_doubles)
        for k, e in enumerate(edges):
            bonds = buf[k, self.L * D * D * 4: self.L * D * D * 4 + self.L + 1].astype(int)
            tens = []
            for t in range(self.L):
                n = bonds[t] * bonds[t + 1] * 4
                tens.append(buf[k, t * D * D * 4: t * D * D * 4 + n].reshape((bonds[t], bonds[t + 1], 2, 2), order="F").copy())
            self.bp.mu[int(e)] = OT.TT(tens, float(buf[k, -1]))


def _worker(rank, world, port, iters, out, balanced=False):
    import torch.distributed as dist
    from mpbp_b200.dist import DistMPBP, LocalProblem, partition_contiguous
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w, phi = _problem()
    from mpbp_b200.dist import partition_balanced
    owner = partition_balanced(N, UND, world) if balanced else partition_contiguous(N, world)
    lp = LocalProblem(N, UND, owner, rank)
    g = O.BiDiGraph(len(lp.nodes), lp.local_und)
    lp.build_exchange(np.array(g.src), np.array(g.dst), world)
    wl = [w[int(i)] for i in lp.nodes]
    pl = [phi[int(i)] for i in lp.nodes]
    bp = O.MPBP(g, wl, [2] * g.N, T, phi=pl)
    drv = DistMPBP(lp, OracleBackend(bp, lp.owned_local, OT.TruncBond(D)), dist, device="cpu")
    its, deltas = drv.iterate(iters)
    bel = {int(lp.nodes[i]): np.array(OT.marginals(bp.b[int(i)])) for i in lp.owned_local}
    f = {int(lp.nodes[i]): float(bp.f[int(i)]) for i in lp.owned_local}
    gathered = [None] * world
    dist.all_gather_object(gathered, (bel, f, deltas))
    if rank == 0:
        out.put(gathered)
    dist.destroy_process_group()


@pytest.mark.parametrize("balanced", [False, True])
def test_two_rank_halo_exchange_matches_single_process(balanced):
    import torch.multiprocessing as mp
    iters = 3
    w, phi = _problem()
    g = O.BiDiGraph(N, UND)
    ref = O.MPBP(g, w, [2] * N, T, phi=phi)
    _, ref_deltas = O.iterate(ref, maxiter=iters, trunc=OT.TruncBond(D), tol=0.0, schedule="parallel")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, iters, out, balanced)) for r in range(2)]
    for p in procs:
        p.start()
    gathered = out.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    bel, f = {}, {}
    for b, ff, deltas in gathered:
        bel.update(b)
        f.update(ff)
        assert np.allclose(deltas, ref_deltas, atol=1e-12)  # all-reduced Delta == single-process CB_BP Delta
    assert sorted(bel) == list(range(N))
    for i in range(N):
        assert np.allclose(bel[i], np.array(OT.marginals(ref.b[i])), atol=1e-12)
        assert abs(f[i] - ref.f[i]) < 1e-10


def test_partition_and_exchange_lists_are_consistent():
    from mpbp_b200.dist import LocalProblem, partition_contiguous
    import mpbp_b200 as M
    world = 3
    owner = partition_contiguous(N, world)
    sends, recvs = {}, {}
    for r in range(world):
        lp = LocalProblem(N, UND, owner, r)
        g = M.IndexedBiDiGraph(len(lp.nodes), lp.local_und)
        lp.build_exchange(g.src, g.dst, world)
        assert set(lp.nodes[lp.owned_local]) == set(np.nonzero(owner == r)[0])
        assert np.all(np.diff(lp.nodes) > 0)
        for p in range(world):
            sends[(r, p)] = [(int(lp.nodes[g.src[e]]), int(lp.nodes[g.dst[e]])) for e in lp.send[p]]
            recvs[(r, p)] = [(int(lp.nodes[g.src[e]]), int(lp.nodes[g.dst[e]])) for e in lp.recv[p]]
    for r in range(world):
        for p in range(world):
            assert sends[(r, p)] == recvs[(p, r)]  # same global edges, same order on both sides
    cut = sum(len(v) for v in sends.values())
    gfull = M.IndexedBiDiGraph(N, UND)
    assert cut == int(np.sum(owner[gfull.src] != owner[gfull.dst]))


def test_balanced_partition_covers_all_nodes_and_balances_cost():
    from mpbp_b200.dist import LocalProblem, node_update_cost, partition_balanced, partition_contiguous
    import networkx as nx
    n = 300
    G = nx.fast_gnp_random_graph(n, 4.0 / n, seed=1)
    und = [(int(a), int(b)) for a, b in G.edges()]
    deg = np.bincount(np.array(und).reshape(-1), minlength=n)
    cost = np.array([node_update_cost(z) for z in deg])
    for world in (2, 3, 8):
        owner = partition_balanced(n, und, world)
        assert owner.shape == (n,) and owner.min() == 0 and owner.max() == world - 1
        load = np.array([cost[owner == r].sum() for r in range(world)])
        lc = np.array([cost[partition_contiguous(n, world) == r].sum() for r in range(world)])
        assert load.max() / load.mean() < 1.01 and load.max() <= lc.max() + 1e-9
        # the local problems built from it are consistent: owned sets partition the nodes, neighbour order preserved
        seen = []
        for r in range(world):
            lp = LocalProblem(n, und, owner, r)
            seen.extend(int(x) for x in lp.nodes[lp.owned_local])
            assert np.all(np.diff(lp.nodes) > 0)
        assert sorted(seen) == list(range(n))
    # monotone in the degree, light nodes cost one unit
    assert node_update_cost(0) == 1.0 and all(node_update_cost(z + 1) > node_update_cost(z) for z in range(1, 12))
