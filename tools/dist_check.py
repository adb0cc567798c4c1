"""2-rank NCCL check (run under torchrun on a 2-GPU box): the node-partitioned run must reproduce the
single-GPU Jacobi run (beliefs, free energy) of the same graph."""
import os, sys
import numpy as np
sys.path.insert(0, '.')
import torch, torch.distributed as dist
import networkx as nx
import mpbp_b200 as M
from mpbp_b200.dist import CudaBackend, DistMPBP, LocalProblem, partition_contiguous

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{lr}"))
N, T, d, iters = 24, 6, 6, 3
G = nx.fast_gnp_random_graph(N, 3.0 / N, seed=2)
und = [(int(a), int(b)) for a, b in G.edges()]
fac = M.HomogeneousGlauberFactor(0.5, 0.1, 1.0)
def make(gl, nloc, dev):
    w = [[fac] * (T + 1)] * nloc
    phi = [[np.array([0.2, 0.8]) if t == 0 else np.ones(2) for t in range(T + 1)] for _ in range(nloc)]
    return M.mpbp(gl, w, [2] * nloc, T, phi=phi, dmax=d, device=dev)
owner = partition_contiguous(N, world)
lp = LocalProblem(N, und, owner, rank)
g = M.IndexedBiDiGraph(len(lp.nodes), lp.local_und)
lp.build_exchange(g.src, g.dst, world)
bp = make(g, g.N, lr)
drv = DistMPBP(lp, CudaBackend(bp, lp.owned_local, M.TruncBond(d)), dist, device=f"cuda:{lr}")
its, deltas = drv.iterate(iters)
bel = M.beliefs(bp); f = M.api.free_energy_contributions(bp)
mine = {int(lp.nodes[i]): (bel[int(i)], float(f[int(i)])) for i in lp.owned_local}
allr = [None] * world
dist.all_gather_object(allr, mine)
if rank == 0:
    gf = M.IndexedBiDiGraph(N, und)
    ref = make(gf, N, lr)
    _, cb = M.iterate_(ref, maxiter=iters, svd_trunc=M.TruncBond(d), tol=0.0, shuffle_nodes=False, schedule="parallel")
    rb = M.beliefs(ref); rf = M.api.free_energy_contributions(ref)
    merged = {}
    for m in allr: merged.update(m)
    eb = max(float(np.max(np.abs(merged[i][0] - rb[i]))) for i in range(N))
    ef = max(abs(merged[i][1] - rf[i]) for i in range(N))
    ed = float(np.max(np.abs(np.array(deltas) - np.array(cb.deltas))))
    print(f"dist_check world={world}: max|belief diff|={eb:.2e} max|f diff|={ef:.2e} max|delta diff|={ed:.2e}", "OK" if max(eb, ef, ed) < 1e-10 else "FAIL")
dist.destroy_process_group()
