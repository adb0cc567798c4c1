"""repro: bipartite infinite graph with a bond overflow (dmax too small for TruncThresh(0)); must raise the loud error and
nothing else (run under compute-sanitizer)"""
import sys
sys.path.insert(0, ".")
import numpy as np
import mpbp_b200 as M
T, k, m0 = 3, (3, 2), 0.5
phi = [[np.array([(1 + m0) / 2, (1 - m0) / 2]) if t == 0 else np.ones(2) for t in range(T + 1)] for _ in range(2)]
wd = [[M.HomogeneousGlauberFactor(1.0, -0.1, 1.0)] * (T + 1), [M.HomogeneousGlauberFactor(-0.2, -0.1, 1.0)] * (T + 1)]
for damp in (0.0, 0.1):
    bp = M.mpbp_infinite_bipartite_graph(k, wd, (2, 2), phi=[[p.copy() for p in ph] for ph in phi], dmax=16)
    try:
        it, _ = M.iterate_(bp, maxiter=5, svd_trunc=M.TruncThresh(0.0), tol=1e-14, damp=damp, shuffle_nodes=False)
        print("damp", damp, "iterations", it)
    except M.MPBPError as e:
        print("damp", damp, "loud error:", e)
    del bp
print("done")
