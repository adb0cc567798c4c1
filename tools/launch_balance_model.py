"""CPU model of the per-launch load balance of the sweep-1 QR launches on the bench graph (no GPU needed).

Rebuilds the heavy-op list of every cavity round after the level staggering (same cost model as plan_level_offsets in
csrc/engine.cu), takes the CTA time of an op proportional to its row multiplier X (m = r*X rows, n = 400 columns) and
list-schedules one launch per round on 148 SMs in LPT order.  Prints, per round, makespan / ideal and the share of the
ideal time that the single longest CTA needs: where that share is >= 1 the launch is bound by ONE matrix and a TSQR
split of the outliers (even in full launches) pays.  Round-1 result: efficiency 0.87 overall, rounds 5-9 bound by
their X = 20-22 ops (share 0.95-1.16)."""
import heapq
import sys

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402

N, und = bench.make_workload(int(sys.argv[1]) if len(sys.argv) > 1 else 384)
deg = np.bincount(np.array(und).reshape(-1), minlength=N)
d, q = 20, 2


def node_ops(z):
    ops = {}

    def add(level, ca, cb, ny):
        if ca * cb == d * d:
            ops.setdefault(level, []).append(ny * q)

    if z >= 2:
        for k in range(1, z):
            add(k, d, d, k + 2)
        for k in range(z - 1, 0, -1):
            add(z - k, d, 1 if k == z - 1 else d, z - k + 1)
        for k in range(1, z):
            add(max(k - 1, z - k - 1) + 1, d, 1 if k == z - 1 else d, z)
    return ops


zmax = int(deg.max())
load = np.zeros(zmax + 1)
rounds = {}
for z in sorted((int(v) for v in deg), reverse=True):
    if z < 2:
        continue
    ops = node_ops(z)
    w = np.zeros(z + 1)
    for l, xs in ops.items():
        w[l] = sum(xs)
    best = min(range(zmax - z + 1), key=lambda o: float((load[1 + o:z + 1 + o] * w[1:]).sum()))
    load[1 + best:z + 1 + best] += w[1:]
    for l, xs in ops.items():
        rounds.setdefault(l + best, []).extend(xs)
tot_ideal = tot_ms = 0.0
for r in sorted(rounds):
    xs = sorted(rounds[r], reverse=True)
    h = [0.0] * 148
    for x in xs:
        heapq.heappush(h, heapq.heappop(h) + x)
    ms, ideal = max(h), sum(xs) / 148
    tot_ideal += ideal
    tot_ms += ms
    print(f"round {r}: {len(xs)} ops, max X {xs[0]}, makespan/ideal {ms / ideal:.2f}, longest CTA / ideal {xs[0] / ideal:.2f}")
print(f"overall efficiency of the QR launches (one stream, one launch per site): {tot_ideal / tot_ms:.2f}")
