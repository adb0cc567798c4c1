#!/usr/bin/env python
"""bench.py -- BP edge-updates/sec, FP64, T=50, bond dim 20 (BASELINE.json metric), Glauber on an Erdos-Renyi
graph c=4 (configs[2] shape) with N scaled so that one step (= one Jacobi BP iteration over every node of the
synthetic graph) fits the time budget.  Weak scaling: --nodes-per-gpu nodes per rank, edges cut by a contiguous
node partition, one halo exchange of cut-edge messages per step.

  python bench.py --gpus N --steps K --warmup W            # this framework (CUDA engine through the C-ABI)
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores (oracle port)

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definition of every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "BP edge-updates/sec, FP64, T=50 bond dim 20"
UNIT = "edge-updates/s"


def make_workload(ntot, seed=1):
    import networkx as nx
    G = nx.fast_gnp_random_graph(ntot, 4.0 / ntot, seed=seed)
    return ntot, [(int(a), int(b)) for a, b in G.edges()]


def xweight(z):
    """sum over the heavy ops (both operands of full bond) of X = nstates*q, Glauber: nstates(l) = l+1, q = 2."""
    w = 0
    for k in range(1, z):  # prefix p[k] -> nstates(k+1)
        w += 2 * (k + 2)
    for k in range(1, z - 1):  # suffix s[k], k <= z-2 -> nstates(z-k)
        w += 2 * (z - k + 1)
    for k in range(1, z - 1):  # dest[k], k <= z-2 -> nstates(z-1)
        w += 2 * z
    return w


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm on a bounded sample (one node update)
# ------------------------------------------------------------------------------------------------
def cpu_sample(T, d, z, seed=0):
    """one node update (src/recursive_bp_factor.jl:146-165) of a degree-z Glauber node whose incoming messages have
    the steady-state bond profile min(4^t, 4^(L-t), d); returns seconds.  Uses every host core through the
    threaded LAPACK/BLAS that numpy links (the reference threads over nodes instead; same silicon)."""
    from oracle import factors as OF, mpbp as O, tt as OT
    L = T + 1
    rng = np.random.default_rng(seed)
    und = [(0, k) for k in range(1, z + 1)]
    g = O.BiDiGraph(z + 1, und)
    w = [[OF.HomogeneousGlauberFactor(0.5, 0.1, 1.0)] * L for _ in range(z + 1)]
    phi = [[np.array([0.2, 0.8]) if t == 0 else np.ones(2) for t in range(L)] for _ in range(z + 1)]
    bp = O.MPBP(g, w, [2] * (z + 1), T, phi=phi)
    bonds = [min(4 ** t, 4 ** (L - t), d) for t in range(L + 1)]
    for e in range(g.ne):
        bp.mu[e] = OT.rand_tt(bonds, 2, 2, rng=rng)
    t0 = time.perf_counter()
    O.onebpiter_recursive(bp, 0, OT.TruncBond(d))
    return time.perf_counter() - t0


def cpu_baseline(T, d, degs, z_sample=2):
    t = cpu_sample(T, d, z_sample)
    edges = float(np.sum(degs))
    w_graph = float(sum(xweight(int(z)) for z in degs)) / max(edges, 1)
    w_sample = xweight(z_sample) / z_sample
    val = (z_sample / t) * (w_sample / w_graph)
    return dict(value=val, unit=UNIT, cores=os.cpu_count(), kind="port",
                sample=f"one oracle node update (degree {z_sample}, T={T}, TruncBond({d}), random full-bond incoming messages): "
                       f"{t:.1f} s for {z_sample} edge updates, extrapolated to the graph's degree sequence by the "
                       f"sum-of-X weight of the heavy ops ({w_sample:.2f} vs {w_graph:.2f} per edge)"), t


def run_reference(args, rank, world):
    if rank != 0:
        return
    n, und = make_workload(args.nodes_per_gpu * world)
    degs = np.bincount(np.array(und).ravel(), minlength=n)
    for _ in range(min(args.warmup, 1)):
        cpu_sample(min(args.T, 4), min(args.d, 4), 2)  # warm the BLAS threads / page in numpy
    ts, vals = [], []
    for _ in range(args.steps):
        cb, t = cpu_baseline(args.T, args.d, degs)
        ts.append(t)
        vals.append(cb["value"])
    cb["value"] = float(np.mean(vals))
    line = dict(metric=METRIC, value=cb["value"], unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * float(np.mean(ts)), higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                data="synthetic", impl="reference",
                config=dict(workload=f"glauber_er_c4 T={args.T} TruncBond({args.d}) (BASELINE configs[2] shape), bounded sample per step",
                            nodes_per_gpu=args.nodes_per_gpu),
                cpu_baseline=cb, e2e=dict(value=cb["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "250"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 7 for k in range(4) if r[3 + k].lower().startswith("active")})
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    power_w_max=max(pw) if pw else None, samples=len(sm))


def run_gpu(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import mpbp_b200 as M
    from mpbp_b200 import _lib
    from mpbp_b200.dist import CudaBackend, DistMPBP, LocalProblem, partition_balanced
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    T, d = args.T, args.d
    ntot, und = make_workload(args.nodes_per_gpu * world)
    owner = partition_balanced(ntot, und, world)
    lp = LocalProblem(ntot, und, owner, rank)
    g = M.IndexedBiDiGraph(len(lp.nodes), lp.local_und)
    lp.build_exchange(g.src, g.dst, world)
    w = [[M.HomogeneousGlauberFactor(0.5, 0.1, 1.0)] * (T + 1)] * g.N
    # pinned host copy of the reweightings: the e2e step uploads it every step
    nphi = g.N * (T + 1) * 2
    phi_host = torch.empty(nphi, dtype=torch.float64).pin_memory()
    phi_np = phi_host.numpy().reshape(g.N, T + 1, 2)
    phi_np[:] = 1.0
    phi_np[:, 0, :] = [0.2, 0.8]
    phi = [[phi_np[i, t] for t in range(T + 1)] for i in range(g.N)]
    bp = M.mpbp(g, w, [2] * g.N, T, phi=phi, dmax=d, device=local_rank)
    stream = torch.cuda.Stream()
    bp.set_stream(stream.cuda_stream)
    if args.arena_gb > 0:
        bp.set_option("arena_gb", args.arena_gb)
    for kv in args.set:  # engine tuning knobs (mpbp_set_option), e.g. --set level_balance=0
        k, v = kv.split("=")
        bp.set_option(k, float(v))
    backend = CudaBackend(bp, lp.owned_local, M.TruncBond(d))
    drv = DistMPBP(lp, backend, dist if world > 1 else None, device=f"cuda:{local_rank}")
    degs_owned = np.array([g.degree(int(i)) for i in lp.owned_local])
    edges_local = int(degs_owned.sum())

    def step():
        with torch.cuda.stream(stream):
            drv.iterate(1)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peak = np.zeros(1)
    _lib.check(_lib.lib().mpbp_measure_fp64_peak(local_rank, peak.ctypes.data_as(_lib.c_dp)))
    for _ in range(args.warmup):
        step()
    barrier()
    bp.counters(reset=True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(stream):
        e0.record()
    for _ in range(args.steps):
        step()
    with torch.cuda.stream(stream):
        e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = int(bp.counters(reset=True)["launches"])
    # ---- one extra PROFILED step for the roofline: single-stream launches so that the CUDA-event duration of every
    # kernel family is its own (the timed steps above run op groups on 4 concurrent streams, where durations overlap)
    bp.set_option("nstreams", 1)
    bp.set_option("profile", 1)
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(stream):
        pe0.record()
    step()
    with torch.cuda.stream(stream):
        pe1.record()
    barrier()
    prof_ms = pe0.elapsed_time(pe1)
    ctr = bp.counters(reset=True)
    fam = bp.kernel_times(reset=True)
    bp.set_option("profile", 0)
    bp.set_option("nstreams", 4)
    # max over ranks of the device time, sum over ranks of the units
    tt = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local_rank}")
    ee = torch.tensor([float(edges_local)], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(ee, op=dist.ReduceOp.SUM)
    ms = float(tt.item())
    edges_total = float(ee.item())
    ms_per_step = ms / args.steps
    value = edges_total / (ms_per_step / 1e3)
    # ---- e2e: same step through the public API with host buffers (H2D of phi, D2H of beliefs + f) ----
    e2e_steps = max(1, min(args.steps, 2))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        with torch.cuda.stream(stream):
            bp.upload_phi(phi_host.numpy())  # host -> device from the pinned buffer
            drv.iterate(1)
            b = M.beliefs(bp)
            f = M.bethe_free_energy(bp)
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    t2 = torch.tensor([e2e_s], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_s = float(t2.item())
    h2d = int(nphi * 8)
    d2h = int(nphi * 8 + g.N * 8)
    if rank == 0:
        qr_tf = ctr["qr_flops"] / (ctr["qr_ms"] * 1e-3) / 1e12 if ctr["qr_ms"] > 0 else 0.0
        traffic, traffic_note = None, None
        tpath = os.path.join(ROOT, "profiles", "qr_traffic.json")
        if os.path.exists(tpath):
            try:
                tj = json.load(open(tpath))
                traffic = tj.get("dram_bytes_per_launch")
                traffic_note = (f"ncu --set full capture of ONE isolated launch ({tj.get('launch')}; {tj.get('kernel')}; algorithmic bytes of that launch "
                                f"{tj.get('algorithmic_bytes_per_launch')}); the H=64 variant used for D >= 200 halves the R re-reads; not re-captured")
            except Exception:
                traffic = None
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                    config=dict(workload=f"glauber_er_c4 T={T} TruncBond({d}) (BASELINE configs[2] shape: nx.fast_gnp_random_graph(N, 4/N, seed=1), "
                                         f"HomogeneousGlauberFactor(J=0.5,h=0.1,beta=1), m0=-0.6), N={ntot} nodes, {int(edges_total)} directed edges; "
                                         "configs[2]'s N=1e5 does not fit one GPU (261 GB of messages) nor the time budget",
                                nodes_per_gpu=args.nodes_per_gpu, schedule="parallel (Jacobi), one halo exchange per step",
                                l2="working set per step >> L2 (126 MB): every heavy op streams ~70 MB of scratch",
                                parallelism=f"node partition x{world} (cost-balanced by degree)"),
                    clocks=clocks,
                    e2e=dict(value=edges_total / e2e_s, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h),
                    gpu_launches=launches,
                    roofline=dict(kernel="k_qr_ft (flat-tree DMMA Q-less QR of the bond-D sweep, csrc/qr_ft.cuh)", bound="tensor", achieved=qr_tf, peak=float(peak[0]),
                                  unit="TFLOP/s", frac=qr_tf / float(peak[0]) if peak[0] > 0 else None, traffic=traffic, traffic_note=traffic_note,
                                  peak_source="FP64 DMMA (mma.sync m8n8k4 f64) measured live by mpbp_measure_fp64_peak; MEASURED_PEAKS.json has no FP64 entry",
                                  algorithmic_flops=ctr["qr_flops"], kernel_ms=ctr["qr_ms"], share_of_step=ctr["qr_ms"] / prof_ms if prof_ms > 0 else None,
                                  measured_on=f"one extra profiled step right after the timed region (single-stream launches, {prof_ms:.0f} ms); the timed steps use 4 concurrent streams",
                                  heavy_ops=int(ctr["ops"]), subspace_svd=dict(calls=int(ctr["svd_calls"]), iters=int(ctr["svd_iters"]), unconverged=int(ctr["svd_unconverged"])), kernel_family_ms={k: round(v, 1) for k, v in fam.items()}))
        if world == 1 and not args.no_cpu:
            degs = np.array([g.degree(i) for i in range(g.N)])
            line["cpu_baseline"], _ = cpu_baseline(T, d, degs)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nodes-per-gpu", type=int, default=int(os.environ.get("MPBP_BENCH_NODES", 384)))
    ap.add_argument("--T", type=int, default=50)
    ap.add_argument("--d", type=int, default=20)
    ap.add_argument("--arena-gb", type=float, default=0.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--set", action="append", default=[], metavar="OPTION=VALUE")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        import __graft_entry__ as G
        G.build()
        run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
