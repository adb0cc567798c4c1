"""bench.py contract checks that need no GPU: the reference arm (oracle port on the host cores) prints ONE JSON line
with the keys the driver reads, only rank 0 does any work, and our arm refuses to run without a CUDA device."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TINY = ["--steps", "1", "--warmup", "0", "--T", "4", "--d", "4", "--nodes-per-gpu", "12"]


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, env=e, capture_output=True, text=True, timeout=600)


def test_reference_arm_prints_one_contract_line():
    r = _run(["--impl", "reference"] + TINY)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "impl", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "edge-updates/s" and d["dtype"] == "f64" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["gpu_launches"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_do_no_work():
    r = _run(["--impl", "reference", "--gpus", "2"] + TINY, env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_our_arm_needs_a_cuda_device():
    import torch
    if torch.cuda.is_available():
        return  # covered by the GPU runs
    r = _run(TINY + ["--no-cpu"])
    assert r.returncode != 0
    assert "CUDA" in (r.stderr + r.stdout)
