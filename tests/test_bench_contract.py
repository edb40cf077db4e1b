"""bench.py contract checks that need no GPU: the reference arm (CPU oracle port) prints exactly one JSON line
with the keys the driver reads, and the product arm refuses to run without a CUDA device."""
import json
import os
import subprocess
import sys

import nsxlib as N

BENCH = os.path.join(N.ROOT, "bench.py")


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--mesh", "20,8", "--warmup", "0", "--steps", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "s per Newton step" and d["unit"] == "s" and d["higher_is_better"] is False
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "StationaryNSSolver -m 20,8" in d["config"]["workload"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--gpus", "2", "--mesh", "20,8"], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, BENCH, "--mesh", "20,8", "--warmup", "0", "--steps", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CPU path" in (r.stderr + r.stdout)


import pytest


@pytest.mark.gpu
def test_product_arm_line_has_the_contract_keys():
    """The product arm under driver-style arguments (small mesh so that it takes seconds): one JSON line with every key the contract
    names, and the budget guard lowers the step COUNT -- never the work in a step -- when the run would not fit."""
    r = subprocess.run([sys.executable, BENCH, "--gpus", "1", "--mesh", "40,14", "--steps", "2", "--warmup", "1", "--cpu-sample-s", "2", "--kernel-reps", "5"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "s per Newton step" and d["unit"] == "s" and d["higher_is_better"] is False and d["dtype"] == "f64"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["value"] > 0 and d["gpu_launches"] > 100
    assert "workload" in d["config"] and d["config"]["budget_guard"] is None
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["value"] >= d["value"]
    assert set(d["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"} and d["roofline"]["bound"] == "hbm"
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and d["cpu_baseline"]["kind"] == "port"
    # a budget that cannot hold the requested steps: fewer steps, said so
    r = subprocess.run([sys.executable, BENCH, "--gpus", "1", "--mesh", "40,14", "--steps", "20", "--warmup", "5", "--no-cpu", "--kernel-reps", "3", "--budget-s", "45"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.strip()][0])
    g = d["config"]["budget_guard"]
    assert g is None or (g["requested_steps"] == 20 and g["requested_warmup"] == 5 and 1 <= d["steps"] < 20 and d["warmup"] <= 5)
