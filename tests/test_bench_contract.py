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
