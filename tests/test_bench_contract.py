"""bench.py contract checks that need no GPU: the reference arm prints ONE JSON line with the agreed keys, and the
product arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=600, env=e, cwd=ROOT)


def test_reference_arm_json_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--res", "16")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert REQUIRED <= set(d), REQUIRED - set(d)
    assert d["impl"] == "reference" and d["unit"] == "images/sec" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_stay_silent():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--res", "16", "--gpus", "2",
             env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_needs_cuda():
    if torch.cuda.is_available():
        return
    r = _run("--steps", "1", "--warmup", "0")
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)


def test_reference_arm_time_budget_stops_early_and_says_so():
    """On a slow or shared host the CPU arm stops after the step that exhausts its time budget (at least one timed step)
    and reports the number of steps it really timed."""
    r = _run("--impl", "reference", "--steps", "6", "--warmup", "1", "--res", "16", env={"SG_REFERENCE_BUDGET_S": "0.0001"})
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    assert d["steps"] == 1 and d["warmup"] == 1 and d["value"] > 0
    assert "of 6 requested" in d["cpu_baseline"]["sample"]
    r = _run("--impl", "reference", "--steps", "3", "--warmup", "1", "--res", "16", env={"SG_REFERENCE_BUDGET_S": "600"})
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    assert d["steps"] == 3
