"""bench.py's reference arm runs without a GPU: its JSON line is checked here on a small box
(the contract of the driver: metric/unit/config of our arm, impl, cpu_baseline, e2e keys)."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

from oracle import oracle as O

ROOT = Path(__file__).resolve().parent.parent


def run_reference_arm(extra_env=None):
    env = dict(os.environ, **(extra_env or {}))
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--n", "32", "--steps", "2",
                        "--warmup", "1", "--ref-iters", "3"], capture_output=True, text=True, timeout=600, env=env,
                       cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [x for x in r.stdout.splitlines() if x.startswith("{")]
    assert len(lines) == 1, r.stdout
    return json.loads(lines[0])


def test_reference_arm_line():
    d = run_reference_arm()
    assert d["impl"] == "reference" and d["unit"] == "iterations/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("PCG iterations/sec") and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["gpu_launches"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"].startswith("box32 PCG+DIC")
    cb = d["cpu_baseline"]
    assert cb["value"] == d["value"] and cb["cores"] >= 1 and cb["sample"]
    assert cb["kind"] == ("reference" if O.ref_available() else "port")
    assert d["e2e"] == {"value": d["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
    if O.ref_par_available() and cb["cores"] > 1:
        assert "coupled" in cb["sample"]


def test_reference_arm_other_ranks_stay_silent():
    """under torchrun only rank 0 runs the reference arm; the others exit 0 without output"""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--n", "16", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""
