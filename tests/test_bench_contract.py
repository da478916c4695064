"""The JSON line `bench.py` prints is a contract with the driver (metric / value / e2e / roofline /
cpu_baseline / clocks / gpu_launches ...).  The reference arm runs on the host cores (the oracle port: the one
place besides the tests where bench.py executes `oracle/`), so its line is checked here on CPU; the GPU arm's
line is checked on the GPU box with a minimal run."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMMON = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
          "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline")


def run_bench(*args, timeout=600):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                         timeout=timeout, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1, "bench.py must print ONE JSON line"
    return json.loads(lines[0])


def test_reference_arm_line():
    d = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0")
    for k in COMMON + ("impl",):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"].startswith("transient F-stat") and d["unit"] == "cells/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    assert d["config"]["name"] == "exp120" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


@pytest.mark.gpu
def test_gpu_arm_line():
    d = run_bench("--steps", "2", "--warmup", "3", "--no-secondary", "--no-strong")
    for k in COMMON + ("clocks", "gpu_launches", "roofline", "stage_ms"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["higher_is_better"] is True
    assert d["config"]["name"] == "exp120" and "L2 flushed" in d["config"]["l2"]
    assert d["gpu_launches"] > 0 and d["value"] > 1e9
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and 0.0 < r["frac"] <= 1.2
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["walk_kernel"]["bound"] == "hbm" and 0.0 < r["walk_kernel"]["frac"] < 1.0
    e = d["e2e"]
    assert e["value"] > 1e9 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] > 0
    c = d["clocks"]
    assert "reasons" in c and not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"]))
