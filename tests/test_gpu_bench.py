"""GPU test of bench.py's JSON contract on a small workload (the real numbers come from the default run)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_line_contract():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--grid", "64", "--Re", "253", "--steps", "2",
                        "--warmup", "3", "--block", "120", "--no-cpu-baseline"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in j, k
    assert j["metric"] == "MLUPS" and j["unit"] == "MLUPS" and j["dtype"] == "f64" and j["data"] == "synthetic"
    assert j["n_gpus"] == 1 and j["steps"] == 2 and j["warmup"] == 3 and j["vs_baseline"] is None
    assert j["value"] > 1000 and j["ms_per_step"] > 0 and "workload" in j["config"] and "model" not in j["config"]
    rf = j["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12
    assert rf["algorithmic_bytes_per_launch"] == 304.0 * 64 ** 3
    e = j["e2e"]
    assert e["h2d_bytes_per_step"] == 19 * 8 * 64 ** 3 and e["d2h_bytes_per_step"] == 8 + 5 * 8 * 64 ** 3
    assert 0 < e["value"] < j["value"] * 1.05
    assert j["gpu_launches"] >= 2 * 120 and j["parity_probe"]["sha256_equals_reference"] is True
    assert set(j["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
