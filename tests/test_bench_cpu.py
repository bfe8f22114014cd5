"""CPU tests of bench.py's reference arm (the CPU implementation timed on the host cores) and of the
JSON contract; the GPU arm itself is exercised on the B200 box."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_ref(*extra, env=None):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        *extra], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout.strip().splitlines()


def test_reference_arm_uses_the_real_reference_binary():
    lines = run_ref("--grid", "24", "--Re", "100")
    j = json.loads(lines[-1])
    assert j["impl"] == "reference" and j["metric"] == "MLUPS" and j["unit"] == "MLUPS" and j["higher_is_better"] is True
    assert j["cpu_baseline"]["kind"] == "reference" and j["cpu_baseline"]["cores"] >= 1 and j["value"] > 0
    assert j["e2e"] == {"value": j["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert j["config"]["grid"] == 24 and j["dtype"] == "f64" and j["vs_baseline"] is None


def test_reference_arm_falls_back_to_the_oracle_port_for_unbuilt_grids():
    j = json.loads(run_ref("--grid", "20", "--Re", "77")[-1])
    assert j["cpu_baseline"]["kind"] == "port" and j["value"] > 0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    assert run_ref("--gpus", "2", "--grid", "20", env=env) == []


def test_both_arms_use_the_same_workload_label():
    """The driver compares config.workload of the two arms (same_config): one function builds both."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    j = json.loads(run_ref("--grid", "24", "--Re", "100")[-1])
    assert j["config"]["workload"] == bench.workload_label(24, 100, 1)
    j = json.loads(run_ref("--grid", "20", "--Re", "77", "--gpus", "4")[-1])
    assert j["config"]["workload"] == bench.workload_label(20, 77, 4) and j["n_gpus"] == 4
    assert j["config"]["cpu_code"] == "port" and j["cpu_baseline"]["kind"] == "port"
