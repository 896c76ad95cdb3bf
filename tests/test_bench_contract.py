"""The bench contract on the CPU: the reference arm runs without a GPU and prints ONE JSON line with the keys the driver
reads; the argument parser defaults finish within minutes; the workload description is the one BASELINE.json names."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    rec = json.loads(lines[0])
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert rec["impl"] == "reference" and rec["n_gpus"] == 1 and rec["steps"] == 1
    assert rec["metric"] == "team_head_fwd_bwd_samples_per_sec" and rec["unit"] == "samples/s" and rec["higher_is_better"] is True
    assert "samples/sec" in base["metric"] and not base["published"]                 # vs_baseline stays null: nothing published
    assert rec["value"] > 0 and abs(rec["value"] - 1024 / (rec["ms_per_step"] * 1e-3)) < 1e-6 * rec["value"]
    assert rec["vs_baseline"] is None and rec["gpu_launches"] == 0
    cb = rec["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == rec["value"] and cb["sample"]
    e2e = rec["e2e"]
    assert e2e["value"] == rec["value"] and e2e["unit"] == rec["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert rec["config"]["tasks"] == 10 and rec["config"]["batch_per_gpu"] == 1024 and "model" not in rec["config"]


def test_reference_arm_other_ranks_exit_without_work():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    assert not [l for l in out.stdout.splitlines() if l.startswith("{")]


def test_defaults_and_workload_description():
    import bench
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        a = bench.parse()
    finally:
        sys.argv = argv
    assert a.gpus == 1 and a.warmup >= 3 and a.steps >= 1 and a.impl != "reference"
    cfg = bench.workload_config(10, 1024, 8)
    assert cfg["global_batch"] == 8192 and cfg["parallelism"] == "dp8" and "T=10" in cfg["workload"]
