"""bench.py's reference arm runs on host cores only, so its JSON contract is checked here (no GPU)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    pr = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                         "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert pr.returncode == 0, pr.stderr[-2000:]
    lines = [l for l in pr.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "fp64_cholesky_tflops" and d["unit"] == "TFLOP/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["dtype"] == "f64"
    assert d["config"]["N"] == 65536 and d["config"]["tile"] == 1024
    # the arm says which bounded sample it really timed (the reference program's N is fixed in its source)
    assert d["config"]["N_timed"] == 12000 and "BOUNDED SAMPLE" in d["config"]["workload"]
    assert d["tiled_single_worker"]["info"] == 0 and d["tiled_single_worker"]["backward_error"] <= 1e-13
    assert d["monolithic_N16384_all_cores"]["info"] == 0 and d["monolithic_N16384_all_cores"]["tflops"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_batched_leg():
    pr = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "batched",
                         "--steps", "1", "--warmup", "0", "--batch", "300"], capture_output=True, text=True,
                        timeout=600, cwd=ROOT)
    assert pr.returncode == 0, pr.stderr[-2000:]
    d = json.loads([l for l in pr.stdout.splitlines() if l.strip()][-1])
    assert d["impl"] == "reference" and d["metric"] == "fp64_batched_cholesky_tflops" and d["value"] > 0
    assert d["config"]["n"] == 256 and d["cpu_baseline"]["cores"] == 1


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    pr = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                        capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert pr.returncode == 0 and pr.stdout.strip() == ""


def test_sweep_writes_reference_csv_rows_even_when_the_driver_fails(tmp_path):
    """benchmark.c:257-285 appends a row for every run, with -1 metrics and the child's exit code when
    the driver fails; on this GPU-less box the driver fails loudly (no CPU path), which exercises that."""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("GPU present: the driver would succeed")
    csv = tmp_path / "results" / "bench.csv"
    pr = subprocess.run([sys.executable, "-m", "dense_linear_app_b200.bench_sweep", "--N", "64", "--NB", "16",
                         "--repeats", "1", "--sched", "lookahead", "--csv", str(csv)],
                        capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert pr.returncode == 0, pr.stderr[-1500:]
    rows = csv.read_text().splitlines()
    assert rows[0] == "timestamp,scheduler,mapping,ncpu,ngpu,N,NB,run_idx,ms,exit_code,gflops,rel_error"
    f = rows[1].split(",")
    assert f[1:8] == ["lookahead", "1_b200", "0", "1", "64", "16", "0"]
    assert int(f[9]) != 0 and float(f[10]) == -1.0 and float(f[11]) == -1.0
