"""bench.py's reference arm (the oracle port on the host cores) prints the driver's JSON contract: one line, the same
metric / unit / config object as the GPU arm, `impl: reference`, a cpu_baseline describing the bounded sample and an
e2e block with zero copies."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("port", [False, True])
def test_reference_arm_json_contract(port):
    """`--impl reference`: the unmodified reference when it is reachable (build container), else the oracle port; forced to
    the port with RGM_BENCH_PORT=1 (what the GPU box runs).  N=1 keeps the test short; the default sample is B=1, N=16."""
    env = dict(os.environ, OMP_NUM_THREADS="1")  # what torchrun exports: the arm must still use every core
    if port:
        env["RGM_BENCH_PORT"] = "1"
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--ref-candidates", "1"], capture_output=True, text=True, timeout=900,
                         cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    sys.path.insert(0, ROOT)
    import bench

    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == bench.UNIT
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["config"] == bench.workload_config(bench.B_FULL, bench.N_FULL, 1)
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] == 1 and d["warmup"] == 0
    # ms_per_step is the measured sampled step; the scaled figure is what `value` inverts
    assert abs(d["ms_per_full_step_scaled"] * d["value"] / 1e3 - 1.0) < 1e-9
    assert d["ms_per_step"] < d["ms_per_full_step_scaled"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == ("port" if port or not os.path.isdir("/root/reference") else "reference")
    assert cb["cores"] == len(os.sched_getaffinity(0)) and cb["value"] == d["value"] and "B=1, N=1" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_workload_flops_and_configs():
    sys.path.insert(0, ROOT)
    import bench

    assert abs(bench.step_flops(64, 16) / 1.196e15 - 1) < 2e-3          # config 3
    assert abs(bench.step_flops(256, 0, "c2") / 60.8e12 - 1) < 2e-3     # config 2
    assert abs(bench.step_flops(1, 16, "c5") / 207.6e12 - 1) < 2e-3     # config 5, per batch element
    for cfg in ("c2", "c3", "c5"):
        c = bench.workload_config(*bench.CONFIG_DEFAULTS[cfg], 1, cfg)
        assert set(c) == {"workload", "global_batch", "candidates", "parallelism", "l2"}
    assert "candidate-sharded x8" in bench.workload_config(8, 64, 8, "c3", "weak", "candidates")["parallelism"]
    assert bench.workload_config(8, 16, 8, "c3", "strong")["global_batch"] == 64


def test_gpu_arm_refuses_to_run_without_cuda():
    import torch

    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
