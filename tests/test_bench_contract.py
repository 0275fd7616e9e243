"""bench.py's output contract: exactly one JSON line on stdout with the keys the driver reads (CPU arm here; the GPU arm is
checked with -m gpu on a B200)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "cpu_baseline", "gpu_launches"}


def run_bench(*argv, env=None):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *argv], capture_output=True, text=True, timeout=900,
                       env={**os.environ, **(env or {})})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, f"stdout must hold exactly one line, got {len(lines)}"
    return json.loads(lines[0])


def test_reference_arm_prints_one_json_line():
    line = run_bench("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert BASE_KEYS <= set(line) and line["impl"] == "reference"
    assert line["metric"] == "sample_voxels_per_s" and line["unit"] == "sample-voxels/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and "model" not in line["config"]


def test_reference_arm_other_ranks_stay_silent():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=300, env={**os.environ, "RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.gpu
def test_gpu_arm_prints_one_json_line():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    line = run_bench("--steps", "3", "--warmup", "3", "--images-per-step", "2", "--e2e-images", "1", "--e2e-steps", "1", "--cpu-images", "1", "--no-configs")
    assert BASE_KEYS | {"roofline", "clocks"} <= set(line)
    rf = line["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and 0 < rf["frac"] < 1.2 and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert line["gpu_launches"] == 3 and line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] > 0
    assert line["value"] > line["e2e"]["value"] > line["cpu_baseline"]["value"]
