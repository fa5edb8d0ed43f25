"""bench.py output contract on CPU: the reference arm (the only arm that runs without a GPU) prints exactly one JSON
line on stdout with the keys the driver reads."""
import json
import os
import subprocess
import sys

from conftest import REPO


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--n0", "16", "--ref-cells", "60"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "cells/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("overlay-cell stiffness assembly cells/s")
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["dtype"] == "f64" and d["data"] == "synthetic"
    staged = os.path.isdir(os.path.join(REPO, "baseline", "_ref", "MOFEM"))      # the unmodified reference when staged, else the oracle port
    assert d["cpu_baseline"]["kind"] == ("reference" if staged else "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_silently():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0", "--n0", "16", "--ref-cells", "60"], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_gpu_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--steps", "1", "--warmup", "0", "--n0", "16"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode != 0 and out.stdout.strip() == ""


def test_scaling_decomposition_only_for_the_default_weak_series():
    """The extra object of the multi-GPU lines exists only where its committed 1-GPU side applies (default weak workload)."""
    import argparse
    import bench
    a = argparse.Namespace(strong=False, family="quad", n0=664)
    assert bench.scaling_decomposition(a, 1, 6558, 10000.0) is None
    d = bench.scaling_decomposition(a, 8, 13036, 9043.0)
    assert d is not None and abs(d["iteration_growth"] - 13036 / d["pcg_iterations_1gpu"]) < 1e-12
    assert 0.0 < d["per_iteration_efficiency"] < 1.2
    assert bench.scaling_decomposition(argparse.Namespace(strong=True, family="quad", n0=664), 8, 1, 1.0) is None
    assert bench.scaling_decomposition(argparse.Namespace(strong=False, family="m3s", n0=664), 8, 1, 1.0) is None
