"""bench.py's contract on the CPU box: the reference arm prints one well-formed JSON line, and the CUDA arm refuses to run
without a device (there is no CPU fallback to fall back to)."""
import json
import os
import subprocess
import sys

import pytest
import torch

import rendering_b200 as rb
from helpers import HAVE_REF_DRIVER

BENCH = os.path.join(rb.REPO_ROOT, "bench.py")


def test_reference_arm_prints_the_contract_line():
    if not HAVE_REF_DRIVER:
        pytest.skip("oracle/_ref/ref_driver not built")
    out = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--steps", "2", "--warmup", "1", "--scene", "cfg1_simple_shapes_256"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mrays/s" and d["unit"] == "Mrays/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["config"]["rays_per_frame"] == 572758 and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    out = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=120, env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_cuda_arm_refuses_to_run_without_a_device():
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    out = subprocess.run([sys.executable, BENCH, "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
