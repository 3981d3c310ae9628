"""CPU suite: the parts of bench.py's contract that can be checked without a GPU — the reference arm (the CPU oracle on the host
cores) prints one JSON line with the keys the driver reads, and the native arm refuses to run without a CUDA device."""
import json
import os
import subprocess
import sys

import pytest

import conftest

BENCH = os.path.join(conftest.ROOT, "bench.py")


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--steps", "1", "--warmup", "3", "--scans-per-step", "8"],
                         capture_output=True, text=True, timeout=600, cwd=conftest.ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "scans/s" and line["value"] > 0 and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--gpus", "2", "--steps", "1"], capture_output=True, text=True,
                         timeout=120, cwd=conftest.ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


@pytest.mark.skipif(conftest.has_gpu(), reason="checks the no-GPU failure path")
def test_native_arm_needs_a_gpu():
    out = subprocess.run([sys.executable, BENCH, "--steps", "1"], capture_output=True, text=True, timeout=300, cwd=conftest.ROOT)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
