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


def test_reference_arm_never_maps_the_cuda_library():
    """The CPU arm loads the generator/parameter library and the oracle only (it used to import the product .so)."""
    code = (
        "import sys, runpy\n"
        f"sys.argv=[{BENCH!r},'--impl','reference','--steps','1','--warmup','0','--scans-per-step','4','--workers','2']\n"
        "try:\n    runpy.run_path(sys.argv[0], run_name='__main__')\nexcept SystemExit:\n    pass\n"
        "print('MAPS', sorted({l.split()[-1].split('/')[-1] for l in open('/proc/self/maps') if 'scvod' in l and '.so' in l}))\n")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=conftest.ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    maps = [l for l in out.stdout.splitlines() if l.startswith("MAPS")][-1]
    assert "libscvod_synth.so" in maps and "libscvod_oracle.so" in maps and "libscvod_b200.so" not in maps
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["config"]["scans_per_chunk"] == 4 and line["config"]["chunks_per_step"] == 2  # the GPU arm's decomposition
    assert "single_chain" in line["cpu_baseline"]


def test_oracle_chunked_run_equals_the_sequence_run(pkg, kitti_params):
    """orc_run_chunks (independent chunks, frames dropped once tracked) gives the labels of orc_run_sequence on each chunk."""
    import numpy as np

    S, nch = 5, 2
    chunks = [[pkg.synth_scan(conftest.SEED + 9, 100 * c + k, rings=16, cols=450) for k in range(S)] for c in range(nch)]
    scans = [s for ch in chunks for s, _ in ch]
    poses = np.stack([p for ch in chunks for _, p in ch])
    off = np.zeros(len(scans) + 1, np.int64)
    off[1:] = np.cumsum([len(s) for s in scans])
    orc = conftest.Oracle(kitti_params)
    _, lab = orc.run_chunks(np.concatenate(scans), off, poses, S, 3, nthreads=2, want_labels=True)  # 3 > nch: chunk 0 runs twice
    for c in range(nch):
        _, ref, o2 = orc.run_sequence(scans[c * S:(c + 1) * S], poses[c * S:(c + 1) * S], nthreads=1)
        assert np.array_equal(lab[off[c * S]:off[(c + 1) * S]], ref)
    orc.close()


def test_numa_binding_degrades_to_not_pinned_without_topology():
    """bench.py binds a rank to the cores of its GPU's NUMA node only when the box exposes that topology; anywhere else
    (no nvidia-smi, numa_node = -1, one node) it must leave the affinity alone and say so."""
    import importlib.util
    import os

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(conftest.ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    before = os.sched_getaffinity(0)
    note = mod.bind_to_gpu_numa_node(0, 2)
    assert note.startswith("not pinned")
    assert os.sched_getaffinity(0) == before
