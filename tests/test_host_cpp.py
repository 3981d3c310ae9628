"""C++ host layer (host/): the SSC / PatchWork / Session / Utility class surface of the reference over the C-ABI.

CPU: it builds, the reference's OWN src/main.cpp compiles against it unchanged (when the checkout is present), the
parameter server reads the launch file's YAML, and the node fails loudly without a GPU.
GPU: the node (batched segDF) and the stage-by-stage driver (process/segment/recognize/tracking, one call at a
time as in the reference's segDF body) give the oracle's per-point classes.
"""
import os
import struct
import subprocess

import numpy as np
import pytest

import conftest

HOST = os.path.join(conftest.ROOT, "host")
BUILD = os.path.join(HOST, "_build")
YAML = os.path.join(HOST, "config", "synthetic.yaml")
REF_MAIN = "/root/reference/src/main.cpp"


@pytest.fixture(scope="module")
def host_build():
    subprocess.check_call(["make", "-C", HOST], stdout=subprocess.DEVNULL)
    return BUILD


def test_host_layer_builds(host_build):
    for f in ("libufo_host.so", "ufo_ufo", "ufo_stagewise"):
        assert os.path.exists(os.path.join(host_build, f))
    out = subprocess.run(["nm", "-DC", os.path.join(host_build, "libufo_host.so")], capture_output=True, text=True).stdout
    for sym in ("SSC::segDF()", "SSC::process(", "SSC::segment()", "SSC::recognize(Frame&)", "SSC::tracking(Frame&, Frame&, PointXYZIRPYT, PointXYZIRPYT)",
                "SSC::extractGroudByPatchWork(", "SSC::makeApriVec(", "SSC::makeHashCloud(", "SSC::findVoxelNeighbors(", "SSC::getPose()",
                "SSC::getCloud()", "SSC::intialization(", "Session::Session()", "Session::getPose(", "Session::getCloudSeg(", "Session::getReloInfo("):
        assert sym in out, sym


def test_loader_voxel_grid(host_build):
    """The 0.08 m VoxelGrid of the KITTI loader (reference src/ssc.cpp:1108-1111, restated in host/include/voxel_grid.h):
    one output per occupied leaf, in ascending leaf index, each the mean of the leaf's points (float sums: 1e-4)."""
    import ctypes

    lib = ctypes.CDLL(os.path.join(host_build, "libufo_host.so"))
    rng = np.random.default_rng(11)
    pts = np.concatenate([rng.uniform(-20, 20, (20000, 3)), rng.uniform(0, 255, (20000, 1))], axis=1).astype(np.float32)
    pts[:, 2] = rng.uniform(-2, 4, 20000).astype(np.float32)
    pts[:5000, :3] = (pts[:5000, :3] / 40).astype(np.float32)  # a dense blob: many points per leaf
    out = np.zeros_like(pts)
    n_out = ctypes.c_int(0)
    leaf = np.float32(0.08)
    rc = lib.ufo_voxel_grid(pts.ctypes.data_as(ctypes.c_void_p), len(pts), ctypes.c_float(float(leaf)), out.ctypes.data_as(ctypes.c_void_p),
                            ctypes.byref(n_out))
    assert rc == 0
    out = out[: n_out.value]
    inv = np.float32(1.0) / leaf
    ijk = np.floor(pts[:, :3] * inv).astype(np.int64)
    mn = ijk.min(axis=0)
    div = ijk.max(axis=0) - mn + 1
    idx = (ijk[:, 0] - mn[0]) + (ijk[:, 1] - mn[1]) * div[0] + (ijk[:, 2] - mn[2]) * div[0] * div[1]
    uniq, inverse, counts = np.unique(idx, return_inverse=True, return_counts=True)
    assert len(out) == len(uniq) and counts.max() > 10
    sums = np.zeros((len(uniq), 4), np.float64)
    np.add.at(sums, inverse, pts.astype(np.float64))
    assert np.allclose(out, sums / counts[:, None], rtol=0, atol=1e-4 * max(1.0, float(np.abs(pts).max())) / 100)
    # an empty cloud stays empty, a single point is returned as it is
    assert lib.ufo_voxel_grid(pts.ctypes.data_as(ctypes.c_void_p), 0, ctypes.c_float(0.08), out.ctypes.data_as(ctypes.c_void_p), ctypes.byref(n_out)) == 0
    assert n_out.value == 0
    one = np.array([[1.0, 2.0, 3.0, 4.0]], np.float32)
    o1 = np.zeros_like(one)
    lib.ufo_voxel_grid(one.ctypes.data_as(ctypes.c_void_p), 1, ctypes.c_float(0.08), o1.ctypes.data_as(ctypes.c_void_p), ctypes.byref(n_out))
    assert n_out.value == 1 and np.array_equal(o1, one)


@pytest.mark.skipif(not os.path.exists(REF_MAIN), reason="reference checkout not present (GPU box)")
def test_reference_main_cpp_compiles_unchanged(host_build):
    subprocess.check_call(["make", "-C", HOST, "ref_main"], stdout=subprocess.DEVNULL)
    assert os.path.exists(os.path.join(host_build, "ufo_ufo_refmain"))


@pytest.mark.skipif(conftest.has_gpu(), reason="checks the no-GPU failure path")
def test_node_reads_yaml_and_fails_loudly_without_gpu(host_build):
    r = subprocess.run([os.path.join(host_build, "ufo_ufo"), f"_params:={YAML}"], capture_output=True, text=True)
    assert r.returncode == 2
    assert "range_num: 72" in r.stdout and "sector_num: 300" in r.stdout and "azimuth_num: 60" in r.stdout  # src/ssc.cpp:36-39 on the YAML values
    assert "synth:8:64:1800" in r.stdout
    assert "no CPU fallback" in r.stderr
    # private-parameter override, as rosrun would pass it
    r = subprocess.run([os.path.join(host_build, "ufo_ufo"), f"_params:={YAML}", "_ssc/max_dis_:=40.0", "_ssc/min_dis_:=0.8"], capture_output=True, text=True)
    assert "range_num: 98" in r.stdout  # the parkinglot grid


def read_dump(path):
    with open(path, "rb") as f:
        nf = struct.unpack("i", f.read(4))[0]
        out = []
        for _ in range(nf):
            n = struct.unpack("i", f.read(4))[0]
            out.append(np.frombuffer(f.read(n), np.uint8))
    return out


def oracle_labels(pkg, nscans, rings, cols):
    o = conftest.Oracle(pkg.semantickitti_params())
    scans, poses = zip(*[pkg.synth_scan(conftest.SEED, k, rings=rings, cols=cols) for k in range(nscans)])
    for s in scans:
        o.push_scan(s)
    o.track(np.stack(poses))
    lab = [o.labels(f) for f in range(nscans)]
    cl = [o.clusters(f) for f in range(nscans)]
    cnt = [o.counts(f) for f in range(nscans)]
    o.close()
    return scans, lab, cl, cnt


@pytest.mark.gpu
def test_node_segdf_labels_match_oracle(host_build, pkg, tmp_path):
    dump = str(tmp_path / "labels.bin")
    env = dict(os.environ, UFO_DUMP_LABELS=dump)
    r = subprocess.run([os.path.join(host_build, "ufo_ufo"), f"_params:={YAML}"], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    got = read_dump(dump)
    scans, lab, _, _ = oracle_labels(pkg, 8, 64, 1800)
    assert len(got) == 8
    for f in range(8):
        assert np.array_equal(got[f], lab[f]), f"frame {f}"
    assert sum(int((g == pkg.PT_DYNAMIC).sum()) for g in got) > 0


@pytest.mark.gpu
def test_stagewise_driver_matches_oracle(host_build, pkg, tmp_path):
    out = str(tmp_path / "stagewise.bin")
    r = subprocess.run([os.path.join(host_build, "ufo_stagewise"), f"_params:={YAML}", "_session/data_path_:=synth:5:32:900", out],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    scans, lab, cl, cnt = oracle_labels(pkg, 5, 32, 900)
    with open(out, "rb") as f:
        nf = struct.unpack("i", f.read(4))[0]
        assert nf == 5
        for k in range(nf):
            nd, ncl, ncar, nuse, nvox = struct.unpack("5i", f.read(20))
            dyn = np.frombuffer(f.read(12 * nd), np.float32).reshape(-1, 3)
            assert ncl == len(cl[k]["name"]) and ncar == int((cl[k]["type"] == 2).sum())
            assert nuse == cnt[k][3] and nvox == cnt[k][4]
            want = scans[k][lab[k] == pkg.PT_DYNAMIC][:, :3]
            assert len(dyn) == len(want)
            a = np.sort(dyn.view([("x", "f4"), ("y", "f4"), ("z", "f4")]).ravel(), order=("x", "y", "z"))
            b = np.sort(np.ascontiguousarray(want).view([("x", "f4"), ("y", "f4"), ("z", "f4")]).ravel(), order=("x", "y", "z"))
            assert np.array_equal(a, b), f"dynamic points of frame {k}"
