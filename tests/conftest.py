"""Shared fixtures: the product package (ctypes over libscvod_b200.so) and the CPU oracle.

The oracle (oracle/_build/libscvod_oracle.so) is test infrastructure: it is only ever loaded here, in
__graft_entry__.smoke() and in bench.py's cpu_baseline / --impl reference legs.
"""
import ctypes
import importlib.util
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "dr-using-scv-od_b200")
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "libscvod_oracle.so")
SEED = 0x5C0D0000


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_package():
    name = "scvod_b200"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(PKG_DIR, "__init__.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_parallel():
    name = "scvod_b200_parallel"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(PKG_DIR, "parallel.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class Oracle:
    """ctypes wrapper of the CPU restatement (oracle/scvod_oracle.cpp)."""

    def __init__(self, params):
        if not os.path.exists(ORACLE_SO):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
        self.lib = ctypes.CDLL(ORACLE_SO)
        self.lib.orc_create.restype = ctypes.c_void_p
        self.lib.orc_run_sequence.restype = ctypes.c_double
        self.lib.orc_atan2f.restype = ctypes.c_float
        self.params = params
        self.h = ctypes.c_void_p(self.lib.orc_create(ctypes.byref(params)))
        self.sizes = []

    def close(self):
        if self.h:
            self.lib.orc_destroy(self.h)
            self.h = None

    def grid_dims(self):
        out = (ctypes.c_int * 4)()
        self.lib.orc_grid_dims(self.h, out)
        return list(out)

    def push_scan(self, xyzi):
        xyzi = np.ascontiguousarray(xyzi, np.float32)
        self.sizes.append(len(xyzi))
        return self.lib.orc_push_scan(self.h, _ptr(xyzi), len(xyzi))

    def track(self, poses):
        poses = np.ascontiguousarray(poses, np.float32).reshape(-1, 6)
        self.lib.orc_track(self.h, _ptr(poses), len(poses))

    def initialization(self, poses):
        """SSC::intialization restated (oracle/scvod_oracle.cpp); the initialised frame is read with f = -1."""
        poses = np.ascontiguousarray(poses, np.float32).reshape(-1, 6)
        self.lib.orc_initialization.restype = ctypes.c_int
        return int(self.lib.orc_initialization(self.h, _ptr(poses), len(poses)))

    def reset(self):
        self.lib.orc_reset_frames(self.h)
        self.sizes = []

    def counts(self, f):
        c = np.zeros(9, np.int32)
        self.lib.orc_frame_counts(self.h, f, _ptr(c))
        return c

    def labels(self, f):
        cls = np.zeros(max(self.sizes[f], 1), np.uint8)
        self.lib.orc_frame_labels(self.h, f, _ptr(cls))
        return cls[: self.sizes[f]]

    def ground_order(self, f):
        c = self.counts(f)
        g = np.zeros(max(int(c[1]), 1), np.int32)
        ng = np.zeros(max(int(c[2]), 1), np.int32)
        self.lib.orc_frame_ground_order(self.h, f, _ptr(g), _ptr(ng))
        return g[: c[1]], ng[: c[2]]

    def apri(self, f):
        m = int(self.counts(f)[3])
        src = np.zeros(max(m, 1), np.int32)
        vid = np.zeros(max(m, 1), np.int32)
        self.lib.orc_frame_apri(self.h, f, _ptr(src), _ptr(vid))
        return src[:m], vid[:m]

    def voxels(self, f):
        v = int(self.counts(f)[4])
        cap = max(v, 1)
        out = {"voxel_idx": np.zeros(cap, np.int32), "count": np.zeros(cap, np.int32), "av": np.zeros(cap, np.float32),
               "cov": np.zeros(cap, np.float32), "center": np.zeros((cap, 3), np.float32), "tri": np.zeros((cap, 3), np.int32),
               "label": np.zeros(cap, np.int32)}
        self.lib.orc_frame_voxels(self.h, f, _ptr(out["voxel_idx"]), _ptr(out["count"]), _ptr(out["av"]), _ptr(out["cov"]),
                                  _ptr(out["center"]), _ptr(out["tri"]), _ptr(out["label"]))
        return {k: a[:v] for k, a in out.items()}

    def point_cluster(self, f, stage):
        m = int(self.counts(f)[3])
        name = np.zeros(max(m, 1), np.int32)
        self.lib.orc_frame_point_cluster(self.h, f, stage, _ptr(name))
        return name[:m]

    def clusters(self, f):
        cap = max(int(self.counts(f)[8]), 1)
        out = {"name": np.zeros(cap, np.int32), "type": np.zeros(cap, np.int32), "state": np.zeros(cap, np.int32),
               "npts": np.zeros(cap, np.int32), "nvox": np.zeros(cap, np.int32), "bbox": np.zeros((cap, 6), np.float32)}
        n = self.lib.orc_frame_clusters(self.h, f, cap, _ptr(out["name"]), _ptr(out["type"]), _ptr(out["state"]), _ptr(out["npts"]),
                                        _ptr(out["nvox"]), _ptr(out["bbox"]))
        return {k: a[:n] for k, a in out.items()}

    # stage-level helpers
    def bin(self, xyzi):
        xyzi = np.ascontiguousarray(xyzi, np.float32).reshape(-1, 4)
        n = len(xyzi)
        out = {"pass": np.zeros(n, np.uint8), "voxel_idx": np.zeros(n, np.int32), "range_idx": np.zeros(n, np.int32),
               "sector_idx": np.zeros(n, np.int32), "azimuth_idx": np.zeros(n, np.int32), "range": np.zeros(n, np.float32),
               "angle": np.zeros(n, np.float32), "azimuth": np.zeros(n, np.float32)}
        self.lib.orc_bin(ctypes.byref(self.params), _ptr(xyzi), n, _ptr(out["pass"]), _ptr(out["voxel_idx"]), _ptr(out["range_idx"]),
                         _ptr(out["sector_idx"]), _ptr(out["azimuth_idx"]), _ptr(out["range"]), _ptr(out["angle"]), _ptr(out["azimuth"]))
        return out

    def ground(self, xyzi, sensor_height=None):
        xyzi = np.ascontiguousarray(xyzi, np.float32).reshape(-1, 4)
        n = len(xyzi)
        g = np.zeros(max(n, 1), np.int32)
        ng = np.zeros(max(n, 1), np.int32)
        cls = np.zeros(max(n, 1), np.uint8)
        rec = np.zeros((504, 15), np.float32)
        cg, cng = ctypes.c_int32(), ctypes.c_int32()
        h = float(np.float32(self.params.sensor_height)) if sensor_height is None else sensor_height
        npatch = self.lib.orc_ground(_ptr(xyzi), n, ctypes.c_double(h), _ptr(g), ctypes.byref(cg), _ptr(ng), ctypes.byref(cng),
                                     _ptr(cls), _ptr(rec), 504)
        return g[: cg.value].copy(), ng[: cng.value].copy(), cls[:n], rec[:npatch]

    def atan2f_many(self, y, x):
        y = np.ascontiguousarray(y, np.float32)
        x = np.ascontiguousarray(x, np.float32)
        out = np.empty_like(y)
        self.lib.orc_atan2f_many(_ptr(y), _ptr(x), _ptr(out), ctypes.c_int64(y.size))
        return out

    def relative_pose(self, pose_next, pose_pre):
        T = np.zeros(12, np.float32)
        a = np.ascontiguousarray(pose_next, np.float32)
        b = np.ascontiguousarray(pose_pre, np.float32)
        self.lib.orc_relative_pose(_ptr(a), _ptr(b), _ptr(T))
        return T.reshape(3, 4)

    def svd3(self, A):
        A = np.ascontiguousarray(A, np.float32).reshape(9)
        U = np.zeros(9, np.float32)
        sv = np.zeros(3, np.float32)
        self.lib.orc_svd3(_ptr(A), _ptr(U), _ptr(sv))
        return U.reshape(3, 3), sv

    # GICP (oracle/gicp_oracle.cpp; docs/gicp_spec.md)
    def gicp_normals(self, xyzi, gparams):
        xyzi = np.ascontiguousarray(xyzi, np.float32).reshape(-1, 4)
        n = len(xyzi)
        nm = np.zeros((max(n, 1), 3), np.float32)
        va = np.zeros(max(n, 1), np.uint8)
        cn = np.zeros(max(n, 1), np.int32)
        self.lib.orc_gicp_normals(_ptr(xyzi), n, ctypes.byref(gparams), _ptr(nm), _ptr(va), _ptr(cn))
        return nm[:n], va[:n], cn[:n]

    def gicp_align(self, src, tgt, T0, gparams):
        pkg = load_package()
        src = np.ascontiguousarray(src, np.float32).reshape(-1, 4)
        tgt = np.ascontiguousarray(tgt, np.float32).reshape(-1, 4)
        T0 = np.ascontiguousarray(T0, np.float32).reshape(12)
        res = pkg.GicpResult()
        self.lib.orc_gicp_align(_ptr(src), len(src), _ptr(tgt), len(tgt), _ptr(T0), ctypes.byref(gparams), ctypes.byref(res))
        return res.as_dict()

    def run_sequence(self, scans, poses, nthreads=1, want_labels=True):
        off = np.zeros(len(scans) + 1, np.int64)
        off[1:] = np.cumsum([len(s) for s in scans])
        flat = np.ascontiguousarray(np.concatenate(scans, axis=0), np.float32)
        poses = np.ascontiguousarray(poses, np.float32).reshape(-1, 6)
        labels = np.zeros(max(int(off[-1]), 1), np.uint8) if want_labels else None
        secs = self.lib.orc_run_sequence(ctypes.byref(self.params), _ptr(flat), _ptr(off), len(scans), _ptr(poses), int(nthreads), _ptr(labels))
        return secs, (labels[: off[-1]] if want_labels else None), off

    def run_chunks(self, flat, off, poses, chunk, nchunks, nthreads=1, want_labels=False):
        """nchunks independent sequences of `chunk` scans (chunk c reads input chunk c % in_chunks), one thread per chunk at a
        time: the decomposition the GPU arm of bench.py uses.  Returns (seconds, labels or None)."""
        off = np.ascontiguousarray(off, np.int64)
        flat = np.ascontiguousarray(flat, np.float32)
        poses = np.ascontiguousarray(poses, np.float32).reshape(-1, 6)
        in_chunks = (len(off) - 1) // chunk
        assert in_chunks >= 1 and len(poses) >= in_chunks * chunk
        labels = np.zeros(max(int(off[in_chunks * chunk] - off[0]), 1), np.uint8) if want_labels else None
        self.lib.orc_run_chunks.restype = ctypes.c_double
        secs = self.lib.orc_run_chunks(ctypes.byref(self.params), _ptr(flat), _ptr(off), in_chunks, int(chunk), _ptr(poses), int(nchunks),
                                       int(nthreads), _ptr(labels))
        return secs, labels


@pytest.fixture(scope="session")
def pkg():
    return load_package()


@pytest.fixture(scope="session")
def kitti_params(pkg):
    return pkg.semantickitti_params()


@pytest.fixture()
def oracle(kitti_params):
    o = Oracle(kitti_params)
    yield o
    o.close()


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def copy_params(pkg, p):
    q = pkg.Params()
    import ctypes as _c
    _c.memmove(_c.byref(q), _c.byref(p), _c.sizeof(q))
    return q


def densest_object_direction(cloud):
    """Direction (rad) in which the scan has the most above-ground returns between 3 and 28 m."""
    r = np.hypot(cloud[:, 0], cloud[:, 1])
    m = (cloud[:, 2] > -1.3) & (r > 3.0) & (r < 28.0)
    h, edges = np.histogram(np.arctan2(cloud[m, 1], cloud[m, 0]), bins=72, range=(-np.pi, np.pi))
    k = int(np.argmax(h))
    return 0.5 * (edges[k] + edges[k + 1])


def taint_scan(cloud, pose=None, turn=None, frac=1.0):
    """Put points with a -1 sector index (y == 0 exactly with x > 0, ssc.cpp:186) INSIDE objects: the scan is turned about z by
    -turn (default: its densest object direction, so that objects sit on the +x axis; the pose's yaw follows), then every
    above-ground point with x > 3 m within 0.15 m of the x axis gets y = +0.0 or -0.0 (both give angle 0).
    Returns (cloud, pose, number of such points)."""
    c = cloud.copy()
    if turn is None:
        turn = densest_object_direction(c)
    cs, sn = np.float32(np.cos(-turn)), np.float32(np.sin(-turn))
    x, y = c[:, 0].copy(), c[:, 1].copy()
    c[:, 0] = cs * x - sn * y
    c[:, 1] = sn * x + cs * y
    k = np.flatnonzero((c[:, 0] > 3.0) & (np.abs(c[:, 1]) < 0.15) & (c[:, 2] > -1.3))
    if frac <= 0.0:
        k = k[:0]
    elif frac < 1.0:
        k = k[:: max(1, int(round(1.0 / frac)))]
    c[k[::2], 1] = 0.0
    c[k[1::2], 1] = -0.0
    if pose is not None:
        pose = np.array(pose, np.float32)
        pose[5] = np.float32(pose[5] + turn)  # p_map = T Rz(turn) p'
    return c, pose, len(k)


def params_with_edge_points(pkg, params, oracle_cls, cloud):
    """Parameters whose min_azimuth / min_dis equal the exact float elevation / range of points that sit inside objects of
    `cloud`, so that those points get azimuth_idx == -1 / range_idx == -1 (ssc.cpp:185,187).  The point chosen for min_dis
    lends its (x, y) to three object points near it, so several points alias."""
    o = oracle_cls(params)
    b = o.bin(cloud)
    o.close()
    c = cloud.copy()
    obj = np.flatnonzero((c[:, 2] > -1.0) & (c[:, 2] < 0.5) & (b["pass"] != 0) & (b["range"] > 6.0) & (b["range"] < 20.0))
    assert len(obj) > 200
    q = copy_params(pkg, params)
    k2 = obj[np.argsort(b["azimuth"][obj])[len(obj) // 10]]  # a low (not the lowest) object point: most of the objects stay in the window
    q.min_azimuth = float(b["azimuth"][k2])
    near = obj[(b["azimuth"][obj] > q.min_azimuth + 1.0) & (b["range"][obj] < 12.0)]
    assert len(near) > 50
    k = near[len(near) // 3]
    d = np.abs(c[near, 0] - c[k, 0]) + np.abs(c[near, 1] - c[k, 1])
    for j in near[np.argsort(d)[1:4]]:
        c[j, 0], c[j, 1] = c[k, 0], c[k, 1]
    q.min_dis = float(b["range"][k])
    return q, c
