"""scvod_b200 — host-side Python mirror of the reference's SSC interface over the C-ABI of
``libscvod_b200.so`` (include/scvod.h).

The reference (Yixin-F/DR-Using-SCV-OD) is a C++ program whose public surface is ``class SSC``
(reference include/ssc.h:7-105): ``process`` -> ``segment`` -> ``recognize`` per scan, then ``tracking``
over consecutive frames, driven by ``segDF`` (reference src/ssc.cpp:1428-1452).  This module keeps
those names and argument meanings; every call goes through ctypes into the CUDA library.  There is no
CPU fallback: importing works without a GPU (so the symbol table can be checked), computing does not.
"""
from __future__ import annotations

import ctypes
import os
from typing import Iterable, List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libscvod_b200.so")
SYNTH_LIB_PATH = os.path.join(_HERE, "libscvod_synth.so")  # scan generator + parameter sets, no CUDA (csrc/Makefile)

# enum scvod_point_class (include/scvod.h)
PT_DROPPED_LOW, PT_DROPPED_RANGE, PT_DROPPED_SPARSE, PT_GROUND, PT_GATED_OUT, PT_UNCLUSTERED, PT_STATIC, PT_DYNAMIC = range(8)
INIT_FRAME = -1  # SCVOD_INIT_FRAME: the frame produced by SSC.intialization


class Params(ctypes.Structure):
    """scvod_params: the Utility fields the path reads (reference include/utility.h:209-240)."""

    _fields_ = (
        [(n, ctypes.c_float) for n in (
            "sensor_height", "min_dis", "max_dis", "min_angle", "max_angle", "min_azimuth", "max_azimuth",
            "range_res", "sector_res", "azimuth_res", "refine_height", "max_z", "min_z", "car_square")]
        + [(n, ctypes.c_int32) for n in ("iteration", "toBeClass", "search_c")]
        + [(n, ctypes.c_float) for n in ("intensity_diff", "intensity_cov", "occupancy")]
        + [(n, ctypes.c_int32) for n in ("building", "tree", "car")]
    )


class GicpParams(ctypes.Structure):
    """scvod_gicp_params (docs/gicp_spec.md §1)."""

    _fields_ = [("cov_radius", ctypes.c_float), ("max_corr_dist", ctypes.c_float), ("cov_eps", ctypes.c_float), ("planarity", ctypes.c_float),
                ("min_neighbors", ctypes.c_int32), ("max_iter", ctypes.c_int32), ("rot_eps", ctypes.c_float), ("trans_eps", ctypes.c_float)]


class GicpResult(ctypes.Structure):
    """scvod_gicp_result."""

    _fields_ = [("T", ctypes.c_float * 12), ("pose6", ctypes.c_float * 6), ("H", ctypes.c_double * 36), ("b", ctypes.c_double * 6),
                ("cost", ctypes.c_double), ("iterations", ctypes.c_int32), ("n_corr", ctypes.c_int32), ("converged", ctypes.c_int32),
                ("n_src_valid", ctypes.c_int32), ("n_tgt_valid", ctypes.c_int32)]

    def as_dict(self) -> dict:
        return {"T": np.array(self.T, np.float32).reshape(3, 4), "pose6": np.array(self.pose6, np.float32),
                "H": np.array(self.H, np.float64).reshape(6, 6), "b": np.array(self.b, np.float64), "cost": float(self.cost),
                "iterations": int(self.iterations), "n_corr": int(self.n_corr), "converged": bool(self.converged)}


class EvalResult(ctypes.Structure):
    """scvod_eval_result (tool/analysis.py:124-194)."""

    _fields_ = ([(n, ctypes.c_int64) for n in ("gt_static", "gt_dynamic", "est_static", "est_dynamic", "preserved", "static_preserved", "dynamic_preserved")]
                + [(n, ctypes.c_double) for n in ("preservation_rate", "rejection_rate", "f1")]
                + [("gt_per_class", ctypes.c_int64 * 8), ("est_per_class", ctypes.c_int64 * 8)])

    def as_dict(self) -> dict:
        d = {n: getattr(self, n) for n, _ in self._fields_[:10]}
        d["gt_per_class"] = list(self.gt_per_class)
        d["est_per_class"] = list(self.est_per_class)
        return d


DYNAMIC_CLASSES = (252, 253, 254, 255, 256, 257, 258, 259)  # tool/analysis.py:6, config/semantickitti.yaml dynamic_label_


class Grid(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("range_num", "sector_num", "azimuth_num", "bin_num")]


EXPORTS = (
    "scvod_params_semantickitti", "scvod_params_parkinglot", "scvod_grid_dims", "scvod_create", "scvod_destroy",
    "scvod_last_error", "scvod_num_kernel_launches", "scvod_set_option", "scvod_ground", "scvod_bin", "scvod_push_scans",
    "scvod_push_scans_dev", "scvod_track", "scvod_num_frames", "scvod_reset_frames", "scvod_frame_labels",
    "scvod_labels_range", "scvod_frame_counts", "scvod_frame_ground_order", "scvod_frame_apri", "scvod_frame_voxels",
    "scvod_frame_point_cluster", "scvod_frame_clusters", "scvod_static_submap_dev", "scvod_last_patch_records",
    "scvod_atan2f_device", "scvod_relative_pose", "scvod_synth_scan", "scvod_host_segment", "scvod_host_segment_pts", "scvod_set_stream", "scvod_kernel_timing", "scvod_kernel_timing_report", "scvod_get_stat",
    "scvod_gicp_default_params", "scvod_gicp_set_target", "scvod_gicp_set_target_dev", "scvod_gicp_align", "scvod_gicp_align_dev",
    "scvod_gicp_normals", "scvod_pose_matrix", "scvod_initialization", "scvod_prefetch_scans",
    "scvod_export_tail", "scvod_track_from_tail", "scvod_apply_tail_states", "scvod_load_kitti", "scvod_load_kitti_dev",
    "scvod_evaluate_map", "scvod_evaluate_confusion", "scvod_synth_scan_labeled",
    "scvod_knn_normals", "scvod_calibrate_intensity", "scvod_region_growing", "scvod_bin_filter_check", "scvod_last_failed_scan",
)

_lib = None


def load_library() -> ctypes.CDLL:
    """Load libscvod_b200.so; raises (loudly) when the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
                "(or __graft_entry__.build()); there is no CPU fallback for the SCV-OD path")
        lib = ctypes.CDLL(LIB_PATH)
        lib.scvod_last_error.restype = ctypes.c_char_p
        lib.scvod_relative_pose.restype = None
        lib.scvod_params_semantickitti.restype = None
        lib.scvod_params_parkinglot.restype = None
        lib.scvod_gicp_default_params.restype = None
        lib.scvod_pose_matrix.restype = None
        _lib = lib
    return _lib


_synth = None


def load_synth_library() -> ctypes.CDLL:
    """libscvod_synth.so: the deterministic scan generator and the two YAML parameter sets, host-only C++.  The CPU reference
    arm of bench.py and the oracle-side tools use nothing else of this package, so they never map the CUDA library."""
    global _synth
    if _synth is None:
        if not os.path.exists(SYNTH_LIB_PATH):
            raise RuntimeError(f"{SYNTH_LIB_PATH} is missing: build it with `make -C {os.path.join(_HERE, 'csrc')}`")
        lib = ctypes.CDLL(SYNTH_LIB_PATH)
        lib.scvod_params_semantickitti.restype = None
        lib.scvod_params_parkinglot.restype = None
        _synth = lib
    return _synth


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class ScvodError(RuntimeError):
    pass


def _check(rc: int):
    if rc < 0:
        raise ScvodError(f"scvod error {rc}: {load_library().scvod_last_error().decode()}")
    return rc


def semantickitti_params() -> Params:
    p = Params()
    load_synth_library().scvod_params_semantickitti(ctypes.byref(p))
    return p


def parkinglot_params() -> Params:
    p = Params()
    load_synth_library().scvod_params_parkinglot(ctypes.byref(p))
    return p


def grid_dims(p: Params) -> Grid:
    g = Grid()
    _check(load_library().scvod_grid_dims(ctypes.byref(p), ctypes.byref(g)))
    return g


def synth_scan(seed: int, scan_id: int, rings: int = 64, cols: int = 1800):
    """Deterministic synthetic scan. Returns (xyzi float32 [n,4], pose6 float32 [6])."""
    buf = np.empty((rings * cols, 4), np.float32)
    n = ctypes.c_int(0)
    pose = np.zeros(6, np.float32)
    if load_synth_library().scvod_synth_scan(ctypes.c_uint64(seed), int(scan_id), int(rings), int(cols), _ptr(buf), ctypes.byref(n), _ptr(pose)) < 0:
        raise ScvodError("scvod_synth_scan: bad arguments")
    return np.ascontiguousarray(buf[: n.value]), pose


def synth_scan_labeled(seed: int, scan_id: int, rings: int = 64, cols: int = 1800):
    """synth_scan plus a SemanticKITTI-style label per point (252 in the low 16 bits = moving car).  Returns (xyzi, pose6, labels)."""
    buf = np.empty((rings * cols, 4), np.float32)
    lab = np.zeros(rings * cols, np.uint32)
    n = ctypes.c_int(0)
    pose = np.zeros(6, np.float32)
    if load_synth_library().scvod_synth_scan_labeled(ctypes.c_uint64(seed), int(scan_id), int(rings), int(cols), _ptr(buf), ctypes.byref(n), _ptr(pose), _ptr(lab)) < 0:
        raise ScvodError("scvod_synth_scan_labeled: bad arguments")
    return np.ascontiguousarray(buf[: n.value]), pose, np.ascontiguousarray(lab[: n.value])


def relative_pose(pose_next: np.ndarray, pose_pre: np.ndarray) -> np.ndarray:
    T = np.zeros(12, np.float32)
    a = np.ascontiguousarray(pose_next, np.float32)
    b = np.ascontiguousarray(pose_pre, np.float32)
    load_library().scvod_relative_pose(_ptr(a), _ptr(b), _ptr(T))
    return T.reshape(3, 4)


def pose_matrix(pose6: np.ndarray) -> np.ndarray:
    """pcl::getTransformation(x,y,z,roll,pitch,yaw) as a 3x4 float matrix (reference src/ssc.cpp:1255)."""
    T = np.zeros(12, np.float32)
    a = np.ascontiguousarray(pose6, np.float32)
    load_library().scvod_pose_matrix(_ptr(a), _ptr(T))
    return T.reshape(3, 4)


def gicp_default_params() -> GicpParams:
    p = GicpParams()
    load_library().scvod_gicp_default_params(ctypes.byref(p))
    return p


def kernel_timing(enable: bool):
    _check(load_library().scvod_kernel_timing(1 if enable else 0))


def kernel_timing_report() -> dict:
    """{kernel: (total_ms, launches)} measured with CUDA events on the launching stream."""
    buf = ctypes.create_string_buffer(1 << 16)
    _check(load_library().scvod_kernel_timing_report(buf, len(buf)))
    out = {}
    for line in buf.value.decode().splitlines():
        name, ms, cnt = line.split()
        out[name] = (float(ms), int(cnt))
    return out


class SSC:
    """Mirror of the reference's ``class SSC`` for the hot path (reference include/ssc.h:55-104).

    ``process(clouds)`` runs process+segment+recognize for a list of scans (the per-scan loop of segDF,
    src/ssc.cpp:1435-1444) in batched GPU passes; ``tracking(poses)`` runs the frame chain
    (src/ssc.cpp:1450-1452); ``segDF(clouds, poses)`` does both and returns per-point classes.
    """

    def __init__(self, params: Optional[Params] = None, device: int = 0, max_points: int = 131072, max_batch: int = 16):
        self._lib = load_library()
        self.params = params if params is not None else semantickitti_params()
        g = grid_dims(self.params)
        self.range_num, self.sector_num, self.azimuth_num, self.bin_num = g.range_num, g.sector_num, g.azimuth_num, g.bin_num
        self._ctx = ctypes.c_void_p()
        _check(self._lib.scvod_create(ctypes.byref(self.params), int(device), int(max_points), int(max_batch), ctypes.byref(self._ctx)))
        self.frame_sizes: List[int] = []

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx:
            self._lib.scvod_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, key: str, value: int):
        _check(self._lib.scvod_set_option(self._ctx, key.encode(), int(value)))

    @property
    def last_failed_scan(self) -> int:
        """Index (inside the failing push call) of the scan that exceeded a per-scan capacity, -1 if none."""
        return int(self._lib.scvod_last_failed_scan(self._ctx))

    def stat(self, key: str) -> int:
        v = ctypes.c_int64()
        _check(self._lib.scvod_get_stat(self._ctx, key.encode(), ctypes.byref(v)))
        return v.value

    def set_stream(self, cuda_stream: int):
        _check(self._lib.scvod_set_stream(self._ctx, ctypes.c_void_p(cuda_stream)))

    # -- per-scan stages ------------------------------------------------------------------------
    def process(self, clouds: Sequence[np.ndarray]):
        """process + segment + recognize for each cloud ([n,4] float32 xyzi); appends to frame_set."""
        clouds = [np.ascontiguousarray(c, np.float32).reshape(-1, 4) for c in clouds]
        if not clouds:
            return
        off = np.zeros(len(clouds) + 1, np.int64)
        off[1:] = np.cumsum([len(c) for c in clouds])
        flat = np.concatenate(clouds, axis=0) if len(clouds) > 1 else clouds[0]
        self.process_flat(flat, off)

    def process_flat(self, flat: np.ndarray, offsets: np.ndarray):
        offsets = np.ascontiguousarray(offsets, np.int64)
        n0 = self.num_frames
        try:
            _check(self._lib.scvod_push_scans(self._ctx, _ptr(flat), _ptr(offsets), len(offsets) - 1))
        finally:  # a failing call keeps the batches it committed before the failing one
            self.frame_sizes.extend(int(x) for x in np.diff(offsets)[: self.num_frames - n0])

    def process_device(self, dev_ptr: int, offsets: np.ndarray):
        offsets = np.ascontiguousarray(offsets, np.int64)
        n0 = self.num_frames
        try:
            _check(self._lib.scvod_push_scans_dev(self._ctx, ctypes.c_void_p(dev_ptr), _ptr(offsets), len(offsets) - 1))
        finally:
            self.frame_sizes.extend(int(x) for x in np.diff(offsets)[: self.num_frames - n0])

    def extractGroudByPatchWork(self, cloud: np.ndarray):
        """PatchWork::estimate_ground: returns (ground_idx, nonground_idx) in the reference's output order."""
        cloud = np.ascontiguousarray(cloud, np.float32).reshape(-1, 4)
        n = len(cloud)
        g = np.empty(max(n, 1), np.int32)
        ng = np.empty(max(n, 1), np.int32)
        cg, cng = ctypes.c_int32(), ctypes.c_int32()
        _check(self._lib.scvod_ground(self._ctx, _ptr(cloud), n, _ptr(g), ctypes.byref(cg), _ptr(ng), ctypes.byref(cng)))
        return g[: cg.value].copy(), ng[: cng.value].copy()

    def last_patch_records(self, slot: int = 0) -> np.ndarray:
        rec = np.zeros((504, 12), np.float32)
        _check(self._lib.scvod_last_patch_records(self._ctx, slot, _ptr(rec)))
        return rec

    def makeApriVec(self, cloud: np.ndarray) -> dict:
        """Polar binning of an arbitrary cloud: dict of pass, voxel_idx, range_idx, sector_idx, azimuth_idx, range, angle, azimuth."""
        cloud = np.ascontiguousarray(cloud, np.float32).reshape(-1, 4)
        n = len(cloud)
        out = {
            "pass": np.zeros(n, np.uint8), "voxel_idx": np.zeros(n, np.int32), "range_idx": np.zeros(n, np.int32),
            "sector_idx": np.zeros(n, np.int32), "azimuth_idx": np.zeros(n, np.int32), "range": np.zeros(n, np.float32),
            "angle": np.zeros(n, np.float32), "azimuth": np.zeros(n, np.float32),
        }
        _check(self._lib.scvod_bin(self._ctx, _ptr(cloud), n, _ptr(out["pass"]), _ptr(out["voxel_idx"]), _ptr(out["range_idx"]),
                                   _ptr(out["sector_idx"]), _ptr(out["azimuth_idx"]), _ptr(out["range"]), _ptr(out["angle"]), _ptr(out["azimuth"])))
        return out

    def bin_filter_check(self, n: int, seed: int = 1, extent: float = 60.0, cloud: Optional[np.ndarray] = None) -> dict:
        """Soundness probe of the in-kernel binning filter: dict(points, exact, mismatches, max_dq_sector, max_dq_azimuth)."""
        st = np.zeros(7, np.uint64)
        if cloud is not None:
            cloud = np.ascontiguousarray(cloud, np.float32).reshape(-1, 4)
            n = len(cloud)
        _check(self._lib.scvod_bin_filter_check(self._ctx, _ptr(cloud) if cloud is not None else None, ctypes.c_int64(n), ctypes.c_uint32(seed),
                                                ctypes.c_float(extent), _ptr(st)))
        return {"points": int(st[0]), "exact": int(st[1]), "mismatches": int(st[2]), "max_dq_sector": float(st[3]) * 1e-9,
                "max_dq_azimuth": float(st[4]) * 1e-9, "patch_exact": int(st[5]), "patch_mismatches": int(st[6])}

    def atan2f_device(self, y: np.ndarray, x: np.ndarray) -> np.ndarray:
        y = np.ascontiguousarray(y, np.float32)
        x = np.ascontiguousarray(x, np.float32)
        out = np.empty_like(y)
        _check(self._lib.scvod_atan2f_device(self._ctx, _ptr(y), _ptr(x), _ptr(out), ctypes.c_int64(y.size)))
        return out

    # -- frame chain ------------------------------------------------------------------------------
    def tracking(self, poses: np.ndarray):
        poses = np.ascontiguousarray(poses, np.float32).reshape(-1, 6)
        _check(self._lib.scvod_track(self._ctx, _ptr(poses), len(poses)))

    # -- loader front end (SSC::getCloud, reference src/ssc.cpp:1060-1111) ----------------------
    def load_kitti(self, raw_scans: Sequence[np.ndarray], labels: Optional[Sequence[np.ndarray]] = None, leaf: float = 0.08,
                   max_intensity: float = 255.0) -> List[np.ndarray]:
        """Label mask (semantic label 0 / 1 dropped), intensity x max_intensity, 0.08 m VoxelGrid on the device.  raw_scans are
        the .bin contents ([n, 4] float32), labels the .label contents ([n] uint32) or None.  Returns the clouds process() takes."""
        raw = [np.ascontiguousarray(r, np.float32).reshape(-1, 4) for r in raw_scans]
        off = np.zeros(len(raw) + 1, np.int64)
        off[1:] = np.cumsum([len(r) for r in raw])
        flat = np.ascontiguousarray(np.concatenate(raw, axis=0)) if raw else np.zeros((0, 4), np.float32)
        lab = None
        if labels is not None:
            lab = np.ascontiguousarray(np.concatenate([np.asarray(l, np.uint32).reshape(-1) for l in labels])) if raw else np.zeros(0, np.uint32)
            if len(lab) != len(flat):
                raise ValueError("labels and points disagree in length")
        out = np.zeros((max(len(flat), 1), 4), np.float32)
        ooff = np.zeros(len(raw) + 1, np.int64)
        _check(self._lib.scvod_load_kitti(self._ctx, _ptr(flat), _ptr(lab), _ptr(off), len(raw), ctypes.c_float(leaf), ctypes.c_float(max_intensity),
                                          _ptr(out), _ptr(ooff)))
        return [out[ooff[b]:ooff[b + 1]].copy() for b in range(len(raw))]

    # -- k-NN normals / intensity calibration / region growing (src/ssc.cpp:98-153, 797-832) -------
    def knn_normals(self, cloud: np.ndarray, k: int = 10):
        """(normals [n,3], curvature [n], neighbours [n,k]) of pcl::NormalEstimation with setKSearch(k) on the cloud itself."""
        cloud = np.ascontiguousarray(cloud, np.float32).reshape(-1, 4)
        n = len(cloud)
        nm = np.zeros((max(n, 1), 3), np.float32)
        cv = np.zeros(max(n, 1), np.float32)
        nb = np.zeros((max(n, 1), k), np.int32)
        _check(self._lib.scvod_knn_normals(self._ctx, _ptr(cloud), n, int(k), _ptr(nm), _ptr(cv), _ptr(nb)))
        return nm[:n], cv[:n], nb[:n]

    def intensityCalibrationByCurvature(self, cloud: np.ndarray, search_num: int = 10, max_intensity: float = 255.0) -> np.ndarray:
        """SSC::intensityCalibrationByCurvature (reference src/ssc.cpp:98-153); returns the calibrated copy."""
        out = np.ascontiguousarray(cloud, np.float32).reshape(-1, 4).copy()
        _check(self._lib.scvod_calibrate_intensity(self._ctx, _ptr(out), len(out), int(search_num), ctypes.c_float(max_intensity)))
        return out

    def regionGrowing(self, cluster_cloud: np.ndarray):
        """SSC::regionGrowing (reference src/ssc.cpp:797-832): (is_building, segment of every point, points in planar segments)."""
        cloud = np.ascontiguousarray(cluster_cloud, np.float32).reshape(-1, 4)
        n = len(cloud)
        flag, planar = ctypes.c_int32(0), ctypes.c_int32(0)
        seg = np.zeros(max(n, 1), np.int32)
        _check(self._lib.scvod_region_growing(self._ctx, _ptr(cloud), n, ctypes.byref(flag), _ptr(seg), ctypes.byref(planar)))
        return bool(flag.value), seg[:n], int(planar.value)

    # -- quality measures (tool/analysis.py:124-194, src/evaluate.cpp:79-145) --------------------
    def evaluate_map(self, gt_xyzl: np.ndarray, est_xyzl: np.ndarray, voxelsize: float = 0.2, dynamic_classes: Sequence[int] = DYNAMIC_CLASSES,
                     want_nn: bool = False):
        """Preservation / rejection rate of an estimated static map against the ground-truth map (labels as floats in column 3)."""
        gt = np.ascontiguousarray(gt_xyzl, np.float32).reshape(-1, 4)
        est = np.ascontiguousarray(est_xyzl, np.float32).reshape(-1, 4)
        dc = np.ascontiguousarray(dynamic_classes, np.int32)
        res = EvalResult()
        nn = np.zeros(max(len(gt), 1), np.int32) if want_nn else None
        _check(self._lib.scvod_evaluate_map(self._ctx, _ptr(gt), ctypes.c_int64(len(gt)), _ptr(est), ctypes.c_int64(len(est)), ctypes.c_float(voxelsize),
                                            _ptr(dc), len(dc), ctypes.byref(res), _ptr(nn)))
        d = res.as_dict()
        return (d, nn[: len(gt)]) if want_nn else d

    def evaluate_confusion(self, pred_xyzs: np.ndarray, static_gt: np.ndarray, dynamic_gt: np.ndarray, r_hit: float = 0.15, r_miss: float = 0.1):
        """TP / FN / TN / FN / not-shown counts and the per-point class (src/evaluate.cpp:79-145)."""
        p = np.ascontiguousarray(pred_xyzs, np.float32).reshape(-1, 4)
        s = np.ascontiguousarray(static_gt, np.float32).reshape(-1, 4)
        d = np.ascontiguousarray(dynamic_gt, np.float32).reshape(-1, 4)
        counts = np.zeros(5, np.int64)
        per = np.zeros(max(len(p), 1), np.uint8)
        _check(self._lib.scvod_evaluate_confusion(self._ctx, _ptr(p), ctypes.c_int64(len(p)), _ptr(s), ctypes.c_int64(len(s)), _ptr(d), ctypes.c_int64(len(d)),
                                                  ctypes.c_float(r_hit), ctypes.c_float(r_miss), _ptr(counts), _ptr(per)))
        return counts, per[: len(p)]

    # -- one unbroken chain over a sequence cut into chunks (include/scvod.h "chain hand-off") ----
    def export_tail(self) -> np.ndarray:
        """The last frame as tracking() sees a frame_pre_ (car clusters in cluster_set order with their clouds): host bytes."""
        n = ctypes.c_size_t(0)
        _check(self._lib.scvod_export_tail(self._ctx, None, ctypes.c_size_t(0), ctypes.byref(n)))
        buf = np.zeros(int(n.value), np.uint8)
        _check(self._lib.scvod_export_tail(self._ctx, _ptr(buf), ctypes.c_size_t(buf.size), ctypes.byref(n)))
        return buf

    def track_from_tail(self, tail: np.ndarray, pose_pre: np.ndarray, pose_next: np.ndarray) -> np.ndarray:
        """tracking(tail of the previous chunk, this context's frame 0); returns (state, type) per exported car cluster."""
        tail = np.ascontiguousarray(tail, np.uint8)
        ncars = int(tail[4:8].view(np.int32)[0]) if tail.size >= 16 else 0
        st = np.zeros((max(ncars, 1), 2), np.int32)
        a = np.ascontiguousarray(pose_pre, np.float32)
        b = np.ascontiguousarray(pose_next, np.float32)
        n = _check(self._lib.scvod_track_from_tail(self._ctx, _ptr(tail), ctypes.c_size_t(tail.size), _ptr(a), _ptr(b), _ptr(st), len(st)))
        return st[:n]

    def apply_tail_states(self, state_type: np.ndarray):
        st = np.ascontiguousarray(state_type, np.int32).reshape(-1, 2)
        _check(self._lib.scvod_apply_tail_states(self._ctx, _ptr(st), len(st)))

    def intialization(self, poses: np.ndarray) -> int:
        """SSC::intialization (reference src/ssc.cpp:1148-1248, spelling as in the reference): fuses the clusters of the base frame
        (the last frame with the fewest clusters) that clusters of the other frames bridge.  Returns the base frame's index; read
        the initialised frame with frame index INIT_FRAME."""
        poses = np.ascontiguousarray(poses, np.float32).reshape(-1, 6)
        out = ctypes.c_int32(-1)
        _check(self._lib.scvod_initialization(self._ctx, _ptr(poses), len(poses), ctypes.byref(out)))
        return int(out.value)

    def segDF(self, clouds: Sequence[np.ndarray], poses: np.ndarray) -> List[np.ndarray]:
        f0 = self.num_frames
        self.process(clouds)
        self.tracking(poses)
        return [self.frame_labels(f) for f in range(f0, self.num_frames)]

    def reset(self):
        _check(self._lib.scvod_reset_frames(self._ctx))
        self.frame_sizes = []

    @property
    def num_frames(self) -> int:
        return int(self._lib.scvod_num_frames(self._ctx))

    @property
    def kernel_launches(self) -> int:
        v = ctypes.c_int64()
        _check(self._lib.scvod_num_kernel_launches(self._ctx, ctypes.byref(v)))
        return v.value

    # -- results / inspection -----------------------------------------------------------------------
    def frame_counts(self, f: int) -> np.ndarray:
        c = np.zeros(9, np.int32)
        _check(self._lib.scvod_frame_counts(self._ctx, f, _ptr(c)))
        return c

    def frame_labels(self, f: int) -> np.ndarray:
        n = int(self.frame_counts(f)[0])
        cls = np.zeros(max(n, 1), np.uint8)
        _check(self._lib.scvod_frame_labels(self._ctx, f, _ptr(cls), n))
        return cls[:n]

    def labels_range(self, f0: int, f1: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        total = sum(self.frame_sizes[f0:f1])
        if out is None:
            out = np.zeros(max(total, 1), np.uint8)
        _check(self._lib.scvod_labels_range(self._ctx, f0, f1, _ptr(out), ctypes.c_int64(out.size)))
        return out[:total]

    def refresh_labels(self, f0: int, f1: int):
        """Bring the device-resident per-point classes of frames [f0,f1) up to date (no host copy)."""
        _check(self._lib.scvod_labels_range(self._ctx, f0, f1, None, ctypes.c_int64(0)))

    def labels_into(self, f0: int, f1: int, host_ptr: int, cap: int):
        _check(self._lib.scvod_labels_range(self._ctx, f0, f1, ctypes.c_void_p(host_ptr), ctypes.c_int64(cap)))

    def prefetch_host_ptr(self, host_ptr: int, offsets: np.ndarray):
        """Start uploading the batch a later process_host_ptr(host_ptr, offsets) will be given (scvod_prefetch_scans)."""
        offsets = np.ascontiguousarray(offsets, np.int64)
        _check(self._lib.scvod_prefetch_scans(self._ctx, ctypes.c_void_p(host_ptr), _ptr(offsets), len(offsets) - 1))

    def process_host_ptr(self, host_ptr: int, offsets: np.ndarray):
        offsets = np.ascontiguousarray(offsets, np.int64)
        _check(self._lib.scvod_push_scans(self._ctx, ctypes.c_void_p(host_ptr), _ptr(offsets), len(offsets) - 1))
        self.frame_sizes.extend(int(x) for x in np.diff(offsets))

    def frame_ground_order(self, f: int):
        c = self.frame_counts(f)
        g = np.zeros(max(int(c[1]), 1), np.int32)
        ng = np.zeros(max(int(c[2]), 1), np.int32)
        _check(self._lib.scvod_frame_ground_order(self._ctx, f, _ptr(g), _ptr(ng)))
        return g[: c[1]], ng[: c[2]]

    def frame_apri(self, f: int):
        m = int(self.frame_counts(f)[3])
        src = np.zeros(max(m, 1), np.int32)
        vid = np.zeros(max(m, 1), np.int32)
        _check(self._lib.scvod_frame_apri(self._ctx, f, _ptr(src), _ptr(vid)))
        return src[:m], vid[:m]

    def frame_voxels(self, f: int) -> dict:
        v = int(self.frame_counts(f)[4])
        cap = max(v, 1)
        out = {"voxel_idx": np.zeros(cap, np.int32), "count": np.zeros(cap, np.int32), "av": np.zeros(cap, np.float32),
               "cov": np.zeros(cap, np.float32), "center": np.zeros((cap, 3), np.float32), "tri": np.zeros((cap, 3), np.int32),
               "label": np.zeros(cap, np.int32)}
        _check(self._lib.scvod_frame_voxels(self._ctx, f, _ptr(out["voxel_idx"]), _ptr(out["count"]), _ptr(out["av"]), _ptr(out["cov"]),
                                            _ptr(out["center"]), _ptr(out["tri"]), _ptr(out["label"])))
        return {k: a[:v] for k, a in out.items()}

    def frame_point_cluster(self, f: int, stage: int) -> np.ndarray:
        m = int(self.frame_counts(f)[3])
        name = np.zeros(max(m, 1), np.int32)
        _check(self._lib.scvod_frame_point_cluster(self._ctx, f, stage, _ptr(name)))
        return name[:m]

    def frame_clusters(self, f: int) -> dict:
        cap = max(int(self.frame_counts(f)[8]), 1)
        out = {"name": np.zeros(cap, np.int32), "type": np.zeros(cap, np.int32), "state": np.zeros(cap, np.int32),
               "npts": np.zeros(cap, np.int32), "nvox": np.zeros(cap, np.int32), "bbox": np.zeros((cap, 6), np.float32)}
        n = _check(self._lib.scvod_frame_clusters(self._ctx, f, cap, _ptr(out["name"]), _ptr(out["type"]), _ptr(out["state"]),
                                                  _ptr(out["npts"]), _ptr(out["nvox"]), _ptr(out["bbox"])))
        return {k: a[:n] for k, a in out.items()}

    def static_submap_device(self, f0: int, f1: int, poses: np.ndarray, out_dev_ptr: int, cap_points: int) -> int:
        poses = np.ascontiguousarray(poses, np.float32).reshape(-1, 6)
        n = ctypes.c_int64()
        _check(self._lib.scvod_static_submap_dev(self._ctx, f0, f1, _ptr(poses), ctypes.c_void_p(out_dev_ptr), ctypes.c_int64(cap_points), ctypes.byref(n)))
        return n.value

    # -- GICP scan-to-map (docs/gicp_spec.md; the reference names the stage but has no code for it) --------
    def gicp_set_target(self, cloud: np.ndarray, params: Optional[GicpParams] = None):
        cloud = np.ascontiguousarray(cloud, np.float32).reshape(-1, 4)
        _check(self._lib.scvod_gicp_set_target(self._ctx, _ptr(cloud), len(cloud), ctypes.byref(params) if params is not None else None))

    def gicp_set_target_device(self, dev_ptr: int, n: int, params: Optional[GicpParams] = None):
        _check(self._lib.scvod_gicp_set_target_dev(self._ctx, ctypes.c_void_p(dev_ptr), int(n), ctypes.byref(params) if params is not None else None))

    def gicp_align(self, cloud: np.ndarray, T0: np.ndarray) -> dict:
        cloud = np.ascontiguousarray(cloud, np.float32).reshape(-1, 4)
        T0 = np.ascontiguousarray(T0, np.float32).reshape(12)
        res = GicpResult()
        _check(self._lib.scvod_gicp_align(self._ctx, _ptr(cloud), len(cloud), _ptr(T0), ctypes.byref(res)))
        return res.as_dict()

    def gicp_align_device(self, dev_ptr: int, n: int, T0: np.ndarray) -> dict:
        T0 = np.ascontiguousarray(T0, np.float32).reshape(12)
        res = GicpResult()
        _check(self._lib.scvod_gicp_align_dev(self._ctx, ctypes.c_void_p(dev_ptr), int(n), _ptr(T0), ctypes.byref(res)))
        return res.as_dict()

    def gicp_normals(self, cloud: np.ndarray, params: Optional[GicpParams] = None):
        """Radius-search kernel alone: (normals [n,3], valid [n], neighbour count [n]) in input order."""
        cloud = np.ascontiguousarray(cloud, np.float32).reshape(-1, 4)
        n = len(cloud)
        nm = np.zeros((max(n, 1), 3), np.float32)
        va = np.zeros(max(n, 1), np.uint8)
        cn = np.zeros(max(n, 1), np.int32)
        _check(self._lib.scvod_gicp_normals(self._ctx, _ptr(cloud), n, ctypes.byref(params) if params is not None else None, _ptr(nm), _ptr(va), _ptr(cn)))
        return nm[:n], va[:n], cn[:n]
