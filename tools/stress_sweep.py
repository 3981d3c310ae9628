#!/usr/bin/env python
"""BASELINE.json configs[4]: 512k-point aggregated dense cloud, curved-voxel binning + radius-kNN kernels only, achieved
GB/s against the HBM roofline.  Prints one JSON line per (kernel, cloud size); replicas only (no exchange): run one
copy per GPU for a multi-GPU sweep.

Algorithmic bytes: binning = 16 B read + 4 B written (voxel_idx) per point; kNN / normals = 16 B read +
16 B written (normal + validity) per query point, 16 B per target point for the grid build.
Timing: CUDA events of the library's own per-kernel timers (scvod_kernel_timing), warm-up first, inputs > L2 for the
large sizes (the 512k cloud itself is L2 resident: 8 MB; that is the nature of the workload).
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import conftest  # noqa: E402

pkg = conftest.load_package()
PEAK = 6650.0  # GB/s, fallback of B200_PROFILING.md (MEASURED_PEAKS.json is absent on this pool)


def dense_cloud(n_target):
    """Aggregated dense cloud as tool/makeScan.cpp builds it: consecutive scans moved into one frame."""
    clouds, k = [], 0
    while sum(len(c) for c in clouds) < n_target:
        s, pose = pkg.synth_scan(conftest.SEED + 9, k, rings=128, cols=2250)
        T = pkg.pose_matrix(pose)
        out = s.copy()
        out[:, :3] = s[:, :3] @ T[:, :3].T + T[:, 3]
        clouds.append(out)
        k += 1
    return np.ascontiguousarray(np.concatenate(clouds)[:n_target], np.float32)


def timed(fn, reps=5):
    fn()  # warm-up (also sizes the buffers)
    pkg.kernel_timing(True)
    for _ in range(reps):
        fn()
    rep = pkg.kernel_timing_report()
    pkg.kernel_timing(False)
    return {k: (v[0] / reps, v[1] / reps) for k, v in rep.items()}  # ms and launches per repetition


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [512_000, 2_048_000, 8_192_000]
    ssc = pkg.SSC(pkg.parkinglot_params(), device=0, max_points=4096, max_batch=1)
    gp = pkg.gicp_default_params()
    for n in sizes:
        cloud = dense_cloud(n)
        rep = timed(lambda: ssc.makeApriVec(cloud))
        ms = rep["k_bin_only"][0]
        # the inspection entry point also writes the seven per-point diagnostics (range/sector/azimuth index and floats, pass flag)
        bytes_bin = n * (16 + 1 + 4 * 4 + 3 * 4)  # xyzi in; pass flag, voxel/range/sector/azimuth index, range/angle/azimuth out
        print(json.dumps({"kernel": "k_bin_only", "points": n, "ms": round(ms, 4), "GBps": round(bytes_bin / ms / 1e6, 1),
                          "frac_of_peak": round(bytes_bin / ms / 1e6 / PEAK, 4), "algorithmic_bytes": bytes_bin,
                          "Gpoints_per_s": round(n / ms / 1e6, 2)}))
        if n <= 2_100_000:
            rep = timed(lambda: ssc.gicp_normals(cloud, gp), reps=3)
            total = sum(v[0] for k, v in rep.items() if k.startswith("k_gicp"))
            knn = rep.get("k_gicp_normals", (0.0, 0))[0]
            bytes_knn = n * 32
            print(json.dumps({"kernel": "k_gicp_normals (radius kNN + 3x3 eigen)", "points": n, "ms": round(knn, 4),
                              "grid_build_ms": round(total - knn, 4), "GBps": round(bytes_knn / max(knn, 1e-9) / 1e6, 1),
                              "frac_of_peak": round(bytes_knn / max(knn, 1e-9) / 1e6 / PEAK, 4), "algorithmic_bytes": bytes_knn,
                              "kernels_ms": {k: round(v[0], 4) for k, v in sorted(rep.items(), key=lambda kv: -kv[1][0]) if k.startswith("k_gicp")}}))
    ssc.close()


if __name__ == "__main__":
    main()
