"""GICP scan-to-map stage (docs/gicp_spec.md).

The reference has no GICP (src/gicp.cpp is a PCD merge tool), so parity here is SELF-CONSISTENCY: the CUDA
path through the C-ABI against the double-precision CPU oracle derived from the same spec, pose within
1e-3 m / 1e-3 rad (BASELINE.json north_star).  The CPU tests pin the oracle itself against known rigid
transforms.
"""
import numpy as np
import pytest

import conftest

POSE_TOL_M = 1e-3
POSE_TOL_RAD = 1e-3


def to_world(pkg, scan, pose):
    T = pkg.pose_matrix(pose)
    out = scan.copy()
    out[:, :3] = scan[:, :3] @ T[:, :3].T + T[:, 3]
    return out


def scan_to_map_case(pkg, seed, first, nmap, rings, cols, dx=0.15, dy=-0.1, dyaw_deg=1.5):
    """map = nmap consecutive scans moved to the world frame; source = the next scan; T0 = its pose perturbed."""
    scans = [pkg.synth_scan(seed, k, rings=rings, cols=cols) for k in range(first, first + nmap + 1)]
    tgt = np.concatenate([to_world(pkg, s, p) for s, p in scans[:nmap]])
    src, pose = scans[nmap]
    pert = pose.copy()
    pert[0] += dx
    pert[1] += dy
    pert[5] += np.deg2rad(dyaw_deg)
    return src, tgt, pkg.pose_matrix(pert), pose


def pose_close(a, b):
    return np.abs(a[:3] - b[:3]).max() <= POSE_TOL_M and np.abs(a[3:] - b[3:]).max() <= POSE_TOL_RAD


# ---- CPU: the oracle against known answers --------------------------------------------------------------------
def test_oracle_recovers_known_rigid_transform(pkg, oracle):
    """src = T_true^-1 (tgt): same surface samples, so the optimum is T_true exactly (residuals vanish)."""
    tgt, _ = pkg.synth_scan(conftest.SEED, 2, rings=32, cols=900)
    pose_true = np.array([0.4, -0.25, 0.05, 0.01, -0.015, 0.03], np.float32)
    T = pkg.pose_matrix(pose_true).astype(np.float64)
    R, t = T[:, :3], T[:, 3]
    src = tgt.copy()
    src[:, :3] = ((tgt[:, :3].astype(np.float64) - t) @ R).astype(np.float32)  # R^T (b - t)
    gp = pkg.gicp_default_params()
    r = oracle.gicp_align(src, tgt, pkg.pose_matrix(np.zeros(6, np.float32)), gp)
    assert r["converged"] and r["n_corr"] > 10000
    assert np.abs(r["pose6"][:3] - pose_true[:3]).max() < 2e-4
    assert np.abs(r["pose6"][3:] - pose_true[3:]).max() < 2e-5
    assert np.allclose(r["H"], r["H"].T) and np.all(np.linalg.eigvalsh(r["H"]) > 0)
    assert np.allclose(pkg.pose_matrix(r["pose6"]), r["T"], atol=2e-5)  # pose6 is the inverse of getTransformation


def test_oracle_normals_on_planes(pkg, oracle):
    rng = np.random.default_rng(4)
    n = 4000
    gp = pkg.gicp_default_params()
    # a horizontal plane and a vertical wall, 5 cm sampling
    a = np.zeros((n, 4), np.float32)
    a[:, 0] = rng.uniform(0, 6, n)
    a[:, 1] = rng.uniform(0, 6, n)
    a[:, 2] = rng.normal(0, 0.005, n)
    b = np.zeros((n, 4), np.float32)
    b[:, 0] = 20 + rng.normal(0, 0.005, n)
    b[:, 1] = rng.uniform(0, 6, n)
    b[:, 2] = rng.uniform(0, 6, n)
    line = np.zeros((200, 4), np.float32)  # a pole: no well-defined plane -> invalid
    line[:, 0], line[:, 1], line[:, 2] = 40, 40, np.linspace(0, 5, 200)
    nm, va, cn = oracle.gicp_normals(np.concatenate([a, b, line]), gp)
    assert va[:n].mean() > 0.99 and va[n:2 * n].mean() > 0.99
    assert np.abs(nm[:n][va[:n] == 1][:, 2]).min() > 0.99       # plane normal = +-z
    assert np.abs(nm[n:2 * n][va[n:2 * n] == 1][:, 0]).min() > 0.99  # wall normal = +-x
    assert va[2 * n:].sum() == 0
    assert cn.min() >= 1  # every point finds at least itself
    # brute-force neighbour counts
    pts = np.concatenate([a, b, line])[:, :3].astype(np.float64)
    sel = rng.choice(len(pts), 200, replace=False)
    d2 = ((pts[sel, None, :] - pts[None, :, :]) ** 2).sum(-1)
    brute = (d2 <= float(np.float32(gp.cov_radius)) ** 2).sum(1)
    assert np.abs(brute - cn[sel]).max() <= 1  # float vs double radius test on the boundary


def test_oracle_scan_to_map_converges_near_truth(pkg, oracle):
    src, tgt, T0, pose = scan_to_map_case(pkg, conftest.SEED, 0, 3, 32, 900)
    r = oracle.gicp_align(src, tgt, T0, pkg.gicp_default_params())
    assert r["converged"] and r["iterations"] <= 20
    # different surface samples + moving cars: centimetre-level agreement with the generator's pose
    assert np.abs(r["pose6"][:3] - pose[:3]).max() < 0.1 and np.abs(r["pose6"][3:] - pose[3:]).max() < 5e-3


# ---- GPU: the CUDA path against the oracle -------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_normals_match_oracle(pkg, oracle):
    s = pkg.SSC(pkg.semantickitti_params(), device=0, max_points=4096, max_batch=1)
    gp = pkg.gicp_default_params()
    cloud, _ = pkg.synth_scan(conftest.SEED, 1, rings=64, cols=1800)
    nm, va, cn = s.gicp_normals(cloud, gp)
    onm, ova, ocn = oracle.gicp_normals(cloud, gp)
    assert (cn != ocn).mean() < 1e-4 and np.abs(cn - ocn).max() <= 1   # radius-boundary flips only
    assert (va != ova).mean() < 1e-3                                    # planarity-threshold flips only
    both = (va == 1) & (ova == 1)
    assert both.sum() > 0.5 * len(cloud)
    dots = np.abs((nm[both] * onm[both]).sum(1))
    assert dots.min() > 1 - 1e-4
    # empty and tiny inputs
    e = s.gicp_normals(np.zeros((0, 4), np.float32), gp)
    assert len(e[0]) == 0
    one = s.gicp_normals(np.array([[1, 2, 3, 0]], np.float32), gp)
    assert one[1][0] == 0 and one[2][0] == 1
    s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("config,rings,cols,nmap", [("semantickitti", 32, 900, 3), ("semantickitti", 64, 1800, 3), ("parkinglot", 128, 2700, 2)])
def test_gpu_gicp_pose_matches_oracle(pkg, config, rings, cols, nmap):
    """configs[2]: dense scan vs accumulated map, pose within 1e-3 m / 1e-3 rad of the CPU GICP."""
    P = pkg.semantickitti_params() if config == "semantickitti" else pkg.parkinglot_params()
    s = pkg.SSC(P, device=0, max_points=4096, max_batch=1)
    o = conftest.Oracle(P)
    gp = pkg.gicp_default_params()
    src, tgt, T0, pose = scan_to_map_case(pkg, conftest.SEED + 5, 2, nmap, rings, cols)
    s.gicp_set_target(tgt, gp)
    g = s.gicp_align(src, T0)
    c = o.gicp_align(src, tgt, T0, gp)
    assert g["converged"] and c["converged"]
    assert pose_close(g["pose6"], c["pose6"]), (g["pose6"], c["pose6"])
    assert np.abs(g["T"] - c["T"]).max() <= 1e-3
    assert abs(g["n_corr"] - c["n_corr"]) <= max(5, 2e-3 * c["n_corr"])
    assert np.allclose(g["H"], c["H"], rtol=5e-3, atol=1e-2 * np.abs(c["H"]).max())
    # second alignment against the same target from another start converges to the same pose
    pert = pose.copy()
    pert[0] -= 0.1
    pert[5] -= np.deg2rad(1.0)
    g2 = s.gicp_align(src, pkg.pose_matrix(pert))
    assert pose_close(g2["pose6"], g["pose6"])
    s.close()
    o.close()


@pytest.mark.gpu
def test_gpu_gicp_is_run_to_run_deterministic(pkg):
    """The Gauss-Newton sums are reduced in 64-bit fixed point: results are bit-identical across runs."""
    s = pkg.SSC(pkg.semantickitti_params(), device=0, max_points=4096, max_batch=1)
    src, tgt, T0, _ = scan_to_map_case(pkg, conftest.SEED + 6, 0, 3, 32, 900)
    s.gicp_set_target(tgt, pkg.gicp_default_params())
    a = s.gicp_align(src, T0)
    for _ in range(3):
        b = s.gicp_align(src, T0)
        assert np.array_equal(a["H"], b["H"]) and np.array_equal(a["b"], b["b"]) and np.array_equal(a["T"], b["T"])
        assert a["iterations"] == b["iterations"] and a["n_corr"] == b["n_corr"]
    s.close()


@pytest.mark.gpu
def test_gpu_gicp_error_paths(pkg):
    s = pkg.SSC(pkg.semantickitti_params(), device=0, max_points=4096, max_batch=1)
    cloud, _ = pkg.synth_scan(conftest.SEED, 0, rings=16, cols=450)
    with pytest.raises(pkg.ScvodError):
        s.gicp_align(cloud, np.eye(4, dtype=np.float32)[:3])  # no target yet
    s.gicp_set_target(cloud, pkg.gicp_default_params())
    far = np.eye(4, dtype=np.float32)[:3].copy()
    far[0, 3] = 500.0  # no overlap: zero correspondences, not converged, pose unchanged
    r = s.gicp_align(cloud, far)
    assert not r["converged"] and r["n_corr"] == 0 and np.allclose(r["T"], far)
    s.close()
