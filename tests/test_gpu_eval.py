"""GPU suite: the quality measures on the device (scvod_evaluate_map, scvod_evaluate_confusion) against
restatements of the reference's own evaluation code: tool/analysis.py:124-194 (sklearn 1-NN, as the reference does it) and
src/evaluate.cpp:79-145 (radius tests, brute force in the kernel's float arithmetic)."""
import numpy as np
import pytest

import conftest

pytestmark = pytest.mark.gpu
DYN = (252, 253, 254, 255, 256, 257, 258, 259)


def analysis_py_evaluate(gt, est, voxelsize):
    """evaluate() + calc_naive_preservation() of tool/analysis.py:124-194 on [n, 4] arrays (column 3 = label as float)."""
    from sklearn.neighbors import NearestNeighbors

    def sem(a):
        return a[:, 3].astype(np.uint32) & 0xFFFF

    gs, es = sem(gt), sem(est)
    nbrs = NearestNeighbors(n_neighbors=1, algorithm="kd_tree").fit(est[:, :3])
    dists, idx = nbrs.kneighbors(gt[:, :3])
    dists, idx = dists.reshape(-1), idx.reshape(-1)
    inl = dists < voxelsize * np.sqrt(3) / 2
    g_in, e_in = np.isin(gs[inl], DYN), np.isin(es[idx][inl], DYN)
    n_gt_dyn, n_est_dyn = int(np.isin(gs, DYN).sum()), int(np.isin(es, DYN).sum())
    res = {"gt_dynamic": n_gt_dyn, "gt_static": len(gt) - n_gt_dyn, "est_dynamic": n_est_dyn, "est_static": len(est) - n_est_dyn,
           "preserved": int(inl.sum()), "static_preserved": int((~g_in & ~e_in).sum()), "dynamic_preserved": int((g_in & e_in).sum())}
    res["preservation_rate"] = res["static_preserved"] / res["gt_static"] * 100
    res["rejection_rate"] = (n_gt_dyn - res["dynamic_preserved"]) / n_gt_dyn * 100
    pr, rr = res["preservation_rate"] / 100, res["rejection_rate"] / 100
    res["f1"] = 2 * pr * rr / (pr + rr)
    return res, dists, idx


def random_maps(rng, n_gt, keep_static, keep_dynamic, jitter):
    gt = np.zeros((n_gt, 4), np.float32)
    gt[:, :3] = rng.uniform([-40, -40, -2], [40, 40, 4], (n_gt, 3))
    sem = rng.choice(np.array([40, 50, 70, 10, 252, 253, 259], np.uint32), n_gt, p=[0.4, 0.2, 0.15, 0.1, 0.1, 0.03, 0.02])
    gt[:, 3] = (sem | (rng.integers(0, 200, n_gt).astype(np.uint32) << 16)).astype(np.float32)
    dyn = np.isin(sem, DYN)
    keep = np.where(dyn, rng.uniform(size=n_gt) < keep_dynamic, rng.uniform(size=n_gt) < keep_static)
    est = gt[keep].copy()
    est[:, :3] += rng.normal(0, jitter, (len(est), 3)).astype(np.float32)
    return gt, est


@pytest.mark.parametrize("n_gt,voxel,jitter", [(200_000, 0.2, 0.05), (60_000, 0.15, 0.12), (5_000, 0.4, 0.3)])
def test_preservation_metric_matches_analysis_py(pkg, n_gt, voxel, jitter):
    rng = np.random.default_rng(n_gt)
    gt, est = random_maps(rng, n_gt, 0.93, 0.08, jitter)
    s = pkg.SSC(pkg.semantickitti_params(), device=0, max_points=4096, max_batch=1)
    got, nn = s.evaluate_map(gt, est, voxel, DYN, want_nn=True)
    exp, dists, idx = analysis_py_evaluate(gt, est, voxel)
    for k in ("gt_static", "gt_dynamic", "est_static", "est_dynamic", "preserved", "static_preserved", "dynamic_preserved"):
        assert got[k] == exp[k], k
    for k in ("preservation_rate", "rejection_rate", "f1"):
        assert abs(got[k] - exp[k]) < 1e-9, k
    inl = dists < voxel * np.sqrt(3) / 2
    assert np.array_equal(nn >= 0, inl)
    assert np.array_equal(nn[inl], idx[inl])  # the same neighbour wherever there is one inside the threshold
    assert got["gt_per_class"][0] == int(((gt[:, 3].astype(np.uint32) & 0xFFFF) == 252).sum())
    # degenerate inputs: an empty estimate preserves nothing and rejects everything
    e = s.evaluate_map(gt, np.zeros((0, 4), np.float32), voxel, DYN)
    assert e["preserved"] == 0 and e["rejection_rate"] == 100.0 and e["preservation_rate"] == 0.0
    s.close()


def brute_any_within(q, t, r):
    out = np.zeros(len(q), bool)
    r2 = np.float32(r) * np.float32(r)
    for i0 in range(0, len(q), 2000):
        d = t[None, :, :3] - q[i0:i0 + 2000, None, :3]
        d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]  # float32, the kernel's order
        out[i0:i0 + 2000] = (d2 < r2).any(axis=1)
    return out


def test_confusion_counts_match_brute_force(pkg):
    rng = np.random.default_rng(9)
    st = np.zeros((12000, 4), np.float32)
    st[:, :3] = rng.uniform([-20, -20, -2], [20, 20, 2], (12000, 3))
    dy = np.zeros((3000, 4), np.float32)
    dy[:, :3] = rng.uniform([-20, -3, -2], [20, 3, 0], (3000, 3))
    src = np.concatenate([st[:6000], dy[:1500], rng.uniform(-20, 20, (1000, 4)).astype(np.float32)])
    pred = src.copy()
    pred[:, :3] += rng.normal(0, 0.06, (len(pred), 3)).astype(np.float32)
    pred[:, 3] = (rng.uniform(size=len(pred)) < 0.8).astype(np.float32)  # predicted static flag (ori.g != 0)
    s = pkg.SSC(pkg.semantickitti_params(), device=0, max_points=4096, max_batch=1)
    counts, per = s.evaluate_confusion(pred, st, dy, 0.15, 0.1)
    ps = pred[:, 3] != 0
    s15, d15, s10, d10 = (brute_any_within(pred, t, r) for t, r in ((st, 0.15), (dy, 0.15), (st, 0.1), (dy, 0.1)))
    exp = np.full(len(pred), 4, np.uint8)
    exp[ps & ~s15 & d10] = 1
    exp[ps & s15] = 0
    exp[~ps & ~d15 & s10] = 3
    exp[~ps & d15] = 2
    assert np.array_equal(per, exp)
    assert list(counts) == [int((exp == k).sum()) for k in range(5)] and counts[:4].min() > 0
    c0, _ = s.evaluate_confusion(pred, np.zeros((0, 4), np.float32), np.zeros((0, 4), np.float32))
    assert list(c0) == [0, 0, 0, 0, len(pred)]
    s.close()


def test_quality_of_the_synthetic_stream(pkg):
    """Quality number for the synthetic stream (ground truth from the generator): static map = every point the path does not class
    DYNAMIC, moved to the map frame; scored like the reference scores its own maps."""
    n = 24
    params = pkg.semantickitti_params()
    data = [pkg.synth_scan_labeled(conftest.SEED + 80, k, rings=32, cols=900) for k in range(n)]
    scans, poses, labs = [d[0] for d in data], np.stack([d[1] for d in data]), [d[2] for d in data]
    for k in range(n):  # the labelled generator returns the same points
        assert np.array_equal(scans[k], pkg.synth_scan(conftest.SEED + 80, k, rings=32, cols=900)[0])
    s = pkg.SSC(params, device=0, max_points=32 * 900, max_batch=n)
    cls = s.segDF(scans, poses)
    gt, est = [], []
    for k in range(n - 1):  # the last frame is never tracked (ssc.cpp:1450)
        T = pkg.pose_matrix(poses[k]).astype(np.float64)
        p = scans[k].astype(np.float64)
        w = np.concatenate([p[:, :3] @ T[:, :3].T + T[:, 3], labs[k][:, None].astype(np.float64)], axis=1).astype(np.float32)
        keep = (labs[k] & 0xFFFF) > 1
        gt.append(w[keep])
        est.append(w[keep & (cls[k] != pkg.PT_DYNAMIC)])
    gt, est = np.concatenate(gt), np.concatenate(est)
    r = s.evaluate_map(gt, est, 0.2, DYN)
    print(f"synthetic stream quality: PR {r['preservation_rate']:.2f} %, RR {r['rejection_rate']:.2f} %, F1 {r['f1']:.4f} "
          f"({r['gt_dynamic']} moving-car points of {len(gt)})")
    assert r["gt_dynamic"] > 1000 and r["preservation_rate"] > 95.0 and r["rejection_rate"] > 20.0
    s.close()
