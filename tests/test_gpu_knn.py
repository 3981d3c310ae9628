"""GPU suite: k-NN normals and the two reference routines on top of them: SSC::intensityCalibrationByCurvature
(src/ssc.cpp:98-153) and SSC::regionGrowing (src/ssc.cpp:797-832), against scipy / numpy restatements (PCL is not available)."""
import numpy as np
import pytest
from scipy.spatial import cKDTree

import conftest

pytestmark = pytest.mark.gpu


def numpy_normals(pts, k):
    tree = cKDTree(pts[:, :3].astype(np.float64))
    d, idx = tree.query(pts[:, :3].astype(np.float64), k=k)
    nb = pts[idx, :3].astype(np.float64)
    c = nb - nb.mean(axis=1, keepdims=True)
    cov = np.einsum("nki,nkj->nij", c, c) / k
    w, v = np.linalg.eigh(cov)
    nrm = v[:, :, 0]
    flip = np.einsum("ni,ni->n", -pts[:, :3].astype(np.float64), nrm) < 0
    nrm[flip] *= -1
    return idx, d, nrm, np.abs(w[:, 0] / w.sum(axis=1)), w


def test_knn_normals_match_kdtree_and_eigh(pkg):
    s = pkg.SSC(pkg.semantickitti_params(), device=0, max_points=4096, max_batch=1)
    cloud, _ = pkg.synth_scan(conftest.SEED + 90, 3, rings=64, cols=1800)
    k = 10
    nm, cv, nb = s.knn_normals(cloud, k)
    idx, d, nrm, curv, w = numpy_normals(cloud, k)
    # the same neighbour sets wherever the k-th and (k+1)-th distances are not (nearly) tied
    tree = cKDTree(cloud[:, :3].astype(np.float64))
    d11, _ = tree.query(cloud[:, :3].astype(np.float64), k=k + 1)
    clear = d11[:, k] - d11[:, k - 1] > 1e-5
    assert clear.mean() > 0.99
    assert np.array_equal(np.sort(nb[clear], axis=1), np.sort(idx[clear], axis=1))
    assert np.all(nb[:, 0] == np.arange(len(cloud)))  # the point itself comes first (distance 0)
    # normals: same direction (and sign: towards the sensor) where the smallest eigenvalue is separated
    sep = (w[:, 1] - w[:, 0]) > 1e-3 * w[:, 2]
    ok = clear & sep
    dots = np.einsum("ni,ni->n", nm[ok].astype(np.float64), nrm[ok])
    assert ok.mean() > 0.8 and np.all(dots > 1 - 1e-5)
    assert np.allclose(cv[clear], curv[clear], atol=2e-5)
    # small and degenerate inputs
    e = s.knn_normals(np.zeros((0, 4), np.float32), k)
    assert len(e[0]) == 0
    nm2, cv2, nb2 = s.knn_normals(np.array([[1, 2, 3, 0], [1.1, 2, 3, 0]], np.float32), k)
    assert np.isnan(nm2).all() and list(nb2[0][:3]) == [0, 1, -1]
    s.close()


def test_intensity_calibration_matches_restatement(pkg):
    s = pkg.SSC(pkg.semantickitti_params(), device=0, max_points=4096, max_batch=1)
    cloud, _ = pkg.synth_scan(conftest.SEED + 91, 1, rings=32, cols=900)
    cloud = cloud.copy()
    cloud[::50, 3] = 300.0  # above max_intensity: clamped first (ssc.cpp:101-105)
    out = s.intensityCalibrationByCurvature(cloud, search_num=10, max_intensity=255.0)
    idx, d, nrm, curv, w = numpy_normals(cloud, 10)
    inten = np.minimum(cloud[:, 3].astype(np.float64), 255.0)
    p = cloud[:, :3].astype(np.float64)
    cosang = np.abs(np.einsum("ni,ni->n", nrm, p) / (np.linalg.norm(nrm, axis=1) * np.linalg.norm(p, axis=1)))
    cosang = np.maximum(cosang, 0.3)
    exp = np.minimum(inten / cosang, 255.0)
    sep = (w[:, 1] - w[:, 0]) > 1e-2 * w[:, 2]
    assert np.array_equal(out[:, :3], cloud[:, :3])
    assert sep.mean() > 0.7 and np.allclose(out[sep, 3], exp[sep], rtol=2e-3, atol=2e-3)
    assert out[:, 3].max() <= 255.0 and (out[:, 3] >= inten.astype(np.float32) - 1e-3).all()
    s.close()


def python_region_growing(pts, k=10):
    """pcl::RegionGrowing::extract as SSC::regionGrowing configures it, with scipy / numpy normals."""
    idx, d, nrm, curv, w = numpy_normals(pts, k)
    n = len(pts)
    order = np.argsort(curv.astype(np.float32), kind="stable")
    label = np.full(n, -1)
    cos_thr = np.cos(np.float32(10.0 / 180.0 * np.pi))
    sizes = []
    pos = 0
    seed = order[0]
    done = 0
    while done < n:
        q = [seed]
        label[seed] = len(sizes)
        cnt = 1
        while q:
            cur = q.pop(0)
            for j in idx[cur]:
                if label[j] != -1:
                    continue
                if abs(float(nrm[cur] @ nrm[j])) < cos_thr:
                    continue
                label[j] = len(sizes)
                cnt += 1
                q.append(j)
        sizes.append(cnt)
        done += cnt
        for t in range(pos + 1, n):
            if label[order[t]] == -1:
                seed, pos = order[t], t
                break
    planar = sum(c for c in sizes if c >= 20)
    return planar >= n * 0.2, label, planar


def test_region_growing_building_vs_tree(pkg):
    s = pkg.SSC(pkg.semantickitti_params(), device=0, max_points=4096, max_batch=1)
    rng = np.random.default_rng(4)
    wall = np.zeros((3000, 4), np.float32)
    wall[:, 0] = 12.0 + rng.normal(0, 0.01, 3000)
    wall[:, 1] = rng.uniform(-8, 8, 3000)
    wall[:, 2] = rng.uniform(-1.7, 4, 3000)
    crown = np.zeros((3000, 4), np.float32)
    crown[:, :3] = rng.normal(0, 1.0, (3000, 3)) + np.array([8.0, 5.0, 3.0])
    corner = np.concatenate([wall[:1500], np.stack([rng.uniform(4, 12, 1500), np.full(1500, 8.0) + rng.normal(0, 0.01, 1500), rng.uniform(-1.7, 4, 1500),
                                                    np.zeros(1500)], 1).astype(np.float32)])
    for name, cl, want in (("wall", wall, True), ("crown", crown, False), ("corner", corner, True)):
        got, seg, planar = s.regionGrowing(cl)
        exp, lab, eplanar = python_region_growing(cl)
        assert got == exp == want, name
        assert abs(planar - eplanar) <= 0.02 * len(cl), (name, planar, eplanar)
        assert seg.min() >= 0
    # a real cluster of the synthetic scene: the points of one building wall and of one tree crown from the labelled generator
    cloud, _, lab = pkg.synth_scan_labeled(conftest.SEED + 92, 2)
    sem, inst = lab & 0xFFFF, lab >> 16
    r = np.hypot(cloud[:, 0], cloud[:, 1])
    b = cloud[(sem == 50) & (r < 25) & (cloud[:, 1] > 0)]
    t = cloud[np.isin(sem, [70, 71]) & (r < 40)]
    assert len(b) > 300
    assert s.regionGrowing(b)[0] is True and python_region_growing(b)[0]
    if len(t) > 50:
        assert s.regionGrowing(t)[0] == python_region_growing(t)[0]
    s.close()
