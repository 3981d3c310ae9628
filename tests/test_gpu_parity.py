"""GPU suite (-m gpu): the CUDA path, called through the C-ABI, against the CPU oracle on identical inputs.

Bars (BASELINE.json north_star): bit-exact voxel indices and per-point dynamic/static classes;
descriptor floats within 1e-4 (intensity mean/variance are in fact bit-exact; only the voxel "centre",
which goes through sinf/cosf/tanf, differs in the last bits).  Nothing here reads /root/reference.
"""
import os

import numpy as np
import pytest

import conftest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DESC_ATOL = 1e-4


@pytest.fixture(scope="module")
def ssc(pkg):
    s = pkg.SSC(pkg.semantickitti_params(), device=0, max_points=131072, max_batch=8)
    yield s
    s.close()


def bits(a):
    return a.view(np.uint32) if a.dtype == np.float32 else a


# ---- scalar arithmetic ---------------------------------------------------------------------------
def test_device_atan2f_equals_host_libm(ssc, oracle):
    rng = np.random.default_rng(2024)
    n = 4_000_000
    y = rng.uniform(-80, 80, n).astype(np.float32)
    x = rng.uniform(-80, 80, n).astype(np.float32)
    # structured edge inputs: zeros, signed zeros, axes, tiny/huge ratios, infinities, NaN
    sp = np.array([0.0, -0.0, 1.0, -1.0, 1e-30, -1e-30, 1e30, -1e30, np.inf, -np.inf, 0.4375, 0.6875, 1.1875, 2.4375, 1.5, 2.0 ** 25,
                   2.0 ** -29], np.float32)
    yy, xx = np.meshgrid(sp, sp)
    y = np.concatenate([y, yy.ravel(), rng.uniform(-3, 3, 100000).astype(np.float32), rng.normal(size=100000).astype(np.float32) * 1e-3])
    x = np.concatenate([x, xx.ravel(), np.ones(100000, np.float32), rng.uniform(1, 80, 100000).astype(np.float32)])
    d = ssc.atan2f_device(y, x)
    h = oracle.atan2f_many(y, x)
    assert int((d.view(np.uint32) != h.view(np.uint32)).sum()) == 0


@pytest.mark.parametrize("config", ["semantickitti", "parkinglot"])
def test_binning_bit_exact(pkg, config):
    P = pkg.semantickitti_params() if config == "semantickitti" else pkg.parkinglot_params()
    s = pkg.SSC(P, device=0, max_points=65536, max_batch=1)
    o = conftest.Oracle(P)
    rng = np.random.default_rng(5)
    n = 1_000_000
    pts = np.zeros((n, 4), np.float32)
    pts[:, 0] = rng.uniform(-45, 45, n)
    pts[:, 1] = rng.uniform(-45, 45, n)
    pts[:, 2] = rng.uniform(-4, 12, n)
    pts[:1000, 1] = 0.0                      # y == 0 quirk rows
    pts[1000:1100, :2] = 0.0                 # x == y == 0
    pts[1100:1200, 0] = P.min_dis
    pts[1100:1200, 1] = 0.0                  # dis == min_dis
    g, c = s.makeApriVec(pts), o.bin(pts)
    for k in g:
        assert np.array_equal(bits(g[k]), bits(c[k])), k
    z = np.load(os.path.join(GOLD, "bin_edge_cases.npz"))
    if config == "semantickitti":
        g = s.makeApriVec(z["xyzi"])
        for k in g:
            assert np.array_equal(bits(g[k]), bits(z["o_" + k])), k
    assert len(s.makeApriVec(np.zeros((0, 4), np.float32))["pass"]) == 0  # empty input
    s.close()
    o.close()


@pytest.mark.parametrize("config", ["semantickitti", "parkinglot"])
def test_filtered_binning_equals_exact_chain(pkg, config):
    """The pipeline kernels bin through a floating-point filter (dev_bin_filtered / dev_patch_filtered: approximate angles decide
    unless they are within a guard band of a bin edge or gate, otherwise the exact glibc-atan2f / double chain).  Soundness: on
    4 x 2^30 generated points (1/8 of them snapped onto axes, the origin, bin edges, gate radii) plus the oracle-checked edge-case
    fixture, the filtered result equals the exact chain's in every index and gate outcome, and the filter-decided points keep a
    >= 4x margin between the measured approximate-vs-exact distance and the guard band."""
    P = pkg.semantickitti_params() if config == "semantickitti" else pkg.parkinglot_params()
    s = pkg.SSC(P, device=0, max_points=65536, max_batch=1)
    worst_s = worst_e = 0.0
    for seed, extent in [(1, 60.0), (77, 35.0), (1234567, 90.0), (99, 8.0)]:
        st = s.bin_filter_check(1 << 30, seed=seed, extent=extent)
        assert st["points"] == 1 << 30
        assert st["mismatches"] == 0, st
        assert st["patch_mismatches"] == 0, st
        assert 0 < st["exact"] < 0.15 * st["points"], st          # the filter decides the bulk (structured cases are 1/8 of the sample)
        assert 0 < st["patch_exact"] < 0.15 * st["points"], st
        worst_s, worst_e = max(worst_s, st["max_dq_sector"]), max(worst_e, st["max_dq_azimuth"])
    # guard bands (scvod_kernel_common.cuh make_bin_params): 1e-3 degrees in bin coordinates + rounding allowance
    assert worst_s * 4 < 1e-3 / P.sector_res, worst_s
    assert worst_e * 4 < 1e-3 / P.azimuth_res, worst_e
    z = np.load(os.path.join(GOLD, "bin_edge_cases.npz"))
    st = s.bin_filter_check(0, cloud=z["xyzi"])
    assert st["mismatches"] == 0 and st["points"] == len(z["xyzi"])
    s.close()


# ---- ground ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("scan_id,rings,cols", [(0, 64, 1800), (11, 64, 1800), (2, 16, 450), (4, 128, 2250)])
def test_ground_order_bit_exact(pkg, oracle, scan_id, rings, cols):
    s = pkg.SSC(pkg.semantickitti_params(), device=0, max_points=rings * cols, max_batch=1)
    cloud, _ = pkg.synth_scan(conftest.SEED, scan_id, rings=rings, cols=cols)
    g, ng = s.extractGroudByPatchWork(cloud)
    og, ong, ocls, orec = oracle.ground(cloud)
    assert np.array_equal(g, og) and np.array_equal(ng, ong)
    rec = s.last_patch_records(0)
    base, secs = [0, 32, 160, 376], [16, 32, 54, 32]
    for r in orec:  # per-patch plane parameters, bit for bit
        pid = base[int(r[0])] + int(r[1]) * secs[int(r[0])] + int(r[2])
        assert np.array_equal(rec[pid, 0:3].view(np.uint32), r[5:8].view(np.uint32)), f"normal of patch {pid}"
        assert np.array_equal(rec[pid, 6:9].view(np.uint32), r[11:14].view(np.uint32)), f"singular values of patch {pid}"
        assert int(rec[pid, 10]) == int(r[4]) and int(rec[pid, 11]) == int(r[3])
    s.close()


def test_ground_with_tma_staged_chain(pkg, oracle):
    """The opt-in variant of the plane-fit kernel (cp.async.bulk + mbarrier ring) gives the same ground / non-ground order."""
    s = pkg.SSC(pkg.semantickitti_params(), device=0, max_points=64 * 1800, max_batch=1)
    s.set_option("chain_tma", 1)
    for scan_id, rings, cols in ((3, 64, 1800), (5, 16, 450)):
        cloud, _ = pkg.synth_scan(conftest.SEED, scan_id, rings=rings, cols=cols)
        g, ng = s.extractGroudByPatchWork(cloud)
        og, ong, _, _ = oracle.ground(cloud)
        assert np.array_equal(g, og) and np.array_equal(ng, ong)
    s.close()


def test_ground_ragged_inputs(ssc, oracle, pkg):
    for cloud in (np.zeros((0, 4), np.float32), np.array([[5, 5, -1.7, 1]], np.float32),
                  np.tile(np.array([[6, 1, -1.7, 1]], np.float32), (11, 1)) + np.arange(11, dtype=np.float32)[:, None] * 1e-3):
        g, ng = ssc.extractGroudByPatchWork(cloud)
        og, ong, _, _ = oracle.ground(cloud)
        assert np.array_equal(g, og) and np.array_equal(ng, ong)


def test_ground_sort_with_z_ties_is_deterministic(pkg):
    """std::sort is unstable, so the reference's order of equal-z points is implementation defined; here it is
    defined as (z, input index).  Patches of every sort tier (<= 1024, <= 4096, above) with many z ties: the result
    must be reproducible, a partition of the kept points, and ascending in (z, index) inside each list of a patch."""
    rng = np.random.default_rng(7)
    parts = []
    for n, ang in ((700, 0.1), (3000, 0.5), (9000, 0.9)):  # three zone-0 patches (r in [2.7, 7.5), sectors 0, 1, 2)
        r = rng.uniform(3.0, 7.0, n)
        a = rng.uniform(ang - 0.05, ang + 0.05, n)
        z = -1.73 + np.round(rng.normal(0, 0.02, n), 2)  # quantised to 1 cm: lots of ties
        parts.append(np.stack([r * np.cos(a), r * np.sin(a), z, np.full(n, 20.0)], 1))
    cloud = np.concatenate(parts).astype(np.float32)
    cloud = cloud[rng.permutation(len(cloud))]
    s = pkg.SSC(pkg.semantickitti_params(), device=0, max_points=len(cloud), max_batch=1)
    g1, ng1 = s.extractGroudByPatchWork(cloud)
    for _ in range(3):
        g2, ng2 = s.extractGroudByPatchWork(cloud)
        assert np.array_equal(g1, g2) and np.array_equal(ng1, ng2)
    both = np.concatenate([g1, ng1])
    assert len(np.unique(both)) == len(both) == len(cloud)
    # the ground list is patch-major; inside a patch it follows the sorted order
    ang = np.arctan2(cloud[:, 1].astype(np.float64), cloud[:, 0].astype(np.float64))
    sector = np.minimum((ang / (2 * np.pi / 16)).astype(int), 15)
    for lst in (g1, ng1):
        for sec in np.unique(sector[lst]):
            idx = lst[sector[lst] == sec]
            zz = cloud[idx, 2]
            assert np.all((np.diff(zz) > 0) | ((np.diff(zz) == 0) & (np.diff(idx) > 0)))
    s.close()


# ---- whole path -------------------------------------------------------------------------------------
def compare_frames(ssc, orc, nframes, pkg):
    for f in range(nframes):
        assert np.array_equal(ssc.frame_counts(f)[:8], orc.counts(f)[:8])
        g, ng = ssc.frame_ground_order(f)
        og, ong = orc.ground_order(f)
        assert np.array_equal(g, og) and np.array_equal(ng, ong)
        src, vid = ssc.frame_apri(f)
        osrc, ovid = orc.apri(f)
        assert np.array_equal(src, osrc) and np.array_equal(vid, ovid)  # bit-exact voxel indices
        vg, vo = ssc.frame_voxels(f), orc.voxels(f)
        for k in ("voxel_idx", "count", "tri", "label"):
            assert np.array_equal(vg[k], vo[k]), k
        for k in ("av", "cov", "center"):
            assert np.allclose(vg[k], vo[k], rtol=0, atol=DESC_ATOL), k
        assert np.array_equal(vg["av"].view(np.uint32), vo["av"].view(np.uint32))
        assert np.array_equal(vg["cov"].view(np.uint32), vo["cov"].view(np.uint32))
        for st in range(3):
            assert np.array_equal(ssc.frame_point_cluster(f, st), orc.point_cluster(f, st)), f"cluster names, stage {st}"
        cg, co = ssc.frame_clusters(f), orc.clusters(f)
        for k in ("name", "type", "npts", "nvox", "bbox"):
            assert np.array_equal(bits(cg[k]), bits(co[k])), k


def test_pipeline_stage_by_stage_kitti(pkg, ssc, oracle):
    ssc.reset()
    scans, poses = zip(*[pkg.synth_scan(conftest.SEED, k) for k in range(6)])
    poses = np.stack(poses)
    ssc.process(scans)
    for s in scans:
        oracle.push_scan(s)
    compare_frames(ssc, oracle, len(scans), pkg)
    ssc.tracking(poses)
    oracle.track(poses)
    ndyn = 0
    for f in range(len(scans)):
        lg, lo = ssc.frame_labels(f), oracle.labels(f)
        assert np.array_equal(lg, lo)  # bit-exact per-point classes
        ndyn += int((lg == pkg.PT_DYNAMIC).sum())
        cg, co = ssc.frame_clusters(f), oracle.clusters(f)
        for k in ("name", "type", "state", "npts", "nvox"):
            assert np.array_equal(cg[k], co[k]), k
    assert ndyn > 0  # the sequence does contain moving cars


def test_degenerate_scans_inside_a_batch(pkg, oracle):
    """Empty, tiny and ground-only scans between normal ones: same classes as the oracle, nothing crashes."""
    rng = np.random.default_rng(3)
    normal = [pkg.synth_scan(conftest.SEED + 40, k, rings=16, cols=450) for k in range(3)]
    flat = np.concatenate([rng.uniform(-20, 20, (3000, 2)), np.full((3000, 1), -1.73) + rng.normal(0, 0.01, (3000, 1)),
                           np.full((3000, 1), 20.0)], axis=1).astype(np.float32)
    scans = [normal[0][0], np.zeros((0, 4), np.float32), normal[1][0], np.array([[5, 5, -1.7, 1], [6, 5, 0.5, 9], [7, 1, 0.2, 3]], np.float32),
             flat, normal[2][0]]
    poses = np.stack([normal[0][1], normal[0][1], normal[1][1], normal[1][1], normal[1][1], normal[2][1]])
    s = pkg.SSC(pkg.semantickitti_params(), device=0, max_points=16 * 450, max_batch=8)
    labels = s.segDF(scans, poses)
    for sc in scans:
        oracle.push_scan(sc)
    oracle.track(poses)
    for f in range(len(scans)):
        assert np.array_equal(s.frame_counts(f)[:8], oracle.counts(f)[:8]), f
        assert np.array_equal(labels[f], oracle.labels(f)), f
    s.close()


def test_name_replay_global_memory_variant(pkg, oracle):
    """Very dense scans keep the union-find state of k_name_replay in global memory: same names, same clusters."""
    s = pkg.SSC(pkg.semantickitti_params(), device=0, max_points=64 * 1800, max_batch=4)
    s.set_option("replay_global", 1)
    scans, _ = zip(*[pkg.synth_scan(conftest.SEED + 5, k) for k in range(3)])
    s.process(scans)
    for sc in scans:
        oracle.push_scan(sc)
    compare_frames(s, oracle, len(scans), pkg)
    s.close()


@pytest.mark.parametrize("config,seed_off,rings,cols,nframes", [("semantickitti", 0, 64, 1800, 6), ("parkinglot", 3, 32, 900, 8),
                                                               ("parkinglot", 0, 64, 1800, 6), ("semantickitti", 1, 32, 900, 10)])
def test_initialization_matches_oracle(pkg, config, seed_off, rings, cols, nframes):
    """SSC::intialization (ssc.cpp:1148-1248): base frame choice, fused clusters (including the nameless
    cluster the reference inserts when no label passes the occupancy test), re-recognised types and voxel labels."""
    params = getattr(pkg, config + "_params")()
    s = pkg.SSC(params, device=0, max_points=rings * cols, max_batch=8)
    orc = conftest.Oracle(params)
    scans, poses = zip(*[pkg.synth_scan(conftest.SEED + seed_off, k, rings=rings, cols=cols) for k in range(nframes)])
    poses = np.stack(poses)
    s.process(scans)
    for sc in scans:
        orc.push_scan(sc)
    base = s.intialization(poses)
    assert base == orc.initialization(poses)
    cg, co = s.frame_clusters(pkg.INIT_FRAME), orc.clusters(-1)
    for k in ("name", "type", "npts", "nvox"):
        assert np.array_equal(cg[k], co[k]), k  # same clusters in the same unordered_map iteration order
    assert np.array_equal(cg["bbox"].view(np.uint32), co["bbox"].view(np.uint32))
    assert np.array_equal(s.frame_voxels(pkg.INIT_FRAME)["label"], orc.voxels(-1)["label"])
    # the sequence itself is untouched: tracking afterwards gives the usual labels
    s.tracking(poses)
    orc.track(poses)
    for f in range(nframes):
        assert np.array_equal(s.frame_labels(f), orc.labels(f))
    with pytest.raises(pkg.ScvodError):
        s.intialization(poses)  # only before tracking
    s.close()
    orc.close()


def test_long_sequence_labels(pkg, oracle):
    """40-frame chain at reduced resolution: exercises carried clouds, splits and fusions of tracking."""
    s = pkg.SSC(pkg.semantickitti_params(), device=0, max_points=32 * 900, max_batch=16)
    scans, poses = zip(*[pkg.synth_scan(conftest.SEED + 1, k, rings=32, cols=900) for k in range(40)])
    poses = np.stack(poses)
    labels = s.segDF(scans, poses)
    for sc in scans:
        oracle.push_scan(sc)
    oracle.track(poses)
    for f in range(len(scans)):
        assert np.array_equal(labels[f], oracle.labels(f)), f"frame {f}"
        cg, co = s.frame_clusters(f), oracle.clusters(f)
        assert np.array_equal(cg["name"], co["name"]) and np.array_equal(cg["state"], co["state"])
    s.close()


def test_incremental_push_and_track_equals_one_shot(pkg):
    scans, poses = zip(*[pkg.synth_scan(conftest.SEED + 2, k, rings=32, cols=900) for k in range(9)])
    poses = np.stack(poses)
    a = pkg.SSC(pkg.semantickitti_params(), device=0, max_points=32 * 900, max_batch=16)
    la = a.segDF(scans, poses)
    b = pkg.SSC(pkg.semantickitti_params(), device=0, max_points=32 * 900, max_batch=2)  # forces 5 batches
    for i in range(0, 9, 3):
        b.process(scans[i:i + 3])
        b.tracking(poses[: i + 3])
    for f in range(9):
        assert np.array_equal(la[f], b.frame_labels(f))
    a.close()
    b.close()


def test_parkinglot_dense_scan(pkg):
    """parkinglot.yaml parameters on a dense (3x) cloud: descriptor and labels vs the oracle."""
    P = pkg.parkinglot_params()
    s = pkg.SSC(P, device=0, max_points=128 * 2700, max_batch=2)
    o = conftest.Oracle(P)
    scans, poses = zip(*[pkg.synth_scan(conftest.SEED + 3, k, rings=128, cols=2700) for k in range(2)])
    poses = np.stack(poses)
    s.process(scans)
    for sc in scans:
        o.push_scan(sc)
    compare_frames(s, o, 2, pkg)
    s.tracking(poses)
    o.track(poses)
    for f in range(2):
        assert np.array_equal(s.frame_labels(f), o.labels(f))
    s.close()
    o.close()


def test_golden_fixture_through_gpu(pkg):
    z = np.load(os.path.join(GOLD, "scan_small.npz"))
    s = pkg.SSC(pkg.semantickitti_params(), device=0, max_points=16 * 450, max_batch=4)
    labels = s.segDF([z[f"xyzi{f}"] for f in range(3)], z["poses"])
    for f in range(3):
        src, vid = s.frame_apri(f)
        assert np.array_equal(src, z[f"apri_src{f}"]) and np.array_equal(vid, z[f"apri_vid{f}"])
        vox = s.frame_voxels(f)
        assert np.array_equal(vox["voxel_idx"], z[f"vox_vid{f}"])
        assert np.allclose(vox["av"], z[f"vox_av{f}"], rtol=0, atol=DESC_ATOL) and np.allclose(vox["cov"], z[f"vox_cov{f}"], rtol=0, atol=DESC_ATOL)
        for st in range(3):
            assert np.array_equal(s.frame_point_cluster(f, st), z[f"names{f}_{st}"])
        assert np.array_equal(labels[f], z[f"labels{f}"])
    s.close()


def test_initialization_golden_fixture_through_gpu(pkg):
    z = np.load(os.path.join(GOLD, "init_small.npz"))
    s = pkg.SSC(pkg.semantickitti_params(), device=0, max_points=16 * 450, max_batch=8)
    s.process([z[f"xyzi{k}"] for k in range(6)])
    assert s.intialization(z["poses"]) == int(z["base"])
    cl = s.frame_clusters(pkg.INIT_FRAME)
    for key in ("name", "type", "npts", "nvox"):
        assert np.array_equal(cl[key], z["cl_" + key]), key
    assert np.array_equal(cl["bbox"].view(np.uint32), z["cl_bbox"].view(np.uint32))
    assert np.array_equal(s.frame_voxels(pkg.INIT_FRAME)["label"], z["vox_label"])
    s.close()


# ---- size-independent properties at the benchmark size ----------------------------------------------------
def test_full_size_batch_properties(pkg):
    n = 16
    scans, poses = zip(*[pkg.synth_scan(conftest.SEED + 4, k) for k in range(n)])
    poses = np.stack(poses)
    s = pkg.SSC(pkg.semantickitti_params(), device=0, max_points=64 * 1800, max_batch=16)
    labels = s.segDF(scans, poses)
    for f in range(n):
        c = s.frame_counts(f)
        lab = labels[f]
        hist = np.bincount(lab, minlength=8)
        assert hist.sum() == len(scans[f]) == c[0]
        assert hist[pkg.PT_GROUND] == c[1]
        assert hist[pkg.PT_GATED_OUT] + hist[pkg.PT_UNCLUSTERED] + hist[pkg.PT_STATIC] + hist[pkg.PT_DYNAMIC] == c[2]
        assert hist[pkg.PT_UNCLUSTERED] + hist[pkg.PT_STATIC] + hist[pkg.PT_DYNAMIC] == c[3]
        g, ng = s.frame_ground_order(f)
        assert len(np.unique(np.concatenate([g, ng]))) == c[1] + c[2]
        vox = s.frame_voxels(f)
        assert np.all(np.diff(vox["voxel_idx"]) > 0) and vox["count"].sum() == c[3]  # sortedness + checksum
    assert np.all(labels[-1] != pkg.PT_DYNAMIC)  # the last frame is never tracked (ssc.cpp:1450)
    # idempotence / batch invariance: same scans one at a time give identical classes
    t = pkg.SSC(pkg.semantickitti_params(), device=0, max_points=64 * 1800, max_batch=1)
    l2 = t.segDF(scans[:4], poses[:4])
    for f in range(3):
        assert np.array_equal(l2[f], labels[f])
    # static submap: default = the instance map (points of non-dynamic clusters, ssc.cpp:531-555) ...
    import torch
    total = sum(len(x) for x in scans)
    out = torch.empty((total, 4), dtype=torch.float32, device="cuda:0")
    cnt = s.static_submap_device(0, n, poses, out.data_ptr(), total)
    assert cnt == sum(int((l == pkg.PT_STATIC).sum()) for l in labels)
    inst = out[:cnt].cpu().numpy().copy()
    # ... or, as an option, one point per non-dynamic input point
    s.set_option("submap_all_static", 1)
    cnt = s.static_submap_device(0, n, poses, out.data_ptr(), total)
    assert cnt == sum(int((l != pkg.PT_DYNAMIC).sum()) for l in labels)
    assert torch.isfinite(out[:cnt]).all()
    # ... moved to the map frame with transformCloud's arithmetic (utility.h:394-406: left to right, float, no FMA);
    # the submap is a concatenation without a defined order, so compare as multisets
    exp = []
    for f in range(n):
        T = pkg.pose_matrix(poses[f]).astype(np.float32)
        p = scans[f][labels[f] != pkg.PT_DYNAMIC].astype(np.float32)
        q = np.empty_like(p)
        for r in range(3):
            q[:, r] = ((T[r, 0] * p[:, 0] + T[r, 1] * p[:, 1]) + T[r, 2] * p[:, 2]) + T[r, 3]
        q[:, 3] = p[:, 3]
        exp.append(q)
    exp = np.concatenate(exp).view(np.uint32)
    got = out[:cnt].cpu().numpy().view(np.uint32)
    key = lambda a: a[np.lexsort((a[:, 3], a[:, 2], a[:, 1], a[:, 0]))]
    assert np.array_equal(key(exp), key(got))
    # the instance map is the subset of it that belongs to clusters
    exp_i = []
    for f in range(n):
        T = pkg.pose_matrix(poses[f]).astype(np.float32)
        p = scans[f][labels[f] == pkg.PT_STATIC].astype(np.float32)
        q = np.empty_like(p)
        for r in range(3):
            q[:, r] = ((T[r, 0] * p[:, 0] + T[r, 1] * p[:, 1]) + T[r, 2] * p[:, 2]) + T[r, 3]
        q[:, 3] = p[:, 3]
        exp_i.append(q)
    assert np.array_equal(key(np.concatenate(exp_i).view(np.uint32)), key(inst.view(np.uint32)))
    s.close()
    t.close()


def test_failing_scan_is_named_and_the_rest_can_be_pushed(pkg, oracle):
    """Per-scan error isolation: a scan that exceeds a per-scan capacity fails the call with its index (scvod_last_failed_scan);
    earlier batches of the call stay, and pushing the remaining scans gives the same result as never having seen the bad one."""
    P = pkg.semantickitti_params()
    good = [pkg.synth_scan(conftest.SEED + 5, k, rings=32, cols=900)[0] for k in range(5)]
    big = pkg.synth_scan(conftest.SEED + 5, 9, rings=64, cols=1800)[0]
    s = pkg.SSC(P, device=0, max_points=32 * 900, max_batch=2)
    with pytest.raises(pkg.ScvodError) as e:
        s.process(good[:3] + [big] + good[3:])  # batches: [g0 g1] [g2] (BIG does not fit next to it), then BIG alone fails
    assert "scan 3" in str(e.value)
    assert s.last_failed_scan == 3
    n_kept = s.num_frames
    assert n_kept == 3  # the batches before the failing scan were committed
    s.process(good[n_kept:])
    assert s.last_failed_scan == -1 and s.num_frames == 5
    for f, g in enumerate(good):
        oracle.push_scan(g)
    for f in range(5):
        assert np.array_equal(s.frame_apri(f)[1], oracle.apri(f)[1])
        for st in range(3):
            assert np.array_equal(s.frame_point_cluster(f, st), oracle.point_cluster(f, st))
    s.close()


def test_concurrent_contexts_give_the_sequential_result(pkg):
    """bench.py runs up to 16 contexts (one per host thread and CUDA stream) on one GPU: results must not depend on it."""
    import threading

    nctx, nfr = 4, 8
    chunks = []
    for w in range(nctx):
        scans, poses = zip(*[pkg.synth_scan(conftest.SEED + 20 + w, k, rings=32, cols=900) for k in range(nfr)])
        chunks.append((scans, np.stack(poses)))
    ref = []
    one = pkg.SSC(pkg.semantickitti_params(), device=0, max_points=32 * 900, max_batch=nfr)
    for scans, poses in chunks:
        one.reset()
        ref.append([l.copy() for l in one.segDF(scans, poses)])
    one.close()
    ctxs = [pkg.SSC(pkg.semantickitti_params(), device=0, max_points=32 * 900, max_batch=nfr) for _ in range(nctx)]
    out, errs = [None] * nctx, []

    def work(w):
        try:
            for _ in range(3):  # several rounds: buffers are reused, kernels of different contexts overlap
                ctxs[w].reset()
                out[w] = [l.copy() for l in ctxs[w].segDF(*chunks[w])]
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    threads = [threading.Thread(target=work, args=(w,)) for w in range(nctx)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errs, errs
    for w in range(nctx):
        for f in range(nfr):
            assert np.array_equal(out[w][f], ref[w][f]), (w, f)
    for c in ctxs:
        c.close()


def test_prefetched_upload_gives_the_same_labels(pkg):
    """scvod_prefetch_scans (double-buffered host -> device upload) only changes where the points come from."""
    import torch

    nfr = 6
    chunks = []
    for w in range(3):
        scans, poses = zip(*[pkg.synth_scan(conftest.SEED + 30 + w, k, rings=32, cols=900) for k in range(nfr)])
        off = np.zeros(nfr + 1, np.int64)
        off[1:] = np.cumsum([len(x) for x in scans])
        chunks.append((torch.from_numpy(np.concatenate(scans)).pin_memory(), off, np.stack(poses), scans))
    s = pkg.SSC(pkg.semantickitti_params(), device=0, max_points=32 * 900, max_batch=nfr)
    ref = []
    for host, off, poses, scans in chunks:
        s.reset()
        ref.append([l.copy() for l in s.segDF(scans, poses)])
    got = []
    s.reset()
    s.prefetch_host_ptr(chunks[0][0].data_ptr(), chunks[0][1])
    for i, (host, off, poses, scans) in enumerate(chunks):
        s.reset()
        s.process_host_ptr(host.data_ptr(), off)  # consumes the prefetched copy
        if i + 1 < len(chunks):
            s.prefetch_host_ptr(chunks[i + 1][0].data_ptr(), chunks[i + 1][1])  # overlaps the tracking below
        s.tracking(poses)
        got.append([s.frame_labels(f) for f in range(nfr)])
    for a, b in zip(ref, got):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    # a prefetch that is NOT followed by a push of the same buffer is simply ignored
    s.reset()
    s.prefetch_host_ptr(chunks[0][0].data_ptr(), chunks[0][1])
    s.process_host_ptr(chunks[1][0].data_ptr(), chunks[1][1])
    s.tracking(chunks[1][2])
    for f in range(nfr):
        assert np.array_equal(s.frame_labels(f), ref[1][f])
    s.close()


def tainted_sequence(pkg, seed, nframes, rings, cols, every=1):
    """Synthetic frames turned so that objects sit on the +x axis, with y == 0 points inside them (conftest.taint_scan)."""
    raw = [pkg.synth_scan(seed, k, rings=rings, cols=cols) for k in range(nframes)]
    turn = conftest.densest_object_direction(raw[0][0])
    scans, poses, nq = [], [], 0
    for k, (s, p) in enumerate(raw):
        if k % every == 0:
            s, p, n = conftest.taint_scan(s, p, turn=turn)
            nq += n
        else:
            s, p, _ = conftest.taint_scan(s, p, turn=turn, frac=0.0)
        scans.append(s)
        poses.append(p)
    return scans, np.stack(poses), nq


def check_sequence_against_oracle(pkg, s, orc, scans, poses):
    s.process(scans)
    for sc in scans:
        orc.push_scan(sc)
    compare_frames(s, orc, len(scans), pkg)
    s.tracking(poses)
    orc.track(poses)
    for f in range(len(scans)):
        assert np.array_equal(s.frame_labels(f), orc.labels(f)), f"per-point classes of frame {f}"
        cg, co = s.frame_clusters(f), orc.clusters(f)
        for k in ("name", "type", "state", "npts", "nvox"):
            assert np.array_equal(cg[k], co[k]), (f, k)


@pytest.mark.parametrize("config,seed_off,rings,cols,nframes", [("semantickitti", 60, 64, 1800, 6), ("parkinglot", 61, 64, 1800, 4),
                                                               ("semantickitti", 62, 32, 900, 12), ("parkinglot", 63, 32, 900, 10)])
def test_minus_one_sector_points_inside_clusters(pkg, config, seed_off, rings, cols, nframes):
    """y == 0 exactly with x > 0 gives angle 0 and sector_idx -1 (ssc.cpp:186): the point hashes into another cell's voxel and
    clusterAndCreateFrame has to be replayed point by point for that voxel (ssc.cpp:299-393).  KITTI .bin coordinates are
    quantised, so such rows occur in real scans.  Everything the reference computes is reproduced: voxel indices, descriptor,
    names after the three clustering stages, clusters, tracking states and per-point classes."""
    params = getattr(pkg, config + "_params")()
    scans, poses, nq = tainted_sequence(pkg, conftest.SEED + seed_off, nframes, rings, cols)
    assert nq >= 10 * nframes
    s = pkg.SSC(params, device=0, max_points=rings * cols, max_batch=8)
    orc = conftest.Oracle(params)
    check_sequence_against_oracle(pkg, s, orc, scans, poses)
    assert s.stat("tainted_voxels") > 0
    s.close()
    orc.close()


def test_minus_one_range_and_azimuth_points_inside_clusters(pkg):
    """dis == min_dis (range_idx -1) and azimuth == min_azimuth (azimuth_idx -1, a NEGATIVE voxel_idx), together with
    y == 0 rows; a batch that mixes such scans with ordinary ones."""
    kp = pkg.semantickitti_params()
    scans, poses, _ = tainted_sequence(pkg, conftest.SEED + 64, 6, 32, 900, every=2)
    params, scans[0] = conftest.params_with_edge_points(pkg, kp, conftest.Oracle, scans[0])
    orc = conftest.Oracle(params)
    b = orc.bin(scans[0])
    assert ((b["pass"] != 0) & (b["range_idx"] == -1)).sum() >= 1 and ((b["pass"] != 0) & (b["azimuth_idx"] == -1)).sum() >= 1
    s = pkg.SSC(params, device=0, max_points=32 * 900, max_batch=8)
    check_sequence_against_oracle(pkg, s, orc, scans, poses)
    s.close()
    orc.close()


def test_minus_one_points_with_global_replay_and_initialization(pkg):
    """The global-memory variant of the name replay and SSC::intialization on frames that hold aliased voxels."""
    params = pkg.semantickitti_params()
    scans, poses, _ = tainted_sequence(pkg, conftest.SEED + 65, 6, 32, 900)
    s = pkg.SSC(params, device=0, max_points=32 * 900, max_batch=8)
    s.set_option("replay_global", 1)
    orc = conftest.Oracle(params)
    s.process(scans)
    for sc in scans:
        orc.push_scan(sc)
    compare_frames(s, orc, len(scans), pkg)
    assert s.intialization(poses) == orc.initialization(poses)
    cg, co = s.frame_clusters(pkg.INIT_FRAME), orc.clusters(-1)
    for k in ("name", "type", "npts", "nvox"):
        assert np.array_equal(cg[k], co[k]), k
    assert np.array_equal(s.frame_voxels(pkg.INIT_FRAME)["label"], orc.voxels(-1)["label"])
    s.tracking(poses)
    orc.track(poses)
    for f in range(len(scans)):
        assert np.array_equal(s.frame_labels(f), orc.labels(f)), f
    s.close()
    orc.close()


def test_aliased_golden_fixture_through_gpu(pkg):
    z = np.load(os.path.join(GOLD, "aliased_small.npz"))
    s = pkg.SSC(pkg.semantickitti_params(), device=0, max_points=16 * 450, max_batch=4)
    labels = s.segDF([z[f"xyzi{k}"] for k in range(4)], z["poses"])
    for k in range(4):
        src, vid = s.frame_apri(k)
        assert np.array_equal(src, z[f"apri_src{k}"]) and np.array_equal(vid, z[f"apri_vid{k}"])
        for st in range(3):
            assert np.array_equal(s.frame_point_cluster(k, st), z[f"names{k}_{st}"])
        assert np.array_equal(labels[k], z[f"labels{k}"])
        assert np.array_equal(s.frame_clusters(k)["state"], z[f"cl_state{k}"])
    assert s.stat("tainted_points") > 0
    s.close()


def test_benchmarked_chunk_matches_oracle(pkg):
    """The exact shape bench.py times: one 64-frame chunk of 64x1800 scans (its batches[0]: generator seed SEED, scans 0..63,
    carried clouds growing over 63 tracked pairs).  Per-point classes, cluster names and states of every frame equal the oracle's."""
    n = 64
    scans, poses = zip(*[pkg.synth_scan(conftest.SEED, k) for k in range(n)])
    poses = np.stack(poses)
    s = pkg.SSC(pkg.semantickitti_params(), device=0, max_points=64 * 1800, max_batch=n)
    labels = s.segDF(scans, poses)
    orc = conftest.Oracle(pkg.semantickitti_params())
    secs, olab, off = orc.run_sequence(scans, poses, nthreads=8)
    ndyn = 0
    for f in range(n):
        assert np.array_equal(labels[f], olab[off[f]:off[f + 1]]), f"frame {f}"
        ndyn += int((labels[f] == pkg.PT_DYNAMIC).sum())
    assert ndyn > 0 and s.stat("track_pairs") == n - 1
    # cluster tables of a few frames deep in the chain (the step-by-step oracle keeps them)
    for sc in scans[:24]:
        orc.push_scan(sc)
    orc.track(poses[:24])
    t = pkg.SSC(pkg.semantickitti_params(), device=0, max_points=64 * 1800, max_batch=24)
    t.segDF(scans[:24], poses[:24])
    for f in (5, 14, 22, 23):
        cg, co = t.frame_clusters(f), orc.clusters(f)
        for k in ("name", "type", "state", "npts", "nvox"):
            assert np.array_equal(cg[k], co[k]), (f, k)
    s.close()
    t.close()
    orc.close()


def test_chunked_sequence_with_tail_handoff_is_one_chain(pkg, oracle):
    """A 40-frame sequence cut into chunks of 7 frames held by different contexts: with the tail hand-off (scvod_export_tail ->
    scvod_track_from_tail -> scvod_apply_tail_states) labels, cluster states and track ids equal the oracle's ONE unbroken chain
    (ssc.cpp:1450-1452); without it, every cut loses the labels of one frame pair."""
    par = conftest.load_parallel()
    n, chunk = 40, 7
    scans, poses = zip(*[pkg.synth_scan(conftest.SEED + 1, k, rings=32, cols=900) for k in range(n)])
    poses = np.stack(poses)
    for sc in scans:
        oracle.push_scan(sc)
    oracle.track(poses)
    cuts = par.chunk_sequence(n, chunk)
    ctxs = [pkg.SSC(pkg.semantickitti_params(), device=0, max_points=32 * 900, max_batch=chunk) for _ in cuts]
    for (a, b), s in zip(cuts, ctxs):
        s.process(scans[a:b])
    par.track_chunks_as_one_chain(ctxs, [poses[a:b] for a, b in cuts])
    for (a, b), s in zip(cuts, ctxs):
        for f in range(a, b):
            assert np.array_equal(s.frame_labels(f - a), oracle.labels(f)), f"frame {f}"
            cg, co = s.frame_clusters(f - a), oracle.clusters(f)
            for k in ("name", "type", "state", "npts", "nvox"):
                assert np.array_equal(cg[k], co[k]), (f, k)
    # the cut chain, for the record: only frames next to a cut may differ, and at least one does in this sequence
    differ = []
    for (a, b), s in zip(cuts, ctxs):
        s.reset()
        lab = s.segDF(scans[a:b], poses[a:b])
        differ += [f for f in range(a, b) if not np.array_equal(lab[f - a], oracle.labels(f))]
    assert len(differ) > 0
    for s in ctxs:
        s.close()


def test_tail_handoff_with_aliased_voxels_and_errors(pkg):
    params = pkg.semantickitti_params()
    scans, poses, _ = tainted_sequence(pkg, conftest.SEED + 62, 9, 32, 900)
    orc = conftest.Oracle(params)
    for sc in scans:
        orc.push_scan(sc)
    orc.track(poses)
    par = conftest.load_parallel()
    a = pkg.SSC(params, device=0, max_points=32 * 900, max_batch=8)
    b = pkg.SSC(params, device=0, max_points=32 * 900, max_batch=8)
    a.process(scans[:5])
    b.process(scans[5:])
    with pytest.raises(pkg.ScvodError):
        a.export_tail()  # its own frames are not tracked yet
    par.track_chunks_as_one_chain([a, b], [poses[:5], poses[5:]])
    with pytest.raises(pkg.ScvodError):
        b.track_from_tail(a.export_tail(), poses[4], poses[5])  # head already tracked
    for f in range(9):
        s, g = (a, f) if f < 5 else (b, f - 5)
        assert np.array_equal(s.frame_labels(g), orc.labels(f)), f
    a.close()
    b.close()
    orc.close()
