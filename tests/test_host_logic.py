"""CPU suite: the product's host-side cluster bookkeeping (csrc/host_cluster.cpp) against the oracle.

The voxel tables the GPU stages would produce (27-neighbour lists in findVoxelNeighbors order, connected
components, similarity edges, clustering events, per-voxel bounding boxes) are rebuilt here with numpy
from the oracle's descriptor, then fed through the host-only hook scvod_host_segment.  Cluster names
after each of the three stages, cluster_set iteration order and car/non-car types must be identical.
"""
import ctypes

import numpy as np
import pytest
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import connected_components

import conftest


def build_tables(params, grid, xyzi, apri_src, apri_vid, vox):
    R, S, A = grid
    vids = vox["voxel_idx"]
    V = len(vids)
    cid = np.searchsorted(vids, apri_vid)
    tri = vox["tri"].astype(np.int64)
    lookup = {int(v): i for i, v in enumerate(vids)}
    nbr = np.full((V, 27), -1, np.int32)
    for v in range(V):
        ri, si, ei = tri[v]
        t = 0
        for x in range(ri - 1, ri + 2):
            for y in range(si - 1, si + 2):
                for z in range(ei - 1, ei + 2):
                    if 0 <= x <= R - 1 and 0 <= y <= S - 1 and 0 <= z <= A - 1:
                        nbr[v, t] = lookup.get(x * S + y + z * R * S, -1)
                    t += 1
    rows, cols = np.nonzero(nbr >= 0)
    ncomp, comp = connected_components(coo_matrix((np.ones(len(rows)), (rows, nbr[rows, cols])), shape=(V, V)), directed=False)
    root_of_comp = np.full(ncomp, V, np.int64)
    np.minimum.at(root_of_comp, comp, np.arange(V))
    root = root_of_comp[comp].astype(np.int32)
    # events: first three points of every voxel, in apri order
    seen = np.zeros(V, np.int32)
    ev = []
    for c in cid:
        if seen[c] < 3:
            ev.append(c)
        seen[c] += 1
    ev = np.array(ev, np.int32)
    # per-voxel bbox
    pts = xyzi[apri_src][:, :3]
    bbox = np.zeros((V, 6), np.float32)
    lo = np.full((V, 3), np.inf, np.float32)
    hi = np.full((V, 3), -np.inf, np.float32)
    np.minimum.at(lo, cid, pts)
    np.maximum.at(hi, cid, pts)
    bbox[:, :3], bbox[:, 3:] = lo, hi
    # similarity edges (ssc.cpp:587-595), directed, component level
    edges = set()
    av, cov = vox["av"], vox["cov"]
    for v in range(V):
        ri, si, ei = tri[v]
        size = 1 if ri > R * 0.6 else params.search_c
        for x in range(ri - size, ri + size + 1):
            if x > R - 1 or x < 0:
                continue
            for y in range(si - size, si + size + 1):
                if y > S - 1 or y < 0:
                    continue
                for z in range(ei - size, ei + size + 1):
                    if z > A - 1 or z < 0:
                        continue
                    u = lookup.get(x * S + y + z * R * S)
                    if u is None:
                        continue
                    if cov[u] <= np.float32(params.intensity_cov) and np.abs(np.float32(av[v] - av[u])) <= np.float32(params.intensity_diff):
                        edges.add((int(root[v]), int(root[u])))
    edges = np.array(sorted(edges), np.int32).reshape(-1, 2)
    return cid, nbr, root, ev, bbox, edges


@pytest.mark.parametrize("scan_id,rings,cols", [(0, 32, 900), (7, 32, 900), (3, 64, 1800)])
def test_host_segment_matches_oracle(pkg, oracle, kitti_params, scan_id, rings, cols):
    s, _ = pkg.synth_scan(conftest.SEED, scan_id, rings=rings, cols=cols)
    oracle.push_scan(s)
    src, vid = oracle.apri(0)
    vox = oracle.voxels(0)
    grid = oracle.grid_dims()[:3]
    cid, nbr, root, ev, bbox, edges = build_tables(kitti_params, grid, s, src, vid, vox)
    V = len(vox["voxel_idx"])
    names = [np.zeros(V, np.int32) for _ in range(3)]
    ncl = np.zeros(3, np.int32)
    cap = 4096
    cname = np.zeros(cap, np.int32)
    ctype = np.zeros(cap, np.int32)
    max_name = ctypes.c_int32()
    cnt = np.ascontiguousarray(vox["count"], np.int32)
    P = conftest._ptr
    lib = pkg.load_library()
    n = lib.scvod_host_segment(ctypes.byref(kitti_params), V, P(cnt), P(root), P(np.ascontiguousarray(nbr)), P(bbox), len(ev), P(ev),
                               len(edges), P(np.ascontiguousarray(edges)), P(names[0]), P(names[1]), P(names[2]), P(ncl), cap, P(cname),
                               P(ctype), ctypes.byref(max_name))
    assert n >= 0, lib.scvod_last_error()
    c = oracle.counts(0)
    assert list(ncl) == [c[5], c[6], c[7]]
    for st in range(3):
        assert np.array_equal(names[st][cid], oracle.point_cluster(0, st)), f"stage {st}"
    oc = oracle.clusters(0)
    assert np.array_equal(cname[:n], oc["name"])  # cluster_set iteration order
    assert np.array_equal(ctype[:n], oc["type"])


def build_taint_tables(grid, xyzi, apri_src, vox, cid, nbr, root, binres):
    """Side tables of the voxels that hold a point with a -1 index ("tainted", include/scvod.h scvod_host_segment_pts)."""
    R, S, A = grid
    V = len(vox["voxel_idx"])
    lookup = {int(v): i for i, v in enumerate(vox["voxel_idx"])}
    ptri = np.stack([binres["range_idx"], binres["sector_idx"], binres["azimuth_idx"]], 1)[apri_src].astype(np.int64)
    quirk = (ptri < 0).any(axis=1)
    tv = np.unique(cid[quirk]).astype(np.int32)
    is_t = np.zeros(V, bool)
    is_t[tv] = True
    tv_base = np.zeros(len(tv) + 1, np.int32)
    tp_m = []
    for i, v in enumerate(tv):
        ms = np.flatnonzero(cid == v)
        tp_m.extend(ms.tolist())
        tv_base[i + 1] = len(tp_m)
    tp_m = np.array(tp_m, np.int32)
    tp_xyz = np.ascontiguousarray(xyzi[apri_src[tp_m]], np.float32)
    tp_nbr = np.full((len(tp_m), 27), -1, np.int32)
    for q, m in enumerate(tp_m):
        ri, si, ei = ptri[m]
        t = 0
        for x in range(ri - 1, ri + 2):
            for y in range(si - 1, si + 2):
                for z in range(ei - 1, ei + 2):
                    if 0 <= x <= R - 1 and 0 <= y <= S - 1 and 0 <= z <= A - 1:
                        tp_nbr[q, t] = lookup.get(x * S + y + z * R * S, -1)
                    t += 1
    # components of the ORDINARY voxels only; a tainted voxel is its own root
    rows, cols = np.nonzero(nbr >= 0)
    keep = ~is_t[rows] & ~is_t[nbr[rows, cols]]
    ncomp, comp = connected_components(coo_matrix((np.ones(int(keep.sum())), (rows[keep], nbr[rows, cols][keep])), shape=(V, V)), directed=False)
    root_of_comp = np.full(ncomp, V, np.int64)
    np.minimum.at(root_of_comp, comp, np.arange(V))
    root2 = root_of_comp[comp].astype(np.int32)
    # events: first three points of an ordinary voxel, every point of a tainted one (node V + position in tp_*)
    pos_of = {int(m): q for q, m in enumerate(tp_m)}
    seen = np.zeros(V, np.int32)
    ev = []
    for m, c in enumerate(cid):
        if is_t[c]:
            ev.append(V + pos_of[m])
        elif seen[c] < 3:
            ev.append(c)
        seen[c] += 1
    return tv, tv_base, tp_m, tp_xyz, tp_nbr, root2, np.array(ev, np.int32)


def run_host_segment_pts(pkg, params, s, orc):
    src, vid = orc.apri(0)
    vox = orc.voxels(0)
    grid = orc.grid_dims()[:3]
    cid, nbr, root, ev, bbox, _ = build_tables(params, grid, s, src, vid, vox)
    tv, tv_base, tp_m, tp_xyz, tp_nbr, root, ev = build_taint_tables(grid, s, src, vox, cid, nbr, root, orc.bin(s))
    # similarity edges between the roots that hold with tainted voxels kept apart
    R, S, A = grid
    lookup = {int(v): i for i, v in enumerate(vox["voxel_idx"])}
    tri = vox["tri"].astype(np.int64)
    av, cov = vox["av"], vox["cov"]
    edges = set()
    for v in range(len(tri)):
        ri, si, ei = tri[v]
        size = 1 if ri > R * 0.6 else params.search_c
        for x in range(ri - size, ri + size + 1):
            if x > R - 1 or x < 0:
                continue
            for y in range(si - size, si + size + 1):
                if y > S - 1 or y < 0:
                    continue
                for z in range(ei - size, ei + size + 1):
                    if z > A - 1 or z < 0:
                        continue
                    u = lookup.get(x * S + y + z * R * S)
                    if u is not None and cov[u] <= np.float32(params.intensity_cov) and np.abs(np.float32(av[v] - av[u])) <= np.float32(params.intensity_diff):
                        edges.add((int(root[v]), int(root[u])))
    edges = np.array(sorted(edges), np.int32).reshape(-1, 2)
    V = len(vox["voxel_idx"])
    names = [np.zeros(V, np.int32) for _ in range(3)]
    tp_stage = np.zeros((3, max(len(tp_m), 1)), np.int32)
    ncl = np.zeros(3, np.int32)
    cap = 4096
    cname, ctype, cnpts, cnvox = (np.zeros(cap, np.int32) for _ in range(4))
    max_name = ctypes.c_int32()
    cnt = np.ascontiguousarray(vox["count"], np.int32)
    P = conftest._ptr
    lib = pkg.load_library()
    n = lib.scvod_host_segment_pts(ctypes.byref(params), V, P(cnt), P(root), P(np.ascontiguousarray(nbr)), P(bbox), len(ev), P(ev), len(edges),
                                   P(np.ascontiguousarray(edges)), len(tv), P(tv), P(tv_base), len(tp_m), P(tp_m), P(tp_xyz),
                                   P(np.ascontiguousarray(tp_nbr)), P(names[0]), P(names[1]), P(names[2]), P(tp_stage), P(ncl), cap, P(cname),
                                   P(ctype), P(cnpts), P(cnvox), ctypes.byref(max_name))
    assert n >= 0, lib.scvod_last_error()
    c = orc.counts(0)
    assert list(ncl) == [c[5], c[6], c[7]]
    for st in range(3):
        got = names[st][cid]
        got[tp_m] = tp_stage[st][: len(tp_m)]  # points of tainted voxels carry their own names
        assert np.array_equal(got, orc.point_cluster(0, st)), f"stage {st}"
        assert np.array_equal(names[st], orc.voxels(0)["label"]) or st < 2
    oc = orc.clusters(0)
    assert np.array_equal(cname[:n], oc["name"])  # cluster_set iteration order
    assert np.array_equal(ctype[:n], oc["type"])
    assert np.array_equal(cnpts[:n], oc["npts"]) and np.array_equal(cnvox[:n], oc["nvox"])
    return len(tv), len(tp_m)


@pytest.mark.parametrize("config,scan_id,rings,cols", [("semantickitti", 0, 32, 900), ("semantickitti", 5, 64, 1800), ("parkinglot", 2, 32, 900),
                                                       ("parkinglot", 9, 64, 1800)])
def test_host_segment_with_minus_one_sector_points(pkg, config, scan_id, rings, cols):
    """y == 0 exactly (x > 0) gives sector_idx -1 (ssc.cpp:186): the point hashes into the voxel of sector 299 one ring closer.
    Names of every point after the three stages, cluster order, types and sizes equal the oracle's."""
    params = getattr(pkg, config + "_params")()
    s, _ = pkg.synth_scan(conftest.SEED + 50, scan_id, rings=rings, cols=cols)
    s, _, nq = conftest.taint_scan(s)
    assert nq > 10
    orc = conftest.Oracle(params)
    orc.push_scan(s)
    ntv, ntp = run_host_segment_pts(pkg, params, s, orc)
    assert ntv > 0 and ntp >= ntv
    orc.close()


@pytest.mark.parametrize("scan_id", [1, 4])
def test_host_segment_with_minus_one_range_and_azimuth_points(pkg, kitti_params, scan_id):
    """dis == min_dis gives range_idx -1, azimuth == min_azimuth gives azimuth_idx -1 (negative voxel_idx)."""
    s, _ = pkg.synth_scan(conftest.SEED + 51, scan_id, rings=32, cols=900)
    s, _, _ = conftest.taint_scan(s, frac=0.5)
    params, s = conftest.params_with_edge_points(pkg, kitti_params, conftest.Oracle, s)
    orc = conftest.Oracle(params)
    orc.push_scan(s)
    b = orc.bin(s)
    assert ((b["pass"] != 0) & (b["range_idx"] == -1)).sum() >= 1 and ((b["pass"] != 0) & (b["azimuth_idx"] == -1)).sum() >= 1
    ntv, _ = run_host_segment_pts(pkg, params, s, orc)
    assert ntv > 0
    orc.close()
