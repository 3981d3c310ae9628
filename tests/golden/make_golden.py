"""Generates the fixtures in this directory.

The reference ships no tests or golden vectors for this path and cannot be built or run here
(SURVEY.md §4, §8c), so these fixtures are produced by the CPU oracle (oracle/scvod_oracle.cpp) on
inputs from the deterministic generator: they pin the oracle (regression) and give the GPU path a
fixed input/output pair that does not depend on the generator.  Independent known answers (grid
dimensions and index quirks verified in SURVEY.md §7/§8) live in tests/test_oracle_golden.py.

Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import conftest  # noqa: E402


def edge_case_points(p):
    """Crafted points for the index quirks of SSC::makeApriVec (SURVEY.md hard part 7)."""
    pts = [
        (5.0, 0.0, -1.0),      # y == 0, x > 0 -> angle 0 -> sector_idx -1
        (0.0, 0.0, -1.0),      # x == y == 0 -> angle 0, dis 0 -> gated out
        (-5.0, 0.0, -1.0),     # angle 180
        (0.0, 5.0, -1.0),      # angle 90
        (0.0, -5.0, -1.0),     # angle 270
        (5.0, -1e-30, -1.0),   # angle just below 360
        (p.min_dis, 0.0, 0.0),  # dis == min_dis -> range_idx -1 (and angle 0)
        (0.0, p.min_dis, 0.0),  # dis == min_dis, angle 90
        (p.max_dis, 1e-3, 0.0),
        (0.0, p.max_dis, 0.0),  # dis == max_dis passes the gate
        (0.0, np.nextafter(np.float32(p.max_dis), np.float32(1e9)), 0.0),  # just outside
        (3.0, 3.0, 30.0), (3.0, 3.0, -30.0), (10.0, -7.0, 0.5), (-12.0, 9.0, -1.6), (29.9, 0.1, 2.0),
    ]
    rng = np.random.default_rng(7)
    rnd = rng.uniform(-35, 35, size=(4000, 3)).astype(np.float32)
    rnd[:, 2] = rng.uniform(-3, 8, size=4000).astype(np.float32)
    arr = np.concatenate([np.array(pts, np.float32), rnd], axis=0)
    return np.concatenate([arr, np.full((len(arr), 1), 50.0, np.float32)], axis=1)


def main():
    pkg = conftest.load_package()
    P = pkg.semantickitti_params()
    orc = conftest.Oracle(P)

    edge = edge_case_points(P)
    b = orc.bin(edge)
    np.savez_compressed(os.path.join(HERE, "bin_edge_cases.npz"), xyzi=edge, **{"o_" + k: v for k, v in b.items()})

    scans, poses = [], []
    for k in range(3):
        s, pose = pkg.synth_scan(conftest.SEED, k, rings=16, cols=450)
        scans.append(s)
        poses.append(pose)
    poses = np.stack(poses)
    for s in scans:
        orc.push_scan(s)
    out = {"poses": poses}
    for f, s in enumerate(scans):
        g, ng = orc.ground_order(f)
        src, vid = orc.apri(f)
        vox = orc.voxels(f)
        out.update({f"xyzi{f}": s, f"ground{f}": g, f"nonground{f}": ng, f"apri_src{f}": src, f"apri_vid{f}": vid,
                    f"vox_vid{f}": vox["voxel_idx"], f"vox_cnt{f}": vox["count"], f"vox_av{f}": vox["av"], f"vox_cov{f}": vox["cov"],
                    f"counts{f}": orc.counts(f)})
        for st in range(3):
            out[f"names{f}_{st}"] = orc.point_cluster(f, st)
    orc.track(poses)
    for f in range(len(scans)):
        out[f"labels{f}"] = orc.labels(f)
    np.savez_compressed(os.path.join(HERE, "scan_small.npz"), **out)

    # SSC::intialization (ssc.cpp:1148-1248): six small scans whose base frame gets two fusions
    orc2 = conftest.Oracle(P)
    init = {}
    iposes = []
    for k in range(6):
        s, pose = pkg.synth_scan(conftest.SEED, k, rings=16, cols=450)
        init[f"xyzi{k}"] = s
        iposes.append(pose)
        orc2.push_scan(s)
    iposes = np.stack(iposes)
    init["poses"] = iposes
    init["base"] = np.int32(orc2.initialization(iposes))
    cl = orc2.clusters(-1)
    for key in ("name", "type", "npts", "nvox", "bbox"):
        init["cl_" + key] = cl[key]
    init["vox_label"] = orc2.voxels(-1)["label"]
    init["n_clusters_before"] = np.int32(len(orc2.clusters(int(init["base"]))["name"]))
    np.savez_compressed(os.path.join(HERE, "init_small.npz"), **init)
    orc2.close()

    # scans with y == 0 rows inside objects (sector_idx -1: the point hashes into another cell's voxel, ssc.cpp:186-188), which
    # clusterAndCreateFrame names point by point; labels after tracking
    orc3 = conftest.Oracle(P)
    raw = [pkg.synth_scan(conftest.SEED + 60, k, rings=16, cols=450) for k in range(4)]
    turn = conftest.densest_object_direction(raw[0][0])
    al = {}
    aposes = []
    nq = 0
    for k, (s, pose) in enumerate(raw):
        s, pose, n = conftest.taint_scan(s, pose, turn=turn)
        nq += n
        al[f"xyzi{k}"] = s
        aposes.append(pose)
        orc3.push_scan(s)
    assert nq > 20
    aposes = np.stack(aposes)
    al["poses"] = aposes
    for k in range(4):
        src, vid = orc3.apri(k)
        al[f"apri_src{k}"], al[f"apri_vid{k}"] = src, vid
        al[f"vox_label{k}"] = orc3.voxels(k)["label"]
        for st in range(3):
            al[f"names{k}_{st}"] = orc3.point_cluster(k, st)
    orc3.track(aposes)
    for k in range(4):
        al[f"labels{k}"] = orc3.labels(k)
        al[f"cl_state{k}"] = orc3.clusters(k)["state"]
    np.savez_compressed(os.path.join(HERE, "aliased_small.npz"), **al)
    orc3.close()

    hashes = {}
    for k in range(3):
        s, pose = pkg.synth_scan(conftest.SEED, k)
        hashes[str(k)] = {"n": int(len(s)), "sha256": hashlib.sha256(s.tobytes()).hexdigest(), "pose": [float(x) for x in pose]}
    json.dump(hashes, open(os.path.join(HERE, "synth_hash.json"), "w"), indent=1)
    print("written", os.listdir(HERE))


if __name__ == "__main__":
    main()
