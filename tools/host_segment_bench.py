#!/usr/bin/env python
"""CPU-only measurement of the host-side cluster bookkeeping (csrc/host_cluster.cpp): builds the voxel tables of one synthetic
64x1800 scan from the oracle (as tests/test_host_logic.py does), gives the product's segment_and_recognize device-style names, and
times it with tools/probe/bench_segment.cpp.  The printed hash covers every cluster and label: it must not change when the
bookkeeping is optimised.  usage: host_segment_bench.py [reps=20000] [--prof]"""
import os, subprocess, sys, ctypes
import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest
import test_host_logic as T

reps = next((a for a in sys.argv[1:] if a.isdigit()), "20000")
prof = "--prof" in sys.argv
pkg = conftest.load_package()
params = pkg.semantickitti_params()
orc = conftest.Oracle(params)
scan, _ = pkg.synth_scan(conftest.SEED, 3, rings=64, cols=1800)
orc.push_scan(scan)
src, vid = orc.apri(0)
vox = orc.voxels(0)
cid, nbr, root, ev, bbox, edges = T.build_tables(params, orc.grid_dims()[:3], scan, src, vid, vox)
cnt = np.ascontiguousarray(vox["count"], np.int32)
V = len(cnt)
# stage-0 names through the host replay hook, then the first event of every name (what k_name_replay hands over)
names = [np.zeros(V, np.int32) for _ in range(3)]
ncl = np.zeros(3, np.int32); cname = np.zeros(4096, np.int32); ctype = np.zeros(4096, np.int32); mx = ctypes.c_int32()
P = conftest._ptr
lib = pkg.load_library()
n = lib.scvod_host_segment(ctypes.byref(params), V, P(cnt), P(root), P(np.ascontiguousarray(nbr)), P(bbox), len(ev), P(ev), len(edges),
                           P(np.ascontiguousarray(edges)), P(names[0]), P(names[1]), P(names[2]), P(ncl), 4096, P(cname), P(ctype), ctypes.byref(mx))
assert n >= 0
maxn = int(names[0].max())
nf = np.full(maxn + 7, 0x7FFFFFFF, np.int32)
for e, v in enumerate(ev):
    if nf[names[0][v]] == 0x7FFFFFFF:
        nf[names[0][v]] = e
tables = "/tmp/scvod_segment_tables.bin"
with open(tables, "wb") as f:
    f.write(np.array([V, len(ev), len(edges), maxn, len(nf)], np.int32).tobytes())
    for a in (cnt, root.astype(np.int32), nbr.astype(np.int32), bbox.astype(np.float32), ev.astype(np.int32), edges.astype(np.int32).reshape(-1), names[0], nf):
        f.write(np.ascontiguousarray(a).tobytes())
csrc = os.path.join(ROOT, "dr-using-scv-od_b200", "csrc")
exe = "/tmp/scvod_bench_segment"
subprocess.check_call(["g++", "-O2", "-std=gnu++17", "-ffp-contract=off"] + (["-DSEG_PROF"] if prof else []) + ["-I", csrc,
                       os.path.join(ROOT, "tools", "probe", "bench_segment.cpp"), os.path.join(csrc, "host_cluster.cpp"), os.path.join(csrc, "scvod_params.cpp"), "-o", exe])
print(f"V={V} events={len(ev)} edges={len(edges)} clusters per stage={list(map(int, ncl))}")
subprocess.check_call([exe, reps, tables])
