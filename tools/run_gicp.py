"""One dense scan-to-map GICP alignment (configs[2]/[4] shape) — a short command for ncu captures and timing."""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import conftest
pkg = conftest.load_package()
rings = int(sys.argv[1]) if len(sys.argv) > 1 else 128
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 2700
nmap = int(sys.argv[3]) if len(sys.argv) > 3 else 2
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
scans = [pkg.synth_scan(conftest.SEED + 5, k, rings=rings, cols=cols) for k in range(nmap + 1)]
def to_world(s, pose):
    T = pkg.pose_matrix(pose); out = s.copy(); out[:, :3] = s[:, :3] @ T[:, :3].T + T[:, 3]; return out
tgt = np.concatenate([to_world(s, p) for s, p in scans[:nmap]])
src, pose = scans[nmap]
pert = pose.copy(); pert[0] += 0.15; pert[1] -= 0.1; pert[5] += np.deg2rad(1.5)
s = pkg.SSC(pkg.parkinglot_params(), device=0, max_points=4096, max_batch=1)
gp = pkg.gicp_default_params()
pkg.kernel_timing(True)
for _ in range(reps):
    t0 = time.time(); s.gicp_set_target(tgt, gp); t1 = time.time()
    r = s.gicp_align(src, pkg.pose_matrix(pert)); t2 = time.time()
    print(f"src {len(src)} tgt {len(tgt)} set_target {1e3*(t1-t0):.1f} ms align {1e3*(t2-t1):.1f} ms iters {r['iterations']} conv {r['converged']} ncorr {r['n_corr']} pose {r['pose6']} true {pose}")
rep = pkg.kernel_timing_report()
for k, v in sorted(rep.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:28s} {v[0]:9.3f} ms {v[1]:5d} launches")
