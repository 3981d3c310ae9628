"""Per-kernel CUDA-event times of one chunk of synthetic 64x1800 scans on a single stream (no overlap): ms per chunk, sorted.
usage: kernel_times.py [scans=64] [reps=5]"""
import os, sys, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import conftest
pkg = conftest.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
from concurrent.futures import ThreadPoolExecutor
with ThreadPoolExecutor(16) as ex:
    res = list(ex.map(lambda k: pkg.synth_scan(conftest.SEED, k), range(n)))
scans = [r[0] for r in res]; poses = np.stack([r[1] for r in res])
s = pkg.SSC(pkg.semantickitti_params(), device=0, max_points=64 * 1800, max_batch=n)
s.set_option("inspect", 0)
for _ in range(2):
    s.reset(); s.process(scans); s.tracking(poses); s.refresh_labels(0, n)
pkg.kernel_timing(True)
for _ in range(reps):
    s.reset(); s.process(scans); s.tracking(poses); s.refresh_labels(0, n)
rep = pkg.kernel_timing_report()
pkg.kernel_timing(False)
tot = 0.0
for k, (ms, cnt) in sorted(rep.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:28s} {ms / reps:8.4f} ms/chunk  {cnt // reps:5d} launches")
    tot += ms / reps
print(f"{'total':28s} {tot:8.4f} ms/chunk")
