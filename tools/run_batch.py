"""Push one batch of synthetic scans (+ tracking) — a short command for ncu captures."""
import os, sys, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import conftest
pkg = conftest.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
from concurrent.futures import ThreadPoolExecutor
with ThreadPoolExecutor(16) as ex:
    res = list(ex.map(lambda k: pkg.synth_scan(conftest.SEED, k), range(n)))
scans = [r[0] for r in res]; poses = np.stack([r[1] for r in res])
s = pkg.SSC(pkg.semantickitti_params(), device=0, max_points=64 * 1800, max_batch=n)
s.set_option("inspect", 0)
if os.environ.get("SCVOD_CHAIN_TMA"):
    s.set_option("chain_tma", 1)  # k_patch_chain<true>: ring staged with cp.async.bulk + mbarrier (UBLKCP) instead of per-lane cp.async
for _ in range(reps):
    s.reset(); s.process(scans); s.tracking(poses); s.refresh_labels(0, n)
print("done", s.kernel_launches)
