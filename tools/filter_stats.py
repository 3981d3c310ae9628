"""Binning-filter statistics quoted in DESIGN.md: share of points that take the exact chain on synthetic scans, and the largest distance
between the approximate and the exact bin coordinates over generated points (scvod_bin_filter_check)."""
import os, sys, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import conftest
pkg = conftest.load_package()
for name, P in (("semantickitti", pkg.semantickitti_params()), ("parkinglot", pkg.parkinglot_params())):
    s = pkg.SSC(P, device=0, max_points=65536, max_batch=1)
    scans = np.concatenate([pkg.synth_scan(conftest.SEED, k)[0] for k in range(4)])
    st = s.bin_filter_check(0, cloud=scans)
    g = [s.bin_filter_check(1 << 30, seed=sd, extent=ex) for sd, ex in ((1, 60.0), (77, 35.0))]
    print(name, "synthetic scans: exact-chain share %.4f %% of %d points, mismatches %d;" % (100.0 * st["exact"] / st["points"], st["points"], st["mismatches"]),
          "generated 2 x 2^30: max dq sector %.2e azimuth %.2e, patch-filter exact share %.3f %%, mismatches %d / %d" % (
              max(x["max_dq_sector"] for x in g), max(x["max_dq_azimuth"] for x in g), 100.0 * g[0]["patch_exact"] / g[0]["points"],
              sum(x["mismatches"] for x in g), sum(x["patch_mismatches"] for x in g)))
    s.close()
