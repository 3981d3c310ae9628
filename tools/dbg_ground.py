import os, sys, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import conftest
pkg = conftest.load_package()
rings, cols = int(sys.argv[1]), int(sys.argv[2])
s = pkg.SSC(pkg.semantickitti_params(), device=0, max_points=rings * cols, max_batch=1)
cloud, _ = pkg.synth_scan(conftest.SEED, 4, rings=rings, cols=cols)
print('points', len(cloud))
g, ng = s.extractGroudByPatchWork(cloud)
print(len(g), len(ng))
