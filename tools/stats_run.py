"""One C2 run with max_iterations = K repeated `runs` times (tile statistics builds print at exit).
Usage: python tools/stats_run.py K [runs]"""
import sys
sys.path.insert(0, '.')
import numpy as np
from srrg2_slam_interfaces_b200 import capi as A, synthetic as syn
K = int(sys.argv[1]); runs = int(sys.argv[2]) if len(sys.argv) > 2 else 1
d = syn.make_icp3d(1000000, 1000000, seed=2)
ctx = A.Context(3)
ctx.set_cloud(A.FIXED, 0, d["fixed"], d["fixed_normals"])
ctx.set_cloud(A.MOVING, 0, d["moving"], d["moving_normals"])
sl = [A.make_slice(3, 0, None, A.finder_params(0.3, 0.8), A.factor_params(A.FACTOR_PLANE, A.ROB_HUBER, 0.01))]
for _ in range(runs):
    ctx.icp_run(sl, A.aligner_params(max_iterations=K), np.eye(4))
print("K", K, "device ms", ctx.last_run_timing()[0])
ctx.close()
