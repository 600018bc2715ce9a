"""Two C2 aligner runs (for ncu captures of single kernels).  Usage: python tools/one_run.py [n] [iters] [runs]"""
import sys
sys.path.insert(0, '.')
import numpy as np
from srrg2_slam_interfaces_b200 import capi as A, synthetic as syn

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
runs = int(sys.argv[3]) if len(sys.argv) > 3 else 2
d = syn.make_icp3d(n, n, seed=2)
ctx = A.Context(3)
ctx.set_cloud(A.FIXED, 0, d["fixed"], d["fixed_normals"])
ctx.set_cloud(A.MOVING, 0, d["moving"], d["moving_normals"])
sl = [A.make_slice(3, 0, None, A.finder_params(0.3, 0.8), A.factor_params(A.FACTOR_PLANE, A.ROB_HUBER, 0.01))]
for rep in range(runs):
    r = ctx.icp_run(sl, A.aligner_params(max_iterations=K), np.eye(4))
    print("run", rep, "device ms", ctx.last_run_timing()[0], "iters", len(r["stats"]), flush=True)
ctx.close()
