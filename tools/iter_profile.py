"""Per-iteration device time of the C2 aligner run (no profiler): icp_run with max_iterations = k for
k = 1..K, device time differences, plus the NN index / work-list diagnostics after each run.
Usage: python tools/iter_profile.py [n_points] [K]"""
import sys
sys.path.insert(0, '.')
import numpy as np
from srrg2_slam_interfaces_b200 import capi as A, synthetic as syn

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
d = syn.make_icp3d(n, n, seed=2)
ctx = A.Context(3)
ctx.set_cloud(A.FIXED, 0, d["fixed"], d["fixed_normals"])
ctx.set_cloud(A.MOVING, 0, d["moving"], d["moving_normals"])
sl = [A.make_slice(3, 0, None, A.finder_params(0.3, 0.8), A.factor_params(A.FACTOR_PLANE, A.ROB_HUBER, 0.01))]
prev = 0.0
for rep in range(2):
    ctx.icp_run(sl, A.aligner_params(max_iterations=K), np.eye(4))
for it in range(1, K + 1):
    best = 1e9
    for rep in range(3):
        r = ctx.icp_run(sl, A.aligner_params(max_iterations=it), np.eye(4))
        best = min(best, ctx.last_run_timing()[0])
    info = ctx.debug_info(0)
    print("iters=%2d total=%8.1f us  last=%7.1f us  far=%d work=%d ncorr=%d" % (
        it, best * 1e3, (best - prev) * 1e3, info["last_far_count"], info["last_work_count"],
        r['stats'][-1]['num_correspondences']), flush=True)
    prev = best
