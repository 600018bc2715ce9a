"""Wall-clock breakdown of one end-to-end step (host buffers in, pose out) of the C2 workload."""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from srrg2_slam_interfaces_b200 import capi as A, synthetic as syn
from bench import bind_to_gpu_numa
print('cores bound to the GPU-local NUMA node:', bind_to_gpu_numa(0))
n = 1000000
d = syn.make_icp3d(n, n, seed=2)
host = {k: torch.from_numpy(np.ascontiguousarray(d[k])).pin_memory().numpy() for k in ("fixed", "fixed_normals", "moving", "moving_normals")}
ctx = A.Context(3)
sl = [A.make_slice(3, 0, None, A.finder_params(0.3, 0.8), A.factor_params(A.FACTOR_PLANE, A.ROB_HUBER, 0.01))]
ap = A.aligner_params(max_iterations=20, min_num_inliers=10)
T0 = np.eye(4, dtype=np.float32)
for rep in range(6):
    t0 = time.perf_counter()
    ctx.set_cloud(A.FIXED, 0, host["fixed"], host["fixed_normals"])
    t1 = time.perf_counter()
    ctx.set_cloud(A.MOVING, 0, host["moving"], host["moving_normals"])
    t2 = time.perf_counter()
    r = ctx.icp_run(sl, ap, T0)
    t3 = time.perf_counter()
    print("step %d: set_fixed %.3f ms  set_moving %.3f ms  icp_run %.3f ms (device loop %.3f ms)  total %.3f ms" % (
        rep, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), ctx.last_run_timing()[0], 1e3 * (t3 - t0)), flush=True)
ctx.close()
