import sys; sys.path.insert(0,'.')
import numpy as np
from srrg2_slam_interfaces_b200 import capi as A, synthetic as syn
d = syn.make_icp3d(1000000, 1000000, seed=2)
ctx = A.Context(3)
ctx.set_cloud(A.FIXED, 0, d["fixed"], d["fixed_normals"]); ctx.set_cloud(A.MOVING, 0, d["moving"], d["moving_normals"])
sl=[A.make_slice(3,0,None,A.finder_params(0.3,0.8),A.factor_params(A.FACTOR_PLANE,A.ROB_HUBER,0.01))]
for it in (3,4,5,6,8,12,20):
    r = ctx.icp_run(sl, A.aligner_params(max_iterations=it), np.eye(4))
    print(it, ctx.debug_info(0), r['stats'][-1]['num_correspondences'])
