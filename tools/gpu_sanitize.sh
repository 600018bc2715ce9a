#!/bin/bash
# compute-sanitizer passes (memcheck, racecheck, synccheck) over a small aligner run, a multi-slice 2D run, a clip and a
# pose-graph optimisation.  Usage: tools/gpu_sanitize.sh <tag>
TAG=${1:-san}; OUT=gpurun_out/$TAG; mkdir -p $OUT
cat > /tmp/san_case.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np
from srrg2_slam_interfaces_b200 import capi as A, synthetic as syn
d = syn.make_icp3d(6000, 5001, seed=2)
ctx = A.Context(3)
ctx.set_cloud(A.FIXED, 0, d["fixed"], d["fixed_normals"]); ctx.set_cloud(A.MOVING, 0, d["moving"], d["moving_normals"])
sl = [A.make_slice(3, 0, None, A.finder_params(0.3, 0.8), A.factor_params(A.FACTOR_PLANE, A.ROB_HUBER, 0.01))]
for _ in range(2):
    r = ctx.icp_run(sl, A.aligner_params(max_iterations=10), np.eye(4))
ctx.get_correspondences(0, 5001)
ctx.scene_set(0, d["moving"], d["moving_normals"]); ctx.scene_clip(0, 0, np.eye(4), 7.0); ctx.icp_run(sl, A.aligner_params(max_iterations=6), np.eye(4))
ctx.close()
m = syn.make_multicue2d(20000, n_beams=360, seed=5)
c2 = A.Context(2)
s2 = []
for k, sc in enumerate(m["scans"]):
    c2.set_cloud(A.FIXED, k, sc["points"], sc["normals"]); c2.set_cloud(A.MOVING, k, m["map"], m["map_normals"])
    s2.append(A.make_slice(2, k, sc["robot_in_sensor"], A.finder_params(0.5, 0.7), A.factor_params(A.FACTOR_PLANE, A.ROB_CAUCHY, 0.05)))
s2.append(A.make_slice(2, prior_measurement=syn.iso2(0.07, -0.04, 0.02), prior_info_diag=np.full(3, 100.0)))
c2.icp_run(s2, A.aligner_params(max_iterations=8), np.eye(3)); c2.close()
# N3: K candidates in flight on their own contexts, the fixed side lent by the source context
src = A.Context(2)
dd = [syn.make_icp2d(3000, 1200 + 100 * k, seed=9, T_star=syn.iso2(0.02 * k, -0.01 * k, 0.005 * k), paired=False) for k in range(4)]
src.set_cloud(A.FIXED, 0, dd[0]["fixed"], dd[0]["fixed_normals"]); src.set_cloud(A.MOVING, 0, dd[0]["moving"], dd[0]["moving_normals"])
sl2 = [A.make_slice(2, 0, None, A.finder_params(0.5, 0.7), A.factor_params(A.FACTOR_PLANE, A.ROB_CAUCHY, 0.05))]
src.icp_run(sl2, A.aligner_params(max_iterations=2), np.eye(3))
cx = []
for x in dd:
    q = A.Context(2); q.share_fixed(0, src, 0); q.set_cloud(A.MOVING, 0, x["moving"], x["moving_normals"]); cx.append(q)
cb = A.closure_batch(cx, sl2, A.aligner_params(max_iterations=8, enable_inlier_only_runs=True, keep_only_inlier_correspondences=True),
                     [np.eye(3, dtype=np.float32)] * 4, A.closure_params(100, 0.01, 0.5))
for q in cx: q.close()
src.close()
g = syn.make_pose_graph3d(300, 1200, seed=4, box=(6, 6, 2))
c3 = A.Context(3); c3.pgo_upload(g["guess"], g["fixed"], g["ij"], g["Z"], g["Omega"]); h = c3.pgo_optimize(max_iterations=8); c3.close()
print("sanitizer case done:", r["status"], len(h))
PY
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_case.py > $OUT/$tool.log 2>&1; echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitizer case done" $OUT/$tool.log
done
