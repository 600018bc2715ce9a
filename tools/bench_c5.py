#!/usr/bin/env python
"""Config C5 (BASELINE.json): multi-cue 2D aligner -- 2 laser scans x 1080 beams (fixed side) + odometry
prior against a local map on the moving side, 10 iterations.  The 10M-point map is sharded over 8 GPUs
in the full configuration; this script times ONE rank's share (default 1.25M map points) on one GPU,
or the sharded run when launched under torch.distributed.run.  Prints one JSON line.
  python tools/bench_c5.py [map_points_per_gpu] [steps]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from srrg2_slam_interfaces_b200 import capi as A, synthetic as syn

n_map = int(sys.argv[1]) if len(sys.argv) > 1 else 1250000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
d = syn.make_multicue2d(n_map * world, n_beams=1080, seed=5)
b, e = rank * n_map, (rank + 1) * n_map
odom = syn.iso2(0.07, -0.04, np.deg2rad(1.2))
ctx = A.Context(2, local_rank)
if world > 1:
    uid = [ctx.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(uid[0], rank, world)
fp, fa = A.finder_params(0.5, 0.7), A.factor_params(A.FACTOR_PLANE, A.ROB_CAUCHY, 0.05)
sl = []
for k, sc in enumerate(d["scans"]):
    ctx.set_cloud(A.FIXED, k, sc["points"], sc["normals"])
    ctx.set_cloud(A.MOVING, k, d["map"][b:e], d["map_normals"][b:e], index_offset=b, n_global=n_map * world)
    sl.append(A.make_slice(2, k, sc["robot_in_sensor"], fp, fa))
sl.append(A.make_slice(2, prior_measurement=odom, prior_info_diag=np.full(3, 100.0)))
ap = A.aligner_params(max_iterations=10, min_num_inliers=10)
ms = []
for rep in range(3 + steps):
    r = ctx.icp_run(sl, ap, np.eye(3))
    if rep >= 3:
        ms.append(ctx.last_run_timing()[0])
t = float(np.mean(ms))
if world > 1:
    tt = torch.tensor([t], dtype=torch.float64, device="cuda")
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t = float(tt.item())
if rank == 0:
    it = len(r["stats"])
    rot, trans = syn.pose_error(r["T"], d["T_star"])
    print(json.dumps({"config": "C5: 2D multi-cue, 2 scans x 1080 beams + odom prior vs %d-point local map (%d per GPU)" % (n_map * world, n_map),
                      "n_gpus": world, "iterations": it, "ms_per_run": t, "iters_per_s": it / (t * 1e-3),
                      "map_points_per_s": n_map * world * it / (t * 1e-3), "status": r["status"],
                      "pose_error_rad_m": [rot, trans], "last_stats": r["stats"][-1]}))
ctx.close()
if world > 1:
    dist.destroy_process_group()
