#!/usr/bin/env python
"""Secondary benchmark: pose-graph GN (config C4: 100k SE(3) poses / 500k factors, synthetic
Manhattan-3D).  Prints one JSON line: device time of linearise and of the PCG solve per GN iteration,
achieved GB/s of the linearise kernel against its algorithmic bytes (312 B x F + 168 B x V),
CG iterations, chi^2 trajectory, and the oracle (numpy/scipy, host cores) on a bounded smaller graph.
Under torchrun the factor list is sharded over the ranks (all-reduce of H/b only)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--poses", type=int, default=100000)
    ap.add_argument("--factors", type=int, default=500000)
    ap.add_argument("--iters", type=int, default=6)
    ap.add_argument("--cpu-poses", type=int, default=3000)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from srrg2_slam_interfaces_b200 import capi as A
    from srrg2_slam_interfaces_b200 import synthetic as syn
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    t0 = time.perf_counter()
    g = syn.make_pose_graph3d(a.poses, a.factors, seed=4)
    t_gen = time.perf_counter() - t0
    ctx = A.Context(3, local)
    if world > 1:
        uid = [ctx.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], rank, world)
    t0 = time.perf_counter()
    ctx.pgo_upload(g["guess"], g["fixed"], g["ij"], g["Z"], g["Omega"])
    t_up = time.perf_counter() - t0
    hist = []
    for _ in range(a.iters):
        hist.append(ctx.pgo_iterate(max_cg_iterations=5000, cg_tolerance=1e-8))
    poses = ctx.pgo_download().astype(np.float64)
    err = np.linalg.norm(poses[:, :3, 3] - g["truth"][:, :3, 3], axis=1)
    err0 = np.linalg.norm(g["guess"][:, :3, 3].astype(np.float64) - g["truth"][:, :3, 3], axis=1)
    if rank == 0:
        F, V = g["ij"].shape[0], a.poses
        lin_bytes = 312 * F + 168 * V
        lin_ms = float(np.median([h["linearize_ms"] for h in hist]))
        line = {"metric": "pose-graph GN iteration (C4)", "n_gpus": world, "poses": V, "factors": F,
                "gn_iterations": len(hist), "linearize_ms": lin_ms,
                "linearize_GBps_algorithmic": lin_bytes / (lin_ms * 1e-3) / 1e9,
                "solve_ms": [round(h["solve_ms"], 3) for h in hist], "cg_iterations": [h["cg_iterations"] for h in hist],
                "cg_residual": [h["cg_relative_residual"] for h in hist], "chi": [h["chi"] for h in hist],
                "dx_norm_inf": [h["dx_norm_inf"] for h in hist], "num_blocks": hist[0]["num_blocks"],
                "position_error_mean_m": [float(err0.mean()), float(err.mean())], "upload_s": t_up, "generate_s": t_gen}
        if world == 1:
            from oracle import pgo_oracle as P
            gs = syn.make_pose_graph3d(a.cpu_poses, 5 * a.cpu_poses, seed=4, box=(14, 14, 3))
            t0 = time.perf_counter()
            _, st = P.gn_step(gs["guess"], gs["ij"], gs["Z"], gs["Omega"], gs["fixed"])
            line["cpu_baseline"] = {"kind": "port", "sample": "one GN iteration of oracle/pgo_oracle.py (numpy + scipy spsolve) "
                                    "on %d poses / %d factors" % (a.cpu_poses, gs["ij"].shape[0]),
                                    "seconds": time.perf_counter() - t0, "cores": os.cpu_count()}
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
