#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: ICP iterations/s on the 1M-vs-1M point SE(3) align (config C2:
point+normal factor, Huber robustifier, 20 iterations) + achieved HBM GB/s of the fused kernel.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl graft|reference]

A "step" is one MultiAligner compute(): 20 _runSolver iterations over the resident clouds.
  value  device-timed (CUDA events on the context stream), clouds resident in HBM
  e2e    same step through the C ABI with HOST (pinned) buffers: H2D of both clouds, index build,
         20 iterations, D2H of pose + IterationStats inside the timed region
N > 1 (torchrun): weak scaling -- the fixed cloud (1M) is replicated, every rank owns a 1M-point
shard of an N x 1M moving cloud, one 256-entry int64 NCCL all-reduce per iteration; value is in
1M-point-equivalent iterations/s (= N x global iterations/s), max-over-ranks time.
--impl reference times the CPU path (the oracle restatement, all host threads) on the same config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_POINTS = 1_000_000
ICP_ITERS = 20
MAX_DISTANCE, NORMAL_COS, HUBER_TAU = 0.3, 0.8, 0.01
BYTES_PER_POINT = 56  # 24 B moving point+normal, 24 B gathered fixed, 8 B idx+response
METRIC = "ICP iters/sec on 1M-pt SE(3) align"
UNIT = "iters/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft")
    ap.add_argument("--points", type=int, default=N_POINTS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_config(n, world):
    return {
        "workload": "C2: SE(3) point+normal ICP, %d vs %d synthetic pts per GPU, %d iters, Huber" % (n, n, ICP_ITERS),
        "n_fixed": n, "n_moving_per_gpu": n, "icp_iterations": ICP_ITERS, "factor": "point+normal (4 rows)",
        "robustifier": "Huber tau=%g" % HUBER_TAU, "max_distance": MAX_DISTANCE, "normal_cos": NORMAL_COS,
        "sharding": "moving cloud sharded by rank, fixed cloud replicated" if world > 1 else "none",
        "l2": "flushed between steps (512 MB memset); the 20 iterations inside a step re-read the same clouds",
    }


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu = gpu
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa(gpu):
    """Pins this process to the CPU cores NVML reports as local to the GPU, BEFORE the pinned host
    buffers are allocated: pinned pages land where the allocating thread runs, and a buffer on a remote
    socket uploads slower.  (A no-op on single-node hosts such as this pool's 16-core boxes.)"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("icp_slice_kernel_dram_bytes_per_launch")
        except Exception:
            return None
    return None


# ------------------------------------------------------------------------------------------------
def cpu_iterations(O, d, threads, iters):
    """Times `iters` _runSolver iterations of the CPU path (oracle) with a prebuilt kd-tree."""
    O.set_threads(threads)
    F = O.CloudRef(d["fixed"], d["fixed_normals"])
    M = O.CloudRef(d["moving"], d["moving_normals"])
    fp = O.finder_params(MAX_DISTANCE, NORMAL_COS)
    fa = O.factor_params(O.FACTOR_PLANE, O.ROB_HUBER, HUBER_TAU)
    t0 = time.perf_counter()
    ix = O.Index(F, O.NN_KDTREE)
    t_build = time.perf_counter() - t0
    T = np.eye(4, dtype=np.float32)
    t0 = time.perf_counter()
    for _ in range(iters):
        fidx, _ = O.find(ix, F, M, T, fp)
        lin = O.linearize(F, M, fidx, T, fp, fa, want_status=False)
        ok, T = O.solve_update(3, O.VAR_SE3_QUAT_RIGHT, lin["H"], lin["b"], T)
    dt = time.perf_counter() - t0
    return dt, t_build


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    from srrg2_slam_interfaces_b200 import synthetic as syn
    n = args.points
    threads = os.cpu_count() or 1
    d = syn.make_icp3d(n, n, seed=2)
    iters_per_step = ICP_ITERS  # one step = the whole 20-iteration _runSolver, like the graft arm's step
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_iterations(O, d, threads, 1)
    total, t_build = 0.0, 0.0
    for _ in range(args.steps):
        dt, tb = cpu_iterations(O, d, threads, iters_per_step)
        total += dt
        t_build = tb
    value = iters_per_step * args.steps / total
    sample = ("%d _runSolver iterations per step at the full %d x %d size, kd-tree prebuilt "
              "(build %.2f s not counted), oracle/srrg2b_oracle.c with OpenMP" % (iters_per_step, n, n, t_build))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(n, 1),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def run_graft(args):
    import torch
    import torch.distributed as dist
    from srrg2_slam_interfaces_b200 import capi as A
    from srrg2_slam_interfaces_b200 import synthetic as syn

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    all_cpus = os.sched_getaffinity(0)
    bind_to_gpu_numa(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n = args.points
    d = syn.make_icp3d(n, n, seed=2, moving_stream=rank)

    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()

    keep, host = [], {}
    for k in ("fixed", "fixed_normals", "moving", "moving_normals"):
        t, v = pinned(d[k])
        keep.append(t)
        host[k] = v

    ctx = A.Context(3, local_rank)
    if world > 1:
        uid = [ctx.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], rank, world)
    sl = [A.make_slice(3, 0, None, A.finder_params(MAX_DISTANCE, NORMAL_COS),
                       A.factor_params(A.FACTOR_PLANE, A.ROB_HUBER, HUBER_TAU))]
    ap = A.aligner_params(max_iterations=ICP_ITERS, min_num_inliers=10)
    T0 = np.eye(4, dtype=np.float32)

    def upload():
        ctx.set_cloud(A.FIXED, 0, host["fixed"], host["fixed_normals"])
        ctx.set_cloud(A.MOVING, 0, host["moving"], host["moving_normals"], index_offset=rank * n, n_global=world * n)

    flush_buf = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def flush_l2():
        flush_buf.zero_()
        torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    upload()
    res = None
    for _ in range(max(args.warmup, 3)):
        res = ctx.icp_run(sl, ap, T0)
    assert res["status"] == A.ALIGNER_SUCCESS, res["status"]

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # ---- value: device time of K steps, clouds resident ----
    launches0 = ctx.launch_count
    dev_ms = 0.0
    iters_done = 0
    barrier()
    for _ in range(args.steps):
        flush_l2()
        barrier()
        res = ctx.icp_run(sl, ap, T0)
        ms, it = ctx.last_run_timing()
        dev_ms += ms
        iters_done += it
    barrier()
    gpu_launches = ctx.launch_count - launches0
    # ---- e2e: host buffers in, pose + stats out, every step ----
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        upload()
        res_e = ctx.icp_run(sl, ap, T0)
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    # ---- roofline: one ICP iteration = one pass of the slice kernels over the resident clouds; its
    # duration is the CUDA-event time of the timed steps above divided by the iterations they ran ----
    kms, kn = dev_ms, iters_done
    t = torch.tensor([dev_ms, e2e_s, kms / max(kn, 1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_s_max, k_ms = [float(x) for x in t.tolist()]

    if rank == 0:
        value = world * iters_done / (dev_ms_max * 1e-3)
        e2e_value = world * ICP_ITERS * args.steps / e2e_s_max
        peak, peak_src = measured_peak()
        achieved = BYTES_PER_POINT * n / (k_ms * 1e-3) / 1e9
        h2d = sum(host[k].nbytes for k in host)
        d2h = 64 + 56 * ICP_ITERS
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(n, world),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": ncu_traffic(), "kernel": "one ICP iteration: linearize_kernel<3,PLANE,CHECK> (+ nn_kernel / nn_far_kernel / linearize_kernel on iterations that search) + icp_solve_kernel; CUDA events around the graph-replayed run / iterations",
                             "kernel_ms": k_ms, "algorithmic_bytes_per_launch": BYTES_PER_POINT * n,
                             "peak_source": peak_src},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": 1e3 * e2e_s_max / args.steps},
                "gpu_launches": int(gpu_launches), "clocks": clocks,
                "result_check": {"status": res["status"], "iterations": len(res["stats"]),
                                 "pose_error_rad_m": list(syn.pose_error(res["T"], d["T_star"])),
                                 "last_num_inliers": res["stats"][-1]["num_inliers"],
                                 "e2e_pose_equal": bool(np.array_equal(res["T"], res_e["T"])),
                                 "nn_index": ctx.debug_info(0)}}
        if world == 1 and not args.no_cpu_baseline:
            from oracle import oracle as O
            os.sched_setaffinity(0, all_cpus)  # the CPU baseline may use every host core
            threads = os.cpu_count() or 1
            dt, tb = cpu_iterations(O, d, threads, ICP_ITERS)
            line["cpu_baseline"] = {"value": ICP_ITERS / dt, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "one whole step: %d _runSolver iterations from the identity guess at the "
                                              "full %d x %d size, kd-tree prebuilt (build %.2f s not counted), "
                                              "oracle/srrg2b_oracle.c with OpenMP on all host cores"
                                              % (ICP_ITERS, n, n, tb)}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_graft(a)
