#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: ICP iterations/s on the 1M-vs-1M point SE(3) align + achieved HBM GB/s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl graft|reference] [--config c2|c5|c3|c4]

A "step" is one MultiAligner compute() (R/registration/aligners/multi_aligner_impl.cpp:46-95): all _runSolver
iterations of the configuration over the resident clouds.
  value  device-timed (CUDA events on the context stream around the graph-replayed run), clouds resident in
         HBM, COLD: the slice's correspondences / warm-start candidates / certified bounds are forgotten before
         every timed step, as a tracker's fresh setMoving() would (the warm figure is reported beside it)
  roofline                       algorithmic bytes of one iteration / (device time of the run / iterations): the whole
                                 run, searching iterations included
  roofline_converged_iteration   the same for the converged iterations alone, measured in the same run: the cold step
                                 repeated with half of the iterations, the difference is the second half (their kernel,
                                 check_tiles_kernel, is the dominant one of the run)
  e2e    the same step through the C ABI with HOST (pinned) buffers: H2D of the clouds, index build, all
         iterations, D2H of pose + IterationStats inside the timed region.  N > 1: the replicated fixed cloud
         crosses PCIe once IN TOTAL (every rank uploads 1/N of it) and is all-gathered over NVLink; every rank
         uploads its own shard of the moving cloud
Configurations (BASELINE.json `configs`):
  c2 (default, the metric's configuration)  SE(3) point+normal, 1M vs 1M, 20 iterations, Huber.  N > 1: weak
     scaling -- the fixed cloud is replicated, every rank owns a 1M-point shard of an N x 1M moving cloud; value
     is in 1M-point-equivalent iterations/s (= N x global iterations/s)
  c5  2D multi-cue: 2 scans x 1080 beams (fixed) + odometry prior against a 10M-point local map (moving), 10
     iterations; STRONG scaling: the map is sharded over the N ranks (R/trackers/multi_tracker_impl.cpp:97-98)
  c3  RGB-D projective association, 640x480, 30-frame sequence (29 aligner calls x 10 iterations per step); N = 1
  c4  pose-graph optimisation, 100k SE(3) poses / 500k factors: damped Gauss-Newton to |dx| < 1e-6, a step = one
     iteration (linearise + PCG solve + update); factor-sharded under torchrun (all-reduce of H / b / chi only)
The accumulators of the ranks are exchanged inside the solve kernel over peer mailboxes (NVLink loads / stores);
NCCL only bootstraps (IPC handles, coordinate bounds).  With N > 1 rank 0 runs the CPU oracle on the GLOBAL
clouds and compares pose bits, IterationStats and the concatenated correspondence lists (`result_check.parity`).
--impl reference times the CPU path (the oracle restatement, all host threads) on the same configuration.
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

METRIC = "ICP iters/sec on 1M-pt SE(3) align"
UNIT = "iters/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft")
    ap.add_argument("--config", default="c2", choices=["c2", "c5", "c3", "c4"])
    ap.add_argument("--points", type=int, default=0, help="c2: points per cloud and GPU; c5: map points in total")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the in-run oracle comparison")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# configurations of the aligner path
# ------------------------------------------------------------------------------------------------
class C2:
    """SE(3) point+normal ICP, 1M vs 1M, 20 iterations, Huber (weak scaling)."""
    name, dim, iters, scaling = "c2", 3, 20, "weak"
    max_distance, normal_cos, tau = 0.3, 0.8, 0.01
    bytes_per_point = 56  # 24 B moving point+normal, 24 B gathered fixed, 8 B idx+response

    def __init__(self, args, rank, world):
        self.n = args.points or 1_000_000
        self.rank, self.world = rank, world

    def workload(self):
        n = self.n
        return {"workload": "C2: SE(3) point+normal ICP, %d vs %d synthetic pts per GPU, %d iters, Huber" % (n, n, self.iters),
                "n_fixed": n, "n_moving_per_gpu": n, "icp_iterations": self.iters, "factor": "point+normal (4 rows)",
                "robustifier": "Huber tau=%g" % self.tau, "max_distance": self.max_distance, "normal_cos": self.normal_cos,
                "sharding": "moving cloud sharded by rank, fixed cloud replicated" if self.world > 1 else "none"}

    def data(self, syn, rank=None):
        return syn.make_icp3d(self.n, self.n, seed=2, moving_stream=self.rank if rank is None else rank)

    def host_arrays(self, d):
        return {"fixed": d["fixed"], "fixed_normals": d["fixed_normals"], "moving": d["moving"],
                "moving_normals": d["moving_normals"]}

    def replicated(self):  # arrays every rank holds identically (fanned out over NVLink in the e2e step)
        return ["fixed", "fixed_normals"]

    def upload(self, A, ctx, host):
        ctx.set_cloud(A.FIXED, 0, host["fixed"], host["fixed_normals"])
        self.upload_sharded(A, ctx, host)

    def upload_sharded(self, A, ctx, host):  # this rank's part of the sharded operand
        n, r, w = self.n, self.rank, self.world
        ctx.set_cloud(A.MOVING, 0, host["moving"], host["moving_normals"], index_offset=r * n, n_global=w * n)

    def upload_replicated_from_device(self, A, ctx, dev):
        ctx.set_cloud_device(A.FIXED, 0, dev["fixed"].data_ptr(), dev["fixed_normals"].data_ptr(), None, self.n)

    # resident-map variant of the e2e step: the moving side (the tracker's local map) stays in HBM
    def scene_arrays(self):
        return "moving", "moving_normals"

    def upload_fixed(self, A, ctx, host):
        ctx.set_cloud(A.FIXED, 0, host["fixed"], host["fixed_normals"])

    def fixed_bytes(self, host):
        return host["fixed"].nbytes + host["fixed_normals"].nbytes

    def slices(self, A):
        return [A.make_slice(3, 0, None, A.finder_params(self.max_distance, self.normal_cos),
                             A.factor_params(A.FACTOR_PLANE, A.ROB_HUBER, self.tau))]

    def point_slices(self):
        return [0]

    def aligner_params(self, A):
        return A.aligner_params(max_iterations=self.iters, min_num_inliers=10)

    def T0(self):
        return np.eye(4, dtype=np.float32)

    def local_moving(self):
        return self.n

    def algorithmic_bytes_per_iteration(self):  # per GPU
        return self.bytes_per_point * self.n

    def units_per_step(self):  # what `value` counts per step and GPU-set
        return self.world * self.iters

    def oracle_run(self, O, syn, iters=None, world=None):
        """The CPU path on the GLOBAL clouds (all shards of the moving cloud)."""
        world = self.world if world is None else world
        d0 = self.data(syn, 0)
        mov = [d0["moving"]] + [self.data(syn, r)["moving"] for r in range(1, world)]
        mnr = [d0["moving_normals"]] + [self.data(syn, r)["moving_normals"] for r in range(1, world)]
        F = O.CloudRef(d0["fixed"], d0["fixed_normals"])
        M = O.CloudRef(np.concatenate(mov), np.concatenate(mnr))
        sl = [O.make_slice(F, M, None, O.finder_params(self.max_distance, self.normal_cos),
                           O.factor_params(O.FACTOR_PLANE, O.ROB_HUBER, self.tau))]
        return O.icp_run(3, sl, O.aligner_params(max_iterations=iters or self.iters, min_num_inliers=10), np.eye(4))


class C5:
    """2D multi-cue aligner: 2 scans x 1080 beams + odometry prior vs a 10M-point local map (strong scaling)."""
    name, dim, iters, scaling = "c5", 2, 10, "strong"
    max_distance, normal_cos, tau = 0.5, 0.7, 0.05
    bytes_per_point = 40  # SE(2) point+normal

    def __init__(self, args, rank, world):
        self.n_map = args.points or 10_000_000
        self.rank, self.world = rank, world
        self.b = (self.n_map * rank) // world
        self.e = (self.n_map * (rank + 1)) // world

    def workload(self):
        return {"workload": "C5: 2D multi-cue, 2 scans x 1080 beams + odometry prior vs a %d-point local map, %d iters, "
                            "Cauchy" % (self.n_map, self.iters),
                "n_fixed": [1080, 1080], "n_moving_total": self.n_map, "n_moving_per_gpu": self.e - self.b,
                "icp_iterations": self.iters, "factor": "2D point+normal (2 rows) x 2 slices + SE(2) prior",
                "robustifier": "Cauchy tau=%g" % self.tau, "max_distance": self.max_distance, "normal_cos": self.normal_cos,
                "sharding": "local map (the aligner's moving side) sharded by rank, scans replicated" if self.world > 1 else "none"}

    def data(self, syn, rank=None):
        if not hasattr(self, "_d"):
            self._d = syn.make_multicue2d(self.n_map, n_beams=1080, seed=5)
        return self._d

    def host_arrays(self, d):
        h = {"map": d["map"][self.b:self.e], "map_normals": d["map_normals"][self.b:self.e]}
        for k, sc in enumerate(d["scans"]):
            h["scan%d" % k] = sc["points"]
            h["scan%d_normals" % k] = sc["normals"]
        self._ris = [sc["robot_in_sensor"] for sc in d["scans"]]
        return h

    def replicated(self):
        return []  # (the scans are 2 x 1080 points: not worth a broadcast)

    def upload(self, A, ctx, host):
        for k in range(2):
            ctx.set_cloud(A.FIXED, k, host["scan%d" % k], host["scan%d_normals" % k])
            ctx.set_cloud(A.MOVING, k, host["map"], host["map_normals"], index_offset=self.b, n_global=self.n_map)

    def scene_arrays(self):
        return "map", "map_normals"

    def upload_fixed(self, A, ctx, host):
        for k in range(2):
            ctx.set_cloud(A.FIXED, k, host["scan%d" % k], host["scan%d_normals" % k])

    def fixed_bytes(self, host):
        return sum(host["scan%d%s" % (k, sfx)].nbytes for k in range(2) for sfx in ("", "_normals"))

    def slices(self, A):
        from srrg2_slam_interfaces_b200 import synthetic as syn
        fp, fa = A.finder_params(self.max_distance, self.normal_cos), A.factor_params(A.FACTOR_PLANE, A.ROB_CAUCHY, self.tau)
        sl = [A.make_slice(2, k, self._ris[k], fp, fa) for k in range(2)]
        sl.append(A.make_slice(2, prior_measurement=syn.iso2(0.07, -0.04, np.deg2rad(1.2)), prior_info_diag=np.full(3, 100.0)))
        return sl

    def point_slices(self):
        return [0, 1]

    def aligner_params(self, A):
        return A.aligner_params(max_iterations=self.iters, min_num_inliers=10)

    def T0(self):
        return np.eye(3, dtype=np.float32)

    def local_moving(self):
        return self.e - self.b

    def algorithmic_bytes_per_iteration(self):
        return 2 * self.bytes_per_point * (self.e - self.b)

    def units_per_step(self):
        return self.iters

    def oracle_run(self, O, syn, iters=None, world=None):
        d = self.data(syn)
        M = O.CloudRef(d["map"], d["map_normals"])
        fp, fa = O.finder_params(self.max_distance, self.normal_cos), O.factor_params(O.FACTOR_PLANE, O.ROB_CAUCHY, self.tau)
        sl = [O.make_slice(O.CloudRef(sc["points"], sc["normals"]), M, sc["robot_in_sensor"], fp, fa, dim=2) for sc in d["scans"]]
        sl.append(O.make_slice(prior_measurement=syn.iso2(0.07, -0.04, np.deg2rad(1.2)), prior_info_diag=np.full(3, 100.0), dim=2))
        return O.icp_run(2, sl, O.aligner_params(max_iterations=iters or self.iters, min_num_inliers=10), np.eye(3))


CONFIGS = {"c2": C2, "c5": C5}


# ------------------------------------------------------------------------------------------------
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


def ncu_traffic(config):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (the
    profile names the commit it was taken at); null for configurations without a capture."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p)).get(config, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
# the CPU arm (reference = the oracle restatement; the upstream sources cannot be built here)
# ------------------------------------------------------------------------------------------------
def cpu_c2_iterations(O, cfg, d, threads, iters):
    """Times `iters` _runSolver iterations of the CPU path (oracle) with a prebuilt kd-tree."""
    O.set_threads(threads)
    F = O.CloudRef(d["fixed"], d["fixed_normals"])
    M = O.CloudRef(d["moving"], d["moving_normals"])
    fp = O.finder_params(cfg.max_distance, cfg.normal_cos)
    fa = O.factor_params(O.FACTOR_PLANE, O.ROB_HUBER, cfg.tau)
    t0 = time.perf_counter()
    ix = O.Index(F, O.NN_KDTREE)
    t_build = time.perf_counter() - t0
    T = np.eye(4, dtype=np.float32)
    t0 = time.perf_counter()
    for _ in range(iters):
        fidx, _ = O.find(ix, F, M, T, fp)
        lin = O.linearize(F, M, fidx, T, fp, fa, want_status=False)
        ok, T = O.solve_update(3, O.VAR_SE3_QUAT_RIGHT, lin["H"], lin["b"], T)
    return time.perf_counter() - t0, t_build


def cpu_step(O, syn, cfg, threads, sample_iters=None):
    """One bounded CPU sample of the configuration's step: (seconds, iterations run, description)."""
    if cfg.name == "c2":
        d = cfg.data(syn, 0)
        it = sample_iters or cfg.iters
        dt, tb = cpu_c2_iterations(O, cfg, d, threads, it)
        return dt, it, ("%d _runSolver iterations from the identity guess at the full %d x %d size, kd-tree prebuilt "
                        "(build %.2f s not counted), oracle/srrg2b_oracle.c with OpenMP on %d threads" % (it, cfg.n, cfg.n, tb, threads))
    O.set_threads(threads)
    it = sample_iters or 2
    t0 = time.perf_counter()
    cfg.oracle_run(O, syn, iters=it, world=1)
    dt = time.perf_counter() - t0
    return dt, it, ("%d _runSolver iterations of the full multi-cue problem (2 x 1080-point scans, %d-point map, prior), "
                    "kd-trees included, oracle/srrg2b_oracle.c with OpenMP on %d threads" % (it, cfg.n_map, threads))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.config in ("c3", "c4"):
        print(json.dumps({"impl": "reference", "unavailable": "the reference arm is implemented for the aligner configurations c2 / c5"}))
        return
    from oracle import oracle as O
    from srrg2_slam_interfaces_b200 import synthetic as syn
    threads = os.cpu_count() or 1
    cfg = CONFIGS[args.config](args, 0, 1)
    if args.warmup > 0:
        cpu_step(O, syn, cfg, threads, 1)
    total, its, sample = 0.0, 0, ""
    for _ in range(args.steps):
        dt, it, sample = cpu_step(O, syn, cfg, threads)
        total += dt
        its += it
    value = its / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": cfg.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg.workload(),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": "per step: " + sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def run_aligner(args):
    import torch
    import torch.distributed as dist
    from srrg2_slam_interfaces_b200 import capi as A
    from srrg2_slam_interfaces_b200 import synthetic as syn

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch with torch.distributed.run --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    all_cpus = os.sched_getaffinity(0)
    bind_to_gpu_numa(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    cfg = CONFIGS[args.config](args, rank, world)
    d = cfg.data(syn)
    keep, host = [], {}
    for k, v in cfg.host_arrays(d).items():
        t = torch.from_numpy(np.ascontiguousarray(v)).pin_memory()
        keep.append(t)
        host[k] = t.numpy()
    # device staging of the replicated arrays for the NVLink fan-out of the e2e step
    dev = None
    if world > 1 and cfg.replicated() and all(host[k].shape[0] % world == 0 for k in cfg.replicated()):
        dev = {k: torch.empty(host[k].shape, dtype=torch.float32, device="cuda") for k in cfg.replicated()}
        shard = {k: torch.empty((host[k].shape[0] // world,) + host[k].shape[1:], dtype=torch.float32, device="cuda") for k in dev}

    ctx = A.Context(cfg.dim, local_rank)
    if world > 1:
        uid = [ctx.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], rank, world)
    sl, ap, T0 = cfg.slices(A), cfg.aligner_params(A), cfg.T0()

    def upload_e2e():
        if not dev:
            cfg.upload(A, ctx, host)
            return
        # the sharded operand first (its H2D copy runs on the library's copy stream), then the replicated cloud:
        # every rank uploads 1/N of it over its own PCIe link and the parts are all-gathered over NVLink
        cfg.upload_sharded(A, ctx, host)
        for k in dev:
            rows = host[k].shape[0] // world
            shard[k].copy_(torch.from_numpy(host[k][rank * rows:(rank + 1) * rows]), non_blocking=True)
            dist.all_gather_into_tensor(dev[k].view(-1), shard[k].view(-1))
        torch.cuda.current_stream().synchronize()
        cfg.upload_replicated_from_device(A, ctx, dev)

    flush_buf = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def flush_l2():
        flush_buf.zero_()
        torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cfg.upload(A, ctx, host)
    res = None
    for _ in range(max(args.warmup, 3)):
        res = ctx.icp_run(sl, ap, T0)
    assert res["status"] == A.ALIGNER_SUCCESS, res["status"]

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # ---- warm figure: the slots still hold the previous step's neighbours and bounds ----
    warm_ms = 0.0
    barrier()
    for _ in range(args.steps):
        flush_l2()
        barrier()
        ctx.icp_run(sl, ap, T0)
        warm_ms += ctx.last_run_timing()[0]
    # ---- value: device time of K cold steps, clouds resident ----
    launches0 = ctx.launch_count
    dev_ms, iters_done = 0.0, 0
    barrier()
    for _ in range(args.steps):
        for s in cfg.point_slices():
            ctx.reset_correspondences(s)
        flush_l2()
        barrier()
        res = ctx.icp_run(sl, ap, T0)
        ms, it = ctx.last_run_timing()
        dev_ms += ms
        iters_done += it
    barrier()
    gpu_launches = ctx.launch_count - launches0
    corr = [ctx.get_correspondences(s, cfg.local_moving()) for s in cfg.point_slices()]
    # ---- the converged iterations alone (their kernel, check_tiles_kernel, is the dominant one of the run): device time
    # of the same cold steps cut off after half of the iterations; the difference is the second half ----
    half_ms, half_iters = 0.0, 0
    ap_half = cfg.aligner_params(A)
    ap_half.max_iterations = max(cfg.iters // 2, 1)
    ctx.icp_run(sl, ap_half, T0)
    for _ in range(args.steps):
        for s in cfg.point_slices():
            ctx.reset_correspondences(s)
        flush_l2()
        barrier()
        ctx.icp_run(sl, ap_half, T0)
        ms, it = ctx.last_run_timing()
        half_ms += ms
        half_iters += it
    barrier()
    # ---- e2e: host buffers in, pose + stats out, every step ----
    upload_e2e()
    ctx.icp_run(sl, ap, T0)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        upload_e2e()
        res_e = ctx.icp_run(sl, ap, T0)
    barrier()
    e2e_s = time.perf_counter() - t0
    # ---- e2e with the local map RESIDENT: only the new scan(s) cross PCIe; the map is clipped on the
    # device into each slice's moving cloud (identity pose, unbounded range here: the workload stays the same) ----
    ck, cn = cfg.scene_arrays()
    ctx.scene_set(0, host[ck], host[cn])
    eye = np.eye(cfg.dim + 1, dtype=np.float32)

    def step_resident():
        cfg.upload_fixed(A, ctx, host)
        for s_ in cfg.point_slices():
            ctx.scene_clip(0, s_, eye, 1e18)
        return ctx.icp_run(sl, ap, T0)

    res_r = step_resident()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res_r = step_resident()
    barrier()
    e2e_res_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    # what the PCIe links give when every rank uploads at the same time (pinned -> device, this rank's sharded operand)
    probe = [torch.empty(host[k].shape, dtype=torch.float32, device="cuda") for k in host if not (dev and k in dev)]
    srcs = [torch.from_numpy(host[k]) for k in host if not (dev and k in dev)]
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        for dst, src in zip(probe, srcs):
            dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    h2d_gbs = 3 * sum(x.numel() * 4 for x in srcs) / (time.perf_counter() - t0) / 1e9
    del probe

    t = torch.tensor([dev_ms, e2e_s, warm_ms, -h2d_gbs, e2e_res_s, half_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_s_max, warm_ms_max, h2d_gbs_min, e2e_res_s_max, half_ms_max = [float(x) for x in t.tolist()]
    h2d_gbs_min = -h2d_gbs_min

    # ---- N > 1: every rank's result must be the single-process oracle's on the global clouds, bit for bit ----
    parity = None
    if world > 1 and not args.no_parity:
        mine = {"T": np.asarray(res["T"]), "status": res["status"], "stats": res["stats"], "corr": corr}
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(mine, gathered, dst=0)
        if rank == 0:
            from oracle import oracle as O
            os.sched_setaffinity(0, all_cpus)
            O.set_threads(os.cpu_count() or 1)
            t0 = time.perf_counter()
            o = cfg.oracle_run(O, syn)
            ok_pose = all(np.array_equal(g["T"], o["T"]) and g["status"] == o["status"] for g in gathered)
            ok_stats = all(g["stats"] == o["stats"] for g in gathered)
            ok_corr = True
            for k in range(len(cfg.point_slices())):
                for j in range(3):
                    cat = np.concatenate([g["corr"][k][j] for g in gathered])
                    ok_corr = ok_corr and np.array_equal(cat, o["correspondences"][k][j])
            parity = {"verdict": "bit-exact" if (ok_pose and ok_stats and ok_corr) else "MISMATCH",
                      "pose_and_status_equal_on_all_ranks": bool(ok_pose), "iteration_stats_equal": bool(ok_stats),
                      "concatenated_correspondences_equal": bool(ok_corr),
                      "against": "oracle/srrg2b_oracle.c on the global clouds, single process", "oracle_seconds": time.perf_counter() - t0}

    if rank == 0:
        units = cfg.units_per_step()
        value = units * args.steps / (dev_ms_max * 1e-3)
        e2e_value = units * args.steps / e2e_s_max
        peak, peak_src = measured_peak()
        k_ms = dev_ms_max / max(iters_done, 1)
        abytes = cfg.algorithmic_bytes_per_iteration()
        achieved = abytes / (k_ms * 1e-3) / 1e9
        converged = None
        if iters_done > half_iters > 0 and dev_ms_max > half_ms_max:
            c_ms = (dev_ms_max - half_ms_max) / (iters_done - half_iters)
            converged = {"bound": "hbm", "achieved": abytes / (c_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": abytes / (c_ms * 1e-3) / 1e9 / peak, "kernel_ms": c_ms, "algorithmic_bytes_per_launch": abytes,
                         "kernel": "iterations %d..%d of the same cold run (full run minus a run cut off after %d iterations, CUDA events): "
                                   "check_tiles_kernel + the lone search kernel + icp_solve_kernel per point slice"
                                   % (half_iters // args.steps + 1, iters_done // args.steps, half_iters // args.steps)}
        h2d = sum(host[k].nbytes for k in host if not (dev and k in dev)) + (sum(host[k].nbytes for k in dev) // world if dev else 0)
        if cfg.name == "c5":
            h2d += host["map"].nbytes + host["map_normals"].nbytes  # the map shard is uploaded once per scanner slice
        d2h = 64 + 64 * cfg.iters
        wl = cfg.workload()
        wl["l2"] = "flushed between steps (512 MB memset); the iterations inside a step re-read the same clouds"
        wl["start"] = "cold: correspondences, warm-start candidates and certified bounds reset before every timed step"
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
                "scaling": cfg.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": wl,
                "value_warm_start": units * args.steps / (warm_ms_max * 1e-3),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": ncu_traffic(cfg.name),
                             "kernel": "one ICP iteration per point slice: check_tiles_kernel (coherence check + linearisation; nn_kernel / "
                                       "nn_far_kernel / lin_after_search_kernel on iterations that search) + icp_solve_kernel; "
                                       "CUDA events around the graph-replayed run / iterations",
                             "kernel_ms": k_ms, "algorithmic_bytes_per_launch": abytes, "peak_source": peak_src},
                "roofline_converged_iteration": converged,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": 1e3 * e2e_s_max / args.steps,
                        "h2d_gbs_per_rank_all_ranks_copying": h2d_gbs_min,  # slowest rank; the limiting copy of the e2e step
                        "fixed_cloud_fan_out": "every rank uploads 1/N of the replicated fixed cloud, NCCL all-gather over NVLink" if dev else "none"},
                "e2e_resident_map": {"value": units * args.steps / e2e_res_s_max, "unit": UNIT, "ms_per_step": 1e3 * e2e_res_s_max / args.steps,
                                     "h2d_bytes_per_step": cfg.fixed_bytes(host), "pose_equal": bool(np.array_equal(res["T"], res_r["T"])),
                                     "what": "the moving side (the tracker's local map) stays in HBM and is clipped on the device "
                                             "(srrg2b_scene_clip); only the fixed scan(s) are uploaded per step"},
                "gpu_launches": int(gpu_launches), "clocks": clocks,
                "result_check": {"status": res["status"], "iterations": len(res["stats"]),
                                 "pose_error_rad_m": list(syn.pose_error(res["T"], d["T_star"])),
                                 "last_num_inliers": res["stats"][-1]["num_inliers"],
                                 "num_saturated": int(sum(s.get("num_saturated", 0) for s in res["stats"])),
                                 "e2e_pose_equal": bool(np.array_equal(res["T"], res_e["T"])),
                                 "parity": parity, "nn_index": ctx.debug_info(cfg.point_slices()[0])}}
        if world == 1 and not args.no_cpu_baseline:
            from oracle import oracle as O
            os.sched_setaffinity(0, all_cpus)  # the CPU baseline may use every host core
            threads = os.cpu_count() or 1
            dt, it, sample = cpu_step(O, syn, cfg, threads)
            line["cpu_baseline"] = {"value": it / dt, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}
            dt1, it1, _ = cpu_step(O, syn, cfg, 1, 2)  # (the reference itself is single-threaded: reported beside it)
            line["cpu_baseline"]["single_thread_value"] = it1 / dt1
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
def run_c3(args):
    """RGB-D projective aligner over the 30-frame synthetic sequence: a step = 29 aligner calls (frame k-1 onto
    frame k) x 10 iterations, clouds resident; e2e uploads both frames of every call."""
    import torch
    from srrg2_slam_interfaces_b200 import capi as A
    from srrg2_slam_interfaces_b200 import synthetic as syn
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        raise SystemExit("config c3 runs on one GPU (frames are processed in sequence)")
    frames, K = syn.make_rgbd_sequence(30, 640, 480, seed=3)
    iters = 10
    fp = A.finder_params(0.1, 0.8, kind=A.FINDER_PROJECTIVE, fx=K["fx"], fy=K["fy"], cx=K["cx"], cy=K["cy"],
                         width=K["width"], height=K["height"], min_depth=0.1, max_depth=20.0)
    fa = A.factor_params(A.FACTOR_PLANE, A.ROB_HUBER, 0.01)
    ap = A.aligner_params(max_iterations=iters, min_num_inliers=10)
    ctx = A.Context(3, 0)
    sl = [A.make_slice(3, 0, None, fp, fa)]
    pin = [{k: torch.from_numpy(np.ascontiguousarray(f[k])).pin_memory().numpy() for k in ("points", "normals", "valid")} for f in frames]

    def pair(k):
        ctx.set_cloud(A.FIXED, 0, pin[k]["points"], pin[k]["normals"], pin[k]["valid"])
        ctx.set_cloud(A.MOVING, 0, pin[k - 1]["points"], pin[k - 1]["normals"], pin[k - 1]["valid"])

    sampler = ClockSampler(0)
    sampler.start()
    n_pairs = len(frames) - 1
    for k in range(1, 4):
        pair(k)
        ctx.icp_run(sl, ap, np.eye(4))
    launches0 = ctx.launch_count
    dev_ms, its, worst = 0.0, 0, (0.0, 0.0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for k in range(1, len(frames)):
            pair(k)
            r = ctx.icp_run(sl, ap, np.eye(4))
            ms, it = ctx.last_run_timing()
            dev_ms += ms
            its += it
            truth = np.linalg.inv(frames[k]["pose"]) @ frames[k - 1]["pose"]
            e = syn.pose_error(r["T"], truth)
            worst = (max(worst[0], e[0]), max(worst[1], e[1]))
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    n_valid = int(np.mean([f["valid"].sum() for f in frames]))
    peak, peak_src = measured_peak()
    k_ms = dev_ms / max(its, 1)
    abytes = 64 * n_valid
    line = {"metric": "ICP iters/sec, RGB-D projective aligner (C3)", "value": its / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": 1,
            "steps": args.steps, "warmup": 3, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C3: RGB-D projective association, 640x480 depth, 30-frame sequence: %d aligner calls x %d iters per step"
                                   % (n_pairs, iters), "valid_points_per_frame": n_valid,
                       "l2": "every call works on freshly uploaded frames (2 x %d points)" % frames[0]["points"].shape[0]},
            "roofline": {"bound": "hbm", "achieved": abytes / (k_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": abytes / (k_ms * 1e-3) / 1e9 / peak, "traffic": None,
                         "kernel": "one ICP iteration: proj_find_kernel + lin_tiles_kernel + icp_solve_kernel (640x480 frames are launch-latency bound)",
                         "kernel_ms": k_ms, "algorithmic_bytes_per_launch": abytes, "peak_source": peak_src},
            "e2e": {"value": its / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(n_pairs * 2 * frames[0]["points"].shape[0] * 25),
                    "d2h_bytes_per_step": n_pairs * (64 + 64 * iters), "ms_per_step": 1e3 * e2e_s / args.steps},
            "gpu_launches": int(ctx.launch_count - launches0), "clocks": sampler.stop(),
            "result_check": {"worst_pose_error_rad_m": list(worst), "iterations": its}, "cpu_baseline": None}
    print(json.dumps(line))
    ctx.close()


def run_c4(args):
    """Pose-graph Gauss-Newton (config C4): one GN iteration per step on the 100k-pose / 500k-factor graph."""
    import torch
    from srrg2_slam_interfaces_b200 import capi as A
    from srrg2_slam_interfaces_b200 import synthetic as syn
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_poses, n_factors = (args.points or 100000), 5 * (args.points or 100000)
    g = syn.make_pose_graph3d(n_poses, n_factors, seed=4)
    ctx = A.Context(3, local)
    if world > 1:
        uid = [ctx.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], rank, world)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    ctx.pgo_upload(g["guess"], g["fixed"], g["ij"], g["Z"], g["Omega"])
    # Solver::compute() as optimize() runs it: damped Gauss-Newton until |dx|_inf < 1e-6 (BASELINE C4: or 10 iterations;
    # the synthetic guess -- integrated noisy odometry, 47 m mean error -- needs more, so the cap is args.steps * 5)
    hist = ctx.pgo_optimize(max_iterations=max(10, 5 * args.steps), dx_tolerance=1e-6, max_cg_iterations=5000)
    poses = ctx.pgo_download().astype(np.float64)
    e2e_s = time.perf_counter() - t0
    if rank == 0:
        F, V = g["ij"].shape[0], n_poses
        lin_bytes = 312 * F + 168 * V
        lin_ms = float(np.median([h["linearize_ms"] for h in hist]))
        step_ms = float(np.mean([h["linearize_ms"] + h["solve_ms"] for h in hist]))
        err = np.linalg.norm(poses[:, :3, 3] - g["truth"][:, :3, 3], axis=1)
        err0 = np.linalg.norm(g["guess"][:, :3, 3].astype(np.float64) - g["truth"][:, :3, 3], axis=1)
        peak, peak_src = measured_peak()
        line = {"metric": "pose-graph GN iterations/s (C4)", "value": 1e3 / step_ms, "unit": "GN iters/s", "n_gpus": world,
                "steps": len(hist), "warmup": 0, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "C4: pose-graph GN, %d SE(3) poses / %d factors (synthetic Manhattan-3D)" % (V, F),
                           "solver": "Levenberg-Marquardt steps, inexact block-Jacobi PCG (fp64), atomics-free assembly", "sharding": "factors round-robin over ranks" if world > 1 else "none"},
                "roofline": {"bound": "hbm", "achieved": lin_bytes / (lin_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": lin_bytes / (lin_ms * 1e-3) / 1e9 / peak, "traffic": None, "kernel": "pgo_linearize_kernel",
                             "kernel_ms": lin_ms, "algorithmic_bytes_per_launch": lin_bytes, "peak_source": peak_src},
                "e2e": {"value": len(hist) / e2e_s, "unit": "GN iters/s", "h2d_bytes_per_step": int(sum(g[k].nbytes for k in ("guess", "ij", "Z", "Omega")) // len(hist)),
                        "d2h_bytes_per_step": int(poses.shape[0] * 64 // len(hist)), "ms_per_step": 1e3 * e2e_s / len(hist)},
                "gpu_launches": int(ctx.launch_count), "clocks": sampler.stop(),
                "result_check": {"converged": bool(hist[-1]["dx_norm_inf"] < 1e-6), "iterations": len(hist),
                                 "accepted": [h["accepted"] for h in hist], "lambda": [h["lambda"] for h in hist],
                                 "chi": [h["chi"] for h in hist], "dx_norm_inf": [h["dx_norm_inf"] for h in hist],
                                 "cg_iterations": [h["cg_iterations"] for h in hist], "solve_ms": [round(h["solve_ms"], 3) for h in hist],
                                 "position_error_mean_m": [float(err0.mean()), float(err.mean())]},
                "cpu_baseline": None}
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.config == "c3":
        run_c3(a)
    elif a.config == "c4":
        run_c4(a)
    else:
        run_aligner(a)
