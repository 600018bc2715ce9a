"""world_size-2 gloo test of loop closing over several GPUs: the candidates of the brute-force
detector are independent alignments, dealt to the ranks in contiguous runs; every rank works through its run (here with
the oracle standing in for srrg2b_closure_batch) and the gathered results, in rank order, are the serial loop's
results in candidate order.  No collective on the data path."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _problem():
    from oracle import oracle as O
    from srrg2_slam_interfaces_b200 import synthetic as syn
    base = syn.make_icp2d(3000, 16, seed=9, paired=False)
    F = O.CloudRef(base["fixed"], base["fixed_normals"])
    keep, slices, guesses = [F], [], []
    for k in range(5):
        T_star = syn.iso2(0.02 * k, -0.015 * k, 0.004 * k)
        d = syn.make_icp2d(3000, 900 + 150 * k, seed=9 if k != 3 else 12, T_star=T_star, paired=False)
        keep.append(O.CloudRef(d["moving"], d["moving_normals"]))
        slices.append([O.make_slice(F, keep[-1], None, O.finder_params(0.5, 0.7), O.factor_params(O.FACTOR_PLANE, O.ROB_CAUCHY, 0.05), dim=2)])
        guesses.append(np.eye(3, dtype=np.float32))
    return O, keep, slices, guesses, O.aligner_params(max_iterations=8, min_num_inliers=10)


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from srrg2_slam_interfaces_b200.sharding import candidate_shard
    O, keep, slices, guesses, ap = _problem()
    O.set_threads(2)
    mine = candidate_shard(len(slices), rank, world)
    res = O.closure_loop(2, [slices[k] for k in mine], ap, [guesses[k] for k in mine], 200, 0.01, 0.5)
    packed = [(k, r["verdict"], r["aligner_status"], r["num_correspondences"], r["num_inliers"], float(r["chi_inliers"]), r["T"].tolist())
              for k, r in zip(mine, res)]
    gathered = [None] * world
    dist.all_gather_object(gathered, packed)
    if rank == 0:
        out.put([x for part in gathered for x in part])
    dist.barrier()
    dist.destroy_process_group()


def test_candidate_shards_cover_in_order():
    sys.path.insert(0, ROOT)
    from srrg2_slam_interfaces_b200.sharding import candidate_shard
    for k in (0, 1, 5, 16, 17):
        for w in (1, 2, 3, 8):
            assert [c for r in range(w) for c in candidate_shard(k, r, w)] == list(range(k))


def test_sharded_closure_loop_equals_serial_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = out.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sys.path.insert(0, ROOT)
    O, keep, slices, guesses, ap = _problem()
    want = O.closure_loop(2, slices, ap, guesses, 200, 0.01, 0.5)
    assert [g[0] for g in got] == list(range(len(slices)))
    for g, w in zip(got, want):
        assert g[1] == w["verdict"] and g[2] == w["aligner_status"] and g[3] == w["num_correspondences"] and g[4] == w["num_inliers"]
        assert np.float32(g[5]) == np.float32(w["chi_inliers"]) and np.array_equal(np.asarray(g[6], dtype=np.float32), w["T"])
    assert any(w["verdict"] == O.CLOSURE_ACCEPT for w in want) and any(w["verdict"] != O.CLOSURE_ACCEPT for w in want)
