"""world_size-2 gloo test of the multi-GPU design: the moving cloud is split by
contiguous index ranges, each rank accumulates its exact fixed-point partial sums (here with the
oracle standing in for the per-rank kernel), and an integer all-reduce reproduces the unsharded
accumulators bit for bit -- which is what lets every rank solve the 6x6 redundantly."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from srrg2_slam_interfaces_b200 import synthetic as syn
    from srrg2_slam_interfaces_b200.sharding import shard_range
    O.set_threads(2)
    d = syn.make_icp3d(6000, 5001, seed=21)
    F = O.CloudRef(d["fixed"], d["fixed_normals"])
    fp, fa = O.finder_params(0.4, 0.8), O.factor_params(O.FACTOR_PLANE, O.ROB_HUBER, 0.02)
    ix = O.Index(F)
    S = np.eye(4, dtype=np.float32)
    b, e = shard_range(5001, rank, world)
    Ms = O.CloudRef(d["moving"][b:e], d["moving_normals"][b:e])
    fidx, _ = O.find(ix, F, Ms, S, fp)
    # the scale exponents must come from GLOBAL quantities: max |m|^2 and max |n|^2 over the whole moving cloud
    bound = torch.tensor([O.radius_bound2(Ms), max(O.normal_bound2(Ms), O.normal_bound2(F))], dtype=torch.float32)
    dist.all_reduce(bound, op=dist.ReduceOp.MAX)
    assert float(bound[0]) == O.radius_bound2(O.CloudRef(d["moving"], d["moving_normals"]))
    part = O.linearize(F, Ms, fidx, S, fp, fa, radius_bound2=float(bound[0]), normal_bound2=float(bound[1]))
    acc = torch.from_numpy(part["acc"].copy())
    dist.all_reduce(acc, op=dist.ReduceOp.SUM)
    if rank == 0:
        M = O.CloudRef(d["moving"], d["moving_normals"])
        gfidx, _ = O.find(ix, F, M, S, fp)
        full = O.linearize(F, M, gfidx, S, fp, fa)
        out.put((acc.numpy().tolist(), full["acc"].tolist(), np.array_equal(gfidx[b:e], fidx)))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_and_balance():
    sys.path.insert(0, ROOT)
    from srrg2_slam_interfaces_b200.sharding import shard_range
    for n in (0, 1, 7, 8, 1000003):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [e - b for b, e in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_sharded_integer_sums_equal_unsharded_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    reduced, full, same_idx = res
    assert same_idx
    assert reduced == full
