"""Factor sharding of the pose-graph path (row e2), host logic on CPU with gloo, world size 2: the ranks take the
factors f = rank, rank + world, ...; each assembles H / b / chi from ITS factors only (here: the numpy oracle's
per-factor terms), one all-reduce(sum) of H values, b and chi gives every rank the single-process system."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from oracle import pgo_oracle as P
    from srrg2_slam_interfaces_b200 import synthetic as syn
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = syn.make_pose_graph3d(300, 1200, seed=4, box=(6, 6, 2))
    poses, Z, Om = g["guess"].astype(np.float64), g["Z"].astype(np.float64), g["Omega"].astype(np.float64)
    mine = np.arange(rank, g["ij"].shape[0], world)
    H, b, chi, _ = P.linearize(poses, g["ij"][mine], Z[mine], Om[mine], g["fixed"])
    # (the gauge rows are identity on every rank: take them out before the sum, put them back after)
    fixed_rows = np.repeat(np.asarray(g["fixed"], bool), 6)
    Hd = H.toarray()
    Hd[fixed_rows, fixed_rows] = 0.0
    t = torch.from_numpy(np.concatenate([Hd.ravel(), b, [chi]]))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    out = t.numpy()
    n = b.size
    Hs = out[:n * n].reshape(n, n)
    Hs[fixed_rows, fixed_rows] = 1.0
    q.put((rank, Hs, out[n * n:n * n + n], out[-1]))
    dist.destroy_process_group()


def test_factor_sharded_system_equals_single_process():
    from oracle import pgo_oracle as P
    from srrg2_slam_interfaces_b200 import synthetic as syn
    world, port = 2, 29533
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = syn.make_pose_graph3d(300, 1200, seed=4, box=(6, 6, 2))
    H, b, chi, _ = P.linearize(g["guess"].astype(np.float64), g["ij"], g["Z"].astype(np.float64), g["Omega"].astype(np.float64), g["fixed"])
    Hd = H.toarray()
    for rank, Hs, bs, cs in res:
        assert np.allclose(Hs, Hd, rtol=1e-12, atol=1e-9 * np.abs(Hd).max())
        assert np.allclose(bs, b, rtol=1e-12, atol=1e-9 * np.abs(b).max())
        assert abs(cs - chi) <= 1e-12 * chi
    assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], res[1][2])
