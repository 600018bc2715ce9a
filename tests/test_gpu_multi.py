"""Multi-GPU parity (needs >= 2 CUDA devices, skipped otherwise): the moving cloud is sharded by
rank, the integer accumulators are all-reduced over the peer mailboxes inside the solve kernel; every
rank must end with the oracle's single-process pose / IterationStats bit for bit, the per-rank
correspondence lists concatenate to the oracle's list, and the stand-alone linearize call returns the
GLOBAL sums on every rank."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

N_FIXED, N_MOVING, SEED = 60000, 50001, 17
KW = dict(max_iterations=12, min_num_inliers=10)


def _worker(rank, world, uid_q, out_q):
    sys.path.insert(0, ROOT)
    from srrg2_slam_interfaces_b200 import capi as A
    from srrg2_slam_interfaces_b200 import synthetic as syn
    from srrg2_slam_interfaces_b200.sharding import shard_range
    d = syn.make_icp3d(N_FIXED, N_MOVING, seed=SEED)
    ctx = A.Context(3, rank)
    if rank == 0:
        uid = ctx.unique_id()
        for _ in range(world - 1):
            uid_q.put(uid)
    else:
        uid = uid_q.get(timeout=60)
    ctx.comm_init(uid, rank, world)
    b, e = shard_range(N_MOVING, rank, world)
    ctx.set_cloud(A.FIXED, 0, d["fixed"], d["fixed_normals"])
    ctx.set_cloud(A.MOVING, 0, d["moving"][b:e], d["moving_normals"][b:e], index_offset=b, n_global=N_MOVING)
    sl = [A.make_slice(3, 0, None, A.finder_params(0.3, 0.8), A.factor_params(A.FACTOR_PLANE, A.ROB_HUBER, 0.01))]
    res = None
    for _ in range(2):
        res = ctx.icp_run(sl, A.aligner_params(**KW), np.eye(4))
    corr = ctx.get_correspondences(0, e - b)
    # stand-alone find + linearize at a fixed transform: every rank must return the global accumulators
    S = syn.iso3([0.02, -0.01, 0.03], [0.004, -0.003, 0.005])
    fp, fa = A.finder_params(0.3, 0.8), A.factor_params(A.FACTOR_PLANE, A.ROB_HUBER, 0.01)
    ctx.find_correspondences(0, S, fp, e - b)
    lin = ctx.linearize(0, S, fp, fa, n_moving=e - b)
    out_q.put((rank, res["T"], res["status"], res["stats"], corr, lin["acc"], lin["H"], lin["b"]))
    ctx.close()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_run_equals_oracle(oracle, world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    from srrg2_slam_interfaces_b200 import synthetic as syn
    ctx = mp.get_context("spawn")
    uid_q, out_q = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, uid_q, out_q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted([out_q.get(timeout=300) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    d = syn.make_icp3d(N_FIXED, N_MOVING, seed=SEED)
    F = oracle.CloudRef(d["fixed"], d["fixed_normals"])
    M = oracle.CloudRef(d["moving"], d["moving_normals"])
    o = oracle.icp_run(3, [oracle.make_slice(F, M, None, oracle.finder_params(0.3, 0.8),
                                             oracle.factor_params(oracle.FACTOR_PLANE, oracle.ROB_HUBER, 0.01))],
                       oracle.aligner_params(**KW), np.eye(4))
    S = syn.iso3([0.02, -0.01, 0.03], [0.004, -0.003, 0.005])
    ofp = oracle.finder_params(0.3, 0.8)
    ofi, _ = oracle.find(oracle.Index(F, oracle.NN_KDTREE), F, M, S, ofp)
    ol = oracle.linearize(F, M, ofi, S, ofp, oracle.factor_params(oracle.FACTOR_PLANE, oracle.ROB_HUBER, 0.01))
    for rank, T, status, stats, corr, acc, H, b in results:
        assert status == o["status"]
        assert np.array_equal(T, o["T"])
        assert stats == o["stats"]
        assert np.array_equal(acc, ol["acc"]), "rank %d: stand-alone linearize did not return the global sums" % rank
        assert np.array_equal(H, ol["H"]) and np.array_equal(b, ol["b"])
    fi = np.concatenate([r[4][0] for r in results])
    mi = np.concatenate([r[4][1] for r in results])
    rs = np.concatenate([r[4][2] for r in results])
    assert np.array_equal(fi, o["correspondences"][0][0])
    assert np.array_equal(mi, o["correspondences"][0][1])
    assert np.array_equal(rs, o["correspondences"][0][2])


# ---- pose graph, factor-sharded (row e2): every rank ends with the single-process result within the north-star
# tolerances (the all-reduce changes the fp64 summation order of H / b, nothing else) ----
def _pgo_worker(rank, world, uid_q, out_q):
    sys.path.insert(0, ROOT)
    from srrg2_slam_interfaces_b200 import capi as A
    from srrg2_slam_interfaces_b200 import synthetic as syn
    g = syn.make_pose_graph3d(1500, 7000, seed=4, box=(10, 10, 3))
    ctx = A.Context(3, rank)
    if world > 1:
        if rank == 0:
            uid = ctx.unique_id()
            for _ in range(world - 1):
                uid_q.put(uid)
        else:
            uid = uid_q.get(timeout=60)
        ctx.comm_init(uid, rank, world)
    ctx.pgo_upload(g["guess"], g["fixed"], g["ij"], g["Z"], g["Omega"])
    st = [ctx.pgo_iterate(max_cg_iterations=4000, cg_tolerance=1e-11) for _ in range(4)]
    out_q.put((rank, ctx.pgo_download(), [(s["chi"], s["dx_norm_inf"]) for s in st]))
    ctx.close()


def test_sharded_pose_graph_equals_oracle():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from oracle import pgo_oracle as P
    from srrg2_slam_interfaces_b200 import synthetic as syn
    world = 2
    ctx = mp.get_context("spawn")
    uid_q, out_q = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_pgo_worker, args=(r, world, uid_q, out_q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted([out_q.get(timeout=300) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = syn.make_pose_graph3d(1500, 7000, seed=4, box=(10, 10, 3))
    poses = g["guess"].astype(np.float64)
    chis = []
    for _ in range(4):
        poses, so = P.gn_step(poses, g["ij"], g["Z"].astype(np.float64), g["Omega"].astype(np.float64), g["fixed"])
        chis.append(so["chi"])
    assert np.array_equal(results[0][1], results[1][1]), "ranks disagree on the optimised poses"
    for rank, got, st in results:
        got = got.astype(np.float64)
        assert np.abs(got[:, :3, 3] - poses[:, :3, 3]).max() < 1e-4
        for (chi, _), co in zip(st, chis):
            assert abs(chi - co) <= 1e-6 * co


# ---- loop closing over several GPUs (SURVEY 8f N3 + 8e): the detector's candidates are independent alignments; every
# rank batches its own contiguous run of them (srrg2b_closure_batch on its GPU), results concatenated in rank order ----
def _closure_problem(syn):
    base = syn.make_icp2d(20000, 16, seed=9, paired=False)
    cands = []
    for k in range(6):
        T_star = syn.iso2(0.02 * k, -0.015 * k, 0.004 * k)
        d = syn.make_icp2d(20000, 4000 + 500 * k, seed=9 if k != 4 else 12, T_star=T_star, paired=False)
        cands.append((d["moving"], d["moving_normals"]))
    return base, cands


def _closure_worker(rank, world, out_q):
    sys.path.insert(0, ROOT)
    from srrg2_slam_interfaces_b200 import capi as A
    from srrg2_slam_interfaces_b200 import synthetic as syn
    from srrg2_slam_interfaces_b200.sharding import candidate_shard
    base, cands = _closure_problem(syn)
    sl = [A.make_slice(2, 0, None, A.finder_params(0.5, 0.7), A.factor_params(A.FACTOR_PLANE, A.ROB_CAUCHY, 0.05))]
    ap = A.aligner_params(max_iterations=10, min_num_inliers=10)
    mine = candidate_shard(len(cands), rank, world)
    src = A.Context(2, rank)
    src.set_cloud(A.FIXED, 0, base["fixed"], base["fixed_normals"])
    src.set_cloud(A.MOVING, 0, cands[0][0][:64], cands[0][1][:64])
    src.icp_run(sl, A.aligner_params(max_iterations=1, min_num_inliers=0), np.eye(3))
    ctxs = []
    for k in mine:
        x = A.Context(2, rank)
        x.share_fixed(0, src, 0)
        x.set_cloud(A.MOVING, 0, cands[k][0], cands[k][1])
        ctxs.append(x)
    res = A.closure_batch(ctxs, sl, ap, [np.eye(3, dtype=np.float32)] * len(mine), A.closure_params(500, 0.01, 0.5))
    out_q.put((rank, [(k, r["verdict"], r["aligner_status"], r["num_correspondences"], r["num_inliers"], float(r["chi_inliers"]), r["T"])
                      for k, r in zip(mine, res)]))
    for x in ctxs:
        x.close()
    src.close()


def test_candidates_sharded_over_gpus_equal_serial_loop(oracle):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from srrg2_slam_interfaces_b200 import synthetic as syn
    world = 2
    ctx = mp.get_context("spawn")
    out_q = ctx.Queue()
    procs = [ctx.Process(target=_closure_worker, args=(r, world, out_q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted([out_q.get(timeout=300) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got = [x for _, part in results for x in part]
    base, cands = _closure_problem(syn)
    F = oracle.CloudRef(base["fixed"], base["fixed_normals"])
    refs = [oracle.CloudRef(m, n) for m, n in cands]
    osl = [[oracle.make_slice(F, m, None, oracle.finder_params(0.5, 0.7), oracle.factor_params(oracle.FACTOR_PLANE, oracle.ROB_CAUCHY, 0.05), dim=2)]
           for m in refs]
    want = oracle.closure_loop(2, osl, oracle.aligner_params(max_iterations=10, min_num_inliers=10), [np.eye(3, dtype=np.float32)] * len(cands),
                               500, 0.01, 0.5)
    assert [g[0] for g in got] == list(range(len(cands)))
    for g, w in zip(got, want):
        assert g[1] == w["verdict"] and g[2] == w["aligner_status"], (g[:3], w["verdict"])
        assert np.array_equal(g[6], w["T"])
        if w["aligner_status"] == 0:
            assert g[3] == w["num_correspondences"] and g[4] == w["num_inliers"] and np.float32(g[5]) == np.float32(w["chi_inliers"])
