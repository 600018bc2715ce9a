"""candidate-batched loop closing.  MultiLoopDetectorBruteForce_::compute
(R/registration/loop_detector/multi_loop_detector_brute_force_impl.cpp:63-133) aligns K candidate local maps against
the source local map one after the other and gates each result; srrg2b_closure_batch has the K runs in flight at once
(one context per candidate, the source map's cloud and index lent to all of them by srrg2b_share_fixed).  Against the
oracle's serial loop: the same verdicts in candidate order, the same poses and IterationStats-derived numbers, bit for
bit, and every candidate's correspondences stay retrievable from its context."""
import numpy as np
import pytest

from srrg2_slam_interfaces_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _candidates3d(n_fixed, sizes, seed=41):
    """source map + K target maps: most are independent samples of the same surfaces seen from different poses, one is
    a different place altogether (has to be dropped), one is too sparse (not enough inliers)."""
    base = syn.make_icp3d(n_fixed, 16, seed=seed)
    cands = []
    rng = np.random.default_rng(7)
    for k, n in enumerate(sizes):
        t = rng.uniform(-0.08, 0.08, size=3)
        rpy = rng.uniform(-0.02, 0.02, size=3)
        T_star = syn.iso3(t, rpy)
        d = syn.make_icp3d(16, n, seed=(seed if k != 2 else seed + 100), T_star=T_star, moving_stream=k + 1,
                           outlier_frac=0.05 if k != 4 else 0.45)
        guess = T_star @ syn.iso3(rng.uniform(-0.03, 0.03, size=3), rng.uniform(-0.01, 0.01, size=3))
        cands.append(dict(moving=d["moving"], normals=d["moving_normals"], guess=guess.astype(np.float32), T_star=T_star))
    return base, cands


def _run_both(oracle, capi, dim, fixed, fixed_n, cands, fp_kw, fa, ap_kw, gates):
    # ---- device: the source map is uploaded and indexed once, every candidate borrows it ----
    src = capi.Context(dim)
    src.set_cloud(capi.FIXED, 0, fixed, fixed_n)
    sl = [capi.make_slice(dim, 0, None, capi.finder_params(*fp_kw), capi.factor_params(*fa(capi)))]
    ap = capi.aligner_params(**ap_kw)
    # (the lender needs an index for the finder radius before it is lent; a stand-alone find on a tiny cloud builds it)
    src.set_cloud(capi.MOVING, 0, cands[0]["moving"][:64], None if cands[0]["normals"] is None else cands[0]["normals"][:64])
    src.icp_run(sl, capi.aligner_params(max_iterations=1, min_num_inliers=0), np.eye(dim + 1))
    ctxs = []
    for c in cands:
        x = capi.Context(dim)
        x.share_fixed(0, src, 0)
        x.set_cloud(capi.MOVING, 0, c["moving"], c["normals"])
        ctxs.append(x)
    res = capi.closure_batch(ctxs, sl, ap, [c["guess"] for c in cands], capi.closure_params(*gates))
    corr = [x.get_correspondences(0, c["moving"].shape[0]) for x, c in zip(ctxs, cands)]
    # the batch must equal K separate compute() calls on the same contexts (and leave them reusable)
    solo = [x.icp_run(sl, ap, c["guess"]) for x, c in zip(ctxs, cands)]
    for x in ctxs:
        x.close()
    src.close()
    # ---- oracle: the reference's serial loop ----
    F = oracle.CloudRef(fixed, fixed_n)
    refs = [oracle.CloudRef(c["moving"], c["normals"]) for c in cands]  # (a slice only holds the pointers)
    osl = [[oracle.make_slice(F, m, None, oracle.finder_params(*fp_kw), oracle.factor_params(*fa(oracle)), dim=dim)] for m in refs]
    ores = oracle.closure_loop(dim, osl, oracle.aligner_params(**ap_kw), [c["guess"] for c in cands], *gates)
    for k, (g, o, s) in enumerate(zip(res, ores, solo)):
        assert g["aligner_status"] == o["aligner_status"] == s["status"], (k, g, o)
        assert np.array_equal(g["T"], o["T"]) and np.array_equal(g["T"], s["T"]), k
        assert g["iterations"] == o["iterations"], k
        assert g["verdict"] == o["verdict"], (k, g["verdict"], o["verdict"], g, o)
        if g["aligner_status"] == 0:
            assert g["num_correspondences"] == o["num_correspondences"], (k, g, o)
            assert g["num_inliers"] == o["num_inliers"], k
            assert np.float32(g["chi_inliers"]) == np.float32(o["chi_inliers"]), (k, g["chi_inliers"], o["chi_inliers"])
    return res, ores, corr, (osl, refs, F)


def test_closure_batch_3d(oracle, capi):
    base, cands = _candidates3d(60000, [30000, 20000, 25000, 300, 30000, 45000])
    ap_kw = dict(max_iterations=12, min_num_inliers=50, enable_inlier_only_runs=True, keep_only_inlier_correspondences=True)
    res, ores, corr, osl = _run_both(oracle, capi, 3, base["fixed"], base["fixed_normals"], cands, (0.3, 0.8),
                                     lambda m: (m.FACTOR_PLANE, m.ROB_HUBER, 0.01), ap_kw, (500, 0.005, 0.7))
    verdicts = [r["verdict"] for r in res]
    assert verdicts[0] == capi.CLOSURE_ACCEPT and verdicts[1] == capi.CLOSURE_ACCEPT and verdicts[5] == capi.CLOSURE_ACCEPT
    assert verdicts[2] != capi.CLOSURE_ACCEPT      # a different place
    assert verdicts[3] != capi.CLOSURE_ACCEPT      # 300 points: below relocalize_min_inliers
    for k in (0, 1, 5):
        assert max(syn.pose_error(res[k]["T"], cands[k]["T_star"])) < 2e-3
    # the candidates' correspondences stayed in their contexts (pruned to inliers)
    o = oracle.icp_run(3, osl[0][0], oracle.aligner_params(**ap_kw), cands[0]["guess"])
    assert np.array_equal(corr[0][0], o["correspondences"][0][0]) and np.array_equal(corr[0][1], o["correspondences"][0][1])
    assert len(corr[0][0]) == res[0]["num_correspondences"]


def test_closure_batch_2d(oracle, capi):
    """SE(2) laser maps (the srrg2_laser_slam_2d loop closer): small clouds, where the K runs overlap on the device; no
    pruning, so the ratio gate sees all correspondences."""
    fixed = None
    cands = []
    rng = np.random.default_rng(11)
    for k in range(8):
        T_star = syn.iso2(*rng.uniform(-0.1, 0.1, size=2), rng.uniform(-0.03, 0.03))
        d = syn.make_icp2d(4000, 1500 + 300 * k, seed=9 if k != 5 else 10, T_star=T_star, noise=0.004 if k != 6 else 0.05, paired=False)
        if fixed is None:
            fixed = (d["fixed"], d.get("fixed_normals"))
        guess = (T_star @ syn.iso2(*rng.uniform(-0.05, 0.05, size=2), rng.uniform(-0.02, 0.02))).astype(np.float32)
        cands.append(dict(moving=d["moving"], normals=d.get("moving_normals"), guess=guess, T_star=T_star))
    ap_kw = dict(max_iterations=15, min_num_inliers=20)
    res, ores, corr, _ = _run_both(oracle, capi, 2, fixed[0], fixed[1], cands, (0.5, 0.7),
                                   lambda m: (m.FACTOR_PLANE if fixed[1] is not None else m.FACTOR_P2P, m.ROB_CAUCHY, 0.05), ap_kw,
                                   (400, 0.01, 0.6))
    assert any(r["verdict"] == capi.CLOSURE_ACCEPT for r in res) and any(r["verdict"] != capi.CLOSURE_ACCEPT for r in res)


def test_share_fixed_contract(capi):
    d = syn.make_icp3d(5000, 5000, seed=3)
    a, b = capi.Context(3), capi.Context(3)
    with pytest.raises(capi.Srrg2bError):
        b.share_fixed(0, a, 0)              # nothing to lend yet
    a.set_cloud(capi.FIXED, 0, d["fixed"], d["fixed_normals"])
    b.share_fixed(0, a, 0)
    b.set_cloud(capi.MOVING, 0, d["moving"], d["moving_normals"])
    a.set_cloud(capi.MOVING, 0, d["moving"], d["moving_normals"])
    sl = [capi.make_slice(3, 0, None, capi.finder_params(0.3, 0.8), capi.factor_params(capi.FACTOR_PLANE, capi.ROB_HUBER, 0.01))]
    ap = capi.aligner_params(max_iterations=6)
    rb = b.icp_run(sl, ap, np.eye(4))       # the borrower builds a private index (the lender had none yet)
    ra = a.icp_run(sl, ap, np.eye(4))
    assert np.array_equal(ra["T"], rb["T"]) and ra["stats"] == rb["stats"]
    # a borrower that sets its own fixed cloud lets go of the loan; the lender is untouched
    b.set_cloud(capi.FIXED, 0, d["fixed"][:2500], d["fixed_normals"][:2500])
    rb2 = b.icp_run(sl, ap, np.eye(4))
    ra2 = a.icp_run(sl, ap, np.eye(4))
    assert np.array_equal(ra2["T"], ra["T"]) and ra2["stats"] == ra["stats"]
    assert rb2["stats"] != ra["stats"]
    b.close()
    a.close()


def test_closure_batch_edge_cases(oracle, capi):
    """Empty batch, a candidate with an empty cloud, a candidate whose every point is masked out, a context listed twice:
    the reference's loop would `continue` on the aligner status for the degenerate maps (:80-84); misuse is an error."""
    d = syn.make_icp3d(8000, 3000, seed=13)
    sl = [capi.make_slice(3, 0, None, capi.finder_params(0.3, 0.8), capi.factor_params(capi.FACTOR_PLANE, capi.ROB_HUBER, 0.01))]
    ap = capi.aligner_params(max_iterations=5, min_num_inliers=10)
    cp = capi.closure_params(50, 0.01, 0.5)
    assert capi.closure_batch([], sl, ap, [], cp) == []
    src = capi.Context(3)
    src.set_cloud(capi.FIXED, 0, d["fixed"], d["fixed_normals"])
    clouds = [(d["moving"], d["moving_normals"], None),
              (np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32), None),
              (d["moving"][:500], d["moving_normals"][:500], np.zeros(500, np.uint8))]
    ctxs = []
    for m, n, v in clouds:
        x = capi.Context(3)
        x.share_fixed(0, src, 0)
        x.set_cloud(capi.MOVING, 0, m, n, v)
        ctxs.append(x)
    res = capi.closure_batch(ctxs, sl, ap, [np.eye(4, dtype=np.float32)] * 3, cp)
    F = oracle.CloudRef(d["fixed"], d["fixed_normals"])
    refs = [oracle.CloudRef(m, n, v) for m, n, v in clouds]
    osl = [[oracle.make_slice(F, r, None, oracle.finder_params(0.3, 0.8), oracle.factor_params(oracle.FACTOR_PLANE, oracle.ROB_HUBER, 0.01))]
           for r in refs]
    want = oracle.closure_loop(3, osl, oracle.aligner_params(max_iterations=5, min_num_inliers=10), [np.eye(4, dtype=np.float32)] * 3, 50, 0.01, 0.5)
    for g, w in zip(res, want):
        assert g["verdict"] == w["verdict"] and g["aligner_status"] == w["aligner_status"], (g, w)
        assert np.array_equal(g["T"], w["T"])
    assert res[1]["verdict"] == capi.CLOSURE_ALIGNER_DROP and res[2]["verdict"] == capi.CLOSURE_ALIGNER_DROP
    with pytest.raises(capi.Srrg2bError):
        capi.closure_batch([ctxs[0], ctxs[0]], sl, ap, [np.eye(4, dtype=np.float32)] * 2, cp)
    for x in ctxs:
        x.close()
    src.close()
