"""GPU parity: CUDA path through the C ABI vs the CPU oracle on the same seeded inputs.
Bars (BASELINE.json north_star): correspondence indices bit-exact; H/b/chi exact (integer sums);
pose / chi^2 far inside 1e-5 rad / 1e-4 m / 1e-6 relative (they are in fact bit-identical)."""
import numpy as np
import pytest

from srrg2_slam_interfaces_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _dense_from_compact(n, fi, mi, rs):
    d = np.full(n, -1, dtype=np.int32)
    r = np.zeros(n, dtype=np.float32)
    d[mi] = fi
    r[mi] = rs
    return d, r


def _setup3d(oracle, capi, nf, nm, seed=2, valid_frac=None):
    d = syn.make_icp3d(nf, nm, seed=seed)
    fv = mv = None
    if valid_frac is not None:
        rng = np.random.default_rng(seed + 100)
        fv = (rng.uniform(size=nf) < valid_frac).astype(np.uint8)
        mv = (rng.uniform(size=nm) < valid_frac).astype(np.uint8)
    F = oracle.CloudRef(d["fixed"], d["fixed_normals"], fv)
    M = oracle.CloudRef(d["moving"], d["moving_normals"], mv)
    ctx = capi.Context(3)
    ctx.set_cloud(capi.FIXED, 0, d["fixed"], d["fixed_normals"], fv)
    ctx.set_cloud(capi.MOVING, 0, d["moving"], d["moving_normals"], mv)
    return d, F, M, ctx


@pytest.mark.parametrize("nf,nm,valid", [(5000, 4000, None), (20000, 30000, 0.9), (1, 10, None), (300, 1, None)])
def test_find_bit_exact_3d(oracle, capi, nf, nm, valid):
    d, F, M, ctx = _setup3d(oracle, capi, nf, nm, valid_frac=valid)
    ix = oracle.Index(F, oracle.NN_KDTREE)
    for S in (np.eye(4), syn.iso3([0.3, -0.2, 0.1], [0.02, -0.03, 0.05]), d["T_star"]):
        for md, nc in ((0.3, 0.8), (1.5, -2.0), (0.05, 0.95)):
            ofi, ors = oracle.find(ix, F, M, S, oracle.finder_params(md, nc))
            fi, mi, rs = ctx.find_correspondences(0, S, capi.finder_params(md, nc), nm)
            assert np.all(np.diff(mi) > 0)
            gfi, grs = _dense_from_compact(nm, fi, mi, rs)
            assert np.array_equal(gfi, ofi)
            assert np.array_equal(grs, ors)
    ctx.close()


@pytest.mark.parametrize("factor", ["P2P", "PLANE"])
@pytest.mark.parametrize("rob", ["NONE", "HUBER", "CAUCHY", "CLAMP", "SATURATED"])
@pytest.mark.parametrize("variable", [0, 1])
def test_linearize_exact_3d(oracle, capi, factor, rob, variable):
    nf, nm = 20000, 20000
    d, F, M, ctx = _setup3d(oracle, capi, nf, nm, valid_frac=0.95)
    S = syn.iso3([0.02, -0.01, 0.03], [0.004, -0.003, 0.005])
    ofp = oracle.finder_params(0.3, 0.8)
    gfp = capi.finder_params(0.3, 0.8)
    ofa = oracle.factor_params(getattr(oracle, "FACTOR_" + factor), getattr(oracle, "ROB_" + rob), 0.01, 2.0, 0.5)
    gfa = capi.factor_params(getattr(capi, "FACTOR_" + factor), getattr(capi, "ROB_" + rob), 0.01, 2.0, 0.5)
    ix = oracle.Index(F, oracle.NN_KDTREE)
    ofi, _ = oracle.find(ix, F, M, S, ofp)
    o = oracle.linearize(F, M, ofi, S, ofp, ofa, variable=variable)
    ctx.find_correspondences(0, S, gfp, nm)
    g = ctx.linearize(0, S, gfp, gfa, variable=variable, n_moving=nm)
    assert np.array_equal(g["acc"], o["acc"])
    assert np.array_equal(g["H"], o["H"]) and np.array_equal(g["b"], o["b"])
    assert g["stats"]["num_inliers"] == o["stats"]["num_inliers"]
    assert g["stats"]["num_outliers"] == o["stats"]["num_outliers"]
    assert g["stats"]["chi_inliers"] == o["stats"]["chi_inliers"]
    assert g["stats"]["chi_outliers"] == o["stats"]["chi_outliers"]
    sel = ofi >= 0
    assert np.array_equal(g["status"], o["status"][sel])
    assert np.array_equal(g["chi"], o["chi"][sel])
    ctx.close()


def _run_both(oracle, capi, dim, d, oap, gap, ofp, gfp, ofa, gfa, T0):
    F = oracle.CloudRef(d["fixed"], d["fixed_normals"])
    M = oracle.CloudRef(d["moving"], d["moving_normals"])
    o = oracle.icp_run(dim, [oracle.make_slice(F, M, None, ofp, ofa, dim=dim)], oap, T0)
    ctx = capi.Context(dim)
    ctx.set_cloud(capi.FIXED, 0, d["fixed"], d["fixed_normals"])
    ctx.set_cloud(capi.MOVING, 0, d["moving"], d["moving_normals"])
    g = ctx.icp_run([capi.make_slice(dim, 0, None, gfp, gfa)], gap, T0)
    corr = ctx.get_correspondences(0, d["moving"].shape[0])
    ctx.close()
    return o, g, corr


def _assert_same_run(o, g, corr):
    assert g["status"] == o["status"]
    assert len(g["stats"]) == len(o["stats"])
    for a, b in zip(g["stats"], o["stats"]):
        assert a == b
    assert np.array_equal(g["T"], o["T"])
    ofi, omi, ors = o["correspondences"][0]
    assert np.array_equal(corr[0], ofi) and np.array_equal(corr[1], omi) and np.array_equal(corr[2], ors)


def test_icp_run_c2_small(oracle, capi):
    """Config C2 at 1/20 scale: 20 iterations, Huber, point+normal factor."""
    d = syn.make_icp3d(50000, 50000, seed=2)
    kw = dict(max_iterations=20, min_num_inliers=10)
    o, g, corr = _run_both(oracle, capi, 3, d, oracle.aligner_params(**kw), capi.aligner_params(**kw),
                           oracle.finder_params(0.3, 0.8), capi.finder_params(0.3, 0.8),
                           oracle.factor_params(oracle.FACTOR_PLANE, oracle.ROB_HUBER, 0.01),
                           capi.factor_params(capi.FACTOR_PLANE, capi.ROB_HUBER, 0.01), np.eye(4))
    _assert_same_run(o, g, corr)
    assert g["status"] == capi.ALIGNER_SUCCESS and len(g["stats"]) == 20
    rot, trans = syn.pose_error(g["T"], d["T_star"])
    assert rot < 2e-3 and trans < 2e-2


def test_icp_run_c1_2d(oracle, capi):
    """Config C1: 2D SE(2), 10k vs 10k, 1 iteration, no robustifier (plumbing config)."""
    d = syn.make_icp2d(10000, seed=1)
    for iters in (1, 10):
        kw = dict(max_iterations=iters, min_num_inliers=10)
        o, g, corr = _run_both(oracle, capi, 2, d, oracle.aligner_params(**kw), capi.aligner_params(**kw),
                               oracle.finder_params(0.5, 0.8), capi.finder_params(0.5, 0.8),
                               oracle.factor_params(oracle.FACTOR_PLANE, oracle.ROB_NONE),
                               capi.factor_params(capi.FACTOR_PLANE, capi.ROB_NONE), np.eye(3))
        _assert_same_run(o, g, corr)
    rot, trans = syn.pose_error(g["T"], d["T_star"])
    assert rot < 1e-3 and trans < 5e-3


def test_icp_run_termination_and_inlier_runs(oracle, capi):
    d = syn.make_icp3d(30000, 30000, seed=7)
    kw = dict(max_iterations=30, min_num_inliers=10, use_termination_criteria=True, enable_inlier_only_runs=True,
              keep_only_inlier_correspondences=True, num_correspondences_range=200, num_inliers_range=200,
              num_outliers_range=200)
    o, g, corr = _run_both(oracle, capi, 3, d, oracle.aligner_params(**kw), capi.aligner_params(**kw),
                           oracle.finder_params(0.3, 0.8), capi.finder_params(0.3, 0.8),
                           oracle.factor_params(oracle.FACTOR_P2P, oracle.ROB_CAUCHY, 0.01),
                           capi.factor_params(capi.FACTOR_P2P, capi.ROB_CAUCHY, 0.01), np.eye(4))
    _assert_same_run(o, g, corr)


def test_not_enough_correspondences(oracle, capi):
    d = syn.make_icp3d(2000, 2000, seed=3)
    far = syn.iso3([500.0, 0, 0], [0, 0, 0])
    kw = dict(max_iterations=5, min_num_inliers=10)
    o, g, corr = _run_both(oracle, capi, 3, d, oracle.aligner_params(**kw), capi.aligner_params(**kw),
                           oracle.finder_params(0.3, 0.8), capi.finder_params(0.3, 0.8),
                           oracle.factor_params(), capi.factor_params(), far)
    assert g["status"] == capi.ALIGNER_FAIL == o["status"]
    assert len(g["stats"]) == 0
    assert np.array_equal(g["T"], o["T"])


@pytest.mark.parametrize("mode", ["0", "1", "2"])
def test_temporal_coherence_modes_are_exact(oracle, capi, mode, monkeypatch):
    """SRRG2B_TRACK2 = 0 (never certify bounds), 1 (always), 2 (automatic): the skip of the NN search
    is an exact optimisation, so every mode must reproduce the oracle bit for bit -- including a
    second compute() that starts from the previous run's bounds, and a 2D multi-iteration run."""
    monkeypatch.setenv("SRRG2B_TRACK2", mode)
    d = syn.make_icp3d(40000, 40000, seed=13)
    kw = dict(max_iterations=14, min_num_inliers=10)
    F = oracle.CloudRef(d["fixed"], d["fixed_normals"])
    M = oracle.CloudRef(d["moving"], d["moving_normals"])
    ofp, ofa = oracle.finder_params(0.3, 0.8), oracle.factor_params(oracle.FACTOR_PLANE, oracle.ROB_HUBER, 0.01)
    ctx = capi.Context(3)
    ctx.set_cloud(capi.FIXED, 0, d["fixed"], d["fixed_normals"])
    ctx.set_cloud(capi.MOVING, 0, d["moving"], d["moving_normals"])
    gsl = [capi.make_slice(3, 0, None, capi.finder_params(0.3, 0.8),
                           capi.factor_params(capi.FACTOR_PLANE, capi.ROB_HUBER, 0.01))]
    T0s = [np.eye(4), syn.iso3([0.09, -0.05, 0.08], np.deg2rad([1.4, -1.0, 1.9])), np.eye(4)]
    for T0 in T0s:
        o = oracle.icp_run(3, [oracle.make_slice(F, M, None, ofp, ofa)], oracle.aligner_params(**kw), T0)
        g = ctx.icp_run(gsl, capi.aligner_params(**kw), T0)
        _assert_same_run(o, g, ctx.get_correspondences(0, 40000))
    # stand-alone finds interleaved with runs keep the bounds consistent
    for S in (d["T_star"], d["T_star"] @ syn.iso3([1e-4, 0, 0], [0, 1e-5, 0]), np.eye(4)):
        ofi, ors = oracle.find(oracle.Index(F), F, M, S, ofp)
        fi, mi, rs = ctx.find_correspondences(0, S, capi.finder_params(0.3, 0.8), 40000)
        gfi, grs = _dense_from_compact(40000, fi, mi, rs)
        assert np.array_equal(gfi, ofi) and np.array_equal(grs, ors)
    ctx.close()
    d2 = syn.make_icp2d(20000, 15000, seed=3, paired=False)
    o, g, corr = _run_both(oracle, capi, 2, d2, oracle.aligner_params(**kw), capi.aligner_params(**kw),
                           oracle.finder_params(0.4, 0.8), capi.finder_params(0.4, 0.8),
                           oracle.factor_params(oracle.FACTOR_P2P, oracle.ROB_SATURATED, 0.05),
                           capi.factor_params(capi.FACTOR_P2P, capi.ROB_SATURATED, 0.05), np.eye(3))
    _assert_same_run(o, g, corr)


def test_multicue_2d_two_scanners_plus_odom_prior(oracle, capi):
    """Config C5 shape at reduced size: 2 laser slices (robot_in_sensor != I) + odometry prior
    (Omega = diag(100,100,100), aligner_slice_odometry_prior.h:20), local map on the moving side."""
    d = syn.make_multicue2d(200000, n_beams=1080, seed=5)
    n_map = d["map"].shape[0]
    odom = syn.iso2(0.07, -0.04, np.deg2rad(1.2))
    kw = dict(max_iterations=10, min_num_inliers=10)
    ofp, ofa = oracle.finder_params(0.5, 0.7), oracle.factor_params(oracle.FACTOR_PLANE, oracle.ROB_CAUCHY, 0.05)
    gfp, gfa = capi.finder_params(0.5, 0.7), capi.factor_params(capi.FACTOR_PLANE, capi.ROB_CAUCHY, 0.05)
    M = oracle.CloudRef(d["map"], d["map_normals"])
    osl, gsl = [], []
    ctx = capi.Context(2)
    keep = []
    for k, sc in enumerate(d["scans"]):
        F = oracle.CloudRef(sc["points"], sc["normals"])
        keep.append(F)
        osl.append(oracle.make_slice(F, M, sc["robot_in_sensor"], ofp, ofa, dim=2))
        ctx.set_cloud(capi.FIXED, k, sc["points"], sc["normals"])
        ctx.set_cloud(capi.MOVING, k, d["map"], d["map_normals"])
        gsl.append(capi.make_slice(2, k, sc["robot_in_sensor"], gfp, gfa))
    osl.append(oracle.make_slice(prior_measurement=odom, prior_info_diag=np.full(3, 100.0), dim=2))
    gsl.append(capi.make_slice(2, prior_measurement=odom, prior_info_diag=np.full(3, 100.0)))
    o = oracle.icp_run(2, osl, oracle.aligner_params(**kw), np.eye(3))
    g = ctx.icp_run(gsl, capi.aligner_params(**kw), np.eye(3))
    assert g["status"] == o["status"] == capi.ALIGNER_SUCCESS
    assert g["stats"] == o["stats"]
    assert np.array_equal(g["T"], o["T"])
    for k in range(2):
        c = ctx.get_correspondences(k, n_map)
        oc = o["correspondences"][k]
        assert np.array_equal(c[0], oc[0]) and np.array_equal(c[1], oc[1]) and np.array_equal(c[2], oc[2])
    rot, trans = syn.pose_error(g["T"], d["T_star"])
    assert rot < 5e-3 and trans < 3e-2
    ctx.close()


def test_prior_only_reference_kat_on_gpu(oracle, capi):
    """The reference's own solver-level test (tests/test_motion_model_slice.cpp:81-85) through the CUDA path."""
    rng = np.random.default_rng(0)
    ctx = capi.Context(3)
    for _ in range(5):
        motion_prev = syn.iso3(rng.uniform(-1, 1, 3), rng.uniform(0, 6, 3)).astype(np.float32)
        inv = syn.inv_iso(motion_prev.astype(np.float64)).astype(np.float32)
        g = ctx.icp_run([capi.make_slice(3, prior_measurement=inv, prior_info_diag=np.ones(6))],
                        capi.aligner_params(max_iterations=10, min_num_inliers=0), np.eye(4))
        o = oracle.icp_run(3, [oracle.make_slice(prior_measurement=inv, prior_info_diag=np.ones(6), dim=3)],
                           oracle.aligner_params(max_iterations=10, min_num_inliers=0), np.eye(4))
        assert g["status"] == 0 and np.array_equal(g["T"], o["T"]) and g["stats"] == o["stats"]
        rot, trans = syn.pose_error(g["T"].astype(np.float64) @ motion_prev.astype(np.float64), np.eye(4))
        assert rot < 1e-5 and trans < 1e-5
    ctx.close()


@pytest.mark.parametrize("env", [{}, {"SRRG2B_NO_GRAPH": "1"}, {"SRRG2B_TRACK2": "1"}, {"SRRG2B_TRACK2": "0"},
                                 {"SRRG2B_FULL_ITERS": "0"}, {"SRRG2B_FULL_ITERS": "1", "SRRG2B_TRACK2": "1"},
                                 {"SRRG2B_FULL_ITERS": "100"}])
def test_execution_modes_match_oracle(oracle, capi, env, monkeypatch):
    """The same run through every execution mode of the device loop -- CUDA-graph replay (default), plain stream
    launches, bounds certified from the first iteration on (every later iteration is a coherence-check pass) or
    never (every iteration searches: from iteration 5 on through the single-kernel fallback), the single search
    kernel from the first iteration on (full searches and long work lists thread-per-query), the three-kernel
    pipeline throughout -- equals the oracle's, bit for bit."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    d = syn.make_icp3d(30000, 27001, seed=11)
    kw = dict(max_iterations=12, min_num_inliers=10)
    o, g, corr = _run_both(oracle, capi, 3, d, oracle.aligner_params(**kw), capi.aligner_params(**kw),
                           oracle.finder_params(0.3, 0.8), capi.finder_params(0.3, 0.8),
                           oracle.factor_params(oracle.FACTOR_PLANE, oracle.ROB_HUBER, 0.01),
                           capi.factor_params(capi.FACTOR_PLANE, capi.ROB_HUBER, 0.01), np.eye(4))
    _assert_same_run(o, g, corr)
    d2 = syn.make_icp2d(9000, 7000, seed=4, paired=False)
    o, g, corr = _run_both(oracle, capi, 2, d2, oracle.aligner_params(**kw), capi.aligner_params(**kw),
                           oracle.finder_params(0.4, 0.8), capi.finder_params(0.4, 0.8),
                           oracle.factor_params(oracle.FACTOR_PLANE, oracle.ROB_CAUCHY, 0.05),
                           capi.factor_params(capi.FACTOR_PLANE, capi.ROB_CAUCHY, 0.05), np.eye(3))
    _assert_same_run(o, g, corr)


def test_warm_context_equals_fresh_context(oracle, capi):
    """compute() twice on the same resident clouds (the second one starts from the first one's slots and
    bounds) must equal a fresh context bit for bit -- and both equal the oracle."""
    d = syn.make_icp3d(40000, 35000, seed=17)
    kw = dict(max_iterations=10, min_num_inliers=10)
    gsl = [capi.make_slice(3, 0, None, capi.finder_params(0.3, 0.8), capi.factor_params(capi.FACTOR_PLANE, capi.ROB_HUBER, 0.01))]
    runs = []
    for fresh in (True, False, False):
        if fresh:
            ctx = capi.Context(3)
            ctx.set_cloud(capi.FIXED, 0, d["fixed"], d["fixed_normals"])
            ctx.set_cloud(capi.MOVING, 0, d["moving"], d["moving_normals"])
        g = ctx.icp_run(gsl, capi.aligner_params(**kw), np.eye(4))
        runs.append((g, ctx.get_correspondences(0, 35000)))
    ctx.close()
    F, M = oracle.CloudRef(d["fixed"], d["fixed_normals"]), oracle.CloudRef(d["moving"], d["moving_normals"])
    o = oracle.icp_run(3, [oracle.make_slice(F, M, None, oracle.finder_params(0.3, 0.8),
                                             oracle.factor_params(oracle.FACTOR_PLANE, oracle.ROB_HUBER, 0.01))],
                       oracle.aligner_params(**kw), np.eye(4))
    for g, corr in runs:
        _assert_same_run(o, g, corr)


def test_supplied_correspondences_and_saturation(oracle, capi):
    """HBST path (multi_loop_detector_hbst_impl.cpp:331-352): externally matched pairs evaluated by
    srrg2b_linearize equal the oracle's sums; pairs whose residual leaves the fixed-point error range are
    suppressed and COUNTED (num_saturated), pairs naming a masked-out moving point are rejected."""
    rng = np.random.default_rng(8)
    n = 6000
    f = rng.uniform(-4, 4, size=(n, 3)).astype(np.float32)
    nf = rng.normal(size=(n, 3)); nf = (nf / np.linalg.norm(nf, axis=1, keepdims=True)).astype(np.float32)
    perm = rng.permutation(n)
    m = (f[perm] + rng.normal(scale=0.01, size=(n, 3))).astype(np.float32)
    nm = nf[perm].copy()
    m[:9] += np.float32(30.0)  # nine pairs far outside the residual bound
    mv = np.ones(n, np.uint8); mv[100] = 0
    sel = np.setdiff1d(rng.choice(n, size=4000, replace=False), [100])
    sel = np.union1d(sel, np.arange(9))
    fidx_dense = np.full(n, -1, np.int32)
    fidx_dense[sel] = perm[sel]
    F, M = oracle.CloudRef(f, nf), oracle.CloudRef(m, nm, mv)
    S = syn.iso3([0.004, -0.003, 0.002], [0.001, 0.002, -0.001]).astype(np.float32)
    ctx = capi.Context(3)
    ctx.set_cloud(capi.FIXED, 0, f, nf)
    ctx.set_cloud(capi.MOVING, 0, m, nm, mv)
    for factor in ("P2P", "PLANE"):
        ofp, gfp = oracle.finder_params(0.5, -2.0), capi.finder_params(0.5, -2.0)
        ofa = oracle.factor_params(getattr(oracle, "FACTOR_" + factor), oracle.ROB_CAUCHY, 0.01)
        gfa = capi.factor_params(getattr(capi, "FACTOR_" + factor), capi.ROB_CAUCHY, 0.01)
        o = oracle.linearize(F, M, fidx_dense, S, ofp, ofa)
        ctx.set_correspondences(0, perm[sel].astype(np.int32), sel.astype(np.int32))
        g = ctx.linearize(0, S, gfp, gfa, n_moving=n)
        assert np.array_equal(g["acc"], o["acc"])
        assert g["stats"] == o["stats"] and g["stats"]["num_saturated"] == 9 == g["stats"]["num_suppressed"]
        assert np.array_equal(g["H"], o["H"]) and np.array_equal(g["b"], o["b"])
    with pytest.raises(capi.Srrg2bError):
        ctx.set_correspondences(0, np.array([5], np.int32), np.array([100], np.int32))  # masked-out moving point
    ctx.close()


def test_icp_run_c2_full_size(oracle, capi):
    """Config C2 at BASELINE.json's full size (1M vs 1M points, 20 iterations, Huber, point+normal):
    the CUDA run equals the oracle's bit for bit -- pose, IterationStats, every correspondence -- and
    the size-independent properties of the correspondence list hold."""
    n = 1000000
    d = syn.make_icp3d(n, n, seed=2)
    kw = dict(max_iterations=20, min_num_inliers=10)
    o, g, corr = _run_both(oracle, capi, 3, d, oracle.aligner_params(**kw), capi.aligner_params(**kw),
                           oracle.finder_params(0.3, 0.8), capi.finder_params(0.3, 0.8),
                           oracle.factor_params(oracle.FACTOR_PLANE, oracle.ROB_HUBER, 0.01),
                           capi.factor_params(capi.FACTOR_PLANE, capi.ROB_HUBER, 0.01), np.eye(4))
    _assert_same_run(o, g, corr)
    assert g["status"] == capi.ALIGNER_SUCCESS and len(g["stats"]) == 20
    fi, mi, rs = corr
    assert np.all(np.diff(mi) > 0), "ascending moving index, at most one entry per moving point"
    assert fi.min() >= 0 and fi.max() < n and np.all(rs <= 0.3 + 1e-6) and np.all(rs >= 0)
    assert g["stats"][-1]["num_correspondences"] == fi.size
    rot, trans = syn.pose_error(g["T"], d["T_star"])
    assert rot < 1e-5 and trans < 1e-4
    # spot check of exactness against brute force: the reported neighbour is the nearest fixed point
    T = np.asarray(g["T"], dtype=np.float64)
    rng = np.random.default_rng(5)
    F = d["fixed"].astype(np.float64)
    for k in rng.choice(fi.size, size=24, replace=False):
        q = T[:3, :3] @ d["moving"][mi[k]].astype(np.float64) + T[:3, 3]
        d2 = ((F - q) ** 2).sum(axis=1)
        assert d2[fi[k]] <= d2.min() * (1 + 1e-5) + 1e-9


def test_fixed_cloud_replaced_between_runs(oracle, capi):
    """setFixed with a new cloud (the eager index build, the cached grid resolution) and a finder
    radius that changes between runs (lazy rebuild): every run equals the oracle's."""
    ctx = capi.Context(3)
    kw = dict(max_iterations=6, min_num_inliers=10)
    for nf, nm, seed, md in ((20000, 15000, 3, 0.3), (21000, 15000, 4, 0.3), (9000, 15000, 5, 0.3), (9000, 8000, 6, 0.45),
                             (30000, 8000, 7, 0.3)):
        d = syn.make_icp3d(nf, nm, seed=seed)
        ctx.set_cloud(capi.FIXED, 0, d["fixed"], d["fixed_normals"])
        ctx.set_cloud(capi.MOVING, 0, d["moving"], d["moving_normals"])
        g = ctx.icp_run([capi.make_slice(3, 0, None, capi.finder_params(md, 0.8),
                                         capi.factor_params(capi.FACTOR_PLANE, capi.ROB_HUBER, 0.01))],
                        capi.aligner_params(**kw), np.eye(4))
        gc = ctx.get_correspondences(0, nm)
        F = oracle.CloudRef(d["fixed"], d["fixed_normals"])
        M = oracle.CloudRef(d["moving"], d["moving_normals"])
        o = oracle.icp_run(3, [oracle.make_slice(F, M, None, oracle.finder_params(md, 0.8),
                                                 oracle.factor_params(oracle.FACTOR_PLANE, oracle.ROB_HUBER, 0.01))],
                           oracle.aligner_params(**kw), np.eye(4))
        assert g["stats"] == o["stats"] and np.array_equal(g["T"], o["T"])
        assert np.array_equal(gc[0], o["correspondences"][0][0]) and np.array_equal(gc[1], o["correspondences"][0][1])
    ctx.close()
