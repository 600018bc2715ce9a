"""MergerCorrespondenceHomo_::compute (R/mapping/merger_correspondence_homo_impl.cpp:11-126) on the
device-resident scene.  One tracker frame end to end on the device -- clip the scene, align the new scan against the
clip, merge the scan into the scene with the aligner's correspondences (flipped and mapped to the global scene through
the clip's indices, R/trackers/tracker_slice_processor_impl.cpp:159-191) -- against the same frame done by the oracle:
the merged scene must be the same bits (coordinates, normals, validity, size)."""
import numpy as np
import pytest

from srrg2_slam_interfaces_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _frame(oracle, capi, dim, scan, scan_n, scan_valid, scene, scene_n, scene_valid, T_clip, rng, T0, fp_kw, fa, merge_kw):
    kw = dict(max_iterations=8, min_num_inliers=10)
    # ---- device ----
    ctx = capi.Context(dim)
    ctx.scene_set(3, scene, scene_n, scene_valid)
    ctx.set_cloud(capi.FIXED, 0, scan, scan_n, scan_valid)
    n = ctx.scene_clip(3, 0, T_clip, rng)
    g = ctx.icp_run([capi.make_slice(dim, 0, None, capi.finder_params(*fp_kw), capi.factor_params(*fa(capi)))], capi.aligner_params(**kw), T0)
    # measurement (scan, robot frame) in scene = inverse of scene_in_robot refined by the aligner: scene_in_robot' = X * T_clip
    meas_in_scene = oracle.mat_op("orc_inverse", dim, oracle.mat_op("orc_mul", dim, g["T"], T_clip))
    nm, na = ctx.scene_merge(3, 0, meas_in_scene, **merge_kw)
    gs = ctx.scene_get(3, normals=True, valid=scene_valid is not None)
    # a second merge without correspondences (first frame of a new local map): every valid scan point is appended
    nm2, na2 = ctx.scene_merge(3, 0, meas_in_scene, without_correspondences=True)
    gs2 = ctx.scene_get(3, normals=True, valid=scene_valid is not None)
    ctx.close()
    # ---- oracle ----
    S = oracle.CloudRef(scene, scene_n, scene_valid)
    oc, on, og = oracle.scene_clip(S, T_clip, rng)
    assert n == og.shape[0]
    F, M = oracle.CloudRef(scan, scan_n, scan_valid), oracle.CloudRef(oc, on)
    o = oracle.icp_run(dim, [oracle.make_slice(F, M, None, oracle.finder_params(*fp_kw), oracle.factor_params(*fa(oracle)), dim=dim)],
                       oracle.aligner_params(**kw), T0)
    assert np.array_equal(g["T"], o["T"]) and g["stats"] == o["stats"]
    fi, mi, rs = o["correspondences"][0]
    corr = (og[mi], fi, rs)  # flipped, mapped to the global scene
    es, en, ev, onm, ona = oracle.scene_merge(scene, scene_n, scene_valid, F, meas_in_scene, corr, **merge_kw)
    assert (nm, na) == (onm, ona), ((nm, na), (onm, ona))
    assert np.array_equal(gs[0], es) and np.array_equal(gs[1], en)
    if scene_valid is not None:
        assert np.array_equal(gs[2], ev)
    es2, en2, ev2, onm2, ona2 = oracle.scene_merge(es, en, ev, F, meas_in_scene, None)
    assert (nm2, na2) == (onm2, ona2) and na2 == (int(scan_valid.sum()) if scan_valid is not None else scan.shape[0])
    assert np.array_equal(gs2[0], es2) and np.array_equal(gs2[1], en2)
    return nm, na


@pytest.mark.parametrize("target", [200, 10**9])
def test_merge_3d_frame(oracle, capi, target):
    d = syn.make_icp3d(20000, 60000, seed=31)
    scan_valid = (np.random.default_rng(2).uniform(size=20000) < 0.98).astype(np.uint8)
    scene_valid = (np.random.default_rng(3).uniform(size=60000) < 0.97).astype(np.uint8)
    T = syn.iso3([0.02, -0.01, 0.03], [0.002, -0.001, 0.003]).astype(np.float32)
    nm, na = _frame(oracle, capi, 3, d["fixed"], d["fixed_normals"], scan_valid, d["moving"], d["moving_normals"], scene_valid, T, 9.0,
                    np.eye(4), (0.3, 0.8), lambda m: (m.FACTOR_PLANE, m.ROB_HUBER, 0.01),
                    dict(maximum_response=0.2, maximum_distance_geometry_squared=0.02, target_number_of_merges=target))
    assert nm > 1000 and ((na == 0) if target == 200 else (na > 0))


def test_merge_2d_frame(oracle, capi):
    d = syn.make_multicue2d(150000, n_beams=1080, seed=5)
    sc = d["scans"][0]
    nm, na = _frame(oracle, capi, 2, sc["points"], sc["normals"], None, d["map"], d["map_normals"], None, sc["robot_in_sensor"].astype(np.float32),
                    12.0, np.eye(3), (0.5, 0.7), lambda m: (m.FACTOR_PLANE, m.ROB_CAUCHY, 0.05),
                    dict(maximum_response=50.0, maximum_distance_geometry_squared=0.25, target_number_of_merges=10**6))
    assert nm > 100
