"""the local map resident on the device and clipped there (srrg2b_scene_set / _clip /
_clip_indices) against the oracle's range clipper: same kept indices, and -- the clipped cloud never leaves the
device -- the aligner run on it equals the oracle's run on the oracle-clipped cloud bit for bit (pose,
IterationStats, correspondences), which it can only do if the clipped points and normals are the same bits."""
import numpy as np
import pytest

from srrg2_slam_interfaces_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _run_pair(oracle, capi, dim, fixed, fixed_n, scene, scene_n, scene_valid, T, rng, T0, kw, fp_kw, fa):
    S = oracle.CloudRef(scene, scene_n, scene_valid)
    oc, on, og = oracle.scene_clip(S, T, rng)
    ctx = capi.Context(dim)
    ctx.scene_set(7, scene, scene_n, scene_valid)
    ctx.set_cloud(capi.FIXED, 0, fixed, fixed_n)
    n = ctx.scene_clip(7, 0, T, rng)
    assert n == og.shape[0]
    assert np.array_equal(ctx.scene_clip_indices(0), og)
    g = ctx.icp_run([capi.make_slice(dim, 0, None, capi.finder_params(*fp_kw), capi.factor_params(*fa(capi)))], capi.aligner_params(**kw), T0)
    gc = ctx.get_correspondences(0, max(n, 1))
    # clip again at another pose / range on the same resident scene: the slice's moving cloud is replaced
    T2 = T.copy(); T2[0, dim] += 0.5
    oc2, on2, og2 = oracle.scene_clip(S, T2, 0.7 * rng)
    assert ctx.scene_clip(7, 0, T2, 0.7 * rng) == og2.shape[0]
    assert np.array_equal(ctx.scene_clip_indices(0), og2)
    ctx.close()
    if n == 0:
        assert g["status"] != capi.ALIGNER_SUCCESS
        return
    F, M = oracle.CloudRef(fixed, fixed_n), oracle.CloudRef(oc, on)
    o = oracle.icp_run(dim, [oracle.make_slice(F, M, None, oracle.finder_params(*fp_kw), oracle.factor_params(*fa(oracle)), dim=dim)],
                       oracle.aligner_params(**kw), T0)
    assert g["status"] == o["status"] and g["stats"] == o["stats"] and np.array_equal(g["T"], o["T"])
    for a, b in zip(gc, o["correspondences"][0]):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("rng", [6.0, 1e6, 1e-3])
def test_clip_3d_then_align(oracle, capi, rng):
    d = syn.make_icp3d(30000, 60000, seed=23)
    valid = (np.random.default_rng(5).uniform(size=60000) < 0.97).astype(np.uint8)
    T = syn.iso3([0.3, -0.2, 0.1], [0.01, -0.02, 0.015]).astype(np.float32)
    # the fixed cloud lives in the robot frame the clip maps into: express it there too
    Tf = T.astype(np.float64)
    fixed = (d["fixed"].astype(np.float64) @ Tf[:3, :3].T + Tf[:3, 3]).astype(np.float32)
    fixed_n = (d["fixed_normals"].astype(np.float64) @ Tf[:3, :3].T).astype(np.float32)
    _run_pair(oracle, capi, 3, fixed, fixed_n, d["moving"], d["moving_normals"], valid, T, rng, np.eye(4),
              dict(max_iterations=8, min_num_inliers=10), (0.3, 0.8), lambda m: (m.FACTOR_PLANE, m.ROB_HUBER, 0.01))


def test_clip_2d_then_align(oracle, capi):
    d = syn.make_multicue2d(200000, n_beams=1080, seed=5)
    sc = d["scans"][0]
    T = syn.iso2(0.05, -0.02, 0.01).astype(np.float32)
    _run_pair(oracle, capi, 2, sc["points"], sc["normals"], d["map"], d["map_normals"], None, T, 9.0, np.eye(3),
              dict(max_iterations=6, min_num_inliers=10), (0.5, 0.7), lambda m: (m.FACTOR_PLANE, m.ROB_CAUCHY, 0.05))
