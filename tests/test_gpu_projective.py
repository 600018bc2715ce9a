"""Config C3 (RGB-D projective association, the srrg2_proslam cue): CUDA vs oracle, bit-exact."""
import numpy as np
import pytest

from srrg2_slam_interfaces_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _params(mod, K, max_distance=0.15, normal_cos=0.8):
    return mod.finder_params(max_distance, normal_cos, kind=mod.FINDER_PROJECTIVE, fx=K["fx"], fy=K["fy"], cx=K["cx"],
                             cy=K["cy"], width=K["width"], height=K["height"], min_depth=0.2, max_depth=15.0)


@pytest.mark.parametrize("width,height,n_frames", [(160, 120, 4), (640, 480, 30)])  # C3: the configured 30 frames
def test_projective_sequence(oracle, capi, width, height, n_frames):
    frames, K = syn.make_rgbd_sequence(n_frames, width, height, seed=3)
    ctx = capi.Context(3)
    ofp, gfp = _params(oracle, K), _params(capi, K)
    ofa = oracle.factor_params(oracle.FACTOR_PLANE, oracle.ROB_HUBER, 0.002, 1.0, 0.2)
    gfa = capi.factor_params(capi.FACTOR_PLANE, capi.ROB_HUBER, 0.002, 1.0, 0.2)
    kw = dict(max_iterations=10, min_num_inliers=100)
    for k in range(1, n_frames):
        fx, mv = frames[k], frames[k - 1]
        n = mv["points"].shape[0]
        F = oracle.CloudRef(fx["points"], fx["normals"], fx["valid"])
        M = oracle.CloudRef(mv["points"], mv["normals"], mv["valid"])
        ctx.set_cloud(capi.FIXED, 0, fx["points"], fx["normals"], fx["valid"])
        ctx.set_cloud(capi.MOVING, 0, mv["points"], mv["normals"], mv["valid"])
        # a3 alone, at identity and at the true motion
        T_true = syn.inv_iso(fx["pose"]) @ mv["pose"]
        for S in (np.eye(4), T_true):
            ofi, ors = oracle.find(None, F, M, S, ofp)
            fi, mi, rs = ctx.find_correspondences(0, S, gfp, n)
            dense = np.full(n, -1, np.int32)
            dense[mi] = fi
            resp = np.zeros(n, np.float32)
            resp[mi] = rs
            assert np.array_equal(dense, ofi) and np.array_equal(resp, ors)
            assert (ofi >= 0).sum() > 0.5 * fx["valid"].sum()
        # the whole aligner call
        o = oracle.icp_run(3, [oracle.make_slice(F, M, None, ofp, ofa)], oracle.aligner_params(**kw), np.eye(4))
        g = ctx.icp_run([capi.make_slice(3, 0, None, gfp, gfa)], capi.aligner_params(**kw), np.eye(4))
        assert g["status"] == o["status"] == capi.ALIGNER_SUCCESS
        assert g["stats"] == o["stats"] and np.array_equal(g["T"], o["T"])
        c = ctx.get_correspondences(0, n)
        oc = o["correspondences"][0]
        assert np.array_equal(c[0], oc[0]) and np.array_equal(c[1], oc[1]) and np.array_equal(c[2], oc[2])
        rot, trans = syn.pose_error(g["T"], T_true)
        assert rot < 3e-3 and trans < 1e-2, (rot, trans)
    ctx.close()
