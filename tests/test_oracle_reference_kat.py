"""Known-answer tests the reference itself holds for the aligner path, restated for the oracle.

The only reference test that reaches Solver + MultiAligner is tests/test_motion_model_slice.cpp
(reference tree): an aligner with ONE prior slice (AlignerSliceMotionModel3D) and min_num_inliers=0
must return movingInFixed() == inverse(previous motion): |t2v(T * motion_previous)| < 1e-5
(:81-85, :139-142; 1e-4 at :220-223).  The slice uses the inverse motion both as initial guess and
as the prior measurement (aligner_slice_motion_model.hpp:69-70,78)."""
import numpy as np

from srrg2_slam_interfaces_b200 import synthetic as syn


def _t2v_norm(oracle, T):
    import ctypes as C
    v = np.zeros(6, dtype=np.float32)
    T = np.ascontiguousarray(T, dtype=np.float32)
    f = oracle.lib().orc_t2v
    f.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
    f(3, T.ctypes.data, v.ctypes.data)
    return float(np.linalg.norm(v))


def _random_motion(rng):
    return syn.iso3(rng.uniform(-1, 1, size=3), rng.uniform(0, 1, size=3) * np.pi * rng.integers(1, 1000, size=3))


def test_motion_model_prior_slice_3d(oracle):
    rng = np.random.default_rng(0)
    for i in range(10):
        motion_previous = _random_motion(rng).astype(np.float32)
        motion_inverse = syn.inv_iso(motion_previous.astype(np.float64)).astype(np.float32)
        sl = oracle.make_slice(prior_measurement=motion_inverse, prior_info_diag=np.ones(6), dim=3)
        ap = oracle.aligner_params(max_iterations=10, min_num_inliers=0)
        r = oracle.icp_run(3, [sl], ap, np.eye(4))
        assert r["status"] == 0  # AlignerBase::Success
        assert len(r["stats"]) == 10 and all(s["solver_status"] == 1 for s in r["stats"])
        assert r["stats"][-1]["num_inliers"] == 1  # a prior counts as one factor
        err = _t2v_norm(oracle, r["T"].astype(np.float64) @ motion_previous.astype(np.float64))
        assert err < 1e-5, err


def test_odometry_prior_pulls_towards_measurement_2d(oracle):
    """AlignerSliceOdom2DPrior: guess := measurement (aligner_slice_odometry_prior.cpp:19), default
    information 1e2 (aligner_slice_odometry_prior.h:17-21); a GN step from the measurement stays put."""
    Z = syn.iso2(0.3, -0.2, 0.4).astype(np.float32)
    sl = oracle.make_slice(prior_measurement=Z, prior_info_diag=np.full(3, 100.0), dim=2)
    r = oracle.icp_run(2, [sl], oracle.aligner_params(max_iterations=3, min_num_inliers=0), np.eye(3))
    assert r["status"] == 0
    rot, trans = syn.pose_error(r["T"], Z)
    assert rot < 1e-6 and trans < 1e-6
    assert abs(r["stats"][-1]["chi_inliers"]) < 1e-10
