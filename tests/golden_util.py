import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["icp3d_plane_huber", "icp3d_p2p_euler_cauchy_prune", "icp2d_plane_c1"]
AP_KEYS = ("variable", "max_iterations", "min_num_inliers", "enable_inlier_only_runs",
           "keep_only_inlier_correspondences", "use_termination_criteria")
INT_KEYS = ["iteration", "solver_status", "num_inliers", "num_outliers", "num_suppressed", "num_correspondences"]


def load(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    g["dim"] = int(g["dim"])
    g["fixed_valid"] = g["fixed_valid"] if g["fixed_valid"].size else None
    g["moving_valid"] = g["moving_valid"] if g["moving_valid"].size else None
    g["ap_kw"] = {k: int(v) for k, v in zip(AP_KEYS, g["ap"])}
    g["fp_kw"] = dict(max_distance=float(g["fp"][0]), normal_cos=float(g["fp"][1]))
    g["fa_kw"] = dict(factor=int(g["fa"][0]), robustifier=int(g["fa"][1]), chi_threshold=float(g["fa_f"][0]),
                      info_point=float(g["fa_f"][1]), info_normal=float(g["fa_f"][2]))
    return g


def check_run(g, T, status, stats, corr):
    assert status == int(g["status"])
    assert np.array_equal(np.asarray(T, np.float32), g["T"])
    assert len(stats) == g["stats_int"].shape[0]
    for s, ints, chis in zip(stats, g["stats_int"], g["stats_chi"]):
        assert [s[k] for k in INT_KEYS] == ints.tolist()
        assert s["chi_inliers"] == chis[0] and s["chi_outliers"] == chis[1]
    assert np.array_equal(corr[0], g["corr_fixed"])
    assert np.array_equal(corr[1], g["corr_moving"])
    assert np.array_equal(corr[2], g["corr_resp"])
