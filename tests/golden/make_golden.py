#!/usr/bin/env python
"""Generates the committed golden vectors (tests/golden/*.npz).

The reference cannot be built or imported here (its arithmetic lives in the un-vendored
srrg2_core / srrg2_solver packages), so these vectors are produced by the CPU oracle
(oracle/srrg2b_oracle.c) and pin ITS behaviour: later oracle edits and the CUDA path are both held
to them.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from srrg2_slam_interfaces_b200 import synthetic as syn  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def stats_array(stats):
    keys = ["iteration", "solver_status", "num_inliers", "num_outliers", "num_suppressed", "num_correspondences"]
    ints = np.array([[s[k] for k in keys] for s in stats], dtype=np.int64).reshape(-1, len(keys))
    chis = np.array([[s["chi_inliers"], s["chi_outliers"]] for s in stats], dtype=np.float64).reshape(-1, 2)
    return ints, chis


def case(name, dim, d, fp_kw, fa_kw, ap_kw, T0, valid_seed=None):
    fv = mv = None
    if valid_seed is not None:
        rng = np.random.default_rng(valid_seed)
        fv = (rng.uniform(size=d["fixed"].shape[0]) < 0.93).astype(np.uint8)
        mv = (rng.uniform(size=d["moving"].shape[0]) < 0.93).astype(np.uint8)
    F = O.CloudRef(d["fixed"], d["fixed_normals"], fv)
    M = O.CloudRef(d["moving"], d["moving_normals"], mv)
    fp, fa, ap = O.finder_params(**fp_kw), O.factor_params(**fa_kw), O.aligner_params(**ap_kw)
    r = O.icp_run(dim, [O.make_slice(F, M, None, fp, fa, dim=dim)], ap, T0)
    ints, chis = stats_array(r["stats"])
    fi, mi, rs = r["correspondences"][0]
    # first-iteration pieces: correspondences at T0 and their linearisation
    ix = O.Index(F)
    fidx0, resp0 = O.find(ix, F, M, T0, fp)
    lin0 = O.linearize(F, M, fidx0, T0, fp, fa, variable=ap.variable)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), dim=dim, fixed=d["fixed"], fixed_normals=d["fixed_normals"],
        moving=d["moving"], moving_normals=d["moving_normals"],
        fixed_valid=fv if fv is not None else np.zeros(0, np.uint8),
        moving_valid=mv if mv is not None else np.zeros(0, np.uint8),
        T0=np.asarray(T0, np.float32), fp=np.array([fp_kw["max_distance"], fp_kw["normal_cos"]], np.float32),
        fa=np.array([fa_kw["factor"], fa_kw["robustifier"]], np.int32),
        fa_f=np.array([fa_kw["chi_threshold"], fa_kw["info_point"], fa_kw["info_normal"]], np.float32),
        ap=np.array([ap_kw[k] for k in ("variable", "max_iterations", "min_num_inliers", "enable_inlier_only_runs",
                                        "keep_only_inlier_correspondences", "use_termination_criteria")], np.int32),
        T=r["T"], status=r["status"], stats_int=ints, stats_chi=chis, corr_fixed=fi, corr_moving=mi, corr_resp=rs,
        find0_fixed=fidx0, find0_resp=resp0, lin0_acc=lin0["acc"], lin0_H=lin0["H"], lin0_b=lin0["b"])
    print(name, "status", r["status"], "iters", len(r["stats"]), "corr", fi.size)


def main():
    base_ap = dict(variable=0, max_iterations=8, min_num_inliers=10, enable_inlier_only_runs=0,
                   keep_only_inlier_correspondences=0, use_termination_criteria=0)
    case("icp3d_plane_huber", 3, syn.make_icp3d(4000, 3500, seed=31, cube=8.0, n_planes=12, n_spheres=3),
         dict(max_distance=0.4, normal_cos=0.8), dict(factor=1, robustifier=4, chi_threshold=0.02, info_point=1.0, info_normal=0.5),
         base_ap, np.eye(4), valid_seed=5)
    case("icp3d_p2p_euler_cauchy_prune", 3, syn.make_icp3d(3000, 3000, seed=32, cube=6.0, n_planes=10, n_spheres=2),
         dict(max_distance=0.5, normal_cos=-2.0), dict(factor=0, robustifier=2, chi_threshold=0.05, info_point=2.0, info_normal=1.0),
         dict(base_ap, variable=1, enable_inlier_only_runs=1, keep_only_inlier_correspondences=1), np.eye(4))
    case("icp2d_plane_c1", 2, syn.make_icp2d(2500, seed=1),
         dict(max_distance=0.5, normal_cos=0.8), dict(factor=1, robustifier=0, chi_threshold=1.0, info_point=1.0, info_normal=1.0),
         dict(base_ap, max_iterations=6), np.eye(3))


if __name__ == "__main__":
    main()
