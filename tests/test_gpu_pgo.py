"""Pose-graph Gauss-Newton (row a10, config C4 at reduced size): CUDA (fp64 linearise + PCG) vs the
numpy/scipy oracle.  Tolerances from BASELINE.json north_star: pose 1e-5 rad / 1e-4 m, chi^2 1e-6 rel."""
import numpy as np
import pytest

from srrg2_slam_interfaces_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _rot_err(A, B):
    E = np.swapaxes(A[:, :3, :3], 1, 2) @ B[:, :3, :3]
    w = np.stack([E[:, 2, 1] - E[:, 1, 2], E[:, 0, 2] - E[:, 2, 0], E[:, 1, 0] - E[:, 0, 1]], axis=1) / 2
    return np.linalg.norm(w, axis=1)


@pytest.mark.parametrize("n,f,box", [(400, 1500, (6, 6, 2)), (1500, 7000, (10, 10, 3))])
def test_gn_iterations_match_oracle(capi, n, f, box):
    from oracle import pgo_oracle as P
    g = syn.make_pose_graph3d(n, f, seed=4, box=box)
    ctx = capi.Context(3)
    ctx.pgo_upload(g["guess"], g["fixed"], g["ij"], g["Z"], g["Omega"])
    poses = g["guess"].astype(np.float64)
    for it in range(5):
        poses, so = P.gn_step(poses, g["ij"], g["Z"].astype(np.float64), g["Omega"].astype(np.float64), g["fixed"])
        sg = ctx.pgo_iterate(max_cg_iterations=4000, cg_tolerance=1e-11)
        assert sg["cg_relative_residual"] <= 1e-10
        assert abs(sg["chi"] - so["chi"]) <= 1e-6 * so["chi"], (it, sg["chi"], so["chi"])
        got = ctx.pgo_download().astype(np.float64)
        assert np.abs(got[:, :3, 3] - poses[:, :3, 3]).max() < 1e-4
        assert _rot_err(got, poses).max() < 1e-5
        assert abs(sg["dx_norm_inf"] - so["dx_norm_inf"]) < 1e-6 * max(1.0, so["dx_norm_inf"])
    assert so["dx_norm_inf"] < 1e-3
    assert np.array_equal(got[0], g["guess"][0].astype(np.float64))
    ctx.close()
