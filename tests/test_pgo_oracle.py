"""CPU tests of the pose-graph oracle (oracle/pgo_oracle.py): Jacobians vs finite differences,
Gauss-Newton convergence on a noisy Manhattan-3D graph, gauge handling."""
import numpy as np

from srrg2_slam_interfaces_b200 import synthetic as syn


def _graph(n=300, f=1200, seed=4):
    return syn.make_pose_graph3d(n, f, seed=seed, box=(6, 6, 2))


def test_pose_pose_jacobians_against_finite_differences():
    from oracle import pgo_oracle as P
    g = _graph()
    poses = g["guess"].astype(np.float64)
    ij, Z, Om = g["ij"][:80], g["Z"][:80].astype(np.float64), g["Omega"][:80].astype(np.float64)
    e, Ji, Jj, chi = P.factor_terms(poses, ij, Z, Om)
    eps = 1e-6
    for which, J in (("i", Ji), ("j", Jj)):
        num = np.zeros_like(J)
        for k in range(6):
            d = np.zeros(6)
            d[k] = eps
            Xi, Xj = poses[ij[:, 0]].copy(), poses[ij[:, 1]].copy()
            if which == "i":
                Xi = Xi @ P.v2t(d)
            else:
                Xj = Xj @ P.v2t(d)
            num[:, :, k] = (P.t2v(P.inv_iso(Z) @ P.inv_iso(Xi) @ Xj) - e) / eps
        assert np.abs(num - J).max() < 2e-5


def test_gauss_newton_converges_to_the_noise_floor():
    from oracle import pgo_oracle as P
    g = _graph()
    sol, hist = P.solve(g["guess"], g["ij"], g["Z"], g["Omega"], g["fixed"], iterations=10)
    assert hist[-1]["dx_norm_inf"] < 1e-6
    assert hist[-1]["chi"] < hist[0]["chi"] * 1e-2
    dof = 6 * (g["ij"].shape[0] - (g["guess"].shape[0] - 1))
    assert 0.6 * dof < hist[-1]["chi"] < 1.5 * dof  # chi^2 of a consistent noise model
    assert np.array_equal(sol[0], g["guess"][0].astype(np.float64))  # the gauge does not move
    err = np.linalg.norm(sol[:, :3, 3] - g["truth"][:, :3, 3], axis=1)
    err0 = np.linalg.norm(g["guess"][:, :3, 3] - g["truth"][:, :3, 3], axis=1)
    assert err.max() < 0.2 and err.mean() < 0.25 * err0.mean()


def test_direct_and_cg_solvers_agree():
    from oracle import pgo_oracle as P
    g = _graph(150, 500, seed=9)
    a, sa = P.gn_step(g["guess"], g["ij"], g["Z"], g["Omega"], g["fixed"], solver="direct")
    b, sb = P.gn_step(g["guess"], g["ij"], g["Z"], g["Omega"], g["fixed"], solver="cg")
    assert abs(sa["chi"] - sb["chi"]) <= 1e-9 * sa["chi"]
    assert np.abs(a - b).max() < 1e-6


def test_two_level_preconditioner_reference():
    """The rigid-motion coarse space (reference for the CUDA solver's next preconditioner): same solution
    as the direct solve, and several times fewer PCG iterations than block-Jacobi."""
    from oracle import pgo_oracle as P
    import scipy.sparse.linalg as spla
    g = syn.make_pose_graph3d(2000, 9000, seed=4, box=(12, 12, 3))
    poses = g["guess"].astype(np.float64)
    H, b, chi, _ = P.linearize(poses, g["ij"], g["Z"].astype(np.float64), g["Omega"].astype(np.float64), g["fixed"])
    x_ref = spla.spsolve(H.tocsc(), -b)
    x_bj, it_bj, rel_bj = P.solve_block_jacobi(H, b, poses.shape[0], rtol=1e-11)
    x_tl, it_tl, rel_tl = P.solve_two_level(H, b, poses, g["fixed"], aggregate_size=16, rtol=1e-11)
    assert rel_bj < 1e-11 and rel_tl < 1e-11
    scale = np.abs(x_ref).max()
    assert np.abs(x_bj - x_ref).max() < 1e-6 * scale and np.abs(x_tl - x_ref).max() < 1e-6 * scale
    assert it_tl * 3 < it_bj, (it_tl, it_bj)
    # the coarse basis spans rigid motions: a global rigid motion of the free poses is reproduced exactly
    agg, Pi = P.rigid_prolongation(poses, np.zeros(poses.shape[0], bool), poses.shape[0])
    w, dt = np.array([0.01, -0.02, 0.015]), np.array([0.3, -0.1, 0.2])
    dx = np.einsum("vij,j->vi", Pi, np.concatenate([dt, w]))
    moved = poses @ P.v2t(dx)
    c = poses[:, :3, 3].mean(axis=0)
    Rw = P.R_from_quat(np.concatenate([w / 2, [np.sqrt(1 - (w / 2) @ (w / 2))]]))
    expect_t = (Rw @ (poses[:, :3, 3] - c).T).T + c + dt
    assert np.abs(moved[:, :3, 3] - expect_t).max() < 5e-3  # first order in |w|
