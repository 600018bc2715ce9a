"""CPU ORACLE (test infrastructure) for the pose-graph Gauss-Newton step -- SURVEY.md section 8 row a10.

Restates what MultiGraphSLAM_::optimize() reaches (R/system/multi_graph_slam_impl.cpp:299-317:
graph->bindFactors(); solver->setGraph(); solver->compute()) for the factor / variable types the
reference instantiates: SE3PosePoseGeodesicErrorFactor (R/registration/loop_closure.h:110-111) between
VariableSE3QuaternionRightAD local maps (R/mapping/local_map.h:64,75), first map fixed as the gauge
(multi_graph_slam_impl.cpp:85-87).  The factor / solver bodies live in srrg2_solver (absent, unpinned):
PARITY UNPINNED; the arithmetic is restated from its conventions:

    e_ij   = t2v( Z_ij^-1 * X_i^-1 * X_j )          t2v = [t ; unit-quaternion vector part, w >= 0]
    X      <- X * v2t(dx)                            right perturbation, dx = [dt ; dq]
    H dx   = -b,  H = sum J^T Omega J,  b = sum J^T Omega e

numpy / scipy, float64, direct sparse solve (scipy.sparse.linalg.spsolve).  Only tests/ and the
cpu_baseline legs of the benchmarks import this module.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def skew(v):
    z = np.zeros(v.shape[:-1])
    return np.stack([np.stack([z, -v[..., 2], v[..., 1]], -1),
                     np.stack([v[..., 2], z, -v[..., 0]], -1),
                     np.stack([-v[..., 1], v[..., 0], z], -1)], -2)


def inv_iso(T):
    Ti = np.zeros_like(T)
    Rt = np.swapaxes(T[..., :3, :3], -1, -2)
    Ti[..., :3, :3] = Rt
    Ti[..., :3, 3] = -np.einsum("...ij,...j->...i", Rt, T[..., :3, 3])
    Ti[..., 3, 3] = 1.0
    return Ti


def quat_from_R(R):
    """(x, y, z, w), unit, w >= 0.  Vectorised four-branch extraction (same branches as the aligner oracle)."""
    R = np.asarray(R, dtype=np.float64)
    tr = R[..., 0, 0] + R[..., 1, 1] + R[..., 2, 2]
    q = np.zeros(R.shape[:-2] + (4,))
    c0 = tr > 0
    c1 = ~c0 & (R[..., 0, 0] > R[..., 1, 1]) & (R[..., 0, 0] > R[..., 2, 2])
    c2 = ~c0 & ~c1 & (R[..., 1, 1] > R[..., 2, 2])
    c3 = ~c0 & ~c1 & ~c2
    with np.errstate(invalid="ignore", divide="ignore"):
        s = np.sqrt(np.maximum(tr + 1.0, 1e-300)) * 2.0
        q0 = np.stack([(R[..., 2, 1] - R[..., 1, 2]) / s, (R[..., 0, 2] - R[..., 2, 0]) / s,
                       (R[..., 1, 0] - R[..., 0, 1]) / s, 0.25 * s], -1)
        s = np.sqrt(np.maximum(1.0 + R[..., 0, 0] - R[..., 1, 1] - R[..., 2, 2], 1e-300)) * 2.0
        q1 = np.stack([0.25 * s, (R[..., 0, 1] + R[..., 1, 0]) / s, (R[..., 0, 2] + R[..., 2, 0]) / s,
                       (R[..., 2, 1] - R[..., 1, 2]) / s], -1)
        s = np.sqrt(np.maximum(1.0 + R[..., 1, 1] - R[..., 0, 0] - R[..., 2, 2], 1e-300)) * 2.0
        q2 = np.stack([(R[..., 0, 1] + R[..., 1, 0]) / s, 0.25 * s, (R[..., 1, 2] + R[..., 2, 1]) / s,
                       (R[..., 0, 2] - R[..., 2, 0]) / s], -1)
        s = np.sqrt(np.maximum(1.0 + R[..., 2, 2] - R[..., 0, 0] - R[..., 1, 1], 1e-300)) * 2.0
        q3 = np.stack([(R[..., 0, 2] + R[..., 2, 0]) / s, (R[..., 1, 2] + R[..., 2, 1]) / s, 0.25 * s,
                       (R[..., 1, 0] - R[..., 0, 1]) / s], -1)
    for c, qq in ((c0, q0), (c1, q1), (c2, q2), (c3, q3)):
        q[c] = qq[c]
    q /= np.linalg.norm(q, axis=-1, keepdims=True)
    q[q[..., 3] < 0] *= -1.0
    return q


def quat_mul(a, b):
    """Hamilton product, (x, y, z, w) layout."""
    av, aw, bv, bw = a[..., :3], a[..., 3:], b[..., :3], b[..., 3:]
    v = aw * bv + bw * av + np.cross(av, bv)
    w = aw * bw - np.sum(av * bv, -1, keepdims=True)
    return np.concatenate([v, w], -1)


def R_from_quat(q):
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - 2 * (y * y + z * z); R[..., 0, 1] = 2 * (x * y - w * z); R[..., 0, 2] = 2 * (x * z + w * y)
    R[..., 1, 0] = 2 * (x * y + w * z); R[..., 1, 1] = 1 - 2 * (x * x + z * z); R[..., 1, 2] = 2 * (y * z - w * x)
    R[..., 2, 0] = 2 * (x * z - w * y); R[..., 2, 1] = 2 * (y * z + w * x); R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def v2t(dx):
    """[dt ; dq] -> isometry, quaternion (w = sqrt(1 - |dq|^2), dq)."""
    dx = np.asarray(dx, dtype=np.float64)
    dq = dx[..., 3:6]
    n2 = np.sum(dq * dq, -1, keepdims=True)
    w = np.sqrt(np.maximum(1.0 - n2, 0.0))
    big = n2[..., 0] >= 1.0
    q = np.concatenate([dq, w], -1)
    if np.any(big):
        q[big, :3] = dq[big] / np.sqrt(n2[big])
        q[big, 3] = 0.0
    T = np.zeros(dx.shape[:-1] + (4, 4))
    T[..., :3, :3] = R_from_quat(q)
    T[..., :3, 3] = dx[..., :3]
    T[..., 3, 3] = 1.0
    return T


def t2v(T):
    return np.concatenate([T[..., :3, 3], quat_from_R(T[..., :3, :3])[..., :3]], -1)


def factor_terms(poses, ij, Z, Omega):
    """Per factor: e (F,6), J_i (F,6,6), J_j (F,6,6), chi (F,)."""
    Xi, Xj = poses[ij[:, 0]], poses[ij[:, 1]]
    A = inv_iso(Xi) @ Xj
    Zi = inv_iso(Z)
    E = Zi @ A
    qe = quat_from_R(E[:, :3, :3])
    e = np.concatenate([E[:, :3, 3], qe[:, :3]], -1)
    F = ij.shape[0]
    Jj = np.zeros((F, 6, 6))
    Jj[:, :3, :3] = E[:, :3, :3]
    Jj[:, 3:, 3:] = qe[:, 3, None, None] * np.eye(3) + skew(qe[:, :3])
    Ji = np.zeros((F, 6, 6))
    Rzi = Zi[:, :3, :3]
    Ji[:, :3, :3] = -Rzi
    Ji[:, :3, 3:] = 2.0 * Rzi @ skew(A[:, :3, 3])
    qz, qa = quat_from_R(Rzi), quat_from_R(A[:, :3, :3])
    prod = quat_mul(qz, qa)
    sgn = np.where(np.sum(prod * qe, -1) < 0, -1.0, 1.0)  # q_z^-1 (x) q_a = +- q_e
    for k in range(3):
        ek = np.zeros((F, 4))
        ek[:, k] = 1.0
        Ji[:, 3:, 3 + k] = -sgn[:, None] * quat_mul(quat_mul(qz, ek), qa)[:, :3]
    chi = np.einsum("fi,fij,fj->f", e, Omega, e)
    return e, Ji, Jj, chi


def linearize(poses, ij, Z, Omega, fixed):
    """Sparse H (6V x 6V, CSR) and b (6V); rows/cols of fixed variables are replaced by identity."""
    poses = np.asarray(poses, np.float64); Z = np.asarray(Z, np.float64); Omega = np.asarray(Omega, np.float64)
    V = poses.shape[0]
    e, Ji, Jj, chi = factor_terms(poses, ij, Z, Omega)
    JiT_O = np.swapaxes(Ji, 1, 2) @ Omega
    JjT_O = np.swapaxes(Jj, 1, 2) @ Omega
    blocks = [(ij[:, 0], ij[:, 0], JiT_O @ Ji), (ij[:, 0], ij[:, 1], JiT_O @ Jj),
              (ij[:, 1], ij[:, 0], JjT_O @ Ji), (ij[:, 1], ij[:, 1], JjT_O @ Jj)]
    free = ~np.asarray(fixed, bool)
    rows, cols, vals = [], [], []
    r6, c6 = np.meshgrid(np.arange(6), np.arange(6), indexing="ij")
    for bi, bj, blk in blocks:
        keep = free[bi] & free[bj]
        rows.append((bi[keep, None, None] * 6 + r6).ravel())
        cols.append((bj[keep, None, None] * 6 + c6).ravel())
        vals.append(blk[keep].ravel())
    fx = np.nonzero(~free)[0]
    rows.append((fx[:, None] * 6 + np.arange(6)).ravel()); cols.append((fx[:, None] * 6 + np.arange(6)).ravel())
    vals.append(np.ones(fx.size * 6))
    H = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(6 * V, 6 * V)).tocsr()
    b = np.zeros((V, 6))
    np.add.at(b, ij[:, 0], np.einsum("fij,fj->fi", JiT_O, e))
    np.add.at(b, ij[:, 1], np.einsum("fij,fj->fi", JjT_O, e))
    b[~free] = 0.0
    return H, b.ravel(), float(chi.sum()), chi


# ---- two-level preconditioner (reference for the CUDA solver's next step) ----
def rigid_prolongation(poses, fixed, aggregate_size):
    """Coarse space of the two-level preconditioner: the poses are grouped into aggregates of
    `aggregate_size` consecutive indices, each aggregate gets the 6 rigid motions (dt, w) of its
    members about the aggregate's centroid c, expressed as the members' RIGHT perturbations:
        dt_i = R_i^T (dt + w x (t_i - c)),   dq_i = R_i^T w / 2.
    Returns (agg[V], Pi[V, 6, 6]); fixed poses get a zero block."""
    poses = np.asarray(poses, np.float64)
    V = poses.shape[0]
    agg = np.arange(V) // int(aggregate_size)
    R, t = poses[:, :3, :3], poses[:, :3, 3]
    c = np.zeros((agg.max() + 1, 3))
    np.add.at(c, agg, t)
    c /= np.bincount(agg)[:, None]
    Rt = np.swapaxes(R, 1, 2)
    Pi = np.zeros((V, 6, 6))
    Pi[:, :3, :3] = Rt
    Pi[:, :3, 3:] = -Rt @ skew(t - c[agg])
    Pi[:, 3:, 3:] = 0.5 * Rt
    Pi[np.asarray(fixed, bool)] = 0.0
    return agg, Pi


def coarse_matrix(H, agg, Pi):
    """Galerkin coarse operator Hc = P^T H P (dense, 6 * n_aggregates square) and P as a sparse matrix."""
    V = agg.shape[0]
    nc = int(agg.max()) + 1
    rows = (np.arange(V)[:, None, None] * 6 + np.arange(6)[None, :, None] + 0 * np.arange(6)[None, None, :]).ravel()
    cols = (agg[:, None, None] * 6 + 0 * np.arange(6)[None, :, None] + np.arange(6)[None, None, :]).ravel()
    Pm = sp.csr_matrix((Pi.ravel(), (rows, cols)), shape=(6 * V, 6 * nc))
    Hc = np.asarray((Pm.T @ H @ Pm).todense())
    # aggregates made of fixed poses only have a zero block: give them a unit diagonal
    dead = np.abs(Hc).sum(axis=1) == 0.0
    Hc[dead, dead] = 1.0
    return Hc, Pm


def pcg(H, b, apply_M, rtol=1e-10, maxiter=20000):
    """Preconditioned conjugate gradients for H x = -b; returns (x, iterations, relative residual)."""
    n = b.shape[0]
    x = np.zeros(n)
    r = -b.copy()
    z = apply_M(r)
    p = z.copy()
    rz = r @ z
    b2 = b @ b
    rel = 1.0
    for it in range(1, maxiter + 1):
        Ap = H @ p
        a = rz / (p @ Ap)
        x += a * p
        r -= a * Ap
        rel = np.sqrt((r @ r) / b2) if b2 > 0 else 0.0
        if rel < rtol:
            return x, it, rel
        z = apply_M(r)
        rz2 = r @ z
        p = z + (rz2 / rz) * p
        rz = rz2
    return x, maxiter, rel


def solve_two_level(H, b, poses, fixed, aggregate_size=32, rtol=1e-10, maxiter=20000):
    """PCG with M^-1 = blockdiag(H)^-1 + P (P^T H P)^-1 P^T (additive two-level Schwarz)."""
    import scipy.linalg as sla
    V = np.asarray(poses).shape[0]
    H = H.tocsr()
    D = np.stack([H[6 * v:6 * v + 6, 6 * v:6 * v + 6].toarray() for v in range(V)])
    Dinv = np.linalg.inv(D)
    agg, Pi = rigid_prolongation(poses, fixed, aggregate_size)
    Hc, Pm = coarse_matrix(H, agg, Pi)
    cho = sla.cho_factor(Hc + 1e-12 * np.trace(Hc) / Hc.shape[0] * np.eye(Hc.shape[0]))

    def apply_M(r):
        return np.einsum("vij,vj->vi", Dinv, r.reshape(V, 6)).ravel() + Pm @ sla.cho_solve(cho, Pm.T @ r)

    return pcg(H, b, apply_M, rtol, maxiter)


def solve_block_jacobi(H, b, n_vars, rtol=1e-10, maxiter=20000):
    H = H.tocsr()
    D = np.stack([H[6 * v:6 * v + 6, 6 * v:6 * v + 6].toarray() for v in range(n_vars)])
    Dinv = np.linalg.inv(D)
    return pcg(H, b, lambda r: np.einsum("vij,vj->vi", Dinv, r.reshape(n_vars, 6)).ravel(), rtol, maxiter)


def gn_step(poses, ij, Z, Omega, fixed, solver="direct"):
    """One Gauss-Newton iteration; returns (new poses, stats)."""
    H, b, chi, _ = linearize(poses, ij, Z, Omega, fixed)
    if solver == "direct":
        dx = spla.spsolve(H.tocsc(), -b)
    elif solver == "two_level":
        dx, _, _ = solve_two_level(H, b, poses, fixed)
    else:
        Minv = spla.LinearOperator(H.shape, matvec=lambda x: x / H.diagonal())
        dx, info = spla.cg(H, -b, rtol=1e-12, maxiter=20000, M=Minv)
    dx = dx.reshape(-1, 6)
    new = np.asarray(poses, np.float64) @ v2t(dx)
    return new, dict(chi=chi, dx_norm_inf=float(np.abs(dx).max()), num_factors=int(ij.shape[0]))


def solve(poses, ij, Z, Omega, fixed, iterations=10, tol=1e-6):
    poses = np.asarray(poses, np.float64)
    hist = []
    for _ in range(iterations):
        poses, st = gn_step(poses, ij, Z, Omega, fixed)
        hist.append(st)
        if st["dx_norm_inf"] < tol:
            break
    return poses, hist


# ----------------------------------------------------------------------------------------------
# SE(2): SE2PosePoseGeodesicErrorFactor between VariableSE2Right poses (R/registration/loop_closure.h:110,
# R/mapping/local_map.h:64).  Poses are 3x3 homogeneous matrices; e = t2v(Z^-1 Xi^-1 Xj) = (t_E, angle(R_E)),
# right perturbation X <- X v2t([dx, dy, dtheta]).
# ----------------------------------------------------------------------------------------------
def v2t2(dx):
    dx = np.asarray(dx, dtype=np.float64)
    c, s = np.cos(dx[..., 2]), np.sin(dx[..., 2])
    T = np.zeros(dx.shape[:-1] + (3, 3))
    T[..., 0, 0] = c; T[..., 0, 1] = -s; T[..., 1, 0] = s; T[..., 1, 1] = c
    T[..., 0, 2] = dx[..., 0]; T[..., 1, 2] = dx[..., 1]; T[..., 2, 2] = 1.0
    return T


def t2v2(T):
    return np.stack([T[..., 0, 2], T[..., 1, 2], np.arctan2(T[..., 1, 0], T[..., 0, 0])], -1)


def inv_iso2(T):
    R = np.swapaxes(T[..., :2, :2], -1, -2)
    o = np.zeros_like(T)
    o[..., :2, :2] = R
    o[..., :2, 2] = -np.einsum("...ij,...j->...i", R, T[..., :2, 2])
    o[..., 2, 2] = 1.0
    return o


def factor_terms2(poses, ij, Z, Omega):
    """Per factor: e (F,3), J_i (F,3,3), J_j (F,3,3), chi (F,)."""
    Xi, Xj = poses[ij[:, 0]], poses[ij[:, 1]]
    A = inv_iso2(Xi) @ Xj
    Zi = inv_iso2(Z)
    E = Zi @ A
    e = t2v2(E)
    F = ij.shape[0]
    Jj = np.zeros((F, 3, 3))
    Jj[:, :2, :2] = E[:, :2, :2]
    Jj[:, 2, 2] = 1.0
    Ji = np.zeros((F, 3, 3))
    Rzi = Zi[:, :2, :2]
    Ji[:, :2, :2] = -Rzi
    St = np.stack([-A[:, 1, 2], A[:, 0, 2]], -1)  # S t_A, S = [0 -1; 1 0]
    Ji[:, :2, 2] = -np.einsum("fij,fj->fi", Rzi, St)
    Ji[:, 2, 2] = -1.0
    chi = np.einsum("fi,fij,fj->f", e, Omega, e)
    return e, Ji, Jj, chi


def linearize2(poses, ij, Z, Omega, fixed):
    poses = np.asarray(poses, np.float64); Z = np.asarray(Z, np.float64); Omega = np.asarray(Omega, np.float64)
    V = poses.shape[0]
    e, Ji, Jj, chi = factor_terms2(poses, ij, Z, Omega)
    JiT_O = np.swapaxes(Ji, 1, 2) @ Omega
    JjT_O = np.swapaxes(Jj, 1, 2) @ Omega
    blocks = [(ij[:, 0], ij[:, 0], JiT_O @ Ji), (ij[:, 0], ij[:, 1], JiT_O @ Jj),
              (ij[:, 1], ij[:, 0], JjT_O @ Ji), (ij[:, 1], ij[:, 1], JjT_O @ Jj)]
    free = ~np.asarray(fixed, bool)
    rows, cols, vals = [], [], []
    r3, c3 = np.meshgrid(np.arange(3), np.arange(3), indexing="ij")
    for bi, bj, blk in blocks:
        keep = free[bi] & free[bj]
        rows.append((bi[keep, None, None] * 3 + r3).ravel())
        cols.append((bj[keep, None, None] * 3 + c3).ravel())
        vals.append(blk[keep].ravel())
    fx = np.nonzero(~free)[0]
    rows.append((fx[:, None] * 3 + np.arange(3)).ravel()); cols.append((fx[:, None] * 3 + np.arange(3)).ravel())
    vals.append(np.ones(fx.size * 3))
    H = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(3 * V, 3 * V)).tocsr()
    b = np.zeros((V, 3))
    np.add.at(b, ij[:, 0], np.einsum("fij,fj->fi", JiT_O, e))
    np.add.at(b, ij[:, 1], np.einsum("fij,fj->fi", JjT_O, e))
    b[~free] = 0.0
    return H, b.ravel(), float(chi.sum()), chi


def gn_step2(poses, ij, Z, Omega, fixed):
    """One SE(2) Gauss-Newton iteration (direct sparse solve); returns (new poses, stats)."""
    H, b, chi, _ = linearize2(poses, ij, Z, Omega, fixed)
    dx = spla.spsolve(H.tocsc(), -b).reshape(-1, 3)
    new = np.asarray(poses, np.float64) @ v2t2(dx)
    return new, dict(chi=chi, dx_norm_inf=float(np.abs(dx).max()), num_factors=int(ij.shape[0]))


def total_chi(poses, ij, Z, Omega):
    poses = np.asarray(poses, np.float64)
    f = factor_terms2 if poses.shape[-1] == 3 else factor_terms
    return float(f(poses, ij, np.asarray(Z, np.float64), np.asarray(Omega, np.float64))[3].sum())
