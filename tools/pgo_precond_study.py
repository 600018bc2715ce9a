#!/usr/bin/env python
"""CPU study (numpy/scipy, uses the pose-graph oracle): PCG iteration counts on the first Gauss-Newton
system of a C4-shaped graph for candidate preconditioners -- what the CUDA solver should implement
instead of block-Jacobi.  Usage: python tools/pgo_precond_study.py [poses] [factors]"""
import sys, time
sys.path.insert(0, '.')
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
from oracle import pgo_oracle as P
from srrg2_slam_interfaces_b200 import synthetic as syn

V = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
F = int(sys.argv[2]) if len(sys.argv) > 2 else 5 * V
g = syn.make_pose_graph3d(V, F, seed=4)
H, b, chi, _ = P.linearize(g["guess"].astype(np.float64), g["ij"], g["Z"].astype(np.float64), g["Omega"].astype(np.float64), g["fixed"])
H = H.tocsr(); n = H.shape[0]
print("system: %d dofs, %d nnz, chi %.3e" % (n, H.nnz, chi))

def pcg(apply_M, tol=1e-10, maxiter=20000):
    x = np.zeros(n); r = -b.copy(); z = apply_M(r); p = z.copy(); rz = r @ z; b2 = b @ b
    for it in range(1, maxiter + 1):
        Ap = H @ p; a = rz / (p @ Ap); x += a * p; r -= a * Ap
        if np.sqrt((r @ r) / b2) < tol: return it
        z = apply_M(r); rz2 = r @ z; p = z + (rz2 / rz) * p; rz = rz2
    return maxiter

# (a) block-Jacobi
D = np.stack([H[6 * v:6 * v + 6, 6 * v:6 * v + 6].toarray() for v in range(V)])
Dinv = np.linalg.inv(D)
def bj(r): return np.einsum("vij,vj->vi", Dinv, r.reshape(V, 6)).ravel()
t = time.time(); print("block-Jacobi: %d iterations (%.1f s)" % (pcg(bj), time.time() - t))

# (b) two-level additive: block-Jacobi + Galerkin coarse correction, piecewise-CONSTANT perturbations on
# aggregates of m consecutive poses
for m in (8, 32, 128):
    agg = np.arange(V) // m; nc = agg.max() + 1
    rows = np.arange(n); cols = agg[rows // 6] * 6 + rows % 6
    Pm = sp.csr_matrix((np.ones(n), (rows, cols)), shape=(n, nc * 6))
    Hc = (Pm.T @ H @ Pm).tocsc()
    lu = spla.splu(Hc)
    def two(r, lu=lu, Pm=Pm): return bj(r) + Pm @ lu.solve(Pm.T @ r)
    t = time.time(); print("block-Jacobi + coarse (aggregates of %d poses, %d coarse dofs): %d iterations (%.1f s)" % (m, nc * 6, pcg(two), time.time() - t))

# (c) same two-level scheme with the RIGID-MOTION coarse space: a left rigid motion (dt, w) of an aggregate,
# expressed in every member pose's own frame (the variables are right perturbations):
#   dt_i = R_i^T (dt + w x (t_i - c)),  dq_i = R_i^T w / 2
X = g["guess"].astype(np.float64)
R = X[:, :3, :3]; tr = X[:, :3, 3]
free = ~np.asarray(g["fixed"], bool)
for m in (8, 32, 128):
    agg = np.arange(V) // m; nc = agg.max() + 1
    c = np.zeros((nc, 3)); np.add.at(c, agg, tr); c /= np.bincount(agg)[:, None]
    Rt = np.swapaxes(R, 1, 2)
    Pi = np.zeros((V, 6, 6))
    Pi[:, :3, :3] = Rt
    Pi[:, :3, 3:] = -Rt @ P.skew(tr - c[agg])
    Pi[:, 3:, 3:] = 0.5 * Rt
    Pi[~free] = 0.0
    rows = (np.arange(V)[:, None, None] * 6 + np.arange(6)[None, :, None] + 0 * np.arange(6)[None, None, :]).ravel()
    cols = (agg[:, None, None] * 6 + 0 * np.arange(6)[None, :, None] + np.arange(6)[None, None, :]).ravel()
    Pm = sp.csr_matrix((Pi.ravel(), (rows, cols)), shape=(n, nc * 6))
    Hc = (Pm.T @ H @ Pm).tocsc() + 1e-9 * sp.identity(nc * 6, format="csc")
    lu = spla.splu(Hc)
    def rigid(r, lu=lu, Pm=Pm): return bj(r) + Pm @ lu.solve(Pm.T @ r)
    t = time.time(); print("block-Jacobi + RIGID-MODE coarse space (aggregates of %d poses, %d coarse dofs): %d iterations (%.1f s)" % (m, nc * 6, pcg(rigid), time.time() - t), flush=True)

# (d) odometry chain (block-tridiagonal part of H) solved exactly
Hcoo = H.tocoo(); keep = np.abs(Hcoo.row // 6 - Hcoo.col // 6) <= 1
T = sp.csc_matrix((Hcoo.data[keep], (Hcoo.row[keep], Hcoo.col[keep])), shape=H.shape)
luT = spla.splu(T)
t = time.time(); print("block-tridiagonal (odometry chain) exact solve: %d iterations (%.1f s)" % (pcg(lambda r: luT.solve(r)), time.time() - t))
