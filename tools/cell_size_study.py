#!/usr/bin/env python
"""CPU study (numpy / scipy + the oracle): how many candidates the warm-started grid search of iteration 2 has to
examine per query at C2, as a function of the cell edge -- the lever DESIGN.md section 5 names for the searching
iterations (their cost is candidates x divergence, not bytes).  A query with warm-start distance d0 examines every
fixed point in the rows (y, z cells) its d0-ball touches, within x in [q_x - d0, q_x + d0] (+ one fine x cell).
Usage: python tools/cell_size_study.py [n_points] [n_sample]"""
import sys
sys.path.insert(0, '.')
import numpy as np
from scipy.spatial import cKDTree
from oracle import oracle as O
from srrg2_slam_interfaces_b200 import synthetic as syn

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
d = syn.make_icp3d(n, n, seed=2)
F, M = O.CloudRef(d["fixed"], d["fixed_normals"]), O.CloudRef(d["moving"], d["moving_normals"])
fp, fa = O.finder_params(0.3, 0.8), O.factor_params(O.FACTOR_PLANE, O.ROB_HUBER, 0.01)
sl = [O.make_slice(F, M, None, fp, fa)]
r1 = O.icp_run(3, sl, O.aligner_params(max_iterations=1, min_num_inliers=10), np.eye(4))
r2 = O.icp_run(3, sl, O.aligner_params(max_iterations=2, min_num_inliers=10), np.eye(4))
fi, mi, _ = r1["correspondences"][0]          # neighbours found in iteration 1 (at the identity guess)
T1 = r1["T"].astype(np.float64)               # pose the searches of iteration 2 run at
rng = np.random.default_rng(0)
sel = rng.choice(len(mi), size=min(ns, len(mi)), replace=False)
q = d["moving"][mi[sel]].astype(np.float64) @ T1[:3, :3].T + T1[:3, 3]
f = d["fixed"].astype(np.float64)
d0 = np.linalg.norm(q - f[fi[sel]], axis=1)   # warm-start radius of iteration 2
tree = cKDTree(f)
print("C2, iteration 2: warm-start distance d0: median %.3f m, mean %.3f m, p90 %.3f m (max_distance 0.3)" % (np.median(d0), d0.mean(), np.quantile(d0, 0.9)))
true_in_ball = np.array([len(tree.query_ball_point(q[k], d0[k])) for k in range(len(sel))])
print("points inside the d0 ball (the floor of any exact search): mean %.1f" % true_in_ball.mean())
for R in (2, 4, 6, 8, 12, 16):
    cell = 1.05 * 0.3 / R
    xf = 4
    cnt = np.zeros(len(sel))
    for k in range(len(sel)):
        lo_yz = np.floor((q[k, 1:] - d0[k]) / cell) * cell
        hi_yz = (np.floor((q[k, 1:] + d0[k]) / cell) + 1) * cell
        xc = cell / xf
        lo_x = np.floor((q[k, 0] - d0[k]) / xc) * xc
        hi_x = (np.floor((q[k, 0] + d0[k]) / xc) + 1) * xc
        c = 0.5 * np.array([lo_x + hi_x, lo_yz[0] + hi_yz[0], lo_yz[1] + hi_yz[1]])
        h = 0.5 * np.array([hi_x - lo_x, hi_yz[0] - lo_yz[0], hi_yz[1] - lo_yz[1]])
        idx = tree.query_ball_point(c, h.max() + 1e-9, p=np.inf)
        p = f[idx]
        cnt[k] = np.count_nonzero(np.all(np.abs(p - c) <= h, axis=1))
    rows = np.mean([(np.floor((q[k, 1] + d0[k]) / cell) - np.floor((q[k, 1] - d0[k]) / cell) + 1) *
                    (np.floor((q[k, 2] + d0[k]) / cell) - np.floor((q[k, 2] - d0[k]) / cell) + 1) for k in range(len(sel))])
    print("R = %2d (cell %.3f m): candidates in the box of touched rows: mean %.1f (p90 %.0f), rows touched: mean %.1f" % (R, cell, cnt.mean(), np.quantile(cnt, 0.9), rows))

# ---- the order of the walk matters more than the cell size: simulate the product's phase 1 (centre row with the
# radius the warm start gives, then the ring-1 rows that survive) against a variant that first looks at the query's own
# fine x cell (+-1) of the centre row, and only then walks the rest of the row with the radius that established ----
def box_points(c_lo, c_hi):
    c = 0.5 * (c_lo + c_hi); h = 0.5 * (c_hi - c_lo)
    idx = tree.query_ball_point(c, h.max() + 1e-9, p=np.inf)
    p = f[idx]
    return p[np.all(np.abs(p - c) <= h + 1e-12, axis=1)]

def simulate(qk, r0, cell, xf, probe_first):
    xc = cell / xf
    cy, cz = np.floor(qk[1] / cell), np.floor(qk[2] / cell)
    r, n = r0, 0
    def scan(y, z, gap2, x_lo, x_hi, skip=None):
        nonlocal r, n
        lo = np.array([np.floor(x_lo / xc) * xc, y * cell, z * cell]); hi = np.array([(np.floor(x_hi / xc) + 1) * xc, (y + 1) * cell, (z + 1) * cell])
        p = box_points(lo, hi)
        if skip is not None:
            p = p[(p[:, 0] < skip[0]) | (p[:, 0] >= skip[1])]
        n += len(p)
        if len(p):
            r = min(r, np.sqrt(((p - qk) ** 2).sum(axis=1).min()))
    skip = None
    if probe_first:
        lo_x, hi_x = qk[0] - xc, qk[0] + xc
        scan(cy, cz, 0.0, lo_x, hi_x)
        skip = (np.floor(lo_x / xc) * xc, (np.floor(hi_x / xc) + 1) * xc)
    scan(cy, cz, 0.0, qk[0] - r, qk[0] + r, skip)
    fy, fz = qk[1] / cell - cy, qk[2] / cell - cz
    for dy in (-1, 0, 1):
        for dz in (-1, 0, 1):
            if dy == 0 and dz == 0:
                continue
            gy = (fy if dy < 0 else (1 - fy if dy > 0 else 0.0)) * cell
            gz = (fz if dz < 0 else (1 - fz if dz > 0 else 0.0)) * cell
            g2 = gy * gy + gz * gz
            if g2 > r * r:
                continue
            w = np.sqrt(max(r * r - g2, 0.0))
            scan(cy + dy, cz + dz, g2, qk[0] - w, qk[0] + w)
    return n, r

cell, xf = 1.05 * 0.3 / 4, 4
for name, r0 in (("iteration 1 (no warm start: radius = max_distance)", np.full(len(sel), 0.3)), ("iteration 2 (warm start d0)", np.minimum(d0, 0.3))):
    a = np.array([simulate(q[k], r0[k], cell, xf, False)[0] for k in range(min(len(sel), 1500))])
    b = np.array([simulate(q[k], r0[k], cell, xf, True)[0] for k in range(min(len(sel), 1500))])
    print("%s, R = 4: candidates per query in rings 0-1: row order as built mean %.1f (p90 %.0f) | own fine x cell first mean %.1f (p90 %.0f)"
          % (name, a.mean(), np.quantile(a, 0.9), b.mean(), np.quantile(b, 0.9)))
