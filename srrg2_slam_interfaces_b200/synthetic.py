"""Seeded synthetic clouds / graphs for the BASELINE.json configurations.

Pure numpy, PCG64 streams: the same seed gives bit-identical inputs in the tests, in bench.py, for
the CPU oracle and for the CUDA path.  Nothing here touches the GPU.
"""
import numpy as np


def rot3(rpy):
    r, p, y = [float(v) for v in rpy]
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def iso3(t, rpy):
    T = np.eye(4)
    T[:3, :3] = rot3(rpy)
    T[:3, 3] = t
    return T


def iso2(x, y, th):
    c, s = np.cos(th), np.sin(th)
    return np.array([[c, -s, x], [s, c, y], [0, 0, 1.0]])


def inv_iso(T):
    d = T.shape[0] - 1
    Ti = np.eye(d + 1)
    Ti[:d, :d] = T[:d, :d].T
    Ti[:d, d] = -T[:d, :d].T @ T[:d, d]
    return Ti


def transform_cloud(T, pts, nrm=None):
    d = T.shape[0] - 1
    p = pts.astype(np.float64) @ T[:d, :d].T + T[:d, d]
    n = None if nrm is None else nrm.astype(np.float64) @ T[:d, :d].T
    return p.astype(np.float32), (None if n is None else n.astype(np.float32))


# ----------------------------------------------------------------------------------------------
# 3D scene: planar patches + spheres inside a cube
# ----------------------------------------------------------------------------------------------
class Scene3D:
    def __init__(self, seed, n_planes=64, n_spheres=16, cube=20.0):
        rng = np.random.default_rng([seed, 0xA11])
        self.cube = float(cube)
        h = cube / 2
        self.pl_c = rng.uniform(-0.7 * h, 0.7 * h, size=(n_planes, 3))
        nrm = rng.normal(size=(n_planes, 3))
        self.pl_n = nrm / np.linalg.norm(nrm, axis=1, keepdims=True)
        a = np.cross(self.pl_n, rng.normal(size=(n_planes, 3)))
        self.pl_u = a / np.linalg.norm(a, axis=1, keepdims=True)
        self.pl_v = np.cross(self.pl_n, self.pl_u)
        self.pl_l = rng.uniform(0.3 * cube, 0.7 * cube, size=(n_planes, 2))
        self.sp_c = rng.uniform(-0.6 * h, 0.6 * h, size=(n_spheres, 3))
        self.sp_r = rng.uniform(0.05 * cube, 0.15 * cube, size=n_spheres)
        area = np.concatenate([self.pl_l[:, 0] * self.pl_l[:, 1], 4 * np.pi * self.sp_r ** 2])
        self.prob = area / area.sum()
        self.n_planes = n_planes

    def sample(self, n, rng, jitter=0.0):
        which = rng.choice(len(self.prob), size=n, p=self.prob)
        pts = np.empty((n, 3))
        nrm = np.empty((n, 3))
        isp = which < self.n_planes
        k = which[isp]
        uv = rng.uniform(-0.5, 0.5, size=(k.size, 2)) * self.pl_l[k]
        pts[isp] = self.pl_c[k] + uv[:, :1] * self.pl_u[k] + uv[:, 1:] * self.pl_v[k]
        nrm[isp] = self.pl_n[k]
        k = which[~isp] - self.n_planes
        d = rng.normal(size=(k.size, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        pts[~isp] = self.sp_c[k] + self.sp_r[k, None] * d
        nrm[~isp] = d
        if jitter > 0:
            pts += rng.uniform(-jitter, jitter, size=pts.shape)
        return pts, nrm


C2_T_STAR = iso3([0.10, -0.05, 0.08], np.deg2rad([1.5, -1.0, 2.0]))


def make_icp3d(n_fixed, n_moving, seed=2, T_star=None, noise=0.002, outlier_frac=0.05, cube=20.0,
               jitter=0.0005, n_planes=64, n_spheres=16, moving_stream=0):
    """Config C2 generator: fixed and moving are INDEPENDENT samples of the same
    surfaces; moving = T*^-1 (sample + N(0,noise)), a fraction replaced by uniform outliers.
    The aligner estimate (moving in fixed) should converge to T*."""
    T_star = C2_T_STAR if T_star is None else T_star
    scene = Scene3D(seed, n_planes, n_spheres, cube)
    rf = np.random.default_rng([seed, 1])
    rm = np.random.default_rng([seed, 2] if moving_stream == 0 else [seed, 2, moving_stream])
    fp, fn = scene.sample(n_fixed, rf, jitter)
    mp, mn = scene.sample(n_moving, rm, 0.0)
    mp = mp + rm.normal(scale=noise, size=mp.shape)
    n_out = int(outlier_frac * n_moving)
    if n_out:
        idx = rm.choice(n_moving, size=n_out, replace=False)
        mp[idx] = rm.uniform(-cube / 2, cube / 2, size=(n_out, 3))
        d = rm.normal(size=(n_out, 3))
        mn[idx] = d / np.linalg.norm(d, axis=1, keepdims=True)
    Ti = inv_iso(T_star)
    mp32, mn32 = transform_cloud(Ti, mp, mn)
    return dict(fixed=fp.astype(np.float32), fixed_normals=fn.astype(np.float32), moving=mp32,
                moving_normals=mn32, T_star=T_star.astype(np.float64))


# ----------------------------------------------------------------------------------------------
# 2D scene: wall segments
# ----------------------------------------------------------------------------------------------
C1_T_STAR = iso2(0.05, -0.03, np.deg2rad(2.0))


def walls2d(seed, n_walls=8, half=10.0):
    rng = np.random.default_rng([seed, 0xB22])
    a = rng.uniform(-half, half, size=(n_walls, 2))
    ang = rng.uniform(0, np.pi, size=n_walls)
    length = rng.uniform(0.4 * half, 1.2 * half, size=n_walls)
    d = np.stack([np.cos(ang), np.sin(ang)], axis=1)
    return a, d, length


def sample_walls2d(walls, n, rng, jitter=0.0):
    a, d, length = walls
    p = length / length.sum()
    which = rng.choice(len(p), size=n, p=p)
    s = rng.uniform(-0.5, 0.5, size=n) * length[which]
    pts = a[which] + s[:, None] * d[which]
    nrm = np.stack([-d[which, 1], d[which, 0]], axis=1)
    if jitter > 0:
        pts += rng.uniform(-jitter, jitter, size=pts.shape)
    return pts, nrm


def make_icp2d(n_fixed, n_moving=None, seed=1, T_star=None, noise=0.005, paired=True, n_walls=8, half=10.0):
    """Config C1 generator: fixed on wall segments; moving = T*^-1 (fixed + N(0,noise)) when
    paired, else independent samples of the same walls."""
    T_star = C1_T_STAR if T_star is None else T_star
    walls = walls2d(seed, n_walls, half)
    rf = np.random.default_rng([seed, 1])
    rm = np.random.default_rng([seed, 2])
    fp, fn = sample_walls2d(walls, n_fixed, rf, 0.001)
    if paired:
        mp, mn = fp.copy(), fn.copy()
        if n_moving is not None and n_moving != n_fixed:
            sel = rm.choice(n_fixed, size=n_moving, replace=n_moving > n_fixed)
            mp, mn = mp[sel], mn[sel]
    else:
        mp, mn = sample_walls2d(walls, n_moving or n_fixed, rm, 0.0)
    mp = mp + rm.normal(scale=noise, size=mp.shape)
    mp32, mn32 = transform_cloud(inv_iso(T_star), mp, mn)
    return dict(fixed=fp.astype(np.float32), fixed_normals=fn.astype(np.float32), moving=mp32,
                moving_normals=mn32, T_star=T_star.astype(np.float64))


def pose_error(T, T_ref):
    """(rotation error [rad], translation error [m]) between two isometries."""
    d = T.shape[0] - 1
    E = inv_iso(np.asarray(T_ref, dtype=np.float64)) @ np.asarray(T, dtype=np.float64)
    if d == 3:
        c = np.clip((np.trace(E[:3, :3]) - 1) / 2, -1, 1)
        ang = float(np.arccos(c))
        if ang < 1e-3:  # arccos loses precision near 0: use the skew part
            w = np.array([E[2, 1] - E[1, 2], E[0, 2] - E[2, 0], E[1, 0] - E[0, 1]]) / 2
            ang = float(np.linalg.norm(w))
    else:
        ang = float(abs(np.arctan2(E[1, 0], E[0, 0])))
    return ang, float(np.linalg.norm(E[:d, d]))


# ----------------------------------------------------------------------------------------------
# Config C5: 2D multi-cue -- two 1080-beam scanners (fixed side) vs a large local map (moving side)
# ----------------------------------------------------------------------------------------------
def corridor2d(seed, length=60.0, width=4.0, n_doors=10):
    """Wall segments of a corridor world with side rooms: (start points, unit directions, lengths)."""
    rng = np.random.default_rng([seed, 0xC55])
    a = [[-length / 2, -width / 2], [-length / 2, width / 2]]
    d = [[1.0, 0.0], [1.0, 0.0]]
    ln = [length, length]
    for k in range(n_doors):
        x = rng.uniform(-length / 2 + 2, length / 2 - 2)
        side = 1.0 if k % 2 else -1.0
        depth = rng.uniform(2.0, 6.0)
        w = rng.uniform(1.5, 4.0)
        y0 = side * width / 2
        for (sx, sy, dx, dy, l) in ((x, y0, 0.0, side, depth), (x + w, y0, 0.0, side, depth), (x, y0 + side * depth, 1.0, 0.0, w)):
            a.append([sx, sy]); d.append([dx, dy]); ln.append(l)
    a, d, ln = np.array(a), np.array(d), np.array(ln)
    return a + 0.5 * ln[:, None] * d, d, ln  # centre form used by sample_walls2d


def make_multicue2d(n_map, n_beams=1080, seed=5, T_star=None, noise=0.004):
    """Two laser scans (fixed, in their sensor frames) against one local map (moving, robot/map
    frame) + the sensor mounting poses.  The aligner estimate (map in robot) converges to T*."""
    T_star = iso2(0.08, -0.05, np.deg2rad(1.5)) if T_star is None else T_star
    walls = corridor2d(seed)
    rm = np.random.default_rng([seed, 2])
    mp, mn = sample_walls2d(walls, n_map, rm, 0.0)
    mp = mp + rm.normal(scale=noise, size=mp.shape)
    map_pts, map_nrm = transform_cloud(inv_iso(T_star), mp, mn)  # map expressed so that T* aligns it
    sensors = [iso2(0.2, 0.0, np.deg2rad(90.0)), iso2(-0.2, 0.0, np.deg2rad(-90.0))]  # sensor in robot
    scans = []
    for k, sir in enumerate(sensors):
        rs = np.random.default_rng([seed, 10 + k])
        sp, sn = sample_walls2d(walls, 8 * n_beams, rs, 0.002)
        keep = np.nonzero(np.abs(sp[:, 0]) < 12.0)[0][:n_beams]  # the scanner sees the world near the robot
        sp, sn = sp[keep], sn[keep]
        p, n = transform_cloud(inv_iso(sir), sp, sn)  # world(robot) -> sensor frame
        scans.append(dict(points=p, normals=n, sensor_in_robot=sir, robot_in_sensor=inv_iso(sir)))
    return dict(map=map_pts, map_normals=map_nrm, scans=scans, T_star=T_star)


# ----------------------------------------------------------------------------------------------
# Config C3: synthetic RGB-D sequence (pinhole depth images of a box room with boxes inside)
# ----------------------------------------------------------------------------------------------
def _ray_aabb(o, d, lo, hi):
    """Slab intersection of rays o + t d with the box [lo, hi]: (t_near, t_far, axis_near, axis_far)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        t1 = (lo - o) / d
        t2 = (hi - o) / d
    tmin, tmax = np.minimum(t1, t2), np.maximum(t1, t2)
    return tmin.max(axis=1), tmax.min(axis=1), tmin.argmax(axis=1), tmax.argmin(axis=1)


def render_depth_cloud(cam_in_world, width=640, height=480, fx=525.0, fy=525.0, cx=319.5, cy=239.5, seed=3,
                       noise=0.0):
    """Point+normal cloud (camera frame) of one depth image: 6-plane room + 3 boxes."""
    rng = np.random.default_rng([seed, 0xD33])
    room_lo, room_hi = np.array([-2.0, -1.4, -3.0]), np.array([2.0, 1.4, 4.0])  # side walls, floor, ceiling in view
    boxes = []
    for _ in range(3):
        c = rng.uniform([-1.2, 0.4, 1.8], [1.2, 1.0, 3.2])
        h = rng.uniform(0.2, 0.4, size=3)
        boxes.append((c - h, c + h))
    v, u = np.mgrid[0:height, 0:width]
    rays_c = np.stack([(u.ravel() - cx) / fx, (v.ravel() - cy) / fy, np.ones(u.size)], axis=1)
    R, o = cam_in_world[:3, :3], cam_in_world[:3, 3]
    d = rays_c @ R.T
    oo = np.broadcast_to(o, d.shape)
    _, tfar, _, afar = _ray_aabb(oo, d, room_lo, room_hi)
    t = tfar.copy()
    nw = np.zeros_like(d)
    nw[np.arange(d.shape[0]), afar] = -np.sign(d[np.arange(d.shape[0]), afar])  # inward room normals
    for lo, hi in boxes:
        tn, tf, an, _ = _ray_aabb(oo, d, lo, hi)
        hit = (tn > 0) & (tn < tf) & (tn < t)
        t[hit] = tn[hit]
        nb = np.zeros_like(d)
        nb[np.arange(d.shape[0]), an] = -np.sign(d[np.arange(d.shape[0]), an])
        nw[hit] = nb[hit]
    if noise > 0:
        t = t + np.random.default_rng([seed, 0xD34, int(abs(o).sum() * 1e6) % 100000]).normal(scale=noise, size=t.shape)
    pts_c = rays_c * t[:, None]
    nrm_c = nw @ R  # world normal -> camera frame (R^T n)
    valid = (np.isfinite(t) & (t > 0.1) & (t < 20.0)).astype(np.uint8)
    pts_c[valid == 0] = 0.0
    return pts_c.astype(np.float32), nrm_c.astype(np.float32), valid


def make_rgbd_sequence(n_frames=30, width=640, height=480, seed=3, noise=0.001):
    """Camera on a smooth path (<= 2 cm / 0.5 deg per frame); frame k's cloud is the fixed side and
    frame k-1's cloud the moving side of aligner call k; ground truth moving-in-fixed = cam_k^-1 cam_{k-1}."""
    s = width / 640.0
    K = dict(fx=525.0 * s, fy=525.0 * s, cx=319.5 * s + (s - 1) * 0.5, cy=239.5 * s + (s - 1) * 0.5,
             width=width, height=height)
    frames = []
    for k in range(n_frames):
        a = 0.05 * k  # path parameter: <= 2 cm and <= 0.5 deg between consecutive frames
        pose = iso3([0.3 * np.sin(a) + 0.004 * k, -0.2 + 0.1 * np.cos(1.5 * a), 0.012 * k],
                    np.deg2rad([4.0 * np.sin(2 * a), 0.4 * k, 3.0 * np.cos(a)]))
        p, n, v = render_depth_cloud(pose, width, height, K["fx"], K["fy"], K["cx"], K["cy"], seed, noise)
        frames.append(dict(points=p, normals=n, valid=v, pose=pose))
    return frames, K


# ----------------------------------------------------------------------------------------------
# Config C4: Manhattan-3D pose graph (lattice walk confined to a box so that places are revisited)
# ----------------------------------------------------------------------------------------------
def make_pose_graph3d(n_poses, n_factors, seed=4, box=(40, 40, 4), sigma_t=0.02, sigma_r_deg=0.5, max_loop_dist=2.0):
    """Axis-aligned 1 m steps / 90 deg turns inside a box (reflecting walls); n_poses-1 odometry
    factors + loop factors between poses within max_loop_dist; measurements = truth (+) noise,
    Omega = diag(1/sigma^2); initial guess = integrated noisy odometry; pose 0 is the gauge."""
    rng = np.random.default_rng([seed, 0xE44])
    box = np.asarray(box)
    dirs = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]])
    pos = np.zeros((n_poses, 3), dtype=np.int64)
    pos[0] = box // 2
    heading = np.zeros(n_poses, dtype=np.int64)
    steps = rng.integers(0, 6, size=n_poses)
    keep_dir = rng.uniform(size=n_poses) < 0.7
    for k in range(1, n_poses):
        h = heading[k - 1] if keep_dir[k] else steps[k]
        p = pos[k - 1] + dirs[h]
        if np.any(p < 0) or np.any(p >= box):
            h = h ^ 1
            p = pos[k - 1] + dirs[h]
            if np.any(p < 0) or np.any(p >= box):
                h = int(steps[k]) % 4
                p = np.clip(pos[k - 1] + dirs[h], 0, box - 1)
        pos[k], heading[k] = p, h
    # orientation: x axis along the heading (yaw / pitch multiples of 90 deg)
    truth = np.zeros((n_poses, 4, 4))
    for h in range(6):
        sel = heading == h
        d = dirs[h].astype(float)
        up = np.array([0.0, 0.0, 1.0]) if h < 4 else np.array([1.0, 0.0, 0.0])
        y = np.cross(up, d)
        R = np.stack([d, y, np.cross(d, y)], axis=1)
        truth[sel, :3, :3] = R
    truth[:, :3, 3] = pos
    truth[:, 3, 3] = 1.0
    # candidate loop pairs via a lattice hash (same or neighbouring site, non-consecutive)
    key = (pos[:, 0] * box[1] + pos[:, 1]) * box[2] + pos[:, 2]
    order = np.argsort(key, kind="stable")
    n_loop = n_factors - (n_poses - 1)
    pairs = []
    if n_loop > 0:
        ks = key[order]
        starts = np.nonzero(np.r_[True, ks[1:] != ks[:-1]])[0]
        ends = np.r_[starts[1:], ks.size]
        site_of = {int(ks[s]): (s, e) for s, e in zip(starts, ends)}
        got = 0
        tries = 0
        while got < n_loop and tries < 50:
            tries += 1
            a = rng.integers(0, n_poses, size=2 * (n_loop - got) + 16)
            off = dirs[rng.integers(0, 6, size=a.size)] * (rng.uniform(size=(a.size, 1)) < 0.5)
            q = pos[a] + off
            ok = np.all((q >= 0) & (q < box), axis=1)
            a, q = a[ok], q[ok]
            kq = (q[:, 0] * box[1] + q[:, 1]) * box[2] + q[:, 2]
            for ai, kk in zip(a, kq):
                se = site_of.get(int(kk))
                if se is None:
                    continue
                bi = int(order[rng.integers(se[0], se[1])])
                if abs(bi - int(ai)) > 1 and np.linalg.norm(pos[bi] - pos[ai]) <= max_loop_dist:
                    pairs.append((min(int(ai), bi), max(int(ai), bi)))
                    got += 1
                    if got >= n_loop:
                        break
    odo = np.stack([np.arange(n_poses - 1), np.arange(1, n_poses)], axis=1)
    ij = np.concatenate([odo, np.array(pairs, dtype=np.int64).reshape(-1, 2)], axis=0).astype(np.int32)
    F = ij.shape[0]
    sr = np.deg2rad(sigma_r_deg)
    noise = np.concatenate([rng.normal(scale=sigma_t, size=(F, 3)), rng.normal(scale=sr / 2, size=(F, 3))], axis=1)
    Zt = _inv_iso_batch(truth[ij[:, 0]]) @ truth[ij[:, 1]]
    Z = Zt @ _v2t_batch(noise)
    omega = np.diag(np.r_[np.full(3, 1 / sigma_t ** 2), np.full(3, 1 / (sr / 2) ** 2)])
    Omega = np.broadcast_to(omega, (F, 6, 6)).copy()
    guess = np.empty_like(truth)
    guess[0] = truth[0]
    for k in range(1, n_poses):  # integrated noisy odometry
        guess[k] = guess[k - 1] @ Z[k - 1]
    fixed = np.zeros(n_poses, dtype=np.uint8)
    fixed[0] = 1
    return dict(truth=truth, guess=guess.astype(np.float32), ij=ij, Z=Z.astype(np.float32),
                Omega=Omega.astype(np.float32), fixed=fixed)


def _inv_iso_batch(T):
    Ti = np.zeros_like(T)
    Rt = np.swapaxes(T[..., :3, :3], -1, -2)
    Ti[..., :3, :3] = Rt
    Ti[..., :3, 3] = -np.einsum("...ij,...j->...i", Rt, T[..., :3, 3])
    Ti[..., 3, 3] = 1.0
    return Ti


def _v2t_batch(dx):
    dq = dx[..., 3:6]
    w = np.sqrt(np.maximum(1.0 - np.sum(dq * dq, -1), 0.0))
    x, y, z = dq[..., 0], dq[..., 1], dq[..., 2]
    T = np.zeros(dx.shape[:-1] + (4, 4))
    T[..., 0, 0] = 1 - 2 * (y * y + z * z); T[..., 0, 1] = 2 * (x * y - w * z); T[..., 0, 2] = 2 * (x * z + w * y)
    T[..., 1, 0] = 2 * (x * y + w * z); T[..., 1, 1] = 1 - 2 * (x * x + z * z); T[..., 1, 2] = 2 * (y * z - w * x)
    T[..., 2, 0] = 2 * (x * z - w * y); T[..., 2, 1] = 2 * (y * z + w * x); T[..., 2, 2] = 1 - 2 * (x * x + y * y)
    T[..., :3, 3] = dx[..., :3]
    T[..., 3, 3] = 1.0
    return T


def make_pose_graph2d(n_poses, n_factors, seed=6, box=(30, 30), sigma_t=0.02, sigma_r_deg=0.5, max_loop_dist=2.0):
    """SE(2) analogue of make_pose_graph3d (LoopClosure2D, R/registration/loop_closure.h:110): lattice walk with
    90 deg turns inside a box, odometry + loop factors between poses within max_loop_dist, measurements =
    truth (+) noise, Omega = diag(1/sigma^2), guess = integrated noisy odometry, pose 0 fixed.  3x3 matrices."""
    rng = np.random.default_rng([seed, 0xE22])
    box = np.asarray(box)
    dirs = np.array([[1, 0], [0, 1], [-1, 0], [0, -1]])
    pos = np.zeros((n_poses, 2), dtype=np.int64)
    pos[0] = box // 2
    head = np.zeros(n_poses, dtype=np.int64)
    turn = rng.integers(-1, 2, size=n_poses)
    keep = rng.uniform(size=n_poses) < 0.7
    for k in range(1, n_poses):
        h = head[k - 1] if keep[k] else (head[k - 1] + turn[k]) % 4
        p = pos[k - 1] + dirs[h]
        if np.any(p < 0) or np.any(p >= box):
            h = (h + 2) % 4
            p = pos[k - 1] + dirs[h]
        pos[k], head[k] = p, h

    def iso(x, y, th):
        T = np.zeros(np.shape(x) + (3, 3))
        T[..., 0, 0] = np.cos(th); T[..., 0, 1] = -np.sin(th); T[..., 1, 0] = np.sin(th); T[..., 1, 1] = np.cos(th)
        T[..., 0, 2] = x; T[..., 1, 2] = y; T[..., 2, 2] = 1.0
        return T

    truth = iso(pos[:, 0].astype(np.float64), pos[:, 1].astype(np.float64), head * (np.pi / 2))
    pairs = [(k, k + 1) for k in range(n_poses - 1)]
    cells = {}
    for k in range(n_poses):
        cells.setdefault((int(pos[k, 0]), int(pos[k, 1])), []).append(k)
    cand = []
    for k in range(n_poses):
        for dx in range(-2, 3):
            for dy in range(-2, 3):
                if dx * dx + dy * dy > max_loop_dist ** 2:
                    continue
                for j in cells.get((int(pos[k, 0]) + dx, int(pos[k, 1]) + dy), ()):
                    if j > k + 1:
                        cand.append((k, j))
    cand = np.asarray(cand, dtype=np.int64).reshape(-1, 2)
    n_loops = max(0, min(n_factors - len(pairs), cand.shape[0]))
    if n_loops:
        pairs += [tuple(c) for c in cand[rng.choice(cand.shape[0], size=n_loops, replace=False)]]
    ij = np.asarray(pairs, dtype=np.int32)

    def inv(T):
        o = np.zeros_like(T)
        R = np.swapaxes(T[..., :2, :2], -1, -2)
        o[..., :2, :2] = R
        o[..., :2, 2] = -np.einsum("...ij,...j->...i", R, T[..., :2, 2])
        o[..., 2, 2] = 1.0
        return o

    rel = inv(truth[ij[:, 0]]) @ truth[ij[:, 1]]
    noise = iso(rng.normal(scale=sigma_t, size=ij.shape[0]), rng.normal(scale=sigma_t, size=ij.shape[0]),
                rng.normal(scale=np.deg2rad(sigma_r_deg), size=ij.shape[0]))
    Z = rel @ noise
    Omega = np.tile(np.diag([1 / sigma_t ** 2, 1 / sigma_t ** 2, 1 / np.deg2rad(sigma_r_deg) ** 2]), (ij.shape[0], 1, 1))
    guess = np.zeros_like(truth)
    guess[0] = truth[0]
    for k in range(1, n_poses):
        guess[k] = guess[k - 1] @ Z[k - 1]
    fixed = np.zeros(n_poses, dtype=np.uint8)
    fixed[0] = 1
    return dict(truth=truth, guess=guess.astype(np.float32), ij=ij, Z=Z.astype(np.float32), Omega=Omega.astype(np.float32), fixed=fixed)
