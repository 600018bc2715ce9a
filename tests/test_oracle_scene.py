"""CPU checks of the oracle's scene functions against plain numpy restatements of the reference
semantics: SceneClipper's contract (R/mapping/scene_clipper.h:104-107) and MergerCorrespondenceHomo_::compute
(R/mapping/merger_correspondence_homo_impl.cpp:11-126).  Indices, counts and control flow must agree exactly; the
coordinates to fp32 rounding (the oracle pins the operation order with explicit fmaf, numpy does not)."""
import numpy as np

from srrg2_slam_interfaces_b200 import synthetic as syn


def _np_merge(scene, scene_n, meas, meas_n, meas_valid, T, corr, max_resp, max_d2, target):
    sc, sn = scene.astype(np.float64).copy(), scene_n.astype(np.float64).copy()
    R, t = T[:-1, :-1].astype(np.float64), T[:-1, -1].astype(np.float64)
    merged = set()
    if corr is not None:
        for g, j, r in zip(*corr):
            if not r < max_resp:
                continue
            q = R @ meas[j] + t
            if not np.sum((q - sc[g]) ** 2) < max_d2:
                continue
            sn[g] = meas_n[j]            # the copied point keeps its un-rotated normal (:71)
            sc[g] = (q + sc[g]) / 2      # :74
            merged.add(int(j))
    added = []
    if corr is None or len(merged) < target:
        for j in range(meas.shape[0]):
            if j in merged or (meas_valid is not None and not meas_valid[j]):
                continue
            added.append(j)
    if added:
        sc = np.vstack([sc, meas[added] @ R.T + t])
        sn = np.vstack([sn, meas_n[added] @ R.T])
    return sc, sn, len(merged), len(added)


def test_scene_clip_matches_numpy(oracle):
    d = syn.make_icp3d(100, 20000, seed=9)
    valid = (np.random.default_rng(1).uniform(size=20000) < 0.9).astype(np.uint8)
    T = syn.iso3([0.4, -0.3, 0.2], [0.05, -0.02, 0.03]).astype(np.float32)
    c, n, g = oracle.scene_clip(oracle.CloudRef(d["moving"], d["moving_normals"], valid), T, 7.5)
    q = d["moving"].astype(np.float64) @ T[:3, :3].T.astype(np.float64) + T[:3, 3]
    r = np.linalg.norm(q, axis=1)
    keep = (valid != 0) & (r <= 7.5)
    border = np.abs(r - 7.5) < 1e-4  # (fp32 vs fp64 at the rim)
    assert np.array_equal(np.isin(np.arange(20000), g)[~border], keep[~border])
    assert np.all(np.diff(g) > 0)
    assert np.allclose(c, q[g], atol=2e-6) and np.allclose(n, (d["moving_normals"].astype(np.float64) @ T[:3, :3].T)[g], atol=1e-6)


def test_scene_merge_matches_numpy(oracle):
    rng = np.random.default_rng(4)
    d = syn.make_icp3d(4000, 9000, seed=5)
    scene, scene_n, meas, meas_n = d["moving"], d["moving_normals"], d["fixed"], d["fixed_normals"]
    meas_valid = (rng.uniform(size=4000) < 0.95).astype(np.uint8)
    T = syn.iso3([0.1, -0.05, 0.08], [0.02, -0.015, 0.03]).astype(np.float32)
    # correspondences: every scene index at most once (the aligner's one entry per moving point), measurement indices repeat
    g = rng.permutation(9000)[:3000].astype(np.int32)
    j = rng.integers(0, 4000, size=3000).astype(np.int32)
    q = meas[j].astype(np.float64) @ T[:3, :3].T + T[:3, 3]
    close = rng.uniform(size=3000) < 0.5
    scene = scene.copy()
    scene[g[close]] = (q[close] + rng.normal(scale=0.02, size=(int(close.sum()), 3))).astype(np.float32)
    r = rng.uniform(0, 0.4, size=3000).astype(np.float32)
    M = oracle.CloudRef(meas, meas_n, meas_valid)
    for target in (100, 10**9):
        es, en, ev, nm, na = oracle.scene_merge(scene, scene_n, None, M, T, (g, j, r), 0.3, 0.01, target)
        rs, rn, rnm, rna = _np_merge(scene, scene_n, meas, meas_n, meas_valid, T, (g, j, r), 0.3, 0.01, target)
        assert (nm, na) == (rnm, rna) and es.shape[0] == rs.shape[0]
        assert np.allclose(es, rs, atol=2e-6) and np.allclose(en, rn, atol=1e-6)
        assert nm > 500 and ((na == 0) if target == 100 else (na > 0))
    es, en, ev, nm, na = oracle.scene_merge(scene, scene_n, None, M, T, None)
    assert nm == 0 and na == int(meas_valid.sum()) and es.shape[0] == 9000 + na
