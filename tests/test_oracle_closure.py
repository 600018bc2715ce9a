"""CPU check of the oracle's restatement of the brute-force loop detector's candidate loop
(R/registration/loop_detector/multi_loop_detector_brute_force_impl.cpp:63-133): gate order and arithmetic against
the aligner results they are derived from."""
import numpy as np

from srrg2_slam_interfaces_b200 import synthetic as syn


def test_closure_loop_gates(oracle):
    O = oracle
    base = syn.make_icp3d(8000, 16, seed=5)
    F = O.CloudRef(base["fixed"], base["fixed_normals"])
    cands, guesses, keep = [], [], []
    for k, (n, seed) in enumerate([(4000, 5), (3000, 77), (120, 5)]):
        d = syn.make_icp3d(16, n, seed=seed, moving_stream=k + 1)
        keep.append(O.CloudRef(d["moving"], d["moving_normals"]))  # (a slice only holds the pointers)
        cands.append([O.make_slice(F, keep[-1], None, O.finder_params(0.3, 0.8),
                                   O.factor_params(O.FACTOR_PLANE, O.ROB_HUBER, 0.01))])
        guesses.append(np.eye(4, dtype=np.float32))
    ap = O.aligner_params(max_iterations=8, min_num_inliers=10)
    res = O.closure_loop(3, cands, ap, guesses, 500, 0.005, 0.7)
    assert [r["verdict"] for r in res][0] == O.CLOSURE_ACCEPT
    assert res[1]["verdict"] != O.CLOSURE_ACCEPT and res[2]["verdict"] == O.CLOSURE_NUM_INLIERS_DROP
    for sl, g, r in zip(cands, guesses, res):
        o = O.icp_run(3, sl, ap, g)
        assert np.array_equal(o["T"], r["T"]) and o["status"] == r["aligner_status"]
        if o["status"] == 0:
            st = o["stats"][-1]
            assert r["num_correspondences"] == len(o["correspondences"][0][0])
            assert r["num_inliers"] == st["num_inliers"]
            assert np.isclose(float(r["chi_inliers"]), st["chi_inliers"] / st["num_inliers"], rtol=1e-6)
    # the gates are tried in the reference's order: loosening one exposes the next
    res2 = O.closure_loop(3, cands[2:], ap, guesses[2:], 10, 1e-9, 0.7)
    assert res2[0]["verdict"] == O.CLOSURE_MAX_CHI_DROP
    res3 = O.closure_loop(3, cands[2:], ap, guesses[2:], 10, 1e9, 1.01)
    assert res3[0]["verdict"] == O.CLOSURE_INLIER_RATIO_DROP


def test_relocalize_selection_matches_the_reference_rule(oracle):
    """MultiRelocalizer_::compute (R/registration/relocalization/multi_relocalizer_impl.cpp:74-138): pre-filter by the
    guess's translation, the detector's gates, then the smallest chi per inlier -- first candidate on ties.  The host-side
    selection of the product (capi.relocalize_select, pure Python over closure_batch results) against the oracle loop."""
    from srrg2_slam_interfaces_b200 import capi as A
    O = oracle
    base = syn.make_icp3d(8000, 16, seed=5)
    F = O.CloudRef(base["fixed"], base["fixed_normals"])
    keep, cands, guesses = [], [], []
    for k, (n, seed, noise) in enumerate([(4000, 5, 0.004), (4000, 5, 0.002), (3000, 77, 0.002), (4000, 5, 0.002)]):
        d = syn.make_icp3d(16, n, seed=seed, moving_stream=k + 1, noise=noise)
        keep.append(O.CloudRef(d["moving"], d["moving_normals"]))
        cands.append([O.make_slice(F, keep[-1], None, O.finder_params(0.3, 0.8), O.factor_params(O.FACTOR_PLANE, O.ROB_HUBER, 0.01))])
        guesses.append(np.eye(4, dtype=np.float32))
    ap = O.aligner_params(max_iterations=8, min_num_inliers=10)
    translations = [0.5, 0.2, 0.1, 9.0]  # the last candidate's guess is too far away: never aligned
    best, res = O.relocalize_loop(3, cands, ap, guesses, translations, 3.0, 500, 0.005, 0.7)
    assert res[3] is None and res[2]["verdict"] != O.CLOSURE_ACCEPT
    assert best == 1 and res[1]["chi_inliers"] < res[0]["chi_inliers"]  # the less noisy of the two accepted maps
    # the product's selection over the same per-candidate results (closure_batch returns them for every candidate)
    full = O.closure_loop(3, cands, ap, guesses, 500, 0.005, 0.7)
    assert A.relocalize_select(full, translations, 3.0) == best
    assert A.relocalize_select(full) == min((k for k in range(4) if full[k]["verdict"] == 0), key=lambda k: (float(full[k]["chi_inliers"]), k))
    # ties: the first accepted candidate wins (strict <, :121)
    tie = [dict(verdict=0, chi_inliers=np.float32(0.001)), dict(verdict=0, chi_inliers=np.float32(0.001)), dict(verdict=2, chi_inliers=np.float32(0.0))]
    assert A.relocalize_select(tie) == 0
    assert A.relocalize_select([dict(verdict=1, chi_inliers=np.float32(0))]) is None
