"""The oracle against the committed golden vectors (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from golden_util import CASES, check_run, load


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_golden(oracle, name):
    g = load(name)
    F = oracle.CloudRef(g["fixed"], g["fixed_normals"], g["fixed_valid"])
    M = oracle.CloudRef(g["moving"], g["moving_normals"], g["moving_valid"])
    fp, fa = oracle.finder_params(**g["fp_kw"]), oracle.factor_params(**g["fa_kw"])
    ap = oracle.aligner_params(**g["ap_kw"])
    for method in (oracle.NN_KDTREE, oracle.NN_BRUTE):
        r = oracle.icp_run(g["dim"], [oracle.make_slice(F, M, None, fp, fa, dim=g["dim"])], ap, g["T0"], nn_method=method)
        check_run(g, r["T"], r["status"], r["stats"], r["correspondences"][0])
    fidx0, resp0 = oracle.find(oracle.Index(F), F, M, g["T0"], fp)
    assert np.array_equal(fidx0, g["find0_fixed"]) and np.array_equal(resp0, g["find0_resp"])
    lin = oracle.linearize(F, M, fidx0, g["T0"], fp, fa, variable=ap.variable)
    assert np.array_equal(lin["acc"], g["lin0_acc"])
