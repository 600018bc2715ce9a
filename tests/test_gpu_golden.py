"""CUDA path (through the C ABI) against the committed golden vectors: bit-exact."""
import numpy as np
import pytest

from golden_util import CASES, check_run, load

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", CASES)
def test_cuda_reproduces_golden(capi, name):
    g = load(name)
    dim, nm = g["dim"], g["moving"].shape[0]
    ctx = capi.Context(dim)
    ctx.set_cloud(capi.FIXED, 3, g["fixed"], g["fixed_normals"], g["fixed_valid"])
    ctx.set_cloud(capi.MOVING, 3, g["moving"], g["moving_normals"], g["moving_valid"])
    fp, fa = capi.finder_params(**g["fp_kw"]), capi.factor_params(**g["fa_kw"])
    ap = capi.aligner_params(**g["ap_kw"])
    fi, mi, rs = ctx.find_correspondences(3, g["T0"], fp, nm)
    dense = np.full(nm, -1, np.int32)
    dense[mi] = fi
    assert np.array_equal(dense, g["find0_fixed"])
    lin = ctx.linearize(3, g["T0"], fp, fa, variable=ap.variable, n_moving=nm)
    assert np.array_equal(lin["acc"], g["lin0_acc"])
    assert np.array_equal(lin["H"], g["lin0_H"]) and np.array_equal(lin["b"], g["lin0_b"])
    for _ in range(2):  # second run re-uses the index and the warm-start state
        r = ctx.icp_run([capi.make_slice(dim, 3, None, fp, fa)], ap, g["T0"])
        check_run(g, r["T"], r["status"], r["stats"], ctx.get_correspondences(3, nm))
    ctx.close()
