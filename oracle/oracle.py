"""ctypes front-end of the CPU ORACLE (oracle/srrg2b_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsrrg2b_oracle.so")

FACTOR_P2P, FACTOR_PLANE = 0, 1
ROB_NONE, ROB_SATURATED, ROB_CAUCHY, ROB_CLAMP, ROB_HUBER = 0, 1, 2, 3, 4
VAR_SE3_QUAT_RIGHT, VAR_SE3_EULER_RIGHT = 0, 1
FINDER_NN, FINDER_PROJECTIVE = 0, 1
SLICE_POINTS, SLICE_PRIOR = 0, 1
NN_BRUTE, NN_KDTREE = 0, 1
STAT_INLIER, STAT_KERNELIZED, STAT_SUPPRESSED, STAT_NONE = 0, 1, 2, 3


def build(force=False):
    src = os.path.join(_HERE, "srrg2b_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


class FinderParams(C.Structure):
    _fields_ = [("kind", C.c_int32), ("max_distance", C.c_float), ("normal_cos", C.c_float),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("width", C.c_int32), ("height", C.c_int32),
                ("min_depth", C.c_float), ("max_depth", C.c_float)]


class FactorParams(C.Structure):
    _fields_ = [("factor", C.c_int32), ("robustifier", C.c_int32), ("chi_threshold", C.c_float),
                ("info_point", C.c_float), ("info_normal", C.c_float)]


class Cloud(C.Structure):
    _fields_ = [("coords", C.c_void_p), ("normals", C.c_void_p), ("valid", C.c_void_p), ("n", C.c_int64)]


class Slice(C.Structure):
    _fields_ = [("kind", C.c_int32), ("min_num_correspondences", C.c_int32),
                ("fixed", Cloud), ("moving", Cloud),
                ("robot_in_sensor", C.c_float * 16),
                ("finder", FinderParams), ("factor", FactorParams),
                ("prior_measurement", C.c_float * 16), ("prior_info_diag", C.c_float * 6)]


class IterStats(C.Structure):
    _fields_ = [("iteration", C.c_int32), ("solver_status", C.c_int32),
                ("num_inliers", C.c_int64), ("num_outliers", C.c_int64),
                ("num_suppressed", C.c_int64), ("num_correspondences", C.c_int64),
                ("chi_inliers", C.c_double), ("chi_outliers", C.c_double),
                ("num_saturated", C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class AlignerParams(C.Structure):
    _fields_ = [("variable", C.c_int32), ("max_iterations", C.c_int32), ("min_num_inliers", C.c_int32),
                ("enable_inlier_only_runs", C.c_int32), ("keep_only_inlier_correspondences", C.c_int32),
                ("use_termination_criteria", C.c_int32), ("window_size", C.c_int32),
                ("num_correspondences_range", C.c_int32), ("num_inliers_range", C.c_int32),
                ("num_outliers_range", C.c_int32), ("chi_epsilon", C.c_float)]


class CorrOut(C.Structure):
    _fields_ = [("fixed_idx", C.c_void_p), ("moving_idx", C.c_void_p), ("response", C.c_void_p),
                ("n", C.c_int64)]


class Scales(C.Structure):
    _fields_ = [("k", C.c_int32 * 7), ("err_bound", C.c_float)]


ACC_SLOTS = 40


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_index_create.restype = C.c_void_p
        _lib.orc_index_create.argtypes = [C.c_int, C.POINTER(Cloud), C.c_int]
        _lib.orc_index_free.argtypes = [C.c_void_p]
        _lib.orc_find.argtypes = [C.c_void_p, C.c_int, C.POINTER(Cloud), C.POINTER(Cloud), C.c_void_p,
                                  C.POINTER(FinderParams), C.c_void_p, C.c_void_p]
        _lib.orc_linearize.argtypes = [C.c_int, C.c_int, C.POINTER(Cloud), C.POINTER(Cloud), C.c_void_p,
                                       C.c_void_p, C.POINTER(FinderParams), C.POINTER(FactorParams),
                                       C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.POINTER(IterStats), C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.orc_icp_run.argtypes = [C.c_int, C.c_int, C.POINTER(Slice), C.POINTER(AlignerParams),
                                     C.c_void_p, C.POINTER(IterStats), C.POINTER(C.c_int32),
                                     C.POINTER(C.c_int32), C.POINTER(CorrOut), C.c_int]
        _lib.orc_scales.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, C.POINTER(FinderParams),
                                    C.POINTER(FactorParams), C.POINTER(Scales)]
        _lib.orc_coord_bound.restype = C.c_float
        _lib.orc_coord_bound.argtypes = [C.c_int, C.POINTER(Cloud)]
        _lib.orc_radius_bound2.restype = C.c_float
        _lib.orc_radius_bound2.argtypes = [C.c_int, C.POINTER(Cloud)]
        _lib.orc_normal_bound2.restype = C.c_float
        _lib.orc_normal_bound2.argtypes = [C.c_int, C.POINTER(Cloud)]
        _lib.orc_sincos.argtypes = [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        _lib.orc_atan2.restype = C.c_double
        _lib.orc_atan2.argtypes = [C.c_double, C.c_double]
        _lib.orc_log.restype = C.c_double
        _lib.orc_log.argtypes = [C.c_double]
        _lib.orc_solve_update.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.orc_scene_merge.restype = C.c_int64
        _lib.orc_scene_merge.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(Cloud), C.c_void_p, C.c_int64,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int64,
                                         C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        _lib.orc_scene_clip.restype = C.c_int64
        _lib.orc_scene_clip.argtypes = [C.c_int, C.POINTER(Cloud), C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    return _lib


def set_threads(n):
    return lib().orc_set_threads(int(n))


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a):
    return None if a is None else a.ctypes.data


class CloudRef:
    """Keeps numpy buffers alive next to the C struct."""

    def __init__(self, coords, normals=None, valid=None):
        self.coords = _f32(coords)
        self.normals = _f32(normals)
        self.valid = None if valid is None else np.ascontiguousarray(valid, dtype=np.uint8)
        self.n = int(self.coords.shape[0])
        self.dim = int(self.coords.shape[1])
        self.c = Cloud(_ptr(self.coords), _ptr(self.normals), _ptr(self.valid), self.n)


def finder_params(max_distance=0.5, normal_cos=0.8, kind=FINDER_NN, fx=0, fy=0, cx=0, cy=0, width=0,
                  height=0, min_depth=0.0, max_depth=1e9):
    return FinderParams(kind, max_distance, normal_cos, fx, fy, cx, cy, width, height, min_depth, max_depth)


def factor_params(factor=FACTOR_PLANE, robustifier=ROB_NONE, chi_threshold=1.0, info_point=1.0,
                  info_normal=1.0):
    return FactorParams(factor, robustifier, chi_threshold, info_point, info_normal)


def aligner_params(variable=VAR_SE3_QUAT_RIGHT, max_iterations=10, min_num_inliers=10,
                   enable_inlier_only_runs=False, keep_only_inlier_correspondences=False,
                   use_termination_criteria=False, window_size=5, num_correspondences_range=20,
                   num_inliers_range=20, num_outliers_range=20, chi_epsilon=0.2):
    return AlignerParams(variable, max_iterations, min_num_inliers, int(enable_inlier_only_runs),
                         int(keep_only_inlier_correspondences), int(use_termination_criteria), window_size,
                         num_correspondences_range, num_inliers_range, num_outliers_range, chi_epsilon)


class Index:
    def __init__(self, fixed: CloudRef, method=NN_KDTREE):
        self.fixed = fixed
        self.h = lib().orc_index_create(fixed.dim, C.byref(fixed.c), method)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_index_free(self.h)
            self.h = None


def find(index, fixed, moving, S, fp):
    """Dense result: fixed_idx[j] (-1 = none), response[j]."""
    S = _f32(S).reshape(-1)
    fidx = np.empty(moving.n, dtype=np.int32)
    resp = np.empty(moving.n, dtype=np.float32)
    rc = lib().orc_find(index.h if index is not None else None, fixed.dim, C.byref(fixed.c),
                        C.byref(moving.c), S.ctypes.data, C.byref(fp), fidx.ctypes.data, resp.ctypes.data)
    assert rc == 0
    return fidx, resp


def linearize(fixed, moving, fidx, S, fp, fa, variable=VAR_SE3_QUAT_RIGHT, want_status=True,
              radius_bound2=0.0, normal_bound2=0.0, want_plain=False):
    """plain (want_plain): un-quantised fp64 sums of the same fp32 terms -- H upper triangle (21 | 6), b (6 | 3)
    packed one after the other, chi_inliers at [27], chi_outliers at [28]."""
    dim = fixed.dim
    P = 6 if dim == 3 else 3
    S = _f32(S).reshape(-1)
    fidx = np.ascontiguousarray(fidx, dtype=np.int32)
    acc = np.zeros(ACC_SLOTS, dtype=np.int64)
    H = np.zeros((P, P), dtype=np.float64)
    b = np.zeros(P, dtype=np.float64)
    st = IterStats()
    status = np.empty(moving.n, dtype=np.uint8) if want_status else None
    chi = np.empty(moving.n, dtype=np.float32) if want_status else None
    plain = np.zeros(32, dtype=np.float64) if want_plain else None
    rc = lib().orc_linearize(dim, variable, C.byref(fixed.c), C.byref(moving.c), fidx.ctypes.data,
                             S.ctypes.data, C.byref(fp), C.byref(fa), radius_bound2, normal_bound2,
                             acc.ctypes.data, H.ctypes.data, b.ctypes.data, C.byref(st), _ptr(status), _ptr(chi),
                             _ptr(plain))
    assert rc == 0
    return dict(acc=acc, H=H, b=b, stats=st.as_dict(), status=status, chi=chi, plain=plain)


def make_slice(fixed=None, moving=None, robot_in_sensor=None, fp=None, fa=None, min_num_correspondences=0,
               prior_measurement=None, prior_info_diag=None, dim=3):
    s = Slice()
    D1 = dim + 1
    eye = np.eye(D1, dtype=np.float32).reshape(-1)
    if prior_measurement is not None:
        s.kind = SLICE_PRIOR
        z = _f32(prior_measurement).reshape(-1)
        for i in range(D1 * D1):
            s.prior_measurement[i] = z[i]
        info = _f32(prior_info_diag).reshape(-1)
        for i in range(len(info)):
            s.prior_info_diag[i] = info[i]
        for i in range(D1 * D1):
            s.robot_in_sensor[i] = eye[i]
        return s
    s.kind = SLICE_POINTS
    s.min_num_correspondences = min_num_correspondences
    s.fixed = fixed.c
    s.moving = moving.c
    r = eye if robot_in_sensor is None else _f32(robot_in_sensor).reshape(-1)
    for i in range(D1 * D1):
        s.robot_in_sensor[i] = r[i]
    s.finder = fp
    s.factor = fa
    return s


def icp_run(dim, slices, ap, T0, nn_method=NN_KDTREE, want_correspondences=True, max_stats=None):
    n = len(slices)
    arr = (Slice * n)(*slices)
    D1 = dim + 1
    T = np.ascontiguousarray(np.asarray(T0, dtype=np.float32).reshape(D1, D1)).copy()
    cap = max_stats or (2 * ap.max_iterations + 2)
    stats = (IterStats * cap)()
    n_stats = C.c_int32(cap)
    status = C.c_int32(-1)
    outs = (CorrOut * n)()
    keep = []
    if want_correspondences:
        for i, s in enumerate(slices):
            if s.kind == SLICE_POINTS:
                nm = s.moving.n
                fi = np.empty(nm, dtype=np.int32)
                mi = np.empty(nm, dtype=np.int32)
                rs = np.empty(nm, dtype=np.float32)
                keep.append((fi, mi, rs))
                outs[i] = CorrOut(fi.ctypes.data, mi.ctypes.data, rs.ctypes.data, 0)
            else:
                keep.append(None)
    rc = lib().orc_icp_run(dim, n, arr, C.byref(ap), T.ctypes.data, stats, C.byref(n_stats), C.byref(status),
                           outs if want_correspondences else None, nn_method)
    assert rc == 0
    corr = []
    if want_correspondences:
        for i, k in enumerate(keep):
            if k is None:
                corr.append(None)
            else:
                m = outs[i].n
                corr.append((k[0][:m].copy(), k[1][:m].copy(), k[2][:m].copy()))
    return dict(T=T, status=status.value, stats=[stats[i].as_dict() for i in range(min(n_stats.value, cap))],
                correspondences=corr)


CLOSURE_ACCEPT, CLOSURE_ALIGNER_DROP, CLOSURE_NUM_INLIERS_DROP, CLOSURE_MAX_CHI_DROP, CLOSURE_INLIER_RATIO_DROP = 0, 1, 2, 3, 4


def closure_loop(dim, candidate_slices, ap, guesses, min_inliers=500, max_chi_inliers=0.005, min_inliers_ratio=0.7):
    """MultiLoopDetectorBruteForce_::compute's candidate loop
    (R/registration/loop_detector/multi_loop_detector_brute_force_impl.cpp:63-133; the same gates in
    R/registration/relocalization/multi_relocalizer_impl.cpp:86-118), restated over the oracle aligner: one compute()
    per candidate, serially, then -- in this order -- status, num_inliers, chi per inlier, inlier ratio.
    candidate_slices[k]: the slice list of candidate k (same configuration, candidate k's moving cloud)."""
    out = []
    for slices, g in zip(candidate_slices, guesses):
        r = icp_run(dim, slices, ap, g)                                   # :75-78 setMoving / setMovingInFixed / compute
        res = dict(aligner_status=r["status"], iterations=len(r["stats"]), T=r["T"], num_correspondences=0,
                   num_inliers=0, chi_inliers=np.float32(0))
        out.append(res)
        if r["status"] != 0:                                               # :80-84
            res["verdict"] = CLOSURE_ALIGNER_DROP
            continue
        istat = r["stats"][-1]                                             # :86
        ncorr = 0                                                          # multi_aligner_impl.cpp:275-285
        for s, c in zip(slices, r["correspondences"]):
            ncorr += 1 if c is None else len(c[0])                         # a prior slice counts 1
        ninl = istat["num_inliers"]
        chi = np.float32(istat["chi_inliers"]) / np.float32(ninl)          # :91
        res.update(num_correspondences=ncorr, num_inliers=ninl, chi_inliers=np.float32(chi))
        if ninl < min_inliers:                                             # :94
            res["verdict"] = CLOSURE_NUM_INLIERS_DROP
        elif chi > np.float32(max_chi_inliers):                            # :99
            res["verdict"] = CLOSURE_MAX_CHI_DROP
        elif np.float32(ninl) / np.float32(ncorr) < np.float32(min_inliers_ratio):  # :105
            res["verdict"] = CLOSURE_INLIER_RATIO_DROP
        else:
            res["verdict"] = CLOSURE_ACCEPT
    return out


def relocalize_loop(dim, candidate_slices, ap, guesses, translations, max_translation, min_inliers=500, max_chi_inliers=0.005,
                    min_inliers_ratio=0.7):
    """MultiRelocalizer_::compute, alignment branch (R/registration/relocalization/multi_relocalizer_impl.cpp:74-138):
    the detector's loop with a translation pre-filter (:79-83) and `best = smallest chi per inlier, first on ties` (:121-131).
    Returns (index of the relocalisation map or None, per-candidate results with None for the pre-filtered ones)."""
    out, best, best_chi = [], None, np.float32(np.finfo(np.float32).max)
    for k, (slices, g) in enumerate(zip(candidate_slices, guesses)):
        if translations[k] > max_translation:                              # :79-83
            out.append(None)
            continue
        r = closure_loop(dim, [slices], ap, [g], min_inliers, max_chi_inliers, min_inliers_ratio)[0]
        out.append(r)
        if r["verdict"] == CLOSURE_ACCEPT and r["chi_inliers"] < best_chi:  # :119-121
            best, best_chi = k, r["chi_inliers"]
    return best, out



    s = Scales()
    lib().orc_scales(dim, variable, radius_bound2, normal_bound2, C.byref(fp), C.byref(fa), C.byref(s))
    return tuple(s.k)


def radius_bound2(cloud):
    return float(lib().orc_radius_bound2(cloud.dim, C.byref(cloud.c)))


def normal_bound2(cloud):
    return float(lib().orc_normal_bound2(cloud.dim, C.byref(cloud.c)))


def solve_update(dim, variable, H, b, T):
    D1 = dim + 1
    H = np.ascontiguousarray(H, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    T = np.ascontiguousarray(np.asarray(T, dtype=np.float32).reshape(D1, D1)).copy()
    ok = lib().orc_solve_update(dim, variable, H.ctypes.data, b.ctypes.data, T.ctypes.data)
    return ok, T


def mat_op(name, dim, *mats):
    D1 = dim + 1
    ms = [np.ascontiguousarray(np.asarray(m, dtype=np.float32).reshape(D1, D1)) for m in mats]
    out = np.zeros((D1, D1), dtype=np.float32)
    f = getattr(lib(), name)
    f.argtypes = [C.c_int] + [C.c_void_p] * (len(ms) + 1)
    f(dim, *[m.ctypes.data for m in ms], out.ctypes.data)
    return out


def scene_clip(scene, T, max_range):
    """Range clip of a resident scene: (coords, normals | None, global_indices) of the kept points."""
    dim = scene.dim
    D1 = dim + 1
    T = np.ascontiguousarray(np.asarray(T, dtype=np.float32).reshape(D1, D1))
    oc = np.empty((scene.n, dim), dtype=np.float32)
    on = np.empty((scene.n, dim), dtype=np.float32) if scene.normals is not None else None
    gi = np.empty(scene.n, dtype=np.int32)
    k = lib().orc_scene_clip(dim, C.byref(scene.c), T.ctypes.data, float(max_range), oc.ctypes.data, _ptr(on), gi.ctypes.data)
    return oc[:k].copy(), (None if on is None else on[:k].copy()), gi[:k].copy()


def scene_merge(scene_coords, scene_normals, scene_valid, meas, T, corr=None, maximum_response=50.0,
                maximum_distance_geometry_squared=0.25, target_number_of_merges=200):
    """MergerCorrespondenceHomo_::compute.  corr = (scene_idx, meas_idx, response) arrays or None (no
    correspondences set).  Returns (coords, normals | None, valid | None, n_merged, n_added) of the merged scene."""
    dim = meas.dim
    D1 = dim + 1
    T = np.ascontiguousarray(np.asarray(T, dtype=np.float32).reshape(D1, D1))
    ns = scene_coords.shape[0]
    cap = ns + meas.n
    sc = np.zeros((cap, dim), np.float32); sc[:ns] = scene_coords
    sn = None
    if scene_normals is not None:
        sn = np.zeros((cap, dim), np.float32); sn[:ns] = scene_normals
    sv = None
    if scene_valid is not None:
        sv = np.zeros(cap, np.uint8); sv[:ns] = scene_valid
    nm, na = C.c_int64(0), C.c_int64(0)
    if corr is None:
        n = lib().orc_scene_merge(dim, sc.ctypes.data, _ptr(sn), _ptr(sv), ns, C.byref(meas.c), T.ctypes.data, -1, None, None, None,
                                  maximum_response, maximum_distance_geometry_squared, target_number_of_merges, C.byref(nm), C.byref(na))
    else:
        cs = np.ascontiguousarray(corr[0], np.int32); cm = np.ascontiguousarray(corr[1], np.int32); cr = np.ascontiguousarray(corr[2], np.float32)
        n = lib().orc_scene_merge(dim, sc.ctypes.data, _ptr(sn), _ptr(sv), ns, C.byref(meas.c), T.ctypes.data, cs.shape[0], cs.ctypes.data,
                                  cm.ctypes.data, cr.ctypes.data, maximum_response, maximum_distance_geometry_squared,
                                  target_number_of_merges, C.byref(nm), C.byref(na))
    return sc[:n], (None if sn is None else sn[:n]), (None if sv is None else sv[:n]), nm.value, na.value
