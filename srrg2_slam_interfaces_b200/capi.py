"""ctypes binding of libsrrg2b.so (include/srrg2b.h).  This is the product path: it fails loudly
when the CUDA library is missing or no GPU is present -- there is no CPU fallback."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SRRG2B_LIB: an alternative build of the same library (tuning experiments); never a fallback
LIB_PATH = os.environ.get("SRRG2B_LIB") or os.path.join(_HERE, "libsrrg2b.so")

OK, ERR_INVALID, ERR_CUDA, ERR_STATE, ERR_NCCL = 0, 1, 2, 3, 4
FIXED, MOVING = 0, 1
FACTOR_P2P, FACTOR_PLANE = 0, 1
ROB_NONE, ROB_SATURATED, ROB_CAUCHY, ROB_CLAMP, ROB_HUBER = 0, 1, 2, 3, 4
VAR_SE3_QUAT_RIGHT, VAR_SE3_EULER_RIGHT = 0, 1
FINDER_NN, FINDER_PROJECTIVE = 0, 1
SLICE_POINTS, SLICE_PRIOR = 0, 1
ALIGNER_SUCCESS, ALIGNER_NOT_ENOUGH_CORRESPONDENCES, ALIGNER_NOT_ENOUGH_INLIERS, ALIGNER_FAIL = 0, 1, 2, 3
STAT_INLIER, STAT_KERNELIZED, STAT_SUPPRESSED, STAT_NONE = 0, 1, 2, 3
MAX_SLICES = 8

EXPORTED_SYMBOLS = [
    "srrg2b_version", "srrg2b_ctx_create", "srrg2b_ctx_destroy", "srrg2b_last_error", "srrg2b_stream",
    "srrg2b_launch_count", "srrg2b_comm_unique_id", "srrg2b_comm_init", "srrg2b_set_cloud", "srrg2b_scene_set", "srrg2b_scene_clip", "srrg2b_scene_clip_indices", "srrg2b_scene_merge", "srrg2b_scene_get",
    "srrg2b_find_correspondences", "srrg2b_set_correspondences", "srrg2b_linearize", "srrg2b_icp_run",
    "srrg2b_icp_iterate", "srrg2b_get_correspondences", "srrg2b_reset_correspondences", "srrg2b_last_run_timing",
    "srrg2b_set_kernel_timing", "srrg2b_last_kernel_timing", "srrg2b_debug_info",
    "srrg2b_share_fixed", "srrg2b_closure_batch",
    "srrg2b_pgo_upload", "srrg2b_pgo_iterate", "srrg2b_pgo_optimize", "srrg2b_pgo_download",
]


class Srrg2bError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("srrg2b error %d: %s" % (code, msg))
        self.code = code


class Cloud(C.Structure):
    _fields_ = [("coords", C.c_void_p), ("normals", C.c_void_p), ("valid", C.c_void_p), ("n", C.c_int64),
                ("index_offset", C.c_int64), ("n_global", C.c_int64), ("on_device", C.c_int32),
                ("reserved", C.c_int32)]


class FinderParams(C.Structure):
    _fields_ = [("kind", C.c_int32), ("max_distance", C.c_float), ("normal_cos", C.c_float),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("width", C.c_int32), ("height", C.c_int32), ("min_depth", C.c_float), ("max_depth", C.c_float)]


class FactorParams(C.Structure):
    _fields_ = [("factor", C.c_int32), ("robustifier", C.c_int32), ("chi_threshold", C.c_float),
                ("info_point", C.c_float), ("info_normal", C.c_float)]


class Slice(C.Structure):
    _fields_ = [("kind", C.c_int32), ("slice_id", C.c_int32), ("min_num_correspondences", C.c_int32),
                ("reserved", C.c_int32), ("robot_in_sensor", C.c_float * 16), ("finder", FinderParams),
                ("factor", FactorParams), ("prior_measurement", C.c_float * 16),
                ("prior_info_diag", C.c_float * 6)]


class IterStats(C.Structure):
    _fields_ = [("iteration", C.c_int32), ("solver_status", C.c_int32), ("num_inliers", C.c_int64),
                ("num_outliers", C.c_int64), ("num_suppressed", C.c_int64), ("num_correspondences", C.c_int64),
                ("chi_inliers", C.c_double), ("chi_outliers", C.c_double), ("num_saturated", C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class PgoStats(C.Structure):
    _fields_ = [("chi", C.c_double), ("chi_after", C.c_double), ("dx_norm_inf", C.c_double),
                ("cg_relative_residual", C.c_double), ("lambda", C.c_double), ("gain_ratio", C.c_double),
                ("cg_iterations", C.c_int32), ("num_factors", C.c_int32), ("num_blocks", C.c_int32),
                ("accepted", C.c_int32), ("linearize_ms", C.c_float), ("solve_ms", C.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class MergeParams(C.Structure):
    _fields_ = [("maximum_response", C.c_float), ("maximum_distance_geometry_squared", C.c_float),
                ("target_number_of_merges", C.c_int32), ("without_correspondences", C.c_int32)]


class ClosureParams(C.Structure):
    _fields_ = [("relocalize_min_inliers", C.c_int32), ("relocalize_max_chi_inliers", C.c_float),
                ("relocalize_min_inliers_ratio", C.c_float)]


class ClosureResult(C.Structure):
    _fields_ = [("verdict", C.c_int32), ("aligner_status", C.c_int32), ("iterations", C.c_int32),
                ("reserved", C.c_int32), ("num_correspondences", C.c_int64), ("num_inliers", C.c_int64),
                ("chi_inliers", C.c_float), ("device_ms", C.c_float), ("moving_in_fixed", C.c_float * 16)]


CLOSURE_ACCEPT, CLOSURE_ALIGNER_DROP, CLOSURE_NUM_INLIERS_DROP, CLOSURE_MAX_CHI_DROP, CLOSURE_INLIER_RATIO_DROP = 0, 1, 2, 3, 4


class AlignerParams(C.Structure):
    _fields_ = [("variable", C.c_int32), ("max_iterations", C.c_int32), ("min_num_inliers", C.c_int32),
                ("enable_inlier_only_runs", C.c_int32), ("keep_only_inlier_correspondences", C.c_int32),
                ("use_termination_criteria", C.c_int32), ("window_size", C.c_int32),
                ("num_correspondences_range", C.c_int32), ("num_inliers_range", C.c_int32),
                ("num_outliers_range", C.c_int32), ("chi_epsilon", C.c_float)]


_lib = None


def load_library():
    """Loads libsrrg2b.so; raises if it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Srrg2bError(ERR_CUDA, "libsrrg2b.so is not built (%s); run __graft_entry__.build()" % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, i32p, i64p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    lib.srrg2b_version.restype = C.c_int
    lib.srrg2b_ctx_create.argtypes = [C.c_int, C.c_int, C.POINTER(vp)]
    lib.srrg2b_ctx_destroy.argtypes = [vp]
    lib.srrg2b_last_error.restype = C.c_char_p
    lib.srrg2b_last_error.argtypes = [vp]
    lib.srrg2b_stream.restype = vp
    lib.srrg2b_stream.argtypes = [vp]
    lib.srrg2b_launch_count.restype = C.c_int64
    lib.srrg2b_launch_count.argtypes = [vp]
    lib.srrg2b_comm_unique_id.argtypes = [vp]
    lib.srrg2b_comm_init.argtypes = [vp, vp, C.c_int, C.c_int]
    lib.srrg2b_set_cloud.argtypes = [vp, C.c_int, C.c_int, C.POINTER(Cloud)]
    lib.srrg2b_find_correspondences.argtypes = [vp, C.c_int, vp, C.POINTER(FinderParams), vp, vp, vp, i64p]
    lib.srrg2b_set_correspondences.argtypes = [vp, C.c_int, vp, vp, C.c_int64]
    lib.srrg2b_linearize.argtypes = [vp, C.c_int, vp, C.c_int, C.POINTER(FinderParams), C.POINTER(FactorParams),
                                     vp, vp, vp, C.POINTER(IterStats), vp, vp]
    lib.srrg2b_icp_run.argtypes = [vp, C.c_int, C.POINTER(Slice), C.POINTER(AlignerParams), vp,
                                   C.POINTER(IterStats), i32p, i32p]
    lib.srrg2b_icp_iterate.argtypes = [vp, C.c_int, C.POINTER(Slice), C.c_int, vp, C.POINTER(IterStats), i32p]
    lib.srrg2b_get_correspondences.argtypes = [vp, C.c_int, vp, vp, vp, i64p]
    lib.srrg2b_reset_correspondences.argtypes = [vp, C.c_int]
    lib.srrg2b_scene_set.argtypes = [vp, C.c_int, C.POINTER(Cloud)]
    lib.srrg2b_scene_clip.argtypes = [vp, C.c_int, C.c_int, vp, C.c_float, i64p]
    lib.srrg2b_scene_clip_indices.argtypes = [vp, C.c_int, vp]
    lib.srrg2b_scene_merge.argtypes = [vp, C.c_int, C.c_int, vp, C.POINTER(MergeParams), i64p, i64p]
    lib.srrg2b_scene_get.argtypes = [vp, C.c_int, vp, vp, vp, i64p]
    lib.srrg2b_share_fixed.argtypes = [vp, C.c_int, vp, C.c_int]
    lib.srrg2b_closure_batch.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.POINTER(Slice), C.POINTER(AlignerParams), vp,
                                         C.POINTER(ClosureParams), C.POINTER(ClosureResult)]
    lib.srrg2b_last_run_timing.argtypes = [vp, C.POINTER(C.c_float), i32p]
    lib.srrg2b_debug_info.argtypes = [vp, C.c_int, vp]
    lib.srrg2b_pgo_upload.argtypes = [vp, C.c_int64, vp, vp, C.c_int64, vp, vp, vp]
    lib.srrg2b_pgo_iterate.argtypes = [vp, C.c_int, C.c_double, C.POINTER(PgoStats)]
    lib.srrg2b_pgo_optimize.argtypes = [vp, C.c_int, C.c_double, C.c_int, C.POINTER(PgoStats), i32p]
    lib.srrg2b_pgo_download.argtypes = [vp, vp]
    lib.srrg2b_set_kernel_timing.argtypes = [vp, C.c_int]
    lib.srrg2b_last_kernel_timing.argtypes = [vp, C.POINTER(C.c_float), i32p]
    _lib = lib
    return lib


def finder_params(max_distance=0.5, normal_cos=0.8, kind=FINDER_NN, fx=0, fy=0, cx=0, cy=0, width=0, height=0,
                  min_depth=0.0, max_depth=1e9):
    return FinderParams(kind, max_distance, normal_cos, fx, fy, cx, cy, width, height, min_depth, max_depth)


def factor_params(factor=FACTOR_PLANE, robustifier=ROB_NONE, chi_threshold=1.0, info_point=1.0, info_normal=1.0):
    return FactorParams(factor, robustifier, chi_threshold, info_point, info_normal)


def aligner_params(variable=VAR_SE3_QUAT_RIGHT, max_iterations=10, min_num_inliers=10,
                   enable_inlier_only_runs=False, keep_only_inlier_correspondences=False,
                   use_termination_criteria=False, window_size=5, num_correspondences_range=20,
                   num_inliers_range=20, num_outliers_range=20, chi_epsilon=0.2):
    return AlignerParams(variable, max_iterations, min_num_inliers, int(enable_inlier_only_runs),
                         int(keep_only_inlier_correspondences), int(use_termination_criteria), window_size,
                         num_correspondences_range, num_inliers_range, num_outliers_range, chi_epsilon)


def make_slice(dim, slice_id=0, robot_in_sensor=None, fp=None, fa=None, min_num_correspondences=0,
               prior_measurement=None, prior_info_diag=None):
    s = Slice()
    D1 = dim + 1
    eye = np.eye(D1, dtype=np.float32).reshape(-1)
    r = eye if robot_in_sensor is None else np.asarray(robot_in_sensor, dtype=np.float32).reshape(-1)
    for i in range(D1 * D1):
        s.robot_in_sensor[i] = r[i]
    if prior_measurement is not None:
        s.kind = SLICE_PRIOR
        z = np.asarray(prior_measurement, dtype=np.float32).reshape(-1)
        for i in range(D1 * D1):
            s.prior_measurement[i] = z[i]
        info = np.asarray(prior_info_diag, dtype=np.float32).reshape(-1)
        for i in range(len(info)):
            s.prior_info_diag[i] = info[i]
        return s
    s.kind = SLICE_POINTS
    s.slice_id = slice_id
    s.min_num_correspondences = min_num_correspondences
    s.finder = fp
    s.factor = fa
    return s


def closure_params(relocalize_min_inliers=500, relocalize_max_chi_inliers=0.005, relocalize_min_inliers_ratio=0.7):
    """Defaults of MultiLoopDetectorBruteForce_ (multi_loop_detector_brute_force.h:25-40)."""
    return ClosureParams(relocalize_min_inliers, relocalize_max_chi_inliers, relocalize_min_inliers_ratio)


def closure_batch(contexts, slices, ap, guesses, cp):
    """srrg2b_closure_batch: the detector's aligner over K candidate contexts, all runs in flight together."""
    k = len(contexts)
    if k == 0:
        return []
    dim = contexts[0].dim
    D1 = dim + 1
    lib = contexts[0].lib
    hs = (C.c_void_p * k)(*[c.h for c in contexts])
    n = len(slices)
    arr = (Slice * n)(*slices)
    g = np.ascontiguousarray(np.asarray(guesses, dtype=np.float32).reshape(k, D1 * D1))
    res = (ClosureResult * k)()
    contexts[0]._check(lib.srrg2b_closure_batch(hs, k, n, arr, C.byref(ap), g.ctypes.data, C.byref(cp), res))
    out = []
    for r in res:
        out.append(dict(verdict=r.verdict, aligner_status=r.aligner_status, iterations=r.iterations,
                        num_correspondences=r.num_correspondences, num_inliers=r.num_inliers,
                        chi_inliers=np.float32(r.chi_inliers), device_ms=r.device_ms,
                        T=np.array(r.moving_in_fixed[:D1 * D1], dtype=np.float32).reshape(D1, D1)))
    return out


def relocalize_select(results, pose_in_target_translations=None, max_translation=None):
    """MultiRelocalizer_::compute's choice among the candidates (R/registration/relocalization/
    multi_relocalizer_impl.cpp:77-131): candidates whose guess lies farther than param_max_translation are not aligned
    at all (:79-83; pass their translation norms and the bound, their results are ignored), the others go through the
    detector's gates (closure_batch did that), and the accepted candidate with the SMALLEST chi per inlier wins --
    strictly smaller, so the first one on ties (:121).  Returns its index, or None."""
    best, best_chi = None, float("inf")
    for k, r in enumerate(results):
        if max_translation is not None and pose_in_target_translations is not None and pose_in_target_translations[k] > max_translation:
            continue
        if r["verdict"] != CLOSURE_ACCEPT:
            continue
        if float(r["chi_inliers"]) < best_chi:
            best, best_chi = k, float(r["chi_inliers"])
    return best


def _ptr(a):
    return None if a is None else a.ctypes.data


class Context:
    """One srrg2b_ctx: one GPU, one CUDA stream, device-resident clouds."""

    def __init__(self, dim, device=0):
        self.lib = load_library()
        self.dim = dim
        h = C.c_void_p()
        rc = self.lib.srrg2b_ctx_create(dim, device, C.byref(h))
        if rc != OK:
            raise Srrg2bError(rc, "srrg2b_ctx_create failed (no CUDA device? dim=%d device=%d)" % (dim, device))
        self.h = h
        self._keep = {}

    def close(self):
        if getattr(self, "h", None):
            self.lib.srrg2b_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != OK:
            raise Srrg2bError(rc, self.lib.srrg2b_last_error(self.h).decode())

    @property
    def stream(self):
        return self.lib.srrg2b_stream(self.h)

    @property
    def launch_count(self):
        return self.lib.srrg2b_launch_count(self.h)

    # ---- multi GPU ----
    def unique_id(self):
        buf = (C.c_char * 128)()
        rc = self.lib.srrg2b_comm_unique_id(buf)
        if rc != OK:
            raise Srrg2bError(rc, "ncclGetUniqueId failed")
        return bytes(buf)

    def comm_init(self, uid, rank, world):
        buf = (C.c_char * 128).from_buffer_copy(uid)
        self._check(self.lib.srrg2b_comm_init(self.h, buf, rank, world))

    # ---- data ----
    def set_cloud(self, slot, slice_id, coords, normals=None, valid=None, index_offset=0, n_global=0):
        coords = np.ascontiguousarray(coords, dtype=np.float32)
        if coords.ndim != 2 or coords.shape[1] != self.dim:
            raise Srrg2bError(ERR_INVALID, "coords must be n x %d" % self.dim)
        normals = None if normals is None else np.ascontiguousarray(normals, dtype=np.float32)
        valid = None if valid is None else np.ascontiguousarray(valid, dtype=np.uint8)
        cl = Cloud(_ptr(coords), _ptr(normals), _ptr(valid), coords.shape[0], index_offset, n_global, 0, 0)
        self._check(self.lib.srrg2b_set_cloud(self.h, slot, slice_id, C.byref(cl)))

    # ---- N1: device-resident local map, clipped on the device ----
    def scene_set(self, scene_id, coords, normals=None, valid=None):
        coords = np.ascontiguousarray(coords, dtype=np.float32)
        normals = None if normals is None else np.ascontiguousarray(normals, dtype=np.float32)
        valid = None if valid is None else np.ascontiguousarray(valid, dtype=np.uint8)
        cl = Cloud(_ptr(coords), _ptr(normals), _ptr(valid), coords.shape[0], 0, 0, 0, 0)
        self._check(self.lib.srrg2b_scene_set(self.h, scene_id, C.byref(cl)))

    def scene_clip(self, scene_id, slice_id, scene_in_robot, max_range):
        """Clips the resident scene into the MOVING cloud of slice_id; returns the number of points kept."""
        T = np.ascontiguousarray(np.asarray(scene_in_robot, dtype=np.float32).reshape(-1))
        n = C.c_int64(0)
        self._check(self.lib.srrg2b_scene_clip(self.h, scene_id, slice_id, T.ctypes.data, float(max_range), C.byref(n)))
        self._clip_n = getattr(self, "_clip_n", {})
        self._clip_n[slice_id] = n.value
        return n.value

    def scene_clip_indices(self, slice_id):
        out = np.empty(self._clip_n[slice_id], dtype=np.int32)
        self._check(self.lib.srrg2b_scene_clip_indices(self.h, slice_id, out.ctypes.data))
        return out

    def scene_merge(self, scene_id, slice_id, measurement_in_scene, maximum_response=50.0, maximum_distance_geometry_squared=0.25,
                    target_number_of_merges=200, without_correspondences=False):
        """MergerCorrespondenceHomo_::compute on the resident scene; returns (n_merged, n_added)."""
        T = np.ascontiguousarray(np.asarray(measurement_in_scene, dtype=np.float32).reshape(-1))
        mp = MergeParams(maximum_response, maximum_distance_geometry_squared, target_number_of_merges, int(without_correspondences))
        a, b = C.c_int64(0), C.c_int64(0)
        self._check(self.lib.srrg2b_scene_merge(self.h, scene_id, slice_id, T.ctypes.data, C.byref(mp), C.byref(a), C.byref(b)))
        return a.value, b.value

    def scene_get(self, scene_id, normals=True, valid=False):
        n = C.c_int64(0)
        self._check(self.lib.srrg2b_scene_get(self.h, scene_id, None, None, None, C.byref(n)))
        co = np.empty((n.value, self.dim), np.float32)
        no = np.empty((n.value, self.dim), np.float32) if normals else None
        va = np.empty(n.value, np.uint8) if valid else None
        self._check(self.lib.srrg2b_scene_get(self.h, scene_id, co.ctypes.data, _ptr(no), _ptr(va), C.byref(n)))
        return co, no, va

    def set_cloud_device(self, slot, slice_id, coords_ptr, normals_ptr, valid_ptr, n, index_offset=0, n_global=0):
        cl = Cloud(coords_ptr, normals_ptr, valid_ptr, n, index_offset, n_global, 1, 0)
        self._check(self.lib.srrg2b_set_cloud(self.h, slot, slice_id, C.byref(cl)))

    # ---- a3 ----
    def find_correspondences(self, slice_id, S, fp, n_moving):
        S = np.ascontiguousarray(np.asarray(S, dtype=np.float32).reshape(-1))
        fi = np.empty(n_moving, dtype=np.int32)
        mi = np.empty(n_moving, dtype=np.int32)
        rs = np.empty(n_moving, dtype=np.float32)
        n = C.c_int64(0)
        self._check(self.lib.srrg2b_find_correspondences(self.h, slice_id, S.ctypes.data, C.byref(fp), fi.ctypes.data,
                                                         mi.ctypes.data, rs.ctypes.data, C.byref(n)))
        return fi[:n.value], mi[:n.value], rs[:n.value]

    def set_correspondences(self, slice_id, fixed_idx, moving_idx):
        fi = np.ascontiguousarray(fixed_idx, dtype=np.int32)
        mi = np.ascontiguousarray(moving_idx, dtype=np.int32)
        self._check(self.lib.srrg2b_set_correspondences(self.h, slice_id, fi.ctypes.data, mi.ctypes.data, fi.size))

    # ---- a5 ----
    def linearize(self, slice_id, S, fp, fa, variable=VAR_SE3_QUAT_RIGHT, n_moving=0, want_status=True):
        P = 6 if self.dim == 3 else 3
        S = np.ascontiguousarray(np.asarray(S, dtype=np.float32).reshape(-1))
        H = np.zeros((P, P), dtype=np.float64)
        b = np.zeros(P, dtype=np.float64)
        acc = np.zeros(40, dtype=np.int64)
        st = IterStats()
        status = np.empty(n_moving, dtype=np.uint8) if want_status else None
        chi = np.empty(n_moving, dtype=np.float32) if want_status else None
        self._check(self.lib.srrg2b_linearize(self.h, slice_id, S.ctypes.data, variable, C.byref(fp), C.byref(fa),
                                              H.ctypes.data, b.ctypes.data, acc.ctypes.data, C.byref(st),
                                              _ptr(status), _ptr(chi)))
        n = st.num_correspondences
        return dict(H=H, b=b, acc=acc, stats=st.as_dict(), status=None if status is None else status[:n],
                    chi=None if chi is None else chi[:n])

    # ---- a1 ----
    def icp_run(self, slices, ap, T0):
        D1 = self.dim + 1
        n = len(slices)
        arr = (Slice * n)(*slices)
        T = np.ascontiguousarray(np.asarray(T0, dtype=np.float32).reshape(D1, D1)).copy()
        cap = 256
        stats = (IterStats * cap)()
        n_stats = C.c_int32(cap)
        status = C.c_int32(-1)
        self._check(self.lib.srrg2b_icp_run(self.h, n, arr, C.byref(ap), T.ctypes.data, stats, C.byref(n_stats),
                                            C.byref(status)))
        return dict(T=T, status=status.value, stats=[stats[i].as_dict() for i in range(min(n_stats.value, cap))])

    def share_fixed(self, slice_id, src, src_slice_id):
        """Borrow the fixed cloud and NN index of another context's slice (srrg2b_share_fixed): no copy."""
        self._check(self.lib.srrg2b_share_fixed(self.h, slice_id, src.h, src_slice_id))

    def icp_iterate(self, slices, variable, T):
        D1 = self.dim + 1
        n = len(slices)
        arr = (Slice * n)(*slices)
        T = np.ascontiguousarray(np.asarray(T, dtype=np.float32).reshape(D1, D1)).copy()
        st = IterStats()
        good = C.c_int32(0)
        self._check(self.lib.srrg2b_icp_iterate(self.h, n, arr, variable, T.ctypes.data, C.byref(st), C.byref(good)))
        return dict(T=T, stats=st.as_dict(), association_good=bool(good.value))

    def get_correspondences(self, slice_id, n_moving):
        fi = np.empty(n_moving, dtype=np.int32)
        mi = np.empty(n_moving, dtype=np.int32)
        rs = np.empty(n_moving, dtype=np.float32)
        n = C.c_int64(0)
        self._check(self.lib.srrg2b_get_correspondences(self.h, slice_id, fi.ctypes.data, mi.ctypes.data,
                                                        rs.ctypes.data, C.byref(n)))
        return fi[:n.value], mi[:n.value], rs[:n.value]

    def reset_correspondences(self, slice_id):
        """Forget the slice's correspondences / warm-start candidates / certified bounds (cold start)."""
        self._check(self.lib.srrg2b_reset_correspondences(self.h, slice_id))

    def debug_info(self, slice_id):
        out = np.zeros(16, dtype=np.int32)
        self._check(self.lib.srrg2b_debug_info(self.h, slice_id, out.ctypes.data))
        return dict(R=int(out[0]), dims=(int(out[1]), int(out[2]), int(out[3])), n_fixed_valid=int(out[4]),
                    n_moving_valid=int(out[5]), last_far_count=int(out[6]), last_work_count=int(out[8]),
                    cell=float(out[7:8].view(np.float32)[0]))

    def set_kernel_timing(self, enable):
        self._check(self.lib.srrg2b_set_kernel_timing(self.h, int(enable)))

    def last_kernel_timing(self):
        ms = C.c_float(0)
        n = C.c_int32(0)
        self._check(self.lib.srrg2b_last_kernel_timing(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    # ---- a10: pose graph ----
    def pgo_upload(self, poses, fixed, ij, Z, Omega):
        """dim 3: poses / Z n x 4 x 4, Omega n x 6 x 6; dim 2: poses / Z n x 3 x 3, Omega n x 3 x 3."""
        m, b = (16, 36) if self.dim == 3 else (9, 9)
        poses = np.ascontiguousarray(poses, dtype=np.float32).reshape(-1, m)
        fixed = np.ascontiguousarray(fixed, dtype=np.uint8)
        ij = np.ascontiguousarray(ij, dtype=np.int32).reshape(-1, 2)
        Z = np.ascontiguousarray(Z, dtype=np.float32).reshape(-1, m)
        Omega = np.ascontiguousarray(Omega, dtype=np.float32).reshape(-1, b)
        self._pgo_n = poses.shape[0]
        self._check(self.lib.srrg2b_pgo_upload(self.h, poses.shape[0], poses.ctypes.data, fixed.ctypes.data,
                                               ij.shape[0], ij.ctypes.data, Z.ctypes.data, Omega.ctypes.data))

    def pgo_iterate(self, max_cg_iterations=2000, cg_tolerance=1e-10):
        st = PgoStats()
        self._check(self.lib.srrg2b_pgo_iterate(self.h, max_cg_iterations, cg_tolerance, C.byref(st)))
        return st.as_dict()

    def pgo_optimize(self, max_iterations=10, dx_tolerance=1e-6, max_cg_iterations=2000):
        """Damped Gauss-Newton until |dx|_inf < dx_tolerance: list of per-iteration stats."""
        arr = (PgoStats * max_iterations)()
        n = C.c_int32(0)
        self._check(self.lib.srrg2b_pgo_optimize(self.h, max_iterations, dx_tolerance, max_cg_iterations, arr, C.byref(n)))
        return [arr[k].as_dict() for k in range(n.value)]

    def pgo_download(self):
        d1 = self.dim + 1
        out = np.empty((self._pgo_n, d1, d1), dtype=np.float32)
        self._check(self.lib.srrg2b_pgo_download(self.h, out.ctypes.data))
        return out

    def last_run_timing(self):
        ms = C.c_float(0)
        it = C.c_int32(0)
        self._check(self.lib.srrg2b_last_run_timing(self.h, C.byref(ms), C.byref(it)))
        return ms.value, it.value
