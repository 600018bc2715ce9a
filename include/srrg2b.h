/*
 * srrg2b.h -- C ABI of libsrrg2b.so: B200 (sm_100a) implementation of the two data-parallel hot
 * paths of srrg2_slam_interfaces (the MultiAligner ICP loop and the pose-graph Gauss-Newton step).
 *
 * Every entry point is extern "C", takes plain pointers and sizes, returns an int error code
 * (0 = ok) and has finished (results visible on the host) when it returns, matching the
 * single-threaded, synchronous calling convention of the reference.
 * Configuration errors map to SRRG2B_ERR_INVALID (the reference throws std::runtime_error, e.g.
 * R/registration/aligners/aligner_slice_processor_impl.cpp:13-16); numeric outcomes are reported
 * through status enums (R/registration/aligners/aligner.h:23-28), never as errors.
 *
 * R/ = srrg2_slam_interfaces/src/srrg2_slam_interfaces/ of the reference tree.
 * Matrices are row-major (dim+1)x(dim+1) float (Isometry2f / Isometry3f .matrix()).
 * There is NO CPU fallback: every compute entry point fails with SRRG2B_ERR_CUDA without a GPU.
 */
#ifndef SRRG2B_H
#define SRRG2B_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SRRG2B_VERSION 100
#define SRRG2B_MAX_SLICES 8

enum { SRRG2B_OK = 0, SRRG2B_ERR_INVALID = 1, SRRG2B_ERR_CUDA = 2, SRRG2B_ERR_STATE = 3, SRRG2B_ERR_NCCL = 4 };
enum { SRRG2B_FIXED = 0, SRRG2B_MOVING = 1 };
/* factor types instantiated by the reference: R/instances.h:27-30,70-73 (Point2Point) and the
 * point+normal factor the laser/proslam pipelines plug into AlignerSliceProcessor_ */
enum { SRRG2B_FACTOR_P2P = 0, SRRG2B_FACTOR_PLANE = 1 };
/* RobustifierBase subclasses (R/registration/aligners/aligner_slice_processor_base.h:34-38,
 * RobustifierClamp at R/registration/aligners/multi_aligner_impl.cpp:194) */
enum { SRRG2B_ROB_NONE = 0, SRRG2B_ROB_SATURATED = 1, SRRG2B_ROB_CAUCHY = 2, SRRG2B_ROB_CLAMP = 3, SRRG2B_ROB_HUBER = 4 };
/* VariableSE3QuaternionRightAD (MultiAligner3DQR) / VariableSE3EulerRightAD (MultiAligner3D):
 * R/registration/aligners/multi_aligner.h:152-158; dim 2 is always VariableSE2RightAD */
enum { SRRG2B_VAR_SE3_QUAT_RIGHT = 0, SRRG2B_VAR_SE3_EULER_RIGHT = 1 };
enum { SRRG2B_FINDER_NN = 0, SRRG2B_FINDER_PROJECTIVE = 1 };
enum { SRRG2B_SLICE_POINTS = 0, SRRG2B_SLICE_PRIOR = 1 };
/* AlignerBase::Status, R/registration/aligners/aligner.h:23-28 */
enum { SRRG2B_ALIGNER_SUCCESS = 0, SRRG2B_ALIGNER_NOT_ENOUGH_CORRESPONDENCES = 1, SRRG2B_ALIGNER_NOT_ENOUGH_INLIERS = 2, SRRG2B_ALIGNER_FAIL = 3 };
/* srrg2_solver FactorStats::Status as consumed at R/registration/aligners/multi_aligner_impl.cpp:244 */
enum { SRRG2B_STAT_INLIER = 0, SRRG2B_STAT_KERNELIZED = 1, SRRG2B_STAT_SUPPRESSED = 2, SRRG2B_STAT_NONE = 3 };

typedef struct srrg2b_ctx srrg2b_ctx;

/* A point(+normal) cloud slice: PointNormal{2,3}fVectorCloud flattened to packed fp32 arrays. */
typedef struct {
  const float* coords;    /* n x dim */
  const float* normals;   /* n x dim or NULL */
  const uint8_t* valid;   /* n (point.status == Valid) or NULL = all valid */
  int64_t n;
  int64_t index_offset;   /* global moving index of element 0 when the moving cloud is sharded */
  int64_t n_global;       /* size of the whole (unsharded) cloud; 0 = n */
  int32_t on_device;      /* 1: the pointers are device pointers on the context's GPU */
  int32_t reserved;
} srrg2b_cloud;

/* CorrespondenceFinder_ parameters (kd-tree NN finder of srrg2_laser_slam_2d, projective finder of
 * srrg2_proslam; interface R/registration/correspondence_finder.h:66-125) */
typedef struct {
  int32_t kind;
  float max_distance;
  float normal_cos;       /* <= -1 disables the normal gate */
  float fx, fy, cx, cy;   /* projective only */
  int32_t width, height;
  float min_depth, max_depth;
} srrg2b_finder_params;

typedef struct {
  int32_t factor;
  int32_t robustifier;
  float chi_threshold;    /* RobustifierBase::param_chi_threshold */
  float info_point;       /* information of the point rows (isotropic for P2P) */
  float info_normal;      /* information of the normal rows (PLANE factor) */
} srrg2b_factor_params;

/* One entry of MultiAlignerBase_::param_slice_processors (R/registration/aligners/multi_aligner.h:34-37):
 * an AlignerSliceProcessor_ (points) or an AlignerSliceProcessorPrior_ (odometry / motion model). */
typedef struct {
  int32_t kind;
  int32_t slice_id;                 /* cloud slot set with srrg2b_set_cloud */
  int32_t min_num_correspondences;  /* aligner_slice_processor.h:61-66, strict > */
  int32_t reserved;
  float robot_in_sensor[16];        /* aligner_slice_processor.h:142-150 */
  srrg2b_finder_params finder;
  srrg2b_factor_params factor;
  float prior_measurement[16];      /* prior: factor measurement AND initial guess (odometry_prior.cpp:19,34) */
  float prior_info_diag[6];         /* param_diagonal_info_matrix */
} srrg2b_slice;

/* srrg2_solver::IterationStats fields the reference reads (aligner_termination_criteria_impl.cpp:30-32) */
typedef struct {
  int32_t iteration;
  int32_t solver_status;            /* 1 = SolverBase::Success */
  int64_t num_inliers, num_outliers, num_suppressed, num_correspondences;
  double chi_inliers, chi_outliers;
  /* correspondences suppressed because |S m - f| exceeds the finder radius the fixed-point ranges were derived
   * for (part of num_suppressed).  Always 0 for finder output; counts bad pairs of srrg2b_set_correspondences. */
  int64_t num_saturated;
} srrg2b_iter_stats;

/* AlignerBase / MultiAlignerBase_ / AlignerTerminationCriteriaStandard_ parameters, same names and
 * defaults: aligner.h:30-35, multi_aligner.h:45-57, aligner_termination_criteria.h:40-56 */
typedef struct {
  int32_t variable;
  int32_t max_iterations;
  int32_t min_num_inliers;
  int32_t enable_inlier_only_runs;
  int32_t keep_only_inlier_correspondences;
  int32_t use_termination_criteria;
  int32_t window_size, num_correspondences_range, num_inliers_range, num_outliers_range;
  float chi_epsilon;
} srrg2b_aligner_params;

/* ---- context ---- */
int srrg2b_version(void);
int srrg2b_ctx_create(int dim, int device, srrg2b_ctx** out);
int srrg2b_ctx_destroy(srrg2b_ctx* ctx);
const char* srrg2b_last_error(const srrg2b_ctx* ctx);
/* CUDA stream the context launches on (cudaStream_t), for callers that time with events */
void* srrg2b_stream(srrg2b_ctx* ctx);
/* number of kernels this context has launched so far */
int64_t srrg2b_launch_count(const srrg2b_ctx* ctx);

/* ---- multi-GPU: one context per rank; only the reduced H/b/stats cross NVLink ---- */
int srrg2b_comm_unique_id(void* id_128_bytes);
int srrg2b_comm_init(srrg2b_ctx* ctx, const void* id_128_bytes, int rank, int world_size);

/* ---- data: CorrespondenceFinder_::setFixed / setMoving (correspondence_finder.h:80-91) ---- */
int srrg2b_set_cloud(srrg2b_ctx* ctx, int slot, int slice_id, const srrg2b_cloud* cloud);

/* ---- N1: the local map stays in HBM between frames and is clipped on the device.
 * Replaces TrackerSliceProcessor_::clip (R/trackers/tracker_slice_processor_impl.cpp:194-205) with a range clipper
 * behind SceneClipper's contract (R/mapping/scene_clipper.h:104-107: clipped scene in the robot frame + the indices
 * of its points in the full scene): a valid scene point p is kept iff |T p| <= max_range (T = scene_in_robot,
 * row-major (dim+1)^2); the kept points T p (normals R n), in ascending scene index, become the MOVING cloud of
 * `slice_id` without crossing PCIe.  srrg2b_scene_clip_indices returns the scene index of every clipped point
 * (n_clipped entries): correspondences of the slice carry indices into the clipped cloud. ---- */
int srrg2b_scene_set(srrg2b_ctx* ctx, int scene_id, const srrg2b_cloud* cloud);
int srrg2b_scene_clip(srrg2b_ctx* ctx, int scene_id, int slice_id, const float* scene_in_robot, float max_range,
                      int64_t* n_clipped);
int srrg2b_scene_clip_indices(srrg2b_ctx* ctx, int slice_id, int32_t* global_indices);

/* ---- N2: MergerCorrespondenceHomo_::compute() (R/mapping/merger_correspondence_homo_impl.cpp:11-126)
 * on the resident scene.  The measurement is the FIXED cloud of `slice_id` (the new scan, already on the device), the
 * correspondences are the ones the aligner left in that slice -- flipped and mapped to the global scene through the
 * clip's indices, as R/trackers/tracker_slice_processor_impl.cpp:159-191 does on the host -- so nothing crosses PCIe.
 * For every correspondence with response < maximum_response whose points lie closer than
 * sqrt(maximum_distance_geometry_squared) in the scene frame, the scene point takes the measurement point's fields and
 * the mid-point as coordinates; if fewer than target_number_of_merges DISTINCT measurement points were merged, the
 * valid unmerged ones are appended (transformed into the scene frame).  without_correspondences = 1: the
 * "no correspondences set" branch (:31-42, first frame of a local map): every valid measurement point is appended.
 * Parameter names and defaults: merger_correspondence_homo.h:22-31, merger.h:126-131. ---- */
typedef struct {
  float maximum_response;                    /* 50 */
  float maximum_distance_geometry_squared;   /* 0.25 */
  int32_t target_number_of_merges;           /* 200 */
  int32_t without_correspondences;           /* 0 */
} srrg2b_merge_params;
int srrg2b_scene_merge(srrg2b_ctx* ctx, int scene_id, int slice_id, const float* measurement_in_scene,
                       const srrg2b_merge_params* params, int64_t* n_merged, int64_t* n_added);
/* the resident scene back on the host (map serialisation, tests); any output pointer may be NULL; *n = its size */
int srrg2b_scene_get(srrg2b_ctx* ctx, int scene_id, float* coords, float* normals, uint8_t* valid, int64_t* n);

/* ---- a3: CorrespondenceFinder_::compute() (correspondence_finder.h:56). S = local_map_in_sensor
 * (:111-114). Output: ascending moving_idx, at most one entry per moving point; buffers must hold
 * n_moving entries; any of them may be NULL. */
int srrg2b_find_correspondences(srrg2b_ctx* ctx, int slice_id, const float* S, const srrg2b_finder_params* fp,
                                int32_t* fixed_idx, int32_t* moving_idx, float* response, int64_t* n_out);
/* HBST path (R/registration/loop_detector/multi_loop_detector_hbst_impl.cpp:331-352): externally
 * matched correspondences, evaluated by srrg2b_linearize (srrg2b_icp_iterate / srrg2b_icp_run search their
 * own correspondences and overwrite supplied ones).  Pairs naming a masked-out moving point are rejected
 * with SRRG2B_ERR_INVALID; pairs naming an unknown fixed point are kept and counted as suppressed. */
int srrg2b_set_correspondences(srrg2b_ctx* ctx, int slice_id, const int32_t* fixed_idx, const int32_t* moving_idx, int64_t n);

/* ---- a5 (linearise): FactorCorrespondenceDriven_ accumulation over the slice's current
 * correspondences at S. H is PxP row-major (P = 6 | 3), b is P; acc (optional) receives the 40
 * exact fixed-point sums (21 H, 6 b, chi in/out as coarse+residual words, 4 counters, padding); status/chi
 * (optional) are per correspondence in ascending moving_idx.  fp->max_distance is the residual bound the
 * fixed-point ranges are derived for: a pair with |S m - f| beyond it is suppressed and counted in
 * num_saturated (never clamped).  With several ranks every rank receives the global sums. */
int srrg2b_linearize(srrg2b_ctx* ctx, int slice_id, const float* S, int variable, const srrg2b_finder_params* fp,
                     const srrg2b_factor_params* fa, double* H, double* b, int64_t* acc,
                     srrg2b_iter_stats* stats, uint8_t* status, float* chi);

/* ---- a1..a9: MultiAlignerBase_::compute() (multi_aligner_impl.cpp:46-95) entirely on the device:
 * all iterations are enqueued back to back, termination is decided on the GPU, one sync at the end.
 * stats_out holds *n_stats entries on input (capacity) and receives the appended IterationStats. */
int srrg2b_icp_run(srrg2b_ctx* ctx, int n_slices, const srrg2b_slice* slices, const srrg2b_aligner_params* ap,
                   float* T_inout, srrg2b_iter_stats* stats_out, int32_t* n_stats, int32_t* aligner_status);
/* one _runSolver iteration (multi_aligner_impl.cpp:103-126 body): correspondences + one GN step */
int srrg2b_icp_iterate(srrg2b_ctx* ctx, int n_slices, const srrg2b_slice* slices, int variable,
                       float* T_inout, srrg2b_iter_stats* stats, int32_t* association_good);
/* correspondences of the last iteration (pruned to inliers if keep_only_inlier_correspondences),
 * Aligner_::storeCorrespondences() (aligner_slice_processor_impl.cpp:50-74) */
int srrg2b_get_correspondences(srrg2b_ctx* ctx, int slice_id, int32_t* fixed_idx, int32_t* moving_idx,
                               float* response, int64_t* n_out);

/* Forgets the slice's correspondences, warm-start candidates and certified bounds, as a fresh setMoving()
 * would (the clouds and the NN index stay): the next compute() starts cold.  Results never depend on it. */
int srrg2b_reset_correspondences(srrg2b_ctx* ctx, int slice_id);

/* ---- N3: candidate-batched loop closing / relocalisation.
 * MultiLoopDetectorBruteForce_::compute (R/registration/loop_detector/multi_loop_detector_brute_force_impl.cpp:63-133)
 * and MultiRelocalizer_::compute (R/registration/relocalization/multi_relocalizer_impl.cpp:74-138) run ONE aligner
 * serially over K candidate local maps: setFixed(source) once, then per candidate setMoving(target),
 * setMovingInFixed(guess), compute(), and the three gates on the last IterationStats.  Here every candidate owns
 * a context (its own stream, solver state and CUDA graph), the K runs are in flight together -- first runs of all
 * candidates, then the inlier-only runs of those that passed -- and the gates are applied in candidate order.
 *
 * srrg2b_share_fixed: the source local map is uploaded and indexed ONCE (srrg2b_set_cloud(FIXED) on `src`) and lent
 * to the candidates' contexts: `dst`'s slice reads src's fixed cloud and NN index in place (no copy).  Contract: same
 * device and dimension; `src`'s slice must not be re-set or destroyed while a borrower still uses it; a borrower that
 * sets its own fixed cloud, or searches with a different radius, silently gets private buffers again. */
int srrg2b_share_fixed(srrg2b_ctx* dst, int dst_slice_id, srrg2b_ctx* src, int src_slice_id);

typedef struct {
  int32_t relocalize_min_inliers;        /* multi_loop_detector_brute_force.h: param_relocalize_min_inliers */
  float relocalize_max_chi_inliers;      /* param_relocalize_max_chi_inliers (chi per inlier) */
  float relocalize_min_inliers_ratio;    /* param_relocalize_min_inliers_ratio */
} srrg2b_closure_params;

#define SRRG2B_CLOSURE_ACCEPT 0
#define SRRG2B_CLOSURE_ALIGNER_DROP 1        /* :80-84  status != Success */
#define SRRG2B_CLOSURE_NUM_INLIERS_DROP 2    /* :94-97 */
#define SRRG2B_CLOSURE_MAX_CHI_DROP 3        /* :99-103 */
#define SRRG2B_CLOSURE_INLIER_RATIO_DROP 4   /* :105-111 */

typedef struct {
  int32_t verdict;                 /* SRRG2B_CLOSURE_* */
  int32_t aligner_status;          /* SRRG2B_ALIGNER_* */
  int32_t iterations;              /* IterationStats entries of the run(s) */
  int32_t reserved;
  int64_t num_correspondences;     /* aligner->numCorrespondences() after compute(): point slices' lists (pruned when
                                      keep_only_inlier_correspondences) + 1 per prior slice (multi_aligner_impl.cpp:275-285) */
  int64_t num_inliers;             /* last IterationStats */
  float chi_inliers;               /* istat.chi_inliers / num_inliers, fp32 division (:91) */
  float device_ms;                 /* device time of this candidate's run(s) */
  float moving_in_fixed[16];       /* aligner->movingInFixed(), row-major (dim+1)^2; the closure's measurement */
} srrg2b_closure_result;

/* ctxs[k]: candidate k's context, holding the slices' clouds (fixed side typically lent by srrg2b_share_fixed);
 * guesses: K row-major (dim+1)^2 initial guesses (h->initial_guess); the slice list and aligner parameters are the
 * detector's one aligner configuration, used for every candidate.  Results in candidate order.  The correspondences
 * of candidate k stay in ctxs[k] (srrg2b_get_correspondences / srrg2b_scene_merge). */
int srrg2b_closure_batch(srrg2b_ctx* const* ctxs, int k, int n_slices, const srrg2b_slice* slices,
                         const srrg2b_aligner_params* ap, const float* guesses, const srrg2b_closure_params* cp,
                         srrg2b_closure_result* results);

/* ---- benchmarking support: device time of the last srrg2b_icp_run between its first and last
 * kernel (CUDA events on the context stream), and how many _runSolver iterations it executed ---- */
int srrg2b_last_run_timing(srrg2b_ctx* ctx, float* device_ms, int32_t* iterations);
/* when enabled, every launch of the fused per-slice ICP kernel is bracketed by CUDA events on the
 * context stream; the sum and count for the last run are returned (roofline measurement) */
/* diagnostics of a slice's NN index: out16 = {R, nx, ny, nz, n_fixed_valid, n_moving_valid,
 * size of the last phase-2 worklist, cell edge (float bits), size of the last coherence worklist, 0...} */
int srrg2b_debug_info(srrg2b_ctx* ctx, int slice_id, int32_t* out16);
int srrg2b_set_kernel_timing(srrg2b_ctx* ctx, int enable);
int srrg2b_last_kernel_timing(srrg2b_ctx* ctx, float* slice_kernel_ms, int32_t* slice_kernel_launches);

/* ---- a10: pose-graph Gauss-Newton, MultiGraphSLAM_::optimize() -> Solver::compute()
 * (R/system/multi_graph_slam_impl.cpp:299-317).  Variables are the local maps' poses (R/mapping/local_map.h:64,75):
 * context dim 3: SE(3), row-major 4x4 float each, factors SE3PosePoseGeodesicErrorFactor
 * (R/registration/loop_closure.h:111) with a 4x4 measurement Z and a 6x6 information; context dim 2: SE(2),
 * row-major 3x3 float each, factors SE2PosePoseGeodesicErrorFactor (R/registration/loop_closure.h:110) with a 3x3
 * measurement and a 3x3 information.  e = t2v(Z^-1 Xi^-1 Xj), right perturbation.  With a communicator the factor
 * list is sharded over the ranks (every rank uploads the whole graph) and only H / b / chi are all-reduced.
 * H and b are assembled without atomics (per-factor records gathered in factor order) and every dot product of
 * the linear solve is summed in a fixed order: results are bit-identical from run to run. ---- */
typedef struct {
  double chi;                   /* sum of e^T Omega e at the linearisation point */
  double chi_after;             /* ... at the updated poses (srrg2b_pgo_optimize; -1 from srrg2b_pgo_iterate) */
  double dx_norm_inf;           /* largest perturbation component of the step */
  double cg_relative_residual;  /* |H dx + b| / |b| reached by the linear solve */
  double lambda;                /* Levenberg-Marquardt damping used (0 for srrg2b_pgo_iterate) */
  double gain_ratio;            /* actual / predicted decrease of chi */
  int32_t cg_iterations;
  int32_t num_factors;
  int32_t num_blocks;           /* DxD blocks of the block-CSR system matrix */
  int32_t accepted;             /* the step was applied (a rejected Levenberg-Marquardt step leaves the poses unchanged) */
  float linearize_ms, solve_ms; /* device times of the two phases */
} srrg2b_pgo_stats;
int srrg2b_pgo_upload(srrg2b_ctx* ctx, int64_t n_vars, const float* poses, const uint8_t* fixed_mask,
                      int64_t n_factors, const int32_t* ij, const float* Z, const float* Omega);
/* one plain Gauss-Newton iteration: linearise, solve H dx = -b (block-Jacobi PCG to cg_tolerance), X <- X [+] dx */
int srrg2b_pgo_iterate(srrg2b_ctx* ctx, int max_cg_iterations, double cg_tolerance, srrg2b_pgo_stats* stats);
/* Solver::compute() as optimize() uses it: iterate until |dx|_inf < dx_tolerance or max_iterations; the steps are
 * guarded by Levenberg-Marquardt damping with a gain-ratio test, the linear solves are inexact (their tolerance
 * follows the outer convergence).  stats: one entry per iteration (NULL allowed), *n_done: iterations run. */
int srrg2b_pgo_optimize(srrg2b_ctx* ctx, int max_iterations, double dx_tolerance, int max_cg_iterations,
                        srrg2b_pgo_stats* stats, int32_t* n_done);
int srrg2b_pgo_download(srrg2b_ctx* ctx, float* poses);

#ifdef __cplusplus
}
#endif
#endif
