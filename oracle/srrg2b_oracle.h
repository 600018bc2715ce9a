/*
 * srrg2b_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the srrg2_slam_interfaces MultiAligner ICP loop and of the
 * arithmetic it reaches in its un-vendored dependencies (srrg2_solver / srrg2_core / the
 * downstream kd-tree finders).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.
 *
 * PARITY STATUS: the control flow is pinned by the in-tree reference sources cited at each
 * function; the per-term arithmetic (factors, robustifiers, NN tie-breaking) lives in
 * dependencies that are absent from /root/reference and unpinned -> "parity unpinned" for
 * those parts.  The only reference test that reaches the solver
 * (tests/test_motion_model_slice.cpp:81-85,139-142,220-223) is restated in
 * tests/test_oracle_reference_kat.py and holds for this oracle.
 *
 * Paths: R/ = /root/reference/srrg2_slam_interfaces/src/srrg2_slam_interfaces/
 */
#ifndef SRRG2B_ORACLE_H
#define SRRG2B_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- enums (values mirror include/srrg2b.h; AlignerBase::Status from
 *      R/registration/aligners/aligner.h:23-28) ---- */
enum { ORC_FACTOR_P2P = 0, ORC_FACTOR_PLANE = 1 };
enum { ORC_ROB_NONE = 0, ORC_ROB_SATURATED = 1, ORC_ROB_CAUCHY = 2, ORC_ROB_CLAMP = 3, ORC_ROB_HUBER = 4 };
enum { ORC_VAR_SE3_QUAT_RIGHT = 0, ORC_VAR_SE3_EULER_RIGHT = 1 };
enum { ORC_FINDER_NN = 0, ORC_FINDER_PROJECTIVE = 1 };
enum { ORC_SLICE_POINTS = 0, ORC_SLICE_PRIOR = 1 };
enum { ORC_ALIGNER_SUCCESS = 0, ORC_ALIGNER_NOT_ENOUGH_CORR = 1, ORC_ALIGNER_NOT_ENOUGH_INLIERS = 2, ORC_ALIGNER_FAIL = 3 };
enum { ORC_STAT_INLIER = 0, ORC_STAT_KERNELIZED = 1, ORC_STAT_SUPPRESSED = 2, ORC_STAT_NONE = 3 };
enum { ORC_NN_BRUTE = 0, ORC_NN_KDTREE = 1 };

typedef struct {
  int32_t kind;
  float max_distance;
  float normal_cos;
  float fx, fy, cx, cy;
  int32_t width, height;
  float min_depth, max_depth;
} orc_finder_params;

typedef struct {
  int32_t factor;
  int32_t robustifier;
  float chi_threshold;
  float info_point;
  float info_normal;
} orc_factor_params;

typedef struct {
  const float* coords;   /* n x dim packed */
  const float* normals;  /* n x dim packed or NULL */
  const uint8_t* valid;  /* n or NULL */
  int64_t n;
} orc_cloud;

typedef struct {
  int32_t kind;          /* ORC_SLICE_POINTS | ORC_SLICE_PRIOR */
  int32_t min_num_correspondences;
  orc_cloud fixed, moving;
  float robot_in_sensor[16]; /* row-major (dim+1)^2 */
  orc_finder_params finder;
  orc_factor_params factor;
  float prior_measurement[16];
  float prior_info_diag[6];
} orc_slice;

typedef struct {
  int32_t iteration;
  int32_t solver_status; /* 1 = Success */
  int64_t num_inliers, num_outliers, num_suppressed, num_correspondences;
  double chi_inliers, chi_outliers;
  int64_t num_saturated; /* suppressed because |S m - f| left the fixed-point error range (part of num_suppressed) */
} orc_iter_stats;

typedef struct {
  int32_t variable;
  int32_t max_iterations;
  int32_t min_num_inliers;
  int32_t enable_inlier_only_runs;
  int32_t keep_only_inlier_correspondences;
  int32_t use_termination_criteria;
  int32_t window_size, num_correspondences_range, num_inliers_range, num_outliers_range;
  float chi_epsilon;
} orc_aligner_params;

/* correspondence output of one slice: arrays sized n_moving, compact, ascending moving_idx */
typedef struct {
  int32_t* fixed_idx;
  int32_t* moving_idx;
  float* response;
  int64_t n;
} orc_corr_out;

/* fixed-point scale exponents per accumulated class (see oracle .c: orc_scales) */
enum { ORC_K_HTT = 0, ORC_K_HTR = 1, ORC_K_HRR = 2, ORC_K_BT = 3, ORC_K_BR = 4, ORC_K_CHI = 5, ORC_K_CHI_LO = 6, ORC_K_COUNT = 7 };
#define ORC_ACC_SLOTS 40
typedef struct { int32_t k[ORC_K_COUNT]; float err_bound; /* largest |S m - f| the ranges cover */ } orc_scales_t;

int orc_set_threads(int n);   /* OpenMP threads for the finder / lineariser loops; returns actual */

/* exact NN finder handle (kd-tree over the valid fixed points) */
typedef struct orc_index orc_index;
orc_index* orc_index_create(int dim, const orc_cloud* fixed, int method);
void orc_index_free(orc_index*);

/* a3: CorrespondenceFinder_::compute(); dense output sized moving->n: fixed_idx[j] = -1 if none */
int orc_find(const orc_index* index, int dim, const orc_cloud* fixed, const orc_cloud* moving,
             const float* S, const orc_finder_params* fp, int32_t* fixed_idx_dense, float* response_dense);

/* a5 (linearise part): from a dense per-moving-point fixed index */
int orc_scales(int dim, int variable, float radius_bound2, float normal_bound2, const orc_finder_params* fp,
               const orc_factor_params* fa, orc_scales_t* out);
float orc_coord_bound(int dim, const orc_cloud* moving);
float orc_radius_bound2(int dim, const orc_cloud* moving);  /* max |m|^2 over the valid points */
float orc_normal_bound2(int dim, const orc_cloud* cloud);  /* max |n|^2 over the valid points, 0 without normals */
int orc_linearize(int dim, int variable, const orc_cloud* fixed, const orc_cloud* moving,
                  const int32_t* fixed_idx_dense, const float* S, const orc_finder_params* fp,
                  const orc_factor_params* fa, float radius_bound2_global /* <=0: from moving */,
                  float normal_bound2_global /* <=0: from both clouds */,
                  int64_t* acc /*[ORC_ACC_SLOTS] fixed point*/, double* H /*36 or 9 full row-major*/, double* b,
                  orc_iter_stats* stats, uint8_t* status_dense, float* chi_dense,
                  double* plain /* NULL or [32]: un-quantised fp64 sums of the same fp32 terms (H upper triangle,
                                   b, chi_in at [27], chi_out at [28]) -- the independent check of the fixed point */);

/* a1..a9: MultiAlignerBase_::compute() */
int orc_icp_run(int dim, int n_slices, const orc_slice* slices, const orc_aligner_params* ap,
                float* T_inout, orc_iter_stats* stats_out, int32_t* n_stats, int32_t* status_out,
                orc_corr_out* corr_out /* n_slices entries or NULL */, int nn_method);

/* deterministic helpers exported for tests */
void orc_sincos(double x, double* s, double* c);
double orc_atan2(double y, double x);
double orc_log(double x);
void orc_fix_transform(int dim, float* T);
void orc_t2v(int dim, const float* T, float* v);
void orc_v2t(int dim, int variable, const float* v, float* T);
void orc_mul(int dim, const float* A, const float* B, float* C);
void orc_inverse(int dim, const float* A, float* Ainv);
/* N2: MergerCorrespondenceHomo_::compute() on flat arrays (see the .c file); returns the new scene size */
int64_t orc_scene_merge(int dim, float* scene_coords, float* scene_normals, uint8_t* scene_valid, int64_t n_scene,
                        const orc_cloud* meas, const float* T, int64_t n_corr, const int32_t* corr_scene, const int32_t* corr_meas,
                        const float* corr_resp, float maximum_response, float maximum_distance_geometry_squared,
                        int64_t target_number_of_merges, int64_t* n_merged_out, int64_t* n_added_out);
/* N1: range clip of a resident scene (see the .c file); returns the number of points kept */
int64_t orc_scene_clip(int dim, const orc_cloud* scene, const float* T, float max_range, float* out_coords, float* out_normals,
                       int32_t* global_indices);
int orc_solve_update(int dim, int variable, const double* H, const double* b, float* T);

#ifdef __cplusplus
}
#endif
#endif
