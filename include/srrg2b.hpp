// srrg2b.hpp -- header-only C++17 host side above the C ABI of libsrrg2b.so (include/srrg2b.h).
//
// The reference is compiled C++ whose dependencies (srrg2_core, srrg2_solver, Eigen, catkin) are not
// available where this repository is built, so the reference-side adapter classes of INTEGRATION.md
// cannot be compiled here.  This header is the same host logic WITHOUT those dependencies: classes
// that mirror the reference's plugin interfaces for the hot path -- same method names, argument
// meaning, defaults and error behaviour -- on plain STL types, so that a test written against them
// reads like the reference's own:
//
//   CorrespondenceFinderB200<Dim>   CorrespondenceFinder_  (R/registration/correspondence_finder.h:41-124)
//   MultiAlignerB200<Dim>           Aligner_ / MultiAlignerBase_ (R/registration/aligners/aligner.h:46-127,
//                                   multi_aligner.h:34-66) with its slice processors
//                                   (aligner_slice_processor.h:56-66, aligner_slice_odometry_prior.h:9-45)
//   PoseGraphSolverB200             Solver as used by MultiGraphSLAM_::optimize()
//                                   (R/system/multi_graph_slam_impl.cpp:299-317)
//
// R/ = srrg2_slam_interfaces/src/srrg2_slam_interfaces/ of the reference tree.  Misconfiguration throws
// std::runtime_error (aligner_slice_processor_impl.cpp:13-16); numeric outcomes are status values
// (aligner.h:23-28).  There is no CPU fallback: without a CUDA device the constructors throw.
#ifndef SRRG2B_HPP
#define SRRG2B_HPP
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "srrg2b.h"

namespace srrg2b {

// srrg2_core::Correspondence(fixed_idx, moving_idx, response): constructor order proven by
// R/registration/loop_detector/multi_loop_detector_hbst_impl.cpp:183-191
struct Correspondence {
  int fixed_idx = -1, moving_idx = -1;
  float response = 0.f;
  Correspondence() = default;
  Correspondence(int f, int m, float r) : fixed_idx(f), moving_idx(m), response(r) {}
};
using CorrespondenceVector = std::vector<Correspondence>;

// PointNormal{2,3}fVectorCloud flattened: coordinates n x Dim, optional normals, optional validity
template <int Dim>
struct PointNormalCloud {
  std::vector<float> coordinates, normals;
  std::vector<uint8_t> valid;  // point.status == Valid; empty = all valid
  size_t size() const { return coordinates.size() / Dim; }
};

// Isometry{2,3}f as the row-major (Dim+1) x (Dim+1) matrix the C ABI documents
template <int Dim>
struct Isometry {
  static constexpr int N = (Dim + 1) * (Dim + 1);
  std::array<float, N> m;
  static Isometry Identity() {
    Isometry T;
    T.m.fill(0.f);
    for (int i = 0; i <= Dim; ++i) T.m[i * (Dim + 1) + i] = 1.f;
    return T;
  }
  const float* data() const { return m.data(); }
  float* data() { return m.data(); }
  float operator()(int r, int c) const { return m[(size_t) (r * (Dim + 1) + c)]; }
  float& operator()(int r, int c) { return m[(size_t) (r * (Dim + 1) + c)]; }
  Isometry inverse() const {  // [R t]^-1 = [R^T, -R^T t]
    Isometry o = Identity();
    for (int r = 0; r < Dim; ++r) {
      float t = 0.f;
      for (int c = 0; c < Dim; ++c) { o(r, c) = (*this)(c, r); t += (*this)(c, r) * (*this)(c, Dim); }
      o(r, Dim) = -t;
    }
    return o;
  }
  Isometry operator*(const Isometry& B) const {
    Isometry o;
    for (int r = 0; r <= Dim; ++r)
      for (int c = 0; c <= Dim; ++c) {
        float t = 0.f;
        for (int k = 0; k <= Dim; ++k) t += (*this)(r, k) * B(k, c);
        o(r, c) = t;
      }
    return o;
  }
};

// srrg2_solver::IterationStats fields the reference reads (aligner_termination_criteria_impl.cpp:30-32)
using IterationStats = srrg2b_iter_stats;
using IterationStatsVector = std::vector<IterationStats>;

// AlignerBase::Status, aligner.h:23-28 (same numeric values)
enum class AlignerStatus { Success = 0, NotEnoughCorrespondences = 1, NotEnoughInliers = 2, Fail = 3 };

// one CUDA context (device, stream, resident clouds); shared by the modules built on it
class Context {
public:
  explicit Context(int dim, int device = 0) {
    const int rc = srrg2b_ctx_create(dim, device, &_ctx);
    if (rc != SRRG2B_OK)
      throw std::runtime_error(rc == SRRG2B_ERR_CUDA ? "srrg2b::Context|no usable CUDA device (there is no CPU fallback)"
                                                     : "srrg2b::Context|invalid arguments");
  }
  ~Context() { srrg2b_ctx_destroy(_ctx); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  srrg2b_ctx* get() const { return _ctx; }
  void check(int rc, const char* who) const {
    if (rc != SRRG2B_OK) throw std::runtime_error(std::string(who) + "|" + srrg2b_last_error(_ctx));
  }

private:
  srrg2b_ctx* _ctx = nullptr;
};
using ContextPtr = std::shared_ptr<Context>;

namespace detail {
template <int Dim>
inline srrg2b_cloud describe(const PointNormalCloud<Dim>& c) {
  if (!c.normals.empty() && c.normals.size() != c.coordinates.size())
    throw std::runtime_error("srrg2b|normals and coordinates differ in size");
  if (!c.valid.empty() && c.valid.size() != c.size()) throw std::runtime_error("srrg2b|validity mask has the wrong size");
  srrg2b_cloud d;
  std::memset(&d, 0, sizeof(d));
  d.coords = c.coordinates.data();
  d.normals = c.normals.empty() ? nullptr : c.normals.data();
  d.valid = c.valid.empty() ? nullptr : c.valid.data();
  d.n = (int64_t) c.size();
  return d;
}
template <int Dim>
inline void embed(const Isometry<Dim>& T, float* out16) {  // the slice descriptors always carry 16 floats
  std::memset(out16, 0, 16 * sizeof(float));
  std::memcpy(out16, T.data(), sizeof(float) * Isometry<Dim>::N);
}
}  // namespace detail

// ---------------------------------------------------------------------------------------------
// CorrespondenceFinder_ (a3): exact nearest neighbour within max_distance + normal gate
// ---------------------------------------------------------------------------------------------
template <int Dim>
class CorrespondenceFinderB200 {
public:
  using CloudType = PointNormalCloud<Dim>;
  using EstimateType = Isometry<Dim>;
  float param_max_distance_m = 0.5f;  // kd-tree finder of srrg2_laser_slam_2d
  float param_normal_cos = 0.8f;

  explicit CorrespondenceFinderB200(ContextPtr ctx, int slice_id = 0) : _ctx(std::move(ctx)), _slice(slice_id) {}
  // correspondence_finder.h:41-56
  void setCorrespondences(CorrespondenceVector* c) { _correspondences = c; }
  // correspondence_finder.h:80-91: change detection is by flag, not by content
  void setFixed(const CloudType* f) { _fixed = f; _fixed_changed_flag = true; }
  void setMoving(const CloudType* m) { _moving = m; _moving_changed_flag = true; }
  // correspondence_finder.h:111-114
  void setLocalMapInSensor(const EstimateType& T) { _local_map_in_sensor = T; }

  void compute() {  // correspondence_finder.h:56
    if (!_fixed || !_moving) throw std::runtime_error("CorrespondenceFinderB200::compute|fixed or moving not set");
    if (!_correspondences) throw std::runtime_error("CorrespondenceFinderB200::compute|correspondence vector not set");
    if (_fixed_changed_flag) {
      const srrg2b_cloud d = detail::describe(*_fixed);
      _ctx->check(srrg2b_set_cloud(_ctx->get(), SRRG2B_FIXED, _slice, &d), "CorrespondenceFinderB200::setFixed");
      _fixed_changed_flag = false;
    }
    if (_moving_changed_flag) {
      const srrg2b_cloud d = detail::describe(*_moving);
      _ctx->check(srrg2b_set_cloud(_ctx->get(), SRRG2B_MOVING, _slice, &d), "CorrespondenceFinderB200::setMoving");
      _moving_changed_flag = false;
    }
    srrg2b_finder_params fp;
    std::memset(&fp, 0, sizeof(fp));
    fp.kind = SRRG2B_FINDER_NN;
    fp.max_distance = param_max_distance_m;
    fp.normal_cos = param_normal_cos;
    const size_t n = _moving->size();
    std::vector<int32_t> fi(n), mi(n);
    std::vector<float> rs(n);
    int64_t m = 0;
    _ctx->check(srrg2b_find_correspondences(_ctx->get(), _slice, _local_map_in_sensor.data(), &fp, fi.data(), mi.data(),
                                            rs.data(), &m),
                "CorrespondenceFinderB200::compute");
    _correspondences->clear();
    _correspondences->reserve((size_t) m);
    for (int64_t k = 0; k < m; ++k) _correspondences->emplace_back(fi[(size_t) k], mi[(size_t) k], rs[(size_t) k]);
  }

private:
  ContextPtr _ctx;
  int _slice;
  const CloudType* _fixed = nullptr;
  const CloudType* _moving = nullptr;
  bool _fixed_changed_flag = false, _moving_changed_flag = false;
  EstimateType _local_map_in_sensor = EstimateType::Identity();
  CorrespondenceVector* _correspondences = nullptr;
};

// ---------------------------------------------------------------------------------------------
// MultiAlignerBase_ (a1-a9, a11): the whole compute() on the device
// ---------------------------------------------------------------------------------------------
template <int Dim>
class MultiAlignerB200 {
public:
  using CloudType = PointNormalCloud<Dim>;
  using EstimateType = Isometry<Dim>;

  // AlignerSliceProcessor_ (points): aligner_slice_processor.h:56-66,142-150
  struct SliceProcessor {
    // param_finder: SRRG2B_FINDER_NN (kd-tree finders of srrg2_laser_slam_2d) or SRRG2B_FINDER_PROJECTIVE (the
    // projective finder of srrg2_proslam: needs the pinhole intrinsics and the image size below, Dim == 3)
    int finder_kind = SRRG2B_FINDER_NN;
    float finder_fx = 0.f, finder_fy = 0.f, finder_cx = 0.f, finder_cy = 0.f, finder_min_depth = 0.f, finder_max_depth = 1e9f;
    int finder_width = 0, finder_height = 0;
    float finder_max_distance_m = 0.5f, finder_normal_cos = 0.8f;
    int factor = SRRG2B_FACTOR_PLANE;
    int robustifier = SRRG2B_ROB_NONE;
    float robustifier_chi_threshold = 1.f;  // RobustifierBase::param_chi_threshold
    float info_point = 1.f, info_normal = 1.f;
    int param_min_num_correspondences = 0;
    EstimateType robot_in_sensor = EstimateType::Identity();
    const CloudType* fixed = nullptr;
    const CloudType* moving = nullptr;
    bool fixed_changed = false, moving_changed = false;
    CorrespondenceVector correspondences;  // owned by the slice (aligner_slice_processor.h:156)
  };
  // AlignerSliceOdom{2,3}DPrior: aligner_slice_odometry_prior.h:9-45 (2D default information 1e2, 3D 1)
  struct PriorSliceProcessor {
    EstimateType measurement = EstimateType::Identity();
    std::array<float, 6> param_diagonal_info_matrix;
    PriorSliceProcessor() { param_diagonal_info_matrix.fill(Dim == 2 ? 100.f : 1.f); }
    // AlignerSliceOdom{2,3}DPrior::setupFactor (aligner_slice_odometry_prior.cpp:6-37): the measurement is the
    // odometry increment fixed^-1 * moving -- but only from the THIRD call on (`_count > 1`, :8,:25); the first two
    // calls leave it at the identity (the tracker has no previous odometry pose yet).
    void setOdometry(const EstimateType& fixed_pose, const EstimateType& moving_pose) {
      measurement = _count > 1 ? fixed_pose.inverse() * moving_pose : EstimateType::Identity();
      ++_count;
    }
    int count() const { return _count; }

  private:
    int _count = 0;
  };
  // AlignerTerminationCriteriaStandard_: aligner_termination_criteria.h:40-56
  struct TerminationCriteria {
    int param_window_size = 5, param_num_correspondences_range = 20, param_num_inliers_range = 20,
        param_num_outliers_range = 20;
    float param_chi_epsilon = 0.2f;
  };

  // aligner.h:30-35, multi_aligner.h:45-57 (same names and defaults)
  int param_max_iterations = 10;
  int param_min_num_inliers = 10;
  bool param_enable_inlier_only_runs = false;
  bool param_keep_only_inlier_correspondences = false;
  std::shared_ptr<TerminationCriteria> param_termination_criteria;  // null: run all iterations
  int variable = SRRG2B_VAR_SE3_QUAT_RIGHT;                         // MultiAligner3DQR / MultiAligner3D

  explicit MultiAlignerB200(ContextPtr ctx) : _ctx(std::move(ctx)) {}

  // param_slice_processors, in order (multi_aligner.h:34-37)
  int addSliceProcessor(const SliceProcessor& s) {
    _order.push_back({false, (int) _slices.size()});
    _slices.push_back(s);
    return (int) _order.size() - 1;
  }
  int addPriorSliceProcessor(const PriorSliceProcessor& p) {
    _order.push_back({true, (int) _priors.size()});
    _priors.push_back(p);
    return (int) _order.size() - 1;
  }
  SliceProcessor& sliceProcessor(int k) { return _slices.at((size_t) entry(k, false).index); }
  PriorSliceProcessor& priorSliceProcessor(int k) { return _priors.at((size_t) entry(k, true).index); }
  // Aligner_::setFixed / setMoving (aligner.h:46-60), per slice
  void setFixed(int k, const CloudType* c) { auto& s = sliceProcessor(k); s.fixed = c; s.fixed_changed = true; }
  void setMoving(int k, const CloudType* c) { auto& s = sliceProcessor(k); s.moving = c; s.moving_changed = true; }
  void setMovingInFixed(const EstimateType& T) { _moving_in_fixed = T; }
  const EstimateType& movingInFixed() const { return _moving_in_fixed; }
  AlignerStatus status() const { return _status; }
  const IterationStatsVector& iterationStats() const { return _iteration_stats; }
  int numCorrespondences() const {
    int n = 0;
    for (const auto& s : _slices) n += (int) s.correspondences.size();
    return n;
  }

  void compute() {  // multi_aligner_impl.cpp:46-95
    if (_order.empty()) throw std::runtime_error("MultiAlignerB200::compute|no slice processors");
    if ((int) _order.size() > SRRG2B_MAX_SLICES) throw std::runtime_error("MultiAlignerB200::compute|too many slices");
    std::vector<srrg2b_slice> sl(_order.size());
    for (size_t k = 0; k < _order.size(); ++k) {
      srrg2b_slice& d = sl[k];
      std::memset(&d, 0, sizeof(d));
      detail::embed(EstimateType::Identity(), d.robot_in_sensor);
      detail::embed(EstimateType::Identity(), d.prior_measurement);
      if (_order[k].prior) {
        const PriorSliceProcessor& p = _priors[(size_t) _order[k].index];
        d.kind = SRRG2B_SLICE_PRIOR;
        detail::embed(p.measurement, d.prior_measurement);
        for (int i = 0; i < 6; ++i) d.prior_info_diag[i] = p.param_diagonal_info_matrix[(size_t) i];
        continue;
      }
      SliceProcessor& s = _slices[(size_t) _order[k].index];
      if (!s.fixed || !s.moving) throw std::runtime_error("MultiAlignerB200::compute|slice without fixed or moving");
      d.kind = SRRG2B_SLICE_POINTS;
      d.slice_id = (int) k;
      d.min_num_correspondences = s.param_min_num_correspondences;
      detail::embed(s.robot_in_sensor, d.robot_in_sensor);
      if (s.finder_kind == SRRG2B_FINDER_PROJECTIVE && Dim != 3)
        throw std::runtime_error("MultiAlignerB200::compute|the projective finder needs a 3D aligner");
      d.finder.kind = s.finder_kind;
      d.finder.max_distance = s.finder_max_distance_m;
      d.finder.normal_cos = s.finder_normal_cos;
      d.finder.fx = s.finder_fx; d.finder.fy = s.finder_fy; d.finder.cx = s.finder_cx; d.finder.cy = s.finder_cy;
      d.finder.width = s.finder_width; d.finder.height = s.finder_height;
      d.finder.min_depth = s.finder_min_depth; d.finder.max_depth = s.finder_max_depth;
      d.factor.factor = s.factor;
      d.factor.robustifier = s.robustifier;
      d.factor.chi_threshold = s.robustifier_chi_threshold;
      d.factor.info_point = s.info_point;
      d.factor.info_normal = s.info_normal;
      if (s.fixed_changed) {  // fixed first: its index build overlaps the upload of the moving cloud
        const srrg2b_cloud c = detail::describe(*s.fixed);
        _ctx->check(srrg2b_set_cloud(_ctx->get(), SRRG2B_FIXED, (int) k, &c), "MultiAlignerB200::setFixed");
        s.fixed_changed = false;
      }
      if (s.moving_changed) {
        const srrg2b_cloud c = detail::describe(*s.moving);
        _ctx->check(srrg2b_set_cloud(_ctx->get(), SRRG2B_MOVING, (int) k, &c), "MultiAlignerB200::setMoving");
        s.moving_changed = false;
      }
    }
    srrg2b_aligner_params ap;
    std::memset(&ap, 0, sizeof(ap));
    ap.variable = variable;
    ap.max_iterations = param_max_iterations;
    ap.min_num_inliers = param_min_num_inliers;
    ap.enable_inlier_only_runs = param_enable_inlier_only_runs ? 1 : 0;
    ap.keep_only_inlier_correspondences = param_keep_only_inlier_correspondences ? 1 : 0;
    const TerminationCriteria tc = param_termination_criteria ? *param_termination_criteria : TerminationCriteria();
    ap.use_termination_criteria = param_termination_criteria ? 1 : 0;
    ap.window_size = tc.param_window_size;
    ap.num_correspondences_range = tc.param_num_correspondences_range;
    ap.num_inliers_range = tc.param_num_inliers_range;
    ap.num_outliers_range = tc.param_num_outliers_range;
    ap.chi_epsilon = tc.param_chi_epsilon;
    std::vector<srrg2b_iter_stats> st(256);
    int32_t n_stats = (int32_t) st.size(), status = SRRG2B_ALIGNER_FAIL;
    EstimateType T = _moving_in_fixed;
    _ctx->check(srrg2b_icp_run(_ctx->get(), (int) sl.size(), sl.data(), &ap, T.data(), st.data(), &n_stats, &status),
                "MultiAlignerB200::compute");
    _moving_in_fixed = T;
    _status = static_cast<AlignerStatus>(status);
    st.resize((size_t) std::min<int32_t>(n_stats, (int32_t) st.size()));
    _iteration_stats = st;
    // storeCorrespondences() (aligner_slice_processor_impl.cpp:50-74): the slices keep the final lists
    for (size_t k = 0; k < _order.size(); ++k) {
      if (_order[k].prior) continue;
      SliceProcessor& s = _slices[(size_t) _order[k].index];
      const size_t n = s.moving->size();
      std::vector<int32_t> fi(n), mi(n);
      std::vector<float> rs(n);
      int64_t m = 0;
      _ctx->check(srrg2b_get_correspondences(_ctx->get(), (int) k, fi.data(), mi.data(), rs.data(), &m),
                  "MultiAlignerB200::storeCorrespondences");
      s.correspondences.clear();
      s.correspondences.reserve((size_t) m);
      for (int64_t j = 0; j < m; ++j) s.correspondences.emplace_back(fi[(size_t) j], mi[(size_t) j], rs[(size_t) j]);
    }
  }

private:
  struct Entry {
    bool prior;
    int index;
  };
  const Entry& entry(int k, bool prior) const {
    const Entry& e = _order.at((size_t) k);
    if (e.prior != prior) throw std::runtime_error("MultiAlignerB200|slice has the other kind");
    return e;
  }
  ContextPtr _ctx;
  std::vector<Entry> _order;
  std::vector<SliceProcessor> _slices;
  std::vector<PriorSliceProcessor> _priors;
  EstimateType _moving_in_fixed = EstimateType::Identity();
  AlignerStatus _status = AlignerStatus::Fail;
  IterationStatsVector _iteration_stats;
};

// ---------------------------------------------------------------------------------------------
// MultiLoopDetectorBruteForce_ / MultiRelocalizer_ candidate loop:
// R/registration/loop_detector/multi_loop_detector_brute_force_impl.cpp:63-133.  The reference aligns the K
// candidate local maps one after the other with ONE aligner; here every candidate gets a context of its own (created
// on demand, kept between calls), the source map is uploaded and indexed once and lent to all of them, and the K
// runs are in flight together.  Same parameters, same gates, same order of the detected closures.
// ---------------------------------------------------------------------------------------------
template <int Dim>
class LoopDetectorBruteForceB200 {
public:
  using CloudType = PointNormalCloud<Dim>;
  using EstimateType = Isometry<Dim>;
  using SliceProcessor = typename MultiAlignerB200<Dim>::SliceProcessor;

  // multi_loop_detector_brute_force.h:25-40 (same names and defaults)
  int param_relocalize_min_inliers = 500;
  float param_relocalize_max_chi_inliers = 0.005f;
  float param_relocalize_min_inliers_ratio = 0.7f;
  // param_relocalize_aligner: one point slice (finder / factor / robustifier) + the aligner's own parameters
  SliceProcessor aligner_slice;
  int aligner_max_iterations = 10, aligner_min_num_inliers = 10;
  bool aligner_enable_inlier_only_runs = false, aligner_keep_only_inlier_correspondences = false;
  int variable = SRRG2B_VAR_SE3_QUAT_RIGHT;

  struct Hint {  // LoopClosureHint: target local map + initial guess (:64-76)
    const CloudType* local_map = nullptr;
    EstimateType initial_guess = EstimateType::Identity();
  };
  struct Closure {  // what the LoopClosure factor is built from (:113-127)
    int target = -1;  // index of the hint
    EstimateType moving_in_fixed = EstimateType::Identity();
    float chi_inliers = 0.f;
    int64_t num_inliers = 0, num_correspondences = 0;
  };

  explicit LoopDetectorBruteForceB200(int device = 0) : _device(device), _source(std::make_shared<Context>(Dim, device)) {}

  // aligner->setFixed(&source_local_map->dynamic_properties) (:63)
  void setSource(const CloudType* source) { _fixed = source; _fixed_changed = true; }
  const std::vector<Closure>& detectedClosures() const { return _detected; }
  const std::vector<srrg2b_closure_result>& results() const { return _results; }
  // MultiRelocalizer_::compute's choice (R/registration/relocalization/multi_relocalizer_impl.cpp:119-131): among the
  // accepted candidates the one with the smallest chi per inlier, the first one on ties; -1: none (index into
  // detectedClosures()).  Its correspondences stay in its context (storeCorrespondences(), :130).
  int bestDetectedByChi() const {
    int best = -1;
    for (size_t k = 0; k < _detected.size(); ++k)
      if (best < 0 || _detected[k].chi_inliers < _detected[(size_t) best].chi_inliers) best = (int) k;
    return best;
  }

  void compute(const std::vector<Hint>& hints) {
    _detected.clear();
    _results.clear();
    if (!_fixed) throw std::runtime_error("LoopDetectorBruteForceB200::compute|no source local map");
    srrg2b_slice d;
    std::memset(&d, 0, sizeof(d));
    detail::embed(aligner_slice.robot_in_sensor, d.robot_in_sensor);
    detail::embed(EstimateType::Identity(), d.prior_measurement);
    d.kind = SRRG2B_SLICE_POINTS;
    d.slice_id = 0;
    d.min_num_correspondences = aligner_slice.param_min_num_correspondences;
    d.finder.kind = aligner_slice.finder_kind;
    d.finder.max_distance = aligner_slice.finder_max_distance_m;
    d.finder.normal_cos = aligner_slice.finder_normal_cos;
    d.factor.factor = aligner_slice.factor;
    d.factor.robustifier = aligner_slice.robustifier;
    d.factor.chi_threshold = aligner_slice.robustifier_chi_threshold;
    d.factor.info_point = aligner_slice.info_point;
    d.factor.info_normal = aligner_slice.info_normal;
    srrg2b_aligner_params ap;
    std::memset(&ap, 0, sizeof(ap));
    ap.variable = variable;
    ap.max_iterations = aligner_max_iterations;
    ap.min_num_inliers = aligner_min_num_inliers;
    ap.enable_inlier_only_runs = aligner_enable_inlier_only_runs ? 1 : 0;
    ap.keep_only_inlier_correspondences = aligner_keep_only_inlier_correspondences ? 1 : 0;
    ap.window_size = 5; ap.num_correspondences_range = 20; ap.num_inliers_range = 20; ap.num_outliers_range = 20;
    ap.chi_epsilon = 0.2f;
    // candidates without a local map are skipped (:70-72)
    std::vector<int> live;
    for (size_t k = 0; k < hints.size(); ++k) if (hints[k].local_map) live.push_back((int) k);
    if (live.empty()) return;
    if (_fixed_changed) {
      const srrg2b_cloud c = detail::describe(*_fixed);
      _source->check(srrg2b_set_cloud(_source->get(), SRRG2B_FIXED, 0, &c), "LoopDetectorBruteForceB200::setSource");
      // build the index for the detector's finder radius before lending it: a one-point find does that
      const srrg2b_cloud m = detail::describe(*hints[(size_t) live[0]].local_map);
      srrg2b_cloud one = m;
      one.n = std::min<int64_t>(m.n, 1);
      _source->check(srrg2b_set_cloud(_source->get(), SRRG2B_MOVING, 0, &one), "LoopDetectorBruteForceB200::setSource");
      int64_t n = 0;
      EstimateType I = EstimateType::Identity();
      _source->check(srrg2b_find_correspondences(_source->get(), 0, I.data(), &d.finder, nullptr, nullptr, nullptr, &n),
                     "LoopDetectorBruteForceB200::setSource");
    }
    while (_candidates.size() < live.size()) _candidates.push_back(std::make_shared<Context>(Dim, _device));
    std::vector<srrg2b_ctx*> ctxs(live.size());
    std::vector<float> guesses(live.size() * EstimateType::N);
    for (size_t i = 0; i < live.size(); ++i) {
      const Hint& h = hints[(size_t) live[i]];
      Context& c = *_candidates[i];
      if (_fixed_changed || i >= _lent) c.check(srrg2b_share_fixed(c.get(), 0, _source->get(), 0), "LoopDetectorBruteForceB200::share");
      const srrg2b_cloud m = detail::describe(*h.local_map);  // aligner->setMoving (:76)
      c.check(srrg2b_set_cloud(c.get(), SRRG2B_MOVING, 0, &m), "LoopDetectorBruteForceB200::setMoving");
      std::memcpy(&guesses[i * EstimateType::N], h.initial_guess.data(), sizeof(float) * EstimateType::N);
      ctxs[i] = c.get();
    }
    _lent = std::max(_lent, live.size());
    _fixed_changed = false;
    srrg2b_closure_params cp;
    cp.relocalize_min_inliers = param_relocalize_min_inliers;
    cp.relocalize_max_chi_inliers = param_relocalize_max_chi_inliers;
    cp.relocalize_min_inliers_ratio = param_relocalize_min_inliers_ratio;
    _results.resize(live.size());
    _candidates[0]->check(srrg2b_closure_batch(ctxs.data(), (int) ctxs.size(), 1, &d, &ap, guesses.data(), &cp, _results.data()),
                          "LoopDetectorBruteForceB200::compute");
    for (size_t i = 0; i < live.size(); ++i) {
      const srrg2b_closure_result& r = _results[i];
      if (r.verdict != SRRG2B_CLOSURE_ACCEPT) continue;
      Closure cl;
      cl.target = live[i];
      std::memcpy(cl.moving_in_fixed.data(), r.moving_in_fixed, sizeof(float) * EstimateType::N);
      cl.chi_inliers = r.chi_inliers;
      cl.num_inliers = r.num_inliers;
      cl.num_correspondences = r.num_correspondences;
      _detected.push_back(cl);
    }
  }

private:
  int _device;
  ContextPtr _source;
  std::vector<ContextPtr> _candidates;
  size_t _lent = 0;
  const CloudType* _fixed = nullptr;
  bool _fixed_changed = false;
  std::vector<Closure> _detected;
  std::vector<srrg2b_closure_result> _results;
};

// ---------------------------------------------------------------------------------------------
// Solver on a pose graph (a10): SE{2,3}PosePoseGeodesicErrorFactor between VariableSE2RightAD /
// VariableSE3QuaternionRightAD poses (R/registration/loop_closure.h:110-111, R/mapping/local_map.h:64,75).
// The context's dim selects the group: 3 -> 4x4 poses, 6x6 informations; 2 -> 3x3 poses, 3x3 informations.
// ---------------------------------------------------------------------------------------------
class PoseGraphSolverB200 {
public:
  int param_max_iterations = 10;      // Solver::param_max_iterations
  double param_dx_epsilon = 1e-6;     // stop when the largest perturbation component falls below
  int param_max_cg_iterations = 0;    // 0 = library default
  bool param_damped = true;           // Levenberg-Marquardt guard + inexact solves (srrg2b_pgo_optimize); false: plain GN
  double param_cg_tolerance = 0.0;    // plain GN only; 0 = library default
  enum Status { Error = 0, Success = 1 };

  PoseGraphSolverB200(ContextPtr ctx, int dim = 3) : _ctx(std::move(ctx)), _m((dim + 1) * (dim + 1)), _b(dim == 3 ? 36 : 9) {
    if (dim != 2 && dim != 3) throw std::runtime_error("PoseGraphSolverB200|dim must be 2 or 3");
  }

  // setGraph(): poses row-major (dim+1)^2 each, fixed mask, factor pairs (i, j), measurements, informations
  void setGraph(std::vector<float> poses, std::vector<uint8_t> fixed, std::vector<int32_t> ij, std::vector<float> Z,
                std::vector<float> Omega) {
    _poses = std::move(poses); _fixed = std::move(fixed); _ij = std::move(ij); _Z = std::move(Z); _Omega = std::move(Omega);
    if (_poses.size() % _m || _fixed.size() != _poses.size() / _m || _ij.size() % 2 || _Z.size() != _ij.size() / 2 * _m ||
        _Omega.size() != _ij.size() / 2 * _b)
      throw std::runtime_error("PoseGraphSolverB200::setGraph|inconsistent array sizes");
  }
  void compute() {
    _ctx->check(srrg2b_pgo_upload(_ctx->get(), (int64_t) _fixed.size(), _poses.data(), _fixed.data(), (int64_t) _ij.size() / 2,
                                  _ij.data(), _Z.data(), _Omega.data()),
                "PoseGraphSolverB200::compute");
    _stats.clear();
    _status = Error;
    if (param_damped) {
      _stats.resize((size_t) std::max(param_max_iterations, 1));
      int32_t n = 0;
      _ctx->check(srrg2b_pgo_optimize(_ctx->get(), param_max_iterations, param_dx_epsilon, param_max_cg_iterations, _stats.data(), &n),
                  "PoseGraphSolverB200::compute");
      _stats.resize((size_t) n);
    } else {
      for (int it = 0; it < param_max_iterations; ++it) {
        srrg2b_pgo_stats st;
        _ctx->check(srrg2b_pgo_iterate(_ctx->get(), param_max_cg_iterations, param_cg_tolerance, &st), "PoseGraphSolverB200::compute");
        _stats.push_back(st);
        if (st.dx_norm_inf < param_dx_epsilon) break;
      }
    }
    _ctx->check(srrg2b_pgo_download(_ctx->get(), _poses.data()), "PoseGraphSolverB200::compute");
    _status = Success;
  }
  Status status() const { return _status; }
  const std::vector<srrg2b_pgo_stats>& iterationStats() const { return _stats; }
  const std::vector<float>& poses() const { return _poses; }

private:
  ContextPtr _ctx;
  size_t _m, _b;
  std::vector<float> _poses, _Z, _Omega;
  std::vector<uint8_t> _fixed;
  std::vector<int32_t> _ij;
  std::vector<srrg2b_pgo_stats> _stats;
  Status _status = Error;
};

}  // namespace srrg2b
#endif
