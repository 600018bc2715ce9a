// multi_aligner_b200.h -- reference-side adapter: the whole MultiAlignerBase_::compute() (rows a1-a9, a11) on the GPU.
// The loop detectors / relocaliser are templated on the concrete MultiAligner2D / 3DQR and reach into
// param_slice_processors, param_solver and slice->param_robustifier
// (R/registration/loop_detector/multi_loop_detector_hbst_impl.cpp:208-227,271,337), so the GPU aligner DERIVES from
// MultiAlignerBase_<V> and overrides the virtual compute() (R/registration/aligners/multi_aligner.h:95):
// Replaces: R/registration/aligners/multi_aligner_impl.cpp:46-263.
#pragma once
#include <srrg2_slam_interfaces/registration/aligners/multi_aligner.h>

#include "srrg2b_flatten.h"

namespace srrg2_slam_interfaces {

template <typename VariableType_, typename CloudType_>
class MultiAlignerB200_ : public MultiAlignerBase_<VariableType_> {
public:
  using BaseType = MultiAlignerBase_<VariableType_>;
  using EstimateType = typename BaseType::EstimateType;
  using PointSlice = AlignerSliceProcessor_<EstimateType, CloudType_>;
  using PriorSlice = AlignerSliceProcessorPrior_<EstimateType>;
  static constexpr int Dim = EstimateType::Dim;
  PARAM(srrg2_core::PropertyInt, device, "CUDA device", 0, nullptr);
  PARAM(srrg2_core::PropertyFloat, max_distance_m, "finder: maximum distance of a correspondence [m]", 0.5f, nullptr);
  PARAM(srrg2_core::PropertyFloat, normal_cos, "finder: minimum cosine between the normals", 0.8f, nullptr);
  PARAM(srrg2_core::PropertyBool, quaternion_variable, "VariableSE3QuaternionRightAD (else Euler)", true, nullptr);

  ~MultiAlignerB200_() override { if (_ctx) srrg2b_ctx_destroy(_ctx); }

  // The aligner's configuration as the C ABI takes it: one srrg2b_slice per slice processor, in order.  No cloud is
  // touched (the loop-detector adapter describes the slices once and runs them on its own contexts).
  void describeSlices(std::vector<srrg2b_slice>& slices) {
    const size_t n_slices = this->param_slice_processors.size();
    if (n_slices == 0 || n_slices > SRRG2B_MAX_SLICES) throw std::runtime_error("MultiAlignerB200_|bad number of slice processors");
    slices.assign(n_slices, srrg2b_slice{});
    for (size_t s = 0; s < n_slices; ++s) {
      srrg2b_slice& d = slices[s];
      std::memset(&d, 0, sizeof(d));
      srrg2b_adapters::embed16(EstimateType::Identity(), d.robot_in_sensor);
      srrg2b_adapters::embed16(EstimateType::Identity(), d.prior_measurement);
      auto* base = this->param_slice_processors.value(s).get();
      if (base->isPrior()) {
        // odometry / motion-model priors: the reference slice refreshes its factor in setupFactor() -- including the
        // `_count > 1` rule of AlignerSliceOdom{2,3}DPrior (aligner_slice_odometry_prior.cpp:8,25) and the inverse
        // motion of AlignerSliceMotionModel_ (aligner_slice_motion_model.hpp:69-78); the adapter only reads the result.
        // The library lets the priors overwrite the initial guess in slice order (multi_aligner_impl.cpp:130-141).
        auto* p = dynamic_cast<PriorSlice*>(base);
        if (!p) throw std::runtime_error("MultiAlignerB200_|unknown prior slice type");
        p->setupFactor();
        d.kind = SRRG2B_SLICE_PRIOR;
        srrg2b_adapters::embed16(p->measurement(), d.prior_measurement);
        for (int k = 0; k < 6; ++k) d.prior_info_diag[k] = p->diagonalInfo()[(size_t) k];
        continue;
      }
      auto* sp = dynamic_cast<PointSlice*>(base);
      if (!sp) throw std::runtime_error("MultiAlignerB200_|unknown point slice type");
      d.kind = SRRG2B_SLICE_POINTS;
      d.slice_id = (int) s;
      d.min_num_correspondences = sp->param_min_num_correspondences.value();
      srrg2b_adapters::embed16(sp->robotInSensor(), d.robot_in_sensor);
      d.finder.kind = SRRG2B_FINDER_NN;
      d.finder.max_distance = param_max_distance_m.value();
      d.finder.normal_cos = param_normal_cos.value();
      d.factor.factor = sp->factorKind();
      d.factor.info_point = 1.f;
      d.factor.info_normal = 1.f;
      // robustifier kind + RobustifierBase::param_chi_threshold (aligner_slice_processor_base.h:34-38)
      const auto& rob = sp->param_robustifier.value();
      d.factor.robustifier = SRRG2B_ROB_NONE;
      if (rob) {
        d.factor.chi_threshold = rob->param_chi_threshold.value();
        if (dynamic_cast<srrg2_solver::RobustifierCauchy*>(rob.get())) d.factor.robustifier = SRRG2B_ROB_CAUCHY;
        else if (dynamic_cast<srrg2_solver::RobustifierClamp*>(rob.get())) d.factor.robustifier = SRRG2B_ROB_CLAMP;
        else d.factor.robustifier = SRRG2B_ROB_SATURATED;
      }
    }
  }

  // AlignerBase / MultiAlignerBase_ / termination-criterion parameters (aligner.h:30-35, multi_aligner.h:45-57)
  srrg2b_aligner_params alignerParams() const {
    srrg2b_aligner_params ap;
    std::memset(&ap, 0, sizeof(ap));
    // (dim 2 contexts use VariableSE2Right whatever this field says)
    ap.variable = param_quaternion_variable.value() ? SRRG2B_VAR_SE3_QUAT_RIGHT : SRRG2B_VAR_SE3_EULER_RIGHT;
    ap.max_iterations = this->param_max_iterations.value();
    ap.min_num_inliers = this->param_min_num_inliers.value();
    ap.enable_inlier_only_runs = this->param_enable_inlier_only_runs.value() ? 1 : 0;
    ap.keep_only_inlier_correspondences = this->param_keep_only_inlier_correspondences.value() ? 1 : 0;
    const auto& tc = this->param_termination_criteria.value();
    ap.use_termination_criteria = tc ? 1 : 0;
    ap.window_size = tc ? tc->param_window_size.value() : 5;
    ap.num_correspondences_range = tc ? tc->param_num_correspondences_range.value() : 20;
    ap.num_inliers_range = tc ? tc->param_num_inliers_range.value() : 20;
    ap.num_outliers_range = tc ? tc->param_num_outliers_range.value() : 20;
    ap.chi_epsilon = tc ? tc->param_chi_epsilon.value() : 0.2f;
    return ap;
  }

  // the point slice behind slot s (null for a prior slice): its clouds are bound by Aligner::setFixed / setMoving
  PointSlice* pointSlice(size_t s) const { return dynamic_cast<PointSlice*>(this->param_slice_processors.value(s).get()); }

  // The library context of this aligner (created on first use): the scene adapters (scene_b200.h) keep the local map
  // on the same device context, so that a clipped scene becomes a slice's moving cloud without crossing PCIe.
  srrg2b_ctx* context() {
    if (!_ctx) srrg2b_adapters::check(nullptr, srrg2b_ctx_create(Dim, param_device.value(), &_ctx) == SRRG2B_OK ? SRRG2B_OK : SRRG2B_ERR_CUDA,
                                      "MultiAlignerB200_|no usable CUDA device (there is no CPU fallback)");
    return _ctx;
  }
  // SceneClipperRangeB200_::compute() has written slice s's moving cloud on the device (n_points of them): compute()
  // neither expects nor uploads a host cloud for it until setMovingResident(s, -1).
  void setMovingResident(size_t s, int64_t n_points) {
    if (_resident.size() <= s) _resident.resize(s + 1, -1);
    _resident[s] = n_points;
  }
  int64_t movingResident(size_t s) const { return s < _resident.size() ? _resident[s] : -1; }

  void compute() override {
    context();
    std::vector<srrg2b_slice> slices;
    describeSlices(slices);
    const size_t n_slices = slices.size();
    for (size_t s = 0; s < n_slices; ++s) {
      if (slices[s].kind != SRRG2B_SLICE_POINTS) continue;
      PointSlice* sp = pointSlice(s);
      const bool resident = movingResident(s) >= 0;
      if (!sp->fixed() || (!resident && !sp->moving())) throw std::runtime_error("MultiAlignerB200_::compute|slice without fixed or moving");
      if (sp->fixedChanged()) upload(SRRG2B_FIXED, (int) s, *sp->fixed());
      if (!resident && sp->movingChanged()) upload(SRRG2B_MOVING, (int) s, *sp->moving());
      sp->clearChanged();
    }
    const srrg2b_aligner_params ap = alignerParams();
    float T[16];
    srrg2b_adapters::to_row_major(this->movingInFixed(), T);
    std::vector<srrg2b_iter_stats> st(256);
    int32_t n_stats = (int32_t) st.size(), status = SRRG2B_ALIGNER_FAIL;
    srrg2b_adapters::check(_ctx, srrg2b_icp_run(_ctx, (int) n_slices, slices.data(), &ap, T, st.data(), &n_stats, &status),
                           "MultiAlignerB200_::compute");
    this->setMovingInFixed(srrg2b_adapters::from_row_major<EstimateType>(T));
    this->_status = static_cast<AlignerBase::Status>(status);  // aligner.h:23-28, same numeric values
    this->_iteration_stats.clear();
    for (int32_t k = 0; k < n_stats && k < (int32_t) st.size(); ++k) {
      srrg2_solver::IterationStats is;
      is.iteration = st[(size_t) k].iteration;
      is.num_inliers = (int) st[(size_t) k].num_inliers;
      is.num_outliers = (int) st[(size_t) k].num_outliers;
      is.num_suppressed = (int) st[(size_t) k].num_suppressed;
      is.chi_inliers = (float) st[(size_t) k].chi_inliers;
      is.chi_outliers = (float) st[(size_t) k].chi_outliers;
      this->_iteration_stats.push_back(is);
    }
    // storeCorrespondences() / merger hand-off (aligner_slice_processor_impl.cpp:50-74): the slices keep the lists
    for (size_t s = 0; s < n_slices; ++s) {
      if (slices[s].kind != SRRG2B_SLICE_POINTS) continue;
      auto* sp = static_cast<PointSlice*>(this->param_slice_processors.value(s).get());
      const size_t n = movingResident(s) >= 0 ? (size_t) movingResident(s) : sp->moving()->size();
      _fi.resize(n); _mi.resize(n); _rs.resize(n);
      int64_t m = 0;
      srrg2b_adapters::check(_ctx, srrg2b_get_correspondences(_ctx, (int) s, _fi.data(), _mi.data(), _rs.data(), &m),
                             "MultiAlignerB200_::storeCorrespondences");
      auto& out = sp->correspondences();
      out.clear();
      out.reserve((size_t) m);
      for (int64_t k = 0; k < m; ++k) out.emplace_back(_fi[(size_t) k], _mi[(size_t) k], _rs[(size_t) k]);
    }
  }

private:
  void upload(int slot, int slice, const CloudType_& cloud) {
    const srrg2b_adapters::FlatCloud f = srrg2b_adapters::flatten(cloud);
    const srrg2b_cloud c = f.describe();
    srrg2b_adapters::check(_ctx, srrg2b_set_cloud(_ctx, slot, slice, &c), "MultiAlignerB200_::upload");
  }
  srrg2b_ctx* _ctx = nullptr;
  std::vector<int64_t> _resident;
  std::vector<int32_t> _fi, _mi;
  std::vector<float> _rs;
};

using MultiAligner2DB200 = MultiAlignerB200_<srrg2_solver::VariableSE2RightAD, srrg2_core::PointNormal2fVectorCloud>;
using MultiAligner3DQRB200 = MultiAlignerB200_<srrg2_solver::VariableSE3QuaternionRightAD, srrg2_core::PointNormal3fVectorCloud>;
// registerTypes(): BOSS_REGISTER_CLASS(MultiAligner2DB200); BOSS_REGISTER_CLASS(MultiAligner3DQRB200);

}  // namespace srrg2_slam_interfaces
