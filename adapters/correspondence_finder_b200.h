// correspondence_finder_b200.h -- reference-side adapter: a CorrespondenceFinder_ subclass (row a3) that a maintainer
// adds next to the kd-tree / projective finders and registers with BOSS; selected by class name in the configuration,
// no caller changes (slot: AlignerSliceProcessor_::param_finder, R/registration/aligners/aligner_slice_processor.h:56-60).
// Replaces: CorrespondenceFinder_::compute(), R/registration/correspondence_finder.h:56 (called at
// R/registration/aligners/aligner_slice_processor_impl.cpp:43).
// Build inside the reference tree against the real headers; `g++ -fsyntax-only -I include -I adapters/stubs` checks it here.
#pragma once
#include <srrg2_slam_interfaces/registration/correspondence_finder.h>

#include "srrg2b_flatten.h"

namespace srrg2_slam_interfaces {

template <typename EstimateType_, typename CloudType_>  // CloudType_ = PointNormal{2,3}fVectorCloud
class CorrespondenceFinderB200_ : public CorrespondenceFinder_<EstimateType_, CloudType_, CloudType_> {
public:
  using BaseType = CorrespondenceFinder_<EstimateType_, CloudType_, CloudType_>;
  static constexpr int Dim = EstimateType_::Dim;
  PARAM(srrg2_core::PropertyFloat, max_distance_m, "maximum distance of a correspondence [m]", 0.5f, nullptr);
  PARAM(srrg2_core::PropertyFloat, normal_cos, "minimum cosine between the normals", 0.8f, nullptr);
  PARAM(srrg2_core::PropertyInt, device, "CUDA device", 0, nullptr);
  // projective association (the srrg2_proslam cue): set projective and the pinhole model of the fixed frame
  PARAM(srrg2_core::PropertyBool, projective, "associate through the index image of the fixed cloud", false, nullptr);
  PARAM(srrg2_core::PropertyFloat, fx, "", 0.f, nullptr);
  PARAM(srrg2_core::PropertyFloat, fy, "", 0.f, nullptr);
  PARAM(srrg2_core::PropertyFloat, cx, "", 0.f, nullptr);
  PARAM(srrg2_core::PropertyFloat, cy, "", 0.f, nullptr);
  PARAM(srrg2_core::PropertyInt, image_cols, "", 0, nullptr);
  PARAM(srrg2_core::PropertyInt, image_rows, "", 0, nullptr);

  ~CorrespondenceFinderB200_() override { if (_ctx) srrg2b_ctx_destroy(_ctx); }

  void compute() override {
    if (!this->_fixed || !this->_moving || !this->_correspondences)
      throw std::runtime_error("CorrespondenceFinderB200_::compute|fixed, moving or correspondences not set");
    if (!_ctx) srrg2b_adapters::check(nullptr, srrg2b_ctx_create(Dim, param_device.value(), &_ctx) == SRRG2B_OK ? SRRG2B_OK : SRRG2B_ERR_CUDA,
                                      "CorrespondenceFinderB200_|no usable CUDA device (there is no CPU fallback)");
    // change detection is by the setters' flags, not by content (correspondence_finder.h:80-91)
    if (this->_fixed_changed_flag) { upload(SRRG2B_FIXED, *this->_fixed); this->_fixed_changed_flag = false; }
    if (this->_moving_changed_flag) { upload(SRRG2B_MOVING, *this->_moving); this->_moving_changed_flag = false; }
    srrg2b_finder_params fp;
    std::memset(&fp, 0, sizeof(fp));
    fp.kind = param_projective.value() ? SRRG2B_FINDER_PROJECTIVE : SRRG2B_FINDER_NN;
    fp.max_distance = param_max_distance_m.value();
    fp.normal_cos = param_normal_cos.value();
    fp.fx = param_fx.value(); fp.fy = param_fy.value(); fp.cx = param_cx.value(); fp.cy = param_cy.value();
    fp.width = param_image_cols.value(); fp.height = param_image_rows.value();
    fp.min_depth = 0.f; fp.max_depth = 1e9f;
    float S[16];
    srrg2b_adapters::to_row_major(this->_local_map_in_sensor, S);
    const size_t n = this->_moving->size();
    _fi.resize(n); _mi.resize(n); _rs.resize(n);
    int64_t m = 0;
    srrg2b_adapters::check(_ctx, srrg2b_find_correspondences(_ctx, 0, S, &fp, _fi.data(), _mi.data(), _rs.data(), &m),
                           "CorrespondenceFinderB200_::compute");
    this->_correspondences->clear();
    this->_correspondences->reserve((size_t) m);
    for (int64_t k = 0; k < m; ++k)  // Correspondence(fixed, moving, response), ascending moving index
      this->_correspondences->emplace_back(_fi[(size_t) k], _mi[(size_t) k], _rs[(size_t) k]);
    this->_local_map_in_sensor_changed_flag = false;
  }

private:
  void upload(int slot, const CloudType_& cloud) {
    const srrg2b_adapters::FlatCloud f = srrg2b_adapters::flatten(cloud);
    const srrg2b_cloud c = f.describe();
    srrg2b_adapters::check(_ctx, srrg2b_set_cloud(_ctx, slot, 0, &c), "CorrespondenceFinderB200_::upload");
  }
  srrg2b_ctx* _ctx = nullptr;
  std::vector<int32_t> _fi, _mi;
  std::vector<float> _rs;
};

using CorrespondenceFinderB2002D = CorrespondenceFinderB200_<srrg2_core::Isometry2f, srrg2_core::PointNormal2fVectorCloud>;
using CorrespondenceFinderB2003D = CorrespondenceFinderB200_<srrg2_core::Isometry3f, srrg2_core::PointNormal3fVectorCloud>;
// in the library's registerTypes() (pattern R/instances.cpp:21-23):
//   BOSS_REGISTER_CLASS(CorrespondenceFinderB2002D); BOSS_REGISTER_CLASS(CorrespondenceFinderB2003D);

}  // namespace srrg2_slam_interfaces
