// scene_b200.h -- reference-side adapters for the steps either side of the aligner: the local map
// stays in HBM between frames.
//   SceneClipperRangeB200_          a SceneClipper_ (R/mapping/scene_clipper.h:16-122) whose compute() clips the resident
//                                   scene on the device straight into the aligner slice's moving cloud
//                                   (replaces the host clip TrackerSliceProcessor_::clip drives,
//                                   R/trackers/tracker_slice_processor_impl.cpp:194-205)
//   MergerCorrespondenceHomoB200_   MergerCorrespondenceHomo_::compute (R/mapping/merger_correspondence_homo_impl.cpp:11-126)
//                                   on that scene, with the correspondences the aligner left on the device
// Both work on the aligner adapter's context (MultiAlignerB200_::context()).
#pragma once
#include <srrg2_slam_interfaces/mapping/merger_correspondence_homo.h>
#include <srrg2_slam_interfaces/mapping/scene_clipper.h>

#include "multi_aligner_b200.h"

namespace srrg2_slam_interfaces {

template <typename EstimateType_, typename SceneType_, typename AlignerType_>
class SceneClipperRangeB200_ : public SceneClipper_<EstimateType_, SceneType_> {
public:
  using BaseType = SceneClipper_<EstimateType_, SceneType_>;
  PARAM(srrg2_core::PropertyFloat, range, "points farther than this from the robot are clipped away [m]", 10.f, nullptr);
  PARAM(srrg2_core::PropertyInt, scene_id, "handle of the resident scene in the library context", 0, nullptr);
  PARAM(srrg2_core::PropertyInt, aligner_slice, "aligner slice whose moving cloud the clipped scene becomes", 0, nullptr);

  void setAligner(std::shared_ptr<AlignerType_> aligner_) { _aligner = std::move(aligner_); }
  // the merger changed the scene on the device, or the tracker switched local maps: upload again at the next compute()
  void invalidateScene() { _uploaded = nullptr; }

  void compute() override {
    if (!_aligner) throw std::runtime_error("SceneClipperRangeB200_::compute|no aligner");
    if (!this->_full_scene) throw std::runtime_error("SceneClipperRangeB200_::compute|no scene");
    srrg2b_ctx* ctx = _aligner->context();
    const int scene = param_scene_id.value(), slice = param_aligner_slice.value();
    if (_uploaded != this->_full_scene) {  // setFullScene(scene): once per local map
      const srrg2b_adapters::FlatCloud f = srrg2b_adapters::flatten(*this->_full_scene);
      const srrg2b_cloud c = f.describe();
      srrg2b_adapters::check(ctx, srrg2b_scene_set(ctx, scene, &c), "SceneClipperRangeB200_::setFullScene");
      _uploaded = this->_full_scene;
    }
    float T[16] = {0};
    srrg2b_adapters::to_row_major(this->_local_map_in_robot, T);  // scene (local map) in robot, scene_clipper.h:80-83
    int64_t n = 0;
    srrg2b_adapters::check(ctx, srrg2b_scene_clip(ctx, scene, slice, T, param_range.value(), &n), "SceneClipperRangeB200_::compute");
    _indices32.resize((size_t) n);
    if (n > 0) srrg2b_adapters::check(ctx, srrg2b_scene_clip_indices(ctx, slice, _indices32.data()), "SceneClipperRangeB200_::globalIndices");
    _global_indices.assign(_indices32.begin(), _indices32.end());
    _aligner->setMovingResident((size_t) slice, n);
    // (the host cloud given to setClippedSceneInRobot is left untouched: the clipped scene lives on the device, in the
    // robot frame, as the aligner slice's moving cloud; globalIndices() names its points in the full scene)
    this->_status = n > 0 ? BaseType::Successful : BaseType::Ready;  // :22-26
  }
  const std::vector<int> globalIndices() const override { return _global_indices; }  // :104-107

private:
  std::shared_ptr<AlignerType_> _aligner;
  const SceneType_* _uploaded = nullptr;
  std::vector<int32_t> _indices32;
  std::vector<int> _global_indices;
};

template <typename EstimateType_, typename SceneType_, typename AlignerType_>
class MergerCorrespondenceHomoB200_ : public MergerCorrespondenceHomo_<EstimateType_, SceneType_> {
public:
  using BaseType = MergerCorrespondenceHomo_<EstimateType_, SceneType_>;
  PARAM(srrg2_core::PropertyInt, scene_id, "handle of the resident scene in the library context", 0, nullptr);
  PARAM(srrg2_core::PropertyInt, aligner_slice, "aligner slice that holds the measurement (fixed) and the correspondences", 0, nullptr);

  void setAligner(std::shared_ptr<AlignerType_> aligner_) { _aligner = std::move(aligner_); }
  int64_t numMerged() const { return _merged; }
  int64_t numAdded() const { return _added; }

  void compute() override {  // merger_correspondence_homo_impl.cpp:11-126
    if (!_aligner) throw std::runtime_error("MergerCorrespondenceHomoB200_::compute|no aligner");
    srrg2b_ctx* ctx = _aligner->context();
    srrg2b_merge_params mp;
    mp.maximum_response = this->param_maximum_response.value();
    mp.maximum_distance_geometry_squared = this->param_maximum_distance_geometry_squared.value();
    mp.target_number_of_merges = this->param_target_number_of_merges.value();
    mp.without_correspondences = this->_correspondences ? 0 : 1;  // :31-42: no correspondences set -> add everything
    float T[16] = {0};
    srrg2b_adapters::to_row_major(this->_measurement_in_scene, T);
    srrg2b_adapters::check(ctx, srrg2b_scene_merge(ctx, param_scene_id.value(), param_aligner_slice.value(), T, &mp, &_merged, &_added),
                           "MergerCorrespondenceHomoB200_::compute");
    this->_status = MergerBase::Success;
  }

  // the scene back on the host (when the local map is serialised or handed to a host-side consumer): fills *_scene
  void downloadScene() {
    if (!_aligner || !this->_scene) throw std::runtime_error("MergerCorrespondenceHomoB200_::downloadScene|no aligner or scene");
    srrg2b_ctx* ctx = _aligner->context();
    using Point = typename SceneType_::value_type;
    constexpr int D = Point::Dim;
    int64_t n = 0;
    srrg2b_adapters::check(ctx, srrg2b_scene_get(ctx, param_scene_id.value(), nullptr, nullptr, nullptr, &n), "MergerCorrespondenceHomoB200_::downloadScene");
    std::vector<float> c((size_t) n * D), nr((size_t) n * D);
    std::vector<uint8_t> v((size_t) n);
    srrg2b_adapters::check(ctx, srrg2b_scene_get(ctx, param_scene_id.value(), c.data(), nr.data(), v.data(), &n), "MergerCorrespondenceHomoB200_::downloadScene");
    this->_scene->resize((size_t) n);
    for (int64_t i = 0; i < n; ++i) {
      Point& p = (*this->_scene)[(size_t) i];
      for (int k = 0; k < D; ++k) { p._c[(size_t) k] = c[(size_t) i * D + k]; p._n[(size_t) k] = nr[(size_t) i * D + k]; }
      p.status = v[(size_t) i] ? srrg2_core::Valid : srrg2_core::Invalid;
    }
  }

private:
  std::shared_ptr<AlignerType_> _aligner;
  int64_t _merged = 0, _added = 0;
};

}  // namespace srrg2_slam_interfaces
