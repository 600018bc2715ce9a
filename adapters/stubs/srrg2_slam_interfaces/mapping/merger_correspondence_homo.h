// STAND-IN for R/mapping/merger.h:14-177 and R/mapping/merger_correspondence_homo.h:8-33: the members a merger subclass
// uses, with the reference's names and parameter defaults.  Shapes only; not upstream code.
#pragma once
#include "srrg2_core/stub.h"

namespace srrg2_slam_interfaces {

class MergerBase : public srrg2_core::Configurable {  // merger.h:14-40
public:
  enum Status { Error = 0x0, Initializing = 0x1, Success = 0x2 };  // :19-23
  Status status() const { return _status; }

protected:
  Status _status = Error;
};

template <typename EstimateType_, typename FixedSceneType_, typename MovingMeasurementType_>
class Merger_ : public MergerBase {  // merger.h:49-112
public:
  using EstimateType = EstimateType_;
  void setMeasurementInScene(const EstimateType& measurement_in_scene_) { _measurement_in_scene = measurement_in_scene_; }  // :59-62
  void setScene(FixedSceneType_* scene_) { _scene = scene_; _scene_changed_flag = true; }                                    // :76-79
  void setMeasurement(const MovingMeasurementType_* measurement_) { _measurement = measurement_; }                           // :85-88
  virtual void compute() = 0;                                                                                               // :97

protected:
  const MovingMeasurementType_* _measurement = nullptr;
  FixedSceneType_* _scene = nullptr;
  EstimateType _measurement_in_scene = EstimateType::Identity();
  bool _scene_changed_flag = true;
};

template <typename EstimateType_, typename FixedSceneType_, typename MovingMeasurementType_>
class MergerCorrespondence_ : public Merger_<EstimateType_, FixedSceneType_, MovingMeasurementType_> {  // merger.h:116-177
public:
  PARAM(srrg2_core::PropertyInt, target_number_of_merges, "target number of points to merge", 200, nullptr);  // :126-131
  void setCorrespondences(const srrg2_core::CorrespondenceVector* correspondences_) { _correspondences = correspondences_; }  // :145-148

protected:
  const srrg2_core::CorrespondenceVector* _correspondences = nullptr;
};

template <typename EstimateType_, typename SceneType_>
class MergerCorrespondenceHomo_ : public MergerCorrespondence_<EstimateType_, SceneType_, SceneType_> {  // merger_correspondence_homo.h:8-33
public:
  PARAM(srrg2_core::PropertyFloat, maximum_response, "maximum permitted correspondence response for merging a point", 50.f, nullptr);  // :22-26
  PARAM(srrg2_core::PropertyFloat, maximum_distance_geometry_squared, "maximum distance in geometry in meters (squared)", 0.25f, nullptr);  // :27-31
  void compute() override {}
};

}  // namespace srrg2_slam_interfaces
