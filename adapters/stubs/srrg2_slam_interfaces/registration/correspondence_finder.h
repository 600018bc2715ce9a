// STAND-IN for R/registration/correspondence_finder.h:41-124 -- the members an adapter subclass touches, with the
// reference's names; the real header replaces this file when the adapters are built inside the reference tree.
#pragma once
#include "srrg2_core/stub.h"

namespace srrg2_slam_interfaces {

class CorrespondenceFinderBase : public srrg2_core::Configurable {  // :41-56
public:
  void setCorrespondences(srrg2_core::CorrespondenceVector* c) { _correspondences = c; }
  virtual void compute() = 0;

protected:
  srrg2_core::CorrespondenceVector* _correspondences = nullptr;
};

template <typename EstimateType_, typename FixedType_, typename MovingType_>
class CorrespondenceFinder_ : public CorrespondenceFinderBase {  // :66-124
public:
  using EstimateType = EstimateType_;
  using FixedType = FixedType_;
  using MovingType = MovingType_;
  void setFixed(FixedType* f) { _fixed = f; _fixed_changed_flag = true; }                       // :80-85
  void setMoving(MovingType* m) { _moving = m; _moving_changed_flag = true; }                   // :86-91
  virtual void setLocalMapInSensor(const EstimateType& T) {                                     // :111-114
    _local_map_in_sensor = T;
    _local_map_in_sensor_changed_flag = true;
  }

protected:
  FixedType* _fixed = nullptr;
  MovingType* _moving = nullptr;
  EstimateType _local_map_in_sensor = EstimateType::Identity();                                 // :122-123
  bool _fixed_changed_flag = true, _moving_changed_flag = true, _local_map_in_sensor_changed_flag = true;
};

}  // namespace srrg2_slam_interfaces
