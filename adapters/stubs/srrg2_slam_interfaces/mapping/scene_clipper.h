// STAND-IN for R/mapping/scene_clipper.h:16-122: the members a SceneClipper subclass uses, with the reference's names.
// Shapes only; not upstream code.
#pragma once
#include <string>
#include <vector>

#include "srrg2_core/stub.h"

namespace srrg2_slam_interfaces {

template <typename EstimateType_, typename SceneType_>
class SceneClipper_ : public srrg2_core::Configurable {
public:
  enum Status { Error = 0x0, Successful = 0x1, Ready = 0x2 };  // :22-26
  using EstimateType = EstimateType_;
  using SceneType = SceneType_;
  virtual void compute() = 0;                                                        // :36
  void setClippedSceneInRobot(SceneType* scene_) { _clipped_scene_in_robot = scene_; }  // :50-52
  void setFullScene(SceneType* scene_) { _full_scene = scene_; }                        // :58-60
  void setRobotInLocalMap(const EstimateType& robot_in_local_map_) {                 // :80-83
    _robot_in_local_map = robot_in_local_map_;
    _local_map_in_robot = _robot_in_local_map.inverse();
  }
  Status status() const { return _status; }                                          // :98-100
  virtual const std::vector<int> globalIndices() const { return std::vector<int>(0); }  // :104-107

protected:
  SceneType* _clipped_scene_in_robot = nullptr;
  SceneType* _full_scene = nullptr;
  EstimateType _robot_in_local_map = EstimateType::Identity();
  EstimateType _local_map_in_robot = EstimateType::Identity();
  Status _status = Error;
};

}  // namespace srrg2_slam_interfaces
