// STAND-IN for R/registration/loop_detector/{loop_detector.h:26-85, multi_loop_detector_brute_force.h:8-47},
// R/registration/local_map_selectors/local_map_selector.h:12-75, R/registration/loop_closure.h:33-71 and
// R/mapping/local_map.h:15-84: the members the GPU loop-detector adapter touches, with the reference's names.
// Shapes only (each proven by the cited line); not upstream code.
#pragma once
#include <memory>
#include <set>
#include <vector>

#include "srrg2_core/stub.h"
#include "srrg2_slam_interfaces/registration/aligners/multi_aligner.h"

namespace srrg2_slam_interfaces {

// LocalMap_: an SE(d) variable that owns a dynamic property container (local_map.h:15-31,40)
template <typename EstimateType_>
struct LocalMapStub {
  using EstimateType = EstimateType_;
  int _graph_id = -1;
  srrg2_core::PropertyContainerDynamic dynamic_properties;
  int graphId() const { return _graph_id; }
};

// LoopClosure_: the pose-pose factor a detected closure becomes (loop_closure.h:45-71)
template <typename LocalMapType_, int InfoDim_>
struct LoopClosureStub {
  using LocalMapType = LocalMapType_;
  using EstimateType = typename LocalMapType_::EstimateType;
  using MeasurementType = EstimateType;
  struct InformationMatrixType {
    std::array<float, InfoDim_ * InfoDim_> a{};
    static InformationMatrixType Identity() { InformationMatrixType I; for (int i = 0; i < InfoDim_; ++i) I.a[(size_t) (i * InfoDim_ + i)] = 1.f; return I; }
    static InformationMatrixType Zero() { return InformationMatrixType(); }
  };
  LoopClosureStub(int graph_id_, LocalMapType* source_, LocalMapType* target_, const MeasurementType& measurement_,
                  const InformationMatrixType& info_, const EstimateType& pose_in_target_, const float& chi_inliers_,
                  const size_t& num_inliers_, const size_t& num_correspondences_)
      : graph_id(graph_id_), source(source_), target(target_), measurement(measurement_), information(info_),
        pose_in_target(pose_in_target_), chi_inliers(chi_inliers_), num_inliers(num_inliers_),
        num_correspondences(num_correspondences_) {}
  int graph_id;
  LocalMapType *source, *target;
  MeasurementType measurement;
  InformationMatrixType information;
  EstimateType pose_in_target;
  float chi_inliers;
  size_t num_inliers, num_correspondences;
};

// what the detectors ask of MultiGraphSLAM_ (multi_loop_detector_brute_force_impl.cpp:19-23,56-57)
template <typename LoopClosureType_>
struct SLAMAlgorithmStub {
  using LoopClosureType = LoopClosureType_;
  using LocalMapType = typename LoopClosureType_::LocalMapType;
  using EstimateType = typename LocalMapType::EstimateType;
  LocalMapType* _current = nullptr;
  EstimateType _robot_in_local_map = EstimateType::Identity();
  std::vector<LocalMapType*> _local_maps;  // graph()->variables() upstream
  LocalMapType* currentLocalMap() const { return _current; }
  const EstimateType& robotInLocalMap() const { return _robot_in_local_map; }
  const std::vector<LocalMapType*>& localMaps() const { return _local_maps; }
};

// LocalMapSelector_: produces the closure hints (local_map_selector.h:24-66)
template <typename SLAMAlgorithmType_>
class LocalMapSelector_ : public srrg2_core::Configurable {
public:
  using LoopClosureType = typename SLAMAlgorithmType_::LoopClosureType;
  using LocalMapType = typename LoopClosureType::LocalMapType;
  using EstimateType = typename LocalMapType::EstimateType;
  struct ClosureHint {
    LocalMapType* local_map;
    EstimateType initial_guess;
    explicit ClosureHint(LocalMapType* lmap, const EstimateType& guess = EstimateType::Identity()) : local_map(lmap), initial_guess(guess) {}
  };
  using ClosureHintPtr = std::shared_ptr<ClosureHint>;
  struct ClosureHintPtrComparator {
    bool operator()(const ClosureHintPtr& a, const ClosureHintPtr& b) const { return a->local_map->graphId() < b->local_map->graphId(); }
  };
  using ClosureHintPtrSet = std::set<ClosureHintPtr, ClosureHintPtrComparator>;
  void setSLAMAlgorithm(SLAMAlgorithmType_* slam_) { _slam = slam_; }
  virtual void compute() = 0;
  ClosureHintPtrSet& hints() { return _hints; }

protected:
  ClosureHintPtrSet _hints;
  SLAMAlgorithmType_* _slam = nullptr;
};

// LoopDetector_ (loop_detector.h:26-85)
template <typename SLAMAlgorithmType_>
class LoopDetector_ : public srrg2_core::Configurable {
public:
  using SLAMAlgorithmType = SLAMAlgorithmType_;
  using LoopClosureType = typename SLAMAlgorithmType_::LoopClosureType;
  using LocalMapType = typename LoopClosureType::LocalMapType;
  using InformationMatrixType = typename LoopClosureType::InformationMatrixType;
  using LocalMapSelectorType = LocalMapSelector_<SLAMAlgorithmType_>;
  using LoopClosurePtrContainer = std::vector<std::shared_ptr<LoopClosureType>>;
  using LocalMapRawPtrContainer = std::set<LocalMapType*>;
  srrg2_core::PropertyConfigurable_<LocalMapSelectorType> param_local_map_selector;  // :39-43
  virtual void compute() = 0;                                                        // :48
  const LocalMapRawPtrContainer& attemptedClosures() const { return _attempted_closures; }
  const LoopClosurePtrContainer& detectedClosures() const { return _detected_closures; }
  void setSLAMAlgorithm(SLAMAlgorithmType_* slam_) { _slam = slam_; }

protected:
  SLAMAlgorithmType_* _slam = nullptr;
  LocalMapRawPtrContainer _attempted_closures;
  LoopClosurePtrContainer _detected_closures;
};

// MultiLoopDetectorBruteForce_ (multi_loop_detector_brute_force.h:8-47): parameters; compute() is what the adapter replaces
template <typename SLAMAlgorithmType_, typename AlignerType_>
class MultiLoopDetectorBruteForce_ : public LoopDetector_<SLAMAlgorithmType_> {
public:
  using AlignerType = AlignerType_;
  srrg2_core::PropertyConfigurable_<AlignerType_> param_relocalize_aligner;                                       // :20-24
  PARAM(srrg2_core::PropertyInt, relocalize_min_inliers, "minimum number of inliers for success [int]", 500, nullptr);       // :25-29
  PARAM(srrg2_core::PropertyFloat, relocalize_max_chi_inliers, "maximum chi per inlier for success [chi]", 0.005f, nullptr); // :30-34
  PARAM(srrg2_core::PropertyFloat, relocalize_min_inliers_ratio, "minimum fraction of inliers over total correspondences", 0.7f, nullptr);  // :35-40
  void compute() override {}
};

}  // namespace srrg2_slam_interfaces
