// STAND-IN for R/registration/aligners/{aligner.h:23-127, aligner_slice_processor_base.h:34-189,
// aligner_slice_processor.h:56-66,142-156, aligner_slice_processor_prior.h:41-98, multi_aligner.h:34-66,95}:
// the members the GPU aligner adapter touches, with the reference's names.
#pragma once
#include "srrg2_core/stub.h"
#include "srrg2_solver/stub.h"
#include "srrg2_slam_interfaces/registration/correspondence_finder.h"

namespace srrg2_slam_interfaces {

class AlignerBase : public srrg2_core::Configurable {  // aligner.h:13-44
public:
  enum Status { Success = 0, NotEnoughCorrespondences = 1, NotEnoughInliers = 2, Fail = 3 };  // :23-28
  PARAM(srrg2_core::PropertyInt, max_iterations, "maximum number of iterations", 10, nullptr);  // :30
  virtual void compute() = 0;
  Status status() const { return _status; }
  const srrg2_solver::IterationStatsVector& iterationStats() const { return _iteration_stats; }

protected:
  Status _status = Fail;
  srrg2_solver::IterationStatsVector _iteration_stats;
};

class AlignerTerminationCriteriaStandard : public srrg2_core::Configurable {  // aligner_termination_criteria.h:40-56
public:
  PARAM(srrg2_core::PropertyInt, window_size, "", 5, nullptr);
  PARAM(srrg2_core::PropertyInt, num_correspondences_range, "", 20, nullptr);
  PARAM(srrg2_core::PropertyInt, num_inliers_range, "", 20, nullptr);
  PARAM(srrg2_core::PropertyInt, num_outliers_range, "", 20, nullptr);
  PARAM(srrg2_core::PropertyFloat, chi_epsilon, "", 0.2f, nullptr);
};

// slice processors: the pure virtuals / accessors the adapter reads (aligner_slice_processor_base.h:115-189)
template <typename EstimateType_>
class AlignerSliceProcessorBase_ : public srrg2_core::Configurable {
public:
  using EstimateType = EstimateType_;
  srrg2_core::PropertyConfigurable_<srrg2_solver::RobustifierBase> param_robustifier;  // :34-38
  PARAM(srrg2_core::PropertyString, fixed_slice_name, "name of the slice in the fixed scene", "", nullptr);    // :40-45
  PARAM(srrg2_core::PropertyString, moving_slice_name, "name of the slice in the moving scene", "", nullptr);  // :46-51
  virtual void setFixed(srrg2_core::PropertyContainerBase*) {}   // :97-101 (binds the slice's cloud by name)
  virtual void setMoving(srrg2_core::PropertyContainerBase*) {}  // :106-110
  virtual bool isPrior() const = 0;
  virtual ~AlignerSliceProcessorBase_() = default;
};
// point slices (aligner_slice_processor.h): finder slot, min_num_correspondences, clouds, sensor pose, correspondences
template <typename EstimateType_, typename CloudType_>
class AlignerSliceProcessor_ : public AlignerSliceProcessorBase_<EstimateType_> {
public:
  using CloudType = CloudType_;
  srrg2_core::PropertyConfigurable_<CorrespondenceFinder_<EstimateType_, CloudType_, CloudType_>> param_finder;  // :56-60
  PARAM(srrg2_core::PropertyInt, min_num_correspondences, "", 0, nullptr);                                       // :62-66
  bool isPrior() const override { return false; }
  void setFixed(srrg2_core::PropertyContainerBase* scene) override {
    _fixed = scene ? scene->template property<CloudType>(this->param_fixed_slice_name.value()) : nullptr;
    _fixed_changed = true;
  }
  void setMoving(srrg2_core::PropertyContainerBase* scene) override {
    _moving = scene ? scene->template property<CloudType>(this->param_moving_slice_name.value()) : nullptr;
    _moving_changed = true;
  }
  CloudType* fixed() const { return _fixed; }      // bound by name from the tracker slice (…_base_impl.cpp:7-51)
  CloudType* moving() const { return _moving; }
  bool fixedChanged() const { return _fixed_changed; }
  bool movingChanged() const { return _moving_changed; }
  void clearChanged() { _fixed_changed = _moving_changed = false; }
  const EstimateType_& robotInSensor() const { return _robot_in_sensor; }  // setSensorInRobot, :142-150
  srrg2_core::CorrespondenceVector& correspondences() { return _correspondences; }  // :156
  int factorKind() const { return _factor_kind; }  // which FactorCorrespondenceDriven_ the slice instantiates (R/instances.h:27-30)

  CloudType* _fixed = nullptr;
  CloudType* _moving = nullptr;
  bool _fixed_changed = true, _moving_changed = true;
  EstimateType_ _robot_in_sensor = EstimateType_::Identity();
  srrg2_core::CorrespondenceVector _correspondences;
  int _factor_kind = 1;
};
// prior slices (aligner_slice_processor_prior.h:41-98, aligner_slice_odometry_prior.cpp:6-37): one SE(d) prior factor
template <typename EstimateType_>
class AlignerSliceProcessorPrior_ : public AlignerSliceProcessorBase_<EstimateType_> {
public:
  bool isPrior() const override { return true; }
  virtual void setupFactor() = 0;  // refreshes measurement() (the `_count > 1` rule lives in the reference class)
  const EstimateType_& measurement() const { return _measurement; }
  const std::array<float, 6>& diagonalInfo() const { return _diag; }

protected:
  EstimateType_ _measurement = EstimateType_::Identity();
  std::array<float, 6> _diag{{1.f, 1.f, 1.f, 1.f, 1.f, 1.f}};
};

template <typename VariableType_>
class MultiAlignerBase_ : public AlignerBase {  // multi_aligner.h:24-150
public:
  using VariableType = VariableType_;
  using EstimateType = typename VariableType_::EstimateType;
  srrg2_core::PropertyConfigurableVector_<AlignerSliceProcessorBase_<EstimateType>> param_slice_processors;  // :34-37
  srrg2_core::PropertyConfigurable_<srrg2_solver::Solver> param_solver;                                     // :39-43
  srrg2_core::PropertyConfigurable_<AlignerTerminationCriteriaStandard> param_termination_criteria;         // aligner.h:31-35
  PARAM(srrg2_core::PropertyInt, min_num_inliers, "", 10, nullptr);                                         // :45-46
  PARAM(srrg2_core::PropertyBool, enable_inlier_only_runs, "", false, nullptr);                             // :47-51
  PARAM(srrg2_core::PropertyBool, keep_only_inlier_correspondences, "", false, nullptr);                    // :53-57
  void setFixed(srrg2_core::PropertyContainerBase* scene) {   // multi_aligner_impl.cpp:8-14
    for (size_t s = 0; s < param_slice_processors.size(); ++s) param_slice_processors.value(s)->setFixed(scene);
  }
  void setMoving(srrg2_core::PropertyContainerBase* scene) {  // multi_aligner_impl.cpp:18-24
    for (size_t s = 0; s < param_slice_processors.size(); ++s) param_slice_processors.value(s)->setMoving(scene);
  }
  void setMovingInFixed(const EstimateType& T) { _moving_in_fixed = T; }  // aligner.h:62-70
  const EstimateType& movingInFixed() const { return _moving_in_fixed; }
  void compute() override {}                                              // :95 (virtual upstream)

protected:
  EstimateType _moving_in_fixed = EstimateType::Identity();
};

}  // namespace srrg2_slam_interfaces
