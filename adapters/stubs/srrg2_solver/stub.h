// MINIMAL STAND-INS for the srrg2_solver symbols the adapters touch: shapes only, proven by
// the reference's call sites (file:line in the comments).  Not upstream code.
#pragma once
#include <memory>
#include <vector>

#include "srrg2_core/stub.h"

namespace srrg2_solver {

// IterationStats fields the reference reads (R/registration/aligners/aligner_termination_criteria_impl.cpp:26-32)
struct IterationStats {
  int iteration = 0, num_inliers = 0, num_outliers = 0, num_suppressed = 0;
  float chi_inliers = 0.f, chi_outliers = 0.f;
};
using IterationStatsVector = std::vector<IterationStats>;

// RobustifierBase::param_chi_threshold, RobustifierClamp (R/registration/aligners/multi_aligner_impl.cpp:189-197)
struct RobustifierBase {
  srrg2_core::PropertyFloat param_chi_threshold{1.f};
  virtual ~RobustifierBase() = default;
};
struct RobustifierClamp : RobustifierBase {};
struct RobustifierCauchy : RobustifierBase {};
struct RobustifierSaturated : RobustifierBase {};

// variables: ::EstimateType, estimate / setEstimate (R/registration/aligners/multi_aligner.h:152-158)
struct VariableSE2RightAD { using EstimateType = srrg2_core::Isometry2f; };
struct VariableSE3EulerRightAD { using EstimateType = srrg2_core::Isometry3f; };
struct VariableSE3QuaternionRightAD { using EstimateType = srrg2_core::Isometry3f; };

// pose-graph side: FactorGraph with SE{2,3} pose variables and pose-pose factors (R/system/multi_graph_slam_impl.cpp:51-90)
template <int D>
struct PoseVariableStub {
  int _id = -1;
  bool _fixed = false;
  srrg2_core::IsometryStub<D> _x = srrg2_core::IsometryStub<D>::Identity();
  int graphId() const { return _id; }
  const srrg2_core::IsometryStub<D>& estimate() const { return _x; }
  void setEstimate(const srrg2_core::IsometryStub<D>& x) { _x = x; }
  bool fixed() const { return _fixed; }  // status() == VariableBase::Fixed upstream
};
template <int D>
struct PosePoseFactorStub {  // SE{2,3}PosePoseGeodesicErrorFactor: variableId(i), measurement(), informationMatrix()
  int _ids[2] = {-1, -1};
  srrg2_core::IsometryStub<D> _z = srrg2_core::IsometryStub<D>::Identity();
  std::vector<float> _omega;  // (D == 3 ? 36 : 9) entries, row-major
  int variableId(int k) const { return _ids[k]; }
  const srrg2_core::IsometryStub<D>& measurement() const { return _z; }
  const std::vector<float>& informationMatrix() const { return _omega; }
};
template <int D>
struct FactorGraphStub {
  std::vector<PoseVariableStub<D>> _variables;
  std::vector<PosePoseFactorStub<D>> _factors;
  std::vector<PoseVariableStub<D>>& variables() { return _variables; }
  std::vector<PosePoseFactorStub<D>>& factors() { return _factors; }
};

// Solver surface the reference uses: setGraph, compute, status() == SolverBase::Success, iterationStats(),
// clearIterationStats(), param_max_iterations (R/registration/aligners/multi_aligner_impl.cpp:59,66,112-118,
// R/system/multi_graph_slam_impl.cpp:314-316)
struct SolverBase {
  enum SolverStatus { Error = 0, Ready = 1, Processing = 2, Success = 3 };
  virtual ~SolverBase() = default;
};
struct Solver : SolverBase, srrg2_core::Configurable {
  srrg2_core::PropertyInt param_max_iterations{10};
  virtual void compute() = 0;
  SolverStatus status() const { return _status; }
  const IterationStatsVector& iterationStats() const { return _iteration_stats; }
  void clearIterationStats() { _iteration_stats.clear(); }

protected:
  SolverStatus _status = Error;
  IterationStatsVector _iteration_stats;
};

}  // namespace srrg2_solver
