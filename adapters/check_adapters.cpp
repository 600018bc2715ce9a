// Instantiates every adapter against the stub headers: `g++ -std=c++17 -fsyntax-only -I include -I adapters/stubs -I adapters`.
#include "correspondence_finder_b200.h"
#include "multi_aligner_b200.h"
#include "solver_b200.h"

template class srrg2_slam_interfaces::CorrespondenceFinderB200_<srrg2_core::Isometry2f, srrg2_core::PointNormal2fVectorCloud>;
template class srrg2_slam_interfaces::CorrespondenceFinderB200_<srrg2_core::Isometry3f, srrg2_core::PointNormal3fVectorCloud>;
template class srrg2_slam_interfaces::MultiAlignerB200_<srrg2_solver::VariableSE2RightAD, srrg2_core::PointNormal2fVectorCloud>;
template class srrg2_slam_interfaces::MultiAlignerB200_<srrg2_solver::VariableSE3QuaternionRightAD, srrg2_core::PointNormal3fVectorCloud>;
template class srrg2_slam_interfaces::PoseGraphSolverB200_<2>;
template class srrg2_slam_interfaces::PoseGraphSolverB200_<3>;
