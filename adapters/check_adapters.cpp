// Instantiates every adapter against the stub headers: `g++ -std=c++17 -fsyntax-only -I include -I adapters/stubs -I adapters`.
#include "correspondence_finder_b200.h"
#include "multi_aligner_b200.h"
#include "solver_b200.h"
#include "multi_loop_detector_b200.h"
#include "scene_b200.h"

template class srrg2_slam_interfaces::CorrespondenceFinderB200_<srrg2_core::Isometry2f, srrg2_core::PointNormal2fVectorCloud>;
template class srrg2_slam_interfaces::CorrespondenceFinderB200_<srrg2_core::Isometry3f, srrg2_core::PointNormal3fVectorCloud>;
template class srrg2_slam_interfaces::MultiAlignerB200_<srrg2_solver::VariableSE2RightAD, srrg2_core::PointNormal2fVectorCloud>;
template class srrg2_slam_interfaces::MultiAlignerB200_<srrg2_solver::VariableSE3QuaternionRightAD, srrg2_core::PointNormal3fVectorCloud>;
template class srrg2_slam_interfaces::PoseGraphSolverB200_<2>;
template class srrg2_slam_interfaces::PoseGraphSolverB200_<3>;

// the loop detector over the stub SLAM surface (LocalMap / LoopClosure / MultiGraphSLAM shapes of the stub headers)
namespace {
using LocalMap2DStub = srrg2_slam_interfaces::LocalMapStub<srrg2_core::Isometry2f>;
using LocalMap3DStub = srrg2_slam_interfaces::LocalMapStub<srrg2_core::Isometry3f>;
using SLAM2DStub = srrg2_slam_interfaces::SLAMAlgorithmStub<srrg2_slam_interfaces::LoopClosureStub<LocalMap2DStub, 3>>;
using SLAM3DStub = srrg2_slam_interfaces::SLAMAlgorithmStub<srrg2_slam_interfaces::LoopClosureStub<LocalMap3DStub, 6>>;
}  // namespace
template class srrg2_slam_interfaces::MultiLoopDetectorBruteForceB200_<SLAM2DStub, srrg2_slam_interfaces::MultiAligner2DB200>;
template class srrg2_slam_interfaces::MultiLoopDetectorBruteForceB200_<SLAM3DStub, srrg2_slam_interfaces::MultiAligner3DQRB200>;
template class srrg2_slam_interfaces::SceneClipperRangeB200_<srrg2_core::Isometry2f, srrg2_core::PointNormal2fVectorCloud, srrg2_slam_interfaces::MultiAligner2DB200>;
template class srrg2_slam_interfaces::SceneClipperRangeB200_<srrg2_core::Isometry3f, srrg2_core::PointNormal3fVectorCloud, srrg2_slam_interfaces::MultiAligner3DQRB200>;
template class srrg2_slam_interfaces::MergerCorrespondenceHomoB200_<srrg2_core::Isometry2f, srrg2_core::PointNormal2fVectorCloud, srrg2_slam_interfaces::MultiAligner2DB200>;
template class srrg2_slam_interfaces::MergerCorrespondenceHomoB200_<srrg2_core::Isometry3f, srrg2_core::PointNormal3fVectorCloud, srrg2_slam_interfaces::MultiAligner3DQRB200>;
