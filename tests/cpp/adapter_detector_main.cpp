// Drives the reference-side loop-detector adapter (adapters/multi_loop_detector_b200.h) over the stub SLAM surface:
// two local maps with one point+normal slice each, a MultiAligner2DB200 as param_relocalize_aligner, compute().
// Without a CUDA device the adapter must fail loudly (exit code 3); with one it prints the verdicts.
#include <cstdio>

#include "multi_loop_detector_b200.h"

using namespace srrg2_slam_interfaces;
using LocalMap = LocalMapStub<srrg2_core::Isometry2f>;
using SLAM = SLAMAlgorithmStub<LoopClosureStub<LocalMap, 3>>;
using Cloud = srrg2_core::PointNormal2fVectorCloud;
using Detector = MultiLoopDetectorBruteForceB200_<SLAM, MultiAligner2DB200>;

static Cloud wall(int n, float dx, float dy) {
  Cloud c((size_t) n);
  for (int i = 0; i < n; ++i) {
    const float t = 10.f * (float) i / (float) n;
    const bool horizontal = i % 2 == 0;  // an L-shaped corner: two walls
    c[(size_t) i]._c = {{(horizontal ? t : 0.f) + dx, (horizontal ? 0.f : t) + dy}};
    c[(size_t) i]._n = {{horizontal ? 0.f : 1.f, horizontal ? 1.f : 0.f}};
  }
  return c;
}

int main() {
  Cloud source = wall(4000, 0.f, 0.f), near_map = wall(3000, 0.03f, -0.02f), far_map = wall(3000, 3.f, 3.f);
  LocalMap a, b, c;
  a._graph_id = 0; b._graph_id = 1; c._graph_id = 2;
  a.dynamic_properties._properties["points"] = &source;
  b.dynamic_properties._properties["points"] = &near_map;
  c.dynamic_properties._properties["points"] = &far_map;
  SLAM slam;
  slam._current = &a;
  slam._local_maps = {&a, &b, &c};
  auto slice = std::make_shared<AlignerSliceProcessor_<srrg2_core::Isometry2f, Cloud>>();
  slice->param_fixed_slice_name.setValue("points");
  slice->param_moving_slice_name.setValue("points");
  auto aligner = std::make_shared<MultiAligner2DB200>();
  aligner->param_slice_processors.pushBack(slice);
  aligner->param_max_distance_m.setValue(0.5f);
  Detector detector;
  detector.param_relocalize_aligner.setValue(aligner);
  detector.param_relocalize_min_inliers.setValue(500);
  detector.setSLAMAlgorithm(&slam);
  try {
    detector.compute();
  } catch (const std::runtime_error& e) {
    fprintf(stderr, "%s\n", e.what());
    return std::string(e.what()).find("no usable CUDA device") != std::string::npos ? 3 : 5;
  }
  printf("attempted %zu detected %zu\n", detector.attemptedClosures().size(), detector.detectedClosures().size());
  for (const auto& r : detector.results()) printf("verdict %d inliers %lld correspondences %lld chi %g\n", r.verdict, (long long) r.num_inliers, (long long) r.num_correspondences, (double) r.chi_inliers);
  for (const auto& cl : detector.detectedClosures()) printf("closure -> map %d\n", cl->target->graphId());
  return 0;
}
