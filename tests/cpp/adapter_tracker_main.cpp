// One tracker frame through the reference-side adapters over the stub surface (adapters/scene_b200.h,
// adapters/multi_aligner_b200.h): clip the resident local map (SceneClipperRangeB200_), align the new scan against the
// clip (MultiAligner2DB200, moving cloud resident), merge the scan into the map (MergerCorrespondenceHomoB200_), bring
// the map back.  Flow: R/trackers/multi_tracker_impl.cpp:82-138, R/trackers/tracker_slice_processor_impl.cpp:159-205.
// Without a CUDA device: exit code 3 (loud failure).
#include <cstdio>

#include "scene_b200.h"

using namespace srrg2_slam_interfaces;
using Cloud = srrg2_core::PointNormal2fVectorCloud;
using Iso = srrg2_core::Isometry2f;
using Clipper = SceneClipperRangeB200_<Iso, Cloud, MultiAligner2DB200>;
using Merger = MergerCorrespondenceHomoB200_<Iso, Cloud, MultiAligner2DB200>;

static Cloud corner(int n, float length, float dx, float dy) {
  Cloud c((size_t) n);
  for (int i = 0; i < n; ++i) {
    const float t = length * (float) (i / 2) / (float) (n / 2);
    const bool horizontal = i % 2 == 0;
    c[(size_t) i]._c = {{(horizontal ? t : 0.f) + dx, (horizontal ? 0.f : t) + dy}};
    c[(size_t) i]._n = {{horizontal ? 0.f : 1.f, horizontal ? 1.f : 0.f}};
  }
  return c;
}

int main() {
  Cloud map = corner(6000, 12.f, 0.f, 0.f);       // the local map: two 12 m walls
  Cloud scan = corner(2000, 8.f, 0.02f, -0.015f);  // the new scan: the first 8 m of them, 2.5 cm off
  srrg2_core::PropertyContainerDynamic measurement;
  measurement._properties["points"] = &scan;
  auto slice = std::make_shared<AlignerSliceProcessor_<Iso, Cloud>>();
  slice->param_fixed_slice_name.setValue("points");
  auto aligner = std::make_shared<MultiAligner2DB200>();
  aligner->param_slice_processors.pushBack(slice);
  aligner->param_max_distance_m.setValue(0.5f);
  Clipper clipper;
  clipper.setAligner(aligner);
  clipper.param_range.setValue(9.f);
  Merger merger;
  merger.setAligner(aligner);
  merger.param_maximum_response.setValue(0.2f);
  merger.param_maximum_distance_geometry_squared.setValue(0.02f);
  merger.param_target_number_of_merges.setValue(1 << 30);
  try {
    clipper.setFullScene(&map);
    clipper.setRobotInLocalMap(Iso::Identity());
    clipper.compute();
    const size_t n_clipped = clipper.globalIndices().size();
    aligner->setFixed(&measurement);               // the measurement is the aligner's fixed side (multi_tracker_impl.cpp:97-98)
    aligner->setMovingInFixed(Iso::Identity());
    aligner->compute();
    const size_t n_corr = slice->correspondences().size();
    merger.setScene(&map);
    merger.setMeasurement(&scan);
    merger.setCorrespondences(&slice->correspondences());
    merger.setMeasurementInScene(aligner->movingInFixed().inverse());  // scan in map = (map in scan)^-1 at an identity clip pose
    merger.compute();
    const size_t before = map.size();
    merger.downloadScene();
    printf("clipped %zu status %d\n", n_clipped, (int) clipper.status());
    printf("aligner status %d correspondences %zu inliers %d\n", (int) aligner->status(), n_corr, aligner->iterationStats().back().num_inliers);
    printf("merged %lld added %lld scene %zu -> %zu\n", (long long) merger.numMerged(), (long long) merger.numAdded(), before, map.size());
    const auto& T = aligner->movingInFixed().matrix();
    printf("map in scan: tx %.4f ty %.4f\n", (double) T(0, 2), (double) T(1, 2));
  } catch (const std::runtime_error& e) {
    fprintf(stderr, "%s\n", e.what());
    return std::string(e.what()).find("no usable CUDA device") != std::string::npos ? 3 : 5;
  }
  return 0;
}
