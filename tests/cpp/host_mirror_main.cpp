// Test driver of the C++ host mirror (include/srrg2b.hpp): reads two point+normal clouds, runs the
// finder and the aligner through the reference-shaped classes, writes the results for the pytest side
// (tests/test_cpp_host_mirror.py) to compare with the oracle.  Exit code 3 = no CUDA device.
//   host_mirror_main <input.bin> <output.bin>
#include <cstdio>
#include <cstdlib>

#include "srrg2b.hpp"

using namespace srrg2b;

template <typename T>
static void rd(FILE* f, T* p, size_t n) {
  if (fread(p, sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
}
template <typename T>
static void wr(FILE* f, const T* p, size_t n) {
  if (fwrite(p, sizeof(T), n, f) != n) { fprintf(stderr, "short write\n"); exit(2); }
}

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s input.bin output.bin\n", argv[0]); return 2; }
  ContextPtr ctx;
  try {
    ctx = std::make_shared<Context>(3, 0);
  } catch (const std::runtime_error& e) {
    fprintf(stderr, "%s\n", e.what());
    return 3;
  }
  FILE* in = fopen(argv[1], "rb");
  if (!in) { fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
  int32_t hdr[2];
  rd(in, hdr, 2);
  PointNormalCloud<3> fixed, moving;
  fixed.coordinates.resize((size_t) hdr[0] * 3); fixed.normals.resize((size_t) hdr[0] * 3);
  moving.coordinates.resize((size_t) hdr[1] * 3); moving.normals.resize((size_t) hdr[1] * 3);
  rd(in, fixed.coordinates.data(), fixed.coordinates.size());
  rd(in, fixed.normals.data(), fixed.normals.size());
  rd(in, moving.coordinates.data(), moving.coordinates.size());
  rd(in, moving.normals.data(), moving.normals.size());
  fclose(in);
  FILE* out = fopen(argv[2], "wb");
  if (!out) { fprintf(stderr, "cannot open %s\n", argv[2]); return 2; }
  try {
    // --- the way a tracker slice drives its finder (aligner_slice_processor_impl.cpp:38-48) ---
    CorrespondenceVector correspondences;
    CorrespondenceFinderB200<3> finder(ctx, 7);
    finder.param_max_distance_m = 0.3f;
    finder.param_normal_cos = 0.8f;
    finder.setCorrespondences(&correspondences);
    finder.setFixed(&fixed);
    finder.setMoving(&moving);
    finder.setLocalMapInSensor(Isometry<3>::Identity());
    finder.compute();
    const int64_t nf = (int64_t) correspondences.size();
    wr(out, &nf, 1);
    for (const Correspondence& c : correspondences) { wr(out, &c.fixed_idx, 1); wr(out, &c.moving_idx, 1); wr(out, &c.response, 1); }

    // --- MultiAligner3DQR::compute() with one point slice (config C2 shape) ---
    MultiAlignerB200<3> aligner(ctx);
    aligner.param_max_iterations = 8;
    aligner.param_min_num_inliers = 10;
    MultiAlignerB200<3>::SliceProcessor sp;
    sp.finder_max_distance_m = 0.3f;
    sp.finder_normal_cos = 0.8f;
    sp.factor = SRRG2B_FACTOR_PLANE;
    sp.robustifier = SRRG2B_ROB_HUBER;
    sp.robustifier_chi_threshold = 0.01f;
    const int k = aligner.addSliceProcessor(sp);
    aligner.setFixed(k, &fixed);
    aligner.setMoving(k, &moving);
    aligner.setMovingInFixed(Isometry<3>::Identity());
    aligner.compute();
    wr(out, aligner.movingInFixed().data(), 16);
    const int32_t status = (int32_t) aligner.status(), ns = (int32_t) aligner.iterationStats().size();
    wr(out, &status, 1);
    wr(out, &ns, 1);
    wr(out, aligner.iterationStats().data(), (size_t) ns);
    const CorrespondenceVector& fc = aligner.sliceProcessor(k).correspondences;
    const int64_t nc = (int64_t) fc.size();
    wr(out, &nc, 1);
    for (const Correspondence& c : fc) { wr(out, &c.fixed_idx, 1); wr(out, &c.moving_idx, 1); wr(out, &c.response, 1); }
    if (aligner.numCorrespondences() != (int) nc) { fprintf(stderr, "numCorrespondences mismatch\n"); return 4; }

    // --- MultiLoopDetectorBruteForce_::compute (multi_loop_detector_brute_force_impl.cpp:63-133): the source map
    // against three hints -- the whole target map, a hint without a local map (skipped), half of the map ---
    {
      PointNormalCloud<3> half;
      half.coordinates.assign(moving.coordinates.begin(), moving.coordinates.begin() + (moving.size() / 2) * 3);
      half.normals.assign(moving.normals.begin(), moving.normals.begin() + (moving.size() / 2) * 3);
      LoopDetectorBruteForceB200<3> detector(0);
      detector.param_relocalize_min_inliers = 100;
      detector.param_relocalize_min_inliers_ratio = 0.5f;
      detector.aligner_slice = sp;
      detector.aligner_max_iterations = 8;
      detector.aligner_enable_inlier_only_runs = true;
      detector.setSource(&fixed);
      std::vector<LoopDetectorBruteForceB200<3>::Hint> hints(3);
      hints[0].local_map = &moving;
      hints[2].local_map = &half;
      detector.compute(hints);
      detector.compute(hints);  // a second call reuses the contexts and the lent index
      const int32_t nr = (int32_t) detector.results().size(), nd = (int32_t) detector.detectedClosures().size();
      wr(out, &nr, 1);
      for (const srrg2b_closure_result& r : detector.results()) {
        wr(out, &r.verdict, 1); wr(out, &r.aligner_status, 1); wr(out, &r.num_correspondences, 1); wr(out, &r.num_inliers, 1);
        wr(out, &r.chi_inliers, 1); wr(out, r.moving_in_fixed, 16);
      }
      wr(out, &nd, 1);
      for (const auto& cl : detector.detectedClosures()) { const int32_t t = cl.target; wr(out, &t, 1); wr(out, cl.moving_in_fixed.data(), 16); }
    }

    // --- misconfiguration throws, like the reference (aligner_slice_processor_impl.cpp:13-16) ---
    bool threw = false;
    try {
      MultiAlignerB200<3> empty(ctx);
      empty.compute();
    } catch (const std::runtime_error&) { threw = true; }
    if (!threw) { fprintf(stderr, "misconfiguration did not throw\n"); return 4; }
  } catch (const std::runtime_error& e) {
    fprintf(stderr, "unexpected error: %s\n", e.what());
    return 5;
  }
  fclose(out);
  return 0;
}
