// multi_loop_detector_b200.h -- reference-side adapter: MultiLoopDetectorBruteForce_::compute()
// (R/registration/loop_detector/multi_loop_detector_brute_force_impl.cpp:13-133) with the K candidate alignments in
// flight together.  Prologue (checks, hints, _attempted_closures) and epilogue (LoopClosure construction) follow the
// reference line by line; the serial `for (auto& h : hints) { setMoving; setMovingInFixed; compute; gates }` body becomes
// one srrg2b_closure_batch call over per-candidate contexts that borrow the source map's cloud and NN index.
// param_relocalize_aligner must be a MultiAlignerB200_ (its slice / finder / robustifier configuration is what runs).
#pragma once
#include <srrg2_slam_interfaces/registration/loop_detector/multi_loop_detector_brute_force.h>

#include "multi_aligner_b200.h"

namespace srrg2_slam_interfaces {

template <typename SLAMAlgorithmType_, typename AlignerType_>
class MultiLoopDetectorBruteForceB200_ : public MultiLoopDetectorBruteForce_<SLAMAlgorithmType_, AlignerType_> {
public:
  using ThisType = MultiLoopDetectorBruteForceB200_<SLAMAlgorithmType_, AlignerType_>;
  using BaseType = MultiLoopDetectorBruteForce_<SLAMAlgorithmType_, AlignerType_>;
  using LoopClosureType = typename SLAMAlgorithmType_::LoopClosureType;
  using LocalMapType = typename SLAMAlgorithmType_::LocalMapType;
  using EstimateType = typename LocalMapType::EstimateType;
  using LocalMapSelectorType = typename BaseType::LocalMapSelectorType;
  using InformationMatrixType = typename LoopClosureType::InformationMatrixType;
  static constexpr int Dim = EstimateType::Dim;
  PARAM(srrg2_core::PropertyInt, device, "CUDA device", 0, nullptr);

  ~MultiLoopDetectorBruteForceB200_() override {
    for (srrg2b_ctx* c : _candidate_ctx) srrg2b_ctx_destroy(c);  // borrowers first, then the lender
    if (_source_ctx) srrg2b_ctx_destroy(_source_ctx);
  }

  // per candidate, in hint order: what the reference prints in its DEBUG lines (verdict, counts, chi)
  const std::vector<srrg2b_closure_result>& results() const { return _results; }

  void compute() override {
    if (!this->_slam) throw std::runtime_error("MultiLoopDetectorBruteForceB200_::compute| no slam selected");  // :15-17
    LocalMapType* source_local_map = this->_slam->currentLocalMap();
    if (!source_local_map) throw std::runtime_error("MultiLoopDetectorBruteForceB200_::compute| _current_local_map is NULL");  // :24-27
    using ClosureHint = typename LocalMapSelectorType::ClosureHint;
    using ClosureHintPtr = typename LocalMapSelectorType::ClosureHintPtr;
    typename LocalMapSelectorType::ClosureHintPtrSet hints;
    this->_attempted_closures.clear();                                  // :35
    if (this->param_local_map_selector.value()) {                       // :36-39
      this->param_local_map_selector->setSLAMAlgorithm(this->_slam);
      this->param_local_map_selector->compute();
      hints = this->param_local_map_selector->hints();
    } else {                                                            // :40-47: every other local map of the graph
      for (LocalMapType* m : this->_slam->localMaps())
        if (m != source_local_map) hints.insert(ClosureHintPtr(new ClosureHint(m)));
    }
    this->_detected_closures.clear();                                   // :49
    _results.clear();
    std::shared_ptr<AlignerType_> aligner = this->param_relocalize_aligner.value();
    if (!aligner) throw std::runtime_error("MultiLoopDetectorBruteForceB200_::compute| no aligner");  // :50-53
    const EstimateType& pose_in_current = this->_slam->robotInLocalMap();  // :56-57

    std::vector<srrg2b_slice> slices;
    aligner->describeSlices(slices);
    const srrg2b_aligner_params ap = aligner->alignerParams();
    const int n_slices = (int) slices.size();

    // aligner->setFixed(&source_local_map->dynamic_properties) (:63): uploaded and indexed ONCE, on the lender context
    if (!_source_ctx) make_context(_source_ctx);
    aligner->setFixed(&source_local_map->dynamic_properties);
    std::vector<LocalMapType*> targets;
    std::vector<float> guesses;
    for (const ClosureHintPtr& h : hints) {                             // :64-72
      LocalMapType* target_local_map = const_cast<LocalMapType*>(h->local_map);
      if (!target_local_map) continue;
      this->_attempted_closures.insert(target_local_map);
      targets.push_back(target_local_map);
      float T[16] = {0};
      srrg2b_adapters::to_row_major(h->initial_guess, T);               // aligner->setMovingInFixed(h->initial_guess), :77
      guesses.insert(guesses.end(), T, T + (Dim + 1) * (Dim + 1));
    }
    if (targets.empty()) return;
    for (int s = 0; s < n_slices; ++s) {
      if (slices[(size_t) s].kind != SRRG2B_SLICE_POINTS) continue;
      auto* sp = aligner->pointSlice((size_t) s);
      if (!sp->fixed()) throw std::runtime_error("MultiLoopDetectorBruteForceB200_::compute|source map lacks a slice's cloud");
      upload(_source_ctx, SRRG2B_FIXED, s, *sp->fixed());
      // the index is built for the finder radius before it is lent: a find over one point of the first target does that
      aligner->setMoving(&targets[0]->dynamic_properties);
      if (sp->moving() && !sp->moving()->empty()) {
        typename std::remove_reference<decltype(*sp->moving())>::type one(sp->moving()->begin(), sp->moving()->begin() + 1);
        upload(_source_ctx, SRRG2B_MOVING, s, one);
        float I[16] = {0};
        srrg2b_adapters::to_row_major(EstimateType::Identity(), I);
        int64_t n = 0;
        srrg2b_adapters::check(_source_ctx, srrg2b_find_correspondences(_source_ctx, s, I, &slices[(size_t) s].finder, nullptr, nullptr, nullptr, &n),
                               "MultiLoopDetectorBruteForceB200_::index");
      }
    }
    while (_candidate_ctx.size() < targets.size()) {
      srrg2b_ctx* c = nullptr;
      make_context(c);
      _candidate_ctx.push_back(c);
    }
    for (size_t k = 0; k < targets.size(); ++k) {
      aligner->setMoving(&targets[k]->dynamic_properties);              // :76
      for (int s = 0; s < n_slices; ++s) {
        if (slices[(size_t) s].kind != SRRG2B_SLICE_POINTS) continue;
        auto* sp = aligner->pointSlice((size_t) s);
        if (!sp->moving()) throw std::runtime_error("MultiLoopDetectorBruteForceB200_::compute|target map lacks a slice's cloud");
        srrg2b_adapters::check(_candidate_ctx[k], srrg2b_share_fixed(_candidate_ctx[k], s, _source_ctx, s), "MultiLoopDetectorBruteForceB200_::share");
        upload(_candidate_ctx[k], SRRG2B_MOVING, s, *sp->moving());
      }
    }
    srrg2b_closure_params cp;
    cp.relocalize_min_inliers = this->param_relocalize_min_inliers.value();
    cp.relocalize_max_chi_inliers = this->param_relocalize_max_chi_inliers.value();
    cp.relocalize_min_inliers_ratio = this->param_relocalize_min_inliers_ratio.value();
    _results.resize(targets.size());
    srrg2b_adapters::check(_candidate_ctx[0],
                           srrg2b_closure_batch(_candidate_ctx.data(), (int) targets.size(), n_slices, slices.data(), &ap, guesses.data(), &cp,
                                                _results.data()),
                           "MultiLoopDetectorBruteForceB200_::compute");
    for (size_t k = 0; k < targets.size(); ++k) {                       // :80-127, candidate order
      const srrg2b_closure_result& r = _results[k];
      if (r.verdict != SRRG2B_CLOSURE_ACCEPT) continue;                 // ALIGNER / NUM_INLIERS / MAX_CHI_INLIERS / MIN_INLIERS_RATIO DROP
      const EstimateType moving_in_fixed = srrg2b_adapters::from_row_major<EstimateType>(r.moving_in_fixed);
      const EstimateType pose_in_target = moving_in_fixed.inverse() * pose_in_current;  // :115
      this->_detected_closures.push_back(std::make_shared<LoopClosureType>(          // :117-127
        -1, source_local_map, targets[k], moving_in_fixed, InformationMatrixType::Identity(), pose_in_target, r.chi_inliers,
        (size_t) r.num_inliers, (size_t) r.num_correspondences));
    }
  }

private:
  void make_context(srrg2b_ctx*& c) {
    srrg2b_adapters::check(nullptr, srrg2b_ctx_create(Dim, param_device.value(), &c) == SRRG2B_OK ? SRRG2B_OK : SRRG2B_ERR_CUDA,
                           "MultiLoopDetectorBruteForceB200_|no usable CUDA device (there is no CPU fallback)");
  }
  template <typename Cloud>
  static void upload(srrg2b_ctx* ctx, int slot, int slice, const Cloud& cloud) {
    const srrg2b_adapters::FlatCloud f = srrg2b_adapters::flatten(cloud);
    const srrg2b_cloud c = f.describe();
    srrg2b_adapters::check(ctx, srrg2b_set_cloud(ctx, slot, slice, &c), "MultiLoopDetectorBruteForceB200_::upload");
  }
  srrg2b_ctx* _source_ctx = nullptr;
  std::vector<srrg2b_ctx*> _candidate_ctx;
  std::vector<srrg2b_closure_result> _results;
};

// SLAM algorithm types as R/system/multi_graph_slam.h names them; registerTypes(): BOSS_REGISTER_CLASS(MultiLoopDetectorBruteForce2DB200) ...
// (R/instances.cpp:72-73)

}  // namespace srrg2_slam_interfaces
