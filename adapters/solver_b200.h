// solver_b200.h -- reference-side adapter for MultiGraphSLAM_::param_global_solver (R/system/multi_graph_slam.h:50-54):
// a Solver subclass whose compute() optimises the pose graph (row a10) on the GPU.
// Replaces: Solver::compute() as called at R/system/multi_graph_slam_impl.cpp:314-316 on the graph that
// :51-90 builds (local-map pose variables, SE{2,3}PosePoseGeodesicErrorFactor closures and odometry links).
#pragma once
#include <srrg2_solver/stub.h>  // the real build includes <srrg2_solver/solvers_core/solver.h> and the factor / variable headers

#include "srrg2b_flatten.h"

namespace srrg2_slam_interfaces {

template <int Dim>
class PoseGraphSolverB200_ : public srrg2_solver::Solver {
public:
  using Graph = srrg2_solver::FactorGraphStub<Dim>;
  PARAM(srrg2_core::PropertyInt, device, "CUDA device", 0, nullptr);
  PARAM(srrg2_core::PropertyFloat, dx_epsilon, "stop when the largest perturbation component falls below", 1e-6f, nullptr);

  ~PoseGraphSolverB200_() override { if (_ctx) srrg2b_ctx_destroy(_ctx); }
  void setGraph(std::shared_ptr<Graph> g) { _graph = std::move(g); }

  void compute() override {
    if (!_graph) throw std::runtime_error("PoseGraphSolverB200_::compute|no graph");
    if (!_ctx) srrg2b_adapters::check(nullptr, srrg2b_ctx_create(Dim, param_device.value(), &_ctx) == SRRG2B_OK ? SRRG2B_OK : SRRG2B_ERR_CUDA,
                                      "PoseGraphSolverB200_|no usable CUDA device (there is no CPU fallback)");
    constexpr int M = (Dim + 1) * (Dim + 1), B = Dim == 3 ? 36 : 9;
    auto& vars = _graph->variables();
    auto& facs = _graph->factors();
    std::vector<float> poses(vars.size() * M), Z(facs.size() * M), Om(facs.size() * B);
    std::vector<uint8_t> fixed(vars.size());
    std::vector<int32_t> ij(facs.size() * 2);
    for (size_t v = 0; v < vars.size(); ++v) {  // graph ids are dense 0..V-1 in MultiGraphSLAM_ (multi_graph_slam_impl.cpp:56)
      srrg2b_adapters::to_row_major(vars[v].estimate(), &poses[v * M]);
      fixed[v] = vars[v].fixed() ? 1 : 0;       // the first local map is the gauge (:85-87)
    }
    for (size_t f = 0; f < facs.size(); ++f) {
      ij[2 * f] = facs[f].variableId(0);
      ij[2 * f + 1] = facs[f].variableId(1);
      srrg2b_adapters::to_row_major(facs[f].measurement(), &Z[f * M]);
      if (facs[f].informationMatrix().size() != (size_t) B) throw std::runtime_error("PoseGraphSolverB200_::compute|bad information matrix");
      std::memcpy(&Om[f * B], facs[f].informationMatrix().data(), sizeof(float) * B);
    }
    srrg2b_adapters::check(_ctx, srrg2b_pgo_upload(_ctx, (int64_t) vars.size(), poses.data(), fixed.data(), (int64_t) facs.size(), ij.data(),
                                                   Z.data(), Om.data()), "PoseGraphSolverB200_::compute");
    const int max_it = this->param_max_iterations.value();
    std::vector<srrg2b_pgo_stats> st((size_t) (max_it > 0 ? max_it : 1));
    int32_t n = 0;
    srrg2b_adapters::check(_ctx, srrg2b_pgo_optimize(_ctx, max_it, param_dx_epsilon.value(), 0, st.data(), &n), "PoseGraphSolverB200_::compute");
    srrg2b_adapters::check(_ctx, srrg2b_pgo_download(_ctx, poses.data()), "PoseGraphSolverB200_::compute");
    for (size_t v = 0; v < vars.size(); ++v) vars[v].setEstimate(srrg2b_adapters::from_row_major<srrg2_core::IsometryStub<Dim>>(&poses[v * M]));
    this->_iteration_stats.clear();
    for (int32_t k = 0; k < n; ++k) {
      srrg2_solver::IterationStats is;
      is.iteration = k;
      is.num_inliers = st[(size_t) k].num_factors;
      is.chi_inliers = (float) st[(size_t) k].chi;
      this->_iteration_stats.push_back(is);
    }
    this->_status = srrg2_solver::SolverBase::Success;
  }

private:
  srrg2b_ctx* _ctx = nullptr;
  std::shared_ptr<Graph> _graph;
};

using PoseGraphSolver2DB200 = PoseGraphSolverB200_<2>;
using PoseGraphSolver3DB200 = PoseGraphSolverB200_<3>;

}  // namespace srrg2_slam_interfaces
