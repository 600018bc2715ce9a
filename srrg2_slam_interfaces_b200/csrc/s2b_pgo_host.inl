// s2b_pgo_host.inl -- host side of the pose-graph path (included at the end of s2b_api.cu):
// block-CSR structure from the factor list, launch sequencing of linearise + PCG + update, and the
// extern "C" entry points srrg2b_pgo_*.
#include "s2b_pgo.cuh"

#include <algorithm>
#include <numeric>

struct PgoState {
  int V = 0, F = 0, nnzb = 0;
  DevBuf<double> poses, Z, Omega, vals, b, x, r, z, p, Ap, Minv;
  DevBuf<unsigned char> fixed;
  DevBuf<int> ij, slots, row_ptr, col_idx, diag_slot;
  DevBuf<float> stage;
  PgoScalars* d_sc = nullptr;
  PgoScalars* h_sc = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
};

static std::map<srrg2b_ctx*, PgoState*> g_pgo;

static PgoState* pgo_of(srrg2b_ctx* c, bool create) {
  auto it = g_pgo.find(c);
  if (it != g_pgo.end()) return it->second;
  if (!create) return nullptr;
  PgoState* s = new PgoState();
  g_pgo[c] = s;
  return s;
}

static void pgo_release(srrg2b_ctx* c) {
  auto it = g_pgo.find(c);
  if (it == g_pgo.end()) return;
  PgoState* s = it->second;
  s->poses.release(); s->Z.release(); s->Omega.release(); s->vals.release(); s->b.release(); s->x.release();
  s->r.release(); s->z.release(); s->p.release(); s->Ap.release(); s->Minv.release(); s->fixed.release();
  s->ij.release(); s->slots.release(); s->row_ptr.release(); s->col_idx.release(); s->diag_slot.release();
  s->stage.release();
  if (s->d_sc) cudaFree(s->d_sc);
  if (s->h_sc) cudaFreeHost(s->h_sc);
  if (s->e0) cudaEventDestroy(s->e0);
  if (s->e1) cudaEventDestroy(s->e1);
  if (s->e2) cudaEventDestroy(s->e2);
  delete s;
  g_pgo.erase(it);
}

extern "C" {

// FactorGraph upload: variables (LocalMap poses, R/mapping/local_map.h:64,75), gauge mask
// (VariableBase::Fixed, R/system/multi_graph_slam_impl.cpp:85-87) and pose-pose factors
// (R/system/multi_graph_slam_impl.cpp:73-79, R/registration/loop_closure.h:68-78)
int srrg2b_pgo_upload(srrg2b_ctx* c, int64_t n_vars, const float* poses16, const uint8_t* fixed_mask, int64_t n_factors,
                      const int32_t* ij, const float* Z16, const float* Omega36) {
  if (!c) return SRRG2B_ERR_INVALID;
  if (c->dim != 3) FAIL(c, SRRG2B_ERR_INVALID, "the pose-graph path is built for dim == 3");
  if (n_vars <= 0 || n_factors < 0 || !poses16 || !fixed_mask || (n_factors > 0 && (!ij || !Z16 || !Omega36)) ||
      n_vars > 0x7fffffff / 36 || n_factors > 0x7fffffff / 36)
    FAIL(c, SRRG2B_ERR_INVALID, "bad pose-graph description");
  CK(c, cudaSetDevice(c->device));
  const int V = (int) n_vars, F = (int) n_factors;
  for (int f = 0; f < F; ++f)
    if (ij[2 * f] < 0 || ij[2 * f] >= V || ij[2 * f + 1] < 0 || ij[2 * f + 1] >= V || ij[2 * f] == ij[2 * f + 1])
      FAIL(c, SRRG2B_ERR_INVALID, "factor references an unknown variable");
  PgoState* s = pgo_of(c, true);
  s->V = V; s->F = F;
  // ---- block-CSR structure (both triangles): unique (row, col) pairs, diagonal always present ----
  std::vector<std::pair<int, int>> pairs;
  pairs.reserve((size_t) V + 2 * (size_t) F);
  for (int v = 0; v < V; ++v) pairs.emplace_back(v, v);
  for (int f = 0; f < F; ++f) {
    pairs.emplace_back(ij[2 * f], ij[2 * f + 1]);
    pairs.emplace_back(ij[2 * f + 1], ij[2 * f]);
  }
  std::sort(pairs.begin(), pairs.end());
  pairs.erase(std::unique(pairs.begin(), pairs.end()), pairs.end());
  const int nnzb = (int) pairs.size();
  s->nnzb = nnzb;
  std::vector<int> row_ptr(V + 1, 0), col_idx(nnzb), diag(V, 0), slots(4 * (size_t) F);
  for (int k = 0; k < nnzb; ++k) {
    row_ptr[pairs[k].first + 1]++;
    col_idx[k] = pairs[k].second;
    if (pairs[k].first == pairs[k].second) diag[pairs[k].first] = k;
  }
  std::partial_sum(row_ptr.begin(), row_ptr.end(), row_ptr.begin());
  auto slot_of = [&](int r, int cc) {
    const int* b = col_idx.data() + row_ptr[r];
    const int* e = col_idx.data() + row_ptr[r + 1];
    return (int) (std::lower_bound(b, e, cc) - col_idx.data());
  };
  for (int f = 0; f < F; ++f) {
    const int i = ij[2 * f], j = ij[2 * f + 1];
    slots[4 * (size_t) f] = diag[i];
    slots[4 * (size_t) f + 1] = slot_of(i, j);
    slots[4 * (size_t) f + 2] = slot_of(j, i);
    slots[4 * (size_t) f + 3] = diag[j];
  }
  const size_t n6 = (size_t) V * 6;
  CK(c, s->poses.ensure((size_t) V * 12)); CK(c, s->fixed.ensure(V)); CK(c, s->ij.ensure(2 * (size_t) F + 2));
  CK(c, s->Z.ensure((size_t) F * 12 + 12)); CK(c, s->Omega.ensure((size_t) F * 36 + 36));
  CK(c, s->slots.ensure(4 * (size_t) F + 4)); CK(c, s->row_ptr.ensure(V + 1)); CK(c, s->col_idx.ensure(nnzb));
  CK(c, s->diag_slot.ensure(V)); CK(c, s->vals.ensure((size_t) nnzb * 36)); CK(c, s->Minv.ensure((size_t) V * 36));
  CK(c, s->b.ensure(n6)); CK(c, s->x.ensure(n6)); CK(c, s->r.ensure(n6)); CK(c, s->z.ensure(n6));
  CK(c, s->p.ensure(n6)); CK(c, s->Ap.ensure(n6));
  CK(c, s->stage.ensure(std::max((size_t) V * 16, (size_t) F * 36) + 16));
  if (!s->d_sc) {
    CK(c, cudaMalloc((void**) &s->d_sc, sizeof(PgoScalars)));
    CK(c, cudaMallocHost((void**) &s->h_sc, sizeof(PgoScalars)));
    CK(c, cudaEventCreate(&s->e0)); CK(c, cudaEventCreate(&s->e1)); CK(c, cudaEventCreate(&s->e2));
  }
  cudaStream_t st = c->stream;
  CK(c, cudaMemcpyAsync(s->stage.p, poses16, sizeof(float) * 16 * (size_t) V, cudaMemcpyHostToDevice, st));
  pgo_pack_poses_kernel<<<blocks_for(V, 256), 256, 0, st>>>(s->stage.p, V, s->poses.p);
  c->launches++;
  if (F > 0) {
    CK(c, cudaMemcpyAsync(s->stage.p, Z16, sizeof(float) * 16 * (size_t) F, cudaMemcpyHostToDevice, st));
    pgo_pack_poses_kernel<<<blocks_for(F, 256), 256, 0, st>>>(s->stage.p, F, s->Z.p);
    CK(c, cudaMemcpyAsync(s->stage.p, Omega36, sizeof(float) * 36 * (size_t) F, cudaMemcpyHostToDevice, st));
    pgo_cast_kernel<<<blocks_for((int64_t) F * 36, 256), 256, 0, st>>>(s->stage.p, (size_t) F * 36, s->Omega.p);
    c->launches += 2;
    CK(c, cudaMemcpyAsync(s->ij.p, ij, sizeof(int) * 2 * (size_t) F, cudaMemcpyHostToDevice, st));
    CK(c, cudaMemcpyAsync(s->slots.p, slots.data(), sizeof(int) * 4 * (size_t) F, cudaMemcpyHostToDevice, st));
  }
  CK(c, cudaMemcpyAsync(s->fixed.p, fixed_mask, V, cudaMemcpyHostToDevice, st));
  CK(c, cudaMemcpyAsync(s->row_ptr.p, row_ptr.data(), sizeof(int) * (V + 1), cudaMemcpyHostToDevice, st));
  CK(c, cudaMemcpyAsync(s->col_idx.p, col_idx.data(), sizeof(int) * (size_t) nnzb, cudaMemcpyHostToDevice, st));
  CK(c, cudaMemcpyAsync(s->diag_slot.p, diag.data(), sizeof(int) * V, cudaMemcpyHostToDevice, st));
  CK(c, cudaStreamSynchronize(st));
  CK(c, cudaGetLastError());
  return SRRG2B_OK;
}

// One Gauss-Newton iteration of Solver::compute() on the uploaded graph: linearise (factor-sharded
// over the ranks of the communicator, all-reduce of H/b only), PCG solve, X <- X [+] dx.
int srrg2b_pgo_iterate(srrg2b_ctx* c, int max_cg_iterations, double cg_tolerance, srrg2b_pgo_stats* out) {
  if (!c) return SRRG2B_ERR_INVALID;
  PgoState* s = pgo_of(c, false);
  if (!s || s->V == 0) FAIL(c, SRRG2B_ERR_STATE, "no pose graph uploaded");
  if (max_cg_iterations <= 0) max_cg_iterations = 2000;
  if (!(cg_tolerance > 0.0)) cg_tolerance = 1e-10;
  CK(c, cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const int V = s->V, F = s->F, n = V * 6;
  CK(c, cudaEventRecord(s->e0, st));
  CK(c, cudaMemsetAsync(s->vals.p, 0, sizeof(double) * 36 * (size_t) s->nnzb, st));
  CK(c, cudaMemsetAsync(s->b.p, 0, sizeof(double) * (size_t) n, st));
  CK(c, cudaMemsetAsync(s->d_sc, 0, sizeof(PgoScalars), st));
  const int mine = (F - c->rank + c->world - 1) / c->world;  // factors f = rank, rank + world, ...
  if (mine > 0) {
    pgo_linearize_kernel<<<blocks_for(mine, 128), 128, 0, st>>>(s->poses.p, s->fixed.p, s->ij.p, s->Z.p, s->Omega.p,
                                                                s->slots.p, F, c->rank, c->world, s->vals.p, s->b.p,
                                                                s->d_sc);
    c->launches++;
  }
  if (c->world > 1) {  // the only exchange of the graph path: H values, b and chi
    const int kF64 = 8;
    if (g_nccl.AllReduce(s->vals.p, s->vals.p, (size_t) s->nnzb * 36, kF64, kNcclSum, c->comm, st) != 0 ||
        g_nccl.AllReduce(s->b.p, s->b.p, (size_t) n, kF64, kNcclSum, c->comm, st) != 0 ||
        g_nccl.AllReduce(&s->d_sc->chi, &s->d_sc->chi, 1, kF64, kNcclSum, c->comm, st) != 0)
      FAIL(c, SRRG2B_ERR_NCCL, "ncclAllReduce of the pose-graph system failed");
  }
  pgo_fix_diag_kernel<<<blocks_for(V, 256), 256, 0, st>>>(s->fixed.p, s->diag_slot.p, V, s->vals.p);
  CK(c, cudaEventRecord(s->e1, st));
  pgo_block_inverse_kernel<<<blocks_for(V, 128), 128, 0, st>>>(s->vals.p, s->diag_slot.p, V, s->Minv.p);
  pgo_cg_init_kernel<<<blocks_for(n, 192), 192, 0, st>>>(s->b.p, s->Minv.p, n, s->x.p, s->r.p, s->z.p, s->p.p, s->d_sc);
  c->launches += 3;
  int it = 0;
  double rel = 1.0;
  const int check_every = 20;
  while (it < max_cg_iterations) {
    for (int k = 0; k < check_every && it < max_cg_iterations; ++k, ++it) {
      const int par = it & 1;
      pgo_cg_spmv_kernel<<<blocks_for(n, 192), 192, 0, st>>>(s->row_ptr.p, s->col_idx.p, s->vals.p, s->p.p, n, par,
                                                             s->Ap.p, s->d_sc);
      pgo_cg_update_kernel<<<blocks_for(n, 192), 192, 0, st>>>(s->Minv.p, s->p.p, s->Ap.p, n, par, s->x.p, s->r.p,
                                                               s->z.p, s->d_sc);
      pgo_cg_direction_kernel<<<blocks_for(n, 192), 192, 0, st>>>(s->z.p, n, par, s->p.p, s->d_sc);
      c->launches += 3;
    }
    CK(c, cudaMemcpyAsync(s->h_sc, s->d_sc, sizeof(PgoScalars), cudaMemcpyDeviceToHost, st));
    CK(c, cudaStreamSynchronize(st));
    const double rr = s->h_sc->rr[it & 1];
    rel = s->h_sc->b_norm2 > 0.0 ? sqrt(rr / s->h_sc->b_norm2) : 0.0;
    if (!(rel > cg_tolerance)) break;
  }
  pgo_update_kernel<<<blocks_for(V, 128), 128, 0, st>>>(s->x.p, s->fixed.p, V, s->poses.p, s->d_sc);
  c->launches++;
  CK(c, cudaEventRecord(s->e2, st));
  CK(c, cudaMemcpyAsync(s->h_sc, s->d_sc, sizeof(PgoScalars), cudaMemcpyDeviceToHost, st));
  CK(c, cudaStreamSynchronize(st));
  CK(c, cudaGetLastError());
  if (out) {
    float ms1 = 0.f, ms2 = 0.f;
    CK(c, cudaEventElapsedTime(&ms1, s->e0, s->e1));
    CK(c, cudaEventElapsedTime(&ms2, s->e1, s->e2));
    out->chi = s->h_sc->chi;
    out->dx_norm_inf = s->h_sc->dx_max;
    out->cg_iterations = it;
    out->cg_relative_residual = rel;
    out->linearize_ms = ms1;
    out->solve_ms = ms2;
    out->num_blocks = s->nnzb;
    out->num_factors = F;
  }
  return SRRG2B_OK;
}

int srrg2b_pgo_download(srrg2b_ctx* c, float* poses16) {
  if (!c) return SRRG2B_ERR_INVALID;
  PgoState* s = pgo_of(c, false);
  if (!s || s->V == 0 || !poses16) FAIL(c, SRRG2B_ERR_STATE, "no pose graph uploaded");
  CK(c, cudaSetDevice(c->device));
  pgo_unpack_poses_kernel<<<blocks_for(s->V, 256), 256, 0, c->stream>>>(s->poses.p, s->V, s->stage.p);
  c->launches++;
  CK(c, cudaMemcpyAsync(poses16, s->stage.p, sizeof(float) * 16 * (size_t) s->V, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  return SRRG2B_OK;
}

}  // extern "C"
