// s2b_pgo_host.inl -- host side of the pose-graph path (included at the end of s2b_api.cu):
// block-CSR structure and gather lists from the factor list, launch sequencing of linearise + PCG + update,
// the Levenberg-Marquardt driver, and the extern "C" entry points srrg2b_pgo_*.
#include "s2b_pgo.cuh"

#include <algorithm>
#include <numeric>

struct PgoState {
  int D = 6;  // degrees of freedom per pose: 6 (SE(3), ctx dim 3) or 3 (SE(2), ctx dim 2)
  int V = 0, F = 0, nnzb = 0, n_local = 0;
  DevBuf<double> poses, poses_try, Z, Omega, vals, b, x, r, z, p, Ap, Minv, diag0, rec, rec_try, chi_f, parts;
  DevBuf<unsigned char> fixed;
  DevBuf<int> ij, row_ptr, col_idx, row_of, diag_slot, blk_ptr, blk_src, var_ptr, var_src;
  DevBuf<float> stage;
  PgoScalars* d_sc = nullptr;
  PgoScalars* h_sc = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
  bool need_assemble = true, diag_saved = false;
  const double* assembled_for_rec = nullptr;
  bool rec_valid = false;     // `rec` holds the factor records of the current poses
  double chi_cur = 0.0;       // ... and this is their chi
  double lambda = 0.0, nu = 2.0, dx_prev = 1e300, cg_tol = 1e-4;
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
  s->poses.release(); s->poses_try.release(); s->Z.release(); s->Omega.release(); s->vals.release(); s->b.release();
  s->x.release(); s->r.release(); s->z.release(); s->p.release(); s->Ap.release(); s->Minv.release(); s->diag0.release();
  s->rec.release(); s->rec_try.release(); s->chi_f.release(); s->parts.release(); s->fixed.release();
  s->ij.release(); s->row_ptr.release(); s->col_idx.release(); s->row_of.release(); s->diag_slot.release();
  s->blk_ptr.release(); s->blk_src.release(); s->var_ptr.release(); s->var_src.release(); s->stage.release();
  if (s->d_sc) cudaFree(s->d_sc);
  if (s->h_sc) cudaFreeHost(s->h_sc);
  if (s->e0) cudaEventDestroy(s->e0);
  if (s->e1) cudaEventDestroy(s->e1);
  if (s->e2) cudaEventDestroy(s->e2);
  delete s;
  g_pgo.erase(it);
}

namespace {

// the factor records of `poses` (this rank's share of the factors) and their chi, summed in a fixed order and --
// several ranks -- all-reduced
template <int D>
int pgo_factors(srrg2b_ctx* c, PgoState* s, const double* poses, bool chi_only, double* rec, double* chi_out) {
  cudaStream_t st = c->stream;
  if (s->n_local > 0) {
    pgo_factor_kernel<D><<<blocks_for(s->n_local, 128), 128, 0, st>>>(poses, s->ij.p, s->Z.p, s->Omega.p, s->F, c->rank, c->world,
                                                                     chi_only ? 1 : 0, rec, s->chi_f.p);
    c->launches++;
  }
  pgo_sum_kernel<<<1, 1024, 0, st>>>(s->chi_f.p, s->n_local, &s->d_sc->chi);
  c->launches++;
  if (c->world > 1 && g_nccl.AllReduce(&s->d_sc->chi, &s->d_sc->chi, 1, 8 /* f64 */, kNcclSum, c->comm, st) != 0)
    FAIL(c, SRRG2B_ERR_NCCL, "ncclAllReduce of chi failed");
  CK(c, cudaMemcpyAsync(&s->h_sc->chi, &s->d_sc->chi, sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(c, cudaStreamSynchronize(st));
  *chi_out = s->h_sc->chi;
  return SRRG2B_OK;
}

// H (block-CSR values) and b from the records: gathers in factor order, then -- several ranks -- the only exchange of
// the graph path: all-reduce of H values and b
template <int D>
int pgo_assemble(srrg2b_ctx* c, PgoState* s) {
  cudaStream_t st = c->stream;
  constexpr int BB = PgoDim<D>::BB;
  pgo_gather_blocks_kernel<D><<<blocks_for((int64_t) s->nnzb * BB, 256), 256, 0, st>>>(s->blk_ptr.p, s->blk_src.p, s->row_of.p, s->col_idx.p,
                                                                                      s->fixed.p, s->rec.p, s->nnzb, s->vals.p);
  pgo_gather_b_kernel<D><<<blocks_for((int64_t) s->V * D, 256), 256, 0, st>>>(s->var_ptr.p, s->var_src.p, s->fixed.p, s->rec.p, s->V, s->b.p);
  c->launches += 2;
  if (c->world > 1) {
    if (g_nccl.AllReduce(s->vals.p, s->vals.p, (size_t) s->nnzb * BB, 8, kNcclSum, c->comm, st) != 0 ||
        g_nccl.AllReduce(s->b.p, s->b.p, (size_t) s->V * D, 8, kNcclSum, c->comm, st) != 0)
      FAIL(c, SRRG2B_ERR_NCCL, "ncclAllReduce of the pose-graph system failed");
  }
  return SRRG2B_OK;
}

// (H + lambda diag(H)) dx = -b by block-Jacobi PCG to the relative residual `tol`; dx in s->x
template <int D>
int pgo_solve(srrg2b_ctx* c, PgoState* s, double lambda, bool save_diag, int max_cg, double tol, int* iters, double* rel_out) {
  cudaStream_t st = c->stream;
  const int V = s->V, n = V * D;
  const int threads = 192;  // a multiple of 3 and 6
  pgo_damp_kernel<D><<<blocks_for(n, 256), 256, 0, st>>>(s->diag_slot.p, V, lambda, save_diag ? 1 : 0, s->diag0.p, s->vals.p);
  pgo_fix_diag_kernel<D><<<blocks_for(V, 256), 256, 0, st>>>(s->fixed.p, s->diag_slot.p, V, s->vals.p);
  pgo_block_inverse_kernel<D><<<blocks_for(V, 128), 128, 0, st>>>(s->vals.p, s->diag_slot.p, V, s->Minv.p);
  pgo_cg_init_kernel<D><<<blocks_for(n, threads), threads, 0, st>>>(s->b.p, s->Minv.p, n, s->x.p, s->r.p, s->z.p, s->p.p, s->d_sc, s->parts.p);
  c->launches += 4;
  int it = 0;
  double rel = 1.0;
  const int check_every = 10;
  while (it < max_cg) {
    for (int k = 0; k < check_every && it < max_cg; ++k, ++it) {
      const int par = it & 1;
      pgo_cg_spmv_kernel<D><<<blocks_for(n, threads), threads, 0, st>>>(s->row_ptr.p, s->col_idx.p, s->vals.p, s->p.p, n, par, s->Ap.p,
                                                                       s->d_sc, s->parts.p);
      pgo_cg_update_kernel<D><<<blocks_for(n, threads), threads, 0, st>>>(s->Minv.p, s->p.p, s->Ap.p, n, par, s->x.p, s->r.p, s->z.p,
                                                                         s->d_sc, s->parts.p);
      pgo_cg_direction_kernel<<<blocks_for(n, threads), threads, 0, st>>>(s->z.p, n, par, s->p.p, s->d_sc);
      c->launches += 3;
    }
    CK(c, cudaMemcpyAsync(s->h_sc, s->d_sc, sizeof(PgoScalars), cudaMemcpyDeviceToHost, st));
    CK(c, cudaStreamSynchronize(st));
    const double rr = s->h_sc->rr[it & 1];
    rel = s->h_sc->b_norm2 > 0.0 ? sqrt(rr / s->h_sc->b_norm2) : 0.0;
    if (!(rel > tol)) break;
  }
  *iters = it;
  *rel_out = rel;
  return SRRG2B_OK;
}

template <int D>
int pgo_iterate_t(srrg2b_ctx* c, PgoState* s, int max_cg, double tol, srrg2b_pgo_stats* out) {
  cudaStream_t st = c->stream;
  CK(c, cudaEventRecord(s->e0, st));
  CK(c, cudaMemsetAsync(s->d_sc, 0, sizeof(PgoScalars), st));
  double chi = 0.0;
  int rcode = pgo_factors<D>(c, s, s->poses.p, false, s->rec.p, &chi);
  if (rcode) return rcode;
  rcode = pgo_assemble<D>(c, s);
  if (rcode) return rcode;
  CK(c, cudaEventRecord(s->e1, st));
  int it = 0;
  double rel = 1.0;
  rcode = pgo_solve<D>(c, s, 0.0, true, max_cg, tol, &it, &rel);
  if (rcode) return rcode;
  pgo_update_kernel<D><<<blocks_for(s->V, 128), 128, 0, st>>>(s->x.p, s->fixed.p, s->V, s->poses.p, s->poses.p, s->d_sc);
  c->launches++;
  s->rec_valid = false;
  CK(c, cudaEventRecord(s->e2, st));
  CK(c, cudaMemcpyAsync(s->h_sc, s->d_sc, sizeof(PgoScalars), cudaMemcpyDeviceToHost, st));
  CK(c, cudaStreamSynchronize(st));
  CK(c, cudaGetLastError());
  if (out) {
    float ms1 = 0.f, ms2 = 0.f;
    CK(c, cudaEventElapsedTime(&ms1, s->e0, s->e1));
    CK(c, cudaEventElapsedTime(&ms2, s->e1, s->e2));
    memset(out, 0, sizeof(*out));
    out->chi = chi;
    out->chi_after = -1.0;
    out->dx_norm_inf = s->h_sc->dx_max;
    out->cg_iterations = it;
    out->cg_relative_residual = rel;
    out->linearize_ms = ms1;
    out->solve_ms = ms2;
    out->num_blocks = s->nnzb;
    out->num_factors = s->F;
    out->accepted = 1;
  }
  return SRRG2B_OK;
}

// One Levenberg-Marquardt iteration (Nielsen's damping update): damped step, candidate chi, gain ratio, accept / reject.
// The linear solve is inexact: its tolerance follows the outer convergence (forcing term), so the early iterations
// -- far from the optimum -- do not pay for digits the next linearisation throws away.
template <int D>
int pgo_lm_iterate_t(srrg2b_ctx* c, PgoState* s, int max_cg, double tol_floor, srrg2b_pgo_stats* out) {
  cudaStream_t st = c->stream;
  constexpr int POSE = PgoDim<D>::POSE;
  const int n = s->V * D;
  CK(c, cudaEventRecord(s->e0, st));
  CK(c, cudaMemsetAsync(s->d_sc, 0, sizeof(PgoScalars), st));
  int rcode;
  const bool fresh = !s->rec_valid;
  if (fresh) {
    rcode = pgo_factors<D>(c, s, s->poses.p, false, s->rec.p, &s->chi_cur);
    if (rcode) return rcode;
    s->rec_valid = true;
  }
  // (a rejected step leaves H and b as they are: only the damping changes; after an accepted one they are rebuilt)
  if (s->assembled_for_rec != s->rec.p || fresh || s->need_assemble) {
    rcode = pgo_assemble<D>(c, s);
    if (rcode) return rcode;
    s->need_assemble = false;
    s->diag_saved = false;
  }
  s->assembled_for_rec = s->rec.p;
  CK(c, cudaEventRecord(s->e1, st));
  // forcing term: loose (1e-4; measured on C4: 1e-2 wanders into a long valley, 1e-6 costs 1.6x the time for the same
  // 14 iterations) while chi still falls by factors; every accepted step that gains less than half
  // tightens it tenfold (s->cg_tol), small steps tighten it further
  double tol = s->cg_tol;
  if (s->dx_prev < 1e-2) tol = std::min(tol, 1e-4);
  if (s->dx_prev < 1e-4) tol = std::min(tol, 1e-7);
  tol = std::max(tol, tol_floor);
  int it = 0;
  double rel = 1.0;
  rcode = pgo_solve<D>(c, s, s->lambda, !s->diag_saved, max_cg, tol, &it, &rel);
  if (rcode) return rcode;
  s->diag_saved = true;
  pgo_model_kernel<D><<<blocks_for(n, 192), 192, 0, st>>>(s->row_ptr.p, s->col_idx.p, s->vals.p, s->x.p, s->b.p, s->diag0.p, n, s->d_sc, s->parts.p);
  pgo_update_kernel<D><<<blocks_for(s->V, 128), 128, 0, st>>>(s->x.p, s->fixed.p, s->V, s->poses.p, s->poses_try.p, s->d_sc);
  c->launches += 2;
  CK(c, cudaMemcpyAsync(s->h_sc, s->d_sc, sizeof(PgoScalars), cudaMemcpyDeviceToHost, st));
  CK(c, cudaStreamSynchronize(st));
  const double dx_max = s->h_sc->dx_max;
  const double bdx = s->h_sc->dot_b_dx, dHd = s->h_sc->dot_dx_Hdx, dDd = s->h_sc->dot_dx_Ddx;
  // model: chi(dx) = chi + 2 b.dx + dx.H dx with H = H_damped - lambda D
  const double predicted = -2.0 * bdx - (dHd - s->lambda * dDd);
  double chi_new = 0.0;
  rcode = pgo_factors<D>(c, s, s->poses_try.p, false, s->rec_try.p, &chi_new);
  if (rcode) return rcode;
  const double rho = predicted > 0.0 ? (s->chi_cur - chi_new) / predicted : 0.0;
  const bool accept = chi_new <= s->chi_cur;  // (equality: the step of a converged estimate)
  const double chi_before = s->chi_cur, lambda_used = s->lambda;
  if (accept) {
    std::swap(s->poses.p, s->poses_try.p);
    std::swap(s->rec.p, s->rec_try.p);
    if (chi_new > 0.5 * s->chi_cur) s->cg_tol = std::max(0.1 * s->cg_tol, 1e-8);
    s->chi_cur = chi_new;
    s->need_assemble = true;
    const double t = 2.0 * std::min(std::max(rho, 0.0), 1.0) - 1.0;
    s->lambda *= std::max(1.0 / 3.0, 1.0 - t * t * t);
    s->nu = 2.0;
    s->dx_prev = dx_max;
  } else {
    s->lambda = std::max(s->lambda, 1e-12) * s->nu;
    s->nu *= 2.0;
  }
  CK(c, cudaEventRecord(s->e2, st));
  CK(c, cudaStreamSynchronize(st));
  CK(c, cudaGetLastError());
  (void) POSE;
  if (out) {
    float ms1 = 0.f, ms2 = 0.f;
    CK(c, cudaEventElapsedTime(&ms1, s->e0, s->e1));
    CK(c, cudaEventElapsedTime(&ms2, s->e1, s->e2));
    memset(out, 0, sizeof(*out));
    out->chi = chi_before;
    out->chi_after = chi_new;
    out->dx_norm_inf = dx_max;
    out->cg_iterations = it;
    out->cg_relative_residual = rel;
    out->linearize_ms = ms1;
    out->solve_ms = ms2;
    out->num_blocks = s->nnzb;
    out->num_factors = s->F;
    out->lambda = lambda_used;
    out->gain_ratio = rho;
    out->accepted = accept ? 1 : 0;
  }
  return SRRG2B_OK;
}

}  // namespace

extern "C" {

// FactorGraph upload: variables (LocalMap poses, R/mapping/local_map.h:64,75), gauge mask
// (VariableBase::Fixed, R/system/multi_graph_slam_impl.cpp:85-87) and pose-pose factors
// (R/system/multi_graph_slam_impl.cpp:73-79, R/registration/loop_closure.h:68-78).
// ctx dim 3: 4x4 poses / measurements, 6x6 informations; ctx dim 2: 3x3 poses / measurements, 3x3 informations.
int srrg2b_pgo_upload(srrg2b_ctx* c, int64_t n_vars, const float* poses, const uint8_t* fixed_mask, int64_t n_factors,
                      const int32_t* ij, const float* Zs, const float* Omegas) {
  if (!c) return SRRG2B_ERR_INVALID;
  if (n_vars <= 0 || n_factors < 0 || !poses || !fixed_mask || (n_factors > 0 && (!ij || !Zs || !Omegas)) ||
      n_vars > 0x7fffffff / 36 || n_factors > 0x7fffffff / 36)
    FAIL(c, SRRG2B_ERR_INVALID, "bad pose-graph description");
  CK(c, cudaSetDevice(c->device));
  const int V = (int) n_vars, F = (int) n_factors;
  for (int f = 0; f < F; ++f)
    if (ij[2 * f] < 0 || ij[2 * f] >= V || ij[2 * f + 1] < 0 || ij[2 * f + 1] >= V || ij[2 * f] == ij[2 * f + 1])
      FAIL(c, SRRG2B_ERR_INVALID, "factor references an unknown variable");
  PgoState* s = pgo_of(c, true);
  const int D = c->dim == 3 ? 6 : 3, BB = D * D, POSE = D == 6 ? 12 : 3, REC = 3 * BB + 2 * D, MAT = c->dim == 3 ? 16 : 9;
  if (blocks_for((int64_t) V * D, 192) > kPgoMaxParts) FAIL(c, SRRG2B_ERR_INVALID, "pose graph too large for the ordered reductions");
  s->D = D; s->V = V; s->F = F;
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
  std::vector<int> row_ptr(V + 1, 0), col_idx(nnzb), row_of(nnzb), diag(V, 0);
  for (int k = 0; k < nnzb; ++k) {
    row_ptr[pairs[k].first + 1]++;
    col_idx[k] = pairs[k].second;
    row_of[k] = pairs[k].first;
    if (pairs[k].first == pairs[k].second) diag[pairs[k].first] = k;
  }
  std::partial_sum(row_ptr.begin(), row_ptr.end(), row_ptr.begin());
  auto slot_of = [&](int r, int cc) {
    const int* b = col_idx.data() + row_ptr[r];
    const int* e = col_idx.data() + row_ptr[r + 1];
    return (int) (std::lower_bound(b, e, cc) - col_idx.data());
  };
  // ---- gather lists of THIS rank's factors (f = rank, rank + world, ...; local ordinal l), in factor order ----
  const int n_local = F > c->rank ? (F - c->rank + c->world - 1) / c->world : 0;
  s->n_local = n_local;
  std::vector<int> blk_cnt(nnzb + 1, 0), var_cnt(V + 1, 0);
  for (int l = 0; l < n_local; ++l) {
    const int f = c->rank + l * c->world, i = ij[2 * f], j = ij[2 * f + 1];
    blk_cnt[diag[i] + 1]++; blk_cnt[diag[j] + 1]++; blk_cnt[slot_of(i, j) + 1]++; blk_cnt[slot_of(j, i) + 1]++;
    var_cnt[i + 1]++; var_cnt[j + 1]++;
  }
  std::partial_sum(blk_cnt.begin(), blk_cnt.end(), blk_cnt.begin());
  std::partial_sum(var_cnt.begin(), var_cnt.end(), var_cnt.begin());
  std::vector<int> blk_src((size_t) 4 * n_local + 1), var_src((size_t) 2 * n_local + 1), blk_at(blk_cnt.begin(), blk_cnt.end() - 1),
    var_at(var_cnt.begin(), var_cnt.end() - 1);
  for (int l = 0; l < n_local; ++l) {
    const int f = c->rank + l * c->world, i = ij[2 * f], j = ij[2 * f + 1];
    blk_src[blk_at[diag[i]]++] = 4 * l + 0;
    blk_src[blk_at[slot_of(i, j)]++] = 4 * l + 1;
    blk_src[blk_at[slot_of(j, i)]++] = 4 * l + 2;
    blk_src[blk_at[diag[j]]++] = 4 * l + 3;
    var_src[var_at[i]++] = 2 * l + 0;
    var_src[var_at[j]++] = 2 * l + 1;
  }
  const size_t nD = (size_t) V * D;
  CK(c, s->poses.ensure((size_t) V * POSE)); CK(c, s->poses_try.ensure((size_t) V * POSE)); CK(c, s->fixed.ensure(V));
  CK(c, s->ij.ensure(2 * (size_t) F + 2));
  CK(c, s->Z.ensure((size_t) F * POSE + POSE)); CK(c, s->Omega.ensure((size_t) F * BB + BB));
  CK(c, s->row_ptr.ensure(V + 1)); CK(c, s->col_idx.ensure(nnzb)); CK(c, s->row_of.ensure(nnzb));
  CK(c, s->diag_slot.ensure(V)); CK(c, s->vals.ensure((size_t) nnzb * BB)); CK(c, s->Minv.ensure((size_t) V * BB));
  CK(c, s->blk_ptr.ensure(nnzb + 1)); CK(c, s->blk_src.ensure(blk_src.size())); CK(c, s->var_ptr.ensure(V + 1));
  CK(c, s->var_src.ensure(var_src.size()));
  CK(c, s->rec.ensure((size_t) std::max(n_local, 1) * REC)); CK(c, s->rec_try.ensure((size_t) std::max(n_local, 1) * REC));
  CK(c, s->chi_f.ensure((size_t) std::max(n_local, 1))); CK(c, s->parts.ensure(3 * (size_t) kPgoMaxParts));
  CK(c, s->b.ensure(nD)); CK(c, s->x.ensure(nD)); CK(c, s->r.ensure(nD)); CK(c, s->z.ensure(nD));
  CK(c, s->p.ensure(nD)); CK(c, s->Ap.ensure(nD)); CK(c, s->diag0.ensure(nD));
  CK(c, s->stage.ensure(std::max((size_t) V * 16, (size_t) F * 36) + 16));
  if (!s->d_sc) {
    CK(c, cudaMalloc((void**) &s->d_sc, sizeof(PgoScalars)));
    CK(c, cudaMallocHost((void**) &s->h_sc, sizeof(PgoScalars)));
    CK(c, cudaEventCreate(&s->e0)); CK(c, cudaEventCreate(&s->e1)); CK(c, cudaEventCreate(&s->e2));
  }
  cudaStream_t st = c->stream;
  CK(c, cudaMemcpyAsync(s->stage.p, poses, sizeof(float) * MAT * (size_t) V, cudaMemcpyHostToDevice, st));
  if (D == 6) pgo_pack_poses_kernel<<<blocks_for(V, 256), 256, 0, st>>>(s->stage.p, V, s->poses.p);
  else pgo_pack_poses2_kernel<<<blocks_for(V, 256), 256, 0, st>>>(s->stage.p, V, s->poses.p);
  c->launches++;
  if (F > 0) {
    CK(c, cudaMemcpyAsync(s->stage.p, Zs, sizeof(float) * MAT * (size_t) F, cudaMemcpyHostToDevice, st));
    if (D == 6) pgo_pack_poses_kernel<<<blocks_for(F, 256), 256, 0, st>>>(s->stage.p, F, s->Z.p);
    else pgo_pack_poses2_kernel<<<blocks_for(F, 256), 256, 0, st>>>(s->stage.p, F, s->Z.p);
    CK(c, cudaMemcpyAsync(s->stage.p, Omegas, sizeof(float) * BB * (size_t) F, cudaMemcpyHostToDevice, st));
    pgo_cast_kernel<<<blocks_for((int64_t) F * BB, 256), 256, 0, st>>>(s->stage.p, (size_t) F * BB, s->Omega.p);
    c->launches += 2;
    CK(c, cudaMemcpyAsync(s->ij.p, ij, sizeof(int) * 2 * (size_t) F, cudaMemcpyHostToDevice, st));
  }
  CK(c, cudaMemcpyAsync(s->fixed.p, fixed_mask, V, cudaMemcpyHostToDevice, st));
  CK(c, cudaMemcpyAsync(s->row_ptr.p, row_ptr.data(), sizeof(int) * (V + 1), cudaMemcpyHostToDevice, st));
  CK(c, cudaMemcpyAsync(s->col_idx.p, col_idx.data(), sizeof(int) * (size_t) nnzb, cudaMemcpyHostToDevice, st));
  CK(c, cudaMemcpyAsync(s->row_of.p, row_of.data(), sizeof(int) * (size_t) nnzb, cudaMemcpyHostToDevice, st));
  CK(c, cudaMemcpyAsync(s->diag_slot.p, diag.data(), sizeof(int) * V, cudaMemcpyHostToDevice, st));
  CK(c, cudaMemcpyAsync(s->blk_ptr.p, blk_cnt.data(), sizeof(int) * (size_t) (nnzb + 1), cudaMemcpyHostToDevice, st));
  CK(c, cudaMemcpyAsync(s->blk_src.p, blk_src.data(), sizeof(int) * blk_src.size(), cudaMemcpyHostToDevice, st));
  CK(c, cudaMemcpyAsync(s->var_ptr.p, var_cnt.data(), sizeof(int) * (size_t) (V + 1), cudaMemcpyHostToDevice, st));
  CK(c, cudaMemcpyAsync(s->var_src.p, var_src.data(), sizeof(int) * var_src.size(), cudaMemcpyHostToDevice, st));
  CK(c, cudaStreamSynchronize(st));
  CK(c, cudaGetLastError());
  s->rec_valid = false; s->need_assemble = true; s->assembled_for_rec = nullptr; s->diag_saved = false;
  s->lambda = 1e-4; s->nu = 2.0; s->dx_prev = 1e300; s->cg_tol = 1e-4;
  if (const char* env = getenv("SRRG2B_PGO_TOL0")) s->cg_tol = atof(env);
  if (const char* env = getenv("SRRG2B_PGO_LAMBDA0")) s->lambda = atof(env);
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
  return s->D == 6 ? pgo_iterate_t<6>(c, s, max_cg_iterations, cg_tolerance, out) : pgo_iterate_t<3>(c, s, max_cg_iterations, cg_tolerance, out);
}

// Solver::compute() as MultiGraphSLAM_::optimize() uses it: iterate until the largest perturbation component
// falls below dx_tolerance (or max_iterations).  Levenberg-Marquardt damping with a gain-ratio test guards the
// steps taken far from the optimum (rejected steps count as iterations); the linear solves are inexact PCG.
// stats: one entry per iteration (may be NULL); *n_done: iterations run; returns SRRG2B_OK whether or not the
// tolerance was reached (stats[n_done - 1].dx_norm_inf tells).
int srrg2b_pgo_optimize(srrg2b_ctx* c, int max_iterations, double dx_tolerance, int max_cg_iterations, srrg2b_pgo_stats* stats,
                        int32_t* n_done) {
  if (!c) return SRRG2B_ERR_INVALID;
  PgoState* s = pgo_of(c, false);
  if (!s || s->V == 0) FAIL(c, SRRG2B_ERR_STATE, "no pose graph uploaded");
  if (max_iterations <= 0) max_iterations = 10;
  if (max_cg_iterations <= 0) max_cg_iterations = 2000;
  if (!(dx_tolerance > 0.0)) dx_tolerance = 1e-6;
  CK(c, cudaSetDevice(c->device));
  int done = 0;
  for (int it = 0; it < max_iterations; ++it) {
    srrg2b_pgo_stats st;
    const int rcode = s->D == 6 ? pgo_lm_iterate_t<6>(c, s, max_cg_iterations, 1e-10, &st) : pgo_lm_iterate_t<3>(c, s, max_cg_iterations, 1e-10, &st);
    if (rcode) return rcode;
    if (stats) stats[done] = st;
    ++done;
    // converged: the proposed step is below the tolerance -- and not merely because the damping blew up
    if (st.dx_norm_inf < dx_tolerance && (st.accepted || st.lambda <= 1.0)) break;
  }
  if (n_done) *n_done = done;
  return SRRG2B_OK;
}

int srrg2b_pgo_download(srrg2b_ctx* c, float* poses) {
  if (!c) return SRRG2B_ERR_INVALID;
  PgoState* s = pgo_of(c, false);
  if (!s || s->V == 0 || !poses) FAIL(c, SRRG2B_ERR_STATE, "no pose graph uploaded");
  CK(c, cudaSetDevice(c->device));
  const int MAT = s->D == 6 ? 16 : 9;
  if (s->D == 6) pgo_unpack_poses_kernel<<<blocks_for(s->V, 256), 256, 0, c->stream>>>(s->poses.p, s->V, s->stage.p);
  else pgo_unpack_poses2_kernel<<<blocks_for(s->V, 256), 256, 0, c->stream>>>(s->poses.p, s->V, s->stage.p);
  c->launches++;
  CK(c, cudaMemcpyAsync(poses, s->stage.p, sizeof(float) * MAT * (size_t) s->V, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  return SRRG2B_OK;
}

}  // extern "C"
