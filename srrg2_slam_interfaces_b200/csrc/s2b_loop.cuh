// s2b_loop.cuh -- the streaming lineariser and the persistent device loop of the aligner.
//
//   lin_tiles_body    tile-pipelined pass over a slice's correspondences: the contiguous arrays of the
//                     moving cloud (points, normals, slots, bounds) travel as TMA bulk copies
//                     (cp.async.bulk + mbarrier, UBLKCP) into a 3-stage shared-memory ring, the gathered
//                     fixed points / normals as 16-byte cp.async copies issued one tile ahead; every
//                     thread linearises TWO correspondences at a time with the packed fp32x2 pipeline
//                     (s2b_lin.cuh).  CHECK fuses the exact temporal-coherence test: failures are searched
//                     and linearised in place by the CTA's warps (nn_far_body), overflow goes to a global
//                     work list.
//   lin_tiles_kernel  stand-alone launch of that body (iterations that search with the dedicated NN
//                     kernels, srrg2b_linearize).
//   icp_loop_kernel   ONE cooperative kernel runs all remaining _runSolver iterations
//                     (R/registration/aligners/multi_aligner_impl.cpp:97-128): per iteration the slices'
//                     passes, one grid barrier, the solve step on CTA 0 (all-reduce over the peers included),
//                     release.  No kernel boundary, no host round trip between iterations.
#pragma once
#include "s2b_icp.cuh"

namespace s2b {

constexpr int kLoopThreads = 512;          // one CTA per SM, 16 warps, <= 128 registers per thread
constexpr int kTile = 2 * kLoopThreads;    // correspondences per tile: two per thread
constexpr int kStages = 3;
constexpr int kFailCap = 128;              // coherence-check failures a CTA resolves in place per pass

struct TileStage {
  float4 m[kTile], nm[kTile], f[kTile], nf[kTile];
  int slot[kTile];
  float lb[kTile];
};
constexpr size_t kLoopSmemBytes = sizeof(TileStage) * kStages;

// static shared memory of the tile pass
struct TileCtl {
  LinConst lk;
  FlushSmem fsm;
  unsigned long long full[kStages];  // mbarriers: the bulk copies of a stage have landed
  int fail[kFailCap];
  int nfail;
  int rows[kRowTable];
  float S[16];
  unsigned phase_bits;  // parity of the next wait per stage
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned done = 0;
  while (!done) {
    asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}

// tiles are dealt to the CTAs round robin: CTA b takes tiles b, b + G, b + 2G, ...
// One elected thread issues the bulk copies of tile `t` into stage `st` (count rounded up to a multiple of
// 4 elements: the arrays are padded, sizes must be multiples of 16 bytes).
template <bool CHECK>
__device__ __forceinline__ void tile_issue_bulk(const SliceArgs& a, TileStage* stages, TileCtl& ctl, int st, int tile) {
  const int base = tile * kTile;
  const int cnt = min(kTile, a.nm - base);
  const unsigned c4 = (unsigned) ((cnt + 3) & ~3);
  TileStage& S = stages[st];
  const unsigned bytes = c4 * (16u + 16u + 4u + (CHECK ? 4u : 0u));
  mbar_expect_tx(&ctl.full[st], bytes);
  bulk_g2s(S.m, a.mp + base, c4 * 16u, &ctl.full[st]);
  bulk_g2s(S.nm, a.mn + base, c4 * 16u, &ctl.full[st]);
  bulk_g2s(S.slot, a.c_fpos + base, c4 * 4u, &ctl.full[st]);
  if (CHECK) bulk_g2s(S.lb, a.c_lb + base, c4 * 4u, &ctl.full[st]);
}

// every thread issues the gathers of its two correspondences of the tile in stage st (slots have landed);
// elements beyond the end of the slice (last tile) are neutralised: zero point, no slot, no bound
__device__ __forceinline__ void tile_issue_gather(const SliceArgs& a, TileStage& S, int base, int tid) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int e = tid + h * kLoopThreads;
    int pn = 0;
    if (base + e < a.nm) {
      // no candidate: position 0 stands in (always initialised, finite data; the half is masked later)
      pn = max(slot_candidate(S.slot[e]), 0);
    } else {
      S.m[e] = make_float4(0.f, 0.f, 0.f, 0.f);
      S.nm[e] = make_float4(0.f, 0.f, 0.f, 0.f);
      S.slot[e] = -1;
      S.lb[e] = 0.f;
    }
    cp_async16(&S.f[e], a.fp + pn);
    cp_async16(&S.nf[e], a.fn + pn);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// One pass over the slice: CHECK = temporal-coherence test fused with the linearisation (bounds certified),
// else every slot is linearised as it is.  The CTA's accumulators end up in A (flushed by the caller).
// Requires blockDim.x == kLoopThreads and all threads of the CTA.
template <int DIM, int FACTOR, bool CHECK>
__device__ __forceinline__ void lin_tiles_body(const SliceArgs& a, TileStage* stages, TileCtl& ctl, LinAcc<DIM>& A,
                                               int cta, int n_ctas) {
  const int tid = threadIdx.x;
  const int n_tiles = (a.nm + kTile - 1) / kTile;
  const int my_tiles = cta < n_tiles ? (n_tiles - cta + n_ctas - 1) / n_ctas : 0;
  const LinConst& k = ctl.lk;
  const float bsub = CHECK ? *reinterpret_cast<const volatile float*>(a.S_lb + 17) : 0.f;
  const bool regate = a.gate != 0;  // gated-out slots are re-checked every iteration
  if (my_tiles == 0) return;
  // prologue: bulk copies of the first kStages tiles, gathers of the first
  if (tid == 0) {
    asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy writes (slots, bounds) before the bulk reads
    for (int j = 0; j < kStages && j < my_tiles; ++j) tile_issue_bulk<CHECK>(a, stages, ctl, j, cta + j * n_ctas);
  }
  unsigned phase = ctl.phase_bits;  // (uniform: every thread tracks the same parities)
  mbar_wait(&ctl.full[0], phase & 1u);
  phase ^= 1u;
  tile_issue_gather(a, stages[0], cta * kTile, tid);
  for (int j = 0; j < my_tiles; ++j) {
    const int st = j % kStages;
    TileStage& S = stages[st];
    if (j + 1 < my_tiles) {  // gathers of the next tile (its bulk copies were issued two tiles ago)
      const int sn = (j + 1) % kStages;
      mbar_wait(&ctl.full[sn], (phase >> sn) & 1u);
      phase ^= 1u << sn;
      tile_issue_gather(a, stages[sn], (cta + (j + 1) * n_ctas) * kTile, tid);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    const int base = (cta + j * n_ctas) * kTile;
    const int iA = base + tid, iB = iA + kLoopThreads;
    const bool inA = iA < a.nm, inB = iB < a.nm;
    const int slotA = S.slot[tid], slotB = S.slot[tid + kLoopThreads];
    const int bposA = regate ? slot_candidate(slotA) : (slotA == kSlotSuppressed ? -1 : slotA);
    const int bposB = regate ? slot_candidate(slotB) : (slotB == kSlotSuppressed ? -1 : slotB);
    bool okA = bposA >= 0, okB = bposB >= 0;
    const float4 mA = S.m[tid], mB = S.m[tid + kLoopThreads];
    const float4 nmA = S.nm[tid], nmB = S.nm[tid + kLoopThreads];
    const float4 fA = S.f[tid], fB = S.f[tid + kLoopThreads];
    const float4 nfA = S.nf[tid], nfB = S.nf[tid + kLoopThreads];
    const P3<F2> m = pack3(mA, mB), nm = pack3(nmA, nmB);
    LinGeo<DIM, FACTOR, F2> G;
    lin_geo<DIM, FACTOR, F2>(k, m, nm, pack3(fA, fB), pack3(nfA, nfB), G);
    if (CHECK) {
      // exact temporal coherence (see the NN kernels): keep the neighbour / the "none" verdict when the
      // certified bound minus the motion budget still proves it; else hand the query to the search.
      // The squared distances come from the same packed evaluation that linearises the pair.
      const float lbA = S.lb[tid], lbB = S.lb[tid + kLoopThreads];
      const float lbnA = lbA - bsub, lbnB = lbB - bsub;
      bool failA = false, failB = false;
      if (inA) {
        const bool cert = lbA > 0.f && lbnA > 0.f;
        const bool keep = cert && bposA >= 0 && G.d2.v.x <= a.md2 && G.d2.v.x * (1.f + 1e-5f) < lbnA * lbnA;
        const bool none = cert && slotA == -1 && lbnA * lbnA > a.md2 * (1.f + 1e-5f);
        failA = !(keep || none);
        if (none && a.c_stat) a.c_stat[iA] = SRRG2B_STAT_NONE;
        okA = keep;
      }
      if (inB) {
        const bool cert = lbB > 0.f && lbnB > 0.f;
        const bool keep = cert && bposB >= 0 && G.d2.v.y <= a.md2 && G.d2.v.y * (1.f + 1e-5f) < lbnB * lbnB;
        const bool none = cert && slotB == -1 && lbnB * lbnB > a.md2 * (1.f + 1e-5f);
        failB = !(keep || none);
        if (none && a.c_stat) a.c_stat[iB] = SRRG2B_STAT_NONE;
        okB = keep;
      }
      if (failA | failB) {
        // the first kFailCap failures of the CTA are searched and linearised by its own warps after the
        // tiles (converged iterations: a handful per CTA); the rest go to the global work list
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (!(h ? failB : failA)) continue;
          const int i = h ? iB : iA;
          const int kf = atomicAdd(&ctl.nfail, 1);
          if (kf < kFailCap) ctl.fail[kf] = i;
          else a.work_list[atomicAdd(a.work_count, 1)] = i;
        }
      }
    } else {
      if (inA && slotA == kSlotSuppressed) { A.n_ss += 1; if (a.c_stat) a.c_stat[iA] = SRRG2B_STAT_SUPPRESSED; }
      else if (inA && !okA && a.c_stat) a.c_stat[iA] = SRRG2B_STAT_NONE;
      if (inB && slotB == kSlotSuppressed) { A.n_ss += 1; if (a.c_stat) a.c_stat[iB] = SRRG2B_STAT_SUPPRESSED; }
      else if (inB && !okB && a.c_stat) a.c_stat[iB] = SRRG2B_STAT_NONE;
    }
    if (__any_sync(0xffffffffu, okA | okB)) {
      PairOut o;
      if (lin_pair_finish<DIM, FACTOR>(k, G, m, nm, okA, okB, A, o)) {
        if (okA) {
          if (regate && o.gateA != (slotA >= 0)) a.c_fpos[iA] = o.gateA ? bposA : -(bposA + 2);
          if (a.c_stat) a.c_stat[iA] = (unsigned char) o.statA;
          if (a.c_chi && o.gateA) a.c_chi[iA] = o.chiA;
        }
        if (okB) {
          if (regate && o.gateB != (slotB >= 0)) a.c_fpos[iB] = o.gateB ? bposB : -(bposB + 2);
          if (a.c_stat) a.c_stat[iB] = (unsigned char) o.statB;
          if (a.c_chi && o.gateB) a.c_chi[iB] = o.chiB;
        }
      } else {  // a non-finite chi in the pair (overflowing / NaN input): one correspondence at a time
        if (okA) lin_one_slot<DIM, FACTOR>(a, k, iA, slotA, bposA, mA, nmA, fA, nfA, A);
        if (okB) lin_one_slot<DIM, FACTOR>(a, k, iB, slotB, bposB, mB, nmB, fB, nfB, A);
      }
    }
    // the stage is free once every thread has read it: refill it with the tile kStages ahead
    if (j + kStages < my_tiles) {
      __syncthreads();
      if (tid == 0) tile_issue_bulk<CHECK>(a, stages, ctl, st, cta + (j + kStages) * n_ctas);
    }
    if ((j & 127) == 127) {  // 32-bit partial sums: a thread stays below 512 terms per flush
      lin_flush<DIM>(a.acc, false, A, ctl.fsm);
      A.clear();
    }
  }
  if (tid == 0) ctl.phase_bits = phase;
}

// the CTA's control block: mbarriers of the ring, lineariser constants of the slice about to be processed
__device__ __forceinline__ void tile_ctl_init(TileCtl& ctl) {
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&ctl.full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    ctl.phase_bits = 0;
    ctl.nfail = 0;
  }
}

// ---------------------------------------------------------------------------------------------
// stand-alone pass: linearise every slot of the slice as it is (iterations whose correspondences come
// from the dedicated NN kernels; srrg2b_linearize)
// ---------------------------------------------------------------------------------------------
template <int DIM, int FACTOR>
__global__ void __launch_bounds__(kLoopThreads, 1) lin_tiles_kernel(const SliceArgs a, const int* skip) {
  if (*a.stop || (skip && *skip)) return;
  extern __shared__ __align__(128) unsigned char loop_smem_raw[];
  TileStage* stages = reinterpret_cast<TileStage*>(loop_smem_raw);
  __shared__ TileCtl ctl;
  tile_ctl_init(ctl);
  if (threadIdx.x == 32) make_lin_const(a, a.S, ctl.lk);
  __syncthreads();
  LinAcc<DIM> A;
  A.clear();
  lin_tiles_body<DIM, FACTOR, false>(a, stages, ctl, A, blockIdx.x, gridDim.x);
  lin_flush<DIM>(a.acc, a.few_terms != 0, A, ctl.fsm);
}

// ---------------------------------------------------------------------------------------------
// the persistent device loop
// ---------------------------------------------------------------------------------------------
struct LoopArgs {
  const SolveArgs* ap;
  DevState* st;
  const PeerExchange* px;
  GridBar* bar;
  long long timeout_cycles;
  int n_slices;
  int factor[SRRG2B_MAX_SLICES];
  int is_points[SRRG2B_MAX_SLICES];
  SliceArgs sl[SRRG2B_MAX_SLICES];
};

__device__ __forceinline__ unsigned ld_volatile_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

struct LoopSync {
  GridBar* bar;
  int* error;
  long long timeout;
  unsigned n_ctas, seq;  // seq: barriers completed so far (identical on every CTA)
  int* bcast;            // one shared word

  // every CTA: everything this CTA wrote is visible before the arrival counts
  __device__ __forceinline__ void arrive() {
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence(); atomicAdd(&bar->count, 1u); }
    ++seq;
  }
  // leader CTA: wait until every CTA has arrived at barrier `seq`
  __device__ __forceinline__ bool wait_all() {
    if (threadIdx.x == 0) {
      const unsigned target = n_ctas * seq;
      const long long t0 = clock64();
      int ok = 1, spins = 0;
      while (ld_volatile_u32(&bar->count) < target) {
        if ((++spins & 255) == 0 && (ld_volatile_u32(&bar->abort) || clock64() - t0 > timeout)) { ok = 0; break; }
      }
      __threadfence();
      if (!ok) { *error = 2; atomicExch(&bar->abort, 1u); __threadfence(); }
      *bcast = ok;
    }
    __syncthreads();
    const bool ok = *bcast != 0;
    __syncthreads();
    return ok;
  }
  // leader CTA: let everybody pass barrier `seq`; mode travels with the release
  __device__ __forceinline__ void release(unsigned mode) {
    __syncthreads();
    if (threadIdx.x == 0) {
      bar->mode = mode;
      __threadfence();
      atomicExch(&bar->gen, seq);
    }
  }
  // every CTA: wait for the release of barrier `seq`; returns the mode, or -1 on abort / timeout
  __device__ __forceinline__ int wait_release() {
    if (threadIdx.x == 0) {
      const long long t0 = clock64();
      int ok = 1, spins = 0;
      while (ld_volatile_u32(&bar->gen) < seq) {
        if ((++spins & 255) == 0 && (ld_volatile_u32(&bar->abort) || clock64() - t0 > timeout)) { ok = 0; break; }
      }
      __threadfence();
      if (!ok) { *error = 2; atomicExch(&bar->abort, 1u); __threadfence(); }
      *bcast = ok ? (int) ld_volatile_u32(&bar->mode) : -1;
    }
    __syncthreads();
    const int m = *bcast;
    __syncthreads();
    return m;
  }
  // plain barrier (leader releases as soon as everybody has arrived)
  __device__ __forceinline__ bool sync_all(bool leader) {
    arrive();
    if (leader) {
      if (!wait_all()) return false;
      release(0);
    }
    return wait_release() >= 0;
  }
};

// projective association of the whole slice (grid-stride), see proj_find_kernel
__device__ __forceinline__ void proj_find_body(const SliceArgs& a, const float* S) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.nm; i += gridDim.x * blockDim.x) {
    NNQuery q;
    nn_transform<3>(S, a.mp[i], q.qx, q.qy, q.qz);
    q.bd2 = a.md2; q.sd2 = a.md2; q.bidx = INT_MAX; q.bpos = -1;
    int pix;
    if (project_pixel(q.qx, q.qy, q.qz, a.fx, a.fy, a.pcx, a.pcy, a.min_depth, a.max_depth, a.width, a.height, pix)) {
      const unsigned long long key = __ldg(a.image + pix);
      if (key != ~0ull) {
        const int idx = (int) (key & 0xffffffffull);
        const float4 c = __ldg(a.fp + idx);
        const float ddx = q.qx - c.x, ddy = q.qy - c.y, ddz = q.qz - c.z;
        const float d2 = fmaf(ddz, ddz, fmaf(ddy, ddy, ddx * ddx));
        if (d2 <= a.md2) { q.bd2 = d2; q.bidx = idx; q.bpos = idx; }
      }
    }
    nn_finish<3>(a, S, q, i, 0.f, __ldcg(a.c_fpos + i));
  }
}

// lineariser constants of slice s into the control block (S from the device state, coherently)
__device__ __forceinline__ void loop_load_lin_const(const SliceArgs& a, const DevState* st, int s, TileCtl& ctl) {
  const int tid = threadIdx.x;
  if (tid == 0) ctl.nfail = 0;
  if (tid < 12) ctl.lk.S[tid] = __ldcg(&st->S[s].m[tid]);
  if (tid == 32) {
    LinConst& k = ctl.lk;
    for (int i = 0; i < kKCount; ++i) k.fS[i] = a.fS[i];
    k.fSinvChi = a.fSinvChi;
    k.ip = a.ip; k.in_ = a.in_; k.rs = a.rs; k.tau = a.tau; k.delta = a.delta;
    k.normal_cos = a.normal_cos; k.eb2 = a.eb2; k.rob = a.rob; k.gate = a.gate;
  }
  __syncthreads();
}

// One warp per listed query: search, slot + bound, linearisation; the sums go to the slice's accumulators.
// (cold path of the loop kernel: its own accumulator registers and flush, kept out of line)
template <int DIM, int FACTOR>
__device__ __noinline__ void loop_search_list(const SliceArgs& a, TileCtl& ctl, int track2, int n, const int* list, int w0, int ws) {
  const float cell = __fdiv_rn(1.f, a.inv_cell);
  const int K = (DIM == 3) ? (2 * a.R + 1) * (2 * a.R + 1) : (2 * a.R + 1);
  LinAcc<DIM> A;
  A.clear();
  if (track2) nn_far_body<DIM, true, FACTOR>(a, ctl.lk.S, ctl.rows, K, cell, n, list, &A, &ctl.lk, w0, ws);
  else nn_far_body<DIM, false, FACTOR>(a, ctl.lk.S, ctl.rows, K, cell, n, list, &A, &ctl.lk, w0, ws);
  lin_flush<DIM>(a.acc, false, A, ctl.fsm);
}

// one pass of the streaming lineariser over a slice inside the loop kernel, flush included
template <int DIM, int FACTOR, bool CHECK>
__device__ __forceinline__ void loop_lin_pass(const SliceArgs& a, TileStage* stages, TileCtl& ctl, int track2) {
  LinAcc<DIM> A;
  A.clear();
  lin_tiles_body<DIM, FACTOR, CHECK>(a, stages, ctl, A, blockIdx.x, gridDim.x);
  lin_flush<DIM>(a.acc, a.few_terms != 0, A, ctl.fsm);
  if (CHECK) {
    // the CTA's coherence-check failures: one warp per query (search, slot + bound, linearisation)
    const int n_local = min(ctl.nfail, kFailCap);
    if (n_local > 0) loop_search_list<DIM, FACTOR>(a, ctl, track2, n_local, ctl.fail, threadIdx.x >> 5, blockDim.x >> 5);
  }
}

template <int DIM, int FACTOR>
__device__ __noinline__ void loop_lin_all(const SliceArgs& a, TileStage* stages, TileCtl& ctl) {
  loop_lin_pass<DIM, FACTOR, false>(a, stages, ctl, 0);
}

// first phase of a from-scratch search of the whole slice (rings 0-1, thread per query)
template <int DIM>
__device__ __noinline__ void loop_search_phase1(const SliceArgs& a, TileCtl& ctl, int track2) {
  const float cell = __fdiv_rn(1.f, a.inv_cell);
  const float ring2 = (a.R >= 2) ? (1.f - 4e-3f) * cell : __fsqrt_rn(a.rho_s2);
  const float ring2_sq = (a.R >= 2) ? ring2 * ring2 : 3.0e38f;
  if (track2) nn_phase1_body<DIM, true>(a, ctl.lk.S, 0.f, cell, ring2, ring2_sq, true, a.nm);
  else nn_phase1_body<DIM, false>(a, ctl.lk.S, 0.f, cell, ring2, ring2_sq, true, a.nm);
}

// second phase: the queries rings 0-1 did not settle
template <int DIM>
__device__ __noinline__ void loop_search_far(const SliceArgs& a, TileCtl& ctl, int track2, int n_far) {
  const float cell = __fdiv_rn(1.f, a.inv_cell);
  const int K = (DIM == 3) ? (2 * a.R + 1) * (2 * a.R + 1) : (2 * a.R + 1);
  const int wpb = blockDim.x >> 5, w0 = blockIdx.x * wpb + (threadIdx.x >> 5), ws = gridDim.x * wpb;
  if (track2) nn_far_body<DIM, true>(a, ctl.lk.S, ctl.rows, K, cell, n_far, a.far_list, nullptr, nullptr, w0, ws);
  else nn_far_body<DIM, false>(a, ctl.lk.S, ctl.rows, K, cell, n_far, a.far_list, nullptr, nullptr, w0, ws);
}

template <int DIM>
__device__ __noinline__ void loop_solve(const LoopArgs& L, SolveSmem& ssm) {
  icp_solve_block<DIM>(L.ap, L.st, L.px, ssm);
}

__device__ __noinline__ void loop_proj_find(const SliceArgs& a, const float* S) { proj_find_body(a, S); }

template <int DIM>
__global__ void __launch_bounds__(kLoopThreads, 1) icp_loop_kernel(const __grid_constant__ LoopArgs L) {
  extern __shared__ __align__(128) unsigned char loop_smem_raw[];
  TileStage* stages = reinterpret_cast<TileStage*>(loop_smem_raw);
  // the solve step's staging area lives in the last ring stage (no tile is in flight while CTA 0 solves)
  SolveSmem& ssm = *reinterpret_cast<SolveSmem*>(&stages[kStages - 1]);
  static_assert(sizeof(SolveSmem) <= sizeof(TileStage), "solve staging must fit a ring stage");
  __shared__ TileCtl ctl;
  __shared__ int s_ctrl[4 + 2 * SRRG2B_MAX_SLICES];  // stop, -, -, bcast | list_all[s] | track2[s]
  const int tid = threadIdx.x;
  const bool leader = blockIdx.x == 0;
  DevState* st = L.st;
  tile_ctl_init(ctl);
  LoopSync sync{L.bar, &st->error, L.timeout_cycles, gridDim.x, 0u, &s_ctrl[3]};
  constexpr int KMAX = (DIM == 3) ? kRowTable : (2 * kMaxR + 1);
  for (int k = tid; k < KMAX; k += blockDim.x)
    ctl.rows[k] = (DIM == 3) ? *reinterpret_cast<const int*>(c_rows3[k]) : *reinterpret_cast<const int*>(c_rows2[k]);
  __syncthreads();

  for (;;) {
    // ---- state of this iteration (written by the solve step of the previous one) ----
    if (tid == 0) s_ctrl[0] = __ldcg(&st->stop);
    if (tid < L.n_slices) {
      s_ctrl[4 + tid] = __ldcg(&st->list_all[tid]);
      s_ctrl[4 + SRRG2B_MAX_SLICES + tid] = __ldcg(&st->track2[tid]);
    }
    __syncthreads();
    if (s_ctrl[0]) break;
    bool fallback = false;  // some slice needs the full search path this iteration (uniform over the grid)

    // ---- phase A: certified slices: coherence check + linearisation; others: first search phase ----
    for (int s = 0; s < L.n_slices; ++s) {
      if (!L.is_points[s]) continue;
      const SliceArgs& a = L.sl[s];
      if (a.nm <= 0) continue;
      const int track2 = s_ctrl[4 + SRRG2B_MAX_SLICES + s];
      loop_load_lin_const(a, st, s, ctl);
      if (a.projective) {
        fallback = true;
        loop_proj_find(a, ctl.lk.S);
      } else if (s_ctrl[4 + s]) {  // no certified bounds: search everything, rings 0-1 here, the rest after the barrier
        fallback = true;
        loop_search_phase1<DIM>(a, ctl, track2);
      } else if (L.factor[s] == SRRG2B_FACTOR_P2P) {
        loop_lin_pass<DIM, SRRG2B_FACTOR_P2P, true>(a, stages, ctl, track2);
      } else {
        loop_lin_pass<DIM, SRRG2B_FACTOR_PLANE, true>(a, stages, ctl, track2);
      }
      __syncthreads();
    }

    // ---- barrier; in the all-certified case CTA 0 goes straight to the solve step ----
    sync.arrive();
    if (leader) {
      if (!sync.wait_all()) break;
      bool lists = false;
      if (!fallback) {
        for (int s = 0; s < L.n_slices; ++s)
          if (L.is_points[s] && L.sl[s].nm > 0 && !L.sl[s].projective && __ldcg(L.sl[s].work_count) > 0) lists = true;
      }
      if (!fallback && !lists) {
        loop_solve<DIM>(L, ssm);
        sync.release(0);
      } else {
        sync.release(1);
      }
    }
    const int mode = sync.wait_release();
    if (mode < 0) break;
    if (mode == 0) continue;

    // ---- phase B (rare): overflow work lists of certified slices, far phase of the full searches ----
    bool far_phase = false;
    for (int s = 0; s < L.n_slices; ++s) {
      if (!L.is_points[s]) continue;
      const SliceArgs& a = L.sl[s];
      if (a.nm <= 0 || a.projective) continue;
      const int track2 = s_ctrl[4 + SRRG2B_MAX_SLICES + s];
      loop_load_lin_const(a, st, s, ctl);
      if (s_ctrl[4 + s]) {
        far_phase = true;
        const int n_far = __ldcg(a.far_count);
        if (n_far > 0) loop_search_far<DIM>(a, ctl, track2, n_far);
      } else {
        const int n_work = __ldcg(a.work_count);
        const int wpb = blockDim.x >> 5, w0 = blockIdx.x * wpb + (tid >> 5), ws = gridDim.x * wpb;
        if (n_work > 0) {
          if (L.factor[s] == SRRG2B_FACTOR_P2P) loop_search_list<DIM, SRRG2B_FACTOR_P2P>(a, ctl, track2, n_work, a.work_list, w0, ws);
          else loop_search_list<DIM, SRRG2B_FACTOR_PLANE>(a, ctl, track2, n_work, a.work_list, w0, ws);
        }
      }
      __syncthreads();
    }
    if (far_phase && !sync.sync_all(leader)) break;

    // ---- phase C (rare): linearise the slices that were searched from scratch ----
    for (int s = 0; s < L.n_slices; ++s) {
      if (!L.is_points[s]) continue;
      const SliceArgs& a = L.sl[s];
      if (a.nm <= 0 || !(a.projective || s_ctrl[4 + s])) continue;
      loop_load_lin_const(a, st, s, ctl);
      if (L.factor[s] == SRRG2B_FACTOR_P2P) loop_lin_all<DIM, SRRG2B_FACTOR_P2P>(a, stages, ctl);
      else loop_lin_all<DIM, SRRG2B_FACTOR_PLANE>(a, stages, ctl);
      __syncthreads();
    }
    sync.arrive();
    if (leader) {
      if (!sync.wait_all()) break;
      loop_solve<DIM>(L, ssm);
      sync.release(0);
    }
    if (sync.wait_release() < 0) break;
  }
}

}  // namespace s2b
