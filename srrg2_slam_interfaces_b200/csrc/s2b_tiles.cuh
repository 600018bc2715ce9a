// s2b_tiles.cuh -- the streaming lineariser of the aligner and the kernels of one _runSolver iteration built on it.
//
//   lin_tiles_body    tile-pipelined pass over a slice's correspondences: the contiguous arrays of the
//                     moving cloud (pair records, slots, bounds) travel as TMA bulk copies (cp.async.bulk +
//                     mbarrier, UBLKCP) into a 4-stage shared-memory ring PER WARP, the gathered fixed
//                     points / normals as 16-byte cp.async copies issued one tile ahead; every thread
//                     linearises TWO correspondences at a time with the packed fp32x2 pipeline (s2b_lin.cuh).
//                     CHECK fuses the exact temporal-coherence test: the first failures of a CTA are kept as
//                     64-byte records and searched + linearised by its warps after the tiles, the rest go to a
//                     global work list.
//   check_tiles_kernel        certified slice: coherence check + linearisation of every correspondence
//   lin_after_search_kernel   linearises what the NN kernels searched (everything / a long work list)
//   lin_tiles_kernel          linearises every slot as it is (srrg2b_linearize, projective slices)
// One iteration (R/registration/aligners/multi_aligner_impl.cpp:97-128) is the launch sequence
// check_tiles -> nn_kernel -> nn_far_kernel -> lin_after_search -> icp_solve_kernel; every kernel decides from
// the device-side control words what it has to do.  (A persistent cooperative kernel that ran all iterations
// behind a grid barrier was measured slower -- 1.51 ms against 1.34 ms per 20-iteration C2 run: grid barrier,
// in-kernel solve and release cost 28 us per iteration against ~10 us of launch boundaries -- and was removed;
// see DESIGN.md section 5.)
#pragma once
#include "s2b_icp.cuh"

namespace s2b {

#ifndef S2B_LOOP_THREADS
#define S2B_LOOP_THREADS 384
#define S2B_LOOP_STAGES 4
#define S2B_LOOP_PPL 1
#endif
#ifndef S2B_FAIL_ROUNDS
#define S2B_FAIL_ROUNDS 2  // failures of a check pass searched in place: at most this many rounds of the CTA's warps
#endif
// One CTA of 12 warps per SM: the lineariser wants registers (35 accumulators + ~90 temporaries of a pair),
// not warps.  Measured on C2 (pass over 1M correspondences inside the loop kernel): 512 threads x 128
// registers spill the accumulators: 30 us; 384 x 168: 19-21 us; 256 x 255: 21 us; two independent pairs
// per lane (256 x 255, kPPL = 2): 24-28 us.
constexpr int kLoopThreads = S2B_LOOP_THREADS;
constexpr int kLoopWarps = kLoopThreads / 32;
constexpr int kPPL = S2B_LOOP_PPL;         // pairs per lane and tile
constexpr int kSubTile = 64;               // correspondences of one sub-tile: one pair per lane
constexpr int kWTile = kSubTile * kPPL;    // correspondences per warp tile
constexpr int kStages = S2B_LOOP_STAGES;
constexpr int kFailInPlace = S2B_FAIL_ROUNDS * (S2B_LOOP_THREADS / 32);  // ... unless the CTA saw more failures than two rounds of its warps
constexpr int kFailCap = 48;               // coherence-check failures a CTA resolves in place per pass (one warp per
                                           // query); the rest go to the global work list

// One sub-tile (one pair per lane) of one ring stage of ONE warp (4 KB).  Every warp runs its own pipeline
// over its own tiles, so the steady state has no CTA-wide barrier and the warps drift apart freely.
struct SubTile {
  float4 mp[kSubTile / 2 * 3];     // pair records of the moving cloud (see pair_pack_kernel): lane l owns [3l, 3l + 3)
  float4 f[kSubTile];              // gathered fixed points: lane l owns [l] (first of its pair) and [32 + l]
  float4 nf[kSubTile];             // gathered fixed normals: lane l owns [l] and [32 + l]
  int slot[kSubTile];              // lane l owns [2l], [2l + 1]
  float lb[kSubTile];
};
static_assert(sizeof(SubTile) == 4096, "one sub-tile is 4 KB");
struct WarpStage {
  SubTile sub[kPPL];
};
struct TileStage {
  WarpStage w[kLoopWarps];
};
constexpr size_t kLoopSmemBytes = sizeof(TileStage) * kStages;
static_assert(kLoopSmemBytes <= 200 * 1024, "ring does not fit");

// a query that failed the coherence check, with everything its tile held about it (64 bytes)
struct FailRec {
  float4 m, nm, f, nf;  // moving point (.w: its index, integer bits) / normal (.w: old slot), the slot's fixed
                        // point (.w: original index) / normal
};
// static shared memory of the tile pass
struct TileCtl {
  LinConst lk;
  FlushSmem fsm;
  unsigned long long full[kLoopWarps][kStages];  // mbarriers: the bulk copies of a warp's stage have landed
  int nfail;                        // coherence-check failures of the CTA in the running pass
  FailRec rec[kFailCap];            // the failures the CTA resolves itself after its tiles (one warp per query)
  int nrec;                         // ... how many
  int ep_cur;                       // id of the epoch this pass certifies bounds for
  int inline_ok;                    // this pass: the warps resolve their (few) failures themselves
  int rows[kRowTable];
  unsigned phase_bits[kLoopWarps];  // parity of the next wait per stage, per warp (survives between passes)
  long long tail[kAcc];  // sums of the correspondences resolved by the CTA's search warps (lin_push_tail)
  float2 dtab[kEpochs];  // displacement table of the slice's epochs (see encode_bound), refreshed per pass: a query
                         // m has moved at most dtab[e].x |m| + dtab[e].y since epoch e
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

// One elected lane issues the bulk copies (TMA) of warp tile `tile` into the warp's stage, sub-tile by
// sub-tile (slot / bound counts rounded up to a multiple of 4 elements: the arrays are padded, sizes must be
// multiples of 16 bytes).  One mbarrier phase covers the whole warp tile.
template <bool CHECK>
__device__ __forceinline__ void tile_issue_bulk(const SliceArgs& a, WarpStage& S, unsigned long long* bar, int tile) {
  unsigned cnt[kPPL], bytes = 0;
#pragma unroll
  for (int q = 0; q < kPPL; ++q) {
    const int base = tile * kWTile + q * kSubTile;
    cnt[q] = (unsigned) max(0, min(kSubTile, a.nm - base));
    const unsigned c4 = (cnt[q] + 3u) & ~3u, pairs = (cnt[q] + 1u) >> 1;
    bytes += pairs * 48u + c4 * (4u + (CHECK ? 4u : 0u));
  }
  mbar_expect_tx(bar, bytes);
#pragma unroll
  for (int q = 0; q < kPPL; ++q) {
    if (cnt[q] == 0) continue;
    const int base = tile * kWTile + q * kSubTile;
    const unsigned c4 = (cnt[q] + 3u) & ~3u, pairs = (cnt[q] + 1u) >> 1;
    bulk_g2s(S.sub[q].mp, a.mpair + (size_t) (base >> 1) * 3, pairs * 48u, bar);
    bulk_g2s(S.sub[q].slot, a.c_fpos + base, c4 * 4u, bar);
    if (CHECK) bulk_g2s(S.sub[q].lb, a.c_lb + base, c4 * 4u, bar);
  }
}

// every lane issues the gathers of its correspondences of the warp tile (slots have landed); elements beyond
// the end of the slice (last tile) are neutralised: zero point, no slot, no bound
__device__ __forceinline__ void tile_issue_gather(const SliceArgs& a, WarpStage& W, int tile_base, int lane) {
#pragma unroll
  for (int q = 0; q < kPPL; ++q) {
    SubTile& S = W.sub[q];
    const int i0 = tile_base + q * kSubTile + 2 * lane;
    int2 sl = *reinterpret_cast<const int2*>(&S.slot[2 * lane]);
    if (i0 + 1 >= a.nm) {
      if (i0 >= a.nm) {
        sl.x = -1;
        S.mp[3 * lane] = S.mp[3 * lane + 1] = S.mp[3 * lane + 2] = make_float4(0.f, 0.f, 0.f, 0.f);
        S.lb[2 * lane] = 0.f;
      }
      sl.y = -1;
      S.lb[2 * lane + 1] = 0.f;
      *reinterpret_cast<int2*>(&S.slot[2 * lane]) = sl;
    }
    // no candidate: zeros stand in (finite data; the half is masked later).  Written by the lane itself -- a
    // gather of some fixed stand-in record by every such lane of the grid would hammer one L2 sector.
    const int pa = slot_candidate(sl.x), pb = slot_candidate(sl.y);
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pa >= 0) {
      cp_async16(&S.f[lane], a.frec + 2 * (size_t) pa);
      cp_async16(&S.nf[lane], a.frec + 2 * (size_t) pa + 1);
    } else {
      S.f[lane] = z4; S.nf[lane] = z4;
    }
    if (pb >= 0) {
      cp_async16(&S.f[32 + lane], a.frec + 2 * (size_t) pb);
      cp_async16(&S.nf[32 + lane], a.frec + 2 * (size_t) pb + 1);
    } else {
      S.f[32 + lane] = z4; S.nf[32 + lane] = z4;
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

#ifdef S2B_FAIL_STATS
__device__ unsigned long long g_fail_stats[8];
#endif

// everything one lane holds of one pair between the two halves of its evaluation
template <int DIM, int FACTOR>
struct PairWork {
  P3<F2> m, nm;
  NF2 nf;
  float4 fA, fB;
  LinGeo<DIM, FACTOR, F2> G;
  int slotA, slotB, bposA, bposB, iA;
  bool okA, okB;
};

// One pass over the slice: CHECK = temporal-coherence test fused with the linearisation (bounds certified),
// else every slot is linearised as it is.  Every warp streams its own tiles (warp w of CTA b takes tiles
// g, g + W, g + 2W, ... with g = w * gridDim + b, W = all warps of the grid) through its own ring; every lane
// evaluates kPPL independent pairs per tile back to back, so that their instruction streams interleave.
// The warp's accumulators end up in A (flushed by the caller); returns the number of terms a thread may have added
// to a slot since A was last cleared.  Requires blockDim.x == kLoopThreads.
// CHECK: the first failures of the CTA are recorded in ctl.rec (the caller has them searched once the tiles are
// done); beyond the cap they go to the global work list.
// the bulk copies of a warp's first kStages tiles: issued by the kernels BEFORE they load the lineariser constants
// (nothing here depends on the transform), so the first tiles are in flight meanwhile
template <bool CHECK>
__device__ __forceinline__ void lin_tiles_prologue(const SliceArgs& a, TileStage* stages, TileCtl& ctl) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_tiles = (a.nm + kWTile - 1) / kWTile;
  const int W = gridDim.x * kLoopWarps, g0 = warp * gridDim.x + blockIdx.x;
  const int my_tiles = g0 < n_tiles ? (n_tiles - g0 + W - 1) / W : 0;
  if (lane == 0 && my_tiles > 0) {
    asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy writes (slots, bounds) before the bulk reads
    for (int j = 0; j < kStages && j < my_tiles; ++j) tile_issue_bulk<CHECK>(a, stages[j].w[warp], &ctl.full[warp][j], g0 + j * W);
  }
}

template <int DIM, int FACTOR, bool CHECK>
__device__ __forceinline__ int lin_tiles_body(const SliceArgs& a, TileStage* stages, TileCtl& ctl, LinAcc<DIM>& A) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_tiles = (a.nm + kWTile - 1) / kWTile;
  const int W = gridDim.x * kLoopWarps, g0 = warp * gridDim.x + blockIdx.x;
  const int my_tiles = g0 < n_tiles ? (n_tiles - g0 + W - 1) / W : 0;
  const LinConst& k = ctl.lk;
  const bool regate = a.gate != 0;  // gated-out slots are re-checked every iteration
  if (my_tiles == 0) return 0;
  unsigned long long* bars = ctl.full[warp];
  // (the bulk copies of the first kStages tiles were issued by lin_tiles_prologue); gathers of the first tile
  unsigned phase = ctl.phase_bits[warp];
  mbar_wait(&bars[0], phase & 1u);
  phase ^= 1u;
  tile_issue_gather(a, stages[0].w[warp], g0 * kWTile, lane);
  for (int j = 0; j < my_tiles; ++j) {
    const int st = j % kStages;
    WarpStage& WS = stages[st].w[warp];
    if (j + 1 < my_tiles) {  // gathers of the next tile (its bulk copies were issued two tiles ago)
      const int sn = (j + 1) % kStages;
      mbar_wait(&bars[sn], (phase >> sn) & 1u);
      phase ^= 1u << sn;
      tile_issue_gather(a, stages[sn].w[warp], (g0 + (j + 1) * W) * kWTile, lane);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    const int base = (g0 + j * W) * kWTile;
    PairWork<DIM, FACTOR> pw[kPPL];
    // ---- first half, all pairs: loads, geometry (transform, residual, chi), coherence verdicts ----
#pragma unroll
    for (int q = 0; q < kPPL; ++q) {
      PairWork<DIM, FACTOR>& w = pw[q];
      const SubTile& S = WS.sub[q];
      w.iA = base + q * kSubTile + 2 * lane;
      const int2 sl = *reinterpret_cast<const int2*>(&S.slot[2 * lane]);
      w.slotA = sl.x; w.slotB = sl.y;
      w.bposA = regate ? slot_candidate(sl.x) : (sl.x == kSlotSuppressed ? -1 : sl.x);
      w.bposB = regate ? slot_candidate(sl.y) : (sl.y == kSlotSuppressed ? -1 : sl.y);
      w.okA = w.bposA >= 0; w.okB = w.bposB >= 0;
      const float4 Q0 = S.mp[3 * lane], Q1 = S.mp[3 * lane + 1], Q2 = S.mp[3 * lane + 2];
      w.fA = S.f[lane]; w.fB = S.f[32 + lane];
      w.nf = NF2{S.nf[lane], S.nf[32 + lane]};
      w.m = P3<F2>{F2{make_float2(Q0.x, Q0.y)}, F2{make_float2(Q0.z, Q0.w)}, F2{make_float2(Q1.x, Q1.y)}};
      w.nm = P3<F2>{F2{make_float2(Q1.z, Q1.w)}, F2{make_float2(Q2.x, Q2.y)}, F2{make_float2(Q2.z, Q2.w)}};
      lin_geo<DIM, FACTOR, F2, NF2>(k, w.m, w.nm, pack3(w.fA, w.fB), w.nf, w.G);
    }
#pragma unroll
    for (int q = 0; q < kPPL; ++q) {
      PairWork<DIM, FACTOR>& w = pw[q];
      const SubTile& S = WS.sub[q];
      const int iA = w.iA, iB = iA + 1;
      if (CHECK) {
        // exact temporal coherence (see the NN kernels): keep the neighbour / the "none" verdict when the
        // certified bound minus the motion budget still proves it; else hand the query to the search.
        // The squared distances come from the same packed evaluation that linearises the pair.
        const float2 lb2 = *reinterpret_cast<const float2*>(&S.lb[2 * lane]);
        const float lbA = lb2.x, lbB = lb2.y;
        // |m| of the two queries, rounded up (approximate root: 2 ulp, covered by the factor)
        const F2 m2 = t_fma(w.m.z, w.m.z, t_fma(w.m.y, w.m.y, t_mul(w.m.x, w.m.x)));
        float rA, rB;
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rA) : "f"(m2.v.x));
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rB) : "f"(m2.v.y));
        rA = rA * (1.f + 1e-6f) + 1e-30f; rB = rB * (1.f + 1e-6f) + 1e-30f;
        float lbnA = 0.f, lbnB = 0.f;
        if (lbA > 0.f) {
          const int b = __float_as_int(lbA);
          const float2 d = ctl.dtab[b & (kEpochs - 1)];
          lbnA = __int_as_float(b & ~(kEpochs - 1)) - fmaf(d.x, rA, d.y) * (1.f + 2.4e-7f);
        }
        if (lbB > 0.f) {
          const int b = __float_as_int(lbB);
          const float2 d = ctl.dtab[b & (kEpochs - 1)];
          lbnB = __int_as_float(b & ~(kEpochs - 1)) - fmaf(d.x, rB, d.y) * (1.f + 2.4e-7f);
        }
        const bool inA = iA < a.nm, inB = iB < a.nm;
        bool failA = false, failB = false;
        if (inA) {
          const bool cert = lbA > 0.f && lbnA > 0.f;
          const bool keep = cert && w.bposA >= 0 && w.G.d2.v.x <= a.md2 && w.G.d2.v.x * (1.f + 1e-5f) < lbnA * lbnA;
          const bool none = cert && w.slotA == -1 && lbnA * lbnA > a.md2 * (1.f + 1e-5f);
          failA = !(keep || none);
          if (none && a.c_stat) a.c_stat[iA] = SRRG2B_STAT_NONE;
          w.okA = keep;
        }
        if (inB) {
          const bool cert = lbB > 0.f && lbnB > 0.f;
          const bool keep = cert && w.bposB >= 0 && w.G.d2.v.y <= a.md2 && w.G.d2.v.y * (1.f + 1e-5f) < lbnB * lbnB;
          const bool none = cert && w.slotB == -1 && lbnB * lbnB > a.md2 * (1.f + 1e-5f);
          failB = !(keep || none);
          if (none && a.c_stat) a.c_stat[iB] = SRRG2B_STAT_NONE;
          w.okB = keep;
        }
#ifdef S2B_FAIL_STATS
        if (failA | failB) {  // why: 0 no bound, 1 budget spent, 2 runner-up too close, 3 "none" too close, 4 out of range
          for (int h = 0; h < 2; ++h) {
            if (!(h ? failB : failA)) continue;
            const float lb = h ? lbB : lbA, lbn = h ? lbnB : lbnA, d2 = h ? w.G.d2.v.y : w.G.d2.v.x;
            const int slot = h ? w.slotB : w.slotA, bpos = h ? w.bposB : w.bposA;
            int why = 2;
            if (!(lb > 0.f)) why = 0; else if (!(lbn > 0.f)) why = 1; else if (slot == -1) why = 3; else if (bpos >= 0 && !(d2 <= a.md2)) why = 4;
            atomicAdd(&g_fail_stats[why], 1ull);
            if (why == 2) { atomicAdd(&g_fail_stats[5], (unsigned long long) (1e6f * (lbn - sqrtf(d2)) + 1000.f)); }
          }
        }
#endif
        const unsigned fmA = __ballot_sync(0xffffffffu, failA), fmB = __ballot_sync(0xffffffffu, failB);
        if (fmA | fmB) {
          // a few failures (converged iterations: a handful per CTA and pass): recorded -- with everything the
          // search and the lineariser need, so that nothing is fetched twice -- in the warp's list; more: the
          // global work list (searched grid-wide after the barrier)
          const int nA = __popc(fmA), n = nA + __popc(fmB);
          const unsigned lt = (1u << lane) - 1u;
          int before = 0;
          if (lane == 0) before = atomicAdd(&ctl.nfail, n);
          before = __shfl_sync(0xffffffffu, before, 0);
          if (ctl.inline_ok && before + n <= kFailCap) {
            if (failA) {
              FailRec& r = ctl.rec[before + __popc(fmA & lt)];
              r.m = make_float4(w.m.x.v.x, w.m.y.v.x, w.m.z.v.x, __int_as_float(iA));
              r.nm = make_float4(w.nm.x.v.x, w.nm.y.v.x, w.nm.z.v.x, __int_as_float(w.slotA));
              r.f = w.fA; r.nf = w.nf.a;
            }
            if (failB) {
              FailRec& r = ctl.rec[before + nA + __popc(fmB & lt)];
              r.m = make_float4(w.m.x.v.y, w.m.y.v.y, w.m.z.v.y, __int_as_float(iB));
              r.nm = make_float4(w.nm.x.v.y, w.nm.y.v.y, w.nm.z.v.y, __int_as_float(w.slotB));
              r.f = w.fB; r.nf = w.nf.b;
            }
            if (lane == 0) atomicMax(&ctl.nrec, before + n);
          } else {
            int at = 0;
            if (lane == 0) at = atomicAdd(a.work_count, n);
            at = __shfl_sync(0xffffffffu, at, 0);
            if (failA) a.work_list[at + __popc(fmA & lt)] = iA;
            if (failB) a.work_list[at + nA + __popc(fmB & lt)] = iB;
          }
        }
      } else {
        if (w.slotA == kSlotSuppressed) { A.n_ss += 1; if (a.c_stat) a.c_stat[iA] = SRRG2B_STAT_SUPPRESSED; }
        else if (!w.okA && a.c_stat && iA < a.nm) a.c_stat[iA] = SRRG2B_STAT_NONE;
        if (w.slotB == kSlotSuppressed) { A.n_ss += 1; if (a.c_stat) a.c_stat[iB] = SRRG2B_STAT_SUPPRESSED; }
        else if (!w.okB && a.c_stat && iB < a.nm) a.c_stat[iB] = SRRG2B_STAT_NONE;
      }
    }
    // ---- second half, all pairs: gate, robustifier, chi words, H / b terms ----
    bool any = false;
#pragma unroll
    for (int q = 0; q < kPPL; ++q) any = any || pw[q].okA || pw[q].okB;
    if (__any_sync(0xffffffffu, any)) {
#pragma unroll
      for (int q = 0; q < kPPL; ++q) {
        PairWork<DIM, FACTOR>& w = pw[q];
        const int iA = w.iA, iB = iA + 1;
        PairOut o;
        if (lin_pair_finish<DIM, FACTOR>(k, w.G, w.m, w.nm, w.okA, w.okB, A, o)) {
          if (w.okA) {
            if (regate && o.gateA != (w.slotA >= 0)) a.c_fpos[iA] = o.gateA ? w.bposA : -(w.bposA + 2);
            if (a.c_stat) a.c_stat[iA] = (unsigned char) o.statA;
            if (a.c_chi && o.gateA) a.c_chi[iA] = o.chiA;
          }
          if (w.okB) {
            if (regate && o.gateB != (w.slotB >= 0)) a.c_fpos[iB] = o.gateB ? w.bposB : -(w.bposB + 2);
            if (a.c_stat) a.c_stat[iB] = (unsigned char) o.statB;
            if (a.c_chi && o.gateB) a.c_chi[iB] = o.chiB;
          }
        } else {  // a non-finite chi in the pair (overflowing / NaN input): one correspondence at a time
          const float4 mA = make_float4(w.m.x.v.x, w.m.y.v.x, w.m.z.v.x, 0.f), mB = make_float4(w.m.x.v.y, w.m.y.v.y, w.m.z.v.y, 0.f);
          const float4 nmA = make_float4(w.nm.x.v.x, w.nm.y.v.x, w.nm.z.v.x, 0.f), nmB = make_float4(w.nm.x.v.y, w.nm.y.v.y, w.nm.z.v.y, 0.f);
          if (w.okA) lin_one_slot<DIM, FACTOR>(a, k, iA, w.slotA, w.bposA, mA, nmA, w.fA, w.nf.a, A);
          if (w.okB) lin_one_slot<DIM, FACTOR>(a, k, iB, w.slotB, w.bposB, mB, nmB, w.fB, w.nf.b, A);
        }
      }
    }
    // the stage is free once every lane has read it: refill it with the warp's tile kStages ahead
    if (j + kStages < my_tiles) {
      __syncwarp();
      if (lane == 0) tile_issue_bulk<CHECK>(a, WS, &bars[st], g0 + (j + kStages) * W);
    }
    if ((j & 127) == 127) {  // 32-bit partial sums: a thread stays below 512 terms per flush (huge slices only)
      lin_push_tail<DIM>(A, ctl.tail);
      A.clear();
    }
  }
  if (lane == 0) ctl.phase_bits[warp] = phase;
  return 2 * kPPL * (my_tiles & 127);
}

// the CTA's control block: mbarriers of the ring, lineariser constants of the slice about to be processed
__device__ __forceinline__ void tile_ctl_init(TileCtl& ctl) {
  if (threadIdx.x < kLoopWarps) {
    for (int s = 0; s < kStages; ++s) mbar_init(&ctl.full[threadIdx.x][s], 1);
    ctl.phase_bits[threadIdx.x] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x == 0) { ctl.nfail = 0; ctl.nrec = 0; ctl.inline_ok = 0; ctl.ep_cur = 0; }
  if (threadIdx.x < kAcc) ctl.tail[threadIdx.x] = 0;
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
  __syncthreads();
  lin_tiles_prologue<false>(a, stages, ctl);
  if (threadIdx.x == 32) make_lin_const(a, a.S, ctl.lk);
  __syncthreads();
  LinAcc<DIM> A;
  A.clear();
  const int terms = lin_tiles_body<DIM, FACTOR, false>(a, stages, ctl, A);
  lin_flush<DIM>(a.acc, terms <= 30, A, ctl.fsm);
  // (sums a warp parked in the shared tail accumulators on the way: huge slices only)
  if (threadIdx.x < kAcc && ctl.tail[threadIdx.x]) atomicAdd(&a.acc[threadIdx.x], (unsigned long long) ctl.tail[threadIdx.x]);
}

// lineariser constants of slice s into the control block (S from the device state, coherently)
__device__ __forceinline__ void load_lin_const(const SliceArgs& a, const float* Sg, int inline_ok, TileCtl& ctl) {
  const int tid = threadIdx.x;
  if (tid == 0) { ctl.nfail = 0; ctl.nrec = 0; ctl.inline_ok = inline_ok; }
  if (tid < 12) ctl.lk.S[tid] = __ldcg(Sg + tid);
  if (tid == 33 && a.S_lb) ctl.ep_cur = __ldcg(reinterpret_cast<const int*>(a.S_lb) + kSlbEpoch);
  if (tid >= 64 && tid < 64 + kEpochs && a.S_lb) {
    // displacement table of the slice's epochs at this pass's transform (epoch ep_cur = this pass)
    const int e = tid - 64, ep_cur = __ldcg(reinterpret_cast<const int*>(a.S_lb) + kSlbEpoch);
    float Sn[12];
    for (int j = 0; j < 12; ++j) Sn[j] = __ldcg(Sg + j);
    ctl.dtab[e] = e <= ep_cur ? epoch_displacement2(Sn, a.S_lb + kSlbEpS + 12 * e, a.radius, e == ep_cur) : make_float2(0.f, 3e38f);
  }
  if (tid == 32) {
    LinConst& k = ctl.lk;
    for (int i = 0; i < kKCount; ++i) k.fS[i] = a.fS[i];
    k.fSinvChi = a.fSinvChi;
    k.ip = a.ip; k.in_ = a.in_; k.rs = a.rs; k.tau = a.tau; k.delta = a.delta;
    k.normal_cos = a.normal_cos; k.eb2 = a.eb2; k.rob = a.rob; k.gate = a.gate;
  }
  __syncthreads();
}
// The failures a warp recorded during its tiles: one query at a time, the whole warp searches (the lanes
// fetch the bounds of the rows, the rows' points are dealt out to the lanes, shuffle arg-min -- nn_far_body's
// scheme), lane 0 writes slot + bound and linearises.  Everything the tile already held (query, old slot, the
// slot's fixed point and both normals) comes from the record: the dependent chain is row bounds -> points.
// The sums travel through the CTA's shared tail accumulators.  Cold path, kept out of line.
template <int DIM, bool TRACK2, int FACTOR>
__device__ __forceinline__ void fail_search_body(const SliceArgs& a, TileCtl& ctl, int n, const FailRec* recs, LinAcc<DIM>& A) {
  const int w0 = threadIdx.x >> 5, ws = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const float* S = ctl.lk.S;
  const float cell = __fdiv_rn(1.f, a.inv_cell);
  const int K = (DIM == 3) ? (2 * a.R + 1) * (2 * a.R + 1) : (2 * a.R + 1);
  const int K1 = min(K, (DIM == 3 ? 9 : 3));
  for (int w = w0; w < n; w += ws) {
    const FailRec& rec = recs[w];
    const float4 m = rec.m;
    const int i = __float_as_int(m.w), old_slot = __float_as_int(rec.nm.w);
    NNQuery q;
    nn_setup<DIM>(a, S, m, q);
    const int p0 = slot_candidate(old_slot);
    if (a.warm && p0 >= 0) { nn_consider_pt<DIM, TRACK2>(q, p0, rec.f); if (TRACK2) nn_limit_bound(q, cell); }
    for (int stage = 0; stage < 2; ++stage) {
      const int kb = stage ? K1 : 0, ke = stage ? K : K1;
      if (kb >= ke) break;
      for (int k0 = kb; k0 < ke; k0 += 32) {
        int ps = 0, cnt = 0;
        const int k = k0 + lane;
        if (k < ke) {
          const int e = ctl.rows[k];
          const int dy = (int) (signed char) (e & 0xff), dz = (int) (signed char) ((e >> 8) & 0xff);
          const int y = q.cy + dy, z = q.cz + dz;
          if (y >= 0 && y < a.ny && z >= 0 && z < a.nz) {
            const float gy = axis_gap(dy, q.fry) * cell;
            float lb2 = gy * gy;
            if (DIM == 3) {
              const float gz = axis_gap(dz, q.frz) * cell;
              lb2 = fmaf(gz, gz, lb2);
            }
            const float pr2 = TRACK2 ? q.sd2 : q.bd2;
            if (!(lb2 > pr2)) {
              const float rr = __fsqrt_rn(fmaxf(pr2 - lb2, 0.f)) * a.inv_cell_x + 2e-3f;
              const int xa = max(max((int) floorf(q.cfx - rr), q.cx - a.Rx), 0);
              const int xb = min(min((int) floorf(q.cfx + rr), q.cx + a.Rx), a.nx - 1);
              if (xa <= xb) {
                const int row = (z * a.ny + y) * a.nx;
                ps = __ldg(a.cell_start + row + xa);
                cnt = __ldg(a.cell_start + row + xb + 1) - ps;
              }
            }
          }
        }
        int incl = cnt;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, incl, off);
          if (lane >= off) incl += v;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        const int excl = incl - cnt;
        for (int t0 = 0; t0 < total; t0 += 32) {
          // candidate t lives in the last row whose exclusive prefix is <= t (empty rows are skipped)
          const int ta = t0 + lane;
          int ra = 0;
#pragma unroll
          for (int step = 16; step; step >>= 1) {
            const int ea = __shfl_sync(0xffffffffu, excl, ra + step);
            if (ea <= ta) ra += step;
          }
          const int pa = __shfl_sync(0xffffffffu, ps, ra) + (ta - __shfl_sync(0xffffffffu, excl, ra));
          if (ta < total) nn_consider_pt<DIM, TRACK2>(q, pa, __ldg(a.fp + pa));
        }
      }
#pragma unroll
      for (int off = 16; off; off >>= 1) {
        const float od2 = __shfl_xor_sync(0xffffffffu, q.bd2, off);
        const float os2 = __shfl_xor_sync(0xffffffffu, q.sd2, off);
        const int oidx = __shfl_xor_sync(0xffffffffu, q.bidx, off);
        const int opos = __shfl_xor_sync(0xffffffffu, q.bpos, off);
        const bool other_wins = od2 < q.bd2 || (od2 == q.bd2 && oidx < q.bidx);
        if (TRACK2) {
          // the loser's best is a runner-up unless both lanes hold the same point (shared warm start)
          float s2 = fminf(q.sd2, os2);
          if (q.bpos >= 0 && opos >= 0 && q.bpos != opos) s2 = fminf(s2, other_wins ? q.bd2 : od2);
          q.sd2 = s2;
        }
        if (other_wins) { q.bd2 = od2; q.bidx = oidx; q.bpos = opos; }
      }
    }
    if (lane == 0) {
      // (inside the ICP loop a pass certifies exactly at its epoch's transform: D_cert = 0)
      a.c_lb[i] = encode_bound_at(TRACK2 ? __fsqrt_rn(q.sd2) * (1.f - 1e-5f) : 0.f, 0.f, ctl.ep_cur);
      const int slot = nn_finish_keep<DIM>(a, S, q, i, old_slot);
      const int bpos = a.gate ? slot_candidate(slot) : slot;
      if (bpos >= 0) {
        const bool same = bpos == p0;  // (the usual outcome: the neighbour stands, only its bound is renewed)
        const float4 f = same ? rec.f : __ldg(a.frec + 2 * (size_t) bpos), nf = same ? rec.nf : __ldg(a.frec + 2 * (size_t) bpos + 1);
        lin_one_slot<DIM, FACTOR>(a, ctl.lk, i, slot, bpos, m, rec.nm, f, nf, A);
      } else if (a.c_stat) {
        a.c_stat[i] = SRRG2B_STAT_NONE;
      }
    }
  }
}
template <int DIM, int FACTOR>
__device__ __noinline__ void loop_search_recs(const SliceArgs& a, TileCtl& ctl, int track2, int n, const FailRec* recs) {
  LinAcc<DIM> A;
  A.clear();
  if (track2) fail_search_body<DIM, true, FACTOR>(a, ctl, n, recs, A);
  else fail_search_body<DIM, false, FACTOR>(a, ctl, n, recs, A);
  lin_push_tail_lane0<DIM>(A, ctl.tail);
}

// linearise the listed correspondences as their slots are now (thread per correspondence, scalar path)
template <int DIM, int FACTOR>
__device__ __noinline__ void loop_lin_list(const SliceArgs& a, TileCtl& ctl, int n, const int* list) {
  LinAcc<DIM> A;
  A.clear();
  const bool regate = a.gate != 0;
  int done = 0;
  for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < n; w += gridDim.x * blockDim.x) {
    const int i = __ldcg(list + w);
    const int slot = __ldcg(a.c_fpos + i);
    const int bpos = regate ? slot_candidate(slot) : (slot == kSlotSuppressed ? -1 : slot);
    if (bpos < 0) {
      if (slot == kSlotSuppressed) A.n_ss += 1;
      if (a.c_stat) a.c_stat[i] = slot == kSlotSuppressed ? SRRG2B_STAT_SUPPRESSED : SRRG2B_STAT_NONE;
      continue;
    }
    lin_one_slot<DIM, FACTOR>(a, ctl.lk, i, slot, bpos, a.mp[i], a.mn[i], __ldg(a.frec + 2 * (size_t) bpos),
                              __ldg(a.frec + 2 * (size_t) bpos + 1), A);
    if (++done >= 400) {  // 32-bit partial sums: a thread stays below 512 terms per flush
      lin_flush<DIM>(a.acc, false, A, ctl.fsm);
      A.clear();
      done = 0;
    }
  }
  lin_flush<DIM>(a.acc, false, A, ctl.fsm);
}

// ---------------------------------------------------------------------------------------------
// the kernels of one _runSolver iteration when it runs as a launch sequence (the default):
//   check_tiles_kernel   certified slice: coherence check + linearisation of every correspondence in one
//                        streaming pass; the failures are searched in place (records) or listed
//   nn_kernel / nn_far_kernel (s2b_icp.cuh)   the listed queries -- or everything while no bounds exist
//   lin_after_search_kernel   linearises what those two searched (everything / a long work list; short lists
//                        are linearised by nn_far_kernel itself)
//   icp_solve_kernel     all-reduce, solve, update, statistics, termination, next transforms
// Every kernel decides from the control words what it has to do; a launch with nothing to do costs the
// latency of those loads.
// ---------------------------------------------------------------------------------------------
template <int DIM, int FACTOR>
__global__ void __launch_bounds__(kLoopThreads, 1) check_tiles_kernel(const __grid_constant__ SliceArgs a) {
  extern __shared__ __align__(128) unsigned char loop_smem_raw[];
  TileStage* stages = reinterpret_cast<TileStage*>(loop_smem_raw);
  __shared__ TileCtl ctl;
  const int stop = __ldcg(a.stop), list_all = __ldcg(a.list_all), track2 = __ldcg(a.track2);
  if (stop || list_all) return;
  tile_ctl_init(ctl);
  constexpr int KMAX = (DIM == 3) ? kRowTable : (2 * kMaxR + 1);
  for (int k = threadIdx.x; k < KMAX; k += blockDim.x)
    ctl.rows[k] = (DIM == 3) ? *reinterpret_cast<const int*>(c_rows3[k]) : *reinterpret_cast<const int*>(c_rows2[k]);
  __syncthreads();
  lin_tiles_prologue<true>(a, stages, ctl);
  load_lin_const(a, a.S, 1, ctl);
  long long mine0, mine1;
  {
    LinAcc<DIM> A;
    A.clear();
    const int terms = lin_tiles_body<DIM, FACTOR, true>(a, stages, ctl, A);
    lin_warp_reduce<DIM>(terms <= 30, A, mine0, mine1);
  }
  __syncthreads();  // every warp is done with its tiles: the record list is complete
  if (ctl.nfail > kFailInPlace) {
    // many failures (the first check pass after a certification): the work-list kernels search them at full
    // occupancy -- several rounds of this CTA's 12 warps would only hold the pass up
    if ((int) threadIdx.x < ctl.nrec) a.work_list[atomicAdd(a.work_count, 1)] = __float_as_int(ctl.rec[threadIdx.x].m.w);
  } else if ((int) (threadIdx.x >> 5) < ctl.nrec) {
    loop_search_recs<DIM, FACTOR>(a, ctl, track2, ctl.nrec, ctl.rec);
  }
  lin_cta_reduce(a.acc, mine0, mine1, ctl.fsm);
  if (threadIdx.x < kAcc && ctl.tail[threadIdx.x]) atomicAdd(&a.acc[threadIdx.x], (unsigned long long) ctl.tail[threadIdx.x]);
}

template <int DIM, int FACTOR>
__global__ void __launch_bounds__(kLoopThreads, 1) lin_after_search_kernel(const __grid_constant__ SliceArgs a) {
  const int stop = *a.stop, list_all = *a.list_all, work_count = *a.work_count;
  if (stop) return;
  const bool all = !a.use_list || list_all;
  if (!all && (work_count == 0 || small_work_list(a, false, work_count))) return;
  extern __shared__ __align__(128) unsigned char loop_smem_raw[];
  TileStage* stages = reinterpret_cast<TileStage*>(loop_smem_raw);
  __shared__ TileCtl ctl;
  tile_ctl_init(ctl);
  __syncthreads();
  if (all) lin_tiles_prologue<false>(a, stages, ctl);
  if (threadIdx.x == 32) make_lin_const(a, a.S, ctl.lk);
  __syncthreads();
  if (all) {
    LinAcc<DIM> A;
    A.clear();
    const int terms = lin_tiles_body<DIM, FACTOR, false>(a, stages, ctl, A);
    lin_flush<DIM>(a.acc, terms <= 30, A, ctl.fsm);
    if (threadIdx.x < kAcc && ctl.tail[threadIdx.x]) atomicAdd(&a.acc[threadIdx.x], (unsigned long long) ctl.tail[threadIdx.x]);
  } else {
    loop_lin_list<DIM, FACTOR>(a, ctl, work_count, a.work_list);
  }
}

}  // namespace s2b
