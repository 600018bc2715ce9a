// s2b_icp.cuh -- device code of the aligner hot path:
//   k0  index build helpers (bounds, cell ids, Morton keys, gathers, cell table)
//   k1a nn_kernel: transform -> exact pruned grid NN (warm-started) -> normal gate
//   k1b linearize_kernel: error/Jacobian -> robust weight -> exact fixed-point accumulation of
//       H, b, chi, counters
//   k1s icp_solve_kernel: slice-ordered assembly, prior factors, 6x6/3x3 LL^T, X <- X [+] dx,
//       IterationStats append, termination criterion, next-iteration finder transforms
//   k2b export kernels (sorted order -> ascending moving_idx, optional inlier pruning)
// Everything is fp32 per term with an explicit operation order (fmaf spelled out, TU compiled with
// -fmad=false); sums are 64-bit integers, so results do not depend on grid shape or GPU count.
#pragma once
#include "s2b_lin.cuh"
#include <cooperative_groups.h>

namespace s2b {
namespace cg = cooperative_groups;

constexpr int kMaxStats = 256;
constexpr int kMaxWindow = 64;

struct Ring {
  int window, count, head, pad_;
  double v[kMaxWindow];
};

// Everything the per-iteration solve step reads and writes, contiguous and 16-byte granular so that
// the solve step can stage it in shared memory with one round of wide loads and write it back the
// same way (the serial part then never waits on global memory).  The IterationStats array follows.
struct DevHeader {
  Mat4f X;                                   // variable 0 estimate (moving in fixed)
  Mat4f S[SRRG2B_MAX_SLICES];                // robot_in_sensor * X per slice (finder transform)
  unsigned long long acc[SRRG2B_MAX_SLICES][kAcc];
  long long ncorr[SRRG2B_MAX_SLICES];
  int track2[SRRG2B_MAX_SLICES];             // NN searches of the next iteration certify bounds
  int list_all[SRRG2B_MAX_SLICES];           // next iteration has no usable bounds: search everything
  int stop;                                  // set on termination / bad association
  int n_stats;
  int not_enough_corr;
  int iterations_run;
  Ring r_ncorr, r_ninl, r_nout, r_chi;
  int tc_iterations;
  int iterations_left;                       // iterations the current run may still execute
  unsigned long long epoch;                  // peer-exchange epoch (PeerExchange), survives across runs
  srrg2b_iter_stats last_stats;              // the newest IterationStats entry (also when the array is full)
  int certified;                             // every NN point slice holds certified bounds: the device loop takes over
  int error;                                 // sticky: 1 = peer exchange timed out, 2 = grid barrier timed out
  int pad_[2];
  int ep_next[SRRG2B_MAX_SLICES];            // epoch of the slice's next pass (see the bound state below)
};
static_assert(sizeof(DevHeader) % 16 == 0, "DevHeader is copied in 16-byte pieces");

struct DevState : DevHeader {
  srrg2b_iter_stats stats[kMaxStats];
};

struct SolveSlice {
  int kind, min_corr;
  Mat4f ris, Z;
  float info[6];
  double invk[kKCount];       // 2^-k of the slice's fixed-point scales, per accumulated class
  float* S_lb;                // slice's bound state (kSlb* layout below)
  float cell, radius;         // NN cell edge / max |m| of the moving cloud (global)
  int track2_mode;            // 0 never, 1 always, 2 automatic (small motion)
  float track2_frac;          // automatic: certify once the per-iteration motion bound is below this many cells
  int* counters;              // slice's counters {far list, work list, tile ticket}: zeroed for the next iteration
  int nn_points;              // point slice searched by the grid NN finder (takes part in `certified`)
  int pad_;
};

struct alignas(16) SolveArgs {
  int dim, variable, n_slices;
  int use_tc, window, range_corr, range_inl, range_out;
  float chi_eps;
  int pad_[3];
  SolveSlice sl[SRRG2B_MAX_SLICES];
};

struct SliceArgs {
  const float4* __restrict__ mp;   // moving points, Hilbert order: x y z | local index bits
  const float4* __restrict__ mn;   // moving normals
  const float4* __restrict__ mpair;  // the same cloud, two consecutive points per 48-byte record (pair_pack_kernel)
  int nm;
  const float4* __restrict__ fp;   // fixed points, cell order: x y z | original index bits (compact: the searches scan it)
  const float4* __restrict__ frec; // the same points as 32-byte records {point | normal} (the gathers of the lineariser)
  const int* __restrict__ cell_start;
  const unsigned* __restrict__ near_bits;  // dilated occupancy (see near_bits_kernel)
  float ox, oy, oz, inv_cell;
  float inv_cell_x;  // cells are XF times finer along x (the direction the rows run): inv_cell_x = XF * inv_cell
  int Rx;            // search radius in x cells = XF * R
  int nx, ny, nz;
  int R;      // search radius in cells (cell edge = 1.05 * max_distance / R)
  int warm;   // c_fpos holds a valid candidate position per query (previous iteration's NN)
  float md2, normal_cos;
  int gate;        // apply the normal gate
  int gate_in_nn;  // 1: the NN kernels gate and write responses (stand-alone find); 0: linearise gates
  int rob;
  float tau, delta, ip, in_, rs, eb2;
  float fS[kKCount];  // 2^(k-22) per class: scale of the fixed-point conversion (to_raw)
  float fSinvChi;     // 2^(22-k) of the coarse chi word
  const float* S;
  int* c_fpos;
  int* far_list;   // phase-2 worklist of the NN search (query positions) and its counter
  int* far_count;
  int* work_list;      // queries whose coherence check failed (need a search + a second linearise pass)
  int* work_count;
  int* tile_ticket;    // next tile of the streaming lineariser (tiles are drawn dynamically by the CTAs)
  const int* list_all; // device flag: no usable bounds -> the work list is implicitly [0, nm)
  int inline_check;    // 1: nn_kernel does the coherence check itself (stand-alone finder)
  int use_list;        // 1: nn / linearise kernels iterate over the work list
  int sole_list;       // 1: nn_far_kernel is the ONLY search kernel of this iteration (see launch_slice_iteration)
  int few_terms;       // every thread of the accumulating kernel adds at most 30 terms per slot
  int small_shift;     // work lists shorter than nm >> small_shift are searched + linearised one warp per query (small_work_list)
  int nn_flat;         // 1: thread-per-query searches walk their rows in chunks of 8 with one flat candidate loop (nn_scan_rows8)
  float* c_lb;         // certified lower bound per query PLUS the motion budget at certification (0: none)
  const float* S_lb;   // bound state of the slice (see SolveSlice::S_lb)
  const int* track2;   // device flag: searches track the second neighbour (certify bounds)
  float radius;        // max |m| over the (whole, unsharded) moving cloud
  float rho_s2;        // squared radius the (2R+1) cell neighbourhood is guaranteed to cover
  float xq_slack;      // x quantum of the cell-order sort (see cell_key_kernel), with rounding slack
  // projective finder (srrg2_proslam cue): pinhole + index image of the fixed cloud
  int projective;
  float fx, fy, pcx, pcy, min_depth, max_depth;
  int width, height;
  const unsigned long long* image;  // per pixel: (depth bits << 32) | fixed index, ~0 = empty
  unsigned char* c_stat;  // may be null
  float* c_chi;           // may be null
  unsigned long long* acc;
  const int* stop;
};

// the lineariser's view of a slice (uniform per slice and iteration)
__device__ __forceinline__ void make_lin_const(const SliceArgs& a, const float* S, LinConst& k) {
#pragma unroll
  for (int i = 0; i < 12; ++i) k.S[i] = S[i];
#pragma unroll
  for (int i = 0; i < kKCount; ++i) k.fS[i] = a.fS[i];
  k.fSinvChi = a.fSinvChi;
  k.ip = a.ip; k.in_ = a.in_; k.rs = a.rs; k.tau = a.tau; k.delta = a.delta;
  k.normal_cos = a.normal_cos; k.eb2 = a.eb2;
  k.rob = a.rob; k.gate = a.gate;
}

// ---------------------------------------------------------------------------------------------
// ordered-int encoding so float min/max can use integer atomics
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int f2ord(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__host__ __device__ __forceinline__ float ord2f(int i) {
  const int j = i >= 0 ? i : i ^ 0x7fffffff;
#ifdef __CUDA_ARCH__
  return __int_as_float(j);
#else
  float f;
  memcpy(&f, &j, 4);
  return f;
#endif
}

// out[0..2] = min, out[3..5] = max (ordered ints), out[6] = max |coord| (float bits, >= 0),
// out[7] = number of valid points, out[8] = max |p|^2 (float bits; fma chain x*x, y, z as in the oracle)
constexpr int kBoundWords = 12;
__global__ void bounds_kernel(const float* __restrict__ xyz, const unsigned char* __restrict__ valid, int n,
                              int dim, int* __restrict__ out) {
  int mn[3] = {INT_MAX, INT_MAX, INT_MAX}, mx[3] = {INT_MIN, INT_MIN, INT_MIN};
  int amax = 0, cnt = 0, r2max = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (valid && !valid[i]) continue;
    ++cnt;
    float r2 = 0.f;
    for (int a = 0; a < dim; ++a) {
      const float v = xyz[(size_t) i * dim + a];
      const int o = f2ord(v);
      mn[a] = min(mn[a], o);
      mx[a] = max(mx[a], o);
      amax = max(amax, __float_as_int(fabsf(v)));
      r2 = a == 0 ? v * v : fmaf(v, v, r2);
    }
    r2max = max(r2max, __float_as_int(r2));  // (non-negative floats order like their bit patterns; NaN sorts above inf)
  }
  // warp reduce (integer min / max / add are REDUX instructions), then one row per warp in shared
  // memory, then ONE set of global atomics per CTA (they all hit the same words)
  __shared__ int red[8][9];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int a = 0; a < 3; ++a) {
    mn[a] = __reduce_min_sync(0xffffffffu, mn[a]);
    mx[a] = __reduce_max_sync(0xffffffffu, mx[a]);
  }
  amax = __reduce_max_sync(0xffffffffu, amax);
  r2max = __reduce_max_sync(0xffffffffu, r2max);
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if (lane == 0) {
    for (int a = 0; a < 3; ++a) { red[warp][a] = mn[a]; red[warp][3 + a] = mx[a]; }
    red[warp][6] = amax; red[warp][7] = cnt; red[warp][8] = r2max;
  }
  __syncthreads();
  if (threadIdx.x < 9) {
    const int k = threadIdx.x, nw = blockDim.x >> 5;
    int v = red[0][k];
    for (int w = 1; w < nw; ++w) {
      const int o = red[w][k];
      v = k < 3 ? min(v, o) : (k == 7 ? v + o : max(v, o));
    }
    if (k < 3) atomicMin(&out[k], v);
    else if (k == 7) { if (v) atomicAdd(&out[7], v); }
    else atomicMax(&out[k], v);
  }
}

// out[0] = max |n|^2 over the valid points (float bits): the fixed-point ranges of the lineariser are
// derived from it (choose_scales)
__global__ void normal_bound_kernel(const float* __restrict__ nrm, const unsigned char* __restrict__ valid, int n,
                                    int dim, int* __restrict__ out) {
  int m = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (valid && !valid[i]) continue;
    float r2 = 0.f;
    for (int a = 0; a < dim; ++a) {
      const float v = nrm[(size_t) i * dim + a];
      r2 = a == 0 ? v * v : fmaf(v, v, r2);
    }
    m = max(m, __float_as_int(r2));
  }
  m = __reduce_max_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

__device__ __forceinline__ int cell_coord(float v, float o, float inv, int n) {
  float c = (v - o) * inv;
  c = fminf(fmaxf(c, -2.f), (float) n + 1.f);
  return (int) floorf(c);
}

__device__ __forceinline__ float cell_coord_f(float v, float o, float inv, int n) {
  const float c = (v - o) * inv;
  return fminf(fmaxf(c, -2.f), (float) n + 1.f);
}

// conservative lower bound (in cells) of the distance along one axis between a query at in-cell
// fraction fr and any point of the cell at integer offset d; 2e-3 cells of slack cover the fp32
// rounding of the cell coordinate (grids are capped at 1024 cells per axis)
__device__ __forceinline__ float axis_gap(int d, float fr) {
  const float g = d > 0 ? (float) d - fr : (d < 0 ? fr - (float) (d + 1) : 0.f);
  return fmaxf(g - 2e-3f, 0.f);
}

// fixed cloud: key = (linear cell id (x fastest) << xbits) | position of x inside its cell quantised to
// xbits bits; invalid points -> all ones.  One 32-bit sort then leaves every run of consecutive cells
// of a grid row ordered by x up to one quantum (cell / 2^xbits), which lets the row scans jump to the
// query's x and stop as soon as |x - q_x| alone exceeds the pruning radius (plus that quantum).
__global__ void cell_key_kernel(const float* __restrict__ xyz, const unsigned char* __restrict__ valid, int n,
                                int dim, float ox, float oy, float oz, float inv, float inv_x, int nx, int ny, int nz,
                                int xbits, unsigned* __restrict__ keys, int* __restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  vals[i] = i;
  if (valid && !valid[i]) {
    keys[i] = 0xffffffffu;
    return;
  }
  const float x = xyz[(size_t) i * dim], y = xyz[(size_t) i * dim + 1];
  const float z = dim == 3 ? xyz[(size_t) i * dim + 2] : 0.f;
  const float cfx = cell_coord_f(x, ox, inv_x, nx);
  int cx = min(max((int) floorf(cfx), 0), nx - 1);
  int cy = min(max(cell_coord(y, oy, inv, ny), 0), ny - 1);
  int cz = min(max(cell_coord(z, oz, inv, nz), 0), nz - 1);
  const float frac = fminf(fmaxf(cfx - (float) cx, 0.f), 1.f);
  const unsigned qmax = (1u << xbits) - 1u;
  unsigned xq = min((unsigned) (frac * (float) (1u << xbits)), qmax);
  unsigned key = ((unsigned) ((cz * ny + cy) * nx + cx) << xbits) | xq;
  if (key == 0xffffffffu) key = 0xfffffffeu;  // all ones is reserved for invalid points
  keys[i] = key;
}

__device__ __forceinline__ unsigned spread3(unsigned v) {  // 10 bits -> every third bit
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x030000ffu;
  v = (v | (v << 8)) & 0x0300f00fu;
  v = (v | (v << 4)) & 0x030c30c3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}
__device__ __forceinline__ unsigned spread2(unsigned v) {  // 15 bits -> every second bit
  v &= 0x7fffu;
  v = (v | (v << 8)) & 0x00ff00ffu;
  v = (v | (v << 4)) & 0x0f0f0f0fu;
  v = (v | (v << 2)) & 0x33333333u;
  v = (v | (v << 1)) & 0x55555555u;
  return v;
}

// Hilbert index of a quantised point (Skilling's transpose algorithm, B bits per axis): unlike the
// Morton order, consecutive cells of the curve are always face neighbours, so ANY run of
// consecutive points is a spatially compact blob -- which is what lets nn_tile_kernel stage the
// whole search neighbourhood of a 256-query tile in shared memory.
template <int N, int B>
__device__ __forceinline__ void hilbert_transpose(unsigned* X) {
  const unsigned M = 1u << (B - 1);
  for (unsigned Q = M; Q > 1; Q >>= 1) {
    const unsigned P = Q - 1;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if (X[i] & Q) {
        X[0] ^= P;
      } else {
        const unsigned t = (X[0] ^ X[i]) & P;
        X[0] ^= t;
        X[i] ^= t;
      }
    }
  }
#pragma unroll
  for (int i = 1; i < N; ++i) X[i] ^= X[i - 1];
  unsigned t = 0;
  for (unsigned Q = M; Q > 1; Q >>= 1)
    if (X[N - 1] & Q) t ^= Q - 1;
#pragma unroll
  for (int i = 0; i < N; ++i) X[i] ^= t;
}

// moving cloud: Hilbert key over its own bounding box (pose independent, so the spatial coherence
// of a tile's queries survives every rigid transform the aligner applies)
__global__ void curve_key_kernel(const float* __restrict__ xyz, const unsigned char* __restrict__ valid, int n,
                                  int dim, float ox, float oy, float oz, float sx, float sy, float sz,
                                  unsigned* __restrict__ keys, int* __restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  vals[i] = i;
  if (valid && !valid[i]) {
    keys[i] = 0xffffffffu;
    return;
  }
  const float x = xyz[(size_t) i * dim], y = xyz[(size_t) i * dim + 1];
  if (dim == 3) {
    const float z = xyz[(size_t) i * dim + 2];
    unsigned X[3];
    X[0] = (unsigned) fminf(fmaxf((x - ox) * sx, 0.f), 1023.f);
    X[1] = (unsigned) fminf(fmaxf((y - oy) * sy, 0.f), 1023.f);
    X[2] = (unsigned) fminf(fmaxf((z - oz) * sz, 0.f), 1023.f);
    hilbert_transpose<3, 10>(X);
    keys[i] = (spread3(X[0]) << 2) | (spread3(X[1]) << 1) | spread3(X[2]);
  } else {
    unsigned X[2];
    X[0] = (unsigned) fminf(fmaxf((x - ox) * sx, 0.f), 32767.f);
    X[1] = (unsigned) fminf(fmaxf((y - oy) * sy, 0.f), 32767.f);
    hilbert_transpose<2, 15>(X);
    keys[i] = (spread2(X[0]) << 1) | spread2(X[1]);
  }
}

// sorted float4 SoA: points carry the original index in .w
// rec (fixed cloud): additionally the interleaved 32-byte record {point | normal} per sorted position, so that
// the lineariser's gather of a correspondence's fixed point AND normal costs exactly one DRAM sector
__global__ void gather_kernel(const float* __restrict__ xyz, const float* __restrict__ nrm,
                              const int* __restrict__ order, int n_valid, int dim, float4* __restrict__ op,
                              float4* __restrict__ on, int* __restrict__ inverse, float4* __restrict__ rec) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_valid) return;
  const int src = order[i];
  float4 p;
  p.x = xyz[(size_t) src * dim];
  p.y = xyz[(size_t) src * dim + 1];
  p.z = dim == 3 ? xyz[(size_t) src * dim + 2] : 0.f;
  p.w = __int_as_float(src);
  op[i] = p;
  float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
  if (nrm) {
    q.x = nrm[(size_t) src * dim];
    q.y = nrm[(size_t) src * dim + 1];
    q.z = dim == 3 ? nrm[(size_t) src * dim + 2] : 0.f;
  }
  if (on) on[i] = q;
  if (rec) { rec[2 * (size_t) i] = p; rec[2 * (size_t) i + 1] = q; }
  if (inverse) inverse[src] = i;
}

// Pair-interleaved copy of the sorted moving cloud for the packed (fp32x2) lineariser: record p holds points
// 2p and 2p + 1 component by component -- (x0 x1 y0 y1) (z0 z1 nx0 nx1) (ny0 ny1 nz0 nz1) -- so that one
// 16-byte shared-memory load yields two ready-made register pairs.  24 bytes per point instead of 32.
__global__ void pair_pack_kernel(const float4* __restrict__ mp, const float4* __restrict__ mn, int n,
                                 float4* __restrict__ out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (2 * p >= n) return;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 a = mp[2 * p], na = mn[2 * p];
  const float4 b = (2 * p + 1 < n) ? mp[2 * p + 1] : z4, nb = (2 * p + 1 < n) ? mn[2 * p + 1] : z4;
  out[3 * p] = make_float4(a.x, b.x, a.y, b.y);
  out[3 * p + 1] = make_float4(a.z, b.z, na.x, nb.x);
  out[3 * p + 2] = make_float4(na.y, nb.y, na.z, nb.z);
}

// cell_start[c] = first sorted position whose cell id >= c (lower bound), c in [0, ncells]:
// cell_head_kernel writes the first position of every occupied cell into a table preset to n_valid,
// a reverse running minimum (host: cub::DeviceScan over reverse iterators) fills the empty cells.
__global__ void cell_head_kernel(const unsigned* __restrict__ keys, int n_valid, int xbits, int* __restrict__ head) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_valid) return;
  const unsigned c = keys[i] >> xbits;
  if (i == 0 || (keys[i - 1] >> xbits) != c) head[c] = i;
}

// number of distinct keys among the first n sorted keys (= occupied cells)
// (cells counted at the isotropic resolution: xf consecutive x cells of a row are one cell here)
__global__ void count_distinct_kernel(const unsigned* __restrict__ keys, int n, int xbits, int nx, int xf,
                                      int* __restrict__ out) {
  int c = 0;
  auto coarse = [=](unsigned key) {
    const unsigned id = key >> xbits, row = id / (unsigned) nx, cx = id - row * (unsigned) nx;
    return (unsigned long long) row * (unsigned) nx + cx / (unsigned) xf;
  };
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    c += (i == 0 || coarse(keys[i]) != coarse(keys[i - 1])) ? 1 : 0;
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// near_bits: the bit of cell c is set iff some fixed point lies in a cell within Chebyshev distance R of
// c (the occupancy grid dilated by the search radius).  A query whose own cell has the bit clear has
// no fixed point within (R - slack) cells, i.e. nothing within max_distance: the finder can answer
// "none" from one load instead of walking (2R+1)^(DIM-1) empty rows.
// Layout: every grid row (y, z) owns nxw = ceil(nx / 32) words, bit x & 31 of word x >> 5, so that rows
// never share a word.  Pass 1 dilates along x (one thread per word), pass 2 ORs the (2R+1)^(DIM-1) rows.
__host__ __device__ __forceinline__ int near_words_per_row(int nx) { return (nx + 31) >> 5; }

__device__ __forceinline__ bool near_bit(const unsigned* __restrict__ bits, int nx, int ny, int cx, int cy, int cz) {
  const int nxw = near_words_per_row(nx);
  return (__ldg(bits + (size_t) (cz * ny + cy) * nxw + (cx >> 5)) >> (cx & 31)) & 1u;
}

__global__ void near_bits_x_kernel(const int* __restrict__ cell_start, int nx, int nrows, int R /* in x cells */,
                                   unsigned* __restrict__ out) {
  // one warp per word, one lane per cell: two coalesced loads of the row's table and a ballot
  const int nxw = near_words_per_row(nx);
  const long long t = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (t >= (long long) nrows * nxw) return;
  const int row = (int) (t / nxw), w = (int) (t - (long long) row * nxw), lane = threadIdx.x & 31;
  const int* cs = cell_start + (size_t) row * nx;
  const int x = (w << 5) + lane;
  bool near = false;
  if (x < nx) {
    const int xa = max(x - R, 0), xb = min(x + R, nx - 1);
    near = cs[xb + 1] > cs[xa];  // some point in cells [x - R, x + R] of this row
  }
  const unsigned word = __ballot_sync(0xffffffffu, near);
  if (lane == 0) out[t] = word;
}

__global__ void near_bits_yz_kernel(const unsigned* __restrict__ in, int nx, int ny, int nz, int R, int dim,
                                    unsigned* __restrict__ out) {
  const int nxw = near_words_per_row(nx);
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ny * nz * nxw) return;
  const int row = t / nxw, w = t - row * nxw;
  const int cy = row % ny, cz = row / ny;
  const int rz = dim == 3 ? R : 0;
  unsigned word = 0;
  for (int z = max(cz - rz, 0); z <= min(cz + rz, nz - 1); ++z)
    for (int y = max(cy - R, 0); y <= min(cy + R, ny - 1); ++y) word |= in[(size_t) (z * ny + y) * nxw + w];
  out[t] = word;
}

__global__ void fill_int_kernel(int* p, int n, int v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ---------------------------------------------------------------------------------------------
// k1a: exact nearest neighbour (a3).  One query per thread, queries in Morton order.
// ---------------------------------------------------------------------------------------------
// Correspondence slot encoding (c_fpos, per moving point in Morton order):
//   >= 0            accepted correspondence, value = position of the fixed point in cell order
//   -1              no fixed point within max_distance
//   <= -2           neighbour at position -(v+2) exists but the normal gate rejected it (kept only
//                   as the warm-start candidate of the next iteration)
//   kSlotSuppressed externally supplied correspondence that cannot be evaluated
constexpr int kSlotSuppressed = INT_MIN;
#ifndef S2B_ROW_WIDTH
#define S2B_ROW_WIDTH 4  // candidates per trip of a row walk (2 or 4)
#endif
constexpr int kMaxR = 4;
constexpr int kRowTable = (2 * kMaxR + 1) * (2 * kMaxR + 1);

// rows (dy, dz) of the search neighbourhood ordered by Chebyshev ring; filled by the host
__constant__ signed char c_rows3[kRowTable][4];  // dy, dz, ring, 0   (3D)
__constant__ signed char c_rows2[2 * kMaxR + 1][4];  // dy, 0, ring, 0 (2D)

// shared pieces of the two NN kernels -----------------------------------------------------------
//
// Exact temporal coherence.  Besides its slot every query keeps a certified lower bound `lb`:
//   slot holds a neighbour p0 : every OTHER fixed point is at least lb away from the query
//   slot == -1 (none)         : EVERY fixed point is at least lb away
// valid for the transform S_lb the bound was computed (or last refreshed) at.  At a new transform
// the query moved by delta = |S m - S_lb m|, so by the triangle inequality the bound lb - delta still
// holds; if d(q, p0) is strictly below it, p0 is the unique nearest neighbour and the search is
// skipped -- same result as the full search, bit for bit (1e-5 relative slack covers the fp32
// rounding of the distances).  Bounds come from searches that track the SECOND nearest point
// (track2); the solve step switches that on once the per-iteration motion is small.
struct NNQuery {
  float qx, qy, qz;      // transformed query
  float cfx;             // float cell coordinate along x
  int cx, cy, cz;        // integer cell
  float fry, frz;        // in-cell fractions along y, z
  float bd2;             // best squared distance so far (starts at max_distance^2, accept radius)
  float sd2;             // smallest squared distance of any OTHER examined point (starts at rho_s^2)
  int bidx, bpos;        // best original index / position in cell order
};

template <int DIM, bool TRACK2>
__device__ __forceinline__ void nn_consider_pt(NNQuery& q, int p, const float4 c) {
  if (TRACK2 && p == q.bpos) return;  // the warm-start candidate met again during the walk
  const float ddx = q.qx - c.x, ddy = q.qy - c.y, ddz = q.qz - c.z;
  float d2 = fmaf(ddy, ddy, ddx * ddx);
  if (DIM == 3) d2 = fmaf(ddz, ddz, d2);
  const int id = __float_as_int(c.w);
  if (d2 < q.bd2 || (d2 == q.bd2 && id < q.bidx)) {
    if (TRACK2 && q.bpos >= 0) q.sd2 = fminf(q.sd2, q.bd2);
    q.bd2 = d2; q.bidx = id; q.bpos = p;
  } else if (TRACK2) {
    q.sd2 = fminf(q.sd2, d2);
  }
}
template <int DIM, bool TRACK2>
__device__ __forceinline__ void nn_consider(const SliceArgs& a, NNQuery& q, int p) {
  nn_consider_pt<DIM, TRACK2>(q, p, __ldg(a.fp + p));
}

// A search that certifies a bound examines everything within sqrt(sd2).  With a warm-start neighbour
// at distance d0 there is no point in certifying more than max(2 d0, cell / 4): later iterations only
// need the bound to exceed d0 by the (sub-millimetre) motion of a converged estimate, and the smaller
// start value prunes the rows of ring 1 like a nearest-only search.  (A smaller sd2 only makes the
// certified bound smaller, never wrong.)
__device__ __forceinline__ void nn_limit_bound(NNQuery& q, float cell) {
  if (q.bpos >= 0) q.sd2 = fminf(q.sd2, fmaxf(4.f * q.bd2, 0.0625f * cell * cell));
}

// scan the part of cell row (y, z) that can still matter, given the conservative squared distance
// lb2 between the query and the row's y/z slab.  Pruning radius: bd2 (nearest only) or sd2 (two
// nearest, needed to certify a bound).
template <int DIM, bool TRACK2>
__device__ __forceinline__ void nn_scan_row(const SliceArgs& a, NNQuery& q, int y, int z, float lb2) {
  const float pr2 = TRACK2 ? q.sd2 : q.bd2;
  const float rr = __fsqrt_rn(fmaxf(pr2 - lb2, 0.f)) * a.inv_cell_x + 2e-3f;
  const int xa = max(max((int) floorf(q.cfx - rr), q.cx - a.Rx), 0);
  const int xb = min(min((int) floorf(q.cfx + rr), q.cx + a.Rx), a.nx - 1);
  if (xa > xb) return;
  const int row = (z * a.ny + y) * a.nx;
  const int ps = __ldg(a.cell_start + row + xa);
  const int pe = __ldg(a.cell_start + row + xb + 1);
  // candidates in ascending position, four loads in flight per step (the compare chain is serial,
  // the loads are not): out-of-range slots re-read the last point and are not considered.
  // (Measured: using the x order of the run here -- a 4-ary search for the query's x, then a windowed
  // walk -- examines 4x fewer candidates but is 20-35% SLOWER: the search adds dependent L2 round trips
  // to a walk whose loads are otherwise all independent.  The x order pays off in shared memory only.)
  const int last = pe - 1;
#if S2B_ROW_WIDTH == 2
  // (experiment, -DS2B_ROW_WIDTH=2: a CPU simulation of the walk -- tools/cell_size_study.py -- finds ~1.3 points per
  // visited run at C2, so a four-wide trip spends most of its slots on padding; measured, two per trip is SLOWER all
  // the same -- iterations 1 / 2 / 3: 277 / 188 / 174 us against 241 / 166 / 172 us -- the loads in flight are worth
  // more than the padded slots cost)
#pragma unroll 1
  for (int p = ps; p < pe; p += 2) {
    const float4 c0 = __ldg(a.fp + p);
    const float4 c1 = __ldg(a.fp + min(p + 1, last));
    nn_consider_pt<DIM, TRACK2>(q, p, c0);
    if (p + 1 < pe) nn_consider_pt<DIM, TRACK2>(q, p + 1, c1);
  }
#else
#pragma unroll 1
  for (int p = ps; p < pe; p += 4) {
    const float4 c0 = __ldg(a.fp + p);
    const float4 c1 = __ldg(a.fp + min(p + 1, last));
    const float4 c2 = __ldg(a.fp + min(p + 2, last));
    const float4 c3 = __ldg(a.fp + min(p + 3, last));
    nn_consider_pt<DIM, TRACK2>(q, p, c0);
    if (p + 1 < pe) nn_consider_pt<DIM, TRACK2>(q, p + 1, c1);
    if (p + 2 < pe) nn_consider_pt<DIM, TRACK2>(q, p + 2, c2);
    if (p + 3 < pe) nn_consider_pt<DIM, TRACK2>(q, p + 3, c3);
  }
#endif
}

// Up to 8 rows of the search neighbourhood in one go, SIMT-friendly.  rows[k0..k1) are packed table entries
// (dy | dz << 8 | ring << 16).  Stage A: every lane prunes its 8 rows arithmetically and fetches the bounds of the
// survivors together -- 16 independent loads in flight instead of 8 dependent pairs one after the other; runs
// that hold points are parked in the lane's column of shared memory (park[0..8) first positions, [8..16) ends,
// [16..24) slab distances; stride = blockDim.x).  Stage B: ONE loop over the parked runs, four candidates per
// trip, so the lanes of a warp diverge by their total candidate count only and no longer per row (the nested
// row / candidate loops ran at 12 of 32 lanes).  A row whose slab distance the shrinking radius has overtaken is
// dropped when its turn comes.  The x range of a row is fixed by the radius at stage A: a few more
// candidates than the one-row-at-a-time walk, same result -- (nearest, second nearest) do not depend on the order or
// on how much beyond the final radius was examined.
constexpr int kRowChunk = 8;
template <int DIM, bool TRACK2>
__device__ __forceinline__ void nn_scan_rows8(const SliceArgs& a, NNQuery& q, const int* rows, int k0, int k1, float cell,
                                              int* park, int stride) {
  const unsigned am = __activemask();
  const float pr2 = TRACK2 ? q.sd2 : q.bd2;
  int ps[kRowChunk], pe[kRowChunk];
  float lbs[kRowChunk];
#pragma unroll
  for (int u = 0; u < kRowChunk; ++u) {
    ps[u] = 0; pe[u] = 0; lbs[u] = 0.f;
    const int k = k0 + u;
    if (k < k1) {  // (uniform)
      const int e = rows[k];
      const int dy = (int) (signed char) (e & 0xff), dz = (int) (signed char) ((e >> 8) & 0xff);
      const int y = q.cy + dy, z = q.cz + dz;
      const float gy = axis_gap(dy, q.fry) * cell;
      float lb2 = gy * gy;
      if (DIM == 3) {
        const float gz = axis_gap(dz, q.frz) * cell;
        lb2 = fmaf(gz, gz, lb2);
      }
      bool ok = y >= 0 && y < a.ny && z >= 0 && z < a.nz && !(lb2 > pr2);
      if (__any_sync(am, ok)) {
        const float rr = __fsqrt_rn(fmaxf(pr2 - lb2, 0.f)) * a.inv_cell_x + 2e-3f;
        const int xa = max(max((int) floorf(q.cfx - rr), q.cx - a.Rx), 0);
        const int xb = min(min((int) floorf(q.cfx + rr), q.cx + a.Rx), a.nx - 1);
        ok = ok && xa <= xb;
        const int row = (z * a.ny + y) * a.nx;
        // (a lane without this row reads cell_start[0] twice: an empty run)
        ps[u] = __ldg(a.cell_start + (ok ? row + xa : 0));
        pe[u] = __ldg(a.cell_start + (ok ? row + xb + 1 : 0));
        lbs[u] = lb2;
      }
    }
  }
  int nr = 0;
#pragma unroll
  for (int u = 0; u < kRowChunk; ++u) {
    if (pe[u] > ps[u]) {
      park[nr * stride] = ps[u];
      park[(kRowChunk + nr) * stride] = pe[u];
      park[(2 * kRowChunk + nr) * stride] = __float_as_int(lbs[u]);
      ++nr;
    }
  }
  int j = 0, p = 0, end = 0;
#pragma unroll 1
  for (;;) {
    if (p >= end) {
      if (j >= nr) break;
      p = park[j * stride];
      end = park[(kRowChunk + j) * stride];
      const float lb2 = __int_as_float(park[(2 * kRowChunk + j) * stride]);
      ++j;
      if (lb2 > (TRACK2 ? q.sd2 : q.bd2)) end = p;
    }
    if (p < end) {
      const int last = end - 1;
#if S2B_ROW_WIDTH == 2
      const float4 c0 = __ldg(a.fp + p);
      const float4 c1 = __ldg(a.fp + min(p + 1, last));
      nn_consider_pt<DIM, TRACK2>(q, p, c0);
      if (p + 1 < end) nn_consider_pt<DIM, TRACK2>(q, p + 1, c1);
      p += 2;
#else
      const float4 c0 = __ldg(a.fp + p);
      const float4 c1 = __ldg(a.fp + min(p + 1, last));
      const float4 c2 = __ldg(a.fp + min(p + 2, last));
      const float4 c3 = __ldg(a.fp + min(p + 3, last));
      nn_consider_pt<DIM, TRACK2>(q, p, c0);
      if (p + 1 < end) nn_consider_pt<DIM, TRACK2>(q, p + 1, c1);
      if (p + 2 < end) nn_consider_pt<DIM, TRACK2>(q, p + 2, c2);
      if (p + 3 < end) nn_consider_pt<DIM, TRACK2>(q, p + 3, c3);
      p += 4;
#endif
    }
  }
}

// rings >= 1 of a thread-per-query search in chunks of 8 rows, nearest ring first (rows[] is ordered by ring)
template <int DIM, bool TRACK2>
__device__ __forceinline__ void nn_scan_rings_flat(const SliceArgs& a, NNQuery& q, const int* rows, int k_first, int K,
                                                   float cell, int* park, int stride) {
  for (int k = k_first; k < K; k += kRowChunk) {
    const int ring = (rows[k] >> 16) & 0xff;
    if (ring >= 2) {  // every row of this and later rings is at least (ring - 1) cells away
      const float g = ((float) (ring - 1) - 4e-3f) * cell;
      if (g * g > (TRACK2 ? q.sd2 : q.bd2)) break;
    }
    nn_scan_rows8<DIM, TRACK2>(a, q, rows, k, min(k + kRowChunk, K), cell, park, stride);
  }
}

template <int DIM>
__device__ __forceinline__ void nn_transform(const float* S, const float4 m, float& x, float& y, float& z) {
  float t;  // q = S m  (operation order is part of the numerics contract)
  t = S[0] * m.x; t = fmaf(S[1], m.y, t); if (DIM == 3) t = fmaf(S[2], m.z, t); x = t + S[3];
  t = S[4] * m.x; t = fmaf(S[5], m.y, t); if (DIM == 3) t = fmaf(S[6], m.z, t); y = t + S[7];
  z = 0.f;
  if (DIM == 3) { t = S[8] * m.x; t = fmaf(S[9], m.y, t); t = fmaf(S[10], m.z, t); z = t + S[11]; }
}

template <int DIM>
__device__ __forceinline__ void nn_setup(const SliceArgs& a, const float* S, const float4 m, NNQuery& q) {
  nn_transform<DIM>(S, m, q.qx, q.qy, q.qz);
  q.cfx = cell_coord_f(q.qx, a.ox, a.inv_cell_x, a.nx);
  const float cfy = cell_coord_f(q.qy, a.oy, a.inv_cell, a.ny);
  const float cfz = (DIM == 3) ? cell_coord_f(q.qz, a.oz, a.inv_cell, a.nz) : 0.f;
  q.cx = (int) floorf(q.cfx); q.cy = (int) floorf(cfy); q.cz = (DIM == 3) ? (int) floorf(cfz) : 0;
  q.fry = cfy - (float) q.cy; q.frz = cfz - (float) q.cz;
  q.bd2 = a.md2; q.sd2 = a.rho_s2; q.bidx = INT_MAX; q.bpos = -1;
}

// work lists shorter than this are handled by one warp per query inside nn_far_kernel (search +
// linearise), where the chain of dependent loads of a search is short; the thread-per-query kernels
// then return at once
__device__ __forceinline__ bool small_work_list(const SliceArgs& a, bool all, int n_work) {
  return a.use_list && !all && n_work < max(a.nm >> a.small_shift, 64);
}

__device__ __forceinline__ int slot_candidate(int slot) {
  const int c = slot >= 0 ? slot : -(slot + 2);  // -1 -> -1; kSlotSuppressed would wrap to a positive value
  return slot == kSlotSuppressed ? -1 : c;
}

// Bound state of a slice (float words of SolveSlice::S_lb / SliceArgs::S_lb):
//   [0..15]  transform of the last pass (export recomputes the responses with it)
//   [16]     D_cert: how far any query is from its position at the CURRENT EPOCH's transform while the running
//            pass certifies bounds (0 inside the ICP loop: a pass certifies exactly at its epoch's transform)
//   [18]     id of the current epoch (integer bits)
//   [32..95] dtab[e]: what the coherence check subtracts from a bound of epoch e = a bound of any query's
//            displacement between the current transform and epoch e's (+ absolute rounding slack); 3e38 = dead
//   [96.. ]  transforms of the epochs (12 floats each)
// Every pass of the ICP loop is an epoch of its own, so a bound is always measured against the transform it
// was certified at -- a converged estimate spends nothing of it, however long the run.  A certified bound
// lb is stored as (lb - D_cert), rounded down, with the epoch id in its 6 low mantissa bits; 0 = no bound.
constexpr int kEpochs = 64, kSlbDcert = 16, kSlbEpoch = 18, kSlbDtab = 32, kSlbEpS = 96, kSlbFloats = 96 + 12 * kEpochs;
__device__ __forceinline__ float encode_bound_at(float lb, float d_cert, int ep) {
  const float v = (lb - d_cert) * (1.f - 2.4e-7f) - 1e-9f;
  if (!(lb > 0.f) || !(v > 1e-30f)) return 0.f;
  return __int_as_float((__float_as_int(v) & ~(kEpochs - 1)) | (ep & (kEpochs - 1)));
}
__device__ __forceinline__ float encode_bound(float lb, const float* S_lb) {
  const float d_cert = *reinterpret_cast<const volatile float*>(S_lb + kSlbDcert);
  const int ep = *reinterpret_cast<const volatile int*>(S_lb + kSlbEpoch);
  return encode_bound_at(lb, d_cert, ep);
}
// bound minus everything the query can have moved since certification (<= 0: nothing left / no bound)
__device__ __forceinline__ float decode_bound(float stored, const float* dtab) {
  if (!(stored > 0.f)) return 0.f;
  const int b = __float_as_int(stored);
  return __int_as_float(b & ~(kEpochs - 1)) - dtab[b & (kEpochs - 1)];
}

// slot / bound of a finished query.  Inside the ICP loop the normal gate is evaluated by the
// lineariser (it has both normals in registers anyway); the stand-alone finder gates here.
template <int DIM>
__device__ __forceinline__ int nn_finish_keep(const SliceArgs& a, const float* S, const NNQuery& q, int i, int old_slot) {
  int slot = q.bpos;
  if (q.bpos >= 0 && a.gate && a.gate_in_nn) {
    const float4 nm = a.mn[i];
    const float4 nf = __ldg(a.frec + 2 * (size_t) q.bpos + 1);
    float t;
    t = S[0] * nm.x; t = fmaf(S[1], nm.y, t); if (DIM == 3) t = fmaf(S[2], nm.z, t); const float nqx = t;
    t = S[4] * nm.x; t = fmaf(S[5], nm.y, t); if (DIM == 3) t = fmaf(S[6], nm.z, t); const float nqy = t;
    float dot = fmaf(nf.y, nqy, nf.x * nqx);
    if (DIM == 3) {
      t = S[8] * nm.x; t = fmaf(S[9], nm.y, t); t = fmaf(S[10], nm.z, t);
      dot = fmaf(nf.z, t, dot);
    }
    if (dot < a.normal_cos) slot = -(q.bpos + 2);
  } else if (q.bpos >= 0 && a.gate && old_slot == -(q.bpos + 2)) {
    slot = old_slot;  // same neighbour as before and it was gated out: the lineariser re-checks it
  }
  if (slot != old_slot) a.c_fpos[i] = slot;
  return slot;
}
template <int DIM>
__device__ __forceinline__ int nn_finish(const SliceArgs& a, const float* S, const NNQuery& q, int i, float lb,
                                         int old_slot) {
  a.c_lb[i] = encode_bound(lb, a.S_lb);
  return nn_finish_keep<DIM>(a, S, q, i, old_slot);
}

// Phase 1: temporal-coherence check, else warm start + the 3^(DIM-1) rows of rings 0 and 1 (each
// lane walks only its own surviving rows).  A query whose best distance is still larger than the
// distance to ring 2 is handed to phase 2 through a worklist, so that the rare expensive queries
// (outliers, large initial misalignment) do not serialise the warps of the cheap ones.
template <int DIM, bool TRACK2, bool FLAT>
__device__ __forceinline__ void nn_phase1_body(const SliceArgs& a, const float* S, float cell,
                                               float ring2, float ring2_sq, bool all, int n_work, const int* rows, int* park) {
  for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < n_work; w += gridDim.x * blockDim.x) {
    const int i = all ? w : __ldcg(a.work_list + w);
    NNQuery q;
    const float4 m = a.mp[i];
    nn_setup<DIM>(a, S, m, q);
    const int old_slot = __ldcg(a.c_fpos + i);
    const int p0 = slot_candidate(old_slot);
    const float lb_old = a.inline_check ? __ldcg(a.c_lb + i) : 0.f;
    if (lb_old > 0.f) {
      // exact temporal coherence: the bound minus everything the query can have moved since certification
      const float lbn = decode_bound(lb_old, a.S_lb + kSlbDtab);
      if (lbn > 0.f) {
        if (p0 >= 0) {
          nn_consider<DIM, false>(a, q, p0);
          if (q.bpos >= 0 && q.bd2 * (1.f + 1e-5f) < lbn * lbn) {  // p0 is still the unique neighbour
            nn_finish_keep<DIM>(a, S, q, i, old_slot);
            continue;
          }
          q.bd2 = a.md2; q.bidx = INT_MAX; q.bpos = -1;
        } else if (old_slot == -1 && lbn * lbn > a.md2 * (1.f + 1e-5f)) {  // still nothing in range
          continue;
        }
      }
    }
    if (p0 < 0 && q.cx >= 0 && q.cx < a.nx && q.cy >= 0 && q.cy < a.ny && q.cz >= 0 && q.cz < a.nz) {
      // nothing occupied within R cells of the query's cell: no fixed point within the covered radius
      if (!near_bit(a.near_bits, a.nx, a.ny, q.cx, q.cy, q.cz)) {
        nn_finish<DIM>(a, S, q, i, TRACK2 ? __fsqrt_rn(a.rho_s2) * (1.f - 1e-5f) : 0.f, old_slot);
        continue;
      }
    }
    if (a.warm && p0 >= 0) {  // a real candidate: exactness untouched
      nn_consider<DIM, TRACK2>(a, q, p0);
      if (TRACK2) nn_limit_bound(q, cell);
    }
    // squared slab gaps for offsets -1 and +1 along y and z (offset 0 has gap 0)
    const float gym = fmaxf(q.fry - 2e-3f, 0.f) * cell, gyp = fmaxf(1.f - q.fry - 2e-3f, 0.f) * cell;
    const float gy2m = gym * gym, gy2p = gyp * gyp;
    float gz2m = 0.f, gz2p = 0.f;
    if (DIM == 3) {
      const float gzm = fmaxf(q.frz - 2e-3f, 0.f) * cell, gzp = fmaxf(1.f - q.frz - 2e-3f, 0.f) * cell;
      gz2m = gzm * gzm; gz2p = gzp * gzp;
    }
    // centre row first: it usually tightens the pruning radius enough to drop most of ring 1
    if (q.cy >= 0 && q.cy < a.ny && q.cz >= 0 && q.cz < a.nz) nn_scan_row<DIM, TRACK2>(a, q, q.cy, q.cz, 0.f);
    // ring 1: bit b = jz * 3 + jy marks a row this lane still has to visit
    unsigned mask = 0;
    // Ring 1 as one flat chunk pays when the centre row leaves much to do -- searches that certify bounds, and warps
    // whose queries start without a warm-start candidate (the first iteration of a run: 248 -> 230 us at C2); with a
    // good candidate most lanes drop all of ring 1 and the per-lane walk below is cheaper (measured: 164 vs 181 us).
    bool flat = FLAT;
    if (FLAT && !TRACK2) {
      const unsigned am = __activemask();
      flat = 2 * __popc(__ballot_sync(am, p0 < 0)) > __popc(am);
    }
    if (flat) {  // the 3^(DIM-1) - 1 rows of ring 1 as one chunk
      nn_scan_rows8<DIM, TRACK2>(a, q, rows, 1, DIM == 3 ? 9 : 3, cell, park, blockDim.x);
    } else {
      const float pr2 = TRACK2 ? q.sd2 : q.bd2;
#pragma unroll
      for (int jz = (DIM == 3 ? 0 : 1); jz < (DIM == 3 ? 3 : 2); ++jz) {
        const int z = q.cz + jz - 1;
        const bool zin = (z >= 0 && z < a.nz);
#pragma unroll
        for (int jy = 0; jy < 3; ++jy) {
          if (jy == 1 && jz == 1) continue;
          const int y = q.cy + jy - 1;
          const float lb2 = (jy == 0 ? gy2m : (jy == 2 ? gy2p : 0.f)) + (jz == 0 ? gz2m : (jz == 2 ? gz2p : 0.f));
          if (zin && y >= 0 && y < a.ny && !(lb2 > pr2)) mask |= 1u << (jz * 3 + jy);
        }
      }
    }
    while (mask) {
      const int b = __ffs(mask) - 1;
      mask &= mask - 1;
      const int jz = (b * 11) >> 5, jy = b - 3 * jz;
      const float lb2 = (jy == 0 ? gy2m : (jy == 2 ? gy2p : 0.f)) + (jz == 0 ? gz2m : (jz == 2 ? gz2p : 0.f));
      if (lb2 > (TRACK2 ? q.sd2 : q.bd2)) continue;
      nn_scan_row<DIM, TRACK2>(a, q, q.cy + jy - 1, q.cz + jz - 1, lb2);
    }
    if (q.bd2 > ring2_sq) {  // not settled by rings 0-1: phase 2 continues this query from ring 2
      // rings 0-1 are done for good: their best point (if any) travels in the slot as the warm-start
      // candidate, and phase 2 -- when it only needs the nearest point -- starts at ring 2
      if (!TRACK2 && q.bpos >= 0 && q.bpos != old_slot) a.c_fpos[i] = q.bpos;
      const int w = atomicAdd(a.far_count, 1);
      a.far_list[w] = i;
      continue;
    }
    // everything within min(sqrt(sd2), ring2) of the query has been examined
    const float lb = TRACK2 ? fminf(__fsqrt_rn(q.sd2), ring2) * (1.f - 1e-5f) : 0.f;
    nn_finish<DIM>(a, S, q, i, lb, old_slot);
  }
}

template <int DIM>
__global__ void __launch_bounds__(256) nn_kernel(const SliceArgs a, const int* skip) {
  // the control words are fetched together: when this launch has nothing to do its cost is the latency
  // of these loads
  const int stop = *a.stop, list_all = *a.list_all, work_count = *a.work_count, track2 = *a.track2;
  const int skipped = skip ? *skip : 0;
  if (stop || skipped) return;
  const bool all = !a.use_list || list_all;
  const int n_work = all ? a.nm : work_count;
  if (small_work_list(a, all, n_work)) return;
  __shared__ float S[16];
  __shared__ int rows[16];
  __shared__ int park_all[DIM == 3 ? 3 * kRowChunk * 256 : 1];  // (2D: 2 rows per ring -- the flat chunks are compiled out)
  if (threadIdx.x < 16) S[threadIdx.x] = a.S[threadIdx.x];
  if (threadIdx.x >= 32 && threadIdx.x < 32 + (DIM == 3 ? 9 : 3))
    rows[threadIdx.x - 32] = (DIM == 3) ? *reinterpret_cast<const int*>(c_rows3[threadIdx.x - 32])
                                        : *reinterpret_cast<const int*>(c_rows2[threadIdx.x - 32]);
  int* park = park_all + threadIdx.x;
  __syncthreads();
  const float cell = __fdiv_rn(1.f, a.inv_cell);
  // distance below which a point cannot lie in ring 2 or beyond (for R == 1: the covered radius)
  const float ring2 = (a.R >= 2) ? (1.f - 4e-3f) * cell : __fsqrt_rn(a.rho_s2);
  const float ring2_sq = (a.R >= 2) ? ring2 * ring2 : 3.0e38f;
  if (DIM == 3 && (a.nn_flat & 1)) {
    if (track2) nn_phase1_body<DIM, true, true>(a, S, cell, ring2, ring2_sq, all, n_work, rows, park);
    else nn_phase1_body<DIM, false, true>(a, S, cell, ring2, ring2_sq, all, n_work, rows, park);
  } else {
    if (track2) nn_phase1_body<DIM, true, false>(a, S, cell, ring2, ring2_sq, all, n_work, rows, park);
    else nn_phase1_body<DIM, false, false>(a, S, cell, ring2, ring2_sq, all, n_work, rows, park);
  }
}

// Phase 2: the queries phase 1 could not settle (worklist).  These are few but expensive
// (typically no neighbour at all, so nothing prunes), so one WARP takes one query: the lanes fetch
// the bounds of the (2R+1)^(DIM-1) rows, the points of all rows are dealt out to the lanes, then a
// shuffle reduction merges the lanes' (nearest, second nearest) pairs and lane 0 applies the gate
// and writes slot and bound.
template <int DIM, int FACTOR>
__device__ __forceinline__ void lin_one_slot(const SliceArgs& a, const LinConst& k, int i, int slot, int bpos,
                                             const float4 m, const float4 nm, const float4 f, const float4 nf,
                                             LinAcc<DIM>& A);

template <int DIM, bool TRACK2, int FACTOR = SRRG2B_FACTOR_P2P>
__device__ __forceinline__ void nn_far_body(const SliceArgs& a, const float* S, const int* rows, int K, float cell,
                                            int n_far, const int* list, LinAcc<DIM>* lin, const LinConst* lk, int w_first,
                                            int w_stride, int* park) {
  const int lane = threadIdx.x & 31;
  if (!lin && n_far > (a.nm >> 4)) {
    // long worklist (large initial misalignment): one THREAD per query, rows nearest ring first
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < n_far; w += gridDim.x * blockDim.x) {
      const int i = *reinterpret_cast<const volatile int*>(list + w);
      NNQuery q;
      nn_setup<DIM>(a, S, a.mp[i], q);
      const int old_slot = __ldcg(a.c_fpos + i);
      const int p0 = slot_candidate(old_slot);
      if (a.warm && p0 >= 0) { nn_consider<DIM, TRACK2>(a, q, p0); if (TRACK2) nn_limit_bound(q, cell); }
      // (far list of phase 1, nearest point only: rings 0-1 were searched exhaustively there)
      if (DIM == 3 && (a.nn_flat & 2)) {
        int k = TRACK2 ? 0 : min(K, (DIM == 3 ? 9 : 3));
        if (k == 0) {
          if (q.cy >= 0 && q.cy < a.ny && q.cz >= 0 && q.cz < a.nz) nn_scan_row<DIM, TRACK2>(a, q, q.cy, q.cz, 0.f);
          k = 1;
        }
        nn_scan_rings_flat<DIM, TRACK2>(a, q, rows, k, K, cell, park, blockDim.x);
      } else
      for (int k = TRACK2 ? 0 : min(K, (DIM == 3 ? 9 : 3)); k < K; ++k) {
        const int e = rows[k];
        const int dy = (int) (signed char) (e & 0xff), dz = (int) (signed char) ((e >> 8) & 0xff);
        const int ring = (e >> 16) & 0xff;
        const float pr2 = TRACK2 ? q.sd2 : q.bd2;
        if (ring >= 2) {  // every row of this and later rings is at least (ring - 1) cells away
          const float g = ((float) (ring - 1) - 4e-3f) * cell;
          if (g * g > pr2) break;
        }
        const int y = q.cy + dy, z = q.cz + dz;
        if (y < 0 || y >= a.ny || z < 0 || z >= a.nz) continue;
        const float gy = axis_gap(dy, q.fry) * cell;
        float lb2 = gy * gy;
        if (DIM == 3) {
          const float gz = axis_gap(dz, q.frz) * cell;
          lb2 = fmaf(gz, gz, lb2);
        }
        if (lb2 > pr2) continue;
        nn_scan_row<DIM, TRACK2>(a, q, y, z, lb2);
      }
      nn_finish<DIM>(a, S, q, i, TRACK2 ? __fsqrt_rn(q.sd2) * (1.f - 1e-5f) : 0.f, old_slot);
    }
    return;
  }
  for (int w = w_first; w < n_far; w += w_stride) {
    const int i = *reinterpret_cast<const volatile int*>(list + w);
    NNQuery q;
    const float4 m = a.mp[i];
    nn_setup<DIM>(a, S, m, q);
    const int old_slot = __ldcg(a.c_fpos + i);
    const int p0 = slot_candidate(old_slot);
    if (a.warm && p0 >= 0) { nn_consider<DIM, TRACK2>(a, q, p0); if (TRACK2) nn_limit_bound(q, cell); }
    // the warp takes the rows 32 at a time: lane -> bounds of one row, then the points of all 32 rows
    // are dealt out to the lanes round robin (a handful of dependent loads per query instead of a
    // serial walk per row); every lane keeps its own (nearest, second nearest) pair
    // two stages: rings 0-1 first, merged across the warp, so that the outer rings are pruned with the
    // radius the near rows established (nearly always to nothing)
    const int K1 = min(K, (DIM == 3 ? 9 : 3));
    // far list of phase 1 and only the nearest point wanted: rings 0-1 were searched exhaustively there
    for (int stage = (!lin && !TRACK2) ? 1 : 0; stage < 2; ++stage) {
      const int kb = stage ? K1 : 0, ke = stage ? K : K1;
      if (kb >= ke) break;
      for (int k0 = kb; k0 < ke; k0 += 32) {
        int ps = 0, cnt = 0;
        const int k = k0 + lane;
        if (k < ke) {
          const int e = rows[k];
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
        for (int t0 = 0; t0 < total; t0 += 64) {
          // candidate t lives in the last row whose exclusive prefix is <= t (empty rows are skipped)
          const int ta = t0 + lane, tb = t0 + 32 + lane;
          int ra = 0, rb = 0;
  #pragma unroll
          for (int step = 16; step; step >>= 1) {
            const int ea = __shfl_sync(0xffffffffu, excl, ra + step);
            const int eb = __shfl_sync(0xffffffffu, excl, rb + step);
            if (ea <= ta) ra += step;
            if (eb <= tb) rb += step;
          }
          const int pa = __shfl_sync(0xffffffffu, ps, ra) + (ta - __shfl_sync(0xffffffffu, excl, ra));
          const int pb = __shfl_sync(0xffffffffu, ps, rb) + (tb - __shfl_sync(0xffffffffu, excl, rb));
          float4 ca, cb;
          if (ta < total) ca = __ldg(a.fp + pa);
          if (tb < total) cb = __ldg(a.fp + pb);
          if (ta < total) nn_consider_pt<DIM, TRACK2>(q, pa, ca);
          if (tb < total) nn_consider_pt<DIM, TRACK2>(q, pb, cb);
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
          float s = fminf(q.sd2, os2);
          if (q.bpos >= 0 && opos >= 0 && q.bpos != opos) s = fminf(s, other_wins ? q.bd2 : od2);
          q.sd2 = s;
        }
        if (other_wins) { q.bd2 = od2; q.bidx = oidx; q.bpos = opos; }
      }
    }
    if (lane == 0) {
      const int slot = nn_finish<DIM>(a, S, q, i, TRACK2 ? __fsqrt_rn(q.sd2) * (1.f - 1e-5f) : 0.f, old_slot);
      if (lin) {  // tail mode: linearise the query right away (the gate is re-evaluated there)
        const int bpos = a.gate ? slot_candidate(slot) : slot;
        if (bpos >= 0) lin_one_slot<DIM, FACTOR>(a, *lk, i, slot, bpos, m, a.mn[i], __ldg(a.frec + 2 * (size_t) bpos), __ldg(a.frec + 2 * (size_t) bpos + 1), *lin);
        else if (a.c_stat) a.c_stat[i] = SRRG2B_STAT_NONE;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// k1p: projective association (a3 for the RGB-D cue).  The fixed cloud (sensor/camera frame) is
// rendered into an index image: per pixel the point of smallest depth wins, lowest index on ties.
// A moving point is transformed, projected, and associated with the pixel's point if it lies
// within max_distance (then the usual normal gate).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool project_pixel(float x, float y, float z, float fx, float fy, float cx, float cy,
                                              float min_depth, float max_depth, int width, int height, int& pix) {
  if (!(z > min_depth) || !(z < max_depth)) return false;
  const float u = fmaf(fx, __fdiv_rn(x, z), cx);
  const float v = fmaf(fy, __fdiv_rn(y, z), cy);
  const float uf = floorf(u + 0.5f), vf = floorf(v + 0.5f);
  if (!(uf >= 0.f) || !(vf >= 0.f) || !(uf < (float) width) || !(vf < (float) height)) return false;
  pix = (int) vf * width + (int) uf;
  return true;
}

__global__ void proj_image_kernel(const float* __restrict__ xyz, const unsigned char* __restrict__ valid, int n,
                                  float fx, float fy, float cx, float cy, float min_depth, float max_depth, int width,
                                  int height, unsigned long long* __restrict__ image) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || (valid && !valid[i])) return;
  const float x = xyz[(size_t) i * 3], y = xyz[(size_t) i * 3 + 1], z = xyz[(size_t) i * 3 + 2];
  int pix;
  if (!project_pixel(x, y, z, fx, fy, cx, cy, min_depth, max_depth, width, height, pix)) return;
  const unsigned long long key = ((unsigned long long) __float_as_uint(z) << 32) | (unsigned) i;
  atomicMin(&image[pix], key);
}

// float4 SoA in ORIGINAL order (position == index) for the projective finder
__global__ void gather_identity_kernel(const float* __restrict__ xyz, const float* __restrict__ nrm, int n, int dim,
                                       float4* __restrict__ op, float4* __restrict__ rec, int* __restrict__ inverse) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p;
  p.x = xyz[(size_t) i * dim];
  p.y = xyz[(size_t) i * dim + 1];
  p.z = dim == 3 ? xyz[(size_t) i * dim + 2] : 0.f;
  p.w = __int_as_float(i);
  op[i] = p;
  float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
  if (nrm) {
    q.x = nrm[(size_t) i * dim];
    q.y = nrm[(size_t) i * dim + 1];
    q.z = dim == 3 ? nrm[(size_t) i * dim + 2] : 0.f;
  }
  rec[2 * (size_t) i] = p;
  rec[2 * (size_t) i + 1] = q;
  if (inverse) inverse[i] = i;
}

__global__ void __launch_bounds__(256) proj_find_kernel(const SliceArgs a, const int* skip) {
  if (*a.stop || (skip && *skip)) return;
  __shared__ float S[16];
  if (threadIdx.x < 16) S[threadIdx.x] = a.S[threadIdx.x];
  __syncthreads();
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
    nn_finish<3>(a, S, q, i, 0.f, a.c_fpos[i]);
  }
}

// S_lb <- S after a stand-alone find (inside the ICP loop the solve kernel does this)
__global__ void commit_S_kernel(const float* S, float* S_lb) {
  if (threadIdx.x < 16 && blockIdx.x == 0) S_lb[threadIdx.x] = S[threadIdx.x];
}

// ---------------------------------------------------------------------------------------------
// k1b: per-correspondence linearisation + exact accumulation (a5); arithmetic in s2b_lin.cuh
// ---------------------------------------------------------------------------------------------
// one correspondence on the scalar path (tails, work lists): gate bookkeeping of the slot, status / chi
// output, accumulation
template <int DIM, int FACTOR>
__device__ __forceinline__ void lin_one_slot(const SliceArgs& a, const LinConst& k, int i, int slot, int bpos,
                                             const float4 m, const float4 nm, const float4 f, const float4 nf,
                                             LinAcc<DIM>& A) {
  int status;
  float chi;
  const bool ok = lin_one_scalar<DIM, FACTOR>(k, m, nm, f, nf, A, status, chi);
  if (a.gate && ok != (slot >= 0)) a.c_fpos[i] = ok ? bpos : -(bpos + 2);
  if (a.c_stat) a.c_stat[i] = (unsigned char) status;
  if (a.c_chi && ok) a.c_chi[i] = chi;
}

// exact integer block reduction of the per-thread partial sums: REDUX per slot (one when the
// per-thread sums are known to stay below 2^26, else two over 16-bit halves, which cannot overflow the
// 32-bit warp sum), the warp's 40 sums parked in lanes, one plain shared store per lane, then 40
// threads add the warps' rows and issue one global atomic per slot and CTA.
constexpr int kMaxWarps = 12;  // CTAs of the accumulating kernels have at most 384 threads (s2b_tiles.cuh: kLoopThreads)
struct FlushSmem {
  long long w[kMaxWarps][kAcc];
};

// first half: the warp's sums, parked in lanes (lane l keeps slot l in mine0 and slot 32 + l in mine1)
template <int DIM>
__device__ __forceinline__ void lin_warp_reduce(bool few, const LinAcc<DIM>& A, long long& mine0, long long& mine1) {
  constexpr int P = LinAcc<DIM>::P, NH = LinAcc<DIM>::NH;
  auto wsum = [few](int v) -> long long {
    if (few) return (long long) __reduce_add_sync(0xffffffffu, v);
    const unsigned lo = __reduce_add_sync(0xffffffffu, (unsigned) v & 0xffffu);
    const int hi = __reduce_add_sync(0xffffffffu, v >> 16);
    return ((long long) hi << 16) + (long long) lo;
  };
  const int lane = threadIdx.x & 31;
  // the accumulators hold raw bit patterns (see to_raw): take out count * bits(3.5f), modulo 2^32
  const int bias = A.n_terms * kFixBias;
  mine0 = 0; mine1 = 0;
#pragma unroll
  for (int k = 0; k < NH; ++k) {
    const long long v = wsum(A.aH[k] - bias);
    if (lane == k) mine0 = v;
  }
#pragma unroll
  for (int k = 0; k < P; ++k) {
    const long long v = wsum(A.ab[k] - bias);
    if (lane == kAccB + k) mine0 = v;
  }
  {
    // chi_in = chi_all - chi_out: both carry n_terms biases, so the difference carries none
    const int cv[4] = {A.chi_all - A.chi_out, A.chi_all_lo - A.chi_out_lo, A.chi_out - bias, A.chi_out_lo - bias};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long v = wsum(cv[k]);
      if (kAccChiIn + k < 32) { if (lane == kAccChiIn + k) mine0 = v; }
      else { if (lane == kAccChiIn + k - 32) mine1 = v; }
    }
    const int cn[4] = {A.n_io & 0xffff, (int) ((unsigned) A.n_io >> 16), A.n_ss & 0xffff, (int) ((unsigned) A.n_ss >> 16)};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long v = (long long) __reduce_add_sync(0xffffffffu, cn[k]);
      if (kAccNIn + k < 32) { if (lane == kAccNIn + k) mine0 = v; }
      else { if (lane == kAccNIn + k - 32) mine1 = v; }
    }
  }
}

// second half: the warps' rows through shared memory, one writer per slot.
// cta_dst: when non-null the CTA's sums are ADDED to these shared-memory words (one writer per slot) instead of
// going to the global accumulators with atomics -- the persistent loop publishes them once per iteration
__device__ __forceinline__ void lin_cta_reduce(unsigned long long* acc, long long mine0, long long mine1, FlushSmem& sm,
                                               long long* cta_dst = nullptr) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  sm.w[warp][lane] = mine0;
  if (lane < kAcc - 32) sm.w[warp][32 + lane] = mine1;
  __syncthreads();
  if (threadIdx.x < kAcc) {
    long long v = 0;
    const int nw = blockDim.x >> 5;
    for (int w = 0; w < nw; ++w) v += sm.w[w][threadIdx.x];
    if (cta_dst) cta_dst[threadIdx.x] += v;
    else if (v) atomicAdd(&acc[threadIdx.x], (unsigned long long) v);
  }
  __syncthreads();  // sm may be reused right away (the persistent loop flushes once per slice and iteration)
}

template <int DIM>
__device__ __forceinline__ void lin_flush(unsigned long long* acc, bool few, const LinAcc<DIM>& A, FlushSmem& sm,
                                          long long* cta_dst = nullptr) {
  long long mine0, mine1;
  lin_warp_reduce<DIM>(few, A, mine0, mine1);
  lin_cta_reduce(acc, mine0, mine1, sm, cta_dst);
}

// one thread's partial sums (a few scalar-path correspondences) into shared-memory accumulators, bias removed
template <int DIM>
__device__ __forceinline__ void lin_push_tail(const LinAcc<DIM>& A, long long* tail) {
  constexpr int P = LinAcc<DIM>::P, NH = LinAcc<DIM>::NH;
  if (A.n_terms == 0 && A.n_ss == 0) return;
  auto add = [tail](int slot, long long v) { if (v) atomicAdd(reinterpret_cast<unsigned long long*>(tail + slot), (unsigned long long) v); };
  const int bias = A.n_terms * kFixBias;
#pragma unroll
  for (int k = 0; k < NH; ++k) add(k, (long long) (A.aH[k] - bias));
#pragma unroll
  for (int k = 0; k < P; ++k) add(kAccB + k, (long long) (A.ab[k] - bias));
  add(kAccChiIn, (long long) (A.chi_all - A.chi_out)); add(kAccChiIn + 1, (long long) (A.chi_all_lo - A.chi_out_lo));
  add(kAccChiOut, (long long) (A.chi_out - bias)); add(kAccChiOut + 1, (long long) (A.chi_out_lo - bias));
  add(kAccNIn, A.n_io & 0xffff); add(kAccNOut, (unsigned) A.n_io >> 16);
  add(kAccNSup, A.n_ss & 0xffff); add(kAccNSat, (unsigned) A.n_ss >> 16);
}

// the same for a warp in which only lane 0 holds sums: the lanes take a slot each (one atomic per lane instead
// of 40 serial ones)
template <int DIM>
__device__ __forceinline__ void lin_push_tail_lane0(const LinAcc<DIM>& A, long long* tail) {
  constexpr int P = LinAcc<DIM>::P, NH = LinAcc<DIM>::NH;
  const int lane = threadIdx.x & 31;
  const int bias = A.n_terms * kFixBias;
  long long mine0 = 0, mine1 = 0;
  auto give = [&](int slot, int v) {
    const int t = __shfl_sync(0xffffffffu, v, 0);
    if (slot < 32) { if (lane == slot) mine0 = (long long) t; }
    else { if (lane == slot - 32) mine1 = (long long) t; }
  };
#pragma unroll
  for (int k = 0; k < NH; ++k) give(k, A.aH[k] - bias);
#pragma unroll
  for (int k = 0; k < P; ++k) give(kAccB + k, A.ab[k] - bias);
  give(kAccChiIn, A.chi_all - A.chi_out); give(kAccChiIn + 1, A.chi_all_lo - A.chi_out_lo);
  give(kAccChiOut, A.chi_out - bias); give(kAccChiOut + 1, A.chi_out_lo - bias);
  give(kAccNIn, A.n_io & 0xffff); give(kAccNOut, (int) ((unsigned) A.n_io >> 16));
  give(kAccNSup, A.n_ss & 0xffff); give(kAccNSat, (int) ((unsigned) A.n_ss >> 16));
  if (mine0) atomicAdd(reinterpret_cast<unsigned long long*>(tail + lane), (unsigned long long) mine0);
  if (lane < kAcc - 32 && mine1) atomicAdd(reinterpret_cast<unsigned long long*>(tail + 32 + lane), (unsigned long long) mine1);
}

// One THREAD per listed query does the whole job: exact search over all rings (nearest ring first, pruned by the
// best / second-best distance), slot + bound, linearisation.  The fallback of the iterations that launch a single
// search kernel (sole_list): long work lists and -- list == nullptr -- full searches that turn up late in a run.
// Slower per query than the phase-1 / phase-2 kernels, but it keeps three kernel launches out of every iteration
// that does not need them.
template <int DIM, bool TRACK2, int FACTOR>
__device__ __forceinline__ void nn_full_lin_body(const SliceArgs& a, const float* S, const int* rows, int K, float cell, int n,
                                                 const int* list, LinAcc<DIM>& A, const LinConst& lk, FlushSmem& fsm, int* park) {
  int done = 0;
  // (every thread of a CTA makes the same number of trips: the mid-loop flush below is a CTA-wide barrier)
  for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
    const int w = base + threadIdx.x;
    if (w < n) {
    const int i = list ? *reinterpret_cast<const volatile int*>(list + w) : w;
    NNQuery q;
    const float4 m = a.mp[i];
    nn_setup<DIM>(a, S, m, q);
    const int old_slot = __ldcg(a.c_fpos + i);
    const int p0 = slot_candidate(old_slot);
    if (a.warm && p0 >= 0) { nn_consider<DIM, TRACK2>(a, q, p0); if (TRACK2) nn_limit_bound(q, cell); }
    if (DIM == 3 && (a.nn_flat & 2)) {
      if (q.cy >= 0 && q.cy < a.ny && q.cz >= 0 && q.cz < a.nz) nn_scan_row<DIM, TRACK2>(a, q, q.cy, q.cz, 0.f);
      nn_scan_rings_flat<DIM, TRACK2>(a, q, rows, 1, K, cell, park, blockDim.x);
    } else
    for (int k = 0; k < K; ++k) {
      const int e = rows[k];
      const int dy = (int) (signed char) (e & 0xff), dz = (int) (signed char) ((e >> 8) & 0xff);
      const int ring = (e >> 16) & 0xff;
      const float pr2 = TRACK2 ? q.sd2 : q.bd2;
      if (ring >= 2) {  // every row of this and later rings is at least (ring - 1) cells away
        const float g = ((float) (ring - 1) - 4e-3f) * cell;
        if (g * g > pr2) break;
      }
      const int y = q.cy + dy, z = q.cz + dz;
      if (y < 0 || y >= a.ny || z < 0 || z >= a.nz) continue;
      const float gy = axis_gap(dy, q.fry) * cell;
      float lb2 = gy * gy;
      if (DIM == 3) {
        const float gz = axis_gap(dz, q.frz) * cell;
        lb2 = fmaf(gz, gz, lb2);
      }
      if (lb2 > pr2) continue;
      nn_scan_row<DIM, TRACK2>(a, q, y, z, lb2);
    }
    const int slot = nn_finish<DIM>(a, S, q, i, TRACK2 ? __fsqrt_rn(q.sd2) * (1.f - 1e-5f) : 0.f, old_slot);
    const int bpos = a.gate ? slot_candidate(slot) : slot;
    if (bpos >= 0) {
      lin_one_slot<DIM, FACTOR>(a, lk, i, slot, bpos, m, a.mn[i], __ldg(a.frec + 2 * (size_t) bpos), __ldg(a.frec + 2 * (size_t) bpos + 1), A);
      ++done;
    } else if (a.c_stat) {
      a.c_stat[i] = SRRG2B_STAT_NONE;
    }
    }
    // 32-bit partial sums: a thread stays below 512 terms per flush
    if (__syncthreads_or(done >= 400)) {
      lin_flush<DIM>(a.acc, false, A, fsm);
      A.clear();
      done = 0;
    }
  }
}

// Phase 2 / tail kernel.  Large work lists: the far list of phase 1 (see nn_far_body).  Short work
// lists: one warp per query does the whole job here -- search of all rows, slot + certified bound, and
// the linearisation of that query.
template <int DIM, int FACTOR>
__global__ void __launch_bounds__(256) nn_far_kernel(const SliceArgs a, const int* skip) {
  const int stop = *a.stop, list_all = *a.list_all, work_count = *a.work_count, far_count = *a.far_count;
  const int track2_flag = *a.track2;
  const int skipped = skip ? *skip : 0;
  if (stop || skipped) return;
  const bool all = !a.use_list || list_all;
  const int n_work = all ? a.nm : work_count;
  const bool tail = small_work_list(a, all, n_work);
  const bool sole_full = a.sole_list && !tail;  // the only search kernel of the iteration, and the list is long (or everything)
  const int n_far = (tail || sole_full) ? n_work : far_count;
  if (n_far == 0) return;
  __shared__ float S[16];
  __shared__ int rows[kRowTable];
  __shared__ FlushSmem fsm;
  __shared__ LinConst lk;
  __shared__ int park_all[DIM == 3 ? 3 * kRowChunk * 256 : 1];  // (2D: 2 rows per ring -- the flat chunks are compiled out)
  int* park = park_all + threadIdx.x;
  if (threadIdx.x < 16) S[threadIdx.x] = a.S[threadIdx.x];
  if (threadIdx.x == 32) make_lin_const(a, a.S, lk);
  const int R = a.R;
  const int K = (DIM == 3) ? (2 * R + 1) * (2 * R + 1) : (2 * R + 1);
  for (int k = threadIdx.x; k < K; k += blockDim.x)
    rows[k] = (DIM == 3) ? *reinterpret_cast<const int*>(c_rows3[k]) : *reinterpret_cast<const int*>(c_rows2[k]);
  __syncthreads();
  const float cell = __fdiv_rn(1.f, a.inv_cell);
  const bool track2 = track2_flag != 0;
  const int wpb = blockDim.x >> 5, w0 = blockIdx.x * wpb + (threadIdx.x >> 5), ws = gridDim.x * wpb;
  if (sole_full) {
    LinAcc<DIM> A;
    A.clear();
    const int* list = all ? nullptr : a.work_list;
    if (track2) nn_full_lin_body<DIM, true, FACTOR>(a, S, rows, K, cell, n_far, list, A, lk, fsm, park);
    else nn_full_lin_body<DIM, false, FACTOR>(a, S, rows, K, cell, n_far, list, A, lk, fsm, park);
    lin_flush<DIM>(a.acc, false, A, fsm);
    return;
  }
  if (!tail) {
    if (track2) nn_far_body<DIM, true>(a, S, rows, K, cell, n_far, a.far_list, nullptr, nullptr, w0, ws, park);
    else nn_far_body<DIM, false>(a, S, rows, K, cell, n_far, a.far_list, nullptr, nullptr, w0, ws, park);
    return;
  }
  LinAcc<DIM> A;
  A.clear();
  if (track2) nn_far_body<DIM, true, FACTOR>(a, S, rows, K, cell, n_far, a.work_list, &A, &lk, w0, ws, park);
  else nn_far_body<DIM, false, FACTOR>(a, S, rows, K, cell, n_far, a.work_list, &A, &lk, w0, ws, park);
  lin_flush<DIM>(a.acc, a.few_terms != 0, A, fsm);
}

// ---------------------------------------------------------------------------------------------
// k1s: per-iteration solve / update / statistics / termination (one thread; O(#slices) work)
// ---------------------------------------------------------------------------------------------
__device__ inline void ring_reset(Ring& r, int w) {
  r.window = w > kMaxWindow ? kMaxWindow : (w < 1 ? 1 : w);
  r.count = 0;
  r.head = 0;
}
__device__ inline void ring_add(Ring& r, double x) {
  r.v[r.head] = x;
  r.head = (r.head + 1) % r.window;
  if (r.count < r.window) r.count++;
}
__device__ inline double ring_max(const Ring& r) {
  double m = r.v[0];
  for (int i = 1; i < r.count; ++i) m = r.v[i] > m ? r.v[i] : m;
  return m;
}
__device__ inline double ring_min(const Ring& r) {
  double m = r.v[0];
  for (int i = 1; i < r.count; ++i) m = r.v[i] < m ? r.v[i] : m;
  return m;
}

// AlignerTerminationCriteriaStandard_::hasToStop,
// R/registration/aligners/aligner_termination_criteria_impl.cpp:24-65 (quirks at :32-33,:46,:53 kept)
__device__ inline bool has_to_stop(DevHeader* st, const SolveArgs& a, const srrg2b_iter_stats& s, long long ncorr) {
  ++st->tc_iterations;
  const int ninl = (int) s.num_inliers, nout = (int) s.num_outliers;
  const float chi = __fdiv_rn((float) s.chi_inliers, (float) ninl);
  if (!ninl) return false;
  ring_add(st->r_ncorr, (double) ncorr);
  ring_add(st->r_ninl, (double) ninl);
  ring_add(st->r_nout, (double) nout);
  ring_add(st->r_chi, (double) chi);
  if (st->r_ncorr.count < a.window) return false;
  if (ring_max(st->r_nout) - ring_min(st->r_nout) > (double) a.range_corr) return false;
  if (ring_max(st->r_ninl) - ring_min(st->r_ninl) > (double) a.range_inl) return false;
  const float crange = (float) ring_max(st->r_chi) - (float) ring_min(st->r_chi);
  const float cmax = (float) ring_max(st->r_chi);
  if (crange > (float) a.range_out) return false;
  if (__fdiv_rn(crange, cmax) > a.chi_eps) return false;
  return true;
}

// compute() prologue: variable <- guess, prior slices overwrite it in slice order
// (multi_aligner_impl.cpp:130-141), finder transforms, counters
// SolveArgs live in device memory (refreshed before every run) so that per-call values such as a
// prior's measurement do not change the launch sequence a cached CUDA graph replays
__device__ __forceinline__ void load_solve_args(const SolveArgs* ap, SolveArgs* sh) {
  const int* src = reinterpret_cast<const int*>(ap);
  int* dst = reinterpret_cast<int*>(sh);
  for (int k = threadIdx.x; k < (int) (sizeof(SolveArgs) / sizeof(int)); k += blockDim.x) dst[k] = src[k];
  __syncthreads();
}

__global__ void icp_init_kernel(const SolveArgs* ap, DevState* st, const Mat4f* T0, int apply_prior_guess, int reset_tc,
                                int keep_stats, int iterations) {
  __shared__ SolveArgs a;
  load_solve_args(ap, &a);
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (!keep_stats) {
    Mat4f X = *T0;
    if (apply_prior_guess) {
      for (int s = 0; s < a.n_slices; ++s)
        if (a.sl[s].kind == SRRG2B_SLICE_PRIOR) X = a.sl[s].Z;
    }
    st->X = X;
    st->n_stats = 0;
    st->iterations_run = 0;
  }
  st->iterations_left = iterations;
  st->error = 0;
  for (int s = 0; s < a.n_slices; ++s) {
    compose(a.sl[s].ris, st->X, st->S[s]);
    for (int k = 0; k < kAcc; ++k) st->acc[s][k] = 0ull;
    st->ncorr[s] = 0;
    // the motion from the previous call's pose is unknown: certify bounds only when forced to
    if (!keep_stats) st->track2[s] = (a.sl[s].track2_mode == 1) ? 1 : 0;
    // a fresh compute() starts from an arbitrary guess: search everything; the inlier-only second
    // run continues from certified bounds
    if (!keep_stats) st->list_all[s] = 1;  // (the last solve step already set it for a continued run)
    if (a.sl[s].kind == SRRG2B_SLICE_POINTS && a.sl[s].counters) { a.sl[s].counters[0] = 0; a.sl[s].counters[1] = 0; a.sl[s].counters[2] = 0; }
    if (!keep_stats && a.sl[s].kind == SRRG2B_SLICE_POINTS && a.sl[s].S_lb) {
      // a fresh run rewrites every bound in its first pass: that pass is epoch 0
      const Mat4f& S = st->S[s];
      const float tn = sqrtf(S.m[3] * S.m[3] + S.m[7] * S.m[7] + S.m[11] * S.m[11]);
      float* B = a.sl[s].S_lb;
      st->ep_next[s] = 1;
      B[kSlbDcert] = 0.f;
      reinterpret_cast<int*>(B)[kSlbEpoch] = 0;
      for (int e = 0; e < kEpochs; ++e) B[kSlbDtab + e] = 3e38f;
      B[kSlbDtab] = 1e-6f * (a.sl[s].radius * 1.0001f + tn) + 2e-7f;
      for (int j = 0; j < 12; ++j) B[kSlbEpS + j] = S.m[j];
    }
  }
  if (!keep_stats) st->certified = 0;
  st->stop = 0;
  st->not_enough_corr = 0;
  if (reset_tc) {
    ring_reset(st->r_ncorr, a.window);
    ring_reset(st->r_ninl, a.window);
    ring_reset(st->r_nout, a.window);
    ring_reset(st->r_chi, a.window);
    st->tc_iterations = 0;
  }
}

// set only the finder transform of one slice (stand-alone find / linearise entry points); motion is the
// host-computed bound of how far any query moved since the slice's previous stand-alone pass
__global__ void set_S_kernel(DevState* st, int slice, Mat4f S, int track2, float* S_lb, float motion, float slack) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    st->S[slice] = S;
    st->track2[slice] = track2;
    st->list_all[slice] = 1;
    for (int k = 0; k < kAcc; ++k) st->acc[slice][k] = 0ull;
    st->stop = 0;
    if (S_lb) {
      // stand-alone passes use epoch 0 only, anchored at the first pass after a reset.  motion: bound of any
      // query's displacement between S and that anchor (host-computed): bounds certified now are stored
      // relative to the anchor, the check subtracts the same displacement (+ slack)
      S_lb[kSlbDcert] = motion;
      reinterpret_cast<int*>(S_lb)[kSlbEpoch] = 0;
      S_lb[kSlbDtab] = (motion + slack) * (1.f + 2.4e-7f);
    }
  }
}

// body of one _runSolver iteration after the per-slice kernels
// (R/registration/aligners/multi_aligner_impl.cpp:106-126), serial part: runs on one thread against the
// shared-memory copy of the state; the IterationStats entry goes straight to global memory
#ifndef S2B_SOLVE_STAMPS
#define S2B_SOLVE_STAMPS 0
#endif
__device__ unsigned long long g_solve_stamps[8];
__device__ __forceinline__ void solve_stamp(int k) {
#if S2B_SOLVE_STAMPS
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  g_solve_stamps[k] = t;
#endif
}

template <int DIM>
// conv[k][slot]: the accumulators of slice k already converted to double and scaled (done by the whole CTA in
// parallel: 35 int64 -> fp64 conversions per slice are the slowest part of the assembly when one thread does them)
__device__ void icp_solve_serial(const SolveArgs& a, DevHeader& st, srrg2b_iter_stats* stats_out, const double (*conv)[kAcc]) {
  constexpr int P = (DIM == 3) ? 6 : 3;
  double H[P * P], b[P];
#pragma unroll
  for (int i = 0; i < P * P; ++i) H[i] = 0.0;
#pragma unroll
  for (int i = 0; i < P; ++i) b[i] = 0.0;
  srrg2b_iter_stats s;
  s.iteration = st.n_stats;
  s.solver_status = 0;
  s.num_inliers = 0; s.num_outliers = 0; s.num_suppressed = 0; s.num_correspondences = 0; s.num_saturated = 0;
  s.chi_inliers = 0.0; s.chi_outliers = 0.0;
  bool good = false;
  long long total = 0;
  Mat4f X = st.X;
  for (int k = 0; k < a.n_slices; ++k) {
    const SolveSlice& sl = a.sl[k];
    if (sl.kind == SRRG2B_SLICE_PRIOR) {
      double chi = 0.0;
      double Hf[36], bf[6];  // prior_accumulate is written for run-time P
      for (int i = 0; i < P * P; ++i) Hf[i] = H[i];
      for (int i = 0; i < P; ++i) bf[i] = b[i];
      prior_accumulate(DIM, a.variable, sl.Z, X, sl.info, Hf, bf, chi);
      for (int i = 0; i < P * P; ++i) H[i] = Hf[i];
      for (int i = 0; i < P; ++i) b[i] = bf[i];
      s.num_inliers += 1; s.num_correspondences += 1; s.chi_inliers += chi;
      good = true;  // aligner_slice_processor_prior.h:65-67
      total += 1;   // :75-77
      st.ncorr[k] = 1;
      continue;
    }
    const unsigned long long* acc = st.acc[k];
    // the slice's H (both triangles) and b are added entry by entry in slice order
    int slot = 0;
#pragma unroll
    for (int i = 0; i < P; ++i) {
#pragma unroll
      for (int j = i; j < P; ++j) {
        const double v = conv[k][slot++];
        H[i * P + j] = H[i * P + j] + v;
        if (j != i) H[j * P + i] = H[j * P + i] + v;
      }
    }
#pragma unroll
    for (int i = 0; i < P; ++i)
      b[i] = b[i] + conv[k][kAccB + i];
    const long long ni = (long long) acc[kAccNIn], no = (long long) acc[kAccNOut], ns = (long long) acc[kAccNSup];
    s.num_inliers += ni; s.num_outliers += no; s.num_suppressed += ns; s.num_correspondences += ni + no + ns;
    s.num_saturated += (long long) acc[kAccNSat];
    s.chi_inliers += conv[k][kAccChiIn] + conv[k][kAccChiIn + 1];
    s.chi_outliers += conv[k][kAccChiOut] + conv[k][kAccChiOut + 1];
    const long long n = ni + no + ns;
    st.ncorr[k] = n;
    total += n;
    good = good || (n > (long long) sl.min_corr);  // aligner_slice_processor_impl.cpp:77-79
  }
  solve_stamp(2);
  st.iterations_run += 1;
  st.iterations_left -= 1;
  if (st.iterations_left <= 0) st.stop = 1;
  if (!good) {  // multi_aligner_impl.cpp:107-111 (estimate already equals the backup)
    st.not_enough_corr = 1;
    st.stop = 1;
    return;
  }
  double dx[6] = {0, 0, 0, 0, 0, 0};
  const bool solved = spd_solve_t<P>(H, b, dx);
  solve_stamp(3);
  if (solved) {
    box_plus(DIM, a.variable, dx, X);
    st.X = X;
    s.solver_status = 1;
  }
  solve_stamp(4);
  if (st.n_stats < kMaxStats) stats_out[st.n_stats] = s;
  st.last_stats = s;
  st.n_stats += 1;
  int certified = 1;
  for (int k = 0; k < a.n_slices; ++k) {
    Mat4f Sn;
    compose(a.sl[k].ris, X, Sn);
    if (a.sl[k].kind == SRRG2B_SLICE_POINTS) {
      // |dS m| <= |dR|_F |m|_max + |dt| bounds how far any query of the slice moves between this iteration and
      // the next: it decides whether the next NN pass certifies bounds (track2: once the motion is small against
      // the cell edge).  The displacement tables of the epochs are refreshed by the whole CTA afterwards
      // (solve_refresh_epochs).
      double dr = 0.0, dt = 0.0;
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) {
          const double d = (double) Sn.m[r * 4 + c] - (double) st.S[k].m[r * 4 + c];
          dr += d * d;
        }
        const double d = (double) Sn.m[r * 4 + 3] - (double) st.S[k].m[r * 4 + 3];
        dt += d * d;
      }
      const float motion = (float) ((sqrt(dr) * (double) a.sl[k].radius + sqrt(dt)) * (1.0 + 1e-6)) * (1.f + 2.4e-7f);
      const int mode = a.sl[k].track2_mode;
      st.list_all[k] = st.track2[k] ? 0 : 1;  // bounds exist only if the pass just done certified them
      st.track2[k] = (mode == 1) || (mode == 2 && motion < a.sl[k].track2_frac * a.sl[k].cell) ? 1 : 0;
      // every pass is an epoch of the slice's bounds; a full pass rewrites every bound and restarts the ids,
      // and so does running out of ids
      if (st.ep_next[k] >= kEpochs) st.list_all[k] = 1;
      if (st.list_all[k]) st.ep_next[k] = 0;
      if (a.sl[k].nn_points && st.list_all[k]) certified = 0;
    }
    st.S[k] = Sn;
  }
  st.certified = certified;
  solve_stamp(5);
  if (a.use_tc && has_to_stop(&st, a, s, total)) st.stop = 1;
  solve_stamp(6);
}

// What the coherence check subtracts from a bound of an epoch whose transform is E (12 floats), at the transform
// Sn: a bound of any query's displacement between the two, |dS m| <= |dR|_F |m|_max + |dt|, plus absolute slack
// for the fp32 rounding of the two transformed queries (each component: 4 roundings of partial sums bounded by
// qmax = |m|_max + |t|, i.e. <= 4.1e-7 qmax per query as a vector) and of the bound arithmetic.
__device__ __forceinline__ float epoch_displacement(const float* Sn, const float* E, float radius, bool same) {
  double dr = 0.0, dt = 0.0, tn = 0.0;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) {
      const double d = same ? 0.0 : (double) Sn[r * 4 + c] - (double) __ldcg(E + r * 4 + c);
      dr += d * d;
    }
    const double d = same ? 0.0 : (double) Sn[r * 4 + 3] - (double) __ldcg(E + r * 4 + 3);
    dt += d * d;
    tn += (double) Sn[r * 4 + 3] * (double) Sn[r * 4 + 3];
  }
  const float Dm = (float) ((sqrt(dr) * (double) radius + sqrt(dt)) * (1.0 + 1e-6)) * (1.f + 2.4e-7f);
  const float qmax = radius * 1.0001f + (float) sqrt(tn);
  return (Dm + (1e-6f * qmax + 2e-7f * (1.f + Dm))) * (1.f + 2.4e-7f);
}
// The same, split per query: the displacement of query m is at most kr |m| + kt (kr = |dR|_F rounded up, kt =
// |dt| + the absolute slack, which is computed for the worst query).
__device__ __forceinline__ float2 epoch_displacement2(const float* Sn, const float* E, float radius, bool same) {
  double dr = 0.0, dt = 0.0, tn = 0.0;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) {
      const double d = same ? 0.0 : (double) Sn[r * 4 + c] - (double) __ldcg(E + r * 4 + c);
      dr += d * d;
    }
    const double d = same ? 0.0 : (double) Sn[r * 4 + 3] - (double) __ldcg(E + r * 4 + 3);
    dt += d * d;
    tn += (double) Sn[r * 4 + 3] * (double) Sn[r * 4 + 3];
  }
  // (the rounding slack 1e-6 (|m| + |t|) of the two transformed queries is split the same way: the partial sums of
  //  query m are bounded by 1.001 |m| + |t|, not only by the slice-wide maximum)
  const float kr0 = (float) (sqrt(dr) * (1.0 + 1e-6)) * (1.f + 2.4e-7f);
  const float Dm = (kr0 * radius + (float) (sqrt(dt) * (1.0 + 1e-6))) * (1.f + 4.8e-7f);
  const float kr = (kr0 + 1.0011e-6f) * (1.f + 2.4e-7f);
  const float kt = ((float) (sqrt(dt) * (1.0 + 1e-6)) + (1e-6f * (float) (sqrt(tn) * (1.0 + 1e-6)) + 2e-7f * (1.f + Dm))) * (1.f + 4.8e-7f);
  return make_float2(kr, kt);
}

// After the serial part: the next pass of every point slice becomes epoch ep_next: its transform and id are
// recorded in the slice's bound state.  The displacement table dtab is computed by whoever runs the pass
// (the loop kernel's CTAs, in parallel with their pipeline prologue: loop_load_lin_const), so the solve step
// only publishes.  (Every thread of the CTA calls this; sh.S[k] already holds the NEXT transform.)
__device__ __forceinline__ void solve_refresh_epochs(const SolveArgs& a, DevHeader& sh) {
  const int tid = threadIdx.x;
  for (int k = 0; k < a.n_slices; ++k) {
    if (a.sl[k].kind != SRRG2B_SLICE_POINTS || !a.sl[k].S_lb) continue;
    float* B = a.sl[k].S_lb;
    const int ep = sh.ep_next[k];  // (< kEpochs: the serial part restarts the ids in time)
    if (tid < 12) B[kSlbEpS + 12 * ep + tid] = sh.S[k].m[tid];
    if (tid == 32) { B[kSlbDcert] = 0.f; reinterpret_cast<int*>(B)[kSlbEpoch] = ep; }
  }
  __syncthreads();
  if (tid == 0)
    for (int k = 0; k < a.n_slices; ++k)
      if (a.sl[k].kind == SRRG2B_SLICE_POINTS && a.sl[k].S_lb) sh.ep_next[k] += 1;
  __syncthreads();
}

constexpr int kSolveThreads = 256;
constexpr int kMaxRanks = 16;
constexpr int kMailWords = 2 * SRRG2B_MAX_SLICES * kAcc;  // two tagged 8-byte words per accumulator

// Multi-GPU exchange of the integer accumulators, fused into the solve step: every rank owns a
// two-slot mailbox in its own HBM that the peers map through CUDA IPC (NVLink / NVSwitch peer loads).
// Low-latency protocol (the idea of NCCL's LL): every 8-byte mailbox word carries 32 bits of payload
// and the 32-bit epoch tag, and an aligned 8-byte store is indivisible, so a reader simply polls the
// word until the tag matches -- no fence, no separate flag, one NVLink traversal.  Epoch e uses slot
// e & 1; a slot is rewritten at epoch e + 2, which a rank reaches only after every peer delivered
// e + 1, i.e. finished reading e -- so two slots suffice.  Integer sums: every rank gets the same bits
// whatever the order.  No NCCL call inside the iteration: the run stays one CUDA graph.
// A peer that never delivers (crashed process, diverged launch sequence) is detected by a clock-based
// timeout of the poll: the rank sets the sticky error flag, stops, and the host returns SRRG2B_ERR_NCCL.
struct PeerExchange {
  int rank, world;
  unsigned long long* mail[kMaxRanks];       // mail[r]: rank r's mailbox (own or IPC-mapped), 2 * kMailWords
  long long timeout_cycles;                  // poll budget (clock64 ticks), ~2 s by default
};

__device__ __forceinline__ void st_volatile_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_volatile_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// shared-memory staging of the solve step
struct SolveSmem {
  alignas(16) SolveArgs a;
  alignas(16) DevHeader sh;
  PeerExchange pe;
  int timed_out;
  double conv[SRRG2B_MAX_SLICES][kAcc];  // accumulators as scaled doubles (see icp_solve_serial)
};

// The solve step of one iteration, executed by ONE CTA with at least kSolveThreads threads (the first
// kSolveThreads take part): stage arguments + state, all-reduce over the peers, serial solve, write back.
// Returns (to every participating thread) whether the iteration loop has to stop.
template <int DIM>
__device__ __forceinline__ bool icp_solve_block(const SolveArgs* ap, DevState* st, const PeerExchange* px, SolveSmem& sm) {
  const int tid = threadIdx.x;
  const bool part = tid < kSolveThreads;
  if (tid == 0) solve_stamp(0);
  SolveArgs& a = sm.a;
  DevHeader& sh = sm.sh;
  PeerExchange& pe = sm.pe;
  if (part) {
    // one round of independent 16-byte loads stages the arguments and the whole mutable state
    const int4* s0 = reinterpret_cast<const int4*>(ap);
    int4* d0 = reinterpret_cast<int4*>(&a);
    for (int k = tid; k < (int) (sizeof(SolveArgs) / 16); k += kSolveThreads) d0[k] = s0[k];
    const int4* s1 = reinterpret_cast<const int4*>(static_cast<const DevHeader*>(st));
    int4* d1 = reinterpret_cast<int4*>(&sh);
    for (int k = tid; k < (int) (sizeof(DevHeader) / 16); k += kSolveThreads) d1[k] = __ldcg(s1 + k);
    if (px) {
      const int* s2 = reinterpret_cast<const int*>(px);
      int* d2 = reinterpret_cast<int*>(&pe);
      for (int k = tid; k < (int) (sizeof(PeerExchange) / sizeof(int)); k += kSolveThreads) d2[k] = s2[k];
    }
    if (tid == 0) sm.timed_out = 0;
  }
  __syncthreads();
  if (sh.stop) return true;  // (every rank holds the same state, so every rank returns here or none does)
  if (px) {
    // all-reduce of the accumulators over peer memory (see PeerExchange)
    const unsigned long long e = sh.epoch + 1ull;
    const unsigned long long tag = (e & 0xffffffffull) << 32;
    const int n_words = a.n_slices * kAcc;
    unsigned long long* mine = pe.mail[pe.rank] + (e & 1ull) * kMailWords;
    if (part) {
      for (int k = tid; k < n_words; k += kSolveThreads) {
        const unsigned long long v = (&sh.acc[0][0])[k];
        st_volatile_sys(mine + 2 * k, (v & 0xffffffffull) | tag);
        st_volatile_sys(mine + 2 * k + 1, (v >> 32) | tag);
      }
    }
    // (peer, word) pairs are dealt to the threads: both halves are requested together and polled until
    // their tags are e, then added to the shared copy.  Everything a thread waits for is in flight at
    // once, so the exchange costs about one NVLink round trip per ceil(pairs / threads).
    __syncthreads();
    const int n_pairs = n_words * (pe.world - 1);
    if (part) {
      const long long t0 = clock64();
      for (int t = tid; t < n_pairs; t += kSolveThreads) {
        int r = t / n_words;
        const int k = t - r * n_words;
        if (r >= pe.rank) ++r;  // peers in rank order, skipping this rank
        const unsigned long long* theirs = pe.mail[r] + (e & 1ull) * kMailWords + 2 * k;
        unsigned long long lo, hi;
        bool ok = true;
        int spins = 0;
        do {
          lo = ld_volatile_sys(theirs);
          hi = ld_volatile_sys(theirs + 1);
          if ((++spins & 1023) == 0 && clock64() - t0 > pe.timeout_cycles) { ok = false; break; }
        } while ((lo & 0xffffffff00000000ull) != tag || (hi & 0xffffffff00000000ull) != tag);
        if (!ok) { sm.timed_out = 1; break; }
        atomicAdd(&sh.acc[0][0] + k, (lo & 0xffffffffull) | ((hi & 0xffffffffull) << 32));
      }
    }
    __syncthreads();
    if (sm.timed_out) {  // a peer never delivered: stop here, the host reports SRRG2B_ERR_NCCL
      if (tid == 0) { st->error = 1; st->stop = 1; __threadfence(); }
      return true;
    }
    if (tid == 0) sh.epoch = e;  // (written back with the rest of the state)
  }
  // the NN pass of this iteration certified its bounds at S: record that before anything can bail out;
  // the work-list counters start the next iteration at zero
  for (int k = 0; k < a.n_slices; ++k) {
    if (a.sl[k].kind != SRRG2B_SLICE_POINTS) continue;
    if (a.sl[k].S_lb && tid < 16) a.sl[k].S_lb[tid] = sh.S[k].m[tid];
    if (a.sl[k].counters && tid >= 32 && tid < 35) a.sl[k].counters[tid - 32] = 0;
  }
  {
    // the accumulators as scaled doubles, one (slice, slot) pair per thread
    constexpr int P = (DIM == 3) ? 6 : 3, NH = P * (P + 1) / 2;
    __syncthreads();
    for (int t = tid; t < a.n_slices * kAcc; t += blockDim.x) {
      const int k = t / kAcc, slot = t - k * kAcc;
      int cls = -1;
      if (slot < NH) {
        int i = 0, rem = slot;  // slot -> (i, j) of the upper triangle, row-major
        while (rem >= P - i) { rem -= P - i; ++i; }
        const int j = i + rem;
        cls = (j < DIM) ? kKHtt : ((i < DIM) ? kKHtr : kKHrr);
      } else if (slot >= kAccB && slot < kAccB + P) {
        cls = (slot - kAccB < DIM) ? kKBt : kKBr;
      } else if (slot == kAccChiIn || slot == kAccChiOut) {
        cls = kKChi;
      } else if (slot == kAccChiIn + 1 || slot == kAccChiOut + 1) {
        cls = kKChiLo;
      }
      sm.conv[k][slot] = cls >= 0 ? __ll2double_rn((long long) sh.acc[k][slot]) * a.sl[k].invk[cls] : 0.0;
    }
    __syncthreads();
  }
  if (tid == 0) { solve_stamp(1); icp_solve_serial<DIM>(a, sh, st->stats, sm.conv); }
  __syncthreads();
  solve_refresh_epochs(a, sh);
  // accumulators restart at zero; everything else goes back as the serial part left it
  if (part)
    for (int k = tid; k < a.n_slices * kAcc; k += kSolveThreads) sh.acc[k / kAcc][k % kAcc] = 0ull;
  __syncthreads();
  if (part) {
    const int4* s1 = reinterpret_cast<const int4*>(&sh);
    int4* d1 = reinterpret_cast<int4*>(static_cast<DevHeader*>(st));
    for (int k = tid; k < (int) (sizeof(DevHeader) / 16); k += kSolveThreads) d1[k] = s1[k];
  }
  if (tid == 0) solve_stamp(7);
  return sh.stop != 0;
}

// stand-alone solve step (the first iterations of a run, which search with the dedicated NN kernels)
template <int DIM>
__global__ void __launch_bounds__(kSolveThreads) icp_solve_kernel(const SolveArgs* ap, DevState* st, const PeerExchange* px,
                                                                  const int* skip) {
  if (skip && *skip) return;
  __shared__ SolveSmem sm;
  icp_solve_block<DIM>(ap, st, px, sm);
}

// ---------------------------------------------------------------------------------------------
// N1: the clip of the tracker slice on the device.  TrackerSliceProcessor_::clip
// (R/trackers/tracker_slice_processor_impl.cpp:194-205) hands the full scene to a SceneClipper, whose contract is
// "clipped scene in the robot frame + the indices of its points in the full scene" (R/mapping/scene_clipper.h:104-107).
// Range clipper: keep a valid scene point iff |T p| <= max_range; output T p, R n in ascending scene index.
// Two passes around a CUB exclusive scan: flag, then transform + compact.
// ---------------------------------------------------------------------------------------------
template <int DIM>
__global__ void scene_clip_flag_kernel(const float* __restrict__ xyz, const unsigned char* __restrict__ valid, int n, Mat4f T,
                                       float r2, int* __restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int keep = 0;
  if (!valid || valid[i]) {
    const float4 m = make_float4(xyz[(size_t) i * DIM], xyz[(size_t) i * DIM + 1], DIM == 3 ? xyz[(size_t) i * DIM + 2] : 0.f, 0.f);
    float qx, qy, qz;
    nn_transform<DIM>(T.m, m, qx, qy, qz);
    float d2 = fmaf(qy, qy, qx * qx);
    if (DIM == 3) d2 = fmaf(qz, qz, d2);
    keep = d2 <= r2 ? 1 : 0;
  }
  flag[i] = keep;
}

template <int DIM>
__global__ void scene_clip_compact_kernel(const float* __restrict__ xyz, const float* __restrict__ nrm, const int* __restrict__ flag,
                                          const int* __restrict__ pos, int n, Mat4f T, float* __restrict__ out_xyz,
                                          float* __restrict__ out_nrm, int* __restrict__ gidx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !flag[i]) return;
  const int k = pos[i];
  const float* S = T.m;
  const float4 m = make_float4(xyz[(size_t) i * DIM], xyz[(size_t) i * DIM + 1], DIM == 3 ? xyz[(size_t) i * DIM + 2] : 0.f, 0.f);
  float qx, qy, qz;
  nn_transform<DIM>(S, m, qx, qy, qz);
  out_xyz[(size_t) k * DIM] = qx;
  out_xyz[(size_t) k * DIM + 1] = qy;
  if (DIM == 3) out_xyz[(size_t) k * DIM + 2] = qz;
  if (nrm) {
    const float nx = nrm[(size_t) i * DIM], ny = nrm[(size_t) i * DIM + 1], nz = DIM == 3 ? nrm[(size_t) i * DIM + 2] : 0.f;
    float t;
    t = S[0] * nx; t = fmaf(S[1], ny, t); if (DIM == 3) t = fmaf(S[2], nz, t); out_nrm[(size_t) k * DIM] = t;
    t = S[4] * nx; t = fmaf(S[5], ny, t); if (DIM == 3) t = fmaf(S[6], nz, t); out_nrm[(size_t) k * DIM + 1] = t;
    if (DIM == 3) { t = S[8] * nx; t = fmaf(S[9], ny, t); t = fmaf(S[10], nz, t); out_nrm[(size_t) k * DIM + 2] = t; }
  }
  gidx[k] = i;
}

// ---------------------------------------------------------------------------------------------
// N2: MergerCorrespondenceHomo_::compute(), R/mapping/merger_correspondence_homo_impl.cpp:11-126, on
// the device-resident scene.  The aligner's correspondences never leave the GPU: d_flag / d_fidx / d_resp are the
// dense export of the slice (per local moving index k: keep flag, index j of the measurement point -- the aligner's
// fixed cloud --, response), gidx maps k to the scene point it was clipped from (the flip + local_to_global mapping of
// R/trackers/tracker_slice_processor_impl.cpp:159-191).  Every scene point appears in at most one correspondence
// (one entry per moving point, the clip is injective), so the merges are conflict free; merged[j] is an idempotent mark.
// ---------------------------------------------------------------------------------------------
template <int DIM>
__global__ void scene_merge_kernel(const int* __restrict__ d_flag, const int* __restrict__ d_fidx, const float* __restrict__ d_resp,
                                   const int* __restrict__ gidx, int n_local, const float* __restrict__ meas_xyz,
                                   const float* __restrict__ meas_nrm, Mat4f T, float max_response, float max_d2,
                                   float* __restrict__ scene_xyz, float* __restrict__ scene_nrm, int* __restrict__ merged) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_local || !d_flag[k]) return;
  const int j = d_fidx[k], g = gidx[k];
  if (j < 0 || !(d_resp[k] < max_response)) return;
  const float4 m = make_float4(meas_xyz[(size_t) j * DIM], meas_xyz[(size_t) j * DIM + 1], DIM == 3 ? meas_xyz[(size_t) j * DIM + 2] : 0.f, 0.f);
  float q[3];
  nn_transform<DIM>(T.m, m, q[0], q[1], q[2]);
  float sp[3] = {scene_xyz[(size_t) g * DIM], scene_xyz[(size_t) g * DIM + 1], DIM == 3 ? scene_xyz[(size_t) g * DIM + 2] : 0.f};
  const float dx = q[0] - sp[0], dy = q[1] - sp[1], dz = q[2] - sp[2];
  float d2 = fmaf(dy, dy, dx * dx);
  if (DIM == 3) d2 = fmaf(dz, dz, d2);
  if (!(d2 < max_d2)) return;
#pragma unroll
  for (int c = 0; c < DIM; ++c) {
    scene_xyz[(size_t) g * DIM + c] = (q[c] + sp[c]) * 0.5f;
    if (scene_nrm && meas_nrm) scene_nrm[(size_t) g * DIM + c] = meas_nrm[(size_t) j * DIM + c];  // (copied as it is: :71-74)
  }
  merged[j] = 1;
}

// the measurement points that were not merged (and are valid): flag for the append
__global__ void scene_append_flag_kernel(const int* __restrict__ merged, const unsigned char* __restrict__ valid, int n, int* __restrict__ flag) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) flag[j] = (!merged[j] && (!valid || valid[j])) ? 1 : 0;
}

// append the flagged measurement points, transformed in place (coordinates and normal), behind the scene's n_scene points
template <int DIM>
__global__ void scene_append_kernel(const int* __restrict__ flag, const int* __restrict__ pos, int n, const float* __restrict__ meas_xyz,
                                    const float* __restrict__ meas_nrm, Mat4f T, int n_scene, float* __restrict__ scene_xyz,
                                    float* __restrict__ scene_nrm, unsigned char* __restrict__ scene_valid) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n || !flag[j]) return;
  const size_t o = (size_t) n_scene + pos[j];
  const float* S = T.m;
  const float4 m = make_float4(meas_xyz[(size_t) j * DIM], meas_xyz[(size_t) j * DIM + 1], DIM == 3 ? meas_xyz[(size_t) j * DIM + 2] : 0.f, 0.f);
  float qx, qy, qz;
  nn_transform<DIM>(S, m, qx, qy, qz);
  scene_xyz[o * DIM] = qx;
  scene_xyz[o * DIM + 1] = qy;
  if (DIM == 3) scene_xyz[o * DIM + 2] = qz;
  if (scene_nrm && meas_nrm) {
    const float nx = meas_nrm[(size_t) j * DIM], ny = meas_nrm[(size_t) j * DIM + 1], nz = DIM == 3 ? meas_nrm[(size_t) j * DIM + 2] : 0.f;
    float t;
    t = S[0] * nx; t = fmaf(S[1], ny, t); if (DIM == 3) t = fmaf(S[2], nz, t); scene_nrm[o * DIM] = t;
    t = S[4] * nx; t = fmaf(S[5], ny, t); if (DIM == 3) t = fmaf(S[6], nz, t); scene_nrm[o * DIM + 1] = t;
    if (DIM == 3) { t = S[8] * nx; t = fmaf(S[9], ny, t); t = fmaf(S[10], nz, t); scene_nrm[o * DIM + 2] = t; }
  }
  if (scene_valid) scene_valid[o] = 1;
}

// ---------------------------------------------------------------------------------------------
// k2b: export (sorted order -> dense by local moving index), then compaction
// ---------------------------------------------------------------------------------------------
__global__ void export_dense_kernel(const float4* __restrict__ mp, const float4* __restrict__ fp,
                                    const int* __restrict__ c_fpos, const int* __restrict__ c_fidx,
                                    const float* __restrict__ S_lb, int dim, const unsigned char* __restrict__ c_stat,
                                    const float* __restrict__ c_chi, int nm, int prune, int* __restrict__ d_fidx,
                                    float* __restrict__ d_resp, unsigned char* __restrict__ d_stat,
                                    float* __restrict__ d_chi, int* __restrict__ d_flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nm) return;
  const int src = __float_as_int(mp[i].w);
  const int slot = c_fpos[i];
  int fi = -1;
  if (slot >= 0) fi = __float_as_int(fp[slot].w);
  else if (slot == kSlotSuppressed) fi = c_fidx[i];  // externally supplied, kept as given
  const bool keep = (slot >= 0 || slot == kSlotSuppressed) && (!prune || (c_stat && c_stat[i] == SRRG2B_STAT_INLIER));
  d_flag[src] = keep ? 1 : 0;
  d_fidx[src] = fi;
  float resp = 0.f;
  if (slot >= 0) {  // response = |S m - f| at the transform of the last NN pass (pinned arithmetic)
    const float4 m = mp[i];
    const float4 f = fp[slot];
    float qx, qy, qz;
    if (dim == 3) nn_transform<3>(S_lb, m, qx, qy, qz); else nn_transform<2>(S_lb, m, qx, qy, qz);
    const float ddx = qx - f.x, ddy = qy - f.y, ddz = qz - f.z;
    float d2 = fmaf(ddy, ddy, ddx * ddx);
    if (dim == 3) d2 = fmaf(ddz, ddz, d2);
    resp = __fsqrt_rn(d2);
  }
  d_resp[src] = resp;
  if (d_stat) d_stat[src] = c_stat ? c_stat[i] : (unsigned char) SRRG2B_STAT_NONE;
  if (d_chi) d_chi[src] = c_chi ? c_chi[i] : 0.f;
}

__global__ void compact_kernel(const int* __restrict__ flag, const int* __restrict__ pos, int n, int index_offset,
                               const int* __restrict__ d_fidx, const float* __restrict__ d_resp,
                               const unsigned char* __restrict__ d_stat, const float* __restrict__ d_chi,
                               int* __restrict__ o_fidx, int* __restrict__ o_midx, float* __restrict__ o_resp,
                               unsigned char* __restrict__ o_stat, float* __restrict__ o_chi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !flag[i]) return;
  const int k = pos[i];
  o_fidx[k] = d_fidx[i];
  o_midx[k] = i + index_offset;
  o_resp[k] = d_resp[i];
  if (o_stat) o_stat[k] = d_stat[i];
  if (o_chi) o_chi[k] = d_chi[i];
}

// external correspondences (HBST path): list -> sorted-order slots; -2 marks "cannot evaluate"
__global__ void import_corr_kernel(const int* __restrict__ fixed_idx, const int* __restrict__ moving_idx, int n,
                                   const int* __restrict__ m_inverse, const int* __restrict__ f_inverse, int nm_raw,
                                   int nf_raw, int index_offset, int* __restrict__ c_fidx, int* __restrict__ c_fpos,
                                   int* __restrict__ n_bad) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int mi = moving_idx[k] - index_offset, fi = fixed_idx[k];
  if (mi < 0 || mi >= nm_raw) return;  // not in this shard
  const int mpos = m_inverse[mi];
  if (mpos < 0) { atomicAdd(n_bad, 1); return; }
  const int fpos = (fi >= 0 && fi < nf_raw) ? f_inverse[fi] : -1;
  c_fidx[mpos] = fi;
  c_fpos[mpos] = fpos >= 0 ? fpos : kSlotSuppressed;
}

}  // namespace s2b
