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
#include "s2b_math.cuh"
#include "../../include/srrg2b.h"
#include <cooperative_groups.h>

namespace s2b {
namespace cg = cooperative_groups;

constexpr int kAcc = 40;  // accumulator slots per slice: 21 H, 6 b, chi in/out (coarse + residual), 3 counters
constexpr int kAccB = 21, kAccChiIn = 27, kAccChiOut = 29, kAccNIn = 31, kAccNOut = 32, kAccNSup = 33;
constexpr int kMaxStats = 256;
constexpr int kMaxWindow = 64;

struct Ring {
  int window, count, head, pad_;
  double v[kMaxWindow];
};

// Everything the per-iteration solve step reads and writes, contiguous and 16-byte granular so that
// icp_solve_kernel can stage it in shared memory with one round of wide loads and write it back the
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
  int pad_;
  unsigned long long epoch;                  // peer-exchange epoch (PeerExchange), survives across runs
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
  float* S_lb;                // slice's bound-validity transform (committed every iteration)
  float cell, coord_bound;    // NN cell edge / max |coordinate| of the moving cloud
  int track2_mode;            // 0 never, 1 always, 2 automatic (small motion)
  float track2_frac;          // automatic: certify once the per-iteration motion bound is below this many cells
  int* counters;              // slice's work-list counters {far, work}: zeroed for the next iteration
};

struct alignas(16) SolveArgs {
  int dim, variable, n_slices;
  int use_tc, window, range_corr, range_inl, range_out;
  float chi_eps;
  SolveSlice sl[SRRG2B_MAX_SLICES];
};

struct SliceArgs {
  const float4* __restrict__ mp;   // moving points, Morton order: x y z | local index bits
  const float4* __restrict__ mn;   // moving normals
  int nm;
  const float4* __restrict__ fp;   // fixed points, cell order: x y z | original index bits
  const float4* __restrict__ fn;
  const int* __restrict__ cell_start;
  const unsigned* __restrict__ near_bits;  // dilated occupancy (see near_bits_kernel)
  float ox, oy, oz, inv_cell;
  float inv_cell_x;  // cells are XF times finer along x (the direction the rows run): inv_cell_x = XF * inv_cell
  int Rx;            // search radius in x cells = XF * R
  int nx, ny, nz;
  int R;      // search radius in cells (cell edge = 1.001 * max_distance / R)
  int warm;   // c_fpos holds a valid candidate position per query (previous iteration's NN)
  float md2, normal_cos;
  int gate;        // apply the normal gate
  int gate_in_nn;  // 1: the NN kernels gate and write responses (stand-alone find); 0: linearise gates
  int rob;
  float tau, ip, in_, rs;
  float fS[kKCount];  // 2^(k-22) per class: scale of the saturating fixed-point conversion (to_raw)
  float fSinvChi;     // 2^(22-k) of the coarse chi word
  const float* S;
  int* c_fpos;
  int* far_list;   // phase-2 worklist of the NN search (query positions) and its counter
  int* far_count;
  int* work_list;      // queries whose coherence check failed (need a search + a second linearise pass)
  int* work_count;
  const int* list_all; // device flag: no usable bounds -> the work list is implicitly [0, nm)
  int inline_check;    // 1: nn_kernel does the coherence check itself (stand-alone finder)
  int use_list;        // 1: nn / linearise kernels iterate over the work list
  unsigned long long* tile_stats;  // S2B_TILE_STATS builds: staged / fallback / idle tiles, cycles, sizes
  int tile;            // 1: "all" mode searches run tiled out of shared memory (nn_tile_body)
  int few_terms;       // every thread of the accumulating kernel adds at most 30 terms per slot
  float* c_lb;         // certified lower bound per query (see nn kernels)
  const float* S_lb;   // transform the bounds are valid for
  const int* track2;   // device flag: searches track the second neighbour (certify bounds)
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
// out[7] = number of valid points
__global__ void bounds_kernel(const float* __restrict__ xyz, const unsigned char* __restrict__ valid, int n,
                              int dim, int* __restrict__ out) {
  int mn[3] = {INT_MAX, INT_MAX, INT_MAX}, mx[3] = {INT_MIN, INT_MIN, INT_MIN};
  int amax = 0, cnt = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (valid && !valid[i]) continue;
    ++cnt;
    for (int a = 0; a < dim; ++a) {
      const float v = xyz[(size_t) i * dim + a];
      const int o = f2ord(v);
      mn[a] = min(mn[a], o);
      mx[a] = max(mx[a], o);
      amax = max(amax, __float_as_int(fabsf(v)));
    }
  }
  // warp reduce (integer min / max / add are REDUX instructions), then one row per warp in shared
  // memory, then ONE set of global atomics per CTA (they all hit the same eight words)
  __shared__ int red[8][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int a = 0; a < 3; ++a) {
    mn[a] = __reduce_min_sync(0xffffffffu, mn[a]);
    mx[a] = __reduce_max_sync(0xffffffffu, mx[a]);
  }
  amax = __reduce_max_sync(0xffffffffu, amax);
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if (lane == 0) {
    for (int a = 0; a < 3; ++a) { red[warp][a] = mn[a]; red[warp][3 + a] = mx[a]; }
    red[warp][6] = amax; red[warp][7] = cnt;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    const int k = threadIdx.x, nw = blockDim.x >> 5;
    int v = red[0][k];
    for (int w = 1; w < nw; ++w) {
      const int o = red[w][k];
      v = k < 3 ? min(v, o) : (k < 7 ? max(v, o) : v + o);
    }
    if (k < 3) atomicMin(&out[k], v);
    else if (k < 7) atomicMax(&out[k], v);
    else if (v) atomicAdd(&out[7], v);
  }
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
__global__ void gather_kernel(const float* __restrict__ xyz, const float* __restrict__ nrm,
                              const int* __restrict__ order, int n_valid, int dim, float4* __restrict__ op,
                              float4* __restrict__ on, int* __restrict__ inverse) {
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
  on[i] = q;
  if (inverse) inverse[src] = i;
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
  return a.use_list && !all && n_work < max(a.nm >> 6, 64);
}

__device__ __forceinline__ int slot_candidate(int slot) {
  const int c = slot >= 0 ? slot : -(slot + 2);  // -1 -> -1; kSlotSuppressed would wrap to a positive value
  return slot == kSlotSuppressed ? -1 : c;
}

// slot / bound of a finished query.  Inside the ICP loop the normal gate is evaluated by the
// lineariser (it has both normals in registers anyway); the stand-alone finder gates here.
template <int DIM>
__device__ __forceinline__ int nn_finish(const SliceArgs& a, const float* S, const NNQuery& q, int i, float lb,
                                         int old_slot) {
  int slot = q.bpos;
  if (q.bpos >= 0 && a.gate && a.gate_in_nn) {
    const float4 nm = a.mn[i];
    const float4 nf = __ldg(a.fn + q.bpos);
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
  a.c_lb[i] = lb;
  return slot;
}

// Phase 1: temporal-coherence check, else warm start + the 3^(DIM-1) rows of rings 0 and 1 (each
// lane walks only its own surviving rows).  A query whose best distance is still larger than the
// distance to ring 2 is handed to phase 2 through a worklist, so that the rare expensive queries
// (outliers, large initial misalignment) do not serialise the warps of the cheap ones.
template <int DIM, bool TRACK2>
__device__ __forceinline__ void nn_phase1_body(const SliceArgs& a, const float* S, const float* Slb, float cell,
                                               float ring2, float ring2_sq, bool all, int n_work) {
  for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < n_work; w += gridDim.x * blockDim.x) {
    const int i = all ? w : a.work_list[w];
    NNQuery q;
    const float4 m = a.mp[i];
    nn_setup<DIM>(a, S, m, q);
    const int old_slot = a.c_fpos[i];
    const int p0 = slot_candidate(old_slot);
    const float lb_old = a.inline_check ? a.c_lb[i] : 0.f;
    if (lb_old > 0.f) {
      float ox, oy, oz;
      nn_transform<DIM>(Slb, m, ox, oy, oz);
      const float ex = q.qx - ox, ey = q.qy - oy, ez = q.qz - oz;
      const float delta = __fsqrt_rn(fmaf(ez, ez, fmaf(ey, ey, ex * ex)));
      const float lbn = lb_old * (1.f - 1e-5f) - delta * (1.f + 1e-5f);
      if (lbn > 0.f) {
        if (p0 >= 0) {
          nn_consider<DIM, false>(a, q, p0);
          if (q.bpos >= 0 && q.bd2 * (1.f + 1e-5f) < lbn * lbn) {  // p0 is still the unique neighbour
            nn_finish<DIM>(a, S, q, i, lbn, old_slot);
            continue;
          }
          q.bd2 = a.md2; q.bidx = INT_MAX; q.bpos = -1;
        } else if (old_slot == -1 && lbn * lbn > a.md2 * (1.f + 1e-5f)) {  // still nothing in range
          a.c_lb[i] = lbn;
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
    {
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

// ---------------------------------------------------------------------------------------------
// Tiled search ("all" mode: every query of the slice is searched).  A CTA takes a tile of 256
// consecutive queries -- a compact blob, thanks to the Hilbert order -- and stages the WHOLE search
// neighbourhood of the tile in shared memory: the bounding box of the queries' cells dilated by R,
// i.e. for each (y, z) row of that box the contiguous run of cell-sorted fixed points and the row's
// slice of the cell table.  Every query then walks its (2R+1)^(DIM-1) rows ring by ring out of
// shared memory (~30 cycles per dependent access instead of an L2 round trip), so the far ring
// queries of a badly aligned first iteration cost the same as the near ones and no phase 2 is
// needed.  Tiles whose box does not fit (sparse regions) fall back to the global-memory walk.
// ---------------------------------------------------------------------------------------------
constexpr int kTileThreads = 256;
constexpr int kTileCapPts = 2560;      // staged fixed points (40 KB)
constexpr int kTileCapEntries = 5632;  // staged cell-table entries, (bx + 1) per row (22 KB)
constexpr int kTileCapRows = 768;      // rows of the dilated box (3 per thread in the scan)

struct TileSmem {
  float4 pts[kTileCapPts];
  int cs[kTileCapEntries];
  int rdelta[kTileCapRows];  // staged index of the row's first point minus its cell-order position
  int rows[kRowTable];
  float S[16], Slb[16];
  int box[8];       // min cx, cy, cz, max cx, cy, cz of the tile's queries
  int wsum[kTileThreads / 32];
  int npts;
  unsigned long long bar;  // mbarrier of the bulk copies (TMA variant)
};

#ifndef S2B_TILE_STATS
#define S2B_TILE_STATS 0  // 1: per-run counters of the tiled search in a.tile_stats (experiments only)
#endif
#ifndef S2B_TILE_TMA
#define S2B_TILE_TMA 1  // 1: rows staged with cp.async.bulk (TMA) + mbarrier; 0: 16-byte cp.async per point
#endif

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned) __cvta_generic_to_shared(p); }

// whole-neighbourhood walk of one query out of global memory (ring-ordered rows, pruned)
template <int DIM, bool TRACK2>
__device__ __forceinline__ void nn_search_global(const SliceArgs& a, NNQuery& q, const int* rows, int K, float cell) {
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
}

// one staged row run: candidates [ps, pe) ordered by x (up to one quantum) in shared memory.  Binary
// search for the query's x, then outwards in both directions until |x - q_x| alone exceeds the pruning
// radius (plus the quantum).
template <int DIM, bool TRACK2>
__device__ __forceinline__ void nn_scan_staged(NNQuery& q, const float4* pr, int ps, int pe, float slack) {
  int lo = ps, hi = pe;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (pr[mid].x < q.qx) lo = mid + 1; else hi = mid;
  }
#pragma unroll 1
  for (int p = lo; p < pe; ++p) {
    const float4 c = pr[p];
    const float ex = (c.x - q.qx) - slack;
    if (ex > 0.f && ex * ex > (TRACK2 ? q.sd2 : q.bd2)) break;
    nn_consider_pt<DIM, TRACK2>(q, p, c);
  }
#pragma unroll 1
  for (int p = lo - 1; p >= ps; --p) {
    const float4 c = pr[p];
    const float ex = (q.qx - c.x) - slack;
    if (ex > 0.f && ex * ex > (TRACK2 ? q.sd2 : q.bd2)) break;
    nn_consider_pt<DIM, TRACK2>(q, p, c);
  }
}

template <int DIM, bool TRACK2>
__device__ __forceinline__ void nn_tile_body(const SliceArgs& a, TileSmem& sm, float cell) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int R = a.R;
  const int K = (DIM == 3) ? (2 * R + 1) * (2 * R + 1) : (2 * R + 1);
  const int n_tiles = (a.nm + kTileThreads - 1) / kTileThreads;
#if S2B_TILE_TMA
  unsigned phase = 0;
#endif
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
#if S2B_TILE_STATS
    const long long t_begin = clock64();
#endif
    const int i = tile * kTileThreads + tid;
    const bool valid = i < a.nm;
    NNQuery q;
    int old_slot = -1;
    if (tid < 3) { sm.box[tid] = INT_MAX; sm.box[3 + tid] = INT_MIN; }
    __syncthreads();  // also fences the previous tile's readers of the staged data
    bool need = false;  // the query has to be searched (not answered by the dilated-occupancy bit)
    float cfy = 0.f, cfz = 0.f;
    if (valid) {
      nn_setup<DIM>(a, sm.S, a.mp[i], q);
      cfy = (float) q.cy + q.fry; cfz = (float) q.cz + q.frz;
      old_slot = a.c_fpos[i];
      const int p0 = slot_candidate(old_slot);
      need = true;
      if (a.warm && p0 >= 0) {
        // last iteration's neighbour: a real candidate whose distance bounds the search radius
        nn_consider<DIM, TRACK2>(a, q, p0);
        if (TRACK2) nn_limit_bound(q, cell);
      } else if (q.cx >= 0 && q.cx < a.nx && q.cy >= 0 && q.cy < a.ny && q.cz >= 0 && q.cz < a.nz) {
        need = near_bit(a.near_bits, a.nx, a.ny, q.cx, q.cy, q.cz);
      }
    }
    {  // bounding box of the cells within reach (the pruning radius) of the queries that search
      int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
      if (need) {
        const float rd = __fsqrt_rn(TRACK2 ? q.sd2 : q.bd2);
        const float rq = rd * a.inv_cell + 2e-3f, rqx = rd * a.inv_cell_x + 2e-3f;
        lo[0] = max((int) floorf(q.cfx - rqx), q.cx - a.Rx); hi[0] = min((int) floorf(q.cfx + rqx), q.cx + a.Rx);
        lo[1] = max((int) floorf(cfy - rq), q.cy - R); hi[1] = min((int) floorf(cfy + rq), q.cy + R);
        if (DIM == 3) { lo[2] = max((int) floorf(cfz - rq), q.cz - R); hi[2] = min((int) floorf(cfz + rq), q.cz + R); }
        else { lo[2] = 0; hi[2] = 0; }
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        lo[k] = __reduce_min_sync(0xffffffffu, lo[k]);
        hi[k] = __reduce_max_sync(0xffffffffu, hi[k]);
      }
      if (lane < 3) {
        const int l = lane == 0 ? lo[0] : (lane == 1 ? lo[1] : lo[2]);
        const int h = lane == 0 ? hi[0] : (lane == 1 ? hi[1] : hi[2]);
        if (l != INT_MAX) { atomicMin(&sm.box[lane], l); atomicMax(&sm.box[3 + lane], h); }
      }
    }
    __syncthreads();
    const bool any = sm.box[0] != INT_MAX;
    const int x0 = min(max(sm.box[0], 0), a.nx - 1), x1 = min(max(sm.box[3], 0), a.nx - 1);
    const int y0 = min(max(sm.box[1], 0), a.ny - 1), y1 = min(max(sm.box[4], 0), a.ny - 1);
    const int z0 = (DIM == 3) ? min(max(sm.box[2], 0), a.nz - 1) : 0;
    const int z1 = (DIM == 3) ? min(max(sm.box[5], 0), a.nz - 1) : 0;
    const int bx1 = x1 - x0 + 2, by = y1 - y0 + 1, bz = z1 - z0 + 1;  // bx1 = table entries per row
    const int nrows = by * bz;
    bool staged = any && nrows <= kTileCapRows && nrows * bx1 <= kTileCapEntries;
    if (staged) {
      // the rows' slices of the cell table, entries dealt to the threads, four independent loads in
      // flight per thread (divisions by multiply-high: exact while e * bx1 < 2^32)
      {
        const int E = nrows * bx1;
        const unsigned inv_bx1 = (unsigned) (0x100000000ull / (unsigned) bx1) + 1u;
        const unsigned inv_by = by > 1 ? (unsigned) (0x100000000ull / (unsigned) by) + 1u : 0u;
        for (int e0 = tid; e0 < E; e0 += kTileThreads * 4) {
          int v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int e = e0 + j * kTileThreads;
            if (e < E) {
              const int r = (int) __umulhi((unsigned) e, inv_bx1), k = e - r * bx1;
              const int rz = by > 1 ? (int) __umulhi((unsigned) r, inv_by) : r, ry = r - rz * by;
              v[j] = __ldg(a.cell_start + (size_t) ((z0 + rz) * a.ny + y0 + ry) * a.nx + x0 + k);
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int e = e0 + j * kTileThreads;
            if (e < E) sm.cs[e] = v[j];
          }
        }
      }
      __syncthreads();
      // exclusive scan of the rows' point counts -> staged offsets (3 consecutive rows per thread)
      int cnt[3], mine = 0;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int r = tid * 3 + j;
        cnt[j] = r < nrows ? sm.cs[r * bx1 + bx1 - 1] - sm.cs[r * bx1] : 0;
        mine += cnt[j];
      }
      int incl = mine;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += v;
      }
      if (lane == 31) sm.wsum[warp] = incl;
      __syncthreads();
      int wbase = 0, total = 0;
#pragma unroll
      for (int w = 0; w < kTileThreads / 32; ++w) {
        const int v = sm.wsum[w];
        if (w < warp) wbase += v;
        total += v;
      }
      staged = total <= kTileCapPts;  // uniform across the CTA
      if (staged) {
        int off = wbase + incl - mine;
#if S2B_TILE_TMA
        // every thread announces the bytes of its own rows before issuing their copies
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&sm.bar)),
                     "r"((unsigned) mine * 16u) : "memory");
#endif
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int r = tid * 3 + j;
          if (r < nrows) {
            const int g0 = sm.cs[r * bx1];
            sm.rdelta[r] = off - g0;
#if S2B_TILE_TMA
            if (cnt[j] > 0) {  // one bulk copy per row: contiguous in cell order, 16-byte granules
              asm volatile(
                "cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                  smem_u32(&sm.pts[off])),
                "l"(a.fp + g0), "r"((unsigned) cnt[j] * 16u), "r"(smem_u32(&sm.bar))
                : "memory");
            }
#endif
            off += cnt[j];
          }
        }
#if S2B_TILE_TMA
        {  // wait for all bytes of this tile's rows
          unsigned done = 0;
          while (!done) {
            asm volatile(
              "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
              : "=r"(done) : "r"(smem_u32(&sm.bar)), "r"(phase) : "memory");
          }
          phase ^= 1u;
        }
#else
        __syncthreads();  // rdelta visible
        for (int r = warp; r < nrows; r += kTileThreads / 32) {
          const int g0 = sm.cs[r * bx1], n = sm.cs[r * bx1 + bx1 - 1] - g0, d = sm.rdelta[r];
          for (int p = lane; p < n; p += 32)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(&sm.pts[d + g0 + p])), "l"(a.fp + g0 + p) : "memory");
        }
        asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
#endif
      }
    }
    __syncthreads();
#if S2B_TILE_STATS
    const long long t_staged = clock64();
    if (tid == 0 && a.tile_stats) {
      atomicAdd(a.tile_stats + (staged ? 0 : (any ? 1 : 2)), 1ull);
      if (!staged && any && nrows <= kTileCapRows && nrows * bx1 <= kTileCapEntries) atomicAdd(a.tile_stats + 7, 1ull);
      atomicAdd(a.tile_stats + 3, (unsigned long long) (t_staged - t_begin));
      if (staged) { atomicAdd(a.tile_stats + 5, (unsigned long long) nrows); atomicAdd(a.tile_stats + 6, (unsigned long long) (nrows * bx1)); }
    }
    struct SearchTimer {
      long long t0; unsigned long long* out; bool on;
      __device__ ~SearchTimer() { if (on) atomicAdd(out, (unsigned long long) (clock64() - t0)); }
    } search_timer{t_staged, a.tile_stats + (staged ? 4 : 8), lane == 0 && a.tile_stats != nullptr};
#endif
    if (!valid) continue;
    if (!need) {
      // nothing occupied within R cells of the query's cell: no fixed point within the covered radius
      nn_finish<DIM>(a, sm.S, q, i, TRACK2 ? __fsqrt_rn(a.rho_s2) * (1.f - 1e-5f) : 0.f, old_slot);
      continue;
    }
    if (!staged) {
      nn_search_global<DIM, TRACK2>(a, q, sm.rows, K, cell);
    } else {
      // one staged row: the cells within reach along x, then the x-sorted run
      auto scan_row = [&](int y, int z, float lb2) {
        const float pr2 = TRACK2 ? q.sd2 : q.bd2;
        const float rr = __fsqrt_rn(fmaxf(pr2 - lb2, 0.f)) * a.inv_cell_x + 2e-3f;
        const int xa = max(max((int) floorf(q.cfx - rr), q.cx - a.Rx), 0);
        const int xb = min(min((int) floorf(q.cfx + rr), q.cx + a.Rx), a.nx - 1);
        if (xa > xb) return;
        const int r = (z - z0) * by + (y - y0);
        const int* csr = sm.cs + r * bx1 - x0;
        nn_scan_staged<DIM, TRACK2>(q, sm.pts + sm.rdelta[r], csr[xa], csr[xb + 1], a.xq_slack);
      };
      // centre row first: it usually shrinks the pruning radius to a fraction of a cell ...
      if (q.cy >= 0 && q.cy < a.ny && q.cz >= 0 && q.cz < a.nz) scan_row(q.cy, q.cz, 0.f);
      // ... so that only the rows within that radius are left (each still pruned by its slab distance)
      const float rq = __fsqrt_rn(TRACK2 ? q.sd2 : q.bd2) * a.inv_cell + 2e-3f;
      const int ya = max(max((int) floorf(cfy - rq), q.cy - R), 0), yb = min(min((int) floorf(cfy + rq), q.cy + R), a.ny - 1);
      const int za = (DIM == 3) ? max(max((int) floorf(cfz - rq), q.cz - R), 0) : 0;
      const int zb = (DIM == 3) ? min(min((int) floorf(cfz + rq), q.cz + R), a.nz - 1) : 0;
      for (int z = za; z <= zb; ++z) {
        const float gz = (DIM == 3) ? axis_gap(z - q.cz, q.frz) * cell : 0.f;
        for (int y = ya; y <= yb; ++y) {
          if (y == q.cy && z == q.cz) continue;
          const float gy = axis_gap(y - q.cy, q.fry) * cell;
          const float lb2 = (DIM == 3) ? fmaf(gz, gz, gy * gy) : gy * gy;
          if (lb2 > (TRACK2 ? q.sd2 : q.bd2)) continue;
          scan_row(y, z, lb2);
        }
      }
    }
    nn_finish<DIM>(a, sm.S, q, i, TRACK2 ? __fsqrt_rn(q.sd2) * (1.f - 1e-5f) : 0.f, old_slot);
  }
}

template <int DIM>
__global__ void __launch_bounds__(kTileThreads, 3) nn_tile_kernel(const SliceArgs a) {
  if (*a.stop) return;
  extern __shared__ __align__(16) unsigned char tile_smem_raw[];
  TileSmem& sm = *reinterpret_cast<TileSmem*>(tile_smem_raw);
  const bool all = !a.use_list || *a.list_all;
  if (!all) return;
  if (threadIdx.x < 16) { sm.S[threadIdx.x] = a.S[threadIdx.x]; sm.Slb[threadIdx.x] = a.S_lb[threadIdx.x]; }
  const float cell = __fdiv_rn(1.f, a.inv_cell);
  const int K = (DIM == 3) ? (2 * a.R + 1) * (2 * a.R + 1) : (2 * a.R + 1);
  for (int k = threadIdx.x; k < K; k += blockDim.x)
    sm.rows[k] = (DIM == 3) ? *reinterpret_cast<const int*>(c_rows3[k]) : *reinterpret_cast<const int*>(c_rows2[k]);
#if S2B_TILE_TMA
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&sm.bar)), "r"(kTileThreads) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
#endif
  __syncthreads();
  if (*a.track2) nn_tile_body<DIM, true>(a, sm, cell);
  else nn_tile_body<DIM, false>(a, sm, cell);
}

template <int DIM>
__global__ void __launch_bounds__(256) nn_kernel(const SliceArgs a) {
  // the control words are fetched together: on converged iterations this kernel has nothing to do and
  // its cost is the latency of these loads
  const int stop = *a.stop, list_all = *a.list_all, work_count = *a.work_count, track2 = *a.track2;
  if (stop) return;
  const bool all = !a.use_list || list_all;
  if (a.tile && all) return;  // nn_tile_kernel searched everything
  const int n_work = all ? a.nm : work_count;
  if (small_work_list(a, all, n_work)) return;
  __shared__ float S[16], Slb[16];
  if (threadIdx.x < 16) { S[threadIdx.x] = a.S[threadIdx.x]; Slb[threadIdx.x] = a.S_lb[threadIdx.x]; }
  __syncthreads();
  const float cell = __fdiv_rn(1.f, a.inv_cell);
  // distance below which a point cannot lie in ring 2 or beyond (for R == 1: the covered radius)
  const float ring2 = (a.R >= 2) ? (1.f - 4e-3f) * cell : __fsqrt_rn(a.rho_s2);
  const float ring2_sq = (a.R >= 2) ? ring2 * ring2 : 3.0e38f;
  if (track2) nn_phase1_body<DIM, true>(a, S, Slb, cell, ring2, ring2_sq, all, n_work);
  else nn_phase1_body<DIM, false>(a, S, Slb, cell, ring2, ring2_sq, all, n_work);
}

// Phase 2: the queries phase 1 could not settle (worklist).  These are few but expensive
// (typically no neighbour at all, so nothing prunes), so one WARP takes one query: the lanes fetch
// the bounds of the (2R+1)^(DIM-1) rows, the points of all rows are dealt out to the lanes, then a
// shuffle reduction merges the lanes' (nearest, second nearest) pairs and lane 0 applies the gate
// and writes slot and bound.
template <int DIM>
struct LinAcc;
template <int DIM, int FACTOR>
__device__ __forceinline__ void lin_one(const SliceArgs& a, const float* Ss, int i, int slot, int bpos, const float4 m,
                                        const float4 nm, const float4 f, const float4 nf, LinAcc<DIM>& A);

template <int DIM, bool TRACK2, int FACTOR = SRRG2B_FACTOR_P2P>
__device__ __forceinline__ void nn_far_body(const SliceArgs& a, const float* S, const int* rows, int K, float cell,
                                            int n_far, const int* list, LinAcc<DIM>* lin, int w_first, int w_stride) {
  const int lane = threadIdx.x & 31;
  if (!lin && n_far > (a.nm >> 4)) {
    // long worklist (large initial misalignment): one THREAD per query, rows nearest ring first
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < n_far; w += gridDim.x * blockDim.x) {
      const int i = list[w];
      NNQuery q;
      nn_setup<DIM>(a, S, a.mp[i], q);
      const int old_slot = a.c_fpos[i];
      const int p0 = slot_candidate(old_slot);
      if (a.warm && p0 >= 0) { nn_consider<DIM, TRACK2>(a, q, p0); if (TRACK2) nn_limit_bound(q, cell); }
      // (far list of phase 1, nearest point only: rings 0-1 were searched exhaustively there)
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
    const int i = list[w];
    NNQuery q;
    const float4 m = a.mp[i];
    nn_setup<DIM>(a, S, m, q);
    const int old_slot = a.c_fpos[i];
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
        if (bpos >= 0) lin_one<DIM, FACTOR>(a, S, i, slot, bpos, m, a.mn[i], __ldg(a.fp + bpos), __ldg(a.fn + bpos), *lin);
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
                                       float4* __restrict__ op, float4* __restrict__ on, int* __restrict__ inverse) {
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
  on[i] = q;
  if (inverse) inverse[i] = i;
}

__global__ void __launch_bounds__(256) proj_find_kernel(const SliceArgs a) {
  if (*a.stop) return;
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

// Fixed-point conversion of one term, k fractional bits, |v| <= B = 2^(21-k):
//   t = sat(v * 2^(k-22) + 0.5)   one FFMA.SAT: maps [-B, B] onto [0, 1] and clamps everything else
//   u = t + 3.0                   in [3, 4] the fp32 spacing is 2^-22, i.e. one unit of 2^-k of v
// so bits(u) - bits(3.5f) is the term in units of 2^-k (t is rounded to a quarter unit or finer, then u
// to the unit, ties to even: the oracle performs the same two fp32 operations).  The accumulators add the
// RAW bit patterns -- integer addition wraps modulo 2^32 -- and lin_flush subtracts count * bits(3.5f).
constexpr int kFixBias = 0x40600000;  // bit pattern of 3.5f
__device__ __forceinline__ int to_raw(float v, float s) {
  return __float_as_int(__saturatef(fmaf(v, s, 0.5f)) + 3.0f);
}
// chi as (coarse, residual): the residual of the coarse rounding is exact in fp32
__device__ __forceinline__ void to_raw2(float v, float s, float inv_s, float s_lo, int& hi, int& lo) {
  const float u = __saturatef(fmaf(v, s, 0.5f)) + 3.0f;
  hi += __float_as_int(u);
  lo += to_raw(v - (u - 3.5f) * inv_s, s_lo);
}

// robustifier on chi (threshold tau): weight, robustified chi, kernelized flag
__device__ __forceinline__ bool robustify(int kind, float tau, float chi, float& w, float& rho) {
  w = 1.f;
  rho = chi;
  if (kind == SRRG2B_ROB_NONE || !(chi > tau)) return false;
  if (kind == SRRG2B_ROB_HUBER) {
    const float delta = __fsqrt_rn(tau);
    const float sc = __fsqrt_rn(chi);
    w = __fdiv_rn(delta, sc);
    rho = fmaf(2.f * delta, sc, -tau);
  } else if (kind == SRRG2B_ROB_CAUCHY) {
    const float r = __fdiv_rn(chi, tau);
    w = __fdiv_rn(1.f, 1.f + r);
    rho = (float) ((double) tau * log_det(1.0 + (double) r));
  } else {  // Saturated / Clamp
    w = 0.f;
    rho = tau;
  }
  return true;
}

// ---------------------------------------------------------------------------------------------
// k1b: per-correspondence linearisation + exact accumulation (a5)
// ---------------------------------------------------------------------------------------------
template <int DIM>
struct LinAcc {  // per-thread partial sums: |term| < 2^21 and a thread stays below 512 terms -> 32 bits
  static constexpr int P = (DIM == 3) ? 6 : 3;
  static constexpr int NH = P * (P + 1) / 2;
  int aH[NH], ab[P];
  int chi_in, chi_in_lo, chi_out, chi_out_lo;
  int n_in, n_out, n_sup;
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int k = 0; k < NH; ++k) aH[k] = 0;
#pragma unroll
    for (int k = 0; k < P; ++k) ab[k] = 0;
    chi_in = chi_in_lo = chi_out = chi_out_lo = 0;
    n_in = n_out = n_sup = 0;
  }
};

// gate + error/Jacobian rows + robust weight + fixed-point accumulation of ONE correspondence
// (moving point i with normal, fixed point at cell-order position bpos with normal)
template <int DIM, int FACTOR>
__device__ __forceinline__ void lin_one(const SliceArgs& a, const float* Ss, int i, int slot, int bpos, const float4 m,
                                        const float4 nm, const float4 f, const float4 nf, LinAcc<DIM>& A) {
  constexpr int P = (DIM == 3) ? 6 : 3;
  const float s00 = Ss[0], s01 = Ss[1], s02 = Ss[2], s03 = Ss[3];
  const float s10 = Ss[4], s11 = Ss[5], s12 = Ss[6], s13 = Ss[7];
  const float s20 = Ss[8], s21 = Ss[9], s22 = Ss[10], s23 = Ss[11];
  float t;
  t = s00 * m.x; t = fmaf(s01, m.y, t); if (DIM == 3) t = fmaf(s02, m.z, t); const float qx = t + s03;
  t = s10 * m.x; t = fmaf(s11, m.y, t); if (DIM == 3) t = fmaf(s12, m.z, t); const float qy = t + s13;
  float qz = 0.f;
  if (DIM == 3) { t = s20 * m.x; t = fmaf(s21, m.y, t); t = fmaf(s22, m.z, t); qz = t + s23; }
  t = s00 * nm.x; t = fmaf(s01, nm.y, t); if (DIM == 3) t = fmaf(s02, nm.z, t); const float nqx = t;
  t = s10 * nm.x; t = fmaf(s11, nm.y, t); if (DIM == 3) t = fmaf(s12, nm.z, t); const float nqy = t;
  float nqz = 0.f;
  if (DIM == 3) { t = s20 * nm.x; t = fmaf(s21, nm.y, t); t = fmaf(s22, nm.z, t); nqz = t; }

  if (a.gate) {  // normal gate of the finder, n_f . (R_S n_m) >= normal_cos
    float dot = fmaf(nf.y, nqy, nf.x * nqx);
    if (DIM == 3) dot = fmaf(nf.z, nqz, dot);
    const bool ok = !(dot < a.normal_cos);
    if (ok != (slot >= 0)) a.c_fpos[i] = ok ? bpos : -(bpos + 2);
    if (!ok) {
      if (a.c_stat) a.c_stat[i] = SRRG2B_STAT_NONE;
      return;
    }
  }
  // ---- error rows e, information om, Jacobian rows J (right perturbation of X) ----
  constexpr int E = (FACTOR == SRRG2B_FACTOR_P2P) ? DIM : DIM + 1;
  float e[E], om[E], J[E][P];
  const float dx = qx - f.x, dy = qy - f.y, dz = qz - f.z;
  const float rs = a.rs;
  if (DIM == 3) {
    const float R[3][3] = {{s00, s01, s02}, {s10, s11, s12}, {s20, s21, s22}};
    if (FACTOR == SRRG2B_FACTOR_P2P) {
      const float d[3] = {dx, dy, dz};
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        J[r][0] = R[r][0]; J[r][1] = R[r][1]; J[r][2] = R[r][2];
        t = R[r][2] * m.y; J[r][3] = -rs * fmaf(R[r][1], m.z, -t);
        t = R[r][0] * m.z; J[r][4] = -rs * fmaf(R[r][2], m.x, -t);
        t = R[r][1] * m.x; J[r][5] = -rs * fmaf(R[r][0], m.y, -t);
        e[r] = d[r]; om[r] = a.ip;
      }
    } else {
      float av[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        t = R[0][c] * nf.x; t = fmaf(R[1][c], nf.y, t); t = fmaf(R[2][c], nf.z, t);
        av[c] = t;
      }
      J[0][0] = av[0]; J[0][1] = av[1]; J[0][2] = av[2];
      t = m.z * av[1]; J[0][3] = rs * fmaf(m.y, av[2], -t);
      t = m.x * av[2]; J[0][4] = rs * fmaf(m.z, av[0], -t);
      t = m.y * av[0]; J[0][5] = rs * fmaf(m.x, av[1], -t);
      e[0] = fmaf(nf.z, dz, fmaf(nf.y, dy, nf.x * dx)); om[0] = a.ip;
      const float nqv[3] = {nqx, nqy, nqz}, nfv[3] = {nf.x, nf.y, nf.z};
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        J[r + 1][0] = 0.f; J[r + 1][1] = 0.f; J[r + 1][2] = 0.f;
        t = R[r][2] * nm.y; J[r + 1][3] = -rs * fmaf(R[r][1], nm.z, -t);
        t = R[r][0] * nm.z; J[r + 1][4] = -rs * fmaf(R[r][2], nm.x, -t);
        t = R[r][1] * nm.x; J[r + 1][5] = -rs * fmaf(R[r][0], nm.y, -t);
        e[r + 1] = nqv[r] - nfv[r]; om[r + 1] = a.in_;
      }
    }
  } else {
    const float R[2][2] = {{s00, s01}, {s10, s11}};
    if (FACTOR == SRRG2B_FACTOR_P2P) {
      const float d[2] = {dx, dy};
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        J[r][0] = R[r][0]; J[r][1] = R[r][1];
        t = R[r][0] * m.y; J[r][2] = fmaf(R[r][1], m.x, -t);
        e[r] = d[r]; om[r] = a.ip;
      }
    } else {
      float av[2];
#pragma unroll
      for (int c = 0; c < 2; ++c) av[c] = fmaf(R[1][c], nf.y, R[0][c] * nf.x);
      t = av[0] * m.y;
      J[0][0] = av[0]; J[0][1] = av[1]; J[0][2] = fmaf(av[1], m.x, -t);
      e[0] = fmaf(nf.y, dy, nf.x * dx); om[0] = a.ip;
      const float nqv[2] = {nqx, nqy}, nfv[2] = {nf.x, nf.y};
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        J[r + 1][0] = 0.f; J[r + 1][1] = 0.f;
        t = R[r][0] * nm.y; J[r + 1][2] = fmaf(R[r][1], nm.x, -t);
        e[r + 1] = nqv[r] - nfv[r]; om[r + 1] = a.in_;
      }
    }
  }
  float chi = (om[0] * e[0]) * e[0];
#pragma unroll
  for (int r = 1; r < E; ++r) chi = fmaf(om[r] * e[r], e[r], chi);
  if (a.c_chi) a.c_chi[i] = chi;
  if (!(chi == chi) || isinf(chi)) {
    ++A.n_sup;
    if (a.c_stat) a.c_stat[i] = SRRG2B_STAT_SUPPRESSED;
    return;
  }
  float w, rho;
  const bool kern = robustify(a.rob, a.tau, chi, w, rho);
  if (kern) { ++A.n_out; to_raw2(rho, a.fS[kKChi], a.fSinvChi, a.fS[kKChiLo], A.chi_out, A.chi_out_lo); }
  else { ++A.n_in; to_raw2(chi, a.fS[kKChi], a.fSinvChi, a.fS[kKChiLo], A.chi_in, A.chi_in_lo); }
  if (a.c_stat) a.c_stat[i] = kern ? SRRG2B_STAT_KERNELIZED : SRRG2B_STAT_INLIER;
  // ---- H += J^T (w Om) J, b += J^T (w Om) e; structural zeros of the normal rows skipped ----
  constexpr int TR = (FACTOR == SRRG2B_FACTOR_P2P) ? 0 : DIM;  // columns < TR are zero in rows >= 1
  float u[E][P];
#pragma unroll
  for (int r = 0; r < E; ++r) {
    const float s = w * om[r];
#pragma unroll
    for (int c = 0; c < P; ++c) u[r][c] = s * J[r][c];
  }
  int hslot = 0;
#pragma unroll
  for (int ii = 0; ii < P; ++ii) {
#pragma unroll
    for (int jj = ii; jj < P; ++jj) {
      float h = u[0][ii] * J[0][jj];
      if (ii >= TR && jj >= TR) {
#pragma unroll
        for (int r = 1; r < E; ++r) h = fmaf(u[r][ii], J[r][jj], h);
      }
      constexpr int T = DIM;  // columns < T: translation part of the perturbation
      const int cls = (jj < T) ? kKHtt : ((ii < T) ? kKHtr : kKHrr);
      A.aH[hslot++] += to_raw(h, a.fS[cls]);
    }
  }
#pragma unroll
  for (int ii = 0; ii < P; ++ii) {
    float g = u[0][ii] * e[0];
    if (ii >= TR) {
#pragma unroll
      for (int r = 1; r < E; ++r) g = fmaf(u[r][ii], e[r], g);
    }
    A.ab[ii] += to_raw(g, a.fS[ii < DIM ? kKBt : kKBr]);
  }
}

// exact integer block reduction of the per-thread partial sums: REDUX per slot (one when the
// per-thread sums are known to stay below 2^26, else two over 16-bit halves, which cannot overflow the
// 32-bit warp sum), the warp's 40 sums parked in lanes, one plain shared store per lane, then 40
// threads add the warps' rows and issue one global atomic per slot and CTA.
constexpr int kMaxWarps = 8;  // CTAs of the accumulating kernels have at most 256 threads
struct FlushSmem {
  long long w[kMaxWarps][kAcc];
};

template <int DIM>
__device__ __forceinline__ void lin_flush(const SliceArgs& a, const LinAcc<DIM>& A, FlushSmem& sm) {
  constexpr int P = LinAcc<DIM>::P, NH = LinAcc<DIM>::NH;
  const bool few = a.few_terms != 0;
  auto wsum = [few](int v) -> long long {
    if (few) return (long long) __reduce_add_sync(0xffffffffu, v);
    const unsigned lo = __reduce_add_sync(0xffffffffu, (unsigned) v & 0xffffu);
    const int hi = __reduce_add_sync(0xffffffffu, v >> 16);
    return ((long long) hi << 16) + (long long) lo;
  };
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // the accumulators hold raw bit patterns (see to_raw): take out count * bits(3.5f), modulo 2^32
  const int bias_in = A.n_in * kFixBias, bias_out = A.n_out * kFixBias, bias_all = bias_in + bias_out;
  long long mine0 = 0, mine1 = 0;  // lane l keeps slot l and slot 32 + l
#pragma unroll
  for (int k = 0; k < NH; ++k) {
    const long long v = wsum(A.aH[k] - bias_all);
    if (lane == k) mine0 = v;
  }
#pragma unroll
  for (int k = 0; k < P; ++k) {
    const long long v = wsum(A.ab[k] - bias_all);
    if (lane == kAccB + k) mine0 = v;
  }
  {
    const int cv[4] = {A.chi_in - bias_in, A.chi_in_lo - bias_in, A.chi_out - bias_out, A.chi_out_lo - bias_out};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long v = wsum(cv[k]);
      if (kAccChiIn + k < 32) { if (lane == kAccChiIn + k) mine0 = v; }
      else { if (lane == kAccChiIn + k - 32) mine1 = v; }
    }
    const int cn[3] = {A.n_in, A.n_out, A.n_sup};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const long long v = (long long) __reduce_add_sync(0xffffffffu, cn[k]);
      if (kAccNIn + k < 32) { if (lane == kAccNIn + k) mine0 = v; }
      else { if (lane == kAccNIn + k - 32) mine1 = v; }
    }
  }
  sm.w[warp][lane] = mine0;
  if (lane < kAcc - 32) sm.w[warp][32 + lane] = mine1;
  __syncthreads();
  if (threadIdx.x < kAcc) {
    long long v = 0;
    const int nw = blockDim.x >> 5;
    for (int w = 0; w < nw; ++w) v += sm.w[w][threadIdx.x];
    if (v) atomicAdd(&a.acc[threadIdx.x], (unsigned long long) v);
  }
}

// Phase 2 / tail kernel.  Large work lists: the far list of phase 1 (see nn_far_body).  Short work
// lists (converged iterations: a few hundred queries fail the coherence check): one warp per query
// does the whole job here -- search of all rows, slot + certified bound, and the linearisation of
// that query -- because at this size the thread-per-query kernels would be pure load latency.
template <int DIM, int FACTOR>
__global__ void __launch_bounds__(256) nn_far_kernel(const SliceArgs a) {
  const int stop = *a.stop, list_all = *a.list_all, work_count = *a.work_count, far_count = *a.far_count;
  const int track2_flag = *a.track2;
  if (stop) return;
  const bool all = !a.use_list || list_all;
  const int n_work = all ? a.nm : work_count;
  const bool tail = small_work_list(a, all, n_work);
  const int n_far = tail ? n_work : far_count;
  if (n_far == 0) return;
  __shared__ float S[16];
  __shared__ int rows[kRowTable];
  __shared__ FlushSmem fsm;
  if (threadIdx.x < 16) S[threadIdx.x] = a.S[threadIdx.x];
  const int R = a.R;
  const int K = (DIM == 3) ? (2 * R + 1) * (2 * R + 1) : (2 * R + 1);
  for (int k = threadIdx.x; k < K; k += blockDim.x)
    rows[k] = (DIM == 3) ? *reinterpret_cast<const int*>(c_rows3[k]) : *reinterpret_cast<const int*>(c_rows2[k]);
  __syncthreads();
  const float cell = __fdiv_rn(1.f, a.inv_cell);
  const bool track2 = track2_flag != 0;
  const int wpb = blockDim.x >> 5, w0 = blockIdx.x * wpb + (threadIdx.x >> 5), ws = gridDim.x * wpb;
  if (!tail) {
    if (track2) nn_far_body<DIM, true>(a, S, rows, K, cell, n_far, a.far_list, nullptr, w0, ws);
    else nn_far_body<DIM, false>(a, S, rows, K, cell, n_far, a.far_list, nullptr, w0, ws);
    return;
  }
  LinAcc<DIM> A;
  A.clear();
  if (track2) nn_far_body<DIM, true, FACTOR>(a, S, rows, K, cell, n_far, a.work_list, &A, w0, ws);
  else nn_far_body<DIM, false, FACTOR>(a, S, rows, K, cell, n_far, a.work_list, &A, w0, ws);
  lin_flush<DIM>(a, A, fsm);
}

// CHECK = true is the first pass of an iteration once bounds exist: the temporal-coherence test is
// fused with the linearisation (one read of the query, its neighbour and both normals serves both);
// queries that fail it are appended to the work list for nn_kernel / nn_far_kernel, and the
// CHECK = false pass then linearises exactly those.
#ifndef S2B_LIN_THREADS
#define S2B_LIN_THREADS 256  // CTA size / resident CTAs per SM of the linearise kernels (register budget)
#define S2B_LIN_CTAS 2
#endif
constexpr int kLinThreads = S2B_LIN_THREADS, kLinCtas = S2B_LIN_CTAS;
#ifndef S2B_LIN_STAGES
#define S2B_LIN_STAGES 4
#endif
constexpr int kLinStages = S2B_LIN_STAGES;  // depth of the per-thread cp.async ring of point data
constexpr size_t kLinSmemBytes = (size_t) kLinStages * kLinThreads * (4 * sizeof(float4) + 3 * sizeof(int));
constexpr int kFailCap = 128;  // coherence-check failures a CTA of the fused kernel resolves in place

template <int DIM, int FACTOR, bool CHECK>
__global__ void __launch_bounds__(kLinThreads, kLinCtas) linearize_kernel(const SliceArgs a) {
  const int stop = *a.stop, list_all = *a.list_all, work_count = *a.work_count;
  if (stop) return;
  const bool all = !a.use_list || list_all;
  if (CHECK && all) return;  // nothing is certified: everything goes through the search path
  if (!CHECK && small_work_list(a, all, all ? a.nm : work_count)) return;  // nn_far_kernel linearises short lists itself
  __shared__ float Ss[16], Sl[16];
  __shared__ FlushSmem fsm;
  __shared__ int s_fail[CHECK ? kFailCap : 1];
  __shared__ int s_nfail;
  __shared__ int s_rows[CHECK ? kRowTable : 1];
  if (threadIdx.x < 16) { Ss[threadIdx.x] = a.S[threadIdx.x]; Sl[threadIdx.x] = CHECK ? a.S_lb[threadIdx.x] : 0.f; }
  if (threadIdx.x == 0) s_nfail = 0;
  constexpr int KMAX = (DIM == 3) ? kRowTable : (2 * kMaxR + 1);
  if (CHECK) {
    for (int k = threadIdx.x; k < KMAX; k += blockDim.x)
      s_rows[k] = (DIM == 3) ? *reinterpret_cast<const int*>(c_rows3[k]) : *reinterpret_cast<const int*>(c_rows2[k]);
  }
  __syncthreads();
  LinAcc<DIM> A;
  A.clear();

  // software pipeline: index (w + D + 1 strides) -> slot / bound (w + D strides) -> point data (w + D - 1
  // strides .. w), the point data travelling through a per-thread ring in shared memory filled by
  // 16-byte cp.async copies, so D - 1 gathers per thread are in flight while one correspondence is
  // linearised and none of them occupies registers
  constexpr int D = kLinStages;
  extern __shared__ __align__(16) unsigned char lin_smem_raw[];
  float4* ring = reinterpret_cast<float4*>(lin_smem_raw);                          // [D][4][threads]
  int* ring_i = reinterpret_cast<int*>(lin_smem_raw + (size_t) D * 4 * kLinThreads * sizeof(float4));  // [D][3][threads]
  const int stride = gridDim.x * blockDim.x;
  const int n_work = CHECK ? a.nm : (all ? a.nm : work_count);
  const bool direct = CHECK || all;
  auto index_of = [&](int w) { return w < n_work ? (direct ? w : a.work_list[w]) : -1; };
  const bool regate = a.gate != 0;  // gated-out slots are re-checked every iteration
  const int tid = threadIdx.x;
  // issue the copies of one element into ring stage st (always commits a group, possibly empty)
  auto issue = [&](int st, int i, int slot, float lb) {
    const int pn = regate ? slot_candidate(slot) : slot;
    ring_i[(st * 3 + 0) * kLinThreads + tid] = i;
    ring_i[(st * 3 + 1) * kLinThreads + tid] = slot;
    ring_i[(st * 3 + 2) * kLinThreads + tid] = __float_as_int(lb);
    float4* dst = ring + (size_t) st * 4 * kLinThreads + tid;
    if (i >= 0 && (pn >= 0 || CHECK))
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(a.mp + i) : "memory");
    if (pn >= 0) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + kLinThreads)), "l"(a.mn + i) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + 2 * kLinThreads)), "l"(a.fp + pn) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + 3 * kLinThreads)), "l"(a.fn + pn) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  // prologue: stages 0 .. D-2 hold elements w .. w + (D-2) strides
#pragma unroll
  for (int d = 0; d < D - 1; ++d) {
    const int i = index_of(w + d * stride);
    issue(d, i, i >= 0 ? a.c_fpos[i] : -1, (CHECK && i >= 0) ? a.c_lb[i] : 0.f);
  }
  // index, slot and bound of the next two elements to issue travel in registers (two iterations of
  // lead: one is not enough to cover a DRAM round trip under load), the index of the third one too
  int i_b = index_of(w + (D - 1) * stride);
  int slot_b = i_b >= 0 ? a.c_fpos[i_b] : -1;
  float lb_b = (CHECK && i_b >= 0) ? a.c_lb[i_b] : 0.f;
  int i_c = index_of(w + D * stride);
  int slot_c = i_c >= 0 ? a.c_fpos[i_c] : -1;
  float lb_c = (CHECK && i_c >= 0) ? a.c_lb[i_c] : 0.f;
  int i_a = index_of(w + (D + 1) * stride);
  int st = 0;                                // ring stage of the element processed now
  for (; w < n_work; w += stride) {
    // refill the stage freed by the previous iteration, advance the register part of the pipeline
    issue(st == 0 ? D - 1 : st - 1, i_b, slot_b, lb_b);
    i_b = i_c; slot_b = slot_c; lb_b = lb_c;
    i_c = i_a;
    slot_c = i_c >= 0 ? a.c_fpos[i_c] : -1;
    lb_c = (CHECK && i_c >= 0) ? a.c_lb[i_c] : 0.f;
    i_a = index_of(w + (D + 2) * stride);
    asm volatile("cp.async.wait_group %0;" ::"n"(D - 1) : "memory");  // the oldest group (this element) has landed
    const int i = ring_i[(st * 3 + 0) * kLinThreads + tid];
    const int slot = ring_i[(st * 3 + 1) * kLinThreads + tid];
    const float lb_old = __int_as_float(ring_i[(st * 3 + 2) * kLinThreads + tid]);
    const int bpos = regate ? slot_candidate(slot) : slot;
    const float4* src = ring + (size_t) st * 4 * kLinThreads + tid;
    const float4 m = src[0], nm = src[kLinThreads], f = src[2 * kLinThreads], nf = src[3 * kLinThreads];
    st = st == D - 1 ? 0 : st + 1;
    if (CHECK) {
      // exact temporal coherence (see the NN kernels): keep the neighbour / the "none" verdict when
      // the certified bound minus the query's motion still proves it; else hand over to the search
      bool keep = false, none = false;
      float lbn = 0.f;
      if (lb_old > 0.f) {
        float qx, qy, qz, ox, oy, oz;
        nn_transform<DIM>(Ss, m, qx, qy, qz);
        nn_transform<DIM>(Sl, m, ox, oy, oz);
        const float ex = qx - ox, ey = qy - oy, ez = qz - oz;
        const float delta = __fsqrt_rn(fmaf(ez, ez, fmaf(ey, ey, ex * ex)));
        lbn = lb_old * (1.f - 1e-5f) - delta * (1.f + 1e-5f);
        if (lbn > 0.f) {
          if (bpos >= 0) {
            const float ddx = qx - f.x, ddy = qy - f.y, ddz = qz - f.z;
            float d2 = fmaf(ddy, ddy, ddx * ddx);
            if (DIM == 3) d2 = fmaf(ddz, ddz, d2);
            keep = d2 <= a.md2 && d2 * (1.f + 1e-5f) < lbn * lbn;
          } else if (slot == -1) {
            none = lbn * lbn > a.md2 * (1.f + 1e-5f);
          }
        }
      }
      if (keep || none) {
        a.c_lb[i] = lbn;
        if (none) {
          if (a.c_stat) a.c_stat[i] = SRRG2B_STAT_NONE;
          continue;
        }
      } else {
        // the first kFailCap failures of the CTA are searched and linearised by its own warps after
        // the loop (converged iterations: a handful per CTA); the rest go to the global work list
        const int k = atomicAdd(&s_nfail, 1);
        if (k < kFailCap) {
          s_fail[k] = i;
        } else {
          cg::coalesced_group g = cg::coalesced_threads();
          int base = 0;
          if (g.thread_rank() == 0) base = atomicAdd(a.work_count, (int) g.size());
          base = g.shfl(base, 0);
          a.work_list[base + g.thread_rank()] = i;
        }
        continue;
      }
    }
    if (bpos < 0) {
      if (slot == kSlotSuppressed) {
        ++A.n_sup;
        if (a.c_stat) a.c_stat[i] = SRRG2B_STAT_SUPPRESSED;
      } else if (a.c_stat) {
        a.c_stat[i] = SRRG2B_STAT_NONE;
      }
      continue;
    }
    lin_one<DIM, FACTOR>(a, Ss, i, slot, bpos, m, nm, f, nf, A);
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  if (CHECK) {
    __syncthreads();
    const int n_local = min(s_nfail, kFailCap);
    if (n_local > 0) {  // one warp per failed query: search, slot + bound, linearisation
      const int K = (DIM == 3) ? (2 * a.R + 1) * (2 * a.R + 1) : (2 * a.R + 1);
      const float cell = __fdiv_rn(1.f, a.inv_cell);
      if (*a.track2) nn_far_body<DIM, true, FACTOR>(a, Ss, s_rows, K, cell, n_local, s_fail, &A, threadIdx.x >> 5, blockDim.x >> 5);
      else nn_far_body<DIM, false, FACTOR>(a, Ss, s_rows, K, cell, n_local, s_fail, &A, threadIdx.x >> 5, blockDim.x >> 5);
    }
  }
  lin_flush<DIM>(a, A, fsm);
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
                                int keep_stats) {
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
  for (int s = 0; s < a.n_slices; ++s) {
    compose(a.sl[s].ris, st->X, st->S[s]);
    for (int k = 0; k < kAcc; ++k) st->acc[s][k] = 0ull;
    st->ncorr[s] = 0;
    // the motion from the previous call's pose is unknown: certify bounds only when forced to
    if (!keep_stats) st->track2[s] = (a.sl[s].track2_mode == 1) ? 1 : 0;
    // a fresh compute() starts from an arbitrary guess: search everything; the inlier-only second
    // run continues from certified bounds
    if (!keep_stats) st->list_all[s] = 1;  // (the last solve step already set it for a continued run)
    if (a.sl[s].kind == SRRG2B_SLICE_POINTS && a.sl[s].counters) { a.sl[s].counters[0] = 0; a.sl[s].counters[1] = 0; }
  }
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

// set only the finder transform of one slice (stand-alone find / linearise entry points)
__global__ void set_S_kernel(DevState* st, int slice, Mat4f S, int track2) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    st->S[slice] = S;
    st->track2[slice] = track2;
    st->list_all[slice] = 1;
    for (int k = 0; k < kAcc; ++k) st->acc[slice][k] = 0ull;
    st->stop = 0;
  }
}

// body of one _runSolver iteration after the per-slice kernels
// (R/registration/aligners/multi_aligner_impl.cpp:106-126), serial part: runs on one thread against the
// shared-memory copy of the state; the IterationStats entry goes straight to global memory
template <int DIM>
__device__ void icp_solve_serial(const SolveArgs& a, DevHeader& st, srrg2b_iter_stats* stats_out) {
  constexpr int P = (DIM == 3) ? 6 : 3;
  double H[P * P], b[P];
#pragma unroll
  for (int i = 0; i < P * P; ++i) H[i] = 0.0;
#pragma unroll
  for (int i = 0; i < P; ++i) b[i] = 0.0;
  srrg2b_iter_stats s;
  s.iteration = st.n_stats;
  s.solver_status = 0;
  s.num_inliers = 0; s.num_outliers = 0; s.num_suppressed = 0; s.num_correspondences = 0;
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
        const double v = __ll2double_rn((long long) acc[slot++]) * sl.invk[(j < DIM) ? kKHtt : ((i < DIM) ? kKHtr : kKHrr)];
        H[i * P + j] = H[i * P + j] + v;
        if (j != i) H[j * P + i] = H[j * P + i] + v;
      }
    }
#pragma unroll
    for (int i = 0; i < P; ++i)
      b[i] = b[i] + __ll2double_rn((long long) acc[kAccB + i]) * sl.invk[(i < DIM) ? kKBt : kKBr];
    const long long ni = (long long) acc[kAccNIn], no = (long long) acc[kAccNOut], ns = (long long) acc[kAccNSup];
    s.num_inliers += ni; s.num_outliers += no; s.num_suppressed += ns; s.num_correspondences += ni + no + ns;
    s.chi_inliers += __ll2double_rn((long long) acc[kAccChiIn]) * sl.invk[kKChi] +
                     __ll2double_rn((long long) acc[kAccChiIn + 1]) * sl.invk[kKChiLo];
    s.chi_outliers += __ll2double_rn((long long) acc[kAccChiOut]) * sl.invk[kKChi] +
                      __ll2double_rn((long long) acc[kAccChiOut + 1]) * sl.invk[kKChiLo];
    const long long n = ni + no + ns;
    st.ncorr[k] = n;
    total += n;
    good = good || (n > (long long) sl.min_corr);  // aligner_slice_processor_impl.cpp:77-79
  }
  st.iterations_run += 1;
  if (!good) {  // multi_aligner_impl.cpp:107-111 (estimate already equals the backup)
    st.not_enough_corr = 1;
    st.stop = 1;
    return;
  }
  double dx[6] = {0, 0, 0, 0, 0, 0};
  if (spd_solve_t<P>(H, b, dx)) {
    box_plus(DIM, a.variable, dx, X);
    st.X = X;
    s.solver_status = 1;
  }
  if (st.n_stats < kMaxStats) stats_out[st.n_stats] = s;
  st.n_stats += 1;
  for (int k = 0; k < a.n_slices; ++k) {
    Mat4f Sn;
    compose(a.sl[k].ris, X, Sn);
    if (a.sl[k].kind == SRRG2B_SLICE_POINTS) {
      // upper bound of how far any query of the slice moves between this iteration and the next;
      // when it is small against the cell edge the next NN pass certifies bounds (track2) so that
      // later iterations can skip their searches
      float dr = 0.f, dt = 0.f;
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) {
          const float d = Sn.m[r * 4 + c] - st.S[k].m[r * 4 + c];
          dr += d * d;
        }
        const float d = Sn.m[r * 4 + 3] - st.S[k].m[r * 4 + 3];
        dt += d * d;
      }
      const float motion = sqrtf(dr) * 1.7321f * a.sl[k].coord_bound + sqrtf(dt);
      const int mode = a.sl[k].track2_mode;
      st.list_all[k] = st.track2[k] ? 0 : 1;  // bounds exist only if the pass just done certified them
      st.track2[k] = (mode == 1) || (mode == 2 && motion < a.sl[k].track2_frac * a.sl[k].cell) ? 1 : 0;
    }
    st.S[k] = Sn;
  }
  if (a.use_tc && has_to_stop(&st, a, s, total)) st.stop = 1;
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
struct PeerExchange {
  int rank, world;
  unsigned long long* mail[kMaxRanks];       // mail[r]: rank r's mailbox (own or IPC-mapped), 2 * kMailWords
};

__device__ __forceinline__ void st_volatile_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_volatile_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

template <int DIM>
__global__ void __launch_bounds__(kSolveThreads) icp_solve_kernel(const SolveArgs* ap, DevState* st, const PeerExchange* px) {
  __shared__ __align__(16) SolveArgs a;
  __shared__ __align__(16) DevHeader sh;
  __shared__ PeerExchange pe;
  // one round of independent 16-byte loads stages the arguments and the whole mutable state
  {
    const int4* s0 = reinterpret_cast<const int4*>(ap);
    int4* d0 = reinterpret_cast<int4*>(&a);
    for (int k = threadIdx.x; k < (int) (sizeof(SolveArgs) / 16); k += kSolveThreads) d0[k] = s0[k];
    const int4* s1 = reinterpret_cast<const int4*>(static_cast<const DevHeader*>(st));
    int4* d1 = reinterpret_cast<int4*>(&sh);
    for (int k = threadIdx.x; k < (int) (sizeof(DevHeader) / 16); k += kSolveThreads) d1[k] = s1[k];
    if (px) {
      const int* s2 = reinterpret_cast<const int*>(px);
      int* d2 = reinterpret_cast<int*>(&pe);
      for (int k = threadIdx.x; k < (int) (sizeof(PeerExchange) / sizeof(int)); k += kSolveThreads) d2[k] = s2[k];
    }
  }
  __syncthreads();
  if (sh.stop) return;  // (every rank holds the same state, so every rank returns here or none does)
  if (px) {
    // all-reduce of the accumulators over peer memory (see PeerExchange)
    const unsigned long long e = sh.epoch + 1ull;
    const unsigned long long tag = (e & 0xffffffffull) << 32;
    const int n_words = a.n_slices * kAcc;
    unsigned long long* mine = pe.mail[pe.rank] + (e & 1ull) * kMailWords;
    for (int k = threadIdx.x; k < n_words; k += kSolveThreads) {
      const unsigned long long v = (&sh.acc[0][0])[k];
      st_volatile_sys(mine + 2 * k, (v & 0xffffffffull) | tag);
      st_volatile_sys(mine + 2 * k + 1, (v >> 32) | tag);
    }
    // (peer, word) pairs are dealt to the threads: both halves are requested together and polled until
    // their tags are e, then added to the shared copy.  Everything a thread waits for is in flight at
    // once, so the exchange costs about one NVLink round trip per ceil(pairs / threads).
    __syncthreads();
    const int n_pairs = n_words * (pe.world - 1);
    for (int t = threadIdx.x; t < n_pairs; t += kSolveThreads) {
      int r = t / n_words;
      const int k = t - r * n_words;
      if (r >= pe.rank) ++r;  // peers in rank order, skipping this rank
      const unsigned long long* theirs = pe.mail[r] + (e & 1ull) * kMailWords + 2 * k;
      unsigned long long lo, hi;
      do {
        lo = ld_volatile_sys(theirs);
        hi = ld_volatile_sys(theirs + 1);
      } while ((lo & 0xffffffff00000000ull) != tag || (hi & 0xffffffff00000000ull) != tag);
      atomicAdd(&sh.acc[0][0] + k, (lo & 0xffffffffull) | ((hi & 0xffffffffull) << 32));
    }
    __syncthreads();
    if (threadIdx.x == 0) sh.epoch = e;  // (written back with the rest of the state)
  }
  // the NN pass of this iteration certified its bounds at S: record that before anything can bail out;
  // the work-list counters start the next iteration at zero
  for (int k = 0; k < a.n_slices; ++k) {
    if (a.sl[k].kind != SRRG2B_SLICE_POINTS) continue;
    if (a.sl[k].S_lb && threadIdx.x < 16) a.sl[k].S_lb[threadIdx.x] = sh.S[k].m[threadIdx.x];
    if (a.sl[k].counters && threadIdx.x >= 32 && threadIdx.x < 34) a.sl[k].counters[threadIdx.x - 32] = 0;
  }
  if (threadIdx.x == 0) icp_solve_serial<DIM>(a, sh, st->stats);
  __syncthreads();
  // accumulators restart at zero; everything else goes back as the serial part left it
  for (int k = threadIdx.x; k < a.n_slices * kAcc; k += kSolveThreads) sh.acc[k / kAcc][k % kAcc] = 0ull;
  __syncthreads();
  {
    const int4* s1 = reinterpret_cast<const int4*>(&sh);
    int4* d1 = reinterpret_cast<int4*>(static_cast<DevHeader*>(st));
    for (int k = threadIdx.x; k < (int) (sizeof(DevHeader) / 16); k += kSolveThreads) d1[k] = s1[k];
  }
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
