// s2b_api.cu -- host side of libsrrg2b.so: context, cloud residency, index builds, launch
// sequencing of the device-driven ICP loop, NCCL plumbing, and the extern "C" boundary declared in
// include/srrg2b.h.  No CPU fallback: without a CUDA device every compute call returns
// SRRG2B_ERR_CUDA.
#include "s2b_tiles.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/reverse_iterator.h>
#include <cuda/functional>
#include <dlfcn.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include <algorithm>
#include <array>

using namespace s2b;

namespace {

// ---- minimal NCCL binding (resolved at run time from the NCCL already loaded in the process) ----
struct NcclId { char internal[128]; };
typedef void* ncclComm_t;
struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, NcclId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool load() {
    if (handle) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so", nullptr};
    const char* env = getenv("SRRG2B_NCCL_LIB");
    if (env) handle = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    for (int i = 0; !handle && names[i]; ++i) handle = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!handle) return false;
    GetUniqueId = (int (*)(NcclId*)) dlsym(handle, "ncclGetUniqueId");
    CommInitRank = (int (*)(ncclComm_t*, int, NcclId, int)) dlsym(handle, "ncclCommInitRank");
    AllReduce = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t)) dlsym(handle, "ncclAllReduce");
    AllGather = (int (*)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t)) dlsym(handle, "ncclAllGather");
    CommDestroy = (int (*)(ncclComm_t)) dlsym(handle, "ncclCommDestroy");
    GetErrorString = (const char* (*) (int) ) dlsym(handle, "ncclGetErrorString");
    return GetUniqueId && CommInitRank && AllReduce && CommDestroy;
  }
};
NcclApi g_nccl;
constexpr int kNcclUint64 = 5, kNcclFloat32 = 7, kNcclSum = 0, kNcclMax = 2;

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  // borrowed: the memory belongs to another context (srrg2b_share_fixed) and is read-only here.  ensure() announces
  // a write, so it lets go of a borrowed buffer and allocates a private one.
  bool borrowed = false;
  cudaError_t ensure(size_t n) {
    if (borrowed) { p = nullptr; cap = 0; borrowed = false; }
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc((void**) &p, sizeof(T) * (n ? n : 1));
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void release() {
    if (p && !borrowed) cudaFree(p);
    p = nullptr;
    cap = 0;
    borrowed = false;
  }
  void borrow(const DevBuf<T>& o) {
    release();
    p = o.p;
    cap = o.cap;
    borrowed = o.p != nullptr;
  }
};

// a device buffer that keeps its first `used` elements when it grows (the resident scene is appended to)
template <typename T>
cudaError_t grow_preserve(DevBuf<T>& b, size_t used, size_t need, cudaStream_t st) {
  if (need <= b.cap) return cudaSuccess;
  const size_t cap = std::max(need, b.cap + b.cap / 2);
  T* np_ = nullptr;
  cudaError_t e = cudaMalloc((void**) &np_, sizeof(T) * (cap ? cap : 1));
  if (e != cudaSuccess) return e;
  if (b.p && used) e = cudaMemcpyAsync(np_, b.p, sizeof(T) * used, cudaMemcpyDeviceToDevice, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (b.p) cudaFree(b.p);
  b.p = np_;
  b.cap = cap;
  return e;
}

struct RawCloud {
  DevBuf<float> xyz, nrm;
  DevBuf<unsigned char> valid;
  int64_t n = 0, n_global = 0, index_offset = 0;
  bool has_normals = false, has_valid = false, present = false;
  // max |n|^2 over the valid points (0 without normals): computed on the device right after the normals
  // landed, read back lazily (nb2_pending) through the pinned word h_nb2
  float nb2 = 0.f;
  bool nb2_pending = false;
  int* h_nb2 = nullptr;
  int* d_nb2 = nullptr;
  // upload progress on the copy stream: coordinates (+ validity) landed / everything landed.  The index
  // builds start on the coordinates while the normals are still crossing PCIe.
  cudaEvent_t ev_coords = nullptr, ev_all = nullptr;
};

struct SliceData {
  RawCloud fixed_raw, moving_raw;
  // fixed index (lazy, keyed by the cell size it was built for)
  DevBuf<float4> f_pts, f_rec;  // compact points (searches) / 32-byte {point | normal} records (lineariser gathers)
  DevBuf<int> f_inverse, cell_start;
  DevBuf<unsigned> near_bits;
  DevBuf<float> clip_xyz, clip_nrm;  // N1: the clipped scene this slice's moving cloud was built from
  DevBuf<int> clip_gidx;             // ... and the indices of its points in the full scene
  int64_t n_clipped = 0;
  int clip_scene_id = -1;
  int nf_valid = 0;
  float built_for_max_distance = -1.f;
  float ox = 0, oy = 0, oz = 0, inv_cell = 1;
  int nx = 1, ny = 1, nz = 1;
  int R = 1;  // cells per max_distance
  int xbits = 0;  // low key bits of the cell-order sort: x inside the cell (cell_key_kernel)
  int nx_coarse = 1;
  int xf = 1;     // the grid is xf times finer along x: row runs are cut by the pruning radius at that resolution
  // resolution chosen by the last build (skips the occupancy read-back while the cloud size is stable)
  int cached_R = 0, cached_R_n = 0;
  float cached_R_md = -1.f;
  float last_max_distance = -1.f;  // finder radius of the last NN index: set_cloud(FIXED) rebuilds for it right away
  // projective index (alternative to the grid): index image + SoA in original order
  bool index_is_projective = false;
  srrg2b_finder_params proj_params = {};
  DevBuf<unsigned long long> image;
  // moving, Morton order
  DevBuf<float4> m_pts, m_nrm, m_pair;
  DevBuf<int> m_inverse;
  int nm_valid = 0;
  float coord_bound = 0.f;
  float radius2 = 0.f;      // max |m|^2 over the valid moving points (all shards once bounds_global)
  float nb2_moving_global = 0.f;
  bool bounds_global = false;
  // correspondences in moving-sorted order
  DevBuf<int> c_fidx, c_fpos, far_list, far_count, work_list;
  DevBuf<float> c_resp, c_chi, c_lb, S_lb;
  DevBuf<unsigned char> c_stat;
  bool corr_valid = false, stat_valid = false;
  int prune_on_export = 0;
  bool have_last_S = false;   // stand-alone finds: last_S is the anchor transform of the slice's certified bounds
  s2b::Mat4f last_S;
};

}  // namespace

struct srrg2b_ctx {
  int dim = 3, device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // cloud uploads: they overlap the index build queued on `stream`
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::string err;
  int64_t launches = 0;
  std::map<int, SliceData> slices;
  std::map<int, RawCloud> scenes;  // N1: local maps resident in HBM between frames (srrg2b_scene_set / _clip)
  DevState* d_state = nullptr;
  DevState* h_state = nullptr;  // pinned mirror (header part is copied back)
  // scratch
  DevBuf<unsigned> keys_a, keys_b;
  DevBuf<int> vals_a, vals_b, flags, positions, bounds;
  DevBuf<unsigned char> cub_tmp;
  DevBuf<int> o_fidx, o_midx, d_fidx;
  DevBuf<float> o_resp, d_resp, o_chi, d_chi;
  DevBuf<unsigned char> o_stat, d_stat;
  DevBuf<int> imp_f, imp_m, imp_bad;
  int* h_bounds = nullptr;  // pinned 8 ints
  int* h_bounds_init = nullptr;  // pinned: initial value of the bounds reduction
  // comm
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  // peer-memory exchange of the accumulators (PeerExchange in s2b_icp.cuh); null: NCCL all-reduce
  s2b::PeerExchange* d_px = nullptr;
  unsigned long long* d_mail = nullptr;
  unsigned long long* d_epoch = nullptr;
  std::vector<void*> peer_maps;
  float last_ms = 0.f;
  int last_iterations = 0;
  int sm_count = 148;
  // optional per-launch timing of the slice kernel (roofline measurement)
  int track2_mode = 2;  // 0 never, 1 always, 2 automatic; env SRRG2B_TRACK2 overrides
  float track2_frac = 1.0f;  // env SRRG2B_TRACK2_FRAC overrides
  bool time_kernels = false;
  std::vector<cudaEvent_t> kev;
  size_t kev_used = 0;
  float last_kernel_ms = 0.f;
  int last_kernel_launches = 0;
  // the launch sequence of a whole run (init + all iterations) is captured once per distinct plan
  // into a CUDA graph and replayed (0.6 us per kernel node instead of 2.9 us per stream launch)
  struct RunGraph {
    std::vector<unsigned char> key;
    cudaGraphExec_t exec = nullptr;
    int64_t launches = 0;
  };
  std::vector<RunGraph> run_graphs;
  bool use_graphs = true;  // env SRRG2B_NO_GRAPH=1 disables
  bool graph_nccl = false;  // env SRRG2B_GRAPH_NCCL=1: capture the all-reduce too (experimental: failed the 2-GPU parity test)
  bool eager_index = true;  // env SRRG2B_EAGER_INDEX=0: build the NN index on first use only
  int small_shift = 6;  // env SRRG2B_SMALL_SHIFT
  int far_sole_ctas = 2;  // CTAs per SM of the lone search kernel of a late iteration (env SRRG2B_FAR_SOLE_CTAS)
  int nn_flat = 3;  // thread-per-query searches walk their rows in chunks of 8 (env SRRG2B_NN_FLAT=0: one row at a time)
  int full_iters = 4;  // iterations of a run that launch the three-kernel search pipeline (env SRRG2B_FULL_ITERS)
  long long timeout_cycles = 4000000000ll;  // ~2 s of SM clock: the peer exchange gives up (env SRRG2B_TIMEOUT_MS)
  s2b::Mat4f* d_T0 = nullptr;  // initial guess of the current run (read by icp_init_kernel)
  s2b::Mat4f* h_T0 = nullptr;  // pinned staging
  s2b::SolveArgs* d_solve = nullptr;  // solve-step arguments of the current run
  s2b::SolveArgs* h_solve = nullptr;
};

namespace {

#define CK(ctx, call)                                                                         \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                       \
      return SRRG2B_ERR_CUDA;                                                                 \
    }                                                                                         \
  } while (0)

#define FAIL(ctx, code, msg) \
  do {                       \
    (ctx)->err = (msg);      \
    return (code);           \
  } while (0)

inline int blocks_for(int64_t n, int threads) { return (int) ((n + threads - 1) / threads); }

// SRRG2B_TRACE=1: wall-clock checkpoints of the index builds on stderr (each one synchronises the stream)
struct Trace {
  srrg2b_ctx* c;
  const char* what;
  bool on;
  std::chrono::steady_clock::time_point t0;
  Trace(srrg2b_ctx* ctx, const char* w);
  void mark(const char* label);
};

Trace::Trace(srrg2b_ctx* ctx, const char* w) : c(ctx), what(w), on(getenv("SRRG2B_TRACE") != nullptr) {
  if (on) { cudaStreamSynchronize(c->stream); t0 = std::chrono::steady_clock::now(); }
}
void Trace::mark(const char* label) {
  if (!on) return;
  cudaStreamSynchronize(c->stream);
  const auto t1 = std::chrono::steady_clock::now();
  fprintf(stderr, "[srrg2b trace] %s: %s %.1f us\n", what, label, std::chrono::duration<double, std::micro>(t1 - t0).count());
  t0 = std::chrono::steady_clock::now();
}

int cub_sort_pairs(srrg2b_ctx* c, int n, int end_bit) {
  size_t bytes = 0;
  CK(c, cub::DeviceRadixSort::SortPairs(nullptr, bytes, c->keys_a.p, c->keys_b.p, c->vals_a.p, c->vals_b.p, n, 0,
                                        end_bit, c->stream));
  CK(c, c->cub_tmp.ensure(bytes));
  CK(c, cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, bytes, c->keys_a.p, c->keys_b.p, c->vals_a.p, c->vals_b.p, n, 0,
                                        end_bit, c->stream));
  return SRRG2B_OK;
}

int compute_bounds(srrg2b_ctx* c, const RawCloud& rc) {
  CK(c, c->bounds.ensure(kBoundWords));
  // (pinned source: a pageable one makes the copy synchronous behind the cloud uploads queued before it)
  CK(c, cudaMemcpyAsync(c->bounds.p, c->h_bounds_init, kBoundWords * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  if (rc.n > 0) {
    const int blocks = std::min(blocks_for(rc.n, 256), c->sm_count * 8);
    bounds_kernel<<<blocks, 256, 0, c->stream>>>(rc.xyz.p, rc.has_valid ? rc.valid.p : nullptr, (int) rc.n, c->dim,
                                                 c->bounds.p);
    c->launches++;
  }
  CK(c, cudaMemcpyAsync(c->h_bounds, c->bounds.p, kBoundWords * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  return SRRG2B_OK;
}

// max |n|^2 of a cloud's normals: queued on the compute stream behind the upload, read back lazily
int queue_normal_bound(srrg2b_ctx* c, RawCloud& rc) {
  rc.nb2 = 0.f;
  rc.nb2_pending = false;
  if (!rc.has_normals || rc.n == 0) return SRRG2B_OK;
  if (!rc.d_nb2) CK(c, cudaMalloc((void**) &rc.d_nb2, sizeof(int)));
  if (!rc.h_nb2) CK(c, cudaMallocHost((void**) &rc.h_nb2, sizeof(int)));
  if (rc.ev_all) CK(c, cudaStreamWaitEvent(c->stream, rc.ev_all, 0));
  CK(c, cudaMemsetAsync(rc.d_nb2, 0, sizeof(int), c->stream));
  const int blocks = std::min(blocks_for(rc.n, 256), c->sm_count * 8);
  normal_bound_kernel<<<blocks, 256, 0, c->stream>>>(rc.nrm.p, rc.has_valid ? rc.valid.p : nullptr, (int) rc.n, c->dim, rc.d_nb2);
  c->launches++;
  CK(c, cudaMemcpyAsync(rc.h_nb2, rc.d_nb2, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  rc.nb2_pending = true;
  return SRRG2B_OK;
}

int resolve_normal_bound(srrg2b_ctx* c, RawCloud& rc) {
  if (!rc.nb2_pending) return SRRG2B_OK;
  CK(c, cudaStreamSynchronize(c->stream));
  memcpy(&rc.nb2, rc.h_nb2, 4);
  rc.nb2_pending = false;
  return SRRG2B_OK;
}

// Queues the copies of a cloud on the copy stream -- validity and coordinates first, then the normals --
// and records the two progress events.  Does NOT wait: the caller's buffers are borrowed until the copy
// stream (or the compute stream after it waited for ev_all) has been synchronised.
int upload_raw(srrg2b_ctx* c, RawCloud& rc, const srrg2b_cloud* cl) {
  const int dim = c->dim;
  const cudaMemcpyKind kind = cl->on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  rc.n = cl->n;
  rc.n_global = cl->n_global > 0 ? cl->n_global : cl->n;
  rc.index_offset = cl->index_offset;
  rc.has_normals = cl->normals != nullptr;
  rc.has_valid = cl->valid != nullptr;
  if (!rc.ev_coords) CK(c, cudaEventCreateWithFlags(&rc.ev_coords, cudaEventDisableTiming));
  if (!rc.ev_all) CK(c, cudaEventCreateWithFlags(&rc.ev_all, cudaEventDisableTiming));
  CK(c, rc.xyz.ensure((size_t) cl->n * dim));
  if (rc.has_normals) CK(c, rc.nrm.ensure((size_t) cl->n * dim));
  if (rc.has_valid) {
    CK(c, rc.valid.ensure((size_t) cl->n));
    if (cl->n) CK(c, cudaMemcpyAsync(rc.valid.p, cl->valid, cl->n, kind, c->copy_stream));
  }
  if (cl->n) CK(c, cudaMemcpyAsync(rc.xyz.p, cl->coords, sizeof(float) * cl->n * dim, kind, c->copy_stream));
  CK(c, cudaEventRecord(rc.ev_coords, c->copy_stream));
  if (rc.has_normals && cl->n)
    CK(c, cudaMemcpyAsync(rc.nrm.p, cl->normals, sizeof(float) * cl->n * dim, kind, c->copy_stream));
  CK(c, cudaEventRecord(rc.ev_all, c->copy_stream));
  rc.present = true;
  return SRRG2B_OK;
}

// moving cloud: Hilbert order over its own bounding box, float4 SoA, inverse permutation
int build_moving(srrg2b_ctx* c, SliceData& sd) {
  RawCloud& rc = sd.moving_raw;
  const int n = (int) rc.n, dim = c->dim;
  Trace tr(c, "build_moving");
  if (rc.ev_coords) CK(c, cudaStreamWaitEvent(c->stream, rc.ev_coords, 0));
  int rcode = compute_bounds(c, rc);
  if (rcode) return rcode;
  tr.mark("bounds");
  sd.nm_valid = c->h_bounds[7];
  memcpy(&sd.coord_bound, &c->h_bounds[6], 4);
  memcpy(&sd.radius2, &c->h_bounds[8], 4);
  sd.bounds_global = false;
  // (+ 4: the streaming lineariser moves these arrays with bulk copies in multiples of 4 elements)
  CK(c, sd.m_pts.ensure((size_t) n + 4));
  CK(c, sd.m_nrm.ensure((size_t) n + 4));
  CK(c, sd.m_pair.ensure(((size_t) n / 2 + 2) * 3));
  CK(c, sd.m_inverse.ensure((size_t) n));
  CK(c, sd.c_fidx.ensure((size_t) n));
  CK(c, sd.c_fpos.ensure((size_t) n + 4));
  CK(c, sd.far_list.ensure((size_t) n));
  CK(c, sd.far_count.ensure(4));  // [0] far worklist size, [1] coherence worklist size, [2] tile ticket
  CK(c, sd.work_list.ensure((size_t) n));
  CK(c, sd.c_lb.ensure((size_t) n + 4));
  CK(c, sd.S_lb.ensure(kSlbFloats));
  if (n) CK(c, cudaMemsetAsync(sd.c_lb.p, 0, sizeof(float) * (size_t) n, c->stream));
  CK(c, cudaMemsetAsync(sd.S_lb.p, 0, sizeof(float) * kSlbFloats, c->stream));
  CK(c, sd.c_resp.ensure((size_t) n));
  CK(c, sd.c_chi.ensure((size_t) n));
  CK(c, sd.c_stat.ensure((size_t) n));
  sd.corr_valid = false;
  sd.stat_valid = false;
  sd.have_last_S = false;
  tr.mark("buffers");
  if (n == 0) return SRRG2B_OK;
  float mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
  for (int a = 0; a < dim; ++a) {
    mn[a] = ord2f(c->h_bounds[a]);
    mx[a] = ord2f(c->h_bounds[3 + a]);
  }
  const float q = dim == 3 ? 1023.f : 32767.f;
  float sc[3];
  for (int a = 0; a < 3; ++a) {
    const float ext = mx[a] - mn[a];
    sc[a] = (sd.nm_valid > 0 && ext > 0.f) ? q / ext : 0.f;
  }
  CK(c, c->keys_a.ensure(n));
  CK(c, c->keys_b.ensure(n));
  CK(c, c->vals_a.ensure(n));
  CK(c, c->vals_b.ensure(n));
  curve_key_kernel<<<blocks_for(n, 256), 256, 0, c->stream>>>(rc.xyz.p, rc.has_valid ? rc.valid.p : nullptr, n, dim,
                                                               mn[0], mn[1], mn[2], sc[0], sc[1], sc[2], c->keys_a.p,
                                                               c->vals_a.p);
  c->launches++;
  tr.mark("keys");
  rcode = cub_sort_pairs(c, n, 32);
  if (rcode) return rcode;
  tr.mark("sort");
  fill_int_kernel<<<blocks_for(n, 256), 256, 0, c->stream>>>(sd.m_inverse.p, n, -1);
  c->launches++;
  if (sd.nm_valid > 0) {
    if (rc.ev_all) CK(c, cudaStreamWaitEvent(c->stream, rc.ev_all, 0));
    gather_kernel<<<blocks_for(sd.nm_valid, 256), 256, 0, c->stream>>>(
      rc.xyz.p, rc.has_normals ? rc.nrm.p : nullptr, c->vals_b.p, sd.nm_valid, dim, sd.m_pts.p, sd.m_nrm.p,
      sd.m_inverse.p, nullptr);
    c->launches++;
    pair_pack_kernel<<<blocks_for((sd.nm_valid + 1) / 2, 256), 256, 0, c->stream>>>(sd.m_pts.p, sd.m_nrm.p, sd.nm_valid, sd.m_pair.p);
    fill_int_kernel<<<blocks_for(sd.nm_valid, 256), 256, 0, c->stream>>>(sd.c_fidx.p, sd.nm_valid, -1);
    fill_int_kernel<<<blocks_for(sd.nm_valid, 256), 256, 0, c->stream>>>(sd.c_fpos.p, sd.nm_valid, -1);
    c->launches += 3;
  }
  CK(c, cudaGetLastError());
  tr.mark("gather");
  return queue_normal_bound(c, rc);
}

// The cell edge is kCellSlack * max_distance / R: the (2R+1)^dim neighbourhood of a query's cell is
// then guaranteed to cover a radius rho_s ~ 1.046 * max_distance, a little MORE than the accept
// radius, so a search that finds nothing certifies 'no point within rho_s' and the verdict survives
// small motions (temporal coherence, see s2b_icp.cuh).
constexpr float kCellSlack = 1.05f;

// k0: uniform grid over the fixed cloud with cell edge >= max_distance (so the 3^dim
// neighbourhood of a query's cell contains every point within max_distance), cell-sorted float4 SoA
int ensure_index(srrg2b_ctx* c, SliceData& sd, float max_distance) {
  RawCloud& rc = sd.fixed_raw;
  if (!rc.present) FAIL(c, SRRG2B_ERR_STATE, "fixed cloud not set for slice");
  if (!(max_distance > 0.f)) FAIL(c, SRRG2B_ERR_INVALID, "max_distance must be > 0");
  if (sd.built_for_max_distance == max_distance) return SRRG2B_OK;
  const int n = (int) rc.n, dim = c->dim;
  Trace tr(c, "ensure_index");
  if (rc.ev_coords) CK(c, cudaStreamWaitEvent(c->stream, rc.ev_coords, 0));
  int rcode = compute_bounds(c, rc);
  if (rcode) return rcode;
  tr.mark("bounds");
  sd.nf_valid = c->h_bounds[7];
  float mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
  if (sd.nf_valid > 0) {
    for (int a = 0; a < dim; ++a) {
      mn[a] = ord2f(c->h_bounds[a]);
      mx[a] = ord2f(c->h_bounds[3 + a]);
    }
  }
  CK(c, sd.f_pts.ensure((size_t) n + 1));
  CK(c, sd.f_rec.ensure(2 * ((size_t) n + 1)));
  CK(c, cudaMemsetAsync(sd.f_rec.p, 0, 2 * sizeof(float4), c->stream));  // (position 0 stands in for "no candidate")
  CK(c, sd.f_inverse.ensure((size_t) n + 1));
  if (n > 0) {
    CK(c, c->keys_a.ensure(n));
    CK(c, c->keys_b.ensure(n));
    CK(c, c->vals_a.ensure(n));
    CK(c, c->vals_b.ensure(n));
  }
  // Cell edge = 1.001 * max_distance / R: the 0.1% slack absorbs the fp32 rounding of the cell
  // coordinate, and the (2R+1)^dim neighbourhood of a query's cell holds every point within
  // max_distance.  R is the finest of 4..1 that fits the table budget and keeps about two or more
  // points per occupied cell (finer cells = fewer candidates examined per query).
  int R = kMaxR;
  if (const char* env = getenv("SRRG2B_GRID_R")) R = std::max(1, std::min(kMaxR, atoi(env)));
  const bool forced = getenv("SRRG2B_GRID_R") != nullptr;
  // same finder radius and about the same number of points as the last build of this slice: start at
  // (and keep) its resolution; a frame-to-frame stream of scans then builds without a host round trip
  // after the bounds
  const bool cached = !forced && sd.cached_R > 0 && sd.cached_R_md == max_distance &&
                      std::abs(sd.nf_valid - sd.cached_R_n) <= sd.cached_R_n / 8;
  if (cached) R = sd.cached_R;
  int dims[3] = {1, 1, 1};
  float cell = max_distance;
  for (;; --R) {
    cell = max_distance * kCellSlack / (float) R;
    double total = 1.0;
    bool fits = true;
    for (int a = 0; a < dim; ++a) {
      const double cnt = floor((double) (mx[a] - mn[a]) / (double) cell) + 1.0;
      if (cnt > 1024.0) fits = false;
      dims[a] = (int) (cnt > 1024.0 ? 1024 : cnt);
      total *= cnt;
    }
    if (!fits || total > 8.0 * 1024 * 1024) {
      if (R > 1) continue;
      // even R = 1 does not fit: grow the cell (the search then covers more than it needs to)
      while (true) {
        cell *= 2.f;
        total = 1.0;
        fits = true;
        for (int a = 0; a < dim; ++a) {
          const double cnt = floor((double) (mx[a] - mn[a]) / (double) cell) + 1.0;
          if (cnt > 1024.0) fits = false;
          dims[a] = (int) (cnt > 1024.0 ? 1024 : cnt);
          total *= cnt;
        }
        if (fits && total <= 8.0 * 1024 * 1024) break;
      }
    }
    sd.ox = mn[0]; sd.oy = mn[1]; sd.oz = mn[2];
    sd.inv_cell = 1.f / cell;
    sd.nx = dims[0]; sd.ny = dims[1]; sd.nz = dims[2];
    sd.R = R;
    {
      // Rows run along x, and a row scan costs two table loads whatever its length, so the grid is made
      // up to 4x finer along x only: the pruning radius then cuts a run at a quarter of the cell edge
      // (4x fewer candidates on a converged query) at the price of a 4x larger table, nothing else.
      int xf = 4;
      if (const char* env = getenv("SRRG2B_GRID_XF")) xf = std::max(1, std::min(4, atoi(env)));
      double total = (double) sd.nx * sd.ny * sd.nz;
      while (xf > 1 && ((double) sd.nx * xf > 4096.0 || total * xf > 32.0 * 1024 * 1024)) xf >>= 1;
      if (xf == 3) xf = 2;
      sd.xf = xf;
      sd.nx_coarse = sd.nx;
      if (xf > 1) sd.nx = std::min(sd.nx * xf, (int) floor((double) (mx[0] - mn[0]) * xf / (double) cell) + 1);
    }
    {  // the key bits the cell id leaves free order the points of a cell by x (at most 16 bits)
      int cell_bits = 1;
      while (((int64_t) 1 << cell_bits) < (int64_t) sd.nx * sd.ny * sd.nz) ++cell_bits;
      sd.xbits = std::max(0, std::min(16, 32 - cell_bits));
    }
    if (n > 0) {
      cell_key_kernel<<<blocks_for(n, 256), 256, 0, c->stream>>>(rc.xyz.p, rc.has_valid ? rc.valid.p : nullptr, n,
                                                                 dim, sd.ox, sd.oy, sd.oz, sd.inv_cell,
                                                                 sd.inv_cell * (float) sd.xf, sd.nx, sd.ny, sd.nz,
                                                                 sd.xbits, c->keys_a.p, c->vals_a.p);
      c->launches++;
      rcode = cub_sort_pairs(c, n, 32);
      if (rcode) return rcode;
    }
    if (R == 1 || forced || sd.nf_valid == 0 || cached) break;
    CK(c, cudaMemsetAsync(c->bounds.p, 0, 4, c->stream));
    count_distinct_kernel<<<std::min(blocks_for(sd.nf_valid, 256), c->sm_count * 8), 256, 0, c->stream>>>(
      c->keys_b.p, sd.nf_valid, sd.xbits, sd.nx, sd.xf, c->bounds.p);
    c->launches++;
    CK(c, cudaMemcpyAsync(c->h_bounds, c->bounds.p, 4, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    const double occupancy = (double) sd.nf_valid / (double) std::max(1, c->h_bounds[0]);
    if (occupancy >= 2.0) break;
  }
  tr.mark("keys + sort + occupancy");
  const int ncells = sd.nx * sd.ny * sd.nz;
  CK(c, sd.cell_start.ensure((size_t) ncells + 1));
  if (n > 0) {
    fill_int_kernel<<<blocks_for(n, 256), 256, 0, c->stream>>>(sd.f_inverse.p, n, -1);
    c->launches++;
    if (sd.nf_valid > 0) {
      if (rc.ev_all) CK(c, cudaStreamWaitEvent(c->stream, rc.ev_all, 0));
      gather_kernel<<<blocks_for(sd.nf_valid, 256), 256, 0, c->stream>>>(
        rc.xyz.p, rc.has_normals ? rc.nrm.p : nullptr, c->vals_b.p, sd.nf_valid, dim, sd.f_pts.p, nullptr,
        sd.f_inverse.p, sd.f_rec.p);
      c->launches++;
    }
  }
  tr.mark("gather");
  {
    // run heads of the sorted cell ids, then a reverse running minimum fills the empty cells
    CK(c, c->vals_a.ensure((size_t) ncells + 1));
    fill_int_kernel<<<blocks_for(ncells + 1, 256), 256, 0, c->stream>>>(c->vals_a.p, ncells + 1, sd.nf_valid);
    c->launches++;
    if (sd.nf_valid > 0) {
      cell_head_kernel<<<blocks_for(sd.nf_valid, 256), 256, 0, c->stream>>>(c->keys_b.p, sd.nf_valid, sd.xbits, c->vals_a.p);
      c->launches++;
    }
    auto rin = thrust::make_reverse_iterator(c->vals_a.p + ncells + 1);
    auto rout = thrust::make_reverse_iterator(sd.cell_start.p + ncells + 1);
    size_t bytes = 0;
    CK(c, cub::DeviceScan::InclusiveScan(nullptr, bytes, rin, rout, ::cuda::minimum<>{}, ncells + 1, c->stream));
    CK(c, c->cub_tmp.ensure(bytes));
    CK(c, cub::DeviceScan::InclusiveScan(c->cub_tmp.p, bytes, rin, rout, ::cuda::minimum<>{}, ncells + 1, c->stream));
  }
  tr.mark("cell table");
  {
    const int nxw = near_words_per_row(sd.nx), nrows = sd.ny * sd.nz;
    const size_t words = (size_t) nrows * nxw;
    CK(c, sd.near_bits.ensure(words));
    CK(c, c->keys_a.ensure(words));
    near_bits_x_kernel<<<blocks_for((int64_t) words * 32, 256), 256, 0, c->stream>>>(sd.cell_start.p, sd.nx, nrows, sd.R * sd.xf,
                                                                     c->keys_a.p);
    near_bits_yz_kernel<<<blocks_for(words, 256), 256, 0, c->stream>>>(c->keys_a.p, sd.nx, sd.ny, sd.nz, sd.R, dim,
                                                                      sd.near_bits.p);
    c->launches += 2;
  }
  // positions into the old ordering are meaningless now: drop the warm-start candidates
  if (sd.moving_raw.present && sd.nm_valid > 0) {
    fill_int_kernel<<<blocks_for(sd.nm_valid, 256), 256, 0, c->stream>>>(sd.c_fpos.p, sd.nm_valid, -1);
    fill_int_kernel<<<blocks_for(sd.nm_valid, 256), 256, 0, c->stream>>>(sd.c_fidx.p, sd.nm_valid, -1);
    c->launches += 2;
    CK(c, cudaMemsetAsync(sd.c_lb.p, 0, sizeof(float) * (size_t) sd.nm_valid, c->stream));
    sd.corr_valid = false;
    sd.have_last_S = false;
  }
  CK(c, cudaGetLastError());
  tr.mark("near bits + resets");
  sd.built_for_max_distance = max_distance;
  sd.last_max_distance = max_distance;
  sd.cached_R = sd.R; sd.cached_R_n = sd.nf_valid; sd.cached_R_md = max_distance;
  return SRRG2B_OK;
}

// k0p: projective index = index image of the fixed cloud + float4 SoA in original order
int ensure_proj_index(srrg2b_ctx* c, SliceData& sd, const srrg2b_finder_params& fp) {
  RawCloud& rc = sd.fixed_raw;
  if (!rc.present) FAIL(c, SRRG2B_ERR_STATE, "fixed cloud not set for slice");
  if (c->dim != 3) FAIL(c, SRRG2B_ERR_INVALID, "the projective finder needs dim == 3");
  if (fp.width <= 0 || fp.height <= 0 || !(fp.fx > 0.f) || !(fp.fy > 0.f) || !(fp.max_distance > 0.f))
    FAIL(c, SRRG2B_ERR_INVALID, "bad projective finder parameters");
  const srrg2b_finder_params& o = sd.proj_params;
  if (sd.index_is_projective && sd.built_for_max_distance > 0.f && o.fx == fp.fx && o.fy == fp.fy && o.cx == fp.cx &&
      o.cy == fp.cy && o.width == fp.width && o.height == fp.height && o.min_depth == fp.min_depth &&
      o.max_depth == fp.max_depth)
    return SRRG2B_OK;
  const int n = (int) rc.n;
  const size_t npx = (size_t) fp.width * fp.height;
  CK(c, sd.f_pts.ensure((size_t) n + 1));
  CK(c, sd.f_rec.ensure(2 * ((size_t) n + 1)));
  CK(c, cudaMemsetAsync(sd.f_rec.p, 0, 2 * sizeof(float4), c->stream));
  CK(c, sd.f_inverse.ensure((size_t) n + 1));
  CK(c, sd.image.ensure(npx));
  CK(c, cudaMemsetAsync(sd.image.p, 0xff, npx * sizeof(unsigned long long), c->stream));
  if (n > 0) {
    if (rc.ev_all) CK(c, cudaStreamWaitEvent(c->stream, rc.ev_all, 0));
    gather_identity_kernel<<<blocks_for(n, 256), 256, 0, c->stream>>>(rc.xyz.p, rc.has_normals ? rc.nrm.p : nullptr, n,
                                                                      c->dim, sd.f_pts.p, sd.f_rec.p, sd.f_inverse.p);
    proj_image_kernel<<<blocks_for(n, 256), 256, 0, c->stream>>>(rc.xyz.p, rc.has_valid ? rc.valid.p : nullptr, n, fp.fx,
                                                                 fp.fy, fp.cx, fp.cy, fp.min_depth, fp.max_depth,
                                                                 fp.width, fp.height, sd.image.p);
    c->launches += 2;
  }
  sd.nf_valid = n;
  sd.R = 1; sd.xf = 1; sd.nx = sd.ny = sd.nz = 1; sd.inv_cell = 1.f / fp.max_distance;
  if (sd.moving_raw.present && sd.nm_valid > 0) {
    fill_int_kernel<<<blocks_for(sd.nm_valid, 256), 256, 0, c->stream>>>(sd.c_fpos.p, sd.nm_valid, -1);
    fill_int_kernel<<<blocks_for(sd.nm_valid, 256), 256, 0, c->stream>>>(sd.c_fidx.p, sd.nm_valid, -1);
    c->launches += 2;
    CK(c, cudaMemsetAsync(sd.c_lb.p, 0, sizeof(float) * (size_t) sd.nm_valid, c->stream));
    sd.corr_valid = false;
    sd.have_last_S = false;
  }
  CK(c, cudaGetLastError());
  sd.index_is_projective = true;
  sd.proj_params = fp;
  sd.built_for_max_distance = fp.max_distance;
  return SRRG2B_OK;
}

int ensure_any_index(srrg2b_ctx* c, SliceData& sd, const srrg2b_finder_params& fp) {
  if (fp.kind == SRRG2B_FINDER_PROJECTIVE) return ensure_proj_index(c, sd, fp);
  if (fp.kind != SRRG2B_FINDER_NN) FAIL(c, SRRG2B_ERR_INVALID, "unknown finder kind");
  if (sd.index_is_projective) {  // switching back to the grid: force a rebuild
    sd.index_is_projective = false;
    sd.built_for_max_distance = -1.f;
  }
  return ensure_index(c, sd, fp.max_distance);
}

// The fixed-point ranges come from global quantities: max |m|^2 and max |n|^2 over ALL shards of the
// moving cloud (and the replicated fixed cloud).  One small max all-reduce per upload.
int ensure_global_bound(srrg2b_ctx* c, SliceData& sd) {
  int rcode = resolve_normal_bound(c, sd.moving_raw);
  if (rcode) return rcode;
  rcode = resolve_normal_bound(c, sd.fixed_raw);
  if (rcode) return rcode;
  if (sd.bounds_global) return SRRG2B_OK;
  sd.nb2_moving_global = sd.moving_raw.nb2;
  if (c->world > 1) {
    CK(c, c->bounds.ensure(kBoundWords));
    const float two[2] = {sd.radius2, sd.moving_raw.nb2};
    CK(c, cudaMemcpyAsync(c->bounds.p, two, 8, cudaMemcpyHostToDevice, c->stream));
    if (g_nccl.AllReduce(c->bounds.p, c->bounds.p, 2, kNcclFloat32, kNcclMax, c->comm, c->stream) != 0)
      FAIL(c, SRRG2B_ERR_NCCL, "ncclAllReduce(max) failed");
    float out[2] = {0.f, 0.f};
    CK(c, cudaMemcpyAsync(out, c->bounds.p, 8, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    sd.radius2 = out[0];
    sd.nb2_moving_global = out[1];
  }
  sd.bounds_global = true;
  return SRRG2B_OK;
}

int fill_slice_args(srrg2b_ctx* c, SliceData& sd, int state_slot, const srrg2b_finder_params& fp,
                    const srrg2b_factor_params& fa, int variable, bool want_status, SliceArgs& a, Scales* sc_out) {
  if (!sd.moving_raw.present) FAIL(c, SRRG2B_ERR_STATE, "moving cloud not set for slice");
  int rcode = ensure_any_index(c, sd, fp);
  if (rcode) return rcode;
  rcode = ensure_global_bound(c, sd);
  if (rcode) return rcode;
  const bool normals = sd.fixed_raw.has_normals && sd.moving_raw.has_normals;
  if (fa.factor == SRRG2B_FACTOR_PLANE && !normals)
    FAIL(c, SRRG2B_ERR_INVALID, "PLANE factor needs normals on both clouds");
  if (!(fa.info_point >= 0.f) || !(fa.info_normal >= 0.f) || !(fa.chi_threshold >= 0.f))
    FAIL(c, SRRG2B_ERR_INVALID, "informations and the robustifier threshold must be non-negative");
  const float nb2 = std::max(sd.nb2_moving_global, sd.fixed_raw.nb2);
  const Scales sc = choose_scales(c->dim, variable, fa.factor, sd.radius2, nb2, fp.max_distance, fa.info_point, fa.info_normal);
  if (sc_out) *sc_out = sc;
  a.mp = sd.m_pts.p; a.mn = sd.m_nrm.p; a.mpair = sd.m_pair.p; a.nm = sd.nm_valid;
  a.fp = sd.f_pts.p; a.frec = sd.f_rec.p; a.cell_start = sd.cell_start.p; a.near_bits = sd.near_bits.p;
  a.ox = sd.ox; a.oy = sd.oy; a.oz = sd.oz; a.inv_cell = sd.inv_cell;
  a.inv_cell_x = sd.inv_cell * (float) sd.xf; a.Rx = sd.R * sd.xf;
  a.nx = sd.nx; a.ny = sd.ny; a.nz = sd.nz;
  a.R = sd.R;
  a.warm = 1;
  a.md2 = fp.max_distance * fp.max_distance;
  a.normal_cos = fp.normal_cos;
  a.gate = (normals && fp.normal_cos > -1.f) ? 1 : 0;
  a.rob = fa.robustifier; a.tau = fa.chi_threshold; a.delta = sqrtf(fa.chi_threshold);
  a.ip = fa.info_point; a.in_ = fa.info_normal;
  a.rs = (c->dim == 3 && variable == SRRG2B_VAR_SE3_QUAT_RIGHT) ? 2.f : 1.f;
  a.eb2 = sc.err_bound * sc.err_bound;
  for (int k = 0; k < kKCount; ++k) a.fS[k] = ldexpf(1.f, sc.k[k] - 22);
  a.fSinvChi = ldexpf(1.f, 22 - sc.k[kKChi]);
  a.S = c->d_state->S[state_slot].m;
  a.c_fpos = sd.c_fpos.p;
  a.gate_in_nn = 0;
  a.far_list = sd.far_list.p; a.far_count = sd.far_count.p;
  a.work_list = sd.work_list.p; a.work_count = sd.far_count.p + 1; a.tile_ticket = sd.far_count.p + 2;
  a.list_all = &c->d_state->list_all[state_slot];
  a.inline_check = 0; a.use_list = 0; a.sole_list = 0;
  a.few_terms = 0;
  a.nn_flat = c->nn_flat;
  a.small_shift = c->small_shift;
  a.projective = fp.kind == SRRG2B_FINDER_PROJECTIVE ? 1 : 0;
  a.fx = fp.fx; a.fy = fp.fy; a.pcx = fp.cx; a.pcy = fp.cy; a.min_depth = fp.min_depth; a.max_depth = fp.max_depth;
  a.width = fp.width; a.height = fp.height; a.image = sd.image.p;
  a.c_lb = sd.c_lb.p; a.S_lb = sd.S_lb.p;
  a.track2 = &c->d_state->track2[state_slot];
  a.radius = sqrtf(sd.radius2) * 1.0001f;
  a.xq_slack = sd.index_is_projective ? 0.f : (1.01f * ldexpf(1.f, -sd.xbits) + 1e-4f) / sd.inv_cell;  // one key quantum + fp32 rounding of the cell coordinate
  {
    const float rho = ((float) sd.R - 4e-3f) / sd.inv_cell;
    a.rho_s2 = rho * rho;
    if (a.rho_s2 < a.md2) a.rho_s2 = a.md2;
  }
  a.c_stat = want_status ? sd.c_stat.p : nullptr;
  a.c_chi = want_status ? sd.c_chi.p : nullptr;
  a.acc = c->d_state->acc[state_slot];
  a.stop = &c->d_state->stop;
  return SRRG2B_OK;
}

// skip: device flag that turns the launch into a no-op (the persistent loop has taken over)
void launch_far(srrg2b_ctx* c, const SliceArgs& a_in, int factor, const int* skip) {
  const int threads = 256;
  SliceArgs a = a_in;
  const int fblocks = std::max(1, std::min(blocks_for((int64_t) a.nm * 32, threads), c->sm_count * ((a.sole_list && a.nm <= (1 << 21)) ? c->far_sole_ctas : 8)));  // (big clouds: their late work lists are not short)
  {  // tail mode: lane 0 of a warp linearises its share of a short (< nm / 64 + 64) work list
    const int64_t warps = (int64_t) fblocks * (threads / 32);
    const int64_t per_lane = (((int64_t) a.nm >> a.small_shift) + 64 + warps - 1) / warps;
    a.few_terms = per_lane <= 30 ? 1 : 0;
  }
  if (c->dim == 3) {
    if (factor == SRRG2B_FACTOR_P2P) nn_far_kernel<3, SRRG2B_FACTOR_P2P><<<fblocks, threads, 0, c->stream>>>(a, skip);
    else nn_far_kernel<3, SRRG2B_FACTOR_PLANE><<<fblocks, threads, 0, c->stream>>>(a, skip);
  } else {
    if (factor == SRRG2B_FACTOR_P2P) nn_far_kernel<2, SRRG2B_FACTOR_P2P><<<fblocks, threads, 0, c->stream>>>(a, skip);
    else nn_far_kernel<2, SRRG2B_FACTOR_PLANE><<<fblocks, threads, 0, c->stream>>>(a, skip);
  }
  c->launches++;
}

// phase 1 of the grid search (rings 0-1, thread per query)
void launch_nn(srrg2b_ctx* c, const SliceArgs& a, const int* skip) {
  const int blocks = std::max(1, std::min(blocks_for(a.nm, 256), c->sm_count * 8));
  if (c->dim == 3) nn_kernel<3><<<blocks, 256, 0, c->stream>>>(a, skip);
  else nn_kernel<2><<<blocks, 256, 0, c->stream>>>(a, skip);
  c->launches++;
}

int launch_find(srrg2b_ctx* c, const SliceArgs& a, const int* skip) {
  if (a.nm <= 0) return SRRG2B_OK;
  const int threads = 256;
  const int blocks = std::max(1, std::min(blocks_for(a.nm, threads), c->sm_count * 8));
  if (a.projective) {
    proj_find_kernel<<<blocks, threads, 0, c->stream>>>(a, skip);
    c->launches++;
    return SRRG2B_OK;
  }
  launch_nn(c, a, skip);
  if (a.R >= 2) launch_far(c, a, SRRG2B_FACTOR_P2P, skip);
  return SRRG2B_OK;
}

inline int lin_grid(const srrg2b_ctx* c, int n) { return std::max(1, std::min(blocks_for(n, kWTile * kLoopWarps), c->sm_count)); }

// streaming lineariser over every slot of the slice as it is
int launch_linearize(srrg2b_ctx* c, const SliceArgs& a_in, int factor, const int* skip) {
  if (a_in.nm <= 0) return SRRG2B_OK;
  SliceArgs a = a_in;
  const int blocks = lin_grid(c, a.nm);
  if (c->dim == 3) {
    if (factor == SRRG2B_FACTOR_P2P) lin_tiles_kernel<3, SRRG2B_FACTOR_P2P><<<blocks, kLoopThreads, kLoopSmemBytes, c->stream>>>(a, skip);
    else lin_tiles_kernel<3, SRRG2B_FACTOR_PLANE><<<blocks, kLoopThreads, kLoopSmemBytes, c->stream>>>(a, skip);
  } else {
    if (factor == SRRG2B_FACTOR_P2P) lin_tiles_kernel<2, SRRG2B_FACTOR_P2P><<<blocks, kLoopThreads, kLoopSmemBytes, c->stream>>>(a, skip);
    else lin_tiles_kernel<2, SRRG2B_FACTOR_PLANE><<<blocks, kLoopThreads, kLoopSmemBytes, c->stream>>>(a, skip);
  }
  c->launches++;
  return SRRG2B_OK;
}

template <typename K3P, typename K3L, typename K2P, typename K2L>
void launch_tiles_kernel(srrg2b_ctx* c, const SliceArgs& a, int factor, K3P k3p, K3L k3l, K2P k2p, K2L k2l) {
  const int blocks = lin_grid(c, a.nm);
  if (c->dim == 3) {
    if (factor == SRRG2B_FACTOR_P2P) k3p<<<blocks, kLoopThreads, kLoopSmemBytes, c->stream>>>(a);
    else k3l<<<blocks, kLoopThreads, kLoopSmemBytes, c->stream>>>(a);
  } else {
    if (factor == SRRG2B_FACTOR_P2P) k2p<<<blocks, kLoopThreads, kLoopSmemBytes, c->stream>>>(a);
    else k2l<<<blocks, kLoopThreads, kLoopSmemBytes, c->stream>>>(a);
  }
  c->launches++;
}

// One pass over a slice inside the ICP loop (R/registration/aligners/aligner_slice_processor_impl.cpp:38-48):
// coherence check fused with the linearisation when the slice holds certified bounds, then the searches of
// whatever failed the check (everything, while no bounds exist) and the linearisation of what they found.
// sole: the iteration launches ONE search kernel (nn_far_kernel: short lists warp-per-query, long lists and late full
// searches thread-per-query, each with its linearisation) instead of nn_kernel + nn_far_kernel +
// lin_after_search_kernel -- three launches that have nothing to do in a converged iteration, yet cost 6 us of
// every iteration in the replayed graph (measured at C2).  The first full_iters iterations of a run, which search
// everything, keep the three-kernel pipeline.
int launch_slice_iteration(srrg2b_ctx* c, const SliceArgs& a0, int factor, const int* skip, bool sole = false, bool no_check = false) {
  if (a0.nm <= 0) return SRRG2B_OK;
  if (a0.projective || skip) {  // no coherence machinery for the index-image finder / the groups in front of the loop kernel
    int rcode = launch_find(c, a0, skip);
    if (rcode) return rcode;
    return launch_linearize(c, a0, factor, skip);
  }
  SliceArgs a = a0;
  a.use_list = 1;  // (the work-list counters were zeroed by icp_init_kernel / the previous solve step)
  // no_check: an iteration that cannot hold certified bounds yet (the first of a fresh run; the second too unless
  // every search certifies) -- the check kernel would read its control words and return: 2.9 us saved per launch
  if (!no_check)
    launch_tiles_kernel(c, a, factor, check_tiles_kernel<3, SRRG2B_FACTOR_P2P>, check_tiles_kernel<3, SRRG2B_FACTOR_PLANE>,
                        check_tiles_kernel<2, SRRG2B_FACTOR_P2P>, check_tiles_kernel<2, SRRG2B_FACTOR_PLANE>);
  if (sole) {
    a.sole_list = 1;
    launch_far(c, a, factor, nullptr);
    return SRRG2B_OK;
  }
  launch_nn(c, a, nullptr);
  launch_far(c, a, factor, nullptr);  // phase 2 of long lists, or the whole job for short ones (any R)
  launch_tiles_kernel(c, a, factor, lin_after_search_kernel<3, SRRG2B_FACTOR_P2P>, lin_after_search_kernel<3, SRRG2B_FACTOR_PLANE>,
                      lin_after_search_kernel<2, SRRG2B_FACTOR_P2P>, lin_after_search_kernel<2, SRRG2B_FACTOR_PLANE>);
  return SRRG2B_OK;
}

// ring-ordered (dy, dz) row offsets of the NN search neighbourhood -> __constant__ tables
int upload_row_tables() {
  signed char t3[kRowTable][4];
  int n = 0;
  for (int ring = 0; ring <= kMaxR; ++ring)
    for (int dz = -ring; dz <= ring; ++dz)
      for (int dy = -ring; dy <= ring; ++dy) {
        if (std::max(std::abs(dy), std::abs(dz)) != ring) continue;
        t3[n][0] = (signed char) dy; t3[n][1] = (signed char) dz; t3[n][2] = (signed char) ring; t3[n][3] = 0;
        ++n;
      }
  signed char t2[2 * kMaxR + 1][4];
  n = 0;
  for (int ring = 0; ring <= kMaxR; ++ring)
    for (int dy = -ring; dy <= ring; dy += (ring ? 2 * ring : 1)) {
      t2[n][0] = (signed char) dy; t2[n][1] = 0; t2[n][2] = (signed char) ring; t2[n][3] = 0;
      ++n;
    }
  if (cudaMemcpyToSymbol(c_rows3, t3, sizeof(t3)) != cudaSuccess) return SRRG2B_ERR_CUDA;
  if (cudaMemcpyToSymbol(c_rows2, t2, sizeof(t2)) != cudaSuccess) return SRRG2B_ERR_CUDA;
  return SRRG2B_OK;
}

// a transform the fixed-point ranges were derived for: rotation block entries bounded by 1 (+ rounding)
bool rigid_enough(int dim, const float* M) {
  const int D1 = dim + 1;
  for (int r = 0; r < dim; ++r)
    for (int cc = 0; cc < dim; ++cc)
      if (!(fabsf(M[r * D1 + cc]) <= 1.001f)) return false;
  for (int k = 0; k < D1 * D1; ++k)
    if (!(M[k] == M[k]) || isinf(M[k])) return false;
  return true;
}

int validate_slices(srrg2b_ctx* c, int n_slices, const srrg2b_slice* slices, int variable) {
  if (n_slices < 1 || n_slices > SRRG2B_MAX_SLICES || !slices)
    FAIL(c, SRRG2B_ERR_INVALID, "n_slices must be in [1, SRRG2B_MAX_SLICES]");
  for (int s = 0; s < n_slices; ++s) {
    const srrg2b_slice& sl = slices[s];
    if (sl.kind == SRRG2B_SLICE_PRIOR) {
      if (c->dim == 3 && variable != SRRG2B_VAR_SE3_QUAT_RIGHT)
        FAIL(c, SRRG2B_ERR_INVALID, "prior slices need the quaternion SE(3) variable (SE3PriorErrorFactorAD)");
      continue;
    }
    if (sl.kind != SRRG2B_SLICE_POINTS) FAIL(c, SRRG2B_ERR_INVALID, "unknown slice kind");
    if (!c->slices.count(sl.slice_id)) FAIL(c, SRRG2B_ERR_STATE, "slice has no clouds (no fixed / no moving)");
    if (sl.factor.factor != SRRG2B_FACTOR_P2P && sl.factor.factor != SRRG2B_FACTOR_PLANE)
      FAIL(c, SRRG2B_ERR_INVALID, "unknown factor kind");
    if (sl.factor.robustifier < 0 || sl.factor.robustifier > SRRG2B_ROB_HUBER)
      FAIL(c, SRRG2B_ERR_INVALID, "unknown robustifier");
    if (!rigid_enough(c->dim, sl.robot_in_sensor))
      FAIL(c, SRRG2B_ERR_INVALID, "robot_in_sensor is not a rigid transform");
  }
  return SRRG2B_OK;
}

struct Plan {
  SolveArgs solve;
  SliceArgs sargs[SRRG2B_MAX_SLICES];
  int factor[SRRG2B_MAX_SLICES];
  bool is_points[SRRG2B_MAX_SLICES];
};

int make_plan(srrg2b_ctx* c, int n_slices, const srrg2b_slice* slices, const srrg2b_aligner_params& ap, bool clamp,
              bool want_status, Plan& plan) {
  memset(&plan, 0, sizeof(plan));  // also the padding: plans are compared bytewise (graph cache key)
  plan.solve.dim = c->dim;
  plan.solve.variable = ap.variable;
  plan.solve.n_slices = n_slices;
  plan.solve.use_tc = ap.use_termination_criteria;
  plan.solve.window = ap.window_size;
  plan.solve.range_corr = ap.num_correspondences_range;
  plan.solve.range_inl = ap.num_inliers_range;
  plan.solve.range_out = ap.num_outliers_range;
  plan.solve.chi_eps = ap.chi_epsilon;
  for (int s = 0; s < n_slices; ++s) {
    const srrg2b_slice& sl = slices[s];
    SolveSlice& ss = plan.solve.sl[s];
    ss.kind = sl.kind;
    ss.min_corr = sl.min_num_correspondences;
    embed(c->dim, sl.robot_in_sensor, ss.ris);
    plan.is_points[s] = sl.kind == SRRG2B_SLICE_POINTS;
    if (sl.kind == SRRG2B_SLICE_PRIOR) {
      embed(c->dim, sl.prior_measurement, ss.Z);
      set_identity(ss.ris);
      for (int k = 0; k < 6; ++k) ss.info[k] = sl.prior_info_diag[k];
      continue;
    }
    set_identity(ss.Z);
    srrg2b_factor_params fa = sl.factor;
    if (clamp && fa.robustifier != SRRG2B_ROB_NONE) fa.robustifier = SRRG2B_ROB_CLAMP;  // multi_aligner_impl.cpp:193-199
    Scales sc;
    int rcode = fill_slice_args(c, c->slices[sl.slice_id], s, sl.finder, fa, ap.variable, want_status, plan.sargs[s], &sc);
    if (rcode) return rcode;
    plan.factor[s] = fa.factor;
    {
      SliceData& sdd = c->slices[sl.slice_id];
      ss.S_lb = sdd.S_lb.p;
      ss.cell = 1.f / sdd.inv_cell;
      ss.radius = sqrtf(sdd.radius2) * 1.0001f;
      ss.track2_mode = c->track2_mode;
      ss.track2_frac = c->track2_frac;
      ss.counters = sdd.far_count.p;
      ss.nn_points = sl.finder.kind == SRRG2B_FINDER_NN ? 1 : 0;
    }
    for (int k = 0; k < kKCount; ++k) ss.invk[k] = ldexp(1.0, -sc.k[k]);
  }
  return SRRG2B_OK;
}

int allreduce_acc(srrg2b_ctx* c, int n_slices, bool in_solve_kernel) {
  // (peer exchange: done by the solve step itself when it runs)
  if (c->world <= 1 || (c->d_px && in_solve_kernel)) return SRRG2B_OK;
  if (g_nccl.AllReduce(c->d_state->acc, c->d_state->acc, (size_t) n_slices * kAcc, kNcclUint64, kNcclSum, c->comm,
                       c->stream) != 0)
    FAIL(c, SRRG2B_ERR_NCCL, "ncclAllReduce(sum) of H/b/stats failed");
  return SRRG2B_OK;
}

// one _runSolver iteration (multi_aligner_impl.cpp:103-126) as a kernel sequence: full search +
// linearisation per point slice, then the solve step.  skip: device flag that makes the whole group a no-op.
int enqueue_iteration_group(srrg2b_ctx* c, const Plan& plan, const int* skip, bool sole = false, bool no_check = false) {
  for (int s = 0; s < plan.solve.n_slices; ++s) {
    if (!plan.is_points[s]) continue;
    if (c->time_kernels) {
      while (c->kev.size() < c->kev_used + 2) {
        cudaEvent_t e;
        CK(c, cudaEventCreate(&e));
        c->kev.push_back(e);
      }
      CK(c, cudaEventRecord(c->kev[c->kev_used], c->stream));
    }
    int rcode = launch_slice_iteration(c, plan.sargs[s], plan.factor[s], skip, sole, no_check);
    if (rcode) return rcode;
    if (c->time_kernels) {
      CK(c, cudaEventRecord(c->kev[c->kev_used + 1], c->stream));
      c->kev_used += 2;
    }
  }
  int rcode = allreduce_acc(c, plan.solve.n_slices, true);
  if (rcode) return rcode;
  if (c->dim == 3) icp_solve_kernel<3><<<1, kSolveThreads, 0, c->stream>>>(c->d_solve, c->d_state, c->d_px, skip);
  else icp_solve_kernel<2><<<1, kSolveThreads, 0, c->stream>>>(c->d_solve, c->d_state, c->d_px, skip);
  c->launches++;
  return SRRG2B_OK;
}

// enqueue `iterations` _runSolver iterations; no host sync inside
int enqueue_iterations(srrg2b_ctx* c, const Plan& plan, int iterations, int keep_stats) {
  for (int it = 0; it < iterations; ++it) {
    // a fresh run (icp_init_kernel: list_all = 1, track2 = (mode == 1)) has no bounds in its first pass, and none in
    // its second unless the first one certified (icp_solve_serial: list_all = !track2)
    const bool no_check = !keep_stats && (it == 0 || (it == 1 && c->track2_mode != 1));
    const int rcode = enqueue_iteration_group(c, plan, nullptr, it >= c->full_iters, no_check);
    if (rcode) return rcode;
  }
  CK(c, cudaGetLastError());
  return SRRG2B_OK;
}

// icp_init_kernel + `iterations` iterations for `plan`, starting from guess T0: replay of the cached CUDA
// graph of that launch sequence (captured on first use).
int run_plan(srrg2b_ctx* c, const Plan& plan, const Mat4f& T0, int iterations, int apply_prior_guess, int reset_tc,
             int keep_stats) {
  // the previous run has been synchronised (fetch_state), so the pinned staging buffers are free
  *c->h_T0 = T0;
  *c->h_solve = plan.solve;
  CK(c, cudaMemcpyAsync(c->d_T0, c->h_T0, sizeof(Mat4f), cudaMemcpyHostToDevice, c->stream));
  CK(c, cudaMemcpyAsync(c->d_solve, c->h_solve, sizeof(SolveArgs), cudaMemcpyHostToDevice, c->stream));
  // (several ranks: graph replay needs the peer-memory exchange -- no NCCL call inside the iteration)
  const bool graph = c->use_graphs && !c->time_kernels && (c->world <= 1 || c->graph_nccl || c->d_px);
  if (!graph) {
    icp_init_kernel<<<1, 32, 0, c->stream>>>(c->d_solve, c->d_state, c->d_T0, apply_prior_guess, reset_tc, keep_stats,
                                             iterations);
    c->launches++;
    return enqueue_iterations(c, plan, iterations, keep_stats);
  }
  const int tail[6] = {iterations, apply_prior_guess, reset_tc, keep_stats, c->full_iters, c->track2_mode};
  // key = everything the launch sequence depends on (slice kernel arguments, factor kinds, counts);
  // the solve-step values are read from device memory and may change freely between replays
  const size_t nb = sizeof(plan.sargs) + sizeof(plan.factor) + sizeof(plan.is_points);
  std::vector<unsigned char> key(nb + sizeof(tail) + sizeof(int));
  memcpy(key.data(), plan.sargs, sizeof(plan.sargs));
  memcpy(key.data() + sizeof(plan.sargs), plan.factor, sizeof(plan.factor));
  memcpy(key.data() + sizeof(plan.sargs) + sizeof(plan.factor), plan.is_points, sizeof(plan.is_points));
  memcpy(key.data() + nb, tail, sizeof(tail));
  memcpy(key.data() + nb + sizeof(tail), &plan.solve.n_slices, sizeof(int));
  srrg2b_ctx::RunGraph* hit = nullptr;
  for (auto& g : c->run_graphs)
    if (g.key == key) { hit = &g; break; }
  if (!hit) {
    if (c->run_graphs.size() >= 16) {  // plans of clouds long gone: drop the oldest
      if (c->run_graphs.front().exec) cudaGraphExecDestroy(c->run_graphs.front().exec);
      c->run_graphs.erase(c->run_graphs.begin());
    }
    const int64_t before = c->launches;
    cudaGraph_t g = nullptr;
    CK(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    icp_init_kernel<<<1, 32, 0, c->stream>>>(c->d_solve, c->d_state, c->d_T0, apply_prior_guess, reset_tc, keep_stats,
                                             iterations);
    c->launches++;
    const int rcode = enqueue_iterations(c, plan, iterations, keep_stats);
    const cudaError_t e = cudaStreamEndCapture(c->stream, &g);
    const int64_t n_launch = c->launches - before;
    c->launches = before;
    if (rcode) { if (g) cudaGraphDestroy(g); return rcode; }
    CK(c, e);
    srrg2b_ctx::RunGraph rg;
    rg.key = key;
    rg.launches = n_launch;
    const cudaError_t ei = cudaGraphInstantiate(&rg.exec, g, 0);
    cudaGraphDestroy(g);
    CK(c, ei);
    c->run_graphs.push_back(rg);
    hit = &c->run_graphs.back();
  }
  CK(c, cudaGraphLaunch(hit->exec, c->stream));
  c->launches += hit->launches;
  return SRRG2B_OK;
}

int fetch_state_async(srrg2b_ctx* c) {
  CK(c, cudaMemcpyAsync(c->h_state, c->d_state, sizeof(DevState), cudaMemcpyDeviceToHost, c->stream));
  return SRRG2B_OK;
}
int fetch_state_wait(srrg2b_ctx* c);
int fetch_state(srrg2b_ctx* c) {
  const int rcode = fetch_state_async(c);
  return rcode ? rcode : fetch_state_wait(c);
}
int fetch_state_wait(srrg2b_ctx* c) {
  CK(c, cudaStreamSynchronize(c->stream));
  if (c->h_state->error == 1) FAIL(c, SRRG2B_ERR_NCCL, "peer exchange timed out: a rank never delivered its accumulators");
  if (c->h_state->error == 2) FAIL(c, SRRG2B_ERR_CUDA, "grid barrier of the device loop timed out");
  if (c->time_kernels) {
    c->last_kernel_ms = 0.f;
    c->last_kernel_launches = 0;
    for (size_t i = 0; i + 1 < c->kev_used; i += 2) {
      float ms = 0.f;
      CK(c, cudaEventElapsedTime(&ms, c->kev[i], c->kev[i + 1]));
      c->last_kernel_ms += ms;
      c->last_kernel_launches++;
    }
    c->kev_used = 0;
  }
  return SRRG2B_OK;
}

int export_corr(srrg2b_ctx* c, SliceData& sd, int prune, bool want_stat, int32_t* fixed_idx, int32_t* moving_idx,
                float* response, uint8_t* status, float* chi, int64_t* n_out) {
  const int n = (int) sd.moving_raw.n;
  *n_out = 0;
  if (n == 0 || !sd.corr_valid) return SRRG2B_OK;
  CK(c, c->flags.ensure(n));
  CK(c, c->positions.ensure(n));
  CK(c, c->d_fidx.ensure(n));
  CK(c, c->d_resp.ensure(n));
  CK(c, c->o_fidx.ensure(n));
  CK(c, c->o_midx.ensure(n));
  CK(c, c->o_resp.ensure(n));
  if (want_stat) {
    CK(c, c->d_stat.ensure(n));
    CK(c, c->d_chi.ensure(n));
    CK(c, c->o_stat.ensure(n));
    CK(c, c->o_chi.ensure(n));
  }
  CK(c, cudaMemsetAsync(c->flags.p, 0, sizeof(int) * n, c->stream));
  if (sd.nm_valid > 0) {
    export_dense_kernel<<<blocks_for(sd.nm_valid, 256), 256, 0, c->stream>>>(
      sd.m_pts.p, sd.f_pts.p, sd.c_fpos.p, sd.c_fidx.p, sd.S_lb.p, c->dim, sd.stat_valid ? sd.c_stat.p : nullptr, sd.stat_valid ? sd.c_chi.p : nullptr,
      sd.nm_valid, prune, c->d_fidx.p, c->d_resp.p, want_stat ? c->d_stat.p : nullptr, want_stat ? c->d_chi.p : nullptr,
      c->flags.p);
    c->launches++;
  }
  size_t bytes = 0;
  CK(c, cub::DeviceScan::ExclusiveSum(nullptr, bytes, c->flags.p, c->positions.p, n, c->stream));
  CK(c, c->cub_tmp.ensure(bytes));
  CK(c, cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, bytes, c->flags.p, c->positions.p, n, c->stream));
  compact_kernel<<<blocks_for(n, 256), 256, 0, c->stream>>>(
    c->flags.p, c->positions.p, n, (int) sd.moving_raw.index_offset, c->d_fidx.p, c->d_resp.p,
    want_stat ? c->d_stat.p : nullptr, want_stat ? c->d_chi.p : nullptr, c->o_fidx.p, c->o_midx.p, c->o_resp.p,
    want_stat ? c->o_stat.p : nullptr, want_stat ? c->o_chi.p : nullptr);
  c->launches++;
  int last_flag = 0, last_pos = 0;
  CK(c, cudaMemcpyAsync(&last_flag, c->flags.p + (n - 1), 4, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaMemcpyAsync(&last_pos, c->positions.p + (n - 1), 4, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  const int m = last_flag + last_pos;
  if (m > 0) {
    if (fixed_idx) CK(c, cudaMemcpyAsync(fixed_idx, c->o_fidx.p, 4 * (size_t) m, cudaMemcpyDeviceToHost, c->stream));
    if (moving_idx) CK(c, cudaMemcpyAsync(moving_idx, c->o_midx.p, 4 * (size_t) m, cudaMemcpyDeviceToHost, c->stream));
    if (response) CK(c, cudaMemcpyAsync(response, c->o_resp.p, 4 * (size_t) m, cudaMemcpyDeviceToHost, c->stream));
    if (status && want_stat) CK(c, cudaMemcpyAsync(status, c->o_stat.p, (size_t) m, cudaMemcpyDeviceToHost, c->stream));
    if (chi && want_stat) CK(c, cudaMemcpyAsync(chi, c->o_chi.p, 4 * (size_t) m, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
  }
  *n_out = m;
  return SRRG2B_OK;
}

// MultiAlignerBase_::compute() (R/registration/aligners/multi_aligner_impl.cpp:46-95, :162-180) in three host steps, so
// that several contexts can have their runs in flight at once (srrg2b_closure_batch): begin() queues the first run
// and the read-back of its state, mid() waits for it, decides the status and queues the inlier-only run when one is
// due, end() waits for that and writes the results.  srrg2b_icp_run is begin + mid + end on one context.
struct RunCall {
  srrg2b_ctx* c;
  int n_slices;
  const srrg2b_slice* slices;
  const srrg2b_aligner_params* ap;
  float* T;
  srrg2b_iter_stats* stats_out;
  int32_t* n_stats;
  int32_t* aligner_status;
  Mat4f T0 = {};
  bool want_status = false, done = false, second = false;
  int status = SRRG2B_ALIGNER_FAIL;
};

int run_begin(RunCall& r) {
  srrg2b_ctx* c = r.c;
  const srrg2b_aligner_params* ap = r.ap;
  if (!ap || !r.T || !r.aligner_status) FAIL(c, SRRG2B_ERR_INVALID, "null argument");
  if (ap->max_iterations < 0) FAIL(c, SRRG2B_ERR_INVALID, "max_iterations < 0");
  int rcode = validate_slices(c, r.n_slices, r.slices, ap->variable);
  if (rcode) return rcode;
  CK(c, cudaSetDevice(c->device));
  r.want_status = ap->keep_only_inlier_correspondences != 0;
  Plan plan;
  rcode = make_plan(c, r.n_slices, r.slices, *ap, false, r.want_status, plan);
  if (rcode) return rcode;
  if (!rigid_enough(c->dim, r.T)) FAIL(c, SRRG2B_ERR_INVALID, "initial guess is not a rigid transform");
  embed(c->dim, r.T, r.T0);
  for (int s = 0; s < r.n_slices; ++s)
    if (r.slices[s].kind == SRRG2B_SLICE_POINTS) c->slices[r.slices[s].slice_id].have_last_S = false;
  CK(c, cudaEventRecord(c->ev0, c->stream));
  rcode = run_plan(c, plan, r.T0, ap->max_iterations, 1, 1, 0);  // multi_aligner_impl.cpp:72
  if (rcode) return rcode;
  CK(c, cudaEventRecord(c->ev1, c->stream));
  return fetch_state_async(c);
}

int run_mid(RunCall& r) {
  srrg2b_ctx* c = r.c;
  const srrg2b_aligner_params* ap = r.ap;
  CK(c, cudaSetDevice(c->device));
  int rcode = fetch_state_wait(c);
  if (rcode) return rcode;
  float ms = 0.f;
  CK(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  c->last_ms = ms;
  c->last_iterations = c->h_state->iterations_run;
  const DevState* hs = c->h_state;
  if (hs->n_stats == 0) {  // :75-78
    r.status = SRRG2B_ALIGNER_FAIL;
    r.done = true;
  } else if (hs->last_stats.num_inliers < ap->min_num_inliers) {  // :81-85 (newest entry, also when the array is full)
    r.status = SRRG2B_ALIGNER_NOT_ENOUGH_INLIERS;
    r.done = true;
  }
  if (!r.done && ap->enable_inlier_only_runs) {  // _postCompute, :162-175
    Plan plan2;
    rcode = make_plan(c, r.n_slices, r.slices, *ap, true, r.want_status, plan2);
    if (rcode) return rcode;
    CK(c, cudaEventRecord(c->ev0, c->stream));
    rcode = run_plan(c, plan2, r.T0, ap->max_iterations, 0, 0, 1);
    if (rcode) return rcode;
    CK(c, cudaEventRecord(c->ev1, c->stream));
    r.second = true;
    return fetch_state_async(c);
  }
  return SRRG2B_OK;
}

int run_end(RunCall& r) {
  srrg2b_ctx* c = r.c;
  CK(c, cudaSetDevice(c->device));
  if (r.second) {
    const int rcode = fetch_state_wait(c);
    if (rcode) return rcode;
    float ms = 0.f;
    CK(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->last_ms += ms;
    c->last_iterations = c->h_state->iterations_run;
  }
  Mat4f X = c->h_state->X;
  if (!r.done) {
    fix_transform(c->dim, X);  // :90-93
    r.status = SRRG2B_ALIGNER_SUCCESS;
  }
  unembed(c->dim, X, r.T);
  const int have = std::min(c->h_state->n_stats, kMaxStats);
  if (r.stats_out && r.n_stats) {
    const int cap = *r.n_stats;
    for (int i = 0; i < have && i < cap; ++i) r.stats_out[i] = c->h_state->stats[i];
  }
  if (r.n_stats) *r.n_stats = c->h_state->n_stats;
  *r.aligner_status = r.status;
  for (int s = 0; s < r.n_slices; ++s) {
    if (r.slices[s].kind != SRRG2B_SLICE_POINTS) continue;
    SliceData& sd = c->slices[r.slices[s].slice_id];
    sd.corr_valid = c->h_state->iterations_run > 0;
    sd.stat_valid = r.want_status && sd.corr_valid;
    sd.prune_on_export = (!r.done && r.want_status) ? 1 : 0;  // :177-180
  }
  return SRRG2B_OK;
}

}  // namespace

// =================================================================================================
// extern "C" boundary
// =================================================================================================
extern "C" {

int srrg2b_version(void) { return SRRG2B_VERSION; }

int srrg2b_ctx_create(int dim, int device, srrg2b_ctx** out) {
  if (!out || (dim != 2 && dim != 3)) return SRRG2B_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device < 0 || device >= count) return SRRG2B_ERR_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return SRRG2B_ERR_CUDA;
  srrg2b_ctx* c = new srrg2b_ctx();
  c->dim = dim;
  c->device = device;
  if (const char* env = getenv("SRRG2B_TRACK2")) c->track2_mode = std::max(0, std::min(2, atoi(env)));
  if (const char* env = getenv("SRRG2B_TRACK2_FRAC")) c->track2_frac = (float) atof(env);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
  bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaEventCreate(&c->ev0) == cudaSuccess && cudaEventCreate(&c->ev1) == cudaSuccess;
  ok = ok && cudaMalloc((void**) &c->d_state, sizeof(DevState)) == cudaSuccess;
  ok = ok && cudaMallocHost((void**) &c->h_state, sizeof(DevState)) == cudaSuccess;
  ok = ok && cudaMallocHost((void**) &c->h_bounds, kBoundWords * sizeof(int)) == cudaSuccess;
  ok = ok && cudaMallocHost((void**) &c->h_bounds_init, kBoundWords * sizeof(int)) == cudaSuccess;
  if (ok) {
    const int init[kBoundWords] = {INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN, 0, 0, 0, 0, 0, 0};
    memcpy(c->h_bounds_init, init, sizeof(init));
  }
  ok = ok && cudaMalloc((void**) &c->d_T0, sizeof(Mat4f)) == cudaSuccess;
  ok = ok && cudaMallocHost((void**) &c->h_T0, sizeof(Mat4f)) == cudaSuccess;
  ok = ok && cudaMalloc((void**) &c->d_solve, sizeof(SolveArgs)) == cudaSuccess;
  ok = ok && cudaMallocHost((void**) &c->h_solve, sizeof(SolveArgs)) == cudaSuccess;
  if (const char* env = getenv("SRRG2B_NO_GRAPH")) c->use_graphs = atoi(env) == 0;
  {
    int khz = 1965000;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
    double ms = 2000.0;
    if (const char* env = getenv("SRRG2B_TIMEOUT_MS")) ms = std::max(1.0, atof(env));
    c->timeout_cycles = (long long) (ms * (double) khz);
  }
  if (const char* env = getenv("SRRG2B_FULL_ITERS")) c->full_iters = std::max(0, atoi(env));
  if (const char* env = getenv("SRRG2B_NN_FLAT")) c->nn_flat = atoi(env);
  if (const char* env = getenv("SRRG2B_FAR_SOLE_CTAS")) c->far_sole_ctas = std::min(8, std::max(1, atoi(env)));
  if (const char* env = getenv("SRRG2B_SMALL_SHIFT")) c->small_shift = std::min(20, std::max(0, atoi(env)));
  if (const char* env = getenv("SRRG2B_EAGER_INDEX")) c->eager_index = atoi(env) != 0;
  if (const char* env = getenv("SRRG2B_GRAPH_NCCL")) c->graph_nccl = atoi(env) != 0;
  {
    const void* lin[] = {(const void*) lin_tiles_kernel<3, SRRG2B_FACTOR_P2P>, (const void*) lin_tiles_kernel<3, SRRG2B_FACTOR_PLANE>,
                         (const void*) lin_tiles_kernel<2, SRRG2B_FACTOR_P2P>, (const void*) lin_tiles_kernel<2, SRRG2B_FACTOR_PLANE>,
                         (const void*) check_tiles_kernel<3, SRRG2B_FACTOR_P2P>, (const void*) check_tiles_kernel<3, SRRG2B_FACTOR_PLANE>,
                         (const void*) check_tiles_kernel<2, SRRG2B_FACTOR_P2P>, (const void*) check_tiles_kernel<2, SRRG2B_FACTOR_PLANE>,
                         (const void*) lin_after_search_kernel<3, SRRG2B_FACTOR_P2P>, (const void*) lin_after_search_kernel<3, SRRG2B_FACTOR_PLANE>,
                         (const void*) lin_after_search_kernel<2, SRRG2B_FACTOR_P2P>, (const void*) lin_after_search_kernel<2, SRRG2B_FACTOR_PLANE>};
    for (const void* f : lin)
      ok = ok && cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kLoopSmemBytes) == cudaSuccess;
    // (static + dynamic shared memory of the tile kernels must fit one CTA per SM: fail loudly here, not at launch)
    int per_sm = 0;
    ok = ok && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, check_tiles_kernel<3, SRRG2B_FACTOR_PLANE>, kLoopThreads, kLoopSmemBytes) == cudaSuccess && per_sm >= 1;
  }
  ok = ok && cudaMemsetAsync(c->d_state, 0, sizeof(DevState), c->stream) == cudaSuccess;
  ok = ok && cudaStreamSynchronize(c->stream) == cudaSuccess;
  ok = ok && upload_row_tables() == SRRG2B_OK;
  if (!ok) {
    srrg2b_ctx_destroy(c);
    return SRRG2B_ERR_CUDA;
  }
  *out = c;
  return SRRG2B_OK;
}

static void pgo_release(srrg2b_ctx* c);

int srrg2b_ctx_destroy(srrg2b_ctx* c) {
  if (!c) return SRRG2B_OK;
  pgo_release(c);
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
  {
#ifdef S2B_FAIL_STATS
    {
      unsigned long long fs[8];
      if (cudaMemcpyFromSymbol(fs, g_fail_stats, sizeof(fs)) == cudaSuccess)
        fprintf(stderr, "[srrg2b fail stats] all loop iterations of all runs: no bound %llu | budget spent %llu | runner-up too close %llu (mean margin %.1f um) | none too close %llu | out of range %llu\n",
                fs[0], fs[1], fs[2], fs[2] ? (double) fs[5] / (double) fs[2] - 1000.0 : 0.0, fs[3], fs[4]);
    }
#endif
#if S2B_SOLVE_STAMPS
    unsigned long long st8[8];
    if (cudaMemcpyFromSymbol(st8, g_solve_stamps, sizeof(st8)) == cudaSuccess) {
      fprintf(stderr, "[srrg2b solve stamps] last solve step (us): stage+exchange %.2f | assemble %.2f | cholesky %.2f | box_plus %.2f | transforms+budget %.2f | termination %.2f | write-back %.2f\n",
              (st8[1] - st8[0]) * 1e-3, (st8[2] - st8[1]) * 1e-3, (st8[3] - st8[2]) * 1e-3, (st8[4] - st8[3]) * 1e-3,
              (st8[5] - st8[4]) * 1e-3, (st8[6] - st8[5]) * 1e-3, (st8[7] - st8[6]) * 1e-3);
    }
#endif
  }
  if (c->d_px && c->comm && g_nccl.AllReduce && c->d_epoch) {
    // a peer may still be adding this rank's last mailbox words: leave together (contexts of a
    // communicator are destroyed collectively, like the communicator itself)
    g_nccl.AllReduce(c->d_epoch, c->d_epoch, 1, kNcclUint64, kNcclMax, c->comm, c->stream);
    cudaStreamSynchronize(c->stream);
  }
  for (void* p : c->peer_maps) cudaIpcCloseMemHandle(p);
  if (c->d_px) cudaFree(c->d_px);
  if (c->d_mail) cudaFree(c->d_mail);
  if (c->d_epoch) cudaFree(c->d_epoch);
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  for (auto& kv : c->slices) {
    SliceData& s = kv.second;
    for (RawCloud* rcp : {&s.fixed_raw, &s.moving_raw}) {
      if (rcp->ev_coords) cudaEventDestroy(rcp->ev_coords);
      if (rcp->ev_all) cudaEventDestroy(rcp->ev_all);
      if (rcp->d_nb2) cudaFree(rcp->d_nb2);
      if (rcp->h_nb2) cudaFreeHost(rcp->h_nb2);
    }
    s.fixed_raw.xyz.release(); s.fixed_raw.nrm.release(); s.fixed_raw.valid.release();
    s.moving_raw.xyz.release(); s.moving_raw.nrm.release(); s.moving_raw.valid.release();
    s.f_pts.release(); s.f_rec.release(); s.f_inverse.release(); s.cell_start.release(); s.near_bits.release();
    s.m_pts.release(); s.m_nrm.release(); s.m_pair.release(); s.m_inverse.release();
    s.image.release(); s.c_lb.release(); s.S_lb.release(); s.c_fidx.release(); s.c_fpos.release(); s.far_list.release(); s.far_count.release(); s.work_list.release(); s.c_resp.release(); s.c_chi.release(); s.c_stat.release();
  }
  c->keys_a.release(); c->keys_b.release(); c->vals_a.release(); c->vals_b.release();
  c->flags.release(); c->positions.release(); c->bounds.release(); c->cub_tmp.release();
  c->o_fidx.release(); c->o_midx.release(); c->d_fidx.release(); c->o_resp.release(); c->d_resp.release();
  c->o_chi.release(); c->d_chi.release(); c->o_stat.release(); c->d_stat.release();
  c->imp_f.release(); c->imp_m.release(); c->imp_bad.release();
  if (c->d_state) cudaFree(c->d_state);
  if (c->h_state) cudaFreeHost(c->h_state);
  if (c->h_bounds) cudaFreeHost(c->h_bounds);
  for (auto& g : c->run_graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
  c->run_graphs.clear();
  if (c->d_T0) cudaFree(c->d_T0);
  if (c->h_T0) cudaFreeHost(c->h_T0);
  if (c->d_solve) cudaFree(c->d_solve);
  if (c->h_solve) cudaFreeHost(c->h_solve);
  for (cudaEvent_t e : c->kev) cudaEventDestroy(e);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return SRRG2B_OK;
}

const char* srrg2b_last_error(const srrg2b_ctx* c) { return c ? c->err.c_str() : "null context"; }
void* srrg2b_stream(srrg2b_ctx* c) { return c ? (void*) c->stream : nullptr; }
int64_t srrg2b_launch_count(const srrg2b_ctx* c) { return c ? c->launches : 0; }

int srrg2b_comm_unique_id(void* id) {
  if (!id) return SRRG2B_ERR_INVALID;
  if (!g_nccl.load()) return SRRG2B_ERR_NCCL;
  NcclId nid;
  if (g_nccl.GetUniqueId(&nid) != 0) return SRRG2B_ERR_NCCL;
  memcpy(id, &nid, sizeof(nid));
  return SRRG2B_OK;
}

int srrg2b_comm_init(srrg2b_ctx* c, const void* id, int rank, int world) {
  if (!c || !id || world < 1 || rank < 0 || rank >= world) return SRRG2B_ERR_INVALID;
  if (world == 1) {
    c->rank = 0;
    c->world = 1;
    return SRRG2B_OK;
  }
  if (!g_nccl.load()) FAIL(c, SRRG2B_ERR_NCCL, "libnccl.so.2 not found");
  CK(c, cudaSetDevice(c->device));
  NcclId nid;
  memcpy(&nid, id, sizeof(nid));
  const int r = g_nccl.CommInitRank(&c->comm, world, nid, rank);
  if (r != 0) FAIL(c, SRRG2B_ERR_NCCL, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
  c->rank = rank;
  c->world = world;
  // Peer-memory exchange (default; SRRG2B_PEER_EXCHANGE=0 keeps the NCCL all-reduce): every rank
  // allocates its mailbox, the IPC handles travel through one ncclAllGather, peers are mapped.
  const char* env = getenv("SRRG2B_PEER_EXCHANGE");
  if ((env && atoi(env) == 0) || world > kMaxRanks || !g_nccl.AllGather) return SRRG2B_OK;
  const size_t mail_bytes = 2 * kMailWords * sizeof(unsigned long long);
  CK(c, cudaMalloc((void**) &c->d_mail, mail_bytes));
  CK(c, cudaMemsetAsync(c->d_mail, 0, mail_bytes, c->stream));
  CK(c, cudaMalloc((void**) &c->d_epoch, sizeof(unsigned long long)));
  CK(c, cudaMemsetAsync(c->d_epoch, 0, sizeof(unsigned long long), c->stream));
  cudaIpcMemHandle_t mine;
  CK(c, cudaIpcGetMemHandle(&mine, c->d_mail));
  unsigned char* d_handles = nullptr;
  CK(c, cudaMalloc((void**) &d_handles, sizeof(mine) * (size_t) world));
  CK(c, cudaMemcpyAsync(d_handles + sizeof(mine) * (size_t) rank, &mine, sizeof(mine), cudaMemcpyHostToDevice, c->stream));
  if (g_nccl.AllGather(d_handles + sizeof(mine) * (size_t) rank, d_handles, sizeof(mine), /*ncclUint8*/ 1, c->comm,
                       c->stream) != 0) {
    cudaFree(d_handles);
    FAIL(c, SRRG2B_ERR_NCCL, "ncclAllGather of the mailbox handles failed");
  }
  std::vector<cudaIpcMemHandle_t> all((size_t) world);
  CK(c, cudaMemcpyAsync(all.data(), d_handles, sizeof(mine) * (size_t) world, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  cudaFree(d_handles);
  PeerExchange px;
  memset(&px, 0, sizeof(px));
  px.rank = rank;
  px.world = world;
  px.timeout_cycles = c->timeout_cycles;
  bool mapped = true;
  for (int r = 0; r < world && mapped; ++r) {
    if (r == rank) { px.mail[r] = c->d_mail; continue; }
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, all[(size_t) r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      mapped = false;
      break;
    }
    c->peer_maps.push_back(p);
    px.mail[r] = (unsigned long long*) p;
  }
  // the decision must be the same on every rank: agree through a one-word all-reduce (min)
  int* d_ok = nullptr;
  CK(c, cudaMalloc((void**) &d_ok, sizeof(int)));
  const int ok_mine = mapped ? 1 : 0;
  CK(c, cudaMemcpyAsync(d_ok, &ok_mine, sizeof(int), cudaMemcpyHostToDevice, c->stream));
  int ok_all = 0;
  if (g_nccl.AllReduce(d_ok, d_ok, 1, /*ncclInt32*/ 2, /*ncclMin*/ 3, c->comm, c->stream) != 0) ok_all = 0;
  else {
    CK(c, cudaMemcpyAsync(&ok_all, d_ok, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
  }
  cudaFree(d_ok);
  if (ok_all) {
    CK(c, cudaMalloc((void**) &c->d_px, sizeof(PeerExchange)));
    CK(c, cudaMemcpyAsync(c->d_px, &px, sizeof(px), cudaMemcpyHostToDevice, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
  }
  return SRRG2B_OK;
}

int srrg2b_set_cloud(srrg2b_ctx* c, int slot, int slice_id, const srrg2b_cloud* cl) {
  if (!c) return SRRG2B_ERR_INVALID;
  if (!cl || cl->n < 0 || (cl->n > 0 && !cl->coords) || cl->n > 0x7fffffff00ll / 16)
    FAIL(c, SRRG2B_ERR_INVALID, "bad cloud descriptor");
  if (slot != SRRG2B_FIXED && slot != SRRG2B_MOVING) FAIL(c, SRRG2B_ERR_INVALID, "slot must be FIXED or MOVING");
  CK(c, cudaSetDevice(c->device));
  SliceData& sd = c->slices[slice_id];
  if (slot == SRRG2B_FIXED) {
    CK(c, cudaStreamSynchronize(c->stream));  // an index build of the previous fixed cloud may still be running
    int rcode = upload_raw(c, sd.fixed_raw, cl);
    if (rcode) return rcode;
    const bool was_nn = !sd.index_is_projective;
    sd.built_for_max_distance = -1.f;  // _fixed_changed_flag: the index has to be rebuilt
    sd.index_is_projective = false;
    sd.corr_valid = false;
    // The NN index is built for the finder radius the slice used last, right away and WITHOUT waiting for
    // it: the kernels run on the compute stream while the caller uploads the moving cloud on the copy
    // stream (a different radius at run time simply rebuilds).  First use of a slice: built lazily.
    if (was_nn && sd.last_max_distance > 0.f && c->eager_index) rcode = ensure_index(c, sd, sd.last_max_distance);
    if (!rcode) rcode = queue_normal_bound(c, sd.fixed_raw);
    CK(c, cudaStreamSynchronize(c->copy_stream));  // the caller's buffers are free again
    return rcode;
  }
  int rcode = upload_raw(c, sd.moving_raw, cl);
  if (rcode) return rcode;
  rcode = build_moving(c, sd);
  CK(c, cudaStreamSynchronize(c->copy_stream));
  if (rcode) return rcode;
  CK(c, cudaStreamSynchronize(c->stream));
  return SRRG2B_OK;
}

// ---- N1: device-resident scene, clipped on the device into a slice's moving cloud ----
int srrg2b_scene_set(srrg2b_ctx* c, int scene_id, const srrg2b_cloud* cl) {
  if (!c) return SRRG2B_ERR_INVALID;
  if (!cl || cl->n < 0 || (cl->n > 0 && !cl->coords) || cl->n > 0x7fffffff00ll / 16) FAIL(c, SRRG2B_ERR_INVALID, "bad cloud descriptor");
  CK(c, cudaSetDevice(c->device));
  CK(c, cudaStreamSynchronize(c->stream));  // a clip of the previous scene may still be running
  const int rcode = upload_raw(c, c->scenes[scene_id], cl);
  CK(c, cudaStreamSynchronize(c->copy_stream));
  return rcode;
}

int srrg2b_scene_clip(srrg2b_ctx* c, int scene_id, int slice_id, const float* scene_in_robot, float max_range, int64_t* n_clipped) {
  if (!c) return SRRG2B_ERR_INVALID;
  if (!scene_in_robot || !(max_range > 0.f)) FAIL(c, SRRG2B_ERR_INVALID, "bad clip arguments");
  if (!c->scenes.count(scene_id) || !c->scenes[scene_id].present) FAIL(c, SRRG2B_ERR_STATE, "unknown scene");
  CK(c, cudaSetDevice(c->device));
  RawCloud& sc = c->scenes[scene_id];
  SliceData& sd = c->slices[slice_id];
  const int n = (int) sc.n, dim = c->dim;
  Mat4f T;
  embed(dim, scene_in_robot, T);
  if (!rigid_enough(dim, scene_in_robot)) FAIL(c, SRRG2B_ERR_INVALID, "scene_in_robot is not a rigid transform");
  int kept = 0;
  if (n > 0) {
    CK(c, c->flags.ensure(n));
    CK(c, c->positions.ensure(n));
    const float r2 = max_range * max_range;
    const unsigned char* valid = sc.has_valid ? sc.valid.p : nullptr;
    if (dim == 3) scene_clip_flag_kernel<3><<<blocks_for(n, 256), 256, 0, c->stream>>>(sc.xyz.p, valid, n, T, r2, c->flags.p);
    else scene_clip_flag_kernel<2><<<blocks_for(n, 256), 256, 0, c->stream>>>(sc.xyz.p, valid, n, T, r2, c->flags.p);
    size_t bytes = 0;
    CK(c, cub::DeviceScan::ExclusiveSum(nullptr, bytes, c->flags.p, c->positions.p, n, c->stream));
    CK(c, c->cub_tmp.ensure(bytes));
    CK(c, cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, bytes, c->flags.p, c->positions.p, n, c->stream));
    int last_flag = 0, last_pos = 0;
    CK(c, cudaMemcpyAsync(&last_flag, c->flags.p + (n - 1), 4, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaMemcpyAsync(&last_pos, c->positions.p + (n - 1), 4, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    kept = last_pos + last_flag;
    CK(c, sd.clip_xyz.ensure((size_t) std::max(kept, 1) * dim));
    CK(c, sd.clip_nrm.ensure((size_t) std::max(kept, 1) * dim));
    CK(c, sd.clip_gidx.ensure((size_t) std::max(kept, 1)));
    const float* nrm = sc.has_normals ? sc.nrm.p : nullptr;
    if (dim == 3) scene_clip_compact_kernel<3><<<blocks_for(n, 256), 256, 0, c->stream>>>(sc.xyz.p, nrm, c->flags.p, c->positions.p, n, T,
                                                                                       sd.clip_xyz.p, sd.clip_nrm.p, sd.clip_gidx.p);
    else scene_clip_compact_kernel<2><<<blocks_for(n, 256), 256, 0, c->stream>>>(sc.xyz.p, nrm, c->flags.p, c->positions.p, n, T,
                                                                                  sd.clip_xyz.p, sd.clip_nrm.p, sd.clip_gidx.p);
    c->launches += 2;
    CK(c, cudaGetLastError());
    CK(c, cudaStreamSynchronize(c->stream));  // the moving-cloud path below copies on the copy stream
  }
  sd.n_clipped = kept;
  sd.clip_scene_id = scene_id;
  if (n_clipped) *n_clipped = kept;
  // the clipped scene becomes the slice's moving cloud without leaving the device
  srrg2b_cloud cl;
  memset(&cl, 0, sizeof(cl));
  cl.coords = sd.clip_xyz.p;
  cl.normals = sc.has_normals ? sd.clip_nrm.p : nullptr;
  cl.n = kept;
  cl.on_device = 1;
  if (kept == 0) { static const float dummy = 0.f; cl.coords = &dummy; cl.on_device = 0; }
  return srrg2b_set_cloud(c, SRRG2B_MOVING, slice_id, &cl);
}

int srrg2b_scene_merge(srrg2b_ctx* c, int scene_id, int slice_id, const float* measurement_in_scene, const srrg2b_merge_params* mp,
                       int64_t* n_merged, int64_t* n_added) {
  if (!c) return SRRG2B_ERR_INVALID;
  if (!measurement_in_scene || !mp) FAIL(c, SRRG2B_ERR_INVALID, "null argument");
  if (!c->scenes.count(scene_id) || !c->scenes[scene_id].present) FAIL(c, SRRG2B_ERR_STATE, "unknown scene");
  if (!c->slices.count(slice_id) || !c->slices[slice_id].fixed_raw.present) FAIL(c, SRRG2B_ERR_STATE, "the slice has no measurement (fixed) cloud");
  if (!rigid_enough(c->dim, measurement_in_scene)) FAIL(c, SRRG2B_ERR_INVALID, "measurement_in_scene is not a rigid transform");
  CK(c, cudaSetDevice(c->device));
  RawCloud& sc = c->scenes[scene_id];
  SliceData& sd = c->slices[slice_id];
  RawCloud& me = sd.fixed_raw;
  const int dim = c->dim, n_meas = (int) me.n, n_scene = (int) sc.n;
  if (sc.has_normals && !me.has_normals) FAIL(c, SRRG2B_ERR_INVALID, "the scene carries normals, the measurement does not");
  Mat4f T;
  embed(dim, measurement_in_scene, T);
  cudaStream_t st = c->stream;
  if (me.ev_all) CK(c, cudaStreamWaitEvent(st, me.ev_all, 0));
  CK(c, c->keys_a.ensure((size_t) std::max(n_meas, 1)));  // merged marks (as ints)
  int* merged = reinterpret_cast<int*>(c->keys_a.p);
  CK(c, cudaMemsetAsync(merged, 0, sizeof(int) * (size_t) std::max(n_meas, 1), st));
  int64_t merged_count = 0, added = 0;
  bool append = true;
  if (!mp->without_correspondences) {
    // the slice's moving cloud must be a clip of THIS scene, and a compute() must have left correspondences
    if (sd.n_clipped != (int64_t) sd.moving_raw.n || sd.clip_scene_id != scene_id) FAIL(c, SRRG2B_ERR_STATE, "the slice's moving cloud is not a clip of this scene");
    if (!sd.corr_valid) FAIL(c, SRRG2B_ERR_STATE, "the slice holds no correspondences (run the aligner first)");
    const int n_local = (int) sd.moving_raw.n;
    if (n_local > 0 && n_meas > 0) {
      CK(c, c->flags.ensure(n_local)); CK(c, c->d_fidx.ensure(n_local)); CK(c, c->d_resp.ensure(n_local));
      CK(c, cudaMemsetAsync(c->flags.p, 0, sizeof(int) * (size_t) n_local, st));
      if (sd.nm_valid > 0) {
        export_dense_kernel<<<blocks_for(sd.nm_valid, 256), 256, 0, st>>>(sd.m_pts.p, sd.f_pts.p, sd.c_fpos.p, sd.c_fidx.p, sd.S_lb.p, dim,
                                                                          sd.stat_valid ? sd.c_stat.p : nullptr, sd.stat_valid ? sd.c_chi.p : nullptr,
                                                                          sd.nm_valid, sd.prune_on_export, c->d_fidx.p, c->d_resp.p, nullptr, nullptr,
                                                                          c->flags.p);
        c->launches++;
      }
      float* s_nrm = sc.has_normals ? sc.nrm.p : nullptr;
      const float* m_nrm = me.has_normals ? me.nrm.p : nullptr;
      if (dim == 3) scene_merge_kernel<3><<<blocks_for(n_local, 256), 256, 0, st>>>(c->flags.p, c->d_fidx.p, c->d_resp.p, sd.clip_gidx.p, n_local, me.xyz.p, m_nrm, T,
                                                                                mp->maximum_response, mp->maximum_distance_geometry_squared, sc.xyz.p, s_nrm, merged);
      else scene_merge_kernel<2><<<blocks_for(n_local, 256), 256, 0, st>>>(c->flags.p, c->d_fidx.p, c->d_resp.p, sd.clip_gidx.p, n_local, me.xyz.p, m_nrm, T,
                                                                           mp->maximum_response, mp->maximum_distance_geometry_squared, sc.xyz.p, s_nrm, merged);
      c->launches++;
      // number of DISTINCT measurement points merged (:79)
      CK(c, c->positions.ensure((size_t) n_meas));
      size_t bytes = 0;
      CK(c, cub::DeviceScan::ExclusiveSum(nullptr, bytes, merged, c->positions.p, n_meas, st));
      CK(c, c->cub_tmp.ensure(bytes));
      CK(c, cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, bytes, merged, c->positions.p, n_meas, st));
      int last_flag = 0, last_pos = 0;
      CK(c, cudaMemcpyAsync(&last_flag, merged + (n_meas - 1), 4, cudaMemcpyDeviceToHost, st));
      CK(c, cudaMemcpyAsync(&last_pos, c->positions.p + (n_meas - 1), 4, cudaMemcpyDeviceToHost, st));
      CK(c, cudaStreamSynchronize(st));
      merged_count = (int64_t) last_flag + last_pos;
    }
    append = merged_count < (int64_t) mp->target_number_of_merges;  // :96
  }
  if (append && n_meas > 0) {
    CK(c, c->flags.ensure((size_t) n_meas)); CK(c, c->positions.ensure((size_t) n_meas));
    scene_append_flag_kernel<<<blocks_for(n_meas, 256), 256, 0, st>>>(merged, me.has_valid ? me.valid.p : nullptr, n_meas, c->flags.p);
    size_t bytes = 0;
    CK(c, cub::DeviceScan::ExclusiveSum(nullptr, bytes, c->flags.p, c->positions.p, n_meas, st));
    CK(c, c->cub_tmp.ensure(bytes));
    CK(c, cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, bytes, c->flags.p, c->positions.p, n_meas, st));
    int last_flag = 0, last_pos = 0;
    CK(c, cudaMemcpyAsync(&last_flag, c->flags.p + (n_meas - 1), 4, cudaMemcpyDeviceToHost, st));
    CK(c, cudaMemcpyAsync(&last_pos, c->positions.p + (n_meas - 1), 4, cudaMemcpyDeviceToHost, st));
    CK(c, cudaStreamSynchronize(st));
    added = (int64_t) last_flag + last_pos;
    if (added > 0) {
      const size_t need = (size_t) n_scene + (size_t) added;
      CK(c, grow_preserve(sc.xyz, (size_t) n_scene * dim, need * dim, st));
      if (sc.has_normals) CK(c, grow_preserve(sc.nrm, (size_t) n_scene * dim, need * dim, st));
      if (sc.has_valid) CK(c, grow_preserve(sc.valid, (size_t) n_scene, need, st));
      float* s_nrm = sc.has_normals ? sc.nrm.p : nullptr;
      const float* m_nrm = me.has_normals ? me.nrm.p : nullptr;
      unsigned char* s_val = sc.has_valid ? sc.valid.p : nullptr;
      if (dim == 3) scene_append_kernel<3><<<blocks_for(n_meas, 256), 256, 0, st>>>(c->flags.p, c->positions.p, n_meas, me.xyz.p, m_nrm, T, n_scene, sc.xyz.p, s_nrm, s_val);
      else scene_append_kernel<2><<<blocks_for(n_meas, 256), 256, 0, st>>>(c->flags.p, c->positions.p, n_meas, me.xyz.p, m_nrm, T, n_scene, sc.xyz.p, s_nrm, s_val);
      c->launches += 2;
      sc.n = (int64_t) need;
    }
  }
  CK(c, cudaGetLastError());
  CK(c, cudaStreamSynchronize(st));
  // the scene changed under the slices clipped from it: their clips are stale (the next frame clips again anyway)
  if (n_merged) *n_merged = merged_count;
  if (n_added) *n_added = added;
  return SRRG2B_OK;
}

int srrg2b_scene_get(srrg2b_ctx* c, int scene_id, float* coords, float* normals, uint8_t* valid, int64_t* n) {
  if (!c) return SRRG2B_ERR_INVALID;
  if (!c->scenes.count(scene_id) || !c->scenes[scene_id].present) FAIL(c, SRRG2B_ERR_STATE, "unknown scene");
  CK(c, cudaSetDevice(c->device));
  RawCloud& sc = c->scenes[scene_id];
  if (n) *n = sc.n;
  const size_t cnt = (size_t) sc.n;
  if (cnt) {
    if (coords) CK(c, cudaMemcpyAsync(coords, sc.xyz.p, sizeof(float) * cnt * c->dim, cudaMemcpyDeviceToHost, c->stream));
    if (normals && sc.has_normals) CK(c, cudaMemcpyAsync(normals, sc.nrm.p, sizeof(float) * cnt * c->dim, cudaMemcpyDeviceToHost, c->stream));
    if (valid) {
      if (sc.has_valid) CK(c, cudaMemcpyAsync(valid, sc.valid.p, cnt, cudaMemcpyDeviceToHost, c->stream));
      else memset(valid, 1, cnt);
    }
  }
  CK(c, cudaStreamSynchronize(c->stream));
  return SRRG2B_OK;
}

int srrg2b_scene_clip_indices(srrg2b_ctx* c, int slice_id, int32_t* global_indices) {
  if (!c) return SRRG2B_ERR_INVALID;
  if (!c->slices.count(slice_id) || !global_indices) FAIL(c, SRRG2B_ERR_STATE, "unknown slice");
  CK(c, cudaSetDevice(c->device));
  SliceData& sd = c->slices[slice_id];
  if (sd.n_clipped > 0)
    CK(c, cudaMemcpyAsync(global_indices, sd.clip_gidx.p, sizeof(int32_t) * (size_t) sd.n_clipped, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  return SRRG2B_OK;
}

int srrg2b_find_correspondences(srrg2b_ctx* c, int slice_id, const float* S, const srrg2b_finder_params* fp,
                                int32_t* fixed_idx, int32_t* moving_idx, float* response, int64_t* n_out) {
  if (!c) return SRRG2B_ERR_INVALID;
  if (!S || !fp || !n_out) FAIL(c, SRRG2B_ERR_INVALID, "null argument");
  if (!c->slices.count(slice_id)) FAIL(c, SRRG2B_ERR_STATE, "unknown slice");
  CK(c, cudaSetDevice(c->device));
  SliceData& sd = c->slices[slice_id];
  srrg2b_factor_params fa = {SRRG2B_FACTOR_P2P, SRRG2B_ROB_NONE, 1.f, 1.f, 1.f};
  SliceArgs a;
  int rcode = fill_slice_args(c, sd, 0, *fp, fa, SRRG2B_VAR_SE3_QUAT_RIGHT, false, a, nullptr);
  if (rcode) return rcode;
  a.gate_in_nn = 1;
  a.inline_check = 1;
  Mat4f S4;
  embed(c->dim, S, S4);
  // how far any query can be from its position at the slice's ANCHOR pass (the first stand-alone pass after a
  // reset): certified bounds are stored relative to the anchor (encode_bound); a loop run in between, or new
  // clouds, reset the bounds
  float motion = 0.f, slack = 0.f;
  {
    double dr = 0.0, dt = 0.0, tn = 0.0;
    const Mat4f& P0 = sd.last_S;
    for (int r = 0; r < 3; ++r) {
      for (int cc = 0; cc < 3; ++cc) {
        const double d = sd.have_last_S ? (double) S4.m[r * 4 + cc] - (double) P0.m[r * 4 + cc] : 0.0;
        dr += d * d;
      }
      const double d = sd.have_last_S ? (double) S4.m[r * 4 + 3] - (double) P0.m[r * 4 + 3] : 0.0;
      dt += d * d;
      tn += (double) S4.m[r * 4 + 3] * (double) S4.m[r * 4 + 3];
    }
    const float radius = sqrtf(sd.radius2) * 1.0001f;
    motion = (float) ((sqrt(dr) * (double) radius + sqrt(dt)) * (1.0 + 1e-6)) * (1.f + 2.4e-7f);
    const float qmax = radius * 1.0001f + (float) sqrt(tn);
    slack = 1e-6f * qmax + 2e-7f;
    if (!sd.have_last_S) {  // no previous pass: whatever bounds exist are not trusted
      CK(c, cudaMemsetAsync(sd.c_lb.p, 0, sizeof(float) * (size_t) std::max(sd.nm_valid, 1), c->stream));
    }
  }
  set_S_kernel<<<1, 32, 0, c->stream>>>(c->d_state, 0, S4, c->track2_mode == 1 ? 1 : 0, sd.S_lb.p, motion, slack);
  c->launches++;
  CK(c, cudaMemsetAsync(a.far_count, 0, 3 * sizeof(int), c->stream));
  rcode = launch_find(c, a, nullptr);
  if (rcode) return rcode;
  commit_S_kernel<<<1, 32, 0, c->stream>>>(a.S, sd.S_lb.p);
  c->launches++;
  CK(c, cudaGetLastError());
  if (!sd.have_last_S) {  // the first stand-alone pass after a reset anchors the bounds
    sd.last_S = S4;
    sd.have_last_S = true;
  }
  sd.corr_valid = true;
  sd.stat_valid = false;
  return export_corr(c, sd, 0, false, fixed_idx, moving_idx, response, nullptr, nullptr, n_out);
}

int srrg2b_set_correspondences(srrg2b_ctx* c, int slice_id, const int32_t* fixed_idx, const int32_t* moving_idx,
                               int64_t n) {
  if (!c) return SRRG2B_ERR_INVALID;
  if (n < 0 || (n > 0 && (!fixed_idx || !moving_idx))) FAIL(c, SRRG2B_ERR_INVALID, "bad correspondence list");
  if (!c->slices.count(slice_id)) FAIL(c, SRRG2B_ERR_STATE, "unknown slice");
  CK(c, cudaSetDevice(c->device));
  SliceData& sd = c->slices[slice_id];
  if (!sd.moving_raw.present || !sd.fixed_raw.present) FAIL(c, SRRG2B_ERR_STATE, "slice needs both clouds");
  if (sd.built_for_max_distance <= 0.f) {  // any cell size will do: only the sorted SoA is needed
    int rcode = ensure_index(c, sd, 1.0f);
    if (rcode) return rcode;
  }
  if (sd.nm_valid > 0) {
    fill_int_kernel<<<blocks_for(sd.nm_valid, 256), 256, 0, c->stream>>>(sd.c_fidx.p, sd.nm_valid, -1);
    fill_int_kernel<<<blocks_for(sd.nm_valid, 256), 256, 0, c->stream>>>(sd.c_fpos.p, sd.nm_valid, -1);
    c->launches += 2;
    CK(c, cudaMemsetAsync(sd.c_lb.p, 0, sizeof(float) * (size_t) sd.nm_valid, c->stream));
  }
  if (n > 0) {
    CK(c, c->imp_f.ensure(n));
    CK(c, c->imp_m.ensure(n));
    CK(c, c->imp_bad.ensure(1));
    CK(c, cudaMemcpyAsync(c->imp_f.p, fixed_idx, 4 * (size_t) n, cudaMemcpyHostToDevice, c->stream));
    CK(c, cudaMemcpyAsync(c->imp_m.p, moving_idx, 4 * (size_t) n, cudaMemcpyHostToDevice, c->stream));
    CK(c, cudaMemsetAsync(c->imp_bad.p, 0, 4, c->stream));
    import_corr_kernel<<<blocks_for(n, 256), 256, 0, c->stream>>>(
      c->imp_f.p, c->imp_m.p, (int) n, sd.m_inverse.p, sd.f_inverse.p, (int) sd.moving_raw.n, (int) sd.fixed_raw.n,
      (int) sd.moving_raw.index_offset, sd.c_fidx.p, sd.c_fpos.p, c->imp_bad.p);
    c->launches++;
  }
  CK(c, cudaGetLastError());
  int n_bad = 0;
  if (n > 0) CK(c, cudaMemcpyAsync(&n_bad, c->imp_bad.p, 4, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  sd.have_last_S = false;
  sd.corr_valid = true;
  sd.stat_valid = false;
  if (n_bad > 0) FAIL(c, SRRG2B_ERR_INVALID, "correspondence list refers to invalid (masked-out) moving points");
  return SRRG2B_OK;
}

int srrg2b_linearize(srrg2b_ctx* c, int slice_id, const float* S, int variable, const srrg2b_finder_params* fp,
                     const srrg2b_factor_params* fa, double* H, double* b, int64_t* acc, srrg2b_iter_stats* stats,
                     uint8_t* status, float* chi) {
  if (!c) return SRRG2B_ERR_INVALID;
  if (!S || !fa || !fp) FAIL(c, SRRG2B_ERR_INVALID, "null argument");
  if (!c->slices.count(slice_id)) FAIL(c, SRRG2B_ERR_STATE, "unknown slice");
  CK(c, cudaSetDevice(c->device));
  SliceData& sd = c->slices[slice_id];
  if (!sd.corr_valid) FAIL(c, SRRG2B_ERR_STATE, "slice has no correspondences (call find / set_correspondences)");
  SliceArgs a;
  Scales sc;
  srrg2b_finder_params fpl = *fp;
  if (sd.built_for_max_distance > 0.f) {
    // the index is only used for its SoA here: do not rebuild it for a different gate / finder
    if (sd.index_is_projective) fpl = sd.proj_params;
    else { fpl.kind = SRRG2B_FINDER_NN; fpl.max_distance = sd.built_for_max_distance; }
  }
  int rcode = fill_slice_args(c, sd, 0, fpl, *fa, variable, true, a, nullptr);
  if (rcode) return rcode;
  a.gate = 0;  // the correspondences are taken as they are (gated by the finder or supplied by the caller)
  // ranges for the caller's error bound (fp->max_distance): pairs farther apart are suppressed and counted
  sc = choose_scales(c->dim, variable, fa->factor, sd.radius2, std::max(sd.nb2_moving_global, sd.fixed_raw.nb2),
                     fp->max_distance, fa->info_point, fa->info_normal);
  for (int k = 0; k < kKCount; ++k) a.fS[k] = ldexpf(1.f, sc.k[k] - 22);
  a.fSinvChi = ldexpf(1.f, 22 - sc.k[kKChi]);
  a.eb2 = sc.err_bound * sc.err_bound;
  if (!rigid_enough(c->dim, S)) FAIL(c, SRRG2B_ERR_INVALID, "S is not a rigid transform");
  Mat4f S4;
  embed(c->dim, S, S4);
  set_S_kernel<<<1, 32, 0, c->stream>>>(c->d_state, 0, S4, 0, nullptr, 0.f, 0.f);
  c->launches++;
  CK(c, cudaMemsetAsync(a.tile_ticket, 0, sizeof(int), c->stream));
  rcode = launch_linearize(c, a, fa->factor, nullptr);
  if (rcode) return rcode;
  CK(c, cudaGetLastError());
  // several ranks: the solve step (which exchanges the sums over peer memory) does not run here, so the
  // partial sums are reduced explicitly -- the result is the global sum on every rank either way
  rcode = allreduce_acc(c, 1, false);
  if (rcode) return rcode;
  rcode = fetch_state(c);
  if (rcode) return rcode;
  sd.stat_valid = true;
  const int P = c->dim == 3 ? 6 : 3;
  const unsigned long long* A = c->h_state->acc[0];
  if (acc) memcpy(acc, A, sizeof(int64_t) * kAcc);
  if (H && b) {
    int slot = 0;
    for (int i = 0; i < P; ++i)
      for (int j = i; j < P; ++j) {
        const double v = ldexp((double) (long long) A[slot++], -sc.k[(j < c->dim) ? kKHtt : ((i < c->dim) ? kKHtr : kKHrr)]);
        H[i * P + j] = v;
        H[j * P + i] = v;
      }
    for (int i = 0; i < P; ++i) b[i] = ldexp((double) (long long) A[kAccB + i], -sc.k[(i < c->dim) ? kKBt : kKBr]);
  }
  if (stats) {
    stats->iteration = 0;
    stats->solver_status = 0;
    stats->num_inliers = (int64_t) A[kAccNIn];
    stats->num_outliers = (int64_t) A[kAccNOut];
    stats->num_suppressed = (int64_t) A[kAccNSup];
    stats->num_saturated = (int64_t) A[kAccNSat];
    stats->num_correspondences = stats->num_inliers + stats->num_outliers + stats->num_suppressed;
    stats->chi_inliers = ldexp((double) (long long) A[kAccChiIn], -sc.k[kKChi]) + ldexp((double) (long long) A[kAccChiIn + 1], -sc.k[kKChiLo]);
    stats->chi_outliers = ldexp((double) (long long) A[kAccChiOut], -sc.k[kKChi]) + ldexp((double) (long long) A[kAccChiOut + 1], -sc.k[kKChiLo]);
  }
  if (status || chi) {
    int64_t m = 0;
    return export_corr(c, sd, 0, true, nullptr, nullptr, nullptr, status, chi, &m);
  }
  return SRRG2B_OK;
}

int srrg2b_icp_run(srrg2b_ctx* c, int n_slices, const srrg2b_slice* slices, const srrg2b_aligner_params* ap,
                   float* T, srrg2b_iter_stats* stats_out, int32_t* n_stats, int32_t* aligner_status) {
  if (!c) return SRRG2B_ERR_INVALID;
  RunCall r{c, n_slices, slices, ap, T, stats_out, n_stats, aligner_status};
  int rcode = run_begin(r);
  if (!rcode) rcode = run_mid(r);
  if (!rcode) rcode = run_end(r);
  return rcode;
}

// ---- N3: candidate-batched loop closing ----
int srrg2b_share_fixed(srrg2b_ctx* dst, int dst_slice_id, srrg2b_ctx* src, int src_slice_id) {
  if (!dst || !src) return SRRG2B_ERR_INVALID;
  if (dst == src && dst_slice_id == src_slice_id) FAIL(dst, SRRG2B_ERR_INVALID, "a slice cannot borrow from itself");
  if (dst->dim != src->dim || dst->device != src->device) FAIL(dst, SRRG2B_ERR_INVALID, "contexts differ in dimension or device");
  if (!src->slices.count(src_slice_id) || !src->slices[src_slice_id].fixed_raw.present)
    FAIL(dst, SRRG2B_ERR_STATE, "source slice has no fixed cloud");
  CK(dst, cudaSetDevice(dst->device));
  SliceData& a = src->slices[src_slice_id];
  // everything the source queued for this cloud (upload, index build, normal bound) has to be complete: the
  // borrower's stream does not know the source's events
  CK(dst, cudaStreamSynchronize(src->copy_stream));
  CK(dst, cudaStreamSynchronize(src->stream));
  int rcode = resolve_normal_bound(src, a.fixed_raw);
  if (rcode) { dst->err = src->err; return rcode; }
  CK(dst, cudaStreamSynchronize(dst->stream));  // a run on the borrower's previous fixed cloud may still be in flight
  SliceData& b = dst->slices[dst_slice_id];
  RawCloud& ra = a.fixed_raw;
  RawCloud& rb = b.fixed_raw;
  rb.xyz.borrow(ra.xyz); rb.nrm.borrow(ra.nrm); rb.valid.borrow(ra.valid);
  rb.n = ra.n; rb.n_global = ra.n_global; rb.index_offset = ra.index_offset;
  rb.has_normals = ra.has_normals; rb.has_valid = ra.has_valid; rb.present = true;
  rb.nb2 = ra.nb2; rb.nb2_pending = false;
  b.f_pts.borrow(a.f_pts); b.f_rec.borrow(a.f_rec); b.f_inverse.borrow(a.f_inverse);
  b.cell_start.borrow(a.cell_start); b.near_bits.borrow(a.near_bits); b.image.borrow(a.image);
  b.nf_valid = a.nf_valid;
  b.built_for_max_distance = a.built_for_max_distance;
  b.ox = a.ox; b.oy = a.oy; b.oz = a.oz; b.inv_cell = a.inv_cell;
  b.nx = a.nx; b.ny = a.ny; b.nz = a.nz; b.R = a.R; b.xbits = a.xbits; b.nx_coarse = a.nx_coarse; b.xf = a.xf;
  b.cached_R = a.cached_R; b.cached_R_n = a.cached_R_n; b.cached_R_md = a.cached_R_md;
  b.last_max_distance = a.last_max_distance;
  b.index_is_projective = a.index_is_projective;
  b.proj_params = a.proj_params;
  b.corr_valid = false;
  b.have_last_S = false;
  return SRRG2B_OK;
}

int srrg2b_closure_batch(srrg2b_ctx* const* ctxs, int k, int n_slices, const srrg2b_slice* slices,
                         const srrg2b_aligner_params* ap, const float* guesses, const srrg2b_closure_params* cp,
                         srrg2b_closure_result* results) {
  if (!ctxs || k < 0) return SRRG2B_ERR_INVALID;
  if (k == 0) return SRRG2B_OK;
  for (int i = 0; i < k; ++i) if (!ctxs[i]) return SRRG2B_ERR_INVALID;
  srrg2b_ctx* c0 = ctxs[0];
  if (!slices || !ap || !guesses || !cp || !results) FAIL(c0, SRRG2B_ERR_INVALID, "null argument");
  for (int i = 0; i < k; ++i) {
    if (ctxs[i]->dim != c0->dim) FAIL(c0, SRRG2B_ERR_INVALID, "contexts differ in dimension");
    // candidates are independent: several GPUs take disjoint runs of them (sharding.py: candidate_shard), they do not
    // shard one candidate's cloud -- a context that belongs to a communicator would wait for its peers in every solve step
    if (ctxs[i]->world > 1) FAIL(c0, SRRG2B_ERR_INVALID, "a context of the batch belongs to a multi-rank communicator");
    for (int j = 0; j < i; ++j) if (ctxs[j] == ctxs[i]) FAIL(c0, SRRG2B_ERR_INVALID, "a context appears twice in the batch");
  }
  const int D1 = c0->dim + 1, DD = D1 * D1;
  std::vector<RunCall> calls;
  std::vector<std::array<float, 16>> T(k);
  std::vector<int32_t> status(k, SRRG2B_ALIGNER_FAIL);
  calls.reserve(k);
  for (int i = 0; i < k; ++i) {
    for (int e = 0; e < DD; ++e) T[i][e] = guesses[(size_t) i * DD + e];
    calls.push_back(RunCall{ctxs[i], n_slices, slices, ap, T[i].data(), nullptr, nullptr, &status[i]});
  }
  // the three host steps of compute(), each over all candidates: while the host waits for candidate i, the runs of
  // the candidates behind it are executing
  int first_error = SRRG2B_OK;
  auto note = [&](int i, int rcode) {
    if (rcode && !first_error) { first_error = rcode; if (ctxs[i] != c0) c0->err = ctxs[i]->err; }
    return rcode;
  };
  std::vector<char> alive(k, 1);
  for (int i = 0; i < k; ++i) if (note(i, run_begin(calls[i]))) alive[i] = 0;
  for (int i = 0; i < k; ++i) if (alive[i] && note(i, run_mid(calls[i]))) alive[i] = 0;
  for (int i = 0; i < k; ++i) if (alive[i] && note(i, run_end(calls[i]))) alive[i] = 0;
  if (first_error) {  // leave no run in flight behind an error
    for (int i = 0; i < k; ++i) cudaStreamSynchronize(ctxs[i]->stream);
    return first_error;
  }
  int n_prior = 0;
  for (int s = 0; s < n_slices; ++s) n_prior += slices[s].kind != SRRG2B_SLICE_POINTS;
  for (int i = 0; i < k; ++i) {
    srrg2b_ctx* c = ctxs[i];
    srrg2b_closure_result& r = results[i];
    memset(&r, 0, sizeof(r));
    const DevState* hs = c->h_state;
    r.aligner_status = status[i];
    r.iterations = hs->n_stats;
    r.device_ms = c->last_ms;
    for (int e = 0; e < DD; ++e) r.moving_in_fixed[e] = T[i][e];
    if (status[i] != SRRG2B_ALIGNER_SUCCESS) {  // :80-84
      r.verdict = SRRG2B_CLOSURE_ALIGNER_DROP;
      continue;
    }
    const srrg2b_iter_stats& is = hs->last_stats;  // iterationStats().back(), :86
    int64_t ncorr = n_prior;  // AlignerSliceProcessorPrior_::numCorrespondences() == 1
    for (int s = 0; s < n_slices; ++s) {
      if (slices[s].kind != SRRG2B_SLICE_POINTS) continue;
      SliceData& sd = c->slices[slices[s].slice_id];
      int64_t m = 0;
      const int rcode = export_corr(c, sd, sd.prune_on_export, false, nullptr, nullptr, nullptr, nullptr, nullptr, &m);
      if (rcode) { if (c != c0) c0->err = c->err; return rcode; }
      ncorr += m;
    }
    r.num_correspondences = ncorr;
    r.num_inliers = is.num_inliers;
    r.chi_inliers = (float) is.chi_inliers / (float) is.num_inliers;  // :91 (IterationStats holds floats upstream)
    if (is.num_inliers < (int64_t) cp->relocalize_min_inliers) { r.verdict = SRRG2B_CLOSURE_NUM_INLIERS_DROP; continue; }
    if (r.chi_inliers > cp->relocalize_max_chi_inliers) { r.verdict = SRRG2B_CLOSURE_MAX_CHI_DROP; continue; }
    const float ratio = (float) is.num_inliers / (float) ncorr;  // :105
    if (ratio < cp->relocalize_min_inliers_ratio) { r.verdict = SRRG2B_CLOSURE_INLIER_RATIO_DROP; continue; }
    r.verdict = SRRG2B_CLOSURE_ACCEPT;
  }
  return SRRG2B_OK;
}

int srrg2b_icp_iterate(srrg2b_ctx* c, int n_slices, const srrg2b_slice* slices, int variable, float* T,
                       srrg2b_iter_stats* stats, int32_t* association_good) {
  if (!c) return SRRG2B_ERR_INVALID;
  if (!T) FAIL(c, SRRG2B_ERR_INVALID, "null argument");
  int rcode = validate_slices(c, n_slices, slices, variable);
  if (rcode) return rcode;
  CK(c, cudaSetDevice(c->device));
  srrg2b_aligner_params ap;
  memset(&ap, 0, sizeof(ap));
  ap.variable = variable;
  ap.max_iterations = 1;
  ap.window_size = 5;
  Plan plan;
  rcode = make_plan(c, n_slices, slices, ap, false, true, plan);
  if (rcode) return rcode;
  if (!rigid_enough(c->dim, T)) FAIL(c, SRRG2B_ERR_INVALID, "T is not a rigid transform");
  Mat4f T0;
  embed(c->dim, T, T0);
  for (int s = 0; s < n_slices; ++s)
    if (slices[s].kind == SRRG2B_SLICE_POINTS) c->slices[slices[s].slice_id].have_last_S = false;
  rcode = run_plan(c, plan, T0, 1, 0, 1, 0);
  if (rcode) return rcode;
  rcode = fetch_state(c);
  if (rcode) return rcode;
  unembed(c->dim, c->h_state->X, T);
  if (association_good) *association_good = c->h_state->not_enough_corr ? 0 : 1;
  if (stats) {
    if (c->h_state->n_stats > 0) *stats = c->h_state->stats[0];
    else memset(stats, 0, sizeof(*stats));
  }
  for (int s = 0; s < n_slices; ++s) {
    if (slices[s].kind != SRRG2B_SLICE_POINTS) continue;
    SliceData& sd = c->slices[slices[s].slice_id];
    sd.corr_valid = true;
    sd.stat_valid = true;
    sd.prune_on_export = 0;
  }
  return SRRG2B_OK;
}

int srrg2b_get_correspondences(srrg2b_ctx* c, int slice_id, int32_t* fixed_idx, int32_t* moving_idx, float* response,
                               int64_t* n_out) {
  if (!c) return SRRG2B_ERR_INVALID;
  if (!n_out) FAIL(c, SRRG2B_ERR_INVALID, "null argument");
  if (!c->slices.count(slice_id)) FAIL(c, SRRG2B_ERR_STATE, "unknown slice");
  CK(c, cudaSetDevice(c->device));
  SliceData& sd = c->slices[slice_id];
  return export_corr(c, sd, sd.prune_on_export, false, fixed_idx, moving_idx, response, nullptr, nullptr, n_out);
}

int srrg2b_reset_correspondences(srrg2b_ctx* c, int slice_id) {
  if (!c) return SRRG2B_ERR_INVALID;
  if (!c->slices.count(slice_id)) FAIL(c, SRRG2B_ERR_STATE, "unknown slice");
  CK(c, cudaSetDevice(c->device));
  SliceData& sd = c->slices[slice_id];
  if (sd.moving_raw.present && sd.nm_valid > 0) {
    fill_int_kernel<<<blocks_for(sd.nm_valid, 256), 256, 0, c->stream>>>(sd.c_fpos.p, sd.nm_valid, -1);
    fill_int_kernel<<<blocks_for(sd.nm_valid, 256), 256, 0, c->stream>>>(sd.c_fidx.p, sd.nm_valid, -1);
    c->launches += 2;
    CK(c, cudaMemsetAsync(sd.c_lb.p, 0, sizeof(float) * (size_t) sd.nm_valid, c->stream));
    CK(c, cudaGetLastError());
  }
  CK(c, cudaStreamSynchronize(c->stream));
  sd.corr_valid = false;
  sd.stat_valid = false;
  sd.have_last_S = false;
  return SRRG2B_OK;
}

int srrg2b_last_run_timing(srrg2b_ctx* c, float* device_ms, int32_t* iterations) {
  if (!c) return SRRG2B_ERR_INVALID;
  if (device_ms) *device_ms = c->last_ms;
  if (iterations) *iterations = c->last_iterations;
  return SRRG2B_OK;
}

int srrg2b_debug_info(srrg2b_ctx* c, int slice_id, int32_t* out16) {
  if (!c || !out16) return SRRG2B_ERR_INVALID;
  if (!c->slices.count(slice_id)) FAIL(c, SRRG2B_ERR_STATE, "unknown slice");
  CK(c, cudaSetDevice(c->device));
  SliceData& sd = c->slices[slice_id];
  memset(out16, 0, 16 * sizeof(int32_t));
  out16[0] = sd.R; out16[1] = sd.nx; out16[2] = sd.ny; out16[3] = sd.nz;
  out16[4] = sd.nf_valid; out16[5] = sd.nm_valid;
  if (sd.far_count.p) {
    int two[2] = {0, 0};
    CK(c, cudaMemcpyAsync(two, sd.far_count.p, 8, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    out16[6] = two[0];
    out16[8] = two[1];
  }
  float cell = 1.f / sd.inv_cell;
  memcpy(&out16[7], &cell, 4);
  return SRRG2B_OK;
}

int srrg2b_set_kernel_timing(srrg2b_ctx* c, int enable) {
  if (!c) return SRRG2B_ERR_INVALID;
  c->time_kernels = enable != 0;
  c->kev_used = 0;
  return SRRG2B_OK;
}

int srrg2b_last_kernel_timing(srrg2b_ctx* c, float* slice_kernel_ms, int32_t* slice_kernel_launches) {
  if (!c) return SRRG2B_ERR_INVALID;
  if (slice_kernel_ms) *slice_kernel_ms = c->last_kernel_ms;
  if (slice_kernel_launches) *slice_kernel_launches = c->last_kernel_launches;
  return SRRG2B_OK;
}

}  // extern "C"

#include "s2b_pgo_host.inl"
