// fclgpu_api.cu — C ABI (include/fclgpu.h): model upload, batched collide / distance,
// contact compaction, host-pointer wrappers and micro-benchmarks.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/fclgpu.h"
#include "bvh_build.hpp"
#include "records.hpp"
#include "traversal.cuh"
#include "collide_ordered.cuh"
#include "broadphase.cuh"
#include "refit.cuh"
#include "continuous.cuh"
#include "query_order.cuh"

using namespace fclgpu;

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------
namespace {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY(expr)                                                                      \
  do {                                                                                      \
    cudaError_t e__ = (expr);                                                               \
    if (e__ != cudaSuccess)                                                                 \
      return fail(e__ == cudaErrorNoDevice || e__ == cudaErrorInsufficientDriver            \
                      ? FCLGPU_ERR_NO_DEVICE                                                \
                      : (e__ == cudaErrorMemoryAllocation ? FCLGPU_ERR_MODEL_OUT_OF_MEMORY  \
                                                          : FCLGPU_ERR_CUDA),               \
                  "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

inline int cuda_status(cudaError_t e) {
  return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver
             ? FCLGPU_ERR_NO_DEVICE
             : (e == cudaErrorMemoryAllocation ? FCLGPU_ERR_MODEL_OUT_OF_MEMORY : FCLGPU_ERR_CUDA);
}

// options
std::mutex g_opt_mu;
std::map<std::string, long long> g_opts = {
    // 4: like 3 with the pooled collide kernel testing both children per BV round (chosen automatically for big BVHs)
    // 3 (default): like 2, and collide with contact generation pools the warp's deferred triangle pairs
    //    over all 32 lanes (collide_pooled_kernel)
    // 2: like 1, but the BV tests that steer the traversal are conservative single-precision tests
    //    (bounds_f32.cuh); every result still comes from the exact FP64 triangle routines
    // 1: collide = thread per query with deferred leaf rounds, distance = warp per query sorted front, FP64 BV tests
    // 0: thread per query, the reference's visiting order exactly (work counters match the reference's)
    {"traversal", 3},
    {"dist_spill_entries", 4096},  // per-warp global overflow entries of the distance front (allocated for big models)
    {"collide_front", 1},      // counts-only collide: warp-per-query front kernel (0 never, 1 big BVHs and small batches, 2 always)
    {"front_small_batch", 131072},  // ... batches of at most this many queries count as small
    {"front_leaf_trigger", 0},  // front kernel: triangle pairs queued before a leaf round while the query can still saturate (0 = 8 for BVHs beyond the caches, else 32)
    {"front_cap", 0},          // front kernel: stack entries per warp (0 = default, 384; measured 256 .. 2048)
    {"host_chunk", 1 << 17},   // queries per stage of the host API's two-stream copy/compute pipeline
    {"contact_stride", 1024},  // contact staging slots per query (per resident warp on the ordered-front path) when num_max_contacts is larger
    // layout of a contact list: 0 (default) = blocks appended in completion order by the one-launch ordered-front kernel
    // (contact_offsets[i] = start of query i's block); 1 = blocks in query order (contact_offsets = exclusive prefix sum
    // of num_contacts): per-query scratch + scan + compaction passes around the lane-per-query kernels
    {"contact_order", 0},
    // evaluate the queries of a device-resident batch in a locality order (query_order.cuh) on the lane-per-query collide
    // path: 0 never (default), else for batches of at least this many queries.  Measured on 1M random poses:
    // slower -- a million random 6-DOF poses are too sparse to be coherent, and coherent stretches have correlated cost
    {"order_queries", 0},
    // ordered-front path, both measured and left off: discard the L2 lines of the contact staging once a
    // query's list is in the pool (DRAM writes 2.11 -> 2.05 GB per 1M queries, time unchanged); stage short lists in
    // shared memory instead of the global per-warp scratch (DRAM writes 2.04 GB, but the smaller L1 costs 1.7 %)
    {"contact_discard", 0},
    {"contact_smem_stage", 0},
    {"scratch_bytes", 2ll << 30},
    {"blocks_per_sm", 0},      // 0 = occupancy query
    {"stats", 1},
    // binary-mode collide at traversal >= 3: 1 = pooled kernel (measured fastest: 3.8 ms per 1M poses),
    // 2 = pooled kernel with the FP32 triangle classification (4.2 ms), 0 = deferred kernel with classification (4.9 ms)
    {"binary_pooled", 1},
    {"refit_warp", 2},         // on-device refit / build: 2 = block-cooperative fit for nodes > 2048 triangles and
                               // warp-cooperative for nodes > 24, 1 = warp-cooperative only, 0 = one thread per node
    {"pool_trigger", 32},      // collide variant P: queued pairs in the warp that trigger a pooled leaf round
    {"leaf_trigger", 20},      // collide variant D: lanes with queued triangle pairs that trigger a leaf round
    {"sphere_blocks", 4},         // mesh <-> sphere distance: resident blocks per SM the kernel is compiled for (3, 4, 5)
    {"sphere_bound32", 1},        // mesh <-> sphere distance: box bound from the 64-byte FP32 records (0 = FP64 records)
    {"sphere_leaf_trigger", 16},  // mesh <-> sphere distance: parked lanes that trigger a leaf round (0 = leaf tests inline)
};
long long opt(const char* k) {
  std::lock_guard<std::mutex> g(g_opt_mu);
  auto it = g_opts.find(k);
  return it == g_opts.end() ? 0 : it->second;
}

// ------------------------------------------------------------------------------------------
// per-device workspace: work counters, sticky status, scan temporaries, contact scratch
// ------------------------------------------------------------------------------------------
struct Workspace {
  int device = -1;
  int sm_count = 0;
  unsigned long long* counters = nullptr;  // ring of work counters (one per launch in flight)
  int counter_slots = 0, counter_next = 0;
  // Sticky status word and running contact total PER STREAM (slot 0 doubles as the overflow slot when more than
  // kStreamSlots streams are in use): concurrent *_batch calls on different streams neither mix their contact offsets
  // nor report / clear each other's errors.
  int* status_pool = nullptr;
  long long* cursor_pool = nullptr;
  std::map<cudaStream_t, int> stream_slot;
  // The contact scratch and the distance front's overflow area are one allocation per device: launches that use them
  // on DIFFERENT streams are ordered with an event (same-stream launches are ordered anyway).
  cudaEvent_t scratch_ev = nullptr, order_ev = nullptr;
  cudaStream_t scratch_stream = nullptr, order_stream = nullptr;
  bool scratch_used = false, order_used = false;
  void* order_buf = nullptr;  // locality order of a batch (query_order.cuh): keys, bucket counters, the order itself
  size_t order_bytes = 0;
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
  void* scan_tmp = nullptr;
  size_t scan_bytes = 0;
  // host-API staging (device side) and the two streams of the chunked copy/compute pipeline
  void* dev_io = nullptr;
  size_t dev_io_bytes = 0;
  cudaStream_t pipe[2] = {nullptr, nullptr};
  // distance kernel: per-warp overflow areas of the sorted front (deep trees).  An area is indexed by the launch-local warp
  // id, so two launches may share one only one after the other; there are two, so that the launches of the host API's two
  // pipeline streams overlap (the tail of a launch -- its longest queries -- runs next to the start of the following one)
  struct SpillArea {
    uint2* pair = nullptr;
    float* bound = nullptr;
    int cap = 0;
    cudaEvent_t ev = nullptr;
    bool used = false;
    cudaStream_t stream = nullptr;
  } spill[2];
  unsigned spill_next = 0;
  unsigned* ready = nullptr;      // device: per-chunk "input has landed" flags of the streamed host path
  unsigned* host_one = nullptr;   // pinned host word (= 1) the copy stream writes into ready[c]
  long long* host_totals = nullptr;  // pinned: running contact totals read back per sub-batch
  cudaEvent_t ev[2] = {nullptr, nullptr};
  std::mutex mu;       // held while enqueuing (work counters, scratch growth)
  std::mutex host_mu;  // held for a whole *_batch_host call: the staging buffers and pipeline streams are shared
};
constexpr int kStreamSlots = 64;
constexpr int kReadySlots = 1 << 16;
constexpr int kTotalSlots = 1 << 12;

std::mutex g_ws_mu;
std::map<int, Workspace*> g_ws;

int get_ws(int device, Workspace** out) {
  std::lock_guard<std::mutex> g(g_ws_mu);
  auto it = g_ws.find(device);
  if (it != g_ws.end()) {
    *out = it->second;
    return 0;
  }
  CUDA_TRY(cudaSetDevice(device));
  Workspace* w = new Workspace;
  w->device = device;
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  w->sm_count = prop.multiProcessorCount;
  w->counter_slots = 4096;
  CUDA_TRY(cudaMalloc(&w->counters, sizeof(unsigned long long) * w->counter_slots));
  CUDA_TRY(cudaMemset(w->counters, 0, sizeof(unsigned long long) * w->counter_slots));
  CUDA_TRY(cudaMalloc(&w->status_pool, sizeof(int) * kStreamSlots));
  CUDA_TRY(cudaMemset(w->status_pool, 0, sizeof(int) * kStreamSlots));
  CUDA_TRY(cudaMalloc(&w->cursor_pool, sizeof(long long) * kStreamSlots));
  CUDA_TRY(cudaMemset(w->cursor_pool, 0, sizeof(long long) * kStreamSlots));
  CUDA_TRY(cudaEventCreateWithFlags(&w->scratch_ev, cudaEventDisableTiming));
  for (auto& a : w->spill) CUDA_TRY(cudaEventCreateWithFlags(&a.ev, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&w->order_ev, cudaEventDisableTiming));
  CUDA_TRY(cudaMalloc(&w->ready, sizeof(unsigned) * kReadySlots));
  CUDA_TRY(cudaMemset(w->ready, 0, sizeof(unsigned) * kReadySlots));
  CUDA_TRY(cudaHostAlloc((void**)&w->host_one, sizeof(unsigned), cudaHostAllocDefault));
  *w->host_one = 1u;
  CUDA_TRY(cudaHostAlloc((void**)&w->host_totals, sizeof(long long) * kTotalSlots, cudaHostAllocDefault));
  CUDA_TRY(cudaEventCreateWithFlags(&w->ev[0], cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&w->ev[1], cudaEventDisableTiming));
  CUDA_TRY(cudaStreamCreateWithFlags(&w->pipe[0], cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&w->pipe[1], cudaStreamNonBlocking));
  g_ws[device] = w;
  *out = w;
  return 0;
}

// per-stream slot (caller holds w->mu)
struct StreamState {
  int* status;
  long long* cursor;
};
StreamState stream_state(Workspace* w, cudaStream_t st) {
  auto it = w->stream_slot.find(st);
  int slot;
  if (it != w->stream_slot.end()) {
    slot = it->second;
  } else {
    slot = (int)w->stream_slot.size() + 1 < kStreamSlots ? (int)w->stream_slot.size() + 1 : 0;
    w->stream_slot[st] = slot;
  }
  return StreamState{w->status_pool + slot, w->cursor_pool + slot};
}
// order a launch that uses a per-device buffer after the previous user of that buffer on ANOTHER stream
void order_after(cudaStream_t st, cudaEvent_t ev, bool used, cudaStream_t last) {
  if (used && last != st) cudaStreamWaitEvent(st, ev, 0);
}
void mark_use(cudaStream_t st, cudaEvent_t ev, bool& used, cudaStream_t& last) {
  cudaEventRecord(ev, st);
  used = true;
  last = st;
}

size_t padded_up(size_t bytes) { return (bytes + 255) & ~size_t(255); }
int ensure(void** p, size_t* have, size_t want) {
  if (*have >= want) return 0;
  if (*p) CUDA_TRY(cudaFree(*p));
  *p = nullptr;
  *have = 0;
  CUDA_TRY(cudaMalloc(p, want));
  *have = want;
  return 0;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// model
// ------------------------------------------------------------------------------------------
struct fclgpu_model {
  int device;
  DeviceModel d;
  double *obb, *rss, *tri;
  RssRec32* rss32;
  ObbRec32* obb32;
  double2* topo;
  // refit topology (optional)
  int32_t num_vertices = 0;
  int32_t *tri_index = nullptr, *node_first = nullptr, *node_count = nullptr, *by_size = nullptr;
  int32_t n_big = 0;   // nodes with more than 24 triangles (front of by_size): a warp each
  int32_t n_huge = 0;  // of those, nodes with more than 2048 triangles: a block each
  uint32_t* prim_order = nullptr;
  double* vert_stage = nullptr;
  int32_t* fc;
  int depth;
  // bottom-up refit schedule (built on first use): node ids grouped by tree height, leaves first
  int32_t* by_height = nullptr;
  std::vector<int32_t> height_start;  // group h = by_height[height_start[h] .. height_start[h + 1])
  LocalAabb aabb;  // BVHModel::computeLocalAABB over the vertices the triangles reference
};

namespace {
// BVHModel::computeLocalAABB (BVH_model-inl.h:1080-1100) over the vertices referenced by the triangles: aabb_local,
// aabb_center = (min + max) * 0.5, aabb_radius = sqrt(max |center - v|^2).  vertex(t, k) -> pointer to 3 doubles.
template <class VertexOf>
void local_aabb_of(int n_tris, VertexOf vertex, LocalAabb& a) {
  const double big = 1.7976931348623157e308;
  for (int k = 0; k < 3; ++k) {
    a.mn[k] = big;
    a.mx[k] = -big;
  }
  for (int t = 0; t < n_tris; ++t)
    for (int c = 0; c < 3; ++c) {
      const double* v = vertex(t, c);
      for (int k = 0; k < 3; ++k) {
        a.mn[k] = std::min(a.mn[k], v[k]);
        a.mx[k] = std::max(a.mx[k], v[k]);
      }
    }
  for (int k = 0; k < 3; ++k) a.c[k] = (a.mn[k] + a.mx[k]) * 0.5;
  double r2 = 0;
  for (int t = 0; t < n_tris; ++t)
    for (int c = 0; c < 3; ++c) {
      const double* v = vertex(t, c);
      const double dx = a.c[0] - v[0], dy = a.c[1] - v[1], dz = a.c[2] - v[2];
      const double r = (dx * dx + dy * dy) + dz * dz;
      if (r > r2) r2 = r;
    }
  a.r = std::sqrt(r2);
}

int tree_depth(const int32_t* fc, int n) {
  std::vector<std::pair<int, int>> st;
  st.emplace_back(0, 0);
  int mx = 0;
  while (!st.empty()) {
    auto [i, d] = st.back();
    st.pop_back();
    mx = std::max(mx, d);
    if (fc[i] >= 0) {
      if (fc[i] + 1 >= n) return -1;
      st.emplace_back(fc[i], d + 1);
      st.emplace_back(fc[i] + 1, d + 1);
    }
  }
  return mx;
}
}  // namespace

extern "C" int fclgpu_model_create_obbrss(int device, int32_t n_nodes, const int32_t* first_child,
                                          const double* axis9, const double* obb_To3,
                                          const double* obb_extent3, const double* rss_To3,
                                          const double* rss_l2, const double* rss_r, int32_t n_tris,
                                          const double* tri_verts9, fclgpu_model** out) {
  return fclgpu_model_create_obbrss2(device, n_nodes, first_child, axis9, obb_To3, obb_extent3, nullptr, rss_To3, rss_l2, rss_r,
                                     n_tris, tri_verts9, out);
}

// rss_axis9 == NULL: the RSS shares the OBB's axes (true for every tree built by endModel() or refitted top-down,
// BV_fitter-inl.h:464).  After the reference's BOTTOM-UP refit the two differ (OBBRSS::operator+ merges the OBB and
// the RSS separately, OBBRSS-inl.h:95-101): such a model must be uploaded with both.
extern "C" int fclgpu_model_create_obbrss2(int device, int32_t n_nodes, const int32_t* first_child, const double* axis9,
                                           const double* obb_To3, const double* obb_extent3, const double* rss_axis9,
                                           const double* rss_To3, const double* rss_l2, const double* rss_r, int32_t n_tris,
                                           const double* tri_verts9, fclgpu_model** out) {
  if (!out) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "out is NULL");
  *out = nullptr;
  if (n_nodes <= 0 || n_tris <= 0) return fail(FCLGPU_ERR_BUILD_EMPTY_MODEL, "empty model");
  if (!first_child || !axis9 || !obb_To3 || !obb_extent3 || !rss_To3 || !rss_l2 || !rss_r || !tri_verts9)
    return fail(FCLGPU_ERR_INVALID_ARGUMENT, "NULL node array");
  if (n_nodes != 2 * n_tris - 1) return fail(FCLGPU_ERR_INCORRECT_DATA, "n_nodes != 2*n_tris-1");
  for (int i = 0; i < n_nodes; ++i) {
    const int fc = first_child[i];
    if (fc >= 0 ? (fc + 1 >= n_nodes || fc <= i) : (-(fc + 1) >= n_tris))
      return fail(FCLGPU_ERR_INCORRECT_DATA, "first_child[%d]=%d out of range", i, fc);
  }
  const int depth = tree_depth(first_child, n_nodes);
  if (depth < 0) return fail(FCLGPU_ERR_INCORRECT_DATA, "malformed tree");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return fail(FCLGPU_ERR_NO_DEVICE, "no CUDA device: %s", cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "device %d out of range", device);
  CUDA_TRY(cudaSetDevice(device));

  // pack the 128-byte node records and 80-byte triangle records
  std::vector<double> obb((size_t)n_nodes * kNodeDoubles), rss((size_t)n_nodes * kNodeDoubles);
  for (int i = 0; i < n_nodes; ++i) {
    double* o = &obb[(size_t)i * kNodeDoubles];
    double* r = &rss[(size_t)i * kNodeDoubles];
    for (int k = 0; k < 9; ++k) {
      o[k] = axis9[9 * (size_t)i + k];
      r[k] = (rss_axis9 ? rss_axis9 : axis9)[9 * (size_t)i + k];
    }
    for (int k = 0; k < 3; ++k) {
      o[9 + k] = obb_To3[3 * (size_t)i + k];
      o[12 + k] = obb_extent3[3 * (size_t)i + k];
      r[9 + k] = rss_To3[3 * (size_t)i + k];
    }
    r[12] = rss_l2[2 * (size_t)i];
    r[13] = rss_l2[2 * (size_t)i + 1];
    r[14] = rss_r[i];
    // OBB::size() = extent.squaredNorm(), summed left to right
    const double size = (o[12] * o[12] + o[13] * o[13]) + o[14] * o[14];
    o[15] = r[15] = size;
  }
  std::vector<double> tri((size_t)n_tris * kTriDoubles, 0.0);
  for (int t = 0; t < n_tris; ++t)
    for (int k = 0; k < 9; ++k) tri[(size_t)t * kTriDoubles + k] = tri_verts9[9 * (size_t)t + k];

  fclgpu_model* m = new fclgpu_model{};
  m->device = device;
  m->depth = depth;
  auto up = [&](double** dst, const std::vector<double>& src) -> int {
    CUDA_TRY(cudaMalloc((void**)dst, src.size() * sizeof(double)));
    CUDA_TRY(cudaMemcpy(*dst, src.data(), src.size() * sizeof(double), cudaMemcpyHostToDevice));
    return 0;
  };
  int rc;
  if ((rc = up(&m->obb, obb)) || (rc = up(&m->rss, rss)) || (rc = up(&m->tri, tri))) {
    fclgpu_model_destroy(m);
    return rc;
  }
  {
    std::vector<RssRec32> r32(n_nodes);
    for (int i = 0; i < n_nodes; ++i)
      pack_rss32((rss_axis9 ? rss_axis9 : axis9) + 9 * (size_t)i, rss_To3 + 3 * (size_t)i, rss_l2 + 2 * (size_t)i, rss_r[i], r32[i]);
    if (cudaMalloc((void**)&m->rss32, sizeof(RssRec32) * n_nodes) != cudaSuccess ||
        cudaMemcpy(m->rss32, r32.data(), sizeof(RssRec32) * n_nodes, cudaMemcpyHostToDevice) != cudaSuccess) {
      fclgpu_model_destroy(m);
      return fail(FCLGPU_ERR_MODEL_OUT_OF_MEMORY, "rss32 upload failed");
    }
  }
  {
    std::vector<ObbRec32> o32(n_nodes);
    std::vector<double2> topo(n_nodes);
    for (int i = 0; i < n_nodes; ++i) {
      pack_obb32(axis9 + 9 * (size_t)i, obb_To3 + 3 * (size_t)i, obb_extent3 + 3 * (size_t)i, o32[i]);
      long long bits = (long long)(unsigned)first_child[i];  // low word = first_child
      std::memcpy(&topo[i].x, &bits, 8);
      topo[i].y = obb[(size_t)i * kNodeDoubles + 15];
    }
    if (cudaMalloc((void**)&m->obb32, sizeof(ObbRec32) * n_nodes) != cudaSuccess ||
        cudaMemcpy(m->obb32, o32.data(), sizeof(ObbRec32) * n_nodes, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMalloc((void**)&m->topo, sizeof(double2) * n_nodes) != cudaSuccess ||
        cudaMemcpy(m->topo, topo.data(), sizeof(double2) * n_nodes, cudaMemcpyHostToDevice) != cudaSuccess) {
      fclgpu_model_destroy(m);
      return fail(FCLGPU_ERR_MODEL_OUT_OF_MEMORY, "obb32/topo upload failed");
    }
  }
  if (cudaMalloc((void**)&m->fc, sizeof(int32_t) * n_nodes) != cudaSuccess ||
      cudaMemcpy(m->fc, first_child, sizeof(int32_t) * n_nodes, cudaMemcpyHostToDevice) != cudaSuccess) {
    fclgpu_model_destroy(m);
    return fail(FCLGPU_ERR_MODEL_OUT_OF_MEMORY, "first_child upload failed");
  }
  m->d = DeviceModel{m->obb, m->rss, m->fc, m->tri, m->rss32, m->obb32, m->topo, n_nodes, n_tris};
  local_aabb_of(n_tris, [&](int t, int c) { return tri_verts9 + 9 * (size_t)t + 3 * c; }, m->aabb);
  *out = m;
  return FCLGPU_OK;
}

namespace {
// Refit schedule: node ids with the huge nodes (> 2048 triangles, a block each) first, then the big ones
// (> 24, a warp each), then the rest (a thread each); huge and big nodes largest first.  O(n) apart from the
// sort of the few large nodes.
void refit_schedule(const int32_t* count, int nn, std::vector<int32_t>& by_size, int& n_huge, int& n_big) {
  by_size.clear();
  by_size.reserve(nn);
  for (int i = 0; i < nn; ++i)
    if (count[i] > 24) by_size.push_back(i);
  std::stable_sort(by_size.begin(), by_size.end(), [&](int a, int b) { return count[a] > count[b]; });
  n_big = (int)by_size.size();
  n_huge = 0;
  while (n_huge < n_big && count[by_size[n_huge]] > 2048) ++n_huge;
  for (int i = 0; i < nn; ++i)
    if (count[i] <= 24) by_size.push_back(i);
}
}  // namespace

extern "C" int fclgpu_model_set_partition(fclgpu_model* m, int32_t num_vertices, const int32_t* tri_indices3,
                                          const int32_t* first_primitive, const int32_t* num_primitives,
                                          const int32_t* primitive_indices) {
  if (!m || !tri_indices3 || !first_primitive || !num_primitives || !primitive_indices || num_vertices <= 0)
    return fail(FCLGPU_ERR_INVALID_ARGUMENT, "NULL partition array");
  const int nn = m->d.n_nodes, nt = m->d.n_tris;
  long long covered = 0;
  for (int i = 0; i < nn; ++i) {
    if (first_primitive[i] < 0 || num_primitives[i] <= 0 || first_primitive[i] + num_primitives[i] > nt)
      return fail(FCLGPU_ERR_INCORRECT_DATA, "node %d: primitive range out of bounds", i);
    if (num_primitives[i] == 1) covered++;
  }
  if (covered != nt) return fail(FCLGPU_ERR_INCORRECT_DATA, "partition does not have one leaf per triangle");
  for (long long i = 0; i < 3ll * nt; ++i)
    if (tri_indices3[i] < 0 || tri_indices3[i] >= num_vertices) return fail(FCLGPU_ERR_INCORRECT_DATA, "triangle index out of range");
  std::vector<int32_t> by_size;
  refit_schedule(num_primitives, nn, by_size, m->n_huge, m->n_big);
  CUDA_TRY(cudaSetDevice(m->device));
  auto up = [&](int32_t** dst, const void* src, size_t count) -> int {
    if (*dst) CUDA_TRY(cudaFree(*dst));
    *dst = nullptr;
    CUDA_TRY(cudaMalloc((void**)dst, count * sizeof(int32_t)));
    CUDA_TRY(cudaMemcpy(*dst, src, count * sizeof(int32_t), cudaMemcpyHostToDevice));
    return 0;
  };
  int rc;
  if ((rc = up(&m->tri_index, tri_indices3, 3 * (size_t)nt)) || (rc = up(&m->node_first, first_primitive, nn)) ||
      (rc = up(&m->node_count, num_primitives, nn)) || (rc = up(&m->by_size, by_size.data(), nn)) ||
      (rc = up((int32_t**)&m->prim_order, primitive_indices, nt)))
    return rc;
  if (m->vert_stage) CUDA_TRY(cudaFree(m->vert_stage));
  m->vert_stage = nullptr;
  CUDA_TRY(cudaMalloc((void**)&m->vert_stage, 3 * (size_t)num_vertices * sizeof(double)));
  m->num_vertices = num_vertices;
  return FCLGPU_OK;
}

extern "C" int fclgpu_model_from_bvh(int device, const fclgpu_bvh* bvh, fclgpu_model** out) {
  if (!bvh) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "bvh is NULL");
  const int nn = fclgpu_bvh_num_nodes(bvh), nt = fclgpu_bvh_num_tris(bvh);
  std::vector<int32_t> fc(nn);
  std::vector<double> axis(9 * (size_t)nn), oT(3 * (size_t)nn), oe(3 * (size_t)nn), rT(3 * (size_t)nn), rl(2 * (size_t)nn),
      rr(nn), tv(9 * (size_t)nt);
  fclgpu_bvh_get(bvh, fc.data(), axis.data(), oT.data(), oe.data(), rT.data(), rl.data(), rr.data(), tv.data());
  std::vector<double> raxis(9 * (size_t)nn);
  const bool separate = fclgpu_bvh_get_rss_axis(bvh, raxis.data()) == 1;  // host model refitted bottom-up
  int rc = fclgpu_model_create_obbrss2(device, nn, fc.data(), axis.data(), oT.data(), oe.data(), separate ? raxis.data() : nullptr,
                                       rT.data(), rl.data(), rr.data(), nt, tv.data(), out);
  if (rc) return rc;
  std::vector<int32_t> nf(nn), nc(nn), po(nt), ti(3 * (size_t)nt);
  fclgpu_bvh_get_partition(bvh, nf.data(), nc.data(), po.data(), ti.data());
  rc = fclgpu_model_set_partition(*out, fclgpu_bvh_num_vertices(bvh), ti.data(), nf.data(), nc.data(), po.data());
  if (rc) {
    fclgpu_model_destroy(*out);
    *out = nullptr;
  }
  return rc;
}

extern "C" int fclgpu_model_refit_topdown(fclgpu_model* m, const double* vertices, int32_t num_vertices,
                                          int32_t vertices_on_device, void* stream) {
  if (!m || !vertices) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "NULL model/vertices");
  if (!m->prim_order) return fail(FCLGPU_ERR_UNSUPPORTED_FUNCTION, "model has no refit topology (fclgpu_model_set_partition)");
  if (num_vertices != m->num_vertices)  // BVH_model-inl.h:602-606
    return fail(FCLGPU_ERR_INCORRECT_DATA, "the replaced model must have the same number of vertices (%d != %d)", num_vertices, m->num_vertices);
  CUDA_TRY(cudaSetDevice(m->device));
  cudaStream_t st = (cudaStream_t)stream;
  const double* dv = vertices;
  if (!vertices_on_device) {
    CUDA_TRY(cudaMemcpyAsync(m->vert_stage, vertices, 3 * (size_t)num_vertices * sizeof(double), cudaMemcpyHostToDevice, st));
    dv = m->vert_stage;
  }
  RefitParams P{m->obb, m->rss, m->tri, m->rss32, m->obb32, m->topo, m->tri_index, m->node_first, m->node_count,
                m->prim_order, m->by_size, m->d.n_nodes, m->d.n_tris};
  gather_tris_kernel<<<(P.n_tris + 255) / 256, 256, 0, st>>>(P, dv);
  if (opt("refit_warp") && m->n_big > 0) {
    const int n_huge = opt("refit_warp") >= 2 ? m->n_huge : 0;
    if (n_huge > 0) refit_huge_nodes_kernel<<<n_huge, kFitBlock, 0, st>>>(P);
    if (m->n_big > n_huge)
      refit_big_nodes_kernel<<<((m->n_big - n_huge) * 32 + 127) / 128, 128, 0, st>>>(P, n_huge, m->n_big);
    if (P.n_nodes > m->n_big) refit_small_nodes_kernel<<<(P.n_nodes - m->n_big + 63) / 64, 64, 0, st>>>(P, m->n_big);
    g_launches += 2 + (n_huge > 0) + (m->n_big > n_huge);
  } else {
    refit_nodes_kernel<<<(P.n_nodes + 63) / 64, 64, 0, st>>>(P);
    g_launches += 2;
  }
  CUDA_TRY(cudaGetLastError());
  return FCLGPU_OK;
}

// On-device bottom-up refit: BVHModel::endReplaceModel(refit = true, bottomup = true), the reference's default
// (BVH_model-inl.h:952-1037).  One launch per tree height (refit.cuh), the same merge routines as the host model.
extern "C" int fclgpu_model_refit_bottomup(fclgpu_model* m, const double* vertices, int32_t num_vertices,
                                           int32_t vertices_on_device, void* stream) {
  if (!m || !vertices) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "NULL model/vertices");
  if (!m->tri_index) return fail(FCLGPU_ERR_UNSUPPORTED_FUNCTION, "model has no refit topology (fclgpu_model_set_partition)");
  if (num_vertices != m->num_vertices)  // BVH_model-inl.h:602-606
    return fail(FCLGPU_ERR_INCORRECT_DATA, "the replaced model must have the same number of vertices (%d != %d)", num_vertices, m->num_vertices);
  CUDA_TRY(cudaSetDevice(m->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int nn = m->d.n_nodes;
  if (!m->by_height) {
    // height of a node = 1 + max height of its children; children have larger ids than their parent
    std::vector<int32_t> fc(nn), height(nn), ids(nn);
    CUDA_TRY(cudaMemcpy(fc.data(), m->fc, sizeof(int32_t) * nn, cudaMemcpyDeviceToHost));
    int hmax = 0;
    for (int i = nn - 1; i >= 0; --i) {
      height[i] = fc[i] < 0 ? 0 : 1 + std::max(height[fc[i]], height[fc[i] + 1]);
      hmax = std::max(hmax, height[i]);
    }
    m->height_start.assign(hmax + 2, 0);
    for (int i = 0; i < nn; ++i) m->height_start[height[i] + 1]++;
    for (int h = 0; h <= hmax; ++h) m->height_start[h + 1] += m->height_start[h];
    std::vector<int32_t> cur(m->height_start.begin(), m->height_start.end() - 1);
    for (int i = 0; i < nn; ++i) ids[cur[height[i]]++] = i;
    CUDA_TRY(cudaMalloc((void**)&m->by_height, sizeof(int32_t) * nn));
    CUDA_TRY(cudaMemcpy(m->by_height, ids.data(), sizeof(int32_t) * nn, cudaMemcpyHostToDevice));
  }
  const double* dv = vertices;
  if (!vertices_on_device) {
    CUDA_TRY(cudaMemcpyAsync(m->vert_stage, vertices, 3 * (size_t)num_vertices * sizeof(double), cudaMemcpyHostToDevice, st));
    dv = m->vert_stage;
  }
  RefitParams P{m->obb, m->rss, m->tri, m->rss32, m->obb32, m->topo, m->tri_index, m->node_first, m->node_count,
                m->prim_order, m->by_size, m->d.n_nodes, m->d.n_tris};
  gather_tris_kernel<<<(P.n_tris + 255) / 256, 256, 0, st>>>(P, dv);
  const int levels = (int)m->height_start.size() - 1;
  for (int h = 0; h < levels; ++h) {
    const int first = m->height_start[h], count = m->height_start[h + 1] - first;
    refit_bottomup_level_kernel<<<(count + 63) / 64, 64, 0, st>>>(P, m->fc, m->by_height + first, count);
  }
  g_launches += 1 + levels;
  CUDA_TRY(cudaGetLastError());
  return FCLGPU_OK;
}

// rss.axis of every node as it is in HBM (equal to the OBB's axes unless the model was refitted bottom-up)
extern "C" int fclgpu_model_download_rss_axis(const fclgpu_model* m, double* rss_axis9) {
  if (!m || !rss_axis9) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "NULL model/array");
  CUDA_TRY(cudaSetDevice(m->device));
  CUDA_TRY(cudaDeviceSynchronize());
  const int nn = m->d.n_nodes;
  std::vector<double> rss((size_t)nn * kNodeDoubles);
  CUDA_TRY(cudaMemcpy(rss.data(), m->rss, rss.size() * sizeof(double), cudaMemcpyDeviceToHost));
  for (int i = 0; i < nn; ++i)
    for (int k = 0; k < 9; ++k) rss_axis9[9 * (size_t)i + k] = rss[(size_t)i * kNodeDoubles + k];
  return FCLGPU_OK;
}

// On-device build: BVHModel::endModel() (BVH_model-inl.h:450-517) for the mean / BV-centre split rules,
// level by level (refit.cuh).  Produces the same tree, node numbering, primitive order and volumes as
// fclgpu_bvh_build_obbrss + fclgpu_model_from_bvh.
extern "C" int fclgpu_model_build_obbrss(int device, const double* vertices, int32_t num_vertices,
                                         const int32_t* triangles, int32_t num_tris, int32_t split_method,
                                         fclgpu_model** out) {
  if (!out) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "out is NULL");
  *out = nullptr;
  if (num_tris <= 0 || num_vertices <= 0) return fail(FCLGPU_ERR_BUILD_EMPTY_MODEL, "empty model");
  if (!vertices || !triangles) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "NULL vertices/triangles");
  if (split_method != FCLGPU_SPLIT_METHOD_MEAN && split_method != FCLGPU_SPLIT_METHOD_BV_CENTER && split_method != FCLGPU_SPLIT_METHOD_MEDIAN)
    return fail(FCLGPU_ERR_INVALID_ARGUMENT, "unknown split method %d", split_method);
  for (int64_t i = 0; i < 3 * (int64_t)num_tris; ++i)
    if (triangles[i] < 0 || triangles[i] >= num_vertices) return fail(FCLGPU_ERR_INCORRECT_DATA, "triangle index out of range");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return fail(FCLGPU_ERR_NO_DEVICE, "no CUDA device: %s", cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "device %d out of range", device);
  CUDA_TRY(cudaSetDevice(device));
  const int nt = num_tris, nn = 2 * nt - 1;

  fclgpu_model* m = new fclgpu_model{};
  m->device = device;
  m->num_vertices = num_vertices;
  uint8_t* flag = nullptr;
  uint32_t* queue = nullptr;
  BuildNode* lists = nullptr;
  int32_t* next_count = nullptr;
  auto cleanup = [&](int rc) {
    cudaFree(flag);
    cudaFree(queue);
    cudaFree(lists);
    cudaFree(next_count);
    if (rc) fclgpu_model_destroy(m);
    return rc;
  };
#define BUILD_TRY(expr)                                                                    \
  do {                                                                                     \
    cudaError_t e_ = (expr);                                                               \
    if (e_ != cudaSuccess) return cleanup(fail(cuda_status(e_), "%s: %s", #expr, cudaGetErrorString(e_))); \
  } while (0)
  BUILD_TRY(cudaMalloc((void**)&m->obb, (size_t)nn * kNodeDoubles * sizeof(double)));
  BUILD_TRY(cudaMalloc((void**)&m->rss, (size_t)nn * kNodeDoubles * sizeof(double)));
  BUILD_TRY(cudaMalloc((void**)&m->tri, (size_t)nt * kTriDoubles * sizeof(double)));
  BUILD_TRY(cudaMemset(m->tri, 0, (size_t)nt * kTriDoubles * sizeof(double)));
  BUILD_TRY(cudaMalloc((void**)&m->rss32, sizeof(RssRec32) * nn));
  BUILD_TRY(cudaMalloc((void**)&m->obb32, sizeof(ObbRec32) * nn));
  BUILD_TRY(cudaMalloc((void**)&m->topo, sizeof(double2) * nn));
  BUILD_TRY(cudaMemset(m->topo, 0, sizeof(double2) * nn));
  BUILD_TRY(cudaMalloc((void**)&m->fc, sizeof(int32_t) * nn));
  BUILD_TRY(cudaMalloc((void**)&m->tri_index, sizeof(int32_t) * 3 * (size_t)nt));
  BUILD_TRY(cudaMalloc((void**)&m->node_first, sizeof(int32_t) * nn));
  BUILD_TRY(cudaMalloc((void**)&m->node_count, sizeof(int32_t) * nn));
  BUILD_TRY(cudaMalloc((void**)&m->by_size, sizeof(int32_t) * nn));
  BUILD_TRY(cudaMalloc((void**)&m->prim_order, sizeof(uint32_t) * nt));
  BUILD_TRY(cudaMalloc((void**)&m->vert_stage, 3 * (size_t)num_vertices * sizeof(double)));
  BUILD_TRY(cudaMalloc((void**)&flag, nt));
  BUILD_TRY(cudaMalloc((void**)&queue, sizeof(uint32_t) * 2 * (size_t)nt));
  BUILD_TRY(cudaMalloc((void**)&lists, sizeof(BuildNode) * 2 * (size_t)nt));
  BUILD_TRY(cudaMalloc((void**)&next_count, sizeof(int32_t)));
  BUILD_TRY(cudaMemcpy(m->vert_stage, vertices, 3 * (size_t)num_vertices * sizeof(double), cudaMemcpyHostToDevice));
  BUILD_TRY(cudaMemcpy(m->tri_index, triangles, sizeof(int32_t) * 3 * (size_t)nt, cudaMemcpyHostToDevice));

  BuildParams B{};
  B.R = RefitParams{m->obb, m->rss, m->tri, m->rss32, m->obb32, m->topo, m->tri_index, m->node_first, m->node_count,
                    m->prim_order, m->by_size, nn, nt};
  B.first_child = m->fc;
  B.node_first = m->node_first;
  B.node_count = m->node_count;
  B.prim_order = m->prim_order;
  B.flag = flag;
  B.queue = queue;
  B.next_count = next_count;
  B.split = split_method;
  gather_tris_kernel<<<(nt + 255) / 256, 256>>>(B.R, m->vert_stage);
  iota_kernel<<<(nt + 255) / 256, 256>>>(m->prim_order, nt);
  g_launches += 2;
  const BuildNode root{0, 0, nt, 1};
  BUILD_TRY(cudaMemcpy(lists, &root, sizeof root, cudaMemcpyHostToDevice));
  int n_level = 1, cur = 0, depth = 0;
  while (true) {
    BUILD_TRY(cudaMemset(next_count, 0, sizeof(int32_t)));
    B.level = lists + (size_t)cur * nt;
    B.next = lists + (size_t)(cur ^ 1) * nt;
    B.n_level = n_level;
    if ((long long)n_level * 2048 < nt && opt("refit_warp") >= 2)  // a handful of huge nodes: a block each
      build_level_block_kernel<<<n_level, kFitBlock>>>(B);
    else if ((long long)n_level * 12 >= nt)  // average node of the level has <= 12 triangles: lanes take whole nodes
      build_level_kernel<true><<<(n_level + 127) / 128, 128>>>(B);
    else
      build_level_kernel<false><<<(int)(((size_t)n_level * 32 + 127) / 128), 128>>>(B);
    g_launches++;
    int32_t n_next = 0;
    BUILD_TRY(cudaMemcpy(&n_next, next_count, sizeof n_next, cudaMemcpyDeviceToHost));
    if (n_next == 0) break;
    if (n_next > nt) return cleanup(fail(FCLGPU_ERR_UNKNOWN, "level overflow in the on-device build"));
    n_level = n_next;
    cur ^= 1;
    ++depth;
  }
  BUILD_TRY(cudaGetLastError());
  // node sizes -> refit schedule
  std::vector<int32_t> cnt(nn), by_size;
  BUILD_TRY(cudaMemcpy(cnt.data(), m->node_count, sizeof(int32_t) * nn, cudaMemcpyDeviceToHost));
  refit_schedule(cnt.data(), nn, by_size, m->n_huge, m->n_big);
  BUILD_TRY(cudaMemcpy(m->by_size, by_size.data(), sizeof(int32_t) * nn, cudaMemcpyHostToDevice));
  m->depth = depth;
  m->d = DeviceModel{m->obb, m->rss, m->fc, m->tri, m->rss32, m->obb32, m->topo, nn, nt};
  local_aabb_of(nt, [&](int t, int c) { return vertices + 3 * (size_t)triangles[3 * (size_t)t + c]; }, m->aabb);
#undef BUILD_TRY
  *out = m;
  return cleanup(FCLGPU_OK);
}

// topology of a device model (first_child per node, and the build partition when present)
extern "C" int fclgpu_model_get_topology(const fclgpu_model* m, int32_t* first_child, int32_t* first_primitive,
                                         int32_t* num_primitives, int32_t* primitive_indices) {
  if (!m) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "NULL model");
  CUDA_TRY(cudaSetDevice(m->device));
  CUDA_TRY(cudaDeviceSynchronize());
  const int nn = m->d.n_nodes, nt = m->d.n_tris;
  if (first_child) CUDA_TRY(cudaMemcpy(first_child, m->fc, sizeof(int32_t) * nn, cudaMemcpyDeviceToHost));
  if ((first_primitive || num_primitives || primitive_indices) && !m->prim_order)
    return fail(FCLGPU_ERR_UNSUPPORTED_FUNCTION, "model has no partition");
  if (first_primitive) CUDA_TRY(cudaMemcpy(first_primitive, m->node_first, sizeof(int32_t) * nn, cudaMemcpyDeviceToHost));
  if (num_primitives) CUDA_TRY(cudaMemcpy(num_primitives, m->node_count, sizeof(int32_t) * nn, cudaMemcpyDeviceToHost));
  if (primitive_indices) CUDA_TRY(cudaMemcpy(primitive_indices, m->prim_order, sizeof(int32_t) * nt, cudaMemcpyDeviceToHost));
  return FCLGPU_OK;
}

extern "C" int fclgpu_model_download(const fclgpu_model* m, double* axis9, double* obb_To3, double* obb_extent3,
                                     double* rss_To3, double* rss_l2, double* rss_r, double* tri_verts9) {
  if (!m) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "NULL model");
  CUDA_TRY(cudaSetDevice(m->device));
  CUDA_TRY(cudaDeviceSynchronize());
  const int nn = m->d.n_nodes, nt = m->d.n_tris;
  std::vector<double> obb((size_t)nn * kNodeDoubles), rss((size_t)nn * kNodeDoubles), tri((size_t)nt * kTriDoubles);
  CUDA_TRY(cudaMemcpy(obb.data(), m->obb, obb.size() * sizeof(double), cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(rss.data(), m->rss, rss.size() * sizeof(double), cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(tri.data(), m->tri, tri.size() * sizeof(double), cudaMemcpyDeviceToHost));
  for (int i = 0; i < nn; ++i) {
    const double* o = &obb[(size_t)i * kNodeDoubles];
    const double* r = &rss[(size_t)i * kNodeDoubles];
    for (int k = 0; k < 9; ++k)
      if (axis9) axis9[9 * (size_t)i + k] = o[k];
    for (int k = 0; k < 3; ++k) {
      if (obb_To3) obb_To3[3 * (size_t)i + k] = o[9 + k];
      if (obb_extent3) obb_extent3[3 * (size_t)i + k] = o[12 + k];
      if (rss_To3) rss_To3[3 * (size_t)i + k] = r[9 + k];
    }
    if (rss_l2) { rss_l2[2 * (size_t)i] = r[12]; rss_l2[2 * (size_t)i + 1] = r[13]; }
    if (rss_r) rss_r[i] = r[14];
  }
  if (tri_verts9)
    for (int t = 0; t < nt; ++t)
      for (int k = 0; k < 9; ++k) tri_verts9[9 * (size_t)t + k] = tri[(size_t)t * kTriDoubles + k];
  return FCLGPU_OK;
}

extern "C" int fclgpu_model_destroy(fclgpu_model* m) {
  if (!m) return FCLGPU_OK;
  cudaSetDevice(m->device);
  cudaFree(m->obb);
  cudaFree(m->rss);
  cudaFree(m->tri);
  cudaFree(m->fc);
  cudaFree(m->rss32);
  cudaFree(m->obb32);
  cudaFree(m->topo);
  cudaFree(m->tri_index);
  cudaFree(m->node_first);
  cudaFree(m->node_count);
  cudaFree(m->by_size);
  cudaFree(m->by_height);
  cudaFree(m->prim_order);
  cudaFree(m->vert_stage);
  delete m;
  return FCLGPU_OK;
}
extern "C" int32_t fclgpu_model_num_nodes(const fclgpu_model* m) { return m ? m->d.n_nodes : 0; }
extern "C" int32_t fclgpu_model_num_tris(const fclgpu_model* m) { return m ? m->d.n_tris : 0; }
extern "C" int fclgpu_model_device(const fclgpu_model* m) { return m ? m->device : -1; }

// ------------------------------------------------------------------------------------------
// contact compaction: exclusive scan of per-query counts -> offsets, then gather from the
// fixed-stride scratch into the dense pool (deterministic: query order, DFS order inside)
// ------------------------------------------------------------------------------------------
namespace {

constexpr int kScanBlock = 1024;

// pass 1: per-block inclusive scan (counts clipped to `stride` = what was actually stored)
__global__ void scan_block_kernel(const int32_t* __restrict__ counts, long long n, long long stride,
                                  long long* __restrict__ local, long long* __restrict__ block_sums) {
  __shared__ long long warp_sums[32];
  const long long i = (long long)blockIdx.x * kScanBlock + threadIdx.x;
  long long v = 0;
  if (i < n) {
    v = counts[i];
    if (v > stride) v = stride;
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  long long x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    long long y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_sums[wid] = x;
  __syncthreads();
  if (wid == 0) {
    long long w = warp_sums[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      long long y = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += y;
    }
    warp_sums[lane] = w;
  }
  __syncthreads();
  const long long prefix = (wid > 0 ? warp_sums[wid - 1] : 0) + x;  // inclusive
  if (i < n) local[i] = prefix - v;                                // exclusive within block
  if (threadIdx.x == kScanBlock - 1) block_sums[blockIdx.x] = prefix;
}

// pass 2: one block scans the block sums in place (exclusive), adds the running base
__global__ void scan_sums_kernel(long long* __restrict__ block_sums, int nblocks, long long* __restrict__ base,
                                 long long* __restrict__ total_out) {
  __shared__ long long carry;
  __shared__ long long warp_sums[32];
  if (threadIdx.x == 0) carry = *base;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int s = 0; s < nblocks; s += kScanBlock) {
    const int i = s + threadIdx.x;
    const long long v = (i < nblocks) ? block_sums[i] : 0;
    long long x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      long long y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
      long long w = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        long long y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      warp_sums[lane] = w;
    }
    __syncthreads();
    const long long incl = (wid > 0 ? warp_sums[wid - 1] : 0) + x;
    if (i < nblocks) block_sums[i] = carry + incl - v;
    __syncthreads();
    if (threadIdx.x == kScanBlock - 1) carry += incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    *base = carry;
    if (total_out) *total_out = carry;
  }
}

// pass 3: offsets[i] = block offset + local; gather contacts (one warp per query, 16-byte copies)
__global__ void compact_kernel(const int32_t* __restrict__ counts, long long n, long long stride,
                               const long long* __restrict__ local, const long long* __restrict__ block_sums,
                               const fclgpu_contact* __restrict__ scratch, fclgpu_contact* __restrict__ pool,
                               long long capacity, long long* __restrict__ offsets, int* status) {
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n) return;
  const long long off = block_sums[warp / kScanBlock] + local[warp];
  long long c = counts[warp];
  if (c > stride) c = stride;
  if (lane == 0 && offsets) offsets[warp] = off;
  if (pool == nullptr) return;
  if (off + c > capacity) {
    if (lane == 0) atomicMin(status, (int)FCLGPU_ERR_CONTACT_OVERFLOW);
    c = capacity > off ? capacity - off : 0;
  }
  const int4* src = reinterpret_cast<const int4*>(scratch + warp * stride);
  int4* dst = reinterpret_cast<int4*>(pool + off);
  for (long long k = lane; k < c * 4; k += 32) dst[k] = src[k];
}

__global__ void write_total_kernel(const long long* base, long long* offsets_n) { *offsets_n = *base; }

template <typename K, typename Pm, typename... Extra>
int launch_persistent(K kernel, const Pm& params, Workspace* w, int block, cudaStream_t st, size_t smem = 0,
                      Extra... extra) {
  if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = (int)opt("blocks_per_sm");
  if (per_sm <= 0) {
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem));
    if (per_sm < 1) per_sm = 1;
  }
  const int grid = w->sm_count * per_sm;
  kernel<<<grid, block, smem, st>>>(params, extra...);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

unsigned long long* next_counter(Workspace* w, cudaStream_t st) {
  unsigned long long* c = w->counters + w->counter_next;
  w->counter_next = (w->counter_next + 1) % w->counter_slots;
  cudaMemsetAsync(c, 0, sizeof(unsigned long long), st);
  return c;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// collide
// ------------------------------------------------------------------------------------------
namespace {
// Hidden knobs of the host wrappers around collide_enqueue.
size_t contact_record_bytes(const fclgpu_collision_request* r) {
  return r->contact_format == FCLGPU_CONTACT_IDS ? sizeof(fclgpu_contact_ids)
         : r->contact_format == FCLGPU_CONTACT_F32 ? sizeof(fclgpu_contact_f32) : sizeof(fclgpu_contact);
}
long long stage_capacity(const fclgpu_collision_request* r) {
  return r->stage_capacity > 0 ? (long long)r->stage_capacity : std::max<long long>(1, opt("contact_stride"));
}
struct CollideExtra {
  const unsigned* ready = nullptr;  // streamed input: per-chunk "poses have landed" flags (see wait_ready)
  int ready_shift = 0;
  long long ready_q0 = 0;           // index of this call's first query in the flagged batch
  bool continue_scan = false;       // keep the running contact offset of the previous call (sub-batches of one result)
  double sphere_radius = -1.0;      // >= 0: model 2 is a sphere of this radius (mesh <-> sphere collide)
  int plane_kind = -1;              // 0 / 1: model 2 is a halfspace / plane {n . x <= d / n . x = d} (normalised)
  double plane_n[3] = {0, 0, 0}, plane_d = 0;
};
int collide_enqueue(const fclgpu_model* m1, const fclgpu_model* m2, int64_t n, const double* tf1, const double* tf2,
                    const fclgpu_collision_request* request, int32_t* num_contacts, fclgpu_contact* contacts,
                    int64_t contact_capacity, int64_t* contact_offsets, uint32_t* n_bv, uint32_t* n_leaf, void* stream,
                    const CollideExtra& X);
}  // namespace

extern "C" int fclgpu_collide_batch(const fclgpu_model* m1, const fclgpu_model* m2, int64_t n, const double* tf1,
                                    const double* tf2, const fclgpu_collision_request* request,
                                    int32_t* num_contacts, fclgpu_contact* contacts, int64_t contact_capacity,
                                    int64_t* contact_offsets, uint32_t* n_bv, uint32_t* n_leaf, void* stream) {
  return collide_enqueue(m1, m2, n, tf1, tf2, request, num_contacts, contacts, contact_capacity, contact_offsets, n_bv,
                         n_leaf, stream, CollideExtra{});
}

namespace {
// Locality order of queries [0, n) of a resident batch into the workspace's order buffer (three small launches).
int enqueue_query_order(Workspace* w, const fclgpu_model* m1, const fclgpu_model* m2, long long n, const double* tf1, const double* tf2,
                        cudaStream_t st, const int32_t** order) {
  const size_t bytes = padded_up(2 * (size_t)n) + 2 * sizeof(uint32_t) * kOrderBuckets + padded_up(4 * (size_t)n);
  int rc = ensure(&w->order_buf, &w->order_bytes, bytes);
  if (rc) return rc;
  order_after(st, w->order_ev, w->order_used, w->order_stream);
  char* base = (char*)w->order_buf;
  OrderParams O;
  O.tf1 = tf1;
  O.tf2 = tf2;
  O.n = n;
  for (int k = 0; k < 3; ++k) {
    O.c1[k] = m1->aabb.c[k];
    O.c2[k] = m2->aabb.c[k];
  }
  O.reach = m1->aabb.r + m2->aabb.r;
  O.hist = (uint32_t*)base;
  O.cursor = O.hist + kOrderBuckets;
  O.key = (uint16_t*)(base + 2 * sizeof(uint32_t) * kOrderBuckets);
  O.order = (int32_t*)(base + 2 * sizeof(uint32_t) * kOrderBuckets + padded_up(2 * (size_t)n));
  CUDA_TRY(cudaMemsetAsync(O.hist, 0, sizeof(uint32_t) * kOrderBuckets, st));
  const unsigned blocks = (unsigned)((n + 255) / 256);
  order_key_kernel<<<blocks, 256, 0, st>>>(O);
  order_scan_kernel<<<1, 1024, 0, st>>>(O);
  order_scatter_kernel<<<blocks, 256, 0, st>>>(O);
  g_launches += 3;
  CUDA_TRY(cudaGetLastError());
  *order = O.order;
  return 0;
}
}  // namespace

namespace {
int collide_enqueue(const fclgpu_model* m1, const fclgpu_model* m2, int64_t n, const double* tf1, const double* tf2,
                    const fclgpu_collision_request* request, int32_t* num_contacts, fclgpu_contact* contacts,
                    int64_t contact_capacity, int64_t* contact_offsets, uint32_t* n_bv, uint32_t* n_leaf, void* stream,
                    const CollideExtra& X) {
  if (!m1 || !m2 || !request || n < 0) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "NULL model/request or n<0");
  if (m1->device != m2->device) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "models live on different devices");
  if (request->enable_cost) return fail(FCLGPU_ERR_UNSUPPORTED_FUNCTION, "cost sources are not supported on this path");
  if (!num_contacts) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "num_contacts is NULL");
  const bool shape2 = X.sphere_radius >= 0 || X.plane_kind >= 0;  // model 2 is a primitive shape: one-sided traversal
  if (m1->depth + (shape2 ? 0 : m2->depth) + 2 > kStackCap)
    return fail(FCLGPU_ERR_STACK_OVERFLOW, "tree depths %d+%d exceed the traversal stack", m1->depth, m2->depth);
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cudaSetDevice(m1->device));
  Workspace* w;
  int rc = get_ws(m1->device, &w);
  if (rc) return rc;
  std::lock_guard<std::mutex> lock(w->mu);
  if (n == 0) {
    if (contact_offsets) CUDA_TRY(cudaMemsetAsync(contact_offsets, 0, sizeof(int64_t), st));
    return FCLGPU_OK;
  }
  if (request->num_max_contacts <= 0) {  // collision-inl.h:111-115: warn and return 0
    CUDA_TRY(cudaMemsetAsync(num_contacts, 0, sizeof(int32_t) * n, st));
    if (contact_offsets) CUDA_TRY(cudaMemsetAsync(contact_offsets, 0, sizeof(int64_t) * (n + 1), st));
    if (n_bv) CUDA_TRY(cudaMemsetAsync(n_bv, 0, sizeof(uint32_t) * n, st));
    if (n_leaf) CUDA_TRY(cudaMemsetAsync(n_leaf, 0, sizeof(uint32_t) * n, st));
    return FCLGPU_OK;
  }
  const bool want_contacts = contacts != nullptr || contact_offsets != nullptr;
  const bool stats = (n_bv || n_leaf);
  const StreamState ss = stream_state(w, st);
  const long long trav0 = opt("traversal");
  if (request->contact_format < FCLGPU_CONTACT_FULL || request->contact_format > FCLGPU_CONTACT_F32)
    return fail(FCLGPU_ERR_INVALID_ARGUMENT, "unknown contact_format %d", request->contact_format);
  if (request->contact_format != FCLGPU_CONTACT_FULL && contacts != nullptr && !(trav0 >= 3 && !opt("contact_order") && !shape2))
    return fail(FCLGPU_ERR_UNSUPPORTED_FUNCTION, "compact contact records are written by the default mesh-mesh contact path only");

  if (want_contacts && trav0 >= 3 && !opt("contact_order") && !shape2) {
    // Contact list, default path: ONE launch of the warp-per-query ordered-front kernel (collide_ordered.cuh).
    // Contacts are staged per RESIDENT WARP (L2-resident) and appended to the caller's pool when a query retires.
    const long long stride = std::min<long long>(request->num_max_contacts, stage_capacity(request));
    size_t smem = sizeof(OrderedFront) * kOrdWarps;
    auto kern = stats ? collide_ordered_kernel<true> : collide_ordered_kernel<false>;
    // short lists (the usual num_max_contacts <= ~100) are staged in shared memory when that does not cost a resident block
    int smem_stage = 0;
    int per_sm = (int)opt("blocks_per_sm");
    if (per_sm <= 0) {
      CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kOrdWarps * 32, smem));
      if (per_sm < 1) per_sm = 1;
      const size_t with_stage = smem + (size_t)kOrdWarps * (size_t)stride * sizeof(fclgpu_contact);
      if (opt("contact_smem_stage") && contacts != nullptr && with_stage <= 100 * 1024) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)with_stage));
        int per_sm2 = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm2, kern, kOrdWarps * 32, with_stage));
        if (per_sm2 >= per_sm) {
          smem = with_stage;
          smem_stage = (int)stride;
        }
      }
    }
    const long long warps = (long long)w->sm_count * per_sm * kOrdWarps;
    rc = ensure(&w->scratch, &w->scratch_bytes, (size_t)(warps * stride) * sizeof(fclgpu_contact));
    if (rc) return rc;
    if (!X.continue_scan) CUDA_TRY(cudaMemsetAsync(ss.cursor, 0, sizeof(long long), st));
    order_after(st, w->scratch_ev, w->scratch_used, w->scratch_stream);
    OrderedParams Q;
    CollideParams& P = Q.C;
    P.m1 = m1->d;
    P.m2 = m2->d;
    P.tf1 = tf1;
    P.tf2 = tf2;
    P.n = n;
    P.max_contacts = request->num_max_contacts;
    P.enable_contact = request->enable_contact ? 1 : 0;
    P.num_contacts = num_contacts;
    P.scratch = (fclgpu_contact*)w->scratch;
    P.stride = stride;
    P.n_bv = n_bv;
    P.n_leaf = n_leaf;
    P.work_counter = next_counter(w, st);
    P.status = ss.status;
    P.ready = X.ready;
    P.ready_shift = X.ready_shift;
    P.ready_q0 = X.ready_q0;
    Q.pool = contacts;
    Q.pool_capacity = contacts ? contact_capacity : 0;
    Q.cursor = (unsigned long long*)ss.cursor;
    Q.starts = (long long*)contact_offsets;
    Q.depth_sum = m1->depth + m2->depth;
    Q.smem_stage = smem_stage;
    Q.format = request->contact_format;
    // (the per-warp staging block must start on a 128-byte line: stride * 64 bytes per warp)
    Q.discard_stage = (opt("contact_discard") && stride % 2 == 0) ? 1 : 0;
    kern<<<(unsigned)(w->sm_count * per_sm), kOrdWarps * 32, smem, st>>>(Q);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    if (contact_offsets) {
      write_total_kernel<<<1, 1, 0, st>>>(ss.cursor, (long long*)contact_offsets + n);
      g_launches++;
    }
    mark_use(st, w->scratch_ev, w->scratch_used, w->scratch_stream);
    CUDA_TRY(cudaGetLastError());
    return FCLGPU_OK;
  }

  long long stride = 0, chunk = n;
  if (want_contacts) {
    stride = std::min<long long>(request->num_max_contacts, stage_capacity(request));
    const long long budget = std::max<long long>(opt("scratch_bytes"), stride * (long long)sizeof(fclgpu_contact));
    chunk = std::max<long long>(1, std::min<long long>(n, budget / (stride * (long long)sizeof(fclgpu_contact))));
    rc = ensure(&w->scratch, &w->scratch_bytes, (size_t)(chunk * stride) * sizeof(fclgpu_contact));
    if (rc) return rc;
    const long long nblk = (chunk + kScanBlock - 1) / kScanBlock;
    rc = ensure(&w->scan_tmp, &w->scan_bytes, (size_t)(chunk + nblk) * sizeof(long long));
    if (rc) return rc;
    if (!X.continue_scan) CUDA_TRY(cudaMemsetAsync(ss.cursor, 0, sizeof(long long), st));
    order_after(st, w->scratch_ev, w->scratch_used, w->scratch_stream);
  }

  for (long long s = 0; s < n; s += chunk) {
    const long long cn = std::min(chunk, n - s);
    CollideParams P;
    P.m1 = m1->d;
    P.m2 = m2->d;
    P.tf1 = tf1 ? tf1 + 12 * s : nullptr;
    P.tf2 = tf2 ? tf2 + 12 * s : nullptr;
    P.n = cn;
    P.max_contacts = request->num_max_contacts;
    P.enable_contact = request->enable_contact ? 1 : 0;
    P.num_contacts = num_contacts + s;
    P.scratch = want_contacts ? (fclgpu_contact*)w->scratch : nullptr;
    P.stride = stride;
    P.n_bv = n_bv ? n_bv + s : nullptr;
    P.n_leaf = n_leaf ? n_leaf + s : nullptr;
    P.work_counter = next_counter(w, st);
    P.status = ss.status;
    P.ready = X.ready;
    P.ready_shift = X.ready_shift;
    P.ready_q0 = X.ready_q0 + s;
    P.order = nullptr;
    P.front_cap = P.front_reserve = P.front_leaf_trigger = 0;
    const long long trav = trav0;
    const int trig = (int)opt("leaf_trigger");
    const long long front = opt("collide_front");  // 0 never, 1 (default) for BVHs beyond the caches, 2 always
    if (X.sphere_radius >= 0) {
      rc = stats ? launch_persistent(collide_mesh_sphere_kernel<true>, P, w, 128, st, 0, X.sphere_radius)
                 : launch_persistent(collide_mesh_sphere_kernel<false>, P, w, 128, st, 0, X.sphere_radius);
    } else if (X.plane_kind >= 0) {
      rc = stats ? launch_persistent(collide_mesh_plane_kernel<true>, P, w, 128, st, 0, X.plane_kind, X.plane_n[0], X.plane_n[1],
                                     X.plane_n[2], X.plane_d)
                 : launch_persistent(collide_mesh_plane_kernel<false>, P, w, 128, st, 0, X.plane_kind, X.plane_n[0], X.plane_n[1],
                                     X.plane_n[2], X.plane_d);
    } else if (!want_contacts && !P.enable_contact && trav >= 1 &&
        (front >= 2 || (front == 1 && ((long long)m1->d.n_nodes + m2->d.n_nodes >= (1 << 17) || cn <= opt("front_small_batch"))))) {
      // counts only: the result does not depend on the visiting order -> warp-per-query front kernel.  Chosen for BVHs
      // beyond the caches (memory-level parallelism) and for SMALL batches: a lane-per-query launch cannot end before
      // its longest query (~0.8 ms on env/rob), a warp-per-query launch spreads every query over 32 lanes
      // (BASELINE cfg1 at its real size, 10k poses: 0.83 ms -> 0.11 ms)
      P.front_reserve = m1->depth + m2->depth + 2 + 32;
      const long long cap_opt = opt("front_cap");
      P.front_cap = (int)std::min<long long>(std::max<long long>(cap_opt > 0 ? cap_opt : 384, P.front_reserve + 64), 3072);
      // A query whose count is within 32 of num_max_contacts (every verdict query) may end at its next leaf round.  With the
      // BVH in L2 / HBM a BV round costs its load latency, an early, partly filled leaf round is cheap next to the rounds it
      // saves (cfg4 21.6 -> 19.9 ms, cfg5 5.44 -> 4.84 ms); with cache-resident models (env/rob) it is not (0.087 -> 0.102 ms).
      const bool beyond_caches = (long long)m1->d.n_nodes + m2->d.n_nodes >= (1 << 17);
      const long long trig_opt = opt("front_leaf_trigger");
      P.front_leaf_trigger = (int)std::min<long long>(trig_opt > 0 ? trig_opt : (beyond_caches ? 8 : 32), 32);
      const size_t fsm = front_bytes_per_warp(P.front_cap) * 4;
      rc = stats ? launch_persistent(collide_front_kernel<true>, P, w, 128, st, fsm)
                 : launch_persistent(collide_front_kernel<false>, P, w, 128, st, fsm);
    } else if (trav >= 2 && !P.enable_contact && (trav == 2 || !opt("binary_pooled"))) {
      rc = stats ? launch_persistent(collide_deferred_kernel<true, true, true>, P, w, 128, st, 0, trig)
                 : launch_persistent(collide_deferred_kernel<false, true, true>, P, w, 128, st, 0, trig);
    } else if (trav >= 3) {  // pooled leaf rounds
      const int ptrig = (int)opt("pool_trigger");
      const size_t psm = sizeof(PoolWarp) * 4;
      const long long min_order = opt("order_queries");
      bool ordered = false;
      if (min_order > 0 && cn >= min_order && X.ready == nullptr && cn < (1ll << 31)) {
        rc = enqueue_query_order(w, m1, m2, cn, P.tf1, P.tf2, st, &P.order);
        if (rc) return rc;
        ordered = true;
      }
      if (!P.enable_contact && opt("binary_pooled") >= 2)
        rc = stats ? launch_persistent(collide_pooled_kernel<true, true>, P, w, 128, st, psm, ptrig)
                   : launch_persistent(collide_pooled_kernel<false, true>, P, w, 128, st, psm, ptrig);
      else if (trav >= 4 || (long long)m1->d.n_nodes + m2->d.n_nodes >= (1 << 17))
        // two children per BV round: 7-15 % faster once the node records no longer fit the caches
        // (cfg4 / cfg5 shapes), 15-20 % slower on cache-resident models, hence chosen by model size
        rc = stats ? launch_persistent(collide_pooled_kernel<true, false, true>, P, w, 128, st, psm, ptrig)
                   : launch_persistent(collide_pooled_kernel<false, false, true>, P, w, 128, st, psm, ptrig);
      else
        rc = stats ? launch_persistent(collide_pooled_kernel<true, false>, P, w, 128, st, psm, ptrig)
                   : launch_persistent(collide_pooled_kernel<false, false>, P, w, 128, st, psm, ptrig);
      if (ordered) mark_use(st, w->order_ev, w->order_used, w->order_stream);
    } else if (trav >= 2) {
      rc = stats ? launch_persistent(collide_deferred_kernel<true, true, false>, P, w, 128, st, 0, trig)
                 : launch_persistent(collide_deferred_kernel<false, true, false>, P, w, 128, st, 0, trig);
    } else if (trav == 1) {
      rc = stats ? launch_persistent(collide_deferred_kernel<true, false, false>, P, w, 128, st, 0, trig)
                 : launch_persistent(collide_deferred_kernel<false, false, false>, P, w, 128, st, 0, trig);
    } else {
      rc = stats ? launch_persistent(collide_thread_kernel<true>, P, w, 128, st)
                 : launch_persistent(collide_thread_kernel<false>, P, w, 128, st);
    }
    if (rc) return rc;
    if (want_contacts) {
      long long* local = (long long*)w->scan_tmp;
      long long* bsums = local + cn;
      const int nblk = (int)((cn + kScanBlock - 1) / kScanBlock);
      scan_block_kernel<<<nblk, kScanBlock, 0, st>>>(num_contacts + s, cn, stride, local, bsums);
      scan_sums_kernel<<<1, kScanBlock, 0, st>>>(bsums, nblk, ss.cursor, nullptr);
      const long long threads = cn * 32;
      compact_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(
          num_contacts + s, cn, stride, local, bsums, (const fclgpu_contact*)w->scratch, contacts, contact_capacity,
          contact_offsets ? (long long*)contact_offsets + s : nullptr, ss.status);
      g_launches += 3;
      CUDA_TRY(cudaGetLastError());
    }
  }
  if (want_contacts && contact_offsets) {
    write_total_kernel<<<1, 1, 0, st>>>(ss.cursor, (long long*)contact_offsets + n);
    g_launches++;
  }
  if (want_contacts) mark_use(st, w->scratch_ev, w->scratch_used, w->scratch_stream);
  CUDA_TRY(cudaGetLastError());
  return FCLGPU_OK;
}
}  // namespace

// ------------------------------------------------------------------------------------------
// mesh <-> sphere collide
// ------------------------------------------------------------------------------------------
extern "C" int fclgpu_collide_mesh_sphere_batch(const fclgpu_model* m1, double radius, int64_t n, const double* tf1,
                                                const double* tf2, const fclgpu_collision_request* request,
                                                int32_t* num_contacts, fclgpu_contact* contacts, int64_t contact_capacity,
                                                int64_t* contact_offsets, uint32_t* n_bv, uint32_t* n_leaf, void* stream) {
  if (!(radius >= 0)) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "sphere radius must be >= 0");
  CollideExtra X;
  X.sphere_radius = radius;
  return collide_enqueue(m1, m1, n, tf1, tf2, request, num_contacts, contacts, contact_capacity, contact_offsets, n_bv,
                         n_leaf, stream, X);
}

// ------------------------------------------------------------------------------------------
// mesh <-> halfspace / plane collide
// ------------------------------------------------------------------------------------------
namespace {
// Halfspace(n, d) / Plane(n, d) constructors -> unitNormalTest (geometry/shape/halfspace-inl.h:144-160)
int plane_extra(int kind, const double* normal3, double d, CollideExtra& X) {
  if (!normal3) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "normal is NULL");
  const double l = std::sqrt((normal3[0] * normal3[0] + normal3[1] * normal3[1]) + normal3[2] * normal3[2]);
  X.plane_kind = kind;
  if (l > 0) {
    const double inv_l = 1.0 / l;
    for (int k = 0; k < 3; ++k) X.plane_n[k] = normal3[k] * inv_l;
    X.plane_d = d * inv_l;
  } else {
    X.plane_n[0] = 1; X.plane_n[1] = 0; X.plane_n[2] = 0;
    X.plane_d = 0;
  }
  return FCLGPU_OK;
}
}  // namespace

extern "C" int fclgpu_collide_mesh_plane_batch(const fclgpu_model* m1, int32_t shape, const double* normal3, double d, int64_t n,
                                               const double* tf1, const double* tf2, const fclgpu_collision_request* request,
                                               int32_t* num_contacts, fclgpu_contact* contacts, int64_t contact_capacity,
                                               int64_t* contact_offsets, uint32_t* n_bv, uint32_t* n_leaf, void* stream) {
  if (shape != FCLGPU_SHAPE_HALFSPACE && shape != FCLGPU_SHAPE_PLANE) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "shape must be FCLGPU_SHAPE_HALFSPACE or FCLGPU_SHAPE_PLANE");
  CollideExtra X;
  int rc = plane_extra(shape == FCLGPU_SHAPE_HALFSPACE ? 0 : 1, normal3, d, X);
  if (rc) return rc;
  return collide_enqueue(m1, m1, n, tf1, tf2, request, num_contacts, contacts, contact_capacity, contact_offsets, n_bv, n_leaf,
                         stream, X);
}

// ------------------------------------------------------------------------------------------
// distance
// ------------------------------------------------------------------------------------------
namespace {
// sphere_radius >= 0: model 2 is a sphere of that radius (mesh <-> sphere distance; m2 == m1 is passed and ignored)
int distance_enqueue(const fclgpu_model* m1, const fclgpu_model* m2, int64_t n, const double* tf1, const double* tf2,
                     const fclgpu_distance_request* request, double* min_distance, double* nearest_p1,
                     double* nearest_p2, int32_t* b1, int32_t* b2, uint32_t* n_bv, uint32_t* n_leaf, void* stream,
                     double sphere_radius, double cutoff = 1.7976931348623157e308, double stop_below = -1.0,
                     uint8_t* within = nullptr) {
  if (!m1 || !m2 || !request || n < 0) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "NULL model/request or n<0");
  if (m1->device != m2->device) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "models live on different devices");
  if (m1->depth + (sphere_radius >= 0 ? 0 : m2->depth) + 2 > kStackCap)
    return fail(FCLGPU_ERR_STACK_OVERFLOW, "tree depths %d+%d exceed the traversal stack", m1->depth, m2->depth);
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cudaSetDevice(m1->device));
  Workspace* w;
  int rc = get_ws(m1->device, &w);
  if (rc) return rc;
  std::lock_guard<std::mutex> lock(w->mu);
  if (n == 0) return FCLGPU_OK;
  DistanceParams P;
  P.m1 = m1->d;
  P.m2 = m2->d;
  P.tf1 = tf1;
  P.tf2 = tf2;
  P.n = n;
  P.enable_nearest_points = request->enable_nearest_points ? 1 : 0;
  P.min_distance = min_distance;
  P.p1 = nearest_p1;
  P.p2 = nearest_p2;
  P.b1 = b1;
  P.b2 = b2;
  P.n_bv = n_bv;
  P.n_leaf = n_leaf;
  P.work_counter = next_counter(w, st);
  P.status = stream_state(w, st).status;
  P.spill_pair = nullptr;
  P.spill_bound = nullptr;
  P.spill_cap = 0;
  P.spill_warps = 0;
  P.cutoff = cutoff;
  P.stop_below = stop_below;
  P.within = within;
  Workspace::SpillArea* area = nullptr;
  if ((long long)m1->d.n_nodes + m2->d.n_nodes >= (1 << 17) && opt("dist_spill_entries") >= kSpillBlock) {
    // big models: the sorted front may outgrow its shared-memory stack; give every warp of the largest
    // possible grid an overflow area in HBM (12 bytes per entry)
    const int cap = (int)opt("dist_spill_entries");
    for (auto& a : w->spill)
      if (a.used && a.stream == st) area = &a;  // the area this stream used last
    if (!area)
      for (auto& a : w->spill)
        if (!a.used) {
          area = &a;
          break;
        }
    if (!area) area = &w->spill[w->spill_next++ & 1u];  // more streams than areas: share one, ordered by its event
    if (area->cap != cap) {
      if (area->pair) CUDA_TRY(cudaFree(area->pair));
      if (area->bound) CUDA_TRY(cudaFree(area->bound));
      area->pair = nullptr;
      area->bound = nullptr;
      area->cap = 0;
      const size_t warps = (size_t)w->sm_count * 8 * kDistWarps;  // up to 8 resident blocks per SM
      CUDA_TRY(cudaMalloc((void**)&area->pair, warps * cap * sizeof(uint2)));
      CUDA_TRY(cudaMalloc((void**)&area->bound, warps * cap * sizeof(float)));
      area->cap = cap;
    }
    P.spill_pair = area->pair;
    P.spill_bound = area->bound;
    P.spill_cap = cap;
    P.spill_warps = w->sm_count * 8 * kDistWarps;
    // the area is indexed by the launch-local warp id: a launch on another stream must not overlap its previous user
    order_after(st, area->ev, area->used, area->stream);
  }
  const bool stats = (n_bv || n_leaf);
  if (sphere_radius >= 0) {
    const int trig = (int)opt("sphere_leaf_trigger");
    const int b32 = (int)opt("sphere_bound32");
    if (trig > 0) {
      const long long occ = opt("sphere_blocks");  // register budget of the kernel: 3, 4 or 5 resident blocks per SM
      if (stats) return launch_persistent(distance_mesh_sphere_rounds_kernel<true, 3>, P, w, 128, st, 0, sphere_radius, trig, b32);
      if (occ >= 5) return launch_persistent(distance_mesh_sphere_rounds_kernel<false, 5>, P, w, 128, st, 0, sphere_radius, trig, b32);
      if (occ == 4) return launch_persistent(distance_mesh_sphere_rounds_kernel<false, 4>, P, w, 128, st, 0, sphere_radius, trig, b32);
      return launch_persistent(distance_mesh_sphere_rounds_kernel<false, 3>, P, w, 128, st, 0, sphere_radius, trig, b32);
    }
    return stats ? launch_persistent(distance_mesh_sphere_kernel<true>, P, w, 128, st, 0, sphere_radius)
                 : launch_persistent(distance_mesh_sphere_kernel<false>, P, w, 128, st, 0, sphere_radius);
  }
  const long long trav = opt("traversal");
  const size_t front_smem = sizeof(WarpFront) * kDistWarps;
  const bool tol = within != nullptr || stop_below >= 0;
  if (trav >= 1 && tol) {  // tolerance verdicts: FP32-steered front with the early exit (also for traversal 1)
    if (P.spill_pair)
      rc = stats ? launch_persistent(distance_warp_kernel<true, true, true, true>, P, w, kDistWarps * 32, st, front_smem)
                 : launch_persistent(distance_warp_kernel<false, true, true, true>, P, w, kDistWarps * 32, st, front_smem);
    else
      rc = stats ? launch_persistent(distance_warp_kernel<true, true, false, true>, P, w, kDistWarps * 32, st, front_smem)
                 : launch_persistent(distance_warp_kernel<false, true, false, true>, P, w, kDistWarps * 32, st, front_smem);
  } else if (trav >= 2 && P.spill_pair) {
    rc = stats ? launch_persistent(distance_warp_kernel<true, true, true>, P, w, kDistWarps * 32, st, front_smem)
               : launch_persistent(distance_warp_kernel<false, true, true>, P, w, kDistWarps * 32, st, front_smem);
  } else if (trav >= 2) {
    rc = stats ? launch_persistent(distance_warp_kernel<true, true>, P, w, kDistWarps * 32, st, front_smem)
               : launch_persistent(distance_warp_kernel<false, true>, P, w, kDistWarps * 32, st, front_smem);
  } else if (trav == 1) {
    rc = stats ? launch_persistent(distance_warp_kernel<true, false>, P, w, kDistWarps * 32, st, front_smem)
               : launch_persistent(distance_warp_kernel<false, false>, P, w, kDistWarps * 32, st, front_smem);
  } else {
    rc = stats ? launch_persistent(distance_thread_kernel<true>, P, w, 128, st)
               : launch_persistent(distance_thread_kernel<false>, P, w, 128, st);
  }
  if (area) mark_use(st, area->ev, area->used, area->stream);
  return rc;
}
}  // namespace

extern "C" int fclgpu_distance_batch(const fclgpu_model* m1, const fclgpu_model* m2, int64_t n, const double* tf1,
                                     const double* tf2, const fclgpu_distance_request* request,
                                     double* min_distance, double* nearest_p1, double* nearest_p2, int32_t* b1,
                                     int32_t* b2, uint32_t* n_bv, uint32_t* n_leaf, void* stream) {
  return distance_enqueue(m1, m2, n, tf1, tf2, request, min_distance, nearest_p1, nearest_p2, b1, b2, n_bv, n_leaf, stream,
                          -1.0);
}

// tolerance verification (extension, BASELINE cfg5): fcl::distance started from min_distance = cutoff
extern "C" int fclgpu_distance_cutoff_batch(const fclgpu_model* m1, const fclgpu_model* m2, int64_t n, const double* tf1,
                                            const double* tf2, const fclgpu_distance_request* request, double cutoff,
                                            double* min_distance, double* nearest_p1, double* nearest_p2, int32_t* b1,
                                            int32_t* b2, uint32_t* n_bv, uint32_t* n_leaf, void* stream) {
  if (!(cutoff > 0)) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "cutoff must be > 0");
  return distance_enqueue(m1, m2, n, tf1, tf2, request, min_distance, nearest_p1, nearest_p2, b1, b2, n_bv, n_leaf, stream,
                          -1.0, cutoff);
}

// tolerance verdicts with early exit (extension, BASELINE cfg5)
extern "C" int fclgpu_within_tolerance_batch(const fclgpu_model* m1, const fclgpu_model* m2, int64_t n, const double* tf1,
                                             const double* tf2, double tolerance, uint8_t* within, double* witness_distance,
                                             uint32_t* n_bv, uint32_t* n_leaf, void* stream) {
  if (!(tolerance >= 0)) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "tolerance must be >= 0");
  if (!within) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "within is NULL");
  const fclgpu_distance_request rq{0, 0, 0.0, 0.0};
  return distance_enqueue(m1, m2, n, tf1, tf2, &rq, witness_distance, nullptr, nullptr, nullptr, nullptr, n_bv, n_leaf, stream,
                          -1.0, std::nextafter(tolerance, 1.7976931348623157e308), tolerance, within);
}

// mesh <-> sphere distance
extern "C" int fclgpu_distance_mesh_sphere_batch(const fclgpu_model* m1, double radius, int64_t n, const double* tf1,
                                                 const double* tf2, const fclgpu_distance_request* request,
                                                 double* min_distance, double* nearest_p1, double* nearest_p2,
                                                 int32_t* b1, int32_t* b2, uint32_t* n_bv, uint32_t* n_leaf, void* stream) {
  if (!(radius >= 0)) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "sphere radius must be >= 0");
  return distance_enqueue(m1, m1, n, tf1, tf2, request, min_distance, nearest_p1, nearest_p2, b1, b2, n_bv, n_leaf, stream,
                          radius);
}

// ------------------------------------------------------------------------------------------
// status / host-pointer wrappers
// ------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------
// continuous collision (conservative advancement, translating bodies) -- continuous.cuh
// ------------------------------------------------------------------------------------------
extern "C" int fclgpu_continuous_collide_batch(const fclgpu_model* m1, const fclgpu_model* m2, int64_t n, const double* tf1_beg,
                                               const double* tf1_end, const double* tf2_beg, const double* tf2_end,
                                               const fclgpu_continuous_request* request, int32_t* is_collide,
                                               double* time_of_contact, double* contact_tf1, double* contact_tf2,
                                               int32_t* iterations, uint32_t* n_bv, uint32_t* n_leaf, void* stream) {
  if (!m1 || !m2 || !request || n < 0) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "NULL model/request or n<0");
  if (!is_collide) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "is_collide is NULL (it also holds the start-configuration verdicts)");
  if (request->ccd_motion_type != FCLGPU_CCDM_TRANS)
    return fail(FCLGPU_ERR_UNSUPPORTED_FUNCTION, "only CCDM_TRANS motions are built (the others need sin / cos / atan2: no bit parity)");
  if (request->ccd_solver_type != FCLGPU_CCDC_CONSERVATIVE_ADVANCEMENT)
    return fail(FCLGPU_ERR_UNSUPPORTED_FUNCTION, "only CCDC_CONSERVATIVE_ADVANCEMENT is built");
  if (m1->device != m2->device) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "models live on different devices");
  if (m1->depth + m2->depth + 2 > kStackCap)
    return fail(FCLGPU_ERR_STACK_OVERFLOW, "tree depths %d+%d exceed the traversal stack", m1->depth, m2->depth);
  if (n == 0) return FCLGPU_OK;
  // conservativeAdvancementMeshOriented starts with collide() at the start configuration (default CollisionRequest)
  fclgpu_collision_request cr{1, 0, 0, 0, 0, 0};
  int rc = collide_enqueue(m1, m2, n, tf1_beg, tf2_beg, &cr, is_collide, nullptr, 0, nullptr, nullptr, nullptr, stream, CollideExtra{});
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cudaSetDevice(m1->device));
  Workspace* w;
  rc = get_ws(m1->device, &w);
  if (rc) return rc;
  std::lock_guard<std::mutex> lock(w->mu);
  ContinuousParams P;
  P.m1 = m1->d;
  P.m2 = m2->d;
  P.tf1_beg = tf1_beg;
  P.tf1_end = tf1_end;
  P.tf2_beg = tf2_beg;
  P.tf2_end = tf2_end;
  P.n = n;
  P.start_hits = is_collide;
  P.is_collide = is_collide;
  P.toc = time_of_contact;
  P.contact_tf1 = contact_tf1;
  P.contact_tf2 = contact_tf2;
  P.iterations = iterations;
  P.n_bv = n_bv;
  P.n_leaf = n_leaf;
  P.work_counter = next_counter(w, st);
  P.status = stream_state(w, st).status;
  return (n_bv || n_leaf) ? launch_persistent(ca_translation_kernel<true>, P, w, 128, st)
                          : launch_persistent(ca_translation_kernel<false>, P, w, 128, st);
}

extern "C" int fclgpu_device_trim(int device, int64_t* released) {
  if (released) *released = 0;
  Workspace* w = nullptr;
  {
    std::lock_guard<std::mutex> g(g_ws_mu);
    auto it = g_ws.find(device);
    if (it == g_ws.end()) return FCLGPU_OK;  // nothing was ever allocated on this device
    w = it->second;
  }
  std::lock_guard<std::mutex> host_lock(w->host_mu);
  std::lock_guard<std::mutex> lock(w->mu);
  CUDA_TRY(cudaSetDevice(device));
  CUDA_TRY(cudaDeviceSynchronize());  // nothing in flight may still use the buffers
  int64_t bytes = 0;
  auto drop = [&](void** p, size_t* have) {
    if (*p) cudaFree(*p);
    bytes += (int64_t)*have;
    *p = nullptr;
    *have = 0;
  };
  drop(&w->scratch, &w->scratch_bytes);
  drop(&w->scan_tmp, &w->scan_bytes);
  drop(&w->dev_io, &w->dev_io_bytes);
  drop(&w->order_buf, &w->order_bytes);
  const size_t spill_warps = (size_t)w->sm_count * 8 * kDistWarps;
  for (auto& a : w->spill) {
    if (a.pair) cudaFree(a.pair);
    if (a.bound) cudaFree(a.bound);
    if (a.pair) bytes += (int64_t)(spill_warps * (size_t)a.cap * (sizeof(uint2) + sizeof(float)));
    a.pair = nullptr;
    a.bound = nullptr;
    a.cap = 0;
    a.used = false;
    a.stream = nullptr;
  }
  w->scratch_used = w->order_used = false;
  w->scratch_stream = w->order_stream = nullptr;
  if (released) *released = bytes;
  CUDA_TRY(cudaGetLastError());
  return FCLGPU_OK;
}

extern "C" int fclgpu_sync_status(int device, void* stream) {
  Workspace* w;
  int rc = get_ws(device, &w);
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(device));
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  int* word;
  {
    std::lock_guard<std::mutex> lock(w->mu);
    word = stream_state(w, (cudaStream_t)stream).status;
  }
  int s = 0;
  CUDA_TRY(cudaMemcpy(&s, word, sizeof(int), cudaMemcpyDeviceToHost));
  if (s != 0) {
    CUDA_TRY(cudaMemset(word, 0, sizeof(int)));
    return fail(s, s == FCLGPU_ERR_CONTACT_OVERFLOW ? "contact capacity exceeded (counts are exact; raise contact "
                                                      "capacity or the contact_stride option)"
                   : s == FCLGPU_ERR_INPUT_STALLED  ? "a pose chunk did not reach the device in time (host API input stream)"
                                                    : "traversal stack overflow");
  }
  return FCLGPU_OK;
}

namespace {
struct DevBuf {  // carve typed sub-buffers out of one device allocation (256-byte aligned)
  char* base;
  size_t off = 0;
  template <typename T>
  T* take(size_t count) {
    T* p = reinterpret_cast<T*>(base + off);
    off += (count * sizeof(T) + 255) & ~size_t(255);
    return p;
  }
};
size_t padded(size_t bytes) { return (bytes + 255) & ~size_t(255); }
}  // namespace

namespace {
// queries per pipeline stage of the host API: option "host_chunk" (default 1 << 17 = 12.6 MB of poses)
static int64_t host_chunk() { return std::max<long long>(1024, opt("host_chunk")); }

int finish_pipeline(Workspace* w, int device) {  // both pipeline streams may have run kernels (distance alternates)
  const int r0 = fclgpu_sync_status(device, w->pipe[0]);
  const int r1 = fclgpu_sync_status(device, w->pipe[1]);
  return r0 ? r0 : r1;
}
}  // namespace

extern "C" int fclgpu_collide_batch_host(const fclgpu_model* m1, const fclgpu_model* m2, int64_t n, const double* tf1,
                                         const double* tf2, const fclgpu_collision_request* request,
                                         int32_t* num_contacts, fclgpu_contact* contacts, int64_t contact_capacity,
                                         int64_t* contact_offsets, uint32_t* n_bv, uint32_t* n_leaf) {
  if (!m1 || !m2 || !request || n < 0 || !num_contacts)
    return fail(FCLGPU_ERR_INVALID_ARGUMENT, "NULL model/request/num_contacts or n<0");
  CUDA_TRY(cudaSetDevice(m1->device));
  Workspace* w;
  int rc = get_ws(m1->device, &w);
  if (rc) return rc;
  std::lock_guard<std::mutex> host_lock(w->host_mu);
  const bool want = contacts != nullptr || contact_offsets != nullptr;
  if (contacts == nullptr) contact_capacity = 0;
  const size_t rec = contact_record_bytes(request);  // bytes per record of `contacts` (fclgpu_collision_request::contact_format)
  if (!want && n > 0) {
    // Counts / verdicts only.  ONE persistent launch over the whole batch on the compute stream while the copy
    // stream brings the poses up chunk by chunk; behind every chunk it writes ready[c] (a 4-byte DMA from a pinned
    // word -- not a memset, which could need an SM the spinning kernel holds), and lanes that fetch a query of a
    // chunk still in flight wait on that flag.  No per-chunk launch tails, copies fully overlapped.
    // chunks of host_chunk / 8 queries: no launch depends on the chunk size here, finer chunks only let the kernel
    // start earlier (1M verdicts: 4.55 ms with 131072-query chunks, 4.40 ms with 16384)
    int shift = 10;
    while ((1ll << (shift + 1)) <= host_chunk() / 8) ++shift;
    while (((n + (1ll << shift) - 1) >> shift) > kReadySlots) ++shift;
    const int64_t C = 1ll << shift;
    const int nchunks = (int)((n + C - 1) / C);
    const size_t bytes = (tf1 ? padded(96 * (size_t)n) : 0) + (tf2 ? padded(96 * (size_t)n) : 0) + 3 * padded(4 * (size_t)n) + 256;
    {
      std::lock_guard<std::mutex> lock(w->mu);
      rc = ensure(&w->dev_io, &w->dev_io_bytes, bytes);
      if (rc) return rc;
    }
    DevBuf B{(char*)w->dev_io};
    double* d_tf1 = tf1 ? B.take<double>(12 * (size_t)n) : nullptr;
    double* d_tf2 = tf2 ? B.take<double>(12 * (size_t)n) : nullptr;
    int32_t* d_cnt = B.take<int32_t>((size_t)n);
    uint32_t* d_bv = n_bv ? B.take<uint32_t>((size_t)n) : nullptr;
    uint32_t* d_leaf = n_leaf ? B.take<uint32_t>((size_t)n) : nullptr;
    cudaStream_t copy = w->pipe[0], compute = w->pipe[1];
    CUDA_TRY(cudaMemsetAsync(w->ready, 0, sizeof(unsigned) * nchunks, copy));
    CUDA_TRY(cudaEventRecord(w->ev[0], copy));
    CUDA_TRY(cudaStreamWaitEvent(compute, w->ev[0], 0));  // the kernel must not see flags of an earlier call
    for (int c = 0; c < nchunks; ++c) {
      const int64_t s = (int64_t)c * C;
      const size_t cn = (size_t)std::min<int64_t>(C, n - s);
      if (tf1) CUDA_TRY(cudaMemcpyAsync(d_tf1 + 12 * s, tf1 + 12 * s, 96 * cn, cudaMemcpyHostToDevice, copy));
      if (tf2) CUDA_TRY(cudaMemcpyAsync(d_tf2 + 12 * s, tf2 + 12 * s, 96 * cn, cudaMemcpyHostToDevice, copy));
      CUDA_TRY(cudaMemcpyAsync(w->ready + c, w->host_one, sizeof(unsigned), cudaMemcpyHostToDevice, copy));
    }
    CollideExtra X;
    X.ready = w->ready;
    X.ready_shift = shift;
    rc = collide_enqueue(m1, m2, n, d_tf1, d_tf2, request, d_cnt, nullptr, 0, nullptr, d_bv, d_leaf, compute, X);
    if (rc) {
      cudaStreamSynchronize(copy);  // never leave a kernel waiting for copies that are not coming
      return rc;
    }
    CUDA_TRY(cudaMemcpyAsync(num_contacts, d_cnt, 4 * (size_t)n, cudaMemcpyDeviceToHost, compute));
    if (n_bv) CUDA_TRY(cudaMemcpyAsync(n_bv, d_bv, 4 * (size_t)n, cudaMemcpyDeviceToHost, compute));
    if (n_leaf) CUDA_TRY(cudaMemcpyAsync(n_leaf, d_leaf, 4 * (size_t)n, cudaMemcpyDeviceToHost, compute));
    return finish_pipeline(w, m1->device);
  }
  if (want && n > 0 && request->num_max_contacts > 0) {
    // Contacts wanted.  The result can be gigabytes (64 B per contact), so the copy back is the long pole: the
    // batch runs as sub-batches that append to one contact array (continue_scan), and while sub-batch k computes
    // the copy stream brings sub-batch k-1's contact range down.  The host learns each range from the running
    // total, read back behind the sub-batch.  Poses go up on the copy stream ahead of everything, flagged per
    // chunk like in the counts-only path.
    int shift = 10;
    while ((1ll << (shift + 1)) <= host_chunk()) ++shift;
    while (((n + (1ll << shift) - 1) >> shift) > kTotalSlots) ++shift;
    const int64_t C = 1ll << shift;
    const int nsub = (int)((n + C - 1) / C);
    const size_t bytes = (tf1 ? padded(96 * (size_t)n) : 0) + (tf2 ? padded(96 * (size_t)n) : 0) + padded(4 * (size_t)n) +
                         padded(8 * (size_t)(n + 1)) + padded(rec * (size_t)contact_capacity) +
                         (n_bv ? padded(4 * (size_t)n) : 0) + (n_leaf ? padded(4 * (size_t)n) : 0) + 256;
    {
      std::lock_guard<std::mutex> lock(w->mu);
      rc = ensure(&w->dev_io, &w->dev_io_bytes, bytes);
      if (rc) return rc;
    }
    DevBuf B{(char*)w->dev_io};
    double* d_tf1 = tf1 ? B.take<double>(12 * (size_t)n) : nullptr;
    double* d_tf2 = tf2 ? B.take<double>(12 * (size_t)n) : nullptr;
    int32_t* d_cnt = B.take<int32_t>((size_t)n);
    int64_t* d_off = B.take<int64_t>((size_t)n + 1);
    fclgpu_contact* d_con = contact_capacity > 0 ? reinterpret_cast<fclgpu_contact*>(B.take<char>(rec * (size_t)contact_capacity)) : nullptr;
    uint32_t* d_bv = n_bv ? B.take<uint32_t>((size_t)n) : nullptr;
    uint32_t* d_leaf = n_leaf ? B.take<uint32_t>((size_t)n) : nullptr;
    cudaStream_t copy = w->pipe[0], compute = w->pipe[1];
    CUDA_TRY(cudaMemsetAsync(w->ready, 0, sizeof(unsigned) * nsub, copy));
    CUDA_TRY(cudaEventRecord(w->ev[0], copy));
    CUDA_TRY(cudaStreamWaitEvent(compute, w->ev[0], 0));
    for (int c = 0; c < nsub; ++c) {
      const int64_t s = (int64_t)c * C;
      const size_t cn = (size_t)std::min<int64_t>(C, n - s);
      if (tf1) CUDA_TRY(cudaMemcpyAsync(d_tf1 + 12 * s, tf1 + 12 * s, 96 * cn, cudaMemcpyHostToDevice, copy));
      if (tf2) CUDA_TRY(cudaMemcpyAsync(d_tf2 + 12 * s, tf2 + 12 * s, 96 * cn, cudaMemcpyHostToDevice, copy));
      CUDA_TRY(cudaMemcpyAsync(w->ready + c, w->host_one, sizeof(unsigned), cudaMemcpyHostToDevice, copy));
    }
    long long* cursor_of_compute;
    {
      std::lock_guard<std::mutex> lock(w->mu);
      cursor_of_compute = stream_state(w, compute).cursor;
    }
    int64_t done = 0;  // contacts already on their way to the host
    auto drain = [&](int k) -> int {  // copy sub-batch k's contact range once its running total is known
      CUDA_TRY(cudaEventSynchronize(w->ev[k & 1]));
      const int64_t upto = std::min<int64_t>(w->host_totals[k], contact_capacity);
      if (contacts && upto > done)
        CUDA_TRY(cudaMemcpyAsync((char*)contacts + rec * (size_t)done, (const char*)d_con + rec * (size_t)done, rec * (size_t)(upto - done),
                                 cudaMemcpyDeviceToHost, copy));
      done = std::max(done, upto);
      return 0;
    };
    for (int k = 0; k < nsub; ++k) {
      const int64_t s = (int64_t)k * C;
      const int64_t cn = std::min<int64_t>(C, n - s);
      CollideExtra X;
      X.ready = w->ready;
      X.ready_shift = shift;
      X.ready_q0 = s;
      X.continue_scan = k > 0;
      rc = collide_enqueue(m1, m2, cn, d_tf1 ? d_tf1 + 12 * s : nullptr, d_tf2 ? d_tf2 + 12 * s : nullptr, request,
                           d_cnt + s, d_con, contact_capacity, d_off + s, d_bv ? d_bv + s : nullptr,
                           d_leaf ? d_leaf + s : nullptr, compute, X);
      if (rc) {
        cudaStreamSynchronize(copy);
        cudaStreamSynchronize(compute);
        return rc;
      }
      CUDA_TRY(cudaMemcpyAsync(w->host_totals + k, cursor_of_compute, sizeof(long long), cudaMemcpyDeviceToHost, compute));
      CUDA_TRY(cudaEventRecord(w->ev[k & 1], compute));
      if (k >= 1 && (rc = drain(k - 1))) return rc;
    }
    CUDA_TRY(cudaMemcpyAsync(num_contacts, d_cnt, 4 * (size_t)n, cudaMemcpyDeviceToHost, compute));
    if (contact_offsets) CUDA_TRY(cudaMemcpyAsync(contact_offsets, d_off, 8 * (size_t)(n + 1), cudaMemcpyDeviceToHost, compute));
    if (n_bv) CUDA_TRY(cudaMemcpyAsync(n_bv, d_bv, 4 * (size_t)n, cudaMemcpyDeviceToHost, compute));
    if (n_leaf) CUDA_TRY(cudaMemcpyAsync(n_leaf, d_leaf, 4 * (size_t)n, cudaMemcpyDeviceToHost, compute));
    if ((rc = drain(nsub - 1))) return rc;
    return finish_pipeline(w, m1->device);
  }
  size_t bytes = (tf1 ? padded(96 * (size_t)n) : 0) + (tf2 ? padded(96 * (size_t)n) : 0) + padded(4 * (size_t)n) +
                 (want ? padded(8 * (size_t)(n + 1)) + padded(rec * (size_t)contact_capacity) : 0) +
                 (n_bv ? padded(4 * (size_t)n) : 0) + (n_leaf ? padded(4 * (size_t)n) : 0) + 256;
  {
    std::lock_guard<std::mutex> lock(w->mu);
    rc = ensure(&w->dev_io, &w->dev_io_bytes, bytes);
    if (rc) return rc;
  }
  DevBuf B{(char*)w->dev_io};
  double* d_tf1 = tf1 ? B.take<double>(12 * (size_t)n) : nullptr;
  double* d_tf2 = tf2 ? B.take<double>(12 * (size_t)n) : nullptr;
  int32_t* d_cnt = B.take<int32_t>((size_t)n);
  int64_t* d_off = want ? B.take<int64_t>((size_t)n + 1) : nullptr;
  fclgpu_contact* d_con = (want && contact_capacity > 0) ? reinterpret_cast<fclgpu_contact*>(B.take<char>(rec * (size_t)contact_capacity)) : nullptr;
  uint32_t* d_bv = n_bv ? B.take<uint32_t>((size_t)n) : nullptr;
  uint32_t* d_leaf = n_leaf ? B.take<uint32_t>((size_t)n) : nullptr;
  cudaStream_t st = 0;
  if (tf1) CUDA_TRY(cudaMemcpyAsync(d_tf1, tf1, 96 * (size_t)n, cudaMemcpyHostToDevice, st));
  if (tf2) CUDA_TRY(cudaMemcpyAsync(d_tf2, tf2, 96 * (size_t)n, cudaMemcpyHostToDevice, st));
  rc = fclgpu_collide_batch(m1, m2, n, d_tf1, d_tf2, request, d_cnt, d_con, contact_capacity, d_off, d_bv, d_leaf, st);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(num_contacts, d_cnt, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
  if (n_bv) CUDA_TRY(cudaMemcpyAsync(n_bv, d_bv, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
  if (n_leaf) CUDA_TRY(cudaMemcpyAsync(n_leaf, d_leaf, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
  int64_t total = 0;
  if (want && n > 0 && request->num_max_contacts > 0) {
    if (contact_offsets) CUDA_TRY(cudaMemcpyAsync(contact_offsets, d_off, 8 * (size_t)(n + 1), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpy(&total, d_off + n, 8, cudaMemcpyDeviceToHost));
    if (contacts && total > 0)
      CUDA_TRY(cudaMemcpyAsync(contacts, d_con, rec * (size_t)std::min<int64_t>(total, contact_capacity),
                               cudaMemcpyDeviceToHost, st));
  } else if (contact_offsets) {
    std::memset(contact_offsets, 0, 8 * (size_t)(n + 1));
  }
  return fclgpu_sync_status(m1->device, st);
}

namespace {
int distance_host(const fclgpu_model* m1, const fclgpu_model* m2, int64_t n, const double* tf1, const double* tf2,
                  const fclgpu_distance_request* request, double* min_distance, double* nearest_p1, double* nearest_p2,
                  int32_t* b1, int32_t* b2, uint32_t* n_bv, uint32_t* n_leaf, double sphere_radius,
                  double cutoff = 1.7976931348623157e308, double stop_below = -1.0, uint8_t* within = nullptr) {
  if (!m1 || !m2 || !request || n < 0) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "NULL model/request or n<0");
  CUDA_TRY(cudaSetDevice(m1->device));
  Workspace* w;
  int rc = get_ws(m1->device, &w);
  if (rc) return rc;
  std::lock_guard<std::mutex> host_lock(w->host_mu);
  if (n == 0) return FCLGPU_OK;
  // Chunked two-stream pipeline: while chunk c computes, chunk c+1's poses go up and chunk
  // c-1's results come down (the copies are asynchronous when the caller's buffers are pinned).
  const size_t C = (size_t)std::min<int64_t>(n, host_chunk());
  const size_t per_stage = (tf1 ? padded(96 * C) : 0) + (tf2 ? padded(96 * C) : 0) + padded(8 * C) + 2 * padded(24 * C) +
                           4 * padded(4 * C) + padded(C) + 256;
  {
    std::lock_guard<std::mutex> lock(w->mu);
    rc = ensure(&w->dev_io, &w->dev_io_bytes, 2 * per_stage);
    if (rc) return rc;
  }
  const bool np = request->enable_nearest_points != 0;
  int stage = 0;
  for (int64_t s = 0; s < n; s += (int64_t)C, stage ^= 1) {
    const size_t cn = (size_t)std::min<int64_t>((int64_t)C, n - s);
    cudaStream_t st = w->pipe[stage];
    DevBuf B{(char*)w->dev_io + stage * per_stage};
    double* d_tf1 = tf1 ? B.take<double>(12 * C) : nullptr;
    double* d_tf2 = tf2 ? B.take<double>(12 * C) : nullptr;
    double* d_dist = B.take<double>(C);
    double* d_p1 = (nearest_p1 && np) ? B.take<double>(3 * C) : nullptr;
    double* d_p2 = (nearest_p2 && np) ? B.take<double>(3 * C) : nullptr;
    int32_t* d_b1 = b1 ? B.take<int32_t>(C) : nullptr;
    int32_t* d_b2 = b2 ? B.take<int32_t>(C) : nullptr;
    uint32_t* d_bv = n_bv ? B.take<uint32_t>(C) : nullptr;
    uint32_t* d_leaf = n_leaf ? B.take<uint32_t>(C) : nullptr;
    uint8_t* d_within = within ? B.take<uint8_t>(C) : nullptr;
    if (tf1) CUDA_TRY(cudaMemcpyAsync(d_tf1, tf1 + 12 * s, 96 * cn, cudaMemcpyHostToDevice, st));
    if (tf2) CUDA_TRY(cudaMemcpyAsync(d_tf2, tf2 + 12 * s, 96 * cn, cudaMemcpyHostToDevice, st));
    rc = distance_enqueue(m1, m2, (int64_t)cn, d_tf1, d_tf2, request, d_dist, d_p1, d_p2, d_b1, d_b2, d_bv, d_leaf, st,
                          sphere_radius, cutoff, stop_below, d_within);
    if (rc) return rc;
    if (within) CUDA_TRY(cudaMemcpyAsync(within + s, d_within, cn, cudaMemcpyDeviceToHost, st));
    if (min_distance) CUDA_TRY(cudaMemcpyAsync(min_distance + s, d_dist, 8 * cn, cudaMemcpyDeviceToHost, st));
    if (d_p1) CUDA_TRY(cudaMemcpyAsync(nearest_p1 + 3 * s, d_p1, 24 * cn, cudaMemcpyDeviceToHost, st));
    if (d_p2) CUDA_TRY(cudaMemcpyAsync(nearest_p2 + 3 * s, d_p2, 24 * cn, cudaMemcpyDeviceToHost, st));
    if (b1) CUDA_TRY(cudaMemcpyAsync(b1 + s, d_b1, 4 * cn, cudaMemcpyDeviceToHost, st));
    if (b2) CUDA_TRY(cudaMemcpyAsync(b2 + s, d_b2, 4 * cn, cudaMemcpyDeviceToHost, st));
    if (n_bv) CUDA_TRY(cudaMemcpyAsync(n_bv + s, d_bv, 4 * cn, cudaMemcpyDeviceToHost, st));
    if (n_leaf) CUDA_TRY(cudaMemcpyAsync(n_leaf + s, d_leaf, 4 * cn, cudaMemcpyDeviceToHost, st));
    // a stage's device buffers are reused two chunks later: that chunk is enqueued on the same
    // stream, so stream order already protects them
  }
  return finish_pipeline(w, m1->device);
}
}  // namespace

extern "C" int fclgpu_distance_batch_host(const fclgpu_model* m1, const fclgpu_model* m2, int64_t n,
                                          const double* tf1, const double* tf2,
                                          const fclgpu_distance_request* request, double* min_distance,
                                          double* nearest_p1, double* nearest_p2, int32_t* b1, int32_t* b2,
                                          uint32_t* n_bv, uint32_t* n_leaf) {
  return distance_host(m1, m2, n, tf1, tf2, request, min_distance, nearest_p1, nearest_p2, b1, b2, n_bv, n_leaf, -1.0);
}

extern "C" int fclgpu_distance_cutoff_batch_host(const fclgpu_model* m1, const fclgpu_model* m2, int64_t n,
                                                 const double* tf1, const double* tf2,
                                                 const fclgpu_distance_request* request, double cutoff,
                                                 double* min_distance, double* nearest_p1, double* nearest_p2, int32_t* b1,
                                                 int32_t* b2, uint32_t* n_bv, uint32_t* n_leaf) {
  if (!(cutoff > 0)) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "cutoff must be > 0");
  return distance_host(m1, m2, n, tf1, tf2, request, min_distance, nearest_p1, nearest_p2, b1, b2, n_bv, n_leaf, -1.0, cutoff);
}

extern "C" int fclgpu_within_tolerance_batch_host(const fclgpu_model* m1, const fclgpu_model* m2, int64_t n, const double* tf1,
                                                  const double* tf2, double tolerance, uint8_t* within,
                                                  double* witness_distance, uint32_t* n_bv, uint32_t* n_leaf) {
  if (!(tolerance >= 0)) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "tolerance must be >= 0");
  if (!within) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "within is NULL");
  const fclgpu_distance_request rq{0, 0, 0.0, 0.0};
  return distance_host(m1, m2, n, tf1, tf2, &rq, witness_distance, nullptr, nullptr, nullptr, nullptr, n_bv, n_leaf, -1.0,
                       std::nextafter(tolerance, 1.7976931348623157e308), tolerance, within);
}

extern "C" int fclgpu_distance_mesh_sphere_batch_host(const fclgpu_model* m1, double radius, int64_t n, const double* tf1,
                                                      const double* tf2, const fclgpu_distance_request* request,
                                                      double* min_distance, double* nearest_p1, double* nearest_p2,
                                                      int32_t* b1, int32_t* b2, uint32_t* n_bv, uint32_t* n_leaf) {
  if (!(radius >= 0)) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "sphere radius must be >= 0");
  return distance_host(m1, m1, n, tf1, tf2, request, min_distance, nearest_p1, nearest_p2, b1, b2, n_bv, n_leaf, radius);
}

namespace {
// host-pointer wrapper shared by the mesh <-> primitive-shape collide entry points (X names the shape)
int mesh_shape_collide_host(const fclgpu_model* m1, const CollideExtra& X, int64_t n, const double* tf1, const double* tf2,
                            const fclgpu_collision_request* request, int32_t* num_contacts, fclgpu_contact* contacts,
                            int64_t contact_capacity, int64_t* contact_offsets, uint32_t* n_bv, uint32_t* n_leaf) {
  if (!m1 || !request || n < 0 || !num_contacts) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "NULL model/request/num_contacts or n<0");
  CUDA_TRY(cudaSetDevice(m1->device));
  Workspace* w;
  int rc = get_ws(m1->device, &w);
  if (rc) return rc;
  std::lock_guard<std::mutex> host_lock(w->host_mu);
  const bool want = contacts != nullptr || contact_offsets != nullptr;
  if (contacts == nullptr) contact_capacity = 0;
  const size_t bytes = (tf1 ? padded(96 * (size_t)n) : 0) + (tf2 ? padded(96 * (size_t)n) : 0) + padded(4 * (size_t)n) +
                       (want ? padded(8 * (size_t)(n + 1)) + padded(64 * (size_t)contact_capacity) : 0) +
                       (n_bv ? padded(4 * (size_t)n) : 0) + (n_leaf ? padded(4 * (size_t)n) : 0) + 256;
  {
    std::lock_guard<std::mutex> lock(w->mu);
    rc = ensure(&w->dev_io, &w->dev_io_bytes, bytes);
    if (rc) return rc;
  }
  DevBuf B{(char*)w->dev_io};
  double* d_tf1 = tf1 ? B.take<double>(12 * (size_t)n) : nullptr;
  double* d_tf2 = tf2 ? B.take<double>(12 * (size_t)n) : nullptr;
  int32_t* d_cnt = B.take<int32_t>((size_t)n);
  int64_t* d_off = want ? B.take<int64_t>((size_t)n + 1) : nullptr;
  fclgpu_contact* d_con = (want && contact_capacity > 0) ? B.take<fclgpu_contact>((size_t)contact_capacity) : nullptr;
  uint32_t* d_bv = n_bv ? B.take<uint32_t>((size_t)n) : nullptr;
  uint32_t* d_leaf = n_leaf ? B.take<uint32_t>((size_t)n) : nullptr;
  cudaStream_t st = w->pipe[1];
  if (tf1) CUDA_TRY(cudaMemcpyAsync(d_tf1, tf1, 96 * (size_t)n, cudaMemcpyHostToDevice, st));
  if (tf2) CUDA_TRY(cudaMemcpyAsync(d_tf2, tf2, 96 * (size_t)n, cudaMemcpyHostToDevice, st));
  rc = collide_enqueue(m1, m1, n, d_tf1, d_tf2, request, d_cnt, d_con, contact_capacity, d_off, d_bv, d_leaf, st, X);
  if (rc) return rc;
  if (n > 0) CUDA_TRY(cudaMemcpyAsync(num_contacts, d_cnt, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
  if (n_bv && n > 0) CUDA_TRY(cudaMemcpyAsync(n_bv, d_bv, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
  if (n_leaf && n > 0) CUDA_TRY(cudaMemcpyAsync(n_leaf, d_leaf, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
  if (want && n > 0 && request->num_max_contacts > 0) {
    if (contact_offsets) CUDA_TRY(cudaMemcpyAsync(contact_offsets, d_off, 8 * (size_t)(n + 1), cudaMemcpyDeviceToHost, st));
    int64_t total = 0;
    CUDA_TRY(cudaMemcpyAsync(&total, d_off + n, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (contacts && total > 0)
      CUDA_TRY(cudaMemcpyAsync(contacts, d_con, 64 * (size_t)std::min<int64_t>(total, contact_capacity), cudaMemcpyDeviceToHost, st));
  } else if (contact_offsets) {
    std::memset(contact_offsets, 0, 8 * (size_t)(n + 1));
  }
  return fclgpu_sync_status(m1->device, st);
}
}  // namespace

extern "C" int fclgpu_collide_mesh_sphere_batch_host(const fclgpu_model* m1, double radius, int64_t n, const double* tf1,
                                                     const double* tf2, const fclgpu_collision_request* request,
                                                     int32_t* num_contacts, fclgpu_contact* contacts,
                                                     int64_t contact_capacity, int64_t* contact_offsets, uint32_t* n_bv,
                                                     uint32_t* n_leaf) {
  if (!(radius >= 0)) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "sphere radius must be >= 0");
  CollideExtra X;
  X.sphere_radius = radius;
  return mesh_shape_collide_host(m1, X, n, tf1, tf2, request, num_contacts, contacts, contact_capacity, contact_offsets, n_bv, n_leaf);
}

extern "C" int fclgpu_collide_mesh_plane_batch_host(const fclgpu_model* m1, int32_t shape, const double* normal3, double d, int64_t n,
                                                    const double* tf1, const double* tf2, const fclgpu_collision_request* request,
                                                    int32_t* num_contacts, fclgpu_contact* contacts, int64_t contact_capacity,
                                                    int64_t* contact_offsets, uint32_t* n_bv, uint32_t* n_leaf) {
  if (shape != FCLGPU_SHAPE_HALFSPACE && shape != FCLGPU_SHAPE_PLANE) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "shape must be FCLGPU_SHAPE_HALFSPACE or FCLGPU_SHAPE_PLANE");
  CollideExtra X;
  const int rc = plane_extra(shape == FCLGPU_SHAPE_HALFSPACE ? 0 : 1, normal3, d, X);
  if (rc) return rc;
  return mesh_shape_collide_host(m1, X, n, tf1, tf2, request, num_contacts, contacts, contact_capacity, contact_offsets, n_bv, n_leaf);
}


// ------------------------------------------------------------------------------------------
// broadphase: N x M AABB culling feeding the batched mesh-mesh kernel
// ------------------------------------------------------------------------------------------
extern "C" int fclgpu_continuous_collide_batch_host(const fclgpu_model* m1, const fclgpu_model* m2, int64_t n, const double* tf1_beg,
                                                    const double* tf1_end, const double* tf2_beg, const double* tf2_end,
                                                    const fclgpu_continuous_request* request, int32_t* is_collide,
                                                    double* time_of_contact, double* contact_tf1, double* contact_tf2,
                                                    int32_t* iterations) {
  if (!m1 || !m2 || !request || n < 0) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "NULL model/request or n<0");
  if (n == 0) return FCLGPU_OK;
  CUDA_TRY(cudaSetDevice(m1->device));
  Workspace* w;
  int rc = get_ws(m1->device, &w);
  if (rc) return rc;
  std::lock_guard<std::mutex> host_lock(w->host_mu);
  // one staging block: 4 pose arrays in, verdict / toc / 2 contact poses / iterations out
  const size_t N = (size_t)n;
  char* dev = nullptr;
  const size_t bytes = 4 * padded(96 * N) + padded(4 * N) + padded(8 * N) + 2 * padded(96 * N) + padded(4 * N) + 256;
  CUDA_TRY(cudaMalloc((void**)&dev, bytes));
  DevBuf B{dev};
  double* d_in[4];
  const double* h_in[4] = {tf1_beg, tf1_end, tf2_beg, tf2_end};
  cudaStream_t st = w->pipe[0];
  auto bail = [&](int code) {
    cudaStreamSynchronize(st);
    cudaFree(dev);
    return code;
  };
  for (int k = 0; k < 4; ++k) {
    d_in[k] = h_in[k] ? B.take<double>(12 * N) : nullptr;
    if (h_in[k] && cudaMemcpyAsync(d_in[k], h_in[k], 96 * N, cudaMemcpyHostToDevice, st) != cudaSuccess)
      return bail(fail(FCLGPU_ERR_UNKNOWN, "pose upload failed"));
  }
  int32_t* d_hit = B.take<int32_t>(N);
  double* d_toc = B.take<double>(N);
  double* d_c1 = contact_tf1 ? B.take<double>(12 * N) : nullptr;
  double* d_c2 = contact_tf2 ? B.take<double>(12 * N) : nullptr;
  int32_t* d_it = iterations ? B.take<int32_t>(N) : nullptr;
  rc = fclgpu_continuous_collide_batch(m1, m2, n, d_in[0], d_in[1], d_in[2], d_in[3], request, d_hit, d_toc, d_c1, d_c2, d_it, nullptr,
                                       nullptr, st);
  if (rc) return bail(rc);
  cudaError_t e = cudaSuccess;
  if (is_collide) e = cudaMemcpyAsync(is_collide, d_hit, 4 * N, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && time_of_contact) e = cudaMemcpyAsync(time_of_contact, d_toc, 8 * N, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && d_c1) e = cudaMemcpyAsync(contact_tf1, d_c1, 96 * N, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && d_c2) e = cudaMemcpyAsync(contact_tf2, d_c2, 96 * N, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && d_it) e = cudaMemcpyAsync(iterations, d_it, 4 * N, cudaMemcpyDeviceToHost, st);
  if (e != cudaSuccess) return bail(fail(FCLGPU_ERR_UNKNOWN, "result download failed: %s", cudaGetErrorString(e)));
  rc = fclgpu_sync_status(m1->device, st);
  cudaFree(dev);
  return rc;
}

extern "C" int fclgpu_model_local_aabb(const fclgpu_model* m, double center3[3], double* radius, double min3[3], double max3[3]) {
  if (!m) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "NULL model");
  for (int k = 0; k < 3; ++k) {
    if (center3) center3[k] = m->aabb.c[k];
    if (min3) min3[k] = m->aabb.mn[k];
    if (max3) max3[k] = m->aabb.mx[k];
  }
  if (radius) *radius = m->aabb.r;
  return FCLGPU_OK;
}

extern "C" int fclgpu_broadphase_collide_host(int32_t n_geoms, const fclgpu_model* const* geoms, int64_t n1, const int32_t* geom1,
                                              const double* tf1, int64_t n2, const int32_t* geom2, const double* tf2,
                                              const fclgpu_collision_request* request, int64_t pair_capacity, int32_t* pairs,
                                              int64_t* num_pairs, int32_t* num_contacts, double* aabb1_out, double* aabb2_out) {
  if (!geoms || n_geoms <= 0 || n1 < 0 || n2 < 0 || !geom1 || !geom2 || !tf1 || !tf2 || !num_pairs || pair_capacity < 0)
    return fail(FCLGPU_ERR_INVALID_ARGUMENT, "NULL / negative argument");
  if (n1 > 0x7fffffff || n2 > 0x7fffffff) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "too many objects");
  if (num_contacts && !request) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "request is NULL");
  *num_pairs = 0;
  for (int g = 0; g < n_geoms; ++g)
    if (!geoms[g] || geoms[g]->device != geoms[0]->device) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "NULL geometry / mixed devices");
  for (int64_t i = 0; i < n1; ++i)
    if (geom1[i] < 0 || geom1[i] >= n_geoms) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "geom1[%lld] out of range", (long long)i);
  for (int64_t j = 0; j < n2; ++j)
    if (geom2[j] < 0 || geom2[j] >= n_geoms) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "geom2[%lld] out of range", (long long)j);
  if (n1 == 0 || n2 == 0) return FCLGPU_OK;
  const int device = geoms[0]->device;
  CUDA_TRY(cudaSetDevice(device));
  Workspace* w;
  int rc = get_ws(device, &w);
  if (rc) return rc;
  std::lock_guard<std::mutex> host_lock(w->host_mu);
  cudaStream_t st = w->pipe[1];
  const int64_t cap = std::max<int64_t>(pair_capacity, 1);
  const int64_t nscan = std::max<int64_t>(std::max(n1, cap), 1);
  const int64_t nblk = (nscan + kScanBlock - 1) / kScanBlock;
  const size_t bytes = padded(sizeof(LocalAabb) * n_geoms) + padded(4 * n1) + padded(4 * n2) + padded(96 * n1) + padded(96 * n2) +
                       padded(48 * n1) + padded(48 * n2) + padded(4 * nscan) + padded(8 * nscan) + padded(8 * nblk) + padded(8 * cap) +
                       2 * padded(96 * cap) + padded(8 * cap) + 2 * padded(4 * cap) + padded(8) + 4096;
  {
    std::lock_guard<std::mutex> lock(w->mu);
    rc = ensure(&w->dev_io, &w->dev_io_bytes, bytes);
    if (rc) return rc;
  }
  DevBuf B{(char*)w->dev_io};
  LocalAabb* d_loc = B.take<LocalAabb>(n_geoms);
  int32_t* d_g1 = B.take<int32_t>(n1);
  int32_t* d_g2 = B.take<int32_t>(n2);
  double* d_tf1 = B.take<double>(12 * n1);
  double* d_tf2 = B.take<double>(12 * n2);
  double* d_a1 = B.take<double>(6 * n1);
  double* d_a2 = B.take<double>(6 * n2);
  int32_t* d_cnt = B.take<int32_t>(nscan);      // per-object overlap counts, later the group flags
  long long* d_local = B.take<long long>(nscan);
  long long* d_bsum = B.take<long long>(nblk);
  int2* d_pairs = B.take<int2>(cap);
  double* d_gtf1 = B.take<double>(12 * cap);
  double* d_gtf2 = B.take<double>(12 * cap);
  long long* d_gidx = B.take<long long>(cap);
  int32_t* d_gcnt = B.take<int32_t>(cap);
  int32_t* d_out = B.take<int32_t>(cap);
  long long* d_base = B.take<long long>(1);

  std::vector<LocalAabb> loc(n_geoms);
  for (int g = 0; g < n_geoms; ++g) loc[g] = geoms[g]->aabb;
  CUDA_TRY(cudaMemcpyAsync(d_loc, loc.data(), sizeof(LocalAabb) * n_geoms, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(d_g1, geom1, 4 * n1, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(d_g2, geom2, 4 * n2, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(d_tf1, tf1, 96 * n1, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(d_tf2, tf2, 96 * n2, cudaMemcpyHostToDevice, st));
  world_aabb_kernel<<<(unsigned)((n1 + 255) / 256), 256, 0, st>>>(d_loc, d_g1, d_tf1, n1, d_a1);
  world_aabb_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(d_loc, d_g2, d_tf2, n2, d_a2);
  const unsigned pair_blocks = (unsigned)((n1 * 32 + 255) / 256);
  pair_kernel<false><<<pair_blocks, 256, 0, st>>>(d_a1, n1, d_a2, n2, d_cnt, nullptr, nullptr, kScanBlock, nullptr, 0);
  auto scan = [&](int64_t n) {  // exclusive scan of d_cnt[0, n): offset(i) = d_bsum[i / kScanBlock] + d_local[i]; total -> d_base
    const int nb = (int)((n + kScanBlock - 1) / kScanBlock);
    cudaMemsetAsync(d_base, 0, sizeof(long long), st);
    scan_block_kernel<<<nb, kScanBlock, 0, st>>>(d_cnt, n, (long long)0x7fffffff, d_local, d_bsum);
    scan_sums_kernel<<<1, kScanBlock, 0, st>>>(d_bsum, nb, d_base, nullptr);
    g_launches += 2;
  };
  scan(n1);
  pair_kernel<true><<<pair_blocks, 256, 0, st>>>(d_a1, n1, d_a2, n2, nullptr, d_local, d_bsum, kScanBlock, d_pairs, cap);
  g_launches += 4;
  long long total = 0;
  CUDA_TRY(cudaMemcpyAsync(&total, d_base, sizeof(long long), cudaMemcpyDeviceToHost, st));
  if (aabb1_out) CUDA_TRY(cudaMemcpyAsync(aabb1_out, d_a1, 48 * n1, cudaMemcpyDeviceToHost, st));
  if (aabb2_out) CUDA_TRY(cudaMemcpyAsync(aabb2_out, d_a2, 48 * n2, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  CUDA_TRY(cudaGetLastError());
  *num_pairs = total;
  if (total > pair_capacity) return fail(FCLGPU_ERR_CONTACT_OVERFLOW, "pair capacity %lld < %lld overlapping pairs", (long long)pair_capacity, total);
  if (total == 0) return FCLGPU_OK;
  if (pairs) CUDA_TRY(cudaMemcpyAsync(pairs, d_pairs, 8 * (size_t)total, cudaMemcpyDeviceToHost, st));
  if (num_contacts) {
    // narrowphase per (geometry 1, geometry 2) group: stable device compaction, pose gather, batched collide, scatter
    std::vector<char> used1(n_geoms, 0), used2(n_geoms, 0);
    for (int64_t i = 0; i < n1; ++i) used1[geom1[i]] = 1;
    for (int64_t j = 0; j < n2; ++j) used2[geom2[j]] = 1;
    const unsigned tb = (unsigned)((total + 255) / 256);
    CUDA_TRY(cudaMemsetAsync(d_out, 0, 4 * (size_t)total, st));
    for (int ga = 0; ga < n_geoms; ++ga) {
      if (!used1[ga]) continue;
      for (int gb = 0; gb < n_geoms; ++gb) {
        if (!used2[gb]) continue;
        group_flag_kernel<<<tb, 256, 0, st>>>(d_pairs, total, d_g1, d_g2, ga, gb, d_cnt);
        scan(total);
        long long m = 0;
        CUDA_TRY(cudaMemcpyAsync(&m, d_base, sizeof(long long), cudaMemcpyDeviceToHost, st));
        group_gather_kernel<<<tb, 256, 0, st>>>(d_pairs, total, d_cnt, d_local, d_bsum, kScanBlock, d_tf1, d_tf2, d_gtf1, d_gtf2, d_gidx);
        g_launches += 2;
        CUDA_TRY(cudaStreamSynchronize(st));
        if (m == 0) continue;
        rc = fclgpu_collide_batch(geoms[ga], geoms[gb], m, d_gtf1, d_gtf2, request, d_gcnt, nullptr, 0, nullptr, nullptr, nullptr, st);
        if (rc) return rc;
        group_scatter_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(d_gidx, d_gcnt, m, d_out);
        g_launches++;
      }
    }
    CUDA_TRY(cudaMemcpyAsync(num_contacts, d_out, 4 * (size_t)total, cudaMemcpyDeviceToHost, st));
  }
  CUDA_TRY(cudaGetLastError());
  return fclgpu_sync_status(device, st);
}

// ------------------------------------------------------------------------------------------
// utilities
// ------------------------------------------------------------------------------------------
extern "C" int fclgpu_abi_version(void) { return FCLGPU_ABI_VERSION; }

extern "C" int fclgpu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

extern "C" const char* fclgpu_last_error(void) { return g_err; }

extern "C" void fclgpu_pose_from_colmajor4x4(const double* m, double* p) {
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) p[3 * r + c] = m[4 * c + r];
    p[9 + r] = m[12 + r];
  }
}

extern "C" int fclgpu_set_option(const char* name, int64_t value) {
  if (!name) return FCLGPU_ERR_INVALID_ARGUMENT;
  std::lock_guard<std::mutex> g(g_opt_mu);
  auto it = g_opts.find(name);
  if (it == g_opts.end()) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "unknown option %s", name);
  it->second = value;
  return FCLGPU_OK;
}
extern "C" int64_t fclgpu_get_option(const char* name) { return name ? opt(name) : 0; }
extern "C" int64_t fclgpu_launch_count(void) { return g_launches.load(); }
// development counters (all zero unless the library was built with an instrumentation flag such as -DFCLGPU_DIST_PROF=1)
extern "C" int fclgpu_debug_counters(int device, uint64_t* out16, int reset) {
  if (!out16) return fail(FCLGPU_ERR_INVALID_ARGUMENT, "out16 is NULL");
  CUDA_TRY(cudaSetDevice(device));
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpyFromSymbol(out16, g_debug_counters, sizeof(uint64_t) * 16));
  if (reset) {
    uint64_t z[16] = {0};
    CUDA_TRY(cudaMemcpyToSymbol(g_debug_counters, z, sizeof(z)));
  }
  return 0;
}

// ------------------------------------------------------------------------------------------
// micro-benchmarks for the roofline denominators
// ------------------------------------------------------------------------------------------
namespace {
template <int MODE>
__global__ void fp64_peak_kernel(double* out, int iters, double seed) {
  // 8 independent dependency chains per thread
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {  // separately rounded multiply then add
      a0 = __dadd_rn(__dmul_rn(a0, m), c); a1 = __dadd_rn(__dmul_rn(a1, m), c);
      a2 = __dadd_rn(__dmul_rn(a2, m), c); a3 = __dadd_rn(__dmul_rn(a3, m), c);
      a4 = __dadd_rn(__dmul_rn(a4, m), c); a5 = __dadd_rn(__dmul_rn(a5, m), c);
      a6 = __dadd_rn(__dmul_rn(a6, m), c); a7 = __dadd_rn(__dmul_rn(a7, m), c);
    } else {
      a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
      a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

__global__ void l2_read_kernel(const double2* __restrict__ buf, size_t n16, int reps, double* out) {
  double acc = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int r = 0; r < reps; ++r)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
      double2 v = __ldcg(buf + i);  // L2-cached, bypass L1
      acc += v.x + v.y;
    }
  if (acc == 123.456) out[0] = acc;
}
}  // namespace

extern "C" int fclgpu_microbench(int device, int kind, double* result) {
  if (!result) return FCLGPU_ERR_INVALID_ARGUMENT;
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  float ms = 0;
  if (kind == 0 || kind == 1) {
    const int block = 256, grid = prop.multiProcessorCount * 8, iters = 20000;
    double* out;
    CUDA_TRY(cudaMalloc(&out, sizeof(double) * grid * block));
    for (int rep = 0; rep < 3; ++rep) {
      CUDA_TRY(cudaEventRecord(e0));
      if (kind == 0) fp64_peak_kernel<0><<<grid, block>>>(out, iters, 1.0);
      else fp64_peak_kernel<1><<<grid, block>>>(out, iters, 1.0);
      CUDA_TRY(cudaEventRecord(e1));
      CUDA_TRY(cudaEventSynchronize(e1));
      CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    }
    const double ops = (double)grid * block * iters * 8.0 * (kind == 0 ? 2.0 : 1.0);
    *result = ops / (ms * 1e-3);
    cudaFree(out);
  } else if (kind == 2) {
    const size_t bytes = 32ull << 20;  // 32 MiB: L2 resident
    double2* buf;
    double* out;
    CUDA_TRY(cudaMalloc(&buf, bytes));
    CUDA_TRY(cudaMalloc(&out, 8));
    CUDA_TRY(cudaMemset(buf, 0, bytes));
    const int reps = 50;
    for (int rep = 0; rep < 3; ++rep) {
      CUDA_TRY(cudaEventRecord(e0));
      l2_read_kernel<<<prop.multiProcessorCount * 8, 256>>>(buf, bytes / 16, reps, out);
      CUDA_TRY(cudaEventRecord(e1));
      CUDA_TRY(cudaEventSynchronize(e1));
      CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    }
    *result = (double)bytes * reps / (ms * 1e-3) / 1e9;
    cudaFree(buf);
    cudaFree(out);
  } else {
    return fail(FCLGPU_ERR_INVALID_ARGUMENT, "unknown microbench kind %d", kind);
  }
  g_launches += 3;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return FCLGPU_OK;
}
