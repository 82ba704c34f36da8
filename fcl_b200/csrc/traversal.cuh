// traversal.cuh — BVTT traversal kernels for BVHModel<OBBRSS<double>> pairs.
//
// Reference semantics implemented here (file:line under /root/reference/include/fcl):
//   collisionRecurse        narrowphase/detail/traversal/traversal_recurse-inl.h:84-130
//   distanceRecurse         narrowphase/detail/traversal/traversal_recurse-inl.h:259-316
//   firstOverSecond         narrowphase/detail/traversal/collision/bvh_collision_traversal_node-inl.h:78-90
//   leaf tests              .../collision/mesh_collision_traversal_node-inl.h:527-620,
//                           .../distance/mesh_distance_traversal_node-inl.h:453-499, 546-603
//   relative pose           math/geometry-inl.h:681-682 (collide), mesh_distance_traversal_node-inl.h:630 (distance)
#pragma once
#include <cstdint>

#include "../../include/fclgpu.h"
#include "bounds_f32.cuh"
#include "device_math.cuh"
#include "mesh_sphere.cuh"

namespace fclgpu {

// ---------------------------------------------------------------------------------------
// HBM layout.  Every record is a multiple of 16 bytes and 16-byte aligned,
// fetched with LDG.128:
//   obb[i]  16 doubles = 128 B (one cache line): axis[9] row-major, To[3], extent[3], size
//   rss[i]  16 doubles = 128 B:                  axis[9] row-major, To[3], l[2], r, size
//   first_child[i] int32  (<0: leaf holding triangle -(fc+1))
//   tri[t]  10 doubles = 80 B: p1 p2 p3 (de-indexed vertices), pad
// size = extent.squaredNorm() (OBB::size(), OBB-inl.h:206-209) precomputed at upload.
// ---------------------------------------------------------------------------------------
struct DeviceModel {
  const double* obb;
  const double* rss;
  const int32_t* first_child;
  const double* tri;
  const RssRec32* rss32;  // single-precision steering records (bounds_f32.cuh), 64 B per node
  const ObbRec32* obb32;
  const double2* topo;    // per node {first_child (int32 in the low word of .x), size}: one LDG.128
  int32_t n_nodes, n_tris;
};

#ifndef FCLGPU_PREFETCH
#define FCLGPU_PREFETCH 0
#endif
constexpr int kNodeDoubles = 16;
constexpr int kTriDoubles = 10;
constexpr int kStackCap = 128;  // per-query DFS stack entries; host checks depth1+depth2+2 <= cap

struct NodeRec {
  M3 axis;
  V3 To;
  double e0, e1, e2;  // obb: extent xyz          | rss: l0, l1, r
  double size;
};

__device__ __forceinline__ NodeRec load_node(const double* __restrict__ base, int idx) {
  const double2* p = reinterpret_cast<const double2*>(base + (size_t)idx * kNodeDoubles);
  double2 v0 = __ldg(p + 0), v1 = __ldg(p + 1), v2 = __ldg(p + 2), v3 = __ldg(p + 3);
  double2 v4 = __ldg(p + 4), v5 = __ldg(p + 5), v6 = __ldg(p + 6), v7 = __ldg(p + 7);
  NodeRec n;
  n.axis.m[0] = v0.x; n.axis.m[1] = v0.y; n.axis.m[2] = v1.x; n.axis.m[3] = v1.y;
  n.axis.m[4] = v2.x; n.axis.m[5] = v2.y; n.axis.m[6] = v3.x; n.axis.m[7] = v3.y;
  n.axis.m[8] = v4.x;
  n.To = mk(v4.y, v5.x, v5.y);
  n.e0 = v6.x; n.e1 = v6.y; n.e2 = v7.x;
  n.size = v7.y;
  return n;
}

__device__ __forceinline__ RssRec32 load_rss32(const RssRec32* __restrict__ base, int idx) {
  const float4* p = reinterpret_cast<const float4*>(base + idx);
  const float4 v0 = __ldg(p + 0), v1 = __ldg(p + 1), v2 = __ldg(p + 2), v3 = __ldg(p + 3);
  RssRec32 n;
  n.a[0] = v0.x; n.a[1] = v0.y; n.a[2] = v0.z; n.a[3] = v0.w;
  n.a[4] = v1.x; n.a[5] = v1.y; n.a[6] = v1.z; n.a[7] = v1.w;
  n.a[8] = v2.x; n.c[0] = v2.y; n.c[1] = v2.z; n.c[2] = v2.w;
  n.h0 = v3.x; n.h1 = v3.y; n.r = v3.z; n.s = v3.w;
  return n;
}

__device__ __forceinline__ ObbRec32 load_obb32(const ObbRec32* __restrict__ base, int idx) {
  const float4* p = reinterpret_cast<const float4*>(base + idx);
  const float4 v0 = __ldg(p + 0), v1 = __ldg(p + 1), v2 = __ldg(p + 2), v3 = __ldg(p + 3);
  ObbRec32 n;
  n.a[0] = v0.x; n.a[1] = v0.y; n.a[2] = v0.z; n.a[3] = v0.w;
  n.a[4] = v1.x; n.a[5] = v1.y; n.a[6] = v1.z; n.a[7] = v1.w;
  n.a[8] = v2.x; n.c[0] = v2.y; n.c[1] = v2.z; n.c[2] = v2.w;
  n.e[0] = v3.x; n.e[1] = v3.y; n.e[2] = v3.z; n.s = v3.w;
  return n;
}

// L1 prefetch of the records the next BV round will need (the next entry is known as soon as the
// current one is decided; the lines arrive while the warp does its round bookkeeping)
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_bv32(const DeviceModel& m1, const DeviceModel& m2, uint2 e) {
  prefetch_l1(m1.obb32 + e.x);
  prefetch_l1(m2.obb32 + e.y);
  prefetch_l1(m1.topo + e.x);
  prefetch_l1(m2.topo + e.y);
}

__device__ __forceinline__ void load_topo(const double2* __restrict__ base, int idx, int& first_child, double& size) {
  const double2 t = __ldg(base + idx);
  first_child = __double2loint(t.x);
  size = t.y;
}

__device__ __forceinline__ void load_tri(const double* __restrict__ base, int t, V3 out[3]) {
  const double2* p = reinterpret_cast<const double2*>(base + (size_t)t * kTriDoubles);
  double2 v0 = __ldg(p + 0), v1 = __ldg(p + 1), v2 = __ldg(p + 2), v3 = __ldg(p + 3), v4 = __ldg(p + 4);
  out[0] = mk(v0.x, v0.y, v1.x);
  out[1] = mk(v1.y, v2.x, v2.y);
  out[2] = mk(v3.x, v3.y, v4.x);
}

struct PoseRT {
  M3 R;
  V3 t;
};

// kCoherent: the streamed host path (CollideParams::ready) has the copy engine write the pose buffer WHILE the kernel
// runs; wait_ready's acquire load orders ordinary loads only, and ld.global.nc (__ldg) assumes data that is read-only for
// the kernel's lifetime.  Those kernels therefore read poses with ld.global.cg (coherent at L2, never from L1/texture);
// resident inputs keep the non-coherent path.
__device__ __forceinline__ double2 ld_pose16(const double2* p, bool coherent) { return coherent ? __ldcg(p) : __ldg(p); }
__device__ __forceinline__ PoseRT load_pose(const double* __restrict__ tf, long long i, bool coherent = false) {
  PoseRT p;
  if (tf == nullptr) {
#pragma unroll
    for (int k = 0; k < 9; ++k) p.R.m[k] = (k % 4 == 0) ? 1.0 : 0.0;
    p.t = mk(0, 0, 0);
  } else {
    const double2* q = reinterpret_cast<const double2*>(tf + 12 * i);  // 96 B records, 16-B aligned
    double2 v0 = ld_pose16(q + 0, coherent), v1 = ld_pose16(q + 1, coherent), v2 = ld_pose16(q + 2, coherent);
    double2 v3 = ld_pose16(q + 3, coherent), v4 = ld_pose16(q + 4, coherent), v5 = ld_pose16(q + 5, coherent);
    p.R.m[0] = v0.x; p.R.m[1] = v0.y; p.R.m[2] = v1.x; p.R.m[3] = v1.y; p.R.m[4] = v2.x;
    p.R.m[5] = v2.y; p.R.m[6] = v3.x; p.R.m[7] = v3.y; p.R.m[8] = v4.x;
    p.t = mk(v4.y, v5.x, v5.y);
  }
  return p;
}

// warp-aggregated fetch of the next unprocessed query index (persistent lanes)
__device__ __forceinline__ long long fetch_work(bool need, unsigned long long* counter) {
  const unsigned mask = __ballot_sync(0xffffffffu, need);
  if (mask == 0) return -1;
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(mask) - 1;
  unsigned long long base = 0;
  if (lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(mask));
  base = __shfl_sync(0xffffffffu, base, leader);
  return need ? (long long)(base + __popc(mask & ((1u << lane) - 1u))) : -1;
}

// The same for kernels whose lanes should stay on NEIGHBOURING units of work (a batch sorted for locality): a warp draws
// kWorkBlock consecutive units at a time and hands them to its lanes as they become free, so the 32 queries in flight in
// a warp always come from one stretch of kWorkBlock units.  next / end: the warp's current stretch (uniform across the
// warp).  Returns the unit, -1 for lanes that did not ask, -2 for a lane that has to ask again (stretch used up).
constexpr int kWorkBlock = 64;
__device__ __forceinline__ long long fetch_work_blocked(bool need, unsigned long long* counter, long long& next, long long& end) {
  const unsigned mask = __ballot_sync(0xffffffffu, need);
  if (mask == 0) return -1;
  const int lane = threadIdx.x & 31;
  if (next >= end) {
    const int leader = __ffs(mask) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(counter, (unsigned long long)kWorkBlock);
    base = __shfl_sync(0xffffffffu, base, leader);
    next = (long long)base;
    end = next + kWorkBlock;
  }
  const long long rem = end - next;
  const int rank = __popc(mask & ((1u << lane) - 1u));
  const int cnt = __popc(mask);
  const long long mine = (need && rank < rem) ? next + rank : (need ? -2 : -1);
  next += cnt < rem ? cnt : rem;
  return mine;
}

struct CollideParams {
  DeviceModel m1, m2;
  const double* tf1;
  const double* tf2;
  long long n;
  long long max_contacts;
  int enable_contact;
  int32_t* num_contacts;      // [n]
  fclgpu_contact* scratch;    // [n * stride] or nullptr
  long long stride;           // slots per query in scratch
  uint32_t* n_bv;             // optional
  uint32_t* n_leaf;           // optional
  unsigned long long* work_counter;
  int* status;                // sticky device status (0 ok)
  // streamed input (host API): poses of queries [c << ready_shift, (c + 1) << ready_shift) may be read once
  // ready[c] != 0; the copy stream sets the flag right behind chunk c's copy.  nullptr = everything is resident.
  const unsigned* ready;
  int ready_shift;
  long long ready_q0;  // index of query 0 of this launch in the flagged batch
  // locality order (query_order.cuh): the i-th unit of work is query order[i]; nullptr = batch order.  Results are
  // written at the query's own index, so the order never shows in an output.
  const int32_t* order;
  // collide_front_kernel: stack entries per warp and the head room kept for the depth-first mode
  // (depth1 + depth2 + 2 for the descent, + 32 for the wide round that crossed the limit)
  int front_cap;
  int front_reserve;
  int front_leaf_trigger;  // queued triangle pairs that start a leaf round while fewer than 32 more contacts end the query
};

// Spin until the chunk that holds query q has landed.  Acquire load: the pose loads that follow cannot be
// hoisted above it.  Chunks are multiples of four poses (3 cache lines), so no line is shared between chunks.
// The wait is bounded (kReadyTimeoutNs): if the copies never arrive the kernel reports it instead of hanging.
constexpr unsigned long long kReadyTimeoutNs = 4000000000ull;
__device__ __forceinline__ bool wait_ready(const unsigned* ready, int shift, long long q) {
  if (ready == nullptr || q < 0) return true;
  const unsigned* f = ready + (q >> shift);
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
  if (v) return true;
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (true) {
    __nanosleep(256);
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
    if (v) return true;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > kReadyTimeoutNs) return false;
  }
}

// ---------------------------------------------------------------------------------------
// Variant T (thread per query, persistent lanes): every lane runs the reference's depth-first
// recursion for its own query with an explicit stack, so the visiting order, the early stop
// (canStop) and therefore contact order / truncation are the reference's by construction.
// Lanes that finish pull the next query with one warp-aggregated atomic.
// ---------------------------------------------------------------------------------------
template <bool kStats>
__global__ void __launch_bounds__(128) collide_thread_kernel(CollideParams P) {
  uint2 stk[kStackCap];
  int sp = 0;
  long long q = -1;
  PoseRT tf1;
  M3 R;
  V3 T;
  long long count = 0;
  uint32_t bv_tests = 0, leaf_tests = 0;
  bool exhausted = false;

  while (true) {
    // ---- refill idle lanes ----
    const bool need = (sp == 0) && !exhausted;
    if (need && q >= 0) {  // retire the finished query
      P.num_contacts[q] = (int32_t)count;
      if (kStats) {
        if (P.n_bv) P.n_bv[q] = bv_tests;
        if (P.n_leaf) P.n_leaf[q] = leaf_tests;
      }
      q = -1;
    }
    const long long nq = fetch_work(need, P.work_counter);
    if (nq >= 0 && nq < P.n && !wait_ready(P.ready, P.ready_shift, nq + P.ready_q0)) atomicMin(P.status, (int)FCLGPU_ERR_INPUT_STALLED);
    if (need) {
      if (nq < P.n) {
        q = nq;
        tf1 = load_pose(P.tf1, q, P.ready != nullptr);
        const PoseRT tf2 = load_pose(P.tf2, q, P.ready != nullptr);
        R = mulTM(tf1.R, tf2.R);                    // R1^T R2
        T = mulTv(tf1.R, tf2.t - tf1.t);            // R1^T (t2 - t1)
        count = 0;
        bv_tests = leaf_tests = 0;
        stk[0] = make_uint2(0u, 0u);
        sp = 1;
      } else {
        exhausted = true;
      }
    }
    if (__all_sync(0xffffffffu, exhausted && sp == 0)) break;
    if (sp == 0) continue;

    // ---- one BVTT node ----
    const uint2 e = stk[--sp];
    const int b1 = (int)e.x, b2 = (int)e.y;
    const NodeRec n1 = load_node(P.m1.obb, b1);
    const NodeRec n2 = load_node(P.m2.obb, b2);
    const int fc1 = __ldg(P.m1.first_child + b1);
    const int fc2 = __ldg(P.m2.first_child + b2);
    if (kStats) bv_tests++;
    if (obb_pair_disjoint(R, T, n1.axis, n1.To, mk(n1.e0, n1.e1, n1.e2), n2.axis, n2.To, mk(n2.e0, n2.e1, n2.e2)))
      continue;
    const bool l1 = fc1 < 0, l2 = fc2 < 0;
    if (l1 && l2) {
      if (kStats) leaf_tests++;
      const int id1 = -(fc1 + 1), id2 = -(fc2 + 1);
      V3 Pt[3], Qt[3];
      load_tri(P.m1.tri, id1, Pt);
      load_tri(P.m2.tri, id2, Qt);
#pragma unroll
      for (int k = 0; k < 3; ++k) Qt[k] = mulv(R, Qt[k]) + T;
      if (tri_intersect(Pt[0], Pt[1], Pt[2], Qt[0], Qt[1], Qt[2])) {
        if (!P.enable_contact) {
          if (count < P.max_contacts) {
            if (P.scratch) {
              if (count < P.stride) {
                fclgpu_contact* c = P.scratch + q * P.stride + count;
                c->b1 = id1;
                c->b2 = id2;
              } else {
                atomicMin(P.status, (int)FCLGPU_ERR_CONTACT_OVERFLOW);
              }
            }
            count++;
          }
        } else {
          V3 cp[2], nrm;
          unsigned nc;
          double depth;
          tri_contact_info(Pt, Qt, cp, nc, depth, nrm);
          if (P.max_contacts < count + (long long)nc)
            nc = (P.max_contacts > count) ? (unsigned)(P.max_contacts - count) : 0u;
          for (unsigned k = 0; k < nc; ++k) {
            if (P.scratch) {
              if (count < P.stride) {
                fclgpu_contact* c = P.scratch + q * P.stride + count;
                const V3 pw = mulv(tf1.R, cp[k]) + tf1.t;  // tf1 * p
                const V3 nw = mulv(tf1.R, nrm);            // tf1.linear() * n
                c->b1 = id1;
                c->b2 = id2;
                c->normal[0] = nw.x; c->normal[1] = nw.y; c->normal[2] = nw.z;
                c->pos[0] = pw.x; c->pos[1] = pw.y; c->pos[2] = pw.z;
                c->penetration_depth = depth;
              } else {
                atomicMin(P.status, (int)FCLGPU_ERR_CONTACT_OVERFLOW);
              }
            }
            count++;
          }
        }
        // canStop(): isCollision && num_max_contacts <= numContacts -> every pending sibling is skipped
        if (count > 0 && P.max_contacts <= count) sp = 0;
      }
      continue;
    }
    // firstOverSecond: descend model1 iff l2 || (!l1 && size1 > size2)
    uint2 left, right;
    if (l2 || (!l1 && (n1.size > n2.size))) {
      left = make_uint2((unsigned)fc1, (unsigned)b2);
      right = make_uint2((unsigned)fc1 + 1u, (unsigned)b2);
    } else {
      left = make_uint2((unsigned)b1, (unsigned)fc2);
      right = make_uint2((unsigned)b1, (unsigned)fc2 + 1u);
    }
    if (sp + 2 > kStackCap) {
      atomicMin(P.status, (int)FCLGPU_ERR_STACK_OVERFLOW);
      sp = 0;
      continue;
    }
    stk[sp++] = right;
    stk[sp++] = left;
  }
}

struct DistanceParams {
  DeviceModel m1, m2;
  const double* tf1;
  const double* tf2;
  long long n;
  int enable_nearest_points;
  double* min_distance;
  double* p1;
  double* p2;
  int32_t* b1;
  int32_t* b2;
  uint32_t* n_bv;
  uint32_t* n_leaf;
  unsigned long long* work_counter;
  int* status;
  // overflow area of the sorted-front kernel (deep trees): spill_cap entries per warp, or nullptr
  uint2* spill_pair;
  float* spill_bound;
  int spill_cap;
  int spill_warps;  // warps the area was sized for
  // value the minimum starts from (DistanceResult's initial min_distance): DBL_MAX for fcl::distance; a finite cutoff
  // for the tolerance-verification extension -- everything whose bound is >= cutoff is pruned from the first round on
  double cutoff;
  // tolerance verdicts: the traversal of a query ends as soon as its minimum is <= stop_below (a triangle pair within
  // the tolerance has been found: the verdict is decided); -1 = never (distances are >= 0).  within[q] = min <= stop_below.
  double stop_below;
  uint8_t* within;
};

struct DistState {
  double min_d;
  V3 p1, p2;
  int b1, b2;
};

__device__ __forceinline__ void dist_leaf(const DeviceModel& m1, const DeviceModel& m2, const M3& R, const V3& T,
                                          int id1, int id2, DistState& s) {
  V3 S[3], Tt[3];
  load_tri(m1.tri, id1, S);
  load_tri(m2.tri, id2, Tt);
#pragma unroll
  for (int k = 0; k < 3; ++k) Tt[k] = mulv(R, Tt[k]) + T;  // tf * p
  V3 Pn, Qn;
  const double d = tri_distance(S, Tt, Pn, Qn);
  if (s.min_d > d) {  // DistanceResult::update keeps strictly smaller (distance_result-inl.h:66-103)
    s.min_d = d;
    s.b1 = id1;
    s.b2 = id2;
    s.p1 = Pn;
    s.p2 = Qn;
  }
}

template <bool kStats>
__global__ void __launch_bounds__(128) distance_thread_kernel(DistanceParams P) {
  uint2 stk[kStackCap];
  double stk_d[kStackCap];
  int sp = 0;
  long long q = -1;
  PoseRT tf1;
  M3 R;
  V3 T;
  DistState s;
  uint32_t bv_tests = 0, leaf_tests = 0;
  bool exhausted = false;

  while (true) {
    const bool need = (sp == 0) && !exhausted;
    if (need && q >= 0) {
      // postprocess: nearest points (model1 frame) -> world with tf1
      if (P.min_distance) P.min_distance[q] = s.min_d;
      if (P.within) P.within[q] = s.min_d <= P.stop_below ? 1 : 0;
      if (P.b1) P.b1[q] = s.b1;
      if (P.b2) P.b2[q] = s.b2;
      if (P.enable_nearest_points) {
        const V3 w1 = mulv(tf1.R, s.p1) + tf1.t, w2 = mulv(tf1.R, s.p2) + tf1.t;
        if (P.p1) { P.p1[3 * q] = w1.x; P.p1[3 * q + 1] = w1.y; P.p1[3 * q + 2] = w1.z; }
        if (P.p2) { P.p2[3 * q] = w2.x; P.p2[3 * q + 1] = w2.y; P.p2[3 * q + 2] = w2.z; }
      }
      if (kStats) {
        if (P.n_bv) P.n_bv[q] = bv_tests;
        if (P.n_leaf) P.n_leaf[q] = leaf_tests;
      }
      q = -1;
    }
    const long long nq = fetch_work(need, P.work_counter);
    if (need) {
      if (nq < P.n) {
        q = nq;
        tf1 = load_pose(P.tf1, q);
        const PoseRT tf2 = load_pose(P.tf2, q);
        // tf = tf1.inverse(Isometry) * tf2: linear R1^T R2, translation R1^T t2 + (-(R1^T t1))
        R = mulTM(tf1.R, tf2.R);
        const V3 it = mulTv(tf1.R, tf1.t);
        T = mulTv(tf1.R, tf2.t) + mk(-it.x, -it.y, -it.z);
        s.min_d = P.cutoff;
        s.b1 = s.b2 = -1;
        s.p1 = s.p2 = mk(0, 0, 0);
        bv_tests = leaf_tests = 0;
        dist_leaf(P.m1, P.m2, R, T, 0, 0, s);  // preprocess: seed with triangle 0 / triangle 0
        stk[0] = make_uint2(0u, 0u);
        stk_d[0] = -1.0;  // the root pair is never bound-tested
        sp = (s.min_d <= P.stop_below) ? 0 : 1;
      } else {
        exhausted = true;
      }
    }
    if (__all_sync(0xffffffffu, exhausted && sp == 0)) break;
    if (sp == 0) continue;

    --sp;
    const uint2 e = stk[sp];
    const double bound = stk_d[sp];
    if (bound >= s.min_d) continue;  // canStop(c) with rel_err = abs_err = 0
    const int b1 = (int)e.x, b2 = (int)e.y;
    const int fc1 = __ldg(P.m1.first_child + b1);
    const int fc2 = __ldg(P.m2.first_child + b2);
    const bool l1 = fc1 < 0, l2 = fc2 < 0;
    if (l1 && l2) {
      if (kStats) leaf_tests++;
      dist_leaf(P.m1, P.m2, R, T, -(fc1 + 1), -(fc2 + 1), s);
      if (s.min_d <= P.stop_below) sp = 0;  // tolerance verdict decided
      continue;
    }
    // firstOverSecond needs both sizes (OBB extents' squared norm, stored in the rss record too)
    const double size1 = __ldg(P.m1.rss + (size_t)b1 * kNodeDoubles + 15);
    const double size2 = __ldg(P.m2.rss + (size_t)b2 * kNodeDoubles + 15);
    int a1, a2, c1, c2;
    if (l2 || (!l1 && (size1 > size2))) {
      a1 = fc1; a2 = b2; c1 = fc1 + 1; c2 = b2;
    } else {
      a1 = b1; a2 = fc2; c1 = b1; c2 = fc2 + 1;
    }
    double d1, d2;
    {
      const NodeRec na1 = load_node(P.m1.rss, a1);
      const NodeRec na2 = load_node(P.m2.rss, a2);
      const double la[2] = {na1.e0, na1.e1}, lb[2] = {na2.e0, na2.e1};
      d1 = rss_pair_distance(R, T, na1.axis, na1.To, la, na1.e2, na2.axis, na2.To, lb, na2.e2);
    }
    {
      const NodeRec nc1 = load_node(P.m1.rss, c1);
      const NodeRec nc2 = load_node(P.m2.rss, c2);
      const double la[2] = {nc1.e0, nc1.e1}, lb[2] = {nc2.e0, nc2.e1};
      d2 = rss_pair_distance(R, T, nc1.axis, nc1.To, la, nc1.e2, nc2.axis, nc2.To, lb, nc2.e2);
    }
    if (kStats) bv_tests += 2;
    if (sp + 2 > kStackCap) {
      atomicMin(P.status, (int)FCLGPU_ERR_STACK_OVERFLOW);
      sp = 0;
      continue;
    }
    // visit the nearer child first: it goes on top of the stack
    if (d2 < d1) {
      stk[sp] = make_uint2((unsigned)a1, (unsigned)a2); stk_d[sp] = d1; ++sp;
      stk[sp] = make_uint2((unsigned)c1, (unsigned)c2); stk_d[sp] = d2; ++sp;
    } else {
      stk[sp] = make_uint2((unsigned)c1, (unsigned)c2); stk_d[sp] = d2; ++sp;
      stk[sp] = make_uint2((unsigned)a1, (unsigned)a2); stk_d[sp] = d1; ++sp;
    }
  }
}

}  // namespace fclgpu

namespace fclgpu {

// ---------------------------------------------------------------------------------------
// Variant W for distance (warp per query, sorted front):
//   * the warp owns one query; its BVTT front lives in a shared-memory stack of
//     (b1, b2, lower bound) entries ordered so that the nearest candidates are on top;
//   * BV round: the top kDistPop entries are popped, entries whose bound no longer beats the
//     current minimum are dropped (canStop), leaf pairs go to a shared leaf queue, the others
//     are expanded: lane 2e / 2e+1 evaluate the RSS distance of the two children of entry e
//     (firstOverSecond picks the side to split), so all 32 lanes run rect_distance together;
//     surviving children are sorted by bound with a warp bitonic network and pushed nearest-
//     on-top;
//   * leaf round (queue >= kLeafTrigger or front empty): up to 32 triangle pairs are tested
//     in parallel (triDistance), warp arg-min updates the query's minimum.
// The minimum distance is the minimum over every leaf pair not excluded by a valid lower
// bound, i.e. exactly the reference's value; which pair is reported first among exact ties
// may differ from the sequential recursion, so ids / points are compared by value.
// Measured against the sequential recursion on env/rob this order needs 0.7x the BV tests and
// 0.5x the leaf tests (nearest-first expansion tightens the bound sooner).
// ---------------------------------------------------------------------------------------

#ifndef FCLGPU_DIST_OUTOFLINE
#define FCLGPU_DIST_OUTOFLINE 0
#endif
// exact triDistance out of line: the hot BV loop of the sorted-front kernel then needs far fewer
// registers than the FP64 leaf routine
__device__ __noinline__ double tri_distance_outofline(const V3* Sv, const V3* Tv, V3* Pn, V3* Qn) {
  return tri_distance(Sv, Tv, *Pn, *Qn);
}

#ifndef FCLGPU_DIST_POP
#define FCLGPU_DIST_POP 16
#endif
constexpr int kDistPop = FCLGPU_DIST_POP;  // entries expanded per BV round (2 lanes each)
constexpr int kDistStackCap = 512;   // entries per warp
constexpr int kSpillBlock = 256;     // entries moved to / from the global overflow area at a time
constexpr int kLeafCap = 64;
constexpr int kLeafTrigger = 32;
constexpr int kDistWarps = 4;        // warps per block
// Triangle-level FP32 screening of leaf pairs before the exact triDistance: a conservative lower bound on the
// triangle distance from the support-function gap along a few directions (tri_lower_bound_dirs_f32); only pairs
// whose bound still beats the minimum go on to the exact queue.  Measured on B200 (1M poses, env/rob), direction
// set -> kernel ms / exact tests per query: none 41.4 / 147; face normals 41.6 / 133; normals + centroid 39.2 / 105;
// + nine edge x edge 41.4 / 70; in-plane edge normals only 37.4 / 78; normals + in-plane 37.3 / 71 (default);
// all 19 directions 46.5 / 59 (instruction fetch).
#ifndef FCLGPU_SCREEN_LEAVES
#define FCLGPU_SCREEN_LEAVES 9
#endif
constexpr bool kScreenLeaves = FCLGPU_SCREEN_LEAVES != 0;  // direction set of the screening bound (bit mask, see tri_lower_bound_dirs_f32)

struct __align__(16) WarpFront {
  float bound[kDistStackCap];   // lower bounds are stored in single precision, rounded down
  uint2 pair[kDistStackCap];
  float leaf_bound[kLeafCap];
  uint2 leaf_pair[kLeafCap];  // triangle ids: pairs waiting for the exact triDistance
  float raw_bound[kLeafCap];
  uint2 raw_pair[kLeafCap];   // triangle ids: leaf pairs not yet screened by the triangle-level bound
  uint2 expand[32];           // child pairs handed from the entry's holder lane to the two testing lanes
  double best[6];
  int best_id[2];
};

__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int m) {
  unsigned lo = (unsigned)v, hi = (unsigned)(v >> 32);
  lo = __shfl_xor_sync(0xffffffffu, lo, m);
  hi = __shfl_xor_sync(0xffffffffu, hi, m);
  return ((unsigned long long)hi << 32) | lo;
}
__device__ __forceinline__ unsigned long long shfl_u64(unsigned long long v, int src) {
  unsigned lo = (unsigned)v, hi = (unsigned)(v >> 32);
  lo = __shfl_sync(0xffffffffu, lo, src);
  hi = __shfl_sync(0xffffffffu, hi, src);
  return ((unsigned long long)hi << 32) | lo;
}

// Ascending bitonic sort of one 32-bit key per lane.  The keys only steer the visiting order
// (nearest candidates first), so a rounded-down float image of the bound with the lane id in
// the low 5 bits is enough: keys are unique and the winner's payload is fetched afterwards.
__device__ __forceinline__ unsigned warp_sort_keys(unsigned key, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const unsigned other = __shfl_xor_sync(0xffffffffu, key, j);
      const bool keep_min = (((lane & j) == 0) == ((lane & k) == 0));
      key = keep_min ? min(key, other) : max(key, other);
    }
  }
  return key;
}

#ifndef FCLGPU_DIST_MINBLOCKS
#define FCLGPU_DIST_MINBLOCKS 5
#endif
#ifndef FCLGPU_DIST_SEED
#define FCLGPU_DIST_SEED 5
#endif
// Pre-expanded front entries (FCLGPU_DIST_PREX): the lane that bounds a child pair also reads the two topology records
// and stores the entry in the form the NEXT expansion needs -- a leaf pair as its two triangle ids, an internal pair as
// (first child of the node firstOverSecond splits, which model that is, the other node) -- so that a popped entry is
// expanded without a dependent global load (pop -> topology -> ballot was the longest chain of a BV round); the loads
// now overlap the key sort.  Encoding: leaf pair {t1 | 1<<31, t2}; internal {fc | side<<30, other}, side 1 = model 1 splits.
#ifndef FCLGPU_DIST_PREX
#define FCLGPU_DIST_PREX 1
#endif
// Development build (-DFCLGPU_DIST_PROF=1): per-phase SM cycles of the sorted-front kernel, summed over warps
// (0 prologue / epilogue, 1 BV rounds, 2 screening rounds, 3 exact rounds, 4 refill; 8.. = round counts), read with
// fclgpu_debug_counters().  Compiled out of the product build.
#ifndef FCLGPU_DIST_PROF
#define FCLGPU_DIST_PROF 0
#endif
__device__ unsigned long long g_debug_counters[16];
#if FCLGPU_DIST_PROF
#define DPROF_MARK(k)                      \
  {                                        \
    const long long t_now = clock64();     \
    prof_acc[k] += t_now - prof_t0;        \
    prof_cnt[k] += 1;                      \
    prof_t0 = t_now;                       \
  }
#else
#define DPROF_MARK(k)
#endif
// front entry of the node pair (x, y) in the pre-expanded form (FCLGPU_DIST_PREX)
__device__ __forceinline__ uint2 prex_entry(unsigned x, unsigned y, int fc1, double size1, int fc2, double size2) {
  const bool l1 = fc1 < 0, l2 = fc2 < 0;
  if (l1 && l2) return make_uint2((unsigned)(-(fc1 + 1)) | 0x80000000u, (unsigned)(-(fc2 + 1)));
  if (l2 || (!l1 && (size1 > size2))) return make_uint2((unsigned)fc1 | 0x40000000u, y);  // firstOverSecond: model 1 splits
  return make_uint2((unsigned)fc2, x);
}
// kSpill: instantiation with the global overflow area for deep trees (kept out of the default instantiation:
// the extra live state costs the hot loop 10 %)
// kTol: tolerance verdicts (stop_below / within) -- a separate instantiation as well: the early-exit test inside the leaf
// round cost the plain query 15 % (34.9 -> 40.3 ms per 1M poses: register allocation of the hot loop)
template <bool kStats, bool kBound32, bool kSpill = false, bool kTol = false>
__global__ void __launch_bounds__(kDistWarps * 32, FCLGPU_DIST_MINBLOCKS) distance_warp_kernel(DistanceParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WarpFront& S = reinterpret_cast<WarpFront*>(smem_raw)[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  // the screening round pays off when triangle tests dominate; on BVHs beyond the caches (the kSpill instantiation)
  // box tests dominate and the later minimum update costs more than the saved tests (cfg5: 48 -> 56 ms)
  constexpr bool kScreen = kScreenLeaves && !kSpill;

  // Seed front (FCLGPU_DIST_SEED levels): while nothing is known about the minimum every pair of the first BVTT levels
  // is expanded whatever the pose, and which node of a pair is split depends on the two trees only (firstOverSecond).
  // Warp 0 therefore expands the root pair level by level once per block (leaf pairs stay as they are, at most 32
  // entries), and every query starts with ONE full round that bounds and sorts those pairs instead of the 1-, 2-, 4-,
  // 8- and 16-lane rounds that lead there.  Same BVTT coverage, so the same minimum.
  constexpr bool kSeed = FCLGPU_DIST_SEED > 0;
  constexpr bool kPrex = FCLGPU_DIST_PREX != 0;
  __shared__ uint2 s_seed[32];
  __shared__ int s_nseed;
  __shared__ uint2 s_root;  // the root pair as a front entry
  if (threadIdx.x == 0) {
    s_root = make_uint2(0u, 0u);
    if (kPrex) {
      int fc1, fc2;
      double size1, size2;
      load_topo(P.m1.topo, 0, fc1, size1);
      load_topo(P.m2.topo, 0, fc2, size2);
      s_root = prex_entry(0u, 0u, fc1, size1, fc2, size2);
    }
  }
  if (!kSeed) __syncthreads();
  if (kSeed) {
    if (threadIdx.x < 32) {
      if (lane == 0) s_seed[0] = make_uint2(0u, 0u);
      __syncwarp();
      int n = 1;
      for (int lv = 0; lv < FCLGPU_DIST_SEED; ++lv) {
        const bool have = lane < n;
        uint2 e = make_uint2(0u, 0u);
        int fc1 = -1, fc2 = -1;
        double size1 = 0.0, size2 = 0.0;
        if (have) {
          e = s_seed[lane];
          load_topo(P.m1.topo, (int)e.x, fc1, size1);
          load_topo(P.m2.topo, (int)e.y, fc2, size2);
        }
        const bool l1 = fc1 < 0, l2 = fc2 < 0;
        const bool exp = have && !(l1 && l2);
        const unsigned em = __ballot_sync(0xffffffffu, exp);
        const int add = __popc(em);
        if (add == 0 || n + add > 32) break;
        __syncwarp();
        if (have) {
          const int pos = lane + __popc(em & lt_mask);
          if (!exp) {
            s_seed[pos] = e;
          } else if (l2 || (!l1 && (size1 > size2))) {  // firstOverSecond
            s_seed[pos] = make_uint2((unsigned)fc1, e.y);
            s_seed[pos + 1] = make_uint2((unsigned)fc1 + 1u, e.y);
          } else {
            s_seed[pos] = make_uint2(e.x, (unsigned)fc2);
            s_seed[pos + 1] = make_uint2(e.x, (unsigned)fc2 + 1u);
          }
        }
        n += add;
        __syncwarp();
      }
      if (lane == 0) s_nseed = n;
    }
    __syncthreads();
  }

  while (true) {
    long long q = 0;
    if (lane == 0) q = (long long)atomicAdd(P.work_counter, 1ull);
    q = __shfl_sync(0xffffffffu, q, 0);
    if (q >= P.n) break;

#if FCLGPU_DIST_PROF
    long long prof_t0 = clock64();
#endif
    M3 R;
    V3 T;
    {
      const PoseRT tf1 = load_pose(P.tf1, q);
      const PoseRT tf2 = load_pose(P.tf2, q);
      R = mulTM(tf1.R, tf2.R);
      const V3 it = mulTv(tf1.R, tf1.t);
      T = mulTv(tf1.R, tf2.t) + mk(-it.x, -it.y, -it.z);
    }
    float Rf[9], Tf[3], t_l1 = 0.0f;
    if (kBound32) {
#pragma unroll
      for (int k = 0; k < 9; ++k) Rf[k] = (float)R.m[k];
      Tf[0] = (float)T.x; Tf[1] = (float)T.y; Tf[2] = (float)T.z;
      t_l1 = __double2float_ru((fabs(T.x) + fabs(T.y)) + fabs(T.z));
    }

#if FCLGPU_DIST_PROF
    long long prof_acc[5] = {0, 0, 0, 0, 0};
    unsigned prof_cnt[5] = {0, 0, 0, 0, 0};
    DPROF_MARK(0)
#endif
    double min_d = P.cutoff;  // DBL_MAX for fcl::distance
    // bounds are floats: (double)b < min_d  <=>  b < min_f with min_f = min_d rounded up (the smallest float >= min_d;
    // +inf for DBL_MAX)
    float min_f = __double2float_ru(min_d);
    // (with a finite start value far queries end in their first round either way: root -> two children, or the seed pairs)
#ifndef FCLGPU_DIST_SEED_CUTOFF
#define FCLGPU_DIST_SEED_CUTOFF 1
#endif
    bool seed_pending = kSeed && (FCLGPU_DIST_SEED_CUTOFF || !(min_d < DBL_MAX));
    int sp = seed_pending ? 0 : 1, nleaf = 1, nraw = 0;
    bool had_exact = false;
    uint32_t bv_tests = 0, leaf_tests = 0;
    if (lane == 0) {
      S.pair[0] = s_root;
      S.bound[0] = -1.0f;                   // the root pair is never bound-tested
      S.leaf_pair[0] = make_uint2(0u, 0u);  // preprocess: triangle 0 / triangle 0 seeds the result
      S.leaf_bound[0] = -1.0f;
      S.best_id[0] = S.best_id[1] = -1;
#pragma unroll
      for (int k = 0; k < 6; ++k) S.best[k] = 0.0;
    }
    __syncwarp();

    int gsp = 0;  // entries parked in the warp's global overflow area
    const int gwarp = blockIdx.x * kDistWarps + (threadIdx.x >> 5);
    const bool can_spill = kSpill && P.spill_pair != nullptr && gwarp < P.spill_warps;
    uint2* const g_pair = can_spill ? P.spill_pair + (size_t)gwarp * P.spill_cap : nullptr;
    float* const g_bound = can_spill ? P.spill_bound + (size_t)gwarp * P.spill_cap : nullptr;

    while (true) {
      if (kSpill && sp == 0 && gsp > 0) {
        // refill: bring the most recently parked block back (oldest first, so the stack order is kept),
        // dropping what the minimum found meanwhile has made irrelevant
        const int take = gsp < kSpillBlock ? gsp : kSpillBlock;
        gsp -= take;
        for (int base = 0; base < take; base += 32) {
          const int i = base + lane;
          uint2 pr = make_uint2(0u, 0u);
          float bd = 0.0f;
          bool live = false;
          if (i < take) {
            pr = g_pair[gsp + i];
            bd = g_bound[gsp + i];
            live = bd < min_f;
          }
          const unsigned m = __ballot_sync(0xffffffffu, live);
          if (live) {
            const int pos = sp + __popc(m & lt_mask);
            S.pair[pos] = pr;
            S.bound[pos] = bd;
          }
          sp += __popc(m);
        }
        __syncwarp();
        DPROF_MARK(4)
        continue;
      }
      // the screening round may add up to min(nraw, 32) pairs to the exact queue: only run it when they fit
      // (otherwise the exact queue is at least half full and the exact round below drains it first)
      const bool can_screen = kBound32 && kScreen && nraw > 0 && (nleaf + (nraw < 32 ? nraw : 32) <= kLeafCap);
      // Early first minimum (FCLGPU_DIST_EAGER > 0): until an exact round has run nothing is pruned, so the first one is
      // not held back until 32 pairs are queued -- a screening round starts at FCLGPU_DIST_EAGER raw pairs and the exact round
      // follows at once.
#ifndef FCLGPU_DIST_EAGER
#define FCLGPU_DIST_EAGER 16
#endif
      // (only with the screening round, i.e. cache-resident models: on cfg5 the early round costs 4 %)
      const bool eager = FCLGPU_DIST_EAGER > 0 && kBound32 && kScreen && !had_exact;
      if (can_screen && (nraw >= 32 || sp == 0 || (eager && nraw >= FCLGPU_DIST_EAGER))) {
        // ---- screening round: triangle-level lower bound (FP32, branch-free) on up to 32 raw leaf pairs;
        // only pairs that can still beat the minimum go on to the exact queue
        const int k = nraw < 32 ? nraw : 32;
        nraw -= k;
        uint2 ids = make_uint2(0u, 0u);
        float b = 0.0f;
        bool mine = lane < k;
        if (mine) {
          ids = S.raw_pair[nraw + lane];
          b = S.raw_bound[nraw + lane];
          mine = b < min_f;
        }
        bool keep = false;
        if (mine) {
          V3 Sv[3], Tv[3];
          load_tri(P.m1.tri, (int)ids.x, Sv);
          load_tri(P.m2.tri, (int)ids.y, Tv);
          float s1[3], s2[3], t0[3], t1[3], t2[3];
          {
            const V3 a = Sv[1] - Sv[0], c = Sv[2] - Sv[0];
            s1[0] = (float)a.x; s1[1] = (float)a.y; s1[2] = (float)a.z;
            s2[0] = (float)c.x; s2[1] = (float)c.y; s2[2] = (float)c.z;
            const V3 u0 = (mulv(R, Tv[0]) + T) - Sv[0], u1 = (mulv(R, Tv[1]) + T) - Sv[0], u2 = (mulv(R, Tv[2]) + T) - Sv[0];
            t0[0] = (float)u0.x; t0[1] = (float)u0.y; t0[2] = (float)u0.z;
            t1[0] = (float)u1.x; t1[1] = (float)u1.y; t1[2] = (float)u1.z;
            t2[0] = (float)u2.x; t2[1] = (float)u2.y; t2[2] = (float)u2.z;
          }
          const float lb = tri_lower_bound_dirs_f32<(FCLGPU_SCREEN_LEAVES ? FCLGPU_SCREEN_LEAVES : 15)>(s1, s2, t0, t1, t2);
          b = fmaxf(b, lb);
          keep = b < min_f;
        }
        const unsigned km = __ballot_sync(0xffffffffu, keep);
        if (keep) {
          const int pos = nleaf + __popc(km & lt_mask);
          S.leaf_pair[pos] = ids;
          S.leaf_bound[pos] = b;
        }
        nleaf += __popc(km);
        __syncwarp();
        DPROF_MARK(2)
        continue;
      }
      const bool do_leaf = (nleaf >= kLeafTrigger) || (sp == 0 && nleaf > 0 && !seed_pending) || (eager && nleaf > 1);
      if (do_leaf) {
        had_exact = true;
        const int k = nleaf < 32 ? nleaf : 32;
        nleaf -= k;
        double d = 1.7976931348623157e308;
        V3 Pn = mk(0, 0, 0), Qn = mk(0, 0, 0);
        uint2 ids = make_uint2(0u, 0u);
        bool mine = lane < k;
        if (mine) {
          ids = S.leaf_pair[nleaf + lane];
          mine = S.leaf_bound[nleaf + lane] < min_f;
        }
        if (kStats) leaf_tests += __popc(__ballot_sync(0xffffffffu, mine));
        if (mine) {
          V3 Sv[3], Tv[3];
          load_tri(P.m1.tri, (int)ids.x, Sv);
          load_tri(P.m2.tri, (int)ids.y, Tv);
#pragma unroll
          for (int c = 0; c < 3; ++c) Tv[c] = mulv(R, Tv[c]) + T;
          d = (FCLGPU_DIST_OUTOFLINE && kBound32) ? tri_distance_outofline(Sv, Tv, &Pn, &Qn) : tri_distance(Sv, Tv, Pn, Qn);
        }
        // warp arg-min (distances are >= 0, so the bit pattern orders like the value); ties -> lowest lane
        unsigned long long key = (unsigned long long)__double_as_longlong(d);
        int who = lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const unsigned long long ok = shfl_xor_u64(key, o);
          const int ow = __shfl_xor_sync(0xffffffffu, who, o);
          if (ok < key || (ok == key && ow < who)) {
            key = ok;
            who = ow;
          }
        }
        const double dmin = __longlong_as_double((long long)key);
        if (dmin < min_d) {  // strictly smaller, like DistanceResult::update
          min_d = dmin;
          min_f = __double2float_ru(dmin);
          if (lane == who) {
            S.best[0] = Pn.x; S.best[1] = Pn.y; S.best[2] = Pn.z;
            S.best[3] = Qn.x; S.best[4] = Qn.y; S.best[5] = Qn.z;
            S.best_id[0] = (int)ids.x;
            S.best_id[1] = (int)ids.y;
          }
          if (kTol && dmin <= P.stop_below) {  // tolerance verdict decided: a pair within the tolerance exists
            sp = 0;
            gsp = 0;
            nleaf = 0;
            nraw = 0;
          }
        }
        __syncwarp();
        DPROF_MARK(3)
        continue;
      }
      if (sp == 0 && !seed_pending) break;

      // ---- BV round ----
      // every lane pops one entry; dead entries (bound no longer beats the minimum) vanish, leaf
      // pairs move to the leaf queue, the first kDistPop internal entries (nearest first) are
      // expanded by two lanes each and the remaining internal entries go back on the stack.
      // Close to the stack limit (a front that nothing prunes, e.g. coincident meshes before the first zero distance
      // is found) a warp without (room in) an overflow area pops and expands ONE entry per round: plain
      // nearest-first depth first, whose growth is bounded by the tree depths (checked on the host).
      int n_test;  // lanes that bound a child pair in this round
      if (kSeed && seed_pending) {
        seed_pending = false;
        n_test = s_nseed;
        if (lane < n_test) S.expand[lane] = s_seed[lane];
        __syncwarp();
      } else {
#ifndef FCLGPU_DIST_TIGHT
#define FCLGPU_DIST_TIGHT 1
#endif
        const bool tight = FCLGPU_DIST_TIGHT && sp > kDistStackCap - 160 &&
                           (!kSpill || g_pair == nullptr || gsp + kSpillBlock > P.spill_cap);  // ... or the overflow area is full
        const int k = tight ? 1 : (sp < 32 ? sp : 32);
        uint2 pr = make_uint2(0u, 0u);
        float bd = 0.0f;
        bool alive = lane < k;
        if (alive) {
          pr = S.pair[sp - 1 - lane];
          bd = S.bound[sp - 1 - lane];
          alive = bd < min_f;  // canStop(c): bound >= min_distance -> skip
        }
        int fc1 = 0, fc2 = 0;
        double size1 = 0.0, size2 = 0.0;
        if (!kPrex && alive) {  // {first_child, size} of both nodes: one 16-byte load each
          load_topo(P.m1.topo, (int)pr.x, fc1, size1);
          load_topo(P.m2.topo, (int)pr.y, fc2, size2);
        }
        const bool l1 = fc1 < 0, l2 = fc2 < 0;
        const bool leafpair = alive && (kPrex ? (pr.x >> 31) != 0u : (l1 && l2));
        const unsigned lm = __ballot_sync(0xffffffffu, leafpair);
        if (leafpair) {
          const uint2 tri_ids = kPrex ? make_uint2(pr.x & 0x7fffffffu, pr.y)
                                      : make_uint2((unsigned)(-(fc1 + 1)), (unsigned)(-(fc2 + 1)));
          if (kBound32 && kScreen) {
            const int pos = nraw + __popc(lm & lt_mask);
            S.raw_pair[pos] = tri_ids;
            S.raw_bound[pos] = bd;
          } else {
            const int pos = nleaf + __popc(lm & lt_mask);
            S.leaf_pair[pos] = tri_ids;
            S.leaf_bound[pos] = bd;
          }
        }
        if (kBound32 && kScreen) nraw += __popc(lm);
        else nleaf += __popc(lm);
        const bool internal = alive && !leafpair;
        const unsigned im = __ballot_sync(0xffffffffu, internal);
        const int n_int = __popc(im), rank = __popc(im & lt_mask);
        sp -= k;
        if (kSpill && kDistStackCap - sp - n_int < kDistPop && g_pair != nullptr && sp >= kSpillBlock && gsp + kSpillBlock <= P.spill_cap) {
          // The front no longer fits (deep trees): park the bottom of the stack -- the farthest, oldest candidates --
          // in the warp's global overflow area and slide the rest down.
          for (int base = 0; base < kSpillBlock; base += 32) {
            g_pair[gsp + base + lane] = S.pair[base + lane];
            g_bound[gsp + base + lane] = S.bound[base + lane];
          }
          gsp += kSpillBlock;
          __syncwarp();
          for (int base = kSpillBlock; base < sp; base += 32) {
            const int i = base + lane;
            uint2 pr2 = make_uint2(0u, 0u);
            float bd2 = 0.0f;
            if (i < sp) {
              pr2 = S.pair[i];
              bd2 = S.bound[i];
            }
            __syncwarp();
            if (i < sp) {
              S.pair[i - kSpillBlock] = pr2;
              S.bound[i - kSpillBlock] = bd2;
            }
          }
          sp -= kSpillBlock;
          __syncwarp();
        }
        int n_exp = n_int < kDistPop ? n_int : kDistPop;
        const int room = kDistStackCap - sp - n_int;  // after re-pushing the leftovers, 2*n_exp children minus n_exp must fit
        if (n_exp > room) n_exp = room;
        if (n_int > 0 && n_exp <= 0) {
          if (lane == 0) atomicMin(P.status, (int)FCLGPU_ERR_STACK_OVERFLOW);
          sp = 0;
          gsp = 0;
          nleaf = 0;
          nraw = 0;
          break;
        }
        __syncwarp();  // every lane has read its popped entry before slots are overwritten
        if (internal) {
          if (rank < n_exp) {
            if (kPrex) {
              const unsigned fc = pr.x & 0x3fffffffu;
              if (pr.x & 0x40000000u) {  // model 1's node splits
                S.expand[2 * rank] = make_uint2(fc, pr.y);
                S.expand[2 * rank + 1] = make_uint2(fc + 1u, pr.y);
              } else {
                S.expand[2 * rank] = make_uint2(pr.y, fc);
                S.expand[2 * rank + 1] = make_uint2(pr.y, fc + 1u);
              }
            } else if (l2 || (!l1 && (size1 > size2))) {  // firstOverSecond
              S.expand[2 * rank] = make_uint2((unsigned)fc1, pr.y);
              S.expand[2 * rank + 1] = make_uint2((unsigned)fc1 + 1u, pr.y);
            } else {
              S.expand[2 * rank] = make_uint2(pr.x, (unsigned)fc2);
              S.expand[2 * rank + 1] = make_uint2(pr.x, (unsigned)fc2 + 1u);
            }
          } else {  // leftover: back on the stack, order preserved (rank n_exp nearest -> top)
            const int pos = sp + (n_int - 1 - rank);
            S.pair[pos] = pr;
            S.bound[pos] = bd;
          }
        }
        sp += n_int - n_exp;
        __syncwarp();
        n_test = 2 * n_exp;
      }
      const bool expand = lane < n_test;
      unsigned key = 0xffffffffu;
      uint2 xy = make_uint2(0u, 0u);
      float d = 0.0f;  // lower bound on the distance between the two child BVs
      if (expand) {
        xy = S.expand[lane];
        if (kBound32) {
          const RssRec32 n1 = load_rss32(P.m1.rss32, (int)xy.x);
          const RssRec32 n2 = load_rss32(P.m2.rss32, (int)xy.y);
          d = rss_lower_bound_f32(Rf, Tf, t_l1, n1, n2);
        } else {
          const NodeRec n1 = load_node(P.m1.rss, (int)xy.x);
          const NodeRec n2 = load_node(P.m2.rss, (int)xy.y);
          const double la[2] = {n1.e0, n1.e1}, lb[2] = {n2.e0, n2.e1};
          d = __double2float_rd(rss_pair_distance(R, T, n1.axis, n1.To, la, n1.e2, n2.axis, n2.To, lb, n2.e2));
        }
        if (d < min_f) key = (__float_as_uint(d) & ~31u) | (unsigned)lane;
      }
      int tfc1 = 0, tfc2 = 0;
      double tsz1 = 0.0, tsz2 = 0.0;
      if (kPrex && key != 0xffffffffu) {  // the survivor's topology: in flight during the sort below
        load_topo(P.m1.topo, (int)xy.x, tfc1, tsz1);
        load_topo(P.m2.topo, (int)xy.y, tfc2, tsz2);
      }
      if (kStats) bv_tests += n_test;
      const int nkeep = __popc(__ballot_sync(0xffffffffu, key != 0xffffffffu));
      if (nkeep > 0) {
        // full 32-key sort: ordering only the two children of each parent instead (one shuffle) was measured at
        // 83 ms vs 44 ms -- the nearest-first order is what keeps the front small
        key = warp_sort_keys(key, lane);  // ascending: lane 0 = nearest
        const int src = (int)(key & 31u);
        if (kPrex) xy = prex_entry(xy.x, xy.y, tfc1, tsz1, tfc2, tsz2);
        const unsigned long long v = shfl_u64(((unsigned long long)xy.x << 32) | xy.y, src);
        const float dv = __shfl_sync(0xffffffffu, d, src);
        if (lane < nkeep) {  // nearest ends on top of the stack
          const int pos = sp + (nkeep - 1 - lane);
          S.pair[pos] = make_uint2((unsigned)(v >> 32), (unsigned)v);
          S.bound[pos] = dv;
        }
        sp += nkeep;
      }
      __syncwarp();
      DPROF_MARK(1)
    }

    // postprocess: nearest points (model1 frame) -> world with tf1
    if (lane == 0) {
      if (P.min_distance) P.min_distance[q] = min_d;
      if (kTol && P.within) P.within[q] = min_d <= P.stop_below ? 1 : 0;
      if (P.b1) P.b1[q] = S.best_id[0];
      if (P.b2) P.b2[q] = S.best_id[1];
      if (P.enable_nearest_points) {
        const PoseRT tf1 = load_pose(P.tf1, q);
        const V3 w1 = mulv(tf1.R, mk(S.best[0], S.best[1], S.best[2])) + tf1.t;
        const V3 w2 = mulv(tf1.R, mk(S.best[3], S.best[4], S.best[5])) + tf1.t;
        if (P.p1) { P.p1[3 * q] = w1.x; P.p1[3 * q + 1] = w1.y; P.p1[3 * q + 2] = w1.z; }
        if (P.p2) { P.p2[3 * q] = w2.x; P.p2[3 * q + 1] = w2.y; P.p2[3 * q + 2] = w2.z; }
      }
      if (kStats) {
        if (P.n_bv) P.n_bv[q] = bv_tests;
        if (P.n_leaf) P.n_leaf[q] = leaf_tests;
      }
    }
    __syncwarp();
#if FCLGPU_DIST_PROF
    DPROF_MARK(0)
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        atomicAdd(&g_debug_counters[k], (unsigned long long)prof_acc[k]);
        atomicAdd(&g_debug_counters[8 + k], (unsigned long long)prof_cnt[k]);
      }
    }
#endif
  }
}

}  // namespace fclgpu

namespace fclgpu {

// ---------------------------------------------------------------------------------------
// Variant D for collide (thread per query, deferred leaf tests):
// every lane still walks its own query depth first in the reference's order, but the
// triangle-pair tests it meets are appended to a small per-lane FIFO instead of being
// evaluated on the spot, and the lane carries on with the next BVTT node.  The warp
// alternates between
//   * BV rounds   -- all lanes with a non-empty stack run the OBB SAT together, and
//   * leaf rounds -- triggered when enough lanes have queued pairs (or a FIFO is full, or
//                    nobody has BV work): every lane with a queued pair tests its OLDEST pair
//                    and appends the contacts,
// so the two routines never serialise inside one warp.  Per query the leaf tests are
// consumed in DFS order, hence contacts keep the reference's order and the num_max_contacts
// prefix is exact; once the budget is reached the query's stack and FIFO are dropped (the
// reference's canStop()).  BV tests done between a deferred hit and its evaluation are
// speculative work the sequential recursion would not do; they never change a result.
// ---------------------------------------------------------------------------------------

// Out-of-line wrappers for the exact FP64 leaf routines: in the FP32-steered collide kernel they
// run rarely (undecided pairs, contact generation), and keeping them out of line lets the hot
// loop be allocated far fewer registers (higher occupancy).
__device__ __noinline__ bool tri_intersect_outofline(const V3* Pt, const V3* Qt) {
  return tri_intersect(Pt[0], Pt[1], Pt[2], Qt[0], Qt[1], Qt[2]);
}
__device__ __noinline__ bool tri_intersect_rolled_outofline(const V3* Pt, const V3* Qt) {
  return tri_intersect_rolled(Pt[0], Pt[1], Pt[2], Qt[0], Qt[1], Qt[2]);
}
__device__ __noinline__ void tri_contact_info_outofline(const V3* Pt, const V3* Qt, V3* cp, unsigned* nc, double* depth, V3* nrm) {
  tri_contact_info(Pt, Qt, cp, *nc, *depth, *nrm);
}

constexpr int kLeafFifo = 8;  // deferred pairs per lane (power of two)

#ifndef FCLGPU_COLLIDE_MINBLOCKS
#define FCLGPU_COLLIDE_MINBLOCKS 3
#endif
#ifndef FCLGPU_COLLIDE32_MINBLOCKS
#define FCLGPU_COLLIDE32_MINBLOCKS 3
#endif
// kSat32: FP32 steering box test.  kClassify32 (binary mode only): FP32 triangle-pair classification in
// front of the exact SAT, exact routines out of line (measured: +10 % for verdict / pair-id queries, but
// -25 % when contact points are generated because most tested pairs are hits there, so the contact-mode
// instantiation keeps the exact SAT inline).
template <bool kStats, bool kSat32, bool kClassify32>
__global__ void __launch_bounds__(128, kClassify32 ? FCLGPU_COLLIDE32_MINBLOCKS : (kSat32 ? 2 : FCLGPU_COLLIDE_MINBLOCKS))
collide_deferred_kernel(CollideParams P, int leaf_trigger) {
  __shared__ uint2 fifo[4][kLeafFifo][32];  // [warp][slot][lane]
  uint2 stk[kStackCap];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int sp = 0, qhead = 0, qcount = 0;
  long long q = -1;
  PoseRT tf1;
  M3 R;
  V3 T;
  float Rf[9], Tf[3], t_l1 = 0.0f;
  long long count = 0;
  uint32_t bv_tests = 0, leaf_tests = 0;
  bool exhausted = false;

  while (true) {
    // ---- retire / refill lanes that have neither BV work nor queued leaf pairs ----
    const bool need = (sp == 0) && (qcount == 0) && !exhausted;
    if (need && q >= 0) {
      P.num_contacts[q] = (int32_t)count;
      if (kStats) {
        if (P.n_bv) P.n_bv[q] = bv_tests;
        if (P.n_leaf) P.n_leaf[q] = leaf_tests;
      }
      q = -1;
    }
    const long long nq = fetch_work(need, P.work_counter);
    if (nq >= 0 && nq < P.n && !wait_ready(P.ready, P.ready_shift, nq + P.ready_q0)) atomicMin(P.status, (int)FCLGPU_ERR_INPUT_STALLED);
    if (need) {
      if (nq < P.n) {
        q = nq;
        tf1 = load_pose(P.tf1, q, P.ready != nullptr);
        const PoseRT tf2 = load_pose(P.tf2, q, P.ready != nullptr);
        R = mulTM(tf1.R, tf2.R);
        T = mulTv(tf1.R, tf2.t - tf1.t);
        if (kSat32) {
#pragma unroll
          for (int k = 0; k < 9; ++k) Rf[k] = (float)R.m[k];
          Tf[0] = (float)T.x; Tf[1] = (float)T.y; Tf[2] = (float)T.z;
          t_l1 = __double2float_ru((fabs(T.x) + fabs(T.y)) + fabs(T.z));
        }
        count = 0;
        bv_tests = leaf_tests = 0;
        stk[0] = make_uint2(0u, 0u);
        sp = 1;
      } else {
        exhausted = true;
      }
    }
    const unsigned bv_mask = __ballot_sync(0xffffffffu, sp > 0 && qcount < kLeafFifo);
    const unsigned leaf_mask = __ballot_sync(0xffffffffu, qcount > 0);
    if (bv_mask == 0u && leaf_mask == 0u) break;  // every lane exhausted and drained

    const bool leaf_round = (leaf_mask != 0u) && (bv_mask == 0u || __popc(leaf_mask) >= leaf_trigger ||
                                                  __any_sync(0xffffffffu, qcount == kLeafFifo));
    if (leaf_round) {
      if (qcount > 0) {
        const uint2 ids = fifo[wid][qhead][lane];
        qhead = (qhead + 1) & (kLeafFifo - 1);
        qcount--;
        if (kStats) leaf_tests++;
        const int id1 = (int)ids.x, id2 = (int)ids.y;
        V3 Pt[3], Qt[3];
        load_tri(P.m1.tri, id1, Pt);
        load_tri(P.m2.tri, id2, Qt);
#pragma unroll
        for (int k = 0; k < 3; ++k) Qt[k] = mulv(R, Qt[k]) + T;
        bool hit;
        if (kClassify32) {
          // single-precision classification first (bounds_f32.cuh): only undecided pairs -- touching,
          // degenerate, parallel edges -- need the exact FP64 SAT
          const V3 a = Pt[1] - Pt[0], b = Pt[2] - Pt[0], c = Qt[0] - Pt[0], d = Qt[1] - Pt[0], e = Qt[2] - Pt[0];
          const float p2[3] = {(float)a.x, (float)a.y, (float)a.z}, p3[3] = {(float)b.x, (float)b.y, (float)b.z};
          const float q1[3] = {(float)c.x, (float)c.y, (float)c.z}, q2[3] = {(float)d.x, (float)d.y, (float)d.z};
          const float q3[3] = {(float)e.x, (float)e.y, (float)e.z};
          const int cls = tri_classify_f32(p2, p3, q1, q2, q3);
          hit = cls < 0;
          if (cls == 0) hit = tri_intersect_outofline(Pt, Qt);
        } else {
          hit = tri_intersect(Pt[0], Pt[1], Pt[2], Qt[0], Qt[1], Qt[2]);
        }
        if (hit) {
          if (!P.enable_contact) {
            if (count < P.max_contacts) {
              if (P.scratch) {
                if (count < P.stride) {
                  fclgpu_contact* c = P.scratch + q * P.stride + count;
                  c->b1 = id1;
                  c->b2 = id2;
                } else {
                  atomicMin(P.status, (int)FCLGPU_ERR_CONTACT_OVERFLOW);
                }
              }
              count++;
            }
          } else {
            V3 cp[2], nrm;
            unsigned nc;
            double depth;
            if (kClassify32) tri_contact_info_outofline(Pt, Qt, cp, &nc, &depth, &nrm);
            else tri_contact_info(Pt, Qt, cp, nc, depth, nrm);
            if (P.max_contacts < count + (long long)nc)
              nc = (P.max_contacts > count) ? (unsigned)(P.max_contacts - count) : 0u;
            for (unsigned k = 0; k < nc; ++k) {
              if (P.scratch) {
                if (count < P.stride) {
                  fclgpu_contact* c = P.scratch + q * P.stride + count;
                  const V3 pw = mulv(tf1.R, cp[k]) + tf1.t;
                  const V3 nw = mulv(tf1.R, nrm);
                  c->b1 = id1;
                  c->b2 = id2;
                  c->normal[0] = nw.x; c->normal[1] = nw.y; c->normal[2] = nw.z;
                  c->pos[0] = pw.x; c->pos[1] = pw.y; c->pos[2] = pw.z;
                  c->penetration_depth = depth;
                } else {
                  atomicMin(P.status, (int)FCLGPU_ERR_CONTACT_OVERFLOW);
                }
              }
              count++;
            }
          }
          if (count > 0 && P.max_contacts <= count) {  // canStop(): drop everything still pending
            sp = 0;
            qcount = 0;
          }
        }
      }
      continue;
    }

    // ---- BV round ----
    if (sp > 0 && qcount < kLeafFifo) {
      const uint2 e = stk[--sp];
      const int b1 = (int)e.x, b2 = (int)e.y;
      int fc1, fc2;
      double size1, size2;
      bool disjoint;
      if (kSat32) {
        const ObbRec32 n1 = load_obb32(P.m1.obb32, b1);
        const ObbRec32 n2 = load_obb32(P.m2.obb32, b2);
        load_topo(P.m1.topo, b1, fc1, size1);
        load_topo(P.m2.topo, b2, fc2, size2);
        disjoint = obb_certainly_disjoint_f32(Rf, Tf, t_l1, n1, n2);
      } else {
        const NodeRec n1 = load_node(P.m1.obb, b1);
        const NodeRec n2 = load_node(P.m2.obb, b2);
        fc1 = __ldg(P.m1.first_child + b1);
        fc2 = __ldg(P.m2.first_child + b2);
        size1 = n1.size;
        size2 = n2.size;
        disjoint = obb_pair_disjoint(R, T, n1.axis, n1.To, mk(n1.e0, n1.e1, n1.e2), n2.axis, n2.To, mk(n2.e0, n2.e1, n2.e2));
      }
      if (kStats) bv_tests++;
      if (!disjoint) {
        const bool l1 = fc1 < 0, l2 = fc2 < 0;
        if (l1 && l2) {
          fifo[wid][(qhead + qcount) & (kLeafFifo - 1)][lane] = make_uint2((unsigned)(-(fc1 + 1)), (unsigned)(-(fc2 + 1)));
          qcount++;
        } else if (sp + 2 > kStackCap) {
          atomicMin(P.status, (int)FCLGPU_ERR_STACK_OVERFLOW);
          sp = 0;
          qcount = 0;
        } else {
          uint2 left, right;
          if (l2 || (!l1 && (size1 > size2))) {
            left = make_uint2((unsigned)fc1, (unsigned)b2);
            right = make_uint2((unsigned)fc1 + 1u, (unsigned)b2);
          } else {
            left = make_uint2((unsigned)b1, (unsigned)fc2);
            right = make_uint2((unsigned)b1, (unsigned)fc2 + 1u);
          }
          stk[sp++] = right;
          stk[sp++] = left;
        }
      }
      if (FCLGPU_PREFETCH && kSat32 && sp > 0) prefetch_bv32(P.m1, P.m2, stk[sp - 1]);
    }
  }
}

}  // namespace fclgpu

namespace fclgpu {

// ---------------------------------------------------------------------------------------
// Variant P for collide with contact generation (thread per query, POOLED leaf rounds):
// like variant D every lane walks its own query depth first and defers the triangle pairs it
// meets to a per-lane FIFO, but a leaf round treats the FIFOs of the whole warp as one pool:
// the first 32 queued pairs (owner-major, FIFO order inside an owner) are handed one per lane,
// so a lane stuck in a dense contact region gets its backlog tested by the idle lanes.  Each
// helper recomputes the owner's relative pose, runs the exact FP64 triangle SAT and contact
// generation, and stages the result in shared memory; afterwards every owner consumes its own
// results in FIFO (= DFS) order and applies the num_max_contacts budget exactly like the
// reference's leaf test, so the contact list and its truncation are unchanged.
// ---------------------------------------------------------------------------------------
constexpr int kPoolFifo = 16;  // deferred pairs per lane (power of two)

struct __align__(16) PoolStage {  // result of one triangle-pair test, 96 bytes
  double pos[2][3];
  double normal[3];
  double depth;
  int32_t id1, id2;
  int32_t nc;   // -1: no intersection, else number of contact points (0..2)
  int32_t pad;
};

struct __align__(16) PoolWarp {
  uint2 fifo[kPoolFifo][32];
  PoolStage stage[32];
  long long pose[32];
  int incl[32];
  int excl[32];
  int head[32];
};

#ifndef FCLGPU_POOLED_MINBLOCKS
#define FCLGPU_POOLED_MINBLOCKS 4
#endif
#ifndef FCLGPU_POOLED_OUTOFLINE
#define FCLGPU_POOLED_OUTOFLINE 1
#endif
// kTwo: a BV round expands an entry whose boxes are already known to overlap and tests BOTH children at once
// (their records are fetched together: one memory latency per two box tests instead of one per test, and the two
// child records of the split node are neighbours).  Stack entries then carry {b1 | split-first flag, b2 | single
// flag, first_child1, first_child2}, so the expansion needs no load before it can address the children.
// The visiting order, hence every result, is unchanged; the right child's test is merely done early (wasted
// only when the query stops before reaching it).
template <bool kStats, bool kClassify32, bool kTwo = false>
__global__ void __launch_bounds__(128, FCLGPU_POOLED_MINBLOCKS) collide_pooled_kernel(CollideParams P, int leaf_trigger) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PoolWarp& S = reinterpret_cast<PoolWarp*>(smem_raw)[threadIdx.x >> 5];
  // DFS stack: the top entry lives in a register (an expanded node's left child is consumed by the very
  // next BV round), the rest in local memory; sp counts both
  uint2 stk[kTwo ? 1 : kStackCap];
  uint2 top = make_uint2(0u, 0u);
  uint4 stk4[kTwo ? kStackCap : 1];
  uint4 top4 = make_uint4(0u, 0u, 0u, 0u);
  const int lane = threadIdx.x & 31;
  int sp = 0, qhead = 0, qcount = 0;
  long long q = -1;
  M3 R;
  V3 T;
  float Rf[9], Tf[3], t_l1 = 0.0f;
  long long count = 0;
  uint32_t bv_tests = 0, leaf_tests = 0;
  bool exhausted = false;
  long long blk_next = 0, blk_end = 0;  // the warp's current stretch of an ordered batch (fetch_work_blocked)

  while (true) {
    const bool need = (sp == 0) && (qcount == 0) && !exhausted;
    if (need && q >= 0) {
      P.num_contacts[q] = (int32_t)count;
      if (kStats) {
        if (P.n_bv) P.n_bv[q] = bv_tests;
        if (P.n_leaf) P.n_leaf[q] = leaf_tests;
      }
      q = -1;
    }
    // a batch ordered for locality is handed out in warp-sized stretches, so that the lanes of a warp stay on neighbours
    const long long nq = P.order ? fetch_work_blocked(need, P.work_counter, blk_next, blk_end) : fetch_work(need, P.work_counter);
    if (nq >= 0 && nq < P.n && !wait_ready(P.ready, P.ready_shift, nq + P.ready_q0)) atomicMin(P.status, (int)FCLGPU_ERR_INPUT_STALLED);
    if (need && nq != -2) {
      if (nq < P.n) {
        q = P.order ? (long long)__ldg(P.order + nq) : nq;
        const PoseRT tf1 = load_pose(P.tf1, q, P.ready != nullptr);
        const PoseRT tf2 = load_pose(P.tf2, q, P.ready != nullptr);
        R = mulTM(tf1.R, tf2.R);
        T = mulTv(tf1.R, tf2.t - tf1.t);
#pragma unroll
        for (int k = 0; k < 9; ++k) Rf[k] = (float)R.m[k];
        Tf[0] = (float)T.x; Tf[1] = (float)T.y; Tf[2] = (float)T.z;
        t_l1 = __double2float_ru((fabs(T.x) + fabs(T.y)) + fabs(T.z));
        count = 0;
        bv_tests = leaf_tests = 0;
        top = make_uint2(0u, 0u);
        // kTwo: a virtual parent whose only child is the root pair (split-first flag, single flag, child index 0)
        top4 = make_uint4(0x80000000u, 0x80000000u, 0u, 0u);
        sp = 1;
      } else {
        exhausted = true;
      }
    }
    const unsigned bv_mask = __ballot_sync(0xffffffffu, sp > 0 && qcount < kPoolFifo);
    const int total = __reduce_add_sync(0xffffffffu, qcount);  // pool size
    if (bv_mask == 0u && total == 0) break;

    const bool leaf_round = (total > 0) && (bv_mask == 0u || total >= leaf_trigger ||
                                            __any_sync(0xffffffffu, qcount == kPoolFifo));
    if (leaf_round) {
      int incl = qcount;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
      }
      S.incl[lane] = incl;
      S.excl[lane] = incl - qcount;
      S.head[lane] = qhead;
      S.pose[lane] = q;
      __syncwarp();
      const int nproc = total < 32 ? total : 32;
      if (lane < nproc) {
        // owner of pooled pair j = lane: number of lanes whose inclusive count is <= j
        int owner = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1)
          if (S.incl[owner + step - 1] <= lane) owner += step;
        const int slot = (S.head[owner] + (lane - S.excl[owner])) & (kPoolFifo - 1);
        const uint2 ids = S.fifo[slot][owner];
        const long long qo = S.pose[owner];
        const PoseRT tf1 = load_pose(P.tf1, qo, P.ready != nullptr);
        const PoseRT tf2 = load_pose(P.tf2, qo, P.ready != nullptr);
        const M3 Ro = mulTM(tf1.R, tf2.R);
        const V3 To = mulTv(tf1.R, tf2.t - tf1.t);
        V3 Pt[3], Qt[3];
        load_tri(P.m1.tri, (int)ids.x, Pt);
        load_tri(P.m2.tri, (int)ids.y, Qt);
#pragma unroll
        for (int k = 0; k < 3; ++k) Qt[k] = mulv(Ro, Qt[k]) + To;
        PoolStage& out = S.stage[lane];
        out.id1 = (int)ids.x;
        out.id2 = (int)ids.y;
        out.nc = -1;
        bool hit;
        if (kClassify32) {  // binary mode: single-precision classification, exact SAT only when undecided
          const V3 a = Pt[1] - Pt[0], b = Pt[2] - Pt[0], c = Qt[0] - Pt[0], d = Qt[1] - Pt[0], e = Qt[2] - Pt[0];
          const float p2[3] = {(float)a.x, (float)a.y, (float)a.z}, p3[3] = {(float)b.x, (float)b.y, (float)b.z};
          const float q1[3] = {(float)c.x, (float)c.y, (float)c.z}, q2[3] = {(float)d.x, (float)d.y, (float)d.z};
          const float q3[3] = {(float)e.x, (float)e.y, (float)e.z};
          const int cls = tri_classify_f32(p2, p3, q1, q2, q3);
          hit = cls < 0;
          if (cls == 0) hit = tri_intersect_outofline(Pt, Qt);
        } else {
          hit = FCLGPU_POOLED_OUTOFLINE ? tri_intersect_outofline(Pt, Qt) : tri_intersect(Pt[0], Pt[1], Pt[2], Qt[0], Qt[1], Qt[2]);
        }
        if (hit) {
          out.nc = 0;
          if (P.enable_contact) {
            V3 cp[2], nrm;
            unsigned nc;
            double depth;
            if (FCLGPU_POOLED_OUTOFLINE) tri_contact_info_outofline(Pt, Qt, cp, &nc, &depth, &nrm);
            else tri_contact_info(Pt, Qt, cp, nc, depth, nrm);
            const V3 nw = mulv(tf1.R, nrm);  // tf1.linear() * n
            out.normal[0] = nw.x; out.normal[1] = nw.y; out.normal[2] = nw.z;
            out.depth = depth;
            out.nc = (int)nc;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const V3 pw = mulv(tf1.R, cp[k]) + tf1.t;  // tf1 * p
              out.pos[k][0] = pw.x; out.pos[k][1] = pw.y; out.pos[k][2] = pw.z;
            }
          }
        }
      }
      __syncwarp();
      // owners consume their results in FIFO order
      int mine = nproc - (incl - qcount);
      mine = mine < 0 ? 0 : (mine > qcount ? qcount : mine);
      const int first = incl - qcount;
      if (kStats) leaf_tests += mine;
      qhead = (qhead + mine) & (kPoolFifo - 1);
      qcount -= mine;
      for (int k = 0; k < mine; ++k) {
        const PoolStage& r = S.stage[first + k];
        if (r.nc < 0) continue;
        if (!P.enable_contact) {
          if (count < P.max_contacts) {
            if (P.scratch) {
              if (count < P.stride) {
                fclgpu_contact* c = P.scratch + q * P.stride + count;
                c->b1 = r.id1;
                c->b2 = r.id2;
              } else {
                atomicMin(P.status, (int)FCLGPU_ERR_CONTACT_OVERFLOW);
              }
            }
            count++;
          }
        } else {
          long long nc = r.nc;
          if (P.max_contacts < count + nc) nc = (P.max_contacts > count) ? (P.max_contacts - count) : 0;
          for (long long c2 = 0; c2 < nc; ++c2) {
            if (P.scratch) {
              if (count < P.stride) {
                fclgpu_contact* c = P.scratch + q * P.stride + count;
                c->b1 = r.id1;
                c->b2 = r.id2;
                c->normal[0] = r.normal[0]; c->normal[1] = r.normal[1]; c->normal[2] = r.normal[2];
                c->pos[0] = r.pos[c2][0]; c->pos[1] = r.pos[c2][1]; c->pos[2] = r.pos[c2][2];
                c->penetration_depth = r.depth;
              } else {
                atomicMin(P.status, (int)FCLGPU_ERR_CONTACT_OVERFLOW);
              }
            }
            count++;
          }
        }
        if (count > 0 && P.max_contacts <= count) {  // canStop(): drop everything still pending
          sp = 0;
          qcount = 0;
          break;
        }
      }
      __syncwarp();
      continue;
    }

    // ---- BV round, two children per round ----
    if (kTwo) {
      if (sp > 0 && qcount < kPoolFifo) {
        const uint4 e = top4;
        --sp;
        bool have_top = false;
        const bool split1 = (e.x >> 31) != 0u, single = (e.y >> 31) != 0u;
        const int b1 = (int)(e.x & 0x7fffffffu), b2 = (int)(e.y & 0x7fffffffu);
        const int fc1 = (int)e.z, fc2 = (int)e.w;
        if (fc1 < 0 && fc2 < 0) {  // an overlapping leaf pair: the next triangle test in DFS order
          S.fifo[(qhead + qcount) & (kPoolFifo - 1)][lane] = make_uint2((unsigned)(-(fc1 + 1)), (unsigned)(-(fc2 + 1)));
          qcount++;
        } else if (sp + 2 > kStackCap) {
          atomicMin(P.status, (int)FCLGPU_ERR_STACK_OVERFLOW);
          sp = 0;
          qcount = 0;
        } else {
          // child A = left, child B = right; the split node's children are records ci and ci + 1
          const int ci = split1 ? fc1 : fc2, oi = split1 ? b2 : b1;
          const int cj = single ? ci : ci + 1;
          const ObbRec32 A1 = load_obb32(P.m1.obb32, split1 ? ci : oi);
          const ObbRec32 A2 = load_obb32(P.m2.obb32, split1 ? oi : ci);
          const ObbRec32 C = load_obb32(split1 ? P.m1.obb32 : P.m2.obb32, cj);
          int fa, fb, fo;
          double sa, sb, so;
          load_topo(split1 ? P.m1.topo : P.m2.topo, ci, fa, sa);
          load_topo(split1 ? P.m1.topo : P.m2.topo, cj, fb, sb);
          load_topo(split1 ? P.m2.topo : P.m1.topo, oi, fo, so);
          ObbRec32 B1, B2;
#pragma unroll
          for (int k = 0; k < 9; ++k) {
            B1.a[k] = split1 ? C.a[k] : A1.a[k];
            B2.a[k] = split1 ? A2.a[k] : C.a[k];
          }
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            B1.c[k] = split1 ? C.c[k] : A1.c[k];
            B2.c[k] = split1 ? A2.c[k] : C.c[k];
            B1.e[k] = split1 ? C.e[k] : A1.e[k];
            B2.e[k] = split1 ? A2.e[k] : C.e[k];
          }
          B1.s = split1 ? C.s : A1.s;
          B2.s = split1 ? A2.s : C.s;
          const bool hitA = !obb_certainly_disjoint_f32(Rf, Tf, t_l1, A1, A2);
          const bool hitB = !single && !obb_certainly_disjoint_f32(Rf, Tf, t_l1, B1, B2);
          if (kStats) bv_tests += single ? 1u : 2u;
          const bool lo = fo < 0;
          if (hitB) {
            const bool lb = fb < 0;
            const uint4 eb = split1 ? make_uint4((unsigned)cj | ((lo || (!lb && sb > so)) ? 0x80000000u : 0u), (unsigned)oi, (unsigned)fb, (unsigned)fo)
                                    : make_uint4((unsigned)oi | ((lb || (!lo && so > sb)) ? 0x80000000u : 0u), (unsigned)cj, (unsigned)fo, (unsigned)fb);
            if (hitA) {
              stk4[sp++] = eb;  // the slot the popped entry had; child A goes into the register
            } else {
              top4 = eb;
              ++sp;
              have_top = true;
            }
          }
          if (hitA) {
            const bool la = fa < 0;
            top4 = split1 ? make_uint4((unsigned)ci | ((lo || (!la && sa > so)) ? 0x80000000u : 0u), (unsigned)oi, (unsigned)fa, (unsigned)fo)
                          : make_uint4((unsigned)oi | ((la || (!lo && so > sa)) ? 0x80000000u : 0u), (unsigned)ci, (unsigned)fo, (unsigned)fa);
            ++sp;
            have_top = true;
          }
        }
        if (!have_top && sp > 0) top4 = stk4[sp - 1];
      }
      continue;
    }

    // ---- BV round (conservative FP32 box test, see bounds_f32.cuh) ----
    if (sp > 0 && qcount < kPoolFifo) {
      const uint2 e = top;
      --sp;
      bool have_top = false;
      const int b1 = (int)e.x, b2 = (int)e.y;
      int fc1, fc2;
      double size1, size2;
      const ObbRec32 n1 = load_obb32(P.m1.obb32, b1);
      const ObbRec32 n2 = load_obb32(P.m2.obb32, b2);
      load_topo(P.m1.topo, b1, fc1, size1);
      load_topo(P.m2.topo, b2, fc2, size2);
      if (kStats) bv_tests++;
      if (!obb_certainly_disjoint_f32(Rf, Tf, t_l1, n1, n2)) {
        const bool l1 = fc1 < 0, l2 = fc2 < 0;
        if (l1 && l2) {
          S.fifo[(qhead + qcount) & (kPoolFifo - 1)][lane] = make_uint2((unsigned)(-(fc1 + 1)), (unsigned)(-(fc2 + 1)));
          qcount++;
        } else if (sp + 2 > kStackCap) {
          atomicMin(P.status, (int)FCLGPU_ERR_STACK_OVERFLOW);
          sp = 0;
          qcount = 0;
        } else {
          uint2 left, right;
          if (l2 || (!l1 && (size1 > size2))) {
            left = make_uint2((unsigned)fc1, (unsigned)b2);
            right = make_uint2((unsigned)fc1 + 1u, (unsigned)b2);
          } else {
            left = make_uint2((unsigned)b1, (unsigned)fc2);
            right = make_uint2((unsigned)b1, (unsigned)fc2 + 1u);
          }
          stk[sp++] = right;  // slot sp-1 (the old top's slot) .. the register holds the new top
          top = left;
          ++sp;
          have_top = true;
        }
      }
      if (!have_top && sp > 0) top = stk[sp - 1];  // entry below becomes the top
    }
  }
}

}  // namespace fclgpu

namespace fclgpu {

// ---------------------------------------------------------------------------------------
// Variant F for collide, COUNTS ONLY (no contact list, enable_contact = false): one warp per query.
// num_contacts = min(number of intersecting triangle pairs, num_max_contacts) does not depend on the
// order in which the pairs are found, so the warp may expand its BVTT front breadth-wise like the
// distance kernel: a shared-memory stack of node pairs, up to 16 entries expanded per round by two
// lanes each (conservative FP32 box test of the two children), surviving children pushed by ballot
// compaction, leaf pairs queued and tested 32 at a time with the exact FP64 triangle SAT.  All
// 32 lanes have independent record loads in flight, which is what the lane-per-query kernels lack
// once the BVH no longer fits the caches (cfg5: 1M-triangle meshes).  Near the stack limit the
// warp falls back to one expansion per round (plain DFS, depth-bounded), so it never overflows.
// ---------------------------------------------------------------------------------------
#ifndef FCLGPU_FRONT_SEED
#define FCLGPU_FRONT_SEED 5
#endif
// Stack entries per warp: a launch parameter (CollideParams::front_cap, dynamic shared memory).  Rounds are 32 wide while the
// stack is below cap - front_reserve and depth first (one entry per round, growth bounded by the two tree depths) above it;
// on BVHs beyond the caches the width of the rounds is what hides the HBM latency, so big models get the larger stack.
// Measured and rejected for this kernel (the code is in the history): box records fetched by four lanes per
// record through a shared-memory staging area (a third of the L1 wavefronts, 25 % slower: a longer dependent chain per round),
// and wide rounds that expand every popped pair with both children per lane (breadth first: a colliding query reaches its
// first intersecting triangle pair much later).
#ifndef FCLGPU_FRONT_NEXP
#define FCLGPU_FRONT_NEXP 16  // pairs expanded per round (two lanes each)
#endif
#ifndef FCLGPU_FRONT_MINBLOCKS
#define FCLGPU_FRONT_MINBLOCKS 4
#endif
constexpr int kFrontLeafCap = 64;
// A front entry carries everything the NEXT round needs to expand it without touching memory: both node ids, both
// first_child fields (fetched, together with the box records, when the pair was tested) and the firstOverSecond decision
// (bit 31 of b1).  A round is then pop -> expand -> ONE dependent global-load phase (records + topo of the two children)
// -> test -> push, instead of two phases (topo of the popped pair, then the children's records): the kernel is latency
// bound once the BVH leaves the caches (ncu on cfg5: 55 % of the stall samples were long_scoreboard).
struct CollideFront {
  uint4* pair;    // [cap] {b1 | split-first flag, b2, first_child1, first_child2}
  uint2* leaf;    // [kFrontLeafCap]
  uint4* expand;  // [32] {node1, node2, first_child of the node that was NOT split (carried), which side was split}
};
inline __host__ __device__ size_t front_bytes_per_warp(int cap) {
  return (size_t)cap * sizeof(uint4) + kFrontLeafCap * sizeof(uint2) + 32 * sizeof(uint4);
}

__device__ __forceinline__ uint4 front_entry(int b1, int b2, int fc1, double size1, int fc2, double size2) {
  const bool l1 = fc1 < 0, l2 = fc2 < 0;
  const bool first = l2 || (!l1 && (size1 > size2));  // firstOverSecond (bvh_collision_traversal_node-inl.h:78-90)
  return make_uint4((unsigned)b1 | (first ? 0x80000000u : 0u), (unsigned)b2, (unsigned)fc1, (unsigned)fc2);
}

template <bool kStats>
__global__ void __launch_bounds__(128, FCLGPU_FRONT_MINBLOCKS) collide_front_kernel(CollideParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CollideFront S;
  S.pair = reinterpret_cast<uint4*>(smem_raw + (size_t)(threadIdx.x >> 5) * front_bytes_per_warp(P.front_cap));
  S.leaf = reinterpret_cast<uint2*>(S.pair + P.front_cap);
  S.expand = reinterpret_cast<uint4*>(S.leaf + kFrontLeafCap);
  const int wide_limit = P.front_cap - P.front_reserve;  // above it: one entry per round
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;

  // Seed front (FCLGPU_FRONT_SEED levels, see the distance kernel): the pose-independent first levels of the BVTT, expanded
  // once per block; every query starts with one full round over them.  A pair whose boxes overlap although an ancestor's
  // are disjoint holds no intersecting triangles, so skipping the ancestors' tests never changes a count.
  constexpr bool kSeed = FCLGPU_FRONT_SEED > 0;
  __shared__ uint2 s_seed[32];
  __shared__ uint4 s_seed_entry[32];
  __shared__ int s_nseed;
  if (kSeed) {
    if (threadIdx.x < 32) {
      if (lane == 0) s_seed[0] = make_uint2(0u, 0u);
      __syncwarp();
      int n = 1;
      for (int lv = 0; lv <= FCLGPU_FRONT_SEED; ++lv) {
        const bool have = lane < n;
        uint2 e = make_uint2(0u, 0u);
        int fc1 = -1, fc2 = -1;
        double size1 = 0.0, size2 = 0.0;
        if (have) {
          e = s_seed[lane];
          load_topo(P.m1.topo, (int)e.x, fc1, size1);
          load_topo(P.m2.topo, (int)e.y, fc2, size2);
        }
        const bool l1 = fc1 < 0, l2 = fc2 < 0;
        const bool exp = have && !(l1 && l2);
        const unsigned em = __ballot_sync(0xffffffffu, exp);
        const int add = __popc(em);
        if (lv == FCLGPU_FRONT_SEED || add == 0 || n + add > 32) {  // final pairs: keep them as ready-to-expand entries
          if (have) s_seed_entry[lane] = front_entry((int)e.x, (int)e.y, fc1, size1, fc2, size2);
          break;
        }
        __syncwarp();
        if (have) {
          const int pos = lane + __popc(em & lt_mask);
          if (!exp) {
            s_seed[pos] = e;
          } else if (l2 || (!l1 && (size1 > size2))) {  // firstOverSecond
            s_seed[pos] = make_uint2((unsigned)fc1, e.y);
            s_seed[pos + 1] = make_uint2((unsigned)fc1 + 1u, e.y);
          } else {
            s_seed[pos] = make_uint2(e.x, (unsigned)fc2);
            s_seed[pos + 1] = make_uint2(e.x, (unsigned)fc2 + 1u);
          }
        }
        n += add;
        __syncwarp();
      }
      if (lane == 0) s_nseed = n;
    }
    __syncthreads();
  }

  while (true) {
    long long q = 0;
    if (lane == 0) q = (long long)atomicAdd(P.work_counter, 1ull);
    q = __shfl_sync(0xffffffffu, q, 0);
    if (q >= P.n) break;
    if (!wait_ready(P.ready, P.ready_shift, q + P.ready_q0) && lane == 0) atomicMin(P.status, (int)FCLGPU_ERR_INPUT_STALLED);

    M3 R;
    V3 T;
    {
      const PoseRT tf1 = load_pose(P.tf1, q, P.ready != nullptr);
      const PoseRT tf2 = load_pose(P.tf2, q, P.ready != nullptr);
      R = mulTM(tf1.R, tf2.R);
      T = mulTv(tf1.R, tf2.t - tf1.t);
    }
    float Rf[9], Tf[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) Rf[k] = (float)R.m[k];
    Tf[0] = (float)T.x; Tf[1] = (float)T.y; Tf[2] = (float)T.z;
    const float t_l1 = __double2float_ru((fabs(T.x) + fabs(T.y)) + fabs(T.z));

    long long count = 0;
    int sp = 0, nleaf = 0;
    uint32_t bv_tests = 1, leaf_tests = 0;
    if (kSeed) {  // one round over the seed pairs
      const int n_test = s_nseed;
      bool keep = false;
      if (lane < n_test) {
        const uint2 xy = s_seed[lane];
        const ObbRec32 n1 = load_obb32(P.m1.obb32, (int)xy.x), n2 = load_obb32(P.m2.obb32, (int)xy.y);
        keep = !obb_certainly_disjoint_f32(Rf, Tf, t_l1, n1, n2);
      }
      const unsigned km = __ballot_sync(0xffffffffu, keep);
      if (keep) S.pair[__popc(km & lt_mask)] = s_seed_entry[lane];
      sp = __popc(km);
      bv_tests = (uint32_t)n_test;
    } else {  // root pair
      const ObbRec32 n1 = load_obb32(P.m1.obb32, 0), n2 = load_obb32(P.m2.obb32, 0);
      int fc1, fc2;
      double size1, size2;
      load_topo(P.m1.topo, 0, fc1, size1);
      load_topo(P.m2.topo, 0, fc2, size2);
      if (!obb_certainly_disjoint_f32(Rf, Tf, t_l1, n1, n2)) {
        if (lane == 0) S.pair[0] = front_entry(0, 0, fc1, size1, fc2, size2);
        sp = 1;
      }
    }
    __syncwarp();

    while (true) {
      const bool do_leaf = (nleaf >= (P.max_contacts - count < 32 ? P.front_leaf_trigger : 32)) || (sp == 0 && nleaf > 0);
      if (do_leaf) {
        const int k = nleaf < 32 ? nleaf : 32;
        nleaf -= k;
        bool hit = false;
        if (lane < k) {
          const uint2 ids = S.leaf[nleaf + lane];
          V3 Pt[3], Qt[3];
          load_tri(P.m1.tri, (int)ids.x, Pt);
          load_tri(P.m2.tri, (int)ids.y, Qt);
#pragma unroll
          for (int c = 0; c < 3; ++c) Qt[c] = mulv(R, Qt[c]) + T;
          hit = tri_intersect_outofline(Pt, Qt);
        }
        if (kStats) leaf_tests += k;
        count += __popc(__ballot_sync(0xffffffffu, hit));
        __syncwarp();
        if (count >= P.max_contacts) {  // canStop()
          count = P.max_contacts;
          break;
        }
        continue;
      }
      if (sp == 0) break;

      // ---- BV round: entries on the stack are known to overlap; no memory access before the expansion ----
      const bool tight = sp > wide_limit;  // close to the limit: one entry per round (depth-first)
      const int k = tight ? 1 : (sp < 32 ? sp : 32);
      uint4 pr = make_uint4(0u, 0u, 0u, 0u);
      const bool have = lane < k;
      if (have) pr = S.pair[sp - 1 - lane];
      const int fc1 = (int)pr.z, fc2 = (int)pr.w;
      const bool leafpair = have && fc1 < 0 && fc2 < 0;
      const unsigned lm = __ballot_sync(0xffffffffu, leafpair);
      if (leafpair) S.leaf[nleaf + __popc(lm & lt_mask)] = make_uint2((unsigned)(-(fc1 + 1)), (unsigned)(-(fc2 + 1)));
      nleaf += __popc(lm);
      const bool internal = have && !leafpair;
      const unsigned im = __ballot_sync(0xffffffffu, internal);
      const int n_int = __popc(im), rank = __popc(im & lt_mask);
      sp -= k;
      const int n_exp = n_int < FCLGPU_FRONT_NEXP ? n_int : FCLGPU_FRONT_NEXP;
      __syncwarp();  // every lane has read its popped entry before slots are overwritten
      if (internal) {
        if (rank < n_exp) {
          const unsigned b1 = pr.x & 0x7fffffffu;
          if (pr.x >> 31) {  // split model 1's node: children fc1, fc1 + 1 against b2 (whose first_child is carried)
            S.expand[2 * rank] = make_uint4((unsigned)fc1, pr.y, pr.w, 1u);
            S.expand[2 * rank + 1] = make_uint4((unsigned)fc1 + 1u, pr.y, pr.w, 1u);
          } else {
            S.expand[2 * rank] = make_uint4(b1, (unsigned)fc2, pr.z, 0u);
            S.expand[2 * rank + 1] = make_uint4(b1, (unsigned)fc2 + 1u, pr.z, 0u);
          }
        } else {  // not expanded this round: back on the stack
          S.pair[sp + (n_int - 1 - rank)] = pr;
        }
      }
      sp += n_int - n_exp;
      __syncwarp();
      bool keep = false;
      uint4 entry = make_uint4(0u, 0u, 0u, 0u);
      if (lane < 2 * n_exp) {
        const uint4 xy = S.expand[lane];
        // one load phase: both box records, and {first_child, size} of both nodes for the entry to be pushed
        const ObbRec32 n1 = load_obb32(P.m1.obb32, (int)xy.x);
        const ObbRec32 n2 = load_obb32(P.m2.obb32, (int)xy.y);
        int f1, f2;
        double s1, s2;
        load_topo(P.m1.topo, (int)xy.x, f1, s1);
        load_topo(P.m2.topo, (int)xy.y, f2, s2);
        keep = !obb_certainly_disjoint_f32(Rf, Tf, t_l1, n1, n2);
        entry = front_entry((int)xy.x, (int)xy.y, f1, s1, f2, s2);
      }
      if (kStats) bv_tests += 2 * n_exp;
      const unsigned km = __ballot_sync(0xffffffffu, keep);
      if (keep) S.pair[sp + __popc(km & lt_mask)] = entry;
      sp += __popc(km);
      __syncwarp();
    }

    if (lane == 0) {
      P.num_contacts[q] = (int32_t)count;
      if (kStats) {
        if (P.n_bv) P.n_bv[q] = bv_tests;
        if (P.n_leaf) P.n_leaf[q] = leaf_tests;
      }
    }
    __syncwarp();
  }
}

}  // namespace fclgpu

namespace fclgpu {

// ---------------------------------------------------------------------------------------
// Mesh <-> sphere collide: fcl::collide(BVHModel<OBBRSS>, tf1, Sphere, tf2)
// = BVHShapeCollider<OBBRSS, Sphere> -> orientedBVHShapeCollide (collision_func_matrix-inl.h:378-430)
// -> collisionRecurse over the mesh tree with the sphere as a single leaf.  One lane per query,
// depth first in the reference's order (left child first), so contacts -- one per intersecting
// triangle, {b1 = triangle, b2 = -1 (Contact::NONE), pos, -normal, depth} -- come in the reference's
// order and the num_max_contacts prefix is exact.
// BV test: the reference tests the node's OBB against an OBB fitted around the sphere's 12 bound
// vertices; any conservative test visits a superset of the nodes that can hold a contact and therefore
// yields the same contacts.  Here: exact point-to-box distance of the sphere centre (in the mesh
// frame) against the node's OBB, with a margin far above the rounding error -- tighter and 8x cheaper
// than the box-box SAT.  Leaf: sphere_tri_intersect on the triangle moved to the world by tf1, like
// the reference's transformed shapeTriangleIntersect (gjk_solver_libccd-inl.h:479-497).
// ---------------------------------------------------------------------------------------
template <bool kStats>
__global__ void __launch_bounds__(128) collide_mesh_sphere_kernel(CollideParams P, double radius) {
  int stk[kStackCap];
  const int lane = threadIdx.x & 31;
  (void)lane;
  bool exhausted = false;
  while (true) {
    const long long q = fetch_work(!exhausted, P.work_counter);
    if (__all_sync(0xffffffffu, exhausted || q >= P.n)) break;
    if (exhausted) continue;
    if (q >= P.n) {
      exhausted = true;
      continue;
    }
    const PoseRT tf1 = load_pose(P.tf1, q);
    const PoseRT tf2 = load_pose(P.tf2, q);
    const V3 c = tf2.t;                             // sphere centre, world
    const V3 cm = mulTv(tf1.R, c - tf1.t);          // ... in the mesh frame
    const double cm_l1 = (fabs(cm.x) + fabs(cm.y)) + fabs(cm.z);
    long long count = 0;
    uint32_t bv_tests = 0, leaf_tests = 0;
    int sp = 0;
    stk[sp++] = 0;
    while (sp > 0) {
      const int b = stk[--sp];
      const NodeRec nd = load_node(P.m1.obb, b);
      const int fc = __ldg(P.m1.first_child + b);
      if (kStats) bv_tests++;
      {
        const V3 l = mulTv(nd.axis, cm - nd.To);
        const double ex = fmax(fabs(l.x) - nd.e0, 0.0), ey = fmax(fabs(l.y) - nd.e1, 0.0), ez = fmax(fabs(l.z) - nd.e2, 0.0);
        const double scale = (((fabs(nd.To.x) + fabs(nd.To.y)) + fabs(nd.To.z)) + ((nd.e0 + nd.e1) + nd.e2)) + cm_l1;
        const double reach = radius * 1.000001 + 1e-6 * scale;
        if ((ex * ex + ey * ey) + ez * ez > reach * reach) continue;  // certainly out of reach
      }
      if (fc >= 0) {
        if (sp + 2 > kStackCap) {
          atomicMin(P.status, (int)FCLGPU_ERR_STACK_OVERFLOW);
          break;
        }
        stk[sp++] = fc + 1;
        stk[sp++] = fc;  // left child first
        continue;
      }
      const int id = -(fc + 1);
      V3 T[3];
      load_tri(P.m1.tri, id, T);
#pragma unroll
      for (int k = 0; k < 3; ++k) T[k] = mulv(tf1.R, T[k]) + tf1.t;
      if (kStats) leaf_tests++;
      V3 cp, nrm;
      double depth;
      if (sphere_tri_intersect(c, radius, T, cp, depth, nrm)) {
        if (count < P.max_contacts) {
          if (P.scratch) {
            if (count < P.stride) {
              fclgpu_contact* o = P.scratch + q * P.stride + count;
              o->b1 = id;
              o->b2 = -1;
              if (P.enable_contact) {
                o->normal[0] = -nrm.x; o->normal[1] = -nrm.y; o->normal[2] = -nrm.z;
                o->pos[0] = cp.x; o->pos[1] = cp.y; o->pos[2] = cp.z;
                o->penetration_depth = depth;
              }
            } else {
              atomicMin(P.status, (int)FCLGPU_ERR_CONTACT_OVERFLOW);
            }
          }
          count++;
        }
        if (count > 0 && P.max_contacts <= count) sp = 0;  // canStop()
      }
    }
    P.num_contacts[q] = (int32_t)count;
    if (kStats) {
      if (P.n_bv) P.n_bv[q] = bv_tests;
      if (P.n_leaf) P.n_leaf[q] = leaf_tests;
    }
  }
}

// ---------------------------------------------------------------------------------------
// Mesh <-> halfspace / plane collide: fcl::collide(BVHModel<OBBRSS>, tf1, Halfspace | Plane, tf2) =
// BVHShapeCollider<OBBRSS, Shape> -> orientedBVHShapeCollide (collision_func_matrix-inl.h:378-430; cells :841-842) with
// the closed-form leaf tests halfspaceTriangleIntersect / planeTriangleIntersect (gjk_solver_libccd-inl.h:502-540).
// One lane per query, depth first, left child first: one contact per intersecting triangle in the reference's order,
// {b1 = triangle, b2 = -1, pos, -normal, depth}.
// BV test: the reference tests the node's OBB against the shape's OBB -- for a halfspace that is the infinite box
// (utility-inl.h:361-373: nothing is ever pruned), for a plane the slab of zero thickness (:657-669: only the plane's
// normal can separate).  Any conservative test yields the same contacts; here: the signed distance of the box to the
// plane along the normal, s - r > margin (halfspace) or |s| - r > margin (plane), evaluated in the mesh frame.
// kind 0 = halfspace, 1 = plane; (n0, d0) = the shape's own normalised parameters.
// ---------------------------------------------------------------------------------------
template <bool kStats>
__global__ void __launch_bounds__(128) collide_mesh_plane_kernel(CollideParams P, int kind, double n0x, double n0y, double n0z, double d0) {
  int stk[kStackCap];
  bool exhausted = false;
  while (true) {
    const long long q = fetch_work(!exhausted, P.work_counter);
    if (__all_sync(0xffffffffu, exhausted || q >= P.n)) break;
    if (exhausted) continue;
    if (q >= P.n) {
      exhausted = true;
      continue;
    }
    const PoseRT tf1 = load_pose(P.tf1, q);
    const PoseRT tf2 = load_pose(P.tf2, q);
    // transform(shape, tf2): n' = R2 n, d' = d + n' . t2 (halfspace-inl.h:168-180)
    const V3 nw = mulv(tf2.R, mk(n0x, n0y, n0z));
    const double dw = d0 + dot(nw, tf2.t);
    // the same plane in the mesh frame (steering only)
    const V3 nm = mulTv(tf1.R, nw);
    const double dm = dw - dot(nw, tf1.t);
    const double base_scale = (fabs(dm) + fabs(dw)) + ((fabs(tf1.t.x) + fabs(tf1.t.y)) + fabs(tf1.t.z));
    long long count = 0;
    uint32_t bv_tests = 0, leaf_tests = 0;
    int sp = 0;
    stk[sp++] = 0;
    while (sp > 0) {
      const int b = stk[--sp];
      const NodeRec nd = load_node(P.m1.obb, b);
      const int fc = __ldg(P.m1.first_child + b);
      if (kStats) bv_tests++;
      {
        const double sdist = dot(nm, nd.To) - dm;
        const V3 pa = mulTv(nd.axis, nm);  // n . a_i
        const double reach = (nd.e0 * fabs(pa.x) + nd.e1 * fabs(pa.y)) + nd.e2 * fabs(pa.z);
        const double scale = (((fabs(nd.To.x) + fabs(nd.To.y)) + fabs(nd.To.z)) + ((nd.e0 + nd.e1) + nd.e2)) + base_scale;
        const double gap = (kind == 0 ? sdist : fabs(sdist)) - reach;
        if (gap > 1e-9 * scale) continue;  // certainly no triangle of this subtree touches the shape
      }
      if (fc >= 0) {
        if (sp + 2 > kStackCap) {
          atomicMin(P.status, (int)FCLGPU_ERR_STACK_OVERFLOW);
          break;
        }
        stk[sp++] = fc + 1;
        stk[sp++] = fc;  // left child first
        continue;
      }
      const int id = -(fc + 1);
      V3 T[3];
      load_tri(P.m1.tri, id, T);
#pragma unroll
      for (int k = 0; k < 3; ++k) T[k] = mulv(tf1.R, T[k]) + tf1.t;  // tf1 * P
      if (kStats) leaf_tests++;
      V3 cp, nrm;
      double depth;
      const bool hit = kind == 0 ? halfspace_tri_intersect(nw, dw, T, cp, depth, nrm) : plane_tri_intersect(nw, dw, T, cp, depth, nrm);
      if (hit) {
        if (count < P.max_contacts) {
          if (P.scratch) {
            if (count < P.stride) {
              fclgpu_contact* o = P.scratch + q * P.stride + count;
              o->b1 = id;
              o->b2 = -1;
              if (P.enable_contact) {
                o->normal[0] = -nrm.x; o->normal[1] = -nrm.y; o->normal[2] = -nrm.z;
                o->pos[0] = cp.x; o->pos[1] = cp.y; o->pos[2] = cp.z;
                o->penetration_depth = depth;
              }
            } else {
              atomicMin(P.status, (int)FCLGPU_ERR_CONTACT_OVERFLOW);
            }
          }
          count++;
        }
        if (count > 0 && P.max_contacts <= count) sp = 0;  // canStop()
      }
    }
    P.num_contacts[q] = (int32_t)count;
    if (kStats) {
      if (P.n_bv) P.n_bv[q] = bv_tests;
      if (P.n_leaf) P.n_leaf[q] = leaf_tests;
    }
  }
}

// ---------------------------------------------------------------------------------------
// Mesh <-> sphere distance: fcl::distance(BVHModel<OBBRSS>, tf1, Sphere, tf2).  One lane per
// query; the per-query traversal, its bound and the leaf are in mesh_sphere.cuh (shared with the host check).
// The point on the triangle goes back to the mesh frame with tf1^-1, the point on the sphere to the sphere
// frame with tf2^-1 -- the reference leaves both local (sphere_triangle-inl.h:485, 499-508; empty postprocess).
// ---------------------------------------------------------------------------------------
struct DeviceMeshAccessor {
  const DeviceModel& m;
  __device__ __forceinline__ int first_child(int b) const { return __ldg(m.first_child + b); }
  __device__ __forceinline__ void box(int b, M3& axis, V3& To, double& e0, double& e1, double& e2) const {
    const NodeRec nd = load_node(m.obb, b);
    axis = nd.axis;
    To = nd.To;
    e0 = nd.e0;
    e1 = nd.e1;
    e2 = nd.e2;
  }
  __device__ __forceinline__ void box32(int b, ObbRec32& n) const { n = load_obb32(m.obb32, b); }
  __device__ __forceinline__ void tri(int id, V3 T[3]) const { load_tri(m.tri, id, T); }
};

template <bool kStats>
__global__ void __launch_bounds__(128) distance_mesh_sphere_kernel(DistanceParams P, double radius) {
  int stk[kStackCap];
  float stk_lb[kStackCap];
  bool exhausted = false;
  const DeviceMeshAccessor acc{P.m1};
  while (true) {
    const long long q = fetch_work(!exhausted, P.work_counter);
    if (__all_sync(0xffffffffu, exhausted || q >= P.n)) break;
    if (exhausted) continue;
    if (q >= P.n) {
      exhausted = true;
      continue;
    }
    const PoseRT tf1 = load_pose(P.tf1, q);
    const PoseRT tf2 = load_pose(P.tf2, q);
    MeshSphereDistance s;
    mesh_sphere_distance_query(acc, tf1.R, tf1.t, tf2.t, radius, stk, stk_lb, kStackCap, s);
    if (s.overflow) atomicMin(P.status, (int)FCLGPU_ERR_STACK_OVERFLOW);
    if (P.min_distance) P.min_distance[q] = s.min_d;
    if (P.b1) P.b1[q] = s.best;
    if (P.b2) P.b2[q] = -1;  // DistanceResult::NONE
    if (P.enable_nearest_points) {
      V3 a, b;
      if (s.min_d < 0.0) {
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        a = b = mk(nan, nan, nan);
      } else {
        a = inverse_apply(tf1.R, tf1.t, s.on_tri);
        b = inverse_apply(tf2.R, tf2.t, s.on_sph);
      }
      if (P.p1) { P.p1[3 * q] = a.x; P.p1[3 * q + 1] = a.y; P.p1[3 * q + 2] = a.z; }
      if (P.p2) { P.p2[3 * q] = b.x; P.p2[3 * q + 1] = b.y; P.p2[3 * q + 2] = b.z; }
    }
    if (kStats) {
      if (P.n_bv) P.n_bv[q] = s.bv_tests;
      if (P.n_leaf) P.n_leaf[q] = s.leaf_tests;
    }
  }
}

// ---------------------------------------------------------------------------------------
// The same query with the leaf tests taken out of the box loop (default).  In the kernel above a warp pays for the
// triangle routine in almost every iteration -- some lane is at a leaf 80 % of the time while the other ~30 lanes
// wait (ncu: profiles/r01_ncu_sphere_distance.txt).  Here a lane that reaches a leaf (or has just fetched a
// query: the preprocess seed, triangle 0) parks the triangle id and waits; the warp runs a LEAF ROUND -- every
// parked lane runs sphere_tri_distance at once -- when `leaf_trigger` lanes are parked or nobody can do box work.
// Bound, leaf and result arithmetic are the shared routines of mesh_sphere.cuh; only the interleaving differs, and
// the minimum over the triangles does not depend on it.
// ---------------------------------------------------------------------------------------
template <bool kStats, int kMinBlocks>
__global__ void __launch_bounds__(128, kMinBlocks) distance_mesh_sphere_rounds_kernel(DistanceParams P, double radius,
                                                                                      int leaf_trigger, int bound32) {
  int stk[kStackCap];
  float stk_lb[kStackCap];
  const DeviceMeshAccessor acc{P.m1};
  int sp = 0, pend = -1;
  long long q = -1;
  bool exhausted = false;
  PoseRT tf1, tf2;
  V3 cm = mk(0, 0, 0);
  double cm_l1 = 0;
  SphereCentre32 q32 = make_centre32(cm, 0.0, radius);
  MeshSphereDistance s;
  s.min_d = 0;
  s.best = -1;
  s.on_tri = s.on_sph = mk(0, 0, 0);
  s.bv_tests = s.leaf_tests = 0;
  s.overflow = false;
  while (true) {
    const bool idle = (sp == 0) && (pend < 0);
    if (idle && q >= 0) {
      if (P.min_distance) P.min_distance[q] = s.min_d;
      if (P.b1) P.b1[q] = s.best;
      if (P.b2) P.b2[q] = -1;  // DistanceResult::NONE
      if (P.enable_nearest_points) {
        V3 a, b;
        if (s.min_d < 0.0) {
          const double nan = __longlong_as_double(0x7ff8000000000000LL);
          a = b = mk(nan, nan, nan);
        } else {
          a = inverse_apply(tf1.R, tf1.t, s.on_tri);
          b = inverse_apply(tf2.R, tf2.t, s.on_sph);
        }
        if (P.p1) { P.p1[3 * q] = a.x; P.p1[3 * q + 1] = a.y; P.p1[3 * q + 2] = a.z; }
        if (P.p2) { P.p2[3 * q] = b.x; P.p2[3 * q + 1] = b.y; P.p2[3 * q + 2] = b.z; }
      }
      if (kStats) {
        if (P.n_bv) P.n_bv[q] = s.bv_tests;
        if (P.n_leaf) P.n_leaf[q] = s.leaf_tests;
      }
      q = -1;
    }
    const bool need = idle && !exhausted;
    const long long nq = fetch_work(need, P.work_counter);
    if (need) {
      if (nq < P.n) {
        q = nq;
        tf1 = load_pose(P.tf1, q);
        tf2 = load_pose(P.tf2, q);
        cm = mulTv(tf1.R, tf2.t - tf1.t);
        cm_l1 = (fabs(cm.x) + fabs(cm.y)) + fabs(cm.z);
        q32 = make_centre32(cm, cm_l1, radius);
        s.min_d = 1.7976931348623157e308;
        s.best = -1;
        s.bv_tests = s.leaf_tests = 0;
        stk[0] = 0;  // the root, never bound-tested
        stk_lb[0] = -3.0e38f;
        sp = 1;
        pend = 0;  // preprocess: seed with triangle 0
      } else {
        exhausted = true;
      }
    }
    if (__all_sync(0xffffffffu, exhausted && sp == 0 && pend < 0)) break;

    // box step
    if (pend < 0 && sp > 0) {
      --sp;
      const int b = stk[sp];
      if (!((double)stk_lb[sp] >= s.min_d)) {  // canStop(c)
        const int fc = acc.first_child(b);
        if (fc < 0) {
          pend = -(fc + 1);
          if (kStats) s.leaf_tests++;
        } else if (sp + 2 > kStackCap) {
          atomicMin(P.status, (int)FCLGPU_ERR_STACK_OVERFLOW);
          sp = 0;
        } else {
          double d1, d2;
          if (bound32) {
            ObbRec32 n;
            acc.box32(fc, n);
            d1 = (double)sphere_box_lower_bound_f32(n, q32);
            acc.box32(fc + 1, n);
            d2 = (double)sphere_box_lower_bound_f32(n, q32);
          } else {
            M3 ax;
            V3 To;
            double e0, e1, e2;
            acc.box(fc, ax, To, e0, e1, e2);
            d1 = sphere_box_lower_bound(ax, To, e0, e1, e2, cm, cm_l1, radius);
            acc.box(fc + 1, ax, To, e0, e1, e2);
            d2 = sphere_box_lower_bound(ax, To, e0, e1, e2, cm, cm_l1, radius);
          }
          if (kStats) s.bv_tests += 2;
          const bool second_first = d2 < d1;  // nearer child on top
          const int far_b = second_first ? fc : fc + 1, near_b = second_first ? fc + 1 : fc;
          const double far_d = second_first ? d1 : d2, near_d = second_first ? d2 : d1;
          if (far_d < s.min_d) {
            stk[sp] = far_b;
            stk_lb[sp] = float_below(far_d);
            sp++;
          }
          if (near_d < s.min_d) {
            stk[sp] = near_b;
            stk_lb[sp] = float_below(near_d);
            sp++;
          }
        }
      }
    }

    // leaf round?
    const unsigned parked = __ballot_sync(0xffffffffu, pend >= 0);
    const unsigned movable = __ballot_sync(0xffffffffu, pend < 0 && (sp > 0 || !exhausted));
    if (parked != 0 && (__popc(parked) >= leaf_trigger || movable == 0)) {
      if (pend >= 0) {
        mesh_sphere_leaf(acc, pend, tf1.R, tf1.t, tf2.t, radius, s);
        pend = -1;
        if (s.min_d == -1.0) sp = 0;  // a triangle within the radius: final, drop the rest of the stack
      }
    }
  }
}

}  // namespace fclgpu
