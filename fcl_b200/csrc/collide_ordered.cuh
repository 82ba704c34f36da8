// collide_ordered.cuh — collide() WITH a contact list: one warp per query, DFS-ORDERED front.
//
// Reference semantics (file:line under /root/reference/include/fcl):
//   collisionRecurse   narrowphase/detail/traversal/traversal_recurse-inl.h:84-130 (left subtree before right, canStop
//                      between the two)
//   leaf test / budget narrowphase/detail/traversal/collision/mesh_collision_traversal_node-inl.h:527-620
//   canStop            collision_request-inl.h:77-82
// The reference's contact list is: the intersecting triangle pairs in the depth-first order of the BVTT, each
// contributing 1 entry (binary mode) or <= 2 contact points (contact mode), cut after num_max_contacts entries.
// That list does not depend on WHEN a box test is evaluated, only on the ORDER in which the leaf pairs are consumed.
// So the warp keeps the query's BVTT front as a list in depth-first order (a shared-memory stack whose top is the
// next node pair of the recursion) and works on its head breadth-wise:
//   * BV round: the top <= 32 entries are popped; the leading run of leaf pairs -- nothing precedes them any more --
//     moves to the leaf FIFO; the first <= 16 internal entries are replaced IN PLACE by those of their two children
//     (firstOverSecond picks the side to split, left child first) whose boxes are not certainly disjoint -- two
//     lanes per entry run the conservative FP32 box test of bounds_f32.cuh; everything else keeps its place.  Output
//     positions come from two ballots (every entry yields 0, 1 or 2 entries).
//   * leaf round (32 queued pairs, or the stack has run empty): 32 lanes run the exact FP64 intersect_Triangle (and
//     the contact computation) at once; an ordered prefix count over the lanes gives every contact its index in the
//     query's list, the budget cuts the list exactly where the reference's leaf test stops appending, and once it is
//     exhausted the query ends (canStop).
// Box tests that the sequential recursion would have skipped after its early stop are speculative work; they never
// change a result.  Near the stack limit the warp expands one entry per round: plain depth first, whose growth is
// bounded by the tree depths (checked on the host).
//
// Output: contacts are staged in a per-warp scratch (resident warps x stride x 64 B: L2-resident) and, when the query
// retires, appended to the caller's dense pool with one atomic reservation: ONE launch per batch, no per-query
// scratch, no scan / compaction passes.  Blocks therefore land in completion order; contact_offsets[i] names the
// start of query i's block (see include/fclgpu.h).
#pragma once
#include "traversal.cuh"

namespace fclgpu {

constexpr int kOrdCap = 384;      // BVTT front entries per warp
constexpr int kOrdLeafCap = 64;   // leaf FIFO (ring), power of two
constexpr int kOrdWarps = 4;      // warps per block
#ifndef FCLGPU_ORD_MINBLOCKS
#define FCLGPU_ORD_MINBLOCKS 4
#endif
#ifndef FCLGPU_ORD_OUTOFLINE
#define FCLGPU_ORD_OUTOFLINE 3   // bit 0: triangle SAT out of line, bit 1: contact computation out of line
#endif
// Seed front (FCLGPU_ORD_SEED levels, like the distance kernel's): the first levels of the BVTT are expanded whatever the pose --
// which node of a pair splits depends on the two trees only -- so one warp per block expands the root pair that many
// levels once (in depth-first order, at most 32 pairs) and every query starts with ONE round that box-tests those pairs,
// instead of the 1-, 2-, 4-, ... lane rounds that lead there.  A pair whose boxes overlap although an ancestor's are
// disjoint cannot hold an intersecting triangle pair (both ancestors' boxes would contain the intersection point), so
// skipping the ancestors' tests never adds a contact.
#ifndef FCLGPU_ORD_SEED
#define FCLGPU_ORD_SEED 5
#endif
// Pre-expanded entries (like the distance kernel's FCLGPU_DIST_PREX): the lane that tests a child pair also reads the two
// topology records -- in the same load phase as the box records -- and stores the pair in the form the next round needs:
// {triangle ids} for a leaf pair, {first child of the node firstOverSecond splits | side, other node} otherwise.  A BV round
// is then pop -> expand -> ONE dependent load phase -> test -> push instead of two (topology of the popped pairs first).
#ifndef FCLGPU_ORD_PREX
#define FCLGPU_ORD_PREX 1
#endif
#ifndef FCLGPU_ORD_NEXP
#define FCLGPU_ORD_NEXP 16          // entries expanded per BV round (two lanes each)
#endif
#ifndef FCLGPU_ORD_LEAF_TRIGGER
#define FCLGPU_ORD_LEAF_TRIGGER 32  // queued triangle pairs that start a leaf round
#endif
#ifndef FCLGPU_ORD_ROLLED
#define FCLGPU_ORD_ROLLED 1      // rolled-loop triangle SAT (compact code)
#endif

// out-of-line leaf routines taking their operands BY VALUE (registers), so that the caller's triangles need no stack slots
struct TriPair {
  V3 p0, p1, p2, q0, q1, q2;
};
struct ContactInfo {
  V3 c0, c1, n;
  double depth;
  unsigned nc;
};
__device__ __noinline__ bool ord_tri_intersect(const TriPair t) {
  return FCLGPU_ORD_ROLLED ? tri_intersect_rolled(t.p0, t.p1, t.p2, t.q0, t.q1, t.q2) : tri_intersect(t.p0, t.p1, t.p2, t.q0, t.q1, t.q2);
}
__device__ __noinline__ ContactInfo ord_contact_info(const TriPair t) {
  const V3 P[3] = {t.p0, t.p1, t.p2}, Q[3] = {t.q0, t.q1, t.q2};
  V3 cp[2];
  ContactInfo o;
  tri_contact_info(P, Q, cp, o.nc, o.depth, o.n);
  o.c0 = cp[0];
  o.c1 = cp[1];
  return o;
}

struct __align__(16) OrderedFront {
  uint2 pair[kOrdCap];
  uint2 leaf[kOrdLeafCap];
  uint2 expand[32];
  int base[16];
  double tf1[12];  // pose of model 1 (contacts go to the world frame with it)
  double rel[12];  // exact relative pose R = R1^T R2, T = R1^T (t2 - t1)
};

struct OrderedParams {
  CollideParams C;             // models, poses, request, num_contacts, counters, status, ready flags
  fclgpu_contact* pool;        // caller's dense contact array (or nullptr: counts + offsets only)
  long long pool_capacity;
  unsigned long long* cursor;  // running number of contacts appended to the pool
  long long* starts;           // [n] start of query i's block in the pool (or nullptr)
  int depth_sum;               // depth(model1) + depth(model2)
  int format;                  // FCLGPU_CONTACT_*: layout of the pool records (the staging always holds full records)
  int discard_stage;           // drop the staging lines from the L2 after the copy to the pool (global staging only)
  int smem_stage;              // > 0: contacts are staged in shared memory (that many slots per warp, >= C.stride) instead
                               // of the global per-warp scratch: the staged list never leaves the SM before it reaches the pool
};

template <bool kStats>
__global__ void __launch_bounds__(kOrdWarps * 32, FCLGPU_ORD_MINBLOCKS) collide_ordered_kernel(OrderedParams Q) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  OrderedFront& S = reinterpret_cast<OrderedFront*>(smem_raw)[threadIdx.x >> 5];
  const CollideParams& P = Q.C;
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const long long gwarp = (long long)blockIdx.x * kOrdWarps + (threadIdx.x >> 5);
  fclgpu_contact* const stage =
      Q.smem_stage > 0 ? reinterpret_cast<fclgpu_contact*>(smem_raw + sizeof(OrderedFront) * kOrdWarps) + (size_t)(threadIdx.x >> 5) * Q.smem_stage
                       : (P.scratch ? P.scratch + gwarp * P.stride : nullptr);
  // a normal round adds at most 16 entries; the depth-first fallback at most depth_sum + 1 in total
  const int normal_limit = kOrdCap - (Q.depth_sum + 2) - 16;
  const bool coherent = P.ready != nullptr;

  constexpr bool kSeed = FCLGPU_ORD_SEED > 0;
  constexpr bool kPrex = FCLGPU_ORD_PREX != 0;
  __shared__ uint2 s_seed[32];
  __shared__ uint2 s_seed_entry[32];
  __shared__ int s_nseed;
  if (kSeed) {
    if (threadIdx.x < 32) {
      if (lane == 0) s_seed[0] = make_uint2(0u, 0u);
      __syncwarp();
      int n = 1;
      for (int lv = 0; lv < FCLGPU_ORD_SEED; ++lv) {
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
          const int pos = lane + __popc(em & lt_mask);  // every expanded entry ahead of this one takes one more slot
          if (!exp) {
            s_seed[pos] = e;
          } else if (l2 || (!l1 && (size1 > size2))) {  // firstOverSecond; left child first = depth-first order
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
      if (kPrex && lane < n) {
        const uint2 e = s_seed[lane];
        int fc1, fc2;
        double size1, size2;
        load_topo(P.m1.topo, (int)e.x, fc1, size1);
        load_topo(P.m2.topo, (int)e.y, fc2, size2);
        s_seed_entry[lane] = prex_entry(e.x, e.y, fc1, size1, fc2, size2);
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

    float Rf[9], Tf[3], t_l1;
    {
      const PoseRT tf1 = load_pose(P.tf1, q, coherent);
      const PoseRT tf2 = load_pose(P.tf2, q, coherent);
      const M3 R = mulTM(tf1.R, tf2.R);             // R1^T R2                 (math/geometry-inl.h:681-682)
      const V3 T = mulTv(tf1.R, tf2.t - tf1.t);     // R1^T (t2 - t1)
#pragma unroll
      for (int k = 0; k < 9; ++k) Rf[k] = (float)R.m[k];
      Tf[0] = (float)T.x; Tf[1] = (float)T.y; Tf[2] = (float)T.z;
      t_l1 = __double2float_ru((fabs(T.x) + fabs(T.y)) + fabs(T.z));
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          S.tf1[k] = tf1.R.m[k];
          S.rel[k] = R.m[k];
        }
        S.tf1[9] = tf1.t.x; S.tf1[10] = tf1.t.y; S.tf1[11] = tf1.t.z;
        S.rel[9] = T.x; S.rel[10] = T.y; S.rel[11] = T.z;
      }
    }

    long long count = 0;
    int sp = 0, nleaf = 0, head = 0;
    uint32_t bv_tests = 1, leaf_tests = 0;
    if (kSeed) {  // one round over the pose-independent seed pairs (depth-first order: the first survivor ends on top)
      const int n_test = s_nseed;
      bool keep = false;
      uint2 xy = make_uint2(0u, 0u);
      if (lane < n_test) {
        xy = s_seed[lane];
        const ObbRec32 n1 = load_obb32(P.m1.obb32, (int)xy.x), n2 = load_obb32(P.m2.obb32, (int)xy.y);
        keep = !obb_certainly_disjoint_f32(Rf, Tf, t_l1, n1, n2);
      }
      const unsigned km = __ballot_sync(0xffffffffu, keep);
      sp = __popc(km);
      if (keep) S.pair[sp - 1 - __popc(km & lt_mask)] = kPrex ? s_seed_entry[lane] : xy;
      bv_tests = (uint32_t)n_test;
    } else {  // root pair
      const ObbRec32 n1 = load_obb32(P.m1.obb32, 0), n2 = load_obb32(P.m2.obb32, 0);
      if (!obb_certainly_disjoint_f32(Rf, Tf, t_l1, n1, n2)) {
        uint2 root = make_uint2(0u, 0u);
        if (kPrex) {
          int fc1, fc2;
          double size1, size2;
          load_topo(P.m1.topo, 0, fc1, size1);
          load_topo(P.m2.topo, 0, fc2, size2);
          root = prex_entry(0u, 0u, fc1, size1, fc2, size2);
        }
        if (lane == 0) S.pair[0] = root;
        sp = 1;
      }
    }
    __syncwarp();

    while (true) {
      if (nleaf >= FCLGPU_ORD_LEAF_TRIGGER || (sp == 0 && nleaf > 0)) {
        // ---- leaf round: the next <= 32 triangle pairs of the depth-first order ----
        const int k = nleaf < 32 ? nleaf : 32;
        const bool mine = lane < k;
        uint2 ids = make_uint2(0u, 0u);
        if (mine) ids = S.leaf[(head + lane) & (kOrdLeafCap - 1)];
        head = (head + k) & (kOrdLeafCap - 1);
        nleaf -= k;
        int ncp = 0;  // entries this pair adds to the list: 1 per hit (binary mode) or its 0..2 contact points
        V3 cp0 = mk(0, 0, 0), cp1 = mk(0, 0, 0), nrm = mk(0, 0, 0);
        double depth = 0.0;
        if (mine) {
          M3 R;
#pragma unroll
          for (int c = 0; c < 9; ++c) R.m[c] = S.rel[c];
          const V3 T = mk(S.rel[9], S.rel[10], S.rel[11]);
          V3 Pt[3], Qt[3];
          load_tri(P.m1.tri, (int)ids.x, Pt);
          load_tri(P.m2.tri, (int)ids.y, Qt);
#pragma unroll
          for (int c = 0; c < 3; ++c) Qt[c] = mulv(R, Qt[c]) + T;
          const TriPair tp{Pt[0], Pt[1], Pt[2], Qt[0], Qt[1], Qt[2]};
          bool hit;
          if (FCLGPU_ORD_OUTOFLINE & 1) hit = ord_tri_intersect(tp);
          else if (FCLGPU_ORD_ROLLED) hit = tri_intersect_rolled(Pt[0], Pt[1], Pt[2], Qt[0], Qt[1], Qt[2]);
          else hit = tri_intersect(Pt[0], Pt[1], Pt[2], Qt[0], Qt[1], Qt[2]);
          if (hit) {
            ncp = 1;
            if (P.enable_contact) {
              if (FCLGPU_ORD_OUTOFLINE & 2) {
                const ContactInfo ci = ord_contact_info(tp);
                cp0 = ci.c0; cp1 = ci.c1; nrm = ci.n; depth = ci.depth;
                ncp = (int)ci.nc;
              } else {
                V3 cp[2];
                unsigned nc;
                tri_contact_info(Pt, Qt, cp, nc, depth, nrm);
                cp0 = cp[0]; cp1 = cp[1];
                ncp = (int)nc;
              }
            }
          }
        }
        if (kStats) leaf_tests += k;
        const unsigned m1 = __ballot_sync(0xffffffffu, ncp & 1), m2 = __ballot_sync(0xffffffffu, ncp >> 1);
        const long long before = __popc(m1 & lt_mask) + 2 * __popc(m2 & lt_mask);
        const long long total = __popc(m1) + 2 * __popc(m2);
        const long long room = P.max_contacts - count;  // > 0 here
        if (ncp > 0 && stage != nullptr && before < room) {
          // mesh_collision_traversal_node-inl.h:594-600: a pair that does not fit entirely contributes its first points
          const bool two = ncp == 2 && before + 2 <= room;
          const long long slot = count + before;
          if (slot + (two ? 1 : 0) < P.stride) {
            fclgpu_contact* c = stage + slot;
            c->b1 = (int)ids.x;
            c->b2 = (int)ids.y;
            if (two) {
              c[1].b1 = (int)ids.x;
              c[1].b2 = (int)ids.y;
            }
            if (P.enable_contact) {
              M3 R1;
#pragma unroll
              for (int k2 = 0; k2 < 9; ++k2) R1.m[k2] = S.tf1[k2];
              const V3 t1 = mk(S.tf1[9], S.tf1[10], S.tf1[11]);
              const V3 nw3 = mulv(R1, nrm);        // tf1.linear() * n
              const V3 pw = mulv(R1, cp0) + t1;    // tf1 * p
              c->normal[0] = nw3.x; c->normal[1] = nw3.y; c->normal[2] = nw3.z;
              c->pos[0] = pw.x; c->pos[1] = pw.y; c->pos[2] = pw.z;
              c->penetration_depth = depth;
              if (two) {
                const V3 pv = mulv(R1, cp1) + t1;
                c[1].normal[0] = nw3.x; c[1].normal[1] = nw3.y; c[1].normal[2] = nw3.z;
                c[1].pos[0] = pv.x; c[1].pos[1] = pv.y; c[1].pos[2] = pv.z;
                c[1].penetration_depth = depth;
              }
            }
          } else {
            atomicMin(P.status, (int)FCLGPU_ERR_CONTACT_OVERFLOW);
          }
        }
        count += total < room ? total : room;
        __syncwarp();
        if (count > 0 && P.max_contacts <= count) break;  // canStop(): everything still pending is dropped
        continue;
      }
      if (sp == 0) break;

      // ---- BV round on the head of the depth-first list ----
      const bool tight = sp > normal_limit;
      const int k = tight ? 1 : (sp < 32 ? sp : 32);
      uint2 pr = make_uint2(0u, 0u);
      int fc1 = 0, fc2 = 0;
      double size1 = 0.0, size2 = 0.0;
      const bool have = lane < k;
      if (have) {
        pr = S.pair[sp - 1 - lane];
        if (!kPrex) {
          load_topo(P.m1.topo, (int)pr.x, fc1, size1);
          load_topo(P.m2.topo, (int)pr.y, fc2, size2);
        }
      }
      const bool l1 = fc1 < 0, l2 = fc2 < 0;
      const bool leafpair = have && (kPrex ? (pr.x >> 31) != 0u : (l1 && l2));
      const bool internal = have && !leafpair;
      const unsigned im = __ballot_sync(0xffffffffu, internal);
      const int lead = im ? (__ffs(im) - 1) : k;  // leaf pairs ahead of every internal entry: next in DFS order
      if (lane < lead)
        S.leaf[(head + nleaf + lane) & (kOrdLeafCap - 1)] =
            kPrex ? make_uint2(pr.x & 0x7fffffffu, pr.y) : make_uint2((unsigned)(-(fc1 + 1)), (unsigned)(-(fc2 + 1)));
      nleaf += lead;
      const int n_int = __popc(im), rank = __popc(im & lt_mask);
      const int n_exp = n_int < FCLGPU_ORD_NEXP ? n_int : FCLGPU_ORD_NEXP;
      const bool expanded = internal && rank < n_exp;
      __syncwarp();  // every lane holds its popped entry before slots are overwritten
      if (expanded) {
        if (kPrex) {
          const unsigned fc = pr.x & 0x3fffffffu;
          if (pr.x & 0x40000000u) {  // model 1's node is split
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
      }
      __syncwarp();
      bool keep = false;
      uint2 xy = make_uint2(0u, 0u);
      if (lane < 2 * n_exp) {
        xy = S.expand[lane];
        const ObbRec32 n1 = load_obb32(P.m1.obb32, (int)xy.x);
        const ObbRec32 n2 = load_obb32(P.m2.obb32, (int)xy.y);
        if (kPrex) {
          int tfc1, tfc2;
          double tsz1, tsz2;
          load_topo(P.m1.topo, (int)xy.x, tfc1, tsz1);
          load_topo(P.m2.topo, (int)xy.y, tfc2, tsz2);
          keep = !obb_certainly_disjoint_f32(Rf, Tf, t_l1, n1, n2);
          xy = prex_entry(xy.x, xy.y, tfc1, tsz1, tfc2, tsz2);
        } else {
          keep = !obb_certainly_disjoint_f32(Rf, Tf, t_l1, n1, n2);
        }
      }
      if (kStats) bv_tests += 2 * n_exp;
      const unsigned km = __ballot_sync(0xffffffffu, keep);
      int cnt = 0;  // entries this popped entry leaves on the list
      if (have && lane >= lead) cnt = expanded ? __popc((km >> (2 * rank)) & 3u) : 1;
      const unsigned c1 = __ballot_sync(0xffffffffu, cnt & 1), c2 = __ballot_sync(0xffffffffu, cnt >> 1);
      const int pre = __popc(c1 & lt_mask) + 2 * __popc(c2 & lt_mask);
      const int new_sp = sp - k + __popc(c1) + 2 * __popc(c2);
      if (have && lane >= lead) {
        if (expanded) S.base[rank] = pre;
        else S.pair[new_sp - 1 - pre] = pr;  // keeps its place
      }
      __syncwarp();
      if (keep) {
        const int p = S.base[lane >> 1] + ((lane & 1) ? (int)((km >> (lane - 1)) & 1u) : 0);  // left child first
        S.pair[new_sp - 1 - p] = xy;
      }
      sp = new_sp;
      __syncwarp();
    }

    // ---- retire: reserve the query's block in the dense pool and move the staged contacts there ----
    if (lane == 0) {
      P.num_contacts[q] = (int32_t)count;
      if (kStats) {
        if (P.n_bv) P.n_bv[q] = bv_tests;
        if (P.n_leaf) P.n_leaf[q] = leaf_tests;
      }
    }
    if (stage != nullptr) {
      long long stored = count < P.stride ? count : P.stride;
      unsigned long long off = 0;
      if (lane == 0 && stored > 0) off = atomicAdd(Q.cursor, (unsigned long long)stored);
      off = shfl_u64(off, 0);
      if (lane == 0 && Q.starts) Q.starts[q] = (long long)off;
      if (Q.pool != nullptr && stored > 0) {
        if ((long long)off + stored > Q.pool_capacity) {
          if (lane == 0) atomicMin(P.status, (int)FCLGPU_ERR_CONTACT_OVERFLOW);
          stored = Q.pool_capacity > (long long)off ? Q.pool_capacity - (long long)off : 0;
        }
        // the pool is written once and never read by this kernel: streaming stores (evict first), so that the 2 GB
        // passing through do not push the staging lines -- which ARE re-read and overwritten -- out of the L2
        if (Q.format == FCLGPU_CONTACT_IDS) {
          uint2* dst = reinterpret_cast<uint2*>(Q.pool) + off;
          for (long long i = lane; i < stored; i += 32) __stcs(dst + i, make_uint2((unsigned)stage[i].b1, (unsigned)stage[i].b2));
        } else if (Q.format == FCLGPU_CONTACT_F32) {
          // 40-byte records = 5 x 8 bytes; one lane per record
          float2* dst = reinterpret_cast<float2*>(reinterpret_cast<fclgpu_contact_f32*>(Q.pool) + off);
          for (long long i = lane; i < stored; i += 32) {
            const fclgpu_contact c = stage[i];
            float2* o = dst + 5 * i;
            __stcs(o + 0, make_float2(__int_as_float(c.b1), __int_as_float(c.b2)));
            __stcs(o + 1, make_float2((float)c.normal[0], (float)c.normal[1]));
            __stcs(o + 2, make_float2((float)c.normal[2], (float)c.pos[0]));
            __stcs(o + 3, make_float2((float)c.pos[1], (float)c.pos[2]));
            __stcs(o + 4, make_float2((float)c.penetration_depth, 0.0f));
          }
        } else {
        const int4* src = reinterpret_cast<const int4*>(stage);
        int4* dst = reinterpret_cast<int4*>(Q.pool + off);
        for (long long i = lane; i < stored * 4; i += 32) __stcs(dst + i, src[i]);
        }
        if (Q.discard_stage && Q.smem_stage == 0) {
          // the staged copy is dead now: drop its L2 lines instead of letting them be written back to HBM some day
          // (discard.global.L2 leaves the contents undetermined; the warp's next query rewrites them before reading)
          __syncwarp();
          const char* base = reinterpret_cast<const char*>(stage);
          const long long lines = (stored * (long long)sizeof(fclgpu_contact)) / 128;  // whole 128-byte lines only
          for (long long l = lane; l < lines; l += 32) asm volatile("discard.global.L2 [%0], 128;" ::"l"(base + 128 * l) : "memory");
        }
      }
    }
    __syncwarp();
  }
}

}  // namespace fclgpu
