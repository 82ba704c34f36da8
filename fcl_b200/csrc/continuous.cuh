// continuous.cuh — continuous collision by conservative advancement, both bodies TRANSLATING.
//
// Reference semantics (file:line under /root/reference/include/fcl):
//   continuousCollide(o1, tf1_beg, tf1_end, o2, tf2_beg, tf2_end, request, result) with request.ccd_motion_type = CCDM_TRANS
//   and request.ccd_solver_type = CCDC_CONSERVATIVE_ADVANCEMENT      narrowphase/continuous_collision-inl.h:441-452, 93-117,
//                                                                     355-377, 302-352
//   BVHConservativeAdvancement<OBBRSS> -> conservativeAdvancementMeshOriented
//                                          narrowphase/detail/conservative_advancement_func_matrix-inl.h:692-712, 149-219
//   MeshConservativeAdvancementTraversalNodeOBBRSS (BVTesting / leafTesting / canStop)
//                                          detail/traversal/distance/mesh_conservative_advancement_traversal_node.h:163-215,
//                                          -inl.h:432-470, 566-713
//   TranslationMotion, its triangle motion bound   math/motion/translation_motion-inl.h:47-134,
//                                                  math/motion/triangle_motion_bound_visitor-inl.h:218-227
//   the box-level motion bound is the generic TBVMotionBoundVisitorVisitImpl = 0 for BV = OBBRSS
//   (math/motion/tbv_motion_bound_visitor-inl.h:52-62: only RSS is specialised), so a pruned node pair never shortens
//   the step and the witness points of the box distance are never read: one advancement step is the reference's
//   distanceRecurse (nearer child first, canStop(c) = c >= min_distance) whose leaf test also lowers
//   delta_t = min(delta_t, d / ((v1 - v2) . n)) along the unit vector n between the closest points, turned by body 1's
//   current rotation.  Which triangle pairs are visited -- hence delta_t and the time of contact -- depends on the
//   visiting order, so this kernel keeps the reference's order and its exact FP64 RSS distance (like traversal 0 of
//   distance()): one lane per query, explicit stack, the advancement loop around it.
// No transcendental function is involved (quaternion <-> matrix conversions, normalisation: + - * / sqrt, all IEEE), so
// the results are the oracle's bit for bit.  The other motion types (InterpMotion, ScrewMotion, SplineMotion) go through
// sin / cos / atan2, whose last bits differ between libm and the device library: not built.
#pragma once
#include "bvh_merge.cuh"
#include "traversal.cuh"

namespace fclgpu {

struct ContinuousParams {
  DeviceModel m1, m2;
  const double* tf1_beg;  // 12 doubles per query (or nullptr = identity)
  const double* tf1_end;
  const double* tf2_beg;
  const double* tf2_end;
  long long n;
  const int32_t* start_hits;  // numContacts of collide() at the start configuration (request: 1 contact, no contact data)
  int32_t* is_collide;        // outputs, any may be nullptr
  double* toc;
  double* contact_tf1;        // 12 doubles per query
  double* contact_tf2;
  int32_t* iterations;        // distance traversals run
  uint32_t* n_bv;
  uint32_t* n_leaf;
  unsigned long long* work_counter;
  int* status;
};

struct TranslationState {  // TranslationMotion<S>
  double rot[4];           // quaternion of tf_beg's rotation (x, y, z, w)
  V3 start, range;
  M3 R;                    // current transform
  V3 t;
};

__device__ __forceinline__ void motion_init(TranslationState& m, const PoseRT& beg, const PoseRT& end) {
  merge_detail::quat_of_matrix(beg.R.m, m.rot);
  m.start = beg.t;
  m.range = end.t - beg.t;
  m.R = beg.R;
  m.t = beg.t;
}
__device__ __forceinline__ void motion_integrate(TranslationState& m, double dt) {  // translation_motion-inl.h:74-84
  if (dt > 1) dt = 1;
  merge_detail::matrix_of_quat(m.rot, m.R.m);
  m.t = m.start + m.range * dt;
}
__device__ __forceinline__ void store_pose(double* dst, long long q, const M3& R, const V3& t) {
  if (!dst) return;
  double* o = dst + 12 * q;
#pragma unroll
  for (int k = 0; k < 9; ++k) o[k] = R.m[k];
  o[9] = t.x; o[10] = t.y; o[11] = t.z;
}

template <bool kStats>
__global__ void __launch_bounds__(128) ca_translation_kernel(ContinuousParams P) {
  uint2 stk[kStackCap];
  double stk_d[kStackCap];
  int sp = 0;
  long long q = -1;
  bool exhausted = false;
  TranslationState mo1, mo2;
  double R0[4] = {0, 0, 0, 1};  // body 1's current rotation as a quaternion (getCurrentRotation)
  M3 R;
  V3 T;
  double min_d = 0, delta_t = 1, toc = 0;
  int iters = 0;
  uint32_t bv_tests = 0, leaf_tests = 0;
  constexpr double t_err = 0.00001;  // MeshConservativeAdvancementTraversalNode(): t_err

  while (true) {
    if (sp == 0 && q >= 0) {
      // ---- one distance traversal has finished: advance or stop (conservative_advancement_func_matrix-inl.h:195-210)
      bool done = false;
      if (delta_t <= t_err) {
        done = true;
      } else {
        toc += delta_t;
        if (toc > 1) {
          toc = 1;
          done = true;
        }
      }
      if (!done) {
        motion_integrate(mo1, toc);
        motion_integrate(mo2, toc);
      } else {
        const bool hit = toc < 1;
        if (P.is_collide) P.is_collide[q] = hit ? 1 : 0;
        if (P.toc) P.toc[q] = toc;
        if (P.iterations) P.iterations[q] = iters;
        if (kStats) {
          if (P.n_bv) P.n_bv[q] = bv_tests;
          if (P.n_leaf) P.n_leaf[q] = leaf_tests;
        }
        if (hit) {  // continuous_collision-inl.h:339-350
          motion_integrate(mo1, toc);
          motion_integrate(mo2, toc);
          store_pose(P.contact_tf1, q, mo1.R, mo1.t);
          store_pose(P.contact_tf2, q, mo2.R, mo2.t);
        }
        q = -1;
      }
    }
    const bool need = (q < 0) && !exhausted;
    const long long nq = fetch_work(need, P.work_counter);
    if (need) {
      if (nq < P.n) {
        q = nq;
        const PoseRT b1 = load_pose(P.tf1_beg, q), e1 = load_pose(P.tf1_end, q);
        const PoseRT b2 = load_pose(P.tf2_beg, q), e2 = load_pose(P.tf2_end, q);
        motion_init(mo1, b1, e1);
        motion_init(mo2, b2, e2);
        // without a contact the result keeps the start poses (the reference leaves contact_tf1 / 2 unset)
        store_pose(P.contact_tf1, q, b1.R, b1.t);
        store_pose(P.contact_tf2, q, b2.R, b2.t);
        toc = 0;
        iters = 0;
        bv_tests = leaf_tests = 0;
        if (P.start_hits[q] > 0) {
          // in collision at the start configuration: toc = 0 (:165-170), contact poses = integrate(0)
          motion_integrate(mo1, 0.0);
          motion_integrate(mo2, 0.0);
          if (P.is_collide) P.is_collide[q] = 1;
          if (P.toc) P.toc[q] = 0.0;
          if (P.iterations) P.iterations[q] = 0;
          if (kStats) {
            if (P.n_bv) P.n_bv[q] = 0;
            if (P.n_leaf) P.n_leaf[q] = 0;
          }
          store_pose(P.contact_tf1, q, mo1.R, mo1.t);
          store_pose(P.contact_tf2, q, mo2.R, mo2.t);
          q = -1;
        }
      } else {
        exhausted = true;
      }
    }
    if (__all_sync(0xffffffffu, exhausted && q < 0)) break;
    if (q < 0) continue;

    if (sp == 0) {
      // ---- start a traversal from the current transforms: tf = tf1.inverse(Isometry) * tf2 (:182-193)
      R = mulTM(mo1.R, mo2.R);
      const V3 it = mulTv(mo1.R, mo1.t);
      T = mulTv(mo1.R, mo2.t) + mk(-it.x, -it.y, -it.z);
      merge_detail::quat_of_matrix(mo1.R.m, R0);
      delta_t = 1;
      min_d = 1.7976931348623157e308;
      ++iters;
      stk[0] = make_uint2(0u, 0u);
      stk_d[0] = -1.0;  // the root pair is never bound-tested
      sp = 1;
    }

    // ---- one BVTT node of distanceRecurse (traversal_recurse-inl.h:259-316)
    --sp;
    const uint2 e = stk[sp];
    if (stk_d[sp] >= min_d) continue;  // canStop(c); its motion bound is 0 for OBBRSS: delta_t unchanged
    const int b1 = (int)e.x, b2 = (int)e.y;
    const int fc1 = __ldg(P.m1.first_child + b1);
    const int fc2 = __ldg(P.m2.first_child + b2);
    const bool l1 = fc1 < 0, l2 = fc2 < 0;
    if (l1 && l2) {
      // meshConservativeAdvancementOrientedNodeLeafTesting (-inl.h:641-713)
      if (kStats) leaf_tests++;
      V3 S[3], Tt[3];
      load_tri(P.m1.tri, -(fc1 + 1), S);
      load_tri(P.m2.tri, -(fc2 + 1), Tt);
#pragma unroll
      for (int k = 0; k < 3; ++k) Tt[k] = mulv(R, Tt[k]) + T;
      V3 P1, P2;
      const double d = tri_distance(S, Tt, P1, P2);
      if (d < min_d) min_d = d;
      const V3 n = P2 - P1;
      // R0 * n (Eigen _transformVector): uv = 2 (vec x n); n + w uv + vec x uv
      const V3 qv = mk(R0[0], R0[1], R0[2]);
      V3 uv = cross(qv, n);
      uv = uv + uv;
      V3 nt = (n + uv * R0[3]) + cross(qv, uv);
      {
        const double z = dot(nt, nt);
        if (z > 0) {
          const double len = sqrt(z);
          nt = mk(nt.x / len, nt.y / len, nt.z / len);
        }
      }
      const double bound1 = dot(mo1.range, nt);                         // TranslationMotion: velocity . n
      const double bound2 = dot(mo2.range, mk(-nt.x, -nt.y, -nt.z));
      const double bound = bound1 + bound2;
      const double cur = (bound <= d) ? 1.0 : d / bound;
      if (cur < delta_t) delta_t = cur;
      continue;
    }
    const double size1 = __ldg(P.m1.rss + (size_t)b1 * kNodeDoubles + 15);
    const double size2 = __ldg(P.m2.rss + (size_t)b2 * kNodeDoubles + 15);
    int a1, a2, c1, c2;
    if (l2 || (!l1 && (size1 > size2))) {  // firstOverSecond
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
      delta_t = 0;  // ends the query
      continue;
    }
    if (d2 < d1) {  // the nearer child is visited first: it goes on top
      stk[sp] = make_uint2((unsigned)a1, (unsigned)a2); stk_d[sp] = d1; ++sp;
      stk[sp] = make_uint2((unsigned)c1, (unsigned)c2); stk_d[sp] = d2; ++sp;
    } else {
      stk[sp] = make_uint2((unsigned)c1, (unsigned)c2); stk_d[sp] = d2; ++sp;
      stk[sp] = make_uint2((unsigned)a1, (unsigned)a2); stk_d[sp] = d1; ++sp;
    }
  }
}

}  // namespace fclgpu
