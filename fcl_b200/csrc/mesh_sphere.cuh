// mesh_sphere.cuh — per-query traversal of the mesh <-> sphere distance, written once for
// device and host: the kernel (traversal.cuh, distance_mesh_sphere_kernel) instantiates it with an accessor
// over the HBM records, tests/hostcheck instantiates the SAME code over host arrays so the CPU suite can
// compare it with the oracle without a GPU.
//
// Reference: BVHShapeDistancer<OBBRSS, Sphere> -> orientedBVHShapeDistance (narrowphase/detail/
// distance_func_matrix-inl.h:259-277) -> preprocess (triangle 0), distanceRecurse over the mesh tree with the
// sphere as a single leaf (traversal/traversal_recurse-inl.h:259-316), empty postprocess
// (traversal/distance/mesh_shape_distance_traversal_node-inl.h:351-364).
#pragma once
#include <cstdint>

#include "bounds_f32.cuh"
#include "device_math.cuh"

namespace fclgpu {

// largest float <= x (a lower bound stays a lower bound)
FD float float_below(double x) {
#ifdef __CUDA_ARCH__
  return __double2float_rd(x);
#else
  float f = (float)x;
  if ((double)f > x) f = nextafterf(f, -3.0e38f);
  return f;
#endif
}

// Lower bound on (distance from the sphere to anything inside the node's OBB): exact distance from the centre
// cm (mesh frame) to the box, minus the radius, minus a margin far above the rounding error of the box fit,
// of cm and of the leaf arithmetic (which runs in the world frame): 2e-7 relative + 1e-9 of the coordinates'
// magnitude, against rounding errors of ~1e-15 of that magnitude.
// The reference bounds a node by the RSS distance between the node's RSS and an RSS fitted around the sphere's
// 12 bound vertices; a bound only decides which triangles get tested, so any valid lower bound yields the
// minimum over all triangles no valid bound excludes.
// The square root runs in single precision (correctly rounded on both host and device, so the two agree bit for
// bit): g = sqrtf(rd(gap^2)) <= sqrt(gap^2) (1 + 2^-24), and g * 0.9999998 <= gap.
FD double sphere_box_lower_bound(const M3& axis, const V3& To, double e0, double e1, double e2, const V3& cm,
                                 double cm_l1, double radius) {
  const V3 l = mulTv(axis, cm - To);
  const double ex = fmax(fabs(l.x) - e0, 0.0), ey = fmax(fabs(l.y) - e1, 0.0), ez = fmax(fabs(l.z) - e2, 0.0);
  const double scale = (((fabs(To.x) + fabs(To.y)) + fabs(To.z)) + ((e0 + e1) + e2)) + cm_l1;
  const float g = sqrtf(float_below((ex * ex + ey * ey) + ez * ez));
  return ((double)g * 0.9999998 - radius) - 1e-9 * scale;
}

// The same bound from the 64-byte single-precision OBB record (half the bytes per box test; the kernel is bound by
// the L1 data pipe, profiles/r01_ncu_sphere_distance.txt).  u = 2^-24, M = |c|_1 + |e|_1 + |cm|_1 (rounded up):
// inputs rounded once (<= uM each), centre difference <= 3uM, box-frame coordinates <= 25uM, per-axis excess <= 27uM
// (extents are stored rounded up), gap <= 47uM + 6uM for the squares and the square root.  Slack used: 128uM.
struct SphereCentre32 {
  float c[3];      // centre in the mesh frame
  float l1;        // |c|_1 rounded up
  float radius;    // rounded up
};
FD float float_above(double x) {
#ifdef __CUDA_ARCH__
  return __double2float_ru(x);
#else
  float f = (float)x;
  if ((double)f < x) f = nextafterf(f, 3.0e38f);
  return f;
#endif
}
FD SphereCentre32 make_centre32(const V3& cm, double cm_l1, double radius) {
  SphereCentre32 q;
  q.c[0] = (float)cm.x;
  q.c[1] = (float)cm.y;
  q.c[2] = (float)cm.z;
  q.l1 = float_above(cm_l1);
  q.radius = float_above(radius);
  return q;
}
FD float sphere_box_lower_bound_f32(const ObbRec32& n, const SphereCentre32& q) {
  const float dx = q.c[0] - n.c[0], dy = q.c[1] - n.c[1], dz = q.c[2] - n.c[2];
  const float lx = (n.a[0] * dx + n.a[3] * dy) + n.a[6] * dz;
  const float ly = (n.a[1] * dx + n.a[4] * dy) + n.a[7] * dz;
  const float lz = (n.a[2] * dx + n.a[5] * dy) + n.a[8] * dz;
  const float ex = fmaxf(fabsf(lx) - n.e[0], 0.0f), ey = fmaxf(fabsf(ly) - n.e[1], 0.0f), ez = fmaxf(fabsf(lz) - n.e[2], 0.0f);
  const float g = sqrtf((ex * ex + ey * ey) + ez * ez);
  const float M = n.s + q.l1;
  return (g * 0.9999995f - q.radius) - 7.62939453125e-6f * M * 1.000001f;
}

struct MeshSphereDistance {
  double min_d;
  int best;          // closest triangle (or a triangle within the radius when min_d = -1)
  V3 on_tri;         // world frame
  V3 on_sph;         // world frame
  uint32_t bv_tests, leaf_tests;
  bool overflow;     // traversal stack too small (caller reports FCLGPU_ERR_STACK_OVERFLOW)
};

// Leaf: sphere_tri_distance on the triangle moved to the world by tf1 = (R1, t1), like the reference's
// transformed shapeTriangleDistance (sphere_triangle-inl.h:499-508).  Centre within the radius of the triangle
// (the reference's solver returns false and its leaf reads an uninitialised distance): DEFINED as -1.
#pragma nv_exec_check_disable
template <class Acc>
FD void mesh_sphere_leaf(const Acc& acc, int id, const M3& R1, const V3& t1, const V3& c, double radius,
                         MeshSphereDistance& s) {
  V3 T[3];
  acc.tri(id, T);
#pragma unroll
  for (int k = 0; k < 3; ++k) T[k] = mulv(R1, T[k]) + t1;
  double d;
  V3 ps, pt;
  if (sphere_tri_distance(c, radius, T, d, ps, pt)) {
    if (s.min_d > d) {  // DistanceResult::update keeps strictly smaller (distance_result-inl.h:66-103)
      s.min_d = d;
      s.best = id;
      s.on_tri = pt;
      s.on_sph = ps;
    }
  } else if (s.min_d > -1.0) {
    s.min_d = -1.0;
    s.best = id;
  }
}

// Depth first, nearer child first.  Acc: int first_child(b); void box(b, axis, To, e0, e1, e2); void tri(id, T[3]).
// The bounds (distance of the centre to a box, minus the radius, minus a margin) go down to about -radius, so after a
// triangle within the radius (minimum -1: nothing can update it any more) the loop is left explicitly instead of
// descending every box within radius - 1 of the centre.
#pragma nv_exec_check_disable
template <class Acc>
FD void mesh_sphere_distance_query(const Acc& acc, const M3& R1, const V3& t1, const V3& c, double radius, int* stk,
                                   float* stk_lb, int cap, MeshSphereDistance& s, bool bound32 = false) {
  const V3 cm = mulTv(R1, c - t1);  // centre in the mesh frame (bounds only)
  const double cm_l1 = (fabs(cm.x) + fabs(cm.y)) + fabs(cm.z);
  const SphereCentre32 q32 = make_centre32(cm, cm_l1, radius);
  s.min_d = 1.7976931348623157e308;
  s.best = -1;
  s.on_tri = s.on_sph = mk(0, 0, 0);
  s.bv_tests = s.leaf_tests = 0;
  s.overflow = false;
  // bottom: the root (never bound-tested); on top of it, as pseudo entry -1, the reference's preprocess step:
  // seed the minimum with triangle 0 (distancePreprocessOrientedNode, :205-236) before the tree is entered
  stk[0] = 0;
  stk[1] = -1;
  stk_lb[0] = stk_lb[1] = -3.0e38f;
  int sp = 2;
  while (sp > 0) {
    --sp;
    const int b = stk[sp];
    if ((double)stk_lb[sp] >= s.min_d) continue;  // canStop(c), rel_err = abs_err = 0
    const int fc = (b < 0) ? -1 : acc.first_child(b);
    if (fc < 0) {
      if (b >= 0) s.leaf_tests++;
      mesh_sphere_leaf(acc, -(fc + 1), R1, t1, c, radius, s);
      if (s.min_d == -1.0) break;  // a triangle within the radius: the result is final
      continue;
    }
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
    s.bv_tests += 2;
    if (sp + 2 > cap) {
      s.overflow = true;
      return;
    }
    // nearer child on top
    const bool second_first = d2 < d1;
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

// tf.inverse(Isometry) * p = R^T p + (-(R^T t))
FD V3 inverse_apply(const M3& R, const V3& t, const V3& p) {
  const V3 it = mulTv(R, t);
  return mulTv(R, p) + mk(-it.x, -it.y, -it.z);
}

}  // namespace fclgpu
