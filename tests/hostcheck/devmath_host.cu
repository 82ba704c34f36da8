// Test-only: compiles the product's device math header for the HOST so that the CPU test
// suite can compare it bit-for-bit with the oracle without a GPU.  Not part of the product.
#include "../../fcl_b200/csrc/device_math.cuh"
#include "../../fcl_b200/csrc/bounds_f32.cuh"
#include "../../fcl_b200/csrc/mesh_sphere.cuh"
#include "../../fcl_b200/csrc/records.hpp"
using namespace fclgpu;

static M3 m3(const double* r) { M3 M; for (int i = 0; i < 9; ++i) M.m[i] = r[i]; return M; }
static V3 v3(const double* v) { return mk(v[0], v[1], v[2]); }

extern "C" {
int hm_obb_disjoint(const double* B9, const double* T3, const double* a3, const double* b3) {
  return obb_disjoint(m3(B9), v3(T3), v3(a3), v3(b3)) ? 1 : 0;
}
double hm_rect_distance(const double* R9, const double* T3, const double* a2, const double* b2) {
  return rect_distance(m3(R9), v3(T3), a2, b2);
}
// n pairs: nodes given as arrays; idx1/idx2 select nodes; R0/T0 per pair (12 doubles)
void hm_obb_pairs(long long n, const double* pose12, const int* idx1, const int* idx2,
                  const double* axis1, const double* To1, const double* ext1,
                  const double* axis2, const double* To2, const double* ext2, int* out) {
  for (long long k = 0; k < n; ++k) {
    int i = idx1[k], j = idx2[k];
    out[k] = obb_pair_disjoint(m3(pose12 + 12 * k), v3(pose12 + 12 * k + 9), m3(axis1 + 9 * i), v3(To1 + 3 * i),
                               v3(ext1 + 3 * i), m3(axis2 + 9 * j), v3(To2 + 3 * j), v3(ext2 + 3 * j));
  }
}
void hm_rss_pairs(long long n, const double* pose12, const int* idx1, const int* idx2,
                  const double* axis1, const double* To1, const double* l1, const double* r1,
                  const double* axis2, const double* To2, const double* l2, const double* r2, double* out) {
  for (long long k = 0; k < n; ++k) {
    int i = idx1[k], j = idx2[k];
    out[k] = rss_pair_distance(m3(pose12 + 12 * k), v3(pose12 + 12 * k + 9), m3(axis1 + 9 * i), v3(To1 + 3 * i),
                               l1 + 2 * i, r1[i], m3(axis2 + 9 * j), v3(To2 + 3 * j), l2 + 2 * j, r2[j]);
  }
}
// triangles: P9/Q9 per pair, pose per pair (Q' = R Q + T done here like the kernels do)
int hm_tri_intersect(const double* P9, const double* Q9, const double* pose12, int want, unsigned* nc,
                     double* contacts6, double* depth, double* normal3) {
  M3 R = m3(pose12); V3 T = v3(pose12 + 9);
  V3 P[3] = {v3(P9), v3(P9 + 3), v3(P9 + 6)};
  V3 Q[3] = {mulv(R, v3(Q9)) + T, mulv(R, v3(Q9 + 3)) + T, mulv(R, v3(Q9 + 6)) + T};
  bool hit = tri_intersect(P[0], P[1], P[2], Q[0], Q[1], Q[2]);
  if (hit != tri_intersect_rolled(P[0], P[1], P[2], Q[0], Q[1], Q[2])) return -1;  // the rolled variant must agree
  if (hit && want) {
    V3 c[2]; unsigned n; double d; V3 nrm;
    tri_contact_info(P, Q, c, n, d, nrm);
    *nc = n; *depth = d;
    contacts6[0] = c[0].x; contacts6[1] = c[0].y; contacts6[2] = c[0].z;
    contacts6[3] = c[1].x; contacts6[4] = c[1].y; contacts6[5] = c[1].z;
    normal3[0] = nrm.x; normal3[1] = nrm.y; normal3[2] = nrm.z;
  }
  return hit ? 1 : 0;
}
double hm_tri_distance(const double* S9, const double* T9, double* P3, double* Q3) {
  V3 S[3] = {v3(S9), v3(S9 + 3), v3(S9 + 6)};
  V3 T[3] = {v3(T9), v3(T9 + 3), v3(T9 + 6)};
  V3 P = mk(0, 0, 0), Q = mk(0, 0, 0);
  double d = tri_distance(S, T, P, Q);
  P3[0] = P.x; P3[1] = P.y; P3[2] = P.z; Q3[0] = Q.x; Q3[1] = Q.y; Q3[2] = Q.z;
  return d;
}
// halfspace (kind 0) / plane (kind 1) {n3, d}, already in the frame of the triangle T9; out7 = contact point, depth, normal
int hm_plane_tri_intersect(int kind, const double* n3, double d, const double* T9, double* out7) {
  V3 T[3] = {v3(T9), v3(T9 + 3), v3(T9 + 6)};
  V3 cp = mk(0, 0, 0), nrm = mk(0, 0, 0);
  double depth = 0;
  const bool hit = kind == 0 ? halfspace_tri_intersect(v3(n3), d, T, cp, depth, nrm) : plane_tri_intersect(v3(n3), d, T, cp, depth, nrm);
  if (hit) {
    out7[0] = cp.x; out7[1] = cp.y; out7[2] = cp.z; out7[3] = depth; out7[4] = nrm.x; out7[5] = nrm.y; out7[6] = nrm.z;
  }
  return hit ? 1 : 0;
}
// sphere (centre c3, radius) vs triangle T9, one frame; out7 = contact point, depth, normal (as the routine writes them)
int hm_sphere_tri_intersect(const double* c3, double radius, const double* T9, double* out7) {
  V3 T[3] = {v3(T9), v3(T9 + 3), v3(T9 + 6)};
  V3 cp = mk(0, 0, 0), nrm = mk(0, 0, 0);
  double depth = 0;
  const bool hit = sphere_tri_intersect(v3(c3), radius, T, cp, depth, nrm);
  if (hit) {
    out7[0] = cp.x; out7[1] = cp.y; out7[2] = cp.z; out7[3] = depth; out7[4] = nrm.x; out7[5] = nrm.y; out7[6] = nrm.z;
  }
  return hit ? 1 : 0;
}
// out7 = distance, point on the sphere, point on the triangle; 0 = centre within the radius
int hm_sphere_tri_distance(const double* c3, double radius, const double* T9, double* out7) {
  V3 T[3] = {v3(T9), v3(T9 + 3), v3(T9 + 6)};
  V3 ps = mk(0, 0, 0), pt = mk(0, 0, 0);
  double d = 0;
  const bool ok = sphere_tri_distance(v3(c3), radius, T, d, ps, pt);
  if (ok) {
    out7[0] = d; out7[1] = ps.x; out7[2] = ps.y; out7[3] = ps.z; out7[4] = pt.x; out7[5] = pt.y; out7[6] = pt.z;
  }
  return ok ? 1 : 0;
}
// The product's mesh <-> sphere distance traversal (mesh_sphere.cuh, the code the kernel inlines) over host arrays:
// first_child[n_nodes], axis 9 / obb_To 3 / obb_ext 3 per node, tri9 = de-indexed triangles.  Outputs as the
// kernel writes them (p1 mesh frame, p2 sphere frame, NaN points and -1 when the centre is within the radius).
struct HostMeshAccessor {
  const int* fc;
  const double *axis, *To, *ext, *tri9;
  int first_child(int b) const { return fc[b]; }
  void box(int b, M3& A, V3& T, double& e0, double& e1, double& e2) const {
    A = m3(axis + 9 * b);
    T = v3(To + 3 * b);
    e0 = ext[3 * b];
    e1 = ext[3 * b + 1];
    e2 = ext[3 * b + 2];
  }
  void box32(int b, ObbRec32& n) const { pack_obb32(axis + 9 * b, To + 3 * b, ext + 3 * b, n); }  // as the upload step packs it
  void tri(int id, V3 T[3]) const {
    for (int k = 0; k < 3; ++k) T[k] = v3(tri9 + 9 * id + 3 * k);
  }
};
int hm_mesh_sphere_distance(long long n, const double* tf1, const double* tf2, double radius, const int* first_child,
                            const double* axis, const double* obb_To, const double* obb_ext, const double* tri9,
                            double* dist, double* p1, double* p2, int* b1, unsigned* n_bv, unsigned* n_leaf, int bound32) {
  const HostMeshAccessor acc{first_child, axis, obb_To, obb_ext, tri9};
  int stk[128];
  float lb[128];
  int overflow = 0;
  for (long long q = 0; q < n; ++q) {
    const M3 R1 = m3(tf1 + 12 * q), R2 = m3(tf2 + 12 * q);
    const V3 t1 = v3(tf1 + 12 * q + 9), t2 = v3(tf2 + 12 * q + 9);
    MeshSphereDistance s;
    mesh_sphere_distance_query(acc, R1, t1, t2, radius, stk, lb, 128, s, bound32 != 0);
    overflow |= s.overflow;
    dist[q] = s.min_d;
    b1[q] = s.best;
    n_bv[q] = s.bv_tests;
    n_leaf[q] = s.leaf_tests;
    V3 a, b;
    if (s.min_d < 0.0) {
      a = b = mk(NAN, NAN, NAN);
    } else {
      a = inverse_apply(R1, t1, s.on_tri);
      b = inverse_apply(R2, t2, s.on_sph);
    }
    p1[3 * q] = a.x; p1[3 * q + 1] = a.y; p1[3 * q + 2] = a.z;
    p2[3 * q] = b.x; p2[3 * q + 1] = b.y; p2[3 * q + 2] = b.z;
  }
  return overflow;
}
// conservative FP32 RSS lower bound on n node pairs (records packed exactly like the upload step does)
void hm_rss_lb32_pairs(long long n, const double* pose12, const int* idx1, const int* idx2,
                       const double* axis1, const double* To1, const double* l1, const double* r1,
                       const double* axis2, const double* To2, const double* l2, const double* r2, float* out) {
  for (long long k = 0; k < n; ++k) {
    int i = idx1[k], j = idx2[k];
    RssRec32 a, b;
    pack_rss32(axis1 + 9 * i, To1 + 3 * i, l1 + 2 * i, r1[i], a);
    pack_rss32(axis2 + 9 * j, To2 + 3 * j, l2 + 2 * j, r2[j], b);
    float R0[9], T0[3], t1;
    pack_pose32(pose12 + 12 * k, pose12 + 12 * k + 9, R0, T0, t1);
    out[k] = rss_lower_bound_f32(R0, T0, t1, a, b);
  }
}
void hm_obb_disjoint32_pairs(long long n, const double* pose12, const int* idx1, const int* idx2,
                             const double* axis1, const double* To1, const double* ext1,
                             const double* axis2, const double* To2, const double* ext2, int* out) {
  for (long long k = 0; k < n; ++k) {
    int i = idx1[k], j = idx2[k];
    ObbRec32 a, b;
    pack_obb32(axis1 + 9 * i, To1 + 3 * i, ext1 + 3 * i, a);
    pack_obb32(axis2 + 9 * j, To2 + 3 * j, ext2 + 3 * j, b);
    float R0[9], T0[3], t1;
    pack_pose32(pose12 + 12 * k, pose12 + 12 * k + 9, R0, T0, t1);
    out[k] = obb_certainly_disjoint_f32(R0, T0, t1, a, b) ? 1 : 0;
  }
}
// triangle-level FP32 lower bound; S9 / T9 in a common frame (FP64); translation by -S0 and rounding done here
float hm_tri_lb32(const double* S9, const double* T9) {
  float s1[3], s2[3], t0[3], t1[3], t2[3];
  for (int c = 0; c < 3; ++c) {
    s1[c] = (float)(S9[3 + c] - S9[c]);
    s2[c] = (float)(S9[6 + c] - S9[c]);
    t0[c] = (float)(T9[c] - S9[c]);
    t1[c] = (float)(T9[3 + c] - S9[c]);
    t2[c] = (float)(T9[6 + c] - S9[c]);
  }
  return tri_lower_bound_f32(s1, s2, t0, t1, t2);
}
// the same with a direction subset (the distance kernel's screening round uses subset 9, see traversal.cuh)
float hm_tri_lb32_dirs(const double* S9, const double* T9, int dirs) {
  float s1[3], s2[3], t0[3], t1[3], t2[3];
  for (int c = 0; c < 3; ++c) {
    s1[c] = (float)(S9[3 + c] - S9[c]);
    s2[c] = (float)(S9[6 + c] - S9[c]);
    t0[c] = (float)(T9[c] - S9[c]);
    t1[c] = (float)(T9[3 + c] - S9[c]);
    t2[c] = (float)(T9[6 + c] - S9[c]);
  }
  switch (dirs) {
    case 9: return tri_lower_bound_dirs_f32<9>(s1, s2, t0, t1, t2);
    case 11: return tri_lower_bound_dirs_f32<11>(s1, s2, t0, t1, t2);
    case 3: return tri_lower_bound_dirs_f32<3>(s1, s2, t0, t1, t2);
    default: return tri_lower_bound_dirs_f32<15>(s1, s2, t0, t1, t2);
  }
}
// FP32 triangle-pair classification exactly as the collide kernel does it: Q' = R Q + T and the
// translation by -P1 in FP64, one rounding to float, then tri_classify_f32
int hm_tri_classify32(const double* P9, const double* Q9, const double* pose12) {
  M3 R = m3(pose12); V3 T = v3(pose12 + 9);
  V3 P[3] = {v3(P9), v3(P9 + 3), v3(P9 + 6)};
  V3 Q[3] = {mulv(R, v3(Q9)) + T, mulv(R, v3(Q9 + 3)) + T, mulv(R, v3(Q9 + 6)) + T};
  const V3 a = P[1] - P[0], b = P[2] - P[0], c = Q[0] - P[0], d = Q[1] - P[0], e = Q[2] - P[0];
  float p2[3] = {(float)a.x, (float)a.y, (float)a.z}, p3[3] = {(float)b.x, (float)b.y, (float)b.z};
  float q1[3] = {(float)c.x, (float)c.y, (float)c.z}, q2[3] = {(float)d.x, (float)d.y, (float)d.z};
  float q3[3] = {(float)e.x, (float)e.y, (float)e.z};
  return tri_classify_f32(p2, p3, q1, q2, q3);
}
}
