// oracle_capi.cpp — CPU ORACLE (test infrastructure, NOT product code).
// extern "C" surface over the oracle so that tests/, smoke() and bench.py's
// cpu_baseline leg can drive it through ctypes.  Nothing under fcl_b200/ links
// or loads this library.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>
#include <thread>

#include "fcl_oracle.hpp"

using namespace oracle;

namespace {

// pose record at the boundary: 12 doubles, R row-major (9) then t (3)
inline Pose pose_from(const double* p) {
  Pose q;
  if (!p) {
    q.R = Mat3{{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}};
    q.t = Vec3{{0, 0, 0}};
    return q;
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) q.R.m[i][j] = p[3 * i + j];
  q.t = Vec3{{p[9], p[10], p[11]}};
  return q;
}

struct OrcContact {  // 64 bytes, same layout as fclgpu_contact
  int32_t b1, b2;
  double normal[3], pos[3], depth;
};

struct CollideBatch {
  std::vector<int32_t> counts;
  std::vector<std::vector<Contact>> per_pose;
  std::vector<long long> n_bv, n_leaf;
  double seconds = 0;
};

template <class F>
void parallel_for(long long n, int nthreads, F f) {
  if (nthreads <= 1 || n < 2) {
    for (long long i = 0; i < n; ++i) f(i);
    return;
  }
  // dynamic chunks: per-pose cost varies by >100x
  std::atomic<long long> next{0};
  const long long chunk = std::max<long long>(1, std::min<long long>(256, n / (nthreads * 8)));
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t)
    th.emplace_back([&] {
      while (true) {
        long long s = next.fetch_add(chunk);
        if (s >= n) break;
        long long e = std::min(n, s + chunk);
        for (long long i = s; i < e; ++i) f(i);
      }
    });
  for (auto& x : th) x.join();
}

}  // namespace

extern "C" {

void* orc_model_from_obj(const char* path, int split) {
  std::vector<Vec3> pts;
  std::vector<Tri> tris;
  if (!load_obj(path, pts, tris)) return nullptr;
  Model* m = new Model;
  build_model(*m, pts, tris, (SplitMethod)split);
  return m;
}

void* orc_model_from_arrays(const double* verts, int nv, const int32_t* tris, int nt, int split) {
  std::vector<Vec3> pts(nv);
  std::vector<Tri> ts(nt);
  for (int i = 0; i < nv; ++i) pts[i] = Vec3{{verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]}};
  for (int i = 0; i < nt; ++i) ts[i] = Tri{{tris[3 * i], tris[3 * i + 1], tris[3 * i + 2]}};
  Model* m = new Model;
  build_model(*m, pts, ts, (SplitMethod)split);
  return m;
}

void orc_model_free(void* h) { delete (Model*)h; }

// endReplaceModel(refit=true, bottomup=false) with nv new vertex positions
int orc_model_refit_topdown(void* h, const double* verts, int nv) {
  Model* m = (Model*)h;
  if ((size_t)nv != m->verts.size()) return -7;  // BVH_ERR_INCORRECT_DATA (BVH_model-inl.h:602-606)
  std::vector<Vec3> pts(nv);
  for (int i = 0; i < nv; ++i) pts[i] = Vec3{{verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]}};
  refit_topdown(*m, pts);
  return 0;
}

// endReplaceModel(refit=true, bottomup=true), the reference's default
int orc_model_refit_bottomup(void* h, const double* verts, int nv) {
  Model* m = (Model*)h;
  if ((size_t)nv != m->verts.size()) return -7;
  std::vector<Vec3> pts(nv);
  for (int i = 0; i < nv; ++i) pts[i] = Vec3{{verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]}};
  refit_bottomup(*m, pts);
  return 0;
}

// rss.axis per node, row-major 9 (equal to the obb's unless the model was refitted bottom-up)
void orc_model_get_rss_axis(void* h, double* axis9) {
  Model* m = (Model*)h;
  for (size_t i = 0; i < m->nodes.size(); ++i)
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) axis9[9 * i + 3 * r + c] = m->nodes[i].rss_axis.m[r][c];
}

// OBBRSS::operator+ on two volumes given as arrays (axis9, obb_To3, obb_ext3, rss_axis9, rss_To3, rss_l2, rss_r = 29 doubles)
static Node node_from29(const double* a) {
  Node n{};
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      n.axis.m[r][c] = a[3 * r + c];
      n.rss_axis.m[r][c] = a[15 + 3 * r + c];
    }
  for (int k = 0; k < 3; ++k) {
    n.obb_To[k] = a[9 + k];
    n.obb_ext[k] = a[12 + k];
    n.rss_To[k] = a[24 + k];
  }
  n.rss_l[0] = a[27];
  n.rss_l[1] = a[28];
  return n;
}
static void node_to30(const Node& n, double* a) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      a[3 * r + c] = n.axis.m[r][c];
      a[15 + 3 * r + c] = n.rss_axis.m[r][c];
    }
  for (int k = 0; k < 3; ++k) {
    a[9 + k] = n.obb_To[k];
    a[12 + k] = n.obb_ext[k];
    a[24 + k] = n.rss_To[k];
  }
  a[27] = n.rss_l[0];
  a[28] = n.rss_l[1];
  a[29] = n.rss_r;
}
// volumes as 30 doubles: axis9, obb_To3, obb_ext3, rss_axis9, rss_To3, rss_l2, rss_r
void orc_merge_obbrss(const double* a30, const double* b30, double* out30) {
  Node a = node_from29(a30), b = node_from29(b30), o{};
  a.rss_r = a30[29];
  b.rss_r = b30[29];
  merge_obbrss(a, b, o);
  node_to30(o, out30);
}
void orc_fit3_obbrss(const double* pts9, double* out30) {
  const Vec3 ps[3] = {Vec3{{pts9[0], pts9[1], pts9[2]}}, Vec3{{pts9[3], pts9[4], pts9[5]}}, Vec3{{pts9[6], pts9[7], pts9[8]}}};
  Node o{};
  fit3_obbrss(ps, o);
  node_to30(o, out30);
}

void orc_model_partition(void* h, int32_t* first_primitive, int32_t* num_primitives, int32_t* primitive_indices) {
  Model* m = (Model*)h;
  for (size_t i = 0; i < m->nodes.size(); ++i) {
    first_primitive[i] = m->nodes[i].first_primitive;
    num_primitives[i] = m->nodes[i].num_primitives;
  }
  for (size_t i = 0; i < m->prim.size(); ++i) primitive_indices[i] = (int32_t)m->prim[i];
}

void orc_model_counts(void* h, int* nv, int* nt, int* nn) {
  Model* m = (Model*)h;
  *nv = (int)m->verts.size();
  *nt = (int)m->tris.size();
  *nn = (int)m->nodes.size();
}

// axis is written row-major 9 per node (axis[3*r+c]; column c = c-th box axis)
void orc_model_get(void* h, double* verts, int32_t* tris, int32_t* first_child, double* axis9,
                   double* obb_To, double* obb_ext, double* rss_To, double* rss_l, double* rss_r) {
  Model* m = (Model*)h;
  if (verts)
    for (size_t i = 0; i < m->verts.size(); ++i)
      for (int k = 0; k < 3; ++k) verts[3 * i + k] = m->verts[i][k];
  if (tris)
    for (size_t i = 0; i < m->tris.size(); ++i)
      for (int k = 0; k < 3; ++k) tris[3 * i + k] = m->tris[i].v[k];
  for (size_t i = 0; i < m->nodes.size(); ++i) {
    const Node& n = m->nodes[i];
    if (first_child) first_child[i] = n.first_child;
    if (axis9)
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) axis9[9 * i + 3 * r + c] = n.axis.m[r][c];
    for (int k = 0; k < 3; ++k) {
      if (obb_To) obb_To[3 * i + k] = n.obb_To[k];
      if (obb_ext) obb_ext[3 * i + k] = n.obb_ext[k];
      if (rss_To) rss_To[3 * i + k] = n.rss_To[k];
    }
    if (rss_l) { rss_l[2 * i] = n.rss_l[0]; rss_l[2 * i + 1] = n.rss_l[1]; }
    if (rss_r) rss_r[i] = n.rss_r;
  }
}

// ---- batched queries (tf arrays: n x 12 doubles, or NULL = identity for all) ----
void* orc_collide_batch(void* h1, void* h2, long long n, const double* tf1, const double* tf2,
                        long long num_max_contacts, int enable_contact, int nthreads) {
  Model* m1 = (Model*)h1;
  Model* m2 = (Model*)h2;
  CollideBatch* b = new CollideBatch;
  b->counts.assign(n, 0);
  b->per_pose.resize(n);
  b->n_bv.assign(n, 0);
  b->n_leaf.assign(n, 0);
  auto t0 = std::chrono::steady_clock::now();
  parallel_for(n, nthreads, [&](long long i) {
    Pose a = pose_from(tf1 ? tf1 + 12 * i : nullptr);
    Pose c = pose_from(tf2 ? tf2 + 12 * i : nullptr);
    CollideStats st;
    collide(*m1, a, *m2, c, (size_t)num_max_contacts, enable_contact != 0, b->per_pose[i], &st);
    b->counts[i] = (int32_t)b->per_pose[i].size();
    b->n_bv[i] = st.n_bv;
    b->n_leaf[i] = st.n_leaf;
  });
  b->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return b;
}

// mesh <-> sphere collide batch: tf1 poses of the mesh, tf2 poses of the sphere (only the translation matters
// for the leaf test; the rotation enters the sphere's fitted OBB)
void* orc_collide_mesh_sphere_batch(void* h1, double radius, long long n, const double* tf1, const double* tf2,
                                    long long num_max_contacts, int enable_contact, int nthreads) {
  Model* m1 = (Model*)h1;
  CollideBatch* b = new CollideBatch;
  b->counts.assign(n, 0);
  b->per_pose.resize(n);
  b->n_bv.assign(n, 0);
  b->n_leaf.assign(n, 0);
  auto t0 = std::chrono::steady_clock::now();
  parallel_for(n, nthreads, [&](long long i) {
    Pose a = pose_from(tf1 ? tf1 + 12 * i : nullptr);
    Pose c = pose_from(tf2 ? tf2 + 12 * i : nullptr);
    CollideStats st;
    collide_mesh_sphere(*m1, a, radius, c, (size_t)num_max_contacts, enable_contact != 0, b->per_pose[i], &st);
    b->counts[i] = (int32_t)b->per_pose[i].size();
    b->n_bv[i] = st.n_bv;
    b->n_leaf[i] = st.n_leaf;
  });
  b->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return b;
}

// fcl::collide(mesh, tf1, Halfspace | Plane {n, d}, tf2) over a batch; kind 0 = halfspace, 1 = plane
void* orc_collide_mesh_plane_batch(void* h1, int kind, const double* n3, double d, long long n, const double* tf1, const double* tf2,
                                   long long num_max_contacts, int enable_contact, int nthreads) {
  Model* m1 = (Model*)h1;
  const PlaneShape shape = make_plane(Vec3{{n3[0], n3[1], n3[2]}}, d);
  CollideBatch* b = new CollideBatch;
  b->counts.assign(n, 0);
  b->per_pose.resize(n);
  b->n_bv.assign(n, 0);
  b->n_leaf.assign(n, 0);
  auto t0 = std::chrono::steady_clock::now();
  parallel_for(n, nthreads, [&](long long i) {
    Pose a = pose_from(tf1 ? tf1 + 12 * i : nullptr);
    Pose c = pose_from(tf2 ? tf2 + 12 * i : nullptr);
    CollideStats st;
    collide_mesh_plane(*m1, a, kind, shape, c, (size_t)num_max_contacts, enable_contact != 0, b->per_pose[i], &st);
    b->counts[i] = (int32_t)b->per_pose[i].size();
    b->n_bv[i] = st.n_bv;
    b->n_leaf[i] = st.n_leaf;
  });
  b->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return b;
}

long long orc_brute_mesh_plane(void* h1, int kind, const double* n3, double d, const double* tf1, const double* tf2, int32_t* out, long long cap) {
  std::vector<int> tris;
  brute_mesh_plane(*(Model*)h1, pose_from(tf1), kind, make_plane(Vec3{{n3[0], n3[1], n3[2]}}, d), pose_from(tf2), tris);
  for (size_t i = 0; i < tris.size() && (long long)i < cap; ++i) out[i] = tris[i];
  return (long long)tris.size();
}

// unit kernel: shape {n, d} posed by tf_shape vs triangle (9 doubles) posed by tf_tri; out7 = contact point, depth, normal
int orc_plane_tri_intersect(int kind, const double* n3, double d, const double* tf_shape, const double* tri9, const double* tf_tri, double* out7) {
  const PlaneShape s = make_plane(Vec3{{n3[0], n3[1], n3[2]}}, d);
  Vec3 cp{{0, 0, 0}}, nrm{{0, 0, 0}};
  double pen = 0;
  const Vec3 P1{{tri9[0], tri9[1], tri9[2]}}, P2{{tri9[3], tri9[4], tri9[5]}}, P3{{tri9[6], tri9[7], tri9[8]}};
  const bool hit = kind == 0 ? halfspace_tri_intersect(s, pose_from(tf_shape), P1, P2, P3, pose_from(tf_tri), &cp, &pen, &nrm)
                             : plane_tri_intersect(s, pose_from(tf_shape), P1, P2, P3, pose_from(tf_tri), &cp, &pen, &nrm);
  if (out7 && hit) {
    for (int k = 0; k < 3; ++k) { out7[k] = cp[k]; out7[4 + k] = nrm[k]; }
    out7[3] = pen;
  }
  return hit ? 1 : 0;
}

// ids of all triangles the sphere intersects (brute force, primitive order); returns the count
long long orc_brute_mesh_sphere(void* h1, double radius, const double* tf1, const double* tf2, int32_t* out, long long cap) {
  std::vector<int> tris;
  brute_mesh_sphere(*(Model*)h1, pose_from(tf1), radius, pose_from(tf2), tris);
  for (size_t i = 0; i < tris.size() && (long long)i < cap; ++i) out[i] = tris[i];
  return (long long)tris.size();
}

// unit kernel: sphere (center, radius) vs triangle (9 doubles) in one frame; out7 = contact point, depth, normal
int orc_sphere_tri_intersect(const double* center, double radius, const double* tri9, double* out7) {
  Vec3 cp{{0, 0, 0}}, nrm{{0, 0, 0}};
  double pen = 0;
  const bool hit = sphere_tri_intersect(Vec3{{center[0], center[1], center[2]}}, radius, Vec3{{tri9[0], tri9[1], tri9[2]}}, Vec3{{tri9[3], tri9[4], tri9[5]}}, Vec3{{tri9[6], tri9[7], tri9[8]}}, &cp, &pen, &nrm);
  if (out7 && hit) {
    for (int k = 0; k < 3; ++k) { out7[k] = cp[k]; out7[4 + k] = nrm[k]; }
    out7[3] = pen;
  }
  return hit ? 1 : 0;
}

// mesh <-> sphere distance batch (brute != 0: every triangle in primitive order instead of the traversal);
// p1 in the mesh frame, p2 in the sphere frame, b1 = triangle; returns wall seconds
double orc_distance_mesh_sphere_batch(void* h1, double radius, long long n, const double* tf1, const double* tf2,
                                      int brute, int nthreads, double* dist, double* p1, double* p2, int32_t* b1,
                                      long long* n_bv, long long* n_leaf) {
  Model* m1 = (Model*)h1;
  auto t0 = std::chrono::steady_clock::now();
  parallel_for(n, nthreads, [&](long long i) {
    Pose a = pose_from(tf1 ? tf1 + 12 * i : nullptr);
    Pose c = pose_from(tf2 ? tf2 + 12 * i : nullptr);
    CollideStats st;
    DistanceOut o;
    if (brute) brute_distance_mesh_sphere(*m1, a, radius, c, o);
    else distance_mesh_sphere(*m1, a, radius, c, o, &st);
    if (dist) dist[i] = o.min_distance;
    for (int k = 0; k < 3; ++k) {
      if (p1) p1[3 * i + k] = o.p1[k];
      if (p2) p2[3 * i + k] = o.p2[k];
    }
    if (b1) b1[i] = o.b1;
    if (n_bv) n_bv[i] = st.n_bv;
    if (n_leaf) n_leaf[i] = st.n_leaf;
  });
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// unit kernel: sphereTriangleDistance with nearest points, centre and triangle in one frame;
// out7 = distance, point on the sphere (same frame), point on the triangle; returns 0 when the centre is within the radius
int orc_sphere_tri_distance(const double* center, double radius, const double* tri9, double* out7) {
  double d = 0;
  Vec3 ps{{0, 0, 0}}, pt{{0, 0, 0}};
  const bool ok = sphere_tri_distance(Vec3{{center[0], center[1], center[2]}}, radius, Vec3{{tri9[0], tri9[1], tri9[2]}},
                                      Vec3{{tri9[3], tri9[4], tri9[5]}}, Vec3{{tri9[6], tri9[7], tri9[8]}}, &d, &ps, &pt);
  if (out7 && ok) {
    out7[0] = d;
    for (int k = 0; k < 3; ++k) { out7[1 + k] = ps[k]; out7[4 + k] = pt[k]; }
  }
  return ok ? 1 : 0;
}

// computeBV<OBBRSS>(Sphere(radius), tf): axis9 row-major, obb_To3, obb_ext3, rss_To3, rss_l2, rss_r
void orc_sphere_bv(double radius, const double* tf, double* axis9, double* obb_To, double* obb_ext, double* rss_To,
                   double* rss_l, double* rss_r) {
  Node bv;
  sphere_obb(radius, pose_from(tf), bv);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) axis9[3 * r + c] = bv.axis.m[r][c];
  for (int k = 0; k < 3; ++k) { obb_To[k] = bv.obb_To[k]; obb_ext[k] = bv.obb_ext[k]; rss_To[k] = bv.rss_To[k]; }
  rss_l[0] = bv.rss_l[0];
  rss_l[1] = bv.rss_l[1];
  *rss_r = bv.rss_r;
}

double orc_collide_seconds(void* hb) { return ((CollideBatch*)hb)->seconds; }

long long orc_collide_total(void* hb) {
  CollideBatch* b = (CollideBatch*)hb;
  long long t = 0;
  for (auto c : b->counts) t += c;
  return t;
}

// counts[n], contacts[total] (DFS order per pose, poses concatenated), n_bv[n], n_leaf[n]
void orc_collide_copy(void* hb, int32_t* counts, void* contacts, long long* n_bv, long long* n_leaf) {
  CollideBatch* b = (CollideBatch*)hb;
  OrcContact* out = (OrcContact*)contacts;
  size_t k = 0;
  for (size_t i = 0; i < b->counts.size(); ++i) {
    if (counts) counts[i] = b->counts[i];
    if (n_bv) n_bv[i] = b->n_bv[i];
    if (n_leaf) n_leaf[i] = b->n_leaf[i];
    if (out)
      for (const Contact& c : b->per_pose[i]) {
        OrcContact& o = out[k++];
        o.b1 = c.b1;
        o.b2 = c.b2;
        for (int d = 0; d < 3; ++d) {
          o.normal[d] = c.normal[d];
          o.pos[d] = c.pos[d];
        }
        o.depth = c.depth;
      }
  }
}

void orc_collide_free(void* hb) { delete (CollideBatch*)hb; }

// returns wall seconds of the batch
double orc_distance_batch(void* h1, void* h2, long long n, const double* tf1, const double* tf2,
                          int enable_nearest_points, int qsize, int nthreads, double* dist,
                          double* p1, double* p2, int32_t* b1, int32_t* b2, long long* n_bv,
                          long long* n_leaf) {
  Model* m1 = (Model*)h1;
  Model* m2 = (Model*)h2;
  auto t0 = std::chrono::steady_clock::now();
  parallel_for(n, nthreads, [&](long long i) {
    Pose a = pose_from(tf1 ? tf1 + 12 * i : nullptr);
    Pose c = pose_from(tf2 ? tf2 + 12 * i : nullptr);
    CollideStats st;
    DistanceOut o;
    distance(*m1, a, *m2, c, enable_nearest_points != 0, o, qsize, &st);
    if (dist) dist[i] = o.min_distance;
    for (int k = 0; k < 3; ++k) {
      if (p1) p1[3 * i + k] = o.p1[k];
      if (p2) p2[3 * i + k] = o.p2[k];
    }
    if (b1) b1[i] = o.b1;
    if (b2) b2[i] = o.b2;
    if (n_bv) n_bv[i] = st.n_bv;
    if (n_leaf) n_leaf[i] = st.n_leaf;
  });
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// continuousCollide (CCDM_TRANS, conservative advancement) over a batch; tf*_beg / tf*_end: 12 doubles per query
// (NULL = identity).  Outputs (any may be NULL): is_collide[n], toc[n], contact_tf1/2[12 n], iterations[n].
double orc_continuous_collide_translation_batch(void* h1, void* h2, long long n, const double* tf1_beg, const double* tf1_end,
                                                const double* tf2_beg, const double* tf2_end, int nthreads, int32_t* is_collide,
                                                double* toc, double* contact_tf1, double* contact_tf2, int32_t* iterations) {
  Model* m1 = (Model*)h1;
  Model* m2 = (Model*)h2;
  auto t0 = std::chrono::steady_clock::now();
  auto put = [](double* dst, const Pose& p) {
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) dst[3 * r + c] = p.R.m[r][c];
    for (int k = 0; k < 3; ++k) dst[9 + k] = p.t[k];
  };
  parallel_for(n, nthreads, [&](long long i) {
    const Pose a0 = pose_from(tf1_beg ? tf1_beg + 12 * i : nullptr), a1 = pose_from(tf1_end ? tf1_end + 12 * i : nullptr);
    const Pose b0 = pose_from(tf2_beg ? tf2_beg + 12 * i : nullptr), b1 = pose_from(tf2_end ? tf2_end + 12 * i : nullptr);
    ContinuousOut o;
    continuous_collide_translation(*m1, a0, a1, *m2, b0, b1, o);
    if (is_collide) is_collide[i] = o.is_collide ? 1 : 0;
    if (toc) toc[i] = o.time_of_contact;
    if (contact_tf1) put(contact_tf1 + 12 * i, o.contact_tf1);
    if (contact_tf2) put(contact_tf2 + 12 * i, o.contact_tf2);
    if (iterations) iterations[i] = o.iterations;
  });
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// ---- brute force --------------------------------------------------------------
long long orc_brute_collide(void* h1, void* h2, const double* tf1, const double* tf2, int32_t* pairs,
                            long long cap_pairs) {
  std::vector<std::pair<int, int>> v;
  brute_collide_pairs(*(Model*)h1, pose_from(tf1), *(Model*)h2, pose_from(tf2), v);
  for (long long i = 0; i < (long long)v.size() && i < cap_pairs; ++i) {
    pairs[2 * i] = v[i].first;
    pairs[2 * i + 1] = v[i].second;
  }
  return (long long)v.size();
}

double orc_brute_distance(void* h1, void* h2, const double* tf1, const double* tf2, double* p1,
                          double* p2, int32_t* b12) {
  DistanceOut o;
  brute_distance(*(Model*)h1, pose_from(tf1), *(Model*)h2, pose_from(tf2), o);
  for (int k = 0; k < 3; ++k) {
    if (p1) p1[k] = o.p1[k];
    if (p2) p2[k] = o.p2[k];
  }
  if (b12) { b12[0] = o.b1; b12[1] = o.b2; }
  return o.min_distance;
}

// ---- unit-level kernels (row-major 3x3) ---------------------------------------
static Mat3 mat_from(const double* r) {
  Mat3 M;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) M.m[i][j] = r[3 * i + j];
  return M;
}
static Vec3 vec_from(const double* v) { return Vec3{{v[0], v[1], v[2]}}; }

int orc_obb_disjoint(const double* B9, const double* T3, const double* a3, const double* b3) {
  return obb_disjoint(mat_from(B9), vec_from(T3), vec_from(a3), vec_from(b3)) ? 1 : 0;
}

static Node node_from(const double* axis9, const double* To, const double* ext_or_l, double r, bool rss) {
  Node n{};
  n.axis = mat_from(axis9);
  n.rss_axis = n.axis;
  if (rss) {
    n.rss_To = vec_from(To);
    n.rss_l[0] = ext_or_l[0];
    n.rss_l[1] = ext_or_l[1];
    n.rss_r = r;
  } else {
    n.obb_To = vec_from(To);
    n.obb_ext = vec_from(ext_or_l);
  }
  return n;
}

int orc_obb_overlap(const double* R0, const double* T0, const double* axis1, const double* To1,
                    const double* ext1, const double* axis2, const double* To2, const double* ext2) {
  Node n1 = node_from(axis1, To1, ext1, 0, false), n2 = node_from(axis2, To2, ext2, 0, false);
  return obb_overlap(mat_from(R0), vec_from(T0), n1, n2) ? 1 : 0;
}

double orc_rect_distance(const double* R9, const double* T3, const double* a2, const double* b2) {
  return rect_distance(mat_from(R9), vec_from(T3), a2, b2);
}

double orc_rss_distance(const double* R0, const double* T0, const double* axis1, const double* To1,
                        const double* l1, double r1, const double* axis2, const double* To2,
                        const double* l2, double r2) {
  Node n1 = node_from(axis1, To1, l1, r1, true), n2 = node_from(axis2, To2, l2, r2, true);
  return rss_distance(mat_from(R0), vec_from(T0), n1, n2);
}

// P9/Q9: three vertices each; returns hit; out: n_contacts, contacts(6), depth, normal(3)
int orc_tri_intersect(const double* P9, const double* Q9, const double* R9, const double* T3,
                      int want_contacts, uint32_t* n_contacts, double* contacts6, double* depth,
                      double* normal3) {
  Vec3 P[3] = {vec_from(P9), vec_from(P9 + 3), vec_from(P9 + 6)};
  Vec3 Q[3] = {vec_from(Q9), vec_from(Q9 + 3), vec_from(Q9 + 6)};
  if (!want_contacts)
    return tri_intersect(P, Q, mat_from(R9), vec_from(T3), nullptr, nullptr, nullptr, nullptr) ? 1 : 0;
  Vec3 c[2] = {Vec3{{0, 0, 0}}, Vec3{{0, 0, 0}}}, nrm{{0, 0, 0}};
  unsigned nc = 0;
  double d = 0;
  bool hit = tri_intersect(P, Q, mat_from(R9), vec_from(T3), c, &nc, &d, &nrm);
  if (hit) {
    *n_contacts = nc;
    *depth = d;
    for (int k = 0; k < 3; ++k) {
      contacts6[k] = c[0][k];
      contacts6[3 + k] = c[1][k];
      normal3[k] = nrm[k];
    }
  }
  return hit ? 1 : 0;
}

double orc_tri_distance(const double* S9, const double* T9, double* P3, double* Q3) {
  Vec3 S[3] = {vec_from(S9), vec_from(S9 + 3), vec_from(S9 + 6)};
  Vec3 T[3] = {vec_from(T9), vec_from(T9 + 3), vec_from(T9 + 6)};
  Vec3 P{{0, 0, 0}}, Q{{0, 0, 0}};
  double d = tri_distance(S, T, P, Q);
  for (int k = 0; k < 3; ++k) {
    P3[k] = P[k];
    Q3[k] = Q[k];
  }
  return d;
}

int orc_hardware_threads() { return (int)std::thread::hardware_concurrency(); }


// brute-force broadphase: pairs (2 ints each, i-major) of overlapping world AABBs and, per pair, numContacts of
// fcl::collide(o1, o2, request) with a fresh result; returns the number of pairs (pairs / counts may be NULL or too small:
// only the first `cap` are written)
long long orc_broadphase(int n_geoms, void** geoms, long long n1, const int32_t* geom1, const double* tf1, long long n2,
                         const int32_t* geom2, const double* tf2, long long num_max_contacts, int enable_contact, int nthreads,
                         int32_t* pairs, int32_t* counts, long long cap, double* aabb1_out) {
  std::vector<const Model*> gs;
  for (int g = 0; g < n_geoms; ++g) gs.push_back((const Model*)geoms[g]);
  std::vector<int> g1(geom1, geom1 + n1), g2(geom2, geom2 + n2);
  std::vector<Pose> p1(n1), p2(n2);
  for (long long i = 0; i < n1; ++i) p1[i] = pose_from(tf1 + 12 * i);
  for (long long j = 0; j < n2; ++j) p2[j] = pose_from(tf2 + 12 * j);
  std::vector<std::pair<int, int>> pr;
  broadphase_pairs(gs, g1, p1, g2, p2, pr);
  if (aabb1_out)
    for (long long i = 0; i < n1; ++i) world_aabb(local_aabb(*gs[g1[i]]), p1[i], aabb1_out + 6 * i);
  const long long m = std::min<long long>((long long)pr.size(), cap);
  if (pairs)
    for (long long k = 0; k < m; ++k) {
      pairs[2 * k] = pr[k].first;
      pairs[2 * k + 1] = pr[k].second;
    }
  if (counts)
    parallel_for(m, nthreads, [&](long long k) {
      std::vector<Contact> out;
      counts[k] = (int32_t)collide(*gs[g1[pr[k].first]], p1[pr[k].first], *gs[g2[pr[k].second]], p2[pr[k].second],
                                   (size_t)num_max_contacts, enable_contact != 0, out);
    });
  return (long long)pr.size();
}
}
