// fcl_oracle.hpp — CPU ORACLE (test infrastructure, NOT product code).
//
// A dependency-free, scalar, double-precision restatement of the reference's
// BVHModel<OBBRSS<double>> mesh-mesh collide()/distance() path.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may build, link or call anything in this directory.  The shipped library
// (fcl_b200/) never includes or links it.
//
// PARITY STATUS: the reference cannot be compiled here (it needs Eigen3 and
// libccd, neither is on the image), so this restatement is pinned against the
// reference's own known-answer tests and invariants (see tests/test_oracle_*.py):
//   * test/test_fcl_math.cpp:256-291   RSS distance known answers 2, 1, sqrt(6)-1, 1
//   * test/test_fcl_collision.cpp:792-886  contact-pair sets equal across split
//     methods and equal to brute-force all-pairs triangle intersection
//   * test/test_fcl_distance.cpp:177-298  distance equal across split methods,
//     qsize 2 vs 20, and equal to brute-force all-pairs triangle distance
//   * test/test_fcl_shape_mesh_consistency.cpp:57-80  tessellated spheres
//   * test/test_fcl_collision.cpp:271-311  OBB overlap == AABB overlap
// At the ULP level (Eigen's internal association order of 3-term sums) parity
// is UNPINNED; this file fixes one canonical order, documented at each helper:
// every 3-term sum is evaluated left to right, (a0*b0 + a1*b1) + a2*b2, with
// separately rounded multiplies and adds (compile with -ffp-contract=off,
// matching the reference's default non-FMA x86-64 build, CMakeLists.txt:82-116).
//
// All file:line citations are relative to /root/reference/.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace oracle {

struct Vec3 {
  double v[3];
  double& operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
};

// Row-major storage: m[r][c].  "axis.col(i)" of the reference is (m[0][i], m[1][i], m[2][i]).
struct Mat3 {
  double m[3][3];
};

struct Tri {
  int v[3];
};

// One BVNode<OBBRSS<double>>  (include/fcl/geometry/bvh/BV_node.h:50-72,
// BV_node_base.h:48-63, math/bv/OBB.h:56-68, RSS.h:63-82, OBBRSS.h:52-60).
struct Node {
  int first_child;      // <0 => leaf holding primitive -(first_child+1)
  int first_primitive;  // build bookkeeping
  int num_primitives;
  Mat3 axis;            // obb.axis; also rss.axis after endModel() / a top-down refit (BV_fitter-inl.h:464)
  Mat3 rss_axis;        // rss.axis: differs from `axis` only after a bottom-up refit (OBBRSS-inl.h:95-101 merges separately)
  Vec3 obb_To, obb_ext;
  Vec3 rss_To;
  double rss_l[2], rss_r;
};

enum SplitMethod { SPLIT_MEAN = 0, SPLIT_MEDIAN = 1, SPLIT_BV_CENTER = 2 };

struct Model {
  std::vector<Vec3> verts;
  std::vector<Tri> tris;
  std::vector<Node> nodes;
  std::vector<unsigned> prim;  // BVHModel::primitive_indices after the build (node i owns prim[first_primitive .. +num_primitives))
  SplitMethod split = SPLIT_MEAN;
};

// rigid pose: p_world = R * p + t
struct Pose {
  Mat3 R;
  Vec3 t;
};

struct Contact {
  int b1, b2;
  Vec3 normal, pos;
  double depth;
};

struct CollideStats {
  long long n_bv = 0, n_leaf = 0;
};

struct DistanceOut {
  double min_distance;
  Vec3 p1, p2;  // world frame (valid when nearest points requested)
  int b1, b2;
};

// ---- inputs ----------------------------------------------------------------
bool load_obj(const std::string& path, std::vector<Vec3>& pts, std::vector<Tri>& tris);
void build_model(Model& m, const std::vector<Vec3>& pts, const std::vector<Tri>& tris,
                 SplitMethod split = SPLIT_MEAN);

// endReplaceModel(refit=true, bottomup=false): new vertex positions, same topology; every node is
// refitted over its stored primitive range (BVH_model-inl.h:594-620, refitTree_topdown :1064-1076)
void refit_topdown(Model& m, const std::vector<Vec3>& new_verts);
// endReplaceModel(refit=true, bottomup=true) -- the reference's default: leaves get the closed-form fit of their
// triangle (fit3), inner nodes the merge of their children's volumes (BVH_model-inl.h:952-1037, OBB-inl.h:161-357,
// RSS-inl.h:313-371).  OBB and RSS are merged separately, so the two no longer share axes.
void refit_bottomup(Model& m, const std::vector<Vec3>& new_verts);
// the two merges on their own (tests): out.{axis, obb_To, obb_ext} = a.obb + b.obb, out.{rss_*} = a.rss + b.rss
void merge_obbrss(const Node& a, const Node& b, Node& out);
void fit3_obbrss(const Vec3 ps[3], Node& out);

// continuousCollide() with ccd_motion_type = CCDM_TRANS and ccd_solver_type = CCDC_CONSERVATIVE_ADVANCEMENT on two
// BVHModel<OBBRSS> (only the translations of tf*_end are used: TranslationMotion keeps tf*_beg's rotation)
struct ContinuousOut {
  bool is_collide;
  double time_of_contact;
  Pose contact_tf1, contact_tf2;  // ContinuousCollisionResult::contact_tf1/2 (the start poses when there is no contact)
  int iterations;                 // distance traversals run (diagnostic)
};
double continuous_collide_translation(const Model& m1, const Pose& tf1_beg, const Pose& tf1_end, const Model& m2,
                                      const Pose& tf2_beg, const Pose& tf2_end, ContinuousOut& out);

// ---- BV / leaf kernels -------------------------------------------------------
bool obb_disjoint(const Mat3& B, const Vec3& T, const Vec3& a, const Vec3& b);
bool obb_overlap(const Mat3& R0, const Vec3& T0, const Node& n1, const Node& n2);
double rect_distance(const Mat3& Rab, const Vec3& Tab, const double a[2], const double b[2]);
double rss_distance(const Mat3& R0, const Vec3& T0, const Node& n1, const Node& n2);
bool tri_intersect(const Vec3 P[3], const Vec3 Qin[3], const Mat3& R, const Vec3& T,
                   Vec3* contact_points, unsigned* num_contact_points, double* depth,
                   Vec3* normal);
double tri_distance(const Vec3 S[3], const Vec3 T[3], Vec3& P, Vec3& Q);

// ---- queries -----------------------------------------------------------------
// collide(): contacts appended in the reference's DFS order; returns numContacts.
size_t collide(const Model& m1, const Pose& tf1, const Model& m2, const Pose& tf2,
               size_t num_max_contacts, bool enable_contact, std::vector<Contact>& out,
               CollideStats* stats = nullptr);
// distance(): qsize<=2 recursive, else queue variant.
double distance(const Model& m1, const Pose& tf1, const Model& m2, const Pose& tf2,
                bool enable_nearest_points, DistanceOut& out, int qsize = 2,
                CollideStats* stats = nullptr);

// ---- mesh <-> sphere (closed-form leaf test, no GJK on this pair) ----------------
// sphereTriangleIntersect, narrowphase/detail/primitive_shape_algorithm/sphere_triangle-inl.h:147-244.
// center / P1..P3 in one frame.  Outputs as the reference writes them (normal is negated by the caller).
bool sphere_tri_intersect(const Vec3& center, double radius, const Vec3& P1, const Vec3& P2, const Vec3& P3,
                          Vec3* contact_point, double* penetration_depth, Vec3* normal);
// computeBV<OBBRSS>(Sphere, tf): OBB and RSS fitted over the 12 bound vertices (sphere-inl.h:95-120, fitn)
void sphere_obb(double radius, const Pose& tf, Node& bv);
// fcl::collide(BVHModel<OBBRSS>, tf1, Sphere(radius), tf2): contacts {b1 = triangle, b2 = -1 (Contact::NONE)}
size_t collide_mesh_sphere(const Model& m1, const Pose& tf1, double radius, const Pose& tf2, size_t num_max_contacts,
                           bool enable_contact, std::vector<Contact>& out, CollideStats* stats = nullptr);
// every triangle tested in primitive order (for the invariants): ids of the intersecting triangles
void brute_mesh_sphere(const Model& m1, const Pose& tf1, double radius, const Pose& tf2, std::vector<int>& tris);

// sphereTriangleDistance with nearest points (sphere_triangle-inl.h:469-496) on top of Project::projectTriangle
// (math/detail/project-inl.h:54-123): centre o and triangle in one frame; false = centre within the radius of the
// triangle (the reference then leaves every output unwritten).
bool sphere_tri_distance(const Vec3& o, double radius, const Vec3& P1, const Vec3& P2, const Vec3& P3, double* dist,
                         Vec3* on_sphere_world, Vec3* on_triangle);
// fcl::distance(BVHModel<OBBRSS>, tf1, Sphere(radius), tf2): out.p1 in the MESH frame, out.p2 in the SPHERE frame (the
// reference's postprocess is empty for this node), b2 = -1; centre within the radius: min_distance = -1, NaN points
double distance_mesh_sphere(const Model& m1, const Pose& tf1, double radius, const Pose& tf2, DistanceOut& out,
                            CollideStats* stats = nullptr);
double brute_distance_mesh_sphere(const Model& m1, const Pose& tf1, double radius, const Pose& tf2, DistanceOut& out);

// ---- mesh <-> halfspace / plane (closed-form leaf tests, no GJK on these pairs) ------------------
// Halfspace / Plane { n . x <= d / n . x = d } with the constructor's normalisation (halfspace-inl.h:144-160 unitNormalTest)
struct PlaneShape {
  Vec3 n;
  double d;
};
PlaneShape make_plane(const Vec3& n, double d);
// transform(Halfspace, tf) / transform(Plane, tf): n' = R n, d' = d + n' . t (halfspace-inl.h:168-180, plane-inl.h:168-180)
PlaneShape transform_plane(const PlaneShape& a, const Pose& tf);
// halfspaceTriangleIntersect (narrowphase/detail/primitive_shape_algorithm/halfspace-inl.h:587-621) and
// planeTriangleIntersect (.../plane-inl.h:683-759): shape posed by tf1, triangle posed by tf2
bool halfspace_tri_intersect(const PlaneShape& s, const Pose& tf1, const Vec3& P1, const Vec3& P2, const Vec3& P3, const Pose& tf2,
                             Vec3* contact_point, double* penetration_depth, Vec3* normal);
bool plane_tri_intersect(const PlaneShape& s, const Pose& tf1, const Vec3& P1, const Vec3& P2, const Vec3& P3, const Pose& tf2,
                         Vec3* contact_point, double* penetration_depth, Vec3* normal);
// fcl::collide(BVHModel<OBBRSS>, tf1, Halfspace | Plane, tf2): one contact per intersecting triangle, in the traversal's
// depth-first order, {b1 = triangle, b2 = -1, pos, -normal, depth}.  kind 0 = halfspace, 1 = plane.
size_t collide_mesh_plane(const Model& m1, const Pose& tf1, int kind, const PlaneShape& s, const Pose& tf2, size_t num_max_contacts,
                          bool enable_contact, std::vector<Contact>& out, CollideStats* stats = nullptr);
void brute_mesh_plane(const Model& m1, const Pose& tf1, int kind, const PlaneShape& s, const Pose& tf2, std::vector<int>& tris);

// ---- broadphase: the brute-force manager -----------------------------------------------------------
// BVHModel::computeLocalAABB (BVH_model-inl.h:1080-1100), over the vertices the triangles reference
struct LocalAABB {
  Vec3 center, mn, mx;
  double radius;
};
LocalAABB local_aabb(const Model& m);
// CollisionObject::computeAABB (collision_object-inl.h:118-131): out6 = min, max
void world_aabb(const LocalAABB& a, const Pose& tf, double out6[6]);
// NaiveCollisionManager::collide(other, ...) (broadphase_bruteforce-inl.h:182-205): every (i, j) whose AABBs overlap
// (AABB::overlap, AABB-inl.h:98-107), i-major in registration order
void broadphase_pairs(const std::vector<const Model*>& geoms, const std::vector<int>& geom1, const std::vector<Pose>& tf1,
                      const std::vector<int>& geom2, const std::vector<Pose>& tf2, std::vector<std::pair<int, int>>& pairs);

// brute force over all triangle pairs (for the invariants)
void brute_collide_pairs(const Model& m1, const Pose& tf1, const Model& m2, const Pose& tf2,
                         std::vector<std::pair<int, int>>& pairs);
double brute_distance(const Model& m1, const Pose& tf1, const Model& m2, const Pose& tf2,
                      DistanceOut& out);

}  // namespace oracle
