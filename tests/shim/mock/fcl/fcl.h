// MOCK of the slice of FCL's public API that include/fclgpu/fcl_shim.hpp touches -- test scaffolding only.
// FCL and Eigen are not installed on the build image, so the shim is compile- and run-checked against
// these stand-ins, which reproduce names, members and call signatures of the reference headers
// (include/fcl/...): geometry/bvh/BVH_model.h:160-203 (vertices, tri_indices, num_tris, getBV, getNumBVs),
// geometry/bvh/BV_node_base.h + BV_node.h:50-72 (first_child, bv), math/bv/OBB.h, RSS.h, OBBRSS.h (axis, To,
// extent, l, r), math/triangle.h (operator[]), narrowphase/contact.h:48-91, collision_request.h:52-106,
// collision_result.h (addContact / numContacts / getContact), distance_request.h:52-113,
// distance_result.h (update overloads: only a smaller distance replaces the stored one).
// The dispatch side is mocked too, so that the shim's cell functions can be installed and reached the way a real
// caller reaches them: geometry/collision_geometry.h:53-54 (NODE_TYPE, getNodeType), narrowphase/detail/
// collision_func_matrix.h:67-78 and distance_func_matrix.h:65-76 (the tables of function pointers),
// collision-inl.h:72-76 / distance-inl.h:65-69 (the function-local static tables), collision-inl.h:95-150 /
// distance-inl.h:92-190 (the solver-taking fcl::collide / fcl::distance that look the cell up), request isSatisfied
// (collision_request-inl.h:77-82, distance_request-inl.h:73-77).
// Nothing here computes anything; it is NOT a substitute for FCL.
#pragma once
#include <cstddef>
#include <limits>
#include <vector>

namespace fcl {

template <typename S>
struct Vector3 {
  S v[3];
  Vector3() : v{0, 0, 0} {}
  Vector3(S x, S y, S z) : v{x, y, z} {}
  S& operator[](int i) { return v[i]; }
  const S& operator[](int i) const { return v[i]; }
};

template <typename S>
struct Matrix3 {
  S m[9];  // row-major storage; only operator()(row, col) is part of the mocked surface
  S& operator()(int r, int c) { return m[3 * r + c]; }
  const S& operator()(int r, int c) const { return m[3 * r + c]; }
};

// Eigen::Transform<S, 3, Isometry>: 4x4 column-major, reachable through matrix().data()
template <typename S>
struct Transform3 {
  S m16[16];
  struct MatrixView {
    const S* p;
    const S* data() const { return p; }
  };
  MatrixView matrix() const { return MatrixView{m16}; }
  static Transform3 Identity() {
    Transform3 t;
    for (int i = 0; i < 16; ++i) t.m16[i] = (i % 5 == 0) ? S(1) : S(0);
    return t;
  }
  void setLinear(const S* row_major9) {
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) m16[4 * c + r] = row_major9[3 * r + c];
  }
  void setTranslation(S x, S y, S z) { m16[12] = x; m16[13] = y; m16[14] = z; }
  // writable views like Eigen's tf.linear()(r, c) and tf.translation()[k] (column-major 4x4 storage)
  struct LinearRef {
    S* p;
    S& operator()(int r, int c) { return p[4 * c + r]; }
  };
  struct TranslationRef {
    S* p;
    S& operator[](int k) { return p[12 + k]; }
  };
  LinearRef linear() { return LinearRef{m16}; }
  TranslationRef translation() { return TranslationRef{m16}; }
};

template <typename S>
struct OBB {
  Matrix3<S> axis;
  Vector3<S> To, extent;
};
template <typename S>
struct RSS {
  Matrix3<S> axis;
  Vector3<S> To;
  S l[2];
  S r;
};
template <typename S>
struct OBBRSS {
  OBB<S> obb;
  RSS<S> rss;
};

struct BVNodeBase {
  int first_child = 0, first_primitive = 0, num_primitives = 0;
};
template <typename BV>
struct BVNode : BVNodeBase {
  BV bv;
};

struct Triangle {
  std::size_t vids[3];
  std::size_t operator[](int i) const { return vids[i]; }
};

enum NODE_TYPE {BV_UNKNOWN, BV_AABB, BV_OBB, BV_RSS, BV_kIOS, BV_OBBRSS, BV_KDOP16, BV_KDOP18, BV_KDOP24,
                GEOM_BOX, GEOM_SPHERE, GEOM_ELLIPSOID, GEOM_CAPSULE, GEOM_CONE, GEOM_CYLINDER, GEOM_CONVEX, GEOM_PLANE, GEOM_HALFSPACE, GEOM_TRIANGLE, GEOM_OCTREE, NODE_COUNT};
enum OBJECT_TYPE {OT_UNKNOWN, OT_BVH, OT_GEOM, OT_OCTREE, OT_COUNT};

template <typename S>
struct CollisionGeometry {
  virtual ~CollisionGeometry() {}
  virtual NODE_TYPE getNodeType() const { return BV_UNKNOWN; }
  virtual OBJECT_TYPE getObjectType() const { return OT_UNKNOWN; }
};

template <typename BV>
struct BVHModel : CollisionGeometry<double> {
  Vector3<double>* vertices = nullptr;
  Triangle* tri_indices = nullptr;
  int num_tris = 0, num_vertices = 0;
  NODE_TYPE getNodeType() const override { return BV_OBBRSS; }  // the mock only instantiates BVHModel<OBBRSS<double>>
  OBJECT_TYPE getObjectType() const override { return OT_BVH; }
  int getNumBVs() const { return (int)bvs_.size(); }
  const BVNode<BV>& getBV(int i) const { return bvs_[i]; }
  // storage of the mock
  std::vector<BVNode<BV>> bvs_;
  std::vector<Vector3<double>> verts_;
  std::vector<Triangle> tris_;
};

// geometry/shape/sphere.h: public member `radius`
template <typename S>
struct Sphere : CollisionGeometry<S> {
  S radius;
  explicit Sphere(S r) : radius(r) {}
  NODE_TYPE getNodeType() const override { return GEOM_SPHERE; }
  OBJECT_TYPE getObjectType() const override { return OT_GEOM; }
};

// geometry/shape/halfspace.h, plane.h: public members n (unit normal) and d
template <typename S>
struct Halfspace : CollisionGeometry<S> {
  Vector3<S> n;
  S d;
  Halfspace(const Vector3<S>& n_, S d_) : n(n_), d(d_) {}
  NODE_TYPE getNodeType() const override { return GEOM_HALFSPACE; }
  OBJECT_TYPE getObjectType() const override { return OT_GEOM; }
};
template <typename S>
struct Plane : CollisionGeometry<S> {
  Vector3<S> n;
  S d;
  Plane(const Vector3<S>& n_, S d_) : n(n_), d(d_) {}
  NODE_TYPE getNodeType() const override { return GEOM_PLANE; }
  OBJECT_TYPE getObjectType() const override { return OT_GEOM; }
};

template <typename S>
struct Contact {
  static const int NONE = -1;  // contact.h: id of "no primitive" (shape side)
  const CollisionGeometry<S>* o1 = nullptr;
  const CollisionGeometry<S>* o2 = nullptr;
  int b1 = -1, b2 = -1;
  Vector3<S> normal, pos;
  S penetration_depth = 0;
  Contact() {}
  Contact(const CollisionGeometry<S>* a, const CollisionGeometry<S>* b, int i, int j) : o1(a), o2(b), b1(i), b2(j) {}
  Contact(const CollisionGeometry<S>* a, const CollisionGeometry<S>* b, int i, int j, const Vector3<S>& p,
          const Vector3<S>& n, S depth)
      : o1(a), o2(b), b1(i), b2(j), normal(n), pos(p), penetration_depth(depth) {}
};

template <typename S>
struct CollisionResult;

template <typename S>
struct CollisionRequest {
  std::size_t num_max_contacts = 1;
  bool enable_contact = false;
  std::size_t num_max_cost_sources = 1;
  bool enable_cost = false;
  CollisionRequest(std::size_t n = 1, bool contact = false) : num_max_contacts(n), enable_contact(contact) {}
  bool isSatisfied(const CollisionResult<S>& result) const;  // collision_request-inl.h:77-82
};

template <typename S>
struct CollisionResult {
  void addContact(const Contact<S>& c) { contacts_.push_back(c); }
  std::size_t numContacts() const { return contacts_.size(); }
  const Contact<S>& getContact(std::size_t i) const { return contacts_[i]; }
  bool isCollision() const { return !contacts_.empty(); }
  std::vector<Contact<S>> contacts_;
};

template <typename S>
bool CollisionRequest<S>::isSatisfied(const CollisionResult<S>& result) const {
  return (!enable_cost) && result.isCollision() && (num_max_contacts <= result.numContacts());
}

// narrowphase/continuous_collision_request.h:48-80, continuous_collision_result.h
enum CCDMotionType { CCDM_TRANS, CCDM_LINEAR, CCDM_SCREW, CCDM_SPLINE };
enum CCDSolverType { CCDC_NAIVE, CCDC_CONSERVATIVE_ADVANCEMENT, CCDC_RAY_SHOOTING, CCDC_POLYNOMIAL_SOLVER };
enum GJKSolverType { GST_LIBCCD, GST_INDEP };
template <typename S>
struct ContinuousCollisionRequest {
  std::size_t num_max_iterations;
  S toc_err;
  CCDMotionType ccd_motion_type;
  GJKSolverType gjk_solver_type;
  CCDSolverType ccd_solver_type;
  ContinuousCollisionRequest(std::size_t it = 10, S err = 0.0001, CCDMotionType m = CCDM_TRANS, GJKSolverType g = GST_LIBCCD,
                             CCDSolverType c = CCDC_NAIVE)
      : num_max_iterations(it), toc_err(err), ccd_motion_type(m), gjk_solver_type(g), ccd_solver_type(c) {}
};
template <typename S>
struct ContinuousCollisionResult {
  bool is_collide = false;
  S time_of_contact = 1.0;
  Transform3<S> contact_tf1, contact_tf2;
};

template <typename S>
struct DistanceResult;

template <typename S>
struct DistanceRequest {
  bool enable_nearest_points, enable_signed_distance = false;
  S rel_err = 0, abs_err = 0;
  explicit DistanceRequest(bool nearest = false) : enable_nearest_points(nearest) {}
  bool isSatisfied(const DistanceResult<S>& result) const;  // distance_request-inl.h:73-77
};

template <typename S>
struct DistanceResult {
  S min_distance = std::numeric_limits<S>::max();
  Vector3<S> nearest_points[2];
  const CollisionGeometry<S>* o1 = nullptr;
  const CollisionGeometry<S>* o2 = nullptr;
  static const int NONE = -1;  // distance_result.h:78
  int b1 = -1, b2 = -1;
  void update(S d, const CollisionGeometry<S>* a, const CollisionGeometry<S>* b, int i, int j) {
    if (min_distance > d) { min_distance = d; o1 = a; o2 = b; b1 = i; b2 = j; }
  }
  void update(S d, const CollisionGeometry<S>* a, const CollisionGeometry<S>* b, int i, int j, const Vector3<S>& p1,
              const Vector3<S>& p2) {
    if (min_distance > d) { min_distance = d; o1 = a; o2 = b; b1 = i; b2 = j; nearest_points[0] = p1; nearest_points[1] = p2; }
  }
};

template <typename S>
bool DistanceRequest<S>::isSatisfied(const DistanceResult<S>& result) const {
  return result.min_distance <= 0;
}

namespace detail {
template <typename S_>
struct GJKSolver_libccd {  // narrowphase/detail/gjk_solver_libccd.h: only the scalar type matters to the tables
  using S = S_;
};

template <typename NarrowPhaseSolver>
struct CollisionFunctionMatrix {  // detail/collision_func_matrix.h:53-84
  using S = typename NarrowPhaseSolver::S;
  using CollisionFunc = std::size_t (*)(const CollisionGeometry<S>* o1, const Transform3<S>& tf1, const CollisionGeometry<S>* o2,
                                        const Transform3<S>& tf2, const NarrowPhaseSolver* nsolver,
                                        const CollisionRequest<S>& request, CollisionResult<S>& result);
  CollisionFunc collision_matrix[NODE_COUNT][NODE_COUNT];
  CollisionFunctionMatrix() {
    for (int i = 0; i < NODE_COUNT; ++i)
      for (int j = 0; j < NODE_COUNT; ++j) collision_matrix[i][j] = nullptr;
  }
};

template <typename NarrowPhaseSolver>
struct DistanceFunctionMatrix {  // detail/distance_func_matrix.h:53-82
  using S = typename NarrowPhaseSolver::S;
  using DistanceFunc = S (*)(const CollisionGeometry<S>* o1, const Transform3<S>& tf1, const CollisionGeometry<S>* o2,
                             const Transform3<S>& tf2, const NarrowPhaseSolver* nsolver, const DistanceRequest<S>& request,
                             DistanceResult<S>& result);
  DistanceFunc distance_matrix[NODE_COUNT][NODE_COUNT];
  DistanceFunctionMatrix() {
    for (int i = 0; i < NODE_COUNT; ++i)
      for (int j = 0; j < NODE_COUNT; ++j) distance_matrix[i][j] = nullptr;
  }
};
}  // namespace detail

template <typename GJKSolver>
detail::CollisionFunctionMatrix<GJKSolver>& getCollisionFunctionLookTable() {  // collision-inl.h:72-76
  static detail::CollisionFunctionMatrix<GJKSolver> table;
  return table;
}
template <typename GJKSolver>
detail::DistanceFunctionMatrix<GJKSolver>& getDistanceFunctionLookTable() {  // distance-inl.h:65-69
  static detail::DistanceFunctionMatrix<GJKSolver> table;
  return table;
}

// the solver-taking fcl::collide (collision-inl.h:95-150): guard on num_max_contacts, (OT_GEOM, OT_BVH) swap, table look-up
template <typename S, typename NarrowPhaseSolver>
std::size_t collide(const CollisionGeometry<S>* o1, const Transform3<S>& tf1, const CollisionGeometry<S>* o2, const Transform3<S>& tf2,
                    const NarrowPhaseSolver* nsolver, const CollisionRequest<S>& request, CollisionResult<S>& result) {
  const auto& looktable = getCollisionFunctionLookTable<NarrowPhaseSolver>();
  if (request.num_max_contacts == 0) return 0;
  const NODE_TYPE t1 = o1->getNodeType(), t2 = o2->getNodeType();
  if (o1->getObjectType() == OT_GEOM && o2->getObjectType() == OT_BVH) {
    if (!looktable.collision_matrix[t2][t1]) return 0;
    return looktable.collision_matrix[t2][t1](o2, tf2, o1, tf1, nsolver, request, result);
  }
  if (!looktable.collision_matrix[t1][t2]) return 0;
  return looktable.collision_matrix[t1][t2](o1, tf1, o2, tf2, nsolver, request, result);
}

// the solver-taking fcl::distance (distance-inl.h:92-190)
template <typename S, typename NarrowPhaseSolver>
S distance(const CollisionGeometry<S>* o1, const Transform3<S>& tf1, const CollisionGeometry<S>* o2, const Transform3<S>& tf2,
           const NarrowPhaseSolver* nsolver, const DistanceRequest<S>& request, DistanceResult<S>& result) {
  const auto& looktable = getDistanceFunctionLookTable<NarrowPhaseSolver>();
  const NODE_TYPE t1 = o1->getNodeType(), t2 = o2->getNodeType();
  if (o1->getObjectType() == OT_GEOM && o2->getObjectType() == OT_BVH) {
    if (!looktable.distance_matrix[t2][t1]) return std::numeric_limits<S>::max();
    return looktable.distance_matrix[t2][t1](o2, tf2, o1, tf1, nsolver, request, result);
  }
  if (!looktable.distance_matrix[t1][t2]) return std::numeric_limits<S>::max();
  return looktable.distance_matrix[t1][t2](o1, tf1, o2, tf2, nsolver, request, result);
}

}  // namespace fcl
