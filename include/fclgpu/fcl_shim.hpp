// fcl_shim.hpp — the literal FCL drop-in above the C ABI (include/fclgpu.h).
//
// Compiled only when FCL and Eigen are available (neither is on the build image, so this header
// is dormant there; INTEGRATION.md shows how a maintainer wires it in).  It
//   * uploads an existing fcl::BVHModel<fcl::OBBRSS<double>> (no rebuild: the node tree FCL
//     built is flattened as is),
//   * evaluates batches of (tf1, tf2) pairs and fills std::vector<fcl::CollisionResult<double>>
//     / std::vector<fcl::DistanceResult<double>>,
//   * offers single-query functions with the exact signature of the reference's dispatch-table
//     cell -- collide_cell<Solver> / distance_cell<Solver> (and the mesh-vs-sphere pair) are
//     CollisionFunc / DistanceFunc (detail/collision_func_matrix.h:67-78,
//     detail/distance_func_matrix.h:65-76) -- and install<Solver>() / uninstall<Solver>(), which
//     put them into collision_matrix[BV_OBBRSS][BV_OBBRSS] / distance_matrix[..][..] (and the
//     [BV_OBBRSS][GEOM_SPHERE] cells) of the look-up tables fcl::collide / fcl::distance dispatch
//     through (collision-inl.h:72-76, distance-inl.h:65-69).
#pragma once
#include "../fclgpu.h"

#if defined(__has_include)
#if __has_include(<fcl/fcl.h>) && __has_include(<Eigen/Core>)
#define FCLGPU_HAVE_FCL 1
#endif
#endif

#ifdef FCLGPU_HAVE_FCL
#include <fcl/fcl.h>

#include <algorithm>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <cmath>
#include <limits>
#include <vector>

namespace fclgpu {

using BVH = fcl::BVHModel<fcl::OBBRSS<double>>;

inline void check(int rc) {
  if (rc != FCLGPU_OK) throw std::runtime_error(std::string("fclgpu: ") + fclgpu_last_error());
}

// 12-double pose record from an Eigen isometry (column-major 4x4)
inline void to_pose(const fcl::Transform3<double>& tf, double* p) { fclgpu_pose_from_colmajor4x4(tf.matrix().data(), p); }

// Device-resident copy of a built BVHModel<OBBRSS<double>> (BVH_model.h:160-203, BV_node.h:50-72).
class DeviceModel {
 public:
  DeviceModel(const BVH& m, int device = 0) : host_(&m) {
    const int n = m.getNumBVs(), nt = m.num_tris;
    std::vector<int32_t> fc(n);
    std::vector<double> axis(9 * n), raxis(9 * n), oT(3 * n), oe(3 * n), rT(3 * n), rl(2 * n), rr(n), tv(9 * nt);
    for (int i = 0; i < n; ++i) {
      const auto& node = m.getBV(i);
      fc[i] = node.first_child;
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
          axis[9 * i + 3 * r + c] = node.bv.obb.axis(r, c);
          raxis[9 * i + 3 * r + c] = node.bv.rss.axis(r, c);  // equal to obb.axis unless FCL refitted the model bottom-up
        }
      for (int k = 0; k < 3; ++k) {
        oT[3 * i + k] = node.bv.obb.To[k];
        oe[3 * i + k] = node.bv.obb.extent[k];
        rT[3 * i + k] = node.bv.rss.To[k];
      }
      rl[2 * i] = node.bv.rss.l[0];
      rl[2 * i + 1] = node.bv.rss.l[1];
      rr[i] = node.bv.rss.r;
    }
    for (int t = 0; t < nt; ++t)
      for (int k = 0; k < 3; ++k)
        for (int c = 0; c < 3; ++c) tv[9 * t + 3 * k + c] = m.vertices[m.tri_indices[t][k]][c];
    check(fclgpu_model_create_obbrss2(device, n, fc.data(), axis.data(), oT.data(), oe.data(), raxis.data(), rT.data(), rl.data(),
                                      rr.data(), nt, tv.data(), &h_));
    // Refit topology.  BVHModel::primitive_indices is private (BVH_model.h:191), but the public node fields give it
    // back: node i owns primitive_indices[first_primitive .. first_primitive + num_primitives), a leaf owns exactly one
    // slot and names its triangle in first_child, so every slot is written by exactly one leaf.
    std::vector<int32_t> first(n), count(n), prim(nt, -1), tidx(3 * (std::size_t)nt);
    bool ok = m.num_vertices > 0;
    for (int i = 0; i < n; ++i) {
      const auto& node = m.getBV(i);
      first[i] = node.first_primitive;
      count[i] = node.num_primitives;
      if (node.first_child < 0) {
        if (node.first_primitive < 0 || node.first_primitive >= nt) ok = false;
        else prim[node.first_primitive] = -(node.first_child + 1);
      }
    }
    for (int t = 0; t < nt; ++t) {
      if (prim[t] < 0) ok = false;
      for (int k = 0; k < 3; ++k) tidx[3 * (std::size_t)t + k] = (int32_t)m.tri_indices[t][k];
    }
    if (ok) refit_ready_ = fclgpu_model_set_partition(h_, m.num_vertices, tidx.data(), first.data(), count.data(), prim.data()) == FCLGPU_OK;
  }
  ~DeviceModel() { fclgpu_model_destroy(h_); }
  DeviceModel(const DeviceModel&) = delete;
  DeviceModel& operator=(const DeviceModel&) = delete;
  const fclgpu_model* handle() const { return h_; }
  const BVH* host() const { return host_; }
  // endReplaceModel(refit = true, bottomup = false) on the device copy: `vertices` = the host model's updated array
  bool refit_ready() const { return refit_ready_; }
  void refit_topdown() {
    std::vector<double> v(3 * (std::size_t)host_->num_vertices);
    for (int i = 0; i < host_->num_vertices; ++i)
      for (int c = 0; c < 3; ++c) v[3 * (std::size_t)i + c] = host_->vertices[i][c];
    check(fclgpu_model_refit_topdown(h_, v.data(), host_->num_vertices, 0, nullptr));
    check(fclgpu_sync_status(fclgpu_model_device(h_), nullptr));
  }
  // endReplaceModel() with its default arguments (refit = true, bottomup = true) on the device copy
  void refit_bottomup() {
    std::vector<double> v(3 * (std::size_t)host_->num_vertices);
    for (int i = 0; i < host_->num_vertices; ++i)
      for (int c = 0; c < 3; ++c) v[3 * (std::size_t)i + c] = host_->vertices[i][c];
    check(fclgpu_model_refit_bottomup(h_, v.data(), host_->num_vertices, 0, nullptr));
    check(fclgpu_sync_status(fclgpu_model_device(h_), nullptr));
  }

 private:
  const BVH* host_;
  fclgpu_model* h_ = nullptr;
  bool refit_ready_ = false;
};

inline void same_size(std::size_t a, std::size_t b) {
  if (a != b) throw std::invalid_argument("fclgpu: tf1 and tf2 must have the same length");
}

// n independent fcl::collide(o1, tf1[i], o2, tf2[i], request, results[i]) calls
inline void collide(const DeviceModel& o1, const std::vector<fcl::Transform3<double>>& tf1, const DeviceModel& o2,
                    const std::vector<fcl::Transform3<double>>& tf2, const fcl::CollisionRequest<double>& request,
                    std::vector<fcl::CollisionResult<double>>& results) {
  const int64_t n = (int64_t)tf1.size();
  same_size(tf1.size(), tf2.size());
  results.assign(n, fcl::CollisionResult<double>());
  if (request.num_max_contacts == 0 || n == 0) return;
  std::vector<double> p1(12 * n), p2(12 * n);
  for (int64_t i = 0; i < n; ++i) {
    to_pose(tf1[i], &p1[12 * i]);
    to_pose(tf2[i], &p2[12 * i]);
  }
  fclgpu_collision_request req{(int64_t)std::min<std::size_t>(request.num_max_contacts, (std::size_t)1 << 62),
                               request.enable_contact ? 1 : 0, request.enable_cost ? 1 : 0, 0, FCLGPU_CONTACT_FULL, 0};
  std::vector<int32_t> counts(n);
  std::vector<int64_t> off(n + 1);
  int64_t cap = std::max<int64_t>(64 * n, 1024);
  std::vector<fclgpu_contact> pool(cap);
  int rc = fclgpu_collide_batch_host(o1.handle(), o2.handle(), n, p1.data(), p2.data(), &req, counts.data(),
                                     pool.data(), cap, off.data(), nullptr, nullptr);
  if (rc == FCLGPU_ERR_CONTACT_OVERFLOW) {  // counts are exact: size the pool and rerun
    cap = 0;
    int32_t mx = 1;
    for (int32_t c : counts) { cap += c; mx = std::max(mx, c); }
    req.stage_capacity = mx;  // per call: the reference is re-entrant, so no process-wide setting is touched
    pool.resize(cap);
    rc = fclgpu_collide_batch_host(o1.handle(), o2.handle(), n, p1.data(), p2.data(), &req, counts.data(), pool.data(),
                                   cap, off.data(), nullptr, nullptr);
  }
  check(rc);
  for (int64_t i = 0; i < n; ++i)
    for (int64_t k = off[i]; k < off[i] + counts[i]; ++k) {
      const fclgpu_contact& c = pool[k];
      if (request.enable_contact)
        results[i].addContact(fcl::Contact<double>(o1.host(), o2.host(), c.b1, c.b2,
                                                   fcl::Vector3<double>(c.pos[0], c.pos[1], c.pos[2]),
                                                   fcl::Vector3<double>(c.normal[0], c.normal[1], c.normal[2]),
                                                   c.penetration_depth));
      else
        results[i].addContact(fcl::Contact<double>(o1.host(), o2.host(), c.b1, c.b2));
    }
}

// n independent fcl::collide(mesh, tf1[i], sphere, tf2[i], request, results[i]) calls
// (collision_matrix[BV_OBBRSS][GEOM_SPHERE]; contacts carry b2 = Contact::NONE)
inline void collide(const DeviceModel& o1, const std::vector<fcl::Transform3<double>>& tf1, const fcl::Sphere<double>& sphere,
                    const std::vector<fcl::Transform3<double>>& tf2, const fcl::CollisionRequest<double>& request,
                    std::vector<fcl::CollisionResult<double>>& results) {
  const int64_t n = (int64_t)tf1.size();
  same_size(tf1.size(), tf2.size());
  results.assign(n, fcl::CollisionResult<double>());
  if (request.num_max_contacts == 0 || n == 0) return;
  std::vector<double> p1(12 * n), p2(12 * n);
  for (int64_t i = 0; i < n; ++i) {
    to_pose(tf1[i], &p1[12 * i]);
    to_pose(tf2[i], &p2[12 * i]);
  }
  fclgpu_collision_request req{(int64_t)std::min<std::size_t>(request.num_max_contacts, (std::size_t)1 << 62),
                               request.enable_contact ? 1 : 0, request.enable_cost ? 1 : 0, 0, FCLGPU_CONTACT_FULL, 0};
  std::vector<int32_t> counts(n);
  std::vector<int64_t> off(n + 1);
  int64_t cap = std::max<int64_t>(64 * n, 1024);
  std::vector<fclgpu_contact> pool(cap);
  int rc = fclgpu_collide_mesh_sphere_batch_host(o1.handle(), sphere.radius, n, p1.data(), p2.data(), &req, counts.data(),
                                                 pool.data(), cap, off.data(), nullptr, nullptr);
  if (rc == FCLGPU_ERR_CONTACT_OVERFLOW) {  // counts are exact: size the pool and rerun
    cap = 0;
    int32_t mx = 1;
    for (int32_t c : counts) { cap += c; mx = std::max(mx, c); }
    req.stage_capacity = mx;  // per call: the reference is re-entrant, so no process-wide setting is touched
    pool.resize(cap);
    rc = fclgpu_collide_mesh_sphere_batch_host(o1.handle(), sphere.radius, n, p1.data(), p2.data(), &req, counts.data(),
                                               pool.data(), cap, off.data(), nullptr, nullptr);
  }
  check(rc);
  for (int64_t i = 0; i < n; ++i)
    for (int64_t k = off[i]; k < off[i] + counts[i]; ++k) {
      const fclgpu_contact& c = pool[k];
      if (request.enable_contact)
        results[i].addContact(fcl::Contact<double>(o1.host(), &sphere, c.b1, fcl::Contact<double>::NONE,
                                                   fcl::Vector3<double>(c.pos[0], c.pos[1], c.pos[2]),
                                                   fcl::Vector3<double>(c.normal[0], c.normal[1], c.normal[2]),
                                                   c.penetration_depth));
      else
        results[i].addContact(fcl::Contact<double>(o1.host(), &sphere, c.b1, fcl::Contact<double>::NONE));
    }
}

// n independent fcl::collide(mesh, tf1[i], Halfspace | Plane, tf2[i], request, results[i]) calls
// (collision_matrix[BV_OBBRSS][GEOM_HALFSPACE] / [GEOM_PLANE], collision_func_matrix-inl.h:841-842; contacts carry b2 = NONE).
// Shape: fcl::Halfspace<double> or fcl::Plane<double> (public members n, d; geometry/shape/halfspace.h, plane.h).
template <typename Shape>
inline void collide_plane_like(const DeviceModel& o1, const std::vector<fcl::Transform3<double>>& tf1, const Shape& shape, int32_t kind,
                               const std::vector<fcl::Transform3<double>>& tf2, const fcl::CollisionRequest<double>& request,
                               std::vector<fcl::CollisionResult<double>>& results) {
  const int64_t n = (int64_t)tf1.size();
  same_size(tf1.size(), tf2.size());
  results.assign(n, fcl::CollisionResult<double>());
  if (request.num_max_contacts == 0 || n == 0) return;
  std::vector<double> p1(12 * n), p2(12 * n);
  for (int64_t i = 0; i < n; ++i) {
    to_pose(tf1[i], &p1[12 * i]);
    to_pose(tf2[i], &p2[12 * i]);
  }
  const double nrm[3] = {shape.n[0], shape.n[1], shape.n[2]};
  fclgpu_collision_request req{(int64_t)std::min<std::size_t>(request.num_max_contacts, (std::size_t)1 << 62),
                               request.enable_contact ? 1 : 0, request.enable_cost ? 1 : 0, 0, FCLGPU_CONTACT_FULL, 0};
  std::vector<int32_t> counts(n);
  std::vector<int64_t> off(n + 1);
  int64_t cap = std::max<int64_t>(64 * n, 1024);
  std::vector<fclgpu_contact> pool(cap);
  int rc = fclgpu_collide_mesh_plane_batch_host(o1.handle(), kind, nrm, shape.d, n, p1.data(), p2.data(), &req, counts.data(), pool.data(),
                                                cap, off.data(), nullptr, nullptr);
  if (rc == FCLGPU_ERR_CONTACT_OVERFLOW) {  // counts are exact: size the pool and rerun
    cap = 0;
    int32_t mx = 1;
    for (int32_t c : counts) { cap += c; mx = std::max(mx, c); }
    req.stage_capacity = mx;
    pool.resize(cap);
    rc = fclgpu_collide_mesh_plane_batch_host(o1.handle(), kind, nrm, shape.d, n, p1.data(), p2.data(), &req, counts.data(), pool.data(), cap,
                                              off.data(), nullptr, nullptr);
  }
  check(rc);
  for (int64_t i = 0; i < n; ++i)
    for (int64_t k = off[i]; k < off[i] + counts[i]; ++k) {
      const fclgpu_contact& c = pool[k];
      if (request.enable_contact)
        results[i].addContact(fcl::Contact<double>(o1.host(), &shape, c.b1, fcl::Contact<double>::NONE,
                                                   fcl::Vector3<double>(c.pos[0], c.pos[1], c.pos[2]),
                                                   fcl::Vector3<double>(c.normal[0], c.normal[1], c.normal[2]), c.penetration_depth));
      else
        results[i].addContact(fcl::Contact<double>(o1.host(), &shape, c.b1, fcl::Contact<double>::NONE));
    }
}
inline void collide(const DeviceModel& o1, const std::vector<fcl::Transform3<double>>& tf1, const fcl::Halfspace<double>& hs,
                    const std::vector<fcl::Transform3<double>>& tf2, const fcl::CollisionRequest<double>& request,
                    std::vector<fcl::CollisionResult<double>>& results) {
  collide_plane_like(o1, tf1, hs, FCLGPU_SHAPE_HALFSPACE, tf2, request, results);
}
inline void collide(const DeviceModel& o1, const std::vector<fcl::Transform3<double>>& tf1, const fcl::Plane<double>& pl,
                    const std::vector<fcl::Transform3<double>>& tf2, const fcl::CollisionRequest<double>& request,
                    std::vector<fcl::CollisionResult<double>>& results) {
  collide_plane_like(o1, tf1, pl, FCLGPU_SHAPE_PLANE, tf2, request, results);
}

inline void distance(const DeviceModel& o1, const std::vector<fcl::Transform3<double>>& tf1, const DeviceModel& o2,
                     const std::vector<fcl::Transform3<double>>& tf2, const fcl::DistanceRequest<double>& request,
                     std::vector<fcl::DistanceResult<double>>& results) {
  const int64_t n = (int64_t)tf1.size();
  same_size(tf1.size(), tf2.size());
  results.assign(n, fcl::DistanceResult<double>());
  if (n == 0) return;
  std::vector<double> p1(12 * n), p2(12 * n), d(n), a(3 * n), b(3 * n);
  std::vector<int32_t> b1(n), b2(n);
  for (int64_t i = 0; i < n; ++i) {
    to_pose(tf1[i], &p1[12 * i]);
    to_pose(tf2[i], &p2[12 * i]);
  }
  fclgpu_distance_request req{request.enable_nearest_points ? 1 : 0, request.enable_signed_distance ? 1 : 0,
                              request.rel_err, request.abs_err};
  check(fclgpu_distance_batch_host(o1.handle(), o2.handle(), n, p1.data(), p2.data(), &req, d.data(), a.data(),
                                   b.data(), b1.data(), b2.data(), nullptr, nullptr));
  for (int64_t i = 0; i < n; ++i) {
    if (request.enable_nearest_points)
      results[i].update(d[i], o1.host(), o2.host(), b1[i], b2[i], fcl::Vector3<double>(a[3 * i], a[3 * i + 1], a[3 * i + 2]),
                        fcl::Vector3<double>(b[3 * i], b[3 * i + 1], b[3 * i + 2]));
    else
      results[i].update(d[i], o1.host(), o2.host(), b1[i], b2[i]);
  }
}

// Tolerance verification (extension, see fclgpu_distance_cutoff_batch in fclgpu.h): within[i] = 1 iff
// fcl::distance(o1, tf1[i], o2, tf2[i]) <= tolerance; node pairs farther apart than the tolerance are never descended and a
// query ends at the first triangle pair found within it (fclgpu_within_tolerance_batch).
inline void within_tolerance(const DeviceModel& o1, const std::vector<fcl::Transform3<double>>& tf1, const DeviceModel& o2,
                             const std::vector<fcl::Transform3<double>>& tf2, double tolerance, std::vector<char>& within) {
  const int64_t n = (int64_t)tf1.size();
  same_size(tf1.size(), tf2.size());
  within.assign(n, 0);
  if (n == 0) return;
  std::vector<double> p1(12 * n), p2(12 * n);
  for (int64_t i = 0; i < n; ++i) {
    to_pose(tf1[i], &p1[12 * i]);
    to_pose(tf2[i], &p2[12 * i]);
  }
  check(fclgpu_within_tolerance_batch_host(o1.handle(), o2.handle(), n, p1.data(), p2.data(), tolerance,
                                           reinterpret_cast<uint8_t*>(within.data()), nullptr, nullptr, nullptr));
}

// n independent fcl::distance(mesh, tf1[i], sphere, tf2[i], request, results[i]) calls
// (distance_matrix[BV_OBBRSS][GEOM_SPHERE]).  Like the reference's mesh-shape leaf, the nearest points are always
// handed to DistanceResult::update and stay in the local frames of the mesh and of the sphere; b2 = NONE.
inline void distance(const DeviceModel& o1, const std::vector<fcl::Transform3<double>>& tf1, const fcl::Sphere<double>& sphere,
                     const std::vector<fcl::Transform3<double>>& tf2, const fcl::DistanceRequest<double>& request,
                     std::vector<fcl::DistanceResult<double>>& results) {
  const int64_t n = (int64_t)tf1.size();
  same_size(tf1.size(), tf2.size());
  results.assign(n, fcl::DistanceResult<double>());
  if (n == 0) return;
  std::vector<double> p1(12 * n), p2(12 * n), d(n), a(3 * n), b(3 * n);
  std::vector<int32_t> b1(n);
  for (int64_t i = 0; i < n; ++i) {
    to_pose(tf1[i], &p1[12 * i]);
    to_pose(tf2[i], &p2[12 * i]);
  }
  fclgpu_distance_request req{1, request.enable_signed_distance ? 1 : 0, request.rel_err, request.abs_err};
  check(fclgpu_distance_mesh_sphere_batch_host(o1.handle(), sphere.radius, n, p1.data(), p2.data(), &req, d.data(), a.data(),
                                               b.data(), b1.data(), nullptr, nullptr, nullptr));
  for (int64_t i = 0; i < n; ++i)
    results[i].update(d[i], o1.host(), &sphere, b1[i], fcl::DistanceResult<double>::NONE,
                      fcl::Vector3<double>(a[3 * i], a[3 * i + 1], a[3 * i + 2]),
                      fcl::Vector3<double>(b[3 * i], b[3 * i + 1], b[3 * i + 2]));
}

// ---------------------------------------------------------------------------------------------------------------------
// Single-query drop-in: the dispatch-table cells.
// fcl::collide / fcl::distance look the pair (getNodeType(), getNodeType()) up in a table of function pointers
// (collision-inl.h:95-150, distance-inl.h:92-190); the functions below have the cell signature (CollisionFunc,
// detail/collision_func_matrix.h:67-78; DistanceFunc, detail/distance_func_matrix.h:65-76) and the cell semantics of
// orientedMeshCollide / orientedMeshDistance (collision_func_matrix-inl.h:571-590, distance_func_matrix-inl.h:386-403):
// early return when the request is already satisfied by the result, otherwise the query's contacts are APPENDED to the
// caller's result (whose contacts count against num_max_contacts) / the result is updated when the distance is smaller.
// A batch of one pays a launch and two PCIe copies: the batched overloads above are the product path, these make
// unmodified single-query callers work.  Device copies of the models come from a process-wide cache keyed by the host
// model's address (upload on first use; evict() when a model is rebuilt or destroyed).
// ---------------------------------------------------------------------------------------------------------------------
// n independent fcl::continuousCollide(o1, tf1_beg[i], tf1_end[i], o2, tf2_beg[i], tf2_end[i], request, results[i]) calls
// (narrowphase/continuous_collision-inl.h:441-452).  Built: ccd_motion_type = CCDM_TRANS with ccd_solver_type =
// CCDC_CONSERVATIVE_ADVANCEMENT; other settings throw (the C ABI answers FCLGPU_ERR_UNSUPPORTED_FUNCTION).
inline void from_pose(const double* p, fcl::Transform3<double>& tf) {
  tf = fcl::Transform3<double>::Identity();
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) tf.linear()(r, c) = p[3 * r + c];
    tf.translation()[r] = p[9 + r];
  }
}
inline void continuous_collide(const DeviceModel& o1, const std::vector<fcl::Transform3<double>>& tf1_beg,
                               const std::vector<fcl::Transform3<double>>& tf1_end, const DeviceModel& o2,
                               const std::vector<fcl::Transform3<double>>& tf2_beg, const std::vector<fcl::Transform3<double>>& tf2_end,
                               const fcl::ContinuousCollisionRequest<double>& request,
                               std::vector<fcl::ContinuousCollisionResult<double>>& results) {
  const std::size_t n = tf1_beg.size();
  same_size(n, tf1_end.size());
  same_size(n, tf2_beg.size());
  same_size(n, tf2_end.size());
  std::vector<double> p(4 * 12 * n), toc(n), c1(12 * n), c2(12 * n);
  std::vector<int32_t> hit(n);
  for (std::size_t i = 0; i < n; ++i) {
    to_pose(tf1_beg[i], &p[12 * i]);
    to_pose(tf1_end[i], &p[12 * (n + i)]);
    to_pose(tf2_beg[i], &p[12 * (2 * n + i)]);
    to_pose(tf2_end[i], &p[12 * (3 * n + i)]);
  }
  fclgpu_continuous_request req{(int64_t)request.num_max_iterations, request.toc_err, (int32_t)request.ccd_motion_type,
                                (int32_t)request.gjk_solver_type, (int32_t)request.ccd_solver_type};
  check(fclgpu_continuous_collide_batch_host(o1.handle(), o2.handle(), (int64_t)n, &p[0], &p[12 * n], &p[24 * n], &p[36 * n], &req,
                                             hit.data(), toc.data(), c1.data(), c2.data(), nullptr));
  results.assign(n, fcl::ContinuousCollisionResult<double>());
  for (std::size_t i = 0; i < n; ++i) {
    results[i].is_collide = hit[i] != 0;
    results[i].time_of_contact = toc[i];
    if (hit[i]) {  // the reference sets the contact transforms only when there is a contact
      from_pose(&c1[12 * i], results[i].contact_tf1);
      from_pose(&c2[12 * i], results[i].contact_tf2);
    }
  }
}

class ModelCache {
 public:
  DeviceModel& get(const BVH* m, int device = 0) {
    std::lock_guard<std::mutex> g(mu_);
    auto& slot = map_[std::make_pair(m, device)];
    if (!slot) slot.reset(new DeviceModel(*m, device));
    return *slot;
  }
  void evict(const BVH* m) {
    std::lock_guard<std::mutex> g(mu_);
    for (auto it = map_.begin(); it != map_.end();) it = (it->first.first == m) ? map_.erase(it) : std::next(it);
  }
  void clear() {
    std::lock_guard<std::mutex> g(mu_);
    map_.clear();
  }

 private:
  std::mutex mu_;
  std::map<std::pair<const BVH*, int>, std::unique_ptr<DeviceModel>> map_;
};
inline ModelCache& model_cache() {
  static ModelCache c;
  return c;
}
// Give back what the library keeps between calls on `device`: the growable workspace buffers (fclgpu_device_trim) and,
// with models = true, the cached device copies of every BVHModel as well.  Returns the workspace bytes released.
inline std::int64_t release_device_memory(int device = 0, bool models = false) {
  if (models) model_cache().clear();
  std::int64_t released = 0;
  check(fclgpu_device_trim(device, &released));
  return released;
}

namespace detail {
// the budget a non-empty result leaves (mesh_collision_traversal_node-inl.h:553-556, 594-600: addContact only while
// request.num_max_contacts > result.numContacts())
inline fcl::CollisionRequest<double> remaining(const fcl::CollisionRequest<double>& request, const fcl::CollisionResult<double>& result) {
  fcl::CollisionRequest<double> sub(request);
  sub.num_max_contacts = request.num_max_contacts - result.numContacts();
  return sub;
}
inline void append(const fcl::CollisionResult<double>& from, fcl::CollisionResult<double>& to) {
  for (std::size_t i = 0; i < from.numContacts(); ++i) to.addContact(from.getContact(i));
}
}  // namespace detail

// collision_matrix[BV_OBBRSS][BV_OBBRSS]
template <typename Solver>
std::size_t collide_cell(const fcl::CollisionGeometry<double>* o1, const fcl::Transform3<double>& tf1,
                         const fcl::CollisionGeometry<double>* o2, const fcl::Transform3<double>& tf2, const Solver*,
                         const fcl::CollisionRequest<double>& request, fcl::CollisionResult<double>& result) {
  if (request.isSatisfied(result)) return result.numContacts();  // collision_func_matrix-inl.h:580
  if (request.num_max_contacts <= result.numContacts()) return result.numContacts();
  const DeviceModel& m1 = model_cache().get(static_cast<const BVH*>(o1));
  const DeviceModel& m2 = model_cache().get(static_cast<const BVH*>(o2));
  std::vector<fcl::CollisionResult<double>> r;
  collide(m1, {tf1}, m2, {tf2}, detail::remaining(request, result), r);
  detail::append(r[0], result);
  return result.numContacts();
}

// collision_matrix[BV_OBBRSS][GEOM_SPHERE] (BVHShapeCollider<OBBRSS, Sphere>, collision_func_matrix-inl.h:378-430)
template <typename Solver>
std::size_t collide_sphere_cell(const fcl::CollisionGeometry<double>* o1, const fcl::Transform3<double>& tf1,
                                const fcl::CollisionGeometry<double>* o2, const fcl::Transform3<double>& tf2, const Solver*,
                                const fcl::CollisionRequest<double>& request, fcl::CollisionResult<double>& result) {
  if (request.isSatisfied(result)) return result.numContacts();  // collision_func_matrix-inl.h:389
  if (request.num_max_contacts <= result.numContacts()) return result.numContacts();
  const DeviceModel& m1 = model_cache().get(static_cast<const BVH*>(o1));
  std::vector<fcl::CollisionResult<double>> r;
  collide(m1, {tf1}, *static_cast<const fcl::Sphere<double>*>(o2), {tf2}, detail::remaining(request, result), r);
  detail::append(r[0], result);
  return result.numContacts();
}

// collision_matrix[BV_OBBRSS][GEOM_HALFSPACE] and [GEOM_PLANE] (BVHShapeCollider<OBBRSS, Halfspace | Plane>)
template <typename Solver, typename Shape>
std::size_t collide_plane_like_cell(const fcl::CollisionGeometry<double>* o1, const fcl::Transform3<double>& tf1,
                                    const fcl::CollisionGeometry<double>* o2, const fcl::Transform3<double>& tf2, const Solver*,
                                    const fcl::CollisionRequest<double>& request, fcl::CollisionResult<double>& result) {
  if (request.isSatisfied(result)) return result.numContacts();  // collision_func_matrix-inl.h:389
  if (request.num_max_contacts <= result.numContacts()) return result.numContacts();
  const DeviceModel& m1 = model_cache().get(static_cast<const BVH*>(o1));
  std::vector<fcl::CollisionResult<double>> r;
  collide(m1, {tf1}, *static_cast<const Shape*>(o2), {tf2}, detail::remaining(request, result), r);
  detail::append(r[0], result);
  return result.numContacts();
}
template <typename Solver>
std::size_t collide_halfspace_cell(const fcl::CollisionGeometry<double>* o1, const fcl::Transform3<double>& tf1,
                                   const fcl::CollisionGeometry<double>* o2, const fcl::Transform3<double>& tf2, const Solver* s,
                                   const fcl::CollisionRequest<double>& request, fcl::CollisionResult<double>& result) {
  return collide_plane_like_cell<Solver, fcl::Halfspace<double>>(o1, tf1, o2, tf2, s, request, result);
}
template <typename Solver>
std::size_t collide_plane_cell(const fcl::CollisionGeometry<double>* o1, const fcl::Transform3<double>& tf1,
                               const fcl::CollisionGeometry<double>* o2, const fcl::Transform3<double>& tf2, const Solver* s,
                               const fcl::CollisionRequest<double>& request, fcl::CollisionResult<double>& result) {
  return collide_plane_like_cell<Solver, fcl::Plane<double>>(o1, tf1, o2, tf2, s, request, result);
}

// distance_matrix[BV_OBBRSS][BV_OBBRSS]
template <typename Solver>
double distance_cell(const fcl::CollisionGeometry<double>* o1, const fcl::Transform3<double>& tf1,
                     const fcl::CollisionGeometry<double>* o2, const fcl::Transform3<double>& tf2, const Solver*,
                     const fcl::DistanceRequest<double>& request, fcl::DistanceResult<double>& result) {
  if (request.isSatisfied(result)) return result.min_distance;  // distance_func_matrix-inl.h:395
  const DeviceModel& m1 = model_cache().get(static_cast<const BVH*>(o1));
  const DeviceModel& m2 = model_cache().get(static_cast<const BVH*>(o2));
  std::vector<fcl::DistanceResult<double>> r;
  distance(m1, {tf1}, m2, {tf2}, request, r);
  if (request.enable_nearest_points)
    result.update(r[0].min_distance, o1, o2, r[0].b1, r[0].b2, r[0].nearest_points[0], r[0].nearest_points[1]);
  else
    result.update(r[0].min_distance, o1, o2, r[0].b1, r[0].b2);
  return result.min_distance;
}

// distance_matrix[BV_OBBRSS][GEOM_SPHERE] (BVHShapeDistancer<OBBRSS, Sphere>, distance_func_matrix-inl.h:259-277, 322-341)
template <typename Solver>
double distance_sphere_cell(const fcl::CollisionGeometry<double>* o1, const fcl::Transform3<double>& tf1,
                            const fcl::CollisionGeometry<double>* o2, const fcl::Transform3<double>& tf2, const Solver*,
                            const fcl::DistanceRequest<double>& request, fcl::DistanceResult<double>& result) {
  if (request.isSatisfied(result)) return result.min_distance;  // distance_func_matrix-inl.h:268
  const DeviceModel& m1 = model_cache().get(static_cast<const BVH*>(o1));
  std::vector<fcl::DistanceResult<double>> r;
  distance(m1, {tf1}, *static_cast<const fcl::Sphere<double>*>(o2), {tf2}, request, r);
  result.update(r[0].min_distance, o1, o2, r[0].b1, fcl::DistanceResult<double>::NONE, r[0].nearest_points[0], r[0].nearest_points[1]);
  return result.min_distance;
}

// What install() replaced, so that uninstall() can put it back.
template <typename Solver>
struct InstalledCells {
  typename fcl::detail::CollisionFunctionMatrix<Solver>::CollisionFunc collide_mesh = nullptr, collide_sphere = nullptr,
                                                                       collide_halfspace = nullptr, collide_plane = nullptr;
  typename fcl::detail::DistanceFunctionMatrix<Solver>::DistanceFunc distance_mesh = nullptr, distance_sphere = nullptr;
  bool installed = false;
  static InstalledCells& saved() {
    static InstalledCells s;
    return s;
  }
};

// Overwrites the six cells of the look-up tables of `Solver` (the public, non-const array members of the function-
// local statics behind fcl::getCollisionFunctionLookTable / getDistanceFunctionLookTable).  Call once at start-up,
// before other threads issue queries.  In a shared-library build of FCL with hidden visibility the library keeps its
// own copy of the tables for the extern-template collide<double>(o1, tf1, o2, tf2, request, result); the overwrite then
// reaches the calls instantiated in the caller's translation units (the solver-taking overloads), see INTEGRATION.md.
template <typename Solver>
void install() {
  auto& c = fcl::getCollisionFunctionLookTable<Solver>();
  auto& d = fcl::getDistanceFunctionLookTable<Solver>();
  auto& s = InstalledCells<Solver>::saved();
  if (!s.installed) {
    s.collide_mesh = c.collision_matrix[fcl::BV_OBBRSS][fcl::BV_OBBRSS];
    s.collide_sphere = c.collision_matrix[fcl::BV_OBBRSS][fcl::GEOM_SPHERE];
    s.collide_halfspace = c.collision_matrix[fcl::BV_OBBRSS][fcl::GEOM_HALFSPACE];
    s.collide_plane = c.collision_matrix[fcl::BV_OBBRSS][fcl::GEOM_PLANE];
    s.distance_mesh = d.distance_matrix[fcl::BV_OBBRSS][fcl::BV_OBBRSS];
    s.distance_sphere = d.distance_matrix[fcl::BV_OBBRSS][fcl::GEOM_SPHERE];
    s.installed = true;
  }
  c.collision_matrix[fcl::BV_OBBRSS][fcl::BV_OBBRSS] = &collide_cell<Solver>;
  c.collision_matrix[fcl::BV_OBBRSS][fcl::GEOM_SPHERE] = &collide_sphere_cell<Solver>;
  c.collision_matrix[fcl::BV_OBBRSS][fcl::GEOM_HALFSPACE] = &collide_halfspace_cell<Solver>;
  c.collision_matrix[fcl::BV_OBBRSS][fcl::GEOM_PLANE] = &collide_plane_cell<Solver>;
  d.distance_matrix[fcl::BV_OBBRSS][fcl::BV_OBBRSS] = &distance_cell<Solver>;
  d.distance_matrix[fcl::BV_OBBRSS][fcl::GEOM_SPHERE] = &distance_sphere_cell<Solver>;
}

template <typename Solver>
void uninstall() {
  auto& s = InstalledCells<Solver>::saved();
  if (!s.installed) return;
  auto& c = fcl::getCollisionFunctionLookTable<Solver>();
  auto& d = fcl::getDistanceFunctionLookTable<Solver>();
  c.collision_matrix[fcl::BV_OBBRSS][fcl::BV_OBBRSS] = s.collide_mesh;
  c.collision_matrix[fcl::BV_OBBRSS][fcl::GEOM_SPHERE] = s.collide_sphere;
  c.collision_matrix[fcl::BV_OBBRSS][fcl::GEOM_HALFSPACE] = s.collide_halfspace;
  c.collision_matrix[fcl::BV_OBBRSS][fcl::GEOM_PLANE] = s.collide_plane;
  d.distance_matrix[fcl::BV_OBBRSS][fcl::BV_OBBRSS] = s.distance_mesh;
  d.distance_matrix[fcl::BV_OBBRSS][fcl::GEOM_SPHERE] = s.distance_sphere;
  s.installed = false;
  model_cache().clear();
}

}  // namespace fclgpu
#endif  // FCLGPU_HAVE_FCL
