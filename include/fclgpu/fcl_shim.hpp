// fcl_shim.hpp — the literal FCL drop-in above the C ABI (include/fclgpu.h).
//
// Compiled only when FCL and Eigen are available (neither is on the build image, so this header
// is dormant there; INTEGRATION.md shows how a maintainer wires it in).  It
//   * uploads an existing fcl::BVHModel<fcl::OBBRSS<double>> (no rebuild: the node tree FCL
//     built is flattened as is),
//   * evaluates batches of (tf1, tf2) pairs and fills std::vector<fcl::CollisionResult<double>>
//     / std::vector<fcl::DistanceResult<double>>,
//   * offers single-query functions with the exact signature of the reference's dispatch-table
//     cell (detail/collision_func_matrix.h:67-78, detail/distance_func_matrix.h:65-76) so they
//     can be installed in collision_matrix[BV_OBBRSS][BV_OBBRSS] / distance_matrix[..][..].
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
#include <memory>
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
    std::vector<double> axis(9 * n), oT(3 * n), oe(3 * n), rT(3 * n), rl(2 * n), rr(n), tv(9 * nt);
    for (int i = 0; i < n; ++i) {
      const auto& node = m.getBV(i);
      fc[i] = node.first_child;
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) axis[9 * i + 3 * r + c] = node.bv.obb.axis(r, c);  // rss.axis is identical
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
    check(fclgpu_model_create_obbrss(device, n, fc.data(), axis.data(), oT.data(), oe.data(), rT.data(), rl.data(),
                                     rr.data(), nt, tv.data(), &h_));
  }
  ~DeviceModel() { fclgpu_model_destroy(h_); }
  DeviceModel(const DeviceModel&) = delete;
  DeviceModel& operator=(const DeviceModel&) = delete;
  const fclgpu_model* handle() const { return h_; }
  const BVH* host() const { return host_; }

 private:
  const BVH* host_;
  fclgpu_model* h_ = nullptr;
};

// n independent fcl::collide(o1, tf1[i], o2, tf2[i], request, results[i]) calls
inline void collide(const DeviceModel& o1, const std::vector<fcl::Transform3<double>>& tf1, const DeviceModel& o2,
                    const std::vector<fcl::Transform3<double>>& tf2, const fcl::CollisionRequest<double>& request,
                    std::vector<fcl::CollisionResult<double>>& results) {
  const int64_t n = (int64_t)tf1.size();
  results.assign(n, fcl::CollisionResult<double>());
  if (request.num_max_contacts == 0 || n == 0) return;
  std::vector<double> p1(12 * n), p2(12 * n);
  for (int64_t i = 0; i < n; ++i) {
    to_pose(tf1[i], &p1[12 * i]);
    to_pose(tf2[i], &p2[12 * i]);
  }
  fclgpu_collision_request req{(int64_t)std::min<std::size_t>(request.num_max_contacts, (std::size_t)1 << 62),
                               request.enable_contact ? 1 : 0, request.enable_cost ? 1 : 0, 0};
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
  results.assign(n, fcl::CollisionResult<double>());
  if (request.num_max_contacts == 0 || n == 0) return;
  std::vector<double> p1(12 * n), p2(12 * n);
  for (int64_t i = 0; i < n; ++i) {
    to_pose(tf1[i], &p1[12 * i]);
    to_pose(tf2[i], &p2[12 * i]);
  }
  fclgpu_collision_request req{(int64_t)std::min<std::size_t>(request.num_max_contacts, (std::size_t)1 << 62),
                               request.enable_contact ? 1 : 0, request.enable_cost ? 1 : 0, 0};
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

inline void distance(const DeviceModel& o1, const std::vector<fcl::Transform3<double>>& tf1, const DeviceModel& o2,
                     const std::vector<fcl::Transform3<double>>& tf2, const fcl::DistanceRequest<double>& request,
                     std::vector<fcl::DistanceResult<double>>& results) {
  const int64_t n = (int64_t)tf1.size();
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
// fcl::distance(o1, tf1[i], o2, tf2[i]) <= tolerance; node pairs farther apart than the tolerance are never descended.
inline void within_tolerance(const DeviceModel& o1, const std::vector<fcl::Transform3<double>>& tf1, const DeviceModel& o2,
                             const std::vector<fcl::Transform3<double>>& tf2, double tolerance, std::vector<char>& within) {
  const int64_t n = (int64_t)tf1.size();
  within.assign(n, 0);
  if (n == 0) return;
  std::vector<double> p1(12 * n), p2(12 * n), d(n);
  for (int64_t i = 0; i < n; ++i) {
    to_pose(tf1[i], &p1[12 * i]);
    to_pose(tf2[i], &p2[12 * i]);
  }
  fclgpu_distance_request req{0, 0, 0.0, 0.0};
  check(fclgpu_distance_cutoff_batch_host(o1.handle(), o2.handle(), n, p1.data(), p2.data(), &req,
                                          std::nextafter(tolerance, std::numeric_limits<double>::infinity()), d.data(), nullptr,
                                          nullptr, nullptr, nullptr, nullptr, nullptr));
  for (int64_t i = 0; i < n; ++i) within[i] = d[i] <= tolerance ? 1 : 0;
}

// n independent fcl::distance(mesh, tf1[i], sphere, tf2[i], request, results[i]) calls
// (distance_matrix[BV_OBBRSS][GEOM_SPHERE]).  Like the reference's mesh-shape leaf, the nearest points are always
// handed to DistanceResult::update and stay in the local frames of the mesh and of the sphere; b2 = NONE.
inline void distance(const DeviceModel& o1, const std::vector<fcl::Transform3<double>>& tf1, const fcl::Sphere<double>& sphere,
                     const std::vector<fcl::Transform3<double>>& tf2, const fcl::DistanceRequest<double>& request,
                     std::vector<fcl::DistanceResult<double>>& results) {
  const int64_t n = (int64_t)tf1.size();
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

}  // namespace fclgpu
#endif  // FCLGPU_HAVE_FCL
