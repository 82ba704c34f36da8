// bvh_build.cpp — host-side OBBRSS BVH construction for the upload step.
//
// Mirrors what BVHModel<OBBRSS<double>>::endModel() produces so that a model built here and a
// model built by the reference have the same tree and the same bounding volumes
// (reference files, /root/reference/include/fcl/...):
//   geometry/bvh/BVH_model-inl.h:450-517, 833-938   endModel / buildTree / recursiveBuildTree
//   geometry/bvh/detail/BV_fitter-inl.h:449-477     FitImpl<OBBRSS>
//   geometry/bvh/detail/BV_splitter-inl.h:422-451, 492-500, 540-657  split rules
//   math/geometry-inl.h:294-362, 477-597, 709-988, 1335-1425  extent/centre, Jacobi, RSS fit, covariance
// Implementation notes: works on de-indexed triangles (9 doubles each), iterative with an
// explicit job stack (pre-order numbering: a node's two children get consecutive ids when the
// node is processed; the left subtree is finished before the right one starts).
// Arithmetic: separately rounded mul/add, three-term sums left to right (built with
// -ffp-contract=off), like the device code.
#include "bvh_build.hpp"
#include "bvh_fit.cuh"
#include "bvh_merge.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

#include "../../include/fclgpu.h"

namespace {

using fclgpu::NodeFit;

struct Builder {
  const double* tv;  // 9 doubles per triangle
  int split;
  std::vector<uint32_t> order;
  std::vector<double> med;

  const double* vert(uint32_t tri, int k) const { return tv + 9 * (size_t)tri + 3 * k; }

  void fit(const uint32_t* idx, int n, NodeFit& f) const { fclgpu::fit_obbrss(tv, 9, idx, n, f); }

  // split threshold along obb.axis.col(0) (BV_splitter-inl.h:540-657)
  double threshold(const NodeFit& f, const uint32_t* idx, int n) {
    const double sv[3] = {f.axis[0], f.axis[3], f.axis[6]};
    if (split == FCLGPU_SPLIT_METHOD_BV_CENTER) return f.obb_To[0];
    if (split == FCLGPU_SPLIT_METHOD_MEDIAN) {
      med.resize(n);
      for (int i = 0; i < n; ++i) {
        const double *p1 = vert(idx[i], 0), *p2 = vert(idx[i], 1), *p3 = vert(idx[i], 2);
        const double c0 = p1[0] + p2[0] + p3[0], c1 = p1[1] + p2[1] + p3[1], c2 = p1[2] + p2[2] + p3[2];
        med[i] = FCL_SUM3(c0 * sv[0], c1 * sv[1], c2 * sv[2]) / 3;
      }
      std::sort(med.begin(), med.end());
      return (n % 2 == 1) ? med[(n - 1) / 2] : (med[n / 2] + med[n / 2 - 1]) / 2;
    }
    double c[3] = {0.0, 0.0, 0.0};
    for (int i = 0; i < n; ++i) {
      const double *p1 = vert(idx[i], 0), *p2 = vert(idx[i], 1), *p3 = vert(idx[i], 2);
      c[0] += (p1[0] + p2[0] + p3[0]);
      c[1] += (p1[1] + p2[1] + p3[1]);
      c[2] += (p1[2] + p2[2] + p3[2]);
    }
    return (c[0] * sv[0] + c[1] * sv[1] + c[2] * sv[2]) / (3 * n);
  }
};

void store_fit(fclgpu_bvh* b, int node, const NodeFit& f) {
  std::memcpy(&b->axis[9 * (size_t)node], f.axis, sizeof f.axis);
  for (int k = 0; k < 3; ++k) {
    b->obb_To[3 * (size_t)node + k] = f.obb_To[k];
    b->obb_ext[3 * (size_t)node + k] = f.obb_ext[k];
    b->rss_To[3 * (size_t)node + k] = f.rss_To[k];
  }
  b->rss_l[2 * (size_t)node] = f.rss_l[0];
  b->rss_l[2 * (size_t)node + 1] = f.rss_l[1];
  b->rss_r[node] = f.rss_r;
}

void deindex(fclgpu_bvh* b, const double* vertices) {
  for (int t = 0; t < b->num_tris; ++t)
    for (int k = 0; k < 3; ++k)
      for (int c = 0; c < 3; ++c)
        b->tri_verts[9 * (size_t)t + 3 * k + c] = vertices[3 * (size_t)b->tri_index[3 * (size_t)t + k] + c];
}

}  // namespace

extern "C" int fclgpu_bvh_build_obbrss(const double* vertices, int32_t num_vertices, const int32_t* triangles,
                                       int32_t num_tris, int32_t split_method, fclgpu_bvh** out) {
  if (!out) return FCLGPU_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  if (num_tris <= 0 || num_vertices <= 0) return FCLGPU_ERR_BUILD_EMPTY_MODEL;  // BVH_model-inl.h:459-463
  if (!vertices || !triangles) return FCLGPU_ERR_INVALID_ARGUMENT;
  if (split_method < 0 || split_method > 2) return FCLGPU_ERR_UNSUPPORTED_FUNCTION;
  for (int64_t i = 0; i < 3 * (int64_t)num_tris; ++i)
    if (triangles[i] < 0 || triangles[i] >= num_vertices) return FCLGPU_ERR_INCORRECT_DATA;

  fclgpu_bvh* b = new fclgpu_bvh;
  b->num_tris = num_tris;
  b->num_vertices = num_vertices;
  b->split = split_method;
  const int nn = 2 * num_tris - 1;
  b->node_first.assign(nn, 0);
  b->node_count.assign(nn, 0);
  b->tri_index.assign(triangles, triangles + 3 * (size_t)num_tris);
  b->first_child.assign(nn, 0);
  b->axis.assign(9 * (size_t)nn, 0);
  b->obb_To.assign(3 * (size_t)nn, 0);
  b->obb_ext.assign(3 * (size_t)nn, 0);
  b->rss_To.assign(3 * (size_t)nn, 0);
  b->rss_l.assign(2 * (size_t)nn, 0);
  b->rss_r.assign(nn, 0);
  b->tri_verts.resize(9 * (size_t)num_tris);
  deindex(b, vertices);

  Builder B;
  B.tv = b->tri_verts.data();
  B.split = split_method;
  B.order.resize(num_tris);
  for (int i = 0; i < num_tris; ++i) B.order[i] = i;

  struct Job {
    int node, first, count;
  };
  std::vector<Job> jobs;
  jobs.push_back({0, 0, num_tris});
  int next_id = 1;
  while (!jobs.empty()) {
    const Job j = jobs.back();
    jobs.pop_back();
    uint32_t* idx = B.order.data() + j.first;
    NodeFit f;
    B.fit(idx, j.count, f);
    store_fit(b, j.node, f);
    b->node_first[j.node] = j.first;
    b->node_count[j.node] = j.count;
    if (j.count == 1) {
      b->first_child[j.node] = -((int32_t)idx[0] + 1);
      continue;
    }
    const double thr = B.threshold(f, idx, j.count);
    const double sv[3] = {f.axis[0], f.axis[3], f.axis[6]};
    const int left = next_id;
    next_id += 2;
    b->first_child[j.node] = left;
    // partition: centroids with sv.c <= thr go to the front, order otherwise preserved by swaps
    int c1 = 0;
    for (int i = 0; i < j.count; ++i) {
      const double *p1 = B.vert(idx[i], 0), *p2 = B.vert(idx[i], 1), *p3 = B.vert(idx[i], 2);
      const double cx = ((p1[0] + p2[0]) + p3[0]) / 3.0, cy = ((p1[1] + p2[1]) + p3[1]) / 3.0,
                   cz = ((p1[2] + p2[2]) + p3[2]) / 3.0;
      if (!(FCL_SUM3(sv[0] * cx, sv[1] * cy, sv[2] * cz) > thr)) {
        std::swap(idx[i], idx[c1]);
        c1++;
      }
    }
    if (c1 == 0 || c1 == j.count) c1 = j.count / 2;
    jobs.push_back({left + 1, j.first + c1, j.count - c1});  // right: processed after the whole left subtree
    jobs.push_back({left, j.first, c1});
  }
  b->prim_order = B.order;
  *out = b;
  return FCLGPU_OK;
}

// BVHModel::beginReplaceModel / replaceSubModel / endReplaceModel(refit = true, bottomup = false)
// (BVH_model-inl.h:521-620): same topology, every node refitted over its stored primitive range in
// primitive_indices order (refitTree_topdown, :1064-1076).
extern "C" int fclgpu_bvh_refit_topdown(fclgpu_bvh* b, const double* vertices, int32_t num_vertices) {
  if (!b || !vertices) return FCLGPU_ERR_INVALID_ARGUMENT;
  if (num_vertices != b->num_vertices) return FCLGPU_ERR_INCORRECT_DATA;  // :602-606
  deindex(b, vertices);
  b->rss_axis.clear();  // FitImpl<OBBRSS>: both volumes share the fitted axes again
  const int nn = 2 * b->num_tris - 1;
  for (int i = 0; i < nn; ++i) {
    NodeFit f;
    fclgpu::fit_obbrss(b->tri_verts.data(), 9, b->prim_order.data() + b->node_first[i], b->node_count[i], f);
    store_fit(b, i, f);
  }
  return FCLGPU_OK;
}

// endReplaceModel(refit = true, bottomup = true), the reference's default (BVH_model-inl.h:952-1037): children before
// parents, leaves by the closed-form triangle fit, inner nodes by merging the children's volumes (bvh_merge.cuh).
// Children always have larger ids than their parent (pre-order pair allocation), so one pass from the last node to the
// first visits every child before its parent.
extern "C" int fclgpu_bvh_refit_bottomup(fclgpu_bvh* b, const double* vertices, int32_t num_vertices) {
  if (!b || !vertices) return FCLGPU_ERR_INVALID_ARGUMENT;
  if (num_vertices != b->num_vertices) return FCLGPU_ERR_INCORRECT_DATA;  // :602-606
  deindex(b, vertices);
  const int nn = 2 * b->num_tris - 1;
  std::vector<fclgpu::NodeBV> bv((size_t)nn);
  for (int i = nn - 1; i >= 0; --i) {
    const int fc = b->first_child[i];
    if (fc < 0) fclgpu::fit3_obbrss(&b->tri_verts[9 * (size_t)(-(fc + 1))], bv[i]);
    else fclgpu::merge_obbrss(bv[fc], bv[fc + 1], bv[i]);
  }
  b->rss_axis.resize(9 * (size_t)nn);
  for (int i = 0; i < nn; ++i) {
    const fclgpu::NodeBV& f = bv[i];
    std::memcpy(&b->axis[9 * (size_t)i], f.axis, sizeof f.axis);
    std::memcpy(&b->rss_axis[9 * (size_t)i], f.rss_axis, sizeof f.rss_axis);
    for (int k = 0; k < 3; ++k) {
      b->obb_To[3 * (size_t)i + k] = f.obb_To[k];
      b->obb_ext[3 * (size_t)i + k] = f.obb_ext[k];
      b->rss_To[3 * (size_t)i + k] = f.rss_To[k];
    }
    b->rss_l[2 * (size_t)i] = f.rss_l[0];
    b->rss_l[2 * (size_t)i + 1] = f.rss_l[1];
    b->rss_r[i] = f.rss_r;
  }
  return FCLGPU_OK;
}

// rss.axis per node (9 doubles, row-major); returns 1 when the model has separate RSS axes (after a bottom-up refit),
// 0 when the RSS shares the OBB's axes (rss_axis9 is then filled with those)
extern "C" int fclgpu_bvh_get_rss_axis(const fclgpu_bvh* b, double* rss_axis9) {
  if (!b) return FCLGPU_ERR_INVALID_ARGUMENT;
  const std::vector<double>& src = b->rss_axis.empty() ? b->axis : b->rss_axis;
  if (rss_axis9) std::memcpy(rss_axis9, src.data(), src.size() * sizeof(double));
  return b->rss_axis.empty() ? 0 : 1;
}

extern "C" int fclgpu_bvh_get_partition(const fclgpu_bvh* b, int32_t* first_primitive, int32_t* num_primitives,
                                        int32_t* primitive_indices, int32_t* tri_indices3) {
  if (!b) return FCLGPU_ERR_INVALID_ARGUMENT;
  if (first_primitive) std::memcpy(first_primitive, b->node_first.data(), b->node_first.size() * sizeof(int32_t));
  if (num_primitives) std::memcpy(num_primitives, b->node_count.data(), b->node_count.size() * sizeof(int32_t));
  if (primitive_indices) std::memcpy(primitive_indices, b->prim_order.data(), b->prim_order.size() * sizeof(int32_t));
  if (tri_indices3) std::memcpy(tri_indices3, b->tri_index.data(), b->tri_index.size() * sizeof(int32_t));
  return FCLGPU_OK;
}

extern "C" int32_t fclgpu_bvh_num_vertices(const fclgpu_bvh* bvh) { return bvh ? bvh->num_vertices : 0; }

extern "C" void fclgpu_bvh_destroy(fclgpu_bvh* bvh) { delete bvh; }
extern "C" int32_t fclgpu_bvh_num_nodes(const fclgpu_bvh* bvh) { return bvh ? 2 * bvh->num_tris - 1 : 0; }
extern "C" int32_t fclgpu_bvh_num_tris(const fclgpu_bvh* bvh) { return bvh ? bvh->num_tris : 0; }

extern "C" int fclgpu_bvh_get(const fclgpu_bvh* b, int32_t* first_child, double* axis9, double* obb_To3,
                              double* obb_extent3, double* rss_To3, double* rss_l2, double* rss_r,
                              double* tri_verts9) {
  if (!b) return FCLGPU_ERR_INVALID_ARGUMENT;
  auto cp = [](auto* dst, const auto& src) {
    if (dst) std::memcpy(dst, src.data(), src.size() * sizeof(src[0]));
  };
  cp(first_child, b->first_child);
  cp(axis9, b->axis);
  cp(obb_To3, b->obb_To);
  cp(obb_extent3, b->obb_ext);
  cp(rss_To3, b->rss_To);
  cp(rss_l2, b->rss_l);
  cp(rss_r, b->rss_r);
  cp(tri_verts9, b->tri_verts);
  return FCLGPU_OK;
}
