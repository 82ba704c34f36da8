// bvh_build.hpp — host-side construction of the flattened OBBRSS node tree
// (the product's counterpart of BVHModel<OBBRSS<double>>::endModel()).
#pragma once
#include <cstdint>
#include <vector>

struct fclgpu_bvh {
  int32_t num_tris = 0, num_vertices = 0, split = 0;
  std::vector<int32_t> first_child;  // per node
  std::vector<double> axis;          // 9 per node, row-major (column c = c-th box axis)
  std::vector<double> obb_To, obb_ext, rss_To;  // 3 per node
  std::vector<double> rss_l;         // 2 per node
  std::vector<double> rss_r;         // 1 per node
  std::vector<double> rss_axis;      // 9 per node after a bottom-up refit (the RSS no longer shares the OBB's axes); else empty
  std::vector<double> tri_verts;     // 9 per triangle (de-indexed)
  // what a top-down refit needs (BVHModel::primitive_indices, BVNodeBase::first_primitive / num_primitives)
  std::vector<int32_t> tri_index;        // 3 per triangle (vertex ids)
  std::vector<int32_t> node_first, node_count;  // per node: range in prim_order
  std::vector<uint32_t> prim_order;      // primitive_indices after the build
};
