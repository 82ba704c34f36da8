// broadphase.cuh — batched N x M broadphase feeding the mesh-mesh kernels.
//
// Reference semantics (file:line under /root/reference/include/fcl):
//   BVHModel::computeLocalAABB         geometry/bvh/BVH_model-inl.h:1080-1100   (aabb_local, aabb_center, aabb_radius)
//   CollisionObject::computeAABB       narrowphase/collision_object-inl.h:118-131
//       rotation == identity (Eigen isIdentity, precision 1e-12): aabb = aabb_local translated by t;
//       otherwise the cube of half side aabb_radius around tf * aabb_center
//   AABB::overlap                      math/bv/AABB-inl.h:98-107
//   NaiveCollisionManager::collide(other, cdata, callback)   broadphase/broadphase_bruteforce-inl.h:182-205
//       for obj1 in this (registration order) for obj2 in other: if the AABBs overlap -> callback(obj1, obj2)
//   DefaultCollisionFunction           broadphase/default_broadphase_callbacks.h:84-103   (fcl::collide on the pair)
// The DynamicAABBTreeCollisionManager reports the same SET of pairs (it is an acceleration structure over the same
// overlap test); its visiting order is an implementation detail, the brute-force manager's order is the documented one,
// and that is the order of the pair list produced here (i-major, j ascending).
//
// Device pipeline: world AABBs of both object sets -> per-object overlap counts (one warp per object of set 1, ballot +
// popc over 32 objects of set 2 at a time) -> exclusive scan -> pair list written at deterministic positions.  For the
// narrowphase the pairs are grouped by (geometry 1, geometry 2) with a stable device compaction per group, the poses
// are gathered on the device, the batched collide kernel runs per group, and the counts are scattered back to pair
// order: the pair list and the poses never visit the host between the two phases.
#pragma once
#include <cstdint>

#include "sum_order.h"

namespace fclgpu {

struct LocalAabb {  // BVHModel::computeLocalAABB: aabb_center, aabb_radius, aabb_local.min_, aabb_local.max_
  double c[3];
  double r;
  double mn[3], mx[3];
};

// Eigen::MatrixBase::isIdentity(prec = 1e-12) on the rotation part of a pose record (row-major R)
__host__ __device__ inline bool rotation_is_identity(const double* R) {
  const double prec = 1e-12;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      const double x = R[3 * i + j];
      if (i == j) {
        const double ax = x < 0 ? -x : x;
        const double d = x - 1.0 < 0 ? 1.0 - x : x - 1.0;
        if (!(d <= (ax < 1.0 ? ax : 1.0) * prec)) return false;  // !isApprox(x, 1, prec)
      } else {
        if (!((x < 0 ? -x : x) <= prec)) return false;           // !isMuchSmallerThan(x, 1, prec)
      }
    }
  return true;
}

__host__ __device__ inline void world_aabb(const LocalAabb& a, const double* tf, double* out6) {
  if (rotation_is_identity(tf)) {  // translate(aabb_local, t)
    for (int k = 0; k < 3; ++k) {
      out6[k] = a.mn[k] + tf[9 + k];
      out6[3 + k] = a.mx[k] + tf[9 + k];
    }
  } else {  // center = tf * aabb_center; min = center - radius, max = center + radius
    for (int k = 0; k < 3; ++k) {
      const double ck = FCL_SUM3(tf[3 * k] * a.c[0], tf[3 * k + 1] * a.c[1], tf[3 * k + 2] * a.c[2]) + tf[9 + k];
      out6[k] = ck - a.r;
      out6[3 + k] = ck + a.r;
    }
  }
}

__global__ void world_aabb_kernel(const LocalAabb* __restrict__ locals, const int32_t* __restrict__ geom,
                                  const double* __restrict__ tf, long long n, double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  world_aabb(locals[geom[i]], tf + 12 * i, out + 6 * i);
}

__device__ __forceinline__ bool aabb_overlap(const double* a, const double* b) {
  // !(min > other.max).any() && !(max < other.min).any()
  return !(a[0] > b[3] || a[1] > b[4] || a[2] > b[5]) && !(a[3] < b[0] || a[4] < b[1] || a[5] < b[2]);
}

// one warp per object of set 1; kWrite = false: counts[i] = number of overlapping objects of set 2;
// kWrite = true: pairs written at offsets[i].. in ascending j
template <bool kWrite>
__global__ void pair_kernel(const double* __restrict__ aabb1, long long n1, const double* __restrict__ aabb2, long long n2,
                            int32_t* __restrict__ counts, const long long* __restrict__ local, const long long* __restrict__ block_sums,
                            int scan_block, int2* __restrict__ pairs, long long capacity) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n1) return;
  double a[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) a[k] = aabb1[6 * i + k];
  long long base = kWrite ? block_sums[i / scan_block] + local[i] : 0;
  int total = 0;
  for (long long j0 = 0; j0 < n2; j0 += 32) {
    const long long j = j0 + lane;
    bool hit = false;
    if (j < n2) {
      double b[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) b[k] = aabb2[6 * j + k];
      hit = aabb_overlap(a, b);
    }
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (kWrite && hit) {
      const long long pos = base + __popc(m & ((1u << lane) - 1u));
      if (pos < capacity) pairs[pos] = make_int2((int)i, (int)j);
    }
    base += __popc(m);
    total += __popc(m);
  }
  if (!kWrite && lane == 0) counts[i] = total;
}

// group selection: flag[k] = pair k is a (ga, gb) pair
__global__ void group_flag_kernel(const int2* __restrict__ pairs, long long n, const int32_t* __restrict__ geom1,
                                  const int32_t* __restrict__ geom2, int ga, int gb, int32_t* __restrict__ flag) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int2 p = pairs[k];
  flag[k] = (geom1[p.x] == ga && geom2[p.y] == gb) ? 1 : 0;
}

// stable gather of the group's poses (and the pair index each compacted slot came from)
__global__ void group_gather_kernel(const int2* __restrict__ pairs, long long n, const int32_t* __restrict__ flag,
                                    const long long* __restrict__ local, const long long* __restrict__ block_sums, int scan_block,
                                    const double* __restrict__ tf1, const double* __restrict__ tf2, double* __restrict__ gtf1,
                                    double* __restrict__ gtf2, long long* __restrict__ gidx) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n || !flag[k]) return;
  const long long dst = block_sums[k / scan_block] + local[k];
  const int2 p = pairs[k];
#pragma unroll
  for (int c = 0; c < 12; ++c) {
    gtf1[12 * dst + c] = tf1[12 * (long long)p.x + c];
    gtf2[12 * dst + c] = tf2[12 * (long long)p.y + c];
  }
  gidx[dst] = k;
}

__global__ void group_scatter_kernel(const long long* __restrict__ gidx, const int32_t* __restrict__ gcounts, long long m,
                                     int32_t* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < m) out[gidx[t]] = gcounts[t];
}

}  // namespace fclgpu
