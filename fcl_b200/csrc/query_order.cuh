// query_order.cuh — locality order of a batch of queries.
//
// The queries of a batch are independent, so they may be EVALUATED in any order as long as every result lands at its
// query's own index.  A lane-per-query traversal runs at the speed of its warp: 32 lanes that walk unrelated parts of the
// two trees execute every branch of every lane and touch 32 different node records per load.  Queries whose relative
// poses are close walk nearly the same BVTT nodes, so the batch is bucketed by a 16-bit key of the relative pose --
// position of model 2's centre in model 1's frame on a 16 x 16 x 16 grid over the reachable cube (Morton-interleaved; a
// centre farther than r1 + r2 along an axis cannot collide and is clamped to the border cells), plus the octant of model
// 2's x axis and the sign of one more rotation entry -- with ONE counting-sort pass (histogram by atomics, scan of the
// 65 536 counters, scatter).  The order inside a bucket is whatever the atomics give: it never shows in a result.
// Nothing like it in the reference (it evaluates one query per call); this is batching glue, not collision code.
#pragma once
#include "traversal.cuh"

namespace fclgpu {

constexpr int kOrderBuckets = 1 << 16;

struct OrderParams {
  const double* tf1;
  const double* tf2;
  long long n;
  double c1[3], c2[3];  // local AABB centres of the two models (BVHModel::aabb_center)
  double reach;         // r1 + r2: farther apart along any axis and the pair cannot collide
  uint16_t* key;        // [n]
  uint32_t* hist;       // [kOrderBuckets], zeroed
  uint32_t* cursor;     // [kOrderBuckets]: exclusive prefix sums, consumed by the scatter
  int32_t* order;       // [n]
};

__device__ __forceinline__ unsigned spread4(unsigned v) {  // 4 bits -> every third bit
  return (v & 1u) | ((v & 2u) << 2) | ((v & 4u) << 4) | ((v & 8u) << 6);
}

__global__ void __launch_bounds__(256) order_key_kernel(OrderParams P) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= P.n) return;
  const PoseRT a = load_pose(P.tf1, q), b = load_pose(P.tf2, q);
  // centre of model 2 in model 1's frame, relative to model 1's centre
  const V3 w2 = mulv(b.R, mk(P.c2[0], P.c2[1], P.c2[2])) + b.t;
  const V3 d = mulTv(a.R, w2 - a.t) - mk(P.c1[0], P.c1[1], P.c1[2]);
  const double inv = P.reach > 0 ? 8.0 / P.reach : 0.0;  // [-reach, reach] -> [0, 16)
  auto cell = [&](double x) {
    const double c = x * inv + 8.0;
    return (unsigned)(c < 0.0 ? 0 : (c > 15.0 ? 15 : (int)c));
  };
  const unsigned m = spread4(cell(d.x)) | (spread4(cell(d.y)) << 1) | (spread4(cell(d.z)) << 2);  // 12 bits
  // x axis of model 2 seen from model 1 (first column of R1^T R2): its octant, and the sign of one entry of the y axis
  const M3 R = mulTM(a.R, b.R);
  const unsigned rot = (R.m[0] < 0 ? 1u : 0u) | (R.m[3] < 0 ? 2u : 0u) | (R.m[6] < 0 ? 4u : 0u) | (R.m[4] < 0 ? 8u : 0u);
  const unsigned key = (m << 4) | rot;
  P.key[q] = (uint16_t)key;
  atomicAdd(P.hist + key, 1u);
}

// exclusive scan of the 65 536 bucket counts: one block of 1024 threads, 64 buckets each
__global__ void __launch_bounds__(1024) order_scan_kernel(OrderParams P) {
  __shared__ uint32_t part[1024];
  const int t = threadIdx.x;
  uint32_t sum = 0;
  for (int k = 0; k < kOrderBuckets / 1024; ++k) sum += P.hist[t * (kOrderBuckets / 1024) + k];
  part[t] = sum;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const uint32_t v = t >= o ? part[t - o] : 0u;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  uint32_t run = part[t] - sum;
  for (int k = 0; k < kOrderBuckets / 1024; ++k) {
    const int i = t * (kOrderBuckets / 1024) + k;
    const uint32_t c = P.hist[i];
    P.cursor[i] = run;
    run += c;
  }
}

__global__ void __launch_bounds__(256) order_scatter_kernel(OrderParams P) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= P.n) return;
  const uint32_t pos = atomicAdd(P.cursor + P.key[q], 1u);
  P.order[pos] = (int32_t)q;
}

}  // namespace fclgpu
