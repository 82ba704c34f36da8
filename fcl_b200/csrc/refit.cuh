// refit.cuh — on-device top-down refit of an OBBRSS BVH.
//
// Reference semantics: BVHModel::endReplaceModel(refit = true, bottomup = false) ->
// refitTree_topdown (include/fcl/geometry/bvh/BVH_model-inl.h:594-620, 1064-1076): the tree
// keeps its shape; every node's OBBRSS is fitted again over the node's primitive range in
// primitive_indices order (FitImpl<OBBRSS>, detail/BV_fitter-inl.h:449-477).
//
// Kernels: gather_tris (vertices -> de-indexed 80-byte triangle records), refit_nodes (one
// thread per node runs fit_obbrss sequentially, so every sum has the reference's order and
// the BVs are bit-identical to a CPU refit; nodes are processed largest first), and the same
// thread repacks the node's FP64 records, its single-precision steering records and `size`.
// The work of a node is proportional to its triangle count, so the launch time is set by the
// root (n triangles, sequential); that is the price of bit-exact sums and is still several
// times faster than the host refit and needs no PCIe round trip.
#pragma once
#include "bvh_fit.cuh"
#include "bvh_merge.cuh"
#include "records.hpp"
#include "traversal.cuh"

namespace fclgpu {

struct RefitParams {
  double* obb;
  double* rss;
  double* tri;  // 10 doubles per triangle
  RssRec32* rss32;
  ObbRec32* obb32;
  double2* topo;
  const int32_t* tri_index;   // 3 per triangle
  const int32_t* node_first;  // per node
  const int32_t* node_count;
  const uint32_t* prim_order;
  const int32_t* by_size;     // node ids sorted by decreasing primitive count
  int32_t n_nodes, n_tris;
};

__global__ void gather_tris_kernel(RefitParams P, const double* __restrict__ vertices) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P.n_tris) return;
  double* o = P.tri + (size_t)t * kTriDoubles;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double* v = vertices + 3 * (size_t)P.tri_index[3 * (size_t)t + k];
    o[3 * k] = v[0];
    o[3 * k + 1] = v[1];
    o[3 * k + 2] = v[2];
  }
  o[9] = 0.0;
}

__global__ void __launch_bounds__(64) refit_nodes_kernel(RefitParams P) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= P.n_nodes) return;
  const int node = P.by_size[k];
  NodeFit f;
  fit_obbrss(P.tri, kTriDoubles, P.prim_order + P.node_first[node], P.node_count[node], f);
  double* o = P.obb + (size_t)node * kNodeDoubles;
  double* r = P.rss + (size_t)node * kNodeDoubles;
#pragma unroll
  for (int i = 0; i < 9; ++i) o[i] = r[i] = f.axis[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    o[9 + i] = f.obb_To[i];
    o[12 + i] = f.obb_ext[i];
    r[9 + i] = f.rss_To[i];
  }
  r[12] = f.rss_l[0];
  r[13] = f.rss_l[1];
  r[14] = f.rss_r;
  const double size = FCL_SUM3(f.obb_ext[0] * f.obb_ext[0], f.obb_ext[1] * f.obb_ext[1], f.obb_ext[2] * f.obb_ext[2]);  // extent.squaredNorm()
  o[15] = r[15] = size;
  RssRec32 r32;
  pack_rss32(f.axis, f.rss_To, f.rss_l, f.rss_r, r32);
  P.rss32[node] = r32;
  ObbRec32 o32;
  pack_obb32(f.axis, f.obb_To, f.obb_ext, o32);
  P.obb32[node] = o32;
  double2 t = P.topo[node];
  t.y = size;
  P.topo[node] = t;
}

}  // namespace fclgpu

namespace fclgpu {

// ---------------------------------------------------------------------------------------
// Warp-cooperative fit for large nodes.  Bit-identical to fit_obbrss:
//   * covariance: the 32 lanes compute the per-triangle terms of a chunk in parallel, then
//     nine lanes (one per accumulator) add the 32 terms IN ORDER, so each running sum sees
//     exactly the sequence of additions of the sequential loop;
//   * extents, z range, extreme points (first index wins ties), rectangle growth: min / max /
//     arg-min reductions, which are exact and order independent (the reference's "skip points
//     inside the current bound" tests only skip candidates that cannot change the result);
//   * corner growth: the rectangle only grows, so only points outside the INITIAL rectangle in
//     both coordinates can ever trigger; they are found in parallel and replayed in order.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double w = __shfl_xor_sync(0xffffffffu, v, o);
    v = (w < v) ? w : v;
  }
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double w = __shfl_xor_sync(0xffffffffu, v, o);
    v = (w > v) ? w : v;
  }
  return v;
}
// (value, index) with the smallest value; ties -> smallest index.  sign = -1 turns it into arg-max.
__device__ __forceinline__ void warp_argmin(double& v, int& idx) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double w = __shfl_xor_sync(0xffffffffu, v, o);
    const int wi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (w < v || (w == v && wi < idx)) {
      v = w;
      idx = wi;
    }
  }
}

// One step of the corner growth of getRadiusAndOriginAndRectangleSize (math/geometry-inl.h:893-968) for a
// point (qx, qy, qz) in the box frame: a point beyond a corner of the rectangle pushes that corner outwards.
__device__ __forceinline__ void rss_corner_grow(double qx, double qy, double qz, double cz, double radsqr, double& minx,
                                                double& maxx, double& miny, double& maxy) {
  const double a = sqrt(0.5);
  double dx, dy, u, t;
  if (qx > maxx) {
    if (qy > maxy) {
      dx = qx - maxx; dy = qy - maxy;
      u = dx * a + dy * a;
      t = (a * u - dx) * (a * u - dx) + (a * u - dy) * (a * u - dy) + (cz - qz) * (cz - qz);
      u = u - sqrt(fmax(radsqr - t, 0.0));
      if (u > 0) { maxx += u * a; maxy += u * a; }
    } else if (qy < miny) {
      dx = qx - maxx; dy = qy - miny;
      u = dx * a - dy * a;
      t = (a * u - dx) * (a * u - dx) + (-a * u - dy) * (-a * u - dy) + (cz - qz) * (cz - qz);
      u = u - sqrt(fmax(radsqr - t, 0.0));
      if (u > 0) { maxx += u * a; miny -= u * a; }
    }
  } else if (qx < minx) {
    if (qy > maxy) {
      dx = qx - minx; dy = qy - maxy;
      u = dy * a - dx * a;
      t = (-a * u - dx) * (-a * u - dx) + (a * u - dy) * (a * u - dy) + (cz - qz) * (cz - qz);
      u = u - sqrt(fmax(radsqr - t, 0.0));
      if (u > 0) { minx -= u * a; maxy += u * a; }
    } else if (qy < miny) {
      dx = qx - minx; dy = qy - miny;
      u = -dx * a - dy * a;
      t = (-a * u - dx) * (-a * u - dx) + (-a * u - dy) * (-a * u - dy) + (cz - qz) * (cz - qz);
      u = u - sqrt(fmax(radsqr - t, 0.0));
      if (u > 0) { minx -= u * a; miny -= u * a; }
    }
  }
}

__device__ inline void fit_obbrss_warp(const double* __restrict__ tv, int tri_stride, const uint32_t* __restrict__ idx,
                                       int n, double (*terms)[9], NodeFit& f) {
  const int lane = threadIdx.x & 31;
  // --- covariance ---
  double acc = 0.0;  // lane k < 9 owns accumulator k: S1[0..2], S2 xx yy zz xy xz yz
  for (int base = 0; base < n; base += 32) {
    const int i = base + lane;
    if (i < n) {
      const double* p1 = tv + (size_t)idx[i] * tri_stride;
      const double* p2 = p1 + 3;
      const double* p3 = p1 + 6;
      double t[9];
      for (int k = 0; k < 3; ++k) t[k] = ((p1[k] + p2[k]) + p3[k]);
      t[3] = (p1[0] * p1[0] + p2[0] * p2[0] + p3[0] * p3[0]);
      t[4] = (p1[1] * p1[1] + p2[1] * p2[1] + p3[1] * p3[1]);
      t[5] = (p1[2] * p1[2] + p2[2] * p2[2] + p3[2] * p3[2]);
      t[6] = (p1[0] * p1[1] + p2[0] * p2[1] + p3[0] * p3[1]);
      t[7] = (p1[0] * p1[2] + p2[0] * p2[2] + p3[0] * p3[2]);
      t[8] = (p1[1] * p1[2] + p2[1] * p2[2] + p3[1] * p3[2]);
#pragma unroll
      for (int k = 0; k < 9; ++k) terms[lane][k] = t[k];
    }
    __syncwarp();
    if (lane < 9) {
      const int cnt = (n - base) < 32 ? (n - base) : 32;
      for (int r = 0; r < cnt; ++r) acc += terms[r][lane];
    }
    __syncwarp();
  }
  double S1[3], S2[6];
#pragma unroll
  for (int k = 0; k < 3; ++k) S1[k] = __shfl_sync(0xffffffffu, acc, k);
#pragma unroll
  for (int k = 0; k < 6; ++k) S2[k] = __shfl_sync(0xffffffffu, acc, 3 + k);
  const int np = 3 * n;
  f.vsum[0] = S1[0];
  f.vsum[1] = S1[1];
  f.vsum[2] = S1[2];
  double M[3][3];
  M[0][0] = S2[0] - S1[0] * S1[0] / np;
  M[1][1] = S2[1] - S1[1] * S1[1] / np;
  M[2][2] = S2[2] - S1[2] * S1[2] / np;
  M[0][1] = M[1][0] = S2[3] - S1[0] * S1[1] / np;
  M[1][2] = M[2][1] = S2[5] - S1[1] * S1[2] / np;
  M[0][2] = M[2][0] = S2[4] - S1[0] * S1[2] / np;
  double ev[3], V[3][3];
  jacobi3(M, ev, V);
  int lo, mid, hi;
  if (ev[0] > ev[1]) { hi = 0; lo = 1; } else { lo = 0; hi = 1; }
  if (ev[2] < ev[lo]) { mid = lo; lo = 2; }
  else if (ev[2] > ev[hi]) { mid = hi; hi = 2; }
  else mid = 2;
  double* A = f.axis;
  for (int r = 0; r < 3; ++r) {
    A[3 * r + 0] = V[r][hi];
    A[3 * r + 1] = V[r][mid];
  }
  A[2] = A[3] * A[7] - A[6] * A[4];
  A[5] = A[6] * A[1] - A[0] * A[7];
  A[8] = A[0] * A[4] - A[3] * A[1];
  const double a00 = A[0], a10 = A[3], a20 = A[6], a01 = A[1], a11 = A[4], a21 = A[7], a02 = A[2], a12 = A[5], a22 = A[8];
  const int m = 3 * n;
#define W_PT(j) (tv + (size_t)idx[(j) / 3] * tri_stride + 3 * ((j) % 3))
#define W_PX(p) FCL_SUM3(a00 * (p)[0], a10 * (p)[1], a20 * (p)[2])
#define W_PY(p) FCL_SUM3(a01 * (p)[0], a11 * (p)[1], a21 * (p)[2])
#define W_PZ(p) FCL_SUM3(a02 * (p)[0], a12 * (p)[1], a22 * (p)[2])
  // --- extents (min / max of the projections) and extreme points along x and y (first index wins) ---
  double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
  double vminx = DBL_MAX, vmaxx = DBL_MAX, vminy = DBL_MAX, vmaxy = DBL_MAX;  // arg-max tracked on the negated value
  int iminx = 0x7fffffff, imaxx = 0x7fffffff, iminy = 0x7fffffff, imaxy = 0x7fffffff;
  for (int j = lane; j < m; j += 32) {
    const double* p = W_PT(j);
    const double c0 = W_PX(p), c1 = W_PY(p), c2 = W_PZ(p);
    if (c0 > mx[0]) mx[0] = c0;
    if (c0 < mn[0]) mn[0] = c0;
    if (c1 > mx[1]) mx[1] = c1;
    if (c1 < mn[1]) mn[1] = c1;
    if (c2 > mx[2]) mx[2] = c2;
    if (c2 < mn[2]) mn[2] = c2;
    if (c0 < vminx) { vminx = c0; iminx = j; }
    if (-c0 < vmaxx) { vmaxx = -c0; imaxx = j; }
    if (c1 < vminy) { vminy = c1; iminy = j; }
    if (-c1 < vmaxy) { vmaxy = -c1; imaxy = j; }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    mn[k] = warp_min(mn[k]);
    mx[k] = warp_max(mx[k]);
  }
  warp_argmin(vminx, iminx);
  warp_argmin(vmaxx, imaxx);
  warp_argmin(vminy, iminy);
  warp_argmin(vmaxy, imaxy);
  const double o[3] = {(mx[0] + mn[0]) / 2, (mx[1] + mn[1]) / 2, (mx[2] + mn[2]) / 2};
  for (int r = 0; r < 3; ++r) {
    f.obb_To[r] = FCL_SUM3(A[3 * r] * o[0], A[3 * r + 1] * o[1], A[3 * r + 2] * o[2]);
    f.obb_ext[r] = (mx[r] - mn[r]) / 2;
  }
  const double minz = mn[2], maxz = mx[2];
  const double r = 0.5 * (maxz - minz), radsqr = r * r, cz = 0.5 * (maxz + minz);
  // initial rectangle from the extreme points
  double minx, maxx, miny, maxy;
  {
    const double* p = W_PT(iminx);
    double dz = W_PZ(p) - cz;
    minx = W_PX(p) + sqrt(fmax(radsqr - dz * dz, 0.0));
    p = W_PT(imaxx);
    dz = W_PZ(p) - cz;
    maxx = W_PX(p) - sqrt(fmax(radsqr - dz * dz, 0.0));
    p = W_PT(iminy);
    dz = W_PZ(p) - cz;
    miny = W_PY(p) + sqrt(fmax(radsqr - dz * dz, 0.0));
    p = W_PT(imaxy);
    dz = W_PZ(p) - cz;
    maxy = W_PY(p) - sqrt(fmax(radsqr - dz * dz, 0.0));
  }
  // growth: lo = min(lo, val + reach) over points with val < lo ; hi = max(hi, val - reach) over val > hi
  {
    double lx = minx, hx = maxx, ly = miny, hy = maxy;
    for (int j = lane; j < m; j += 32) {
      const double* p = W_PT(j);
      const double px = W_PX(p), py = W_PY(p);
      if (px < minx || px > maxx || py < miny || py > maxy) {
        const double dz = W_PZ(p) - cz;
        const double reach = sqrt(fmax(radsqr - dz * dz, 0.0));
        if (px < minx) { const double x = px + reach; if (x < lx) lx = x; }
        if (px > maxx) { const double x = px - reach; if (x > hx) hx = x; }
        if (py < miny) { const double y = py + reach; if (y < ly) ly = y; }
        if (py > maxy) { const double y = py - reach; if (y > hy) hy = y; }
      }
    }
    // NOTE: the reference grows minx first and then tests maxx against the ORIGINAL maxx etc.; each of the
    // four bounds depends only on its own initial value, so they can be reduced together.
    minx = warp_min(lx);
    maxx = warp_max(hx);
    miny = warp_min(ly);
    maxy = warp_max(hy);
  }
  // corner growth, replayed in point order over the candidates
  const double minx0 = minx, maxx0 = maxx, miny0 = miny, maxy0 = maxy;
  for (int base = 0; base < m; base += 32) {
    const int j = base + lane;
    double px = 0, py = 0, pz = 0;
    bool cand = false;
    if (j < m) {
      const double* p = W_PT(j);
      px = W_PX(p); py = W_PY(p); pz = W_PZ(p);
      cand = (px > maxx0 || px < minx0) && (py > maxy0 || py < miny0);
    }
    unsigned mask = __ballot_sync(0xffffffffu, cand);
    while (mask) {
      const int src = __ffs(mask) - 1;
      mask &= mask - 1;
      const double qx = __shfl_sync(0xffffffffu, px, src), qy = __shfl_sync(0xffffffffu, py, src), qz = __shfl_sync(0xffffffffu, pz, src);
      rss_corner_grow(qx, qy, qz, cz, radsqr, minx, maxx, miny, maxy);
    }
  }
  for (int k = 0; k < 3; ++k) f.rss_To[k] = (A[3 * k] * minx + A[3 * k + 1] * miny) + A[3 * k + 2] * cz;
  f.rss_l[0] = maxx - minx;
  if (f.rss_l[0] < 0) f.rss_l[0] = 0;
  f.rss_l[1] = maxy - miny;
  if (f.rss_l[1] < 0) f.rss_l[1] = 0;
  f.rss_r = r;
#undef W_PT
#undef W_PX
#undef W_PY
#undef W_PZ
}


// ---------------------------------------------------------------------------------------
// Block-cooperative fit for the few huge nodes near the root (a warp would spend tens of
// milliseconds on a 200k-triangle root).  Same construction as the warp version: parallel
// terms + in-order accumulation for the covariance, exact min / max / arg-min reductions, and
// the corner growth replayed in point order over the (few) candidate points.
// ---------------------------------------------------------------------------------------
constexpr int kFitBlock = 256;
constexpr int kCornerCap = 1024;
struct BlockFitSmem {
  double terms[kFitBlock][9];
  double red[kFitBlock / 32][14];
  double bc[16];
  int redi[kFitBlock / 32][4];
  int cand[kCornerCap], sorted[kCornerCap];
  int wcnt[kFitBlock / 32];
  int ncand;
};

__device__ inline void fit_obbrss_block(const double* __restrict__ tv, int tri_stride, const uint32_t* __restrict__ idx,
                                        int n, BlockFitSmem& sm, NodeFit& f) {
  constexpr int NW = kFitBlock / 32;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) sm.ncand = 0;
  // --- covariance ---
  double acc = 0.0;  // thread k < 9 owns accumulator k
  for (int base = 0; base < n; base += kFitBlock) {
    const int i = base + tid;
    if (i < n) {
      const double* p1 = tv + (size_t)idx[i] * tri_stride;
      const double* p2 = p1 + 3;
      const double* p3 = p1 + 6;
      double* t = sm.terms[tid];
      for (int k = 0; k < 3; ++k) t[k] = ((p1[k] + p2[k]) + p3[k]);
      t[3] = (p1[0] * p1[0] + p2[0] * p2[0] + p3[0] * p3[0]);
      t[4] = (p1[1] * p1[1] + p2[1] * p2[1] + p3[1] * p3[1]);
      t[5] = (p1[2] * p1[2] + p2[2] * p2[2] + p3[2] * p3[2]);
      t[6] = (p1[0] * p1[1] + p2[0] * p2[1] + p3[0] * p3[1]);
      t[7] = (p1[0] * p1[2] + p2[0] * p2[2] + p3[0] * p3[2]);
      t[8] = (p1[1] * p1[2] + p2[1] * p2[2] + p3[1] * p3[2]);
    }
    __syncthreads();
    if (tid < 9) {
      const int cnt = (n - base) < kFitBlock ? (n - base) : kFitBlock;
#pragma unroll 8
      for (int r = 0; r < cnt; ++r) acc += sm.terms[r][tid];
    }
    __syncthreads();
  }
  if (tid < 9) sm.bc[tid] = acc;
  __syncthreads();
  double S1[3], S2[6];
  for (int k = 0; k < 3; ++k) S1[k] = sm.bc[k];
  for (int k = 0; k < 6; ++k) S2[k] = sm.bc[3 + k];
  const int np = 3 * n;
  f.vsum[0] = S1[0];
  f.vsum[1] = S1[1];
  f.vsum[2] = S1[2];
  double M[3][3];
  M[0][0] = S2[0] - S1[0] * S1[0] / np;
  M[1][1] = S2[1] - S1[1] * S1[1] / np;
  M[2][2] = S2[2] - S1[2] * S1[2] / np;
  M[0][1] = M[1][0] = S2[3] - S1[0] * S1[1] / np;
  M[1][2] = M[2][1] = S2[5] - S1[1] * S1[2] / np;
  M[0][2] = M[2][0] = S2[4] - S1[0] * S1[2] / np;
  double ev[3], V[3][3];
  jacobi3(M, ev, V);
  int lo, mid, hi;
  if (ev[0] > ev[1]) { hi = 0; lo = 1; } else { lo = 0; hi = 1; }
  if (ev[2] < ev[lo]) { mid = lo; lo = 2; }
  else if (ev[2] > ev[hi]) { mid = hi; hi = 2; }
  else mid = 2;
  double* A = f.axis;
  for (int r = 0; r < 3; ++r) {
    A[3 * r + 0] = V[r][hi];
    A[3 * r + 1] = V[r][mid];
  }
  A[2] = A[3] * A[7] - A[6] * A[4];
  A[5] = A[6] * A[1] - A[0] * A[7];
  A[8] = A[0] * A[4] - A[3] * A[1];
  const double a00 = A[0], a10 = A[3], a20 = A[6], a01 = A[1], a11 = A[4], a21 = A[7], a02 = A[2], a12 = A[5], a22 = A[8];
  const int m = 3 * n;
#define W_PT(j) (tv + (size_t)idx[(j) / 3] * tri_stride + 3 * ((j) % 3))
#define W_PX(p) FCL_SUM3(a00 * (p)[0], a10 * (p)[1], a20 * (p)[2])
#define W_PY(p) FCL_SUM3(a01 * (p)[0], a11 * (p)[1], a21 * (p)[2])
#define W_PZ(p) FCL_SUM3(a02 * (p)[0], a12 * (p)[1], a22 * (p)[2])
  // --- extents and extreme points ---
  double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
  double vminx = DBL_MAX, vmaxx = DBL_MAX, vminy = DBL_MAX, vmaxy = DBL_MAX;
  int iminx = 0x7fffffff, imaxx = 0x7fffffff, iminy = 0x7fffffff, imaxy = 0x7fffffff;
  for (int j = tid; j < m; j += kFitBlock) {
    const double* p = W_PT(j);
    const double c0 = W_PX(p), c1 = W_PY(p), c2 = W_PZ(p);
    if (c0 > mx[0]) mx[0] = c0;
    if (c0 < mn[0]) mn[0] = c0;
    if (c1 > mx[1]) mx[1] = c1;
    if (c1 < mn[1]) mn[1] = c1;
    if (c2 > mx[2]) mx[2] = c2;
    if (c2 < mn[2]) mn[2] = c2;
    if (c0 < vminx) { vminx = c0; iminx = j; }
    if (-c0 < vmaxx) { vmaxx = -c0; imaxx = j; }
    if (c1 < vminy) { vminy = c1; iminy = j; }
    if (-c1 < vmaxy) { vmaxy = -c1; imaxy = j; }
  }
  for (int k = 0; k < 3; ++k) {
    mn[k] = warp_min(mn[k]);
    mx[k] = warp_max(mx[k]);
  }
  warp_argmin(vminx, iminx);
  warp_argmin(vmaxx, imaxx);
  warp_argmin(vminy, iminy);
  warp_argmin(vmaxy, imaxy);
  if (lane == 0) {
    for (int k = 0; k < 3; ++k) {
      sm.red[wid][k] = mn[k];
      sm.red[wid][3 + k] = mx[k];
    }
    sm.red[wid][6] = vminx; sm.red[wid][7] = vmaxx; sm.red[wid][8] = vminy; sm.red[wid][9] = vmaxy;
    sm.redi[wid][0] = iminx; sm.redi[wid][1] = imaxx; sm.redi[wid][2] = iminy; sm.redi[wid][3] = imaxy;
  }
  __syncthreads();
  for (int k = 0; k < 3; ++k) {
    mn[k] = sm.red[0][k];
    mx[k] = sm.red[0][3 + k];
  }
  double av[4];
  int ai[4];
  for (int q = 0; q < 4; ++q) {
    av[q] = sm.red[0][6 + q];
    ai[q] = sm.redi[0][q];
  }
  for (int w = 1; w < NW; ++w) {
    for (int k = 0; k < 3; ++k) {
      const double a = sm.red[w][k], b = sm.red[w][3 + k];
      if (a < mn[k]) mn[k] = a;
      if (b > mx[k]) mx[k] = b;
    }
    for (int q = 0; q < 4; ++q) {
      const double v = sm.red[w][6 + q];
      const int vi = sm.redi[w][q];
      if (v < av[q] || (v == av[q] && vi < ai[q])) {
        av[q] = v;
        ai[q] = vi;
      }
    }
  }
  iminx = ai[0]; imaxx = ai[1]; iminy = ai[2]; imaxy = ai[3];
  const double o[3] = {(mx[0] + mn[0]) / 2, (mx[1] + mn[1]) / 2, (mx[2] + mn[2]) / 2};
  for (int r = 0; r < 3; ++r) {
    f.obb_To[r] = FCL_SUM3(A[3 * r] * o[0], A[3 * r + 1] * o[1], A[3 * r + 2] * o[2]);
    f.obb_ext[r] = (mx[r] - mn[r]) / 2;
  }
  const double minz = mn[2], maxz = mx[2];
  const double r = 0.5 * (maxz - minz), radsqr = r * r, cz = 0.5 * (maxz + minz);
  double minx, maxx, miny, maxy;
  {
    const double* p = W_PT(iminx);
    double dz = W_PZ(p) - cz;
    minx = W_PX(p) + sqrt(fmax(radsqr - dz * dz, 0.0));
    p = W_PT(imaxx);
    dz = W_PZ(p) - cz;
    maxx = W_PX(p) - sqrt(fmax(radsqr - dz * dz, 0.0));
    p = W_PT(iminy);
    dz = W_PZ(p) - cz;
    miny = W_PY(p) + sqrt(fmax(radsqr - dz * dz, 0.0));
    p = W_PT(imaxy);
    dz = W_PZ(p) - cz;
    maxy = W_PY(p) - sqrt(fmax(radsqr - dz * dz, 0.0));
  }
  {
    double lx = minx, hx = maxx, ly = miny, hy = maxy;
    for (int j = tid; j < m; j += kFitBlock) {
      const double* p = W_PT(j);
      const double px = W_PX(p), py = W_PY(p);
      if (px < minx || px > maxx || py < miny || py > maxy) {
        const double dz = W_PZ(p) - cz;
        const double reach = sqrt(fmax(radsqr - dz * dz, 0.0));
        if (px < minx) { const double x = px + reach; if (x < lx) lx = x; }
        if (px > maxx) { const double x = px - reach; if (x > hx) hx = x; }
        if (py < miny) { const double y = py + reach; if (y < ly) ly = y; }
        if (py > maxy) { const double y = py - reach; if (y > hy) hy = y; }
      }
    }
    lx = warp_min(lx); hx = warp_max(hx); ly = warp_min(ly); hy = warp_max(hy);
    if (lane == 0) {
      sm.red[wid][10] = lx; sm.red[wid][11] = hx; sm.red[wid][12] = ly; sm.red[wid][13] = hy;
    }
    __syncthreads();
    minx = sm.red[0][10]; maxx = sm.red[0][11]; miny = sm.red[0][12]; maxy = sm.red[0][13];
    for (int w = 1; w < NW; ++w) {
      if (sm.red[w][10] < minx) minx = sm.red[w][10];
      if (sm.red[w][11] > maxx) maxx = sm.red[w][11];
      if (sm.red[w][12] < miny) miny = sm.red[w][12];
      if (sm.red[w][13] > maxy) maxy = sm.red[w][13];
    }
  }
  // --- corner growth: candidates in parallel, replay in point order ---
  const double minx0 = minx, maxx0 = maxx, miny0 = miny, maxy0 = maxy;
  for (int j = tid; j < m; j += kFitBlock) {
    const double* p = W_PT(j);
    const double px = W_PX(p), py = W_PY(p);
    if ((px > maxx0 || px < minx0) && (py > maxy0 || py < miny0)) {
      const int slot = atomicAdd(&sm.ncand, 1);
      if (slot < kCornerCap) sm.cand[slot] = j;
    }
  }
  __syncthreads();
  const int nc = sm.ncand;
  if (nc > 0) {
    if (nc <= kCornerCap) {
      for (int c = tid; c < nc; c += kFitBlock) {  // rank sort (indices are distinct)
        const int mine = sm.cand[c];
        int rank = 0;
        for (int q = 0; q < nc; ++q) rank += sm.cand[q] < mine;
        sm.sorted[rank] = mine;
      }
      __syncthreads();
      if (tid == 0) {
        for (int c = 0; c < nc; ++c) {
          const double* p = W_PT(sm.sorted[c]);
          rss_corner_grow(W_PX(p), W_PY(p), W_PZ(p), cz, radsqr, minx, maxx, miny, maxy);
        }
        sm.bc[0] = minx; sm.bc[1] = maxx; sm.bc[2] = miny; sm.bc[3] = maxy;
      }
    } else if (wid == 0) {  // too many candidates for the list: warp 0 walks all points in order
      for (int base = 0; base < m; base += 32) {
        const int j = base + lane;
        double px = 0, py = 0, pz = 0;
        bool cand = false;
        if (j < m) {
          const double* p = W_PT(j);
          px = W_PX(p); py = W_PY(p); pz = W_PZ(p);
          cand = (px > maxx0 || px < minx0) && (py > maxy0 || py < miny0);
        }
        unsigned mask = __ballot_sync(0xffffffffu, cand);
        while (mask) {
          const int src = __ffs(mask) - 1;
          mask &= mask - 1;
          rss_corner_grow(__shfl_sync(0xffffffffu, px, src), __shfl_sync(0xffffffffu, py, src),
                          __shfl_sync(0xffffffffu, pz, src), cz, radsqr, minx, maxx, miny, maxy);
        }
      }
      if (lane == 0) { sm.bc[0] = minx; sm.bc[1] = maxx; sm.bc[2] = miny; sm.bc[3] = maxy; }
    }
    __syncthreads();
    minx = sm.bc[0]; maxx = sm.bc[1]; miny = sm.bc[2]; maxy = sm.bc[3];
  }
  for (int k = 0; k < 3; ++k) f.rss_To[k] = (A[3 * k] * minx + A[3 * k + 1] * miny) + A[3 * k + 2] * cz;
  f.rss_l[0] = maxx - minx;
  if (f.rss_l[0] < 0) f.rss_l[0] = 0;
  f.rss_l[1] = maxy - miny;
  if (f.rss_l[1] < 0) f.rss_l[1] = 0;
  f.rss_r = r;
  __syncthreads();  // the shared scratch may be reused by the caller
#undef W_PT
#undef W_PX
#undef W_PY
#undef W_PZ
}

__device__ __forceinline__ void store_node_records(const RefitParams& P, int node, const NodeFit& f) {
  double* o = P.obb + (size_t)node * kNodeDoubles;
  double* r = P.rss + (size_t)node * kNodeDoubles;
#pragma unroll
  for (int i = 0; i < 9; ++i) o[i] = r[i] = f.axis[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    o[9 + i] = f.obb_To[i];
    o[12 + i] = f.obb_ext[i];
    r[9 + i] = f.rss_To[i];
  }
  r[12] = f.rss_l[0];
  r[13] = f.rss_l[1];
  r[14] = f.rss_r;
  const double size = FCL_SUM3(f.obb_ext[0] * f.obb_ext[0], f.obb_ext[1] * f.obb_ext[1], f.obb_ext[2] * f.obb_ext[2]);  // extent.squaredNorm()
  o[15] = r[15] = size;
  RssRec32 r32;
  pack_rss32(f.axis, f.rss_To, f.rss_l, f.rss_r, r32);
  P.rss32[node] = r32;
  ObbRec32 o32;
  pack_obb32(f.axis, f.obb_To, f.obb_ext, o32);
  P.obb32[node] = o32;
  double2 t = P.topo[node];
  t.y = size;
  P.topo[node] = t;
}

// ---------------------------------------------------------------------------------------
// Bottom-up refit (the reference's default endReplaceModel(), BVH_model-inl.h:952-1037): one launch per tree HEIGHT.
// ids[] lists the nodes of one height (0 = leaves: closed-form triangle fit; h > 0: merge of the two children, both of
// smaller height and therefore finished by an earlier launch on the same stream), one thread per node running the very
// routines of bvh_merge.cuh the host model runs.  OBB and RSS get their own axes from here on.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void load_node_bv(const RefitParams& P, int node, NodeBV& f) {
  const double* o = P.obb + (size_t)node * kNodeDoubles;
  const double* r = P.rss + (size_t)node * kNodeDoubles;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    f.axis[i] = o[i];
    f.rss_axis[i] = r[i];
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    f.obb_To[i] = o[9 + i];
    f.obb_ext[i] = o[12 + i];
    f.rss_To[i] = r[9 + i];
  }
  f.rss_l[0] = r[12];
  f.rss_l[1] = r[13];
  f.rss_r = r[14];
}

__device__ __forceinline__ void store_node_bv(const RefitParams& P, int node, const NodeBV& f) {
  double* o = P.obb + (size_t)node * kNodeDoubles;
  double* r = P.rss + (size_t)node * kNodeDoubles;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    o[i] = f.axis[i];
    r[i] = f.rss_axis[i];
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    o[9 + i] = f.obb_To[i];
    o[12 + i] = f.obb_ext[i];
    r[9 + i] = f.rss_To[i];
  }
  r[12] = f.rss_l[0];
  r[13] = f.rss_l[1];
  r[14] = f.rss_r;
  const double size = FCL_SUM3(f.obb_ext[0] * f.obb_ext[0], f.obb_ext[1] * f.obb_ext[1], f.obb_ext[2] * f.obb_ext[2]);  // extent.squaredNorm()
  o[15] = r[15] = size;
  RssRec32 r32;
  pack_rss32(f.rss_axis, f.rss_To, f.rss_l, f.rss_r, r32);
  P.rss32[node] = r32;
  ObbRec32 o32;
  pack_obb32(f.axis, f.obb_To, f.obb_ext, o32);
  P.obb32[node] = o32;
  double2 t = P.topo[node];
  t.y = size;
  P.topo[node] = t;
}

__global__ void __launch_bounds__(64) refit_bottomup_level_kernel(RefitParams P, const int32_t* __restrict__ first_child,
                                                                  const int32_t* __restrict__ ids, int count) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const int node = ids[k];
  const int fc = first_child[node];
  NodeBV f;
  if (fc < 0) {
    fit3_obbrss(P.tri + (size_t)(-(fc + 1)) * kTriDoubles, f);
  } else {
    NodeBV a, b;
    load_node_bv(P, fc, a);
    load_node_bv(P, fc + 1, b);
    merge_obbrss(a, b, f);
  }
  store_node_bv(P, node, f);
}

// one warp per node for the next largest nodes (by_size[n_huge .. n_big))
__global__ void __launch_bounds__(128) refit_big_nodes_kernel(RefitParams P, int n_huge, int n_big) {
  __shared__ double terms[4][32][9];
  const int warp = n_huge + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (warp >= n_big) return;
  const int node = P.by_size[warp];
  NodeFit f;
  fit_obbrss_warp(P.tri, kTriDoubles, P.prim_order + P.node_first[node], P.node_count[node], terms[threadIdx.x >> 5], f);
  if ((threadIdx.x & 31) == 0) store_node_records(P, node, f);
}

// one block per node for the n_huge largest nodes (by_size[0 .. n_huge))
__global__ void __launch_bounds__(kFitBlock) refit_huge_nodes_kernel(RefitParams P) {
  __shared__ BlockFitSmem sm;
  const int node = P.by_size[blockIdx.x];
  NodeFit f;
  fit_obbrss_block(P.tri, kTriDoubles, P.prim_order + P.node_first[node], P.node_count[node], sm, f);
  if (threadIdx.x == 0) store_node_records(P, node, f);
}

// one thread per node for the rest (by_size[n_big .. n_nodes))
__global__ void __launch_bounds__(64) refit_small_nodes_kernel(RefitParams P, int n_big) {
  const int k = n_big + blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= P.n_nodes) return;
  const int node = P.by_size[k];
  NodeFit f;
  fit_obbrss(P.tri, kTriDoubles, P.prim_order + P.node_first[node], P.node_count[node], f);
  store_node_records(P, node, f);
}

}  // namespace fclgpu

namespace fclgpu {

// ---------------------------------------------------------------------------------------
// On-device top-down BUILD (BVHModel::buildTree / recursiveBuildTree, BVH_model-inl.h:833-938)
// for the mean and BV-centre split rules, level by level.  Per node of the current level:
//   1. fit (same routines as the refit);
//   2. split threshold: mean rule = (vsum . axis0) / (3 n) with vsum from the fit (the rule sums
//      the same terms in the same order, BV_splitter-inl.h:578-589); BV-centre rule = obb.To[0];
//   3. side flag of every primitive (parallel): !(axis0 . centroid > threshold);
//   4. the reference's in-place swap partition (:888-926), reproduced exactly: elements going left
//      keep their order; the elements staying right behave as a queue that is rotated by one each
//      time a left element arrives.  One lane replays that queue on a scratch array;
//   5. children: ids follow the reference's pre-order pair allocation -- a node whose pair is
//      (c, c+1) gives its left child the pair c+2 and its right child the pair c + 2*n_left.
// ---------------------------------------------------------------------------------------
struct BuildNode {
  int32_t id, first, count, pair;  // pair = first_child id to assign if the node is internal
};

struct BuildParams {
  RefitParams R;          // record arrays to fill (obb, rss, tri, rss32, obb32, topo, prim_order ...)
  int32_t* first_child;   // per node
  int32_t* node_first;    // per node (writable aliases of R.node_first / R.node_count)
  int32_t* node_count;
  uint32_t* prim_order;   // writable alias of R.prim_order
  uint8_t* flag;          // per primitive slot: 1 = goes left
  uint32_t* queue;        // 2 slots per primitive slot
  const BuildNode* level; // nodes of this level
  BuildNode* next;        // nodes of the next level
  int32_t* next_count;
  int32_t n_level;
  int32_t split;          // FCLGPU_SPLIT_METHOD_MEAN or _BV_CENTER
};

// ---------------------------------------------------------------------------------------
// Median split rule (computeSplitValue_median, BV_splitter-inl.h:603-657): the reference sorts the projections of the
// node's triangle centroids on the split axis and takes the middle one (odd count) or the mean of the two middle ones
// (even).  Only those one or two order statistics matter, so nothing is sorted here: the projections go to the node's
// slice of the queue scratch as order-preserving 64-bit keys and the k-th smallest is found bit by bit -- the answer is
// the largest t with |{key < t}| <= k, built from the most significant bit down, one cooperative counting pass per bit.
// kMode 0: one thread (n <= 24: insertion sort of a local array instead), 1: one warp, 2: one block of kFitBlock threads.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ordered_key_of(double x) {
  const long long b = __double_as_longlong(x);
  return b < 0 ? ~(unsigned long long)b : ((unsigned long long)b | 0x8000000000000000ull);
}
__device__ __forceinline__ double double_of_ordered_key(unsigned long long k) {
  return __longlong_as_double((long long)((k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k));
}
__device__ __forceinline__ double centroid_projection(const double* __restrict__ tri, uint32_t prim, double sv0, double sv1, double sv2) {
  const double* p1 = tri + (size_t)prim * kTriDoubles;
  const double* p2 = p1 + 3;
  const double* p3 = p1 + 6;
  const double c0 = (p1[0] + p2[0]) + p3[0], c1 = (p1[1] + p2[1]) + p3[1], c2 = (p1[2] + p2[2]) + p3[2];
  return FCL_SUM3(c0 * sv0, c1 * sv1, c2 * sv2) / 3;
}

template <int kMode>
__device__ inline long long group_sum(long long v, int* wcnt) {
  if constexpr (kMode == 0) {
    return v;
  } else {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if constexpr (kMode == 2) {
      __syncthreads();  // the previous round's readers are done with wcnt
      if ((threadIdx.x & 31) == 0) wcnt[threadIdx.x >> 5] = (int)v;
      __syncthreads();
      v = 0;
      for (int w = 0; w < kFitBlock / 32; ++w) v += wcnt[w];
    }
    return v;
  }
}

template <int kMode>
__device__ inline double median_split_value(const BuildParams& B, const BuildNode nd, double sv0, double sv1, double sv2, int* wcnt) {
  const RefitParams& P = B.R;
  const uint32_t* idx = B.prim_order + nd.first;
  const int n = nd.count;
  if constexpr (kMode == 0) {  // n <= 24
    double proj[24];
    for (int i = 0; i < n; ++i) {
      const double v = centroid_projection(P.tri, idx[i], sv0, sv1, sv2);
      int j = i;
      while (j > 0 && v < proj[j - 1]) {
        proj[j] = proj[j - 1];
        --j;
      }
      proj[j] = v;
    }
    return (n % 2 == 1) ? proj[(n - 1) / 2] : (proj[n / 2] + proj[n / 2 - 1]) / 2;
  } else {
  const int tid = (kMode == 1) ? (threadIdx.x & 31) : threadIdx.x;
  const int stride = (kMode == 1) ? 32 : kFitBlock;
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(B.queue + 2 * (size_t)nd.first);
  for (int i = tid; i < n; i += stride) keys[i] = ordered_key_of(centroid_projection(P.tri, idx[i], sv0, sv1, sv2));
  if (kMode == 1) __syncwarp();
  else __syncthreads();
  const long long k = (n % 2 == 1) ? (n - 1) / 2 : n / 2 - 1;  // lower middle (0-based rank)
  unsigned long long ans = 0;
  for (int bit = 63; bit >= 0; --bit) {
    const unsigned long long t = ans | (1ull << bit);
    long long c = 0;
    for (int i = tid; i < n; i += stride) c += keys[i] < t ? 1 : 0;
    c = group_sum<kMode>(c, wcnt);
    if (c <= k) ans = t;
  }
  double lo = double_of_ordered_key(ans);
  if (n % 2 == 1) return lo;
  // upper middle: the same value when it occurs more than once past rank k, else the smallest key above it
  long long le = 0;
  unsigned long long above = ~0ull;
  for (int i = tid; i < n; i += stride) {
    const unsigned long long x = keys[i];
    le += x <= ans ? 1 : 0;
    if (x > ans && x < above) above = x;
  }
  le = group_sum<kMode>(le, wcnt);
  // min over the group, as two 32-bit halves through the same sum-free path: shuffles inside a warp, wcnt across warps
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = shfl_u64(above, (threadIdx.x & 31) ^ o);
    above = other < above ? other : above;
  }
  if (kMode == 2) {
    __shared__ unsigned long long s_above[kFitBlock / 32];
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_above[threadIdx.x >> 5] = above;
    __syncthreads();
    for (int w = 0; w < kFitBlock / 32; ++w) above = s_above[w] < above ? s_above[w] : above;
  }
  const double hi = (le > k + 1) ? lo : double_of_ordered_key(above);
  return (hi + lo) / 2;
  }
}

__device__ inline void build_finish_node(const BuildParams& B, const BuildNode nd, const NodeFit& f, int lane, bool warp_mode) {
  const RefitParams& P = B.R;
  uint32_t* idx = B.prim_order + nd.first;
  const int n = nd.count;
  if (!warp_mode || lane == 0) {
    store_node_records(P, nd.id, f);
    B.node_first[nd.id] = nd.first;
    B.node_count[nd.id] = n;
  }
  if (n == 1) {
    if (!warp_mode || lane == 0) {
      const int fc = -((int)idx[0] + 1);
      B.first_child[nd.id] = fc;
      double2 t = P.topo[nd.id];
      long long bits = (long long)(unsigned)fc;
      t.x = __longlong_as_double(bits);
      P.topo[nd.id] = t;
    }
    return;
  }
  const double sv0 = f.axis[0], sv1 = f.axis[3], sv2 = f.axis[6];
  double thr;
  if (B.split == FCLGPU_SPLIT_METHOD_BV_CENTER) thr = f.obb_To[0];
  else if (B.split == FCLGPU_SPLIT_METHOD_MEDIAN)
    thr = warp_mode ? median_split_value<1>(B, nd, sv0, sv1, sv2, nullptr) : median_split_value<0>(B, nd, sv0, sv1, sv2, nullptr);
  else thr = (f.vsum[0] * sv0 + f.vsum[1] * sv1 + f.vsum[2] * sv2) / (3 * n);
  if (warp_mode) __syncwarp();  // (median) every lane is done with the key scratch before it is reused below
  // side flags
  for (int i = warp_mode ? lane : 0; i < n; i += warp_mode ? 32 : 1) {
    const double* p1 = P.tri + (size_t)idx[i] * kTriDoubles;
    const double* p2 = p1 + 3;
    const double* p3 = p1 + 6;
    const double cx = ((p1[0] + p2[0]) + p3[0]) / 3.0, cy = ((p1[1] + p2[1]) + p3[1]) / 3.0, cz = ((p1[2] + p2[2]) + p3[2]) / 3.0;
    B.flag[nd.first + i] = (FCL_SUM3(sv0 * cx, sv1 * cy, sv2 * cz) > thr) ? 0 : 1;
  }
  int c1 = 0;
  if (warp_mode) {
    // Parallel form of the replay.  Let i0 be the first right element; number the operations from there
    // (time 1 = i0).  Every operation enqueues exactly one item (a push, or the re-queue of the served head)
    // and, the queue being FIFO, the j-th rotation serves the item enqueued at time j.  With J rotations in
    // total the final queue holds the items enqueued at times J+1..T, in that order; an item enqueued by a
    // rotation is the item of an earlier time, so each output follows a chain of strictly decreasing times
    // until it reaches a push.  Chains of different outputs are disjoint (total work <= T).
    __syncwarp();
    uint32_t* old = B.queue + 2 * (size_t)nd.first;  // copy of the incoming order
    uint32_t* lrank = old + n;                        // inclusive count of left flags
    const uint8_t* fl = B.flag + nd.first;
    int run = 0, i0 = n;
    for (int base = 0; base < n; base += 32) {
      const int i = base + lane;
      const bool in = i < n;
      const bool left = in && fl[i];
      if (in) old[i] = idx[i];
      const unsigned m = __ballot_sync(0xffffffffu, left);
      if (in) lrank[i] = run + __popc(m & (0xffffffffu >> (31 - lane)));
      if (i0 == n) {
        const unsigned rm = __ballot_sync(0xffffffffu, in && !left);
        if (rm) i0 = base + __ffs(rm) - 1;
      }
      run += __popc(m);
    }
    c1 = run;
    __syncwarp();
    if (c1 != 0 && c1 != n) {
      const int J = c1 - i0, T = n - i0;
      for (int i = lane; i < n; i += 32)
        if (fl[i]) idx[lrank[i] - 1] = old[i];
      for (int k = J + 1 + lane; k <= T; k += 32) {
        int i = i0 + k - 1;
        while (fl[i]) i = i0 + ((int)lrank[i] - i0) - 1;  // rotation number j -> item of time j
        idx[c1 + (k - J - 1)] = old[i];
      }
    } else {
      c1 = n / 2;  // degenerate split: the reference's loop leaves the order untouched
    }
    __syncwarp();
  } else {
    // replay the swap partition: left elements in order, right elements through the rotating queue
    uint32_t* Q = B.queue + 2 * (size_t)nd.first;
    int head = 0, tail = 0;
    for (int i = 0; i < n; ++i) {
      const uint32_t x = idx[i];
      if (B.flag[nd.first + i]) {
        if (tail > head) Q[tail++] = Q[head++];  // the oldest right element jumps behind the block
        idx[c1++] = x;                           // c1 <= i: slot already consumed
      } else {
        Q[tail++] = x;
      }
    }
    if (c1 != 0 && c1 != n)
      for (int k = 0; k < n - c1; ++k) idx[c1 + k] = Q[head + k];
    else {
      // degenerate split: the reference's loop leaves the order untouched (self swaps / no swaps)
      if (c1 == 0)
        for (int k = 0; k < n; ++k) idx[k] = Q[head + k];
      c1 = n / 2;
    }
  }
  if (!warp_mode || lane == 0) {
    B.first_child[nd.id] = nd.pair;
    double2 t = P.topo[nd.id];
    long long bits = (long long)(unsigned)nd.pair;
    t.x = __longlong_as_double(bits);
    P.topo[nd.id] = t;
    const int slot = atomicAdd(B.next_count, 2);
    B.next[slot] = BuildNode{nd.pair, nd.first, c1, nd.pair + 2};
    B.next[slot + 1] = BuildNode{nd.pair + 1, nd.first + c1, n - c1, nd.pair + 2 * c1};
  }
}

// Block-cooperative version of build_finish_node for huge nodes (same results; see the warp branch above
// for the closed form of the swap-partition replay).
__device__ inline void build_finish_node_block(const BuildParams& B, const BuildNode nd, const NodeFit& f, BlockFitSmem& sm) {
  constexpr int NW = kFitBlock / 32;
  const RefitParams& P = B.R;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  uint32_t* idx = B.prim_order + nd.first;
  const int n = nd.count;  // > 24
  if (tid == 0) {
    store_node_records(P, nd.id, f);
    B.node_first[nd.id] = nd.first;
    B.node_count[nd.id] = n;
  }
  const double sv0 = f.axis[0], sv1 = f.axis[3], sv2 = f.axis[6];
  double thr;
  if (B.split == FCLGPU_SPLIT_METHOD_BV_CENTER) thr = f.obb_To[0];
  else if (B.split == FCLGPU_SPLIT_METHOD_MEDIAN) {
    thr = median_split_value<2>(B, nd, sv0, sv1, sv2, sm.wcnt);
    __syncthreads();  // every thread is done with the key scratch before it is reused below
  } else thr = (f.vsum[0] * sv0 + f.vsum[1] * sv1 + f.vsum[2] * sv2) / (3 * n);
  uint32_t* old = B.queue + 2 * (size_t)nd.first;
  uint32_t* lrank = old + n;
  uint8_t* fl = B.flag + nd.first;
  int run = 0, my_i0 = n;
  for (int base = 0; base < n; base += kFitBlock) {
    const int i = base + tid;
    const bool in = i < n;
    bool left = false;
    if (in) {
      const uint32_t pi = idx[i];
      const double* p1 = P.tri + (size_t)pi * kTriDoubles;
      const double* p2 = p1 + 3;
      const double* p3 = p1 + 6;
      const double cx = ((p1[0] + p2[0]) + p3[0]) / 3.0, cy = ((p1[1] + p2[1]) + p3[1]) / 3.0, cz = ((p1[2] + p2[2]) + p3[2]) / 3.0;
      left = !(FCL_SUM3(sv0 * cx, sv1 * cy, sv2 * cz) > thr);
      fl[i] = left ? 1 : 0;
      old[i] = pi;
      if (!left && my_i0 == n) my_i0 = i;
    }
    const unsigned m = __ballot_sync(0xffffffffu, left);
    if (lane == 0) sm.wcnt[wid] = __popc(m);
    __syncthreads();
    int pre = 0, tot = 0;
    for (int w = 0; w < NW; ++w) {
      const int c = sm.wcnt[w];
      pre += (w < wid) ? c : 0;
      tot += c;
    }
    if (in) lrank[i] = run + pre + __popc(m & (0xffffffffu >> (31 - lane)));
    run += tot;
    __syncthreads();
  }
  // first right element
  for (int o = 16; o > 0; o >>= 1) {
    const int w = __shfl_xor_sync(0xffffffffu, my_i0, o);
    my_i0 = w < my_i0 ? w : my_i0;
  }
  if (lane == 0) sm.wcnt[wid] = my_i0;
  __syncthreads();  // also makes old / lrank / fl visible to the whole block
  int i0 = n;
  for (int w = 0; w < NW; ++w) i0 = sm.wcnt[w] < i0 ? sm.wcnt[w] : i0;
  int c1 = run;
  if (c1 != 0 && c1 != n) {
    const int J = c1 - i0, T = n - i0;
    for (int i = tid; i < n; i += kFitBlock)
      if (fl[i]) idx[lrank[i] - 1] = old[i];
    for (int k = J + 1 + tid; k <= T; k += kFitBlock) {
      int i = i0 + k - 1;
      while (fl[i]) i = (int)lrank[i] - 1;  // = i0 + (lrank - i0) - 1: rotation number j -> item of time j
      idx[c1 + (k - J - 1)] = old[i];
    }
  } else {
    c1 = n / 2;
  }
  if (tid == 0) {
    B.first_child[nd.id] = nd.pair;
    double2 t = P.topo[nd.id];
    long long bits = (long long)(unsigned)nd.pair;
    t.x = __longlong_as_double(bits);
    P.topo[nd.id] = t;
    const int slot = atomicAdd(B.next_count, 2);
    B.next[slot] = BuildNode{nd.pair, nd.first, c1, nd.pair + 2};
    B.next[slot + 1] = BuildNode{nd.pair + 1, nd.first + c1, n - c1, nd.pair + 2 * c1};
  }
}

// one block per node of the level (top levels: a handful of huge nodes)
__global__ void __launch_bounds__(kFitBlock) build_level_block_kernel(BuildParams B) {
  __shared__ BlockFitSmem sm;
  const BuildNode nd = B.level[blockIdx.x];
  NodeFit f;
  if (nd.count > 24) {
    fit_obbrss_block(B.R.tri, kTriDoubles, B.prim_order + nd.first, nd.count, sm, f);
    build_finish_node_block(B, nd, f, sm);
  } else if (threadIdx.x == 0) {
    fit_obbrss(B.R.tri, kTriDoubles, B.prim_order + nd.first, nd.count, f);
    build_finish_node(B, nd, f, 0, false);
  }
}

// kGrouped = false: one warp per node of the level (upper levels: few, large nodes).
// kGrouped = true: a warp takes 32 consecutive nodes; nodes of <= 24 triangles are done one per lane,
// the larger ones among the 32 afterwards by the whole warp, one at a time (lower levels: many small nodes).
template <bool kGrouped>
__global__ void __launch_bounds__(128) build_level_kernel(BuildParams B) {
  __shared__ double terms[4][32][9];
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  NodeFit f;
  if (!kGrouped) {
    if (warp >= B.n_level) return;
    const BuildNode nd = B.level[warp];
    if (nd.count > 24) {
      fit_obbrss_warp(B.R.tri, kTriDoubles, B.prim_order + nd.first, nd.count, terms[threadIdx.x >> 5], f);
      build_finish_node(B, nd, f, lane, true);
    } else if (lane == 0) {
      fit_obbrss(B.R.tri, kTriDoubles, B.prim_order + nd.first, nd.count, f);
      build_finish_node(B, nd, f, 0, false);
    }
  } else {
    if (warp * 32 >= B.n_level) return;
    const int k = warp * 32 + lane;
    const bool valid = k < B.n_level;
    BuildNode nd{0, 0, 0, 0};
    if (valid) nd = B.level[k];
    const bool small = valid && nd.count <= 24;
    if (small) {
      fit_obbrss(B.R.tri, kTriDoubles, B.prim_order + nd.first, nd.count, f);
      build_finish_node(B, nd, f, 0, false);
    }
    __syncwarp();
    unsigned big = __ballot_sync(0xffffffffu, valid && !small);
    while (big) {
      const int src = __ffs(big) - 1;
      big &= big - 1;
      BuildNode b;
      b.id = __shfl_sync(0xffffffffu, nd.id, src);
      b.first = __shfl_sync(0xffffffffu, nd.first, src);
      b.count = __shfl_sync(0xffffffffu, nd.count, src);
      b.pair = __shfl_sync(0xffffffffu, nd.pair, src);
      fit_obbrss_warp(B.R.tri, kTriDoubles, B.prim_order + b.first, b.count, terms[threadIdx.x >> 5], f);
      build_finish_node(B, b, f, lane, true);
      __syncwarp();
    }
  }
}

__global__ void iota_kernel(uint32_t* p, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (uint32_t)i;
}

}  // namespace fclgpu
