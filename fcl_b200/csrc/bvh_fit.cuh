// bvh_fit.cuh — OBBRSS fit of one BVH node over a set of triangles, shared by the host builder
// (bvh_build.cpp) and the on-device top-down refit kernel (refit.cuh).
//
// What it computes (reference, /root/reference/include/fcl/...):
//   FitImpl<OBBRSS>::run                 geometry/bvh/detail/BV_fitter-inl.h:449-477
//   getCovariance (triangles)            math/geometry-inl.h:1335-1425
//   eigen_old (cyclic Jacobi), axisFromEigen   math/geometry-inl.h:477-558, 563-597
//   getExtentAndCenter_mesh              math/geometry-inl.h:294-362
//   getRadiusAndOriginAndRectangleSize   math/geometry-inl.h:709-988
// The reference stores the projected points in a temporary vector; here every pass recomputes
// the three projections of a vertex (same expression, hence the same bits), so the routine
// needs no scratch memory and can run inside a kernel.  Sums run sequentially in primitive
// order, like the reference, so the result is bit-identical to a CPU fit.
#pragma once
#include <cuda_runtime.h>

#include <cfloat>
#include <cmath>
#include <cstdint>

#include "sum_order.h"

namespace fclgpu {

#ifndef FD
#define FD __host__ __device__ __forceinline__
#endif

struct NodeFit {
  double axis[9];  // row-major, columns = box axes
  double obb_To[3], obb_ext[3], rss_To[3], rss_l[2], rss_r;
  double vsum[3];  // sum over the node's triangles of ((p1 + p2) + p3), in primitive order: the covariance's S1
                   // and, term for term, the centroid sum of the mean split rule (BV_splitter-inl.h:578-589)
};

// symmetric 3x3 eigen-decomposition, cyclic Jacobi; v[r][k]: k-th eigenvector in column k
__host__ __device__ inline void jacobi3(const double M[3][3], double d[3], double v[3][3]) {
  double R[3][3], b[3], z[3];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) {
      R[i][j] = M[i][j];
      v[i][j] = (i == j) ? 1.0 : 0.0;
    }
    b[i] = d[i] = R[i][i];
    z[i] = 0;
  }
  for (int sweep = 0; sweep < 50; ++sweep) {
    double sm = 0;
    sm += fabs(R[0][1]);
    sm += fabs(R[0][2]);
    sm += fabs(R[1][2]);
    if (sm == 0.0) return;
    const double tresh = (sweep < 3) ? 0.2 * sm / 9 : 0.0;
    for (int ip = 0; ip < 3; ++ip)
      for (int iq = ip + 1; iq < 3; ++iq) {
        double g = 100.0 * fabs(R[ip][iq]);
        if (sweep > 3 && fabs(d[ip]) + g == fabs(d[ip]) && fabs(d[iq]) + g == fabs(d[iq])) {
          R[ip][iq] = 0.0;
          continue;
        }
        if (!(fabs(R[ip][iq]) > tresh)) continue;
        double h = d[iq] - d[ip], t;
        if (fabs(h) + g == fabs(h)) {
          t = R[ip][iq] / h;
        } else {
          const double theta = 0.5 * h / R[ip][iq];
          t = 1.0 / (fabs(theta) + sqrt(1.0 + theta * theta));
          if (theta < 0.0) t = -t;
        }
        const double c = 1.0 / sqrt(1 + t * t), s = t * c, tau = s / (1.0 + c);
        h = t * R[ip][iq];
        z[ip] -= h;
        z[iq] += h;
        d[ip] -= h;
        d[iq] += h;
        R[ip][iq] = 0.0;
#define FCLGPU_GIVENS(x, y)                \
  {                                        \
    const double gx = (x), hy = (y);       \
    (x) = gx - s * (hy + gx * tau);        \
    (y) = hy + s * (gx - hy * tau);        \
  }
        for (int j = 0; j < ip; ++j) FCLGPU_GIVENS(R[j][ip], R[j][iq]);
        for (int j = ip + 1; j < iq; ++j) FCLGPU_GIVENS(R[ip][j], R[j][iq]);
        for (int j = iq + 1; j < 3; ++j) FCLGPU_GIVENS(R[ip][j], R[iq][j]);
        for (int j = 0; j < 3; ++j) FCLGPU_GIVENS(v[j][ip], v[j][iq]);
#undef FCLGPU_GIVENS
      }
    for (int i = 0; i < 3; ++i) {
      b[i] += z[i];
      d[i] = b[i];
      z[i] = 0.0;
    }
  }
}

// Centre and half extents of the box with the given axes around m points (getExtentAndCenter_mesh /
// _pointcloud: identical arithmetic once the points are enumerated in the same order).  pt(j) -> pointer to 3 doubles.
template <class PointOf>
__host__ __device__ inline void extent_center_from_points(PointOf pt, int m, const double* A, double To[3], double ext[3]) {
  const double a00 = A[0], a10 = A[3], a20 = A[6], a01 = A[1], a11 = A[4], a21 = A[7], a02 = A[2], a12 = A[5], a22 = A[8];
  double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
  for (int j = 0; j < m; ++j) {
    const double* p = pt(j);
    const double c0 = FCL_SUM3(a00 * p[0], a10 * p[1], a20 * p[2]);  // axis.col(0).dot(p)
    const double c1 = FCL_SUM3(a01 * p[0], a11 * p[1], a21 * p[2]);
    const double c2 = FCL_SUM3(a02 * p[0], a12 * p[1], a22 * p[2]);
    if (c0 > mx[0]) mx[0] = c0;
    if (c0 < mn[0]) mn[0] = c0;
    if (c1 > mx[1]) mx[1] = c1;
    if (c1 < mn[1]) mn[1] = c1;
    if (c2 > mx[2]) mx[2] = c2;
    if (c2 < mn[2]) mn[2] = c2;
  }
  const double o[3] = {(mx[0] + mn[0]) / 2, (mx[1] + mn[1]) / 2, (mx[2] + mn[2]) / 2};
  for (int r = 0; r < 3; ++r) {
    To[r] = FCL_SUM3(A[3 * r] * o[0], A[3 * r + 1] * o[1], A[3 * r + 2] * o[2]);  // axis * o
    ext[r] = (mx[r] - mn[r]) / 2;
  }
}

// Rectangle-swept-sphere with the given axes around m points (getRadiusAndOriginAndRectangleSize after the projections
// have been gathered, math/geometry-inl.h:782-988): radius from the thin direction, rectangle grown to cover.  The
// reference keeps the projections in a temporary vector; every pass here recomputes them (same expression, same bits).
template <class PointOf>
__host__ __device__ inline void rss_from_points(PointOf pt, int m, const double* A, double To[3], double l[2], double& rad) {
  const double a00 = A[0], a10 = A[3], a20 = A[6], a01 = A[1], a11 = A[4], a21 = A[7], a02 = A[2], a12 = A[5], a22 = A[8];
#define FCLGPU_PX(p) FCL_SUM3(a00 * (p)[0], a10 * (p)[1], a20 * (p)[2])
#define FCLGPU_PY(p) FCL_SUM3(a01 * (p)[0], a11 * (p)[1], a21 * (p)[2])
#define FCLGPU_PZ(p) FCL_SUM3(a02 * (p)[0], a12 * (p)[1], a22 * (p)[2])
  double minz = 0, maxz = 0;
  for (int j = 0; j < m; ++j) {
    const double* p = pt(j);
    const double c2 = FCLGPU_PZ(p);
    if (j == 0) {
      minz = maxz = c2;
    } else {
      if (c2 < minz) minz = c2;
      else if (c2 > maxz) maxz = c2;
    }
  }
  const double r = 0.5 * (maxz - minz), radsqr = r * r, cz = 0.5 * (maxz + minz);
  double lo2[2], hi2[2];
  for (int c = 0; c < 2; ++c) {
    int imin = 0, imax = 0;
    double vmin, vmax;
    {
      const double* p = pt(0);
      vmin = vmax = (c == 0) ? FCLGPU_PX(p) : FCLGPU_PY(p);
    }
    for (int j = 1; j < m; ++j) {
      const double* p = pt(j);
      const double val = (c == 0) ? FCLGPU_PX(p) : FCLGPU_PY(p);
      if (val < vmin) { imin = j; vmin = val; }
      else if (val > vmax) { imax = j; vmax = val; }
    }
    double lo_c, hi_c;
    {
      const double* p = pt(imin);
      const double dz = FCLGPU_PZ(p) - cz;
      lo_c = ((c == 0) ? FCLGPU_PX(p) : FCLGPU_PY(p)) + sqrt(fmax(radsqr - dz * dz, 0.0));
      const double* q = pt(imax);
      const double dz2 = FCLGPU_PZ(q) - cz;
      hi_c = ((c == 0) ? FCLGPU_PX(q) : FCLGPU_PY(q)) - sqrt(fmax(radsqr - dz2 * dz2, 0.0));
    }
    for (int j = 0; j < m; ++j) {
      const double* p = pt(j);
      const double val = (c == 0) ? FCLGPU_PX(p) : FCLGPU_PY(p);
      if (val < lo_c) {
        const double dz = FCLGPU_PZ(p) - cz;
        const double x = val + sqrt(fmax(radsqr - dz * dz, 0.0));
        if (x < lo_c) lo_c = x;
      }
    }
    for (int j = 0; j < m; ++j) {
      const double* p = pt(j);
      const double val = (c == 0) ? FCLGPU_PX(p) : FCLGPU_PY(p);
      if (val > hi_c) {
        const double dz = FCLGPU_PZ(p) - cz;
        const double x = val - sqrt(fmax(radsqr - dz * dz, 0.0));
        if (x > hi_c) hi_c = x;
      }
    }
    lo2[c] = lo_c;
    hi2[c] = hi_c;
  }
  double minx = lo2[0], maxx = hi2[0], miny = lo2[1], maxy = hi2[1];
  const double a = sqrt(0.5);
  for (int j = 0; j < m; ++j) {
    const double* p = pt(j);
    const double px = FCLGPU_PX(p), py = FCLGPU_PY(p), pz = FCLGPU_PZ(p);
    double dx, dy, u, t;
    if (px > maxx) {
      if (py > maxy) {
        dx = px - maxx; dy = py - maxy;
        u = dx * a + dy * a;
        t = (a * u - dx) * (a * u - dx) + (a * u - dy) * (a * u - dy) + (cz - pz) * (cz - pz);
        u = u - sqrt(fmax(radsqr - t, 0.0));
        if (u > 0) { maxx += u * a; maxy += u * a; }
      } else if (py < miny) {
        dx = px - maxx; dy = py - miny;
        u = dx * a - dy * a;
        t = (a * u - dx) * (a * u - dx) + (-a * u - dy) * (-a * u - dy) + (cz - pz) * (cz - pz);
        u = u - sqrt(fmax(radsqr - t, 0.0));
        if (u > 0) { maxx += u * a; miny -= u * a; }
      }
    } else if (px < minx) {
      if (py > maxy) {
        dx = px - minx; dy = py - maxy;
        u = dy * a - dx * a;
        t = (-a * u - dx) * (-a * u - dx) + (a * u - dy) * (a * u - dy) + (cz - pz) * (cz - pz);
        u = u - sqrt(fmax(radsqr - t, 0.0));
        if (u > 0) { minx -= u * a; maxy += u * a; }
      } else if (py < miny) {
        dx = px - minx; dy = py - miny;
        u = -dx * a - dy * a;
        t = (-a * u - dx) * (-a * u - dx) + (-a * u - dy) * (-a * u - dy) + (cz - pz) * (cz - pz);
        u = u - sqrt(fmax(radsqr - t, 0.0));
        if (u > 0) { minx -= u * a; miny -= u * a; }
      }
    }
  }
  for (int k = 0; k < 3; ++k) To[k] = (A[3 * k] * minx + A[3 * k + 1] * miny) + A[3 * k + 2] * cz;
  l[0] = maxx - minx;
  if (l[0] < 0) l[0] = 0;
  l[1] = maxy - miny;
  if (l[1] < 0) l[1] = 0;
  rad = r;
#undef FCLGPU_PX
#undef FCLGPU_PY
#undef FCLGPU_PZ
}

// principal axes from a covariance matrix: largest, middle eigenvector (eigen_old + axisFromEigen), then their cross
// product.  A row-major, column c = c-th axis.
__host__ __device__ inline void axes_from_covariance(const double M[3][3], double* A) {
  double ev[3], V[3][3];
  jacobi3(M, ev, V);
  int lo, mid, hi;
  if (ev[0] > ev[1]) { hi = 0; lo = 1; } else { lo = 0; hi = 1; }
  if (ev[2] < ev[lo]) { mid = lo; lo = 2; }
  else if (ev[2] > ev[hi]) { mid = hi; hi = 2; }
  else mid = 2;
  for (int r = 0; r < 3; ++r) {
    A[3 * r + 0] = V[r][hi];
    A[3 * r + 1] = V[r][mid];
  }
  A[2] = A[3] * A[7] - A[6] * A[4];
  A[5] = A[6] * A[1] - A[0] * A[7];
  A[8] = A[0] * A[4] - A[3] * A[1];
}

// tv: de-indexed triangle vertices, 9 doubles per triangle with stride `tri_stride` doubles.
// idx[0..n): the node's primitives in the order the sums must run.
__host__ __device__ inline void fit_obbrss(const double* tv, int tri_stride, const uint32_t* idx, int n, NodeFit& f) {
  // --- covariance of the 3n vertices ---
  double S1[3] = {0, 0, 0}, S2[6] = {0, 0, 0, 0, 0, 0};  // xx yy zz xy xz yz
  for (int i = 0; i < n; ++i) {
    const double* p1 = tv + (size_t)idx[i] * tri_stride;
    const double* p2 = p1 + 3;
    const double* p3 = p1 + 6;
    for (int k = 0; k < 3; ++k) S1[k] += ((p1[k] + p2[k]) + p3[k]);
    S2[0] += (p1[0] * p1[0] + p2[0] * p2[0] + p3[0] * p3[0]);
    S2[1] += (p1[1] * p1[1] + p2[1] * p2[1] + p3[1] * p3[1]);
    S2[2] += (p1[2] * p1[2] + p2[2] * p2[2] + p3[2] * p3[2]);
    S2[3] += (p1[0] * p1[1] + p2[0] * p2[1] + p3[0] * p3[1]);
    S2[4] += (p1[0] * p1[2] + p2[0] * p2[2] + p3[0] * p3[2]);
    S2[5] += (p1[1] * p1[2] + p2[1] * p2[2] + p3[1] * p3[2]);
  }
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
  axes_from_covariance(M, f.axis);
  // point j = vertex j % 3 of the node's (j / 3)-th primitive
  auto pt = [&](int j) { return tv + (size_t)idx[j / 3] * tri_stride + 3 * (j % 3); };
  extent_center_from_points(pt, 3 * n, f.axis, f.obb_To, f.obb_ext);
  rss_from_points(pt, 3 * n, f.axis, f.rss_To, f.rss_l, f.rss_r);
}

}  // namespace fclgpu
