// bvh_merge.cuh — the two operations of the reference's BOTTOM-UP refit, shared by the host model
// (bvh_build.cpp) and the level-by-level refit kernel (refit.cuh): same source, same bits.
//
// Reference (file:line under /root/reference/include/fcl):
//   BVHModel::recursiveRefitTree_bottomup     geometry/bvh/BVH_model-inl.h:961-1037  (leaf: fit(v, 3, bv); inner: left + right)
//   OBBRSS_fit_functions::fit3                math/bv/utility-inl.h:92-117 (OBB), 208-230 (RSS), 507-511
//   OBBRSS::operator+                         math/bv/OBBRSS-inl.h:95-101
//   OBB::operator+, merge_largedist / _smalldist, computeVertices   math/bv/OBB-inl.h:161-174, 233-369
//   RSS::operator+                            math/bv/RSS-inl.h:313-371
// The reference's quirks are part of the contract and are kept (each one is marked where it happens): after this
// refit the volumes are NOT guaranteed to contain their subtrees, exactly as in the reference.
// Eigen's quaternion <-> matrix conversions and normalize() are restated from Eigen 3.3's published algorithms.
#pragma once
#include "bvh_fit.cuh"

namespace fclgpu {

struct NodeBV {
  double axis[9];  // obb.axis, row-major (column c = c-th box axis)
  double obb_To[3], obb_ext[3];
  double rss_axis[9];  // rss.axis: no longer the OBB's after a merge
  double rss_To[3], rss_l[2], rss_r;
};

namespace merge_detail {

__host__ __device__ inline double dot3p(const double* a, const double* b) { return FCL_SUM3(a[0] * b[0], a[1] * b[1], a[2] * b[2]); }
__host__ __device__ inline void normalize3(double* v) {  // Eigen: z = squaredNorm(); if (z > 0) v /= sqrt(z)
  const double z = dot3p(v, v);
  if (z > 0) {
    const double n = sqrt(z);
    v[0] /= n;
    v[1] /= n;
    v[2] /= n;
  }
}
__host__ __device__ inline void set_col(double* A, int c, const double* v) {
  A[c] = v[0];
  A[3 + c] = v[1];
  A[6 + c] = v[2];
}
__host__ __device__ inline void get_col(const double* A, int c, double* v) {
  v[0] = A[c];
  v[1] = A[3 + c];
  v[2] = A[6 + c];
}

// covariance of n plain points (getCovariance, point branch, math/geometry-inl.h:1383-1425)
__host__ __device__ inline void covariance_of_points(const double (*ps)[3], int n, double M[3][3]) {
  double S1[3] = {0, 0, 0}, S2[6] = {0, 0, 0, 0, 0, 0};  // xx yy zz xy xz yz
  for (int i = 0; i < n; ++i) {
    const double* p = ps[i];
    for (int k = 0; k < 3; ++k) S1[k] += p[k];
    S2[0] += (p[0] * p[0]);
    S2[1] += (p[1] * p[1]);
    S2[2] += (p[2] * p[2]);
    S2[3] += (p[0] * p[1]);
    S2[4] += (p[0] * p[2]);
    S2[5] += (p[1] * p[2]);
  }
  M[0][0] = S2[0] - S1[0] * S1[0] / n;
  M[1][1] = S2[1] - S1[1] * S1[1] / n;
  M[2][2] = S2[2] - S1[2] * S1[2] / n;
  M[0][1] = M[1][0] = S2[3] - S1[0] * S1[1] / n;
  M[1][2] = M[2][1] = S2[5] - S1[1] * S1[2] / n;
  M[0][2] = M[2][0] = S2[4] - S1[0] * S1[2] / n;
}

__host__ __device__ inline void order_eigenvalues(const double s[3], int& mid, int& hi) {
  int lo;
  if (s[0] > s[1]) { hi = 0; lo = 1; } else { lo = 0; hi = 1; }
  if (s[2] < s[lo]) { mid = lo; lo = 2; }
  else if (s[2] > s[hi]) { mid = hi; hi = 2; }
  else mid = 2;
}

// the 8 corners of a box in the reference's order (computeVertices): signs of (e0, e1, e2) per corner
__host__ __device__ inline void box_corners(const double* A, const double* To, const double* ext, double (*v)[3]) {
  double e0[3], e1[3], e2[3];
  for (int k = 0; k < 3; ++k) {
    e0[k] = A[3 * k] * ext[0];
    e1[k] = A[3 * k + 1] * ext[1];
    e2[k] = A[3 * k + 2] * ext[2];
  }
  // corner i: To (+/-) e0 (+/-) e1 (+/-) e2, evaluated left to right
  const int s0[8] = {-1, 1, 1, -1, -1, 1, 1, -1}, s1[8] = {-1, -1, 1, 1, -1, -1, 1, 1}, s2[8] = {-1, -1, -1, -1, 1, 1, 1, 1};
  for (int i = 0; i < 8; ++i)
    for (int k = 0; k < 3; ++k) {
      double x = To[k];
      x = s0[i] > 0 ? x + e0[k] : x - e0[k];
      x = s1[i] > 0 ? x + e1[k] : x - e1[k];
      x = s2[i] > 0 ? x + e2[k] : x - e2[k];
      v[i][k] = x;
    }
}

// Eigen: Quaternion(Matrix3) — coefficients (x, y, z, w)
__host__ __device__ inline void quat_of_matrix(const double* A, double q[4]) {
  double t = (A[0] + A[4]) + A[8];
  if (t > 0) {
    t = sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (A[7] - A[5]) * t;
    q[1] = (A[2] - A[6]) * t;
    q[2] = (A[3] - A[1]) * t;
  } else {
    int i = 0;
    if (A[4] > A[0]) i = 1;
    if (A[8] > A[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(((A[4 * i] - A[4 * j]) - A[4 * k]) + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (A[3 * k + j] - A[3 * j + k]) * t;
    q[j] = (A[3 * j + i] + A[3 * i + j]) * t;
    q[k] = (A[3 * k + i] + A[3 * i + k]) * t;
  }
}
// Eigen: QuaternionBase::toRotationMatrix
__host__ __device__ inline void matrix_of_quat(const double q[4], double* R) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz);
  R[1] = txy - twz;
  R[2] = txz + twy;
  R[3] = txy + twz;
  R[4] = 1 - (txx + tzz);
  R[5] = tyz - twx;
  R[6] = txz - twy;
  R[7] = tyz + twx;
  R[8] = 1 - (txx + tyy);
}

}  // namespace merge_detail

// leaf: closed-form fit of one triangle (p: 9 doubles)
__host__ __device__ inline void fit3_obbrss(const double* p, NodeBV& f) {
  using namespace merge_detail;
  const double* p1 = p;
  const double* p2 = p + 3;
  const double* p3 = p + 6;
  double e[3][3], len[3];
  for (int k = 0; k < 3; ++k) {
    e[0][k] = p1[k] - p2[k];
    e[1][k] = p2[k] - p3[k];
    e[2][k] = p3[k] - p1[k];
  }
  for (int i = 0; i < 3; ++i) len[i] = dot3p(e[i], e[i]);
  int imax = 0;
  if (len[1] > len[0]) imax = 1;
  if (len[2] > len[imax]) imax = 2;
  double c2[3] = {e[0][1] * e[1][2] - e[0][2] * e[1][1], e[0][2] * e[1][0] - e[0][0] * e[1][2], e[0][0] * e[1][1] - e[0][1] * e[1][0]};
  normalize3(c2);
  double c0[3] = {e[imax][0], e[imax][1], e[imax][2]};
  normalize3(c0);
  const double c1[3] = {c2[1] * c0[2] - c2[2] * c0[1], c2[2] * c0[0] - c2[0] * c0[2], c2[0] * c0[1] - c2[1] * c0[0]};
  set_col(f.axis, 0, c0);
  set_col(f.axis, 1, c1);
  set_col(f.axis, 2, c2);
  for (int k = 0; k < 9; ++k) f.rss_axis[k] = f.axis[k];
  auto pt = [&](int j) { return p + 3 * j; };
  extent_center_from_points(pt, 3, f.axis, f.obb_To, f.obb_ext);
  rss_from_points(pt, 3, f.rss_axis, f.rss_To, f.rss_l, f.rss_r);
}

// inner node: a = left child (`*this` of the reference's operator+), b = right child (`other`)
__host__ __device__ inline void merge_obbrss(const NodeBV& a, const NodeBV& b, NodeBV& out) {
  using namespace merge_detail;
  double v[16][3];
  // ---------------- OBB ----------------
  {
    double diff[3] = {a.obb_To[0] - b.obb_To[0], a.obb_To[1] - b.obb_To[1], a.obb_To[2] - b.obb_To[2]};
    const double ma = fmax(fmax(a.obb_ext[0], a.obb_ext[1]), a.obb_ext[2]);
    const double mb = fmax(fmax(b.obb_ext[0], b.obb_ext[1]), b.obb_ext[2]);
    if (sqrt(dot3p(diff, diff)) > 2 * (ma + mb)) {
      // merge_largedist: first axis along the centre difference, the other two from the covariance of the 16 corners
      // projected on the plane normal to it
      box_corners(a.axis, a.obb_To, a.obb_ext, v);
      box_corners(b.axis, b.obb_To, b.obb_ext, v + 8);
      normalize3(diff);
      double pr[16][3];
      for (int i = 0; i < 16; ++i) {
        const double d = dot3p(v[i], diff);
        for (int k = 0; k < 3; ++k) pr[i][k] = v[i][k] - diff[k] * d;
      }
      double M[3][3], s[3], V[3][3];
      covariance_of_points(pr, 16, M);
      jacobi3(M, s, V);
      int mid, hi;
      order_eigenvalues(s, mid, hi);
      const double a1[3] = {V[0][hi], V[1][hi], V[2][hi]}, a2[3] = {V[0][mid], V[1][mid], V[2][mid]};
      set_col(out.axis, 0, diff);
      set_col(out.axis, 1, a1);
      set_col(out.axis, 2, a2);
      auto pt = [&](int j) { return (const double*)v[j]; };
      extent_center_from_points(pt, 16, out.axis, out.obb_To, out.obb_ext);
    } else {
      // merge_smalldist: average orientation (normalised quaternion sum), extents over the corners of both boxes
      double To[3] = {(a.obb_To[0] + b.obb_To[0]) * 0.5, (a.obb_To[1] + b.obb_To[1]) * 0.5, (a.obb_To[2] + b.obb_To[2]) * 0.5};
      double q0[4], q1[4], q[4];
      quat_of_matrix(a.axis, q0);
      quat_of_matrix(b.axis, q1);
      if ((((q0[0] * q1[0] + q0[1] * q1[1]) + q0[2] * q1[2]) + q0[3] * q1[3]) < 0)
        for (int k = 0; k < 4; ++k) q1[k] = -q1[k];
      for (int k = 0; k < 4; ++k) q[k] = q0[k] + q1[k];
      const double z = ((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3];
      if (z > 0) {
        const double n = sqrt(z);
        for (int k = 0; k < 4; ++k) q[k] /= n;
      }
      matrix_of_quat(q, out.axis);
      double pmin[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, pmax[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
      box_corners(a.axis, a.obb_To, a.obb_ext, v);
      box_corners(b.axis, b.obb_To, b.obb_ext, v + 8);
      for (int i = 0; i < 16; ++i) {
        const double d3[3] = {v[i][0] - To[0], v[i][1] - To[1], v[i][2] - To[2]};
        for (int j = 0; j < 3; ++j) {
          const double d = FCL_SUM3(d3[0] * out.axis[j], d3[1] * out.axis[3 + j], d3[2] * out.axis[6 + j]);
          // reference quirk (OBB-inl.h:339-342): `else if` -- a point that raises the maximum is not tried as a minimum,
          // so the very first corner never lowers pmin
          if (d > pmax[j]) pmax[j] = d;
          else if (d < pmin[j]) pmin[j] = d;
        }
      }
      for (int j = 0; j < 3; ++j) {
        const double h = 0.5 * (pmax[j] + pmin[j]);
        for (int k = 0; k < 3; ++k) To[k] += out.axis[3 * k + j] * h;
        out.obb_ext[j] = 0.5 * (pmax[j] - pmin[j]);
      }
      for (int k = 0; k < 3; ++k) out.obb_To[k] = To[k];
    }
  }
  // ---------------- RSS ----------------
  {
    // corners of the boxes around the two swept rectangles: `other` first (v[0..7]), then `*this` (v[8..15])
    for (int w = 0; w < 2; ++w) {
      const NodeBV& s = (w == 0) ? b : a;
      double c0[3], c1[3], c2[3];
      get_col(s.rss_axis, 0, c0);
      get_col(s.rss_axis, 1, c1);
      get_col(s.rss_axis, 2, c2);
      const double f0p = s.rss_l[0] + s.rss_r, f1p = s.rss_l[1] + s.rss_r, fn = -s.rss_r;
      for (int i = 0; i < 8; ++i)
        for (int k = 0; k < 3; ++k) {
          const double d0 = (i & 4) ? c0[k] * fn : c0[k] * f0p;
          const double d1 = (i & 2) ? c1[k] * fn : c1[k] * f1p;
          const double d2 = (i & 1) ? c2[k] * fn : c2[k] * s.rss_r;
          v[8 * w + i][k] = ((s.rss_To[k] + d0) + d1) + d2;
        }
    }
    double M[3][3], sv[3], V[3][3];
    covariance_of_points(v, 16, M);
    jacobi3(M, sv, V);
    int mid, hi;
    order_eigenvalues(sv, mid, hi);
    // reference quirks (RSS-inl.h:362-364): the in-plane axes are COLUMNS of eigen_old's output matrix, which holds the
    // eigenvectors in its rows (vout.col(k) = v[k][*]), i.e. ROWS of the Jacobi matrix; and the third axis is the cross
    // product of *this*'s old in-plane axes, not of the new ones
    const double a0[3] = {V[hi][0], V[hi][1], V[hi][2]}, a1[3] = {V[mid][0], V[mid][1], V[mid][2]};
    double t0[3], t1[3];
    get_col(a.rss_axis, 0, t0);
    get_col(a.rss_axis, 1, t1);
    const double a2[3] = {t0[1] * t1[2] - t0[2] * t1[1], t0[2] * t1[0] - t0[0] * t1[2], t0[0] * t1[1] - t0[1] * t1[0]};
    set_col(out.rss_axis, 0, a0);
    set_col(out.rss_axis, 1, a1);
    set_col(out.rss_axis, 2, a2);
    auto pt = [&](int j) { return (const double*)v[j]; };
    rss_from_points(pt, 16, out.rss_axis, out.rss_To, out.rss_l, out.rss_r);
  }
}

}  // namespace fclgpu
