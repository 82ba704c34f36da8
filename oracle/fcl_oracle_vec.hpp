// fcl_oracle_vec.hpp — CPU ORACLE (test infrastructure, NOT product code).
// 3-vector / 3x3-matrix helpers with ONE fixed evaluation order.  The reference
// routes these through Eigen expression templates (include/fcl/common/types.h:
// 70-92); Eigen is not available here, so the canonical order is: every
// three-term sum is evaluated left to right, (x0 + x1) + x2, products rounded
// separately (build with -ffp-contract=off).  `sum3` is the one place that fixes
// that order; -DFCL_SUM3_ORDER=1 selects x0 + (x1 + x2) instead, and the product's
// device math (csrc/sum_order.h `FCL_SUM3`) honours the same macro, so both sides can
// be flipped together and re-verified (tests/test_sum_order_hook.py).
#pragma once
#include <cmath>

#include "fcl_oracle.hpp"

namespace oracle {

#ifndef FCL_SUM3_ORDER
#define FCL_SUM3_ORDER 0
#endif
static inline double sum3(double x0, double x1, double x2) {
#if FCL_SUM3_ORDER == 0
  return (x0 + x1) + x2;
#else
  return x0 + (x1 + x2);
#endif
}

static inline Vec3 add(const Vec3& a, const Vec3& b) { return Vec3{{a[0] + b[0], a[1] + b[1], a[2] + b[2]}}; }
static inline Vec3 sub(const Vec3& a, const Vec3& b) { return Vec3{{a[0] - b[0], a[1] - b[1], a[2] - b[2]}}; }
static inline Vec3 scale(const Vec3& a, double s) { return Vec3{{a[0] * s, a[1] * s, a[2] * s}}; }
static inline double dot(const Vec3& a, const Vec3& b) { return sum3(a[0] * b[0], a[1] * b[1], a[2] * b[2]); }
static inline double sqnorm(const Vec3& a) { return sum3(a[0] * a[0], a[1] * a[1], a[2] * a[2]); }
static inline double norm(const Vec3& a) { return std::sqrt(sqnorm(a)); }
static inline Vec3 cross(const Vec3& a, const Vec3& b) {
  return Vec3{{a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]}};
}

// A * B
static inline Mat3 mul(const Mat3& A, const Mat3& B) {
  Mat3 C;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      C.m[i][j] = sum3(A.m[i][0] * B.m[0][j], A.m[i][1] * B.m[1][j], A.m[i][2] * B.m[2][j]);
  return C;
}
// A^T * B
static inline Mat3 mulTN(const Mat3& A, const Mat3& B) {
  Mat3 C;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      C.m[i][j] = sum3(A.m[0][i] * B.m[0][j], A.m[1][i] * B.m[1][j], A.m[2][i] * B.m[2][j]);
  return C;
}
// A * v
static inline Vec3 mul(const Mat3& A, const Vec3& v) {
  Vec3 r;
  for (int i = 0; i < 3; ++i) r[i] = sum3(A.m[i][0] * v[0], A.m[i][1] * v[1], A.m[i][2] * v[2]);
  return r;
}
// A^T * v   (== v^T * A)
static inline Vec3 mulTv(const Mat3& A, const Vec3& v) {
  Vec3 r;
  for (int j = 0; j < 3; ++j) r[j] = sum3(A.m[0][j] * v[0], A.m[1][j] * v[1], A.m[2][j] * v[2]);
  return r;
}
static inline Vec3 col(const Mat3& A, int j) { return Vec3{{A.m[0][j], A.m[1][j], A.m[2][j]}}; }

}  // namespace oracle
