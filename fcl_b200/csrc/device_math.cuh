// device_math.cuh — FP64 BV-pair and triangle-pair tests evaluated in registers.
//
// Arithmetic contract: IEEE binary64, every multiply and add rounded separately
// (the translation unit is compiled with -fmad=false; SASS is checked for the
// absence of DFMA outside the correctly-rounded div/sqrt sequences), three-term
// sums associated left to right.  This is the same contract the reference's
// default (non-FMA) x86-64 build follows for hand-written expressions.
//
// What each routine implements (reference file:line, /root/reference/...):
//   obb_disjoint      include/fcl/math/bv/OBB-inl.h:399-523
//   rect_distance     include/fcl/math/bv/RSS-inl.h:450-509, 513-1225
//   tri_intersect     include/fcl/narrowphase/detail/traversal/collision/intersect-inl.h:597-617,727-885,1032-1106
//   tri_distance      include/fcl/narrowphase/detail/primitive_shape_algorithm/triangle_distance-inl.h:55-394
#pragma once
#include <cuda_runtime.h>

#include <cmath>

#include "sum_order.h"

namespace fclgpu {

struct V3 {
  double x, y, z;
};

#define FD __host__ __device__ __forceinline__

FD V3 mk(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
FD V3 operator+(const V3& a, const V3& b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
FD V3 operator-(const V3& a, const V3& b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
FD V3 operator*(const V3& a, double s) { return mk(a.x * s, a.y * s, a.z * s); }
FD double dot(const V3& a, const V3& b) { return FCL_SUM3(a.x * b.x, a.y * b.y, a.z * b.z); }
FD V3 cross(const V3& a, const V3& b) {
  return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
FD double dmin(double a, double b) { return (b < a) ? b : a; }  // std::min
FD double dmax(double a, double b) { return (a < b) ? b : a; }  // std::max
FD double dabs(double x) { return (x < 0.0) ? -x : x; }

// 3x3 matrix, row-major: m[3*r+c]
struct M3 {
  double m[9];
};
FD double dot3(double a0, double a1, double a2, double b0, double b1, double b2) { return FCL_SUM3(a0 * b0, a1 * b1, a2 * b2); }
FD V3 mulv(const M3& A, const V3& v) {  // A v
  return mk(dot3(A.m[0], A.m[1], A.m[2], v.x, v.y, v.z), dot3(A.m[3], A.m[4], A.m[5], v.x, v.y, v.z),
            dot3(A.m[6], A.m[7], A.m[8], v.x, v.y, v.z));
}
FD V3 mulTv(const M3& A, const V3& v) {  // A^T v
  return mk(dot3(A.m[0], A.m[3], A.m[6], v.x, v.y, v.z), dot3(A.m[1], A.m[4], A.m[7], v.x, v.y, v.z),
            dot3(A.m[2], A.m[5], A.m[8], v.x, v.y, v.z));
}
FD M3 mulMM(const M3& A, const M3& B) {  // A B
  M3 C;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      C.m[3 * i + j] = dot3(A.m[3 * i], A.m[3 * i + 1], A.m[3 * i + 2], B.m[j], B.m[3 + j], B.m[6 + j]);
  return C;
}
FD M3 mulTM(const M3& A, const M3& B) {  // A^T B
  M3 C;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      C.m[3 * i + j] = dot3(A.m[i], A.m[3 + i], A.m[6 + i], B.m[j], B.m[3 + j], B.m[6 + j]);
  return C;
}

// ---------------------------------------------------------------------------------------
// 15-axis OBB SAT.  B = rotation of box b in a's frame (row-major), T = centre of b in a's
// frame, a/b = half extents.  true = disjoint.  Axis order A0,B0,A1,A2,B1,B2, then A_i x B_j.
// ---------------------------------------------------------------------------------------
FD bool obb_disjoint(const M3& B, const V3& T, const V3& a, const V3& b) {
  const double reps = 1e-6;
  double f[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) f[i] = fabs(B.m[i]) + reps;
  double t, s;

  if (dabs(T.x) > (a.x + dot3(f[0], f[1], f[2], b.x, b.y, b.z))) return true;                // A0
  s = dot3(B.m[0], B.m[3], B.m[6], T.x, T.y, T.z);
  if (dabs(s) > (b.x + dot3(f[0], f[3], f[6], a.x, a.y, a.z))) return true;                  // B0
  if (dabs(T.y) > (a.y + dot3(f[3], f[4], f[5], b.x, b.y, b.z))) return true;                // A1
  if (dabs(T.z) > (a.z + dot3(f[6], f[7], f[8], b.x, b.y, b.z))) return true;                // A2
  s = dot3(B.m[1], B.m[4], B.m[7], T.x, T.y, T.z);
  if (dabs(s) > (b.y + dot3(f[1], f[4], f[7], a.x, a.y, a.z))) return true;                  // B1
  s = dot3(B.m[2], B.m[5], B.m[8], T.x, T.y, T.z);
  if (dabs(s) > (b.z + dot3(f[2], f[5], f[8], a.x, a.y, a.z))) return true;                  // B2

  // A0 x B0..B2
  s = T.z * B.m[3] - T.y * B.m[6];
  t = ((a.y * f[6] + a.z * f[3]) + b.y * f[2]) + b.z * f[1];
  if (dabs(s) > t) return true;
  s = T.z * B.m[4] - T.y * B.m[7];
  t = ((a.y * f[7] + a.z * f[4]) + b.x * f[2]) + b.z * f[0];
  if (dabs(s) > t) return true;
  s = T.z * B.m[5] - T.y * B.m[8];
  t = ((a.y * f[8] + a.z * f[5]) + b.x * f[1]) + b.y * f[0];
  if (dabs(s) > t) return true;
  // A1 x B0..B2
  s = T.x * B.m[6] - T.z * B.m[0];
  t = ((a.x * f[6] + a.z * f[0]) + b.y * f[5]) + b.z * f[4];
  if (dabs(s) > t) return true;
  s = T.x * B.m[7] - T.z * B.m[1];
  t = ((a.x * f[7] + a.z * f[1]) + b.x * f[5]) + b.z * f[3];
  if (dabs(s) > t) return true;
  s = T.x * B.m[8] - T.z * B.m[2];
  t = ((a.x * f[8] + a.z * f[2]) + b.x * f[4]) + b.y * f[3];
  if (dabs(s) > t) return true;
  // A2 x B0..B2
  s = T.y * B.m[0] - T.x * B.m[3];
  t = ((a.x * f[3] + a.y * f[0]) + b.y * f[8]) + b.z * f[7];
  if (dabs(s) > t) return true;
  s = T.y * B.m[1] - T.x * B.m[4];
  t = ((a.x * f[4] + a.y * f[1]) + b.x * f[8]) + b.z * f[6];
  if (dabs(s) > t) return true;
  s = T.y * B.m[2] - T.x * B.m[5];
  t = ((a.x * f[5] + a.y * f[2]) + b.x * f[7]) + b.y * f[6];
  if (dabs(s) > t) return true;
  return false;
}

// ---------------------------------------------------------------------------------------
// Rectangle-rectangle distance (PQP RectDist).  Rab row-major, Tab, a[2] = sides of
// rectangle A, b[2] = sides of B.
//
// Same arithmetic as the reference's 16 hand-unrolled blocks (RSS-inl.h:513-1225), different
// control structure, chosen so that the 32 lanes of a warp -- each working on a different BV
// pair -- stay converged:
//   1. all 16 (edge of A, edge of B) candidates are classified at once from the projected
//      corner coordinates: a 16-bit mask of open gates and two masks of interval shortcuts
//      (straight-line code, no divisions);
//   2. the first open candidate whose Voronoi conditions hold is found by walking the gate
//      mask; the (rare, ~0.6 per call) in_voronoi evaluations use parameters selected from
//      the candidate index, so lanes on different candidates share the same instructions;
//   3. one shared tail computes the segment-segment closest points and the norm.
// Candidate index k = 4*g + c: group g = (ia,jb) in the reference's order (1,1),(1,0),(0,1),
// (0,0) -- A's edges parallel to A-axis ia against B's edges parallel to B-axis jb -- and
// c = (upper A, upper B), (upper A, lower B), (lower A, upper B), (lower A, lower B).
// ---------------------------------------------------------------------------------------
FD void clip_to_range(double& v, double lo, double hi) {
  if (v < lo) v = lo;
  else if (v > hi) v = hi;
}

FD bool in_voronoi(double a, double b, double Anorm_dot_B, double Anorm_dot_T, double A_dot_B,
                   double A_dot_T, double B_dot_T) {
  if (fabs(Anorm_dot_B) < 1e-7) return false;
  double u = -Anorm_dot_T / Anorm_dot_B;
  clip_to_range(u, 0.0, b);
  double t = u * A_dot_B + A_dot_T;
  clip_to_range(t, 0.0, a);
  double v = t * A_dot_B - B_dot_T;
  if (Anorm_dot_B > 0) {
    if (v > (u + 1e-7)) return true;
  } else {
    if (v < (u - 1e-7)) return true;
  }
  return false;
}

FD int first_set_bit(unsigned m) {
#ifdef __CUDA_ARCH__
  return __ffs((int)m) - 1;
#else
  return __builtin_ctz(m);
#endif
}

// gates / shortcuts of the four candidates of one group, packed into bits [shift, shift+4)
FD void rect_group_masks(double lo0, double lo1, double up0, double up1,      // A: lower edge ends, upper edge ends
                         double blo0, double blo1, double bup0, double bup1,  // B likewise
                         double bq, double ap, int shift, unsigned& gate, unsigned& sc1, unsigned& sc2) {
  double LA_l, LA_u, UA_l, UA_u, LB_l, LB_u, UB_l, UB_u;
  if (lo0 < lo1) { LA_l = lo0; LA_u = lo1; UA_l = up0; UA_u = up1; }
  else           { LA_l = lo1; LA_u = lo0; UA_l = up1; UA_u = up0; }
  if (blo0 < blo1) { LB_l = blo0; LB_u = blo1; UB_l = bup0; UB_u = bup1; }
  else             { LB_l = blo1; LB_u = blo0; UB_l = bup1; UB_u = bup0; }
  const unsigned g = (((UA_u > bq) && (UB_u > ap)) ? 1u : 0u) | (((UA_l < 0) && (LB_u > ap)) ? 2u : 0u) |
                     (((LA_u > bq) && (UB_l < 0)) ? 4u : 0u) | (((LA_l < 0) && (LB_l < 0)) ? 8u : 0u);
  const unsigned s1 = ((UA_l > bq) ? 1u : 0u) | ((UA_u < 0) ? 2u : 0u) | ((LA_l > bq) ? 4u : 0u) | ((LA_u < 0) ? 8u : 0u);
  const unsigned s2 = ((UB_l > ap) ? 1u : 0u) | ((LB_l > ap) ? 2u : 0u) | ((UB_u < 0) ? 4u : 0u) | ((LB_u < 0) ? 8u : 0u);
  gate |= g << shift;
  sc1 |= s1 << shift;
  sc2 |= s2 << shift;
}

FD double rect_distance(const M3& Rab, const V3& Tabv, const double a[2], const double b[2]) {
  const double R00 = Rab.m[0], R01 = Rab.m[1], R02 = Rab.m[2];
  const double R10 = Rab.m[3], R11 = Rab.m[4], R12 = Rab.m[5];
  const double R20 = Rab.m[6], R21 = Rab.m[7];
  const double a0 = a[0], a1 = a[1], b0 = b[0], b1 = b[1];
  const double Tab0 = Tabv.x, Tab1 = Tabv.y, Tab2 = Tabv.z;
  const V3 Tbav = mulTv(Rab, Tabv);
  const double Tba0 = Tbav.x, Tba1 = Tbav.y, Tba2 = Tbav.z;

  const double aA0B0 = a0 * R00, aA0B1 = a0 * R01, aA1B0 = a1 * R10, aA1B1 = a1 * R11;
  const double bA0B0 = b0 * R00, bA1B0 = b0 * R10, bA0B1 = b1 * R01, bA1B1 = b1 * R11;

  // corners of A on B's axes (x: B-axis 0, y: B-axis 1) and of B on A's axes; first letter =
  // position along the rectangle's axis 0 (Lower/Upper), second along its axis 1
  const double ALL_x = -Tba0, ALU_x = ALL_x + aA1B0, AUL_x = ALL_x + aA0B0, AUU_x = ALU_x + aA0B0;
  const double ALL_y = -Tba1, ALU_y = ALL_y + aA1B1, AUL_y = ALL_y + aA0B1, AUU_y = ALU_y + aA0B1;
  const double BLL_x = Tab0, BLU_x = BLL_x + bA0B1, BUL_x = BLL_x + bA0B0, BUU_x = BLU_x + bA0B0;
  const double BLL_y = Tab1, BLU_y = BLL_y + bA1B1, BUL_y = BLL_y + bA1B0, BUU_y = BLU_y + bA1B0;

  unsigned gate = 0, sc1 = 0, sc2 = 0;
  rect_group_masks(ALL_x, ALU_x, AUL_x, AUU_x, BLL_x, BLU_x, BUL_x, BUU_x, b0, a0, 0, gate, sc1, sc2);   // (1,1)
  rect_group_masks(ALL_y, ALU_y, AUL_y, AUU_y, BLL_x, BUL_x, BLU_x, BUU_x, b1, a0, 4, gate, sc1, sc2);   // (1,0)
  rect_group_masks(ALL_x, AUL_x, ALU_x, AUU_x, BLL_y, BLU_y, BUL_y, BUU_y, b0, a1, 8, gate, sc1, sc2);   // (0,1)
  rect_group_masks(ALL_y, AUL_y, ALU_y, AUU_y, BLL_y, BUL_y, BLU_y, BUU_y, b1, a1, 12, gate, sc1, sc2);  // (0,0)

  // candidate-indexed parameters (selects, no memory indexing)
  int hit = -1;
  bool ia = true, jb = true, ua = true, ub = true;
  double la = a1, lb = b1, ap = a0, bq = b0, RiaJ = R11, AdT_U = 0, AdT_L = 0, BdT_U = 0, BdT_L = 0;
  unsigned m = gate;
  while (m != 0u) {
    const int k = first_set_bit(m);
    const int g = k >> 2, c = k & 3;
    ia = g < 2;
    jb = (g & 1) == 0;
    ua = c < 2;
    ub = (c & 1) == 0;
    la = ia ? a1 : a0;
    lb = jb ? b1 : b0;
    ap = ia ? a0 : a1;
    bq = jb ? b0 : b1;
    RiaJ = ia ? (jb ? R11 : R10) : (jb ? R01 : R00);
    const double RiaQ = ia ? (jb ? R10 : R11) : (jb ? R00 : R01);
    const double RpJ = ia ? (jb ? R01 : R00) : (jb ? R11 : R10);
    const double RpQ = ia ? (jb ? R00 : R01) : (jb ? R10 : R11);
    const double Tab_ia = ia ? Tab1 : Tab0, Tab_p = ia ? Tab0 : Tab1;
    const double Tba_jb = jb ? Tba1 : Tba0, Tba_q = jb ? Tba0 : Tba1;
    const double aPQ = ap * RpQ, aPJ = ap * RpJ, bQiQ = bq * RiaQ, bQpQ = bq * RpQ;
    AdT_U = Tab_ia + bQiQ;
    AdT_L = Tab_ia;
    BdT_U = Tba_jb - aPJ;
    BdT_L = Tba_jb;
    bool ok1 = ((sc1 >> k) & 1u) != 0u, ok2 = ((sc2 >> k) & 1u) != 0u;
    if (!ok1) {
      double n_dot_T;
      if (c == 0) n_dot_T = (g == 0) ? (aPQ - bq - Tba_q) : (aPQ - Tba_q - bq);  // RSS-inl.h:588 vs :742,896,1041
      else if (c == 1) n_dot_T = Tba_q - aPQ;
      else if (c == 2) n_dot_T = -Tba_q - bq;
      else n_dot_T = Tba_q;
      ok1 = in_voronoi(lb, la, (c & 1) ? -RiaQ : RiaQ, n_dot_T, RiaJ, ua ? (aPJ - Tba_jb) : (-Tba_jb),
                       ub ? (-Tab_ia - bQiQ) : (-Tab_ia));
    }
    if (ok1 && !ok2) {
      double n_dot_T;
      if (c == 0) n_dot_T = (g == 0) ? (Tab_p + bQpQ - ap) : (Tab_p - ap + bQpQ);  // RSS-inl.h:592 vs :746,900,1045
      else if (c == 1) n_dot_T = Tab_p - ap;
      else if (c == 2) n_dot_T = -Tab_p - bQpQ;
      else n_dot_T = -Tab_p;
      ok2 = in_voronoi(la, lb, ua ? RpJ : -RpJ, n_dot_T, RiaJ, ub ? AdT_U : AdT_L, ua ? BdT_U : BdT_L);
    }
    if (ok1 && ok2) {
      hit = k;
      break;
    }
    m &= m - 1u;
  }

  // shared tail: closest points of the two edge segments (segCoords, RSS-inl.h:458-482) and
  // their difference vector; executed by every lane (lanes without a hit discard the result)
  const double AdT = ub ? AdT_U : AdT_L, BdT = ua ? BdT_U : BdT_L;
  double t, u;
  const double denom = 1 - RiaJ * RiaJ;
  if (denom == 0) t = 0;
  else {
    t = (AdT - BdT * RiaJ) / denom;
    clip_to_range(t, 0.0, la);
  }
  u = t * RiaJ - BdT;
  if (u < 0) {
    u = 0;
    t = AdT;
    clip_to_range(t, 0.0, la);
  } else if (u > lb) {
    u = lb;
    t = u * RiaJ + AdT;
    clip_to_range(t, 0.0, la);
  }
  // D_k = Tab[k] (+ R[k][q]*b[q] if upper B) + R[k][jb]*u - (point on A)_k
  double D0 = Tab0, D1 = Tab1, D2 = Tab2;
  if (ub) {
    D0 = D0 + (jb ? R00 : R01) * bq;
    D1 = D1 + (jb ? R10 : R11) * bq;
    D2 = D2 + (jb ? R20 : R21) * bq;
  }
  D0 = D0 + (jb ? R01 : R00) * u;
  D1 = D1 + (jb ? R11 : R10) * u;
  D2 = D2 + (jb ? R21 : R20) * u;
  if (ia) {
    D1 = D1 - t;
    if (ua) D0 = D0 - ap;
  } else {
    D0 = D0 - t;
    if (ua) D1 = D1 - ap;
  }
  const double edge_dist = sqrt(FCL_SUM3(D0 * D0, D1 * D1, D2 * D2));  // S.norm()

  // no edge pair: separation along the two face normals (RSS-inl.h:1152-1224)
  double sep1, sep2;
  if (Tab2 > 0.0) {
    sep1 = Tab2;
    if (R20 < 0.0) sep1 += b0 * R20;
    if (R21 < 0.0) sep1 += b1 * R21;
  } else {
    sep1 = -Tab2;
    if (R20 > 0.0) sep1 -= b0 * R20;
    if (R21 > 0.0) sep1 -= b1 * R21;
  }
  if (Tba2 < 0) {
    sep2 = -Tba2;
    if (R02 < 0.0) sep2 += a0 * R02;
    if (R12 < 0.0) sep2 += a1 * R12;
  } else {
    sep2 = Tba2;
    if (R02 > 0.0) sep2 -= a0 * R02;
    if (R12 > 0.0) sep2 -= a1 * R12;
  }
  const double sep = (sep1 > sep2 ? sep1 : sep2);
  const double face_dist = (sep > 0 ? sep : 0);
  return hit >= 0 ? edge_dist : face_dist;
}

// ---------------------------------------------------------------------------------------
// Triangle-triangle intersection: 17-axis SAT in model1's frame, everything translated by
// -P1.  Q must already be transformed (Q' = R Q + T).
// ---------------------------------------------------------------------------------------
FD bool axis_overlaps(const V3& ax, const V3& p2, const V3& p3, const V3& q1, const V3& q2, const V3& q3) {
  // p1 is the origin after translation; ax.dot(p1) is an exact (signed) zero
  const double P1 = FCL_SUM3(ax.x * 0.0, ax.y * 0.0, ax.z * 0.0);
  const double P2 = dot(ax, p2), P3 = dot(ax, p3);
  const double Q1 = dot(ax, q1), Q2 = dot(ax, q2), Q3 = dot(ax, q3);
  const double mn1 = dmin(P1, dmin(P2, P3));
  const double mx2 = dmax(Q1, dmax(Q2, Q3));
  if (mn1 > mx2) return false;
  const double mx1 = dmax(P1, dmax(P2, P3));
  const double mn2 = dmin(Q1, dmin(Q2, Q3));
  if (mn2 > mx1) return false;
  return true;
}

FD bool tri_intersect(const V3& P1, const V3& P2, const V3& P3, const V3& Q1, const V3& Q2, const V3& Q3) {
  const V3 p2 = P2 - P1, p3 = P3 - P1;
  const V3 q1 = Q1 - P1, q2 = Q2 - P1, q3 = Q3 - P1;
  // p1 = P1 - P1 = 0, so e1 = p2 - p1 = p2 and e3 = p1 - p3 = -p3 exactly
  const V3 e1 = mk(p2.x - 0.0, p2.y - 0.0, p2.z - 0.0);
  const V3 e2 = p3 - p2;
  const V3 e3 = mk(0.0 - p3.x, 0.0 - p3.y, 0.0 - p3.z);
  const V3 f1 = q2 - q1, f2 = q3 - q2, f3 = q1 - q3;
  const V3 n1 = cross(e1, e2);
  if (!axis_overlaps(n1, p2, p3, q1, q2, q3)) return false;
  const V3 m1 = cross(f1, f2);
  if (!axis_overlaps(m1, p2, p3, q1, q2, q3)) return false;
  if (!axis_overlaps(cross(e1, f1), p2, p3, q1, q2, q3)) return false;
  if (!axis_overlaps(cross(e1, f2), p2, p3, q1, q2, q3)) return false;
  if (!axis_overlaps(cross(e1, f3), p2, p3, q1, q2, q3)) return false;
  if (!axis_overlaps(cross(e2, f1), p2, p3, q1, q2, q3)) return false;
  if (!axis_overlaps(cross(e2, f2), p2, p3, q1, q2, q3)) return false;
  if (!axis_overlaps(cross(e2, f3), p2, p3, q1, q2, q3)) return false;
  if (!axis_overlaps(cross(e3, f1), p2, p3, q1, q2, q3)) return false;
  if (!axis_overlaps(cross(e3, f2), p2, p3, q1, q2, q3)) return false;
  if (!axis_overlaps(cross(e3, f3), p2, p3, q1, q2, q3)) return false;
  if (!axis_overlaps(cross(e1, n1), p2, p3, q1, q2, q3)) return false;
  if (!axis_overlaps(cross(e2, n1), p2, p3, q1, q2, q3)) return false;
  if (!axis_overlaps(cross(e3, n1), p2, p3, q1, q2, q3)) return false;
  if (!axis_overlaps(cross(f1, m1), p2, p3, q1, q2, q3)) return false;
  if (!axis_overlaps(cross(f2, m1), p2, p3, q1, q2, q3)) return false;
  if (!axis_overlaps(cross(f3, m1), p2, p3, q1, q2, q3)) return false;
  return true;
}

// The same 17-axis test with ROLLED loops (compact code: one copy of the axis test per loop instead of 17): the edges of
// each triangle rotate through three register sets, so no operand is selected at run time.  The verdict of a
// separating-axis test does not depend on the order in which the axes are tried (separated iff SOME axis separates;
// every axis is evaluated by the very same expression as above), so this returns exactly what tri_intersect returns;
// only the early exit may come at another axis.  Used where instruction fetch, not arithmetic, limits the kernel
// (ncu: 27 % of the stall samples of the contact kernel were no_instructions with the unrolled version).
FD bool tri_intersect_rolled(const V3& P1, const V3& P2, const V3& P3, const V3& Q1, const V3& Q2, const V3& Q3) {
  const V3 p2 = P2 - P1, p3 = P3 - P1;
  const V3 q1 = Q1 - P1, q2 = Q2 - P1, q3 = Q3 - P1;
  V3 ea = mk(p2.x - 0.0, p2.y - 0.0, p2.z - 0.0);  // e1, e2, e3 exactly as in tri_intersect
  V3 eb = p3 - p2;
  V3 ec = mk(0.0 - p3.x, 0.0 - p3.y, 0.0 - p3.z);
  V3 fa = q2 - q1, fb = q3 - q2, fc = q1 - q3;
  const V3 n1 = cross(ea, eb);
  if (!axis_overlaps(n1, p2, p3, q1, q2, q3)) return false;
  const V3 m1 = cross(fa, fb);
  if (!axis_overlaps(m1, p2, p3, q1, q2, q3)) return false;
#pragma unroll 1
  for (int i = 0; i < 3; ++i) {
#pragma unroll 1
    for (int j = 0; j < 3; ++j) {
      if (!axis_overlaps(cross(ea, fa), p2, p3, q1, q2, q3)) return false;  // e_i x f_j
      const V3 t = fa;
      fa = fb;
      fb = fc;
      fc = t;  // after three steps f is back at f1
    }
    if (!axis_overlaps(cross(ea, n1), p2, p3, q1, q2, q3)) return false;    // e_i x n1
    const V3 t = ea;
    ea = eb;
    eb = ec;
    ec = t;
  }
#pragma unroll 1
  for (int j = 0; j < 3; ++j) {
    if (!axis_overlaps(cross(fa, m1), p2, p3, q1, q2, q3)) return false;    // f_j x m1
    const V3 t = fa;
    fa = fb;
    fb = fc;
    fc = t;
  }
  return true;
}

// Contact information for an intersecting pair (intersect-inl.h:800-842): plane of each
// triangle, deepest vertices of the other one, the shallower side wins, <= 2 points.
FD void triangle_plane(const V3& v1, const V3& v2, const V3& v3, V3& n, double& t) {
  V3 c = cross(v2 - v1, v3 - v1);
  const double sq = FCL_SUM3(c.x * c.x, c.y * c.y, c.z * c.z);  // squaredNorm()
  if (sq > 0) {
    const double len = sqrt(sq);
    n = mk(c.x / len, c.y / len, c.z / len);
    t = dot(n, v1);
  } else {  // degenerate triangle: the reference leaves n,t uninitialised; defined as 0 here
    n = mk(0, 0, 0);
    t = 0;
  }
}

FD void deepest_points(const V3 pts[3], const V3& n, double t, double& depth, V3 out[3], unsigned& num) {
  const double eps = 1e-5;
  double max_depth = -1.7976931348623157e308;
  unsigned nd = 0, n_neg = 0, n_pos = 0, n_zero = 0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double dist = -(dot(n, pts[i]) - t);
    if (dist > eps) n_pos++;
    else if (dist < -eps) n_neg++;
    else n_zero++;
    if (dist > max_depth) {
      max_depth = dist;
      nd = 1;
      out[0] = pts[i];
    } else if (dist + 1e-6 >= max_depth) {
      nd++;
      if (nd == 2) out[1] = pts[i];
      else out[2] = pts[i];
    }
  }
  if (max_depth < -eps) nd = 0;
  if (n_zero == 0 && ((n_neg == 0) || (n_pos == 0))) nd = 0;
  depth = max_depth;
  num = nd;
}

FD void tri_contact_info(const V3 P[3], const V3 Q[3], V3 contacts[2], unsigned& n_contacts, double& depth,
                         V3& normal) {
  V3 n1, n2;
  double t1, t2;
  triangle_plane(P[0], P[1], P[2], n1, t1);
  triangle_plane(Q[0], Q[1], Q[2], n2, t2);
  V3 deep1[3], deep2[3];
  unsigned nd1, nd2;
  double depth1, depth2;
  deepest_points(Q, n1, t1, depth2, deep2, nd2);
  deepest_points(P, n2, t2, depth1, deep1, nd1);
  if (depth1 > depth2) {
    n_contacts = nd2 < 2u ? nd2 : 2u;
    contacts[0] = deep2[0];
    contacts[1] = deep2[1];
    normal = n1;
    depth = depth2;
  } else {
    n_contacts = nd1 < 2u ? nd1 : 2u;
    contacts[0] = deep1[0];
    contacts[1] = deep1[1];
    normal = mk(-n2.x, -n2.y, -n2.z);
    depth = depth1;
  }
}

// ---------------------------------------------------------------------------------------
// Triangle-triangle distance (PQP TriDist).  T2 already in T1's frame.
// ---------------------------------------------------------------------------------------
FD void seg_points(const V3& P, const V3& A, const V3& Q, const V3& B, V3& VEC, V3& X, V3& Y) {
  V3 T = Q - P;
  const double A_dot_A = dot(A, A), B_dot_B = dot(B, B), A_dot_B = dot(A, B);
  const double A_dot_T = dot(A, T), B_dot_T = dot(B, T);
  const double denom = A_dot_A * B_dot_B - A_dot_B * A_dot_B;
  double t = (A_dot_T * B_dot_B - B_dot_T * A_dot_B) / denom;
  if ((t < 0) || isnan(t)) t = 0;
  else if (t > 1) t = 1;
  const double u = (t * A_dot_B - B_dot_T) / B_dot_B;

  if ((u <= 0) || isnan(u)) {
    Y = Q;
    t = A_dot_T / A_dot_A;
    if ((t <= 0) || isnan(t)) {
      X = P;
      VEC = Q - P;
    } else if (t >= 1) {
      X = P + A;
      VEC = Q - X;
    } else {
      X = P + A * t;
      VEC = cross(A, cross(T, A));
    }
  } else if (u >= 1) {
    Y = Q + B;
    t = (A_dot_B + A_dot_T) / A_dot_A;
    if ((t <= 0) || isnan(t)) {
      X = P;
      VEC = Y - P;
    } else if (t >= 1) {
      X = P + A;
      VEC = Y - X;
    } else {
      X = P + A * t;
      T = Y - P;
      VEC = cross(A, cross(T, A));
    }
  } else {
    Y = Q + B * u;
    if ((t <= 0) || isnan(t)) {
      X = P;
      VEC = cross(B, cross(T, B));
    } else if (t >= 1) {
      X = P + A;
      T = Q - X;
      VEC = cross(B, cross(T, B));
    } else {
      X = P + A * t;
      VEC = cross(A, B);
      if (dot(VEC, T) < 0) VEC = VEC * (-1.0);
    }
  }
}

FD V3 sel3(const V3 v[3], int i) { return i == 0 ? v[0] : (i == 1 ? v[1] : v[2]); }

// vertex-face candidate shared by both triangles' planes: F = face triangle with edge
// vectors Fe, O = the other triangle.  Returns true when the closest points are
// (vertex of O, its projection on F's plane); sets shown_disjoint like the reference.
FD bool vertex_face(const V3 F[3], const V3 Fe[3], const V3 O[3], int& shown_disjoint, V3& onF, V3& onO) {
  const V3 N = cross(Fe[0], Fe[1]);
  const double Nl = dot(N, N);
  if (!(Nl > 1e-15)) return false;
  double pr[3];
  pr[0] = dot(F[0] - O[0], N);
  pr[1] = dot(F[0] - O[1], N);
  pr[2] = dot(F[0] - O[2], N);
  int point = -1;
  if ((pr[0] > 0) && (pr[1] > 0) && (pr[2] > 0)) {
    point = (pr[0] < pr[1]) ? 0 : 1;
    if (pr[2] < (point == 0 ? pr[0] : pr[1])) point = 2;
  } else if ((pr[0] < 0) && (pr[1] < 0) && (pr[2] < 0)) {
    point = (pr[0] > pr[1]) ? 0 : 1;
    if (pr[2] > (point == 0 ? pr[0] : pr[1])) point = 2;
  }
  if (point < 0) return false;
  shown_disjoint = 1;
  const V3 Op = sel3(O, point);
  const double prp = point == 0 ? pr[0] : (point == 1 ? pr[1] : pr[2]);
  if (dot(Op - F[0], cross(N, Fe[0])) > 0)
    if (dot(Op - F[1], cross(N, Fe[1])) > 0)
      if (dot(Op - F[2], cross(N, Fe[2])) > 0) {
        onF = Op + N * (prp / Nl);
        onO = Op;
        return true;
      }
  return false;
}

FD double tri_distance(const V3 T1[3], const V3 T2[3], V3& P, V3& Q) {
  V3 minP = mk(0, 0, 0), minQ = mk(0, 0, 0);
  int shown_disjoint = 0;
  const V3 d00 = T1[0] - T2[0];
  double mindd = dot(d00, d00) + 1;

  // edge i of T1 against edge j of T2, i outer / j inner like the reference.  Both loops stay rolled (unrolling
  // the inner one triples the code and was measured 35 % slower: instruction fetch).  The vertices of each
  // triangle rotate through three register sets, so no operand is selected at run time: edge i starts at A0
  // with vector A1 - A0 (the expression of Sv[i]) and A2 is the third vertex; likewise B0, B1 - B0, B2 for T2.
  V3 A0 = T1[0], A1 = T1[1], A2 = T1[2];
#pragma unroll 1
  for (int i = 0; i < 3; ++i) {
    const V3 Ai = A1 - A0;
    V3 B0 = T2[0], B1 = T2[1], B2 = T2[2];
#pragma unroll 1
    for (int j = 0; j < 3; ++j) {
      V3 VEC;
      seg_points(A0, Ai, B0, B1 - B0, VEC, P, Q);
      const V3 V = Q - P;
      const double dd = dot(V, V);
      if (dd <= mindd) {
        minP = P;
        minQ = Q;
        mindd = dd;
        double a = dot(A2 - P, VEC);
        double b = dot(B2 - Q, VEC);
        if ((a <= 0) && (b >= 0)) return sqrt(dd);
        const double p = dot(V, VEC);
        if (a < 0) a = 0;
        if (b > 0) b = 0;
        if ((p - a + b) > 0) shown_disjoint = 1;
      }
      const V3 t = B0;
      B0 = B1;
      B1 = B2;
      B2 = t;
    }
    const V3 t = A0;
    A0 = A1;
    A1 = A2;
    A2 = t;
  }

  {
    V3 Sv[3], Tv[3];
    Sv[0] = T1[1] - T1[0];
    Sv[1] = T1[2] - T1[1];
    Sv[2] = T1[0] - T1[2];
    Tv[0] = T2[1] - T2[0];
    Tv[1] = T2[2] - T2[1];
    Tv[2] = T2[0] - T2[2];
    V3 onF, onO;
    if (vertex_face(T1, Sv, T2, shown_disjoint, onF, onO)) {  // vertex of T2 against T1's face
      P = onF;
      Q = onO;
      const V3 d = P - Q;
      return sqrt(dot(d, d));
    }
    if (vertex_face(T2, Tv, T1, shown_disjoint, onF, onO)) {  // vertex of T1 against T2's face
      P = onO;
      Q = onF;
      const V3 d = P - Q;
      return sqrt(dot(d, d));
    }
  }
  if (shown_disjoint) {
    P = minP;
    Q = minQ;
    return sqrt(mindd);
  }
  return 0;
}

// ---------------------------------------------------------------------------------------
// Sphere vs triangle (mesh <-> sphere collide): sphereTriangleIntersect,
// narrowphase/detail/primitive_shape_algorithm/sphere_triangle-inl.h:85-244.  c = sphere centre, P[3] = triangle,
// same frame.  On a hit: cp = contact point, depth = -(radius - distance) (<= 0), nrm = unit vector from the
// centre to the contact point (the caller negates it, mesh_shape_collision_traversal_node-inl.h:243).
// Structure differs from the oracle's restatement (one edge routine, last in-reach edge wins by a loop),
// arithmetic order is the same.
// ---------------------------------------------------------------------------------------
FD double point_segment_sq(const V3& a, const V3& b, const V3& p, V3& nearest) {
  V3 diff = p - a;
  const V3 v = b - a;
  double t = dot(v, diff);
  if (t > 0) {
    const double vv = dot(v, v);
    if (t < vv) {
      t /= vv;
      diff = diff - v * t;
    } else {
      t = 1;
      diff = diff - v;
    }
  } else {
    t = 0;
  }
  nearest = a + v * t;
  return dot(diff, diff);
}

FD bool sphere_tri_intersect(const V3& c, double radius, const V3 P[3], V3& cp, double& depth, V3& nrm) {
  V3 n = cross(P[1] - P[0], P[2] - P[0]);
  {
    const double len = sqrt(dot(n, n));
    n = mk(n.x / len, n.y / len, n.z / len);
  }
  const double reach = radius + 2.220446049250313e-16;  // radius + epsilon
  double h = dot(c - P[0], n);
  if (h < 0) {
    h *= -1;
    n = n * (-1.0);
  }
  if (!(h < reach)) return false;
  bool touched = false;
  V3 q = mk(0, 0, 0);
  {
    const double r1 = dot(cross(P[1] - P[0], n), c - P[0]);
    const double r2 = dot(cross(P[2] - P[1], n), c - P[1]);
    const double r3 = dot(cross(P[0] - P[2], n), c - P[2]);
    if ((r1 > 0 && r2 > 0 && r3 > 0) || (r1 <= 0 && r2 <= 0 && r3 <= 0)) {
      touched = true;
      q = c - n * h;
    } else {
      const double reach2 = reach * reach;
#pragma unroll 1
      for (int e = 0; e < 3; ++e) {  // edges P1P2, P2P3, P3P1: the last one within reach supplies the point
        V3 near_e;
        const double d2 = point_segment_sq(sel3(P, e), sel3(P, (e + 1) % 3), c, near_e);
        if (d2 < reach2) {
          touched = true;
          q = near_e;
        }
      }
    }
  }
  if (!touched) return false;
  const V3 w = q - c;
  const double d2 = dot(w, w);
  if (!(d2 < reach * reach)) return false;
  cp = q;
  if (d2 > 0) {
    const double d = sqrt(d2);
    nrm = mk(w.x / d, w.y / d, w.z / d);
    depth = -(radius - d);
  } else {
    nrm = n * (-1.0);
    depth = -radius;
  }
  return true;
}

// ---------------------------------------------------------------------------------------
// Halfspace / Plane vs triangle (mesh <-> halfspace / plane collide): halfspaceTriangleIntersect,
// narrowphase/detail/primitive_shape_algorithm/halfspace-inl.h:587-621, and planeTriangleIntersect, plane-inl.h:683-759.
// The shape is given ALREADY TRANSFORMED to the frame of the vertices (n = tf.linear() * n0, d = d0 + n . tf.translation(),
// geometry/shape/halfspace-inl.h:168-180); V[3] = the triangle's vertices in that frame.  Outputs as the reference writes
// them (the traversal node negates the normal).  Structure differs from the oracle's restatement (the plane routine
// keeps the lone vertex and the two others in named registers instead of index arrays), arithmetic order is the same.
// ---------------------------------------------------------------------------------------
FD bool halfspace_tri_intersect(const V3& n, double d, const V3 V[3], V3& cp, double& depth_out, V3& nrm) {
  V3 v = V[0];
  double depth = dot(n, v) - d;
#pragma unroll
  for (int k = 1; k < 3; ++k) {
    const double dk = dot(n, V[k]) - d;
    if (dk < depth) {
      depth = dk;
      v = V[k];
    }
  }
  if (!(depth <= 0)) return false;
  depth_out = -depth;
  nrm = n;
  cp = v - n * (0.5 * depth);
  return true;
}

FD bool plane_tri_intersect(const V3& n, double d, const V3 V[3], V3& cp, double& depth_out, V3& nrm) {
  const double d0 = dot(n, V[0]) - d, d1 = dot(n, V[1]) - d, d2 = dot(n, V[2]) - d;
  if ((d0 >= 0 && d1 >= 0 && d2 >= 0) || (d0 <= 0 && d1 <= 0 && d2 <= 0)) return false;
  const bool p0 = d0 > 0, p1 = d1 > 0, p2 = d2 > 0;
  const int n_positive = (int)p0 + (int)p1 + (int)p2;
  double d_positive = 0, d_negative = 0;
  // running maxima in vertex order, `<=` like the reference
  if (p0) { if (d_positive <= d0) d_positive = d0; } else { if (d_negative <= -d0) d_negative = -d0; }
  if (p1) { if (d_positive <= d1) d_positive = d1; } else { if (d_negative <= -d1) d_negative = -d1; }
  if (p2) { if (d_positive <= d2) d_positive = d2; } else { if (d_negative <= -d2) d_negative = -d2; }
  depth_out = dmin(d_positive, d_negative);
  nrm = (d_positive > d_negative) ? n : mk(-n.x, -n.y, -n.z);
  // the lone vertex q (the only one on its side) and the two others a, b in vertex order
  const bool lone_positive = n_positive == 1;
  const bool q0 = p0 == lone_positive, q1 = p1 == lone_positive;  // is vertex k the lone one?
  const V3 q = q0 ? V[0] : (q1 ? V[1] : V[2]);
  const double qd = q0 ? d0 : (q1 ? d1 : d2);
  const V3 a = q0 ? V[1] : V[0];
  const double ad = q0 ? d1 : d0;
  const V3 b = (q0 || q1) ? V[2] : V[1];
  const double bd = (q0 || q1) ? d2 : d1;
  V3 t1, t2;
  if (n_positive == 2) {  // a, b positive: t = (-p * q_d + q * p_d) / (-q_d + p_d)
    const double den1 = -qd + ad, den2 = -qd + bd;
    t1 = mk(((-a.x) * qd + q.x * ad) / den1, ((-a.y) * qd + q.y * ad) / den1, ((-a.z) * qd + q.z * ad) / den1);
    t2 = mk(((-b.x) * qd + q.x * bd) / den2, ((-b.y) * qd + q.y * bd) / den2, ((-b.z) * qd + q.z * bd) / den2);
  } else {                // a, b not positive: t = (p * q_d - q * p_d) / (q_d - p_d)
    const double den1 = qd - ad, den2 = qd - bd;
    t1 = mk((a.x * qd - q.x * ad) / den1, (a.y * qd - q.y * ad) / den1, (a.z * qd - q.z * ad) / den1);
    t2 = mk((b.x * qd - q.x * bd) / den2, (b.y * qd - q.y * bd) / den2, (b.z * qd - q.z * bd) / den2);
  }
  cp = (t1 + t2) * 0.5;
  return true;
}

// ---------------------------------------------------------------------------------------
// Sphere vs triangle distance (mesh <-> sphere distance): the nearest-point overload of
// sphereTriangleDistance (sphere_triangle-inl.h:469-496) on top of Project::projectTriangle / projectLine
// (math/detail/project-inl.h:54-123).  o = sphere centre, P[3] = triangle, same frame.
// Returns false when the centre is within the radius of the triangle (or the triangle has zero area);
// otherwise d = distance, on_tri = projection of the centre, on_sphere = o - dir * radius.
// Structure differs from the oracle's restatement (edge loop carries the two barycentric weights of the best
// edge and its index instead of a parameterisation array), arithmetic order is the same.
// ---------------------------------------------------------------------------------------
FD bool sphere_tri_distance(const V3& o, double radius, const V3 P[3], double& d, V3& on_sphere, V3& on_tri) {
  const V3 e0 = P[0] - P[1], e1 = P[1] - P[2], e2 = P[2] - P[0];
  const V3 E[3] = {e0, e1, e2};
  const V3 n = cross(e0, e1);
  const double l = dot(n, n);
  if (!(l > 0)) return false;
  double best = -1, wa = 0, wb = 0;
  int best_i = -1;
#pragma unroll 1
  for (int i = 0; i < 3; ++i) {
    const V3 a = sel3(P, i);
    if (dot(a - o, cross(sel3(E, i), n)) > 0) {  // outside edge i: candidate = closest point of that edge
      const V3 b = sel3(P, (i + 1) % 3);
      const V3 dv = b - a;
      const double ll = dot(dv, dv);
      double s = -1, t1 = 0;
      if (ll > 0) {
        const double t = dot(o - a, dv);
        if (t >= ll) {
          t1 = 1;
          const V3 w = o - b;
          s = dot(w, w);
        } else if (t <= 0) {
          t1 = 0;
          const V3 w = o - a;
          s = dot(w, w);
        } else {
          t1 = t / ll;
          const V3 w = (a + dv * t1) - o;
          s = dot(w, w);
        }
      }
      if (best < 0 || s < best) {
        best = s;
        best_i = i;
        wa = (ll > 0) ? 1 - t1 : 0.0;
        wb = t1;
      }
    }
  }
  double w0, w1, w2;
  if (best < 0) {  // the projection falls inside the triangle
    const double h = dot(P[0] - o, n);
    const double s = sqrt(l);
    const V3 pp = n * (h / l);
    best = dot(pp, pp);
    const V3 c1 = cross(e1, (P[1] - o) - pp), c2 = cross(e2, (P[2] - o) - pp);
    w0 = sqrt(dot(c1, c1)) / s;
    w1 = sqrt(dot(c2, c2)) / s;
    w2 = 1 - w0 - w1;
  } else {
    w0 = (best_i == 0) ? wa : ((best_i == 2) ? wb : 0.0);
    w1 = (best_i == 1) ? wa : ((best_i == 0) ? wb : 0.0);
    w2 = (best_i == 2) ? wa : ((best_i == 1) ? wb : 0.0);
  }
  if (!(best > radius * radius)) return false;
  d = sqrt(best) - radius;
  on_tri = (P[0] * w0 + P[1] * w1) + P[2] * w2;
  V3 dir = o - on_tri;
  {
    const double len = sqrt(dot(dir, dir));
    dir = mk(dir.x / len, dir.y / len, dir.z / len);
  }
  on_sphere = o - dir * radius;
  return true;
}

// ---------------------------------------------------------------------------------------
// BV-pair tests in the relative pose (R0, T0) of model2 in model1's frame.
//   overlap(R0,T0,OBB,OBB)   include/fcl/math/bv/OBB-inl.h:384-395
//   distance(R0,T0,RSS,RSS)  include/fcl/math/bv/RSS-inl.h:1957-1974
// axis1/axis2 row-major (column c = c-th box axis).
// ---------------------------------------------------------------------------------------
FD bool obb_pair_disjoint(const M3& R0, const V3& T0, const M3& axis1, const V3& To1, const V3& ext1,
                          const M3& axis2, const V3& To2, const V3& ext2) {
  const M3 R0b2 = mulMM(R0, axis2);
  const M3 R = mulTM(axis1, R0b2);
  const V3 Ttemp = (mulv(R0, To2) + T0) - To1;
  const V3 T = mulTv(axis1, Ttemp);
  return obb_disjoint(R, T, ext1, ext2);
}

FD double rss_pair_distance(const M3& R0, const V3& T0, const M3& axis1, const V3& To1, const double l1[2],
                            double r1, const M3& axis2, const V3& To2, const double l2[2], double r2) {
  const M3 R0b2 = mulMM(R0, axis2);
  const M3 R = mulTM(axis1, R0b2);
  const V3 Ttemp = (mulv(R0, To2) + T0) - To1;
  const V3 T = mulTv(axis1, Ttemp);
  double dist = rect_distance(R, T, l1, l2);
  dist -= (r1 + r2);
  return (dist < 0.0) ? 0.0 : dist;
}

}  // namespace fclgpu
