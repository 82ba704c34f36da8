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

namespace fclgpu {

struct V3 {
  double x, y, z;
};

#define FD __host__ __device__ __forceinline__

FD V3 mk(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
FD V3 operator+(const V3& a, const V3& b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
FD V3 operator-(const V3& a, const V3& b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
FD V3 operator*(const V3& a, double s) { return mk(a.x * s, a.y * s, a.z * s); }
FD double dot(const V3& a, const V3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
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
FD double dot3(double a0, double a1, double a2, double b0, double b1, double b2) {
  return (a0 * b0 + a1 * b1) + a2 * b2;
}
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
// Structure (differs from the reference's 16 hand-unrolled blocks, same arithmetic): the
// 16 (edge of A, edge of B) candidates are 4 groups (ia, jb) = (1,1), (1,0), (0,1), (0,0)
// -- A's edges parallel to A-axis ia against B's edges parallel to B-axis jb -- times the
// 4 (upper/lower A edge, upper/lower B edge) combinations.  rect_group<> evaluates one
// group; the first candidate whose Voronoi conditions hold yields the segment-segment
// parameters, and ONE shared tail computes the closest-point vector, so lanes that stop
// at different candidates re-converge for the divisions / sqrt.
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

struct RectHit {
  int ia, jb;     // edge axes
  bool ua, ub;    // upper A edge / upper B edge
  double AdT, BdT;  // segment parameters' right-hand sides
};

// One group.  R = Rab (row-major), Tab, Tba = Rab^T Tab.  Returns true and fills `h` when
// one of the 4 candidates of the group contains the closest points.
template <int IA, int JB>
FD bool rect_group(const double* R, const double* Tab, const double* Tba, const double* a,
                   const double* b, RectHit& h) {
  constexpr int P = 1 - IA;  // A's other axis: index into a[], Tab[], rows of R
  constexpr int Q = 1 - JB;  // B's other axis: index into b[], Tba[], cols of R
  const double RiaQ = R[3 * IA + Q];    // A_ia . B_q
  const double RiaJ = R[3 * IA + JB];   // A_ia . B_jb
  const double RpJ = R[3 * P + JB];     // A_p  . B_jb
  const double RpQ = R[3 * P + Q];      // A_p  . B_q
  const double aPQ = a[P] * RpQ;        // a[p]  * (A_p . B_q)
  const double aPJ = a[P] * RpJ;        // a[p]  * (A_p . B_jb)
  const double aIQ = a[IA] * RiaQ;      // a[ia] * (A_ia . B_q)
  const double bQiQ = b[Q] * RiaQ;      // b[q]  * (A_ia . B_q)
  const double bQpQ = b[Q] * RpQ;       // b[q]  * (A_p . B_q)
  const double bJpJ = b[JB] * RpJ;      // b[jb] * (A_p . B_jb)

  // A's four corners projected on B-axis q (origin of B at 0), B's on A-axis p.
  // Corner naming: first letter = position along axis 0 (L/U), second along axis 1.
  const double ALL = -Tba[Q];
  const double A_step1 = a[1] * R[3 * 1 + Q], A_step0 = a[0] * R[3 * 0 + Q];
  const double ALU = ALL + A_step1, AUL = ALL + A_step0, AUU = ALU + A_step0;
  const double BLL = Tab[P];
  const double B_step1 = b[1] * R[3 * P + 1], B_step0 = b[0] * R[3 * P + 0];
  const double BLU = BLL + B_step1, BUL = BLL + B_step0, BUU = BLU + B_step0;

  // lower / upper edge of A parallel to axis IA, each as an ordered interval [l,u]
  double LA_l, LA_u, UA_l, UA_u;
  {
    const double lo0 = ALL, lo1 = (IA == 1) ? ALU : AUL;  // the lower edge's two endpoints
    const double up0 = (IA == 1) ? AUL : ALU, up1 = AUU;
    if (lo0 < lo1) { LA_l = lo0; LA_u = lo1; UA_l = up0; UA_u = up1; }
    else           { LA_l = lo1; LA_u = lo0; UA_l = up1; UA_u = up0; }
  }
  double LB_l, LB_u, UB_l, UB_u;
  {
    const double lo0 = BLL, lo1 = (JB == 1) ? BLU : BUL;
    const double up0 = (JB == 1) ? BUL : BLU, up1 = BUU;
    if (lo0 < lo1) { LB_l = lo0; LB_u = lo1; UB_l = up0; UB_u = up1; }
    else           { LB_l = lo1; LB_u = lo0; UB_l = up1; UB_u = up0; }
  }

  const double la = a[IA], lb = b[JB];
  const double AdT_U = Tab[IA] + bQiQ;  // B's upper edge: origin shifted by b[q] along B_q
  const double AdT_L = Tab[IA];
  const double BdT_U = Tba[JB] - aPJ;   // A's upper edge: origin shifted by a[p] along A_p
  const double BdT_L = Tba[JB];

  h.ia = IA;
  h.jb = JB;
  // (upper A, upper B)
  if ((UA_u > b[Q]) && (UB_u > a[P])) {
    // group (1,1) associates these two sums differently from the other three groups
    // (RSS-inl.h:588,592 vs :742,746 / :896,900 / :1041,1045)
    const double x1 = (IA == 1 && JB == 1) ? (aPQ - b[Q] - Tba[Q]) : (aPQ - Tba[Q] - b[Q]);
    const double x2 = (IA == 1 && JB == 1) ? (Tab[P] + bQpQ - a[P]) : (Tab[P] - a[P] + bQpQ);
    if (((UA_l > b[Q]) || in_voronoi(lb, la, RiaQ, x1, RiaJ, aPJ - Tba[JB], -Tab[IA] - bQiQ)) &&
        ((UB_l > a[P]) || in_voronoi(la, lb, RpJ, x2, RiaJ, AdT_U, BdT_U))) {
      h.ua = true; h.ub = true; h.AdT = AdT_U; h.BdT = BdT_U;
      return true;
    }
  }
  // (upper A, lower B)
  if ((UA_l < 0) && (LB_u > a[P])) {
    if (((UA_u < 0) || in_voronoi(lb, la, -RiaQ, Tba[Q] - aPQ, RiaJ, aPJ - Tba[JB], -Tab[IA])) &&
        ((LB_l > a[P]) || in_voronoi(la, lb, RpJ, Tab[P] - a[P], RiaJ, AdT_L, BdT_U))) {
      h.ua = true; h.ub = false; h.AdT = AdT_L; h.BdT = BdT_U;
      return true;
    }
  }
  // (lower A, upper B)
  if ((LA_u > b[Q]) && (UB_l < 0)) {
    if (((LA_l > b[Q]) || in_voronoi(lb, la, RiaQ, -Tba[Q] - b[Q], RiaJ, -Tba[JB], -Tab[IA] - bQiQ)) &&
        ((UB_u < 0) || in_voronoi(la, lb, -RpJ, -Tab[P] - bQpQ, RiaJ, AdT_U, BdT_L))) {
      h.ua = false; h.ub = true; h.AdT = AdT_U; h.BdT = BdT_L;
      return true;
    }
  }
  // (lower A, lower B)
  if ((LA_l < 0) && (LB_l < 0)) {
    if (((LA_u < 0) || in_voronoi(lb, la, -RiaQ, Tba[Q], RiaJ, -Tba[JB], -Tab[IA])) &&
        ((LB_u < 0) || in_voronoi(la, lb, -RpJ, -Tab[P], RiaJ, AdT_L, BdT_L))) {
      h.ua = false; h.ub = false; h.AdT = AdT_L; h.BdT = BdT_L;
      return true;
    }
  }
  (void)aIQ; (void)bJpJ;
  return false;
}

FD double rect_distance(const M3& Rab, const V3& Tabv, const double a[2], const double b[2]) {
  const double* R = Rab.m;
  const double Tab[3] = {Tabv.x, Tabv.y, Tabv.z};
  const V3 Tbav = mulTv(Rab, Tabv);
  const double Tba[3] = {Tbav.x, Tbav.y, Tbav.z};

  RectHit h;
  bool found = rect_group<1, 1>(R, Tab, Tba, a, b, h);
  if (!found) found = rect_group<1, 0>(R, Tab, Tba, a, b, h);
  if (!found) found = rect_group<0, 1>(R, Tab, Tba, a, b, h);
  if (!found) found = rect_group<0, 0>(R, Tab, Tba, a, b, h);

  if (found) {
    // shared tail: closest points of the two edge segments (segCoords, RSS-inl.h:458-482)
    const int ia = h.ia, jb = h.jb, p = 1 - ia, q = 1 - jb;
    const double la = a[ia], lb = b[jb];
    const double A_dot_B = R[3 * ia + jb];
    double t, u;
    const double denom = 1 - A_dot_B * A_dot_B;
    if (denom == 0) t = 0;
    else {
      t = (h.AdT - h.BdT * A_dot_B) / denom;
      clip_to_range(t, 0.0, la);
    }
    u = t * A_dot_B - h.BdT;
    if (u < 0) {
      u = 0;
      t = h.AdT;
      clip_to_range(t, 0.0, la);
    } else if (u > lb) {
      u = lb;
      t = u * A_dot_B + h.AdT;
      clip_to_range(t, 0.0, la);
    }
    // D_k = Tab[k] (+ R[k][q]*b[q] if upper B) + R[k][jb]*u - (point on A)_k
    double D[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double d = Tab[k];
      if (h.ub) d = d + R[3 * k + q] * b[q];
      d = d + R[3 * k + jb] * u;
      D[k] = d;
    }
    D[ia] = D[ia] - t;
    if (h.ua) D[p] = D[p] - a[p];
    return sqrt((D[0] * D[0] + D[1] * D[1]) + D[2] * D[2]);
  }

  // no edge pair: separation along the two face normals (RSS-inl.h:1152-1224)
  double sep1, sep2;
  if (Tab[2] > 0.0) {
    sep1 = Tab[2];
    if (R[6] < 0.0) sep1 += b[0] * R[6];
    if (R[7] < 0.0) sep1 += b[1] * R[7];
  } else {
    sep1 = -Tab[2];
    if (R[6] > 0.0) sep1 -= b[0] * R[6];
    if (R[7] > 0.0) sep1 -= b[1] * R[7];
  }
  if (Tba[2] < 0) {
    sep2 = -Tba[2];
    if (R[2] < 0.0) sep2 += a[0] * R[2];
    if (R[5] < 0.0) sep2 += a[1] * R[5];
  } else {
    sep2 = Tba[2];
    if (R[2] > 0.0) sep2 -= a[0] * R[2];
    if (R[5] > 0.0) sep2 -= a[1] * R[5];
  }
  const double sep = (sep1 > sep2 ? sep1 : sep2);
  return (sep > 0 ? sep : 0);
}

// ---------------------------------------------------------------------------------------
// Triangle-triangle intersection: 17-axis SAT in model1's frame, everything translated by
// -P1.  Q must already be transformed (Q' = R Q + T).
// ---------------------------------------------------------------------------------------
FD bool axis_overlaps(const V3& ax, const V3& p2, const V3& p3, const V3& q1, const V3& q2, const V3& q3) {
  // p1 is the origin after translation; ax.dot(p1) is an exact (signed) zero
  const double P1 = (ax.x * 0.0 + ax.y * 0.0) + ax.z * 0.0;
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

// Contact information for an intersecting pair (intersect-inl.h:800-842): plane of each
// triangle, deepest vertices of the other one, the shallower side wins, <= 2 points.
FD void triangle_plane(const V3& v1, const V3& v2, const V3& v3, V3& n, double& t) {
  V3 c = cross(v2 - v1, v3 - v1);
  const double sq = (c.x * c.x + c.y * c.y) + c.z * c.z;
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
  V3 Sv[3], Tv[3];
  Sv[0] = T1[1] - T1[0];
  Sv[1] = T1[2] - T1[1];
  Sv[2] = T1[0] - T1[2];
  Tv[0] = T2[1] - T2[0];
  Tv[1] = T2[2] - T2[1];
  Tv[2] = T2[0] - T2[2];

  V3 minP = mk(0, 0, 0), minQ = mk(0, 0, 0);
  int shown_disjoint = 0;
  const V3 d00 = T1[0] - T2[0];
  double mindd = dot(d00, d00) + 1;

#pragma unroll 1
  for (int ij = 0; ij < 9; ++ij) {
    const int i = ij / 3, j = ij - 3 * i;
    V3 VEC;
    seg_points(sel3(T1, i), sel3(Sv, i), sel3(T2, j), sel3(Tv, j), VEC, P, Q);
    const V3 V = Q - P;
    const double dd = dot(V, V);
    if (dd <= mindd) {
      minP = P;
      minQ = Q;
      mindd = dd;
      const int i2 = (i + 2) % 3, j2 = (j + 2) % 3;
      double a = dot(sel3(T1, i2) - P, VEC);
      double b = dot(sel3(T2, j2) - Q, VEC);
      if ((a <= 0) && (b >= 0)) return sqrt(dd);
      const double p = dot(V, VEC);
      if (a < 0) a = 0;
      if (b > 0) b = 0;
      if ((p - a + b) > 0) shown_disjoint = 1;
    }
  }

  {
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
