// fcl_oracle_math.cpp — CPU ORACLE (test infrastructure, NOT product code).
// BV-pair tests and triangle-pair leaf tests of the OBBRSS mesh-mesh path.
// See fcl_oracle.hpp for the parity statement.  Citations: /root/reference/.
#include <cmath>
#include <limits>

#include "fcl_oracle.hpp"
#include "fcl_oracle_vec.hpp"

namespace oracle {

// -----------------------------------------------------------------------------
// OBB separating-axis test — include/fcl/math/bv/OBB-inl.h:399-523.
// B: rotation of box b in box a's frame, T: centre of b in a's frame,
// a, b: half extents.  Returns true when DISJOINT.  15 axes, in the reference's
// order A0,B0,A1,A2,B1,B2, then the nine A_i x B_j; every |B| entry is padded by
// reps=1e-6 (:403-406).
// -----------------------------------------------------------------------------
bool obb_disjoint(const Mat3& B, const Vec3& T, const Vec3& a, const Vec3& b) {
  const double reps = 1e-6;
  double Bf[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Bf[i][j] = std::fabs(B.m[i][j]) + reps;

  auto absd = [](double x) { return (x < 0.0) ? -x : x; };
  auto row_dot_b = [&](int i) { return (Bf[i][0] * b[0] + Bf[i][1] * b[1]) + Bf[i][2] * b[2]; };
  auto col_dot_a = [&](int j) { return (Bf[0][j] * a[0] + Bf[1][j] * a[1]) + Bf[2][j] * a[2]; };
  auto col_dot_T = [&](int j) { return (B.m[0][j] * T[0] + B.m[1][j] * T[1]) + B.m[2][j] * T[2]; };

  // face axes, reference order: A0, B0, A1, A2, B1, B2
  if (absd(T[0]) > (a[0] + row_dot_b(0))) return true;
  if (absd(col_dot_T(0)) > (b[0] + col_dot_a(0))) return true;
  if (absd(T[1]) > (a[1] + row_dot_b(1))) return true;
  if (absd(T[2]) > (a[2] + row_dot_b(2))) return true;
  if (absd(col_dot_T(1)) > (b[1] + col_dot_a(1))) return true;
  if (absd(col_dot_T(2)) > (b[2] + col_dot_a(2))) return true;

  // edge-edge axes A_i x B_j.  With (i1,i2) = the two indices != i in increasing
  // order and (j1,j2) likewise:
  //   s   = T[i2']*B[i1'][j] - T[i1']*B[i2'][j]   (cyclic: i1'=(i+1)%3, i2'=(i+2)%3)
  //   rhs = a[i1]*Bf[i2][j] + a[i2]*Bf[i1][j] + b[j1]*Bf[i][j2] + b[j2]*Bf[i][j1]
  for (int i = 0; i < 3; ++i) {
    const int ic1 = (i + 1) % 3, ic2 = (i + 2) % 3;            // cyclic successors
    const int i1 = (i == 0) ? 1 : 0, i2 = (i == 2) ? 1 : 2;    // increasing others
    for (int j = 0; j < 3; ++j) {
      const int j1 = (j == 0) ? 1 : 0, j2 = (j == 2) ? 1 : 2;
      double s = T[ic2] * B.m[ic1][j] - T[ic1] * B.m[ic2][j];
      double rhs = ((a[i1] * Bf[i2][j] + a[i2] * Bf[i1][j]) + b[j1] * Bf[i][j2]) + b[j2] * Bf[i][j1];
      if (absd(s) > rhs) return true;
    }
  }
  return false;
}

// overlap(R0,T0,OBB,OBB) — OBB-inl.h:384-395 (via OBBRSS-inl.h:155-160).
bool obb_overlap(const Mat3& R0, const Vec3& T0, const Node& n1, const Node& n2) {
  Mat3 R0b2 = mul(R0, n2.axis);
  Mat3 R = mulTN(n1.axis, R0b2);
  Vec3 Ttemp = sub(add(mul(R0, n2.obb_To), T0), n1.obb_To);
  Vec3 T = mulTv(n1.axis, Ttemp);  // Ttemp^T * axis
  return !obb_disjoint(R, T, n1.obb_ext, n2.obb_ext);
}

// -----------------------------------------------------------------------------
// RSS rectangle distance — include/fcl/math/bv/RSS-inl.h:450-509 (helpers) and
// :513-1225 (rectDistance, P=Q=nullptr on this path).
// -----------------------------------------------------------------------------
static inline void clip_to_range(double& val, double lo, double hi) {  // :450-454
  if (val < lo) val = lo;
  else if (val > hi) val = hi;
}

static inline void seg_coords(double& t, double& u, double a, double b, double A_dot_B,
                              double A_dot_T, double B_dot_T) {  // :458-482
  double denom = 1 - A_dot_B * A_dot_B;
  if (denom == 0) t = 0;
  else {
    t = (A_dot_T - B_dot_T * A_dot_B) / denom;
    clip_to_range(t, 0.0, a);
  }
  u = t * A_dot_B - B_dot_T;
  if (u < 0) {
    u = 0;
    t = A_dot_T;
    clip_to_range(t, 0.0, a);
  } else if (u > b) {
    u = b;
    t = u * A_dot_B + A_dot_T;
    clip_to_range(t, 0.0, a);
  }
}

static inline bool in_voronoi(double a, double b, double Anorm_dot_B, double Anorm_dot_T,
                              double A_dot_B, double A_dot_T, double B_dot_T) {  // :486-509
  if (std::fabs(Anorm_dot_B) < 1e-7) return false;
  double t, u, v;
  u = -Anorm_dot_T / Anorm_dot_B;
  clip_to_range(u, 0.0, b);
  t = u * A_dot_B + A_dot_T;
  clip_to_range(t, 0.0, a);
  v = t * A_dot_B - B_dot_T;
  if (Anorm_dot_B > 0) {
    if (v > (u + 1e-7)) return true;
  } else {
    if (v < (u - 1e-7)) return true;
  }
  return false;
}

static inline double norm3(double x, double y, double z) { return std::sqrt(sum3(x * x, y * y, z * z)); }  // Vector3::norm()

double rect_distance(const Mat3& Rab, const Vec3& Tab, const double a[2], const double b[2]) {
  const double (*R)[3] = Rab.m;
  const double A0B0 = R[0][0], A0B1 = R[0][1], A1B0 = R[1][0], A1B1 = R[1][1];

  const double aA0B0 = a[0] * A0B0, aA0B1 = a[0] * A0B1, aA1B0 = a[1] * A1B0, aA1B1 = a[1] * A1B1;
  const double bA0B0 = b[0] * A0B0, bA1B0 = b[0] * A1B0, bA0B1 = b[1] * A0B1, bA1B1 = b[1] * A1B1;

  Vec3 Tba = mulTv(Rab, Tab);  // Rab^T * Tab  (:534)

  double t, u;

  // ---- group 1: A's edges along axis 1  vs  B's edges along axis 1 (:541-697) ----
  const double ALL_x = -Tba[0];
  const double ALU_x = ALL_x + aA1B0;
  const double AUL_x = ALL_x + aA0B0;
  const double AUU_x = ALU_x + aA0B0;

  double LA1_lx, LA1_ux, UA1_lx, UA1_ux;
  if (ALL_x < ALU_x) { LA1_lx = ALL_x; LA1_ux = ALU_x; UA1_lx = AUL_x; UA1_ux = AUU_x; }
  else               { LA1_lx = ALU_x; LA1_ux = ALL_x; UA1_lx = AUU_x; UA1_ux = AUL_x; }

  const double BLL_x = Tab[0];
  const double BLU_x = BLL_x + bA0B1;
  const double BUL_x = BLL_x + bA0B0;
  const double BUU_x = BLU_x + bA0B0;

  double LB1_lx, LB1_ux, UB1_lx, UB1_ux;
  if (BLL_x < BLU_x) { LB1_lx = BLL_x; LB1_ux = BLU_x; UB1_lx = BUL_x; UB1_ux = BUU_x; }
  else               { LB1_lx = BLU_x; LB1_ux = BLL_x; UB1_lx = BUU_x; UB1_ux = BUL_x; }

  // UA1, UB1 (:585)
  if ((UA1_ux > b[0]) && (UB1_ux > a[0])) {
    if (((UA1_lx > b[0]) ||
         in_voronoi(b[1], a[1], A1B0, aA0B0 - b[0] - Tba[0], A1B1, aA0B1 - Tba[1], -Tab[1] - bA1B0)) &&
        ((UB1_lx > a[0]) ||
         in_voronoi(a[1], b[1], A0B1, Tab[0] + bA0B0 - a[0], A1B1, Tab[1] + bA1B0, Tba[1] - aA0B1))) {
      seg_coords(t, u, a[1], b[1], A1B1, Tab[1] + bA1B0, Tba[1] - aA0B1);
      return norm3(Tab[0] + R[0][0] * b[0] + R[0][1] * u - a[0],
                   Tab[1] + R[1][0] * b[0] + R[1][1] * u - t,
                   Tab[2] + R[2][0] * b[0] + R[2][1] * u);
    }
  }
  // UA1, LB1 (:616)
  if ((UA1_lx < 0) && (LB1_ux > a[0])) {
    if (((UA1_ux < 0) ||
         in_voronoi(b[1], a[1], -A1B0, Tba[0] - aA0B0, A1B1, aA0B1 - Tba[1], -Tab[1])) &&
        ((LB1_lx > a[0]) ||
         in_voronoi(a[1], b[1], A0B1, Tab[0] - a[0], A1B1, Tab[1], Tba[1] - aA0B1))) {
      seg_coords(t, u, a[1], b[1], A1B1, Tab[1], Tba[1] - aA0B1);
      return norm3(Tab[0] + R[0][1] * u - a[0],
                   Tab[1] + R[1][1] * u - t,
                   Tab[2] + R[2][1] * u);
    }
  }
  // LA1, UB1 (:644)
  if ((LA1_ux > b[0]) && (UB1_lx < 0)) {
    if (((LA1_lx > b[0]) ||
         in_voronoi(b[1], a[1], A1B0, -Tba[0] - b[0], A1B1, -Tba[1], -Tab[1] - bA1B0)) &&
        ((UB1_ux < 0) ||
         in_voronoi(a[1], b[1], -A0B1, -Tab[0] - bA0B0, A1B1, Tab[1] + bA1B0, Tba[1]))) {
      seg_coords(t, u, a[1], b[1], A1B1, Tab[1] + bA1B0, Tba[1]);
      return norm3(Tab[0] + R[0][0] * b[0] + R[0][1] * u,
                   Tab[1] + R[1][0] * b[0] + R[1][1] * u - t,
                   Tab[2] + R[2][0] * b[0] + R[2][1] * u);
    }
  }
  // LA1, LB1 (:672)
  if ((LA1_lx < 0) && (LB1_lx < 0)) {
    if (((LA1_ux < 0) ||
         in_voronoi(b[1], a[1], -A1B0, Tba[0], A1B1, -Tba[1], -Tab[1])) &&
        ((LB1_ux < 0) ||
         in_voronoi(a[1], b[1], -A0B1, -Tab[0], A1B1, Tab[1], Tba[1]))) {
      seg_coords(t, u, a[1], b[1], A1B1, Tab[1], Tba[1]);
      return norm3(Tab[0] + R[0][1] * u,
                   Tab[1] + R[1][1] * u - t,
                   Tab[2] + R[2][1] * u);
    }
  }

  // ---- group 2: A's edges along axis 1  vs  B's edges along axis 0 (:700-852) ----
  const double ALL_y = -Tba[1];
  const double ALU_y = ALL_y + aA1B1;
  const double AUL_y = ALL_y + aA0B1;
  const double AUU_y = ALU_y + aA0B1;

  double LA1_ly, LA1_uy, UA1_ly, UA1_uy;
  if (ALL_y < ALU_y) { LA1_ly = ALL_y; LA1_uy = ALU_y; UA1_ly = AUL_y; UA1_uy = AUU_y; }
  else               { LA1_ly = ALU_y; LA1_uy = ALL_y; UA1_ly = AUU_y; UA1_uy = AUL_y; }

  double LB0_lx, LB0_ux, UB0_lx, UB0_ux;
  if (BLL_x < BUL_x) { LB0_lx = BLL_x; LB0_ux = BUL_x; UB0_lx = BLU_x; UB0_ux = BUU_x; }
  else               { LB0_lx = BUL_x; LB0_ux = BLL_x; UB0_lx = BUU_x; UB0_ux = BLU_x; }

  // UA1, UB0 (:739)
  if ((UA1_uy > b[1]) && (UB0_ux > a[0])) {
    if (((UA1_ly > b[1]) ||
         in_voronoi(b[0], a[1], A1B1, aA0B1 - Tba[1] - b[1], A1B0, aA0B0 - Tba[0], -Tab[1] - bA1B1)) &&
        ((UB0_lx > a[0]) ||
         in_voronoi(a[1], b[0], A0B0, Tab[0] - a[0] + bA0B1, A1B0, Tab[1] + bA1B1, Tba[0] - aA0B0))) {
      seg_coords(t, u, a[1], b[0], A1B0, Tab[1] + bA1B1, Tba[0] - aA0B0);
      return norm3(Tab[0] + R[0][1] * b[1] + R[0][0] * u - a[0],
                   Tab[1] + R[1][1] * b[1] + R[1][0] * u - t,
                   Tab[2] + R[2][1] * b[1] + R[2][0] * u);
    }
  }
  // UA1, LB0 (:768)
  if ((UA1_ly < 0) && (LB0_ux > a[0])) {
    if (((UA1_uy < 0) ||
         in_voronoi(b[0], a[1], -A1B1, Tba[1] - aA0B1, A1B0, aA0B0 - Tba[0], -Tab[1])) &&
        ((LB0_lx > a[0]) ||
         in_voronoi(a[1], b[0], A0B0, Tab[0] - a[0], A1B0, Tab[1], Tba[0] - aA0B0))) {
      seg_coords(t, u, a[1], b[0], A1B0, Tab[1], Tba[0] - aA0B0);
      return norm3(Tab[0] + R[0][0] * u - a[0],
                   Tab[1] + R[1][0] * u - t,
                   Tab[2] + R[2][0] * u);
    }
  }
  // LA1, UB0 (:796)
  if ((LA1_uy > b[1]) && (UB0_lx < 0)) {
    if (((LA1_ly > b[1]) ||
         in_voronoi(b[0], a[1], A1B1, -Tba[1] - b[1], A1B0, -Tba[0], -Tab[1] - bA1B1)) &&
        ((UB0_ux < 0) ||
         in_voronoi(a[1], b[0], -A0B0, -Tab[0] - bA0B1, A1B0, Tab[1] + bA1B1, Tba[0]))) {
      seg_coords(t, u, a[1], b[0], A1B0, Tab[1] + bA1B1, Tba[0]);
      return norm3(Tab[0] + R[0][1] * b[1] + R[0][0] * u,
                   Tab[1] + R[1][1] * b[1] + R[1][0] * u - t,
                   Tab[2] + R[2][1] * b[1] + R[2][0] * u);
    }
  }
  // LA1, LB0 (:826)
  if ((LA1_ly < 0) && (LB0_lx < 0)) {
    if (((LA1_uy < 0) ||
         in_voronoi(b[0], a[1], -A1B1, Tba[1], A1B0, -Tba[0], -Tab[1])) &&
        ((LB0_ux < 0) ||
         in_voronoi(a[1], b[0], -A0B0, -Tab[0], A1B0, Tab[1], Tba[0]))) {
      seg_coords(t, u, a[1], b[0], A1B0, Tab[1], Tba[0]);
      return norm3(Tab[0] + R[0][0] * u,
                   Tab[1] + R[1][0] * u - t,
                   Tab[2] + R[2][0] * u);
    }
  }

  // ---- group 3: A's edges along axis 0  vs  B's edges along axis 1 (:854-1002) ----
  const double BLL_y = Tab[1];
  const double BLU_y = BLL_y + bA1B1;
  const double BUL_y = BLL_y + bA1B0;
  const double BUU_y = BLU_y + bA1B0;

  double LA0_lx, LA0_ux, UA0_lx, UA0_ux;
  if (ALL_x < AUL_x) { LA0_lx = ALL_x; LA0_ux = AUL_x; UA0_lx = ALU_x; UA0_ux = AUU_x; }
  else               { LA0_lx = AUL_x; LA0_ux = ALL_x; UA0_lx = AUU_x; UA0_ux = ALU_x; }

  double LB1_ly, LB1_uy, UB1_ly, UB1_uy;
  if (BLL_y < BLU_y) { LB1_ly = BLL_y; LB1_uy = BLU_y; UB1_ly = BUL_y; UB1_uy = BUU_y; }
  else               { LB1_ly = BLU_y; LB1_uy = BLL_y; UB1_ly = BUU_y; UB1_uy = BUL_y; }

  // UA0, UB1 (:893)
  if ((UA0_ux > b[0]) && (UB1_uy > a[1])) {
    if (((UA0_lx > b[0]) ||
         in_voronoi(b[1], a[0], A0B0, aA1B0 - Tba[0] - b[0], A0B1, aA1B1 - Tba[1], -Tab[0] - bA0B0)) &&
        ((UB1_ly > a[1]) ||
         in_voronoi(a[0], b[1], A1B1, Tab[1] - a[1] + bA1B0, A0B1, Tab[0] + bA0B0, Tba[1] - aA1B1))) {
      seg_coords(t, u, a[0], b[1], A0B1, Tab[0] + bA0B0, Tba[1] - aA1B1);
      return norm3(Tab[0] + R[0][0] * b[0] + R[0][1] * u - t,
                   Tab[1] + R[1][0] * b[0] + R[1][1] * u - a[1],
                   Tab[2] + R[2][0] * b[0] + R[2][1] * u);
    }
  }
  // UA0, LB1 (:922)
  if ((UA0_lx < 0) && (LB1_uy > a[1])) {
    if (((UA0_ux < 0) ||
         in_voronoi(b[1], a[0], -A0B0, Tba[0] - aA1B0, A0B1, aA1B1 - Tba[1], -Tab[0])) &&
        ((LB1_ly > a[1]) ||
         in_voronoi(a[0], b[1], A1B1, Tab[1] - a[1], A0B1, Tab[0], Tba[1] - aA1B1))) {
      seg_coords(t, u, a[0], b[1], A0B1, Tab[0], Tba[1] - aA1B1);
      return norm3(Tab[0] + R[0][1] * u - t,
                   Tab[1] + R[1][1] * u - a[1],
                   Tab[2] + R[2][1] * u);
    }
  }
  // LA0, UB1 (:950)
  if ((LA0_ux > b[0]) && (UB1_ly < 0)) {
    if (((LA0_lx > b[0]) ||
         in_voronoi(b[1], a[0], A0B0, -b[0] - Tba[0], A0B1, -Tba[1], -bA0B0 - Tab[0])) &&
        ((UB1_uy < 0) ||
         in_voronoi(a[0], b[1], -A1B1, -Tab[1] - bA1B0, A0B1, Tab[0] + bA0B0, Tba[1]))) {
      seg_coords(t, u, a[0], b[1], A0B1, Tab[0] + bA0B0, Tba[1]);
      return norm3(Tab[0] + R[0][0] * b[0] + R[0][1] * u - t,
                   Tab[1] + R[1][0] * b[0] + R[1][1] * u,
                   Tab[2] + R[2][0] * b[0] + R[2][1] * u);
    }
  }
  // LA0, LB1 (:978)
  if ((LA0_lx < 0) && (LB1_ly < 0)) {
    if (((LA0_ux < 0) ||
         in_voronoi(b[1], a[0], -A0B0, Tba[0], A0B1, -Tba[1], -Tab[0])) &&
        ((LB1_uy < 0) ||
         in_voronoi(a[0], b[1], -A1B1, -Tab[1], A0B1, Tab[0], Tba[1]))) {
      seg_coords(t, u, a[0], b[1], A0B1, Tab[0], Tba[1]);
      return norm3(Tab[0] + R[0][1] * u - t,
                   Tab[1] + R[1][1] * u,
                   Tab[2] + R[2][1] * u);
    }
  }

  // ---- group 4: A's edges along axis 0  vs  B's edges along axis 0 (:1006-1150) ----
  double LA0_ly, LA0_uy, UA0_ly, UA0_uy;
  if (ALL_y < AUL_y) { LA0_ly = ALL_y; LA0_uy = AUL_y; UA0_ly = ALU_y; UA0_uy = AUU_y; }
  else               { LA0_ly = AUL_y; LA0_uy = ALL_y; UA0_ly = AUU_y; UA0_uy = ALU_y; }

  double LB0_ly, LB0_uy, UB0_ly, UB0_uy;
  if (BLL_y < BUL_y) { LB0_ly = BLL_y; LB0_uy = BUL_y; UB0_ly = BLU_y; UB0_uy = BUU_y; }
  else               { LB0_ly = BUL_y; LB0_uy = BLL_y; UB0_ly = BUU_y; UB0_uy = BLU_y; }

  // UA0, UB0 (:1038)
  if ((UA0_uy > b[1]) && (UB0_uy > a[1])) {
    if (((UA0_ly > b[1]) ||
         in_voronoi(b[0], a[0], A0B1, aA1B1 - Tba[1] - b[1], A0B0, aA1B0 - Tba[0], -Tab[0] - bA0B1)) &&
        ((UB0_ly > a[1]) ||
         in_voronoi(a[0], b[0], A1B0, Tab[1] - a[1] + bA1B1, A0B0, Tab[0] + bA0B1, Tba[0] - aA1B0))) {
      seg_coords(t, u, a[0], b[0], A0B0, Tab[0] + bA0B1, Tba[0] - aA1B0);
      return norm3(Tab[0] + R[0][1] * b[1] + R[0][0] * u - t,
                   Tab[1] + R[1][1] * b[1] + R[1][0] * u - a[1],
                   Tab[2] + R[2][1] * b[1] + R[2][0] * u);
    }
  }
  // UA0, LB0 (:1067)
  if ((UA0_ly < 0) && (LB0_uy > a[1])) {
    if (((UA0_uy < 0) ||
         in_voronoi(b[0], a[0], -A0B1, Tba[1] - aA1B1, A0B0, aA1B0 - Tba[0], -Tab[0])) &&
        ((LB0_ly > a[1]) ||
         in_voronoi(a[0], b[0], A1B0, Tab[1] - a[1], A0B0, Tab[0], Tba[0] - aA1B0))) {
      seg_coords(t, u, a[0], b[0], A0B0, Tab[0], Tba[0] - aA1B0);
      return norm3(Tab[0] + R[0][0] * u - t,
                   Tab[1] + R[1][0] * u - a[1],
                   Tab[2] + R[2][0] * u);
    }
  }
  // LA0, UB0 (:1095)
  if ((LA0_uy > b[1]) && (UB0_ly < 0)) {
    if (((LA0_ly > b[1]) ||
         in_voronoi(b[0], a[0], A0B1, -Tba[1] - b[1], A0B0, -Tba[0], -Tab[0] - bA0B1)) &&
        ((UB0_uy < 0) ||
         in_voronoi(a[0], b[0], -A1B0, -Tab[1] - bA1B1, A0B0, Tab[0] + bA0B1, Tba[0]))) {
      seg_coords(t, u, a[0], b[0], A0B0, Tab[0] + bA0B1, Tba[0]);
      return norm3(Tab[0] + R[0][1] * b[1] + R[0][0] * u - t,
                   Tab[1] + R[1][1] * b[1] + R[1][0] * u,
                   Tab[2] + R[2][1] * b[1] + R[2][0] * u);
    }
  }
  // LA0, LB0 (:1124)
  if ((LA0_ly < 0) && (LB0_ly < 0)) {
    if (((LA0_uy < 0) ||
         in_voronoi(b[0], a[0], -A0B1, Tba[1], A0B0, -Tba[0], -Tab[0])) &&
        ((LB0_uy < 0) ||
         in_voronoi(a[0], b[0], -A1B0, -Tab[1], A0B0, Tab[0], Tba[0]))) {
      seg_coords(t, u, a[0], b[0], A0B0, Tab[0], Tba[0]);
      return norm3(Tab[0] + R[0][0] * u - t,
                   Tab[1] + R[1][0] * u,
                   Tab[2] + R[2][0] * u);
    }
  }

  // ---- no edge pair: max separation along the two face normals (:1152-1224) ----
  double sep1, sep2;
  if (Tab[2] > 0.0) {
    sep1 = Tab[2];
    if (R[2][0] < 0.0) sep1 += b[0] * R[2][0];
    if (R[2][1] < 0.0) sep1 += b[1] * R[2][1];
  } else {
    sep1 = -Tab[2];
    if (R[2][0] > 0.0) sep1 -= b[0] * R[2][0];
    if (R[2][1] > 0.0) sep1 -= b[1] * R[2][1];
  }
  if (Tba[2] < 0) {
    sep2 = -Tba[2];
    if (R[0][2] < 0.0) sep2 += a[0] * R[0][2];
    if (R[1][2] < 0.0) sep2 += a[1] * R[1][2];
  } else {
    sep2 = Tba[2];
    if (R[0][2] > 0.0) sep2 -= a[0] * R[0][2];
    if (R[1][2] > 0.0) sep2 -= a[1] * R[1][2];
  }
  double sep = (sep1 > sep2 ? sep1 : sep2);
  return (sep > 0 ? sep : 0);
}

// distance(R0,T0,RSS,RSS) — RSS-inl.h:1957-1974 (via OBBRSS-inl.h:164-171).
double rss_distance(const Mat3& R0, const Vec3& T0, const Node& n1, const Node& n2) {
  Mat3 R0b2 = mul(R0, n2.rss_axis);
  Mat3 R = mulTN(n1.rss_axis, R0b2);
  Vec3 Ttemp = sub(add(mul(R0, n2.rss_To), T0), n1.rss_To);
  Vec3 T = mulTv(n1.rss_axis, Ttemp);
  double dist = rect_distance(R, T, n1.rss_l, n2.rss_l);
  dist -= (n1.rss_r + n2.rss_r);
  return (dist < 0.0) ? 0.0 : dist;
}

// -----------------------------------------------------------------------------
// Triangle-triangle intersection —
// include/fcl/narrowphase/detail/traversal/collision/intersect-inl.h:597-617
// (R,T overload), :727-845 (core), :848-885 (computeDeepestPoints),
// :1032-1053 (distanceToPlane, buildTrianglePlane), :1085-1106 (project6).
// -----------------------------------------------------------------------------
static inline bool project6(const Vec3& ax, const Vec3& p1, const Vec3& p2, const Vec3& p3,
                            const Vec3& q1, const Vec3& q2, const Vec3& q3) {
  double P1 = dot(ax, p1), P2 = dot(ax, p2), P3 = dot(ax, p3);
  double Q1 = dot(ax, q1), Q2 = dot(ax, q2), Q3 = dot(ax, q3);
  double mn1 = std::min(P1, std::min(P2, P3));
  double mx2 = std::max(Q1, std::max(Q2, Q3));
  if (mn1 > mx2) return false;
  double mx1 = std::max(P1, std::max(P2, P3));
  double mn2 = std::min(Q1, std::min(Q2, Q3));
  if (mn2 > mx1) return false;
  return true;
}

// normalize() — include/fcl/math/geometry-inl.h:413-425
static inline bool build_triangle_plane(const Vec3& v1, const Vec3& v2, const Vec3& v3, Vec3* n,
                                        double* t) {
  Vec3 n_ = cross(sub(v2, v1), sub(v3, v1));
  double sqr_length = sqnorm(n_);
  if (sqr_length > 0) {
    double len = std::sqrt(sqr_length);
    n_ = Vec3{{n_[0] / len, n_[1] / len, n_[2] / len}};
    *n = n_;
    *t = dot(n_, v1);
    return true;
  }
  return false;
}

static void compute_deepest_points(const Vec3* pts, unsigned num, const Vec3& n, double t,
                                   double* penetration_depth, Vec3* deepest, unsigned* num_deepest) {
  const double eps = 1e-5;  // Intersect::getEpsilon(), :1117-1120
  *num_deepest = 0;
  double max_depth = -std::numeric_limits<double>::max();
  unsigned nd = 0, num_neg = 0, num_pos = 0, num_zero = 0;
  for (unsigned i = 0; i < num; ++i) {
    double dist = -(dot(n, pts[i]) - t);
    if (dist > eps) num_pos++;
    else if (dist < -eps) num_neg++;
    else num_zero++;
    if (dist > max_depth) {
      max_depth = dist;
      nd = 1;
      deepest[nd - 1] = pts[i];
    } else if (dist + 1e-6 >= max_depth) {
      nd++;
      deepest[nd - 1] = pts[i];
    }
  }
  if (max_depth < -eps) nd = 0;
  if (num_zero == 0 && ((num_neg == 0) || (num_pos == 0))) nd = 0;
  *penetration_depth = max_depth;
  *num_deepest = nd;
}

bool tri_intersect(const Vec3 Pin[3], const Vec3 Qin[3], const Mat3& R, const Vec3& T,
                   Vec3* contact_points, unsigned* num_contact_points, double* penetration_depth,
                   Vec3* normal) {
  // Q' = R*Q + T  (:612-614)
  const Vec3 P1 = Pin[0], P2 = Pin[1], P3 = Pin[2];
  const Vec3 Q1 = add(mul(R, Qin[0]), T), Q2 = add(mul(R, Qin[1]), T), Q3 = add(mul(R, Qin[2]), T);

  Vec3 p1 = sub(P1, P1), p2 = sub(P2, P1), p3 = sub(P3, P1);
  Vec3 q1 = sub(Q1, P1), q2 = sub(Q2, P1), q3 = sub(Q3, P1);

  Vec3 e1 = sub(p2, p1), e2 = sub(p3, p2);
  Vec3 n1 = cross(e1, e2);
  if (!project6(n1, p1, p2, p3, q1, q2, q3)) return false;

  Vec3 f1 = sub(q2, q1), f2 = sub(q3, q2);
  Vec3 m1 = cross(f1, f2);
  if (!project6(m1, p1, p2, p3, q1, q2, q3)) return false;

  if (!project6(cross(e1, f1), p1, p2, p3, q1, q2, q3)) return false;
  if (!project6(cross(e1, f2), p1, p2, p3, q1, q2, q3)) return false;
  Vec3 f3 = sub(q1, q3);
  if (!project6(cross(e1, f3), p1, p2, p3, q1, q2, q3)) return false;
  if (!project6(cross(e2, f1), p1, p2, p3, q1, q2, q3)) return false;
  if (!project6(cross(e2, f2), p1, p2, p3, q1, q2, q3)) return false;
  if (!project6(cross(e2, f3), p1, p2, p3, q1, q2, q3)) return false;
  Vec3 e3 = sub(p1, p3);
  if (!project6(cross(e3, f1), p1, p2, p3, q1, q2, q3)) return false;
  if (!project6(cross(e3, f2), p1, p2, p3, q1, q2, q3)) return false;
  if (!project6(cross(e3, f3), p1, p2, p3, q1, q2, q3)) return false;
  if (!project6(cross(e1, n1), p1, p2, p3, q1, q2, q3)) return false;
  if (!project6(cross(e2, n1), p1, p2, p3, q1, q2, q3)) return false;
  if (!project6(cross(e3, n1), p1, p2, p3, q1, q2, q3)) return false;
  if (!project6(cross(f1, m1), p1, p2, p3, q1, q2, q3)) return false;
  if (!project6(cross(f2, m1), p1, p2, p3, q1, q2, q3)) return false;
  if (!project6(cross(f3, m1), p1, p2, p3, q1, q2, q3)) return false;

  if (contact_points && num_contact_points && penetration_depth && normal) {
    // The reference leaves pn1/pn2 uninitialised when a triangle is degenerate
    // (:1045-1051); the oracle (and the GPU) define them as zero in that case.
    Vec3 pn1{{0, 0, 0}}, pn2{{0, 0, 0}};
    double t1 = 0, t2 = 0;
    build_triangle_plane(P1, P2, P3, &pn1, &t1);
    build_triangle_plane(Q1, Q2, Q3, &pn2, &t2);

    Vec3 deepest1[3], deepest2[3];
    unsigned nd1 = 0, nd2 = 0;
    double depth1, depth2;
    Vec3 P[3] = {P1, P2, P3};
    Vec3 Q[3] = {Q1, Q2, Q3};
    compute_deepest_points(Q, 3, pn1, t1, &depth2, deepest2, &nd2);
    compute_deepest_points(P, 3, pn2, t2, &depth1, deepest1, &nd1);

    if (depth1 > depth2) {
      *num_contact_points = std::min(nd2, 2u);
      for (unsigned i = 0; i < *num_contact_points; ++i) contact_points[i] = deepest2[i];
      *normal = pn1;
      *penetration_depth = depth2;
    } else {
      *num_contact_points = std::min(nd1, 2u);
      for (unsigned i = 0; i < *num_contact_points; ++i) contact_points[i] = deepest1[i];
      *normal = Vec3{{-pn2[0], -pn2[1], -pn2[2]}};
      *penetration_depth = depth1;
    }
  }
  return true;
}

// -----------------------------------------------------------------------------
// Triangle-triangle distance (PQP TriDist) —
// include/fcl/narrowphase/detail/primitive_shape_algorithm/triangle_distance-inl.h
// :55-167 (segPoints), :171-394 (triDistance).
// -----------------------------------------------------------------------------
static void seg_points(const Vec3& P, const Vec3& A, const Vec3& Q, const Vec3& B, Vec3& VEC,
                       Vec3& X, Vec3& Y) {
  Vec3 T = sub(Q, P);
  double A_dot_A = dot(A, A), B_dot_B = dot(B, B), A_dot_B = dot(A, B);
  double A_dot_T = dot(A, T), B_dot_T = dot(B, T);
  double t, u;

  double denom = A_dot_A * B_dot_B - A_dot_B * A_dot_B;
  t = (A_dot_T * B_dot_B - B_dot_T * A_dot_B) / denom;
  if ((t < 0) || std::isnan(t)) t = 0;
  else if (t > 1) t = 1;

  u = (t * A_dot_B - B_dot_T) / B_dot_B;

  if ((u <= 0) || std::isnan(u)) {
    Y = Q;
    t = A_dot_T / A_dot_A;
    if ((t <= 0) || std::isnan(t)) {
      X = P;
      VEC = sub(Q, P);
    } else if (t >= 1) {
      X = add(P, A);
      VEC = sub(Q, X);
    } else {
      X = add(P, scale(A, t));
      Vec3 TMP = cross(T, A);
      VEC = cross(A, TMP);
    }
  } else if (u >= 1) {
    Y = add(Q, B);
    t = (A_dot_B + A_dot_T) / A_dot_A;
    if ((t <= 0) || std::isnan(t)) {
      X = P;
      VEC = sub(Y, P);
    } else if (t >= 1) {
      X = add(P, A);
      VEC = sub(Y, X);
    } else {
      X = add(P, scale(A, t));
      T = sub(Y, P);
      Vec3 TMP = cross(T, A);
      VEC = cross(A, TMP);
    }
  } else {
    Y = add(Q, scale(B, u));
    if ((t <= 0) || std::isnan(t)) {
      X = P;
      Vec3 TMP = cross(T, B);
      VEC = cross(B, TMP);
    } else if (t >= 1) {
      X = add(P, A);
      T = sub(Q, X);
      Vec3 TMP = cross(T, B);
      VEC = cross(B, TMP);
    } else {
      X = add(P, scale(A, t));
      VEC = cross(A, B);
      if (dot(VEC, T) < 0) VEC = scale(VEC, -1.0);
    }
  }
}

double tri_distance(const Vec3 T1[3], const Vec3 T2[3], Vec3& P, Vec3& Q) {
  Vec3 Sv[3], Tv[3], VEC;
  Sv[0] = sub(T1[1], T1[0]);
  Sv[1] = sub(T1[2], T1[1]);
  Sv[2] = sub(T1[0], T1[2]);
  Tv[0] = sub(T2[1], T2[0]);
  Tv[1] = sub(T2[2], T2[1]);
  Tv[2] = sub(T2[0], T2[2]);

  Vec3 V, Z;
  Vec3 minP{{0, 0, 0}}, minQ{{0, 0, 0}};
  int shown_disjoint = 0;
  double mindd = sqnorm(sub(T1[0], T2[0])) + 1;

  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) {
      seg_points(T1[i], Sv[i], T2[j], Tv[j], VEC, P, Q);
      V = sub(Q, P);
      double dd = dot(V, V);
      if (dd <= mindd) {
        minP = P;
        minQ = Q;
        mindd = dd;

        Z = sub(T1[(i + 2) % 3], P);
        double a = dot(Z, VEC);
        Z = sub(T2[(j + 2) % 3], Q);
        double b = dot(Z, VEC);

        if ((a <= 0) && (b >= 0)) return std::sqrt(dd);

        double p = dot(V, VEC);
        if (a < 0) a = 0;
        if (b > 0) b = 0;
        if ((p - a + b) > 0) shown_disjoint = 1;
      }
    }
  }

  Vec3 Sn = cross(Sv[0], Sv[1]);
  double Snl = dot(Sn, Sn);
  if (Snl > 1e-15) {
    double Tp[3];
    V = sub(T1[0], T2[0]); Tp[0] = dot(V, Sn);
    V = sub(T1[0], T2[1]); Tp[1] = dot(V, Sn);
    V = sub(T1[0], T2[2]); Tp[2] = dot(V, Sn);

    int point = -1;
    if ((Tp[0] > 0) && (Tp[1] > 0) && (Tp[2] > 0)) {
      point = (Tp[0] < Tp[1]) ? 0 : 1;
      if (Tp[2] < Tp[point]) point = 2;
    } else if ((Tp[0] < 0) && (Tp[1] < 0) && (Tp[2] < 0)) {
      point = (Tp[0] > Tp[1]) ? 0 : 1;
      if (Tp[2] > Tp[point]) point = 2;
    }
    if (point >= 0) {
      shown_disjoint = 1;
      V = sub(T2[point], T1[0]);
      Z = cross(Sn, Sv[0]);
      if (dot(V, Z) > 0) {
        V = sub(T2[point], T1[1]);
        Z = cross(Sn, Sv[1]);
        if (dot(V, Z) > 0) {
          V = sub(T2[point], T1[2]);
          Z = cross(Sn, Sv[2]);
          if (dot(V, Z) > 0) {
            P = add(T2[point], scale(Sn, Tp[point] / Snl));
            Q = T2[point];
            return norm(sub(P, Q));
          }
        }
      }
    }
  }

  Vec3 Tn = cross(Tv[0], Tv[1]);
  double Tnl = dot(Tn, Tn);
  if (Tnl > 1e-15) {
    double Sp[3];
    V = sub(T2[0], T1[0]); Sp[0] = dot(V, Tn);
    V = sub(T2[0], T1[1]); Sp[1] = dot(V, Tn);
    V = sub(T2[0], T1[2]); Sp[2] = dot(V, Tn);

    int point = -1;
    if ((Sp[0] > 0) && (Sp[1] > 0) && (Sp[2] > 0)) {
      point = (Sp[0] < Sp[1]) ? 0 : 1;
      if (Sp[2] < Sp[point]) point = 2;
    } else if ((Sp[0] < 0) && (Sp[1] < 0) && (Sp[2] < 0)) {
      point = (Sp[0] > Sp[1]) ? 0 : 1;
      if (Sp[2] > Sp[point]) point = 2;
    }
    if (point >= 0) {
      shown_disjoint = 1;
      V = sub(T1[point], T2[0]);
      Z = cross(Tn, Tv[0]);
      if (dot(V, Z) > 0) {
        V = sub(T1[point], T2[1]);
        Z = cross(Tn, Tv[1]);
        if (dot(V, Z) > 0) {
          V = sub(T1[point], T2[2]);
          Z = cross(Tn, Tv[2]);
          if (dot(V, Z) > 0) {
            P = T1[point];
            Q = add(T1[point], scale(Tn, Sp[point] / Tnl));
            return norm(sub(P, Q));
          }
        }
      }
    }
  }

  if (shown_disjoint) {
    P = minP;
    Q = minQ;
    return std::sqrt(mindd);
  }
  return 0;
}

// ---------------------------------------------------------------------------------------
// sphereTriangleIntersect and its helpers segmentSqrDistance / projectInTriangle --
// narrowphase/detail/primitive_shape_algorithm/sphere_triangle-inl.h:85-244
// ---------------------------------------------------------------------------------------
namespace {
double segment_sqr_distance(const Vec3& from, const Vec3& to, const Vec3& p, Vec3& nearest) {  // :85-111
  Vec3 diff = sub(p, from);
  const Vec3 v = sub(to, from);
  double t = dot(v, diff);
  if (t > 0) {
    const double dotVV = dot(v, v);
    if (t < dotVV) {
      t /= dotVV;
      diff = sub(diff, scale(v, t));
    } else {
      t = 1;
      diff = sub(diff, v);
    }
  } else {
    t = 0;
  }
  nearest = add(from, scale(v, t));
  return dot(diff, diff);
}

bool project_in_triangle(const Vec3& p1, const Vec3& p2, const Vec3& p3, const Vec3& normal, const Vec3& p) {  // :115-139
  const Vec3 edge1 = sub(p2, p1), edge2 = sub(p3, p2), edge3 = sub(p1, p3);
  const Vec3 p1_to_p = sub(p, p1), p2_to_p = sub(p, p2), p3_to_p = sub(p, p3);
  const double r1 = dot(cross(edge1, normal), p1_to_p);
  const double r2 = dot(cross(edge2, normal), p2_to_p);
  const double r3 = dot(cross(edge3, normal), p3_to_p);
  return (r1 > 0 && r2 > 0 && r3 > 0) || (r1 <= 0 && r2 <= 0 && r3 <= 0);
}
}  // namespace

bool sphere_tri_intersect(const Vec3& center, double radius, const Vec3& P1, const Vec3& P2, const Vec3& P3,
                          Vec3* contact_point_out, double* penetration_depth, Vec3* normal_out) {  // :143-244
  Vec3 normal = cross(sub(P2, P1), sub(P3, P1));
  {  // Eigen normalize(): v /= sqrt(v.squaredNorm()), a true division per component
    const double nn = norm(normal);
    normal = Vec3{{normal[0] / nn, normal[1] / nn, normal[2] / nn}};
  }
  const double radius_with_threshold = radius + std::numeric_limits<double>::epsilon();
  const Vec3 p1_to_center = sub(center, P1);
  double distance_from_plane = dot(p1_to_center, normal);
  if (distance_from_plane < 0) {
    distance_from_plane *= -1;
    normal = scale(normal, -1.0);
  }
  const bool is_inside_contact_plane = distance_from_plane < radius_with_threshold;
  bool has_contact = false;
  Vec3 contact_point{{0, 0, 0}};
  if (is_inside_contact_plane) {
    if (project_in_triangle(P1, P2, P3, normal, center)) {
      has_contact = true;
      contact_point = sub(center, scale(normal, distance_from_plane));
    } else {
      const double contact_capsule_radius_sqr = radius_with_threshold * radius_with_threshold;
      Vec3 nearest_on_edge;
      double distance_sqr = segment_sqr_distance(P1, P2, center, nearest_on_edge);
      if (distance_sqr < contact_capsule_radius_sqr) {
        has_contact = true;
        contact_point = nearest_on_edge;
      }
      distance_sqr = segment_sqr_distance(P2, P3, center, nearest_on_edge);
      if (distance_sqr < contact_capsule_radius_sqr) {
        has_contact = true;
        contact_point = nearest_on_edge;
      }
      distance_sqr = segment_sqr_distance(P3, P1, center, nearest_on_edge);
      if (distance_sqr < contact_capsule_radius_sqr) {
        has_contact = true;
        contact_point = nearest_on_edge;
      }
    }
  }
  if (has_contact) {
    const Vec3 contact_to_center = sub(contact_point, center);
    const double distance_sqr = sqnorm(contact_to_center);
    if (distance_sqr < radius_with_threshold * radius_with_threshold) {
      if (distance_sqr > 0) {
        const double distance = std::sqrt(distance_sqr);
        if (normal_out) *normal_out = Vec3{{contact_to_center[0] / distance, contact_to_center[1] / distance, contact_to_center[2] / distance}};  // normalized()
        if (contact_point_out) *contact_point_out = contact_point;
        if (penetration_depth) *penetration_depth = -(radius - distance);
      } else {
        if (normal_out) *normal_out = scale(normal, -1.0);
        if (contact_point_out) *contact_point_out = contact_point;
        if (penetration_depth) *penetration_depth = -radius;
      }
      return true;
    }
  }
  return false;
}

// ---------------------------------------------------------------------------------------
// Project<S>::projectLine / projectTriangle -- include/fcl/math/detail/project-inl.h:54-123,
// and the nearest-point overload of sphereTriangleDistance (sphere_triangle-inl.h:469-508), which is
// what the mesh <-> sphere distance leaf always reaches (it passes both nearest-point pointers,
// mesh_shape_distance_traversal_node-inl.h:186-190).
// ---------------------------------------------------------------------------------------
namespace {
struct ProjectResult {  // project.h:53-66, ctor project-inl.h:311-315
  double parameterization[4] = {0.0, 0.0, 0.0, 0.0};
  double sqr_distance = -1;
  unsigned encode = 0;
};

ProjectResult project_line(const Vec3& a, const Vec3& b, const Vec3& p) {  // :54-73
  ProjectResult res;
  const Vec3 d = sub(b, a);
  const double l = sqnorm(d);
  if (l > 0) {
    const double t = dot(sub(p, a), d);
    res.parameterization[1] = (t >= l) ? 1 : ((t <= 0) ? 0 : (t / l));
    res.parameterization[0] = 1 - res.parameterization[1];
    if (t >= l) {
      res.sqr_distance = sqnorm(sub(p, b));
      res.encode = 2;
    } else if (t <= 0) {
      res.sqr_distance = sqnorm(sub(p, a));
      res.encode = 1;
    } else {
      res.sqr_distance = sqnorm(sub(add(a, scale(d, res.parameterization[1])), p));
      res.encode = 3;
    }
  }
  return res;
}

ProjectResult project_triangle(const Vec3& a, const Vec3& b, const Vec3& c, const Vec3& p) {  // :77-123
  ProjectResult res;
  static const int nexti[3] = {1, 2, 0};
  const Vec3* vt[] = {&a, &b, &c};
  const Vec3 dl[] = {sub(a, b), sub(b, c), sub(c, a)};
  const Vec3 n = cross(dl[0], dl[1]);
  const double l = sqnorm(n);
  if (l > 0) {
    double mindist = -1;
    for (int i = 0; i < 3; ++i) {
      if (dot(sub(*vt[i], p), cross(dl[i], n)) > 0) {  // outside this edge: the optimum can only be on the edge
        const int j = nexti[i];
        const ProjectResult res_line = project_line(*vt[i], *vt[j], p);
        if (mindist < 0 || res_line.sqr_distance < mindist) {
          mindist = res_line.sqr_distance;
          res.encode = ((res_line.encode & 1) ? 1u << i : 0u) + ((res_line.encode & 2) ? 1u << j : 0u);
          res.parameterization[i] = res_line.parameterization[0];
          res.parameterization[j] = res_line.parameterization[1];
          res.parameterization[nexti[j]] = 0;
        }
      }
    }
    if (mindist < 0) {  // the projection falls inside the triangle
      const double d = dot(sub(a, p), n);
      const double s = std::sqrt(l);
      const Vec3 p_to_project = scale(n, d / l);
      mindist = sqnorm(p_to_project);
      res.encode = 7;
      res.parameterization[0] = norm(cross(dl[1], sub(sub(b, p), p_to_project))) / s;
      res.parameterization[1] = norm(cross(dl[2], sub(sub(c, p), p_to_project))) / s;
      res.parameterization[2] = 1 - res.parameterization[0] - res.parameterization[1];
    }
    res.sqr_distance = mindist;
  }
  return res;
}
}  // namespace

// sphereTriangleDistance(sp, tf, P1, P2, P3, dist, p1, p2), :469-496: centre o = tf.translation(), triangle in
// the same (world) frame.  Returns false -- and, in the reference, leaves *dist and the points UNWRITTEN --
// when the centre is within the radius of the triangle (or the triangle has zero area: sqr_distance = -1).
// on_sphere_world = o - dir * radius (before the reference maps it into the sphere's frame), on_triangle = project_p.
bool sphere_tri_distance(const Vec3& o, double radius, const Vec3& P1, const Vec3& P2, const Vec3& P3, double* dist,
                         Vec3* on_sphere_world, Vec3* on_triangle) {
  const ProjectResult result = project_triangle(P1, P2, P3, o);
  if (result.sqr_distance > radius * radius) {
    if (dist) *dist = std::sqrt(result.sqr_distance) - radius;
    const Vec3 project_p = add(add(scale(P1, result.parameterization[0]), scale(P2, result.parameterization[1])),
                               scale(P3, result.parameterization[2]));
    Vec3 dir = sub(o, project_p);
    {  // Eigen normalize()
      const double nn = norm(dir);
      dir = Vec3{{dir[0] / nn, dir[1] / nn, dir[2] / nn}};
    }
    if (on_sphere_world) *on_sphere_world = sub(o, scale(dir, radius));
    if (on_triangle) *on_triangle = project_p;
    return true;
  }
  return false;
}

// -----------------------------------------------------------------------------
// Halfspace / Plane vs triangle -- closed form, so these pairs can be pinned without libccd.
// -----------------------------------------------------------------------------
// Halfspace(n, d) / Plane(n, d) constructors -> unitNormalTest (geometry/shape/halfspace-inl.h:144-160, plane-inl.h:144-160)
PlaneShape make_plane(const Vec3& n_in, double d_in) {
  PlaneShape s;
  const double l = norm(n_in);
  if (l > 0) {
    const double inv_l = 1.0 / l;
    s.n = scale(n_in, inv_l);
    s.d = d_in * inv_l;
  } else {
    s.n = Vec3{{1, 0, 0}};
    s.d = 0;
  }
  return s;
}

// transform(): n' = tf.linear() * n, d' = d + n'.dot(tf.translation())
PlaneShape transform_plane(const PlaneShape& a, const Pose& tf) {
  PlaneShape r;
  r.n = mul(tf.R, a.n);
  r.d = a.d + dot(r.n, tf.t);
  return r;
}

static inline double signed_distance(const PlaneShape& s, const Vec3& p) { return dot(s.n, p) - s.d; }  // halfspace-inl.h:83-86

// halfspaceTriangleIntersect, narrowphase/detail/primitive_shape_algorithm/halfspace-inl.h:587-621
bool halfspace_tri_intersect(const PlaneShape& s1, const Pose& tf1, const Vec3& P1, const Vec3& P2, const Vec3& P3, const Pose& tf2,
                             Vec3* contact_point, double* penetration_depth, Vec3* normal) {
  const PlaneShape new_s1 = transform_plane(s1, tf1);
  Vec3 v = add(mul(tf2.R, P1), tf2.t);
  double depth = signed_distance(new_s1, v);
  Vec3 p = add(mul(tf2.R, P2), tf2.t);
  double d = signed_distance(new_s1, p);
  if (d < depth) {
    depth = d;
    v = p;
  }
  p = add(mul(tf2.R, P3), tf2.t);
  d = signed_distance(new_s1, p);
  if (d < depth) {
    depth = d;
    v = p;
  }
  if (depth <= 0) {
    if (penetration_depth) *penetration_depth = -depth;
    if (normal) *normal = new_s1.n;
    if (contact_point) *contact_point = sub(v, scale(new_s1.n, 0.5 * depth));
    return true;
  }
  return false;
}

// planeTriangleIntersect, narrowphase/detail/primitive_shape_algorithm/plane-inl.h:683-759
bool plane_tri_intersect(const PlaneShape& s1, const Pose& tf1, const Vec3& P1, const Vec3& P2, const Vec3& P3, const Pose& tf2,
                         Vec3* contact_point, double* penetration_depth, Vec3* normal) {
  const PlaneShape new_s1 = transform_plane(s1, tf1);
  Vec3 c[3];
  c[0] = add(mul(tf2.R, P1), tf2.t);
  c[1] = add(mul(tf2.R, P2), tf2.t);
  c[2] = add(mul(tf2.R, P3), tf2.t);
  double d[3];
  for (int i = 0; i < 3; ++i) d[i] = signed_distance(new_s1, c[i]);
  if ((d[0] >= 0 && d[1] >= 0 && d[2] >= 0) || (d[0] <= 0 && d[1] <= 0 && d[2] <= 0)) return false;
  bool positive[3];
  for (int i = 0; i < 3; ++i) positive[i] = (d[i] > 0);
  int n_positive = 0;
  double d_positive = 0, d_negative = 0;
  for (int i = 0; i < 3; ++i) {
    if (positive[i]) {
      n_positive++;
      if (d_positive <= d[i]) d_positive = d[i];
    } else {
      if (d_negative <= -d[i]) d_negative = -d[i];
    }
  }
  if (penetration_depth) *penetration_depth = std::min(d_positive, d_negative);
  if (normal) *normal = (d_positive > d_negative) ? new_s1.n : Vec3{{-new_s1.n[0], -new_s1.n[1], -new_s1.n[2]}};
  if (contact_point) {
    Vec3 p[2] = {Vec3{{0, 0, 0}}, Vec3{{0, 0, 0}}};
    Vec3 q{{0, 0, 0}};
    double p_d[2] = {0, 0};
    double q_d = 0;
    if (n_positive == 2) {
      for (int i = 0, j = 0; i < 3; ++i) {
        if (positive[i]) { p[j] = c[i]; p_d[j] = d[i]; j++; }
        else { q = c[i]; q_d = d[i]; }
      }
      // t = (-p * q_d + q * p_d) / (-q_d + p_d): the unary minus applies to the vector, then the scalar product
      Vec3 t1, t2;
      for (int k = 0; k < 3; ++k) {
        t1[k] = ((-p[0][k]) * q_d + q[k] * p_d[0]) / (-q_d + p_d[0]);
        t2[k] = ((-p[1][k]) * q_d + q[k] * p_d[1]) / (-q_d + p_d[1]);
      }
      *contact_point = scale(add(t1, t2), 0.5);
    } else {
      for (int i = 0, j = 0; i < 3; ++i) {
        if (!positive[i]) { p[j] = c[i]; p_d[j] = d[i]; j++; }
        else { q = c[i]; q_d = d[i]; }
      }
      Vec3 t1, t2;
      for (int k = 0; k < 3; ++k) {
        t1[k] = (p[0][k] * q_d - q[k] * p_d[0]) / (q_d - p_d[0]);
        t2[k] = (p[1][k] * q_d - q[k] * p_d[1]) / (q_d - p_d[1]);
      }
      *contact_point = scale(add(t1, t2), 0.5);
    }
  }
  return true;
}

}  // namespace oracle
