// fcl_oracle_bvh.cpp — CPU ORACLE (test infrastructure, NOT product code).
// OBJ loading, BVHModel<OBBRSS> construction and the collide / distance
// traversals.  See fcl_oracle.hpp for the parity statement.
// Citations are relative to /root/reference/.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <queue>

#include "fcl_oracle.hpp"
#include "fcl_oracle_vec.hpp"

namespace oracle {

// -----------------------------------------------------------------------------
// OBJ loader with the semantics of test/test_fcl_utility.h:194-280: the first
// whitespace token selects the record; 'v' (not vn/vt) adds a vertex, 'f' adds a
// triangle fan with 1-based indices ("a/b/c" -> atoi reads a); anything else
// (including the "6540 2180" first line of env.obj) is ignored.
// -----------------------------------------------------------------------------
bool load_obj(const std::string& path, std::vector<Vec3>& pts, std::vector<Tri>& tris) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return false;
  char line[2000];
  while (std::fgets(line, sizeof line, f)) {
    char* tok = std::strtok(line, "\r\n\t ");
    if (!tok || tok[0] == '#' || tok[0] == 0) continue;
    if (tok[0] == 'v') {
      if (tok[1] == 'n' || tok[1] == 't') continue;
      double c[3] = {0, 0, 0};
      for (int k = 0; k < 3; ++k) {
        char* s = std::strtok(nullptr, "\t ");
        c[k] = s ? std::atof(s) : 0.0;
      }
      pts.push_back(Vec3{{c[0], c[1], c[2]}});
    } else if (tok[0] == 'f') {
      int idx[30];
      int n = 0;
      char* s;
      while (n < 30 && (s = std::strtok(nullptr, "\t \r\n")) != nullptr)
        if (std::strlen(s)) idx[n++] = std::atoi(s) - 1;
      // fan: (0, t+1, t+2).  (For plain "f a b c" records the reference emits the
      // same triangle; its no-normal branch only differs for polygons, :250-255.)
      for (int t = 0; t < n - 2; ++t) tris.push_back(Tri{{idx[0], idx[t + 1], idx[t + 2]}});
    }
  }
  std::fclose(f);
  return true;
}

// -----------------------------------------------------------------------------
// Fitting — include/fcl/geometry/bvh/detail/BV_fitter-inl.h:449-477
// -----------------------------------------------------------------------------
namespace {

struct Builder {
  Model& m;
  SplitMethod split;
  std::vector<unsigned> prim;
  int num_bvs = 0;

  // getCovariance — include/fcl/math/geometry-inl.h:1335-1425 (triangle branch)
  void covariance(const unsigned* idx, int n, double M[3][3]) const {
    double S1[3] = {0, 0, 0};
    double S2[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int i = 0; i < n; ++i) {
      const Tri& t = m.tris[idx[i]];
      const Vec3& p1 = m.verts[t.v[0]];
      const Vec3& p2 = m.verts[t.v[1]];
      const Vec3& p3 = m.verts[t.v[2]];
      for (int k = 0; k < 3; ++k) S1[k] += ((p1[k] + p2[k]) + p3[k]);
      S2[0][0] += (p1[0] * p1[0] + p2[0] * p2[0] + p3[0] * p3[0]);
      S2[1][1] += (p1[1] * p1[1] + p2[1] * p2[1] + p3[1] * p3[1]);
      S2[2][2] += (p1[2] * p1[2] + p2[2] * p2[2] + p3[2] * p3[2]);
      S2[0][1] += (p1[0] * p1[1] + p2[0] * p2[1] + p3[0] * p3[1]);
      S2[0][2] += (p1[0] * p1[2] + p2[0] * p2[2] + p3[0] * p3[2]);
      S2[1][2] += (p1[1] * p1[2] + p2[1] * p2[2] + p3[1] * p3[2]);
    }
    int n_points = 3 * n;
    M[0][0] = S2[0][0] - S1[0] * S1[0] / n_points;
    M[1][1] = S2[1][1] - S1[1] * S1[1] / n_points;
    M[2][2] = S2[2][2] - S1[2] * S1[2] / n_points;
    M[0][1] = S2[0][1] - S1[0] * S1[1] / n_points;
    M[1][2] = S2[1][2] - S1[1] * S1[2] / n_points;
    M[0][2] = S2[0][2] - S1[0] * S1[2] / n_points;
    M[1][0] = M[0][1];
    M[2][0] = M[0][2];
    M[2][1] = M[1][2];
  }

  // eigen_old — include/fcl/math/geometry-inl.h:477-558 (cyclic Jacobi, <=50 sweeps).
  // On return evec[k] is the k-th eigenvector (the reference's vout.col(k) written
  // from v[k][*] and read back with eigenV.row(), :503-505 and :590-591).
  static void jacobi(const double Min[3][3], double dout[3], double v[3][3]) {
    double R[3][3];
    std::memcpy(R, Min, sizeof R);
    const int n = 3;
    double b[3], z[3], d[3];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) v[i][j] = (i == j) ? 1.0 : 0.0;
    for (int ip = 0; ip < n; ++ip) {
      b[ip] = d[ip] = R[ip][ip];
      z[ip] = 0;
    }
    for (int i = 0; i < 50; ++i) {
      double sm = 0;
      for (int ip = 0; ip < n; ++ip)
        for (int iq = ip + 1; iq < n; ++iq) sm += std::abs(R[ip][iq]);
      if (sm == 0.0) {
        dout[0] = d[0]; dout[1] = d[1]; dout[2] = d[2];
        return;
      }
      double tresh = (i < 3) ? 0.2 * sm / (n * n) : 0.0;
      for (int ip = 0; ip < n; ++ip) {
        for (int iq = ip + 1; iq < n; ++iq) {
          double g = 100.0 * std::abs(R[ip][iq]);
          if (i > 3 && std::abs(d[ip]) + g == std::abs(d[ip]) && std::abs(d[iq]) + g == std::abs(d[iq]))
            R[ip][iq] = 0.0;
          else if (std::abs(R[ip][iq]) > tresh) {
            double h = d[iq] - d[ip], t;
            if (std::abs(h) + g == std::abs(h)) t = (R[ip][iq]) / h;
            else {
              double theta = 0.5 * h / (R[ip][iq]);
              t = 1.0 / (std::abs(theta) + std::sqrt(1.0 + theta * theta));
              if (theta < 0.0) t = -t;
            }
            double c = 1.0 / std::sqrt(1 + t * t);
            double s = t * c;
            double tau = s / (1.0 + c);
            h = t * R[ip][iq];
            z[ip] -= h; z[iq] += h; d[ip] -= h; d[iq] += h;
            R[ip][iq] = 0.0;
            auto rot = [&](double& x, double& y) {
              double gg = x, hh = y;
              x = gg - s * (hh + gg * tau);
              y = hh + s * (gg - hh * tau);
            };
            for (int j = 0; j < ip; ++j) rot(R[j][ip], R[j][iq]);
            for (int j = ip + 1; j < iq; ++j) rot(R[ip][j], R[j][iq]);
            for (int j = iq + 1; j < n; ++j) rot(R[ip][j], R[iq][j]);
            for (int j = 0; j < n; ++j) rot(v[j][ip], v[j][iq]);
          }
        }
      }
      for (int ip = 0; ip < n; ++ip) {
        b[ip] += z[ip];
        d[ip] = b[ip];
        z[ip] = 0.0;
      }
    }
    std::fprintf(stderr, "oracle: too many iterations in Jacobi transform.\n");
    dout[0] = d[0]; dout[1] = d[1]; dout[2] = d[2];
  }

  // axisFromEigen — include/fcl/math/geometry-inl.h:563-597
  static void axis_from_eigen(const double v[3][3], const double s[3], Mat3& axis) {
    int mn, mid, mx;
    if (s[0] > s[1]) { mx = 0; mn = 1; } else { mn = 0; mx = 1; }
    if (s[2] < s[mn]) { mid = mn; mn = 2; }
    else if (s[2] > s[mx]) { mid = mx; mx = 2; }
    else { mid = 2; }
    (void)mn;
    for (int r = 0; r < 3; ++r) {
      axis.m[r][0] = v[r][mx];
      axis.m[r][1] = v[r][mid];
    }
    Vec3 c2 = cross(col(axis, 0), col(axis, 1));
    for (int r = 0; r < 3; ++r) axis.m[r][2] = c2[r];
  }

  // getExtentAndCenter_mesh — include/fcl/math/geometry-inl.h:294-362
  void extent_and_center(const unsigned* idx, int n, const Mat3& axis, Vec3& center, Vec3& extent) const {
    const double real_max = std::numeric_limits<double>::max();
    double mn[3] = {real_max, real_max, real_max}, mx[3] = {-real_max, -real_max, -real_max};
    for (int i = 0; i < n; ++i) {
      const Tri& t = m.tris[idx[i]];
      for (int j = 0; j < 3; ++j) {
        const Vec3& p = m.verts[t.v[j]];
        for (int k = 0; k < 3; ++k) {
          double proj = dot(col(axis, k), p);
          if (proj > mx[k]) mx[k] = proj;
          if (proj < mn[k]) mn[k] = proj;
        }
      }
    }
    Vec3 o{{(mx[0] + mn[0]) / 2, (mx[1] + mn[1]) / 2, (mx[2] + mn[2]) / 2}};
    center = mul(axis, o);
    extent = Vec3{{(mx[0] - mn[0]) / 2, (mx[1] - mn[1]) / 2, (mx[2] - mn[2]) / 2}};
  }

  // getRadiusAndOriginAndRectangleSize — include/fcl/math/geometry-inl.h:709-988
  void rss_fit(const unsigned* idx, int n, const Mat3& axis, Vec3& origin, double l[2], double& r) const {
    const int size_P = 3 * n;
    std::vector<Vec3> P(size_P);
    int id = 0;
    for (int i = 0; i < n; ++i) {
      const Tri& t = m.tris[idx[i]];
      for (int j = 0; j < 3; ++j) {
        const Vec3& p = m.verts[t.v[j]];
        P[id][0] = dot(col(axis, 0), p);
        P[id][1] = dot(col(axis, 1), p);
        P[id][2] = dot(col(axis, 2), p);
        id++;
      }
    }
    rss_from_projections(P, axis, origin, l, r);
  }

  // the part of getRadiusAndOriginAndRectangleSize after the projections P have been gathered (:782-988); the point
  // branch (:757-779) builds P from plain points and continues the same way
  static void rss_from_projections(const std::vector<Vec3>& P, const Mat3& axis, Vec3& origin, double l[2], double& r) {
    const int size_P = (int)P.size();
    double minx, maxx, miny, maxy, minz, maxz, cz, radsqr;
    minz = maxz = P[0][2];
    for (int i = 1; i < size_P; ++i) {
      double zv = P[i][2];
      if (zv < minz) minz = zv;
      else if (zv > maxz) maxz = zv;
    }
    r = 0.5 * (maxz - minz);
    radsqr = r * r;
    cz = 0.5 * (maxz + minz);

    auto cap = [&](int i) {  // sqrt(max(radsqr - dz*dz, 0)) for point i
      double dz = P[i][2] - cz;
      return std::sqrt(std::max<double>(radsqr - dz * dz, 0));
    };
    auto extreme = [&](int c, int& minindex, int& maxindex) {
      minindex = maxindex = 0;
      double mintmp = P[0][c], maxtmp = P[0][c];
      for (int i = 1; i < size_P; ++i) {
        double val = P[i][c];
        if (val < mintmp) { minindex = i; mintmp = val; }
        else if (val > maxtmp) { maxindex = i; maxtmp = val; }
      }
    };
    int minindex, maxindex;
    extreme(0, minindex, maxindex);
    minx = P[minindex][0] + cap(minindex);
    maxx = P[maxindex][0] - cap(maxindex);
    for (int i = 0; i < size_P; ++i)
      if (P[i][0] < minx) { double x = P[i][0] + cap(i); if (x < minx) minx = x; }
    for (int i = 0; i < size_P; ++i)
      if (P[i][0] > maxx) { double x = P[i][0] - cap(i); if (x > maxx) maxx = x; }

    extreme(1, minindex, maxindex);
    miny = P[minindex][1] + cap(minindex);
    maxy = P[maxindex][1] - cap(maxindex);
    for (int i = 0; i < size_P; ++i)
      if (P[i][1] < miny) { double y = P[i][1] + cap(i); if (y < miny) miny = y; }
    for (int i = 0; i < size_P; ++i)
      if (P[i][1] > maxy) { double y = P[i][1] - cap(i); if (y > maxy) maxy = y; }

    // corner growth (:883-965)
    double dx, dy, u, t;
    const double a = std::sqrt(0.5);
    for (int i = 0; i < size_P; ++i) {
      if (P[i][0] > maxx) {
        if (P[i][1] > maxy) {
          dx = P[i][0] - maxx; dy = P[i][1] - maxy;
          u = dx * a + dy * a;
          t = (a * u - dx) * (a * u - dx) + (a * u - dy) * (a * u - dy) + (cz - P[i][2]) * (cz - P[i][2]);
          u = u - std::sqrt(std::max<double>(radsqr - t, 0));
          if (u > 0) { maxx += u * a; maxy += u * a; }
        } else if (P[i][1] < miny) {
          dx = P[i][0] - maxx; dy = P[i][1] - miny;
          u = dx * a - dy * a;
          t = (a * u - dx) * (a * u - dx) + (-a * u - dy) * (-a * u - dy) + (cz - P[i][2]) * (cz - P[i][2]);
          u = u - std::sqrt(std::max<double>(radsqr - t, 0));
          if (u > 0) { maxx += u * a; miny -= u * a; }
        }
      } else if (P[i][0] < minx) {
        if (P[i][1] > maxy) {
          dx = P[i][0] - minx; dy = P[i][1] - maxy;
          u = dy * a - dx * a;
          t = (-a * u - dx) * (-a * u - dx) + (a * u - dy) * (a * u - dy) + (cz - P[i][2]) * (cz - P[i][2]);
          u = u - std::sqrt(std::max<double>(radsqr - t, 0));
          if (u > 0) { minx -= u * a; maxy += u * a; }
        } else if (P[i][1] < miny) {
          dx = P[i][0] - minx; dy = P[i][1] - miny;
          u = -dx * a - dy * a;
          t = (-a * u - dx) * (-a * u - dx) + (-a * u - dy) * (-a * u - dy) + (cz - P[i][2]) * (cz - P[i][2]);
          u = u - std::sqrt(std::max<double>(radsqr - t, 0));
          if (u > 0) { minx -= u * a; miny -= u * a; }
        }
      }
    }
    for (int k = 0; k < 3; ++k) origin[k] = (axis.m[k][0] * minx + axis.m[k][1] * miny) + axis.m[k][2] * cz;
    l[0] = maxx - minx; if (l[0] < 0) l[0] = 0;
    l[1] = maxy - miny; if (l[1] < 0) l[1] = 0;
  }

  void fit(const unsigned* idx, int n, Node& nd) const {
    double M[3][3], s[3], E[3][3];
    covariance(idx, n, M);
    jacobi(M, s, E);
    axis_from_eigen(E, s, nd.axis);
    nd.rss_axis = nd.axis;  // bv.rss.axis = bv.obb.axis (BV_fitter-inl.h:464)
    extent_and_center(idx, n, nd.axis, nd.obb_To, nd.obb_ext);
    rss_fit(idx, n, nd.axis, nd.rss_To, nd.rss_l, nd.rss_r);
  }

  // split rule — include/fcl/geometry/bvh/detail/BV_splitter-inl.h:422-451 (OBBRSS
  // specialisations), :540-547 (split vector = obb.axis.col(0)), :550-555,
  // :558-599 (mean), :603-657 (median); apply(): :492-500.
  double split_value(const Node& nd, const unsigned* idx, int n, const Vec3& sv) const {
    if (split == SPLIT_BV_CENTER) return nd.obb_To[0];
    if (split == SPLIT_MEAN) {
      double c[3] = {0.0, 0.0, 0.0};
      for (int i = 0; i < n; ++i) {
        const Tri& t = m.tris[idx[i]];
        const Vec3& p1 = m.verts[t.v[0]];
        const Vec3& p2 = m.verts[t.v[1]];
        const Vec3& p3 = m.verts[t.v[2]];
        c[0] += (p1[0] + p2[0] + p3[0]);
        c[1] += (p1[1] + p2[1] + p3[1]);
        c[2] += (p1[2] + p2[2] + p3[2]);
      }
      return (c[0] * sv[0] + c[1] * sv[1] + c[2] * sv[2]) / (3 * n);
    }
    std::vector<double> proj(n);
    for (int i = 0; i < n; ++i) {
      const Tri& t = m.tris[idx[i]];
      const Vec3& p1 = m.verts[t.v[0]];
      const Vec3& p2 = m.verts[t.v[1]];
      const Vec3& p3 = m.verts[t.v[2]];
      Vec3 c3{{p1[0] + p2[0] + p3[0], p1[1] + p2[1] + p3[1], p1[2] + p2[2] + p3[2]}};
      proj[i] = dot(c3, sv) / 3;
    }
    std::sort(proj.begin(), proj.end());
    if (n % 2 == 1) return proj[(n - 1) / 2];
    return (proj[n / 2] + proj[n / 2 - 1]) / 2;
  }

  // recursiveBuildTree — include/fcl/geometry/bvh/BVH_model-inl.h:868-938
  void recurse(int bv_id, int first, int n) {
    unsigned* cur = prim.data() + first;
    Node nd;
    fit(cur, n, nd);
    Vec3 sv = col(nd.axis, 0);
    double sval = split_value(nd, cur, n, sv);
    nd.first_primitive = first;
    nd.num_primitives = n;
    if (n == 1) {
      nd.first_child = -((int)cur[0] + 1);
      m.nodes[bv_id] = nd;
      return;
    }
    nd.first_child = num_bvs;
    num_bvs += 2;
    m.nodes[bv_id] = nd;

    int c1 = 0;
    for (int i = 0; i < n; ++i) {
      const Tri& t = m.tris[cur[i]];
      const Vec3& p1 = m.verts[t.v[0]];
      const Vec3& p2 = m.verts[t.v[1]];
      const Vec3& p3 = m.verts[t.v[2]];
      Vec3 p{{((p1[0] + p2[0]) + p3[0]) / 3.0, ((p1[1] + p2[1]) + p3[1]) / 3.0, ((p1[2] + p2[2]) + p3[2]) / 3.0}};
      if (dot(sv, p) > sval) {
        // right side: stays
      } else {
        std::swap(cur[i], cur[c1]);
        c1++;
      }
    }
    if ((c1 == 0) || (c1 == n)) c1 = n / 2;
    recurse(nd.first_child, first, c1);
    recurse(nd.first_child + 1, first + c1, n - c1);
  }
};


// ---------------------------------------------------------------------------------------
// computeBV<OBBRSS<S>>(Sphere, tf) -- geometry/shape/utility-inl.h generic ComputeBVImpl:
// fit(getBoundVertices(tf)) with n = 12 > 3 -> fitn (BV_fitter-inl.h): getCovariance (point branch,
// math/geometry-inl.h:1383-1409), eigen_old, axisFromEigen, getExtentAndCenter_pointcloud (:229-292).
// Only the OBB part is needed by the collide traversal.
// ---------------------------------------------------------------------------------------
void sphere_obb_impl(double radius, const Pose& tf, Node& bv) {
  const double m = (1 + std::sqrt(5.0)) / 2.0;
  const double edge = radius * 6 / (std::sqrt(27.0) + std::sqrt(15.0));
  const double a = edge, b = m * edge;
  const double L[12][3] = {{0, a, b}, {0, -a, b}, {0, a, -b}, {0, -a, -b}, {a, b, 0}, {-a, b, 0},
                           {a, -b, 0}, {-a, -b, 0}, {b, 0, a}, {b, 0, -a}, {-b, 0, a}, {-b, 0, -a}};
  Vec3 ps[12];
  for (int i = 0; i < 12; ++i) ps[i] = add(mul(tf.R, Vec3{{L[i][0], L[i][1], L[i][2]}}), tf.t);
  double S1[3] = {0, 0, 0}, S2[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int i = 0; i < 12; ++i) {
    const Vec3& p = ps[i];
    for (int k = 0; k < 3; ++k) S1[k] += p[k];
    S2[0][0] += (p[0] * p[0]);
    S2[1][1] += (p[1] * p[1]);
    S2[2][2] += (p[2] * p[2]);
    S2[0][1] += (p[0] * p[1]);
    S2[0][2] += (p[0] * p[2]);
    S2[1][2] += (p[1] * p[2]);
  }
  const int n_points = 12;
  double M[3][3];
  M[0][0] = S2[0][0] - S1[0] * S1[0] / n_points;
  M[1][1] = S2[1][1] - S1[1] * S1[1] / n_points;
  M[2][2] = S2[2][2] - S1[2] * S1[2] / n_points;
  M[0][1] = S2[0][1] - S1[0] * S1[1] / n_points;
  M[1][2] = S2[1][2] - S1[1] * S1[2] / n_points;
  M[0][2] = S2[0][2] - S1[0] * S1[2] / n_points;
  M[1][0] = M[0][1];
  M[2][0] = M[0][2];
  M[2][1] = M[1][2];
  double sv[3], E[3][3];
  Builder::jacobi(M, sv, E);
  Builder::axis_from_eigen(E, sv, bv.axis);
  bv.rss_axis = bv.axis;
  const double real_max = std::numeric_limits<double>::max();
  double mn[3] = {real_max, real_max, real_max}, mx[3] = {-real_max, -real_max, -real_max};
  for (int i = 0; i < 12; ++i)
    for (int k = 0; k < 3; ++k) {
      const double proj = dot(col(bv.axis, k), ps[i]);
      if (proj > mx[k]) mx[k] = proj;
      if (proj < mn[k]) mn[k] = proj;
    }
  const Vec3 o{{(mx[0] + mn[0]) / 2, (mx[1] + mn[1]) / 2, (mx[2] + mn[2]) / 2}};
  bv.obb_To = mul(bv.axis, o);
  bv.obb_ext = Vec3{{(mx[0] - mn[0]) * 0.5, (mx[1] - mn[1]) * 0.5, (mx[2] - mn[2]) * 0.5}};
  bv.first_child = -1;
  bv.first_primitive = 0;
  bv.num_primitives = 0;
  // RSS part (RSS_fit_functions::fitn, math/bv/utility-inl.h:244-259): same covariance / eigen / axes, then
  // getRadiusAndOriginAndRectangleSize over the 12 points (point branch, math/geometry-inl.h:725-760: P[i] =
  // projections of ps[i]).  The triangle routine fed with 4 "triangles" (0,1,2) .. (9,10,11) builds the same P.
  {
    Model tmp;
    tmp.verts.assign(ps, ps + 12);
    for (int i = 0; i < 4; ++i) tmp.tris.push_back(Tri{{3 * i, 3 * i + 1, 3 * i + 2}});
    const unsigned idx[4] = {0, 1, 2, 3};
    Builder b{tmp, SPLIT_MEAN, {}, 0};
    b.rss_fit(idx, 4, bv.axis, bv.rss_To, bv.rss_l, bv.rss_r);
  }
}

}  // namespace

// beginModel/addSubModel/endModel/buildTree — BVH_model-inl.h:207-253,383-517,833-864
void build_model(Model& m, const std::vector<Vec3>& pts, const std::vector<Tri>& tris, SplitMethod split) {
  m.verts = pts;
  m.tris = tris;
  const int nt = (int)tris.size();
  m.nodes.assign(nt > 0 ? 2 * nt - 1 : 0, Node{});
  if (nt == 0) return;
  Builder b{m, split, {}, 1};
  b.prim.resize(nt);
  for (int i = 0; i < nt; ++i) b.prim[i] = i;
  b.recurse(0, 0, nt);
  m.prim = b.prim;
  m.split = split;
}

// refitTree_topdown — BVH_model-inl.h:1064-1076 (prev_vertices == nullptr after beginReplaceModel, :530-534)
void refit_topdown(Model& m, const std::vector<Vec3>& new_verts) {
  m.verts = new_verts;
  Builder b{m, m.split, m.prim, (int)m.nodes.size()};
  for (size_t i = 0; i < m.nodes.size(); ++i) {
    Node& nd = m.nodes[i];
    b.fit(b.prim.data() + nd.first_primitive, nd.num_primitives, nd);
  }
}

// -----------------------------------------------------------------------------
// Bottom-up refit — BVHModel::refitTree_bottomup / recursiveRefitTree_bottomup, BVH_model-inl.h:952-1037.
// prev_vertices == nullptr after beginReplaceModel (:530-534), so a leaf is fit(v, 3, bv) = fit3 and an inner node is
// bvs[left].bv + bvs[right].bv, children first.
//
// Eigen pieces restated from Eigen 3.3's published algorithms (Eigen itself is not on this image; the 4-term quaternion
// sums are evaluated left to right in coefficient order x, y, z, w -- same "association order unpinned" caveat as
// fcl_oracle_vec.hpp):
//   Vector3::normalize()            z = squaredNorm(); if (z > 0) v /= sqrt(z)          (Core/Dot.h)
//   Quaternion(Matrix3)             trace branch / largest-diagonal branch              (Geometry/Quaternion.h)
//   Quaternion::toRotationMatrix()  tx = 2x ... res(0,0) = 1 - (tyy + tzz) ...          (Geometry/Quaternion.h)
// -----------------------------------------------------------------------------
namespace {

inline void normalize_in_place(Vec3& v) {
  const double z = sqnorm(v);
  if (z > 0) {
    const double n = std::sqrt(z);
    v = Vec3{{v[0] / n, v[1] / n, v[2] / n}};
  }
}
inline void set_col(Mat3& A, int j, const Vec3& v) {
  for (int r = 0; r < 3; ++r) A.m[r][j] = v[r];
}

// getCovariance, point branch without indices — math/geometry-inl.h:1383-1425
void covariance_points(const Vec3* ps, int n, double M[3][3]) {
  double S1[3] = {0, 0, 0}, S2[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int i = 0; i < n; ++i) {
    const Vec3& p = ps[i];
    for (int k = 0; k < 3; ++k) S1[k] += p[k];
    S2[0][0] += (p[0] * p[0]);
    S2[1][1] += (p[1] * p[1]);
    S2[2][2] += (p[2] * p[2]);
    S2[0][1] += (p[0] * p[1]);
    S2[0][2] += (p[0] * p[2]);
    S2[1][2] += (p[1] * p[2]);
  }
  const int n_points = n;
  M[0][0] = S2[0][0] - S1[0] * S1[0] / n_points;
  M[1][1] = S2[1][1] - S1[1] * S1[1] / n_points;
  M[2][2] = S2[2][2] - S1[2] * S1[2] / n_points;
  M[0][1] = S2[0][1] - S1[0] * S1[1] / n_points;
  M[1][2] = S2[1][2] - S1[1] * S1[2] / n_points;
  M[0][2] = S2[0][2] - S1[0] * S1[2] / n_points;
  M[1][0] = M[0][1];
  M[2][0] = M[0][2];
  M[2][1] = M[1][2];
}

// getExtentAndCenter_pointcloud without indices / second frame — math/geometry-inl.h:229-292
void extent_and_center_points(const Vec3* ps, int n, const Mat3& axis, Vec3& center, Vec3& extent) {
  const double real_max = std::numeric_limits<double>::max();
  double mn[3] = {real_max, real_max, real_max}, mx[3] = {-real_max, -real_max, -real_max};
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k) {
      const double proj = dot(col(axis, k), ps[i]);
      if (proj > mx[k]) mx[k] = proj;
      if (proj < mn[k]) mn[k] = proj;
    }
  const Vec3 o{{(mx[0] + mn[0]) / 2, (mx[1] + mn[1]) / 2, (mx[2] + mn[2]) / 2}};
  center = mul(axis, o);
  extent = Vec3{{(mx[0] - mn[0]) * 0.5, (mx[1] - mn[1]) * 0.5, (mx[2] - mn[2]) * 0.5}};
}

// getRadiusAndOriginAndRectangleSize, point branch — math/geometry-inl.h:757-779, then :782-988
void rss_fit_points(const Vec3* ps, int n, const Mat3& axis, Vec3& origin, double l[2], double& r) {
  std::vector<Vec3> P(n);
  for (int i = 0; i < n; ++i) {
    P[i][0] = dot(col(axis, 0), ps[i]);
    P[i][1] = dot(col(axis, 1), ps[i]);
    P[i][2] = dot(col(axis, 2), ps[i]);
  }
  Builder::rss_from_projections(P, axis, origin, l, r);
}

// eigenvalue order of axisFromEigen / merge_largedist / RSS::operator+ (same if-chain in all three)
inline void order3(const double s[3], int& mn, int& mid, int& mx) {
  if (s[0] > s[1]) { mx = 0; mn = 1; } else { mn = 0; mx = 1; }
  if (s[2] < s[mn]) { mid = mn; mn = 2; }
  else if (s[2] > s[mx]) { mid = mx; mx = 2; }
  else { mid = 2; }
}

// computeVertices — math/bv/OBB-inl.h:233-250
void obb_vertices(const Mat3& axis, const Vec3& To, const Vec3& extent, Vec3 v[8]) {
  const Vec3 e0 = scale(col(axis, 0), extent[0]), e1 = scale(col(axis, 1), extent[1]), e2 = scale(col(axis, 2), extent[2]);
  v[0] = sub(sub(sub(To, e0), e1), e2);
  v[1] = sub(sub(add(To, e0), e1), e2);
  v[2] = sub(add(add(To, e0), e1), e2);
  v[3] = sub(add(sub(To, e0), e1), e2);
  v[4] = add(sub(sub(To, e0), e1), e2);
  v[5] = add(sub(add(To, e0), e1), e2);
  v[6] = add(add(add(To, e0), e1), e2);
  v[7] = add(add(sub(To, e0), e1), e2);
}

// merge_largedist — math/bv/OBB-inl.h:254-312
void obb_merge_largedist(const Node& b1, const Node& b2, Node& b) {
  Vec3 vertex[16];
  obb_vertices(b1.axis, b1.obb_To, b1.obb_ext, vertex);
  obb_vertices(b2.axis, b2.obb_To, b2.obb_ext, vertex + 8);
  Vec3 a0 = sub(b1.obb_To, b2.obb_To);
  normalize_in_place(a0);
  Vec3 proj[16];
  for (int i = 0; i < 16; ++i) proj[i] = sub(vertex[i], scale(a0, dot(vertex[i], a0)));
  double M[3][3], s[3], v[3][3];
  covariance_points(proj, 16, M);
  Builder::jacobi(M, s, v);
  int mn, mid, mx;
  order3(s, mn, mid, mx);
  // E = vout with vout.col(k) = (v[k][0], v[k][1], v[k][2]) (eigen_old, :503-505), i.e. E(r, c) = v[c][r];
  // axis.col(1) << E.col(0)[max], E.col(1)[max], E.col(2)[max] = (E(max,0), E(max,1), E(max,2)) = (v[0][max], v[1][max], v[2][max])
  set_col(b.axis, 0, a0);
  set_col(b.axis, 1, Vec3{{v[0][mx], v[1][mx], v[2][mx]}});
  set_col(b.axis, 2, Vec3{{v[0][mid], v[1][mid], v[2][mid]}});
  extent_and_center_points(vertex, 16, b.axis, b.obb_To, b.obb_ext);
}

// merge_smalldist — math/bv/OBB-inl.h:316-369
struct Quat {
  double x, y, z, w;
};
Quat quat_from(const Mat3& A) {  // Eigen quaternionbase_assign_impl<Matrix3>
  Quat q;
  double t = (A.m[0][0] + A.m[1][1]) + A.m[2][2];
  if (t > 0) {
    t = std::sqrt(t + 1.0);
    q.w = 0.5 * t;
    t = 0.5 / t;
    q.x = (A.m[2][1] - A.m[1][2]) * t;
    q.y = (A.m[0][2] - A.m[2][0]) * t;
    q.z = (A.m[1][0] - A.m[0][1]) * t;
  } else {
    int i = 0;
    if (A.m[1][1] > A.m[0][0]) i = 1;
    if (A.m[2][2] > A.m[i][i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(((A.m[i][i] - A.m[j][j]) - A.m[k][k]) + 1.0);
    double c[3];
    c[i] = 0.5 * t;
    t = 0.5 / t;
    q.w = (A.m[k][j] - A.m[j][k]) * t;
    c[j] = (A.m[j][i] + A.m[i][j]) * t;
    c[k] = (A.m[k][i] + A.m[i][k]) * t;
    q.x = c[0];
    q.y = c[1];
    q.z = c[2];
  }
  return q;
}
Mat3 quat_to_matrix(const Quat& q) {  // Eigen QuaternionBase::toRotationMatrix
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  Mat3 R;
  R.m[0][0] = 1 - (tyy + tzz);
  R.m[0][1] = txy - twz;
  R.m[0][2] = txz + twy;
  R.m[1][0] = txy + twz;
  R.m[1][1] = 1 - (txx + tzz);
  R.m[1][2] = tyz - twx;
  R.m[2][0] = txz - twy;
  R.m[2][1] = tyz + twx;
  R.m[2][2] = 1 - (txx + tyy);
  return R;
}
void obb_merge_smalldist(const Node& b1, const Node& b2, Node& b) {
  Vec3 To = scale(add(b1.obb_To, b2.obb_To), 0.5);
  const Quat q0 = quat_from(b1.axis);
  Quat q1 = quat_from(b2.axis);
  if ((((q0.x * q1.x + q0.y * q1.y) + q0.z * q1.z) + q0.w * q1.w) < 0) q1 = Quat{-q1.x, -q1.y, -q1.z, -q1.w};
  Quat q{q0.x + q1.x, q0.y + q1.y, q0.z + q1.z, q0.w + q1.w};
  {
    const double z = ((q.x * q.x + q.y * q.y) + q.z * q.z) + q.w * q.w;
    if (z > 0) {
      const double n = std::sqrt(z);
      q = Quat{q.x / n, q.y / n, q.z / n, q.w / n};
    }
  }
  b.axis = quat_to_matrix(q);
  const double real_max = std::numeric_limits<double>::max();
  double pmin[3] = {real_max, real_max, real_max}, pmax[3] = {-real_max, -real_max, -real_max};
  Vec3 vertex[8];
  for (int which = 0; which < 2; ++which) {
    const Node& s = which == 0 ? b1 : b2;
    obb_vertices(s.axis, s.obb_To, s.obb_ext, vertex);
    for (int i = 0; i < 8; ++i) {
      const Vec3 diff = sub(vertex[i], To);
      for (int j = 0; j < 3; ++j) {
        const double d = dot(diff, col(b.axis, j));
        if (d > pmax[j]) pmax[j] = d;
        else if (d < pmin[j]) pmin[j] = d;   // `else if`, like the reference (:339-342)
      }
    }
  }
  for (int j = 0; j < 3; ++j) {
    To = add(To, scale(col(b.axis, j), 0.5 * (pmax[j] + pmin[j])));
    b.obb_ext[j] = 0.5 * (pmax[j] - pmin[j]);
  }
  b.obb_To = To;
}

// OBB::operator+ — math/bv/OBB-inl.h:161-174
void obb_merge(const Node& a, const Node& o, Node& out) {
  const Vec3 center_diff = sub(a.obb_To, o.obb_To);
  const double max_extent = std::max(std::max(a.obb_ext[0], a.obb_ext[1]), a.obb_ext[2]);
  const double max_extent2 = std::max(std::max(o.obb_ext[0], o.obb_ext[1]), o.obb_ext[2]);
  if (norm(center_diff) > 2 * (max_extent + max_extent2)) obb_merge_largedist(a, o, out);
  else obb_merge_smalldist(a, o, out);
}

// RSS::operator+ — math/bv/RSS-inl.h:313-371.  Two things are kept exactly as the reference has them: the new in-plane
// axes are E.col(max) / E.col(mid) of eigen_old's output (E(r, c) = v[c][r], so E.col(k) = ROW k of the Jacobi
// eigenvector matrix, not the eigenvector axisFromEigen would pick), and the third axis is the cross product of
// *this*'s first two axes (`axis.col(0).cross(axis.col(1))`, :364), not of the new ones.
void rss_merge(const Node& a, const Node& o, Node& out) {
  Vec3 v[16];
  auto corners = [](const Node& s, Vec3* dst) {
    const Vec3 d0_pos = scale(col(s.rss_axis, 0), s.rss_l[0] + s.rss_r);
    const Vec3 d1_pos = scale(col(s.rss_axis, 1), s.rss_l[1] + s.rss_r);
    const Vec3 d0_neg = scale(col(s.rss_axis, 0), -s.rss_r);
    const Vec3 d1_neg = scale(col(s.rss_axis, 1), -s.rss_r);
    const Vec3 d2_pos = scale(col(s.rss_axis, 2), s.rss_r);
    const Vec3 d2_neg = scale(col(s.rss_axis, 2), -s.rss_r);
    dst[0] = add(add(add(s.rss_To, d0_pos), d1_pos), d2_pos);
    dst[1] = add(add(add(s.rss_To, d0_pos), d1_pos), d2_neg);
    dst[2] = add(add(add(s.rss_To, d0_pos), d1_neg), d2_pos);
    dst[3] = add(add(add(s.rss_To, d0_pos), d1_neg), d2_neg);
    dst[4] = add(add(add(s.rss_To, d0_neg), d1_pos), d2_pos);
    dst[5] = add(add(add(s.rss_To, d0_neg), d1_pos), d2_neg);
    dst[6] = add(add(add(s.rss_To, d0_neg), d1_neg), d2_pos);
    dst[7] = add(add(add(s.rss_To, d0_neg), d1_neg), d2_neg);
  };
  corners(o, v);      // v[0..7]: other
  corners(a, v + 8);  // v[8..15]: *this
  double M[3][3], s[3], e[3][3];
  covariance_points(v, 16, M);
  Builder::jacobi(M, s, e);
  int mn, mid, mx;
  order3(s, mn, mid, mx);
  Mat3 A;
  set_col(A, 0, Vec3{{e[mx][0], e[mx][1], e[mx][2]}});     // E.col(max)
  set_col(A, 1, Vec3{{e[mid][0], e[mid][1], e[mid][2]}});  // E.col(mid)
  set_col(A, 2, cross(col(a.rss_axis, 0), col(a.rss_axis, 1)));
  out.rss_axis = A;
  rss_fit_points(v, 16, A, out.rss_To, out.rss_l, out.rss_r);
}

}  // namespace

// OBBRSS_fit_functions::fit3 — math/bv/utility-inl.h:92-117 (OBB), :208-230 (RSS), :507-511
void fit3_obbrss(const Vec3 ps[3], Node& out) {
  Vec3 e[3] = {sub(ps[0], ps[1]), sub(ps[1], ps[2]), sub(ps[2], ps[0])};
  const double len[3] = {sqnorm(e[0]), sqnorm(e[1]), sqnorm(e[2])};
  int imax = 0;
  if (len[1] > len[0]) imax = 1;
  if (len[2] > len[imax]) imax = 2;
  Vec3 c2 = cross(e[0], e[1]);
  normalize_in_place(c2);
  Vec3 c0 = e[imax];
  normalize_in_place(c0);
  set_col(out.axis, 2, c2);
  set_col(out.axis, 0, c0);
  set_col(out.axis, 1, cross(c2, c0));
  extent_and_center_points(ps, 3, out.axis, out.obb_To, out.obb_ext);
  // RSS: e[0].cross(e[1]).normalized(), e[imax].normalized() -- the same values (normalized() = copy + normalize())
  out.rss_axis = out.axis;
  rss_fit_points(ps, 3, out.rss_axis, out.rss_To, out.rss_l, out.rss_r);
}

// OBBRSS::operator+ — math/bv/OBBRSS-inl.h:95-101
void merge_obbrss(const Node& a, const Node& b, Node& out) {
  Node r = out;
  obb_merge(a, b, r);
  rss_merge(a, b, r);
  out.axis = r.axis;
  out.obb_To = r.obb_To;
  out.obb_ext = r.obb_ext;
  out.rss_axis = r.rss_axis;
  out.rss_To = r.rss_To;
  out.rss_l[0] = r.rss_l[0];
  out.rss_l[1] = r.rss_l[1];
  out.rss_r = r.rss_r;
}

void refit_bottomup(Model& m, const std::vector<Vec3>& new_verts) {
  m.verts = new_verts;
  // recursiveRefitTree_bottomup(0): post-order; iterative with an explicit stack (trees can be 10^6 nodes deep in the
  // worst case)
  if (m.nodes.empty()) return;
  std::vector<std::pair<int, int>> st;  // (node, state)
  st.push_back({0, 0});
  while (!st.empty()) {
    auto [id, state] = st.back();
    Node& nd = m.nodes[id];
    if (nd.first_child < 0) {
      const Tri& t = m.tris[-(nd.first_child + 1)];
      const Vec3 ps[3] = {m.verts[t.v[0]], m.verts[t.v[1]], m.verts[t.v[2]]};
      fit3_obbrss(ps, nd);
      st.pop_back();
    } else if (state == 0) {
      st.back().second = 1;
      st.push_back({nd.first_child + 1, 0});  // right is pushed first so that left is processed first
      st.push_back({nd.first_child, 0});
    } else {
      merge_obbrss(m.nodes[nd.first_child], m.nodes[nd.first_child + 1], nd);
      st.pop_back();
    }
  }
}

// -----------------------------------------------------------------------------
// Collision traversal
// -----------------------------------------------------------------------------
namespace {

inline double node_size(const Node& n) {  // OBB::size() = extent.squaredNorm(), OBB-inl.h:206-209
  return sqnorm(n.obb_ext);
}

// firstOverSecond — bvh_collision_traversal_node-inl.h:78-90 (same in distance: bvh_distance_…:78-90)
inline bool first_over_second(const Node& n1, const Node& n2) {
  bool l1 = n1.first_child < 0, l2 = n2.first_child < 0;
  return l2 || (!l1 && (node_size(n1) > node_size(n2)));
}

struct CollideCtx {
  const Model& m1;
  const Model& m2;
  Pose tf1;
  Mat3 R;  // relative rotation  R1^T R2
  Vec3 T;  // relative translation R1^T (t2 - t1)
  size_t max_contacts;
  bool enable_contact;
  std::vector<Contact>& out;
  long long n_bv = 0, n_leaf = 0;

  bool can_stop() const {  // collision_request-inl.h:77-82 (enable_cost == false)
    return !out.empty() && max_contacts <= out.size();
  }

  // meshCollisionOrientedNodeLeafTesting — mesh_collision_traversal_node-inl.h:527-620
  void leaf(int b1, int b2) {
    n_leaf++;
    int id1 = -(m1.nodes[b1].first_child + 1);
    int id2 = -(m2.nodes[b2].first_child + 1);
    const Tri& t1 = m1.tris[id1];
    const Tri& t2 = m2.tris[id2];
    Vec3 P[3] = {m1.verts[t1.v[0]], m1.verts[t1.v[1]], m1.verts[t1.v[2]]};
    Vec3 Q[3] = {m2.verts[t2.v[0]], m2.verts[t2.v[1]], m2.verts[t2.v[2]]};
    if (!enable_contact) {
      if (tri_intersect(P, Q, R, T, nullptr, nullptr, nullptr, nullptr)) {
        if (out.size() < max_contacts) {
          Contact c{};
          c.b1 = id1;
          c.b2 = id2;
          out.push_back(c);
        }
      }
    } else {
      double penetration;
      Vec3 normal;
      unsigned n_contacts;
      Vec3 contacts[2];
      if (tri_intersect(P, Q, R, T, contacts, &n_contacts, &penetration, &normal)) {
        if (max_contacts < out.size() + n_contacts)
          n_contacts = (max_contacts > out.size()) ? (unsigned)(max_contacts - out.size()) : 0;
        for (unsigned i = 0; i < n_contacts; ++i) {
          Contact c;
          c.b1 = id1;
          c.b2 = id2;
          c.pos = add(mul(tf1.R, contacts[i]), tf1.t);  // tf1 * p
          c.normal = mul(tf1.R, normal);                // tf1.linear() * n
          c.depth = penetration;
          out.push_back(c);
        }
      }
    }
  }

  bool bv_disjoint(int b1, int b2) {  // MeshCollisionTraversalNodeOBBRSS::BVTesting, :495-500
    n_bv++;
    return !obb_overlap(R, T, m1.nodes[b1], m2.nodes[b2]);
  }

  // collisionRecurse — traversal/traversal_recurse-inl.h:84-130 (front_list == nullptr)
  void recurse(int b1, int b2) {
    const Node& n1 = m1.nodes[b1];
    const Node& n2 = m2.nodes[b2];
    bool l1 = n1.first_child < 0, l2 = n2.first_child < 0;
    if (l1 && l2) {
      if (bv_disjoint(b1, b2)) return;
      leaf(b1, b2);
      return;
    }
    if (bv_disjoint(b1, b2)) return;
    if (first_over_second(n1, n2)) {
      int c1 = n1.first_child, c2 = n1.first_child + 1;
      recurse(c1, b2);
      if (can_stop()) return;
      recurse(c2, b2);
    } else {
      int c1 = n2.first_child, c2 = n2.first_child + 1;
      recurse(b1, c1);
      if (can_stop()) return;
      recurse(b1, c2);
    }
  }
};

}  // namespace

// fcl::collide → BVHCollide → orientedMeshCollide — collision-inl.h:95-150,
// detail/collision_func_matrix-inl.h:571-590; setup: mesh_collision_traversal_node-inl.h:716-745
// and relativeTransform, math/geometry-inl.h:681-682.
size_t collide(const Model& m1, const Pose& tf1, const Model& m2, const Pose& tf2,
               size_t num_max_contacts, bool enable_contact, std::vector<Contact>& out,
               CollideStats* stats) {
  if (num_max_contacts == 0) return 0;                              // collision-inl.h:111-115
  if (!out.empty() && num_max_contacts <= out.size()) return out.size();  // isSatisfied on entry
  if (m1.nodes.empty() || m2.nodes.empty()) return out.size();
  CollideCtx ctx{m1, m2, tf1, mulTN(tf1.R, tf2.R), mulTv(tf1.R, sub(tf2.t, tf1.t)),
                 num_max_contacts, enable_contact, out};
  ctx.recurse(0, 0);
  if (stats) {
    stats->n_bv += ctx.n_bv;
    stats->n_leaf += ctx.n_leaf;
  }
  return out.size();
}

void brute_collide_pairs(const Model& m1, const Pose& tf1, const Model& m2, const Pose& tf2,
                         std::vector<std::pair<int, int>>& pairs) {
  Mat3 R = mulTN(tf1.R, tf2.R);
  Vec3 T = mulTv(tf1.R, sub(tf2.t, tf1.t));
  for (int i = 0; i < (int)m1.tris.size(); ++i) {
    const Tri& t1 = m1.tris[i];
    Vec3 P[3] = {m1.verts[t1.v[0]], m1.verts[t1.v[1]], m1.verts[t1.v[2]]};
    for (int j = 0; j < (int)m2.tris.size(); ++j) {
      const Tri& t2 = m2.tris[j];
      Vec3 Q[3] = {m2.verts[t2.v[0]], m2.verts[t2.v[1]], m2.verts[t2.v[2]]};
      if (tri_intersect(P, Q, R, T, nullptr, nullptr, nullptr, nullptr)) pairs.emplace_back(i, j);
    }
  }
}

// -----------------------------------------------------------------------------
// Distance traversal
// -----------------------------------------------------------------------------
namespace {

struct DistCtx {
  const Model& m1;
  const Model& m2;
  Mat3 R;  // tf = tf1^-1 * tf2  (mesh_distance_traversal_node-inl.h:630)
  Vec3 T;
  bool nearest;
  double min_distance = std::numeric_limits<double>::max();  // distance_result.h:94
  Vec3 p1{{0, 0, 0}}, p2{{0, 0, 0}};
  int rb1 = -1, rb2 = -1;
  long long n_bv = 0, n_leaf = 0;

  void update(double d, int id1, int id2, const Vec3& P1, const Vec3& P2) {  // distance_result-inl.h:66-103
    if (min_distance > d) {
      min_distance = d;
      rb1 = id1;
      rb2 = id2;
      if (nearest) { p1 = P1; p2 = P2; }
    }
  }

  // triDistance(S1,S2,S3,T1,T2,T3,tf,P,Q) — triangle_distance-inl.h:453-462; tf*p = R p + T
  void tri_pair(int id1, int id2) {
    const Tri& t1 = m1.tris[id1];
    const Tri& t2 = m2.tris[id2];
    Vec3 S[3] = {m1.verts[t1.v[0]], m1.verts[t1.v[1]], m1.verts[t1.v[2]]};
    Vec3 Tt[3];
    for (int k = 0; k < 3; ++k) Tt[k] = add(mul(R, m2.verts[t2.v[k]]), T);
    Vec3 P1, P2;
    double d = tri_distance(S, Tt, P1, P2);
    update(d, id1, id2, P1, P2);
  }

  void leaf(int b1, int b2) {  // meshDistanceOrientedNodeLeafTesting, :453-499
    n_leaf++;
    tri_pair(-(m1.nodes[b1].first_child + 1), -(m2.nodes[b2].first_child + 1));
  }

  double bv(int b1, int b2) {  // mesh_distance_traversal_node.h:188-193
    n_bv++;
    return rss_distance(R, T, m1.nodes[b1], m2.nodes[b2]);
  }

  // MeshDistanceTraversalNode::canStop with rel_err = abs_err = 0 (latched from a
  // default request in the constructor, mesh_distance_traversal_node-inl.h:96-105,148-153)
  bool can_stop(double c) const { return (c >= min_distance - 0.0) && (c * (1 + 0.0) >= min_distance); }

  // distanceRecurse — traversal_recurse-inl.h:259-316
  void recurse(int b1, int b2) {
    const Node& n1 = m1.nodes[b1];
    const Node& n2 = m2.nodes[b2];
    bool l1 = n1.first_child < 0, l2 = n2.first_child < 0;
    if (l1 && l2) {
      leaf(b1, b2);
      return;
    }
    int a1, a2, c1, c2;
    if (first_over_second(n1, n2)) {
      a1 = n1.first_child; a2 = b2; c1 = n1.first_child + 1; c2 = b2;
    } else {
      a1 = b1; a2 = n2.first_child; c1 = b1; c2 = n2.first_child + 1;
    }
    double d1 = bv(a1, a2);
    double d2 = bv(c1, c2);
    if (d2 < d1) {
      if (!can_stop(d2)) recurse(c1, c2);
      if (!can_stop(d1)) recurse(a1, a2);
    } else {
      if (!can_stop(d1)) recurse(a1, a2);
      if (!can_stop(d2)) recurse(c1, c2);
    }
  }

  // distanceQueueRecurse — traversal_recurse-inl.h:321-460
  struct BVT {
    double d;
    int b1, b2;
  };
  struct Cmp {
    bool operator()(const BVT& l, const BVT& r) const { return l.d > r.d; }
  };
  void queue_recurse(int b1, int b2, unsigned qsize) {
    std::priority_queue<BVT, std::vector<BVT>, Cmp> pq;
    BVT cur{0, b1, b2};
    while (true) {
      const Node& n1 = m1.nodes[cur.b1];
      const Node& n2 = m2.nodes[cur.b2];
      bool l1 = n1.first_child < 0, l2 = n2.first_child < 0;
      if (l1 && l2) {
        leaf(cur.b1, cur.b2);
      } else if (pq.size() + 1 >= qsize) {
        queue_recurse(cur.b1, cur.b2, qsize);
      } else {
        BVT x, y;
        if (first_over_second(n1, n2)) {
          x.b1 = n1.first_child; x.b2 = cur.b2;
          x.d = bv(x.b1, x.b2);
          y.b1 = n1.first_child + 1; y.b2 = cur.b2;
          y.d = bv(y.b1, y.b2);
        } else {
          x.b1 = cur.b1; x.b2 = n2.first_child;
          x.d = bv(x.b1, x.b2);
          y.b1 = cur.b1; y.b2 = n2.first_child + 1;
          y.d = bv(y.b1, y.b2);
        }
        pq.push(x);
        pq.push(y);
      }
      if (pq.empty()) break;
      cur = pq.top();
      pq.pop();
      if (can_stop(cur.d)) break;
    }
  }
};

inline void relative_for_distance(const Pose& tf1, const Pose& tf2, Mat3& R, Vec3& T) {
  // tf1.inverse(Isometry) * tf2:  linear = R1^T R2,  translation = R1^T t2 + (-(R1^T t1))
  R = mulTN(tf1.R, tf2.R);
  Vec3 inv_t = mulTv(tf1.R, tf1.t);
  inv_t = Vec3{{-inv_t[0], -inv_t[1], -inv_t[2]}};
  T = add(mulTv(tf1.R, tf2.t), inv_t);
}

}  // namespace

// fcl::distance → BVHDistance → orientedMeshDistance — distance-inl.h:92-190,
// detail/distance_func_matrix-inl.h:386-403; driver collision_node-inl.h:137-146.
double distance(const Model& m1, const Pose& tf1, const Model& m2, const Pose& tf2,
                bool enable_nearest_points, DistanceOut& out, int qsize, CollideStats* stats) {
  DistCtx ctx{m1, m2, Mat3{}, Vec3{}, enable_nearest_points};
  relative_for_distance(tf1, tf2, ctx.R, ctx.T);
  // preprocess: seed with triangle 0 / triangle 0 (:352-366 → :546-586)
  ctx.tri_pair(0, 0);
  if (qsize <= 2) ctx.recurse(0, 0);
  else ctx.queue_recurse(0, 0, (unsigned)qsize);
  // postprocess: both points are in model1's frame → world with tf1 (:590-603)
  out.min_distance = ctx.min_distance;
  out.b1 = ctx.rb1;
  out.b2 = ctx.rb2;
  if (enable_nearest_points) {
    out.p1 = add(mul(tf1.R, ctx.p1), tf1.t);
    out.p2 = add(mul(tf1.R, ctx.p2), tf1.t);
  } else {
    out.p1 = out.p2 = Vec3{{0, 0, 0}};
  }
  if (stats) {
    stats->n_bv += ctx.n_bv;
    stats->n_leaf += ctx.n_leaf;
  }
  return out.min_distance;
}

double brute_distance(const Model& m1, const Pose& tf1, const Model& m2, const Pose& tf2, DistanceOut& out) {
  DistCtx ctx{m1, m2, Mat3{}, Vec3{}, true};
  relative_for_distance(tf1, tf2, ctx.R, ctx.T);
  for (int i = 0; i < (int)m1.tris.size(); ++i)
    for (int j = 0; j < (int)m2.tris.size(); ++j) ctx.tri_pair(i, j);
  out.min_distance = ctx.min_distance;
  out.b1 = ctx.rb1;
  out.b2 = ctx.rb2;
  out.p1 = add(mul(tf1.R, ctx.p1), tf1.t);
  out.p2 = add(mul(tf1.R, ctx.p2), tf1.t);
  return out.min_distance;
}

// ---------------------------------------------------------------------------------------
// Continuous collision by conservative advancement, both bodies translating (CCDM_TRANS):
//   continuousCollide(o1, tf1_beg, tf1_end, o2, tf2_beg, tf2_end, request{CCDM_TRANS, CCDC_CONSERVATIVE_ADVANCEMENT})
//     narrowphase/continuous_collision-inl.h:441-452 (getMotionBase :93-117 -> TranslationMotion), :355-377, :302-352
//   -> conservative_advancement_matrix[BV_OBBRSS][BV_OBBRSS] = BVHConservativeAdvancement<OBBRSS>
//     detail/conservative_advancement_func_matrix-inl.h:692-712, 675-690
//   -> conservativeAdvancementMeshOriented<OBBRSS, MeshConservativeAdvancementTraversalNodeOBBRSS>   :149-219
//   traversal node: detail/traversal/distance/mesh_conservative_advancement_traversal_node.h:163-215 (BVTesting),
//     -inl.h:432-470 (OBBRSS leafTesting / canStop), :641-713 (meshConservativeAdvancementOrientedNodeLeafTesting),
//     :566-637 (meshConservativeAdvancementOrientedNodeCanStop)
//   TranslationMotion: math/motion/translation_motion-inl.h:47-134; bounds: triangle_motion_bound_visitor-inl.h:218-227,
//     tbv_motion_bound_visitor-inl.h:52-62 -- TBVMotionBoundVisitorVisitImpl is specialised for RSS only, so for
//     BV = OBBRSS the generic template answers 0: a pruned node pair never shortens the step (bound 0 <= c gives
//     cur_delta_t = 1) and only the triangle pairs actually visited do.  Kept as the reference has it.
// Eigen pieces (published algorithms, restated above for the bottom-up refit): Quaternion(Matrix3), toRotationMatrix,
// Quaternion * Vector3 (_transformVector: uv = 2 (vec x v); v + w uv + vec x uv), Vector3::normalize.
// ---------------------------------------------------------------------------------------
namespace {

struct TransMotion {  // TranslationMotion<S>
  Quat rot;
  Vec3 trans_start, trans_range;
  Pose tf;  // current transform
  TransMotion(const Pose& tf1, const Pose& tf2) : rot(quat_from(tf1.R)), trans_start(tf1.t), trans_range(sub(tf2.t, tf1.t)), tf(tf1) {}
  void integrate(double dt) {
    if (dt > 1) dt = 1;
    tf.R = quat_to_matrix(rot);
    tf.t = add(trans_start, scale(trans_range, dt));
  }
};

inline Vec3 quat_rotate(const Quat& q, const Vec3& v) {  // Eigen QuaternionBase::_transformVector
  const Vec3 qv{{q.x, q.y, q.z}};
  Vec3 uv = cross(qv, v);
  uv = add(uv, uv);
  return add(add(v, scale(uv, q.w)), cross(qv, uv));
}

struct CaStackData {  // ConservativeAdvancementStackData (P1, P2 of the box pair are never read for OBBRSS: bound = 0)
  int c1, c2;
  double d;
};

struct CaCtx {
  const Model& m1;
  const Model& m2;
  const TransMotion* motion1;
  const TransMotion* motion2;
  Mat3 R{};
  Vec3 T{};
  double min_distance = std::numeric_limits<double>::max();
  double delta_t = 1, toc = 0, t_err = 0.00001, w = 1;  // constructor defaults, -inl.h:84-97
  CaCtx(const Model& a, const Model& b, const TransMotion* x, const TransMotion* y) : m1(a), m2(b), motion1(x), motion2(y) {}
  std::vector<CaStackData> stack;
  long long n_bv = 0, n_leaf = 0;

  double bv(int b1, int b2) {  // BVTesting, mesh_conservative_advancement_traversal_node.h:172-189
    n_bv++;
    const double d = rss_distance(R, T, m1.nodes[b1], m2.nodes[b2]);
    stack.push_back({b1, b2, d});
    return d;
  }

  void leaf(int b1, int b2) {  // meshConservativeAdvancementOrientedNodeLeafTesting
    n_leaf++;
    const Tri& t1 = m1.tris[-(m1.nodes[b1].first_child + 1)];
    const Tri& t2 = m2.tris[-(m2.nodes[b2].first_child + 1)];
    const Vec3 S[3] = {m1.verts[t1.v[0]], m1.verts[t1.v[1]], m1.verts[t1.v[2]]};
    Vec3 Tt[3];
    for (int k = 0; k < 3; ++k) Tt[k] = add(mul(R, m2.verts[t2.v[k]]), T);  // triDistance(..., R, T, P, Q)
    Vec3 P1, P2;
    const double d = tri_distance(S, Tt, P1, P2);
    if (d < min_distance) min_distance = d;
    const Vec3 n = sub(P2, P1);
    const Quat R0 = quat_from(motion1->tf.R);  // getCurrentRotation(Quaternion&): Q = tf.linear()
    Vec3 nt = quat_rotate(R0, n);
    normalize_in_place(nt);
    const Vec3 neg{{-nt[0], -nt[1], -nt[2]}};
    const double bound1 = dot(motion1->trans_range, nt);   // TriangleMotionBoundVisitor / TranslationMotion: velocity . n
    const double bound2 = dot(motion2->trans_range, neg);
    const double bound = bound1 + bound2;
    double cur_delta_t;
    if (bound <= d) cur_delta_t = 1;
    else cur_delta_t = d / bound;
    if (cur_delta_t < delta_t) delta_t = cur_delta_t;
  }

  bool can_stop(double c) {  // meshConservativeAdvancementOrientedNodeCanStop with abs_err = rel_err = 0
    if ((c >= w * (min_distance - 0.0)) && (c * (1 + 0.0) >= w * min_distance)) {
      const CaStackData& data = stack.back();
      if (data.d > c) stack[stack.size() - 2] = stack[stack.size() - 1];
      // bound1 = bound2 = 0 (generic TBVMotionBoundVisitorVisitImpl for OBBRSS): bound <= c, cur_delta_t = 1
      const double bound = 0.0 + 0.0;
      double cur_delta_t;
      if (bound <= c) cur_delta_t = 1;
      else cur_delta_t = c / bound;
      if (cur_delta_t < delta_t) delta_t = cur_delta_t;
      stack.pop_back();
      return true;
    }
    const CaStackData& data = stack.back();
    if (data.d > c) stack[stack.size() - 2] = stack[stack.size() - 1];
    stack.pop_back();
    return false;
  }

  void recurse(int b1, int b2) {  // distanceRecurse, traversal_recurse-inl.h:259-316
    const Node& n1 = m1.nodes[b1];
    const Node& n2 = m2.nodes[b2];
    const bool l1 = n1.first_child < 0, l2 = n2.first_child < 0;
    if (l1 && l2) {
      leaf(b1, b2);
      return;
    }
    int a1, a2, c1, c2;
    if (first_over_second(n1, n2)) {
      a1 = n1.first_child; a2 = b2; c1 = n1.first_child + 1; c2 = b2;
    } else {
      a1 = b1; a2 = n2.first_child; c1 = b1; c2 = n2.first_child + 1;
    }
    const double d1 = bv(a1, a2);
    const double d2 = bv(c1, c2);
    if (d2 < d1) {
      if (!can_stop(d2)) recurse(c1, c2);
      if (!can_stop(d1)) recurse(a1, a2);
    } else {
      if (!can_stop(d1)) recurse(a1, a2);
      if (!can_stop(d2)) recurse(c1, c2);
    }
  }
};

}  // namespace

double continuous_collide_translation(const Model& m1, const Pose& tf1_beg, const Pose& tf1_end, const Model& m2,
                                      const Pose& tf2_beg, const Pose& tf2_end, ContinuousOut& out) {
  TransMotion motion1(tf1_beg, tf1_end), motion2(tf2_beg, tf2_end);
  out.is_collide = false;
  out.time_of_contact = 1.0;  // ContinuousCollisionResult()
  out.iterations = 0;
  out.contact_tf1 = tf1_beg;
  out.contact_tf2 = tf2_beg;
  double toc;
  bool is_collide;
  // conservativeAdvancementMeshOriented: collision at the start configuration?
  std::vector<Contact> contacts;
  if (collide(m1, motion1.tf, m2, motion2.tf, 1, false, contacts) > 0) {
    toc = 0;
    is_collide = true;
  } else {
    CaCtx node(m1, m2, &motion1, &motion2);
    do {
      relative_for_distance(motion1.tf, motion2.tf, node.R, node.T);  // tf1.inverse(Isometry) * tf2
      node.delta_t = 1;
      node.min_distance = std::numeric_limits<double>::max();
      node.stack.clear();
      node.recurse(0, 0);
      out.iterations++;
      if (node.delta_t <= node.t_err) break;
      node.toc += node.delta_t;
      if (node.toc > 1) {
        node.toc = 1;
        break;
      }
      motion1.integrate(node.toc);
      motion2.integrate(node.toc);
    } while (1);
    toc = node.toc;
    is_collide = node.toc < 1;
  }
  out.is_collide = is_collide;
  out.time_of_contact = toc;
  if (is_collide) {  // continuousCollideConservativeAdvancement, continuous_collision-inl.h:339-350
    motion1.integrate(toc);
    motion2.integrate(toc);
    out.contact_tf1 = motion1.tf;
    out.contact_tf2 = motion2.tf;
  }
  return toc;
}

// ---------------------------------------------------------------------------------------
// Mesh <-> sphere collide: BVHShapeCollider<OBBRSS, Sphere>::collide -> orientedBVHShapeCollide
// (detail/collision_func_matrix-inl.h:378-430) -> initialize / setupMeshShapeCollisionOrientedNode
// (computeBV(model2, tf2, model2_bv)) -> collisionRecurse with a leaf second node.
// ---------------------------------------------------------------------------------------
void sphere_obb(double radius, const Pose& tf, Node& bv) { sphere_obb_impl(radius, tf, bv); }

namespace {
struct MeshSphereCtx {
  const Model& m1;
  Pose tf1, tf2;
  double radius;
  Node shape_bv;
  size_t max_contacts;
  bool enable_contact;
  std::vector<Contact>& out;
  long long n_bv = 0, n_leaf = 0;

  bool can_stop() const { return !out.empty() && max_contacts <= out.size(); }

  // MeshShapeCollisionTraversalNodeOBBRSS::BVTesting (mesh_shape_collision_traversal_node-inl.h:398-404):
  // !overlap(tf1.linear(), tf1.translation(), model2_bv, model1->getBV(b1).bv)
  bool bv_disjoint(int b1) {
    n_bv++;
    return !obb_overlap(tf1.R, tf1.t, shape_bv, m1.nodes[b1]);
  }

  // meshShapeCollisionOrientedNodeLeafTesting (:193-262) with the sphere specialisation of the transformed
  // shapeTriangleIntersect (gjk_solver_libccd-inl.h:479-497): triangle moved to the world by tf1
  void leaf(int b1) {
    n_leaf++;
    const int id = -(m1.nodes[b1].first_child + 1);
    const Tri& t = m1.tris[id];
    const Vec3 p1 = add(mul(tf1.R, m1.verts[t.v[0]]), tf1.t);
    const Vec3 p2 = add(mul(tf1.R, m1.verts[t.v[1]]), tf1.t);
    const Vec3 p3 = add(mul(tf1.R, m1.verts[t.v[2]]), tf1.t);
    if (!enable_contact) {
      if (sphere_tri_intersect(tf2.t, radius, p1, p2, p3, nullptr, nullptr, nullptr)) {
        if (max_contacts > out.size()) {
          Contact c{};
          c.b1 = id;
          c.b2 = -1;
          out.push_back(c);
        }
      }
    } else {
      Vec3 cp, nrm;
      double pen;
      if (sphere_tri_intersect(tf2.t, radius, p1, p2, p3, &cp, &pen, &nrm)) {
        if (max_contacts > out.size()) {
          Contact c;
          c.b1 = id;
          c.b2 = -1;
          c.pos = cp;
          c.normal = Vec3{{-nrm[0], -nrm[1], -nrm[2]}};
          c.depth = pen;
          out.push_back(c);
        }
      }
    }
  }

  // collisionRecurse (traversal_recurse-inl.h:84-130): the second node is always a leaf, firstOverSecond = true
  void recurse(int b1) {
    const Node& n1 = m1.nodes[b1];
    if (n1.first_child < 0) {
      if (bv_disjoint(b1)) return;
      leaf(b1);
      return;
    }
    if (bv_disjoint(b1)) return;
    recurse(n1.first_child);
    if (can_stop()) return;
    recurse(n1.first_child + 1);
  }
};
}  // namespace

size_t collide_mesh_sphere(const Model& m1, const Pose& tf1, double radius, const Pose& tf2, size_t num_max_contacts,
                           bool enable_contact, std::vector<Contact>& out, CollideStats* stats) {
  if (num_max_contacts == 0) return 0;
  if (!out.empty() && num_max_contacts <= out.size()) return out.size();
  if (m1.nodes.empty()) return out.size();
  MeshSphereCtx ctx{m1, tf1, tf2, radius, Node{}, num_max_contacts, enable_contact, out};
  sphere_obb(radius, tf2, ctx.shape_bv);
  ctx.recurse(0);
  if (stats) {
    stats->n_bv += ctx.n_bv;
    stats->n_leaf += ctx.n_leaf;
  }
  return out.size();
}

// -----------------------------------------------------------------------------
// mesh <-> halfspace / plane: fcl::collide(BVHModel<OBBRSS>, tf1, Halfspace | Plane, tf2) =
// BVHShapeCollider<OBBRSS, Shape> -> orientedBVHShapeCollide (collision_func_matrix-inl.h:378-430, cells :841-842).
// BV test: the shape's OBB against the node's OBB.  computeBV<OBB>(Halfspace) is the infinite box (axis = I, To = 0,
// extent = DBL_MAX, geometry/shape/utility-inl.h:361-373): obbDisjoint can never separate it, every node is visited.
// computeBV<OBB>(Plane) has extent (0, DBL_MAX, DBL_MAX) around the plane (:657-669): the only axis of obbDisjoint that
// can separate is the plane's normal, |T0| > sum_j b_j (|B0j| + 1e-6) (OBB-inl.h:409-412) -- restated directly on the
// transformed plane (its in-plane axes come from Eigen's unitOrthogonal() and cannot matter).
// -----------------------------------------------------------------------------
namespace {
struct MeshPlaneCtx {
  const Model& m1;
  Pose tf1, tf2;
  int kind;  // 0 halfspace, 1 plane
  PlaneShape shape;
  PlaneShape world;  // transform(shape, tf2)
  size_t max_contacts;
  bool enable_contact;
  std::vector<Contact>& out;
  long long n_bv = 0, n_leaf = 0;

  bool can_stop() const { return !out.empty() && max_contacts <= out.size(); }

  bool bv_disjoint(int b1) {
    n_bv++;
    if (kind == 0) return false;
    // the plane's normal axis: T0 = n' . (centre of the node's box in the world) - d', B0j = n' . (world axis j of the box)
    const Node& nd = m1.nodes[b1];
    const Vec3 cw = add(mul(tf1.R, nd.obb_To), tf1.t);
    const double T0 = dot(world.n, cw) - world.d;
    double reach = 0;
    for (int j = 0; j < 3; ++j) {
      const Vec3 aw = mul(tf1.R, col(nd.axis, j));
      reach += nd.obb_ext[j] * (std::fabs(dot(world.n, aw)) + 1e-6);
    }
    return std::fabs(T0) > reach;
  }

  void leaf(int b1) {
    n_leaf++;
    const int id = -(m1.nodes[b1].first_child + 1);
    const Tri& t = m1.tris[id];
    const Vec3 &p1 = m1.verts[t.v[0]], &p2 = m1.verts[t.v[1]], &p3 = m1.verts[t.v[2]];
    Vec3 cp{{0, 0, 0}}, nrm{{0, 0, 0}};
    double pen = 0;
    const bool want = enable_contact;
    const bool hit = kind == 0 ? halfspace_tri_intersect(shape, tf2, p1, p2, p3, tf1, want ? &cp : nullptr, want ? &pen : nullptr, want ? &nrm : nullptr)
                               : plane_tri_intersect(shape, tf2, p1, p2, p3, tf1, want ? &cp : nullptr, want ? &pen : nullptr, want ? &nrm : nullptr);
    if (hit && max_contacts > out.size()) {
      Contact c{};
      c.b1 = id;
      c.b2 = -1;
      if (enable_contact) {
        c.pos = cp;
        c.normal = Vec3{{-nrm[0], -nrm[1], -nrm[2]}};
        c.depth = pen;
      }
      out.push_back(c);
    }
  }

  void recurse(int b1) {
    const Node& n1 = m1.nodes[b1];
    if (bv_disjoint(b1)) return;
    if (n1.first_child < 0) {
      leaf(b1);
      return;
    }
    recurse(n1.first_child);
    if (can_stop()) return;
    recurse(n1.first_child + 1);
  }
};
}  // namespace

size_t collide_mesh_plane(const Model& m1, const Pose& tf1, int kind, const PlaneShape& s, const Pose& tf2, size_t num_max_contacts,
                          bool enable_contact, std::vector<Contact>& out, CollideStats* stats) {
  if (num_max_contacts == 0) return 0;
  if (!out.empty() && num_max_contacts <= out.size()) return out.size();
  if (m1.nodes.empty()) return out.size();
  MeshPlaneCtx ctx{m1, tf1, tf2, kind, s, transform_plane(s, tf2), num_max_contacts, enable_contact, out};
  ctx.recurse(0);
  if (stats) {
    stats->n_bv += ctx.n_bv;
    stats->n_leaf += ctx.n_leaf;
  }
  return out.size();
}

void brute_mesh_plane(const Model& m1, const Pose& tf1, int kind, const PlaneShape& s, const Pose& tf2, std::vector<int>& tris) {
  tris.clear();
  for (int i = 0; i < (int)m1.tris.size(); ++i) {
    const Tri& t = m1.tris[i];
    const bool hit = kind == 0 ? halfspace_tri_intersect(s, tf2, m1.verts[t.v[0]], m1.verts[t.v[1]], m1.verts[t.v[2]], tf1, nullptr, nullptr, nullptr)
                               : plane_tri_intersect(s, tf2, m1.verts[t.v[0]], m1.verts[t.v[1]], m1.verts[t.v[2]], tf1, nullptr, nullptr, nullptr);
    if (hit) tris.push_back(i);
  }
}

void brute_mesh_sphere(const Model& m1, const Pose& tf1, double radius, const Pose& tf2, std::vector<int>& tris) {
  for (int i = 0; i < (int)m1.tris.size(); ++i) {
    const Tri& t = m1.tris[i];
    const Vec3 p1 = add(mul(tf1.R, m1.verts[t.v[0]]), tf1.t);
    const Vec3 p2 = add(mul(tf1.R, m1.verts[t.v[1]]), tf1.t);
    const Vec3 p3 = add(mul(tf1.R, m1.verts[t.v[2]]), tf1.t);
    if (sphere_tri_intersect(tf2.t, radius, p1, p2, p3, nullptr, nullptr, nullptr)) tris.push_back(i);
  }
}

// ---------------------------------------------------------------------------------------
// Mesh <-> sphere distance: BVHShapeDistancer<OBBRSS, Sphere>::distance -> orientedBVHShapeDistance
// (detail/distance_func_matrix-inl.h:259-277, 322-341) -> setupMeshShapeDistanceOrientedNode
// (mesh_shape_distance_traversal_node-inl.h:386-413: computeBV(model2, tf2, model2_bv)) -> detail::distance:
// preprocess (triangle 0), distanceRecurse with a leaf second node, empty postprocess (:351-364).
// Nearest points stay in the LOCAL frames: the solver maps the point on the sphere with tf2^-1 and the point on
// the triangle with tf1^-1 (sphere_triangle-inl.h:485, 499-508).
// Centre within the radius of a triangle: the reference's leaf passes an uninitialised distance to
// DistanceResult::update (the solver returned false without writing it).  The restatement DEFINES that case
// as distance -1 (what the distance-only overload writes, sphere_triangle-inl.h:462) and NaN points.
// ---------------------------------------------------------------------------------------
namespace {
struct MeshSphereDistCtx {
  const Model& m1;
  Pose tf1, tf2;
  double radius;
  Node shape_bv;
  double min_distance = std::numeric_limits<double>::max();
  Vec3 p1{{0, 0, 0}}, p2{{0, 0, 0}};
  int rb1 = -1;
  long long n_bv = 0, n_leaf = 0;

  static Vec3 inverse_apply(const Pose& tf, const Vec3& p) {  // tf.inverse(Isometry) * p
    Vec3 it = mulTv(tf.R, tf.t);
    it = Vec3{{-it[0], -it[1], -it[2]}};
    return add(mulTv(tf.R, p), it);
  }

  // meshShapeDistanceOrientedNodeLeafTesting / distancePreprocessOrientedNode (:163-236)
  void tri(int id) {
    const Tri& t = m1.tris[id];
    const Vec3 P1 = add(mul(tf1.R, m1.verts[t.v[0]]), tf1.t);
    const Vec3 P2 = add(mul(tf1.R, m1.verts[t.v[1]]), tf1.t);
    const Vec3 P3 = add(mul(tf1.R, m1.verts[t.v[2]]), tf1.t);
    double d;
    Vec3 on_sphere, on_tri;
    if (sphere_tri_distance(tf2.t, radius, P1, P2, P3, &d, &on_sphere, &on_tri)) {
      if (min_distance > d) {  // DistanceResult::update, distance_result-inl.h:66-103
        min_distance = d;
        rb1 = id;
        p1 = inverse_apply(tf1, on_tri);
        p2 = inverse_apply(tf2, on_sphere);
      }
    } else if (min_distance > -1) {
      const double nan = std::numeric_limits<double>::quiet_NaN();
      min_distance = -1;
      rb1 = id;
      p1 = p2 = Vec3{{nan, nan, nan}};
    }
  }

  bool can_stop(double c) const { return c >= min_distance; }  // :100-106 with rel_err = abs_err = 0

  // MeshShapeDistanceTraversalNodeOBBRSS::BVTesting (:366-375)
  double bv(int b1) {
    n_bv++;
    return rss_distance(tf1.R, tf1.t, shape_bv, m1.nodes[b1]);
  }

  // distanceRecurse (traversal_recurse-inl.h:259-316); the second node is always a leaf, firstOverSecond = true
  void recurse(int b1) {
    const Node& n1 = m1.nodes[b1];
    if (n1.first_child < 0) {
      n_leaf++;
      tri(-(n1.first_child + 1));
      return;
    }
    const int a1 = n1.first_child, c1 = n1.first_child + 1;
    const double d1 = bv(a1);
    const double d2 = bv(c1);
    if (d2 < d1) {
      if (!can_stop(d2)) recurse(c1);
      if (!can_stop(d1)) recurse(a1);
    } else {
      if (!can_stop(d1)) recurse(a1);
      if (!can_stop(d2)) recurse(c1);
    }
  }
};
}  // namespace

double distance_mesh_sphere(const Model& m1, const Pose& tf1, double radius, const Pose& tf2, DistanceOut& out,
                            CollideStats* stats) {
  MeshSphereDistCtx ctx{m1, tf1, tf2, radius, Node{}};
  sphere_obb(radius, tf2, ctx.shape_bv);
  ctx.tri(0);  // preprocess
  ctx.recurse(0);
  out.min_distance = ctx.min_distance;
  out.b1 = ctx.rb1;
  out.b2 = -1;  // DistanceResult::NONE
  out.p1 = ctx.p1;
  out.p2 = ctx.p2;
  if (stats) {
    stats->n_bv += ctx.n_bv;
    stats->n_leaf += ctx.n_leaf;
  }
  return out.min_distance;
}

double brute_distance_mesh_sphere(const Model& m1, const Pose& tf1, double radius, const Pose& tf2, DistanceOut& out) {
  MeshSphereDistCtx ctx{m1, tf1, tf2, radius, Node{}};
  for (int i = 0; i < (int)m1.tris.size(); ++i) ctx.tri(i);
  out.min_distance = ctx.min_distance;
  out.b1 = ctx.rb1;
  out.b2 = -1;
  out.p1 = ctx.p1;
  out.p2 = ctx.p2;
  return out.min_distance;
}

// -----------------------------------------------------------------------------
// Broadphase: local / world AABBs and the brute-force manager's pair enumeration
// -----------------------------------------------------------------------------
LocalAABB local_aabb(const Model& m) {
  LocalAABB a;
  const double big = std::numeric_limits<double>::max();
  a.mn = Vec3{{big, big, big}};
  a.mx = Vec3{{-big, -big, -big}};
  for (const Tri& t : m.tris)
    for (int c = 0; c < 3; ++c)
      for (int k = 0; k < 3; ++k) {
        a.mn[k] = std::min(a.mn[k], m.verts[t.v[c]][k]);
        a.mx[k] = std::max(a.mx[k], m.verts[t.v[c]][k]);
      }
  a.center = scale(add(a.mn, a.mx), 0.5);  // AABB::center()
  double r2 = 0;
  for (const Tri& t : m.tris)
    for (int c = 0; c < 3; ++c) {
      const double r = sqnorm(sub(a.center, m.verts[t.v[c]]));
      if (r > r2) r2 = r;
    }
  a.radius = std::sqrt(r2);
  return a;
}

// Eigen's MatrixBase::isIdentity(prec = NumTraits<double>::dummy_precision() = 1e-12): diagonal isApprox(x, 1, prec),
// off-diagonal isMuchSmallerThan(x, 1, prec)  (Eigen/src/Core/CwiseNullaryOp.h; third-party, version unpinned)
static bool is_identity(const Mat3& R) {
  const double prec = 1e-12;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      const double x = R.m[i][j];
      if (i == j) {
        const double ax = std::fabs(x), one = 1.0;
        if (!(std::fabs(x - one) <= (ax < one ? ax : one) * prec)) return false;
      } else {
        if (!(std::fabs(x) <= prec)) return false;
      }
    }
  return true;
}

void world_aabb(const LocalAABB& a, const Pose& tf, double out6[6]) {
  if (is_identity(tf.R)) {
    for (int k = 0; k < 3; ++k) {
      out6[k] = a.mn[k] + tf.t[k];
      out6[3 + k] = a.mx[k] + tf.t[k];
    }
  } else {
    const Vec3 c = add(mul(tf.R, a.center), tf.t);
    for (int k = 0; k < 3; ++k) {
      out6[k] = c[k] - a.radius;
      out6[3 + k] = c[k] + a.radius;
    }
  }
}

void broadphase_pairs(const std::vector<const Model*>& geoms, const std::vector<int>& geom1, const std::vector<Pose>& tf1,
                      const std::vector<int>& geom2, const std::vector<Pose>& tf2, std::vector<std::pair<int, int>>& pairs) {
  std::vector<LocalAABB> loc;
  for (const Model* m : geoms) loc.push_back(local_aabb(*m));
  std::vector<double> a1(6 * geom1.size()), a2(6 * geom2.size());
  for (size_t i = 0; i < geom1.size(); ++i) world_aabb(loc[geom1[i]], tf1[i], &a1[6 * i]);
  for (size_t j = 0; j < geom2.size(); ++j) world_aabb(loc[geom2[j]], tf2[j], &a2[6 * j]);
  pairs.clear();
  for (size_t i = 0; i < geom1.size(); ++i)
    for (size_t j = 0; j < geom2.size(); ++j) {
      const double *a = &a1[6 * i], *b = &a2[6 * j];
      if (a[0] > b[3] || a[1] > b[4] || a[2] > b[5]) continue;
      if (a[3] < b[0] || a[4] < b[1] || a[5] < b[2]) continue;
      pairs.emplace_back((int)i, (int)j);
    }
}

}  // namespace oracle
