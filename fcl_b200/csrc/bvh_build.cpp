// bvh_build.cpp — host-side OBBRSS BVH construction for the upload step.
//
// Mirrors what BVHModel<OBBRSS<double>>::endModel() produces so that a model built here and a
// model built by the reference have the same tree and the same bounding volumes
// (reference files, /root/reference/include/fcl/...):
//   geometry/bvh/BVH_model-inl.h:450-517, 833-938   endModel / buildTree / recursiveBuildTree
//   geometry/bvh/detail/BV_fitter-inl.h:449-477     FitImpl<OBBRSS>
//   geometry/bvh/detail/BV_splitter-inl.h:422-451, 492-500, 540-657  split rules
//   math/geometry-inl.h:294-362, 477-597, 709-988, 1335-1425  extent/centre, Jacobi, RSS fit, covariance
// Implementation notes: works on de-indexed triangles (9 doubles each), iterative with an
// explicit job stack (pre-order numbering: a node's two children get consecutive ids when the
// node is processed; the left subtree is finished before the right one starts).
// Arithmetic: separately rounded mul/add, three-term sums left to right (built with
// -ffp-contract=off), like the device code.
#include "bvh_build.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

#include "../../include/fclgpu.h"

namespace {

struct Fit {
  double axis[9];  // row-major, columns = box axes
  double obb_To[3], obb_ext[3], rss_To[3], rss_l[2], rss_r;
};

inline double dot_col(const double* A, int c, const double* p) {  // axis.col(c) . p
  return (A[c] * p[0] + A[3 + c] * p[1]) + A[6 + c] * p[2];
}

// symmetric 3x3 eigen-decomposition, cyclic Jacobi (math/geometry-inl.h:477-558);
// v[r][k]: k-th eigenvector in column k
void jacobi3(const double M[3][3], double d[3], double v[3][3]) {
  double R[3][3];
  std::memcpy(R, M, sizeof R);
  double b[3], z[3];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) v[i][j] = (i == j);
    b[i] = d[i] = R[i][i];
    z[i] = 0;
  }
  for (int sweep = 0; sweep < 50; ++sweep) {
    double sm = 0;
    sm += std::abs(R[0][1]);
    sm += std::abs(R[0][2]);
    sm += std::abs(R[1][2]);
    if (sm == 0.0) return;
    const double tresh = (sweep < 3) ? 0.2 * sm / 9 : 0.0;
    for (int ip = 0; ip < 3; ++ip)
      for (int iq = ip + 1; iq < 3; ++iq) {
        double g = 100.0 * std::abs(R[ip][iq]);
        if (sweep > 3 && std::abs(d[ip]) + g == std::abs(d[ip]) && std::abs(d[iq]) + g == std::abs(d[iq])) {
          R[ip][iq] = 0.0;
          continue;
        }
        if (!(std::abs(R[ip][iq]) > tresh)) continue;
        double h = d[iq] - d[ip], t;
        if (std::abs(h) + g == std::abs(h)) {
          t = R[ip][iq] / h;
        } else {
          const double theta = 0.5 * h / R[ip][iq];
          t = 1.0 / (std::abs(theta) + std::sqrt(1.0 + theta * theta));
          if (theta < 0.0) t = -t;
        }
        const double c = 1.0 / std::sqrt(1 + t * t), s = t * c, tau = s / (1.0 + c);
        h = t * R[ip][iq];
        z[ip] -= h;
        z[iq] += h;
        d[ip] -= h;
        d[iq] += h;
        R[ip][iq] = 0.0;
        auto givens = [s, tau](double& x, double& y) {
          const double gx = x, hy = y;
          x = gx - s * (hy + gx * tau);
          y = hy + s * (gx - hy * tau);
        };
        for (int j = 0; j < ip; ++j) givens(R[j][ip], R[j][iq]);
        for (int j = ip + 1; j < iq; ++j) givens(R[ip][j], R[j][iq]);
        for (int j = iq + 1; j < 3; ++j) givens(R[ip][j], R[iq][j]);
        for (int j = 0; j < 3; ++j) givens(v[j][ip], v[j][iq]);
      }
    for (int i = 0; i < 3; ++i) {
      b[i] += z[i];
      d[i] = b[i];
      z[i] = 0.0;
    }
  }
}

struct Builder {
  const double* tv;  // 9 doubles per triangle
  int split;
  std::vector<uint32_t> order;
  std::vector<double> proj;  // scratch: projected points, 3 doubles each
  std::vector<double> med;

  const double* vert(uint32_t tri, int k) const { return tv + 9 * (size_t)tri + 3 * k; }

  void fit(const uint32_t* idx, int n, Fit& f) {
    // --- covariance of the 3n vertices (math/geometry-inl.h:1335-1425) ---
    double S1[3] = {0, 0, 0}, S2[6] = {0, 0, 0, 0, 0, 0};  // xx yy zz xy xz yz
    for (int i = 0; i < n; ++i) {
      const double *p1 = vert(idx[i], 0), *p2 = vert(idx[i], 1), *p3 = vert(idx[i], 2);
      for (int k = 0; k < 3; ++k) S1[k] += ((p1[k] + p2[k]) + p3[k]);
      S2[0] += (p1[0] * p1[0] + p2[0] * p2[0] + p3[0] * p3[0]);
      S2[1] += (p1[1] * p1[1] + p2[1] * p2[1] + p3[1] * p3[1]);
      S2[2] += (p1[2] * p1[2] + p2[2] * p2[2] + p3[2] * p3[2]);
      S2[3] += (p1[0] * p1[1] + p2[0] * p2[1] + p3[0] * p3[1]);
      S2[4] += (p1[0] * p1[2] + p2[0] * p2[2] + p3[0] * p3[2]);
      S2[5] += (p1[1] * p1[2] + p2[1] * p2[2] + p3[1] * p3[2]);
    }
    const int np = 3 * n;
    double M[3][3];
    M[0][0] = S2[0] - S1[0] * S1[0] / np;
    M[1][1] = S2[1] - S1[1] * S1[1] / np;
    M[2][2] = S2[2] - S1[2] * S1[2] / np;
    M[0][1] = M[1][0] = S2[3] - S1[0] * S1[1] / np;
    M[1][2] = M[2][1] = S2[5] - S1[1] * S1[2] / np;
    M[0][2] = M[2][0] = S2[4] - S1[0] * S1[2] / np;

    // --- principal axes: largest, middle eigenvector, then their cross product (:563-597) ---
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
    A[2] = A[3] * A[7] - A[6] * A[4];   // col0 x col1, component x = a1*b2 - a2*b1
    A[5] = A[6] * A[1] - A[0] * A[7];   // y = a2*b0 - a0*b2
    A[8] = A[0] * A[4] - A[3] * A[1];   // z = a0*b1 - a1*b0

    // --- project every vertex on the axes once (used by both fits) ---
    proj.resize(9 * (size_t)n);
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < 3; ++k) {
        const double* p = vert(idx[i], k);
        double* o = &proj[9 * (size_t)i + 3 * k];
        o[0] = dot_col(A, 0, p);
        o[1] = dot_col(A, 1, p);
        o[2] = dot_col(A, 2, p);
      }
    const int m = 3 * n;
    const double* P = proj.data();

    // --- OBB centre and half extents (math/geometry-inl.h:294-362) ---
    const double big = std::numeric_limits<double>::max();
    double mn[3] = {big, big, big}, mx[3] = {-big, -big, -big};
    for (int i = 0; i < m; ++i)
      for (int k = 0; k < 3; ++k) {
        const double c = P[3 * i + k];
        if (c > mx[k]) mx[k] = c;
        if (c < mn[k]) mn[k] = c;
      }
    const double o[3] = {(mx[0] + mn[0]) / 2, (mx[1] + mn[1]) / 2, (mx[2] + mn[2]) / 2};
    for (int r = 0; r < 3; ++r) {
      f.obb_To[r] = (A[3 * r] * o[0] + A[3 * r + 1] * o[1]) + A[3 * r + 2] * o[2];
      f.obb_ext[r] = (mx[r] - mn[r]) / 2;
    }

    // --- RSS: radius from the thin direction, rectangle grown to cover (:709-988) ---
    double minz = P[2], maxz = P[2];
    for (int i = 1; i < m; ++i) {
      const double zv = P[3 * i + 2];
      if (zv < minz) minz = zv;
      else if (zv > maxz) maxz = zv;
    }
    const double r = 0.5 * (maxz - minz), radsqr = r * r, cz = 0.5 * (maxz + minz);
    auto reach = [&](int i) {  // half chord of the sphere-swept slab at this point's height
      const double dz = P[3 * i + 2] - cz;
      return std::sqrt(std::max<double>(radsqr - dz * dz, 0));
    };
    double lo2[2], hi2[2];
    for (int c = 0; c < 2; ++c) {
      int imin = 0, imax = 0;
      double vmin = P[c], vmax = P[c];
      for (int i = 1; i < m; ++i) {
        const double val = P[3 * i + c];
        if (val < vmin) { imin = i; vmin = val; }
        else if (val > vmax) { imax = i; vmax = val; }
      }
      double lo_c = P[3 * imin + c] + reach(imin);
      double hi_c = P[3 * imax + c] - reach(imax);
      for (int i = 0; i < m; ++i)
        if (P[3 * i + c] < lo_c) {
          const double x = P[3 * i + c] + reach(i);
          if (x < lo_c) lo_c = x;
        }
      for (int i = 0; i < m; ++i)
        if (P[3 * i + c] > hi_c) {
          const double x = P[3 * i + c] - reach(i);
          if (x > hi_c) hi_c = x;
        }
      lo2[c] = lo_c;
      hi2[c] = hi_c;
    }
    double minx = lo2[0], maxx = hi2[0], miny = lo2[1], maxy = hi2[1];
    const double a = std::sqrt(0.5);
    for (int i = 0; i < m; ++i) {
      const double px = P[3 * i], py = P[3 * i + 1], pz = P[3 * i + 2];
      double dx, dy, u, t;
      if (px > maxx) {
        if (py > maxy) {
          dx = px - maxx; dy = py - maxy;
          u = dx * a + dy * a;
          t = (a * u - dx) * (a * u - dx) + (a * u - dy) * (a * u - dy) + (cz - pz) * (cz - pz);
          u = u - std::sqrt(std::max<double>(radsqr - t, 0));
          if (u > 0) { maxx += u * a; maxy += u * a; }
        } else if (py < miny) {
          dx = px - maxx; dy = py - miny;
          u = dx * a - dy * a;
          t = (a * u - dx) * (a * u - dx) + (-a * u - dy) * (-a * u - dy) + (cz - pz) * (cz - pz);
          u = u - std::sqrt(std::max<double>(radsqr - t, 0));
          if (u > 0) { maxx += u * a; miny -= u * a; }
        }
      } else if (px < minx) {
        if (py > maxy) {
          dx = px - minx; dy = py - maxy;
          u = dy * a - dx * a;
          t = (-a * u - dx) * (-a * u - dx) + (a * u - dy) * (a * u - dy) + (cz - pz) * (cz - pz);
          u = u - std::sqrt(std::max<double>(radsqr - t, 0));
          if (u > 0) { minx -= u * a; maxy += u * a; }
        } else if (py < miny) {
          dx = px - minx; dy = py - miny;
          u = -dx * a - dy * a;
          t = (-a * u - dx) * (-a * u - dx) + (-a * u - dy) * (-a * u - dy) + (cz - pz) * (cz - pz);
          u = u - std::sqrt(std::max<double>(radsqr - t, 0));
          if (u > 0) { minx -= u * a; miny -= u * a; }
        }
      }
    }
    for (int k = 0; k < 3; ++k) f.rss_To[k] = (A[3 * k] * minx + A[3 * k + 1] * miny) + A[3 * k + 2] * cz;
    f.rss_l[0] = maxx - minx;
    if (f.rss_l[0] < 0) f.rss_l[0] = 0;
    f.rss_l[1] = maxy - miny;
    if (f.rss_l[1] < 0) f.rss_l[1] = 0;
    f.rss_r = r;
  }

  // split threshold along obb.axis.col(0) (BV_splitter-inl.h:540-657)
  double threshold(const Fit& f, const uint32_t* idx, int n) {
    const double sv[3] = {f.axis[0], f.axis[3], f.axis[6]};
    if (split == FCLGPU_SPLIT_METHOD_BV_CENTER) return f.obb_To[0];
    if (split == FCLGPU_SPLIT_METHOD_MEDIAN) {
      med.resize(n);
      for (int i = 0; i < n; ++i) {
        const double *p1 = vert(idx[i], 0), *p2 = vert(idx[i], 1), *p3 = vert(idx[i], 2);
        const double c0 = p1[0] + p2[0] + p3[0], c1 = p1[1] + p2[1] + p3[1], c2 = p1[2] + p2[2] + p3[2];
        med[i] = ((c0 * sv[0] + c1 * sv[1]) + c2 * sv[2]) / 3;
      }
      std::sort(med.begin(), med.end());
      return (n % 2 == 1) ? med[(n - 1) / 2] : (med[n / 2] + med[n / 2 - 1]) / 2;
    }
    double c[3] = {0.0, 0.0, 0.0};
    for (int i = 0; i < n; ++i) {
      const double *p1 = vert(idx[i], 0), *p2 = vert(idx[i], 1), *p3 = vert(idx[i], 2);
      c[0] += (p1[0] + p2[0] + p3[0]);
      c[1] += (p1[1] + p2[1] + p3[1]);
      c[2] += (p1[2] + p2[2] + p3[2]);
    }
    return (c[0] * sv[0] + c[1] * sv[1] + c[2] * sv[2]) / (3 * n);
  }
};

}  // namespace

extern "C" int fclgpu_bvh_build_obbrss(const double* vertices, int32_t num_vertices, const int32_t* triangles,
                                       int32_t num_tris, int32_t split_method, fclgpu_bvh** out) {
  if (!out) return FCLGPU_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  if (num_tris <= 0 || num_vertices <= 0) return FCLGPU_ERR_BUILD_EMPTY_MODEL;  // BVH_model-inl.h:459-463
  if (!vertices || !triangles) return FCLGPU_ERR_INVALID_ARGUMENT;
  if (split_method < 0 || split_method > 2) return FCLGPU_ERR_UNSUPPORTED_FUNCTION;
  for (int64_t i = 0; i < 3 * (int64_t)num_tris; ++i)
    if (triangles[i] < 0 || triangles[i] >= num_vertices) return FCLGPU_ERR_INCORRECT_DATA;

  fclgpu_bvh* b = new fclgpu_bvh;
  b->num_tris = num_tris;
  const int nn = 2 * num_tris - 1;
  b->first_child.assign(nn, 0);
  b->axis.assign(9 * (size_t)nn, 0);
  b->obb_To.assign(3 * (size_t)nn, 0);
  b->obb_ext.assign(3 * (size_t)nn, 0);
  b->rss_To.assign(3 * (size_t)nn, 0);
  b->rss_l.assign(2 * (size_t)nn, 0);
  b->rss_r.assign(nn, 0);
  b->tri_verts.resize(9 * (size_t)num_tris);
  for (int t = 0; t < num_tris; ++t)
    for (int k = 0; k < 3; ++k)
      for (int c = 0; c < 3; ++c) b->tri_verts[9 * (size_t)t + 3 * k + c] = vertices[3 * (size_t)triangles[3 * t + k] + c];

  Builder B;
  B.tv = b->tri_verts.data();
  B.split = split_method;
  B.order.resize(num_tris);
  for (int i = 0; i < num_tris; ++i) B.order[i] = i;

  struct Job {
    int node, first, count;
  };
  std::vector<Job> jobs;
  jobs.push_back({0, 0, num_tris});
  int next_id = 1;
  while (!jobs.empty()) {
    const Job j = jobs.back();
    jobs.pop_back();
    uint32_t* idx = B.order.data() + j.first;
    Fit f;
    B.fit(idx, j.count, f);
    std::memcpy(&b->axis[9 * (size_t)j.node], f.axis, sizeof f.axis);
    for (int k = 0; k < 3; ++k) {
      b->obb_To[3 * (size_t)j.node + k] = f.obb_To[k];
      b->obb_ext[3 * (size_t)j.node + k] = f.obb_ext[k];
      b->rss_To[3 * (size_t)j.node + k] = f.rss_To[k];
    }
    b->rss_l[2 * (size_t)j.node] = f.rss_l[0];
    b->rss_l[2 * (size_t)j.node + 1] = f.rss_l[1];
    b->rss_r[j.node] = f.rss_r;
    if (j.count == 1) {
      b->first_child[j.node] = -((int32_t)idx[0] + 1);
      continue;
    }
    const double thr = B.threshold(f, idx, j.count);
    const double sv[3] = {f.axis[0], f.axis[3], f.axis[6]};
    const int left = next_id;
    next_id += 2;
    b->first_child[j.node] = left;
    // partition: centroids with sv.c <= thr go to the front, order otherwise preserved by swaps
    int c1 = 0;
    for (int i = 0; i < j.count; ++i) {
      const double *p1 = B.vert(idx[i], 0), *p2 = B.vert(idx[i], 1), *p3 = B.vert(idx[i], 2);
      const double cx = ((p1[0] + p2[0]) + p3[0]) / 3.0, cy = ((p1[1] + p2[1]) + p3[1]) / 3.0,
                   cz = ((p1[2] + p2[2]) + p3[2]) / 3.0;
      if (!(((sv[0] * cx + sv[1] * cy) + sv[2] * cz) > thr)) {
        std::swap(idx[i], idx[c1]);
        c1++;
      }
    }
    if (c1 == 0 || c1 == j.count) c1 = j.count / 2;
    jobs.push_back({left + 1, j.first + c1, j.count - c1});  // right: processed after the whole left subtree
    jobs.push_back({left, j.first, c1});
  }
  *out = b;
  return FCLGPU_OK;
}

extern "C" void fclgpu_bvh_destroy(fclgpu_bvh* bvh) { delete bvh; }
extern "C" int32_t fclgpu_bvh_num_nodes(const fclgpu_bvh* bvh) { return bvh ? 2 * bvh->num_tris - 1 : 0; }
extern "C" int32_t fclgpu_bvh_num_tris(const fclgpu_bvh* bvh) { return bvh ? bvh->num_tris : 0; }

extern "C" int fclgpu_bvh_get(const fclgpu_bvh* b, int32_t* first_child, double* axis9, double* obb_To3,
                              double* obb_extent3, double* rss_To3, double* rss_l2, double* rss_r,
                              double* tri_verts9) {
  if (!b) return FCLGPU_ERR_INVALID_ARGUMENT;
  auto cp = [](auto* dst, const auto& src) {
    if (dst) std::memcpy(dst, src.data(), src.size() * sizeof(src[0]));
  };
  cp(first_child, b->first_child);
  cp(axis9, b->axis);
  cp(obb_To3, b->obb_To);
  cp(obb_extent3, b->obb_ext);
  cp(rss_To3, b->rss_To);
  cp(rss_l2, b->rss_l);
  cp(rss_r, b->rss_r);
  cp(tri_verts9, b->tri_verts);
  return FCLGPU_OK;
}
