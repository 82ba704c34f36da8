// bounds_f32.cuh — conservative single-precision BV tests that only STEER the traversal.
//
// The reference prunes the BVTT with double-precision RSS distances / OBB overlap tests
// (math/bv/RSS-inl.h:1957-1974, math/bv/OBB-inl.h:384-523).  A pruning test does not have to
// reproduce those numbers: any test that never discards a node pair whose subtree could matter
// gives the same result, because results are produced only by the exact FP64 leaf routines
// (triDistance / intersect_Triangle):
//   * distance: a VALID LOWER BOUND lb <= dist(RSS1, RSS2) prunes (lb >= current minimum) only
//     pairs that cannot improve the minimum, so the final minimum is the minimum over all
//     triangle pairs, which is the reference's value;
//   * collide: a test that answers "disjoint" only when the two boxes are certainly disjoint
//     visits a superset of the reference's BVTT nodes in the same depth-first order, so the
//     same triangle pairs intersect in the same order.
// The bounds below are evaluated in FP32 (the FP32 pipe is otherwise idle on this path and
// issues twice as fast as FP64), are branch-free so all 32 lanes stay converged, and subtract
// a slack that dominates the worst-case rounding error (derivation in DESIGN.md §4.5):
//   every intermediate has magnitude <= M = s1 + s2 + |T0|_1 (s = |To|_1 + l0 + l1 + r per
//   node), inputs are rounded once (relative 2^-24), accumulated absolute error of every
//   candidate gap <= 1000 u M with u = 2^-24; slack = 1536 u M plus 0.1 % of the bound.
#pragma once
#include <cuda_runtime.h>

#include <cmath>

namespace fclgpu {

#ifndef FD
#define FD __host__ __device__ __forceinline__
#endif

// 64-byte single-precision RSS record: axis[9] row-major, rectangle CENTRE c[3] (not the
// corner), half side lengths h[2] (rounded up), radius r (rounded up), magnitude scale s
// (rounded up).
struct RssRec32 {
  float a[9];
  float c[3];
  float h0, h1, r, s;
};

FD float f32_rsqrt(float x) {
#ifdef __CUDA_ARCH__
  return rsqrtf(x);
#else
  return 1.0f / sqrtf(x);
#endif
}
// approximate reciprocal (MUFU.RCP): only used where the result's accuracy cannot invalidate a bound
FD float f32_rcp(float x) {
#ifdef __CUDA_ARCH__
  return __fdividef(1.0f, x);
#else
  return 1.0f / x;
#endif
}

// lower bound on the distance between two RSS; R0 (row-major), T0: pose of model2 in model1's
// frame, t0_l1 >= |T0|_1.  Valid for any inputs: every candidate is the support-function gap
// along some direction, which never exceeds the true distance between the rectangles.
FD float rss_lower_bound_f32(const float* R0, const float* T0, float t0_l1, const RssRec32& n1, const RssRec32& n2) {
  // B = R0 * A2 (columns = B's axes in model1's frame), cB = T0 + R0 * c2
  float B[9], D[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
      B[3 * r + c] = fmaf(R0[3 * r + 2], n2.a[6 + c], fmaf(R0[3 * r + 1], n2.a[3 + c], R0[3 * r] * n2.a[c]));
    const float cb = fmaf(R0[3 * r + 2], n2.c[2], fmaf(R0[3 * r + 1], n2.c[1], fmaf(R0[3 * r], n2.c[0], T0[r])));
    D[r] = cb - n1.c[r];
  }
  // C[i][j] = a_i . b_j ; DA = A1^T D ; DB = B^T D
  float C[9], DA[3], DB[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) C[3 * i + j] = fmaf(n1.a[6 + i], B[6 + j], fmaf(n1.a[3 + i], B[3 + j], n1.a[i] * B[j]));
    DA[i] = fmaf(n1.a[6 + i], D[2], fmaf(n1.a[3 + i], D[1], n1.a[i] * D[0]));
    DB[i] = fmaf(B[6 + i], D[2], fmaf(B[3 + i], D[1], B[i] * D[0]));
  }
  const float hA[3] = {n1.h0, n1.h1, 0.0f}, hB[3] = {n2.h0, n2.h1, 0.0f};
  float best = 0.0f;
  // A's three axes, B's three axes
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float ga = fabsf(DA[i]) - hA[i] - fmaf(hB[1], fabsf(C[3 * i + 1]), hB[0] * fabsf(C[3 * i]));
    const float gb = fabsf(DB[i]) - hB[i] - fmaf(hA[1], fabsf(C[3 + i]), hA[0] * fabsf(C[i]));
    best = fmaxf(best, fmaxf(ga, gb));
  }
  // nine edge-direction cross products a_i x b_j (skipped when nearly parallel: |a_i x b_j|^2 < 1/16)
// Measured on env/rob (1M poses): all nine 34.6+2.6 ms / 837 box bounds per query; the four products of the
// in-plane axes (i, j < 2) 34.6 ms / 844; one 34.3 / 872; none 35.6 / 917 -- the products that involve a
// rectangle normal never decide a bound here, so they are left out (any subset is a valid bound).
#ifndef FCLGPU_RSS_CROSS
#define FCLGPU_RSS_CROSS 2
#endif
#pragma unroll
  for (int i = 0; i < FCLGPU_RSS_CROSS; ++i) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
#ifndef FCLGPU_RSS_CROSS_J
#define FCLGPU_RSS_CROSS_J 2
#endif
#pragma unroll
    for (int j = 0; j < FCLGPU_RSS_CROSS_J; ++j) {
      const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      const float len2 = fmaf(-C[3 * i + j], C[3 * i + j], 1.0f);
      const float num = fabsf(fmaf(DA[i2], C[3 * i1 + j], -(DA[i1] * C[3 * i2 + j])));
      const float ra = fmaf(hA[i2], fabsf(C[3 * i1 + j]), hA[i1] * fabsf(C[3 * i2 + j]));
      const float rb = fmaf(hB[j2], fabsf(C[3 * i + j1]), hB[j1] * fabsf(C[3 * i + j2]));
      const float g = (num - ra - rb) * f32_rsqrt(fmaxf(len2, 0.0625f));
      best = fmaxf(best, len2 >= 0.0625f ? g : 0.0f);
    }
  }
  const float M = n1.s + n2.s + t0_l1;
  // centre-to-centre direction (skipped when the centres are close relative to the scale)
  {
    const float len2 = fmaf(D[2], D[2], fmaf(D[1], D[1], D[0] * D[0]));
    const float ra = fmaf(hA[1], fabsf(DA[1]), hA[0] * fabsf(DA[0]));
    const float rb = fmaf(hB[1], fabsf(DB[1]), hB[0] * fabsf(DB[0]));
    const float lim = M * (1.0f / 128.0f);
    const bool ok = len2 >= lim * lim;
    const float g = (len2 - ra - rb) * f32_rsqrt(ok ? len2 : 1.0f);
    best = fmaxf(best, ok ? g : 0.0f);
  }
  // one refined direction: between the point of rectangle A nearest to B's centre and the point of
  // rectangle B nearest to A's centre (any direction gives a valid bound; this one is usually tight)
  {
    const float pa0 = fminf(fmaxf(DA[0], -hA[0]), hA[0]), pa1 = fminf(fmaxf(DA[1], -hA[1]), hA[1]);
    const float pb0 = fminf(fmaxf(-DB[0], -hB[0]), hB[0]), pb1 = fminf(fmaxf(-DB[1], -hB[1]), hB[1]);
    float L[3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
      L[r] = fmaf(-pa1, n1.a[3 * r + 1], fmaf(-pa0, n1.a[3 * r], fmaf(pb1, B[3 * r + 1], fmaf(pb0, B[3 * r], D[r]))));
    const float len2 = fmaf(L[2], L[2], fmaf(L[1], L[1], L[0] * L[0]));
    const float dl = fmaf(L[2], D[2], fmaf(L[1], D[1], L[0] * D[0]));
    float ra = 0.0f, rb = 0.0f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float la = fmaf(L[2], n1.a[6 + i], fmaf(L[1], n1.a[3 + i], L[0] * n1.a[i]));
      const float lb = fmaf(L[2], B[6 + i], fmaf(L[1], B[3 + i], L[0] * B[i]));
      ra = fmaf(hA[i], fabsf(la), ra);
      rb = fmaf(hB[i], fabsf(lb), rb);
    }
    const float lim = M * (1.0f / 128.0f);
    const bool ok = len2 >= lim * lim;
    const float g = (fabsf(dl) - ra - rb) * f32_rsqrt(ok ? len2 : 1.0f);
    best = fmaxf(best, ok ? g : 0.0f);
  }
  const float slack = 9.1552734375e-05f * M;  // 1536 * 2^-24 * M
#ifndef FCLGPU_RSS_REL
#define FCLGPU_RSS_REL 0.999f
#endif
  const float lb = fmaf(best, FCLGPU_RSS_REL, -slack) - (n1.r + n2.r) * 1.000001f;
  return fmaxf(lb, 0.0f);
}


// 64-byte single-precision OBB record: axis[9] row-major, centre c[3], half extents e[3]
// (rounded up), magnitude scale s = |c|_1 + |e|_1 (rounded up).
struct ObbRec32 {
  float a[9];
  float c[3];
  float e[3];
  float s;
};

// true only when the two boxes are CERTAINLY disjoint: some axis of the 15-axis SAT separates
// them by more than the worst-case single-precision rounding error (256 u M, unnormalised
// axes; see the error budget at the top of this file).  "false" means "maybe overlapping":
// the traversal then descends, which can only add work, never change a result.
FD bool obb_certainly_disjoint_f32(const float* R0, const float* T0, float t0_l1, const ObbRec32& n1, const ObbRec32& n2) {
  float B[9], D[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
      B[3 * r + c] = fmaf(R0[3 * r + 2], n2.a[6 + c], fmaf(R0[3 * r + 1], n2.a[3 + c], R0[3 * r] * n2.a[c]));
    const float cb = fmaf(R0[3 * r + 2], n2.c[2], fmaf(R0[3 * r + 1], n2.c[1], fmaf(R0[3 * r], n2.c[0], T0[r])));
    D[r] = cb - n1.c[r];
  }
  float C[9], DA[3], DB[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j)
      C[3 * i + j] = fabsf(fmaf(n1.a[6 + i], B[6 + j], fmaf(n1.a[3 + i], B[3 + j], n1.a[i] * B[j])));
    DA[i] = fmaf(n1.a[6 + i], D[2], fmaf(n1.a[3 + i], D[1], n1.a[i] * D[0]));
    DB[i] = fmaf(B[6 + i], D[2], fmaf(B[3 + i], D[1], B[i] * D[0]));
  }
  // signed products for the cross axes need the signed C: recompute sign-carrying terms from A1, B
  float best = -3.0e38f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float ga = fabsf(DA[i]) - n1.e[i] - fmaf(n2.e[2], C[3 * i + 2], fmaf(n2.e[1], C[3 * i + 1], n2.e[0] * C[3 * i]));
    const float gb = fabsf(DB[i]) - n2.e[i] - fmaf(n1.e[2], C[6 + i], fmaf(n1.e[1], C[3 + i], n1.e[0] * C[i]));
    best = fmaxf(best, fmaxf(ga, gb));
  }
// Cross-product axes a_i x b_j, i < FCLGPU_OBB_CROSS.  A missing axis can only turn "certainly disjoint" into
// "maybe overlapping" (more traversal work, same results).  Measured (1M poses env/rob, verdicts / contacts):
// all nine 3.87 / 24.9 ms; the three with box 1's major axis (i = 0) 3.76 / 23.9 ms; six 3.80 / 24.3; none 4.97 / 29.4.
#ifndef FCLGPU_OBB_CROSS
#define FCLGPU_OBB_CROSS 1
#endif
#pragma unroll
  for (int i = 0; i < FCLGPU_OBB_CROSS; ++i) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      // D . (a_i x b_j) = (D x a_i) . b_j, evaluated directly (no signed C needed)
      const float ux = fmaf(D[1], n1.a[6 + i], -(D[2] * n1.a[3 + i]));
      const float uy = fmaf(D[2], n1.a[i], -(D[0] * n1.a[6 + i]));
      const float uz = fmaf(D[0], n1.a[3 + i], -(D[1] * n1.a[i]));
      const float num = fabsf(fmaf(uz, B[6 + j], fmaf(uy, B[3 + j], ux * B[j])));
      const float ra = fmaf(n1.e[i2], C[3 * i1 + j], n1.e[i1] * C[3 * i2 + j]);
      const float rb = fmaf(n2.e[j2], C[3 * i + j1], n2.e[j1] * C[3 * i + j2]);
      best = fmaxf(best, num - ra - rb);
    }
  }
  const float M = n1.s + n2.s + t0_l1;
  return best > 1.52587890625e-05f * M;  // 256 * 2^-24 * M
}


// Lower bound on the distance between two triangles, used to skip exact triDistance calls.
// Inputs are LOCAL single-precision coordinates: both triangles translated (in FP64, by the
// caller) so that S0 is the origin, then rounded once.  The bound is the largest support-
// function gap over 19 directions (both normals, the nine edge-edge cross products, the six
// in-plane edge normals, the centroid direction) -- any direction yields a valid bound, so
// rounding errors in the directions themselves are harmless; only the six projections and
// the norm of each candidate contribute error: <= 13 u Lmax + 3 u gap.  Slack: 32 u Lsum
// (Lsum >= |p|_2 for every vertex) plus 1e-6 relative.
FD float tri_lower_bound_f32(const float* s1, const float* s2, const float* t0, const float* t1, const float* t2) {
  // candidate evaluation: S = {0, s1, s2}, T = {t0, t1, t2}
  float best = 0.0f;
  auto consider = [&](float lx, float ly, float lz) {
    const float len2 = fmaf(lz, lz, fmaf(ly, ly, lx * lx));
    const float a1 = fmaf(lz, s1[2], fmaf(ly, s1[1], lx * s1[0]));
    const float a2 = fmaf(lz, s2[2], fmaf(ly, s2[1], lx * s2[0]));
    const float b0 = fmaf(lz, t0[2], fmaf(ly, t0[1], lx * t0[0]));
    const float b1 = fmaf(lz, t1[2], fmaf(ly, t1[1], lx * t1[0]));
    const float b2 = fmaf(lz, t2[2], fmaf(ly, t2[1], lx * t2[0]));
    const float mnS = fminf(0.0f, fminf(a1, a2)), mxS = fmaxf(0.0f, fmaxf(a1, a2));
    const float mnT = fminf(b0, fminf(b1, b2)), mxT = fmaxf(b0, fmaxf(b1, b2));
    const float g = fmaxf(mnT - mxS, mnS - mxT);
    const bool ok = (len2 > 1e-30f) && (len2 < 1e30f);
    const float v = g * f32_rsqrt(ok ? len2 : 1.0f);
    best = fmaxf(best, ok ? v : 0.0f);
  };
  const float e0[3] = {s1[0], s1[1], s1[2]};
  const float e1[3] = {s2[0] - s1[0], s2[1] - s1[1], s2[2] - s1[2]};
  const float e2[3] = {-s2[0], -s2[1], -s2[2]};
  const float f0[3] = {t1[0] - t0[0], t1[1] - t0[1], t1[2] - t0[2]};
  const float f1[3] = {t2[0] - t1[0], t2[1] - t1[1], t2[2] - t1[2]};
  const float f2[3] = {t0[0] - t2[0], t0[1] - t2[1], t0[2] - t2[2]};
  const float* E[3] = {e0, e1, e2};
  const float* Fv[3] = {f0, f1, f2};
  float n1[3], n2[3];
  n1[0] = fmaf(e0[1], e1[2], -(e0[2] * e1[1])); n1[1] = fmaf(e0[2], e1[0], -(e0[0] * e1[2])); n1[2] = fmaf(e0[0], e1[1], -(e0[1] * e1[0]));
  n2[0] = fmaf(f0[1], f1[2], -(f0[2] * f1[1])); n2[1] = fmaf(f0[2], f1[0], -(f0[0] * f1[2])); n2[2] = fmaf(f0[0], f1[1], -(f0[1] * f1[0]));
  consider(n1[0], n1[1], n1[2]);
  consider(n2[0], n2[1], n2[2]);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j)
      consider(fmaf(E[i][1], Fv[j][2], -(E[i][2] * Fv[j][1])), fmaf(E[i][2], Fv[j][0], -(E[i][0] * Fv[j][2])),
               fmaf(E[i][0], Fv[j][1], -(E[i][1] * Fv[j][0])));
    consider(fmaf(E[i][1], n1[2], -(E[i][2] * n1[1])), fmaf(E[i][2], n1[0], -(E[i][0] * n1[2])),
             fmaf(E[i][0], n1[1], -(E[i][1] * n1[0])));
    consider(fmaf(Fv[i][1], n2[2], -(Fv[i][2] * n2[1])), fmaf(Fv[i][2], n2[0], -(Fv[i][0] * n2[2])),
             fmaf(Fv[i][0], n2[1], -(Fv[i][1] * n2[0])));
  }
  consider((t0[0] + t1[0] + t2[0]) - (s1[0] + s2[0]), (t0[1] + t1[1] + t2[1]) - (s1[1] + s2[1]),
           (t0[2] + t1[2] + t2[2]) - (s1[2] + s2[2]));
  auto l1 = [](const float* p) { return fabsf(p[0]) + fabsf(p[1]) + fabsf(p[2]); };
  const float Lsum = fmaxf(fmaxf(l1(s1), l1(s2)), fmaxf(l1(t0), fmaxf(l1(t1), l1(t2))));
  const float lb = fmaf(best, 0.999999f, -(1.9073486328125e-06f * Lsum));  // 32 * 2^-24 * Lsum
  return fmaxf(lb, 0.0f);
}


// Selectable subset of the 19 directions of tri_lower_bound_f32 (same validity argument and slack; every
// direction only adds cost and tightness): kDirs bit 0 = the two face normals, bit 1 = centroid direction,
// bit 2 = the nine edge x edge products, bit 3 = the six in-plane edge normals.
template <int kDirs>
FD float tri_lower_bound_dirs_f32(const float* s1, const float* s2, const float* t0, const float* t1, const float* t2) {
  float best = 0.0f;
  auto consider = [&](float lx, float ly, float lz) {
    const float len2 = fmaf(lz, lz, fmaf(ly, ly, lx * lx));
    const float a1 = fmaf(lz, s1[2], fmaf(ly, s1[1], lx * s1[0]));
    const float a2 = fmaf(lz, s2[2], fmaf(ly, s2[1], lx * s2[0]));
    const float b0 = fmaf(lz, t0[2], fmaf(ly, t0[1], lx * t0[0]));
    const float b1 = fmaf(lz, t1[2], fmaf(ly, t1[1], lx * t1[0]));
    const float b2 = fmaf(lz, t2[2], fmaf(ly, t2[1], lx * t2[0]));
    const float mnS = fminf(0.0f, fminf(a1, a2)), mxS = fmaxf(0.0f, fmaxf(a1, a2));
    const float mnT = fminf(b0, fminf(b1, b2)), mxT = fmaxf(b0, fmaxf(b1, b2));
    const float g = fmaxf(mnT - mxS, mnS - mxT);
    const bool ok = (len2 > 1e-30f) && (len2 < 1e30f);
    const float v = g * f32_rsqrt(ok ? len2 : 1.0f);
    best = fmaxf(best, ok ? v : 0.0f);
  };
  const float e0[3] = {s1[0], s1[1], s1[2]};
  const float e1[3] = {s2[0] - s1[0], s2[1] - s1[1], s2[2] - s1[2]};
  const float e2[3] = {-s2[0], -s2[1], -s2[2]};
  const float f0[3] = {t1[0] - t0[0], t1[1] - t0[1], t1[2] - t0[2]};
  const float f1[3] = {t2[0] - t1[0], t2[1] - t1[1], t2[2] - t1[2]};
  const float f2[3] = {t0[0] - t2[0], t0[1] - t2[1], t0[2] - t2[2]};
  const float* E[3] = {e0, e1, e2};
  const float* Fv[3] = {f0, f1, f2};
  float n1[3], n2[3];
  n1[0] = fmaf(e0[1], e1[2], -(e0[2] * e1[1])); n1[1] = fmaf(e0[2], e1[0], -(e0[0] * e1[2])); n1[2] = fmaf(e0[0], e1[1], -(e0[1] * e1[0]));
  n2[0] = fmaf(f0[1], f1[2], -(f0[2] * f1[1])); n2[1] = fmaf(f0[2], f1[0], -(f0[0] * f1[2])); n2[2] = fmaf(f0[0], f1[1], -(f0[1] * f1[0]));
  if (kDirs & 1) {
    consider(n1[0], n1[1], n1[2]);
    consider(n2[0], n2[1], n2[2]);
  }
  if (kDirs & 2)
    consider((t0[0] + t1[0] + t2[0]) - (s1[0] + s2[0]), (t0[1] + t1[1] + t2[1]) - (s1[1] + s2[1]),
             (t0[2] + t1[2] + t2[2]) - (s1[2] + s2[2]));
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if (kDirs & 4) {
#pragma unroll
      for (int j = 0; j < 3; ++j)
        consider(fmaf(E[i][1], Fv[j][2], -(E[i][2] * Fv[j][1])), fmaf(E[i][2], Fv[j][0], -(E[i][0] * Fv[j][2])),
                 fmaf(E[i][0], Fv[j][1], -(E[i][1] * Fv[j][0])));
    }
    if (kDirs & 8) {
      consider(fmaf(E[i][1], n1[2], -(E[i][2] * n1[1])), fmaf(E[i][2], n1[0], -(E[i][0] * n1[2])),
               fmaf(E[i][0], n1[1], -(E[i][1] * n1[0])));
      consider(fmaf(Fv[i][1], n2[2], -(Fv[i][2] * n2[1])), fmaf(Fv[i][2], n2[0], -(Fv[i][0] * n2[2])),
               fmaf(Fv[i][0], n2[1], -(Fv[i][1] * n2[0])));
    }
  }
  auto l1 = [](const float* p) { return fabsf(p[0]) + fabsf(p[1]) + fabsf(p[2]); };
  const float Lsum = fmaxf(fmaxf(l1(s1), l1(s2)), fmaxf(l1(t0), fmaxf(l1(t1), l1(t2))));
  const float lb = fmaf(best, 0.999999f, -(1.9073486328125e-06f * Lsum));  // 32 * 2^-24 * Lsum
  return fmaxf(lb, 0.0f);
}

// Two-sided single-precision bounds on the distance between two triangles: lo <= d <= hi.
// Same local coordinates as tri_lower_bound_f32 (S = {0, s1, s2}, T = {t0, t1, t2}, translated in FP64, rounded once).
//   * Fifteen candidate point pairs (X on S, Y on T), each an ACTUAL pair of points of the two (rounded) triangles
//     whatever the rounding of its parameters: the nine edge / edge pairs (clamped closed form) and the six
//     vertex / face pairs (barycentric coordinates of the projection, clamped into the triangle).
//   * hi = the smallest candidate length, plus slack: an upper bound because the two points exist.
//   * lo = the support-function gap along V = Y - X of that candidate, minus slack: a lower bound along ANY
//     direction, and tight along this one (in exact arithmetic V of the true closest pair gives gap = d).
// Error budget (u = 2^-24, L = largest |coordinate| <= Lsum): a candidate point is evaluated with <= 2 u L per
// coordinate, V and its length with <= 3 u |V|, rounding the inputs moves every point by <= sqrt(3) u L:
// |hi_computed - d(X, Y)| <= 11 u L + 3 u |V|; the projections of lo as in tri_lower_bound_f32 (<= 13 u L + 3 u gap).
// Slack: 64 u Lsum + 2e-6 relative on hi (covers the 2-ulp rsqrt), 32 u Lsum + 1e-6 relative on lo.
// `trust_hi` is false when a face normal is so short that the reference's triDistance skips its vertex / face stage
// (|N|^2 <= 1e-15, triangle_distance-inl.h:262,318) and may return MORE than the true distance: hi bounds the true
// distance only, so the caller must not prune with it then.
FD void tri_closest_bounds_f32(const float* s1, const float* s2, const float* t0, const float* t1, const float* t2,
                               float& lo, float& hi, bool& trust_hi) {
  const float P[3][3] = {{0.0f, 0.0f, 0.0f}, {s1[0], s1[1], s1[2]}, {s2[0], s2[1], s2[2]}};
  const float Q[3][3] = {{t0[0], t0[1], t0[2]}, {t1[0], t1[1], t1[2]}, {t2[0], t2[1], t2[2]}};
  float E[3][3], Fv[3][3], a[3], e[3], ra[3], re[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int n = (i + 1) % 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      E[i][c] = P[n][c] - P[i][c];
      Fv[i][c] = Q[n][c] - Q[i][c];
    }
    a[i] = fmaf(E[i][2], E[i][2], fmaf(E[i][1], E[i][1], E[i][0] * E[i][0]));
    e[i] = fmaf(Fv[i][2], Fv[i][2], fmaf(Fv[i][1], Fv[i][1], Fv[i][0] * Fv[i][0]));
    ra[i] = f32_rcp(fmaxf(a[i], 1e-30f));
    re[i] = f32_rcp(fmaxf(e[i], 1e-30f));
  }
  float best = 3.0e38f, V[3] = {0.0f, 0.0f, 0.0f};
  auto take = [&](float vx, float vy, float vz) {
    const float dd = fmaf(vz, vz, fmaf(vy, vy, vx * vx));
    const bool better = dd < best;
    best = better ? dd : best;
    V[0] = better ? vx : V[0];
    V[1] = better ? vy : V[1];
    V[2] = better ? vz : V[2];
  };
  auto clamp01 = [](float x) { return fminf(fmaxf(x, 0.0f), 1.0f); };
  // edge i of S against edge j of T
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float r0 = P[i][0] - Q[j][0], r1 = P[i][1] - Q[j][1], r2 = P[i][2] - Q[j][2];
      const float b = fmaf(E[i][2], Fv[j][2], fmaf(E[i][1], Fv[j][1], E[i][0] * Fv[j][0]));
      const float c = fmaf(E[i][2], r2, fmaf(E[i][1], r1, E[i][0] * r0));
      const float f = fmaf(Fv[j][2], r2, fmaf(Fv[j][1], r1, Fv[j][0] * r0));
      const float denom = fmaf(a[i], e[j], -(b * b));
      float s = clamp01(fmaf(b, f, -(c * e[j])) * f32_rcp(fmaxf(denom, 1e-30f)));
      s = denom > 1e-6f * (a[i] * e[j]) ? s : 0.0f;  // nearly parallel: any s is a valid point
      float t = fmaf(b, s, f) * re[j];
      const float s_lo = clamp01(-c * ra[i]), s_hi = clamp01((b - c) * ra[i]);
      s = t < 0.0f ? s_lo : (t > 1.0f ? s_hi : s);
      t = clamp01(t);
      // V = Y - X = (Q_j + t F_j) - (P_i + s E_i) = -r + t F_j - s E_i
      take(fmaf(-s, E[i][0], fmaf(t, Fv[j][0], -r0)), fmaf(-s, E[i][1], fmaf(t, Fv[j][1], -r1)),
           fmaf(-s, E[i][2], fmaf(t, Fv[j][2], -r2)));
    }
  }
  // vertex of one triangle against the face of the other: barycentric coordinates (u, v) of the projection with respect
  // to origin O and edge vectors g1, g2, clamped into the triangle; sign = +1 when the vertex belongs to T (V = Y - X)
  auto vertex_face = [&](const float* O, const float* g1, const float* g2, float d11, float d12, float d22, float rdet,
                         const float* Y, float sign) {
    const float w0 = Y[0] - O[0], w1 = Y[1] - O[1], w2 = Y[2] - O[2];
    const float p1 = fmaf(g1[2], w2, fmaf(g1[1], w1, g1[0] * w0));
    const float p2 = fmaf(g2[2], w2, fmaf(g2[1], w1, g2[0] * w0));
    float uu = fmaf(d22, p1, -(d12 * p2)) * rdet;
    float vv = fmaf(d11, p2, -(d12 * p1)) * rdet;
    uu = fminf(fmaxf(uu, 0.0f), 1.0f);
    vv = fminf(fmaxf(vv, 0.0f), 1.0f - uu);
    take(sign * fmaf(-vv, g2[0], fmaf(-uu, g1[0], w0)), sign * fmaf(-vv, g2[1], fmaf(-uu, g1[1], w1)),
         sign * fmaf(-vv, g2[2], fmaf(-uu, g1[2], w2)));
  };
  float gS2[3], gT2[3];  // second edge vector from the first vertex: S2 - S0, T2 - T0
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    gS2[c] = P[2][c];
    gT2[c] = Q[2][c] - Q[0][c];
  }
  const float dS12 = fmaf(E[0][2], gS2[2], fmaf(E[0][1], gS2[1], E[0][0] * gS2[0]));
  const float dT12 = fmaf(Fv[0][2], gT2[2], fmaf(Fv[0][1], gT2[1], Fv[0][0] * gT2[0]));
  float nS[3], nT[3];
  nS[0] = fmaf(E[0][1], gS2[2], -(E[0][2] * gS2[1])); nS[1] = fmaf(E[0][2], gS2[0], -(E[0][0] * gS2[2])); nS[2] = fmaf(E[0][0], gS2[1], -(E[0][1] * gS2[0]));
  nT[0] = fmaf(Fv[0][1], gT2[2], -(Fv[0][2] * gT2[1])); nT[1] = fmaf(Fv[0][2], gT2[0], -(Fv[0][0] * gT2[2])); nT[2] = fmaf(Fv[0][0], gT2[1], -(Fv[0][1] * gT2[0]));
  const float nlS = fmaf(nS[2], nS[2], fmaf(nS[1], nS[1], nS[0] * nS[0]));
  const float nlT = fmaf(nT[2], nT[2], fmaf(nT[1], nT[1], nT[0] * nT[0]));
  const float rdS = f32_rcp(fmaxf(nlS, 1e-30f)), rdT = f32_rcp(fmaxf(nlT, 1e-30f));  // |g1 x g2|^2 = d11 d22 - d12^2
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    vertex_face(P[0], E[0], gS2, a[0], dS12, a[2], rdS, Q[k], 1.0f);
    vertex_face(Q[0], Fv[0], gT2, e[0], dT12, e[2], rdT, P[k], -1.0f);
  }
  auto l1 = [](const float* p) { return fabsf(p[0]) + fabsf(p[1]) + fabsf(p[2]); };
  const float Lsum = fmaxf(fmaxf(l1(s1), l1(s2)), fmaxf(l1(t0), fmaxf(l1(t1), l1(t2))));
  const bool ok = (best > 1e-30f) && (best < 1e30f);
  const float rlen = f32_rsqrt(ok ? best : 1.0f);
  const float len = ok ? best * rlen : (best < 1.0f ? 1e-15f : 3.0e38f);
  hi = fmaf(len, 1.000002f, 3.814697265625e-06f * Lsum);  // 64 * 2^-24 * Lsum
  // support-function gap along V
  const float a1 = fmaf(V[2], s1[2], fmaf(V[1], s1[1], V[0] * s1[0]));
  const float a2 = fmaf(V[2], s2[2], fmaf(V[1], s2[1], V[0] * s2[0]));
  const float b0 = fmaf(V[2], t0[2], fmaf(V[1], t0[1], V[0] * t0[0]));
  const float b1 = fmaf(V[2], t1[2], fmaf(V[1], t1[1], V[0] * t1[0]));
  const float b2 = fmaf(V[2], t2[2], fmaf(V[1], t2[1], V[0] * t2[0]));
  const float g = fminf(b0, fminf(b1, b2)) - fmaxf(0.0f, fmaxf(a1, a2));
  const float v = ok ? g * rlen : 0.0f;
  lo = fmaxf(fmaf(v, 0.999999f, -(1.9073486328125e-06f * Lsum)), 0.0f);  // 32 * 2^-24 * Lsum
  // |N| computed with <= 8 u |g1| |g2| absolute error; the reference compares |N|^2 with 1e-15 (|N| > 3.17e-8)
  const float limS = 3.2e-8f + 5e-7f * sqrtf(a[0] * a[2]), limT = 3.2e-8f + 5e-7f * sqrtf(e[0] * e[2]);
  trust_hi = (nlS > limS * limS) && (nlT > limT * limT) && (best < 1e30f);
}

// Single-precision classification of a triangle pair for the collide leaf test:
//   +1  certainly separated   (some axis of the reference's 17-axis SAT separates the projections
//                              by more than the rounding margin => intersect_Triangle is false)
//   -1  certainly intersecting (on all 17 axes the projections overlap by more than the margin
//                              => intersect_Triangle is true)
//    0  undecided: run the exact FP64 test.
// Inputs: local coordinates (both triangles translated by -P1 in FP64 by the caller, so p1 = 0,
// and rounded once).  Error budget with Lm = largest |coordinate| (u = 2^-24): edge vectors
// <= 3 u Lm, first-level axes (n1, m1, e_i x f_j) <= 28 u Lm^2, their projections <= 50 u Lm^3;
// second-level axes (e_i x n1, f_j x m1) <= 140 u Lm^3, projections <= 260 u Lm^4.  Margins used:
// 1024 u Lm^3 and 8192 u Lm^4.
FD int tri_classify_f32(const float* p2, const float* p3, const float* q1, const float* q2, const float* q3) {
  auto amax3 = [](const float* v) { return fmaxf(fabsf(v[0]), fmaxf(fabsf(v[1]), fabsf(v[2]))); };
  const float Lm = fmaxf(fmaxf(amax3(p2), amax3(p3)), fmaxf(amax3(q1), fmaxf(amax3(q2), amax3(q3))));
  const float L3 = Lm * Lm * Lm;
  const float m1 = 6.103515625e-05f * L3;          // 1024 * 2^-24 * Lm^3
  const float m2 = 4.8828125e-04f * L3 * Lm;       // 8192 * 2^-24 * Lm^4
  bool sep = false, all_overlap = true;
  auto axis = [&](float ax, float ay, float az, float margin) {
    const float P2 = fmaf(az, p2[2], fmaf(ay, p2[1], ax * p2[0]));
    const float P3 = fmaf(az, p3[2], fmaf(ay, p3[1], ax * p3[0]));
    const float Q1 = fmaf(az, q1[2], fmaf(ay, q1[1], ax * q1[0]));
    const float Q2 = fmaf(az, q2[2], fmaf(ay, q2[1], ax * q2[0]));
    const float Q3 = fmaf(az, q3[2], fmaf(ay, q3[1], ax * q3[0]));
    const float mnP = fminf(0.0f, fminf(P2, P3)), mxP = fmaxf(0.0f, fmaxf(P2, P3));
    const float mnQ = fminf(Q1, fminf(Q2, Q3)), mxQ = fmaxf(Q1, fmaxf(Q2, Q3));
    const float gap = fmaxf(mnP - mxQ, mnQ - mxP);  // > 0: separated on this axis
    sep = sep || (gap > margin);
    all_overlap = all_overlap && (gap < -margin);
  };
  const float e1[3] = {p2[0], p2[1], p2[2]};
  const float e2[3] = {p3[0] - p2[0], p3[1] - p2[1], p3[2] - p2[2]};
  const float e3[3] = {-p3[0], -p3[1], -p3[2]};
  const float f1[3] = {q2[0] - q1[0], q2[1] - q1[1], q2[2] - q1[2]};
  const float f2[3] = {q3[0] - q2[0], q3[1] - q2[1], q3[2] - q2[2]};
  const float f3[3] = {q1[0] - q3[0], q1[1] - q3[1], q1[2] - q3[2]};
  const float* E[3] = {e1, e2, e3};
  const float* Fv[3] = {f1, f2, f3};
  float n1[3], mm[3];
  n1[0] = fmaf(e1[1], e2[2], -(e1[2] * e2[1])); n1[1] = fmaf(e1[2], e2[0], -(e1[0] * e2[2])); n1[2] = fmaf(e1[0], e2[1], -(e1[1] * e2[0]));
  mm[0] = fmaf(f1[1], f2[2], -(f1[2] * f2[1])); mm[1] = fmaf(f1[2], f2[0], -(f1[0] * f2[2])); mm[2] = fmaf(f1[0], f2[1], -(f1[1] * f2[0]));
  axis(n1[0], n1[1], n1[2], m1);
  axis(mm[0], mm[1], mm[2], m1);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      axis(fmaf(E[i][1], Fv[j][2], -(E[i][2] * Fv[j][1])), fmaf(E[i][2], Fv[j][0], -(E[i][0] * Fv[j][2])),
           fmaf(E[i][0], Fv[j][1], -(E[i][1] * Fv[j][0])), m1);
#pragma unroll
  for (int i = 0; i < 3; ++i)
    axis(fmaf(E[i][1], n1[2], -(E[i][2] * n1[1])), fmaf(E[i][2], n1[0], -(E[i][0] * n1[2])),
         fmaf(E[i][0], n1[1], -(E[i][1] * n1[0])), m2);
#pragma unroll
  for (int j = 0; j < 3; ++j)
    axis(fmaf(Fv[j][1], mm[2], -(Fv[j][2] * mm[1])), fmaf(Fv[j][2], mm[0], -(Fv[j][0] * mm[2])),
         fmaf(Fv[j][0], mm[1], -(Fv[j][1] * mm[0])), m2);
  return sep ? 1 : (all_overlap ? -1 : 0);
}

}  // namespace fclgpu
