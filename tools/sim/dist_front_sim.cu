// Host model of distance_warp_kernel's schedule (csrc/traversal.cuh): same rounds, same pops, same single-precision
// bounds (the product headers compiled for the host), exact triDistance from device_math.cuh.  It counts what a warp
// would issue -- rounds by kind, lanes in use, seg_points iterations -- so that schedule changes can be compared for
// their WORK before they cost GPU time.  Development tool: not product code, not the oracle.
#include <algorithm>
#include <cfloat>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cstdio>
#include <cstdlib>

#include "../../fcl_b200/csrc/device_math.cuh"
#include "../../fcl_b200/csrc/bounds_f32.cuh"
#include "../../fcl_b200/csrc/records.hpp"
using namespace fclgpu;

struct SimModel {
  int n_nodes, n_tris;
  const int32_t* fc;
  const double* size;     // OBB size per node
  std::vector<RssRec32> rss32;
  const double* tri;      // 9 doubles per triangle
  const double *axis, *rTo, *rl, *rr;
  std::vector<float> wit;  // witness point per node (3 floats): a point ON a triangle of the node's subtree
};

struct Stats {
  double bv_rounds, bv_tests, bv_lanes_pop;
  double scr_rounds, scr_lanes;       // direction screen
  double cls_rounds, cls_lanes;       // closest-point screen
  double ex_rounds, ex_lanes, ex_iter_sum, ex_iter_max, ex_tail;
  double max_sp, queries;
};

struct Opt {
  int mode;           // 0 = kernel as shipped (direction screen -> exact); 1 = closest-point screen (lo/hi), exact deferred
  int pop;            // entries expanded per BV round
  int leaf_trigger;   // exact queue length that triggers an exact round (mode 0)
  int raw_trigger;    // raw queue length that triggers a screening round
  int eager_first;    // mode 1: screen as soon as any raw pair exists while no finite upper bound is known
  int dirs_first;     // mode 1: run the cheap direction screen in front of the closest-point screen
  int use_hi;         // mode 1: prune with the upper bounds
  int seed_levels;    // start from the pose-independent front after this many unconditional expansions (0 = root pair)
  int exact_rss;      // bound children with the exact FP64 RSS distance (traversal 1)
  int eager0;         // mode 0: first exact round as soon as a screened pair exists and no minimum is known
  int witness;        // upper bound from the witness points of every tested node pair
  int cheap_hi;       // mode 0: upper bound next to the direction screen: 1 = nearest vertex pair, 2 = six vertex/face clamps, 3 = full closest-point routine
  int sort_bits;      // 0 = full sort of the children; else rank by a linear bucket of this many bits between the round's min and max bound
  int quad;           // expand `pop` entries by TWO levels (up to 4 grandchild pairs each, no test at the level between)
};

static double g_hist[4][33];
static inline float fl_ru(double x) { return round_up_f32(x); }

// exact triDistance with the number of edge-pair iterations executed (early return) and whether the tail ran
static double tri_distance_counted(const V3 T1[3], const V3 T2[3], V3& P, V3& Q, int& iters, int& tail) {
  V3 minP = mk(0, 0, 0), minQ = mk(0, 0, 0);
  int shown_disjoint = 0;
  const V3 d00 = T1[0] - T2[0];
  double mindd = dot(d00, d00) + 1;
  iters = 0;
  tail = 0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      ++iters;
      V3 VEC;
      const V3 A0 = T1[i], A1 = T1[(i + 1) % 3], A2 = T1[(i + 2) % 3];
      const V3 B0 = T2[j], B1 = T2[(j + 1) % 3], B2 = T2[(j + 2) % 3];
      seg_points(A0, A1 - A0, B0, B1 - B0, VEC, P, Q);
      const V3 V = Q - P;
      const double dd = dot(V, V);
      if (dd <= mindd) {
        minP = P; minQ = Q; mindd = dd;
        double a = dot(A2 - P, VEC), b = dot(B2 - Q, VEC);
        if ((a <= 0) && (b >= 0)) return sqrt(dd);
        const double p = dot(V, VEC);
        if (a < 0) a = 0;
        if (b > 0) b = 0;
        if ((p - a + b) > 0) shown_disjoint = 1;
      }
    }
  tail = 1;
  V3 P2, Q2;
  const double d = tri_distance(T1, T2, P2, Q2);  // the product routine gives the final answer
  P = P2; Q = Q2;
  (void)minP; (void)minQ; (void)shown_disjoint;
  return d;
}

struct Entry { uint32_t x, y; float b; };
struct Leaf { uint32_t x, y; float b; };

static void local_f32(const SimModel& m1, const SimModel& m2, const M3& R, const V3& T, uint32_t ix, uint32_t iy, float* s1,
                      float* s2, float* t0, float* t1, float* t2, V3* Sv, V3* Tv) {
  for (int k = 0; k < 3; ++k) {
    Sv[k] = mk(m1.tri[9 * (size_t)ix + 3 * k], m1.tri[9 * (size_t)ix + 3 * k + 1], m1.tri[9 * (size_t)ix + 3 * k + 2]);
    Tv[k] = mk(m2.tri[9 * (size_t)iy + 3 * k], m2.tri[9 * (size_t)iy + 3 * k + 1], m2.tri[9 * (size_t)iy + 3 * k + 2]);
  }
  const V3 a = Sv[1] - Sv[0], c = Sv[2] - Sv[0];
  s1[0] = (float)a.x; s1[1] = (float)a.y; s1[2] = (float)a.z;
  s2[0] = (float)c.x; s2[1] = (float)c.y; s2[2] = (float)c.z;
  const V3 u0 = (mulv(R, Tv[0]) + T) - Sv[0], u1 = (mulv(R, Tv[1]) + T) - Sv[0], u2 = (mulv(R, Tv[2]) + T) - Sv[0];
  t0[0] = (float)u0.x; t0[1] = (float)u0.y; t0[2] = (float)u0.z;
  t1[0] = (float)u1.x; t1[1] = (float)u1.y; t1[2] = (float)u1.z;
  t2[0] = (float)u2.x; t2[1] = (float)u2.y; t2[2] = (float)u2.z;
}

static double run_query(const SimModel& m1, const SimModel& m2, const double* pose12, const Opt& o, Stats& st, double init_hi) {
  // pose of model 2 (identity) in model 1's frame: R = R1^T, T = -R1^T t1 (kernel prologue with tf2 = identity)
  M3 R1; for (int k = 0; k < 9; ++k) R1.m[k] = pose12[k];
  M3 I; for (int k = 0; k < 9; ++k) I.m[k] = (k % 4 == 0) ? 1.0 : 0.0;
  const V3 t1v = mk(pose12[9], pose12[10], pose12[11]);
  const M3 R = mulTM(R1, I);
  const V3 it = mulTv(R1, t1v);
  const V3 T = mulTv(R1, mk(0, 0, 0)) + mk(-it.x, -it.y, -it.z);
  float Rf[9], Tf[3];
  for (int k = 0; k < 9; ++k) Rf[k] = (float)R.m[k];
  Tf[0] = (float)T.x; Tf[1] = (float)T.y; Tf[2] = (float)T.z;
  const float t_l1 = fl_ru((fabs(T.x) + fabs(T.y)) + fabs(T.z));

  double min_d = DBL_MAX;
  float min_f = fl_ru(init_hi);  // pruning bound (mode 1: also tightened by the upper bounds)
  std::vector<Entry> stack;  // back() = top
  std::vector<Leaf> raw, mid, exq;
  float wit_hi = 3.4e38f;
  auto bound_of = [&](uint32_t x, uint32_t y) -> float {
    if (o.witness == 1) {
      const float* w1 = &m1.wit[3 * (size_t)x];
      const float* w2 = &m2.wit[3 * (size_t)y];
      const float px = fmaf(Rf[2], w2[2], fmaf(Rf[1], w2[1], fmaf(Rf[0], w2[0], Tf[0]))) - w1[0];
      const float py = fmaf(Rf[5], w2[2], fmaf(Rf[4], w2[1], fmaf(Rf[3], w2[0], Tf[1]))) - w1[1];
      const float pz = fmaf(Rf[8], w2[2], fmaf(Rf[7], w2[1], fmaf(Rf[6], w2[0], Tf[2]))) - w1[2];
      const float len = sqrtf(fmaf(pz, pz, fmaf(py, py, px * px)));
      const float M = m1.rss32[x].s + m2.rss32[y].s + t_l1;
      wit_hi = fminf(wit_hi, fmaf(len, 1.00001f, 1e-5f * M));
    }
    if (o.exact_rss) {
      M3 a1, a2;
      for (int k = 0; k < 9; ++k) { a1.m[k] = m1.axis[9 * (size_t)x + k]; a2.m[k] = m2.axis[9 * (size_t)y + k]; }
      const double d = rss_pair_distance(R, T, a1, mk(m1.rTo[3 * x], m1.rTo[3 * x + 1], m1.rTo[3 * x + 2]), m1.rl + 2 * x, m1.rr[x], a2,
                                         mk(m2.rTo[3 * y], m2.rTo[3 * y + 1], m2.rTo[3 * y + 2]), m2.rl + 2 * y, m2.rr[y]);
      float f = (float)d;
      if ((double)f > d) f = std::nextafterf(f, -1.0f);
      return f;
    }
    return rss_lower_bound_f32(Rf, Tf, t_l1, m1.rss32[x], m2.rss32[y]);
  };
  exq.push_back({0u, 0u, -1.0f});
  if (o.seed_levels > 0) {
    // pose-independent: expand every entry seed_levels times (firstOverSecond), leaf pairs stay as they are
    std::vector<Entry> cur{{0u, 0u, -1.0f}};
    for (int lv = 0; lv < o.seed_levels; ++lv) {
      std::vector<Entry> nxt;
      for (auto& en : cur) {
        const int fc1 = m1.fc[en.x], fc2 = m2.fc[en.y];
        const bool l1 = fc1 < 0, l2 = fc2 < 0;
        if (l1 && l2) { nxt.push_back(en); continue; }
        if (l2 || (!l1 && (m1.size[en.x] > m2.size[en.y]))) { nxt.push_back({(uint32_t)fc1, en.y, 0}); nxt.push_back({(uint32_t)fc1 + 1, en.y, 0}); }
        else { nxt.push_back({en.x, (uint32_t)fc2, 0}); nxt.push_back({en.x, (uint32_t)fc2 + 1, 0}); }
      }
      cur.swap(nxt);
    }
    for (auto& en : cur) en.b = bound_of(en.x, en.y);
    std::stable_sort(cur.begin(), cur.end(), [](const Entry& a, const Entry& b) { return a.b < b.b; });
    for (int r = (int)cur.size() - 1; r >= 0; --r) stack.push_back(cur[r]);
    st.bv_rounds += (cur.size() + 31) / 32; st.bv_tests += cur.size();
  } else {
    stack.push_back({0u, 0u, -1.0f});
  }
  size_t max_sp = 1;

  auto exact_round = [&]() {
    // filter the whole queue by the current bound first (mode 1 compaction), then test up to 32 from the back
    if (o.mode == 1) {
      std::vector<Leaf> keep;
      for (auto& l : exq) if (l.b < min_f) keep.push_back(l);
      exq.swap(keep);
      if (exq.empty()) return;
    }
    const int k = (int)std::min<size_t>(32, exq.size());
    int lanes = 0, itmax = 0, tails = 0;
    double itsum = 0;
    double best = DBL_MAX;
    for (int l = 0; l < k; ++l) {
      const Leaf lf = exq[exq.size() - k + l];
      if (!(lf.b < min_f)) continue;
      ++lanes;
      V3 Sv[3], Tv[3];
      for (int c = 0; c < 3; ++c) {
        Sv[c] = mk(m1.tri[9 * (size_t)lf.x + 3 * c], m1.tri[9 * (size_t)lf.x + 3 * c + 1], m1.tri[9 * (size_t)lf.x + 3 * c + 2]);
        Tv[c] = mulv(R, mk(m2.tri[9 * (size_t)lf.y + 3 * c], m2.tri[9 * (size_t)lf.y + 3 * c + 1], m2.tri[9 * (size_t)lf.y + 3 * c + 2])) + T;
      }
      V3 P, Q;
      int it, tl;
      const double d = tri_distance_counted(Sv, Tv, P, Q, it, tl);
      itsum += it; itmax = std::max(itmax, it); tails += tl;
      best = std::min(best, d);
    }
    exq.resize(exq.size() - k);
    st.ex_rounds += 1; st.ex_lanes += lanes; st.ex_iter_sum += itsum; st.ex_iter_max += itmax; st.ex_tail += tails ? 1 : 0;
    if (best < min_d) {
      min_d = best;
      min_f = std::min(min_f, fl_ru(best));
    }
  };

  while (true) {
    // ---- screening rounds
    if (o.mode == 0) {
      const bool can_screen = !raw.empty() && (exq.size() + std::min<size_t>(raw.size(), 32) <= 64);
      if (can_screen && ((int)raw.size() >= o.raw_trigger || stack.empty() || (o.eager0 >= 2 && (int)raw.size() >= o.eager0 && !(min_f < 3.0e38f)))) {
        const int k = (int)std::min<size_t>(32, raw.size());
        int lanes = 0;
        float round_hi0 = 3.4e38f;
        for (int l = 0; l < k; ++l) {
          Leaf lf = raw[raw.size() - k + l];
          if (!(lf.b < min_f)) continue;
          ++lanes;
          float s1[3], s2[3], t0[3], t1[3], t2[3];
          V3 Sv[3], Tv[3];
          local_f32(m1, m2, R, T, lf.x, lf.y, s1, s2, t0, t1, t2, Sv, Tv);
          const float lb = tri_lower_bound_dirs_f32<9>(s1, s2, t0, t1, t2);
          lf.b = fmaxf(lf.b, lb);
          if (lf.b < min_f) exq.push_back(lf);
          if (o.cheap_hi) {
            const float z[3] = {0, 0, 0};
            const float* SP[3] = {z, s1, s2};
            const float* TP[3] = {t0, t1, t2};
            float best = 3e38f;
            auto l1n = [](const float* p) { return fabsf(p[0]) + fabsf(p[1]) + fabsf(p[2]); };
            const float Lsum = fmaxf(fmaxf(l1n(s1), l1n(s2)), fmaxf(l1n(t0), fmaxf(l1n(t1), l1n(t2))));
            if (o.cheap_hi == 1) {
              for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) {
                const float dx = SP[a][0] - TP[b][0], dy = SP[a][1] - TP[b][1], dz = SP[a][2] - TP[b][2];
                best = fminf(best, dx * dx + dy * dy + dz * dz);
              }
              round_hi0 = fminf(round_hi0, sqrtf(best) * 1.00001f + 4e-6f * Lsum);
            } else if (o.cheap_hi == 2) {
              auto vf = [&](const float* O, const float* A, const float* B, const float* Y) {
                float g1[3], g2[3], w[3];
                for (int c = 0; c < 3; ++c) { g1[c] = A[c] - O[c]; g2[c] = B[c] - O[c]; w[c] = Y[c] - O[c]; }
                const float d11 = g1[0]*g1[0]+g1[1]*g1[1]+g1[2]*g1[2], d12 = g1[0]*g2[0]+g1[1]*g2[1]+g1[2]*g2[2], d22 = g2[0]*g2[0]+g2[1]*g2[1]+g2[2]*g2[2];
                const float p1 = g1[0]*w[0]+g1[1]*w[1]+g1[2]*w[2], p2 = g2[0]*w[0]+g2[1]*w[1]+g2[2]*w[2];
                const float rdet = 1.0f / fmaxf(d11 * d22 - d12 * d12, 1e-30f);
                float uu = (d22 * p1 - d12 * p2) * rdet, vv = (d11 * p2 - d12 * p1) * rdet;
                uu = fminf(fmaxf(uu, 0.0f), 1.0f); vv = fminf(fmaxf(vv, 0.0f), 1.0f - uu);
                float dd = 0; for (int c = 0; c < 3; ++c) { const float v = w[c] - uu * g1[c] - vv * g2[c]; dd += v * v; }
                best = fminf(best, dd);
              };
              for (int b = 0; b < 3; ++b) { vf(z, s1, s2, TP[b]); vf(t0, t1, t2, SP[b]); }
              round_hi0 = fminf(round_hi0, sqrtf(best) * 1.00001f + 4e-6f * Lsum);
            } else {
              float lo, hi; bool trust;
              tri_closest_bounds_f32(s1, s2, t0, t1, t2, lo, hi, trust);
              if (trust) round_hi0 = fminf(round_hi0, hi);
            }
          }
        }
        min_f = std::min(min_f, round_hi0);
        raw.resize(raw.size() - k);
        st.scr_rounds += 1; st.scr_lanes += lanes;
        continue;
      }
      if (o.cheap_hi && (int)exq.size() >= o.leaf_trigger) {  // compact the exact queue before deciding on a round
        std::vector<Leaf> keep;
        for (auto& l : exq) if (l.b < min_f) keep.push_back(l);
        exq.swap(keep);
      }
      const bool do_leaf = ((int)exq.size() >= o.leaf_trigger) || (stack.empty() && !exq.empty()) ||
                           (o.eager0 && exq.size() > 1 && !(min_f < 3.0e38f));
      if (do_leaf) { exact_round(); continue; }
    } else {
      const bool eager = o.eager_first && !raw.empty() && !(min_f < 3.0e38f);
      if (o.dirs_first) {
        if (!raw.empty() && ((int)raw.size() >= o.raw_trigger || stack.empty() || eager)) {
          const int k = (int)std::min<size_t>(32, raw.size());
          int lanes = 0;
          for (int l = 0; l < k; ++l) {
            Leaf lf = raw[raw.size() - k + l];
            if (!(lf.b < min_f)) continue;
            ++lanes;
            float s1[3], s2[3], t0[3], t1[3], t2[3];
            V3 Sv[3], Tv[3];
            local_f32(m1, m2, R, T, lf.x, lf.y, s1, s2, t0, t1, t2, Sv, Tv);
            const float lb = tri_lower_bound_dirs_f32<9>(s1, s2, t0, t1, t2);
            lf.b = fmaxf(lf.b, lb);
            if (lf.b < min_f) mid.push_back(lf);
          }
          raw.resize(raw.size() - k);
          st.scr_rounds += 1; st.scr_lanes += lanes;
          continue;
        }
      } else {
        for (auto& l : raw) mid.push_back(l);
        raw.clear();
      }
      if (!mid.empty() && ((int)mid.size() >= o.raw_trigger || (stack.empty() && raw.empty()) || eager)) {
        if (exq.size() + 32 > 64) { exact_round(); continue; }
        const int k = (int)std::min<size_t>(32, mid.size());
        int lanes = 0;
        float round_hi = 3.4e38f;
        for (int l = 0; l < k; ++l) {
          Leaf lf = mid[mid.size() - k + l];
          if (!(lf.b < min_f)) continue;
          ++lanes;
          float s1[3], s2[3], t0[3], t1[3], t2[3];
          V3 Sv[3], Tv[3];
          local_f32(m1, m2, R, T, lf.x, lf.y, s1, s2, t0, t1, t2, Sv, Tv);
          float lo, hi; bool trust;
          tri_closest_bounds_f32(s1, s2, t0, t1, t2, lo, hi, trust);
          lf.b = fmaxf(lf.b, lo);
          if (lf.b < min_f) exq.push_back(lf);
          if (o.use_hi && trust) round_hi = std::min(round_hi, hi);
        }
        min_f = std::min(min_f, round_hi);  // the warp applies the round's smallest upper bound after the round
        mid.resize(mid.size() - k);
        st.cls_rounds += 1; st.cls_lanes += lanes;
        continue;
      }
      if (stack.empty() && raw.empty() && mid.empty()) {
        if (!exq.empty()) { exact_round(); continue; }
      }
    }
    if (stack.empty()) break;

    // ---- BV round
    const int k = (int)std::min<size_t>(32, stack.size());
    std::vector<Entry> popped(stack.end() - k, stack.end());
    std::reverse(popped.begin(), popped.end());  // lane 0 = top
    stack.resize(stack.size() - k);
    std::vector<Entry> internal;
    int alive_n = 0;
    for (auto& en : popped) {
      if (!(en.b < min_f)) continue;
      ++alive_n;
      const int fc1 = m1.fc[en.x], fc2 = m2.fc[en.y];
      if (fc1 < 0 && fc2 < 0) raw.push_back({(uint32_t)(-(fc1 + 1)), (uint32_t)(-(fc2 + 1)), en.b});
      else internal.push_back(en);
    }
    const int n_int = (int)internal.size();
    const int n_exp = std::min(n_int, o.pop);
    // leftovers back on the stack: rank n_exp nearest -> top
    for (int r = n_int - 1; r >= n_exp; --r) stack.push_back(internal[r]);
    std::vector<Entry> kids;
    int n_tests = 0;
    auto split = [&](uint32_t x, uint32_t y, uint32_t* cx, uint32_t* cy) -> int {  // children of a pair (1 if leaf-leaf)
      const int fc1 = m1.fc[x], fc2 = m2.fc[y];
      const bool l1 = fc1 < 0, l2 = fc2 < 0;
      if (l1 && l2) { cx[0] = x; cy[0] = y; return 1; }
      if (l2 || (!l1 && ((float)m1.size[x] > (float)m2.size[y]))) { cx[0] = fc1; cx[1] = fc1 + 1; cy[0] = cy[1] = y; }
      else { cx[0] = cx[1] = x; cy[0] = fc2; cy[1] = fc2 + 1; }
      return 2;
    };
    for (int r = 0; r < n_exp; ++r) {
      const Entry& en = internal[r];
      uint32_t cx[2], cy[2];
      const int nc = split(en.x, en.y, cx, cy);
      for (int c = 0; c < nc; ++c) {
        if (o.quad) {
          uint32_t gx[2], gy[2];
          const int ng = split(cx[c], cy[c], gx, gy);
          for (int g = 0; g < ng; ++g) {
            const float d = bound_of(gx[g], gy[g]);
            ++n_tests;
            if (d < min_f) kids.push_back({gx[g], gy[g], d});
          }
        } else {
          const float d = bound_of(cx[c], cy[c]);
          ++n_tests;
          if (d < min_f) kids.push_back({cx[c], cy[c], d});
        }
      }
    }
    if (o.witness == 1) {  // applied after the round, like a warp would; children just computed are re-filtered
      min_f = std::min(min_f, wit_hi);
      std::vector<Entry> k2;
      for (auto& kd : kids) if (kd.b < min_f) k2.push_back(kd);
      kids.swap(k2);
    }
    if (getenv("SIM_TRACE")) printf("  bv round: sp %zu popped %d alive %d n_exp %d kids %zu raw %zu mid %zu exq %zu min_f %g\n", stack.size(), k, alive_n, n_exp, kids.size(), raw.size(), mid.size(), exq.size(), (double)min_f);
    st.bv_rounds += 1; st.bv_tests += n_tests; st.bv_lanes_pop += alive_n;
#pragma omp atomic
    g_hist[0][n_exp] += 1;
#pragma omp atomic
    g_hist[1][k] += 1;
#pragma omp atomic
    g_hist[2][alive_n] += 1;
#pragma omp atomic
    g_hist[3][std::min<size_t>(32, raw.size() / 4)] += 1;
    // ascending by the masked key like the kernel (stable enough for a work model)
    if (o.sort_bits > 0 && !kids.empty()) {
      float mn = 3e38f, mx = -3e38f;
      for (auto& kd : kids) { mn = fminf(mn, kd.b); mx = fmaxf(mx, kd.b); }
      const float sc = (float)((1 << o.sort_bits) - 1) / fmaxf(mx - mn, 1e-30f);
      std::stable_sort(kids.begin(), kids.end(), [&](const Entry& a, const Entry& b) {
        return (int)((a.b - mn) * sc) < (int)((b.b - mn) * sc);
      });
    } else
    std::stable_sort(kids.begin(), kids.end(), [](const Entry& a, const Entry& b) {
      uint32_t ka, kb; std::memcpy(&ka, &a.b, 4); std::memcpy(&kb, &b.b, 4);
      return (ka & ~31u) < (kb & ~31u);
    });
    for (int r = (int)kids.size() - 1; r >= 0; --r) stack.push_back(kids[r]);  // nearest on top
    max_sp = std::max(max_sp, stack.size());
    if (o.mode == 0) { /* raw pairs go straight to the screen queue */ }
  }
  st.max_sp = std::max<double>(st.max_sp, (double)max_sp);
  st.queries += 1;
  return min_d;
}

extern "C" void sim_run(long long n, const double* poses12, int nn1, const int32_t* fc1, const double* axis1, const double* ext1,
                        const double* rTo1, const double* rl1, const double* rr1, int nt1, const double* tri1, int nn2,
                        const int32_t* fc2, const double* axis2, const double* ext2, const double* rTo2, const double* rl2,
                        const double* rr2, int nt2, const double* tri2, const int* opt7, double* out_dist, double* stats13) {
  auto mk_model = [](int nn, const int32_t* fc, const double* axis, const double* ext, const double* rTo, const double* rl,
                     const double* rr, int nt, const double* tri, std::vector<double>& size) {
    SimModel m;
    m.n_nodes = nn; m.n_tris = nt; m.fc = fc; m.tri = tri;
    m.axis = axis; m.rTo = rTo; m.rl = rl; m.rr = rr;
    // witness: centroid of the subtree triangle whose centroid is nearest to the node's RSS centre (bottom-up candidates:
    // a node picks the better of its children's witnesses)
    m.wit.assign(3 * (size_t)nn, 0.0f);
    std::vector<double> wd(3 * (size_t)nn);
    for (int i = nn - 1; i >= 0; --i) {   // children have larger indices than their parent in this layout
      double c[3];
      for (int k = 0; k < 3; ++k) c[k] = rTo[3 * i + k] + 0.5 * rl[2 * i] * axis[9 * i + 3 * k] + 0.5 * rl[2 * i + 1] * axis[9 * i + 3 * k + 1];
      if (fc[i] < 0) {
        const int t = -(fc[i] + 1);
        for (int k = 0; k < 3; ++k) wd[3 * i + k] = (tri[9 * t + k] + tri[9 * t + 3 + k] + tri[9 * t + 6 + k]) / 3.0;
      } else {
        double best = 1e300; int bi = fc[i];
        for (int ch = fc[i]; ch <= fc[i] + 1; ++ch) {
          double d2 = 0; for (int k = 0; k < 3; ++k) d2 += (wd[3 * ch + k] - c[k]) * (wd[3 * ch + k] - c[k]);
          if (d2 < best) { best = d2; bi = ch; }
        }
        for (int k = 0; k < 3; ++k) wd[3 * i + k] = wd[3 * bi + k];
      }
    }
    for (size_t k = 0; k < wd.size(); ++k) m.wit[k] = (float)wd[k];
    size.resize(nn);
    m.rss32.resize(nn);
    for (int i = 0; i < nn; ++i) {
      size[i] = (ext[3 * i] * ext[3 * i] + ext[3 * i + 1] * ext[3 * i + 1]) + ext[3 * i + 2] * ext[3 * i + 2];
      pack_rss32(axis + 9 * (size_t)i, rTo + 3 * (size_t)i, rl + 2 * (size_t)i, rr[i], m.rss32[i]);
    }
    m.size = size.data();
    return m;
  };
  std::vector<double> sz1, sz2;
  SimModel m1 = mk_model(nn1, fc1, axis1, ext1, rTo1, rl1, rr1, nt1, tri1, sz1);
  SimModel m2 = mk_model(nn2, fc2, axis2, ext2, rTo2, rl2, rr2, nt2, tri2, sz2);
  Opt o{opt7[0], opt7[1], opt7[2], opt7[3], opt7[4], opt7[5], opt7[6], opt7[7], opt7[8], opt7[9], opt7[10], opt7[12], opt7[13], opt7[11]};
  Stats total{};
#pragma omp parallel
  {
    Stats st{};
#pragma omp for schedule(dynamic, 64)
    for (long long q = 0; q < n; ++q) { const double ih = (o.witness == 2) ? out_dist[q] * 1.000001 + 1e-9 : DBL_MAX; out_dist[q] = run_query(m1, m2, poses12 + 12 * q, o, st, ih); }
#pragma omp critical
    {
      double* a = reinterpret_cast<double*>(&total);
      const double* b = reinterpret_cast<const double*>(&st);
      for (int k = 0; k < 14; ++k) a[k] = (k == 12) ? std::max(a[k], b[k]) : a[k] + b[k];
    }
  }
  std::memcpy(stats13, &total, sizeof(double) * 14);
}

extern "C" void sim_hist(double* out, int reset) {
  std::memcpy(out, g_hist, sizeof(g_hist));
  if (reset) std::memset(g_hist, 0, sizeof(g_hist));
}

// validity sample of tri_closest_bounds_f32 against the exact routine: returns the number of violations
extern "C" long long sim_check_bounds(long long n, const double* S9, const double* T9, double* worst3, double* lohid) {
  long long bad = 0;
  double worst_lo = 0, worst_hi = 0, loose = 0;
  for (long long k = 0; k < n; ++k) {
    V3 S[3], T[3];
    for (int c = 0; c < 3; ++c) {
      S[c] = mk(S9[9 * k + 3 * c], S9[9 * k + 3 * c + 1], S9[9 * k + 3 * c + 2]);
      T[c] = mk(T9[9 * k + 3 * c], T9[9 * k + 3 * c + 1], T9[9 * k + 3 * c + 2]);
    }
    float s1[3], s2[3], t0[3], t1[3], t2[3];
    const V3 a = S[1] - S[0], c2 = S[2] - S[0], u0 = T[0] - S[0], u1 = T[1] - S[0], u2 = T[2] - S[0];
    s1[0] = (float)a.x; s1[1] = (float)a.y; s1[2] = (float)a.z;
    s2[0] = (float)c2.x; s2[1] = (float)c2.y; s2[2] = (float)c2.z;
    t0[0] = (float)u0.x; t0[1] = (float)u0.y; t0[2] = (float)u0.z;
    t1[0] = (float)u1.x; t1[1] = (float)u1.y; t1[2] = (float)u1.z;
    t2[0] = (float)u2.x; t2[1] = (float)u2.y; t2[2] = (float)u2.z;
    float lo, hi; bool trust;
    tri_closest_bounds_f32(s1, s2, t0, t1, t2, lo, hi, trust);
    V3 P, Q;
    const double d = tri_distance(S, T, P, Q);
    if (lohid) { lohid[3 * k] = lo; lohid[3 * k + 1] = trust ? hi : -1.0; lohid[3 * k + 2] = d; }
    if ((double)lo > d) { ++bad; worst_lo = std::max(worst_lo, (double)lo - d); }
    if (trust && (double)hi < d) { ++bad; worst_hi = std::max(worst_hi, d - (double)hi); }
    if (trust && d > 0) loose = std::max(loose, ((double)hi - (double)lo) / d);
  }
  worst3[0] = worst_lo; worst3[1] = worst_hi; worst3[2] = loose;
  return bad;
}
