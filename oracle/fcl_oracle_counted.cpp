// fcl_oracle_counted.cpp — CPU ORACLE (test infrastructure, NOT product code): EXECUTED-operation counters.
//
// SURVEY.md 8(d) defines the arithmetic side of the roofline from the operations the reference's SEQUENTIAL traversal
// actually executes: flops = 63 + n_bv x F_bv + n_leaf x F_leaf with F_* the executed mul + add + cmp counts (not the
// all-early-outs-fail upper bounds).  This translation unit compiles the very same oracle sources a second time with
// `double` replaced by a counting scalar, inside another namespace, so the numbers come from the restatement itself
// (same control flow, same early exits) and need no hand-maintained formulas.  Results are bit-identical to the plain
// build (the scalar wraps a double and forwards every operation); tests/test_oracle_counters.py checks that.
//
// Counted: multiplications, additions/subtractions (incl. unary minus), comparisons, divisions, square roots of the
// scalar type.  Not counted: integer work, loads/stores, fabs/min/max selects (their compares are counted).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <queue>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

namespace oracle_counted {

struct OpCounts {
  long long mul = 0, add = 0, cmp = 0, div = 0, sqrt = 0;
};
thread_local OpCounts g_ops;

typedef double plain;

struct Real {
  plain v;
  Real() = default;
  template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>
  Real(T x) : v((plain)x) {}
  explicit operator plain() const { return v; }
  explicit operator int() const { return (int)v; }
  explicit operator float() const { return (float)v; }
  Real operator-() const { ++g_ops.add; return Real(-v); }
  Real& operator+=(const Real& o) { ++g_ops.add; v += o.v; return *this; }
  Real& operator-=(const Real& o) { ++g_ops.add; v -= o.v; return *this; }
  Real& operator*=(const Real& o) { ++g_ops.mul; v *= o.v; return *this; }
  Real& operator/=(const Real& o) { ++g_ops.div; v /= o.v; return *this; }
};
#define ORC_ARITH(T) typename std::enable_if<std::is_arithmetic<T>::value, int>::type = 0
#define ORC_BIN(op, field)                                                                          \
  inline Real operator op(const Real& a, const Real& b) { ++g_ops.field; return Real(a.v op b.v); } \
  template <class T, ORC_ARITH(T)>                                                                  \
  inline Real operator op(const Real& a, T b) { ++g_ops.field; return Real(a.v op (plain)b); }      \
  template <class T, ORC_ARITH(T)>                                                                  \
  inline Real operator op(T a, const Real& b) { ++g_ops.field; return Real((plain)a op b.v); }
ORC_BIN(+, add)
ORC_BIN(-, add)
ORC_BIN(*, mul)
ORC_BIN(/, div)
#define ORC_CMP(op)                                                                        \
  inline bool operator op(const Real& a, const Real& b) { ++g_ops.cmp; return a.v op b.v; } \
  template <class T, ORC_ARITH(T)>                                                          \
  inline bool operator op(const Real& a, T b) { ++g_ops.cmp; return a.v op (plain)b; }      \
  template <class T, ORC_ARITH(T)>                                                          \
  inline bool operator op(T a, const Real& b) { ++g_ops.cmp; return (plain)a op b.v; }
ORC_CMP(<)
ORC_CMP(>)
ORC_CMP(<=)
ORC_CMP(>=)
ORC_CMP(==)
ORC_CMP(!=)

}  // namespace oracle_counted

namespace std {
inline oracle_counted::Real sqrt(const oracle_counted::Real& x) { ++oracle_counted::g_ops.sqrt; return oracle_counted::Real(std::sqrt(x.v)); }
inline oracle_counted::Real fabs(const oracle_counted::Real& x) { return oracle_counted::Real(std::fabs(x.v)); }
inline oracle_counted::Real abs(const oracle_counted::Real& x) { return oracle_counted::Real(std::fabs(x.v)); }
inline bool isnan(const oracle_counted::Real& x) { return std::isnan(x.v); }
template <>
struct numeric_limits<oracle_counted::Real> {
  static constexpr bool is_specialized = true;
  static oracle_counted::Real max() { return oracle_counted::Real(numeric_limits<double>::max()); }
  static oracle_counted::Real min() { return oracle_counted::Real(numeric_limits<double>::min()); }
  static oracle_counted::Real lowest() { return oracle_counted::Real(numeric_limits<double>::lowest()); }
  static oracle_counted::Real epsilon() { return oracle_counted::Real(numeric_limits<double>::epsilon()); }
  static oracle_counted::Real quiet_NaN() { return oracle_counted::Real(numeric_limits<double>::quiet_NaN()); }
  static oracle_counted::Real infinity() { return oracle_counted::Real(numeric_limits<double>::infinity()); }
};
}  // namespace std

// ---- the oracle, recompiled over the counting scalar --------------------------------------------------------------
#define double oracle_counted::Real
#define oracle oracle_cnt
#include "fcl_oracle_math.cpp"
#include "fcl_oracle_bvh.cpp"
#undef oracle
#undef double

namespace {
using oracle_counted::OpCounts;
using oracle_counted::g_ops;

oracle_cnt::Pose pose_from(const double* p) {
  oracle_cnt::Pose q;
  if (!p) {
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) q.R.m[i][j] = (i == j) ? 1.0 : 0.0;
    q.t = oracle_cnt::Vec3{{0.0, 0.0, 0.0}};
    return q;
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) q.R.m[i][j] = p[3 * i + j];
  q.t = oracle_cnt::Vec3{{p[9], p[10], p[11]}};
  return q;
}

oracle_cnt::Model* make_model(const double* verts, int nv, const int32_t* tris, int nt) {
  std::vector<oracle_cnt::Vec3> pts(nv);
  std::vector<oracle_cnt::Tri> ts(nt);
  for (int i = 0; i < nv; ++i) pts[i] = oracle_cnt::Vec3{{verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]}};
  for (int i = 0; i < nt; ++i) ts[i] = oracle_cnt::Tri{{tris[3 * i], tris[3 * i + 1], tris[3 * i + 2]}};
  auto* m = new oracle_cnt::Model;
  oracle_cnt::build_model(*m, pts, ts, oracle_cnt::SPLIT_MEAN);
  return m;
}

template <class F>
void parallel_for(long long n, int nthreads, F f) {
  std::atomic<long long> next{0};
  std::vector<std::thread> th;
  for (int t = 0; t < std::max(1, nthreads); ++t)
    th.emplace_back([&] {
      while (true) {
        const long long s = next.fetch_add(64);
        if (s >= n) break;
        for (long long i = s; i < std::min(n, s + 64); ++i) f(i);
      }
    });
  for (auto& x : th) x.join();
}
}  // namespace

extern "C" {

void* orcc_model(const double* verts, int nv, const int32_t* tris, int nt) { return make_model(verts, nv, tris, nt); }
void orcc_model_free(void* m) { delete (oracle_cnt::Model*)m; }

// Per query: ops[i][0..4] = executed mul, add, cmp, div, sqrt of the WHOLE query (pose set-up, every BV test, every leaf
// test, contact generation); nbv/nleaf = the traversal's test counters; value[i] = numContacts (collide) or the
// minimum distance (distance), for the check against the plain build.
void orcc_collide(void* m1, void* m2, const double* tf1, const double* tf2, long long n, long long num_max_contacts,
                  int enable_contact, int nthreads, long long* ops5, long long* nbv, long long* nleaf, double* value) {
  parallel_for(n, nthreads, [&](long long i) {
    g_ops = OpCounts();
    std::vector<oracle_cnt::Contact> out;
    oracle_cnt::CollideStats st;
    const size_t c = oracle_cnt::collide(*(oracle_cnt::Model*)m1, pose_from(tf1 ? tf1 + 12 * i : nullptr), *(oracle_cnt::Model*)m2,
                                         pose_from(tf2 ? tf2 + 12 * i : nullptr), (size_t)num_max_contacts, enable_contact != 0,
                                         out, &st);
    const OpCounts o = g_ops;
    ops5[5 * i] = o.mul; ops5[5 * i + 1] = o.add; ops5[5 * i + 2] = o.cmp; ops5[5 * i + 3] = o.div; ops5[5 * i + 4] = o.sqrt;
    nbv[i] = st.n_bv;
    nleaf[i] = st.n_leaf;
    value[i] = (double)c;
  });
}

void orcc_distance(void* m1, void* m2, const double* tf1, const double* tf2, long long n, int enable_nearest_points,
                   int nthreads, long long* ops5, long long* nbv, long long* nleaf, double* value) {
  parallel_for(n, nthreads, [&](long long i) {
    g_ops = OpCounts();
    oracle_cnt::DistanceOut out;
    oracle_cnt::CollideStats st;
    const oracle_counted::Real d = oracle_cnt::distance(*(oracle_cnt::Model*)m1, pose_from(tf1 ? tf1 + 12 * i : nullptr),
                                                        *(oracle_cnt::Model*)m2, pose_from(tf2 ? tf2 + 12 * i : nullptr),
                                                        enable_nearest_points != 0, out, 2, &st);
    const OpCounts o = g_ops;
    ops5[5 * i] = o.mul; ops5[5 * i + 1] = o.add; ops5[5 * i + 2] = o.cmp; ops5[5 * i + 3] = o.div; ops5[5 * i + 4] = o.sqrt;
    nbv[i] = st.n_bv;
    nleaf[i] = st.n_leaf;
    value[i] = d.v;
  });
}

}  // extern "C"
