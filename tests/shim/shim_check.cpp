// Compile- and run-check of include/fclgpu/fcl_shim.hpp against the mock FCL API surface (tests/shim/mock).
//   g++ -std=c++17 -Itests/shim/mock -Iinclude tests/shim/shim_check.cpp -Lfcl_b200/lib -lfclgpu -o shim_check
// Builds two procedural meshes with the library's own host builder, exposes them through the mock
// fcl::BVHModel<OBBRSS<double>>, runs the shim's batched collide / distance and compares every result with
// direct C-ABI calls on the same inputs.  argv[1] = "compile-only" skips the GPU part, "tolerance" adds the extension check.
#include <fclgpu/fcl_shim.hpp>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

#ifndef FCLGPU_HAVE_FCL
#error "the mock <fcl/fcl.h> / <Eigen/Core> must be on the include path"
#endif

using Model = fcl::BVHModel<fcl::OBBRSS<double>>;

static void sphere(double radius, int seg, int ring, std::vector<double>& v, std::vector<int32_t>& t) {
  const double pi = 3.14159265358979323846;
  for (int i = 1; i < ring; ++i)
    for (int j = 0; j < seg; ++j) {
      const double th = pi * i / ring, ph = 2 * pi * j / seg;
      v.insert(v.end(), {radius * std::sin(th) * std::cos(ph), radius * std::sin(th) * std::sin(ph) * 0.7, radius * std::cos(th)});
    }
  const int top = (int)v.size() / 3;
  v.insert(v.end(), {0, 0, radius});
  const int bot = top + 1;
  v.insert(v.end(), {0, 0, -radius});
  for (int j = 0; j < seg; ++j) {
    const int base = (ring - 2) * seg;
    t.insert(t.end(), {top, j, (j + 1) % seg});
    t.insert(t.end(), {bot, base + (j + 1) % seg, base + j});
  }
  for (int i = 0; i < ring - 2; ++i)
    for (int j = 0; j < seg; ++j) {
      const int a = i * seg + j, b = i * seg + (j + 1) % seg, c = (i + 1) * seg + j, d = (i + 1) * seg + (j + 1) % seg;
      t.insert(t.end(), {a, c, b});
      t.insert(t.end(), {b, c, d});
    }
}

// fill the mock model from the library's host builder (stands in for BVHModel::endModel())
static fclgpu_bvh* fill(Model& m, const std::vector<double>& v, const std::vector<int32_t>& t) {
  fclgpu_bvh* b = nullptr;
  fclgpu::check(fclgpu_bvh_build_obbrss(v.data(), (int32_t)v.size() / 3, t.data(), (int32_t)t.size() / 3, FCLGPU_SPLIT_METHOD_MEAN, &b));
  const int n = fclgpu_bvh_num_nodes(b), nt = fclgpu_bvh_num_tris(b);
  std::vector<int32_t> fc(n);
  std::vector<double> ax(9 * n), oT(3 * n), oe(3 * n), rT(3 * n), rl(2 * n), rr(n);
  fclgpu::check(fclgpu_bvh_get(b, fc.data(), ax.data(), oT.data(), oe.data(), rT.data(), rl.data(), rr.data(), nullptr));
  std::vector<int32_t> nf(n), nc(n);
  fclgpu::check(fclgpu_bvh_get_partition(b, nf.data(), nc.data(), nullptr, nullptr));
  m.bvs_.resize(n);
  for (int i = 0; i < n; ++i) {
    auto& node = m.bvs_[i];
    node.first_child = fc[i];
    node.first_primitive = nf[i];  // public BVNodeBase fields: the shim derives the (private) primitive_indices from them
    node.num_primitives = nc[i];
    for (int k = 0; k < 9; ++k) node.bv.obb.axis.m[k] = node.bv.rss.axis.m[k] = ax[9 * i + k];
    for (int k = 0; k < 3; ++k) {
      node.bv.obb.To[k] = oT[3 * i + k];
      node.bv.obb.extent[k] = oe[3 * i + k];
      node.bv.rss.To[k] = rT[3 * i + k];
    }
    node.bv.rss.l[0] = rl[2 * i];
    node.bv.rss.l[1] = rl[2 * i + 1];
    node.bv.rss.r = rr[i];
  }
  for (size_t i = 0; i < v.size() / 3; ++i) m.verts_.emplace_back(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
  for (int i = 0; i < nt; ++i) m.tris_.push_back(fcl::Triangle{{(size_t)t[3 * i], (size_t)t[3 * i + 1], (size_t)t[3 * i + 2]}});
  m.vertices = m.verts_.data();
  m.tri_indices = m.tris_.data();
  m.num_tris = nt;
  m.num_vertices = (int)m.verts_.size();
  return b;
}

int main(int argc, char** argv) {
  if (argc > 1 && std::string(argv[1]) == "compile-only") {
    std::printf("shim compiled against the mock FCL surface\n");
    return 0;
  }
  std::vector<double> v1, v2;
  std::vector<int32_t> t1, t2;
  sphere(1.0, 24, 20, v1, t1);
  sphere(0.6, 16, 12, v2, t2);
  Model m1, m2;
  fclgpu_bvh* b1 = fill(m1, v1, t1);
  fclgpu_bvh* b2 = fill(m2, v2, t2);
  fclgpu::DeviceModel d1(m1), d2(m2);

  const int n = 2000;
  std::vector<fcl::Transform3<double>> tf1(n, fcl::Transform3<double>::Identity()), tf2(n, fcl::Transform3<double>::Identity());
  std::vector<double> p2(12 * (size_t)n);
  unsigned long long s = 12345;
  auto rnd = [&]() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (double)(s >> 11) / 9007199254740992.0; };
  for (int i = 0; i < n; ++i) {
    const double a = 6.283185307179586 * rnd(), b = 6.283185307179586 * rnd();
    const double ca = std::cos(a), sa = std::sin(a), cb = std::cos(b), sb = std::sin(b);
    const double R[9] = {ca * cb, -sa, ca * sb, sa * cb, ca, sa * sb, -sb, 0, cb};
    tf2[i].setLinear(R);
    tf2[i].setTranslation(3.2 * (rnd() - 0.5), 3.2 * (rnd() - 0.5), 3.2 * (rnd() - 0.5));
    fclgpu_pose_from_colmajor4x4(tf2[i].m16, &p2[12 * (size_t)i]);
  }

  // direct C-ABI results on a model uploaded from the builder's own arrays
  fclgpu_model *g1 = nullptr, *g2 = nullptr;
  fclgpu::check(fclgpu_model_from_bvh(0, b1, &g1));
  fclgpu::check(fclgpu_model_from_bvh(0, b2, &g2));
  fclgpu_collision_request creq{20, 1, 0, 0, FCLGPU_CONTACT_FULL, 0};
  std::vector<int32_t> cnt(n);
  std::vector<int64_t> off(n + 1);
  std::vector<fclgpu_contact> pool(20 * (size_t)n);
  fclgpu::check(fclgpu_collide_batch_host(g1, g2, n, nullptr, p2.data(), &creq, cnt.data(), pool.data(), (int64_t)pool.size(), off.data(), nullptr, nullptr));
  fclgpu_distance_request dreq{1, 0, 0.0, 0.0};
  std::vector<double> dist(n), q1(3 * (size_t)n), q2(3 * (size_t)n);
  std::vector<int32_t> i1(n), i2(n);
  fclgpu::check(fclgpu_distance_batch_host(g1, g2, n, nullptr, p2.data(), &dreq, dist.data(), q1.data(), q2.data(), i1.data(), i2.data(), nullptr, nullptr));

  // continuous collision (both bodies translating): shim against the C ABI
  {
    std::vector<fcl::Transform3<double>> end1(tf1), end2(tf2);
    std::vector<double> pe1(12 * (size_t)n), pe2(12 * (size_t)n), pb1(12 * (size_t)n);
    for (int i = 0; i < n; ++i) {
      end1[i].setTranslation(0.3 * (rnd() - 0.5), 0.3 * (rnd() - 0.5), 0.3 * (rnd() - 0.5));
      end2[i].setTranslation(tf2[i].m16[12] + 2.0 * (rnd() - 0.5), tf2[i].m16[13] + 2.0 * (rnd() - 0.5), tf2[i].m16[14] + 2.0 * (rnd() - 0.5));
      fclgpu_pose_from_colmajor4x4(tf1[i].m16, &pb1[12 * (size_t)i]);
      fclgpu_pose_from_colmajor4x4(end1[i].m16, &pe1[12 * (size_t)i]);
      fclgpu_pose_from_colmajor4x4(end2[i].m16, &pe2[12 * (size_t)i]);
    }
    fclgpu_continuous_request rq{10, 0.0001, FCLGPU_CCDM_TRANS, 0, FCLGPU_CCDC_CONSERVATIVE_ADVANCEMENT};
    std::vector<int32_t> hit(n);
    std::vector<double> toc(n), c2(12 * (size_t)n);
    fclgpu::check(fclgpu_continuous_collide_batch_host(g1, g2, n, pb1.data(), pe1.data(), p2.data(), pe2.data(), &rq, hit.data(), toc.data(),
                                                       nullptr, c2.data(), nullptr));
    std::vector<fcl::ContinuousCollisionResult<double>> cc;
    fclgpu::continuous_collide(d1, tf1, end1, d2, tf2, end2,
                               fcl::ContinuousCollisionRequest<double>(10, 0.0001, fcl::CCDM_TRANS, fcl::GST_LIBCCD, fcl::CCDC_CONSERVATIVE_ADVANCEMENT), cc);
    long long moving_hits = 0;
    for (int i = 0; i < n; ++i) {
      if (cc[i].is_collide != (hit[i] != 0) || cc[i].time_of_contact != toc[i]) { std::printf("FAIL continuous %d\n", i); return 1; }
      if (hit[i] && toc[i] > 0) {
        ++moving_hits;
        double q[12];
        fclgpu_pose_from_colmajor4x4(cc[i].contact_tf2.m16, q);
        if (std::memcmp(q, &c2[12 * (size_t)i], sizeof q)) { std::printf("FAIL continuous contact pose %d\n", i); return 1; }
      }
    }
    if (moving_hits == 0) { std::printf("FAIL: no contact found in motion\n"); return 1; }
    bool threw = false;
    try {
      fclgpu::continuous_collide(d1, tf1, end1, d2, tf2, end2, fcl::ContinuousCollisionRequest<double>(), cc);  // CCDC_NAIVE: not built
    } catch (const std::exception&) {
      threw = true;
    }
    if (!threw) { std::printf("FAIL: unsupported continuous setting accepted\n"); return 1; }
    std::printf("shim continuous collide OK: %lld contacts found in motion\n", moving_hits);
  }

  // the same through the shim
  std::vector<fcl::CollisionResult<double>> cres;
  fclgpu::collide(d1, tf1, d2, tf2, fcl::CollisionRequest<double>(20, true), cres);
  std::vector<fcl::DistanceResult<double>> dres;
  fclgpu::distance(d1, tf1, d2, tf2, fcl::DistanceRequest<double>(true), dres);

  long long colliding = 0, contacts = 0;
  for (int i = 0; i < n; ++i) {
    if ((int)cres[i].numContacts() != cnt[i]) { std::printf("FAIL count %d: %zu vs %d\n", i, cres[i].numContacts(), cnt[i]); return 1; }
    colliding += cnt[i] > 0;
    for (int k = 0; k < cnt[i]; ++k) {
      const fclgpu_contact& c = pool[off[i] + k];
      const auto& r = cres[i].getContact(k);
      ++contacts;
      if (r.b1 != c.b1 || r.b2 != c.b2 || r.o1 != &m1 || r.o2 != &m2 || std::memcmp(r.pos.v, c.pos, 24) || std::memcmp(r.normal.v, c.normal, 24) ||
          r.penetration_depth != c.penetration_depth) { std::printf("FAIL contact %d/%d\n", i, k); return 1; }
    }
    if (dres[i].min_distance != dist[i] || dres[i].b1 != i1[i] || dres[i].b2 != i2[i] ||
        std::memcmp(dres[i].nearest_points[0].v, &q1[3 * (size_t)i], 24) || std::memcmp(dres[i].nearest_points[1].v, &q2[3 * (size_t)i], 24)) {
      std::printf("FAIL distance %d\n", i);
      return 1;
    }
  }
  // mesh <-> sphere through the shim vs the C ABI
  {
    fcl::Sphere<double> sph(0.5);
    std::vector<fcl::CollisionResult<double>> sres;
    fclgpu::collide(d1, tf1, sph, tf2, fcl::CollisionRequest<double>(20, true), sres);
    std::vector<int32_t> scnt(n);
    std::vector<int64_t> soff(n + 1);
    std::vector<fclgpu_contact> spool(20 * (size_t)n);
    fclgpu::check(fclgpu_collide_mesh_sphere_batch_host(g1, 0.5, n, nullptr, p2.data(), &creq, scnt.data(), spool.data(), (int64_t)spool.size(), soff.data(), nullptr, nullptr));
    long long shits = 0;
    for (int i = 0; i < n; ++i) {
      if ((int)sres[i].numContacts() != scnt[i]) { std::printf("FAIL sphere count %d\n", i); return 1; }
      shits += scnt[i] > 0;
      for (int k = 0; k < scnt[i]; ++k) {
        const fclgpu_contact& c = spool[soff[i] + k];
        const auto& r = sres[i].getContact(k);
        if (r.b1 != c.b1 || r.b2 != fcl::Contact<double>::NONE || r.o2 != &sph || std::memcmp(r.pos.v, c.pos, 24) || r.penetration_depth != c.penetration_depth) {
          std::printf("FAIL sphere contact %d/%d\n", i, k);
          return 1;
        }
      }
    }
    if (shits == 0) { std::printf("FAIL: no sphere hits\n"); return 1; }
    std::printf("shim sphere OK: %lld colliding\n", shits);
    // mesh <-> sphere distance through the shim vs the C ABI
    std::vector<fcl::DistanceResult<double>> sd;
    fclgpu::distance(d1, tf1, sph, tf2, fcl::DistanceRequest<double>(false), sd);
    std::vector<double> sdist(n), sa(3 * (size_t)n), sb(3 * (size_t)n);
    std::vector<int32_t> sb1(n);
    fclgpu_distance_request sreq{1, 0, 0.0, 0.0};
    fclgpu::check(fclgpu_distance_mesh_sphere_batch_host(g1, 0.5, n, nullptr, p2.data(), &sreq, sdist.data(), sa.data(), sb.data(), sb1.data(), nullptr, nullptr, nullptr));
    long long separated = 0;
    for (int i = 0; i < n; ++i) {
      if (sd[i].min_distance != sdist[i] || sd[i].b1 != sb1[i] || sd[i].b2 != fcl::DistanceResult<double>::NONE || sd[i].o2 != &sph ||
          std::memcmp(sd[i].nearest_points[0].v, &sa[3 * (size_t)i], 24) || std::memcmp(sd[i].nearest_points[1].v, &sb[3 * (size_t)i], 24)) {
        std::printf("FAIL sphere distance %d\n", i);
        return 1;
      }
      separated += sdist[i] > 0;
      if ((sdist[i] > 0) != (scnt[i] == 0)) { std::printf("FAIL sphere distance vs collide verdict %d\n", i); return 1; }
    }
    if (separated == 0) { std::printf("FAIL: no separated sphere\n"); return 1; }
    std::printf("shim sphere distance OK: %lld separated\n", separated);
    // tolerance verification through the shim vs the plain distances above (argv[1] = "tolerance": run by the last GPU test file)
    if (argc > 1 && std::string(argv[1]) == "tolerance") {
    std::vector<char> within;
    fclgpu::within_tolerance(d1, tf1, d2, tf2, 0.25, within);
    long long n_within = 0;
    for (int i = 0; i < n; ++i) {
      if ((within[i] != 0) != (dist[i] <= 0.25)) { std::printf("FAIL within tolerance %d\n", i); return 1; }
      n_within += within[i] != 0;
    }
    if (n_within == 0 || n_within == n) { std::printf("FAIL: degenerate tolerance sample\n"); return 1; }
    std::printf("shim tolerance OK: %lld within 0.25\n", n_within);
    }
  }
  // ---- the dispatch cells: install(), then reach them through the (mock) fcl::collide / fcl::distance look-up ----
  {
    using Solver = fcl::detail::GJKSolver_libccd<double>;
    Solver solver;
    if (fcl::getCollisionFunctionLookTable<Solver>().collision_matrix[fcl::BV_OBBRSS][fcl::BV_OBBRSS] != nullptr) { std::printf("FAIL: mock table not empty\n"); return 1; }
    fclgpu::install<Solver>();
    // exact CollisionFunc / DistanceFunc types (collision_func_matrix.h:67-78, distance_func_matrix.h:65-76)
    fcl::detail::CollisionFunctionMatrix<Solver>::CollisionFunc cf = &fclgpu::collide_cell<Solver>;
    fcl::detail::DistanceFunctionMatrix<Solver>::DistanceFunc df = &fclgpu::distance_cell<Solver>;
    (void)cf; (void)df;
    long long cells = 0, accumulated = 0, early = 0;
    for (int i = 0; i < n && cells < 40; ++i) {
      if (cnt[i] < 3) continue;  // queries with several contacts exercise the budget
      ++cells;
      const fcl::CollisionGeometry<double>*g1p = &m1, *g2p = &m2;
      fcl::CollisionResult<double> res;
      const std::size_t got = fcl::collide(g1p, tf1[i], g2p, tf2[i], &solver, fcl::CollisionRequest<double>(20, true), res);
      if ((int)got != cnt[i] || (int)res.numContacts() != cnt[i]) { std::printf("FAIL cell count %d: %zu vs %d\n", i, got, cnt[i]); return 1; }
      for (int k = 0; k < cnt[i]; ++k) {
        const fclgpu_contact& c = pool[off[i] + k];
        const auto& r = res.getContact(k);
        if (r.b1 != c.b1 || r.b2 != c.b2 || std::memcmp(r.pos.v, c.pos, 24) || r.penetration_depth != c.penetration_depth) { std::printf("FAIL cell contact %d/%d\n", i, k); return 1; }
      }
      // accumulation into a NON-EMPTY result: two contacts present, budget 2 + 1 -> exactly one more (the query's first)
      fcl::CollisionResult<double> acc;
      acc.addContact(res.getContact(0));
      acc.addContact(res.getContact(1));
      const std::size_t got2 = fcl::collide(g1p, tf1[i], g2p, tf2[i], &solver, fcl::CollisionRequest<double>(3, true), acc);
      if (got2 != 3 || acc.numContacts() != 3 || acc.getContact(2).b1 != pool[off[i]].b1 || acc.getContact(2).b2 != pool[off[i]].b2 ||
          std::memcmp(acc.getContact(2).pos.v, pool[off[i]].pos, 24)) { std::printf("FAIL cell accumulation %d\n", i); return 1; }
      ++accumulated;
      // isSatisfied early return (collision_func_matrix-inl.h:580): the result already holds num_max_contacts -> no GPU work
      const long long launches = fclgpu_launch_count();
      const std::size_t got3 = fcl::collide(g1p, tf1[i], g2p, tf2[i], &solver, fcl::CollisionRequest<double>(3, true), acc);
      if (got3 != 3 || fclgpu_launch_count() != launches) { std::printf("FAIL cell early return %d\n", i); return 1; }
      ++early;
      // distance cell: same value as the batch; a result that already holds a smaller distance keeps it; <= 0 returns early
      fcl::DistanceResult<double> dr;
      const double dd = fcl::distance(g1p, tf1[i], g2p, tf2[i], &solver, fcl::DistanceRequest<double>(true), dr);
      if (dd != dist[i] || dr.b1 != i1[i] || dr.b2 != i2[i]) { std::printf("FAIL distance cell %d\n", i); return 1; }
    }
    int far = -1;
    for (int i = 0; i < n; ++i) if (dist[i] > 0.1) { far = i; break; }
    if (far >= 0) {
      const fcl::CollisionGeometry<double>*g1p = &m1, *g2p = &m2;
      fcl::DistanceResult<double> dr;
      dr.min_distance = dist[far] * 0.5;  // already smaller: the cell must not overwrite it (DistanceResult::update)
      if (fcl::distance(g1p, tf1[far], g2p, tf2[far], &solver, fcl::DistanceRequest<double>(true), dr) != dist[far] * 0.5) { std::printf("FAIL distance cell update\n"); return 1; }
      dr.min_distance = 0.0;
      const long long launches = fclgpu_launch_count();
      if (fcl::distance(g1p, tf1[far], g2p, tf2[far], &solver, fcl::DistanceRequest<double>(true), dr) != 0.0 || fclgpu_launch_count() != launches) { std::printf("FAIL distance cell early return\n"); return 1; }
      // sphere cells, both argument orders ((OT_GEOM, OT_BVH) is dispatched with swapped arguments, collision-inl.h:124-133)
      fcl::Sphere<double> sph(0.5);
      const fcl::CollisionGeometry<double>* sp = &sph;
      fcl::DistanceResult<double> s1, s2;
      const double a = fcl::distance(g1p, tf1[far], sp, tf2[far], &solver, fcl::DistanceRequest<double>(true), s1);
      const double b = fcl::distance(sp, tf2[far], g1p, tf1[far], &solver, fcl::DistanceRequest<double>(true), s2);
      if (a != b || s1.b1 != s2.b1 || s1.b2 != fcl::DistanceResult<double>::NONE) { std::printf("FAIL sphere cells\n"); return 1; }
    }
    {  // halfspace / plane cells against the C ABI (single query each)
      const fcl::CollisionGeometry<double>* g1p = &m1;
      fcl::Halfspace<double> hs(fcl::Vector3<double>(0.0, 0.0, 1.0), 0.3);
      fcl::Plane<double> pl(fcl::Vector3<double>(0.0, 1.0, 0.0), -0.2);
      const double nh[3] = {0, 0, 1}, np_[3] = {0, 1, 0};
      fclgpu_collision_request rq{1000, 1, 0, 0, FCLGPU_CONTACT_FULL, 0};
      for (int which = 0; which < 2; ++which) {
        int32_t c = 0;
        std::vector<fclgpu_contact> pc(4000);
        int64_t o2[2];
        double idp[12];
        fclgpu_pose_from_colmajor4x4(tf1[0].m16, idp);
        fclgpu::check(fclgpu_collide_mesh_plane_batch_host(g1, which ? FCLGPU_SHAPE_PLANE : FCLGPU_SHAPE_HALFSPACE, which ? np_ : nh, which ? -0.2 : 0.3,
                                                           1, idp, idp, &rq, &c, pc.data(), (int64_t)pc.size(), o2, nullptr, nullptr));
        fcl::CollisionResult<double> res;
        const fcl::CollisionGeometry<double>* sh = which ? (const fcl::CollisionGeometry<double>*)&pl : (const fcl::CollisionGeometry<double>*)&hs;
        const std::size_t got = fcl::collide(g1p, tf1[0], sh, tf1[0], &solver, fcl::CollisionRequest<double>(1000, true), res);
        if ((int)got != c || c == 0) { std::printf("FAIL plane-like cell %d: %zu vs %d\n", which, got, c); return 1; }
        for (int k = 0; k < c; ++k)
          if (res.getContact(k).b1 != pc[k].b1 || res.getContact(k).penetration_depth != pc[k].penetration_depth) { std::printf("FAIL plane-like contact %d\n", k); return 1; }
      }
      std::printf("shim halfspace / plane cells OK\n");
    }
    if (cells == 0) { std::printf("FAIL: no multi-contact query for the cell check\n"); return 1; }
    std::printf("shim cells OK: %lld single queries through the look-up table, %lld accumulations, %lld early returns\n", cells, accumulated, early);
    // refit through the shim: the partition was derived from the public node fields
    if (!d1.refit_ready()) { std::printf("FAIL: partition not derived\n"); return 1; }
    for (auto& v : m1.verts_) v = fcl::Vector3<double>(v[0] * 1.01, v[1] * 1.01, v[2] * 1.01);
    d1.refit_topdown();
    fclgpu::check(fclgpu_bvh_refit_topdown(b1, &m1.verts_[0].v[0], m1.num_vertices));
    std::vector<double> devT(3 * (size_t)m1.getNumBVs()), hostT(3 * (size_t)m1.getNumBVs());
    fclgpu::check(fclgpu_model_download(d1.handle(), nullptr, devT.data(), nullptr, nullptr, nullptr, nullptr, nullptr));
    fclgpu::check(fclgpu_bvh_get(b1, nullptr, nullptr, hostT.data(), nullptr, nullptr, nullptr, nullptr, nullptr));
    if (std::memcmp(devT.data(), hostT.data(), devT.size() * 8)) { std::printf("FAIL: shim refit differs from the host refit\n"); return 1; }
    std::printf("shim refit OK\n");
    // bottom-up refit (FCL's default endReplaceModel()) through the shim against the host model's
    d1.refit_bottomup();
    fclgpu::check(fclgpu_bvh_refit_bottomup(b1, &m1.verts_[0].v[0], m1.num_vertices));
    std::vector<double> devA(9 * (size_t)m1.getNumBVs()), hostA(9 * (size_t)m1.getNumBVs());
    fclgpu::check(fclgpu_model_download_rss_axis(d1.handle(), devA.data()));
    if (fclgpu_bvh_get_rss_axis(b1, hostA.data()) != 1) { std::printf("FAIL: host model has no separate RSS axes after the bottom-up refit\n"); return 1; }
    fclgpu::check(fclgpu_model_download(d1.handle(), nullptr, devT.data(), nullptr, nullptr, nullptr, nullptr, nullptr));
    fclgpu::check(fclgpu_bvh_get(b1, nullptr, nullptr, hostT.data(), nullptr, nullptr, nullptr, nullptr, nullptr));
    if (std::memcmp(devA.data(), hostA.data(), devA.size() * 8) || std::memcmp(devT.data(), hostT.data(), devT.size() * 8)) {
      std::printf("FAIL: shim bottom-up refit differs from the host refit\n");
      return 1;
    }
    std::printf("shim bottom-up refit OK\n");
    fclgpu::uninstall<Solver>();
    if (fcl::getCollisionFunctionLookTable<Solver>().collision_matrix[fcl::BV_OBBRSS][fcl::BV_OBBRSS] != nullptr) { std::printf("FAIL: uninstall\n"); return 1; }
  }
  if (colliding < n / 20 || colliding > n - n / 20) { std::printf("FAIL: degenerate pose sample (%lld colliding)\n", colliding); return 1; }
  std::printf("shim OK: %d queries, %lld colliding, %lld contacts compared, distances identical\n", n, colliding, contacts);
  {  // what the library kept between the calls goes back to the driver; a second call has nothing left to release
    const std::int64_t released = fclgpu::release_device_memory(0, /*models=*/true);
    if (released <= 0 || fclgpu::release_device_memory(0) != 0) { std::printf("FAIL: release_device_memory (%lld)\n", (long long)released); return 1; }
    std::printf("shim release OK: %lld workspace bytes\n", (long long)released);
  }
  fclgpu_model_destroy(g1);
  fclgpu_model_destroy(g2);
  fclgpu_bvh_destroy(b1);
  fclgpu_bvh_destroy(b2);
  return 0;
}
