"""The product's device math header, compiled for the host, must agree with the oracle
bit for bit (CPU only; the same header is what the CUDA kernels inline)."""
import ctypes as C
import shutil

import numpy as np
import pytest

from fcl_b200.poses import random_poses

pytestmark = pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not available")


@pytest.fixture(scope="module")
def hm():
    from tests import hostcheck

    return hostcheck


def _rel_pose(P):
    """(R, T) of identity model2 in posed model1's frame, same arithmetic as collide setup."""
    n = len(P)
    R1 = P[:, :9].reshape(n, 3, 3)
    t1 = P[:, 9:]
    out = np.empty((n, 12))
    out[:, :9] = np.transpose(R1, (0, 2, 1)).reshape(n, 9)  # R1^T * I
    d = 0.0 - t1
    RT = np.transpose(R1, (0, 2, 1))
    out[:, 9:] = np.stack([(RT[:, i, 0] * d[:, 0] + RT[:, i, 1] * d[:, 1]) + RT[:, i, 2] * d[:, 2] for i in range(3)], 1)
    return out


def test_obb_and_rss_pairs_match_oracle(hm, oracle, oracle_env_rob):
    env, rob = oracle_env_rob
    a1, a2 = env.arrays(), rob.arrays()
    rng = np.random.default_rng(11)
    n = 200000
    rel = _rel_pose(random_poses(n, seed=9))
    # bias towards deep nodes being near each other: also shrink translations for half of them
    rel[: n // 2, 9:] *= rng.uniform(0, 0.3, size=(n // 2, 1))
    i1 = rng.integers(0, env.num_bvs, n).astype(np.int32)
    i2 = rng.integers(0, rob.num_bvs, n).astype(np.int32)
    got = np.empty(n, np.int32)
    L = hm.lib()
    L.hm_obb_pairs(n, hm.dptr(rel), hm.iptr(i1), hm.iptr(i2), hm.dptr(a1["axis"]), hm.dptr(a1["obb_To"]),
                   hm.dptr(a1["obb_ext"]), hm.dptr(a2["axis"]), hm.dptr(a2["obb_To"]), hm.dptr(a2["obb_ext"]), hm.iptr(got))
    dist = np.empty(n)
    L.hm_rss_pairs(n, hm.dptr(rel), hm.iptr(i1), hm.iptr(i2), hm.dptr(a1["axis"]), hm.dptr(a1["rss_To"]),
                   hm.dptr(a1["rss_l"]), hm.dptr(a1["rss_r"]), hm.dptr(a2["axis"]), hm.dptr(a2["rss_To"]),
                   hm.dptr(a2["rss_l"]), hm.dptr(a2["rss_r"]), hm.dptr(dist))
    step = 7
    n_dis = 0
    n_pos = 0
    for k in range(0, n, step):
        R0, T0 = rel[k, :9], rel[k, 9:]
        i, j = i1[k], i2[k]
        ov = oracle.obb_overlap(R0, T0, a1["axis"][i], a1["obb_To"][i], a1["obb_ext"][i], a2["axis"][j], a2["obb_To"][j], a2["obb_ext"][j])
        assert bool(got[k]) == (not ov)
        n_dis += got[k]
        d = oracle.rss_distance(R0, T0, a1["axis"][i], a1["rss_To"][i], a1["rss_l"][i], a1["rss_r"][i],
                                a2["axis"][j], a2["rss_To"][j], a2["rss_l"][j], a2["rss_r"][j])
        assert d == dist[k], (k, d, dist[k])
        n_pos += d > 0
    assert 0 < n_dis < n // step and n_pos > 100


def test_rect_distance_random_configs_match_oracle(hm, oracle):
    """Synthetic rectangles in general position exercise all 16 edge cases + the face fallback."""
    rng = np.random.default_rng(5)
    from fcl_b200.poses import euler_to_matrix

    L = hm.lib()
    n = 60000
    ang = rng.uniform(0, 2 * np.pi, size=(n, 3))
    # a third of the rotations nearly axis aligned (exercises the 1e-7 parallel thresholds)
    ang[: n // 3] = np.round(ang[: n // 3] / (np.pi / 2)) * (np.pi / 2) + rng.normal(0, 1e-8, size=(n // 3, 3))
    R = euler_to_matrix(ang[:, 0], ang[:, 1], ang[:, 2]).reshape(n, 9)
    T = rng.uniform(-3, 3, size=(n, 3))
    a = rng.uniform(0, 2, size=(n, 2))
    b = rng.uniform(0, 2, size=(n, 2))
    a[::17] = 0.0  # degenerate (segment / point) rectangles
    vals = set()
    for k in range(n):
        g = L.hm_rect_distance(hm.dptr(R[k]), hm.dptr(T[k]), hm.dptr(a[k]), hm.dptr(b[k]))
        o = oracle.rect_distance(R[k], T[k], a[k], b[k])
        assert g == o, (k, g, o)
        vals.add(g > 0)
    assert vals == {True, False}


def _tri_pairs(rng, n, spread):
    P = rng.normal(0, 1, size=(n, 3, 3))
    Q = rng.normal(0, 1, size=(n, 3, 3)) + rng.normal(0, spread, size=(n, 1, 3))
    return P.reshape(n, 9), Q.reshape(n, 9)


def test_tri_intersect_and_contacts_match_oracle(hm, oracle):
    rng = np.random.default_rng(3)
    L = hm.lib()
    n = 40000
    P, Q = _tri_pairs(rng, n, 0.8)
    # shared-vertex / coplanar cases (touching counts as intersecting)
    Q[:2000, :3] = P[:2000, :3]
    Q[2000:3000, 2::3] = 0.0
    P[2000:3000, 2::3] = 0.0
    poses = random_poses(n, seed=21)
    poses[:, 9:] *= 1e-4
    poses[: n // 2, :9] = np.eye(3).reshape(9)
    hits = 0
    for k in range(n):
        nc = C.c_uint32(0)
        c6, dep, nrm = np.zeros(6), np.zeros(1), np.zeros(3)
        g = L.hm_tri_intersect(hm.dptr(P[k]), hm.dptr(Q[k]), hm.dptr(poses[k]), 1, C.byref(nc), hm.dptr(c6), hm.dptr(dep), hm.dptr(nrm))
        o = oracle.tri_intersect(P[k], Q[k], poses[k, :9], poses[k, 9:], want_contacts=True)
        assert bool(g) == o[0]
        if g:
            hits += 1
            assert nc.value == o[1]
            assert dep[0] == o[3]
            assert nrm.tobytes() == o[4].tobytes()
            assert c6[: 3 * nc.value].tobytes() == o[2].reshape(-1)[: 3 * nc.value].tobytes()
    assert n // 20 < hits < n


def test_tri_distance_matches_oracle(hm, oracle):
    rng = np.random.default_rng(4)
    L = hm.lib()
    n = 40000
    S, T = _tri_pairs(rng, n, 2.5)
    T[:1500, :3] = S[:1500, :3]          # shared vertex
    S[1500:2500, 3:6] = S[1500:2500, :3]  # degenerate (zero-length edge -> NaN guards)
    T[2500:3500] = S[2500:3500] + np.tile(rng.normal(0, 2, size=(1000, 1, 3)), (1, 3, 1)).reshape(1000, 9)  # parallel
    zero = 0
    for k in range(n):
        Pg, Qg = np.zeros(3), np.zeros(3)
        g = L.hm_tri_distance(hm.dptr(S[k]), hm.dptr(T[k]), hm.dptr(Pg), hm.dptr(Qg))
        o, Po, Qo = oracle.tri_distance(S[k], T[k])
        assert g == o or (np.isnan(g) and np.isnan(o)), (k, g, o)
        assert Pg.tobytes() == Po.tobytes() and Qg.tobytes() == Qo.tobytes(), k
        zero += g == 0
    assert 0 < zero < n


def _sphere_cases(rng, n):
    """Triangles of size ~1 and sphere centres placed over faces, edges, vertices, far away and on the plane."""
    T = rng.normal(0, 1, size=(n, 9))
    w = rng.dirichlet([1, 1, 1], size=n)
    w[: n // 4] = rng.dirichlet([0.05, 0.05, 0.05], size=n // 4)  # close to vertices / edges
    on = np.einsum("nk,nkj->nj", w, T.reshape(n, 3, 3))
    nrm = np.cross(T[:, 3:6] - T[:, :3], T[:, 6:] - T[:, :3])
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    h = rng.normal(0, 1.0, size=(n, 1))
    h[n // 2: n // 2 + n // 8] = 0.0  # centre in the triangle's plane
    c = on + nrm * h + rng.normal(0, 0.7, size=(n, 3)) * (rng.random((n, 1)) < 0.6)
    radius = rng.choice([0.0, 0.05, 0.5, 1.5], size=n)
    T[n - 200:, 3:6] = T[n - 200:, :3]  # zero-area triangles
    return T, c, radius


def test_sphere_triangle_routines_match_oracle(hm, oracle):
    rng = np.random.default_rng(21)
    L = hm.lib()
    n = 40000
    T, c, radius = _sphere_cases(rng, n)
    hits = seps = 0
    for k in range(n):
        out = np.zeros(7)
        g = L.hm_sphere_tri_intersect(hm.dptr(c[k]), float(radius[k]), hm.dptr(T[k]), hm.dptr(out))
        hit, cp, depth, nrm = oracle.sphere_tri_intersect(c[k], radius[k], T[k])
        assert bool(g) == hit, k
        if hit:
            assert out.tobytes() == np.concatenate([cp, [depth], nrm]).tobytes(), k
        hits += hit
        out = np.zeros(7)
        g = L.hm_sphere_tri_distance(hm.dptr(c[k]), float(radius[k]), hm.dptr(T[k]), hm.dptr(out))
        ok, d, ps, pt = oracle.sphere_tri_distance(c[k], radius[k], T[k])
        assert bool(g) == ok, k
        if ok:
            assert out.tobytes() == np.concatenate([[d], ps, pt]).tobytes(), k
            if k < n - 200:  # the closed form really is the point-triangle distance (checked against sampling-free geometry)
                assert abs(np.linalg.norm(pt - c[k]) - radius[k] - d) < 1e-9 * max(1.0, d)
        seps += ok
    assert 0.05 * n < hits < 0.95 * n and 0.05 * n < seps < 0.95 * n


def _host_mesh_sphere(hm, model, verts, tris, radius, M, S, bound32=False):
    a = model.arrays()
    tri9 = np.ascontiguousarray(np.asarray(verts, np.float64)[np.asarray(tris)].reshape(-1, 9))
    n = len(S)
    dist, p1, p2 = np.empty(n), np.empty((n, 3)), np.empty((n, 3))
    b1, n_bv, n_leaf = np.empty(n, np.int32), np.empty(n, np.uint32), np.empty(n, np.uint32)
    M, S = np.ascontiguousarray(M), np.ascontiguousarray(S)
    ov = hm.lib().hm_mesh_sphere_distance(n, hm.dptr(M), hm.dptr(S), float(radius), hm.iptr(a["first_child"]), hm.dptr(a["axis"]),
                                          hm.dptr(a["obb_To"]), hm.dptr(a["obb_ext"]), hm.dptr(tri9), hm.dptr(dist), hm.dptr(p1),
                                          hm.dptr(p2), hm.iptr(b1), n_bv.ctypes.data_as(C.POINTER(C.c_uint32)),
                                          n_leaf.ctypes.data_as(C.POINTER(C.c_uint32)), int(bound32))
    assert ov == 0
    return dict(min_distance=dist, p1=p1, p2=p2, b1=b1, n_bv=n_bv, n_leaf=n_leaf)


def test_mesh_sphere_distance_traversal_matches_oracle(hm, oracle, env_rob_npz):
    """The traversal the kernel inlines (csrc/mesh_sphere.cuh), run on the host over the oracle's node arrays:
    minimum distance bit-exact against the oracle's brute-force pass over all triangles (the point-to-box bound
    never excludes the closest triangle), equal to the reference-order traversal up to ties between adjacent
    triangles, nearest points bit-exact whenever the same triangle is reported."""
    from fcl_b200.poses import identity_poses
    from tests.meshes import heightfield

    (ev, et), _ = env_rob_npz
    hv, ht = heightfield(40, size=10.0, seed=5, amp=0.5)
    rng = np.random.default_rng(5)
    cases = []
    n = 6000
    S = random_poses(n, seed=31)
    M = identity_poses(n)
    M[: n // 2] = random_poses(n // 2, seed=37)
    S[: n // 2, 9:] = np.einsum("nij,nj->ni", M[: n // 2, :9].reshape(-1, 3, 3), S[: n // 2, 9:]) + M[: n // 2, 9:]
    cases.append((ev, et, (0.0, 10.0, 150.0, 600.0), M, S))
    Sh = identity_poses(3000)
    Sh[:, 9:11] = rng.uniform(-6, 6, size=(3000, 2))
    Sh[:, 11] = rng.uniform(-1.0, 3.0, size=3000)
    cases.append((hv, ht, (0.05, 0.3), identity_poses(3000), Sh))
    for v, t, radii, M, S in cases:
        o = oracle.Model(v, t)
        for radius, bound32 in [(r, b) for r in radii for b in (False, True)]:
            got = _host_mesh_sphere(hm, o, v, t, radius, M, S, bound32)
            brute = oracle.distance_mesh_sphere_batch(o, radius, M, S, brute=True, nthreads=8)
            trav = oracle.distance_mesh_sphere_batch(o, radius, M, S, nthreads=8)
            assert np.array_equal(got["min_distance"], brute["min_distance"]), radius
            pos = brute["min_distance"] > 0
            assert np.array_equal(brute["min_distance"] < 0, trav["min_distance"] < 0)
            rel = np.abs(brute["min_distance"][pos] - trav["min_distance"][pos]) / trav["min_distance"][pos]
            assert rel.max() <= 1e-12
            same = pos & (got["b1"] == brute["b1"])
            assert same.sum() > 0  # exact ties at shared vertices / edges may name a neighbouring triangle: those go through allclose below
            assert got["p1"][same].tobytes() == brute["p1"][same].tobytes()
            assert got["p2"][same].tobytes() == brute["p2"][same].tobytes()
            assert np.allclose(got["p1"][pos], brute["p1"][pos], rtol=1e-6, atol=1e-6 * (1 + np.abs(brute["p1"][pos]).max()))
            assert np.isnan(got["p1"][~pos]).all() and np.isnan(got["p2"][~pos]).all()
            # tighter bound than the reference's RSS-to-RSS test: no more leaf tests than its traversal
            assert got["n_leaf"].mean() <= trav["n_leaf"].mean() + 1
            assert 0 < pos.sum() and (radius < 100 or (~pos).sum() > 0)


def test_f32_rss_lower_bound_is_conservative(hm, oracle, oracle_env_rob):
    """The single-precision steering bound must never exceed the exact RSS distance (it may be
    smaller): checked on real node pairs at the benchmark's pose distribution and on close /
    near-parallel / huge-offset configurations."""
    import ctypes as C

    env, rob = oracle_env_rob
    a1, a2 = env.arrays(), rob.arrays()
    rng = np.random.default_rng(23)
    n = 300000
    rel = _rel_pose(random_poses(n, seed=31))
    rel[: n // 3, 9:] *= rng.uniform(0, 0.3, size=(n // 3, 1))           # close configurations
    q = n // 3
    ang = np.round(rng.uniform(0, 2 * np.pi, size=(q, 3)) / (np.pi / 2)) * (np.pi / 2) + rng.normal(0, 1e-6, size=(q, 3))
    from fcl_b200.poses import euler_to_matrix
    rel[q:2 * q, :9] = euler_to_matrix(ang[:, 0], ang[:, 1], ang[:, 2]).reshape(q, 9)  # near axis-aligned
    rel[-2000:, 9:] *= 50.0                                              # far away (large magnitudes)
    i1 = rng.integers(0, env.num_bvs, n).astype(np.int32)
    i2 = rng.integers(0, rob.num_bvs, n).astype(np.int32)
    exact = np.empty(n)
    L = hm.lib()
    L.hm_rss_pairs(n, hm.dptr(rel), hm.iptr(i1), hm.iptr(i2), hm.dptr(a1["axis"]), hm.dptr(a1["rss_To"]),
                   hm.dptr(a1["rss_l"]), hm.dptr(a1["rss_r"]), hm.dptr(a2["axis"]), hm.dptr(a2["rss_To"]),
                   hm.dptr(a2["rss_l"]), hm.dptr(a2["rss_r"]), hm.dptr(exact))
    lb = np.empty(n, np.float32)
    L.hm_rss_lb32_pairs(n, hm.dptr(rel), hm.iptr(i1), hm.iptr(i2), hm.dptr(a1["axis"]), hm.dptr(a1["rss_To"]),
                        hm.dptr(a1["rss_l"]), hm.dptr(a1["rss_r"]), hm.dptr(a2["axis"]), hm.dptr(a2["rss_To"]),
                        hm.dptr(a2["rss_l"]), hm.dptr(a2["rss_r"]), lb.ctypes.data_as(C.POINTER(C.c_float)))
    lb = lb.astype(np.float64)
    assert (lb <= exact).all(), float((lb - exact).max())
    # and it is a useful bound: on separated pairs it recovers most of the exact distance
    sep = exact > 50
    assert sep.sum() > 10000
    assert np.median(lb[sep] / exact[sep]) > 0.9


def test_f32_obb_sat_is_conservative(hm, oracle_env_rob):
    """"certainly disjoint" in single precision must imply disjoint for the exact (FP64) SAT, and
    it should recognise almost every disjoint pair (otherwise the traversal would do extra work)."""
    env, rob = oracle_env_rob
    a1, a2 = env.arrays(), rob.arrays()
    rng = np.random.default_rng(29)
    n = 400000
    rel = _rel_pose(random_poses(n, seed=37))
    rel[: n // 2, 9:] *= rng.uniform(0, 0.4, size=(n // 2, 1))
    i1 = rng.integers(0, env.num_bvs, n).astype(np.int32)
    i2 = rng.integers(0, rob.num_bvs, n).astype(np.int32)
    # a third of the pairs: put box 2 next to box 1 (touching / barely separated / overlapping)
    m = n // 3
    R0 = rel[:m, :9].reshape(m, 3, 3)
    c1, c2 = a1["obb_To"][i1[:m]], a2["obb_To"][i2[:m]]
    reach = np.linalg.norm(a1["obb_ext"][i1[:m]], axis=1) + np.linalg.norm(a2["obb_ext"][i2[:m]], axis=1)
    dirs = rng.normal(size=(m, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    rel[:m, 9:] = c1 - np.einsum("nij,nj->ni", R0, c2) + dirs * (reach * rng.uniform(0.0, 1.1, size=m))[:, None]
    L = hm.lib()
    exact = np.empty(n, np.int32)
    L.hm_obb_pairs(n, hm.dptr(rel), hm.iptr(i1), hm.iptr(i2), hm.dptr(a1["axis"]), hm.dptr(a1["obb_To"]),
                   hm.dptr(a1["obb_ext"]), hm.dptr(a2["axis"]), hm.dptr(a2["obb_To"]), hm.dptr(a2["obb_ext"]), hm.iptr(exact))
    got = np.empty(n, np.int32)
    L.hm_obb_disjoint32_pairs(n, hm.dptr(rel), hm.iptr(i1), hm.iptr(i2), hm.dptr(a1["axis"]), hm.dptr(a1["obb_To"]),
                              hm.dptr(a1["obb_ext"]), hm.dptr(a2["axis"]), hm.dptr(a2["obb_To"]), hm.dptr(a2["obb_ext"]), hm.iptr(got))
    assert not np.any((got == 1) & (exact == 0))          # never "disjoint" when the exact test says overlap
    assert exact.sum() > n // 10 and (exact == 0).sum() > n // 20
    missed = ((got == 0) & (exact == 1)).sum()
    assert missed < 2e-3 * exact.sum(), missed            # sharp: < 0.2 % of disjoint pairs left undecided


def test_f32_triangle_lower_bound_is_conservative(hm, oracle, env_rob_npz):
    """lb32(tri, tri) <= exact triDistance, on synthetic pairs at several scales and on real
    env/rob triangle pairs under benchmark poses; and it is tight enough to be useful."""
    L = hm.lib()
    rng = np.random.default_rng(41)
    ratios = []
    for scale, offset in ((1.0, 0.0), (1000.0, 3000.0), (0.01, 100.0)):
        n = 20000
        S, T = _tri_pairs(rng, n, 2.5)
        S, T = S * scale + offset, T * scale + offset
        for k in range(n):
            lb = float(L.hm_tri_lb32(hm.dptr(S[k]), hm.dptr(T[k])))
            d, _, _ = oracle.tri_distance(S[k], T[k])
            assert lb <= d, (scale, k, lb, d)
            if d > 0.5 * scale:
                ratios.append(lb / d)
    (ev, et), (rv, rt) = env_rob_npz
    P = random_poses(4000, seed=43)
    ti = rng.integers(0, len(et), len(P))
    tj = rng.integers(0, len(rt), len(P))
    for k in range(len(P)):
        R1, t1 = P[k, :9].reshape(3, 3), P[k, 9:]
        S = ev[et[ti[k]]].reshape(9)
        Tw = (R1.T @ (rv[rt[tj[k]]] - t1).T).T.reshape(9)  # rob triangle in env's frame
        lb = float(L.hm_tri_lb32(hm.dptr(S), hm.dptr(np.ascontiguousarray(Tw))))
        d, _, _ = oracle.tri_distance(S, Tw)
        assert lb <= d, (k, lb, d)
        if d > 10:
            ratios.append(lb / d)
    assert np.median(ratios) > 0.95


def test_f32_triangle_screening_subset_is_a_valid_lower_bound(hm, oracle, env_rob_npz):
    """The direction subset the distance kernel's screening round uses (FCLGPU_SCREEN_LEAVES = 9: face normals +
    in-plane edge normals) never exceeds the exact triangle distance -- also for touching / overlapping pairs and
    at large coordinate offsets -- and the full set reproduces tri_lower_bound_f32."""
    import ctypes as C

    L = hm.lib()
    L.hm_tri_lb32_dirs.restype = C.c_float
    L.hm_tri_lb32_dirs.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    rng = np.random.default_rng(53)
    ratios = []
    for scale, offset, spread in ((1.0, 0.0, 2.5), (1000.0, 3000.0, 2.5), (0.01, 100.0, 2.5), (1.0, 0.0, 0.3)):
        n = 12000
        S, T = _tri_pairs(rng, n, spread)
        S, T = S * scale + offset, T * scale + offset
        T[:500, :3] = S[:500, :3]  # shared vertex: distance 0
        for k in range(n):
            d, _, _ = oracle.tri_distance(S[k], T[k])
            for dirs in (9, 3, 11):
                lb = float(L.hm_tri_lb32_dirs(hm.dptr(S[k]), hm.dptr(T[k]), dirs))
                assert lb <= d, (scale, k, dirs, lb, d)
            if k < 2000:
                assert float(L.hm_tri_lb32_dirs(hm.dptr(S[k]), hm.dptr(T[k]), 15)) == float(L.hm_tri_lb32(hm.dptr(S[k]), hm.dptr(T[k])))
            if d > 0.5 * scale:
                ratios.append(float(L.hm_tri_lb32_dirs(hm.dptr(S[k]), hm.dptr(T[k]), 9)) / d)
    (ev, et), (rv, rt) = env_rob_npz
    P = random_poses(3000, seed=59)
    ti = rng.integers(0, len(et), len(P))
    tj = rng.integers(0, len(rt), len(P))
    for k in range(len(P)):
        R1, t1 = P[k, :9].reshape(3, 3), P[k, 9:]
        S = ev[et[ti[k]]].reshape(9)
        Tw = np.ascontiguousarray((R1.T @ (rv[rt[tj[k]]] - t1).T).T.reshape(9))
        d, _, _ = oracle.tri_distance(S, Tw)
        assert float(L.hm_tri_lb32_dirs(hm.dptr(S), hm.dptr(Tw), 9)) <= d, k
    assert np.median(ratios) > 0.5


def test_f32_triangle_classification_is_sound(hm, oracle, env_rob_npz):
    """+1 (certainly separated) => the exact test says no intersection; -1 (certainly intersecting)
    => the exact test says intersection; 0 (undecided) must be rare on generic input."""
    L = hm.lib()
    rng = np.random.default_rng(47)
    stats = {1: 0, -1: 0, 0: 0}
    for scale, offset, spread in ((1.0, 0.0, 0.8), (500.0, 4000.0, 0.8), (0.01, 50.0, 1.5)):
        n = 15000
        P, Q = _tri_pairs(rng, n, spread)
        P, Q = P * scale + offset, Q * scale + offset
        Q[:1000, :3] = P[:1000, :3]            # shared vertex (touching) -> must never be classified wrongly
        Q[1000:1500] = P[1000:1500]            # identical triangles (coplanar, parallel edges)
        pose = np.concatenate([np.eye(3).reshape(9), np.zeros(3)])
        for k in range(n):
            c = L.hm_tri_classify32(hm.dptr(P[k]), hm.dptr(Q[k]), hm.dptr(pose))
            hit = oracle.tri_intersect(P[k], Q[k])
            assert not (c == 1 and hit), (scale, k)
            assert not (c == -1 and not hit), (scale, k)
            if k >= 1500:  # generic pairs only (the injected touching / coplanar ones are legitimately undecided)
                stats[c] += 1
    # real mesh triangles under benchmark poses (model coordinates ~1e3, poses ~1e3)
    (ev, et), (rv, rt) = env_rob_npz
    poses = random_poses(6000, seed=53)
    rel = _rel_pose(poses)
    ti, tj = rng.integers(0, len(et), len(poses)), rng.integers(0, len(rt), len(poses))
    for k in range(len(poses)):
        Pk = np.ascontiguousarray(ev[et[ti[k]]].reshape(9))
        Qk = np.ascontiguousarray(rv[rt[tj[k]]].reshape(9))
        c = L.hm_tri_classify32(hm.dptr(Pk), hm.dptr(Qk), hm.dptr(rel[k]))
        hit = oracle.tri_intersect(Pk, Qk, rel[k, :9], rel[k, 9:])
        assert not (c == 1 and hit) and not (c == -1 and not hit), k
    assert stats[1] > 10000 and stats[-1] > 3000
    assert stats[0] < 0.002 * sum(stats.values()), stats   # undecided only for touching / degenerate pairs
