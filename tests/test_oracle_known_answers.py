"""Pins the CPU oracle against the reference's own known answers and invariants
(SURVEY.md section 8c).  CPU only."""
import numpy as np
import pytest

from fcl_b200.poses import identity_poses, random_poses
from tests.meshes import box_mesh, uv_sphere

I3 = np.eye(3)
Z3 = np.zeros(3)


def test_rss_distance_known_answers(oracle):
    """test/test_fcl_math.cpp:256-291 -- 2.0, 1.0, sqrt(6)-1, 1.0 (eps_78)."""
    eps = 1e-7
    # case 1: two triangles in the planes y=+1 / y=-1, fitted by the mesh RSS fitter
    m0 = oracle.Model([[1, 1, 1], [1, 1, -1], [0, 1, -1]], [[0, 1, 2]])
    m1 = oracle.Model([[1, -1, 1], [1, -1, -1], [0, -1, -1]], [[0, 1, 2]])
    a0, a1 = m0.arrays(), m1.arrays()
    d = oracle.rss_distance(I3, Z3, a0["axis"][0], a0["rss_To"][0], a0["rss_l"][0], a0["rss_r"][0],
                            a1["axis"][0], a1["rss_To"][0], a1["rss_l"][0], a1["rss_r"][0])
    assert abs(d - 2.0) < eps
    # cases 2-4: hand-built RSS pairs
    d = oracle.rss_distance(I3, Z3, I3, [0, 0, 0.5], [1, 1], 0.5, I3, [-1, -1, 2.5], [1, 1], 0.5)
    assert abs(d - 1.0) < eps
    d = oracle.rss_distance(I3, Z3, I3, [0, 0, 0.5], [1, 1], 0.5, -I3, [-1, -1, 2.5], [1, 1], 0.5)
    assert abs(d - (np.sqrt(6) - 1.0)) < eps
    d = oracle.rss_distance(I3, Z3, I3, [0, 0, 0.5], [1, 1], 0.5, -I3, [0, 0, 2.5], [1, 1], 0.5)
    assert abs(d - 1.0) < eps


def test_obb_overlap_equals_aabb_for_axis_aligned(oracle):
    """test/test_fcl_collision.cpp:271-311 -- OBB::overlap == AABB overlap for axis-aligned boxes."""
    rng = np.random.default_rng(7)
    n_over = 0
    for _ in range(2000):
        c1, c2 = rng.uniform(-100, 100, 3), rng.uniform(-100, 100, 3)
        e1, e2 = rng.uniform(1, 60, 3), rng.uniform(1, 60, 3)
        aabb = bool(np.all(np.abs(c1 - c2) <= e1 + e2))
        gap = np.abs(np.abs(c1 - c2) - (e1 + e2)).min()
        if gap < 1e-3:  # the SAT pads |R| by 1e-6; skip razor-edge samples
            continue
        got = oracle.obb_overlap(I3, Z3, I3, c1, e1, I3, c2, e2)
        assert got == aabb
        n_over += aabb
    assert n_over > 50


def test_bvh_tree_invariants(oracle_env_rob):
    """2n-1 nodes, children adjacent, every triangle in exactly one leaf."""
    for m in oracle_env_rob:
        a = m.arrays()
        fc = a["first_child"]
        assert len(fc) == 2 * m.num_tris - 1
        leaves = fc[fc < 0]
        assert sorted((-(leaves + 1)).tolist()) == list(range(m.num_tris))
        internal = fc[fc >= 0]
        assert len(set(internal.tolist())) == len(internal)
        assert internal.max() + 1 == len(fc) - 1
        # rotation matrices
        ax = a["axis"].reshape(-1, 3, 3)
        assert np.allclose(np.einsum("nij,nik->njk", ax, ax), np.eye(3), atol=1e-9)
        assert (a["obb_ext"] >= 0).all() and (a["rss_l"] >= 0).all() and (a["rss_r"] >= 0).all()


def test_collide_pairs_equal_bruteforce_and_split_methods(oracle, env_rob_npz):
    """test/test_fcl_collision.cpp:792-886 invariant: sorted contact pair sets identical across
    split methods and equal to exhaustive triangle-pair testing."""
    (ev, et), (rv, rt) = env_rob_npz
    P = random_poses(24, seed=3)
    sets = []
    for split in (oracle.SPLIT_MEAN, oracle.SPLIT_MEDIAN, oracle.SPLIT_BV_CENTER):
        env, rob = oracle.Model(ev, et, split), oracle.Model(rv, rt, split)
        r = oracle.collide_batch(env, rob, P, None, 2**31 - 1, True, nthreads=4)
        per = []
        for i in range(len(P)):
            c = r["contacts"][r["offsets"][i]:r["offsets"][i + 1]]
            per.append(sorted(set(zip(c["b1"].tolist(), c["b2"].tolist()))))
        sets.append(per)
    assert sets[0] == sets[1] == sets[2]
    env, rob = oracle.Model(ev, et), oracle.Model(rv, rt)
    n_col = 0
    for i in range(len(P)):
        brute = sorted(map(tuple, oracle.brute_collide(env, rob, P[i]).tolist()))
        assert brute == sets[0][i]
        n_col += bool(brute)
    assert n_col >= 3


def test_distance_equal_bruteforce_queue_and_split_methods(oracle, env_rob_npz):
    """test/test_fcl_distance.cpp:177-298 invariant (tolerance there 1e-3; here exact for the
    distance because every variant evaluates the same triangle pair)."""
    (ev, et), (rv, rt) = env_rob_npz
    P = random_poses(24, seed=4)
    res = []
    for split in (oracle.SPLIT_MEAN, oracle.SPLIT_MEDIAN, oracle.SPLIT_BV_CENTER):
        env, rob = oracle.Model(ev, et, split), oracle.Model(rv, rt, split)
        for q in (2, 20):
            res.append(oracle.distance_batch(env, rob, P, None, True, q, nthreads=4))
    ref = res[0]
    for r in res[1:]:
        assert np.array_equal(r["min_distance"], ref["min_distance"])
        pos = ref["min_distance"] > 0
        assert np.allclose(r["p1"][pos], ref["p1"][pos], rtol=0, atol=1e-6)
        assert np.allclose(r["p2"][pos], ref["p2"][pos], rtol=0, atol=1e-6)
    env, rob = oracle.Model(ev, et), oracle.Model(rv, rt)
    for i in range(8):
        d, p1, p2, _ = oracle.brute_distance(env, rob, P[i])
        assert d == ref["min_distance"][i]
        if d > 0:
            assert abs(np.linalg.norm(ref["p1"][i] - ref["p2"][i]) - d) < 1e-9 * max(1.0, d)


def test_tessellated_spheres_analytic(oracle):
    """test/test_fcl_shape_mesh_consistency.cpp:57-80: two r=20 spheres (16x16), centres 50 apart:
    mesh distance within 5% of 10; moved to 22.6 apart they still do not touch... and overlap at 20."""
    v1, t1 = uv_sphere(20, 16, 16)
    m1 = oracle.Model(v1, t1)
    m2 = oracle.Model(v1, t1)
    tf2 = identity_poses(1)
    tf2[0, 9] = 50.0
    r = oracle.distance_batch(m1, m2, identity_poses(1), tf2, True)
    assert abs(r["min_distance"][0] - 10.0) < 0.05 * 10.0 + 1.0
    assert abs(np.linalg.norm(r["p1"][0] - r["p2"][0]) - r["min_distance"][0]) < 1e-9
    c = oracle.collide_batch(m1, m2, identity_poses(1), tf2, 1, False)
    assert c["counts"][0] == 0
    tf2[0, 9] = 30.0
    r = oracle.distance_batch(m1, m2, identity_poses(1), tf2, True)
    assert r["min_distance"][0] == 0.0
    c = oracle.collide_batch(m1, m2, identity_poses(1), tf2, 1, False)
    assert c["counts"][0] == 1


def test_collide_request_semantics(oracle, oracle_env_rob):
    """num_max_contacts truncation (DFS prefix), enable_contact budget clamp, num_max_contacts == 0."""
    env, rob = oracle_env_rob
    P = random_poses(64, seed=5)
    full = oracle.collide_batch(env, rob, P, None, 2**31 - 1, True)
    assert oracle.collide_batch(env, rob, P, None, 0, True)["counts"].sum() == 0
    for k in (1, 3, 100):
        part = oracle.collide_batch(env, rob, P, None, k, True)
        assert (part["counts"] == np.minimum(full["counts"], k)).all()
        for i in range(len(P)):
            a = part["contacts"][part["offsets"][i]:part["offsets"][i + 1]]
            b = full["contacts"][full["offsets"][i]:full["offsets"][i] + len(a)]
            assert a.tobytes() == b.tobytes()
    # binary mode reports the same pair sequence (one entry per intersecting pair)
    binm = oracle.collide_batch(env, rob, P, None, 2**31 - 1, False)
    for i in range(len(P)):
        a = binm["contacts"][binm["offsets"][i]:binm["offsets"][i + 1]]
        b = full["contacts"][full["offsets"][i]:full["offsets"][i + 1]]
        pa = list(zip(a["b1"].tolist(), a["b2"].tolist()))
        pb = list(dict.fromkeys(zip(b["b1"].tolist(), b["b2"].tolist())))
        assert pa == pb


def test_contact_geometry_box_box(oracle):
    """Two overlapping boxes: contact normals unit length, positions inside both boxes' hull, depth >= 0."""
    v, t = box_mesh(1, 1, 1)
    m1, m2 = oracle.Model(v, t), oracle.Model(v, t)
    tf2 = identity_poses(1)
    tf2[0, 9:] = [1.5, 0.2, 0.1]
    r = oracle.collide_batch(m1, m2, identity_poses(1), tf2, 1000, True)
    assert r["counts"][0] > 0
    c = r["contacts"]
    assert np.allclose(np.linalg.norm(c["normal"], axis=1), 1.0, atol=1e-12)
    assert (np.abs(c["pos"]) <= 2.5 + 1e-9).all()


# ---------------------------------------------------------------------------------------------
# mesh <-> sphere (closed-form sphereTriangleIntersect; no GJK on this pair)
# ---------------------------------------------------------------------------------------------
def test_sphere_triangle_known_answers(oracle):
    """test/test_fcl_geometric_shapes.cpp:1248-1287 (shapeIntersection_spheretriangle): sphere r = 10 at the
    origin hits both triangles; for the second one the normal is (1, 0, 0); the same after a rigid motion of
    both (normal rotated accordingly)."""
    from fcl_b200.poses import random_poses

    t_a = np.array([[20.0, 0, 0], [-20, 0, 0], [0, 20, 0]])
    t_b = np.array([[30.0, 0, 0], [9.9, -20, 0], [9.9, 20, 0]])
    hit, _, _, _ = oracle.sphere_tri_intersect([0, 0, 0], 10.0, t_a)
    assert hit
    hit, cp, depth, n = oracle.sphere_tri_intersect([0, 0, 0], 10.0, t_b)
    assert hit and np.allclose(n, [1, 0, 0], atol=1e-9) and np.allclose(cp, [9.9, 0, 0]) and abs(depth + 0.1) < 1e-12
    for pose in random_poses(20, seed=71):
        R, t = pose[:9].reshape(3, 3), pose[9:]
        hit, _, _, _ = oracle.sphere_tri_intersect(t, 10.0, t_a @ R.T + t)
        assert hit
        hit, _, _, n = oracle.sphere_tri_intersect(t, 10.0, t_b @ R.T + t)
        assert hit and np.allclose(n, R @ np.array([1.0, 0, 0]), atol=1e-9)
    # just outside (gap beyond the epsilon threshold) / touching an edge from outside the face
    assert not oracle.sphere_tri_intersect([0, 0, 10.0 + 1e-9], 10.0, t_a)[0]
    assert oracle.sphere_tri_intersect([0, -5.0, 0], 10.0, t_a)[0]       # projects outside, edge within reach
    assert not oracle.sphere_tri_intersect([0, -10.5, 0], 10.0, t_a)[0]


def test_mesh_sphere_traversal_equals_brute_force(oracle, oracle_env_rob):
    """The OBBRSS traversal with the sphere's fitted OBB reports exactly the triangles the brute-force loop finds
    (in DFS order), budgets truncate that list, and every contact has b2 = -1 (Contact::NONE)."""
    from fcl_b200.poses import identity_poses, random_poses

    env, _ = oracle_env_rob
    n = 300
    S = random_poses(n, seed=73)       # sphere poses inside env's extents
    M = identity_poses(n)
    M[: n // 2] = random_poses(n // 2, seed=79)  # half of the queries also move the mesh
    S[: n // 2, 9:] = np.einsum("nij,nj->ni", M[: n // 2, :9].reshape(-1, 3, 3), S[: n // 2, 9:]) + M[: n // 2, 9:]
    radius = 400.0
    full = oracle.collide_mesh_sphere_batch(env, radius, M, S, 1 << 30, True, nthreads=4)
    few = oracle.collide_mesh_sphere_batch(env, radius, M, S, 3, False, nthreads=4)
    hits = 0
    for i in range(n):
        ids = full["contacts"]["b1"][full["offsets"][i]:full["offsets"][i + 1]]
        brute = oracle.brute_mesh_sphere(env, radius, M[i], S[i])
        assert sorted(ids.tolist()) == brute.tolist(), i
        assert (full["contacts"]["b2"][full["offsets"][i]:full["offsets"][i + 1]] == -1).all()
        k = few["contacts"]["b1"][few["offsets"][i]:few["offsets"][i + 1]]
        assert k.tolist() == ids[:3].tolist()
        hits += len(brute) > 0
    assert 0.15 * n < hits < 0.95 * n
    c = full["contacts"]
    assert (c["depth"] <= 0).all() and (c["depth"] >= -radius).all()
    assert np.allclose(np.linalg.norm(c["normal"], axis=1), 1.0, atol=1e-12)


def test_mesh_sphere_distance_reference_known_answers(oracle):
    """test/test_fcl_shape_mesh_consistency.cpp:57-140, the `distance(&s1_rss, pose1, &s2, pose2)` legs: an r=20
    tessellated sphere (16x16) against an r=20 Sphere, centres 50 apart -> within 5 % of the analytic 10; 40.1 apart
    -> within 200 % of 0.1; both also under ten common random rigid motions (extents 0..10).  Then the oracle's own
    invariants: traversal == brute force over all triangles, |tf1 p1 - tf2 p2| == distance, p2 on the sphere."""
    v, t = uv_sphere(20, 16, 16)
    m = oracle.Model(v, t)
    T = random_poses(10, seed=7, extents=(0, 0, 0, 10, 10, 10))
    for gap, tol in ((50.0, 0.05), (40.1, 2.0)):
        true = gap - 40.0
        tf1 = np.concatenate([identity_poses(1), T])
        tf2 = tf1.copy()
        R = tf1[:, :9].reshape(-1, 3, 3)
        tf2[:, 9:] = np.einsum("nij,j->ni", R, np.array([gap, 0.0, 0.0])) + tf1[:, 9:]
        r = oracle.distance_mesh_sphere_batch(m, 20.0, tf1, tf2)
        assert np.all(np.abs(r["min_distance"] - true) / true < tol)
        assert np.allclose(r["min_distance"], r["min_distance"][0], rtol=1e-9)
        b = oracle.distance_mesh_sphere_batch(m, 20.0, tf1, tf2, brute=True)
        assert np.array_equal(r["min_distance"], b["min_distance"])
        w1 = np.einsum("nij,nj->ni", R, r["p1"]) + tf1[:, 9:]
        w2 = np.einsum("nij,nj->ni", tf2[:, :9].reshape(-1, 3, 3), r["p2"]) + tf2[:, 9:]
        assert np.allclose(np.linalg.norm(w1 - w2, axis=1), r["min_distance"], atol=1e-9)
        assert np.allclose(np.linalg.norm(r["p2"], axis=1), 20.0, atol=1e-9)
    # centre within the radius of a triangle: the reference leaves the result unwritten; defined here as -1 / NaN
    tf2 = identity_poses(1)
    tf2[0, 9] = 30.0
    r = oracle.distance_mesh_sphere_batch(m, 20.0, identity_poses(1), tf2)
    assert r["min_distance"][0] == -1.0 and np.isnan(r["p1"]).all()


def test_sphere_fitted_obbrss_encloses_its_bound_vertices(oracle):
    """computeBV<OBBRSS>(Sphere, tf) = fit over the 12 bound vertices (geometry/shape/sphere-inl.h:95-120,
    math/bv/utility-inl.h:516-521): every bound vertex lies inside the OBB and within r of the RSS rectangle."""
    m = (1 + np.sqrt(5.0)) / 2.0
    for radius, seed in ((1.0, 1), (20.0, 2), (350.0, 3)):
        tf = random_poses(1, seed=seed)[0]
        bv = oracle.sphere_bv(radius, tf)
        edge = radius * 6 / (np.sqrt(27.0) + np.sqrt(15.0))
        a, b = edge, m * edge
        L = np.array([[0, a, b], [0, -a, b], [0, a, -b], [0, -a, -b], [a, b, 0], [-a, b, 0], [a, -b, 0], [-a, -b, 0],
                      [b, 0, a], [b, 0, -a], [-b, 0, a], [-b, 0, -a]])
        P = L @ tf[:9].reshape(3, 3).T + tf[9:]
        loc = (P - bv["obb_To"]) @ bv["axis"]
        assert np.all(np.abs(loc) <= bv["obb_ext"] * (1 + 1e-12) + 1e-9)
        q = (P - bv["rss_To"]) @ bv["axis"]
        dx = np.maximum(np.maximum(-q[:, 0], q[:, 0] - bv["rss_l"][0]), 0)
        dy = np.maximum(np.maximum(-q[:, 1], q[:, 1] - bv["rss_l"][1]), 0)
        assert np.all(np.sqrt(dx * dx + dy * dy + q[:, 2] ** 2) <= bv["rss_r"] * (1 + 1e-9) + 1e-9)
        assert np.all(np.linalg.norm(P - tf[9:], axis=1) >= radius * (1 - 1e-12))  # the icosahedron encloses the sphere
