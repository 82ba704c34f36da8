"""SURVEY 8f rank 2, distance half: fcl::distance(BVHModel<OBBRSS>, tf1, Sphere, tf2) on the GPU, through the C ABI
(fclgpu_distance_mesh_sphere_batch_host), against the oracle's restatement of MeshShapeDistanceTraversalNodeOBBRSS +
sphereTriangleDistance.

Parity statement.  The minimum distance is the minimum of (distance(centre, triangle) - radius) over every triangle no
valid bound excludes; the kernel bounds nodes more tightly than the reference (point-to-OBB instead of RSS-to-RSS), so
it returns the minimum over ALL triangles: bit-exact against the oracle's brute-force pass, and equal to the
oracle's traversal except where two triangles sharing the closest edge differ by an ULP and the reference's bound
prunes the smaller one (tolerance 1e-12 relative, far inside north_star's 1e-6).  Nearest points: bit-exact whenever
the same triangle is reported, else 1e-6 (ties at shared vertices / edges)."""
import numpy as np
import pytest

import fcl_b200 as F
from fcl_b200.poses import identity_poses, random_poses

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[(16, 1, 3), (16, 0, 3), (0, 0, 3), (32, 1, 4), (1, 1, 5)],
                ids=lambda v: "leaf_trigger=%d,bound32=%d,blocks=%d" % v, autouse=True)
def kernel_variant(request):
    """Kernel variants: leaf rounds triggered by 16 (default) / 32 / 1 parked lanes, 0 = leaf tests inline; box bound
    from the packed FP32 records (default) or from the FP64 records; register budget for 3 / 4 / 5 blocks per SM."""
    from fcl_b200 import _capi

    saved = {k: _capi.get_option(k) for k in ("sphere_leaf_trigger", "sphere_bound32", "sphere_blocks")}
    _capi.set_option("sphere_leaf_trigger", request.param[0])
    _capi.set_option("sphere_bound32", request.param[1])
    _capi.set_option("sphere_blocks", request.param[2])
    yield request.param
    for k, v in saved.items():
        _capi.set_option(k, v)


def _check(got, brute, trav, radius):
    assert np.array_equal(got.min_distance, brute["min_distance"])  # bit-exact, -1 cases included
    pos = brute["min_distance"] > 0
    rel = np.abs(got.min_distance[pos] - trav["min_distance"][pos]) / trav["min_distance"][pos]
    assert np.array_equal(got.min_distance < 0, trav["min_distance"] < 0) and (rel.size == 0 or rel.max() <= 1e-12)
    assert (got.b2 == -1).all() and (got.b1 >= 0).all()
    same = pos & (got.b1 == brute["b1"])
    assert same.sum() > 0  # exact ties at shared vertices / edges may name a neighbouring triangle: those go through allclose below
    assert got.nearest_p1[same].tobytes() == brute["p1"][same].tobytes()
    assert got.nearest_p2[same].tobytes() == brute["p2"][same].tobytes()
    scale = 1.0 + np.abs(brute["p1"][pos]).max() if pos.any() else 1.0
    assert np.allclose(got.nearest_p1[pos], brute["p1"][pos], rtol=1e-6, atol=1e-6 * scale)
    assert np.allclose(got.nearest_p2[pos], brute["p2"][pos], rtol=1e-6, atol=1e-6 * max(radius, 1.0))
    neg = ~pos
    assert np.isnan(got.nearest_p1[neg]).all() and np.isnan(got.nearest_p2[neg]).all()


def test_mesh_sphere_distance_matches_oracle(oracle, env_rob_npz):
    (ev, et), _ = env_rob_npz
    env, oenv = F.BVHModel.from_arrays(ev, et), oracle.Model(ev, et)
    n = 20000
    S = random_poses(n, seed=101)
    M = identity_poses(n)
    M[: n // 2] = random_poses(n // 2, seed=103)
    S[: n // 2, 9:] = np.einsum("nij,nj->ni", M[: n // 2, :9].reshape(-1, 3, 3), S[: n // 2, 9:]) + M[: n // 2, 9:]
    seen_neg = seen_pos = 0
    for radius in (0.0, 10.0, 150.0, 600.0):
        got = F.distance_mesh_sphere_batch(env, M, F.Sphere(radius), S, F.DistanceRequest(True), stats=True)
        brute = oracle.distance_mesh_sphere_batch(oenv, radius, M, S, brute=True, nthreads=8)
        trav = oracle.distance_mesh_sphere_batch(oenv, radius, M, S, nthreads=8)
        _check(got, brute, trav, radius)
        # every query walks the tree, except that a triangle within the radius ends it on the spot (already the seed
        # triangle, before any box test, when the centre is that close to triangle 0)
        inside = got.min_distance == -1.0
        assert (got.n_bv[~inside] > 0).all() and (got.n_leaf[~inside] > 0).all()
        seen_neg += int((got.min_distance < 0).sum())
        seen_pos += int((got.min_distance > 0).sum())
    assert seen_neg > 1000 and seen_pos > 10000
    # without nearest points: only distances and ids are written
    g2 = F.distance_mesh_sphere_batch(env, M, F.Sphere(150.0), S, F.DistanceRequest(False))
    b2 = oracle.distance_mesh_sphere_batch(oenv, 150.0, M, S, brute=True, nthreads=8)
    assert np.array_equal(g2.min_distance, b2["min_distance"])


def test_mesh_sphere_distance_reference_known_answers(oracle):
    """test/test_fcl_shape_mesh_consistency.cpp:57-140 (distance(&s1_rss, I, &s2, pose)): r=20 tessellated sphere vs
    r=20 Sphere, centres 50 apart -> within 5 % of 10; 40.1 apart -> within 200 % of 0.1; the same under a common
    random rigid motion."""
    from tests.meshes import uv_sphere

    v, t = uv_sphere(20, 16, 16)
    mesh = F.BVHModel.from_arrays(v, t)
    T = random_poses(10, seed=7, extents=(0, 0, 0, 10, 10, 10))
    for gap, tol in ((50.0, 0.05), (40.1, 2.0)):
        true = gap - 40.0
        tf1 = np.concatenate([identity_poses(1), T])
        tf2 = tf1.copy()
        R = tf1[:, :9].reshape(-1, 3, 3)
        tf2[:, 9:] = np.einsum("nij,j->ni", R, np.array([gap, 0.0, 0.0])) + tf1[:, 9:]
        got = F.distance_mesh_sphere_batch(mesh, tf1, F.Sphere(20.0), tf2, F.DistanceRequest(True))
        assert np.all(np.abs(got.min_distance - true) / true < tol)
        assert np.allclose(got.min_distance, got.min_distance[0], rtol=1e-9)  # rigid-motion invariance


def test_mesh_sphere_distance_large_mesh_and_single_query(oracle):
    from tests.meshes import heightfield

    v, t = heightfield(200, size=10.0, seed=7, amp=0.5)  # 79,202 triangles
    m, o = F.BVHModel.from_arrays(v, t, build_on_device=True), oracle.Model(v, t)
    n = 3000
    rng = np.random.default_rng(107)
    S = identity_poses(n)
    S[:, 9:11] = rng.uniform(-6, 6, size=(n, 2))
    S[:, 11] = rng.uniform(-1.0, 3.0, size=n)
    got = F.distance_mesh_sphere_batch(m, None, F.Sphere(0.3), S, F.DistanceRequest(True))
    Mi = identity_poses(n)
    brute = oracle.distance_mesh_sphere_batch(o, 0.3, Mi, S, brute=True, nthreads=8)
    trav = oracle.distance_mesh_sphere_batch(o, 0.3, Mi, S, nthreads=8)
    _check(got, brute, trav, 0.3)
    assert 50 < (got.min_distance < 0).sum() < n - 50

    # single-query entry points read like the reference (distance-inl.h:92-246), both argument orders
    i = int(np.argmax(got.min_distance))
    sphere = F.Sphere(0.3)
    res = F.DistanceResult()
    d = F.distance(m, F.Transform3.from_pose12(Mi[i]), sphere, F.Transform3.from_pose12(S[i]), F.DistanceRequest(True), res)
    assert d == got.min_distance[i] == res.min_distance and res.b1 == got.b1[i] and res.b2 == -1
    assert res.o1 is m and res.o2 is sphere and np.array_equal(res.nearest_points[0], got.nearest_p1[i])
    res2 = F.DistanceResult()
    d2 = F.distance(sphere, F.Transform3.from_pose12(S[i]), m, F.Transform3.from_pose12(Mi[i]), F.DistanceRequest(), res2)
    assert d2 == d and res2.o1 is m and np.array_equal(res2.nearest_points[1], got.nearest_p2[i])
    # a satisfied result returns early (orientedBVHShapeDistance, distance_func_matrix-inl.h:268)
    res3 = F.DistanceResult(0.0)
    assert F.distance(m, None, sphere, F.Transform3.from_pose12(S[i]), F.DistanceRequest(), res3) == 0.0


def test_mesh_sphere_distance_tiny_meshes(oracle):
    """1- and 2-triangle models (the root is a leaf / has two leaf children), sphere over face, edge and vertex."""
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0.5]], dtype=np.float64)
    C = np.array([[0.25, 0.25, 2.0], [0.5, -1.0, 0.0], [-1.0, -1.0, 0.3], [2.0, 2.0, 2.0], [0.3, 0.3, 0.05], [0.6, 0.6, -3.0]])
    S = identity_poses(len(C))
    S[:, 9:] = C
    for tris in (np.array([[0, 1, 2]], np.int32), np.array([[0, 1, 2], [1, 3, 2]], np.int32)):
        m, o = F.BVHModel.from_arrays(v, tris), oracle.Model(v, tris)
        for radius in (0.0, 0.1, 0.75):
            got = F.distance_mesh_sphere_batch(m, None, F.Sphere(radius), S, F.DistanceRequest(True))
            brute = oracle.distance_mesh_sphere_batch(o, radius, identity_poses(len(C)), S, brute=True)
            assert np.array_equal(got.min_distance, brute["min_distance"]), (len(tris), radius)
            pos = brute["min_distance"] > 0
            assert np.array_equal(got.b1[pos], brute["b1"][pos])
            assert got.nearest_p1[pos].tobytes() == brute["p1"][pos].tobytes()
            assert got.nearest_p2[pos].tobytes() == brute["p2"][pos].tobytes()
