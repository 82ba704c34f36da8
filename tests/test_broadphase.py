"""Broadphase pair generation feeding the batched narrowphase.  Oracle = the brute-force manager
(NaiveCollisionManager::collide(other), broadphase_bruteforce-inl.h:182-205) with the default callback
(default_broadphase_callbacks.h:84-103); parity on the pair set like test_fcl_broadphase_collision_1.cpp:136-170, here
even on the pair ORDER (the brute-force manager's) and on the per-pair contact counts."""
import numpy as np
import pytest

import fcl_b200 as F
from fcl_b200.poses import random_poses
from fcl_b200.workloads import box_mesh, uv_sphere


def _scene(seed, n1, n2):
    """n1 small meshes of three kinds scattered in a box, n2 of two kinds; every fifth pose is a pure translation (the
    identity-rotation branch of CollisionObject::computeAABB)."""
    meshes = [uv_sphere(0.4, 8, 6), box_mesh(0.5, 0.2, 0.3), uv_sphere(0.25, 6, 5), box_mesh(0.15, 0.6, 0.2)]
    rng = np.random.default_rng(seed)

    def poses(n, s):
        P = random_poses(n, seed=s)
        P[:, 9:] = rng.uniform(-3.0, 3.0, size=(n, 3))
        P[::5, :9] = np.eye(3).reshape(9)
        return P

    g1 = rng.integers(0, 3, size=n1).astype(np.int32)
    g2 = rng.integers(2, 4, size=n2).astype(np.int32)
    return meshes, g1, poses(n1, seed + 1), g2, poses(n2, seed + 2)


def test_oracle_local_and_world_aabb(oracle):
    v, t = box_mesh(0.5, 0.2, 0.3, center=(1.0, 0.0, -2.0))
    m = oracle.Model(v, t)
    P = random_poses(3, seed=4)
    P[0, :9] = np.eye(3).reshape(9)
    r = oracle.broadphase([m], [0, 0, 0], P, [0], P[:1], narrowphase=False)
    # identity rotation: the local box translated (collision_object-inl.h:120-123)
    assert np.allclose(r["aabb1"][0], np.concatenate([[0.5, -0.2, -2.3], [1.5, 0.2, -1.7]]) + np.tile(P[0, 9:], 2))
    # rotated: the cube of half side aabb_radius around tf * aabb_center (:124-130)
    radius = np.sqrt(0.25 + 0.04 + 0.09)
    c = P[1, :9].reshape(3, 3) @ np.array([1.0, 0.0, -2.0]) + P[1, 9:]
    assert np.allclose(r["aabb1"][1], np.concatenate([c - radius, c + radius]))
    assert [0, 0] in r["pairs"].tolist()  # an object overlaps itself


def test_oracle_pairs_equal_python_enumeration(oracle):
    meshes, g1, P1, g2, P2 = _scene(3, 60, 40)
    models = [oracle.Model(v, t) for v, t in meshes]
    r = oracle.broadphase(models, g1, P1, g2, P2, num_max_contacts=4, enable_contact=True)
    a1 = r["aabb1"]
    a2 = oracle.broadphase(models, g2, P2, g1[:1], P1[:1], narrowphase=False)["aabb1"]
    want = [(i, j) for i in range(len(g1)) for j in range(len(g2))
            if not (a1[i, :3] > a2[j, 3:]).any() and not (a1[i, 3:] < a2[j, :3]).any()]
    assert r["pairs"].tolist() == [list(p) for p in want] and 20 < len(want) < len(g1) * len(g2)
    # the narrowphase of a culled pair is plain fcl::collide on that pair
    for k in range(0, len(want), 7):
        i, j = want[k]
        c = oracle.collide_batch(models[g1[i]], models[g2[j]], P1[i:i + 1], P2[j:j + 1], 4, True)
        assert c["counts"][0] == r["counts"][k]
    assert (r["counts"] > 0).any() and (r["counts"] == 0).any()


@pytest.mark.gpu
def test_gpu_broadphase_pairs_and_counts_equal_the_brute_force_manager(oracle):
    meshes, g1, P1, g2, P2 = _scene(5, 700, 300)
    omodels = [oracle.Model(v, t) for v, t in meshes]
    ref = oracle.broadphase(omodels, g1, P1, g2, P2, num_max_contacts=6, enable_contact=False, nthreads=8)
    geoms = [F.BVHModel.from_arrays(v, t) for v, t in meshes]
    A, B = F.NaiveCollisionManager(), F.DynamicAABBTreeCollisionManager()
    A.registerObjects([F.CollisionObject(geoms[g], F.Transform3.from_pose12(p)) for g, p in zip(g1, P1)])
    B.registerObjects([F.CollisionObject(geoms[g], F.Transform3.from_pose12(p)) for g, p in zip(g2, P2)])
    A.setup()
    B.setup()
    got = A.collide_batch(B, F.CollisionRequest(6, False), pair_capacity=64)  # too small on purpose: grows and reruns
    assert got.aabb1.tobytes() == ref["aabb1"].tobytes()
    assert np.array_equal(got.pairs, ref["pairs"]) and len(got.pairs) > 2000
    assert np.array_equal(got.num_contacts, ref["counts"]) and (got.num_contacts > 0).sum() > 50
    # the reference's callback protocol: DefaultCollisionFunction accumulates into one result and stops the evaluation
    data = F.DefaultCollisionData(F.CollisionRequest(10, False))
    A.collide(B, data, F.DefaultCollisionFunction)
    assert data.done and data.result.numContacts() == 10
    want, total = [], 0
    for (i, j), c in zip(ref["pairs"], ref["counts"]):  # contacts come from the culled pairs in visiting order
        if total >= 10:
            break
        if c:
            total += min(int(c), 10 - total)
            want.append((i, j))
    seen = []
    for c in data.result.getContacts():
        key = (id(c.o1), id(c.o2))
        if not seen or seen[-1] != key:
            seen.append(key)
    assert len(seen) <= len(want)
    # empty managers
    E = F.NaiveCollisionManager()
    assert len(E.collide_batch(B).pairs) == 0 and len(A.collide_batch(E).pairs) == 0
    # self-collision (collide(cdata, callback), broadphase_bruteforce-inl.h:140-160): every unordered pair once, in order
    sp = A.self_pairs(F.CollisionRequest(6, False))
    aabb = got.aabb1
    ov = np.all((aabb[:, None, :3] <= aabb[None, :, 3:]) & (aabb[None, :, :3] <= aabb[:, None, 3:]), axis=2)
    iu, ju = np.nonzero(np.triu(ov, 1))
    assert np.array_equal(sp.pairs, np.stack([iu, ju], axis=1).astype(np.int32)) and len(sp.pairs) > 50
    visited = []
    A.collide(visited, lambda o1, o2, data: data.append((A.objs.index(o1), A.objs.index(o2))) or False)
    assert visited == [tuple(p) for p in sp.pairs.tolist()]
    visited2 = []
    A.collide(A, visited2, lambda o1, o2, data: data.append(1) or len(data) >= 7)  # same manager on both sides; early stop
    assert len(visited2) == 7


@pytest.mark.gpu
def test_gpu_manager_distance_equals_all_pairs_minimum(oracle):
    """NaiveCollisionManager::distance(other, cdata, DefaultDistanceFunction) (broadphase_bruteforce-inl.h:208-233,
    default_broadphase_callbacks.h:190-211) and its batched form against the oracle's distance over ALL pairs."""
    meshes, g1, P1, g2, P2 = _scene(9, 40, 30)
    P2[:, 9:] += 8.0  # apart: a positive minimum
    omodels = [oracle.Model(v, t) for v, t in meshes]
    geoms = [F.BVHModel.from_arrays(v, t) for v, t in meshes]
    A, B = F.NaiveCollisionManager(), F.NaiveCollisionManager()
    A.registerObjects([F.CollisionObject(geoms[g], F.Transform3.from_pose12(p)) for g, p in zip(g1, P1)])
    B.registerObjects([F.CollisionObject(geoms[g], F.Transform3.from_pose12(p)) for g, p in zip(g2, P2)])
    want = np.inf
    for ga in np.unique(g1):
        for gb in np.unique(g2):
            ia, ib = np.nonzero(g1 == ga)[0], np.nonzero(g2 == gb)[0]
            t1 = np.repeat(P1[ia], len(ib), axis=0)
            t2 = np.tile(P2[ib], (len(ia), 1))
            want = min(want, oracle.distance_batch(omodels[ga], omodels[gb], t1, t2, True, 2, nthreads=8)["min_distance"].min())
    assert want > 0
    got = A.distance_batch(B, chunk=64)  # small chunks: the AABB lower bounds end the evaluation early
    assert got.min_distance == want and got.evaluated < len(g1) * len(g2) // 2
    assert A.distance_batch(B).min_distance == want
    i, j = got.pair
    d = F.DistanceResult()
    assert F.distance(A.objs[i], B.objs[j], F.DistanceRequest(True), d) == want
    assert abs(np.linalg.norm(got.nearest_points[0] - got.nearest_points[1]) - want) <= 1e-9 * max(1.0, want)
    data = F.DefaultDistanceData(F.DistanceRequest(True))
    A.distance(B, data, F.DefaultDistanceFunction)
    assert data.result.min_distance == want
    # touching / colliding sets: the callback protocol stops at the first pair with distance <= 0
    P2[:, 9:] -= 8.0
    B2 = F.NaiveCollisionManager()
    B2.registerObjects([F.CollisionObject(geoms[g], F.Transform3.from_pose12(p)) for g, p in zip(g2, P2)])
    assert A.distance_batch(B2).min_distance == 0.0
    data = F.DefaultDistanceData()
    A.distance(B2, data, F.DefaultDistanceFunction)
    assert data.result.min_distance == 0.0
