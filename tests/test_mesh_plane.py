"""Mesh <-> Halfspace / Plane collide: the two other shape-triangle pairs whose test is closed form
in the reference (gjk_solver_libccd-inl.h:502-540 -> halfspace-inl.h:587-621, plane-inl.h:683-759), so they can be pinned
without libccd.  CPU part: the oracle against the reference's own known answers (test_fcl_geometric_shapes.cpp:1295-1394)
and against brute force; the product's device math (host build) against the oracle, bit for bit.  GPU part: the batched
entry point against the oracle."""
import os

import numpy as np
import pytest

import fcl_b200 as F
from fcl_b200.poses import identity_poses, random_poses
from tests import hostcheck as H

IDENT = identity_poses(1)[0]


_FLIP = os.environ.get("FCL_SUM3_ORDER", "0") != "0"  # tests/test_sum_order_hook.py re-runs this file with both sides flipped


def _sum3(a, b, c):
    """The kernel prologue restated below uses the same association order as the library it is compared with."""
    return a + (b + c) if _FLIP else (a + b) + c


def _random_tf(seed):
    return random_poses(1, seed=seed)[0]


def test_reference_known_answers_halfspace_triangle(oracle):
    """test_shapeIntersection_halfspacetriangle (test_fcl_geometric_shapes.cpp:1295-1340)."""
    n, d = (1.0, 0.0, 0.0), 0.0
    tf = _random_tf(5)
    for tri in (np.array([[20.0, 0, 0], [-20, 0, 0], [0, 20, 0]]), np.array([[20.0, 0, 0], [0, -20, 0], [0, 20, 0]])):
        hit, _, _, normal = oracle.plane_tri_intersect("halfspace", n, d, IDENT, tri, IDENT)
        assert hit and np.allclose(normal, [1, 0, 0], atol=1e-9)
        hit, _, _, normal = oracle.plane_tri_intersect("halfspace", n, d, tf, tri, tf)
        assert hit and np.allclose(normal, tf[:9].reshape(3, 3) @ np.array([1.0, 0, 0]), atol=1e-9)
    # a triangle strictly inside the open side x > 0 does not touch the halfspace x <= 0
    assert not oracle.plane_tri_intersect("halfspace", n, d, IDENT, np.array([[1.0, 0, 0], [2, 1, 0], [2, 0, 1]]), IDENT)[0]


def test_reference_known_answers_plane_triangle(oracle):
    """test_shapeIntersection_planetriangle (test_fcl_geometric_shapes.cpp:1348-1388)."""
    n, d = (1.0, 0.0, 0.0), 0.0
    tf = _random_tf(6)
    for tri in (np.array([[20.0, 0, 0], [-20, 0, 0], [0, 20, 0]]), np.array([[20.0, 0, 0], [-0.1, -20, 0], [-0.1, 20, 0]])):
        hit, _, _, _ = oracle.plane_tri_intersect("plane", n, d, IDENT, tri, IDENT)
        assert hit
        assert oracle.plane_tri_intersect("plane", n, d, tf, tri, tf)[0]
    hit, _, _, normal = oracle.plane_tri_intersect("plane", n, d, IDENT, np.array([[20.0, 0, 0], [-0.1, -20, 0], [-0.1, 20, 0]]), IDENT)
    assert hit and np.allclose(normal, [1, 0, 0], atol=1e-9)
    hit, _, _, normal = oracle.plane_tri_intersect("plane", n, d, tf, np.array([[20.0, 0, 0], [-0.1, -20, 0], [-0.1, 20, 0]]), tf)
    assert hit and np.allclose(normal, tf[:9].reshape(3, 3) @ np.array([1.0, 0, 0]), atol=1e-9)
    assert not oracle.plane_tri_intersect("plane", n, d, IDENT, np.array([[1.0, 0, 0], [2, 1, 0], [2, 0, 1]]), IDENT)[0]


def test_device_math_equals_oracle_bit_for_bit(oracle):
    """fcl_b200/csrc/device_math.cuh (host build) takes the shape already transformed; the oracle transforms inside."""
    L = H.lib()
    rng = np.random.default_rng(7)
    hits = {"halfspace": 0, "plane": 0}
    for i in range(3000):
        kind = ("halfspace", "plane")[i % 2]
        nrm = rng.normal(size=3)
        d = float(rng.normal() * 0.5)
        tri = rng.normal(size=(3, 3)) * (1000.0 if i % 7 == 0 else 1.0)
        tf_s, tf_t = _random_tf(1000 + i), _random_tf(5000 + i)
        if i % 7 != 0:
            tf_s[9:] *= 1e-3
            tf_t[9:] *= 1e-3
        hit, cp, depth, normal = oracle.plane_tri_intersect(kind, nrm, d, tf_s, tri, tf_t)
        # what the kernel does: n' = R n0, d' = d0 + n' . t ; vertices to the world
        l = np.sqrt(_sum3(nrm[0] * nrm[0], nrm[1] * nrm[1], nrm[2] * nrm[2]))  # (1 / l) * n like unitNormalTest
        inv_l = 1.0 / l
        n0, d0 = nrm * inv_l, d * inv_l
        Rs, ts = tf_s[:9].reshape(3, 3), tf_s[9:]
        nw = np.array([_sum3(Rs[r, 0] * n0[0], Rs[r, 1] * n0[1], Rs[r, 2] * n0[2]) for r in range(3)])
        dw = d0 + _sum3(nw[0] * ts[0], nw[1] * ts[1], nw[2] * ts[2])
        Rt, tt = tf_t[:9].reshape(3, 3), tf_t[9:]
        V = np.array([[_sum3(Rt[r, 0] * p[0], Rt[r, 1] * p[1], Rt[r, 2] * p[2]) + tt[r] for r in range(3)] for p in tri])
        out = np.zeros(7)
        got = L.hm_plane_tri_intersect(0 if kind == "halfspace" else 1, H.dptr(np.ascontiguousarray(nw)), float(dw),
                                       H.dptr(np.ascontiguousarray(V.reshape(-1))), H.dptr(out))
        assert bool(got) == hit, i
        if hit:
            hits[kind] += 1
            assert out[:3].tobytes() == cp.tobytes() and out[3] == depth and out[4:].tobytes() == normal.tobytes(), (i, kind)
    assert hits["halfspace"] > 300 and hits["plane"] > 300


def test_oracle_traversal_equals_brute_force(oracle, oracle_env_rob):
    """Contacts = every intersecting triangle, in the tree's depth-first leaf order; truncation keeps a prefix."""
    env, _ = oracle_env_rob
    P = random_poses(40, seed=9)
    S = random_poses(40, seed=10)
    for kind in ("halfspace", "plane"):
        nrm, d = (0.3, -0.2, 1.0), 150.0
        full = oracle.collide_mesh_plane_batch(env, kind, nrm, d, P, S, 10**9, True)
        few = oracle.collide_mesh_plane_batch(env, kind, nrm, d, P, S, 7, True)
        some = 0
        for i in range(len(P)):
            ids = full["contacts"]["b1"][full["offsets"][i]:full["offsets"][i + 1]]
            assert sorted(ids.tolist()) == sorted(oracle.brute_mesh_plane(env, kind, nrm, d, P[i], S[i]).tolist())
            assert (full["contacts"]["b2"][full["offsets"][i]:full["offsets"][i + 1]] == -1).all()
            k = few["contacts"][few["offsets"][i]:few["offsets"][i + 1]]
            assert k.tobytes() == full["contacts"][full["offsets"][i]:full["offsets"][i] + len(k)].tobytes()
            assert len(k) == min(7, len(ids))
            some += len(ids) > 0
        assert some > 5


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["halfspace", "plane"])
def test_gpu_mesh_plane_collide_equals_oracle(kind, oracle, env_rob_npz):
    (ev, et), _ = env_rob_npz
    env, oenv = F.BVHModel.from_arrays(ev, et), oracle.Model(ev, et)
    shape = (F.Halfspace if kind == "halfspace" else F.Plane)((0.3, -0.2, 1.0), 150.0)
    n = 4000
    P, S = random_poses(n, seed=21), random_poses(n, seed=22)
    for mx, en in ((10**9, True), (5, True), (1, False), (50, False)):
        got = F.collide_mesh_plane_batch(env, P, shape, S, F.CollisionRequest(mx, en), contact_capacity=64 * n, grow_on_overflow=True)
        ref = oracle.collide_mesh_plane_batch(oenv, kind, (0.3, -0.2, 1.0), 150.0, P, S, mx, en, nthreads=8)
        assert np.array_equal(got.num_contacts, ref["counts"]), (kind, mx, en)
        assert np.array_equal(got.offsets, ref["offsets"])
        assert np.array_equal(got.contacts["b1"], ref["contacts"]["b1"]) and (got.contacts["b2"] == -1).all()
        if en:
            assert got.contacts.tobytes() == ref["contacts"].tobytes(), (kind, mx)
    assert 0.05 * n < (ref["counts"] > 0).sum() < 0.999 * n
    # the fixed-mesh form (tf1 = None) and the single-query entry point, both argument orders
    got = F.collide_mesh_plane_batch(env, None, shape, S[:500], F.CollisionRequest(20, True), contact_capacity=20 * 500)
    ref = oracle.collide_mesh_plane_batch(oenv, kind, (0.3, -0.2, 1.0), 150.0, identity_poses(500), S[:500], 20, True, nthreads=8)
    assert np.array_equal(got.num_contacts, ref["counts"]) and got.contacts.tobytes() == ref["contacts"].tobytes()
    i = int(np.argmax(ref["counts"] > 2))
    res = F.CollisionResult()
    F.collide(env, None, shape, S[i], F.CollisionRequest(20, True), res)
    assert res.numContacts() == ref["counts"][i]
    res2 = F.CollisionResult()
    F.collide(shape, S[i], env, None, F.CollisionRequest(20, True), res2)
    assert res2.numContacts() == res.numContacts() and res2.getContact(0).o1 is env and res2.getContact(0).b2 == -1
