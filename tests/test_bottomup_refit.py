"""Bottom-up refit -- the reference's default endReplaceModel() (BVH_model.h:128, BVH_model-inl.h:952-1037):
fit3 at the leaves (math/bv/utility-inl.h:92-117, 208-230), OBB::operator+ (OBB-inl.h:161-369) and RSS::operator+
(RSS-inl.h:313-371) above.  CPU part: the oracle's restatement against closed-form answers and against the product's
independently structured host code (csrc/bvh_merge.cuh); GPU part: the level-by-level kernel against both, and queries
on the refitted model."""
import numpy as np
import pytest

import fcl_b200 as F
from fcl_b200 import _capi
from fcl_b200.poses import random_poses

KEYS = ("axis", "obb_To", "obb_ext", "rss_axis", "rss_To", "rss_l", "rss_r")


def vol(axis=np.eye(3), To=(0, 0, 0), ext=(1, 1, 1), raxis=None, rTo=(0, 0, 0), l=(1, 1), r=0.5):
    raxis = axis if raxis is None else raxis
    return np.concatenate([np.asarray(axis, float).reshape(9), np.asarray(To, float), np.asarray(ext, float),
                           np.asarray(raxis, float).reshape(9), np.asarray(rTo, float), np.asarray(l, float), [float(r)]])


def unpack(v):
    return dict(axis=v[0:9].reshape(3, 3), obb_To=v[9:12], obb_ext=v[12:15], rss_axis=v[15:24].reshape(3, 3), rss_To=v[24:27],
                rss_l=v[27:29], rss_r=v[29])


def test_fit3_known_answers(oracle):
    """A 3-4-5 right triangle in the z = 2 plane: first axis along the longest edge (the hypotenuse), third axis the
    normal, zero thickness, OBB centred on the hypotenuse's bounding strip; RSS with zero radius covering the triangle."""
    p = np.array([[0.0, 0, 2], [3, 0, 2], [0, 4, 2]])
    f = unpack(oracle.fit3_obbrss(p))
    A = f["axis"]
    assert np.allclose(A.T @ A, np.eye(3), atol=1e-15)
    assert np.allclose(np.abs(A[:, 2]), [0, 0, 1])  # e0 x e1 normalised
    # e = (p1 - p2, p2 - p3, p3 - p1) = ((-3,0,0), (3,-4,0), (0,4,0)): the longest is e[1], normalised (0.6, -0.8, 0)
    assert np.allclose(A[:, 0], [0.6, -0.8, 0.0], atol=1e-15)
    assert np.allclose(f["obb_ext"], [2.5, 1.2, 0.0], atol=1e-14)  # half the hypotenuse, half the height over it
    loc = (p - f["obb_To"]) @ A
    assert (np.abs(loc) <= f["obb_ext"] + 1e-12).all()
    assert f["rss_r"] == 0.0 and np.allclose(f["rss_l"], [5.0, 2.4], atol=1e-14)
    assert f["rss_axis"].tobytes() == f["axis"].tobytes()
    # every vertex lies on the rectangle
    q = (p - f["rss_To"]) @ f["rss_axis"]
    assert (q[:, 0] >= -1e-12).all() and (q[:, 0] <= f["rss_l"][0] + 1e-12).all()
    assert (q[:, 1] >= -1e-12).all() and (q[:, 1] <= f["rss_l"][1] + 1e-12).all() and np.allclose(q[:, 2], 0, atol=1e-12)


def test_obb_merge_known_answers(oracle):
    """OBB::operator+: two unit cubes 10 apart along x merge by merge_largedist (centre distance 10 > 2 (1 + 1)): first
    axis along the centre difference, extents (6, 1, 1) around the midpoint.  A box merged with an equal box at the same
    place goes through merge_smalldist (quaternion average = the same orientation) and comes back unchanged."""
    a, b = vol(To=(0, 0, 0)), vol(To=(10, 0, 0))
    m = unpack(oracle.merge_obbrss(a, b))
    assert np.allclose(np.abs(m["axis"][:, 0]), [1, 0, 0], atol=1e-15)
    assert np.allclose(m["obb_To"], [5, 0, 0], atol=1e-12)
    assert np.isclose(m["obb_ext"][0], 6.0) and np.allclose(np.sort(m["obb_ext"][1:]), [1, 1], atol=1e-12)
    c = np.cos(0.3), np.sin(0.3)
    R = np.array([[c[0], -c[1], 0], [c[1], c[0], 0], [0, 0, 1.0]])
    s = vol(axis=R, To=(1, 2, 3), ext=(3, 2, 1))
    m = unpack(oracle.merge_obbrss(s, s))
    assert np.allclose(m["axis"], R, atol=1e-15) and np.allclose(m["obb_To"], [1, 2, 3], atol=1e-12)
    assert np.allclose(m["obb_ext"], [3, 2, 1], atol=1e-12)
    # the reference's `else if` (OBB-inl.h:339-342): the first corner of the first box is never tried as a minimum.
    # A small first box inside a big second one is harmless; a big first box whose low corner is the overall minimum on
    # every axis loses that corner -- the merged box is smaller than it.
    big, small = vol(ext=(4, 4, 4)), vol(ext=(1, 1, 1))
    ok = unpack(oracle.merge_obbrss(small, big))
    assert np.allclose(ok["obb_ext"], [4, 4, 4], atol=1e-12)
    far = vol(To=(3, 3, 3), ext=(1, 1, 1))  # centre distance 5.2 <= 2 (4 + 1): merge_smalldist
    quirk = unpack(oracle.merge_obbrss(big, far))
    lo = quirk["obb_To"] - quirk["obb_ext"]
    assert np.allclose(lo, [-4, -4, -4], atol=1e-12)  # saved here only because other corners share each minimum


def test_rss_merge_keeps_the_reference_quirks(oracle):
    """RSS::operator+ (RSS-inl.h:313-371): the third axis of the result is the cross product of *this*'s in-plane axes,
    whatever the new in-plane axes are (:364)."""
    c, s = np.cos(0.7), np.sin(0.7)
    R1 = np.array([[1, 0, 0], [0, c, -s], [0, s, c]])
    R2 = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
    a = vol(raxis=R1, rTo=(0, 0, 0), l=(2, 1), r=0.25)
    b = vol(raxis=R2, rTo=(1, 1, 1), l=(1, 3), r=0.5)
    m = unpack(oracle.merge_obbrss(a, b))
    assert np.allclose(m["rss_axis"][:, 2], np.cross(R1[:, 0], R1[:, 1]), atol=1e-15)
    m2 = unpack(oracle.merge_obbrss(b, a))
    assert np.allclose(m2["rss_axis"][:, 2], np.cross(R2[:, 0], R2[:, 1]), atol=1e-15)
    # the in-plane axes are orthonormal (rows of an orthogonal matrix)
    assert np.isclose(np.linalg.norm(m["rss_axis"][:, 0]), 1) and np.isclose(m["rss_axis"][:, 0] @ m["rss_axis"][:, 1], 0, atol=1e-12)


@pytest.mark.parametrize("split", [F.SPLIT_METHOD_MEAN, F.SPLIT_METHOD_MEDIAN, F.SPLIT_METHOD_BV_CENTER])
def test_host_bottomup_refit_matches_oracle_bitwise(oracle, env_rob_npz, split):
    """endReplaceModel() with the reference's defaults (refit = true, bottomup = true): the product's host model
    (csrc/bvh_merge.cuh, flat arrays and point functors) and the oracle (fcl_oracle_bvh.cpp, the reference's structure)
    give the same bits for every node of env.obj and rob.obj."""
    rng = np.random.default_rng(10 + split)
    for v, t in env_rob_npz:
        o, m = oracle.Model(v, t, split), F.BVHModel.from_arrays(v, t, split)
        for step in range(2):  # a second refit starts from volumes whose OBB and RSS axes already differ
            v2 = v + rng.normal(0, 3.0, size=v.shape)
            assert o.refit_bottomup(v2) == 0
            assert m.beginReplaceModel() == F.BVH_OK and m.replaceSubModel(v2) == F.BVH_OK
            assert m.endReplaceModel() == F.BVH_OK
            a, b = o.arrays(), m.node_arrays()
            for k in KEYS:
                assert a[k].tobytes() == np.ascontiguousarray(b[k]).tobytes(), (split, step, k)
        leaves = a["first_child"] < 0
        assert (a["axis"][leaves] == a["rss_axis"][leaves]).all() and (a["axis"][~leaves] != a["rss_axis"][~leaves]).any()
        # a top-down refit brings the shared axes back
        assert m.beginReplaceModel() == F.BVH_OK and m.replaceSubModel(v) == F.BVH_OK
        assert m.endReplaceModel(True, False) == F.BVH_OK and o.refit_topdown(v) == 0
        a, b = o.arrays(), m.node_arrays()
        for k in KEYS:
            assert a[k].tobytes() == np.ascontiguousarray(b[k]).tobytes(), (split, "topdown", k)
        assert (b["axis"] == b["rss_axis"]).all()


def test_bottomup_refit_protocol(capfd):
    m = F.BVHModel.from_arrays(np.array([[0.0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 1]]), np.array([[0, 1, 2], [1, 2, 3]], np.int32))
    assert m.endReplaceModel() == F.BVH_ERR_BUILD_OUT_OF_SEQUENCE
    assert m.beginReplaceModel() == F.BVH_OK
    assert m.replaceSubModel(np.zeros((3, 3))) == F.BVH_OK
    assert m.endReplaceModel() == F.BVH_ERR_INCORRECT_DATA  # vertex count differs (BVH_model-inl.h:602-606)
    capfd.readouterr()


@pytest.mark.gpu
def test_device_bottomup_refit_is_bit_exact(oracle, env_rob_npz):
    """The refit kernel (one launch per tree height) writes the oracle's bits into HBM; collide on the refitted models
    is the reference's, contact for contact, with the exact box test (traversal 1) -- the merged boxes need not contain
    their subtrees (the reference's `else if`), so only a traversal that evaluates the SAME box test visits the same
    nodes -- and distance is the reference's with its own visiting order and exact RSS distance (traversal 0)."""
    import torch

    (ev, et), (rv, rt) = env_rob_npz
    rng = np.random.default_rng(5)
    env, rob = F.BVHModel.from_arrays(ev, et), F.BVHModel.from_arrays(rv, rt)
    oenv, orob = oracle.Model(ev, et), oracle.Model(rv, rt)
    env.device_model()
    rob.device_model()
    ev2 = ev * (1.0 + 0.02 * np.sin(ev[:, [1, 2, 0]] / 500.0))
    rv2 = rv + rng.normal(0, 5.0, size=rv.shape)
    for m, o, v2 in ((env, oenv, ev2), (rob, orob, rv2)):
        assert m.beginReplaceModel() == F.BVH_OK and m.replaceSubModel(v2) == F.BVH_OK
        assert m.endReplaceModel() == F.BVH_OK
        assert o.refit_bottomup(v2) == 0
        dev, ref, host = m.download_device_arrays(), o.arrays(), m.node_arrays()
        for k in KEYS:
            assert dev[k].tobytes() == ref[k].tobytes(), k
            assert np.ascontiguousarray(host[k]).tobytes() == ref[k].tobytes(), k
    P = random_poses(4000, seed=23)
    refc = oracle.collide_batch(oenv, orob, P, None, 50, True, nthreads=8)
    rd = oracle.distance_batch(oenv, orob, P, None, True, 2, nthreads=8)
    try:
        _capi.set_option("traversal", 1)
        got = F.collide_batch(env, P, rob, None, F.CollisionRequest(50, True), contact_capacity=50 * len(P))
        assert np.array_equal(got.num_contacts, refc["counts"]) and got.contacts.tobytes() == refc["contacts"].tobytes()
        _capi.set_option("traversal", 0)
        gd = F.distance_batch(env, P, rob, None, F.DistanceRequest(True))
        assert np.array_equal(gd.min_distance, rd["min_distance"])
        assert np.array_equal(gd.b1, rd["b1"]) and np.array_equal(gd.b2, rd["b2"])
    finally:
        _capi.set_option("traversal", 3)
    # default kernels: conservative steering tests visit a superset of the nodes, so nothing the reference reports is lost
    got3 = F.collide_batch(env, P, rob, None, F.CollisionRequest(100000, False), contact_capacity=4_000_000)
    ref3 = oracle.collide_batch(oenv, orob, P, None, 100000, False, nthreads=8)
    assert (got3.num_contacts >= ref3["counts"]).all()
    # an upload made AFTER the refit carries the separate axes too
    env2 = F.BVHModel.from_arrays(ev, et)
    assert env2.beginReplaceModel() == F.BVH_OK and env2.replaceSubModel(ev2) == F.BVH_OK and env2.endReplaceModel() == F.BVH_OK
    dev = env2.download_device_arrays()
    ref = oenv.arrays()
    for k in KEYS:
        assert dev[k].tobytes() == ref[k].tobytes(), k
    # device-resident vertices
    rv3 = rv2 + 1.0
    rob.refit_device(torch.from_numpy(rv3).cuda(), bottomup=True)
    F.sync_status()
    assert orob.refit_bottomup(rv3) == 0
    dev, ref = rob.download_device_arrays(), orob.arrays()
    for k in KEYS:
        assert dev[k].tobytes() == ref[k].tobytes(), k


@pytest.mark.gpu
def test_device_bottomup_refit_on_a_deep_tree(oracle):
    from tests.meshes import heightfield

    v, t = heightfield(60, size=10.0, seed=3, amp=0.6)  # 7200 triangles
    m, o = F.BVHModel.from_arrays(v, t, build_on_device=True), oracle.Model(v, t)
    m.device_model()
    v2 = v + np.random.default_rng(2).normal(0, 0.05, size=v.shape)
    assert m.beginReplaceModel() == F.BVH_OK and m.replaceSubModel(v2) == F.BVH_OK and m.endReplaceModel() == F.BVH_OK
    assert o.refit_bottomup(v2) == 0
    dev, ref = m.download_device_arrays(), o.arrays()
    for k in KEYS:
        assert dev[k].tobytes() == ref[k].tobytes(), k
