"""CPU-only: the C-ABI library loads and exports every declared symbol, the host-side mirror of
the reference interface behaves like the reference (build protocol, return codes, request/result
semantics), the product's BVH builder reproduces the oracle's tree bit for bit, and compute
entry points fail loudly without a GPU (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

import fcl_b200 as F
from fcl_b200 import _capi
from fcl_b200.poses import random_poses, splitmix64_uniform
from tests.meshes import box_mesh, random_soup, uv_sphere

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "fclgpu.h")).read()
    declared = set(re.findall(r"\b(fclgpu_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    L = _capi.lib()
    for s in sorted(declared):
        assert hasattr(L, s), f"libfclgpu.so does not export {s}"
    assert declared == set(_capi.SYMBOLS)
    assert L.fclgpu_abi_version() == 3


def test_build_protocol_and_return_codes(capfd):
    v, t = box_mesh(1, 2, 3)
    m = F.BVHModel()
    assert m.endModel() == F.BVH_ERR_BUILD_OUT_OF_SEQUENCE          # endModel before beginModel
    assert m.beginModel() == F.BVH_OK
    assert m.endModel() == F.BVH_ERR_BUILD_EMPTY_MODEL              # nothing added
    assert m.addSubModel(v, t) == F.BVH_OK
    assert m.addTriangle([0, 0, 0], [1, 0, 0], [0, 1, 0]) == F.BVH_OK
    assert m.endModel() == F.BVH_OK
    assert m.build_state == F.BVH_BUILD_STATE_PROCESSED
    assert m.num_tris == 13 and m.num_vertices == 11 and m.getNumBVs() == 25
    assert m.addSubModel(v, t) == F.BVH_ERR_BUILD_OUT_OF_SEQUENCE   # after endModel
    assert m.beginModel() == F.BVH_ERR_BUILD_OUT_OF_SEQUENCE        # non-empty model: cleared + warning
    assert m.build_state == F.BVH_BUILD_STATE_EMPTY and m.num_tris == 0
    assert m.beginModel() == F.BVH_OK
    assert "BVH Warning" in capfd.readouterr().err
    bad = F.BVHModel()
    bad.beginModel()
    bad.addSubModel(v, t + 100)
    assert bad.endModel() == F.BVH_ERR_INCORRECT_DATA


@pytest.mark.parametrize("split", [0, 1, 2])
def test_product_bvh_builder_matches_oracle_bitwise(oracle, env_rob_npz, split):
    meshes = list(env_rob_npz) + [uv_sphere(20, 16, 16), box_mesh(1, 2, 3), random_soup(300, 3), ([[0, 0, 0], [1, 0, 0], [0, 1, 0]], [[0, 1, 2]])]
    # degenerate inputs: 64 copies of one triangle (every split ties), collinear vertices (zero-area triangles,
    # rank-deficient covariance), two far-apart triangles
    line = np.arange(40, dtype=np.float64)[:, None] * np.array([[1.0, 2.0, 3.0]])
    meshes += [(np.array([[0.0, 0, 0], [1, 0, 0], [0, 1, 0]]), np.tile(np.array([[0, 1, 2]], np.int32), (64, 1))),
               (line, np.stack([np.arange(38), np.arange(38) + 1, np.arange(38) + 2], axis=1).astype(np.int32)),
               (np.array([[0.0, 0, 0], [1, 0, 0], [0, 1, 0], [5, 5, 5], [6, 5, 5], [5, 6, 5]]), np.array([[0, 1, 2], [3, 4, 5]], np.int32))]
    for v, t in meshes:
        got = F.BVHModel.from_arrays(v, t, split).node_arrays()
        ref = oracle.Model(v, t, split).arrays()
        for k in ref:
            assert got[k].tobytes() == ref[k].tobytes(), k
        tv = np.asarray(v, dtype=np.float64)[np.asarray(t).reshape(-1)].reshape(-1, 9)
        assert got["tri_verts"].tobytes() == tv.tobytes()


def test_request_result_semantics():
    req = F.CollisionRequest()
    assert req.num_max_contacts == 1 and req.enable_contact is False and req.enable_cost is False
    res = F.CollisionResult()
    assert not res.isCollision() and not req.isSatisfied(res)
    res.addContact(F.Contact(None, None, 3, 4))
    assert res.isCollision() and res.numContacts() == 1 and req.isSatisfied(res)
    assert res.getContact(7).b1 == 3  # out-of-range index returns the last contact (collision_result-inl.h:96-101)
    res.clear()
    assert res.numContacts() == 0
    a, b = F.Contact(b1=1, b2=5), F.Contact(b1=1, b2=7)
    assert a < b and not (b < a) and F.Contact(b1=0, b2=9) < a
    d = F.DistanceResult()
    assert d.min_distance == F.DBL_MAX and d.b1 == -1
    d.update(2.0, None, None, 1, 2)
    d.update(2.0, None, None, 5, 6)  # strict '>' keeps the first
    assert (d.b1, d.b2) == (1, 2)
    assert F.DistanceRequest().isSatisfied(F.DistanceResult(0.0))


def test_collide_zero_max_contacts_and_unsupported(capfd):
    v, t = box_mesh(1, 1, 1)
    m = F.BVHModel.from_arrays(v, t)
    res = F.CollisionResult()
    assert F.collide(m, F.Transform3(), m, F.Transform3(), F.CollisionRequest(0), res) == 0
    assert "should stop early" in capfd.readouterr().err
    assert F.collide(object(), None, m, None, F.CollisionRequest(), res) == 0
    assert F.distance(object(), None, m, None, F.DistanceRequest(), F.DistanceResult()) == F.DBL_MAX


def test_no_cpu_fallback_without_device():
    if _capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    v, t = box_mesh(1, 1, 1)
    m = F.BVHModel.from_arrays(v, t)
    with pytest.raises(F.FclGpuError) as ei:
        F.collide_batch(m, random_poses(4), m, None, F.CollisionRequest())
    assert ei.value.code == _capi.ERR_NO_DEVICE
    with pytest.raises(F.FclGpuError):
        F.distance_batch(m, random_poses(4), m, None, F.DistanceRequest())
    with pytest.raises(F.FclGpuError) as ei:
        F.distance_mesh_sphere_batch(m, random_poses(4), F.Sphere(1.0), random_poses(4, seed=2), F.DistanceRequest(True))
    assert ei.value.code == _capi.ERR_NO_DEVICE
    with pytest.raises(F.FclGpuError):
        F.collide_mesh_sphere_batch(m, random_poses(4), F.Sphere(1.0), random_poses(4, seed=2), F.CollisionRequest())


def test_device_trim_on_an_untouched_device_is_a_no_op():
    """fclgpu_device_trim before anything ran on the device: nothing to release, no CUDA call, no error (also without a GPU)."""
    import ctypes as C

    n = C.c_int64(-1)
    assert _capi.lib().fclgpu_device_trim(7, C.byref(n)) == 0 and n.value == 0
    assert _capi.lib().fclgpu_device_trim(7, None) == 0


def test_pose_generator_is_reproducible_and_shardable():
    a = random_poses(1000, seed=1)
    b = np.concatenate([random_poses(400, seed=1, start=0), random_poses(600, seed=1, start=400)])
    assert a.tobytes() == b.tobytes()
    R = a[:, :9].reshape(-1, 3, 3)
    assert np.allclose(R @ np.transpose(R, (0, 2, 1)), np.eye(3), atol=1e-12)
    assert (a[:, 9:11] >= -3000).all() and (a[:, 9:11] < 3000).all() and (a[:, 11] >= 0).all()
    u = splitmix64_uniform(1, 4)
    assert (u >= 0).all() and (u < 1).all()
    # splitmix64 known answer: seed 0, first output 0xE220A8397B1DCDAF
    assert splitmix64_uniform(0, 1)[0] == (0xE220A8397B1DCDAF >> 11) / 2.0**53


def test_transform_conversion():
    rng = np.random.default_rng(0)
    M = np.eye(4)
    M[:3, :3] = rng.normal(size=(3, 3))
    M[:3, 3] = rng.normal(size=3)
    tf = F.Transform3.from_matrix4_colmajor(M.T.reshape(16))  # column-major storage
    assert np.array_equal(tf.R, M[:3, :3]) and np.array_equal(tf.t, M[:3, 3])


def test_replace_model_topdown_refit_matches_oracle(oracle, env_rob_npz, capfd):
    """beginReplaceModel / replaceSubModel / endReplaceModel(refit=True, bottomup=False)
    (BVH_model-inl.h:521-620, refitTree_topdown :1064-1076): host refit is bit-identical to the oracle's."""
    (ev, et), (rv, rt) = env_rob_npz
    rng = np.random.default_rng(2)
    for v, t in ((rv, rt), random_soup(400, 9), uv_sphere(5.0, 12, 12)):
        v = np.asarray(v, dtype=np.float64)
        m, o = F.BVHModel.from_arrays(v, t), oracle.Model(v, t)
        assert [a.tobytes() for a in m.partition()] == [a.tobytes() for a in o.partition()]
        v2 = v + rng.normal(0, 0.05 * np.abs(v).max(), size=v.shape)
        assert m.beginReplaceModel() == F.BVH_OK
        assert m.replaceSubModel(v2[: len(v2) // 2]) == F.BVH_OK
        assert m.endReplaceModel(True, False) == F.BVH_ERR_INCORRECT_DATA  # vertex count mismatch (:602-606)
        assert m.replaceSubModel(v2[len(v2) // 2:]) == F.BVH_OK
        assert m.endReplaceModel(True, False) == F.BVH_OK
        assert o.refit_topdown(v2) == 0
        got, ref = m.node_arrays(), o.arrays()
        for k in ref:
            assert got[k].tobytes() == ref[k].tobytes(), k
        # refit=False rebuilds the tree: equal to a fresh build on the new vertices
        assert m.beginReplaceModel() == F.BVH_OK and m.replaceSubModel(v2) == F.BVH_OK
        assert m.endReplaceModel(False) == F.BVH_OK
        fresh = oracle.Model(v2, t).arrays()
        got = m.node_arrays()
        for k in fresh:
            assert got[k].tobytes() == fresh[k].tobytes(), k
    fresh = F.BVHModel()
    assert fresh.beginReplaceModel() == F.BVH_ERR_BUILD_EMPTY_PREVIOUS_FRAME
    assert fresh.replaceVertex([0, 0, 0]) == F.BVH_ERR_BUILD_OUT_OF_SEQUENCE
    assert "no previous frame" in capfd.readouterr().err
