"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on
the same seeded inputs.  Bit-exact for verdicts, contact lists (ids, positions, normals, depths,
in the reference's DFS order) and minimum distances; nearest points / closest ids are exact for
the thread-per-query traversal and within 1e-6 relative for the warp-front traversal."""
import numpy as np
import pytest

import fcl_b200 as F
from fcl_b200 import _capi
from fcl_b200.poses import identity_poses, random_poses
from tests.meshes import box_mesh, random_soup, uv_sphere

pytestmark = pytest.mark.gpu

INT_MAX = 2**31 - 1
REL_TOL = 1e-6  # north_star: minimum distance and nearest points within 1e-6 relative


@pytest.fixture(scope="module")
def models(env_rob_npz, oracle):
    (ev, et), (rv, rt) = env_rob_npz
    return (F.BVHModel.from_arrays(ev, et), F.BVHModel.from_arrays(rv, rt)), (oracle.Model(ev, et), oracle.Model(rv, rt))


@pytest.fixture(params=[0, 1, 2, 3, 4], ids=["thread", "front64", "front32", "pooled", "pooled2"])
def traversal(request):
    _capi.set_option("traversal", request.param)
    yield request.param
    _capi.set_option("traversal", 3)  # library default


def _contacts_equal(got, ref):
    assert np.array_equal(got.num_contacts, ref["counts"])
    assert np.array_equal(got.offsets, ref["offsets"])
    g, r = got.contacts, ref["contacts"]
    assert len(g) == len(r)
    assert np.array_equal(g["b1"], r["b1"]) and np.array_equal(g["b2"], r["b2"])
    return g, r


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5])
def test_cfg1_binary_verdict_10k(models, oracle, traversal, seed):
    (env, rob), (oenv, orob) = models
    P = random_poses(10000, seed=seed)
    got = F.collide_batch(env, P, rob, None, F.CollisionRequest(), want_contacts=False, stats=True)
    ref = oracle.collide_batch(oenv, orob, P, None, 1, False, nthreads=8)
    assert np.array_equal(got.num_contacts, ref["counts"])
    assert 3000 < got.num_contacts.sum() < 5000
    if traversal == 0:  # same visiting order as the reference => same work counters
        assert np.array_equal(got.n_bv, ref["n_bv"]) and np.array_equal(got.n_leaf, ref["n_leaf"])


def test_cfg1_first_contact_pair_ids(models, oracle, traversal):
    (env, rob), (oenv, orob) = models
    P = random_poses(10000, seed=6)
    got = F.collide_batch(env, P, rob, None, F.CollisionRequest(1, False), contact_capacity=len(P))
    ref = oracle.collide_batch(oenv, orob, P, None, 1, False, nthreads=8)
    _contacts_equal(got, ref)


def test_cfg3_contacts_max100(models, oracle, traversal):
    (env, rob), (oenv, orob) = models
    P = random_poses(20000, seed=7)
    got = F.collide_batch(env, P, rob, None, F.CollisionRequest(100, True), contact_capacity=100 * len(P))
    ref = oracle.collide_batch(oenv, orob, P, None, 100, True, nthreads=8)
    g, r = _contacts_equal(got, ref)
    assert g.tobytes() == r.tobytes()  # positions, normals, depths bit-exact, DFS order, truncated at 100
    assert (got.num_contacts == 100).sum() > 100  # truncation is exercised


def test_compact_contact_records(models, oracle):
    """fclgpu_collision_request::contact_format (extension): the same contact lists -- same order, same truncation -- as
    8-byte id records or 40-byte single-precision records (the FP64 values rounded to nearest float)."""
    (env, rob), (oenv, orob) = models
    P = random_poses(6000, seed=17)
    ref = oracle.collide_batch(oenv, orob, P, None, 100, True, nthreads=8)
    full = F.collide_batch(env, P, rob, None, F.CollisionRequest(100, True), contact_capacity=100 * len(P))
    assert full.contacts.tobytes() == ref["contacts"].tobytes()
    for pinned in (False, True):
        ids = F.collide_batch(env, P, rob, None, F.CollisionRequest(100, True), contact_capacity=100 * len(P), contact_format=F.CONTACT_IDS,
                              pinned=pinned)
        assert np.array_equal(ids.num_contacts, ref["counts"]) and np.array_equal(ids.offsets, ref["offsets"])
        assert ids.contacts.dtype.itemsize == 8
        assert np.array_equal(ids.contacts["b1"], ref["contacts"]["b1"]) and np.array_equal(ids.contacts["b2"], ref["contacts"]["b2"])
        f32 = F.collide_batch(env, P, rob, None, F.CollisionRequest(100, True), contact_capacity=100 * len(P), contact_format=F.CONTACT_F32,
                              pinned=pinned)
        c = f32.contacts
        assert c.dtype.itemsize == 40 and np.array_equal(f32.num_contacts, ref["counts"])
        assert np.array_equal(c["b1"], ref["contacts"]["b1"]) and np.array_equal(c["b2"], ref["contacts"]["b2"])
        for k, rk in (("normal", "normal"), ("pos", "pos"), ("penetration_depth", "depth")):
            assert c[k].tobytes() == ref["contacts"][rk].astype(np.float32).tobytes(), k
    # the other contact paths do not write compact records
    _capi.set_option("contact_order", 1)
    try:
        with pytest.raises(F.FclGpuError):
            F.collide_batch(env, P[:10], rob, None, F.CollisionRequest(100, True), contact_capacity=1000, contact_format=F.CONTACT_IDS)
    finally:
        _capi.set_option("contact_order", 0)


def test_exhaustive_contact_pair_sets(models, oracle, traversal):
    (env, rob), (oenv, orob) = models
    P = random_poses(4000, seed=8)
    got = F.collide_batch(env, P, rob, None, F.CollisionRequest(INT_MAX, True), contact_capacity=1024 * len(P) // 4)
    ref = oracle.collide_batch(oenv, orob, P, None, INT_MAX, True, nthreads=8)
    g, r = _contacts_equal(got, ref)
    assert g.tobytes() == r.tobytes()
    assert got.num_contacts.max() > 300
    # and the pair set of a few poses equals brute force over all 2180 x 216 triangle pairs
    for i in np.flatnonzero(got.num_contacts > 0)[:5]:
        c = got.contacts_of(i)
        brute = set(map(tuple, oracle.brute_collide(oenv, orob, P[i]).tolist()))
        assert set(zip(c["b1"].tolist(), c["b2"].tolist())) == brute


def test_binary_mode_many_contacts(models, oracle, traversal):
    (env, rob), (oenv, orob) = models
    P = random_poses(3000, seed=9)
    got = F.collide_batch(env, P, rob, None, F.CollisionRequest(50, False), contact_capacity=50 * len(P))
    ref = oracle.collide_batch(oenv, orob, P, None, 50, False, nthreads=8)
    _contacts_equal(got, ref)


def test_cfg2_distance_nearest_points(models, oracle, traversal):
    (env, rob), (oenv, orob) = models
    P = random_poses(20000, seed=10)
    got = F.distance_batch(env, P, rob, None, F.DistanceRequest(True), stats=True)
    ref = oracle.distance_batch(oenv, orob, P, None, True, 2, nthreads=8)
    assert np.array_equal(got.min_distance, ref["min_distance"])  # bit-exact minimum distance
    pos = ref["min_distance"] > 0
    assert pos.sum() > 5000 and (~pos).sum() > 5000
    scale = np.maximum(1.0, np.abs(ref["p1"][pos]).max(axis=1, keepdims=True))
    assert (np.abs(got.nearest_p1[pos] - ref["p1"][pos]) <= REL_TOL * scale).all()
    assert (np.abs(got.nearest_p2[pos] - ref["p2"][pos]) <= REL_TOL * scale).all()
    if traversal == 0:  # identical visiting order: ids, points and counters are identical too
        assert np.array_equal(got.b1, ref["b1"]) and np.array_equal(got.b2, ref["b2"])
        assert got.nearest_p1.tobytes() == ref["p1"].tobytes() and got.nearest_p2.tobytes() == ref["p2"].tobytes()
        assert np.array_equal(got.n_bv, ref["n_bv"]) and np.array_equal(got.n_leaf, ref["n_leaf"])
    else:
        # front traversals: the closest ids and the nearest points are the reference's, bit for bit, wherever the closest
        # pair is unique; where another pair is named it must be an EXACT tie -- its own triangle distance, recomputed by
        # the oracle's triDistance at that pose, equals the minimum (adjacent triangles sharing the closest vertex / edge)
        # (env.obj's faces are pairs of coplanar triangles and its boxes share edges and corners, so a closest point on a
        # shared edge or vertex -- an exact tie between neighbours -- is the common case, not the exception)
        same = (got.b1 == ref["b1"]) & (got.b2 == ref["b2"])
        assert same[pos].mean() > 0.2
        both = pos & same
        assert got.nearest_p1[both].tobytes() == ref["p1"][both].tobytes()
        assert got.nearest_p2[both].tobytes() == ref["p2"][both].tobytes()
        (ev, et), (rv, rt) = (oenv.verts, oenv.tris), (orob.verts, orob.tris)
        other = np.where(pos & ~same)[0]
        for q in other[:300]:
            R, t = P[q, :9].reshape(3, 3), P[q, 9:]
            S = ev[et[got.b1[q]]] @ R.T + t   # model 1 carries the pose, model 2 sits at the identity
            d, _, _ = oracle.tri_distance(S, rv[rt[got.b2[q]]])
            # (recomputed in the world frame instead of model 1's: rounding differs around 1e-13 of the coordinates)
            assert abs(d - ref["min_distance"][q]) <= 1e-9 * max(1.0, ref["min_distance"][q]), (q, d, ref["min_distance"][q])


def test_distance_without_nearest_points(models, oracle, traversal):
    (env, rob), (oenv, orob) = models
    P = random_poses(2000, seed=11)
    got = F.distance_batch(env, P, rob, None, F.DistanceRequest(False))
    ref = oracle.distance_batch(oenv, orob, P, None, False, 2, nthreads=8)
    assert np.array_equal(got.min_distance, ref["min_distance"])


def test_both_objects_posed(models, oracle, traversal):
    """General (tf1, tf2): the collide and distance paths build the relative pose with different
    rounding sequences (SURVEY 8 a3 vs a11); both must match."""
    (env, rob), (oenv, orob) = models
    P1 = random_poses(3000, seed=12)
    P2 = random_poses(3000, seed=13)
    P2[:, 9:] = P1[:, 9:] + 0.2 * (P2[:, 9:] - 1500.0)
    got = F.collide_batch(env, P1, rob, P2, F.CollisionRequest(20, True), contact_capacity=20 * len(P1))
    ref = oracle.collide_batch(oenv, orob, P1, P2, 20, True, nthreads=8)
    g, r = _contacts_equal(got, ref)
    assert g.tobytes() == r.tobytes()
    assert got.num_contacts.sum() > 1000
    gd = F.distance_batch(env, P1, rob, P2, F.DistanceRequest(True))
    rd = oracle.distance_batch(oenv, orob, P1, P2, True, 2, nthreads=8)
    assert np.array_equal(gd.min_distance, rd["min_distance"])
    # model2 posed, model1 identity
    got = F.collide_batch(env, None, rob, P2, F.CollisionRequest(5, True), contact_capacity=5 * len(P2))
    ref = oracle.collide_batch(oenv, orob, None, P2, 5, True, nthreads=8)
    g, r = _contacts_equal(got, ref)
    assert g.tobytes() == r.tobytes()


def test_edge_cases(oracle, traversal):
    """Single-triangle models (the root pair is a leaf pair), n = 1, n = 0, num_max_contacts = 0."""
    tri = ([[0, 0, 0], [1, 0, 0], [0, 1, 0]], [[0, 1, 2]])
    a, b = F.BVHModel.from_arrays(*tri), F.BVHModel.from_arrays(*tri)
    oa, ob = oracle.Model(*tri), oracle.Model(*tri)
    P = identity_poses(3)
    P[1, 9:] = [0.2, 0.2, 0.0]   # coplanar overlap
    P[2, 9:] = [0.0, 0.0, 3.0]   # apart
    got = F.collide_batch(a, P, b, None, F.CollisionRequest(10, True), contact_capacity=30)
    ref = oracle.collide_batch(oa, ob, P, None, 10, True)
    g, r = _contacts_equal(got, ref)
    assert g.tobytes() == r.tobytes()
    gd = F.distance_batch(a, P, b, None, F.DistanceRequest(True))
    rd = oracle.distance_batch(oa, ob, P, None, True)
    assert np.array_equal(gd.min_distance, rd["min_distance"]) and gd.min_distance[2] == 3.0
    one = F.collide_batch(a, P[2:], b, None, F.CollisionRequest(), want_contacts=False)
    assert one.num_contacts.tolist() == [0]
    zero = F.collide_batch(a, P, b, None, F.CollisionRequest(0, True), contact_capacity=8)
    assert zero.num_contacts.sum() == 0 and zero.offsets.tolist() == [0, 0, 0, 0]
    empty = F.collide_batch(a, np.zeros((0, 12)), b, None, F.CollisionRequest(), want_contacts=False)
    assert len(empty.num_contacts) == 0
    assert len(F.distance_batch(a, np.zeros((0, 12)), b, None, F.DistanceRequest()).min_distance) == 0


def test_contact_capacity_overflow_is_reported(models):
    (env, rob), _ = models
    P = random_poses(2000, seed=14)
    with pytest.raises(F.FclGpuError) as ei:
        F.collide_batch(env, P, rob, None, F.CollisionRequest(100, True), contact_capacity=10)
    assert ei.value.code == _capi.ERR_CONTACT_OVERFLOW


def test_synthetic_meshes(oracle, traversal):
    """Tessellated spheres (test_fcl_shape_mesh_consistency.cpp:57-80 analytic check) and random soups."""
    v, t = uv_sphere(20, 16, 16)
    s1, s2 = F.BVHModel.from_arrays(v, t), F.BVHModel.from_arrays(v, t)
    tf2 = identity_poses(2)
    tf2[0, 9] = 50.0
    tf2[1, 9] = 30.0
    d = F.distance_batch(s1, identity_poses(2), s2, tf2, F.DistanceRequest(True))
    assert abs(d.min_distance[0] - 10.0) < 1.5 and d.min_distance[1] == 0.0
    c = F.collide_batch(s1, identity_poses(2), s2, tf2, F.CollisionRequest(), want_contacts=False)
    assert c.num_contacts.tolist() == [0, 1]
    va, ta = random_soup(700, 1, scale=3.0)
    vb, tb = random_soup(300, 2, scale=1.0)
    A, B = F.BVHModel.from_arrays(va, ta), F.BVHModel.from_arrays(vb, tb)
    OA, OB = oracle.Model(va, ta), oracle.Model(vb, tb)
    P = random_poses(3000, seed=15, extents=(-4, -4, -4, 4, 4, 4))
    got = F.collide_batch(A, P, B, None, F.CollisionRequest(INT_MAX, True), contact_capacity=600 * len(P))
    ref = oracle.collide_batch(OA, OB, P, None, INT_MAX, True, nthreads=8)
    g, r = _contacts_equal(got, ref)
    assert g.tobytes() == r.tobytes()
    gd = F.distance_batch(A, P, B, None, F.DistanceRequest(True))
    rd = oracle.distance_batch(OA, OB, P, None, True, 2, nthreads=8)
    assert np.array_equal(gd.min_distance, rd["min_distance"])


def test_single_query_api_reads_like_the_reference(models, oracle):
    """test/test_fcl_collision.cpp:1107-1150 (test_collide_func) and test_fcl_distance.cpp style."""
    (env, rob), (oenv, orob) = models
    P = random_poses(40, seed=16)
    for i in range(len(P)):
        pose = F.Transform3.from_pose12(P[i])
        request = F.CollisionRequest(INT_MAX, True)
        result = F.CollisionResult()
        n = F.collide(env, pose, rob, F.Transform3.Identity(), request, result)
        ref = oracle.collide_batch(oenv, orob, P[i:i + 1], None, INT_MAX, True)
        assert n == ref["counts"][0] == result.numContacts()
        pairs = sorted((c.b1, c.b2) for c in result.getContacts())
        assert pairs == sorted(zip(ref["contacts"]["b1"].tolist(), ref["contacts"]["b2"].tolist()))
        # results accumulate; a satisfied request returns early
        assert F.collide(env, pose, rob, F.Transform3.Identity(), F.CollisionRequest(1), result) == n or n == 0
        dres = F.DistanceResult()
        d = F.distance(env, pose, rob, F.Transform3.Identity(), F.DistanceRequest(True), dres)
        rd = oracle.distance_batch(oenv, orob, P[i:i + 1], None, True)
        assert d == rd["min_distance"][0]
        if d > 0:
            assert abs(np.linalg.norm(dres.nearest_points[0] - dres.nearest_points[1]) - d) < 1e-6 * max(1, d)


def test_full_size_properties_1m(models, oracle):
    """BASELINE full size (1M poses): device-resident API; shard-additivity, idempotence,
    and a random sample checked against the oracle."""
    import torch

    (env, rob), (oenv, orob) = models
    n = 1_000_000
    P = random_poses(n, seed=1)
    dP = torch.from_numpy(P).cuda()
    cnt = torch.zeros(n, dtype=torch.int32, device="cuda")
    cnt2 = torch.zeros(n, dtype=torch.int32, device="cuda")
    req = F.CollisionRequest()
    F.collide_batch_device(env, dP, rob, None, req, cnt)
    h = n // 2
    F.collide_batch_device(env, dP[:h], rob, None, req, cnt2[:h])
    F.collide_batch_device(env, dP[h:], rob, None, req, cnt2[h:])
    F.sync_status()
    assert torch.equal(cnt, cnt2)
    dist = torch.zeros(n, dtype=torch.float64, device="cuda")
    dist2 = torch.zeros(n, dtype=torch.float64, device="cuda")
    p1 = torch.zeros(n, 3, dtype=torch.float64, device="cuda")
    p2 = torch.zeros(n, 3, dtype=torch.float64, device="cuda")
    F.distance_batch_device(env, dP, rob, None, F.DistanceRequest(True), dist, p1, p2)
    F.distance_batch_device(env, dP, rob, None, F.DistanceRequest(False), dist2)
    F.sync_status()
    assert torch.equal(dist, dist2)
    sep = (p1 - p2).norm(dim=1)
    pos = dist > 0
    assert torch.allclose(sep[pos], dist[pos], rtol=1e-9, atol=1e-9)
    # colliding (triangle intersection) implies zero triangle distance for the overwhelming majority
    col = cnt > 0
    assert ((dist[col] == 0).float().mean() > 0.999) and ((cnt[dist > 1e-6] == 0).all())
    idx = np.random.default_rng(0).choice(n, 3000, replace=False)
    ref = oracle.collide_batch(oenv, orob, P[idx], None, 1, False, nthreads=8)
    assert np.array_equal(cnt.cpu().numpy()[idx], ref["counts"])
    rd = oracle.distance_batch(oenv, orob, P[idx], None, True, 2, nthreads=8)
    assert np.array_equal(dist.cpu().numpy()[idx], rd["min_distance"])


def test_mesh_mesh_consistency_across_split_methods(env_rob_npz, oracle):
    """FCL_COLLISION.mesh_mesh / FCL_DISTANCE.mesh_distance on the GPU path
    (test/test_fcl_collision.cpp:792-886, test/test_fcl_distance.cpp:177-298): with
    num_max_contacts = INT_MAX and enable_contact the sorted contact (b1, b2) sets are identical for
    MEAN / MEDIAN / BV_CENTER trees and for "pose carried by the node" vs "pose baked into the
    vertices"; distances agree (the reference tolerates 1e-3, here they are equal)."""
    (ev, et), (rv, rt) = env_rob_npz
    P = random_poses(10, seed=17)  # the reference uses 10 poses
    per_split, dists = [], []
    for split in (F.SPLIT_METHOD_MEAN, F.SPLIT_METHOD_MEDIAN, F.SPLIT_METHOD_BV_CENTER):
        env, rob = F.BVHModel.from_arrays(ev, et, split), F.BVHModel.from_arrays(rv, rt, split)
        r = F.collide_batch(env, P, rob, None, F.CollisionRequest(INT_MAX, True), contact_capacity=4000, grow_on_overflow=True)
        per_split.append([sorted(set(zip(r.contacts_of(i)["b1"].tolist(), r.contacts_of(i)["b2"].tolist()))) for i in range(len(P))])
        dists.append(F.distance_batch(env, P, rob, None, F.DistanceRequest(True)).min_distance)
    assert per_split[0] == per_split[1] == per_split[2]
    assert np.array_equal(dists[0], dists[1]) and np.array_equal(dists[0], dists[2])
    assert any(len(s) > 0 for s in per_split[0])
    # collide_Test2 style: transform env's vertices on the host, identity pose
    for i in range(len(P)):
        R, t = P[i, :9].reshape(3, 3), P[i, 9:]
        env_t = F.BVHModel.from_arrays(ev @ R.T + t, et)
        r = F.collide_batch(env_t, identity_poses(1), F.BVHModel.from_arrays(rv, rt), None, F.CollisionRequest(INT_MAX, True),
                            contact_capacity=4000, grow_on_overflow=True)
        got = sorted(set(zip(r.contacts["b1"].tolist(), r.contacts["b2"].tolist())))
        # baking the pose rounds the vertices, so razor-edge pairs may flip: require near-identity
        a, b = set(got), set(per_split[0][i])
        assert len(a ^ b) <= max(2, len(b) // 100)


def test_collision_object_overloads(models, oracle):
    (env, rob), (oenv, orob) = models
    P = random_poses(5, seed=18)
    for i in range(len(P)):
        o1 = F.CollisionObject(env, F.Transform3.from_pose12(P[i]))
        o2 = F.CollisionObject(rob)
        res = F.CollisionResult()
        n = F.collide(o1, o2, F.CollisionRequest(1000, True), res)
        ref = oracle.collide_batch(oenv, orob, P[i:i + 1], None, 1000, True)
        assert n == ref["counts"][0]
        dres = F.DistanceResult()
        d = F.distance(o1, o2, F.DistanceRequest(True), dres)
        assert d == oracle.distance_batch(oenv, orob, P[i:i + 1], None, True)["min_distance"][0]


def test_device_refit_topdown_is_bit_exact(oracle, env_rob_npz):
    """on-device top-down refit.  After moving the vertices the BV records in HBM
    (recomputed by refit_nodes_kernel) equal the oracle's refitted BVs bit for bit, and queries on the
    refitted model match the oracle's."""
    (ev, et), (rv, rt) = env_rob_npz
    rng = np.random.default_rng(3)
    env, rob = F.BVHModel.from_arrays(ev, et), F.BVHModel.from_arrays(rv, rt)
    oenv, orob = oracle.Model(ev, et), oracle.Model(rv, rt)
    env.device_model()
    rob.device_model()  # upload before the refit so that the device copies are refitted by the kernel
    ev2 = ev * (1.0 + 0.02 * np.sin(ev[:, [1, 2, 0]] / 500.0))  # smooth deformation
    rv2 = rv + rng.normal(0, 10.0, size=rv.shape)
    for m, o, v2 in ((env, oenv, ev2), (rob, orob, rv2)):
        assert m.beginReplaceModel() == F.BVH_OK and m.replaceSubModel(v2) == F.BVH_OK
        assert m.endReplaceModel(True, False) == F.BVH_OK
        assert o.refit_topdown(v2) == 0
        dev, ref = m.download_device_arrays(), o.arrays()
        for k in ("axis", "obb_To", "obb_ext", "rss_To", "rss_l", "rss_r"):
            assert dev[k].tobytes() == ref[k].tobytes(), k
        host = m.node_arrays()
        assert dev["tri_verts"].tobytes() == host["tri_verts"].tobytes()
    P = random_poses(5000, seed=19)
    got = F.collide_batch(env, P, rob, None, F.CollisionRequest(50, True), contact_capacity=50 * len(P))
    refc = oracle.collide_batch(oenv, orob, P, None, 50, True, nthreads=8)
    assert np.array_equal(got.num_contacts, refc["counts"]) and got.contacts.tobytes() == refc["contacts"].tobytes()
    gd = F.distance_batch(env, P, rob, None, F.DistanceRequest(True))
    rd = oracle.distance_batch(oenv, orob, P, None, True, 2, nthreads=8)
    assert np.array_equal(gd.min_distance, rd["min_distance"])
    # device-resident vertices (no host round trip)
    import torch

    rv3 = rv2 + 5.0
    rob.refit_device(torch.from_numpy(rv3).cuda())
    F.sync_status()
    assert orob.refit_topdown(rv3) == 0
    dev, ref = rob.download_device_arrays(), orob.arrays()
    for k in ("axis", "obb_To", "obb_ext", "rss_To", "rss_l", "rss_r"):
        assert dev[k].tobytes() == ref[k].tobytes(), k


def test_device_refit_thread_and_warp_variants_agree(oracle):
    """All refit schedules (one thread per node / warp-cooperative for large nodes / block-cooperative for the
    huge ones) give the oracle's BVs, also on a mesh large enough for deep trees and many large nodes."""
    import torch
    from tests.meshes import heightfield

    v, t = heightfield(60, size=10.0, seed=3, amp=0.6)  # 7200 triangles
    m, o = F.BVHModel.from_arrays(v, t), oracle.Model(v, t)
    m.device_model()
    rng = np.random.default_rng(8)
    for variant in (2, 1, 0):
        _capi.set_option("refit_warp", variant)
        v2 = v + rng.normal(0, 0.05, size=v.shape)
        m.refit_device(torch.from_numpy(v2).cuda())
        F.sync_status()
        assert o.refit_topdown(v2) == 0
        dev, ref = m.download_device_arrays(), o.arrays()
        for k in ("axis", "obb_To", "obb_ext", "rss_To", "rss_l", "rss_r"):
            assert dev[k].tobytes() == ref[k].tobytes(), (variant, k)
    _capi.set_option("refit_warp", 2)


@pytest.mark.parametrize("split", [F.SPLIT_METHOD_MEAN, F.SPLIT_METHOD_MEDIAN, F.SPLIT_METHOD_BV_CENTER])
def test_device_build_equals_host_build(oracle, env_rob_npz, split):
    """BVHModel::endModel() on the device (level-by-level fit / split / swap-partition replay).
    Tree, node numbering, primitive order and every BV equal the host builder's and the oracle's bit for bit;
    queries and a later refit behave like on an uploaded model."""
    from tests.meshes import heightfield, random_soup, uv_sphere

    meshes = [env_rob_npz[0], env_rob_npz[1], heightfield(60, size=10.0, seed=3, amp=0.6), uv_sphere(1.0, 24, 24),
              random_soup(3000, seed=5), random_soup(20000, seed=6, scale=3.0, tri_size=0.05)]
    one_tri = (np.array([[0.0, 0, 0], [1, 0, 0], [0, 1, 0]]), np.array([[0, 1, 2]], np.int32))
    coplanar = heightfield(12, size=4.0, seed=1, amp=0.0)  # flat grid: degenerate covariance, many split ties
    same = (np.array([[0.0, 0, 0], [1, 0, 0], [0, 1, 0]]), np.tile(np.array([[0, 1, 2]], np.int32), (64, 1)))  # every split ties
    line = np.arange(40, dtype=np.float64)[:, None] * np.array([[1.0, 2.0, 3.0]])  # collinear vertices: zero-area triangles
    degenerate = (line, np.stack([np.arange(38), np.arange(38) + 1, np.arange(38) + 2], axis=1).astype(np.int32))
    two = (np.array([[0.0, 0, 0], [1, 0, 0], [0, 1, 0], [5, 5, 5], [6, 5, 5], [5, 6, 5]]), np.array([[0, 1, 2], [3, 4, 5]], np.int32))
    for v, t in meshes + [one_tri, coplanar, same, degenerate, two]:
        host = F.BVHModel.from_arrays(v, t, split)
        dev = F.BVHModel.from_arrays(v, t, split, build_on_device=True)
        assert dev.getNumBVs() == host.getNumBVs()
        a, b = host.node_arrays(), dev.node_arrays()
        assert np.array_equal(a["first_child"], b["first_child"])
        for k in ("axis", "obb_To", "obb_ext", "rss_To", "rss_l", "rss_r", "tri_verts"):
            assert a[k].tobytes() == b[k].tobytes(), k
        for x, y in zip(host.partition(), dev.partition()):
            assert np.array_equal(x, y)
    (ev, et), (rv, rt) = env_rob_npz[0], env_rob_npz[1]
    env = F.BVHModel.from_arrays(ev, et, split, build_on_device=True)
    rob = F.BVHModel.from_arrays(rv, rt, split, build_on_device=True)
    oenv, orob = oracle.Model(ev, et, split), oracle.Model(rv, rt, split)
    P = random_poses(4000, seed=23)
    got = F.collide_batch(env, P, rob, None, F.CollisionRequest(50, True), contact_capacity=50 * len(P))
    refc = oracle.collide_batch(oenv, orob, P, None, 50, True, nthreads=8)
    assert np.array_equal(got.num_contacts, refc["counts"]) and got.contacts.tobytes() == refc["contacts"].tobytes()
    gd = F.distance_batch(env, P, rob, None, F.DistanceRequest(True))
    rd = oracle.distance_batch(oenv, orob, P, None, True, 2, nthreads=8)
    assert np.array_equal(gd.min_distance, rd["min_distance"])
    # refit of a device-built model
    rv2 = rv * 1.01
    assert rob.beginReplaceModel() == F.BVH_OK and rob.replaceSubModel(rv2) == F.BVH_OK
    assert rob.endReplaceModel(True, False) == F.BVH_OK
    assert orob.refit_topdown(rv2) == 0
    d, r = rob.download_device_arrays(), orob.arrays()
    for k in ("axis", "obb_To", "obb_ext", "rss_To", "rss_l", "rss_r"):
        assert d[k].tobytes() == r[k].tobytes(), k


def test_device_build_rejects_what_the_reference_rejects():
    import ctypes as C
    from fcl_b200 import _capi

    L = _capi.lib()

    h = C.c_void_p()
    v = np.zeros((3, 3))
    t = np.array([[0, 1, 5]], np.int32)
    assert L.fclgpu_model_build_obbrss(0, v.ctypes.data, 3, t.ctypes.data, 1, 0, C.byref(h)) == F.BVH_ERR_INCORRECT_DATA
    assert L.fclgpu_model_build_obbrss(0, v.ctypes.data, 3, t.ctypes.data, 0, 0, C.byref(h)) == F.BVH_ERR_BUILD_EMPTY_MODEL
    t[0, 2] = 2
    assert L.fclgpu_model_build_obbrss(0, v.ctypes.data, 3, t.ctypes.data, 1, 7, C.byref(h)) == _capi.ERR_INVALID_ARGUMENT  # no such split rule


@pytest.mark.parametrize("host_chunk", [4096, 1 << 17])
def test_host_api_pinned_pipelines_match_oracle(oracle, env_rob_npz, host_chunk):
    """The host API with page-locked buffers (truly asynchronous copies): streamed-input single launch for
    verdicts, sub-batch pipeline with overlapped contact download, chunked distance pipeline.  Small chunks
    force many chunks / sub-batches per call."""
    import torch

    (ev, et), (rv, rt) = env_rob_npz
    env, rob = F.BVHModel.from_arrays(ev, et), F.BVHModel.from_arrays(rv, rt)
    oenv, orob = oracle.Model(ev, et), oracle.Model(rv, rt)
    n = 30000
    P = torch.from_numpy(random_poses(n, seed=29)).pin_memory().numpy()
    _capi.set_option("host_chunk", host_chunk)
    try:
        for rep in range(2):  # second call reuses flags / staging of the first
            got = F.collide_batch(env, P, rob, None, F.CollisionRequest(), want_contacts=False, pinned=True)
            ref = oracle.collide_batch(oenv, orob, P, None, 1, False, nthreads=8)
            assert np.array_equal(got.num_contacts, ref["counts"])
            gc = F.collide_batch(env, P, rob, None, F.CollisionRequest(20, True), contact_capacity=20 * n, pinned=True)
            rc = oracle.collide_batch(oenv, orob, P, None, 20, True, nthreads=8)
            assert np.array_equal(gc.num_contacts, rc["counts"])
            assert np.array_equal(gc.offsets, rc["offsets"])
            assert gc.contacts.tobytes() == rc["contacts"].tobytes()
            gd = F.distance_batch(env, P, rob, None, F.DistanceRequest(True), pinned=True)
            rd = oracle.distance_batch(oenv, orob, P, None, True, 2, nthreads=8)
            assert np.array_equal(gd.min_distance, rd["min_distance"])
        # capacity overflow keeps the counts exact and reports the status
        with pytest.raises(F.FclGpuError) as ei:
            F.collide_batch(env, P, rob, None, F.CollisionRequest(20, True), contact_capacity=1000, pinned=True)
        assert ei.value.code == _capi.ERR_CONTACT_OVERFLOW
    finally:
        _capi.set_option("host_chunk", 1 << 17)


def test_counts_only_front_kernel_matches_oracle(models, oracle):
    """Counts-only collide through the warp-per-query front kernel (chosen automatically for BVHs beyond the
    caches, forced here): num_contacts = min(#intersecting pairs, num_max_contacts) for every budget."""
    (env, rob), (oenv, orob) = models
    P = random_poses(20000, seed=31)
    _capi.set_option("collide_front", 2)
    try:
        for max_contacts in (1, 7, 100000):
            got = F.collide_batch(env, P, rob, None, F.CollisionRequest(max_contacts, False), want_contacts=False, stats=True)
            ref = oracle.collide_batch(oenv, orob, P, None, max_contacts, False, nthreads=8)
            assert np.array_equal(got.num_contacts, ref["counts"]), max_contacts
        assert got.num_contacts.max() > 100  # the exhaustive budget really counted every pair
        assert (got.n_bv > 0).all()
    finally:
        _capi.set_option("collide_front", 1)


def test_front_kernel_stack_and_leaf_trigger_settings_do_not_change_counts(models, oracle):
    """The front kernel's schedule knobs only move work around: with the smallest stack the library accepts (the rounds are
    depth first almost all the time), with a large one, and with leaf rounds started at 1 / 8 / 32 queued triangle pairs the
    counts equal the oracle's for every budget."""
    (env, rob), (oenv, orob) = models
    P = random_poses(6000, seed=77)
    refs = {m: oracle.collide_batch(oenv, orob, P, None, m, False, nthreads=8)["counts"] for m in (1, 40, 100000)}
    _capi.set_option("collide_front", 2)
    try:
        for cap, trig in ((1, 32), (1, 1), (384, 8), (2048, 1), (100000, 32)):  # the library clamps cap to [depths + 98, 3072]
            _capi.set_option("front_cap", cap)
            _capi.set_option("front_leaf_trigger", trig)
            for max_contacts, ref in refs.items():
                got = F.collide_batch(env, P, rob, None, F.CollisionRequest(max_contacts, False), want_contacts=False)
                assert np.array_equal(got.num_contacts, ref), (cap, trig, max_contacts)
    finally:
        _capi.set_option("collide_front", 1)
        _capi.set_option("front_cap", 0)
        _capi.set_option("front_leaf_trigger", 0)


def test_tiny_and_coincident_models_all_variants(oracle):
    """Edge cases of the traversals: models of 1..9 triangles (single-node trees, leaf-vs-internal pairs), identical
    poses (coincident meshes: every tie-break and touching rule is exercised), far-apart poses (root boxes disjoint).
    Every kernel variant must reproduce the oracle's contacts bit for bit and its distances exactly."""
    rng = np.random.default_rng(61)
    cases = []
    for nt1, nt2 in ((1, 1), (1, 7), (2, 2), (5, 1), (9, 3), (4, 8)):
        v1, t1 = random_soup(nt1, seed=int(rng.integers(1 << 30)), scale=1.0, tri_size=0.8)
        v2, t2 = random_soup(nt2, seed=int(rng.integers(1 << 30)), scale=1.0, tri_size=0.8)
        cases.append((v1, t1, v2, t2))
    v, t = box_mesh(1.0, 0.5, 0.25)
    cases.append((v, t, v, t))  # the same box twice
    P = random_poses(600, seed=67, extents=(-1.5, -1.5, -1.5, 1.5, 1.5, 1.5))
    P[:50] = identity_poses(50)          # coincident
    P[50:100, 9:] += 100.0               # far apart
    try:
        for v1, t1, v2, t2 in cases:
            m1, m2 = F.BVHModel.from_arrays(v1, t1), F.BVHModel.from_arrays(v2, t2)
            o1, o2 = oracle.Model(v1, t1), oracle.Model(v2, t2)
            rc = oracle.collide_batch(o1, o2, P, None, 1000, True, nthreads=4)
            rb = oracle.collide_batch(o1, o2, P, None, 1, False, nthreads=4)
            rd = oracle.distance_batch(o1, o2, P, None, True, 2, nthreads=4)
            for trav in (0, 1, 2, 3, 4):
                _capi.set_option("traversal", trav)
                for front in (1, 2):
                    _capi.set_option("collide_front", front)
                    gb = F.collide_batch(m1, P, m2, None, F.CollisionRequest(), want_contacts=False)
                    assert np.array_equal(gb.num_contacts, rb["counts"]), (trav, front, len(t1), len(t2))
                _capi.set_option("collide_front", 1)
                gc = F.collide_batch(m1, P, m2, None, F.CollisionRequest(1000, True), contact_capacity=1000 * 20, grow_on_overflow=True)
                assert np.array_equal(gc.num_contacts, rc["counts"]), (trav, len(t1), len(t2))
                assert gc.contacts.tobytes() == rc["contacts"].tobytes(), (trav, len(t1), len(t2))
                gd = F.distance_batch(m1, P, m2, None, F.DistanceRequest(True))
                assert np.array_equal(gd.min_distance, rd["min_distance"]), (trav, len(t1), len(t2))
    finally:
        _capi.set_option("traversal", 3)
        _capi.set_option("collide_front", 1)


def test_mesh_sphere_collide_matches_oracle(oracle, env_rob_npz):
    """fcl::collide(BVHModel<OBBRSS>, tf1, Sphere, tf2) on the GPU.  Contacts (triangle id,
    b2 = -1, contact point, normal, depth), their order and the num_max_contacts truncation equal the oracle's bit
    for bit, for fixed and moving meshes and through the single-query entry point."""
    (ev, et), _ = env_rob_npz
    env, oenv = F.BVHModel.from_arrays(ev, et), oracle.Model(ev, et)
    n = 20000
    S = random_poses(n, seed=83)
    M = identity_poses(n)
    M[: n // 2] = random_poses(n // 2, seed=89)
    S[: n // 2, 9:] = np.einsum("nij,nj->ni", M[: n // 2, :9].reshape(-1, 3, 3), S[: n // 2, 9:]) + M[: n // 2, 9:]
    sphere = F.Sphere(350.0)
    for max_contacts, enable_contact in ((1, False), (5, True), (1 << 20, True)):
        got = F.collide_mesh_sphere_batch(env, M, sphere, S, F.CollisionRequest(max_contacts, enable_contact),
                                          contact_capacity=64 * n, grow_on_overflow=True, stats=True)
        ref = oracle.collide_mesh_sphere_batch(oenv, sphere.radius, M, S, max_contacts, enable_contact, nthreads=8)
        assert np.array_equal(got.num_contacts, ref["counts"]), (max_contacts, enable_contact)
        assert np.array_equal(got.offsets, ref["offsets"])
        assert np.array_equal(got.contacts["b1"], ref["contacts"]["b1"]) and (got.contacts["b2"] == -1).all()
        if enable_contact:
            assert got.contacts.tobytes() == ref["contacts"].tobytes()
        assert (got.n_bv > 0).all()
    assert 0.1 * n < (got.num_contacts > 0).sum() < 0.95 * n
    # counts only
    cnt = F.collide_mesh_sphere_batch(env, M, sphere, S, F.CollisionRequest(7, False), want_contacts=False)
    ref7 = oracle.collide_mesh_sphere_batch(oenv, sphere.radius, M, S, 7, False, nthreads=8)
    assert np.array_equal(cnt.num_contacts, ref7["counts"])
    # single-query entry point, result accumulation like the reference
    i = int(np.argmax(got.num_contacts))
    res = F.CollisionResult()
    k = F.collide(env, F.Transform3.from_pose12(M[i]), sphere, F.Transform3.from_pose12(S[i]), F.CollisionRequest(1000, True), res)
    assert k == got.num_contacts[i] == res.numContacts()
    c0 = res.getContact(0)
    assert c0.b2 == -1 and c0.b1 == got.contacts_of(i)[0]["b1"] and c0.penetration_depth <= 0
    # (sphere, mesh) argument order: the reference swaps the arguments and leaves the contacts as they are
    res2 = F.CollisionResult()
    k2 = F.collide(sphere, F.Transform3.from_pose12(S[i]), env, F.Transform3.from_pose12(M[i]), F.CollisionRequest(1000, True), res2)
    assert k2 == k and res2.getContact(0).b1 == c0.b1 and res2.getContact(0).o1 is env
    assert np.array_equal(res2.getContact(k - 1).pos, res.getContact(k - 1).pos)


def test_mesh_sphere_on_a_large_mesh(oracle):
    from tests.meshes import heightfield

    v, t = heightfield(200, size=10.0, seed=7, amp=0.5)  # 79,202 triangles
    m, o = F.BVHModel.from_arrays(v, t, build_on_device=True), oracle.Model(v, t)
    n = 5000
    rng = np.random.default_rng(97)
    S = identity_poses(n)
    S[:, 9:11] = rng.uniform(-5, 5, size=(n, 2))
    S[:, 11] = rng.uniform(-0.5, 1.0, size=n)
    got = F.collide_mesh_sphere_batch(m, None, F.Sphere(0.3), S, F.CollisionRequest(1 << 20, True), contact_capacity=400 * n,
                                      grow_on_overflow=True)
    ref = oracle.collide_mesh_sphere_batch(o, 0.3, None, S, 1 << 20, True, nthreads=8)
    assert np.array_equal(got.num_contacts, ref["counts"]) and got.contacts.tobytes() == ref["contacts"].tobytes()
    assert got.num_contacts.max() > 20


def test_distance_front_survives_unprunable_fronts(oracle):
    """Regression (found by tests/stress/stress_parity.py): a 2,048-triangle heightfield against a 384-triangle sphere, both
    built with the BV-centre split, under poses for which the sorted front of the FP64-bound variant outgrew its
    512-entry shared-memory stack and reported FCLGPU_ERR_STACK_OVERFLOW.  Near the limit the warp now degrades to
    nearest-first depth first.  tests/golden/front_overflow.npz holds the meshes and 32 poses (8 of them overflowed)."""
    import os

    from tests.conftest import GOLDEN

    d = np.load(os.path.join(GOLDEN, "front_overflow.npz"))
    split = int(d["split"])
    m1, m2 = F.BVHModel.from_arrays(d["v1"], d["t1"], split), F.BVHModel.from_arrays(d["v2"], d["t2"], split)
    o1, o2 = oracle.Model(d["v1"], d["t1"], split), oracle.Model(d["v2"], d["t2"], split)
    P1, P2 = np.ascontiguousarray(d["P1"]), np.ascontiguousarray(d["P2"])
    assert int(d["n_bad"]) == 8
    rd = oracle.distance_batch(o1, o2, P1, P2, True, 2, nthreads=4)
    for trav in (0, 1, 2, 3):
        _capi.set_option("traversal", trav)
        try:
            gd = F.distance_batch(m1, P1, m2, P2, F.DistanceRequest(True))
        finally:
            _capi.set_option("traversal", 3)
        assert np.all(np.abs(gd.min_distance - rd["min_distance"]) <= 1e-14 * np.abs(rd["min_distance"])), trav


def test_upload_with_separate_rss_axes(models, oracle):
    """fclgpu_model_create_obbrss2: after the reference's bottom-up refit the RSS of a node no longer shares the OBB's axes
    (OBBRSS::operator+ merges the two volumes separately, OBBRSS-inl.h:95-101).  Here every RSS is re-expressed in another
    frame -- same rectangle, corner moved to the opposite vertex, both in-plane axes negated -- and uploaded next to the
    untouched OBBs: collide (OBB side) is unchanged byte for byte and distance returns the same minima (box tests only
    steer; results come from the triangles)."""
    import ctypes as C

    (env, rob), (oenv, orob) = models
    L = _capi.lib()
    a = env.node_arrays()
    nn, nt = len(a["first_child"]), env.num_tris
    tv = np.zeros((nt, 9))
    check = _capi.check
    h = env._bvh
    fc = np.zeros(nn, np.int32)
    check(L.fclgpu_bvh_get(h, _capi.addr(fc), None, None, None, None, None, None, _capi.addr(tv)))
    axis = np.ascontiguousarray(a["axis"]).reshape(nn, 3, 3)
    rax = axis.copy()
    rax[:, :, 0] *= -1.0
    rax[:, :, 1] *= -1.0
    rTo = a["rss_To"] + axis[:, :, 0] * a["rss_l"][:, :1] + axis[:, :, 1] * a["rss_l"][:, 1:2]
    m2 = C.c_void_p()
    check(L.fclgpu_model_create_obbrss2(0, nn, _capi.addr(fc), _capi.addr(np.ascontiguousarray(axis)), _capi.addr(np.ascontiguousarray(a["obb_To"])),
                                        _capi.addr(np.ascontiguousarray(a["obb_ext"])), _capi.addr(np.ascontiguousarray(rax)),
                                        _capi.addr(np.ascontiguousarray(rTo)), _capi.addr(np.ascontiguousarray(a["rss_l"])),
                                        _capi.addr(np.ascontiguousarray(a["rss_r"])), nt, _capi.addr(tv), C.byref(m2)))
    try:
        n = 4000
        P = random_poses(n, seed=31)
        ref = oracle.distance_batch(oenv, orob, P, None, True, 2, nthreads=8)
        dist = np.zeros(n)
        rq = F.DistanceRequest(False)._c()
        check(L.fclgpu_distance_batch_host(m2, rob.device_model(0), n, _capi.addr(P), None, C.byref(rq), _capi.addr(dist), None, None, None,
                                           None, None, None))
        assert np.array_equal(dist, ref["min_distance"])
        cnt = np.zeros(n, np.int32)
        cq = F.CollisionRequest()._c()
        check(L.fclgpu_collide_batch_host(m2, rob.device_model(0), n, _capi.addr(P), None, C.byref(cq), _capi.addr(cnt), None, 0, None, None, None))
        assert np.array_equal(cnt, oracle.collide_batch(oenv, orob, P, None, 1, False, nthreads=8)["counts"])
    finally:
        L.fclgpu_model_destroy(m2)


def test_concurrent_contact_queries_on_two_streams(models, oracle):
    """Two contact-generating fclgpu_collide_batch calls in flight on DIFFERENT streams of one device (and two distance
    calls): the contact staging is one allocation per device, so the library orders such launches with an event; running
    totals, offsets and the sticky status are per stream.  Both results equal the oracle's; an overflow on one stream is
    reported on that stream only."""
    import torch

    (env, rob), (oenv, orob) = models
    dev = torch.device("cuda", 0)
    n = 6000
    sets = [random_poses(n, seed=61), random_poses(n, seed=62)]
    refs = [oracle.collide_batch(oenv, orob, P, None, 20, True, nthreads=8) for P in sets]
    streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
    outs = []
    for rep in range(3):  # several rounds back to back: the second call is enqueued while the first is running
        outs = []
        for P, st in zip(sets, streams):
            dP = torch.from_numpy(P).to(dev)
            cnt = torch.zeros(n, dtype=torch.int32, device=dev)
            con = torch.zeros(20 * n * 64, dtype=torch.uint8, device=dev)
            off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
            st.wait_stream(torch.cuda.current_stream(dev))
            F.collide_batch_device(env, dP, rob, None, F.CollisionRequest(20, True), cnt, con, off, stream=st)
            outs.append((dP, cnt, con, off))
        for st in streams:
            F.sync_status(0, st)
    for (dP, cnt, con, off), ref in zip(outs, refs):
        res = F.BatchCollisionResult(cnt.cpu().numpy(), np.frombuffer(con.cpu().numpy().tobytes(), dtype=F.api.CONTACT_DTYPE), off.cpu().numpy())
        assert np.array_equal(res.num_contacts, ref["counts"])
        assert res.contacts[: ref["offsets"][-1]].tobytes() == ref["contacts"].tobytes()
    # an overflow on one stream does not show on the other
    small = torch.zeros(10 * 64, dtype=torch.uint8, device=dev)
    dP, cnt, con, off = outs[0]
    F.collide_batch_device(env, dP, rob, None, F.CollisionRequest(20, True), cnt, small, off, stream=streams[0])
    dP1, cnt1, con1, off1 = outs[1]
    F.collide_batch_device(env, dP1, rob, None, F.CollisionRequest(20, True), cnt1, con1, off1, stream=streams[1])
    F.sync_status(0, streams[1])  # clean
    with pytest.raises(F.FclGpuError):
        F.sync_status(0, streams[0])
    F.sync_status(0, streams[0])  # the sticky word was cleared by the read


def test_device_trim_releases_the_workspace_and_the_next_call_regrows_it(models, oracle):
    """fclgpu_device_trim: the growable per-device buffers go back to the device (the reference keeps nothing between calls,
    SURVEY 8b "Ownership"); models survive and the next query allocates what it needs and returns the same results."""
    torch = pytest.importorskip("torch")
    (env, rob), (oenv, orob) = models
    P = random_poses(3000, seed=91)
    req = F.CollisionRequest(50, True)
    a = F.collide_batch(env, P, rob, None, req)
    da = F.distance_batch(env, P, rob, None, F.DistanceRequest(True))
    free0 = torch.cuda.mem_get_info(0)[0]
    released = F.trim_device(0)
    assert released > 0
    assert torch.cuda.mem_get_info(0)[0] >= free0 + released // 2  # the driver really got (most of) it back
    assert F.trim_device(0) == 0  # nothing left to release
    b = F.collide_batch(env, P, rob, None, req)
    db = F.distance_batch(env, P, rob, None, F.DistanceRequest(True))
    assert np.array_equal(a.num_contacts, b.num_contacts) and np.array_equal(a.contacts, b.contacts)
    assert np.array_equal(da.min_distance, db.min_distance)
    ref = oracle.collide_batch(oenv, orob, P, None, 50, True, nthreads=8)
    assert np.array_equal(b.num_contacts, ref["counts"])

