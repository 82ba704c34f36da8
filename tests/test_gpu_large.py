"""Larger BASELINE-style configurations as parity cases (scaled so the CPU oracle finishes in
seconds): cfg4 -- a 7-link arm (7 meshes, ~5k triangles each) against a 200k-triangle scene over
a batch of robot configurations; cfg5 -- two large synthetic meshes whose BVHs are far bigger than
L1/L2-friendly sizes, collide + distance."""
import numpy as np
import pytest

import fcl_b200 as F
from fcl_b200.poses import euler_to_matrix, identity_poses
from tests.meshes import heightfield, noisy_sphere, serial_chain_poses

pytestmark = pytest.mark.gpu
INT_MAX = 2**31 - 1


def test_cfg4_seven_link_arm_vs_200k_scene(oracle):
    sv, st = heightfield(316, size=10.0, seed=1, amp=0.5)  # 199,712 triangles
    scene, oscene = F.BVHModel.from_arrays(sv, st), oracle.Model(sv, st)
    assert scene.num_tris > 199000
    rng = np.random.default_rng(4)
    ncfg = 1500
    q = rng.uniform(-np.pi, np.pi, size=(ncfg, 7))
    link_poses = serial_chain_poses(q)
    total_hits = 0
    for j in range(7):
        lv, lt = noisy_sphere(0.12, 50, 51, seed=10 + j, scale=(2.6, 1.0, 1.0))  # ~5k triangles
        link, olink = F.BVHModel.from_arrays(lv, lt), oracle.Model(lv, lt)
        assert 4500 < link.num_tris < 5500
        P = np.ascontiguousarray(link_poses[:, j])
        got = F.collide_batch(scene, None, link, P, F.CollisionRequest(), want_contacts=False)
        ref = oracle.collide_batch(oscene, olink, None, P, 1, False, nthreads=8)
        assert np.array_equal(got.num_contacts, ref["counts"])
        total_hits += int(got.num_contacts.sum())
        if j == 6:  # end effector: contacts and distance as well
            gc = F.collide_batch(scene, None, link, P[:400], F.CollisionRequest(30, True), contact_capacity=30 * 400)
            rc = oracle.collide_batch(oscene, olink, None, P[:400], 30, True, nthreads=8)
            assert np.array_equal(gc.num_contacts, rc["counts"]) and gc.contacts.tobytes() == rc["contacts"].tobytes()
            gd = F.distance_batch(scene, None, link, P[:400], F.DistanceRequest(True))
            rd = oracle.distance_batch(oscene, olink, None, P[:400], True, 2, nthreads=8)
            assert np.array_equal(gd.min_distance, rd["min_distance"])
    assert total_hits > 100


def test_cfg5_two_large_meshes(oracle):
    va, ta = noisy_sphere(1.0, 230, 221, seed=11)  # ~101k triangles
    vb, tb = noisy_sphere(1.0, 230, 221, seed=12)
    A, B = F.BVHModel.from_arrays(va, ta), F.BVHModel.from_arrays(vb, tb)
    OA, OB = oracle.Model(va, ta), oracle.Model(vb, tb)
    assert A.num_tris > 100000
    rng = np.random.default_rng(5)
    n = 1500
    ang = rng.uniform(0, 2 * np.pi, size=(n, 3))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    P = np.empty((n, 12))
    P[:, :9] = euler_to_matrix(ang[:, 0], ang[:, 1], ang[:, 2]).reshape(n, 9)
    P[:, 9:] = d * rng.uniform(1.5, 3.0, size=(n, 1))  # centre distance 1.5 .. 3 radii
    got = F.collide_batch(A, identity_poses(n), B, P, F.CollisionRequest(), want_contacts=False)
    ref = oracle.collide_batch(OA, OB, identity_poses(n), P, 1, False, nthreads=8)
    assert np.array_equal(got.num_contacts, ref["counts"])
    assert 0.2 * n < got.num_contacts.sum() < 0.8 * n
    gd = F.distance_batch(A, identity_poses(n), B, P, F.DistanceRequest(True))
    rd = oracle.distance_batch(OA, OB, identity_poses(n), P, True, 2, nthreads=8)
    assert np.array_equal(gd.min_distance, rd["min_distance"])
    pos = rd["min_distance"] > 0
    assert (np.abs(gd.nearest_p1[pos] - rd["p1"][pos]) <= 1e-6 * 3).all()
    sub = slice(0, 200)
    gc = F.collide_batch(A, identity_poses(200), B, P[sub], F.CollisionRequest(INT_MAX, True), contact_capacity=200 * 2000,
                         grow_on_overflow=True)
    rc = oracle.collide_batch(OA, OB, identity_poses(200), P[sub], INT_MAX, True, nthreads=8)
    assert np.array_equal(gc.num_contacts, rc["counts"]) and gc.contacts.tobytes() == rc["contacts"].tobytes()


def test_cfg5_full_size_one_million_triangles(oracle):
    """BASELINE cfg5 at its real size: two ~1M-triangle meshes (2M-node BVHs, 0.5 GB of node records each),
    built ON the device; every node record equals the oracle's builder output byte for byte, then collide +
    distance agree with the oracle on a pose sample."""
    va, ta = noisy_sphere(1.0, 710, 705, seed=21)  # 999,680 triangles
    vb, tb = noisy_sphere(1.0, 710, 705, seed=22)
    A = F.BVHModel.from_arrays(va, ta, build_on_device=True)
    B = F.BVHModel.from_arrays(vb, tb, build_on_device=True)
    OA, OB = oracle.Model(va, ta), oracle.Model(vb, tb)
    assert A.num_tris > 990000 and A.getNumBVs() == OA.num_bvs
    dev, ref = A.node_arrays(), OA.arrays()
    assert np.array_equal(dev["first_child"], ref["first_child"])
    for k in ("axis", "obb_To", "obb_ext", "rss_To", "rss_l", "rss_r"):
        assert dev[k].tobytes() == ref[k].tobytes(), k
    del dev, ref
    rng = np.random.default_rng(6)
    n = 600
    ang = rng.uniform(0, 2 * np.pi, size=(n, 3))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    P = np.empty((n, 12))
    P[:, :9] = euler_to_matrix(ang[:, 0], ang[:, 1], ang[:, 2]).reshape(n, 9)
    P[:, 9:] = d * rng.uniform(1.6, 2.6, size=(n, 1))
    got = F.collide_batch(A, identity_poses(n), B, P, F.CollisionRequest(), want_contacts=False)
    ref = oracle.collide_batch(OA, OB, identity_poses(n), P, 1, False, nthreads=8)
    assert np.array_equal(got.num_contacts, ref["counts"])
    assert 0.1 * n < got.num_contacts.sum() < 0.9 * n
    gd = F.distance_batch(A, identity_poses(n), B, P, F.DistanceRequest(True))
    rd = oracle.distance_batch(OA, OB, identity_poses(n), P, True, 2, nthreads=8)
    assert np.array_equal(gd.min_distance, rd["min_distance"])
    sub = slice(0, 60)
    gc = F.collide_batch(A, identity_poses(60), B, P[sub], F.CollisionRequest(INT_MAX, True), contact_capacity=60 * 4000,
                         grow_on_overflow=True)
    rc = oracle.collide_batch(OA, OB, identity_poses(60), P[sub], INT_MAX, True, nthreads=8)
    assert np.array_equal(gc.num_contacts, rc["counts"]) and gc.contacts.tobytes() == rc["contacts"].tobytes()


def test_distance_overflow_area_with_several_host_chunks(oracle):
    """The distance front's HBM overflow area is one allocation per device while fclgpu_distance_batch_host alternates its
    chunks on two streams: with a BVH pair of >= 2^17 nodes (the kSpill instantiation), a SMALL overflow area and small
    host chunks, consecutive chunks' launches overlap in time and must still not disturb each other's parked entries
    (launches that use the area on different streams are ordered by an event).  Distances = the oracle's, bit for bit."""
    from fcl_b200 import _capi

    va, ta = noisy_sphere(1.0, 190, 181, seed=31)  # ~68k triangles -> 137k nodes: >= 2^17 together with the second mesh
    vb, tb = noisy_sphere(1.0, 60, 61, seed=32)
    A, B = F.BVHModel.from_arrays(va, ta), F.BVHModel.from_arrays(vb, tb)
    assert A.getNumBVs() + B.getNumBVs() >= (1 << 17)
    OA, OB = oracle.Model(va, ta), oracle.Model(vb, tb)
    rng = np.random.default_rng(8)
    n = 6000
    ang = rng.uniform(0, 2 * np.pi, size=(n, 3))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    P = np.empty((n, 12))
    P[:, :9] = euler_to_matrix(ang[:, 0], ang[:, 1], ang[:, 2]).reshape(n, 9)
    P[:, 9:] = d * rng.uniform(0.0, 2.6, size=(n, 1))  # from coincident centres (wide, unprunable fronts) to apart
    ref = oracle.distance_batch(OA, OB, identity_poses(n), P, True, 2, nthreads=8)
    try:
        _capi.set_option("dist_spill_entries", 256)  # one block of parked entries per warp: refills happen all the time
        _capi.set_option("host_chunk", 1024)          # six chunks alternating on the two pipeline streams
        for pinned in (False, True):
            got = F.distance_batch(A, identity_poses(n), B, P, F.DistanceRequest(True), pinned=pinned)
            assert np.array_equal(got.min_distance, ref["min_distance"]), pinned
    finally:
        _capi.set_option("dist_spill_entries", 4096)
        _capi.set_option("host_chunk", 1 << 17)
