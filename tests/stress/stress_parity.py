"""Randomised differential test GPU vs oracle over many mesh pairs / pose distributions / requests (GPU box only).
    python tests/stress/stress_parity.py [seconds] [seed]     -- exits non-zero on the first mismatch"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import fcl_b200 as F
from fcl_b200 import _capi
from fcl_b200.poses import random_poses, identity_poses
from oracle import pyoracle as O
from tests.meshes import box_mesh, heightfield, noisy_sphere, random_soup, uv_sphere

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 12345)
t_end = time.time() + budget
rounds = checks = 0
ulp_diffs = dist_checks = 0


def make_mesh():
    kind = rng.integers(0, 5)
    s = int(rng.integers(1 << 30))
    if kind == 0:
        return random_soup(int(rng.integers(1, 4000 if rng.integers(0, 8) == 0 else 400)), seed=s, scale=float(rng.uniform(0.5, 3)), tri_size=float(rng.uniform(0.05, 1.0)))
    if kind == 1:
        return noisy_sphere(float(rng.uniform(0.3, 2)), int(rng.integers(4, 30)), int(rng.integers(3, 30)), seed=s, noise=float(rng.uniform(0, 0.2)),
                            scale=tuple(rng.uniform(0.3, 3, size=3)))
    if kind == 2:
        return heightfield(int(rng.integers(2, 100 if rng.integers(0, 8) == 0 else 40)), size=float(rng.uniform(1, 6)), seed=s, amp=float(rng.uniform(0, 1)))
    if kind == 3:
        return box_mesh(*rng.uniform(0.1, 2, size=3))
    return uv_sphere(float(rng.uniform(0.2, 2)), int(rng.integers(3, 20)), int(rng.integers(2, 20)))


while time.time() < t_end:
    (v1, t1), (v2, t2) = make_mesh(), make_mesh()
    split = int(rng.integers(0, 3))
    on_dev = bool(rng.integers(0, 2))  # all three split rules build on the device
    m1 = F.BVHModel.from_arrays(v1, t1, split, build_on_device=on_dev)
    m2 = F.BVHModel.from_arrays(v2, t2, split)
    o1, o2 = O.Model(v1, t1, split), O.Model(v2, t2, split)
    n = int(rng.integers(50, 20000 if rng.integers(0, 6) == 0 else 3000))
    ext = float(rng.uniform(0.5, 6))
    P1 = random_poses(n, seed=int(rng.integers(1 << 30)), extents=(-ext, -ext, -ext, ext, ext, ext))
    P2 = random_poses(n, seed=int(rng.integers(1 << 30)), extents=(-ext, -ext, -ext, ext, ext, ext)) if rng.integers(0, 2) else None
    if rng.integers(0, 4) == 0:
        P1[: n // 4] = identity_poses(n // 4)
        if P2 is not None:
            P2[: n // 4] = identity_poses(n // 4)
    mx = int(rng.choice([1, 2, 3, 10, 100, 1 << 20]))
    ec = bool(rng.integers(0, 2))
    trav = int(rng.choice([0, 1, 2, 3, 3, 3, 4]))
    _capi.set_option("traversal", trav)
    _capi.set_option("collide_front", int(rng.choice([1, 2])))
    scratch = int(rng.choice([2 << 30, 8 << 20]))   # small scratch: contact-mode batches run in several chunks
    chunk = int(rng.choice([1 << 17, 2048]))          # small host chunks: many pipeline stages / sub-batches
    pinned = bool(rng.integers(0, 2))
    _capi.set_option("scratch_bytes", scratch)
    _capi.set_option("host_chunk", chunk)
    tag = dict(nt1=len(t1), nt2=len(t2), split=split, on_dev=on_dev, n=n, ext=ext, mx=mx, ec=ec, trav=trav, scratch=scratch, chunk=chunk, pinned=pinned)
    if on_dev:  # the device-built tree must be the oracle's tree
        a, b = m1.node_arrays(), o1.arrays()
        if not (np.array_equal(a["first_child"], b["first_child"]) and all(a[k].tobytes() == b[k].tobytes() for k in ("axis", "obb_To", "obb_ext", "rss_To", "rss_l", "rss_r"))):
            bad = [k for k in ("first_child", "axis", "obb_To", "obb_ext", "rss_To", "rss_l", "rss_r") if a[k].tobytes() != b[k].tobytes()]
            print("DEVICE BUILD MISMATCH", tag, bad)
            np.savez("gpurun_out/stress_fail_mesh.npz", v=v1, t=t1, split=split)
            sys.exit(2)
    ref = O.collide_batch(o1, o2, P1, P2, mx, ec, nthreads=8)
    got = F.collide_batch(m1, P1, m2, P2, F.CollisionRequest(mx, ec), contact_capacity=max(64 * n, 1024), grow_on_overflow=True, pinned=pinned)
    got_counts, got_contacts = got.num_contacts.copy(), got.contacts.copy()  # pinned results live until the next call
    ok = np.array_equal(got_counts, ref["counts"]) and np.array_equal(got_contacts["b1"], ref["contacts"]["b1"]) and np.array_equal(got_contacts["b2"], ref["contacts"]["b2"])
    if ec:
        ok = ok and got_contacts.tobytes() == ref["contacts"].tobytes()
    cnt = F.collide_batch(m1, P1, m2, P2, F.CollisionRequest(mx, False), want_contacts=False, pinned=pinned)
    refc = ref["counts"] if not ec else O.collide_batch(o1, o2, P1, P2, mx, False, nthreads=8)["counts"]
    ok = ok and np.array_equal(cnt.num_contacts, refc)
    rd = O.distance_batch(o1, o2, P1, P2, True, 2, nthreads=8)
    try:
        gd = F.distance_batch(m1, P1, m2, P2, F.DistanceRequest(True), pinned=pinned)
    except F.FclGpuError as ex:
        print("DISTANCE ERROR", ex, tag)
        np.savez("gpurun_out/stress_fail.npz", v1=v1, t1=t1, v2=v2, t2=t2, P1=P1, P2=P2 if P2 is not None else np.zeros(0), split=split)
        sys.exit(3)
    # Minimum distance: bit-identical with traversal 0 (the reference's visiting order).  The front traversals may
    # meet mathematically tied candidates (adjacent triangles sharing the closest vertex / edge) in another order;
    # their computed distances can differ in the last bit and the reference's tight bounds prune whichever comes
    # second, so allow a few ULP there and count how often it happens.
    if trav == 0:
        ok = ok and np.array_equal(gd.min_distance, rd["min_distance"])
    else:
        ok = ok and bool(np.all(np.abs(gd.min_distance - rd["min_distance"]) <= 1e-14 * np.abs(rd["min_distance"])))
        ulp_diffs += int((gd.min_distance != rd["min_distance"]).sum())
        dist_checks += n
    pos = rd["min_distance"] > 0
    if pos.any():
        scale = np.abs(rd["p1"][pos]).max() + 1.0
        ok = ok and bool((np.abs(gd.nearest_p1[pos] - rd["p1"][pos]) <= 1e-6 * scale).all()) if trav == 0 else ok
    # mesh <-> sphere on the first mesh
    r = float(rng.uniform(0.05, 1.5))
    S = random_poses(n, seed=int(rng.integers(1 << 30)), extents=(-ext, -ext, -ext, ext, ext, ext))
    rs = O.collide_mesh_sphere_batch(o1, r, P1, S, mx, True, nthreads=8)
    gs = F.collide_mesh_sphere_batch(m1, P1, F.Sphere(r), S, F.CollisionRequest(mx, True), contact_capacity=max(64 * n, 1024), grow_on_overflow=True)
    ok = ok and np.array_equal(gs.num_contacts, rs["counts"]) and gs.contacts.tobytes() == rs["contacts"].tobytes()
    # mesh <-> sphere distance: bit-exact against the oracle's pass over all triangles, within tie rounding of its traversal
    _capi.set_option("sphere_leaf_trigger", int(rng.choice([0, 1, 16, 16, 32])))
    _capi.set_option("sphere_bound32", int(rng.integers(0, 2)))
    bs = O.distance_mesh_sphere_batch(o1, r, P1, S, brute=len(t1) * n <= 40_000_000, nthreads=8)
    ds = F.distance_mesh_sphere_batch(m1, P1, F.Sphere(r), S, F.DistanceRequest(True), pinned=pinned)
    if len(t1) * n <= 40_000_000:
        ok = ok and np.array_equal(ds.min_distance, bs["min_distance"])
    else:
        ok = ok and np.array_equal(ds.min_distance < 0, bs["min_distance"] < 0) and bool(
            np.all(np.abs(ds.min_distance - bs["min_distance"]) <= 1e-12 * np.abs(bs["min_distance"])))
    sep = bs["min_distance"] > 0
    if sep.any():
        ok = ok and bool((np.abs(ds.nearest_p1[sep] - bs["p1"][sep]) <= 1e-6 * (np.abs(bs["p1"][sep]).max() + 1.0)).all())
    # continuous collision (translating bodies, conservative advancement): verdict, time of contact, traversal count and
    # contact poses bit-identical
    if rng.integers(0, 2) == 0:
        nc = min(n, 1500)
        E1 = P1[:nc].copy()
        E1[:, 9:] += rng.normal(0, 0.5 * ext, size=(nc, 3))
        B0 = P2[:nc] if P2 is not None else None
        B1 = None
        if B0 is not None:
            B1 = B0.copy()
            B1[:, 9:] += rng.normal(0, 0.2 * ext, size=(nc, 3))
        rc_ = O.continuous_collide_translation_batch(o1, o2, P1[:nc], E1, B0, B1, nthreads=8)
        gc_ = F.continuous_collide_batch(m1, P1[:nc], E1, m2, B0, B1, F.ContinuousCollisionRequest(ccd_solver_type=F.CCDC_CONSERVATIVE_ADVANCEMENT))
        ok = ok and np.array_equal(gc_.is_collide, rc_["is_collide"]) and gc_.time_of_contact.tobytes() == rc_["time_of_contact"].tobytes()
        ok = ok and np.array_equal(gc_.iterations, rc_["iterations"]) and gc_.contact_tf1.tobytes() == rc_["contact_tf1"].tobytes()
        ok = ok and gc_.contact_tf2.tobytes() == rc_["contact_tf2"].tobytes()
        checks += nc
    # bottom-up refit (the reference's default endReplaceModel()): device records = oracle's, then collide with the
    # exact box test (the merged boxes need not contain their subtrees, so only the same test visits the same nodes)
    if ok and rng.integers(0, 3) == 0:
        vn = v1 + rng.normal(0, 0.02, size=v1.shape)
        o1.refit_bottomup(vn)
        m1.device_model()
        if not (m1.beginReplaceModel() == 0 and m1.replaceSubModel(vn) == 0 and m1.endReplaceModel() == 0):
            print("REFIT PROTOCOL ERROR", tag)
            sys.exit(4)
        a, b = m1.download_device_arrays(), o1.arrays()
        bad = [k for k in ("axis", "obb_To", "obb_ext", "rss_axis", "rss_To", "rss_l", "rss_r") if a[k].tobytes() != b[k].tobytes()]
        if bad:
            print("BOTTOM-UP REFIT MISMATCH", tag, bad)
            np.savez("gpurun_out/stress_fail_refit.npz", v=v1, t=t1, vn=vn, split=split)
            sys.exit(5)
        _capi.set_option("traversal", 1)
        nr = min(n, 2000)
        ref2 = O.collide_batch(o1, o2, P1[:nr], None if P2 is None else P2[:nr], mx, True, nthreads=8)
        got2 = F.collide_batch(m1, P1[:nr], m2, None if P2 is None else P2[:nr], F.CollisionRequest(mx, True), contact_capacity=max(64 * nr, 1024),
                               grow_on_overflow=True)
        ok = ok and np.array_equal(got2.num_contacts, ref2["counts"]) and got2.contacts.tobytes() == ref2["contacts"].tobytes()
        checks += nr
    rounds += 1
    checks += 5 * n
    if not ok:
        print("MISMATCH", tag)
        np.savez("gpurun_out/stress_fail.npz", v1=v1, t1=t1, v2=v2, t2=t2, P1=P1, P2=P2 if P2 is not None else np.zeros(0), S=S, r=r)
        sys.exit(1)
_capi.set_option("traversal", 3)
_capi.set_option("collide_front", 1)
_capi.set_option("scratch_bytes", 2 << 30)
_capi.set_option("host_chunk", 1 << 17)
_capi.set_option("sphere_leaf_trigger", 16)
_capi.set_option("sphere_bound32", 1)
print("stress parity OK: %d random configurations, %d query comparisons; front-traversal distances differing from the oracle in the last bits: %d of %d" % (rounds, checks, ulp_diffs, dist_checks))
