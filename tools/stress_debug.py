import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fcl_b200 as F
from fcl_b200 import _capi
from oracle import pyoracle as O
d = np.load(sys.argv[1])
split, mx, ec = int(sys.argv[2]), int(sys.argv[3]), bool(int(sys.argv[4]))
v1, t1, v2, t2, P1, S, r = d["v1"], d["t1"], d["v2"], d["t2"], d["P1"], d["S"], float(d["r"])
P2 = d["P2"] if d["P2"].size else None
m1, m2 = F.BVHModel.from_arrays(v1, t1, split), F.BVHModel.from_arrays(v2, t2, split)
o1, o2 = O.Model(v1, t1, split), O.Model(v2, t2, split)
n = len(P1)
for trav in (0, 1, 2, 3, 4):
    _capi.set_option("traversal", trav)
    for front in (1, 2):
        _capi.set_option("collide_front", front)
        ref = O.collide_batch(o1, o2, P1, P2, mx, ec, nthreads=8)
        got = F.collide_batch(m1, P1, m2, P2, F.CollisionRequest(mx, ec), contact_capacity=max(64 * n, 1024), grow_on_overflow=True)
        a = np.array_equal(got.num_contacts, ref["counts"])
        b = a and np.array_equal(got.contacts["b1"], ref["contacts"]["b1"]) and np.array_equal(got.contacts["b2"], ref["contacts"]["b2"])
        c = b and (not ec or got.contacts.tobytes() == ref["contacts"].tobytes())
        cnt = F.collide_batch(m1, P1, m2, P2, F.CollisionRequest(mx, False), want_contacts=False)
        refc = O.collide_batch(o1, o2, P1, P2, mx, False, nthreads=8)["counts"]
        e = np.array_equal(cnt.num_contacts, refc)
        rd = O.distance_batch(o1, o2, P1, P2, True, 2, nthreads=8)
        gd = F.distance_batch(m1, P1, m2, P2, F.DistanceRequest(True))
        f = np.array_equal(gd.min_distance, rd["min_distance"])
        print("trav", trav, "front", front, "counts", a, "ids", b, "bytes", c, "counts-only", e, "distance", f)
        if not e:
            bad = np.nonzero(cnt.num_contacts != refc)[0]
            print("  counts-only mismatches:", len(bad), "first", bad[:5], cnt.num_contacts[bad[:5]], refc[bad[:5]])
        if not a:
            bad = np.nonzero(got.num_contacts != ref["counts"])[0]
            print("  count mismatches:", len(bad), bad[:5], got.num_contacts[bad[:5]], ref["counts"][bad[:5]])
        if not f:
            bad = np.nonzero(gd.min_distance != rd["min_distance"])[0]
            print("  distance mismatches:", len(bad), bad[:5], gd.min_distance[bad[:5]], rd["min_distance"][bad[:5]], "rel diff", (gd.min_distance[bad[:5]] - rd["min_distance"][bad[:5]]) / rd["min_distance"][bad[:5]], "ids gpu", gd.b1[bad[:5]], gd.b2[bad[:5]], "ids oracle", rd["b1"][bad[:5]], rd["b2"][bad[:5]])
rs = O.collide_mesh_sphere_batch(o1, r, P1, S, mx, True, nthreads=8)
gs = F.collide_mesh_sphere_batch(m1, P1, F.Sphere(r), S, F.CollisionRequest(mx, True), contact_capacity=max(64 * n, 1024), grow_on_overflow=True)
print("sphere counts", np.array_equal(gs.num_contacts, rs["counts"]), "bytes", gs.contacts.tobytes() == rs["contacts"].tobytes())
if gs.contacts.tobytes() != rs["contacts"].tobytes() and np.array_equal(gs.num_contacts, rs["counts"]):
    for k in ("b1", "b2", "normal", "pos", "penetration_depth"):
        kk = "depth" if k == "penetration_depth" else k
        x, y = gs.contacts[k], rs["contacts"][kk]
        bad = np.nonzero((x != y).reshape(len(x), -1).any(axis=1))[0]
        print("  field", k, "mismatching contacts", len(bad), bad[:3], x[bad[:3]], y[bad[:3]])
