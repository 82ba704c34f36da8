"""Small pass over every kernel family for compute-sanitizer (memcheck / racecheck); GPU box only.
    compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fcl_b200 as F
from fcl_b200 import _capi
from tests.meshes import heightfield
g = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
e, r = np.load(os.path.join(g, "env.npz")), np.load(os.path.join(g, "rob.npz"))
env, rob = F.BVHModel.from_arrays(e["verts"], e["tris"]), F.BVHModel.from_arrays(r["verts"], r["tris"])
n = 3000
P = torch.from_numpy(F.random_poses(n, seed=5)).pin_memory().numpy()
_capi.set_option("host_chunk", 1024)
sig = []
for trav in (0, 1, 2, 3, 4):
    _capi.set_option("traversal", trav)
    c = F.collide_batch(env, P, rob, None, F.CollisionRequest(50, True), contact_capacity=50 * n, pinned=True)
    nc = int(c.num_contacts.sum())  # pinned results are valid until the next call of the same shape
    b = F.collide_batch(env, P, rob, None, F.CollisionRequest(), want_contacts=False, pinned=True)
    d = F.distance_batch(env, P, rob, None, F.DistanceRequest(True), pinned=True)
    sig.append((trav, nc, int(b.num_contacts.sum()), float(d.min_distance.sum())))
    print(*sig[-1])
_capi.set_option("traversal", 3)
_capi.set_option("collide_front", 2)
f = F.collide_batch(env, P, rob, None, F.CollisionRequest(100000, False), want_contacts=False)
print("front", int(f.num_contacts.sum()))
_capi.set_option("collide_front", 1)
ms = F.collide_mesh_sphere_batch(env, None, F.Sphere(350.0), P, F.CollisionRequest(50, True), contact_capacity=50 * n, stats=True)
print("mesh-sphere", int(ms.num_contacts.sum()))
# on-device build and refit (block-, warp- and thread-cooperative paths), distance overflow area
v, t = heightfield(40, size=10.0, seed=3, amp=0.6)
for variant in (2, 1, 0):
    _capi.set_option("refit_warp", variant)
    m = F.BVHModel.from_arrays(v, t, build_on_device=True)
    m.device_model()
    m.refit_device(torch.from_numpy(v * 1.01).cuda())
    F.sync_status()
_capi.set_option("refit_warp", 2)
big = F.BVHModel.from_arrays(*heightfield(190, size=10.0, seed=4, amp=0.5), build_on_device=True)  # 72k triangles: >= 2^17 nodes
Q = F.random_poses(300, seed=6, extents=(-4, 4, -4, 4, 0.2, 1.5)) if "extents" in F.random_poses.__code__.co_varnames else F.random_poses(300, seed=6)
dd = F.distance_batch(big, None, rob, Q, F.DistanceRequest(True))
cc = F.collide_batch(big, None, rob, Q, F.CollisionRequest(), want_contacts=False)
print("big", float(np.nansum(dd.min_distance)), int(cc.num_contacts.sum()))
# round 2: median split on the device, bottom-up refit, compact contact records, continuous collision, plane / halfspace,
# tolerance verdicts, broadphase
m = F.BVHModel.from_arrays(v, t, F.SPLIT_METHOD_MEDIAN, build_on_device=True)
m.device_model()
assert m.beginReplaceModel() == 0 and m.replaceSubModel(v * 1.01) == 0 and m.endReplaceModel() == 0  # bottom-up
for fmt in (F.CONTACT_IDS, F.CONTACT_F32):
    cf = F.collide_batch(env, P, rob, None, F.CollisionRequest(50, True), contact_capacity=50 * n, contact_format=fmt)
    print("compact", fmt, int(cf.num_contacts.sum()))
P1 = P.copy()
P1[:, 9:] += 200.0
cc2 = F.continuous_collide_batch(env, None, None, rob, P, P1, F.ContinuousCollisionRequest(ccd_solver_type=F.CCDC_CONSERVATIVE_ADVANCEMENT))
print("continuous", int(cc2.is_collide.sum()), float(cc2.time_of_contact.sum()))
hp = F.collide_mesh_plane_batch(env, None, F.Halfspace([0, 0, 1.0], 100.0), P, F.CollisionRequest(50, True), contact_capacity=50 * n)
pp = F.collide_mesh_plane_batch(env, None, F.Plane([0, 0, 1.0], 100.0), P, F.CollisionRequest(50, True), contact_capacity=50 * n)
print("halfspace / plane", int(hp.num_contacts.sum()), int(pp.num_contacts.sum()))
w, _ = F.within_tolerance_batch(env, P, rob, None, 50.0)
print("within tolerance", int(w.sum()))
mgr1, mgr2 = F.NaiveCollisionManager(), F.NaiveCollisionManager()
for i in range(40):
    mgr1.registerObject(F.CollisionObject(env if i % 2 else rob, F.Transform3.from_pose12(P[i])))
    mgr2.registerObject(F.CollisionObject(rob, F.Transform3.from_pose12(P[100 + i])))
bp = mgr1.collide_batch(mgr2)
print("broadphase pairs", len(bp.pairs))
