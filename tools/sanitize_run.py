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
