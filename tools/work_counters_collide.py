"""Mean per-query BV / leaf test counts of the collide kernel variants (GPU box only)."""
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, fcl_b200 as F
from fcl_b200 import _capi
g = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden")
e, r = np.load(os.path.join(g, "env.npz")), np.load(os.path.join(g, "rob.npz"))
env, rob = F.BVHModel.from_arrays(e["verts"], e["tris"]), F.BVHModel.from_arrays(r["verts"], r["tris"])
P = F.random_poses(100000, seed=1)
for t in (0, 3, 4):
    _capi.set_option("traversal", t)
    d = F.collide_batch(env, P, rob, None, F.CollisionRequest(), want_contacts=False, stats=True)
    c = F.collide_batch(env, P, rob, None, F.CollisionRequest(100, True), contact_capacity=40 * len(P), stats=True)
    print("traversal", t, "binary: n_bv %.1f n_leaf %.2f (max bv %d)   contacts: n_bv %.1f n_leaf %.2f" % (d.n_bv.mean(), d.n_leaf.mean(), d.n_bv.max(), c.n_bv.mean(), c.n_leaf.mean()))
