"""Mean per-query BV / leaf test counts of each traversal variant (GPU box only)."""
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, fcl_b200 as F
from fcl_b200 import _capi
e,r=np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),"tests/golden/env.npz")),np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),"tests/golden/rob.npz"))
env,rob=F.BVHModel.from_arrays(e["verts"],e["tris"]),F.BVHModel.from_arrays(r["verts"],r["tris"])
P=F.random_poses(100000,seed=1)
for t in (0,1,2):
    _capi.set_option("traversal",t)
    d=F.distance_batch(env,P,rob,None,F.DistanceRequest(True),stats=True)
    print("traversal",t,"mean n_bv %.1f mean n_leaf %.1f"%(d.n_bv.mean(), d.n_leaf.mean()))
