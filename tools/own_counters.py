"""Own work counters of the ordered-front contact kernel vs the reference traversal's (GPU box only)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fcl_b200 as F
from fcl_b200 import _capi
g = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
e, r = np.load(os.path.join(g, "env.npz")), np.load(os.path.join(g, "rob.npz"))
env, rob = F.BVHModel.from_arrays(e["verts"], e["tris"]), F.BVHModel.from_arrays(r["verts"], r["tris"])
n = 200000
dP = torch.from_numpy(F.random_poses(n, seed=1)).cuda()
cnt = torch.zeros(n, dtype=torch.int32, device="cuda")
con = torch.empty(64 * n * 64, dtype=torch.uint8, device="cuda")
off = torch.empty(n + 1, dtype=torch.int64, device="cuda")
req = F.CollisionRequest(100, True)
out = {}
for trav in (3, 0):
    _capi.set_option("traversal", trav)
    nbv = torch.zeros(n, dtype=torch.int32, device="cuda"); nlf = torch.zeros(n, dtype=torch.int32, device="cuda")
    F.collide_batch_device(env, dP, rob, None, req, cnt, con, off, nbv, nlf)
    torch.cuda.synchronize()
    out[trav] = (nbv.cpu().numpy().astype(np.int64), nlf.cpu().numpy().astype(np.int64), cnt.cpu().numpy())
_capi.set_option("traversal", 3)
c = out[0][2]
for name, m in (("all", c >= 0), ("non-colliding", c == 0), ("colliding < 100", (c > 0) & (c < 100)), ("saturated", c >= 100)):
    print("%-16s n=%6d  ordered kernel: n_bv %.1f n_leaf %.1f | reference traversal: n_bv %.1f n_leaf %.1f" % (
        name, m.sum(), out[3][0][m].mean(), out[3][1][m].mean(), out[0][0][m].mean(), out[0][1][m].mean()))
