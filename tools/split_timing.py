"""Where does the contact kernel's time go?  Times collide(max 100 contacts) on all poses, on the non-colliding ones and
on the colliding ones (GPU box only)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import fcl_b200 as F  # noqa: E402

g = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
e, r = np.load(os.path.join(g, "env.npz")), np.load(os.path.join(g, "rob.npz"))
env, rob = F.BVHModel.from_arrays(e["verts"], e["tris"]), F.BVHModel.from_arrays(r["verts"], r["tris"])
n = 1_000_000
P = F.random_poses(n, seed=1)
dP = torch.from_numpy(P).cuda()
cnt = torch.zeros(n, dtype=torch.int32, device="cuda")
con = torch.empty(64 * n * 64, dtype=torch.uint8, device="cuda")
off = torch.empty(n + 1, dtype=torch.int64, device="cuda")
req = F.CollisionRequest(100, True)


def timed(dp, reps=5):
    m = dp.shape[0]
    F.collide_batch_device(env, dp, rob, None, req, cnt[:m], con, off[:m + 1])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        F.collide_batch_device(env, dp, rob, None, req, cnt[:m], con, off[:m + 1])
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


t_all = timed(dP)
c = cnt.clone()
hit = c > 0
sat = c >= 100
print("all %d poses: %.3f ms" % (n, t_all))
for name, mask in (("non-colliding", ~hit), ("colliding", hit), ("colliding, < 100 contacts", hit & ~sat), ("saturated (100 contacts)", sat)):
    sub = dP[mask].contiguous()
    print("%-28s %7d poses: %.3f ms  (%.2f us per query)" % (name, sub.shape[0], timed(sub), 1e3 * timed(sub) / max(1, sub.shape[0])))
F.sync_status()
