"""Quick kernel-only timing of one workload across option values (GPU box only; not a benchmark)."""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fcl_b200 as F
from fcl_b200 import _capi
ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="collide")
ap.add_argument("--poses", type=int, default=1000000)
ap.add_argument("--opt", default="leaf_trigger")
ap.add_argument("--values", default="8,16,20,24,28,32")
ap.add_argument("--traversal", type=int, default=2)
a = ap.parse_args()
g = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
e, r = np.load(os.path.join(g, "env.npz")), np.load(os.path.join(g, "rob.npz"))
env, rob = F.BVHModel.from_arrays(e["verts"], e["tris"]), F.BVHModel.from_arrays(r["verts"], r["tris"])
n = a.poses
dP = torch.from_numpy(F.random_poses(n, seed=1)).cuda()
cnt = torch.empty(n, dtype=torch.int32, device="cuda"); dist = torch.empty(n, dtype=torch.float64, device="cuda")
p1 = torch.empty(n, 3, dtype=torch.float64, device="cuda"); p2 = torch.empty(n, 3, dtype=torch.float64, device="cuda")
con = torch.empty(64 * n * 64, dtype=torch.uint8, device="cuda") if a.workload == "contacts" else None
off = torch.empty(n + 1, dtype=torch.int64, device="cuda")
_capi.set_option("traversal", a.traversal)
def run():
    if a.workload == "distance": F.distance_batch_device(env, dP, rob, None, F.DistanceRequest(True), dist, p1, p2)
    elif a.workload == "collide": F.collide_batch_device(env, dP, rob, None, F.CollisionRequest(), cnt)
    else: F.collide_batch_device(env, dP, rob, None, F.CollisionRequest(100, True), cnt, con, off)
for v in a.values.split(","):
    _capi.set_option(a.opt, int(v))
    for _ in range(2): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): run()
    e1.record(); e1.synchronize()
    print("%s %s=%s: %.3f ms" % (a.workload, a.opt, v, e0.elapsed_time(e1) / 5))
