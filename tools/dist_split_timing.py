"""Distance kernel time by query class (colliding: distance 0 / separated), 1M env/rob poses; GPU box only."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fcl_b200 as F

g = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
e, r = np.load(os.path.join(g, "env.npz")), np.load(os.path.join(g, "rob.npz"))
env, rob = F.BVHModel.from_arrays(e["verts"], e["tris"]), F.BVHModel.from_arrays(r["verts"], r["tris"])
n = 1_000_000
P = F.random_poses(n, seed=1)
dP = torch.from_numpy(P).cuda()
dist = torch.empty(n, dtype=torch.float64, device="cuda")
p1 = torch.empty(n, 3, dtype=torch.float64, device="cuda"); p2 = torch.empty_like(p1)
b1 = torch.empty(n, dtype=torch.int32, device="cuda"); b2 = torch.empty_like(b1)
nbv = torch.zeros(n, dtype=torch.int32, device="cuda"); nlf = torch.zeros_like(nbv)
F.distance_batch_device(env, dP, rob, None, F.DistanceRequest(True), dist, p1, p2, b1, b2, nbv, nlf)
torch.cuda.synchronize()
d = dist.cpu().numpy()
classes = {"all": np.ones(n, bool), "colliding (d = 0)": d == 0, "separated": d > 0, "near (0 < d < 200)": (d > 0) & (d < 200), "far (d >= 200)": d >= 200}
hb, hl = nbv.cpu().numpy(), nlf.cpu().numpy()
for name, m in classes.items():
    k = int(m.sum())
    sub = torch.from_numpy(np.ascontiguousarray(P[m])).cuda()
    o = [torch.empty(k, dtype=torch.float64, device="cuda"), torch.empty(k, 3, dtype=torch.float64, device="cuda"), torch.empty(k, 3, dtype=torch.float64, device="cuda"),
         torch.empty(k, dtype=torch.int32, device="cuda"), torch.empty(k, dtype=torch.int32, device="cuda")]
    for _ in range(2):
        F.distance_batch_device(env, sub, rob, None, F.DistanceRequest(True), *o)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(3):
        F.distance_batch_device(env, sub, rob, None, F.DistanceRequest(True), *o)
    ev1.record(); torch.cuda.synchronize()
    print("%-22s %8d poses: %7.3f ms   own box tests %.0f, exact triangle tests %.1f per query" % (name, k, ev0.elapsed_time(ev1) / 3, hb[m].mean(), hl[m].mean()))
