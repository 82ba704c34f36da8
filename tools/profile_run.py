"""One or a few launches of a traversal kernel, for ncu (never a benchmark).
    ncu ... python tools/profile_run.py --workload distance --poses 100000 --traversal 0 --launches 3"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import fcl_b200 as F  # noqa: E402
from fcl_b200 import _capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="distance")
ap.add_argument("--poses", type=int, default=100000)
ap.add_argument("--traversal", type=int, default=0)
ap.add_argument("--launches", type=int, default=3)
ap.add_argument("--opt", action="append", default=[], help="name=value fclgpu options")
a = ap.parse_args()
_capi.set_option("traversal", a.traversal)
for kv in a.opt:
    k, v = kv.split("=")
    _capi.set_option(k, int(v))
g = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
e, r = np.load(os.path.join(g, "env.npz")), np.load(os.path.join(g, "rob.npz"))
env, rob = F.BVHModel.from_arrays(e["verts"], e["tris"]), F.BVHModel.from_arrays(r["verts"], r["tris"])
n = a.poses
dP = torch.from_numpy(F.random_poses(n, seed=1)).cuda()
cnt = torch.empty(n, dtype=torch.int32, device="cuda")
dist = torch.empty(n, dtype=torch.float64, device="cuda")
p1 = torch.empty(n, 3, dtype=torch.float64, device="cuda")
p2 = torch.empty(n, 3, dtype=torch.float64, device="cuda")
b1 = torch.empty(n, dtype=torch.int32, device="cuda")
b2 = torch.empty(n, dtype=torch.int32, device="cuda")
con = torch.empty(64 * n * 64, dtype=torch.uint8, device="cuda") if a.workload == "contacts" else None
off = torch.empty(n + 1, dtype=torch.int64, device="cuda")
for _ in range(a.launches):
    if a.workload == "distance":
        F.distance_batch_device(env, dP, rob, None, F.DistanceRequest(True), dist, p1, p2, b1, b2)
    elif a.workload == "sphere_distance":  # sphere centres = the pose translations, radius 100
        import ctypes as C
        rq = F.DistanceRequest(True)._c()
        rc = _capi.lib().fclgpu_distance_mesh_sphere_batch(env.device_model(0), 100.0, n, None, dP.data_ptr(), C.byref(rq), dist.data_ptr(),
                                                           p1.data_ptr(), p2.data_ptr(), b1.data_ptr(), b2.data_ptr(), None, None,
                                                           torch.cuda.current_stream().cuda_stream)
        assert rc == 0, rc
    elif a.workload == "collide":
        F.collide_batch_device(env, dP, rob, None, F.CollisionRequest(), cnt)
    else:
        F.collide_batch_device(env, dP, rob, None, F.CollisionRequest(100, True), cnt, con, off)
F.sync_status()
print("done", a.workload, n, float(dist.sum()) if a.workload in ("distance", "sphere_distance") else int(cnt.sum()))
