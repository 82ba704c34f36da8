"""One launch of each cfg5 / cfg4 kernel for ncu (never a benchmark).
    ncu ... python tools/profile_big.py --workload cfg5 --poses 100000"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import fcl_b200 as F  # noqa: E402
from fcl_b200 import _capi  # noqa: E402
from fcl_b200 import workloads as W  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg5")
ap.add_argument("--poses", type=int, default=100000)
ap.add_argument("--opt", action="append", default=[])
a = ap.parse_args()
for kv in a.opt:
    k, v = kv.split("=")
    _capi.set_option(k, int(v))
n = a.poses
cnt = torch.zeros(n, dtype=torch.int32, device="cuda")
nbv = torch.zeros(n, dtype=torch.int32, device="cuda")
nlf = torch.zeros(n, dtype=torch.int32, device="cuda")
if a.workload == "cfg5":
    (va, ta), (vb, tb) = W.cfg5_meshes()
    A = F.BVHModel.from_arrays(va, ta, build_on_device=True)
    B = F.BVHModel.from_arrays(vb, tb, build_on_device=True)
    dP = torch.from_numpy(W.shell_poses(n, 1.5, 3.0, seed=6)).cuda()
    dist = torch.zeros(n, dtype=torch.float64, device="cuda")
    p1 = torch.zeros(n, 3, dtype=torch.float64, device="cuda")
    p2 = torch.zeros(n, 3, dtype=torch.float64, device="cuda")
    b1 = torch.zeros(n, dtype=torch.int32, device="cuda")
    b2 = torch.zeros(n, dtype=torch.int32, device="cuda")
    within = torch.zeros(n, dtype=torch.uint8, device="cuda")
    F.collide_batch_device(A, None, B, dP, F.CollisionRequest(), cnt)
    F.distance_batch_device(A, None, B, dP, F.DistanceRequest(True), dist, p1, p2, b1, b2)
    rc = _capi.lib().fclgpu_within_tolerance_batch(A.device_model(0), B.device_model(0), n, None, dP.data_ptr(), 0.05, within.data_ptr(),
                                                   dist.data_ptr(), None, None, torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    # the kernels' OWN work counters (not profiled: separate stats launches)
    F.collide_batch_device(A, None, B, dP, F.CollisionRequest(), cnt, None, None, nbv, nlf)
    torch.cuda.synchronize()
    print("collide own counters: n_bv %.1f n_leaf %.2f" % (nbv.float().mean().item(), nlf.float().mean().item()))
    F.distance_batch_device(A, None, B, dP, F.DistanceRequest(True), dist, p1, p2, b1, b2, nbv, nlf)
    torch.cuda.synchronize()
    print("distance own counters: n_bv %.1f n_leaf %.2f" % (nbv.float().mean().item(), nlf.float().mean().item()))
else:
    (sv, st), links = W.cfg4_meshes()
    scene = F.BVHModel.from_arrays(sv, st, build_on_device=True)
    link = F.BVHModel.from_arrays(*links[6], build_on_device=True)
    LP = W.arm_configurations(n, seed=4)
    dP = torch.from_numpy(np.ascontiguousarray(LP[:, 6])).cuda()
    F.collide_batch_device(scene, None, link, dP, F.CollisionRequest(), cnt)
    torch.cuda.synchronize()
    F.collide_batch_device(scene, None, link, dP, F.CollisionRequest(), cnt, None, None, nbv, nlf)
    torch.cuda.synchronize()
    print("cfg4 collide own counters: n_bv %.1f n_leaf %.2f" % (nbv.float().mean().item(), nlf.float().mean().item()))
F.sync_status()
print("done", a.workload, n, int((cnt > 0).sum()))
