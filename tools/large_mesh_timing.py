"""Kernel timings on BVHs far larger than the caches (BASELINE cfg4 / cfg5 shapes); GPU box only.
    python tools/large_mesh_timing.py [poses]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fcl_b200 as F
from fcl_b200.poses import euler_to_matrix
from tests.meshes import heightfield, noisy_sphere
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
if len(sys.argv) > 2:
    from fcl_b200 import _capi
    _capi.set_option("traversal", int(sys.argv[2]))
    print("traversal", sys.argv[2])
def poses(n, lo, hi, seed):
    rng = np.random.default_rng(seed)
    ang = rng.uniform(0, 2 * np.pi, size=(n, 3))
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    P = np.empty((n, 12)); P[:, :9] = euler_to_matrix(ang[:, 0], ang[:, 1], ang[:, 2]).reshape(n, 9)
    P[:, 9:] = d * rng.uniform(lo, hi, size=(n, 1))
    return torch.from_numpy(P).cuda()
def timed(f, reps=3):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps
cases = []
va, ta = noisy_sphere(1.0, 710, 705, seed=21); vb, tb = noisy_sphere(1.0, 710, 705, seed=22)
cases.append(("cfg5: two 1M-triangle meshes", va, ta, vb, tb, poses(n, 1.6, 2.6, 6)))
sv, st = heightfield(316, size=10.0, seed=1, amp=0.5); lv, lt = noisy_sphere(0.12, 50, 51, seed=10, scale=(2.6, 1.0, 1.0))
Pl = poses(n, 0.0, 4.0, 7); Pl[:, 11] = Pl[:, 11].abs() * 0.3 + 0.2
cases.append(("cfg4-like: 5k-triangle link vs 200k-triangle scene", sv, st, lv, lt, Pl))
for name, v1, t1, v2, t2, P in cases:
    t0 = time.perf_counter()
    A = F.BVHModel.from_arrays(v1, t1, build_on_device=True); B = F.BVHModel.from_arrays(v2, t2, build_on_device=True)
    A.device_model(); B.device_model(); torch.cuda.synchronize()
    tb_ = time.perf_counter() - t0
    cnt = torch.empty(n, dtype=torch.int32, device="cuda"); dist = torch.empty(n, dtype=torch.float64, device="cuda")
    p1 = torch.empty(n, 3, dtype=torch.float64, device="cuda"); p2 = torch.empty_like(p1)
    b1 = torch.empty(n, dtype=torch.int32, device="cuda"); b2 = torch.empty_like(b1)
    nbv = torch.empty(n, dtype=torch.int32, device="cuda"); nlf = torch.empty_like(nbv)
    tc = timed(lambda: F.collide_batch_device(A, None, B, P, F.CollisionRequest(), cnt))
    F.collide_batch_device(A, None, B, P, F.CollisionRequest(), cnt, None, None, nbv, nlf); torch.cuda.synchronize()
    cb, cl = nbv.float().mean().item(), nlf.float().mean().item()
    td = timed(lambda: F.distance_batch_device(A, None, B, P, F.DistanceRequest(True), dist, p1, p2, b1, b2))
    F.distance_batch_device(A, None, B, P, F.DistanceRequest(True), dist, p1, p2, b1, b2, nbv, nlf); torch.cuda.synchronize()
    db, dl = nbv.float().mean().item(), nlf.float().mean().item()
    F.sync_status()
    print("%s  (%d + %d triangles, device build of both %.0f ms, %d poses)" % (name, len(t1), len(t2), tb_ * 1e3, n))
    print("   collide  %.2f ms  %.3g q/s   (%.0f BV tests, %.1f leaf tests per query; colliding %.0f %%)" % (tc, n / tc * 1e3, cb, cl, 100.0 * (cnt > 0).float().mean().item()))
    print("   distance %.2f ms  %.3g q/s   (%.0f BV tests, %.1f leaf tests per query)" % (td, n / td * 1e3, db, dl))
