"""Kernel timing of the mesh <-> sphere collide (GPU box only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fcl_b200 as F
from fcl_b200 import _capi
import ctypes as C
g = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
e = np.load(os.path.join(g, "env.npz"))
env = F.BVHModel.from_arrays(e["verts"], e["tris"])
n = 1_000_000
S = torch.from_numpy(F.random_poses(n, seed=3)).cuda()
cnt = torch.empty(n, dtype=torch.int32, device="cuda")
L = _capi.lib()
for radius in (100.0, 350.0, 800.0):
    for req, label in ((F.CollisionRequest(), "verdict"), (F.CollisionRequest(100, True), "contacts<=100")):
        rq = req._c()
        con = torch.empty(64 * 40 * n, dtype=torch.uint8, device="cuda") if req.enable_contact else None
        off = torch.empty(n + 1, dtype=torch.int64, device="cuda") if req.enable_contact else None
        def run():
            rc = L.fclgpu_collide_mesh_sphere_batch(env.device_model(0), radius, n, None, S.data_ptr(), C.byref(rq), cnt.data_ptr(),
                                                    con.data_ptr() if con is not None else None, 40 * n if con is not None else 0,
                                                    off.data_ptr() if off is not None else None, None, None, torch.cuda.current_stream().cuda_stream)
            assert rc == 0, rc
        run(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): run()
        e1.record(); e1.synchronize()
        ms = e0.elapsed_time(e1) / 3
        F.sync_status()
        print("sphere r=%-5g %-14s %.2f ms per 1M queries  %.3g q/s  (colliding %.0f %%, contacts/query %.1f)" % (radius, label, ms, n / ms * 1e3, 100.0 * (cnt > 0).float().mean().item(), cnt.float().mean().item()))

# ---- mesh <-> sphere distance (same poses): kernel time, separated fraction, work per query ----
dist = torch.empty(n, dtype=torch.float64, device="cuda")
p1 = torch.empty(3 * n, dtype=torch.float64, device="cuda")
p2 = torch.empty(3 * n, dtype=torch.float64, device="cuda")
b1 = torch.empty(n, dtype=torch.int32, device="cuda")
nbv = torch.empty(n, dtype=torch.int32, device="cuda")
nlf = torch.empty(n, dtype=torch.int32, device="cuda")
dq = F.DistanceRequest(True)._c()
default_blocks = _capi.get_option("sphere_blocks")
for trig, b32, blocks, radius in [(t, b, k, r) for (t, b, k) in ((16, 1, 3), (16, 1, 4), (16, 1, 5), (16, 0, 3), (0, 0, 3)) for r in (10.0, 100.0, 350.0)]:
    _capi.set_option("sphere_leaf_trigger", trig)
    _capi.set_option("sphere_bound32", b32)
    _capi.set_option("sphere_blocks", blocks)
    def run_d(stats=False):
        rc = L.fclgpu_distance_mesh_sphere_batch(env.device_model(0), radius, n, None, S.data_ptr(), C.byref(dq), dist.data_ptr(),
                                                 p1.data_ptr(), p2.data_ptr(), b1.data_ptr(), None,
                                                 nbv.data_ptr() if stats else None, nlf.data_ptr() if stats else None,
                                                 torch.cuda.current_stream().cuda_stream)
        assert rc == 0, rc
    run_d(True); run_d(); torch.cuda.synchronize()  # both instantiations loaded before the timed launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): run_d()
    e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1) / 3
    F.sync_status()
    print("sphere r=%-5g leaf_trigger=%-2d bound32=%d blocks=%d distance+points %.2f ms per 1M queries  %.3g q/s  (separated %.0f %%, box tests/query %.1f, triangle tests/query %.1f)" % (
        radius, trig, b32, blocks, ms, n / ms * 1e3, 100.0 * (dist > 0).float().mean().item(), nbv.float().mean().item(), nlf.float().mean().item()))
_capi.set_option("sphere_leaf_trigger", 16)
_capi.set_option("sphere_bound32", 1)
_capi.set_option("sphere_blocks", default_blocks)
