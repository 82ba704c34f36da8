"""End-to-end (host buffers) timing of the host API vs pipeline chunk size, plus raw pinned copy bandwidth (GPU box only)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fcl_b200 as F
from fcl_b200 import _capi
g = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
e, r = np.load(os.path.join(g, "env.npz")), np.load(os.path.join(g, "rob.npz"))
env, rob = F.BVHModel.from_arrays(e["verts"], e["tris"]), F.BVHModel.from_arrays(r["verts"], r["tris"])
n = 1_000_000
Ph = torch.from_numpy(F.random_poses(n, seed=1)).pin_memory()
d = torch.empty(n, 12, dtype=torch.float64, device="cuda")
for nbytes in (96 * n, 96 * n // 8):
    k = nbytes // 96
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        d[:k].copy_(Ph[:k], non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print("pinned H2D %6.1f MB: %.2f ms  %.1f GB/s" % (nbytes / 1e6, dt * 1e3, nbytes / dt / 1e9))
out = torch.empty(n, dtype=torch.int32).pin_memory()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    out.copy_(torch.empty(n, dtype=torch.int32, device="cuda"), non_blocking=True)
torch.cuda.synchronize(); print("pinned D2H 4 MB: %.2f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
P = Ph.numpy()
for chunk in (1 << 14, 1 << 15, 1 << 16, 1 << 17, 1 << 18, 1 << 19):
    _capi.set_option("host_chunk", chunk)
    for wl in ("collide", "contacts", "distance"):
        f = {"collide": lambda: F.collide_batch(env, P, rob, None, F.CollisionRequest(), want_contacts=False, pinned=True),
             "contacts": lambda: F.collide_batch(env, P, rob, None, F.CollisionRequest(100, True), contact_capacity=40 * n, pinned=True),
             "distance": lambda: F.distance_batch(env, P, rob, None, F.DistanceRequest(True), pinned=True)}[wl]
        f(); f()
        t0 = time.perf_counter()
        for _ in range(3):
            f()
        dt = (time.perf_counter() - t0) / 3
        print("host_chunk %7d %-8s %.2f ms  %.3g q/s" % (chunk, wl, dt * 1e3, n / dt))
