"""One end-to-end (host buffers, pinned) timing: python -u tools/e2e_one.py <workload> <host_chunk> [poses]"""
import os, sys, time, faulthandler
faulthandler.dump_traceback_later(50, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fcl_b200 as F
from fcl_b200 import _capi
wl, chunk = sys.argv[1], int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1_000_000
g = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
e, r = np.load(os.path.join(g, "env.npz")), np.load(os.path.join(g, "rob.npz"))
env, rob = F.BVHModel.from_arrays(e["verts"], e["tris"]), F.BVHModel.from_arrays(r["verts"], r["tris"])
P = torch.from_numpy(F.random_poses(n, seed=1)).pin_memory().numpy()
_capi.set_option("host_chunk", chunk)
f = {"collide": lambda: F.collide_batch(env, P, rob, None, F.CollisionRequest(), want_contacts=False, pinned=True),
     "contacts": lambda: F.collide_batch(env, P, rob, None, F.CollisionRequest(100, True), contact_capacity=40 * n, pinned=True),
     "distance": lambda: F.distance_batch(env, P, rob, None, F.DistanceRequest(True), pinned=True)}[wl]
print("start", wl, chunk, n, flush=True)
f(); print("first call done", flush=True); f()
t0 = time.perf_counter()
for _ in range(3):
    f()
dt = (time.perf_counter() - t0) / 3
print("host_chunk %7d %-8s n=%d  %.2f ms  %.3g q/s" % (chunk, wl, n, dt * 1e3, n / dt), flush=True)
