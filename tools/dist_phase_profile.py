"""Per-phase cycles of the sorted-front distance kernel (needs a library built with -DFCLGPU_DIST_PROF=1, selected with
FCLGPU_LIB_PATH).  Development tool.  usage: FCLGPU_LIB_PATH=.../libfclgpu_prof.so python tools/dist_phase_profile.py [poses]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import fcl_b200 as F  # noqa: E402
from fcl_b200 import _capi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
g = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
e, r = np.load(os.path.join(g, "env.npz")), np.load(os.path.join(g, "rob.npz"))
env, rob = F.BVHModel.from_arrays(e["verts"], e["tris"]), F.BVHModel.from_arrays(r["verts"], r["tris"])
dP = torch.from_numpy(F.random_poses(n, seed=1)).cuda()
dist = torch.empty(n, dtype=torch.float64, device="cuda")
p1 = torch.empty(n, 3, dtype=torch.float64, device="cuda")
p2 = torch.empty(n, 3, dtype=torch.float64, device="cuda")
b1 = torch.empty(n, dtype=torch.int32, device="cuda")
b2 = torch.empty(n, dtype=torch.int32, device="cuda")
out = (C.c_uint64 * 16)()
L = _capi.lib()
for rep in range(2):
    L.fclgpu_debug_counters(0, out, 1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    F.distance_batch_device(env, dP, rob, None, F.DistanceRequest(True), dist, p1, p2, b1, b2)
    ev1.record()
    torch.cuda.synchronize()
    L.fclgpu_debug_counters(0, out, 0)
v = np.array(list(out), dtype=np.float64)
tot = v[:5].sum()
print("kernel %.3f ms, %d poses; warp-cycles per query %.0f" % (ev0.elapsed_time(ev1), n, tot / n))
for k, nm in enumerate(("prologue/epilogue", "BV rounds", "screening rounds", "exact rounds", "refill")):
    print("  %-18s %5.1f %% of warp cycles, %6.2f per query, %7.0f cycles each" % (nm, 100 * v[k] / max(tot, 1), v[8 + k] / n, v[k] / max(v[8 + k], 1)))
