"""One on-device BVH build of a heightfield (argv[1] = grid side; 316 -> 200k triangles); for ncu launch lists."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fcl_b200 as F
from tests.meshes import heightfield
n = int(sys.argv[1]) if len(sys.argv) > 1 else 316
v, t = heightfield(n)
torch.zeros(1).cuda()
for rep in range(3):
    t0 = time.perf_counter()
    m = F.BVHModel.from_arrays(v, t, build_on_device=True)
    m.device_model()
    torch.cuda.synchronize()
    print("build %d tris: %.2f ms" % (len(t), (time.perf_counter() - t0) * 1e3))
