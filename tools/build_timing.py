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
# the C entry point alone (no Python-side array handling)
import ctypes as C
from fcl_b200 import _capi
L = _capi.lib()
vv = np.ascontiguousarray(v, np.float64); tt = np.ascontiguousarray(t, np.int32)
for rep in range(3):
    h = C.c_void_p()
    t0 = time.perf_counter()
    rc = L.fclgpu_model_build_obbrss(0, vv.ctypes.data, len(vv), tt.ctypes.data, len(tt), 0, C.byref(h))
    dt = time.perf_counter() - t0
    print("fclgpu_model_build_obbrss rc=%d: %.2f ms" % (rc, dt * 1e3))
    L.fclgpu_model_destroy(h)
for rep in range(2):
    h = C.c_void_p()
    t0 = time.perf_counter()
    rc = L.fclgpu_bvh_build_obbrss(vv.ctypes.data, len(vv), tt.ctypes.data, len(tt), 0, C.byref(h))
    t1 = time.perf_counter()
    hm = C.c_void_p()
    L.fclgpu_model_from_bvh(0, h, C.byref(hm))
    t2 = time.perf_counter()
    print("fclgpu_bvh_build_obbrss (host): %.2f ms + upload %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
    L.fclgpu_model_destroy(hm); L.fclgpu_bvh_destroy(h)
