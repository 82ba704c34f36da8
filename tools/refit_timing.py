"""Timing of the on-device top-down refit vs the host refit (GPU box only)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fcl_b200 as F
from tests.meshes import heightfield
g = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
e = np.load(os.path.join(g, "env.npz"))
for name, (v, t) in (("env.obj 2180 tris", (e["verts"], e["tris"])), ("heightfield 20k tris", heightfield(100)), ("heightfield 200k tris", heightfield(316))):
    m = F.BVHModel.from_arrays(v, t)
    m.device_model()
    dv = torch.from_numpy(np.ascontiguousarray(v * 1.01)).cuda()
    for _ in range(2):
        m.refit_device(dv)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        m.refit_device(dv)
    e1.record(); e1.synchronize()
    t0 = time.perf_counter()
    F._capi.lib().fclgpu_bvh_refit_topdown(m._bvh, F._capi.addr(np.ascontiguousarray(v * 1.01)), m.num_vertices)
    th = time.perf_counter() - t0
    print("%-24s device refit %.2f ms   host refit %.2f ms" % (name, e0.elapsed_time(e1) / 3, th * 1e3))
    # build: host builder + upload vs on-device build (wall clock, both synchronous)
    t0 = time.perf_counter()
    mh = F.BVHModel.from_arrays(v, t)
    mh.device_model()
    torch.cuda.synchronize()
    tb_host = time.perf_counter() - t0
    F.BVHModel.from_arrays(v, t, build_on_device=True).device_model()  # warm the kernels
    t0 = time.perf_counter()
    md = F.BVHModel.from_arrays(v, t, build_on_device=True)
    md.device_model()
    torch.cuda.synchronize()
    tb_dev = time.perf_counter() - t0
    print("%-24s device build %.2f ms   host build + upload %.2f ms" % (name, tb_dev * 1e3, tb_host * 1e3))
