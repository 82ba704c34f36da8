"""Bench line for the mesh <-> sphere distance in bench.py's format (GPU box only):
env.obj posed at identity vs Sphere(radius) at 1M random centres.  value = kernel throughput with the inputs resident
in HBM, e2e = through fclgpu_distance_mesh_sphere_batch_host with pinned host buffers, roofline from the REFERENCE
traversal's counters (the oracle's n_bv / n_leaf on a sample, SURVEY 8d: 2 poses + 120 B per node test + 72 B per
triangle test + 64 B out), cpu_baseline = the oracle on the host cores.
    python tools/bench_mesh_sphere.py [--radius 100] [--poses 1000000] [--steps 20]"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import ctypes as C
import fcl_b200 as F
from fcl_b200 import _capi
from oracle import pyoracle as O
import bench as B

ap = argparse.ArgumentParser()
ap.add_argument("--radius", type=float, default=100.0)
ap.add_argument("--poses", type=int, default=1_000_000)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--cpu-sample", type=int, default=20000)
a = ap.parse_args()
(ev, et), _ = B.load_meshes()
env = F.BVHModel.from_arrays(ev, et)
n = a.poses
S = F.random_poses(n, seed=1)
hS = torch.from_numpy(S).pin_memory()
dS = hS.cuda()
dist = torch.empty(n, dtype=torch.float64, device="cuda")
p1 = torch.empty(n, 3, dtype=torch.float64, device="cuda")
p2 = torch.empty(n, 3, dtype=torch.float64, device="cuda")
b1 = torch.empty(n, dtype=torch.int32, device="cuda")
rq = F.DistanceRequest(True)._c()
L = _capi.lib()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def kernel():
    rc = L.fclgpu_distance_mesh_sphere_batch(env.device_model(0), a.radius, n, None, dS.data_ptr(), C.byref(rq), dist.data_ptr(),
                                             p1.data_ptr(), p2.data_ptr(), b1.data_ptr(), None, None, None,
                                             torch.cuda.current_stream().cuda_stream)
    assert rc == 0, rc


for _ in range(a.warmup):
    kernel()
torch.cuda.synchronize()
sampler = B.ClockSampler(0)
sampler.start()
launches0 = _capi.launch_count()
ms = 0.0
for _ in range(a.steps):
    flush.fill_(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); kernel(); e1.record(); e1.synchronize()
    ms += e0.elapsed_time(e1)
launches = _capi.launch_count() - launches0
clocks = sampler.stop()
F.sync_status(0)
k_ms = ms / a.steps
sphere = F.Sphere(a.radius)
hs = hS.numpy()
for _ in range(2):
    F.distance_mesh_sphere_batch(env, None, sphere, hs, F.DistanceRequest(True), pinned=True)
t0 = time.perf_counter()
for _ in range(a.steps):
    r = F.distance_mesh_sphere_batch(env, None, sphere, hs, F.DistanceRequest(True), pinned=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
# reference-traversal counters and CPU baseline on a sample
s = min(a.cpu_sample, n)
oenv = O.Model(ev, et)
threads = O.hardware_threads()
ident = F.identity_poses(s)
ref = O.distance_mesh_sphere_batch(oenv, a.radius, ident, S[:s], nthreads=threads)
brute = O.distance_mesh_sphere_batch(oenv, a.radius, ident, S[:s], brute=True, nthreads=threads)
got = dist.cpu().numpy()[:s]
bytes_per_query = 2 * 96 + ref["n_bv"].astype(np.float64) * 120 + ref["n_leaf"].astype(np.float64) * 72 + 64
alg = float(bytes_per_query.mean()) * n
peak, which = B.measured_peaks()
line = {
    "metric": "mesh-sphere distance queries/sec (SURVEY 8f rank 2)", "value": n / (k_ms * 1e-3), "unit": "queries/s", "n_gpus": 1,
    "steps": a.steps, "warmup": a.warmup, "ms_per_step": k_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
    "dtype": "f64", "data": "synthetic",
    "config": {"workload": "env.obj (2180 tris) at identity vs Sphere(r=%g) at %d random centres (seed-1 pose translations), distance() with nearest points" % (a.radius, n),
               "l2_flush_between_steps": True, "sphere_leaf_trigger": _capi.get_option("sphere_leaf_trigger"), "sphere_bound32": _capi.get_option("sphere_bound32")},
    "clocks": clocks,
    "e2e": {"value": n * a.steps / dt, "unit": "queries/s", "h2d_bytes_per_step": 96 * n, "d2h_bytes_per_step": n * (8 + 24 + 24 + 4 + 4)},
    "gpu_launches": launches,
    "roofline": {"bound": "hbm", "achieved": alg / (k_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / (k_ms * 1e-3) / 1e9 / peak,
                 "traffic": None, "peak_source": which, "kernel_ms": k_ms, "algorithmic_bytes_per_launch": alg,
                 "mean_n_bv_reference": float(ref["n_bv"].mean()), "mean_n_leaf_reference": float(ref["n_leaf"].mean()),
                 "note": "counters of the reference's traversal from the oracle on a %d-query sample, scaled to the batch; records are L1/L2 resident" % s},
    "cpu_baseline": {"value": s / ref["seconds"], "unit": "queries/s", "cores": threads, "kind": "port",
                     "sample": "first %d queries, %d host threads, one pass" % (s, threads),
                     "matches_gpu_all_triangles_minimum": bool(np.array_equal(got, brute["min_distance"])),
                     "matches_gpu_traversal_1e-12": bool(np.all(np.abs(got - ref["min_distance"]) <= 1e-12 * np.abs(ref["min_distance"])))},
}
print(json.dumps(line))
