#!/usr/bin/env python
"""bench.py — batched OBBRSS mesh-mesh queries/second on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--workload distance|collide|contacts|sphere_distance]
    python bench.py --impl reference ...     # the CPU oracle (the only CPU FCL buildable here)

A step = one pass of the hot path over one batch of synthetic poses (env.obj vs rob.obj).
Default workload = BASELINE configs[1]: distance() with nearest points, 1M poses per GPU.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "mesh-mesh collide/distance queries/sec at 1/2/4/8 B200 vs host-core FCL"
UNIT = "queries/s"
WORKLOADS = {
    # name: (BASELINE config it corresponds to, description)
    "distance": "cfg2: env.obj vs rob.obj distance() with nearest points, 1M random poses per GPU, double",
    "collide": "cfg1-style: env.obj vs rob.obj collide() binary verdict (CollisionRequest()), 1M random poses per GPU",
    "contacts": "cfg3: env.obj vs rob.obj collide() enable_contact, num_max_contacts=100, 1M random poses per GPU",
    # SURVEY 8f rank 2 (the row next to the path), single GPU: see run_sphere_distance()
    "sphere_distance": "env.obj at identity vs Sphere(r=100) at the seed-1 pose translations, distance() with nearest points, 1M queries",
}
SPHERE_RADIUS = 100.0


def load_meshes():
    g = os.path.join(ROOT, "tests", "golden")
    e, r = np.load(os.path.join(g, "env.npz")), np.load(os.path.join(g, "rob.npz"))
    return (e["verts"], e["tris"]), (r["verts"], r["tris"])


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 7:
                    self.rows.append(f)
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm, mx, reasons = [], [], set()
        for f in self.rows:
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_bytes(workload, n_bv, n_leaf, n_contacts=None):
    """bytes per query from the reference traversal's counters.  Node records are
    128 B here (120 B of fields + the precomputed size), triangles 72 B."""
    n_bv = n_bv.astype(np.float64)
    n_leaf = n_leaf.astype(np.float64)
    if workload == "distance":
        return 96 + n_bv * 2 * 128 + n_leaf * 2 * 72 + 64
    out = 1.0 if workload == "collide" else 4.0 + 64.0 * n_contacts.astype(np.float64)
    return 96 + n_bv * 2 * 120 + n_leaf * 2 * 72 + out


def run_reference(args, rank, world):
    """--impl reference: the CPU oracle (restatement of the reference; real FCL needs Eigen/libccd,
    absent from this image) on all host threads, bounded sample per step."""
    if rank != 0:
        return
    from fcl_b200.poses import random_poses
    from oracle import pyoracle as O

    O.build()
    (ev, et), (rv, rt) = load_meshes()
    env, rob = O.Model(ev, et), O.Model(rv, rt)
    threads = O.hardware_threads()
    sample = args.cpu_sample
    P = random_poses(sample, seed=1)

    ident = np.zeros((sample, 12))
    ident[:, 0] = ident[:, 4] = ident[:, 8] = 1.0

    def step():
        if args.workload == "sphere_distance":
            return O.distance_mesh_sphere_batch(env, SPHERE_RADIUS, ident, P, nthreads=threads)["seconds"]
        if args.workload == "distance":
            return O.distance_batch(env, rob, P, None, True, 2, nthreads=threads)["seconds"]
        if args.workload == "collide":
            return O.collide_batch(env, rob, P, None, 1, False, nthreads=threads)["seconds"]
        return O.collide_batch(env, rob, P, None, 100, True, nthreads=threads)["seconds"]

    for _ in range(args.warmup):
        step()
    secs = [step() for _ in range(args.steps)]
    total = float(sum(secs))
    value = sample * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload], "sample_poses_per_step": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"first {sample} poses of the seed-1 batch per step, {threads} host threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


_RESULT_FD = None


def quiet_stdout():
    """The contract is ONE JSON line on stdout.  Libraries (NCCL prints its version banner there when NCCL_DEBUG is
    set on the box) write to file descriptor 1 directly, so point fd 1 at stderr for the whole run and keep the
    real stdout for the result line."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def ensure_built(local_rank):
    """The CUDA library normally travels with the tree (built by __graft_entry__.build()); on a bare checkout build it
    once (local rank 0) instead of failing -- there is still no CPU fallback, only a compile step."""
    lib = os.path.join(ROOT, "fcl_b200", "lib", "libfclgpu.so")
    if os.path.exists(lib):
        return
    if local_rank == 0:
        tmp_out = "../lib/libfclgpu.so.building"
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "fcl_b200", "csrc"), "-s", "OUT=" + tmp_out], stdout=sys.stderr)
        os.replace(os.path.join(ROOT, "fcl_b200", "lib", "libfclgpu.so.building"), lib)
    else:
        for _ in range(1800):
            if os.path.exists(lib):
                return
            time.sleep(1.0)
        raise SystemExit("libfclgpu.so was not built")


def run_sphere_distance(args, local):
    """--workload sphere_distance (single GPU): the row next to the path, SURVEY 8f rank 2.  value = kernel throughput with
    the inputs resident in HBM; e2e = through fclgpu_distance_mesh_sphere_batch_host with pinned host buffers; roofline from
    the REFERENCE traversal's counters (the oracle's n_bv / n_leaf on the CPU sample, SURVEY 8d accounting: 2 poses + 120 B
    per node test + 72 B per triangle test + 64 B out); cpu_baseline = the oracle on the host cores."""
    import ctypes as C

    import torch

    import fcl_b200 as F
    from fcl_b200 import _capi

    (ev, et), _ = load_meshes()
    env = F.BVHModel.from_arrays(ev, et)
    n = args.poses
    S = F.random_poses(n, seed=1)
    hS = torch.from_numpy(S).pin_memory()
    dS = hS.cuda()
    dist_d = torch.empty(n, dtype=torch.float64, device="cuda")
    p1 = torch.empty(n, 3, dtype=torch.float64, device="cuda")
    p2 = torch.empty(n, 3, dtype=torch.float64, device="cuda")
    b1 = torch.empty(n, dtype=torch.int32, device="cuda")
    rq = F.DistanceRequest(True)._c()
    L = _capi.lib()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def kernel():
        rc = L.fclgpu_distance_mesh_sphere_batch(env.device_model(local), SPHERE_RADIUS, n, None, dS.data_ptr(), C.byref(rq),
                                                 dist_d.data_ptr(), p1.data_ptr(), p2.data_ptr(), b1.data_ptr(), None, None, None,
                                                 torch.cuda.current_stream().cuda_stream)
        assert rc == 0, rc

    for _ in range(args.warmup):
        kernel()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = _capi.launch_count()
    ms = 0.0
    for _ in range(args.steps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        kernel()
        e1.record()
        e1.synchronize()
        ms += e0.elapsed_time(e1)
    launches = _capi.launch_count() - launches0
    clocks = sampler.stop()
    F.sync_status(local)
    k_ms = ms / args.steps
    sphere = F.Sphere(SPHERE_RADIUS)
    hs = hS.numpy()
    e2e = None
    if not args.no_e2e:
        for _ in range(2):
            F.distance_mesh_sphere_batch(env, None, sphere, hs, F.DistanceRequest(True), pinned=True)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            F.distance_mesh_sphere_batch(env, None, sphere, hs, F.DistanceRequest(True), pinned=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e2e = {"value": n * args.steps / dt, "unit": UNIT, "h2d_bytes_per_step": 96 * n, "d2h_bytes_per_step": n * (8 + 24 + 24 + 4 + 4)}
    from oracle import pyoracle as O  # checker and CPU baseline only

    O.build()
    s = min(args.cpu_sample, n)
    oenv = O.Model(ev, et)
    threads = O.hardware_threads()
    ident = F.identity_poses(s)
    ref = O.distance_mesh_sphere_batch(oenv, SPHERE_RADIUS, ident, S[:s], nthreads=threads)
    brute = O.distance_mesh_sphere_batch(oenv, SPHERE_RADIUS, ident, S[:s], brute=True, nthreads=threads)
    got = dist_d.cpu().numpy()[:s]
    per_query = 2 * 96 + ref["n_bv"].astype(np.float64) * 120 + ref["n_leaf"].astype(np.float64) * 72 + 64
    alg = float(per_query.mean()) * n
    peak, which = measured_peaks()
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            t = json.load(f).get("sphere_distance")
        if t and t["poses"] == n:
            traffic = t["dram_bytes_per_launch"] * t["launches_per_step"]
            traffic_src = "profiles/r01_traffic.json (%s, %d launch(es) per step)" % (t["kernel"], t["launches_per_step"])
    except Exception:  # pragma: no cover
        pass
    achieved = alg / (k_ms * 1e-3) / 1e9
    emit({
        "metric": METRIC, "value": n / (k_ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": k_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS["sphere_distance"], "poses_per_gpu": n, "pose_seed": 1, "l2_flush_between_steps": True,
                   "sphere_leaf_trigger": _capi.get_option("sphere_leaf_trigger"), "sphere_bound32": _capi.get_option("sphere_bound32"),
                   "sphere_blocks": _capi.get_option("sphere_blocks"),
                   "multi_gpu": "single GPU"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "traffic_source": traffic_src, "peak_source": which, "kernel_ms": k_ms, "algorithmic_bytes_per_launch": alg,
                     "mean_n_bv": float(ref["n_bv"].mean()), "mean_n_leaf": float(ref["n_leaf"].mean()),
                     "note": "counters of the reference's traversal from the oracle on the CPU sample, scaled to the batch; records are "
                             "L1/L2 resident, the binding resource is the L1 data pipe (profiles/r01_ncu_sphere_distance.txt)"},
        "cpu_baseline": {"value": s / ref["seconds"], "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"first {s} queries of the batch, {threads} host threads, one pass",
                         "matches_gpu": bool(np.array_equal(got, brute["min_distance"])),
                         "matches_gpu_traversal_1e-12": bool(np.all(np.abs(got - ref["min_distance"]) <= 1e-12 * np.abs(ref["min_distance"])))},
    })


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="distance", choices=sorted(WORKLOADS))
    ap.add_argument("--poses", type=int, default=1_000_000, help="poses per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=20000)
    ap.add_argument("--traversal", type=int, default=3, help="kernel variant (fclgpu option 'traversal', see DESIGN.md)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="name=value library option (A/B runs; recorded in config)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    ensure_built(local)
    import fcl_b200 as F
    from fcl_b200 import _capi
    from fcl_b200.sharding import all_gather_records

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _capi.set_option("traversal", args.traversal)
    for kv in args.opt:
        k, v = kv.split("=")
        _capi.set_option(k, int(v))

    if args.workload == "sphere_distance":
        if world > 1:
            raise SystemExit("--workload sphere_distance is a single-GPU line (run it without torchrun)")
        run_sphere_distance(args, local)
        return

    (ev, et), (rv, rt) = load_meshes()
    env, rob = F.BVHModel.from_arrays(ev, et), F.BVHModel.from_arrays(rv, rt)
    env.device_model(local)
    rob.device_model(local)  # BVHs replicated on every GPU

    n = args.poses
    P = F.random_poses(n, seed=1, start=rank * n)  # rank r owns poses [r*n, (r+1)*n) of the global batch
    hP = torch.from_numpy(P).pin_memory()
    dP = hP.to(dev)
    wl = args.workload
    creq = F.CollisionRequest() if wl == "collide" else F.CollisionRequest(100, True)
    dreq = F.DistanceRequest(True)

    # resident outputs
    if wl == "distance":
        o_dist = torch.empty(n, dtype=torch.float64, device=dev)
        o_p1 = torch.empty(n, 3, dtype=torch.float64, device=dev)
        o_p2 = torch.empty(n, 3, dtype=torch.float64, device=dev)
        o_b1 = torch.empty(n, dtype=torch.int32, device=dev)
        o_b2 = torch.empty(n, dtype=torch.int32, device=dev)
    else:
        o_cnt = torch.empty(n, dtype=torch.int32, device=dev)
        if wl == "contacts":
            cap = 64 * n
            o_con = torch.empty(cap * 64, dtype=torch.uint8, device=dev)
            o_off = torch.empty(n + 1, dtype=torch.int64, device=dev)

    def step_resident():
        if wl == "distance":
            F.distance_batch_device(env, dP, rob, None, dreq, o_dist, o_p1, o_p2, o_b1, o_b2)
            if world > 1:  # per-GPU results gathered with NCCL allgather over NVLink
                all_gather_records(o_dist, world * n)
                all_gather_records(torch.cat([o_p1, o_p2], dim=1), world * n)
        elif wl == "collide":
            F.collide_batch_device(env, dP, rob, None, creq, o_cnt)
            if world > 1:
                all_gather_records(o_cnt, world * n)
        else:
            F.collide_batch_device(env, dP, rob, None, creq, o_cnt, o_con, o_off)
            if world > 1:
                all_gather_records(o_cnt, world * n)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def timed(fn, steps, warmup, do_flush=True):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = 0.0
        for _ in range(steps):
            if do_flush:
                flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            ms += e0.elapsed_time(e1)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- untimed stats pass: per-query work counters for the roofline accounting ----
    nbv = torch.zeros(n, dtype=torch.int32, device=dev)
    nleaf = torch.zeros(n, dtype=torch.int32, device=dev)
    _capi.set_option("traversal", 0)  # the thread-per-query traversal visits exactly the reference's BVTT nodes
    if wl == "distance":
        F.distance_batch_device(env, dP, rob, None, dreq, o_dist, None, None, None, None, nbv, nleaf)
    else:
        F.collide_batch_device(env, dP, rob, None, creq, o_cnt, None, None, nbv, nleaf)
    F.sync_status(local)
    _capi.set_option("traversal", args.traversal)
    h_nbv, h_nleaf = nbv.cpu().numpy().astype(np.int64), nleaf.cpu().numpy().astype(np.int64)

    # ---- device-resident throughput (value) ----
    sampler = ClockSampler(local)
    launches0 = _capi.launch_count()
    sampler.start()
    total_ms = timed(step_resident, args.steps, args.warmup)
    clocks = sampler.stop()
    launches = _capi.launch_count() - launches0
    launches_timed = launches * args.steps // (args.steps + args.warmup)
    F.sync_status(local)
    value = world * n * args.steps / (total_ms * 1e-3)

    # ---- dominant-kernel roofline: traversal kernel timed alone on this rank ----
    def kernel_only():
        if wl == "distance":
            F.distance_batch_device(env, dP, rob, None, dreq, o_dist, o_p1, o_p2, o_b1, o_b2)
        elif wl == "collide":
            F.collide_batch_device(env, dP, rob, None, creq, o_cnt)
        else:
            F.collide_batch_device(env, dP, rob, None, creq, o_cnt, o_con, o_off)

    world_save, k_ms = world, None
    if rank == 0:
        world = 1
        k_ms = timed(kernel_only, args.steps, 1) / args.steps
        world = world_save
    if world > 1:
        dist.barrier()

    # ---- end to end through the host-pointer API (pinned host buffers, copies inside) ----
    e2e = None
    if not args.no_e2e:
        hp = hP.numpy()

        def step_e2e():
            if wl == "distance":
                r = F.distance_batch(env, hp, rob, None, dreq, device=local, pinned=True)
                return r.min_distance
            if wl == "collide":
                return F.collide_batch(env, hp, rob, None, creq, want_contacts=False, device=local, pinned=True).num_contacts
            return F.collide_batch(env, hp, rob, None, creq, contact_capacity=40 * n, device=local, pinned=True).num_contacts

        for _ in range(2):
            step_e2e()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        if wl == "distance":
            d2h = n * (8 + 24 + 24 + 4 + 4)
        elif wl == "collide":
            d2h = n * 4
        else:
            d2h = n * 4 + (n + 1) * 8 + int(h_nleaf.sum() * 0)  # + contacts, counted below
        e2e = {"value": world * n * args.steps / dt, "unit": UNIT, "h2d_bytes_per_step": n * 96, "d2h_bytes_per_step": d2h}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline numbers ----
    peak, which = measured_peaks()
    if wl == "contacts":
        ncon = o_cnt.cpu().numpy().astype(np.int64)
        if e2e:
            e2e["d2h_bytes_per_step"] += int(ncon.sum()) * 64
    else:
        ncon = None
    alg_bytes = float(algorithmic_bytes(wl, h_nbv, h_nleaf, ncon).sum())
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    try:  # measured DRAM bytes of the dominant kernel (one ncu --set full capture, committed under profiles/)
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            t = json.load(f).get(wl)
        if t and t["poses"] == n and t["traversal"] == args.traversal:
            traffic = t["dram_bytes_per_launch"] * t["launches_per_step"]
            traffic_src = "profiles/r01_traffic.json (%s, %d launch(es) per step)" % (t["kernel"], t["launches_per_step"])
    except Exception:  # pragma: no cover
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": which + " (MEASURED_PEAKS.json hbm_gbs)" if which == "measured" else "fallback",
                "kernel_ms": k_ms, "algorithmic_bytes_per_launch": alg_bytes,
                "mean_n_bv": float(h_nbv.mean()), "mean_n_leaf": float(h_nleaf.mean()),
                "note": "working set (~1 MB of BVH records) is L2/L1 resident; the binding resource is the FP64 pipe, see fp64"}
    # FP64 pipe: measured unfused DMUL+DADD issue rate vs nominal per-test operation counts
    fp64 = None
    try:
        unfused = _capi.microbench(0, local)
        fused = _capi.microbench(1, local)
        l2 = _capi.microbench(2, local)
        F_BV = 430.0 if wl == "distance" else 300.0   # upper-bound mul+add per BV test
        F_LEAF = 1100.0 if wl == "distance" else 860.0
        flops = float((h_nbv * F_BV + h_nleaf * F_LEAF).sum())
        fp64 = {"unfused_ops_per_s_peak": unfused, "dfma_per_s_peak": fused, "l2_read_gbs": l2,
                "algorithmic_ops_upper_bound_per_launch": flops, "achieved_ops_per_s": flops / (k_ms * 1e-3),
                "frac_of_unfused_peak_upper_bound": flops / (k_ms * 1e-3) / unfused}
    except Exception as ex:  # pragma: no cover
        fp64 = {"error": str(ex)}

    # ---- CPU baseline: the oracle on this box's host cores, bounded sample ----
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import pyoracle as O

        O.build()
        oenv, orob = O.Model(ev, et), O.Model(rv, rt)
        threads = O.hardware_threads()
        s = min(args.cpu_sample, n)
        if wl == "distance":
            r = O.distance_batch(oenv, orob, P[:s], None, True, 2, nthreads=threads)
            ok = bool(np.array_equal(r["min_distance"], o_dist.cpu().numpy()[:s]))
        elif wl == "collide":
            r = O.collide_batch(oenv, orob, P[:s], None, 1, False, nthreads=threads)
            ok = bool(np.array_equal(r["counts"], o_cnt.cpu().numpy()[:s]))
        else:
            r = O.collide_batch(oenv, orob, P[:s], None, 100, True, nthreads=threads)
            ok = bool(np.array_equal(r["counts"], o_cnt.cpu().numpy()[:s]))
        counters_ok = bool(np.array_equal(r["n_bv"], h_nbv[:s]) and np.array_equal(r["n_leaf"], h_nleaf[:s]))
        cpu = {"value": s / r["seconds"], "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"first {s} poses of rank 0's batch, {threads} host threads, one pass",
               "matches_gpu": ok, "counters_match_gpu": counters_ok}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS[wl], "poses_per_gpu": n, "global_poses": world * n, "pose_seed": 1,
                   "models": "env.obj (2180 tris, 4359 nodes) posed vs rob.obj (216 tris, 431 nodes) at identity",
                   "l2_flush_between_steps": True, "traversal": _capi.get_option("traversal"), **({"options": args.opt} if args.opt else {}),
                   "multi_gpu": "BVHs replicated, poses partitioned, results all-gathered with NCCL" if world > 1 else "single GPU"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches_timed, "roofline": roofline, "fp64": fp64, "cpu_baseline": cpu,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
