#!/usr/bin/env python
"""bench.py — batched OBBRSS mesh-mesh queries/second on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--workload all|distance|collide|cfg1|contacts|sphere_distance|cfg4|cfg5]
    python bench.py --impl reference ...     # the CPU oracle (the only CPU FCL buildable here)

A step = one pass of the hot path over one batch of synthetic poses.  The headline (top-level keys of the ONE JSON
line rank 0 prints) is BASELINE configs[1]: env.obj vs rob.obj distance() with nearest points, 1M poses per GPU.
With the default --workload all the same line carries, under "workloads", the other env/rob configurations measured
the same way: cfg1 (collide, binary verdict, at its real size of 10k poses and at 1M), cfg3 (contacts, max 100).
cfg4 (7-link arm vs 200k-triangle scene) and cfg5 (two 1M-triangle meshes: collide, distance, tolerance
verification) ride along at reduced sizes (250k configurations / 100k poses per GPU; --no-big skips them) and are
separate invocations at any size (--workload cfg4 | cfg5 --poses N, --scaling weak|strong).

Roofline: per workload, the bound is the SLOWER of
  fp64 : executed mul+add+cmp of the reference's sequential traversal (instrumented oracle, oracle/fcl_oracle_counted.cpp)
         over the measured unfused FP64 issue rate (fclgpu_microbench kind 0), and
  l2   : algorithmic record bytes (96 + n_bv*2*(120|128) + n_leaf*2*72 + out) over the measured L2 read bandwidth
         (fclgpu_microbench kind 2) -- or `hbm` (MEASURED_PEAKS.json hbm_gbs) when the BVHs exceed the L2;
frac = that bound's time / the traversal kernel's measured time (CUDA events around the launch).
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
    "distance": "cfg2: env.obj vs rob.obj distance() with nearest points, 1M random poses per GPU, double",
    "collide": "cfg1 at 1M: env.obj vs rob.obj collide() binary verdict (CollisionRequest()), 1M random poses per GPU",
    "cfg1": "cfg1 at its real size: env.obj vs rob.obj collide() binary verdict, 10k random poses per GPU",
    "contacts": "cfg3: env.obj vs rob.obj collide() enable_contact, num_max_contacts=100, 1M random poses per GPU",
    "sphere_distance": "env.obj at identity vs Sphere(r=100) at the seed-1 pose translations, distance() with nearest points, 1M queries",
    "cfg4": "cfg4: 7-link arm (7 x 4,900-triangle links) vs 199,712-triangle scene, collide() verdict per link, robot configurations sharded across GPUs",
    "cfg5": "cfg5: two 999,680-triangle meshes, collide() verdict + distance() + tolerance verification, poses sharded across GPUs",
}
ENV_ROB = ("distance", "collide", "cfg1", "contacts")
SPHERE_RADIUS = 100.0
L2_BYTES = 126 << 20


def load_meshes():
    g = os.path.join(ROOT, "tests", "golden")
    e, r = np.load(os.path.join(g, "env.npz")), np.load(os.path.join(g, "rob.npz"))
    return (e["verts"], e["tris"]), (r["verts"], r["tris"])


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            pass
    return 6650.0, "fallback of /opt/skills/guides/B200_PROFILING.md"


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


def algorithmic_bytes(kind, n_bv, n_leaf, n_contacts=None):
    """bytes per query from the REFERENCE traversal's counters (120 B of OBB fields / 128 B of RSS
    fields per node, 72 B per triangle, 96 B pose in, result out)."""
    n_bv = np.asarray(n_bv, np.float64)
    n_leaf = np.asarray(n_leaf, np.float64)
    if kind == "distance":
        return 96 + n_bv * 2 * 128 + n_leaf * 2 * 72 + 64
    out = 1.0 if n_contacts is None else 4.0 + 64.0 * np.asarray(n_contacts, np.float64)
    return 96 + n_bv * 2 * 120 + n_leaf * 2 * 72 + out


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


# ------------------------------------------------------------------------------------------------
# CPU side: the oracle as baseline and as the source of the reference traversal's executed work
# ------------------------------------------------------------------------------------------------
def oracle_pair_run(kind, O, m1, m2, tf1, tf2, threads, **kw):
    if kind == "distance":
        return O.distance_batch(m1, m2, tf1, tf2, True, 2, nthreads=threads)
    return O.collide_batch(m1, m2, tf1, tf2, kw.get("num_max_contacts", 1), kw.get("enable_contact", False), nthreads=threads)


def request_of(wl):
    if wl == "contacts":
        return {"num_max_contacts": 100, "enable_contact": True}
    return {"num_max_contacts": 1, "enable_contact": False}


def run_reference(args, rank):
    """--impl reference: the CPU oracle (restatement of the reference; real FCL needs Eigen/libccd, absent from this
    image) on all host threads, bounded sample per step, for the headline workload and -- under "workloads" -- the
    other env/rob configurations of the default line."""
    if rank != 0:
        return
    from fcl_b200.poses import identity_poses, random_poses
    from oracle import pyoracle as O

    O.build()
    (ev, et), (rv, rt) = load_meshes()
    env, rob = O.Model(ev, et), O.Model(rv, rt)
    threads = O.hardware_threads()
    sample = args.cpu_sample
    P = random_poses(sample, seed=1)

    def one(wl, steps, warmup):
        s = min(sample, 10000) if wl == "cfg1" else sample

        def step():
            if wl == "sphere_distance":
                return O.distance_mesh_sphere_batch(env, SPHERE_RADIUS, identity_poses(s), P[:s], nthreads=threads)["seconds"]
            kind = "distance" if wl == "distance" else "collide"
            return oracle_pair_run(kind, O, env, rob, P[:s], None, threads, **request_of(wl))["seconds"]

        for _ in range(warmup):
            step()
        total = float(sum(step() for _ in range(steps)))
        return s * steps / total, 1e3 * total / steps, s

    head = "distance" if args.workload == "all" else args.workload
    if head in ("cfg4", "cfg5"):
        emit(reference_big(args, head, O, threads))
        return
    value, ms, s = one(head, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS[head], "sample_poses_per_step": s},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"first {s} poses of the seed-1 batch per step, {threads} host threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if args.workload == "all":
        subs = {}
        for wl in ENV_ROB:
            v, m, ss = (value, ms, s) if wl == head else one(wl, max(2, min(args.steps, 5)), 1)
            subs[wl] = {"workload": WORKLOADS[wl], "value": v, "unit": UNIT, "ms_per_step": m, "sample_poses_per_step": ss,
                        "cores": threads, "kind": "port"}
        line["workloads"] = subs
    emit(line)


def reference_big(args, head, O, threads):
    """--impl reference for cfg4 / cfg5: the oracle on a bounded pose sample of the same meshes."""
    from fcl_b200 import workloads as W
    from fcl_b200.poses import identity_poses

    s = min(args.cpu_sample, 2000)
    subs = {}
    t0 = time.perf_counter()
    if head == "cfg4":
        (sv, st), links = W.cfg4_meshes()
        scene = O.Model(sv, st)
        olinks = [O.Model(v, t) for v, t in links]
        LP = W.arm_configurations(s, seed=4)
        secs = 0.0
        for _ in range(max(1, min(args.steps, 3))):
            secs = sum(O.collide_batch(scene, olinks[j], None, np.ascontiguousarray(LP[:, j]), 1, False, nthreads=threads)["seconds"]
                       for j in range(7))
        value = 7 * s / secs
        subs["collide"] = {"value": value, "unit": UNIT, "configurations_per_s": s / secs}
    else:
        (va, ta), (vb, tb) = W.cfg5_meshes()
        A, B = O.Model(va, ta), O.Model(vb, tb)
        P = W.shell_poses(s, 1.5, 3.0, seed=6)
        rc = O.collide_batch(A, B, identity_poses(s), P, 1, False, nthreads=threads)
        rd = O.distance_batch(A, B, identity_poses(s), P, True, 2, nthreads=threads)
        value = s / rc["seconds"]
        subs["collide"] = {"value": value, "unit": UNIT}
        subs["distance"] = {"value": s / rd["seconds"], "unit": UNIT}
        subs["tolerance"] = {"value": s / rd["seconds"], "unit": UNIT,
                             "note": "the reference has no tolerance query: a caller compares fcl::distance with the tolerance"}
    return {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * (7 if head == "cfg4" else 1) * s / value, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOADS[head], "sample_poses_per_step": s, "setup_s": time.perf_counter() - t0},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{s} poses / configurations per step, {threads} host threads (BVH build excluded)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "workloads": subs}


# ------------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------------
class Ctx:
    """Process-wide state of one bench run (one rank = one GPU)."""

    def __init__(self, args, rank, world, local):
        import torch
        import torch.distributed as dist

        self.args, self.rank, self.world, self.local = args, rank, world, local
        self.torch, self.dist = torch, dist
        self.dev = torch.device("cuda", local)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)  # > 126 MB L2
        self.comm_stream = torch.cuda.Stream(self.dev) if world > 1 else None
        self.comm = None
        self.peaks = None
        if world > 1:
            # data-path collective through the C ABI (fclgpu_comm_* on raw NCCL); torch.distributed only carries the
            # 128-byte NCCL id to the other ranks and reduces the timing
            import ctypes as C

            from fcl_b200 import _capi

            L = _capi.lib()
            ident = C.create_string_buffer(128)
            if rank == 0:
                rc = L.fclgpu_comm_unique_id(ident)
                assert rc == 0, L.fclgpu_comm_last_error()
            box = [ident.raw]
            dist.broadcast_object_list(box, src=0)
            h = C.c_void_p()
            rc = L.fclgpu_comm_init(local, rank, world, box[0], C.byref(h))
            assert rc == 0, L.fclgpu_comm_last_error()
            self.comm = h

    # ---- step loop on the device clock, L2 flushed before every step, max over ranks.  One GPU: the sum of the per-step
    # event intervals (flush outside the intervals).  Several GPUs: steps overlap (step k's results leave over NVLink
    # while step k+1 computes), so ONE interval brackets all K steps, flushes and the final drain included. ----
    def timed(self, fn, steps, warmup, drain=None):
        torch, dist = self.torch, self.dist
        for k in range(warmup):
            fn(k)
        if drain:
            drain()
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if self.world == 1:
            ms = 0.0
            for k in range(steps):
                self.flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn(k)
                e1.record()
                e1.synchronize()
                ms += e0.elapsed_time(e1)
            torch.cuda.synchronize()
            return ms
        # N > 1: the K steps are enqueued back to back (no host synchronisation inside, so the gather of step k overlaps step
        # k + 1); like at N = 1 the L2 flush between steps is outside the events, and the part of the last gather that is
        # still running when the last step ends is timed by its own pair of events
        if os.environ.get("FCLGPU_BENCH_TIMING") == "loop":  # diagnosis: one pair of events around the loop, flushes included
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for k in range(steps):
                self.flush.fill_(1)
                fn(k)
            if drain:
                drain()
            e1.record()
            e1.synchronize()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
        else:
            evs = []
            for k in range(steps):
                self.flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn(k)
                e1.record()
                evs.append((e0, e1))
            d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            d0.record()
            if drain:
                drain()
            d1.record()
            d1.synchronize()
            torch.cuda.synchronize()
            ms = sum(a.elapsed_time(b) for a, b in evs) + d0.elapsed_time(d1)
        dist.barrier()
        t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- the traversal kernel alone: CUDA events around each launch (flush outside the events) ----
    def timed_kernel(self, fn, steps):
        torch = self.torch
        fn(0)
        torch.cuda.synchronize()
        ms = 0.0
        for k in range(steps):
            self.flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(k)
            e1.record()
            e1.synchronize()
            ms += e0.elapsed_time(e1)
        return ms / steps

    # ---- results of step k leave over NVLink while step k+1 computes (one packed record buffer per step) ----
    def gather_async(self, local_buf, out_buf):
        torch, dist = self.torch, self.dist
        if self.world == 1 or os.environ.get("FCLGPU_BENCH_NO_GATHER"):  # (the switch is for diagnosis only)
            return
        ev = torch.cuda.Event()
        ev.record()
        self.comm_stream.wait_event(ev)
        from fcl_b200 import _capi

        rc = _capi.lib().fclgpu_comm_allgather(self.comm, local_buf.data_ptr(), out_buf.data_ptr(), local_buf.numel() * local_buf.element_size(),
                                               self.comm_stream.cuda_stream)
        assert rc == 0, _capi.lib().fclgpu_comm_last_error()
        # the step after the next one reuses local_buf: it may not start before this gather has read it
        done = torch.cuda.Event()
        done.record(self.comm_stream)
        prev, self._gather_done = getattr(self, "_gather_done", None), done
        if prev is not None:
            torch.cuda.current_stream().wait_event(prev)

    def drain(self):
        if self.world > 1:
            self.torch.cuda.current_stream().wait_stream(self.comm_stream)

    def microbench(self):
        if self.peaks is None:
            from fcl_b200 import _capi

            hbm, src = measured_peaks()
            self.peaks = {"fp64_unfused_ops_per_s": _capi.microbench(0, self.local), "dfma_per_s": _capi.microbench(1, self.local),
                          "l2_read_gbs": _capi.microbench(2, self.local), "hbm_gbs": hbm, "hbm_source": src}
        return self.peaks


def roofline_of(ctx, kind, n, k_ms, nbv_sum, nleaf_sum, bytes_sum, flops_sum, flops_note, model_bytes, traffic=None, extra=None,
                own=None):
    """bound = the slower of the FP64 pipe and the memory level that holds the BVH records, with the work of
    the REFERENCE's sequential traversal.  That model stops being a bound when the kernel's traversal needs less work
    than the reference's (cfg5: the sorted front needs 3.7x fewer box tests than distanceRecurse, and the steering tests
    read 64-byte FP32 records): the reference-work figure then exceeds 1 and is reported under "reference_work", and
    the headline fraction falls back to the kernel's OWN counters (`own` = sums of its n_bv / n_leaf): bytes it
    actually requests -- 144 B per box test (two 64 B records + topo), 160 B per triangle-pair test -- over the
    measured L2 bandwidth, the level every record request goes through."""
    pk = ctx.microbench()
    mem_level = "l2" if model_bytes <= L2_BYTES else "hbm"
    mem_peak = pk["l2_read_gbs"] if mem_level == "l2" else pk["hbm_gbs"]
    t_fp64 = flops_sum / pk["fp64_unfused_ops_per_s"]
    t_mem = bytes_sum / (mem_peak * 1e9)
    sec = k_ms * 1e-3
    if t_fp64 >= t_mem:
        r = {"bound": "fp64", "achieved": flops_sum / sec / 1e12, "peak": pk["fp64_unfused_ops_per_s"] / 1e12, "unit": "TFLOP/s",
             "peak_source": "fclgpu_microbench(0): separately rounded DMUL+DADD issue rate measured on this GPU in this run"}
    else:
        r = {"bound": mem_level, "achieved": bytes_sum / sec / 1e9, "peak": mem_peak, "unit": "GB/s",
             "peak_source": ("fclgpu_microbench(2): L2-resident read bandwidth measured on this GPU in this run" if mem_level == "l2"
                             else pk["hbm_source"])}
    r["frac"] = r["achieved"] / r["peak"]
    r["basis"] = "reference traversal's work (SURVEY 8d)"
    ref = {"fp64": {"executed_ops_per_launch": flops_sum, "bound_ms": 1e3 * t_fp64, "frac": t_fp64 / sec, "ops_source": flops_note},
           mem_level: {"algorithmic_bytes_per_launch": bytes_sum, "bound_ms": 1e3 * t_mem, "frac": t_mem / sec, "peak_gbs": mem_peak},
           "mean_n_bv": nbv_sum / n, "mean_n_leaf": nleaf_sum / n}
    if r["frac"] > 1.0 and own is not None:
        out_b = 64.0 if kind == "distance" else 4.0
        own_bytes = own["nbv_sum"] * 144.0 + own["nleaf_sum"] * 160.0 + n * (96.0 + out_b)
        t_own = own_bytes / (pk["l2_read_gbs"] * 1e9)
        r = {"bound": "l2", "achieved": own_bytes / sec / 1e9, "peak": pk["l2_read_gbs"], "unit": "GB/s", "frac": t_own / sec,
             "peak_source": "fclgpu_microbench(2): L2-resident read bandwidth measured on this GPU in this run",
             "basis": ("the kernel's OWN traversal counters: the reference-work bound exceeds 1 because this traversal needs %.2fx fewer "
                       "box tests than the reference's recursion and reads 64 B FP32 records" % (nbv_sum / max(1.0, own["nbv_sum"]))),
             "own_work": {"mean_n_bv": own["nbv_sum"] / n, "mean_n_leaf": own["nleaf_sum"] / n, "bytes_per_launch": own_bytes},
             "reference_work": ref}
    else:
        r.update(ref)
    r["traffic"] = traffic
    r.update({"kernel_ms": k_ms, "bvh_record_bytes": model_bytes,
              "definition": "frac = max(executed FP64 mul+add+cmp / unfused FP64 rate, algorithmic bytes / bandwidth of the level "
                            "holding the BVH) / kernel time; work counted on the REFERENCE's sequential traversal (SURVEY 8d) "
                            "unless `basis` says otherwise"})
    if extra:
        r.update(extra)
    return r


def traffic_of(wl, n, traversal):
    """Measured DRAM bytes of the dominant kernel from the committed ncu --set full capture (profiles/*traffic.json)."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                t = json.load(f).get(wl)
            if t and t["poses"] == n and t.get("traversal", traversal) == traversal:
                return t["dram_bytes_per_launch"] * t["launches_per_step"], "profiles/%s (%s, %d launch(es) per step)" % (
                    name, t["kernel"], t["launches_per_step"])
        except Exception:
            pass
    return None, None


def run_env_rob(ctx, wl, n, steps, warmup, models, meshes, with_cpu):
    """One env.obj-vs-rob.obj workload: device-resident throughput (value), the traversal kernel alone (roofline),
    end to end through the host-buffer API (e2e), the CPU oracle beside it (cpu_baseline)."""
    import fcl_b200 as F
    from fcl_b200 import _capi

    torch, dist = ctx.torch, ctx.dist
    args, world, rank, local, dev = ctx.args, ctx.world, ctx.rank, ctx.local, ctx.dev
    env, rob = models
    (ev, et), (rv, rt) = meshes
    kind = "distance" if wl == "distance" else "collide"
    P = F.random_poses(n, seed=1, start=rank * n)  # rank r owns poses [r*n, (r+1)*n) of the global batch
    hP = torch.from_numpy(P).pin_memory()
    dP = hP.to(dev)
    rq = request_of(wl)
    creq = F.CollisionRequest(rq["num_max_contacts"], rq["enable_contact"])
    dreq = F.DistanceRequest(True)

    # resident outputs: ONE packed record buffer per step parity (64 B per query for distance, 4 B for collide), so that a
    # step's results leave with a single all-gather while the next step computes into the other buffer
    rec = 64 if kind == "distance" else 4
    bufs, gath = [], []
    for _ in range(2 if world > 1 else 1):
        b = torch.empty(n * rec, dtype=torch.uint8, device=dev)
        bufs.append(b)
        gath.append(torch.empty(world * n * rec, dtype=torch.uint8, device=dev) if world > 1 else None)

    def views(b):
        if kind == "distance":
            return (b[:8 * n].view(torch.float64), b[8 * n:32 * n].view(torch.float64).view(n, 3),
                    b[32 * n:56 * n].view(torch.float64).view(n, 3), b[56 * n:60 * n].view(torch.int32), b[60 * n:].view(torch.int32))
        return (b.view(torch.int32),)

    if wl == "contacts":
        cap = 64 * n
        o_con = torch.empty(cap * 64, dtype=torch.uint8, device=dev)
        o_off = torch.empty(n + 1, dtype=torch.int64, device=dev)

    def compute(k, nbv=None, nleaf=None):
        v = views(bufs[k % len(bufs)])
        if kind == "distance":
            F.distance_batch_device(env, dP, rob, None, dreq, v[0], v[1], v[2], v[3], v[4], nbv, nleaf)
        elif wl == "contacts":
            F.collide_batch_device(env, dP, rob, None, creq, v[0], o_con, o_off, nbv, nleaf)
        else:
            F.collide_batch_device(env, dP, rob, None, creq, v[0], None, None, nbv, nleaf)

    def step(k):
        compute(k)
        if world > 1:  # per-GPU results gathered with NCCL all-gather over NVLink (contact blocks stay on their GPU)
            ctx.gather_async(bufs[k % 2], gath[k % 2])

    # ---- untimed stats pass: the REFERENCE traversal's counters for the roofline accounting ----
    nbv = torch.zeros(n, dtype=torch.int32, device=dev)
    nleaf = torch.zeros(n, dtype=torch.int32, device=dev)
    trav = _capi.get_option("traversal")
    _capi.set_option("traversal", 0)  # the thread-per-query traversal visits exactly the reference's BVTT nodes
    if kind == "distance":
        F.distance_batch_device(env, dP, rob, None, dreq, views(bufs[0])[0], None, None, None, None, nbv, nleaf)
    else:
        F.collide_batch_device(env, dP, rob, None, creq, views(bufs[0])[0], None, None, nbv, nleaf)
    F.sync_status(local)
    _capi.set_option("traversal", trav)
    h_nbv, h_nleaf = nbv.cpu().numpy().astype(np.int64), nleaf.cpu().numpy().astype(np.int64)

    # ---- device-resident throughput ----
    launches0 = _capi.launch_count()
    total_ms = ctx.timed(step, steps, warmup, ctx.drain)
    launches = (_capi.launch_count() - launches0) * steps // (steps + warmup)
    F.sync_status(local)
    value = world * n * steps / (total_ms * 1e-3)

    # ---- dominant kernel alone (rank 0) ----
    k_ms = ctx.timed_kernel(compute, steps) if rank == 0 else None
    F.sync_status(local)
    result0 = views(bufs[0])[0].clone()
    ncon = result0.cpu().numpy().astype(np.int64) if wl == "contacts" else None

    # ---- end to end through the host-pointer API (pinned host buffers, copies inside the timed region) ----
    e2e = None
    if not args.no_e2e:
        hp = hP.numpy()

        def step_e2e():
            if kind == "distance":
                return F.distance_batch(env, hp, rob, None, dreq, device=local, pinned=True).min_distance
            if wl == "contacts":
                return F.collide_batch(env, hp, rob, None, creq, contact_capacity=40 * n, device=local, pinned=True).num_contacts
            return F.collide_batch(env, hp, rob, None, creq, want_contacts=False, device=local, pinned=True).num_contacts

        for _ in range(2):
            step_e2e()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step_e2e()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        if kind == "distance":
            d2h = n * (8 + 24 + 24 + 4 + 4)
        elif wl == "contacts":
            d2h = n * 4 + (n + 1) * 8 + int(ncon.sum()) * 64 if ncon is not None else None
        else:
            d2h = n * 4
        e2e = {"value": world * n * steps / dt, "unit": UNIT, "h2d_bytes_per_step": n * 96, "d2h_bytes_per_step": d2h,
               "note": "per-rank host buffers; at N > 1 every rank runs its shard through the host API concurrently"}
        if wl == "contacts" and world == 1:
            # the same lists in the compact record formats (extension, include/fclgpu.h): the copy back is the long pole
            for name, fmt, rec in (("ids", F.CONTACT_IDS, 8), ("f32", F.CONTACT_F32, 40)):
                def step_c():
                    return F.collide_batch(env, hp, rob, None, creq, contact_capacity=40 * n, device=local, pinned=True, contact_format=fmt).num_contacts
                step_c()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                reps = max(2, steps // 2)
                for _ in range(reps):
                    step_c()
                torch.cuda.synchronize()
                e2e["compact_" + name] = {"value": n * reps / (time.perf_counter() - t0), "unit": UNIT, "record_bytes": rec,
                                          "d2h_bytes_per_step": n * 4 + (n + 1) * 8 + int(ncon.sum()) * rec}
    if rank != 0:
        return None

    # ---- CPU baseline + executed work of the reference traversal (oracle, bounded sample) ----
    s = min(args.cpu_sample, n)
    cpu, flops_sum, flops_note = None, None, None
    if with_cpu:
        from oracle import pyoracle as O

        O.build()
        oenv, orob = O.Model(ev, et), O.Model(rv, rt)
        threads = O.hardware_threads()
        r = oracle_pair_run(kind, O, oenv, orob, P[:s], None, threads, **rq)
        got = result0.cpu().numpy()[:s]
        ok = bool(np.array_equal(r["min_distance"] if kind == "distance" else r["counts"], got))
        counters_ok = bool(np.array_equal(r["n_bv"], h_nbv[:s]) and np.array_equal(r["n_leaf"], h_nleaf[:s]))
        cpu = {"value": s / r["seconds"], "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"first {s} poses of rank 0's batch, {threads} host threads, one pass",
               "matches_gpu": ok, "counters_match_gpu": counters_ok}
        c = O.counted_query(kind, O.CountedModel(ev, et), O.CountedModel(rv, rt), P[:s], None, nthreads=threads, **rq)
        ops = c["ops"].sum(axis=0).astype(np.float64)
        sample_flops = float(ops[:3].sum())
        # scale the sample's executed operations to the batch by the reference traversal's own work counters
        scale = 0.5 * float(h_nbv.sum() / max(1, c["n_bv"].sum()) + h_nleaf.sum() / max(1, c["n_leaf"].sum()))
        flops_sum = sample_flops * scale
        flops_note = ("instrumented oracle (oracle/fcl_oracle_counted.cpp) on the first %d poses: executed mul %.4g add %.4g cmp %.4g "
                      "(div %.3g, sqrt %.3g not counted) per query, scaled x%.2f to the batch by the reference traversal's n_bv / n_leaf"
                      % (s, ops[0] / s, ops[1] / s, ops[2] / s, ops[3] / s, ops[4] / s, scale))
    else:  # nominal fallback when the oracle leg is switched off: SURVEY 8(d) upper bounds halved (measured ratio)
        F_BV, F_LEAF = (430.0, 1100.0) if kind == "distance" else (300.0, 860.0)
        flops_sum = 0.5 * float((h_nbv * F_BV + h_nleaf * F_LEAF).sum())
        flops_note = "no oracle leg (--no-cpu-baseline): 0.5 x SURVEY 8(d) upper bounds"

    bytes_sum = float(algorithmic_bytes(kind, h_nbv, h_nleaf, ncon).sum())
    model_bytes = (env.getNumBVs() + rob.getNumBVs()) * 128 + (env.num_tris + rob.num_tris) * 72
    traffic, traffic_src = traffic_of(wl, n, trav)
    roof = roofline_of(ctx, kind, n, k_ms, float(h_nbv.sum()), float(h_nleaf.sum()), bytes_sum, flops_sum, flops_note, model_bytes,
                       traffic, {"traffic_source": traffic_src})
    return {"workload": WORKLOADS[wl], "poses_per_gpu": n, "value": value, "unit": UNIT, "ms_per_step": total_ms / steps,
            "steps": steps, "e2e": e2e, "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu,
            **({"mean_contacts_per_query": float(ncon.mean())} if ncon is not None else {})}


def run_sphere_distance(ctx):
    """--workload sphere_distance (single GPU): the row next to the path, SURVEY 8f rank 2."""
    import ctypes as C

    import fcl_b200 as F
    from fcl_b200 import _capi

    torch, args, local = ctx.torch, ctx.args, ctx.local
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

    def kernel(k=0):
        rc = L.fclgpu_distance_mesh_sphere_batch(env.device_model(local), SPHERE_RADIUS, n, None, dS.data_ptr(), C.byref(rq),
                                                 dist_d.data_ptr(), p1.data_ptr(), p2.data_ptr(), b1.data_ptr(), None, None, None,
                                                 torch.cuda.current_stream().cuda_stream)
        assert rc == 0, rc

    sampler = ClockSampler(local)
    sampler.start()
    launches0 = _capi.launch_count()
    total_ms = ctx.timed(kernel, args.steps, args.warmup)
    launches = (_capi.launch_count() - launches0) * args.steps // (args.steps + args.warmup)
    clocks = sampler.stop()
    k_ms = ctx.timed_kernel(kernel, args.steps)
    F.sync_status(local)
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
    pk = ctx.microbench()
    achieved = alg / (k_ms * 1e-3) / 1e9
    traffic, traffic_src = traffic_of("sphere_distance", n, None)
    emit({
        "metric": METRIC, "value": n * args.steps / (total_ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS["sphere_distance"], "poses_per_gpu": n, "pose_seed": 1, "l2_flush_between_steps": True,
                   "sphere_leaf_trigger": _capi.get_option("sphere_leaf_trigger"), "sphere_bound32": _capi.get_option("sphere_bound32"),
                   "sphere_blocks": _capi.get_option("sphere_blocks"), "multi_gpu": "single GPU"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
        "roofline": {"bound": "l2", "achieved": achieved, "peak": pk["l2_read_gbs"], "unit": "GB/s", "frac": achieved / pk["l2_read_gbs"],
                     "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": "fclgpu_microbench(2): L2-resident read bandwidth measured on this GPU in this run",
                     "kernel_ms": k_ms, "algorithmic_bytes_per_launch": alg,
                     "mean_n_bv": float(ref["n_bv"].mean()), "mean_n_leaf": float(ref["n_leaf"].mean()),
                     "note": "counters of the reference's traversal from the oracle on the CPU sample, scaled to the batch; records "
                             "are L1/L2 resident (1 MB of BVH)"},
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
    ap.add_argument("--workload", default="all", choices=["all"] + sorted(WORKLOADS))
    ap.add_argument("--poses", type=int, default=1_000_000, help="poses per GPU per step (cfg4: robot configurations)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak = --poses per GPU (default), strong = --poses in total, split across the GPUs")
    ap.add_argument("--cpu-sample", type=int, default=20000)
    ap.add_argument("--traversal", type=int, default=3, help="kernel variant (fclgpu option 'traversal', see DESIGN.md)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-big", action="store_true", help="--workload all: skip the cfg4 / cfg5 legs")
    ap.add_argument("--opt", action="append", default=[], help="name=value library option (A/B runs; recorded in config)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    ensure_built(local)
    import fcl_b200 as F
    from fcl_b200 import _capi

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
    ctx = Ctx(args, rank, world, local)

    if args.workload == "sphere_distance":
        if world > 1:
            raise SystemExit("--workload sphere_distance is a single-GPU line (run it without torchrun)")
        run_sphere_distance(ctx)
        return
    if args.workload in ("cfg4", "cfg5"):
        import bench_big

        line = bench_big.run(ctx, args.workload)
        if rank == 0:
            emit(line)
        if world > 1:
            dist.destroy_process_group()
        return

    meshes = load_meshes()
    (ev, et), (rv, rt) = meshes
    env, rob = F.BVHModel.from_arrays(ev, et), F.BVHModel.from_arrays(rv, rt)
    env.device_model(local)
    rob.device_model(local)  # BVHs replicated on every GPU
    head = "distance" if args.workload == "all" else args.workload
    names = list(ENV_ROB) if args.workload == "all" else [head]
    with_cpu = not args.no_cpu_baseline
    results = {}
    sampler = ClockSampler(local)
    sampler.start()
    for wl in names:
        n = 10_000 if wl == "cfg1" else (args.poses // world if args.scaling == "strong" else args.poses)
        steps = args.steps if wl == head else max(3, min(args.steps, 10))
        results[wl] = run_env_rob(ctx, wl, n, steps, args.warmup, (env, rob), meshes, with_cpu)
    clocks = sampler.stop()
    if args.workload == "all" and not args.no_big:
        # BASELINE configs[3] and [4] in the same line (their own invocations, --workload cfg4 | cfg5, take --poses): cfg4 at
        # 250k robot configurations (1.75M link queries) per GPU, which keeps the default run short (the rate does not depend
        # on the batch there: 21.4 ms per 1M), cfg5 at BASELINE's full 1M poses per GPU: its launches are tail-sensitive --
        # single queries take up to ~2 ms and a 100k-pose launch is only 2.3 .. 24 ms long, so the longest query of a shard
        # decides up to 40 % of such a launch (shards of the same distribution: 2.34 .. 4.02 ms), at 1M poses a few per cent
        import copy

        import bench_big

        keep = ("value", "unit", "ms_per_step", "e2e", "roofline", "cpu_baseline", "gpu_launches", "config", "workloads", "clocks")
        for which, poses in (("cfg4", 250_000), ("cfg5", 1_000_000)):
            a2 = copy.copy(args)
            a2.poses = poses // world if args.scaling == "strong" else poses
            a2.steps = max(3, min(args.steps, 5))
            ctx.args = a2
            big = bench_big.run(ctx, which)
            ctx.args = args
            if rank == 0 and big is not None:
                results[which] = {k: big[k] for k in keep if k in big}
    if rank == 0:
        h = results[head]
        line = {
            "metric": METRIC, "value": h["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": h["ms_per_step"], "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOADS[head], "poses_per_gpu": h["poses_per_gpu"], "global_poses": world * h["poses_per_gpu"],
                       "pose_seed": 1, "models": "env.obj (2180 tris, 4359 nodes) posed vs rob.obj (216 tris, 431 nodes) at identity",
                       "l2_flush_between_steps": True, "l2_flush_inside_timed_region": os.environ.get("FCLGPU_BENCH_TIMING") == "loop" and world > 1,
                       "timing": "per-step CUDA events on the launching stream, summed" + (" + exposed tail of the last gather" if world > 1 else ""),
                       "traversal": _capi.get_option("traversal"), **({"options": args.opt} if args.opt else {}),
                       "multi_gpu": ("BVHs replicated, poses partitioned, one packed 64 B/query result record all-gathered per step with "
                                     "fclgpu_comm_allgather (C ABI, raw NCCL) on a second stream, overlapped with the next step's traversal")
                                    if world > 1 else "single GPU"},
            "clocks": clocks, "e2e": h["e2e"], "gpu_launches": h["gpu_launches"], "roofline": h["roofline"],
            "cpu_baseline": h["cpu_baseline"], "peaks": ctx.microbench(),
        }
        if args.workload == "all":
            line["workloads"] = results
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
