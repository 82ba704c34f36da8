"""bench_big.py — BASELINE configs[3] (cfg4) and configs[4] (cfg5) as bench.py workloads (bench.py --workload cfg4|cfg5).

cfg4: 7-link arm (7 meshes of 4,900 triangles) against a 199,712-triangle scene; a batch of robot configurations drawn
      uniformly in joint space, 7 link poses each through a fixed serial-chain forward kinematics; the configurations are
      sharded across the GPUs with the 7 link queries of one configuration kept on one GPU; a step = the 7
      collide() verdict batches + the per-configuration verdict (any link hits) + its all-gather.
cfg5: two 999,680-triangle meshes; a step = collide() verdicts, distance() with nearest points and the tolerance
      verification extension over the rank's poses, each timed on its own.
--scaling weak: --poses per GPU; --scaling strong: --poses in total (split across the GPUs).
Meshes are built ON the device (fclgpu_model_build_obbrss); BVHs are replicated on every GPU.
"""
import os
import time

import numpy as np

from bench import METRIC, UNIT, WORKLOADS, ClockSampler, algorithmic_bytes, roofline_of


def _shard(total, rank, world):
    base, rem = divmod(int(total), int(world))
    s = rank * base + min(rank, rem)
    return s, base + (1 if rank < rem else 0)


def _cpu_leg(ctx, kind, meshes1, meshes2, tf1, tf2, rq, got, what):
    """Oracle on a bounded sample: baseline q/s, parity flag, the reference traversal's counters and executed ops."""
    from oracle import pyoracle as O

    O.build()
    threads = O.hardware_threads()
    t0 = time.perf_counter()
    m1, m2 = O.Model(*meshes1), O.Model(*meshes2)
    build_s = time.perf_counter() - t0
    if kind == "distance":
        r = O.distance_batch(m1, m2, tf1, tf2, True, 2, nthreads=threads)
        ok = bool(np.array_equal(r["min_distance"], got))
    else:
        r = O.collide_batch(m1, m2, tf1, tf2, rq.get("num_max_contacts", 1), rq.get("enable_contact", False), nthreads=threads)
        ok = bool(np.array_equal(r["counts"], got))
    s = len(got)
    c = O.counted_query(kind, O.CountedModel.share(m1), O.CountedModel.share(m2), tf1, tf2, nthreads=threads,
                        **({} if kind == "distance" else rq))
    ops = c["ops"].sum(axis=0).astype(np.float64)
    cpu = {"value": s / r["seconds"], "unit": UNIT, "cores": threads, "kind": "port",
           "sample": f"{what}: first {s} queries of rank 0's batch, {threads} host threads, one pass (host BVH build {build_s:.1f} s excluded)",
           "matches_gpu": ok}
    note = ("instrumented oracle on the same %d-query sample: executed mul %.4g add %.4g cmp %.4g per query (div %.3g, sqrt %.3g not "
            "counted), scaled to the batch by the query count" % (s, ops[0] / s, ops[1] / s, ops[2] / s, ops[3] / s, ops[4] / s))
    return cpu, r["n_bv"].astype(np.float64), r["n_leaf"].astype(np.float64), float(ops[:3].sum()), note, (m1, m2)


def run(ctx, which):
    import fcl_b200 as F
    from fcl_b200 import _capi
    from fcl_b200 import workloads as W

    torch, dist = ctx.torch, ctx.dist
    args, rank, world, local, dev = ctx.args, ctx.rank, ctx.world, ctx.local, ctx.dev
    strong = args.scaling == "strong"
    total = args.poses if strong else args.poses * world
    start, n = _shard(total, rank, world)
    steps, warmup = args.steps, args.warmup
    sampler = ClockSampler(local)
    t_setup = time.perf_counter()
    sub = {}

    if which == "cfg4":
        (sv, st), links = W.cfg4_meshes()
        scene = F.BVHModel.from_arrays(sv, st, build_on_device=True)
        scene.device_model(local)
        lmods = []
        for lv, lt in links:
            m = F.BVHModel.from_arrays(lv, lt, build_on_device=True)
            m.device_model(local)
            lmods.append(m)
        LP = W.arm_configurations(n, seed=4, start=start)  # (n, 7, 12): rank r owns configurations [start, start + n)
        hP = torch.from_numpy(np.ascontiguousarray(LP.transpose(1, 0, 2))).pin_memory()  # link-major: 7 batches of n poses
        dP = hP.to(dev)
        cnt = torch.zeros(7, n, dtype=torch.int32, device=dev)
        bufs = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(2)]
        gath = [torch.empty(world * n, dtype=torch.uint8, device=dev) if world > 1 else None for _ in range(2)]
        req = F.CollisionRequest()
        torch.cuda.synchronize()
        setup_s = time.perf_counter() - t_setup

        def compute(k):
            for j in range(7):
                F.collide_batch_device(scene, None, lmods[j], dP[j], req, cnt[j])

        def step(k):
            compute(k)
            torch.any(cnt > 0, dim=0, out=bufs[k % 2].view(torch.bool))  # configuration in collision: any link hits
            if world > 1 and not strong_ragged:
                ctx.gather_async(bufs[k % 2], gath[k % 2])

        strong_ragged = world > 1 and (total % world != 0)
        sampler.start()
        launches0 = _capi.launch_count()
        total_ms = ctx.timed(step, steps, warmup, ctx.drain)
        launches = (_capi.launch_count() - launches0) * steps // (steps + warmup)
        clocks = sampler.stop()
        F.sync_status(local)
        nq_global = 7 * total
        value = nq_global * steps / (total_ms * 1e-3)
        k_ms = ctx.timed_kernel(compute, max(3, steps // 2)) if rank == 0 else None
        e2e = None
        if not args.no_e2e:
            hp = hP.numpy()

            def step_e2e():
                hit = np.zeros(n, bool)
                for j in range(7):
                    hit |= F.collide_batch(scene, None, lmods[j], hp[j], req, want_contacts=False, device=local, pinned=True).num_contacts > 0
                return hit

            step_e2e()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(max(2, steps // 4)):
                step_e2e()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / max(2, steps // 4)
            if world > 1:
                tt = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                dt = float(tt.item())
            e2e = {"value": nq_global / dt, "unit": UNIT, "h2d_bytes_per_step": 7 * n * 96, "d2h_bytes_per_step": 7 * n * 4}
        if rank != 0:
            return None
        roof, cpu = None, None
        if not args.no_cpu_baseline:
            s = min(args.cpu_sample, 2000, n)
            j = 6  # the end effector's link: the oracle leg runs one link (same mesh family, same pose distribution)
            got = cnt[j, :s].cpu().numpy()
            cpu, nbv, nleaf, flops_s, note, _ = _cpu_leg(ctx, "collide", (sv, st), links[j], None, np.ascontiguousarray(LP[:s, j]),
                                                         {"num_max_contacts": 1, "enable_contact": False}, got, "link 6 vs scene")
            nq_rank = 7 * n
            bytes_sum = float(algorithmic_bytes("collide", nbv, nleaf).mean()) * nq_rank
            model_bytes = (scene.getNumBVs() + lmods[0].getNumBVs()) * 128 + (scene.num_tris + lmods[0].num_tris) * 72
            roof = roofline_of(ctx, "collide", nq_rank, k_ms, float(nbv.mean()) * nq_rank, float(nleaf.mean()) * nq_rank, bytes_sum,
                               flops_s / s * nq_rank, note, model_bytes)
        return {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
                "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": WORKLOADS["cfg4"], "configurations_per_gpu": n, "global_configurations": total,
                           "link_queries_per_step": nq_global, "configurations_per_s": total * steps / (total_ms * 1e-3),
                           "colliding_configurations_frac": float(bufs[0].float().mean().item()),
                           "seeds": {"scene": 1, "links": "10..16", "joint_angles": 4}, "setup_s": setup_s,
                           "l2_flush_between_steps": True, "traversal": _capi.get_option("traversal"),
                           "multi_gpu": ("configurations partitioned (7 link queries of one configuration on one GPU), 1 B/configuration "
                                         "verdict all-gathered with NCCL, overlapped with the next step") if world > 1 else "single GPU"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu, "peaks": ctx.microbench()}

    # ------------------------------------------------------------------ cfg5
    (va, ta), (vb, tb) = W.cfg5_meshes()
    A = F.BVHModel.from_arrays(va, ta, build_on_device=True)
    B = F.BVHModel.from_arrays(vb, tb, build_on_device=True)
    A.device_model(local)
    B.device_model(local)
    start += int(os.environ.get("FCLGPU_BENCH_SHARD_SHIFT", "0")) * n  # diagnosis: time another rank's shard on one GPU
    P = W.shell_poses(n, 1.5, 3.0, seed=6, start=start)  # centre distance 1.5 .. 3 radii: about half collide
    hP = torch.from_numpy(P).pin_memory()
    dP = hP.to(dev)
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t_setup
    tol = 0.05  # tolerance verification: "is the clearance <= 0.05 radii?"
    cutoff = float(np.nextafter(tol, np.inf))
    creq, dreq, dreq0 = F.CollisionRequest(), F.DistanceRequest(True), F.DistanceRequest(False)
    rec = {"collide": 4, "distance": 64, "tolerance": 8}
    bufs = {k: [torch.empty(n * r, dtype=torch.uint8, device=dev) for _ in range(2)] for k, r in rec.items()}
    gath = {k: [torch.empty(world * n * r, dtype=torch.uint8, device=dev) if world > 1 else None for _ in range(2)] for k, r in rec.items()}
    L = _capi.lib()
    within_buf = [torch.zeros(n, dtype=torch.uint8, device=dev) for _ in range(2)]

    def dviews(b):
        return (b[:8 * n].view(torch.float64), b[8 * n:32 * n].view(torch.float64).view(n, 3), b[32 * n:56 * n].view(torch.float64).view(n, 3),
                b[56 * n:60 * n].view(torch.int32), b[60 * n:].view(torch.int32))

    def compute(kind):
        def f(k):
            b = bufs[kind][k % 2]
            if kind == "collide":
                F.collide_batch_device(A, None, B, dP, creq, b.view(torch.int32))
            elif kind == "distance":
                v = dviews(b)
                F.distance_batch_device(A, None, B, dP, dreq, v[0], v[1], v[2], v[3], v[4])
            else:
                rc = L.fclgpu_within_tolerance_batch(A.device_model(local), B.device_model(local), n, None, dP.data_ptr(), tol,
                                                     within_buf[k % 2].data_ptr(), b.view(torch.float64).data_ptr(), None, None,
                                                     torch.cuda.current_stream().cuda_stream)
                assert rc == 0, rc
        return f

    ragged = world > 1 and (total % world != 0)

    def stepper(kind):
        c = compute(kind)

        def f(k):
            c(k)
            if world > 1 and not ragged:
                ctx.gather_async(bufs[kind][k % 2], gath[kind][k % 2])
        return f

    sampler.start()
    launches0 = _capi.launch_count()
    for kind in ("collide", "distance", "tolerance"):
        ms = ctx.timed(stepper(kind), steps, warmup, ctx.drain)
        sub[kind] = {"value": total * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps}
    launches = (_capi.launch_count() - launches0) * steps // (steps + warmup)
    clocks = sampler.stop()
    F.sync_status(local)
    kms, own = {}, {}
    if rank == 0:
        for kind in ("collide", "distance", "tolerance"):
            kms[kind] = ctx.timed_kernel(compute(kind), max(3, steps // 2))
        # the kernels' own work counters (one stats launch each, not timed)
        nbv = torch.zeros(n, dtype=torch.int32, device=dev)
        nlf = torch.zeros(n, dtype=torch.int32, device=dev)
        F.collide_batch_device(A, None, B, dP, creq, bufs["collide"][0].view(torch.int32), None, None, nbv, nlf)
        own["collide"] = {"nbv_sum": float(nbv.sum().item()), "nleaf_sum": float(nlf.sum().item())}
        v = dviews(bufs["distance"][0])
        F.distance_batch_device(A, None, B, dP, dreq, v[0], v[1], v[2], v[3], v[4], nbv, nlf)
        own["distance"] = {"nbv_sum": float(nbv.sum().item()), "nleaf_sum": float(nlf.sum().item())}
        F.sync_status(local)
    if not args.no_e2e:
        hp = hP.numpy()
        ident = None
        for kind in ("collide", "distance", "tolerance"):
            def one():
                if kind == "collide":
                    return F.collide_batch(A, ident, B, hp, creq, want_contacts=False, device=local, pinned=True)
                if kind == "distance":
                    return F.distance_batch(A, ident, B, hp, dreq, device=local, pinned=True)
                return F.within_tolerance_batch(A, ident, B, hp, tol, device=local, pinned=True)
            one()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            reps = max(2, steps // 4)
            t0 = time.perf_counter()
            for _ in range(reps):
                one()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / reps
            if world > 1:
                tt = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                dt = float(tt.item())
            sub[kind]["e2e"] = {"value": total / dt, "unit": UNIT, "h2d_bytes_per_step": n * 96,
                                "d2h_bytes_per_step": n * {"collide": 4, "distance": 64, "tolerance": 8}[kind]}
    if rank != 0:
        return None
    dist_res = dviews(bufs["distance"][0])[0]
    tol_res = bufs["tolerance"][0].view(torch.float64)
    cnt_res = bufs["collide"][0].view(torch.int32)
    within = within_buf[0] != 0
    checks = {"tolerance_equals_distance_le_tol": bool(torch.equal(within, dist_res <= tol)),
              # within: true distance <= witness <= tol (the witness is a real pair's distance); not within: witness = cutoff
              "witness_consistent": bool((((tol_res >= dist_res) & (tol_res <= tol)) | (~within & (tol_res == cutoff))).all().item()),
              "colliding_frac": float((cnt_res > 0).float().mean().item()), "within_tolerance_frac": float(within.float().mean().item()),
              "colliding_implies_zero_distance": bool((dist_res[cnt_res > 0] == 0).all().item())}
    model_bytes = (A.getNumBVs() + B.getNumBVs()) * 128 + (A.num_tris + B.num_tris) * 72
    if not args.no_cpu_baseline:
        s = min(args.cpu_sample, 1000, n)
        ident_s = F.identity_poses(s)
        got = cnt_res[:s].cpu().numpy()
        cpu_c, nbv, nleaf, fl, note, oms = _cpu_leg(ctx, "collide", (va, ta), (vb, tb), ident_s, P[:s],
                                                    {"num_max_contacts": 1, "enable_contact": False}, got, "collide")
        sub["collide"]["cpu_baseline"] = cpu_c
        sub["collide"]["roofline"] = roofline_of(ctx, "collide", n, kms["collide"], float(nbv.mean()) * n, float(nleaf.mean()) * n,
                                                 float(algorithmic_bytes("collide", nbv, nleaf).mean()) * n, fl / s * n, note, model_bytes,
                                                 own=own["collide"])
        sub["collide"]["box_tests_per_s"] = own["collide"]["nbv_sum"] / (kms["collide"] * 1e-3)
        from oracle import pyoracle as O

        threads = O.hardware_threads()
        rd = O.distance_batch(oms[0], oms[1], ident_s, P[:s], True, 2, nthreads=threads)
        cd = O.counted_query("distance", O.CountedModel.share(oms[0]), O.CountedModel.share(oms[1]), ident_s, P[:s], nthreads=threads)
        opsd = cd["ops"].sum(axis=0).astype(np.float64)
        gd = dist_res[:s].cpu().numpy()
        sub["distance"]["cpu_baseline"] = {"value": s / rd["seconds"], "unit": UNIT, "cores": threads, "kind": "port",
                                           "sample": f"distance: first {s} poses of rank 0's batch, {threads} host threads, one pass",
                                           "matches_gpu": bool(np.array_equal(rd["min_distance"], gd))}
        nb, nl = rd["n_bv"].astype(np.float64), rd["n_leaf"].astype(np.float64)
        noted = "instrumented oracle on the same %d-pose sample: executed mul %.4g add %.4g cmp %.4g per query" % (
            s, opsd[0] / s, opsd[1] / s, opsd[2] / s)
        sub["distance"]["roofline"] = roofline_of(ctx, "distance", n, kms["distance"], float(nb.mean()) * n, float(nl.mean()) * n,
                                                  float(algorithmic_bytes("distance", nb, nl).mean()) * n, float(opsd[:3].sum()) / s * n,
                                                  noted, model_bytes, own=own["distance"])
        sub["distance"]["box_tests_per_s"] = own["distance"]["nbv_sum"] / (kms["distance"] * 1e-3)
        sub["tolerance"]["cpu_baseline"] = dict(sub["distance"]["cpu_baseline"],
                                                sample="the reference has no tolerance query: a caller runs fcl::distance and compares (same sample)")
        sub["tolerance"]["kernel_ms"] = kms["tolerance"]
        sub["tolerance"]["speedup_over_plain_distance"] = kms["distance"] / kms["tolerance"]
    head = sub["collide"]
    return {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOADS["cfg5"], "headline": "collide() verdicts", "poses_per_gpu": n, "global_poses": total,
                       "tolerance": tol, "seeds": {"mesh1": 21, "mesh2": 22, "poses": 6}, "setup_s": setup_s, "checks": checks,
                       "l2_flush_between_steps": True, "traversal": _capi.get_option("traversal"),
                       "multi_gpu": ("BVHs replicated (built on every GPU), poses partitioned, packed result records all-gathered with "
                                     "NCCL, overlapped with the next step") if world > 1 else "single GPU"},
            "clocks": clocks, "e2e": head.get("e2e"), "gpu_launches": launches, "roofline": head.get("roofline"),
            "cpu_baseline": head.get("cpu_baseline"), "workloads": sub, "peaks": ctx.microbench()}
