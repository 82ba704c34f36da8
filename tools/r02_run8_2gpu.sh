#!/bin/bash
# round 2, GPU call 8 (2 GPUs): the multi-GPU path through fclgpu_comm_* (raw NCCL behind the C ABI)
O=gpurun_out/r02_h
mkdir -p $O
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) bench.py --gpus $1 "${@:2}"; }
timeout 600 python bench.py --workload distance --steps 10 --no-cpu-baseline > $O/dist_n1.json 2> $O/dist_n1.err
run 2 --workload distance --steps 10 --no-cpu-baseline > $O/dist_n2.json 2> $O/dist_n2.err; echo "dist n2 rc=$?"; tail -2 $O/dist_n2.err
run 2 --workload all --steps 5 --no-cpu-baseline > $O/all_n2.json 2> $O/all_n2.err; echo "all n2 rc=$?"; tail -2 $O/all_n2.err
run 2 --workload distance --steps 10 --no-cpu-baseline --scaling strong > $O/dist_n2_strong.json 2> $O/dist_n2_strong.err; echo "strong rc=$?"
run 2 --workload cfg5 --poses 100000 --steps 3 --no-cpu-baseline > $O/cfg5_n2.json 2> $O/cfg5_n2.err; echo "cfg5 n2 rc=$?"; tail -2 $O/cfg5_n2.err
run 2 --workload cfg4 --poses 200000 --steps 3 --no-cpu-baseline > $O/cfg4_n2.json 2> $O/cfg4_n2.err; echo "cfg4 n2 rc=$?"; tail -2 $O/cfg4_n2.err
python - <<'PY'
import json
for f in ("dist_n1","dist_n2","all_n2","dist_n2_strong","cfg5_n2","cfg4_n2"):
    try:
        d=json.load(open("gpurun_out/r02_h/%s.json"%f))
        print(f, "n_gpus", d["n_gpus"], "value %.4g ms %.3f e2e %s scaling %s" % (d["value"], d["ms_per_step"], d["e2e"] and "%.4g"%d["e2e"]["value"], d["scaling"]))
        for k,v in (d.get("workloads") or {}).items():
            if v: print("    ", k, "value %.4g ms %.3f" % (v["value"], v["ms_per_step"]))
    except Exception as e:
        print(f, "parse failed", e)
PY
