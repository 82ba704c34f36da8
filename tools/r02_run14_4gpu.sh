#!/bin/bash
# round 2, GPU call 14 (4 GPUs): scaling lines for the default line, cfg4 and cfg5 (weak and strong)
N=${1:-4}
O=gpurun_out/r02_n$N
mkdir -p $O
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) bench.py --gpus $1 "${@:2}"; }
run $N --workload all --steps 5 --no-cpu-baseline > $O/all.json 2> $O/all.err; echo "all rc=$?"; tail -2 $O/all.err
run $N --workload distance --steps 10 --no-cpu-baseline --scaling strong > $O/distance_strong.json 2> $O/distance_strong.err; echo "strong rc=$?"
run $N --workload cfg5 --poses 100000 --steps 3 --no-cpu-baseline > $O/cfg5.json 2> $O/cfg5.err; echo "cfg5 rc=$?"; tail -2 $O/cfg5.err
run $N --workload cfg4 --poses 200000 --steps 3 --no-cpu-baseline > $O/cfg4.json 2> $O/cfg4.err; echo "cfg4 rc=$?"; tail -2 $O/cfg4.err
python - <<PY
import json
for f in ("all","distance_strong","cfg5","cfg4"):
    try:
        d=json.load(open("$O/%s.json"%f))
        print(f, "n_gpus", d["n_gpus"], "value %.4g ms %.3f e2e %s scaling %s" % (d["value"], d["ms_per_step"], d["e2e"] and "%.4g"%d["e2e"]["value"], d["scaling"]))
        for k,v in (d.get("workloads") or {}).items():
            if v: print("    ", k, "value %.4g ms %.3f" % (v["value"], v["ms_per_step"]))
    except Exception as e:
        print(f, "parse failed", e)
PY
