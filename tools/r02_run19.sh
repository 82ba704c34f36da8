#!/bin/bash
# round 2, GPU call 19: default bench line with the cfg4 / cfg5 legs (wall time), full GPU suite
O=gpurun_out/r02_s
mkdir -p $O
SECONDS=0; timeout 1200 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$? wall ${SECONDS}s"; grep -E "Elapsed|Maximum resident" $O/bench_time.txt; tail -3 $O/bench_default.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_s/bench_default.json"))
print("headline value %.4g e2e %.4g frac %.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"]))
for k,w in d["workloads"].items():
    r=w.get("roofline") or {}
    print("%-9s value %.4g e2e %s ms %.3f bound %s frac %s" % (k, w["value"], (w.get("e2e") or {}).get("value"), w["ms_per_step"], r.get("bound"), r.get("frac")))
    for kk,ww in (w.get("workloads") or {}).items():
        rr=ww.get("roofline") or {}
        print("     %-9s value %.4g ms %.3f bound %s frac %s basis %s" % (kk, ww["value"], ww["ms_per_step"], rr.get("bound"), rr.get("frac"), rr.get("basis")))
PY
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -3 $O/pytest_gpu.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"; cut -c1-400 $O/bench_reference.json
