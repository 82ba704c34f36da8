#!/bin/bash
# round 2, GPU call 10: seed front as the default -- full GPU suite, default bench line, ncu launch list + full capture of the distance kernel
O=gpurun_out/r02_j
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -3 $O/pytest_gpu.log
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"; tail -3 $O/bench_default.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r02_j/bench_default.json"))
    print("headline value %.4g e2e %.4g frac %.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"]))
    for k,w in d["workloads"].items():
        r=w["roofline"]
        print("%-9s value %.4g e2e %.4g kernel_ms %.3f bound %s frac %.3f cpu %.4g launches %s" % (k, w["value"], w["e2e"]["value"], r["kernel_ms"], r["bound"], r["frac"], w["cpu_baseline"]["value"], w["gpu_launches"]))
except Exception as e:
    print("bench parse failed", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"distance_warp_kernel" -c 1 -f -o $O/full_distance \
    python tools/profile_run.py --workload distance --poses 1000000 --traversal 3 --launches 1 > $O/full_distance.log 2>&1
python tools/ncu_summary.py $O/full_distance.ncu-rep > $O/full_distance.summary.txt 2>&1
python tools/ncu_by_function.py $O/full_distance.ncu-rep >> $O/full_distance.summary.txt 2>&1
head -30 $O/full_distance.summary.txt
