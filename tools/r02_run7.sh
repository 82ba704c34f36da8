#!/bin/bash
# round 2, GPU call 7: new GPU tests (broadphase, mesh-plane), full default bench line
O=gpurun_out/r02_g
mkdir -p $O
timeout 600 python -m pytest tests/test_broadphase.py tests/test_mesh_plane.py tests/test_zz_gpu_tolerance.py -m gpu -x -q > $O/pytest_new.log 2>&1; echo "pytest rc=$?"; tail -12 $O/pytest_new.log
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"; tail -3 $O/bench_default.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r02_g/bench_default.json"))
    for k,w in d["workloads"].items():
        r=w["roofline"]
        print("%-9s value %.4g e2e %.4g kernel_ms %.3f bound %s frac %.3f cpu %.4g match %s launches %s" % (k, w["value"], w["e2e"]["value"], r["kernel_ms"], r["bound"], r["frac"], w["cpu_baseline"]["value"], w["cpu_baseline"]["matches_gpu"], w["gpu_launches"]))
except Exception as e:
    print("bench parse failed", e)
PY
