#!/bin/bash
# round 2, GPU call 1: parity suite, the default bench line, A/B of the ordered-front contact kernel variants
mkdir -p gpurun_out/r02_a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02_a/smi.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_a/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_a/pytest_gpu.log
tail -5 gpurun_out/r02_a/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r02_a/bench_default.json 2> gpurun_out/r02_a/bench_default.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_a/bench_default.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r02_a/bench_default.json"))
    for k,w in d["workloads"].items():
        r=w["roofline"]
        print("%-9s value %.4g e2e %.4g kernel_ms %.3f bound %s frac %.3f (fp64 %.3f mem %.3f) cpu %.4g match %s launches %s" % (k, w["value"], w["e2e"]["value"], r["kernel_ms"], r["bound"], r["frac"], r["fp64"]["frac"], (r.get("l2") or r.get("hbm"))["frac"], w["cpu_baseline"]["value"], w["cpu_baseline"]["matches_gpu"], w["gpu_launches"]))
    print(d["peaks"])
except Exception as e:
    print("bench parse failed", e)
PY
export TRAV=3
bash tools/ab_bench.sh "default inl mb5 mb3 inl3" "contacts"
bash tools/ab_bench.sh "default" "contacts" --opt contact_order=1
timeout 300 python bench.py --steps 5 --warmup 3 --workload cfg1 --no-cpu-baseline --no-e2e --opt collide_front=2 > gpurun_out/r02_a/cfg1_front.json 2> gpurun_out/r02_a/cfg1_front.err
python -c "
import json; d=json.load(open('gpurun_out/r02_a/cfg1_front.json')); print('cfg1 front kernel: value %.4g kernel_ms %.4f' % (d['value'], d['roofline']['kernel_ms']))"
