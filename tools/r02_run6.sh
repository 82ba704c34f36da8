#!/bin/bash
# round 2, GPU call 6: full GPU suite (front kernel with self-contained entries, mesh-plane, tolerance), own counters of the
# contact kernel, cfg1 / cfg4 / cfg5 after the front-kernel change
O=gpurun_out/r02_f
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
timeout 300 python tools/own_counters.py 2>&1 | tee $O/own_counters.log
timeout 300 python bench.py --steps 5 --warmup 3 --workload cfg1 --no-cpu-baseline --no-e2e > $O/cfg1.json 2> $O/cfg1.err
python -c "
import json; d=json.load(open('$O/cfg1.json')); print('cfg1 (10k): value %.4g kernel_ms %.4f' % (d['value'], d['roofline']['kernel_ms']))"
timeout 900 python bench.py --workload cfg4 --poses 250000 --steps 3 --warmup 3 --cpu-sample 500 > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "cfg4 rc=$?"
timeout 900 python bench.py --workload cfg5 --poses 100000 --steps 3 --warmup 3 --cpu-sample 300 > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "cfg5 rc=$?"; tail -2 $O/bench_cfg5.err
python - <<'PY'
import json
for w in ("cfg4","cfg5"):
    try:
        d=json.load(open("gpurun_out/r02_f/bench_%s.json"%w))
        print(w, "value %.4g ms %.3f e2e %s" % (d["value"], d["ms_per_step"], d["e2e"] and "%.4g"%d["e2e"]["value"]), "cpu", d["cpu_baseline"] and (round(d["cpu_baseline"]["value"]), d["cpu_baseline"]["matches_gpu"]))
        for k,v in (d.get("workloads") or {}).items():
            print("   ", k, "value %.4g ms %.3f" % (v["value"], v["ms_per_step"]), "e2e %.4g" % v["e2e"]["value"] if "e2e" in v else "", "cpu", v.get("cpu_baseline") and (round(v["cpu_baseline"]["value"]), v["cpu_baseline"]["matches_gpu"]), v.get("speedup_over_plain_distance"))
        print("   checks", d["config"].get("checks"))
    except Exception as e:
        print(w, "parse failed", e)
PY
