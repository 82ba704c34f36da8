#!/bin/bash
# round 2, GPU call 2: ncu --set full of the ordered-front contact kernel and the distance kernel; cfg4 / cfg5 bench
# workloads on one GPU (reduced batches: this is a functional check of bench_big.py); small-batch crossover; shim test
O=gpurun_out/r02_b
mkdir -p $O
timeout 300 python -m pytest tests/test_fcl_shim.py tests/test_c_abi.py -m gpu -x -q > $O/pytest_shim.log 2>&1; echo "shim pytest rc=$?"; tail -3 $O/pytest_shim.log
for w in contacts distance; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"distance_warp_kernel|collide_ordered_kernel" \
      -c 1 -f -o $O/full_$w python tools/profile_run.py --workload $w --poses 1000000 --traversal 3 --launches 1 > $O/full_$w.log 2>&1
  python tools/ncu_summary.py $O/full_$w.ncu-rep > $O/full_$w.summary.txt 2>&1
  python tools/ncu_by_function.py $O/full_$w.ncu-rep >> $O/full_$w.summary.txt 2>&1
  head -40 $O/full_$w.summary.txt
done
for n in 25000 50000 100000 200000 400000; do for f in 0 2; do
  timeout 120 python bench.py --steps 5 --warmup 3 --workload collide --poses $n --no-cpu-baseline --no-e2e --opt collide_front=$f > $O/cross_${n}_$f.json 2> $O/cross_${n}_$f.err
  python -c "
import json; d=json.load(open('$O/cross_${n}_$f.json')); print('collide n=$n front=$f kernel_ms %.4f value %.4g' % (d['roofline']['kernel_ms'], d['value']))"
done; done
timeout 900 python bench.py --workload cfg4 --poses 250000 --steps 3 --warmup 3 --cpu-sample 500 > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "cfg4 rc=$?"; tail -2 $O/bench_cfg4.err
timeout 900 python bench.py --workload cfg5 --poses 100000 --steps 3 --warmup 3 --cpu-sample 300 > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "cfg5 rc=$?"; tail -2 $O/bench_cfg5.err
python - <<'PY'
import json
for w in ("cfg4","cfg5"):
    try:
        d=json.load(open("gpurun_out/r02_b/bench_%s.json"%w))
        print(w, "value %.4g ms %.3f e2e %s" % (d["value"], d["ms_per_step"], d["e2e"] and "%.4g"%d["e2e"]["value"]), "roofline", d["roofline"] and (d["roofline"]["bound"], round(d["roofline"]["frac"],3), round(d["roofline"]["kernel_ms"],3)), "cpu", d["cpu_baseline"] and (round(d["cpu_baseline"]["value"]), d["cpu_baseline"]["matches_gpu"]))
        print("   config", {k:v for k,v in d["config"].items() if k in ("configurations_per_s","colliding_configurations_frac","checks","setup_s")})
        for k,v in (d.get("workloads") or {}).items():
            print("   ", k, "value %.4g ms %.3f" % (v["value"], v["ms_per_step"]), "e2e %.4g" % v["e2e"]["value"] if "e2e" in v else "", "roof", v.get("roofline") and (v["roofline"]["bound"], round(v["roofline"]["frac"],3)), "cpu", v.get("cpu_baseline") and (round(v["cpu_baseline"]["value"]), v["cpu_baseline"]["matches_gpu"]), v.get("speedup_over_plain_distance"))
    except Exception as e:
        print(w, "parse failed", e)
PY
