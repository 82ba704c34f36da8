#!/bin/bash
# A/B: run bench.py (kernel-only numbers) for each library variant and workload.
# usage: tools/ab_bench.sh "<lib1> <lib2> ..." "<workloads>" [extra bench args]
libs="$1"; wls="$2"; shift 2
for lib in $libs; do
  for w in $wls; do
    if [ "$lib" = default ]; then unset FCLGPU_LIB_PATH; else export FCLGPU_LIB_PATH=$PWD/fcl_b200/lib/variants/libfclgpu_$lib.so; fi
    timeout 300 python bench.py --steps 3 --warmup 3 --workload $w --traversal ${TRAV:-3} --no-cpu-baseline --no-e2e "$@" > gpurun_out/ab_${lib}_$w.json 2> gpurun_out/ab_${lib}_$w.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_${lib}_$w.json"))
    print("%-8s %-9s value %.4g q/s  kernel_ms %.3f" % ("$lib", "$w", d["value"], d["roofline"]["kernel_ms"]))
except Exception as e:
    print("$lib $w FAILED", e); print(open("gpurun_out/ab_${lib}_$w.err").read()[-600:])
PY
  done
done
