#!/bin/bash
# round 2, GPU call 34: counts-only front kernel, resident blocks per SM (4 .. 8) x stack capacity
O=gpurun_out/r02_ak
mkdir -p $O
run() {  # label, bench args...
  local label=$1; shift
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e "$@" 2>> $O/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
w=d.get('workloads')
print('%-28s' % '$label', {k: round(v['ms_per_step'],4) for k,v in w.items()} if w else round(d['ms_per_step'],4))"
}
for lib in default front5 front6 front7 front8; do
  if [ "$lib" = default ]; then unset FCLGPU_LIB_PATH; else export FCLGPU_LIB_PATH=$PWD/fcl_b200/lib/variants/libfclgpu_$lib.so; fi
  for cap in 384 256 192; do
    run "$lib cfg4 cap=$cap" --workload cfg4 --opt front_cap=$cap; run "$lib cfg1 cap=$cap" --workload cfg1 --opt front_cap=$cap
  done
  for cap in 384 256; do run "$lib cfg5 cap=$cap" --workload cfg5 --poses 100000 --opt front_cap=$cap; done
done
