#!/bin/bash
# round 2, GPU call 36: front kernel -- earlier leaf rounds for queries about to saturate (verdicts), fewer expansions per round
O=gpurun_out/r02_am
mkdir -p $O
run() {  # label, bench args...
  local label=$1; shift
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e "$@" 2>> $O/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
w=d.get('workloads')
print('%-28s' % '$label', {k: round(v['ms_per_step'],4) for k,v in w.items()} if w else round(d['ms_per_step'],4))"
}
for trig in 32 16 8 4 1; do
  run "trig=$trig cfg4" --workload cfg4 --opt front_leaf_trigger=$trig; run "trig=$trig cfg1" --workload cfg1 --opt front_leaf_trigger=$trig
  run "trig=$trig cfg5" --workload cfg5 --poses 100000 --opt front_leaf_trigger=$trig
done
for lib in nexp8 nexp12; do
  export FCLGPU_LIB_PATH=$PWD/fcl_b200/lib/variants/libfclgpu_$lib.so
  run "$lib cfg4" --workload cfg4; run "$lib cfg1" --workload cfg1
  run "$lib cfg5" --workload cfg5 --poses 100000
  run "$lib trig=8 cfg4" --workload cfg4 --opt front_leaf_trigger=8
done
