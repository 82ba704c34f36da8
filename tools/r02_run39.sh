#!/bin/bash
# round 2, GPU call 39: ordered contact kernel schedule variants (seed levels, leaf-round trigger, expansions per round)
O=gpurun_out/r02_ap
mkdir -p $O
run() {  # label, bench args...
  local label=$1; shift
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e "$@" 2>> $O/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
w=d.get('workloads')
print('%-28s' % '$label', {k: round(v['ms_per_step'],4) for k,v in w.items()} if w else round(d['ms_per_step'],4))"
}
for lib in default ordseed4 ordseed3 ordtrig24 ordtrig16 ordnexp12 ordnexp8 default; do
  if [ "$lib" = default ]; then unset FCLGPU_LIB_PATH; else export FCLGPU_LIB_PATH=$PWD/fcl_b200/lib/variants/libfclgpu_$lib.so; fi
  run "contacts $lib" --workload contacts
done
