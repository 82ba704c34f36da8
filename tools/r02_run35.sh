#!/bin/bash
# round 2, GPU call 35: front kernel wide rounds that expand every popped pair (both children per lane); distance kernel
# occupancy on cfg5
O=gpurun_out/r02_al
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "front or count or cfg4 or cfg5 or large or small or tiny or edge or verdict or collide" > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
run() {  # label, bench args...
  local label=$1; shift
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e "$@" 2>> $O/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
w=d.get('workloads')
print('%-28s' % '$label', {k: round(v['ms_per_step'],4) for k,v in w.items()} if w else round(d['ms_per_step'],4))"
}
for lib in default nodual dual8 dual24 dual16b5; do
  if [ "$lib" = default ]; then unset FCLGPU_LIB_PATH; else export FCLGPU_LIB_PATH=$PWD/fcl_b200/lib/variants/libfclgpu_$lib.so; fi
  run "$lib cfg4" --workload cfg4; run "$lib cfg1" --workload cfg1
  run "$lib cfg5" --workload cfg5 --poses 100000
done
export FCLGPU_LIB_PATH=$PWD/fcl_b200/lib/variants/libfclgpu_dual16b5.so
run "dual16b5 cfg4 cap=256" --workload cfg4 --opt front_cap=256
for lib in dist4 dist6; do
  export FCLGPU_LIB_PATH=$PWD/fcl_b200/lib/variants/libfclgpu_$lib.so
  run "$lib cfg5" --workload cfg5 --poses 100000
done
