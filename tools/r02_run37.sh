#!/bin/bash
# round 2, GPU call 37: pre-expanded entries in the ordered contact kernel -- parity + A/B
O=gpurun_out/r02_an
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q -k "contact or cfg3 or collide or broadphase or compact or stream or large or edge or tiny" > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
run() {  # label, bench args...
  local label=$1; shift
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e "$@" 2>> $O/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
w=d.get('workloads')
print('%-28s' % '$label', {k: round(v['ms_per_step'],4) for k,v in w.items()} if w else round(d['ms_per_step'],4))"
}
for lib in default noprex default noprex; do
  if [ "$lib" = default ]; then unset FCLGPU_LIB_PATH; else export FCLGPU_LIB_PATH=$PWD/fcl_b200/lib/variants/libfclgpu_$lib.so; fi
  run "contacts $lib" --workload contacts
done
