#!/bin/bash
# round 2, GPU call 33: front kernel stack capacity as a launch parameter (sweep), front kernel forced on 1M env/rob verdicts,
# ordered contact kernel at 3 / 5 blocks per SM
O=gpurun_out/r02_aj
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
for cap in 0 384 512 768 1024 1536 2048; do run "cfg5 cap=$cap" --workload cfg5 --poses 100000 --opt front_cap=$cap; done
for cap in 0 256 384 768 1024; do run "cfg4 cap=$cap" --workload cfg4 --opt front_cap=$cap; run "cfg1 cap=$cap" --workload cfg1 --opt front_cap=$cap; done
run "collide pooled" --workload collide
run "collide front=2" --workload collide --opt collide_front=2
run "collide front=2 cap=256" --workload collide --opt collide_front=2 --opt front_cap=256
for lib in default ord3 ord5 default ord3; do
  if [ "$lib" = default ]; then unset FCLGPU_LIB_PATH; else export FCLGPU_LIB_PATH=$PWD/fcl_b200/lib/variants/libfclgpu_$lib.so; fi
  run "contacts $lib" --workload contacts
done
