#!/bin/bash
# round 2, GPU call 32: staged (four lanes per record) box-record loads in the counts-only front kernel -- parity + A/B
O=gpurun_out/r02_ai
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "front or count or cfg4 or cfg5 or large or small or tiny or edge or verdict or collide" > $O/pytest_stage.log 2>&1; echo "pytest stage rc=$?"; tail -3 $O/pytest_stage.log
for lib in default nostage nostage320 stage448 default nostage; do
  if [ "$lib" = default ]; then unset FCLGPU_LIB_PATH; else export FCLGPU_LIB_PATH=$PWD/fcl_b200/lib/variants/libfclgpu_$lib.so; fi
  for wl in cfg1 cfg4 "cfg5 --poses 100000"; do
    timeout 600 python bench.py --steps 5 --warmup 3 --workload $wl --no-cpu-baseline --no-e2e 2> $O/ab_${lib}.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
w=d.get('workloads')
print('%-10s %-6s' % ('$lib', '$wl'.split()[0]), {k: round(v['ms_per_step'],4) for k,v in w.items()} if w else round(d['ms_per_step'],4))"
  done
done
