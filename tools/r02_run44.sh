#!/bin/bash
# round 2, GPU call 44: tolerance verdicts end to end with pinned result buffers + the tolerance tests
O=gpurun_out/r02_at
mkdir -p $O
timeout 300 python -m pytest tests -m gpu -x -q -k "tolerance" > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
timeout 300 python bench.py --workload cfg5 --poses 500000 --steps 3 --warmup 2 --no-cpu-baseline 2>> $O/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for k,v in d['workloads'].items(): print(k, 'device %.4g  e2e %.4g' % (v['value'], v['e2e']['value']))"
