#!/bin/bash
# round 2, GPU call 43: two overflow areas for the distance front (launches of the two pipeline streams overlap) -- parity + cfg5 e2e
O=gpurun_out/r02_as
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q -k "overflow or unprunable or tolerance or large or concurrent" > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
timeout 600 python bench.py --workload cfg5 --poses 500000 --steps 3 --warmup 2 --no-cpu-baseline 2>> $O/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for k,v in d['workloads'].items(): print(k, 'device %.4g  e2e %.4g' % (v['value'], v['e2e']['value']))"
