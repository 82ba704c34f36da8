#!/bin/bash
# round 2, GPU call 31: eager first exact round only with the screening instantiation -- check
O=gpurun_out/r02_ah
mkdir -p $O
timeout 300 python bench.py --steps 5 --warmup 3 --workload distance --no-cpu-baseline --no-e2e 2> $O/d.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('distance kernel_ms %.3f' % d['roofline']['kernel_ms'])"
timeout 600 python bench.py --steps 5 --warmup 3 --workload cfg5 --poses 100000 --no-cpu-baseline --no-e2e 2> $O/cfg5.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('cfg5', {k: round(v['ms_per_step'],3) for k,v in d['workloads'].items()})"
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -3 $O/pytest_gpu.log
