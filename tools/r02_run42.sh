#!/bin/bash
# round 2, GPU call 42: cfg5 on ONE GPU with the pose shards the other ranks get at N > 1 (is the N > 1 step the slowest shard?)
O=gpurun_out/r02_ar
mkdir -p $O
for shift in 0 1 2 5; do
  FCLGPU_BENCH_SHARD_SHIFT=$shift timeout 600 python bench.py --workload cfg5 --poses 100000 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>> $O/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('shard %s' % '$shift', {k: round(v['ms_per_step'],3) for k,v in d['workloads'].items()})"
done
timeout 600 python bench.py --workload cfg5 --poses 400000 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>> $O/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('400k poses', {k: round(v['ms_per_step'],3) for k,v in d['workloads'].items()})"
