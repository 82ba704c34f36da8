#!/bin/bash
# round 2, GPU call 26: verdict collide at 1M poses -- pooled lane-per-query kernel vs the seeded front kernel
O=gpurun_out/r02_ac
mkdir -p $O
for v in 1 2 1 2; do
  timeout 300 python bench.py --steps 10 --warmup 3 --workload collide --no-cpu-baseline --no-e2e --opt collide_front=$v 2> $O/ab_$v.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('collide_front=$v value %.4g q/s kernel_ms %.3f' % (d['value'], d['roofline']['kernel_ms']))"
done
