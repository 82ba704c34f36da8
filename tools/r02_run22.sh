#!/bin/bash
# round 2, GPU call 22: pooled collide kernel occupancy A/B (verdicts)
O=gpurun_out/r02_x
mkdir -p $O
for lib in default pool5 pool6 default pool5 pool6; do
  if [ "$lib" = default ]; then unset FCLGPU_LIB_PATH; else export FCLGPU_LIB_PATH=$PWD/fcl_b200/lib/variants/libfclgpu_$lib.so; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --workload collide --no-cpu-baseline --no-e2e 2> $O/ab_${lib}.err | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('%-8s value %.4g q/s  kernel_ms %.3f' % ('$lib', d['value'], d['roofline']['kernel_ms']))
except Exception as e: print('$lib FAILED', e)"
done
