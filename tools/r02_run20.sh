#!/bin/bash
# round 2, GPU call 20: distance kernel occupancy A/B (6 resident blocks per SM at 80 registers)
O=gpurun_out/r02_t
mkdir -p $O
for lib in default mb6 default mb6; do
  if [ "$lib" = default ]; then unset FCLGPU_LIB_PATH; else export FCLGPU_LIB_PATH=$PWD/fcl_b200/lib/variants/libfclgpu_$lib.so; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --workload distance --no-cpu-baseline --no-e2e 2> $O/ab_${lib}.err | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('%-8s value %.4g q/s  kernel_ms %.3f' % ('$lib', d['value'], d['roofline']['kernel_ms']))
except Exception as e: print('$lib FAILED', e)"
done
