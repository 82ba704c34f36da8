#!/bin/bash
# like ab_bench.sh with the default traversal: tools/ab_bench3.sh "<libs>" "<workloads>"
libs="$1"; wls="$2"
for lib in $libs; do for w in $wls; do
  if [ "$lib" = default ]; then unset FCLGPU_LIB_PATH; else export FCLGPU_LIB_PATH=$PWD/fcl_b200/lib/variants/libfclgpu_$lib.so; fi
  timeout 300 python bench.py --steps 3 --warmup 3 --workload $w --no-cpu-baseline --no-e2e 2> gpurun_out/ab3_${lib}_$w.err | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('%-8s %-9s value %.4g q/s  kernel_ms %.3f' % ('$lib', '$w', d['value'], d['roofline']['kernel_ms']))
except Exception as e: print('$lib $w FAILED', e)"
done; done
