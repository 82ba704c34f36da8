#!/bin/bash
# round 2, GPU call 16: pre-expanded front entries A/B + parity
O=gpurun_out/r02_p
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_tolerance.py tests/test_gpu_large.py -m gpu -x -q -k "distance or tolerance or both_objects or edge or tiny or large or unprunable or 1m or upload" > $O/pytest_dist.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_dist.log
for lib in default prex0 default prex0; do
  if [ "$lib" = default ]; then unset FCLGPU_LIB_PATH; else export FCLGPU_LIB_PATH=$PWD/fcl_b200/lib/variants/libfclgpu_$lib.so; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --workload distance --no-cpu-baseline --no-e2e 2> $O/ab_${lib}.err | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('%-8s value %.4g q/s  kernel_ms %.3f' % ('$lib', d['value'], d['roofline']['kernel_ms']))
except Exception as e: print('$lib FAILED', e)"
done
for lib in default prex0; do
  if [ "$lib" = default ]; then unset FCLGPU_LIB_PATH; else export FCLGPU_LIB_PATH=$PWD/fcl_b200/lib/variants/libfclgpu_$lib.so; fi
  timeout 600 python bench.py --steps 3 --warmup 2 --workload cfg5 --poses 100000 --no-cpu-baseline --no-e2e 2> $O/cfg5_${lib}.err | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$lib cfg5', {k: round(v['ms_per_step'],3) for k,v in d['workloads'].items()})
except Exception as e: print('$lib cfg5 FAILED', e)"
done
