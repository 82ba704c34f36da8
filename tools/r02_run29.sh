#!/bin/bash
# round 2, GPU call 29: early first exact round in the distance kernel -- parity + A/B
O=gpurun_out/r02_af
mkdir -p $O
for lib in eager8; do
  FCLGPU_LIB_PATH=$PWD/fcl_b200/lib/variants/libfclgpu_$lib.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_tolerance.py tests/test_gpu_large.py -m gpu -x -q -k "distance or tolerance or both_objects or edge or tiny or large or unprunable or 1m or upload" > $O/pytest_$lib.log 2>&1; echo "pytest $lib rc=$?"; tail -3 $O/pytest_$lib.log
done
for lib in default eager8 eager16 default eager8 eager16; do
  if [ "$lib" = default ]; then unset FCLGPU_LIB_PATH; else export FCLGPU_LIB_PATH=$PWD/fcl_b200/lib/variants/libfclgpu_$lib.so; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --workload distance --no-cpu-baseline --no-e2e 2> $O/ab_${lib}.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('%-8s value %.4g q/s  kernel_ms %.3f' % ('$lib', d['value'], d['roofline']['kernel_ms']))"
done
