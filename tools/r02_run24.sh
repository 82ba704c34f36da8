#!/bin/bash
# round 2, GPU call 24: seed front for the ordered contact kernel -- parity + A/B
O=gpurun_out/r02_aa
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py -m gpu -x -q -k "contacts or exhaustive or binary_mode or pinned or overflow or synthetic or tiny or single_query or compact or both_objects or edge or large or device_build" > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest.log
for lib in default ordseed0 default ordseed0; do
  if [ "$lib" = default ]; then unset FCLGPU_LIB_PATH; else export FCLGPU_LIB_PATH=$PWD/fcl_b200/lib/variants/libfclgpu_$lib.so; fi
  timeout 300 python bench.py --steps 6 --warmup 3 --workload contacts --no-cpu-baseline --no-e2e 2> $O/ab_${lib}.err | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('%-9s value %.4g q/s  kernel_ms %.3f' % ('$lib', d['value'], d['roofline']['kernel_ms']))
except Exception as e: print('$lib FAILED', e)"
done
unset FCLGPU_LIB_PATH
timeout 300 python tools/split_timing.py 2>&1 | tail -6
