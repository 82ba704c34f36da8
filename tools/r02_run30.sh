#!/bin/bash
# round 2, GPU call 30: early first exact round as the default -- distance parity, cfg5 A/B
O=gpurun_out/r02_ag
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_tolerance.py tests/test_gpu_large.py -m gpu -x -q -k "distance or tolerance or both_objects or edge or tiny or large or unprunable or 1m or upload or overflow" > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
for lib in default eager0 default eager0; do
  if [ "$lib" = default ]; then unset FCLGPU_LIB_PATH; else export FCLGPU_LIB_PATH=$PWD/fcl_b200/lib/variants/libfclgpu_$lib.so; fi
  timeout 600 python bench.py --steps 5 --warmup 3 --workload cfg5 --poses 100000 --no-cpu-baseline --no-e2e 2> $O/cfg5_${lib}.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('%-8s cfg5' % '$lib', {k: round(v['ms_per_step'],3) for k,v in d['workloads'].items()})"
done
