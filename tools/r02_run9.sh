#!/bin/bash
# round 2, GPU call 9: distance kernel -- seed front A/B, per-phase cycle profile, parity of the seed variants
O=gpurun_out/r02_i
mkdir -p $O
V=$PWD/fcl_b200/lib/variants
for lib in seed5; do
  FCLGPU_LIB_PATH=$V/libfclgpu_$lib.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_tolerance.py -m gpu -x -q -k "distance or tolerance or both_objects or edge or upload" > $O/pytest_$lib.log 2>&1; echo "pytest $lib rc=$?"; tail -3 $O/pytest_$lib.log
done
for lib in default seed4 seed5; do
  if [ "$lib" = default ]; then unset FCLGPU_LIB_PATH; else export FCLGPU_LIB_PATH=$V/libfclgpu_$lib.so; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --workload distance --no-cpu-baseline --no-e2e 2> $O/ab_${lib}.err | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('%-8s value %.4g q/s  kernel_ms %.3f' % ('$lib', d['value'], d['roofline']['kernel_ms']))
except Exception as e: print('$lib FAILED', e)"
done
FCLGPU_LIB_PATH=$V/libfclgpu_prof.so timeout 300 python tools/dist_phase_profile.py 2>&1 | tee $O/phase_profile.log
unset FCLGPU_LIB_PATH
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -3 $O/pytest_gpu.log
