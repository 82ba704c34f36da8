#!/bin/bash
# round 2, GPU call 25: seed front for the counts-only front kernel -- parity + A/B (cfg1 at 10k, cfg4, cfg5)
O=gpurun_out/r02_ab
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py tests/test_broadphase.py tests/test_continuous.py -m gpu -x -q -k "cfg1 or counts_only or tiny or large or edge or synthetic or full_size or broadphase or continuous or binary or pinned" > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest.log
for lib in default frontseed0 default frontseed0; do
  if [ "$lib" = default ]; then unset FCLGPU_LIB_PATH; else export FCLGPU_LIB_PATH=$PWD/fcl_b200/lib/variants/libfclgpu_$lib.so; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --workload cfg1 --no-cpu-baseline --no-e2e 2> $O/cfg1_${lib}.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('%-10s cfg1 kernel_ms %.4f' % ('$lib', d['roofline']['kernel_ms']))"
  timeout 600 python bench.py --steps 5 --warmup 3 --workload cfg4 --poses 250000 --no-cpu-baseline --no-e2e 2> $O/cfg4_${lib}.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('%-10s cfg4 ms %.3f' % ('$lib', d['ms_per_step']))"
  timeout 600 python bench.py --steps 5 --warmup 3 --workload cfg5 --poses 100000 --no-cpu-baseline --no-e2e 2> $O/cfg5_${lib}.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('%-10s cfg5' % '$lib', {k: round(v['ms_per_step'],3) for k,v in d['workloads'].items()})"
done
