#!/bin/bash
# round 2, GPU call 27: seed front also for distance queries with a finite cutoff (tolerance verdicts) -- parity + A/B
O=gpurun_out/r02_ad
mkdir -p $O
timeout 900 python -m pytest tests/test_zz_gpu_tolerance.py tests/test_gpu_large.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
for lib in default seedcut0 default seedcut0; do
  if [ "$lib" = default ]; then unset FCLGPU_LIB_PATH; else export FCLGPU_LIB_PATH=$PWD/fcl_b200/lib/variants/libfclgpu_$lib.so; fi
  timeout 600 python bench.py --steps 5 --warmup 3 --workload cfg5 --poses 100000 --no-cpu-baseline --no-e2e 2> $O/cfg5_${lib}.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('%-10s cfg5' % '$lib', {k: round(v['ms_per_step'],3) for k,v in d['workloads'].items()})"
done
unset FCLGPU_LIB_PATH
python - <<'PY'
import time, numpy as np, torch
import fcl_b200 as F
g = "tests/golden"
e, r = np.load(g + "/env.npz"), np.load(g + "/rob.npz")
env, rob = F.BVHModel.from_arrays(e["verts"], e["tris"]), F.BVHModel.from_arrays(r["verts"], r["tris"])
P = F.random_poses(1000000, seed=1)
for tol in (50.0, 400.0):
    F.within_tolerance_batch(env, P[:1000], rob, None, tol)
    t0 = time.perf_counter(); w, _ = F.within_tolerance_batch(env, P, rob, None, tol); dt = time.perf_counter() - t0
    print("env/rob tolerance %.0f: %.1f ms end to end, within %.3f" % (tol, dt * 1e3, w.mean()))
PY
