#!/bin/bash
# round 2, GPU call 12: median split on the device
O=gpurun_out/r02_l
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py -m gpu -x -q -k "device_build or large or 1m" > $O/pytest_build.log 2>&1; echo "pytest rc=$?"; tail -25 $O/pytest_build.log
python - <<'PY'
import time, numpy as np
import fcl_b200 as F
from tests.meshes import heightfield
for n in (100, 316):
    v, t = heightfield(n, size=10.0, seed=3, amp=0.6)
    for split in (0, 1, 2):
        t0 = time.perf_counter()
        m = F.BVHModel.from_arrays(v, t, split, build_on_device=True); m.device_model(); F.sync_status()
        t1 = time.perf_counter()
        m2 = F.BVHModel.from_arrays(v, t, split, build_on_device=True); m2.device_model(); F.sync_status()
        t2 = time.perf_counter()
        print("tris %7d split %d device build %.1f ms (first %.1f)" % (len(t), split, (t2 - t1) * 1e3, (t1 - t0) * 1e3))
PY
