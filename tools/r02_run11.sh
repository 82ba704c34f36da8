#!/bin/bash
# round 2, GPU call 11: bottom-up refit on the device
O=gpurun_out/r02_k
mkdir -p $O
timeout 900 python -m pytest tests/test_bottomup_refit.py tests/test_gpu_parity.py -m gpu -x -q -k "bottomup or refit or device_build or separate" > $O/pytest_refit.log 2>&1; echo "pytest rc=$?"; tail -25 $O/pytest_refit.log
python - <<'PY'
import time, numpy as np, torch
import fcl_b200 as F
from tests.meshes import heightfield
for n in (33, 100, 316):
    v, t = heightfield(n, size=10.0, seed=3, amp=0.6)
    m = F.BVHModel.from_arrays(v, t, build_on_device=True)
    m.device_model()
    dv = torch.from_numpy(v + 0.01).cuda()
    for bottomup in (False, True):
        for _ in range(2):
            m.refit_device(dv, bottomup=bottomup)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            m.refit_device(dv, bottomup=bottomup)
        torch.cuda.synchronize()
        print("tris %7d  %s refit %.3f ms" % (len(t), "bottom-up" if bottomup else "top-down ", (time.perf_counter() - t0) / 5 * 1e3))
PY
