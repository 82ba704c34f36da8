#!/bin/bash
# round 2, GPU call 13: continuous collision
O=gpurun_out/r02_m
mkdir -p $O
timeout 900 python -m pytest tests/test_continuous.py -m gpu -x -q > $O/pytest_ca.log 2>&1; echo "pytest rc=$?"; tail -25 $O/pytest_ca.log
python - <<'PY'
import time, numpy as np
import fcl_b200 as F
from oracle import pyoracle as O
g = "tests/golden"
e, r = np.load(g + "/env.npz"), np.load(g + "/rob.npz")
env, rob = F.BVHModel.from_arrays(e["verts"], e["tris"]), F.BVHModel.from_arrays(r["verts"], r["tris"])
n = 200000
P0 = F.random_poses(n, seed=1); P1 = P0.copy(); P1[:, 9:] += np.random.default_rng(1).normal(0, 250.0, size=(n, 3))
req = F.ContinuousCollisionRequest(ccd_solver_type=F.CCDC_CONSERVATIVE_ADVANCEMENT)
F.continuous_collide_batch(env, None, None, rob, P0[:1000], P1[:1000], req)
t0 = time.perf_counter(); got = F.continuous_collide_batch(env, None, None, rob, P0, P1, req); t1 = time.perf_counter()
print("GPU: %d queries in %.1f ms (%.3g q/s end to end), hits %.3f, moving hits %.3f, mean traversals %.2f" % (
    n, (t1 - t0) * 1e3, n / (t1 - t0), got.is_collide.mean(), (got.is_collide & (got.time_of_contact > 0)).mean(), got.iterations.mean()))
oenv, orob = O.Model(e["verts"], e["tris"]), O.Model(r["verts"], r["tris"])
k = 20000
ref = O.continuous_collide_translation_batch(oenv, orob, None, None, P0[:k], P1[:k], nthreads=O.hardware_threads())
print("oracle: %d queries in %.1f ms (%.3g q/s on %d threads); equal: %s" % (k, ref["seconds"] * 1e3, k / ref["seconds"], O.hardware_threads(),
      got.time_of_contact[:k].tobytes() == ref["time_of_contact"].tobytes()))
PY
