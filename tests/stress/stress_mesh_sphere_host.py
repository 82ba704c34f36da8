"""Randomised differential run of the mesh <-> sphere distance traversal (csrc/mesh_sphere.cuh, host build of the very
code the kernel inlines) against the oracle, no GPU needed:  python tests/stress/stress_mesh_sphere_host.py [seconds] [seed]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from fcl_b200.poses import random_poses, identity_poses
from oracle import pyoracle as O
from tests import hostcheck as hm
from tests.meshes import box_mesh, heightfield, noisy_sphere, random_soup, uv_sphere
from tests.test_device_math_host import _host_mesh_sphere

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 4321)
t_end = time.time() + budget
rounds = checks = ties = 0
while time.time() < t_end:
    kind = rng.integers(0, 5)
    s = int(rng.integers(1 << 30))
    if kind == 0:
        v, t = random_soup(int(rng.integers(1, 600)), seed=s, scale=float(rng.uniform(0.5, 3)), tri_size=float(rng.uniform(0.05, 1.0)))
    elif kind == 1:
        v, t = noisy_sphere(float(rng.uniform(0.3, 2)), int(rng.integers(4, 30)), int(rng.integers(3, 30)), seed=s, noise=float(rng.uniform(0, 0.2)), scale=tuple(rng.uniform(0.3, 3, size=3)))
    elif kind == 2:
        v, t = heightfield(int(rng.integers(2, 40)), size=float(rng.uniform(1, 6)), seed=s, amp=float(rng.uniform(0, 1)))
    elif kind == 3:
        v, t = box_mesh(*rng.uniform(0.1, 2, size=3))
    else:
        v, t = uv_sphere(float(rng.uniform(0.2, 2)), int(rng.integers(3, 20)), int(rng.integers(2, 20)))
    unit = float(10.0 ** rng.uniform(-3, 4))  # the same scene in millimetres ... tens of kilometres
    v = np.asarray(v) * unit
    split = int(rng.integers(0, 3))
    o = O.Model(v, t, split)
    n = int(rng.integers(50, 1500))
    ext = float(rng.uniform(0.5, 6)) * unit
    M = random_poses(n, seed=int(rng.integers(1 << 30)), extents=(-ext, -ext, -ext, ext, ext, ext)) if rng.integers(0, 2) else identity_poses(n)
    S = random_poses(n, seed=int(rng.integers(1 << 30)), extents=(-ext, -ext, -ext, ext, ext, ext))
    r = float(rng.choice([0.0, rng.uniform(0.01, 2.0)])) * unit
    b32 = bool(rng.integers(0, 2))  # box bound from the FP64 records or from the packed FP32 records
    got = _host_mesh_sphere(hm, o, v, t, r, M, S, b32)
    brute = O.distance_mesh_sphere_batch(o, r, M, S, brute=True, nthreads=8)
    trav = O.distance_mesh_sphere_batch(o, r, M, S, nthreads=8)
    pos = brute["min_distance"] > 0
    ok = np.array_equal(got["min_distance"], brute["min_distance"])
    ok = ok and np.array_equal(trav["min_distance"] < 0, brute["min_distance"] < 0)
    ok = ok and bool(np.all(np.abs(trav["min_distance"][pos] - brute["min_distance"][pos]) <= 1e-12 * brute["min_distance"][pos]))
    ties += int((trav["min_distance"] != brute["min_distance"]).sum())
    same = pos & (got["b1"] == brute["b1"])
    ok = ok and got["p1"][same].tobytes() == brute["p1"][same].tobytes() and got["p2"][same].tobytes() == brute["p2"][same].tobytes()
    rounds += 1
    checks += n
    if not ok:
        print("MISMATCH", dict(kind=int(kind), nt=len(t), unit=unit, split=split, n=n, ext=ext, r=r, b32=b32))
        sys.exit(1)
print("mesh-sphere distance host stress OK: %d random configurations, %d queries; reference-order traversal differing from the all-triangles minimum in the last bits: %d" % (rounds, checks, ties))
