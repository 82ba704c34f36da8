"""CPU-only: the association order of the three-term sums (the one thing about the reference's arithmetic that cannot
be inspected here, because Eigen is not on the image) is a single compile-time switch honoured by BOTH the oracle
(oracle/fcl_oracle_vec.hpp `sum3`) and the product's device math (fcl_b200/csrc/sum_order.h `FCL_SUM3`).  This test builds both
with -DFCL_SUM3_ORDER=1 and checks that (a) the switch is live -- some results change in their last bits -- and (b) oracle
and device math still agree bit for bit, i.e. the same set of sums goes through the switch on both sides, so the order can
be flipped and re-verified in one step the day a real libfcl is available."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(shutil.which("nvcc") is None or shutil.which("g++") is None, reason="needs g++ and nvcc")

DP = C.POINTER(C.c_double)


def _p(a):
    return a.ctypes.data_as(DP)


def _build(tmp, order):
    orc = os.path.join(tmp, "liboracle_sum3_%d.so" % order)
    src = [os.path.join(ROOT, "oracle", f) for f in ("fcl_oracle_math.cpp", "fcl_oracle_bvh.cpp", "oracle_capi.cpp")]
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-pthread", "-DFCL_SUM3_ORDER=%d" % order,
                           "-shared", "-o", orc] + src)
    dev = os.path.join(tmp, "libdevmath_sum3_%d.so" % order)
    subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-Xcompiler", "-fPIC,-ffp-contract=off",
                           "-DFCL_SUM3_ORDER=%d" % order, "-shared", "-o", dev,
                           os.path.join(ROOT, "tests", "hostcheck", "devmath_host.cu")])
    O, D = C.CDLL(orc), C.CDLL(dev)
    O.orc_tri_distance.restype = D.hm_tri_distance.restype = C.c_double
    O.orc_rect_distance.restype = D.hm_rect_distance.restype = C.c_double
    return O, D


def _rot(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _run(O, D, n=4000, seed=5):
    """Results of both sides on the same seeded inputs: triangle distance (value + points), triangle intersection with
    contacts, OBB SAT verdicts, rectangle distances."""
    rng = np.random.default_rng(seed)
    out_o, out_d = [], []
    for _ in range(n):
        S = np.ascontiguousarray(rng.uniform(-1, 1, 9) * 37.3)
        T = np.ascontiguousarray(rng.uniform(-1, 1, 9) * 37.3 + rng.uniform(-30, 30, 3).repeat(3).reshape(3, 3).T.ravel())
        for L, fn, acc in ((O, "orc_tri_distance", out_o), (D, "hm_tri_distance", out_d)):
            P, Q = np.zeros(3), np.zeros(3)
            d = getattr(L, fn)(_p(S), _p(T), _p(P), _p(Q))
            acc.append(np.concatenate([[d], P, Q]))
        R, t = _rot(rng), rng.uniform(-20, 20, 3)
        pose = np.ascontiguousarray(np.concatenate([R.ravel(), t]))
        Rm, tv = np.ascontiguousarray(R.ravel()), np.ascontiguousarray(t)
        nc, c6, dep, nrm = C.c_uint32(0), np.zeros(6), C.c_double(0), np.zeros(3)
        h = O.orc_tri_intersect(_p(S), _p(T), _p(Rm), _p(tv), 1, C.byref(nc), _p(c6), C.byref(dep), _p(nrm))
        out_o.append(np.concatenate([[h, nc.value if h else 0, dep.value if h else 0], c6 if h else np.zeros(6), nrm if h else np.zeros(3)]))
        nc, c6, dep, nrm = C.c_uint32(0), np.zeros(6), C.c_double(0), np.zeros(3)
        h = D.hm_tri_intersect(_p(S), _p(T), _p(pose), 1, C.byref(nc), _p(c6), C.byref(dep), _p(nrm))
        assert h >= 0  # the rolled SAT agreed with the unrolled one
        if h and nc.value < 2:
            c6[3:] = 0.0
        o = out_o[-1]
        if o[0] and o[1] < 2:
            o[6:9] = 0.0
        out_d.append(np.concatenate([[h, nc.value if h else 0, dep.value if h else 0], c6 if h else np.zeros(6), nrm if h else np.zeros(3)]))
        a, b = np.ascontiguousarray(rng.uniform(0.1, 9, 3)), np.ascontiguousarray(rng.uniform(0.1, 9, 3))
        out_o.append(np.array([float(O.orc_obb_disjoint(_p(Rm), _p(tv), _p(a), _p(b)))]))
        out_d.append(np.array([float(D.hm_obb_disjoint(_p(Rm), _p(tv), _p(a), _p(b)))]))
        a2, b2 = np.ascontiguousarray(a[:2]), np.ascontiguousarray(b[:2])
        out_o.append(np.array([O.orc_rect_distance(_p(Rm), _p(tv), _p(a2), _p(b2))]))
        out_d.append(np.array([D.hm_rect_distance(_p(Rm), _p(tv), _p(a2), _p(b2))]))
    return np.concatenate(out_o), np.concatenate(out_d)


def test_the_order_switch_is_live_and_both_sides_follow_it(tmp_path):
    res = {}
    for order in (0, 1):
        O, D = _build(str(tmp_path), order)
        o, d = _run(O, D)
        same = (o.view(np.uint64) == d.view(np.uint64)) | (np.isnan(o) & np.isnan(d))
        assert same.all(), "order %d: oracle and device math differ in %d of %d values" % (order, (~same).sum(), same.size)
        res[order] = o
    changed = (res[0].view(np.uint64) != res[1].view(np.uint64)).sum()
    assert changed > 0, "FCL_SUM3_ORDER has no effect"
    rel = np.abs(res[0] - res[1]) / np.maximum(np.abs(res[0]), 1e-300)
    assert np.nanmax(np.where(np.abs(res[0]) > 1e-9, rel, 0.0)) < 1e-9  # last bits only: far inside north_star's 1e-6


def test_the_whole_cpu_parity_suite_passes_with_both_sides_flipped(tmp_path):
    """The product library (host BVH builder, top-down and bottom-up refit, OBJ path), the host build of its device math and
    the oracle (plain and counting builds), all compiled with -DFCL_SUM3_ORDER=1: every CPU test that compares the two sides
    bit for bit must still pass -- the same set of sums goes through the switch on both sides.  (The GPU kernels take the
    macro from the same headers; with the default order their SASS is unchanged.)"""
    import sys

    csrc = os.path.join(ROOT, "fcl_b200", "csrc")
    lib = str(tmp_path / "libfclgpu_sum3_1.so")
    # PTX only (code=compute_100a): the device code of this build is never run here, and skipping ptxas halves the build time
    subprocess.check_call(["nvcc", "-O1", "-std=c++17", "-fmad=false", "-gencode", "arch=compute_100a,code=compute_100a",
                           "-Xcompiler", "-fPIC,-ffp-contract=off", "-Wno-deprecated-gpu-targets", "-DFCL_SUM3_ORDER=1", "-shared", "-o", lib] +
                          [os.path.join(csrc, f) for f in ("fclgpu_api.cu", "bvh_build.cpp", "comm.cpp", "mesh_io.cpp")] + ["-ldl"])
    env = dict(os.environ, FCL_SUM3_ORDER="1", FCLGPU_LIB_PATH=lib)
    files = ["test_host_api.py", "test_bottomup_refit.py", "test_device_math_host.py", "test_mesh_plane.py", "test_mesh_io.py",
             "test_exact_golden.py", "test_oracle_counters.py", "test_oracle_known_answers.py", "test_continuous.py", "test_broadphase.py"]
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "not gpu", "-p", "no:cacheprovider"] +
                       [os.path.join(ROOT, "tests", f) for f in files], env=env, cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
