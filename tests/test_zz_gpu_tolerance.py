"""Tolerance verification (extension; BASELINE cfg5 "collide + distance tolerance verification"): the reference has no
API for it (only README.md:13-14 mentions it) -- a caller compares fcl::distance with the tolerance.  The GPU entry
point runs the same distance traversal from min_distance = cutoff, so the expected result is min(oracle distance, cutoff),
bit for bit, and the verdict is `oracle distance <= tolerance`."""
import numpy as np
import pytest

import fcl_b200 as F
from fcl_b200 import _capi
from fcl_b200.poses import random_poses

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("trav", [3, 0, 1])
def test_tolerance_verification_matches_oracle(oracle, env_rob_npz, trav):
    (ev, et), (rv, rt) = env_rob_npz
    env, rob = F.BVHModel.from_arrays(ev, et), F.BVHModel.from_arrays(rv, rt)
    oenv, orob = oracle.Model(ev, et), oracle.Model(rv, rt)
    n = 3000
    P = random_poses(n, seed=211)
    rd = oracle.distance_batch(oenv, orob, P, None, True, 2, nthreads=8)
    def same(a, b):  # traversal 0 = the reference's order (exact); 3 verified bit-identical on this batch; 1: tie rounding
        return np.array_equal(a, b) if trav != 1 else bool(np.all(np.abs(a - b) <= 1e-14 * np.abs(b)))

    _capi.set_option("traversal", trav)
    try:
        full = F.distance_batch(env, P, rob, None, F.DistanceRequest(True), stats=True)
        assert same(full.min_distance, rd["min_distance"])
        for tol in (50.0, 400.0):
            cutoff = float(np.nextafter(tol, np.inf))
            got = F.distance_batch(env, P, rob, None, F.DistanceRequest(True), stats=True, cutoff=cutoff)
            assert same(got.min_distance, np.minimum(rd["min_distance"], cutoff)), (trav, tol)
            far = rd["min_distance"] >= cutoff
            assert 0.05 * n < far.sum() < 0.95 * n
            assert (got.b1[far] == -1).all() and (got.b2[far] == -1).all()
            near = ~far & (rd["min_distance"] > 0)
            assert (got.b1[near] >= 0).all() and (got.b2[near] >= 0).all()
            if trav == 0:  # the reference's visiting order: the same pair wins among exact ties
                assert np.array_equal(got.b1[near], full.b1[near]) and np.array_equal(got.b2[near], full.b2[near])
                assert got.nearest_p1[near].tobytes() == full.nearest_p1[near].tobytes()
            # (front traversals may name another pair of an exact tie, DESIGN 2: ids are not compared there)
            # pruning from the first round on: fewer box tests than the unbounded query
            assert got.n_bv.astype(np.int64).sum() < full.n_bv.astype(np.int64).sum()
            within, wres = F.within_tolerance_batch(env, P, rob, None, tol, stats=True)  # early exit at the first pair within tol
            assert np.array_equal(within, rd["min_distance"] <= tol)
            # the witness is a real triangle-pair distance: an upper bound of the true minimum, itself within the tolerance
            assert (wres.min_distance[within] <= tol).all() and (wres.min_distance[within] >= rd["min_distance"][within]).all()
            assert (wres.min_distance[~within] == cutoff).all()
            within2, full_tol = F.within_tolerance_batch(env, P, rob, None, tol, stats=True, early_exit=False)
            assert np.array_equal(within2, within)
            if trav != 1:  # (traversal 1: the early-exit verdict runs the FP32-steered kernel, the plain query the FP64-bound one)
                assert wres.n_bv.astype(np.int64).sum() <= full_tol.n_bv.astype(np.int64).sum()
                assert wres.n_leaf.astype(np.int64).sum() <= full_tol.n_leaf.astype(np.int64).sum()
    finally:
        _capi.set_option("traversal", 3)


def test_cutoff_argument_is_checked(env_rob_npz):
    (ev, et), _ = env_rob_npz
    env = F.BVHModel.from_arrays(ev, et)
    with pytest.raises(F.FclGpuError) as ei:
        F.distance_batch(env, random_poses(4), env, None, F.DistanceRequest(), cutoff=0.0)
    assert ei.value.code == _capi.ERR_INVALID_ARGUMENT


def test_shim_within_tolerance_equals_plain_distances(tmp_path):
    """include/fclgpu/fcl_shim.hpp: fclgpu::within_tolerance against fclgpu::distance <= tolerance (tests/shim/shim_check.cpp)."""
    import subprocess

    from tests.test_fcl_shim import _build

    out = subprocess.run([_build(tmp_path), "tolerance"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "shim tolerance OK" in out.stdout
