"""The counting build of the oracle (oracle/fcl_oracle_counted.cpp) -- SURVEY 8(d): the arithmetic side of the roofline
uses the operations the reference's sequential traversal EXECUTES.  The counting build must be the same algorithm:
identical results and identical n_bv / n_leaf counters, and its per-test operation counts must lie under the survey's
all-early-outs-fail upper bounds."""
import numpy as np
import pytest

from fcl_b200.poses import random_poses
from oracle import pyoracle as O


@pytest.fixture(scope="module")
def pair(env_rob_npz):
    (ev, et), (rv, rt) = env_rob_npz
    return (O.Model(ev, et), O.Model(rv, rt)), (O.CountedModel(ev, et), O.CountedModel(rv, rt))


def test_counting_build_is_the_same_algorithm(pair):
    (env, rob), (cenv, crob) = pair
    P = random_poses(600, seed=3)
    for mx, en in ((1, False), (100, True)):
        ref = O.collide_batch(env, rob, P, None, mx, en, nthreads=4)
        got = O.counted_query("collide", cenv, crob, P, None, mx, en, nthreads=4)
        assert np.array_equal(got["value"].astype(np.int64), ref["counts"])
        assert np.array_equal(got["n_bv"], ref["n_bv"]) and np.array_equal(got["n_leaf"], ref["n_leaf"])
    ref = O.distance_batch(env, rob, P, None, True, 2, nthreads=4)
    got = O.counted_query("distance", cenv, crob, P, nthreads=4)
    assert np.array_equal(got["value"], ref["min_distance"])
    assert np.array_equal(got["n_bv"], ref["n_bv"]) and np.array_equal(got["n_leaf"], ref["n_leaf"])


def test_executed_counts_are_below_the_survey_upper_bounds(pair):
    """SURVEY 8(d) upper bounds (every early-out failing): F_obb ~ 126 + 175, F_tri ~ 63 + 17 x 47,
    F_rss ~ 126 + 15 + 450, F_td ~ 45 + 9 x 85 + 250 (mul + add; the executed figure also counts compares)."""
    _, (cenv, crob) = pair
    P = random_poses(400, seed=5)
    c = O.counted_query("collide", cenv, crob, P, None, 1, False, nthreads=4)
    flops = c["ops"][:, :3].sum(axis=1)  # mul + add + cmp
    # a query that ends at the root pair: pose set-up (63 mul/add) + one OBB test that exits on an early axis
    root_only = c["n_bv"] == 1
    assert root_only.any()
    assert (flops[root_only] >= 63).all() and (flops[root_only] <= 63 + 126 + 175 + 60).all()
    ub = 63 + c["n_bv"] * (126 + 175 + 60) + c["n_leaf"] * (63 + 17 * 60 + 40)
    assert (flops <= ub).all()
    d = O.counted_query("distance", cenv, crob, P, nthreads=4)
    flops = d["ops"][:, :3].sum(axis=1)
    ub = 200 + d["n_bv"] * (126 + 15 + 700) + (d["n_leaf"] + 1) * (45 + 9 * 140 + 400)
    assert (flops <= ub).all() and (flops > 63).all()
    assert (d["ops"][:, 4] >= 1).all()  # at least the seed pair's sqrt
