"""Independent pin of the leaf / box routines ("parity is green against the oracle, not against FCL"):
tests/golden/exact_vectors.npz holds verdicts and distances computed in exact rational arithmetic by ALGORITHMS THAT
DIFFER from the reference's (segment-triangle orientation tests instead of the 17-axis SAT; minimum over vertex-face /
edge-edge squared distances instead of PQP's case analysis), from float64 inputs, by tests/golden/make_exact_vectors.py
(committed; it imports neither oracle/ nor fcl_b200/).  Both the oracle and the product's device math (compiled for
the host by tests/hostcheck) must reproduce them: verdicts exactly wherever the exact margin is beyond rounding reach,
distances to a few ulps (tolerance stated below)."""
import os

import numpy as np
import pytest

from tests import hostcheck as H

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "exact_vectors.npz")
IDENT = np.array([1.0, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0])


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_generator_is_reproducible_and_independent():
    src = open(os.path.join(os.path.dirname(GOLD), "make_exact_vectors.py")).read()
    assert "import oracle" not in src and "from oracle" not in src and "fcl_b200" not in src.split('"""', 2)[2]


def test_obb_sat_verdicts_equal_exact_arithmetic(gold, oracle):
    """OBB-inl.h:399-523: every case whose nearest axis inequality is farther than 1e-9 of the scale from flipping must get
    the exact verdict (float64 rounding of a 4-term sum of O(scale) terms is ~1e-15 of the scale: six orders of room)."""
    L = H.lib()
    n = len(gold["obb_disjoint"])
    checked = 0
    for i in range(n):
        B, T, a, b = (np.ascontiguousarray(gold[k][i]) for k in ("obb_B", "obb_T", "obb_a", "obb_b"))
        scale = np.abs(T).max() + a.max() + b.max()
        if gold["obb_margin"][i] <= 1e-9 * scale:
            continue
        checked += 1
        want = bool(gold["obb_disjoint"][i])
        assert oracle.obb_disjoint(B, T, a, b) == want, i
        assert bool(L.hm_obb_disjoint(H.dptr(B.reshape(-1)), H.dptr(T), H.dptr(a), H.dptr(b))) == want, i
    assert checked > 0.98 * n and 0.3 * n < gold["obb_disjoint"].sum() < 0.7 * n


def test_triangle_intersection_verdicts_equal_exact_arithmetic(gold, oracle):
    """intersect-inl.h:727-845 (17-axis SAT) against exact segment-triangle orientation tests: same verdict wherever the
    smallest orientation determinant is more than 1e-9 (relative to scale^3) away from zero."""
    L = H.lib()
    n = len(gold["tri_intersect"])
    checked = hits = 0
    for i in range(n):
        if gold["tri_margin"][i] <= 1e-9:
            continue
        checked += 1
        P, Q = np.ascontiguousarray(gold["tri_P"][i]), np.ascontiguousarray(gold["tri_Q"][i])
        want = bool(gold["tri_intersect"][i])
        hits += want
        assert oracle.tri_intersect(P, Q) == want, i
        nc = H.C.c_uint32(0)
        six, one, three = np.zeros(6), np.zeros(1), np.zeros(3)
        got = L.hm_tri_intersect(H.dptr(P.reshape(-1)), H.dptr(Q.reshape(-1)), H.dptr(IDENT), 0, H.C.byref(nc), H.dptr(six),
                                 H.dptr(one), H.dptr(three))
        assert bool(got) == want, i
    assert checked > 0.97 * n and hits > 100


def test_triangle_distance_within_ulps_of_exact(gold, oracle):
    """triangle_distance-inl.h:171-394 (PQP TriDist) against the exact minimum over vertex-face and edge-edge distances.
    Tolerance: 16 ulp of the distance plus 4 ulp of the coordinate scale (the routine subtracts coordinates of that size
    before it squares), i.e. ~1e-15 relative -- nine orders inside the 1e-6 of BASELINE.json's north_star.  Pairs the exact
    test classifies as intersecting must return exactly 0 unless they are within rounding reach of merely touching."""
    L = H.lib()
    n = len(gold["tri_distance"])
    worst = 0.0
    for i in range(n):
        P, Q = np.ascontiguousarray(gold["tri_P"][i]), np.ascontiguousarray(gold["tri_Q"][i])
        want = float(gold["tri_distance"][i])
        d_or, _, _ = oracle.tri_distance(P, Q)
        p3, q3 = np.zeros(3), np.zeros(3)
        d_dev = L.hm_tri_distance(H.dptr(P.reshape(-1)), H.dptr(Q.reshape(-1)), H.dptr(p3), H.dptr(q3))
        assert d_dev == d_or, i  # product math == oracle, bit for bit
        if want == 0.0:
            if gold["tri_margin"][i] > 1e-9:
                assert d_or == 0.0, i
            continue
        scale = max(np.abs(P).max(), np.abs(Q).max())
        tol = 16 * np.spacing(want) + 4 * np.spacing(scale)
        assert abs(d_or - want) <= tol, (i, d_or, want)
        worst = max(worst, abs(d_or - want) / want)
        # the reported nearest points realise the distance
        assert abs(np.linalg.norm(p3 - q3) - want) <= tol + 8 * np.spacing(scale), i
    assert worst < 1e-13
