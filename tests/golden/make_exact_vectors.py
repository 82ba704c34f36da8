"""Golden vectors for the leaf / box routines computed INDEPENDENTLY of the oracle, in exact rational arithmetic.

    python tests/golden/make_exact_vectors.py      # writes tests/golden/exact_vectors.npz

Every input is a float64 (exactly representable), every predicate is evaluated with fractions.Fraction, so the golden
answers carry no rounding at all; square roots are taken with mpmath at 60 digits and rounded once to float64.
Nothing here imports oracle/ or fcl_b200/.  What is stored:

  obb_*   : the 15-axis box-box separating-axis test of the reference (math/bv/OBB-inl.h:399-523) as a mathematical
            statement: boxes with half extents a, b, relative rotation B, translation T are reported disjoint iff for
            one of the axes A0, A1, A2, B0, B1, B2, Ai x Bj:  |T . L| > ra + rb  with every |B_ij| padded by 1e-6.
            Stored: verdict, and `margin` = min over the axes of | |T.L| - (ra + rb) | (how far the nearest
            inequality is from flipping; the test skips cases whose margin is within rounding reach).
  tri_*   : triangle-triangle intersection by exact segment-triangle tests (orient3d signs) -- a different algorithm
            from the reference's 17-axis SAT (traversal/collision/intersect-inl.h:727-845); both decide the same
            geometric fact away from degenerate / touching configurations.  Stored: verdict, margin = the smallest
            |orient3d| met, normalised by the cube of the coordinate scale.
  dist_*  : triangle-triangle distance as the minimum over the 6 vertex-triangle and 9 edge-edge exact squared
            distances (0 when the triangles intersect) -- again not the PQP case analysis of
            primitive_shape_algorithm/triangle_distance-inl.h:171-394.  Stored: the distance correctly rounded.
"""
import os
from fractions import Fraction as Fr

import mpmath
import numpy as np

mpmath.mp.dps = 60
HERE = os.path.dirname(os.path.abspath(__file__))


def F(x):
    return Fr(float(x))


def vec(a):
    return [F(a[0]), F(a[1]), F(a[2])]


def sub(a, b):
    return [a[0] - b[0], a[1] - b[1], a[2] - b[2]]


def add(a, b):
    return [a[0] + b[0], a[1] + b[1], a[2] + b[2]]


def mul(a, s):
    return [a[0] * s, a[1] * s, a[2] * s]


def dot(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def cross(a, b):
    return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]


# ---------------------------------------------------------------------------------------------- OBB SAT
def obb_disjoint_exact(B, T, a, b):
    """B[i][j] = a_i . b_j (relative rotation), T = centre of box b in a's frame, a / b = half extents."""
    reps = Fr(1e-6)
    Bf = [[abs(B[i][j]) + reps for j in range(3)] for i in range(3)]
    gaps = []
    for i in range(3):  # A_i
        gaps.append(abs(T[i]) - (a[i] + sum(b[j] * Bf[i][j] for j in range(3))))
    for j in range(3):  # B_j
        s = sum(T[i] * B[i][j] for i in range(3))
        gaps.append(abs(s) - (b[j] + sum(a[i] * Bf[i][j] for i in range(3))))
    for i in range(3):  # A_i x B_j
        i1, i2 = (i + 1) % 3, (i + 2) % 3
        for j in range(3):
            j1, j2 = (j + 1) % 3, (j + 2) % 3
            s = T[i2] * B[i1][j] - T[i1] * B[i2][j]
            ra = a[i1] * Bf[i2][j] + a[i2] * Bf[i1][j]
            rb = b[j1] * Bf[i][j2] + b[j2] * Bf[i][j1]
            gaps.append(abs(s) - (ra + rb))
    return any(g > 0 for g in gaps), min(abs(g) for g in gaps)


# ---------------------------------------------------------------------------------------------- triangle intersection
def orient3d(a, b, c, d):
    return dot(sub(a, d), cross(sub(b, d), sub(c, d)))


def seg_tri_intersect(p, q, t, track):
    """closed segment pq vs closed triangle t (non-coplanar configurations; coplanar ones are flagged degenerate)."""
    dp, dq = orient3d(t[0], t[1], t[2], p), orient3d(t[0], t[1], t[2], q)
    track.extend([abs(dp), abs(dq)])
    if (dp > 0 and dq > 0) or (dp < 0 and dq < 0):
        return False
    if dp == 0 and dq == 0:
        track.append(Fr(0))  # coplanar segment: degenerate, margin 0
        return False
    s = [orient3d(p, q, t[0], t[1]), orient3d(p, q, t[1], t[2]), orient3d(p, q, t[2], t[0])]
    track.extend(abs(x) for x in s)
    return all(x >= 0 for x in s) or all(x <= 0 for x in s)


def tri_intersect_exact(P, Q):
    track = []
    hit = False
    for k in range(3):
        hit |= seg_tri_intersect(P[k], P[(k + 1) % 3], Q, track)
        hit |= seg_tri_intersect(Q[k], Q[(k + 1) % 3], P, track)
    return hit, min(track)


# ---------------------------------------------------------------------------------------------- triangle distance
def point_segment_d2(p, a, b):
    ab = sub(b, a)
    l2 = dot(ab, ab)
    t = dot(sub(p, a), ab) / l2 if l2 != 0 else Fr(0)
    t = min(max(t, Fr(0)), Fr(1))
    d = sub(p, add(a, mul(ab, t)))
    return dot(d, d)


def point_triangle_d2(p, t):
    n = cross(sub(t[1], t[0]), sub(t[2], t[0]))
    nn = dot(n, n)
    best = min(point_segment_d2(p, t[k], t[(k + 1) % 3]) for k in range(3))
    if nn != 0:
        h = dot(sub(p, t[0]), n)
        proj = sub(p, mul(n, h / nn))
        inside = all(dot(cross(sub(t[(k + 1) % 3], t[k]), sub(proj, t[k])), n) >= 0 for k in range(3))
        if inside:
            best = min(best, h * h / nn)
    return best


def segment_segment_d2(p, a, q, b):
    """min over s, t in [0,1] of |(p + s a) - (q + t b)|^2: interior critical point or a clamped 1-D minimum on an edge
    of the parameter square (the function is convex)."""
    cands = [point_segment_d2(p, q, add(q, b)), point_segment_d2(add(p, a), q, add(q, b)),
             point_segment_d2(q, p, add(p, a)), point_segment_d2(add(q, b), p, add(p, a))]
    aa, bb, ab = dot(a, a), dot(b, b), dot(a, b)
    den = aa * bb - ab * ab
    if den != 0:
        w = sub(p, q)
        s = (ab * dot(b, w) - bb * dot(a, w)) / den
        t = (aa * dot(b, w) - ab * dot(a, w)) / den
        if 0 <= s <= 1 and 0 <= t <= 1:
            d = sub(add(p, mul(a, s)), add(q, mul(b, t)))
            cands.append(dot(d, d))
    return min(cands)


def tri_distance2_exact(P, Q):
    hit, _ = tri_intersect_exact(P, Q)
    if hit:
        return Fr(0)
    c = [point_triangle_d2(P[k], Q) for k in range(3)] + [point_triangle_d2(Q[k], P) for k in range(3)]
    for i in range(3):
        for j in range(3):
            c.append(segment_segment_d2(P[i], sub(P[(i + 1) % 3], P[i]), Q[j], sub(Q[(j + 1) % 3], Q[j])))
    return min(c)


def fr_sqrt_to_float(x):
    return float(mpmath.sqrt(mpmath.mpf(x.numerator) / mpmath.mpf(x.denominator)))


# ---------------------------------------------------------------------------------------------- case generation
def random_rotation(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def main():
    rng = np.random.default_rng(20261017)
    n = 1000
    # OBB: boxes of size ~1 at centre distances that straddle the touching range, random relative rotation (a float64
    # matrix that is orthonormal only to rounding: the SAT is evaluated on exactly these numbers, like the reference does)
    obb_B = np.stack([random_rotation(rng) for _ in range(n)])
    obb_a = rng.uniform(0.05, 1.0, size=(n, 3))
    obb_b = rng.uniform(0.05, 1.0, size=(n, 3))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    obb_T = d * rng.uniform(0.2, 2.6, size=(n, 1))
    obb_T[: n // 10] *= 1000.0  # scene-scale coordinates too
    obb_a[: n // 10] *= 1000.0
    obb_b[: n // 10] *= 1000.0
    obb_v, obb_m = np.zeros(n, bool), np.zeros(n)
    for i in range(n):
        B = [[F(obb_B[i, r, c]) for c in range(3)] for r in range(3)]
        v, m = obb_disjoint_exact(B, vec(obb_T[i]), vec(obb_a[i]), vec(obb_b[i]))
        obb_v[i], obb_m[i] = v, float(m)

    # triangles: pairs at distances around their size (about half intersect), plus far pairs and scene-scale pairs
    tri_P = rng.normal(size=(n, 3, 3))
    tri_Q = rng.normal(size=(n, 3, 3)) + (rng.normal(size=(n, 1, 3)) * rng.uniform(0.0, 1.0, size=(n, 1, 1)))
    tri_Q[n // 2: n // 2 + n // 10] += 8.0
    tri_P[-n // 10:] = tri_P[-n // 10:] * 800.0 + 2500.0
    tri_Q[-n // 10:] = tri_Q[-n // 10:] * 800.0 + 2500.0
    tri_v, tri_m, dist = np.zeros(n, bool), np.zeros(n), np.zeros(n)
    for i in range(n):
        P, Q = [vec(p) for p in tri_P[i]], [vec(q) for q in tri_Q[i]]
        v, m = tri_intersect_exact(P, Q)
        scale = max(np.abs(tri_P[i]).max(), np.abs(tri_Q[i]).max())
        tri_v[i], tri_m[i] = v, float(m) / scale ** 3
        dist[i] = fr_sqrt_to_float(tri_distance2_exact(P, Q))
    out = os.path.join(HERE, "exact_vectors.npz")
    np.savez_compressed(out, obb_B=obb_B, obb_T=obb_T, obb_a=obb_a, obb_b=obb_b, obb_disjoint=obb_v, obb_margin=obb_m,
                        tri_P=tri_P, tri_Q=tri_Q, tri_intersect=tri_v, tri_margin=tri_m, tri_distance=dist)
    print("wrote", out, "| obb disjoint", int(obb_v.sum()), "of", n, "| triangles intersecting", int(tri_v.sum()), "of", n,
          "| zero distances", int((dist == 0).sum()))


if __name__ == "__main__":
    main()
