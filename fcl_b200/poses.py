"""Synthetic pose batches for the benchmark and the parity tests.

Distribution follows the reference's mesh tests (test/test_fcl_utility.h:312-327,
348-372; extents test/test_fcl_collision.cpp:323): translation uniform in the
extents box, rotation = eulerToMatrix(a, b, c) with a, b, c ~ U[0, 2*pi).  The
reference draws from unseeded libc rand(); here the stream is an explicit
splitmix64 generator so every run is reproducible from (seed, n).

A pose record is 12 float64: R row-major (9) then t (3); p_world = R p + t.
"""
import numpy as np

ENV_EXTENTS = (-3000.0, -3000.0, 0.0, 3000.0, 3000.0, 3000.0)

_GOLDEN = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def splitmix64_uniform(seed, count, offset=0):
    """count doubles in [0,1): draw k uses state seed + (offset+k+1)*golden (mod 2^64)."""
    with np.errstate(over="ignore"):
        k = np.arange(offset + 1, offset + count + 1, dtype=np.uint64)
        z = np.uint64(seed) + k * _GOLDEN
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def euler_to_matrix(a, b, c):
    """eulerToMatrix of test/test_fcl_utility.h:312-327, vectorised; returns (...,3,3)."""
    c1, c2, c3 = np.cos(a), np.cos(b), np.cos(c)
    s1, s2, s3 = np.sin(a), np.sin(b), np.sin(c)
    R = np.empty(np.shape(a) + (3, 3))
    R[..., 0, 0] = c1 * c2
    R[..., 0, 1] = -c2 * s1
    R[..., 0, 2] = s2
    R[..., 1, 0] = c3 * s1 + c1 * s2 * s3
    R[..., 1, 1] = c1 * c3 - s1 * s2 * s3
    R[..., 1, 2] = -c2 * s3
    R[..., 2, 0] = s1 * s3 - c1 * c3 * s2
    R[..., 2, 1] = c3 * s1 * s2 + c1 * s3
    R[..., 2, 2] = c2 * c3
    return R


def random_poses(n, seed=1, extents=ENV_EXTENTS, start=0):
    """(n, 12) float64 pose records; pose i consumes draws 6*(start+i) .. 6*(start+i)+5
    in the order x, y, z, a, b, c, so shards of one logical batch can be generated
    independently (rank r generates start=r*n_local)."""
    u = splitmix64_uniform(seed, 6 * n, offset=6 * start).reshape(n, 6)
    e = np.asarray(extents, dtype=np.float64)
    t = u[:, :3] * (e[3:] - e[:3]) + e[:3]
    ang = u[:, 3:] * (2.0 * np.pi)
    R = euler_to_matrix(ang[:, 0], ang[:, 1], ang[:, 2])
    out = np.empty((n, 12))
    out[:, :9] = R.reshape(n, 9)
    out[:, 9:] = t
    return out


def identity_poses(n):
    out = np.zeros((n, 12))
    out[:, 0] = out[:, 4] = out[:, 8] = 1.0
    return out
