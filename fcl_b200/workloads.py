"""Synthetic meshes and pose batches for the BASELINE configurations that have no fixture in the reference
(cfg4: 7-link arm vs a 200k-triangle scene; cfg5: two 1M-triangle meshes) and for the tests
(test/test_fcl_shape_mesh_consistency.cpp style tessellations).  Everything is seeded numpy: bench.py and the
parity tests generate identical inputs on any box."""
import numpy as np


def uv_sphere(radius, seg=16, ring=16, center=(0.0, 0.0, 0.0)):
    """Tessellated sphere in the spirit of generateBVHModel(Sphere, seg, ring)
    (include/fcl/geometry/geometric_shape_to_BVH_model-inl.h): ring latitudes x seg longitudes."""
    i = np.arange(1, ring, dtype=np.float64)[:, None]
    j = np.arange(seg, dtype=np.float64)[None, :]
    theta, phi = np.pi * i / ring, 2 * np.pi * j / seg
    body = np.stack([radius * np.sin(theta) * np.cos(phi), radius * np.sin(theta) * np.sin(phi),
                     np.broadcast_to(radius * np.cos(theta), (ring - 1, seg))], axis=-1).reshape(-1, 3)
    top = len(body)
    bot = top + 1
    verts = np.concatenate([body, [[0, 0, radius]], [[0, 0, -radius]]])
    jj = np.arange(seg)
    jn = (jj + 1) % seg
    base = (ring - 2) * seg
    caps = np.stack([np.stack([np.full(seg, top), jj, jn], axis=1), np.stack([np.full(seg, bot), base + jn, base + jj], axis=1)],
                    axis=1).reshape(-1, 3)
    ii = np.arange(ring - 2)[:, None] * seg
    a, bq, c, d = ii + jj[None, :], ii + jn[None, :], ii + seg + jj[None, :], ii + seg + jn[None, :]
    quads = np.stack([np.stack([a, c, bq], axis=-1), np.stack([bq, c, d], axis=-1)], axis=2).reshape(-1, 3)
    tris = np.concatenate([caps, quads])
    v = np.asarray(verts, dtype=np.float64) + np.asarray(center, dtype=np.float64)
    return v, np.asarray(tris, dtype=np.int32)


def box_mesh(hx, hy, hz, center=(0.0, 0.0, 0.0)):
    s = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], dtype=np.float64)
    v = s * np.array([hx, hy, hz]) + np.asarray(center, dtype=np.float64)
    t = np.array([[0, 2, 1], [0, 3, 2], [4, 5, 6], [4, 6, 7], [0, 1, 5], [0, 5, 4], [2, 3, 7], [2, 7, 6], [1, 2, 6], [1, 6, 5], [0, 4, 7], [0, 7, 3]], dtype=np.int32)
    return v, t


def random_soup(n_tris, seed, scale=1.0, tri_size=0.3):
    rng = np.random.default_rng(seed)
    c = rng.uniform(-scale, scale, size=(n_tris, 1, 3))
    v = (c + rng.normal(0, tri_size, size=(n_tris, 3, 3))).reshape(-1, 3)
    t = np.arange(3 * n_tris, dtype=np.int32).reshape(n_tris, 3)
    return v, t


def heightfield(n, size=10.0, seed=0, amp=0.3):
    """(n x n x 2) triangle height field over a size x size square (cfg4-style scene mesh)."""
    rng = np.random.default_rng(seed)
    xs = np.linspace(-size / 2, size / 2, n + 1)
    X, Y = np.meshgrid(xs, xs, indexing="ij")
    Z = amp * np.sin(1.7 * X) * np.cos(1.3 * Y) + 0.15 * amp * rng.normal(size=X.shape)
    V = np.stack([X, Y, Z], -1).reshape(-1, 3)
    idx = np.arange((n + 1) * (n + 1)).reshape(n + 1, n + 1)
    a, b, c, d = idx[:-1, :-1].ravel(), idx[1:, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, 1:].ravel()
    T = np.concatenate([np.stack([a, b, c], 1), np.stack([b, d, c], 1)]).astype(np.int32)
    return V, T


def noisy_sphere(radius, seg, ring, seed, noise=0.02, scale=(1.0, 1.0, 1.0)):
    """Noise-displaced tessellated sphere (cfg5-style synthetic mesh); `scale` stretches it into a link shape."""
    v, t = uv_sphere(radius, seg, ring)
    rng = np.random.default_rng(seed)
    v = v * (1.0 + noise * rng.normal(size=(len(v), 1)))
    return v * np.asarray(scale, dtype=np.float64), t


def serial_chain_poses(q, link_len=0.6):
    """Forward kinematics of a 7-joint serial arm (alternating z / y revolute joints, links along x):
    q (n,7) joint angles -> (n,7,12) link pose records, base at the origin raised by 1."""
    n = len(q)
    R = np.tile(np.eye(3), (n, 1, 1))
    p = np.tile(np.array([0.0, 0.0, 1.0]), (n, 1))
    out = np.empty((n, 7, 12))
    for j in range(7):
        c, s = np.cos(q[:, j]), np.sin(q[:, j])
        J = np.zeros((n, 3, 3))
        if j % 2 == 0:
            J[:, 0, 0], J[:, 0, 1], J[:, 1, 0], J[:, 1, 1], J[:, 2, 2] = c, -s, s, c, 1.0
        else:
            J[:, 0, 0], J[:, 0, 2], J[:, 2, 0], J[:, 2, 2], J[:, 1, 1] = c, s, -s, c, 1.0
        R = R @ J
        centre = p + 0.5 * link_len * R[:, :, 0]
        out[:, j, :9] = R.reshape(n, 9)
        out[:, j, 9:] = centre
        p = p + link_len * R[:, :, 0]
    return out


def shell_poses(n, lo, hi, seed, start=0):
    """cfg5 pose batch: random orientation (eulerToMatrix of three U[0, 2 pi) angles), centre at a uniformly random
    direction and U[lo, hi) distance.  Pose i depends on (seed, start + i) only, so ranks generate disjoint shards."""
    from .poses import euler_to_matrix, splitmix64_uniform

    u = splitmix64_uniform(seed, 7 * n, offset=7 * start).reshape(n, 7)
    ang = u[:, :3] * (2.0 * np.pi)
    # uniform direction from two uniforms (cylinder projection)
    z = 2.0 * u[:, 3] - 1.0
    phi = 2.0 * np.pi * u[:, 4]
    rxy = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    d = np.stack([rxy * np.cos(phi), rxy * np.sin(phi), z], axis=1)
    P = np.empty((n, 12))
    P[:, :9] = euler_to_matrix(ang[:, 0], ang[:, 1], ang[:, 2]).reshape(n, 9)
    P[:, 9:] = d * (lo + (hi - lo) * u[:, 5:6])
    return P


def arm_configurations(n, seed, start=0):
    """cfg4: n robot configurations drawn uniformly in joint space (7 joints, U[-pi, pi)) -> (n, 7, 12) link poses
    through the fixed serial-chain forward kinematics above."""
    from .poses import splitmix64_uniform

    q = (splitmix64_uniform(seed, 7 * n, offset=7 * start).reshape(n, 7) * 2.0 - 1.0) * np.pi
    return serial_chain_poses(q)


def cfg4_meshes():
    """Scene: 199,712-triangle height field over a 10 m square; 7 links: stretched noisy spheres, 4,900 triangles each."""
    scene = heightfield(316, size=10.0, seed=1, amp=0.5)
    links = [noisy_sphere(0.12, 50, 51, seed=10 + j, scale=(2.6, 1.0, 1.0)) for j in range(7)]
    return scene, links


def cfg5_meshes(seg=710, ring=705):
    """Two noise-displaced unit spheres of 999,680 triangles each (seeds 21 / 22)."""
    return noisy_sphere(1.0, seg, ring, seed=21), noisy_sphere(1.0, seg, ring, seed=22)
