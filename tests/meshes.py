"""Small synthetic meshes for tests (test/test_fcl_shape_mesh_consistency.cpp style)."""
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
