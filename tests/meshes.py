"""Small synthetic meshes for tests (test/test_fcl_shape_mesh_consistency.cpp style)."""
import numpy as np


def uv_sphere(radius, seg=16, ring=16, center=(0.0, 0.0, 0.0)):
    """Tessellated sphere in the spirit of generateBVHModel(Sphere, seg, ring)
    (include/fcl/geometry/geometric_shape_to_BVH_model-inl.h): ring latitudes x seg longitudes."""
    verts = []
    for i in range(1, ring):
        theta = np.pi * i / ring
        for j in range(seg):
            phi = 2 * np.pi * j / seg
            verts.append([radius * np.sin(theta) * np.cos(phi), radius * np.sin(theta) * np.sin(phi), radius * np.cos(theta)])
    top = len(verts)
    verts.append([0, 0, radius])
    bot = len(verts)
    verts.append([0, 0, -radius])
    tris = []
    for j in range(seg):
        tris.append([top, j, (j + 1) % seg])
        base = (ring - 2) * seg
        tris.append([bot, base + (j + 1) % seg, base + j])
    for i in range(ring - 2):
        for j in range(seg):
            a = i * seg + j
            b = i * seg + (j + 1) % seg
            c = (i + 1) * seg + j
            d = (i + 1) * seg + (j + 1) % seg
            tris.append([a, c, b])
            tris.append([b, c, d])
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
