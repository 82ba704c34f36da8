"""ctypes front-end of the CPU ORACLE (test infrastructure, NOT product code).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  Nothing under fcl_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# FCL_SUM3_ORDER=1 in the environment selects the oracle built with the other association order of the three-term sums
# (fcl_oracle_vec.hpp `sum3`; tests/test_sum_order_hook.py re-runs the parity tests with both sides flipped)
_SUM3 = os.environ.get("FCL_SUM3_ORDER", "0")
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so" if _SUM3 == "0" else "liboracle_sum3_%s.so" % _SUM3)

SPLIT_MEAN, SPLIT_MEDIAN, SPLIT_BV_CENTER = 0, 1, 2

CONTACT_DTYPE = np.dtype(
    [("b1", "<i4"), ("b2", "<i4"), ("normal", "<f8", (3,)), ("pos", "<f8", (3,)), ("depth", "<f8")]
)
assert CONTACT_DTYPE.itemsize == 64


def build(force=False):
    """Compile the oracle with its Makefile (g++ -O2 -ffp-contract=off)."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".hpp"))]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs
    )
    if (force or stale) and _SUM3 != "0":
        os.makedirs(os.path.dirname(_LIB_PATH), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-pthread", "-DFCL_SUM3_ORDER=" + _SUM3,
                               "-shared", "-o", _LIB_PATH] +
                              [os.path.join(_HERE, f) for f in ("fcl_oracle_math.cpp", "fcl_oracle_bvh.cpp", "oracle_capi.cpp")])
    elif force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        vp, dp, ip, lp = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_longlong)
        L.orc_model_from_obj.restype = vp
        L.orc_model_from_obj.argtypes = [C.c_char_p, C.c_int]
        L.orc_model_from_arrays.restype = vp
        L.orc_model_from_arrays.argtypes = [dp, C.c_int, ip, C.c_int, C.c_int]
        L.orc_model_free.argtypes = [vp]
        L.orc_model_refit_topdown.restype = C.c_int
        L.orc_model_refit_topdown.argtypes = [vp, dp, C.c_int]
        L.orc_model_refit_bottomup.restype = C.c_int
        L.orc_model_refit_bottomup.argtypes = [vp, dp, C.c_int]
        L.orc_model_get_rss_axis.argtypes = [vp, dp]
        L.orc_merge_obbrss.argtypes = [dp, dp, dp]
        L.orc_fit3_obbrss.argtypes = [dp, dp]
        L.orc_model_partition.argtypes = [vp, ip, ip, ip]
        L.orc_model_counts.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_model_get.argtypes = [vp, dp, ip, ip, dp, dp, dp, dp, dp, dp]
        L.orc_collide_batch.restype = vp
        L.orc_collide_batch.argtypes = [vp, vp, C.c_longlong, dp, dp, C.c_longlong, C.c_int, C.c_int]
        L.orc_collide_seconds.restype = C.c_double
        L.orc_collide_seconds.argtypes = [vp]
        L.orc_collide_total.restype = C.c_longlong
        L.orc_collide_total.argtypes = [vp]
        L.orc_collide_copy.argtypes = [vp, ip, vp, lp, lp]
        L.orc_collide_free.argtypes = [vp]
        L.orc_distance_batch.restype = C.c_double
        L.orc_distance_batch.argtypes = [vp, vp, C.c_longlong, dp, dp, C.c_int, C.c_int, C.c_int,
                                         dp, dp, dp, ip, ip, lp, lp]
        L.orc_brute_collide.restype = C.c_longlong
        L.orc_brute_collide.argtypes = [vp, vp, dp, dp, ip, C.c_longlong]
        L.orc_brute_distance.restype = C.c_double
        L.orc_brute_distance.argtypes = [vp, vp, dp, dp, dp, dp, ip]
        L.orc_obb_disjoint.restype = C.c_int
        L.orc_obb_disjoint.argtypes = [dp, dp, dp, dp]
        L.orc_obb_overlap.restype = C.c_int
        L.orc_obb_overlap.argtypes = [dp] * 8
        L.orc_rect_distance.restype = C.c_double
        L.orc_rect_distance.argtypes = [dp, dp, dp, dp]
        L.orc_rss_distance.restype = C.c_double
        L.orc_rss_distance.argtypes = [dp, dp, dp, dp, dp, C.c_double, dp, dp, dp, C.c_double]
        L.orc_tri_intersect.restype = C.c_int
        L.orc_tri_intersect.argtypes = [dp, dp, dp, dp, C.c_int, C.POINTER(C.c_uint32), dp, dp, dp]
        L.orc_tri_distance.restype = C.c_double
        L.orc_tri_distance.argtypes = [dp, dp, dp, dp]
        L.orc_hardware_threads.restype = C.c_int
        _lib = L
    return _lib


def _d(a):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int32))


def _lp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_longlong))


def hardware_threads():
    return int(lib().orc_hardware_threads())


class Model:
    """Oracle BVHModel<OBBRSS<double>>."""

    def __init__(self, verts, tris, split=SPLIT_MEAN):
        self.verts = np.ascontiguousarray(verts, dtype=np.float64).reshape(-1, 3)
        self.tris = np.ascontiguousarray(tris, dtype=np.int32).reshape(-1, 3)
        self.h = lib().orc_model_from_arrays(_dp(self.verts), len(self.verts), _ip(self.tris), len(self.tris), split)
        nv, nt, nn = C.c_int(), C.c_int(), C.c_int()
        lib().orc_model_counts(self.h, C.byref(nv), C.byref(nt), C.byref(nn))
        self.num_vertices, self.num_tris, self.num_bvs = nv.value, nt.value, nn.value

    @classmethod
    def from_npz(cls, path, split=SPLIT_MEAN):
        z = np.load(path)
        return cls(z["verts"], z["tris"], split)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().orc_model_free(self.h)
                self.h = None
        except Exception:
            pass

    def refit_topdown(self, new_verts):
        """endReplaceModel(refit=True, bottomup=False): same topology, every BV refitted."""
        v = np.ascontiguousarray(new_verts, dtype=np.float64).reshape(-1, 3)
        rc = lib().orc_model_refit_topdown(self.h, _dp(v), len(v))
        if rc == 0:
            self.verts = v
        return rc

    def refit_bottomup(self, new_verts):
        """endReplaceModel(refit=True, bottomup=True), the reference's default (BVH_model-inl.h:952-1037)."""
        v = np.ascontiguousarray(new_verts, dtype=np.float64).reshape(-1, 3)
        rc = lib().orc_model_refit_bottomup(self.h, _dp(v), len(v))
        if rc == 0:
            self.verts = v
        return rc

    def partition(self):
        fp, npr, pi = np.empty(self.num_bvs, np.int32), np.empty(self.num_bvs, np.int32), np.empty(self.num_tris, np.int32)
        lib().orc_model_partition(self.h, _ip(fp), _ip(npr), _ip(pi))
        return fp, npr, pi

    def arrays(self):
        """Flattened node tree: dict of numpy arrays (axis row-major 9 per node)."""
        n = self.num_bvs
        out = dict(
            first_child=np.empty(n, np.int32), axis=np.empty((n, 9)), obb_To=np.empty((n, 3)),
            obb_ext=np.empty((n, 3)), rss_To=np.empty((n, 3)), rss_l=np.empty((n, 2)), rss_r=np.empty(n),
        )
        lib().orc_model_get(self.h, None, None, _ip(out["first_child"]), _dp(out["axis"]), _dp(out["obb_To"]),
                            _dp(out["obb_ext"]), _dp(out["rss_To"]), _dp(out["rss_l"]), _dp(out["rss_r"]))
        out["rss_axis"] = np.empty((n, 9))
        lib().orc_model_get_rss_axis(self.h, _dp(out["rss_axis"]))
        return out


def _poses(tf, n=None):
    if tf is None:
        return None
    tf = np.ascontiguousarray(tf, dtype=np.float64).reshape(-1, 12)
    return tf


def collide_batch(m1, m2, tf1, tf2=None, num_max_contacts=1, enable_contact=False, nthreads=1):
    """Returns dict(counts[n], contacts[total] (CONTACT_DTYPE), offsets[n+1], n_bv[n], n_leaf[n], seconds)."""
    tf1 = _poses(tf1)
    tf2 = _poses(tf2)
    n = len(tf1) if tf1 is not None else len(tf2)
    L = lib()
    hb = L.orc_collide_batch(m1.h, m2.h, n, _dp(tf1), _dp(tf2), int(num_max_contacts), int(enable_contact), nthreads)
    try:
        total = L.orc_collide_total(hb)
        counts = np.empty(n, np.int32)
        contacts = np.zeros(total, CONTACT_DTYPE)
        n_bv = np.empty(n, np.int64)
        n_leaf = np.empty(n, np.int64)
        L.orc_collide_copy(hb, _ip(counts), contacts.ctypes.data_as(C.c_void_p), _lp(n_bv), _lp(n_leaf))
        secs = L.orc_collide_seconds(hb)
    finally:
        L.orc_collide_free(hb)
    offsets = np.zeros(n + 1, np.int64)
    np.cumsum(counts, out=offsets[1:])
    return dict(counts=counts, contacts=contacts, offsets=offsets, n_bv=n_bv, n_leaf=n_leaf, seconds=secs)


def collide_mesh_sphere_batch(m1, radius, tf1, tf2, num_max_contacts=1, enable_contact=False, nthreads=1):
    """fcl::collide(BVHModel<OBBRSS>, tf1[i], Sphere(radius), tf2[i]); contacts carry b2 = -1 (Contact::NONE)."""
    tf1 = _poses(tf1)
    tf2 = _poses(tf2)
    n = len(tf1) if tf1 is not None else len(tf2)
    L = lib()
    L.orc_collide_mesh_sphere_batch.restype = C.c_void_p
    L.orc_collide_mesh_sphere_batch.argtypes = [C.c_void_p, C.c_double, C.c_longlong, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                                C.c_longlong, C.c_int, C.c_int]
    hb = L.orc_collide_mesh_sphere_batch(m1.h, float(radius), n, _dp(tf1), _dp(tf2), int(num_max_contacts), int(enable_contact), nthreads)
    try:
        total = L.orc_collide_total(hb)
        counts = np.empty(n, np.int32)
        contacts = np.zeros(total, CONTACT_DTYPE)
        n_bv = np.empty(n, np.int64)
        n_leaf = np.empty(n, np.int64)
        L.orc_collide_copy(hb, _ip(counts), contacts.ctypes.data_as(C.c_void_p), _lp(n_bv), _lp(n_leaf))
    finally:
        L.orc_collide_free(hb)
    offsets = np.zeros(n + 1, np.int64)
    np.cumsum(counts, out=offsets[1:])
    return dict(counts=counts, contacts=contacts, offsets=offsets, n_bv=n_bv, n_leaf=n_leaf)


def collide_mesh_plane_batch(m1, kind, normal, d, tf1, tf2, num_max_contacts=1, enable_contact=False, nthreads=1):
    """fcl::collide(BVHModel<OBBRSS>, tf1[i], Halfspace(normal, d) | Plane(normal, d), tf2[i]); kind 'halfspace' | 'plane'."""
    tf1 = _poses(tf1)
    tf2 = _poses(tf2)
    n = len(tf1) if tf1 is not None else len(tf2)
    L = lib()
    L.orc_collide_mesh_plane_batch.restype = C.c_void_p
    L.orc_collide_mesh_plane_batch.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_double, C.c_longlong, C.POINTER(C.c_double),
                                               C.POINTER(C.c_double), C.c_longlong, C.c_int, C.c_int]
    nv = np.ascontiguousarray(normal, np.float64).reshape(3)
    hb = L.orc_collide_mesh_plane_batch(m1.h, 0 if kind == "halfspace" else 1, _dp(nv), float(d), n, _dp(tf1), _dp(tf2),
                                        int(min(num_max_contacts, 2**62)), int(enable_contact), nthreads)
    try:
        total = L.orc_collide_total(hb)
        counts = np.empty(n, np.int32)
        contacts = np.zeros(total, CONTACT_DTYPE)
        n_bv = np.empty(n, np.int64)
        n_leaf = np.empty(n, np.int64)
        L.orc_collide_copy(hb, _ip(counts), contacts.ctypes.data_as(C.c_void_p), _lp(n_bv), _lp(n_leaf))
    finally:
        L.orc_collide_free(hb)
    offsets = np.zeros(n + 1, np.int64)
    np.cumsum(counts, out=offsets[1:])
    return dict(counts=counts, contacts=contacts, offsets=offsets, n_bv=n_bv, n_leaf=n_leaf)


def brute_mesh_plane(m1, kind, normal, d, tf1, tf2):
    L = lib()
    L.orc_brute_mesh_plane.restype = C.c_longlong
    L.orc_brute_mesh_plane.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                       C.POINTER(C.c_int32), C.c_longlong]
    nv = np.ascontiguousarray(normal, np.float64).reshape(3)
    a = np.ascontiguousarray(tf1, dtype=np.float64).reshape(12)
    b = np.ascontiguousarray(tf2, dtype=np.float64).reshape(12)
    out = np.empty(m1.num_tris, np.int32)
    k = L.orc_brute_mesh_plane(m1.h, 0 if kind == "halfspace" else 1, _dp(nv), float(d), _dp(a), _dp(b), _ip(out), len(out))
    return out[:k]


def plane_tri_intersect(kind, normal, d, tf_shape, tri9, tf_tri):
    """halfspaceTriangleIntersect / planeTriangleIntersect: (hit, contact point, depth, normal)."""
    L = lib()
    L.orc_plane_tri_intersect.restype = C.c_int
    L.orc_plane_tri_intersect.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                          C.POINTER(C.c_double), C.POINTER(C.c_double)]
    nv = np.ascontiguousarray(normal, np.float64).reshape(3)
    a = None if tf_shape is None else np.ascontiguousarray(tf_shape, np.float64).reshape(12)
    b = None if tf_tri is None else np.ascontiguousarray(tf_tri, np.float64).reshape(12)
    t = np.ascontiguousarray(tri9, np.float64).reshape(9)
    out = np.zeros(7)
    hit = L.orc_plane_tri_intersect(0 if kind == "halfspace" else 1, _dp(nv), float(d), _dp(a), _dp(t), _dp(b), _dp(out))
    return bool(hit), out[:3].copy(), float(out[3]), out[4:].copy()


def brute_mesh_sphere(m1, radius, tf1, tf2):
    """Ids of every triangle of m1 (posed by tf1) the sphere (centre tf2's translation) intersects."""
    L = lib()
    L.orc_brute_mesh_sphere.restype = C.c_longlong
    L.orc_brute_mesh_sphere.argtypes = [C.c_void_p, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int32), C.c_longlong]
    a = np.ascontiguousarray(tf1, dtype=np.float64).reshape(12)
    b = np.ascontiguousarray(tf2, dtype=np.float64).reshape(12)
    out = np.empty(m1.num_tris, np.int32)
    k = L.orc_brute_mesh_sphere(m1.h, float(radius), _dp(a), _dp(b), _ip(out), len(out))
    return out[:k]


def sphere_tri_intersect(center, radius, tri9):
    """(hit, contact_point[3], depth, normal[3]) of sphereTriangleIntersect in one frame."""
    L = lib()
    L.orc_sphere_tri_intersect.argtypes = [C.POINTER(C.c_double), C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    c = np.ascontiguousarray(center, dtype=np.float64).reshape(3)
    t = np.ascontiguousarray(tri9, dtype=np.float64).reshape(9)
    out = np.zeros(7)
    hit = L.orc_sphere_tri_intersect(_dp(c), float(radius), _dp(t), _dp(out))
    return bool(hit), out[:3].copy(), float(out[3]), out[4:].copy()


def distance_batch(m1, m2, tf1, tf2=None, enable_nearest_points=True, qsize=2, nthreads=1):
    tf1 = _poses(tf1)
    tf2 = _poses(tf2)
    n = len(tf1) if tf1 is not None else len(tf2)
    dist = np.empty(n)
    p1 = np.empty((n, 3))
    p2 = np.empty((n, 3))
    b1 = np.empty(n, np.int32)
    b2 = np.empty(n, np.int32)
    n_bv = np.empty(n, np.int64)
    n_leaf = np.empty(n, np.int64)
    secs = lib().orc_distance_batch(m1.h, m2.h, n, _dp(tf1), _dp(tf2), int(enable_nearest_points), qsize, nthreads,
                                    _dp(dist), _dp(p1), _dp(p2), _ip(b1), _ip(b2), _lp(n_bv), _lp(n_leaf))
    return dict(min_distance=dist, p1=p1, p2=p2, b1=b1, b2=b2, n_bv=n_bv, n_leaf=n_leaf, seconds=secs)


def continuous_collide_translation_batch(m1, m2, tf1_beg, tf1_end, tf2_beg=None, tf2_end=None, nthreads=1):
    """fcl::continuousCollide(o1, tf1_beg, tf1_end, o2, tf2_beg, tf2_end, request) with ccd_motion_type = CCDM_TRANS and
    ccd_solver_type = CCDC_CONSERVATIVE_ADVANCEMENT, one call per row."""
    L = lib()
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    L.orc_continuous_collide_translation_batch.restype = C.c_double
    L.orc_continuous_collide_translation_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, dp, dp, dp, dp, C.c_int, ip, dp, dp, dp, ip]
    a0, a1, b0, b1 = _poses(tf1_beg), _poses(tf1_end), _poses(tf2_beg), _poses(tf2_end)
    n = max(len(x) for x in (a0, a1, b0, b1) if x is not None)
    hit, toc, it = np.empty(n, np.int32), np.empty(n), np.empty(n, np.int32)
    c1, c2 = np.empty((n, 12)), np.empty((n, 12))
    secs = L.orc_continuous_collide_translation_batch(m1.h, m2.h, n, _dp(a0), _dp(a1), _dp(b0), _dp(b1), nthreads, _ip(hit), _dp(toc),
                                                      _dp(c1), _dp(c2), _ip(it))
    return dict(is_collide=hit.astype(bool), time_of_contact=toc, contact_tf1=c1, contact_tf2=c2, iterations=it, seconds=secs)


def brute_collide(m1, m2, tf1, tf2=None):
    tf1 = None if tf1 is None else np.ascontiguousarray(tf1, dtype=np.float64).reshape(12)
    tf2 = None if tf2 is None else np.ascontiguousarray(tf2, dtype=np.float64).reshape(12)
    cap = 1 << 16
    while True:
        pairs = np.empty((cap, 2), np.int32)
        k = lib().orc_brute_collide(m1.h, m2.h, _dp(tf1), _dp(tf2), _ip(pairs), cap)
        if k <= cap:
            return pairs[:k].copy()
        cap = int(k)


def distance_mesh_sphere_batch(m1, radius, tf1, tf2, brute=False, nthreads=1):
    """fcl::distance(BVHModel<OBBRSS>, tf1[i], Sphere(radius), tf2[i]): p1 in the mesh frame, p2 in the sphere frame
    (the reference leaves them local), b1 = closest triangle; centre within the radius of a triangle: -1 and NaN
    points (the reference leaves that case undefined).  brute=True tests every triangle in primitive order."""
    L = lib()
    dp, lp, ip = C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.POINTER(C.c_int32)
    L.orc_distance_mesh_sphere_batch.restype = C.c_double
    L.orc_distance_mesh_sphere_batch.argtypes = [C.c_void_p, C.c_double, C.c_longlong, dp, dp, C.c_int, C.c_int, dp, dp, dp,
                                                 ip, lp, lp]
    tf1 = _poses(tf1)
    tf2 = _poses(tf2)
    n = len(tf1) if tf1 is not None else len(tf2)
    dist = np.empty(n)
    p1 = np.empty((n, 3))
    p2 = np.empty((n, 3))
    b1 = np.empty(n, np.int32)
    n_bv = np.zeros(n, np.int64)
    n_leaf = np.zeros(n, np.int64)
    secs = L.orc_distance_mesh_sphere_batch(m1.h, float(radius), n, _dp(tf1), _dp(tf2), int(brute), nthreads, _dp(dist),
                                            _dp(p1), _dp(p2), _ip(b1), _lp(n_bv), _lp(n_leaf))
    return dict(min_distance=dist, p1=p1, p2=p2, b1=b1, n_bv=n_bv, n_leaf=n_leaf, seconds=secs)


def sphere_tri_distance(center, radius, tri9):
    """(separated, distance, point_on_sphere[3], point_on_triangle[3]) of sphereTriangleDistance in one frame."""
    L = lib()
    L.orc_sphere_tri_distance.argtypes = [C.POINTER(C.c_double), C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    c = np.ascontiguousarray(center, dtype=np.float64).reshape(3)
    t = np.ascontiguousarray(tri9, dtype=np.float64).reshape(9)
    out = np.zeros(7)
    ok = L.orc_sphere_tri_distance(_dp(c), float(radius), _dp(t), _dp(out))
    return bool(ok), float(out[0]), out[1:4].copy(), out[4:].copy()


def sphere_bv(radius, tf):
    """computeBV<OBBRSS>(Sphere(radius), tf) -> dict(axis, obb_To, obb_ext, rss_To, rss_l, rss_r)."""
    L = lib()
    dp = C.POINTER(C.c_double)
    L.orc_sphere_bv.restype = None
    L.orc_sphere_bv.argtypes = [C.c_double, dp, dp, dp, dp, dp, dp, dp]
    tf = np.ascontiguousarray(tf, dtype=np.float64).reshape(12)
    axis, oT, oe, rT, rl, rr = np.zeros(9), np.zeros(3), np.zeros(3), np.zeros(3), np.zeros(2), np.zeros(1)
    L.orc_sphere_bv(float(radius), _dp(tf), _dp(axis), _dp(oT), _dp(oe), _dp(rT), _dp(rl), _dp(rr))
    return dict(axis=axis.reshape(3, 3), obb_To=oT, obb_ext=oe, rss_To=rT, rss_l=rl, rss_r=float(rr[0]))


def brute_distance(m1, m2, tf1, tf2=None):
    tf1 = None if tf1 is None else np.ascontiguousarray(tf1, dtype=np.float64).reshape(12)
    tf2 = None if tf2 is None else np.ascontiguousarray(tf2, dtype=np.float64).reshape(12)
    p1 = np.empty(3)
    p2 = np.empty(3)
    b = np.empty(2, np.int32)
    d = lib().orc_brute_distance(m1.h, m2.h, _dp(tf1), _dp(tf2), _dp(p1), _dp(p2), _ip(b))
    return d, p1, p2, b


def _c9(a):
    return np.ascontiguousarray(a, dtype=np.float64).reshape(-1)


def obb_disjoint(B, T, a, b):
    B, T, a, b = _c9(B), _c9(T), _c9(a), _c9(b)
    return bool(lib().orc_obb_disjoint(_dp(B), _dp(T), _dp(a), _dp(b)))


def obb_overlap(R0, T0, axis1, To1, ext1, axis2, To2, ext2):
    args = [_c9(x) for x in (R0, T0, axis1, To1, ext1, axis2, To2, ext2)]
    return bool(lib().orc_obb_overlap(*[_dp(x) for x in args]))


def rect_distance(R, T, a, b):
    R, T, a, b = _c9(R), _c9(T), _c9(a), _c9(b)
    return float(lib().orc_rect_distance(_dp(R), _dp(T), _dp(a), _dp(b)))


def rss_distance(R0, T0, axis1, To1, l1, r1, axis2, To2, l2, r2):
    R0, T0, axis1, To1, l1, axis2, To2, l2 = [_c9(x) for x in (R0, T0, axis1, To1, l1, axis2, To2, l2)]
    return float(lib().orc_rss_distance(_dp(R0), _dp(T0), _dp(axis1), _dp(To1), _dp(l1), float(r1),
                                        _dp(axis2), _dp(To2), _dp(l2), float(r2)))


def tri_intersect(P, Q, R=None, T=None, want_contacts=False):
    P, Q = _c9(P), _c9(Q)
    R = _c9(np.eye(3) if R is None else R)
    T = _c9(np.zeros(3) if T is None else T)
    nc = C.c_uint32(0)
    contacts = np.zeros(6)
    depth = np.zeros(1)
    normal = np.zeros(3)
    hit = lib().orc_tri_intersect(_dp(P), _dp(Q), _dp(R), _dp(T), int(want_contacts), C.byref(nc),
                                  _dp(contacts), _dp(depth), _dp(normal))
    if not want_contacts:
        return bool(hit)
    return bool(hit), int(nc.value), contacts.reshape(2, 3), float(depth[0]), normal


def tri_distance(S, T):
    S, T = _c9(S), _c9(T)
    P = np.zeros(3)
    Q = np.zeros(3)
    d = lib().orc_tri_distance(_dp(S), _dp(T), _dp(P), _dp(Q))
    return float(d), P, Q


def broadphase(models, geom1, tf1, geom2, tf2, num_max_contacts=1, enable_contact=False, nthreads=1, narrowphase=True):
    """NaiveCollisionManager::collide(other) with the default callback evaluated for EVERY overlapping pair:
    dict(pairs (m, 2) int32 in the brute-force manager's order, counts[m] = numContacts per pair, aabb1 (n1, 6))."""
    L = lib()
    L.orc_broadphase.restype = C.c_longlong
    L.orc_broadphase.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.c_longlong, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.c_longlong,
                                 C.POINTER(C.c_int32), C.POINTER(C.c_double), C.c_longlong, C.c_int, C.c_int, C.POINTER(C.c_int32),
                                 C.POINTER(C.c_int32), C.c_longlong, C.POINTER(C.c_double)]
    hs = (C.c_void_p * len(models))(*[m.h for m in models])
    g1, g2 = np.ascontiguousarray(geom1, np.int32), np.ascontiguousarray(geom2, np.int32)
    t1, t2 = _poses(tf1), _poses(tf2)
    aabb1 = np.zeros((len(g1), 6))
    total = L.orc_broadphase(len(models), hs, len(g1), _ip(g1), _dp(t1), len(g2), _ip(g2), _dp(t2), 1, 0, 1, None, None, 0, _dp(aabb1))
    pairs = np.zeros((total, 2), np.int32)
    counts = np.zeros(total, np.int32) if narrowphase else None
    L.orc_broadphase(len(models), hs, len(g1), _ip(g1), _dp(t1), len(g2), _ip(g2), _dp(t2), int(min(num_max_contacts, 2**62)),
                     int(enable_contact), int(nthreads), _ip(pairs), _ip(counts), total, None)
    return dict(pairs=pairs, counts=counts, aabb1=aabb1)


# ------------------------------------------------------------------------------------------------
# executed-operation counters (fcl_oracle_counted.cpp): the oracle recompiled over a counting scalar
# ------------------------------------------------------------------------------------------------
_CNT_PATH = os.path.join(_HERE, "_build", "liboracle_counted.so" if _SUM3 == "0" else "liboracle_counted_sum3_%s.so" % _SUM3)
_cnt = None
OP_NAMES = ("mul", "add", "cmp", "div", "sqrt")


def counted_lib():
    global _cnt
    if _cnt is None:
        build()
        if not os.path.exists(_CNT_PATH) and _SUM3 != "0":
            subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-pthread", "-Wno-unused-function",
                                   "-DFCL_SUM3_ORDER=" + _SUM3, "-shared", "-o", _CNT_PATH, os.path.join(_HERE, "fcl_oracle_counted.cpp")])
        elif not os.path.exists(_CNT_PATH):
            subprocess.check_call(["make", "-C", _HERE, "-s", "_build/liboracle_counted.so"])
        L = C.CDLL(_CNT_PATH)
        vp, dp, ip, lp = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_longlong)
        L.orcc_model.restype = vp
        L.orcc_model.argtypes = [dp, C.c_int, ip, C.c_int]
        L.orcc_model_free.argtypes = [vp]
        L.orcc_collide.argtypes = [vp, vp, dp, dp, C.c_longlong, C.c_longlong, C.c_int, C.c_int, lp, lp, lp, dp]
        L.orcc_distance.argtypes = [vp, vp, dp, dp, C.c_longlong, C.c_int, C.c_int, lp, lp, lp, dp]
        _cnt = L
    return _cnt


class CountedModel:
    """A model of the counting build (its BVH is built with the same arithmetic, so it equals Model's)."""

    def __init__(self, verts, tris):
        v = np.ascontiguousarray(verts, np.float64)
        t = np.ascontiguousarray(tris, np.int32)
        self._h = counted_lib().orcc_model(_dp(v), len(v), _ip(t), len(t))
        self._owner = None

    @classmethod
    def share(cls, model):
        """The counting scalar is a struct holding one double, so a Model built by the plain library has the very
        same layout: large meshes (cfg4 / cfg5) are built once and read by both builds."""
        counted_lib()
        m = cls.__new__(cls)
        m._h, m._owner = model.h, model
        return m

    def __del__(self):
        if getattr(self, "_h", None) and getattr(self, "_owner", None) is None and _cnt is not None:
            _cnt.orcc_model_free(self._h)
        self._h = None


def counted_query(kind, m1, m2, tf1, tf2=None, num_max_contacts=1, enable_contact=False, enable_nearest_points=True,
                  nthreads=1):
    """kind = 'collide' | 'distance'.  Returns ops[n,5] (executed mul, add, cmp, div, sqrt per query of the reference's
    sequential traversal), n_bv[n], n_leaf[n], value[n] (numContacts / min_distance)."""
    tf1, tf2 = _poses(tf1), _poses(tf2)
    n = len(tf1) if tf1 is not None else len(tf2)
    ops = np.zeros((n, 5), np.int64)
    nbv, nleaf, val = np.zeros(n, np.int64), np.zeros(n, np.int64), np.zeros(n, np.float64)
    L = counted_lib()
    if kind == "collide":
        L.orcc_collide(m1._h, m2._h, _dp(tf1), _dp(tf2), n, int(min(num_max_contacts, 2**62)), int(enable_contact),
                       int(nthreads), _lp(ops), _lp(nbv), _lp(nleaf), _dp(val))
    else:
        L.orcc_distance(m1._h, m2._h, _dp(tf1), _dp(tf2), n, int(enable_nearest_points), int(nthreads), _lp(ops),
                        _lp(nbv), _lp(nleaf), _dp(val))
    return {"ops": ops, "n_bv": nbv, "n_leaf": nleaf, "value": val}


def merge_obbrss(a30, b30):
    """OBBRSS::operator+ on two volumes (30 doubles each: axis9, obb_To3, obb_ext3, rss_axis9, rss_To3, rss_l2, rss_r)."""
    a, b = np.ascontiguousarray(a30, dtype=np.float64), np.ascontiguousarray(b30, dtype=np.float64)
    out = np.empty(30)
    lib().orc_merge_obbrss(_dp(a), _dp(b), _dp(out))
    return out


def fit3_obbrss(pts):
    """OBBRSS_fit_functions::fit3 over three points -> 30 doubles."""
    p = np.ascontiguousarray(pts, dtype=np.float64).reshape(9)
    out = np.empty(30)
    lib().orc_fit3_obbrss(_dp(p), _dp(out))
    return out
