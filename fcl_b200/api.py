"""Host-side mirror of the reference's interface for the OBBRSS mesh-mesh path.

Names, argument meaning and error behaviour follow the reference
(/root/reference/include/fcl/...):
  BVHModel            geometry/bvh/BVH_model.h:62-330 (beginModel/addSubModel/addTriangle/endModel)
  CollisionRequest    narrowphase/collision_request.h:52-106
  CollisionResult     narrowphase/collision_result.h:52-93
  Contact             narrowphase/contact.h:48-91
  DistanceRequest     narrowphase/distance_request.h:52-113
  DistanceResult      narrowphase/distance_result.h:51-107
  collide / distance  narrowphase/collision-inl.h:95-207, narrowphase/distance-inl.h:92-246
plus the batched entry points (new API beside them) which are the product: one call
evaluates n pose pairs on the GPU through the C ABI (include/fclgpu.h).
"""
import ctypes as C
import sys

import numpy as np

from . import _capi
from ._capi import CONTACT_DTYPE, CONTACT_F32_DTYPE, CONTACT_IDS_DTYPE, FclGpuError, check, addr

CONTACT_FULL, CONTACT_IDS, CONTACT_F32 = 0, 1, 2
CONTACT_DTYPES = {CONTACT_FULL: CONTACT_DTYPE, CONTACT_IDS: CONTACT_IDS_DTYPE, CONTACT_F32: CONTACT_F32_DTYPE}

# BVHReturnCode (geometry/bvh/BVH_internal.h:61-72)
BVH_OK = 0
BVH_ERR_MODEL_OUT_OF_MEMORY = -1
BVH_ERR_BUILD_OUT_OF_SEQUENCE = -2
BVH_ERR_BUILD_EMPTY_MODEL = -3
BVH_ERR_BUILD_EMPTY_PREVIOUS_FRAME = -4
BVH_ERR_UNSUPPORTED_FUNCTION = -5
BVH_ERR_UNUPDATED_MODEL = -6
BVH_ERR_INCORRECT_DATA = -7
BVH_ERR_UNKNOWN = -8

# BVHBuildState (BVH_internal.h:48-57)
BVH_BUILD_STATE_EMPTY, BVH_BUILD_STATE_BEGUN, BVH_BUILD_STATE_PROCESSED = 0, 1, 2
BVH_BUILD_STATE_UPDATE_BEGUN, BVH_BUILD_STATE_UPDATED, BVH_BUILD_STATE_REPLACE_BEGUN = 3, 4, 5

SPLIT_METHOD_MEAN, SPLIT_METHOD_MEDIAN, SPLIT_METHOD_BV_CENTER = 0, 1, 2

DBL_MAX = sys.float_info.max


class Transform3:
    """Rigid transform p -> R p + t (the reference's Transform3<double>, an Eigen Isometry)."""

    __slots__ = ("R", "t")

    def __init__(self, R=None, t=None):
        self.R = np.eye(3) if R is None else np.asarray(R, dtype=np.float64).reshape(3, 3).copy()
        self.t = np.zeros(3) if t is None else np.asarray(t, dtype=np.float64).reshape(3).copy()

    @staticmethod
    def Identity():
        return Transform3()

    def linear(self):
        return self.R

    def translation(self):
        return self.t

    def to_pose12(self):
        return np.concatenate([self.R.reshape(9), self.t])

    @staticmethod
    def from_pose12(p):
        p = np.asarray(p, dtype=np.float64).reshape(12)
        return Transform3(p[:9].reshape(3, 3), p[9:])

    @staticmethod
    def from_matrix4_colmajor(m16):
        m16 = np.ascontiguousarray(m16, dtype=np.float64).reshape(16)
        out = np.empty(12)
        _capi.lib().fclgpu_pose_from_colmajor4x4(addr(m16), addr(out))
        return Transform3.from_pose12(out)


def _poses(tf, n=None):
    """Accepts None (identity), a Transform3, an (n,12) array or a torch tensor; returns (array_or_tensor, n)."""
    if tf is None:
        return None, n
    if isinstance(tf, Transform3):
        return tf.to_pose12().reshape(1, 12), 1
    if isinstance(tf, np.ndarray) or isinstance(tf, (list, tuple)):
        a = np.ascontiguousarray(tf, dtype=np.float64).reshape(-1, 12)
        return a, len(a)
    # torch tensor
    if tf.dtype != _torch().float64 or not tf.is_contiguous():
        raise ValueError("pose tensors must be contiguous float64 of shape (n, 12)")
    return tf, tf.numel() // 12


def _torch():
    import torch

    return torch


class BVHModel:
    """BVHModel<OBBRSS<double>> (triangle meshes only).

    Build protocol and return codes as the reference: beginModel() -> addSubModel()/
    addTriangle()* -> endModel() (BVH_model-inl.h:207-253, 256-446, 450-517).  endModel()
    builds the OBBRSS tree on the host; the flattened tree is uploaded to a GPU on first use.
    With build_on_device=True endModel() only records the mesh and the tree is built by kernels on
    the GPU that first uses the model (same tree, bit for bit; mean / BV-centre split rules).
    """

    def __init__(self, split_method=SPLIT_METHOD_MEAN, build_on_device=False):
        self.build_on_device = bool(build_on_device)
        self.split_method = split_method
        self.build_state = BVH_BUILD_STATE_EMPTY
        self._verts = []
        self._tris = []
        self.vertices = np.zeros((0, 3))
        self.tri_indices = np.zeros((0, 3), np.int32)
        self.num_vertices = 0
        self.num_tris = 0
        self._bvh = None
        self._dev = {}
        # CollisionGeometry defaults (geometry/collision_geometry.h): always "occupied"
        self.cost_density = 1.0
        self.threshold_occupied = 1.0
        self.threshold_free = 0.0

    # -- CollisionGeometry surface used by the dispatch --
    def getObjectType(self):
        return "OT_BVH"

    def getNodeType(self):
        return "BV_OBBRSS"

    def isOccupied(self):
        return self.cost_density >= self.threshold_occupied

    def isFree(self):
        return self.cost_density <= self.threshold_free

    # -- build protocol --
    def beginModel(self, num_tris=0, num_vertices=0):
        if self.build_state != BVH_BUILD_STATE_EMPTY:
            self._release()
            self._verts, self._tris = [], []
            self.num_vertices = self.num_tris = 0
        if self.build_state != BVH_BUILD_STATE_EMPTY:
            sys.stderr.write("BVH Warning! Call beginModel() on a BVHModel that is not empty. This model was cleared "
                             "and previous triangles/vertices were lost.\n")
            self.build_state = BVH_BUILD_STATE_EMPTY
            return BVH_ERR_BUILD_OUT_OF_SEQUENCE
        self.build_state = BVH_BUILD_STATE_BEGUN
        return BVH_OK

    def addSubModel(self, ps, ts=None):
        if self.build_state == BVH_BUILD_STATE_PROCESSED:
            sys.stderr.write("BVH Warning! Call addSubModel() in a wrong order. addSubModel() was ignored. Must do a "
                             "beginModel() to clear the model for addition of new vertices.\n")
            return BVH_ERR_BUILD_OUT_OF_SEQUENCE
        ps = np.asarray(ps, dtype=np.float64).reshape(-1, 3)
        offset = self.num_vertices
        self._verts.append(ps)
        self.num_vertices += len(ps)
        if ts is not None:
            ts = np.asarray(ts, dtype=np.int64).reshape(-1, 3) + offset
            self._tris.append(ts.astype(np.int32))
            self.num_tris += len(ts)
        return BVH_OK

    def addTriangle(self, p1, p2, p3):
        if self.build_state == BVH_BUILD_STATE_PROCESSED:
            sys.stderr.write("BVH Warning! Call addTriangle() in a wrong order. addTriangle() was ignored. Must do a "
                             "beginModel() to clear the model for addition of new triangles.\n")
            return BVH_ERR_BUILD_OUT_OF_SEQUENCE
        return self.addSubModel(np.array([p1, p2, p3], dtype=np.float64), np.array([[0, 1, 2]]))

    def endModel(self):
        if self.build_state != BVH_BUILD_STATE_BEGUN:
            sys.stderr.write("BVH Warning! Call endModel() in wrong order. endModel() was ignored.\n")
            return BVH_ERR_BUILD_OUT_OF_SEQUENCE
        if self.num_tris == 0 and self.num_vertices == 0:
            sys.stderr.write("BVH Error! endModel() called on model with no triangles and vertices.\n")
            return BVH_ERR_BUILD_EMPTY_MODEL
        if self.num_tris == 0:
            sys.stderr.write("BVH Error! point-cloud models are not supported on the OBBRSS mesh-mesh path.\n")
            return BVH_ERR_UNSUPPORTED_FUNCTION
        self.vertices = np.ascontiguousarray(np.concatenate(self._verts), dtype=np.float64)
        self.tri_indices = np.ascontiguousarray(np.concatenate(self._tris), dtype=np.int32)
        if self.build_on_device:
            if self.split_method not in (SPLIT_METHOD_MEAN, SPLIT_METHOD_MEDIAN, SPLIT_METHOD_BV_CENTER):
                return BVH_ERR_UNSUPPORTED_FUNCTION
            if self.tri_indices.min() < 0 or self.tri_indices.max() >= self.num_vertices:
                return BVH_ERR_INCORRECT_DATA
            self.build_state = BVH_BUILD_STATE_PROCESSED
            return BVH_OK
        h = C.c_void_p()
        rc = _capi.lib().fclgpu_bvh_build_obbrss(addr(self.vertices), self.num_vertices, addr(self.tri_indices),
                                                 self.num_tris, self.split_method, C.byref(h))
        if rc != 0:
            return rc
        self._bvh = h
        self.build_state = BVH_BUILD_STATE_PROCESSED
        return BVH_OK

    # -- replace protocol (BVH_model-inl.h:521-620) --
    def beginReplaceModel(self):
        if self.build_state != BVH_BUILD_STATE_PROCESSED:
            sys.stderr.write("BVH Error! Call beginReplaceModel() on a BVHModel that has no previous frame.\n")
            return BVH_ERR_BUILD_EMPTY_PREVIOUS_FRAME
        self._replace = []
        self.num_vertex_updated = 0
        self.build_state = BVH_BUILD_STATE_REPLACE_BEGUN
        return BVH_OK

    def replaceSubModel(self, ps):
        if self.build_state != BVH_BUILD_STATE_REPLACE_BEGUN:
            sys.stderr.write("BVH Warning! Call replaceSubModel() in a wrong order. replaceSubModel() was ignored. Must do "
                             "a beginReplaceModel() for initialization.\n")
            return BVH_ERR_BUILD_OUT_OF_SEQUENCE
        ps = np.asarray(ps, dtype=np.float64).reshape(-1, 3)
        self._replace.append(ps)
        self.num_vertex_updated += len(ps)
        return BVH_OK

    def replaceVertex(self, p):
        if self.build_state != BVH_BUILD_STATE_REPLACE_BEGUN:
            sys.stderr.write("BVH Warning! Call replaceVertex() in a wrong order. replaceVertex() was ignored. Must do a "
                             "beginReplaceModel() for initialization.\n")
            return BVH_ERR_BUILD_OUT_OF_SEQUENCE
        return self.replaceSubModel(np.asarray(p, dtype=np.float64).reshape(1, 3))

    def replaceTriangle(self, p1, p2, p3):
        if self.build_state != BVH_BUILD_STATE_REPLACE_BEGUN:
            sys.stderr.write("BVH Warning! Call replaceTriangle() in a wrong order. replaceTriangle() was ignored. Must do a "
                             "beginReplaceModel() for initialization.\n")
            return BVH_ERR_BUILD_OUT_OF_SEQUENCE
        return self.replaceSubModel(np.array([p1, p2, p3], dtype=np.float64))

    def endReplaceModel(self, refit=True, bottomup=True):
        """refit=True, bottomup=True (the reference's default, BVH_model.h:128): bottom-up refit (refitTree_bottomup,
        BVH_model-inl.h:952-1037: triangle fit at the leaves, OBB / RSS merging above) on the host copy AND, by a kernel,
        on every uploaded device copy (bit-identical BVs, the reference's merge quirks included).
        refit=True, bottomup=False: top-down refit (refitTree_topdown), likewise.
        refit=False: rebuild the tree (buildTree) and re-upload."""
        if self.build_state != BVH_BUILD_STATE_REPLACE_BEGUN:
            sys.stderr.write("BVH Warning! Call endReplaceModel() in a wrong order. endReplaceModel() was ignored. \n")
            return BVH_ERR_BUILD_OUT_OF_SEQUENCE
        if self.num_vertex_updated != self.num_vertices:
            sys.stderr.write("BVH Error! The replaced model should have the same number of vertices as the old model.\n")
            return BVH_ERR_INCORRECT_DATA
        new_v = np.ascontiguousarray(np.concatenate(self._replace), dtype=np.float64)
        self.vertices = new_v
        L = _capi.lib()
        if refit:
            host_refit = L.fclgpu_bvh_refit_bottomup if bottomup else L.fclgpu_bvh_refit_topdown
            dev_refit = L.fclgpu_model_refit_bottomup if bottomup else L.fclgpu_model_refit_topdown
            if self._bvh is not None:
                rc = host_refit(self._bvh, addr(new_v), self.num_vertices)
                if rc != 0:
                    return rc
            for dev, h in self._dev.items():
                check(dev_refit(h, addr(new_v), self.num_vertices, 0, None))
                check(L.fclgpu_sync_status(int(dev), None))
            self._bottomup = bool(bottomup)
        else:
            self._release()
            self._bottomup = False
            if not self.build_on_device:
                h = C.c_void_p()
                rc = L.fclgpu_bvh_build_obbrss(addr(self.vertices), self.num_vertices, addr(self.tri_indices),
                                               self.num_tris, self.split_method, C.byref(h))
                if rc != 0:
                    return rc
                self._bvh = h
        self.build_state = BVH_BUILD_STATE_PROCESSED
        return BVH_OK

    def refit_device(self, vertices, device=None, stream=None, bottomup=False):
        """Device-resident update: `vertices` is a CUDA float64 tensor (num_vertices, 3); only the device copy
        on that GPU is refitted (asynchronous on `stream`); the host copy is left untouched."""
        torch = _torch()
        dev = vertices.device.index if device is None else device
        st = (stream or torch.cuda.current_stream(dev)).cuda_stream
        L = _capi.lib()
        fn = L.fclgpu_model_refit_bottomup if bottomup else L.fclgpu_model_refit_topdown
        check(fn(self.device_model(dev), addr(vertices), self.num_vertices, 1, st))

    def download_device_arrays(self, device=None):
        """FP64 node records as they currently are in HBM (for tests / inspection)."""
        h = self.device_model(device)
        n, nt = self.getNumBVs(), self.num_tris
        out = dict(axis=np.empty((n, 9)), obb_To=np.empty((n, 3)), obb_ext=np.empty((n, 3)), rss_To=np.empty((n, 3)),
                   rss_l=np.empty((n, 2)), rss_r=np.empty(n), tri_verts=np.empty((nt, 9)))
        check(_capi.lib().fclgpu_model_download(h, addr(out["axis"]), addr(out["obb_To"]), addr(out["obb_ext"]),
                                                addr(out["rss_To"]), addr(out["rss_l"]), addr(out["rss_r"]),
                                                addr(out["tri_verts"])))
        out["rss_axis"] = np.empty((n, 9))
        check(_capi.lib().fclgpu_model_download_rss_axis(h, addr(out["rss_axis"])))
        return out

    def partition(self):
        n, nt = self.getNumBVs(), self.num_tris
        fp, npr, pi = np.empty(n, np.int32), np.empty(n, np.int32), np.empty(nt, np.int32)
        if self._bvh is None:
            check(_capi.lib().fclgpu_model_get_topology(self.device_model(), None, addr(fp), addr(npr), addr(pi)))
        else:
            check(_capi.lib().fclgpu_bvh_get_partition(self._bvh, addr(fp), addr(npr), addr(pi), None))
        return fp, npr, pi

    @classmethod
    def from_arrays(cls, verts, tris, split_method=SPLIT_METHOD_MEAN, build_on_device=False):
        m = cls(split_method, build_on_device)
        m.beginModel()
        m.addSubModel(verts, tris)
        rc = m.endModel()
        if rc != BVH_OK:
            raise FclGpuError(rc, "endModel failed")
        return m

    @classmethod
    def from_obj(cls, path, split_method=SPLIT_METHOD_MEAN, build_on_device=False):
        """The reference's test set-up: loadOBJFile + beginModel / addSubModel / endModel (test_fcl_collision.cpp:792-815)."""
        verts, tris = loadOBJFile(path)
        return cls.from_arrays(verts, tris, split_method, build_on_device)

    def getNumBVs(self):
        if self._bvh is None:
            return 2 * self.num_tris - 1 if (self.build_on_device and self.build_state == BVH_BUILD_STATE_PROCESSED) else 0
        return int(_capi.lib().fclgpu_bvh_num_nodes(self._bvh))

    def node_arrays(self):
        """The flattened node tree as numpy arrays (what the upload step sends to HBM)."""
        n, nt = self.getNumBVs(), self.num_tris
        if self._bvh is None:  # built on the device: read the records back
            out = self.download_device_arrays()
            out["first_child"] = np.empty(n, np.int32)
            check(_capi.lib().fclgpu_model_get_topology(self.device_model(), addr(out["first_child"]), None, None, None))
            return out
        out = dict(first_child=np.empty(n, np.int32), axis=np.empty((n, 9)), obb_To=np.empty((n, 3)),
                   obb_ext=np.empty((n, 3)), rss_To=np.empty((n, 3)), rss_l=np.empty((n, 2)), rss_r=np.empty(n),
                   tri_verts=np.empty((nt, 9)))
        check(_capi.lib().fclgpu_bvh_get(self._bvh, addr(out["first_child"]), addr(out["axis"]), addr(out["obb_To"]),
                                         addr(out["obb_ext"]), addr(out["rss_To"]), addr(out["rss_l"]),
                                         addr(out["rss_r"]), addr(out["tri_verts"])))
        out["rss_axis"] = np.empty((n, 9))  # = axis unless the model was refitted bottom-up
        rc = _capi.lib().fclgpu_bvh_get_rss_axis(self._bvh, addr(out["rss_axis"]))
        if rc < 0:
            check(rc)
        return out

    def device_model(self, device=None):
        """Upload (once per device) and return the fclgpu_model handle."""
        if self.build_state != BVH_BUILD_STATE_PROCESSED:
            raise FclGpuError(BVH_ERR_BUILD_OUT_OF_SEQUENCE, "model is not built (call endModel())")
        if device is None:
            device = _current_device()
        h = self._dev.get(device)
        if h is None:
            h = C.c_void_p()
            if self._bvh is None:
                check(_capi.lib().fclgpu_model_build_obbrss(int(device), addr(self.vertices), self.num_vertices,
                                                            addr(self.tri_indices), self.num_tris, self.split_method,
                                                            C.byref(h)))
                if getattr(self, "_bottomup", False):  # a copy made after a bottom-up refit of a device-built model
                    check(_capi.lib().fclgpu_model_refit_bottomup(h, addr(self.vertices), self.num_vertices, 0, None))
                    check(_capi.lib().fclgpu_sync_status(int(device), None))
            else:
                check(_capi.lib().fclgpu_model_from_bvh(int(device), self._bvh, C.byref(h)))
            self._dev[device] = h
        return h

    def _release(self):
        L = _capi.lib()
        for h in self._dev.values():
            L.fclgpu_model_destroy(h)
        self._dev = {}
        if self._bvh is not None:
            L.fclgpu_bvh_destroy(self._bvh)
            self._bvh = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass


def _current_device():
    try:
        torch = _torch()
        if torch.cuda.is_available():
            return torch.cuda.current_device()
    except ImportError:
        pass
    return 0


class Sphere:
    """fcl::Sphere<double> (geometry/shape/sphere.h): centred at the origin of its own frame.  On this path it is
    the second geometry of a mesh <-> sphere collide / distance."""

    def __init__(self, radius):
        self.radius = float(radius)
        self.cost_density = 1.0
        self.threshold_occupied = 1.0
        self.threshold_free = 0.0

    def getObjectType(self):
        return "OT_GEOM"

    def getNodeType(self):
        return "GEOM_SPHERE"


class _PlaneShape:
    """n . x <= d (Halfspace) / n . x = d (Plane) in the shape's own frame; the constructor normalises like the
    reference's unitNormalTest (geometry/shape/halfspace-inl.h:144-160, plane-inl.h:144-160)."""

    def __init__(self, n, d=0.0, *rest):
        if rest:  # Halfspace(a, b, c, d)
            n, d = (n, d, rest[0]), rest[1]
        n = np.asarray(n, np.float64).reshape(3)
        l = float(np.sqrt((n[0] * n[0] + n[1] * n[1]) + n[2] * n[2]))
        if l > 0:
            inv_l = 1.0 / l
            self.n, self.d = n * inv_l, float(d) * inv_l
        else:
            self.n, self.d = np.array([1.0, 0.0, 0.0]), 0.0
        self.cost_density = 1.0
        self.threshold_occupied = 1.0
        self.threshold_free = 0.0

    def signedDistance(self, p):
        p = np.asarray(p, np.float64)
        return float((self.n[0] * p[0] + self.n[1] * p[1]) + self.n[2] * p[2]) - self.d

    def getObjectType(self):
        return "OT_GEOM"


class Halfspace(_PlaneShape):
    """fcl::Halfspace<double> (geometry/shape/halfspace.h): second geometry of a mesh <-> halfspace collide."""

    _kind = _capi.SHAPE_HALFSPACE

    def getNodeType(self):
        return "GEOM_HALFSPACE"


class Plane(_PlaneShape):
    """fcl::Plane<double> (geometry/shape/plane.h): second geometry of a mesh <-> plane collide."""

    _kind = _capi.SHAPE_PLANE

    def getNodeType(self):
        return "GEOM_PLANE"

    def distance(self, p):
        return abs(self.signedDistance(p))


class CollisionObject:
    """fcl::CollisionObject<double> for BVHModel geometries (narrowphase/collision_object.h): a geometry
    plus a transform; collide(o1, o2, request, result) / distance(o1, o2, request, result) forward the
    geometry and getTransform() exactly like collision-inl.h:81-91,154-175."""

    def __init__(self, cgeom, tf=None):
        self.cgeom = cgeom
        self.t = Transform3() if tf is None else tf
        self.user_data = None

    def collisionGeometry(self):
        return self.cgeom

    def getTransform(self):
        return self.t

    def setTransform(self, R_or_tf, T=None):
        self.t = R_or_tf if isinstance(R_or_tf, Transform3) else Transform3(R_or_tf, T)

    def getRotation(self):
        return self.t.R

    def getTranslation(self):
        return self.t.t

    def setRotation(self, R):
        self.t = Transform3(R, self.t.t)

    def setTranslation(self, T):
        self.t = Transform3(self.t.R, T)

    def getObjectType(self):
        return self.cgeom.getObjectType()

    def getNodeType(self):
        return self.cgeom.getNodeType()


class CollisionRequest:
    def __init__(self, num_max_contacts=1, enable_contact=False, num_max_cost_sources=1, enable_cost=False,
                 use_approximate_cost=True, gjk_solver_type="GST_LIBCCD", gjk_tolerance=1e-6):
        self.num_max_contacts = num_max_contacts
        self.enable_contact = enable_contact
        self.num_max_cost_sources = num_max_cost_sources
        self.enable_cost = enable_cost
        self.use_approximate_cost = use_approximate_cost
        self.gjk_solver_type = gjk_solver_type
        self.gjk_tolerance = gjk_tolerance

    def isSatisfied(self, result):
        return (not self.enable_cost) and result.isCollision() and self.num_max_contacts <= result.numContacts()

    def _c(self, stage_capacity=0, contact_format=0):
        return _capi.CollisionRequestC(int(min(self.num_max_contacts, 2**62)), int(bool(self.enable_contact)),
                                       int(bool(self.enable_cost)), int(stage_capacity), int(contact_format), 0)


class Contact:
    __slots__ = ("o1", "o2", "b1", "b2", "normal", "pos", "penetration_depth")

    def __init__(self, o1=None, o2=None, b1=-1, b2=-1, pos=None, normal=None, depth=0.0):
        self.o1, self.o2, self.b1, self.b2 = o1, o2, b1, b2
        self.pos, self.normal, self.penetration_depth = pos, normal, depth

    def __lt__(self, other):  # contact-inl.h:98-104
        if self.b1 == other.b1:
            return self.b2 < other.b2
        return self.b1 < other.b1


class CollisionResult:
    def __init__(self):
        self.contacts = []

    def addContact(self, c):
        self.contacts.append(c)

    def isCollision(self):
        return len(self.contacts) > 0

    def numContacts(self):
        return len(self.contacts)

    def getContact(self, i):
        return self.contacts[i] if i < len(self.contacts) else self.contacts[-1]

    def getContacts(self):
        return list(self.contacts)

    def clear(self):
        self.contacts = []


class DistanceRequest:
    def __init__(self, enable_nearest_points=False, enable_signed_distance=False, rel_err=0.0, abs_err=0.0,
                 distance_tolerance=1e-6, gjk_solver_type="GST_LIBCCD"):
        self.enable_nearest_points = enable_nearest_points
        self.enable_signed_distance = enable_signed_distance
        self.rel_err = rel_err  # ignored on this path, like the reference (see include/fclgpu.h)
        self.abs_err = abs_err
        self.distance_tolerance = distance_tolerance
        self.gjk_solver_type = gjk_solver_type

    def isSatisfied(self, result):
        return result.min_distance <= 0

    def _c(self):
        return _capi.DistanceRequestC(int(bool(self.enable_nearest_points)), int(bool(self.enable_signed_distance)),
                                      float(self.rel_err), float(self.abs_err))


class DistanceResult:
    def __init__(self, min_distance=DBL_MAX):
        self.min_distance = min_distance
        self.nearest_points = [np.zeros(3), np.zeros(3)]
        self.o1 = self.o2 = None
        self.b1 = self.b2 = -1  # DistanceResult::NONE

    def update(self, distance, o1, o2, b1, b2, p1=None, p2=None):  # distance_result-inl.h:66-103
        if self.min_distance > distance:
            self.min_distance = distance
            self.o1, self.o2, self.b1, self.b2 = o1, o2, b1, b2
            if p1 is not None:
                self.nearest_points = [np.array(p1), np.array(p2)]

    def clear(self):
        self.__init__()


# ------------------------------------------------------------------------------------------------
# batched entry points (the product)
# ------------------------------------------------------------------------------------------------
_PINNED_POOL = {}


def _out(shape, dtype, pinned, tag=""):
    """Output buffer: plain numpy, or page-locked (pinned) host memory so that the device->host copy is
    asynchronous and overlaps the next chunk's kernel.  Pinned buffers come from a small pool keyed by
    (tag, size) and are REUSED by the next call with the same shapes (page-locking gigabytes per call
    would cost more than the copy), so a pinned result is valid until that next call.
    Returns (array, keepalive)."""
    if not pinned:
        return np.zeros(shape, dtype), None
    torch = _torch()
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    key = (tag, nbytes)
    t = _PINNED_POOL.get(key)
    if t is None:
        t = torch.empty(max(nbytes, 1), dtype=torch.uint8, pin_memory=True)
        _PINNED_POOL[key] = t
    return t.numpy()[:nbytes].view(dtype).reshape(shape), t


def loadOBJFile(path):
    """loadOBJFile (test/test_fcl_utility.h:194-280) through the C ABI: (vertices (nv, 3) float64, triangles (nt, 3) int32)."""
    L = _capi.lib()
    v, t = C.POINTER(C.c_double)(), C.POINTER(C.c_int32)()
    nv, nt = C.c_int32(0), C.c_int32(0)
    rc = L.fclgpu_load_obj(str(path).encode(), C.byref(v), C.byref(nv), C.byref(t), C.byref(nt))
    if rc == _capi.ERR_INCORRECT_DATA:
        sys.stderr.write("file not exist\n")  # the reference's message; it returns empty arrays
        return np.zeros((0, 3)), np.zeros((0, 3), np.int32)
    check(rc)
    try:
        verts = np.ctypeslib.as_array(v, shape=(nv.value, 3)).copy() if nv.value else np.zeros((0, 3))
        tris = np.ctypeslib.as_array(t, shape=(nt.value, 3)).copy() if nt.value else np.zeros((0, 3), np.int32)
    finally:
        L.fclgpu_free(v)
        L.fclgpu_free(t)
    return verts, tris


def saveOBJFile(path, verts, tris):
    """saveOBJFile (test/test_fcl_utility.h:283-309); coordinates with 17 significant digits (exact round trip)."""
    v = np.ascontiguousarray(verts, np.float64).reshape(-1, 3)
    t = np.ascontiguousarray(tris, np.int32).reshape(-1, 3)
    check(_capi.lib().fclgpu_save_obj(str(path).encode(), addr(v), len(v), addr(t), len(t)))


class BatchCollisionResult:
    """num_contacts[n] plus the contact list of every query.

    The library appends each query's contacts as one contiguous block to a dense pool; `starts[i]` is the first slot of
    query i's block (`starts[n]` = total), blocks come in the order the queries retire on the GPU unless the library
    option `contact_order` is 1 (include/fclgpu.h).  `contacts_of(i)` reads a block in place.  `contacts` / `offsets`
    present the same list in QUERY order (offsets = exclusive prefix sum of num_contacts, n+1 entries) -- a host-side
    gather done on first use, for callers (and the parity tests) that want one flat, deterministic array."""

    def __init__(self, num_contacts, pool, starts, n_bv=None, n_leaf=None):
        self.num_contacts = num_contacts
        self.pool = pool
        self.starts = starts
        self.n_bv = n_bv
        self.n_leaf = n_leaf
        self._ordered = None

    def contacts_of(self, i):
        return self.pool[self.starts[i]:self.starts[i] + self.num_contacts[i]]

    def _order(self):
        if self._ordered is None and self.pool is not None:
            n = len(self.num_contacts)
            cnt = self.num_contacts.astype(np.int64)
            off = np.zeros(n + 1, np.int64)
            np.cumsum(cnt, out=off[1:])
            st = np.asarray(self.starts[:n], np.int64)
            if off[n] != self.starts[n]:  # truncated by a capacity overflow: keep the library's layout
                self._ordered = (self.pool, self.starts)
            elif np.array_equal(st[cnt > 0], off[:n][cnt > 0]):
                self._ordered = (self.pool, off)
            else:
                idx = np.repeat(st - off[:n], cnt) + np.arange(off[n], dtype=np.int64)
                self._ordered = (self.pool[idx], off)
        return self._ordered

    @property
    def contacts(self):
        return None if self.pool is None else self._order()[0]

    @property
    def offsets(self):
        return None if self.pool is None else self._order()[1]


class BatchDistanceResult:
    def __init__(self, min_distance, p1, p2, b1, b2, n_bv=None, n_leaf=None):
        self.min_distance, self.nearest_p1, self.nearest_p2, self.b1, self.b2 = min_distance, p1, p2, b1, b2
        self.n_bv, self.n_leaf = n_bv, n_leaf


def collide_batch(o1, tf1, o2, tf2, request, contact_capacity=None, want_contacts=True, stats=False, device=None,
                  grow_on_overflow=False, pinned=False, stage_capacity=0, contact_format=CONTACT_FULL):
    """Host arrays in, host arrays out (copies inside): n independent fcl::collide() calls.

    tf1 / tf2: (n,12) float64 pose records, a Transform3, or None (identity).
    contact_format (extension): CONTACT_FULL = the reference's 64-byte contacts; CONTACT_IDS = the two primitive ids only
    (8 bytes); CONTACT_F32 = ids + single-precision normal / position / depth (40 bytes) -- same lists, same order."""
    tf1, n1 = _poses(tf1)
    tf2, n2 = _poses(tf2)
    n = n1 if n1 is not None else n2
    if n is None:
        raise ValueError("at least one of tf1/tf2 must be given")
    if n1 is not None and n2 is not None and n1 != n2:
        raise ValueError("tf1 and tf2 must have the same length")
    m1, m2 = o1.device_model(device), o2.device_model(device)
    req = request._c(stage_capacity, contact_format)
    keep = []
    counts, k = _out(n, np.int32, pinned, "counts")
    keep.append(k)
    if want_contacts:
        if contact_capacity is None:
            contact_capacity = int(min(max(request.num_max_contacts, 0), 64)) * n
        contact_capacity = max(int(contact_capacity), 1)
        contacts, k = _out(contact_capacity, CONTACT_DTYPES[contact_format], pinned, "contacts%d" % contact_format)
        keep.append(k)
        offsets = np.zeros(n + 1, np.int64)
    else:
        contact_capacity, contacts, offsets = 0, None, None
    n_bv = np.zeros(n, np.uint32) if stats else None
    n_leaf = np.zeros(n, np.uint32) if stats else None
    rc = _capi.lib().fclgpu_collide_batch_host(m1, m2, n, addr(tf1), addr(tf2), C.byref(req), addr(counts),
                                               addr(contacts), contact_capacity, addr(offsets), addr(n_bv),
                                               addr(n_leaf))
    if rc == _capi.ERR_CONTACT_OVERFLOW and grow_on_overflow:
        # counts are exact even when the pool / the per-query staging was too small: size both (per call) and rerun
        return collide_batch(o1, tf1, o2, tf2, request, contact_capacity=int(counts.sum(dtype=np.int64)),
                             want_contacts=True, stats=stats, device=device, grow_on_overflow=False,
                             stage_capacity=max(int(counts.max()), 1), contact_format=contact_format)
    check(rc)
    if want_contacts:
        contacts = contacts[: offsets[n]]
    res = BatchCollisionResult(counts, contacts, offsets, n_bv, n_leaf)
    res._keepalive = keep
    return res


def collide_mesh_sphere_batch(o1, tf1, sphere, tf2, request, contact_capacity=None, want_contacts=True, stats=False,
                              device=None, grow_on_overflow=False, stage_capacity=0):
    """n independent fcl::collide(mesh, tf1[i], Sphere, tf2[i]) calls (host arrays in and out).  Contacts: one per
    intersecting triangle, b2 = -1 (Contact::NONE), in the reference's traversal order."""
    tf1, n1 = _poses(tf1)
    tf2, n2 = _poses(tf2)
    n = n1 if n1 is not None else n2
    if n is None:
        raise ValueError("at least one of tf1/tf2 must be given")
    if n1 is not None and n2 is not None and n1 != n2:
        raise ValueError("tf1 and tf2 must have the same length")
    m1 = o1.device_model(device)
    req = request._c(stage_capacity)
    counts = np.zeros(n, np.int32)
    if want_contacts:
        if contact_capacity is None:
            contact_capacity = int(min(max(request.num_max_contacts, 0), 64)) * n
        contact_capacity = max(int(contact_capacity), 1)
        contacts = np.zeros(contact_capacity, CONTACT_DTYPE)
        offsets = np.zeros(n + 1, np.int64)
    else:
        contact_capacity, contacts, offsets = 0, None, None
    n_bv = np.zeros(n, np.uint32) if stats else None
    n_leaf = np.zeros(n, np.uint32) if stats else None
    rc = _capi.lib().fclgpu_collide_mesh_sphere_batch_host(m1, float(sphere.radius), n, addr(tf1), addr(tf2), C.byref(req),
                                                           addr(counts), addr(contacts), contact_capacity, addr(offsets),
                                                           addr(n_bv), addr(n_leaf))
    if rc == _capi.ERR_CONTACT_OVERFLOW and grow_on_overflow:
        return collide_mesh_sphere_batch(o1, tf1, sphere, tf2, request, contact_capacity=int(counts.sum(dtype=np.int64)),
                                         want_contacts=True, stats=stats, device=device, grow_on_overflow=False,
                                         stage_capacity=max(int(counts.max()), 1))
    check(rc)
    if want_contacts:
        contacts = contacts[: offsets[n]]
    return BatchCollisionResult(counts, contacts, offsets, n_bv, n_leaf)


def collide_mesh_plane_batch(o1, tf1, shape, tf2, request, contact_capacity=None, want_contacts=True, stats=False,
                             device=None, grow_on_overflow=False, stage_capacity=0):
    """n independent fcl::collide(mesh, tf1[i], Halfspace | Plane, tf2[i]) calls (host arrays in and out).  Contacts: one
    per intersecting triangle, b2 = -1 (Contact::NONE), in the reference's traversal order."""
    tf1, n1 = _poses(tf1)
    tf2, n2 = _poses(tf2)
    n = n1 if n1 is not None else n2
    if n is None:
        raise ValueError("at least one of tf1/tf2 must be given")
    if n1 is not None and n2 is not None and n1 != n2:
        raise ValueError("tf1 and tf2 must have the same length")
    m1 = o1.device_model(device)
    req = request._c(stage_capacity)
    counts = np.zeros(n, np.int32)
    if want_contacts:
        if contact_capacity is None:
            contact_capacity = int(min(max(request.num_max_contacts, 0), 64)) * n
        contact_capacity = max(int(contact_capacity), 1)
        contacts = np.zeros(contact_capacity, CONTACT_DTYPE)
        offsets = np.zeros(n + 1, np.int64)
    else:
        contact_capacity, contacts, offsets = 0, None, None
    n_bv = np.zeros(n, np.uint32) if stats else None
    n_leaf = np.zeros(n, np.uint32) if stats else None
    nrm = np.ascontiguousarray(shape.n, np.float64)
    rc = _capi.lib().fclgpu_collide_mesh_plane_batch_host(m1, shape._kind, addr(nrm), float(shape.d), n, addr(tf1), addr(tf2),
                                                          C.byref(req), addr(counts), addr(contacts), contact_capacity,
                                                          addr(offsets), addr(n_bv), addr(n_leaf))
    if rc == _capi.ERR_CONTACT_OVERFLOW and grow_on_overflow:
        return collide_mesh_plane_batch(o1, tf1, shape, tf2, request, contact_capacity=int(counts.sum(dtype=np.int64)),
                                        want_contacts=True, stats=stats, device=device, grow_on_overflow=False,
                                        stage_capacity=max(int(counts.max()), 1))
    check(rc)
    if want_contacts:
        contacts = contacts[: offsets[n]]
    return BatchCollisionResult(counts, contacts, offsets, n_bv, n_leaf)


def distance_batch(o1, tf1, o2, tf2, request, stats=False, device=None, pinned=False, cutoff=None):
    """n independent fcl::distance() calls (host arrays in and out).  cutoff (extension, see include/fclgpu.h): the
    traversal starts from min_distance = cutoff, so results >= cutoff come back as cutoff with b1 = b2 = -1."""
    tf1, n1 = _poses(tf1)
    tf2, n2 = _poses(tf2)
    n = n1 if n1 is not None else n2
    if n is None:
        raise ValueError("at least one of tf1/tf2 must be given")
    if n1 is not None and n2 is not None and n1 != n2:
        raise ValueError("tf1 and tf2 must have the same length")
    m1, m2 = o1.device_model(device), o2.device_model(device)
    req = request._c()
    (dist, k0), (p1, k1), (p2, k2) = _out(n, np.float64, pinned, "dist"), _out((n, 3), np.float64, pinned, "p1"), _out((n, 3), np.float64, pinned, "p2")
    (b1, k3), (b2, k4) = _out(n, np.int32, pinned, "b1"), _out(n, np.int32, pinned, "b2")
    n_bv = np.zeros(n, np.uint32) if stats else None
    n_leaf = np.zeros(n, np.uint32) if stats else None
    if cutoff is None:
        check(_capi.lib().fclgpu_distance_batch_host(m1, m2, n, addr(tf1), addr(tf2), C.byref(req), addr(dist), addr(p1),
                                                     addr(p2), addr(b1), addr(b2), addr(n_bv), addr(n_leaf)))
    else:
        check(_capi.lib().fclgpu_distance_cutoff_batch_host(m1, m2, n, addr(tf1), addr(tf2), C.byref(req), float(cutoff),
                                                            addr(dist), addr(p1), addr(p2), addr(b1), addr(b2), addr(n_bv),
                                                            addr(n_leaf)))
    res = BatchDistanceResult(dist, p1, p2, b1, b2, n_bv, n_leaf)
    res._keepalive = [k0, k1, k2, k3, k4]
    return res


def within_tolerance_batch(o1, tf1, o2, tf2, tolerance, stats=False, device=None, early_exit=True, pinned=False):
    """Tolerance verification (extension; BASELINE cfg5): bool[n], query i is True iff fcl::distance(o1, tf1[i], o2,
    tf2[i]) <= tolerance.  Node pairs farther apart than the tolerance are pruned from the first round on and (with
    early_exit, the default: fclgpu_within_tolerance_batch) a query ends at the first triangle pair found within the
    tolerance.  Returns (within, BatchDistanceResult); with early_exit the result's min_distance is the witness pair's
    distance (an upper bound of the true distance), without it min(fcl::distance, nextafter(tolerance))."""
    if not early_exit:
        r = distance_batch(o1, tf1, o2, tf2, DistanceRequest(False), stats=stats, device=device, pinned=pinned,
                           cutoff=float(np.nextafter(float(tolerance), np.inf)))
        return r.min_distance <= float(tolerance), r
    tf1, n1 = _poses(tf1)
    tf2, n2 = _poses(tf2)
    n = n1 if n1 is not None else n2
    if n is None:
        raise ValueError("at least one of tf1/tf2 must be given")
    if n1 is not None and n2 is not None and n1 != n2:
        raise ValueError("tf1 and tf2 must have the same length")
    m1, m2 = o1.device_model(device), o2.device_model(device)
    (within, k0), (dist, k1) = _out(n, np.uint8, pinned, "within"), _out(n, np.float64, pinned, "wdist")  # pinned: see _out
    n_bv = np.zeros(n, np.uint32) if stats else None
    n_leaf = np.zeros(n, np.uint32) if stats else None
    check(_capi.lib().fclgpu_within_tolerance_batch_host(m1, m2, n, addr(tf1), addr(tf2), float(tolerance), addr(within),
                                                         addr(dist), addr(n_bv), addr(n_leaf)))
    res = BatchDistanceResult(dist, None, None, None, None, n_bv, n_leaf)
    res._keepalive = (k0, k1)
    return within.astype(bool), res


def distance_mesh_sphere_batch(o1, tf1, sphere, tf2, request, stats=False, device=None, pinned=False):
    """n independent fcl::distance(mesh, tf1[i], Sphere, tf2[i]) calls (host arrays in and out).  nearest_p1 is in the
    mesh frame and nearest_p2 in the sphere frame (the reference's postprocess is empty for this node), b1 = closest
    triangle, b2 = -1 (DistanceResult::NONE).  Centre within the radius of a triangle: min_distance = -1, NaN points."""
    tf1, n1 = _poses(tf1)
    tf2, n2 = _poses(tf2)
    n = n1 if n1 is not None else n2
    if n is None:
        raise ValueError("at least one of tf1/tf2 must be given")
    if n1 is not None and n2 is not None and n1 != n2:
        raise ValueError("tf1 and tf2 must have the same length")
    m1 = o1.device_model(device)
    req = request._c()
    (dist, k0), (p1, k1), (p2, k2) = _out(n, np.float64, pinned, "dist"), _out((n, 3), np.float64, pinned, "p1"), _out((n, 3), np.float64, pinned, "p2")
    (b1, k3), (b2, k4) = _out(n, np.int32, pinned, "b1"), _out(n, np.int32, pinned, "b2")
    n_bv = np.zeros(n, np.uint32) if stats else None
    n_leaf = np.zeros(n, np.uint32) if stats else None
    check(_capi.lib().fclgpu_distance_mesh_sphere_batch_host(m1, float(sphere.radius), n, addr(tf1), addr(tf2), C.byref(req),
                                                             addr(dist), addr(p1), addr(p2), addr(b1), addr(b2),
                                                             addr(n_bv), addr(n_leaf)))
    res = BatchDistanceResult(dist, p1, p2, b1, b2, n_bv, n_leaf)
    res._keepalive = [k0, k1, k2, k3, k4]
    return res


def collide_batch_device(o1, tf1, o2, tf2, request, num_contacts, contacts=None, contact_offsets=None, n_bv=None,
                         n_leaf=None, stream=None):
    """Device-resident variant: every argument is a CUDA torch tensor (or None); asynchronous on
    `stream` (default: torch's current stream).  contacts: uint8/any tensor of capacity*64 bytes."""
    torch = _torch()
    n = (tf1 if tf1 is not None else tf2).numel() // 12
    dev = num_contacts.device.index
    m1, m2 = o1.device_model(dev), o2.device_model(dev)
    req = request._c()
    cap = 0 if contacts is None else (contacts.numel() * contacts.element_size()) // 64
    st = (stream or torch.cuda.current_stream(dev)).cuda_stream
    check(_capi.lib().fclgpu_collide_batch(m1, m2, n, addr(tf1), addr(tf2), C.byref(req), addr(num_contacts),
                                           addr(contacts), cap, addr(contact_offsets), addr(n_bv), addr(n_leaf), st))


def distance_batch_device(o1, tf1, o2, tf2, request, min_distance, p1=None, p2=None, b1=None, b2=None, n_bv=None,
                          n_leaf=None, stream=None):
    torch = _torch()
    n = (tf1 if tf1 is not None else tf2).numel() // 12
    dev = min_distance.device.index
    m1, m2 = o1.device_model(dev), o2.device_model(dev)
    req = request._c()
    st = (stream or torch.cuda.current_stream(dev)).cuda_stream
    check(_capi.lib().fclgpu_distance_batch(m1, m2, n, addr(tf1), addr(tf2), C.byref(req), addr(min_distance),
                                            addr(p1), addr(p2), addr(b1), addr(b2), addr(n_bv), addr(n_leaf), st))


def sync_status(device=None, stream=None):
    torch = _torch()
    dev = _current_device() if device is None else device
    st = (stream or torch.cuda.current_stream(dev)).cuda_stream
    check(_capi.lib().fclgpu_sync_status(dev, st))


def trim_device(device=None):
    """Give the per-device workspace buffers (contact staging, host-API staging, overflow areas of the distance front) back to
    the device after it has gone idle; models stay.  Returns the number of bytes released (fclgpu_device_trim)."""
    dev = _current_device() if device is None else device
    n = C.c_int64(0)
    check(_capi.lib().fclgpu_device_trim(dev, C.byref(n)))
    _PINNED_POOL.clear()  # the page-locked result buffers of the pinned=True calls go as well
    return int(n.value)


# ------------------------------------------------------------------------------------------------
# broadphase
# ------------------------------------------------------------------------------------------------
class DefaultCollisionData:
    """fcl::DefaultCollisionData (broadphase/default_broadphase_callbacks.h:59-70)."""

    def __init__(self, request=None):
        self.request = request if request is not None else CollisionRequest()
        self.result = CollisionResult()
        self.done = False


def DefaultCollisionFunction(o1, o2, data):
    """fcl::DefaultCollisionFunction (default_broadphase_callbacks.h:84-103)."""
    if data.done:
        return True
    collide(o1, o2, data.request, data.result)
    if (not data.request.enable_cost) and data.result.isCollision() and data.result.numContacts() >= data.request.num_max_contacts:
        data.done = True
    return data.done


class DefaultDistanceData:
    """fcl::DefaultDistanceData (broadphase/default_broadphase_callbacks.h:160-167)."""

    def __init__(self, request=None):
        self.request = request if request is not None else DistanceRequest()
        self.result = DistanceResult()
        self.done = False


def DefaultDistanceFunction(o1, o2, data, dist):
    """fcl::DefaultDistanceFunction (default_broadphase_callbacks.h:190-211).  `dist` stands for the reference's `S& dist`:
    a one-element list the callback writes the current minimum into."""
    if data.done:
        dist[0] = data.result.min_distance
        return True
    distance(o1, o2, data.request, data.result)
    dist[0] = data.result.min_distance
    if dist[0] <= 0:
        return True  # in collision or in touch
    return data.done


def _aabb_distance_matrix(a, b):
    """AABB::distance (math/bv/AABB-inl.h:292-315) for every pair of rows of a (n1, 6) and b (n2, 6): {min3, max3}."""
    res = np.zeros((len(a), len(b)))
    for k in range(3):
        amin, amax = a[:, None, k], a[:, None, 3 + k]
        bmin, bmax = b[None, :, k], b[None, :, 3 + k]
        d1 = bmax - amin
        d2 = amax - bmin
        res = res + np.where(amin > bmax, d1 * d1, np.where(bmin > amax, d2 * d2, 0.0))
    return np.sqrt(res)


class BatchBroadPhaseDistance:
    """min_distance over all pairs (o1 in this manager, o2 in the other), the pair (i, j) that attains it, its nearest points
    (world frame) and closest primitive ids, and how many of the n1 x n2 pairs needed an exact query."""

    def __init__(self, min_distance, pair, nearest_points, ids, evaluated):
        self.min_distance, self.pair, self.nearest_points, self.ids, self.evaluated = min_distance, pair, nearest_points, ids, evaluated


class BatchBroadPhaseResult:
    """pairs (m, 2): (index in this manager, index in the other manager) of every pair whose world AABBs overlap, in the
    brute-force manager's visiting order; num_contacts[m]: fcl::collide on each pair with a fresh result (or None)."""

    def __init__(self, pairs, num_contacts, aabb1, aabb2):
        self.pairs, self.num_contacts, self.aabb1, self.aabb2 = pairs, num_contacts, aabb1, aabb2


class NaiveCollisionManager:
    """fcl::NaiveCollisionManager (broadphase/broadphase_bruteforce.h) over CollisionObjects whose geometries are
    BVHModel<OBBRSS>: registerObject(s) / setup / update / clear / size / getObjects, collide(other, cdata, callback)
    with the reference's callback protocol, and the batched form collide_batch(other, request): culling AND the
    narrowphase of every culled pair on the GPU (fclgpu_broadphase_collide_host).  The dynamic AABB tree manager of the
    reference reports the same set of pairs; DynamicAABBTreeCollisionManager is an alias."""

    def __init__(self):
        self.objs = []

    def registerObject(self, obj):
        self.objs.append(obj)

    def registerObjects(self, objs):
        self.objs.extend(objs)

    def unregisterObject(self, obj):
        self.objs.remove(obj)

    def setup(self):
        pass

    def update(self, *args):
        pass

    def clear(self):
        self.objs = []

    def getObjects(self):
        return list(self.objs)

    def empty(self):
        return not self.objs

    def size(self):
        return len(self.objs)

    def _tables(self, other, device):
        geoms, index = [], {}
        for o in list(self.objs) + list(other.objs):
            g = o.collisionGeometry()
            if not isinstance(g, BVHModel):
                raise FclGpuError(BVH_ERR_UNSUPPORTED_FUNCTION, "the batched broadphase handles BVHModel<OBBRSS> geometries")
            if id(g) not in index:
                index[id(g)] = len(geoms)
                geoms.append(g)

        def side(objs):
            gi = np.array([index[id(o.collisionGeometry())] for o in objs], np.int32)
            tf = np.ascontiguousarray(np.stack([_poses(o.getTransform())[0][0] for o in objs])) if objs else np.zeros((0, 12))
            return gi, tf

        handles = (C.c_void_p * len(geoms))(*[g.device_model(device) for g in geoms])
        return geoms, handles, side(self.objs), side(other.objs)

    def collide_batch(self, other, request=None, device=None, pair_capacity=None, narrowphase=True):
        """All pairs (o1 in self, o2 in other) with overlapping AABBs, and numContacts of fcl::collide on each."""
        n1, n2 = len(self.objs), len(other.objs)
        if n1 == 0 or n2 == 0:
            return BatchBroadPhaseResult(np.zeros((0, 2), np.int32), np.zeros(0, np.int32), np.zeros((n1, 6)), np.zeros((n2, 6)))
        geoms, handles, (g1, tf1), (g2, tf2) = self._tables(other, device)
        req = (request if request is not None else CollisionRequest())._c()
        cap = int(pair_capacity) if pair_capacity is not None else max(1024, 4 * (n1 + n2))
        while True:
            pairs = np.zeros((cap, 2), np.int32)
            counts = np.zeros(cap, np.int32) if narrowphase else None
            a1, a2 = np.zeros((n1, 6)), np.zeros((n2, 6))
            m = C.c_int64(0)
            rc = _capi.lib().fclgpu_broadphase_collide_host(len(geoms), handles, n1, addr(g1), addr(tf1), n2, addr(g2), addr(tf2),
                                                            C.byref(req), cap, addr(pairs), C.byref(m), addr(counts), addr(a1), addr(a2))
            if rc == _capi.ERR_CONTACT_OVERFLOW and m.value > cap:
                cap = int(m.value)
                continue
            check(rc)
            k = int(m.value)
            return BatchBroadPhaseResult(pairs[:k], counts[:k] if narrowphase else None, a1, a2)

    def collide(self, other, cdata, callback=None):
        """collide(other_manager, cdata, callback) (broadphase_bruteforce-inl.h:182-205): the callback sees the culled pairs
        in the brute-force manager's order and may end the evaluation by returning True.  Culling runs on the GPU.
        collide(cdata, callback) -- or the same manager on both sides, :188-192 -- is the self-collision form (:140-160):
        every unordered pair of this manager's objects once, it1 before it2 in registration order."""
        if callback is None:  # collide(cdata, callback)
            other, cdata, callback = self, other, cdata
        if self.size() == 0 or other.size() == 0:
            return
        r = self.collide_batch(other, narrowphase=False)
        own = other is self
        for i, j in r.pairs:
            if own and j <= i:
                continue
            if callback(self.objs[i], other.objs[j], cdata):
                return

    def distance(self, other, cdata, callback=None):
        """distance(other_manager, cdata, callback) (broadphase_bruteforce-inl.h:208-233): pairs in registration order, a
        pair is handed to the callback only while the distance of its AABBs is below the running minimum the callback
        reports; distance(cdata, callback) is the self form (:163-178)."""
        if callback is None:
            other, cdata, callback = self, other, cdata
        if self.size() == 0 or other.size() == 0:
            return
        r = self.collide_batch(other, narrowphase=False)  # world AABBs as CollisionObject::computeAABB builds them
        D = _aabb_distance_matrix(r.aabb1, r.aabb2)
        own = other is self
        min_dist = [DBL_MAX]
        for i, o1 in enumerate(self.objs):
            for j, o2 in enumerate(other.objs):
                if own and j <= i:
                    continue
                if D[i, j] < min_dist[0]:
                    if callback(o1, o2, cdata, min_dist):
                        return

    def distance_batch(self, other, request=None, device=None, chunk=4096):
        """The minimum distance between any object of this manager and any object of the other one, batched: pairs in the
        order of their AABB distance (a lower bound), exact distance() queries on the GPU a chunk at a time (grouped by
        geometry pair), until the next AABB distance is no smaller than the minimum found.  The value is what
        distance(other, DefaultDistanceData, DefaultDistanceFunction) ends with; among exact ties another pair may be named."""
        n1, n2 = len(self.objs), len(other.objs)
        if n1 == 0 or n2 == 0:
            return BatchBroadPhaseDistance(DBL_MAX, (-1, -1), None, (-1, -1), 0)
        r = self.collide_batch(other, narrowphase=False, device=device)
        D = _aabb_distance_matrix(r.aabb1, r.aabb2)
        own = other is self
        if own:
            D[np.tril_indices(n1)] = np.inf
        flat = np.argsort(D, axis=None, kind="stable")
        req = request if request is not None else DistanceRequest(True)
        geoms, _, (g1, tf1), (g2, tf2) = self._tables(other, device)
        best = (DBL_MAX, (-1, -1), None, (-1, -1))
        pos = evaluated = 0
        while pos < len(flat):
            take = flat[pos:pos + chunk]
            take = take[D.ravel()[take] < best[0]]
            if len(take) == 0:
                break
            pos += chunk
            ii, jj = np.unravel_index(take, D.shape)
            key = g1[ii].astype(np.int64) * len(geoms) + g2[jj]
            for k in np.unique(key):
                sel = np.nonzero(key == k)[0]
                a, b = geoms[int(k) // len(geoms)], geoms[int(k) % len(geoms)]
                res = distance_batch(a, np.ascontiguousarray(tf1[ii[sel]]), b, np.ascontiguousarray(tf2[jj[sel]]), req, device=device)
                m = int(np.argmin(res.min_distance))
                evaluated += len(sel)
                if res.min_distance[m] < best[0]:
                    pts = (res.nearest_p1[m].copy(), res.nearest_p2[m].copy()) if req.enable_nearest_points else None
                    best = (float(res.min_distance[m]), (int(ii[sel[m]]), int(jj[sel[m]])), pts, (int(res.b1[m]), int(res.b2[m])))
        return BatchBroadPhaseDistance(best[0], best[1], best[2], best[3], evaluated)

    def self_pairs(self, request=None, device=None, narrowphase=True):
        """Batched self-collision: the pairs (i < j) of this manager's objects whose AABBs overlap, in the brute-force
        manager's order, with numContacts of fcl::collide on each."""
        r = self.collide_batch(self, request, device, narrowphase=narrowphase)
        keep = r.pairs[:, 0] < r.pairs[:, 1]
        return BatchBroadPhaseResult(r.pairs[keep], None if r.num_contacts is None else r.num_contacts[keep], r.aabb1, r.aabb2)


DynamicAABBTreeCollisionManager = NaiveCollisionManager


# ------------------------------------------------------------------------------------------------
# single-query entry points with the reference's signatures
# ------------------------------------------------------------------------------------------------
def collide(o1, tf1, o2=None, tf2=None, request=None, result=None):
    """fcl::collide(o1, tf1, o2, tf2, request, result) for two BVHModel<OBBRSS> (collision-inl.h:95-207),
    or the CollisionObject overload collide(obj1, obj2, request, result) (collision-inl.h:154-175).
    Appends to `result` (results accumulate across calls unless cleared) and returns numContacts()."""
    if isinstance(o1, CollisionObject):  # collide(obj1, obj2, request, result)
        obj1, obj2, request, result = o1, tf1, o2, tf2
        return collide(obj1.collisionGeometry(), obj1.getTransform(), obj2.collisionGeometry(), obj2.getTransform(),
                       request, result)
    if request.num_max_contacts == 0:
        sys.stderr.write(f"Warning: should stop early as num_max_contact is {request.num_max_contacts} !\n")
        return 0
    if isinstance(o1, (Sphere, _PlaneShape)) and isinstance(o2, BVHModel):
        # (OT_GEOM, OT_BVH): the reference calls the [BVH][GEOM] cell with the arguments swapped and does not flip
        # the contacts (collision-inl.h:124-134), so o1 of every contact is the mesh
        o1, tf1, o2, tf2 = o2, tf2, o1, tf1
    mesh_sphere = isinstance(o1, BVHModel) and isinstance(o2, Sphere)
    mesh_plane = isinstance(o1, BVHModel) and isinstance(o2, _PlaneShape)
    if not (isinstance(o1, BVHModel) and isinstance(o2, BVHModel)) and not mesh_sphere and not mesh_plane:
        sys.stderr.write("Warning: collision function between these node types is not supported\n")
        return 0
    if request.isSatisfied(result):  # orientedMeshCollide / orientedBVHShapeCollide, collision_func_matrix-inl.h:580,389
        return result.numContacts()
    if request.enable_cost:
        raise FclGpuError(BVH_ERR_UNSUPPORTED_FUNCTION, "cost sources are not supported on this path")
    # a non-empty result consumes part of the contact budget
    budget = request.num_max_contacts - result.numContacts()
    sub = CollisionRequest(budget, request.enable_contact)
    ident = np.array([[1.0, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0]])
    if mesh_sphere:
        r = collide_mesh_sphere_batch(o1, tf1, o2, tf2, sub, contact_capacity=min(budget, 256), grow_on_overflow=True)
    elif mesh_plane:
        r = collide_mesh_plane_batch(o1, ident if tf1 is None else tf1, o2, ident if tf2 is None else tf2, sub,
                                     contact_capacity=min(budget, 256), grow_on_overflow=True)
    else:
        r = collide_batch(o1, tf1, o2, tf2, sub, contact_capacity=min(budget, 256), grow_on_overflow=True)
    for c in r.contacts_of(0):
        if request.enable_contact:
            result.addContact(Contact(o1, o2, int(c["b1"]), int(c["b2"]), c["pos"].copy(), c["normal"].copy(),
                                      float(c["penetration_depth"])))
        else:
            result.addContact(Contact(o1, o2, int(c["b1"]), int(c["b2"])))
    return result.numContacts()


def distance(o1, tf1, o2=None, tf2=None, request=None, result=None):
    """fcl::distance(o1, tf1, o2, tf2, request, result) (distance-inl.h:92-246), or the CollisionObject
    overload distance(obj1, obj2, request, result) (distance-inl.h:196-218); returns min_distance."""
    if isinstance(o1, CollisionObject):
        obj1, obj2, request, result = o1, tf1, o2, tf2
        return distance(obj1.collisionGeometry(), obj1.getTransform(), obj2.collisionGeometry(), obj2.getTransform(),
                        request, result)
    if isinstance(o1, Sphere) and isinstance(o2, BVHModel):
        # (OT_GEOM, OT_BVH): the reference calls the [BVH][GEOM] cell with the arguments swapped and does not flip
        # the result (distance-inl.h:117-127), so o1 / nearest_points[0] of the result belong to the mesh
        o1, tf1, o2, tf2 = o2, tf2, o1, tf1
    mesh_sphere = isinstance(o1, BVHModel) and isinstance(o2, Sphere)
    if not (isinstance(o1, BVHModel) and isinstance(o2, BVHModel)) and not mesh_sphere:
        sys.stderr.write("Warning: distance function between these node types is not supported\n")
        return DBL_MAX
    if request.isSatisfied(result):  # orientedMeshDistance / orientedBVHShapeDistance, distance_func_matrix-inl.h:395,268
        return result.min_distance
    if mesh_sphere:
        ident = np.array([[1.0, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0]])
        # the mesh-shape leaf always hands its nearest points to DistanceResult::update
        # (mesh_shape_distance_traversal_node-inl.h:186-199), whatever the request says
        r = distance_mesh_sphere_batch(o1, ident if tf1 is None else tf1, o2, ident if tf2 is None else tf2,
                                       DistanceRequest(True))
    else:
        r = distance_batch(o1, tf1, o2, tf2, request)
    if request.enable_nearest_points or mesh_sphere:
        result.update(float(r.min_distance[0]), o1, o2, int(r.b1[0]), int(r.b2[0]), r.nearest_p1[0], r.nearest_p2[0])
    else:
        result.update(float(r.min_distance[0]), o1, o2, int(r.b1[0]), int(r.b2[0]))
    return result.min_distance


# ---------------------------------------------------------------------------------------------------------------------
# Continuous collision: narrowphase/continuous_collision.h, continuous_collision_request.h:48-80,
# continuous_collision_result.h.  Built: CCDM_TRANS motions with the conservative-advancement solver on two
# BVHModel<OBBRSS> (csrc/continuous.cuh); everything else answers like the reference's "not supported" branch.
# ---------------------------------------------------------------------------------------------------------------------
CCDM_TRANS, CCDM_LINEAR, CCDM_SCREW, CCDM_SPLINE = 0, 1, 2, 3
CCDC_NAIVE, CCDC_CONSERVATIVE_ADVANCEMENT, CCDC_RAY_SHOOTING, CCDC_POLYNOMIAL_SOLVER = 0, 1, 2, 3


class ContinuousCollisionRequest:
    def __init__(self, num_max_iterations=10, toc_err=0.0001, ccd_motion_type=CCDM_TRANS, gjk_solver_type="GST_LIBCCD",
                 ccd_solver_type=CCDC_NAIVE):
        self.num_max_iterations = num_max_iterations
        self.toc_err = toc_err
        self.ccd_motion_type = ccd_motion_type
        self.gjk_solver_type = gjk_solver_type
        self.ccd_solver_type = ccd_solver_type

    def _c(self):
        return _capi.ContinuousRequestC(int(self.num_max_iterations), float(self.toc_err), int(self.ccd_motion_type), 0,
                                        int(self.ccd_solver_type))


class ContinuousCollisionResult:
    def __init__(self):
        self.is_collide = False
        self.time_of_contact = 1.0
        self.contact_tf1 = Transform3()
        self.contact_tf2 = Transform3()


class BatchContinuousResult:
    def __init__(self, is_collide, time_of_contact, contact_tf1, contact_tf2, iterations):
        self.is_collide = is_collide
        self.time_of_contact = time_of_contact
        self.contact_tf1 = contact_tf1
        self.contact_tf2 = contact_tf2
        self.iterations = iterations


def continuous_collide_batch(o1, tf1_beg, tf1_end, o2, tf2_beg, tf2_end, request, device=None):
    """n independent fcl::continuousCollide(o1, tf1_beg[i], tf1_end[i], o2, tf2_beg[i], tf2_end[i], request, result_i)
    calls (host arrays in and out).  Any pose argument may be None (identity)."""
    arrs, ns = zip(*[_poses(t) for t in (tf1_beg, tf1_end, tf2_beg, tf2_end)])
    ns = [k for k in ns if k is not None]
    if not ns:
        raise ValueError("at least one pose array must be given")
    if len(set(ns)) != 1:
        raise ValueError("the pose arrays must have the same length")
    n = ns[0]
    m1, m2 = o1.device_model(device), o2.device_model(device)
    req = request._c()
    hit, toc, it = np.zeros(n, np.int32), np.zeros(n), np.zeros(n, np.int32)
    c1, c2 = np.zeros((n, 12)), np.zeros((n, 12))
    check(_capi.lib().fclgpu_continuous_collide_batch_host(m1, m2, n, addr(arrs[0]), addr(arrs[1]), addr(arrs[2]), addr(arrs[3]),
                                                           C.byref(req), addr(hit), addr(toc), addr(c1), addr(c2), addr(it)))
    return BatchContinuousResult(hit.astype(bool), toc, c1, c2, it)


def continuousCollide(o1, tf1_beg, tf1_end, o2=None, tf2_beg=None, tf2_end=None, request=None, result=None):
    """fcl::continuousCollide(o1, tf1_beg, tf1_end, o2, tf2_beg, tf2_end, request, result)
    (continuous_collision-inl.h:441-452), or the CollisionObject overload continuousCollide(obj1, tf1_end, obj2, tf2_end,
    request, result) (:455-470); returns the time of contact (-1 for unsupported settings, like the reference)."""
    if isinstance(o1, CollisionObject):
        obj1, end1, obj2, end2, request, result = o1, tf1_beg, tf1_end, o2, tf2_beg, tf2_end
        return continuousCollide(obj1.collisionGeometry(), obj1.getTransform(), end1, obj2.collisionGeometry(),
                                 obj2.getTransform(), end2, request, result)
    if request.ccd_solver_type != CCDC_CONSERVATIVE_ADVANCEMENT or request.ccd_motion_type != CCDM_TRANS or not (
            isinstance(o1, BVHModel) and isinstance(o2, BVHModel)):
        sys.stderr.write("Warning! Invalid continuous collision setting\n")
        return -1.0
    r = continuous_collide_batch(o1, tf1_beg, tf1_end, o2, tf2_beg, tf2_end, request)
    result.is_collide = bool(r.is_collide[0])
    result.time_of_contact = float(r.time_of_contact[0])
    if result.is_collide:
        result.contact_tf1 = Transform3.from_pose12(r.contact_tf1[0])
        result.contact_tf2 = Transform3.from_pose12(r.contact_tf2[0])
    return result.time_of_contact
