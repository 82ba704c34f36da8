"""ctypes binding of the C ABI in include/fclgpu.h (fcl_b200/lib/libfclgpu.so).

There is no CPU fallback: if the shared library is missing this module raises at import
time with the build command, and every compute call fails with FCLGPU_ERR_NO_DEVICE when no
CUDA device is present.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FCLGPU_LIB_PATH") or os.path.join(_HERE, "lib", "libfclgpu.so")  # override: A/B builds

OK = 0
ERR_BUILD_OUT_OF_SEQUENCE = -2
ERR_BUILD_EMPTY_MODEL = -3
ERR_UNSUPPORTED_FUNCTION = -5
ERR_INCORRECT_DATA = -7
ERR_INVALID_ARGUMENT = -20
ERR_NO_DEVICE = -21
ERR_CUDA = -22
ERR_CONTACT_OVERFLOW = -23
ERR_STACK_OVERFLOW = -24
ERR_INPUT_STALLED = -25
ERR_COMM = -26
SHAPE_PLANE, SHAPE_HALFSPACE = 16, 17  # fcl::GEOM_PLANE / GEOM_HALFSPACE

CONTACT_DTYPE = np.dtype(
    [("b1", "<i4"), ("b2", "<i4"), ("normal", "<f8", (3,)), ("pos", "<f8", (3,)), ("penetration_depth", "<f8")]
)
assert CONTACT_DTYPE.itemsize == 64
# compact records (include/fclgpu.h: fclgpu_contact_ids, fclgpu_contact_f32)
CONTACT_IDS_DTYPE = np.dtype([("b1", np.int32), ("b2", np.int32)])
CONTACT_F32_DTYPE = np.dtype([("b1", np.int32), ("b2", np.int32), ("normal", np.float32, 3), ("pos", np.float32, 3),
                              ("penetration_depth", np.float32), ("reserved", np.float32)])
assert CONTACT_IDS_DTYPE.itemsize == 8 and CONTACT_F32_DTYPE.itemsize == 40


class ContinuousRequestC(C.Structure):  # fclgpu_continuous_request
    _fields_ = [("num_max_iterations", C.c_int64), ("toc_err", C.c_double), ("ccd_motion_type", C.c_int32),
                ("gjk_solver_type", C.c_int32), ("ccd_solver_type", C.c_int32)]


class CollisionRequestC(C.Structure):
    _fields_ = [("num_max_contacts", C.c_int64), ("enable_contact", C.c_int32), ("enable_cost", C.c_int32),
                ("stage_capacity", C.c_int64), ("contact_format", C.c_int32), ("reserved", C.c_int32)]


class DistanceRequestC(C.Structure):
    _fields_ = [("enable_nearest_points", C.c_int32), ("enable_signed_distance", C.c_int32),
                ("rel_err", C.c_double), ("abs_err", C.c_double)]


class FclGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"fclgpu error {code}: {msg}")
        self.code = code


# every symbol include/fclgpu.h declares (tests check the library exports all of them)
SYMBOLS = [
    "fclgpu_bvh_build_obbrss", "fclgpu_bvh_destroy", "fclgpu_bvh_num_nodes", "fclgpu_bvh_num_tris", "fclgpu_bvh_get",
    "fclgpu_bvh_refit_topdown", "fclgpu_bvh_num_vertices", "fclgpu_bvh_get_partition", "fclgpu_model_set_partition",
    "fclgpu_continuous_collide_batch", "fclgpu_continuous_collide_batch_host",
    "fclgpu_bvh_refit_bottomup", "fclgpu_bvh_get_rss_axis", "fclgpu_model_refit_bottomup", "fclgpu_model_download_rss_axis",
    "fclgpu_model_refit_topdown", "fclgpu_model_download", "fclgpu_model_build_obbrss", "fclgpu_model_get_topology",
    "fclgpu_collide_mesh_sphere_batch", "fclgpu_collide_mesh_sphere_batch_host",
    "fclgpu_distance_mesh_sphere_batch", "fclgpu_distance_mesh_sphere_batch_host",
    "fclgpu_distance_cutoff_batch", "fclgpu_distance_cutoff_batch_host",
    "fclgpu_within_tolerance_batch", "fclgpu_within_tolerance_batch_host",
    "fclgpu_load_obj", "fclgpu_save_obj", "fclgpu_free", "fclgpu_model_create_obbrss2", "fclgpu_model_local_aabb", "fclgpu_broadphase_collide_host",
    "fclgpu_collide_mesh_plane_batch", "fclgpu_collide_mesh_plane_batch_host",
    "fclgpu_shard_range", "fclgpu_comm_unique_id", "fclgpu_comm_init", "fclgpu_comm_rank", "fclgpu_comm_world",
    "fclgpu_comm_allgather", "fclgpu_comm_allgather_ragged", "fclgpu_comm_destroy", "fclgpu_comm_last_error",
    "fclgpu_model_create_obbrss", "fclgpu_model_from_bvh", "fclgpu_model_destroy", "fclgpu_model_num_nodes",
    "fclgpu_model_num_tris", "fclgpu_model_device", "fclgpu_collide_batch", "fclgpu_collide_batch_host",
    "fclgpu_distance_batch", "fclgpu_distance_batch_host", "fclgpu_abi_version", "fclgpu_device_count",
    "fclgpu_last_error", "fclgpu_pose_from_colmajor4x4", "fclgpu_sync_status", "fclgpu_set_option",
    "fclgpu_get_option", "fclgpu_launch_count", "fclgpu_debug_counters", "fclgpu_microbench", "fclgpu_device_trim",
]

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA library first (python -c 'import __graft_entry__ as g; g.build()' "
            "or make -C fcl_b200/csrc). fcl_b200 has no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    vp, dp, ip, up = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p  # raw addresses (host or device)
    L.fclgpu_bvh_build_obbrss.argtypes = [vp, C.c_int32, vp, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
    L.fclgpu_bvh_destroy.argtypes = [vp]
    L.fclgpu_bvh_destroy.restype = None
    L.fclgpu_bvh_num_nodes.argtypes = [vp]
    L.fclgpu_bvh_num_tris.argtypes = [vp]
    L.fclgpu_bvh_get.argtypes = [vp] * 9
    L.fclgpu_bvh_refit_topdown.argtypes = [vp, vp, C.c_int32]
    L.fclgpu_bvh_num_vertices.argtypes = [vp]
    L.fclgpu_bvh_get_partition.argtypes = [vp] * 5
    L.fclgpu_model_set_partition.argtypes = [vp, C.c_int32, vp, vp, vp, vp]
    L.fclgpu_model_refit_topdown.argtypes = [vp, vp, C.c_int32, C.c_int32, vp]
    L.fclgpu_model_download.argtypes = [vp] * 8
    L.fclgpu_bvh_refit_bottomup.argtypes = [vp, vp, C.c_int32]
    L.fclgpu_continuous_collide_batch.argtypes = [vp, vp, C.c_int64, vp, vp, vp, vp, C.POINTER(ContinuousRequestC), vp, vp, vp, vp, vp,
                                                  vp, vp, vp]
    L.fclgpu_continuous_collide_batch_host.argtypes = [vp, vp, C.c_int64, vp, vp, vp, vp, C.POINTER(ContinuousRequestC), vp, vp, vp,
                                                       vp, vp]
    L.fclgpu_bvh_get_rss_axis.argtypes = [vp, vp]
    L.fclgpu_model_refit_bottomup.argtypes = [vp, vp, C.c_int32, C.c_int32, vp]
    L.fclgpu_model_download_rss_axis.argtypes = [vp, vp]
    L.fclgpu_model_build_obbrss.argtypes = [C.c_int, vp, C.c_int32, vp, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
    L.fclgpu_model_get_topology.argtypes = [vp] * 5
    L.fclgpu_model_create_obbrss.argtypes = [C.c_int, C.c_int32, vp, vp, vp, vp, vp, vp, vp, C.c_int32, vp,
                                             C.POINTER(C.c_void_p)]
    L.fclgpu_model_from_bvh.argtypes = [C.c_int, vp, C.POINTER(C.c_void_p)]
    L.fclgpu_model_destroy.argtypes = [vp]
    L.fclgpu_model_num_nodes.argtypes = [vp]
    L.fclgpu_model_num_tris.argtypes = [vp]
    L.fclgpu_model_device.argtypes = [vp]
    L.fclgpu_collide_batch.argtypes = [vp, vp, C.c_int64, dp, dp, C.POINTER(CollisionRequestC), ip, vp, C.c_int64,
                                       vp, up, up, vp]
    L.fclgpu_collide_batch_host.argtypes = [vp, vp, C.c_int64, dp, dp, C.POINTER(CollisionRequestC), ip, vp,
                                            C.c_int64, vp, up, up]
    L.fclgpu_collide_mesh_sphere_batch.argtypes = [vp, C.c_double, C.c_int64, dp, dp, C.POINTER(CollisionRequestC), ip, vp,
                                                   C.c_int64, vp, up, up, vp]
    L.fclgpu_collide_mesh_sphere_batch_host.argtypes = [vp, C.c_double, C.c_int64, dp, dp, C.POINTER(CollisionRequestC), ip,
                                                        vp, C.c_int64, vp, up, up]
    L.fclgpu_distance_batch.argtypes = [vp, vp, C.c_int64, dp, dp, C.POINTER(DistanceRequestC), dp, dp, dp, ip, ip,
                                        up, up, vp]
    L.fclgpu_distance_batch_host.argtypes = [vp, vp, C.c_int64, dp, dp, C.POINTER(DistanceRequestC), dp, dp, dp, ip,
                                             ip, up, up]
    L.fclgpu_distance_mesh_sphere_batch.argtypes = [vp, C.c_double, C.c_int64, dp, dp, C.POINTER(DistanceRequestC), dp, dp,
                                                    dp, ip, ip, up, up, vp]
    L.fclgpu_distance_mesh_sphere_batch_host.argtypes = [vp, C.c_double, C.c_int64, dp, dp, C.POINTER(DistanceRequestC), dp,
                                                         dp, dp, ip, ip, up, up]
    L.fclgpu_distance_cutoff_batch.argtypes = [vp, vp, C.c_int64, dp, dp, C.POINTER(DistanceRequestC), C.c_double, dp, dp, dp,
                                               ip, ip, up, up, vp]
    L.fclgpu_distance_cutoff_batch_host.argtypes = [vp, vp, C.c_int64, dp, dp, C.POINTER(DistanceRequestC), C.c_double, dp, dp,
                                                    dp, ip, ip, up, up]
    L.fclgpu_within_tolerance_batch.argtypes = [vp, vp, C.c_int64, dp, dp, C.c_double, vp, dp, up, up, vp]
    L.fclgpu_within_tolerance_batch_host.argtypes = [vp, vp, C.c_int64, dp, dp, C.c_double, vp, dp, up, up]
    L.fclgpu_collide_mesh_plane_batch.argtypes = [vp, C.c_int32, dp, C.c_double, C.c_int64, dp, dp, C.POINTER(CollisionRequestC), ip,
                                                  vp, C.c_int64, vp, up, up, vp]
    L.fclgpu_collide_mesh_plane_batch_host.argtypes = [vp, C.c_int32, dp, C.c_double, C.c_int64, dp, dp, C.POINTER(CollisionRequestC),
                                                       ip, vp, C.c_int64, vp, up, up]
    L.fclgpu_model_create_obbrss2.argtypes = [C.c_int, C.c_int32, ip, dp, dp, dp, dp, dp, dp, dp, C.c_int32, dp, C.POINTER(vp)]
    L.fclgpu_model_local_aabb.argtypes = [vp, dp, dp, dp, dp]
    L.fclgpu_broadphase_collide_host.argtypes = [C.c_int32, C.POINTER(vp), C.c_int64, ip, dp, C.c_int64, ip, dp,
                                                 C.POINTER(CollisionRequestC), C.c_int64, ip, C.POINTER(C.c_int64), ip, dp, dp]
    L.fclgpu_load_obj.argtypes = [C.c_char_p, C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.c_int32),
                                  C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.c_int32)]
    L.fclgpu_save_obj.argtypes = [C.c_char_p, dp, C.c_int32, ip, C.c_int32]
    L.fclgpu_free.argtypes = [vp]
    L.fclgpu_free.restype = None
    L.fclgpu_shard_range.argtypes = [C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.fclgpu_shard_range.restype = None
    L.fclgpu_comm_unique_id.argtypes = [C.c_char_p]
    L.fclgpu_comm_init.argtypes = [C.c_int, C.c_int, C.c_int, C.c_char_p, C.POINTER(vp)]
    L.fclgpu_comm_rank.argtypes = [vp]
    L.fclgpu_comm_world.argtypes = [vp]
    L.fclgpu_comm_allgather.argtypes = [vp, vp, vp, C.c_size_t, vp]
    L.fclgpu_comm_allgather_ragged.argtypes = [vp, vp, vp, C.POINTER(C.c_int64), vp]
    L.fclgpu_comm_destroy.argtypes = [vp]
    L.fclgpu_comm_last_error.restype = C.c_char_p
    L.fclgpu_last_error.restype = C.c_char_p
    L.fclgpu_pose_from_colmajor4x4.argtypes = [vp, vp]
    L.fclgpu_pose_from_colmajor4x4.restype = None
    L.fclgpu_sync_status.argtypes = [C.c_int, vp]
    L.fclgpu_device_trim.argtypes = [C.c_int, C.POINTER(C.c_int64)]
    L.fclgpu_set_option.argtypes = [C.c_char_p, C.c_int64]
    L.fclgpu_get_option.argtypes = [C.c_char_p]
    L.fclgpu_get_option.restype = C.c_int64
    L.fclgpu_launch_count.restype = C.c_int64
    L.fclgpu_debug_counters.argtypes = [C.c_int, C.POINTER(C.c_uint64), C.c_int]
    L.fclgpu_microbench.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double)]
    _lib = L
    return L


def check(code):
    if code != OK:
        raise FclGpuError(code, lib().fclgpu_last_error().decode("utf-8", "replace"))


def addr(a):
    """Address of a numpy array (host) or torch tensor (host or device); None -> NULL."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()


def set_option(name, value):
    check(lib().fclgpu_set_option(name.encode(), int(value)))


def get_option(name):
    return int(lib().fclgpu_get_option(name.encode()))


def launch_count():
    return int(lib().fclgpu_launch_count())


def device_count():
    return int(lib().fclgpu_device_count())


def microbench(kind, device=0):
    out = C.c_double(0)
    check(lib().fclgpu_microbench(device, kind, C.byref(out)))
    return out.value
