"""Test-only host build of the product's device math (fcl_b200/csrc/device_math.cuh)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "devmath_host.cu")
_CSRC = os.path.join(_HERE, "..", "..", "fcl_b200", "csrc")
_HDR = os.path.join(_CSRC, "device_math.cuh")
_HDRS = [os.path.join(_CSRC, f) for f in ("device_math.cuh", "sum_order.h", "mesh_sphere.cuh", "bounds_f32.cuh", "records.hpp")]
_SUM3 = os.environ.get("FCL_SUM3_ORDER", "0")  # see tests/test_sum_order_hook.py
_OUT = os.path.join(_HERE, "_build", "libdevmath_host.so" if _SUM3 == "0" else "libdevmath_host_sum3_%s.so" % _SUM3)


def build():
    stale = (not os.path.exists(_OUT)) or any(os.path.getmtime(s) > os.path.getmtime(_OUT) for s in [_SRC] + _HDRS)
    if stale:
        os.makedirs(os.path.dirname(_OUT), exist_ok=True)
        subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-Xcompiler",
                               "-fPIC,-ffp-contract=off", "-DFCL_SUM3_ORDER=" + _SUM3, "-shared", "-o", _OUT, _SRC])
    return _OUT


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        L.hm_obb_disjoint.restype = C.c_int
        L.hm_obb_disjoint.argtypes = [dp] * 4
        L.hm_rect_distance.restype = C.c_double
        L.hm_rect_distance.argtypes = [dp] * 4
        L.hm_obb_pairs.argtypes = [C.c_longlong, dp, ip, ip, dp, dp, dp, dp, dp, dp, ip]
        L.hm_rss_pairs.argtypes = [C.c_longlong, dp, ip, ip, dp, dp, dp, dp, dp, dp, dp, dp, dp]
        L.hm_tri_intersect.restype = C.c_int
        L.hm_tri_intersect.argtypes = [dp, dp, dp, C.c_int, C.POINTER(C.c_uint32), dp, dp, dp]
        L.hm_tri_distance.restype = C.c_double
        L.hm_tri_distance.argtypes = [dp, dp, dp, dp]
        L.hm_plane_tri_intersect.restype = C.c_int
        L.hm_plane_tri_intersect.argtypes = [C.c_int, dp, C.c_double, dp, dp]
        L.hm_sphere_tri_intersect.restype = C.c_int
        L.hm_sphere_tri_intersect.argtypes = [dp, C.c_double, dp, dp]
        L.hm_sphere_tri_distance.restype = C.c_int
        L.hm_sphere_tri_distance.argtypes = [dp, C.c_double, dp, dp]
        L.hm_mesh_sphere_distance.restype = C.c_int
        L.hm_mesh_sphere_distance.argtypes = [C.c_longlong, dp, dp, C.c_double, ip, dp, dp, dp, dp, dp, dp, dp, ip,
                                              C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_int]
        L.hm_obb_disjoint32_pairs.argtypes = [C.c_longlong, dp, ip, ip, dp, dp, dp, dp, dp, dp, ip]
        L.hm_tri_classify32.restype = C.c_int
        L.hm_tri_classify32.argtypes = [dp, dp, dp]
        L.hm_tri_lb32.restype = C.c_float
        L.hm_tri_lb32.argtypes = [dp, dp]
        L.hm_rss_lb32_pairs.argtypes = [C.c_longlong, dp, ip, ip, dp, dp, dp, dp, dp, dp, dp, dp, C.POINTER(C.c_float)]
        _lib = L
    return _lib


def dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def iptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))
