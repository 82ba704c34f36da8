"""fcl_b200 — B200-native batched BVHModel<OBBRSS<double>> mesh-mesh collide()/distance().

Only the hot path lives here: csrc/ (CUDA kernels + the C ABI of include/fclgpu.h),
api.py (host-side mirror of the reference's interface for this path) and poses.py
(synthetic pose batches).  There is no CPU fallback.
"""
from .poses import ENV_EXTENTS, identity_poses, random_poses  # noqa: F401


def __getattr__(name):  # api needs the built shared library; import it lazily
    if name.startswith("__"):
        raise AttributeError(name)
    import importlib

    api = importlib.import_module(__name__ + ".api")
    if name == "api":
        return api
    if name == "_capi":
        return importlib.import_module(__name__ + "._capi")
    return getattr(api, name)
