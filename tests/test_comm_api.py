"""fclgpu_comm_* (C ABI multi-GPU wrapper, include/fclgpu.h): what can be checked without a GPU -- the pose partition
agrees with the Python one used by the gloo test, and the entry points fail with a status code (never crash) when there
is no device.  The collective itself runs under torchrun on the GPU box (bench.py --gpus N uses it for every gather)."""
import ctypes as C

import pytest

from fcl_b200 import _capi
from fcl_b200.sharding import shard_range


def test_shard_range_matches_python_partition():
    L = _capi.lib()
    for n in (0, 1, 7, 8, 1000, 1_000_003):
        for world in (1, 2, 3, 8):
            covered = 0
            for rank in range(world):
                s, c = C.c_int64(), C.c_int64()
                L.fclgpu_shard_range(n, rank, world, C.byref(s), C.byref(c))
                ps, pe = shard_range(n, rank, world)
                assert (s.value, s.value + c.value) == (ps, pe)
                assert s.value == covered
                covered += c.value
            assert covered == n


def test_comm_entry_points_fail_cleanly_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("needs a box without a GPU")
    L = _capi.lib()
    h = C.c_void_p()
    ident = C.create_string_buffer(128)
    rc = L.fclgpu_comm_init(0, 0, 1, ident, C.byref(h))
    assert rc in (_capi.ERR_NO_DEVICE, -5) and not h.value  # no device, or NCCL itself not installed
    assert L.fclgpu_comm_init(0, 3, 2, ident, C.byref(h)) == _capi.ERR_INVALID_ARGUMENT
    assert L.fclgpu_comm_allgather(None, None, None, 8, None) == _capi.ERR_INVALID_ARGUMENT
    assert L.fclgpu_comm_destroy(None) == 0
    assert L.fclgpu_comm_rank(None) == -1 and L.fclgpu_comm_world(None) == 0
