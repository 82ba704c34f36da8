import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # the product library normally travels with the tree; on a bare checkout compile it (nvcc cross-compiles
    # sm_100a without a GPU) -- the package itself never falls back to anything else
    if not os.path.exists(os.path.join(ROOT, "fcl_b200", "lib", "libfclgpu.so")):
        import subprocess

        subprocess.check_call(["make", "-C", os.path.join(ROOT, "fcl_b200", "csrc"), "-s"])


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle

    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def env_rob_npz():
    import numpy as np

    e = np.load(os.path.join(GOLDEN, "env.npz"))
    r = np.load(os.path.join(GOLDEN, "rob.npz"))
    return (e["verts"], e["tris"]), (r["verts"], r["tris"])


@pytest.fixture(scope="session")
def oracle_env_rob(oracle, env_rob_npz):
    (ev, et), (rv, rt) = env_rob_npz
    return oracle.Model(ev, et), oracle.Model(rv, rt)
