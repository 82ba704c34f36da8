"""include/fclgpu/fcl_shim.hpp (the C++ drop-in for fcl::collide / fcl::distance on BVHModel<OBBRSS<double>>)
is dormant without FCL + Eigen.  It is compile-checked here against a mock of the FCL API surface it touches
(tests/shim/mock) and, on a GPU, run end to end against direct C-ABI calls (tests/shim/shim_check.cpp)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "fcl_b200", "lib")


def _build(tmp_path):
    exe = str(tmp_path / "shim_check")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "tests", "shim", "mock"),
                           "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "shim", "shim_check.cpp"),
                           "-L", LIBDIR, "-lfclgpu", "-Wl,-rpath," + LIBDIR, "-o", exe])
    return exe


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_shim_compiles_against_the_fcl_api_surface(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe, "compile-only"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr


@pytest.mark.gpu
def test_shim_results_equal_direct_c_abi_calls(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "shim OK" in out.stdout
