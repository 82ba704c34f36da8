"""The boundary is a C ABI: a plain-C translation unit must compile against include/fclgpu.h,
link against libfclgpu.so and run (CPU part only when no GPU is present)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "fcl_b200", "lib")


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_plain_c_client_builds_and_runs(tmp_path):
    exe = str(tmp_path / "abi_check")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_abi", "abi_check.c"), "-o", exe, "-L", LIBDIR, "-lfclgpu",
                           "-Wl,-rpath," + LIBDIR])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "abi_check ok" in out.stdout


def test_header_is_valid_cxx_too(tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text('#include "fclgpu.h"\nint main() { return fclgpu_abi_version() == FCLGPU_ABI_VERSION ? 0 : 1; }\n')
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)])
