"""C++ host mirror of src/api.rs (vector-indexer_b200/csrc/api.hpp): compiles against the C ABI;
error behaviour checked without a GPU, the full tests/api_tests.rs flow on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "vector-indexer_b200", "lib")


def build(tmp_path):
    exe = str(tmp_path / "api_test")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([gxx, "-std=c++17", "-O1", os.path.join(ROOT, "tests", "cpp", "api_test.cpp"), "-o", exe,
                           "-L" + LIBDIR, "-lvidx_b200", "-Wl,-rpath," + LIBDIR])
    return exe


def test_cpp_api_errors_without_device(tmp_path):
    out = subprocess.run([build(tmp_path), "errors"], capture_output=True, text=True)
    assert out.returncode == 0 and "OK errors" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_cpp_api_full(tmp_path):
    out = subprocess.run([build(tmp_path), "full", str(tmp_path / "work")], capture_output=True, text=True)
    assert out.returncode == 0 and "OK full" in out.stdout, out.stdout + out.stderr
