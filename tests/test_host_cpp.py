"""Builds and runs the C++ tests of the host-side GATB API mirror (tests/cpp/test_host_api.cpp)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "test_host_api")


def build():
    src = os.path.join(ROOT, "tests", "cpp", "test_host_api.cpp")
    lib = os.path.join(ROOT, "gatb_core_b200")
    if not os.path.exists(BIN) or os.path.getmtime(BIN) < max(os.path.getmtime(src), os.path.getmtime(os.path.join(lib, "host", "gatb", "gatb_core_b200.hpp"))):
        subprocess.run(["g++", "-std=c++11", "-O1", "-Wall", "-I" + os.path.join(lib, "host"), src, "-o", BIN,
                        "-L" + lib, "-lgatb_b200", "-Wl,-rpath," + lib, "-Wl,-rpath,/usr/local/cuda/lib64"], check=True)


def test_host_models_cpu():
    build()
    out = subprocess.run([BIN, "cpu"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr


@pytest.mark.gpu
def test_host_api_gpu():
    build()
    out = subprocess.run([BIN, "gpu"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
