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


def _run_device_logic_test(name):
    """The register scanner (k1_scan.cuh) and the chunked record decoder (k2_decode.cuh) are __host__ __device__:
    the same source the kernels use is compiled with g++ and checked against nucleotide-by-nucleotide restatements."""
    src = os.path.join(ROOT, "tests", "cpp", name + ".cpp")
    exe = os.path.join(ROOT, "tests", "cpp", name)
    inc = os.path.join(ROOT, "gatb_core_b200", "csrc")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wno-unknown-pragmas", "-I" + inc, src, "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr


def test_k1_scanner_cpu():
    _run_device_logic_test("test_k1_scan")


def test_k2_decoder_cpu():
    _run_device_logic_test("test_k2_decode")


def test_coarse_layout_cpu():
    _run_device_logic_test("test_layout")
