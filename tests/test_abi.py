"""CPU-only checks of the drop-in boundary: the library loads and exports every symbol include/gatb_gpu.h declares;
without a device the compute path fails loudly (no CPU fallback)."""
import os
import re

import pytest

import gatb_core_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "gatb_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gatb_gpu_[a-z0-9_]+)\s*\(", text)))


def test_header_and_exports_agree():
    lib = gatb_core_b200.load_library()
    decl = declared_symbols()
    assert decl == sorted(gatb_core_b200.EXPORTS)
    for s in decl:
        assert hasattr(lib, s), s


def test_struct_layout_matches_header():
    # gatb_gpu_params: 11 named int32 + 5 reserved; gatb_gpu_result: see header
    assert gatb_core_b200.C.sizeof(gatb_core_b200.Params) == 16 * 4
    assert gatb_core_b200.C.sizeof(gatb_core_b200.Result) == 8 * 7 + 16 * 8 + 8 * 8 + 8 * 8 + 4 + 4 + 8


def test_host_only_entry_points():
    lib = gatb_core_b200.load_library()
    import ctypes as C
    s, h = C.c_uint64(), C.c_int32()
    assert lib.gatb_gpu_bloom_params(21, 49972, C.byref(s), C.byref(h)) == 0
    assert (s.value, h.value) == (293528, 4)           # reference's Bloom at config 1
    nb, bits = C.c_uint64(), C.c_uint64()
    assert lib.gatb_gpu_bloom_layout(2, 293528, C.byref(nb), C.byref(bits)) == 0
    assert (nb.value, bits.value) == (37716, 293528)   # SURVEY.md 8(a) row G [probe]: 37 716 bytes


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(gatb_core_b200.GatbGpuError):
        gatb_core_b200.GatbGpu(0)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "gatb_core_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                # comments may cite the oracle's mirrored synthetic generator; code must not include/link/load it
                assert not re.search(r'#include\s*[<"][^>"]*oracle', text), f
                assert "liboracle" not in text and "libgatbref" not in text and "oracle_lib" not in text, f
