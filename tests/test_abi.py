"""The C-ABI library loads and exports every symbol include/eventcalib_b200.h declares (no compute calls here)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import eventcalib_b200 as ecb
    hdr = open(os.path.join(ROOT, "include", "eventcalib_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(ecb_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 30
    lib = ecb.load_library()
    for name in declared:
        assert hasattr(lib, name), "libecb.so does not export %s" % name
    assert sorted(set(ecb.SYMBOLS)) == declared, "python SYMBOLS list and header disagree"
    assert lib.ecb_version().decode().startswith("eventcalib_b200")


def test_struct_layouts_match_header():
    import ctypes as C
    import eventcalib_b200 as ecb
    assert C.sizeof(ecb.FrontendParams) == 48
    assert C.sizeof(ecb.WindowSummary) == 64
    assert C.sizeof(ecb.LmOptions) == 16 + 9 * 8
    assert C.sizeof(ecb.LmSummary) == 16 + 4 * 8
    o = ecb.lm_options()
    assert o.max_iterations == 50 and o.function_tolerance == 1e-10 and o.initial_radius == 1e4


def test_no_gpu_fails_loudly():
    import pytest
    import torch
    import eventcalib_b200 as ecb
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(ecb.EcbError):
        ecb.Context(0)


def test_product_does_not_import_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "eventcalib_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src, f
