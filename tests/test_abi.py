"""CPU-side checks of the boundary: the shared library loads and exports every symbol include/obk_b200.h
declares; without a GPU the engine refuses to start instead of falling back."""
import ctypes as C
import os
import re

import pytest

from obake_b200 import _capi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "obk_b200.h")).read()
    return sorted(set(re.findall(r"\b(obk_[a-z_0-9]+)\s*\(", hdr)))


def test_header_symbols_all_exported():
    build.build_native()
    lib = C.CDLL(_capi.LIB_PATH)
    declared = _declared_symbols()
    assert declared, "no symbols parsed from the header"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in obk_b200.h but not exported"
    assert sorted(_capi.SYMBOLS) == declared


def test_struct_sizes_match_header_layout():
    # fixed-width fields only: the ctypes mirrors must have the C sizes
    assert C.sizeof(_capi.PolyView) == 8 + 8 * 4 + 16
    assert C.sizeof(_capi.Trunc) == 24
    assert C.sizeof(_capi.PolyResult) == 8 + 8 * 4 + 4 * 8
    assert C.sizeof(_capi.Stats) == 5 * 8 + 10 * 4 + 7 * 4 + 4  # padded to 8


def test_version_and_status_strings():
    lib = _capi.lib()
    assert b"obk_b200" in lib.obk_version()
    assert lib.obk_status_string(2) == b"monomial exponent overflow"


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from obake_b200 import engine
    with pytest.raises(engine.ObkCudaError):
        engine.Engine(0)


def test_product_path_never_touches_the_oracle():
    # the oracle is test infrastructure: nothing under obake_b200/ may reference it
    for dirpath, _, files in os.walk(os.path.join(ROOT, "obake_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "obk_oracle" not in txt and "pyoracle" not in txt and "import orc" not in txt, f
