"""include/obk_b200.h is a C header (the drop-in boundary is a C ABI): it must compile as C99 and as C++17 on its
own, and the ctypes mirrors in obake_b200/_capi.py must have the sizes the C compiler gives the structs."""
import ctypes as C
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "obk_b200.h")

PROBE = r"""
#include <stdio.h>
#include "obk_b200.h"
int main(void)
{
    printf("obk_poly_view %zu\n", sizeof(obk_poly_view));
    printf("obk_trunc %zu\n", sizeof(obk_trunc));
    printf("obk_poly_result %zu\n", sizeof(obk_poly_result));
    printf("obk_stats %zu\n", sizeof(obk_stats));
    printf("obk_peer_windows %zu\n", sizeof(obk_peer_windows));
    return 0;
}
"""


def _sizes(compiler, std, suffix):
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "probe" + suffix)
        exe = os.path.join(d, "probe")
        with open(src, "w") as f:
            f.write(PROBE)
        subprocess.check_call([compiler, std, "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.dirname(HDR),
                               "-o", exe, src])
        out = subprocess.check_output([exe], text=True)
    return {ln.split()[0]: int(ln.split()[1]) for ln in out.strip().splitlines()}


def test_header_compiles_as_c_and_cpp_and_matches_ctypes():
    from obake_b200 import _capi

    c = _sizes("gcc", "-std=c99", ".c")
    cpp = _sizes("g++", "-std=c++17", ".cpp")
    assert c == cpp
    assert c["obk_poly_view"] == C.sizeof(_capi.PolyView)
    assert c["obk_trunc"] == C.sizeof(_capi.Trunc)
    assert c["obk_poly_result"] == C.sizeof(_capi.PolyResult)
    assert c["obk_stats"] == C.sizeof(_capi.Stats)
    assert c["obk_peer_windows"] == C.sizeof(_capi.PeerWindows)


def test_every_declared_function_is_bound():
    """Every obk_* function the header declares is listed in _capi.SYMBOLS (and therefore checked for export)."""
    import re

    from obake_b200 import _capi

    text = open(HDR).read()
    declared = set(re.findall(r"\b(obk_[a-z_0-9]+)\s*\(", text))
    declared -= {"obk_engine", "obk_status"}
    assert declared == set(_capi.SYMBOLS), (declared ^ set(_capi.SYMBOLS))
