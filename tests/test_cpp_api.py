"""Builds tests/cpp/test_polynomial.cpp (the obake-shaped C++ host API over the C ABI) and runs it: the
host-only subset on CPU, everything (operator*, truncated_mul, known answers) on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "_build", "test_polynomial")
SRC = os.path.join(ROOT, "tests", "cpp", "test_polynomial.cpp")


def build_exe():
    from obake_b200 import build

    build.build_native()
    inc = os.path.join(ROOT, "obake_b200", "include")
    newest = max(os.path.getmtime(os.path.join(dp, f)) for dp, _, fs in os.walk(inc) for f in fs)
    newest = max(newest, os.path.getmtime(SRC), os.path.getmtime(os.path.join(ROOT, "include", "obk_b200.h")))
    if os.path.exists(EXE) and os.path.getmtime(EXE) >= newest:
        return EXE
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    cmd = ["g++", "-std=c++20", "-O2", "-Wall", "-I", inc, "-o", EXE, SRC, "-L", os.path.join(ROOT, "obake_b200", "lib"),
           "-lobk_b200", "-pthread", "-Wl,-rpath,$ORIGIN/../../obake_b200/lib"]
    subprocess.check_call(cmd)
    return EXE


def test_cpp_host_only_subset():
    exe = build_exe()
    r = subprocess.run([exe, "--cpu"], capture_output=True, text=True, timeout=300)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0 and "OK:" in r.stdout


@pytest.mark.gpu
def test_cpp_full_suite_on_gpu():
    exe = build_exe()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=580)
    print(r.stdout[-4000:], r.stderr[-2000:])
    assert r.returncode == 0 and "OK:" in r.stdout


def test_tile_geometry_helpers_on_host():
    """tests/cpp/test_tiles_host.cu: the tile path's geometry helpers (tile enumeration and its inverse, tile groups,
    ring-stage layout) against brute force, compiled with nvcc and run on the CPU."""
    src = os.path.join(ROOT, "tests", "cpp", "test_tiles_host.cu")
    exe = os.path.join(ROOT, "tests", "_build", "test_tiles_host")
    dep = os.path.join(ROOT, "obake_b200", "csrc")
    newest = max([os.path.getmtime(src)] + [os.path.getmtime(os.path.join(dep, f)) for f in os.listdir(dep)])
    if not (os.path.exists(exe) and os.path.getmtime(exe) >= newest):
        os.makedirs(os.path.dirname(exe), exist_ok=True)
        nvcc = "/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else "nvcc"
        subprocess.check_call([nvcc, "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe, src])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0 and "OK:" in r.stdout


def test_container_rebuild_benchmark_builds_and_runs():
    """tools/bench_adopt.cpp (the host container rebuild timed alone, no GPU): compiles against the mirror and adopts a
    small synthetic grouped result, once with the block cache and once without."""
    import json

    src = os.path.join(ROOT, "tools", "bench_adopt.cpp")
    exe = os.path.join(ROOT, "tests", "_build", "bench_adopt")
    inc = os.path.join(ROOT, "obake_b200", "include")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++20", "-O2", "-Wall", "-I", inc, "-I", os.path.join(ROOT, "include"), "-o", exe, src,
                           "-pthread"])
    for env in (dict(os.environ), dict(os.environ, OBAKE_B200_SLAB_CACHE_MB="0")):
        r = subprocess.run([exe, "200000", "6", "2", "2"], capture_output=True, text=True, timeout=120, env=env)
        assert r.returncode == 0, r.stderr[-1000:]
        d = json.loads(r.stdout.strip().splitlines()[-1])
        assert d["terms"] == 200000 and d["slot_bytes"] == 40 and d["adopt_ms_mean"] > 0
