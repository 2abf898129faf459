"""Builds libobk_b200.so (the CUDA engine + C ABI) in-tree for sm_100a with nvcc."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "lib", "libobk_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "static", "-Xptxas", "-v",
]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def sources() -> list[str]:
    return [os.path.join(CSRC, "obk_engine.cu")]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(ROOT, "include", "obk_b200.h"), os.path.join(PKG, "include", "obake", "kpack.hpp")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_native(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = [_nvcc()] + NVCC_FLAGS + ["-o", LIB] + sources()
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(os.path.join(PKG, "lib", "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if r.returncode != 0:
        sys.stderr.write(log[-8000:])
        raise RuntimeError("nvcc failed")
    if verbose:
        print(log[-3000:])
    return LIB


BENCH_CPP = os.path.join(ROOT, "tools", "_build", "bench_cpp")


def build_bench_cpp(force: bool = False) -> str:
    """tools/bench_cpp.cpp: obake's benchmark drivers (benchmark/dense.hpp, sparse.hpp) against the C++ host mirror --
    the end-to-end timing of operator* that bench.py reports as e2e_cpp."""
    src = os.path.join(ROOT, "tools", "bench_cpp.cpp")
    inc = os.path.join(PKG, "include")
    build_native()
    newest = max(os.path.getmtime(os.path.join(dp, f)) for dp, _, fs in os.walk(inc) for f in fs)
    newest = max(newest, os.path.getmtime(src), os.path.getmtime(os.path.join(ROOT, "include", "obk_b200.h")), os.path.getmtime(LIB))
    if not force and os.path.exists(BENCH_CPP) and os.path.getmtime(BENCH_CPP) >= newest:
        return BENCH_CPP
    os.makedirs(os.path.dirname(BENCH_CPP), exist_ok=True)
    cmd = ["g++", "-std=c++20", "-O2", "-Wall", "-I", inc, "-o", BENCH_CPP, src, "-L", os.path.join(PKG, "lib"), "-lobk_b200", "-pthread",
           "-Wl,-rpath,$ORIGIN/../../obake_b200/lib"]
    subprocess.check_call(cmd)
    return BENCH_CPP


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose=True))
