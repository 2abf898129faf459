"""Multi-GPU parity (needs >= 2 GPUs on the box): torchrun tools/dist_check.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_product_matches_oracle():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=500)
    sys.stdout.write(r.stdout[-3000:])
    sys.stderr.write(r.stderr[-3000:])
    assert r.returncode == 0
    assert "ok=False" not in r.stdout


def test_single_process_multi_device_engine():
    """obk_engine_create_multi: one process, several devices, `f * g` with host operands and a host result (what a C++
    obake program sees).  Dense (tile path, owner mode), sparse, truncated, fp64, wide coefficients -- against the oracle."""
    import numpy as np
    import torch

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    from obake_b200 import engine, workloads as wl
    from test_engine_gpu import as_key_dict, check_f64, check_segments

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    devs = list(range(min(n, 4)))
    eng = engine.Engine(devs)
    try:
        rng = np.random.default_rng(3)
        cases = [wl.fateman(20, 4, "u64") + (None, None), wl.fateman(9, 5, "u64") + (None, None),
                 wl.sparse_mp(8, "i64") + (None, None), wl.sparse_mp(8, "i64") + (40, None), wl.sparse_mp(7, "u64") + (9, [0, 3]),
                 (wl.random_poly(rng, 400, 3, 7, 120, "i64"), wl.random_poly(rng, 300, 3, 7, 100, "i64"), None, None)]
        for f, g, limit, idx in cases:
            p = eng.mul(f, g, limit=limit, idx=idx)
            r = orc.poly_mul(f, g, limit=limit, idx=idx, nthreads=os.cpu_count() or 1)
            assert p.as_dict() == as_key_dict(r["keys"], r["cfs"])
            assert len(p.as_dict()) == p.nterms
            check_segments(p)
        fa, ga = wl.audi("u64", 0, nvars=6, n=6)
        check_f64(eng, fa, ga, limit=6)
    finally:
        eng.close()


def test_cpp_operator_star_on_several_devices():
    """The C++ mirror with OBAKE_B200_DEVICES=0,1: obake::operator* spreads over the devices from one process."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_cpp_api import build_exe

    exe = build_exe()
    env = dict(os.environ, OBAKE_B200_DEVICES="0,1")
    r = subprocess.run([exe, "--multi"], capture_output=True, text=True, timeout=580, env=env)
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0 and "OK:" in r.stdout and "2 devices" in r.stdout
