import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from obake_b200 import engine, workloads as wl, kpack
from test_engine_gpu import dense_poly
e = engine.Engine(0)
rng = np.random.default_rng(31)
for cfg in ((3, 9, 5, 0.9, 20, "u64", True), (3, 9, 5, 1.0, 20, "u64", False), (3, 3, 2, 1.0, 5, "u64", False)):
    nvars, max0, maxrest, fill, bits, tn, sgn = cfg
    x = dense_poly(rng, nvars, max0, maxrest, fill, bits, tn, signed_cfs=sgn)
    y = dense_poly(rng, nvars, max0, maxrest, fill, bits, tn, signed_cfs=sgn)
    e.set_option("kernel", 0); ref = e.mul(x, y).as_dict()
    e.set_option("kernel", 2); p = e.mul(x, y); got = p.as_dict()
    print(cfg, "kernel", p.stats["kernel"], "terms", len(got), len(ref), "equal", got == ref)
    if got != ref:
        bad = [k for k in set(got) | set(ref) if got.get(k) != ref.get(k)]
        print(" mismatches", len(bad))
        for k in sorted(bad)[:12]:
            print("  ", kpack.unpack(k, nvars, tn), got.get(k), ref.get(k))
