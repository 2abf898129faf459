#!/usr/bin/env python3
"""One GPU playing rank r of N in owner mode (no NCCL needed: ownership is a function of the keys): per-rank kernel and
call times of the zero-exchange product.  usage: owner_probe1.py F30 8 [opt=value ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from obake_b200 import engine, workloads as wl  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "F30"
    world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    eng = engine.Engine(0)
    for a in sys.argv[3:]:
        k, v = a.split("=")
        eng.set_option(k, int(v))
    f, g, limit, _ = wl.named(name)
    lin = None if f.is_f64 else max(engine.needed_limbs(f.cfs), engine.needed_limbs(g.cfs))
    hx, hy = engine.HostOperand(f, lin), engine.HostOperand(g, lin)
    eng.set_option("mgpu_owner", 1)
    tot = 0
    for r in range(world):
        best = None
        for _ in range(3):
            p, _c = eng.mul_partial_views(hx.view, hy.view, r, world, limit, None)
            st = p.stats
            best = st if best is None or st["ms_mul"] < best["ms_mul"] else best
            n = p.nterms
            p.free()
        tot += n
        print(f"rank {r}/{world}: ms_mul {best['ms_mul']:.3f} ms_total {best['ms_total']:.3f} units {best['reserved']} "
              f"tiles/unit {best['table_slots']} stages {best['group_lanes']} terms {n}", flush=True)
    print("terms over all ranks", tot)


if __name__ == "__main__":
    main()
