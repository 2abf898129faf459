#!/usr/bin/env python3
"""Option sweep for the product kernel: prints ms_mul per configuration (development tool)."""
import itertools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from obake_b200 import _capi, build, engine, workloads as wl  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "F30"
    grid = {}
    for a in sys.argv[2:]:
        k, v = a.split("=")
        grid[k] = [int(x) for x in v.split(",")]
    build.build_native()
    e = engine.Engine(0)
    f, g, limit, desc = wl.named(name)
    lin = None if f.is_f64 else max(engine.needed_limbs(f.cfs), engine.needed_limbs(g.cfs))
    hx, hy = engine.HostOperand(f, lin), engine.HostOperand(g, lin)
    keys = list(grid)
    for combo in itertools.product(*[grid[k] for k in keys]):
        for k, v in zip(keys, combo):
            e.set_option(k, v)
        best = None
        try:
            for _ in range(3):
                p = e.mul_views(hx.view, hy.view, limit=limit, result_mem=_capi.OBK_MEM_DEVICE)
                st = p.stats
                best = st if best is None or st["ms_mul"] < best["ms_mul"] else best
                p.free()
            print(dict(zip(keys, combo)), "ms_mul %.3f ms_total %.3f nsegs %d slots %d slices %d retries %d Gprod/s %.2f" % (
                best["ms_mul"], best["ms_total"], best["nsegs"], best["table_slots"], best["group_lanes"], best["retries"],
                best["term_products"] / best["ms_mul"] / 1e6), flush=True)
        except Exception as ex:  # noqa: BLE001
            print(dict(zip(keys, combo)), "FAILED", ex, flush=True)


if __name__ == "__main__":
    main()
