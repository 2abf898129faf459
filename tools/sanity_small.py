#!/usr/bin/env python3
"""Small end-to-end run touching every kernel (for compute-sanitizer): both product kernels, truncation
(one- and two-sided), d_packed keys, fp64, two-limb inputs, merge."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from obake_b200 import _capi, build, engine, workloads as wl  # noqa: E402


def main():
    build.build_native()
    e = engine.Engine(0)
    rng = np.random.default_rng(5)
    f, g = wl.fateman(9, 4, "u64")
    e.set_option("kernel", 1)
    p = e.mul(f, g)
    assert p.stats["kernel"] == 1 and p.nterms == 7315
    e.set_option("kernel", 0)
    assert e.mul(f, g).nterms == 7315
    e.set_option("kernel", -1)
    fs, gs = wl.sparse_mp(6, "i64")
    assert e.mul(fs, gs).nterms > 0
    e.mul(fs, gs, limit=15)
    e.mul(fs, gs, limit=9, idx=[0, 3])
    fa, ga = wl.audi("u64", 0, nvars=6, n=6)
    e.set_option("two_sided", 2)
    e.mul(fa, ga, limit=6)
    e.mul(fa.with_layout("u64", 4), ga.with_layout("u64", 4), limit=6)
    e.set_option("two_sided", 1)
    x = wl.random_poly(rng, 300, 3, 6, 90, "u64")
    y = wl.random_poly(rng, 300, 3, 6, 80, "u64")
    e.mul(x, y)
    # merge of a device list with duplicates
    hx = engine.HostOperand(wl.Operand(fs.symbols, np.concatenate([fs.expos, fs.expos]),
                                       np.concatenate([fs.cfs, fs.cfs]), "i64"), 2)
    m = e.merge_view(hx.view, 0, _capi.OBK_MEM_HOST)
    assert m.nterms == fs.nterms
    # rectangular
    r = wl.rect_base("u64", True)
    big = wl.random_poly(rng, 5000, 3, 30, f64=True)
    big = wl.Operand(r.symbols, big.expos, big.cfs, "u64")
    e.mul(r, big)
    e.close()
    print("sanity ok")


if __name__ == "__main__":
    main()
