"""Multi-GPU decomposition checked on ONE GPU: the device plays rank 0 .. N-1 in turn (obk_poly_mul_partial), the
shares are put together on the host and compared with the oracle's product.

Ownership in both multi-GPU designs is a function of the keys alone, so no second
device and no NCCL are needed to check that
  * owner mode: the ranks' shares are pairwise disjoint and their union is exactly the product, whatever kernel takes
    the product (tile path: canonical deal of output planes; scatter kernel: contiguous ranges of output segments,
    one wave of CTAs PER RANK -- the segmentation depends on the number of ranks);
  * partial mode (left operand sharded, the producer side of the NCCL exchange / the peer-window push): the partial
    products add up, key by key, to the product, and every partial term sits in the owner group the exchange assumes.
tools/dist_check.py repeats this with real peers (profiles/r02_dist_check_*gpu.log)."""
import os

import numpy as np
import pytest

import orc
from obake_b200 import _capi, engine as eng_mod, workloads as wl
from test_engine_gpu import as_key_dict, check_segments

pytestmark = pytest.mark.gpu
NT = os.cpu_count() or 1


def _cases():
    rng = np.random.default_rng(11)
    fa, ga = wl.audi("u64", 0, nvars=6, n=6)
    fd, gd = wl.audi("u64", 8, nvars=10, n=4)
    return [
        ("fateman12 (tile path)", *wl.fateman(12, 4, "u64"), None, None),
        ("fateman8x5 (tile path, 5 symbols)", *wl.fateman(8, 5, "u64"), None, None),
        ("sparse7 (scatter)", *wl.sparse_mp(7, "i64"), None, None),
        ("sparse7 truncated 30", *wl.sparse_mp(7, "i64"), 30, None),
        ("sparse7 partial degree", *wl.sparse_mp(7, "u64"), 9, [0, 3]),
        ("random signed", wl.random_poly(rng, 500, 3, 7, 120, "i64"), wl.random_poly(rng, 300, 3, 7, 100, "i64"), None, None),
        ("short left operand", wl.random_poly(rng, 3, 4, 7, 50, "u64"), wl.random_poly(rng, 900, 4, 7, 50, "u64"), None, None),
        ("audi 6 vars fp64, two-sided", fa, ga, 6, None),
        ("audi d_packed fp64", fd, gd, 4, None),
    ]


def _views(f, g):
    lin = None if f.is_f64 else max(eng_mod.needed_limbs(f.cfs), eng_mod.needed_limbs(g.cfs))
    return eng_mod.HostOperand(f, lin), eng_mod.HostOperand(g, lin)


def _compare(name, f, got, ref, bound):
    """Bit-exact for integers; fp64: relative 1e-12 per coefficient, a key present on one side only must be an exact
    cancellation's residue (below 1e-13 of the largest possible sum)."""
    if not f.is_f64:
        assert got == ref, name
        return
    for k in set(got) | set(ref):
        a, b = got.get(k, 0.0), ref.get(k, 0.0)
        if k in got and k in ref:
            assert abs(a - b) <= 1e-12 * max(abs(a), abs(b)) + 1e-15 * bound, (name, k, a, b)
        else:
            assert abs(a - b) <= 1e-13 * bound, (name, k, a, b)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_owner_mode_shares_partition_the_product(engine, world):
    engine.set_option("mgpu_owner", 1)
    try:
        for name, f, g, limit, idx in _cases():
            hx, hy = _views(f, g)
            r = orc.poly_mul(f, g, limit=limit, idx=idx, nthreads=NT)
            assert r["status"] == 0
            ref = as_key_dict(r["keys"], r["cfs"])
            union, nsum, kernels = {}, 0, set()  # kernels: of the ranks that formed something
            for rank in range(world):
                p, _counts = engine.mul_partial_views(hx.view, hy.view, rank, world, limit, idx)
                share = p.as_dict()
                assert len(share) == p.nterms, (name, "duplicate keys inside a share")
                nsum += p.nterms
                if p.nterms:
                    kernels.add(p.stats["kernel"])
                union.update(share)
                if world == 2:
                    # the download bench.py's multi-GPU e2e leg uses: obk_poly_copy of the share into host series tables
                    hp = engine.copy_view(engine.view_of(p, hx.view), _capi.OBK_MEM_HOST)
                    assert not hp.on_device and hp.as_dict() == share, name
                    check_segments(hp)
                    hp.free()
                p.free()
            assert len(kernels) <= 1, (name, "the ranks must take the same kernel decision", kernels)
            assert nsum == len(union), (name, "two ranks hold the same key")
            bound = 0.0 if not f.is_f64 else float(np.abs(f.cfs).max() * np.abs(g.cfs).max()) * min(f.nterms, g.nterms)
            _compare(name, f, union, ref, bound)
    finally:
        engine.set_option("mgpu_owner", 0)


@pytest.mark.parametrize("world", [2, 5])
def test_partial_products_add_up_and_sit_in_their_owner_groups(engine, world):
    m = _capi.OBK_EXCHANGE_MOD
    for name, f, g, limit, idx in _cases():
        hx, hy = _views(f, g)
        r = orc.poly_mul(f, g, limit=limit, idx=idx, nthreads=NT)
        ref = as_key_dict(r["keys"], r["cfs"])
        total = {}
        for rank in range(world):
            p, counts = engine.mul_partial_views(hx.view, hy.view, rank, world, limit, idx)
            assert sum(counts) == p.nterms, name
            if p.nterms:
                keys = p.keys
                grp = (keys % np.uint64(m)).sum(axis=1) % np.uint64(m)
                if not f.tname.startswith("i"):
                    # group s belongs to the rank whose range [m*g/world, m*(g+1)/world) holds it (dist.owner_ranges);
                    # rows destined to one owner are contiguous, in owner order: counts[] cuts the list
                    bounds = np.array([m * gidx // world for gidx in range(world + 1)], dtype=np.int64)
                    owner = np.searchsorted(bounds, grp.astype(np.int64), side="right") - 1
                    assert (np.diff(owner) >= 0).all(), (name, "partial terms are not grouped by owner")
                    assert np.bincount(owner, minlength=world).tolist() == counts, name
            for k, c in p.as_dict().items():
                total[k] = total.get(k, 0) + c
            p.free()
        if f.is_f64:
            bound = float(np.abs(f.cfs).max() * np.abs(g.cfs).max()) * min(f.nterms, g.nterms)
            total = {k: c for k, c in total.items() if c != 0.0 or k in ref}
            _compare(name, f, total, ref, bound)
        else:
            total = {k: c for k, c in total.items() if c != 0}
            assert total == ref, name
