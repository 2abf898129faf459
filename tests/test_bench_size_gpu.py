"""Parity at the sizes bench.py times ("Next 1"): the exact configurations of BASELINE.md section 2,
checked without the CPU oracle where it cannot finish in seconds -- Fateman / dense_02 coefficient by coefficient
against the closed form (benchmark/dense_4_vars.cpp:28-48, dense_02.cpp), the sparse pair by the term
counts the reference pins plus the evaluation homomorphism over three primes (benchmark/sparse.hpp,
sparse_02_truncated.cpp:24-67), audi_01 against its closed form AND the oracle (benchmark/audi_01.cpp:44-68),
rectangular_01 by its final term count and f(1)^70 (benchmark/rectangular_01.cpp:42-47)."""
import numpy as np
import pytest

import orc
from obake_b200 import _capi, verify as V, workloads as wl
from obake_b200.engine import raw_view

pytestmark = pytest.mark.gpu


def _check(engine, name, kernel=-1, opts=None):
    f, g, limit, _ = wl.named(name)
    opts = dict(opts or {})
    opts["kernel"] = kernel
    try:
        for k, v in opts.items():
            engine.set_option(k, v)
        p = engine.mul(f, g, limit=limit)
    finally:
        engine.set_option("kernel", -1)
        for k in opts:
            if k == "fence":
                engine.set_option("fence", 1)
    cfs = p.cfs if p.is_f64 else p.cf_limb_array
    res = V.check_product(name, f, g, limit, p.keys, cfs, p.is_f64)
    assert res["checked"] and res["ok"], res
    return p, res, (f, g, limit)


@pytest.mark.parametrize("kernel", [-1, 0])
def test_f30_closed_form(engine, kernel):
    """The bench's headline launch: 2.15e9 products, 3-limb accumulators, 635 376 terms, every coefficient exact;
    through the dense path (auto) and forced onto the segment-owner scatter kernel."""
    p, res, _ = _check(engine, "F30", kernel)
    assert p.nterms == 635376 and p.stats["acc_limbs"] == 3 and res["closed_form_mismatches"] == 0
    assert p.stats["kernel"] == (0 if kernel == 0 else p.stats["kernel"])
    if kernel == -1:
        assert p.stats["kernel"] >= 1, "F30 did not take a dense path"


def test_f20_and_d5_closed_form(engine):
    p, res, _ = _check(engine, "F20")
    assert p.nterms == 135751
    p, res, _ = _check(engine, "D5")
    assert p.nterms == 237336 and p.stats["acc_limbs"] in (1, 2)  # values fit 62 bits; the a-priori bound says 2 limbs


@pytest.mark.parametrize("fence", [1, 0])
def test_s12_checksum_both_lock_protocols(engine, fence):
    """5.8 M distinct terms through the shared-memory slot lock: the memory-model form (default) and the
    hardware-order form (option fence=0) must both give the exact product."""
    p, res, _ = _check(engine, "S12", opts={"fence": fence})
    assert p.nterms == 5821335 and p.stats["kernel"] == 0
    assert sum(V.ints_from_limbs(p.cf_limb_array[:1000])) > 0


@pytest.mark.parametrize("fence", [1, 0])
def test_s16t_count_and_checksum(engine, fence):
    """sparse_02_truncated at full size: 28 398 035 terms, no zero / duplicate term, evaluation
    checksum over three primes (implies sum(cf) == 13^32)."""
    p, res, _ = _check(engine, "S16T", opts={"fence": fence})
    assert p.nterms == 28398035
    p.free()


@pytest.mark.parametrize("name", ["AUDI", "AUDID"])
def test_audi_full_size(engine, name):
    p, res, (f, g, limit) = _check(engine, name)
    assert p.nterms == 184756 and res["max_rel_err"] <= 1e-12
    # and against the oracle on the same inputs (0.5 s of CPU): same key set, relative 1e-12 per coefficient
    r = orc.poly_mul(f, g, limit=limit, nthreads=8)
    got = p.as_dict()
    rk = r["keys"]
    ref = dict(zip(rk[:, 0].tolist(), r["cfs"].tolist())) if rk.shape[1] == 1 else \
        {tuple(k): c for k, c in zip(rk.tolist(), r["cfs"].tolist())}
    assert set(got) == set(ref)
    assert all(abs(got[k] - ref[k]) <= 1e-12 * abs(ref[k]) for k in ref)


def test_rect_chain_final(engine):
    """rectangular_01: curr *= f seventy times with every intermediate resident in HBM."""
    f = wl.rect_base("u64", True)
    cur = None
    one = wl.from_dict(f.symbols, {(0, 0, 0): 1.0}, "u64", 0, True)
    from obake_b200.engine import HostOperand

    hf, h1 = HostOperand(f), HostOperand(one)
    vcur = h1.view
    for i in range(70):
        nxt = engine.mul_views(hf.view, vcur, result_mem=_capi.OBK_MEM_DEVICE)
        if cur is not None:
            cur.free()
        cur = nxt
        vcur = raw_view(cur.nterms, 1, 0, 3, 0, _capi.OBK_CF_F64, 1, _capi.OBK_MEM_DEVICE, cur.keys_ptr, cur.cfs_ptr)
    assert cur.nterms == 1284816
    tot = float(cur.cfs.sum())
    assert abs(tot - float(16 ** 70)) <= 1e-12 * float(16 ** 70)
    assert len(np.unique(cur.keys[:, 0])) == cur.nterms
    cur.free()


def test_addsub_large_disjoint_support(engine):
    """a + b with (almost) no common monomial fills the merge tables to the sizing bound; the finer
    re-segmentation retry must make it succeed."""
    rng = np.random.default_rng(77)
    n = 400000
    ea = rng.integers(0, 2000, size=(n, 3))
    eb = rng.integers(0, 2000, size=(n, 3)) + np.array([2001, 0, 0])
    ea = np.unique(ea, axis=0)
    eb = np.unique(eb, axis=0)
    a = wl.Operand(("x", "y", "z"), ea.astype(np.int64), np.array([int(v) + 1 for v in range(len(ea))], dtype=object), "u64")
    b = wl.Operand(("x", "y", "z"), eb.astype(np.int64), np.array([-(int(v) + 1) for v in range(len(eb))], dtype=object), "u64")
    p = engine.add(a, b)
    assert p.nterms == len(ea) + len(eb)
    d = p.as_dict()
    ka = a.keys[:, 0].tolist()
    assert all(d[k] == i + 1 for i, k in list(enumerate(ka))[:: 997])
    p.free()
