"""The steps either side of the product on the engine's layout: series addition / subtraction
(series.hpp:2195-2991), truncate_degree / truncate_p_degree (polynomial.hpp:2364-2445) and device-resident chains
(benchmark/dense.hpp:28-32).  Checked against the pure-Python oracle (oracle/pyoracle.py) through the C ABI."""
import os
import sys

import numpy as np
import pytest

from obake_b200 import _capi, kpack, workloads as wl
from obake_b200.engine import HostOperand

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import pyoracle  # noqa: E402

pytestmark = pytest.mark.gpu


def packed(d: dict, op) -> dict:
    """{exponent tuple: c} -> {key (int, or tuple of W words): c} in the operand's layout."""
    if not d:
        return {}
    expos = np.array(list(d.keys()), dtype=np.int64).reshape(len(d), op.nvars)
    keys = wl.Operand(op.symbols, expos, np.zeros(len(d)), op.tname, op.psize).keys
    if keys.shape[1] == 1:
        return dict(zip(keys[:, 0].tolist(), d.values()))
    return {tuple(r): c for r, c in zip(keys.tolist(), d.values())}


@pytest.mark.parametrize("tn,psize", [("u64", 0), ("i64", 0), ("u64", 3)])
def test_addsub_exact(engine, tn, psize):
    rng = np.random.default_rng(5)
    x = wl.random_poly(rng, 3000, 5, 9, 60, tn, psize)
    y = wl.random_poly(rng, 2500, 5, 9, 60, tn, psize)
    # make about a third of y's terms cancel against / coincide with x's
    xd, yd = x.to_dict(), y.to_dict()
    for i, (e, c) in enumerate(list(xd.items())[:1500]):
        yd[e] = -c if i % 3 == 0 else (c if i % 3 == 1 else c + 7)
    y = wl.from_dict(x.symbols, yd, tn, psize)
    for subtract in (False, True):
        p = engine.sub(x, y) if subtract else engine.add(x, y)
        ref = packed(pyoracle.poly_addsub(xd, yd, subtract), x)
        got = p.as_dict()
        assert len(got) == p.nterms
        assert got == ref
        assert p.seg_kind == 0  # host results come grouped the way a series stores them


def test_addsub_limb_growth_and_empties(engine):
    x = wl.from_dict(("a", "b"), {(1, 0): 2 ** 62, (0, 1): -(2 ** 62), (2, 2): 5}, "u64")
    y = wl.from_dict(("a", "b"), {(1, 0): 2 ** 62, (0, 1): -(2 ** 62), (3, 0): -1}, "u64")
    p = engine.add(x, y)
    assert p.as_dict() == packed({(1, 0): 2 ** 63, (0, 1): -(2 ** 63), (2, 2): 5, (3, 0): -1}, x)
    assert p.cf_limbs == 2
    e0 = wl.from_dict(("a", "b"), {}, "u64")
    assert engine.add(x, e0).as_dict() == packed(x.to_dict(), x)
    assert engine.sub(e0, x).as_dict() == packed({k: -v for k, v in x.to_dict().items()}, x)
    assert engine.add(e0, e0).nterms == 0
    assert engine.sub(x, x).nterms == 0
    # beyond three limbs: sums of 190-bit and 250-bit coefficients are carried in five limbs
    rng = np.random.default_rng(8)
    for bits in (190, 250, 300):
        u = wl.random_poly(rng, 300, 3, 7, bits, "i64")
        v = wl.random_poly(rng, 300, 3, 7, bits - 7, "i64")
        for subtract in (False, True):
            p = engine.sub(u, v) if subtract else engine.add(u, v)
            assert p.as_dict() == packed(pyoracle.poly_addsub(u.to_dict(), v.to_dict(), subtract), u)
            assert p.cf_limbs == 5


def test_addsub_f64(engine):
    rng = np.random.default_rng(6)
    x = wl.random_poly(rng, 2000, 4, 8, 20, "u64", f64=True)
    xd = x.to_dict()
    yd = {e: (-c if i % 2 else 0.5 * c) for i, (e, c) in enumerate(xd.items())}
    y = wl.from_dict(x.symbols, yd, "u64", f64=True)
    got = engine.add(x, y).as_dict()
    ref = packed(pyoracle.poly_addsub(xd, yd), x)
    assert got == ref  # one rounding per term: bit-identical
    got = engine.sub(x, y).as_dict()
    assert got == packed(pyoracle.poly_addsub(xd, yd, True), x)


@pytest.mark.parametrize("tn", ["u64", "i64"])
def test_truncate_degree(engine, tn):
    rng = np.random.default_rng(8)
    x = wl.random_poly(rng, 5000, 5, 12, 40, tn, signed_expos=(tn == "i64"))
    xd = x.to_dict()
    for lim in (100, 30, 12, 0, -3):
        p = engine.truncate_degree(x, lim)
        # unsigned T: a negative limit converts to a huge unsigned value and keeps everything (C++ conversions)
        eff = lim if (tn == "i64" or lim >= 0) else 2 ** 64 - 1
        assert p.as_dict() == packed(pyoracle.truncate_degree(xd, eff), x), (tn, lim)
    for idx in ([0, 3], [4], []):
        p = engine.truncate_degree(x, 9, idx=idx)
        assert p.as_dict() == packed(pyoracle.truncate_degree(xd, 9, idx), x), idx


@pytest.mark.parametrize("tn,psize", [("u64", 0), ("i64", 0), ("u64", 2), ("i64", 4)])
def test_extend_symbols(engine, tn, psize):
    """series_sym_extender on the device: keys re-packed over a larger symbol set (series.hpp:539-625)."""
    rng = np.random.default_rng(12)
    x = wl.random_poly(rng, 1500, 3, 11, 30, tn, psize, signed_expos=(tn == "i64"))
    pos, nn = [1, 2, 5], 7
    big = wl.Operand(tuple("abcdefg"), np.zeros((1, nn), dtype=np.int64), np.zeros(1), tn, psize)
    p = engine.extend_symbols(x, nn, pos)
    assert p.key_words == big.key_words
    assert p.as_dict() == packed(pyoracle.extend_symbols(x.to_dict(), nn, pos), big)
    # identity extension and device-resident operand
    hx = HostOperand(x)
    d = engine.extend_symbols_view(hx.view, 3, [0, 1, 2], result_mem=_capi.OBK_MEM_DEVICE)
    assert d.on_device and d.as_dict() == packed(x.to_dict(), x)
    d.free()
    # an exponent that fits 3 packed symbols but not 7 (per word: psize symbols): the reference's kpacker throws
    if psize == 0:
        lim7 = kpack.lims(tn, nn)[1]
        assert kpack.lims(tn, 3)[1] > lim7
        y = wl.from_dict(("a", "b", "c"), {(lim7 + 1, 0, 1): 3, (1, 1, 1): 2}, tn)
        with pytest.raises(OverflowError):
            engine.extend_symbols(y, nn, pos)
        with pytest.raises(ValueError):
            engine.extend_symbols(y, kpack.tables(tn)["max_size"] + 1, pos)  # too many symbols for one word
    with pytest.raises(ValueError):
        engine.extend_symbols(x, nn, [2, 1, 5])


def test_product_after_extension(engine):
    """x(a,b) * y(b,c): both operands extended to (a,b,c) on the device, then multiplied without leaving HBM."""
    rng = np.random.default_rng(13)
    x = wl.random_poly(rng, 800, 2, 40, 20, "u64")
    y = wl.random_poly(rng, 700, 2, 40, 20, "u64")
    hx, hy = HostOperand(x), HostOperand(y)
    dx = engine.extend_symbols_view(hx.view, 3, [0, 1], result_mem=_capi.OBK_MEM_DEVICE)
    dy = engine.extend_symbols_view(hy.view, 3, [1, 2], result_mem=_capi.OBK_MEM_DEVICE)
    like = wl.Operand(("a", "b", "c"), np.zeros((1, 3), dtype=np.int64), np.zeros(1), "u64")
    hl = HostOperand(like)
    p = engine.mul_views(engine.view_of(dx, hl.view), engine.view_of(dy, hl.view))
    ref = pyoracle.poly_mul(pyoracle.extend_symbols(x.to_dict(), 3, [0, 1]), pyoracle.extend_symbols(y.to_dict(), 3, [1, 2]))
    assert p.as_dict() == packed(ref, like)
    for q in (dx, dy, p):
        q.free()


def test_device_resident_chain(engine):
    """f *= g twice, + g, truncated: every intermediate stays in HBM; only the last result is read back."""
    f, g = wl.fateman(6, 3, "u64")
    hf, hg = HostOperand(f), HostOperand(g)
    p1 = engine.mul_views(hf.view, hg.view, result_mem=_capi.OBK_MEM_DEVICE)
    assert p1.on_device
    p2 = engine.mul_views(engine.view_of(p1, hf.view), hg.view, result_mem=_capi.OBK_MEM_DEVICE)
    # the accumulator width of p2 differs from g's input width: widen g through the sum's limb rule by adding
    s = engine.addsub_views(engine.view_of(p2, hf.view), hg.view, False, result_mem=_capi.OBK_MEM_DEVICE)
    t = engine.truncate_degree_view(engine.view_of(s, hf.view), 10, result_mem=_capi.OBK_MEM_HOST)
    fd, gd = f.to_dict(), g.to_dict()
    ref = pyoracle.poly_mul(pyoracle.poly_mul(fd, gd), gd)
    ref = pyoracle.truncate_degree(pyoracle.poly_addsub(ref, gd), 10)
    assert t.as_dict() == packed(ref, f)
    for p in (p1, p2, s, t):
        p.free()
