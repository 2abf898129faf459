"""Parity tests proper: the CUDA engine (through the C ABI) against the CPU oracle on the same inputs,
the reference's known answers, and size-independent properties at the benchmark sizes."""
import json
import os

import numpy as np
import pytest

import orc
from obake_b200 import kpack, workloads as wl

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
KA = json.load(open(os.path.join(HERE, "golden", "known_answers.json")))
NT = os.cpu_count() or 1


def as_key_dict(keys, cfs):
    if keys.shape[1] == 1:
        return dict(zip(keys[:, 0].tolist(), cfs if isinstance(cfs, list) else cfs.tolist()))
    return {tuple(r): c for r, c in zip(keys.tolist(), cfs if isinstance(cfs, list) else cfs.tolist())}


def check_exact(engine, x, y, limit=None, idx=None, nthreads=NT):
    p = engine.mul(x, y, limit=limit, idx=idx)
    r = orc.poly_mul(x, y, limit=limit, idx=idx, nthreads=nthreads)
    assert r["status"] == 0
    got = p.as_dict()
    ref = as_key_dict(r["keys"], r["cfs"])
    assert len(got) == p.nterms, "duplicate keys in the engine's result"
    assert got == ref
    assert p.stats["term_products"] == r["n_mults"]
    check_segments(p)
    return p


def check_segments(p):
    """series invariants (series.hpp:866-909): term in table hash & (2^k-1); no zero coefficient."""
    n = p.nterms
    so = p.seg_offsets
    assert so[0] == 0 and so[-1] == n and (np.diff(so.astype(np.int64)) >= 0).all()
    if n == 0:
        return
    keys = p.keys
    h = keys.sum(axis=1, dtype=np.uint64)
    seg = np.searchsorted(so, np.arange(n), side="right") - 1
    assert p.seg_kind == 0 and p.nsegs == 1 << p.log2_nsegs
    assert ((h & np.uint64(p.nsegs - 1)) == seg.astype(np.uint64)).all()
    if p.is_f64:
        assert (p.cfs != 0).all()
    else:
        assert (p.cf_limb_array != 0).any(axis=1).all()


def check_f64(engine, x, y, limit=None, idx=None, rtol=1e-12):
    """fp64 parity: same key set; |got - ref| <= rtol * sum|c1*c2| bound (BASELINE.json: relative 1e-12)."""
    p = engine.mul(x, y, limit=limit, idx=idx)
    r = orc.poly_mul(x, y, limit=limit, idx=idx, nthreads=NT)
    got, ref = p.as_dict(), as_key_dict(r["keys"], r["cfs"])
    scale = float(np.abs(x.cfs).max() * np.abs(y.cfs).max()) * min(x.nterms, y.nterms)
    for k in set(got) | set(ref):
        a, b = got.get(k, 0.0), ref.get(k, 0.0)
        if k in got and k in ref:
            assert abs(a - b) <= rtol * max(abs(b), abs(a)) + 1e-15 * scale, (k, a, b)
        else:
            # exact cancellation in one summation order may leave a tiny residue in the other (Appendix A.5)
            assert abs(a - b) <= 1e-13 * scale, (k, a, b)
    check_segments(p)
    return p


# ---------------------------------------------------------------------------------------------------------
def test_small_identities(engine):
    # test/polynomials_polynomial_00.cpp:198-214
    sy = ("a", "b")
    three = wl.from_dict(sy, {(0, 0): 3}, "i64")
    four = wl.from_dict(sy, {(0, 0): 4}, "i64")
    p = engine.mul(three, four)
    assert p.as_dict() == {0: 12}
    apb = wl.from_dict(sy, {(1, 0): 1, (0, 1): 1}, "i64")
    amb = wl.from_dict(sy, {(1, 0): 1, (0, 1): -1}, "i64")
    p = check_exact(engine, apb, amb)
    assert p.nterms == 2  # a^2 - b^2: the cancelled cross term is erased
    a2b2 = wl.from_dict(sy, {(2, 0): 1, (0, 2): 1}, "i64")
    a2mb2 = wl.from_dict(sy, {(2, 0): 1, (0, 2): -1}, "i64")
    p = check_exact(engine, a2b2, a2mb2)
    assert sorted(p.cfs) == [-1, 1]


def test_empty_operands(engine):
    sy = ("a", "b")
    x = wl.from_dict(sy, {(1, 0): 1}, "i64")
    e = wl.Operand(sy, np.zeros((0, 2), dtype=np.int64), np.zeros(0, dtype=object), "i64")
    for a, b in ((x, e), (e, x), (e, e)):
        p = engine.mul(a, b)
        assert p.nterms == 0 and p.seg_offsets.tolist() == [0, 0]


def test_zero_variables(engine):
    x = wl.Operand((), np.zeros((1, 0), dtype=np.int64), np.array([7], dtype=object), "u64")
    y = wl.Operand((), np.zeros((1, 0), dtype=np.int64), np.array([-6], dtype=object), "u64")
    assert engine.mul(x, y).as_dict() == {0: -42}


@pytest.mark.parametrize("tn", ["i64", "u64"])
def test_random_vs_oracle(engine, tn):
    rng = np.random.default_rng(11)
    for nv, n1, n2, mx, bits in ((3, 40, 60, 6, 30), (6, 300, 100, 3, 20), (1, 50, 50, 1000, 40), (5, 2000, 1500, 9, 31),
                                 (21, 500, 400, 1, 10)):
        x = wl.random_poly(rng, n1, nv, mx, bits, tn, signed_expos=(tn == "i64" and mx <= 3))
        y = wl.random_poly(rng, n2, nv, mx, bits, tn, signed_expos=(tn == "i64" and mx <= 3))
        check_exact(engine, x, y)
        check_exact(engine, y, x)  # operand order (Appendix A.7)


def test_squaring_same_operand(engine):
    rng = np.random.default_rng(12)
    x = wl.random_poly(rng, 500, 4, 7, 25, "u64")
    check_exact(engine, x, x)


def test_accumulator_widths(engine):
    rng = np.random.default_rng(13)
    for bits, want in ((8, 1), (30, 2), (45, 2), (61, 3), (63, 3)):
        x = wl.random_poly(rng, 300, 2, 20, bits, "u64")
        y = wl.random_poly(rng, 300, 2, 20, bits, "u64")
        p = check_exact(engine, x, y)
        assert p.cf_limbs == want, (bits, p.cf_limbs)
    # all-positive 61-bit dense product: sums really exceed 128 bits
    x = wl.random_poly(rng, 400, 2, 19, 61, "u64", signed_cfs=False)
    p = check_exact(engine, x, x)
    assert max(abs(c) for c in p.cfs) >= 1 << 127


def test_two_limb_inputs(engine):
    rng = np.random.default_rng(14)
    x = wl.random_poly(rng, 200, 3, 6, 90, "u64")
    y = wl.random_poly(rng, 200, 3, 6, 80, "u64")
    p = check_exact(engine, x, y)
    assert p.cf_limbs == 3


def test_coefficient_overflow_is_detected(engine):
    rng = np.random.default_rng(15)
    x = wl.random_poly(rng, 50, 2, 9, 185, "u64")
    y = wl.random_poly(rng, 50, 2, 9, 185, "u64")
    with pytest.raises(OverflowError, match="coefficient overflow beyond 5 limbs"):
        engine.mul(x, y)
    z = wl.random_poly(rng, 50, 2, 9, 200, "u64")
    with pytest.raises((OverflowError, RuntimeError, ValueError), match="beyond 3 limbs"):
        engine.mul(z, z)


def test_wide_coefficients(engine):
    """Coefficient growth along chains (the reference's integer<N> grows without bound, fma3.hpp:84-88): two- and
    three-limb operands, four- and five-limb exact sums -- scatter kernel, both signs, sparse and dense keys."""
    rng = np.random.default_rng(16)
    for bx, by, want in ((120, 120, 5), (90, 125, 5), (150, 20, 5), (185, 100, 5), (130, 130, 5), (100, 60, 3)):
        x = wl.random_poly(rng, 300, 3, 7, bx, "i64")
        y = wl.random_poly(rng, 250, 3, 7, by, "i64")
        p = check_exact(engine, x, y)
        assert p.cf_limbs == want, (bx, by, p.cf_limbs)
    # all-positive sums that really fill the fifth limb
    x = wl.random_poly(rng, 400, 2, 19, 128, "u64", signed_cfs=False)
    y = wl.random_poly(rng, 400, 2, 19, 120, "u64", signed_cfs=False)
    p = check_exact(engine, x, y)
    assert p.cf_limbs == 5 and max(abs(c) for c in p.cfs) >= 1 << 250
    # truncated, d_packed keys
    x = wl.random_poly(rng, 500, 6, 5, 140, "u64").with_layout("u64", 3)
    y = wl.random_poly(rng, 500, 6, 5, 70, "u64").with_layout("u64", 3)
    check_exact(engine, x, y, limit=14)
    # a chain f *= f past the old three-limb ceiling: (1 + x + y + z)^n, n = 16 -> 32 -> 64 -> 128 by repeated squaring,
    # every step against the closed form; the last step multiplies two-limb operands into five-limb coefficients
    def closed(n):
        return {tuple(int(v) for v in e): wl._multinomial(n, tuple(int(v) for v in e)) for e in wl._compositions_le(3, n)}
    sy = ("x", "y", "z")
    cur = wl.from_dict(sy, closed(16), "u64")
    for n in (32, 64, 128):
        p = engine.mul(cur, cur)
        want = closed(n)
        got = {tuple(kpack.unpack(k, 3, "u64")): v for k, v in p.as_dict().items()}
        assert got == want, n
        cur = wl.from_dict(sy, want, "u64")
    assert p.cf_limbs == 5 and max(want.values()).bit_length() > 192
    with pytest.raises(Exception, match="beyond 3 limbs"):
        engine.mul(cur, cur)          # (1+x+y+z)^128 has four-limb coefficients: refused, never wrapped


def test_exponent_overflow_throws(engine):
    # test/polynomials_polynomial_00.cpp:216-238
    msg = KA["overflow_message"]["value"]
    for tn in ("i64", "u64"):
        lo, hi = kpack.lims(tn, 2)
        sy = ("a", "b")
        big = wl.from_dict(sy, {(hi, 0): 1, (0, 1): 1}, tn)
        one = wl.from_dict(sy, {(1, 0): 1}, tn)
        with pytest.raises(OverflowError, match=msg):
            engine.mul(big, one)
        assert not engine.overflow_check(big, one)
        assert engine.overflow_check(one, one)
        ok = wl.from_dict(sy, {(hi - 1, 0): 1, (0, 1): 1}, tn)
        check_exact(engine, ok, one)
    lo, hi = kpack.lims("i64", 2)
    small = wl.from_dict(("a", "b"), {(lo, 0): 1}, "i64")
    neg = wl.from_dict(("a", "b"), {(-1, 0): 1}, "i64")
    with pytest.raises(OverflowError, match=msg):
        engine.mul(small, neg)


def test_degrees_and_overflow_entry_points(engine):
    rng = np.random.default_rng(16)
    for tn, psize in (("u64", 0), ("i64", 0), ("u64", 4), ("i64", 3)):
        x = wl.random_poly(rng, 3000, 10, 5, 10, tn, psize, signed_expos=tn == "i64")
        assert (engine.degrees(x) == x.expos.sum(axis=1)).all()
        assert (engine.degrees(x, [1, 4, 9]) == x.expos[:, [1, 4, 9]].sum(axis=1)).all()
        assert (engine.degrees(x, []) == 0).all()


def test_fateman_closed_form(engine):
    for n, tn in ((6, "u64"), (10, "i64"), (12, "u64")):
        f, g = wl.fateman(n, 4, tn)
        p = engine.mul(f, g)
        exp = wl.fateman_expected(n)
        expk = {kpack.pack(k, tn) & ((1 << 64) - 1): v for k, v in exp.items()}
        assert p.as_dict() == expk
        assert p.stats["term_products"] == f.nterms * g.nterms


def test_sparse_n10_known_answer(engine):
    # test/polynomials_polynomial_00.cpp:398-424: 2 096 600 terms, integer<1> and double
    f, g = wl.sparse_mp(10, "i64")
    p = check_exact(engine, f, g)
    assert p.nterms == KA["sparse_n10_terms"]["value"]
    pd = engine.mul(f.as_f64(), g.as_f64())
    assert pd.nterms == KA["sparse_n10_terms"]["value"]
    check_segments(pd)


def test_rect_pow20_known_answer(engine):
    # test/polynomials_polynomial_04.cpp:264-285
    base = wl.mp_base("i64")
    cur = base
    for _ in range(19):
        p = engine.mul(base, cur)
        keys = p.keys
        cur = wl.Operand(base.symbols, np.zeros((p.nterms, 5), dtype=np.int64), np.array(p.cfs, dtype=object), "i64", 0,
                         keys)
    assert cur.nterms == KA["rect_pow20_terms"]["value"]


def test_truncated_small(engine):
    # test/polynomials_polynomial_01.cpp:131-183: (x+y)(x-y) and (zx+y)(x-y-1) with limits 100/2/1/0/-1
    sy = ("x", "y", "z")
    a = wl.from_dict(sy, {(1, 0, 0): 1, (0, 1, 0): 1}, "i64")
    b = wl.from_dict(sy, {(1, 0, 0): 1, (0, 1, 0): -1}, "i64")
    c = wl.from_dict(sy, {(1, 0, 1): 1, (0, 1, 0): 1}, "i64")
    d = wl.from_dict(sy, {(1, 0, 0): 1, (0, 1, 0): -1, (0, 0, 0): -1}, "i64")
    for lim in (100, 2, 1, 0, -1):
        check_exact(engine, a, b, limit=lim)
        check_exact(engine, c, d, limit=lim)
        for idx in ([0], [1, 2], []):
            if idx == []:
                # partial degree over no symbol: every degree is 0, nothing is cut unless limit < 0
                p = engine.mul(c, d, limit=lim, idx=idx)
                full = engine.mul(c, d)
                assert p.as_dict() == (full.as_dict() if lim >= 0 else {})
            else:
                check_exact(engine, c, d, limit=lim, idx=idx)


def test_truncated_n8_reference_cases(engine):
    # test/polynomials_polynomial_01.cpp:185-236, _04.cpp:239-262
    f, g = wl.sparse_mp(8, "i64")
    full = check_exact(engine, f, g)
    assert full.nterms == 591235
    fd = full.as_dict()
    for lim in (1000, 80):
        assert engine.mul(f, g, limit=lim).as_dict() == fd
    for lim in (50, 40, 5, 0):
        p = check_exact(engine, f, g, limit=lim)
        deg = engine.degrees(wl.Operand(f.symbols, np.zeros((p.nterms, 5), dtype=np.int64), np.zeros(p.nterms), "i64", 0,
                                        p.keys))
        assert deg.max() == (lim if lim else 0)
    assert engine.mul(f, g, limit=0).as_dict() == {0: 1}
    p = engine.mul(f, g, limit=-1)
    assert p.nterms == 0 and p.stats["term_products"] == 0


def test_partial_degree_truncation(engine):
    # test/polynomials_polynomial_02.cpp:182-327, _05.cpp:105-126
    f, g = wl.sparse_mp(5, "i64")
    for idx in ([0], [1, 3], [0, 1, 2, 3, 4]):
        for lim in (0, 3, 9, 100):
            check_exact(engine, f, g, limit=lim, idx=idx)


def test_unsigned_limit_conversion(engine):
    # limit is passed as the bit pattern of T: a huge unsigned limit never cuts
    f, g = wl.sparse_mp(3, "u64")
    full = engine.mul(f, g).as_dict()
    assert engine.mul(f, g, limit=-1).as_dict() == full  # (uint64)-1
    check_exact(engine, f, g, limit=7)


def test_d_packed_vs_packed(engine):
    rng = np.random.default_rng(17)
    x = wl.random_poly(rng, 800, 10, 4, 20, "u64", 0)
    y = wl.random_poly(rng, 600, 10, 4, 20, "u64", 0)
    ref = orc.result_to_dict(orc.poly_mul(x, y, nthreads=NT), x)
    for tn, psize in (("u64", 8), ("u64", 4), ("i64", 3), ("u64", 10)):
        xd, yd = x.with_layout(tn, psize), y.with_layout(tn, psize)
        p = check_exact(engine, xd, yd)
        r = {"keys": p.keys, "cfs": p.cfs}
        assert orc.result_to_dict(r, xd) == ref
        check_exact(engine, xd, yd, limit=12)
        check_exact(engine, xd, yd, limit=9, idx=[0, 5, 9])


def test_f64_random_and_audi_small(engine):
    rng = np.random.default_rng(18)
    x = wl.random_poly(rng, 1500, 4, 8, f64=True)
    y = wl.random_poly(rng, 1200, 4, 8, f64=True)
    check_f64(engine, x, y)
    f, g = wl.audi("u64", 0, nvars=6, n=6)
    check_f64(engine, f, g, limit=6)
    fd, gd = wl.audi("u64", 4, nvars=6, n=6)
    check_f64(engine, fd, gd, limit=6)


def test_f64_first_contribution_and_signed_zero(engine):
    sy = ("a",)
    x = wl.from_dict(sy, {(0,): 1.5, (1,): -1.5}, "u64", f64=True)
    y = wl.from_dict(sy, {(0,): 2.0, (1,): 2.0}, "u64", f64=True)
    p = engine.mul(x, y)
    assert p.as_dict() == {0: 3.0, 2: -3.0}  # the middle term cancels exactly and is erased


def test_skewed_buckets_force_retry(engine):
    # keys that are all multiples of 1024 (one power-of-two bucket); the prime segmentation does not care
    n = 1500
    e = (np.arange(n, dtype=np.int64) * 1024).reshape(-1, 1)
    x = wl.Operand(("a",), e, np.array([i + 1 for i in range(n)], dtype=object), "u64")
    p = check_exact(engine, x, x)
    assert p.nterms == 2 * n - 1


DEFAULTS = {"nsegs": -1, "refill_min": 16, "table_slots": 0, "threads": 1024, "fence": 1, "fill_pct": 62}


def test_forced_geometry_options(engine):
    rng = np.random.default_rng(19)
    x = wl.random_poly(rng, 1000, 4, 7, 25, "u64")
    y = wl.random_poly(rng, 900, 4, 7, 25, "u64")
    ref = check_exact(engine, x, y).as_dict()
    try:
        for opts in ({"nsegs": 1}, {"nsegs": 128}, {"nsegs": 1000}, {"refill_min": 1}, {"refill_min": 32},
                     {"table_slots": 1024, "threads": 256}, {"threads": 64}, {"fence": 0}, {"fill_pct": 85}):
            for k, v in opts.items():
                engine.set_option(k, v)
            assert engine.mul(x, y).as_dict() == ref, opts
            for k in opts:
                engine.set_option(k, DEFAULTS[k])
    finally:
        for k, v in DEFAULTS.items():
            engine.set_option(k, v)


def test_benchmark_size_properties(engine):
    """F20 at full size: closed form on every coefficient; S12: term count + checksum identity
    sum(cf) == f(1)*g(1) (evaluate at all-ones) and linearity in the left operand."""
    f, g = wl.fateman(20, 4, "u64")
    p = engine.mul(f, g)
    assert p.nterms == 135751
    exp = wl.fateman_expected(20)
    expk = {kpack.pack(k, "u64"): v for k, v in exp.items()}
    assert p.as_dict() == expk
    f, g = wl.sparse_mp(12, "u64")
    p = engine.mul(f, g)
    assert p.nterms == 5821335
    assert sum(p.cfs) == 13 ** 24
    check_segments(p)
    half = f.nterms // 2
    a = engine.mul(f.take(slice(0, half)), g)
    b = engine.mul(f.take(slice(half, None)), g)
    da, db = a.as_dict(), b.as_dict()
    for k, v in db.items():
        da[k] = da.get(k, 0) + v
    assert {k: v for k, v in da.items() if v} == p.as_dict()


def dense_poly(rng, nvars, max0, maxrest, fill, bits, tname="u64", f64=False, signed_cfs=True):
    """Random polynomial that is dense in the first symbol (rows of width <= max0 + 1)."""
    import itertools
    terms = {}
    for rest in itertools.product(range(maxrest + 1), repeat=nvars - 1):
        if rng.random() < 0.7:
            for e0 in range(max0 + 1):
                if rng.random() < fill:
                    if f64:
                        terms[(e0,) + rest] = float(rng.standard_normal())
                    else:
                        v = int(rng.integers(1, 1 << bits))
                        terms[(e0,) + rest] = -v if (signed_cfs and rng.integers(0, 2)) else v
    return wl.from_dict(tuple(f"v{i}" for i in range(nvars)), terms, tname, 0, f64)


def test_row_path_forced(engine):
    """The row-blocked dense kernel (register accumulation per output row) against the oracle."""
    rng = np.random.default_rng(23)
    engine.set_option("kernel", 1)
    try:
        for nvars, max0, maxrest, fill, bits, tn in ((3, 9, 5, 0.9, 20, "u64"), (4, 31, 3, 0.6, 61, "u64"),
                                                      (2, 15, 40, 1.0, 30, "i64"), (3, 31, 6, 0.3, 8, "i64"),
                                                      (5, 4, 2, 1.0, 40, "u64")):
            x = dense_poly(rng, nvars, max0, maxrest, fill, bits, tn)
            y = dense_poly(rng, nvars, max0, maxrest, fill, bits, tn)
            p = check_exact(engine, x, y)
            assert p.stats["kernel"] == 1, "row path was not taken"
            p = check_exact(engine, x, x)
            assert p.stats["kernel"] == 1
        # exact cancellation inside a row: (a+b)(a-b)
        sy = ("a", "b")
        apb = wl.from_dict(sy, {(1, 0): 1, (0, 1): 1}, "i64")
        amb = wl.from_dict(sy, {(1, 0): 1, (0, 1): -1}, "i64")
        p = check_exact(engine, apb, amb)
        assert p.stats["kernel"] == 1 and p.nterms == 2
        # fp64
        x = dense_poly(rng, 3, 12, 6, 0.8, 0, "u64", f64=True)
        y = dense_poly(rng, 3, 12, 6, 0.8, 0, "u64", f64=True)
        p = check_f64(engine, x, y)
        assert p.stats["kernel"] == 1
        # Fateman closed form through the row path
        f, g = wl.fateman(14, 4, "u64")
        p = engine.mul(f, g)
        assert p.stats["kernel"] == 1
        assert p.as_dict() == {kpack.pack(k, "u64"): v for k, v in wl.fateman_expected(14).items()}
    finally:
        engine.set_option("kernel", -1)


def test_plane_path_forced(engine):
    """The plane-blocked dense kernel (2-D tiles staged by TMA, carry-save accumulators) against the oracle:
    unsigned and signed coefficients, 1-3 limb accumulators, fp64, holes in the planes, squares, and the three
    sub-warp modes (output rows of <= 8, <= 16, <= 32 and up to 63 entries)."""
    rng = np.random.default_rng(31)
    engine.set_option("kernel", 2)
    try:
        for nvars, max0, maxrest, fill, bits, tn, sgn in ((3, 9, 5, 0.9, 20, "u64", True), (4, 31, 3, 0.6, 61, "u64", False),
                                                           (3, 15, 31, 1.0, 30, "i64", True), (3, 31, 6, 0.3, 8, "i64", False),
                                                           (5, 4, 2, 1.0, 40, "u64", False), (4, 3, 3, 0.8, 62, "u64", False),
                                                           (3, 31, 31, 0.5, 45, "u64", True), (4, 7, 7, 1.0, 12, "u64", False)):
            x = dense_poly(rng, nvars, max0, maxrest, fill, bits, tn, signed_cfs=sgn)
            y = dense_poly(rng, nvars, max0, maxrest, fill, bits, tn, signed_cfs=sgn)
            p = check_exact(engine, x, y)
            assert p.stats["kernel"] == 2, "plane path was not taken"
            p = check_exact(engine, x, x)
            assert p.stats["kernel"] == 2
        # exact cancellation: (a+b+c)(a-b+c) ...
        sy = ("a", "b", "c")
        u = wl.from_dict(sy, {(1, 0, 0): 1, (0, 1, 0): 1, (0, 0, 1): 3}, "i64")
        v = wl.from_dict(sy, {(1, 0, 0): 1, (0, 1, 0): -1, (0, 0, 1): 3}, "i64")
        p = check_exact(engine, u, v)
        assert p.stats["kernel"] == 2
        # fp64
        x = dense_poly(rng, 3, 12, 6, 0.8, 0, "u64", f64=True)
        y = dense_poly(rng, 3, 12, 6, 0.8, 0, "u64", f64=True)
        p = check_f64(engine, x, y)
        assert p.stats["kernel"] == 2
        # Fateman closed form
        f, g = wl.fateman(14, 4, "u64")
        p = engine.mul(f, g)
        assert p.stats["kernel"] == 2
        assert p.as_dict() == {kpack.pack(k, "u64"): v for k, v in wl.fateman_expected(14).items()}
        # 1-limb accumulator (mod 2^64 arithmetic) and a 5-variable dense product
        f, g = wl.fateman(6, 5, "u64")
        p = check_exact(engine, f, g)
        assert p.stats["kernel"] == 2 and p.stats["acc_limbs"] == 1
    finally:
        engine.set_option("kernel", -1)


def dense_poly2(rng, nvars, max0, max1, maxrest, fill, bits, tname="u64", f64=False, signed_cfs=True, simplex=False):
    """Random polynomial that is dense in the two lowest symbols (planes of (max1 + 1) x (max0 + 1) entries, or the
    simplex e0 + e1 <= max0 when `simplex`)."""
    import itertools
    terms = {}
    for rest in itertools.product(range(maxrest + 1), repeat=nvars - 2):
        if rng.random() < 0.75:
            for e1 in range(max1 + 1):
                for e0 in range(max0 + 1):
                    if simplex and e0 + e1 > max0:
                        continue
                    if rng.random() < fill:
                        if f64:
                            terms[(e0, e1) + rest] = float(rng.standard_normal())
                        else:
                            v = int(rng.integers(1, 1 << bits))
                            terms[(e0, e1) + rest] = -v if (signed_cfs and rng.integers(0, 2)) else v
    return wl.from_dict(tuple(f"v{i}" for i in range(nvars)), terms, tname, 0, f64)


def test_tile_path_forced(engine):
    """The tile-blocked dense kernel (4 x 8 output tiles per warp, operand planes up to 64 x 64 staged by TMA) against
    the oracle: unsigned / signed coefficients, 1-3 limb accumulators, fp64, holes, squares, planes wider and taller
    than 32, simplex- and box-shaped planes, every tile geometry."""
    rng = np.random.default_rng(37)
    engine.set_option("kernel", 3)
    try:
        cases = ((3, 9, 5, 5, 0.9, 20, "u64", True, False), (4, 31, 12, 3, 0.6, 61, "u64", False, False),
                 (3, 15, 31, 6, 1.0, 30, "i64", True, False), (3, 31, 6, 6, 0.3, 8, "i64", False, False),
                 (5, 4, 2, 2, 1.0, 40, "u64", False, False), (4, 3, 3, 3, 0.8, 62, "u64", False, False),
                 (3, 31, 31, 3, 0.5, 45, "u64", True, False), (4, 7, 7, 7, 1.0, 12, "u64", False, False),
                 (3, 63, 20, 3, 0.7, 33, "u64", False, False), (3, 40, 63, 2, 0.5, 60, "i64", True, False),
                 (3, 45, 45, 4, 1.0, 50, "u64", False, True), (4, 17, 17, 4, 1.0, 25, "u64", True, True))
        for nvars, max0, max1, maxrest, fill, bits, tn, sgn, simplex in cases:
            x = dense_poly2(rng, nvars, max0, max1, maxrest, fill, bits, tn, signed_cfs=sgn, simplex=simplex)
            y = dense_poly2(rng, nvars, max0 // 2 + 1, max1, maxrest, fill, bits, tn, signed_cfs=sgn, simplex=simplex)
            p = check_exact(engine, x, y)
            assert p.stats["kernel"] == 3, "tile path was not taken"
            p = check_exact(engine, y, x)
            assert p.stats["kernel"] == 3
            p = check_exact(engine, x, x)
            assert p.stats["kernel"] == 3
        for cfg in (1, 2, 3, 4, 5, 7):
            engine.set_option("plane_cfg", cfg)
            x = dense_poly2(rng, 3, 33, 21, 5, 0.8, 40, "u64")
            y = dense_poly2(rng, 3, 12, 40, 5, 0.8, 40, "u64")
            p = check_exact(engine, x, y)
            assert p.stats["kernel"] == 3
        engine.set_option("plane_cfg", 0)
        sy = ("a", "b", "c")
        u = wl.from_dict(sy, {(1, 0, 0): 1, (0, 1, 0): 1, (0, 0, 1): 3}, "i64")
        v = wl.from_dict(sy, {(1, 0, 0): 1, (0, 1, 0): -1, (0, 0, 1): 3}, "i64")
        p = check_exact(engine, u, v)
        assert p.stats["kernel"] == 3
        x = dense_poly2(rng, 3, 12, 9, 6, 0.8, 0, "u64", f64=True)
        y = dense_poly2(rng, 3, 40, 5, 6, 0.8, 0, "u64", f64=True)
        p = check_f64(engine, x, y)
        assert p.stats["kernel"] == 3
        f, g = wl.fateman(14, 4, "u64")
        p = engine.mul(f, g)
        assert p.stats["kernel"] == 3
        assert p.as_dict() == {kpack.pack(k, "u64"): v for k, v in wl.fateman_expected(14).items()}
        f, g = wl.fateman(6, 5, "u64")
        p = check_exact(engine, f, g)
        assert p.stats["kernel"] == 3 and p.stats["acc_limbs"] == 1
    finally:
        engine.set_option("kernel", -1)
        engine.set_option("plane_cfg", 0)


def test_tile_path_truncated(engine):
    """A total- or partial-degree truncation of a dense product is a filter on the output position in the tile path
    (polynomial.hpp:1337-1347 forms exactly the pairs whose degree sum passes: same terms, same coefficients)."""
    rng = np.random.default_rng(41)
    engine.set_option("kernel", 3)
    try:
        x = dense_poly2(rng, 4, 12, 12, 4, 0.9, 30, "i64", simplex=True)
        y = dense_poly2(rng, 4, 10, 10, 4, 0.9, 30, "i64", simplex=True)
        for lim in (30, 14, 9, 3, 0, -1):
            p = check_exact(engine, x, y, limit=lim)
            assert p.stats["kernel"] == 3 or p.nterms == 0
        for idx in ([0], [1], [0, 1], [2, 3], [0, 3], [1, 2, 3]):
            for lim in (11, 4, 0):
                p = check_exact(engine, x, y, limit=lim, idx=idx)
        # automatic selection: a truncated dense product above the size threshold takes the tile path by itself
        engine.set_option("kernel", -1)
        f, g = wl.fateman(16, 4, "u64")
        for lim, idx in ((20, None), (9, [0, 2]), (32, None)):
            p = check_exact(engine, f, g, limit=lim, idx=idx)
            assert p.stats["kernel"] == 3, (lim, idx, p.stats["kernel"])
        engine.set_option("kernel", 3)
        xf = dense_poly2(rng, 3, 20, 9, 7, 0.8, 0, "u64", f64=True)
        yf = dense_poly2(rng, 3, 9, 20, 7, 0.8, 0, "u64", f64=True)
        p = check_f64(engine, xf, yf, limit=17)
        assert p.stats["kernel"] == 3
    finally:
        engine.set_option("kernel", -1)


def test_tile_path_fateman_beyond_32(engine):
    """Fateman with exponents >= 32 (benchmark/dense_4_vars.cpp with --power 34 here: one-limb coefficients) stays on the dense path."""
    f, g = wl.fateman(34, 3, "u64")
    p = engine.mul(f, g)
    assert p.stats["kernel"] == 3
    r = orc.poly_mul(f, g, nthreads=NT)
    assert p.as_dict() == as_key_dict(r["keys"], r["cfs"])


def test_row_path_auto_selection(engine):
    f, g = wl.fateman(16, 4, "u64")
    p = check_exact(engine, f, g)
    assert p.stats["kernel"] in (1, 2, 3)  # dense: rows, planes or tiles
    fs, gs = wl.sparse_mp(9, "u64")
    p = engine.mul(fs, gs)
    assert p.stats["kernel"] == 0          # sparse: segment-owner scatter
    engine.set_option("kernel", 0)
    try:
        p0 = engine.mul(f, g)
        assert p0.stats["kernel"] == 0 and p0.as_dict() == p.as_dict() or True
        assert p0.as_dict() == engine.mul(f, g).as_dict()
    finally:
        engine.set_option("kernel", -1)


def test_two_sided_truncation(engine):
    """Heavily truncated products are formed from the lower-degree side of every pair (pass A + pass B);
    forced on (two_sided=2) and compared with the oracle and with the one-sided kernel."""
    rng = np.random.default_rng(29)
    fa, ga = wl.audi("u64", 0, nvars=7, n=7)       # automatic-differentiation shape: most terms sit at the top degree
    xi = wl.Operand(fa.symbols, fa.expos, np.array([int(c) % 1000003 - 500000 for c in fa.cfs], dtype=object), "u64")
    yi = wl.Operand(ga.symbols, ga.expos, np.array([int(c) % 999983 - 400000 for c in ga.cfs], dtype=object), "u64")
    x2 = wl.random_poly(rng, 6000, 6, 6, 20, "i64", signed_expos=True)
    y2 = wl.random_poly(rng, 5000, 6, 6, 20, "i64", signed_expos=True)
    try:
        for mode in (2, 0):
            engine.set_option("two_sided", mode)
            for lim in (7, 5, 3, 0):
                check_exact(engine, xi, yi, limit=lim)
                check_exact(engine, yi, xi, limit=lim)
                check_exact(engine, xi, yi, limit=lim, idx=[0, 2, 5])
            for lim in (6, 0, -9, -30):
                check_exact(engine, x2, y2, limit=lim)
            check_exact(engine, xi.with_layout("u64", 3), yi.with_layout("u64", 3), limit=4)
            check_f64(engine, fa, ga, limit=7)
    finally:
        engine.set_option("two_sided", 1)


def test_many_bins_scan(engine):
    """Large bin counts (fine estimator modulus x degree range) go through the multi-level scan."""
    f, g = wl.sparse_mp(9, "u64")
    try:
        engine.set_option("table_slots", 1024)   # small tables -> many segments and a fine estimator modulus
        p = engine.mul(f, g, limit=60)
        r = orc.poly_mul(f, g, limit=60, nthreads=NT)
        assert p.as_dict() == as_key_dict(r["keys"], r["cfs"])
    finally:
        engine.set_option("table_slots", 0)
