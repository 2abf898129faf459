"""Pins the CPU oracle (oracle/obk_oracle.cpp) against the reference's known answers, the closed form of
the Fateman product and the independent Python restatement (oracle/pyoracle.py)."""
import json
import os

import numpy as np
import pytest

import orc
from obake_b200 import kpack, workloads as wl
from oracle import pyoracle

HERE = os.path.dirname(os.path.abspath(__file__))
KA = json.load(open(os.path.join(HERE, "golden", "known_answers.json")))


def test_oracle_kpack_tables_equal_reference():
    import ctypes as C
    gold = json.load(open(os.path.join(HERE, "golden", "kpack_tables.json")))
    L = orc.lib()
    for signed, tn in ((0, "u64"), (1, "i64")):
        d = (C.c_uint64 * 24)()
        l = (C.c_uint64 * 24)()
        k = (C.c_uint64 * 24)()
        mp = (C.c_uint64 * (24 * 25))()
        s1 = (C.c_uint32 * (24 * 25))()
        s2 = (C.c_uint32 * (24 * 25))()
        n = L.orc_kpack_tables(signed, d, l, k, mp, s1, s2)
        g = gold[tn]
        assert n == len(g["deltas"]) == 21
        assert [str(v) for v in d[:n]] == g["deltas"]
        assert [str(v) for v in l[:n]] == g["lims"]
        assert [str(v) for v in k[:n]] == g["klims"]
        for s in range(n):
            for j in range(n + 1):
                b = g["divcnst"][s][j]
                assert (str(mp[s * (n + 1) + j]), s1[s * (n + 1) + j], s2[s * (n + 1) + j]) == (b[0], b[1], b[2])


def test_oracle_unpack_matches_python():
    rng = np.random.default_rng(1)
    for tn, signed in (("u64", False), ("i64", True)):
        for nv in (1, 3, 5, 10, 21):
            lo, hi = kpack.lims(tn, nv)
            lo, hi = max(lo, -1000), min(hi, 1000)
            e = rng.integers(lo, hi + 1, size=(40, nv))
            op = wl.Operand(tuple(f"v{i:02d}" for i in range(nv)), e, np.ones(40, dtype=object), tn, 0)
            assert (orc.unpack_keys(op) == e).all()
            assert (orc.degrees(op) == e.sum(axis=1)).all()
            if nv >= 3:
                assert (orc.degrees(op, [0, 2]) == e[:, [0, 2]].sum(axis=1)).all()
        # d_packed layout
        e = rng.integers(0, 50, size=(40, 10))
        op = wl.Operand(tuple(f"v{i:02d}" for i in range(10)), e, np.ones(40, dtype=object), tn, 4)
        assert op.key_words == 3
        assert (orc.unpack_keys(op) == e).all()


def test_small_identities():
    # test/polynomials_polynomial_00.cpp:198-214 via direct mt_hm calls
    sy = ("a", "b")
    a = wl.from_dict(sy, {(1, 0): 1}, "i64")
    b = wl.from_dict(sy, {(0, 1): 1}, "i64")
    three = wl.from_dict(sy, {(0, 0): 3}, "i64")
    four = wl.from_dict(sy, {(0, 0): 4}, "i64")
    r = orc.result_to_dict(orc.poly_mul(three, four), three)
    assert r == {(0, 0): 12}
    apb = wl.from_dict(sy, {(1, 0): 1, (0, 1): 1}, "i64")
    amb = wl.from_dict(sy, {(1, 0): 1, (0, 1): -1}, "i64")
    r = orc.result_to_dict(orc.poly_mul(apb, amb), apb)
    assert r == {(2, 0): 1, (0, 2): -1}  # the cross terms cancel and are erased
    a2b2 = wl.from_dict(sy, {(2, 0): 1, (0, 2): 1}, "i64")
    r2 = wl.from_dict(sy, r, "i64")
    r = orc.result_to_dict(orc.poly_mul(a2b2, r2), a2b2)
    assert r == {(4, 0): 1, (0, 4): -1}


@pytest.mark.parametrize("tn", ["i64", "u64"])
def test_vs_python_oracle_random(tn):
    rng = np.random.default_rng(3)
    for nv, nt1, nt2, mx in ((3, 40, 60, 6), (6, 100, 30, 3), (1, 20, 20, 100)):
        x = wl.random_poly(rng, nt1, nv, mx, 30, tn)
        y = wl.random_poly(rng, nt2, nv, mx, 30, tn)
        got = orc.result_to_dict(orc.poly_mul(x, y, nthreads=3), x)
        assert got == pyoracle.poly_mul(x.to_dict(), y.to_dict())


def test_fateman_closed_form():
    f, g = wl.fateman(8, 4, "u64")
    r = orc.poly_mul(f, g, nthreads=4)
    assert orc.result_to_dict(r, f) == wl.fateman_expected(8)
    assert r["n_mults"] == f.nterms * g.nterms


def test_sparse_n10_known_answer():
    # test/polynomials_polynomial_00.cpp:398-424: 2 096 600 terms for integer<1> and double
    f, g = wl.sparse_mp(10, "i64")
    assert f.nterms == 3003
    r = orc.poly_mul(f, g, nthreads=os.cpu_count() or 1)
    assert len(r["keys"]) == KA["sparse_n10_terms"]["value"]
    # series invariant: every term sits in table hash & (nsegs-1)  (series.hpp:396-400)
    so = r["seg_offsets"]
    seg = np.searchsorted(so, np.arange(len(r["keys"])), side="right") - 1
    assert ((r["keys"][:, 0] & np.uint64((1 << r["log2_nsegs"]) - 1)) == seg.astype(np.uint64)).all()
    rd = orc.poly_mul(f.as_f64(), g.as_f64(), nthreads=os.cpu_count() or 1)
    assert len(rd["keys"]) == KA["sparse_n10_terms"]["value"]


def test_rect_pow20_known_answer():
    # test/polynomials_polynomial_04.cpp:264-285: 19 rectangular products -> 53 130 terms
    base = wl.mp_base("i64")
    cur = base
    for _ in range(19):
        r = orc.poly_mul(base, cur, nthreads=4)
        d = orc.result_to_dict(r, base)
        cur = wl.from_dict(base.symbols, d, "i64")
    assert cur.nterms == KA["rect_pow20_terms"]["value"]


def test_truncated_n8_reference_cases():
    # test/polynomials_polynomial_01.cpp:185-236 and _04.cpp:239-262
    f, g = wl.sparse_mp(8, "i64")
    full = orc.result_to_dict(orc.poly_mul(f, g, nthreads=8), f)
    assert len(full) == 591235
    for lim in (1000, 80):
        assert orc.result_to_dict(orc.poly_mul(f, g, limit=lim, nthreads=8), f) == full
    for lim in (40, 5):
        t = orc.result_to_dict(orc.poly_mul(f, g, limit=lim, nthreads=8), f)
        assert max(sum(k) for k in t) == lim
        assert t == pyoracle.truncate_degree(full, lim)
    t = orc.result_to_dict(orc.poly_mul(f, g, limit=50, nthreads=8), f)
    assert t == pyoracle.truncate_degree(full, 50)
    t = orc.result_to_dict(orc.poly_mul(f, g, limit=0, nthreads=8), f)
    assert t == {(0, 0, 0, 0, 0): 1}
    r = orc.poly_mul(f, g, limit=-1, nthreads=8)
    assert len(r["keys"]) == 0 and r["n_mults"] == 0


def test_partial_degree_truncation():
    # test/polynomials_polynomial_02.cpp:182-327 (partial degree over a subset of the symbols)
    f, g = wl.sparse_mp(4, "i64")
    fd, gd = f.to_dict(), g.to_dict()
    for idx in ([0], [1, 3], [0, 1, 2, 3, 4]):
        for lim in (0, 3, 9, 100):
            got = orc.result_to_dict(orc.poly_mul(f, g, limit=lim, idx=idx, nthreads=2), f)
            assert got == pyoracle.poly_mul(fd, gd, lim, idx), (idx, lim)


def test_overflow_check():
    # test/polynomials_polynomial_00.cpp:216-238; packed_monomial_00.cpp:496-552
    lo, hi = kpack.lims("i64", 2)
    sy = ("a", "b")
    big = wl.from_dict(sy, {(hi, 0): 1, (0, 1): 1}, "i64")
    one = wl.from_dict(sy, {(1, 0): 1}, "i64")
    assert not orc.overflow_check(big, one)
    assert orc.poly_mul(big, one)["status"] == 2
    small = wl.from_dict(sy, {(lo, 0): 1}, "i64")
    neg = wl.from_dict(sy, {(-1, 0): 1}, "i64")
    assert not orc.overflow_check(small, neg)
    assert orc.overflow_check(one, one)
    ok = wl.from_dict(sy, {(hi - 1, 0): 1}, "i64")
    assert orc.overflow_check(ok, one)
    # d_packed with the same check per component
    _, hi8 = kpack.lims("u64", 8)
    syd = tuple(f"v{i}" for i in range(10))
    e = [0] * 10
    e[9] = hi8
    bigd = wl.from_dict(syd, {tuple(e): 1}, "u64", 8)
    e2 = [0] * 10
    e2[9] = 1
    oned = wl.from_dict(syd, {tuple(e2): 1}, "u64", 8)
    assert not orc.overflow_check(bigd, oned)
    assert orc.overflow_check(oned, oned)


def test_f64_against_python():
    rng = np.random.default_rng(5)
    x = wl.random_poly(rng, 50, 3, 5, f64=True)
    y = wl.random_poly(rng, 70, 3, 5, f64=True)
    got = orc.result_to_dict(orc.poly_mul(x, y, nthreads=2), x)
    ref = pyoracle.poly_mul(x.to_dict(), y.to_dict())
    assert got.keys() == ref.keys()
    for k in ref:
        assert abs(got[k] - ref[k]) <= 1e-12 * max(1.0, abs(ref[k]))


def test_wide_coefficients_three_limbs():
    # F30-like widths: 61-bit inputs -> 3-limb accumulators
    rng = np.random.default_rng(6)
    x = wl.random_poly(rng, 30, 2, 7, 61, "u64")
    y = wl.random_poly(rng, 30, 2, 7, 61, "u64")
    r = orc.poly_mul(x, y)
    assert orc.result_to_dict(r, x) == pyoracle.poly_mul(x.to_dict(), y.to_dict())
    # two-limb inputs
    x2 = wl.random_poly(rng, 20, 2, 7, 100, "u64")
    y2 = wl.random_poly(rng, 20, 2, 7, 70, "u64")
    r = orc.poly_mul(x2, y2)
    assert orc.result_to_dict(r, x2) == pyoracle.poly_mul(x2.to_dict(), y2.to_dict())


def test_series_ops_golden():
    """The Python oracle's series add/sub, degree truncation and symbol-set extension against answers written out
    by hand, and against the committed fixture (tests/golden/series_ops.json, tests/golden/make_golden.py)."""
    x = {(1, 0): 1, (0, 1): 1}   # x + y
    y = {(1, 0): 1, (0, 1): -1}  # x - y
    assert pyoracle.poly_addsub(x, y) == {(1, 0): 2}            # 2x: the y terms cancel and are erased
    assert pyoracle.poly_addsub(x, y, True) == {(0, 1): 2}      # 2y
    assert pyoracle.poly_addsub({}, y, True) == {(1, 0): -1, (0, 1): 1}
    b = {(0, 0): 1, (1, 0): 1, (0, 1): 1}
    p = pyoracle.poly_mul(pyoracle.poly_mul(b, b), b)           # (1 + x + y)^3
    assert pyoracle.truncate_degree(p, 2) == {(0, 0): 1, (1, 0): 3, (0, 1): 3, (2, 0): 3, (1, 1): 6, (0, 2): 3}
    assert pyoracle.truncate_degree(p, -1) == {}
    # partial degree over x only: keep x^0 and x^1 with every power of y
    assert pyoracle.truncate_degree(p, 1, [0]) == {(0, 0): 1, (0, 1): 3, (0, 2): 3, (0, 3): 1, (1, 0): 3, (1, 1): 6,
                                                   (1, 2): 3}
    assert pyoracle.extend_symbols(b, 4, [1, 3]) == {(0, 0, 0, 0): 1, (0, 1, 0, 0): 1, (0, 0, 0, 1): 1}
    with open(os.path.join(HERE, "golden", "series_ops.json")) as f:
        cases = json.load(f)

    def unpack(lst):
        return {tuple(k): v for k, v in lst}

    for c in cases:
        if c["op"] in ("add", "sub"):
            got = pyoracle.poly_addsub(unpack(c["x"]), unpack(c["y"]), c["op"] == "sub")
        elif c["op"] == "truncate":
            got = pyoracle.truncate_degree(unpack(c["x"]), c["limit"], c["idx"])
        else:
            got = pyoracle.extend_symbols(unpack(c["x"]), c["new_nvars"], c["pos"])
        assert got == unpack(c["out"]), c["op"]


def test_three_limb_inputs_against_python_bigints():
    """Inputs beyond 128 bits (three limbs, five-limb sums): the C++ oracle against exact Python integers."""
    rng = np.random.default_rng(77)
    for bx, by in ((150, 30), (185, 100), (129, 129), (64, 190)):
        x = wl.random_poly(rng, 120, 3, 6, bx, "i64")
        y = wl.random_poly(rng, 100, 3, 6, by, "i64")
        r = orc.poly_mul(x, y, nthreads=2)
        assert r["status"] == 0
        assert orc.result_to_dict(r, x) == pyoracle.poly_mul(x.to_dict(), y.to_dict())
