"""Kronecker packing: tables vs the reference's numbers, round trips and limit errors
(restates test/kpack.cpp:41-237 of the reference)."""
import json
import os
import random

import numpy as np
import pytest

from obake_b200 import kpack

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "kpack_tables.json")))


@pytest.mark.parametrize("tn", ["i32", "u32", "i64", "u64"])
def test_tables_equal_reference(tn):
    t = kpack.tables(tn)
    g = GOLD[tn]
    assert t["max_size"] == len(g["deltas"])
    assert [str(v) for v in t["deltas"]] == g["deltas"]
    assert [str(v) for v in t["lims"]] == g["lims"]
    assert [str(v) for v in t["klims"]] == g["klims"]
    for s in range(t["max_size"]):
        for j in range(t["max_size"] + 1):
            a, b = t["divcnst"][s][j], g["divcnst"][s][j]
            assert (str(a[0]), a[1], a[2]) == (b[0], b[1], b[2]), (tn, s, j)


@pytest.mark.parametrize("tn", ["i32", "u32", "i64", "u64"])
def test_roundtrip_random(tn):
    # test/kpack.cpp:101-134 (10000 trials there; a few hundred per size here keeps the CPU suite short)
    rng = random.Random(42)
    t = kpack.tables(tn)
    for size in range(1, t["max_size"] + 1):
        lo, hi = kpack.lims(tn, size)
        for _ in range(200):
            e = [rng.randint(lo, hi) for _ in range(size)]
            v = kpack.pack(e, tn)
            assert kpack.unpack(v, size, tn) == e
            assert kpack.unpack_divcnst(v, size, tn) == e


@pytest.mark.parametrize("tn", ["i32", "u32", "i64", "u64"])
def test_min_max_and_errors(tn):
    # test/kpack.cpp:146-219
    t = kpack.tables(tn)
    for size in range(1, t["max_size"] + 1):
        lo, hi = kpack.lims(tn, size)
        kmin, kmax = kpack.klims(tn, size)
        assert kpack.pack([lo] * size, tn) == kmin
        assert kpack.pack([hi] * size, tn) == kmax
        assert kpack.unpack(kmin, size, tn) == [lo] * size
        assert kpack.unpack(kmax, size, tn) == [hi] * size
        with pytest.raises(OverflowError):
            kpack.pack([hi + 1] + [0] * (size - 1), tn)
        if size > 1 or t["signed"]:
            with pytest.raises(OverflowError):
                kpack.pack([lo - 1] + [0] * (size - 1), tn)
        with pytest.raises(OverflowError):
            kpack.unpack(kmax + 1, size, tn)
    with pytest.raises(OverflowError):
        kpack.pack([0] * (t["max_size"] + 1), tn)
    with pytest.raises(ValueError):
        kpack.unpack(1, 0, tn)
    assert kpack.unpack(0, 0, tn) == []


def test_homomorphic_hash():
    # h(v1) + h(v2) == h(v1 + v2): test/polynomials_packed_monomial_00.cpp:574-618
    rng = random.Random(7)
    for tn in ("i64", "u64"):
        for size in (1, 2, 4, 5, 10):
            lo, hi = kpack.lims(tn, size)
            lo, hi = lo // 2, hi // 2
            for _ in range(100):
                a = [rng.randint(lo, hi) for _ in range(size)]
                b = [rng.randint(lo, hi) for _ in range(size)]
                s = [x + y for x, y in zip(a, b)]
                m = (1 << 64) - 1
                assert (kpack.pack(a, tn) + kpack.pack(b, tn)) & m == kpack.pack(s, tn) & m


def test_monomial_mul_known_answer():
    ka = json.load(open(os.path.join(HERE, "golden", "known_answers.json")))["monomial_mul"]
    for tn in ("i64", "u64", "i32", "u32"):
        assert kpack.pack(ka["a"], tn) + kpack.pack(ka["b"], tn) == kpack.pack(ka["c"], tn)


def test_pack_many_matches_scalar():
    rng = np.random.default_rng(0)
    e = rng.integers(0, 30, size=(50, 5))
    k = kpack.pack_many(e, "u64")
    for row, v in zip(e, k):
        assert int(v) == kpack.pack(row.tolist(), "u64")
    kd = kpack.pack_many_d(e, 2, "u64")
    assert kd.shape == (50, 3)
    for row, ws in zip(e, kd):
        assert int(ws[0]) == kpack.pack(row[:2].tolist(), "u64", 2)
        assert int(ws[2]) == kpack.pack(row[4:].tolist(), "u64", 2)
