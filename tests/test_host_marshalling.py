"""Host-side marshalling of the Python layer (no GPU): exact integers <-> two's-complement limbs as the C ABI carries
them (include/obk_b200.h: `cf_limbs` 64-bit limbs per term, little endian), operand views, truncation arguments."""
import ctypes as C
import random

import numpy as np
import pytest

from obake_b200 import _capi, engine, workloads as wl


def test_limb_round_trip_and_width():
    rnd = random.Random(7)
    for L in (1, 2, 3):
        lo, hi = -(1 << (64 * L - 1)), (1 << (64 * L - 1)) - 1
        vals = [lo, hi, 0, -1, 1, lo + 1, hi - 1] + [rnd.randint(lo, hi) for _ in range(200)]
        arr = engine.int_to_limbs(vals, L)
        assert arr.shape == (len(vals), L) and arr.dtype == np.uint64
        assert engine.limbs_to_int(arr) == vals
    # limbs needed: magnitude bits + the sign bit
    assert engine.needed_limbs([0]) == 1
    assert engine.needed_limbs([(1 << 63) - 1]) == 1 and engine.needed_limbs([1 << 63]) == 2
    assert engine.needed_limbs([-(1 << 63)]) == 1 and engine.needed_limbs([-(1 << 63) - 1]) == 2
    assert engine.needed_limbs([1 << 127]) == 3
    # sign extension: -1 is all ones in every limb
    assert engine.int_to_limbs([-1], 3).tolist() == [[(1 << 64) - 1] * 3]


@pytest.mark.parametrize("tn,psize", [("u64", 0), ("i64", 0), ("u64", 4)])
def test_host_operand_view(tn, psize):
    rng = np.random.default_rng(1)
    x = wl.random_poly(rng, 50, 6, 5, 70, tn, psize, signed_expos=(tn == "i64"))
    h = engine.HostOperand(x)
    v = h.view
    assert v.nterms == 50 and v.nvars == 6 and v.psize == psize
    assert v.key_words == (1 if psize == 0 else 2) and v.key_signed == (1 if tn == "i64" else 0)
    assert v.cf_kind == _capi.OBK_CF_INT and v.cf_limbs == 2 and v.mem == _capi.OBK_MEM_HOST
    assert v.keys == h.keys.ctypes.data and v.cfs == h.cfs.ctypes.data
    assert engine.limbs_to_int(h.cfs) == [int(c) for c in x.cfs]
    f = engine.HostOperand(x.as_f64()).view
    assert f.cf_kind == _capi.OBK_CF_F64 and f.cf_limbs == 1


def test_trunc_argument():
    keep = []
    assert engine.Engine._trunc(None, None, keep) is None
    t = engine.Engine._trunc(-3, None, keep)
    assert (t.limit, t.nidx, t.partial) == (-3, 0, 0)
    t = engine.Engine._trunc(7, [4, 1, 1 + 1], keep)
    assert (t.limit, t.nidx, t.partial) == (7, 3, 1)
    idx = np.ctypeslib.as_array(C.cast(t.idx, C.POINTER(C.c_uint32)), (3,))
    assert idx.tolist() == [1, 2, 4]  # sorted positions
    t = engine.Engine._trunc(7, [], keep)
    assert (t.nidx, t.partial) == (0, 1)  # partial degree over no symbol: every degree is 0
