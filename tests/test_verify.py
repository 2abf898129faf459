"""The oracle-free checks bench.py and the bench-size GPU tests rely on (obake_b200/verify.py) are themselves checked
here against the CPU oracle: they must accept the oracle's products and reject a corrupted coefficient, a dropped term
and a duplicated term."""
import numpy as np
import pytest

import orc
from obake_b200 import verify as V, workloads as wl


def _limbs(cfs, n=3):
    return orc.int_to_limbs(cfs, n)


@pytest.mark.parametrize("name", ["F8", "S5"])
def test_accepts_oracle_and_rejects_corruption(name):
    f, g, limit, _ = wl.named(name)
    r = orc.poly_mul(f, g, nthreads=4)
    keys, limbs = r["keys"], _limbs(r["cfs"])
    res = V.check_product(name, f, g, None, keys, limbs, False)
    assert res["ok"] and res["terms"] == len(keys)
    bad = limbs.copy()
    bad[len(bad) // 2, 0] += np.uint64(1)
    assert not V.check_product(name, f, g, None, keys, bad, False)["ok"]
    assert not V.check_product(name, f, g, None, keys[1:], limbs[1:], False)["ok"]
    dk, dl = np.concatenate([keys, keys[:1]]), np.concatenate([limbs, limbs[:1]])
    assert not V.check_product(name, f, g, None, dk, dl, False)["ok"]


def test_shares_add_up():
    # the multi-GPU form: every rank checks its share, the checksums are summed over the ranks
    f, g, _, _ = wl.named("F6")
    r = orc.poly_mul(f, g, nthreads=2)
    keys, limbs = r["keys"], _limbs(r["cfs"])
    half = len(keys) // 2
    shares = [(keys[:half], limbs[:half]), (keys[half:], limbs[half:])]
    captured = []

    def run(share, total):
        return V.check_product("F6", f, g, None, share[0], share[1], False, world_sum=total)

    # first pass collects every share's local vector, second pass feeds the sums back
    for sh in shares:
        run(sh, lambda v: (captured.append(list(v)), v)[1])
    per_call = len(captured) // 2
    sums = [[a + b for a, b in zip(captured[i], captured[per_call + i])] for i in range(per_call)]
    for sh in shares:
        it = iter(sums)
        assert run(sh, lambda v: next(it))["ok"]


def test_signed_coefficients_and_signed_keys():
    rng = np.random.default_rng(5)
    x = wl.random_poly(rng, 300, 4, 9, 50, "u64")
    y = wl.random_poly(rng, 200, 4, 9, 60, "u64")
    r = orc.poly_mul(x, y, nthreads=2)
    assert V.check_product("R", x, y, None, r["keys"], _limbs(r["cfs"]), False)["ok"]
    xi = wl.random_poly(rng, 100, 3, 5, 20, "i64")
    ex = V.unpack_keys(xi.keys, 3, "i64")
    assert (ex == xi.expos).all()


def test_limbs_mod_matches_python_ints():
    vals = [0, 1, -1, 2**64 - 1, -(2**64), 2**127 - 5, -(2**130) + 12345, 3**100 % 2**190]
    limbs = orc.int_to_limbs(np.array(vals, dtype=object), 3)
    for p in V.PRIMES:
        assert V.limbs_mod(limbs, p).tolist() == [v % p for v in vals]


def test_audi_closed_form_matches_oracle_small():
    f, g = wl.audi("u64", 0, nvars=4, n=6)
    r = orc.poly_mul(f, g, limit=6, nthreads=2)
    ex = V.unpack_keys(r["keys"], 4, "u64")
    want = V.audi_closed_form(ex, 4, 6)
    got = np.asarray(r["cfs"], dtype=np.float64)
    assert np.all(np.abs(got - want) <= 1e-12 * np.abs(want))
