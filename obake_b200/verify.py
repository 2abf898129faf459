"""Oracle-free checks of a product at benchmark size (numpy only; nothing here touches oracle/).

bench.py runs them on the result of the launch it has just timed and the -m gpu tests at the sizes the CPU oracle
cannot finish in seconds.  Three kinds:

* closed forms -- Fateman f*(f+1) = (1+..)^(2n) + (1+..)^n coefficient by coefficient (SURVEY 8c,
  benchmark/dense.hpp:21-44); audi_01's truncated product (benchmark/audi_01.cpp:44-68);
* the evaluation homomorphism: for an untruncated product over Z,  sum_t c_t * m_t(v) == f(v) * g(v)  (mod p)  at a
  point v and for several primes p < 2^31.  One wrong, missing, duplicated or spurious term changes the left side
  except with probability ~ deg/p per prime (Schwartz-Zippel), so three primes make a 2^-80-ish checksum of the whole
  term set that costs O(terms) and also works on a rank's SHARE of the result (the shares' sums add up);
* counts the reference's own tests and benchmarks pin.
"""
from __future__ import annotations

import math

import numpy as np

from . import kpack

PRIMES = (2147483647, 2147483629, 2147483587)  # the three largest primes below 2^31
KNOWN_TERMS = {"F20": 135751, "F30": 635376, "D5": 237336, "S10": 2096600, "S12": 5821335, "S16T": 28398035,
               "RECT": 1284816}


def unpack_keys(keys: np.ndarray, nvars: int, tname: str = "u64", psize: int = 0) -> np.ndarray:
    """(n, W) uint64 key words -> (n, nvars) int64 exponents (kunpacker semantics, vectorised)."""
    keys = np.asarray(keys, dtype=np.uint64).reshape(len(keys), -1)
    n = keys.shape[0]
    out = np.zeros((n, nvars), dtype=np.int64)
    if nvars == 0:
        return out
    wsize = psize if psize else nvars
    d = np.uint64(kpack.delta(tname, wsize))
    kmin = kpack.klims(tname, wsize)[0]
    lmin = kpack.lims(tname, wsize)[0]
    vi = 0
    for w in range(keys.shape[1]):
        v = keys[:, w] - np.uint64(kmin & ((1 << 64) - 1))  # wraps like the reference's unsigned subtraction
        for _ in range(wsize):
            if vi >= nvars:
                break
            q = v // d
            out[:, vi] = (v - q * d).astype(np.int64) + lmin
            v = q
            vi += 1
    return out


def limbs_mod(limbs: np.ndarray, p: int) -> np.ndarray:
    """Two's-complement little-endian limb rows (n, L) uint64 -> value mod p (uint64), vectorised."""
    limbs = np.asarray(limbs, dtype=np.uint64)
    n, L = limbs.shape
    P = np.uint64(p)
    two32 = np.uint64((1 << 32) % p)
    acc = np.zeros(n, dtype=np.uint64)
    for l in range(L - 1, -1, -1):
        hi = (limbs[:, l] >> np.uint64(32)) % P
        lo = (limbs[:, l] & np.uint64(0xFFFFFFFF)) % P
        acc = (acc * two32 + hi) % P
        acc = (acc * two32 + lo) % P
    neg = (limbs[:, L - 1] >> np.uint64(63)).astype(bool)
    corr = np.uint64(pow(2, 64 * L, p))
    acc = np.where(neg, (acc + P - corr) % P, acc)
    return acc


def _point(nvars: int) -> list[int]:
    # fixed evaluation point: distinct small primes (deterministic, the same on every rank)
    pr = [3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71, 73, 79, 83, 89, 97, 101, 103, 107,
          109, 113, 127, 131, 137]
    return [pr[i % len(pr)] + 1000 * (i // len(pr)) for i in range(nvars)]


def eval_mod(expos: np.ndarray, cf_mod: np.ndarray, p: int) -> int:
    """sum_t cf_t * prod_j v_j^e_tj  mod p for non-negative exponents (vectorised with power tables)."""
    n, nv = expos.shape
    if n == 0:
        return 0
    if expos.min() < 0:
        raise ValueError("evaluation checksum needs non-negative exponents")
    P = np.uint64(p)
    acc = np.asarray(cf_mod, dtype=np.uint64) % P
    pt = _point(nv)
    for j in range(nv):
        emax = int(expos[:, j].max())
        tab = np.ones(emax + 1, dtype=np.uint64)
        for e in range(1, emax + 1):
            tab[e] = (int(tab[e - 1]) * pt[j]) % p
        acc = (acc * tab[expos[:, j]]) % P
    # sum in chunks: each addend < 2^31
    return int(np.add.reduce(acc.reshape(-1), dtype=np.uint64) % P) if n < (1 << 32) else int(sum(int(x) for x in acc) % p)


def operand_eval(op, p: int) -> int:
    """f(v) mod p for a workloads.Operand with exact integer coefficients."""
    cf = np.array([int(c) % p for c in op.cfs], dtype=np.uint64)
    return eval_mod(np.asarray(op.expos, dtype=np.int64), cf, p)


def product_eval_sums(keys, limbs, nvars, tname="u64", psize=0) -> list[int]:
    """This share's sum_t c_t m_t(v) mod p for every prime of PRIMES."""
    ex = unpack_keys(keys, nvars, tname, psize)
    return [eval_mod(ex, limbs_mod(limbs, p), p) for p in PRIMES]


def expected_eval(f, g) -> list[int]:
    return [(operand_eval(f, p) * operand_eval(g, p)) % p for p in PRIMES]


# ---- closed forms -------------------------------------------------------------------------------------------------
def fateman_closed_form(expos: np.ndarray, n: int) -> list[int]:
    """Exact coefficient of every monomial `expos[i]` in (1+sum x)^(2n) + (1+sum x)^n."""
    fact = [math.factorial(i) for i in range(2 * n + 1)]
    out = []
    for row in expos.tolist():
        s = sum(row)
        den = 1
        for e in row:
            den *= fact[e]
        c = fact[2 * n] // (den * fact[2 * n - s]) if 0 <= s <= 2 * n else 0
        if s <= n:
            c += fact[n] // (den * fact[n - s])
        out.append(c)
    return out


def ints_from_limbs(limbs: np.ndarray) -> list[int]:
    limbs = np.asarray(limbs, dtype=np.uint64)
    n, L = limbs.shape
    cols = [limbs[:, l].tolist() for l in range(L)]
    top = 1 << (64 * L)
    half = top >> 1
    out = []
    for i in range(n):
        v = 0
        for l in range(L - 1, -1, -1):
            v = (v << 64) | cols[l][i]
        out.append(v - top if v >= half else v)
    return out


def check_fateman_share(keys, limbs, n: int, nvars: int = 4, tname: str = "u64") -> dict:
    """Every term of this share of f*(f+1) against the closed form; also that no key repeats inside the share."""
    keys = np.asarray(keys, dtype=np.uint64).reshape(len(keys), -1)
    ex = unpack_keys(keys, nvars, tname, 0)
    want = fateman_closed_form(ex, n)
    got = ints_from_limbs(limbs)
    bad = sum(1 for a, b in zip(got, want) if a != b or b == 0)
    dup = int(len(keys) - len(np.unique(keys[:, 0])))
    inside = bool(((ex >= 0).all() and (ex.sum(axis=1) <= 2 * n).all())) if len(keys) else True
    return {"terms": int(len(keys)), "mismatches": int(bad), "duplicates": dup, "ok": bad == 0 and dup == 0 and inside}


def audi_closed_form(expos: np.ndarray, nvars: int = 10, n: int = 10) -> np.ndarray:
    """Coefficient of every monomial of degree d <= n in (nvars+1 + S)^n * (-(nvars-1) - S)^n truncated at degree n,
    S = sum x_i:  d!/prod(e_i!) * sum_{a+b=d} C(n,a)(nvars+1)^(n-a) * C(n,b)(-1)^n (nvars-1)^(n-b)  (exact, then float)."""
    fact = [math.factorial(i) for i in range(2 * n + 1)]
    sgn = (-1) ** n
    by_deg = []
    for d in range(n + 1):
        s = 0
        for a in range(d + 1):
            b = d - a
            s += math.comb(n, a) * (nvars + 1) ** (n - a) * math.comb(n, b) * sgn * (nvars - 1) ** (n - b)
        by_deg.append(s)
    out = np.empty(len(expos), dtype=np.float64)
    for i, row in enumerate(expos.tolist()):
        d = sum(row)
        den = 1
        for e in row:
            den *= fact[e]
        out[i] = float(by_deg[d] * (fact[d] // den)) if d <= n else float("nan")
    return out


def check_product(name: str, f, g, limit, keys, cfs, is_f64: bool, world_sum=None) -> dict:
    """The strongest oracle-free verdict available for the named workload.  `keys`/`cfs` are this process's share
    ((n, W) uint64 and (n, L) uint64 limbs or (n,) float64); `world_sum(list[int]) -> list[int]` adds per-rank
    integers over the ranks (None = single process).  Returns {"checked", "ok", "terms", "how", ...}."""
    name = name.upper()
    ws = world_sum or (lambda v: v)
    n_local = int(len(keys))
    res = {"checked": True, "terms_local": n_local}
    oks, how = [], []
    if is_f64:
        if name in ("AUDI", "AUDID"):
            ex = unpack_keys(keys, f.nvars, f.tname, f.psize)
            want = audi_closed_form(ex, f.nvars, 10)
            got = np.asarray(cfs, dtype=np.float64)
            rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-300)
            bad = int((~(rel <= 1e-12)).sum())
            uniq = len(np.unique(np.asarray(keys, dtype=np.uint64).reshape(n_local, -1), axis=0))
            tot, badw, dupw = ws([n_local, bad, n_local - uniq])
            oks.append(badw == 0 and dupw == 0 and tot == math.comb(f.nvars + 10, 10))
            res.update(terms=int(tot), mismatches=int(badw), max_rel_err=float(rel.max()) if n_local else 0.0)
            how.append("closed form of audi_01's truncated product, relative 1e-12 per coefficient, all C(20,10) terms present")
        else:
            tot, = ws([n_local])
            res.update(terms=int(tot))
            res["checked"] = False
            how.append("no oracle-free check for this fp64 workload")
    else:
        limbs = np.asarray(cfs, dtype=np.uint64).reshape(n_local, -1)
        nz = int(((limbs != 0).any(axis=1)).sum()) if n_local else 0
        untruncated = limit is None or name == "S16T"  # sparse_02_truncated's limit (300) exceeds every degree sum (160)
        sums = product_eval_sums(keys, limbs, f.nvars, f.tname, f.psize) if untruncated else [0, 0, 0]
        uniq = len(np.unique(np.asarray(keys, dtype=np.uint64).reshape(n_local, -1), axis=0)) if n_local else 0
        agg = ws([n_local, n_local - nz, n_local - uniq] + sums)
        tot, zeros, dups = agg[0], agg[1], agg[2]
        res.update(terms=int(tot), zero_coefficients=int(zeros), duplicates=int(dups))
        oks.append(zeros == 0 and dups == 0)
        if untruncated:
            want = expected_eval(f, g)
            got = [int(s) % p for s, p in zip(agg[3:6], PRIMES)]
            oks.append(got == want)
            res["eval_checksum"] = {"primes": list(PRIMES), "got": got, "want": want}
            how.append("evaluation homomorphism sum c_t m_t(v) == f(v) g(v) mod three 31-bit primes")
        if name in KNOWN_TERMS:
            oks.append(tot == KNOWN_TERMS[name])
            res["terms_expected"] = KNOWN_TERMS[name]
            how.append("term count pinned by the reference / closed form")
        if name.startswith("F") and name[1:].isdigit() or name == "D5":
            nn, nv = (14, 5) if name == "D5" else (int(name[1:]), 4)
            r = check_fateman_share(keys, limbs, nn, nv, f.tname)
            badw, = ws([r["mismatches"]])
            oks.append(badw == 0)
            res["closed_form_mismatches"] = int(badw)
            how.append("every coefficient against the closed form (1+..)^(2n) + (1+..)^n")
    res["ok"] = bool(all(oks)) if res["checked"] else None
    res["how"] = "; ".join(how)
    return res
