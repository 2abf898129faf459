"""Deterministic synthetic operands for the reference's series-product benchmarks and tests.

Every generator returns :class:`Operand` objects (exponent matrix + exact coefficients) built from
closed forms, so no multiplication is needed to *make* the inputs:

* ``fateman(n, nvars)``   benchmark/dense.hpp:21-69  f=(1+x+y+z+t[+u])^n, g=f+1
* ``sparse_mp(n)``        benchmark/sparse.hpp:21-47 f=(1+x+y+2z^2+3t^3+5u^5)^n, g=(1+u+t+2z^2+3y^3+5x^5)^n
* ``audi()``              benchmark/audi_01.cpp:44-68 f=(11+sum x_i)^10, g=(-9-sum x_i)^10, 10 vars, fp64
* ``rect_base()``         benchmark/rectangular_01.cpp:35-38, the 13-term f in 3 variables
* ``mp_base()``           the 6-term base of the sparse pair (test/polynomials_polynomial_04.cpp:264-285)

Variables are always ordered by their sorted names (symbols.hpp:36), which fixes the digit each
exponent occupies in the packed key.
"""
from __future__ import annotations

import itertools
import math
from dataclasses import dataclass, field

import numpy as np

from . import kpack


@dataclass
class Operand:
    """One polynomial as SoA host data, ready to be viewed by the C-ABI."""

    symbols: tuple            # sorted variable names
    expos: np.ndarray         # (n, nvars) int64 exponent matrix
    cfs: np.ndarray           # (n,) object (exact Python ints) or float64
    tname: str = "u64"        # key word type: 'u64' | 'i64'
    psize: int = 0            # 0 -> packed_monomial (one word); k>0 -> d_packed_monomial<T,k>
    _keys: np.ndarray | None = field(default=None, repr=False)

    @property
    def nterms(self) -> int:
        return int(self.expos.shape[0])

    @property
    def nvars(self) -> int:
        return int(self.expos.shape[1])

    @property
    def is_f64(self) -> bool:
        return self.cfs.dtype == np.float64

    @property
    def key_words(self) -> int:
        if self.psize == 0:
            return 1
        return max(1, -(-self.nvars // self.psize))

    @property
    def keys(self) -> np.ndarray:
        """(n, W) uint64 packed keys (two's complement bit pattern for signed types)."""
        if self._keys is None:
            if self.psize == 0:
                self._keys = kpack.pack_many(self.expos, self.tname).reshape(-1, 1)
            else:
                k = kpack.pack_many_d(self.expos, self.psize, self.tname)
                if k.shape[1] == 0:
                    k = np.zeros((self.nterms, 1), dtype=np.uint64)
                self._keys = k
        return self._keys

    def with_layout(self, tname: str | None = None, psize: int | None = None) -> "Operand":
        return Operand(self.symbols, self.expos, self.cfs, tname or self.tname,
                       self.psize if psize is None else psize)

    def as_f64(self) -> "Operand":
        return Operand(self.symbols, self.expos, np.array([float(c) for c in self.cfs], dtype=np.float64),
                       self.tname, self.psize)

    def take(self, idx) -> "Operand":
        return Operand(self.symbols, self.expos[idx], self.cfs[idx], self.tname, self.psize)

    def to_dict(self) -> dict:
        return {tuple(int(v) for v in e): c for e, c in zip(self.expos, self.cfs)}


def from_dict(symbols, d: dict, tname="u64", psize=0, f64=False) -> Operand:
    nv = len(symbols)
    items = sorted(d.items())
    expos = np.array([k for k, _ in items], dtype=np.int64).reshape(len(items), nv)
    if f64:
        cfs = np.array([float(v) for _, v in items], dtype=np.float64)
    else:
        cfs = np.array([int(v) for _, v in items], dtype=object)
    return Operand(tuple(symbols), expos, cfs, tname, psize)


def _compositions_le(nvars: int, n: int) -> np.ndarray:
    """All exponent tuples of length nvars with sum <= n, lexicographic."""
    out = []

    def rec(prefix, left, k):
        if k == 0:
            out.append(prefix)
            return
        for e in range(left + 1):
            rec(prefix + (e,), left - e, k - 1)

    rec((), n, nvars)
    return np.array(out, dtype=np.int64).reshape(len(out), nvars)


def _multinomial(n: int, parts) -> int:
    r = n - sum(parts)
    v = math.factorial(n) // math.factorial(r)
    for p in parts:
        v //= math.factorial(p)
    return v


def fateman(n: int, nvars: int = 4, tname: str = "u64", psize: int = 0):
    """f = (1 + x_1 + ... + x_nvars)^n, g = f + 1.  |f| = C(n+nvars, nvars)."""
    names = tuple(sorted(["x", "y", "z", "t", "u", "v", "w"][:nvars]))
    expos = _compositions_le(nvars, n)
    cfs = np.array([_multinomial(n, tuple(int(v) for v in e)) for e in expos], dtype=object)
    f = Operand(names, expos, cfs, tname, psize)
    gc = cfs.copy()
    zero = np.flatnonzero(expos.sum(axis=1) == 0)
    gc[zero[0]] = gc[zero[0]] + 1
    g = Operand(names, expos.copy(), gc, tname, psize)
    return f, g


def fateman_expected(n: int, nvars: int = 4) -> dict:
    """Closed form of f*(f+1) = (1+..)^(2n) + (1+..)^n  (SURVEY 8c): exponent tuple -> exact coefficient."""
    res = {}
    for e in _compositions_le(nvars, 2 * n):
        t = tuple(int(v) for v in e)
        c = _multinomial(2 * n, t)
        if sum(t) <= n:
            c += _multinomial(n, t)
        res[t] = c
    return res


def _mp_side(n: int, var_of_choice, scales=(1, 1, 2, 3, 5), powers=(1, 1, 2, 3, 5)):
    """(1 + a + b + 2 c^2 + 3 d^3 + 5 e^5)^n with choices mapped to variable indices in (t,u,x,y,z)."""
    comp = _compositions_le(5, n)
    expos = np.zeros((comp.shape[0], 5), dtype=np.int64)
    cfs = np.empty(comp.shape[0], dtype=object)
    for r, cnt in enumerate(comp):
        c = _multinomial(n, tuple(int(v) for v in cnt))
        for j in range(5):
            k = int(cnt[j])
            expos[r, var_of_choice[j]] += powers[j] * k
            c *= scales[j] ** k
        cfs[r] = c
    return expos, cfs


def sparse_mp(n: int, tname: str = "u64", psize: int = 0):
    """Monagan-Pearce sparse pair; symbols sorted as (t,u,x,y,z)."""
    names = ("t", "u", "x", "y", "z")
    idx = {s: i for i, s in enumerate(names)}
    fe, fc = _mp_side(n, [idx["x"], idx["y"], idx["z"], idx["t"], idx["u"]])
    ge, gc = _mp_side(n, [idx["u"], idx["t"], idx["z"], idx["y"], idx["x"]])
    return Operand(names, fe, fc, tname, psize), Operand(names, ge, gc, tname, psize)


def mp_base(tname: str = "i64"):
    """The 6-term base 1+x+y+2z^2+3t^3+5u^5 (symbols t,u,x,y,z)."""
    f, _ = sparse_mp(1, tname)
    return f


def audi(tname: str = "u64", psize: int = 0, nvars: int = 10, n: int = 10):
    """f = (11 + sum x_i)^n, g = (-9 - sum x_i)^n as fp64 polynomials; multiply truncated at degree n."""
    names = tuple(sorted(f"x{i}" for i in range(1, nvars + 1)))
    expos = _compositions_le(nvars, n)
    fc = np.empty(expos.shape[0], dtype=np.float64)
    gc = np.empty(expos.shape[0], dtype=np.float64)
    for r, e in enumerate(expos):
        t = tuple(int(v) for v in e)
        d = sum(t)
        m = _multinomial(n, t)
        fc[r] = float(m * (nvars + 1) ** (n - d))
        gc[r] = float(m * (-(nvars - 1)) ** (n - d) * (-1) ** d)
    return Operand(names, expos, fc, tname, psize), Operand(names, expos.copy(), gc, tname, psize)


def rect_base(tname: str = "u64", f64: bool = True):
    """13-term f of benchmark/rectangular_01.cpp:35-38 in (x,y,z)."""
    terms = {  # (x, y, z) -> coefficient
        (1, 3, 2): 1, (2, 2, 1): 1, (1, 3, 1): 1, (1, 2, 2): 1, (0, 3, 2): 1, (0, 3, 1): 1,
        (0, 2, 2): 2, (1, 1, 1): 2, (0, 2, 1): 1, (0, 1, 2): 1, (0, 2, 0): 1, (0, 1, 1): 2, (0, 0, 1): 1,
    }
    return from_dict(("x", "y", "z"), terms, tname, 0, f64)


def random_poly(rng: np.random.Generator, nterms: int, nvars: int, max_expo: int, cf_bits: int = 20,
                tname: str = "u64", psize: int = 0, f64: bool = False, signed_expos: bool = False,
                signed_cfs: bool = True) -> Operand:
    """Random sparse polynomial with distinct monomials (seeded by the caller's Generator)."""
    seen = {}
    lo = -max_expo if signed_expos else 0
    if nterms > (max_expo - lo + 1) ** nvars:
        raise ValueError("more terms requested than distinct monomials exist")
    while len(seen) < nterms:
        need = nterms - len(seen)
        e = rng.integers(lo, max_expo + 1, size=(need * 2 + 4, nvars))
        for row in e:
            t = tuple(int(v) for v in row)
            if t not in seen:
                if f64:
                    seen[t] = float(rng.standard_normal())
                else:
                    if cf_bits <= 62:
                        mag = int(rng.integers(1, 1 << cf_bits))
                    else:
                        # 62 random leading bits, then the rest in chunks of at most 62 bits
                        mag, left = int(rng.integers(1, 1 << 62)), cf_bits - 62
                        while left > 0:
                            c = min(left, 62)
                            mag = mag << c | int(rng.integers(0, 1 << c))
                            left -= c
                    sgn = -1 if (signed_cfs and rng.integers(0, 2)) else 1
                    seen[t] = sgn * mag
                if len(seen) == nterms:
                    break
    names = tuple(f"v{i:02d}" for i in range(nvars))
    return from_dict(names, seen, tname, psize, f64)


# Named workloads used by bench.py / tests; sizes from BASELINE.md section 2.
def named(name: str):
    """Return (f, g, trunc_limit_or_None, description)."""
    name = name.upper()
    if name.startswith("F") and name[1:].isdigit():
        n = int(name[1:])
        f, g = fateman(n, 4, "u64")
        return f, g, None, f"Fateman (1+x+y+z+t)^{n} * ((1+x+y+z+t)^{n}+1), packed u64, exact integer"
    if name == "F30W":
        # BASELINE config 3 variant: every coefficient times 2^40 -> two-limb inputs, 3-limb accumulators
        # would overflow, so n is 24 here: inputs <= 88 bits, products <= 192 bits with the term-count margin
        f, g = fateman(24, 4, "u64")
        f = Operand(f.symbols, f.expos, np.array([int(c) << 40 for c in f.cfs], dtype=object), f.tname, f.psize)
        g = Operand(g.symbols, g.expos, np.array([int(c) << 40 for c in g.cfs], dtype=object), g.tname, g.psize)
        return f, g, None, "Fateman n=24 with every coefficient scaled by 2^40 (two-limb inputs, three-limb accumulators)"
    if name == "D5":
        f, g = fateman(14, 5, "u64")
        return f, g, None, "dense_02: (1+x+y+z+t+u)^14 * (..+1), packed u64, exact integer"
    if name.startswith("S") and name[1:].isdigit():
        n = int(name[1:])
        f, g = sparse_mp(n, "u64")
        return f, g, None, f"Monagan-Pearce sparse n={n}, packed u64, exact integer"
    if name == "S16T":
        f, g = sparse_mp(16, "u64")
        return f, g, 300, "sparse n=16 truncated_mul(f,g,300), packed u64, exact integer"
    if name == "AUDI":
        f, g = audi("u64", 0)
        return f, g, 10, "audi_01: (11+sum x)^10 * (-9-sum x)^10 truncated at degree 10, packed u64, fp64"
    if name == "AUDID":
        f, g = audi("u64", 8)
        return f, g, 10, "audi_01 as d_packed_monomial<u64,8> (2 words), fp64"
    raise KeyError(name)
