"""Kronecker packing of exponent vectors into one machine integer (host side, Python).

Mirrors the reference's ``kpacker`` / ``kunpacker`` (include/obake/kpack.hpp:219-375) and the
constant tables of src/kpack.cpp:39-47,161-169,289-338,786-836.  The tables are NOT copied: they are
regenerated here from the recipe of tools/kpack.ipynb and checked against the reference's numbers in
tests/test_kpack.py (fixture tests/golden/kpack_tables.json).

    bw      = bit width (unsigned) or bit width - 1 (signed)
    delta_1 = 2**bw - 1
    delta_s = largest prime <= floor(2**(bw/s)) - 1            for s >= 2 while bw // s >= 3
    lim_s   = delta_s - 1 (unsigned)  |  (delta_s - 1) // 2 (signed, exponents in [-lim, lim])
    klim_s  = sum_{j<s} lim_s * delta_s**j
    key     = sum_j e_j * delta_s**j      (first symbol = lowest digit, kpack.hpp:262-267)
"""
from __future__ import annotations

import functools

import numpy as np

TYPES = {
    "i32": (32, True),
    "u32": (32, False),
    "i64": (64, True),
    "u64": (64, False),
}


def _iroot(n: int, k: int) -> int:
    """floor(n ** (1/k)) for non-negative integers."""
    if k == 1:
        return n
    lo, hi = 0, 1 << (n.bit_length() // k + 1)
    while lo < hi:
        mid = (lo + hi + 1) >> 1
        if mid**k <= n:
            lo = mid
        else:
            hi = mid - 1
    return lo


def _is_prime(n: int) -> bool:
    if n < 2:
        return False
    small = (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37)
    for p in small:
        if n % p == 0:
            return n == p
    d, r = n - 1, 0
    while d % 2 == 0:
        d //= 2
        r += 1
    for a in small:  # deterministic for n < 3.3e24
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(r - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


@functools.lru_cache(maxsize=None)
def tables(tname: str) -> dict:
    """Return {'deltas', 'lims', 'klims', 'divcnst', 'max_size', 'bits', 'signed'} for a packable type."""
    bits, signed = TYPES[tname]
    bw = bits - 1 if signed else bits
    deltas, lims, klims, divcnst = [], [], [], []
    s = 1
    while bw // s >= 3:
        if s == 1:
            d = (1 << bw) - 1
        else:
            d = _iroot(1 << bw, s) - 1
            while not _is_prime(d):
                d -= 1
        lim = (d - 1) // 2 if signed else d - 1
        klim = sum(lim * d**j for j in range(s))
        deltas.append(d)
        lims.append(lim)
        klims.append(klim)
        s += 1
    max_size = len(deltas)
    n_bits = bits
    for s in range(1, max_size + 1):
        row = []
        for j in range(s + 1):
            dd = deltas[s - 1] ** j
            ell = (dd - 1).bit_length()  # ceil(log2 dd)
            mp = (1 << (n_bits + ell)) // dd - (1 << n_bits) + 1
            row.append((mp, min(ell, 1), max(ell - 1, 0)))
        while len(row) < max_size + 1:
            row.append((0, 0, 0))
        divcnst.append(row)
    return {
        "deltas": deltas,
        "lims": lims,
        "klims": klims,
        "divcnst": divcnst,
        "max_size": max_size,
        "bits": bits,
        "signed": signed,
    }


def delta(tname: str, size: int) -> int:
    return tables(tname)["deltas"][size - 1]


def lims(tname: str, size: int) -> tuple[int, int]:
    """(lim_min, lim_max) for one packed component (kpack_get_lims)."""
    t = tables(tname)
    lim = t["lims"][size - 1]
    return (-lim, lim) if t["signed"] else (0, lim)


def klims(tname: str, size: int) -> tuple[int, int]:
    """(klim_min, klim_max) for the packed value (kpack_get_klims)."""
    t = tables(tname)
    k = t["klims"][size - 1]
    return (-k, k) if t["signed"] else (0, k)


def pack(expos, tname: str = "u64", size: int | None = None) -> int:
    """Pack one exponent vector; missing trailing components count as zero.  Raises like kpacker."""
    expos = list(expos)
    size = len(expos) if size is None else size
    t = tables(tname)
    if size == 0:
        return 0
    if size > t["max_size"]:
        raise OverflowError(f"Invalid size {size}: the maximum possible size is {t['max_size']}")
    if len(expos) > size:
        raise IndexError("Cannot push any more values to this Kronecker packer")
    lo, hi = lims(tname, size)
    d = t["deltas"][size - 1]
    v, cur = 0, 1
    for e in expos:
        e = int(e)
        if e < lo or e > hi:
            raise OverflowError(f"Cannot push the value {e}: outside the allowed range [{lo}, {hi}]")
        v += e * cur
        cur *= d
    return v


def unpack(value: int, size: int, tname: str = "u64") -> list[int]:
    """Inverse of :func:`pack` (kunpacker semantics, incl. the range check on the coded value)."""
    t = tables(tname)
    if size == 0:
        if value != 0:
            raise ValueError("Only a value of zero can be used in a Kronecker unpacker with a size of zero")
        return []
    if size > t["max_size"]:
        raise OverflowError(f"Invalid size {size}")
    kmin, kmax = klims(tname, size)
    if value < kmin or value > kmax:
        raise OverflowError(f"The value {value} is outside the allowed range [{kmin}, {kmax}]")
    d = t["deltas"][size - 1]
    lmin, _ = lims(tname, size)
    n = value - kmin
    out = []
    for _ in range(size):
        out.append(n % d + lmin)
        n //= d
    return out


def unpack_divcnst(value: int, size: int, tname: str = "u64") -> list[int]:
    """Unpack exactly as the reference does it: two division-by-constant steps per component
    (Granlund-Montgomery fig. 4.1, kpack.hpp:344-369) using the divcnst table."""
    t = tables(tname)
    bits = t["bits"]
    mask = (1 << bits) - 1
    kmin, _ = klims(tname, size)
    lmin, _ = lims(tname, size)
    d = t["deltas"][size - 1]
    n = (value - kmin) & mask

    def div(nn, mp, sh1, sh2):
        t1 = (mp * nn) >> bits
        return ((t1 + (((nn - t1) & mask) >> sh1)) & mask) >> sh2

    out = []
    cur = 1
    for idx in range(size):
        cur *= d
        mp_d, s1d, s2d = t["divcnst"][size - 1][idx]
        mp_r, s1r, s2r = t["divcnst"][size - 1][idx + 1]
        q_r = div(n, mp_r, s1r, s2r)
        rem = (n - q_r * (cur & mask)) & mask
        q_d = div(rem, mp_d, s1d, s2d)
        out.append(q_d + lmin)
    return out


# ---------------------------------------------------------------------------------------------
# numpy-vectorised helpers used by the workload generators and the ctypes host layer
# ---------------------------------------------------------------------------------------------


def pack_many(expos: np.ndarray, tname: str = "u64", size: int | None = None) -> np.ndarray:
    """Pack the rows of an (n, nvars) integer array into n keys (Python-int exact, returned as uint64 bits)."""
    expos = np.asarray(expos)
    n, nv = expos.shape
    size = nv if size is None else size
    if size == 0:
        return np.zeros(n, dtype=np.uint64)
    t = tables(tname)
    lo, hi = lims(tname, size)
    if expos.size and (expos.min() < lo or expos.max() > hi):
        raise OverflowError("exponent outside the allowed range")
    d = t["deltas"][size - 1]
    acc = np.zeros(n, dtype=object)
    cur = 1
    for j in range(nv):
        acc = acc + expos[:, j].astype(object) * cur
        cur *= d
    mask = (1 << 64) - 1
    return np.array([int(v) & mask for v in acc], dtype=np.uint64)


def pack_many_d(expos: np.ndarray, psize: int, tname: str = "u64") -> np.ndarray:
    """d_packed_monomial layout: ceil(nvars/psize) words, each packing psize exponents with the
    size-psize delta; the last word is zero padded (d_packed_monomial.hpp:110-132)."""
    expos = np.asarray(expos)
    n, nv = expos.shape
    nw = (nv + psize - 1) // psize
    out = np.zeros((n, nw), dtype=np.uint64)
    for w in range(nw):
        chunk = expos[:, w * psize : (w + 1) * psize]
        out[:, w] = pack_many(chunk, tname, psize)
    return out
