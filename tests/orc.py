"""ctypes access to the CPU oracle (oracle/_build/libobk_oracle.so) -- for tests, smoke() and the
cpu_baseline / --impl reference legs of bench.py ONLY.  Never imported by obake_b200/."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


class OrcPoly(C.Structure):
    _fields_ = [("nterms", C.c_uint64), ("key_words", C.c_uint32), ("key_signed", C.c_uint32),
                ("nvars", C.c_uint32), ("psize", C.c_uint32), ("cf_kind", C.c_uint32), ("cf_limbs", C.c_uint32),
                ("keys", C.c_void_p), ("cfs", C.c_void_p)]


class OrcTrunc(C.Structure):
    _fields_ = [("limit", C.c_int64), ("nidx", C.c_uint32), ("idx", C.c_void_p)]


class OrcResult(C.Structure):
    _fields_ = [("nterms", C.c_uint64), ("key_words", C.c_uint32), ("cf_kind", C.c_uint32),
                ("cf_limbs", C.c_uint32), ("log2_nsegs", C.c_uint32), ("keys", C.c_void_p), ("cfs", C.c_void_p),
                ("seg_offsets", C.c_void_p), ("n_mults", C.c_uint64), ("est_nterms", C.c_uint64),
                ("seconds", C.c_double)]


def build(force: bool = False) -> str:
    so = os.path.join(ROOT, "oracle", "_build", "libobk_oracle.so")
    src = os.path.join(ROOT, "oracle", "obk_oracle.cpp")
    if force or not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_poly_mul.restype = C.c_int
        _LIB.orc_overflow_check.restype = C.c_int
        _LIB.orc_hardware_concurrency.restype = C.c_int
    return _LIB


def int_to_limbs(vals, nlimbs: int) -> np.ndarray:
    """Exact Python ints -> (n, nlimbs) uint64 two's complement, little endian limbs."""
    n = len(vals)
    out = np.zeros((n, nlimbs), dtype=np.uint64)
    mod = 1 << (64 * nlimbs)
    m64 = (1 << 64) - 1
    for i, v in enumerate(vals):
        v = int(v) % mod
        for j in range(nlimbs):
            out[i, j] = (v >> (64 * j)) & m64
    return out


def limbs_to_int(arr: np.ndarray) -> list:
    """(n, L) uint64 two's complement -> list of Python ints."""
    n, L = arr.shape
    half = 1 << (64 * L - 1)
    mod = 1 << (64 * L)
    res = []
    for i in range(n):
        v = 0
        for j in range(L):
            v |= int(arr[i, j]) << (64 * j)
        if v >= half:
            v -= mod
        res.append(v)
    return res


def needed_limbs(vals) -> int:
    m = 0
    for v in vals:
        v = int(v)
        m = max(m, (v if v >= 0 else ~v).bit_length() + 1)
    return max(1, -(-m // 64))


class _View:
    """Keeps the numpy buffers alive next to the C struct."""

    def __init__(self, op, cf_limbs=None):
        self.keys = np.ascontiguousarray(op.keys, dtype=np.uint64)
        if op.is_f64:
            self.cfs = np.ascontiguousarray(op.cfs, dtype=np.float64)
            kind, limbs = 1, 1
        else:
            limbs = cf_limbs or needed_limbs(op.cfs)
            self.cfs = int_to_limbs(op.cfs, limbs)
            kind = 0
        self.s = OrcPoly(op.nterms, self.keys.shape[1] if self.keys.ndim == 2 else 1,
                         1 if op.tname.startswith("i") else 0, op.nvars, op.psize, kind, limbs,
                         self.keys.ctypes.data, self.cfs.ctypes.data)


def acc_limbs_bound(x, y) -> int:
    """A-priori accumulator width: bits(max|c1|)+bits(max|c2|)+bits(min(nx,ny))+1 sign bit."""
    if x.is_f64:
        return 1
    b1 = max((abs(int(c)).bit_length() for c in x.cfs), default=0)
    b2 = max((abs(int(c)).bit_length() for c in y.cfs), default=0)
    bits = b1 + b2 + max(1, min(x.nterms, y.nterms)).bit_length() + 1
    L = -(-bits // 64)
    return L if L <= 3 else 5


def poly_mul(x, y, limit=None, idx=None, nthreads=1, lacc=None):
    """Run the C++ oracle.  Returns dict(keys=(n,W) uint64, cfs=list[int] | float64 array, seg_offsets, log2_nsegs,
    n_mults, seconds, status)."""
    L = lib()
    lin = None
    if not x.is_f64:
        lin = max(needed_limbs(x.cfs), needed_limbs(y.cfs))
    vx, vy = _View(x, lin), _View(y, lin)
    if lacc is None:
        lacc = acc_limbs_bound(x, y)
        if lin == 2 and lacc < 3:
            lacc = 3
        if lin == 3:
            lacc = 5
    tr = None
    if limit is not None:
        ia = np.ascontiguousarray(sorted(idx), dtype=np.uint32) if idx is not None else None
        tr = OrcTrunc(int(limit), 0 if ia is None else len(ia), None if ia is None else ia.ctypes.data)
        if ia is not None and len(ia) == 0:
            # partial degree over no symbol: every degree is 0; express as total degree of nothing
            raise NotImplementedError("empty partial-degree set: handle in the caller")
    res = OrcResult()
    rc = L.orc_poly_mul(C.byref(vx.s), C.byref(vy.s), C.byref(tr) if tr is not None else None, int(nthreads),
                        int(lacc), C.byref(res))
    out = {"status": rc, "n_mults": int(res.n_mults), "seconds": float(res.seconds)}
    if rc != 0:
        return out
    n, W = int(res.nterms), int(res.key_words)
    nseg = 1 << int(res.log2_nsegs)
    out["log2_nsegs"] = int(res.log2_nsegs)
    out["seg_offsets"] = np.ctypeslib.as_array(C.cast(res.seg_offsets, C.POINTER(C.c_uint64)), (nseg + 1,)).copy()
    if n:
        out["keys"] = np.ctypeslib.as_array(C.cast(res.keys, C.POINTER(C.c_uint64)), (n, W)).copy()
        if res.cf_kind == 1:
            out["cfs"] = np.ctypeslib.as_array(C.cast(res.cfs, C.POINTER(C.c_double)), (n,)).copy()
        else:
            lim = np.ctypeslib.as_array(C.cast(res.cfs, C.POINTER(C.c_uint64)), (n, int(res.cf_limbs))).copy()
            out["cf_limbs"] = lim
            out["cfs"] = limbs_to_int(lim)
    else:
        out["keys"] = np.zeros((0, W), dtype=np.uint64)
        out["cfs"] = np.zeros(0) if res.cf_kind == 1 else []
    L.orc_result_free(C.byref(res))
    return out


def overflow_check(x, y) -> bool:
    vx, vy = _View(x, 1 if not x.is_f64 else None), _View(y, 1 if not y.is_f64 else None)
    return bool(lib().orc_overflow_check(C.byref(vx.s), C.byref(vy.s)))


def degrees(x, idx=None) -> np.ndarray:
    vx = _View(x, 1 if not x.is_f64 else None)
    out = np.zeros(x.nterms, dtype=np.int64)
    if idx is None:
        lib().orc_degrees(C.byref(vx.s), None, 0, out.ctypes.data_as(C.c_void_p))
    else:
        ia = np.ascontiguousarray(sorted(idx), dtype=np.uint32)
        lib().orc_degrees(C.byref(vx.s), ia.ctypes.data_as(C.c_void_p), len(ia), out.ctypes.data_as(C.c_void_p))
    return out


def unpack_keys(x) -> np.ndarray:
    vx = _View(x, 1 if not x.is_f64 else None)
    out = np.zeros((x.nterms, x.nvars), dtype=np.int64)
    lib().orc_unpack_keys(C.byref(vx.s), out.ctypes.data_as(C.c_void_p))
    return out


def result_to_dict(res, like) -> dict:
    """Oracle / engine result -> {exponent tuple: coefficient} using the operand's key layout."""
    from obake_b200 import workloads as wl

    keys = res["keys"]
    tmp = wl.Operand(like.symbols, np.zeros((keys.shape[0], like.nvars), dtype=np.int64),
                     np.zeros(keys.shape[0]), like.tname, like.psize, keys)
    ex = unpack_keys(tmp)
    return {tuple(int(v) for v in e): c for e, c in zip(ex, res["cfs"])}
