"""Python host layer over the C ABI (include/obk_b200.h): the call a user of the Python package makes.

It mirrors the reference's entry points for the series-product path -- ``mul`` is
``obake::polynomials::series_mul`` / ``operator*`` (polynomial.hpp:2026-2032) and ``mul(..., limit=,
symbols=)`` is ``truncated_mul`` (polynomial.hpp:2113-2127) -- and maps the C status codes onto the
exception types the reference throws (std::overflow_error -> OverflowError, std::invalid_argument ->
ValueError).  All arithmetic happens in the CUDA library; nothing here computes a product.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import OBK_MEM_DEVICE, OBK_MEM_HOST, PolyResult, PolyView, Stats, Trunc


class ObkCudaError(RuntimeError):
    pass


class ObkUnsupported(NotImplementedError):
    pass


def _raise(status: int, msg: str):
    if status in (_capi.OBK_ERR_KEY_OVERFLOW, _capi.OBK_ERR_TABLE_OVERFLOW, _capi.OBK_ERR_CF_OVERFLOW):
        raise OverflowError(msg)
    if status == _capi.OBK_ERR_INVALID_ARG:
        raise ValueError(msg)
    if status == _capi.OBK_ERR_UNSUPPORTED:
        raise ObkUnsupported(msg)
    if status == _capi.OBK_ERR_NOMEM:
        raise MemoryError(msg)
    raise ObkCudaError(msg)


def int_to_limbs(vals, nlimbs: int) -> np.ndarray:
    """Exact Python ints -> (n, nlimbs) uint64, two's complement, little-endian limbs."""
    n = len(vals)
    if nlimbs == 1:
        try:
            return np.array([int(v) for v in vals], dtype=np.int64).view(np.uint64).reshape(n, 1)
        except OverflowError:
            pass
    out = np.zeros((n, nlimbs), dtype=np.uint64)
    mod = 1 << (64 * nlimbs)
    m64 = (1 << 64) - 1
    for i, v in enumerate(vals):
        v = int(v) % mod
        for j in range(nlimbs):
            out[i, j] = (v >> (64 * j)) & m64
    return out


def limbs_to_int(arr: np.ndarray) -> list:
    n, L = arr.shape
    if L == 1:
        return [int(v) for v in arr[:, 0].view(np.int64)]
    half = 1 << (64 * L - 1)
    mod = 1 << (64 * L)
    cols = [arr[:, j].tolist() for j in range(L)]
    res = []
    for i in range(n):
        v = 0
        for j in range(L):
            v |= cols[j][i] << (64 * j)
        res.append(v - mod if v >= half else v)
    return res


def needed_limbs(vals) -> int:
    m = 1
    for v in vals:
        v = int(v)
        m = max(m, (v if v >= 0 else ~v).bit_length() + 1)
    return -(-m // 64)


class HostOperand:
    """An :class:`obake_b200.workloads.Operand` marshalled to the SoA arrays the C ABI views."""

    def __init__(self, op, cf_limbs: int | None = None):
        self.op = op
        self.keys = np.ascontiguousarray(op.keys, dtype=np.uint64)
        if op.is_f64:
            self.cfs = np.ascontiguousarray(op.cfs, dtype=np.float64)
            kind, limbs = _capi.OBK_CF_F64, 1
        else:
            limbs = cf_limbs or needed_limbs(op.cfs)
            self.cfs = int_to_limbs(op.cfs, limbs)
            kind = _capi.OBK_CF_INT
        self.view = PolyView(op.nterms, self.keys.shape[1], 1 if op.tname.startswith("i") else 0, op.nvars,
                             op.psize, kind, limbs, OBK_MEM_HOST, 0, self.keys.ctypes.data, self.cfs.ctypes.data)


def raw_view(nterms, key_words, key_signed, nvars, psize, cf_kind, cf_limbs, mem, keys_ptr, cfs_ptr,
             key_bits=64) -> PolyView:
    """View over caller-owned buffers (torch tensors: pass ``t.data_ptr()``), host-pinned or device."""
    return PolyView(int(nterms), int(key_words), int(key_signed), int(nvars), int(psize), int(cf_kind),
                    int(cf_limbs), int(mem), int(key_bits), int(keys_ptr), int(cfs_ptr))


class Product:
    """Result of one product; owns engine memory until :meth:`free` (or garbage collection)."""

    def __init__(self, engine: "Engine", res: PolyResult, stats: Stats):
        self._engine = engine
        self._res = res
        self.stats = stats.as_dict()

    @property
    def nterms(self) -> int:
        return int(self._res.nterms)

    @property
    def key_words(self) -> int:
        return int(self._res.key_words)

    @property
    def cf_limbs(self) -> int:
        return int(self._res.cf_limbs)

    @property
    def is_f64(self) -> bool:
        return self._res.cf_kind == _capi.OBK_CF_F64

    @property
    def log2_nsegs(self) -> int:
        return int(self._res.log2_nsegs)

    @property
    def nsegs(self) -> int:
        """Number of groups of seg_offsets."""
        return int(self._res.nsegs)

    @property
    def seg_kind(self) -> int:
        """0: group = hash & (nsegs-1) (series tables); 1: group = key mod nsegs (multi-GPU partials)."""
        return int(self._res.seg_kind)

    @property
    def on_device(self) -> bool:
        return self._res.mem == OBK_MEM_DEVICE

    @property
    def keys_ptr(self) -> int:
        return int(self._res.keys or 0)

    @property
    def cfs_ptr(self) -> int:
        return int(self._res.cfs or 0)

    @property
    def seg_offsets(self) -> np.ndarray:
        n = self.nsegs + 1 if self._res.seg_offsets else 0
        if n == 0:
            return np.zeros(2, dtype=np.uint64)
        return np.ctypeslib.as_array(C.cast(self._res.seg_offsets, C.POINTER(C.c_uint64)), (n,)).copy()

    def table_stats(self) -> dict:
        """What series::table_stats() reports for a series rebuilt from this result (series.hpp:1386-1409): number
        of terms and tables, smallest / largest / average table, bytes of the SoA term arrays."""
        so = self.seg_offsets.astype(np.int64)
        sizes = np.diff(so) if len(so) > 1 else np.zeros(1, dtype=np.int64)
        return {"total_terms": self.nterms, "n_tables": int(len(sizes)), "min_table": int(sizes.min()),
                "max_table": int(sizes.max()), "avg_table": float(sizes.mean()),
                "bytes": self.nterms * 8 * (self.key_words + self.cf_limbs)}

    def _host_array(self, ptr, shape, ctype):
        if self.on_device:
            import torch  # plumbing only: copy a device buffer back for inspection

            class _Dev:
                pass

            n = int(np.prod(shape))
            if n == 0:
                return np.empty(shape, dtype=np.dtype(ctype))
            d = _Dev()
            d.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (int(ptr), False), "version": 3,
                                          "strides": None}
            t = torch.as_tensor(d, device=torch.device("cuda", self._engine.device))
            return t.cpu().numpy().view(np.dtype(ctype)).reshape(shape).copy()
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape).copy()

    @property
    def keys(self) -> np.ndarray:
        n, W = self.nterms, self.key_words
        if n == 0:
            return np.zeros((0, W), dtype=np.uint64)
        return self._host_array(self._res.keys, (n, W), C.c_uint64)

    @property
    def cf_limb_array(self) -> np.ndarray:
        n = self.nterms
        if n == 0:
            return np.zeros((0, self.cf_limbs), dtype=np.uint64)
        return self._host_array(self._res.cfs, (n, self.cf_limbs), C.c_uint64)

    @property
    def cfs(self):
        """Exact Python ints (integer coefficients) or a float64 array."""
        n = self.nterms
        if self.is_f64:
            if n == 0:
                return np.zeros(0, dtype=np.float64)
            return self._host_array(self._res.cfs, (n,), C.c_double)
        return limbs_to_int(self.cf_limb_array)

    def as_dict(self) -> dict:
        """{key tuple (W words as Python ints): coefficient}."""
        k = self.keys
        c = self.cfs
        if self.key_words == 1:
            return dict(zip(k[:, 0].tolist(), c if not self.is_f64 else c.tolist()))
        return {tuple(row): v for row, v in zip(k.tolist(), c if not self.is_f64 else c.tolist())}

    def free(self):
        if self._res is not None and self._res.owner:
            _capi.lib().obk_result_free(self._engine._h, C.byref(self._res))
        self._res = None

    def __del__(self):
        try:
            if self._res is not None and self._engine._h:
                self.free()
        except Exception:
            pass


class Engine:
    """One CUDA device + stream + grow-only workspace (obk_engine)."""

    def __init__(self, device=0):
        """`device`: one CUDA device ordinal, or a list of ordinals for the single-process multi-device engine
        (obk_engine_create_multi: host operands in, host result out, every device forms the output it owns)."""
        L = _capi.lib()
        h = C.c_void_p()
        if isinstance(device, (list, tuple)):
            devs = (C.c_int * len(device))(*[int(d) for d in device])
            st = L.obk_engine_create_multi(devs, len(device), C.byref(h))
            what = "obk_engine_create_multi"
            self.devices = [int(d) for d in device]
            device = self.devices[0] if self.devices else 0
        else:
            st = L.obk_engine_create(int(device), C.byref(h))
            what = "obk_engine_create"
            self.devices = [int(device)]
        if st != 0:
            raise ObkCudaError(what + " failed: " + L.obk_last_error(None).decode() + " (no CPU fallback exists)")
        self._h = h
        self.device = device

    def close(self):
        if self._h:
            _capi.lib().obk_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, name: str, value: int):
        st = _capi.lib().obk_engine_set_option(self._h, name.encode(), int(value))
        if st != 0:
            _raise(st, self.last_error())

    def set_stream(self, cuda_stream_ptr: int):
        _capi.lib().obk_engine_set_stream(self._h, C.c_void_p(int(cuda_stream_ptr)))

    def last_error(self) -> str:
        return _capi.lib().obk_last_error(self._h).decode()

    @staticmethod
    def _trunc(limit, idx, keepalive: list):
        if limit is None:
            return None
        if idx is None:
            return Trunc(int(limit), 0, 0, None)
        ia = np.ascontiguousarray(sorted(int(i) for i in idx), dtype=np.uint32)
        keepalive.append(ia)
        return Trunc(int(limit), len(ia), 1, ia.ctypes.data if len(ia) else None)

    def mul_views(self, vx: PolyView, vy: PolyView, limit=None, idx=None, result_mem=OBK_MEM_HOST) -> Product:
        keep = []
        tr = self._trunc(limit, idx, keep)
        res, st = PolyResult(), Stats()
        rc = _capi.lib().obk_poly_mul(self._h, C.byref(vx), C.byref(vy), C.byref(tr) if tr is not None else None,
                                      int(result_mem), C.byref(res), C.byref(st))
        if rc != 0:
            _raise(rc, self.last_error())
        return Product(self, res, st)

    def mul(self, x, y, limit=None, idx=None, result_mem=OBK_MEM_HOST) -> Product:
        """x * y (series_mul) or truncated_mul(x, y, limit[, symbols as positions idx])."""
        lin = None
        if not x.is_f64:
            lin = max(needed_limbs(x.cfs), needed_limbs(y.cfs))
        hx = HostOperand(x, lin)
        hy = hx if y is x else HostOperand(y, lin)
        return self.mul_views(hx.view, hy.view, limit, idx, result_mem)

    def mul_partial_views(self, vx, vy, rank, nranks, limit=None, idx=None):
        keep = []
        tr = self._trunc(limit, idx, keep)
        res, st = PolyResult(), Stats()
        counts = (C.c_uint64 * nranks)()
        rc = _capi.lib().obk_poly_mul_partial(self._h, C.byref(vx), C.byref(vy),
                                              C.byref(tr) if tr is not None else None, int(rank), int(nranks),
                                              C.byref(res), counts, C.byref(st))
        if rc != 0:
            _raise(rc, self.last_error())
        return Product(self, res, st), [int(c) for c in counts]

    def view_of(self, prod: "Product", like: PolyView) -> PolyView:
        """View over a product this engine left in DEVICE (or host) memory, so that it can be the operand of the
        next call without leaving HBM (f *= g chains).  `like` supplies the key layout."""
        return raw_view(prod.nterms, prod.key_words, like.key_signed, like.nvars, like.psize,
                        _capi.OBK_CF_F64 if prod.is_f64 else _capi.OBK_CF_INT, prod.cf_limbs,
                        OBK_MEM_DEVICE if prod.on_device else OBK_MEM_HOST, prod.keys_ptr, prod.cfs_ptr,
                        like.key_bits or 64)

    def copy_view(self, vx: PolyView, result_mem=OBK_MEM_HOST) -> Product:
        """obk_poly_copy: the terms of `vx` in engine-owned HOST (pinned, grouped the way a series stores its terms) or
        DEVICE memory -- the upload that starts a device-resident chain and the download that ends it."""
        res = PolyResult()
        rc = _capi.lib().obk_poly_copy(self._h, C.byref(vx), int(result_mem), C.byref(res))
        if rc != 0:
            _raise(rc, self.last_error())
        return Product(self, res, Stats())

    def addsub_views(self, vx: PolyView, vy: PolyView, subtract=False, result_mem=OBK_MEM_HOST) -> Product:
        """x + y / x - y over the same symbol set (series.hpp:2195-2991) on the engine's layout."""
        res, st = PolyResult(), Stats()
        rc = _capi.lib().obk_poly_addsub(self._h, C.byref(vx), C.byref(vy), 1 if subtract else 0, int(result_mem),
                                         C.byref(res), C.byref(st))
        if rc != 0:
            _raise(rc, self.last_error())
        return Product(self, res, st)

    def add(self, x, y, result_mem=OBK_MEM_HOST) -> Product:
        hx, hy = HostOperand(x), HostOperand(y)  # keep the marshalled arrays alive across the call
        return self.addsub_views(hx.view, hy.view, False, result_mem)

    def sub(self, x, y, result_mem=OBK_MEM_HOST) -> Product:
        hx, hy = HostOperand(x), HostOperand(y)
        return self.addsub_views(hx.view, hy.view, True, result_mem)

    def truncate_degree_view(self, vx: PolyView, limit, idx=None, result_mem=OBK_MEM_HOST) -> Product:
        """truncate_degree(x, limit) / truncate_p_degree(x, limit, symbols at positions idx)
        (polynomial.hpp:2364-2445): keeps the terms with not (limit < degree)."""
        keep = []
        tr = self._trunc(limit, idx, keep)
        res, st = PolyResult(), Stats()
        rc = _capi.lib().obk_poly_truncate_degree(self._h, C.byref(vx), C.byref(tr), int(result_mem), C.byref(res),
                                                  C.byref(st))
        if rc != 0:
            _raise(rc, self.last_error())
        return Product(self, res, st)

    def truncate_degree(self, x, limit, idx=None, result_mem=OBK_MEM_HOST) -> Product:
        hx = HostOperand(x)
        return self.truncate_degree_view(hx.view, limit, idx, result_mem)

    def extend_symbols_view(self, vx: PolyView, new_nvars: int, pos, result_mem=OBK_MEM_HOST) -> Product:
        """Re-express x over a larger symbol set (series_sym_extender, series.hpp:539-625): pos[j] is the position
        of x's j-th symbol in the new set; every key is re-packed on the device."""
        pa = np.ascontiguousarray([int(p) for p in pos], dtype=np.uint32)
        res = PolyResult()
        rc = _capi.lib().obk_poly_extend_symbols(self._h, C.byref(vx), int(new_nvars),
                                                 pa.ctypes.data if len(pa) else None, int(result_mem), C.byref(res))
        if rc != 0:
            _raise(rc, self.last_error())
        return Product(self, res, Stats())

    def extend_symbols(self, x, new_nvars: int, pos, result_mem=OBK_MEM_HOST) -> Product:
        hx = HostOperand(x)
        return self.extend_symbols_view(hx.view, new_nvars, pos, result_mem)

    def merge_view(self, vterms: PolyView, nsegs: int, result_mem=OBK_MEM_HOST) -> Product:
        res, st = PolyResult(), Stats()
        rc = _capi.lib().obk_poly_merge(self._h, C.byref(vterms), int(nsegs), int(result_mem), C.byref(res),
                                        C.byref(st))
        if rc != 0:
            _raise(rc, self.last_error())
        return Product(self, res, st)

    def microbench(self, name: str = "int_mac") -> float:
        """Measured ceiling of this GPU for the named primitive (include/obk_b200.h, obk_microbench)."""
        v = C.c_double(0.0)
        rc = _capi.lib().obk_microbench(self._h, name.encode(), C.byref(v))
        if rc != 0:
            _raise(rc, self.last_error())
        return float(v.value)

    def overflow_check(self, x, y) -> bool:
        hx, hy = HostOperand(x, 1 if not x.is_f64 else None), HostOperand(y, 1 if not y.is_f64 else None)
        ok = C.c_int(1)
        rc = _capi.lib().obk_overflow_check(self._h, C.byref(hx.view), C.byref(hy.view), C.byref(ok))
        if rc != 0:
            _raise(rc, self.last_error())
        return bool(ok.value)

    def degrees(self, x, idx=None) -> np.ndarray:
        hx = HostOperand(x, 1 if not x.is_f64 else None)
        keep = []
        tr = self._trunc(0, idx, keep) if idx is not None else None
        out = np.zeros(x.nterms, dtype=np.int64)
        rc = _capi.lib().obk_key_degrees(self._h, C.byref(hx.view), C.byref(tr) if tr is not None else None,
                                         out.ctypes.data_as(C.c_void_p))
        if rc != 0:
            _raise(rc, self.last_error())
        return out
