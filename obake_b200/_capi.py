"""ctypes binding of include/obk_b200.h (libobk_b200.so).  There is no fallback: if the CUDA library is
missing or cannot be loaded, importing this module's `lib()` raises."""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "lib", "libobk_b200.so")

OBK_OK = 0
OBK_ERR_INVALID_ARG = 1
OBK_ERR_KEY_OVERFLOW = 2
OBK_ERR_TABLE_OVERFLOW = 3
OBK_ERR_CF_OVERFLOW = 4
OBK_ERR_CUDA = 5
OBK_ERR_UNSUPPORTED = 6
OBK_ERR_NOMEM = 7
OBK_CF_INT, OBK_CF_F64 = 0, 1
OBK_MEM_HOST, OBK_MEM_DEVICE = 0, 1
OBK_EXCHANGE_MOD = 4093  # partial products of the multi-GPU path are grouped by key mod this prime


class PolyView(C.Structure):
    _fields_ = [("nterms", C.c_uint64), ("key_words", C.c_uint32), ("key_signed", C.c_uint32),
                ("nvars", C.c_uint32), ("psize", C.c_uint32), ("cf_kind", C.c_uint32), ("cf_limbs", C.c_uint32),
                ("mem", C.c_uint32), ("key_bits", C.c_uint32), ("keys", C.c_void_p), ("cfs", C.c_void_p)]


class Trunc(C.Structure):
    _fields_ = [("limit", C.c_int64), ("nidx", C.c_uint32), ("partial", C.c_uint32), ("idx", C.c_void_p)]


class PeerWindows(C.Structure):
    _fields_ = [("nranks", C.c_uint32), ("rank", C.c_uint32), ("parity", C.c_uint32), ("reserved", C.c_uint32),
                ("window_bytes", C.c_uint64), ("base", C.POINTER(C.c_void_p))]


class PolyResult(C.Structure):
    _fields_ = [("nterms", C.c_uint64), ("key_words", C.c_uint32), ("cf_kind", C.c_uint32),
                ("cf_limbs", C.c_uint32), ("log2_nsegs", C.c_uint32), ("mem", C.c_uint32), ("nsegs", C.c_uint32),
                ("seg_kind", C.c_uint32), ("reserved", C.c_uint32),
                ("keys", C.c_void_p), ("cfs", C.c_void_p), ("seg_offsets", C.c_void_p), ("owner", C.c_void_p)]


class Stats(C.Structure):
    _fields_ = [("term_products", C.c_uint64), ("nterms_out", C.c_uint64), ("est_nterms", C.c_uint64),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("log2_nsegs", C.c_uint32), ("nsegs", C.c_uint32), ("acc_limbs", C.c_uint32), ("retries", C.c_uint32),
                ("kernel", C.c_uint32), ("table_slots", C.c_uint32), ("group_lanes", C.c_uint32),
                ("mul_launches", C.c_uint32), ("total_launches", C.c_uint32), ("reserved", C.c_uint32),
                ("ms_total", C.c_float), ("ms_prep", C.c_float), ("ms_estimate", C.c_float), ("ms_mul", C.c_float),
                ("ms_compact", C.c_float), ("ms_h2d", C.c_float), ("ms_d2h", C.c_float)]

    def as_dict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_}


# Every symbol include/obk_b200.h declares (tests/test_abi.py checks the library exports all of them).
SYMBOLS = [
    "obk_engine_create", "obk_engine_destroy", "obk_engine_set_stream", "obk_engine_set_option",
    "obk_last_error", "obk_status_string", "obk_version", "obk_poly_mul", "obk_result_free",
    "obk_overflow_check", "obk_key_degrees", "obk_poly_mul_partial", "obk_poly_merge",
    "obk_poly_addsub", "obk_poly_truncate_degree", "obk_poly_copy", "obk_poly_extend_symbols",
    "obk_peer_window_alloc", "obk_peer_window_open", "obk_peer_window_close", "obk_poly_mul_push", "obk_peer_merge",
    "obk_peer_window_reset", "obk_microbench", "obk_engine_create_multi",
]

_LIB = None


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m obake_b200.build` (nvcc, sm_100a). "
            "The engine has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.obk_engine_create.restype = C.c_int
    L.obk_engine_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    L.obk_engine_destroy.restype = None
    L.obk_engine_destroy.argtypes = [C.c_void_p]
    L.obk_engine_set_stream.restype = C.c_int
    L.obk_engine_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.obk_engine_set_option.restype = C.c_int
    L.obk_engine_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
    L.obk_last_error.restype = C.c_char_p
    L.obk_last_error.argtypes = [C.c_void_p]
    L.obk_status_string.restype = C.c_char_p
    L.obk_status_string.argtypes = [C.c_int]
    L.obk_version.restype = C.c_char_p
    L.obk_poly_mul.restype = C.c_int
    L.obk_poly_mul.argtypes = [C.c_void_p, C.POINTER(PolyView), C.POINTER(PolyView), C.POINTER(Trunc), C.c_uint32,
                               C.POINTER(PolyResult), C.POINTER(Stats)]
    L.obk_result_free.restype = None
    L.obk_result_free.argtypes = [C.c_void_p, C.POINTER(PolyResult)]
    L.obk_overflow_check.restype = C.c_int
    L.obk_overflow_check.argtypes = [C.c_void_p, C.POINTER(PolyView), C.POINTER(PolyView), C.POINTER(C.c_int)]
    L.obk_key_degrees.restype = C.c_int
    L.obk_key_degrees.argtypes = [C.c_void_p, C.POINTER(PolyView), C.POINTER(Trunc), C.c_void_p]
    L.obk_poly_mul_partial.restype = C.c_int
    L.obk_poly_mul_partial.argtypes = [C.c_void_p, C.POINTER(PolyView), C.POINTER(PolyView), C.POINTER(Trunc),
                                       C.c_uint32, C.c_uint32, C.POINTER(PolyResult), C.POINTER(C.c_uint64),
                                       C.POINTER(Stats)]
    L.obk_poly_merge.restype = C.c_int
    L.obk_poly_merge.argtypes = [C.c_void_p, C.POINTER(PolyView), C.c_uint32, C.c_uint32, C.POINTER(PolyResult),
                                 C.POINTER(Stats)]
    L.obk_poly_addsub.restype = C.c_int
    L.obk_poly_addsub.argtypes = [C.c_void_p, C.POINTER(PolyView), C.POINTER(PolyView), C.c_int, C.c_uint32,
                                  C.POINTER(PolyResult), C.POINTER(Stats)]
    L.obk_poly_truncate_degree.restype = C.c_int
    L.obk_poly_truncate_degree.argtypes = [C.c_void_p, C.POINTER(PolyView), C.POINTER(Trunc), C.c_uint32,
                                           C.POINTER(PolyResult), C.POINTER(Stats)]
    L.obk_poly_copy.restype = C.c_int
    L.obk_poly_copy.argtypes = [C.c_void_p, C.POINTER(PolyView), C.c_uint32, C.POINTER(PolyResult)]
    L.obk_poly_extend_symbols.restype = C.c_int
    L.obk_poly_extend_symbols.argtypes = [C.c_void_p, C.POINTER(PolyView), C.c_uint32, C.c_void_p, C.c_uint32,
                                          C.POINTER(PolyResult)]
    L.obk_peer_window_alloc.restype = C.c_int
    L.obk_peer_window_alloc.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p), C.c_void_p]
    L.obk_peer_window_open.restype = C.c_int
    L.obk_peer_window_open.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
    L.obk_peer_window_close.restype = C.c_int
    L.obk_peer_window_close.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.obk_poly_mul_push.restype = C.c_int
    L.obk_poly_mul_push.argtypes = [C.c_void_p, C.POINTER(PolyView), C.POINTER(PolyView), C.POINTER(Trunc),
                                    C.POINTER(PeerWindows), C.POINTER(Stats)]
    L.obk_peer_merge.restype = C.c_int
    L.obk_peer_merge.argtypes = [C.c_void_p, C.POINTER(PeerWindows), C.POINTER(PolyView), C.c_uint32, C.c_uint32,
                                 C.POINTER(PolyResult), C.POINTER(Stats)]
    L.obk_engine_create_multi.restype = C.c_int
    L.obk_engine_create_multi.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]
    L.obk_microbench.restype = C.c_int
    L.obk_microbench.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_double)]
    L.obk_peer_window_reset.restype = C.c_int
    L.obk_peer_window_reset.argtypes = [C.c_void_p, C.POINTER(PeerWindows)]
    _LIB = L
    return L
