"""obake_b200 -- B200-native engine for ONE path of bluescarni/obake: the product of two sparse multivariate
polynomials with Kronecker-packed monomials (operator* / truncated_mul).

    from obake_b200 import Engine, workloads
    e = Engine(0)                         # one CUDA device; raises without a B200 (no CPU fallback)
    f, g = workloads.fateman(20)          # or build an Operand from exponent tuples: workloads.from_dict(...)
    p = e.mul(f, g)                       # series_mul;  e.mul(f, g, limit=10[, idx=[...]]) is truncated_mul
    p.nterms, p.keys, p.cfs, p.stats

The C ABI underneath is include/obk_b200.h (obake_b200/lib/libobk_b200.so, built by `python -m obake_b200.build`);
the obake-shaped C++20 host API lives in obake_b200/include/obake/.
"""
from . import kpack, workloads  # noqa: F401


def __getattr__(name):
    # engine (ctypes + the CUDA library) is imported lazily so that CPU-only tooling can use kpack / workloads
    if name in ("Engine", "Product"):
        import importlib

        return getattr(importlib.import_module(".engine", __name__), name)
    raise AttributeError(name)
