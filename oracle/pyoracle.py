"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Pure-Python restatement of what obake's polynomial product *means*, on exponent tuples and
arbitrary-precision Python ints (or Python floats).  It is deliberately independent of the Kronecker
packing and of the segmented algorithm, so it pins the C++ oracle (oracle/obk_oracle.cpp) and the
CUDA engine from a second direction.

Follows the parity rules of the reference (file:line under /root/reference):
  * product set: all |x|*|y| pairs, or -- truncated -- pair (i, j) contributes iff
    not (max_deg < deg_i + deg_j)                    include/obake/polynomials/polynomial.hpp:604-610,1222-1229
  * partial degree: only the symbols in the given set  include/obake/series.hpp:3911-3924
  * accumulation: first hit assigns c1*c2, later hits add (fp64: two roundings, no fma)
                                                     polynomial.hpp:1368-1384, math/fma3.hpp:65-72
  * terms whose coefficient ends up == 0 are erased  polynomial.hpp:1528-1540

Parity: pinned (tests/test_oracle.py checks it against the reference's known answers in
tests/golden/known_answers.json and the closed form for Fateman products).
"""
from __future__ import annotations


def degree(expo, idx=None) -> int:
    if idx is None:
        return sum(expo)
    return sum(expo[i] for i in idx)


def poly_mul(x: dict, y: dict, max_deg=None, idx=None) -> dict:
    """x, y: {exponent tuple: coefficient}.  Returns the product as the same kind of dict."""
    out = {}
    if not x or not y:
        return out
    if max_deg is not None:
        ydeg = [(degree(e, idx), e, c) for e, c in y.items()]
    for e1, c1 in x.items():
        if max_deg is None:
            for e2, c2 in y.items():
                k = tuple(a + b for a, b in zip(e1, e2))
                p = c1 * c2
                if k in out:
                    out[k] = out[k] + p
                else:
                    out[k] = p
        else:
            d1 = degree(e1, idx)
            for d2, e2, c2 in ydeg:
                if max_deg < d1 + d2:
                    continue
                k = tuple(a + b for a, b in zip(e1, e2))
                p = c1 * c2
                if k in out:
                    out[k] = out[k] + p
                else:
                    out[k] = p
    return {k: v for k, v in out.items() if v != 0}


def n_mults(x: dict, y: dict, max_deg=None, idx=None) -> int:
    """tot_n_mults of poly_mul_estimate_product_size (polynomial.hpp:574-623)."""
    if max_deg is None:
        return len(x) * len(y)
    yd = sorted(degree(e, idx) for e in y)
    import bisect
    tot = 0
    for e in x:
        tot += bisect.bisect_right(yd, max_deg - degree(e, idx))
    return tot


def truncate_degree(p: dict, max_deg, idx=None) -> dict:
    """polynomial.hpp:2364-2445 -- drop terms whose (partial) degree exceeds max_deg."""
    return {e: c for e, c in p.items() if not (max_deg < degree(e, idx))}


def poly_pow(base: dict, n: int) -> dict:
    nv = len(next(iter(base)))
    r = {tuple([0] * nv): 1}
    for _ in range(n):
        r = poly_mul(r, base)
    return r


def overflow_check(x_expos, y_expos, lim_min: int, lim_max: int) -> bool:
    """packed_monomial.hpp:462-487: interval test on per-variable min/max of each operand."""
    if len(x_expos) == 0 or len(y_expos) == 0:
        return True
    nv = len(x_expos[0])
    for j in range(nv):
        mn = min(e[j] for e in x_expos) + min(e[j] for e in y_expos)
        mx = max(e[j] for e in x_expos) + max(e[j] for e in y_expos)
        if mn < lim_min or mx > lim_max:
            return False
    return True


def poly_addsub(x: dict, y: dict, subtract: bool = False) -> dict:
    """x + y or x - y for two polynomials over the SAME symbol set: the identical-symbol-set branch of
    series_default_addsub_impl (include/obake/series.hpp:2195-2991): every term of the second operand is inserted
    into a copy of the first with its coefficient added (or subtracted, a new term enters negated), and terms whose
    coefficient becomes zero are erased (series_add_term_table, series.hpp:683-792)."""
    out = dict(x)
    for e, c in y.items():
        if e in out:
            out[e] = out[e] - c if subtract else out[e] + c
        else:
            out[e] = -c if subtract else c
    return {k: v for k, v in out.items() if v != 0}


def truncate_degree(x: dict, max_deg, idx=None) -> dict:
    """truncate_degree / truncate_p_degree (include/obake/polynomials/polynomial.hpp:2364-2445): filter() keeping
    the terms with not (max_deg < degree); partial degree over the symbol positions idx."""
    return {e: c for e, c in x.items() if not (max_deg < degree(e, idx))}


def extend_symbols(x: dict, new_nvars: int, pos) -> dict:
    """series_sym_extender (include/obake/series.hpp:539-625) on exponent tuples: x's j-th exponent moves to
    position pos[j] of the larger symbol set, new symbols get exponent 0 (key_merge_symbols,
    src/polynomials/packed_monomial.cpp:234-292)."""
    out = {}
    for e, c in x.items():
        n = [0] * new_nvars
        for j, v in enumerate(e):
            n[pos[j]] = v
        out[tuple(n)] = c
    return out
