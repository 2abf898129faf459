#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ from the reference tree.

Run in the build container only (it reads /root/reference, which does not exist
on the GPU box):  python tests/golden/make_golden.py

Outputs
  kpack_tables.json  -- the four Kronecker-packing tables (deltas, lims, klims,
                        divcnst) for int32/uint32/int64/uint64, parsed as NUMBERS
                        out of /root/reference/src/kpack.cpp:39-47,161-169,
                        289-338,786-836 and the divcnst blocks that follow.
  known_answers.json -- the known-answer constants asserted by the reference's
                        own tests for the series-product path, transcribed with
                        the file:line they come from.
"""
import json
import os
import re
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def parse_kpack_tables():
    src = open(os.path.join(REF, "src/kpack.cpp")).read()
    out = {}
    tnames = {"::std::int32_t": "i32", "::std::uint32_t": "u32", "::std::int64_t": "i64", "::std::uint64_t": "u64"}
    for cxx, short in tnames.items():
        tab = {}
        for field in ("deltas", "lims", "klims"):
            m = re.search(r"kpack_data<%s>::%s\[(\d+)\]\s*=\s*\{(.*?)\};" % (re.escape(cxx), field), src, re.S)
            assert m, (cxx, field)
            n = int(m.group(1))
            nums = [int(x) for x in re.findall(r"(-?\d+)(?:u?l?l?)", m.group(2))]
            assert len(nums) == n, (cxx, field, len(nums), n)
            tab[field] = [str(v) for v in nums]
        m = re.search(r"kpack_data<%s>::divcnst\[(\d+)\]\[(\d+)\]\s*=\s*\{(.*?)\};" % re.escape(cxx), src, re.S)
        assert m, (cxx, "divcnst")
        n1, n2 = int(m.group(1)), int(m.group(2))
        triples = re.findall(r"\{(\d+)ul*,\s*(\d+)u,\s*(\d+)u\}", m.group(3))
        # Rows are brace-initialised and may be shorter than n2 (the rest is zero-init).
        rows = []
        body = m.group(3)
        depth = 0
        cur = None
        i = 0
        # split top-level rows
        row_texts = []
        start = None
        for i, ch in enumerate(body):
            if ch == "{":
                depth += 1
                if depth == 1:
                    start = i
            elif ch == "}":
                depth -= 1
                if depth == 0:
                    row_texts.append(body[start : i + 1])
        assert len(row_texts) == n1, (cxx, len(row_texts), n1)
        for rt in row_texts:
            tr = re.findall(r"\{(\d+)ul*,\s*(\d+)u,\s*(\d+)u\}", rt)
            row = [[str(int(a)), int(b), int(c)] for a, b, c in tr]
            while len(row) < n2:
                row.append(["0", 0, 0])
            assert len(row) == n2
            rows.append(row)
        tab["divcnst"] = rows
        out[short] = tab
    return out


KNOWN = {
    "_comment": "Known-answer constants asserted by the reference's own tests (file:line under /root/reference).",
    "sparse_n10_terms": {"value": 2096600, "src": "test/polynomials_polynomial_00.cpp:398-424",
                         "what": "(1+x+y+2z^2+3t^3+5u^5)^10 * (1+u+t+2z^2+3y^3+5x^5)^10, size() for double and integer<1>"},
    "rect_pow20_terms": {"value": 53130, "src": "test/polynomials_polynomial_04.cpp:264-285",
                         "what": "(1+x+y+2z^2+3t^3+5u^5)^20 built by 19 rectangular products, size()"},
    "monomial_mul": {"a": [1, 2, 3], "b": [4, 5, 6], "c": [5, 7, 9],
                     "src": "test/polynomials_packed_monomial_00.cpp:486-492"},
    "overflow_message": {"value": "An overflow in the monomial exponents was detected while attempting to multiply two polynomials",
                         "src": "include/obake/polynomials/polynomial.hpp:847-850; test/polynomials_polynomial_00.cpp:225-238"},
    "truncated_n8": {"src": "test/polynomials_polynomial_01.cpp:185-236",
                     "what": "n=8 sparse pair, limits 1000 and 80 equal the untruncated product; 40 -> degree 40; 5 -> degree 5; 0 -> the single constant term; -1 -> empty",
                     "limits": [1000, 80, 40, 5, 0, -1]},
    "truncate_degree_50": {"src": "test/polynomials_polynomial_04.cpp:239-262",
                           "what": "truncated_mul(f,g,50) == truncate_degree(f*g,50) for the n=8 sparse pair"},
}


def series_ops_cases():
    """Hand-checkable cases for the steps either side of the product (series add/sub, series.hpp:2195-2991;
    truncate_degree / truncate_p_degree, polynomial.hpp:2364-2445; series_sym_extender, series.hpp:539-625), in the
    exponent-tuple form the Python oracle uses.  The expected outputs are written out by hand in
    tests/test_oracle.py::test_series_ops_golden; this file carries the same cases for the GPU tests."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
    import pyoracle

    def pack(d):
        return [[list(k), v] for k, v in sorted(d.items())]

    x = {(1, 0): 1, (0, 1): 1}
    y = {(1, 0): 1, (0, 1): -1}
    b = {(0, 0): 1, (1, 0): 1, (0, 1): 1}
    p = pyoracle.poly_mul(pyoracle.poly_mul(b, b), b)
    return [
        {"op": "add", "x": pack(x), "y": pack(y), "out": pack(pyoracle.poly_addsub(x, y))},
        {"op": "sub", "x": pack(x), "y": pack(y), "out": pack(pyoracle.poly_addsub(x, y, True))},
        {"op": "truncate", "x": pack(p), "limit": 2, "idx": None, "out": pack(pyoracle.truncate_degree(p, 2))},
        {"op": "truncate", "x": pack(p), "limit": 1, "idx": [0], "out": pack(pyoracle.truncate_degree(p, 1, [0]))},
        {"op": "extend", "x": pack(b), "new_nvars": 4, "pos": [1, 3], "out": pack(pyoracle.extend_symbols(b, 4, [1, 3]))},
    ]


def main():
    with open(os.path.join(HERE, "series_ops.json"), "w") as f:
        json.dump(series_ops_cases(), f, indent=1)
    if not os.path.isdir(REF):
        print("reference tree not present; nothing regenerated", file=sys.stderr)
        return 1
    tabs = parse_kpack_tables()
    with open(os.path.join(HERE, "kpack_tables.json"), "w") as f:
        json.dump(tabs, f, indent=0, separators=(",", ":"))
    with open(os.path.join(HERE, "known_answers.json"), "w") as f:
        json.dump(KNOWN, f, indent=1)
    print("wrote kpack_tables.json, known_answers.json")
    return 0


if __name__ == "__main__":
    sys.exit(main())
