// bench_cpp.cpp -- end-to-end timing of obake's operator* through the C++ host mirror over the C ABI.
//
// What a user of the reference runs (benchmark/dense.hpp:21-44 `dense_benchmark_4_vars`, benchmark/sparse.hpp:21-50
// `sparse_benchmark`): the operands are built with the library's own arithmetic, then ONE product `ret = f * g` is
// timed.  Here the timed call is the mirror's operator*: marshal the host containers to SoA (polynomial.hpp:828-833),
// obk_poly_mul with HOST buffers (H2D + product + D2H), and rebuild the host container term by term
// (series.hpp:1294-1305) -- the whole path, not the C-ABI call alone.  One JSON object per workload on stdout.
//
//   bench_cpp [F<n> | S<n>]... [--reps R]        e.g.  bench_cpp F30 S12
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <obake/integer.hpp>
#include <obake/polynomials/packed_monomial.hpp>
#include <obake/polynomials/polynomial.hpp>

using namespace obake;
using poly_t = polynomial<packed_monomial<std::uint64_t>, integer>;
using clk = std::chrono::steady_clock;

static double ms_since(clk::time_point t0)
{
    return std::chrono::duration<double, std::milli>(clk::now() - t0).count();
}

struct Operands {
    poly_t f, g;
    double build_ms;
};

// (1 + x + y + z + t)^n and the same plus one
static Operands dense4(int n)
{
    const auto t0 = clk::now();
    auto [t, x, y, z] = make_polynomials<poly_t>(symbol_set{"t", "x", "y", "z"}, "t", "x", "y", "z");
    poly_t f = x + y + z + t + 1;
    const poly_t base = f;
    for (int i = 1; i < n; ++i) f *= base;
    poly_t g = f + 1;
    return {std::move(f), std::move(g), ms_since(t0)};
}

// (1 + x + y + 2 z^2 + 3 t^3 + 5 u^5)^n  and  (1 + u + t + 2 z^2 + 3 y^3 + 5 x^5)^n
static Operands sparse5(int n)
{
    const auto t0 = clk::now();
    auto [t, u, x, y, z] = make_polynomials<poly_t>(symbol_set{"t", "u", "x", "y", "z"}, "t", "u", "x", "y", "z");
    poly_t f = x + y + z * z * 2 + t * t * t * 3 + u * u * u * u * u * 5 + 1;
    poly_t g = u + t + z * z * 2 + y * y * y * 3 + x * x * x * x * x * 5 + 1;
    const poly_t bf = f, bg = g;
    for (int i = 1; i < n; ++i) {
        f *= bf;
        g *= bg;
    }
    return {std::move(f), std::move(g), ms_since(t0)};
}

int main(int argc, char **argv)
{
    int reps = 5;
    std::vector<std::string> names;
    for (int i = 1; i < argc; ++i) {
        if (std::strcmp(argv[i], "--reps") == 0 && i + 1 < argc) {
            reps = std::atoi(argv[++i]);
        } else {
            names.emplace_back(argv[i]);
        }
    }
    if (names.empty()) names = {"F30"};
    for (const auto &nm : names) {
        if (nm.size() < 2 || (nm[0] != 'F' && nm[0] != 'S')) {
            std::fprintf(stderr, "unknown workload %s (F<n> or S<n>)\n", nm.c_str());
            return 2;
        }
        const int n = std::atoi(nm.c_str() + 1);
        Operands op = nm[0] == 'F' ? dense4(n) : sparse5(n);
        double best = 1e300, sum = 0;
        std::size_t terms = 0;
        for (int r = 0; r < reps + 1; ++r) { // the first repetition warms the engine up (allocations, module load)
            const auto t0 = clk::now();
            const poly_t ret = op.f * op.g;
            const double ms = ms_since(t0);
            terms = ret.size();
            if (r == 0) continue;
            best = ms < best ? ms : best;
            sum += ms;
        }
        if (std::getenv("OBK_BENCH_TABLE_STATS")) {
            // what the reference's drivers print after the product (benchmark/dense.hpp:41, sparse.hpp:44)
            const poly_t ret = op.f * op.g;
            std::fprintf(stderr, "%s", ret.table_stats().c_str());
        }
        const double products = (double)op.f.size() * (double)op.g.size();
        std::printf("{\"workload\": \"%s\", \"e2e_cpp_ms\": %.4f, \"e2e_cpp_best_ms\": %.4f, \"reps\": %d, \"terms_f\": %zu, "
                    "\"terms_g\": %zu, \"terms_out\": %zu, \"products\": %.0f, \"value\": %.6g, \"build_operands_ms\": %.2f, "
                    "\"what\": \"obake::operator* of the C++ host mirror: marshal + obk_poly_mul (host buffers) + container rebuild\"}\n",
                    nm.c_str(), sum / reps, best, reps, op.f.size(), op.g.size(), terms, products, products / (sum / reps * 1e-3),
                    op.build_ms);
        std::fflush(stdout);
    }
    return 0;
}
