// bench_adopt.cpp -- CPU-only timing of the host container rebuild that follows obk_poly_mul in the C++ mirror
// (polynomial<>::_adopt_grouped, the stand-in for the reference's in-place table fill, polynomial.hpp:1441-1450 /
// series.hpp:1294-1305).  The engine's host result is imitated: `n` distinct keys grouped by hash & (2^k - 1) with
// offsets, `limbs` two's-complement coefficient limbs per term.  No GPU, no engine: this isolates the part of
// e2e_cpp (bench.py) that is host work.
//
//   bench_adopt [n_terms] [log2_tables] [limbs] [reps] [threads]      defaults: S12's result shape 5821335 10 2 5 0
//   (threads 0 = all hardware threads)
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <vector>

#include <obake/integer.hpp>
#include <obake/polynomials/packed_monomial.hpp>
#include <obake/polynomials/polynomial.hpp>

using namespace obake;
using K = packed_monomial<std::uint64_t>;
using poly_t = polynomial<K, integer>;
using clk = std::chrono::steady_clock;

int main(int argc, char **argv)
{
    const std::size_t n = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 5821335u;
    const unsigned k = argc > 2 ? (unsigned)std::atoi(argv[2]) : 10u;
    const unsigned L = argc > 3 ? (unsigned)std::atoi(argv[3]) : 2u;
    const int reps = argc > 4 ? std::atoi(argv[4]) : 5;
    const unsigned nthreads = argc > 5 ? (unsigned)std::atoi(argv[5]) : 0u;
    // distinct keys: a multiplicative walk modulo 2^61 (odd multiplier: a bijection), then grouped by the low k bits
    std::vector<std::uint64_t> keys(n), cfs(n * L);
    std::uint64_t x = 12345;
    for (std::size_t i = 0; i < n; ++i) {
        x = (x * 6364136223846793005ull + 1442695040888963407ull);
        keys[i] = ((std::uint64_t)i * 0x9E3779B97F4A7C15ull) >> 3; // distinct: odd multiplier, top bits kept
        for (unsigned l = 0; l < L; ++l) cfs[i * L + l] = l + 1u < L ? x + l : (x >> 40); // positive, L limbs
    }
    std::vector<std::uint64_t> cnt((std::size_t(1) << k) + 1u, 0);
    const std::uint64_t mask = (std::uint64_t(1) << k) - 1u;
    for (auto v : keys) ++cnt[(v & mask) + 1u];
    std::partial_sum(cnt.begin(), cnt.end(), cnt.begin());
    std::vector<std::uint64_t> gk(n), gc(n * L);
    {
        auto pos = cnt;
        for (std::size_t i = 0; i < n; ++i) {
            const auto p = pos[keys[i] & mask]++;
            gk[p] = keys[i];
            for (unsigned l = 0; l < L; ++l) gc[p * L + l] = cfs[i * L + l];
        }
    }
    symbol_set ss{"t", "u", "x", "y", "z"};
    double best = 1e300, sum = 0, dsum = 0;
    std::size_t sz = 0;
    for (int r = 0; r <= reps; ++r) {
        auto t0 = clk::now();
        double ms;
        {
            poly_t p;
            p.set_symbol_set(ss);
            p._adopt_grouped(n, k, cnt.data(), [&](std::size_t i) {
                return std::pair<K, integer>(K::from_words(&gk[i], 5), integer::from_limbs(&gc[i * L], L));
            }, nthreads);
            ms = std::chrono::duration<double, std::milli>(clk::now() - t0).count();
            sz = p.size();
            t0 = clk::now();
        }
        const double dms = std::chrono::duration<double, std::milli>(clk::now() - t0).count();
        if (r == 0) continue;
        best = std::min(best, ms);
        sum += ms;
        dsum += dms;
    }
    std::printf("{\"terms\": %zu, \"log2_tables\": %u, \"limbs\": %u, \"adopt_ms_mean\": %.3f, \"adopt_ms_best\": %.3f, "
                "\"destroy_ms_mean\": %.3f, \"slot_bytes\": %zu}\n",
                sz, k, L, sum / reps, best, dsum / reps, sizeof(std::pair<K, integer>));
    return 0;
}
