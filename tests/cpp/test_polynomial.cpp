// C++ parity tests for the drop-in boundary, written after the reference's own Catch2 tests for the
// series-product path (file:line cited per case).  Everything that multiplies goes through
// obake::polynomial::operator* / truncated_mul -> C ABI -> CUDA engine.
//   ./test_polynomial          all tests (needs a B200)
//   ./test_polynomial --cpu    host-only tests (kpack, monomials, integer, symbol sets)
#include <cstdio>
#include <algorithm>
#include <cstring>
#include <unordered_map>
#include <iostream>
#include <random>
#include <sstream>
#include <stdexcept>
#include <string>

#include <obake/polynomials/polynomial.hpp>
#include <obake/b200/device_polynomial.hpp>
#include <obake/power_series/power_series.hpp>

using namespace obake;

static int g_fail = 0, g_checks = 0;
#define REQUIRE(...)                                                                                                   \
    do {                                                                                                               \
        ++g_checks;                                                                                                    \
        if (!(__VA_ARGS__)) {                                                                                          \
            ++g_fail;                                                                                                  \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #__VA_ARGS__);                                       \
        }                                                                                                              \
    } while (0)
// test/test_utils.hpp:36-47 -- exception type AND message substring are part of the contract
#define REQUIRES_THROWS_CONTAINS(expr, exc, msg)                                                                       \
    do {                                                                                                               \
        ++g_checks;                                                                                                    \
        bool ok_ = false;                                                                                              \
        try {                                                                                                          \
            (void)(expr);                                                                                              \
        } catch (const exc &e_) {                                                                                      \
            ok_ = std::string(e_.what()).find(msg) != std::string::npos;                                               \
            if (!ok_) std::printf("  wrong message: %s\n", e_.what());                                                 \
        } catch (...) {                                                                                                \
        }                                                                                                              \
        if (!ok_) {                                                                                                    \
            ++g_fail;                                                                                                  \
            std::printf("FAILED %s:%d: %s should throw %s containing '%s'\n", __FILE__, __LINE__, #expr, #exc, msg);   \
        }                                                                                                              \
    } while (0)

template <typename T>
static void kpack_tests()
{
    // test/kpack.cpp:101-219 (round trips, extrema, range errors)
    std::mt19937 rng;
    for (unsigned size = 1; size <= detail::kpack_max_size<T>(); ++size) {
        const auto [lo, hi] = detail::kpack_get_lims<T>(size);
        const long double hi_ld = static_cast<long double>(hi);
        std::uniform_int_distribution<long long> dist(static_cast<long long>(lo),
                                                      hi_ld > 4.0e18L ? 4000000000000000000ll : static_cast<long long>(hi));
        for (int t = 0; t < 200; ++t) {
            std::vector<T> v(size), w(size);
            kpacker<T> kp(size);
            for (auto &e : v) {
                e = static_cast<T>(dist(rng));
                kp << e;
            }
            kunpacker<T> ku(kp.get(), size);
            for (auto &e : w) ku >> e;
            REQUIRE(v == w);
        }
        kpacker<T> kmin(size), kmax(size);
        for (unsigned i = 0; i < size; ++i) {
            kmin << lo;
            kmax << hi;
        }
        REQUIRE(kmin.get() == detail::kpack_get_klims<T>(size).first);
        REQUIRE(kmax.get() == detail::kpack_get_klims<T>(size).second);
        if (size > 1u) {
            kpacker<T> kp(size);
            REQUIRES_THROWS_CONTAINS(kp << static_cast<T>(hi + 1), std::overflow_error, "the value is outside the allowed range");
        }
    }
    REQUIRES_THROWS_CONTAINS(kpacker<T>(detail::kpack_max_size<T>() + 1u), std::overflow_error, "the maximum possible size is");
    {
        kpacker<T> kp(2);
        kp << T(1) << T(2);
        REQUIRES_THROWS_CONTAINS(kp << T(3), std::out_of_range, "Cannot push any more values to this Kronecker packer");
    }
    REQUIRES_THROWS_CONTAINS(kunpacker<T>(T(1), 0), std::invalid_argument, "Only a value of zero can be used in a Kronecker unpacker");
}

template <typename T>
static void monomial_tests()
{
    using pm_t = packed_monomial<T>;
    symbol_set ss{"x", "y", "z"};
    // test/polynomials_packed_monomial_00.cpp:472-494: {1,2,3}*{4,5,6} == {5,7,9}
    pm_t a{1, 2, 3}, b{4, 5, 6}, c;
    monomial_mul(c, a, b, ss);
    REQUIRE(c == pm_t{5, 7, 9});
    // :198-210 hash == value
    REQUIRE(hash(a) == static_cast<std::size_t>(a.get_value()));
    // :574-618 homomorphic hash
    std::mt19937 rng;
    std::uniform_int_distribution<int> d(0, 5);
    for (int t = 0; t < 500; ++t) {
        std::vector<T> v1(6), v2(6), v3(6);
        for (int i = 0; i < 6; ++i) {
            v1[i] = static_cast<T>(d(rng));
            v2[i] = static_cast<T>(d(rng));
            v3[i] = static_cast<T>(v1[i] + v2[i]);
        }
        REQUIRE(hash(pm_t(v1)) + hash(pm_t(v2)) == hash(pm_t(v3)));
        using dpm_t = d_packed_monomial<T, 4>;
        REQUIRE(hash(dpm_t(v1)) + hash(dpm_t(v2)) == hash(dpm_t(v3)));
        dpm_t out;
        monomial_mul(out, dpm_t(v1), dpm_t(v2), symbol_set{"a", "b", "c", "d", "e", "f"});
        REQUIRE(out == dpm_t(v3));
        REQUIRE(dpm_t(v1).unpack(6) == v1);
        REQUIRE(key_degree(dpm_t(v1), symbol_set{"a", "b", "c", "d", "e", "f"}) == static_cast<T>(v1[0] + v1[1] + v1[2] + v1[3] + v1[4] + v1[5]));
    }
    REQUIRE(key_degree(a, ss) == T(6));
    REQUIRE(key_p_degree(a, std::vector<unsigned>{0, 2}, ss) == T(4));
    REQUIRE(key_is_compatible(a, ss));
    // overflow check (polynomials_packed_monomial_00.cpp:496-552)
    const auto [lo, hi] = detail::kpack_get_lims<T>(3);
    std::vector<pm_t> r1{pm_t{1, 2, 3}, pm_t{hi, 0, 0}}, r2{pm_t{1, 0, 0}}, r3{pm_t{0, 1, 1}};
    REQUIRE(!monomial_range_overflow_check(r1, r2, ss));
    REQUIRE(monomial_range_overflow_check(r1, r3, ss));
    REQUIRE(monomial_range_overflow_check(std::vector<pm_t>{}, r3, ss));
}

static void integer_tests()
{
    integer a(std::string("123456789012345678901234567890")), b(-42);
    REQUIRE((a * b).to_string() == "-5185185138518518513851851851380");
    REQUIRE((a + b - a) == b);
    REQUIRE(pow(integer(13), 24).to_string() == "542800770374370512771595361");
    std::uint64_t l[3];
    (a * b).to_limbs(l, 3);
    REQUIRE(integer::from_limbs(l, 3) == a * b);
    REQUIRE(integer(-1).limbs_needed() == 1u);
    REQUIRE((integer(1) * integer(std::string("18446744073709551616"))).limbs_needed() == 2u);
    REQUIRE(integer(5) < integer(7) && integer(-7) < integer(-5) && integer(-1) < integer(1));
    REQUIRES_THROWS_CONTAINS(a.to_limbs(l, 1), std::overflow_error, "does not fit");
    // across the inline-limb boundary (three limbs inline, the heap beyond): 13^60 has 223 bits, 13^120 has 445
    const integer p60 = pow(integer(13), 60), p120 = pow(integer(13), 120);
    REQUIRE(p60.to_string() == "6864377172744689378196133203444067624537070830997366604446306636401");
    REQUIRE(p60 * p60 == p120);
    REQUIRE((p120 - p60 * p60).is_zero() && (p120 - p60 * p60).limbs_needed() == 1u);
    integer copy = p120; // heap storage copied, moved, shrunk back
    integer moved = std::move(copy);
    REQUIRE(moved == p120);
    moved = moved - p120 + integer(7);
    REQUIRE(moved == integer(7) && moved.limbs_needed() == 1u);
    std::uint64_t l5[5];
    (integer(0) - p60).to_limbs(l5, 5);
    REQUIRE(integer::from_limbs(l5, 5) == integer(0) - p60);
    REQUIRE(p60.limbs_needed() == 4u);
}

// The term container of the mirror (detail::seg_table, the reference's segmented tables series.hpp:396-400) against
// std::unordered_map: insertion with growth, lookup, accumulate-to-zero erasure (backward-shift deletion), iteration,
// and the parallel adoption of terms grouped by hash & (2^k - 1).
static void seg_table_tests()
{
    using K = packed_monomial<std::uint64_t>;
    using poly_t = polynomial<K, integer>;
    std::uint64_t st = 88172645463325252ull;
    auto rnd = [&st]() {
        st ^= st << 13;
        st ^= st >> 7;
        st ^= st << 17;
        return st;
    };
    const symbol_set ss{"a", "b", "c"};
    poly_t p;
    p.set_symbol_set(ss);
    std::unordered_map<std::uint64_t, long long> ref;
    for (int it = 0; it < 60000; ++it) {
        const std::uint64_t e0 = rnd() % 40u, e1 = rnd() % 40u, e2 = rnd() % 40u;
        const K k{e0, e1, e2};
        std::uint64_t w;
        k.to_words(&w, 3);
        const long long c = static_cast<long long>(rnd() % 7u) - 3;
        if (c == 0) {
            continue;
        }
        p.add_term(k, integer(c));
        auto [pos, fresh] = ref.try_emplace(w, c);
        if (!fresh) {
            pos->second += c;
        }
        if (pos->second == 0) {
            ref.erase(pos);
        }
    }
    REQUIRE(p.size() == ref.size());
    std::size_t seen = 0;
    bool same = true;
    for (const auto &[k, c] : p) {
        std::uint64_t w;
        k.to_words(&w, 3);
        const auto f = ref.find(w);
        same = same && f != ref.end() && c == integer(f->second);
        ++seen;
    }
    REQUIRE(same && seen == ref.size());
    for (const auto &[w, c] : ref) {
        const auto f = p.find(K::from_words(&w, 3));
        same = same && f != p.end() && f->second == integer(c);
    }
    REQUIRE(same);
    REQUIRE(p.find(K{std::uint64_t(41), std::uint64_t(0), std::uint64_t(0)}) == p.end());
    // grouped adoption: the same terms sorted by hash & 63, adopted in parallel, must compare equal
    std::vector<std::pair<std::uint64_t, long long>> terms(ref.begin(), ref.end());
    std::sort(terms.begin(), terms.end(), [](const auto &x, const auto &y) { return (x.first & 63u) < (y.first & 63u); });
    std::vector<std::uint64_t> offs(65, 0);
    for (const auto &t : terms) {
        ++offs[(t.first & 63u) + 1u];
    }
    for (int i = 0; i < 64; ++i) {
        offs[i + 1] += offs[i];
    }
    poly_t q;
    q.set_symbol_set(ss);
    q._adopt_grouped(terms.size(), 6, offs.data(), [&](std::size_t i) {
        return std::pair<K, integer>(K::from_words(&terms[i].first, 3), integer(terms[i].second));
    });
    REQUIRE(q.size() == p.size() && q == p);
    // series::table_stats (series.hpp:1383-1408): the reference's labels, this container's numbers
    {
        const std::string ts = q.table_stats();
        REQUIRE(ts.find("Total number of terms             : " + std::to_string(q.size()) + "\n") != std::string::npos);
        REQUIRE(ts.find("Total number of tables            : 64\n") != std::string::npos);
        std::size_t mn = ~std::size_t(0), mx = 0;
        for (int i = 0; i < 64; ++i) {
            mn = std::min<std::size_t>(mn, offs[i + 1] - offs[i]);
            mx = std::max<std::size_t>(mx, offs[i + 1] - offs[i]);
        }
        REQUIRE(ts.find("Min/max terms per table           : " + std::to_string(mn) + "/" + std::to_string(mx) + "\n")
                != std::string::npos);
        REQUIRE(ts.find("Average terms per table           : ") != std::string::npos);
        REQUIRE(ts.find("Total size in bytes               : ") != std::string::npos);
        REQUIRE(poly_t{}.table_stats().find("Total number of terms             : 0\nTotal number of tables            : 1\nTotal size") == 0u);
    }
    q.add_term(K::from_words(&terms[0].first, 3), integer(-terms[0].second)); // erase inside a grouped table
    REQUIRE(q.size() + 1u == p.size() && q.find(K::from_words(&terms[0].first, 3)) == q.end());
    // value semantics of the slab-backed container: copies are deep, moves leave an empty polynomial behind
    {
        poly_t c(p);
        REQUIRE(c == p);
        c.add_term(K{std::uint64_t(50), std::uint64_t(50), std::uint64_t(50)}, integer(7));
        REQUIRE(c.size() == p.size() + 1u && !(c == p));
        poly_t m(std::move(c));
        REQUIRE(c.empty() && m.size() == p.size() + 1u);
        c = m;
        REQUIRE(c == m);
        m = poly_t{};
        REQUIRE(m.empty() && c.size() == p.size() + 1u);
        c.clear();
        REQUIRE(c.empty() && c.begin() == c.end());
        c.add_term(K{std::uint64_t(1), std::uint64_t(2), std::uint64_t(3)}, integer(5)); // usable after clear()
        REQUIRE(c.size() == 1u);
    }
    // a product-sized adoption (the block comes from the kernel, is recycled by the next adoption, and grows table by
    // table afterwards), with coefficients that spill to the heap in one of every 1000 terms
    {
        const std::size_t n = 300000;
        const unsigned k = 5;
        std::vector<std::uint64_t> keys(n);
        for (std::size_t i = 0; i < n; ++i) {
            const K key{static_cast<std::uint64_t>(i % 100u), static_cast<std::uint64_t>((i / 100u) % 100u),
                        static_cast<std::uint64_t>(i / 10000u)};
            key.to_words(&keys[i], 3);
        }
        std::sort(keys.begin(), keys.end(), [](auto x, auto y) { return (x & 31u) < (y & 31u); });
        std::vector<std::uint64_t> off(33, 0);
        for (auto v : keys) {
            ++off[(v & 31u) + 1u];
        }
        for (int i = 0; i < 32; ++i) {
            off[i + 1] += off[i];
        }
        auto cf = [](std::size_t i) {
            integer c(static_cast<long long>(i % 977u) + 1);
            if (i % 1000u == 0u) {
                c = c * pow(integer(2), 300); // five limbs: heap
            }
            return (i & 1u) ? integer(0) - c : c;
        };
        for (int round = 0; round < 3; ++round) {
            poly_t big;
            big.set_symbol_set(ss);
            big._adopt_grouped(n, k, off.data(),
                               [&](std::size_t i) { return std::pair<K, integer>(K::from_words(&keys[i], 3), cf(i)); },
                               round == 0 ? 1u : 0u);
            REQUIRE(big.size() == n);
            bool good = true;
            for (std::size_t i = 0; i < n; i += 7) {
                const auto f = big.find(K::from_words(&keys[i], 3));
                good = good && f != big.end() && f->second == cf(i);
            }
            REQUIRE(good);
            std::size_t cnt = 0;
            for (const auto &t : big) {
                cnt += t.second.is_zero() ? 0u : 1u;
            }
            REQUIRE(cnt == n);
            // cancel every third term, then add new ones: erasure and growth inside slab-backed tables
            for (std::size_t i = 0; i < n; i += 3) {
                big.add_term<false>(K::from_words(&keys[i], 3), cf(i));
            }
            REQUIRE(big.size() == n - (n + 2u) / 3u);
            for (std::size_t i = 0; i < n; i += 3) {
                big.add_term(K::from_words(&keys[i], 3), integer(1));
            }
            REQUIRE(big.size() == n);
            const poly_t copy(big);
            REQUIRE(copy == big);
        }
    }
}

// The marshalling of the operands for the C ABI (detail::marshal_operands; what polynomial.hpp:828-833 does in the
// reference): the multi-threaded one-round form must give the very arrays the serial two-pass form gives -- same
// order, same limb count, same two's-complement limbs -- for one- to three-limb and wider coefficients of both
// signs, doubles, mixed coefficient types, multi-word keys, single-table and multi-table containers.
template <typename K, typename C>
static polynomial<K, C> marshal_test_poly(const symbol_set &ss, std::size_t n, unsigned log2, std::uint64_t seed, int wide_every)
{
    std::uint64_t st = seed;
    auto rnd = [&st]() {
        st ^= st << 13;
        st ^= st >> 7;
        st ^= st << 17;
        return st;
    };
    const std::size_t nv = ss.size();
    polynomial<K, C> p;
    p.set_symbol_set(ss);
    auto cf = [&](std::size_t i) -> C {
        if constexpr (std::is_same_v<C, integer>) {
            integer c(static_cast<long long>(rnd() >> (24 + i % 30u)) | 1); // at most 40 bits
            if (i % 7u == 3u) {
                c = c * integer(static_cast<long long>(rnd() >> 24) | 1); // up to 80 bits: two limbs
            }
            if (i % 31u == 5u) {
                c = c * pow(integer(2), 100); // up to 180 bits: three limbs
            }
            if (wide_every > 0 && i % static_cast<std::size_t>(wide_every) == 1u) {
                c = c * pow(integer(2), 250); // beyond the three inline limbs: heap
            }
            return (i & 1u) ? integer(0) - c : c;
        } else {
            return static_cast<double>(static_cast<long long>(rnd() >> 12)) * ((i & 1u) ? -0.25 : 0.5) + 1.0;
        }
    };
    std::vector<std::vector<std::uint64_t>> words; // distinct keys: mixed-radix digits of i
    std::vector<std::size_t> order(n);
    words.reserve(n);
    for (std::size_t i = 0; i < n; ++i) {
        std::vector<typename K::value_type> e(nv);
        std::size_t v = i;
        for (std::size_t j = 0; j < nv; ++j) {
            e[j] = static_cast<typename K::value_type>(v % 19u);
            v /= 19u;
        }
        const K key(e);
        std::vector<std::uint64_t> w(K::n_words(nv));
        key.to_words(w.data(), nv);
        words.push_back(std::move(w));
        order[i] = i;
    }
    if (log2 == 0u) {
        p.reserve(n / 3u); // grow a few times on the way
        for (std::size_t i = 0; i < n; ++i) {
            p._insert_unique(K::from_words(words[i].data(), nv), cf(i));
        }
        return p;
    }
    auto tab_of = [&](std::size_t i) {
        std::uint64_t h = 0;
        for (auto w : words[i]) {
            h += w;
        }
        return h & ((std::uint64_t(1) << log2) - 1u);
    };
    std::stable_sort(order.begin(), order.end(), [&](std::size_t a, std::size_t b) { return tab_of(a) < tab_of(b); });
    std::vector<std::uint64_t> offs((std::size_t(1) << log2) + 1u, 0);
    for (std::size_t i = 0; i < n; ++i) {
        ++offs[tab_of(i) + 1u];
    }
    for (std::size_t t = 0; t + 1u < offs.size(); ++t) {
        offs[t + 1u] += offs[t];
    }
    std::vector<C> cfs(n);
    for (std::size_t i = 0; i < n; ++i) {
        cfs[i] = cf(i);
    }
    p._adopt_grouped(n, log2, offs.data(), [&](std::size_t r) {
        return std::pair<K, C>(K::from_words(words[order[r]].data(), nv), cfs[order[r]]);
    });
    return p;
}

template <typename R, typename K, typename C0, typename C1>
static void marshal_case(const polynomial<K, C0> &x, const polynomial<K, C1> &y, unsigned want_lin)
{
    const std::size_t nv = x.get_symbol_set().size();
    const unsigned W = K::n_words(nv);
    std::vector<std::uint64_t> kx, cx, ky, cy, kx1, cx1, ky1, cy1;
    const unsigned lin1 = obake::detail::marshal_operands<R>(x, y, W, nv, kx1, cx1, ky1, cy1, 1); // serial two-pass form
    for (unsigned th : {0u, 2u, 5u}) {
        const unsigned lin = obake::detail::marshal_operands<R>(x, y, W, nv, kx, cx, ky, cy, th);
        REQUIRE(lin == lin1 && (want_lin == 0u || lin == want_lin));
        REQUIRE(kx == kx1 && ky == ky1);
        REQUIRE(cx == cx1 && cy == cy1);
        REQUIRE(kx.size() == x.size() * W && cx.size() == x.size() * lin && cy.size() == y.size() * lin);
    }
    // and the serial form is what iteration gives
    std::size_t i = 0;
    bool same = true;
    for (const auto &[k, c] : x) {
        std::vector<std::uint64_t> w(W);
        k.to_words(w.data(), nv);
        same = same && std::equal(w.begin(), w.end(), kx1.begin() + static_cast<std::ptrdiff_t>(i * W));
        if constexpr (std::is_same_v<C0, R>) {
            same = same && obake::b200::cf_traits<R>::load(&cx1[i * lin1], lin1) == c;
        }
        ++i;
    }
    REQUIRE(same && i == x.size());
}

static void marshal_tests()
{
    using K1 = packed_monomial<std::uint64_t>;
    using K2 = d_packed_monomial<std::int64_t, 2>;
    const symbol_set s4{"a", "b", "c", "d"};
    {
        // integers of one to three limbs, multi-table x single-table
        const auto x = marshal_test_poly<K1, integer>(s4, 60000, 6, 11, 0);
        const auto y = marshal_test_poly<K1, integer>(s4, 45000, 0, 12, 0);
        marshal_case<integer>(x, y, 3);
        marshal_case<integer>(y, x, 3);
        marshal_case<integer>(x, x, 3);
    }
    {
        // one-limb coefficients only: the three-limb stride is narrowed to one
        polynomial<K1, integer> x, y;
        x.set_symbol_set(s4);
        y.set_symbol_set(s4);
        for (std::uint64_t i = 0; i < 40000; ++i) {
            x._insert_unique(K1{i % 19u, (i / 19u) % 19u, (i / 361u) % 19u, i / 6859u}, integer(static_cast<long long>(i) - 20000ll) * integer(3) + integer(1));
            y._insert_unique(K1{i % 17u, (i / 17u) % 17u, (i / 289u) % 17u, i / 4913u}, integer(static_cast<long long>(i % 1000u) + 1));
        }
        marshal_case<integer>(x, y, 1);
    }
    {
        // a few coefficients beyond three limbs (heap-backed integers): the limb count follows the widest one
        const auto x = marshal_test_poly<K1, integer>(s4, 50000, 5, 13, 977);
        const auto y = marshal_test_poly<K1, integer>(s4, 30000, 4, 14, 0);
        marshal_case<integer>(x, y, 0);
        marshal_case<integer>(y, x, 0);
    }
    {
        // doubles, and integer x double (converted to the result type on the host)
        const auto x = marshal_test_poly<K1, double>(s4, 50000, 6, 15, 0);
        const auto y = marshal_test_poly<K1, double>(s4, 50000, 0, 16, 0);
        marshal_case<double>(x, y, 1);
        const auto z = marshal_test_poly<K1, integer>(s4, 40000, 3, 17, 0);
        marshal_case<double>(z, y, 1);
        marshal_case<double>(x, z, 1);
    }
    {
        // two-word keys
        const auto x = marshal_test_poly<K2, integer>(s4, 40000, 5, 18, 0);
        const auto y = marshal_test_poly<K2, integer>(s4, 40000, 0, 19, 0);
        marshal_case<integer>(x, y, 3);
    }
    {
        // small operands stay serial; empty operands
        const auto x = marshal_test_poly<K1, integer>(s4, 100, 2, 20, 0);
        polynomial<K1, integer> e;
        e.set_symbol_set(s4);
        marshal_case<integer>(x, x, 0);
        marshal_case<integer>(x, e, 0);
    }
}

// ---------------------------------------------------------------------------------------------------------
template <typename K, typename C>
static void mul_mt_hm_test()
{
    // test/polynomials_polynomial_00.cpp:189-240 (polynomial_mul_mt_hm_test)
    using poly_t = polynomial<K, C>;
    auto [a, b] = make_polynomials<poly_t>(symbol_set{"a", "b"}, "a", "b");
    REQUIRE(poly_t(3) * poly_t(4) == poly_t(12));
    REQUIRE((a + b) * (a - b) == a * a - b * b);
    REQUIRE((a * a + b * b) * (a + b) * (a - b) == a * a * a * a - b * b * b * b);
    REQUIRE(((a + b) * (a - b)).size() == 2u);
    // empty operands (:1892-1895)
    REQUIRE((poly_t() * a).empty());
    REQUIRE((a * poly_t()).empty());
    // exponent overflow (:216-238)
    using T = typename K::value_type;
    const unsigned sz = K::psize ? K::psize : 2u;
    const auto [lo, hi] = detail::kpack_get_lims<T>(sz);
    poly_t big;
    big.set_symbol_set(symbol_set{"a", "b"});
    big.add_term(K{hi, T(0)}, C(1));
    big.add_term(K{T(0), T(1)}, C(1));
    REQUIRES_THROWS_CONTAINS(big * a, std::overflow_error,
                             "An overflow in the monomial exponents was detected while attempting to multiply two polynomials");
    if constexpr (std::is_signed_v<T>) {
        poly_t small;
        small.set_symbol_set(symbol_set{"a", "b"});
        small.add_term(K{lo, T(0)}, C(1));
        poly_t inv;
        inv.set_symbol_set(symbol_set{"a", "b"});
        inv.add_term(K{T(-1), T(0)}, C(1));
        REQUIRES_THROWS_CONTAINS(small * inv, std::overflow_error,
                                 "An overflow in the monomial exponents was detected while attempting to multiply two polynomials");
    }
}

template <typename K>
static void mul_general_test()
{
    // test/polynomials_polynomial_00.cpp:244-394: symbol-set merging in all three cases, *=, mixed coefficients
    using p_int = polynomial<K, integer>;
    using p_dbl = polynomial<K, double>;
    auto [x] = make_polynomials<p_int>("x");
    auto [y] = make_polynomials<p_int>("y");
    auto [xx, yy, zz] = make_polynomials<p_int>(symbol_set{"x", "y", "z"}, "x", "y", "z");
    auto r = x * y; // both extended
    REQUIRE(r.get_symbol_set() == symbol_set{"x", "y"});
    REQUIRE(r.size() == 1u);
    auto r2 = xx * y; // only the second extended
    REQUIRE(r2.get_symbol_set() == symbol_set{"x", "y", "z"});
    REQUIRE(r2 == xx * yy);
    auto r3 = y * zz; // only the first extended
    REQUIRE(r3 == yy * zz);
    auto acc = x + 1;
    acc *= x - 1;
    REQUIRE(acc == x * x - 1);
    // integer x double -> double
    auto [xd] = make_polynomials<p_dbl>("x");
    auto m = (x * 3 + 1) * (xd * 0.5 + 2.);
    static_assert(std::is_same_v<decltype(m), p_dbl>);
    REQUIRE(m == xd * xd * 1.5 + xd * 6.5 + 2.);
}

template <typename K, typename C>
static void truncated_tests()
{
    // test/polynomials_polynomial_01.cpp:131-183
    using poly_t = polynomial<K, C>;
    auto [x, y, z] = make_polynomials<poly_t>(symbol_set{"x", "y", "z"}, "x", "y", "z");
    REQUIRE(truncated_mul(x + y, x - y, 100) == x * x - y * y);
    REQUIRE(truncated_mul(x + y, x - y, 2) == x * x - y * y);
    REQUIRE(truncated_mul(x + y, x - y, 1).empty());
    REQUIRE(truncated_mul(x + y, x - y, 0).empty());
    REQUIRE(truncated_mul(x + y, x - y, -1).empty() || !std::is_signed_v<typename K::value_type>);
    REQUIRE(truncated_mul(z * x + y, x - y - 1, 100) == z * x * x - z * x * y - z * x + y * x - y * y - y);
    REQUIRE(truncated_mul(z * x + y, x - y - 1, 2) == -z * x + y * x - y * y - y);
    REQUIRE(truncated_mul(z * x + y, x - y - 1, integer(2)) == -z * x + y * x - y * y - y);
    REQUIRE(truncated_mul(z * x + y, x - y - 1, 1) == -y);
    REQUIRE(truncated_mul(z * x + y, x - y - 1, integer(1)) == -y);
    REQUIRE(truncated_mul(z * x + y, x - y - 1, 0).empty());
    REQUIRE(truncated_mul(z * x + y, x - y - 1, integer(-1)).empty());
    // partial degree (test/polynomials_polynomial_02.cpp:182-252)
    REQUIRE(truncated_mul(z * x + y, x - y - 1, 100, symbol_set{"x"}) == z * x * x - z * x * y - z * x + y * x - y * y - y);
    REQUIRE(truncated_mul(z * x + y, x - y - 1, 1, symbol_set{"x"}) == -z * x * y - z * x + y * x - y * y - y);
    REQUIRE(truncated_mul(z * x + y, x - y - 1, 0, symbol_set{"x"}) == -y * y - y);
    REQUIRE(truncated_mul(z * x + y, x - y - 1, 0, symbol_set{"x", "y"}).empty());
    REQUIRE(truncated_mul(z * x + y, x - y - 1, 1, symbol_set{"y", "zzz"}) == z * x * x - z * x * y - z * x + y * x - y);
}

template <typename K, typename C>
static void large_tests()
{
    using poly_t = polynomial<K, C>;
    auto [t, u, x, y, z] = make_polynomials<poly_t>(symbol_set{"t", "u", "x", "y", "z"}, "t", "u", "x", "y", "z");
    auto f = x + y + z * z * 2 + t * t * t * 3 + u * u * u * u * u * 5 + 1;
    auto g = u + t + z * z * 2 + y * y * y * 3 + x * x * x * x * x * 5 + 1;
    const auto tf = f, tg = g;
    // n = 8: truncation cases of test/polynomials_polynomial_01.cpp:185-236 and _04.cpp:239-262
    for (int i = 1; i < 8; ++i) {
        f *= tf;
        g *= tg;
    }
    const auto full = f * g;
    REQUIRE(full.size() == 591235u);
    REQUIRE(truncated_mul(f, g, 1000) == full);
    REQUIRE(truncated_mul(f, g, 80) == full);
    REQUIRE(degree(truncated_mul(f, g, 40)) == 40);
    REQUIRE(degree(truncated_mul(f, g, 5)) == 5);
    REQUIRE(truncated_mul(f, g, 0) == poly_t(1));
    if constexpr (std::is_signed_v<typename K::value_type>) {
        REQUIRE(truncated_mul(f, g, -1).empty());
    }
    REQUIRE(truncated_mul(f, g, 50) == truncate_degree(full, 50));
    REQUIRE(truncated_mul(f, g, 20, symbol_set{"x", "z"}) == truncate_p_degree(full, 20, symbol_set{"x", "z"}));
    // n = 10: 2 096 600 terms (test/polynomials_polynomial_00.cpp:398-424)
    for (int i = 8; i < 10; ++i) {
        f *= tf;
        g *= tg;
    }
    REQUIRE(f.size() == 3003u);
    REQUIRE((f * g).size() == 2096600u);
    // (..)^20 by 19 rectangular products: 53 130 terms (test/polynomials_polynomial_04.cpp:264-285)
    auto h = tf;
    for (int i = 0; i < 19; ++i) {
        h *= tf;
    }
    REQUIRE(h.size() == 53130u);
}

static void fateman_test()
{
    // benchmark/dense.hpp:21-44 at n = 12, checked against (1+..)^(2n) + (1+..)^n coefficient by coefficient
    using poly_t = polynomial<packed_monomial<std::uint64_t>, integer>;
    auto [t, x, y, z] = make_polynomials<poly_t>(symbol_set{"t", "x", "y", "z"}, "t", "x", "y", "z");
    auto f = x + y + z + t + 1;
    const auto tmp = f;
    for (int i = 1; i < 12; ++i) {
        f *= tmp;
    }
    auto g = f + 1;
    const auto ret = f * g;
    REQUIRE(ret.size() == 20475u); // C(28,4)
    auto f24 = f * f;              // (1+..)^24
    REQUIRE(ret == f24 + f);
    // sum of the coefficients = 5^24 + 5^12
    integer sum(0);
    for (const auto &[k, c] : ret) {
        sum += c;
    }
    REQUIRE(sum == pow(integer(5), 24) + pow(integer(5), 12));
}

// Device-resident chains: the operand-building loops of the reference's benchmarks (f *= tmp,
// benchmark/dense.hpp:28-32; g = f + 1, :34; truncated products, audi_01.cpp:32-42) with every intermediate in HBM,
// compared with the same chain through the host container.
template <typename K, typename C>
static void device_chain_test()
{
    using poly_t = polynomial<K, C>;
    using dpoly_t = b200::device_polynomial<K, C>;
    const symbol_set ss{"t", "x", "y", "z"};
    auto [t, x, y, z] = make_polynomials<poly_t>(ss, "t", "x", "y", "z");
    const poly_t tmp = x + y * 2 - z + t * 3 + 1;
    poly_t f = tmp;
    dpoly_t df(tmp);
    const dpoly_t dtmp(tmp);
    REQUIRE(df.size() == tmp.size());
    REQUIRE(df.to_host() == tmp);
    for (int i = 1; i < 7; ++i) {
        f *= tmp;
        df = df * dtmp; // never leaves the GPU
    }
    REQUIRE(df.size() == f.size());
    REQUIRE(df.to_host() == f);
    // g = f + tmp, h = f - f, products and truncation
    const poly_t g = f + tmp;
    const dpoly_t dg = df + dtmp;
    REQUIRE(dg.to_host() == g);
    REQUIRE((df - df).empty());
    REQUIRE((dg - dtmp).to_host() == f);
    const poly_t full = f * g;
    const dpoly_t dfull = df * dg;
    REQUIRE(dfull.to_host() == full);
    REQUIRE(truncated_mul(df, dg, 9).to_host() == truncated_mul(f, g, 9));
    REQUIRE(truncated_mul(df, dg, 5, symbol_set{"x", "z"}).to_host() == truncated_mul(f, g, 5, symbol_set{"x", "z"}));
    REQUIRE(truncate_degree(dfull, 8).to_host() == truncate_degree(full, 8));
    REQUIRE(truncate_p_degree(dfull, 3, symbol_set{"t", "y"}).to_host() == truncate_p_degree(full, 3, symbol_set{"t", "y"}));
    REQUIRE(truncate_degree(dfull, -1).to_host() == truncate_degree(full, -1));
    // empty operands and mismatched symbol sets
    poly_t e0;
    e0.set_symbol_set(ss);
    const dpoly_t de(e0);
    REQUIRE((df * de).empty());
    REQUIRE((df + de).to_host() == f);
    // different symbol sets: the keys are re-packed over the union on the device (series.hpp:539-625)
    auto [a, y2] = make_polynomials<poly_t>(symbol_set{"a", "y"}, "a", "y");
    const poly_t q = a * 3 + y2 * y2 - 2;
    const dpoly_t dq(q);
    REQUIRE((df * dq).get_symbol_set() == symbol_set{"a", "t", "x", "y", "z"});
    REQUIRE((df * dq).to_host() == f * q);
    REQUIRE((dq + df).to_host() == q + f);
    REQUIRE((dq - df).to_host() == q - f);
    REQUIRE(truncated_mul(df, dq, 6, symbol_set{"a", "x"}).to_host() == truncated_mul(f, q, 6, symbol_set{"a", "x"}));
}

// test/power_series_01.cpp:384-560 ("multiplication"): truncation policies of p_series products
static void p_series_tests()
{
    using pm_t = packed_monomial<std::int32_t>;
    using ps_t = p_series<pm_t, double>;
    {
        auto [x, y] = make_p_series<ps_t>("x", "y");
        auto ret = x * y;
        REQUIRE(ret.get_symbol_set() == symbol_set{"x", "y"});
        REQUIRE(ret.size() == 1u);
        REQUIRE(get_truncation(ret).index() == 0u);
        REQUIRE(ret.begin()->first == pm_t{1, 1});
        REQUIRE(ret.begin()->second == 1);
    }
    {
        auto [x, y] = make_p_series_t<ps_t>(3, "x", "y");
        auto ret = x * y;
        REQUIRE(ret.size() == 1u);
        REQUIRE(get_truncation(ret).index() == 1u);
        REQUIRE(std::get<1>(get_truncation(ret)) == 3);
        REQUIRE(ret.begin()->first == pm_t{1, 1});
    }
    {
        auto [x, y] = make_p_series_t<ps_t>(1, "x", "y");
        auto ret = x * y;
        REQUIRE(ret.get_symbol_set() == symbol_set{"x", "y"});
        REQUIRE(ret.empty());
        REQUIRE(std::get<1>(get_truncation(ret)) == 1);
    }
    {
        auto [x, y] = make_p_series_p<ps_t>(3, symbol_set{"a", "b"}, "x", "y");
        auto ret = x * y;
        REQUIRE(ret.size() == 1u);
        REQUIRE(get_truncation(ret).index() == 2u);
        REQUIRE(std::get<2>(get_truncation(ret)) == std::pair{std::int32_t(3), symbol_set{"a", "b"}});
    }
    {
        auto [x, y] = make_p_series_p<ps_t>(1, symbol_set{"x", "y", "z"}, "x", "y");
        auto ret = x * y;
        REQUIRE(ret.empty());
        REQUIRE(get_truncation(ret).index() == 2u);
    }
    {
        auto [x] = make_p_series_t<ps_t>(3, "x");
        auto [y] = make_p_series_t<ps_t>(2, "y");
        REQUIRES_THROWS_CONTAINS(x * y, std::invalid_argument,
                                 "Unable to multiply two power series if their truncation levels do not match");
    }
    {
        auto [x] = make_p_series_p<ps_t>(3, symbol_set{"a", "b"}, "x");
        auto [y] = make_p_series_p<ps_t>(3, symbol_set{"a", "c"}, "y");
        REQUIRES_THROWS_CONTAINS(x * y, std::invalid_argument,
                                 "Unable to multiply two power series if their truncation levels do not match");
    }
    {
        auto [x] = make_p_series_t<ps_t>(3, "x");
        auto [y] = make_p_series_p<ps_t>(2, symbol_set{"a"}, "y");
        REQUIRES_THROWS_CONTAINS(x * y, std::invalid_argument,
                                 "Unable to multiply two power series if their truncation policies do not match");
    }
    {
        auto [x, y] = make_p_series_t<ps_t>(3, "x", "y");
        unset_truncation(y);
        auto ret = x * y;
        REQUIRE(ret.size() == 1u);
        REQUIRE(std::get<1>(get_truncation(ret)) == 3);
    }
    {
        auto [x, y] = make_p_series_t<ps_t>(1, "x", "y");
        unset_truncation(x);
        auto ret = x * y;
        REQUIRE(ret.empty());
        REQUIRE(std::get<1>(get_truncation(ret)) == 1);
    }
    // a truncated power: (1 + x + y)^8 kept to total degree 4 at every step == truncate_degree of the full power
    {
        using pi_t = p_series<packed_monomial<std::int64_t>, integer>;
        auto [x, y] = make_p_series_t<pi_t>(symbol_set{"x", "y"}, 4, "x", "y");
        pi_t one(polynomial<packed_monomial<std::int64_t>, integer>(1));
        one.poly() = x.poly() * 0 + 1; // constant over {x, y}
        pi_t b = x + y + one;
        pi_t acc = b;
        auto full = b.poly();
        for (int i = 1; i < 8; ++i) {
            acc *= b;
            full = full * b.poly();
        }
        REQUIRE(acc.poly() == truncate_degree(full, 4));
        REQUIRE(std::get<1>(get_truncation(acc)) == 4);
    }
}

// obake::pow (series pow, series.hpp:1628-1712: the previous power times the base) as a device-resident chain, and the
// truncated power loop of benchmark/audi_01.cpp:32-42; both against the same chains through the host container.
template <typename K, typename C>
static void pow_test()
{
    using poly_t = polynomial<K, C>;
    using dpoly_t = b200::device_polynomial<K, C>;
    const symbol_set ss{"t", "x", "y", "z"};
    auto [t, x, y, z] = make_polynomials<poly_t>(ss, "t", "x", "y", "z");
    const poly_t base = x - y * 3 + z * 2 + t + 1;
    REQUIRE(pow(base, 0) == poly_t(1));
    REQUIRE(pow(base, 1) == base);
    poly_t h = base;
    for (unsigned n = 2; n <= 9; ++n) {
        h = h * base;
        REQUIRE(pow(base, n) == h);
    }
    poly_t empty;
    empty.set_symbol_set(ss);
    REQUIRE(pow(empty, 3).empty());
    // truncated powers
    const dpoly_t dbase(base);
    poly_t ht = base;
    for (unsigned n = 2; n <= 8; ++n) {
        ht = truncated_mul(ht, base, 5);
        REQUIRE(b200::truncated_pow(dbase, n, 5).to_host() == ht);
    }
    REQUIRE(b200::pow(dbase, 6).to_host() == pow(base, 6));
}

int main(int argc, char **argv)
{
    const bool cpu_only = argc > 1 && std::strcmp(argv[1], "--cpu") == 0;
    if (argc > 1 && std::strcmp(argv[1], "--multi") == 0) {
        // OBAKE_B200_DEVICES=0,1,...: the host-container paths through the single-process multi-device engine
        const unsigned nd = b200::engine::instance().n_devices();
        mul_mt_hm_test<packed_monomial<std::int64_t>, integer>();
        mul_mt_hm_test<packed_monomial<std::int64_t>, double>();
        mul_mt_hm_test<d_packed_monomial<std::int64_t, 8>, integer>();
        mul_general_test<packed_monomial<std::int64_t>>();
        truncated_tests<packed_monomial<std::int64_t>, integer>();
        truncated_tests<d_packed_monomial<std::int64_t, 2>, integer>();
        large_tests<packed_monomial<std::int64_t>, integer>();
        fateman_test();
        p_series_tests();
        std::printf("%s: %d checks, %d failures (%u devices)\n", g_fail ? "FAILED" : "OK", g_checks, g_fail, nd);
        return g_fail || nd < 2u ? 1 : 0;
    }
    kpack_tests<std::int32_t>();
    kpack_tests<std::uint32_t>();
    kpack_tests<std::int64_t>();
    kpack_tests<std::uint64_t>();
    monomial_tests<std::int64_t>();
    monomial_tests<std::uint64_t>();
    monomial_tests<std::int32_t>();
    integer_tests();
    seg_table_tests();
    marshal_tests();
    if (!cpu_only) {
        mul_mt_hm_test<packed_monomial<std::int64_t>, integer>();
        mul_mt_hm_test<packed_monomial<std::int64_t>, double>();
        mul_mt_hm_test<packed_monomial<std::uint64_t>, integer>();
        mul_mt_hm_test<packed_monomial<std::int32_t>, integer>();
        mul_mt_hm_test<d_packed_monomial<std::int64_t, 8>, integer>();
        mul_mt_hm_test<d_packed_monomial<std::uint64_t, 1>, double>();
        mul_general_test<packed_monomial<std::int64_t>>();
        mul_general_test<d_packed_monomial<std::int64_t, 2>>();
        truncated_tests<packed_monomial<std::int64_t>, integer>();
        truncated_tests<packed_monomial<std::int64_t>, double>();
        truncated_tests<d_packed_monomial<std::int64_t, 2>, integer>();
        large_tests<packed_monomial<std::int64_t>, integer>();
        large_tests<d_packed_monomial<std::int64_t, 3>, double>();
        fateman_test();
        p_series_tests();
        device_chain_test<packed_monomial<std::uint64_t>, integer>();
        device_chain_test<packed_monomial<std::int64_t>, double>();
        device_chain_test<d_packed_monomial<std::int64_t, 2>, integer>();
        pow_test<packed_monomial<std::uint64_t>, integer>();
        pow_test<packed_monomial<std::int64_t>, double>();
        pow_test<d_packed_monomial<std::int64_t, 2>, integer>();
    }
    std::printf("%s: %d checks, %d failures%s\n", g_fail ? "FAILED" : "OK", g_checks, g_fail, cpu_only ? " (host-only subset)" : "");
    return g_fail ? 1 : 0;
}
