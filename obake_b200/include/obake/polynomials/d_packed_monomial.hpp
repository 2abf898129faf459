// obake/polynomials/d_packed_monomial.hpp -- vector-of-packed-words monomial (host side).
// Same surface as the reference's include/obake/polynomials/d_packed_monomial.hpp for the series-product
// path: layout ceil(nvars/PSize) words, each a kpack of PSize exponents, last word zero padded (:97-132);
// hash = sum of the words (:274-284); monomial_mul = word-wise addition (:532-549); key_degree (:901-920),
// key_p_degree (:930-958); overflow check incl. the degree range (:583-897).
#ifndef OBAKE_B200_D_PACKED_MONOMIAL_HPP
#define OBAKE_B200_D_PACKED_MONOMIAL_HPP

#include <cstddef>
#include <cstdint>
#include <initializer_list>
#include <type_traits>
#include <vector>

#include "../kpack.hpp"
#include "../symbols.hpp"
#include "packed_monomial.hpp"

namespace obake
{

template <typename T, unsigned PSize>
class d_packed_monomial
{
    static_assert(detail::is_kpackable_v<T>);
    static_assert(PSize > 0u);
    std::vector<T> m_container;

public:
    using value_type = T;
    static constexpr unsigned psize = PSize;
    static std::size_t n_expos_to_vsize(std::size_t n)
    {
        return n / PSize + static_cast<std::size_t>(n % PSize != 0u);
    }
    d_packed_monomial() = default;
    explicit d_packed_monomial(const symbol_set &ss) : m_container(n_expos_to_vsize(ss.size()), T(0)) {}
    template <typename It>
    explicit d_packed_monomial(It it, std::size_t n) : m_container(n_expos_to_vsize(n))
    {
        std::size_t counter = 0;
        for (auto &out : m_container) {
            kpacker<T> kp(PSize);
            for (unsigned j = 0; j < PSize && counter < n; ++j, ++counter, ++it) {
                kp << static_cast<T>(*it);
            }
            out = kp.get();
        }
    }
    d_packed_monomial(std::initializer_list<T> l) : d_packed_monomial(l.begin(), l.size()) {}
    explicit d_packed_monomial(const std::vector<T> &v) : d_packed_monomial(v.begin(), v.size()) {}

    const std::vector<T> &_container() const
    {
        return m_container;
    }
    std::vector<T> &_container()
    {
        return m_container;
    }
    std::vector<T> unpack(std::size_t nvars) const
    {
        std::vector<T> r(nvars);
        std::size_t idx = 0;
        for (const auto &n : m_container) {
            kunpacker<T> ku(n, PSize);
            for (unsigned j = 0; j < PSize && idx < nvars; ++j, ++idx) {
                ku >> r[idx];
            }
        }
        return r;
    }
    static unsigned n_words(std::size_t nvars)
    {
        const auto n = n_expos_to_vsize(nvars);
        return static_cast<unsigned>(n == 0u ? 1u : n);
    }
    void to_words(std::uint64_t *out, std::size_t nvars) const
    {
        const unsigned W = n_words(nvars);
        for (unsigned w = 0; w < W; ++w) {
            const T v = w < m_container.size() ? m_container[w] : T(0);
            if constexpr (std::is_signed_v<T>) {
                out[w] = static_cast<std::uint64_t>(static_cast<std::int64_t>(v));
            } else {
                out[w] = static_cast<std::uint64_t>(v);
            }
        }
    }
    static d_packed_monomial from_words(const std::uint64_t *in, std::size_t nvars)
    {
        d_packed_monomial r;
        r.m_container.resize(n_expos_to_vsize(nvars));
        for (std::size_t w = 0; w < r.m_container.size(); ++w) {
            r.m_container[w] = static_cast<T>(in[w]);
        }
        return r;
    }
    friend bool operator==(const d_packed_monomial &a, const d_packed_monomial &b)
    {
        return a.m_container == b.m_container;
    }
    friend bool operator!=(const d_packed_monomial &a, const d_packed_monomial &b)
    {
        return !(a == b);
    }
};

template <typename T, unsigned P>
inline std::size_t hash(const d_packed_monomial<T, P> &d)
{
    std::size_t ret = 0;
    for (const auto &n : d._container()) {
        ret += static_cast<std::size_t>(n);
    }
    return ret;
}

template <typename T, unsigned P>
inline bool key_is_compatible(const d_packed_monomial<T, P> &d, const symbol_set &s)
{
    if (d._container().size() != d_packed_monomial<T, P>::n_expos_to_vsize(s.size())) {
        return false;
    }
    const auto [kmin, kmax] = detail::kpack_get_klims<T>(P);
    for (const auto &n : d._container()) {
        if (n < kmin || n > kmax) {
            return false;
        }
    }
    // the padding exponents of the last word must be zero
    const auto e = d.unpack(d._container().size() * P);
    for (std::size_t i = s.size(); i < e.size(); ++i) {
        if (e[i] != T(0)) {
            return false;
        }
    }
    return true;
}

template <typename T, unsigned P>
inline void monomial_mul(d_packed_monomial<T, P> &out, const d_packed_monomial<T, P> &a, const d_packed_monomial<T, P> &b,
                         const symbol_set &)
{
    using U = std::make_unsigned_t<T>;
    out._container().resize(a._container().size());
    for (std::size_t i = 0; i < a._container().size(); ++i) {
        out._container()[i] = static_cast<T>(static_cast<U>(a._container()[i]) + static_cast<U>(b._container()[i]));
    }
}

template <typename T, unsigned P>
inline T key_degree(const d_packed_monomial<T, P> &d, const symbol_set &ss)
{
    T r(0);
    for (const auto &e : d.unpack(ss.size())) {
        r = static_cast<T>(r + e);
    }
    return r;
}

template <typename T, unsigned P>
inline T key_p_degree(const d_packed_monomial<T, P> &d, const std::vector<unsigned> &idx, const symbol_set &ss)
{
    const auto e = d.unpack(ss.size());
    T r(0);
    for (auto i : idx) {
        r = static_cast<T>(r + e[i]);
    }
    return r;
}

template <typename T, unsigned P>
inline bool monomial_range_overflow_check(const std::vector<d_packed_monomial<T, P>> &r1,
                                          const std::vector<d_packed_monomial<T, P>> &r2, const symbol_set &ss)
{
    if (ss.size() == 0u) {
        return true;
    }
    const auto [lmin, lmax] = detail::kpack_get_lims<T>(P);
    return monomial_range_overflow_check_impl(r1, r2, ss, P, std::is_signed_v<T>, static_cast<long double>(lmin),
                                              static_cast<long double>(lmax));
}

// The reference's default dynamic packed monomial: d_monomial = d_packed_monomial<uint64_t, 8>
// (d_packed_monomial.hpp:220-226, 1405).
using d_monomial = d_packed_monomial<std::uint64_t, 8>;

} // namespace obake

#endif
