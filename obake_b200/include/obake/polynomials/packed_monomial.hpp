// obake/polynomials/packed_monomial.hpp -- one-integer Kronecker-packed monomial (host side).
// Same surface as the reference's include/obake/polynomials/packed_monomial.hpp for the series-product
// path: construction (:56-142), hash == value (:172-176), key_is_compatible (:179-204), monomial_mul ==
// integer addition (:249-262), monomial_range_overflow_check (:289-488), key_degree / key_p_degree
// (src/polynomials/packed_monomial.cpp:334-350, 388-412).
#ifndef OBAKE_B200_PACKED_MONOMIAL_HPP
#define OBAKE_B200_PACKED_MONOMIAL_HPP

#include <cstddef>
#include <cstdint>
#include <initializer_list>
#include <stdexcept>
#include <type_traits>
#include <vector>

#include "../kpack.hpp"
#include "../symbols.hpp"

namespace obake
{

template <typename T>
class packed_monomial
{
    static_assert(detail::is_kpackable_v<T>);
    T m_value = 0;

public:
    using value_type = T;
    static constexpr unsigned psize = 0; // tag used by the engine marshalling: one word packs every exponent
    packed_monomial() = default;
    explicit packed_monomial(const symbol_set &) : m_value(0) {}
    explicit packed_monomial(const T &n) : m_value(n) {}
    template <typename It>
    explicit packed_monomial(It it, std::size_t n)
    {
        kpacker<T> kp(static_cast<unsigned>(n));
        for (std::size_t i = 0; i < n; ++i, ++it) {
            kp << static_cast<T>(*it);
        }
        m_value = kp.get();
    }
    packed_monomial(std::initializer_list<T> l) : packed_monomial(l.begin(), l.size()) {}
    explicit packed_monomial(const std::vector<T> &v) : packed_monomial(v.begin(), v.size()) {}

    const T &get_value() const
    {
        return m_value;
    }
    void _set_value(const T &n)
    {
        m_value = n;
    }
    std::vector<T> unpack(std::size_t nvars) const
    {
        std::vector<T> r(nvars);
        kunpacker<T> ku(m_value, static_cast<unsigned>(nvars));
        for (auto &e : r) {
            ku >> e;
        }
        return r;
    }
    // engine marshalling: the key as 64-bit words (32-bit types are sign/zero extended)
    static constexpr unsigned n_words(std::size_t)
    {
        return 1u;
    }
    void to_words(std::uint64_t *out, std::size_t) const
    {
        if constexpr (std::is_signed_v<T>) {
            out[0] = static_cast<std::uint64_t>(static_cast<std::int64_t>(m_value));
        } else {
            out[0] = static_cast<std::uint64_t>(m_value);
        }
    }
    static packed_monomial from_words(const std::uint64_t *in, std::size_t)
    {
        return packed_monomial(static_cast<T>(in[0]));
    }
    friend bool operator==(const packed_monomial &a, const packed_monomial &b)
    {
        return a.m_value == b.m_value;
    }
    friend bool operator!=(const packed_monomial &a, const packed_monomial &b)
    {
        return a.m_value != b.m_value;
    }
};

template <typename T>
inline std::size_t hash(const packed_monomial<T> &m)
{
    return static_cast<std::size_t>(m.get_value());
}

template <typename T>
inline bool key_is_compatible(const packed_monomial<T> &m, const symbol_set &s)
{
    const auto n = s.size();
    if (n == 0u) {
        return m.get_value() == T(0);
    }
    if (n > detail::kpack_max_size<T>()) {
        return false;
    }
    const auto [kmin, kmax] = detail::kpack_get_klims<T>(static_cast<unsigned>(n));
    return m.get_value() >= kmin && m.get_value() <= kmax;
}

template <typename T>
inline void monomial_mul(packed_monomial<T> &out, const packed_monomial<T> &a, const packed_monomial<T> &b,
                         const symbol_set &)
{
    using U = std::make_unsigned_t<T>;
    out._set_value(static_cast<T>(static_cast<U>(a.get_value()) + static_cast<U>(b.get_value())));
}

template <typename T>
inline T key_degree(const packed_monomial<T> &m, const symbol_set &ss)
{
    T d(0);
    for (const auto &e : m.unpack(ss.size())) {
        d = static_cast<T>(d + e);
    }
    return d;
}

template <typename T>
inline T key_p_degree(const packed_monomial<T> &m, const std::vector<unsigned> &idx, const symbol_set &ss)
{
    const auto e = m.unpack(ss.size());
    T d(0);
    for (auto i : idx) {
        d = static_cast<T>(d + e[i]);
    }
    return d;
}

// Host restatement of the interval test (used for tiny inputs and by the tests; the engine runs the same
// test on the device before every product).
template <typename R1, typename R2>
inline bool monomial_range_overflow_check_impl(const R1 &r1, const R2 &r2, const symbol_set &ss, unsigned wsize,
                                               bool signed_t, long double lim_min, long double lim_max)
{
    (void)wsize;
    (void)signed_t;
    const auto n = ss.size();
    if (n == 0u || r1.empty() || r2.empty()) {
        return true;
    }
    std::vector<long double> mn1(n), mx1(n), mn2(n), mx2(n);
    auto scan = [n](const auto &r, auto &mn, auto &mx) {
        bool first = true;
        for (const auto &m : r) {
            const auto e = m.unpack(n);
            for (std::size_t j = 0; j < n; ++j) {
                const auto v = static_cast<long double>(e[j]);
                if (first || v < mn[j]) mn[j] = v;
                if (first || v > mx[j]) mx[j] = v;
            }
            first = false;
        }
    };
    scan(r1, mn1, mx1);
    scan(r2, mn2, mx2);
    for (std::size_t j = 0; j < n; ++j) {
        if (mn1[j] + mn2[j] < lim_min || mx1[j] + mx2[j] > lim_max) {
            return false;
        }
    }
    return true;
}

template <typename T>
inline bool monomial_range_overflow_check(const std::vector<packed_monomial<T>> &r1,
                                          const std::vector<packed_monomial<T>> &r2, const symbol_set &ss)
{
    if (ss.size() == 0u) {
        return true;
    }
    const auto [lmin, lmax] = detail::kpack_get_lims<T>(static_cast<unsigned>(ss.size()));
    return monomial_range_overflow_check_impl(r1, r2, ss, static_cast<unsigned>(ss.size()), std::is_signed_v<T>,
                                              static_cast<long double>(lmin), static_cast<long double>(lmax));
}

namespace polynomials
{
using p_monomial_value_t = std::uint64_t;
}
// The reference's default packed monomial (packed_monomial.hpp:767-773).
using p_monomial = packed_monomial<std::uint64_t>;

} // namespace obake

#endif
