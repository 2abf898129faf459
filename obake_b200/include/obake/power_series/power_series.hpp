// obake/power_series/power_series.hpp -- p_series<K, C>: a polynomial with a truncation policy whose product runs
// on the B200 engine.
//
// Mirrors the multiplication rules of the reference (include/obake/power_series/power_series.hpp:1284-1357):
// the truncation state is one of {none, total degree d, partial degree (d, symbols)}; two series multiply when their
// policies match (and then their levels must match too) or when one of them has no truncation; the product is
// polynomials::detail::poly_mul_impl_switch(ps0, ps1[, d[, symbols]]) -- i.e. the engine's truncated product with
// the filter fused into the key-add kernel -- and keeps the truncation tag.  set_truncation / unset_truncation /
// get_truncation / truncate follow :463-560, make_p_series / make_p_series_t / make_p_series_p follow :567-891.
// Only the product path is on the GPU; everything else here is host bookkeeping around `polynomial`.
#ifndef OBAKE_POWER_SERIES_POWER_SERIES_HPP
#define OBAKE_POWER_SERIES_POWER_SERIES_HPP

#include <array>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <variant>

#include "../polynomials/polynomial.hpp"

namespace obake
{

namespace power_series
{
struct no_truncation {
    friend bool operator==(const no_truncation &, const no_truncation &)
    {
        return true;
    }
    friend bool operator!=(const no_truncation &, const no_truncation &)
    {
        return false;
    }
};
} // namespace power_series

template <typename K, typename C>
class p_series
{
public:
    using poly_t = polynomial<K, C>;
    using key_type = K;
    using cf_type = C;
    using deg_t = typename K::value_type; // psk_deg_t: the key's (partial) degree type
    using trunc_t = std::variant<power_series::no_truncation, deg_t, std::pair<deg_t, symbol_set>>;

private:
    poly_t m_poly;
    trunc_t m_trunc;

public:
    p_series() = default;
    explicit p_series(poly_t p) : m_poly(std::move(p)) {}
    p_series(poly_t p, trunc_t t) : m_poly(std::move(p)), m_trunc(std::move(t)) {}
    const poly_t &poly() const
    {
        return m_poly;
    }
    poly_t &poly()
    {
        return m_poly;
    }
    const trunc_t &trunc() const
    {
        return m_trunc;
    }
    trunc_t &trunc()
    {
        return m_trunc;
    }
    const symbol_set &get_symbol_set() const
    {
        return m_poly.get_symbol_set();
    }
    std::size_t size() const
    {
        return m_poly.size();
    }
    bool empty() const
    {
        return m_poly.empty();
    }
    auto begin() const
    {
        return m_poly.begin();
    }
    auto end() const
    {
        return m_poly.end();
    }
    friend bool operator==(const p_series &a, const p_series &b)
    {
        return a.m_trunc == b.m_trunc && a.m_poly == b.m_poly;
    }
};

// get_truncation (:542-546)
template <typename K, typename C>
inline const typename p_series<K, C>::trunc_t &get_truncation(const p_series<K, C> &ps)
{
    return ps.trunc();
}

// truncate (:548-560): apply the stored policy to the terms
template <typename K, typename C>
inline void truncate(p_series<K, C> &ps)
{
    using ps_t = p_series<K, C>;
    std::visit(
        [&ps](const auto &v) {
            using type = std::decay_t<decltype(v)>;
            if constexpr (std::is_same_v<type, typename ps_t::deg_t>) {
                ps.poly() = truncate_degree(ps.poly(), v);
            } else if constexpr (!std::is_same_v<type, power_series::no_truncation>) {
                ps.poly() = truncate_p_degree(ps.poly(), v.first, v.second);
            }
        },
        ps.trunc());
}

// set_truncation (:463-532): store the policy, then truncate
template <typename K, typename C, typename T>
inline p_series<K, C> &set_truncation(p_series<K, C> &ps, const T &d)
{
    ps.trunc() = static_cast<typename p_series<K, C>::deg_t>(d);
    truncate(ps);
    return ps;
}
template <typename K, typename C, typename T>
inline p_series<K, C> &set_truncation(p_series<K, C> &ps, const T &d, symbol_set ss)
{
    ps.trunc() = std::make_pair(static_cast<typename p_series<K, C>::deg_t>(d), std::move(ss));
    truncate(ps);
    return ps;
}
// unset_truncation (:535-539)
template <typename K, typename C>
inline p_series<K, C> &unset_truncation(p_series<K, C> &ps)
{
    ps.trunc() = power_series::no_truncation{};
    return ps;
}

// series_mul(p_series, p_series) (:1284-1357)
template <typename K, typename C0, typename C1>
inline auto series_mul(const p_series<K, C0> &ps0, const p_series<K, C1> &ps1)
{
    using R = decltype(std::declval<C0>() * std::declval<C1>());
    using ret_t = p_series<K, R>;
    using deg_t = typename ret_t::deg_t;
    using nt = power_series::no_truncation;
    auto run = [&](const auto &policy) -> ret_t {
        using type = std::decay_t<decltype(policy)>;
        if constexpr (std::is_same_v<type, nt>) {
            return ret_t(polynomials::series_mul(ps0.poly(), ps1.poly()));
        } else if constexpr (std::is_same_v<type, deg_t>) {
            return ret_t(polynomials::truncated_mul(ps0.poly(), ps1.poly(), policy), typename ret_t::trunc_t(policy));
        } else {
            return ret_t(polynomials::truncated_mul(ps0.poly(), ps1.poly(), policy.first, policy.second),
                         typename ret_t::trunc_t(policy));
        }
    };
    return std::visit(
        [&](const auto &v0, const auto &v1) -> ret_t {
            using type0 = std::decay_t<decltype(v0)>;
            using type1 = std::decay_t<decltype(v1)>;
            if constexpr (std::is_same_v<type0, type1>) {
                if (v0 != v1) {
                    throw std::invalid_argument("Unable to multiply two power series if their truncation levels do not match");
                }
                return run(v0);
            } else if constexpr (std::is_same_v<type0, nt>) {
                return run(v1);
            } else if constexpr (std::is_same_v<type1, nt>) {
                return run(v0);
            } else {
                throw std::invalid_argument("Unable to multiply two power series if their truncation policies do not match");
            }
        },
        ps0.trunc(), ps1.trunc());
}

template <typename K, typename C0, typename C1>
inline auto operator*(const p_series<K, C0> &a, const p_series<K, C1> &b)
{
    return series_mul(a, b);
}
template <typename K, typename C>
inline p_series<K, C> &operator*=(p_series<K, C> &a, const p_series<K, C> &b)
{
    a = a * b;
    return a;
}

// addition / subtraction keep the same policy rules (:951-1099); the sum is truncated by the resulting policy
namespace detail
{
template <bool Sign, typename K, typename C>
inline p_series<K, C> ps_addsub(const p_series<K, C> &a, const p_series<K, C> &b)
{
    using ps_t = p_series<K, C>;
    using nt = power_series::no_truncation;
    auto sum = [&]() {
        if constexpr (Sign) {
            return a.poly() + b.poly();
        } else {
            return a.poly() - b.poly();
        }
    };
    return std::visit(
        [&](const auto &v0, const auto &v1) -> ps_t {
            using type0 = std::decay_t<decltype(v0)>;
            using type1 = std::decay_t<decltype(v1)>;
            const char *what = Sign ? "add" : "subtract";
            if constexpr (std::is_same_v<type0, type1>) {
                if (v0 != v1) {
                    throw std::invalid_argument(std::string("Unable to ") + what + " two power series if their truncation levels do not match");
                }
                return ps_t(sum(), typename ps_t::trunc_t(v0));
            } else if constexpr (std::is_same_v<type0, nt> || std::is_same_v<type1, nt>) {
                ps_t r(sum());
                if constexpr (std::is_same_v<type0, nt>) {
                    r.trunc() = typename ps_t::trunc_t(v1);
                } else {
                    r.trunc() = typename ps_t::trunc_t(v0);
                }
                truncate(r);
                return r;
            } else {
                throw std::invalid_argument(std::string("Unable to ") + what + " two power series if their truncation policies do not match");
            }
        },
        a.trunc(), b.trunc());
}
} // namespace detail
template <typename K, typename C>
inline p_series<K, C> operator+(const p_series<K, C> &a, const p_series<K, C> &b)
{
    return detail::ps_addsub<true>(a, b);
}
template <typename K, typename C>
inline p_series<K, C> operator-(const p_series<K, C> &a, const p_series<K, C> &b)
{
    return detail::ps_addsub<false>(a, b);
}

// make_p_series / make_p_series_t / make_p_series_p (:567-891)
template <typename PS, typename... Names, std::enable_if_t<(std::is_convertible_v<Names, std::string> && ...), int> = 0>
inline std::array<PS, sizeof...(Names)> make_p_series(const Names &...names)
{
    auto polys = make_polynomials<typename PS::poly_t>(names...);
    std::array<PS, sizeof...(Names)> r;
    for (std::size_t i = 0; i < r.size(); ++i) r[i] = PS(std::move(polys[i]));
    return r;
}
template <typename PS, typename... Names>
inline std::array<PS, sizeof...(Names)> make_p_series(const symbol_set &ss, const Names &...names)
{
    auto polys = make_polynomials<typename PS::poly_t>(ss, names...);
    std::array<PS, sizeof...(Names)> r;
    for (std::size_t i = 0; i < r.size(); ++i) r[i] = PS(std::move(polys[i]));
    return r;
}
template <typename PS, typename U, typename... Names, std::enable_if_t<std::is_integral_v<U>, int> = 0>
inline std::array<PS, sizeof...(Names)> make_p_series_t(const U &d, const Names &...names)
{
    auto r = make_p_series<PS>(names...);
    for (auto &p : r) set_truncation(p, d);
    return r;
}
template <typename PS, typename U, typename... Names>
inline std::array<PS, sizeof...(Names)> make_p_series_t(const symbol_set &ss, const U &d, const Names &...names)
{
    auto r = make_p_series<PS>(ss, names...);
    for (auto &p : r) set_truncation(p, d);
    return r;
}
template <typename PS, typename U, typename... Names, std::enable_if_t<std::is_integral_v<U>, int> = 0>
inline std::array<PS, sizeof...(Names)> make_p_series_p(const U &d, const symbol_set &tss, const Names &...names)
{
    auto r = make_p_series<PS>(names...);
    for (auto &p : r) set_truncation(p, d, tss);
    return r;
}
template <typename PS, typename U, typename... Names>
inline std::array<PS, sizeof...(Names)> make_p_series_p(const symbol_set &ss, const U &d, const symbol_set &tss, const Names &...names)
{
    auto r = make_p_series<PS>(ss, names...);
    for (auto &p : r) set_truncation(p, d, tss);
    return r;
}

} // namespace obake

#endif
