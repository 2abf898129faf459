// obake/kpack.hpp -- Kronecker packing of exponent vectors (host side of the B200 engine).
//
// Same public surface as the reference's include/obake/kpack.hpp:219-375 (kpacker<T>, kunpacker<T>,
// detail::kpack_get_{delta,lims,klims}, detail::kpack_max_size) with the same error behaviour
// (std::overflow_error / std::out_of_range / std::invalid_argument and the same message substrings).
// The constant tables of the reference's src/kpack.cpp are regenerated at first use from the recipe of
// tools/kpack.ipynb (largest prime <= 2^(bw/s) - 1 per size); tests/test_kpack.py checks them
// number-for-number against the reference (tests/golden/kpack_tables.json).
#ifndef OBAKE_B200_KPACK_HPP
#define OBAKE_B200_KPACK_HPP

#include <array>
#include <cstdint>
#include <limits>
#include <stdexcept>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>

namespace obake
{

namespace detail
{

template <typename T>
inline constexpr bool is_kpackable_v = std::is_same_v<T, std::int32_t> || std::is_same_v<T, std::uint32_t>
                                       || std::is_same_v<T, std::int64_t> || std::is_same_v<T, std::uint64_t>;

using kp_u128 = unsigned __int128;

inline std::uint64_t kp_mulmod(std::uint64_t a, std::uint64_t b, std::uint64_t m)
{
    return static_cast<std::uint64_t>(static_cast<kp_u128>(a) * b % m);
}

inline bool kp_is_prime(std::uint64_t n)
{
    if (n < 2u) {
        return false;
    }
    constexpr std::uint64_t bases[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    for (auto p : bases) {
        if (n % p == 0u) {
            return n == p;
        }
    }
    std::uint64_t d = n - 1u;
    int r = 0;
    while ((d & 1u) == 0u) {
        d >>= 1;
        ++r;
    }
    for (auto a : bases) {
        std::uint64_t x = 1, b = a % n, e = d;
        while (e) {
            if (e & 1u) {
                x = kp_mulmod(x, b, n);
            }
            b = kp_mulmod(b, b, n);
            e >>= 1;
        }
        if (x == 1u || x == n - 1u) {
            continue;
        }
        bool composite = true;
        for (int i = 1; i < r; ++i) {
            x = kp_mulmod(x, x, n);
            if (x == n - 1u) {
                composite = false;
                break;
            }
        }
        if (composite) {
            return false;
        }
    }
    return true;
}

// All the per-size constants for one packable type.
template <typename T>
struct kpack_tables_t {
    using U = std::make_unsigned_t<T>;
    static constexpr int nbits = std::numeric_limits<U>::digits;
    static constexpr int bw = std::numeric_limits<T>::digits; // value bits (sign excluded)
    static constexpr unsigned cap = 24;

    unsigned max_size = 0;
    std::array<T, cap> deltas{}, lims{}, klims{};
    // divcnst[size-1][j] = (m', sh1, sh2) for the divisor delta^j, j = 0..size
    std::array<std::array<std::tuple<U, unsigned, unsigned>, cap + 1>, cap> divcnst{};

    kpack_tables_t()
    {
        unsigned s = 1;
        while (static_cast<unsigned>(bw) / s >= 3u) {
            std::uint64_t d;
            if (s == 1u) {
                d = static_cast<std::uint64_t>(std::numeric_limits<T>::max());
            } else {
                // floor(2^(bw/s)) - 1 via an integer s-th root of 2^bw
                std::uint64_t lo = 0, hi = std::uint64_t(1) << (static_cast<unsigned>(bw) / s + 1u);
                while (lo < hi) {
                    const std::uint64_t mid = lo + (hi - lo + 1u) / 2u;
                    kp_u128 p = 1;
                    bool over = false;
                    for (unsigned i = 0; i < s; ++i) {
                        p *= mid;
                        if (p > (kp_u128(1) << bw)) {
                            over = true;
                            break;
                        }
                    }
                    if (over) {
                        hi = mid - 1u;
                    } else {
                        lo = mid;
                    }
                }
                d = lo - 1u;
                while (!kp_is_prime(d)) {
                    --d;
                }
            }
            const std::uint64_t lim = std::is_signed_v<T> ? (d - 1u) / 2u : d - 1u;
            kp_u128 kl = 0, cur = 1;
            for (unsigned j = 0; j < s; ++j) {
                kl += static_cast<kp_u128>(lim) * cur;
                cur *= d;
            }
            deltas[s - 1u] = static_cast<T>(d);
            lims[s - 1u] = static_cast<T>(lim);
            klims[s - 1u] = static_cast<T>(static_cast<std::uint64_t>(kl));
            ++s;
        }
        max_size = s - 1u;
        for (unsigned sz = 1; sz <= max_size; ++sz) {
            kp_u128 dd = 1;
            for (unsigned j = 0; j <= sz; ++j) {
                unsigned l = 0;
                while ((kp_u128(1) << l) < dd) {
                    ++l;
                }
                const kp_u128 num = ((kp_u128(1) << l) - dd) << nbits;
                const U mp = static_cast<U>(static_cast<U>(num / dd) + 1u);
                divcnst[sz - 1u][j] = std::make_tuple(mp, l < 1u ? l : 1u, l > 1u ? l - 1u : 0u);
                if (j < sz) {
                    dd *= static_cast<std::uint64_t>(deltas[sz - 1u]);
                }
            }
        }
    }
};

template <typename T>
inline const kpack_tables_t<T> &kpack_tables()
{
    static const kpack_tables_t<T> t;
    return t;
}

template <typename T>
inline unsigned kpack_max_size()
{
    return kpack_tables<T>().max_size;
}

template <typename T>
inline T kpack_get_delta(unsigned size)
{
    return kpack_tables<T>().deltas[size - 1u];
}

// (lim_min, lim_max) of one component.
template <typename T>
inline std::pair<T, T> kpack_get_lims(unsigned size)
{
    const auto lim = kpack_tables<T>().lims[size - 1u];
    if constexpr (std::is_signed_v<T>) {
        return {static_cast<T>(-lim), lim};
    } else {
        return {T(0), lim};
    }
}

// (klim_min, klim_max) of the coded value.
template <typename T>
inline std::pair<T, T> kpack_get_klims(unsigned size)
{
    const auto klim = kpack_tables<T>().klims[size - 1u];
    if constexpr (std::is_signed_v<T>) {
        return {static_cast<T>(-klim), klim};
    } else {
        return {T(0), klim};
    }
}

template <typename U>
inline U kp_mulhi(U a, U b)
{
    if constexpr (sizeof(U) == 4) {
        return static_cast<U>((static_cast<std::uint64_t>(a) * b) >> 32);
    } else {
        return static_cast<U>((static_cast<kp_u128>(a) * b) >> 64);
    }
}

template <typename T>
inline const char *kp_type_name()
{
    if constexpr (std::is_same_v<T, std::int32_t>) {
        return "int";
    } else if constexpr (std::is_same_v<T, std::uint32_t>) {
        return "unsigned int";
    } else if constexpr (std::is_same_v<T, std::int64_t>) {
        return "long";
    } else {
        return "unsigned long";
    }
}

} // namespace detail

template <typename T>
class kpacker
{
    static_assert(detail::is_kpackable_v<T>);
    T m_value = 0;
    T m_cur_prod = 1;
    unsigned m_index = 0;
    unsigned m_size;

public:
    explicit kpacker(unsigned size) : m_size(size)
    {
        if (size > detail::kpack_max_size<T>()) {
            throw std::overflow_error(std::string("Invalid size specified in the constructor of a Kronecker packer for the type '")
                                      + detail::kp_type_name<T>() + "': the maximum possible size is "
                                      + std::to_string(detail::kpack_max_size<T>()) + ", but a size of "
                                      + std::to_string(size) + " was specified instead");
        }
    }

    kpacker &operator<<(const T &n)
    {
        if (m_index == m_size) {
            throw std::out_of_range(std::string("Cannot push any more values to this Kronecker packer for the type '")
                                    + detail::kp_type_name<T>()
                                    + "': the number of values already pushed to the packer is equal to the packer's size ("
                                    + std::to_string(m_size) + ")");
        }
        const auto [lim_min, lim_max] = detail::kpack_get_lims<T>(m_size);
        if (n < lim_min || n > lim_max) {
            throw std::overflow_error("Cannot push the value " + std::to_string(n) + " to this Kronecker packer for the type '"
                                      + detail::kp_type_name<T>() + "': the value is outside the allowed range ["
                                      + std::to_string(lim_min) + ", " + std::to_string(lim_max) + "]");
        }
        using U = std::make_unsigned_t<T>;
        m_value = static_cast<T>(static_cast<U>(m_value) + static_cast<U>(n) * static_cast<U>(m_cur_prod));
        m_cur_prod = static_cast<T>(static_cast<U>(m_cur_prod) * static_cast<U>(detail::kpack_get_delta<T>(m_size)));
        ++m_index;
        return *this;
    }

    const T &get() const
    {
        return m_value;
    }
};

template <typename T>
class kunpacker
{
    static_assert(detail::is_kpackable_v<T>);
    T m_value;
    T m_cur_prod = 1;
    unsigned m_index = 0;
    unsigned m_size;

public:
    explicit kunpacker(const T &n, unsigned size) : m_value(n), m_size(size)
    {
        if (size == 0u) {
            if (n != T(0)) {
                throw std::invalid_argument("Only a value of zero can be used in a Kronecker unpacker "
                                            "with a size of zero, but a value of "
                                            + std::to_string(n) + " was provided instead");
            }
        } else {
            if (size > detail::kpack_max_size<T>()) {
                throw std::overflow_error(
                    std::string("Invalid size specified in the constructor of a Kronecker unpacker for the type '")
                    + detail::kp_type_name<T>() + "': the maximum possible size is "
                    + std::to_string(detail::kpack_max_size<T>()) + ", but a size of " + std::to_string(size)
                    + " was specified instead");
            }
            const auto [klim_min, klim_max] = detail::kpack_get_klims<T>(m_size);
            if (n < klim_min || n > klim_max) {
                throw std::overflow_error("The value " + std::to_string(n) + " passed to a Kronecker unpacker for the type '"
                                          + detail::kp_type_name<T>() + "' is outside the allowed range ["
                                          + std::to_string(klim_min) + ", " + std::to_string(klim_max) + "]");
            }
        }
    }

    kunpacker &operator>>(T &out)
    {
        if (m_index == m_size) {
            throw std::out_of_range("Cannot unpack any more values from this Kronecker unpacker: the number of "
                                    "values already unpacked is equal to the unpacker's size ("
                                    + std::to_string(m_size) + ")");
        }
        using U = std::make_unsigned_t<T>;
        const auto &tab = detail::kpack_tables<T>();
        const U delta = static_cast<U>(tab.deltas[m_size - 1u]);
        m_cur_prod = static_cast<T>(static_cast<U>(m_cur_prod) * delta);
        const auto [mp_d, sh1_d, sh2_d] = tab.divcnst[m_size - 1u][m_index];
        const auto [mp_r, sh1_r, sh2_r] = tab.divcnst[m_size - 1u][m_index + 1u];
        const U n = static_cast<U>(m_value) - static_cast<U>(detail::kpack_get_klims<T>(m_size).first);
        auto divcnst = [](U nn, U mp, unsigned sh1, unsigned sh2) {
            const U t1 = detail::kp_mulhi<U>(mp, nn);
            const U tmp = static_cast<U>(nn - t1) >> sh1;
            return static_cast<U>(static_cast<U>(t1 + tmp) >> sh2);
        };
        const U q_r = divcnst(n, mp_r, sh1_r, sh2_r);
        const U rem = static_cast<U>(n - q_r * static_cast<U>(m_cur_prod));
        const T q_d = static_cast<T>(divcnst(rem, mp_d, sh1_d, sh2_d));
        out = static_cast<T>(q_d + detail::kpack_get_lims<T>(m_size).first);
        ++m_index;
        return *this;
    }
};

} // namespace obake

#endif
