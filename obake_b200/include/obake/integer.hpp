// obake/integer.hpp -- exact integer coefficient type for the host side of the B200 engine.
//
// The reference's integer coefficients are mppp::integer<N> from mp++ (an external dependency that is
// not installed in this image; N is only the number of inline limbs before mp++ spills to GMP).  This
// header provides a small sign-magnitude multiprecision integer with the operations the series-product
// path and its tests need (ring operations, comparison, nbits, pow, streaming) and the two's-complement
// limb views the C ABI exchanges.  When mp++ IS available, specialise obake::b200::cf_traits for
// mppp::integer<N> instead -- nothing below the ABI changes.
#ifndef OBAKE_B200_INTEGER_HPP
#define OBAKE_B200_INTEGER_HPP

#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <ostream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

namespace obake
{

namespace detail
{
// Little-endian limbs with room for three of them inline (mppp::integer<N> keeps N limbs inline as well): the
// coefficients of the benchmark products (at most 190 bits) never touch the heap, which is what the term-by-term
// rebuild of a multi-million-term result pays for otherwise.  The object is 32 bytes, holds no pointer into
// itself (the inline limbs and the heap pointer share storage) and its all-zero byte pattern is the empty vector,
// so a table of terms can be allocated as zeroed pages without running a constructor per slot
// (detail::zero_init_ok, polynomial.hpp).  One spare bit is kept for the owner: integer stores its sign there.
class limb_vec
{
    static constexpr std::uint32_t N = 3;
    union {
        std::uint64_t m_inl[N];
        struct {
            std::uint64_t *ptr;
            std::uint64_t cap;
        } m_heap;
    };
    std::uint32_t m_size = 0;
    std::uint16_t m_on_heap = 0;
    std::uint16_t m_flag = 0;

    std::size_t capacity() const
    {
        return m_on_heap ? static_cast<std::size_t>(m_heap.cap) : N;
    }
    void grow(std::size_t cap)
    {
        const std::size_t cur = capacity();
        if (cap <= cur) {
            return;
        }
        cap = std::max(cap, cur * 2u);
        auto *p = new std::uint64_t[cap];
        std::copy(data(), data() + m_size, p);
        release();
        m_heap.ptr = p;
        m_heap.cap = cap;
        m_on_heap = 1;
    }
    void release() noexcept
    {
        if (m_on_heap) {
            delete[] m_heap.ptr;
            m_on_heap = 0;
        }
    }
    void steal(limb_vec &o) noexcept
    {
        // bitwise hand-over: whatever o held (inline limbs or the heap block) is now ours
        m_inl[0] = o.m_inl[0];
        m_inl[1] = o.m_inl[1];
        m_inl[2] = o.m_inl[2];
        m_size = o.m_size;
        m_on_heap = o.m_on_heap;
        m_flag = o.m_flag;
        o.m_size = 0;
        o.m_on_heap = 0;
        o.m_flag = 0;
    }

public:
    limb_vec() : m_inl{0, 0, 0} {}
    explicit limb_vec(std::size_t n) : m_inl{0, 0, 0}
    {
        resize(n);
    }
    limb_vec(const limb_vec &o) : m_inl{0, 0, 0}
    {
        assign(o.data(), o.data() + o.m_size);
        m_flag = o.m_flag;
    }
    limb_vec(limb_vec &&o) noexcept
    {
        steal(o);
    }
    limb_vec &operator=(const limb_vec &o)
    {
        if (this != &o) {
            assign(o.data(), o.data() + o.m_size);
            m_flag = o.m_flag;
        }
        return *this;
    }
    limb_vec &operator=(limb_vec &&o) noexcept
    {
        if (this != &o) {
            release();
            steal(o);
        }
        return *this;
    }
    ~limb_vec()
    {
        release();
    }
    // true when the limbs live in a heap block (destroying or relocating the object then needs more than a memcpy)
    bool on_heap() const
    {
        return m_on_heap != 0;
    }
    // the owner's spare bit (not interpreted here; copied and moved with the value)
    bool flag() const
    {
        return m_flag != 0;
    }
    void set_flag(bool f)
    {
        m_flag = f ? 1 : 0;
    }
    std::uint64_t *data()
    {
        return m_on_heap ? m_heap.ptr : m_inl;
    }
    const std::uint64_t *data() const
    {
        return m_on_heap ? m_heap.ptr : m_inl;
    }
    std::size_t size() const
    {
        return m_size;
    }
    bool empty() const
    {
        return m_size == 0u;
    }
    std::uint64_t &operator[](std::size_t i)
    {
        return data()[i];
    }
    const std::uint64_t &operator[](std::size_t i) const
    {
        return data()[i];
    }
    std::uint64_t &back()
    {
        return data()[m_size - 1u];
    }
    const std::uint64_t &back() const
    {
        return data()[m_size - 1u];
    }
    void push_back(std::uint64_t v)
    {
        grow(m_size + 1u);
        data()[m_size++] = v;
    }
    void pop_back()
    {
        --m_size;
    }
    void clear()
    {
        m_size = 0;
    }
    void resize(std::size_t n)
    {
        grow(n);
        std::uint64_t *p = data();
        for (std::size_t i = m_size; i < n; ++i) {
            p[i] = 0u;
        }
        m_size = static_cast<std::uint32_t>(n);
    }
    template <typename It>
    void assign(It first, It last)
    {
        const auto n = static_cast<std::size_t>(last - first);
        m_size = 0;
        grow(n);
        std::copy(first, last, data());
        m_size = static_cast<std::uint32_t>(n);
    }
    std::uint64_t *begin()
    {
        return data();
    }
    std::uint64_t *end()
    {
        return data() + m_size;
    }
    const std::uint64_t *begin() const
    {
        return data();
    }
    const std::uint64_t *end() const
    {
        return data() + m_size;
    }
    // value comparison of the limbs (the owner's bit is the owner's business)
    friend bool operator==(const limb_vec &a, const limb_vec &b)
    {
        return a.m_size == b.m_size && std::equal(a.data(), a.data() + a.m_size, b.data());
    }
    friend bool operator!=(const limb_vec &a, const limb_vec &b)
    {
        return !(a == b);
    }
};
static_assert(sizeof(limb_vec) == 32, "limb_vec is meant to be 32 bytes");
} // namespace detail

class integer
{
    // magnitude, little endian, no leading zero limbs; zero is the empty vector
    // (the sign lives in the limb vector's spare bit: sizeof(integer) == 32, all-zero bytes == 0)
    detail::limb_vec m_mag;
    bool neg() const
    {
        return m_mag.flag();
    }
    void set_neg(bool n)
    {
        m_mag.set_flag(n);
    }
    // replace the magnitude, keep the sign
    void set_mag(detail::limb_vec &&r)
    {
        const bool n = neg();
        m_mag = std::move(r);
        set_neg(n);
    }
    using u128 = unsigned __int128;

    void trim()
    {
        while (!m_mag.empty() && m_mag.back() == 0u) {
            m_mag.pop_back();
        }
        if (m_mag.empty()) {
            set_neg(false);
        }
    }
    static int cmp_mag(const detail::limb_vec &a, const detail::limb_vec &b)
    {
        if (a.size() != b.size()) {
            return a.size() < b.size() ? -1 : 1;
        }
        for (std::size_t i = a.size(); i-- > 0;) {
            if (a[i] != b[i]) {
                return a[i] < b[i] ? -1 : 1;
            }
        }
        return 0;
    }
    static detail::limb_vec add_mag(const detail::limb_vec &a, const detail::limb_vec &b)
    {
        const auto &l = a.size() >= b.size() ? a : b;
        const auto &s = a.size() >= b.size() ? b : a;
        detail::limb_vec r(l.size() + 1u);
        std::uint64_t c = 0;
        for (std::size_t i = 0; i < l.size(); ++i) {
            const u128 t = static_cast<u128>(l[i]) + (i < s.size() ? s[i] : 0u) + c;
            r[i] = static_cast<std::uint64_t>(t);
            c = static_cast<std::uint64_t>(t >> 64);
        }
        r[l.size()] = c;
        return r;
    }
    // a - b, |a| >= |b|
    static detail::limb_vec sub_mag(const detail::limb_vec &a, const detail::limb_vec &b)
    {
        detail::limb_vec r(a.size());
        std::uint64_t brw = 0;
        for (std::size_t i = 0; i < a.size(); ++i) {
            const std::uint64_t bi = i < b.size() ? b[i] : 0u;
            const u128 t = static_cast<u128>(a[i]) - bi - brw;
            r[i] = static_cast<std::uint64_t>(t);
            brw = static_cast<std::uint64_t>(t >> 64) & 1u;
        }
        return r;
    }

public:
    integer() = default;
    template <typename I, std::enable_if_t<std::is_integral_v<I>, int> = 0>
    integer(I v)
    {
        if (v != 0) {
            if constexpr (std::is_signed_v<I>) {
                set_neg(v < 0);
                const auto u = static_cast<std::uint64_t>(static_cast<std::int64_t>(v));
                m_mag.push_back(neg() ? std::uint64_t(0) - u : u);
            } else {
                m_mag.push_back(static_cast<std::uint64_t>(v));
            }
        }
    }
    explicit integer(double d)
    {
        if (!(d == d) || d >= 9223372036854775808.0 || d <= -9223372036854775808.0) {
            throw std::domain_error("Cannot construct an integer from a non-finite or too large floating-point value");
        }
        *this = integer(static_cast<std::int64_t>(d));
    }
    explicit integer(const std::string &s)
    {
        std::size_t i = 0;
        bool neg = false;
        if (i < s.size() && (s[i] == '-' || s[i] == '+')) {
            neg = s[i] == '-';
            ++i;
        }
        if (i == s.size()) {
            throw std::invalid_argument("invalid integer string");
        }
        for (; i < s.size(); ++i) {
            if (s[i] < '0' || s[i] > '9') {
                throw std::invalid_argument("invalid integer string");
            }
            *this *= integer(10);
            *this += integer(s[i] - '0');
        }
        set_neg(neg && !m_mag.empty());
    }

    // Two's-complement view on `n` little-endian limbs; throws if the value does not fit.
    void to_limbs(std::uint64_t *out, unsigned n) const
    {
        if (nbits() + 1u > 64u * n) {
            // exactly -2^(64n-1) also fits; keep the simple, slightly conservative rule
            if (!(neg() && nbits() == 64u * n && is_pow2())) {
                throw std::overflow_error("integer does not fit in the requested number of limbs");
            }
        }
        for (unsigned i = 0; i < n; ++i) {
            out[i] = i < m_mag.size() ? m_mag[i] : 0u;
        }
        if (neg()) {
            std::uint64_t c = 1;
            for (unsigned i = 0; i < n; ++i) {
                out[i] = ~out[i] + c;
                c = (c != 0u && out[i] == 0u) ? 1u : 0u;
            }
        }
    }
    static integer from_limbs(const std::uint64_t *in, unsigned n)
    {
        integer r;
        const bool neg = n > 0u && (in[n - 1u] >> 63) != 0u;
        r.m_mag.assign(in, in + n);
        if (neg) {
            std::uint64_t c = 1;
            for (auto &l : r.m_mag) {
                l = ~l + c;
                c = (c != 0u && l == 0u) ? 1u : 0u;
            }
        }
        r.set_neg(neg);
        r.trim();
        return r;
    }
    // Limbs needed for a two's-complement representation.
    unsigned limbs_needed() const
    {
        const unsigned b = nbits() + 1u;
        return std::max(1u, (b + 63u) / 64u);
    }

    // true when the limbs live in a heap block (more than three limbs): the term containers use it to skip the
    // destructor walk over tables whose coefficients are all inline
    bool _on_heap() const
    {
        return m_mag.on_heap();
    }
    bool is_zero() const
    {
        return m_mag.empty();
    }
    bool is_pow2() const
    {
        if (m_mag.empty()) {
            return false;
        }
        for (std::size_t i = 0; i + 1u < m_mag.size(); ++i) {
            if (m_mag[i] != 0u) {
                return false;
            }
        }
        return (m_mag.back() & (m_mag.back() - 1u)) == 0u;
    }
    int sgn() const
    {
        return m_mag.empty() ? 0 : (neg() ? -1 : 1);
    }
    unsigned nbits() const
    {
        if (m_mag.empty()) {
            return 0;
        }
        return static_cast<unsigned>(64u * (m_mag.size() - 1u) + (64u - static_cast<unsigned>(__builtin_clzll(m_mag.back()))));
    }
    std::size_t size() const
    {
        return m_mag.size();
    }
    explicit operator double() const
    {
        long double r = 0;
        for (std::size_t i = m_mag.size(); i-- > 0;) {
            r = r * 18446744073709551616.0L + static_cast<long double>(m_mag[i]);
        }
        return static_cast<double>(neg() ? -r : r);
    }
    explicit operator std::int64_t() const
    {
        if (nbits() > 63u && !(neg() && nbits() == 64u && is_pow2())) {
            throw std::overflow_error("integer does not fit in int64_t");
        }
        const std::uint64_t u = m_mag.empty() ? 0u : m_mag[0];
        return static_cast<std::int64_t>(neg() ? std::uint64_t(0) - u : u);
    }

    integer operator-() const
    {
        integer r(*this);
        if (!r.m_mag.empty()) {
            r.set_neg(!r.neg());
        }
        return r;
    }
    integer &operator+=(const integer &o)
    {
        if (neg() == o.neg()) {
            set_mag(add_mag(m_mag, o.m_mag));
        } else {
            const int c = cmp_mag(m_mag, o.m_mag);
            if (c == 0) {
                m_mag.clear();
            } else if (c > 0) {
                set_mag(sub_mag(m_mag, o.m_mag));
            } else {
                m_mag = sub_mag(o.m_mag, m_mag);
                set_neg(o.neg());
            }
        }
        trim();
        return *this;
    }
    integer &operator-=(const integer &o)
    {
        return *this += -o;
    }
    integer &operator*=(const integer &o)
    {
        if (m_mag.empty() || o.m_mag.empty()) {
            m_mag.clear();
            set_neg(false);
            return *this;
        }
        detail::limb_vec r(m_mag.size() + o.m_mag.size());
        for (std::size_t i = 0; i < m_mag.size(); ++i) {
            std::uint64_t carry = 0;
            for (std::size_t j = 0; j < o.m_mag.size(); ++j) {
                const u128 t = static_cast<u128>(m_mag[i]) * o.m_mag[j] + r[i + j] + carry;
                r[i + j] = static_cast<std::uint64_t>(t);
                carry = static_cast<std::uint64_t>(t >> 64);
            }
            r[i + o.m_mag.size()] += carry;
        }
        const bool n = neg() != o.neg();
        m_mag = std::move(r);
        set_neg(n);
        trim();
        return *this;
    }
    friend integer operator+(integer a, const integer &b)
    {
        return a += b;
    }
    friend integer operator-(integer a, const integer &b)
    {
        return a -= b;
    }
    friend integer operator*(integer a, const integer &b)
    {
        return a *= b;
    }
    friend bool operator==(const integer &a, const integer &b)
    {
        return a.neg() == b.neg() && a.m_mag == b.m_mag;
    }
    friend bool operator!=(const integer &a, const integer &b)
    {
        return !(a == b);
    }
    friend bool operator<(const integer &a, const integer &b)
    {
        if (a.neg() != b.neg()) {
            return a.neg();
        }
        const int c = cmp_mag(a.m_mag, b.m_mag);
        return a.neg() ? c > 0 : c < 0;
    }
    friend bool operator>(const integer &a, const integer &b)
    {
        return b < a;
    }
    friend bool operator<=(const integer &a, const integer &b)
    {
        return !(b < a);
    }
    friend bool operator>=(const integer &a, const integer &b)
    {
        return !(a < b);
    }

    // divide the magnitude by a small value, return the remainder (for printing)
    std::uint64_t div_small(std::uint64_t d)
    {
        std::uint64_t rem = 0;
        for (std::size_t i = m_mag.size(); i-- > 0;) {
            const u128 cur = (static_cast<u128>(rem) << 64) | m_mag[i];
            m_mag[i] = static_cast<std::uint64_t>(cur / d);
            rem = static_cast<std::uint64_t>(cur % d);
        }
        trim();
        return rem;
    }
    std::string to_string() const
    {
        if (m_mag.empty()) {
            return "0";
        }
        integer t(*this);
        const bool neg = t.neg();
        t.set_neg(false);
        std::string s;
        while (!t.m_mag.empty()) {
            std::uint64_t r = t.div_small(10000000000000000000ull);
            for (int i = 0; i < 19; ++i) {
                s.push_back(static_cast<char>('0' + r % 10u));
                r /= 10u;
                if (t.m_mag.empty() && r == 0u) {
                    break;
                }
            }
        }
        if (neg) {
            s.push_back('-');
        }
        std::reverse(s.begin(), s.end());
        return s;
    }
    friend std::ostream &operator<<(std::ostream &os, const integer &n)
    {
        return os << n.to_string();
    }
};

static_assert(sizeof(integer) == 32, "integer is meant to be 32 bytes (three inline limbs, size, sign)");

inline integer pow(const integer &b, unsigned n)
{
    integer r(1), a(b);
    while (n) {
        if (n & 1u) {
            r *= a;
        }
        n >>= 1;
        if (n) {
            a *= a;
        }
    }
    return r;
}

} // namespace obake

// Spelling used by the reference's sources and tests.  With the real mp++ installed, define
// OBAKE_B200_NO_MPPP_SHIM and include <mp++/integer.hpp> instead.
#if !defined(OBAKE_B200_NO_MPPP_SHIM)
namespace mppp
{
template <std::size_t SSize>
using integer = ::obake::integer;
}
#endif

#endif
