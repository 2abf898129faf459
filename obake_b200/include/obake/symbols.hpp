// obake/symbols.hpp -- sorted symbol sets (host bookkeeping of the B200 engine).
// Mirrors include/obake/symbols.hpp:36 of the reference: a symbol_set is a sorted set of unique names,
// and the position of a name in it is the digit its exponent occupies in a packed monomial.
#ifndef OBAKE_B200_SYMBOLS_HPP
#define OBAKE_B200_SYMBOLS_HPP

#include <algorithm>
#include <cstddef>
#include <initializer_list>
#include <string>
#include <vector>

namespace obake
{

using symbol_idx = std::size_t;

class symbol_set
{
    std::vector<std::string> m_v;

public:
    symbol_set() = default;
    symbol_set(std::initializer_list<std::string> l) : m_v(l)
    {
        normalise();
    }
    explicit symbol_set(std::vector<std::string> v) : m_v(std::move(v))
    {
        normalise();
    }
    void normalise()
    {
        std::sort(m_v.begin(), m_v.end());
        m_v.erase(std::unique(m_v.begin(), m_v.end()), m_v.end());
    }
    void insert(const std::string &s)
    {
        m_v.push_back(s);
        normalise();
    }
    std::size_t size() const
    {
        return m_v.size();
    }
    bool empty() const
    {
        return m_v.empty();
    }
    auto begin() const
    {
        return m_v.begin();
    }
    auto end() const
    {
        return m_v.end();
    }
    const std::string &operator[](std::size_t i) const
    {
        return m_v[i];
    }
    // position of s, or size() if absent
    std::size_t index_of(const std::string &s) const
    {
        const auto it = std::lower_bound(m_v.begin(), m_v.end(), s);
        return (it != m_v.end() && *it == s) ? static_cast<std::size_t>(it - m_v.begin()) : m_v.size();
    }
    friend bool operator==(const symbol_set &a, const symbol_set &b)
    {
        return a.m_v == b.m_v;
    }
    friend bool operator!=(const symbol_set &a, const symbol_set &b)
    {
        return !(a == b);
    }
};

// sorted positions of the symbols of `s` that are present in `ss` (series.hpp:3911-3924: symbols of s
// absent from the series are ignored).
inline std::vector<unsigned> ss_intersect_idx(const symbol_set &s, const symbol_set &ss)
{
    std::vector<unsigned> r;
    for (const auto &n : s) {
        const auto i = ss.index_of(n);
        if (i != ss.size()) {
            r.push_back(static_cast<unsigned>(i));
        }
    }
    std::sort(r.begin(), r.end());
    return r;
}

// union of two symbol sets plus, for each input set, the position in the union of each of its symbols
inline symbol_set merge_symbol_sets(const symbol_set &a, const symbol_set &b, std::vector<std::size_t> &pos_a,
                                    std::vector<std::size_t> &pos_b)
{
    std::vector<std::string> all(a.begin(), a.end());
    all.insert(all.end(), b.begin(), b.end());
    symbol_set u(std::move(all));
    pos_a.clear();
    pos_b.clear();
    for (const auto &n : a) {
        pos_a.push_back(u.index_of(n));
    }
    for (const auto &n : b) {
        pos_b.push_back(u.index_of(n));
    }
    return u;
}

} // namespace obake

#endif
