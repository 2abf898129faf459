// obake/polynomials/polynomial.hpp -- polynomial<Key, Cf> whose product runs on the B200 engine.
//
// Drop-in for the series-product path of the reference (include/obake/polynomials/polynomial.hpp):
//   operator* / series_mul(polynomial, polynomial)                      :2026-2032
//   truncated_mul(x, y, max_degree[, symbol_set])                       :2113-2127
//   poly_mul_impl_switch (shorter operand first)                        :2014-2022
//   poly_mul_impl (merge the symbol sets, re-pack the keys)             :1955-2010
//   poly_mul_impl_identical_ss (empty operand -> empty result)          :1878-1929   <- the seam
// Everything below the seam (overflow check, size estimate, segmentation, the multiply itself,
// :797-1582) is the CUDA engine behind include/obk_b200.h; there is NO CPU multiplication here
// (the reference's poly_mul_impl_simple small-input path, :1602-1874, is routed to the engine too).
// The container (detail::seg_table) has the reference's shape -- 2^k open-addressing tables, a term in table
// hash & (2^k-1) (series.hpp:396-400, 637-639) -- because the engine returns its terms grouped exactly that way and
// the rebuild of the host container is part of what a caller of operator* waits for.
#ifndef OBAKE_B200_POLYNOMIAL_HPP
#define OBAKE_B200_POLYNOMIAL_HPP

#include <algorithm>
#include <array>
#include <atomic>
#include <barrier>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <mutex>
#include <new>
#include <limits>
#include <iterator>
#include <stdexcept>
#include <sstream>
#include <string>
#include <thread>
#include <type_traits>
#include <unordered_map>
#include <utility>
#include <vector>

#if defined(__linux__)
#include <sys/mman.h>
#endif

#include "../b200/engine.hpp"
#include "../integer.hpp"
#include "../symbols.hpp"
#include "d_packed_monomial.hpp"
#include "packed_monomial.hpp"

namespace obake
{

// is_zero (math/is_zero.hpp): x == T(0)
template <typename C>
inline bool is_zero(const C &c)
{
    return c == C(0);
}

// mixed coefficient products (polynomial_mul_general_test: integer<1> x double -> double)
inline double operator*(const integer &a, double b)
{
    return static_cast<double>(a) * b;
}
inline double operator*(double a, const integer &b)
{
    return a * static_cast<double>(b);
}

namespace detail
{
template <typename K>
struct key_hasher {
    std::size_t operator()(const K &k) const
    {
        return ::obake::hash(k);
    }
};
template <typename T, typename = void>
struct is_poly : std::false_type {
};

// f(w) for w in [0, th) on th threads (the caller's included); the first exception thrown by any of them is
// rethrown here after all have finished.
template <typename F>
inline void parallel_for(std::size_t th, const F &f)
{
    if (th <= 1u) {
        f(std::size_t(0));
        return;
    }
    std::exception_ptr err;
    std::mutex mtx;
    auto guarded = [&](std::size_t w) {
        try {
            f(w);
        } catch (...) {
            std::lock_guard<std::mutex> g(mtx);
            if (!err) {
                err = std::current_exception();
            }
        }
    };
    std::vector<std::thread> pool;
    pool.reserve(th - 1u);
    for (std::size_t w = 1; w < th; ++w) {
        pool.emplace_back(guarded, w);
    }
    guarded(0);
    for (auto &p : pool) {
        p.join();
    }
    if (err) {
        std::rethrow_exception(err);
    }
}

// ---- storage traits of the term container --------------------------------------------------------------------
// holds_no_resource(v): destroying v is a no-op (nothing on the heap).  A table whose terms were all adopted in that
// state is released without a destructor walk over millions of slots.
template <typename T>
inline bool holds_no_resource(const T &)
{
    return std::is_trivially_destructible_v<T>;
}
inline bool holds_no_resource(const integer &n)
{
    return !n._on_heap();
}

// Storage of a table group.  Large blocks come straight from the kernel as an anonymous mapping (transparent huge
// pages requested: a 340 MB result is 170 page faults instead of 83 000), small ones from malloc.  The contents are
// NOT defined: slots are raw storage (a term is constructed in place when it is inserted); a table clears its part
// of the block right before it is filled (seg_table::bind).
//
// The last large block that was released is kept for the next product (one block, at most
// OBAKE_B200_SLAB_CACHE_MB megabytes, default 1024; 0 disables): a chain of products, or a benchmark loop, then
// reuses the pages instead of having the kernel unmap, remap and zero a few hundred MB per product -- measured on
// the sparse benchmark's 5.8 M-term result, fresh pages cost 145 ms of system time per product, more than filling
// the tables.
struct raw_block {
    void *ptr = nullptr;
    std::size_t bytes = 0;
    bool mapped = false;
    static constexpr std::size_t map_threshold = std::size_t(1) << 21;

    struct cache_t {
        std::mutex mtx;
        void *ptr = nullptr;
        std::size_t bytes = 0, limit = std::size_t(1024) << 20;
        cache_t()
        {
            if (const char *e = std::getenv("OBAKE_B200_SLAB_CACHE_MB")) {
                limit = static_cast<std::size_t>(std::strtoull(e, nullptr, 10)) << 20;
            }
        }
    };
    static cache_t &cache()
    {
        // never destroyed: a polynomial with static storage duration may be released after any other static
        static cache_t *c = new cache_t;
        return *c;
    }

    raw_block() = default;
    explicit raw_block(std::size_t n) : bytes(n)
    {
        if (n == 0u) {
            return;
        }
#if defined(__linux__)
        if (n >= map_threshold) {
            {
                // a cached block that is large enough, and not absurdly larger than the request
                auto &c = cache();
                std::lock_guard<std::mutex> g(c.mtx);
                if (c.ptr != nullptr && c.bytes >= n && c.bytes / 4u <= n) {
                    ptr = c.ptr;
                    bytes = c.bytes;
                    mapped = true;
                    c.ptr = nullptr;
                    c.bytes = 0;
                    return;
                }
            }
            void *p = ::mmap(nullptr, n, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
            if (p != MAP_FAILED) {
#if defined(MADV_HUGEPAGE)
                ::madvise(p, n, MADV_HUGEPAGE);
#endif
                ptr = p;
                mapped = true;
                return;
            }
        }
#endif
        ptr = std::malloc(n);
        if (ptr == nullptr) {
            throw std::bad_alloc();
        }
    }
    raw_block(const raw_block &) = delete;
    raw_block &operator=(const raw_block &) = delete;
    raw_block(raw_block &&o) noexcept : ptr(o.ptr), bytes(o.bytes), mapped(o.mapped)
    {
        o.ptr = nullptr;
        o.bytes = 0;
        o.mapped = false;
    }
    raw_block &operator=(raw_block &&o) noexcept
    {
        if (this != &o) {
            reset();
            ptr = o.ptr;
            bytes = o.bytes;
            mapped = o.mapped;
            o.ptr = nullptr;
            o.bytes = 0;
            o.mapped = false;
        }
        return *this;
    }
    ~raw_block()
    {
        reset();
    }
    void reset() noexcept
    {
        if (ptr != nullptr) {
#if defined(__linux__)
            if (mapped) {
                void *drop = ptr;
                std::size_t drop_bytes = bytes;
                {
                    // keep the larger of this block and the cached one
                    auto &c = cache();
                    std::lock_guard<std::mutex> g(c.mtx);
                    if (bytes <= c.limit && bytes > c.bytes) {
                        drop = c.ptr;
                        drop_bytes = c.bytes;
                        c.ptr = ptr;
                        c.bytes = bytes;
                    }
                }
                if (drop != nullptr) {
                    ::munmap(drop, drop_bytes);
                }
            } else
#endif
                std::free(ptr);
        }
        ptr = nullptr;
        bytes = 0;
        mapped = false;
    }
};

// The term container of a series (series.hpp:396-400, `s_table`): 2^log2 open-addressing tables, a term lives in table
// hash(key) & (2^log2 - 1).  The engine hands its results over grouped exactly that way (obk_poly_result::seg_offsets),
// so a product is adopted table by table, in parallel, without rehashing anything (_adopt_grouped) -- the
// reference's multiplication threads fill the same tables in place (polynomial.hpp:1441-1450).
//
// Storage.  A table is an array of `cap` RAW slots (any size, not a power of two: the home slot is the high half of
// hash * cap, so a table holds n / 0.7 slots, not up to twice that) plus one occupancy byte per slot; a term object
// exists in a slot exactly while its occupancy byte is set.  The tables of an adopted product share ONE block
// (`m_slab`): one allocation and one release per product instead of two per table, nothing to construct or clear
// but the occupancy bytes.  A table that has to grow later gets a block of its own.  (The first version kept two std::vectors per table and 64-byte slots: rebuilding the
// 5.8 M-term result of the sparse benchmark took 70 ms on 8 cores and destroying it 80 ms, against 0.9 ms for the
// product itself on the GPU; tools/bench_adopt.cpp measures this part alone.)
template <typename K, typename C>
class seg_table
{
public:
    using value_type = std::pair<K, C>;

private:
    static constexpr bool trivial_dtor = std::is_trivially_destructible_v<value_type>;
    struct tab {
        value_type *slots = nullptr;
        unsigned char *full = nullptr;
        std::size_t cap = 0, n = 0;
        raw_block own; // storage of a table that is not (or no longer) part of the slab
    };
    std::vector<tab> m_tabs = std::vector<tab>(1);
    raw_block m_slab;
    unsigned m_log2 = 0;
    std::size_t m_size = 0;
    // false: every stored term satisfied holds_no_resource when it was adopted and nothing has been handed out for
    // modification since, so the slots need no destructor calls
    bool m_walk = !trivial_dtor;

    static std::size_t slot_bytes(std::size_t cap)
    {
        // slots first (alignment of value_type), occupancy bytes behind them, rounded to 64 bytes
        return (cap * sizeof(value_type) + cap + 63u) & ~std::size_t(63);
    }
    static void bind(tab &t, void *base, std::size_t cap)
    {
        t.slots = static_cast<value_type *>(base);
        t.full = reinterpret_cast<unsigned char *>(t.slots + cap);
        t.cap = cap;
        // Only the occupancy bytes NEED clearing (the slots are raw storage), but the whole table is cleared: the
        // sequential pass pulls the table into this core's cache, and the random insertions that follow then hit
        // it instead of missing to DRAM one line at a time (measured: filling 5.8 M terms single-threaded takes
        // 130 ms this way and 266 ms when only the bytes are cleared).
        std::memset(base, 0, cap * sizeof(value_type) + cap);
    }
    static std::size_t cap_for(std::size_t n)
    {
        // load factor 0.7
        return std::max<std::size_t>(8u, n + (n * 3u + 6u) / 7u + 1u);
    }
    static std::size_t slot_of(std::size_t h, std::size_t cap, unsigned log2)
    {
        // the low bits chose the table; the rest is scrambled (packed keys are far from uniform) and mapped to
        // [0, cap) by the high half of the product with cap
        const std::uint64_t x = (static_cast<std::uint64_t>(h) >> log2) * 0x9E3779B97F4A7C15ull;
        return static_cast<std::size_t>((static_cast<unsigned __int128>(x ^ (x >> 29)) * cap) >> 64);
    }
    static std::size_t next(std::size_t j, std::size_t cap)
    {
        return j + 1u == cap ? 0u : j + 1u;
    }
    void destroy_slots(tab &t) noexcept
    {
        if constexpr (!trivial_dtor) {
            if (t.slots != nullptr && m_walk) {
                for (std::size_t i = 0; i < t.cap; ++i) {
                    if (t.full[i]) {
                        t.slots[i].~value_type();
                    }
                }
            }
        }
    }
    // move table t into a block of its own with room for `want` terms
    void grow(tab &t, std::size_t want)
    {
        if (t.cap != 0u && want * 10u <= t.cap * 7u) {
            return;
        }
        std::size_t cap = std::max(cap_for(want), t.cap * 2u);
        tab nt;
        nt.own = raw_block(slot_bytes(cap));
        bind(nt, nt.own.ptr, cap);
        for (std::size_t i = 0; i < t.cap; ++i) {
            if (t.full[i]) {
                std::size_t j = slot_of(key_hasher<K>{}(t.slots[i].first), cap, m_log2);
                while (nt.full[j]) {
                    j = next(j, cap);
                }
                nt.full[j] = 1;
                ::new (static_cast<void *>(nt.slots + j)) value_type(std::move(t.slots[i]));
                t.slots[i].~value_type();
                t.full[i] = 0;
            }
        }
        nt.n = t.n;
        // the old block goes (a slab part simply stays unused until the slab does)
        t = std::move(nt);
    }
    void release() noexcept
    {
        for (auto &t : m_tabs) {
            destroy_slots(t);
        }
        m_tabs.clear();
        m_slab.reset();
    }
    void copy_from(const seg_table &o)
    {
        // same geometry, every term at the same slot
        m_tabs = std::vector<tab>(o.m_tabs.size());
        m_log2 = o.m_log2;
        m_size = o.m_size;
        m_walk = !trivial_dtor;
        std::vector<std::size_t> off(o.m_tabs.size() + 1u, 0);
        for (std::size_t i = 0; i < o.m_tabs.size(); ++i) {
            off[i + 1u] = off[i] + (o.m_tabs[i].cap ? slot_bytes(o.m_tabs[i].cap) : 0u);
        }
        m_slab = raw_block(off.back());
        for (std::size_t i = 0; i < o.m_tabs.size(); ++i) {
            const tab &s = o.m_tabs[i];
            if (s.cap == 0u) {
                continue;
            }
            tab &t = m_tabs[i];
            bind(t, static_cast<unsigned char *>(m_slab.ptr) + off[i], s.cap);
            for (std::size_t j = 0; j < s.cap; ++j) {
                if (s.full[j]) {
                    ::new (static_cast<void *>(t.slots + j)) value_type(s.slots[j]);
                    t.full[j] = 1;
                }
            }
            t.n = s.n;
        }
    }
    template <bool Const>
    class iter
    {
        friend class seg_table;
        using owner_t = std::conditional_t<Const, const seg_table, seg_table>;
        owner_t *m_o = nullptr;
        std::size_t m_t = 0, m_i = 0;
        void settle()
        {
            while (m_t < m_o->m_tabs.size()) {
                const auto &t = m_o->m_tabs[m_t];
                while (m_i < t.cap && !t.full[m_i]) {
                    ++m_i;
                }
                if (m_i < t.cap) {
                    return;
                }
                ++m_t;
                m_i = 0;
            }
        }

    public:
        using iterator_category = std::forward_iterator_tag;
        using value_type = typename seg_table::value_type;
        using difference_type = std::ptrdiff_t;
        using reference = std::conditional_t<Const, const value_type &, value_type &>;
        using pointer = std::conditional_t<Const, const value_type *, value_type *>;
        iter() = default;
        iter(owner_t *o, std::size_t t, std::size_t i) : m_o(o), m_t(t), m_i(i) {}
        template <bool B = Const, std::enable_if_t<B, int> = 0>
        iter(const iter<false> &o) : m_o(o.m_o), m_t(o.m_t), m_i(o.m_i)
        {
        }
        reference operator*() const
        {
            return m_o->m_tabs[m_t].slots[m_i];
        }
        pointer operator->() const
        {
            return &m_o->m_tabs[m_t].slots[m_i];
        }
        iter &operator++()
        {
            ++m_i;
            settle();
            return *this;
        }
        iter operator++(int)
        {
            iter c(*this);
            ++*this;
            return c;
        }
        friend bool operator==(const iter &a, const iter &b)
        {
            return a.m_t == b.m_t && a.m_i == b.m_i;
        }
        friend bool operator!=(const iter &a, const iter &b)
        {
            return !(a == b);
        }
    };

public:
    using iterator = iter<false>;
    using const_iterator = iter<true>;
    seg_table() = default;
    seg_table(const seg_table &o)
    {
        copy_from(o);
    }
    seg_table(seg_table &&o) noexcept
        : m_tabs(std::move(o.m_tabs)), m_slab(std::move(o.m_slab)), m_log2(o.m_log2), m_size(o.m_size), m_walk(o.m_walk)
    {
        o.m_tabs = std::vector<tab>(1);
        o.m_log2 = 0;
        o.m_size = 0;
        o.m_walk = !trivial_dtor;
    }
    seg_table &operator=(const seg_table &o)
    {
        if (this != &o) {
            seg_table tmp(o);
            *this = std::move(tmp);
        }
        return *this;
    }
    seg_table &operator=(seg_table &&o) noexcept
    {
        if (this != &o) {
            release();
            m_tabs = std::move(o.m_tabs);
            m_slab = std::move(o.m_slab);
            m_log2 = o.m_log2;
            m_size = o.m_size;
            m_walk = o.m_walk;
            o.m_tabs = std::vector<tab>(1);
            o.m_log2 = 0;
            o.m_size = 0;
            o.m_walk = !trivial_dtor;
        }
        return *this;
    }
    ~seg_table()
    {
        release();
    }
    std::size_t size() const
    {
        return m_size;
    }
    bool empty() const
    {
        return m_size == 0u;
    }
    unsigned log2_tables() const
    {
        return m_log2;
    }
    // terms in table i, number of tables, bytes held (series::table_stats, series.hpp:1383-1408)
    std::size_t n_tables() const
    {
        return m_tabs.size();
    }
    std::size_t table_size(std::size_t i) const
    {
        return m_tabs[i].n;
    }
    std::size_t byte_size() const
    {
        std::size_t b = sizeof(*this) + m_tabs.capacity() * sizeof(tab) + m_slab.bytes;
        for (const auto &t : m_tabs) {
            b += t.own.bytes;
        }
        return b;
    }
    // Indexed enumeration for the marshalling of an operand (polynomial.hpp:828-833 copies the terms into vectors
    // before every product): the terms are numbered 0 .. size()-1 in (table, slot) order, and the numbering is known
    // per CHUNK (a whole table, or a slot range of a large one) before any term is visited, so that several threads
    // can each write their part of the SoA arrays.
    struct enum_chunk {
        std::size_t ti, b, e, first;
    };
    std::vector<enum_chunk> enum_plan() const
    {
        constexpr std::size_t CH = std::size_t(1) << 15;
        std::vector<enum_chunk> ch;
        std::size_t idx = 0;
        for (std::size_t ti = 0; ti < m_tabs.size(); ++ti) {
            const tab &t = m_tabs[ti];
            if (t.n == 0u) {
                continue;
            }
            if (t.cap <= CH) {
                ch.push_back({ti, 0, t.cap, idx});
                idx += t.n;
                continue;
            }
            for (std::size_t b = 0; b < t.cap; b += CH) {
                const std::size_t e = std::min(b + CH, t.cap);
                std::size_t cnt = 0;
                for (std::size_t j = b; j < e; ++j) {
                    cnt += t.full[j];
                }
                if (cnt) {
                    ch.push_back({ti, b, e, idx});
                    idx += cnt;
                }
            }
        }
        return ch;
    }
    // share w of th of a plan: fn(index, term) for every term of the chunks whose first index falls in
    // [size * w / th, size * (w + 1) / th)
    template <typename Fn>
    void enum_run(const std::vector<enum_chunk> &ch, std::size_t w, std::size_t th, const Fn &fn) const
    {
        const std::size_t lo = static_cast<std::size_t>(static_cast<unsigned __int128>(m_size) * w / th),
                          hi = static_cast<std::size_t>(static_cast<unsigned __int128>(m_size) * (w + 1u) / th);
        auto first_ge = [&](std::size_t v) {
            return static_cast<std::size_t>(
                std::lower_bound(ch.begin(), ch.end(), v, [](const enum_chunk &c, std::size_t x) { return c.first < x; })
                - ch.begin());
        };
        const std::size_t c0 = w == 0u ? 0u : first_ge(lo), c1 = w + 1u == th ? ch.size() : first_ge(hi);
        for (std::size_t c = c0; c < c1; ++c) {
            const tab &t = m_tabs[ch[c].ti];
            std::size_t i = ch[c].first;
            for (std::size_t j = ch[c].b; j < ch[c].e; ++j) {
                if (t.full[j]) {
                    fn(i++, static_cast<const value_type &>(t.slots[j]));
                }
            }
        }
    }
    // slots allocated over all tables (table_stats-style diagnostics and tests)
    std::size_t capacity() const
    {
        std::size_t c = 0;
        for (const auto &t : m_tabs) {
            c += t.cap;
        }
        return c;
    }
    const_iterator begin() const
    {
        const_iterator it(this, 0, 0);
        it.settle();
        return it;
    }
    const_iterator end() const
    {
        return const_iterator(this, m_tabs.size(), 0);
    }
    iterator begin()
    {
        m_walk = !trivial_dtor; // the caller may store anything through a mutable iterator
        iterator it(this, 0, 0);
        it.settle();
        return it;
    }
    iterator end()
    {
        return iterator(this, m_tabs.size(), 0);
    }
    void clear()
    {
        release();
        m_tabs = std::vector<tab>(1);
        m_log2 = 0;
        m_size = 0;
        m_walk = !trivial_dtor;
    }
    void reserve(std::size_t n)
    {
        if (m_log2 == 0u) {
            grow(m_tabs[0], n);
        }
    }

private:
    template <typename Self>
    static auto find_impl(Self &self, const K &k)
    {
        using it_t = std::conditional_t<std::is_const_v<Self>, const_iterator, iterator>;
        const std::size_t h = key_hasher<K>{}(k), ti = h & ((std::size_t(1) << self.m_log2) - 1u);
        auto &t = self.m_tabs[ti];
        if (t.cap != 0u) {
            for (std::size_t j = slot_of(h, t.cap, self.m_log2); t.full[j]; j = next(j, t.cap)) {
                if (t.slots[j].first == k) {
                    return it_t(&self, ti, j);
                }
            }
        }
        return it_t(&self, self.m_tabs.size(), 0);
    }

public:
    const_iterator find(const K &k) const
    {
        return find_impl(*this, k);
    }
    iterator find(const K &k)
    {
        m_walk = !trivial_dtor;
        return find_impl(*this, k);
    }
    std::pair<iterator, bool> try_emplace(const K &k, const C &c)
    {
        m_walk = !trivial_dtor;
        const std::size_t h = key_hasher<K>{}(k), ti = h & ((std::size_t(1) << m_log2) - 1u);
        auto &t = m_tabs[ti];
        grow(t, t.n + 1u);
        const std::size_t cap = t.cap;
        std::size_t j = slot_of(h, cap, m_log2);
        for (; t.full[j]; j = next(j, cap)) {
            if (t.slots[j].first == k) {
                return {iterator(this, ti, j), false};
            }
        }
        ::new (static_cast<void *>(t.slots + j)) value_type(k, c);
        t.full[j] = 1;
        ++t.n;
        ++m_size;
        return {iterator(this, ti, j), true};
    }
    void emplace(K k, C c)
    {
        try_emplace(k, c);
    }
    void erase(iterator it)
    {
        // backward-shift deletion: no tombstones
        auto &t = m_tabs[it.m_t];
        const std::size_t cap = t.cap;
        std::size_t i = it.m_i, j = i;
        for (;;) {
            j = next(j, cap);
            if (!t.full[j]) {
                break;
            }
            const std::size_t home = slot_of(key_hasher<K>{}(t.slots[j].first), cap, m_log2);
            // can the entry at j move to the hole at i?  only if its home is not cyclically inside (i, j]
            const bool between = i <= j ? (home > i && home <= j) : (home > i || home <= j);
            if (!between) {
                t.slots[i] = std::move(t.slots[j]);
                i = j;
            }
        }
        t.slots[i].~value_type(); // the erased term, or the moved-from husk the last shift left behind
        t.full[i] = 0;
        --t.n;
        --m_size;
    }
    // Adopt `n` distinct, non-zero terms that arrive grouped by hash & (2^log2 - 1) (offs[2^log2 + 1]); make(i) builds
    // term i.  One block holds every table; up to `nthreads` threads fill them, each a contiguous range of tables
    // with about the same number of terms (so a thread also faults in a contiguous part of the block).
    template <typename Make>
    void adopt_grouped(std::size_t n, unsigned log2, const std::uint64_t *offs, const Make &make, unsigned nthreads)
    {
        const std::size_t nt = std::size_t(1) << log2;
        release();
        m_tabs = std::vector<tab>(nt);
        m_log2 = log2;
        m_size = n;
        std::vector<std::size_t> boff(nt + 1u, 0);
        for (std::size_t ti = 0; ti < nt; ++ti) {
            const std::size_t cnt = static_cast<std::size_t>(offs[ti + 1u] - offs[ti]);
            boff[ti + 1u] = boff[ti] + (cnt ? slot_bytes(cap_for(cnt)) : 0u);
        }
        m_slab = raw_block(boff[nt]);
        const std::size_t th = std::max<std::size_t>(1, std::min<std::size_t>({nthreads, nt, n / 8192u + 1u}));
        std::vector<unsigned char> resource(th, 0);
        auto work = [&](std::size_t w) {
            // tables [t0, t1): the w-th share of the terms
            const std::uint64_t lo = static_cast<std::uint64_t>(n) * w / th, hi = static_cast<std::uint64_t>(n) * (w + 1u) / th;
            const std::size_t t0 = w == 0u ? 0u : static_cast<std::size_t>(std::lower_bound(offs, offs + nt, lo) - offs);
            const std::size_t t1 = w + 1u == th ? nt : static_cast<std::size_t>(std::lower_bound(offs, offs + nt, hi) - offs);
            bool res = false;
            for (std::size_t ti = t0; ti < t1; ++ti) {
                auto &t = m_tabs[ti];
                const std::size_t b = static_cast<std::size_t>(offs[ti]), e = static_cast<std::size_t>(offs[ti + 1u]);
                if (b == e) {
                    continue;
                }
                const std::size_t cap = cap_for(e - b);
                bind(t, static_cast<unsigned char *>(m_slab.ptr) + boff[ti], cap);
                for (std::size_t i = b; i < e; ++i) {
                    value_type v = make(i);
                    res = res || !holds_no_resource(v.first) || !holds_no_resource(v.second);
                    std::size_t j = slot_of(key_hasher<K>{}(v.first), cap, log2);
                    while (t.full[j]) {
                        j = next(j, cap);
                    }
                    ::new (static_cast<void *>(t.slots + j)) value_type(std::move(v));
                    t.full[j] = 1;
                }
                t.n = e - b;
            }
            resource[w] = res ? 1 : 0;
        };
        if (th == 1u) {
            work(0);
        } else {
            std::vector<std::thread> pool;
            pool.reserve(th - 1u);
            for (std::size_t w = 1; w < th; ++w) {
                pool.emplace_back(work, w);
            }
            work(0);
            for (auto &p : pool) {
                p.join();
            }
        }
        m_walk = false;
        for (auto r : resource) {
            m_walk = m_walk || r != 0;
        }
    }
};
} // namespace detail

template <typename K, typename C>
class polynomial
{
public:
    using key_type = K;
    using cf_type = C;
    using table_type = detail::seg_table<K, C>;

private:
    symbol_set m_ss;
    table_type m_terms;

public:
    polynomial() = default;
    template <typename S, std::enable_if_t<std::is_arithmetic_v<S> || std::is_same_v<S, integer>, int> = 0>
    explicit polynomial(const S &c)
    {
        C cc = [&] {
            if constexpr (std::is_same_v<S, integer> && std::is_same_v<C, double>) {
                return static_cast<double>(c);
            } else {
                return C(c);
            }
        }();
        if (!is_zero(cc)) {
            m_terms.emplace(K(m_ss), std::move(cc));
        }
    }
    const symbol_set &get_symbol_set() const
    {
        return m_ss;
    }
    void set_symbol_set(const symbol_set &s)
    {
        if (!m_terms.empty()) {
            throw std::invalid_argument("A symbol set can be set only in an empty series");
        }
        m_ss = s;
    }
    std::size_t size() const
    {
        return m_terms.size();
    }
    bool empty() const
    {
        return m_terms.empty();
    }
    auto begin() const
    {
        return m_terms.begin();
    }
    auto end() const
    {
        return m_terms.end();
    }
    void clear()
    {
        m_terms.clear();
    }
    void reserve(std::size_t n)
    {
        m_terms.reserve(n);
    }
    const table_type &_terms() const
    {
        return m_terms;
    }
    // series::table_stats (series.hpp:1383-1408), the summary the reference's benchmark drivers print after the
    // product (benchmark/dense.hpp:41, sparse.hpp:44): same lines, same labels.  "Total size in bytes" counts this
    // container's storage (inline-limb coefficients included; heap limbs of coefficients beyond three limbs are not
    // walked), where the reference reports its own byte_size().
    std::string table_stats() const
    {
        std::ostringstream oss;
        const auto s = m_terms.size();
        const auto ntables = m_terms.n_tables();
        oss << "Total number of terms             : " << s << '\n';
        oss << "Total number of tables            : " << ntables << '\n';
        if (s != 0u) {
            oss << "Average terms per table           : " << static_cast<double>(s) / static_cast<double>(ntables) << '\n';
            std::size_t mn = m_terms.table_size(0), mx = mn;
            for (std::size_t i = 1; i < ntables; ++i) {
                mn = std::min(mn, m_terms.table_size(i));
                mx = std::max(mx, m_terms.table_size(i));
            }
            oss << "Min/max terms per table           : " << mn << '/' << mx << '\n';
        }
        oss << "Total size in bytes               : " << sizeof(*this) - sizeof(m_terms) + m_terms.byte_size() << '\n';
        return oss.str();
    }
    auto find(const K &k) const
    {
        return m_terms.find(k);
    }
    // add_term<Sign>: accumulate, erase on zero (series.hpp add_term semantics)
    template <bool Sign = true>
    void add_term(const K &k, const C &c)
    {
        if (!key_is_compatible(k, m_ss)) {
            throw std::invalid_argument("Cannot add a term to a series: the term's key is not compatible with the "
                                        "series' symbol set");
        }
        auto [it, inserted] = m_terms.try_emplace(k, c);
        if (inserted) {
            if constexpr (!Sign) {
                it->second = C(0) - it->second;
            }
        } else {
            if constexpr (Sign) {
                it->second = it->second + c;
            } else {
                it->second = it->second - c;
            }
        }
        if (is_zero(it->second)) {
            m_terms.erase(it);
        }
    }
    // unchecked insertion of a term known to be new and non-zero (engine results)
    void _insert_unique(K k, C c)
    {
        m_terms.emplace(std::move(k), std::move(c));
    }
    // a whole product from the engine: terms grouped by hash & (2^log2 - 1), distinct and non-zero
    template <typename Make>
    void _adopt_grouped(std::size_t n, unsigned log2, const std::uint64_t *offs, const Make &make, unsigned nthreads = 0)
    {
        m_terms.adopt_grouped(n, log2, offs, make, nthreads ? nthreads : std::max(1u, std::thread::hardware_concurrency()));
    }
    auto begin()
    {
        return m_terms.begin();
    }
    auto end()
    {
        return m_terms.end();
    }
};

namespace detail
{
template <typename K, typename C>
struct is_poly<polynomial<K, C>> : std::true_type {
};

// Re-express p over the larger symbol set `u` (series_sym_extender, series.hpp:539-625): zero exponents
// are inserted for the new symbols and every key is re-packed.
template <typename K, typename C>
inline polynomial<K, C> extend_symbols(const polynomial<K, C> &p, const symbol_set &u, const std::vector<std::size_t> &pos)
{
    using T = typename K::value_type;
    polynomial<K, C> r;
    r.set_symbol_set(u);
    r.reserve(p.size());
    const auto n_old = p.get_symbol_set().size();
    std::vector<T> e(u.size());
    for (const auto &[k, c] : p) {
        std::fill(e.begin(), e.end(), T(0));
        const auto old = k.unpack(n_old);
        for (std::size_t i = 0; i < n_old; ++i) {
            e[pos[i]] = old[i];
        }
        r._insert_unique(K(e.begin(), e.size()), c);
    }
    return r;
}

template <typename P>
inline void unify(const P &x, const P &y, P &xs, P &ys, const P *&px, const P *&py)
{
    if (x.get_symbol_set() == y.get_symbol_set()) {
        px = &x;
        py = &y;
        return;
    }
    std::vector<std::size_t> pa, pb;
    const auto u = merge_symbol_sets(x.get_symbol_set(), y.get_symbol_set(), pa, pb);
    if (u == x.get_symbol_set()) {
        px = &x;
    } else {
        xs = extend_symbols(x, u, pa);
        px = &xs;
    }
    if (u == y.get_symbol_set()) {
        py = &y;
    } else {
        ys = extend_symbols(y, u, pb);
        py = &ys;
    }
}

struct trunc_args {
    bool on = false;
    bool reject_all = false;
    std::int64_t limit = 0;
    bool partial = false;
    std::vector<unsigned> idx;
};

// Reduce the truncation limit V to the key's integer type T with the conversions the reference's comparison
// `max_deg < d_i + d_j` performs (polynomial.hpp:604-610): built-in integers follow the usual arithmetic
// conversions (a negative int against an unsigned T wraps), multiprecision integers compare by value.
template <typename T, typename V>
inline void reduce_limit(trunc_args &ta, const V &lim)
{
    ta.on = true;
    if constexpr (std::is_same_v<V, integer>) {
        if constexpr (std::is_signed_v<T>) {
            if (lim < integer(std::numeric_limits<std::int64_t>::min())) {
                ta.reject_all = true;
            } else if (lim > integer(std::numeric_limits<std::int64_t>::max())) {
                ta.limit = std::numeric_limits<std::int64_t>::max();
            } else {
                ta.limit = static_cast<std::int64_t>(lim);
            }
        } else {
            if (lim.sgn() < 0) {
                ta.reject_all = true;
            } else if (lim.nbits() > 64u) {
                ta.limit = -1; // UINT64_MAX
            } else {
                std::uint64_t l[2] = {0, 0};
                lim.to_limbs(l, 2);
                ta.limit = static_cast<std::int64_t>(l[0]);
            }
        }
    } else {
        static_assert(std::is_integral_v<V>, "the truncation limit must be an integer");
        if constexpr (std::is_signed_v<T>) {
            if constexpr (std::is_unsigned_v<V> && sizeof(V) >= sizeof(std::int64_t)) {
                ta.limit = lim > static_cast<V>(std::numeric_limits<std::int64_t>::max())
                               ? std::numeric_limits<std::int64_t>::max()
                               : static_cast<std::int64_t>(lim);
            } else {
                ta.limit = static_cast<std::int64_t>(lim);
            }
        } else {
            // T unsigned: int -> unsigned conversion (wraps for negative values), 64-bit bit pattern
            if constexpr (std::is_signed_v<V>) {
                if constexpr (sizeof(T) < 8) {
                    // 32-bit unsigned T: int converts to uint32 when ranks are equal
                    ta.limit = static_cast<std::int64_t>(static_cast<std::uint64_t>(static_cast<std::uint32_t>(lim)));
                    if (sizeof(V) > sizeof(T)) {
                        // wider signed V: the comparison happens in V; a negative limit rejects everything
                        if (lim < 0) {
                            ta.reject_all = true;
                        } else {
                            ta.limit = static_cast<std::int64_t>(lim);
                        }
                    }
                } else {
                    ta.limit = static_cast<std::int64_t>(static_cast<std::uint64_t>(static_cast<std::int64_t>(lim)));
                }
            } else {
                ta.limit = static_cast<std::int64_t>(static_cast<std::uint64_t>(lim));
            }
        }
    }
}

// The operands of a product as SoA arrays for the C ABI: keys[n][W], cfs[n][lin] (two's-complement limbs, or the bit
// pattern of a double), lin = the widest coefficient of EITHER operand (the engine takes one limb count per call).
// Returns lin.  What the reference does at polynomial.hpp:828-833 (copy the terms into vectors) -- here it is the
// first thing a caller of operator* waits for, so large operands are marshalled by several threads in ONE round of
// two phases with a barrier between them: every thread first finds the widest coefficient among its chunks of both
// operands (seg_table::enum_plan), the barrier's completion step sizes the coefficient arrays for the limb count
// now known, then every thread writes its chunks.  Small operands take the serial form of the same two passes.
// `nthreads` = 0: all hardware threads.
template <typename R, typename K, typename C0, typename C1>
inline unsigned marshal_operands(const polynomial<K, C0> &x, const polynomial<K, C1> &y, unsigned W, std::size_t nvars,
                                 std::vector<std::uint64_t> &kx, std::vector<std::uint64_t> &cx, std::vector<std::uint64_t> &ky,
                                 std::vector<std::uint64_t> &cy, unsigned nthreads = 0)
{
    using RT = b200::cf_traits<R>;
    constexpr bool is_int = RT::kind == OBK_CF_INT;
    auto needed = [](const auto &c) -> unsigned {
        using CP = std::decay_t<decltype(c)>;
        if constexpr (is_int) {
            return b200::cf_traits<CP>::limbs_needed(c);
        } else {
            return 1u;
        }
    };
    auto store = [](const auto &c, std::uint64_t *out, unsigned n) {
        using CP = std::decay_t<decltype(c)>;
        if constexpr (std::is_same_v<CP, R>) {
            RT::store(c, out, n);
        } else {
            RT::store(static_cast<R>(c), out, n); // mixed coefficient types: convert to the result type on the host
        }
    };
    const std::size_t nx = x.size(), ny = y.size();
    kx.resize(nx * W);
    ky.resize(ny * W);
    const std::size_t hw = nthreads ? nthreads : std::max(1u, std::thread::hardware_concurrency());
    const std::size_t th = std::min<std::size_t>(hw, (nx + ny) / 16384u + 1u);
    if (th > 1u) {
        const auto px = x._terms().enum_plan(), py = y._terms().enum_plan();
        std::atomic<unsigned> widest{1u};
        unsigned lin = 1;
        bool failed = false;
        auto sized = [&]() noexcept {
            try {
                lin = widest.load();
                cx.resize(nx * lin);
                cy.resize(ny * lin);
            } catch (...) {
                failed = true;
            }
        };
        std::barrier sync(static_cast<std::ptrdiff_t>(th), sized);
        parallel_for(th, [&](std::size_t w) {
            if constexpr (is_int) {
                unsigned mine = 1;
                auto scan = [&](std::size_t, const auto &term) { mine = std::max(mine, needed(term.second)); };
                x._terms().enum_run(px, w, th, scan);
                y._terms().enum_run(py, w, th, scan);
                unsigned cur = widest.load(std::memory_order_relaxed);
                while (mine > cur && !widest.compare_exchange_weak(cur, mine, std::memory_order_relaxed)) {
                }
            }
            sync.arrive_and_wait(); // `sized` has run: lin, cx and cy are final
            if (failed) {
                return;
            }
            const unsigned l = lin;
            auto put = [&](std::vector<std::uint64_t> &keys, std::vector<std::uint64_t> &cfs) {
                return [&keys, &cfs, &store, W, nvars, l](std::size_t i, const auto &term) {
                    term.first.to_words(&keys[i * W], nvars);
                    store(term.second, &cfs[i * l], l);
                };
            };
            x._terms().enum_run(px, w, th, put(kx, cx));
            y._terms().enum_run(py, w, th, put(ky, cy));
        });
        if (failed) {
            throw std::bad_alloc();
        }
        return lin;
    }
    unsigned lin = 1;
    if constexpr (is_int) {
        for (const auto &t : x) {
            lin = std::max(lin, needed(t.second));
        }
        for (const auto &t : y) {
            lin = std::max(lin, needed(t.second));
        }
    }
    auto serial = [&](const auto &p, std::vector<std::uint64_t> &keys, std::vector<std::uint64_t> &cfs) {
        cfs.resize(p.size() * lin);
        std::size_t i = 0;
        for (const auto &t : p) {
            t.first.to_words(&keys[i * W], nvars);
            store(t.second, &cfs[i * lin], lin);
            ++i;
        }
    };
    serial(x, kx, cx);
    serial(y, ky, cy);
    return lin;
}

// poly_mul_impl_identical_ss: both operands share the symbol set.
template <typename K, typename C0, typename C1, typename R = decltype(std::declval<C0>() * std::declval<C1>())>
inline polynomial<K, R> poly_mul_impl_identical_ss(const polynomial<K, C0> &x, const polynomial<K, C1> &y,
                                                   const trunc_args &ta)
{
    using T = typename K::value_type;
    polynomial<K, R> ret;
    ret.set_symbol_set(x.get_symbol_set());
    // empty operand -> empty result (polynomial.hpp:1892-1895)
    if (x.empty() || y.empty() || ta.reject_all) {
        return ret;
    }
    const auto &ss = x.get_symbol_set();
    const std::size_t nvars = ss.size();
    const unsigned W = K::n_words(nvars);
    using RT = b200::cf_traits<R>;

    // Marshal the terms to SoA (what the reference copies into vectors at :828-833).  The four arrays are kept per
    // calling thread from one product to the next (fresh vectors of this size are fresh pages: faulting them in cost
    // more than filling them); an array that grew beyond 64 MB is given back afterwards.
    static thread_local std::vector<std::uint64_t> kx, cx, ky, cy;
    struct scratch_guard {
        ~scratch_guard()
        {
            for (auto *v : {&kx, &cx, &ky, &cy}) {
                if (v->capacity() > (std::size_t(64) << 20) / 8u) {
                    std::vector<std::uint64_t>().swap(*v);
                }
            }
        }
    } guard;
    const unsigned lin = marshal_operands<R>(x, y, W, nvars, kx, cx, ky, cy);
    auto view = [&](std::size_t n, const std::vector<std::uint64_t> &k, const std::vector<std::uint64_t> &c) {
        obk_poly_view v{};
        v.nterms = n;
        v.key_words = W;
        v.key_signed = std::is_signed_v<T> ? 1u : 0u;
        v.nvars = static_cast<std::uint32_t>(nvars);
        v.psize = K::psize;
        v.cf_kind = RT::kind;
        v.cf_limbs = lin;
        v.mem = OBK_MEM_HOST;
        v.key_bits = sizeof(T) * 8u;
        v.keys = k.data();
        v.cfs = c.data();
        return v;
    };
    const obk_poly_view vx = view(x.size(), kx, cx), vy = view(y.size(), ky, cy);
    obk_trunc tr{};
    if (ta.on) {
        tr.limit = ta.limit;
        tr.partial = ta.partial ? 1u : 0u;
        tr.nidx = static_cast<std::uint32_t>(ta.idx.size());
        tr.idx = ta.idx.data();
    }
    auto &eng = b200::engine::instance();
    obk_poly_result res{};
    {
        std::lock_guard<std::mutex> g(eng.mutex());
        eng.check(::obk_poly_mul(eng.handle(), &vx, &vy, ta.on ? &tr : nullptr, OBK_MEM_HOST, &res, nullptr));
    }
    // Rebuild the host container (terms arrive grouped by table, without zeros or duplicates): table by table, in
    // parallel when the engine grouped them by hash & (2^k - 1) (it always does for host results).
    try {
        const auto *rc = static_cast<const std::uint64_t *>(res.cfs);
        auto make = [&](std::size_t i) {
            return std::pair<K, R>(K::from_words(res.keys + i * W, nvars), RT::load(rc + i * res.cf_limbs, res.cf_limbs));
        };
        if (res.seg_kind == 0u && res.seg_offsets != nullptr && res.nsegs == (std::uint64_t(1) << res.log2_nsegs)) {
            ret._adopt_grouped(static_cast<std::size_t>(res.nterms), res.log2_nsegs, res.seg_offsets, make);
        } else {
            ret.reserve(res.nterms);
            for (std::uint64_t i = 0; i < res.nterms; ++i) {
                auto t = make(i);
                ret._insert_unique(std::move(t.first), std::move(t.second));
            }
        }
    } catch (...) {
        ::obk_result_free(eng.handle(), &res);
        ret.clear(); // polynomial.hpp:1575-1580
        throw;
    }
    ::obk_result_free(eng.handle(), &res);
    return ret;
}

// poly_mul_impl + poly_mul_impl_switch
template <typename K, typename C0, typename C1>
inline auto poly_mul_impl_switch(const polynomial<K, C0> &x, const polynomial<K, C1> &y, const trunc_args &ta)
{
    if constexpr (std::is_same_v<C0, C1>) {
        polynomial<K, C0> xs, ys;
        const polynomial<K, C0> *px, *py;
        unify(x, y, xs, ys, px, py);
        return poly_mul_impl_identical_ss(*px, *py, ta);
    } else {
        if (x.get_symbol_set() == y.get_symbol_set()) {
            return poly_mul_impl_identical_ss(x, y, ta);
        }
        std::vector<std::size_t> pa, pb;
        const auto u = merge_symbol_sets(x.get_symbol_set(), y.get_symbol_set(), pa, pb);
        return poly_mul_impl_identical_ss(extend_symbols(x, u, pa), extend_symbols(y, u, pb), ta);
    }
}

} // namespace detail

// ---- the public product entry points ---------------------------------------------------------------------
namespace polynomials
{
template <typename K, typename C0, typename C1>
inline auto series_mul(const polynomial<K, C0> &x, const polynomial<K, C1> &y)
{
    return detail::poly_mul_impl_switch(x, y, detail::trunc_args{});
}

template <typename K, typename C0, typename C1, typename V>
inline auto truncated_mul(const polynomial<K, C0> &x, const polynomial<K, C1> &y, const V &max_degree)
{
    detail::trunc_args ta;
    detail::reduce_limit<typename K::value_type>(ta, max_degree);
    return detail::poly_mul_impl_switch(x, y, ta);
}

template <typename K, typename C0, typename C1, typename V>
inline auto truncated_mul(const polynomial<K, C0> &x, const polynomial<K, C1> &y, const V &max_degree, const symbol_set &s)
{
    detail::trunc_args ta;
    detail::reduce_limit<typename K::value_type>(ta, max_degree);
    ta.partial = true;
    // positions in the MERGED symbol set of the symbols of s that are present
    std::vector<std::size_t> pa, pb;
    const auto u = merge_symbol_sets(x.get_symbol_set(), y.get_symbol_set(), pa, pb);
    ta.idx = ss_intersect_idx(s, u);
    return detail::poly_mul_impl_switch(x, y, ta);
}
} // namespace polynomials

using polynomials::series_mul;
using polynomials::truncated_mul;

template <typename K, typename C0, typename C1>
inline auto operator*(const polynomial<K, C0> &x, const polynomial<K, C1> &y)
{
    return polynomials::series_mul(x, y);
}

// series x scalar (series_default_mul_impl, series.hpp:3098-3182) -- host arithmetic, outside the path
template <typename K, typename C, typename S, std::enable_if_t<std::is_arithmetic_v<S> || std::is_same_v<S, integer>, int> = 0>
inline polynomial<K, C> operator*(const polynomial<K, C> &x, const S &s)
{
    polynomial<K, C> r;
    r.set_symbol_set(x.get_symbol_set());
    const C sc = [&] {
        if constexpr (std::is_same_v<S, integer> && std::is_same_v<C, double>) {
            return static_cast<double>(s);
        } else {
            return C(s);
        }
    }();
    if (is_zero(sc)) {
        return r;
    }
    r.reserve(x.size());
    for (const auto &[k, c] : x) {
        C v = c * sc;
        if (!is_zero(v)) {
            r._insert_unique(k, std::move(v));
        }
    }
    return r;
}
template <typename K, typename C, typename S, std::enable_if_t<std::is_arithmetic_v<S> || std::is_same_v<S, integer>, int> = 0>
inline polynomial<K, C> operator*(const S &s, const polynomial<K, C> &x)
{
    return x * s;
}
template <typename K, typename C, typename U>
inline polynomial<K, C> &operator*=(polynomial<K, C> &x, const U &y)
{
    x = x * y; // series.hpp:3200-3207
    return x;
}

// ---- addition / subtraction (series.hpp:2195-2991; host bookkeeping used to build operands) ---------------
namespace detail
{
template <bool Sign, typename K, typename C>
inline polynomial<K, C> poly_addsub(const polynomial<K, C> &x, const polynomial<K, C> &y)
{
    polynomial<K, C> xs, ys;
    const polynomial<K, C> *px, *py;
    unify(x, y, xs, ys, px, py);
    polynomial<K, C> r(*px);
    for (const auto &[k, c] : *py) {
        r.template add_term<Sign>(k, c);
    }
    return r;
}
} // namespace detail

template <typename K, typename C>
inline polynomial<K, C> operator+(const polynomial<K, C> &x, const polynomial<K, C> &y)
{
    return detail::poly_addsub<true>(x, y);
}
template <typename K, typename C>
inline polynomial<K, C> operator-(const polynomial<K, C> &x, const polynomial<K, C> &y)
{
    return detail::poly_addsub<false>(x, y);
}
template <typename K, typename C, typename S, std::enable_if_t<std::is_arithmetic_v<S> || std::is_same_v<S, integer>, int> = 0>
inline polynomial<K, C> operator+(const polynomial<K, C> &x, const S &s)
{
    polynomial<K, C> c(s);
    return x + c;
}
template <typename K, typename C, typename S, std::enable_if_t<std::is_arithmetic_v<S> || std::is_same_v<S, integer>, int> = 0>
inline polynomial<K, C> operator+(const S &s, const polynomial<K, C> &x)
{
    return x + s;
}
template <typename K, typename C, typename S, std::enable_if_t<std::is_arithmetic_v<S> || std::is_same_v<S, integer>, int> = 0>
inline polynomial<K, C> operator-(const polynomial<K, C> &x, const S &s)
{
    polynomial<K, C> c(s);
    return x - c;
}
template <typename K, typename C, typename S, std::enable_if_t<std::is_arithmetic_v<S> || std::is_same_v<S, integer>, int> = 0>
inline polynomial<K, C> operator-(const S &s, const polynomial<K, C> &x)
{
    polynomial<K, C> c(s);
    return c - x;
}
template <typename K, typename C>
inline polynomial<K, C> operator-(const polynomial<K, C> &x)
{
    return polynomial<K, C>() - x;
}
template <typename K, typename C, typename U>
inline polynomial<K, C> &operator+=(polynomial<K, C> &x, const U &y)
{
    x = x + y;
    return x;
}
template <typename K, typename C, typename U>
inline polynomial<K, C> &operator-=(polynomial<K, C> &x, const U &y)
{
    x = x - y;
    return x;
}

template <typename K, typename C>
inline bool operator==(const polynomial<K, C> &x, const polynomial<K, C> &y)
{
    polynomial<K, C> xs, ys;
    const polynomial<K, C> *px, *py;
    detail::unify(x, y, xs, ys, px, py);
    if (px->size() != py->size()) {
        return false;
    }
    for (const auto &[k, c] : *px) {
        const auto it = py->find(k);
        if (it == py->end() || !(it->second == c)) {
            return false;
        }
    }
    return true;
}
template <typename K, typename C>
inline bool operator!=(const polynomial<K, C> &x, const polynomial<K, C> &y)
{
    return !(x == y);
}
template <typename K, typename C, typename S, std::enable_if_t<std::is_arithmetic_v<S> || std::is_same_v<S, integer>, int> = 0>
inline bool operator==(const polynomial<K, C> &x, const S &s)
{
    return x == polynomial<K, C>(s);
}

// ---- degrees and truncation as test oracles (polynomial.hpp:2364-2445; series.hpp degree) -------------------
template <typename K, typename C>
inline typename K::value_type degree(const polynomial<K, C> &p)
{
    using T = typename K::value_type;
    bool first = true;
    T best(0);
    for (const auto &[k, c] : p) {
        const T d = key_degree(k, p.get_symbol_set());
        if (first || d > best) {
            best = d;
            first = false;
        }
    }
    return best;
}
template <typename K, typename C>
inline typename K::value_type p_degree(const polynomial<K, C> &p, const symbol_set &s)
{
    using T = typename K::value_type;
    const auto idx = ss_intersect_idx(s, p.get_symbol_set());
    bool first = true;
    T best(0);
    for (const auto &[k, c] : p) {
        const T d = key_p_degree(k, idx, p.get_symbol_set());
        if (first || d > best) {
            best = d;
            first = false;
        }
    }
    return best;
}
template <typename K, typename C, typename V>
inline polynomial<K, C> truncate_degree(const polynomial<K, C> &p, const V &lim)
{
    polynomial<K, C> r;
    r.set_symbol_set(p.get_symbol_set());
    for (const auto &[k, c] : p) {
        if (!(lim < key_degree(k, p.get_symbol_set()))) {
            r._insert_unique(k, c);
        }
    }
    return r;
}
template <typename K, typename C, typename V>
inline polynomial<K, C> truncate_p_degree(const polynomial<K, C> &p, const V &lim, const symbol_set &s)
{
    const auto idx = ss_intersect_idx(s, p.get_symbol_set());
    polynomial<K, C> r;
    r.set_symbol_set(p.get_symbol_set());
    for (const auto &[k, c] : p) {
        if (!(lim < key_p_degree(k, idx, p.get_symbol_set()))) {
            r._insert_unique(k, c);
        }
    }
    return r;
}

// obake::pow by repeated multiplication (series pow, series.hpp:1628-1712: v.back() * b), every intermediate power
// resident in HBM (b200/device_polynomial.hpp, included at the end of this header)
namespace b200
{
template <typename K, typename C>
polynomial<K, C> pow_device_chain(const polynomial<K, C> &, unsigned);
}
template <typename K, typename C>
inline polynomial<K, C> pow(const polynomial<K, C> &b, unsigned n)
{
    return b200::pow_device_chain(b, n);
}

// ---- make_polynomials (polynomial.hpp:97-227) --------------------------------------------------------------
namespace detail
{
template <typename P>
inline P make_generator(const symbol_set &ss, const std::string &name)
{
    using K = typename P::key_type;
    using T = typename K::value_type;
    const auto pos = ss.index_of(name);
    if (pos == ss.size()) {
        throw std::invalid_argument("Cannot create a polynomial with symbol set " + std::string("{...}") + " from the generator '"
                                    + name + "': the generator is not in the symbol set");
    }
    std::vector<T> e(ss.size(), T(0));
    e[pos] = T(1);
    P p;
    p.set_symbol_set(ss);
    p.add_term(K(e.begin(), e.size()), typename P::cf_type(1));
    return p;
}
} // namespace detail

template <typename P, typename... Names>
inline std::array<P, sizeof...(Names)> make_polynomials(const symbol_set &ss, const Names &...names)
{
    return {{detail::make_generator<P>(ss, std::string(names))...}};
}
template <typename P, typename... Names, std::enable_if_t<(std::is_convertible_v<Names, std::string> && ...), int> = 0>
inline std::array<P, sizeof...(Names)> make_polynomials(const Names &...names)
{
    return {{detail::make_generator<P>(symbol_set{std::string(names)}, std::string(names))...}};
}

} // namespace obake

// device-resident chains: the definition of b200::pow_device_chain (this header's guard is already set)
#include "../b200/device_polynomial.hpp"

#endif
