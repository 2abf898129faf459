// obake/b200/device_polynomial.hpp -- a polynomial whose terms live in B200 HBM.
//
// The reference's benchmarks build their operands with chains of products and sums (f *= tmp in
// benchmark/dense.hpp:28-32 and sparse.hpp:33-36, 70 dependent products in rectangular_01.cpp:42-47,
// truncated powers in audi_01.cpp:32-42); with a host container every link of the chain pays a term copy
// (polynomial.hpp:828-833), an upload, a download and a hash-table rebuild.  device_polynomial<K, C> keeps the
// engine's SoA term arrays on the device between operations:
//     device_polynomial<K, C> f(host_poly);          upload once
//     f * g, truncated_mul(f, g, limit[, symbols])    obk_poly_mul             (polynomial.hpp:2026-2032, 2113-2127)
//     f + g, f - g                                    obk_poly_addsub          (series.hpp:2195-2991, identical symbol sets)
//     truncate_degree / truncate_p_degree             obk_poly_truncate_degree (polynomial.hpp:2364-2445)
//     f.to_host()                                     download + container rebuild, once
//     operands over different symbol sets           obk_poly_extend_symbols  (series.hpp:539-625: keys re-packed on the device)
// Errors are the host mirror's:
// std::overflow_error / std::invalid_argument / std::runtime_error (obake/b200/engine.hpp).
#ifndef OBAKE_B200_DEVICE_POLYNOMIAL_HPP
#define OBAKE_B200_DEVICE_POLYNOMIAL_HPP

#include <cstdint>
#include <mutex>
#include <stdexcept>
#include <utility>
#include <vector>

#include "../polynomials/polynomial.hpp"
#include "engine.hpp"

namespace obake::b200
{

template <typename K, typename C>
class device_polynomial
{
    using T = typename K::value_type;
    using CT = cf_traits<C>;
    symbol_set m_ss;
    obk_poly_result m_res{};

    void release()
    {
        if (m_res.owner) {
            ::obk_result_free(engine::instance().handle(), &m_res);
        }
        m_res = obk_poly_result{};
    }
    // Operands over different symbol sets are re-expressed over the union ON THE DEVICE (series_sym_extender,
    // series.hpp:539-625 -> obk_poly_extend_symbols) as poly_mul_impl does on the host (polynomial.hpp:1955-2010).
    static void unify(const device_polynomial &a, const device_polynomial &b, device_polynomial &ta, device_polynomial &tb,
                      const device_polynomial *&pa, const device_polynomial *&pb)
    {
        pa = &a;
        pb = &b;
        if (a.m_ss == b.m_ss) {
            return;
        }
        std::vector<std::size_t> posa, posb;
        const symbol_set u = merge_symbol_sets(a.m_ss, b.m_ss, posa, posb);
        if (!(u == a.m_ss)) {
            ta = a.extend(u, posa);
            pa = &ta;
        }
        if (!(u == b.m_ss)) {
            tb = b.extend(u, posb);
            pb = &tb;
        }
    }
    static device_polynomial adopt(const symbol_set &ss, const obk_poly_result &r)
    {
        device_polynomial p;
        p.m_ss = ss;
        p.m_res = r;
        return p;
    }

public:
    device_polynomial() = default;
    device_polynomial(const device_polynomial &) = delete;
    device_polynomial &operator=(const device_polynomial &) = delete;
    device_polynomial(device_polynomial &&o) noexcept : m_ss(std::move(o.m_ss)), m_res(o.m_res)
    {
        o.m_res = obk_poly_result{};
    }
    device_polynomial &operator=(device_polynomial &&o) noexcept
    {
        if (this != &o) {
            release();
            m_ss = std::move(o.m_ss);
            m_res = o.m_res;
            o.m_res = obk_poly_result{};
        }
        return *this;
    }
    ~device_polynomial()
    {
        release();
    }

    // upload: the SoA term copy of polynomial.hpp:828-833, done once for a whole chain
    explicit device_polynomial(const polynomial<K, C> &p) : m_ss(p.get_symbol_set())
    {
        const std::size_t nvars = m_ss.size();
        const unsigned W = K::n_words(nvars);
        unsigned lin = 1;
        for (const auto &[k, c] : p) lin = std::max(lin, CT::limbs_needed(c));
        if (lin > 2) {
            throw std::invalid_argument("input coefficients wider than two limbs are not supported by the engine");
        }
        std::vector<std::uint64_t> keys(p.size() * W), cfs(p.size() * lin);
        std::size_t i = 0;
        for (const auto &[k, c] : p) {
            k.to_words(&keys[i * W], nvars);
            CT::store(c, &cfs[i * lin], lin);
            ++i;
        }
        obk_poly_view v = layout();
        v.nterms = p.size();
        v.cf_limbs = lin;
        v.mem = OBK_MEM_HOST;
        v.keys = keys.data();
        v.cfs = cfs.data();
        auto &eng = engine::instance();
        std::lock_guard<std::mutex> g(eng.mutex());
        eng.check(::obk_poly_copy(eng.handle(), &v, OBK_MEM_DEVICE, &m_res));
    }

    const symbol_set &get_symbol_set() const
    {
        return m_ss;
    }
    std::size_t size() const
    {
        return static_cast<std::size_t>(m_res.nterms);
    }
    bool empty() const
    {
        return m_res.nterms == 0;
    }
    // key layout of this polynomial as the C ABI describes it (no terms attached)
    obk_poly_view layout() const
    {
        obk_poly_view v{};
        const std::size_t nvars = m_ss.size();
        v.key_words = K::n_words(nvars);
        v.key_signed = std::is_signed_v<T> ? 1u : 0u;
        v.nvars = static_cast<std::uint32_t>(nvars);
        v.psize = K::psize;
        v.cf_kind = CT::kind;
        v.cf_limbs = 1;
        v.key_bits = sizeof(T) * 8u;
        return v;
    }
    // the terms where they are: device pointers owned by the engine
    obk_poly_view view() const
    {
        obk_poly_view v = layout();
        v.nterms = m_res.nterms;
        v.cf_limbs = m_res.nterms ? m_res.cf_limbs : 1u;
        v.mem = OBK_MEM_DEVICE;
        v.keys = m_res.keys;
        v.cfs = m_res.cfs;
        return v;
    }
    // this polynomial over the larger symbol set u; pos[j] = position in u of the j-th symbol of this one
    device_polynomial extend(const symbol_set &u, const std::vector<std::size_t> &pos) const
    {
        device_polynomial r;
        r.m_ss = u;
        if (empty()) {
            return r;
        }
        std::vector<std::uint32_t> p32(pos.begin(), pos.end());
        const obk_poly_view v = view();
        auto &eng = engine::instance();
        std::lock_guard<std::mutex> g(eng.mutex());
        eng.check(::obk_poly_extend_symbols(eng.handle(), &v, static_cast<std::uint32_t>(u.size()), p32.data(), OBK_MEM_DEVICE,
                                            &r.m_res));
        return r;
    }
    // download + host container rebuild
    polynomial<K, C> to_host() const
    {
        polynomial<K, C> ret;
        ret.set_symbol_set(m_ss);
        if (empty()) {
            return ret;
        }
        auto &eng = engine::instance();
        obk_poly_result h{};
        const obk_poly_view v = view();
        {
            std::lock_guard<std::mutex> g(eng.mutex());
            eng.check(::obk_poly_copy(eng.handle(), &v, OBK_MEM_HOST, &h));
        }
        const std::size_t nvars = m_ss.size();
        const unsigned W = K::n_words(nvars);
        try {
            ret.reserve(h.nterms);
            const auto *rc = static_cast<const std::uint64_t *>(h.cfs);
            for (std::uint64_t i = 0; i < h.nterms; ++i) {
                ret._insert_unique(K::from_words(h.keys + i * W, nvars), CT::load(rc + i * h.cf_limbs, h.cf_limbs));
            }
        } catch (...) {
            ::obk_result_free(eng.handle(), &h);
            throw;
        }
        ::obk_result_free(eng.handle(), &h);
        return ret;
    }

    // ---- the product (poly_mul_impl_identical_ss seam, polynomial.hpp:1878-1929) -------------------------------
    static device_polynomial mul(const device_polynomial &a0, const device_polynomial &b0, const ::obake::detail::trunc_args &ta)
    {
        device_polynomial xa, xb;
        const device_polynomial *pa, *pb;
        unify(a0, b0, xa, xb, pa, pb);
        const device_polynomial &a = *pa, &b = *pb;
        if (a.empty() || b.empty() || ta.reject_all) {
            device_polynomial r;
            r.m_ss = a.m_ss;
            return r;
        }
        obk_trunc tr{};
        if (ta.on) {
            tr.limit = ta.limit;
            tr.partial = ta.partial ? 1u : 0u;
            tr.nidx = static_cast<std::uint32_t>(ta.idx.size());
            tr.idx = ta.idx.data();
        }
        const obk_poly_view va = a.view(), vb = b.view();
        obk_poly_result r{};
        auto &eng = engine::instance();
        std::lock_guard<std::mutex> g(eng.mutex());
        eng.check(::obk_poly_mul(eng.handle(), &va, &vb, ta.on ? &tr : nullptr, OBK_MEM_DEVICE, &r, nullptr));
        return adopt(a.m_ss, r);
    }
    friend device_polynomial operator*(const device_polynomial &a, const device_polynomial &b)
    {
        return mul(a, b, ::obake::detail::trunc_args{});
    }
    template <typename V>
    friend device_polynomial truncated_mul(const device_polynomial &a, const device_polynomial &b, const V &max_degree)
    {
        ::obake::detail::trunc_args ta;
        ::obake::detail::reduce_limit<T>(ta, max_degree);
        return mul(a, b, ta);
    }
    template <typename V>
    friend device_polynomial truncated_mul(const device_polynomial &a, const device_polynomial &b, const V &max_degree,
                                           const symbol_set &s)
    {
        ::obake::detail::trunc_args ta;
        ::obake::detail::reduce_limit<T>(ta, max_degree);
        ta.partial = true;
        std::vector<std::size_t> posa, posb;
        ta.idx = ss_intersect_idx(s, merge_symbol_sets(a.m_ss, b.m_ss, posa, posb)); // positions in the MERGED set
        return mul(a, b, ta);
    }

    // ---- addition / subtraction (series.hpp:2195-2991, identical symbol sets) ----------------------------------
    static device_polynomial addsub(const device_polynomial &a0, const device_polynomial &b0, bool subtract)
    {
        device_polynomial xa, xb;
        const device_polynomial *pa, *pb;
        unify(a0, b0, xa, xb, pa, pb);
        const device_polynomial &a = *pa, &b = *pb;
        const obk_poly_view va = a.view(), vb = b.view();
        obk_poly_result r{};
        auto &eng = engine::instance();
        std::lock_guard<std::mutex> g(eng.mutex());
        eng.check(::obk_poly_addsub(eng.handle(), &va, &vb, subtract ? 1 : 0, OBK_MEM_DEVICE, &r, nullptr));
        return adopt(a.m_ss, r);
    }
    friend device_polynomial operator+(const device_polynomial &a, const device_polynomial &b)
    {
        return addsub(a, b, false);
    }
    friend device_polynomial operator-(const device_polynomial &a, const device_polynomial &b)
    {
        return addsub(a, b, true);
    }

    // ---- truncate_degree / truncate_p_degree (polynomial.hpp:2364-2445) ----------------------------------------
    static device_polynomial truncate(const device_polynomial &a, const ::obake::detail::trunc_args &ta)
    {
        if (a.empty() || ta.reject_all) {
            device_polynomial r;
            r.m_ss = a.m_ss;
            return r;
        }
        obk_trunc tr{};
        tr.limit = ta.limit;
        tr.partial = ta.partial ? 1u : 0u;
        tr.nidx = static_cast<std::uint32_t>(ta.idx.size());
        tr.idx = ta.idx.data();
        const obk_poly_view va = a.view();
        obk_poly_result r{};
        auto &eng = engine::instance();
        std::lock_guard<std::mutex> g(eng.mutex());
        eng.check(::obk_poly_truncate_degree(eng.handle(), &va, &tr, OBK_MEM_DEVICE, &r, nullptr));
        return adopt(a.m_ss, r);
    }
    template <typename V>
    friend device_polynomial truncate_degree(const device_polynomial &a, const V &lim)
    {
        ::obake::detail::trunc_args ta;
        ::obake::detail::reduce_limit<T>(ta, lim);
        return truncate(a, ta);
    }
    template <typename V>
    friend device_polynomial truncate_p_degree(const device_polynomial &a, const V &lim, const symbol_set &s)
    {
        ::obake::detail::trunc_args ta;
        ::obake::detail::reduce_limit<T>(ta, lim);
        ta.partial = true;
        ta.idx = ss_intersect_idx(s, a.m_ss);
        return truncate(a, ta);
    }
};

// base^n with every intermediate power resident in HBM: one upload, n - 1 products, one download.  The reference's
// series pow multiplies the previous power by the base (series.hpp:1628-1712, `v.back() * b`); its natural-power
// cache (one host series per power) is replaced by keeping the running power on the device.
template <typename K, typename C>
inline device_polynomial<K, C> pow(const device_polynomial<K, C> &base, unsigned n)
{
    if (n == 0u) {
        throw std::invalid_argument("pow(device_polynomial, 0): the unit polynomial has no device form -- use the host pow");
    }
    device_polynomial<K, C> acc = base * base; // n == 1 is handled by the callers that can copy
    for (unsigned i = 2; i < n; ++i) {
        acc = acc * base;
    }
    return acc;
}
// The truncated power loop of benchmark/audi_01.cpp:32-42 (f = truncated_mul(f, base, limit), n - 1 times).
template <typename K, typename C, typename V>
inline device_polynomial<K, C> truncated_pow(const device_polynomial<K, C> &base, unsigned n, const V &max_degree)
{
    if (n < 2u) {
        throw std::invalid_argument("truncated_pow(device_polynomial, n): n must be at least 2");
    }
    device_polynomial<K, C> acc = truncated_mul(base, base, max_degree);
    for (unsigned i = 2; i < n; ++i) {
        acc = truncated_mul(acc, base, max_degree);
    }
    return acc;
}

// obake::pow of a host polynomial (declared in polynomials/polynomial.hpp) as a device-resident chain.
template <typename K, typename C>
inline polynomial<K, C> pow_device_chain(const polynomial<K, C> &b, unsigned n)
{
    if (n == 0u) {
        return polynomial<K, C>(1);
    }
    if (n == 1u) {
        return polynomial<K, C>(1) * b; // what the reference's cache holds as v[1] = v[0] * b
    }
    if (b.empty()) {
        polynomial<K, C> r;
        r.set_symbol_set(b.get_symbol_set());
        return r;
    }
    if (engine::instance().n_devices() > 1u) {
        // the single-process multi-device engine takes host operands: the chain goes through the host container
        polynomial<K, C> r = b * b;
        for (unsigned i = 2; i < n; ++i) {
            r = r * b;
        }
        return r;
    }
    return pow(device_polynomial<K, C>(b), n).to_host();
}

} // namespace obake::b200

#endif
