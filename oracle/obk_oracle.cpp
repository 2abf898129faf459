// TEST INFRASTRUCTURE ONLY -- the product path (obake_b200/) never links, loads or calls this file.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
//
// CPU restatement of obake's multithreaded homomorphic polynomial product for the series-product
// hot path, written against the reference sources under /root/reference (file:line cited per
// function).  The reference itself cannot be built in this image: it needs mp++ >= 0.27,
// Boost >= 1.81, oneTBB, {fmt} and GMP headers, none of which are installed, so there
// is no oracle/_ref; the third-party arithmetic on the path is mp++'s integer<N> (exact ring
// operations in Z), restated here with fixed-width two's-complement limbs that are wide enough never
// to wrap (the caller picks LACC from an a-priori bound; the widest is 5 limbs = 320 bits).
//
// Parity: PINNED.  tests/test_oracle.py checks this file against (i) the reference's Kronecker tables
// (tests/golden/kpack_tables.json, parsed from src/kpack.cpp), (ii) the reference tests' known
// answers (tests/golden/known_answers.json: 2 096 600 terms, 53 130 terms, the monomial and
// truncation identities), (iii) the closed form of the Fateman product, and (iv) an independent
// big-integer Python restatement (oracle/pyoracle.py).
//
// Build: make -C oracle   (g++ -O3 -std=c++17 -pthread -shared -fPIC)
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <unordered_set>
#include <vector>

typedef unsigned __int128 u128;
typedef __int128 i128;

// ------------------------------------------------------------------------------------------------
// Kronecker packing tables -- recipe of tools/kpack.ipynb, results equal src/kpack.cpp:289-338,786-836.
// ------------------------------------------------------------------------------------------------
namespace
{

struct kp_tables {
    int max_size = 0;
    uint64_t deltas[24], lims[24], klims[24];
    // divcnst[size-1][j] = (m', sh1, sh2) for the divisor delta^j  (kpack.hpp:347-356)
    uint64_t mp[24][25];
    unsigned sh1[24][25], sh2[24][25];
};

uint64_t mulmod(uint64_t a, uint64_t b, uint64_t m) { return (uint64_t)((u128)a * b % m); }
uint64_t powmod(uint64_t a, uint64_t e, uint64_t m)
{
    uint64_t r = 1;
    a %= m;
    while (e) {
        if (e & 1) r = mulmod(r, a, m);
        a = mulmod(a, a, m);
        e >>= 1;
    }
    return r;
}
bool is_prime(uint64_t n)
{
    if (n < 2) return false;
    static const uint64_t sp[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    for (uint64_t p : sp) {
        if (n % p == 0) return n == p;
    }
    uint64_t d = n - 1;
    int r = 0;
    while ((d & 1) == 0) {
        d >>= 1;
        ++r;
    }
    for (uint64_t a : sp) {
        uint64_t x = powmod(a, d, n);
        if (x == 1 || x == n - 1) continue;
        bool comp = true;
        for (int i = 1; i < r; ++i) {
            x = mulmod(x, x, n);
            if (x == n - 1) {
                comp = false;
                break;
            }
        }
        if (comp) return false;
    }
    return true;
}
// floor((2^bw)^(1/s))
uint64_t iroot_pow2(int bw, int s)
{
    if (s == 1) return 0; // unused
    uint64_t lo = 0, hi = (uint64_t)1 << (bw / s + 1);
    while (lo < hi) {
        uint64_t mid = lo + (hi - lo + 1) / 2;
        // mid^s <= 2^bw ?
        u128 p = 1;
        bool over = false;
        for (int i = 0; i < s; ++i) {
            p *= mid;
            if (p > ((u128)1 << bw)) {
                over = true;
                break;
            }
        }
        if (!over) lo = mid; else hi = mid - 1;
    }
    return lo;
}

kp_tables make_tables(bool is_signed)
{
    kp_tables t;
    const int bits = 64;
    const int bw = is_signed ? 63 : 64;
    int s = 1;
    while (bw / s >= 3) {
        uint64_t d;
        if (s == 1) {
            d = is_signed ? (uint64_t)INT64_MAX : UINT64_MAX;
        } else {
            d = iroot_pow2(bw, s) - 1;
            while (!is_prime(d)) --d;
        }
        uint64_t lim = is_signed ? (d - 1) / 2 : d - 1;
        u128 kl = 0, cur = 1;
        for (int j = 0; j < s; ++j) {
            kl += (u128)lim * cur;
            cur *= d;
        }
        t.deltas[s - 1] = d;
        t.lims[s - 1] = lim;
        t.klims[s - 1] = (uint64_t)kl;
        ++s;
    }
    t.max_size = s - 1;
    for (int sz = 1; sz <= t.max_size; ++sz) {
        u128 dd = 1;
        for (int j = 0; j <= sz; ++j) {
            // l = ceil(log2 dd)
            int l = 0;
            while (((u128)1 << l) < dd) ++l;
            // m' = floor(2^(N+l)/dd) - 2^N + 1, computed as floor(2^N*(2^l - dd)/dd) + 1
            u128 num = (((u128)1 << l) - dd) << bits; // (2^l-dd) < dd <= 2^64 so this fits in 128 bits
            uint64_t mp = (uint64_t)(num / dd) + 1;
            t.mp[sz - 1][j] = mp;
            t.sh1[sz - 1][j] = l < 1 ? l : 1;
            t.sh2[sz - 1][j] = l > 1 ? l - 1 : 0;
            if (j < sz) dd *= t.deltas[sz - 1];
        }
    }
    return t;
}

const kp_tables &tabs(bool is_signed)
{
    static const kp_tables tu = make_tables(false), ts = make_tables(true);
    return is_signed ? ts : tu;
}

inline uint64_t mulhi64(uint64_t a, uint64_t b) { return (uint64_t)(((u128)a * b) >> 64); }

// kunpacker::operator>> (kpack.hpp:321-375): remainder and quotient via division by constants.
// Unpacks `size` components of `value`; results are returned as int64 (unsigned exponents < 2^63
// in every multi-component size; size==1 values are returned bit-for-bit).
void kunpack(bool is_signed, uint64_t value, unsigned size, int64_t *out)
{
    if (size == 0) return;
    const kp_tables &t = tabs(is_signed);
    const uint64_t delta = t.deltas[size - 1];
    const uint64_t klim_min = is_signed ? (uint64_t)(-(int64_t)t.klims[size - 1]) : 0;
    const int64_t lim_min = is_signed ? -(int64_t)t.lims[size - 1] : 0;
    const uint64_t n = value - klim_min;
    uint64_t cur_prod = 1;
    auto divc = [](uint64_t nn, uint64_t mp, unsigned s1, unsigned s2) {
        const uint64_t t1 = mulhi64(mp, nn);
        return (t1 + ((nn - t1) >> s1)) >> s2;
    };
    for (unsigned idx = 0; idx < size; ++idx) {
        const uint64_t prev = cur_prod;
        (void)prev;
        cur_prod *= delta;
        const uint64_t q_r = divc(n, t.mp[size - 1][idx + 1], t.sh1[size - 1][idx + 1], t.sh2[size - 1][idx + 1]);
        const uint64_t rem = n - q_r * cur_prod;
        const uint64_t q_d = divc(rem, t.mp[size - 1][idx], t.sh1[size - 1][idx], t.sh2[size - 1][idx]);
        out[idx] = (int64_t)q_d + lim_min;
    }
}

// Layout of a key: W words; packed_monomial => W==1, word packs nvars exponents with the size-nvars
// delta (packed_monomial.hpp:56-142); d_packed_monomial<T,psize> => word w packs exponents
// [w*psize, (w+1)*psize) with the size-psize delta, last word zero padded (d_packed_monomial.hpp:110-132).
struct layout {
    bool is_signed;
    unsigned nvars, psize, W;
    unsigned word_size() const { return psize ? psize : nvars; }
};

void unpack_key(const layout &L, const uint64_t *key, int64_t *expos)
{
    if (L.nvars == 0) return;
    if (L.psize == 0) {
        kunpack(L.is_signed, key[0], L.nvars, expos);
        return;
    }
    int64_t tmp[24];
    unsigned idx = 0;
    for (unsigned w = 0; w < L.W && idx < L.nvars; ++w) {
        kunpack(L.is_signed, key[w], L.psize, tmp);
        for (unsigned j = 0; j < L.psize && idx < L.nvars; ++j, ++idx) expos[idx] = tmp[j];
    }
}

// hash(): packed_monomial.hpp:172-176 (the value itself); d_packed_monomial.hpp:274-284 (sum of words).
inline uint64_t key_hash(const uint64_t *k, unsigned W)
{
    uint64_t h = 0;
    for (unsigned w = 0; w < W; ++w) h += k[w];
    return h;
}

// xoroshiro128+ (include/obake/detail/xoroshiro128_plus.hpp:33-45) -- estimator RNG.
struct xoroshiro128p {
    uint64_t s0, s1;
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next()
    {
        const uint64_t a = s0;
        uint64_t b = s1;
        const uint64_t r = a + b;
        b ^= a;
        s0 = rotl(a, 24) ^ b ^ (b << 16);
        s1 = rotl(b, 37);
        return r;
    }
    // unbiased-enough bounded draw for an estimator
    uint64_t below(uint64_t n) { return n ? (uint64_t)(((u128)next() * n) >> 64) : 0; }
};

} // namespace

// ------------------------------------------------------------------------------------------------
// C interface
// ------------------------------------------------------------------------------------------------
extern "C" {

typedef struct {
    uint64_t nterms;
    uint32_t key_words; // W
    uint32_t key_signed;
    uint32_t nvars;
    uint32_t psize;    // 0: packed_monomial
    uint32_t cf_kind;  // 0: integer limbs, 1: f64
    uint32_t cf_limbs; // L_in (1, 2 or 3) for integer
    const uint64_t *keys;
    const void *cfs;
} orc_poly;

typedef struct {
    int64_t limit;         // max degree (compared in the key's own integer type)
    uint32_t nidx;         // 0 => total degree
    const uint32_t *idx;   // sorted variable positions for the partial degree
} orc_trunc;

typedef struct {
    uint64_t nterms;
    uint32_t key_words;
    uint32_t cf_kind;
    uint32_t cf_limbs; // L_acc limbs per integer coefficient
    uint32_t log2_nsegs;
    uint64_t *keys;
    void *cfs;
    uint64_t *seg_offsets; // 2^log2_nsegs + 1
    uint64_t n_mults;      // tot_n_mults
    uint64_t est_nterms;
    double seconds;        // wall-clock of the product proper (as simple_timer would see it)
} orc_result;

// status codes
enum { ORC_OK = 0, ORC_KEY_OVERFLOW = 2, ORC_BAD_ARG = 1, ORC_DEGREE_OVERFLOW = 3 };

int orc_kpack_tables(int is_signed, uint64_t *deltas, uint64_t *lims, uint64_t *klims, uint64_t *mp, uint32_t *sh1,
                     uint32_t *sh2)
{
    const kp_tables &t = tabs(is_signed != 0);
    for (int s = 0; s < t.max_size; ++s) {
        deltas[s] = t.deltas[s];
        lims[s] = t.lims[s];
        klims[s] = t.klims[s];
        for (int j = 0; j <= t.max_size; ++j) {
            const bool live = j <= s + 1;
            mp[s * (t.max_size + 1) + j] = live ? t.mp[s][j] : 0;
            sh1[s * (t.max_size + 1) + j] = live ? t.sh1[s][j] : 0;
            sh2[s * (t.max_size + 1) + j] = live ? t.sh2[s][j] : 0;
        }
    }
    return t.max_size;
}

// kpacker (kpack.hpp:219-280).  Returns 0, or 1 = size too large, 2 = component out of range.
int orc_kpack(int is_signed, const int64_t *expos, uint32_t n, uint32_t size, uint64_t *out)
{
    const kp_tables &t = tabs(is_signed != 0);
    if (size == 0) {
        *out = 0;
        return 0;
    }
    if ((int)size > t.max_size || n > size) return 1;
    const uint64_t delta = t.deltas[size - 1], lim = t.lims[size - 1];
    uint64_t v = 0, cur = 1;
    for (uint32_t i = 0; i < n; ++i) {
        if (is_signed) {
            if (expos[i] < -(int64_t)lim || expos[i] > (int64_t)lim) return 2;
        } else {
            if (expos[i] < 0 || (uint64_t)expos[i] > lim) return 2;
        }
        v += (uint64_t)expos[i] * cur;
        cur *= delta;
    }
    *out = v;
    return 0;
}

// kunpacker (kpack.hpp:283-375).  Returns 0, or 2 when the coded value is outside [klim_min, klim_max].
int orc_kunpack(int is_signed, uint64_t value, uint32_t size, int64_t *out)
{
    const kp_tables &t = tabs(is_signed != 0);
    if (size == 0) return value == 0 ? 0 : 1;
    if ((int)size > t.max_size) return 1;
    if (is_signed) {
        const int64_t k = (int64_t)t.klims[size - 1];
        if ((int64_t)value < -k || (int64_t)value > k) return 2;
    } else if (value > t.klims[size - 1]) {
        return 2;
    }
    kunpack(is_signed != 0, value, size, out);
    return 0;
}

// key_degree / key_p_degree (src/polynomials/packed_monomial.cpp:334-350,388-412;
// d_packed_monomial.hpp:901-958).  idx == NULL => total degree.
void orc_degrees(const orc_poly *p, const uint32_t *idx, uint32_t nidx, int64_t *out)
{
    layout L{p->key_signed != 0, p->nvars, p->psize, p->key_words};
    std::vector<int64_t> e(L.nvars + 1);
    for (uint64_t i = 0; i < p->nterms; ++i) {
        unpack_key(L, p->keys + i * L.W, e.data());
        int64_t d = 0;
        if (idx) {
            for (uint32_t j = 0; j < nidx; ++j) d += e[idx[j]];
        } else {
            for (unsigned j = 0; j < L.nvars; ++j) d += e[j];
        }
        out[i] = d;
    }
}

void orc_unpack_keys(const orc_poly *p, int64_t *out /* nterms x nvars */)
{
    layout L{p->key_signed != 0, p->nvars, p->psize, p->key_words};
    for (uint64_t i = 0; i < p->nterms; ++i) unpack_key(L, p->keys + i * L.W, out + i * L.nvars);
}

// monomial_range_overflow_check: packed_monomial.hpp:289-488, d_packed_monomial.hpp:583-897.
// Returns 1 when the product is safe, 0 on overflow.
int orc_overflow_check(const orc_poly *x, const orc_poly *y)
{
    layout L{x->key_signed != 0, x->nvars, x->psize, x->key_words};
    if (L.nvars == 0) return 1;
    if (x->nterms == 0 || y->nterms == 0) return 1;
    const unsigned nv = L.nvars;
    std::vector<int64_t> mn1(nv), mx1(nv), mn2(nv), mx2(nv), e(nv);
    i128 dmin1 = 0, dmax1 = 0, dmin2 = 0, dmax2 = 0;
    auto scan = [&](const orc_poly *p, std::vector<int64_t> &mn, std::vector<int64_t> &mx, i128 &dmin, i128 &dmax) {
        for (uint64_t i = 0; i < p->nterms; ++i) {
            unpack_key(L, p->keys + i * L.W, e.data());
            i128 d = 0;
            for (unsigned j = 0; j < nv; ++j) {
                // unsigned single-component keys can exceed int64: compare as the proper type
                if (i == 0) {
                    mn[j] = mx[j] = e[j];
                } else if (L.is_signed) {
                    mn[j] = std::min(mn[j], e[j]);
                    mx[j] = std::max(mx[j], e[j]);
                } else {
                    if ((uint64_t)e[j] < (uint64_t)mn[j]) mn[j] = e[j];
                    if ((uint64_t)e[j] > (uint64_t)mx[j]) mx[j] = e[j];
                }
                d += L.is_signed ? (i128)e[j] : (i128)(uint64_t)e[j];
            }
            if (i == 0) {
                dmin = dmax = d;
            } else {
                dmin = std::min(dmin, d);
                dmax = std::max(dmax, d);
            }
        }
    };
    scan(x, mn1, mx1, dmin1, dmax1);
    scan(y, mn2, mx2, dmin2, dmax2);
    const kp_tables &t = tabs(L.is_signed);
    const uint64_t lim = t.lims[L.word_size() - 1];
    for (unsigned j = 0; j < nv; ++j) {
        if (L.is_signed) {
            const i128 amin = (i128)mn1[j] + mn2[j], amax = (i128)mx1[j] + mx2[j];
            if (amin < -(i128)lim || amax > (i128)lim) return 0;
        } else {
            const u128 amax = (u128)(uint64_t)mx1[j] + (uint64_t)mx2[j];
            if (amax > lim) return 0;
        }
    }
    if (L.psize != 0) {
        // d_packed only: the degree of the product must stay in the range of T (d_packed_monomial.hpp:879-894)
        if (L.is_signed) {
            if (dmin1 + dmin2 < (i128)INT64_MIN || dmax1 + dmax2 > (i128)INT64_MAX) return 0;
        } else {
            if (dmax1 + dmax2 > (i128)UINT64_MAX) return 0;
        }
    }
    return 1;
}

} // extern "C"

// ------------------------------------------------------------------------------------------------
// The multiply: poly_mul_impl_mt_hm, include/obake/polynomials/polynomial.hpp:797-1582
// ------------------------------------------------------------------------------------------------
namespace
{

template <int W>
struct keyw {
    uint64_t w[W];
};

// Exact accumulator: LACC two's-complement limbs.
template <int LACC>
struct iacc {
    uint64_t l[LACC];
    bool is_zero() const
    {
        for (int i = 0; i < LACC; ++i)
            if (l[i]) return false;
        return true;
    }
};

// cf = c1*c2 (first hit) / fma3(cf, c1, c2) (later hits): polynomial.hpp:1368-1384, fma3.hpp:84-88.
template <int LIN, int LACC>
inline void mul_to_limbs(const uint64_t *a, const uint64_t *b, uint64_t *p /* LACC limbs */)
{
    if (LIN == 1) {
        const i128 pr = (i128)(int64_t)a[0] * (int64_t)b[0];
        p[0] = (uint64_t)pr;
        if (LACC > 1) p[1] = (uint64_t)((u128)pr >> 64);
        const uint64_t ext = pr < 0 ? ~(uint64_t)0 : 0;
        for (int i = 2; i < LACC; ++i) p[i] = ext;
    } else {
        // LIN x LIN limbs -> 2 LIN limbs signed, via magnitudes (mp++ integers are sign + magnitude as well)
        uint64_t x[LIN], y[LIN];
        const bool na = (int64_t)a[LIN - 1] < 0, nb = (int64_t)b[LIN - 1] < 0;
        auto magnitude = [](const uint64_t *v, bool negv, uint64_t *m) {
            uint64_t c = 1;
            for (int i = 0; i < LIN; ++i) {
                if (negv) {
                    m[i] = ~v[i] + c;
                    c = (c && m[i] == 0) ? 1 : 0;
                } else {
                    m[i] = v[i];
                }
            }
        };
        magnitude(a, na, x);
        magnitude(b, nb, y);
        const bool neg = na != nb;
        uint64_t r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < LIN; ++i) {
            uint64_t carry = 0;
            for (int j = 0; j < LIN; ++j) {
                u128 cur = (u128)x[i] * y[j] + r[i + j] + carry;
                r[i + j] = (uint64_t)cur;
                carry = (uint64_t)(cur >> 64);
            }
            r[i + LIN] += carry;
        }
        if (neg) {
            // two's complement negate over 8 limbs
            uint64_t c = 1;
            for (int i = 0; i < 8; ++i) {
                r[i] = ~r[i] + c;
                c = (c && r[i] == 0) ? 1 : 0;
            }
        }
        for (int i = 0; i < LACC; ++i) p[i] = r[i];
    }
}

template <int LACC>
inline void add_limbs(uint64_t *acc, const uint64_t *p)
{
    unsigned char c = 0;
    for (int i = 0; i < LACC; ++i) {
        u128 s = (u128)acc[i] + p[i] + c;
        acc[i] = (uint64_t)s;
        c = (unsigned char)(s >> 64);
    }
}

inline uint64_t mix64(uint64_t x)
{
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

// One output segment's table (stands in for boost::unordered_flat_map<K,C>, series.hpp:637-639).
template <int W, typename ACC>
struct seg_table {
    std::vector<keyw<W>> keys;
    std::vector<ACC> acc;
    std::vector<unsigned char> used;
    size_t cap = 0, count = 0;
    unsigned shift = 0;
    void init(size_t c, unsigned sh)
    {
        cap = c;
        shift = sh;
        keys.assign(cap, keyw<W>{});
        acc.assign(cap, ACC{});
        used.assign(cap, 0);
        count = 0;
    }
    size_t slot_of(const keyw<W> &k) const
    {
        uint64_t h = 0;
        for (int i = 0; i < W; ++i) h = h * 0x9E3779B97F4A7C15ULL + k.w[i];
        return (size_t)(mix64(h >> shift) & (cap - 1));
    }
    void grow()
    {
        seg_table n;
        n.init(cap * 2, shift);
        for (size_t i = 0; i < cap; ++i) {
            if (!used[i]) continue;
            size_t s = n.slot_of(keys[i]);
            while (n.used[s]) s = (s + 1) & (n.cap - 1);
            n.used[s] = 1;
            n.keys[s] = keys[i];
            n.acc[s] = acc[i];
        }
        n.count = count;
        *this = std::move(n);
    }
    // try_emplace: returns (slot, inserted)
    std::pair<size_t, bool> find_or_insert(const keyw<W> &k)
    {
        if ((count + 1) * 2 > cap) grow();
        size_t s = slot_of(k);
        for (;;) {
            if (!used[s]) {
                used[s] = 1;
                keys[s] = k;
                ++count;
                return {s, true};
            }
            bool eq = true;
            for (int i = 0; i < W; ++i) eq &= keys[s].w[i] == k.w[i];
            if (eq) return {s, false};
            s = (s + 1) & (cap - 1);
        }
    }
};

struct term_ref {
    uint32_t idx;
    uint32_t bucket;
};

// Degree comparison `max_deg < d_i + d_j` in the key's own integer type T (polynomial.hpp:604-610).
struct trunc_ctx {
    bool on = false, is_signed = false;
    int64_t limit = 0;
    inline bool rejects(int64_t di, int64_t dj) const
    {
        if (is_signed) return limit < di + dj;
        // unsigned T: a negative int limit converts to a huge unsigned value in C++ only when
        // compared against unsigned; the reference compares `const int& < T` which performs the usual
        // arithmetic conversions, i.e. a negative int becomes 2^64-|limit| and never rejects... the
        // host wrapper resolves that case; here limit is already expressed in T.
        return (uint64_t)limit < (uint64_t)di + (uint64_t)dj;
    }
};

template <int W, int LIN, int LACC, bool F64>
int mul_impl(const orc_poly *x, const orc_poly *y, const orc_trunc *tr, int nthreads, orc_result *out)
{
    using ACC = typename std::conditional<F64, double, iacc<LACC>>::type;
    const layout L{x->key_signed != 0, x->nvars, x->psize, (unsigned)W};
    const size_t n1 = x->nterms, n2 = y->nterms;
    const uint64_t *k1 = x->keys, *k2 = y->keys;

    // Degree data (series.hpp:3757-3785 / 3911-3945) for the truncated variants.
    trunc_ctx tc;
    std::vector<int64_t> d1, d2;
    if (tr) {
        tc.on = true;
        tc.is_signed = L.is_signed;
        tc.limit = tr->limit;
        d1.resize(n1);
        d2.resize(n2);
        orc_degrees(x, tr->nidx ? tr->idx : nullptr, tr->nidx, d1.data());
        orc_degrees(y, tr->nidx ? tr->idx : nullptr, tr->nidx, d2.data());
    }
    auto deg_less = [&](int64_t a, int64_t b) { return L.is_signed ? a < b : (uint64_t)a < (uint64_t)b; };

    // tot_n_mults (polynomial.hpp:574-623).
    uint64_t tot = 0;
    std::vector<int64_t> d2s;
    if (tc.on) {
        d2s = d2;
        std::sort(d2s.begin(), d2s.end(), deg_less);
        for (size_t i = 0; i < n1; ++i) {
            const int64_t di = d1[i];
            tot += std::upper_bound(d2s.begin(), d2s.end(), 0, [&](int64_t, int64_t dj) { return tc.rejects(di, dj); })
                   - d2s.begin();
        }
    } else {
        tot = (uint64_t)n1 * n2;
    }
    out->n_mults = tot;
    if (tc.on && tot == 0) return ORC_OK; // polynomial.hpp:860-862

    // Product-size estimate (polynomial.hpp:625-794): random trials, stop at the first repeated
    // key, multiplier 3/2 * count^2.  The reference's value depends on TBB chunking (Appendix A.8);
    // only its order of magnitude matters (it sizes the segment count).
    uint64_t est = 1;
    {
        unsigned ntrials = (unsigned)std::max(5.0, 5e-8 * (double)tot);
        if (ntrials > 64) ntrials = 64;
        std::vector<uint32_t> vidx1(n1);
        std::vector<uint32_t> ysorted;
        if (tc.on) {
            ysorted.resize(n2);
            for (size_t i = 0; i < n2; ++i) ysorted[i] = (uint32_t)i;
            std::sort(ysorted.begin(), ysorted.end(), [&](uint32_t a, uint32_t b) { return deg_less(d2[a], d2[b]); });
        }
        long double acc = 0;
        for (unsigned t = 0; t < ntrials; ++t) {
            xoroshiro128p rng{(uint64_t)t + 18379758338774109289ULL, (uint64_t)t + 15967298767098049689ULL};
            for (size_t i = 0; i < n1; ++i) vidx1[i] = (uint32_t)i;
            for (size_t i = n1; i > 1; --i) std::swap(vidx1[i - 1], vidx1[rng.below(i)]);
            std::unordered_set<uint64_t> seen;
            uint64_t count = 0, ylim_acc = 0;
            bool collided = false;
            for (size_t ii = 0; ii < n1; ++ii) {
                const uint32_t i = vidx1[ii];
                size_t lim = n2;
                if (tc.on) {
                    const int64_t di = d1[i];
                    lim = std::upper_bound(d2s.begin(), d2s.end(), 0,
                                           [&](int64_t, int64_t dj) { return tc.rejects(di, dj); })
                          - d2s.begin();
                    if (lim == 0) continue;
                }
                ylim_acc += lim;
                uint32_t j = (uint32_t)rng.below(lim);
                if (tc.on) j = ysorted[j];
                uint64_t h = 0;
                for (int w = 0; w < W; ++w) h = h * 0x9E3779B97F4A7C15ULL + (k1[(size_t)i * W + w] + k2[(size_t)j * W + w]);
                if (!seen.insert(h).second) {
                    collided = true;
                    break;
                }
                ++count;
            }
            acc += collided ? (long double)3 * count * count / 2 : (long double)ylim_acc;
        }
        est = (uint64_t)std::max((long double)1, acc / ntrials);
    }
    out->est_nterms = est;

    // Segment count (polynomial.hpp:867-908).
    const size_t avg_term_size = (size_t)W * 8 + (F64 ? 8 : 8 * (size_t)LACC);
    const double est_sp = (double)est / (double)tot;
    const uint64_t seg_size = (!std::isfinite(est_sp) || est_sp >= 1e-3) ? 200 : 20;
    const uint64_t est_nsegs = est * avg_term_size / (seg_size * 1024);
    unsigned log2_nsegs = 0;
    while ((est_nsegs >> log2_nsegs) != 0) ++log2_nsegs; // nbits()
    if (log2_nsegs > 24) log2_nsegs = 24;
    const size_t nsegs = (size_t)1 << log2_nsegs;
    const uint64_t bmask = nsegs - 1;

    // Sort by bucket = hash % nsegs and segment (polynomial.hpp:914-992); truncated: sort each bucket
    // by degree (seg_sorter, :1005-1105).
    auto prepare = [&](const uint64_t *keys, size_t n, const std::vector<int64_t> &deg, std::vector<uint32_t> &order,
                       std::vector<uint32_t> &off) {
        order.resize(n);
        off.assign(nsegs + 1, 0);
        std::vector<uint32_t> b(n);
        for (size_t i = 0; i < n; ++i) {
            b[i] = (uint32_t)(key_hash(keys + i * W, W) & bmask);
            ++off[b[i] + 1];
        }
        for (size_t s = 0; s < nsegs; ++s) off[s + 1] += off[s];
        std::vector<uint32_t> cur(off.begin(), off.end() - 1);
        for (size_t i = 0; i < n; ++i) order[cur[b[i]]++] = (uint32_t)i;
        if (tc.on) {
            for (size_t s = 0; s < nsegs; ++s)
                std::stable_sort(order.begin() + off[s], order.begin() + off[s + 1],
                                 [&](uint32_t a, uint32_t c) { return deg_less(deg[a], deg[c]); });
        }
    };
    std::vector<uint32_t> o1, o2, off1, off2;
    prepare(k1, n1, d1, o1, off1);
    prepare(k2, n2, d2, o2, off2);

    // Gather sorted SoA copies (the reference copies the terms into vectors, :828-833).
    std::vector<keyw<W>> sk1(n1), sk2(n2);
    std::vector<int64_t> sd1(tc.on ? n1 : 0), sd2(tc.on ? n2 : 0);
    const size_t cw = F64 ? 1 : LIN;
    std::vector<uint64_t> sc1(n1 * cw), sc2(n2 * cw);
    for (size_t i = 0; i < n1; ++i) {
        for (int w = 0; w < W; ++w) sk1[i].w[w] = k1[(size_t)o1[i] * W + w];
        std::memcpy(&sc1[i * cw], (const uint64_t *)x->cfs + (size_t)o1[i] * cw, cw * 8);
        if (tc.on) sd1[i] = d1[o1[i]];
    }
    for (size_t i = 0; i < n2; ++i) {
        for (int w = 0; w < W; ++w) sk2[i].w[w] = k2[(size_t)o2[i] * W + w];
        std::memcpy(&sc2[i * cw], (const uint64_t *)y->cfs + (size_t)o2[i] * cw, cw * 8);
        if (tc.on) sd2[i] = d2[o2[i]];
    }

    // One owner per output segment (polynomial.hpp:1421-1553): all bucket pairs (i, (seg-i) mod nsegs).
    std::vector<std::vector<keyw<W>>> okeys(nsegs);
    std::vector<std::vector<ACC>> ocfs(nsegs);
    std::atomic<size_t> next{0};
    const size_t init_cap = 1024;
    auto worker = [&]() {
        seg_table<W, ACC> table;
        for (;;) {
            const size_t seg = next.fetch_add(1);
            if (seg >= nsegs) break;
            table.init(init_cap, log2_nsegs);
            for (size_t i = 0; i < nsegs; ++i) {
                const size_t j = seg >= i ? seg - i : nsegs - i + seg;
                const uint32_t r1s = off1[i], r1e = off1[i + 1], r2s = off2[j], r2e = off2[j + 1];
                if (r1s == r1e || r2s == r2e) continue;
                for (uint32_t a = r1s; a != r1e; ++a) {
                    uint32_t end2 = r2e;
                    if (tc.on) {
                        // compute_end_idx2 (polynomial.hpp:1187-1238)
                        const int64_t di = sd1[a];
                        end2 = (uint32_t)(std::upper_bound(sd2.begin() + r2s, sd2.begin() + r2e, 0,
                                                           [&](int64_t, int64_t dj) { return tc.rejects(di, dj); })
                                          - sd2.begin());
                        if (end2 == r2s) break; // both ranges are degree-sorted (:1480-1482)
                    }
                    const keyw<W> ka = sk1[a];
                    for (uint32_t b = r2s; b != end2; ++b) {
                        keyw<W> k;
                        for (int w = 0; w < W; ++w) k.w[w] = ka.w[w] + sk2[b].w[w]; // monomial_mul
                        auto res = table.find_or_insert(k);
                        if (F64) {
                            double c1, c2;
                            std::memcpy(&c1, &sc1[a], 8);
                            std::memcpy(&c2, &sc2[b], 8);
                            double *dst = (double *)&table.acc[res.first];
                            // no FP_FAST_FMA on default x86-64 flags: cf += c1*c2 with two roundings
                            volatile double prod = c1 * c2;
                            if (res.second) *dst = prod; else *dst += prod;
                        } else {
                            uint64_t p[LACC > 0 ? LACC : 1];
                            mul_to_limbs<LIN, LACC>(&sc1[(size_t)a * cw], &sc2[(size_t)b * cw], p);
                            uint64_t *dst = (uint64_t *)&table.acc[res.first];
                            if (res.second) std::memcpy(dst, p, sizeof(uint64_t) * LACC); else add_limbs<LACC>(dst, p);
                        }
                    }
                }
            }
            // erase zero coefficients (polynomial.hpp:1528-1540) and emit
            auto &ok = okeys[seg];
            auto &oc = ocfs[seg];
            ok.reserve(table.count);
            oc.reserve(table.count);
            for (size_t s = 0; s < table.cap; ++s) {
                if (!table.used[s]) continue;
                bool zero;
                if (F64) zero = (*(double *)&table.acc[s] == 0.0); else zero = ((iacc<LACC> *)&table.acc[s])->is_zero();
                if (zero) continue;
                ok.push_back(table.keys[s]);
                oc.push_back(table.acc[s]);
            }
        }
    };
    if (nthreads < 1) nthreads = 1;
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; ++t) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();

    // Flatten, grouped by segment.
    size_t total = 0;
    out->seg_offsets = (uint64_t *)std::malloc(sizeof(uint64_t) * (nsegs + 1));
    for (size_t s = 0; s < nsegs; ++s) {
        out->seg_offsets[s] = total;
        total += okeys[s].size();
    }
    out->seg_offsets[nsegs] = total;
    out->nterms = total;
    out->log2_nsegs = log2_nsegs;
    const size_t cfb = F64 ? 8 : 8 * (size_t)LACC;
    out->keys = (uint64_t *)std::malloc(std::max<size_t>(1, total * W * 8));
    out->cfs = std::malloc(std::max<size_t>(1, total * cfb));
    for (size_t s = 0; s < nsegs; ++s) {
        if (okeys[s].empty()) continue;
        std::memcpy(out->keys + out->seg_offsets[s] * W, okeys[s].data(), okeys[s].size() * W * 8);
        std::memcpy((char *)out->cfs + out->seg_offsets[s] * cfb, ocfs[s].data(), ocfs[s].size() * cfb);
    }
    return ORC_OK;
}

template <int W>
int dispatch_cf(const orc_poly *x, const orc_poly *y, const orc_trunc *tr, int nthreads, int lacc, orc_result *out)
{
    if (x->cf_kind == 1) {
        out->cf_kind = 1;
        out->cf_limbs = 1;
        return mul_impl<W, 1, 1, true>(x, y, tr, nthreads, out);
    }
    out->cf_kind = 0;
    out->cf_limbs = lacc;
    const int lin = (int)x->cf_limbs;
    if (lin == 1) {
        switch (lacc) {
            case 1: return mul_impl<W, 1, 1, false>(x, y, tr, nthreads, out);
            case 2: return mul_impl<W, 1, 2, false>(x, y, tr, nthreads, out);
            case 3: return mul_impl<W, 1, 3, false>(x, y, tr, nthreads, out);
            case 5: return mul_impl<W, 1, 5, false>(x, y, tr, nthreads, out);
        }
    } else if (lin == 2) {
        switch (lacc) {
            case 3: return mul_impl<W, 2, 3, false>(x, y, tr, nthreads, out);
            case 5: return mul_impl<W, 2, 5, false>(x, y, tr, nthreads, out);
        }
    } else if (lin == 3) {
        if (lacc == 5) return mul_impl<W, 3, 5, false>(x, y, tr, nthreads, out);
    }
    return ORC_BAD_ARG;
}

} // namespace

extern "C" {

// poly_mul_impl_mt_hm (polynomial.hpp:797-1582) behind the dispatch of poly_mul_impl_identical_ss
// (:1878-1929: empty operand -> empty result) and poly_mul_impl_switch (:2014-2022: shorter first).
// lacc: accumulator limbs for integer coefficients (1,2,3 or 5); ignored for f64.
int orc_poly_mul(const orc_poly *x, const orc_poly *y, const orc_trunc *tr, int nthreads, int lacc, orc_result *out)
{
    std::memset(out, 0, sizeof(*out));
    if (x->key_words != y->key_words || x->cf_kind != y->cf_kind || x->nvars != y->nvars || x->psize != y->psize
        || x->key_signed != y->key_signed)
        return ORC_BAD_ARG;
    if (x->cf_kind == 0 && x->cf_limbs != y->cf_limbs) return ORC_BAD_ARG;
    out->key_words = x->key_words;
    out->cf_kind = x->cf_kind;
    out->cf_limbs = x->cf_kind ? 1 : lacc;
    if (x->nterms == 0 || y->nterms == 0) {
        out->seg_offsets = (uint64_t *)std::calloc(2, sizeof(uint64_t));
        out->keys = (uint64_t *)std::malloc(8);
        out->cfs = std::malloc(8);
        return ORC_OK;
    }
    if (x->nterms > y->nterms) std::swap(x, y);
    if (!orc_overflow_check(x, y)) return ORC_KEY_OVERFLOW; // polynomial.hpp:839-851
    int rc;
    const auto t0 = std::chrono::steady_clock::now();
    switch (x->key_words) {
        case 1: rc = dispatch_cf<1>(x, y, tr, nthreads, lacc, out); break;
        case 2: rc = dispatch_cf<2>(x, y, tr, nthreads, lacc, out); break;
        case 3: rc = dispatch_cf<3>(x, y, tr, nthreads, lacc, out); break;
        case 4: rc = dispatch_cf<4>(x, y, tr, nthreads, lacc, out); break;
        default: return ORC_BAD_ARG;
    }
    out->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (rc == ORC_OK && out->seg_offsets == nullptr) {
        // truncated to nothing before segmenting
        out->seg_offsets = (uint64_t *)std::calloc(2, sizeof(uint64_t));
        out->keys = (uint64_t *)std::malloc(8);
        out->cfs = std::malloc(8);
    }
    return rc;
}

void orc_result_free(orc_result *r)
{
    std::free(r->keys);
    std::free(r->cfs);
    std::free(r->seg_offsets);
    std::memset(r, 0, sizeof(*r));
}

int orc_hardware_concurrency() { return (int)std::thread::hardware_concurrency(); }

} // extern "C"
