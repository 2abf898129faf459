// obk_engine.cu -- host orchestration + C ABI (include/obk_b200.h) of the B200 series-product engine.
//
// Replaces, behind the same seam, poly_mul_impl_identical_ss -> poly_mul_impl_mt_hm of the reference
// (include/obake/polynomials/polynomial.hpp:1878-1929, 797-1582).  There is NO CPU fallback: without a
// CUDA device every entry point fails with OBK_ERR_CUDA.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/obk_b200.h"
#include "../include/obake/kpack.hpp"
#include "obk_kernels.cuh"
#include "obk_planes.cuh"
#include "obk_tiles.cuh"

using namespace obk;

#define OBK_VERSION_STRING "obk_b200 0.1 (sm_100a)"
#define OBK_EXCHANGE_MOD 4093u

namespace
{

typedef __int128 i128;

struct PinnedBlock {
    void *ptr;
    size_t size;
    bool in_use;
    uint64_t stamp = 0; // release order (the idle block released longest ago is the one evicted)
};

struct ArenaChunk {
    char *base;
    size_t size;
};

} // namespace

struct obk_engine {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    std::string err;
    int sm_count = 148;
    size_t smem_optin = 0;
    cudaEvent_t ev[16];
    std::mutex mtx;
    std::vector<PinnedBlock> pinned;
    uint64_t pinned_clock = 0;
    // options
    int64_t opt_nsegs = -1, opt_fence = 1, opt_regroup = 1, opt_row_threads = 1024, opt_two_sided = 1, opt_ctas_per_sm = 1, opt_refill_min = 16, opt_mgpu_owner = 0, opt_table_slots = 0, opt_threads = 1024, opt_kernel = -1, opt_plane_cfg = 0, opt_tile_split = 1, opt_tile_units = 0,
            opt_max_retries = 6, opt_fill_pct = 62;
    unsigned launches = 0;
    uint32_t last_rslices = 1;
    // multi-device engine (obk_engine_create_multi): one single-device engine per GPU; this object then owns no stream
    std::vector<obk_engine *> subs;
    // scratch arena (see Scratch)
    std::vector<ArenaChunk> arena;
    size_t arena_chunk = 0, arena_off = 0;
    int arena_depth = 0;
};

namespace
{

struct ResultOwner {
    obk_engine *e;
    void *dkeys = nullptr, *dcfs = nullptr; // device allocations (cudaMallocAsync)
    int pk = -1, pc = -1;                   // pinned block indices
    uint64_t *seg_offsets = nullptr;        // malloc
    void *hk = nullptr, *hc = nullptr;      // malloc'ed host term arrays (multi-device engine)
};

#define CU_CHECK(e, call)                                                                                              \
    do {                                                                                                               \
        cudaError_t _err = (call);                                                                                     \
        if (_err != cudaSuccess) {                                                                                     \
            (e)->err = std::string("CUDA error: ") + cudaGetErrorString(_err) + " at " #call;                          \
            throw obk_fail{OBK_ERR_CUDA};                                                                              \
        }                                                                                                              \
    } while (0)

struct obk_fail {
    obk_status st;
};

// Scratch memory of one call.  Small buffers (the dozens of offset tables, flags and sorted copies a call needs) are
// bump-allocated from a per-engine arena that persists across calls -- a call used to spend more host time in
// cudaMallocAsync / cudaFreeAsync (about 60 pairs) than the GPU spent in its kernels for the small benchmarks.
// Large buffers and everything that may be handed to the caller (alloc_out) come from the driver's stream-ordered
// pool.  The arena is reset when the outermost Scratch of a call dies; all work is ordered on the engine's stream,
// so the next call may reuse it.
constexpr size_t ARENA_MAX_ALLOC = 8u << 20;
constexpr size_t ARENA_MIN_CHUNK = 64u << 20;

struct Scratch {
    obk_engine *e;
    std::vector<void *> ptrs; // stream-ordered allocations owned by this object
    explicit Scratch(obk_engine *e_) : e(e_)
    {
        ++e->arena_depth;
    }
    void *pool_alloc(size_t bytes)
    {
        void *p = nullptr;
        cudaError_t err = cudaMallocAsync(&p, bytes, e->stream);
        if (err != cudaSuccess) {
            e->err = std::string("device allocation of ") + std::to_string(bytes) + " bytes failed: " + cudaGetErrorString(err);
            cudaGetLastError();
            throw obk_fail{OBK_ERR_NOMEM};
        }
        ptrs.push_back(p);
        return p;
    }
    void *arena_alloc(size_t bytes)
    {
        for (;;) {
            if (e->arena_chunk < e->arena.size()) {
                ArenaChunk &c = e->arena[e->arena_chunk];
                if (e->arena_off + bytes <= c.size) {
                    void *p = c.base + e->arena_off;
                    e->arena_off += bytes;
                    return p;
                }
                ++e->arena_chunk;
                e->arena_off = 0;
                continue;
            }
            const size_t last = e->arena.empty() ? 0 : e->arena.back().size;
            const size_t sz = std::max(std::max(ARENA_MIN_CHUNK, 2 * last), bytes);
            void *base = nullptr;
            if (cudaMalloc(&base, sz) != cudaSuccess) {
                cudaGetLastError();
                e->err = "device allocation of the scratch arena failed";
                throw obk_fail{OBK_ERR_NOMEM};
            }
            e->arena.push_back({static_cast<char *>(base), sz});
        }
    }
    static size_t round_up(size_t bytes)
    {
        return (std::max<size_t>(bytes, 16) + 255) / 256 * 256;
    }
    template <typename T>
    T *alloc(size_t n)
    {
        const size_t bytes = round_up(n * sizeof(T));
        return static_cast<T *>(bytes <= ARENA_MAX_ALLOC ? arena_alloc(bytes) : pool_alloc(bytes));
    }
    // buffers that may outlive the call (results): never from the arena
    template <typename T>
    T *alloc_out(size_t n)
    {
        return static_cast<T *>(pool_alloc(round_up(n * sizeof(T))));
    }
    void release(void *p)
    {
        for (auto &q : ptrs)
            if (q == p && q) {
                cudaFreeAsync(q, e->stream);
                q = nullptr;
            }
        // arena memory is reclaimed wholesale at the end of the call
    }
    // hand ownership to the caller
    void detach(void *p)
    {
        for (auto &q : ptrs)
            if (q == p) {
                q = nullptr;
                return;
            }
        e->err = "internal: result buffer was not allocated with alloc_out";
        throw obk_fail{OBK_ERR_INVALID_ARG};
    }
    ~Scratch()
    {
        for (auto q : ptrs)
            if (q) cudaFreeAsync(q, e->stream);
        if (--e->arena_depth == 0) {
            if (e->arena.size() > 1) {
                // the call outgrew the arena: replace the chunks by one of the total size (rare; cudaFree synchronises)
                size_t total = 0;
                for (auto &c : e->arena) {
                    total += c.size;
                    cudaFree(c.base);
                }
                e->arena.clear();
                void *base = nullptr;
                if (cudaMalloc(&base, total) == cudaSuccess) e->arena.push_back({static_cast<char *>(base), total});
                else cudaGetLastError();
            }
            e->arena_chunk = 0;
            e->arena_off = 0;
        }
    }
};

int pinned_acquire(obk_engine *e, size_t bytes)
{
    std::lock_guard<std::mutex> g(e->mtx);
    bytes = std::max<size_t>(bytes, 64);
    int best = -1;
    for (size_t i = 0; i < e->pinned.size(); ++i) {
        auto &b = e->pinned[i];
        if (!b.in_use && b.size >= bytes && (best < 0 || b.size < e->pinned[best].size)) best = (int)i;
    }
    if (best >= 0 && e->pinned[best].size <= bytes * 2 + (1u << 20)) {
        e->pinned[best].in_use = true;
        return best;
    }
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        e->err = "pinned host allocation failed";
        throw obk_fail{OBK_ERR_NOMEM};
    }
    // reuse a dead slot if any
    for (size_t i = 0; i < e->pinned.size(); ++i)
        if (e->pinned[i].ptr == nullptr) {
            e->pinned[i] = {p, bytes, true, 0};
            return (int)i;
        }
    e->pinned.push_back({p, bytes, true, 0});
    return (int)e->pinned.size() - 1;
}

void pinned_release(obk_engine *e, int idx)
{
    std::lock_guard<std::mutex> g(e->mtx);
    if (idx < 0 || idx >= (int)e->pinned.size()) return;
    e->pinned[idx].in_use = false;
    e->pinned[idx].stamp = ++e->pinned_clock;
    // Keep at most 8 idle blocks; evict the one released longest ago.  (Evicting the block that has just been
    // released -- the first version -- meant that once eight blocks of other sizes sat idle, every call paid a
    // cudaMallocHost and a cudaFreeHost: D5 measured 6.2 ms end to end against a 1.35 ms product when it ran after
    // the sparse workloads.)
    size_t idle = 0;
    int oldest = -1;
    for (size_t i = 0; i < e->pinned.size(); ++i) {
        const auto &b = e->pinned[i];
        if (b.in_use || !b.ptr) continue;
        ++idle;
        if (oldest < 0 || b.stamp < e->pinned[oldest].stamp) oldest = (int)i;
    }
    if (idle > 8 && oldest >= 0) {
        cudaFreeHost(e->pinned[oldest].ptr);
        e->pinned[oldest] = {nullptr, 0, false, 0};
    }
}

Layout make_layout(const obk_poly_view &v)
{
    Layout L{};
    L.nvars = v.nvars;
    L.psize = v.psize;
    L.W = v.key_words;
    L.is_signed = v.key_signed ? 1u : 0u;
    L.wsize = v.psize ? v.psize : v.nvars;
    if (L.wsize == 0) {
        L.delta = 1;
        L.klim_min = 0;
        L.lim_min = 0;
        return L;
    }
    const bool b32 = v.key_bits == 32;
    if (v.key_signed) {
        if (b32) {
            L.delta = (uint64_t)obake::detail::kpack_get_delta<std::int32_t>(L.wsize);
            L.klim_min = (uint64_t)(int64_t)obake::detail::kpack_get_klims<std::int32_t>(L.wsize).first;
            L.lim_min = obake::detail::kpack_get_lims<std::int32_t>(L.wsize).first;
        } else {
            L.delta = (uint64_t)obake::detail::kpack_get_delta<std::int64_t>(L.wsize);
            L.klim_min = (uint64_t)obake::detail::kpack_get_klims<std::int64_t>(L.wsize).first;
            L.lim_min = obake::detail::kpack_get_lims<std::int64_t>(L.wsize).first;
        }
    } else {
        L.delta = b32 ? (uint64_t)obake::detail::kpack_get_delta<std::uint32_t>(L.wsize)
                      : obake::detail::kpack_get_delta<std::uint64_t>(L.wsize);
        L.klim_min = 0;
        L.lim_min = 0;
    }
    return L;
}

// (lim_min, lim_max) of one packed component for the view's key type
void component_lims(const obk_poly_view &v, unsigned wsize, i128 &lim_min, i128 &lim_max)
{
    const bool b32 = v.key_bits == 32;
    if (v.key_signed) {
        if (b32) {
            auto lm = obake::detail::kpack_get_lims<std::int32_t>(wsize);
            lim_min = lm.first;
            lim_max = lm.second;
        } else {
            auto lm = obake::detail::kpack_get_lims<std::int64_t>(wsize);
            lim_min = lm.first;
            lim_max = lm.second;
        }
    } else {
        lim_min = 0;
        lim_max = b32 ? (i128)obake::detail::kpack_get_lims<std::uint32_t>(wsize).second
                      : (i128)obake::detail::kpack_get_lims<std::uint64_t>(wsize).second;
    }
}

unsigned key_max_size(const obk_poly_view &v)
{
    const bool b32 = v.key_bits == 32;
    if (v.key_signed) return b32 ? obake::detail::kpack_max_size<std::int32_t>() : obake::detail::kpack_max_size<std::int64_t>();
    return b32 ? obake::detail::kpack_max_size<std::uint32_t>() : obake::detail::kpack_max_size<std::uint64_t>();
}

inline i128 unord(long long o, bool is_signed)
{
    return is_signed ? (i128)o : (i128)(uint64_t)((uint64_t)o ^ 0x8000000000000000ULL);
}

unsigned bitlen64(uint64_t v)
{
    unsigned b = 0;
    while (v) {
        ++b;
        v >>= 1;
    }
    return b;
}

void validate_view(obk_engine *e, const obk_poly_view *v, const char *name)
{
    auto bad = [&](const std::string &m) {
        e->err = std::string("invalid operand ") + name + ": " + m;
        throw obk_fail{OBK_ERR_INVALID_ARG};
    };
    if (!v) bad("null view");
    if (v->key_words < 1 || v->key_words > 4) {
        e->err = std::string("operand ") + name + ": key_words must be in [1,4]";
        throw obk_fail{OBK_ERR_UNSUPPORTED};
    }
    if (v->cf_kind != OBK_CF_INT && v->cf_kind != OBK_CF_F64) bad("cf_kind");
    if (v->cf_kind == OBK_CF_INT && (v->cf_limbs < 1 || v->cf_limbs > (uint32_t)MAX_CF_LIMBS)) bad("cf_limbs must be in [1,5]");
    if (v->mem != OBK_MEM_HOST && v->mem != OBK_MEM_DEVICE) bad("mem");
    if (v->nterms && (!v->keys || !v->cfs)) bad("null keys/cfs");
    if (v->key_bits != 0 && v->key_bits != 32 && v->key_bits != 64) bad("key_bits must be 32 or 64");
    const unsigned maxsz = key_max_size(*v);
    if (v->psize == 0) {
        if (v->key_words != 1) bad("packed_monomial keys have exactly one word");
        if (v->nvars > maxsz) bad("too many variables for one packed word");
    } else {
        if (v->psize > maxsz) bad("psize too large");
        const unsigned need = std::max(1u, (v->nvars + v->psize - 1) / v->psize);
        if (v->nvars && need != v->key_words) bad("key_words != ceil(nvars/psize)");
    }
    if (v->nvars > (unsigned)MAX_VARS) {
        e->err = "more than 64 variables are not supported";
        throw obk_fail{OBK_ERR_UNSUPPORTED};
    }
    if (v->nterms >= (1ull << 31)) {
        e->err = "operands with 2^31 or more terms are not supported";
        throw obk_fail{OBK_ERR_UNSUPPORTED};
    }
}

// Exclusive scan of n uint32 into out[0..n] (out[n] = total).
void device_scan(obk_engine *e, Scratch &sc, const uint32_t *in, uint32_t *out, uint64_t n)
{
    if (n <= 8 * (uint64_t)SCAN_BLOCK) {
        k_scan_small<<<1, SCAN_THREADS, 0, e->stream>>>(in, out, n);
        e->launches += 1;
        return;
    }
    const uint64_t nb = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    if (nb > (1ull << 22)) { // 2^34 items
        e->err = "scan too large";
        throw obk_fail{OBK_ERR_UNSUPPORTED};
    }
    uint32_t *bs = sc.alloc<uint32_t>(nb + 1);
    uint32_t *bp = sc.alloc<uint32_t>(nb + 2);
    k_scan_local<<<(unsigned)nb, SCAN_THREADS, 0, e->stream>>>(in, out, bs, n);
    k_scan_small<<<1, SCAN_THREADS, 0, e->stream>>>(bs, bp, nb); // any nb; bp[nb] = grand total
    k_scan_add<<<(unsigned)nb, SCAN_THREADS, 0, e->stream>>>(out, bp, n, bp + nb);
    e->launches += 3;
    sc.release(bs);
    sc.release(bp);
}

struct SortedY {
    uint64_t *keys = nullptr;
    uint64_t *cfs = nullptr;
    uint32_t *off = nullptr; // (nsegs << dbits) + 1
};

// Counting sort of an operand by (segment, degree index).  pow2: segment = hash & (m-1) (m a power of two,
// used to regroup the result the way a series stores it); otherwise segment = key mod m.
SortedY sort_operand(obk_engine *e, Scratch &sc, const uint64_t *keys, const uint64_t *cfs, const int64_t *deg, uint64_t n,
                     int W, int CW, uint32_t m, uint32_t is_signed, bool pow2, int dbits, int64_t degmin, bool with_cfs,
                     bool persistent = false)
{
    SortedY r;
    const uint64_t nbins = (uint64_t)m << dbits;
    uint32_t *cnt = sc.alloc<uint32_t>(nbins + 1);
    uint32_t *bin_of = sc.alloc<uint32_t>(n);
    uint32_t *rank_of = sc.alloc<uint32_t>(n);
    r.off = sc.alloc<uint32_t>(nbins + 1);
    r.keys = persistent ? sc.alloc_out<uint64_t>(n * W) : sc.alloc<uint64_t>(n * W);
    if (with_cfs) r.cfs = persistent ? sc.alloc_out<uint64_t>(n * CW) : sc.alloc<uint64_t>(n * CW);
    CU_CHECK(e, cudaMemsetAsync(cnt, 0, (nbins + 1) * sizeof(uint32_t), e->stream));
    const unsigned grid = (unsigned)std::min<uint64_t>((n + 255) / 256, 4096);
    if (nbins <= 8192 && n >= 8192) {
        if (n >= (1u << 21)) {
            const unsigned g2 = (unsigned)((n + 256 * 16 - 1) / (256 * 16));
            k_bin_count_smem<16><<<g2, 256, nbins * sizeof(uint32_t), e->stream>>>(keys, deg, n, W, m, is_signed, pow2 ? 1 : 0,
                                                                                  dbits, degmin, (uint32_t)nbins, cnt, bin_of, rank_of);
        } else {
            const unsigned g2 = (unsigned)((n + 256 * 2 - 1) / (256 * 2));
            k_bin_count_smem<2><<<g2, 256, nbins * sizeof(uint32_t), e->stream>>>(keys, deg, n, W, m, is_signed, pow2 ? 1 : 0,
                                                                                 dbits, degmin, (uint32_t)nbins, cnt, bin_of, rank_of);
        }
    } else {
        k_bin_count<<<grid, 256, 0, e->stream>>>(keys, deg, n, W, m, is_signed, pow2 ? 1 : 0, dbits, degmin, cnt, bin_of,
                                                 rank_of);
    }
    e->launches += 1;
    device_scan(e, sc, cnt, r.off, nbins);
    k_bin_scatter<<<grid, 256, 0, e->stream>>>(keys, with_cfs ? cfs : nullptr, nullptr, n, W, CW, r.off, bin_of, rank_of,
                                               r.keys, with_cfs ? r.cfs : nullptr, nullptr);
    e->launches += 1;
    sc.release(cnt);
    sc.release(bin_of);
    sc.release(rank_of);
    return r;
}

uint32_t prev_prime(uint64_t n)
{
    if (n < 2) return 1;
    while (n > 2 && !obake::detail::kp_is_prime(n)) --n;
    return (uint32_t)n;
}
uint32_t next_prime(uint64_t n)
{
    if (n < 2) return 1;
    while (!obake::detail::kp_is_prime(n)) ++n;
    return (uint32_t)n;
}

template <int W>
void launch_mul_w(obk_engine *e, int mode, const MulParams &P, unsigned grid, unsigned threads, size_t smem)
{
#define OBK_CASE(M)                                                                                                    \
    case M: {                                                                                                          \
        if (P.npass > 1) {                                                                                             \
            if constexpr (M == CF_ADD1 || M == CF_ADD2 || M == CF_ADD3 || M == CF_ADDF || M == CF_ADD5) {                              \
                e->err = "internal: two-sided pass in merge mode";                                                     \
                throw obk_fail{OBK_ERR_INVALID_ARG};                                                                   \
            } else {                                                                                                   \
                auto fn = k_mul<W, M, true>;                                                                          \
                CU_CHECK(e, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
                fn<<<grid, threads, smem, e->stream>>>(P);                                                             \
            }                                                                                                          \
        } else {                                                                                                       \
            auto fn = k_mul<W, M, false>;                                                                             \
            CU_CHECK(e, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));             \
            fn<<<grid, threads, smem, e->stream>>>(P);                                                                 \
        }                                                                                                              \
        break;                                                                                                         \
    }
    switch (mode) {
        OBK_CASE(CF_SYM)
        OBK_CASE(CF_I1A1)
        OBK_CASE(CF_I1A2)
        OBK_CASE(CF_I1A3)
        OBK_CASE(CF_I2A3)
        OBK_CASE(CF_F64)
        OBK_CASE(CF_ADD1)
        OBK_CASE(CF_ADD2)
        OBK_CASE(CF_ADD3)
        OBK_CASE(CF_ADDF)
        OBK_CASE(CF_I2A5)
        OBK_CASE(CF_I3A5)
        OBK_CASE(CF_ADD5)
        default:
            e->err = "internal: bad coefficient mode";
            throw obk_fail{OBK_ERR_INVALID_ARG};
    }
#undef OBK_CASE
    e->launches += 1;
    CU_CHECK(e, cudaGetLastError());
}

// Work units of k_mul: 32 left terms x one slice of their runs.  Long runs are cut so that a segment yields at
// least ~4 units per warp (a rectangular product has a handful of left terms, the owner-side merge has one).
void plan_units(obk_engine *e, MulParams &P, unsigned threads, double avg_run, uint64_t max_run)
{
    const uint64_t nch = std::max<uint64_t>(1, ((P.nx + 31) >> 5) + (P.npass > 1 ? ((P.b.nx + 31) >> 5) : 0));
    const uint64_t want = 4ull * (threads / 32);
    uint64_t R = 1;
    if (nch < want) {
        R = (want + nch - 1) / nch;
        const uint64_t by_run = (uint64_t)std::max(1.0, avg_run / 8.0);
        R = std::max<uint64_t>(1, std::min<uint64_t>(R, by_run));
    }
    // a unit's products are indexed with 32 bits: 32 slices of at most ceil(max_run / R) terms each
    R = std::max<uint64_t>(R, (max_run >> 25) + 1);
    if (nch * R >= (1ull << 31)) {
        e->err = "operands too large for the work-unit index";
        throw obk_fail{OBK_ERR_UNSUPPORTED};
    }
    P.rslices = (uint32_t)R;
    e->last_rslices = (uint32_t)R;
    P.refill_min = (uint32_t)std::max<int64_t>(1, std::min<int64_t>(32, e->opt_refill_min));
}

void launch_mul(obk_engine *e, int W, int mode, const MulParams &P, unsigned grid, unsigned threads, size_t smem)
{
    switch (W) {
        case 1: launch_mul_w<1>(e, mode, P, grid, threads, smem); break;
        case 2: launch_mul_w<2>(e, mode, P, grid, threads, smem); break;
        case 3: launch_mul_w<3>(e, mode, P, grid, threads, smem); break;
        case 4: launch_mul_w<4>(e, mode, P, grid, threads, smem); break;
        default: e->err = "unsupported key_words"; throw obk_fail{OBK_ERR_UNSUPPORTED};
    }
}

int acc_words_of(int mode)
{
    switch (mode) {
        case CF_SYM: return 0;
        case CF_I1A1: case CF_F64: case CF_ADD1: case CF_ADDF: return 1;
        case CF_I1A2: case CF_ADD2: return 2;
        case CF_I2A5: case CF_I3A5: case CF_ADD5: return 5;
        default: return 3;
    }
}

// Table geometry: the largest power-of-two slot count whose table fits the opt-in shared memory.
uint32_t pick_log2_cap(obk_engine *e, int W, int accw)
{
    if (e->opt_table_slots > 0) {
        uint32_t l = 0;
        while ((1ll << (l + 1)) <= e->opt_table_slots) ++l;
        return l;
    }
    const size_t slot = (size_t)8 * W + (size_t)8 * accw + 4;
    const size_t budget = std::min<size_t>(e->smem_optin, 227 * 1024) - 2048;
    uint32_t l = 4;
    while (((size_t)2 << l) * slot <= budget) ++l;
    return l;
}

// Two-sided truncation plan (see obk_kernels.cuh, MulParams::PassB): pass A multiplies the left terms of degree <= H
// by the degree-sorted right operand, pass B the right terms of degree <= L-H-1 by the degree-sorted left operand
// restricted to degrees > H.  Every admissible pair is formed exactly once, from its lower-degree side.
struct TwoSided {
    bool on = false;
    // left lists (compacted)
    uint64_t *ak = nullptr, *ac = nullptr, *bk = nullptr, *bc = nullptr;
    int64_t *adeg = nullptr, *bdeg = nullptr;
    uint64_t na = 0, nb = 0;
    // geometry of the degree-sorted left operand (right side of pass B)
    int dbits_x = 0;
    uint32_t ndeg_x = 1, dlo_b = 0;
    int64_t degmin_x = 0, lim_base_b = 0;
};

struct EstimateIn {
    const uint64_t *xk;
    const int64_t *xdeg;
    uint64_t nx;
    const uint64_t *yk;
    const int64_t *ydeg;
    uint64_t ny;
    int W;
    uint32_t is_signed;
    int dbits;
    int64_t degmin_y;
    uint32_t ndeg;
    bool trunc;
    int64_t lim_base;
    uint64_t tot_mults;
    unsigned threads;
    const TwoSided *ts = nullptr; // optional: two-sided truncation (xk/xdeg/nx then describe the FULL left operand)
};

// Product-size estimate (stands in for poly_mul_estimate_product_size, polynomial.hpp:478-794): a sampled
// SYMBOLIC product.  The operands are segmented by a fine prime modulus mE (about sym_cap/4 pairs per
// segment, so a sampled table can never fill up), S evenly spread segments are multiplied key-only and
// their distinct keys counted exactly; the total is the sample mean times mE.  Deterministic.
uint64_t estimate_terms(obk_engine *e, Scratch &sc, const EstimateIn &in)
{
    cudaStream_t s = e->stream;
    const int W = in.W;
    const uint64_t max_bins = 1ull << 25;
    const uint32_t sym_log2_cap = pick_log2_cap(e, W, 0);
    const uint32_t sym_cap = 1u << sym_log2_cap;
    uint64_t want = (in.tot_mults + sym_cap / 4 - 1) / (sym_cap / 4);
    want = std::min<uint64_t>(want, max_bins >> in.dbits);
    const uint32_t mE = want <= 1 ? 1u : prev_prime(want);
    const double per_seg = (double)in.tot_mults / (double)mE;
    uint64_t S = (uint64_t)std::ceil(1.5e6 / std::max(per_seg, 1.0));
    S = std::max<uint64_t>(S, 64);
    S = std::min<uint64_t>(S, 1024);
    S = std::min<uint64_t>(S, mE);
    SortedY ye = sort_operand(e, sc, in.yk, nullptr, in.ydeg, in.ny, W, 0, mE, in.is_signed, false, in.dbits, in.degmin_y,
                              false);
    uint32_t *xsegE = sc.alloc<uint32_t>(in.nx);
    k_segments<<<(unsigned)std::min<uint64_t>((in.nx + 255) / 256, 4096), 256, 0, s>>>(in.xk, in.nx, W, mE, in.is_signed,
                                                                                     xsegE);
    e->launches += 1;
    std::vector<uint32_t> hlist(S);
    for (uint64_t i = 0; i < S; ++i) hlist[i] = (uint32_t)((i * (uint64_t)mE) / S); // evenly spread, distinct
    uint32_t *dlist = sc.alloc<uint32_t>(S);
    uint32_t *dcounts = sc.alloc<uint32_t>(S);
    uint32_t *dflags = sc.alloc<uint32_t>(4);
    CU_CHECK(e, cudaMemcpyAsync(dlist, hlist.data(), S * 4, cudaMemcpyHostToDevice, s));
    CU_CHECK(e, cudaMemsetAsync(dflags, 0, 16, s));
    MulParams P{};
    P.xk = in.xk;
    P.xdeg = in.xdeg;
    P.xseg = xsegE;
    P.nx = in.nx;
    P.yk = ye.keys;
    P.yoff = ye.off;
    P.npass = 1;
    SortedY xe{};
    uint32_t *asegE = nullptr, *bsegE = nullptr;
    if (in.ts && in.ts->on) {
        const TwoSided &ts = *in.ts;
        asegE = sc.alloc<uint32_t>(std::max<uint64_t>(ts.na, 1));
        bsegE = sc.alloc<uint32_t>(std::max<uint64_t>(ts.nb, 1));
        if (ts.na) k_segments<<<(unsigned)std::min<uint64_t>((ts.na + 255) / 256, 4096), 256, 0, s>>>(ts.ak, ts.na, W, mE, in.is_signed, asegE);
        if (ts.nb) k_segments<<<(unsigned)std::min<uint64_t>((ts.nb + 255) / 256, 4096), 256, 0, s>>>(ts.bk, ts.nb, W, mE, in.is_signed, bsegE);
        xe = sort_operand(e, sc, in.xk, nullptr, in.xdeg, in.nx, W, 0, mE, in.is_signed, false, ts.dbits_x, ts.degmin_x, false);
        P.xk = ts.ak;
        P.xdeg = ts.adeg;
        P.xseg = asegE;
        P.nx = ts.na;
        P.npass = 2;
        P.b.xk = ts.bk;
        P.b.xdeg = ts.bdeg;
        P.b.xseg = bsegE;
        P.b.nx = ts.nb;
        P.b.yk = xe.keys;
        P.b.yoff = xe.off;
        P.b.dbits = ts.dbits_x;
        P.b.ndeg = ts.ndeg_x;
        P.b.dlo = ts.dlo_b;
        P.b.lim_base = ts.lim_base_b;
    }
    P.nsegs = mE;
    P.fence = e->opt_fence ? 1u : 0u;
    P.dbits = in.dbits;
    P.ndeg = in.ndeg;
    P.trunc = in.trunc;
    P.lim_base = in.lim_base;
    P.seg_list = dlist;
    P.nseg_list = (uint32_t)S;
    P.cursor = dflags + 1;
    P.log2_cap = sym_log2_cap;
    const double run = (double)in.tot_mults / (double)in.nx / (double)mE;
    (void)run;
    P.max_fill = (uint32_t)((uint64_t)sym_cap * 9 / 10);
    P.counts = dcounts;
    P.flags = dflags;
    const size_t smem_sym = (size_t)sym_cap * (8 * W + 4);
    plan_units(e, P, in.threads, (double)in.tot_mults / (double)std::max<uint64_t>(P.nx + (P.npass > 1 ? P.b.nx : 0), 1) / (double)mE,
               std::max(in.nx, in.ny));
    launch_mul(e, W, CF_SYM, P, (unsigned)std::min<uint64_t>(S, (uint64_t)e->sm_count), in.threads, smem_sym);
    std::vector<uint32_t> hcounts(S);
    CU_CHECK(e, cudaMemcpyAsync(hcounts.data(), dcounts, S * 4, cudaMemcpyDeviceToHost, s));
    CU_CHECK(e, cudaStreamSynchronize(s));
    double sum = 0;
    for (uint64_t i = 0; i < S; ++i) sum += hcounts[i] == 0xffffffffu ? (double)sym_cap : (double)hcounts[i];
    uint64_t est = (uint64_t)std::ceil(sum * (double)mE / (double)S);
    sc.release(ye.keys);
    sc.release(ye.off);
    sc.release(xsegE);
    if (xe.keys) {
        sc.release(xe.keys);
        sc.release(xe.off);
        sc.release(asegE);
        sc.release(bsegE);
    }
    sc.release(dlist);
    sc.release(dcounts);
    sc.release(dflags);
    return std::max<uint64_t>(est, 1);
}

// Segment count for a product with `est` distinct terms whose tables hold `slots_target` terms each: a prime
// modulus (flat residues) just below a whole number of waves of resident CTAs; small products get fewer.
uint32_t pick_nsegs(uint64_t est, double slots_target, uint64_t tot_mults, uint32_t wave)
{
    uint64_t needed = (uint64_t)std::ceil((double)est / slots_target);
    needed = std::max<uint64_t>(needed, 1);
    const uint64_t by_work = std::max<uint64_t>(1, std::min<uint64_t>(wave, tot_mults / 65536));
    if (needed <= 1 && by_work <= 1) return 1;
    if (needed <= wave && by_work < wave) {
        uint32_t n = next_prime(std::max<uint64_t>(needed, by_work));
        if (n > wave) n = prev_prime(wave) >= needed ? prev_prime(wave) : n;
        return n;
    }
    const uint64_t waves = (needed + wave - 1) / wave;
    uint32_t m = prev_prime(waves * wave);
    if (m < needed) m = prev_prime((waves + 1) * wave);
    return m;
}

// ------------------------------------------------------------------------------------------------
// Row-blocked dense path (obk_kernels.cuh, "kernel 1")
// ------------------------------------------------------------------------------------------------
struct RowSide {
    uint64_t *tab = nullptr;
    uint32_t log2_cap = 0;
    uint32_t *rank = nullptr;
    uint64_t *row_hi = nullptr;
    uint64_t *mat = nullptr;
    uint32_t *row_len = nullptr;
    uint32_t nrows = 0;
    uint64_t sum_len = 0;
};

void rows_phase_a(obk_engine *e, Scratch &sc, const uint64_t *keys, uint64_t n, uint64_t delta, RowSide &r)
{
    cudaStream_t s = e->stream;
    r.log2_cap = 4;
    while ((1ull << r.log2_cap) < 2 * n) ++r.log2_cap;
    const uint32_t cap = 1u << r.log2_cap;
    r.tab = sc.alloc<uint64_t>(cap);
    CU_CHECK(e, cudaMemsetAsync(r.tab, 0xFF, (size_t)cap * 8, s));
    const unsigned grid = (unsigned)std::min<uint64_t>((n + 255) / 256, 4096);
    k_row_insert<<<grid, 256, 0, s>>>(keys, n, delta, r.tab, r.log2_cap);
    uint32_t *flags = sc.alloc<uint32_t>((size_t)cap + 1);
    k_row_flags<<<(cap + 255) / 256, 256, 0, s>>>(r.tab, cap, flags);
    e->launches += 2;
    r.rank = sc.alloc<uint32_t>((size_t)cap + 2);
    device_scan(e, sc, flags, r.rank, cap);
    sc.release(flags);
}

void rows_phase_b(obk_engine *e, Scratch &sc, const uint64_t *keys, const uint64_t *cfs, uint64_t n, uint64_t delta,
                  RowSide &r)
{
    cudaStream_t s = e->stream;
    const uint32_t cap = 1u << r.log2_cap;
    r.row_hi = sc.alloc<uint64_t>(r.nrows);
    r.mat = sc.alloc<uint64_t>((size_t)r.nrows * 32);
    r.row_len = sc.alloc<uint32_t>(r.nrows);
    CU_CHECK(e, cudaMemsetAsync(r.mat, 0, (size_t)r.nrows * 32 * 8, s));
    CU_CHECK(e, cudaMemsetAsync(r.row_len, 0, (size_t)r.nrows * 4, s));
    k_row_keys<<<(cap + 255) / 256, 256, 0, s>>>(r.tab, cap, r.rank, r.row_hi);
    const unsigned grid = (unsigned)std::min<uint64_t>((n + 255) / 256, 4096);
    k_row_fill<<<grid, 256, 0, s>>>(keys, cfs, n, delta, r.tab, r.log2_cap, r.rank, r.mat, r.row_len);
    e->launches += 2;
}

__global__ void k_iota64(uint64_t *p, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

template <int LACC, bool F64, bool UNS = false>
void launch_rows_t(obk_engine *e, const RowParams &P, unsigned grid, size_t smem, const uint32_t *offs, uint64_t delta,
                   uint64_t *rk, uint64_t *rc, bool expand)
{
    if (!expand) {
        if (e->opt_row_threads <= 512) {
            auto fn = k_rowconv<LACC, F64, UNS, 512>;
            CU_CHECK(e, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
            fn<<<grid, 512, smem, e->stream>>>(P);
        } else {
            auto fn = k_rowconv<LACC, F64, UNS, 1024>;
            CU_CHECK(e, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
            fn<<<grid, 1024, smem, e->stream>>>(P);
        }
    } else {
        k_row_expand<LACC, F64><<<(unsigned)std::min<uint32_t>((P.nro + 7) / 8, 8192), 256, 0, e->stream>>>(
            P.or_hi, P.omat, offs, P.nro, delta, rk, rc);
    }
    e->launches += 1;
    CU_CHECK(e, cudaGetLastError());
}

void launch_rows(obk_engine *e, int mode, bool uns, const RowParams &P, unsigned grid, size_t smem, const uint32_t *offs,
                 uint64_t delta, uint64_t *rk, uint64_t *rc, bool expand)
{
    if (uns && !expand) {
        // no negative coefficient anywhere: unsigned multiply, no sign fix-ups, no sign extension limb
        switch (mode) {
            case CF_I1A1: launch_rows_t<1, false, true>(e, P, grid, smem, offs, delta, rk, rc, expand); return;
            case CF_I1A2: launch_rows_t<2, false, true>(e, P, grid, smem, offs, delta, rk, rc, expand); return;
            case CF_I1A3: launch_rows_t<3, false, true>(e, P, grid, smem, offs, delta, rk, rc, expand); return;
            default: break;
        }
    }
    switch (mode) {
        case CF_I1A1: launch_rows_t<1, false>(e, P, grid, smem, offs, delta, rk, rc, expand); break;
        case CF_I1A2: launch_rows_t<2, false>(e, P, grid, smem, offs, delta, rk, rc, expand); break;
        case CF_I1A3: launch_rows_t<3, false>(e, P, grid, smem, offs, delta, rk, rc, expand); break;
        case CF_F64: launch_rows_t<1, true>(e, P, grid, smem, offs, delta, rk, rc, expand); break;
        default: e->err = "internal: coefficient mode not supported by the row path"; throw obk_fail{OBK_ERR_INVALID_ARG};
    }
}


struct RowsResult {
    bool done = false;
    uint64_t *rk = nullptr, *rc = nullptr;
    uint64_t total = 0;
    // plane path: the term count stays on the device until the call's last synchronisation when nobody needs it
    // earlier (a product left in HBM, not regrouped): `total` is then an upper bound (the size of rk / rc)
    const uint32_t *total_dev = nullptr;
    uint64_t est_rows = 0;
    uint32_t nro = 0, m = 0;
    float ms_rowconv = 0;
};

// Returns done=false when the operands are not dense enough (the caller then runs the scatter kernel).
RowsResult run_rows(obk_engine *e, Scratch &sc, const uint64_t *dxk, const uint64_t *dxc, uint64_t nx, const uint64_t *dyk,
                    const uint64_t *dyc, uint64_t ny, const Layout &L, int mode, int accw, uint64_t tot_mults,
                    unsigned threads, bool forced, bool uns, uint32_t rank, uint32_t nranks, bool owner, obk_stats &st)
{
    RowsResult R;
    cudaStream_t s = e->stream;
    const uint64_t delta = L.delta;
    RowSide X, Y;
    const bool same = dxk == dyk && nx == ny;
    rows_phase_a(e, sc, dxk, nx, delta, X);
    if (!same) rows_phase_a(e, sc, dyk, ny, delta, Y);
    uint32_t hx = 0, hy = 0;
    CU_CHECK(e, cudaMemcpyAsync(&hx, X.rank + (1u << X.log2_cap), 4, cudaMemcpyDeviceToHost, s));
    if (!same) CU_CHECK(e, cudaMemcpyAsync(&hy, Y.rank + (1u << Y.log2_cap), 4, cudaMemcpyDeviceToHost, s));
    CU_CHECK(e, cudaStreamSynchronize(s));
    X.nrows = hx;
    rows_phase_b(e, sc, dxk, dxc, nx, delta, X);
    if (same) {
        Y = X;
    } else {
        Y.nrows = hy;
        rows_phase_b(e, sc, dyk, dyc, ny, delta, Y);
    }
    if (nranks > 1 && !owner) {
        // multi-GPU: this rank multiplies a contiguous slice of the left ROWS (row numbers come out of a hash
        // table, so a slice is a pseudo-random, balanced subset) by every right row
        const uint32_t b = (uint32_t)((uint64_t)X.nrows * rank / nranks);
        const uint32_t c = (uint32_t)((uint64_t)X.nrows * (rank + 1) / nranks) - b;
        X.row_hi += b;
        X.mat += (size_t)b * 32;
        X.row_len += b;
        X.nrows = c;
        if (c == 0) {
            R.done = true;
            R.rk = sc.alloc_out<uint64_t>(1);
            R.rc = sc.alloc_out<uint64_t>(accw);
            return R;
        }
    }
    // density test on the host: the convolution loops over the shorter row of every pair
    std::vector<uint32_t> lx(X.nrows), ly(Y.nrows);
    CU_CHECK(e, cudaMemcpyAsync(lx.data(), X.row_len, (size_t)X.nrows * 4, cudaMemcpyDeviceToHost, s));
    CU_CHECK(e, cudaMemcpyAsync(ly.data(), Y.row_len, (size_t)Y.nrows * 4, cudaMemcpyDeviceToHost, s));
    CU_CHECK(e, cudaStreamSynchronize(s));
    for (auto v : lx) X.sum_len += v;
    for (auto v : ly) Y.sum_len += v;
    const double iters = std::min((double)X.sum_len * (double)Y.nrows, (double)Y.sum_len * (double)X.nrows);
    // measured costs: ~40 warp instructions per convolution step vs ~22 per product in the scatter kernel
    if (!forced && iters * 40.0 * (double)(owner ? 1u : nranks) > (double)tot_mults * 22.0 * 0.7) return R;

    const uint32_t wave = (uint32_t)e->sm_count;
    // ---- symbolic row product: the set of output rows ------------------------------------------------------
    const uint64_t row_pairs = (uint64_t)X.nrows * Y.nrows;
    if (row_pairs <= (8ull << 20)) {
        // A small symbolic product (a multi-GPU slice, Fateman n <= 20) costs less than estimating it: the sampled
        // estimate is three launches and a host round trip.  Size for 32 output rows per operand row (dense operands
        // produce ~7) -- the segment count is then set by the amount of work anyway -- and let the overflow retry
        // below handle a product that turns out sparser.
        R.est_rows = std::min<uint64_t>(row_pairs, 32ull * std::max(X.nrows, Y.nrows));
    } else {
        EstimateIn ein{X.row_hi, nullptr, X.nrows, Y.row_hi, nullptr, Y.nrows, 1, 0, 0, 0, 1, false, 0, row_pairs, threads};
        R.est_rows = estimate_terms(e, sc, ein);
    }
    const uint32_t sym_log2_cap = pick_log2_cap(e, 1, 0);
    const uint32_t sym_cap = 1u << sym_log2_cap;
    uint32_t mR = pick_nsegs(R.est_rows, 0.5 * sym_cap, row_pairs, wave);
    uint32_t *counts = nullptr, *offs = nullptr;
    uint64_t *stage = nullptr;
    uint32_t nro = 0;
    for (int attempt = 0;; ++attempt) {
        SortedY ys = sort_operand(e, sc, Y.row_hi, nullptr, nullptr, Y.nrows, 1, 0, mR, 0, false, 0, 0, false);
        uint32_t *xseg = sc.alloc<uint32_t>(X.nrows);
        k_segments<<<(X.nrows + 255) / 256, 256, 0, s>>>(X.row_hi, X.nrows, 1, mR, 0, xseg);
        e->launches += 1;
        stage = sc.alloc<uint64_t>((size_t)mR * sym_cap);
        counts = sc.alloc<uint32_t>((size_t)mR + 1);
        offs = sc.alloc<uint32_t>((size_t)mR + 2);
        uint32_t *dflags = sc.alloc<uint32_t>(4);
        CU_CHECK(e, cudaMemsetAsync(dflags, 0, 16, s));
        CU_CHECK(e, cudaMemsetAsync(counts, 0, ((size_t)mR + 1) * 4, s));
        MulParams P{};
        P.xk = X.row_hi;
        P.xseg = xseg;
        P.nx = X.nrows;
        P.yk = ys.keys;
        P.yoff = ys.off;
        P.nsegs = mR;
        P.fence = e->opt_fence ? 1u : 0u;
        P.ndeg = 1;
        P.nseg_list = mR;
        P.cursor = dflags + 1;
        P.log2_cap = sym_log2_cap;
            P.max_fill = (uint32_t)((uint64_t)sym_cap * 9 / 10);
        P.okeys = stage;
        P.counts = counts;
        P.flags = dflags;
        plan_units(e, P, threads, (double)Y.nrows / (double)mR, Y.nrows);
        launch_mul(e, 1, CF_SYM, P, std::min<uint32_t>(mR, wave), threads, (size_t)sym_cap * 12);
        device_scan(e, sc, counts, offs, mR);
        uint32_t hflag[4], htot = 0;
        CU_CHECK(e, cudaMemcpyAsync(hflag, dflags, 16, cudaMemcpyDeviceToHost, s));
        CU_CHECK(e, cudaMemcpyAsync(&htot, offs + mR, 4, cudaMemcpyDeviceToHost, s));
        CU_CHECK(e, cudaStreamSynchronize(s));
        sc.release(ys.keys);
        sc.release(ys.off);
        sc.release(xseg);
        sc.release(dflags);
        if (hflag[0] == 0) {
            nro = htot;
            break;
        }
        sc.release(stage);
        sc.release(counts);
        sc.release(offs);
        if (attempt > 6) {
            e->err = "row path: symbolic row product overflowed its tables";
            throw obk_fail{OBK_ERR_TABLE_OVERFLOW};
        }
        mR = prev_prime((uint64_t)mR * 2 + wave);
    }
    uint64_t *or_hi = sc.alloc<uint64_t>(std::max<uint32_t>(nro, 1));
    uint64_t *off64 = sc.alloc<uint64_t>((size_t)mR + 1);
    k_compact<<<std::min<uint32_t>(mR, 4096), 256, 0, s>>>(stage, nullptr, counts, offs, sym_cap, 1, 0, or_hi, nullptr, off64, mR);
    e->launches += 1;
    if (owner) {
        // owner-computes: this rank forms the output rows of a contiguous range of the symbolic product's segments.
        // The CONTENT of a segment (row keys congruent mod mR) is the same on every rank, the order inside it is
        // not (it comes out of a hash table filled by racing threads), so the cut must fall on segment boundaries;
        // mR itself is a function of the operands only (estimate and retries count distinct keys).
        const uint32_t sa = (uint32_t)((uint64_t)mR * rank / nranks), sb = (uint32_t)((uint64_t)mR * (rank + 1) / nranks);
        uint32_t hb[2] = {0, 0};
        CU_CHECK(e, cudaMemcpyAsync(&hb[0], offs + sa, 4, cudaMemcpyDeviceToHost, s));
        CU_CHECK(e, cudaMemcpyAsync(&hb[1], offs + sb, 4, cudaMemcpyDeviceToHost, s));
        CU_CHECK(e, cudaStreamSynchronize(s));
        or_hi += hb[0];
        nro = hb[1] - hb[0];
    }
    sc.release(stage);
    sc.release(counts);
    sc.release(offs);
    sc.release(off64);
    // ---- gather directory: right rows sorted by a fine prime modulus (2-3 rows per run) ------------------------
    const uint32_t mG = next_prime(std::max<uint32_t>(Y.nrows / 2, 3));
    uint64_t *iota = sc.alloc<uint64_t>(Y.nrows);
    k_iota64<<<(Y.nrows + 255) / 256, 256, 0, s>>>(iota, Y.nrows);
    SortedY yg = sort_operand(e, sc, Y.row_hi, iota, nullptr, Y.nrows, 1, 1, mG, 0, false, 0, 0, true);
    uint32_t *xsegG = sc.alloc<uint32_t>(X.nrows);
    k_segments<<<(X.nrows + 255) / 256, 256, 0, s>>>(X.row_hi, X.nrows, 1, mG, 0, xsegG);
    uint32_t *or_seg = sc.alloc<uint32_t>(std::max<uint32_t>(nro, 1));
    if (nro) k_segments<<<(nro + 255) / 256, 256, 0, s>>>(or_hi, nro, 1, mG, 0, or_seg);
    e->launches += 3;

    // ---- one warp per output row ------------------------------------------------------------------------------
    uint64_t *omat = sc.alloc<uint64_t>((size_t)std::max<uint32_t>(nro, 1) * 64 * accw);
    uint32_t *or_cnt = sc.alloc<uint32_t>((size_t)nro + 1);
    uint32_t *roffs = sc.alloc<uint32_t>((size_t)nro + 2);
    uint32_t *cursor = sc.alloc<uint32_t>(4);
    CU_CHECK(e, cudaMemsetAsync(cursor, 0, 16, s));
    // packed directories, padded to the 16-byte granularity of the bulk copies
    auto pad16 = [](size_t b) { return (uint32_t)((b + 15) / 16 * 16); };
    uint32_t *xsl = sc.alloc<uint32_t>((size_t)X.nrows + 4);
    uint32_t *yil = sc.alloc<uint32_t>((size_t)Y.nrows + 4);
    k_row_pack<<<(std::max(X.nrows, Y.nrows) + 255) / 256, 256, 0, s>>>(xsegG, X.row_len, X.nrows, yg.cfs, Y.row_len, Y.nrows,
                                                                      xsl, yil);
    e->launches += 1;
    RowParams P{};
    P.xr_hi = X.row_hi;
    P.xr_sl = xsl;
    P.xmat = X.mat;
    P.nrx = X.nrows;
    P.yr_hi = yg.keys;
    P.yr_il = yil;
    P.yr_off = yg.off;
    P.ymat = Y.mat;
    P.nry = Y.nrows;
    P.m = mG;
    P.or_hi = or_hi;
    P.or_seg = or_seg;
    P.nro = nro;
    P.cursor = cursor;
    P.omat = omat;
    P.or_cnt = or_cnt;
    P.b_xhi = pad16((size_t)X.nrows * 8);
    P.b_yhi = pad16((size_t)Y.nrows * 8);
    P.b_yil = pad16((size_t)Y.nrows * 4);
    P.b_xsl = pad16((size_t)X.nrows * 4);
    P.b_yoff = pad16(((size_t)mG + 1) * 4);
    const size_t dir_bytes = (size_t)P.b_xhi + P.b_yhi + P.b_yil + P.b_xsl + P.b_yoff + 128;
    // a multi-GPU slice offsets X.row_hi by a row count: bulk copies need 16-byte aligned sources
    const bool aligned = ((uintptr_t)X.row_hi % 16 == 0);
    // the per-warp row staging (96 words x 2 buffers) and short-pair queue (64 words) sit at the start of the
    // dynamic shared memory
    const size_t stage_bytes = (size_t)(e->opt_row_threads <= 512 ? 16 : 32) * (2 * 96 + 64) * 8;
    P.dir_in_smem = (dir_bytes + stage_bytes + 2048 <= e->smem_optin && aligned) ? 1u : 0u;
    CU_CHECK(e, cudaEventRecord(e->ev[5], s));
    if (nro) launch_rows(e, mode, uns, P, wave, stage_bytes + (P.dir_in_smem ? dir_bytes : 0), nullptr, delta, nullptr, nullptr, false);
    CU_CHECK(e, cudaEventRecord(e->ev[6], s));
    st.mul_launches += 1;
    uint64_t total = 0;
    if (nro) {
        device_scan(e, sc, or_cnt, roffs, nro);
        uint32_t htot = 0;
        CU_CHECK(e, cudaMemcpyAsync(&htot, roffs + nro, 4, cudaMemcpyDeviceToHost, s));
        CU_CHECK(e, cudaStreamSynchronize(s));
        total = htot;
    }
    R.rk = sc.alloc_out<uint64_t>(std::max<uint64_t>(total, 1));
    R.rc = sc.alloc_out<uint64_t>(std::max<uint64_t>(total, 1) * accw);
    if (total) launch_rows(e, mode, uns, P, 0, 0, roffs, delta, R.rk, R.rc, true);
    R.total = total;
    R.nro = nro;
    R.m = mG;
    R.done = true;
    st.table_slots = 64;
    st.group_lanes = 32;
    return R;
}

// ------------------------------------------------------------------------------------------------
// Plane-blocked dense path (obk_planes.cuh, "kernel 2")
// ------------------------------------------------------------------------------------------------
struct PlaneSide {
    uint64_t *tab = nullptr;       // open-addressing set of the plane keys hi2 = key / delta^2
    uint32_t log2_cap = 0;
    uint32_t *slot_id = nullptr;   // slot -> plane number
    uint64_t *plane_key = nullptr; // plane number -> key
    uint64_t *mat = nullptr;       // [planes][PL_BLOB]: row-length table + rows in their staged layout
    uint32_t *pinfo = nullptr;     // [planes][4]: rows, terms, longest row, -
    uint32_t nplanes = 0;
};

// The plane sets of both operands are built BEFORE the host reads the term statistics (they only need delta), so that
// the plane counts travel with the statistics in one round trip.
struct PlanePre {
    bool launched = false, same = false;
    PlaneSide X, Y;
    uint32_t *counters = nullptr; // device: [0] planes of x, [1] planes of y
    uint32_t h[2] = {0, 0};       // host copies (valid after the next synchronisation)
};

void planes_insert(obk_engine *e, Scratch &sc, const uint64_t *keys, uint64_t n, uint64_t d2, PlaneSide &p, uint32_t *counter)
{
    cudaStream_t s = e->stream;
    p.log2_cap = 4;
    while ((1ull << p.log2_cap) < 2 * n) ++p.log2_cap;
    const uint32_t cap = 1u << p.log2_cap;
    p.tab = sc.alloc<uint64_t>(cap);
    p.slot_id = sc.alloc<uint32_t>(cap);
    p.plane_key = sc.alloc<uint64_t>(n);
    CU_CHECK(e, cudaMemsetAsync(p.tab, 0xFF, (size_t)cap * 8, s));
    k_plane_insert<<<(unsigned)std::min<uint64_t>((n + 255) / 256, 4096), 256, 0, s>>>(keys, n, d2, p.tab, p.log2_cap, p.slot_id, p.plane_key,
                                                                                     counter);
    e->launches += 1;
}

void planes_prepare(obk_engine *e, Scratch &sc, const uint64_t *dxk, uint64_t nx, const uint64_t *dyk, uint64_t ny, uint64_t delta,
                    PlanePre &pre)
{
    pre.same = dxk == dyk && nx == ny;
    pre.counters = sc.alloc<uint32_t>(4);
    CU_CHECK(e, cudaMemsetAsync(pre.counters, 0, 16, e->stream));
    planes_insert(e, sc, dxk, nx, delta * delta, pre.X, pre.counters);
    if (!pre.same) planes_insert(e, sc, dyk, ny, delta * delta, pre.Y, pre.counters + 1);
    CU_CHECK(e, cudaMemcpyAsync(pre.h, pre.counters, 8, cudaMemcpyDeviceToHost, e->stream));
    pre.launched = true;
}

void planes_build(obk_engine *e, Scratch &sc, const uint64_t *keys, const uint64_t *cfs, uint64_t n, uint64_t delta, PlaneSide &p)
{
    cudaStream_t s = e->stream;
    p.mat = sc.alloc<uint64_t>((size_t)p.nplanes * PL_BLOB);
    p.pinfo = sc.alloc<uint32_t>((size_t)p.nplanes * 4);
    CU_CHECK(e, cudaMemsetAsync(p.mat, 0, (size_t)p.nplanes * PL_BLOB * 8, s));
    CU_CHECK(e, cudaMemsetAsync(p.pinfo, 0, (size_t)p.nplanes * 16, s));
    const unsigned grid = (unsigned)std::min<uint64_t>((n + 255) / 256, 4096);
    k_plane_fill<<<grid, 256, 0, s>>>(keys, cfs, n, delta, p.tab, p.log2_cap, p.slot_id, p.mat, p.pinfo);
    e->launches += 1;
}

// Geometry of the plane kernel: consumer warps (= virtual output rows per unit) and CTAs per SM.
struct PlaneCfg {
    uint32_t nw, ctas;
};
PlaneCfg plane_cfg(const obk_engine *e)
{
    switch (e->opt_plane_cfg) {
        case 1: return {16, 1};
        case 2: return {16, 2};
        default: return {8, 3};
    }
}

template <int LACC, bool F64, bool UNS, int NW, int MINB>
void launch_planes_c(obk_engine *e, const PlaneParams &P, unsigned grid)
{
    auto fn = k_planeconv<LACC, F64, UNS, NW, MINB>;
    const size_t smem = (size_t)P.nstage * pl_stage_bytes(P.ra_max, P.rb_max);
    CU_CHECK(e, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fn<<<grid, (NW + 1) * 32, smem, e->stream>>>(P);
    e->launches += 1;
    CU_CHECK(e, cudaGetLastError());
}

template <int LACC, bool F64, bool UNS>
void launch_planes_t(obk_engine *e, const PlaneParams &P, unsigned grid)
{
    switch (e->opt_plane_cfg) {
        case 1: launch_planes_c<LACC, F64, UNS, 16, 1>(e, P, grid); break;
        case 2: launch_planes_c<LACC, F64, UNS, 16, 2>(e, P, grid); break;
        default: launch_planes_c<LACC, F64, UNS, 8, 3>(e, P, grid); break;
    }
}

void launch_planes(obk_engine *e, int mode, bool uns, const PlaneParams &P, unsigned grid)
{
    switch (mode) {
        case CF_I1A1: launch_planes_t<1, false, false>(e, P, grid); break;
        case CF_I1A2: uns ? launch_planes_t<2, false, true>(e, P, grid) : launch_planes_t<2, false, false>(e, P, grid); break;
        case CF_I1A3: uns ? launch_planes_t<3, false, true>(e, P, grid) : launch_planes_t<3, false, false>(e, P, grid); break;
        case CF_F64: launch_planes_t<1, true, false>(e, P, grid); break;
        default: e->err = "internal: coefficient mode not supported by the plane path"; throw obk_fail{OBK_ERR_INVALID_ARG};
    }
}


// Returns done=false when the operands do not qualify (the caller then tries the row path / the scatter kernel).
// Host round trips: the plane counts (with the term statistics, see PlanePre), the unit / row totals of the symbolic
// plane product, and the term count -- the last one only if `need_total`.
RowsResult run_planes(obk_engine *e, Scratch &sc, PlanePre &pre, const uint64_t *dxk, const uint64_t *dxc, uint64_t nx,
                      const uint64_t *dyk, const uint64_t *dyc, uint64_t ny, const Layout &L, int mode, int accw, bool forced, bool uns,
                      uint32_t rank, uint32_t nranks, bool owner, uint32_t rows_max_x, uint32_t rows_max_y, bool need_total,
                      obk_stats &st)
{
    RowsResult R;
    cudaStream_t s = e->stream;
    const uint64_t delta = L.delta;
    if (delta >= (1ull << 32)) return R; // delta^2 must fit 64 bits
    const PlaneCfg cfg = plane_cfg(e);
    const uint32_t stage_bytes = pl_stage_bytes(rows_max_x, rows_max_y);
    // shared memory of one SM (228 KB) shared by cfg.ctas resident CTAs, 1 KB reserved per CTA
    const uint32_t nstage = (uint32_t)std::min<size_t>(PL_MAX_STAGES, (std::min<size_t>(e->smem_optin, (228u << 10) / cfg.ctas - 1024) - 512) / stage_bytes);
    if (nstage < 2 || rows_max_x > (uint32_t)PL_ROWS || rows_max_y > (uint32_t)PL_ROWS) return R;
    if (!pre.launched) {
        planes_prepare(e, sc, dxk, nx, dyk, ny, delta, pre);
        CU_CHECK(e, cudaStreamSynchronize(s));
    }
    PlaneSide &X = pre.X, &Y = pre.Y;
    const uint32_t hx = pre.h[0], hy = pre.same ? pre.h[0] : pre.h[1];
    // dense enough?  a plane costs 16 KB of HBM and a staged pair is only worth its copy with enough terms in it
    const uint64_t plane_pairs = (uint64_t)hx * hy;
    const bool dense = (double)nx >= 8.0 * hx && (double)ny >= 8.0 * hy;
    if (plane_pairs > (1ull << 21) || hx >= (1u << 24) || hy >= (1u << 24) || (!forced && !dense)) return R;
    X.nplanes = hx;
    planes_build(e, sc, dxk, dxc, nx, delta, X);
    if (pre.same) {
        Y = X;
    } else {
        Y.nplanes = hy;
        planes_build(e, sc, dyk, dyc, ny, delta, Y);
    }
    // multi-GPU partial product (left operand sharded): this rank forms the pairs of the left planes it owns
    const uint32_t part_nranks = (nranks > 1 && !owner) ? nranks : 1u;
    const uint64_t npairs = plane_pairs;
    // ---- symbolic plane product ------------------------------------------------------------------------------
    uint32_t l2 = 4;
    while ((1ull << l2) < 2 * npairs) ++l2;
    const uint32_t cap = 1u << l2;
    uint64_t *tab = sc.alloc<uint64_t>(cap);
    // one zeroed block: per-slot records and work, per-plane fill cursors, counters
    const size_t zero_words = (size_t)cap * 4 + (size_t)cap * 2 + npairs + 8;
    uint32_t *zero = sc.alloc<uint32_t>(zero_words);
    uint4 *rec = reinterpret_cast<uint4 *>(zero);
    unsigned long long *work = reinterpret_cast<unsigned long long *>(zero + (size_t)cap * 4);
    uint32_t *fill = zero + (size_t)cap * 6;
    uint32_t *counters = fill + npairs; // [0] output planes, [1] product-kernel unit cursor
    CU_CHECK(e, cudaMemsetAsync(tab, 0xFF, (size_t)cap * 8, s));
    CU_CHECK(e, cudaMemsetAsync(zero, 0, zero_words * 4, s));
    // everything indexed by output plane is sized by the bound nout <= npairs
    uint64_t *okey = sc.alloc<uint64_t>(npairs);
    uint32_t *ocnt = sc.alloc<uint32_t>(npairs + 1), *onrows = sc.alloc<uint32_t>(npairs + 1);
    unsigned long long *owork = sc.alloc<unsigned long long>(npairs);
    uint32_t *pair_off = sc.alloc<uint32_t>(npairs + 2), *row_base = sc.alloc<uint32_t>(npairs + 2);
    uint2 *pairs = sc.alloc<uint2>(npairs);
    uint32_t *totals = sc.alloc<uint32_t>(4);
    const uint32_t blocks_max = (2 * (2 * PL_ROWS - 1) + cfg.nw - 1) / cfg.nw; // at most 63 rows x 2 halves per plane
    uint32_t *units = sc.alloc<uint32_t>(std::max<uint64_t>(npairs * blocks_max, 1));
    const unsigned gp = (unsigned)std::min<uint64_t>((npairs + 255) / 256, 2048);
    k_plane_sym_insert<<<gp, 256, 0, s>>>(X.plane_key, X.pinfo, X.nplanes, Y.plane_key, Y.pinfo, Y.nplanes, rank, part_nranks, tab, l2, rec,
                                          work, okey, counters);
    k_plane_sym_compact<<<(cap + 255) / 256, 256, 0, s>>>(tab, cap, rec, work, ocnt, owork, onrows);
    k_plane_units<<<1, 1024, 0, s>>>(owork, onrows, ocnt, okey, counters, cfg.nw, rank, owner ? nranks : 1u, row_base, pair_off, units, totals);
    k_plane_sym_fill<<<gp, 256, 0, s>>>(X.plane_key, X.nplanes, Y.plane_key, Y.nplanes, rank, part_nranks, tab, l2, rec, pair_off, fill, pairs);
    e->launches += 4;
    uint32_t htot[3] = {0, 0, 0};
    CU_CHECK(e, cudaMemcpyAsync(htot, totals, 12, cudaMemcpyDeviceToHost, s));
    CU_CHECK(e, cudaStreamSynchronize(s));
    const uint32_t nunits = htot[0], nro = htot[1], nout = htot[2];
    // ---- the product ----------------------------------------------------------------------------------------
    uint64_t *omat = sc.alloc<uint64_t>((size_t)std::max<uint32_t>(nro, 1) * 64 * accw);
    uint64_t *or_hi = sc.alloc<uint64_t>(std::max<uint32_t>(nro, 1));
    uint32_t *or_cnt = sc.alloc<uint32_t>(2 * (size_t)nro + 1);
    uint32_t *roffs = sc.alloc<uint32_t>(2 * (size_t)nro + 2);
    CU_CHECK(e, cudaMemsetAsync(or_cnt, 0, (2 * (size_t)nro + 1) * 4, s)); // rows of units this rank does not own stay empty
    PlaneParams P{};
    P.xmat = X.mat;
    P.ymat = Y.mat;
    P.xinfo = X.pinfo;
    P.yinfo = Y.pinfo;
    P.pairs = pairs;
    P.pair_off = pair_off;
    P.okey = okey;
    P.onrows = onrows;
    P.row_base = row_base;
    P.units = units;
    P.nunits = nunits;
    P.cursor = counters + 1;
    P.delta = delta;
    P.or_hi = or_hi;
    P.omat = omat;
    P.or_cnt = or_cnt;
    P.ra_max = rows_max_x;
    P.rb_max = rows_max_y;
    P.nstage = nstage;
    const unsigned grid = (unsigned)std::max<uint32_t>(1, std::min<uint32_t>(nunits, (uint32_t)e->sm_count * cfg.ctas));
    CU_CHECK(e, cudaEventRecord(e->ev[5], s));
    launch_planes(e, mode, uns, P, grid);
    CU_CHECK(e, cudaEventRecord(e->ev[6], s));
    st.mul_launches += 1;
    uint64_t total = 0;
    if (nro) {
        device_scan(e, sc, or_cnt, roffs, 2 * (uint64_t)nro);
        if (need_total) {
            uint32_t ht = 0;
            CU_CHECK(e, cudaMemcpyAsync(&ht, roffs + 2 * (size_t)nro, 4, cudaMemcpyDeviceToHost, s));
            CU_CHECK(e, cudaStreamSynchronize(s));
            total = ht;
        } else {
            total = (uint64_t)nro * 64; // bound: the arrays are sized for it, the count is read with the call's last sync
            R.total_dev = roffs + 2 * (size_t)nro;
        }
    }
    R.rk = sc.alloc_out<uint64_t>(std::max<uint64_t>(total, 1));
    R.rc = sc.alloc_out<uint64_t>(std::max<uint64_t>(total, 1) * accw);
    if (total) {
        const unsigned ge = (unsigned)std::min<uint32_t>((2 * nro + 7) / 8, 8192);
        switch (mode) {
            case CF_I1A1: k_plane_expand<1, false><<<ge, 256, 0, s>>>(or_hi, omat, roffs, nro, delta, R.rk, R.rc); break;
            case CF_I1A2: k_plane_expand<2, false><<<ge, 256, 0, s>>>(or_hi, omat, roffs, nro, delta, R.rk, R.rc); break;
            case CF_I1A3: k_plane_expand<3, false><<<ge, 256, 0, s>>>(or_hi, omat, roffs, nro, delta, R.rk, R.rc); break;
            default: k_plane_expand<1, true><<<ge, 256, 0, s>>>(or_hi, omat, roffs, nro, delta, R.rk, R.rc); break;
        }
        e->launches += 1;
        CU_CHECK(e, cudaGetLastError());
    }
    R.total = total;
    R.nro = nro;
    R.m = nout;
    R.est_rows = nro;
    R.done = true;
    st.table_slots = 64;
    st.group_lanes = 32;
    return R;
}

// ------------------------------------------------------------------------------------------------
// Tile-blocked dense path (obk_tiles.cuh, "kernel 3")
// ------------------------------------------------------------------------------------------------
struct TileCfg {
    uint32_t nw, ctas; // consumer warps per CTA, CTAs per SM
};
TileCfg tile_cfg(const obk_engine *e, uint32_t tiles_max)
{
    switch (e->opt_plane_cfg) {
        case 1: return {10, 2};
        case 2: return {16, 1};
        case 3: return {7, 3};
        case 4: return {12, 1};
        case 5: return {15, 2};
        case 7: return {20, 1};
        // Measured (F30, F20: output planes of up to 121 / 66 tiles in the bounding box; D5: 32): large planes want the deeper ring and the
        // larger accumulator area of 2 CTAs per SM, small ones the cheaper end-of-unit barrier of 3 small CTAs.
        default: return tiles_max <= 32u ? TileCfg{7, 3} : TileCfg{10, 2};
    }
}

template <int LACC, bool F64, bool UNS>
void launch_tiles_t(obk_engine *e, const TileParams &P, unsigned grid, unsigned threads, unsigned ctas, size_t smem)
{
    // register budget by geometry: the <= 64-register build when the resident CTAs would not fit the register file at 80
    auto fn = threads * ctas * 80u > 65536u ? k_tileconv<LACC, F64, UNS, 512, 2> : k_tileconv<LACC, F64, UNS, TL_MAX_THREADS, 1>;
    CU_CHECK(e, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fn<<<grid, threads, smem, e->stream>>>(P);
    e->launches += 1;
    CU_CHECK(e, cudaGetLastError());
}

void launch_tiles(obk_engine *e, int mode, bool uns, const TileParams &P, unsigned grid, unsigned threads, unsigned ctas, size_t smem)
{
    switch (mode) {
        case CF_I1A1: launch_tiles_t<1, false, false>(e, P, grid, threads, ctas, smem); break;
        case CF_I1A2: uns ? launch_tiles_t<2, false, true>(e, P, grid, threads, ctas, smem) : launch_tiles_t<2, false, false>(e, P, grid, threads, ctas, smem); break;
        case CF_I1A3: uns ? launch_tiles_t<3, false, true>(e, P, grid, threads, ctas, smem) : launch_tiles_t<3, false, false>(e, P, grid, threads, ctas, smem); break;
        case CF_F64: launch_tiles_t<1, true, false>(e, P, grid, threads, ctas, smem); break;
        default: e->err = "internal: coefficient mode not supported by the tile path"; throw obk_fail{OBK_ERR_INVALID_ARG};
    }
}

struct TileDims {
    uint32_t wx, wy, rx, ry; // row width (first symbol's largest exponent + 1) and rows (second symbol's) of each operand
    // truncation as an output filter: on, weights of the two low symbols, limit, the symbols it counts
    bool trunc = false;
    uint32_t w0 = 0, w1 = 0;
    int64_t limit = 0;
    uint64_t var_mask = 0;
};

// dense enough for the plane / tile paths?  (a plane costs KBs of HBM and a staged pair is only worth its copy with
// enough terms in it)
bool planes_dense(const PlanePre &pre, uint64_t nx, uint64_t ny)
{
    const uint32_t hx = pre.h[0], hy = pre.same ? pre.h[0] : pre.h[1];
    return (double)nx >= 8.0 * hx && (double)ny >= 8.0 * hy && (uint64_t)hx * hy <= (1ull << 21);
}
// The tile path pays a fixed cost per staged plane pair (measured: ~2-9 ns chip-wide) against ~15 ps per term product
// saved over the scatter kernel: it wins from a few hundred term products per plane pair on (D5: 292 -> 1.0 ms vs
// 2.3 ms; S12: 185, where most pairs feed an output plane of their own -> 1.9 ms vs 0.64 ms).
bool tiles_dense(const PlanePre &pre, uint64_t nx, uint64_t ny)
{
    const uint32_t hx = pre.h[0], hy = pre.same ? pre.h[0] : pre.h[1];
    return planes_dense(pre, nx, ny) && (double)nx * (double)ny >= 256.0 * (double)hx * (double)hy;
}

// Returns done=false when the operands do not qualify (the caller then tries the other paths).
RowsResult run_tiles(obk_engine *e, Scratch &sc, PlanePre &pre, const uint64_t *dxk, const uint64_t *dxc, uint64_t nx,
                     const uint64_t *dyk, const uint64_t *dyc, uint64_t ny, const Layout &L, int mode, int accw, bool forced, bool uns,
                     uint32_t rank, uint32_t nranks, bool owner, const TileDims &D, bool need_total, obk_stats &st)
{
    RowsResult R;
    cudaStream_t s = e->stream;
    const uint64_t delta = L.delta;
    // OBK_TILE_TIMING=1: device time between the phases of this path (development aid)
    static const bool timing = std::getenv("OBK_TILE_TIMING") != nullptr;
    std::vector<std::pair<const char *, cudaEvent_t>> marks;
    std::vector<std::chrono::steady_clock::time_point> hmarks; // host time when the mark was enqueued
    auto mark = [&](const char *what) {
        if (!timing) return;
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        hmarks.push_back(std::chrono::steady_clock::now());
        cudaEventRecord(ev, s);
        marks.emplace_back(what, ev);
    };
    mark("start");
    if (delta >= (1ull << 32)) return R; // delta^2 must fit 64 bits
    if (D.wx > (uint32_t)TL_MAX_DIM || D.wy > (uint32_t)TL_MAX_DIM || D.rx > (uint32_t)TL_MAX_DIM || D.ry > (uint32_t)TL_MAX_DIM) return R;
    const uint32_t P = tl_pitch(std::max(D.wx, D.wy)), rcap = std::max(D.rx, D.ry);
    const uint32_t blob_words = tl_blob_words(rcap, P);
    const TileStage G = tl_stage(rcap, P);
    // Shared memory of one SM (228 KB) shared by cfg.ctas resident CTAs, 1 KB reserved per CTA: the tile accumulators of
    // a unit (all tiles of the largest output plane if they fit) and a ring of at least 3 staged plane pairs.
    const uint32_t wo_raw = D.wx + D.wy - 1u, ro_raw = D.rx + D.ry - 1u;
    const uint32_t shape_max = tl_shape(ro_raw, wo_raw, std::min(ro_raw + wo_raw - 1u, 255u));
    const uint32_t tiles_max = tl_ntiles(ro_raw, wo_raw, std::min(ro_raw + wo_raw - 1u, 255u));
    TileCfg cfg = tile_cfg(e, tiles_max);
    uint32_t slots = 0, nstage = 0;
    size_t per_cta = 0;
    for (;; --cfg.ctas) {
        per_cta = std::min<size_t>(e->smem_optin, (228u << 10) / cfg.ctas - 1024) - 512;
        slots = std::max<uint32_t>(16u, tiles_max);
        while (slots > 16u && tl_acc_bytes(slots, accw) + 3u * (size_t)G.bytes > per_cta) --slots;
        const size_t accb = tl_acc_bytes(slots, accw);
        nstage = accb < per_cta ? (uint32_t)std::min<size_t>(TL_MAX_STAGES, (per_cta - accb) / G.bytes) : 0;
        if (nstage >= 3 || cfg.ctas == 1) break;
    }
    if (nstage < 2) return R;
    if (!pre.launched) {
        planes_prepare(e, sc, dxk, nx, dyk, ny, delta, pre);
        CU_CHECK(e, cudaStreamSynchronize(s));
    }
    PlaneSide &X = pre.X, &Y = pre.Y;
    const uint32_t hx = pre.h[0], hy = pre.same ? pre.h[0] : pre.h[1];
    const uint64_t plane_pairs = (uint64_t)hx * hy;
    if (plane_pairs > (1ull << 21) || hx >= (1u << 24) || hy >= (1u << 24) || (!forced && !tiles_dense(pre, nx, ny))) return R;
    auto build = [&](const uint64_t *keys, const uint64_t *cfs, uint64_t n, PlaneSide &p) {
        // one zeroed block: blobs, plane info, row lengths
        const size_t mat_words = (size_t)p.nplanes * blob_words, info_words = ((size_t)p.nplanes * 4 + 1) / 2, len_words = ((size_t)p.nplanes * rcap + 1) / 2;
        p.mat = sc.alloc<uint64_t>(mat_words + info_words + len_words);
        p.pinfo = reinterpret_cast<uint32_t *>(p.mat + mat_words);
        uint32_t *rowlen = reinterpret_cast<uint32_t *>(p.mat + mat_words + info_words);
        CU_CHECK(e, cudaMemsetAsync(p.mat, 0, (mat_words + info_words + len_words) * 8, s));
        const unsigned grid = (unsigned)std::min<uint64_t>((n + 255) / 256, 4096);
        k_tile_fill<<<grid, 256, 0, s>>>(keys, cfs, n, delta, p.tab, p.log2_cap, p.slot_id, p.mat, blob_words, P, rcap, rowlen, p.pinfo);
        k_tile_finish<<<(p.nplanes + 7) / 8, 256, 0, s>>>(p.nplanes, rcap, rowlen, p.mat, blob_words, p.pinfo);
        e->launches += 2;
    };
    X.nplanes = hx;
    build(dxk, dxc, nx, X);
    if (pre.same) {
        Y = X;
    } else {
        Y.nplanes = hy;
        build(dyk, dyc, ny, Y);
    }
    mark("planes built");
    // multi-GPU partial product (left operand sharded): this rank forms the pairs of the left planes it owns
    const uint32_t part_nranks = (nranks > 1 && !owner) ? nranks : 1u;
    const uint64_t npairs = plane_pairs;
    // ---- symbolic plane product ------------------------------------------------------------------------------
    uint32_t l2 = 4;
    while ((1ull << l2) < 2 * npairs) ++l2;
    const uint32_t cap = 1u << l2;
    uint64_t *tab = sc.alloc<uint64_t>(cap);
    // one zeroed block: per-slot records, diagonal bounds and work, per-plane fill cursors, counters
    const size_t zero_words = (size_t)cap * 4 + (size_t)cap * 2 + cap + npairs + 8;
    uint32_t *zero = sc.alloc<uint32_t>(zero_words);
    uint4 *rec = reinterpret_cast<uint4 *>(zero);
    unsigned long long *work = reinterpret_cast<unsigned long long *>(zero + (size_t)cap * 4);
    uint32_t *dmax = zero + (size_t)cap * 6;
    uint32_t *fill = dmax + cap;
    uint32_t *counters = fill + npairs; // [0] output planes, [1] product-kernel unit cursor
    CU_CHECK(e, cudaMemsetAsync(tab, 0xFF, (size_t)cap * 8, s));
    CU_CHECK(e, cudaMemsetAsync(zero, 0, zero_words * 4, s));
    uint64_t *okey = sc.alloc<uint64_t>(npairs);
    uint32_t *ocnt = sc.alloc<uint32_t>(npairs + 1), *oshape = sc.alloc<uint32_t>(npairs + 1);
    unsigned long long *owork = sc.alloc<unsigned long long>(npairs);
    uint32_t *pair_off = sc.alloc<uint32_t>(npairs + 2), *row_base = sc.alloc<uint32_t>(npairs + 2);
    uint2 *pairs = sc.alloc<uint2>(npairs);
    uint32_t *totals = sc.alloc<uint32_t>(8);
    const uint32_t groups_max = tl_ngroups(shape_max, 16u);
    // a plane yields at most groups x chunks units, and chunks hold at least 4 pairs each: <= groups_max * (1 + pairs / 4)
    uint2 *units = sc.alloc<uint2>(std::max<uint64_t>(npairs * groups_max + (npairs / 4 + 1) * groups_max, 1));
    const unsigned gp = (unsigned)std::min<uint64_t>((npairs + 255) / 256, 2048);
    k_tile_sym_insert<<<gp, 256, 0, s>>>(X.plane_key, X.pinfo, X.nplanes, Y.plane_key, Y.pinfo, Y.nplanes, rank, part_nranks, tab, l2, rec,
                                         dmax, work, okey, counters);
    k_tile_sym_compact<<<(cap + 255) / 256, 256, 0, s>>>(tab, cap, rec, dmax, work, ocnt, owork, oshape);
    TileTrunc tr{};
    if (D.trunc) {
        int64_t *dhi = sc.alloc<int64_t>(npairs);
        k_tile_plane_degrees<<<gp, 256, 0, s>>>(okey, counters, delta, L.nvars, D.var_mask, dhi);
        e->launches += 1;
        tr.on = 1;
        tr.w0 = D.w0;
        tr.w1 = D.w1;
        tr.limit = D.limit;
        tr.dhi = dhi;
    }
    {
        uint32_t *qbase = sc.alloc<uint32_t>(npairs + 1);
        const size_t sort_bytes = owner && nranks > 1 ? (size_t)TL_SORT_MAX * 20 : 0;
        if (sort_bytes) CU_CHECK(e, cudaFuncSetAttribute(k_tile_units, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_bytes));
        k_tile_units<<<1, 1024, sort_bytes, s>>>(owork, oshape, ocnt, okey, counters, slots, (uint32_t)std::max<int64_t>(1, e->opt_tile_units) * (uint32_t)e->sm_count * cfg.ctas,
                                                 e->opt_tile_units <= 0 ? 1u : 0u, rank,
                                                 owner ? nranks : 1u, e->opt_tile_split != 0 ? 1u : 0u, tr, qbase, row_base, pair_off, units, totals);
    }
    k_plane_sym_fill<<<gp, 256, 0, s>>>(X.plane_key, X.nplanes, Y.plane_key, Y.nplanes, rank, part_nranks, tab, l2, rec, pair_off, fill, pairs);
    e->launches += 4;
    mark("symbolic product");
    uint32_t htot[5] = {0, 0, 0, 0, 0};
    CU_CHECK(e, cudaMemcpyAsync(htot, totals, 20, cudaMemcpyDeviceToHost, s));
    CU_CHECK(e, cudaStreamSynchronize(s));
    const uint32_t nunits = htot[0], nro = htot[1], nout = htot[2];
    slots = htot[3]; // tiles per unit chosen on the device (<= the accumulator area); the ring takes what it leaves
    nstage = (uint32_t)std::min<size_t>(TL_MAX_STAGES, (per_cta - tl_acc_bytes(slots, accw)) / G.bytes);
    // ---- the product ----------------------------------------------------------------------------------------
    const bool split = htot[4] != 0;
    const bool is_f64 = mode == CF_F64;
    const uint32_t ls = split && !is_f64 ? 2u * (uint32_t)accw : (uint32_t)accw; // words per output entry
    const uint32_t wo = (wo_raw + 31u) & ~31u, nch = wo / 32u;
    const size_t omat_words = (size_t)std::max<uint32_t>(nro, 1) * wo * ls;
    uint64_t *omat = sc.alloc<uint64_t>(omat_words);
    uint64_t *or_hi = sc.alloc<uint64_t>(std::max<uint32_t>(nro, 1));
    const uint64_t nitems = (uint64_t)nro * nch;
    if (nitems >= (1ull << 31)) {
        e->err = "dense product too large for the tile path's row index";
        throw obk_fail{OBK_ERR_UNSUPPORTED};
    }
    uint32_t *or_cnt = sc.alloc<uint32_t>(nitems + 1);
    uint32_t *roffs = sc.alloc<uint32_t>(nitems + 2);
    CU_CHECK(e, cudaMemsetAsync(omat, 0, omat_words * 8, s)); // tiles outside an output plane's support are never written
    CU_CHECK(e, cudaMemsetAsync(or_cnt, 0, (nitems + 1) * 4, s));
    TileParams Pm{};
    Pm.xmat = X.mat;
    Pm.ymat = Y.mat;
    Pm.xinfo = X.pinfo;
    Pm.yinfo = Y.pinfo;
    Pm.pairs = pairs;
    Pm.pair_off = pair_off;
    Pm.okey = okey;
    Pm.oshape = oshape;
    Pm.row_base = row_base;
    Pm.units = units;
    Pm.nunits = nunits;
    Pm.cursor = counters + 1;
    Pm.delta = delta;
    Pm.or_hi = or_hi;
    Pm.omat = omat;
    Pm.or_cnt = or_cnt;
    Pm.P = P;
    Pm.rcap = rcap;
    Pm.blob_words = blob_words;
    Pm.nstage = nstage;
    Pm.wo = wo;
    Pm.slots = slots;
    Pm.split = split ? 1u : 0u;
    Pm.tr = tr;
    const unsigned grid = (unsigned)std::max<uint32_t>(1, std::min<uint32_t>(nunits, (uint32_t)e->sm_count * cfg.ctas));
    mark("sync + output set-up");
    CU_CHECK(e, cudaEventRecord(e->ev[5], s));
    if (nunits) launch_tiles(e, mode, uns, Pm, grid, (cfg.nw + 1) * 32, cfg.ctas, tl_acc_bytes(slots, accw) + (size_t)nstage * G.bytes);
    CU_CHECK(e, cudaEventRecord(e->ev[6], s));
    mark("k_tileconv");
    st.mul_launches += 1;
    uint64_t total = 0;
    const unsigned ge = (unsigned)std::min<uint64_t>((nitems + 7) / 8, 8192);
    if (nro) {
        if (split) {
            switch (mode) {
                case CF_I1A1: k_tile_norm<1, false><<<ge, 256, 0, s>>>(omat, (uint32_t)nitems, nch, wo, or_cnt); break;
                case CF_I1A2: k_tile_norm<2, false><<<ge, 256, 0, s>>>(omat, (uint32_t)nitems, nch, wo, or_cnt); break;
                case CF_I1A3: k_tile_norm<3, false><<<ge, 256, 0, s>>>(omat, (uint32_t)nitems, nch, wo, or_cnt); break;
                default: k_tile_norm<1, true><<<ge, 256, 0, s>>>(omat, (uint32_t)nitems, nch, wo, or_cnt); break;
            }
            e->launches += 1;
        }
        device_scan(e, sc, or_cnt, roffs, nitems);
        if (need_total) {
            uint32_t ht = 0;
            CU_CHECK(e, cudaMemcpyAsync(&ht, roffs + nitems, 4, cudaMemcpyDeviceToHost, s));
            CU_CHECK(e, cudaStreamSynchronize(s));
            total = ht;
        } else {
            total = (uint64_t)nro * wo; // bound: the arrays are sized for it, the count is read with the call's last sync
            R.total_dev = roffs + nitems;
        }
    }
    R.rk = sc.alloc_out<uint64_t>(std::max<uint64_t>(total, 1));
    R.rc = sc.alloc_out<uint64_t>(std::max<uint64_t>(total, 1) * accw);
    if (total) {
        switch (mode) {
            case CF_I1A1: k_tile_expand<1, false><<<ge, 256, 0, s>>>(or_hi, omat, roffs, (uint32_t)nitems, nch, wo, ls, delta, R.rk, R.rc); break;
            case CF_I1A2: k_tile_expand<2, false><<<ge, 256, 0, s>>>(or_hi, omat, roffs, (uint32_t)nitems, nch, wo, ls, delta, R.rk, R.rc); break;
            case CF_I1A3: k_tile_expand<3, false><<<ge, 256, 0, s>>>(or_hi, omat, roffs, (uint32_t)nitems, nch, wo, ls, delta, R.rk, R.rc); break;
            default: k_tile_expand<1, true><<<ge, 256, 0, s>>>(or_hi, omat, roffs, (uint32_t)nitems, nch, wo, ls, delta, R.rk, R.rc); break;
        }
        e->launches += 1;
        CU_CHECK(e, cudaGetLastError());
    }
    mark("normalise + scan + expand");
    if (timing) {
        cudaStreamSynchronize(s);
        std::string line = "[obk tile timing]";
        for (size_t k = 1; k < marks.size(); ++k) {
            float ms = 0;
            cudaEventElapsedTime(&ms, marks[k - 1].second, marks[k].second);
            char buf[96];
            std::snprintf(buf, sizeof(buf), " %s %.3f ms (host %.3f) |", marks[k].first, ms,
                          std::chrono::duration<double, std::milli>(hmarks[k] - hmarks[k - 1]).count());
            line += buf;
        }
        std::fprintf(stderr, "%s units %u planes %u rows %u split %d\n", line.c_str(), nunits, nout, nro, (int)split);
        for (auto &m : marks) cudaEventDestroy(m.second);
    }
    R.total = total;
    R.nro = nro;
    R.m = nout;
    R.est_rows = nro;
    R.done = true;
    st.table_slots = slots;  // tiles per unit
    st.group_lanes = nstage; // ring depth
    st.reserved = nunits;
    return R;
}

void set_empty_result(obk_poly_result *out, const obk_poly_view *x, int cf_limbs, uint32_t mem)
{
    std::memset(out, 0, sizeof(*out));
    out->key_words = x->key_words;
    out->cf_kind = x->cf_kind;
    out->cf_limbs = cf_limbs;
    out->mem = mem;
    out->nsegs = 1;
    out->seg_kind = 0;
    auto *own = new ResultOwner{nullptr};
    own->seg_offsets = (uint64_t *)std::calloc(2, sizeof(uint64_t));
    out->seg_offsets = own->seg_offsets;
    out->owner = own;
}

struct TailIn {
    uint64_t *rk, *rc; // device term arrays (scratch-owned)
    uint64_t total;
    int W, accw;
    bool f64;
    uint32_t is_signed;
    uint64_t tot_mults; // sparsity hint for the series-table sizing rule
    uint32_t result_mem;
    uint32_t nranks; // > 1 and !owner: multi-GPU partial product, grouped by key mod OBK_EXCHANGE_MOD
    bool owner;
    uint32_t nsegs_out; // out: groups when a multi-GPU partial
    // non-null: `total` is only a bound, the term count is still on the device and is fetched with the last
    // synchronisation of the call (only for products left in DEVICE memory as they are: no regrouping, no partial)
    const uint32_t *total_dev = nullptr;
    int force_log2 = -1; // >= 0: regroup into exactly 2^force_log2 series tables (shares of a multi-device product)
};

// Common tail of every entry point that returns terms: optional regrouping (series tables for the host, exchange
// groups for a multi-GPU partial), result ownership, device -> host copy.  Records events 7 and 8.
void emit_result(obk_engine *e, Scratch &sc, TailIn &in, obk_poly_result *out, obk_stats &st)
{
    cudaStream_t s = e->stream;
    uint64_t *rk = in.rk, *rc = in.rc;
    const uint64_t total = in.total, tot_mults = in.tot_mults;
    const int W = in.W, accw = in.accw;
    const bool f64 = in.f64, owner = in.owner;
    uint32_t nsegs = 1;
    // ---- regroup the way a series stores its terms: table hash & (2^k - 1) (series.hpp:396-400) ---------
    // k follows the reference's sizing rule (polynomial.hpp:870-902).  Partial products of the multi-GPU
    // path keep the engine's own segmentation (key mod nsegs): the all-to-all needs owner ranges contiguous.
    uint32_t out_groups = nsegs, out_log2 = 0, seg_kind = 1;
    // A device-resident product keeps the engine's own segmentation (seg_kind OBK_SEG_KEY_MOD); the series-table
    // grouping is produced when the result is handed to the host (the container rebuild needs it) or on request
    // (option regroup = 2).  regroup = 0 never regroups.
    const bool regroup = (in.nranks <= 1 || owner) && (e->opt_regroup == 2 || (e->opt_regroup == 1 && in.result_mem == OBK_MEM_HOST));
    if (in.nranks > 1 && !owner) {
        // multi-GPU partial product: group by key mod OBK_EXCHANGE_MOD, a fixed prime every rank uses, so the
        // rows destined to one owner are contiguous whatever kernel and segmentation produced them
        const uint32_t mx = OBK_EXCHANGE_MOD;
        auto *own = new ResultOwner{e};
        out->owner = own;
        own->seg_offsets = (uint64_t *)std::calloc((size_t)mx + 1, sizeof(uint64_t));
        if (total > 0) {
            SortedY g = sort_operand(e, sc, rk, rc, nullptr, total, W, accw, mx, in.is_signed, false, 0, 0, true, true);
            sc.release(rk);
            sc.release(rc);
            rk = g.keys;
            rc = g.cfs;
            std::vector<uint32_t> h32((size_t)mx + 1);
            CU_CHECK(e, cudaMemcpyAsync(h32.data(), g.off, h32.size() * 4, cudaMemcpyDeviceToHost, s));
            CU_CHECK(e, cudaStreamSynchronize(s));
            sc.release(g.off);
            for (size_t i = 0; i <= mx; ++i) own->seg_offsets[i] = h32[i];
        }
        out_groups = mx;
        nsegs = mx;
        seg_kind = 1;
    }
    if (regroup) {
        const double term_bytes = 8.0 * W + 8.0 * accw;
        const double est_sp = tot_mults ? (double)total / (double)tot_mults : 1.0;
        const double seg_kb = est_sp >= 1e-3 ? 200.0 : 20.0;
        const uint64_t est_nsegs = (uint64_t)((double)total * term_bytes / (seg_kb * 1024.0));
        while ((est_nsegs >> out_log2) != 0 && out_log2 < 22) ++out_log2; // nbits()
        if (in.force_log2 >= 0) out_log2 = (uint32_t)in.force_log2;
        if (total > 0 && out_log2 > 0) {
            SortedY g = sort_operand(e, sc, rk, rc, nullptr, total, W, accw, 1u << out_log2, 0, true, 0, 0, true, true);
            sc.release(rk);
            sc.release(rc);
            rk = g.keys;
            rc = g.cfs;
            // widen the 32-bit offsets
            std::vector<uint32_t> h32((1u << out_log2) + 1);
            CU_CHECK(e, cudaMemcpyAsync(h32.data(), g.off, h32.size() * 4, cudaMemcpyDeviceToHost, s));
            CU_CHECK(e, cudaStreamSynchronize(s));
            sc.release(g.off);
            out_groups = 1u << out_log2;
            seg_kind = 0;
            auto *own = new ResultOwner{e};
            out->owner = own;
            own->seg_offsets = (uint64_t *)std::malloc(sizeof(uint64_t) * ((size_t)out_groups + 1));
            for (size_t i = 0; i <= out_groups; ++i) own->seg_offsets[i] = h32[i];
        } else {
            out_groups = 1;
            out_log2 = 0;
            seg_kind = 0;
            auto *own = new ResultOwner{e};
            out->owner = own;
            own->seg_offsets = (uint64_t *)std::malloc(sizeof(uint64_t) * 2);
            own->seg_offsets[0] = 0;
            own->seg_offsets[1] = total;
        }
    }
    CU_CHECK(e, cudaEventRecord(e->ev[7], s));

    ResultOwner *own = static_cast<ResultOwner *>(out->owner);
    if (!own) {
        // a product left on the device is one group: both kernels write their terms in no particular order
        own = new ResultOwner{e};
        out->owner = own;
        out_groups = 1;
        own->seg_offsets = (uint64_t *)std::malloc(sizeof(uint64_t) * 2);
        own->seg_offsets[0] = 0;
        own->seg_offsets[1] = total;
    }
    st.d2h_bytes += ((uint64_t)out_groups + 1) * 8;
    out->nterms = total;
    out->key_words = W;
    out->cf_kind = f64 ? OBK_CF_F64 : OBK_CF_INT;
    out->cf_limbs = f64 ? 1 : accw;
    out->log2_nsegs = out_log2;
    out->nsegs = out_groups;
    out->seg_kind = seg_kind;
    out->mem = in.result_mem;
    out->seg_offsets = own->seg_offsets;
    st.log2_nsegs = out_log2;
    if (in.result_mem == OBK_MEM_DEVICE) {
        sc.detach(rk);
        sc.detach(rc);
        own->dkeys = rk;
        own->dcfs = rc;
        out->keys = rk;
        out->cfs = rc;
    } else {
        own->pk = pinned_acquire(e, total * W * 8);
        own->pc = pinned_acquire(e, total * accw * 8);
        out->keys = (uint64_t *)e->pinned[own->pk].ptr;
        out->cfs = e->pinned[own->pc].ptr;
        if (total) {
            CU_CHECK(e, cudaMemcpyAsync(out->keys, rk, total * W * 8, cudaMemcpyDeviceToHost, s));
            CU_CHECK(e, cudaMemcpyAsync(out->cfs, rc, total * accw * 8, cudaMemcpyDeviceToHost, s));
        }
        st.d2h_bytes += total * (W + accw) * 8;
    }
    uint32_t pend = 0;
    if (in.total_dev) CU_CHECK(e, cudaMemcpyAsync(&pend, in.total_dev, 4, cudaMemcpyDeviceToHost, s));
    CU_CHECK(e, cudaEventRecord(e->ev[8], s));
    CU_CHECK(e, cudaStreamSynchronize(s));
    if (in.total_dev) {
        if (regroup || in.result_mem != OBK_MEM_DEVICE || (in.nranks > 1 && !owner)) {
            e->err = "internal: pending term count on a path that needs it early";
            throw obk_fail{OBK_ERR_INVALID_ARG};
        }
        in.total = pend;
        out->nterms = pend;
        own->seg_offsets[1] = pend;
        st.nterms_out = pend;
    }
    in.nsegs_out = nsegs;
    in.rk = rk;
    in.rc = rc;
}

struct MulRequest {
    const obk_poly_view *x, *y;
    const obk_trunc *t;
    uint32_t result_mem;
    // sharding of the left operand (multi-GPU): this call multiplies x[shard_rank] slice only
    uint32_t rank = 0, nranks = 1;
    uint64_t *send_counts = nullptr;
    // push mode (multi-GPU): the partial terms go straight to their owners' windows, nothing is returned
    const obk_peer_windows *push = nullptr;
    // merge mode: x is ignored, y holds acc-limb coefficients to be summed by key
    bool merge = false;
    uint32_t merge_nsegs = 1;
};

obk_status run_mul(obk_engine *e, MulRequest rq, obk_poly_result *out, obk_stats *stats)
{
    obk_stats st{};
    e->launches = 0;
    const obk_poly_view *x = rq.x, *y = rq.y;
    if (!out) {
        e->err = "null result pointer";
        return OBK_ERR_INVALID_ARG;
    }
    std::memset(out, 0, sizeof(*out));
    try {
        CU_CHECK(e, cudaSetDevice(e->device));
        validate_view(e, y, "y");
        if (!rq.merge) {
            validate_view(e, x, "x");
            if (x->key_words != y->key_words || x->key_signed != y->key_signed || x->nvars != y->nvars
                || x->psize != y->psize || x->cf_kind != y->cf_kind || (x->key_bits == 32) != (y->key_bits == 32)) {
                e->err = "operands have different key layouts or coefficient kinds";
                return OBK_ERR_INVALID_ARG;
            }
            // empty operand -> empty result (polynomial.hpp:1892-1895)
            if (x->nterms == 0 || y->nterms == 0) {
                set_empty_result(out, x, x->cf_kind == OBK_CF_F64 ? 1 : 1, rq.result_mem);
                if (rq.send_counts) std::memset(rq.send_counts, 0, sizeof(uint64_t) * rq.nranks);
                if (stats) *stats = st;
                return OBK_OK;
            }
            // shorter operand first (poly_mul_impl_switch, polynomial.hpp:2014-2022)
            if (x->nterms > y->nterms) std::swap(x, y);
        } else {
            x = y;
            if (y->nterms == 0) {
                set_empty_result(out, y, y->cf_limbs, rq.result_mem);
                if (stats) *stats = st;
                return OBK_OK;
            }
        }
        const int W = (int)y->key_words;
        const bool f64 = y->cf_kind == OBK_CF_F64;
        int CWx = f64 ? 1 : (int)x->cf_limbs, CWy = f64 ? 1 : (int)y->cf_limbs;
        const Layout L = make_layout(*y);
        Scratch sc(e);
        cudaStream_t s = e->stream;
        CU_CHECK(e, cudaEventRecord(e->ev[0], s));

        // ---- operands on the device -------------------------------------------------------------------
        const uint64_t nx_full = x->nterms, ny = y->nterms;
        const uint64_t *dxk, *dxc, *dyk, *dyc;
        auto upload = [&](const obk_poly_view *v, int cw, const uint64_t *&dk, const uint64_t *&dc) {
            if (v->mem == OBK_MEM_DEVICE) {
                dk = v->keys;
                dc = (const uint64_t *)v->cfs;
            } else {
                uint64_t *k = sc.alloc<uint64_t>(v->nterms * W);
                uint64_t *c = sc.alloc<uint64_t>(v->nterms * cw);
                CU_CHECK(e, cudaMemcpyAsync(k, v->keys, v->nterms * W * 8, cudaMemcpyHostToDevice, s));
                CU_CHECK(e, cudaMemcpyAsync(c, v->cfs, v->nterms * cw * 8, cudaMemcpyHostToDevice, s));
                st.h2d_bytes += v->nterms * (W + cw) * 8;
                dk = k;
                dc = c;
            }
        };
        upload(y, CWy, dyk, dyc);
        if (rq.merge) {
            dxk = nullptr;
            dxc = nullptr;
        } else if (x == y) {
            dxk = dyk;
            dxc = dyc;
        } else {
            upload(x, CWx, dxk, dxc);
        }
        CU_CHECK(e, cudaEventRecord(e->ev[1], s));

        // ---- term statistics: overflow check, degrees, coefficient width ----------------------------
        int mode;
        uint32_t log2_nsegs = 0;
        int dbits = 0;
        uint32_t ndeg = 1;
        int64_t lim_base = 0, degmin_y = 0;
        int64_t *xdeg = nullptr, *ydeg = nullptr;
        TwoSided ts;
        // multi-GPU "owner-computes" mode (zero-exchange alternative): both operands are used whole on every
        // rank and rank r forms only the output segments / output rows it owns -- no partial products, no exchange.
        const bool owner = rq.nranks > 1 && e->opt_mgpu_owner != 0 && !rq.merge;
        const uint64_t *sxk = nullptr, *sxc = nullptr; // this rank's slice of the left operand
        const int64_t *sxdeg = nullptr;
        uint64_t snx = 0, x_begin = 0;
        bool trunc = false;
        unsigned hs_anyneg = 1;
        bool rows_ok = false; // every exponent >= 0 and the first symbol's exponents < 32 in both operands
        bool planes_ok = false; // ... and the second symbol's exponents < 32 as well (plane-blocked path)
        uint32_t plane_rows[2] = {PL_ROWS, PL_ROWS}; // largest second-symbol exponent + 1 of each operand
        bool tiles_ok = false; // every exponent >= 0 and the two lowest symbols' exponents < 64 (tile-blocked path)
        bool tiles_take = false; // a truncated product that the tile path will take (no two-sided plan needed)
        TileDims tile_dims;
        PlanePre plane_pre;
        uint64_t tot_mults = 0;
        uint64_t var_mask_t = 0; // symbols a truncation counts
        if (!rq.merge) {
            trunc = rq.t != nullptr;
            uint64_t var_mask = 0;
            if (trunc) {
                if (rq.t->partial) {
                    for (uint32_t i = 0; i < rq.t->nidx; ++i) {
                        if (rq.t->idx[i] >= L.nvars) {
                            e->err = "partial-degree index outside the symbol set";
                            return OBK_ERR_INVALID_ARG;
                        }
                        var_mask |= 1ull << rq.t->idx[i];
                    }
                } else {
                    var_mask = L.nvars >= 64 ? ~0ull : ((1ull << L.nvars) - 1);
                }
                var_mask_t = var_mask;
                xdeg = sc.alloc<int64_t>(nx_full);
                ydeg = (x == y) ? xdeg : sc.alloc<int64_t>(ny);
            } else {
                // untruncated: the statistics pass still reports the total-degree extrema (d_packed degree check)
                var_mask = L.nvars >= 64 ? ~0ull : ((1ull << L.nvars) - 1);
            }
            const bool full_mask = var_mask == (L.nvars >= 64 ? ~0ull : ((1ull << L.nvars) - 1));
            StatsOut *dstats = sc.alloc<StatsOut>(2);
            k_stats_init<<<2, 64, 0, s>>>(dstats, 2);
            // NOTE: k_stats_init indexes by global thread; fix up: one block per struct
            const unsigned gx = (unsigned)std::min<uint64_t>((nx_full + 255) / 256, 1024);
            const unsigned gy = (unsigned)std::min<uint64_t>((ny + 255) / 256, 1024);
            k_term_stats<<<gx, 256, 0, s>>>(dxk, dxc, nx_full, L, var_mask, f64 ? 1 : 0, CWx, xdeg, dstats);
            if (x != y)
                k_term_stats<<<gy, 256, 0, s>>>(dyk, dyc, ny, L, var_mask, f64 ? 1 : 0, CWy, ydeg, dstats + 1);
            e->launches += 3;
            // dense candidates: number the planes of both operands now, so that their counts come back with the
            // statistics (one host round trip instead of two)
            if (!trunc && W == 1 && L.nvars >= 3 && L.delta < (1ull << 32) && (e->opt_kernel == -1 || e->opt_kernel == 2 || e->opt_kernel == 3)
                && (e->opt_kernel >= 2 || (double)nx_full * (double)ny >= (double)(1ull << 22)))
                planes_prepare(e, sc, dxk, nx_full, dyk, ny, L.delta, plane_pre);
            StatsOut hs[2];
            CU_CHECK(e, cudaMemcpyAsync(hs, dstats, sizeof(StatsOut) * (x == y ? 1 : 2), cudaMemcpyDeviceToHost, s));
            CU_CHECK(e, cudaStreamSynchronize(s));
            if (x == y) hs[1] = hs[0];
            st.d2h_bytes += sizeof(StatsOut) * 2;
            if (!f64) {
                // Input limbs are whatever the VALUES need (1, 2 or 3), not what the caller's arrays carry: a product
                // left in HBM with a wide accumulator is a valid operand of the next product as long as its
                // coefficients fit three limbs, and operands of different widths are brought to a common one.
                const unsigned bits_in = std::max(hs[0].cfbits, hs[1].cfbits) + 1;
                const int need = (int)std::max(1u, (bits_in + 63) / 64);
                if (need > 3) {
                    e->err = "input coefficients beyond 3 limbs are not supported: the operands need " + std::to_string(bits_in)
                             + " bits";
                    return OBK_ERR_UNSUPPORTED;
                }
                auto renorm = [&](const uint64_t *&dc, uint64_t n, int cw) {
                    uint64_t *o = sc.alloc<uint64_t>(n * need);
                    k_widen_cfs<<<(unsigned)std::min<uint64_t>((n + 255) / 256, 4096), 256, 0, s>>>(dc, n, cw, need, 0, 0, o);
                    e->launches += 1;
                    dc = o;
                };
                const bool same_c = dxc == dyc;
                if (CWy != need) renorm(dyc, ny, CWy);
                if (same_c) dxc = dyc;
                else if (CWx != need) renorm(dxc, nx_full, CWx);
                CWx = CWy = need;
            }
            // multi-GPU: from here on this rank only sees its slice of the left operand's terms;
            // the checks above and the accumulator width below use the FULL operands so that every rank agrees.
            if (rq.nranks > 1 && !owner) {
                x_begin = nx_full * rq.rank / rq.nranks;
                snx = nx_full * (rq.rank + 1) / rq.nranks - x_begin;
                sxk = dxk + x_begin * W;
                sxc = dxc + x_begin * CWx;
                if (xdeg) sxdeg = xdeg + x_begin;
                // snx may be 0 (more ranks than left-operand terms): the rank still takes every decision below --
                // overflow verdict, accumulator width -- so that all ranks agree, and returns an empty share after them
            } else {
                sxk = dxk;
                sxc = dxc;
                sxdeg = xdeg;
                snx = nx_full;
            }

            // monomial_range_overflow_check (packed_monomial.hpp:462-487; d_packed_monomial.hpp:861-894)
            if (L.nvars) {
                i128 lim_min, lim_max;
                component_lims(*y, L.wsize, lim_min, lim_max);
                bool ok = true;
                for (uint32_t v = 0; v < L.nvars && ok; ++v) {
                    const i128 mn = unord(hs[0].emin[v], L.is_signed) + unord(hs[1].emin[v], L.is_signed);
                    const i128 mx = unord(hs[0].emax[v], L.is_signed) + unord(hs[1].emax[v], L.is_signed);
                    if (mn < lim_min || mx > lim_max) ok = false;
                }
                if (ok && L.psize != 0) {
                    // d_packed: the total degree of a product must stay representable in T.  The total-degree
                    // extrema are only computed when var_mask covers every variable; recompute cheaply from
                    // the per-variable extrema as a conservative bound only if that bound fails.
                    i128 dmax = 0, dmin = 0;
                    if (full_mask) {
                        // the reference's own quantity: extrema of the total degrees (d_packed_monomial.hpp:880-894)
                        dmax = unord(hs[0].dmax, L.is_signed) + unord(hs[1].dmax, L.is_signed);
                        dmin = unord(hs[0].dmin, L.is_signed) + unord(hs[1].dmin, L.is_signed);
                    } else {
                        for (uint32_t v = 0; v < L.nvars; ++v) {
                            dmax += unord(hs[0].emax[v], L.is_signed) + unord(hs[1].emax[v], L.is_signed);
                            dmin += unord(hs[0].emin[v], L.is_signed) + unord(hs[1].emin[v], L.is_signed);
                        }
                    }
                    const bool b32 = y->key_bits == 32;
                    const i128 tmax = L.is_signed ? (b32 ? (i128)INT32_MAX : (i128)INT64_MAX)
                                                  : (b32 ? (i128)UINT32_MAX : (i128)UINT64_MAX);
                    const i128 tmin = L.is_signed ? (b32 ? (i128)INT32_MIN : (i128)INT64_MIN) : (i128)0;
                    if (dmax > tmax || dmin < tmin) ok = false; // cannot trigger for psize >= 2 (sums of < 2^33 values)
                }
                if (!ok) {
                    e->err = "An overflow in the monomial exponents was detected while attempting to multiply two "
                             "polynomials";
                    return OBK_ERR_KEY_OVERFLOW;
                }
            }
            hs_anyneg = hs[0].anyneg | hs[1].anyneg;
            if (L.nvars) {
                rows_ok = unord(hs[0].emax[0], L.is_signed) < 32 && unord(hs[1].emax[0], L.is_signed) < 32
                          && hs[0].nonfinite == 0 && hs[1].nonfinite == 0;
                for (uint32_t v = 0; v < L.nvars; ++v)
                    if (unord(hs[0].emin[v], L.is_signed) < 0 || unord(hs[1].emin[v], L.is_signed) < 0) rows_ok = false;
                planes_ok = L.nvars >= 3 && unord(hs[0].emax[1], L.is_signed) < 32 && unord(hs[1].emax[1], L.is_signed) < 32;
                if (planes_ok && rows_ok) {
                    plane_rows[0] = (uint32_t)unord(hs[0].emax[1], L.is_signed) + 1u;
                    plane_rows[1] = (uint32_t)unord(hs[1].emax[1], L.is_signed) + 1u;
                }
                tiles_ok = L.nvars >= 3 && hs[0].nonfinite == 0 && hs[1].nonfinite == 0;
                for (uint32_t v = 0; v < L.nvars; ++v)
                    if (unord(hs[0].emin[v], L.is_signed) < 0 || unord(hs[1].emin[v], L.is_signed) < 0) tiles_ok = false;
                for (int o = 0; o < 2 && tiles_ok; ++o)
                    for (uint32_t v = 0; v < 2; ++v)
                        if (unord(hs[o].emax[v], L.is_signed) >= TL_MAX_DIM) tiles_ok = false;
                if (tiles_ok) {
                    tile_dims.wx = (uint32_t)unord(hs[0].emax[0], L.is_signed) + 1u;
                    tile_dims.wy = (uint32_t)unord(hs[1].emax[0], L.is_signed) + 1u;
                    tile_dims.rx = (uint32_t)unord(hs[0].emax[1], L.is_signed) + 1u;
                    tile_dims.ry = (uint32_t)unord(hs[1].emax[1], L.is_signed) + 1u;
                }
            }
            // a-priori accumulator width: |sum| <= min(nx,ny) * max|c1| * max|c2|
            if (f64) {
                mode = CF_F64;
                st.acc_limbs = 1;
            } else {
                const unsigned bits = hs[0].cfbits + hs[1].cfbits + bitlen64(std::min(nx_full, ny)) + 1;
                const unsigned lacc = (bits + 63) / 64;
                if (lacc > (unsigned)MAX_CF_LIMBS) {
                    e->err = "coefficient overflow beyond 5 limbs: the exact product needs " + std::to_string(bits)
                             + " bits";
                    return OBK_ERR_CF_OVERFLOW;
                }
                if (CWx != CWy) {
                    e->err = "integer operands must use the same number of input limbs";
                    return OBK_ERR_INVALID_ARG;
                }
                if (CWx == 1) {
                    mode = lacc == 1 ? CF_I1A1 : (lacc == 2 ? CF_I1A2 : CF_I1A3);
                    st.acc_limbs = lacc;
                } else if (CWx == 2) {
                    mode = lacc <= 3 ? CF_I2A3 : CF_I2A5;
                    st.acc_limbs = lacc <= 3 ? 3 : 5;
                } else {
                    mode = CF_I3A5;
                    st.acc_limbs = 5;
                }
            }
            if (snx == 0) {
                // nothing to multiply on this rank; the stats carry the limb count every other rank uses
                set_empty_result(out, x, f64 ? 1 : (int)st.acc_limbs, rq.result_mem);
                if (rq.send_counts) std::memset(rq.send_counts, 0, sizeof(uint64_t) * rq.nranks);
                st.total_launches = e->launches;
                if (stats) *stats = st;
                return OBK_OK;
            }
            // A limit no pair can exceed (limit >= max degree of x + max degree of y; the reference's own
            // sparse_02_truncated benchmark runs with one) is no truncation: drop the degree sort, the per-term degree
            // lookups and the (segment x degree) offset table, and let dense operands take the row path.
            if (trunc) {
                const i128 lim0 = L.is_signed ? (i128)rq.t->limit : (i128)(uint64_t)rq.t->limit;
                const i128 top = unord(hs[0].dmax, L.is_signed) + unord(hs[1].dmax, L.is_signed);
                const i128 big0 = (i128)1 << 62;
                if (top < big0 && top > -big0 && lim0 >= top) trunc = false;
            }
            // truncation geometry
            if (trunc) {
                const i128 dminx = unord(hs[0].dmin, L.is_signed), dmaxx = unord(hs[0].dmax, L.is_signed);
                const i128 dminy = unord(hs[1].dmin, L.is_signed), dmaxy = unord(hs[1].dmax, L.is_signed);
                const i128 big = (i128)1 << 62;
                if (dminx < -big || dmaxx > big || dminy < -big || dmaxy > big) {
                    e->err = "degrees beyond 2^62 are not supported in truncated products";
                    return OBK_ERR_UNSUPPORTED;
                }
                const i128 range = dmaxy - dminy + 1;
                if (range > ((i128)1 << 20)) {
                    e->err = "degree range of the right operand exceeds 2^20 (unsupported in truncated products)";
                    return OBK_ERR_UNSUPPORTED;
                }
                ndeg = (uint32_t)range;
                while ((1u << dbits) < ndeg) ++dbits;
                degmin_y = (int64_t)dminy;
                const i128 lim = L.is_signed ? (i128)rq.t->limit : (i128)(uint64_t)rq.t->limit;
                i128 lb = lim - dminy; // right-degree index d passes for left term i iff d <= lb - deg_i
                // clamp: below (dminx - 1) nothing passes, above (dmaxx + ndeg) everything passes
                lb = std::max(lb, dminx - 1);
                lb = std::min(lb, dmaxx + (i128)ndeg);
                lim_base = (int64_t)lb;
                // tot_n_mults (polynomial.hpp:574-623)
                uint32_t *hist = sc.alloc<uint32_t>(ndeg + 1);
                uint32_t *pref = sc.alloc<uint32_t>(ndeg + 2);
                unsigned long long *dtot = sc.alloc<unsigned long long>(1);
                CU_CHECK(e, cudaMemsetAsync(hist, 0, (ndeg + 1) * 4, s));
                CU_CHECK(e, cudaMemsetAsync(dtot, 0, 8, s));
                k_deg_hist<<<gy, 256, ndeg <= 8192 ? ndeg * 4 : 0, s>>>(ydeg, ny, degmin_y, hist, ndeg <= 8192 ? ndeg : 0);
                e->launches += 1;
                device_scan(e, sc, hist, pref, ndeg);
                k_count_mults<<<gx, 256, 0, s>>>(sxdeg, snx, lim_base, ndeg, pref, dtot);
                e->launches += 1;
                unsigned long long htot = 0;
                CU_CHECK(e, cudaMemcpyAsync(&htot, dtot, 8, cudaMemcpyDeviceToHost, s));
                // A truncated product of DENSE operands goes to the tile path, where the truncation is a filter on the
                // output position: number the planes now so that their counts arrive with the next synchronisation.
                const bool tile_cand = W == 1 && tiles_ok && CWx == 1 && L.delta < (1ull << 32) && (e->opt_kernel == -1 || e->opt_kernel == 3)
                                       && (mode == CF_I1A1 || mode == CF_I1A2 || mode == CF_I1A3 || mode == CF_F64)
                                       && (e->opt_kernel == 3 || (double)nx_full * (double)ny >= (double)(1ull << 22));
                if (tile_cand) planes_prepare(e, sc, dxk, nx_full, dyk, ny, L.delta, plane_pre);
                // two-sided plan: degree histograms of both operands on the host
                const i128 rangex = dmaxx - dminx + 1;
                std::vector<uint32_t> hpy, hpx;
                uint32_t *prefx = nullptr;
                if (e->opt_two_sided != 0 && rangex <= ((i128)1 << 20) && snx >= 4096) {
                    const uint32_t ndx = (uint32_t)rangex;
                    uint32_t *histx = sc.alloc<uint32_t>(ndx + 1);
                    prefx = sc.alloc<uint32_t>(ndx + 2);
                    CU_CHECK(e, cudaMemsetAsync(histx, 0, (ndx + 1) * 4, s));
                    k_deg_hist<<<gx, 256, ndx <= 8192 ? ndx * 4 : 0, s>>>(sxdeg, snx, (int64_t)dminx, histx, ndx <= 8192 ? ndx : 0);
                    e->launches += 1;
                    device_scan(e, sc, histx, prefx, ndx);
                    hpx.resize((size_t)ndx + 1);
                    hpy.resize((size_t)ndeg + 1);
                    CU_CHECK(e, cudaMemcpyAsync(hpx.data(), prefx, hpx.size() * 4, cudaMemcpyDeviceToHost, s));
                    CU_CHECK(e, cudaMemcpyAsync(hpy.data(), pref, hpy.size() * 4, cudaMemcpyDeviceToHost, s));
                }
                CU_CHECK(e, cudaStreamSynchronize(s));
                tot_mults = htot;
                tiles_take = tile_cand && (e->opt_kernel == 3 || tiles_dense(plane_pre, nx_full, ny));
                if (!hpx.empty() && tot_mults > 0 && !tiles_take) {
                    // choose the left-degree threshold H that minimises the number of left terms visited
                    const uint32_t ndx = (uint32_t)rangex;
                    uint64_t best = ~0ull;
                    i128 bestH = dminx - 1;
                    for (i128 H = dminx - 1; H <= dmaxx; ++H) {
                        const uint64_t na = hpx[(size_t)(H - dminx + 1)];
                        const i128 yl = lim - H - 1 - dminy; // right-degree index bound of pass B's left list
                        const uint64_t nb = yl < 0 ? 0 : hpy[(size_t)std::min<i128>(yl, (i128)ndeg - 1) + 1];
                        if (na + nb < best) {
                            best = na + nb;
                            bestH = H;
                        }
                    }
                    if (best * 2 < snx || e->opt_two_sided == 2) {
                        ts.on = true;
                        const int64_t H = (int64_t)bestH;
                        const i128 yl128 = lim - bestH - 1;
                        const int64_t YL = (int64_t)std::max<i128>(std::min<i128>(yl128, dmaxy), dminy - 1);
                        ts.na = hpx[(size_t)(bestH - dminx + 1)];
                        ts.nb = YL < (int64_t)dminy ? 0 : hpy[(size_t)(YL - (int64_t)dminy) + 1];
                        ts.degmin_x = (int64_t)dminx;
                        ts.ndeg_x = ndx;
                        while ((1u << ts.dbits_x) < ndx) ++ts.dbits_x;
                        ts.dlo_b = (uint32_t)std::min<i128>(std::max<i128>(bestH + 1 - dminx, 0), (i128)ndx);
                        i128 lbb = lim - dminx; // pass B: left-degree index d passes for right term j iff d <= lbb - deg_j
                        lbb = std::max(lbb, dminy - 1);
                        lbb = std::min(lbb, dmaxy + (i128)ndx);
                        ts.lim_base_b = (int64_t)lbb;
                        // compact the two left lists
                        auto select = [&](const uint64_t *k, const uint64_t *c, const int64_t *d, uint64_t n, int cw, int64_t thr,
                                          uint64_t cnt, uint64_t *&ok, uint64_t *&oc, int64_t *&od) {
                            uint32_t *flag = sc.alloc<uint32_t>(n + 1);
                            uint32_t *pos = sc.alloc<uint32_t>(n + 2);
                            const unsigned g = (unsigned)std::min<uint64_t>((n + 255) / 256, 4096);
                            k_flag_le<<<g, 256, 0, s>>>(d, n, thr, flag);
                            device_scan(e, sc, flag, pos, n);
                            ok = sc.alloc<uint64_t>(std::max<uint64_t>(cnt, 1) * W);
                            oc = sc.alloc<uint64_t>(std::max<uint64_t>(cnt, 1) * cw);
                            od = sc.alloc<int64_t>(std::max<uint64_t>(cnt, 1));
                            k_select<<<g, 256, 0, s>>>(k, c, d, n, W, cw, pos, ok, oc, od);
                            e->launches += 2;
                            sc.release(flag);
                            sc.release(pos);
                        };
                        select(sxk, sxc, sxdeg, snx, CWx, H, ts.na, ts.ak, ts.ac, ts.adeg);
                        select(dyk, dyc, ydeg, ny, CWy, YL, ts.nb, ts.bk, ts.bc, ts.bdeg);
                        if (ts.dlo_b >= ndx || ts.nb == 0) ts.nb = 0; // pass B has nothing to do
                    }
                }
                if (tot_mults == 0) {
                    // the truncation admits no pair: empty result before segmenting (polynomial.hpp:860-862)
                    set_empty_result(out, x, st.acc_limbs, rq.result_mem);
                    if (rq.send_counts) std::memset(rq.send_counts, 0, sizeof(uint64_t) * rq.nranks);
                    st.total_launches = e->launches;
                    if (stats) *stats = st;
                    return OBK_OK;
                }
            } else {
                tot_mults = snx * ny;
            }
        } else {
            // merge mode: sum acc-limb coefficients by key
            mode = f64 ? CF_ADDF : (y->cf_limbs == 1 ? CF_ADD1 : (y->cf_limbs == 2 ? CF_ADD2 : (y->cf_limbs == 3 ? CF_ADD3 : CF_ADD5)));
            if (!f64 && y->cf_limbs == 4) {
                e->err = "internal: merge input must carry 1, 2, 3 or 5 limbs";
                return OBK_ERR_INVALID_ARG;
            }
            st.acc_limbs = f64 ? 1 : y->cf_limbs;
            tot_mults = ny;
        }
        st.term_products = tot_mults;
        const int accw = acc_words_of(mode);
        CU_CHECK(e, cudaEventRecord(e->ev[2], s));

        // ---- geometry of the shared-memory segment table ---------------------------------------------
        const unsigned threads = (unsigned)std::max<int64_t>(32, std::min<int64_t>(1024, e->opt_threads)) / 32 * 32;
        const uint32_t log2_cap = pick_log2_cap(e, W, accw);
        const uint32_t cap = 1u << log2_cap;
        const size_t smem_mul = (size_t)cap * (8 * W + 8 * accw + 4);
        if (smem_mul + 1024 > e->smem_optin) {
            e->err = "segment table does not fit in shared memory";
            return OBK_ERR_INVALID_ARG;
        }
        st.table_slots = cap;
        const uint32_t max_fill = (uint32_t)((uint64_t)cap * 9 / 10);
        const uint64_t max_bins = 1ull << 25;
        const uint32_t wave = (uint32_t)e->sm_count * (uint32_t)std::max<int64_t>(1, e->opt_ctas_per_sm); // resident CTAs

        // ---- dense operands: row-blocked path (register accumulation, no table) --------------------------------
        RowsResult rows;
        int dense_kernel = 1;
        if (!rq.merge && (!trunc || tiles_take) && snx > 0 && W == 1 && L.nvars >= 3 && CWx == 1 && (e->opt_kernel == -1 || e->opt_kernel == 3)
            && (mode == CF_I1A1 || mode == CF_I1A2 || mode == CF_I1A3 || mode == CF_F64) && tiles_ok
            && (e->opt_kernel == 3 || tot_mults >= (1ull << 22) || tiles_take)) {
            const bool need_total = rq.push != nullptr || (rq.nranks > 1 && !owner) || rq.result_mem != OBK_MEM_DEVICE || e->opt_regroup == 2;
            if (trunc) {
                tile_dims.trunc = true;
                tile_dims.limit = rq.t->limit;
                tile_dims.var_mask = var_mask_t;
                tile_dims.w0 = (uint32_t)(var_mask_t & 1ull);
                tile_dims.w1 = (uint32_t)((var_mask_t >> 1) & 1ull);
            }
            rows = run_tiles(e, sc, plane_pre, dxk, dxc, nx_full, dyk, dyc, ny, L, mode, accw, e->opt_kernel == 3, !f64 && hs_anyneg == 0,
                             rq.rank, rq.nranks, owner, tile_dims, need_total, st);
            if (rows.done) dense_kernel = 3;
        }
        if (!rows.done && !rq.merge && !trunc && snx > 0 && W == 1 && L.nvars >= 3 && CWx == 1 && (e->opt_kernel == -1 || e->opt_kernel == 2)
            && (mode == CF_I1A1 || mode == CF_I1A2 || mode == CF_I1A3 || mode == CF_F64) && rows_ok && planes_ok
            && (e->opt_kernel == 2 || tot_mults >= (1ull << 22))) {
            // the term count may stay on the device until the end of the call when the product is left in HBM as it is
            const bool need_total = rq.push != nullptr || (rq.nranks > 1 && !owner) || rq.result_mem != OBK_MEM_DEVICE || e->opt_regroup == 2;
            rows = run_planes(e, sc, plane_pre, dxk, dxc, nx_full, dyk, dyc, ny, L, mode, accw, e->opt_kernel == 2, !f64 && hs_anyneg == 0,
                              rq.rank, rq.nranks, owner, plane_rows[0], plane_rows[1], need_total, st);
            if (rows.done) dense_kernel = 2;
        }
        if (!rows.done && e->opt_kernel != 2 && !rq.merge && !trunc && snx > 0 && W == 1 && L.nvars >= 2 && CWx == 1 && e->opt_kernel != 0
            && (mode == CF_I1A1 || mode == CF_I1A2 || mode == CF_I1A3 || mode == CF_F64) && rows_ok
            && (e->opt_kernel == 1 || tot_mults >= (1ull << 22))) {
            rows = run_rows(e, sc, dxk, dxc, nx_full, dyk, dyc, ny, L, mode, accw, nx_full * ny, threads, e->opt_kernel == 1,
                            !f64 && hs_anyneg == 0, rq.rank, rq.nranks, owner, st);
        }

        uint64_t *rk = nullptr, *rc = nullptr;
        uint64_t total = 0;
        uint32_t nsegs = 1;
        float ms_prep_extra = 0;
        if (rows.done) {
            rk = rows.rk;
            rc = rows.rc;
            total = rows.total;
            st.est_nterms = rows.est_rows;
            st.nsegs = rows.m;
            st.nterms_out = total;
            st.kernel = dense_kernel;
            CU_CHECK(e, cudaEventRecord(e->ev[3], s));
        } else {
            // ---- product-size estimate (poly_mul_estimate_product_size, polynomial.hpp:478-794) -------------
            // Sampled symbolic product: count the distinct product keys of a few fine-grained segments exactly.
            uint64_t est = 0;
            if (rq.merge) {
                // distinct keys <= input terms: a segmentation sized for that bound can never overflow
                est = ny;
                nsegs = pick_nsegs(ny, (double)cap * 0.85, ny * 64, wave);
            } else if (e->opt_nsegs > 0) {
                nsegs = (uint32_t)e->opt_nsegs;
            } else {
                EstimateIn ein{sxk, sxdeg, snx, dyk, ydeg, ny, W, L.is_signed, dbits, degmin_y, ndeg, trunc, lim_base,
                               tot_mults, threads};
                ein.ts = &ts;
                est = estimate_terms(e, sc, ein);
                // owner-computes over several GPUs: a wave is the resident CTAs of ALL ranks (a small product cut into one
                // wave of one GPU leaves every rank with 1/nranks of its SMs busy: audi_01 has 139 segments)
                nsegs = pick_nsegs(est, (double)cap * (double)e->opt_fill_pct / 100.0, tot_mults, owner ? wave * rq.nranks : wave);
            }
            st.est_nterms = est;
            CU_CHECK(e, cudaEventRecord(e->ev[3], s));

            // ---- sort the right operand, multiply, retry with a finer segmentation on table overflow --------
            uint32_t *dflags = nullptr;
            for (unsigned attempt = 0;; ++attempt) {
                if (((uint64_t)nsegs << dbits) > max_bins) {
                    e->err = "The homomorphic multiplication of two polynomials needs more segment bins than the maximum "
                             "allowed value (table size overflow)";
                    return OBK_ERR_TABLE_OVERFLOW;
                }
                if ((uint64_t)nsegs * cap >= (1ull << 32)) {
                    // offsets of the staged terms are 32-bit
                    e->err = "The homomorphic multiplication of two polynomials would produce more than 2^32 terms "
                             "(table size overflow)";
                    return OBK_ERR_TABLE_OVERFLOW;
                }
                CU_CHECK(e, cudaEventRecord(e->ev[4], s));
                SortedY ys, xs{};
                uint32_t *xseg = nullptr, *aseg = nullptr, *bseg = nullptr;
                if (rq.merge) {
                    ys = sort_operand(e, sc, dyk, dyc, nullptr, ny, W, accw, nsegs, L.is_signed, false, 0, 0, true);
                } else {
                    ys = sort_operand(e, sc, dyk, dyc, ydeg, ny, W, CWy, nsegs, L.is_signed, false, dbits, degmin_y, true);
                    xseg = sc.alloc<uint32_t>(std::max<uint64_t>(snx, 1));
                    k_segments<<<(unsigned)std::min<uint64_t>((snx + 255) / 256 + 1, 4096), 256, 0, s>>>(sxk, snx, W, nsegs,
                                                                                                       L.is_signed, xseg);
                    e->launches += 1;
                }
                // final term arrays, sized for the worst case that does not overflow a table (max_fill terms per
                // segment): the kernel writes every segment straight to its reserved range, no staging or compaction
                const uint64_t out_bound = std::max<uint64_t>((uint64_t)nsegs * max_fill, 1);
                rk = sc.alloc_out<uint64_t>(out_bound * W);
                rc = sc.alloc_out<uint64_t>(out_bound * accw);
                dflags = sc.alloc<uint32_t>(4);
                CU_CHECK(e, cudaMemsetAsync(dflags, 0, 16, s));
                MulParams P{};
                uint64_t *zero_key = nullptr;
                if (rq.merge) {
                    zero_key = sc.alloc<uint64_t>(W + 1);
                    CU_CHECK(e, cudaMemsetAsync(zero_key, 0, (W + 1) * 8, s));
                    P.xk = zero_key;
                    P.xseg = (const uint32_t *)(zero_key + W);
                    P.xc = nullptr;
                    P.nx = 1;
                } else {
                    P.xk = sxk;
                    P.xc = sxc;
                    P.xdeg = sxdeg;
                    P.xseg = xseg;
                    P.nx = snx;
                    P.npass = 1;
                    if (ts.on) {
                        aseg = sc.alloc<uint32_t>(std::max<uint64_t>(ts.na, 1));
                        bseg = sc.alloc<uint32_t>(std::max<uint64_t>(ts.nb, 1));
                        if (ts.na) k_segments<<<(unsigned)std::min<uint64_t>((ts.na + 255) / 256, 4096), 256, 0, s>>>(ts.ak, ts.na, W, nsegs, L.is_signed, aseg);
                        if (ts.nb) k_segments<<<(unsigned)std::min<uint64_t>((ts.nb + 255) / 256, 4096), 256, 0, s>>>(ts.bk, ts.nb, W, nsegs, L.is_signed, bseg);
                        xs = sort_operand(e, sc, sxk, sxc, sxdeg, snx, W, CWx, nsegs, L.is_signed, false, ts.dbits_x, ts.degmin_x, true);
                        e->launches += 2;
                        P.xk = ts.ak;
                        P.xc = ts.ac;
                        P.xdeg = ts.adeg;
                        P.xseg = aseg;
                        P.nx = ts.na;
                        P.npass = 2;
                        P.b.xk = ts.bk;
                        P.b.xc = ts.bc;
                        P.b.xdeg = ts.bdeg;
                        P.b.xseg = bseg;
                        P.b.nx = ts.nb;
                        P.b.yk = xs.keys;
                        P.b.yc = xs.cfs;
                        P.b.yoff = xs.off;
                        P.b.dbits = ts.dbits_x;
                        P.b.ndeg = ts.ndeg_x;
                        P.b.dlo = ts.dlo_b;
                        P.b.lim_base = ts.lim_base_b;
                    }
                }
                P.yk = ys.keys;
                P.yc = ys.cfs;
                P.yoff = ys.off;
                P.nsegs = nsegs;
                P.dbits = dbits;
                P.ndeg = ndeg;
                P.trunc = trunc;
                P.lim_base = lim_base;
                P.seg_list = nullptr;
                P.nseg_list = nsegs;
                if (owner) {
                    // owner-computes: this rank forms the contiguous range [a, b) of the output segments
                    const uint32_t a = (uint32_t)((uint64_t)nsegs * rq.rank / rq.nranks);
                    const uint32_t b = (uint32_t)((uint64_t)nsegs * (rq.rank + 1) / rq.nranks);
                    P.seg_base = a;
                    P.nseg_list = b - a;
                }
                P.cursor = dflags + 1;
                P.log2_cap = log2_cap;
                P.max_fill = max_fill;
                P.fence = e->opt_fence ? 1u : 0u;
                P.okeys = rk;
                P.ocfs = rc;
                P.counts = nullptr;
                P.flags = dflags;
                P.out_cursor = dflags + 2;
                P.uns = (!f64 && !rq.merge && hs_anyneg == 0) ? 1u : 0u;
                plan_units(e, P, threads,
                           (double)tot_mults / (double)std::max<uint64_t>(P.nx + (P.npass > 1 ? P.b.nx : 0), 1) / (double)nsegs,
                           std::max(snx, ny));
                st.group_lanes = e->last_rslices;
                CU_CHECK(e, cudaEventRecord(e->ev[5], s));
                launch_mul(e, W, mode, P, (unsigned)std::min<uint64_t>(std::max<uint32_t>(P.nseg_list, 1), (uint64_t)e->sm_count * (uint64_t)std::max<int64_t>(1, e->opt_ctas_per_sm)), threads, smem_mul);
                CU_CHECK(e, cudaEventRecord(e->ev[6], s));
                st.mul_launches += 1;
                uint32_t hflag[4];
                CU_CHECK(e, cudaMemcpyAsync(hflag, dflags, 16, cudaMemcpyDeviceToHost, s));
                CU_CHECK(e, cudaStreamSynchronize(s));
                const uint32_t htotal = hflag[2];
                float ms = 0;
                cudaEventElapsedTime(&ms, e->ev[4], e->ev[5]);
                ms_prep_extra += ms;
                if (zero_key) sc.release(zero_key);
                if (xseg) sc.release(xseg);
                if (xs.keys) {
                    sc.release(xs.keys);
                    sc.release(xs.cfs);
                    sc.release(xs.off);
                    sc.release(aseg);
                    sc.release(bseg);
                }
                sc.release(ys.keys);
                sc.release(ys.cfs);
                sc.release(ys.off);
                if (hflag[0] == 0) {
                    total = htotal;
                    break;
                }
                // some segment's table filled up: finer segmentation
                sc.release(rk);
                sc.release(rc);
                sc.release(dflags);
                st.retries += 1;
                if ((int64_t)st.retries > e->opt_max_retries || owner) {
                    e->err = "The homomorphic multiplication of two polynomials resulted in a segment table whose size "
                             "is larger than the maximum allowed value (shared-memory table overflow after "
                             + std::to_string(st.retries) + " finer segmentations)";
                    return OBK_ERR_TABLE_OVERFLOW;
                }
                nsegs = prev_prime(((uint64_t)nsegs * 2 + wave - 1) / wave * wave);
                if (nsegs < 2) nsegs = 2;
            }
            st.nsegs = nsegs;
            st.nterms_out = total;

            sc.release(dflags);
        }

        if (rq.push) {
            const obk_peer_windows &pw = *rq.push;
            PushParams PP{};
            PP.keys = rk;
            PP.cfs = rc;
            PP.n = total;
            PP.W = W;
            PP.L = accw;
            PP.is_signed = L.is_signed;
            PP.mod = OBK_EXCHANGE_MOD;
            PP.nranks = pw.nranks;
            PP.cap = (pw.window_bytes - PUSH_HEADER) / ((uint64_t)8 * (W + accw));
            for (uint32_t g = 0; g < pw.nranks; ++g) {
                PP.win[g] = static_cast<unsigned char *>(pw.base[g]) + (size_t)pw.parity * pw.window_bytes;
                PP.first_group[g] = (uint32_t)((uint64_t)OBK_EXCHANGE_MOD * g / pw.nranks);
            }
            PP.first_group[pw.nranks] = OBK_EXCHANGE_MOD;
            uint32_t *dover = sc.alloc<uint32_t>(4);
            CU_CHECK(e, cudaMemsetAsync(dover, 0, 16, s));
            PP.local_overflow = dover;
            if (total) {
                const unsigned g = (unsigned)std::min<uint64_t>((total + 256 * PUSH_ITEMS - 1) / (256 * PUSH_ITEMS), 4 * (uint64_t)e->sm_count);
                const size_t psm = (size_t)256 * PUSH_ITEMS * 8 * (W + accw);
                CU_CHECK(e, cudaFuncSetAttribute(k_push, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm));
                k_push<<<g, 256, psm, s>>>(PP);
                e->launches += 1;
                CU_CHECK(e, cudaGetLastError());
            }
            uint32_t hover = 0;
            CU_CHECK(e, cudaMemcpyAsync(&hover, dover, 4, cudaMemcpyDeviceToHost, s));
            CU_CHECK(e, cudaEventRecord(e->ev[7], s));
            CU_CHECK(e, cudaEventRecord(e->ev[8], s));
            CU_CHECK(e, cudaStreamSynchronize(s));
            if (hover) {
                e->err = "peer window too small for the pushed partial terms (window_bytes = " + std::to_string(pw.window_bytes) + ")";
                return OBK_ERR_TABLE_OVERFLOW;
            }
            set_empty_result(out, x, accw, OBK_MEM_DEVICE);
            st.nterms_out = total;
            float msx;
            cudaEventElapsedTime(&msx, e->ev[0], e->ev[8]);
            st.ms_total = msx;
            cudaEventElapsedTime(&msx, e->ev[5], e->ev[6]);
            st.ms_mul = msx;
            st.total_launches = e->launches;
            if (stats) *stats = st;
            return OBK_OK;
        }
        TailIn tin{rk, rc, total, W, accw, f64, L.is_signed, tot_mults, rq.result_mem, rq.nranks, owner, 1};
        tin.total_dev = rows.done ? rows.total_dev : nullptr;
        emit_result(e, sc, tin, out, st);
        total = tin.total;
        ResultOwner *own = static_cast<ResultOwner *>(out->owner);
        if (rq.nranks > 1 && !owner) nsegs = tin.nsegs_out;
        if (owner) {
            // this rank's share of the whole job's term products (the shares of all ranks add up to tot_n_mults)
            st.term_products = tot_mults * (rq.rank + 1) / rq.nranks - tot_mults * rq.rank / rq.nranks;
            if (rq.send_counts) {
                std::memset(rq.send_counts, 0, sizeof(uint64_t) * rq.nranks);
                rq.send_counts[rq.rank] = total;
            }
        } else if (rq.send_counts) {
            for (uint32_t g = 0; g < rq.nranks; ++g) {
                const uint64_t a = (uint64_t)nsegs * g / rq.nranks, b = (uint64_t)nsegs * (g + 1) / rq.nranks;
                rq.send_counts[g] = own->seg_offsets[b] - own->seg_offsets[a];
            }
        }
        float ms;
        cudaEventElapsedTime(&ms, e->ev[0], e->ev[1]);
        st.ms_h2d = ms;
        cudaEventElapsedTime(&ms, e->ev[1], e->ev[2]);
        st.ms_prep = ms + ms_prep_extra;
        cudaEventElapsedTime(&ms, e->ev[2], e->ev[3]);
        st.ms_estimate = ms;
        cudaEventElapsedTime(&ms, e->ev[5], e->ev[6]);
        st.ms_mul = ms;
        cudaEventElapsedTime(&ms, e->ev[6], e->ev[7]);
        st.ms_compact = ms;
        cudaEventElapsedTime(&ms, e->ev[7], e->ev[8]);
        st.ms_d2h = ms;
        cudaEventElapsedTime(&ms, e->ev[0], e->ev[8]);
        st.ms_total = ms;
        st.total_launches = e->launches;
        if (stats) *stats = st;
        return OBK_OK;
    } catch (const obk_fail &f) {
        // strong-ish guarantee of the reference: the result is left empty (polynomial.hpp:1575-1580)
        if (out->owner) {
            obk_result_free(e, out);
        }
        std::memset(out, 0, sizeof(*out));
        cudaStreamSynchronize(e->stream);
        return f.st;
    }
}

} // namespace

// A multi-device engine forwards everything but the product to its first device (the O(terms) steps either side of the
// product do not need more than one GPU) and reports that device's error text.
#define OBK_FORWARD_MULTI(e, call)                                                                                     \
    do {                                                                                                               \
        if (!(e)->subs.empty()) {                                                                                      \
            const obk_status _st = (call);                                                                             \
            if (_st != OBK_OK) (e)->err = (e)->subs[0]->err;                                                           \
            return _st;                                                                                                \
        }                                                                                                              \
    } while (0)
#define OBK_REJECT_MULTI(e)                                                                                            \
    do {                                                                                                               \
        if (!(e)->subs.empty()) {                                                                                      \
            (e)->err = "not available on a multi-device engine (it has no stream of its own): use a single-device engine"; \
            return OBK_ERR_UNSUPPORTED;                                                                                \
        }                                                                                                              \
    } while (0)

extern "C" {

const char *obk_version(void)
{
    return OBK_VERSION_STRING;
}

const char *obk_status_string(obk_status s)
{
    switch (s) {
        case OBK_OK: return "ok";
        case OBK_ERR_INVALID_ARG: return "invalid argument";
        case OBK_ERR_KEY_OVERFLOW: return "monomial exponent overflow";
        case OBK_ERR_TABLE_OVERFLOW: return "segment table overflow";
        case OBK_ERR_CF_OVERFLOW: return "coefficient overflow beyond 3 limbs";
        case OBK_ERR_CUDA: return "CUDA error";
        case OBK_ERR_UNSUPPORTED: return "unsupported";
        case OBK_ERR_NOMEM: return "out of memory";
    }
    return "unknown";
}

static thread_local std::string g_create_err;

obk_status obk_engine_create(int device, obk_engine **out)
{
    if (!out) return OBK_ERR_INVALID_ARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t err = cudaGetDeviceCount(&ndev);
    if (err != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
        cudaGetLastError();
        g_create_err = "no usable CUDA device (the engine has no CPU fallback)";
        return OBK_ERR_CUDA;
    }
    auto *e = new obk_engine();
    e->device = device;
    if (cudaSetDevice(device) != cudaSuccess) {
        delete e;
        return OBK_ERR_CUDA;
    }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    e->sm_count = prop.multiProcessorCount;
    e->smem_optin = prop.sharedMemPerBlockOptin;
    if (prop.major < 10) {
        // built for sm_100a only: refuse anything else loudly
        g_create_err = "device is not a Blackwell (sm_100) GPU";
        delete e;
        return OBK_ERR_CUDA;
    }
    if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete e;
        return OBK_ERR_CUDA;
    }
    for (auto &ev : e->ev) cudaEventCreate(&ev);
    // keep freed scratch in the driver's stream-ordered pool (grow-only workspace)
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    *out = e;
    return OBK_OK;
}

obk_status obk_engine_create_multi(const int *devices, int ndev, obk_engine **out)
{
    if (!out || !devices || ndev < 1 || ndev > 16) return OBK_ERR_INVALID_ARG;
    *out = nullptr;
    if (ndev == 1) return obk_engine_create(devices[0], out);
    for (int i = 0; i < ndev; ++i)
        for (int j = 0; j < i; ++j)
            if (devices[i] == devices[j]) {
                g_create_err = "obk_engine_create_multi: a device is listed twice";
                return OBK_ERR_INVALID_ARG;
            }
    auto *m = new obk_engine();
    m->device = devices[0];
    m->own_stream = false;
    for (int i = 0; i < ndev; ++i) {
        obk_engine *sub = nullptr;
        const obk_status st = obk_engine_create(devices[i], &sub);
        if (st != OBK_OK) {
            for (auto *q : m->subs) obk_engine_destroy(q);
            delete m;
            return st;
        }
        m->subs.push_back(sub);
    }
    m->sm_count = m->subs[0]->sm_count;
    m->smem_optin = m->subs[0]->smem_optin;
    *out = m;
    return OBK_OK;
}

void obk_engine_destroy(obk_engine *e)
{
    if (!e) return;
    if (!e->subs.empty()) {
        for (auto *q : e->subs) obk_engine_destroy(q);
        delete e;
        return;
    }
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    for (auto &b : e->pinned)
        if (b.ptr) cudaFreeHost(b.ptr);
    for (auto &ev : e->ev) cudaEventDestroy(ev);
    for (auto &c : e->arena) cudaFree(c.base);
    if (e->own_stream) cudaStreamDestroy(e->stream);
    delete e;
}

obk_status obk_engine_set_stream(obk_engine *e, void *cuda_stream)
{
    if (!e) return OBK_ERR_INVALID_ARG;
    OBK_REJECT_MULTI(e);
    if (!e) return OBK_ERR_INVALID_ARG;
    cudaStreamSynchronize(e->stream);
    if (e->own_stream) cudaStreamDestroy(e->stream);
    e->stream = (cudaStream_t)cuda_stream;
    e->own_stream = false;
    return OBK_OK;
}

obk_status obk_engine_set_option(obk_engine *e, const char *name, int64_t value)
{
    if (!e || !name) return OBK_ERR_INVALID_ARG;
    for (auto *q : e->subs) {
        const obk_status st = obk_engine_set_option(q, name, value);
        if (st != OBK_OK) {
            e->err = q->err;
            return st;
        }
    }
    const std::string n(name);
    if (n == "nsegs") e->opt_nsegs = value;
    else if (n == "fence") e->opt_fence = value;
    else if (n == "regroup") e->opt_regroup = value;
    else if (n == "row_threads") e->opt_row_threads = value;
    else if (n == "two_sided") e->opt_two_sided = value;
    else if (n == "ctas_per_sm") e->opt_ctas_per_sm = value;
    else if (n == "refill_min") e->opt_refill_min = value;
    else if (n == "mgpu_owner") e->opt_mgpu_owner = value;
    else if (n == "table_slots") e->opt_table_slots = value;
    else if (n == "threads") e->opt_threads = value;
    else if (n == "kernel") e->opt_kernel = value;
    else if (n == "plane_cfg") e->opt_plane_cfg = value;
    else if (n == "tile_split") e->opt_tile_split = value;
    else if (n == "tile_units") e->opt_tile_units = value;
    else if (n == "max_retries") e->opt_max_retries = value;
    else if (n == "fill_pct") e->opt_fill_pct = value;
    else {
        e->err = "unknown option " + n;
        return OBK_ERR_INVALID_ARG;
    }
    return OBK_OK;
}

const char *obk_last_error(const obk_engine *e)
{
    return e ? e->err.c_str() : g_create_err.c_str();
}

static obk_status multi_mul(obk_engine *m, const obk_poly_view *x, const obk_poly_view *y, const obk_trunc *t, uint32_t result_mem,
                            obk_poly_result *out, obk_stats *stats);

obk_status obk_poly_mul(obk_engine *e, const obk_poly_view *x, const obk_poly_view *y, const obk_trunc *t,
                        uint32_t result_mem, obk_poly_result *out, obk_stats *stats)
{
    if (!e) return OBK_ERR_INVALID_ARG;
    if (!e->subs.empty()) return multi_mul(e, x, y, t, result_mem, out, stats);
    MulRequest rq{x, y, t, result_mem};
    return run_mul(e, rq, out, stats);
}

obk_status obk_poly_mul_partial(obk_engine *e, const obk_poly_view *x, const obk_poly_view *y, const obk_trunc *t,
                                uint32_t rank, uint32_t nranks, obk_poly_result *out_partial, uint64_t *send_counts,
                                obk_stats *stats)
{    if (!e) return OBK_ERR_INVALID_ARG;
    OBK_REJECT_MULTI(e);

    if (!e || nranks == 0 || rank >= nranks) return OBK_ERR_INVALID_ARG;
    MulRequest rq{x, y, t, OBK_MEM_DEVICE};
    rq.rank = rank;
    rq.nranks = nranks;
    rq.send_counts = send_counts;
    return run_mul(e, rq, out_partial, stats);
}

obk_status obk_poly_merge(obk_engine *e, const obk_poly_view *terms, uint32_t nsegs, uint32_t result_mem,
                          obk_poly_result *out, obk_stats *stats)
{    if (!e) return OBK_ERR_INVALID_ARG;
    OBK_REJECT_MULTI(e);

    if (!e || !terms) return OBK_ERR_INVALID_ARG;
    MulRequest rq{nullptr, terms, nullptr, result_mem};
    rq.merge = true;
    rq.merge_nsegs = nsegs ? nsegs : 1;
    return run_mul(e, rq, out, stats);
}

obk_status obk_peer_window_alloc(obk_engine *e, uint64_t window_bytes, void **base, void *ipc_handle64)
{    if (!e) return OBK_ERR_INVALID_ARG;
    OBK_REJECT_MULTI(e);

    if (!e || !base || window_bytes < 4096) return OBK_ERR_INVALID_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    try {
        CU_CHECK(e, cudaSetDevice(e->device));
        void *p = nullptr;
        CU_CHECK(e, cudaMalloc(&p, 2 * window_bytes));
        CU_CHECK(e, cudaMemset(p, 0, PUSH_HEADER));
        CU_CHECK(e, cudaMemset(static_cast<char *>(p) + window_bytes, 0, PUSH_HEADER));
        if (ipc_handle64) {
            cudaIpcMemHandle_t h;
            cudaError_t err = cudaIpcGetMemHandle(&h, p);
            if (err != cudaSuccess) {
                cudaGetLastError();
                cudaFree(p);
                e->err = std::string("cudaIpcGetMemHandle failed: ") + cudaGetErrorString(err);
                return OBK_ERR_CUDA;
            }
            std::memcpy(ipc_handle64, &h, 64);
        }
        *base = p;
        return OBK_OK;
    } catch (const obk_fail &f) {
        return f.st;
    }
}

obk_status obk_peer_window_open(obk_engine *e, const void *ipc_handle64, void **base)
{    if (!e) return OBK_ERR_INVALID_ARG;
    OBK_REJECT_MULTI(e);

    if (!e || !ipc_handle64 || !base) return OBK_ERR_INVALID_ARG;
    cudaSetDevice(e->device);
    cudaIpcMemHandle_t h;
    std::memcpy(&h, ipc_handle64, 64);
    void *p = nullptr;
    cudaError_t err = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (err != cudaSuccess) {
        cudaGetLastError();
        e->err = std::string("cudaIpcOpenMemHandle failed: ") + cudaGetErrorString(err);
        return OBK_ERR_CUDA;
    }
    *base = p;
    return OBK_OK;
}

obk_status obk_peer_window_close(obk_engine *e, void *base, int own)
{    if (!e) return OBK_ERR_INVALID_ARG;
    OBK_REJECT_MULTI(e);

    if (!e || !base) return OBK_ERR_INVALID_ARG;
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    cudaError_t err = own ? cudaFree(base) : cudaIpcCloseMemHandle(base);
    if (err != cudaSuccess) {
        cudaGetLastError();
        return OBK_ERR_CUDA;
    }
    return OBK_OK;
}

static bool check_windows(obk_engine *e, const obk_peer_windows *w)
{
    if (!w || !w->base || w->nranks == 0 || w->nranks > (uint32_t)PUSH_MAX_RANKS || w->rank >= w->nranks || w->parity > 1
        || w->window_bytes < 4096) {
        e->err = "invalid peer windows (1..16 ranks, parity 0|1, window_bytes >= 4096)";
        return false;
    }
    for (uint32_t g = 0; g < w->nranks; ++g)
        if (!w->base[g]) {
            e->err = "invalid peer windows: null base pointer";
            return false;
        }
    return true;
}

obk_status obk_peer_window_reset(obk_engine *e, const obk_peer_windows *w)
{    if (!e) return OBK_ERR_INVALID_ARG;
    OBK_REJECT_MULTI(e);

    if (!e) return OBK_ERR_INVALID_ARG;
    if (!check_windows(e, w)) return OBK_ERR_INVALID_ARG;
    try {
        CU_CHECK(e, cudaSetDevice(e->device));
        unsigned char *win = static_cast<unsigned char *>(w->base[w->rank]) + (size_t)w->parity * w->window_bytes;
        CU_CHECK(e, cudaMemsetAsync(win, 0, 8, e->stream));
        CU_CHECK(e, cudaStreamSynchronize(e->stream));
        return OBK_OK;
    } catch (const obk_fail &f) {
        return f.st;
    }
}

obk_status obk_poly_mul_push(obk_engine *e, const obk_poly_view *x, const obk_poly_view *y, const obk_trunc *t,
                             const obk_peer_windows *w, obk_stats *stats)
{    if (!e) return OBK_ERR_INVALID_ARG;
    OBK_REJECT_MULTI(e);

    if (!e) return OBK_ERR_INVALID_ARG;
    if (!check_windows(e, w)) return OBK_ERR_INVALID_ARG;
    MulRequest rq{x, y, t, OBK_MEM_DEVICE};
    rq.rank = w->rank;
    rq.nranks = w->nranks;
    rq.push = w;
    obk_poly_result dummy{};
    const obk_status st = run_mul(e, rq, &dummy, stats);
    if (dummy.owner) obk_result_free(e, &dummy);
    return st;
}

obk_status obk_peer_merge(obk_engine *e, const obk_peer_windows *w, const obk_poly_view *like, uint32_t acc_limbs,
                          uint32_t result_mem, obk_poly_result *out, obk_stats *stats)
{    if (!e) return OBK_ERR_INVALID_ARG;
    OBK_REJECT_MULTI(e);

    if (!e || !like || !out) return OBK_ERR_INVALID_ARG;
    if (!check_windows(e, w)) return OBK_ERR_INVALID_ARG;
    std::memset(out, 0, sizeof(*out));
    try {
        CU_CHECK(e, cudaSetDevice(e->device));
        unsigned char *win = static_cast<unsigned char *>(w->base[w->rank]) + (size_t)w->parity * w->window_bytes;
        uint32_t hdr[2] = {0, 0};
        CU_CHECK(e, cudaMemcpyAsync(hdr, win, 8, cudaMemcpyDeviceToHost, e->stream));
        CU_CHECK(e, cudaStreamSynchronize(e->stream));
        const bool f64 = like->cf_kind == OBK_CF_F64;
        const uint32_t L = f64 ? 1u : acc_limbs;
        if (L < 1 || L > (uint32_t)MAX_CF_LIMBS || L == 4) {
            e->err = "acc_limbs must be 1, 2, 3 or 5";
            return OBK_ERR_INVALID_ARG;
        }
        const uint64_t cap = (w->window_bytes - PUSH_HEADER) / ((uint64_t)8 * (like->key_words + L));
        if (hdr[1] != 0 || hdr[0] > cap) {
            CU_CHECK(e, cudaMemsetAsync(win, 0, 8, e->stream));
            e->err = "peer window overflow: " + std::to_string(hdr[0]) + " terms were pushed into a window of " + std::to_string(cap);
            return OBK_ERR_TABLE_OVERFLOW;
        }
        obk_poly_view terms = *like;
        terms.nterms = hdr[0];
        terms.cf_limbs = L;
        terms.mem = OBK_MEM_DEVICE;
        terms.keys = reinterpret_cast<const uint64_t *>(win + PUSH_HEADER);
        terms.cfs = terms.keys + cap * like->key_words;
        MulRequest rq{nullptr, &terms, nullptr, result_mem};
        rq.merge = true;
        const obk_status st = run_mul(e, rq, out, stats);
        CU_CHECK(e, cudaMemsetAsync(win, 0, 8, e->stream));
        return st;
    } catch (const obk_fail &f) {
        return f.st;
    }
}

obk_status obk_poly_addsub(obk_engine *e, const obk_poly_view *x, const obk_poly_view *y, int subtract, uint32_t result_mem,
                           obk_poly_result *out, obk_stats *stats)
{    if (!e) return OBK_ERR_INVALID_ARG;
    OBK_FORWARD_MULTI(e, obk_poly_addsub(e->subs[0], x, y, subtract, result_mem, out, stats));

    if (!e || !out) return OBK_ERR_INVALID_ARG;
    std::memset(out, 0, sizeof(*out));
    uint64_t *cat_k = nullptr, *cat_c = nullptr;
    obk_status rc_status = OBK_OK;
    try {
        CU_CHECK(e, cudaSetDevice(e->device));
        validate_view(e, x, "x");
        validate_view(e, y, "y");
        if (x->key_words != y->key_words || x->key_signed != y->key_signed || x->nvars != y->nvars || x->psize != y->psize
            || x->cf_kind != y->cf_kind || (x->key_bits == 32) != (y->key_bits == 32)) {
            e->err = "operands have different key layouts or coefficient kinds";
            return OBK_ERR_INVALID_ARG;
        }
        const bool f64 = x->cf_kind == OBK_CF_F64;
        const int W = (int)x->key_words;
        const uint64_t nx = x->nterms, ny = y->nterms, n = nx + ny;
        if (n == 0) {
            set_empty_result(out, x, 1, result_mem);
            if (stats) *stats = obk_stats{};
            return OBK_OK;
        }
        if (n >= (1ull << 31)) {
            e->err = "operands with 2^31 or more terms are not supported";
            return OBK_ERR_UNSUPPORTED;
        }
        cudaStream_t s = e->stream;
        int lout = 1;
        {
            Scratch sc(e);
            const obk_poly_view *vs[2] = {x, y};
            const uint64_t *dk[2] = {nullptr, nullptr}, *dc[2] = {nullptr, nullptr};
            for (int i = 0; i < 2; ++i) {
                const obk_poly_view *v = vs[i];
                const int cw = f64 ? 1 : (int)v->cf_limbs;
                if (v->nterms == 0) continue;
                if (v->mem == OBK_MEM_DEVICE) {
                    dk[i] = v->keys;
                    dc[i] = (const uint64_t *)v->cfs;
                } else {
                    uint64_t *k = sc.alloc<uint64_t>(v->nterms * W);
                    uint64_t *c = sc.alloc<uint64_t>(v->nterms * cw);
                    CU_CHECK(e, cudaMemcpyAsync(k, v->keys, v->nterms * W * 8, cudaMemcpyHostToDevice, s));
                    CU_CHECK(e, cudaMemcpyAsync(c, v->cfs, v->nterms * cw * 8, cudaMemcpyHostToDevice, s));
                    dk[i] = k;
                    dc[i] = c;
                }
            }
            if (!f64) {
                // exact sum: one more bit than the wider operand, plus the sign
                const Layout L = make_layout(*x);
                StatsOut *dstats = sc.alloc<StatsOut>(2);
                k_stats_init<<<2, 64, 0, s>>>(dstats, 2);
                for (int i = 0; i < 2; ++i) {
                    if (vs[i]->nterms == 0) continue;
                    const unsigned g = (unsigned)std::min<uint64_t>((vs[i]->nterms + 255) / 256, 1024);
                    k_term_stats<<<g, 256, 0, s>>>(dk[i], dc[i], vs[i]->nterms, L, 0, 0, (int)vs[i]->cf_limbs, nullptr, dstats + i);
                }
                StatsOut hs[2];
                CU_CHECK(e, cudaMemcpyAsync(hs, dstats, sizeof(hs), cudaMemcpyDeviceToHost, s));
                CU_CHECK(e, cudaStreamSynchronize(s));
                const unsigned bits = std::max(hs[0].cfbits, hs[1].cfbits) + 2;
                lout = (int)((bits + 63) / 64);
                if (lout == 4) lout = 5; // the merge kernel carries 1, 2, 3 or 5 limbs
                if (lout > MAX_CF_LIMBS) {
                    e->err = "coefficient overflow beyond 5 limbs: the exact sum needs " + std::to_string(bits) + " bits";
                    return OBK_ERR_CF_OVERFLOW;
                }
            }
            // merge input: x's terms followed by (+/-) y's terms
            if (cudaMallocAsync((void **)&cat_k, n * W * 8, s) != cudaSuccess
                || cudaMallocAsync((void **)&cat_c, n * (size_t)lout * 8, s) != cudaSuccess) {
                cudaGetLastError();
                e->err = "device allocation failed";
                throw obk_fail{OBK_ERR_NOMEM};
            }
            uint64_t off = 0;
            for (int i = 0; i < 2; ++i) {
                const uint64_t ni = vs[i]->nterms;
                if (ni == 0) continue;
                CU_CHECK(e, cudaMemcpyAsync(cat_k + off * W, dk[i], ni * W * 8, cudaMemcpyDeviceToDevice, s));
                const unsigned g = (unsigned)std::min<uint64_t>((ni + 255) / 256, 4096);
                k_widen_cfs<<<g, 256, 0, s>>>(dc[i], ni, f64 ? 1 : (int)vs[i]->cf_limbs, lout, (i == 1 && subtract) ? 1 : 0, f64 ? 1 : 0,
                                             cat_c + off * lout);
                off += ni;
            }
            CU_CHECK(e, cudaGetLastError());
        }
        obk_poly_view terms = *x;
        terms.nterms = n;
        terms.cf_limbs = f64 ? 1 : (uint32_t)lout;
        terms.mem = OBK_MEM_DEVICE;
        terms.keys = cat_k;
        terms.cfs = cat_c;
        MulRequest rq{nullptr, &terms, nullptr, result_mem};
        rq.merge = true;
        rc_status = run_mul(e, rq, out, stats);
    } catch (const obk_fail &f) {
        rc_status = f.st;
    }
    if (cat_k) cudaFreeAsync(cat_k, e->stream);
    if (cat_c) cudaFreeAsync(cat_c, e->stream);
    return rc_status;
}

static obk_status copy_impl(obk_engine *e, const obk_poly_view *x, uint32_t result_mem, int force_log2, obk_poly_result *out)
{
    std::memset(out, 0, sizeof(*out));
    obk_stats st{};
    e->launches = 0;
    try {
        CU_CHECK(e, cudaSetDevice(e->device));
        validate_view(e, x, "x");
        const bool f64 = x->cf_kind == OBK_CF_F64;
        const int W = (int)x->key_words, cw = f64 ? 1 : (int)x->cf_limbs;
        const uint64_t n = x->nterms;
        if (n == 0) {
            set_empty_result(out, x, cw, result_mem);
            return OBK_OK;
        }
        const Layout L = make_layout(*x);
        Scratch sc(e);
        cudaStream_t s = e->stream;
        CU_CHECK(e, cudaEventRecord(e->ev[0], s));
        uint64_t *rk = sc.alloc_out<uint64_t>(n * W);
        uint64_t *rc = sc.alloc_out<uint64_t>(n * cw);
        const cudaMemcpyKind kind = x->mem == OBK_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
        CU_CHECK(e, cudaMemcpyAsync(rk, x->keys, n * W * 8, kind, s));
        CU_CHECK(e, cudaMemcpyAsync(rc, x->cfs, n * cw * 8, kind, s));
        TailIn tin{rk, rc, n, W, cw, f64, L.is_signed, n, result_mem, 1, false, 1};
        tin.force_log2 = force_log2;
        emit_result(e, sc, tin, out, st);
        return OBK_OK;
    } catch (const obk_fail &f) {
        if (out->owner) obk_result_free(e, out);
        std::memset(out, 0, sizeof(*out));
        cudaStreamSynchronize(e->stream);
        return f.st;
    }
}

// ---- single-process multi-device product -------------------------------------------------------------------------
// obake is a single-process library: `f * g` (series.hpp:3186-3194) must be able to use every GPU of the box without
// a launcher.  The multi-device engine runs the zero-exchange design of SURVEY 8(e) from host threads: every device
// gets both operands, forms only the output segments / output planes it owns (ownership is a function of the keys, so
// the devices agree without talking), regroups its share into the SAME 2^k series tables and copies it to the host,
// where the shares of a table are concatenated.  No term crosses NVLink; the only host-side work is one memcpy per
// (table, device).
static obk_status multi_mul(obk_engine *m, const obk_poly_view *x, const obk_poly_view *y, const obk_trunc *t, uint32_t result_mem,
                            obk_poly_result *out, obk_stats *stats)
{
    if (!x || !y || !out) return OBK_ERR_INVALID_ARG;
    std::memset(out, 0, sizeof(*out));
    if (result_mem != OBK_MEM_HOST || x->mem != OBK_MEM_HOST || y->mem != OBK_MEM_HOST) {
        m->err = "the multi-device engine takes HOST operands and returns a HOST result (device-resident chains belong to one "
                 "device: use a single-device engine)";
        return OBK_ERR_UNSUPPORTED;
    }
    const uint32_t n = (uint32_t)m->subs.size();
    std::vector<obk_poly_result> share(n), host(n);
    std::vector<obk_stats> sst(n);
    std::vector<obk_status> rc(n, OBK_OK);
    auto free_all = [&](std::vector<obk_poly_result> &v) {
        for (uint32_t d = 0; d < n; ++d)
            if (v[d].owner) obk_result_free(m->subs[d], &v[d]);
    };
    auto t0 = std::chrono::steady_clock::now();
    int64_t fill = m->opt_fill_pct;
    for (int attempt = 0;; ++attempt) {
        std::vector<std::thread> th;
        for (uint32_t d = 0; d < n; ++d) {
            th.emplace_back([&, d]() {
                obk_engine *e = m->subs[d];
                e->opt_mgpu_owner = 1;
                e->opt_fill_pct = fill;
                MulRequest rq{x, y, t, OBK_MEM_DEVICE};
                rq.rank = d;
                rq.nranks = n;
                rc[d] = run_mul(e, rq, &share[d], &sst[d]);
                e->opt_mgpu_owner = 0;
                e->opt_fill_pct = m->opt_fill_pct;
            });
        }
        for (auto &tt : th) tt.join();
        bool overflow = false, failed = false;
        for (uint32_t d = 0; d < n; ++d) {
            if (rc[d] == OBK_ERR_TABLE_OVERFLOW && m->subs[d]->err.find("segment table") != std::string::npos) overflow = true;
            else if (rc[d] != OBK_OK) failed = true;
        }
        if (!overflow && !failed) break;
        free_all(share);
        if (failed || attempt >= 5) {
            for (uint32_t d = 0; d < n; ++d)
                if (rc[d] != OBK_OK) {
                    m->err = m->subs[d]->err;
                    return rc[d];
                }
        }
        // the devices must agree on the segmentation (ownership is defined by it): all of them retry finer
        fill = std::max<int64_t>(4, fill / 2);
    }
    uint64_t total = 0, tot_mults = 0;
    for (uint32_t d = 0; d < n; ++d) {
        total += share[d].nterms;
        tot_mults += sst[d].term_products;
    }
    const int W = (int)x->key_words;
    const bool f64 = x->cf_kind == OBK_CF_F64;
    uint32_t cw = 1;
    for (uint32_t d = 0; d < n; ++d) cw = std::max(cw, share[d].cf_limbs);
    // the reference's sizing rule for the number of series tables (polynomial.hpp:870-902), on the WHOLE result
    uint32_t k = 0;
    {
        const double term_bytes = 8.0 * W + 8.0 * cw;
        const double est_sp = tot_mults ? (double)total / (double)tot_mults : 1.0;
        const double seg_kb = est_sp >= 1e-3 ? 200.0 : 20.0;
        const uint64_t est_nsegs = (uint64_t)((double)total * term_bytes / (seg_kb * 1024.0));
        while ((est_nsegs >> k) != 0 && k < 22) ++k;
    }
    {
        std::vector<std::thread> th;
        for (uint32_t d = 0; d < n; ++d) {
            th.emplace_back([&, d]() {
                obk_poly_view v = *x;
                v.nterms = share[d].nterms;
                v.cf_limbs = f64 ? 1 : share[d].cf_limbs;
                v.mem = OBK_MEM_DEVICE;
                v.keys = share[d].keys;
                v.cfs = share[d].cfs;
                rc[d] = copy_impl(m->subs[d], &v, OBK_MEM_HOST, (int)k, &host[d]);
            });
        }
        for (auto &tt : th) tt.join();
    }
    free_all(share);
    for (uint32_t d = 0; d < n; ++d)
        if (rc[d] != OBK_OK) {
            m->err = m->subs[d]->err;
            free_all(host);
            return rc[d];
        }
    // ---- concatenate the shares table by table ----------------------------------------------------------------
    const size_t groups = (size_t)1 << k;
    auto *own = new ResultOwner{nullptr};
    own->seg_offsets = (uint64_t *)std::calloc(groups + 1, sizeof(uint64_t));
    own->hk = std::malloc(std::max<uint64_t>(total, 1) * W * 8);
    own->hc = std::malloc(std::max<uint64_t>(total, 1) * cw * 8);
    auto seg_lo = [&](uint32_t d, size_t g) -> uint64_t {
        const obk_poly_result &h = host[d];
        if (h.nterms == 0) return 0;
        return h.nsegs == groups ? h.seg_offsets[g] : (g == 0 ? 0 : h.nterms); // k == 0: one group
    };
    for (size_t g = 0; g < groups; ++g) {
        uint64_t c = 0;
        for (uint32_t d = 0; d < n; ++d) c += seg_lo(d, g + 1) - seg_lo(d, g);
        own->seg_offsets[g + 1] = own->seg_offsets[g] + c;
    }
    {
        std::vector<std::thread> th;
        for (uint32_t d = 0; d < n; ++d) {
            th.emplace_back([&, d]() {
                if (host[d].nterms == 0) return;
                const uint32_t hl = host[d].cf_limbs;
                for (size_t g = 0; g < groups; ++g) {
                    const uint64_t a = seg_lo(d, g), b = seg_lo(d, g + 1);
                    if (a == b) continue;
                    uint64_t dst = own->seg_offsets[g];
                    for (uint32_t q = 0; q < d; ++q) dst += seg_lo(q, g + 1) - seg_lo(q, g);
                    std::memcpy((uint64_t *)own->hk + dst * W, host[d].keys + a * W, (b - a) * W * 8);
                    if (hl == cw) {
                        std::memcpy((uint64_t *)own->hc + dst * cw, (const uint64_t *)host[d].cfs + a * cw, (b - a) * cw * 8);
                    } else { // a share whose accumulators were narrower: sign-extend
                        for (uint64_t i = a; i < b; ++i)
                            for (uint32_t l = 0; l < cw; ++l)
                                ((uint64_t *)own->hc)[(dst + i - a) * cw + l] =
                                    l < hl ? ((const uint64_t *)host[d].cfs)[i * hl + l]
                                           : (uint64_t)((int64_t)((const uint64_t *)host[d].cfs)[i * hl + hl - 1] >> 63);
                    }
                }
            });
        }
        for (auto &tt : th) tt.join();
    }
    obk_stats st{};
    st.nterms_out = total;
    st.log2_nsegs = k;
    for (uint32_t d = 0; d < n; ++d) {
        st.term_products += sst[d].term_products;
        st.h2d_bytes += sst[d].h2d_bytes;
        st.d2h_bytes += (uint64_t)host[d].nterms * (W + host[d].cf_limbs) * 8;
        st.ms_mul = std::max(st.ms_mul, sst[d].ms_mul);
        st.total_launches += sst[d].total_launches;
        st.mul_launches += sst[d].mul_launches;
        st.retries = std::max(st.retries, sst[d].retries);
    }
    st.est_nterms = sst[0].est_nterms;
    st.nsegs = sst[0].nsegs;
    st.acc_limbs = cw;
    st.kernel = sst[0].kernel;
    st.table_slots = sst[0].table_slots;
    st.group_lanes = sst[0].group_lanes;
    st.ms_total = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    free_all(host);
    out->nterms = total;
    out->key_words = W;
    out->cf_kind = f64 ? OBK_CF_F64 : OBK_CF_INT;
    out->cf_limbs = cw;
    out->log2_nsegs = k;
    out->mem = OBK_MEM_HOST;
    out->nsegs = (uint32_t)groups;
    out->seg_kind = OBK_SEG_HASH_POW2;
    out->keys = (uint64_t *)own->hk;
    out->cfs = own->hc;
    out->seg_offsets = own->seg_offsets;
    out->owner = own;
    if (stats) *stats = st;
    return OBK_OK;
}

obk_status obk_poly_copy(obk_engine *e, const obk_poly_view *x, uint32_t result_mem, obk_poly_result *out)
{
    if (!e || !out || !x) return OBK_ERR_INVALID_ARG;
    OBK_FORWARD_MULTI(e, obk_poly_copy(e->subs[0], x, result_mem, out));
    return copy_impl(e, x, result_mem, -1, out);
}

obk_status obk_poly_extend_symbols(obk_engine *e, const obk_poly_view *x, uint32_t new_nvars, const uint32_t *pos,
                                   uint32_t result_mem, obk_poly_result *out)
{    if (!e) return OBK_ERR_INVALID_ARG;
    OBK_FORWARD_MULTI(e, obk_poly_extend_symbols(e->subs[0], x, new_nvars, pos, result_mem, out));

    if (!e || !out || !x) return OBK_ERR_INVALID_ARG;
    std::memset(out, 0, sizeof(*out));
    obk_stats st{};
    e->launches = 0;
    try {
        CU_CHECK(e, cudaSetDevice(e->device));
        validate_view(e, x, "x");
        if (new_nvars < x->nvars || new_nvars > (uint32_t)MAX_VARS || (x->nvars && !pos)) {
            e->err = "symbol-set extension: the new symbol set must contain the old one (at most 64 symbols)";
            return OBK_ERR_INVALID_ARG;
        }
        for (uint32_t j = 0; j < x->nvars; ++j)
            if (pos[j] >= new_nvars || (j && pos[j] <= pos[j - 1])) {
                e->err = "symbol-set extension: positions must be strictly increasing and inside the new symbol set";
                return OBK_ERR_INVALID_ARG;
            }
        obk_poly_view nv = *x;
        nv.nvars = new_nvars;
        nv.key_words = x->psize ? std::max(1u, (new_nvars + x->psize - 1) / x->psize) : 1u;
        if (x->psize == 0 && new_nvars > key_max_size(*x)) {
            e->err = "Invalid size specified in the constructor of a Kronecker packer: too many symbols for one packed word";
            return OBK_ERR_INVALID_ARG;
        }
        if (nv.key_words > 4) {
            e->err = "symbol-set extension: key_words must be in [1,4]";
            return OBK_ERR_UNSUPPORTED;
        }
        const bool f64 = x->cf_kind == OBK_CF_F64;
        const int W = (int)x->key_words, Wn = (int)nv.key_words, cw = f64 ? 1 : (int)x->cf_limbs;
        const uint64_t n = x->nterms;
        if (n == 0) {
            set_empty_result(out, &nv, cw, result_mem);
            return OBK_OK;
        }
        ExtendParams P{};
        P.lo = make_layout(*x);
        P.ln = make_layout(nv);
        i128 lmin = 0, lmax = 0;
        component_lims(nv, P.ln.wsize ? P.ln.wsize : 1, lmin, lmax);
        P.lim_max_new = (uint64_t)lmax;
        for (uint32_t j = 0; j < x->nvars; ++j) P.pos[j] = pos[j];
        P.n = n;
        Scratch sc(e);
        cudaStream_t s = e->stream;
        CU_CHECK(e, cudaEventRecord(e->ev[0], s));
        const uint64_t *dk = x->keys;
        uint64_t *rk = sc.alloc_out<uint64_t>(n * Wn);
        uint64_t *rc = sc.alloc_out<uint64_t>(n * cw);
        if (x->mem == OBK_MEM_HOST) {
            uint64_t *k = sc.alloc<uint64_t>(n * W);
            CU_CHECK(e, cudaMemcpyAsync(k, x->keys, n * W * 8, cudaMemcpyHostToDevice, s));
            dk = k;
        }
        CU_CHECK(e, cudaMemcpyAsync(rc, x->cfs, n * cw * 8, x->mem == OBK_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, s));
        uint32_t *dover = sc.alloc<uint32_t>(4);
        CU_CHECK(e, cudaMemsetAsync(dover, 0, 16, s));
        P.keys = dk;
        P.okeys = rk;
        P.overflow = dover;
        k_extend_symbols<<<(unsigned)std::min<uint64_t>((n + 127) / 128, 4096), 128, 0, s>>>(P);
        e->launches += 1;
        CU_CHECK(e, cudaGetLastError());
        uint32_t hover = 0;
        CU_CHECK(e, cudaMemcpyAsync(&hover, dover, 4, cudaMemcpyDeviceToHost, s));
        CU_CHECK(e, cudaStreamSynchronize(s));
        if (hover) {
            e->err = "Cannot push a value to this Kronecker packer: an exponent is outside the range allowed by the "
                     "extended symbol set (overflow while merging symbols)";
            return OBK_ERR_KEY_OVERFLOW;
        }
        TailIn tin{rk, rc, n, Wn, cw, f64, P.ln.is_signed, n, result_mem, 1, false, 1};
        emit_result(e, sc, tin, out, st);
        return OBK_OK;
    } catch (const obk_fail &f) {
        if (out->owner) obk_result_free(e, out);
        std::memset(out, 0, sizeof(*out));
        cudaStreamSynchronize(e->stream);
        return f.st;
    }
}

obk_status obk_poly_truncate_degree(obk_engine *e, const obk_poly_view *x, const obk_trunc *t, uint32_t result_mem,
                                    obk_poly_result *out, obk_stats *stats)
{    if (!e) return OBK_ERR_INVALID_ARG;
    OBK_FORWARD_MULTI(e, obk_poly_truncate_degree(e->subs[0], x, t, result_mem, out, stats));

    if (!e || !out || !x || !t) return OBK_ERR_INVALID_ARG;
    std::memset(out, 0, sizeof(*out));
    obk_stats st{};
    e->launches = 0;
    try {
        CU_CHECK(e, cudaSetDevice(e->device));
        validate_view(e, x, "x");
        const bool f64 = x->cf_kind == OBK_CF_F64;
        const int W = (int)x->key_words, cw = f64 ? 1 : (int)x->cf_limbs;
        const uint64_t n = x->nterms;
        if (n == 0) {
            set_empty_result(out, x, cw, result_mem);
            if (stats) *stats = st;
            return OBK_OK;
        }
        const Layout L = make_layout(*x);
        uint64_t var_mask = L.nvars >= 64 ? ~0ull : ((1ull << L.nvars) - 1);
        if (t->partial) {
            var_mask = 0;
            for (uint32_t i = 0; i < t->nidx; ++i) {
                if (t->idx[i] >= L.nvars) {
                    e->err = "partial-degree index outside the symbol set";
                    return OBK_ERR_INVALID_ARG;
                }
                var_mask |= 1ull << t->idx[i];
            }
        }
        Scratch sc(e);
        cudaStream_t s = e->stream;
        CU_CHECK(e, cudaEventRecord(e->ev[0], s));
        const uint64_t *dk = x->keys, *dc = (const uint64_t *)x->cfs;
        if (x->mem == OBK_MEM_HOST) {
            uint64_t *k = sc.alloc<uint64_t>(n * W);
            uint64_t *c = sc.alloc<uint64_t>(n * cw);
            CU_CHECK(e, cudaMemcpyAsync(k, x->keys, n * W * 8, cudaMemcpyHostToDevice, s));
            CU_CHECK(e, cudaMemcpyAsync(c, x->cfs, n * cw * 8, cudaMemcpyHostToDevice, s));
            st.h2d_bytes += n * (W + cw) * 8;
            dk = k;
            dc = c;
        }
        StatsOut *dstats = sc.alloc<StatsOut>(1);
        int64_t *deg = sc.alloc<int64_t>(n);
        uint32_t *flag = sc.alloc<uint32_t>(n + 1);
        uint32_t *pos = sc.alloc<uint32_t>(n + 2);
        const unsigned g = (unsigned)std::min<uint64_t>((n + 255) / 256, 1024);
        k_stats_init<<<1, 64, 0, s>>>(dstats, 1);
        k_term_stats<<<g, 256, 0, s>>>(dk, dk, n, L, var_mask, 2 /* keys only */, 1, deg, dstats);
        k_flag_keep<<<g, 256, 0, s>>>(deg, n, t->limit, L.is_signed ? 0 : 1, flag);
        e->launches += 3;
        device_scan(e, sc, flag, pos, n);
        uint32_t htot = 0;
        CU_CHECK(e, cudaMemcpyAsync(&htot, pos + n, 4, cudaMemcpyDeviceToHost, s));
        CU_CHECK(e, cudaStreamSynchronize(s));
        const uint64_t total = htot;
        uint64_t *rk = sc.alloc_out<uint64_t>(std::max<uint64_t>(total, 1) * W);
        uint64_t *rc = sc.alloc_out<uint64_t>(std::max<uint64_t>(total, 1) * cw);
        int64_t *odeg = sc.alloc<int64_t>(std::max<uint64_t>(total, 1));
        k_select<<<g, 256, 0, s>>>(dk, dc, deg, n, W, cw, pos, rk, rc, odeg);
        e->launches += 1;
        CU_CHECK(e, cudaGetLastError());
        st.nterms_out = total;
        st.acc_limbs = cw;
        TailIn tin{rk, rc, total, W, cw, f64, L.is_signed, n, result_mem, 1, false, 1};
        emit_result(e, sc, tin, out, st);
        float ms = 0;
        cudaEventElapsedTime(&ms, e->ev[0], e->ev[8]);
        st.ms_total = ms;
        st.total_launches = e->launches;
        if (stats) *stats = st;
        return OBK_OK;
    } catch (const obk_fail &f) {
        if (out->owner) obk_result_free(e, out);
        std::memset(out, 0, sizeof(*out));
        cudaStreamSynchronize(e->stream);
        return f.st;
    }
}

void obk_result_free(obk_engine *e, obk_poly_result *r)
{
    if (!r || !r->owner) return;
    auto *own = static_cast<ResultOwner *>(r->owner);
    obk_engine *eng = own->e ? own->e : e;
    if (eng) {
        if (own->dkeys) cudaFreeAsync(own->dkeys, eng->stream);
        if (own->dcfs) cudaFreeAsync(own->dcfs, eng->stream);
        if (own->pk >= 0) pinned_release(eng, own->pk);
        if (own->pc >= 0) pinned_release(eng, own->pc);
    }
    std::free(own->seg_offsets);
    std::free(own->hk);
    std::free(own->hc);
    delete own;
    std::memset(r, 0, sizeof(*r));
}

obk_status obk_microbench(obk_engine *e, const char *name, double *value)
{    if (!e) return OBK_ERR_INVALID_ARG;
    OBK_FORWARD_MULTI(e, obk_microbench(e->subs[0], name, value));

    if (!e || !name || !value) return OBK_ERR_INVALID_ARG;
    *value = 0;
    if (std::string(name) != "int_mac") {
        e->err = "unknown micro-benchmark (known: int_mac)";
        return OBK_ERR_INVALID_ARG;
    }
    try {
        CU_CHECK(e, cudaSetDevice(e->device));
        Scratch sc(e);
        cudaStream_t s = e->stream;
        const unsigned grid = (unsigned)e->sm_count * 8u, threads = 256;
        const uint32_t iters = 1u << 15;
        uint64_t *in = sc.alloc<uint64_t>(256), *out = sc.alloc<uint64_t>((size_t)grid * threads);
        std::vector<uint64_t> h(256);
        for (size_t i = 0; i < h.size(); ++i) h[i] = 0x9E3779B97F4A7C15ULL * (i + 1) >> 3; // 61-bit unsigned operands
        CU_CHECK(e, cudaMemcpyAsync(in, h.data(), 256 * 8, cudaMemcpyHostToDevice, s));
        double best = 0;
        for (int rep = 0; rep < 4; ++rep) { // first repetition warms up
            CU_CHECK(e, cudaEventRecord(e->ev[0], s));
            k_bench_int_mac<<<grid, threads, 0, s>>>(in, iters, out);
            CU_CHECK(e, cudaEventRecord(e->ev[1], s));
            CU_CHECK(e, cudaStreamSynchronize(s));
            float ms = 0;
            cudaEventElapsedTime(&ms, e->ev[0], e->ev[1]);
            const double rate = (double)grid * threads * iters * 4.0 / (ms * 1e-3);
            if (rep > 0) best = std::max(best, rate);
        }
        CU_CHECK(e, cudaGetLastError());
        *value = best;
        return OBK_OK;
    } catch (const obk_fail &f) {
        return f.st;
    }
}

obk_status obk_overflow_check(obk_engine *e, const obk_poly_view *x, const obk_poly_view *y, int *ok)
{    if (!e) return OBK_ERR_INVALID_ARG;
    OBK_FORWARD_MULTI(e, obk_overflow_check(e->subs[0], x, y, ok));

    if (!e || !x || !y || !ok) return OBK_ERR_INVALID_ARG;
    // Same statistics pass as the product; report the verdict only.
    *ok = 1;
    if (x->nterms == 0 || y->nterms == 0 || x->nvars == 0) return OBK_OK;
    try {
        CU_CHECK(e, cudaSetDevice(e->device));
        validate_view(e, x, "x");
        validate_view(e, y, "y");
        const Layout L = make_layout(*x);
        const int W = x->key_words;
        Scratch sc(e);
        cudaStream_t s = e->stream;
        StatsOut *dstats = sc.alloc<StatsOut>(2);
        k_stats_init<<<2, 64, 0, s>>>(dstats, 2);
        const obk_poly_view *vs[2] = {x, y};
        for (int i = 0; i < 2; ++i) {
            const obk_poly_view *v = vs[i];
            const uint64_t *dk = v->keys;
            if (v->mem == OBK_MEM_HOST) {
                uint64_t *k = sc.alloc<uint64_t>(v->nterms * W);
                CU_CHECK(e, cudaMemcpyAsync(k, v->keys, v->nterms * W * 8, cudaMemcpyHostToDevice, s));
                dk = k;
            }
            const unsigned g = (unsigned)std::min<uint64_t>((v->nterms + 255) / 256, 1024);
            k_term_stats<<<g, 256, 0, s>>>(dk, dk, v->nterms, L, 0, 2 /* no coefficient pass */, 1, nullptr, dstats + i);
        }
        StatsOut hs[2];
        CU_CHECK(e, cudaMemcpyAsync(hs, dstats, sizeof(hs), cudaMemcpyDeviceToHost, s));
        CU_CHECK(e, cudaStreamSynchronize(s));
        i128 lim_min = 0, lim_max;
        component_lims(*x, L.wsize, lim_min, lim_max);
        for (uint32_t v = 0; v < L.nvars; ++v) {
            const i128 mn = unord(hs[0].emin[v], L.is_signed) + unord(hs[1].emin[v], L.is_signed);
            const i128 mx = unord(hs[0].emax[v], L.is_signed) + unord(hs[1].emax[v], L.is_signed);
            if (mn < lim_min || mx > lim_max) *ok = 0;
        }
        return OBK_OK;
    } catch (const obk_fail &f) {
        return f.st;
    }
}

obk_status obk_key_degrees(obk_engine *e, const obk_poly_view *x, const obk_trunc *which, int64_t *out_host)
{    if (!e) return OBK_ERR_INVALID_ARG;
    OBK_FORWARD_MULTI(e, obk_key_degrees(e->subs[0], x, which, out_host));

    if (!e || !x || !out_host) return OBK_ERR_INVALID_ARG;
    if (x->nterms == 0) return OBK_OK;
    try {
        CU_CHECK(e, cudaSetDevice(e->device));
        validate_view(e, x, "x");
        const Layout L = make_layout(*x);
        const int W = x->key_words;
        uint64_t var_mask = L.nvars >= 64 ? ~0ull : ((1ull << L.nvars) - 1);
        if (which && which->partial) {
            var_mask = 0;
            for (uint32_t i = 0; i < which->nidx; ++i) {
                if (which->idx[i] >= L.nvars) {
                    e->err = "partial-degree index outside the symbol set";
                    return OBK_ERR_INVALID_ARG;
                }
                var_mask |= 1ull << which->idx[i];
            }
        }
        Scratch sc(e);
        cudaStream_t s = e->stream;
        StatsOut *dstats = sc.alloc<StatsOut>(1);
        k_stats_init<<<1, 64, 0, s>>>(dstats, 1);
        const uint64_t *dk = x->keys;
        if (x->mem == OBK_MEM_HOST) {
            uint64_t *k = sc.alloc<uint64_t>(x->nterms * W);
            CU_CHECK(e, cudaMemcpyAsync(k, x->keys, x->nterms * W * 8, cudaMemcpyHostToDevice, s));
            dk = k;
        }
        int64_t *dd = sc.alloc<int64_t>(x->nterms);
        const unsigned g = (unsigned)std::min<uint64_t>((x->nterms + 255) / 256, 1024);
        k_term_stats<<<g, 256, 0, s>>>(dk, dk, x->nterms, L, var_mask, 2, 1, dd, dstats);
        CU_CHECK(e, cudaMemcpyAsync(out_host, dd, x->nterms * 8, cudaMemcpyDeviceToHost, s));
        CU_CHECK(e, cudaStreamSynchronize(s));
        return OBK_OK;
    } catch (const obk_fail &f) {
        return f.st;
    }
}

} // extern "C"
