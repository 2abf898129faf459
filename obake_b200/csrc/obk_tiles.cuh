// obk_tiles.cuh -- tile-blocked dense path ("kernel 3") of the B200 series-product engine (sm_100a).
//
// Same algebra as the plane path (obk_planes.cuh): key = lo0 + delta * lo1 + delta^2 * hi2 (kpack.hpp:262-267), the terms
// sharing hi2 form a PLANE indexed [lo1][lo0], and the product of two planes is a 2-D convolution that lands in the
// single output plane hi2 + hi2' (digit-wise sums never carry once the exponent-overflow check of
// polynomial.hpp:839-851 has passed).  What changes is how the convolution is mapped onto a warp:
//   * a warp owns a TILE of 4 output rows x 8 output columns of the output plane (lane = (r, c)), not one output row.
//     For a scalar S[b][i] of one plane the 32 lanes read the 4 x 8 window V[j0 + r - b][c0 + c - i] of the other one:
//     a window of a simplex-shaped plane is far better filled than a 32-wide row of it (Fateman n=30: 1.26e8 warp
//     steps for 2.15e9 term products instead of 1.48e8), and there is ONE code path -- no half / quarter-warp cases,
//     no high / low halves of long rows, no choice of the shorter row per row pair;
//   * because x sits at  S0 + (b P + i) 8  and y at  V0 + ((j0 + r - b) P + c0 + c - i) 8, the two addresses add up to
//     a constant of the (plane pair, tile, lane): y = C - x.  The row loop is two byte loads (row length of S, widest
//     row of the 4-row window of V), four integer operations for the scalar range and the multiply-accumulate steps;
//   * planes may be up to 64 x 64 (exponents of the two lowest symbols up to 63), rows at pitch P = W16 + 8 words
//     (P / 8 odd: the four rows of a tile fall into different banks); the zero tail of every row is the left / right
//     padding of the windows, three zero rows above and below the plane are their top / bottom padding;
//   * a total- or partial-degree truncation (polynomial.hpp:1337-1347) is a filter on the OUTPUT position: the degree
//     of a product term is linear in its exponents, so every term of an output plane beyond the limit is simply never
//     formed (tiles beyond the limit are not enumerated, the rest is masked when the rows are expanded to terms).
// Staging (TMA bulk copies into a shared-memory ring, one producer warp), the symbolic plane product, exact
// carry-save accumulators and the multi-GPU ownership rules are those of the plane path.
#pragma once
#include "obk_planes.cuh"

namespace obk
{

constexpr int TL_R = 4, TL_C = 8;         // tile: 4 output rows x 8 output columns
constexpr int TL_MAX_DIM = 64;            // rows / row width of an operand plane
constexpr int TL_HDR_BYTES = 144;         // len[64] u8 (row lengths), wm[72] u8 (widest row of the window k-3 .. k), pad
constexpr int TL_HDR = TL_HDR_BYTES / 8;  // words
constexpr int TL_PAD_ROWS = TL_R - 1;     // zero rows above / below a plane in its vector role
constexpr int TL_LEAD_BYTES = 64;         // zero words in front of the first zero row (left padding of its windows)
constexpr int TL_MAX_STAGES = 8;
constexpr int TL_DESC_BYTES = 128;        // descriptor of a staged pair: 16 words + the tile-row prefix table u16[32]
constexpr int TL_MAX_THREADS = 704;

// Row pitch in words: the widest row rounded to 16 plus 8 zero words; P / 8 is odd.
__host__ __device__ inline uint32_t tl_pitch(uint32_t wmax)
{
    return ((wmax + 15u) & ~15u) + 8u;
}
// A plane in HBM: header, rows [0, rcap) at pitch P, TL_PAD_ROWS zero rows.
__host__ __device__ inline uint32_t tl_blob_words(uint32_t rcap, uint32_t P)
{
    return TL_HDR + (rcap + TL_PAD_ROWS) * P;
}
// Shared-memory ring stage:  [S: header | rows]  [V: header | 64 B zeros | 3 zero rows | rows + 3 zero rows]  [descriptor]
struct TileStage {
    uint32_t s_rows, v_hdr, v_rows, desc, bytes;
};
__host__ __device__ inline TileStage tl_stage(uint32_t rcap, uint32_t P)
{
    TileStage g;
    const uint32_t P8 = P * 8u;
    g.s_rows = TL_HDR_BYTES;
    g.v_hdr = g.s_rows + rcap * P8;
    g.v_rows = g.v_hdr + TL_HDR_BYTES + TL_LEAD_BYTES + TL_PAD_ROWS * P8;
    g.desc = g.v_rows + (rcap + TL_PAD_ROWS) * P8;
    g.bytes = (g.desc + (uint32_t)TL_DESC_BYTES + 127u) & ~127u;
    return g;
}
// Tiles of an output plane with `rows` rows, rows at most `w` wide and lo1 + lo0 < d: row-major over the tile rows;
// tile row tr holds ceil(min(w, d - 4 tr) / 8) tiles.
__host__ __device__ inline uint32_t tl_ntiles(uint32_t rows, uint32_t w, uint32_t d)
{
    uint32_t n = 0;
    const uint32_t jend = rows < d ? rows : d;
    for (uint32_t j0 = 0; j0 < jend; j0 += TL_R) {
        const uint32_t ww = w < d - j0 ? w : d - j0;
        n += (ww + TL_C - 1u) / TL_C;
    }
    return n;
}
// tile number -> origin (j0, c0); false: no such tile
__host__ __device__ inline bool tl_locate(uint32_t idx, uint32_t rows, uint32_t w, uint32_t d, uint32_t &j0, uint32_t &c0)
{
    const uint32_t jend = rows < d ? rows : d;
    for (uint32_t j = 0; j < jend; j += TL_R) {
        const uint32_t ww = w < d - j ? w : d - j;
        const uint32_t n = (ww + TL_C - 1u) / TL_C;
        if (idx < n) {
            j0 = j;
            c0 = idx * TL_C;
            return true;
        }
        idx -= n;
    }
    return false;
}
// shape of an output plane: rows | widest row << 8 | (largest lo1 + lo0) + 1 << 16
__host__ __device__ inline uint32_t tl_shape(uint32_t rows, uint32_t w, uint32_t d)
{
    return rows | (w << 8) | (d << 16);
}

// Groups of an output plane: whole tile rows, packed greedily into at most `slots` tiles (slots >= tiles of one tile row).
// Returns the tile-row range [tr_lo, tr_hi) of group g and its tile count; tr_lo == tr_hi: no such group.
__host__ __device__ inline uint32_t tl_group(uint32_t shape, uint32_t slots, uint32_t g, uint32_t &tr_lo, uint32_t &tr_hi)
{
    const uint32_t rows = shape & 0xffu, w = (shape >> 8) & 0xffu, d = shape >> 16;
    const uint32_t jend = rows < d ? rows : d;
    uint32_t cur = 0, n = 0, lo = 0, tr = 0;
    for (uint32_t j = 0; j < jend; j += TL_R, ++tr) {
        const uint32_t ww = w < d - j ? w : d - j;
        const uint32_t c = (ww + TL_C - 1u) / TL_C;
        if (n + c > slots) {
            if (cur == g) {
                tr_lo = lo;
                tr_hi = tr;
                return n;
            }
            ++cur;
            lo = tr;
            n = 0;
        }
        n += c;
    }
    if (cur == g && n) {
        tr_lo = lo;
        tr_hi = tr;
        return n;
    }
    tr_lo = tr_hi = 0;
    return 0;
}
__host__ __device__ inline uint32_t tl_ngroups(uint32_t shape, uint32_t slots)
{
    const uint32_t rows = shape & 0xffu, w = (shape >> 8) & 0xffu, d = shape >> 16;
    const uint32_t jend = rows < d ? rows : d;
    uint32_t cur = 0, n = 0;
    for (uint32_t j = 0; j < jend; j += TL_R) {
        const uint32_t ww = w < d - j ? w : d - j;
        const uint32_t c = (ww + TL_C - 1u) / TL_C;
        if (n + c > slots) {
            ++cur;
            n = 0;
        }
        n += c;
    }
    return cur + (n ? 1u : 0u);
}
// tiles of tile row tr of a shape
__host__ __device__ inline uint32_t tl_row_tiles(uint32_t shape, uint32_t tr)
{
    const uint32_t rows = shape & 0xffu, w = (shape >> 8) & 0xffu, d = shape >> 16;
    const uint32_t jend = rows < d ? rows : d, j = tr * TL_R;
    if (j >= jend) return 0u;
    const uint32_t ww = w < d - j ? w : d - j;
    return (ww + TL_C - 1u) / TL_C;
}

// ---- plane building ------------------------------------------------------------------------------------------------
// Scatter every term's coefficient into its plane's blob; row lengths and term counts by atomics.
__global__ void __launch_bounds__(256) k_tile_fill(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ cfs, uint64_t n,
                                                   uint64_t delta, const uint64_t *__restrict__ tab, uint32_t log2_cap,
                                                   const uint32_t *__restrict__ slot_id, uint64_t *__restrict__ mat, uint32_t blob_words,
                                                   uint32_t P, uint32_t rcap, uint32_t *__restrict__ rowlen, uint32_t *__restrict__ pinfo)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint32_t mask = (1u << log2_cap) - 1u;
    const uint64_t d2 = delta * delta;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t key = keys[i];
        const uint64_t hi = key / d2;
        const uint64_t rem = key - hi * d2;
        const uint32_t lo1 = (uint32_t)(rem / delta);
        const uint32_t lo0 = (uint32_t)(rem - (uint64_t)lo1 * delta);
        uint32_t h = plane_hash(hi, log2_cap);
        while (tab[h] != hi) h = (h + 1u) & mask;
        const uint32_t p = slot_id[h];
        mat[(size_t)p * blob_words + TL_HDR + lo1 * P + lo0] = cfs[i];
        atomicMax(&rowlen[(size_t)p * rcap + lo1], lo0 + 1u);
        atomicAdd(&pinfo[4 * (size_t)p + 1], 1u);
    }
}

// One warp per plane: rows, widest row, largest lo1 + lo0 (+1), and the header bytes of the blob.
//   pinfo[4 p + 0] rows, [1] terms, [2] widest row, [3] max(b + len[b])
__global__ void __launch_bounds__(256) k_tile_finish(uint32_t nplanes, uint32_t rcap, const uint32_t *__restrict__ rowlen,
                                                     uint64_t *__restrict__ mat, uint32_t blob_words, uint32_t *__restrict__ pinfo)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= nplanes) return;
    const uint32_t *len = rowlen + (size_t)p * rcap;
    uint8_t *hdr = reinterpret_cast<uint8_t *>(mat + (size_t)p * blob_words);
    uint32_t rows = 0, w = 0, d = 0;
    // row lengths: lanes over the rows (at most 64)
    for (uint32_t b = lane; b < (uint32_t)TL_MAX_DIM; b += 32u) {
        const uint32_t l = b < rcap ? len[b] : 0u;
        hdr[b] = (uint8_t)l;
        if (l) {
            rows = max(rows, b + 1u);
            w = max(w, l);
            d = max(d, b + l);
        }
    }
    // wm[k] = widest of the rows k-3 .. k (rows outside the plane count as empty), k = 0 .. 71
    for (uint32_t k = lane; k < 72u; k += 32u) {
        uint32_t m = 0;
        for (uint32_t q = 0; q < (uint32_t)TL_R; ++q) {
            const uint32_t b = k - q; // wraps for k < q
            if (b < rcap) m = max(m, len[b]);
        }
        hdr[64 + k] = (uint8_t)m;
    }
    rows = __reduce_max_sync(0xffffffffu, rows);
    w = __reduce_max_sync(0xffffffffu, w);
    d = __reduce_max_sync(0xffffffffu, d);
    if (lane == 0) {
        pinfo[4 * (size_t)p] = rows;
        pinfo[4 * (size_t)p + 2] = w;
        pinfo[4 * (size_t)p + 3] = d;
    }
}

// ---- symbolic plane product ----------------------------------------------------------------------------------------
// Per-slot records of the output-plane set: rec = {pairs, rows, widest row, output plane number}, dmax, work.
__global__ void __launch_bounds__(256) k_tile_sym_insert(const uint64_t *__restrict__ xh, const uint32_t *__restrict__ xinfo, uint32_t nx,
                                                         const uint64_t *__restrict__ yh, const uint32_t *__restrict__ yinfo, uint32_t ny,
                                                         uint32_t part_rank, uint32_t part_nranks, uint64_t *__restrict__ tab,
                                                         uint32_t log2_cap, uint4 *__restrict__ rec, uint32_t *__restrict__ dmax,
                                                         unsigned long long *__restrict__ work, uint64_t *__restrict__ okey,
                                                         uint32_t *__restrict__ counters)
{
    const uint64_t total = (uint64_t)nx * ny, stride = (uint64_t)gridDim.x * blockDim.x;
    const uint32_t mask = (1u << log2_cap) - 1u;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const uint32_t a = (uint32_t)(i / ny), b = (uint32_t)(i - (uint64_t)a * ny);
        const uint64_t ka = xh[a];
        if (part_nranks > 1u && plane_owner(ka, part_nranks) != part_rank) continue;
        const uint64_t key = ka + yh[b];
        uint32_t h = plane_hash(key, log2_cap);
        for (;;) {
            const unsigned long long old = atomicCAS((unsigned long long *)&tab[h], (unsigned long long)ROW_EMPTY, (unsigned long long)key);
            if (old == ROW_EMPTY) {
                const uint32_t o = atomicAdd(&counters[0], 1u);
                rec[h].w = o;
                okey[o] = key;
                break;
            }
            if (old == key) break;
            h = (h + 1u) & mask;
        }
        const uint4 ia = *reinterpret_cast<const uint4 *>(xinfo + 4 * (size_t)a), ib = *reinterpret_cast<const uint4 *>(yinfo + 4 * (size_t)b);
        uint32_t *r = reinterpret_cast<uint32_t *>(&rec[h]);
        atomicAdd(&r[0], 1u);
        atomicAdd(&work[h], (unsigned long long)ia.y * ib.y);
        atomicMax(&r[1], ia.x + ib.x - 1u);
        atomicMax(&r[2], ia.z + ib.z - 1u);
        atomicMax(&dmax[h], ia.w + ib.w - 1u);
    }
}

__global__ void __launch_bounds__(256) k_tile_sym_compact(const uint64_t *__restrict__ tab, uint32_t cap, const uint4 *__restrict__ rec,
                                                          const uint32_t *__restrict__ dmax, const unsigned long long *__restrict__ work,
                                                          uint32_t *__restrict__ ocnt, unsigned long long *__restrict__ owork,
                                                          uint32_t *__restrict__ oshape)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap || tab[i] == ROW_EMPTY) return;
    const uint4 r = rec[i];
    ocnt[r.w] = r.x;
    owork[r.w] = work[i];
    oshape[r.w] = tl_shape(r.y, r.z, dmax[i]);
}

// Truncation of a dense product (see the file header): the degree of the term at (lo1, lo0) of output plane o is
// dhi[o] + w1 lo1 + w0 lo0 with w0, w1 in {0, 1}; terms with degree > limit are never formed.  `on` = 0: no truncation.
struct TileTrunc {
    uint32_t on, w0, w1;
    int64_t limit;
    const int64_t *dhi; // [output planes] degree of the hi2 part (over the truncation's symbols)
};
// largest admissible value of w1 lo1 + w0 lo0 on a plane, clamped to [-1, 2^20]
__device__ __forceinline__ int32_t tl_budget(const TileTrunc &t, uint32_t o)
{
    if (!t.on) return 1 << 20;
    const int64_t v = t.limit - t.dhi[o];
    return (int32_t)(v < -1 ? -1 : (v > (1 << 20) ? (1 << 20) : v));
}
// the shape of an output plane cut by the truncation: only rows / columns / diagonals that hold admissible terms
__device__ __forceinline__ uint32_t tl_cut_shape(uint32_t shape, const TileTrunc &t, int32_t budget)
{
    if (!t.on) return shape;
    uint32_t rows = shape & 0xffu, w = (shape >> 8) & 0xffu, d = shape >> 16;
    if (budget < 0) return 0u;
    if (t.w1 && !t.w0) rows = min(rows, (uint32_t)budget + 1u);
    if (t.w0 && !t.w1) w = min(w, (uint32_t)budget + 1u);
    if (t.w0 && t.w1) {
        d = min(d, (uint32_t)budget + 1u);
        rows = min(rows, d);
        w = min(w, d);
    }
    return tl_shape(rows, w, d);
}

// One CTA: pair_off = exclusive scan of the pairs per plane, row_base = exclusive scan of the rows per plane, and the
// work units = (output plane, group of tile rows holding at most `tpu` tiles), heaviest planes first (counting sort by the bit length of the
// work).  own_nranks > 1: only the units this rank owns (multi-GPU owner mode).  The shapes are cut by the truncation.
//   units[u] = {plane | group << 24, chunk | chunks << 16};  totals[0] = units, [1] = output rows, [2] = output planes,
//   [3] = tiles per unit, [4] = 1 if some plane is cut into chunks of pairs (split mode)
constexpr uint32_t TL_SORT_MAX = 4096; // output planes the owner-mode assignment sorts (dynamic shared memory: 20 B each)

// rank of position q of the canonical unit sequence: 0 1 .. n-1 n-1 .. 1 0 0 1 ..
__device__ __forceinline__ uint32_t tl_snake(uint32_t q, uint32_t n)
{
    const uint32_t m = q % (2u * n);
    return m < n ? m : 2u * n - 1u - m;
}

__global__ void __launch_bounds__(1024) k_tile_units(const unsigned long long *__restrict__ owork, uint32_t *__restrict__ oshape,
                                                     const uint32_t *__restrict__ ocnt, const uint64_t *__restrict__ okey,
                                                     const uint32_t *__restrict__ counters, uint32_t tpu_max, uint32_t want_units, uint32_t want_scale,
                                                     uint32_t own_rank, uint32_t own_nranks, uint32_t allow_split, TileTrunc tr,
                                                     uint32_t *__restrict__ qbase, uint32_t *__restrict__ row_base,
                                                     uint32_t *__restrict__ pair_off, uint2 *__restrict__ units,
                                                     uint32_t *__restrict__ totals)
{
    extern __shared__ __align__(16) unsigned char sort_raw[];
    // unit counts per (work class, warp): the planes of a Fateman product fall into five or six classes, and 1 891
    // atomics on the same shared word serialise (most of this kernel's time in its first version)
    __shared__ uint32_t s_cls[65][33], s_tot, s_carry, s_cand[4];
    const uint32_t tid = threadIdx.x;
    const uint32_t nout = counters[0];
    for (uint32_t i = tid; i < 65u * 33u; i += blockDim.x) (&s_cls[0][0])[i] = 0;
    if (tid == 0) s_carry = 0;
    if (tid < 4) s_cand[tid] = 0;
    for (uint32_t o = tid; o < nout; o += blockDim.x) oshape[o] = tl_cut_shape(oshape[o], tr, tl_budget(tr, o));
    __syncthreads();
    // Tiles per unit: as many as the accumulator area holds (every pair of a plane is then staged once), unless that
    // leaves fewer than `want_units` units for the resident CTAs of this GPU -- a small product, or one rank's share of
    // a product spread over several GPUs: then the planes are cut into groups of a half, a quarter, an eighth of that.
    if (!allow_split) {
        uint32_t all[4] = {0, 0, 0, 0};
        for (uint32_t o = tid; o < nout; o += blockDim.x) {
#pragma unroll
            for (int k = 0; k < 4; ++k) all[k] += tl_ngroups(oshape[o], max(16u, tpu_max >> k));
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (all[k]) atomicAdd(&s_cand[k], all[k]);
    }
    __syncthreads();
    uint32_t tpu = tpu_max;
    for (int k = 0; k < 4 && !allow_split; ++k) { // with pair chunks (below) the planes need not be cut into tile groups
        tpu = max(16u, tpu_max >> k);
        if (s_cand[k] >= want_units * max(own_nranks, 1u) || tpu == 16u) break;
    }
    __syncthreads();
    if (tid == 0) {
        totals[3] = tpu;
        s_cand[0] = 0;
    }
    auto groups = [&](uint32_t o) { return tl_ngroups(oshape[o], tpu); };
    // Multi-GPU owner mode: who forms unit (o, b).  Every rank runs the same symbolic product but numbers the planes in
    // the order its own atomics resolved, so the rule must be a function of the plane KEYS and works alone: the planes
    // are sorted by (work descending, key), and dealt to the ranks back and forth along that order
    // (a longest-processing-time deal of whole planes: every output term is complete on one rank).  Beyond TL_SORT_MAX
    // planes a hash of the key decides (the law of large numbers does the balancing there).
    const bool sorted = own_nranks > 1u && nout <= TL_SORT_MAX;
    if (own_nranks > 1u) {
        if (sorted) {
            uint32_t n2 = 1;
            while (n2 < nout) n2 <<= 1;
            uint64_t *ka = reinterpret_cast<uint64_t *>(sort_raw), *kb = ka + n2;
            uint32_t *id = reinterpret_cast<uint32_t *>(kb + n2);
            for (uint32_t i = tid; i < n2; i += blockDim.x) {
                ka[i] = i < nout ? ~(uint64_t)owork[i] : ~0ull;
                kb[i] = i < nout ? okey[i] : ~0ull;
                id[i] = i < nout ? i : 0xffffffffu;
            }
            __syncthreads();
            for (uint32_t k = 2; k <= n2; k <<= 1) {
                for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                    // comparator c of this pass exchanges i = (c / j) 2j + c % j with i + j: every thread has one to do
                    for (uint32_t cidx = tid; cidx < (n2 >> 1); cidx += blockDim.x) {
                        const uint32_t i = ((cidx & ~(j - 1u)) << 1) | (cidx & (j - 1u)), l = i | j;
                        {
                            const bool up = (i & k) == 0u;
                            const bool gt = ka[i] > ka[l] || (ka[i] == ka[l] && kb[i] > kb[l]);
                            if (gt == up) {
                                const uint64_t ta = ka[i], tb = kb[i];
                                const uint32_t ti = id[i];
                                ka[i] = ka[l];
                                kb[i] = kb[l];
                                id[i] = id[l];
                                ka[l] = ta;
                                kb[l] = tb;
                                id[l] = ti;
                            }
                        }
                    }
                    __syncthreads();
                }
            }
            for (uint32_t i = tid; i < n2; i += blockDim.x)
                if (id[i] != 0xffffffffu) qbase[id[i]] = i; // position in the canonical order
        } else {
            for (uint32_t o = tid; o < nout; o += blockDim.x) qbase[o] = plane_owner(okey[o], own_nranks);
        }
        __syncthreads();
    }
    auto owned = [&](uint32_t o) {
        if (own_nranks <= 1u) return true;
        return (sorted ? tl_snake(qbase[o], own_nranks) : qbase[o]) == own_rank;
    };
    // Pair chunks: a unit stages the pairs of its plane one after the other (a few us each), so a plane with hundreds
    // of pairs would occupy one CTA for milliseconds while the others run dry.  Heavy planes are cut into K chunks of
    // pairs that different CTAs accumulate independently and add to the output rows with 64-bit atomics on 32-bit
    // digits (k_tileconv, split mode); the target is want_units units of equal work on this GPU.
    __shared__ unsigned long long s_work;
    if (tid == 0) s_work = 0;
    __syncthreads();
    {
        // (a 64-bit shared-memory atomicAdd is a compare-and-swap loop: 1 024 threads on one word took 50 us)
        unsigned long long w = 0;
        for (uint32_t o = tid; o < nout; o += blockDim.x)
            if (owned(o)) w += owork[o];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) w += __shfl_xor_sync(0xffffffffu, w, d);
        if ((tid & 31u) == 0u && w) atomicAdd(&s_work, w);
    }
    __syncthreads();
    // `want_units` = units per resident CTA x resident CTAs.  The more work a CTA has, the coarser its units may be (every
    // unit pays a ring fill and a flush); the less, the finer they must be for the last wave to stay short.  Measured
    // on F30 (7.3 M term products per CTA) and on one rank's share of it at 8 GPUs (0.9 M): 4 units per CTA are best
    // for the first (16: +1 %), 16 for the second (4: +14 %).
    if (want_scale) {
        const unsigned long long per_cta = s_work / max(want_units, 1u);
        const unsigned long long f = per_cta ? (16ull << 20) / per_cta : 16ull;
        want_units *= (uint32_t)min(16ull, max(4ull, f));
    }
    const unsigned long long w_unit = s_work / max(want_units, 1u) + 1ull;
    auto chunks = [&](uint32_t o, uint32_t nb) {
        if (!allow_split || nb == 0u) return 1u;
        const unsigned long long k = (owork[o] / nb + w_unit / 2ull) / w_unit;
        const uint32_t kmax = min(64u, max(1u, ocnt[o] / 4u));
        return (uint32_t)min((unsigned long long)kmax, max(1ull, k));
    };
    // (row_base and pair_off are scratch until the scans at the end fill them: groups | chunks << 8 and the work class of
    // every plane this rank forms are computed once)
    for (uint32_t o = tid; o < nout; o += blockDim.x) {
        uint32_t nb = 0, K = 1, c = 0;
        if (owned(o)) {
            c = 64u - (uint32_t)__clzll((long long)(owork[o] | 1ull));
            nb = groups(o);
            K = chunks(o, nb);
            if (nb) atomicAdd(&s_cls[64u - c][tid >> 5], nb * K); // class 0 = heaviest
            if (K > 1u) s_cand[0] = 1u;                  // (re-used as the split flag: same value from every writer)
        }
        row_base[o] = nb | (K << 8) | (c << 16);
    }
    __syncthreads();
    // exclusive offsets in (class, warp) order: every class row first (one thread each), then the class bases
    if (tid < 65) {
        uint32_t acc = 0;
        for (int w = 0; w < 32; ++w) {
            const uint32_t v = s_cls[tid][w];
            s_cls[tid][w] = acc;
            acc += v;
        }
        s_cls[tid][32] = acc; // class total
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t acc = 0;
        for (int c = 0; c < 65; ++c) {
            const uint32_t v = s_cls[c][32];
            s_cls[c][32] = acc;
            acc += v;
        }
        totals[0] = acc;
        totals[2] = nout;
        totals[4] = s_cand[0];
    }
    __syncthreads();
    for (uint32_t o = tid; o < nout; o += blockDim.x) {
        const uint32_t pk = row_base[o], nb = pk & 0xffu, K = (pk >> 8) & 0xffu, c = pk >> 16;
        if (nb == 0u) continue;
        uint32_t at = s_cls[64u - c][32] + atomicAdd(&s_cls[64u - c][tid >> 5], nb * K);
        for (uint32_t b = 0; b < nb; ++b)
            for (uint32_t k = 0; k < K; ++k, ++at) units[at] = make_uint2(o | (b << 24), k | (K << 16));
    }
    for (int which = 0; which < 2; ++which) {
        uint32_t *out = which == 0 ? row_base : pair_off;
        __syncthreads();
        if (tid == 0) s_carry = 0;
        __syncthreads();
        for (uint32_t base0 = 0; base0 < nout; base0 += SCAN_BLOCK) {
            const uint32_t base = base0 + tid * SCAN_ITEMS;
            uint32_t v[SCAN_ITEMS], sum = 0;
#pragma unroll
            for (int j = 0; j < SCAN_ITEMS; ++j) {
                // output rows exist only for the planes this rank forms (owner mode: 1 / nranks of them)
                v[j] = base + j < nout ? (which == 0 ? (owned(base + j) ? (oshape[base + j] & 0xffu) : 0u) : ocnt[base + j]) : 0;
                sum += v[j];
            }
            uint32_t ex = block_exclusive_scan(sum, &s_tot) + s_carry;
#pragma unroll
            for (int j = 0; j < SCAN_ITEMS; ++j) {
                if (base + j < nout) out[base + j] = ex;
                ex += v[j];
            }
            __syncthreads();
            if (tid == 0) s_carry += s_tot;
            __syncthreads();
        }
        if (tid == 0) {
            out[nout] = s_carry;
            if (which == 0) totals[1] = s_carry;
        }
    }
}

// degree of the hi2 part of every output plane over the truncation's symbols (var_mask, symbols 2 .. nvars-1)
__global__ void __launch_bounds__(256) k_tile_plane_degrees(const uint64_t *__restrict__ okey, const uint32_t *__restrict__ counters,
                                                            uint64_t delta, uint32_t nvars, uint64_t var_mask, int64_t *__restrict__ dhi)
{
    const uint32_t nout = counters[0];
    for (uint32_t o = blockIdx.x * blockDim.x + threadIdx.x; o < nout; o += gridDim.x * blockDim.x) {
        uint64_t k = okey[o];
        int64_t d = 0;
        for (uint32_t v = 2; v < nvars; ++v) {
            const uint64_t q = k / delta;
            if ((var_mask >> v) & 1ull) d += (int64_t)(k - q * delta);
            k = q;
        }
        dhi[o] = d;
    }
}

// `np` pairs of steps of one row pair: z[lane] += x[i] y[-i]; px / py are the shared addresses of x[0] and of y[0].
// Running pointers (two additions per four steps), every load at an immediate offset.
template <class Acc>
__device__ __forceinline__ void conv_run(Acc &z, uint32_t px, uint32_t py, uint32_t np)
{
    for (; np >= 2u; np -= 2u) {
        uint64_t x0, x1, x2, x3;
        lds_2u64(px, x0, x1);
        lds_2u64(px + 16u, x2, x3);
        const uint64_t y0 = lds_u64(py), y1 = lds_u64(py - 8u), y2 = lds_u64(py - 16u), y3 = lds_u64(py - 24u);
        z.mac(x0, y0);
        z.mac(x1, y1);
        z.mac(x2, y2);
        z.mac(x3, y3);
        px += 32u;
        py -= 32u;
    }
    if (np) {
        uint64_t x0, x1;
        lds_2u64(px, x0, x1);
        const uint64_t y0 = lds_u64(py), y1 = lds_u64(py - 8u);
        z.mac(x0, y0);
        z.mac(x1, y1);
    }
}

struct TileParams {
    const uint64_t *xmat, *ymat;   // [planes][blob_words]
    const uint32_t *xinfo, *yinfo; // [planes][4]: rows, terms, widest row, max(b + len[b])
    const uint2 *pairs;            // plane pairs grouped by output plane
    const uint32_t *pair_off;      // [nout + 1]
    const uint64_t *okey;          // [nout] output plane keys (hi2 sums)
    const uint32_t *oshape;        // [nout] tl_shape(), already cut by the truncation
    const uint32_t *row_base;      // [nout + 1]
    const uint2 *units;            // {plane | group << 24, chunk | chunks << 16}, heaviest first
    uint32_t nunits;
    uint32_t *cursor;
    uint64_t delta;
    uint64_t *or_hi;  // [rows] row key = okey * delta + j
    uint64_t *omat;   // [rows][LACC][wo]  (split mode: [rows][2 LACC][wo] digits; fp64: [rows][1][wo])
    uint32_t *or_cnt; // [rows][wo / 32] non-zero entries per 32-column chunk
    uint32_t P, rcap, blob_words, nstage, wo;
    uint32_t slots; // accumulator tiles in shared memory = most tiles of a unit
    uint32_t split; // 1: units add 32-bit digits to omat[rows][2 LACC][wo] with atomics (k_tile_norm finishes the rows)
    TileTrunc tr;
};


// Shared memory of the product kernel:  accumulators [slots][LACC][32] u64 | locks [slots] u32 | item counters
// [TL_MAX_STAGES] u32 | ring stages.
__host__ __device__ inline uint32_t tl_acc_bytes(uint32_t slots, uint32_t lacc)
{
    return (slots * lacc * 256u + slots * 4u + TL_MAX_STAGES * 4u + 127u) & ~127u;
}

// nw consumer warps (blockDim / 32 - 1) + one producer warp.  The unit of work of a consumer warp is an ITEM = (staged
// plane pair, output tile): items are handed out by a shared-memory counter per ring stage, so a warp never waits for
// a slower one -- tiles in the middle of a simplex-shaped plane carry 1.6x the work of the average tile, which is what
// a static tile-per-warp assignment loses.  An item is accumulated in registers and added to the tile's accumulator
// in shared memory under a per-tile lock; when the unit's last pair has been consumed the tiles are written out.
template <int LACC, bool F64, bool UNS, int BT, int MINB>
__global__ void __launch_bounds__(BT, MINB) k_tileconv(const TileParams P)
{
    using Acc = PlaneAcc<LACC, F64, UNS>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_full[TL_MAX_STAGES], s_empty[TL_MAX_STAGES];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t NW = (blockDim.x >> 5) - 1u;
    const uint32_t nstage = P.nstage;
    const uint32_t Pw = P.P, P8 = Pw * 8u;
    const TileStage G = tl_stage(P.rcap, Pw);
    const uint32_t acc_bytes = tl_acc_bytes(P.slots, LACC);
    uint64_t *s_acc = reinterpret_cast<uint64_t *>(smem_raw);
    uint32_t *s_lock = reinterpret_cast<uint32_t *>(smem_raw + (size_t)P.slots * LACC * 256u);
    uint32_t *s_next = s_lock + P.slots;
    unsigned char *ring = smem_raw + acc_bytes;
    // zero everything once: accumulators, locks, and the padding above a vector plane that is never written again
    for (uint32_t i = tid; i < (acc_bytes + nstage * G.bytes) / 16u; i += blockDim.x)
        reinterpret_cast<uint4 *>(smem_raw)[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy zeros before the bulk copies land
    if (tid == 0) {
        for (uint32_t s = 0; s < nstage; ++s) {
            mbar_init(&s_full[s], 1);
            mbar_init(&s_empty[s], NW);
        }
    }
    __syncthreads();

    if (warp == NW) {
        // ================= producer =================
        uint32_t s = 0, ph = 0;
        for (;;) {
            uint32_t u = 0;
            if (lane == 0) u = atomicAdd(P.cursor, 1u);
            u = __shfl_sync(0xffffffffu, u, 0);
            const bool quit = u >= P.nunits;
            uint32_t o = 0, g = 0, p0 = 0, p1 = 1, shape = 0, tr_lo = 0, tr_hi = 0;
            int32_t budget = 0;
            if (!quit) {
                const uint2 ud = __ldg(P.units + u);
                o = ud.x & 0x00ffffffu;
                g = ud.x >> 24;
                const uint32_t q0 = __ldg(P.pair_off + o), cnt = __ldg(P.pair_off + o + 1) - q0;
                const uint32_t ch = ud.y & 0xffffu, K = ud.y >> 16;
                p0 = q0 + (uint32_t)(((uint64_t)cnt * ch) / K);
                p1 = q0 + (uint32_t)(((uint64_t)cnt * (ch + 1u)) / K);
                shape = __ldg(P.oshape + o);
                tl_group(shape, P.slots, g, tr_lo, tr_hi);
                budget = tl_budget(P.tr, o);
            }
            bool first = true;
            for (uint32_t pb = p0; pb < p1; pb += 32u) {
                uint2 pr = make_uint2(0, 0);
                uint4 ia = make_uint4(0, 0, 0, 0), ib = ia;
                if (!quit && pb + lane < p1) {
                    pr = __ldg(P.pairs + pb + lane);
                    ia = __ldg(reinterpret_cast<const uint4 *>(P.xinfo) + pr.x);
                    ib = __ldg(reinterpret_cast<const uint4 *>(P.yinfo) + pr.y);
                }
                // what the stage needs: rows of both planes, which one supplies the scalars, the pair's output shape
                const bool a_is_s = ia.y <= ib.y; // fewer terms
                const uint32_t meta = (a_is_s ? ia.x : ib.x) | ((a_is_s ? ib.x : ia.x) << 8) | (a_is_s ? 0u : 1u << 16);
                const uint32_t pshape = tl_shape(min(ia.x + ib.x - 1u, shape & 0xffu), min(ia.z + ib.z - 1u, (shape >> 8) & 0xffu),
                                                 min(ia.w + ib.w - 1u, shape >> 16));
                const uint32_t cntp = quit ? 1u : min(32u, p1 - pb);
                for (uint32_t q = 0; q < cntp; ++q) {
                    const uint32_t px = __shfl_sync(0xffffffffu, pr.x, (int)q), py = __shfl_sync(0xffffffffu, pr.y, (int)q);
                    const uint32_t mt = __shfl_sync(0xffffffffu, meta, (int)q), psh = __shfl_sync(0xffffffffu, pshape, (int)q);
                    const bool last = pb + q + 1u == p1;
                    // items of this pair in this group: lane i looks at tile row tr_lo + i
                    const uint32_t mine = tr_lo + lane < tr_hi ? tl_row_tiles(psh, tr_lo + lane) : 0u;
                    uint32_t incl = mine;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
                        if ((int)lane >= d) incl += v;
                    }
                    const uint32_t nitems = __shfl_sync(0xffffffffu, incl, 31);
                    // a pair that cannot reach this group of tiles is not staged (unless it carries the end-of-unit flag)
                    if (!quit && !last && nitems == 0u) continue;
                    mbar_wait(&s_empty[s], ph ^ 1u);
                    unsigned char *st = ring + (size_t)s * G.bytes;
                    reinterpret_cast<uint16_t *>(st + G.desc + 64)[lane] = (uint16_t)incl;
                    __syncwarp();
                    if (lane == 0) {
                        uint32_t *d = reinterpret_cast<uint32_t *>(st + G.desc);
                        const uint32_t nS = mt & 0xffu, nV = (mt >> 8) & 0xffu;
                        d[0] = o;
                        d[1] = g;
                        d[2] = quit ? 2u : ((last ? 1u : 0u) | (first ? 4u : 0u));
                        d[3] = shape;
                        d[4] = mt;
                        d[5] = psh;
                        d[6] = (uint32_t)budget;
                        d[7] = quit ? 0u : nitems;
                        d[8] = tr_lo | (tr_hi << 8);
                        s_next[s] = 0;
                        const uint32_t bS = TL_HDR_BYTES + nS * P8, bV = (nV + TL_PAD_ROWS) * P8;
                        mbar_expect_tx(&s_full[s], quit ? 0u : bS + TL_HDR_BYTES + bV);
                        if (!quit) {
                            const uint64_t *A = P.xmat + (size_t)px * P.blob_words, *B = P.ymat + (size_t)py * P.blob_words;
                            const uint64_t *S = (mt >> 16) ? B : A, *V = (mt >> 16) ? A : B;
                            tma_bulk_g2s(st, S, bS, &s_full[s]);
                            tma_bulk_g2s(st + G.v_hdr, V, TL_HDR_BYTES, &s_full[s]);
                            tma_bulk_g2s(st + G.v_rows, V + TL_HDR, bV, &s_full[s]);
                        }
                    }
                    first = false;
                    if (++s == nstage) {
                        s = 0;
                        ph ^= 1u;
                    }
                }
            }
            if (quit) break;
        }
        return;
    }

    // ================= consumers =================
    const uint32_t r = lane >> 3, c = lane & 7u;
    const uint32_t smem0 = smem_u32(ring);
    uint32_t s = 0, ph = 0;
    uint32_t gp_excl = 0; // tiles of the unit's group in front of tile row tr_lo + lane (the slot of its first tile)
    uint32_t gtiles = 0;  // tiles of the unit's group
    for (;;) {
        mbar_wait(&s_full[s], ph);
        const uint32_t base = smem0 + s * G.bytes;
        const uint4 d0 = lds_u32x4(base + G.desc);
        if (d0.z & 2u) break;
        const uint4 d1 = lds_u32x4(base + G.desc + 16u);
        const uint32_t trr = lds_u32(base + G.desc + 32u), tr_lo = trr & 0xffu, tr_hi = trr >> 8;
        if (d0.z & 4u) {
            // a new unit: slots of the tile rows of its group
            const uint32_t mine = tr_lo + lane < tr_hi ? tl_row_tiles(d0.w, tr_lo + lane) : 0u;
            uint32_t incl = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
                if ((int)lane >= d) incl += v;
            }
            gp_excl = incl - mine;
            gtiles = __shfl_sync(0xffffffffu, incl, 31);
        }
        const uint32_t meta = d1.x, nitems = d1.w;
        const uint32_t nS = meta & 0xffu, nV = (meta >> 8) & 0xffu;
        const uint32_t prow = d1.y & 0xffu, pd = d1.y >> 16; // the pair's output rows / diagonal bound (clipped to the plane's)
        const uint32_t aS = base + G.s_rows, aV = base + G.v_rows; // shared addresses of S[0][0] and V[0][0]
        const uint32_t aWm = base + G.v_hdr + 64u + (uint32_t)TL_PAD_ROWS;
        const uint32_t pin = lds_u16(base + G.desc + 64u + 2u * lane); // inclusive tile count up to tile row tr_lo + lane
        for (;;) {
            uint32_t idx = 0;
            if (lane == 0) idx = atomicAdd(&s_next[s], 1u);
            idx = __shfl_sync(0xffffffffu, idx, 0);
            if (idx >= nitems) break;
            const uint32_t trel = __popc(__ballot_sync(0xffffffffu, idx >= pin));
            const uint32_t prev = __shfl_sync(0xffffffffu, pin, (int)((trel + 31u) & 31u));
            const uint32_t tcn = idx - (trel ? prev : 0u);
            const uint32_t j0 = (tr_lo + trel) * (uint32_t)TL_R, c0 = tcn * (uint32_t)TL_C;
            const uint32_t slot = __shfl_sync(0xffffffffu, gp_excl, (int)trel) + tcn;
            Acc z;
            z.clear();
            bool any = false;
            if (j0 < prow && j0 + c0 < pd) {
                const uint32_t blo = j0 + 1u > nV ? j0 + 1u - nV : 0u, bhi = min(nS - 1u, j0 + (uint32_t)TL_PAD_ROWS);
                // x at aS + (b P + i) 8, y at aV + ((j0 + r - b) P + c0 + c - i) 8:  y = Cl - x
                const uint32_t Cl = aS + aV + ((j0 + r) * Pw + c0 + c) * 8u;
                // the scalar ranges of up to 32 rows at a time, one row per lane; then the non-empty rows one by one
                for (uint32_t bb = blo; bb <= bhi; bb += 32u) {
                    const uint32_t bl = bb + lane;
                    uint32_t pk = 0;
                    if (bl <= bhi) {
                        const uint32_t la = lds_u8(base + bl), wv = lds_u8(aWm + j0 - bl);
                        const uint32_t ilo = (c0 + 1u > wv ? c0 + 1u - wv : 0u) & ~1u, ihi = min(la, c0 + (uint32_t)TL_C);
                        if (ilo < ihi) pk = (aS + bl * P8 + ilo * 8u) | ((ihi - ilo + 1u) >> 1) << 24; // x address | step pairs
                    }
                    uint32_t m = __ballot_sync(0xffffffffu, pk != 0u);
                    any |= m != 0u;
                    while (m) {
                        const uint32_t q = (uint32_t)__ffs(m) - 1u;
                        m &= m - 1u;
                        const uint32_t v = __shfl_sync(0xffffffffu, pk, (int)q);
                        const uint32_t ax = v & 0x00ffffffu;
                        conv_run(z, ax, Cl - ax, v >> 24);
                    }
                }
            }
            if (any) {
                // add the item to the tile's accumulator under the tile's lock
                uint64_t o[LACC];
                z.limbs(o);
                uint64_t *ac = s_acc + (size_t)slot * LACC * 32u + lane;
                if (lane == 0) {
                    while (atomicCAS(&s_lock[slot], 0u, 1u) != 0u) __nanosleep(32);
                    __threadfence_block();
                }
                __syncwarp();
                uint64_t v[LACC];
#pragma unroll
                for (int l = 0; l < LACC; ++l) v[l] = ac[l * 32];
                cf_add<LACC, F64>(v, o);
#pragma unroll
                for (int l = 0; l < LACC; ++l) ac[l * 32] = v[l];
                __syncwarp();
                if (lane == 0) {
                    __threadfence_block();
                    atomicExch(&s_lock[slot], 0u);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[s]);
        if (++s == nstage) {
            s = 0;
            ph ^= 1u;
        }
        if (d0.z & 1u) {
            // ---- the unit's last pair: when every warp has left it, write the tiles and clear the accumulators ----
            asm volatile("bar.sync 1, %0;" ::"r"(NW * 32u) : "memory");
            const uint32_t nch = P.wo >> 5, orows = d0.w & 0xffu;
            const uint32_t rb = __ldg(P.row_base + d0.x);
            const uint64_t okd = __ldg(P.okey + d0.x) * P.delta;
            const uint32_t gin = gp_excl + (tr_lo + lane < tr_hi ? tl_row_tiles(d0.w, tr_lo + lane) : 0u); // inclusive
            for (uint32_t slot = warp; slot < gtiles; slot += NW) {
                const uint32_t trel = __popc(__ballot_sync(0xffffffffu, slot >= gin));
                const uint32_t j = (tr_lo + trel) * (uint32_t)TL_R + r;
                const uint32_t c0 = (slot - __shfl_sync(0xffffffffu, gp_excl, (int)trel)) * (uint32_t)TL_C, col = c0 + c;
                uint64_t *ac = s_acc + (size_t)slot * LACC * 32u + lane;
                uint64_t o[LACC];
#pragma unroll
                for (int l = 0; l < LACC; ++l) {
                    o[l] = ac[l * 32];
                    ac[l * 32] = 0;
                }
                // a term beyond the truncation limit is not part of the product: it is stored as zero
                if (P.tr.on && (int32_t)(P.tr.w1 * j + P.tr.w0 * col) > (int32_t)d1.z) {
#pragma unroll
                    for (int l = 0; l < LACC; ++l) o[l] = 0;
                }
                const bool inside = j < orows;
                const uint32_t row = rb + j;
                bool nz = false;
                if (inside && P.split) {
                    // another unit may hold the other pairs of this plane: add, 32 bits per 64-bit word
                    if constexpr (F64) {
                        if (__longlong_as_double((long long)o[0]) != 0.0)
                            atomicAdd(reinterpret_cast<double *>(P.omat + (size_t)row * P.wo + col), __longlong_as_double((long long)o[0]));
                    } else {
                        unsigned long long *om = reinterpret_cast<unsigned long long *>(P.omat + (size_t)row * (2 * LACC) * P.wo + col);
#pragma unroll
                        for (int l = 0; l < LACC; ++l) {
                            if (o[l] & 0xffffffffull) atomicAdd(om + (size_t)(2 * l) * P.wo, o[l] & 0xffffffffull);
                            if (o[l] >> 32) atomicAdd(om + (size_t)(2 * l + 1) * P.wo, o[l] >> 32);
                        }
                    }
                    if (c == 0u) P.or_hi[row] = okd + j;
                } else if (inside) {
                    uint64_t *om = P.omat + (size_t)row * LACC * P.wo + col;
#pragma unroll
                    for (int l = 0; l < LACC; ++l) {
                        om[(size_t)l * P.wo] = o[l];
                        nz |= o[l] != 0;
                    }
                    if constexpr (F64) nz = __longlong_as_double((long long)o[0]) != 0.0;
                }
                const uint32_t bal = __ballot_sync(0xffffffffu, nz);
                if (inside && c == 0u && !P.split) {
                    const uint32_t cnt = __popc((bal >> (8u * r)) & 0xffu);
                    if (cnt) atomicAdd(P.or_cnt + (size_t)row * nch + (c0 >> 5), cnt);
                    P.or_hi[row] = okd + j;
                }
            }
            asm volatile("bar.sync 1, %0;" ::"r"(NW * 32u) : "memory");
        }
    }
}

// Split mode: the 32-bit digits of every output entry -> limbs (in place, words 0 .. LACC-1 of the entry), and the
// non-zero count of every (row, 32-column chunk).  Two's-complement contributions add up modulo 2^(64 LACC).
template <int LACC, bool F64>
__global__ void __launch_bounds__(256) k_tile_norm(uint64_t *__restrict__ omat, uint32_t nitems, uint32_t nch, uint32_t wo,
                                                   uint32_t *__restrict__ or_cnt)
{
    constexpr int LS = F64 ? 1 : 2 * LACC;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t wpb = blockDim.x >> 5;
    for (uint32_t it = blockIdx.x * wpb + (threadIdx.x >> 5); it < nitems; it += gridDim.x * wpb) {
        const uint32_t row = it / nch, col = (it - row * nch) * 32u + lane;
        uint64_t *o = omat + (size_t)row * LS * wo + col;
        bool nz = false;
        if constexpr (F64) {
            nz = __longlong_as_double((long long)o[0]) != 0.0;
        } else {
            uint64_t dg[LS], carry = 0, any = 0;
#pragma unroll
            for (int d = 0; d < LS; ++d) {
                dg[d] = o[(size_t)d * wo];
                any |= dg[d];
            }
            if (any) {
                uint64_t v[LACC];
#pragma unroll
                for (int l = 0; l < LACC; ++l) {
                    const uint64_t t0 = dg[2 * l] + carry;                // < 2^64: digits are sums of < 2^32 values
                    const uint64_t t1 = dg[2 * l + 1] + (t0 >> 32);
                    v[l] = (t0 & 0xffffffffull) | (t1 << 32);
                    carry = t1 >> 32;
                    nz |= v[l] != 0;
                }
#pragma unroll
                for (int l = 0; l < LACC; ++l) o[(size_t)l * wo] = v[l];
            }
        }
        const uint32_t b = __ballot_sync(0xffffffffu, nz);
        if (lane == 0) or_cnt[it] = __popc(b);
    }
}

// Expand the output rows into terms: key = hi * delta + lo, zero coefficients dropped (terms beyond a truncation limit
// were stored as zeros); off[row * nch + chunk] is the exclusive scan of the per-chunk counts; `ls` = words per entry.
template <int LACC, bool F64>
__global__ void __launch_bounds__(256) k_tile_expand(const uint64_t *__restrict__ or_hi, const uint64_t *__restrict__ omat,
                                                     const uint32_t *__restrict__ off, uint32_t nitems, uint32_t nch, uint32_t wo,
                                                     uint32_t ls, uint64_t delta, uint64_t *__restrict__ okeys,
                                                     uint64_t *__restrict__ ocfs)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t wpb = blockDim.x >> 5;
    for (uint32_t it = blockIdx.x * wpb + (threadIdx.x >> 5); it < nitems; it += gridDim.x * wpb) {
        const uint32_t base = off[it], cnt = off[it + 1] - base;
        if (cnt == 0u) continue;
        const uint32_t row = it / nch, col = (it - row * nch) * 32u + lane;
        const uint64_t *o = omat + (size_t)row * ls * wo + col;
        uint64_t v[LACC];
        bool nz = false;
#pragma unroll
        for (int l = 0; l < LACC; ++l) {
            v[l] = o[(size_t)l * wo];
            nz |= v[l] != 0;
        }
        if constexpr (F64) nz = __longlong_as_double((long long)v[0]) != 0.0;
        const uint32_t b = __ballot_sync(0xffffffffu, nz);
        if (nz) {
            const uint32_t pos = base + __popc(b & ((1u << lane) - 1u));
            okeys[pos] = or_hi[row] * delta + (uint64_t)col;
#pragma unroll
            for (int l = 0; l < LACC; ++l) ocfs[(size_t)pos * LACC + l] = v[l];
        }
    }
}

} // namespace obk
