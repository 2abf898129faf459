// obk_planes.cuh -- plane-blocked dense path ("kernel 2") of the B200 series-product engine (sm_100a).
//
// Same algebra as the row path (obk_kernels.cuh, "kernel 1"), one level up.  A packed key is
//     key = lo0 + delta * lo1 + delta^2 * hi2            (kpack.hpp:262-267: the first symbol is the lowest digit)
// and because the exponent-overflow check (polynomial.hpp:839-851) guarantees that digit-wise sums never carry,
//     (hi2, lo1, lo0) * (hi2', lo1', lo0') = (hi2 + hi2', lo1 + lo1', lo0 + lo0').
// All terms sharing hi2 form a PLANE: a dense 32 x 32 tile of coefficients indexed [lo1][lo0].  The product of two
// planes is a 2-D convolution that lands in the single output plane hi2 + hi2'.  What the reference does with one
// hash-table probe per term product (polynomial.hpp:1467-1525) becomes, for dense operands:
//   * a symbolic product of the PLANE keys only (Fateman n=30: 496 x 496 plane pairs -> 1 891 output planes,
//     instead of 5 456 x 5 456 row pairs or 2.15e9 term pairs) that lists, per output plane, its plane pairs;
//   * one CTA per (output plane, block of output rows): a producer warp streams the plane pairs of the list through
//     a shared-memory ring with TMA bulk copies (cp.async.bulk + mbarrier complete_tx); every consumer warp owns ONE
//     output row j of the plane and, for each staged pair (A, B), convolves row b of A with row j - b of B for every
//     b -- no directory search, no per-pair global loads, no per-pair staging: the row pair is two shared-memory
//     addresses;
//   * exact multiply-accumulate in carry-save form: the 64 x 64 -> 128-bit product is four IMAD.WIDE.U32 whose 64-bit
//     results accumulate in three column accumulators (weights 2^0, 2^32, 2^64), the carry-outs are counted in three
//     32-bit counters by IADD3.X with two carry-in predicates each: 4 FMA-pipe + 2 ALU-pipe instructions per product
//     (the 128-bit multiply + 192-bit add of the row path is 12), normalised to limbs once per output row;
//   * short rows share a warp: output rows of at most 16 (8) entries are convolved two (four) row pairs at a time,
//     one per half (quarter) warp, and the partial rows are folded with shuffles when the output row is written.
// No atomics, no locks, no table in the product kernel.
#pragma once
#include "obk_kernels.cuh"

namespace obk
{

constexpr int PL_ROWS = 32;                 // rows (lo1) and row width (lo0) of a plane
constexpr int PL_PITCH = 64;                // words per staged row: [0,32) zeros, [32,64) data + zero fill
constexpr int PL_HDR = 16;                  // words of the row-length table (32 x u32) in front of the rows
constexpr int PL_BLOB = PL_HDR + PL_ROWS * PL_PITCH; // words of one plane in HBM, already in its staged layout
constexpr int PL_MAX_STAGES = 8;

// A plane lives in HBM exactly as it is staged in shared memory -- the row-length table followed by its rows at pitch
// 64 with the zero halves in place -- so that ONE bulk copy of (16 + rows * 64) words stages it.  (The first version
// copied row by row: ~40 small cp.async.bulk per plane pair took 2.2 us per pair and starved the consumers.)
// Shared-memory geometry of one ring stage: plane A with ra_max rows and one trailing zero row, plane B likewise,
// the descriptor.
__host__ __device__ inline uint32_t pl_plane_words(uint32_t rows_max)
{
    return PL_HDR + (rows_max + 1u) * PL_PITCH;
}
// onrows encoding: rows | 0x100 when the plane has rows longer than 32 entries.  Warps are assigned to VIRTUAL rows:
// first the low halves (outputs 0..31) of all rows, then, for such planes, the high halves (outputs 32..63).
__host__ __device__ inline uint32_t pl_virtual_rows(uint32_t enc)
{
    return (enc & 0xffu) * ((enc & 0x100u) ? 2u : 1u);
}
__host__ __device__ inline uint32_t pl_stage_bytes(uint32_t ra_max, uint32_t rb_max)
{
    return (pl_plane_words(ra_max) + pl_plane_words(rb_max)) * 8u + 128u;
}

// ---- plane building ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t plane_hash(uint64_t key, uint32_t log2_cap)
{
    return (uint32_t)((key * 0x9E3779B97F4A7C15ULL) >> (64 - log2_cap));
}
// owner rank of a plane key (multi-GPU): a function of the key alone, so every rank takes the same decision whatever
// numbers its own tables gave the planes
__host__ __device__ inline uint32_t plane_owner(uint64_t key, uint32_t nranks)
{
    return (uint32_t)(((key * 0xD6E8FEB86659FD93ULL) >> 33) % nranks);
}

// The set of plane keys hi2 = key / delta^2 of an operand (open addressing, global memory).  The thread that claims a
// slot numbers the plane: no flag / scan / gather passes (the row path spends five launches on them).  Plane numbers
// therefore depend on the race; nothing downstream may depend on them beyond "distinct and dense".
__global__ void __launch_bounds__(256) k_plane_insert(const uint64_t *__restrict__ keys, uint64_t n, uint64_t d2, uint64_t *__restrict__ tab,
                                                      uint32_t log2_cap, uint32_t *__restrict__ slot_id, uint64_t *__restrict__ plane_key,
                                                      uint32_t *__restrict__ counter)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint32_t mask = (1u << log2_cap) - 1u;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t hi = keys[i] / d2;
        uint32_t h = plane_hash(hi, log2_cap);
        for (;;) {
            const unsigned long long old = atomicCAS((unsigned long long *)&tab[h], (unsigned long long)ROW_EMPTY, (unsigned long long)hi);
            if (old == ROW_EMPTY) {
                const uint32_t id = atomicAdd(counter, 1u);
                slot_id[h] = id;
                plane_key[id] = hi;
                break;
            }
            if (old == hi) break;
            h = (h + 1u) & mask;
        }
    }
}

// Scatter every term's coefficient into its plane's blob and record row lengths / row counts.
__global__ void __launch_bounds__(256) k_plane_fill(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ cfs, uint64_t n,
                                                    uint64_t delta, const uint64_t *__restrict__ tab, uint32_t log2_cap,
                                                    const uint32_t *__restrict__ slot_id, uint64_t *__restrict__ mat,
                                                    uint32_t *__restrict__ pinfo)
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
        uint64_t *blob = mat + (size_t)p * PL_BLOB;
        blob[PL_HDR + lo1 * PL_PITCH + 32 + lo0] = cfs[i];
        atomicMax(reinterpret_cast<uint32_t *>(blob) + lo1, lo0 + 1u);
        atomicMax(&pinfo[4 * (size_t)p], lo1 + 1u);     // rows of the plane
        atomicAdd(&pinfo[4 * (size_t)p + 1], 1u);       // terms of the plane
        atomicMax(&pinfo[4 * (size_t)p + 2], lo0 + 1u); // longest row
    }
}

// ---- symbolic plane product ----------------------------------------------------------------------------------------
// Per-slot records of the output-plane set: [0] pairs, [1] rows, [2] long-row flag (0x100), [3] output plane number.
// Pass 1: every plane pair's key sum is inserted; the claiming thread numbers the output plane.  part_nranks > 1: this
// rank only forms the pairs of the left planes it owns (multi-GPU partial product).
__global__ void __launch_bounds__(256) k_plane_sym_insert(const uint64_t *__restrict__ xh, const uint32_t *__restrict__ xinfo, uint32_t nx,
                                                          const uint64_t *__restrict__ yh, const uint32_t *__restrict__ yinfo, uint32_t ny,
                                                          uint32_t part_rank, uint32_t part_nranks, uint64_t *__restrict__ tab,
                                                          uint32_t log2_cap, uint4 *__restrict__ rec, unsigned long long *__restrict__ work,
                                                          uint64_t *__restrict__ okey, uint32_t *__restrict__ counters)
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
        uint32_t *r = reinterpret_cast<uint32_t *>(&rec[h]);
        atomicAdd(&r[0], 1u);
        atomicAdd(&work[h], (unsigned long long)xinfo[4 * a + 1] * yinfo[4 * b + 1]);
        atomicMax(&r[1], xinfo[4 * a] + yinfo[4 * b] - 1u);
        // some output row is longer than 32 entries: its high half gets a warp of its own
        if (xinfo[4 * a + 2] + yinfo[4 * b + 2] - 1u > 32u) atomicOr(&r[2], 0x100u);
    }
}

// Pass 2: the per-slot records by output plane number.
__global__ void __launch_bounds__(256) k_plane_sym_compact(const uint64_t *__restrict__ tab, uint32_t cap, const uint4 *__restrict__ rec,
                                                           const unsigned long long *__restrict__ work, uint32_t *__restrict__ ocnt,
                                                           unsigned long long *__restrict__ owork, uint32_t *__restrict__ onrows)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap || tab[i] == ROW_EMPTY) return;
    const uint4 r = rec[i];
    ocnt[r.w] = r.x;
    owork[r.w] = work[i];
    onrows[r.w] = r.y | r.z;
}

// Pass 3 (one CTA): pair_off = exclusive scan of the pairs per plane, row_base = exclusive scan of the rows per plane,
// and the work units = (output plane, block of `nw` virtual rows), heaviest planes first (work classes by bit length:
// a counting sort).  own_nranks > 1: only the units this rank owns (multi-GPU owner mode; ownership is a function of
// the plane key and the block, the same on every rank).
//   units[u] = plane | block << 24     (planes < 2^24, blocks < 2^8)
//   totals[0] = units, totals[1] = output rows, totals[2] = output planes
__global__ void __launch_bounds__(1024) k_plane_units(const unsigned long long *__restrict__ owork, const uint32_t *__restrict__ onrows,
                                                      const uint32_t *__restrict__ ocnt, const uint64_t *__restrict__ okey,
                                                      const uint32_t *__restrict__ counters, uint32_t nw, uint32_t own_rank,
                                                      uint32_t own_nranks, uint32_t *__restrict__ row_base,
                                                      uint32_t *__restrict__ pair_off, uint32_t *__restrict__ units,
                                                      uint32_t *__restrict__ totals)
{
    __shared__ uint32_t s_cls[65], s_tot, s_carry;
    const uint32_t tid = threadIdx.x;
    const uint32_t nout = counters[0];
    auto owned = [&](uint32_t o, uint32_t b) { return own_nranks <= 1u || (plane_owner(okey[o], own_nranks) + b) % own_nranks == own_rank; };
    if (tid < 65) s_cls[tid] = 0;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    // units per work class
    for (uint32_t o = tid; o < nout; o += blockDim.x) {
        const uint32_t c = 64u - (uint32_t)__clzll((long long)(owork[o] | 1ull));
        const uint32_t nb = (pl_virtual_rows(onrows[o]) + nw - 1u) / nw;
        uint32_t mine = 0;
        for (uint32_t b = 0; b < nb; ++b) mine += owned(o, b) ? 1u : 0u;
        if (mine) atomicAdd(&s_cls[64u - c], mine); // class 0 = heaviest
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t acc = 0;
        for (int c = 0; c < 65; ++c) {
            const uint32_t v = s_cls[c];
            s_cls[c] = acc;
            acc += v;
        }
        totals[0] = acc;
        totals[2] = nout;
    }
    __syncthreads();
    // unit slots (the order inside a class does not matter)
    for (uint32_t o = tid; o < nout; o += blockDim.x) {
        const uint32_t c = 64u - (uint32_t)__clzll((long long)(owork[o] | 1ull));
        const uint32_t nb = (pl_virtual_rows(onrows[o]) + nw - 1u) / nw;
        uint32_t mine = 0;
        for (uint32_t b = 0; b < nb; ++b) mine += owned(o, b) ? 1u : 0u;
        if (mine == 0u) continue;
        uint32_t at = atomicAdd(&s_cls[64u - c], mine);
        for (uint32_t b = 0; b < nb; ++b)
            if (owned(o, b)) units[at++] = o | (b << 24);
    }
    // the two scans
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
                v[j] = base + j < nout ? (which == 0 ? (onrows[base + j] & 0xffu) : ocnt[base + j]) : 0;
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

// Pass 4: the pair list, grouped by output plane.
__global__ void __launch_bounds__(256) k_plane_sym_fill(const uint64_t *__restrict__ xh, uint32_t nx, const uint64_t *__restrict__ yh, uint32_t ny,
                                                        uint32_t part_rank, uint32_t part_nranks, const uint64_t *__restrict__ tab,
                                                        uint32_t log2_cap, const uint4 *__restrict__ rec,
                                                        const uint32_t *__restrict__ pair_off, uint32_t *__restrict__ fill,
                                                        uint2 *__restrict__ pairs)
{
    const uint64_t total = (uint64_t)nx * ny, stride = (uint64_t)gridDim.x * blockDim.x;
    const uint32_t mask = (1u << log2_cap) - 1u;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const uint32_t a = (uint32_t)(i / ny), b = (uint32_t)(i - (uint64_t)a * ny);
        const uint64_t ka = xh[a];
        if (part_nranks > 1u && plane_owner(ka, part_nranks) != part_rank) continue;
        const uint64_t key = ka + yh[b];
        uint32_t h = plane_hash(key, log2_cap);
        while (tab[h] != key) h = (h + 1u) & mask;
        const uint32_t o = rec[h].w;
        const uint32_t pos = atomicAdd(&fill[o], 1u);
        pairs[(size_t)pair_off[o] + pos] = make_uint2(a, b);
    }
}

// ---- exact accumulators --------------------------------------------------------------------------------------------
// One accumulator per lane and output position.  KIND 0: unsigned carry-save columns (no negative coefficient, 2 or 3
// limbs), 1: signed two's-complement limbs (128-bit multiply + multi-limb add, as the row path), 2: one limb (mod 2^64
// arithmetic is exact when the a-priori bound says the result fits), 3: fp64 with two roundings (fma3.hpp:65-72).
template <int LACC, bool F64, bool UNS>
struct PlaneAcc {
    static constexpr int KIND = F64 ? 3 : (LACC == 1 ? 2 : (UNS ? 0 : 1));
    uint64_t A, B, B2, C; // KIND 0: column accumulators, weights 2^0, 2^32 (x0 y1), 2^32 (x1 y0), 2^64
    uint32_t ca, cb, cc;  // KIND 0: carry-outs of the columns (weights 2^64, 2^96, 2^128)
    uint64_t z[3];        // KIND 1, 2
    double d;             // KIND 3

    __device__ __forceinline__ void clear()
    {
        A = B = B2 = C = 0;
        ca = cb = cc = 0;
        z[0] = z[1] = z[2] = 0;
        d = 0.0;
    }
    __device__ __forceinline__ void mac(uint64_t x, uint64_t y)
    {
        if constexpr (KIND == 0) {
            // ptxas fuses every mad.lo.cc / madc.hi.cc pair into ONE IMAD.WIDE.U32 with a carry-out predicate and
            // the four addc into two IADD3.X with two carry-in predicates each: 4 FMA-pipe + 2 ALU-pipe instructions
            // per product.  The accumulators travel as 64-bit operands so that they sit in aligned register pairs and
            // no MOV is needed around the IMAD.WIDE.  The two middle products have a column each (B, B2): the four
            // IMAD.WIDE of a product are independent, and the dependent chain through a column is one per product.
            asm("{\n\t.reg .u32 x0, x1, y0, y1, l, h;\n\t"
                "mov.b64 {x0, x1}, %7;\n\tmov.b64 {y0, y1}, %8;\n\t"
                "mov.b64 {l, h}, %0;\n\t"
                "mad.lo.cc.u32  l, x0, y0, l;\n\tmadc.hi.cc.u32 h, x0, y0, h;\n\taddc.u32 %4, %4, 0;\n\t"
                "mov.b64 %0, {l, h};\n\t"
                "mov.b64 {l, h}, %1;\n\t"
                "mad.lo.cc.u32  l, x0, y1, l;\n\tmadc.hi.cc.u32 h, x0, y1, h;\n\taddc.u32 %5, %5, 0;\n\t"
                "mov.b64 %1, {l, h};\n\t"
                "mov.b64 {l, h}, %2;\n\t"
                "mad.lo.cc.u32  l, x1, y0, l;\n\tmadc.hi.cc.u32 h, x1, y0, h;\n\taddc.u32 %5, %5, 0;\n\t"
                "mov.b64 %2, {l, h};\n\t"
                "mov.b64 {l, h}, %3;\n\t"
                "mad.lo.cc.u32  l, x1, y1, l;\n\tmadc.hi.cc.u32 h, x1, y1, h;\n\taddc.u32 %6, %6, 0;\n\t"
                "mov.b64 %3, {l, h};\n\t}"
                : "+l"(A), "+l"(B), "+l"(B2), "+l"(C), "+r"(ca), "+r"(cb), "+r"(cc)
                : "l"(x), "l"(y));
        } else if constexpr (KIND == 1) {
            uint64_t(&zz)[LACC] = reinterpret_cast<uint64_t(&)[LACC]>(z);
            row_mac<LACC, false>(zz, x, y);
        } else if constexpr (KIND == 2) {
            z[0] += x * y;
        } else {
            d = __dadd_rn(d, __dmul_rn(__longlong_as_double((long long)x), __longlong_as_double((long long)y)));
        }
    }
    // the value as LACC two's-complement limbs (fp64: the bit pattern)
    __device__ __forceinline__ void limbs(uint64_t (&o)[LACC]) const
    {
        if constexpr (KIND == 0) {
            const unsigned __int128 lo = (unsigned __int128)A + ((unsigned __int128)B << 32) + ((unsigned __int128)B2 << 32);
            const unsigned __int128 mid = (lo >> 64) + (unsigned __int128)C + ca + ((unsigned __int128)cb << 32);
            o[0] = (uint64_t)lo;
            o[1] = (uint64_t)mid;
            if constexpr (LACC >= 3) o[2] = (uint64_t)(mid >> 64) + cc;
        } else if constexpr (KIND == 3) {
            o[0] = (uint64_t)__double_as_longlong(d);
        } else {
#pragma unroll
            for (int l = 0; l < LACC; ++l) o[l] = z[l];
        }
    }
};

// o += v in the coefficient ring (exact limbs, or fp64)
template <int LACC, bool F64>
__device__ __forceinline__ void cf_add(uint64_t (&o)[LACC], const uint64_t (&v)[LACC])
{
    if constexpr (F64) {
        o[0] = (uint64_t)__double_as_longlong(__dadd_rn(__longlong_as_double((long long)o[0]), __longlong_as_double((long long)v[0])));
    } else {
        limbs_add<LACC>(o, v);
    }
}

struct PlaneParams {
    const uint64_t *xmat, *ymat;   // [planes][PL_BLOB]: row lengths, then the rows in their staged layout
    const uint32_t *xinfo, *yinfo; // [planes][4]: rows, terms, longest row, -
    const uint2 *pairs;            // plane pairs grouped by output plane
    const uint32_t *pair_off;      // [nout + 1]
    const uint64_t *okey;          // [nout] output plane keys (hi2 sums)
    const uint32_t *onrows;        // [nout] rows | 0x100 (see pl_virtual_rows)
    const uint32_t *row_base;      // [nout + 1]
    const uint32_t *units;         // plane | block << 24, heaviest first
    uint32_t nunits;
    uint32_t *cursor;
    uint64_t delta;
    uint64_t *or_hi;  // [rows] row key = okey * delta + j
    uint64_t *omat;   // [rows][LACC][64]
    uint32_t *or_cnt; // [rows][2] non-zero entries of the low / high half
    uint32_t ra_max, rb_max; // most rows of a left / right plane (stage geometry)
    uint32_t nstage;         // ring depth (<= PL_MAX_STAGES)
};

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// descriptor of a staged plane pair (written by the producer before it arms the full barrier)
struct PlaneDesc {
    uint32_t plane;  // output plane
    uint32_t v0;     // first virtual row of this CTA's block
    uint32_t flags;  // 1: last pair of the unit (write the rows), 2: quit
    uint32_t rows;   // onrows[plane]
};

// Shared-memory loads by 32-bit shared-window address.  (With generic pointers the compiler re-derived the window base
// -- S2R SR_CgaCtaId, LEA -- at every use: a quarter of the per-pair bookkeeping.)  asm volatile keeps them between the
// mbarrier wait and the mbarrier arrive of their stage.
__device__ __forceinline__ uint64_t lds_u64(uint32_t a)
{
    uint64_t v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void lds_2u64(uint32_t a, uint64_t &v0, uint64_t &v1)
{
    asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(v0), "=l"(v1) : "r"(a));
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint4 lds_u32x4(uint32_t a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}

// steps i0, i0+2, ... < n of one row pair: z[lane] += x[i] * y[lane - i]; px / py are the shared addresses of x[0]
// and of y[lane].  Steps run in pairs (one LDS.128 for two scalars): a scalar row is zero beyond its length and so is
// the pad after it, an odd count is rounded up for free.
template <class Acc>
__device__ __forceinline__ void conv_steps(Acc &z, uint32_t px, uint32_t py, uint32_t i0, uint32_t n)
{
#pragma unroll 2
    for (uint32_t i = i0; i < n; i += 2) {
        uint64_t x0, x1;
        lds_2u64(px + 8u * i, x0, x1);
        const uint64_t y0 = lds_u64(py - 8u * i), y1 = lds_u64(py - 8u * i - 8u);
        z.mac(x0, y0);
        z.mac(x1, y1);
    }
}

// NW consumer warps (= virtual output rows per unit) + one producer warp.
template <int LACC, bool F64, bool UNS, int NW, int MINB>
__global__ void __launch_bounds__((NW + 1) * 32, MINB) k_planeconv(const PlaneParams P)
{
    using Acc = PlaneAcc<LACC, F64, UNS>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_full[PL_MAX_STAGES], s_empty[PL_MAX_STAGES];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t nstage = P.nstage;
    const uint32_t wordsA = pl_plane_words(P.ra_max), wordsB = pl_plane_words(P.rb_max);
    const uint32_t stage_bytes = pl_stage_bytes(P.ra_max, P.rb_max);
    // zero everything once (the trailing zero rows are never written again)
    for (uint32_t i = tid; i < nstage * stage_bytes / 16u; i += blockDim.x)
        reinterpret_cast<uint4 *>(smem_raw)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        for (uint32_t s = 0; s < nstage; ++s) {
            mbar_init(&s_full[s], 1);
            mbar_init(&s_empty[s], NW);
        }
    }
    __syncthreads();
    // plane blobs of a stage; the rows start PL_HDR words into a blob
    auto blobA = [&](uint32_t s) { return reinterpret_cast<uint64_t *>(smem_raw + (size_t)s * stage_bytes); };
    auto blobB = [&](uint32_t s) { return blobA(s) + wordsA; };
    auto stageDesc = [&](uint32_t s) { return reinterpret_cast<PlaneDesc *>(blobA(s) + wordsA + wordsB); };

    if (warp == NW) {
        // ================= producer: stream the plane pairs of this CTA's units through the ring =================
        uint32_t s = 0, ph = 0; // ring slot and the parity of its current use
        for (;;) {
            uint32_t u = 0;
            if (lane == 0) u = atomicAdd(P.cursor, 1u);
            u = __shfl_sync(0xffffffffu, u, 0);
            const uint64_t uidx = u;
            const bool quit = uidx >= P.nunits;
            uint32_t o = 0, v0 = 0, p0 = 0, p1 = 1, enc = 0, jmin = 0;
            if (!quit) {
                const uint32_t ud = __ldg(P.units + uidx);
                o = ud & 0x00ffffffu;
                v0 = (ud >> 24) * NW;
                p0 = __ldg(P.pair_off + o);
                p1 = __ldg(P.pair_off + o + 1);
                enc = __ldg(P.onrows + o);
                // lowest output row any warp of this block works on
                const uint32_t R = enc & 0xffu;
                jmin = v0 < R ? (v0 + NW > R && (enc & 0x100u) ? 0u : v0) : v0 - R;
            }
            for (uint32_t pb = p0; pb < p1; pb += 32u) {
                // 32 pairs' worth of descriptors at a time
                uint2 pr = make_uint2(0, 0);
                uint32_t na = 0, nb = 0;
                if (!quit && pb + lane < p1) {
                    pr = __ldg(P.pairs + pb + lane);
                    na = __ldg(P.xinfo + 4 * (size_t)pr.x);
                    nb = __ldg(P.yinfo + 4 * (size_t)pr.y);
                }
                const uint32_t cntp = quit ? 1u : min(32u, p1 - pb);
                for (uint32_t q = 0; q < cntp; ++q) {
                    const uint32_t px = __shfl_sync(0xffffffffu, pr.x, (int)q), py = __shfl_sync(0xffffffffu, pr.y, (int)q);
                    const uint32_t ra = __shfl_sync(0xffffffffu, na, (int)q), rb = __shfl_sync(0xffffffffu, nb, (int)q);
                    const bool last = pb + q + 1u == p1;
                    // a pair whose rows cannot reach this block of output rows is not staged (unless it must carry the
                    // end-of-unit flag)
                    if (!quit && !last && ra + rb - 1u <= jmin) continue;
                    mbar_wait(&s_empty[s], ph ^ 1u); // first pass over the ring: passes immediately
                    if (lane == 0) {
                        PlaneDesc *d = stageDesc(s);
                        d->plane = o;
                        d->v0 = v0;
                        d->flags = quit ? 2u : (last ? 1u : 0u);
                        d->rows = enc;
                        mbar_expect_tx(&s_full[s], quit ? 0u : (2u * PL_HDR + (ra + rb) * PL_PITCH) * 8u);
                        if (!quit) {
                            tma_bulk_g2s(blobA(s), P.xmat + (size_t)px * PL_BLOB, (PL_HDR + ra * PL_PITCH) * 8u, &s_full[s]);
                            tma_bulk_g2s(blobB(s), P.ymat + (size_t)py * PL_BLOB, (PL_HDR + rb * PL_PITCH) * 8u, &s_full[s]);
                        }
                    }
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

    // ================= consumers: one warp = one half (32 outputs) of one output row of the unit's block =========
    Acc z, zs;           // whole-warp accumulator (output = lane); sub-warp partial rows (output = lane % gsub)
    uint64_t base[LACC]; // folded sub-warp partials, output = lane
    uint32_t gsub = 0;   // group width the sub-warp accumulator currently holds (0: empty)
    z.clear();
    zs.clear();
#pragma unroll
    for (int l = 0; l < LACC; ++l) base[l] = 0;

    auto fold_sub = [&]() {
        if (gsub == 0) return;
        uint64_t t[LACC];
        zs.limbs(t);
        for (uint32_t o = 16; o >= gsub; o >>= 1) {
            uint64_t v[LACC];
#pragma unroll
            for (int l = 0; l < LACC; ++l) v[l] = shfl_u64(t[l], lane ^ o);
            cf_add<LACC, F64>(t, v);
        }
        if (lane < gsub) cf_add<LACC, F64>(base, t);
        zs.clear();
        gsub = 0;
    };

    const uint32_t smem0 = smem_u32(smem_raw);
    uint32_t s = 0, ph = 0;
    for (;;) {
        mbar_wait(&s_full[s], ph);
        const uint32_t aA = smem0 + s * stage_bytes, aB = aA + wordsA * 8u; // shared addresses of the two blobs
        const uint4 dd = lds_u32x4(aB + wordsB * 8u);
        PlaneDesc d;
        d.plane = dd.x;
        d.v0 = dd.y;
        d.flags = dd.z;
        d.rows = dd.w;
        if (d.flags & 2u) break;
        const uint32_t R = d.rows & 0xffu, V = pl_virtual_rows(d.rows);
        const uint32_t v = d.v0 + warp;
        const uint32_t h = v >= R ? 1u : 0u;   // 1: the high half (outputs 32..63) of the row
        const uint32_t j = h ? v - R : v;       // my output row inside the plane
        if (v < V) {
            // row pairs (b, j - b): lane b looks at its candidate
            uint32_t la = lds_u32(aA + 4u * lane);
            const uint32_t jb = j - lane;
            uint32_t lb = jb < (uint32_t)PL_ROWS ? lds_u32(aB + 4u * jb) : 0u;
            // the high half only sees pairs whose output reaches beyond 32 entries
            const bool valid = la != 0u && lb != 0u && (h == 0u || la + lb > 33u);
            if (!valid) la = lb = 0u;
            const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
            if (vmask) {
                const uint32_t L = __reduce_max_sync(0xffffffffu, valid ? la + lb - 1u : 0u);
                // shared addresses of A[0][0] and of B[j][lane (+32)]: row b of A is + b * 512, row j - b of B is - b * 512
                const uint32_t rowA0 = aA + (PL_HDR + 32u) * 8u, rowBj = aB + (PL_HDR + 32u) * 8u + j * (PL_PITCH * 8u);
                if (L > 16u) {
                    // ---- one row pair per trip, the whole warp on one pair; the SHORTER row supplies the scalars.
                    // First the pairs whose B row is the shorter one, then the others ----
                    const uint32_t bmask = __ballot_sync(0xffffffffu, valid && lb < la);
                    const uint32_t voff = (lane + 32u * h) * 8u;
                    {
                        uint32_t m = bmask; // B row shorter: scalars from B, vector = A row
                        while (m) {
                            const uint32_t b = (uint32_t)__ffs(m) - 1u;
                            m &= m - 1u;
                            const uint32_t n = __shfl_sync(0xffffffffu, lb, (int)b), ly = __shfl_sync(0xffffffffu, la, (int)b);
                            const uint32_t i0 = h ? ((33u - ly) & ~1u) : 0u; // outputs reach the high half from step 33 - ly on
                            conv_steps(z, rowBj - b * (PL_PITCH * 8u), rowA0 + b * (PL_PITCH * 8u) + voff, i0, n);
                        }
                    }
                    {
                        uint32_t m = vmask & ~bmask; // scalars from A, vector = B row
                        while (m) {
                            const uint32_t b = (uint32_t)__ffs(m) - 1u;
                            m &= m - 1u;
                            const uint32_t n = __shfl_sync(0xffffffffu, la, (int)b), ly = __shfl_sync(0xffffffffu, lb, (int)b);
                            const uint32_t i0 = h ? ((33u - ly) & ~1u) : 0u;
                            conv_steps(z, rowA0 + b * (PL_PITCH * 8u), rowBj - b * (PL_PITCH * 8u) + voff, i0, n);
                        }
                    }
                } else {
                    // ---- short output row: 2 (L <= 16) or 4 (L <= 8) row pairs per trip, one per half / quarter warp ----
                    const uint32_t blo = (uint32_t)__ffs(vmask) - 1u, bhi = 31u - (uint32_t)__clz(vmask);
                    const uint32_t G = L <= 8u ? 8u : 16u, per = 32u / G;
                    if (gsub != G) {
                        fold_sub();
                        gsub = G;
                    }
                    const uint32_t g = lane / G, lg = lane % G;
                    for (uint32_t b0 = blo; b0 <= bhi; b0 += per) {
                        const uint32_t b = b0 + g;
                        const uint32_t a1 = __shfl_sync(0xffffffffu, la, (int)(b & 31u)), b1 = __shfl_sync(0xffffffffu, lb, (int)(b & 31u));
                        const bool live = b <= bhi && a1 != 0u;
                        const uint32_t ra = rowA0 + b * (PL_PITCH * 8u), rb = rowBj - b * (PL_PITCH * 8u);
                        const bool a_short = a1 <= b1;
                        // a dead group multiplies the zero run in front of A's row 0
                        const uint32_t px = live ? (a_short ? ra : rb) : rowA0 - 32u * 8u;
                        const uint32_t py = (live ? (a_short ? rb : ra) : rowA0) + lg * 8u;
                        const uint32_t n = live ? (a_short ? a1 : b1) : 0u;
                        const uint32_t trips = __reduce_max_sync(0xffffffffu, n);
                        conv_steps(zs, px, py, 0u, trips);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[s]);
        if (++s == nstage) {
            s = 0;
            ph ^= 1u;
        }
        if (d.flags & 1u) {
            // ---- the unit is complete: write my half row once, count its non-zero entries ----
            fold_sub();
            if (v < V) {
                const uint32_t row = __ldg(P.row_base + d.plane) + j;
                uint64_t o[LACC];
                z.limbs(o);
                cf_add<LACC, F64>(o, base);
                uint64_t *om = P.omat + (size_t)row * 64 * LACC + 32u * h + lane;
                bool nz = false;
#pragma unroll
                for (int l = 0; l < LACC; ++l) {
                    om[(size_t)l * 64] = o[l];
                    nz |= o[l] != 0;
                    // a plane without long rows has no warp for the high halves: they are all zero
                    if (!(d.rows & 0x100u)) om[(size_t)l * 64 + 32] = 0;
                }
                if constexpr (F64) nz = __longlong_as_double((long long)o[0]) != 0.0;
                const uint32_t c = __popc(__ballot_sync(0xffffffffu, nz));
                if (lane == 0) {
                    P.or_cnt[2 * (size_t)row + h] = c;
                    P.or_hi[row] = __ldg(P.okey + d.plane) * P.delta + j; // both halves: in owner mode they may sit on different ranks
                }
            }
            z.clear();
#pragma unroll
            for (int l = 0; l < LACC; ++l) base[l] = 0;
        }
    }
}

// Expand the output rows into terms: key = hi * delta + lo, zero coefficients dropped; off[2 * row + half] is the
// exclusive scan of the per-half counts.
template <int LACC, bool F64>
__global__ void __launch_bounds__(256) k_plane_expand(const uint64_t *__restrict__ or_hi, const uint64_t *__restrict__ omat,
                                                      const uint32_t *__restrict__ off, uint32_t nro, uint64_t delta,
                                                      uint64_t *__restrict__ okeys, uint64_t *__restrict__ ocfs)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t wpb = blockDim.x >> 5;
    for (uint32_t hr = blockIdx.x * wpb + (threadIdx.x >> 5); hr < 2u * nro; hr += gridDim.x * wpb) {
        const uint32_t base = off[hr], cnt = off[hr + 1] - base;
        if (cnt == 0u) continue;
        const uint32_t row = hr >> 1, half = hr & 1u;
        const uint64_t *o = omat + (size_t)row * 64 * LACC + half * 32u + lane;
        uint64_t v[LACC];
        bool nz = false;
#pragma unroll
        for (int l = 0; l < LACC; ++l) {
            v[l] = o[(size_t)l * 64];
            nz |= v[l] != 0;
        }
        if constexpr (F64) nz = __longlong_as_double((long long)v[0]) != 0.0;
        const uint32_t b = __ballot_sync(0xffffffffu, nz);
        if (nz) {
            const uint32_t pos = base + __popc(b & ((1u << lane) - 1u));
            okeys[pos] = or_hi[row] * delta + (uint64_t)(half * 32u + lane);
#pragma unroll
            for (int l = 0; l < LACC; ++l) ocfs[(size_t)pos * LACC + l] = v[l];
        }
    }
}

// ---- micro-benchmark: the exact multiply-accumulate rate of the chip -------------------------------------------------
// Register-resident z += x * y (64 x 64 -> carry-save 160 bits, the plane kernel's accumulator) with four independent
// accumulators per thread and nothing else in the loop: the compute roofline the dense kernels are reported against
// (bench.py, roofline.int_mac).  out[thread] keeps the result alive.
__global__ void __launch_bounds__(256) k_bench_int_mac(const uint64_t *__restrict__ in, uint32_t iters, uint64_t *__restrict__ out)
{
    PlaneAcc<3, false, true> a0, a1, a2, a3;
    a0.clear();
    a1.clear();
    a2.clear();
    a3.clear();
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t x0 = in[t & 255u], x1 = in[(t + 1u) & 255u], y0 = in[(t + 2u) & 255u], y1 = in[(t + 3u) & 255u];
#pragma unroll 4
    for (uint32_t i = 0; i < iters; ++i) {
        a0.mac(x0, y0);
        a1.mac(x1, y0);
        a2.mac(x0, y1);
        a3.mac(x1, y1);
    }
    uint64_t l0[3], l1[3], l2[3], l3[3];
    a0.limbs(l0);
    a1.limbs(l1);
    a2.limbs(l2);
    a3.limbs(l3);
    out[t] = l0[0] ^ l1[1] ^ l2[2] ^ l3[0] ^ l0[2] ^ l1[0] ^ l2[1] ^ l3[2];
}

} // namespace obk
