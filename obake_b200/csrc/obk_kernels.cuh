// obk_kernels.cuh -- device code of the B200 series-product engine (sm_100a).
//
// What the reference does on CPU threads (include/obake/polynomials/polynomial.hpp:797-1582,
// poly_mul_impl_mt_hm) is re-designed here around one idea it shares with the reference: the hash of a
// Kronecker-packed monomial is the key itself (packed_monomial.hpp:172-176), monomial multiplication
// is integer addition (:249-262), hence bucket(a*b) = (bucket(a)+bucket(b)) mod nsegs
// (polynomial.hpp:1441-1450) and every output segment can be owned by exactly one CTA.
//
// Kernels in this file
//   k_term_stats   per-variable exponent min/max (overflow check), (partial) degrees, coefficient width
//   k_bin_count / k_scan_* / k_bin_scatter   counting sort of the right operand by (bucket, degree)
//   k_deg_hist / k_count_mults               tot_n_mults of a truncated product
//   k_mul          the product: one CTA owns one output segment's open-addressing table in shared memory; term
//                  products are expanded warp-cooperatively and every segment is written straight to the final
//                  arrays (also: symbolic key-only products, the owner-side merge of the multi-GPU path)
//   k_rowconv      row-blocked dense path of round 1 (option kernel=1): one warp owns one output row, register
//                  accumulation (k_row_insert / k_row_flags / k_row_keys / k_row_fill / k_row_pack build the rows,
//                  k_compact orders the symbolic row product, k_row_expand turns rows back into terms).  The dense
//                  path in use is the tile-blocked kernel of obk_tiles.cuh (k_tileconv, kernel 3) on top of the plane
//                  sets, exact accumulators and symbolic plane product of obk_planes.cuh (k_planeconv, kernel 2)
//   k_push         multi-GPU: partial terms written straight into the owners' NVLink peer windows
//   k_widen_cfs / k_flag_keep / k_select / k_extend_symbols   series add/sub, degree truncation, symbol extension
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace obk
{

// ------------------------------------------------------------------------------------------------
// Key layout (Kronecker packing; kpack.hpp:219-375, tables src/kpack.cpp)
// ------------------------------------------------------------------------------------------------
struct Layout {
    uint32_t nvars, psize, W, is_signed;
    uint32_t wsize; // exponents per word: psize, or nvars for packed_monomial
    uint32_t pad;
    uint64_t delta;    // kpack delta for wsize components
    uint64_t klim_min; // two's-complement bit pattern of the smallest coded value (0 if unsigned)
    int64_t lim_min;   // smallest component value (0 if unsigned)
};

constexpr int MAX_VARS = 64;

struct StatsOut {
    // exponent extrema, stored order-preserving as signed 64-bit (unsigned values are biased by 2^63)
    long long emin[MAX_VARS], emax[MAX_VARS];
    long long dmin, dmax; // (partial) degree extrema, same encoding
    unsigned cfbits;      // max bit length of |coefficient| (integers)
    unsigned nonfinite;   // fp64: number of non-finite coefficients
    unsigned anyneg;      // integers: 1 if some coefficient is negative
    unsigned pad;
};

__device__ __forceinline__ long long ord64(int64_t v, uint32_t is_signed)
{
    return is_signed ? (long long)v : (long long)((uint64_t)v ^ 0x8000000000000000ULL);
}

__device__ __forceinline__ long long shfl_min_ll(long long v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        long long t = __shfl_xor_sync(0xffffffffu, v, o);
        v = t < v ? t : v;
    }
    return v;
}
__device__ __forceinline__ long long shfl_max_ll(long long v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        long long t = __shfl_xor_sync(0xffffffffu, v, o);
        v = t > v ? t : v;
    }
    return v;
}

// Unpack every exponent of every term once: feeds monomial_range_overflow_check
// (packed_monomial.hpp:289-488 / d_packed_monomial.hpp:583-897), key_degree / key_p_degree
// (src/polynomials/packed_monomial.cpp:334-412, d_packed_monomial.hpp:901-958) and the a-priori
// coefficient bound.  var_mask selects the variables that count towards the degree.
__global__ void __launch_bounds__(256) k_term_stats(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ cfs,
                                                    uint64_t n, Layout L, uint64_t var_mask, int cf_kind, int cf_limbs,
                                                    int64_t *__restrict__ deg_out, StatsOut *__restrict__ out)
{
    __shared__ long long s_min[MAX_VARS], s_max[MAX_VARS];
    __shared__ long long s_dmin, s_dmax;
    __shared__ unsigned s_bits, s_nonf, s_neg;
    const int tid = threadIdx.x;
    if (tid < MAX_VARS) {
        s_min[tid] = 0x7fffffffffffffffLL;
        s_max[tid] = (long long)0x8000000000000000ULL;
    }
    if (tid == 0) {
        s_dmin = 0x7fffffffffffffffLL;
        s_dmax = (long long)0x8000000000000000ULL;
        s_bits = 0;
        s_nonf = 0;
        s_neg = 0;
    }
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t nround = (n + 31) / 32 * 32; // keep warps converged for the shuffles
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + tid; i < nround; i += stride) {
        const bool live = i < n;
        int64_t deg = 0;
        uint32_t vi = 0;
        for (uint32_t w = 0; w < L.W; ++w) {
            uint64_t v = live ? keys[i * L.W + w] - L.klim_min : 0;
            for (uint32_t j = 0; j < L.wsize && vi < L.nvars; ++j, ++vi) {
                const uint64_t q = v / L.delta;
                const int64_t e = (int64_t)(v - q * L.delta) + L.lim_min;
                v = q;
                if ((var_mask >> vi) & 1) deg += e;
                long long o = ord64(e, L.is_signed);
                long long mn = shfl_min_ll(live ? o : 0x7fffffffffffffffLL);
                long long mx = shfl_max_ll(live ? o : (long long)0x8000000000000000ULL);
                if ((tid & 31) == 0) {
                    atomicMin(&s_min[vi], mn);
                    atomicMax(&s_max[vi], mx);
                }
            }
        }
        if (live && deg_out) deg_out[i] = deg;
        {
            long long o = ord64(deg, L.is_signed);
            long long mn = shfl_min_ll(live ? o : 0x7fffffffffffffffLL);
            long long mx = shfl_max_ll(live ? o : (long long)0x8000000000000000ULL);
            if ((tid & 31) == 0) {
                atomicMin(&s_dmin, mn);
                atomicMax(&s_dmax, mx);
            }
        }
        unsigned bits = 0, nonf = 0, isneg = 0;
        if (live) {
            if (cf_kind == 0) {
                // bit length of |c| for a two's-complement cf_limbs-limb integer
                const uint64_t *c = cfs + i * cf_limbs;
                const bool neg = (int64_t)c[cf_limbs - 1] < 0;
                for (int l = cf_limbs - 1; l >= 0; --l) {
                    uint64_t m = neg ? ~c[l] : c[l];
                    if (m) {
                        bits = 64u * l + (64u - __clzll((long long)m));
                        break;
                    }
                }
                if (neg) bits += 1; // |-(2^k)| needs k+1 bits; cheap upper bound for every negative value
                isneg = neg;
            } else if (cf_kind == 1) {
                const double d = __longlong_as_double((long long)cfs[i]);
                nonf = !isfinite(d);
            }
        }
        bits = __reduce_max_sync(0xffffffffu, bits);
        nonf = __reduce_add_sync(0xffffffffu, nonf);
        isneg = __any_sync(0xffffffffu, isneg);
        if ((tid & 31) == 0) {
            atomicMax(&s_bits, bits);
            if (nonf) atomicAdd(&s_nonf, nonf);
            if (isneg) s_neg = 1;
        }
    }
    __syncthreads();
    if (tid < (int)L.nvars) {
        atomicMin(&out->emin[tid], s_min[tid]);
        atomicMax(&out->emax[tid], s_max[tid]);
    }
    if (tid == 0) {
        atomicMin(&out->dmin, s_dmin);
        atomicMax(&out->dmax, s_dmax);
        atomicMax(&out->cfbits, s_bits);
        if (s_nonf) atomicAdd(&out->nonfinite, s_nonf);
        if (s_neg) atomicOr(&out->anyneg, 1u);
    }
}

__global__ void k_stats_init(StatsOut *s, int n)
{
    const int i = blockIdx.x;
    if (i >= n) return;
    for (int v = threadIdx.x; v < MAX_VARS; v += blockDim.x) {
        s[i].emin[v] = 0x7fffffffffffffffLL;
        s[i].emax[v] = (long long)0x8000000000000000ULL;
    }
    if (threadIdx.x == 0) {
        s[i].dmin = 0x7fffffffffffffffLL;
        s[i].dmax = (long long)0x8000000000000000ULL;
        s[i].cfbits = 0;
        s[i].nonfinite = 0;
        s[i].anyneg = 0;
        s[i].pad = 0;
    }
}

// ------------------------------------------------------------------------------------------------
// hash / segment of a monomial.
//   hash = value (packed_monomial.hpp:172-176) or the sum of the words (d_packed_monomial.hpp:274-284);
//   the reference routes a product to table hash & (2^k - 1) because hash(a*b) = hash(a) + hash(b)
//   (polynomial.hpp:1441-1450).  The same homomorphism holds modulo ANY m as long as the key addition
//   does not wrap -- which the exponent-overflow check guarantees (the product key stays within the
//   kpack limits) -- so the engine segments by  seg(key) = sum_w (word_w mod m) mod m  with m a prime:
//   the low bits of Kronecker keys are badly clumped (Fateman n=30: fullest 2^8-segment is 1.53x the
//   mean), residues modulo a prime are flat to ~1 %.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t word_mod(uint64_t w, uint32_t m, uint32_t is_signed)
{
    if (is_signed) {
        const long long r = (long long)w % (long long)m;
        return (uint32_t)(r < 0 ? r + (long long)m : r);
    }
    return (uint32_t)(w % m);
}

__device__ __forceinline__ uint32_t key_segment(const uint64_t *k, int W, uint32_t m, uint32_t is_signed)
{
    uint32_t s = 0;
    for (int w = 0; w < W; ++w) {
        s += word_mod(k[w], m, is_signed);
        if (s >= m) s -= m;
    }
    return s;
}

// Segment index of every term of the left operand (read once per visit by the product kernel).
__global__ void __launch_bounds__(256) k_segments(const uint64_t *__restrict__ keys, uint64_t n, int W, uint32_t m,
                                                  uint32_t is_signed, uint32_t *__restrict__ seg_of)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        seg_of[i] = key_segment(keys + i * W, W, m, is_signed);
}

// Counting sort of the right operand by bin = seg(key) << dbits | (deg - degmin)
// (replaces tbb::parallel_sort by bucket + compute_vseg + seg_sorter, polynomial.hpp:914-1105).
// Pass 1: histogram; the value returned by the atomic is the term's rank inside its bin, so pass 3
// scatters without a second round of atomics.  pow2 != 0: bin = hash & (m-1) (output regrouping).
__global__ void __launch_bounds__(256) k_bin_count(const uint64_t *__restrict__ keys, const int64_t *__restrict__ deg,
                                                   uint64_t n, int W, uint32_t m, uint32_t is_signed, int pow2, int dbits,
                                                   int64_t degmin, uint32_t *__restrict__ cnt,
                                                   uint32_t *__restrict__ bin_of, uint32_t *__restrict__ rank_of)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t bin;
        if (pow2) {
            uint64_t h = 0;
            for (int w = 0; w < W; ++w) h += keys[i * W + w];
            bin = (uint32_t)(h & (uint64_t)(m - 1));
        } else {
            bin = key_segment(keys + i * W, W, m, is_signed);
        }
        if (dbits) bin = (bin << dbits) | (uint32_t)(deg[i] - degmin);
        bin_of[i] = bin;
        rank_of[i] = atomicAdd(&cnt[bin], 1u);
    }
}

// Same as k_bin_count for a small number of bins: per-block histogram in shared memory (the block's rank of a
// term comes from the shared atomic), one global atomic per (block, non-empty bin) for the block's base.  Avoids the
// same-address global atomics that dominate k_bin_count when millions of terms fall into ~1000 bins.
template <int BINC_ITEMS>
__global__ void __launch_bounds__(256) k_bin_count_smem(const uint64_t *__restrict__ keys, const int64_t *__restrict__ deg,
                                                        uint64_t n, int W, uint32_t m, uint32_t is_signed, int pow2,
                                                        int dbits, int64_t degmin, uint32_t nbins,
                                                        uint32_t *__restrict__ cnt, uint32_t *__restrict__ bin_of,
                                                        uint32_t *__restrict__ rank_of)
{
    extern __shared__ uint32_t sh_hist[];
    for (uint32_t b = threadIdx.x; b < nbins; b += blockDim.x) sh_hist[b] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * blockDim.x * BINC_ITEMS;
    uint32_t mybin[BINC_ITEMS], myrank[BINC_ITEMS];
#pragma unroll
    for (int j = 0; j < BINC_ITEMS; ++j) {
        const uint64_t i = base + (uint64_t)j * blockDim.x + threadIdx.x;
        mybin[j] = 0xffffffffu;
        if (i < n) {
            uint32_t bin;
            if (pow2) {
                uint64_t h = 0;
                for (int w = 0; w < W; ++w) h += keys[i * W + w];
                bin = (uint32_t)(h & (uint64_t)(m - 1));
            } else {
                bin = key_segment(keys + i * W, W, m, is_signed);
            }
            if (dbits) bin = (bin << dbits) | (uint32_t)(deg[i] - degmin);
            mybin[j] = bin;
        }
        // lanes of a warp that fall into the same bin share ONE shared-memory atomic (degree-skewed operands put
        // most terms into a few bins: one atomic per lane serialises on them)
        const uint32_t peers = __match_any_sync(0xffffffffu, mybin[j]);
        if (mybin[j] != 0xffffffffu) {
            const uint32_t lane = threadIdx.x & 31u;
            const uint32_t leader = (uint32_t)(__ffs(peers) - 1);
            uint32_t r0 = 0;
            if (lane == leader) r0 = atomicAdd(&sh_hist[mybin[j]], (uint32_t)__popc(peers));
            r0 = __shfl_sync(peers, r0, (int)leader);
            myrank[j] = r0 + (uint32_t)__popc(peers & ((1u << lane) - 1u));
        }
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < nbins; b += blockDim.x) {
        const uint32_t c = sh_hist[b];
        sh_hist[b] = c ? atomicAdd(&cnt[b], c) : 0u;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < BINC_ITEMS; ++j) {
        const uint64_t i = base + (uint64_t)j * blockDim.x + threadIdx.x;
        if (mybin[j] != 0xffffffffu) {
            bin_of[i] = mybin[j];
            rank_of[i] = sh_hist[mybin[j]] + myrank[j];
        }
    }
}

// Exclusive scan, three passes (block-local scan, scan of block sums, add back); 4096 items per block.
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_BLOCK = SCAN_ITEMS * SCAN_THREADS;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *total)
{
    __shared__ uint32_t warp_sums[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t s = lane < (int)(blockDim.x >> 5) ? warp_sums[lane] : 0;
        uint32_t si = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, si, o);
            if (lane >= o) si += t;
        }
        warp_sums[lane] = si - s; // exclusive prefix of warp sums
        if (lane == 31) *total = si;
    }
    __syncthreads();
    const uint32_t r = inc - v + warp_sums[wid];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_local(const uint32_t *__restrict__ in, uint32_t *__restrict__ out,
                                                             uint32_t *__restrict__ block_sums, uint64_t n)
{
    __shared__ uint32_t total;
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_BLOCK + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS], s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        v[j] = base + j < n ? in[base + j] : 0;
        s += v[j];
    }
    uint32_t ex = block_exclusive_scan(s, &total);
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        if (base + j < n) out[base + j] = ex;
        ex += v[j];
    }
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// Exclusive scan of a short array by ONE block (n up to a few hundred thousand): chunks of SCAN_BLOCK items with a
// running carry.  Replaces the three launches of the general scan for the many small scans of a call (bucket
// offsets, segment counts), which were launch-latency bound.  out[n] = grand total.
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_small(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, uint64_t n)
{
    __shared__ uint32_t total, carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint64_t base0 = 0; base0 < n; base0 += SCAN_BLOCK) {
        const uint64_t base = base0 + (uint64_t)threadIdx.x * SCAN_ITEMS;
        uint32_t v[SCAN_ITEMS], sum = 0;
#pragma unroll
        for (int j = 0; j < SCAN_ITEMS; ++j) {
            v[j] = base + j < n ? in[base + j] : 0;
            sum += v[j];
        }
        uint32_t ex = block_exclusive_scan(sum, &total) + carry;
#pragma unroll
        for (int j = 0; j < SCAN_ITEMS; ++j) {
            if (base + j < n) out[base + j] = ex;
            ex += v[j];
        }
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry;
}

// out has n+1 entries: out[n] = grand total.
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_add(uint32_t *__restrict__ out,
                                                           const uint32_t *__restrict__ block_prefix, uint64_t n,
                                                           const uint32_t *__restrict__ grand_total)
{
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_BLOCK + (uint64_t)threadIdx.x * SCAN_ITEMS;
    const uint32_t add = block_prefix[blockIdx.x];
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j)
        if (base + j < n) out[base + j] += add;
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = *grand_total;
}

__global__ void __launch_bounds__(256) k_bin_scatter(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ cfs,
                                                     const int64_t *__restrict__ deg, uint64_t n, int W, int CW,
                                                     const uint32_t *__restrict__ off, const uint32_t *__restrict__ bin_of,
                                                     const uint32_t *__restrict__ rank_of, uint64_t *__restrict__ okeys,
                                                     uint64_t *__restrict__ ocfs, int64_t *__restrict__ odeg)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t dst = (uint64_t)off[bin_of[i]] + rank_of[i];
        for (int w = 0; w < W; ++w) okeys[dst * W + w] = keys[i * W + w];
        if (ocfs)
            for (int c = 0; c < CW; ++c) ocfs[dst * CW + c] = cfs[i * CW + c];
        if (odeg) odeg[dst] = deg[i];
    }
}

// Two-sided truncation: flag the terms whose degree is <= thr, then (after a scan of the flags) copy them out.
__global__ void __launch_bounds__(256) k_flag_le(const int64_t *__restrict__ deg, uint64_t n, int64_t thr, uint32_t *__restrict__ flag)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) flag[i] = deg[i] <= thr;
}
__global__ void __launch_bounds__(256) k_select(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ cfs,
                                                const int64_t *__restrict__ deg, uint64_t n, int W, int CW,
                                                const uint32_t *__restrict__ pos /* n+1, exclusive scan of flags */,
                                                uint64_t *__restrict__ okeys, uint64_t *__restrict__ ocfs, int64_t *__restrict__ odeg)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t d = pos[i];
        if (pos[i + 1] == d) continue;
        for (int w = 0; w < W; ++w) okeys[(size_t)d * W + w] = keys[i * W + w];
        if (ocfs)
            for (int c = 0; c < CW; ++c) ocfs[(size_t)d * CW + c] = cfs[i * CW + c];
        odeg[d] = deg[i];
    }
}

// truncate_degree / truncate_p_degree (polynomial.hpp:2364-2445): keep term t iff !(limit < deg(t)), compared in
// the key's integer type T (unsigned T: unsigned comparison).
__global__ void __launch_bounds__(256) k_flag_keep(const int64_t *__restrict__ deg, uint64_t n, int64_t limit, int uns,
                                                   uint32_t *__restrict__ flag)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        flag[i] = uns ? ((uint64_t)deg[i] <= (uint64_t)limit) : (deg[i] <= limit);
}

// Series addition / subtraction (series.hpp:2195-2991) is a merge of the two term lists: copy one operand's
// coefficients into the merge input, sign-extended from lin to lout limbs and negated if asked (fp64: sign flip).
__global__ void __launch_bounds__(256) k_widen_cfs(const uint64_t *__restrict__ in, uint64_t n, int lin, int lout, int negate,
                                                   int f64, uint64_t *__restrict__ out)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (f64) {
            out[i] = negate ? in[i] ^ 0x8000000000000000ULL : in[i];
            continue;
        }
        uint64_t v[5]; // MAX_CF_LIMBS
        const uint64_t ext = (uint64_t)((long long)in[i * lin + lin - 1] >> 63);
        for (int l = 0; l < 5; ++l) v[l] = l < lin ? in[i * lin + l] : ext;
        if (negate) {
            uint64_t c = 1;
            for (int l = 0; l < 5; ++l) {
                const uint64_t t = ~v[l] + c;
                c = (c && t == 0) ? 1 : 0;
                v[l] = t;
            }
        }
        for (int l = 0; l < lout; ++l) out[i * lout + l] = v[l];
    }
}

// Symbol-set extension (series_sym_extender, series.hpp:539-625; key_merge_symbols,
// src/polynomials/packed_monomial.cpp:234-292, d_packed_monomial.hpp:473-520): every key is unpacked, its exponents
// are moved to their positions in the larger symbol set (zeros for the new symbols) and packed again with the
// deltas of the new size.  An exponent outside the new limits sets *overflow (the reference's kpacker throws).
struct ExtendParams {
    const uint64_t *keys;
    uint64_t n;
    Layout lo, ln;          // old and new key layouts
    uint64_t lim_max_new;   // largest component value of the new layout (bit pattern of T)
    uint32_t pos[MAX_VARS]; // position of old variable j in the new symbol set
    uint64_t *okeys;
    uint32_t *overflow;
};

__global__ void __launch_bounds__(128) k_extend_symbols(const ExtendParams P)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += stride) {
        int64_t en[MAX_VARS];
        for (uint32_t v = 0; v < P.ln.nvars; ++v) en[v] = 0;
        uint32_t vi = 0;
        bool bad = false;
        for (uint32_t w = 0; w < P.lo.W; ++w) {
            uint64_t v = P.keys[i * P.lo.W + w] - P.lo.klim_min;
            for (uint32_t j = 0; j < P.lo.wsize && vi < P.lo.nvars; ++j, ++vi) {
                const uint64_t q = v / P.lo.delta;
                const int64_t e = (int64_t)(v - q * P.lo.delta) + P.lo.lim_min;
                v = q;
                en[P.pos[vi]] = e;
                bad |= P.ln.is_signed ? (e < P.ln.lim_min || e > (int64_t)P.lim_max_new) : ((uint64_t)e > P.lim_max_new);
            }
        }
        if (bad) *P.overflow = 1u;
        uint32_t base = 0;
        for (uint32_t w = 0; w < P.ln.W; ++w) {
            // Horner from the highest digit of the word: value = sum_j e_j * delta^j (kpack.hpp:262-267)
            const uint32_t cnt = min(P.ln.wsize, P.ln.nvars - base);
            uint64_t val = 0;
            for (uint32_t j = cnt; j-- > 0;) val = val * P.ln.delta + (uint64_t)en[base + j];
            P.okeys[i * P.ln.W + w] = val;
            base += P.ln.wsize;
        }
    }
}

// tot_n_mults of a truncated product (polynomial.hpp:574-623): for each left term the number of right
// terms whose degree passes the limit, via the degree histogram's prefix sums.
__global__ void __launch_bounds__(256) k_deg_hist(const int64_t *__restrict__ deg, uint64_t n, int64_t degmin,
                                                  uint32_t *__restrict__ hist, uint32_t ndeg_smem)
{
    // ndeg_smem != 0: the histogram has that many bins and fits in shared memory (the usual case: tens of degrees,
    // hundreds of thousands of terms -- global atomics on a dozen addresses would serialise)
    extern __shared__ uint32_t sh_dh[];
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    if (ndeg_smem) {
        for (uint32_t b = threadIdx.x; b < ndeg_smem; b += blockDim.x) sh_dh[b] = 0;
        __syncthreads();
        // a handful of degrees shared by 10^5 terms: lanes holding the same degree elect one to add their count
        const uint64_t nround = (n + 31) / 32 * 32;
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += stride) {
            const uint32_t d = i < n ? (uint32_t)(deg[i] - degmin) : 0xffffffffu;
            const uint32_t peers = __match_any_sync(0xffffffffu, d);
            if (d != 0xffffffffu && (threadIdx.x & 31u) == (uint32_t)(__ffs(peers) - 1)) atomicAdd(&sh_dh[d], (uint32_t)__popc(peers));
        }
        __syncthreads();
        for (uint32_t b = threadIdx.x; b < ndeg_smem; b += blockDim.x)
            if (sh_dh[b]) atomicAdd(&hist[b], sh_dh[b]);
    } else {
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
            atomicAdd(&hist[(uint32_t)(deg[i] - degmin)], 1u);
    }
}

// lim_base = limit - degmin_y expressed so that right-degree index d passes for left term i iff
// d <= lim_base - xdeg[i]  (all in int64; the host has already ruled out wrap-around).
__global__ void __launch_bounds__(256) k_count_mults(const int64_t *__restrict__ xdeg, uint64_t nx, int64_t lim_base,
                                                     uint32_t ndeg, const uint32_t *__restrict__ prefix /* ndeg+1 */,
                                                     unsigned long long *__restrict__ tot)
{
    unsigned long long local = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nx; i += stride) {
        const int64_t r = lim_base - xdeg[i];
        if (r < 0) continue;
        const uint32_t d = r >= (int64_t)ndeg ? ndeg - 1 : (uint32_t)r;
        local += prefix[d + 1];
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(tot, local);
}

// ------------------------------------------------------------------------------------------------
// Coefficient arithmetic (polynomial.hpp:1368-1384; fma3.hpp:84-88 mppp::addmul; :65-72 double)
// ------------------------------------------------------------------------------------------------
enum CfMode { CF_SYM = 0, CF_I1A1 = 1, CF_I1A2 = 2, CF_I1A3 = 3, CF_F64 = 4, CF_I2A3 = 5, CF_ADD1 = 6, CF_ADD2 = 7,
              CF_ADD3 = 8, CF_ADDF = 9, CF_I2A5 = 10, CF_I3A5 = 11, CF_ADD5 = 12 };
constexpr int MAX_CF_LIMBS = 5; // widest coefficient the engine carries (320 bits): inputs up to 3 limbs, sums up to 5

template <int MODE>
struct CfTraits;
// LIN: limbs of one right-operand coefficient; LACC: accumulator limbs; ADD: merge mode (the "product" is the
// right coefficient itself, there is no left coefficient).
#define OBK_TRAITS(M, LIN_, LACC_, F64_, SYM_, ADD_)                                                                   \
    template <>                                                                                                        \
    struct CfTraits<M> {                                                                                               \
        static constexpr int LIN = LIN_, LACC = LACC_;                                                                 \
        static constexpr bool F64 = F64_, SYM = SYM_, ADD = ADD_;                                                      \
    };
OBK_TRAITS(CF_SYM, 0, 0, false, true, false)
OBK_TRAITS(CF_I1A1, 1, 1, false, false, false)
OBK_TRAITS(CF_I1A2, 1, 2, false, false, false)
OBK_TRAITS(CF_I1A3, 1, 3, false, false, false)
OBK_TRAITS(CF_I2A3, 2, 3, false, false, false)
OBK_TRAITS(CF_F64, 1, 1, true, false, false)
OBK_TRAITS(CF_ADD1, 1, 1, false, false, true)
OBK_TRAITS(CF_ADD2, 2, 2, false, false, true)
OBK_TRAITS(CF_ADD3, 3, 3, false, false, true)
OBK_TRAITS(CF_ADDF, 1, 1, true, false, true)
// coefficient growth along chains of products (the reference's integer<N> grows without bound, fma3.hpp:84-88): two- and
// three-limb operands with a five-limb accumulator
OBK_TRAITS(CF_I2A5, 2, 5, false, false, false)
OBK_TRAITS(CF_I3A5, 3, 5, false, false, false)
OBK_TRAITS(CF_ADD5, 5, 5, false, false, true)
#undef OBK_TRAITS

// p (LACC limbs, two's complement) = a * b for LIN-limb two's-complement inputs.
template <int LIN, int LACC>
__device__ __forceinline__ void cf_mul(const uint64_t *a, const uint64_t *b, uint64_t *p, bool uns = false)
{
    if constexpr (LIN == 1) {
        if (uns) {
            // no negative coefficient in either operand: plain unsigned product, no sign fix-ups
            p[0] = a[0] * b[0];
            if constexpr (LACC >= 2) {
                p[1] = __umul64hi(a[0], b[0]);
                if constexpr (LACC >= 3) p[2] = 0;
            }
            return;
        }
        const long long x = (long long)a[0], y = (long long)b[0];
        p[0] = (uint64_t)(x * y);
        if constexpr (LACC >= 2) {
            const long long hi = __mul64hi(x, y);
            p[1] = (uint64_t)hi;
            if constexpr (LACC >= 3) p[2] = (uint64_t)(hi >> 63);
        }
    } else if constexpr (LIN == 2 && LACC <= 3) {
        // 128 x 128 -> low 192 bits of the signed product (the a-priori bound guarantees it fits).
        // Signed product mod 2^192 = unsigned schoolbook on the sign-extended operands, truncated.
        const uint64_t a0 = a[0], a1 = a[1], b0 = b[0], b1 = b[1];
        const uint64_t a2 = (uint64_t)((long long)a1 >> 63), b2 = (uint64_t)((long long)b1 >> 63);
        uint64_t r0 = a0 * b0;
        uint64_t c = __umul64hi(a0, b0);
        // column 1: a0*b1 + a1*b0 + c
        uint64_t t0 = a0 * b1, t1 = a1 * b0;
        uint64_t r1 = c + t0;
        uint64_t cy = r1 < t0;
        r1 += t1;
        cy += r1 < t1;
        // column 2 (low parts only matter mod 2^192)
        uint64_t r2 = __umul64hi(a0, b1) + __umul64hi(a1, b0) + a1 * b1 + a0 * b2 + a2 * b0 + cy;
        p[0] = r0;
        p[1] = r1;
        if constexpr (LACC >= 3) p[2] = r2;
    } else {
        // General case: the signed product modulo 2^(64 LACC) is the unsigned schoolbook product of the operands
        // sign-extended to LACC limbs, truncated.  Column k sums a_i b_(k-i) into a three-word carry-save window.
        uint64_t ax[LACC], bx[LACC];
        const uint64_t ea = (uint64_t)((long long)a[LIN - 1] >> 63), eb = (uint64_t)((long long)b[LIN - 1] >> 63);
#pragma unroll
        for (int l = 0; l < LACC; ++l) {
            ax[l] = l < LIN ? a[l] : ea;
            bx[l] = l < LIN ? b[l] : eb;
        }
        uint64_t c0 = 0, c1 = 0, c2 = 0;
#pragma unroll
        for (int k = 0; k < LACC; ++k) {
#pragma unroll
            for (int i = 0; i <= k; ++i) {
                const uint64_t lo = ax[i] * bx[k - i], hi = __umul64hi(ax[i], bx[k - i]);
                asm("add.cc.u64 %0, %0, %3;\n\taddc.cc.u64 %1, %1, %4;\n\taddc.u64 %2, %2, 0;" : "+l"(c0), "+l"(c1), "+l"(c2) : "l"(lo), "l"(hi));
            }
            p[k] = c0;
            c0 = c1;
            c1 = c2;
            c2 = 0;
        }
    }
}

template <int LACC>
__device__ __forceinline__ void limbs_add(uint64_t *acc, const uint64_t *p)
{
    if constexpr (LACC == 1) {
        acc[0] += p[0];
    } else if constexpr (LACC == 2) {
        asm("add.cc.u64 %0, %0, %2;\n\taddc.u64 %1, %1, %3;" : "+l"(acc[0]), "+l"(acc[1]) : "l"(p[0]), "l"(p[1]));
    } else if constexpr (LACC == 3) {
        asm("add.cc.u64 %0, %0, %3;\n\taddc.cc.u64 %1, %1, %4;\n\taddc.u64 %2, %2, %5;"
            : "+l"(acc[0]), "+l"(acc[1]), "+l"(acc[2])
            : "l"(p[0]), "l"(p[1]), "l"(p[2]));
    } else {
        uint64_t carry = 0;
#pragma unroll
        for (int l = 0; l < LACC; ++l) {
            const uint64_t t = acc[l] + p[l];
            const uint64_t c1 = t < p[l];
            const uint64_t u = t + carry;
            carry = c1 | (uint64_t)(u < t);
            acc[l] = u;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// The product kernel
// ------------------------------------------------------------------------------------------------
struct MulParams {
    // left operand, original order (streamed by every CTA)
    const uint64_t *xk;
    const uint64_t *xc;
    const int64_t *xdeg;
    const uint32_t *xseg; // segment of each left term
    uint64_t nx;
    // right operand, sorted by (segment, degree); yoff has (nsegs << dbits) + 1 entries
    const uint64_t *yk;
    const uint64_t *yc;
    const uint32_t *yoff;
    uint32_t nsegs; // the modulus m
    uint32_t dbits;
    uint32_t ndeg;
    uint32_t trunc;
    int64_t lim_base;
    uint32_t dlo;   // lowest right-degree index a left term may pair with (0 unless two-sided truncation)
    uint32_t npass; // 2: a second pass with the operands' roles swapped follows (two-sided truncation)
    // Second pass of a heavily truncated product (every admissible pair has its LOWER-degree factor on the left):
    // left = the low-degree terms of the right operand, right = the left operand sorted by (segment, degree),
    // degree window [dlo, lim_base - deg].  Same fields as above.
    struct PassB {
        const uint64_t *xk;
        const uint64_t *xc;
        const int64_t *xdeg;
        const uint32_t *xseg;
        uint64_t nx;
        const uint64_t *yk;
        const uint64_t *yc;
        const uint32_t *yoff;
        uint32_t dbits, ndeg, dlo, pad;
        int64_t lim_base;
    } b;
    // work list: segment ids to process (NULL = all nsegs) and the dynamic cursor
    const uint32_t *seg_list;
    uint32_t nseg_list;
    uint32_t seg_base; // seg_list == NULL: the segments processed are seg_base .. seg_base + nseg_list - 1
    uint32_t *cursor;
    // shared-memory table geometry
    uint32_t log2_cap;
    uint32_t max_fill;
    uint32_t fence; // 1: fence.acq_rel.cta around the slot lock (validation build of the protocol)
    uint32_t lx_batch; // lanes that must be waiting for a new left term before the LX step is issued (1 = always)
    // k_mul (warp-cooperative expansion): every left term's run is cut into `rslices` slices (work unit = 32 left
    // terms x one slice) and idle lanes are refilled once at least `refill_min` of a warp's lanes are idle
    uint32_t rslices, refill_min;
    uint32_t uns; // 1: no coefficient is negative (unsigned 64x64 products)
    // staging output: [segment][cap] slots, counts per segment
    uint64_t *okeys;
    uint64_t *ocfs;
    uint32_t *counts;
    uint32_t *flags; // [0] = number of segments whose table overflowed
    // k_mul: when set, okeys/ocfs are the FINAL term arrays and every segment reserves its range with one atomic
    // on this counter (its final value is the number of terms); counts is then unused
    uint32_t *out_cursor;
};

enum : uint32_t { ST_EMPTY = 0, ST_INIT = 1, ST_FULL = 2, ST_BUSY = 3 };

template <int W>
__device__ __forceinline__ uint32_t slot_hash(const uint64_t (&k)[W], uint32_t log2_cap)
{
    if constexpr (W == 1) {
        // one packed word: multilinear hash of its two halves (2 IMAD + shift); the top bits of
        // lo * A + hi * B mod 2^32 are flat for the arithmetic progressions a segment's keys form
        const uint32_t x = (uint32_t)k[0] * 0x9E3779B1u + (uint32_t)(k[0] >> 32) * 0x85EBCA77u;
        return (x ^ (x >> 15)) * 0x2C1B3C6Du >> (32 - log2_cap);
    }
    // multiply-shift over the whole key: the top bits of the product depend on every key bit
    uint64_t h = k[0] * 0x9E3779B97F4A7C15ULL;
#pragma unroll
    for (int w = 1; w < W; ++w) h = (h ^ (h >> 29)) * 0xBF58476D1CE4E5B9ULL + k[w] * 0x94D049BB133111EBULL;
    h ^= h >> 32;
    h *= 0xD6E8FEB86659FD93ULL;
    return (uint32_t)(h >> (64 - log2_cap));
}

// Slot lock protocol.  A slot's state word goes EMPTY -> INIT -> FULL and then FULL <-> BUSY; key and
// accumulator limbs are only written by the thread that moved the state to INIT or BUSY.  In PTX-memory-model
// terms: the state is read with ld.acquire.cta, locked with atom.acquire.cta.cas and released with
// st.release.cta, so the key / limb accesses of successive owners of a slot are ordered by release-acquire
// pairs on the state word.  For shared memory ptxas lowers the two acquire forms to the plain LDS / ATOMS.CAS
// (the shared-memory pipe of one SM performs accesses in issue order) and the release to MEMBAR.ALL.CTA + STS
// (profiles/r02_sass_slot_lock.txt): the only cost of the model-clean protocol is one MEMBAR per accumulated
// product.  fence == 0 (option "fence") replaces the release by a plain store behind a compiler barrier -- the
// hardware-order argument alone -- and exists to measure that cost; the default is fence = 1.
__device__ __forceinline__ uint32_t slot_ld_acquire(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t slot_cas_acquire(uint32_t *p, uint32_t cmp, uint32_t val)
{
    uint32_t old;
    asm volatile("atom.acquire.cta.shared.cas.b32 %0, [%1], %2, %3;"
                 : "=r"(old)
                 : "r"((uint32_t)__cvta_generic_to_shared(p)), "r"(cmp), "r"(val)
                 : "memory");
    return old;
}
__device__ __forceinline__ void slot_st_release(uint32_t *p, uint32_t val, uint32_t fence)
{
    if (fence) {
        asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(val) : "memory");
    } else {
        asm volatile("" ::: "memory");
        *(volatile uint32_t *)p = val;
    }
}

// ------------------------------------------------------------------------------------------------
// k_mul -- the segment-owner product.
//
// One CTA owns one output segment (polynomial.hpp:1421-1553, dense_par_functor): its open-addressing table lives
// in shared memory; every left term x_i pairs with the right bucket (seg - bucket(x_i)) mod nsegs (:1449-1456);
// truncation keeps the prefix of that bucket with deg_j <= limit - deg_i (compute_end_idx2, :1187-1238 -- the
// bucket is degree-sorted, so the prefix is one table lookup).
//
// Term products are expanded WARP-COOPERATIVELY.  (The first two versions of this kernel gave every lane its own
// left term and let it walk that term's run; lanes then sit in different phases of their walk and each trip of the
// warp loop issues the "next left term", "next product" and "probe" bodies for a third of the lanes each -- ncu:
// 12 of 32 threads active, 13 warp instructions per product; profiles/r01_k_mul_S12_*.)  Here a WARP takes a work
// unit of 32 left terms, scans their run lengths, and the products of the unit form one index space [0, T): any
// idle lane is handed the next product index, finds the owning left term by a 5-step binary search over the
// scanned lengths (shuffles), fetches that term's key and coefficient from the owner lane and forms the product.
// Product formation therefore runs with every idle lane busy whatever the run lengths are (5 terms in S12, 10^5
// in a rectangular product, 0 in most runs of a heavily truncated one), and the only divergent code left is the
// probe attempt itself.  Units are handed out by a per-CTA counter; `rslices` cuts long runs so that a product
// with a handful of left terms (or the one-term "left operand" of the owner-side merge) still yields enough units
// for 32 warps.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, uint32_t src)
{
    const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, (int)src);
    const uint32_t hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), (int)src);
    return ((uint64_t)hi << 32) | lo;
}

template <int W, int MODE, bool TWO>
__global__ void __launch_bounds__(1024) k_mul(const MulParams P)
{
    using T = CfTraits<MODE>;
    constexpr int ACCW = T::SYM ? 0 : T::LACC;
    constexpr int CW = T::SYM ? 0 : T::LIN;
    constexpr int XCW = (T::SYM || T::ADD) ? 0 : CW;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t cap = 1u << P.log2_cap;
    uint64_t *tkeys = reinterpret_cast<uint64_t *>(smem_raw);                    // [W][cap]
    uint64_t *tacc = tkeys + (size_t)W * cap;                                   // [ACCW][cap]
    uint32_t *tstate = reinterpret_cast<uint32_t *>(tacc + (size_t)ACCW * cap); // [cap]
    __shared__ uint32_t s_seg, s_count, s_out, s_over, s_unit, s_base, s_pos;

    const uint32_t tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31u;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t cmask = cap - 1u;
    const uint32_t R = P.rslices ? P.rslices : 1u;
    const uint32_t nchA = (uint32_t)((P.nx + 31u) >> 5);
    const uint32_t nchB = (TWO && P.npass > 1) ? (uint32_t)((P.b.nx + 31u) >> 5) : 0u;
    const uint32_t nunits = (nchA + nchB) * R;
    const uint32_t refill_min = P.refill_min ? P.refill_min : 1u;

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            s_seg = atomicAdd(P.cursor, 1u);
            s_count = 0;
            s_out = 0;
            s_over = 0;
            s_unit = 0;
        }
        for (uint32_t h = tid; h < cap; h += nthr) tstate[h] = ST_EMPTY;
        __syncthreads();
        const uint32_t work = s_seg;
        if (work >= P.nseg_list) break;
        const uint32_t seg = P.seg_list ? P.seg_list[work] : P.seg_base + work;

        // ---- warp state: the current unit -----------------------------------------------------------------
        uint32_t Tn = 0, next = 0; // products of the unit, next product index to hand out
        bool more = true;          // units remain
        uint64_t kx[W];
        uint64_t cx[XCW > 0 ? XCW : 1];
        uint32_t ybase = 0, sincl = 0; // first right term of the lane's run minus its exclusive offset; inclusive scan
        const uint64_t *ykp = P.yk, *ycp = P.yc;
#pragma unroll
        for (int w = 0; w < W; ++w) kx[w] = 0;
        cx[0] = 0;
        // ---- lane state: one product in flight ------------------------------------------------------------
        bool pend = false;
        uint64_t k[W];
        uint64_t p[ACCW > 0 ? ACCW : 1];
        uint32_t h = 0, nins = 0;
        for (uint32_t trip = 0;; ++trip) {
            // Every 16 trips: report this warp's insertions (one shared atomic per warp instead of one per insertion;
            // the table has 10 % of headroom above max_fill and a lane never writes outside it) and poll the overflow
            // flag, so a full table cannot trap a lane for longer than that.
            if ((trip & 15u) == 0u) {
                const uint32_t tot = __reduce_add_sync(0xffffffffu, nins);
                nins = 0;
                if (tot && lane == 0 && atomicAdd(&s_count, tot) + tot > P.max_fill) s_over = 1;
                if (__any_sync(0xffffffffu, *(volatile uint32_t *)&s_over != 0u)) break; // warp-uniform exit
            }
            const uint32_t need = __ballot_sync(0xffffffffu, !pend);
            const uint32_t nneed = __popc(need);
            if (nneed >= refill_min || nneed == 32u) {
                while (next >= Tn && more) {
                    // ---- next unit: 32 left terms x one slice of their runs (converged, coalesced) ---------
                    uint32_t u = 0;
                    if (lane == 0) u = atomicAdd(&s_unit, 1u);
                    u = __shfl_sync(0xffffffffu, u, 0);
                    if (u >= nunits) {
                        more = false;
                        break;
                    }
                    uint32_t c = u, r = 0;
                    if (R > 1u) {
                        c = u / R;
                        r = u - c * R;
                    }
                    const bool pb = TWO && c >= nchA;
                    // lane l takes left term c + l * nch: the 32 terms of a unit are far apart in the operand's order.
                    // Neighbouring terms differ by a small exponent vector d, and x + y = (x + d) + (y - d) whenever
                    // y - d is a term too, so products formed side by side would keep hitting the same slot and lose
                    // the slot lock to each other (ncu: 27 % of the lock attempts failed with contiguous units).
                    const uint64_t xi = (uint64_t)(pb ? c - nchA : c) + (uint64_t)lane * (pb ? nchB : nchA);
                    uint32_t len = 0, y0 = 0;
                    if (xi < (pb ? P.b.nx : P.nx)) {
                        const uint64_t *xkp = pb ? P.b.xk : P.xk;
                        const uint32_t *yoffp = pb ? P.b.yoff : P.yoff;
#pragma unroll
                        for (int w = 0; w < W; ++w) kx[w] = __ldg(xkp + xi * W + w);
                        const uint32_t bx = __ldg((pb ? P.b.xseg : P.xseg) + xi);
                        const uint32_t by = seg >= bx ? seg - bx : seg + P.nsegs - bx; // (seg - bx) mod m
                        uint32_t y1;
                        bool any = true;
                        if (P.trunc) {
                            const uint32_t dbits = pb ? P.b.dbits : P.dbits, ndeg = pb ? P.b.ndeg : P.ndeg;
                            const uint32_t dlo = pb ? P.b.dlo : P.dlo;
                            const int64_t rr = (pb ? P.b.lim_base : P.lim_base) - __ldg((pb ? P.b.xdeg : P.xdeg) + xi);
                            any = rr >= (int64_t)dlo;
                            const uint32_t d = rr >= (int64_t)ndeg ? ndeg - 1u : (uint32_t)(rr < 0 ? 0 : rr);
                            y0 = __ldg(yoffp + ((size_t)by << dbits) + dlo);
                            y1 = __ldg(yoffp + ((size_t)by << dbits) + d + 1u);
                        } else {
                            y0 = __ldg(yoffp + by);
                            y1 = __ldg(yoffp + by + 1u);
                        }
                        if (any && y1 > y0) {
                            len = y1 - y0;
                            if (R > 1u) {
                                const uint32_t a = (uint32_t)(((uint64_t)len * r) / R);
                                const uint32_t b = (uint32_t)(((uint64_t)len * (r + 1u)) / R);
                                y0 += a;
                                len = b - a;
                            }
                        }
                        if constexpr (XCW > 0) {
                            const uint64_t *xcp = pb ? P.b.xc : P.xc;
#pragma unroll
                            for (int cc = 0; cc < XCW; ++cc) cx[cc] = __ldg(xcp + xi * XCW + cc);
                        }
                    }
                    uint32_t sc = len;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t t = __shfl_up_sync(0xffffffffu, sc, o);
                        if ((int)lane >= o) sc += t;
                    }
                    sincl = sc;
                    ybase = y0 - (sc - len);
                    Tn = __shfl_sync(0xffffffffu, sc, 31);
                    next = 0;
                    ykp = (TWO && pb) ? P.b.yk : P.yk;
                    ycp = (TWO && pb) ? P.b.yc : P.yc;
                }
                if (next < Tn) {
                    // ---- hand the next product indices to the idle lanes ------------------------------------
                    const uint32_t avail = Tn - next;
                    const uint32_t rank = __popc(need & lt_mask);
                    const bool get = !pend && rank < avail;
                    const uint32_t idx = next + (get ? rank : 0u);
                    uint32_t own = 0; // number of lanes whose inclusive offset is <= idx = the owner lane
#pragma unroll
                    for (uint32_t step = 16; step; step >>= 1) {
                        const uint32_t v = __shfl_sync(0xffffffffu, sincl, (int)(own + step - 1u));
                        if (v <= idx) own += step;
                    }
                    const uint32_t yb = __shfl_sync(0xffffffffu, ybase, (int)own);
                    uint64_t kxo[W];
#pragma unroll
                    for (int w = 0; w < W; ++w) kxo[w] = shfl_u64(kx[w], own);
                    uint64_t cxo[XCW > 0 ? XCW : 1];
#pragma unroll
                    for (int cc = 0; cc < XCW; ++cc) cxo[cc] = shfl_u64(cx[cc], own);
                    if (get) {
                        const size_t t = (size_t)(yb + idx);
#pragma unroll
                        for (int w = 0; w < W; ++w) k[w] = kxo[w] + __ldg(ykp + t * W + w); // monomial_mul
                        if constexpr (!T::SYM) {
                            uint64_t cy[CW];
#pragma unroll
                            for (int cc = 0; cc < CW; ++cc) cy[cc] = __ldg(ycp + t * CW + cc);
                            if constexpr (T::ADD) {
#pragma unroll
                                for (int cc = 0; cc < CW; ++cc) p[cc] = cy[cc];
                            } else if constexpr (T::F64) {
                                // cf += c1*c2 with two roundings (no FP_FAST_FMA in the reference build, fma3.hpp:65-72)
                                p[0] = (uint64_t)__double_as_longlong(
                                    __dmul_rn(__longlong_as_double((long long)cxo[0]), __longlong_as_double((long long)cy[0])));
                            } else {
                                cf_mul<T::LIN, T::LACC>(cxo, cy, p, P.uns != 0u);
                            }
                        }
                        h = slot_hash<W>(k, P.log2_cap);
                        pend = true;
                    }
                    next += nneed < avail ? nneed : avail;
                } else if (nneed == 32u) {
                    break; // no unit left and nothing in flight
                }
            }
            bool inserted = false;
            if (pend) {
                // ---- one find-or-insert attempt: a single path for "accumulate into my key's slot" and "claim an
                // empty slot" (the two differ only in the lock value and in where the old accumulator comes from), so
                // the lanes of a warp do not serialise on the kind of slot they found ---------------------------------
                const uint32_t st = slot_ld_acquire(&tstate[h]);
                bool eq = true;
#pragma unroll
                for (int w = 0; w < W; ++w) eq &= (*(volatile uint64_t *)&tkeys[(size_t)w * cap + h] == k[w]);
                const bool ins = st == ST_EMPTY;
                if constexpr (T::SYM) {
                    if (st >= ST_FULL) {
                        if (eq) pend = false; else h = (h + 1u) & cmask;
                    } else if (ins && slot_cas_acquire(&tstate[h], ST_EMPTY, ST_INIT) == ST_EMPTY) {
#pragma unroll
                        for (int w = 0; w < W; ++w) *(volatile uint64_t *)&tkeys[(size_t)w * cap + h] = k[w];
                        slot_st_release(&tstate[h], ST_FULL, P.fence);
                        inserted = true;
                        pend = false;
                    }
                } else {
                    if (ins || (st == ST_FULL && eq)) {
                        if (slot_cas_acquire(&tstate[h], st, ins ? (uint32_t)ST_INIT : (uint32_t)ST_BUSY) == st) {
                            uint64_t a[ACCW];
#pragma unroll
                            for (int l = 0; l < ACCW; ++l) a[l] = ins ? 0ull : *(volatile uint64_t *)&tacc[(size_t)l * cap + h];
                            if constexpr (T::F64) {
                                a[0] = ins ? p[0]
                                           : (uint64_t)__double_as_longlong(__dadd_rn(__longlong_as_double((long long)a[0]),
                                                                                      __longlong_as_double((long long)p[0])));
                            } else {
                                limbs_add<ACCW>(a, p);
                            }
                            if (ins) {
#pragma unroll
                                for (int w = 0; w < W; ++w) *(volatile uint64_t *)&tkeys[(size_t)w * cap + h] = k[w];
                            }
#pragma unroll
                            for (int l = 0; l < ACCW; ++l) *(volatile uint64_t *)&tacc[(size_t)l * cap + h] = a[l];
                            slot_st_release(&tstate[h], ST_FULL, P.fence);
                            inserted = ins;
                            pend = false;
                        }
                        // else: lost the race for the slot; look again next trip
                    } else if (st >= ST_FULL && !eq) {
                        h = (h + 1u) & cmask; // the table is never completely full (max_fill < cap): probing ends
                    }
                    // ST_INIT, or my key's slot is BUSY: another thread is writing it; look again next trip
                }
            }
            nins += inserted ? 1u : 0u;
        }
        // insertions not yet reported (the loop was left between two polls)
        {
            const uint32_t tot = __reduce_add_sync(0xffffffffu, nins);
            if (tot && lane == 0 && atomicAdd(&s_count, tot) + tot > P.max_fill) s_over = 1;
        }
        __syncthreads();
        // ---- flush: erase zeros (polynomial.hpp:1528-1540) and write the segment out --------------------------
        // Products (P.out_cursor set): the segment's non-zero terms are counted first, one global atomic reserves
        // their range in the FINAL arrays and the terms are written there -- no staging area, no compaction pass.
        // Symbolic products that keep their keys (row path) stage them per segment for k_compact.
        const bool over = s_over != 0;
        uint64_t *ok = nullptr, *oc = nullptr;
        bool direct = false;
        if constexpr (T::SYM) {
            if (P.okeys != nullptr) ok = P.okeys + (size_t)work * cap * W;
        } else {
            direct = P.out_cursor != nullptr;
            ok = direct ? P.okeys : P.okeys + (size_t)seg * cap * W;
            oc = direct ? P.ocfs : P.ocfs + (size_t)seg * cap * ACCW;
        }
        auto slot_live = [&](uint32_t hh, uint64_t(&a)[ACCW > 0 ? ACCW : 1]) -> bool {
            bool live = hh < cap && tstate[hh] == ST_FULL;
            if constexpr (!T::SYM) {
                if (live) {
                    bool nz = false;
#pragma unroll
                    for (int l = 0; l < ACCW; ++l) {
                        a[l] = tacc[(size_t)l * cap + hh];
                        nz |= a[l] != 0;
                    }
                    if constexpr (T::F64) nz = __longlong_as_double((long long)a[0]) != 0.0;
                    live = nz;
                }
            }
            return live;
        };
        if (direct && !over) {
            uint32_t mine = 0;
            for (uint32_t h0 = 0; h0 < cap; h0 += nthr) {
                uint64_t a[ACCW > 0 ? ACCW : 1];
                mine += slot_live(h0 + tid, a) ? 1u : 0u;
            }
            mine = __reduce_add_sync(0xffffffffu, mine);
            if (lane == 0 && mine) atomicAdd(&s_out, mine);
            __syncthreads();
            if (tid == 0) {
                s_base = atomicAdd(P.out_cursor, s_out);
                s_pos = 0;
            }
            __syncthreads();
        }
        if (ok != nullptr && !over) {
            const size_t obase = direct ? (size_t)s_base : 0;
            uint32_t *ctr = direct ? &s_pos : &s_out;
            for (uint32_t h0 = 0; h0 < cap; h0 += nthr) {
                const uint32_t hh = h0 + tid;
                uint64_t a[ACCW > 0 ? ACCW : 1];
                const bool live = slot_live(hh, a);
                const uint32_t b = __ballot_sync(0xffffffffu, live);
                if (b == 0u) continue;
                uint32_t base = 0;
                const uint32_t leader = (uint32_t)(__ffs(b) - 1);
                if (lane == leader) base = atomicAdd(ctr, __popc(b));
                base = __shfl_sync(0xffffffffu, base, (int)leader);
                if (live) {
                    const size_t pos = obase + base + __popc(b & lt_mask);
#pragma unroll
                    for (int w = 0; w < W; ++w) ok[pos * W + w] = tkeys[(size_t)w * cap + hh];
#pragma unroll
                    for (int l = 0; l < ACCW; ++l) oc[pos * ACCW + l] = a[l];
                }
            }
        }
        __syncthreads();
        if (tid == 0) {
            if constexpr (T::SYM) {
                P.counts[work] = over ? 0xffffffffu : s_count;
            } else {
                if (!direct) P.counts[seg] = over ? 0u : s_out;
            }
            if (over) atomicAdd(&P.flags[0], 1u);
        }
    }
}

// Segment-ordered dense output: one CTA copies one segment's staged terms to its final offset.
__global__ void __launch_bounds__(256) k_compact(const uint64_t *__restrict__ skeys, const uint64_t *__restrict__ scfs,
                                                 const uint32_t *__restrict__ counts, const uint32_t *__restrict__ offs,
                                                 uint32_t cap, int W, int ACCW, uint64_t *__restrict__ okeys,
                                                 uint64_t *__restrict__ ocfs, uint64_t *__restrict__ off64, uint32_t nsegs)
{
    for (uint32_t seg = blockIdx.x; seg < nsegs; seg += gridDim.x) {
        const uint32_t n = counts[seg], o = offs[seg];
        const uint64_t *sk = skeys + (size_t)seg * cap * W;
        const uint64_t *sc = scfs + (size_t)seg * cap * ACCW;
        for (uint32_t i = threadIdx.x; i < n * W; i += blockDim.x) okeys[(size_t)o * W + i] = sk[i];
        for (uint32_t i = threadIdx.x; i < n * ACCW; i += blockDim.x) ocfs[(size_t)o * ACCW + i] = sc[i];
        if (threadIdx.x == 0) {
            off64[seg] = o;
            if (seg == nsegs - 1) off64[nsegs] = (uint64_t)o + n;
        }
    }
}


// ------------------------------------------------------------------------------------------------
// k_push -- multi-GPU exchange fused into the producer: every term of this rank's partial product is written
// straight into the receive window of the rank that owns it, over NVLink peer memory (plain P2P stores; the slot
// range is reserved with ONE remote atomic per destination and CTA).  Replaces regroup-by-owner + D2H of the
// group offsets + all-to-all of counts + two NCCL all-to-alls.  Ownership is the same rule as the NCCL path:
// group = key mod OBK_EXCHANGE_MOD, contiguous group ranges per rank.
// Window layout (one per rank and step parity): [0] term counter, [4] overflow flag, records from byte 256:
// keys[cap][W] then cfs[cap][L].
// ------------------------------------------------------------------------------------------------
constexpr int PUSH_MAX_RANKS = 16;
constexpr int PUSH_ITEMS = 4;
constexpr uint32_t PUSH_HEADER = 256;

struct PushParams {
    const uint64_t *keys;
    const uint64_t *cfs;
    uint64_t n;
    int W, L;
    uint32_t is_signed, mod, nranks;
    uint64_t cap;                        // terms one window can hold
    unsigned char *win[PUSH_MAX_RANKS];  // rank g's window (this step's parity) as mapped in this process
    uint32_t first_group[PUSH_MAX_RANKS + 1]; // rank g owns groups [first_group[g], first_group[g+1])
    uint32_t *local_overflow;            // set when some window was too small
};

__global__ void __launch_bounds__(256) k_push(const PushParams P)
{
    // The records of a tile are first sorted by destination in shared memory, then each destination's run is copied
    // out with consecutive threads writing consecutive words: NVLink sees 256-byte contiguous stores instead of
    // scattered 8-byte ones (the scattered version spent 0.57 ms on 0.55 M terms at 8 GPUs).
    extern __shared__ __align__(16) unsigned char push_smem[];
    __shared__ uint32_t s_cnt[PUSH_MAX_RANKS], s_base[PUSH_MAX_RANKS], s_off[PUSH_MAX_RANKS + 1];
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const uint32_t tile_terms = blockDim.x * PUSH_ITEMS;
    uint64_t *sk = reinterpret_cast<uint64_t *>(push_smem);      // [tile_terms][W]
    uint64_t *scf = sk + (size_t)tile_terms * P.W;              // [tile_terms][L]
    for (uint64_t tile = (uint64_t)blockIdx.x * tile_terms; tile < P.n; tile += (uint64_t)gridDim.x * tile_terms) {
        if (tid < PUSH_MAX_RANKS) s_cnt[tid] = 0;
        __syncthreads();
        uint32_t dest[PUSH_ITEMS], rank_in[PUSH_ITEMS];
#pragma unroll
        for (int j = 0; j < PUSH_ITEMS; ++j) {
            const uint64_t i = tile + (uint64_t)j * blockDim.x + tid;
            uint32_t g = 0xffffffffu;
            if (i < P.n) {
                const uint32_t grp = key_segment(P.keys + i * P.W, P.W, P.mod, P.is_signed);
                g = (uint32_t)(((uint64_t)grp * P.nranks) / P.mod);
                while (g + 1 < P.nranks && grp >= P.first_group[g + 1]) ++g;
                while (g > 0 && grp < P.first_group[g]) --g;
            }
            dest[j] = g;
            rank_in[j] = 0;
            const uint32_t peers = __match_any_sync(0xffffffffu, g);
            if (g != 0xffffffffu) {
                const uint32_t leader = (uint32_t)(__ffs(peers) - 1);
                uint32_t b0 = 0;
                if (lane == leader) b0 = atomicAdd(&s_cnt[g], (uint32_t)__popc(peers));
                b0 = __shfl_sync(peers, b0, (int)leader);
                rank_in[j] = b0 + (uint32_t)__popc(peers & ((1u << lane) - 1u));
            }
        }
        __syncthreads();
        if (tid < P.nranks) {
            const uint32_t c = s_cnt[tid];
            uint32_t b = 0;
            if (c) {
                b = atomicAdd(reinterpret_cast<uint32_t *>(P.win[tid]), c); // remote atomic over NVLink (local for own rank)
                if ((uint64_t)b + c > P.cap) {
                    reinterpret_cast<volatile uint32_t *>(P.win[tid])[1] = 1u;
                    *P.local_overflow = 1u;
                    b = 0xffffffffu;
                }
            }
            s_base[tid] = b;
        }
        if (tid == 0) {
            uint32_t o = 0;
            for (uint32_t g = 0; g < P.nranks; ++g) {
                s_off[g] = o;
                o += s_cnt[g];
            }
            s_off[P.nranks] = o;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < PUSH_ITEMS; ++j) {
            const uint32_t g = dest[j];
            if (g == 0xffffffffu) continue;
            const uint64_t i = tile + (uint64_t)j * blockDim.x + tid;
            const uint32_t pos = s_off[g] + rank_in[j];
            for (int w = 0; w < P.W; ++w) sk[(size_t)pos * P.W + w] = P.keys[i * P.W + w];
            for (int l = 0; l < P.L; ++l) scf[(size_t)pos * P.L + l] = P.cfs[i * P.L + l];
        }
        __syncthreads();
        const uint32_t nrec = s_off[P.nranks];
        for (int part = 0; part < 2; ++part) {
            const uint32_t wpr = part == 0 ? (uint32_t)P.W : (uint32_t)P.L; // words per record
            const uint64_t *src = part == 0 ? sk : scf;
            for (uint32_t q = tid; q < nrec * wpr; q += blockDim.x) {
                const uint32_t r = q / wpr, wi = q - r * wpr;
                uint32_t g = 0;
                while (g + 1 < P.nranks && r >= s_off[g + 1]) ++g;
                if (s_base[g] == 0xffffffffu) continue;
                uint64_t *wk = reinterpret_cast<uint64_t *>(P.win[g] + PUSH_HEADER);
                uint64_t *dstp = part == 0 ? wk : wk + P.cap * (uint64_t)P.W;
                dstp[((uint64_t)s_base[g] + (r - s_off[g])) * wpr + wi] = src[q];
            }
        }
        __syncthreads();
    }
}

// ================================================================================================
// Row-blocked dense path ("kernel 1").
//
// A packed key is hi * delta + lo with lo = the exponent of the first symbol (kpack.hpp:262-267: the
// first symbol is the lowest digit) and, because the exponent-overflow check guarantees digit-wise sums
// stay below delta, (hi1, lo1) * (hi2, lo2) = (hi1 + hi2, lo1 + lo2): the product of two ROWS (all terms
// sharing hi) is a 1-D convolution of their coefficient vectors that lands in the single output row
// hi1 + hi2.  For dense operands (Fateman: ~8.5 terms per row, 31 wide) this turns ~10^3 hash-table
// read-modify-writes into register accumulation: one warp owns one OUTPUT row, finds every contributing
// row pair with the homomorphic segment trick on the row keys, convolves them with warp shuffles into
// 64 exact accumulators held in registers, and writes the row once.  No atomics, no locks, no table.
// ================================================================================================

constexpr uint64_t ROW_EMPTY = 0xFFFFFFFFFFFFFFFFULL;

// Pass 1: insert hi = key / delta of every term into an open-addressing set (global memory).
__global__ void __launch_bounds__(256) k_row_insert(const uint64_t *__restrict__ keys, uint64_t n, uint64_t delta,
                                                    uint64_t *__restrict__ tab, uint32_t log2_cap)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint32_t mask = (1u << log2_cap) - 1u;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t hi = keys[i] / delta;
        uint32_t h = (uint32_t)((hi * 0x9E3779B97F4A7C15ULL) >> (64 - log2_cap));
        for (;;) {
            const unsigned long long old = atomicCAS((unsigned long long *)&tab[h], (unsigned long long)ROW_EMPTY, (unsigned long long)hi);
            if (old == ROW_EMPTY || old == hi) break;
            h = (h + 1u) & mask;
        }
    }
}

// Pass 2: occupancy flags of the set (scanned into row numbers by the host-side scan).
__global__ void __launch_bounds__(256) k_row_flags(const uint64_t *__restrict__ tab, uint32_t cap, uint32_t *__restrict__ flag)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cap) flag[i] = tab[i] != ROW_EMPTY;
}

// Pass 3: row key per row number.
__global__ void __launch_bounds__(256) k_row_keys(const uint64_t *__restrict__ tab, uint32_t cap, const uint32_t *__restrict__ rank,
                                                  uint64_t *__restrict__ row_hi)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cap && tab[i] != ROW_EMPTY) row_hi[rank[i]] = tab[i];
}

// Pass 4: scatter every term's coefficient into the dense row matrix [nrows][32] and record row lengths.
__global__ void __launch_bounds__(256) k_row_fill(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ cfs, uint64_t n,
                                                  uint64_t delta, const uint64_t *__restrict__ tab, uint32_t log2_cap,
                                                  const uint32_t *__restrict__ rank, uint64_t *__restrict__ mat,
                                                  uint32_t *__restrict__ row_len)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint32_t mask = (1u << log2_cap) - 1u;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t key = keys[i];
        const uint64_t hi = key / delta;
        const uint32_t lo = (uint32_t)(key - hi * delta);
        uint32_t h = (uint32_t)((hi * 0x9E3779B97F4A7C15ULL) >> (64 - log2_cap));
        while (tab[h] != hi) h = (h + 1u) & mask;
        const uint32_t r = rank[h];
        mat[(size_t)r * 32 + lo] = cfs[i];
        atomicMax(&row_len[r], lo + 1u);
    }
}

struct RowParams {
    // left rows (original order): key, and segment | (len-1) << 27
    const uint64_t *xr_hi;
    const uint32_t *xr_sl;
    const uint64_t *xmat; // [nrx][32]
    uint32_t nrx;
    // right rows sorted by segment: key, and original row number | (len-1) << 27; offsets [m + 1]
    const uint64_t *yr_hi;
    const uint32_t *yr_il;
    const uint32_t *yr_off;
    const uint64_t *ymat; // [nry][32], by original row number
    uint32_t nry;
    uint32_t m; // segment modulus of the row keys
    // output rows (from the symbolic row product): key and segment
    const uint64_t *or_hi;
    const uint32_t *or_seg;
    uint32_t nro;
    uint32_t *cursor;
    // dense output [nro][LACC][64] and the number of non-zero entries per row
    uint64_t *omat;
    uint32_t *or_cnt;
    uint32_t dir_in_smem; // 1: the row directories are staged in shared memory (TMA bulk copies)
    // byte sizes of the five directory arrays, each padded to a multiple of 16 (bulk-copy granularity)
    uint32_t b_xhi, b_yhi, b_yil, b_xsl, b_yoff;
};

// z += x*y for 64-bit x, y (signed, or unsigned when no coefficient is negative) and an LACC-limb accumulator.
// The 64x64->128 product is written as four 32x32->64 multiply-adds (IMAD.WIDE with a 64-bit addend) so that the
// partial products are formed once: asking for x*y and __umul64hi(x,y) separately makes the compiler form them twice.
template <int LACC, bool UNS>
__device__ __forceinline__ void row_mac(uint64_t (&z)[LACC], uint64_t x, uint64_t y)
{
    if constexpr (LACC == 1) {
        z[0] += x * y;
    } else {
        // The 64x64->128 product as ONE 128-bit multiply: the compiler lowers it to 3 IMAD.WIDE + 1 IMAD.WIDE.X with
        // the carries chained through predicates (20 SASS instructions per convolution step including the operand
        // fetch; the hand-split 32-bit form took 26, x*y next to __umul64hi 24).
        uint64_t lo, hi;
        if constexpr (UNS) {
            const unsigned __int128 p = (unsigned __int128)x * y;
            lo = (uint64_t)p;
            hi = (uint64_t)(p >> 64);
        } else {
            const __int128 p = (__int128)(long long)x * (long long)y;
            lo = (uint64_t)p;
            hi = (uint64_t)((unsigned __int128)p >> 64);
        }
        if constexpr (LACC == 2) {
            asm("add.cc.u64 %0, %0, %2;\n\taddc.u64 %1, %1, %3;" : "+l"(z[0]), "+l"(z[1]) : "l"(lo), "l"(hi));
        } else {
            const uint64_t ext = UNS ? 0ull : (uint64_t)((long long)hi >> 63);
            asm("add.cc.u64 %0, %0, %3;\n\taddc.cc.u64 %1, %1, %4;\n\taddc.u64 %2, %2, %5;"
                : "+l"(z[0]), "+l"(z[1]), "+l"(z[2])
                : "l"(lo), "l"(hi), "l"(ext));
        }
    }
}

// ---- TMA (bulk async copy) helpers: global -> shared, completion on an mbarrier --------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(
                     smem_u32(bar)),
                 "r"(parity)
                 : "memory");
}

// One warp = one output row.  The x-row and y-row directories (keys, segments, lengths, offsets) are staged in
// shared memory once per CTA by TMA bulk copies when they fit; coefficient rows stream from L2 (256 B per row,
// coalesced) and the rows of the NEXT matching pair are requested before the current pair is convolved.
template <int LACC, bool F64, bool UNS, int BT>
__global__ void __launch_bounds__(BT, 1) k_rowconv(const RowParams P)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_bar;
    // Per-warp staging of the current row pair at the start of the dynamic shared memory (double-buffered, 96 words
    // per buffer: x[32], then y TWICE so that y[(lane - i) mod 32] is the plain address yy[32 + lane - i]): the
    // convolution reads x[i] as a broadcast LDS and the rotated y as a conflict-free LDS with an immediate offset
    // instead of four SHFL (or an LDS plus three address instructions) per step.
    //
    // SHORT pairs (len1 + len2 <= 17: 57 % of the F30 pairs, 3 convolution steps on average) are not convolved one
    // by one -- their fixed cost (hit extraction, two row loads, staging, loop dispatch: ~80 instructions) is twice the
    // arithmetic.  They are queued per warp (descriptors in shared memory) and processed TWO AT A TIME: lanes 0-15
    // convolve one pair, lanes 16-31 another, each half with its own 16-wide staged rows, into a temporary
    // accumulator that is folded into the row's accumulators (lane l += half 0's l + half 1's l) once per batch.
    constexpr uint32_t PAIR_WORDS = 96, QCAP = 64;
    const uint32_t lane = threadIdx.x & 31u;
    uint64_t *wp = reinterpret_cast<uint64_t *>(smem_raw) + (size_t)(threadIdx.x >> 5) * 2u * PAIR_WORDS;
    uint64_t *wq = reinterpret_cast<uint64_t *>(smem_raw) + (size_t)(BT / 32) * 2u * PAIR_WORDS + (size_t)(threadIdx.x >> 5) * QCAP;
    uint32_t pbuf = 0;
    const uint64_t *xr_hi = P.xr_hi, *yr_hi = P.yr_hi;
    const uint32_t *xr_sl = P.xr_sl, *yr_off = P.yr_off, *yr_il = P.yr_il;
    if (P.dir_in_smem) {
        unsigned char *p = smem_raw + (size_t)(BT / 32) * (2u * PAIR_WORDS + QCAP) * 8u;
        uint64_t *s_xhi = reinterpret_cast<uint64_t *>(p);
        p += P.b_xhi;
        uint64_t *s_yhi = reinterpret_cast<uint64_t *>(p);
        p += P.b_yhi;
        uint32_t *s_yil = reinterpret_cast<uint32_t *>(p);
        p += P.b_yil;
        uint32_t *s_xsl = reinterpret_cast<uint32_t *>(p);
        p += P.b_xsl;
        uint32_t *s_yoff = reinterpret_cast<uint32_t *>(p);
        if (threadIdx.x == 0) mbar_init(&s_bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(&s_bar, P.b_xhi + P.b_yhi + P.b_yil + P.b_xsl + P.b_yoff);
            tma_bulk_g2s(s_xhi, P.xr_hi, P.b_xhi, &s_bar);
            tma_bulk_g2s(s_yhi, P.yr_hi, P.b_yhi, &s_bar);
            tma_bulk_g2s(s_yil, P.yr_il, P.b_yil, &s_bar);
            tma_bulk_g2s(s_xsl, P.xr_sl, P.b_xsl, &s_bar);
            tma_bulk_g2s(s_yoff, P.yr_off, P.b_yoff, &s_bar);
        }
        mbar_wait(&s_bar, 0);
        xr_hi = s_xhi;
        yr_hi = s_yhi;
        yr_il = s_yil;
        xr_sl = s_xsl;
        yr_off = s_yoff;
    }
    for (;;) {
        uint32_t row = 0;
        if (lane == 0) row = atomicAdd(P.cursor, 1u);
        row = __shfl_sync(0xffffffffu, row, 0);
        if (row >= P.nro) break;
        const uint64_t H = __ldg(P.or_hi + row);
        const uint32_t sg = __ldg(P.or_seg + row);
        uint64_t z0[LACC], z1[LACC];
#pragma unroll
        for (int l = 0; l < LACC; ++l) z0[l] = z1[l] = 0;
        double d0 = 0.0, d1 = 0.0;
        uint32_t qn = 0; // queued short pairs
        // two queued pairs per trip, one per half-warp; the batch's sum is folded into z0 / d0 at the end
        auto flush_short = [&]() {
            __syncwarp();
            const uint32_t half = lane >> 4, l16 = lane & 15u;
            uint64_t t[LACC];
#pragma unroll
            for (int l = 0; l < LACC; ++l) t[l] = 0;
            double dt = 0.0;
            for (uint32_t j = 0; j < qn; j += 2u) {
                const uint32_t qi = j + half;
                uint64_t X = 0, Y = 0;
                uint32_t la = 0, lb = 0;
                if (qi < qn) {
                    const uint64_t d = wq[qi];
                    const uint32_t dh = (uint32_t)(d >> 32), dl = (uint32_t)d;
                    la = (dh >> 27) + 1u;
                    lb = (dl >> 27) + 1u;
                    X = __ldg(P.xmat + (size_t)(dh & 0x07ffffffu) * 32 + l16);
                    Y = __ldg(P.ymat + (size_t)(dl & 0x07ffffffu) * 32 + l16);
                }
                if (lb < la) { // loop over the shorter row of MY pair
                    const uint64_t tt = X;
                    X = Y;
                    Y = tt;
                    la = lb;
                }
                uint64_t *pb = wp + pbuf * PAIR_WORDS + half * 48u;
                const uint64_t *px = pb, *py = pb + 32u + l16; // py[-i] = y[(l16 - i) mod 16]
                pb[l16] = X;
                pb[16u + l16] = Y;
                pb[32u + l16] = Y;
                pbuf ^= 1u;
                const uint32_t steps = __reduce_max_sync(0xffffffffu, la);
                __syncwarp();
                for (uint32_t i = 0; i < steps; ++i) {
                    const uint64_t xi = px[i];
                    const uint64_t yv = *(py - i);
                    if constexpr (F64) {
                        dt = __dadd_rn(dt, __dmul_rn(__longlong_as_double((long long)xi), __longlong_as_double((long long)yv)));
                    } else {
                        row_mac<LACC, UNS>(t, xi, yv);
                    }
                }
            }
            if constexpr (F64) {
                const double u = __longlong_as_double((long long)shfl_u64((uint64_t)__double_as_longlong(dt), (lane + 16u) & 31u));
                if (lane < 16u) d0 = __dadd_rn(d0, __dadd_rn(dt, u));
            } else {
                uint64_t u[LACC];
#pragma unroll
                for (int l = 0; l < LACC; ++l) u[l] = shfl_u64(t[l], (lane + 16u) & 31u);
                if (lane < 16u) {
                    limbs_add<LACC>(z0, t);
                    limbs_add<LACC>(z0, u);
                }
            }
            qn = 0;
            __syncwarp();
        };
        for (uint32_t base = 0; base < P.nrx; base += 32u) {
            const uint32_t r1 = base + lane;
            uint32_t found = 0xffffffffu, sl1 = 0;
            if (r1 < P.nrx) {
                const uint64_t h1 = xr_hi[r1];
                if (h1 <= H) {
                    const uint64_t target = H - h1;
                    sl1 = xr_sl[r1];
                    const uint32_t s1 = sl1 & 0x07ffffffu;
                    const uint32_t b = sg >= s1 ? sg - s1 : sg + P.m - s1;
                    const uint32_t o1 = yr_off[b + 1];
                    for (uint32_t o = yr_off[b]; o < o1; ++o) {
                        if (yr_hi[o] == target) {
                            found = yr_il[o];
                            break;
                        }
                    }
                }
            }
            const bool hit = found != 0xffffffffu;
            const bool shortp = hit && ((sl1 >> 27) + (found >> 27) + 2u <= 17u);
            const uint32_t smask = __ballot_sync(0xffffffffu, shortp);
            if (smask) {
                if (shortp) wq[qn + __popc(smask & ((1u << lane) - 1u))] = ((uint64_t)(r1 | (sl1 & 0xF8000000u)) << 32) | found;
                qn += __popc(smask);
                if (qn > QCAP - 32u) flush_short();
            }
            uint32_t mask = __ballot_sync(0xffffffffu, hit && !shortp);
            // software pipeline over the hits of this chunk: request the next pair's rows, convolve the current
            uint64_t Xn = 0, Yn = 0;
            uint32_t l1n = 0, l2n = 0;
            bool have = mask != 0;
            if (have) {
                const int src = __ffs(mask) - 1;
                mask &= mask - 1u;
                const uint32_t il = __shfl_sync(0xffffffffu, found, src);
                l1n = (__shfl_sync(0xffffffffu, sl1, src) >> 27) + 1u;
                l2n = (il >> 27) + 1u;
                Xn = __ldg(P.xmat + (size_t)(base + (uint32_t)src) * 32 + lane);
                Yn = __ldg(P.ymat + (size_t)(il & 0x07ffffffu) * 32 + lane);
            }
            while (have) {
                uint64_t X = Xn, Y = Yn;
                uint32_t len1 = l1n, len2 = l2n;
                have = mask != 0;
                if (have) {
                    const int src = __ffs(mask) - 1;
                    mask &= mask - 1u;
                    const uint32_t il = __shfl_sync(0xffffffffu, found, src);
                    l1n = (__shfl_sync(0xffffffffu, sl1, src) >> 27) + 1u;
                    l2n = (il >> 27) + 1u;
                    Xn = __ldg(P.xmat + (size_t)(base + (uint32_t)src) * 32 + lane);
                    Yn = __ldg(P.ymat + (size_t)(il & 0x07ffffffu) * 32 + lane);
                }
                if (len2 < len1) { // the convolution is symmetric: loop over the shorter row
                    const uint64_t t = X;
                    X = Y;
                    Y = t;
                    const uint32_t tl = len1;
                    len1 = len2;
                    len2 = tl;
                }
                uint64_t *pb = wp + pbuf * PAIR_WORDS;
                const uint64_t *px = pb, *py = pb + 64u + lane; // py[-i] = y[(lane - i) mod 32]
                pb[lane] = X;
                pb[32u + lane] = Y;
                pb[64u + lane] = Y;
                pbuf ^= 1u;
                __syncwarp(); // the other buffer is free: every lane passed this point after finishing the pair before last
                if (len1 + len2 <= 33u) {
                    // the output row fits 32 lanes: the wrapped-around lanes of the rotation read zeros of the
                    // longer row, so every lane accumulates unconditionally into its low output
#pragma unroll 4
                    for (uint32_t i = 0; i < len1; ++i) {
                        const uint64_t xi = px[i];
                        const uint64_t yv = *(py - i);
                        if constexpr (F64) {
                            d0 = __dadd_rn(d0, __dmul_rn(__longlong_as_double((long long)xi), __longlong_as_double((long long)yv)));
                        } else {
                            row_mac<LACC, UNS>(z0, xi, yv);
                        }
                    }
                } else {
#pragma unroll 4
                    for (uint32_t i = 0; i < len1; ++i) {
                        const uint64_t xi = px[i];
                        const uint64_t yv = *(py - i);
                        // lane >= i: the rotated value belongs to output lo = lane, else to output lo = lane + 32;
                        // mask the INPUT so both accumulators run branch-free multiply-adds
                        const bool lowhalf = lane >= i;
                        if constexpr (F64) {
                            const double p = __dmul_rn(__longlong_as_double((long long)xi), __longlong_as_double((long long)yv));
                            if (lowhalf) d0 = __dadd_rn(d0, p); else d1 = __dadd_rn(d1, p);
                        } else {
                            const uint64_t y0 = lowhalf ? yv : 0ull, y1 = lowhalf ? 0ull : yv;
                            row_mac<LACC, UNS>(z0, xi, y0);
                            row_mac<LACC, UNS>(z1, xi, y1);
                        }
                    }
                }
            }
        }
        if (qn) flush_short();
        // write the row once; count its non-zero entries
        uint64_t *o = P.omat + (size_t)row * 64 * LACC;
        bool nz0, nz1;
        if constexpr (F64) {
            o[lane] = (uint64_t)__double_as_longlong(d0);
            o[32 + lane] = (uint64_t)__double_as_longlong(d1);
            nz0 = d0 != 0.0;
            nz1 = d1 != 0.0;
        } else {
            nz0 = nz1 = false;
#pragma unroll
            for (int l = 0; l < LACC; ++l) {
                o[(size_t)l * 64 + lane] = z0[l];
                o[(size_t)l * 64 + 32 + lane] = z1[l];
                nz0 |= z0[l] != 0;
                nz1 |= z1[l] != 0;
            }
        }
        const uint32_t c = __popc(__ballot_sync(0xffffffffu, nz0)) + __popc(__ballot_sync(0xffffffffu, nz1));
        if (lane == 0) P.or_cnt[row] = c;
    }
}

// Directory packing: segment | (len-1) << 27 for the left rows, row number | (len-1) << 27 for the sorted right rows.
__global__ void __launch_bounds__(256) k_row_pack(const uint32_t *__restrict__ xseg, const uint32_t *__restrict__ xlen, uint32_t nrx,
                                                  const uint64_t *__restrict__ yidx, const uint32_t *__restrict__ ylen, uint32_t nry,
                                                  uint32_t *__restrict__ xsl, uint32_t *__restrict__ yil)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nrx) xsl[i] = xseg[i] | ((xlen[i] - 1u) << 27);
    if (i < nry) {
        const uint32_t r = (uint32_t)yidx[i];
        yil[i] = r | ((ylen[r] - 1u) << 27);
    }
}

// Expand the dense output rows into terms: key = hi * delta + lo, zero coefficients dropped.
template <int LACC, bool F64>
__global__ void __launch_bounds__(256) k_row_expand(const uint64_t *__restrict__ or_hi, const uint64_t *__restrict__ omat,
                                                    const uint32_t *__restrict__ off, uint32_t nro, uint64_t delta,
                                                    uint64_t *__restrict__ okeys, uint64_t *__restrict__ ocfs)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t wpb = blockDim.x >> 5;
    for (uint32_t row = blockIdx.x * wpb + (threadIdx.x >> 5); row < nro; row += gridDim.x * wpb) {
        const uint64_t *o = omat + (size_t)row * 64 * LACC;
        uint32_t base = off[row];
        const uint64_t hi = or_hi[row];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint64_t v[LACC];
            bool nz = false;
#pragma unroll
            for (int l = 0; l < LACC; ++l) {
                v[l] = o[(size_t)l * 64 + half * 32 + lane];
                nz |= v[l] != 0;
            }
            if constexpr (F64) nz = __longlong_as_double((long long)v[0]) != 0.0;
            const uint32_t b = __ballot_sync(0xffffffffu, nz);
            if (nz) {
                const uint32_t pos = base + __popc(b & ((1u << lane) - 1u));
                okeys[pos] = hi * delta + (uint64_t)(half * 32 + lane);
#pragma unroll
                for (int l = 0; l < LACC; ++l) ocfs[(size_t)pos * LACC + l] = v[l];
            }
            base += __popc(b);
        }
    }
}

} // namespace obk
