/* obk_b200.h -- C ABI of the B200-native series-product engine (libobk_b200.so).
 *
 * This is the drop-in boundary for ONE path of bluescarni/obake: the product of two sparse
 * multivariate polynomials whose monomials are Kronecker-packed integers.  Nothing like this ABI
 * exists in the reference (it is a header-only C++20 template library); every entry point below
 * names the reference interface it replaces (file:line under the obake source tree) and
 * INTEGRATION.md shows the binding a maintainer adds on the obake side (a C++ overload of
 * obake::customisation::series_mul / the body of polynomials::detail::poly_mul_impl_identical_ss).
 *
 * Conventions
 *  - plain C types only; no torch / CUDA types in the signatures (a CUDA stream travels as void*).
 *  - keys: `key_words` 64-bit words per term, term-major.  One word = packed_monomial<T>
 *    (include/obake/polynomials/packed_monomial.hpp:56-142), several = d_packed_monomial<T,psize>
 *    (d_packed_monomial.hpp:83-217).  32-bit key types are widened (sign- or zero-extended) by the
 *    caller; signed keys travel as their two's-complement bit pattern.
 *  - integer coefficients: `cf_limbs` 64-bit limbs per term, two's complement, little endian --
 *    the value of an mppp::integer<N>.  fp64 coefficients: one double per term.
 *  - `mem` says where `keys`/`cfs` live: host memory (pinned or pageable) or this engine's device.
 *  - every call is synchronous with respect to the engine's stream when it returns.
 */
#ifndef OBK_B200_H
#define OBK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct obk_engine obk_engine;

typedef enum obk_status {
    OBK_OK = 0,
    OBK_ERR_INVALID_ARG = 1,
    /* std::overflow_error "An overflow in the monomial exponents was detected while attempting to
     * multiply two polynomials" -- polynomial.hpp:839-851 */
    OBK_ERR_KEY_OVERFLOW = 2,
    /* std::overflow_error, table larger than the series allows -- polynomial.hpp:1544-1550 */
    OBK_ERR_TABLE_OVERFLOW = 3,
    /* the exact result needs more than 5 accumulator limbs = 320 bits (the reference would spill to GMP) */
    OBK_ERR_CF_OVERFLOW = 4,
    OBK_ERR_CUDA = 5,
    OBK_ERR_UNSUPPORTED = 6,
    OBK_ERR_NOMEM = 7
} obk_status;

enum { OBK_CF_INT = 0, OBK_CF_F64 = 1 };
enum { OBK_MEM_HOST = 0, OBK_MEM_DEVICE = 1 };

/* One operand: the terms of a polynomial<Key, Cf> as SoA arrays (what poly_mul_impl_mt_hm copies
 * out of the series at polynomial.hpp:828-833). */
typedef struct obk_poly_view {
    uint64_t nterms;
    uint32_t key_words;  /* W >= 1 */
    uint32_t key_signed; /* 1: T is a signed integer type */
    uint32_t nvars;      /* size of the symbol set */
    uint32_t psize;      /* 0: packed_monomial (one word packs nvars exponents); k: d_packed_monomial<T,k> */
    uint32_t cf_kind;    /* OBK_CF_INT | OBK_CF_F64 */
    uint32_t cf_limbs;   /* integer: 1..5 limbs carried, values of at most 3 limbs (192 bits) as product operands; f64: 1 */
    uint32_t mem;        /* OBK_MEM_HOST | OBK_MEM_DEVICE */
    uint32_t key_bits;   /* width of T: 64 (or 0) | 32.  Selects the Kronecker tables (delta, limits); 32-bit
                          * keys still travel widened to 64-bit words */
    const uint64_t *keys; /* [nterms * key_words] */
    const void *cfs;      /* [nterms * cf_limbs] uint64 or [nterms] double */
} obk_poly_view;

/* Truncation arguments of truncated_mul (polynomial.hpp:2113-2127): pair (i,j) is kept iff
 * !(limit < deg_i + deg_j), degrees taken in the key's integer type T (polynomial.hpp:604-610). */
typedef struct obk_trunc {
    int64_t limit;       /* already converted to T (bit pattern for unsigned T) */
    uint32_t nidx;       /* 0: total degree; else number of entries of idx */
    uint32_t partial;    /* 1: partial degree over idx (idx may be empty: every degree is 0) */
    const uint32_t *idx; /* sorted positions (in the symbol set) of the variables that count (host) */
} obk_trunc;

enum { OBK_SEG_HASH_POW2 = 0, OBK_SEG_KEY_MOD = 1 };

/* The product, grouped by segment: the terms of group s sit in [seg_offsets[s], seg_offsets[s+1]).
 * seg_kind == OBK_SEG_HASH_POW2 (every product returned in HOST memory, or option regroup = 2): nsegs =
 * 2^log2_nsegs and a term of group s satisfies hash(key) & (nsegs - 1) == s, i.e. it belongs in table s of a
 * series with that many segments (series.hpp:396-400,1294-1305), so a host series can be rebuilt table-parallel.
 * seg_kind == OBK_SEG_KEY_MOD (products left in DEVICE memory, partial products of the multi-GPU path): group s
 * holds the keys with sum_w(word_w mod nsegs) mod nsegs == s; nsegs = 1 for a product left on the device (both
 * kernels write their terms straight to the final arrays in no particular order; `keys`/`cfs` may be allocated
 * larger than nterms), 4093 for multi-GPU partials.
 * No coefficient is zero (polynomial.hpp:1528-1540). */
typedef struct obk_poly_result {
    uint64_t nterms;
    uint32_t key_words;
    uint32_t cf_kind;
    uint32_t cf_limbs;    /* integer: accumulator limbs (1, 2, 3 or 5), two's complement */
    uint32_t log2_nsegs;
    uint32_t mem;         /* where keys/cfs live */
    uint32_t nsegs;       /* number of groups */
    uint32_t seg_kind;
    uint32_t reserved;
    uint64_t *keys;
    void *cfs;
    uint64_t *seg_offsets; /* host memory, nsegs + 1 entries */
    void *owner;           /* engine bookkeeping; release with obk_result_free */
} obk_poly_result;

typedef struct obk_stats {
    uint64_t term_products; /* tot_n_mults, polynomial.hpp:574-623 */
    uint64_t nterms_out;
    uint64_t est_nterms;    /* sampled estimate that sized the segments */
    uint64_t h2d_bytes, d2h_bytes;
    uint32_t log2_nsegs;    /* grouping of the returned terms (hash & (2^k - 1)) */
    uint32_t nsegs;         /* the engine's segment count: a prime modulus m, one CTA-owned table per segment */
    uint32_t acc_limbs;
    uint32_t retries;       /* segment-table overflows that forced a finer segmentation */
    uint32_t kernel;        /* 0: scatter (segment-owner hash tables), 1: row-blocked dense path */
    uint32_t table_slots;   /* slots of one shared-memory segment table */
    uint32_t group_lanes;   /* scatter kernel: slices each left term's run is cut into (work unit = 32 left terms x
                             * one slice); row kernel: 32 (one warp per output row) */
    uint32_t mul_launches;  /* launches of the product kernel (1 + retries) */
    uint32_t total_launches;/* all kernels launched by this call */
    uint32_t reserved;
    float ms_total;         /* device time of the whole call (first to last kernel, CUDA events) */
    float ms_prep;          /* stats + degree + bucket sort */
    float ms_estimate;
    float ms_mul;           /* the product kernel alone (last attempt) */
    float ms_compact;
    float ms_h2d, ms_d2h;
} obk_stats;

/* Engine lifetime.  One engine = one CUDA device + one stream + a grow-only workspace. */
obk_status obk_engine_create(int device, obk_engine **out);
/* Single-process multi-GPU engine (obake is a single-process library: `f * g`, series.hpp:3186-3194, must be able to use
 * the whole box without a launcher).  obk_poly_mul on such an engine takes HOST operands and returns a HOST result: one
 * host thread per device, both operands on every device, each device forms the output segments / planes it owns (the
 * zero-exchange design of SURVEY 8e -- ownership is a function of the keys, no term crosses NVLink), regroups its
 * share into the same 2^k series tables and copies it back; the shares of a table are concatenated on the host.  The
 * O(terms) entry points (addsub, truncate, copy, extend, degrees, overflow check) run on devices[0]; the partial /
 * push / merge entry points and set_stream belong to the one-process-per-GPU design and are refused. */
obk_status obk_engine_create_multi(const int *devices, int ndev, obk_engine **out);
void obk_engine_destroy(obk_engine *e);
/* Launch on the caller's stream (e.g. torch's current stream) instead of the engine's own. */
obk_status obk_engine_set_stream(obk_engine *e, void *cuda_stream);
/* Tunables (development / tests): "nsegs" (force the segment modulus, -1 = auto), "table_slots", "threads",
 * "refill_min" (idle lanes of a warp that trigger the next hand-out of products), "fill_pct", "ctas_per_sm",
 * "max_retries",
 * "kernel"    -1 auto, 0 segment-owner scatter, 1 row-blocked dense path (if the operands qualify),
 * "row_threads" 512 | 1024, "two_sided" 0 off, 1 auto, 2 always (truncated products),
 * "fence"     1: fence.acq_rel.cta around the slot lock (memory-model form of the protocol),
 * "regroup"   1 (default): series-table grouping for HOST results only; 2: always; 0: never. */
obk_status obk_engine_set_option(obk_engine *e, const char *name, int64_t value);
const char *obk_last_error(const obk_engine *e);
const char *obk_status_string(obk_status s);
const char *obk_version(void);

/* series_mul(polynomial, polynomial) / truncated_mul (polynomial.hpp:2026-2032, 2113-2127) from the
 * seam poly_mul_impl_identical_ss (polynomial.hpp:1878-1929): empty operand -> empty result, shorter
 * operand first (:2014-2022), exponent-overflow check before any product (:839-851), early empty
 * result when the truncation admits no pair (:860-862), then the segmented product (:797-1582).
 * `t` may be NULL (untruncated).  `result_mem` selects where the product is left. */
obk_status obk_poly_mul(obk_engine *e, const obk_poly_view *x, const obk_poly_view *y, const obk_trunc *t,
                        uint32_t result_mem, obk_poly_result *out, obk_stats *stats);
void obk_result_free(obk_engine *e, obk_poly_result *r);

/* monomial_range_overflow_check (packed_monomial.hpp:289-488, d_packed_monomial.hpp:583-897).
 * *ok = 1 when every product monomial is representable. */
obk_status obk_overflow_check(obk_engine *e, const obk_poly_view *x, const obk_poly_view *y, int *ok);

/* make_degree_vector / make_p_degree_vector (series.hpp:3757-3785, 3911-3945) with key_degree /
 * key_p_degree (src/polynomials/packed_monomial.cpp:334-350,388-412; d_packed_monomial.hpp:901-958).
 * out_host: nterms int64 (bit pattern of T). */
obk_status obk_key_degrees(obk_engine *e, const obk_poly_view *x, const obk_trunc *which, int64_t *out_host);

/* Multi-GPU: rank r multiplies its slice of the shorter operand's terms by the whole other
 * operand and leaves a DEVICE partial result grouped by key mod 4093 (seg_kind OBK_SEG_KEY_MOD, a fixed
 * prime every rank uses).  send_counts[g] = number of partial terms owned by rank g, i.e. those of groups
 * [g*nsegs/G, (g+1)*nsegs/G) -- contiguous rows.  After the all-to-all the owner merges what it received
 * with obk_poly_merge.  The exponent-overflow check and the accumulator width use the FULL operands, so
 * every rank takes the same decisions and produces the same limb count. */
obk_status obk_poly_mul_partial(obk_engine *e, const obk_poly_view *x, const obk_poly_view *y, const obk_trunc *t,
                                uint32_t rank, uint32_t nranks, obk_poly_result *out_partial,
                                uint64_t *send_counts, obk_stats *stats);
/* Sum terms with equal keys (exact / fp64), drop zeros; `terms` holds acc-limb coefficients
 * (cf_limbs = accumulator limbs).  `nsegs` is a hint (0 = auto): the merge sizes its own segment tables
 * from the bound distinct <= nterms. */
obk_status obk_poly_merge(obk_engine *e, const obk_poly_view *terms, uint32_t nsegs, uint32_t result_mem,
                          obk_poly_result *out, obk_stats *stats);

/* ---- multi-GPU exchange fused into the producer (NVLink peer memory) -------------------------------------------
 * Instead of returning its partial product, a rank writes every partial term straight into the receive window of
 * the rank that owns it (P2P stores over NVLink, the slot range reserved with one remote atomic per destination and
 * CTA).  A window is 2 halves ("parities", alternated step by step so that a fast rank may already push step k+1
 * while a slow one still merges step k) of `window_bytes` each: [0] term counter, [4] overflow flag, records from
 * byte 256.  How the peers' windows get mapped into this process is the caller's business (obk_peer_window_alloc /
 * _open use CUDA IPC; any symmetric-memory allocator works): the engine only sees device pointers.
 * Protocol per step: obk_poly_mul_push on every rank; one stream-ordered collective on the engine's stream as the
 * barrier (e.g. an all-reduce of one word); obk_peer_merge on every rank. */
typedef struct obk_peer_windows {
    uint32_t nranks;
    uint32_t rank;
    uint32_t parity;       /* 0 | 1: the half used by this step */
    uint32_t reserved;
    uint64_t window_bytes; /* bytes of ONE half */
    void *const *base;     /* nranks device pointers: rank g's window as mapped in this process (base[rank] = own) */
} obk_peer_windows;

/* cudaMalloc'ed window of 2 * window_bytes, zeroed headers, and its 64-byte CUDA IPC handle. */
obk_status obk_peer_window_alloc(obk_engine *e, uint64_t window_bytes, void **base, void *ipc_handle64);
obk_status obk_peer_window_open(obk_engine *e, const void *ipc_handle64, void **base);
obk_status obk_peer_window_close(obk_engine *e, void *base, int own);
/* obk_poly_mul_partial + push of the partial terms to their owners; stats->acc_limbs is the limb count pushed. */
obk_status obk_poly_mul_push(obk_engine *e, const obk_poly_view *x, const obk_poly_view *y, const obk_trunc *t,
                             const obk_peer_windows *w, obk_stats *stats);
/* Sum what arrived in this rank's window (same as obk_poly_merge on the received records) and clear the window's
 * header for its next use.  `like` supplies the key layout and coefficient kind, `acc_limbs` the pushed limb count. */
obk_status obk_peer_merge(obk_engine *e, const obk_peer_windows *w, const obk_poly_view *like, uint32_t acc_limbs,
                          uint32_t result_mem, obk_poly_result *out, obk_stats *stats);

/* Error path of the protocol: when ANY rank's obk_poly_mul_push failed, no rank calls obk_peer_merge for that step;
 * every rank instead clears the header of its own window (this step's parity) after the barrier, so that what the
 * peers had already pushed is not merged when the parity is reused two steps later. */
obk_status obk_peer_window_reset(obk_engine *e, const obk_peer_windows *w);

/* ---- the steps either side of the product, on the same device layout ------------------------------
 * Operands and results may stay in DEVICE memory, so a chain such as f *= g; h = f + 1; truncate(h) never leaves
 * HBM (benchmark/dense.hpp:28-32, rectangular_01.cpp:42-47, audi_01.cpp:32-42). */

/* Copy a polynomial's terms into engine-owned HOST or DEVICE memory: the upload that starts a device-resident
 * chain (what poly_mul_impl_mt_hm's term copy, polynomial.hpp:828-833, does once per product in the reference) and
 * the download that ends it (a HOST copy arrives grouped the way a series stores its terms). */
obk_status obk_poly_copy(obk_engine *e, const obk_poly_view *x, uint32_t result_mem, obk_poly_result *out);

/* Symbol-set extension (series_sym_extender, series.hpp:539-625, with key_merge_symbols,
 * src/polynomials/packed_monomial.cpp:234-292 / d_packed_monomial.hpp:473-520): re-express x over a larger symbol
 * set of `new_nvars` symbols; pos[j] (strictly increasing, host memory, x->nvars entries) is the position of x's
 * j-th symbol in the new set, the new symbols get exponent 0.  Every key is re-packed with the Kronecker deltas of
 * the new size; an exponent outside the new limits fails with OBK_ERR_KEY_OVERFLOW (the reference's kpacker throws
 * std::overflow_error), too many symbols for one packed word with OBK_ERR_INVALID_ARG. */
obk_status obk_poly_extend_symbols(obk_engine *e, const obk_poly_view *x, uint32_t new_nvars, const uint32_t *pos,
                                   uint32_t result_mem, obk_poly_result *out);

/* Series addition / subtraction of two polynomials over the SAME symbol set (series.hpp:2195-2991, the
 * identical-symbol-set branch of series_default_addsub_impl): out = x + y or x - y, exact for integers (the result
 * uses as many limbs as the sum needs: 1, 2, 3 or 5), terms that cancel are erased. */
obk_status obk_poly_addsub(obk_engine *e, const obk_poly_view *x, const obk_poly_view *y, int subtract, uint32_t result_mem,
                           obk_poly_result *out, obk_stats *stats);

/* truncate_degree(x, limit) / truncate_p_degree(x, limit, symbols) (polynomial.hpp:2364-2445): keeps the terms
 * with !(limit < degree), the degree taken in the key's integer type T as in obk_trunc. */
obk_status obk_poly_truncate_degree(obk_engine *e, const obk_poly_view *x, const obk_trunc *t, uint32_t result_mem,
                                    obk_poly_result *out, obk_stats *stats);

/* Measurement support (bench.py, roofline.int_mac): "int_mac" = exact 64 x 64 -> 160-bit multiply-accumulates per second
 * of this GPU with register-resident operands (the accumulator form of the plane-blocked dense kernel, 4 IMAD.WIDE + 2
 * IADD3.X per product) -- the compute ceiling of the dense paths, which move almost no HBM bytes. */
obk_status obk_microbench(obk_engine *e, const char *name, double *value);

#ifdef __cplusplus
}
#endif
#endif /* OBK_B200_H */
