// Host-side invariants of the tile path's geometry helpers (obk_tiles.cuh): tile enumeration, tile groups, stage layout.
// Compiled with nvcc, runs on the CPU (no kernel is launched).
#include <cstdio>
#include <cstdlib>

#include "../../obake_b200/csrc/obk_tiles.cuh"

using namespace obk;

static int g_fail = 0, g_checks = 0;
#define CHECK(c)                                                                                                       \
    do {                                                                                                               \
        ++g_checks;                                                                                                    \
        if (!(c)) {                                                                                                    \
            ++g_fail;                                                                                                  \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c);                                                 \
        }                                                                                                              \
    } while (0)

int main()
{
    // pitch: multiple of 8 words with P / 8 odd, at least 8 words beyond the widest row
    for (uint32_t w = 1; w <= (uint32_t)TL_MAX_DIM; ++w) {
        const uint32_t P = tl_pitch(w);
        CHECK(P % 8u == 0u && (P / 8u) % 2u == 1u && P >= w + 8u);
    }
    for (uint32_t rows = 1; rows <= 127u; rows += 3u) {
        for (uint32_t w = 1; w <= 127u; w += 5u) {
            for (uint32_t d : {1u, 2u, 7u, 8u, 9u, 31u, 64u, 127u, 200u, 253u}) {
                const uint32_t shape = tl_shape(rows, w, d);
                const uint32_t nt = tl_ntiles(rows, w, d);
                // tiles by brute force: tile (j0, c0) exists iff it holds a position with j < rows, l < w, j + l < d
                uint32_t brute = 0;
                for (uint32_t j0 = 0; j0 < rows; j0 += TL_R)
                    for (uint32_t c0 = 0; c0 < w; c0 += TL_C)
                        if (j0 + c0 < d) ++brute;
                CHECK(nt == brute);
                // the rows of tile row tr add up, tl_locate inverts the enumeration
                uint32_t sum = 0, idx = 0;
                for (uint32_t tr = 0; tr < 40u; ++tr) {
                    const uint32_t n = tl_row_tiles(shape, tr);
                    for (uint32_t t = 0; t < n; ++t, ++idx) {
                        uint32_t j0 = 99999, c0 = 99999;
                        CHECK(tl_locate(idx, rows, w, d, j0, c0) && j0 == tr * TL_R && c0 == t * TL_C);
                    }
                    sum += n;
                }
                uint32_t j0, c0;
                CHECK(sum == nt && !tl_locate(nt, rows, w, d, j0, c0));
                // groups: whole tile rows, at most `slots` tiles each, together all tiles, in order
                for (uint32_t slots : {16u, 23u, 62u, 128u}) {
                    const uint32_t ng = tl_ngroups(shape, slots);
                    uint32_t tot = 0, next_tr = 0;
                    for (uint32_t g = 0; g < ng; ++g) {
                        uint32_t lo, hi;
                        const uint32_t n = tl_group(shape, slots, g, lo, hi);
                        CHECK(n > 0u && n <= slots && lo == next_tr && hi > lo);
                        uint32_t m = 0;
                        for (uint32_t tr = lo; tr < hi; ++tr) m += tl_row_tiles(shape, tr);
                        CHECK(m == n);
                        tot += n;
                        next_tr = hi;
                    }
                    uint32_t lo, hi;
                    CHECK(tot == nt && tl_group(shape, slots, ng, lo, hi) == 0u && (nt == 0u) == (ng == 0u));
                }
            }
        }
    }
    // stage layout: regions in order, 16-byte aligned (bulk copies), zero padding above the vector plane
    for (uint32_t rcap = 1; rcap <= (uint32_t)TL_MAX_DIM; rcap += 7u) {
        for (uint32_t w : {1u, 16u, 31u, 33u, 64u}) {
            const uint32_t P = tl_pitch(w), P8 = P * 8u;
            const TileStage G = tl_stage(rcap, P);
            CHECK(G.s_rows == (uint32_t)TL_HDR_BYTES && G.v_hdr == G.s_rows + rcap * P8);
            CHECK(G.v_rows == G.v_hdr + TL_HDR_BYTES + TL_LEAD_BYTES + TL_PAD_ROWS * P8);
            CHECK(G.desc == G.v_rows + (rcap + TL_PAD_ROWS) * P8 && G.bytes >= G.desc + TL_DESC_BYTES && G.bytes % 128u == 0u);
            CHECK(G.s_rows % 16u == 0u && G.v_hdr % 16u == 0u && G.v_rows % 16u == 0u && G.desc % 16u == 0u);
            CHECK(tl_blob_words(rcap, P) == TL_HDR + (rcap + TL_PAD_ROWS) * P);
        }
    }
    std::printf("%s: %d checks, %d failures\n", g_fail ? "FAILED" : "OK", g_checks, g_fail);
    return g_fail ? 1 : 0;
}
