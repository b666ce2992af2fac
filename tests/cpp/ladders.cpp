// Host-side check of the bin ladders in csrc/mhb_config.h (compiled with g++, no CUDA):
// every row lands in a bin whose kernel can hold it -- hash tables are never filled beyond
// 5/8 (the bucket-sort scratch and the claim list rely on it), windows / bitmaps fit the
// shared memory their launch requests, and the forced paths only move rows between families.
#include <cstdio>
#include <cstdlib>
#include <initializer_list>
#include "../../mh-spgemm_b200/csrc/mhb_config.h"

#define CHECK(c)                                                                                    \
    do                                                                                              \
    {                                                                                               \
        if (!(c))                                                                                   \
        {                                                                                           \
            std::printf("FAIL %s (line %d) n=%d ip=%d w=%d\n", #c, __LINE__, n, ip, w);             \
            return 1;                                                                               \
        }                                                                                           \
    } while (0)

int main()
{
    const int spans[] = {1, 24, 64, 65, 256, 257, 1024, 1025, 6144, 6145, 27648, 27649, 1 << 20, 1 << 30};
    const int sizes[] = {1, 2, 16, 24, 25, 80, 81, 160, 161, 320, 321, 448, 449, 640, 641, 2560, 2561, 10240, 10241, 1 << 20};
    long long checked = 0;
    for (int w : spans)
        for (int n : sizes)
            for (int mult : {1, 2, 40})
                for (int force : {0, 1, 2})
                {
                    if (n > w)
                        continue;
                    const int ip = n * mult, cmin = 7, cmax = cmin + w - 1;
                    const int b = mhb_classify_num(n, ip, cmin, cmax, force, ip, 0, 1, 8);
                    ++checked;
                    CHECK(b > NB_EMPTY && b < NB_COUNT);
                    switch (b)
                    {
                    case NB_TINY: CHECK(force == 0 && n <= NB_TINY_MAX && ip <= NB_TINY_PRODUCTS && ip > NB_TINY_M_PRODUCTS); break;
                    case NB_TINY_M: CHECK(force == 0 && n <= NB_TINY_MAX && ip <= NB_TINY_M_PRODUCTS && ip > NB_TINY_S_PRODUCTS); break;
                    case NB_TINY_S: CHECK(force == 0 && n <= NB_TINY_MAX && ip <= NB_TINY_S_PRODUCTS); break;
                    case NB_WIN_G8: CHECK(w <= NB_WIN_G8_COLS); break;
                    case NB_WIN_WARP: CHECK(w <= NB_WIN_WARP_COLS); break;
                    case NB_WIN_COMPACT: CHECK(w <= NB_WIN_WARP_COLS && n <= NB_WIN_COMPACT_MAXN); break;
                    case NB_WIN_BLOCK_S: CHECK(w <= NB_WIN_BLOCK_S_COLS); break;
                    case NB_WIN_BLOCK_L: CHECK(w <= NB_WIN_BLOCK_L_COLS && (size_t)w * 8 + w / 32 * 8 <= MHB_SMEM_MAX); break;
                    case NB_H_G8: CHECK(n <= NB_H_G8_MAX && n * 8 <= NB_H_G8_SLOTS * 6); break;
                    case NB_H_WARP_XS: CHECK(n <= NB_H_WARP_XS_MAX && n * 8 <= NB_H_WARP_XS_SLOTS * 5); break;
                    case NB_H_WARP_S: CHECK(n <= NB_H_WARP_S_MAX && n * 8 <= NB_H_WARP_S_SLOTS * 5); break;
                    case NB_H_WARP_M: CHECK(n <= NB_H_WARP_M_MAX && n * 8 <= NB_H_WARP_M_SLOTS * 5); break;
                    case NB_H_WARP_L: CHECK(n <= NB_H_WARP_L_MAX && n * 8 <= NB_H_WARP_L_SLOTS * 5); break;
                    case NB_H_BLOCK_XS: CHECK(n <= NB_H_BLOCK_XS_MAX && n * 8 <= NB_H_BLOCK_XS_SLOTS * 5); break;
                    case NB_H_BLOCK_S: CHECK(n <= NB_H_BLOCK_S_MAX && n * 8 <= NB_H_BLOCK_S_SLOTS * 5); break;
                    case NB_H_BLOCK_M: CHECK(n <= NB_H_BLOCK_M_MAX && n * 8 <= NB_H_BLOCK_M_SLOTS * 5); break;
                    case NB_H_BLOCK_L: CHECK(n <= NB_H_BLOCK_L_MAX && n * 8 <= NB_H_BLOCK_L_SLOTS * 5); break;
                    case NB_H_GLOBAL: CHECK(n > NB_H_BLOCK_L_MAX); break;
                    default: CHECK(false);
                    }
                    if (force == 2)
                        CHECK(b >= NB_H_G8 && b != NB_WIN_COMPACT && b != NB_TINY && b != NB_TINY_S && b != NB_TINY_M && b != NB_WIN_G8 && b != NB_WIN_WARP &&
                              b != NB_WIN_BLOCK_S && b != NB_WIN_BLOCK_L);
                    // symbolic: tile-flop tf = ip, words spanned wt
                    const int sb = mhb_classify_sym(ip, ip, cmin, cmax, force);
                    const long long wt = (long long)(cmax >> MHB_TILE_SHIFT) - (cmin >> MHB_TILE_SHIFT) + 1;
                    const long long ub = wt < ip ? wt : ip;
                    CHECK(sb > SB_EMPTY && sb < SB_COUNT);
                    switch (sb)
                    {
                    case SB_TINY: CHECK(force == 0 && ip <= SB_TINY_MAX && ip > SB_TINY_M_MAX); break;
                    case SB_TINY_M: CHECK(force == 0 && ip <= SB_TINY_M_MAX && ip > SB_TINY_S_MAX); break;
                    case SB_TINY_S: CHECK(force == 0 && ip <= SB_TINY_S_MAX); break;
                    case SB_BM_G8: CHECK(wt <= SB_BM_G8_WORDS); break;
                    case SB_BM_WARP: CHECK(wt <= SB_BM_WARP_WORDS); break;
                    case SB_BM_BLOCK: CHECK(wt <= SB_BM_BLOCK_WORDS && wt * 4 <= MHB_SMEM_MAX); break;
                    case SB_H_G8: CHECK(ub <= SB_H_G8_MAX && ub * 4 <= SB_H_G8_SLOTS * 3); break;
                    case SB_H_G16: CHECK(ub <= SB_H_G16_MAX && ub * 4 <= SB_H_G16_SLOTS * 3); break;
                    case SB_H_WARP: CHECK(ub <= SB_H_WARP_MAX && ub * 4 <= SB_H_WARP_SLOTS * 3); break;
                    case SB_H_BLOCK_S: CHECK(ub <= SB_H_BLOCK_S_MAX && ub * 4 <= SB_H_BLOCK_S_SLOTS * 3); break;
                    case SB_H_BLOCK_L: CHECK(ub <= SB_H_BLOCK_L_MAX && ub * 4 <= SB_H_BLOCK_L_SLOTS * 3); break;
                    case SB_H_GLOBAL: CHECK(ub > SB_H_BLOCK_L_MAX); break;
                    default: CHECK(false);
                    }
                }
    int n = 0, ip = 0, w = 0;
    CHECK(mhb_classify_num(0, 0, 0, 0, 0) == NB_EMPTY);
    CHECK(mhb_classify_sym(0, 0, 0, 0, 0) == SB_EMPTY);
    CHECK(NB_COUNT <= MHB_MAX_BINS && SB_COUNT <= MHB_MAX_BINS);
    std::printf("OK %lld combinations\n", checked);
    return 0;
}
