// mhb_config.h -- bin ladders and capacities shared by host and device code.
//
// The reference routes rows to kernels through hard-coded threshold tables sized for a
// 99 KB shared-memory block (inc/binning.cuh:5-63, inc/common.h:20-36).  These ladders
// are re-derived for sm_100a: 227 KB of shared memory per block lets a whole-row bitmap
// (symbolic) or a dense fp64 window (numeric) replace hashing whenever the column span of
// a C row is small relative to its work, and lets the largest in-SMEM hash table hold
// 16 K (col,val) slots instead of 8 447.
#pragma once
#include <stdint.h>

#define MHB_TILE_SHIFT 5 // 32-column tiles, as inc/common.h:74-75
#define MHB_TILE_BITS 32

#define MHB_MAX_BINS 24
#define MHB_SMEM_MAX 232448 // 227 KB opt-in dynamic shared memory per block on sm_100

// ---- symbolic bins (family 3). Wt = 32-column words spanned by the C row, tf = tile-flop,
//      ub = min(tf, Wt) = upper bound on distinct C tiles of the row. ----
enum MhbSymBin
{
    SB_EMPTY = 0,   // no intermediate products: nnz = 0
    SB_BM_G8,       // bitmap, 8 lanes/row,  Wt <= 64
    SB_BM_WARP,     // bitmap, warp/row,     Wt <= 2048
    SB_BM_BLOCK,    // bitmap, block/row,    Wt <= 57344 (224 KB)
    SB_H_G8,        // tile hash, 8 lanes/row, ub <= 24   (32 slots)
    SB_H_WARP,      // tile hash, warp/row,    ub <= 384  (512 slots)
    SB_H_BLOCK_S,   // tile hash, block/row,   ub <= 3072 (4096 slots)
    SB_H_BLOCK_L,   // tile hash, block/row,   ub <= 12288 (16384 slots)
    SB_H_GLOBAL,    // tile hash in global memory
    SB_TINY,        // one thread per row, tile list in shared memory, tile-flop <= 24
    SB_TINY_S,      // the same kernel, rows with tile-flop <= 4 ...
    SB_TINY_M,      // ... and <= 12.  The three classes are adjacent in the bin list and run as ONE launch
                    // over rows sorted by class: the threads of a warp then carry rows of similar cost (on
                    // power-law inputs the rows of the old single bin differed 100x in cost and a warp ran
                    // 4 of its 32 lanes, profiles/r2l_hash_kernels_R.md)
    SB_H_G16,       // tile hash, 16 lanes/row, ub <= 96 (256 slots): two rows per warp share the per-row
                    // fixed cost (~600 warp instructions) that dominates rows this short
    SB_COUNT
};
#define SB_BM_G8_WORDS 64
#define SB_BM_WARP_WORDS 2048
#define SB_BM_BLOCK_WORDS 57344
#define SB_H_G8_SLOTS 32
#define SB_H_G8_MAX 24
#define SB_H_G16_SLOTS 256
#define SB_H_G16_MAX 96
#define SB_H_WARP_SLOTS 512
#define SB_H_WARP_MAX 384
#define SB_H_BLOCK_S_SLOTS 4096
#define SB_H_BLOCK_S_MAX 3072
#define SB_H_BLOCK_L_SLOTS 16384
#define SB_H_BLOCK_L_MAX 12288
#define SB_TINY_MAX 24
#define SB_TINY_S_MAX 4
#define SB_TINY_M_MAX 12
#define SB_BITMAP_WORK_FACTOR 8 // bitmap when Wt <= 8 * tile-flop (or Wt <= 64)

// ---- numeric bins (family 4). W = column span of the C row, n = nnz of the C row ----
enum MhbNumBin
{
    NB_EMPTY = 0,
    NB_WIN_G8,      // dense window, 8 lanes/row, W <= 256
    NB_WIN_WARP,    // dense window, warp/row,    W <= 1024
    NB_WIN_BLOCK_S, // dense window, block/row,   W <= 6144
    NB_WIN_BLOCK_L, // dense window, block/row,   W <= 27648 (216 KB of fp64 + flags)
    NB_H_G8,        // hash, 8 lanes/row, n <= 24  (32 slots)
    NB_H_WARP_S,    // hash, warp/row,    n <= 160 (256 slots; fill <= 5/8 leaves room for the sort scratch)
    NB_H_WARP_L,    // hash, warp/row,    n <= 640 (1024 slots)
    NB_H_BLOCK_S,   // hash, block/row,   n <= 2560 (4096 slots)
    NB_H_BLOCK_L,   // hash, block/row,   n <= 10240 (16384 slots)
    NB_H_GLOBAL,    // hash in global memory
    NB_TINY,        // one thread per row, n <= 24 and <= 128 products
    NB_TINY_S,      // NB_TINY's kernel, rows with <= 8 products ...
    NB_TINY_M,      // ... and <= 32 products: the three classes are adjacent in the bin list and run as ONE
                    // launch over rows sorted by class (see SB_TINY_S)
    NB_H_WARP_XS,   // hash, warp/row,    n <= 80  (128 slots)
    NB_H_WARP_M,    // hash, warp/row,    n <= 320 (512 slots)
    NB_WIN_COMPACT, // window rows with a stored symbolic bitmap, n <= 448: rank-mapped accumulators,
                    // up to three twin rows of A per warp
    NB_H_BLOCK_M,   // hash, block/row,   n <= 5120 (8192 slots, claim list)
    NB_H_BLOCK_XS,  // hash, block/row,   n <= 1280 (2048 slots, claim list, 128 threads)
    NB_COUNT
};
#define NB_WIN_G8_COLS 256
#define NB_WIN_WARP_COLS 1024
#define NB_WIN_BLOCK_S_COLS 6144
#define NB_WIN_BLOCK_L_COLS 27648
#define NB_H_G8_SLOTS 32
#define NB_H_G8_MAX 24
#define NB_H_WARP_XS_SLOTS 128
#define NB_H_WARP_XS_MAX 80
#define NB_H_WARP_M_SLOTS 512
#define NB_H_WARP_M_MAX 320
#define NB_H_WARP_S_SLOTS 256
#define NB_H_WARP_S_MAX 160
#define NB_H_WARP_L_SLOTS 1024
#define NB_H_WARP_L_MAX 640
#define NB_H_BLOCK_XS_SLOTS 2048
#define NB_H_BLOCK_XS_MAX 1280
#define NB_H_BLOCK_S_SLOTS 4096
#define NB_H_BLOCK_S_MAX 2560
#define NB_H_BLOCK_M_SLOTS 8192
#define NB_H_BLOCK_M_MAX 5120
#define NB_H_BLOCK_L_SLOTS 16384
#define NB_H_BLOCK_L_MAX 10240
#define NB_WIN_COMPACT_MAXN 448
#define SB_BM_STORE_WORDS 64 // stride of a stored symbolic bitmap (the SB_BM_G8 bin)
#define NB_TINY_MAX 24
#define NB_TINY_PRODUCTS 128
#define NB_TINY_S_PRODUCTS 8
#define NB_TINY_M_PRODUCTS 32
#define NB_WINDOW_WORK_FACTOR 32 // window when W <= 32 * n (or W <= 64)

// path forcing (mhb_set_option "force_sym_path"/"force_num_path")
#define MHB_PATH_AUTO 0
#define MHB_PATH_DENSE 1 // bitmap / window wherever it fits
#define MHB_PATH_HASH 2  // hash only

#if defined(__CUDACC__)
#define MHB_HD __host__ __device__ __forceinline__
#else
#define MHB_HD inline
#endif

// Row metrics -> symbolic bin.  ip = intermediate products (saturating), tf = tile-flop.
MHB_HD int mhb_classify_sym(int ip, int tf, int cmin, int cmax, int force)
{
    if (ip <= 0)
        return SB_EMPTY;
    if (force == MHB_PATH_AUTO && tf <= SB_TINY_MAX)
        return tf <= SB_TINY_S_MAX ? SB_TINY_S : (tf <= SB_TINY_M_MAX ? SB_TINY_M : SB_TINY);
    long long wt = (long long)(cmax >> MHB_TILE_SHIFT) - (cmin >> MHB_TILE_SHIFT) + 1;
    bool fits = wt <= SB_BM_BLOCK_WORDS;
    bool dense = fits && (wt <= SB_BM_G8_WORDS || wt <= (long long)SB_BITMAP_WORK_FACTOR * tf);
    if (force == MHB_PATH_DENSE)
        dense = fits;
    if (force == MHB_PATH_HASH)
        dense = false;
    if (dense)
        return wt <= SB_BM_G8_WORDS ? SB_BM_G8 : (wt <= SB_BM_WARP_WORDS ? SB_BM_WARP : SB_BM_BLOCK);
    long long ub = wt < tf ? wt : tf;
    if (ub <= SB_H_G8_MAX)
        return SB_H_G8;
    if (ub <= SB_H_G16_MAX)
        return SB_H_G16;
    if (ub <= SB_H_WARP_MAX)
        return SB_H_WARP;
    if (ub <= SB_H_BLOCK_S_MAX)
        return SB_H_BLOCK_S;
    if (ub <= SB_H_BLOCK_L_MAX)
        return SB_H_BLOCK_L;
    return SB_H_GLOBAL;
}

// Row metrics -> numeric bin.  n = nnz of the C row.
// avg_b = intermediate products / nonzeros of the A row = mean length of the B rows it selects
MHB_HD int mhb_classify_num(int n, int ip, int cmin, int cmax, int force, int tf = 0, int force_sym = 0,
                            int compact_ok = 0, int avg_b = 0)
{
    if (n <= 0)
        return NB_EMPTY;
    if (force == MHB_PATH_AUTO && n <= NB_TINY_MAX && ip <= NB_TINY_PRODUCTS)
        return ip <= NB_TINY_S_PRODUCTS ? NB_TINY_S : (ip <= NB_TINY_M_PRODUCTS ? NB_TINY_M : NB_TINY);
    long long w = (long long)cmax - cmin + 1;
    bool fits = w <= NB_WIN_BLOCK_L_COLS;
    bool dense = fits && (w <= 64 || w <= (long long)NB_WINDOW_WORK_FACTOR * n);
    if (force == MHB_PATH_DENSE)
        dense = fits;
    if (force == MHB_PATH_HASH)
        dense = false;
    if (dense)
    {
        if (w <= NB_WIN_G8_COLS && avg_b <= 16) // short B rows: 8 lanes per row are enough
            return NB_WIN_G8;
        if (w <= NB_WIN_WARP_COLS)
            return (compact_ok && n <= NB_WIN_COMPACT_MAXN &&
                    mhb_classify_sym(ip, tf, cmin, cmax, force_sym) == SB_BM_G8)
                       ? NB_WIN_COMPACT
                       : NB_WIN_WARP;
        return w <= NB_WIN_BLOCK_S_COLS ? NB_WIN_BLOCK_S : NB_WIN_BLOCK_L;
    }
    if (n <= NB_H_G8_MAX)
        return NB_H_G8;
    if (n <= NB_H_WARP_XS_MAX)
        return NB_H_WARP_XS;
    if (n <= NB_H_WARP_S_MAX)
        return NB_H_WARP_S;
    if (n <= NB_H_WARP_M_MAX)
        return NB_H_WARP_M;
    if (n <= NB_H_WARP_L_MAX)
        return NB_H_WARP_L;
    if (n <= NB_H_BLOCK_XS_MAX)
        return NB_H_BLOCK_XS;
    if (n <= NB_H_BLOCK_S_MAX)
        return NB_H_BLOCK_S;
    if (n <= NB_H_BLOCK_M_MAX)
        return NB_H_BLOCK_M;
    if (n <= NB_H_BLOCK_L_MAX)
        return NB_H_BLOCK_L;
    return NB_H_GLOBAL;
}
