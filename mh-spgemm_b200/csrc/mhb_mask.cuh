// mhb_mask.cuh -- kernel family 1: the B mask-matrix builder.
//
// Replaces Form_mask_matrix_B (inc/MH_spgemm.cuh:242-295) and its 15 kernels
// (inc/Form_mask_matrix_B.cuh): per row of B, one (tile column = col>>5, 32-bit occupancy
// mask, bit = col&31) pair per non-empty 32-column tile, plus exclusive tile offsets.
//
// The reference bins B's rows twice and hashes every column into a per-row shared-memory
// table (10 + 9 size-class kernels).  B's rows are sorted and duplicate-free (the CSR
// contract), so the tiles of a row are simply the runs of equal col>>5.  That makes the
// builder a flat, perfectly coalesced pass over B.col that is immune to row-length skew:
//   k_mask_flags      one bit per nonzero: "starts a new tile run" (ignoring row borders)
//   k_mask_rowstarts  OR in the row-start bits (a row start always starts a tile)
//   scan              popc prefix over the flag words  -> tile index of every nonzero
//   k_mask_tileptr    tile offsets + the per-row descriptor {nnz, tiles, first, last col}
//   k_mask_fill       segmented OR of the bit masks by warp shuffles, one write per run
// Tiles come out ascending inside a row (the reference's order is hash-slot order).
#pragma once
#include "mhb_common.cuh"

namespace mhb
{

__global__ void __launch_bounds__(256) k_mask_flags(const int *__restrict__ Bc, long long nnz, long long nwords,
                                                    unsigned *__restrict__ flags)
{
    pdl_prologue();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < nwords * 32; j += stride)
    {
        int c = (j < nnz) ? __ldg(&Bc[j]) : -1;
        int p = __shfl_up_sync(kFull, c, 1);
        if (lane_id() == 0)
            p = (j > 0 && j < nnz) ? __ldg(&Bc[j - 1]) : -1;
        bool f = (j < nnz) && (j == 0 || (c >> MHB_TILE_SHIFT) != (p >> MHB_TILE_SHIFT));
        unsigned w = __ballot_sync(kFull, f);
        if (lane_id() == 0)
            flags[j >> 5] = w;
    }
}

__global__ void __launch_bounds__(256) k_mask_rowstarts(int K, const int *__restrict__ Bp,
                                                        unsigned *__restrict__ flags)
{
    pdl_prologue();
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K)
        return;
    int s = __ldg(&Bp[k]), e = __ldg(&Bp[k + 1]);
    if (e > s)
    {
        unsigned bit = 1u << (s & 31);
        if (!(flags[s >> 5] & bit))
            atomicOr(&flags[s >> 5], bit);
    }
}

// tile index (exclusive count of tile starts) in front of nonzero position s
__device__ __forceinline__ int tiles_before(int s, long long nnz, const unsigned *__restrict__ flags,
                                            const int *__restrict__ wordprefix, int total)
{
    if (s >= nnz)
        return total;
    return wordprefix[s >> 5] + __popc(flags[s >> 5] & ((1u << (s & 31)) - 1u));
}

__global__ void __launch_bounds__(256) k_mask_tileptr(int K, long long nnz, const int *__restrict__ Bp,
                                                      const int *__restrict__ Bc,
                                                      const unsigned *__restrict__ flags,
                                                      const int *__restrict__ wordprefix,
                                                      const long long *__restrict__ total64,
                                                      int *__restrict__ tileptr, int4 *__restrict__ binfo)
{
    pdl_prologue();
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > K)
        return;
    int total = (int)*total64;
    if (k == K)
    {
        tileptr[K] = total;
        return;
    }
    int s = __ldg(&Bp[k]), e = __ldg(&Bp[k + 1]);
    int ts = tiles_before(s, nnz, flags, wordprefix, total);
    int te = tiles_before(e, nnz, flags, wordprefix, total);
    tileptr[k] = ts;
    int first = INT_MAX, last = -1;
    if (e > s)
    {
        first = __ldg(&Bc[s]);
        last = __ldg(&Bc[e - 1]);
    }
    binfo[k] = make_int4(e - s, te - ts, first, last);
}

// tilemask must be zero on entry (runs that cross a 32-nonzero word are merged with atomicOr).
__global__ void __launch_bounds__(256) k_mask_fill(const int *__restrict__ Bc, long long nnz, long long nwords,
                                                   const unsigned *__restrict__ flags,
                                                   const int *__restrict__ wordprefix,
                                                   int *__restrict__ tilecol, unsigned *__restrict__ tilemask)
{
    pdl_prologue();
    const long long stride = (long long)gridDim.x * blockDim.x;
    const int lane = lane_id();
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < nwords * 32; j += stride)
    {
        const bool valid = j < nnz;
        const unsigned f = flags[j >> 5];
        const int c = valid ? __ldg(&Bc[j]) : 0;
        // run this lane belongs to, and the lane where that run starts inside the word
        const unsigned upto = f & lanemask_le();
        const int t = wordprefix[j >> 5] + __popc(upto) - 1;
        const int head = upto ? (31 - __clz(upto)) : 0;
        unsigned bits = valid ? (1u << (c & 31)) : 0u;
        // segmented inclusive OR-scan over the lanes of the run
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            unsigned v = __shfl_up_sync(kFull, bits, d);
            if (lane - d >= head)
                bits |= v;
        }
        const bool next_starts = (lane == 31) || ((f >> (lane + 1)) & 1u);
        const bool next_valid = (lane < 31) && (j + 1 < nnz);
        const bool is_tail = valid && (next_starts || !next_valid);
        const bool starts_here = (f >> head) & 1u;                  // run began in this word
        const bool ends_here = (lane < 31) && (next_starts || !next_valid); // and ends in it
        if (valid && ((f >> lane) & 1u))
            tilecol[t] = c >> MHB_TILE_SHIFT;
        if (is_tail)
        {
            if (starts_here && ends_here)
                tilemask[t] = bits;
            else
                atomicOr(&tilemask[t], bits);
        }
    }
}

// same[k] = 1 when row k of B has exactly the column pattern of row k-1 (equal tile lists).
// Typical of multi-dof FEM matrices, where the rows of one node share their pattern.  The
// symbolic pass skips such a row when it directly follows its twin in a row of A (OR is
// idempotent); the numeric pass can fold the twins' products into one accumulator update.
// One thread per row; the tile comparison stops at the first difference, so rows that are
// not twins cost O(1).
__global__ void __launch_bounds__(256) k_mask_same(int K, const int4 *__restrict__ binfo,
                                                   const int *__restrict__ tileptr,
                                                   const int *__restrict__ tilecol,
                                                   const unsigned *__restrict__ tilemask,
                                                   unsigned char *__restrict__ same)
{
    pdl_prologue();
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K)
        return;
    unsigned char r = 0;
    if (k > 0)
    {
        const int4 a = binfo[k - 1], b = binfo[k];
        if (b.x > 0 && a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w)
        {
            const int pa = tileptr[k - 1], pb = tileptr[k];
            r = 1;
            for (int t = 0; t < b.y; ++t)
                if (tilecol[pa + t] != tilecol[pb + t] || tilemask[pa + t] != tilemask[pb + t])
                {
                    r = 0;
                    break;
                }
        }
    }
    same[k] = r;
}

// same[r] = 1 when row r of a CSR matrix has exactly the column list of row r-1 (used for the
// rows of A when A is not the same array as B; B's flags come from its tile lists above).
// 8 lanes per row compare the two column lists side by side (coalesced), group vote.
__global__ void __launch_bounds__(256) k_rows_same_cols(int M, const int *__restrict__ ptr,
                                                        const int *__restrict__ col,
                                                        unsigned char *__restrict__ same)
{
    pdl_prologue();
    constexpr int G = 8;
    const int l = threadIdx.x % G;
    const unsigned gm = group_mask<G>();
    const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const bool valid = gid < M && gid > 0;
    const int r = valid ? (int)gid : 1;
    bool eq = false;
    if (valid)
    {
        const int a0 = ptr[r - 1], b0 = ptr[r], b1 = ptr[r + 1];
        const int len = b1 - b0;
        eq = len > 0 && len == b0 - a0;
        if (eq)
            for (int t = l; t < len; t += G)
                if (col[a0 + t] != col[b0 + t])
                {
                    eq = false;
                    break;
                }
    }
    const bool all_eq = __all_sync(gm, eq);
    if (l == 0 && gid < M)
        same[gid] = (valid && all_eq) ? 1 : 0;
}

} // namespace mhb
