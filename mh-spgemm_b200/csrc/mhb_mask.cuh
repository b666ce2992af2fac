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

// ---------------------------------------------------------------------------------------
// One-pass builder (round 2): flags + scan + fill in ONE kernel that reads B.col ONCE.
//
//   k_mask_rowstarts   row-start bits into a zeroed flag array (shares one memset with the scalars)
//   k_mask_build       per 2 048-nonzero chunk: tile-start flags, chunk tile count, decoupled
//                      look-back for the tile index in front of the chunk (single-pass chained scan,
//                      chunks taken in ticket order so a chunk only ever waits for chunks that
//                      are already running), then tile columns and masks straight from the columns
//                      still held in registers
//   k_mask_rows        tile offsets + row descriptors + twin flags, one thread per row
//
// A tile run holds at most 32 nonzeros (distinct columns of one 32-column tile), so it spans
// at most two 32-nonzero words.  The warp that owns the word where a run STARTS also reads the
// leading lanes of the next word and writes the whole mask with one plain store: no atomicOr,
// and tilemask needs no zero fill (round 1 cleared nnz(B) words per call).
// ---------------------------------------------------------------------------------------
constexpr int kMaskThreads = 256;
constexpr int kMaskWordsPerWarp = 8;
constexpr int kMaskChunkWords = (kMaskThreads / 32) * kMaskWordsPerWarp; // 64 words = 2 048 nonzeros

// chunk status of the chained scan: (state << 62) | value
constexpr unsigned long long kScanAggregate = 1ull << 62; // value = tiles of this chunk only
constexpr unsigned long long kScanPrefix = 2ull << 62;    // value = tiles of all chunks up to and including this one
constexpr unsigned long long kScanValueMask = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Exclusive prefix of chunk `chunk` (sum of the aggregates of all chunks in front of it), by
// the WHOLE block: 256 predecessors per round, newest first (thread t looks at chunk - 1 - t);
// stops at the nearest one that already knows its inclusive prefix.  Publishes this chunk's
// aggregate first and its inclusive prefix last.  All chunks of a wave publish their aggregates
// at about the same time, so the depth of the walk is what costs: a one-warp window needed
// chunk/32 rounds of L2 latency (44 us for 2 070 chunks, r2f launch list); 256-wide it is 8x
// shallower.  Must be called by all threads of the block; returns the exclusive prefix.
// sh: NW + 2 long long of shared memory.
template <int NW>
__device__ __forceinline__ long long chained_scan_lookback(unsigned long long *status, int chunk, long long aggregate,
                                                           long long *sh)
{
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    if (threadIdx.x == 0)
        st_status(status + chunk, (chunk == 0 ? kScanPrefix : kScanAggregate) | (unsigned long long)aggregate);
    long long excl = 0;
    int idx = chunk - 1;
    while (idx >= 0)
    {
        const int j = idx - (int)threadIdx.x;
        unsigned long long sv = kScanPrefix; // in front of chunk 0: a prefix of 0
        if (j >= 0)
        {
            sv = ld_status(status + j);
            while ((sv >> 62) == 0)
            {
                __nanosleep(20);
                sv = ld_status(status + j);
            }
        }
        const unsigned has_prefix = __ballot_sync(kFull, (sv >> 62) == 2);
        const int stop = has_prefix ? __ffs(has_prefix) - 1 : 31; // nearest predecessor with a prefix in this warp's window
        long long v = (lane <= stop) ? (long long)(sv & kScanValueMask) : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            v += __shfl_xor_sync(kFull, v, o);
        if (lane == 0)
            sh[warp] = has_prefix ? -v - 1 : v; // negative: this window ends the walk
        __syncthreads();
        bool done = false;
#pragma unroll
        for (int w = 0; w < NW; ++w)
        {
            const long long x = sh[w];
            if (!done)
            {
                excl += (x < 0) ? -(x + 1) : x;
                done = x < 0;
            }
        }
        __syncthreads();
        if (done)
            break;
        idx -= NW * 32;
    }
    if (threadIdx.x == 0 && chunk != 0)
    {
        __threadfence();
        st_status(status + chunk, kScanPrefix | (unsigned long long)(excl + aggregate));
    }
    return excl;
}

// ctrl[0] = chunk ticket counter (zeroed with the scalars); status[nchunks] zeroed likewise;
// flags[] holds the row-start bits on entry and the complete tile-start flags on exit.
// MODE 0: the one-pass builder (chunks by ticket, chained scan).  MODE 1 / MODE 2: the same work as two
// passes around an ordinary scan of the per-chunk tile counts -- 1 forms the flags and counts the
// tiles of chunk blockIdx.x (chunk_tiles), 2 takes the chunk's offset from chunk_prefix and does the
// rest.  B.col is read twice (the second time from L2), but no block waits for another: the look-back
// chain is what bounds MODE 0 (issue 30 %, two thirds of the stall samples at its barriers).
#pragma nv_diag_suppress 128 // MODE 1 returns before the second half of the kernel
template <int MODE>
__global__ void __launch_bounds__(kMaskThreads, 4)
    k_mask_build(const int *__restrict__ Bc, long long nnz, long long nwords, unsigned *flags,
                 int *__restrict__ wordprefix, int *__restrict__ tilecol, unsigned *__restrict__ tilemask,
                 unsigned *__restrict__ ctrl, unsigned long long *__restrict__ status, int nchunks,
                 long long *__restrict__ total64, int *__restrict__ chunk_tiles, const int *__restrict__ chunk_prefix)
{
    pdl_prologue();
    constexpr int WPW = kMaskWordsPerWarp, NW = kMaskThreads / 32;
    __shared__ int sh_chunk;
    __shared__ int sh_warp[NW];
    __shared__ long long sh_look[NW + 2];
    if (MODE == 0)
    {
        if (threadIdx.x == 0)
            sh_chunk = (int)atomicAdd(ctrl, 1u);
        __syncthreads();
    }
    const int chunk = MODE == 0 ? sh_chunk : (int)blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = lane_id();
    // positions are 32-bit: nnz(B) < 2^31 by the int32 CSR contract, and a chunk overruns it by < 2^12
    const unsigned w0 = (unsigned)chunk * kMaskChunkWords + (unsigned)warp * WPW; // first word of this warp
    const unsigned n32 = (unsigned)nnz;
    // columns of the warp's WPW words plus one look-ahead word, all loads issued up front
    int c[WPW + 1];
#pragma unroll
    for (int i = 0; i <= WPW; ++i)
    {
        const unsigned j = (w0 + i) * 32 + lane;
        c[i] = (j < n32) ? __ldg(&Bc[j]) : -1;
    }
    int prev_last = -1; // column in front of the warp's first nonzero
    if (w0 > 0 && w0 * 32 - 1 < n32)
        prev_last = __ldg(&Bc[w0 * 32 - 1]);
    unsigned rowbits[WPW + 1];
#pragma unroll
    for (int i = 0; i <= WPW; ++i)
        rowbits[i] = (w0 + i < (unsigned)nwords) ? flags[w0 + i] : 0u; // row starts (or already-complete flags: a superset)
    // tile-start flags: the tile column changes, or a row starts
    unsigned f[WPW + 1];
    int tiles = 0;
#pragma unroll
    for (int i = 0; i <= WPW; ++i)
    {
        int p = __shfl_up_sync(kFull, c[i], 1);
        const int carry = (i == 0) ? prev_last : __shfl_sync(kFull, c[i - 1], 31);
        if (lane == 0)
            p = carry;
        const bool valid = c[i] >= 0;
        const bool start = valid && (p < 0 || (c[i] >> MHB_TILE_SHIFT) != (p >> MHB_TILE_SHIFT));
        f[i] = __ballot_sync(kFull, start) | rowbits[i];
        if (i < WPW)
            tiles += __popc(f[i]);
    }
    if (MODE == 1)
    {
#pragma unroll
        for (int i = 0; i < WPW; ++i)
            if (lane == 0 && w0 + i < (unsigned)nwords)
                flags[w0 + i] = f[i];
        if (lane == 0)
            sh_warp[warp] = tiles;
        __syncthreads();
        if (threadIdx.x == 0)
        {
            int agg = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w)
                agg += sh_warp[w];
            chunk_tiles[chunk] = agg;
        }
        return;
    }
    // The masks are formed BEFORE the look-back (they need no global position), so that the shuffle
    // work of this chunk overlaps the latency of its predecessors' status words; behind the
    // look-back only the stores are left.
    // Mask of every run that STARTS in word i: segmented suffix-OR inside the word.  (A match_any +
    // redux.or formulation was measured and lost -- 44 -> 54 us on the cant-like input, 36 -> 121 us
    // on the R-MAT where every lane is its own run: MATCH.ANY takes one round per distinct value.)
    unsigned m[WPW];
#pragma unroll
    for (int i = 0; i < WPW; ++i)
    {
        const unsigned fi = f[i];
        unsigned bits = (c[i] >= 0) ? (1u << (c[i] & 31)) : 0u;
        const unsigned above = (lane == 31) ? 0u : (fi >> (lane + 1));
        const int seg_end = above ? lane + __ffs(above) - 1 : 31; // last lane of this lane's run inside the word
        // step d is needed only while some run of the word is longer than d (warp-uniform test on
        // the flag word: z has bit l set when lanes l .. l+d all belong to one run)
        unsigned z = ~fi >> 1;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            if (z == 0)
                break;
            const unsigned v = __shfl_down_sync(kFull, bits, d);
            if (lane + d <= seg_end)
                bits |= v;
            z &= z >> d;
        }
        // ... plus the run's continuation in the leading lanes of the next word
        const unsigned fn = f[i + 1];
        const int lead = fn ? __ffs(fn) - 1 : 32;
        const unsigned ext = (lead == 0) ? 0u : __reduce_or_sync(kFull, (lane < lead && c[i + 1] >= 0) ? (1u << (c[i + 1] & 31)) : 0u);
        m[i] = (seg_end == 31) ? (bits | ext) : bits;
    }
    // chunk aggregate -> look-back -> tile index in front of every word
    if (lane == 0)
        sh_warp[warp] = tiles;
    __syncthreads();
    int agg = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w)
        agg += sh_warp[w];
    long long excl;
    if (MODE == 0)
    {
        excl = chained_scan_lookback<NW>(status, chunk, agg, sh_look);
        if (threadIdx.x == 0 && chunk == nchunks - 1)
            *total64 = excl + agg;
    }
    else
        excl = chunk_prefix[chunk]; // (the scan has written the total)
    long long run = excl;
#pragma unroll
    for (int w = 0; w < NW; ++w)
        if (w < warp)
            run += sh_warp[w];
#pragma unroll
    for (int i = 0; i < WPW; ++i)
    {
        const unsigned word = w0 + i;
        if (word >= (unsigned)nwords)
            break;
        const unsigned fi = f[i];
        if (lane == 0)
        {
            flags[word] = fi;
            wordprefix[word] = (int)run;
        }
        if (c[i] >= 0 && ((fi >> lane) & 1u))
        {
            const int t = (int)run + __popc(fi & lanemask_lt());
            tilecol[t] = c[i] >> MHB_TILE_SHIFT;
            tilemask[t] = m[i];
        }
        run += __popc(fi);
    }
}

#pragma nv_diag_default 128

// Tile offsets, the per-row descriptor {nnz, tiles, first col, last col} and the twin flag
// (same tile list as the previous row) of every row of B, one thread per row.
__global__ void __launch_bounds__(256) k_mask_rows(int K, long long nnz, const int *__restrict__ Bp,
                                                   const int *__restrict__ Bc, const unsigned *__restrict__ flags,
                                                   const int *__restrict__ wordprefix,
                                                   const long long *__restrict__ total64,
                                                   const int *__restrict__ tilecol,
                                                   const unsigned *__restrict__ tilemask,
                                                   int *__restrict__ tileptr, int4 *__restrict__ binfo,
                                                   unsigned char *__restrict__ same)
{
    pdl_prologue();
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > K)
        return;
    const int total = (int)*total64;
    if (k == K)
    {
        tileptr[K] = total;
        return;
    }
    const int s = __ldg(&Bp[k]), e = __ldg(&Bp[k + 1]);
    const int ts = tiles_before(s, nnz, flags, wordprefix, total);
    const int te = tiles_before(e, nnz, flags, wordprefix, total);
    tileptr[k] = ts;
    int first = INT_MAX, last = -1;
    if (e > s)
    {
        first = __ldg(&Bc[s]);
        last = __ldg(&Bc[e - 1]);
    }
    binfo[k] = make_int4(e - s, te - ts, first, last);
    unsigned char r = 0;
    if (k > 0 && e > s)
    {
        const int ps = __ldg(&Bp[k - 1]);
        if (s - ps == e - s && __ldg(&Bc[ps]) == first && __ldg(&Bc[s - 1]) == last)
        {
            const int pts = tiles_before(ps, nnz, flags, wordprefix, total);
            if (ts - pts == te - ts)
            {
                r = 1;
                for (int t = 0; t < te - ts; ++t)
                    if (tilecol[pts + t] != tilecol[ts + t] || tilemask[pts + t] != tilemask[ts + t])
                    {
                        r = 0;
                        break;
                    }
            }
        }
    }
    same[k] = r;
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
