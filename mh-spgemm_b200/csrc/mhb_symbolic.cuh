// mhb_symbolic.cuh -- kernel family 3: the symbolic nnz(C) pass.
//
// Replaces Calculate_C_nnz (inc/MH_spgemm.cuh:297-362) and its 10 kernels
// (inc/Calculate_C_nnz.cuh): nnz of every row of C = A*B from B's mask matrix.
//
// The reference hashes B's tiles twice per C row (count distinct C tiles, re-bin, then OR
// the masks and popc).  Here one traversal suffices, and two accumulators are used:
//
//  * BITMAP (sm_100a: 227 KB of shared memory holds a 1.8 M-column bitmap): when the C
//    row spans few 32-column words relative to its work, the row's occupancy bitmap lives
//    in shared memory, word index = tile column - first tile column, and every B tile is
//    one OR.  nnz = popcount of the bitmap.  No keys, no probing, no second pass.
//  * TILE HASH (key = tile column, value = OR of masks) in shared memory sized per bin,
//    with a global-memory table for rows whose tile count exceeds 12 K.
//
// Group kernels (8 lanes or a warp per row) walk A's row one B row at a time; the tiles of
// one B row are distinct, so lanes never touch the same word/slot in the same step and the
// OR is a plain read-modify-write ordered by __syncwarp -- no shared-memory atomics.
// Block kernels (one block per row, warps on different B rows) use atomicOr.
#pragma once
#include "mhb_common.cuh"
#include "mhb_stream.cuh"

namespace mhb
{

constexpr int kSymThreads = 256;

// ---- bitmap, G lanes per row -----------------------------------------------------------
// Walks A's row G nonzeros at a time; the first tile chunk of B row i+1 is loaded before the
// tiles of B row i are OR-ed in (L2 latency overlaps the shared-memory updates).  A B row
// whose pattern equals the previous B row's (same[k]) and that directly follows it in A's
// row contributes nothing new and is skipped.
// Twin rows of A (asame[i]: row i has the column list of row i-1, multi-dof FEM) have the
// same C pattern: only the first row of a run is computed, its count (and stored bitmap slot)
// is copied to the rows that follow it in the bin list.  Work is handed out per warp in
// windows of 3 * (32/G) list entries; the run leaders of a window are dealt to the groups by
// ballot + find-nth-set-bit, so that the groups of a warp stay busy although two of every
// three rows need no work.
template <int G>
__global__ void __launch_bounds__(kSymThreads)
    k_sym_bitmap_group(RowList list, const int *__restrict__ Ap,
                       const int *__restrict__ Ac, const int *__restrict__ tileptr,
                       const int *__restrict__ tilecol, const unsigned *__restrict__ tilemask,
                       const int4 *__restrict__ arow, int *__restrict__ counts, int wcap,
                       const unsigned char *__restrict__ same, unsigned *__restrict__ bm_store,
                       int *__restrict__ bm_slot, const unsigned char *__restrict__ asame, int bm_cap,
                       int *__restrict__ scal)
{
    extern __shared__ unsigned sm_u[];
    const int *__restrict__ rows = list.begin();
    const int nrows = list.size();
    if (bm_store && nrows > bm_cap) // speculative launch sized the bitmap store for fewer rows: redo
    {
        if (blockIdx.x == 0 && threadIdx.x == 0)
            atomicMax(scal + SC_SPEC_MISS, 1);
        bm_store = nullptr;
        bm_slot = nullptr;
    }
    constexpr int NG = 32 / G, WIN = 3 * NG, WPB = kSymThreads / 32;
    const int g = threadIdx.x / G, l = threadIdx.x % G;
    const unsigned gm = group_mask<G>();
    unsigned *bm = sm_u + (size_t)g * wcap;
    const int gbase = lane_id() & ~(G - 1);
    const int gw = (lane_id() / G); // group inside the warp
    for (long long r0 = (long long)(blockIdx.x * WPB + (threadIdx.x >> 5)) * WIN; r0 < nrows;
         r0 += (long long)gridDim.x * WPB * WIN)
    {
      // run leaders of this window (a row that repeats its predecessor in the list is skipped)
      bool lead = false;
      {
          const long long ri = r0 + lane_id();
          if (lane_id() < WIN && ri < nrows)
          {
              const int rw = __ldg(&rows[ri]);
              lead = !(asame && ri > 0 && __ldg(&rows[ri - 1]) == rw - 1 && __ldg(&asame[rw]));
          }
      }
      const unsigned leaders = __ballot_sync(kFull, lead);
      const int nlead = __popc(leaders);
      for (int li = gw; li < nlead; li += NG)
      {
        const int r = (int)r0 + (int)__fns(leaders, 0, li + 1);
        const int row = rows[r];
        const int4 info = arow[row];
        const int tbase = info.z >> MHB_TILE_SHIFT;
        const int wt = (info.w >> MHB_TILE_SHIFT) - tbase + 1;
        for (int w = l; w < wt; w += G)
            bm[w] = 0u;
        __syncwarp(gm);
        const int s = __ldg(&Ap[row]), e = __ldg(&Ap[row + 1]);
        int ts = 0, te = 0, kk = -2, nts = 0, nte = 0, nkk = -2;
        auto load_meta = [&](int j, int &ms, int &me, int &mk) {
            ms = 0, me = 0, mk = -2;
            if (j < e)
            {
                mk = __ldg(&Ac[j]);
                ms = __ldg(&tileptr[mk]);
                me = __ldg(&tileptr[mk + 1]);
                if (__ldg(&same[mk]))
                    mk |= kTwinTag; // twin of row mk-1
            }
        };
        load_meta(s + l, ts, te, kk);
        int prev_k = -2; // column of the last nonzero of the previous chunk
        for (int j0 = s; j0 < e; j0 += G)
        {
            load_meta(j0 + G + l, nts, nte, nkk);
            // drop rows that repeat the pattern of the nonzero just before them
            int pk = __shfl_up_sync(gm, kk & kTwinMask, 1, G);
            if (l == 0)
                pk = prev_k;
            if ((kk & kTwinTag) && (kk & kTwinMask) == pk + 1)
                te = ts;
            prev_k = __shfl_sync(gm, kk & kTwinMask, G - 1, G);
            // visit only the nonzeros that still have tiles to contribute (twins were emptied)
            unsigned live = (__ballot_sync(gm, te > ts) >> gbase) & (G == 32 ? 0xffffffffu : ((1u << G) - 1u));
            int pc = -1, nq = 0, nqe = 0;
            unsigned pm = 0u;
            auto issue = [&](int i) {
                nq = __shfl_sync(gm, ts, i, G);
                nqe = __shfl_sync(gm, te, i, G);
                pc = -1;
                if (nq + l < nqe)
                {
                    pc = __ldg(&tilecol[nq + l]);
                    pm = __ldg(&tilemask[nq + l]);
                }
            };
            if (live)
                issue(__ffs(live) - 1);
            while (live)
            {
                live &= live - 1;
                const int c = pc, q = nq, qe = nqe;
                const unsigned m = pm;
                if (live)
                    issue(__ffs(live) - 1);
                if (c >= 0)
                    bm[c - tbase] |= m; // tiles of one B row are distinct: no atomic
                for (int p = q + G + l; p < qe; p += G)
                    bm[__ldg(&tilecol[p]) - tbase] |= __ldg(&tilemask[p]);
                __syncwarp(gm);
            }
            ts = nts, te = nte, kk = nkk;
        }
        int c = 0;
        for (int w = l; w < wt; w += G)
        {
            const unsigned m = bm[w];
            c += __popc(m);
            if (bm_store) // keep the row's pattern for the numeric pass (NB_WIN_COMPACT)
                bm_store[(size_t)r * SB_BM_STORE_WORDS + w] = m;
        }
        c = group_sum<G>(c, gm);
        if (l == 0)
        {
            counts[row] = c;
            if (bm_store)
                bm_slot[row] = r;
            if (asame) // the twins that follow in the list share the result
                for (int f = 1; r + f < nrows; ++f)
                {
                    const int rf = __ldg(&rows[r + f]);
                    if (rf != row + f || !__ldg(&asame[rf]))
                        break;
                    counts[rf] = c;
                    if (bm_store)
                        bm_slot[rf] = r;
                }
        }
        __syncwarp(gm);
      }
    }
}

__device__ __forceinline__ int block_sum_int(int v, int *sh /*32*/)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(kFull, v, o);
    __syncthreads(); // sh may still be read from a previous call
    if (lane_id() == 0)
        sh[threadIdx.x >> 5] = v;
    __syncthreads();
    int t = 0;
    const int nw = blockDim.x >> 5;
    for (int w = 0; w < nw; ++w)
        t += sh[w];
    return t;
}

// ---- bitmap, one block per row (Wt up to 57 344 words) ---------------------------------
__global__ void __launch_bounds__(kSymThreads)
    k_sym_bitmap_block(RowList list, const int *__restrict__ Ap,
                       const int *__restrict__ Ac, const int *__restrict__ tileptr,
                       const int *__restrict__ tilecol, const unsigned *__restrict__ tilemask,
                       const int4 *__restrict__ arow, int *__restrict__ counts)
{
    extern __shared__ unsigned sm_u[];
    __shared__ int red[32];
    const int *__restrict__ rows = list.begin();
    const int nrows = list.size();
    unsigned *bm = sm_u;
    const int warp = threadIdx.x >> 5, lane = lane_id(), nwarp = blockDim.x >> 5;
    for (int r = blockIdx.x; r < nrows; r += gridDim.x)
    {
        const int row = rows[r];
        const int4 info = arow[row];
        const int tbase = info.z >> MHB_TILE_SHIFT;
        const int wt = (info.w >> MHB_TILE_SHIFT) - tbase + 1;
        for (int w = threadIdx.x; w < wt; w += blockDim.x)
            bm[w] = 0u;
        __syncthreads();
        const int s = __ldg(&Ap[row]), e = __ldg(&Ap[row + 1]);
        // every warp takes 32 nonzeros of A at a time and expands their tile lists flat
        walk_flat<32, NoVal, unsigned>(kFull, lane, s, e, warp, nwarp, Ac, (const NoVal *)nullptr,
                                       tileptr, tilecol, tilemask, [&](int tc, unsigned m, NoVal) {
                                           unsigned *w = &bm[tc - tbase];
                                           if ((*w & m) != m)
                                               atomicOr(w, m);
                                       });
        __syncthreads();
        int c = 0;
        for (int w = threadIdx.x; w < wt; w += blockDim.x)
            c += __popc(bm[w]);
        c = block_sum_int(c, red);
        if (threadIdx.x == 0)
            counts[row] = c;
        __syncthreads();
    }
}

// ---- tile hash: insert (tilecol -> OR mask) with linear probing -------------------------
template <bool ATOMIC_OR>
__device__ __forceinline__ void tile_insert(int *keys, unsigned *masks, int logS, int tc, unsigned m, int *scal)
{
    const unsigned S1 = (1u << logS) - 1u;
    unsigned h = hash_slot((unsigned)tc, logS);
    for (unsigned it = 0; it <= S1; ++it)
    {
        int old = keys[h];
        if (old != tc)
        {
            if (old != -1)
            {
                h = (h + 1) & S1;
                continue;
            }
            old = atomicCAS(&keys[h], -1, tc);
            if (old != -1 && old != tc)
            {
                h = (h + 1) & S1;
                continue;
            }
        }
        if (ATOMIC_OR)
            atomicOr(&masks[h], m);
        else
            masks[h] |= m;
        return;
    }
    atomicMax(scal + SC_ERROR, (int)DEVERR_TABLE_FULL);
}

// The same insertion with the lanes of a group in lockstep (tc < 0: nothing to insert): the
// probe loop leaves when every lane has its slot, so the mask update runs once per step
// instead of once per divergent exit path.  Must be called by ALL lanes of `gm`.
__device__ __forceinline__ int tile_insert_lockstep(unsigned gm, int *keys, unsigned *masks, int logS, int tc,
                                                    unsigned m, int *scal, int &np)
{
    const unsigned S1 = (1u << logS) - 1u;
    unsigned h = hash_slot((unsigned)tc, logS);
    bool more = tc >= 0;
    for (unsigned it = 0; it <= S1 && __any_sync(gm, more); ++it)
    {
        if (more)
        {
            const int old = atomicCAS(&keys[h], -1, tc);
            if (old == -1 || old == tc)
                more = false;
            else
            {
                ++np;
                h = (h + 1) & S1;
            }
        }
    }
    if (more)
        atomicMax(scal + SC_ERROR, (int)DEVERR_TABLE_FULL);
    else if (tc >= 0)
        atomicOr(&masks[h], m);
    return -1;
}

template <int G>
__global__ void __launch_bounds__(kSymThreads)
    k_sym_hash_group(RowList list, const int *__restrict__ Ap,
                     const int *__restrict__ Ac, const int *__restrict__ tileptr,
                     const int *__restrict__ tilecol, const unsigned *__restrict__ tilemask,
                     const int4 *__restrict__ arow, int *__restrict__ counts, int logS, int *__restrict__ scal,
                     unsigned long long *__restrict__ probes)
{
    extern __shared__ unsigned sm_u[];
    constexpr int GPB = kSymThreads / G;
    const int *__restrict__ rows = list.begin();
    const int nrows = list.size();
    int np = 0;
    const int g = threadIdx.x / G, l = threadIdx.x % G;
    const unsigned gm = group_mask<G>();
    const int Smax = 1 << logS;
    int *keys = (int *)sm_u + (size_t)g * 2 * Smax;
    unsigned *masks = (unsigned *)(keys + Smax);
    for (int r = blockIdx.x * GPB + g; r < nrows; r += gridDim.x * GPB)
    {
        const int row = rows[r];
        // table of THIS row: power of two >= 2x its tile upper bound min(tile-flop, spanned words),
        // at most the bin's; initialisation and the popcount sweep then scale with the row, not
        // the bin (a 4/3 bound was tried first: the lockstep probe loop runs as long as the
        // unluckiest lane, and at fill 3/4 that cost more than the sweeps saved -- r2g)
        const int4 info = __ldg(&arow[row]);
        const int wt = (info.w >> MHB_TILE_SHIFT) - (info.z >> MHB_TILE_SHIFT) + 1;
        const int ub = min(info.y, wt);
        int lS = 5;
        while (lS < logS && (1 << lS) < ub * 2)
            ++lS;
        const int S = 1 << lS;
        for (int w = l; w < S; w += G)
        {
            keys[w] = -1;
            masks[w] = 0u;
        }
        __syncwarp(gm);
        const int s = __ldg(&Ap[row]), e = __ldg(&Ap[row + 1]);
        walk_flat_post<G, NoVal, unsigned>(
            gm, l, s, e, 0, 1, Ac, (const NoVal *)nullptr, tileptr, tilecol, tilemask,
            [&](int tc, unsigned m, NoVal) { return tile_insert_lockstep(gm, keys, masks, lS, tc, m, scal, np); },
            [](int) {});
        __syncwarp(gm);
        int c = 0;
        for (int w = l; w < S; w += G)
            c += __popc(masks[w]);
        c = group_sum<G>(c, gm);
        if (l == 0)
            counts[row] = c;
        __syncwarp(gm);
    }
    flush_probes(probes, np);
}

// One block per row; table in shared memory (pool == nullptr) or in a per-block slice of
// the global pool (2 * pool_slots ints per block), sized per row from its tile upper bound.
__global__ void __launch_bounds__(kSymThreads)
    k_sym_hash_block(RowList list, const int *__restrict__ Ap,
                     const int *__restrict__ Ac, const int *__restrict__ tileptr,
                     const int *__restrict__ tilecol, const unsigned *__restrict__ tilemask,
                     const int4 *__restrict__ arow, int *__restrict__ counts, int logS_fixed,
                     int *__restrict__ pool, long long pool_slots, int *__restrict__ scal,
                     unsigned long long *__restrict__ probes)
{
    extern __shared__ unsigned sm_u[];
    __shared__ int red[32];
    const int *__restrict__ rows = list.begin();
    const int nrows = list.size();
    int np = 0;
    const int warp = threadIdx.x >> 5, lane = lane_id(), nwarp = blockDim.x >> 5;
    for (int r = blockIdx.x; r < nrows; r += gridDim.x)
    {
        const int row = rows[r];
        int logS = logS_fixed;
        int *keys;
        if (pool)
        {
            // per-row size: next power of two >= 1.5 * min(tile-flop, spanned words)
            const int4 info = arow[row];
            long long wt = (long long)(info.w >> MHB_TILE_SHIFT) - (info.z >> MHB_TILE_SHIFT) + 1;
            long long ub = wt < info.y ? wt : info.y;
            logS = 10;
            while ((1LL << logS) < ub + (ub >> 1) + 1)
                ++logS;
            if ((1LL << logS) > pool_slots) // the speculative launch sized the pool for smaller rows: redo
            {
                if (threadIdx.x == 0)
                    atomicMax(scal + SC_SPEC_MISS, 1);
                continue;
            }
            keys = pool + (size_t)blockIdx.x * 2 * pool_slots;
        }
        else
        {
            const int4 info = arow[row];
            const int wt = (info.w >> MHB_TILE_SHIFT) - (info.z >> MHB_TILE_SHIFT) + 1;
            const int ub = min(info.y, wt);
            logS = 8; // the row's own table: power of two >= 2 ub, at most the bin's
            while (logS < logS_fixed && (1 << logS) < ub * 2)
                ++logS;
            keys = (int *)sm_u;
        }
        const int S = 1 << logS;
        unsigned *masks = (unsigned *)(keys + S);
        for (int w = threadIdx.x; w < S; w += blockDim.x)
        {
            keys[w] = -1;
            masks[w] = 0u;
        }
        __syncthreads();
        const int s = __ldg(&Ap[row]), e = __ldg(&Ap[row + 1]);
        walk_flat_post<32, NoVal, unsigned>(
            kFull, lane, s, e, warp, nwarp, Ac, (const NoVal *)nullptr, tileptr, tilecol, tilemask,
            [&](int tc, unsigned m, NoVal) { return tile_insert_lockstep(kFull, keys, masks, logS, tc, m, scal, np); },
            [](int) {});
        __syncthreads();
        int c = 0;
        for (int w = threadIdx.x; w < S; w += blockDim.x)
            c += __popc(pool ? __ldcg(&masks[w]) : masks[w]); // pool: read at L2, where the atomics landed
        c = block_sum_int(c, red);
        if (threadIdx.x == 0)
            counts[row] = c;
        __syncthreads();
    }
    flush_probes(probes, np);
}

// ---- tiny rows: one thread per row ------------------------------------------------------
// Rows with at most SB_TINY_MAX tile visits (stencils, road networks, the bulk of power-law
// rows).  A group of lanes per row spends most of its instructions on shuffles, table
// initialisation and reductions for a dozen tiles; here a thread keeps the row's (tile,
// mask) list in its own shared-memory column (slot j of thread t at [j*blockDim + t]: bank
// = t, conflict-free whatever j each lane is at) and scans it linearly.
constexpr int kTinyThreads = 256;

__global__ void __launch_bounds__(kTinyThreads)
    k_sym_tiny(RowList list, const int *__restrict__ Ap, const int *__restrict__ Ac,
               const int *__restrict__ tileptr, const int *__restrict__ tilecol,
               const unsigned *__restrict__ tilemask, int *__restrict__ counts)
{
    const int *__restrict__ rows = list.begin();
    const int nrows = list.size();
    __shared__ int keys[SB_TINY_MAX * kTinyThreads];
    __shared__ unsigned masks[SB_TINY_MAX * kTinyThreads];
    const int t = threadIdx.x;
    for (int r = blockIdx.x * kTinyThreads + t; r < nrows; r += gridDim.x * kTinyThreads)
    {
        const int row = rows[r];
        int n = 0;
        const int s = __ldg(&Ap[row]), e = __ldg(&Ap[row + 1]);
        // one loop over the tile visits of the row (see k_num_tiny: the nested form reconverges the
        // warp at the end of every B row)
        int j = s, q = 0, qe = 0;
        while (true)
        {
            while (q == qe && j < e)
            {
                const int k = __ldg(&Ac[j]);
                q = __ldg(&tileptr[k]), qe = __ldg(&tileptr[k + 1]);
                ++j;
            }
            if (q == qe)
                break;
            const int tc = __ldg(&tilecol[q]);
            const unsigned m = __ldg(&tilemask[q]);
            ++q;
            int p = 0;
            while (p < n && keys[p * kTinyThreads + t] != tc)
                ++p;
            if (p < n)
                masks[p * kTinyThreads + t] |= m;
            else
            {
                keys[n * kTinyThreads + t] = tc;
                masks[n * kTinyThreads + t] = m;
                ++n;
            }
        }
        int c = 0;
        for (int p = 0; p < n; ++p)
            c += __popc(masks[p * kTinyThreads + t]);
        counts[row] = c;
    }
}

} // namespace mhb
