// mhb_binning.cuh -- kernel family 2: per-row intermediate-product counting, binning and
// the exclusive scans, built on warp-level prefix sums (ballot / popc / shuffle scans).
//
// Replaces: k_calculate_flop, k_calculate_flop_tmp (inc/Form_mask_matrix_B.cuh:14-95),
// k_binning1/k_binning2 + the host round trip of binning<TYPE> (inc/binning.cuh:67-155,
// inc/MH_spgemm.cuh:26-43) and the three cub::DeviceScan::ExclusiveSum call sites
// (inc/MH_spgemm.cuh:269,335; src/main.cu:55).
//
// Differences by design: bin offsets are computed on the device (no D2H/H2D per binning
// call), the scatter is a stable partition (row ids ascending inside a bin, so runs are
// reproducible -- the reference's atomic cursors are not), and one pass over A yields all
// four per-row metrics (products, tile-flop, min / max column of the C row).
#pragma once
#include "mhb_common.cuh"

namespace mhb
{

// Zero `n16` 16-byte words (the per-call scalar / ticket / status / flag block).  A kernel, not
// cudaMemsetAsync: the driver's memset of this 0.05-0.5 MB block measured 50 us on the stream
// (r2c: mem_alloc 0.005 -> 0.057 ms), a grid-stride store loop takes 2-3 us and chains with
// programmatic dependent launch like the rest of the pipeline.
__global__ void __launch_bounds__(256) k_zero16(uint4 *__restrict__ p, long long n16)
{
    pdl_prologue();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride)
        p[i] = make_uint4(0u, 0u, 0u, 0u);
}

// ---------------------------------------------------------------------------------------
// Exclusive scan of n ints (n up to 2^31) in two launches: block sums, then scan + apply.
// ---------------------------------------------------------------------------------------
constexpr int kScanThreads = 1024;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

struct LoadInt
{
    const int *p;
    __device__ __forceinline__ int operator()(long long i) const { return p[i]; }
};
struct LoadPopc
{
    const unsigned *p;
    __device__ __forceinline__ int operator()(long long i) const { return __popc(p[i]); }
};

__device__ __forceinline__ long long block_reduce_sum(long long v, long long *sh /*32*/)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(kFull, v, o);
    if (lane_id() == 0)
        sh[threadIdx.x >> 5] = v;
    __syncthreads();
    long long t = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0;
    if (threadIdx.x < 32)
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            t += __shfl_xor_sync(kFull, t, o);
        if (threadIdx.x == 0)
            sh[0] = t;
    }
    __syncthreads();
    t = sh[0];
    __syncthreads();
    return t;
}

template <class Load>
__global__ void __launch_bounds__(kScanThreads) k_scan_blocksums(Load in, long long n, long long *blocksums)
{
    pdl_prologue();
    __shared__ long long sh[32];
    long long base = (long long)blockIdx.x * kScanTile;
    long long s = 0;
#pragma unroll
    for (int it = 0; it < kScanItems; ++it)
    {
        long long i = base + it * kScanThreads + threadIdx.x;
        if (i < n)
            s += in(i);
    }
    s = block_reduce_sum(s, sh);
    if (threadIdx.x == 0)
        blocksums[blockIdx.x] = s;
}

// out[i] = sum_{j<i} in(j) for i in [0, n); out[n] = total when write_total; *total64 = total.
// In-place (out aliasing the input array) is safe: every item is read before any write of
// the same block and blocks touch disjoint ranges.
template <class Load>
__global__ void __launch_bounds__(kScanThreads) k_scan_apply(Load in, long long n, const long long *blocksums,
                                                             int nblocks, int *out, int write_total,
                                                             long long *total64)
{
    pdl_prologue();
    __shared__ long long sh[32];
    __shared__ int warp_tot[32];
    // offset of this block = sum of the preceding block sums (blocksums == nullptr: the scan is one
    // block, launched alone -- no block-sum pass in front of it; the total falls out of the scan below)
    long long pre = 0, all = 0;
    if (blocksums)
    {
        for (int b = threadIdx.x; b < nblocks; b += kScanThreads)
        {
            long long v = blocksums[b];
            all += v;
            if (b < (int)blockIdx.x)
                pre += v;
        }
        pre = block_reduce_sum(pre, sh);
        if (blockIdx.x == gridDim.x - 1)
        {
            all = block_reduce_sum(all, sh);
            if (threadIdx.x == 0)
            {
                if (total64)
                    *total64 = all;
                if (write_total)
                    out[n] = sat_i32(all);
            }
        }
    }
    // thread t owns kScanItems consecutive items
    long long base = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanItems;
    int v[kScanItems];
    int tsum = 0;
#pragma unroll
    for (int it = 0; it < kScanItems; ++it)
    {
        long long i = base + it;
        v[it] = (i < n) ? in(i) : 0;
        tsum += v[it];
    }
    // exclusive scan of tsum across the block: warp shuffle scan + warp totals
    int incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        int t = __shfl_up_sync(kFull, incl, o);
        if (lane_id() >= o)
            incl += t;
    }
    if (lane_id() == 31)
        warp_tot[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32)
    {
        int w = warp_tot[threadIdx.x];
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            int t = __shfl_up_sync(kFull, wi, o);
            if (lane_id() >= o)
                wi += t;
        }
        warp_tot[threadIdx.x] = wi - w;
    }
    __syncthreads();
    long long run = pre + warp_tot[threadIdx.x >> 5] + (incl - tsum);
    if (!blocksums && threadIdx.x == kScanThreads - 1)
    {
        const long long total = run + tsum;
        if (total64)
            *total64 = total;
        if (write_total)
            out[n] = sat_i32(total);
    }
#pragma unroll
    for (int it = 0; it < kScanItems; ++it)
    {
        long long i = base + it;
        if (i < n)
            out[i] = sat_i32(run);
        run += v[it];
    }
}

// ---------------------------------------------------------------------------------------
// Row metrics of C = A*B in one pass over A (G lanes per row, whole warp for long rows).
//   arow[i] = {intermediate products (saturating), tile-flop, min column, max column}
// also: symbolic bin id of the row, zero nnz for rows without products, global totals.
// binfo[k] = {nnz, tiles, first column, last column} of B row k (written by family 1).
// ---------------------------------------------------------------------------------------
struct RowAcc
{
    long long ip;
    long long tf;
    int cmin;
    int cmax;
};

__device__ __forceinline__ void row_acc_range(RowAcc &a, const int *__restrict__ Ac, const int4 *__restrict__ binfo,
                                              int s, int e, int lane, int stride)
{
    for (int j = s + lane; j < e; j += stride)
    {
        int4 b = __ldg(&binfo[__ldg(&Ac[j])]);
        a.ip += b.x;
        a.tf += b.y;
        if (b.x > 0)
        {
            a.cmin = min(a.cmin, b.z);
            a.cmax = max(a.cmax, b.w);
        }
    }
}

template <int G>
__global__ void __launch_bounds__(256) k_arow_metrics(int M, const int *__restrict__ Ap, const int *__restrict__ Ac,
                                                      const int4 *__restrict__ binfo, int4 *__restrict__ arow,
                                                      unsigned char *__restrict__ binid, int *__restrict__ counts,
                                                      int *__restrict__ scal, int force_path,
                                                      const unsigned char *__restrict__ twins)
{
    pdl_prologue();
    // twins (optional): twins[r] != 0 when row r has the column list of row r-1 (multi-dof FEM).  Such
    // a row has the metrics of its run's first row: only run leaders are computed, and the group that
    // computed a leader writes the followers' records too.
    constexpr int kLong = 32 * G; // rows longer than this are walked by the whole warp
    constexpr int GPW = 32 / G;   // rows per warp and step
    __shared__ long long sh_ip[8], sh_tf[8];
    __shared__ int sh_mx[8];
    const int l = threadIdx.x % G;
    const unsigned gm = group_mask<G>();
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    long long ip_tot = 0, tf_tot = 0;
    int tf_max = 0;
    // persistent grid: the trip count is warp-uniform so the ballots below are convergent
    for (long long base = warp0 * GPW; base < M; base += nwarps * GPW)
    {
        const long long gid = base + lane_id() / G;
        const bool in_range = gid < M;
        const int row = in_range ? (int)gid : 0;
        const bool valid = in_range && !(twins && row > 0 && __ldg(&twins[row])); // followers are written by their leader
        int s = 0, e = 0;
        if (valid)
        {
            s = __ldg(&Ap[row]);
            e = __ldg(&Ap[row + 1]);
        }
        RowAcc a{0, 0, INT_MAX, -1};
        const bool is_long = (G < 32) && (e - s) > kLong;
        if (!is_long)
            row_acc_range(a, Ac, binfo, s, e, l, G);
        a.ip = group_sum<G>(a.ip, gm);
        a.tf = group_sum<G>(a.tf, gm);
        a.cmin = group_min<G>(a.cmin, gm);
        a.cmax = group_max<G>(a.cmax, gm);
        if (G < 32)
        {
            // long rows: every lane of the warp helps, one row at a time
            unsigned todo = __ballot_sync(kFull, is_long && l == 0);
            while (todo)
            {
                const int src = __ffs(todo) - 1;
                todo &= todo - 1;
                const int rs = __shfl_sync(kFull, s, src), re = __shfl_sync(kFull, e, src);
                RowAcc w{0, 0, INT_MAX, -1};
                row_acc_range(w, Ac, binfo, rs, re, lane_id(), 32);
                w.ip = group_sum<32>(w.ip, kFull);
                w.tf = group_sum<32>(w.tf, kFull);
                w.cmin = group_min<32>(w.cmin, kFull);
                w.cmax = group_max<32>(w.cmax, kFull);
                if ((lane_id() & ~(G - 1)) == src)
                    a = w;
            }
        }
        if (valid && l == 0)
        {
            const int ip = sat_i32(a.ip), tf = sat_i32(a.tf);
            const int b = mhb_classify_sym(ip, tf, a.cmin, a.cmax, force_path);
            if (row == 0)
                counts[M] = 0;
            int r2 = row;
            do // the row and the twins that follow it
            {
                arow[r2] = make_int4(ip, tf, a.cmin, a.cmax);
                binid[r2] = (unsigned char)b;
                if (b == SB_EMPTY)
                    counts[r2] = 0;
                ip_tot += a.ip;
                tf_tot += a.tf;
                ++r2;
            } while (twins && r2 < M && __ldg(&twins[r2]));
            tf_max = max(tf_max, tf);
        }
    }
    // one set of atomics per block (per-warp atomics on three hot addresses serialise in L2)
    ip_tot = group_sum<32>(ip_tot, kFull);
    tf_tot = group_sum<32>(tf_tot, kFull);
    tf_max = group_max<32>(tf_max, kFull);
    if (lane_id() == 0)
    {
        sh_ip[threadIdx.x >> 5] = ip_tot;
        sh_tf[threadIdx.x >> 5] = tf_tot;
        sh_mx[threadIdx.x >> 5] = tf_max;
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
        {
            ip_tot += sh_ip[w];
            tf_tot += sh_tf[w];
            tf_max = max(tf_max, sh_mx[w]);
        }
        if (ip_tot | tf_tot)
        {
            atomicAdd((unsigned long long *)(scal + SC_INTPROD_LO), (unsigned long long)ip_tot);
            atomicAdd((unsigned long long *)(scal + SC_TILEFLOP_LO), (unsigned long long)tf_tot);
            atomicMax(scal + SC_MAX_TILEFLOP, tf_max);
        }
    }
}

// Numeric bin id from the exact nnz of every C row (counts, before the scan).
__global__ void __launch_bounds__(256) k_classify_num(int M, const int *__restrict__ Ap,
                                                      const int *__restrict__ counts,
                                                      const int4 *__restrict__ arow,
                                                      unsigned char *__restrict__ binid, int *__restrict__ scal,
                                                      int force_path, int force_sym, int compact_ok)
{
    pdl_prologue();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int n = 0;
    if (i < M)
    {
        n = counts[i];
        int4 info = arow[i];
        const int na = Ap[i + 1] - Ap[i];
        binid[i] = (unsigned char)mhb_classify_num(n, info.x, info.z, info.w, force_path, info.y, force_sym, compact_ok,
                                                   na > 0 ? info.x / na : 0);
    }
    __shared__ int sh_mx[8];
    n = group_max<32>(n, kFull);
    if (lane_id() == 0)
        sh_mx[threadIdx.x >> 5] = n;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
            n = max(n, sh_mx[w]);
        if (n > 0)
            atomicMax(scal + SC_MAX_ROWNNZ, n);
    }
}

// Fused call (mhb_spgemm_into_*): the numeric kernels are launched before the host has read
// anything of this call.  This one-thread kernel, between the row-offset scan and the numeric
// phase, verifies on the device what do_symbolic otherwise verifies on the host -- the
// speculative symbolic launch covered the input, every populated numeric bin has a kernel coming,
// the pool is large enough, C fits the caller's buffers -- and leaves the verdict in scal[SC_GATE].
__global__ void k_fused_gate(int *__restrict__ scal, unsigned sym_launched, int planned_tileflop,
                             unsigned num_launched, int planned_rownnz, long long capacity, int sb_global,
                             int nb_global, int sb_count, int nb_count)
{
    pdl_prologue();
    if (threadIdx.x != 0 || blockIdx.x != 0)
        return;
    int g = 0;
    if (scal[SC_SPEC_MISS])
        g |= GATE_SYM_MISS;
    for (int b = 1; b < sb_count; ++b) // bin 0 (rows without products) has no kernel
        if (scal[SC_SYM_SIZE + b] > 0 && !((sym_launched >> b) & 1u))
            g |= GATE_SYM_MISS;
    if (scal[SC_SYM_SIZE + sb_global] > 0 && scal[SC_MAX_TILEFLOP] > planned_tileflop)
        g |= GATE_SYM_MISS;
    for (int b = 1; b < nb_count; ++b)
        if (scal[SC_NUM_SIZE + b] > 0 && !((num_launched >> b) & 1u))
            g |= GATE_NUM_MISS;
    if (scal[SC_NUM_SIZE + nb_global] > 0 && scal[SC_MAX_ROWNNZ] > planned_rownnz)
        g |= GATE_NUM_MISS;
    const long long nnz = *reinterpret_cast<const long long *>(scal + SC_NNZC_LO);
    if (nnz > capacity)
        g |= GATE_CAPACITY;
    if (scal[SC_ERROR] != DEVERR_NONE)
        g |= GATE_ERROR;
    scal[SC_GATE] = g;
}

// Host-buffer path: the numeric phase is cut into row chunks so that the download of a finished
// chunk of C overlaps the computation of the next.  A bin's row list is ascending, so the rows of
// chunk c (rows [r[c], r[c+1]) of the matrix) are a contiguous piece of every bin's list: this
// kernel finds its ends by binary search.  out[b * (nc + 1) + c] = first position (in the
// concatenated bin array) of bin b whose row is >= r[c].
constexpr int kMaxRowChunks = 8;
struct ChunkRows
{
    int r[kMaxRowChunks + 1];
};
// Bins whose kernel fuses runs of twin rows (twin_bin_a / twin_bin_b, -1: none) never get a chunk
// boundary inside a run: the boundary moves forward past the followers (they are then computed one
// chunk EARLIER than their download needs them).
__global__ void k_chunk_bounds(const int *__restrict__ bins, const int *__restrict__ off, int nbins, int nc, ChunkRows rows,
                               int *__restrict__ out, const unsigned char *__restrict__ twins, int twin_bin_a,
                               int twin_bin_b)
{
    pdl_prologue();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nbins * (nc + 1))
        return;
    const int b = t / (nc + 1), c = t % (nc + 1);
    int lo = off[b], hi = off[b + 1];
    const int target = rows.r[c];
    while (lo < hi)
    {
        const int mid = (lo + hi) >> 1;
        if (bins[mid] < target)
            lo = mid + 1;
        else
            hi = mid;
    }
    if (twins && (b == twin_bin_a || b == twin_bin_b))
    {
        const int first = off[b], end = off[b + 1];
        while (lo > first && lo < end && bins[lo] == bins[lo - 1] + 1 && twins[bins[lo]])
            ++lo;
    }
    out[t] = lo;
}

// ---------------------------------------------------------------------------------------
// Stable binning: count per block, scan per bin across blocks on the device, scatter with
// warp-level ranks (match_any + popc) and a warp prefix in shared memory.
// blockhist layout: [bin][block].
// ---------------------------------------------------------------------------------------
constexpr int kBinThreads = 1024;

__global__ void __launch_bounds__(kBinThreads) k_bin_count(int M, const unsigned char *__restrict__ binid,
                                                           int *__restrict__ blockhist, int nblocks)
{
    pdl_prologue();
    __shared__ int hist[MHB_MAX_BINS + 1];
    if (threadIdx.x <= MHB_MAX_BINS)
        hist[threadIdx.x] = 0;
    __syncthreads();
    int i = blockIdx.x * kBinThreads + threadIdx.x;
    int b = (i < M) ? binid[i] : MHB_MAX_BINS;
    unsigned peers = __match_any_sync(kFull, b);
    if ((peers & lanemask_lt()) == 0)
        atomicAdd(&hist[b], __popc(peers));
    __syncthreads();
    if (threadIdx.x < MHB_MAX_BINS)
        blockhist[threadIdx.x * nblocks + blockIdx.x] = hist[threadIdx.x];
}

// One block, one warp per bin: exclusive scan of that bin's per-block counts (warp shuffle
// scan, 32 blocks per step), then the bin bases from the 16 totals; sizes / offsets to scal.
__global__ void __launch_bounds__(32 * MHB_MAX_BINS) k_bin_offsets(int *__restrict__ blockhist, int nblocks,
                                                                   int nbins, int *__restrict__ size_out,
                                                                   int *__restrict__ off_out)
{
    pdl_prologue();
    __shared__ int total[MHB_MAX_BINS], base[MHB_MAX_BINS + 1];
    const int b = threadIdx.x >> 5, lane = lane_id();
    int *h = blockhist + (size_t)b * nblocks;
    int carry = 0;
    if (b < nbins)
        for (int c0 = 0; c0 < nblocks; c0 += 32)
        {
            const int i = c0 + lane;
            const int v = (i < nblocks) ? h[i] : 0;
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const int t = __shfl_up_sync(kFull, incl, o);
                if (lane >= o)
                    incl += t;
            }
            if (i < nblocks)
                h[i] = carry + incl - v; // exclusive inside the bin; the bin base is added below
            carry += __shfl_sync(kFull, incl, 31);
        }
    if (lane == 0)
        total[b] = (b < nbins) ? carry : 0;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        int run = 0;
        for (int q = 0; q < MHB_MAX_BINS; ++q)
        {
            base[q] = run;
            size_out[q] = total[q];
            off_out[q] = run;
            run += total[q];
        }
        base[MHB_MAX_BINS] = run;
        off_out[MHB_MAX_BINS] = run;
    }
    __syncthreads();
    if (b < nbins && base[b] != 0)
        for (int i = lane; i < nblocks; i += 32)
            h[i] += base[b];
}

__global__ void __launch_bounds__(kBinThreads) k_bin_scatter(int M, const unsigned char *__restrict__ binid,
                                                             const int *__restrict__ blockhist, int nblocks,
                                                             int *__restrict__ bins)
{
    pdl_prologue();
    __shared__ int warpcnt[32][MHB_MAX_BINS + 1];
    for (int t = threadIdx.x; t < 32 * (MHB_MAX_BINS + 1); t += kBinThreads)
        (&warpcnt[0][0])[t] = 0;
    __syncthreads();
    int i = blockIdx.x * kBinThreads + threadIdx.x;
    int w = threadIdx.x >> 5;
    int b = (i < M) ? binid[i] : MHB_MAX_BINS;
    unsigned peers = __match_any_sync(kFull, b);
    int rank = __popc(peers & lanemask_lt());
    if (rank == 0)
        warpcnt[w][b] = __popc(peers);
    __syncthreads();
    // exclusive prefix over warps, per bin (column-wise), done by 32 x nbins threads
    if (threadIdx.x < 32 * MHB_MAX_BINS)
    {
        // thread (bin = t / 32, lane = warp index): warp-level scan over the 32 warps
        int bin = threadIdx.x >> 5, wl = threadIdx.x & 31;
        int v = warpcnt[wl][bin], incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            int t = __shfl_up_sync(kFull, incl, o);
            if (wl >= o)
                incl += t;
        }
        warpcnt[wl][bin] = incl - v;
    }
    __syncthreads();
    if (i < M)
        bins[blockhist[b * nblocks + blockIdx.x] + warpcnt[w][b] + rank] = i;
}

} // namespace mhb
