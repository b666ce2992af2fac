// mhb_numeric.cuh -- kernel family 4: the numeric accumulate / compact / sort pass.
//
// Replaces h_numeric (inc/MH_spgemm.cuh:364-430) and its kernels (inc/numeric.cuh):
// column indices (ascending) and values of every row of C = A*B, given C's row offsets.
//
// The reference does, per intermediate product, a 64-bit modulo hash, an ATOMS.CAS on the
// key and an LDS.64 + DADD + ATOMS.CAST.SPIN.64 loop on the value (shared-memory fp64 add
// has no native atomic on sm_100), then compacts through an atomic cursor and rank-sorts
// in O(n^2).  Two accumulators replace that here:
//
//  * DENSE WINDOW: a C row whose column span W fits shared memory (227 KB holds 27 648
//    fp64 columns) accumulates into acc[col - first_col].  No key, no hash, no probing,
//    and the compaction sweep emits the row already column-sorted -- no sort at all.
//  * HASH (multiplicative hash, power-of-two table, linear probing) for rows whose span is
//    large relative to their nnz; compaction by ballot/prefix and a bitonic (or, for tiny
//    rows, rank) sort; a global-memory table for rows above 12 288 nnz.
//
// Group kernels (8 lanes or one warp per row) walk A's row one B row at a time.  The
// columns of one B row are distinct, so in one step no two lanes update the same
// accumulator: the value update is a plain LDS / FMA / STS ordered by __syncwarp, with no
// shared-memory atomic at all.  Block kernels (one block per row, warps on different B
// rows) fall back to atomicAdd.  Loads of B are coalesced along the B row.
#pragma once
#include "mhb_common.cuh"
#include "mhb_stream.cuh"

namespace mhb
{

constexpr int kNumGroupThreads = 256;
constexpr int kNumDepth = 4; // items (chunks of B) in flight per group, see mhb_stream.cuh
constexpr int kPre = 3;      // chunks of the next B row prefetched by the dense-window kernel

// =========================================================================================
// Dense window, G lanes per row.
// =========================================================================================
template <int G, typename T>
__global__ void __launch_bounds__(kNumGroupThreads)
    k_num_win_group(const int *__restrict__ rows, int nrows, const int *__restrict__ Ap,
                    const int *__restrict__ Ac, const T *__restrict__ Av, const int *__restrict__ Bp,
                    const int *__restrict__ Bc, const T *__restrict__ Bv, const int4 *__restrict__ arow,
                    const int *__restrict__ Cp, int *__restrict__ Cc, T *__restrict__ Cv, int wcap)
{
    extern __shared__ __align__(16) unsigned char sm_raw[];
    constexpr int GPB = kNumGroupThreads / G;
    const int g = threadIdx.x / G, l = threadIdx.x % G;
    const unsigned gm = group_mask<G>();
    T *acc = reinterpret_cast<T *>(sm_raw) + (size_t)g * wcap;
    for (int r = blockIdx.x * GPB + g; r < nrows; r += gridDim.x * GPB)
    {
        const int row = rows[r];
        const int4 info = arow[row];
        const int cmin = info.z;
        const int W = info.w - cmin + 1;
        for (int i = l; i < W; i += G)
            acc[i] = Unset<T>::value();
        __syncwarp(gm);
        const int s = __ldg(&Ap[row]), e = __ldg(&Ap[row + 1]);
        // Walk A's row 32 (G) nonzeros at a time.  The first kPre chunks of the B row of
        // nonzero i+1 are loaded into registers before nonzero i is accumulated, so the L2
        // latency of B overlaps the shared-memory updates; the A-side metadata of the next
        // G nonzeros is prefetched the same way.
        int bs = 0, be = 0, nbs = 0, nbe = 0;
        T av = T(0), nav = T(0);
        auto load_meta = [&](int j, int &ms, int &me, T &ma) {
            ms = 0, me = 0, ma = T(0);
            if (j < e)
            {
                const int k = __ldg(&Ac[j]);
                ma = __ldg(&Av[j]);
                ms = __ldg(&Bp[k]);
                me = __ldg(&Bp[k + 1]);
            }
        };
        load_meta(s + l, bs, be, av);
        for (int j0 = s; j0 < e; j0 += G)
        {
            load_meta(j0 + G + l, nbs, nbe, nav);
            const int cnt = min(G, e - j0);
            int pc[kPre], nq = 0, nqe = 0;
            T pv[kPre], na = T(0);
            auto issue = [&](int i) {
                nq = __shfl_sync(gm, bs, i, G);
                nqe = __shfl_sync(gm, be, i, G);
                na = __shfl_sync(gm, av, i, G);
#pragma unroll
                for (int t = 0; t < kPre; ++t)
                {
                    const int p = nq + t * G + l;
                    pc[t] = -1;
                    if (p < nqe)
                    {
                        pc[t] = __ldg(&Bc[p]);
                        pv[t] = __ldg(&Bv[p]);
                    }
                }
            };
            issue(0);
            for (int i = 0; i < cnt; ++i)
            {
                int cc[kPre];
                T cv[kPre];
                const int q = nq, qe = nqe;
                const T a = na;
#pragma unroll
                for (int t = 0; t < kPre; ++t)
                    cc[t] = pc[t], cv[t] = pv[t];
                if (i + 1 < cnt)
                    issue(i + 1);
                // the columns of one B row are distinct: plain read-modify-write, no atomics
#pragma unroll
                for (int t = 0; t < kPre; ++t)
                    if (cc[t] >= 0)
                    {
                        const int idx = cc[t] - cmin;
                        const T o = acc[idx];
                        acc[idx] = Unset<T>::is(o) ? a * cv[t] : fma(a, cv[t], o);
                    }
                for (int p = q + kPre * G + l; p < qe; p += G) // rows longer than kPre*G
                {
                    const int idx = __ldg(&Bc[p]) - cmin;
                    const T v = __ldg(&Bv[p]);
                    const T o = acc[idx];
                    acc[idx] = Unset<T>::is(o) ? a * v : fma(a, v, o);
                }
                __syncwarp(gm); // order this B row's stores before the next row's loads
            }
            bs = nbs, be = nbe, av = nav;
        }
        // ordered compaction: the window is already sorted by column
        int out = __ldg(&Cp[row]);
        for (int i0 = 0; i0 < W; i0 += G)
        {
            const int i = i0 + l;
            const T v = (i < W) ? acc[i] : Unset<T>::value();
            const bool p = !Unset<T>::is(v);
            const unsigned bal = __ballot_sync(gm, p);
            if (p)
            {
                const int pos = out + __popc(bal & lanemask_lt());
                Cc[pos] = cmin + i;
                Cv[pos] = v;
            }
            out += __popc(bal);
        }
        __syncwarp(gm);
    }
}

// =========================================================================================
// Block-wide helpers
// =========================================================================================
// exclusive scan of one int per thread across the block; returns exclusive prefix, *total
__device__ __forceinline__ int block_excl_scan(int v, int *warp_tot /*32*/, int *total)
{
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        int t = __shfl_up_sync(kFull, incl, o);
        if (lane_id() >= o)
            incl += t;
    }
    __syncthreads();
    if (lane_id() == 31)
        warp_tot[threadIdx.x >> 5] = incl;
    __syncthreads();
    const int nw = blockDim.x >> 5;
    int pre = 0, tot = 0;
    for (int w = 0; w < nw; ++w)
    {
        int t = warp_tot[w];
        if (w < (int)(threadIdx.x >> 5))
            pre += t;
        tot += t;
    }
    *total = tot;
    return pre + incl - v;
}

// =========================================================================================
// Dense window, one block per row (W up to 27 648 columns).  Warps work on different B rows
// at once, so values use atomicAdd and presence is a bitmap set with atomicOr.
// smem: acc[wcap] | flags[wcap/32] | wpre[wcap/32]
// =========================================================================================
template <typename T>
__global__ void k_num_win_block(const int *__restrict__ rows, int nrows, const int *__restrict__ Ap,
                                const int *__restrict__ Ac, const T *__restrict__ Av,
                                const int *__restrict__ Bp, const int *__restrict__ Bc,
                                const T *__restrict__ Bv, const int4 *__restrict__ arow,
                                const int *__restrict__ Cp, int *__restrict__ Cc, T *__restrict__ Cv, int wcap)
{
    extern __shared__ __align__(16) unsigned char sm_raw[];
    __shared__ int warp_tot[32];
    T *acc = reinterpret_cast<T *>(sm_raw);
    unsigned *flags = reinterpret_cast<unsigned *>(acc + wcap);
    int *wpre = reinterpret_cast<int *>(flags + wcap / 32);
    const int warp = threadIdx.x >> 5, lane = lane_id(), nwarp = blockDim.x >> 5;
    for (int r = blockIdx.x; r < nrows; r += gridDim.x)
    {
        const int row = rows[r];
        const int4 info = arow[row];
        const int cmin = info.z;
        const int W = info.w - cmin + 1;
        const int nwords = (W + 31) >> 5;
        for (int i = threadIdx.x; i < W; i += blockDim.x)
            acc[i] = T(0);
        for (int i = threadIdx.x; i < nwords; i += blockDim.x)
            flags[i] = 0u;
        __syncthreads();
        const int s = __ldg(&Ap[row]), e = __ldg(&Ap[row + 1]);
        for (int j0 = s + warp * 32; j0 < e; j0 += nwarp * 32)
        {
            int bs = 0, be = 0;
            T av = T(0);
            if (j0 + lane < e)
            {
                const int k = __ldg(&Ac[j0 + lane]);
                av = __ldg(&Av[j0 + lane]);
                bs = __ldg(&Bp[k]);
                be = __ldg(&Bp[k + 1]);
            }
            const int cnt = min(32, e - j0);
            for (int i = 0; i < cnt; ++i)
            {
                const int qs = __shfl_sync(kFull, bs, i), qe = __shfl_sync(kFull, be, i);
                const T a = __shfl_sync(kFull, av, i);
                for (int q = qs + lane; q < qe; q += 32)
                {
                    const int idx = __ldg(&Bc[q]) - cmin;
                    atomicAdd(&acc[idx], a * __ldg(&Bv[q]));
                    const unsigned bit = 1u << (idx & 31);
                    if (!(flags[idx >> 5] & bit))
                        atomicOr(&flags[idx >> 5], bit);
                }
            }
        }
        __syncthreads();
        // prefix of the per-word popcounts -> output position of every present column
        int carry = 0;
        for (int w0 = 0; w0 < nwords; w0 += blockDim.x)
        {
            const int w = w0 + threadIdx.x;
            const int c = (w < nwords) ? __popc(flags[w]) : 0;
            int tot;
            const int ex = block_excl_scan(c, warp_tot, &tot);
            if (w < nwords)
                wpre[w] = carry + ex;
            carry += tot;
        }
        __syncthreads();
        const int out = __ldg(&Cp[row]);
        for (int i = threadIdx.x; i < W; i += blockDim.x)
        {
            const unsigned f = flags[i >> 5];
            if ((f >> (i & 31)) & 1u)
            {
                const int pos = out + wpre[i >> 5] + __popc(f & ((1u << (i & 31)) - 1u));
                Cc[pos] = cmin + i;
                Cv[pos] = acc[i];
            }
        }
        __syncthreads();
    }
}

// =========================================================================================
// Hash accumulation
// =========================================================================================
// Find-or-claim the slot of `key`; returns the slot, or -1 if the table is full.
__device__ __forceinline__ int key_slot(int *keys, int logS, int key)
{
    const unsigned S1 = (1u << logS) - 1u;
    unsigned h = hash_slot((unsigned)key, logS);
    for (unsigned it = 0; it <= S1; ++it)
    {
        int old = keys[h];
        if (old == key)
            return (int)h;
        if (old == -1)
        {
            old = atomicCAS(&keys[h], -1, key);
            if (old == -1 || old == key)
                return (int)h;
        }
        h = (h + 1) & S1;
    }
    return -1;
}

// In-place bitonic sort of (keys[0..P), vals[0..P)) by key, P a power of two, executed by
// `nthr` cooperating threads (tid in [0, nthr)); `sync` orders the stages.
template <typename T, class Sync>
__device__ __forceinline__ void bitonic_sort_kv(int *keys, T *vals, int P, int tid, int nthr, Sync sync)
{
    for (int k = 2; k <= P; k <<= 1)
    {
        for (int j = k >> 1; j > 0; j >>= 1)
        {
            for (int t = tid; t < (P >> 1); t += nthr)
            {
                // t-th compare-exchange pair of this stage: i has bit j clear
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int x = i | j;
                const int a = keys[i], b = keys[x];
                const bool asc = (i & k) == 0;
                if ((a > b) == asc)
                {
                    keys[i] = b;
                    keys[x] = a;
                    const T va = vals[i];
                    vals[i] = vals[x];
                    vals[x] = va;
                }
            }
            sync();
        }
    }
}

// ---- hash, G lanes per row, table of 2^logS (key, value) slots per group -----------------
template <int G, typename T>
__global__ void __launch_bounds__(kNumGroupThreads)
    k_num_hash_group(const int *__restrict__ rows, int nrows, const int *__restrict__ Ap,
                     const int *__restrict__ Ac, const T *__restrict__ Av, const int *__restrict__ Bp,
                     const int *__restrict__ Bc, const T *__restrict__ Bv, const int *__restrict__ Cp,
                     int *__restrict__ Cc, T *__restrict__ Cv, int logS, int *__restrict__ scal)
{
    extern __shared__ __align__(16) unsigned char sm_raw[];
    constexpr int GPB = kNumGroupThreads / G;
    const int g = threadIdx.x / G, l = threadIdx.x % G;
    const unsigned gm = group_mask<G>();
    const int S = 1 << logS;
    // layout: all value tables first (8-byte aligned), then all key tables
    T *vals = reinterpret_cast<T *>(sm_raw) + (size_t)g * S;
    int *keys = reinterpret_cast<int *>(reinterpret_cast<T *>(sm_raw) + (size_t)GPB * S) + (size_t)g * S;
    for (int r = blockIdx.x * GPB + g; r < nrows; r += gridDim.x * GPB)
    {
        const int row = rows[r];
        for (int i = l; i < S; i += G)
        {
            keys[i] = -1;
            vals[i] = T(0);
        }
        __syncwarp(gm);
        const int s = __ldg(&Ap[row]), e = __ldg(&Ap[row + 1]);
        ItemStream<G, T, T> st{Ac, Av, Bp, Bc, Bv, gm, l};
        st.init(s, e);
        int rc[kNumDepth];
        T rv[kNumDepth], ra[kNumDepth];
        bool live[kNumDepth];
#pragma unroll
        for (int d = 0; d < kNumDepth; ++d)
            live[d] = st.next(rc[d], rv[d], ra[d]);
        while (live[0])
        {
#pragma unroll
            for (int d = 0; d < kNumDepth; ++d)
            {
                if (!live[d])
                    break;
                if (rc[d] >= 0)
                {
                    const int h = key_slot(keys, logS, rc[d]);
                    if (h >= 0)
                        vals[h] = fma(ra[d], rv[d], vals[h]); // columns of one item are distinct: no atomic
                    else
                        atomicMax(scal + SC_ERROR, (int)DEVERR_TABLE_FULL);
                }
                __syncwarp(gm);
                live[d] = st.next(rc[d], rv[d], ra[d]);
            }
        }
        // in-place compaction to the front of the table, G slots per step
        int n = 0;
        for (int i0 = 0; i0 < S; i0 += G)
        {
            const int k = keys[i0 + l];
            const T v = vals[i0 + l];
            __syncwarp(gm);
            const bool p = k != -1;
            const unsigned bal = __ballot_sync(gm, p);
            if (p)
            {
                const int pos = n + __popc(bal & lanemask_lt());
                keys[pos] = k;
                vals[pos] = v;
            }
            n += __popc(bal);
            __syncwarp(gm);
        }
        const int out = __ldg(&Cp[row]);
        if (G < 32)
        {
            // tiny rows: rank sort straight into C
            for (int i = l; i < n; i += G)
            {
                const int k = keys[i];
                int rank = 0;
                for (int j = 0; j < n; ++j)
                    rank += keys[j] < k;
                Cc[out + rank] = k;
                Cv[out + rank] = vals[i];
            }
        }
        else
        {
            int P = 2;
            while (P < n)
                P <<= 1;
            for (int i = n + l; i < P; i += G)
                keys[i] = INT_MAX;
            __syncwarp(gm);
            bitonic_sort_kv(keys, vals, P, l, G, [&]() { __syncwarp(gm); });
            for (int i = l; i < n; i += G)
            {
                Cc[out + i] = keys[i];
                Cv[out + i] = vals[i];
            }
        }
        __syncwarp(gm);
    }
}

// ---- hash, one block per row; table in shared memory or in the global pool ---------------
template <typename T>
__global__ void k_num_hash_block(const int *__restrict__ rows, int nrows, const int *__restrict__ Ap,
                                 const int *__restrict__ Ac, const T *__restrict__ Av,
                                 const int *__restrict__ Bp, const int *__restrict__ Bc,
                                 const T *__restrict__ Bv, const int *__restrict__ Cp, int *__restrict__ Cc,
                                 T *__restrict__ Cv, int logS_fixed, unsigned char *__restrict__ pool,
                                 long long pool_slots, int *__restrict__ scal)
{
    extern __shared__ __align__(16) unsigned char sm_raw[];
    __shared__ int warp_tot[32];
    const int warp = threadIdx.x >> 5, lane = lane_id(), nwarp = blockDim.x >> 5;
    for (int r = blockIdx.x; r < nrows; r += gridDim.x)
    {
        const int row = rows[r];
        const int out = __ldg(&Cp[row]);
        const int n_row = __ldg(&Cp[row + 1]) - out;
        int logS = logS_fixed;
        T *vals;
        int *keys;
        if (pool)
        {
            logS = 10;
            while ((1LL << logS) < 2LL * n_row)
                ++logS;
            unsigned char *base = pool + (size_t)blockIdx.x * (size_t)pool_slots * (sizeof(T) + sizeof(int));
            vals = reinterpret_cast<T *>(base);
            keys = reinterpret_cast<int *>(vals + pool_slots);
        }
        else
        {
            vals = reinterpret_cast<T *>(sm_raw);
            keys = reinterpret_cast<int *>(vals + ((size_t)1 << logS));
        }
        const int S = 1 << logS;
        for (int i = threadIdx.x; i < S; i += blockDim.x)
        {
            keys[i] = -1;
            vals[i] = T(0);
        }
        __syncthreads();
        const int s = __ldg(&Ap[row]), e = __ldg(&Ap[row + 1]);
        for (int j0 = s + warp * 32; j0 < e; j0 += nwarp * 32)
        {
            int bs = 0, be = 0;
            T av = T(0);
            if (j0 + lane < e)
            {
                const int k = __ldg(&Ac[j0 + lane]);
                av = __ldg(&Av[j0 + lane]);
                bs = __ldg(&Bp[k]);
                be = __ldg(&Bp[k + 1]);
            }
            const int cnt = min(32, e - j0);
            for (int i = 0; i < cnt; ++i)
            {
                const int qs = __shfl_sync(kFull, bs, i), qe = __shfl_sync(kFull, be, i);
                const T a = __shfl_sync(kFull, av, i);
                for (int q = qs + lane; q < qe; q += 32)
                {
                    const int c = __ldg(&Bc[q]);
                    const T v = __ldg(&Bv[q]);
                    const int h = key_slot(keys, logS, c);
                    if (h >= 0)
                        atomicAdd(&vals[h], a * v);
                    else
                        atomicMax(scal + SC_ERROR, (int)DEVERR_TABLE_FULL);
                }
            }
        }
        __syncthreads();
        // in-place compaction, blockDim slots per step
        int n = 0;
        for (int i0 = 0; i0 < S; i0 += blockDim.x)
        {
            const int i = i0 + threadIdx.x;
            int k = -1;
            T v = T(0);
            if (i < S)
            {
                k = pool ? __ldcg(&keys[i]) : keys[i];
                v = pool ? __ldcg(&vals[i]) : vals[i];
            }
            const bool p = k != -1;
            int tot;
            const int ex = block_excl_scan(p ? 1 : 0, warp_tot, &tot); // has __syncthreads inside
            if (p)
            {
                keys[n + ex] = k;
                vals[n + ex] = v;
            }
            n += tot;
            __syncthreads();
        }
        int P = 2;
        while (P < n)
            P <<= 1;
        for (int i = n + threadIdx.x; i < P; i += blockDim.x)
            keys[i] = INT_MAX;
        __syncthreads();
        bitonic_sort_kv(keys, vals, P, (int)threadIdx.x, (int)blockDim.x, [&]() { __syncthreads(); });
        for (int i = threadIdx.x; i < n; i += blockDim.x)
        {
            Cc[out + i] = keys[i];
            Cv[out + i] = vals[i];
        }
        __syncthreads();
    }
}

} // namespace mhb
