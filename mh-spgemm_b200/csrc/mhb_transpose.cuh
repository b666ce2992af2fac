// mhb_transpose.cuh -- device-side CSR transpose (T = A^T), the `AAT` mode's B operand.
//
// Replaces matrix_transposition (src/utils.cpp:20-46), which the reference runs on the host
// before CSR::H2D (src/main.cu:98-101).  The host loop visits A's nonzeros in (row, column)
// order and appends each to its column's list, so every row of T holds A's row indices in
// ascending order: T is canonical CSR (what MH_spgemm needs for its B operand).
//
// Here: that is a STABLE sort of the nonzeros by column.  The nonzeros are already ordered by
// (row, column), so a stable least-significant-digit radix sort on the column alone (8 bits
// per pass, ceil(log2 N / 8) passes) leaves equal columns in ascending row order -- no
// atomic cursors, so the result is deterministic and needs no per-row sort afterwards.
//   k_tr_count_cols   column histogram  -> scan -> T.ptr
//   k_radix_count     per-block digit histogram [digit][block]  -> scan (run_scan)
//   k_radix_scatter   stable scatter (match_any ranks inside a warp, warp prefix per digit);
//                     the last pass writes T.col (= A's row, found by binary search in A.ptr)
//                     and T.val instead of the (key, position) pair.
#pragma once
#include "mhb_common.cuh"

namespace mhb
{

constexpr int kRadixBits = 8;
constexpr int kRadixBins = 1 << kRadixBits;
constexpr int kRadixThreads = 512;
constexpr int kRadixItems = 8;
constexpr int kRadixTile = kRadixThreads * kRadixItems;

__global__ void __launch_bounds__(256) k_tr_count_cols(const int *__restrict__ Ac, long long nnz,
                                                       int *__restrict__ counts)
{
    pdl_prologue();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < nnz; j += stride)
        atomicAdd(&counts[__ldg(&Ac[j])], 1);
}

// blockhist[d * nblocks + b] = items of tile b whose digit is d
__global__ void __launch_bounds__(kRadixThreads) k_radix_count(const int *__restrict__ keys, long long n, int shift,
                                                               int *__restrict__ blockhist, int nblocks)
{
    pdl_prologue();
    __shared__ int hist[kRadixBins];
    for (int t = threadIdx.x; t < kRadixBins; t += kRadixThreads)
        hist[t] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * kRadixTile;
#pragma unroll
    for (int it = 0; it < kRadixItems; ++it)
    {
        const long long i = base + it * kRadixThreads + threadIdx.x;
        if (i < n)
            atomicAdd(&hist[(__ldg(&keys[i]) >> shift) & (kRadixBins - 1)], 1);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < kRadixBins; t += kRadixThreads)
        blockhist[(size_t)t * nblocks + blockIdx.x] = hist[t];
}

// row of A that owns nonzero position j: the last r with Ap[r] <= j
__device__ __forceinline__ int row_of_position(const int *__restrict__ Ap, int M, int j)
{
    int lo = 0, hi = M; // invariant: Ap[lo] <= j < Ap[hi]
    while (hi - lo > 1)
    {
        const int mid = (lo + hi) >> 1;
        if (__ldg(&Ap[mid]) <= j)
            lo = mid;
        else
            hi = mid;
    }
    return lo;
}

// Stable scatter of one radix pass.  blockhist holds the exclusive scan of the [digit][block]
// counts.  idx_in == nullptr: the payload is the item's own position (first pass).
// FINAL: write T.col / T.val instead of (keys_out, idx_out).
template <typename T, bool FINAL>
__global__ void __launch_bounds__(kRadixThreads)
    k_radix_scatter(const int *__restrict__ keys_in, const int *__restrict__ idx_in, long long n, int shift,
                    const int *__restrict__ blockhist, int nblocks, int *__restrict__ keys_out,
                    int *__restrict__ idx_out, const int *__restrict__ Ap, int M, const T *__restrict__ Av,
                    int *__restrict__ Tc, T *__restrict__ Tv)
{
    pdl_prologue();
    constexpr int NW = kRadixThreads / 32;
    __shared__ int running[kRadixBins];
    __shared__ int warpcnt[NW][kRadixBins];
    const int w = threadIdx.x >> 5;
    for (int t = threadIdx.x; t < kRadixBins; t += kRadixThreads)
        running[t] = 0;
    const long long base = (long long)blockIdx.x * kRadixTile;
    for (int it = 0; it < kRadixItems; ++it)
    {
        for (int t = threadIdx.x; t < NW * kRadixBins; t += kRadixThreads)
            (&warpcnt[0][0])[t] = 0;
        __syncthreads();
        const long long i = base + it * kRadixThreads + threadIdx.x;
        const bool valid = i < n;
        int key = 0, d = -1;
        if (valid)
        {
            key = __ldg(&keys_in[i]);
            d = (key >> shift) & (kRadixBins - 1);
        }
        const unsigned peers = __match_any_sync(kFull, d);
        const int rank = __popc(peers & lanemask_lt());
        if (valid && rank == 0)
            warpcnt[w][d] = __popc(peers);
        __syncthreads();
        // exclusive prefix over the warps, per digit, continued from the earlier rounds
        if (threadIdx.x < kRadixBins)
        {
            int acc = running[threadIdx.x];
#pragma unroll
            for (int ww = 0; ww < NW; ++ww)
            {
                const int c = warpcnt[ww][threadIdx.x];
                warpcnt[ww][threadIdx.x] = acc;
                acc += c;
            }
            running[threadIdx.x] = acc;
        }
        __syncthreads();
        if (valid)
        {
            const int pos = __ldg(&blockhist[(size_t)d * nblocks + blockIdx.x]) + warpcnt[w][d] + rank;
            const int src = idx_in ? __ldg(&idx_in[i]) : (int)i;
            if (FINAL)
            {
                Tc[pos] = row_of_position(Ap, M, src);
                Tv[pos] = __ldg(&Av[src]);
            }
            else
            {
                keys_out[pos] = key;
                idx_out[pos] = src;
            }
        }
        __syncthreads();
    }
}

} // namespace mhb
