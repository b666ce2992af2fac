// mhb_common.cuh -- device helpers shared by the four kernel families.
#pragma once
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>

#include "mhb_config.h"

namespace mhb
{

constexpr unsigned kFull = 0xffffffffu;

// Twin tag of a B-row index held in a lane register: the SIGN bit, so that every row index of
// the int32 contract (K up to INT_MAX) keeps all of its 31 value bits.
constexpr int kTwinTag = (int)0x80000000u;
constexpr int kTwinMask = 0x7fffffff;

__device__ __forceinline__ unsigned lanemask_lt()
{
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
__device__ __forceinline__ unsigned lanemask_le()
{
    unsigned m;
    asm("mov.u32 %0, %%lanemask_le;" : "=r"(m));
    return m;
}
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// Lanes of the G-wide group this thread belongs to (G = 8 or 32).
template <int G>
__device__ __forceinline__ unsigned group_mask()
{
    if constexpr (G == 32)
        return kFull;
    else
        return ((1u << G) - 1u) << (lane_id() & ~(G - 1));
}

template <int G, typename T>
__device__ __forceinline__ T group_sum(T v, unsigned gm)
{
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1)
        v += __shfl_xor_sync(gm, v, o, G);
    return v;
}
template <int G>
__device__ __forceinline__ int group_min(int v, unsigned gm)
{
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1)
        v = min(v, __shfl_xor_sync(gm, v, o, G));
    return v;
}
template <int G>
__device__ __forceinline__ int group_max(int v, unsigned gm)
{
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1)
        v = max(v, __shfl_xor_sync(gm, v, o, G));
    return v;
}

// Fibonacci hashing into a power-of-two table of 2^logS slots.  The reference hashes with
// (key*107) % prime and cumulative-square probing (inc/common.h:72, inc/numeric.cuh:233),
// which reaches only ~67 % of the slots; multiplicative hashing with
// linear probing over a power-of-two table is a full cycle and needs no integer modulo.
// Programmatic dependent launch (the small kernels of the binning / mask chain): wait for the
// previous kernel of the stream -- its memory is visible after this -- then let the next
// kernel of the chain start scheduling its blocks while this one runs.  A kernel launched the
// ordinary way passes straight through both instructions.
__device__ __forceinline__ void pdl_prologue()
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

__device__ __forceinline__ unsigned hash_slot(unsigned key, int logS)
{
    return (key * 2654435761u) >> (32 - logS);
}

__device__ __forceinline__ int sat_i32(long long v) { return v > INT_MAX ? INT_MAX : (int)v; }

// "Never written" marker of the dense numeric window: a signalling-NaN bit pattern.  Every
// value stored into the window is the result of a multiply or fused multiply-add, and the
// FPU never produces a signalling NaN, so the marker cannot collide with a real value --
// entries whose products cancel to 0.0 (or are NaN/Inf) stay structurally present, as in
// the reference (inc/numeric.cuh:237-241 accumulates without testing for zero).
template <typename T>
struct Unset;
template <>
struct Unset<double>
{
    static __device__ __forceinline__ double value() { return __longlong_as_double(0x7FF0000000000001LL); }
    static __device__ __forceinline__ bool is(double v) { return __double_as_longlong(v) == 0x7FF0000000000001LL; }
};
template <>
struct Unset<float>
{
    static __device__ __forceinline__ float value() { return __int_as_float(0x7F800001); }
    static __device__ __forceinline__ bool is(float v) { return __float_as_int(v) == 0x7F800001; }
};

// The rows of one bin, as a kernel argument.  Normally the host knows the bin's range
// (n >= 0, rows points at the bin's first entry).  In a SPECULATIVE symbolic launch (the
// host has not read the bin sizes of this call yet and sizes its grids from the previous call
// on the handle) n is -1 and the range is read from the device-side offsets: rows = the start of
// the whole bin list, off = &offsets[bin].
struct RowList
{
    const int *rows;
    const int *off;
    int n;
    const int *gate; // optional (fused call): when *gate != 0 the list reads as empty and the kernel stands down
    int span = 1;    // adjacent bins the list covers (the three cost classes of the thread-per-row kernels)
    __device__ __forceinline__ const int *begin() const { return n >= 0 ? rows : rows + off[0]; }
    __device__ __forceinline__ int size() const
    {
        if (gate && *gate)
            return 0;
        return n >= 0 ? n : off[span] - off[0];
    }
};

// bits of scal[SC_GATE]
enum GateBit
{
    GATE_SYM_MISS = 1, // the speculative symbolic launch did not cover this input (see do_symbolic)
    GATE_NUM_MISS = 2, // a numeric bin is populated that the previous call did not launch, or a row outgrew the pool
    GATE_CAPACITY = 4, // nnz(C) exceeds the caller's C.col / C.val capacity (or the int32 contract)
    GATE_ERROR = 8,    // a symbolic kernel raised a device error
};

// error flags raised by kernels (checked by the host at the next synchronisation point)
enum DevError
{
    DEVERR_NONE = 0,
    DEVERR_TABLE_FULL = 1, // a hash table filled up: row metrics and table ladder disagree
};

// Layout of the device scalar block (ints) mirrored to pinned host memory.
enum Scalar
{
    SC_INTPROD_LO = 0, // unsigned long long at [0..1]
    SC_TILEFLOP_LO = 2, // unsigned long long at [2..3]
    SC_NTILES_LO = 4,   // long long at [4..5]
    SC_NNZC_LO = 6,     // long long at [6..7]
    SC_MAX_TILEFLOP = 8,
    SC_MAX_ROWNNZ = 9,
    SC_ERROR = 10,
    SC_SPEC_MISS = 11,     // a speculative launch met a capacity it had not planned for: redo
    SC_PROBES_LO = 12,     // unsigned long long at [12..13]: failed probes of the numeric hash kernels (option count_probes)
    SC_SYM_PROBES_LO = 14, // unsigned long long at [14..15]: same for the symbolic tile hash
    SC_SYM_SIZE = 16,               // MHB_MAX_BINS ints
    SC_SYM_OFF = 40,                // MHB_MAX_BINS + 1 ints
    SC_NUM_SIZE = 72,               // MHB_MAX_BINS ints
    SC_NUM_OFF = 96,                // MHB_MAX_BINS + 1 ints
    // fused call (mhb_spgemm_into_*): reasons why the speculatively launched numeric kernels must
    // stand down (GateBit); written by k_fused_gate between the row-offset scan and the numeric phase
    SC_GATE = 124,
    SC_COUNT = 128
};

// Adds this thread's failed-probe count to the statistics counter (option "count_probes";
// probes == nullptr when the option is off).  The reference's HASH_CONFLICT counter
// (inc/common.h:18, inc/numeric.cuh:116-118) counts the same event: a probe that found the
// slot taken by another key.
__device__ __forceinline__ void flush_probes(unsigned long long *probes, int np)
{
    if (probes && np)
        atomicAdd(probes, (unsigned long long)np);
}

} // namespace mhb
