// mhb_stream.cuh -- the two ways a G-lane group walks the intermediate products of one C row.
//
// A group owns row i of C = A*B.  Its work: for every nonzero (i,k) of A, every entry of B's
// row k (CSR arrays for the numeric pass; tileptr/tilecol/tilemask for the symbolic pass).
//
//  * walk_sequential: one B row at a time, G consecutive entries per step.  The entries of
//    one B row are distinct columns, so within a step no two lanes touch the same
//    accumulator slot and the caller's update can be a plain read-modify-write; steps are
//    ordered by __syncwarp.  The first kPre chunks of the NEXT B row are loaded into
//    registers before the current one is consumed (the first ncu capture showed ~50 % of the
//    stalls on the scoreboard of the B loads: L2 latency with 24 warps/SM), and so is the
//    A-side metadata of the next G nonzeros.  Best when B rows are >= G/2 long (FEM-like).
//
//  * walk_flat: load-balanced expansion.  The B-row ranges of G nonzeros of A are laid end
//    to end (warp prefix sum of the lengths) and lane l takes product t0+l, finding its B
//    row by a log2(G)-step binary search over the prefix held in the lanes (shuffles).  All
//    lanes stay busy however short the B rows are (power-law graphs: ~3 entries per row,
//    where the sequential walk keeps 3 of 32 lanes busy), but lanes of one step may now
//    hit the same slot, so the caller's update must be atomic.
//
// All control state is uniform across the group, so the shuffles inside are convergent.
#pragma once
#include <type_traits>

#include "mhb_common.cuh"

namespace mhb
{

struct NoVal
{
};

template <typename TA>
__device__ __forceinline__ TA group_bcast(unsigned gm, TA v, int src, int G)
{
    if constexpr (std::is_same<TA, NoVal>::value)
        return v;
    else
        return __shfl_sync(gm, v, src, G);
}

// Per-lane metadata of nonzero j of A's row: B-row range [ms, me) and A's value.
template <typename TA>
__device__ __forceinline__ void load_meta(int j, int e, const int *__restrict__ Ac, const TA *__restrict__ Av,
                                          const int *__restrict__ Bp, int &ms, int &me, TA &ma)
{
    ms = 0, me = 0;
    if constexpr (!std::is_same<TA, NoVal>::value)
        ma = TA(0);
    if (j < e)
    {
        const int k = __ldg(&Ac[j]);
        if constexpr (!std::is_same<TA, NoVal>::value)
            ma = __ldg(&Av[j]);
        ms = __ldg(&Bp[k]);
        me = __ldg(&Bp[k + 1]);
    }
}

// update(c, v, a): c = column / tile column, v = B payload, a = A value (NoVal for symbolic)
template <int G, int kPre, typename TA, typename TB, class Update>
__device__ __forceinline__ void walk_sequential(unsigned gm, int l, int s, int e, const int *__restrict__ Ac,
                                                const TA *__restrict__ Av, const int *__restrict__ Bp,
                                                const int *__restrict__ Bc, const TB *__restrict__ Bv,
                                                Update update)
{
    int bs, be, nbs, nbe;
    TA av, nav;
    load_meta<TA>(s + l, e, Ac, Av, Bp, bs, be, av);
    for (int j0 = s; j0 < e; j0 += G)
    {
        load_meta<TA>(j0 + G + l, e, Ac, Av, Bp, nbs, nbe, nav);
        const int cnt = min(G, e - j0);
        int pc[kPre], nq = 0, nqe = 0;
        TB pv[kPre];
        TA na = av;
        auto issue = [&](int i) {
            nq = __shfl_sync(gm, bs, i, G);
            nqe = __shfl_sync(gm, be, i, G);
            na = group_bcast<TA>(gm, av, i, G);
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
            TB cv[kPre];
            const int q = nq, qe = nqe;
            const TA a = na;
#pragma unroll
            for (int t = 0; t < kPre; ++t)
                cc[t] = pc[t], cv[t] = pv[t];
            if (i + 1 < cnt)
                issue(i + 1);
#pragma unroll
            for (int t = 0; t < kPre; ++t)
                if (cc[t] >= 0)
                    update(cc[t], cv[t], a);
            for (int p = q + kPre * G + l; p < qe; p += G) // B rows longer than kPre*G
                update(__ldg(&Bc[p]), __ldg(&Bv[p]), a);
            __syncwarp(gm); // order this B row's stores before the next row's loads
        }
        bs = nbs, be = nbe, av = nav;
    }
}

// update(c, v, a) must be atomic with respect to the other lanes of the group.
// Every group walks ALL chunks of G nonzeros of A (j0 = s, s + G, ...); the products of a
// chunk are split between the `tparts` groups sharing the row: this group takes products
// tpart*G + l, (tpart + tparts)*G + l, ...  (tparts = 1 for a group that owns its row; the
// warps of a block that share one row pass their warp index / warp count, so that rows of A
// with few nonzeros still keep every warp of the block busy).
template <int G, typename TA, typename TB, class Update>
__device__ __forceinline__ void walk_flat(unsigned gm, int l, int s, int e, int tpart, int tparts,
                                          const int *__restrict__ Ac, const TA *__restrict__ Av,
                                          const int *__restrict__ Bp, const int *__restrict__ Bc,
                                          const TB *__restrict__ Bv, Update update)
{
    int bs, be, nbs, nbe;
    TA av, nav;
    load_meta<TA>(s + l, e, Ac, Av, Bp, bs, be, av);
    for (int j0 = s; j0 < e; j0 += G)
    {
        load_meta<TA>(j0 + G + l, e, Ac, Av, Bp, nbs, nbe, nav);
        const int len = be - bs;
        int incl = len;
#pragma unroll
        for (int o = 1; o < G; o <<= 1)
        {
            const int t = __shfl_up_sync(gm, incl, o, G);
            if (l >= o)
                incl += t;
        }
        const int off = incl - len;
        const int total = __shfl_sync(gm, incl, G - 1, G);
        const int base = bs - off;
        for (int t0 = tpart * G; t0 < total; t0 += tparts * G)
        {
            const int t = t0 + l;
            int ent = 0;
#pragma unroll
            for (int step = G / 2; step > 0; step >>= 1)
            {
                const int o = __shfl_sync(gm, off, ent + step, G);
                if (o <= t)
                    ent += step;
            }
            const int q = t + __shfl_sync(gm, base, ent, G);
            const TA a = group_bcast<TA>(gm, av, ent, G);
            if (t < total)
                update(__ldg(&Bc[q]), __ldg(&Bv[q]), a);
        }
        bs = nbs, be = nbe, av = nav;
    }
}

} // namespace mhb
