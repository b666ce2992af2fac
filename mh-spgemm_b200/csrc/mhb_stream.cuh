// mhb_stream.cuh -- software-pipelined walk over the intermediate products of one C row.
//
// A G-lane group owns one row i of C = A*B.  Its work is the sequence of "items": for every
// nonzero (i,k) of A, the entries of B's row k in chunks of G consecutive elements (one per
// lane).  The first profile of the numeric kernel (profiles/r1_numwin_baseline.md) showed
// ~50 % of the warp stalls on the scoreboard of the B loads: L2 latency, not bandwidth, with
// only 24 resident warps per SM (shared-memory limited).  ItemStream therefore decouples
// *issuing* the loads of an item from *consuming* it: the kernels keep a ring of D items
// in registers, so D chunks of B (and the A-side metadata of the next 32 nonzeros of A) are
// in flight per group while the accumulator is being updated.
//
// All iterator state is uniform across the group, so the shuffles inside are convergent.
#pragma once
#include <type_traits>

#include "mhb_common.cuh"

namespace mhb
{

struct NoVal
{
};

// G lanes; TA = value type of A (NoVal for the symbolic pass); TB = payload type of B
// (value for numeric, tile mask for symbolic).  Bp/Bc/Bv describe B's rows (CSR arrays for
// numeric; tileptr/tilecol/tilemask for symbolic).
template <int G, typename TA, typename TB>
struct ItemStream
{
    const int *__restrict__ Ac;
    const TA *__restrict__ Av;
    const int *__restrict__ Bp;
    const int *__restrict__ Bc;
    const TB *__restrict__ Bv;
    unsigned gm;
    int l;
    // per-lane staging of up to G nonzeros of A: current chunk and the prefetched next one
    int bs, be, nbs, nbe;
    TA av, nav;
    // uniform iterator state
    int jn, e;   // start of the chunk after `next`, end of A's row
    int cnt, ai; // nonzeros in the current chunk, next one to open
    int ncnt;    // nonzeros in the prefetched chunk
    int q, qe;   // next element / end of the open B row
    TA a;        // A value of the open B row

    __device__ __forceinline__ void load_chunk(int j0, int &s_, int &e_, TA &a_, int &n_)
    {
        n_ = min(G, e - j0);
        s_ = 0;
        e_ = 0;
        if (n_ > 0 && j0 + l < e)
        {
            const int k = __ldg(&Ac[j0 + l]);
            if constexpr (!std::is_same<TA, NoVal>::value)
                a_ = __ldg(&Av[j0 + l]);
            s_ = __ldg(&Bp[k]);
            e_ = __ldg(&Bp[k + 1]);
        }
        if (n_ < 0)
            n_ = 0;
    }

    __device__ __forceinline__ void init(int s, int e_)
    {
        e = e_;
        load_chunk(s, bs, be, av, cnt);
        load_chunk(s + G, nbs, nbe, nav, ncnt);
        jn = s + 2 * G;
        ai = 0;
        q = qe = 0;
    }

    // Issue the loads of the next item.  Returns false (uniformly) when the row is exhausted.
    // c < 0 means this lane has no element in the item.
    __device__ __forceinline__ bool next(int &c, TB &v, TA &a_out)
    {
        while (q >= qe)
        {
            if (ai == cnt)
            {
                if (ncnt == 0)
                    return false;
                bs = nbs, be = nbe, av = nav, cnt = ncnt;
                load_chunk(jn, nbs, nbe, nav, ncnt);
                jn += G;
                ai = 0;
            }
            q = __shfl_sync(gm, bs, ai, G);
            qe = __shfl_sync(gm, be, ai, G);
            if constexpr (!std::is_same<TA, NoVal>::value)
                a = __shfl_sync(gm, av, ai, G);
            ++ai;
        }
        const int p = q + l;
        c = -1;
        if (p < qe)
        {
            c = __ldg(&Bc[p]);
            v = __ldg(&Bv[p]);
        }
        a_out = a;
        q += G;
        return true;
    }
};

} // namespace mhb
