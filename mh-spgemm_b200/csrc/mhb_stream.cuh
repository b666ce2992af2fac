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

// walk_sequential for the numeric pass with TWIN folding.  Multi-dof FEM matrices hold runs
// of B rows with identical column patterns (same[k] = row k repeats row k-1, family 1), and
// the nonzeros of A that select them sit next to each other.  Up to three such nonzeros are
// consumed as one step: the columns are loaded once (from the first twin), the three value
// rows are combined in registers, v = a0*b0 + a1*b1 + a2*b2, and the accumulator sees ONE
// read-modify-write instead of three.  For the cant-like input this removes 2/3 of the
// shared-memory traffic, which is what bounds the kernel (profiles/r1a_numwin_baseline.md:
// LSU data pipe 59 %, 2-way bank conflicts on the fp64 window).
// update(c, v): v already contains the A factor.
template <typename T>
__device__ __forceinline__ int4 pack_meta(int bs, int be, T av)
{
    if constexpr (sizeof(T) == 8)
    {
        const long long bits = __double_as_longlong((double)av);
        return make_int4(bs, be, (int)(bits & 0xffffffffLL), (int)(bits >> 32));
    }
    else
        return make_int4(bs, be, __float_as_int((float)av), 0);
}
template <typename T>
__device__ __forceinline__ T meta_val(const int4 &m)
{
    if constexpr (sizeof(T) == 8)
        return (T)__longlong_as_double(((long long)m.w << 32) | (unsigned)m.z);
    else
        return (T)__int_as_float(m.z);
}

// `stage`: G + 2 int4 of shared memory owned by this group.  The per-nonzero metadata of the
// current chunk of A ({B-row start, end, A value}) is parked there once per chunk and read
// back with one broadcast LDS.128 per twin -- the first version fetched it with 11 shuffles
// per step, and shuffles travel the same data pipe as the accumulator traffic that bounds
// this kernel (profiles/r1b_numwin_twins.md).
template <int G, int kPre, typename T, class Update>
__device__ __forceinline__ void walk_sequential_twins(unsigned gm, int l, int s, int e,
                                                      const int *__restrict__ Ac, const T *__restrict__ Av,
                                                      const int *__restrict__ Bp, const int *__restrict__ Bc,
                                                      const T *__restrict__ Bv,
                                                      const unsigned char *__restrict__ same, int4 *stage,
                                                      Update update)
{
    int bs, be, kk, nbs, nbe, nkk;
    T av, nav;
    const int gbase = lane_id() & ~(G - 1);
    auto meta = [&](int j, int &ms, int &me, int &mk, T &ma) {
        ms = 0, me = 0, mk = -2, ma = T(0);
        if (j < e)
        {
            mk = __ldg(&Ac[j]);
            ma = __ldg(&Av[j]);
            ms = __ldg(&Bp[mk]);
            me = __ldg(&Bp[mk + 1]);
            if (__ldg(&same[mk]))
                mk |= kTwinTag;
        }
    };
    meta(s + l, bs, be, kk, av);
    if (l < 2)
        stage[G + l] = make_int4(0, 0, 0, 0);
    for (int j0 = s; j0 < e; j0 += G)
    {
        meta(j0 + G + l, nbs, nbe, nkk, nav);
        const int cnt = min(G, e - j0);
        stage[l] = pack_meta<T>(bs, be, av);
        // follower = repeats the pattern of the nonzero just before it (inside this chunk)
        const int kprev = __shfl_up_sync(gm, kk & kTwinMask, 1, G);
        const bool fol = l > 0 && l < cnt && (kk & kTwinTag) && (kk & kTwinMask) == kprev + 1;
        const unsigned fmask = (__ballot_sync(gm, fol) >> gbase) & (G == 32 ? 0xffffffffu : ((1u << G) - 1u));
        __syncwarp(gm); // stage[] visible to the group
        int pc[kPre], nq = 0, nqe = 0, nsz = 1, nb1 = 0, nb2 = 0;
        T pv0[kPre], pv1[kPre], pv2[kPre], na0 = T(0), na1 = T(0), na2 = T(0);
        // group starting at entry i: size 1 + (following follower bits, at most 2)
        auto issue = [&](int i) {
            nsz = 1 + ((fmask >> (i + 1)) & 1u);
            if (nsz == 2)
                nsz += (fmask >> (i + 2)) & 1u;
            if (i + nsz > cnt)
                nsz = cnt - i;
            const int4 m0 = stage[i], m1 = stage[i + 1], m2 = stage[i + 2];
            nq = m0.x;
            nqe = m0.y;
            na0 = meta_val<T>(m0);
            // twins: offsets of their value rows relative to the first twin's
            nb1 = m1.x - nq;
            nb2 = m2.x - nq;
            na1 = meta_val<T>(m1);
            na2 = meta_val<T>(m2);
#pragma unroll
            for (int t = 0; t < kPre; ++t)
            {
                const int p = nq + t * G + l;
                pc[t] = -1;
                if (p < nqe)
                {
                    pc[t] = __ldg(&Bc[p]);
                    pv0[t] = __ldg(&Bv[p]);
                    if (nsz > 1)
                        pv1[t] = __ldg(&Bv[p + nb1]);
                    if (nsz > 2)
                        pv2[t] = __ldg(&Bv[p + nb2]);
                }
            }
        };
        issue(0);
        for (int i = 0; i < cnt;)
        {
            int cc[kPre];
            T cv[kPre];
            const int q = nq, qe = nqe, sz = nsz, b1 = nb1, b2 = nb2;
            const T a0 = na0, a1 = na1, a2 = na2;
#pragma unroll
            for (int t = 0; t < kPre; ++t)
            {
                // lanes without an element compute on stale registers; their result is dropped
                cc[t] = pc[t];
                T v = a0 * pv0[t];
                if (sz > 1)
                    v = fma(a1, pv1[t], v);
                if (sz > 2)
                    v = fma(a2, pv2[t], v);
                cv[t] = v;
            }
            i += sz;
            if (i < cnt)
                issue(i);
#pragma unroll
            for (int t = 0; t < kPre; ++t)
                update(cc[t], cv[t], cc[t] >= 0); // predicated inside: no divergent region
            for (int p = q + kPre * G + l; p < qe; p += G) // B rows longer than kPre*G
            {
                T v = a0 * __ldg(&Bv[p]);
                if (sz > 1)
                    v = fma(a1, __ldg(&Bv[p + b1]), v);
                if (sz > 2)
                    v = fma(a2, __ldg(&Bv[p + b2]), v);
                update(__ldg(&Bc[p]), v, true);
            }
            __syncwarp(gm); // order this step's stores before the next step's loads
        }
        bs = nbs, be = nbe, kk = nkk, av = nav;
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

// walk_flat that also tells every product its position in the row's expansion order (products of
// the row's first nonzero of A first, each B row front to back): update(pos, c, v, a).
template <int G, typename TA, typename TB, class Update>
__device__ __forceinline__ void walk_flat_indexed(unsigned gm, int l, int s, int e, int tpart, int tparts,
                                                  const int *__restrict__ Ac, const TA *__restrict__ Av,
                                                  const int *__restrict__ Bp, const int *__restrict__ Bc,
                                                  const TB *__restrict__ Bv, Update update)
{
    int bs, be, nbs, nbe;
    TA av, nav;
    int rowbase = 0;
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
                update(rowbase + t, __ldg(&Bc[q]), __ldg(&Bv[q]), a);
        }
        rowbase += total;
        bs = nbs, be = nbe, av = nav;
    }
}

// walk_flat with convergent hooks: update(c, v, a) is executed by ALL lanes of the group
// (c = -1 for lanes that have no product in this step) and returns an int per product (e.g. a
// newly claimed hash slot, or -1); post(r) follows, also convergent, so both may use
// ballots/shuffles.  The B entries of a step are loaded one step before they are used.
template <int G, typename TA, typename TB, class Update, class Post>
__device__ __forceinline__ void walk_flat_post(unsigned gm, int l, int s, int e, int tpart, int tparts,
                                               const int *__restrict__ Ac, const TA *__restrict__ Av,
                                               const int *__restrict__ Bp, const int *__restrict__ Bc,
                                               const TB *__restrict__ Bv, Update update, Post post)
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
        // item t of the chunk -> (entry of A it belongs to, position in B); one step ahead of its use
        auto fetch = [&](int t, int &c, TB &v, TA &a) {
            int ent = 0;
#pragma unroll
            for (int step = G / 2; step > 0; step >>= 1)
            {
                const int o = __shfl_sync(gm, off, ent + step, G);
                if (o <= t)
                    ent += step;
            }
            const int q = t + __shfl_sync(gm, base, ent, G);
            a = group_bcast<TA>(gm, av, ent, G);
            c = -1;
            if (t < total)
            {
                c = __ldg(&Bc[q]);
                v = __ldg(&Bv[q]);
            }
        };
        int cc = -1, nc = -1;
        TB cv, nv;
        TA ca, na;
        int t0 = tpart * G;
        if (t0 < total)
            fetch(t0 + l, cc, cv, ca);
        for (; t0 < total; t0 += tparts * G)
        {
            nc = -1;
            if (t0 + tparts * G < total)
                fetch(t0 + tparts * G + l, nc, nv, na);
            const int r = update(cc, cv, ca); // ALL lanes; cc = -1 marks a lane without a product
            post(r);
            cc = nc, cv = nv, ca = na;
        }
        bs = nbs, be = nbe, av = nav;
    }
}

} // namespace mhb
